NAME          SimpleMIP
ROWS
 N  OBJ
 G  C1
COLUMNS
    MARK0000  'MARKER'                 'INTORG'
    x1        OBJ       1
    x1        C1        1
    x2        OBJ       1
    x2        C1        2
    MARK0001  'MARKER'                 'INTEND'
RHS
    RHS1      C1        3
BOUNDS
 LO BND1      x1        0
 LO BND1      x2        0
ENDATA
