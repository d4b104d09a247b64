NAME
ROWS
 N  OBJ
COLUMNS
    MARKER    'MARKER'                 'INTORG'
    x1        OBJ       -2
    MARKER    'MARKER'                 'INTEND'
RHS
RANGES
BOUNDS
 BV bounds    x1
ENDATA
