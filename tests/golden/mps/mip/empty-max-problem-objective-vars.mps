NAME
OBJSENSE
 MAX
ROWS
 N  OBJ
COLUMNS
    MARKER    'MARKER'                 'INTORG'
    x1        OBJ       -2
    x2        OBJ       5.5
    MARKER    'MARKER'                 'INTEND'
RHS
RANGES
BOUNDS
 BV bounds    x1
 LI bounds    x2 -2
 UI bounds    x2 2
ENDATA
