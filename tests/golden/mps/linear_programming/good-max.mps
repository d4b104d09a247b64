NAME          MAXEX
OBJSENSE
 MAX
ROWS
 N  OBJ
 L  LIM1
COLUMNS
    X1    OBJ     2    LIM1    1
    X2    OBJ     3    LIM1    2
    X3    OBJ     1    LIM1    1
RHS
    RHS1  LIM1    10
BOUNDS
 LO BND1  X1      0
 UP BND1  X1      4
 LO BND1  X2      0
 UP BND1  X2      6
 LO BND1  X3      0
ENDATA