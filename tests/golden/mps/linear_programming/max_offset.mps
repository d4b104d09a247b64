NAME
OBJSENSE
  MAXIMIZE
ROWS
 N  OBJ
 L  c1
COLUMNS
    x1        c1        3
    x1        OBJ       1
    x2        c1        2
    x2        OBJ       2
    x3        c1        1
    x3        OBJ       3
RHS
    rhs       c1        2
    rhs       OBJ       4
RANGES
BOUNDS
 LO bounds    x1        0
 PL bounds    x1
 LO bounds    x2        0
 PL bounds    x2
 LO bounds    x3        0
 UP bounds    x3        1
ENDATA