NAME
ROWS
  N  OBJ
COLUMNS
     x1        OBJ       -3
RHS
RANGES
BOUNDS
  MI bounds    x1
  UP bounds    x1        2
ENDATA
