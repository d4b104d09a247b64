NAME          EXAMPLE
ROWS
 N  OBJ
 E  C1
COLUMNS
    X1        C1           1
    X1        OBJ          0
    X2        OBJ         -1
RHS
    RHS1      C1           0 
BOUNDS
 LO BND1      X1           0 
 UP BND1      X1           1       
 LO BND1      X2           0 
 UP BND1      X2           1
ENDATA
