NAME          minrep_inf
ROWS
 N  Obj
 L  c0
 L  c1
 L  c2
 L  c3
 G  c4
 G  c5
COLUMNS
    var0       c0                              1
    var0       c1                              1
    var0       c2                              1
    var0       c3                              1
    var0       Obj                            -2
    var1       c4                             -1
    var1       Obj                            -10
    var2       c4                             0.0210643
    var2       c5                             -260
    var3       c0                             -1
    var3       c1                             -1
    var3       c2                             -1
    var3       c3                             -1
    var3       c4                             0.978936
    var3       c5                             260
    var3       Obj                            12
RHS
    RHS       c5                            260
BOUNDS
 LO BOUND     var0                             0
 UP BOUND     var0                          inf
 LO BOUND     var1                             0
 UP BOUND     var1                          inf
 LO BOUND     var2                             0
 UP BOUND     var2                          inf
 LO BOUND     var3                             0
 UP BOUND     var3                          inf
ENDATA
