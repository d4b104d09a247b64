* unknown row found in RHS
NAME   bad-10
ROWS
 N  COST
 L  ROW1
 L  ROW2
COLUMNS
    VAR1      ROW1      3              ROW2       4
RHS
    RHS1      ROW3      5.4
