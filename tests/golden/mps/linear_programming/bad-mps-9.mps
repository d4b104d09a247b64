* incorrect RHS value
NAME   bad-9
ROWS
 N  COST
 L  ROW1
 L  ROW2
COLUMNS
    VAR1      ROW1      3              ROW2      4
RHS
    RHS1      ROW1      a.xx1
