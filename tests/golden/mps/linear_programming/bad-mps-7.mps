* incorrect RHS entry
NAME   bad-7
ROWS
 N  COST
 L  ROW1
 L  ROW2
COLUMNS
    VAR1      ROW1      3              ROW2      4
RHS
    RHS1      ROW1
