* has bad value in a column
NAME   bad-5
ROWS
 N  COST
 L  ROW1
 L  ROW2
COLUMNS
    VAR1      ROW2      x.aa1
