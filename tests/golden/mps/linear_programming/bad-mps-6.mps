* unknown row found in a column
NAME   bad-6
ROWS
 N  COST
 L  ROW1
 L  ROW2
COLUMNS
    VAR1      ROW3      3
