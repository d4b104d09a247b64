* has a column with less than the expected number of chars
NAME   bad-4
ROWS
 N  COST
 L  ROW1
 L  ROW2
COLUMNS
    VAR1      ROW2
