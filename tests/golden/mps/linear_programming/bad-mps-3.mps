* has duplicate rows
NAME   bad-3
ROWS
 N  COST
 L  ROW1
 L  ROW1
