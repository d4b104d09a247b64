* no objective found
NAME   bad-11
ROWS
 L  ROW1
 L  ROW2
COLUMNS
    VAR1      ROW1      3              ROW2      4
RHS
    RHS1      ROW1      5.4
