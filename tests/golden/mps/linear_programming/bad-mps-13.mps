* columns do not occur contiguously
NAME   good-1
ROWS
 N  COST
 L  ROW1
 L  ROW2
COLUMNS
    VAR1      COST      0.2
    VAR2      COST      0.1
    VAR1      ROW1      3              ROW2      2.7
    VAR2      ROW1      4              ROW2      10.1
RHS
    RHS1      ROW1      5.4            ROW2      4.9
ENDATA
