* objective found in RHS
NAME   good-mps-rhs-cost
ROWS
 N  COST
 L  ROW1
 L  ROW2
COLUMNS
    VAR1      ROW1      3              ROW2      4
RHS
    RHS1      COST      5
