* has multiple objectives
NAME   bad-2
ROWS
 N  COST1
 N  COST2
