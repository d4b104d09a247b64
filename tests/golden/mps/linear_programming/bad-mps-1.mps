* NAME doesn't start on the first column
 NAME bad-1
