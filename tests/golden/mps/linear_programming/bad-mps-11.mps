* unknown section name found
NAME   bad-11
SECTION
