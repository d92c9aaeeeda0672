/* placeholder: SQP outer-loop oracle (row 8f-1) is added in a later commit */
int oracle_sqp_placeholder(void) { return 0; }
