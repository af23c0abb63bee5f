int lj_oracle_glue_placeholder(void){return 0;}
