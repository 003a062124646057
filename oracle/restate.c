int ccu_restate_placeholder(void){return 0;}
