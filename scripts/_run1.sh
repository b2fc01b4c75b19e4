scripts/ncu_dram.sh fp16c8 c8_h3 CCSM_TC_L2HINT=3 | grep gru
scripts/ncu_dram.sh fp16c8 c8_h4 CCSM_TC_L2HINT=4 | grep gru
scripts/ncu_dram.sh bf16 bf_h3 CCSM_TC_L2HINT=3 | grep gru
CCSM_TC_L2HINT=3 scripts/ab_quick.sh fp16c8 ld | tail -1
CCSM_TC_L2HINT=4 scripts/ab_quick.sh fp16c8 ld | tail -1
CCSM_TC_L2HINT=0 scripts/ab_quick.sh fp16c8 ld | tail -1
CCSM_TC_L2HINT=3 scripts/ab_quick.sh bf16 ld | tail -1
CCSM_TC_L2HINT=0 scripts/ab_quick.sh bf16 ld | tail -1
