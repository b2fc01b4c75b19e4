scripts/ab_quick.sh fp16c8 ld ln 2>&1 | tail -2
scripts/ab_quick.sh bf16 ld ln 2>&1 | tail -2
