CCSM_TC_VARIANT=pn timeout 300 python -m pytest tests/test_tc_gpu.py -m gpu -x -q -k "test_tc_matches_reference_synth or test_tc_multi_tile_and_chunks or test_tc_layers_match_oracle" 2>&1 | tail -2
CCSM_TC_VARIANT=oo timeout 300 python -m pytest tests/test_tc_gpu.py -m gpu -x -q -k "test_tc_matches_reference_synth" 2>&1 | tail -1
scripts/ab_quick.sh fp16c8 ln pn 2>&1 | tail -2
scripts/ab_quick.sh bf16 ln pn 2>&1 | tail -2
