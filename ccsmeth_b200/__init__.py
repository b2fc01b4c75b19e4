"""ccsmeth_b200 -- B200-native (sm_100a) implementation of ccsmeth's per-site methylation-call
inference path: ``call_mods`` with the ``attbigru2s`` model, plus the ``call_freqb`` aggregate model.

Only what the hot path needs lives here: ``csrc/`` (CUDA kernels + the C ABI of
libccsm.so), ``models.py`` / ``call_modifications.py`` / ``call_mods_freq_bam.py`` (host-side mirrors of
the reference interface), ``parallel.py`` (one process per GPU, read-stream sharding, NCCL count
all-reduce) and ``synth.py`` (the synthetic workloads BASELINE.json names).
"""
VERSION = "0.1.0"
