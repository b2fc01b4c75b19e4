"""CPU oracle for the ccsmeth per-site methylation-call inference path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker (or, for the
CPU baseline, as the thing timed *beside* the CUDA path) -- never as a fallback.

Contents
  * ``att2s_numpy``  -- numpy restatement of ``ModelAttRNN(attbigru2s).forward``
                        (reference ``ccsmeth/models.py:89-150``,
                        ``ccsmeth/utils/attention.py:48-70``).
  * ``aggr_numpy``   -- numpy restatement of ``AggrAttRNN.forward``
                        (reference ``ccsmeth/models.py:673-694``).
  * ``torch_port``   -- the same two forwards restated with torch CPU ops
                        (``aten::gru`` etc., which is where the reference's own
                        arithmetic lives); this is the "port" that the CPU
                        baseline times on the host cores.
  * ``extract_numpy`` -- numpy restatement of the per-read feature extraction
                        (reference ``ccsmeth/extract_features.py:181-199,261-406``), the checker
                        for the device extractor (``ccsm_reads_*``).
  * ``refimport``    -- imports the *unmodified* reference from /root/reference
                        with stub modules for its absent I/O deps; used only by
                        ``scripts/gen_golden.py`` in the build container.

Parity pinning: the reference ships no tests / golden vectors for this path
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference
itself, generated in the build container by ``scripts/gen_golden.py`` and
committed under ``tests/golden/``.
"""
