"""Host logic of the batch loop that needs no GPU: feature batching layout and the h0 stream order."""
import numpy as np
import torch

from ccsmeth_b200 import call_modifications as cm
from tests.test_batchloop_gpu import _feature_list


def test_batch_feature_layout_matches_reference_contract(golden_batchloop):
    g = {k: v[:40] if v.ndim == 2 else v for k, v in golden_batchloop.items()}
    fb = cm._batch_feature_list2s(_feature_list(g))
    assert len(fb) == 18
    sampleinfo, fkmers, fpasss, fipdms, fipdsds = fb[0], fb[1], fb[2], fb[3], fb[4]
    assert sampleinfo[3] == "\t".join([".", "-1", ".", "hole0", "24"])
    assert np.array_equal(np.array(fkmers), g["kmer"].astype(np.int64))
    assert np.array_equal(np.array(fpasss), g["kpass"].astype(np.int64))
    assert np.allclose(np.array(fipdms), g["ipd"])
    assert fipdsds[0] == 0  # unused slots are the scalar 0 (reference call_modifications.py:106-110)
    assert np.array_equal(np.array(fb[9]), g["kmer2"].astype(np.int64))


def test_kmer_codes_collapse_iupac():
    assert list(cm._kmer_codes("ACGTNRYKM")) == [0, 1, 2, 3, 4, 4, 4, 4, 4]


def test_h0_stream_is_chunk_ordered_like_reference():
    torch.manual_seed(1234)
    a, b = cm.draw_h0_stream(1100, 512, 3, 256)
    torch.manual_seed(1234)
    for s, e in ((0, 512), (512, 1024), (1024, 1100)):
        f = torch.randn(6, e - s, 256)
        r = torch.randn(6, e - s, 256)
        assert torch.equal(a[:, s:e], f) and torch.equal(b[:, s:e], r)
