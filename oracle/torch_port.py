"""torch-CPU restatement of the two forwards -- TEST ORACLE / CPU BASELINE, not product code.

The reference owns no arithmetic of its own on this path: ``ModelAttRNN.forward``
(reference ccsmeth/models.py:89-150) is a sequence of PyTorch calls (``nn.Embedding``,
``torch.cat``, ``nn.GRU``, ``nn.Linear``, ``softmax``), and the time goes into ``aten::gru``
(SURVEY.md section 6).  This port issues the same ATen calls on the same shapes with the
same threading, so it is what ``bench.py`` times as the host-core CPU baseline
(``cpu_baseline.kind == "port"``): the reference source cannot travel to the GPU box.

Like the reference (call_modifications.py:170-227) it does NOT wrap the forward in
``torch.no_grad()`` and draws h0 with ``requires_grad=True`` unless h0 is given.
"""
import torch
import torch.nn as nn


class _Att(nn.Module):
    def __init__(self, q, k, h):
        super().__init__()
        self.Wa = nn.Linear(q, h, bias=False)
        self.Ua = nn.Linear(k, h, bias=False)
        self.va = nn.Linear(h, 1, bias=False)

    def forward(self, query, keys):  # reference utils/attention.py:48-70
        e = self.va(torch.tanh(self.Wa(query) + self.Ua(keys))).squeeze(2)
        w = torch.softmax(e, dim=1).unsqueeze(2)
        return torch.matmul(torch.transpose(keys, 1, 2), w).squeeze(2)


class Att2sPort(nn.Module):
    """Same parameter names/shapes as the reference checkpoint."""

    def __init__(self, seq_len=21, num_layers=3, num_classes=2, hidden_size=256, feas=11):
        super().__init__()
        self.seq_len, self.num_layers, self.hidden_size = seq_len, num_layers, hidden_size
        self.embed = nn.Embedding(5, 8)
        self.rnn = nn.GRU(feas, hidden_size, num_layers, dropout=0, batch_first=True, bidirectional=True)
        self._att3 = _Att(2 * hidden_size, 2 * hidden_size, hidden_size)
        self.fc1 = nn.Linear(4 * hidden_size, num_classes)

    def _strand(self, kmer, kpass, ipd, pw, h0):
        L = self.seq_len
        x = torch.cat((self.embed(kmer.int()), ipd.reshape(-1, L, 1).float(), pw.reshape(-1, L, 1).float()), 2)
        x = torch.cat((x, kpass.reshape(-1, L, 1).float()), 2)
        if h0 is None:
            h0 = torch.randn(self.num_layers * 2, x.size(0), self.hidden_size, requires_grad=True)
        out, h_n = self.rnn(x, h0)
        q = h_n.reshape(self.num_layers, 2, -1, self.hidden_size)[-1].transpose(0, 1).reshape(-1, 1, 2 * self.hidden_size)
        return self._att3(q, out)

    def forward(self, kmer, kpass, ipd, pw, kmer2, kpass2, ipd2, pw2, h0_f=None, h0_r=None):
        c1 = self._strand(kmer, kpass, ipd, pw, h0_f)
        c2 = self._strand(kmer2, kpass2, ipd2, pw2, h0_r)
        logits = self.fc1(torch.cat((c1, c2), 1))
        return logits, torch.softmax(logits, 1)


class AggrPort(nn.Module):
    """AggrAttRNN restated (reference ccsmeth/models.py:625-694)."""

    def __init__(self, seq_len=11, hidden_size=32, binsize=20):
        super().__init__()
        self.seq_len, self.hidden_size = seq_len, hidden_size
        self.rnn = nn.GRU(binsize + 1, hidden_size, 1, dropout=0, batch_first=True, bidirectional=True)
        self._att3 = _Att(2 * hidden_size, 2 * hidden_size, hidden_size)
        self.fc1 = nn.Linear(2 * hidden_size, 1)

    def forward(self, offsets, histos, h0=None):
        x = torch.cat((histos.float(), offsets.reshape(-1, self.seq_len, 1).float()), 2)
        if h0 is None:
            h0 = torch.randn(2, x.size(0), self.hidden_size, requires_grad=True)
        out, h_n = self.rnn(x, h0)
        q = h_n.reshape(1, 2, -1, self.hidden_size)[-1].transpose(0, 1).reshape(-1, 1, 2 * self.hidden_size)
        return self.fc1(self._att3(q, out))


def load_numpy_state(module, sd):
    """sd: dict of numpy arrays keyed like the reference checkpoint (optional 'module.' prefix)."""
    t = {(k[7:] if k.startswith("module.") else k): torch.from_numpy(v.copy()) for k, v in sd.items()}
    module.load_state_dict(t)
    module.eval()
    return module
