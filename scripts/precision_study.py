#!/usr/bin/env python
"""CPU simulation of the tensor-core kernel's arithmetic (operand rounding / split schemes /
activation approximations) against the fp32 reference outputs in tests/golden/att2s_synth.npz.
Dev tool: decides which precision modes can meet the 1e-4 parity bar before writing CUDA.

    python scripts/precision_study.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
G = os.path.join(ROOT, "tests", "golden")


def rnd(x, dt):
    return x.to(dt).to(torch.float32)


def split(x, dt, terms):
    """x ~= hi (+ lo): returns list of fp32 tensors exactly representable in dt."""
    hi = rnd(x, dt)
    if terms == 1:
        return [hi]
    lo = rnd(x - hi, dt)
    return [hi, lo]


def make_mm(dt, a_terms, b_terms, cross):
    """Returns mm(a, wT) emulating sum of tensor-core passes with fp32 accumulate.
    cross: list of (ai, bi) index pairs to include."""
    def mm(a, w_parts):
        a_parts = split(a, dt, a_terms)
        acc = None
        for ai, bi in cross:
            p = a_parts[ai].double() @ w_parts[bi].double()  # exact products, ~fp32 accumulate noise ignored
            acc = p if acc is None else acc + p
        return acc.float()
    return mm


def run(sd, g, dt, a_terms, b_terms, cross, state_quant, act="exact"):
    H, L, NL = 256, 21, 3
    mm = make_mm(dt, a_terms, b_terms, cross)
    if dt is None:
        mm = lambda a, wp: (a.double() @ wp[0].double()).float()
        wsplit = lambda w: [w.t().contiguous()]
        q_state = lambda h: h
    else:
        wsplit = lambda w: [p.t().contiguous() for p in split(w, dt, b_terms)]
        if state_quant:
            q_state = lambda h: sum(split(h, dt, a_terms))
        else:
            q_state = lambda h: h
    if act == "exact":
        sig, tanh = torch.sigmoid, torch.tanh
    else:  # tanh.approx.f32 ~ 2^-11 relative error: emulate by rounding result to 11 bits
        def tanh(x):
            y = torch.tanh(x)
            return y * (1 + (torch.rand_like(y) - 0.5) * 2 ** -10.5)
        sig = lambda x: 0.5 * tanh(0.5 * x) + 0.5
    ctxs = []
    for s, sfx in enumerate(("", "2")):
        x = torch.cat([sd["embed.weight"][g["kmer" + sfx].int().long()], g["ipd" + sfx][:, :, None],
                       g["pw" + sfx][:, :, None], g["kpass" + sfx][:, :, None]], 2)
        h0 = g["h0_f"] if s == 0 else g["h0_r"]
        inp = x
        hn_last = []
        for l in range(NL):
            outs = []
            for d, dsfx in enumerate(("", "_reverse")):
                wih = wsplit(sd[f"rnn.weight_ih_l{l}{dsfx}"])
                whh = wsplit(sd[f"rnn.weight_hh_l{l}{dsfx}"])
                bih, bhh = sd[f"rnn.bias_ih_l{l}{dsfx}"], sd[f"rnn.bias_hh_l{l}{dsfx}"]
                h = q_state(h0[2 * l + d])
                out = torch.empty(inp.shape[0], L, H)
                for t in (range(L - 1, -1, -1) if d else range(L)):
                    gi = mm(inp[:, t], wih) + bih
                    gh = mm(h, whh) + bhh
                    r = sig(gi[:, :H] + gh[:, :H])
                    z = sig(gi[:, H:2 * H] + gh[:, H:2 * H])
                    n = tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
                    h = q_state(n + z * (h - n))
                    out[:, t] = h
                outs.append(out)
                if l == NL - 1:
                    hn_last.append(h)
            inp = torch.cat(outs, 2)
        q = torch.cat(hn_last, 1)
        wa, ua = wsplit(sd["_att3.Wa.weight"]), wsplit(sd["_att3.Ua.weight"])
        va = sd["_att3.va.weight"][0]
        qa = mm(q, wa)
        e = torch.stack([(tanh(qa + mm(inp[:, t], ua)) * va).sum(1) for t in range(L)], 1)
        w = torch.softmax(e, 1)
        ctxs.append((inp * w[:, :, None]).sum(1))
    logits = torch.cat(ctxs, 1) @ sd["fc1.weight"].t() + sd["fc1.bias"]
    return logits, torch.softmax(logits, 1)


def main():
    torch.set_num_threads(8)
    sd = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(G, "ckpt_att2s_v3.npz")).items()}
    z = np.load(os.path.join(G, "att2s_synth.npz"))
    n = 256
    g = {k: torch.from_numpy(z[k][:, :n] if k.startswith("h0") else z[k][:n]) for k in z.files}
    ref = g["probs"]
    X3 = [(0, 0), (0, 1), (1, 0)]
    X4 = X3 + [(1, 1)]
    cfgs = [
        ("fp32 restatement", None, 1, 1, [(0, 0)], False, "exact"),
        ("bf16 x1", torch.bfloat16, 1, 1, [(0, 0)], True, "exact"),
        ("fp16 x1", torch.float16, 1, 1, [(0, 0)], True, "exact"),
        ("fp16 x1 + tanh.approx", torch.float16, 1, 1, [(0, 0)], True, "approx"),
        ("bf16 x1 + tanh.approx", torch.bfloat16, 1, 1, [(0, 0)], True, "approx"),
        ("fp16 A x1, W hi+lo (2 pass)", torch.float16, 1, 2, [(0, 0), (0, 1)], True, "exact"),
        ("bf16 x3, state fp32", torch.bfloat16, 2, 2, X3, False, "exact"),
        ("bf16 x3, state hi+lo", torch.bfloat16, 2, 2, X3, True, "exact"),
        ("fp16 x3, state hi+lo", torch.float16, 2, 2, X3, True, "exact"),
        ("fp16 x3, state hi+lo, tanh.approx", torch.float16, 2, 2, X3, True, "approx"),
    ]
    for name, dt, at, bt, cross, sq, act in cfgs:
        torch.manual_seed(0)
        _, probs = run(sd, g, dt, at, bt, cross, sq, act)
        d = (probs - ref).abs()
        p1 = probs[:, 1] / (probs[:, 0] + probs[:, 1])
        r1 = ref[:, 1] / (ref[:, 0] + ref[:, 1])
        mlb = lambda p: torch.where(p >= 1, torch.tensor(255.), torch.floor(p * 256))
        flips = (mlb(p1) != mlb(r1)).sum().item()
        print("%-38s max|dprob| %.2e  mean %.2e  ML-byte flips %d/%d" % (name, d.max(), d.mean(), flips, n))


if __name__ == "__main__":
    main()
