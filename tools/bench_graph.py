"""Micro-benchmark of the graph kernels of the COGMEN step (K1, K3 gather fwd/bwd, K4 edge attention fwd/bwd, the BN /
classifier tail) on the BASELINE config-5 graph, through the same autograd ops the model uses.

    python tools/bench_graph.py [--utts 1048576] [--reps 10]
    ERCG_LIB_PATH=variants/<name>.so python tools/bench_graph.py       # A/B a variant build of the library

One line per C-ABI entry point: average launch time over ``reps`` (CUDA events around every call, `_lib.KernelTimer`;
the operands of every kernel are > 400 MB, i.e. larger than L2) and, per op, an int32 bit checksum of the outputs and
gradients so that a variant that changes a single bit is seen in the same run.  Diagnostics only.
"""
import argparse
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import erc_b200  # noqa: E402,F401
from erc_b200 import _lib, graph as G, ops, synth  # noqa: E402


def bits(t):
    return int(t.contiguous().view(torch.int32).to(torch.int64).sum().item()) & 0xFFFFFFFFFFFF


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--utts", type=int, default=1 << 20)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--H", type=int, default=100)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    H = args.H
    lengths = synth.config5_lengths(args.utts)
    N = int(lengths.sum())
    spk = torch.zeros(N, dtype=torch.int64)
    gen = torch.Generator(device=dev).manual_seed(3)
    ids = G.relation_ids_for_speakers([0], 2)

    def build():
        return G.build_graph(lengths, spk, 5, 5, 2, device=dev, relation_ids=ids)

    g = build()
    rel = g.relation_slots()
    P = len(rel[0])
    qkvs = torch.randn(N, 4 * H, device=dev, generator=gen).requires_grad_()
    Y = torch.randn(N, (P + 1) * H, device=dev, generator=gen).requires_grad_()
    bias = torch.randn(H, device=dev, generator=gen).requires_grad_()
    dout = torch.randn(N, H, device=dev, generator=gen)
    sums = {}

    def step(record):
        gg = build()
        o1 = ops.edge_attention(qkvs, gg, H, 1.0 / math.sqrt(H))
        (dq,) = torch.autograd.grad(o1, qkvs, dout)
        o2 = ops.gather(Y, gg, H, 8, w=gg.mean_weight(), bias=bias, root_off=P * H, rel_slot=rel[1], n_slots=P)
        dY, db = torch.autograd.grad(o2, (Y, bias), dout)
        if record:
            sums.update(attn_out=bits(o1), attn_dqkvs=bits(dq), gather_out=bits(o2), gather_dY=bits(dY), gather_db=bits(db),
                        col=bits(gg.col), t_eid=bits(gg.t_eid), inv_cnt=bits(gg.inv_cnt))

    step(True)
    torch.cuda.synchronize()
    with _lib.KernelTimer() as kt:
        for _ in range(args.reps):
            step(False)
    res = {k: round(t / c, 4) for k, (c, t) in sorted(kt.summary().items(), key=lambda kv: -kv[1][1])}
    for k, v in res.items():
        print("%-32s %.4f ms" % (k, v), flush=True)
    print(json.dumps({"lib": os.environ.get("ERCG_LIB_PATH", "in-tree"), "utterances": N, "edges": g.E, "avg_ms": res,
                      "bit_checksums": sums}))


if __name__ == "__main__":
    main()
