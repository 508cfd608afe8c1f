"""Micro-benchmark of the tensor-core GEMMs on the shapes of the bench step (BASELINE config 5), through the C ABI.

    python tools/bench_gemm.py [--rows 1048576] [--reps 10] [--only tn|nn]
    ERCG_LIB_PATH=/path/to/variant.so python tools/bench_gemm.py        # A/B a variant build of the library

One line per shape: average launch time (CUDA events on the launching stream, L2 flushed by the 6 GB operand or by a
256 MB write between launches), algorithmic GB/s, and a checksum-of-checksums error against fp64
(colsum(A @ B) = colsum(A) @ B,  (A^T B) 1 = A^T (B 1)) so that a fast-but-wrong variant is caught in the same run.
Diagnostics only: nothing here is imported by the product or by bench.py.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import erc_b200  # noqa: E402,F401
from erc_b200 import ops  # noqa: E402

NN_SHAPES = [(1443, 100), (100, 300), (100, 400), (300, 100), (400, 100), (100, 100)]
TN_SHAPES = [(1443, 100), (100, 300), (100, 400), (100, 100)]


def timeit(fn, reps, flush):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return sum(ts) / len(ts), ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1 << 20)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    M = args.rows
    g = torch.Generator(device=dev).manual_seed(1)
    flush = torch.empty(64 << 20, device=dev)
    out = []
    if args.only in ("", "nn"):
        for K, N in NN_SHAPES:
            ld = (K + 3) // 4 * 4
            A = torch.randn(M, ld, device=dev, generator=g)[:, :K]
            B = torch.randn(K, N, device=dev, generator=g) / K ** 0.5
            C = ops.gemm_nn(A, B)
            ref = A.double().sum(0) @ B.double()
            err = float((C.double().sum(0) - ref).abs().max() / ref.abs().max())
            avg, best = timeit(lambda: ops.gemm_nn(A, B), args.reps, flush)
            out.append({"kernel": "nn", "K": K, "N": N, "avg_ms": round(avg, 4), "min_ms": round(best, 4),
                        "gbs": round(4.0 * M * (K + N) / avg / 1e6, 1), "checksum_rel_err": err})
            print(json.dumps(out[-1]), flush=True)
            del A, B, C
    if args.only in ("", "step"):
        # the NN calls of the COGMEN step with their epilogues: forward transforms with bias / ReLU+dropout, input
        # gradients with the fused bias-gradient column sums
        cases = [("qkvs fwd", 100, 400, dict(bias=True)), ("rgcn fwd", 100, 300, dict()), ("cls0 fwd", 100, 100, dict(bias=True, act=ops.ACT_RELU_DROPOUT, drop_p=0.5)),
                 ("qkvs dx", 400, 100, dict(want_colsum=True)), ("rgcn dx", 300, 100, dict(want_colsum=True)),
                 ("cls0 dx", 100, 100, dict(want_colsum=True)), ("proj fwd", 1443, 100, dict(bias=True))]
        for name, K, N, kw in cases:
            ld = (K + 3) // 4 * 4
            A = torch.randn(M, ld, device=dev, generator=g)[:, :K]
            B = torch.randn(K, N, device=dev, generator=g) / K ** 0.5
            bias = torch.randn(N, device=dev, generator=g) if kw.get("bias") else None
            call = lambda: ops.gemm_nn(A, B, bias, act=kw.get("act", ops.ACT_NONE), drop_p=kw.get("drop_p", 0.0), seed=1,
                                       want_colsum=kw.get("want_colsum", False))
            avg, best = timeit(call, args.reps, flush)
            out.append({"kernel": "nn-step", "call": name, "K": K, "N": N, "avg_ms": round(avg, 4), "min_ms": round(best, 4)})
            print(json.dumps(out[-1]), flush=True)
            del A, B
    if args.only in ("", "tn"):
        for K1, N1 in TN_SHAPES:
            ld = (K1 + 3) // 4 * 4
            A = torch.randn(M, ld, device=dev, generator=g)[:, :K1]
            B = torch.randn(M, N1, device=dev, generator=g)
            C = ops.gemm_tn(A, B)
            ref = A.double().t() @ B.double().sum(1)
            err = float((C.double().sum(1) - ref).abs().max() / ref.abs().max())
            # and one exact column against fp64
            col = A.double().t() @ B[:, N1 // 2].double()
            err_col = float((C[:, N1 // 2].double() - col).abs().max() / col.abs().max())
            avg, best = timeit(lambda: ops.gemm_tn(A, B), args.reps, flush)
            out.append({"kernel": "tn", "K1": K1, "N1": N1, "avg_ms": round(avg, 4), "min_ms": round(best, 4),
                        "gbs": round(4.0 * M * (K1 + N1) / avg / 1e6, 1), "checksum_rel_err": err, "column_rel_err": err_col})
            print(json.dumps(out[-1]), flush=True)
            del A, B, C
    print(json.dumps({"lib": os.environ.get("ERCG_LIB_PATH", "in-tree"), "rows": M, "results": out}))


if __name__ == "__main__":
    main()
