"""(needs a library built with `make -C emotion-recognition-in-conversation_b200/csrc EXTRA=-DERCG_TRACE`)
Pipeline timeline of the NN tensor-core GEMM (ERCG_TC_TRACE=1): per k-chunk clock deltas of CTA 0."""
import os, sys
os.environ["ERCG_TC_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import erc_b200
from erc_b200 import ops, _lib
M = 1 << 20
for K, N in [(1443, 100), (100, 400), (400, 100)]:
    ld = (K + 3) // 4 * 4
    A = torch.randn(M, ld, device="cuda")[:, :K]
    B = torch.randn(K, N, device="cuda")
    for _ in range(3):
        C = ops.gemm_nn(A, B)
    torch.cuda.synchronize()
    buf = np.zeros((5, 160, 4), dtype=np.int64)
    _lib.check(_lib.lib().ercg_gemm_nn_tc_trace(buf.ctypes.data), "trace")
    t0 = buf[buf > 0].min()
    rel = np.where(buf > 0, buf - t0, -1)
    print("== K=%d N=%d   (clocks since first mark; -1 = not recorded)" % (K, N))
    print("chunk | Aprod wait_end | split: A_FULL_end TA_FREE_end done | MMA: start accE_end TAfull_end Bfull_end | Bprod wait_end")
    for n in range(48, 112):
        print("%4d | %8d | %8d %8d %8d | %8d %8d %8d %8d | %8d" % (n, rel[0, n, 1], rel[1, n, 1], rel[1, n, 2], rel[1, n, 3],
              rel[2, n, 0], rel[2, n, 1], rel[2, n, 2], rel[2, n, 3], rel[4, n, 1]))
    print("epilogue groups: wait_start acc_full drained")
    for g in range(10, 30):
        print("  g%3d %8d %8d %8d" % (g, rel[3, g, 0], rel[3, g, 1], rel[3, g, 2]))
    # per-chunk steady-state period
    d = np.diff(rel[2, 48:150, 3])
    print("MMA ready-to-ready period: mean %.0f clk, median %.0f" % (d.mean(), np.median(d)))
