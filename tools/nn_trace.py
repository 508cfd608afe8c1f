"""Pipeline timeline of CTA 0 of the NN tensor-core GEMM (library built with -DERCG_TRACE; run with ERCG_TC_TRACE=1).

    tools/build_variant.sh trace WORK -DERCG_TRACE
    ERCG_TC_TRACE=1 ERCG_LIB_PATH=$PWD/variants/trace.so python tools/nn_trace.py [K N [first_chunk]]
"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import erc_b200  # noqa: E402,F401
from erc_b200 import ops, _lib  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 1443
N = int(sys.argv[2]) if len(sys.argv) > 2 else 100
first = int(sys.argv[3]) if len(sys.argv) > 3 else 100
M = 1 << 20
dev = torch.device("cuda:0")
A = torch.randn(M, (K + 3) // 4 * 4, device=dev)[:, :K]
B = torch.randn(K, N, device=dev)
for _ in range(3):
    ops.gemm_nn(A, B)
torch.cuda.synchronize()
buf = np.zeros((5, 160, 4), dtype=np.int64)
rc = _lib.lib().ercg_gemm_nn_tc_trace(buf.ctypes.data_as(ctypes.c_void_p))
assert rc == 0, rc
t0 = buf[buf > 0].min()
r = np.where(buf > 0, buf - t0, -1)
print("== NN K=%d N=%d  (clk of CTA 0)" % (K, N))
print("chunk | Aprod: top issue | split: top A_FULL TA_FREE done | MMA: top ACCempty TAfull Bfull | Bprod: top issue")
for n in range(first, min(first + 40, 160)):
    print("%4d | %7d %7d | %7d %7d %7d %7d | %7d %7d %7d %7d | %7d %7d" % (
        n, r[0, n, 0], r[0, n, 1], r[1, n, 0], r[1, n, 1], r[1, n, 2], r[1, n, 3], r[2, n, 0], r[2, n, 1], r[2, n, 2], r[2, n, 3],
        r[4, n, 0], r[4, n, 1]))
d = np.diff(r[2, first:159, 3])
print("MMA ready-to-ready period: mean %.0f clk, median %.0f" % (d.mean(), np.median(d)))
w = lambda a, b: float(np.mean((b - a)[first:159]))
print("mean per chunk: A TMA issue->landed(seen by splitter) %.0f | splitter waits A_FULL %.0f, TA_FREE(+split work) %.0f, store %.0f | "
      "MMA waits ACC_EMPTY %.0f, TA_FULL %.0f, B_FULL %.0f | B issue -> MMA sees it %.0f"
      % (w(r[0, :, 1], r[1, :, 1]), w(r[1, :, 0], r[1, :, 1]), w(r[1, :, 1], r[1, :, 2]), w(r[1, :, 2], r[1, :, 3]),
         w(r[2, :, 0], r[2, :, 1]), w(r[2, :, 1], r[2, :, 2]), w(r[2, :, 2], r[2, :, 3]), w(r[4, :, 1], r[2, :, 3])))
print("epilogue groups (first 12): wait ACC_FULL, got it, drained")
for g in range(12):
    print("  g%d: %7d %7d %7d" % (g, r[3, g, 0], r[3, g, 1], r[3, g, 2]))
