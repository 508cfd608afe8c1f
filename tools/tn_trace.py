"""Pipeline timeline of CTA 0 of the TN tensor-core GEMM (library built with -DERCG_TRACE; run with ERCG_TC_TRACE=2).

    tools/build_variant.sh trace WORK -DERCG_TRACE
    ERCG_TC_TRACE=2 ERCG_LIB_PATH=$PWD/variants/trace.so python tools/tn_trace.py [K1 N1 [first_chunk]]

Columns are clock64() of CTA 0 relative to its first mark; diagnostics only."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import erc_b200  # noqa: E402,F401
from erc_b200 import ops, _lib  # noqa: E402

K1 = int(sys.argv[1]) if len(sys.argv) > 1 else 1443
N1 = int(sys.argv[2]) if len(sys.argv) > 2 else 100
first = int(sys.argv[3]) if len(sys.argv) > 3 else 100
M = 1 << 20
dev = torch.device("cuda:0")
A = torch.randn(M, (K1 + 3) // 4 * 4, device=dev)[:, :K1]
B = torch.randn(M, N1, device=dev)
for _ in range(3):
    ops.gemm_tn(A, B)
torch.cuda.synchronize()
buf = np.zeros((5, 160, 4), dtype=np.int64)
rc = _lib.lib().ercg_gemm_nn_tc_trace(buf.ctypes.data_as(ctypes.c_void_p))
assert rc == 0, rc
t0 = buf[buf > 0].min()
r = np.where(buf > 0, buf - t0, -1)
print("== TN K1=%d N1=%d  (clk of CTA 0)" % (K1, N1))
print("chunk | Wprod | split: top W_FULL TA_FREE done | MMA: top ACCempty TAfull Bsplit | conv: top B_FULL done [drain_end] | Nprod")
for n in range(first, min(first + 48, 160)):
    print("%4d | %7d | %7d %7d %7d %7d | %7d %7d %7d %7d | %7d %7d %7d %7d | %7d" % (
        n, r[0, n, 1], r[1, n, 0], r[1, n, 1], r[1, n, 2], r[1, n, 3], r[2, n, 0], r[2, n, 1], r[2, n, 2], r[2, n, 3],
        r[3, n, 0], r[3, n, 1], r[3, n, 2], r[3, n, 3], r[4, n, 1]))
m = r[2, first:159, 3]
d = np.diff(m)
print("MMA ready-to-ready period: mean %.0f clk, median %.0f" % (d.mean(), np.median(d)))
# where each role waits (mean clk per chunk)
w = lambda a, b: float(np.mean((b - a)[first:159]))
print("mean waits per chunk: splitter W_FULL %.0f, TA_FREE %.0f | MMA ACC_EMPTY %.0f, TA_FULL %.0f, B_SPLIT %.0f | conv B_FULL %.0f, work %.0f"
      % (w(r[1, :, 0], r[1, :, 1]), w(r[1, :, 1], r[1, :, 2]), w(r[2, :, 0], r[2, :, 1]), w(r[2, :, 1], r[2, :, 2]),
         w(r[2, :, 2], r[2, :, 3]), w(r[3, :, 0], r[3, :, 1]), w(r[3, :, 1], r[3, :, 2])))
