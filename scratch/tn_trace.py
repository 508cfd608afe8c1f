# needs a library built with: make -C emotion-recognition-in-conversation_b200/csrc clean all EXTRA=-DERCG_TRACE
import os, sys
os.environ["ERCG_TC_TRACE"] = "2"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import erc_b200
from erc_b200 import ops, _lib
M = 1 << 20
for K1, N1 in [(1443, 100), (100, 400)]:
    ld = (K1 + 3) // 4 * 4
    A = torch.randn(M, ld, device="cuda")[:, :K1]
    B = torch.randn(M, N1, device="cuda")
    for _ in range(3):
        C = ops.gemm_tn(A, B)
    torch.cuda.synchronize()
    buf = np.zeros((5, 160, 4), dtype=np.int64)
    _lib.check(_lib.lib().ercg_gemm_nn_tc_trace(buf.ctypes.data), "trace")
    t0 = buf[buf > 0].min()
    rel = np.where(buf > 0, buf - t0, -1)
    print("== TN K1=%d N1=%d" % (K1, N1))
    print("chunk | Wprod | split: W_FULL TA_FREE done | MMA: start TAfull Bsplit | narrow split (epi): B_FULL done | Nprod")
    for n in range(60, 110):
        print("%4d | %8d | %8d %8d %8d | %8d %8d %8d | %8d %8d | %8d" % (n, rel[0, n, 1], rel[1, n, 1], rel[1, n, 2], rel[1, n, 3],
              rel[2, n, 0], rel[2, n, 2], rel[2, n, 3], rel[3, n, 1], rel[3, n, 2], rel[4, n, 1]))
    d = np.diff(rel[2, 60:150, 3])
    print("MMA ready-to-ready period: mean %.0f clk, median %.0f" % (d.mean(), np.median(d)))
