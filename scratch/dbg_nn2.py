import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import erc_b200
from erc_b200 import ops
for (M, K, N) in [(1376, 800, 200), (3000, 256, 384)]:
    g = torch.Generator().manual_seed(M + N)
    A, B = torch.randn(M, K, generator=g).cuda(), torch.randn(K, N, generator=g).cuda()
    want = A.double() @ B.double()
    scale = want.abs().max()
    outs = []
    for i in range(25):
        outs.append(ops.gemm_nn(A, B))
        ops.gemm_nn(A[: M // 2], B)
    torch.cuda.synchronize()
    errs = [float((o.double() - want).abs().max() / scale) for o in outs]
    print(M, K, N, ["%.1e" % e for e in errs])
    bad = [i for i, e in enumerate(errs) if e > 1e-4]
    if bad:
        i = bad[0]
        err = (outs[i].double() - want).abs() / scale
        mt, ns = (M + 127) // 128, (N + 31) // 32
        for t in range(mt):
            row = " ".join("%.0e" % float(err[t * 128:(t + 1) * 128, j * 32:(j + 1) * 32].max()) for j in range(ns))
            if float(err[t * 128:(t + 1) * 128].max()) > 1e-4:
                print("   out", i, "mtile", t, row, "quadrants", ["%.0e" % float(err[t * 128 + q * 32:t * 128 + (q + 1) * 32].max()) for q in range(4)])
