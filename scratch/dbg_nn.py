import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import erc_b200
from erc_b200 import ops
torch.manual_seed(0)
for (M, K, N) in [(3000, 256, 384), (1376, 800, 200), (260, 200, 800), (512, 256, 256), (512, 224, 256), (128, 256, 256), (128, 1024, 100)]:
    A, B = torch.randn(M, K).cuda(), torch.randn(K, N).cuda()
    want = (A.double() @ B.double())
    got = ops.gemm_nn(A, B).double()
    err = (got - want).abs()
    scale = want.abs().max()
    print("M,K,N", M, K, N, "max rel err %.2e" % float(err.max() / scale))
    if err.max() / scale > 1e-4:
        mt, ns = (M + 127) // 128, (N + 31) // 32
        for i in range(min(mt, 4)):
            print("   mtile", i, " ".join("%.0e" % float(err[i * 128:(i + 1) * 128, j * 32:(j + 1) * 32].max() / scale) for j in range(ns)))
        # row quadrants of tile 0
        print("   tile0 row-quadrants", [("%.0e" % float(err[q * 32:(q + 1) * 32].max() / scale)) for q in range(4)])
        # is the wrong block equal to a partial sum (missing k chunks)?
        j = int(err[:128].max(0).values.argmax())
        for kk in range(32, K + 1, 32):
            part = A[:128, :kk].double() @ B[:kk].double()
            d = float((got[:128, j] - part[:, j]).abs().max() / scale)
            if d < 1e-5:
                print("   column", j, "of tile 0 equals the partial sum over k <", kk)
