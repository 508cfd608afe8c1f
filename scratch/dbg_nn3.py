import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import erc_b200
from erc_b200 import ops
M, K, N = 3000, 256, 384
g = torch.Generator().manual_seed(M + N)
A, B = torch.randn(M, K, generator=g).cuda(), torch.randn(K, N, generator=g).cuda()
want = A.double() @ B.double()
scale = want.abs().max()
found = 0
for it in range(40):
    out = ops.gemm_nn(A, B)
    ops.gemm_nn(A[: M // 2], B)
    torch.cuda.synchronize()
    err = (out.double() - want).abs() / scale
    if err.max() < 1e-4:
        continue
    for t in range((M + 127) // 128):
        for q in range(4):
            r0 = t * 128 + q * 32
            for nt in range(3):
                blk = err[r0:r0 + 32, nt * 128:(nt + 1) * 128]
                if blk.numel() and blk.max() > 1e-4:
                    rows = slice(r0, min(r0 + 32, M)); cols = slice(nt * 128, (nt + 1) * 128)
                    diff = (out[rows, cols].double() - want[rows, cols])
                    # per-chunk contributions
                    P = torch.stack([A[rows, c * 32:(c + 1) * 32].double() @ B[c * 32:(c + 1) * 32, cols].double() for c in range(K // 32)])
                    X = P.reshape(P.shape[0], -1).t()           # [elements, chunks]
                    coef = torch.linalg.lstsq(X, diff.reshape(-1, 1)).solution.flatten()
                    res = float((X @ coef - diff.reshape(-1)).abs().max() / scale)
                    print("iter", it, "mtile", t, "quad", q, "ntile", nt, "chunk coefficients", [round(float(c), 2) for c in coef], "residual %.1e" % res)
                    found += 1
    if found >= 6:
        break
print("done, found", found)
