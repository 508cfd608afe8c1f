import torch, sys
sys.path.insert(0, "/root/repo")
import erc_b200
from erc_b200 import ops
torch.manual_seed(0)
for K in (32, 128, 512, 1443, 4096):
    M, N = 2048, 100
    ld = (K + 3)//4*4
    A = torch.randn(M, ld)[:, :K]; B = torch.randn(K, N)
    want = A.double() @ B.double()
    ops.GEMM_ENGINE = "tc"; got = ops.gemm_nn(A.cuda().contiguous() if K % 4 == 0 else torch.nn.functional.pad(A, (0, ld-K)).cuda()[:, :K], B.cuda()).double().cpu()
    ops.GEMM_ENGINE = "simt"; g2 = ops.gemm_nn(A.cuda(), B.cuda()).double().cpu()
    big = want.abs() > want.abs().max() * 0.2
    for name, g in (("tc", got), ("simt", g2)):
        e = (g - want)
        print(K, name, "max rel(maxnorm) %.2e" % float(e.abs().max() / want.abs().max()),
              "mean signed shrink (|g|-|w|)/|w| on big entries %.2e" % float(((g.abs() - want.abs()) / want.abs())[big].mean()))
# all-positive data: bias shows up fully
K=1443; M=2048; N=100
A = torch.rand(M, 1444)[:, :K]; B = torch.rand(K, N)
want = A.double() @ B.double()
ops.GEMM_ENGINE = "tc"; got = ops.gemm_nn(A.cuda(), B.cuda()).double().cpu()
print("positive data: mean rel signed err %.2e  max %.2e" % (float(((got-want)/want).mean()), float(((got-want)/want).abs().max())))
