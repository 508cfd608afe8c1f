import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import erc_b200
from erc_b200 import ops
M = 1 << 18
dev = torch.device("cuda")
for K1, N1 in [(1443, 100), (100, 900)]:
    A = torch.randn(M, (K1 + 3) // 4 * 4, device=dev)[:, :K1]
    B = torch.randn(M, N1, device=dev)
    for _ in range(2):
        ops.gemm_tn(A, B)
    torch.cuda.synchronize()
