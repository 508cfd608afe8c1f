import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import erc_b200
from erc_b200 import ops
M = 1 << 18
dev = torch.device("cuda")
for K, N in [(1443, 100), (100, 900)]:
    A = torch.randn(M, (K + 3) // 4 * 4, device=dev)[:, :K]
    B = torch.randn(K, N, device=dev)
    C = torch.empty(M, N, device=dev)
    for _ in range(2):
        ops.gemm_nn(A, B, out=C)
    torch.cuda.synchronize()
