"""Micro-benchmark of the dense-transform kernels on the COGMEN config-5 shapes (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import erc_b200
from erc_b200 import ops

M = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
dev = torch.device("cuda")
PEAK = 6548.2


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


print("NN: C[M,N] = A[M,K] @ B[K,N]")
for K, N in [(1443, 100), (100, 300), (100, 400), (300, 100), (400, 100), (100, 100)]:
    ld = (K + 3) // 4 * 4
    A = torch.randn(M, ld, device=dev)[:, :K]
    B = torch.randn(K, N, device=dev)
    C = torch.empty(M, N, device=dev)
    ms = timeit(lambda: ops.gemm_nn(A, B, out=C))
    gb = 4 * M * (K + N) / 1e9
    print("  K=%4d N=%4d  %.3f ms  %.0f GB/s  %.1f%% of HBM peak" % (K, N, ms, gb / ms * 1e3, 100 * gb / ms * 1e3 / PEAK))
    del A, B, C
print("TN: C[K1,N1] = A[M,K1]^T @ B[M,N1]")
for K1, N1 in [(1443, 100), (100, 300), (100, 400), (100, 100)]:
    ld = (K1 + 3) // 4 * 4
    A = torch.randn(M, ld, device=dev)[:, :K1]
    B = torch.randn(M, N1, device=dev)
    ms = timeit(lambda: ops.gemm_tn(A, B))
    gb = 4 * M * (K1 + N1) / 1e9
    print("  K1=%4d N1=%4d  %.3f ms  %.0f GB/s  %.1f%% of HBM peak" % (K1, N1, ms, gb / ms * 1e3, 100 * gb / ms * 1e3 / PEAK))
    del A, B
