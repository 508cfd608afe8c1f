"""Is the tensor-core GEMM power/clock limited?  Samples nvidia-smi while looping one GEMM shape."""
import sys, os, subprocess, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import erc_b200
from erc_b200 import ops

M = 1 << 20
dev = torch.device("cuda")
samples = []
stop = False


def sampler():
    p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown,temperature.gpu",
                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
    for line in p.stdout:
        samples.append(line.strip())
        if stop:
            break
    p.terminate()


for K, N in [(1443, 100), (100, 900)]:
    A = torch.randn(M, (K + 3) // 4 * 4, device=dev)[:, :K]
    B = torch.randn(K, N, device=dev)
    C = torch.empty(M, N, device=dev)
    for _ in range(5):
        ops.gemm_nn(A, B, out=C)
    torch.cuda.synchronize()
    samples.clear()
    stop = False
    th = threading.Thread(target=sampler, daemon=True)
    th.start()
    time.sleep(0.3)
    n0 = len(samples)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(600):
        ops.gemm_nn(A, B, out=C)
    e1.record()
    torch.cuda.synchronize()
    n1 = len(samples)
    stop = True
    print("K=%d N=%d: %.3f ms/iter; idle samples %s; under load: %s" % (K, N, e0.elapsed_time(e1) / 600, samples[max(0, n0 - 2):n0], samples[n0 + 2:n1][::4]))
    time.sleep(0.3)
