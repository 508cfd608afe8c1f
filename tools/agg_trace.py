"""Pipeline timeline of CTA 0 of the aggregate-first RGCN kernel (library built with -DERCG_TRACE; ERCG_TC_TRACE=4).
    tools/build_variant.sh trace WORK -DERCG_TRACE
    ERCG_TC_TRACE=4 ERCG_LIB_PATH=$PWD/variants/trace.so python tools/agg_trace.py"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import erc_b200  # noqa: E402,F401
from erc_b200 import ops, _lib, synth  # noqa: E402
from erc_b200.graph import build_graph  # noqa: E402
from erc_b200.pyg_nn import RGCNConv  # noqa: E402

dev = torch.device("cuda:0")
lengths = synth.config5_lengths(1 << 20, seed=0)
N = int(lengths.sum())
g = build_graph(lengths, torch.zeros(N, dtype=torch.int64, device=dev), 5, 5, 2)
conv = RGCNConv(100, 100, 8).to(dev)
x = torch.randn(N, 100, device=dev)
ei, et = g.attach(), g.edge_type
for _ in range(3):
    out = conv(x, ei, et)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); out = conv(x, ei, et); e1.record(); torch.cuda.synchronize()
print("forward call: %.3f ms" % e0.elapsed_time(e1))
buf = np.zeros((5, 160, 4), dtype=np.int64)
assert _lib.lib().ercg_gemm_nn_tc_trace(buf.ctypes.data_as(ctypes.c_void_p)) == 0
t0 = buf[buf > 0].min()
r = np.where(buf > 0, buf - t0, -1)
print("chunk | split: top aggregated TA_FREE done | MMA: top ACCempty TAfull Bfull | Bprod: top issue")
for n in range(60, 100):
    print("%4d | %7d %7d %7d %7d | %7d %7d %7d %7d | %7d %7d" % (n, r[1, n, 0], r[1, n, 1], r[1, n, 2], r[1, n, 3], r[2, n, 0], r[2, n, 1],
                                                                   r[2, n, 2], r[2, n, 3], r[4, n, 0], r[4, n, 1]))
print("A producer (stage): top issue")
for n in range(20, 34):
    print("%4d | %7d %7d" % (n, r[0, n, 0], r[0, n, 1]))
