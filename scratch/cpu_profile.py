"""CPU-side cost of one COGMEN train step (python + ctypes + autograd), measured on the GPU box with a SMALL batch so the
GPU is never the bottleneck: wall time per step == host time per step."""
import cProfile, pstats, io, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import erc_b200
from erc_b200 import ops, synth
from erc_b200.graph import build_graph, graph_sizes
from erc_b200.track_mm.cogmen import COGMENModule

dev = torch.device("cuda")
lengths = synth.config5_lengths(4096, seed=0)
N = int(lengths.sum())
x_store = torch.randn(N, 1444, device=dev)
x = x_store[:, :1443]
spk = torch.zeros(N, dtype=torch.int64, device=dev)
labels = torch.randint(0, 6, (N,), device=dev)
sizes = graph_sizes(lengths, 5, 5)
model = COGMENModule(1443, 100, 17, 2, 6).to(dev).train()
optim = torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=1e-8)

def step():
    g = build_graph(lengths, spk, 5, 5, 2, device=dev, sizes=sizes)
    logits, _ = model.forward_packed(x, spk, lengths, graph=g)
    loss = ops.cross_entropy(logits, labels)
    optim.zero_grad(set_to_none=True)
    loss.backward()
    optim.step()
    return loss

for _ in range(5):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50):
    step()
torch.cuda.synchronize()
print("host ms/step (tiny batch, N=%d): %.3f" % (N, (time.perf_counter() - t0) / 50 * 1e3))

def phase(name, fn, n=50):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    print("  %-28s %.3f ms" % (name, (time.perf_counter() - t0) / n * 1e3))

phase("build_graph", lambda: build_graph(lengths, spk, 5, 5, 2, device=dev, sizes=sizes))
g = build_graph(lengths, spk, 5, 5, 2, device=dev, sizes=sizes)
phase("build_graph+census wait", lambda: build_graph(lengths, spk, 5, 5, 2, device=dev, sizes=sizes).relation_slots())
def fwd():
    logits, _ = model.forward_packed(x, spk, lengths, graph=g)
    return ops.cross_entropy(logits, labels)
phase("forward+loss", fwd)
def fb():
    loss = fwd()
    optim.zero_grad(set_to_none=True)
    loss.backward()
phase("forward+loss+backward", fb)
def fbo():
    fb()
    optim.step()
phase("forward+loss+backward+adam", fbo)

pr = cProfile.Profile()
pr.enable()
for _ in range(30):
    step()
torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(45)
print(s.getvalue()[:9000])
