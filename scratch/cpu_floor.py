"""CPU cost of one COGMEN train step (tiny batch, so the GPU is idle): where does the host time go?"""
import sys, os, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import erc_b200
from erc_b200 import ops, synth
from erc_b200.graph import build_graph, graph_sizes
from erc_b200.track_mm.cogmen import COGMENModule
dev = torch.device("cuda")
lengths = synth.config5_lengths(16384, seed=0)
N = int(lengths.sum())
x = torch.randn(N, 1444, device=dev)[:, :1443]
spk = torch.zeros(N, dtype=torch.int64, device=dev)
labels = torch.randint(0, 6, (N,), device=dev)
sizes = graph_sizes(lengths, 5, 5)
model = COGMENModule(1443, 100, 17, 2, 6).to(dev)
model.train()
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
print("ms per step (CPU floor): %.2f" % ((time.perf_counter() - t0) / 50 * 1e3))
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    step()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
