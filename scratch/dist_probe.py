"""Per-rank host enqueue time vs step time for the sharded COGMEN step."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import erc_b200
from erc_b200 import ops, synth
from erc_b200.graph import build_graph, graph_sizes
from erc_b200.track_mm.cogmen import COGMENModule
from erc_b200.dist import shard_dialogues, StatSync, LossSync, GradSync

world, rank, lr = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
total = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
mode = sys.argv[2] if len(sys.argv) > 2 else "full"
lengths_all = synth.config5_lengths(total, seed=0)
total_utts = int(lengths_all.sum())
lengths = lengths_all[shard_dialogues(lengths_all, world)[rank]].contiguous()
N = int(lengths.sum())
x = torch.randn(N, 1444, device=dev)[:, :1443]
spk = torch.zeros(N, dtype=torch.int64, device=dev)
labels = torch.randint(0, 6, (N,), device=dev)
sizes = graph_sizes(lengths, 5, 5)
model = COGMENModule(1443, 100, 17, 2, 6).to(dev)
model.train()
optim = torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=1e-8)
loss_sync = grad_sync = None
if world > 1 and mode != "nosync":
    model.gcn.stat_sync = StatSync(global_count=total_utts)
    loss_sync = LossSync()
    grad_sync = GradSync(model)


def step():
    g = build_graph(lengths, spk, 5, 5, 2, device=dev, sizes=sizes)
    logits, _ = model.forward_packed(x, spk, lengths, graph=g)
    loss = ops.cross_entropy(logits, labels, reduce_sync=loss_sync)
    optim.zero_grad(set_to_none=True)
    loss.backward()
    if grad_sync is not None:
        grad_sync()
    optim.step()


for _ in range(5):
    step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
host = []
t0 = time.perf_counter()
for _ in range(20):
    a = time.perf_counter()
    step()
    host.append(time.perf_counter() - a)
torch.cuda.synchronize()
tot = (time.perf_counter() - t0) / 20 * 1e3
host.sort()
print("rank %d N=%d mode=%s: step %.2f ms, host enqueue median %.2f ms max %.2f ms" % (rank, N, mode, tot, host[10] * 1e3, host[-1] * 1e3), flush=True)
if world > 1:
    dist.destroy_process_group()
