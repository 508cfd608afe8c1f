"""Multi-GPU check of the peer-memory all-reduce (csrc/p2p.cu, erc_b200.p2p.PeerComm) against torch.distributed / NCCL.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/p2p_check.py

Per payload (the four exchanges of the COGMEN step and a few odd sizes, fp32 and fp64): the result must equal, BIT FOR BIT, the
rank-ordered sum of the all-gathered inputs on every rank; then a skewed back-to-back stress run (ranks delayed by different
amounts, payload sizes alternating so that slots and flags are reused in every pattern), a CUDA-graph capture with replays, and
the latency next to NCCL's all-reduce of the same tensor.  Rank 0 prints one JSON line; exit code 1 on any mismatch.
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import erc_b200  # noqa: E402,F401
from erc_b200.p2p import PeerComm  # noqa: E402


def ordered_sum(t, world, group):
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t, group=group)
    acc = parts[0].clone()
    for p in parts[1:]:
        acc += p
    return acc


def timeit(fn, reps=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3     # us


def timeit_graph(fn, calls=20, reps=10):
    """per-call time with the launches replayed from a CUDA graph (no host launch cost): what a captured step sees"""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(calls):
            fn()
    g.replay()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * calls) * 1e3


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    group = dist.new_group(backend="nccl")
    comms = PeerComm.create(group, dev, max_bytes=8 << 20, n=2)
    if comms is None:
        if rank == 0:
            print(json.dumps({"p2p": "unavailable (no peer access / IPC refused)", "world": world}))
        dist.barrier()
        return 0
    comm, comm2 = comms
    gen = torch.Generator(device=dev).manual_seed(100 + rank)
    bad = []
    sizes = [(1, torch.float32), (3, torch.float32), (201, torch.float64), (200, torch.float32), (4097, torch.float32),
             (134913, torch.float32), (144401, torch.float32), (1 << 20, torch.float32), (65537, torch.float64)]
    # ---- exactness
    for n, dt in sizes:
        x = torch.randn(n, device=dev, dtype=dt, generator=gen)
        want = ordered_sum(x, world, group)
        got = comm.all_reduce(x.clone())
        if not torch.equal(got, want):
            bad.append(("exact", n, str(dt), float((got - want).abs().max())))
        y = x.clone()
        dist.all_reduce(y, group=group)
        if not torch.allclose(got, y, rtol=1e-5, atol=1e-5):
            bad.append(("vs nccl", n, str(dt)))
        # unaligned views (odd offset): the element-by-element path
        base = torch.randn(n + 1, device=dev, dtype=dt, generator=gen)
        v = base[1:]
        want = ordered_sum(v.contiguous(), world, group)
        got = comm.all_reduce(v.clone()[:])      # clone is aligned; now the truly offset view, in place
        if not torch.equal(got, want):
            bad.append(("exact (clone of offset view)", n, str(dt)))
        comm.all_reduce(v)
        if not torch.equal(v, want):
            bad.append(("exact (offset view, in place)", n, str(dt)))
    # ---- skewed stress: alternate payloads, delay a different rank every iteration, check every result
    xs = [torch.randn(n, device=dev, dtype=dt, generator=gen) for n, dt in sizes[:7]]
    wants = [ordered_sum(x, world, group) for x in xs]
    torch.cuda.synchronize()
    dist.barrier()
    outs = []
    for it in range(400):
        k = it % len(xs)
        if it % world == rank:
            torch.cuda._sleep(200000 + 50000 * (it % 7))      # ~0.1-0.3 ms: this rank arrives late
        outs.append((k, comm.all_reduce(xs[k].clone())))
        if it % 5 == 0:                                       # the second communicator on its own stream, concurrently
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                o2 = comm2.all_reduce(xs[(k + 1) % len(xs)].clone())
            torch.cuda.current_stream().wait_stream(s)
            outs.append(((k + 1) % len(xs), o2))
    torch.cuda.synchronize()
    nbad = sum(0 if torch.equal(o, wants[k]) else 1 for k, o in outs)
    if nbad:
        bad.append(("stress", nbad, len(outs)))
    # ---- CUDA graph: three collectives captured once, replayed on changing data
    a = torch.zeros(201, device=dev, dtype=torch.float64)
    b = torch.zeros(134913, device=dev)
    c = torch.zeros(200, device=dev)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            comm.all_reduce(a); comm.all_reduce(b); comm.all_reduce(c)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    ra, rb, rc = a.clone(), b.clone(), c.clone()
    with torch.cuda.graph(g):
        ra.copy_(a); rb.copy_(b); rc.copy_(c)
        comm.all_reduce(ra); comm.all_reduce(rb); comm.all_reduce(rc)
    for it in range(30):
        a.fill_(float(rank + it)); b.fill_(float(rank) + 0.5 * it); c.fill_(1.0 + it)
        g.replay()
        torch.cuda.synchronize()
        sa = sum(float(r + it) for r in range(world)); sb = sum(float(r) + 0.5 * it for r in range(world)); sc = world * (1.0 + it)
        if not (bool((ra == sa).all()) and bool((rb == sb).all()) and bool((rc == sc).all())):
            bad.append(("graph replay", it))
            break
    # ---- latency next to NCCL
    lat = {}
    for n, dt in [(201, torch.float64), (200, torch.float32), (134913, torch.float32), (144401, torch.float32), (1 << 20, torch.float32)]:
        x = torch.randn(n, device=dev, dtype=dt, generator=gen)
        lat["%d x %s" % (n, str(dt).replace("torch.", ""))] = {
            "p2p_us": round(timeit(lambda: comm.all_reduce(x)), 2),
            "nccl_us": round(timeit(lambda: dist.all_reduce(x, group=group)), 2),
            "p2p_in_graph_us": round(timeit_graph(lambda: comm.all_reduce(x)), 2),
            "nccl_in_graph_us": round(timeit_graph(lambda: dist.all_reduce(x, group=group)), 2)}
    st = comm.status() | comm2.status()
    flag = torch.tensor([len(bad) + (1 if st else 0)], device=dev)
    dist.all_reduce(flag, group=group)
    if rank == 0:
        print(json.dumps({"world": world, "mismatches_all_ranks": int(flag.item()), "rank0_bad": bad[:10], "status": st,
                          "latency_per_call": lat, "stress_calls": len(outs)}))
    dist.barrier()
    return 1 if int(flag.item()) else 0


if __name__ == "__main__":
    sys.exit(main())
