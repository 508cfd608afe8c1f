"""The optimizer tail (FlatAdam: ercg_adam_step / ercg_sumsq on one flat buffer) against torch.optim, and the whole COGMEN
train step as a replayed CUDA graph against the same steps run eagerly (cogmen.py:50,179-195,
dagerc.py:39,229-231)."""
import copy

import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _mlp(seed):
    torch.manual_seed(seed)
    return torch.nn.Sequential(torch.nn.Linear(37, 64), torch.nn.Tanh(), torch.nn.Linear(64, 5), torch.nn.Linear(5, 3, bias=False)).cuda()


@pytest.mark.parametrize("decoupled,wd,max_norm", [(False, 1e-8, None), (False, 0.0, None), (True, 1e-2, None), (True, 1e-2, 5.0),
                                                   (True, 1e-2, 0.05)])
def test_flat_adam_matches_torch_optim(decoupled, wd, max_norm):
    import erc_b200  # noqa: F401
    from erc_b200.optim import FlatAdam
    a, b = _mlp(0), _mlp(0)
    dead = torch.nn.Parameter(torch.ones(7, device="cuda"))            # never gets a gradient, like the reference's dead encoder
    ours = FlatAdam(list(a.parameters()) + [dead], lr=3e-3, weight_decay=wd, decoupled=decoupled, max_norm=max_norm)
    cls = torch.optim.AdamW if decoupled else torch.optim.Adam
    ref = cls(list(b.parameters()), lr=3e-3, weight_decay=wd, foreach=False, fused=False)
    gen = torch.Generator(device="cuda").manual_seed(1)
    for it in range(12):
        x = torch.randn(50, 37, device="cuda", generator=gen) * (10.0 if it == 5 else 1.0)   # one step with a large gradient norm
        for m in (a, b):
            for p in m.parameters():
                p.grad = None
            m(x).square().sum().backward()
        if max_norm is not None:
            want_norm = torch.nn.utils.clip_grad_norm_(b.parameters(), max_norm)
        ours.step()
        if max_norm is not None:
            assert rel_err(ours.grad_norm().reshape(()), want_norm) < 1e-5
        ref.step()
    assert int(ours.step_dev.item()) == 12 and dead.grad is None and torch.equal(dead.data, torch.ones(7, device="cuda"))
    for (k, p), q in zip(a.named_parameters(), b.parameters()):
        assert rel_err(p, q) < 2e-6, k
        assert p.data_ptr() >= ours.flat_p.data_ptr() and p.data_ptr() < ours.flat_p.data_ptr() + 4 * ours.n   # lives in the flat buffer


def _cogmen_setup(total, seed, p_drop):
    import erc_b200  # noqa: F401
    from erc_b200 import synth
    from erc_b200.track_mm.cogmen import COGMENModule
    lengths = synth.config5_lengths(total, seed=seed)
    b = synth.packed_batch(lengths, 1443, 2, 6, torch.Generator().manual_seed(seed), one_speaker=True)
    torch.manual_seed(seed)
    m = COGMENModule(1443, 100, 17, 2, 6, build_dead_encoder=False).cuda()
    m.cls[2].p = p_drop
    m.train()
    x = b["x_storage"].cuda()[:, :1443]
    return m, lengths, x, b["speaker_packed"].cuda(), b["label"].cuda()


def test_train_step_graph_replay_equals_eager_and_torch_adam():
    from erc_b200 import ops
    from erc_b200.train_step import CogmenTrainStep
    m0, lengths, x, spk, y = _cogmen_setup(6000, 3, 0.0)
    m1, m2 = copy.deepcopy(m0), copy.deepcopy(m0)
    # (a) eager steps through the train-step object
    ts0 = CogmenTrainStep(m0, lengths, (0,))
    losses0 = [float(ts0.step(x, spk, y)) for _ in range(5)]
    ts0.check()
    # (b) the same five steps: two eager warm-up steps inside capture(), then three replays of the captured graph
    ts1 = CogmenTrainStep(m1, lengths, (0,)).capture(x, spk, y, warmup=2)
    losses1 = [float(ts1.replay()) for _ in range(3)]
    assert losses1 == losses0[2:]                                      # same kernels, same order: bit-identical
    for (k, p), q in zip(m0.named_parameters(), m1.parameters()):
        assert torch.equal(p, q), k
    assert int(ts1.opt.step_dev.item()) == 5
    # (c) the reference formulation: ops.cross_entropy + torch.optim.Adam (cogmen.py:50,185-189) on the legacy module path
    opt = torch.optim.Adam(m2.parameters(), lr=1e-4, weight_decay=1e-8, foreach=False, fused=False)
    for i in range(5):
        logits, _ = m2.forward_packed(x, spk, lengths)
        loss = ops.cross_entropy(logits, y)
        opt.zero_grad()
        loss.backward()
        opt.step()
        assert abs(float(loss) - losses0[i]) <= 2e-6 * abs(losses0[i])
    # Same optimizer arithmetic (test_flat_adam_matches_torch_optim feeds both the SAME gradients: 2e-6), but here the two
    # paths compute their gradients with different kernels (window vs generic graph kernels, fused tails): 1e-6 of gradient
    # noise, and Adam's update lr * m / (sqrt(v) + eps) is sign-like where |g| ~ eps.  So: every loss agrees to 2e-6 (above),
    # and the parameters stay within a small fraction of the distance 5 steps can travel (5 * lr), tightly on average.
    travel = 5 * 1e-4
    for (k, p), q in zip(m0.named_parameters(), m2.parameters()):
        d = (p - q).abs()
        assert float(d.max()) <= 0.5 * travel and float(d.mean()) <= 1e-2 * travel, (k, float(d.max()), float(d.mean()))
    assert losses0[-1] < losses0[0]


def test_train_step_graph_replay_draws_a_fresh_dropout_mask_each_replay():
    from erc_b200.train_step import CogmenTrainStep
    m, lengths, x, spk, y = _cogmen_setup(3000, 4, 0.5)
    ts = CogmenTrainStep(m, lengths, (0,), lr=0.0, weight_decay=0.0).capture(x, spk, y)       # lr 0: only the mask changes
    losses = [float(ts.replay()) for _ in range(4)]
    assert len(set(losses)) == 4 and all(l == l for l in losses)


def _single_rank_group():
    import torch.distributed as dist
    if not dist.is_initialized():
        import socket
        sk = socket.socket()
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
        sk.close()
        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=0, world_size=1)
    return dist.group.WORLD


def test_peer_allreduce_single_rank_slots_flags_and_graph_replay():
    """ercg_p2p_allreduce with a one-rank communicator (what a 1-GPU box can run): region set-up through CUDA IPC handles, the
    staging slots / flag words / call counter over many calls of changing size (aligned, odd counts, offset views, fp64), a
    second communicator on a side stream, and replays from a CUDA graph.  With one rank the sum is the input, bit for bit.
    The multi-GPU behaviour is covered by tools/p2p_check.py (test below when the box has two GPUs)."""
    import erc_b200
    from erc_b200.p2p import PeerComm, Reducer
    group = _single_rank_group()
    comms = PeerComm.create(group, "cuda:0", max_bytes=1 << 20, n=2)
    assert comms is not None and len(comms) == 2
    comm, comm2 = comms
    gen = torch.Generator(device="cuda").manual_seed(4)
    before = __import__("erc_b200")._lib.launch_count()
    for it in range(40):
        for n, dt in ((1, torch.float32), (202, torch.float64), (201, torch.float64), (200, torch.float32), (4097, torch.float32),
                      (134913, torch.float32), (1 << 18, torch.float32)):
            x = torch.randn(n + 1, device="cuda", dtype=dt, generator=gen)
            for v in (x[:n], x[1:]):                                  # 16-byte aligned and offset views
                want = v.clone()
                assert comm.all_reduce(v) is v and torch.equal(v, want)
        if it % 4 == 0:
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            y = torch.randn(5000, device="cuda", generator=gen)
            want = y.clone()
            with torch.cuda.stream(s):
                comm2.all_reduce(y)
            torch.cuda.current_stream().wait_stream(s)
            assert torch.equal(y, want)
    assert __import__("erc_b200")._lib.launch_count() - before >= 40 * 14
    with pytest.raises(Exception):
        comm.all_reduce(torch.zeros((1 << 20) // 4 + 64, device="cuda"))   # payload larger than the region's slots
    a = torch.zeros(1000, device="cuda")
    out = torch.zeros(1000, device="cuda")
    red = Reducer(group, comm)
    assert "peer memory" in red.transport
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        red(out)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out.copy_(a)
        red(out)
    for it in range(5):
        a.fill_(float(it))
        g.replay()
        torch.cuda.synchronize()
        assert bool((out == float(it)).all())
    assert comm.status() == 0 and comm2.status() == 0
    for c in comms:
        c.close()


@pytest.mark.parametrize("world", [2, 3, 4, 5, 8])
def test_peer_allreduce_every_world_size_on_one_gpu(world):
    """The flag / slot protocol and all three reduce-loop shapes (W <= 2, W <= 4, W > 4) with W endpoints inside one process:
    rank r's kernel runs on its own stream, the W kernels meet through the flag words exactly as across GPUs
    (PeerComm.simulate).  Result on every rank = the rank-ordered sum, bit for bit; alternating payloads reuse slots and
    flags in every pattern; a late rank (a sleep kernel ahead of its collective) must be waited for."""
    import erc_b200
    from erc_b200.p2p import PeerComm
    comms = PeerComm.simulate(world, "cuda:0", max_bytes=2 << 20)
    streams = [torch.cuda.Stream() for _ in range(world)]
    gen = torch.Generator(device="cuda").manual_seed(world)
    sizes = [(202, torch.float64), (200, torch.float32), (3, torch.float32), (134913, torch.float32), (4097, torch.float64),
             (144401, torch.float32), (1 << 19, torch.float32)]
    for it in range(6):
        for n, dt in sizes:
            xs = [torch.randn(n, device="cuda", dtype=dt, generator=gen) for _ in range(world)]
            want = xs[0].clone()
            for x in xs[1:]:
                want += x
            torch.cuda.synchronize()
            for r in range(world):
                with torch.cuda.stream(streams[r]):
                    if r == (it + n) % world:
                        torch.cuda._sleep(300000)                      # this rank arrives ~0.15 ms late
                    comms[r].all_reduce(xs[r])
            torch.cuda.synchronize()
            for r in range(world):
                assert torch.equal(xs[r], want), (world, it, n, dt, r)
    assert all(c.status() == 0 for c in comms)
    for c in comms:
        c.close()


@pytest.mark.parametrize("world", [2, 3, 8])
def test_fused_bn_statistics_exchange_equals_the_unfused_path(world):
    """ercg_p2p_bn_stats (column reduction + exchange over peer memory + global statistics + running statistics in one
    kernel) on W simulated ranks with shards of different sizes: identical on every rank, bit-identical to bn_stats ->
    pack -> rank-ordered sum -> unpack -> running update, equal to the statistics of the concatenated batch, and the running
    statistics equal nn.BatchNorm1d's on that batch."""
    import erc_b200
    from erc_b200 import ops
    from erc_b200._lib import lib, check
    from erc_b200.p2p import PeerComm
    H = 100
    comms = PeerComm.simulate(world, "cuda:0", max_bytes=1 << 16)
    streams = [torch.cuda.Stream() for _ in range(world)]
    gen = torch.Generator(device="cuda").manual_seed(11 * world)
    ref = torch.nn.BatchNorm1d(H).double().train()
    bns = [torch.nn.BatchNorm1d(H).cuda().train() for _ in range(world)]
    bn_unfused = torch.nn.BatchNorm1d(H).cuda().train()
    for it in range(3):
        shards = [torch.randn(2000 + 977 * r + 131 * it, H, device="cuda", generator=gen) * (1 + r) + 0.3 * r for r in range(world)]
        total = sum(x.size(0) for x in shards)
        torch.cuda.synchronize()
        got = [None] * world
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                got[r] = ops.bn_stats_sync(shards[r], comms[r], total, bns[r])
        torch.cuda.synchronize()
        # unfused: local statistics, pack, sum in rank order, unpack, running update
        acc = None
        for r in range(world):
            m, v = ops.bn_stats(shards[r])
            buf = torch.empty(2 * H + 2, dtype=torch.float64, device="cuda")
            check(lib().ercg_bn_sync_pack(m.data_ptr(), v.data_ptr(), float(shards[r].size(0)), H, buf.data_ptr(), None), "pack")
            acc = buf if acc is None else acc + buf
        gm, gv = torch.empty(H, device="cuda"), torch.empty(H, device="cuda")
        check(lib().ercg_bn_sync_unpack(acc.data_ptr(), H, gm.data_ptr(), gv.data_ptr(), None), "unpack")
        ops.bn_running_update(bn_unfused, gm, gv, float(total))
        allx = torch.cat(shards).double().cpu()
        ref(allx)
        for r in range(world):
            assert torch.equal(got[r][0], gm) and torch.equal(got[r][1], gv), (world, it, r)
            assert torch.equal(bns[r].running_mean, bn_unfused.running_mean) and torch.equal(bns[r].running_var, bn_unfused.running_var)
            assert int(bns[r].num_batches_tracked) == it + 1
        assert rel_err(gm, allx.mean(0)) < 1e-5 and rel_err(gv, allx.var(0, unbiased=False)) < 1e-5
        assert rel_err(bns[0].running_mean, ref.running_mean) < 1e-5 and rel_err(bns[0].running_var, ref.running_var) < 1e-5
    assert all(c.status() == 0 for c in comms)
    for c in comms:
        c.close()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_fused_bn_backward_sums_exchange_equals_the_unfused_path(world):
    """BatchNorm + LeakyReLU backward through ops.bn_leaky_relu with a StatSync whose Reducer owns a peer-memory communicator
    (ercg_p2p_bn_act_bwd_reduce: reduction + exchange in one kernel) on W simulated ranks: the input gradients and the
    dgamma / dbeta of every rank are bit-identical to the unfused path (local sums, rank-ordered total), and their totals
    equal the fp64 gradients of the concatenated batch."""
    import erc_b200
    from erc_b200 import ops
    from erc_b200.dist import StatSync
    from erc_b200.p2p import PeerComm, Reducer
    H = 100
    comms = PeerComm.simulate(world, "cuda:0", max_bytes=1 << 16)
    streams = [torch.cuda.Stream() for _ in range(world)]
    gen = torch.Generator(device="cuda").manual_seed(7 * world)
    shards = [torch.randn(1500 + 611 * r, H, device="cuda", generator=gen) * (1 + 0.5 * r) + 0.2 * r for r in range(world)]
    douts = [torch.randn(x.size(0), H, device="cuda", generator=gen) for x in shards]
    gamma = torch.rand(H, device="cuda", generator=gen) + 0.5
    beta = torch.randn(H, device="cuda", generator=gen)
    total = sum(x.size(0) for x in shards)
    allx = torch.cat(shards).double().cpu().requires_grad_()
    gd, bd = gamma.double().cpu().requires_grad_(), beta.double().cpu().requires_grad_()
    mean64, var64 = allx.mean(0), allx.var(0, unbiased=False)
    want = torch.nn.functional.leaky_relu((allx - mean64) / (var64 + 1e-5).sqrt() * gd + bd, 0.01)
    want.backward(torch.cat(douts).double().cpu())
    mean, var = mean64.detach().float().cuda(), var64.detach().float().cuda()

    class OrderedSum:                      # stands in for the all-reduce of the unfused path: rank-ordered total, computed up front
        def __init__(self, tot):
            self.tot = tot

        def grads(self, sums):
            return self.tot.clone()

    def run(r, sync):
        x, g, b = shards[r].clone().requires_grad_(), gamma.clone().requires_grad_(), beta.clone().requires_grad_()
        out = ops.bn_leaky_relu(x, g, b, mean, var, 1e-5, 0.01, True, float(total), sync)
        out.backward(douts[r])
        return x.grad, g.grad, b.grad

    # unfused reference: every rank's local sums first (dgamma | dbeta with a no-op sync), then the rank-ordered total
    local = [run(r, None) for r in range(world)]
    tot = None
    for r in range(world):
        s_r = torch.cat([local[r][2], local[r][1]])           # sums = (sum dy | sum dy*xhat) = (dbeta | dgamma)
        tot = s_r.clone() if tot is None else tot + s_r
    unfused = [run(r, OrderedSum(tot).grads) for r in range(world)]
    torch.cuda.synchronize()
    fused = [None] * world
    for r in range(world):
        with torch.cuda.stream(streams[r]):
            fused[r] = run(r, StatSync(global_count=total, reduce=Reducer(None, comms[r])).grads)
    torch.cuda.synchronize()
    for r in range(world):
        for a, b in zip(fused[r], unfused[r]):
            assert torch.equal(a, b), (world, r)
    dx = torch.cat([f[0] for f in fused]).double().cpu()
    assert rel_err(dx, allx.grad) < 5e-5
    assert rel_err(sum(f[1] for f in fused), gd.grad, floor=1e-3) < 1e-5 and rel_err(sum(f[2] for f in fused), bd.grad) < 1e-5
    assert all(c.status() == 0 for c in comms)
    for c in comms:
        c.close()


def test_peer_allreduce_missing_peer_times_out_and_reports():
    """The failure path: a peer that never arrives.  With the wait bound lowered to 300 ms (ERCG_P2P_TIMEOUT_MS, read once
    per process -- hence a subprocess) rank 0 of a two-endpoint communicator runs its collective ALONE: the kernel must
    return (no hang), the communicator's sticky status must say ERCG_P2P_ETIMEOUT, check() must raise, and the GPU must
    still run kernels afterwards."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r"""
import sys, time, torch
sys.path.insert(0, %r)
import erc_b200
from erc_b200.p2p import PeerComm
from erc_b200._lib import ErcgError
comms = PeerComm.simulate(2, "cuda:0", max_bytes=1 << 16)
x = torch.ones(1000, device="cuda")
t0 = time.time()
comms[0].all_reduce(x)                  # rank 1 never calls
torch.cuda.synchronize()
dt = time.time() - t0
st = comms[0].status()
try:
    comms[0].check()
    raised = False
except ErcgError:
    raised = True
y = (torch.arange(10, device="cuda") * 2).sum().item()      # the device is alive
print("RESULT", st, raised, round(dt, 2), y)
""" % root
    env = dict(os.environ, ERCG_P2P_TIMEOUT_MS="300")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT")][-1].split()
    assert int(line[1]) == -6 and line[2] == "True" and 0.25 < float(line[3]) < 20.0 and int(line[4]) == 90, r.stdout


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_peer_allreduce_two_gpus_equals_rank_ordered_sum():
    """tools/p2p_check.py on two GPUs: bit-exact against the rank-ordered sum, skewed stress, graph replays, no time-outs."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29547", os.path.join(root, "tools", "p2p_check.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
    assert line.get("mismatches_all_ranks") == 0 and line.get("status") == 0, line
