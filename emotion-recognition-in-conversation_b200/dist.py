"""Dialogue-sharded data parallelism.

The reference reaches NCCL only through HF accelerate -> DDP (lumo/trainer/trainer.py:60-67,315-327): a
bucketed gradient all-reduce in ``accelerate.backward`` (track_mm/cogmen.py:188).  Dialogues are independent
graphs (edges never cross dialogues, cogmen_utils.py:125-126), so the path shards by whole dialogues with
no data-path exchange; what crosses GPUs per step is
  1. ONE all-reduce (sum) of the flat buffer of LIVE gradients (COGMEN: 279 304 floats = 1.1 MB),
  2. BatchNorm statistics: (sum x, sum x^2, count) = 2H+1 floats forward, (sum dy, sum dy*xhat) = 2H floats
     backward, so that N-GPU results equal the 1-GPU result on the same global batch (the reference's DDP
     uses per-rank statistics; ``global_stats=False`` reproduces that),
  3. the loss numerator/denominator (2 floats) so the mean is over the global utterance count.
On one NVSwitch node they go through libercgraph's own peer-memory kernels (p2p.py / csrc/p2p.cu: a one-shot all-reduce over
CUDA-IPC regions; the BatchNorm statistics and their backward sums are exchanged by the very kernels that reduce them,
``StatSync.fused_stats`` / ``ops._BnAct.backward``); torch.distributed (NCCL on GPUs, gloo in the CPU tests) is the fallback
transport and the set-up channel.
"""
import torch
import torch.distributed as dist


def shard_dialogues(lengths, world_size):
    """Greedy balance of whole dialogues by utterance count: longest first onto the lightest rank.

    Returns a list (one entry per rank) of int64 index tensors into ``lengths`` (ascending, so every rank keeps
    the original dialogue order).  Deterministic."""
    lengths = torch.as_tensor(lengths, dtype=torch.int64)
    if world_size == 1:
        return [torch.arange(lengths.numel())]
    order = torch.argsort(lengths, descending=True, stable=True).tolist()
    loads = [0] * world_size
    buckets = [[] for _ in range(world_size)]
    # block-greedy: equal-length dialogues are dealt round-robin, which is what greedy does anyway and is O(B)
    ll = lengths.tolist()
    import heapq
    heap = [(0, r) for r in range(world_size)]
    heapq.heapify(heap)
    for d in order:
        load, r = heapq.heappop(heap)
        buckets[r].append(d)
        heapq.heappush(heap, (load + ll[d], r))
        loads[r] = load + ll[d]
    return [torch.tensor(sorted(b), dtype=torch.int64) for b in buckets]


class StatSync:
    """All-reduce hooks for BatchNorm statistics (see GNN._bn_relu in track_mm/cogmen.py of this package)."""

    def __init__(self, group=None, global_count=None, reduce=None):
        self.group = group
        self.global_count = global_count       # host-known global utterance count: avoids a device sync
        # in-place all-reduce(sum): torch.distributed on ``group``, or a p2p.Reducer (peer-memory kernel on one NVSwitch node)
        self.reduce = reduce if reduce is not None else (lambda t: dist.all_reduce(t, group=self.group))

    def fused_stats(self, x, bn=None):
        """(global mean, global biased var, global count) of the rows of x over all ranks in ONE collective kernel
        (ops.bn_stats_sync), running statistics of ``bn`` included -- or None when there is no peer-memory communicator
        or the global count is not known on the host (the caller then goes bn_stats -> stats())."""
        comm = getattr(self.reduce, "comm", None)
        if comm is None or not x.is_cuda or not self.global_count:
            return None
        from .ops import bn_stats_sync
        mean, var = bn_stats_sync(x, comm, self.global_count, bn)
        return mean, var, float(self.global_count)

    def stats(self, mean, var, n_local):
        """local (mean, biased var, n) -> global (mean, biased var, count)."""
        H = mean.numel()
        if mean.is_cuda:
            # two launches around the collective (ercg_bn_sync_pack / _unpack) instead of a dozen elementwise ones: with the
            # batch sharded over 8 GPUs the step is 1.7 ms and every 3 us launch on its critical path shows
            from ._lib import lib, check
            from .ops import _p, _stream
            buf = torch.empty(2 * H + 2, dtype=torch.float64, device=mean.device)   # even count: whole 16-byte units for the peer kernel
            mean, var = mean.contiguous(), var.contiguous()
            check(lib().ercg_bn_sync_pack(_p(mean), _p(var), float(n_local), H, _p(buf), _stream()), "ercg_bn_sync_pack")
            self.reduce(buf)
            gmean = torch.empty(H, dtype=torch.float32, device=mean.device)
            gvar = torch.empty(H, dtype=torch.float32, device=mean.device)
            check(lib().ercg_bn_sync_unpack(_p(buf), H, _p(gmean), _p(gvar), _stream()), "ercg_bn_sync_unpack")
            return gmean, gvar, (float(self.global_count) if self.global_count else float(buf[2 * H].item()))
        # CPU tensors (gloo tests of the host logic): the same arithmetic as elementwise torch expressions
        buf = torch.empty(2 * H + 1, dtype=torch.float64, device=mean.device)
        buf[:H] = mean.double() * n_local
        buf[H:2 * H] = (var.double() + mean.double() ** 2) * n_local
        buf[2 * H:].fill_(float(n_local))           # (indexed assignment of a Python number is a CPU->GPU copy: not capturable)
        self.reduce(buf)
        count = buf[2 * H]
        gmean = buf[:H] / count
        gvar = (buf[H:2 * H] / count - gmean ** 2).clamp_(min=0)
        return gmean.float(), gvar.float(), (float(self.global_count) if self.global_count else float(count.item()))

    def grads(self, sums):
        self.reduce(sums)
        return sums


class LossSync:
    def __init__(self, group=None):
        self.group = group

    def __call__(self, num_den):
        out = num_den.clone()
        dist.all_reduce(out, group=self.group)
        return out


class GradSync:
    """One flat all-reduce(sum) of every parameter that received a gradient."""

    def __init__(self, module, group=None, average=False):
        self.module, self.group, self.average = module, group, average
        self._flat = None
        self._live = None

    def __call__(self):
        live = [p for p in self.module.parameters() if p.grad is not None]
        n = sum(p.numel() for p in live)
        if self._flat is None or self._flat.numel() != n or self._flat.device != live[0].device:
            self._flat = torch.empty(n, dtype=torch.float32, device=live[0].device)
        flat, off = self._flat, 0
        torch._foreach_copy_([flat[o:o + p.numel()].view_as(p) for o, p in _offsets(live)], [p.grad for p in live])
        dist.all_reduce(flat, group=self.group)
        if self.average:
            flat.div_(dist.get_world_size(self.group))
        torch._foreach_copy_([p.grad for p in live], [flat[o:o + p.numel()].view_as(p) for o, p in _offsets(live)])
        return n


def _offsets(params):
    off = 0
    for p in params:
        yield off, p
        off += p.numel()
