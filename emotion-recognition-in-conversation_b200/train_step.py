"""The COGMEN train step as ONE replayable unit (reference: CogmenTrainer.train_step,
track_mm/cogmen.py:179-195, under accelerate/DDP, lumo/trainer/trainer.py:62-64,315-327).

    graph build (K1) -> forward_packed -> cross entropy -> backward -> gradient all-reduce -> Adam

* Nothing in the step touches the host: sizes and the possible relation ids come from the host-side batch description,
  the loss normaliser is the host-known global utterance count, the optimizer's step counter and the dropout seed live on
  the device.  ``capture()`` records the step once into a CUDA graph (collectives included) and ``replay()`` relaunches
  it: ~50 kernel launches + autograd bookkeeping per step become one graph launch -- what makes 2^20 utterances split over
  8 GPUs (1.7 ms of kernels per rank) scale.
* Collectives per step with world > 1: BatchNorm statistics (2H+1 doubles, forward), their backward sums (2H floats) --
  both only in ``bn_sync="global"`` mode; ``"local"`` is the reference's DDP behaviour (per-rank statistics) -- and the flat
  gradient buffer.  The loss numerator rides in the tail of the gradient buffer; there is no separate loss collective.
* Transport: libercgraph's peer-memory kernels over NVLink (p2p.PeerComm: ercg_p2p_allreduce for the gradient buckets,
  ercg_p2p_bn_stats / ercg_p2p_bn_act_bwd_reduce = the BatchNorm reductions fused with their exchange) when every rank can set
  them up, else torch.distributed on ``group`` -- the results are the same to the last bit (rank-ordered sums).
* The gradient all-reduce is bucketed in two: everything except the input projection is reduced on a side stream as soon as
  those gradients exist, under the projection's weight-gradient GEMM (the longest kernel of the step); the projection's own
  gradient follows on the main stream.
"""
import os

import torch
import torch.distributed as dist

from . import ops
from .graph import build_graph, graph_sizes, relation_ids_for_speakers
from .optim import FlatAdam


class CogmenTrainStep:
    def __init__(self, model, lengths, speakers_present=(0,), lr=1e-4, weight_decay=1e-8, world=1, global_utterances=None,
                 bn_sync="global", overlap=True, group=None, transport="auto"):
        """lengths: CPU int64 [B] dialogue lengths of THIS rank's batch (a step re-run on new data of the same lengths reuses
        the capture; new lengths need a new capture).  speakers_present: the speaker ids the data set uses.
        transport: "p2p" = peer-memory all-reduce kernels (ercg_p2p_allreduce, one NVSwitch node), "nccl" = torch.distributed
        on ``group``, "auto" = p2p when every rank can set it up, else torch.distributed."""
        from .dist import StatSync
        from .p2p import PeerComm, Reducer
        self.model, self.world, self.group = model, world, group
        self.dev = next(model.parameters()).device
        comms = None
        if world > 1 and transport in ("auto", "p2p") and self.dev.type == "cuda":
            # two communicators: the early gradient bucket runs on a side stream next to the main stream's collectives
            comms = PeerComm.create(group, self.dev, max_bytes=8 << 20, n=2)
            if comms is None and transport == "p2p":
                raise RuntimeError("CogmenTrainStep: transport='p2p' but the peer-memory communicator could not be set up")
        self.comms = comms
        self.reduce_main = Reducer(group, comms[0] if comms else None)
        self.reduce_side = Reducer(group, comms[1] if comms else None)
        self.lengths = lengths.to(torch.int64).cpu().contiguous()
        self.sizes = graph_sizes(self.lengths, model.wp, model.wf)
        self.lengths_dev = self.lengths.to(self.dev)
        self.rel_ids = relation_ids_for_speakers(speakers_present, model.n_speakers)
        self.global_utts = int(global_utterances if global_utterances is not None else self.sizes[0])
        self.opt = FlatAdam(model.parameters(), lr=lr, weight_decay=weight_decay, extra_slots=1)
        self.overlap = bool(overlap) and world > 1
        self.side = torch.cuda.Stream(device=self.dev) if self.overlap else None
        if world > 1 and bn_sync == "global":
            model.gcn.stat_sync = StatSync(group=group, global_count=self.global_utts, reduce=self.reduce_main)
        self._graph = None
        self._static = None
        self._pending = 0
        self._n_early = None
        self._hooks = []
        self.loss_share = None

    # ------------------------------------------------------------------------------------------ one step, eager
    def _forward_backward(self, x, spk, labels):
        m = self.model
        g = build_graph(self.lengths_dev, spk, m.wp, m.wf, m.n_speakers, device=self.dev, sizes=self.sizes,
                        relation_ids=self.rel_ids)
        logits, _ = m.forward_packed(x, spk, self.lengths_dev, graph=g)
        num_out = self.opt.extra if self.opt.live is not None else None
        loss = ops.cross_entropy(logits, labels, denom=self.global_utts if self.world > 1 else None, num_out=num_out)
        self.opt.zero_grad()
        self._pending = self._n_early if self._n_early is not None else -1
        loss.backward()
        self._last_graph = g
        return loss

    def _late_first(self, p):
        # flat layout: gradients that exist EARLY in the backward first, the input projection (computed last) at the end
        proj = self.model.rnn[1]
        return 1 if (p is proj.weight or p is proj.bias) else 0

    def _attach(self):
        self.opt.attach(order=self._late_first)
        live = self.opt.live
        self._n_early = sum(1 for p in live if self._late_first(p) == 0)
        if self.overlap:
            for p in live[:self._n_early]:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))

    def _on_grad(self, _p):
        """post-accumulate hook of every non-projection parameter: when the last one has its gradient, reduce that bucket on
        the side stream while the main stream goes on to the projection's weight gradient."""
        self._pending -= 1
        if self._pending != 0:
            return
        opt = self.opt
        opt.gather(0, self._n_early)
        ev = torch.cuda.Event()
        ev.record()
        self.side.wait_event(ev)
        with torch.cuda.stream(self.side):
            self.reduce_side(opt.span(0, self._n_early))

    def step(self, x, spk, labels):
        """One eager step.  Returns this rank's share of the global mean loss (device scalar)."""
        loss = self._forward_backward(x, spk, labels)
        if self.opt.live is None:
            self._attach()                                   # first step: decide the flat layout, then finish it un-bucketed
            if self.world > 1:
                self.opt.extra.copy_(loss.detach().reshape(1) * float(self.global_utts))
            self.opt.gather()
            if self.world > 1:
                self.reduce_main(self.opt.flat_g)
            self.opt.update()
            return loss.detach()
        opt = self.opt
        if self.world > 1:
            if self.overlap:
                opt.gather(self._n_early, None)
                self.reduce_main(opt.flat_g[opt.offsets[self._n_early]:])                    # projection bucket + loss slot
                torch.cuda.current_stream(self.dev).wait_stream(self.side)
            else:
                opt.gather()
                self.reduce_main(opt.flat_g)
        else:
            opt.gather()
        opt.update()
        return loss.detach()

    def global_loss(self):
        """Global mean loss of the last step (valid after the gradient all-reduce; world == 1: the local loss)."""
        return self.opt.extra[0] / float(self.global_utts)

    # ------------------------------------------------------------------------------------------ capture / replay
    def capture(self, x, spk, labels, warmup=3):
        """Record the step on the given STATIC input tensors (refill them in place between replays)."""
        self._static = (x, spk, labels)
        s = torch.cuda.Stream(device=self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(s):                           # warm-up on a side stream, as torch.cuda.graph requires
            for _ in range(max(warmup, 2)):                  # (first call attaches the flat buffers)
                self.step(x, spk, labels)
        torch.cuda.current_stream(self.dev).wait_stream(s)
        torch.cuda.synchronize(self.dev)
        ops.SEED_DEV = self.opt.step_dev                     # dropout seed = host seed + device step counter: fresh mask per replay
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._loss = self.step(x, spk, labels)
        ops.SEED_DEV = None
        return self

    def replay(self):
        self._graph.replay()
        return self._loss

    def check(self):
        """Raise if K1 flagged the batch of the last step (bad lengths / speaker ids / relation id outside the hint), or if
        a peer-memory collective timed out."""
        self._last_graph.check_inputs()
        for c in self.comms or ():
            c.check()

    @property
    def transport(self):
        return self.reduce_main.transport
