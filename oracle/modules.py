"""CPU (PyTorch fp32) restatement of the reference's model forward functions.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Vectorised closed forms, so
it finishes in seconds at test sizes; validated against the *real* reference
modules through tests/golden (oracle/make_golden.py).  A floating-point path,
hence a torch fp32 reference rather than numpy.

Follows
  COGMEN  track_mm/cogmen.py:61-74 (GNN), :77-160 (COGMENModule)
  DGCN    track_mm/dgcn_models.py:10-33 (SeqContext), :36-48 (GCN), :121-152 (EdgeAtt),
          :155-170 (Classifier); track_mm/dgcn.py:53-93 (DGCNModule);
          models/rgcn.py:324-355 (vendored RGCNConv with edge_norm)
Parameter names equal the reference's ``state_dict`` keys for every LIVE
parameter.  The COGMEN encoder ``rnn.0`` (cogmen.py:94-101) is dead code -- its
output is overwritten at cogmen.py:146-147 and its parameters never get a
gradient -- and is not instantiated here.
"""
import math

import numpy as np
import torch
from torch import nn
import torch.nn.functional as F

from . import graph_np
from .pyg_standin import RGCNConv as PygRGCNConv, TransformerConv, GraphConv


def _graph_tensors(lengths, speaker_tensor, wp, wf, n_speakers):
    g = graph_np.batch_graphify_np(lengths.cpu().numpy(), speaker_tensor.cpu().numpy(), wp, wf, n_speakers)
    return g, torch.from_numpy(g["edge_index"]), torch.from_numpy(g["edge_type"])


def pack_nodes(features, lengths):
    """[B,Lmax,D] -> [N,D] dialogue-major (cogmen_utils.py:123,139)."""
    return torch.cat([features[b, : int(lengths[b])] for b in range(features.size(0))], 0)


class CogmenGNN(nn.Module):  # cogmen.py:61-74
    def __init__(self, g_dim, h1_dim, h2_dim, n_speakers=2):
        super().__init__()
        self.conv1 = PygRGCNConv(g_dim, h1_dim, 2 * n_speakers ** 2)
        self.conv2 = TransformerConv(h1_dim, h2_dim, heads=1, concat=True)
        self.bn = nn.BatchNorm1d(h2_dim)
        self.relu = nn.LeakyReLU()

    def forward(self, x, edge_index, edge_type):
        x = self.conv1(x, edge_index, edge_type)
        return self.relu(self.bn(self.conv2(x, edge_index)))


class _Holder(nn.Module):
    pass


class CogmenOracle(nn.Module):
    """COGMENModule (cogmen.py:77-160) without the dead encoder rnn.0."""

    def __init__(self, input_size, hidden_size=100, n_speakers=2, n_classes=4, dropout=0.5, wp=5, wf=5):
        super().__init__()
        self.rnn = nn.ModuleDict({"1": nn.Linear(input_size, hidden_size)})   # key 'rnn.1.*'
        self.gcn = CogmenGNN(hidden_size, hidden_size, hidden_size)   # cogmen.py:114: DEFAULT n_speakers = 2, always 8 relations
        self.cls = nn.Sequential(nn.Linear(100, 100), nn.ReLU(), nn.Dropout(dropout), nn.Linear(100, n_classes))
        self.n_speakers, self.wp, self.wf = n_speakers, wp, wf

    def forward(self, input_tensor, speaker_tensor, text_length, *a, **k):
        feats = pack_nodes(self.rnn["1"](input_tensor), text_length)
        _, ei, et = _graph_tensors(text_length, speaker_tensor, self.wp, self.wf, self.n_speakers)
        return self.cls(self.gcn(feats, ei, et)), feats

    def forward_packed(self, x_packed, speaker_packed, text_length, drop_mask=None, act_masks=None):
        """Same computation on rows that are already packed [N, hidden_all] (cogmen_utils.py:123,139 only re-packs the
        padded tensor; Linear is row-wise, so packing before or after it is the same arithmetic).

        ``drop_mask`` [N,100] (entries 0 or 1/(1-p)) replaces nn.Dropout's own random mask in the classifier (cogmen.py:119),
        so a run of the CUDA path WITH dropout can be checked by handing its mask to the oracle.

        ``act_masks`` = (leaky_pos [N,100] bool, relu_pos [N,100] bool) pins the branch of the two piecewise-linear
        activations (LeakyReLU after the BatchNorm, cogmen.py:68,72; ReLU in the classifier, :118).  With ~10^6 x 100
        activations some pre-activation lands within rounding distance of the kink, where two correct implementations may
        pick different branches; the VALUE moves by ~1e-7 but the GRADIENT of that element jumps, which in a sum over N
        rows shows up as a 1e-4 .. 1e-3 relative change of whole weight gradients (measured: fp32 vs fp64 of this very
        oracle).  Pinning the branch compares the implementations on the same piecewise-linear piece; ``self.kinks``
        records how many elements were on the other side and how far (they must be within rounding distance of 0)."""
        feats = self.rnn["1"](x_packed)
        _, ei, et = _graph_tensors(text_length, speaker_packed, self.wp, self.wf, self.n_speakers)
        if drop_mask is None and act_masks is None:
            return self.cls(self.gcn(feats, ei, et)), feats
        g = self.gcn
        y = g.bn(g.conv2(g.conv1(feats, ei, et), ei))
        self.kinks = {}
        if act_masks is None:
            h = g.relu(y)
        else:
            pos = act_masks[0]
            off = (y.detach() > 0) != pos
            self.kinks["leaky"] = (int(off.sum()), float(y.detach()[off].abs().max()) if off.any() else 0.0)
            h = y * torch.where(pos, torch.ones((), dtype=y.dtype), torch.full((), g.relu.negative_slope, dtype=y.dtype))
        z = self.cls[0](h)
        if act_masks is None:
            a = self.cls[1](z)
        else:
            pos = act_masks[1]
            off = (z.detach() > 0) != pos
            self.kinks["relu"] = (int(off.sum()), float(z.detach()[off].abs().max()) if off.any() else 0.0)
            a = z * pos.to(z.dtype)
        if drop_mask is not None:
            a = a * drop_mask.to(a.dtype)
        return self.cls[3](a), feats


class VendoredRGCNConv(nn.Module):
    """models/rgcn.py:264-361 (basis decomposition, aggr='add', edge_norm), transform-then-gather."""

    def __init__(self, in_channels, out_channels, num_relations, num_bases):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.num_relations, self.num_bases = num_relations, num_bases
        self.basis = nn.Parameter(torch.empty(num_bases, in_channels, out_channels))
        self.att = nn.Parameter(torch.empty(num_relations, num_bases))
        self.root = nn.Parameter(torch.empty(in_channels, out_channels))
        self.bias = nn.Parameter(torch.empty(out_channels))
        bound = 1.0 / math.sqrt(num_bases * in_channels)       # models/rgcn.py:316-321
        for p in self.parameters():
            p.data.uniform_(-bound, bound)

    def forward(self, x, edge_index, edge_type, edge_norm=None):
        w = (self.att @ self.basis.view(self.num_bases, -1)).view(self.num_relations, self.in_channels, -1)
        y = torch.einsum("ni,rio->nro", x, w)
        msg = y[edge_index[0], edge_type]
        if edge_norm is not None:
            msg = msg * edge_norm.view(-1, 1)
        out = torch.zeros(x.size(0), self.out_channels, dtype=x.dtype).index_add_(0, edge_index[1], msg)
        return out + x @ self.root + self.bias


class EdgeAttOracle(nn.Module):
    """EdgeAtt (dgcn_models.py:121-152) in closed form on packed nodes: per-SOURCE window softmax of
    <x_j, W x_k>; returns the weight of every edge in the order of ``edge_index``."""

    def __init__(self, g_dim, wp, wf):
        super().__init__()
        self.wp, self.wf = wp, wf
        self.weight = nn.Parameter(torch.zeros(g_dim, g_dim))
        self.weight.data.normal_(0, 2.0 / (2 * g_dim))         # dgcn_models.py:129-130 (var used as std)

    def forward(self, x, edge_index):
        src, dst = edge_index[0], edge_index[1]
        u = x @ self.weight.t()                                 # u_k = W x_k  (dgcn_models.py:136-137)
        score = (x.index_select(0, src) * u.index_select(0, dst)).sum(-1)
        n = x.size(0)
        smax = torch.full((n,), float("-inf"), dtype=x.dtype).scatter_reduce(0, src, score, reduce="amax", include_self=True)
        ex = (score - smax[src]).exp()
        den = torch.zeros(n, dtype=x.dtype).index_add_(0, src, ex)
        return ex / den[src]


class DgcnGCN(nn.Module):  # dgcn_models.py:36-48
    def __init__(self, g_dim, h1_dim, h2_dim, n_speakers):
        super().__init__()
        self.conv1 = VendoredRGCNConv(g_dim, h1_dim, 2 * n_speakers ** 2, num_bases=30)
        self.conv2 = GraphConv(h1_dim, h2_dim)

    def forward(self, x, edge_index, edge_norm, edge_type):
        return self.conv2(self.conv1(x, edge_index, edge_type, edge_norm), edge_index)


class SeqContextOracle(nn.Module):  # dgcn_models.py:10-33
    def __init__(self, u_dim, g_dim, dropout=0.4):
        super().__init__()
        self.rnn = nn.LSTM(u_dim, g_dim // 2, dropout=dropout, bidirectional=True, num_layers=2, batch_first=True)

    def forward(self, lengths, x):
        packed = nn.utils.rnn.pack_padded_sequence(x, lengths.cpu(), batch_first=True, enforce_sorted=False)
        out, _ = self.rnn(packed, None)
        out, _ = nn.utils.rnn.pad_packed_sequence(out, batch_first=True)
        return out


class DgcnClassifier(nn.Module):  # dgcn_models.py:155-170 (emotion_att is dead: never called)
    def __init__(self, input_dim, hidden_size, tag_size, dropout):
        super().__init__()
        self.lin1 = nn.Linear(input_dim, hidden_size)
        self.drop = nn.Dropout(dropout)
        self.lin2 = nn.Linear(hidden_size, tag_size)

    def forward(self, h):
        return self.lin2(self.drop(F.relu(self.lin1(h))))


class DgcnOracle(nn.Module):
    """DGCNModule (dgcn.py:53-93)."""

    def __init__(self, n_speakers, input_size=100, hidden_size=200, context=(10, 10), dropout=0.4, n_classes=4):
        super().__init__()
        self.wp, self.wf = context
        self.n_speakers = n_speakers
        self.rnn = SeqContextOracle(input_size, hidden_size, dropout)
        self.edge_att = EdgeAttOracle(hidden_size, self.wp, self.wf)
        self.gcn = DgcnGCN(hidden_size, 100, 100, n_speakers)
        self.clf = DgcnClassifier(hidden_size + 100, 100, n_classes, dropout)

    def forward(self, input_tensor, speaker_tensor, text_length, **kw):
        ctx = self.rnn(text_length, input_tensor)
        feats = pack_nodes(ctx, text_length)
        _, ei, et = _graph_tensors(text_length, speaker_tensor, self.wp, self.wf, self.n_speakers)
        norm = self.edge_att(feats, ei)
        graph_out = self.gcn(feats, ei, norm, et)
        return self.clf(torch.cat([feats, graph_out], -1)), graph_out


def load_live(module, state_dict):
    """Copy every key of ``state_dict`` that ``module`` owns; returns the list of skipped keys."""
    own = module.state_dict()
    skipped = []
    for k, v in state_dict.items():
        if k in own:
            own[k].copy_(torch.as_tensor(v))
        else:
            skipped.append(k)
    return skipped
