"""Pure-PyTorch stand-ins for the third-party layers the reference imports.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference uses, but does not vendor:
  * ``torch_geometric.nn.RGCNConv``       -- track_mm/cogmen.py:23,65
  * ``torch_geometric.nn.TransformerConv`` -- track_mm/cogmen.py:23,66
  * ``torch_geometric.nn.GraphConv``       -- track_mm/dgcn_models.py:6,42
  * ``torch_scatter.scatter_add``          -- models/rgcn.py:12,37-38
``torch_geometric`` is unpinned (requirements.txt:12, next to torch~=1.11 => a
PyG 2.0.x era install).  The classes below restate the *published* PyG-2.x
semantics of those layers with PyG's parameter names so that a reference
``state_dict`` keeps its keys:
    RGCNConv:        weight (R,in,out), root (in,out), bias (out)
    TransformerConv: lin_key/lin_query/lin_value/lin_skip .weight/.bias
    GraphConv:       lin_rel.weight/.bias, lin_root.weight
Convention everywhere: edge_index[0] = source j, edge_index[1] = target i,
messages flow source -> target and are aggregated at the target.
"""
import math

import torch
from torch import nn


def scatter_add(src, index, dim=0, out=None, dim_size=None):
    """torch_scatter.scatter_add as called at models/rgcn.py:37-38 (dim=0, out=None)."""
    assert dim == 0 and out is None
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    res = torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    return res.index_add_(0, index, src)


def _segment_softmax(score, index, num_nodes):
    """PyG ``torch_geometric.utils.softmax``: max-subtracted, denominator + 1e-16."""
    smax = torch.full((num_nodes,) + tuple(score.shape[1:]), float("-inf"), dtype=score.dtype)
    smax = smax.scatter_reduce(0, index.view(-1, *([1] * (score.dim() - 1))).expand_as(score), score,
                               reduce="amax", include_self=True)
    ex = (score - smax.index_select(0, index)).exp()
    den = torch.zeros_like(smax).index_add_(0, index, ex) + 1e-16
    return ex / den.index_select(0, index)


class RGCNConv(nn.Module):
    """PyG-2.x RGCNConv(in, out, num_relations), no bases/blocks, aggr='mean'.

    out_i = sum_r mean_{j in N_r(i)} x_j @ weight[r] + x_i @ root + bias
    (aggregate-then-transform per relation, as PyG does).
    """

    def __init__(self, in_channels, out_channels, num_relations):
        super().__init__()
        self.in_channels, self.out_channels, self.num_relations = in_channels, out_channels, num_relations
        self.weight = nn.Parameter(torch.empty(num_relations, in_channels, out_channels))
        self.root = nn.Parameter(torch.empty(in_channels, out_channels))
        self.bias = nn.Parameter(torch.empty(out_channels))
        self.reset_parameters()

    def reset_parameters(self):
        # PyG: glorot(weight), glorot(root), zeros(bias)
        for t in (self.weight, self.root):
            a = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
            t.data.uniform_(-a, a)
        self.bias.data.zero_()

    def forward(self, x, edge_index, edge_type):
        n = x.size(0)
        src, dst = edge_index[0], edge_index[1]
        out = torch.zeros(n, self.out_channels, dtype=x.dtype, device=x.device)
        for r in range(self.num_relations):
            sel = edge_type == r
            s, d = src[sel], dst[sel]
            agg = torch.zeros(n, self.in_channels, dtype=x.dtype).index_add_(0, d, x.index_select(0, s))
            cnt = torch.zeros(n, dtype=x.dtype).index_add_(0, d, torch.ones(d.numel(), dtype=x.dtype))
            agg = agg / cnt.clamp(min=1).unsqueeze(-1)
            out = out + agg @ self.weight[r]
        out = out + x @ self.root
        return out + self.bias


class TransformerConv(nn.Module):
    """PyG-2.x TransformerConv(in, out, heads, concat=True), beta=False, dropout=0, root_weight=True."""

    def __init__(self, in_channels, out_channels, heads=1, concat=True):
        super().__init__()
        self.in_channels, self.out_channels, self.heads, self.concat = in_channels, out_channels, heads, concat
        self.lin_key = nn.Linear(in_channels, heads * out_channels)
        self.lin_query = nn.Linear(in_channels, heads * out_channels)
        self.lin_value = nn.Linear(in_channels, heads * out_channels)
        self.lin_skip = nn.Linear(in_channels, heads * out_channels if concat else out_channels)

    def forward(self, x, edge_index):
        n, h, c = x.size(0), self.heads, self.out_channels
        src, dst = edge_index[0], edge_index[1]
        q = self.lin_query(x).view(n, h, c)
        k = self.lin_key(x).view(n, h, c)
        v = self.lin_value(x).view(n, h, c)
        score = (q.index_select(0, dst) * k.index_select(0, src)).sum(-1) / math.sqrt(c)  # [E,H]
        alpha = _segment_softmax(score, dst, n)
        msg = v.index_select(0, src) * alpha.unsqueeze(-1)
        out = torch.zeros(n, h, c, dtype=x.dtype).index_add_(0, dst, msg)
        out = out.reshape(n, h * c) if self.concat else out.mean(1)
        return out + self.lin_skip(x)


class GraphConv(nn.Module):
    """PyG-2.x GraphConv(in, out), aggr='add': lin_rel(sum_j x_j) + lin_root(x_i)."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.lin_rel = nn.Linear(in_channels, out_channels, bias=True)
        self.lin_root = nn.Linear(in_channels, out_channels, bias=False)

    def forward(self, x, edge_index):
        src, dst = edge_index[0], edge_index[1]
        agg = torch.zeros_like(x).index_add_(0, dst, x.index_select(0, src))
        return self.lin_rel(agg) + self.lin_root(x)
