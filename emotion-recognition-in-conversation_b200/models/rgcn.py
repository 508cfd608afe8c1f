"""Drop-in for the reference's vendored RGCNConv with ``edge_norm`` (models/rgcn.py:264-361).

Same constructor, parameters (basis, att, root, bias), initialisation (models/rgcn.py:316-321) and
forward signature.  The reference materialises one [in,out] weight per EDGE (index_select + bmm,
80 KB/edge, models/rgcn.py:338-341); here the basis decomposition is applied once
(W = att @ basis), the node features are transformed once per relation (one GEMM against
[in,(R+1)*out]) and a deterministic gather applies edge_norm at the destination.
"""
import math

import torch
from torch import nn

from .. import ops
from ..graph import graph_from_edge_index


class RGCNConv(nn.Module):
    def __init__(self, in_channels, out_channels, num_relations, num_bases, root_weight=True, bias=True, **kwargs):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.num_relations, self.num_bases = num_relations, num_bases
        self.basis = nn.Parameter(torch.empty(num_bases, in_channels, out_channels))
        self.att = nn.Parameter(torch.empty(num_relations, num_bases))
        if root_weight:
            self.root = nn.Parameter(torch.empty(in_channels, out_channels))
        else:
            self.register_parameter("root", None)
        if bias:
            self.bias = nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        bound = 1.0 / math.sqrt(self.num_bases * self.in_channels)
        with torch.no_grad():
            for p in (self.basis, self.att, self.root, self.bias):
                if p is not None:
                    p.uniform_(-bound, bound)

    def forward(self, x, edge_index, edge_type, edge_norm=None, size=None):
        if x is None:
            raise NotImplementedError("featureless (embedding-lookup) mode is not used by the reference's models")
        R, H, K = self.num_relations, self.out_channels, self.in_channels
        attached = getattr(edge_index, "_ercg_graph", None)
        n_graph = attached.N if attached is not None else (int(edge_index.max()) + 1 if edge_index.numel() else 0)
        if (attached is not None and x.size(0) != n_graph) or x.size(0) < n_graph:   # models/rgcn.py:126-130
            raise ValueError("Encountered node tensor with size %d in dimension 0, but expected size %d." % (x.size(0), n_graph))
        g = graph_from_edge_index(edge_index, edge_type, x.size(0), R)
        w = ops.matmul_kn(self.att, self.basis.reshape(self.num_bases, K * H))          # [R, K*H]
        w = w.view(R, K, H).permute(1, 0, 2).reshape(K, R * H)
        root_off = -1
        if self.root is not None:
            w = torch.cat([w, self.root], dim=1)
            root_off = R * H
        y = ops.matmul_kn(x, w)
        if edge_norm is not None and g.perm is not None:
            edge_norm = edge_norm[g.perm]
        return ops.gather(y, g, H, R, w=edge_norm, bias=self.bias, root_off=root_off)

    def __repr__(self):
        return "{}({}, {}, num_relations={})".format(self.__class__.__name__, self.in_channels, self.out_channels,
                                                     self.num_relations)
