"""Drop-in for the model classes of track_mm/cogmen.py (reference :61-160): ``GNN`` and ``COGMENModule``
with the reference's constructor arguments, forward signatures, return values and state_dict keys.

Forward/backward run in libercgraph kernels:
  Linear(hidden_all -> 100) fused with the padded->packed row gather (K2, a_rows)  cogmen.py:103-105,146-147
  batch_graphify -> packed CSR (K1)                                                 cogmen.py:149-156
  RGCNConv = relation GEMM (K2) + gather (K3); TransformerConv = QKVS GEMM + K4     cogmen.py:65-66,71-72
  BatchNorm1d(batch stats) + LeakyReLU fused                                        cogmen.py:67-68,72
  cls: Linear+ReLU(+Dropout) fused epilogue, Linear                                 cogmen.py:116-122
The TransformerEncoder ``rnn[0]`` is DEAD in the reference (its output is overwritten by
``rnn[1](input_tensor)`` at cogmen.py:146-147 and its parameters never receive a gradient); it is
instantiated for state_dict compatibility and only executed when ``run_dead_encoder=True``.
"""
import torch
from torch import nn

from .. import ops
from ..graph import build_graph, standard_edge_dict
from ..pyg_nn import RGCNConv, TransformerConv
from .cogmen_utils import batch_graphify  # noqa: F401  (re-exported like the reference)


def _fresh_seed():
    return int(torch.empty((), dtype=torch.int64).random_().item())


class GNN(nn.Module):
    def __init__(self, g_dim, h1_dim, h2_dim, n_speakers=2):
        super().__init__()
        num_relations = 2 * n_speakers ** 2
        self.conv1 = RGCNConv(g_dim, h1_dim, num_relations)
        self.conv2 = TransformerConv(h1_dim, h2_dim, heads=1, concat=True)
        self.bn = nn.BatchNorm1d(h2_dim)
        self.relu = nn.LeakyReLU()
        self.stat_sync = None        # set by the data-parallel wrapper: all-reduce of BN sums across ranks

    def _bn_relu(self, x):
        bn = self.bn
        n_local = x.size(0)
        if self.training or not bn.track_running_stats:
            fused = self.stat_sync.fused_stats(x, bn) if self.stat_sync is not None else None
            if fused is not None:                               # reduction + exchange over peer memory + running statistics
                mean, var, count = fused
            else:
                mean, var = ops.bn_stats(x)
                count = float(n_local)
                if self.stat_sync is not None:
                    mean, var, count = self.stat_sync.stats(mean, var, n_local)
                if bn.track_running_stats:
                    ops.bn_running_update(bn, mean, var, count)     # one launch, no host read of num_batches_tracked
            return ops.bn_leaky_relu(x, bn.weight, bn.bias, mean, var, bn.eps, self.relu.negative_slope, True, count,
                                     self.stat_sync.grads if self.stat_sync is not None else None)
        return ops.bn_leaky_relu(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps,
                                 self.relu.negative_slope, False)

    def forward(self, node_features, edge_index, edge_type):
        x = self.conv1(node_features, edge_index, edge_type)
        return self._bn_relu(self.conv2(x, edge_index))


class COGMENModule(nn.Module):
    def __init__(self, input_size, hidden_size, num_head, n_speakers, n_classes, build_dead_encoder=True,
                 run_dead_encoder=False):
        super().__init__()
        find_head = False
        for h in range(6, num_head):
            if input_size % h == 0:
                num_head, find_head = h, True
                break
        assert find_head, input_size
        mods = []
        if build_dead_encoder:
            layer = nn.TransformerEncoderLayer(d_model=input_size, nhead=num_head, dropout=0.5, batch_first=True)
            mods.append(nn.TransformerEncoder(layer, num_layers=2, enable_nested_tensor=False))
        else:
            mods.append(nn.Identity())
        mods.append(nn.Linear(input_size, hidden_size, bias=True))
        self.rnn = nn.ModuleList(mods)
        self.run_dead_encoder = run_dead_encoder and build_dead_encoder
        # reference quirk, reproduced (cogmen.py:114): GNN is built with its DEFAULT n_speakers = 2 -- conv1.weight is
        # [8,100,100] for every data set, so reference checkpoints load unchanged (MELD, n_speakers = 9, included);
        # relation ids >= 8 then carry no message (see pyg_nn.RGCNConv.forward)
        self.gcn = GNN(hidden_size, hidden_size, hidden_size)
        self.cls = nn.Sequential(nn.Linear(100, 100), nn.ReLU(), nn.Dropout(p=0.5), nn.Linear(100, n_classes))
        self.n_speakers = n_speakers
        self.edge_type_to_idx = standard_edge_dict(n_speakers)
        self.wp, self.wf = 5, 5                      # hard-coded in the reference (cogmen.py:153-154)

    # -- pieces shared by the padded (reference) and packed (resident) entry points
    def _classify(self, graph_out):
        lin0, drop, lin3 = self.cls[0], self.cls[2], self.cls[3]
        p = drop.p if self.training else 0.0
        # Linear -> ReLU -> Dropout -> Linear as one autograd node (one backward pass over the hidden activations)
        return ops.mlp_head(graph_out, lin0.weight, lin0.bias, lin3.weight, lin3.bias, p, _fresh_seed() if p > 0 else 0)

    def _graph_forward(self, features, g):
        g.attach()
        graph_out = self.gcn(features, g.edge_index, g.edge_type)
        return self._classify(graph_out), features

    def forward(self, input_tensor=None, speaker_tensor=None, text_length=None, *args, x_packed=None, speaker_packed=None, **kwargs):
        """Reference signature (cogmen.py:138-160): padded ``input_tensor [B,Lmax,hidden_all]``, ``speaker_tensor [B,Lmax]``,
        ``text_length [B]``; extra batch keys are swallowed.  A batch from erc_b200.collate.DeviceCollate also carries the
        packed rows (``x_packed``, ``speaker_packed``): then the padded tensors are never touched (nor built)."""
        if x_packed is not None and speaker_packed is not None:
            return self.forward_packed(x_packed, speaker_packed, text_length)
        if self.run_dead_encoder:
            self.rnn[0](input_tensor)                # result discarded, exactly like cogmen.py:146-147
        B, Lmax, D = input_tensor.shape
        g = build_graph(text_length, speaker_tensor, self.wp, self.wf, self.n_speakers, device=input_tensor.device)
        lin = self.rnn[1]
        flat = input_tensor.reshape(B * Lmax, D)
        if (D & 3) == 0 and B * Lmax >= 8192 and B * Lmax <= 2 * g.N and ops.GEMM_ENGINE == "tc":
            # big padded batches: tensor-core projection over ALL padded rows (padding wastes < 2x), then pack the 100-wide
            # result rows.  Below ~8 k rows the step is launch-bound and the exact-fp32 gather GEMM is just as fast.
            y = ops.linear(flat, lin.weight, lin.bias)
            features = ops.pack_rows(y.view(B, Lmax, -1), g)
        else:
            features = ops.linear(flat, lin.weight, lin.bias, a_rows=g.pad_row)   # Linear fused with the node packing (SIMT)
        return self._graph_forward(features, g)

    def forward_packed(self, x_packed, speaker_packed, text_length, graph=None):
        """Resident-data entry point: utterance rows already packed [N, hidden_all] (no padding)."""
        g = graph if graph is not None else build_graph(text_length, speaker_packed, self.wp, self.wf, self.n_speakers,
                                                        device=x_packed.device)
        lin = self.rnn[1]
        features = ops.linear(x_packed, lin.weight, lin.bias)
        return self._graph_forward(features, g)
