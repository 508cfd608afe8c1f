"""Drop-in for the DialogueGCN part of track_mm/dgcnv2_models.py (reference :517-773): ``MatchingAttention``
('general2'), ``MaskedEdgeAttention`` ('attn1'), ``edge_perms``, ``batch_graphify`` and ``GraphNetwork`` with the reference's
constructor arguments, signatures and state_dict keys.

Kernel mapping
  MaskedEdgeAttention.scalar (Linear 2D -> 110)   K2 over the dialogue-major full-length rows
  softmax over the sequence axis, mask, renorm    K11 (csrc/dgcnv2.cu): closed form per source node, edge order = packed CSR
  edge_perms / batch_graphify                     K1
  GraphNetwork.conv1 / conv2                      vendored RGCNConv(edge_norm, 30 bases) / GraphConv: K2 + K3
  nodal MatchingAttention over each dialogue      transform = K2, then the generic K4 kernels with the tanh option on the
                                                  fully connected per-dialogue graph (K1 with window -1 / -1)
  linear -> ReLU -> dropout -> smax_fc            ops.mlp_head
The unused attention variants of the file (SimpleAttention, Attention 'mlp', MatchingAttention 'dot' / 'general' / 'concat')
and the DialogueRNN / GRU base models are parameter holders or NotImplementedError: DGCNModule's default path (LSTM base,
'attn1', nodal 'general2') never executes them.
"""
import torch
from torch import nn

from .. import ops, ops_dgcnv2
from ..graph import build_graph, speakers_from_edge_dict
from ..models.rgcn import RGCNConv
from ..pyg_nn import GraphConv
from .cogmen_utils import edge_perms  # noqa: F401  (same function in the reference, dgcnv2_models.py:612-635)
from .dgcn_models import _fresh_seed


class MatchingAttention(nn.Module):
    def __init__(self, mem_dim, cand_dim, alpha_dim=None, att_type="general"):
        super().__init__()
        if att_type != "general2":
            raise NotImplementedError("only att_type='general2' is instantiated by the DialogueGCN path (dgcnv2_models.py:531,761)")
        self.mem_dim, self.cand_dim, self.att_type = mem_dim, cand_dim, att_type
        self.transform = nn.Linear(cand_dim, mem_dim, bias=True)

    def pooled(self, emotions, graph_full):
        """All nodes at once: out_t = sum_j softmax_j(tanh(<transform(e_t), e_j>)) e_j over node t's own dialogue."""
        xt = ops.linear(emotions, self.transform.weight, self.transform.bias)
        return ops.matching_attention(xt, emotions, graph_full)

    def forward(self, M, x, mask=None):
        """Reference signature (M [L,B,D], x [B,D], mask [B,L]) -> one query per dialogue.  Only the unused LSTMModel /
        DialogRNNModel classes and the per-position loop of attentive_node_features (:709-712) call it that way; the product
        path evaluates all positions of all dialogues at once (``pooled``)."""
        raise NotImplementedError("MatchingAttention.forward(M, x, mask): use pooled(emotions, graph_full) -- every node of "
                                  "every dialogue in one launch (same arithmetic as the reference's loop over positions)")


class MaskedEdgeAttention(nn.Module):
    def __init__(self, input_dim, max_seq_len):
        super().__init__()
        self.input_dim, self.max_seq_len = input_dim, max_seq_len
        self.scalar = nn.Linear(input_dim, max_seq_len, bias=False)
        self.matchatt = MatchingAttention(input_dim, input_dim, att_type="general2")       # dead in 'attn1' (keys kept)
        self.simpleatt = nn.Module()
        self.simpleatt.scalar = nn.Linear(input_dim, 1, bias=False)
        self.att = nn.Module()                                                             # Attention(score_function='mlp'): dead
        self.att.weight = nn.Parameter(torch.zeros(2 * input_dim))
        self.att.w_k, self.att.w_q, self.att.proj = (nn.Linear(input_dim, input_dim) for _ in range(3))
        self.window = (10, 10)                     # set by DGCNModule / batch_graphify (window_past, window_future)

    def edge_weights(self, m_rows, graph, Lmax):
        """m_rows [B*Lmax, 2D]: dialogue-major, all Lmax positions of every dialogue (padding included -- the reference's
        softmax runs over them, dgcnv2_models.py:544) -> edge_norm [E] in the packed by-destination order."""
        if Lmax > self.max_seq_len:
            raise IndexError("dialogue longer than max_seq_len=%d (the reference indexes a [B,%d,L] mask, dgcnv2_models.py:546-557)"
                             % (self.max_seq_len, self.max_seq_len))
        S = ops.linear(m_rows, self.scalar.weight)
        return ops_dgcnv2.masked_edge_att(S, graph, Lmax, self.window[0], self.window[1])

    def forward(self, M, lengths, edge_ind):
        """Reference signature: M [L,B,2D] -> dense scores [B, max_seq_len, L] (non-zero on the window edges only).  The edges
        are the window graph of ``self.window`` (what batch_graphify passes in ``edge_ind``)."""
        L, B, D = M.shape
        dev = M.device
        lens = torch.as_tensor(lengths).to(torch.int64).cpu()
        g = build_graph(lens, torch.zeros((B, L), dtype=torch.int64, device=dev), self.window[0], self.window[1], 1, device=dev,
                        mean_weight=False)
        nu = self.edge_weights(M.transpose(0, 1).reshape(B * L, D), g, L)
        scores = torch.zeros((B, self.max_seq_len, L), dtype=torch.float32, device=dev)
        ei = g.edge_index
        off = g.node_off.long()
        b = g.node_dlg.long()[ei[0]]
        scores[b, ei[0] - off[b], ei[1] - off[b]] = nu          # API-compat scatter (ATen); the module path never builds this
        return scores


def batch_graphify(features, qmask, lengths, window_past, window_future, edge_type_mapping, att_model):
    """Reference signature (dgcnv2_models.py:638-690): features [L,B,D] seq-first, qmask one-hot [L,B,n] ->
    (node_features [N,D], edge_index [2,E], edge_norm [E], edge_type [E], edge_index_lengths list).
    Edges come back in canonical (dialogue, destination, source) order."""
    n_speakers = speakers_from_edge_dict(edge_type_mapping)
    L, B, D = features.shape
    dev = features.device
    ids = qmask.argmax(-1).t().contiguous()
    g = build_graph(torch.as_tensor(lengths), ids, window_past, window_future, n_speakers, device=dev, mean_weight=False)
    att_model.window = (window_past, window_future)
    rows = features.transpose(0, 1).reshape(B * L, D)
    edge_norm = att_model.edge_weights(rows, g, L)
    node_features = ops.pack_rows(features, g, seq_first=True)
    g.attach()
    return node_features, g.edge_index, edge_norm, g.edge_type, [int(v) for v in g.edge_index_lengths.tolist()]


class GraphNetwork(nn.Module):
    def __init__(self, num_features, num_classes, num_relations, max_seq_len, hidden_size=64, dropout=0.5):
        super().__init__()
        self.conv1 = RGCNConv(num_features, hidden_size, num_relations, num_bases=30)
        self.conv2 = GraphConv(hidden_size, hidden_size)
        self.matchatt = MatchingAttention(num_features + hidden_size, num_features + hidden_size, att_type="general2")
        self.linear = nn.Linear(num_features + hidden_size, hidden_size)
        self.dropout = nn.Dropout(dropout)
        self.smax_fc = nn.Linear(hidden_size, num_classes)

    def packed_logits(self, x, edge_index, edge_norm, edge_type, graph_full, nodal_attn=True):
        """-> logits [N, C] on the packed nodes."""
        out = self.conv1(x, edge_index, edge_type, edge_norm=edge_norm)
        out = self.conv2(out, edge_index)
        emotions = torch.cat([x, out], dim=-1)
        if nodal_attn:
            emotions = self.matchatt.pooled(emotions, graph_full)
        p = self.dropout.p if self.training else 0.0
        return ops.mlp_head(emotions, self.linear.weight, self.linear.bias, self.smax_fc.weight, self.smax_fc.bias, p,
                            _fresh_seed() if p > 0 else 0)

    def forward(self, x, edge_index, edge_norm, edge_type, seq_lengths, umask, nodal_attn, avec):
        """Reference signature -> logits [max_len, B, C] (seq-first, zero at the padded positions; the reference leaves the
        classifier's output for zero rows there and DGCNModule drops them through the attention mask)."""
        if avec:
            raise NotImplementedError("avec regression head is not used by DGCNModule (dgcnv2.py:63)")
        lens = torch.as_tensor(seq_lengths).to(torch.int64).cpu()
        gfull = build_graph(lens, torch.zeros(int(lens.sum()), dtype=torch.int64, device=x.device), -1, -1, 1, device=x.device,
                            reference_layout=False, mean_weight=False)
        logits = self.packed_logits(x, edge_index, edge_norm, edge_type, gfull, nodal_attn)
        return ops.unpack_rows(logits, gfull, int(lens.max()), seq_first=True)
