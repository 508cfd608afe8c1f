"""CPU (PyTorch) restatement of the declare-lab DialogueGCN variant: ``DGCNModule`` of track_mm/dgcnv2.py:54-181 with
``MaskedEdgeAttention`` 'attn1' (dgcnv2_models.py:517-562), ``batch_graphify`` (:638-690), ``GraphNetwork`` (:753-773) and the
nodal ``MatchingAttention`` 'general2' (:119-148, via attentive_node_features :693-720), base model 'LSTM' (unpacked).
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py); pinned to the real classes by tests/golden/dgcnv2_small.npz
(oracle/make_golden.py) and, when the reference is present, live."""
import numpy as np
import torch
from torch import nn
import torch.nn.functional as F

from . import graph_np
from .modules import VendoredRGCNConv
from .pyg_standin import GraphConv


class _MatchAtt(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.transform = nn.Linear(d, d, bias=True)


class _EdgeAtt(nn.Module):
    def __init__(self, d, max_seq_len):
        super().__init__()
        self.scalar = nn.Linear(d, max_seq_len, bias=False)


class _GraphNet(nn.Module):
    def __init__(self, nf, C, R, hidden, dropout):
        super().__init__()
        self.conv1 = VendoredRGCNConv(nf, hidden, R, num_bases=30)
        self.conv2 = GraphConv(hidden, hidden)
        self.matchatt = _MatchAtt(nf + hidden)
        self.linear = nn.Linear(nf + hidden, hidden)
        self.dropout = nn.Dropout(dropout)
        self.smax_fc = nn.Linear(hidden, C)


class Dgcnv2Oracle(nn.Module):
    def __init__(self, input_size, hidden_size=100, n_speakers=2, wp=10, wf=10, n_classes=7, dropout=0.0, base_model="LSTM"):
        super().__init__()
        self.base_model = base_model
        if base_model == "LSTM":
            self.lstm = nn.LSTM(input_size=input_size, hidden_size=hidden_size, num_layers=2, bidirectional=True, dropout=dropout)
        else:
            self.base_linear = nn.Linear(input_size, 2 * hidden_size)
        self.att_model = _EdgeAtt(2 * hidden_size, 110)
        self.graph_net = _GraphNet(2 * hidden_size, n_classes, 2 * n_speakers ** 2, 100, dropout)
        self.n_speakers, self.wp, self.wf = n_speakers, wp, wf

    def forward(self, input_tensor, speaker_tensor, attention_mask, text_length, **kw):
        L, B, _ = input_tensor.shape
        lens = [int(v) for v in text_length]
        emo = self.lstm(input_tensor)[0] if self.base_model == "LSTM" else self.base_linear(input_tensor)      # [L,B,2h]
        ids = speaker_tensor.argmax(-1).t()                                                                     # [B,L]
        g = graph_np.batch_graphify_np(np.asarray(lens), ids.numpy(), self.wp, self.wf, self.n_speakers)
        ei, et = torch.from_numpy(g["edge_index"]), torch.from_numpy(g["edge_type"])
        off = torch.from_numpy(g["node_off"])
        b = torch.from_numpy(g["dlg"]).long()[ei[0]]
        i, j = ei[0] - off[b], ei[1] - off[b]
        # MaskedEdgeAttention 'attn1'
        alpha = F.softmax(self.att_model.scalar(emo), dim=0).permute(1, 2, 0)              # [B,110,L]: softmax over ALL L positions
        mask = torch.full_like(alpha, 1e-10)
        mask[b, i, j] = 1
        sums = (alpha * mask).sum(-1, keepdim=True)
        edge_norm = (alpha * mask / sums)[b, i, j]
        feats = torch.cat([emo[:lens[d], d] for d in range(B)], 0)
        gn = self.graph_net
        out = gn.conv2(gn.conv1(feats, ei, et, edge_norm), ei)
        emotions = torch.cat([feats, out], -1)
        pooled = []
        for d in range(B):                                                                  # attentive_node_features, per dialogue
            e = emotions[int(off[d]):int(off[d + 1])]
            a = F.softmax(torch.tanh(gn.matchatt.transform(e) @ e.t()), dim=1)              # masked softmax + renorm == softmax over the valid rows
            pooled.append(a @ e)
        hidden = gn.dropout(F.relu(gn.linear(torch.cat(pooled, 0))))
        return gn.smax_fc(hidden), feats


def inputs(lengths, dim, n_speakers, n_classes, seed):
    """Seq-first batch with one-hot speakers (dgcnv2.py:44-45); zero padding, padded speaker id 0."""
    B, L = len(lengths), max(lengths)
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(L, B, dim, generator=g)
    spk = torch.randint(0, n_speakers, (L, B), generator=g)
    mask = torch.zeros(B, L)
    for b, n in enumerate(lengths):
        x[n:, b] = 0
        spk[n:, b] = 0
        mask[b, :n] = 1
    y = torch.randint(0, n_classes, (sum(lengths),), generator=g)
    return dict(input_tensor=x, speaker_tensor=F.one_hot(spk, n_speakers).float(), attention_mask=mask,
                text_length=torch.tensor(lengths), label=y)
