"""CPU (PyTorch fp32) restatement of the reference's MMGCN forward.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Follows
  track_mm/mmgcn.py:56-122           MMGCNModule (linear_{a,v,l}, UNPACKED text LSTM over the zero padding, smax_fc)
  track_mm/mmgcn_utils.py:5-21       simple_batch_graphify (seq-first padded -> packed, dialogue-major)
  track_mm/mmgcn_models.py:27-39     GraphConvolution (variant=True, residual=False)
  track_mm/mmgcn_models.py:373-394   GCNII_lyc.forward (64 layers, use_residue, return_feature)
  track_mm/mmgcn_models.py:530-580   MMGCN.forward (speaker embedding added to the text nodes only)
  track_mm/mmgcn_models.py:582-646   create_big_adj (angular similarity blocks + cross-modal diagonals, D^-1/2 A D^-1/2)
Pinned against the real reference through tests/golden/mmgcn_small.npz (oracle/make_golden.py).  Parameter names
equal the reference's state_dict keys for every LIVE parameter; the dead ones (att_model.*, gatedatt.*,
graph_model.{a_fc,v_fc,l_fc,feature_fc,final_fc,modal_embeddings,*_spk_embs}) are not instantiated.
"""
import math

import numpy as np
import torch
from torch import nn
import torch.nn.functional as F


def simple_pack(features, lengths):
    """[Lmax,B,D] -> [N,D]  (mmgcn_utils.py:5-21)."""
    return torch.cat([features[: int(lengths[j]), j, :] for j in range(features.size(1))], 0)


def big_adj(feats, lengths):
    """create_big_adj (mmgcn_models.py:582-646) for M = len(feats) modalities: dense [M*N, M*N]."""
    M, N = len(feats), feats[0].size(0)
    adj = torch.zeros(M * N, M * N, dtype=feats[0].dtype)
    normed = [x / torch.sqrt((x * x).sum(1, keepdim=True)) for x in feats]
    start = 0
    blocks = []
    for L in [int(v) for v in lengths]:
        sl = slice(start, start + L)
        for m in range(M):
            for n in range(M):
                r0, c0 = start + N * m, start + N * n
                if m == n:
                    c = (normed[m][sl] @ normed[m][sl].t()) * 0.99999
                    blocks.append((r0, c0, 1 - torch.acos(c) / np.pi))
                else:
                    c = (normed[m][sl] * normed[n][sl]).sum(1) * 0.99999
                    blocks.append((r0, c0, torch.diag(1 - torch.acos(c) / np.pi)))
        start += L
    # assemble without in-place writes into a leaf (keeps autograd simple)
    rows = []
    adj = torch.zeros(M * N, M * N, dtype=feats[0].dtype)
    for r0, c0, blk in blocks:
        L = blk.size(0)
        pad = torch.zeros(M * N, M * N, dtype=blk.dtype)
        adj = adj + F.pad(blk, (c0, M * N - c0 - L, r0, M * N - r0 - L))
    d = adj.sum(1)
    dinv = d.pow(-0.5)
    return dinv[:, None] * adj * dinv[None, :]


class _GraphConvolution(nn.Module):
    def __init__(self, nhidden):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(2 * nhidden, nhidden))
        stdv = 1.0 / math.sqrt(nhidden)
        self.weight.data.uniform_(-stdv, stdv)


class _GCNII(nn.Module):
    def __init__(self, nfeat, nlayers, nhidden, dropout, lamda, alpha):
        super().__init__()
        self.convs = nn.ModuleList([_GraphConvolution(nhidden) for _ in range(nlayers)])
        self.fcs = nn.ModuleList([nn.Linear(nfeat, nhidden)])
        self.dropout, self.lamda, self.alpha = dropout, lamda, alpha

    def forward(self, x, adj):
        x = F.dropout(x, self.dropout, training=self.training)
        h0 = F.relu(self.fcs[0](x))
        h = h0
        for i, con in enumerate(self.convs):
            h = F.dropout(h, self.dropout, training=self.training)
            theta = math.log(self.lamda / (i + 1) + 1)
            hi = adj @ h
            support = torch.cat([hi, h0], 1)
            r = (1 - self.alpha) * hi + self.alpha * h0
            h = F.relu(theta * (support @ con.weight) + (1 - theta) * r)
        h = F.dropout(h, self.dropout, training=self.training)
        return torch.cat([x, h], -1)


class _MMGCNGraph(nn.Module):
    def __init__(self, n_dim, nlayers, nhidden, dropout, n_speakers):
        super().__init__()
        self.graph_net = _GCNII(n_dim, nlayers, nhidden, dropout, 0.5, 0.1)
        self.speaker_embeddings = nn.Embedding(n_speakers, n_dim)

    def forward(self, a, v, l, lengths, qmask):
        q = torch.cat([qmask[: int(x), i, :] for i, x in enumerate(lengths)], 0)
        l = l + self.speaker_embeddings(q.argmax(-1))
        adj = big_adj([a, v, l], lengths)
        out = self.graph_net(torch.cat([a, v, l], 0), adj)
        N = l.size(0)
        return torch.cat([out[:N], out[N:2 * N], out[2 * N:]], -1)


class MmgcnOracle(nn.Module):
    def __init__(self, hidden_text, hidden_audio, hidden_visual, n_speakers=2, n_classes=6, dropout=0.4, nlayers=64):
        super().__init__()
        self.linear_l = nn.Linear(hidden_text, 200)
        self.lstm_l = nn.LSTM(200, 100, 2, bidirectional=True, dropout=dropout)
        self.linear_a = nn.Linear(hidden_audio, 200)
        self.linear_v = nn.Linear(hidden_visual, 200)
        self.graph_model = _MMGCNGraph(200, nlayers, 200, dropout, n_speakers)
        self.dropout_ = nn.Dropout(dropout)
        self.smax_fc = nn.Linear(1200, n_classes)

    def forward(self, text_feature, audio_feature, visual_feature, speaker_tensor, text_length, **kw):
        fa = simple_pack(self.linear_a(audio_feature), text_length)
        fv = simple_pack(self.linear_v(visual_feature), text_length)
        fl = simple_pack(self.lstm_l(self.linear_l(text_feature))[0], text_length)
        feat = self.graph_model(fa, fv, fl, text_length, speaker_tensor)
        return self.smax_fc(F.relu(self.dropout_(feat))), None
