"""CPU (numpy + PyTorch fp32) restatement of the reference's DAG-ERC forward.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Follows
  track_mm/dagerc.py:109-129          get_adj_v1   (windowp = 1: back to and including the latest same-speaker utterance)
  track_mm/dagerc.py:131-154          get_s_mask   (computed over ALL Lmax positions, padded speaker = one-hot of 0)
  track_mm/dagerc_models.py:83-90     mask_logic   (alpha - (1 - adj) * 1e30)
  track_mm/dagerc_models.py:312-365   GAT_dialoggcn_v1.forward
  track_mm/dagerc.py:156-198          DAGERCModule.forward (per-utterance loop, two GRUCells per layer, out_mlp on
                                      cat(H0..H4, x); logits for every padded position)
Pinned against the real reference through tests/golden/dagerc_small.npz (oracle/make_golden.py).  Parameter names equal
the reference's state_dict keys for every LIVE parameter; the dead ones (fcs.*, attentive_node_features.transform) are
not instantiated.
"""
import numpy as np
import torch
from torch import nn
import torch.nn.functional as F


def speaker_ids(speaker_onehot):
    """[B,Lmax,n] one-hot (or [B,Lmax] ids) -> int64 ids [B,Lmax]; the reference compares the one-hot LISTS for equality."""
    s = np.asarray(speaker_onehot)
    return s.argmax(-1) if s.ndim == 3 else s.astype(np.int64)


def adj_v1(spk, windowp=1):
    """get_adj_v1 on ids [B,L]: adj[b,i,j] = 1 for j = i-1 down to the windowp-th latest j with spk[j] == spk[i]."""
    B, L = spk.shape
    adj = np.zeros((B, L, L), dtype=np.float32)
    for b in range(B):
        for i in range(L):
            cnt = 0
            for j in range(i - 1, -1, -1):
                adj[b, i, j] = 1
                if spk[b, j] == spk[b, i]:
                    cnt += 1
                    if cnt == windowp:
                        break
    return adj


def s_mask(spk):
    """get_s_mask: [B,L,L] int64, 1 where the speakers agree."""
    return (spk[:, :, None] == spk[:, None, :]).astype(np.int64)


class _Gat(nn.Module):
    def __init__(self, hidden):
        super().__init__()
        self.linear = nn.Linear(2 * hidden, 1)
        self.Wr0 = nn.Linear(hidden, hidden, bias=False)
        self.Wr1 = nn.Linear(hidden, hidden, bias=False)

    def forward(self, Q, K, V, adj, smask):
        n = K.size(1)
        X = torch.cat([Q.unsqueeze(1).expand(-1, n, -1), K], 2)
        alpha = self.linear(X).permute(0, 2, 1)
        alpha = alpha - (1 - adj.unsqueeze(1)) * 1e30
        w = F.softmax(alpha, dim=2)
        sm = smask.unsqueeze(2).float()
        Vr = self.Wr0(V) * sm + self.Wr1(V) * (1 - sm)
        return w, torch.bmm(w, Vr).squeeze(1)


class DagercOracle(nn.Module):
    def __init__(self, emb_dim, n_classes=6, dropout=0.0, gnn_layers=4, hidden=300):
        super().__init__()
        self.gnn_layers = gnn_layers
        self.gather = nn.ModuleList([_Gat(hidden) for _ in range(gnn_layers)])
        self.grus_c = nn.ModuleList([nn.GRUCell(hidden, hidden) for _ in range(gnn_layers)])
        self.grus_p = nn.ModuleList([nn.GRUCell(hidden, hidden) for _ in range(gnn_layers)])
        self.fc1 = nn.Linear(emb_dim, hidden)
        in_dim = hidden * (gnn_layers + 1) + emb_dim
        self.out_mlp = nn.Sequential(nn.Linear(in_dim, hidden), nn.ReLU(), nn.Linear(hidden, hidden), nn.ReLU(),
                                     nn.Dropout(dropout), nn.Linear(hidden, n_classes))

    def forward(self, input_tensor, text_length, speaker_tensor, **kw):
        L = input_tensor.size(1)
        spk = speaker_ids(speaker_tensor.numpy())
        adj = torch.from_numpy(adj_v1(spk))
        sm = torch.from_numpy(s_mask(spk))
        H = [F.relu(self.fc1(input_tensor))]
        for l in range(self.gnn_layers):
            h = H[l]
            C = self.grus_c[l](h[:, 0])
            M = torch.zeros_like(C)
            P = self.grus_p[l](M, h[:, 0])
            H1 = (C + P).unsqueeze(1)
            for i in range(1, L):
                _, M = self.gather[l](h[:, i], H1, H1, adj[:, i, :i], sm[:, i, :i])
                C = self.grus_c[l](h[:, i], M)
                P = self.grus_p[l](M, h[:, i])
                H1 = torch.cat([H1, (C + P).unsqueeze(1)], 1)
            H.append(H1)
        H.append(input_tensor)
        return self.out_mlp(torch.cat(H, 2)), None
