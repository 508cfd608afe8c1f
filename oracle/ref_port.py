"""CPU port of the reference's COGMEN train-step path, structured like the reference (python loops and
all) so that TIMING it is a fair stand-in for timing the reference on the GPU box, where
/root/reference does not exist.

TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/__init__.py): used by bench.py's ``cpu_baseline`` leg and
``--impl reference`` arm ("kind": "port") and by tests that pin it against the real reference.

What is kept faithful to the reference (because it is where the reference spends its time):
  * batch_graphify walks every edge in Python, builds one small tensor per edge and calls ``.item()`` twice per
    edge for the speakers, then looks the relation up in the string-keyed dict (cogmen_utils.py:121-137);
  * edge_perms builds python sets per utterance (cogmen_utils.py:147-172);
  * the 2-layer TransformerEncoder over the padded batch runs and its output is thrown away
    (cogmen.py:94-109,146-147); ``skip_dead_encoder=True`` gives the second, un-inflated CPU figure;
  * RGCNConv / TransformerConv are the pure-PyTorch stand-ins of oracle/pyg_standin.py (PyG itself is not
    installable here), BatchNorm in train mode, dropout on, Adam step as in cogmen.py:187-189.
"""
import numpy as np
import torch
from torch import nn
import torch.nn.functional as F

from .modules import CogmenGNN


def window_pairs(length, past, future):
    """Set-of-pairs construction, one utterance at a time (cogmen_utils.py:147-172)."""
    pairs = set()
    idx = np.arange(length)
    for j in range(length):
        if past == -1 and future == -1:
            nb = idx
        elif past == -1:
            nb = idx[: min(length, j + future + 1)]
        elif future == -1:
            nb = idx[max(0, j - past):]
        else:
            nb = idx[max(0, j - past): min(length, j + future + 1)]
        mine = set()
        for k in nb:
            mine.add((j, k))
        pairs = pairs.union(mine)
    return list(pairs)


def graphify_loop(features, lengths, speakers, past, future, rel_ids):
    """Per-edge python loop with two ``.item()`` calls per edge (cogmen_utils.py:109-144)."""
    dev = features.device
    rows, ei, et, per_dialogue = [], [], [], []
    base = 0
    for b in range(features.size(0)):
        n = lengths[b].item()
        rows.append(features[b, :n, :])
        local = window_pairs(n, past, future)
        shifted = [(p[0] + base, p[1] + base) for p in local]
        base += n
        per_dialogue.append(len(local))
        for p, q in zip(local, shifted):
            ei.append(torch.tensor([q[0], q[1]]))
            s0 = speakers[b, p[0]].item()
            s1 = speakers[b, p[1]].item()
            tag = "0" if p[0] < p[1] else "1"
            et.append(rel_ids[str(s0) + str(s1) + tag])
    return (torch.cat(rows, dim=0).to(dev), torch.stack(ei).t().contiguous().to(dev),
            torch.tensor(et).long().to(dev), torch.tensor(per_dialogue).long().to(dev))


class CogmenRefPort(nn.Module):
    """COGMENModule (cogmen.py:77-160) incl. the dead encoder; parameter names as in the reference."""

    def __init__(self, input_size, hidden_size=100, num_head=17, n_speakers=2, n_classes=4, skip_dead_encoder=False):
        super().__init__()
        head = None
        for h in range(6, num_head):
            if input_size % h == 0:
                head = h
                break
        assert head is not None, input_size
        layer = nn.TransformerEncoderLayer(d_model=input_size, nhead=head, dropout=0.5, batch_first=True)
        self.rnn = nn.ModuleList([nn.TransformerEncoder(layer, num_layers=2, enable_nested_tensor=False),
                                  nn.Linear(input_size, hidden_size)])
        self.gcn = CogmenGNN(hidden_size, hidden_size, hidden_size)   # cogmen.py:114 (default n_speakers)
        self.cls = nn.Sequential(nn.Linear(100, 100), nn.ReLU(), nn.Dropout(0.5), nn.Linear(100, n_classes))
        self.rel_ids = {}
        for a in range(n_speakers):
            for b in range(n_speakers):
                self.rel_ids[str(a) + str(b) + "0"] = len(self.rel_ids)
                self.rel_ids[str(a) + str(b) + "1"] = len(self.rel_ids)
        self.skip_dead_encoder = skip_dead_encoder

    def forward(self, input_tensor, speaker_tensor, text_length, *a, **k):
        node = input_tensor
        for i, mod in enumerate(self.rnn):
            if i == 0 and self.skip_dead_encoder:
                continue
            node = mod(input_tensor)                  # every module sees input_tensor: the encoder result is dropped
        feats, ei, et, _ = graphify_loop(node, text_length, speaker_tensor, 5, 5, self.rel_ids)
        return self.cls(self.gcn(feats, ei, et)), feats


def train_step(model, optim, batch):
    """cogmen.py:179-195."""
    logits, _ = model(batch["input_tensor"], batch["speaker_tensor"], batch["text_length"])
    loss = F.cross_entropy(logits, batch["label"])
    optim.zero_grad()
    loss.backward()
    optim.step()
    return loss
