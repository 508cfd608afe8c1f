"""Drop-in for track_mm/mmgcn_utils.py:5-21 (``simple_batch_graphify``): seq-first padded [Lmax,B,D] -> packed [N,D]
(dialogue-major) in one kernel launch instead of a python loop of B slices.  The other four return values are None,
as in the reference."""
import torch

from .. import ops
from ..graph import build_graph


def lengths_graph(lengths, device):
    """Node bookkeeping (node_off / node_dlg) for a batch of dialogue lengths; no window edges beyond the self loop."""
    B = lengths.numel()
    spk = torch.zeros(int(lengths.sum()), dtype=torch.int32, device=device)
    return build_graph(lengths, spk, 0, 0, 1, device=device, reference_layout=False, mean_weight=False)


def simple_batch_graphify(features, lengths):
    g = lengths_graph(lengths, features.device)
    return ops.pack_rows(features, g, seq_first=True), None, None, None, None
