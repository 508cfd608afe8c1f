"""Drop-in for track_mm/cogmen_utils.py (reference :109-172): same names, arguments and return values.

``batch_graphify`` launches kernel K1 once per batch instead of looping over edges in Python.
Edges come back in canonical order (dialogue, destination, source) instead of the reference's
CPython set-hash order; every consumer is order-invariant.  The returned ``edge_index`` carries the
packed CSR (``edge_index._ercg_graph``) so the conv layers do not rebuild it.
"""
import torch

from .. import ops
from ..graph import build_graph, speakers_from_edge_dict


def edge_perms(length, window_past, window_future):
    """List of (vertex j, neighbour k) pairs of one dialogue (reference :147-172), sorted.

    Host-side helper kept for API compatibility; the kernels never call it (they use the closed form
    k in [max(0, j-wp), min(L-1, j+wf)], -1 = unbounded)."""
    out = []
    for j in range(length):
        lo = 0 if window_past == -1 else max(0, j - window_past)
        hi = length - 1 if window_future == -1 else min(length - 1, j + window_future)
        out.extend((j, k) for k in range(lo, hi + 1))
    return out


def batch_graphify(features, lengths, speaker_tensor, wp, wf, edge_type_to_idx):
    """-> (node_features [N,D], edge_index [2,E] i64, edge_type [E] i64, edge_index_lengths [B] i64)."""
    n_speakers = speakers_from_edge_dict(edge_type_to_idx)
    g = build_graph(lengths, speaker_tensor, wp, wf, n_speakers, device=features.device)
    node_features = ops.pack_rows(features, g)
    g.attach()
    return node_features, g.edge_index, g.edge_type, g.edge_index_lengths
