"""numpy restatement of the reference's conversation-graph construction (integer work).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Follows
  * ``edge_perms``      track_mm/cogmen_utils.py:147-172 (= dgcn_models.py:95-118)
  * ``batch_graphify``  track_mm/cogmen_utils.py:109-144 (COGMEN)
                        track_mm/dgcn_models.py:51-92    (DialogueGCN, + edge_norm)
  * ``edge_type_to_idx`` numbering  track_mm/cogmen.py:124-129
The reference returns edges in CPython set-hash order; every function here
returns them in CANONICAL order (ascending dialogue, then destination k, then
source j), which is also the packed-CSR-by-destination order the CUDA kernel
emits.  Use :func:`canonical_order` to bring reference output into that order.
"""
import numpy as np


def window(L, wp, wf):
    """Effective (past, future) reach for one dialogue; -1 means unbounded (cogmen_utils.py:158-167)."""
    return (L if wp < 0 else wp), (L if wf < 0 else wf)


def edge_perms_loop(L, wp, wf):
    """Literal loop form of edge_perms, result sorted by (j, k). Small L only."""
    out = []
    for j in range(L):
        lo = 0 if wp < 0 else max(0, j - wp)
        hi = L - 1 if wf < 0 else min(L - 1, j + wf)
        out.extend((j, k) for k in range(lo, hi + 1))
    return sorted(out)


def edge_count(L, wp, wf):
    """|E_d| in closed form."""
    P, F = window(L, wp, wf)
    P, F = min(P, L - 1), min(F, L - 1)
    if L <= 0:
        return 0
    return L * (P + F + 1) - P * (P + 1) // 2 - F * (F + 1) // 2


def canonical_order(edge_index):
    """Permutation that sorts reference edges by (dst, src). edge_index: [2,E] (row0=src j, row1=dst k)."""
    src, dst = np.asarray(edge_index[0]), np.asarray(edge_index[1])
    return np.lexsort((src, dst))


def batch_graphify_np(lengths, speakers, wp, wf, n_speakers):
    """Closed-form batch graph.

    lengths  [B] ints; speakers either padded [B,Lmax] or packed [N] ints.
    Returns dict with (all numpy):
      node_off [B+1], N, E, edge_off [B+1]
      edge_index [2,E] int64 (row0 = src j, row1 = dst k; canonical order)
      edge_type  [E] int64   ((s_j*n + s_k)*2 + [j >= k])
      edge_index_lengths [B] int64
      rowptr [N+1] int32 (CSR by destination), col [E] int32 (= edge_index[0]), etype [E] uint8
      inv_cnt [E] float32 = 1 / |{e' : dst(e') = dst(e), type(e') = type(e)}|  (PyG RGCN mean weight)
      t_rowptr [N+1], t_col [E] (dst of each out-edge, sorted by (src,dst)), t_eid [E] (index into the
      by-destination order) -- the by-source transpose used by the backward kernels.
    """
    lengths = np.asarray(lengths, dtype=np.int64).reshape(-1)
    B = lengths.shape[0]
    node_off = np.zeros(B + 1, dtype=np.int64)
    np.cumsum(lengths, out=node_off[1:])
    N = int(node_off[-1])
    speakers = np.asarray(speakers)
    if speakers.ndim == 2:
        mask = np.arange(speakers.shape[1])[None, :] < lengths[:, None]
        spk = speakers[mask].astype(np.int64)
    else:
        spk = speakers.astype(np.int64)
    assert spk.shape[0] == N

    dlg = np.repeat(np.arange(B), lengths)               # dialogue of each node
    pos = np.arange(N) - node_off[dlg]                   # position inside its dialogue
    Ln = lengths[dlg]
    P = Ln if wp < 0 else np.full(N, wp, dtype=np.int64)
    F = Ln if wf < 0 else np.full(N, wf, dtype=np.int64)
    # in-edges of k: j in [k-F, k+P] clipped
    lo = np.maximum(0, pos - F)
    hi = np.minimum(Ln - 1, pos + P)
    deg = hi - lo + 1
    rowptr = np.zeros(N + 1, dtype=np.int64)
    np.cumsum(deg, out=rowptr[1:])
    E = int(rowptr[-1])
    dst = np.repeat(np.arange(N), deg)
    src = (lo + node_off[dlg])[dst] + (np.arange(E) - rowptr[dst])
    etype = (spk[src] * n_speakers + spk[dst]) * 2 + (src >= dst)
    edge_off = rowptr[node_off]
    R = 2 * n_speakers * n_speakers
    key = dst * R + etype
    _, inv, cnt = np.unique(key, return_inverse=True, return_counts=True)
    inv_cnt = (1.0 / cnt[inv].astype(np.float32)).astype(np.float32)
    # transpose (by source): sort edges by (src, dst)
    t_eid = np.lexsort((dst, src))
    t_deg = np.bincount(src, minlength=N)
    t_rowptr = np.zeros(N + 1, dtype=np.int64)
    np.cumsum(t_deg, out=t_rowptr[1:])
    return dict(
        node_off=node_off, N=N, E=E, edge_off=edge_off,
        edge_index=np.stack([src, dst]).astype(np.int64), edge_type=etype.astype(np.int64),
        edge_index_lengths=np.diff(edge_off).astype(np.int64),
        rowptr=rowptr.astype(np.int32), col=src.astype(np.int32), etype=etype.astype(np.uint8),
        inv_cnt=inv_cnt, t_rowptr=t_rowptr.astype(np.int32), t_col=dst[t_eid].astype(np.int32),
        t_eid=t_eid.astype(np.int32), t_etype=etype[t_eid].astype(np.uint8), spk=spk.astype(np.int32),
        dlg=dlg.astype(np.int32), pos=pos.astype(np.int32),
    )
