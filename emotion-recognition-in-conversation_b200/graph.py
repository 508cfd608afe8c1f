"""Packed conversation graph (CSR by destination + by-source transpose) built by kernel K1.

Host-side mirror of the reference's graph construction:
  batch_graphify / edge_perms   track_mm/cogmen_utils.py:109-172, track_mm/dgcn_models.py:51-118
The per-edge python loop (2 device syncs per edge in the reference) becomes one cooperative
integer kernel launch (csrc/graphify.cu).
"""
import ctypes

import torch

from ._lib import lib, check, GraphOut


def _p(t):
    return None if t is None else t.data_ptr()


from .ops import _stream  # noqa: E402  (raw handle of torch's current stream)


import os as _os
_NO_CENSUS = bool(int(_os.environ.get("ERCG_NO_CENSUS", "0")))     # diagnostics: RGCNConv keeps all R relation slots


class PackedGraph:
    """All arrays live on the GPU.  Canonical edge order: dialogue, then destination, then source."""

    __slots__ = ("B", "N", "E", "wp", "wf", "n_speakers", "num_relations", "device", "node_off", "edge_off",
                 "rowptr", "col", "etype", "t_rowptr", "t_col", "t_etype", "t_eid", "spk", "node_dlg", "inv_cnt",
                 "edge_index", "edge_type", "edge_index_lengths", "totals", "pad_row", "perm", "Lpad", "rel_info",
                 "_rel_host", "_rel_event", "_rel_cache", "_rel_gen", "_rel_sel", "_core")

    def __init__(self):
        for s in self.__slots__:
            setattr(self, s, None)

    def mean_weight(self):
        """1/|N_r(k)| per edge -- the aggr='mean' weight of PyG RGCNConv."""
        if self.inv_cnt is None:
            self.inv_cnt = _inv_count_from_csr(self)
        return self.inv_cnt

    def attach(self):
        """Hang the packed CSR on the reference-layout ``edge_index`` tensor (``edge_index._ercg_graph``) so that the conv
        layers find it when they are handed that tensor (graph_from_edge_index), and return ``edge_index``.

        What is attached is a shallow copy WITHOUT the reference-layout tensors: ``edge_index -> core -> arrays`` has no
        cycle, so a step's graph (a few hundred MB of index arrays) is released by reference counting the moment the step
        drops it.  (Attaching ``self`` made a cycle that only the cyclic garbage collector could free; the resulting
        irregular frees sent the caching allocator back to cudaMalloc inside steady-state steps -- 50-110 ms stalls.)"""
        core = self._core
        if core is None:
            core = PackedGraph()
            for name in self.__slots__:
                if name not in ("edge_index", "edge_type", "edge_index_lengths", "_core"):
                    setattr(core, name, getattr(self, name))
            self._core = core
        if self.edge_index is None:
            # built with reference_layout=False (no int64 edge_index / edge_type): the conv layers only need a handle that
            # carries the packed graph, so an empty [2, 0] tensor stands in.  (K1 is issue-bound, not store-bound: dropping
            # the 24 bytes per edge measured 0.302 vs 0.308 ms, so the module path keeps the reference's outputs.)
            self.edge_index = torch.empty((2, 0), dtype=torch.int64, device=self.device)
            self.edge_type = torch.empty(0, dtype=torch.int64, device=self.device)
        self.edge_index._ercg_graph = core
        return self.edge_index

    def check_inputs(self):
        """Raise ValueError if K1 flagged its inputs (length beyond the padded width, speaker id out of range, sizes too
        small).  Costs one wait for the 2 KB census copy that left the GPU right behind K1 -- not a stream drain."""
        if self.rel_info is None:
            return
        if self._rel_cache is None and not _NO_CENSUS:
            self.relation_slots()
            return
        if self._rel_event is None:              # host-supplied relation ids: no census copy in flight, read the flag word
            _raise_graph_errors(int(self.rel_info[513].item()))
            return
        if self._rel_event is not None:
            self._rel_event.synchronize()
            _raise_graph_errors(int(self.rel_info[513].item()) if _CENSUS_GEN.get(id(self._rel_host)) != self._rel_gen
                                else int(self._rel_host[513]))

    def relation_sel(self):
        """Device int64 [P]: the relation ids behind the compact slots (index for selecting their weights)."""
        if self._rel_sel is not None:
            return self._rel_sel
        ids, _ = self.relation_slots()
        return self.rel_info[257:257 + len(ids)].long()

    def relation_slots(self):
        """K1's relation census: (ids, rel_slot) -- the sorted relation ids that occur on at least one edge (python list)
        and the device int32 table id -> compact slot (-1 = absent) -- or None for graphs that did not come from K1.
        The census left the GPU with an async copy right behind K1; waiting for it here does not drain the stream."""
        if self.rel_info is None or _NO_CENSUS:
            return None
        if self._rel_cache is None:
            self._rel_event.synchronize()
            info = self._rel_host.tolist()          # read NOW: the ring buffer is reused by later graphs
            if _CENSUS_GEN.get(id(self._rel_host)) != self._rel_gen:
                info = self.rel_info.cpu().tolist()  # the landing buffer was recycled before anyone asked: read the device copy
            _raise_graph_errors(info[513])
            P = info[0]
            ids = info[257:257 + P]
            self._rel_cache = (ids, self.rel_info[1:1 + self.num_relations])
        return self._rel_cache


def _raise_graph_errors(flags):
    if not flags:
        return
    what = []
    if flags & 1:
        what.append("a dialogue is longer than the padded speaker width (text_length > speaker_tensor.size(1))")
    if flags & 2:
        what.append("a speaker id lies outside [0, n_speakers)")
    if flags & 4:
        what.append("the given (N, E) sizes are smaller than the lengths imply")
    if flags & 8:
        what.append("an edge carries a relation id that is not in the relation_ids given to build_graph")
    raise ValueError("batch_graphify: invalid input -- " + "; ".join(what) +
                     " (the reference raises IndexError / KeyError at cogmen_utils.py:131-137)")


def graph_sizes(lengths_cpu, wp, wf):
    """Closed-form (N, E) from a CPU int64 lengths tensor (no device work, no sync)."""
    lengths_cpu = lengths_cpu.to(torch.int64).contiguous()
    n, e = ctypes.c_int64(), ctypes.c_int64()
    check(lib().ercg_graphify_sizes_host(lengths_cpu.data_ptr(), lengths_cpu.numel(), wp, wf, ctypes.byref(n),
                                         ctypes.byref(e)), "ercg_graphify_sizes_host")
    return n.value, e.value


def relation_ids_for_speakers(present_speakers, n_speakers):
    """Every relation id ((s_j * n + s_k) * 2 + dir) that edges between the given speaker ids can carry -- a superset of
    what a batch uses, computed on the host from the data set's speaker ids (MOSEI: {0} -> [0, 1])."""
    sp = sorted(set(int(v) for v in present_speakers))
    return sorted(((a * n_speakers + b) * 2 + d) for a in sp for b in sp for d in (0, 1))


_HINT_TABLES = {}       # (device, num_relations, ids) -> (id -> slot table int32 [num_relations], ids int64 [P]) on the device


def _hint_tables(device, num_relations, ids):
    key = (str(device), num_relations, tuple(ids))
    t = _HINT_TABLES.get(key)
    if t is None:
        table = torch.full((num_relations,), -1, dtype=torch.int32)
        table[torch.tensor(list(ids), dtype=torch.int64)] = torch.arange(len(ids), dtype=torch.int32)
        t = (table.to(device), torch.tensor(list(ids), dtype=torch.int64).to(device))
        _HINT_TABLES[key] = t
    return t


def build_graph(lengths, speakers, wp, wf, n_speakers, device=None, reference_layout=True, mean_weight=True,
                sizes=None, relation_ids=None):
    """Run K1.

    lengths   [B] int64/int32 (CPU or CUDA).  A CPU tensor avoids the single size sync.
    speakers  padded [B,Lmax] or packed [N] int64/int32 speaker ids
    sizes     optional (N, E) if the caller already knows them
    relation_ids  optional list of the relation ids that CAN occur (relation_ids_for_speakers): RGCNConv then takes its
              relation slots from this list instead of waiting for K1's census to reach the host -- nothing in the step
              touches the host, so it can be captured in a CUDA graph.  A batch that uses an id outside the list is
              flagged on the device (ERCG_GRAPH_ECENSUS, raised by check_inputs()).
    """
    B = lengths.numel()
    if not lengths.is_cuda and B:
        # host-side validation where it costs nothing (B integers already on the host); device-resident inputs are
        # checked by the kernel itself (flags in rel_info[513], raised by check_inputs() / relation_slots())
        lmin, lmax = int(lengths.min()), int(lengths.max())
        if lmin < 0:
            raise ValueError("build_graph: negative dialogue length %d" % lmin)
        if speakers.dim() == 2 and lmax > speakers.size(1):
            raise ValueError("build_graph: dialogue length %d exceeds the padded speaker width %d" % (lmax, speakers.size(1)))
    if not speakers.is_cuda and speakers.numel():
        smin, smax = int(speakers.min()), int(speakers.max())
        if smin < 0 or smax >= n_speakers:
            raise ValueError("build_graph: speaker ids must lie in [0, %d), got [%d, %d]" % (n_speakers, smin, smax))
    if device is None:
        device = speakers.device if speakers.is_cuda else torch.device("cuda", torch.cuda.current_device())
    if sizes is None:
        if lengths.is_cuda:
            tot = torch.empty(2, dtype=torch.int64, device=device)
            ldev = lengths.contiguous()
            check(lib().ercg_graphify_count(_p(ldev), 1 if ldev.dtype == torch.int64 else 0, B, wp, wf, _p(tot), _stream()),
                  "ercg_graphify_count")
            N, E = (int(v) for v in tot.tolist())      # the one host sync of the graph build
        else:
            N, E = graph_sizes(lengths, wp, wf)
    else:
        N, E = sizes
    ldev = lengths.to(device=device, non_blocking=True).contiguous()
    assert ldev.dtype in (torch.int64, torch.int32)
    sdev = speakers.to(device=device, non_blocking=True)
    assert sdev.dtype in (torch.int64, torch.int32)
    if sdev.dim() == 2:
        sdev = sdev.contiguous()
        spk_ld = sdev.size(1)
        assert sdev.size(0) == B
    else:
        sdev = sdev.contiguous()
        spk_ld = 0
        assert sdev.numel() == N

    g = PackedGraph()
    g.B, g.N, g.E, g.wp, g.wf, g.n_speakers, g.device = B, N, E, wp, wf, n_speakers, device
    g.num_relations = 2 * n_speakers * n_speakers
    i32 = dict(dtype=torch.int32, device=device)
    g.node_off = torch.empty(B + 1, **i32)
    g.edge_off = torch.empty(B + 1, **i32)
    g.rowptr = torch.empty(N + 1, **i32)
    g.t_rowptr = torch.empty(N + 1, **i32)
    g.col = torch.empty(E, **i32)
    g.t_col = torch.empty(E, **i32)
    g.t_eid = torch.empty(E, **i32)
    g.etype = torch.empty(E, dtype=torch.uint8, device=device)
    g.t_etype = torch.empty(E, dtype=torch.uint8, device=device)
    g.spk = torch.empty(N, **i32)
    g.node_dlg = torch.empty(N, **i32)
    g.totals = torch.empty(2, dtype=torch.int64, device=device)
    g.pad_row = torch.empty(N, **i32)
    g.Lpad = spk_ld
    if mean_weight:
        g.inv_cnt = torch.empty(E, dtype=torch.float32, device=device)
    if reference_layout:
        g.edge_index = torch.empty((2, E), dtype=torch.int64, device=device)
        g.edge_type = torch.empty(E, dtype=torch.int64, device=device)
        g.edge_index_lengths = torch.empty(B, dtype=torch.int64, device=device)
    g.rel_info = torch.empty(516, **i32)
    out = GraphOut(*[_p(getattr(g, n)) for n, _ in GraphOut._fields_])
    ws_bytes = lib().ercg_graphify_workspace_bytes(B)
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=device)
    check(lib().ercg_graphify_csr(_p(ldev), 1 if ldev.dtype == torch.int64 else 0, B, _p(sdev),
                                  1 if sdev.dtype == torch.int64 else 0, spk_ld, wp, wf, n_speakers, N, E,
                                  ctypes.byref(out), _p(ws), ws.numel(), _stream()), "ercg_graphify_csr")
    if relation_ids is not None:
        ids = [int(i) for i in relation_ids]
        table, sel = _hint_tables(device, g.num_relations, ids)
        check(lib().ercg_graphify_check_census(_p(table), table.numel(), _p(g.rel_info), _stream()), "ercg_graphify_check_census")
        g._rel_cache, g._rel_sel = (ids, table), sel
        return g
    g._rel_host, g._rel_event, g._rel_gen = _census_slot()
    g._rel_host.copy_(g.rel_info, non_blocking=True)
    g._rel_event.record()
    return g


# Pinned landing buffers for the relation census, allocated ONCE and reused round-robin: a fresh pinned allocation per
# graph (cudaHostAlloc when the caching host allocator has no free block) can stall the device for tens of ms.
_CENSUS_RING = []
_CENSUS_GEN = {}                     # id(host buffer) -> generation of the graph that currently owns it
_census_next = 0


def _census_slot(depth=8):
    global _census_next
    if not _CENSUS_RING:
        for _ in range(depth):
            _CENSUS_RING.append((torch.empty(516, dtype=torch.int32, pin_memory=True), torch.cuda.Event()))
        _census_next = 0
    host, ev = _CENSUS_RING[_census_next]
    _census_next = (_census_next + 1) % len(_CENSUS_RING)
    ev.synchronize()                 # the copy that last used this buffer (depth graphs ago) has landed
    gen = _CENSUS_GEN.get(id(host), 0) + 1
    _CENSUS_GEN[id(host)] = gen
    return host, ev, gen


def _inv_count_from_csr(g):
    """Mean weights for a graph that did not come from K1 (generic edge_index path); O(E) torch ops."""
    dst = torch.repeat_interleave(torch.arange(g.N, device=g.device), (g.rowptr[1:] - g.rowptr[:-1]).long())
    key = dst * g.num_relations + g.etype.long()
    _, inv, cnt = torch.unique(key, return_inverse=True, return_counts=True)
    return (1.0 / cnt[inv].float()).contiguous()


def graph_from_edge_index(edge_index, edge_type, num_nodes, num_relations=1):
    """Compatibility path: arbitrary COO edge_index [2,E] (+ edge_type) -> PackedGraph.

    Used when the drop-in layers are handed an edge_index that did not come from our batch_graphify.
    The sort runs on the GPU through torch.sort (plumbing, not the hot path); the result feeds the
    same kernels.  Canonical order = (dst, src) stable.
    """
    g = getattr(edge_index, "_ercg_graph", None)
    if g is not None:
        return g
    dev = edge_index.device
    assert dev.type == "cuda", "libercgraph has no CPU path"
    src, dst = edge_index[0].long(), edge_index[1].long()
    E = src.numel()
    et = edge_type.long() if edge_type is not None else torch.zeros(E, dtype=torch.int64, device=dev)
    order = torch.argsort(dst * num_nodes + src, stable=True)
    g = PackedGraph()
    g.B, g.N, g.E, g.device, g.num_relations = 0, num_nodes, E, dev, num_relations
    s, d, t = src[order], dst[order], et[order]
    g.col = s.int().contiguous()
    g.etype = t.to(torch.uint8).contiguous()
    g.rowptr = torch.zeros(num_nodes + 1, dtype=torch.int64, device=dev)
    g.rowptr[1:] = torch.cumsum(torch.bincount(d, minlength=num_nodes), 0)
    g.rowptr = g.rowptr.int().contiguous()
    torder = torch.argsort(s * num_nodes + d, stable=True)        # positions in by-dst order, sorted by (src,dst)
    g.t_eid = torder.int().contiguous()
    g.t_col = d[torder].int().contiguous()
    g.t_etype = t[torder].to(torch.uint8).contiguous()
    g.t_rowptr = torch.zeros(num_nodes + 1, dtype=torch.int64, device=dev)
    g.t_rowptr[1:] = torch.cumsum(torch.bincount(s, minlength=num_nodes), 0)
    g.t_rowptr = g.t_rowptr.int().contiguous()
    g.edge_index = torch.stack([s, d]).contiguous()
    g.edge_type = t.contiguous()
    # map from the caller's edge order to canonical order, for per-edge inputs such as edge_norm
    g.perm = order
    return g


def standard_edge_dict(n_speakers):
    """edge_type_to_idx exactly as the reference builds it (track_mm/cogmen.py:124-129, dgcn.py:72-77)."""
    d = {}
    for j in range(n_speakers):
        for k in range(n_speakers):
            d[str(j) + str(k) + "0"] = len(d)
            d[str(j) + str(k) + "1"] = len(d)
    return d


def speakers_from_edge_dict(edge_type_to_idx):
    n = 1
    while 2 * n * n < len(edge_type_to_idx):
        n += 1
    if edge_type_to_idx != standard_edge_dict(n):
        raise ValueError("edge_type_to_idx is not the reference's numbering ((s_j*n+s_k)*2+[j>=k]); "
                         "the graph kernel only implements that one")
    return n
