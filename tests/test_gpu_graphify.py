"""K1 (csrc/graphify.cu) bit-exact against the numpy oracle and the reference-generated fixture."""
import numpy as np
import pytest
import torch

from oracle import graph_np

pytestmark = pytest.mark.gpu


def _build(lengths, speakers, wp, wf, n, lengths_on_gpu=False):
    import erc_b200
    from erc_b200.graph import build_graph
    lt = torch.as_tensor(np.asarray(lengths), dtype=torch.int64)
    st = torch.as_tensor(np.asarray(speakers), dtype=torch.int64)
    if lengths_on_gpu:
        lt = lt.cuda()
    g = build_graph(lt, st.cuda(), wp, wf, n)
    torch.cuda.synchronize()
    return g


def _check(g, b):
    c = lambda t: t.cpu().numpy()
    assert (g.N, g.E) == (b["N"], b["E"])
    assert c(g.totals).tolist() == [b["N"], b["E"]]
    for name in ("rowptr", "col", "etype", "t_rowptr", "t_col", "t_etype", "t_eid", "spk"):
        assert np.array_equal(c(getattr(g, name)), b[name]), name
    assert np.array_equal(c(g.node_dlg), b["dlg"])
    assert np.array_equal(c(g.node_off), b["node_off"])
    assert np.array_equal(c(g.edge_off), b["edge_off"])
    assert np.array_equal(c(g.edge_index), b["edge_index"])
    assert np.array_equal(c(g.edge_type), b["edge_type"])
    assert np.array_equal(c(g.edge_index_lengths), b["edge_index_lengths"])
    assert np.array_equal(c(g.inv_cnt), b["inv_cnt"])          # 1/c is exact in both
    # relation census: ids that occur on at least one edge, compact numbering in ascending id order
    ids = np.unique(b["edge_type"]).astype(np.int64)
    info = c(g.rel_info)
    assert info[0] == len(ids)
    assert np.array_equal(info[257:257 + len(ids)], ids)
    want_slot = np.full(256, -1, dtype=np.int64)
    want_slot[ids] = np.arange(len(ids))
    assert np.array_equal(info[1:257], want_slot)
    got_ids, slot_dev = g.relation_slots()
    assert got_ids == ids.tolist() and np.array_equal(c(slot_dev), want_slot[:g.num_relations])


def test_reference_fixture_cases(golden):
    fx = golden("graph")
    for i in range(int(fx["n_graph_cases"])):
        n, wp, wf = (int(v) for v in fx["g%d_meta" % i])
        g = _build(fx["g%d_lengths" % i], fx["g%d_speakers" % i], wp, wf, n)
        c = lambda t: t.cpu().numpy()
        assert np.array_equal(c(g.edge_index), fx["g%d_edge_index" % i])
        assert np.array_equal(c(g.edge_type), fx["g%d_edge_type" % i])
        assert np.array_equal(c(g.edge_index_lengths), fx["g%d_edge_index_lengths" % i])
        _check(g, graph_np.batch_graphify_np(fx["g%d_lengths" % i], fx["g%d_speakers" % i], wp, wf, n))


@pytest.mark.parametrize("wp,wf,n", [(5, 5, 2), (10, 10, 2), (0, 0, 1), (-1, -1, 2), (-1, 3, 3), (4, -1, 9), (2, 7, 2),
                                      (200, 200, 2)])
def test_random_batches(wp, wf, n):
    rng = np.random.default_rng((wp + 1) * 1000 + (wf + 1) * 10 + n)
    for B in (1, 2, 33, 300):
        lengths = rng.integers(1, 111, size=B)
        lengths[rng.integers(0, B)] = 1
        spk = rng.integers(0, n, size=(B, int(lengths.max())))
        b = graph_np.batch_graphify_np(lengths, spk, wp, wf, n)
        _check(_build(lengths, spk, wp, wf, n), b)
        _check(_build(lengths, spk, wp, wf, n, lengths_on_gpu=True), b)


def test_packed_speakers_and_empty_dialogues():
    rng = np.random.default_rng(5)
    lengths = np.array([0, 5, 0, 0, 12, 1, 0])
    N = int(lengths.sum())
    spk = rng.integers(0, 2, size=N)
    b = graph_np.batch_graphify_np(lengths, spk, 5, 5, 2)
    _check(_build(lengths, spk, 5, 5, 2), b)


def test_many_dialogues_multi_tile():
    """B large enough for several scan tiles and every SM: offsets must chain across tiles."""
    rng = np.random.default_rng(9)
    B = 70_000
    lengths = np.minimum(1 + rng.geometric(1 / 7.0, size=B) - 1, 40)
    lengths = np.maximum(lengths, 1)
    N = int(lengths.sum())
    spk = np.zeros(N, dtype=np.int64)
    b = graph_np.batch_graphify_np(lengths, spk, 5, 5, 2)
    g = _build(lengths, spk, 5, 5, 2)
    _check(g, b)
    # size-independent properties (also hold at BASELINE's full size): every node has a self loop, rows sorted
    ei = g.edge_index.cpu().numpy()
    assert (ei[0] == ei[1]).sum() == N
    assert np.all(np.diff(ei[1]) >= 0)
    assert set(np.unique(g.edge_type.cpu().numpy())) <= {0, 1}     # one speaker => only relations 0 (j<k), 1 (j>=k)


def test_drop_in_batch_graphify_signature(golden):
    import erc_b200
    from erc_b200.track_mm.cogmen_utils import batch_graphify, edge_perms
    from erc_b200.graph import standard_edge_dict
    fx = golden("graph")
    n, wp, wf = (int(v) for v in fx["g0_meta"])
    feats = torch.from_numpy(fx["g0_features"]).cuda()
    nf, ei, et, el = batch_graphify(feats, torch.from_numpy(fx["g0_lengths"]), torch.from_numpy(fx["g0_speakers"]).cuda(),
                                    wp, wf, standard_edge_dict(n))
    assert np.array_equal(nf.cpu().numpy(), fx["g0_node_features"])
    assert np.array_equal(ei.cpu().numpy(), fx["g0_edge_index"]) and ei.dtype == torch.int64
    assert np.array_equal(et.cpu().numpy(), fx["g0_edge_type"])
    assert np.array_equal(el.cpu().numpy(), fx["g0_edge_index_lengths"])
    for i, (L, p, f) in enumerate(fx["perm_cases"]):
        assert sorted(edge_perms(int(L), int(p), int(f))) == [tuple(r) for r in fx["perm_%d" % i].tolist()]
