"""oracle/graph_np.py against the reference's own edge_perms / batch_graphify (tests/golden/graph.npz)."""
import numpy as np
import pytest

from oracle import graph_np, ref_loader


def test_edge_perms_closed_form_vs_reference_fixture(golden):
    g = golden("graph")
    for i, (L, wp, wf) in enumerate(g["perm_cases"]):
        L, wp, wf = int(L), int(wp), int(wf)
        want = g["perm_%d" % i]
        got = np.array(graph_np.edge_perms_loop(L, wp, wf)).reshape(-1, 2)
        assert np.array_equal(got, want), (L, wp, wf)
        assert graph_np.edge_count(L, wp, wf) == want.shape[0]
        # closed-form batch builder, single dialogue, one speaker
        b = graph_np.batch_graphify_np([L], np.zeros((1, L), dtype=np.int64), wp, wf, 1)
        pairs = sorted(zip(b["edge_index"][0].tolist(), b["edge_index"][1].tolist()))
        assert pairs == [tuple(p) for p in want.tolist()]


def test_batch_graphify_vs_reference_fixture(golden):
    g = golden("graph")
    for i in range(int(g["n_graph_cases"])):
        n, wp, wf = (int(v) for v in g["g%d_meta" % i])
        b = graph_np.batch_graphify_np(g["g%d_lengths" % i], g["g%d_speakers" % i], wp, wf, n)
        assert np.array_equal(b["edge_index"], g["g%d_edge_index" % i])
        assert np.array_equal(b["edge_type"], g["g%d_edge_type" % i])
        assert np.array_equal(b["edge_index_lengths"], g["g%d_edge_index_lengths" % i])
        # CSR consistency
        assert b["rowptr"][-1] == b["E"] and np.array_equal(b["col"], b["edge_index"][0])
        dst = np.repeat(np.arange(b["N"]), np.diff(b["rowptr"]))
        assert np.array_equal(dst, b["edge_index"][1])
        # transpose consistency
        assert np.array_equal(b["edge_index"][1][b["t_eid"]], b["t_col"])
        tsrc = np.repeat(np.arange(b["N"]), np.diff(b["t_rowptr"]))
        assert np.array_equal(b["edge_index"][0][b["t_eid"]], tsrc)


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")
def test_live_reference_random_cases():
    ref = ref_loader.load()
    import torch
    rng = np.random.default_rng(3)
    for _ in range(6):
        B = int(rng.integers(1, 5))
        n = int(rng.integers(1, 4))
        lengths = rng.integers(1, 20, size=B)
        wp, wf = int(rng.integers(-1, 7)), int(rng.integers(-1, 7))
        spk = rng.integers(0, n, size=(B, int(lengths.max())))
        d = {}
        for j in range(n):
            for k in range(n):
                d[str(j) + str(k) + "0"] = len(d)
                d[str(j) + str(k) + "1"] = len(d)
        feats = torch.zeros(B, int(lengths.max()), 2)
        _, ei, et, el = ref.cogmen_utils.batch_graphify(feats, torch.tensor(lengths), torch.tensor(spk), wp, wf, d)
        order = graph_np.canonical_order(ei.numpy())
        b = graph_np.batch_graphify_np(lengths, spk, wp, wf, n)
        assert np.array_equal(b["edge_index"], ei.numpy()[:, order])
        assert np.array_equal(b["edge_type"], et.numpy()[order])
        assert np.array_equal(b["edge_index_lengths"], el.numpy())
