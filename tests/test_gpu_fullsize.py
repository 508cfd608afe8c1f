"""BASELINE.json config 5 at FULL size (~2^20 MOSEI-shaped utterances, hidden_all 1443, one speaker id): the oracle cannot
run there, so the CUDA path is checked through size-independent properties of the domain --

  K1   closed-form node / edge counts and degrees, contiguous sorted rows with a self-loop, the by-source transpose is a
       permutation of the by-destination edges, relation ids follow the reference formula, the PyG mean weights of every
       (destination, relation) bucket sum to one, the relation census is exactly {past, future-or-self}
  K2   checksum of checksums: colsum(X @ W) == colsum(X) @ W and (X^T @ dF) @ 1 == X^T @ (dF @ 1) against fp64
  K3   linearity of the gather in Y; K4  attention weights of every destination sum to one and a constant value row is
       reproduced exactly (out = c + skip)
  step two identical train steps (same dropout seed) give bit-identical loss and gradients
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

HIDDEN, H, WP, WF = 1443, 100, 5, 5


@pytest.fixture(scope="module")
def big():
    import erc_b200
    from erc_b200 import synth
    from erc_b200.graph import build_graph, graph_sizes
    lengths = synth.config5_lengths(1 << 20, seed=0)
    N, E = graph_sizes(lengths, WP, WF)
    spk = torch.zeros(N, dtype=torch.int64, device="cuda")
    g = build_graph(lengths, spk, WP, WF, 2, sizes=(N, E))
    torch.cuda.synchronize()
    return dict(lengths=lengths, N=N, E=E, g=g, spk=spk)


def test_graph_properties_at_full_size(big):
    g, L = big["g"], big["lengths"].cuda()
    N, E = big["N"], big["E"]
    assert N >= 1 << 20 and N == int(L.sum())
    P = torch.minimum(torch.full_like(L, WP), L - 1)
    F = torch.minimum(torch.full_like(L, WF), L - 1)
    assert E == int((L * (P + F + 1) - P * (P + 1) // 2 - F * (F + 1) // 2).sum())      # closed form of edge_perms
    assert g.totals.tolist() == [N, E]
    off = g.node_off.long()
    assert torch.equal(off[1:] - off[:-1], L) and int(off[-1]) == N
    rp, col = g.rowptr.long(), g.col.long()
    node = torch.arange(N, device="cuda")
    d = g.node_dlg.long()
    k = node - off[d]
    lo = torch.clamp(k - WF, min=0) + off[d]                       # in-edges of k: [k - wf, k + wp] within the dialogue
    hi = torch.minimum(k + WP, L[d] - 1) + off[d]
    assert torch.equal(rp[1:] - rp[:-1], hi - lo + 1) and int(rp[0]) == 0 and int(rp[-1]) == E
    row = torch.repeat_interleave(node, rp[1:] - rp[:-1])
    e = torch.arange(E, device="cuda")
    assert torch.equal(col, lo[row] + (e - rp[row]))               # rows are contiguous ascending windows (incl. the self-loop)
    # relation id = ((s_j * n + s_k) * 2) + [j >= k]; one speaker id 0 -> {0, 1}
    assert torch.equal(g.etype.long(), (col >= row).long())
    ids, _ = g.relation_slots()
    assert ids == [0, 1]
    # by-source transpose: a permutation of the edges, consistent with the by-destination arrays
    teid = g.t_eid.long()
    assert torch.equal(torch.sort(teid).values, e)
    trow = torch.repeat_interleave(node, g.t_rowptr.long()[1:] - g.t_rowptr.long()[:-1])
    assert torch.equal(col[teid], trow) and torch.equal(row[teid], g.t_col.long())
    assert torch.equal(g.etype[teid], g.t_etype)
    # PyG mean weights: every non-empty (destination, relation) bucket sums to one
    bucket = row * 2 + g.etype.long()
    sums = torch.zeros(2 * N, dtype=torch.float64, device="cuda").index_add_(0, bucket, g.inv_cnt.double())
    nonempty = torch.zeros(2 * N, dtype=torch.bool, device="cuda")
    nonempty[bucket] = True
    assert float((sums[nonempty] - 1.0).abs().max()) < 1e-6 and float(sums[~nonempty].abs().max()) == 0.0


def test_dense_transform_checksums_at_full_size(big):
    from erc_b200 import ops
    N = big["N"]
    gen = torch.Generator(device="cuda").manual_seed(1)
    store = torch.empty((N, HIDDEN + 1), dtype=torch.float32, device="cuda").normal_(generator=gen)
    X = store[:, :HIDDEN]
    W = (torch.randn(HIDDEN, H, generator=gen, device="cuda") * 0.05)
    Y = ops.gemm_nn(X, W)                                           # tcgen05 path, M = N rows
    xs = torch.zeros(HIDDEN, dtype=torch.float64, device="cuda")
    for c in range(0, N, 1 << 17):
        xs += X[c:c + (1 << 17)].double().sum(0)
    want = xs @ W.double()
    got = Y.double().sum(0)
    scale = float((X[: 1 << 17].abs().double().sum(0) @ W.abs().double()).max()) * (N / float(1 << 17))
    assert float((got - want).abs().max()) < 1e-5 * scale           # colsum(X W) == colsum(X) W
    dF = torch.randn(N, H, generator=gen, device="cuda")
    G = ops.gemm_tn(X, dF)                                          # [HIDDEN, H] weight gradient, contraction over N rows
    r = dF.double().sum(1)
    want2 = torch.zeros(HIDDEN, dtype=torch.float64, device="cuda")
    for c in range(0, N, 1 << 17):
        want2 += X[c:c + (1 << 17)].double().t() @ r[c:c + (1 << 17)]
    got2 = G.double().sum(1)
    scale2 = float((X[: 1 << 17].abs().double().t() @ dF[: 1 << 17].abs().double().sum(1)).max()) * (N / float(1 << 17))
    assert float((got2 - want2).abs().max()) < 1e-5 * scale2        # (X^T dF) 1 == X^T (dF 1)


def test_gather_linearity_and_attention_invariants_at_full_size(big):
    from erc_b200 import ops
    g, N = big["g"], big["N"]
    gen = torch.Generator(device="cuda").manual_seed(2)
    ids, rel_slot = g.relation_slots()
    Pn = len(ids)
    Y1 = torch.randn(N, (Pn + 1) * H, generator=gen, device="cuda")
    Y2 = torch.randn(N, (Pn + 1) * H, generator=gen, device="cuda")
    w = g.mean_weight()
    f = lambda Y: ops.gather(Y, g, H, 8, w=w, root_off=Pn * H, rel_slot=rel_slot, n_slots=Pn)
    lhs = f(2.0 * Y1 - 0.5 * Y2)
    rhs = 2.0 * f(Y1) - 0.5 * f(Y2)
    assert float((lhs - rhs).abs().max()) < 1e-5 * float(rhs.abs().max())
    # a convex combination of the neighbours' rows (+ the node's own row): constant rows are reproduced
    ones = torch.ones(N, (Pn + 1) * H, device="cuda")
    buckets = f(ones)                                               # = (number of non-empty relation buckets) + 1 (root)
    assert float((buckets - buckets.round()).abs().max()) < 1e-5 and float(buckets.min()) >= 2.0 - 1e-5
    del Y1, Y2, lhs, rhs, ones
    qkvs = torch.randn(N, 4 * H, generator=gen, device="cuda")
    c = torch.randn(H, generator=gen, device="cuda")
    qkvs[:, 2 * H:3 * H] = c                                        # every value row = c
    out = ops.edge_attention(qkvs, g, H, 0.1)
    want = c[None, :] + qkvs[:, 3 * H:]
    assert float((out - want).abs().max()) < 1e-5 * float(want.abs().max())   # sum_j alpha = 1 for every destination


def test_train_step_is_bit_reproducible_at_full_size(big):
    import erc_b200
    from erc_b200 import ops
    from erc_b200.track_mm import cogmen as cg
    N, lengths, spk = big["N"], big["lengths"], big["spk"]
    gen = torch.Generator(device="cuda").manual_seed(3)
    store = torch.empty((N, HIDDEN + 1), dtype=torch.float32, device="cuda").normal_(generator=gen)
    x = store[:, :HIDDEN]
    labels = torch.randint(0, 6, (N,), device="cuda", generator=gen)
    torch.manual_seed(0)
    model = cg.COGMENModule(HIDDEN, 100, 17, 2, 6, build_dead_encoder=False).cuda().train()

    def step():
        torch.manual_seed(123)                                      # same dropout seed (_fresh_seed draws from torch's CPU RNG)
        for p in model.parameters():
            p.grad = None
        logits, _ = model.forward_packed(x, spk, lengths, graph=big["g"])
        loss = ops.cross_entropy(logits, labels)
        loss.backward()
        return loss.detach().clone(), {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}

    l1, g1 = step()
    l2, g2 = step()
    assert torch.isfinite(l1) and torch.equal(l1, l2)
    assert g1.keys() == g2.keys() and len(g1) >= 15
    for k in g1:
        assert torch.isfinite(g1[k]).all() and torch.equal(g1[k], g2[k]), k
    w = g1["gcn.conv1.weight"]                                      # relation ids 2..7 never occur: exactly zero
    assert float(w[2:].abs().max()) == 0.0 and float(w[:2].abs().max()) > 0.0
