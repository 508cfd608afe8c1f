"""MMGCN through the drop-in modules: the block-adjacency kernels (K7), the GCNII layer
kernels (K8) and the whole MMGCNModule against the CPU oracle and the reference-generated fixture."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import copy

from conftest import rel_err, parity_check
from oracle import mmgcn_oracle, seeded
from oracle.make_golden import MMGCN_SEED, mmgcn_inputs
from test_oracle_mmgcn import check_against_fixture

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _layout(lengths, M):
    import erc_b200  # noqa: F401
    from erc_b200.ops_mmgcn import BlockLayout
    from erc_b200.track_mm.mmgcn_utils import lengths_graph
    lens = torch.tensor(lengths)
    g = lengths_graph(lens, torch.device("cuda"))
    return BlockLayout(g, M, lens)


@pytest.mark.parametrize("lengths,M,D", [([5, 1, 9, 3], 3, 200), ([1], 3, 8), ([37, 2, 64], 2, 100), ([110, 8], 3, 200)])
def test_block_adjacency_matches_dense_oracle(lengths, M, D):
    from erc_b200 import ops_mmgcn
    N = sum(lengths)
    gen = torch.Generator().manual_seed(5)
    feats = [torch.randn(N, D, generator=gen).requires_grad_() for _ in range(M)]
    want = mmgcn_oracle.big_adj(feats, lengths)
    cot = torch.randn(M * N, M * N, generator=gen)
    (want * cot).sum().backward()
    lay = _layout(lengths, M)
    x = torch.cat([f.detach() for f in feats], 0).cuda().requires_grad_()
    flat = ops_mmgcn.big_adj(x, lay)
    dense = lay.dense(flat.detach())
    assert rel_err(dense, want.detach()) < 2e-5    # acos at c ~ 0.99999 turns 1 ulp of a cosine into 7e-6 of the diagonal
    # cotangent restricted to the pattern, in the flat layout
    mask = lay.dense(torch.ones_like(flat.detach())) > 0
    assert torch.equal(mask.cpu(), want.detach() != 0)
    cflat = torch.zeros_like(flat.detach())
    idx = torch.arange(lay.nflat, device="cuda", dtype=torch.float32)
    pos = lay.dense(idx + 1).long() - 1                       # dense position -> flat index (or -1)
    sel = pos >= 0
    cflat[pos[sel]] = cot.cuda()[sel]
    flat.backward(cflat)
    got = x.grad.cpu()
    ref = torch.cat([f.grad for f in feats], 0)
    assert rel_err(got, ref) < 2e-4        # d acos at c = 0.99999 amplifies 1 ulp of a cosine by ~220


def test_spmm_and_sddmm_match_dense():
    from erc_b200 import ops_mmgcn
    lengths, M, H = [7, 33, 1, 12], 3, 200
    lay = _layout(lengths, M)
    N = sum(lengths)
    gen = torch.Generator().manual_seed(6)
    flat = torch.randn(lay.nflat, generator=gen).cuda()
    h = torch.randn(M * N, H, generator=gen).cuda()
    dense = lay.dense(flat).double()
    assert rel_err(ops_mmgcn.spmm(flat, h, lay), dense @ h.double()) < 1e-6
    assert rel_err(ops_mmgcn.spmm(flat, h, lay, transpose=True), dense.t() @ h.double()) < 1e-6
    dhi = torch.randn(M * N, H, generator=gen).cuda()
    G = torch.empty(lay.nflat, device="cuda")
    ops_mmgcn.sddmm(dhi, h, lay, G, accumulate=False)
    ops_mmgcn.sddmm(dhi, h, lay, G, accumulate=True)
    full = 2 * (dhi.double() @ h.double().t())
    mask = lay.dense(torch.ones(lay.nflat, device="cuda")) > 0
    assert rel_err(lay.dense(G), torch.where(mask, full, torch.zeros_like(full))) < 1e-6
    # the side accumulation of the backward
    acc = torch.ones(M * N, H, device="cuda")
    ops_mmgcn.spmm(flat, h, lay, acc_src=dhi, acc_dst=acc)
    assert rel_err(acc, 1 + dhi) < 1e-7


def test_graph_convolution_layer_vs_torch():
    """GraphConvolution.forward (variant) with autograd, against the dense formula of mmgcn_models.py:27-39."""
    from erc_b200 import ops_mmgcn
    from erc_b200.track_mm.mmgcn_models import GraphConvolution, BlockAdjacency
    lengths, M, H = [9, 4, 21], 3, 200
    lay = _layout(lengths, M)
    N = sum(lengths)
    gen = torch.Generator().manual_seed(7)
    flat = (torch.rand(lay.nflat, generator=gen) * 0.1).cuda().requires_grad_()
    h = torch.randn(M * N, H, generator=gen).cuda().requires_grad_()
    h0 = torch.randn(M * N, H, generator=gen).cuda().requires_grad_()
    conv = GraphConvolution(H, H, variant=True).cuda()
    out = conv(h, BlockAdjacency(flat, lay), h0, 0.5, 0.1, 3)
    cot = torch.randn(M * N, H, generator=gen).cuda()
    out.backward(cot)
    # dense fp64 restatement
    fd, hd, h0d, wd = (t.detach().double().requires_grad_() for t in (flat, h, h0, conv.weight))
    theta = math.log(0.5 / 3 + 1)
    hi = lay.dense(fd) @ hd
    ref = theta * (torch.cat([hi, h0d], 1) @ wd) + (1 - theta) * (0.9 * hi + 0.1 * h0d)
    ref.backward(cot.double())
    assert rel_err(out, ref) < 2e-6
    assert rel_err(h.grad, hd.grad) < 2e-6 and rel_err(h0.grad, h0d.grad) < 2e-6
    assert rel_err(conv.weight.grad, wd.grad) < 2e-6 and rel_err(flat.grad, fd.grad) < 2e-6


def test_simple_batch_graphify_and_speaker_embedding():
    from erc_b200 import ops_mmgcn
    from erc_b200.track_mm.mmgcn_utils import simple_batch_graphify, lengths_graph
    lengths = [4, 1, 6]
    b = mmgcn_inputs(lengths, (8, 8, 8), 6, seed=3)
    x = b["text_feature"].cuda()
    nodes, *rest = simple_batch_graphify(x, b["text_length"])
    assert rest == [None] * 4
    assert torch.equal(nodes.cpu(), mmgcn_oracle.simple_pack(b["text_feature"], lengths))
    g = lengths_graph(b["text_length"], x.device)
    rows = ops_mmgcn.node_rows(g, max(lengths), seq_first=True)
    emb = torch.randn(2, 8).cuda().requires_grad_()
    out, ids = ops_mmgcn.speaker_embed_add(nodes, b["speaker_tensor"].cuda(), rows, emb)
    q = torch.cat([b["speaker_tensor"][:L, i] for i, L in enumerate(lengths)], 0).argmax(-1)
    assert torch.equal(ids.cpu().long(), q)
    assert torch.equal(out.detach().cpu(), nodes.cpu() + emb.detach().cpu()[q])
    out.sum().backward()
    assert torch.equal(emb.grad.cpu(), torch.stack([(q == s).sum() * torch.ones(8) for s in range(2)]))


def _oracle_runs(dims, C, b):
    """MmgcnOracle (same name-seeded weights) in fp32 and in fp64 -> ((logits, grads) fp32, (logits, grads) fp64)."""
    o = mmgcn_oracle.MmgcnOracle(dims[0], dims[1], dims[2], n_classes=C, dropout=0.0)
    seeded.fill_by_name(o, MMGCN_SEED)
    res = []
    for dt in (torch.float32, torch.float64):
        oo = copy.deepcopy(o).to(dt)
        oo.train()
        kw = {k: (v.to(dt) if v.is_floating_point() else v) for k, v in b.items() if k != "label"}
        ol, _ = oo(**kw)
        F.cross_entropy(ol, b["label"]).backward()
        res.append((ol.detach().numpy(), {k: p.grad.numpy() for k, p in oo.named_parameters() if p.grad is not None}))
    return res


def _run_module(m, b):
    from erc_b200 import ops
    logits, none = m(**{k: v.cuda() if k != "text_length" else v for k, v in b.items() if k != "label"})
    assert none is None
    loss = ops.cross_entropy(logits, b["label"].cuda())
    loss.backward()
    return logits, loss


def test_mmgcn_module_vs_reference_fixture(golden):
    import erc_b200  # noqa: F401
    from erc_b200.track_mm.mmgcn import MMGCNModule
    fx = golden("mmgcn_small")
    dt, da, dv = (int(v) for v in fx["dims"])
    m = MMGCNModule(hidden_text=dt, hidden_audio=da, hidden_visual=dv, n_speakers=2, n_classes=fx["logits"].shape[1],
                    modals="atv")
    seeded.fill_by_name(m, MMGCN_SEED)          # same names as the reference module => same values as the fixture run
    m = m.cuda()
    m.lstm_l.dropout = 0.0
    m.graph_model.graph_net.dropout = 0.0
    m.dropout_.p = 0.0
    m.train()
    b = {k: torch.from_numpy(fx[k]) for k in ("text_feature", "audio_feature", "visual_feature", "speaker_tensor",
                                              "text_length", "label")}
    logits, loss = _run_module(m, b)
    assert rel_err(logits, fx["logits"]) < TOL
    assert abs(float(loss.detach()) - float(fx["loss"])) < TOL * float(fx["loss"])
    grads = {k: p.grad.cpu().numpy() for k, p in m.named_parameters() if p.grad is not None}
    # parameters that are mathematically dead get exact zeros here; the fixture lists only the live ones
    live = set(str(k) for k in fx["live"])
    assert live <= set(grads)
    worst = check_against_fixture(fx, {k: grads[k] for k in live}, 2e-4)
    print("mmgcn fixture: worst grad rel err", worst)
    # the 2e-4 bar above, backed by data (acos at c = 0.99999 amplifies fp32 rounding ~224x in the reference itself):
    (l32, g32), (l64, g64) = _oracle_runs((dt, da, dv), fx["logits"].shape[1], b)
    parity_check("mmgcn/fixture/logits", {"logits": logits}, {"logits": fx["logits"]}, {"logits": l64})
    parity_check("mmgcn/fixture/grads", {k: grads[k] for k in g32}, g32, g64)


def test_mmgcn_config3_shape_vs_oracle():
    """BASELINE config 3: MMGCN 6-way, B=16 IEMOCAP-shaped dialogues, text 768 / audio 100 / visual 512, 64 layers."""
    import erc_b200  # noqa: F401
    from erc_b200 import synth
    from erc_b200.track_mm.mmgcn import MMGCNModule
    gen = torch.Generator().manual_seed(0)
    lengths = [int(v) for v in synth.iemocap_lengths(16, gen)]
    b = mmgcn_inputs(lengths, (768, 100, 512), 6, seed=9)
    (ol, want), (l64, want64) = _oracle_runs((768, 100, 512), 6, b)
    ol = torch.from_numpy(ol)
    oloss = F.cross_entropy(ol, b["label"])
    m = MMGCNModule(hidden_text=768, hidden_audio=100, hidden_visual=512, n_speakers=2, n_classes=6, modals="atv")
    seeded.fill_by_name(m, MMGCN_SEED)
    m = m.cuda()
    m.lstm_l.dropout = 0.0
    m.graph_model.graph_net.dropout = 0.0
    m.dropout_.p = 0.0
    m.train()
    logits, loss = _run_module(m, b)
    assert rel_err(logits, ol.detach()) < TOL
    assert abs(float(loss.detach()) - float(oloss.detach())) < TOL * float(oloss.detach())
    got = {k: p.grad.cpu().numpy() for k, p in m.named_parameters() if p.grad is not None and k in want}
    parity_check("mmgcn/config3/logits", {"logits": logits}, {"logits": ol}, {"logits": l64})
    parity_check("mmgcn/config3/grads", got, want, want64)


def test_mmgcn_dropout_training_step_runs():
    """Training mode with every dropout on: finite loss / gradients, masks regenerated consistently in backward."""
    import erc_b200  # noqa: F401
    from erc_b200.track_mm.mmgcn import MMGCNModule
    b = mmgcn_inputs([6, 3, 11], (24, 10, 12), 6, seed=2)
    torch.manual_seed(0)
    m = MMGCNModule(hidden_text=24, hidden_audio=10, hidden_visual=12, n_speakers=2, n_classes=6, modals="atv").cuda()
    m.train()
    logits, loss = _run_module(m, b)
    assert torch.isfinite(loss) and logits.shape == (20, 6)
    for k, p in m.named_parameters():
        if p.grad is not None:
            assert torch.isfinite(p.grad).all(), k
