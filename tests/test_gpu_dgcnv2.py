"""the declare-lab DialogueGCN variant (track_mm/dgcnv2.py, dgcnv2_models.py) through the drop-in classes:
K11 edge weights, batch_graphify, the module against the reference-generated fixture and the CPU oracle (fp32 + fp64)."""
import copy

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err, parity_check
from oracle import dgcnv2_oracle, seeded, graph_np
from oracle.make_golden import DGCNV2_SEED
from test_oracle_mmgcn import check_against_fixture

pytestmark = pytest.mark.gpu
W = [1 / 0.086747, 1 / 0.144406, 1 / 0.227883, 1 / 0.160585, 1 / 0.127711, 1 / 0.252668]


def _run(m, b, w):
    from erc_b200 import ops
    logits, feats = m(input_tensor=b["input_tensor"].cuda(), speaker_tensor=b["speaker_tensor"].cuda(),
                      attention_mask=b["attention_mask"].cuda(), text_length=b["text_length"])
    loss = ops.cross_entropy(logits, b["label"].cuda(), w.cuda())
    loss.backward()
    return logits, feats, loss


def test_dgcnv2_module_vs_reference_fixture(golden):
    import erc_b200  # noqa: F401
    from erc_b200.track_mm.dgcnv2 import DGCNModule
    fx = golden("dgcnv2_small")
    D, C = fx["input_tensor"].shape[-1], fx["logits"].shape[1]
    m = DGCNModule("LSTM", input_size=D, hidden_size=100, n_speakers=2, n_classes=C, dropout=0.0)
    m.graph_net.dropout.p = 0.0
    seeded.fill_by_name(m, DGCNV2_SEED)                       # same parameter names as the reference module => same values
    m = m.cuda().train()
    b = {k: torch.from_numpy(fx[k]) for k in ("input_tensor", "speaker_tensor", "attention_mask", "text_length", "label")}
    logits, feats, loss = _run(m, b, torch.from_numpy(fx["class_weights"]))
    assert rel_err(logits, fx["logits"]) < 1e-5 and rel_err(feats, fx["features"]) < 1e-5
    assert abs(float(loss.detach()) - float(fx["loss"])) < 1e-5 * float(fx["loss"])
    live = set(str(k) for k in fx["live"])
    grads = {k: p.grad.cpu().numpy() for k, p in m.named_parameters() if p.grad is not None}
    assert live <= set(grads), sorted(live - set(grads))
    check_against_fixture(fx, {k: grads[k] for k in live}, 5e-5)


def test_dgcnv2_iemocap_shape_vs_oracle_fp64():
    """6-way, IEMOCAP-shaped batch of 16 dialogues (hidden_all 1380), window 10/10, class weights (dgcnv2.py:200-203)."""
    import erc_b200  # noqa: F401
    from erc_b200 import synth
    from erc_b200.track_mm.dgcnv2 import DGCNModule
    lengths = [int(v) for v in synth.iemocap_lengths(16, torch.Generator().manual_seed(3))]
    b = dgcnv2_oracle.inputs(lengths, 1380, 2, 6, seed=5)
    w = torch.tensor(W)
    o = dgcnv2_oracle.Dgcnv2Oracle(1380, n_classes=6, dropout=0.0)
    seeded.fill_by_name(o, DGCNV2_SEED)
    res = []
    for dt in (torch.float32, torch.float64):
        oo = copy.deepcopy(o).to(dt).train()
        lg, ft = oo(b["input_tensor"].to(dt), b["speaker_tensor"].to(dt), b["attention_mask"], b["text_length"])
        F.cross_entropy(lg, b["label"], weight=w.to(dt)).backward()
        res.append(({"logits": lg.detach().numpy(), "features": ft.detach().numpy()},
                    {k: p.grad.numpy() for k, p in oo.named_parameters() if p.grad is not None}))
    m = DGCNModule("LSTM", input_size=1380, hidden_size=100, n_speakers=2, n_classes=6, dropout=0.0)
    m.graph_net.dropout.p = 0.0
    seeded.fill_by_name(m, DGCNV2_SEED)
    m = m.cuda().train()
    logits, feats, loss = _run(m, b, w)
    parity_check("dgcnv2/iemocap16/outputs", {"logits": logits, "features": feats}, res[0][0], res[1][0])
    grads = {k: p.grad.cpu().numpy() for k, p in m.named_parameters() if p.grad is not None and k in res[0][1]}
    parity_check("dgcnv2/iemocap16/grads", grads, res[0][1], res[1][1])


def test_dgcnv2_batch_graphify_and_dense_edge_attention_signatures():
    """Reference signatures: batch_graphify(features [L,B,D], qmask one-hot, lengths, wp, wf, mapping, att_model) and
    MaskedEdgeAttention.forward(M, lengths, edge_ind) -> dense [B, 110, L] scores whose window entries sum to ~1 per source."""
    import erc_b200  # noqa: F401
    from erc_b200.graph import standard_edge_dict
    from erc_b200.track_mm.dgcnv2_models import MaskedEdgeAttention, batch_graphify
    lengths = [7, 1, 12]
    b = dgcnv2_oracle.inputs(lengths, 40, 2, 6, seed=9)
    att = MaskedEdgeAttention(40, 110).cuda()
    feats = b["input_tensor"].cuda()
    nf, ei, en, et, el = batch_graphify(feats, b["speaker_tensor"].cuda(), b["text_length"], 3, 2, standard_edge_dict(2), att)
    g = graph_np.batch_graphify_np(np.asarray(lengths), b["speaker_tensor"].argmax(-1).t().numpy(), 3, 2, 2)
    assert np.array_equal(ei.cpu().numpy(), g["edge_index"]) and np.array_equal(et.cpu().numpy(), g["edge_type"])
    assert el == [int(v) for v in g["edge_index_lengths"]] and nf.shape == (sum(lengths), 40)
    L, B = feats.shape[0], feats.shape[1]
    alpha = torch.softmax(att.scalar(feats), 0).permute(1, 2, 0)
    mask = torch.full_like(alpha, 1e-10)
    off = torch.from_numpy(g["node_off"]).cuda()
    bb = torch.from_numpy(g["dlg"]).long().cuda()[ei[0]]
    mask[bb, ei[0] - off[bb], ei[1] - off[bb]] = 1
    want = (alpha * mask / (alpha * mask).sum(-1, keepdim=True))[bb, ei[0] - off[bb], ei[1] - off[bb]]
    assert rel_err(en, want) < 1e-5
    dense = att(feats, b["text_length"], None)
    assert dense.shape == (B, 110, L) and rel_err(dense[bb, ei[0] - off[bb], ei[1] - off[bb]], want) < 1e-5
    assert float(dense.sum()) == pytest.approx(float(want.sum()), rel=1e-5)


def test_dgcnv2_training_step_with_dropout_runs():
    import erc_b200  # noqa: F401
    from erc_b200.track_mm.dgcnv2 import DGCNModule
    b = dgcnv2_oracle.inputs([6, 3, 11], 24, 2, 6, seed=2)
    torch.manual_seed(0)
    m = DGCNModule("LSTM", input_size=24, n_classes=6).cuda().train()
    logits, feats, loss = _run(m, b, torch.tensor(W))
    assert torch.isfinite(loss) and logits.shape == (20, 6) and feats.shape == (20, 200)
    assert all(torch.isfinite(p.grad).all() for p in m.parameters() if p.grad is not None)
    m2 = DGCNModule("None", input_size=24, n_classes=6).cuda().train()
    l2, f2, loss2 = _run(m2, b, torch.tensor(W))
    assert torch.isfinite(loss2) and l2.shape == (20, 6)
