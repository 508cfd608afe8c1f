"""DAG-ERC through the drop-in module: K9 predecessor structure bit-exact against the
reference's get_adj_v1 / get_s_mask, K10 layer kernel + module against the reference-generated fixture and the CPU oracle."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err, check_grads
from oracle import dagerc_oracle, seeded
from oracle.make_golden import DAGERC_SEED, dagerc_inputs
from test_oracle_mmgcn import check_against_fixture

pytestmark = pytest.mark.gpu
TOL = 1e-5


def test_dense_masks_bit_exact_vs_reference_fixture(golden):
    import erc_b200  # noqa: F401
    from erc_b200.track_mm.dagerc import DAGERCModule
    fx = golden("dagerc_small")
    m = DAGERCModule(emb_dim=fx["input_tensor"].shape[-1], dropout=0.0, n_classes=6).cuda()
    L = fx["input_tensor"].shape[1]
    adj = m.get_adj_v1(fx["speaker_tensor"].tolist(), L)           # nested lists, as the reference passes them
    sm, onehot = m.get_s_mask(torch.from_numpy(fx["speaker_tensor"]), L)
    assert adj.dtype == torch.float32 and sm.dtype == torch.int64
    assert np.array_equal(adj.cpu().numpy(), fx["adj"]) and np.array_equal(sm.cpu().numpy(), fx["s_mask"])
    assert onehot.shape == sm.shape + (2,)


@pytest.mark.parametrize("windowp", [1, 2])
def test_dag_structure_random_speakers(windowp):
    import erc_b200  # noqa: F401
    from erc_b200 import ops_dagerc
    from erc_b200.graph import build_graph
    gen = torch.Generator().manual_seed(3)
    lengths = torch.tensor([1, 17, 110, 2, 64, 5000 // 100])
    B, Lmax = lengths.numel(), int(lengths.max())
    spk = torch.randint(0, 3, (B, Lmax), generator=gen)
    g = build_graph(lengths, spk.cuda(), 0, 0, 3, reference_layout=False, mean_weight=False)
    dag = ops_dagerc.DagStructure(g, lengths, windowp)
    want = dagerc_oracle.adj_v1(spk.numpy(), windowp)
    lo, cnt, eoff = dag.lo.cpu().numpy(), dag.cnt.cpu().numpy(), dag.eoff.cpu().numpy()
    off = np.concatenate([[0], np.cumsum(lengths.numpy())])
    for b in range(B):
        for i in range(int(lengths[b])):
            row = want[b, i, :int(lengths[b])]
            nz = np.nonzero(row)[0]
            n = off[b] + i
            assert cnt[n] == len(nz)
            if len(nz):
                assert lo[n] == nz[0] and nz[-1] == i - 1 and len(nz) == i - nz[0]
    assert np.array_equal(eoff, np.concatenate([[0], np.cumsum(cnt)[:-1]])) and dag.E == int(cnt.sum())
    dense, sm = ops_dagerc.dense_masks(spk.cuda(), windowp)
    assert np.array_equal(dense.cpu().numpy(), want) and np.array_equal(sm.cpu().numpy(), dagerc_oracle.s_mask(spk.numpy()))


def _run(m, b):
    from erc_b200 import ops
    logits, none = m(input_tensor=b["input_tensor"].cuda(), text_length=b["text_length"],
                     speaker_tensor=b["speaker_tensor"].cuda())
    assert none is None
    sel = logits[b["attention_mask"].cuda().bool()]                 # dagerc.py:223-226
    loss = ops.cross_entropy(sel, b["label"].cuda())
    loss.backward()
    return logits, loss


def test_dagerc_module_vs_reference_fixture(golden):
    import erc_b200  # noqa: F401
    from erc_b200.track_mm.dagerc import DAGERCModule
    fx = golden("dagerc_small")
    emb, C = fx["input_tensor"].shape[-1], fx["logits"].shape[-1]
    m = DAGERCModule(emb_dim=emb, dropout=0.0, n_classes=C, gnn_layers=4)
    seeded.fill_by_name(m, DAGERC_SEED)
    m = m.cuda()
    m.train()
    b = {k: torch.from_numpy(fx[k]) for k in ("input_tensor", "speaker_tensor", "text_length", "attention_mask", "label")}
    logits, loss = _run(m, b)
    assert logits.shape == fx["logits"].shape
    assert rel_err(logits, fx["logits"]) < TOL                      # padded positions included
    assert abs(float(loss.detach()) - float(fx["loss"])) < TOL * float(fx["loss"])
    grads = {k: p.grad.cpu().numpy() for k, p in m.named_parameters() if p.grad is not None}
    live = set(str(k) for k in fx["live"])
    assert live <= set(grads), sorted(live - set(grads))
    worst = check_against_fixture(fx, {k: grads[k] for k in live}, 1e-4)
    print("dagerc fixture: worst grad rel err", worst)
    # packed mode: same logits at the real positions, zeros at the padding, same gradients
    m2 = DAGERCModule(emb_dim=emb, dropout=0.0, n_classes=C, gnn_layers=4, compute_padding=False)
    seeded.fill_by_name(m2, DAGERC_SEED)
    m2 = m2.cuda()
    m2.train()
    logits2, loss2 = _run(m2, b)
    mask = b["attention_mask"].bool()
    assert rel_err(logits2[mask.cuda()], logits[mask.cuda()]) < 1e-6 and float(logits2[~mask.cuda()].abs().max()) == 0.0
    for k, p in m2.named_parameters():
        if p.grad is not None:
            assert rel_err(p.grad, grads[k], floor=1e-8) < 1e-5, k


def test_dagerc_config4_shape_vs_oracle():
    """BASELINE config 4: DAG-ERC 6-way, B=16 IEMOCAP-shaped dialogues, hidden_all 1380, 4 layers, dropout 0."""
    import erc_b200  # noqa: F401
    from erc_b200 import synth
    from erc_b200.track_mm.dagerc import DAGERCModule
    gen = torch.Generator().manual_seed(1)
    lengths = [int(v) for v in synth.iemocap_lengths(16, gen)]
    b = dagerc_inputs(lengths, 1380, 6, seed=11)
    o = dagerc_oracle.DagercOracle(1380, n_classes=6, dropout=0.0)
    seeded.fill_by_name(o, DAGERC_SEED)
    o.train()
    ol, _ = o(b["input_tensor"], b["text_length"], b["speaker_tensor"])
    oloss = F.cross_entropy(ol[b["attention_mask"].bool()], b["label"])
    oloss.backward()
    m = DAGERCModule(emb_dim=1380, dropout=0.0, n_classes=6, gnn_layers=4)
    seeded.fill_by_name(m, DAGERC_SEED)
    m = m.cuda()
    m.train()
    logits, loss = _run(m, b)
    assert rel_err(logits, ol.detach()) < TOL
    assert abs(float(loss.detach()) - float(oloss.detach())) < TOL * float(oloss.detach())
    want = {k: p.grad.numpy() for k, p in o.named_parameters() if p.grad is not None}
    got = {k: p.grad.cpu().numpy() for k, p in m.named_parameters() if p.grad is not None and k in want}
    worst = check_grads(got, want, 1e-4)
    print("dagerc config 4: worst grad rel err", worst)


def test_dagerc_dropout_training_step_runs():
    import erc_b200  # noqa: F401
    from erc_b200.track_mm.dagerc import DAGERCModule
    b = dagerc_inputs([5, 9, 1], 24, 6, seed=2)
    torch.manual_seed(0)
    m = DAGERCModule(emb_dim=24, dropout=0.3, n_classes=6, gnn_layers=2).cuda()
    m.train()
    logits, loss = _run(m, b)
    assert torch.isfinite(loss) and logits.shape == (3, 9, 6)
    assert all(torch.isfinite(p.grad).all() for p in m.parameters() if p.grad is not None)
