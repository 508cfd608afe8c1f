"""COGMEN end to end through the drop-in modules: CUDA path vs the reference-generated fixture and vs the
CPU oracle on the BASELINE config-1 shape (fp32, 1e-5 relative, max-norm per tensor)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import copy

from conftest import rel_err, parity_check
from oracle import modules as om

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _oracle_fp64(o, x, spk, lens, y, class_weight=None):
    """The same oracle in double precision: the truth against which both fp32 results are measured (conftest.parity_check)."""
    o64 = copy.deepcopy(o).double()
    o64.zero_grad()
    o64.train()
    ol, of = o64(x.double(), spk, lens)
    F.cross_entropy(ol, y, weight=None if class_weight is None else class_weight.double()).backward()
    return ({"logits": ol.detach().numpy(), "features": of.detach().numpy()},
            {k: p.grad.numpy() for k, p in o64.named_parameters() if p.grad is not None})


def _run_ours(m, x, spk, lens, y, class_weight=None):
    from erc_b200 import ops
    dev = torch.device("cuda")
    logits, feats = m(x.to(dev), spk.to(dev), lens)
    loss = ops.cross_entropy(logits, y.to(dev), class_weight)
    loss.backward()
    torch.cuda.synchronize()
    grads = {k: p.grad.cpu().numpy() for k, p in m.named_parameters() if p.grad is not None}
    return logits.detach().cpu(), feats.detach().cpu(), float(loss.detach()), grads


def test_cogmen_vs_reference_fixture(golden):
    import erc_b200
    from erc_b200.track_mm.cogmen import COGMENModule
    fx = golden("cogmen_small")
    D, C = fx["input_tensor"].shape[-1], fx["logits"].shape[1]
    params = {k[6:]: torch.from_numpy(v) for k, v in fx.items() if k.startswith("param/")}
    m = COGMENModule(D, 100, 17, 2, C).cuda()
    missing, unexpected = m.load_state_dict(params, strict=False)
    assert not unexpected and all(k.startswith("rnn.0") for k in missing)      # only the dead encoder is absent
    m.cls[2].p = 0.0
    m.train()
    x, spk, lens, y = (torch.from_numpy(fx[k]) for k in ("input_tensor", "speaker_tensor", "text_length", "label"))
    logits, feats, loss, grads = _run_ours(m, x, spk, lens, y)
    assert rel_err(logits, fx["logits"]) < TOL
    assert rel_err(feats, fx["features"]) < TOL
    assert abs(loss - float(fx["loss"])) < TOL * abs(float(fx["loss"]))
    want = {k[5:]: v for k, v in fx.items() if k.startswith("grad/")}
    assert not any(k.startswith("rnn.0") for k in grads)                          # dead encoder stays grad-free
    o = om.CogmenOracle(D, n_classes=C, dropout=0.0)
    om.load_live(o, params)
    out64, want64 = _oracle_fp64(o, x, spk, lens, y)
    parity_check("cogmen/fixture/outputs", {"logits": logits.numpy(), "features": feats.numpy()},
                 {"logits": fx["logits"], "features": fx["features"]}, out64)
    parity_check("cogmen/fixture/grads", grads, want, want64)
    assert rel_err(m.gcn.bn.running_mean.cpu(), fx["bn_running_mean"]) < TOL
    assert rel_err(m.gcn.bn.running_var.cpu(), fx["bn_running_var"]) < TOL
    assert int(m.gcn.bn.num_batches_tracked) == 1
    m.eval()
    with torch.no_grad():
        le = m(x.cuda(), spk.cuda(), lens)[0]
    assert rel_err(le.cpu(), fx["logits_eval"]) < TOL


def test_cogmen_config1_shape_vs_oracle():
    """BASELINE config 1: iemocap-cogmen-sbert-4 shape, atv = 1380, batch 32, 4-way."""
    import erc_b200
    from erc_b200.track_mm.cogmen import COGMENModule
    from erc_b200 import synth
    batch = synth.config1(seed=0)
    torch.manual_seed(0)
    o = om.CogmenOracle(1380, n_classes=4, dropout=0.0)
    o.train()
    ol, of = o(batch["input_tensor"], batch["speaker_tensor"], batch["text_length"])
    oloss = F.cross_entropy(ol, batch["label"])
    oloss.backward()
    m = COGMENModule(1380, 100, 17, 2, 4, build_dead_encoder=False).cuda()
    m.load_state_dict(o.state_dict(), strict=False)
    m.cls[2].p = 0.0
    m.train()
    logits, feats, loss, grads = _run_ours(m, batch["input_tensor"], batch["speaker_tensor"], batch["text_length"],
                                           batch["label"])
    assert rel_err(feats, of.detach()) < TOL
    assert rel_err(logits, ol.detach()) < TOL
    assert abs(loss - float(oloss)) < TOL * float(oloss)
    out64, want64 = _oracle_fp64(o, batch["input_tensor"], batch["speaker_tensor"], batch["text_length"], batch["label"])
    parity_check("cogmen/config1/outputs", {"logits": logits.numpy(), "features": feats.numpy()},
                 {"logits": ol.detach().numpy(), "features": of.detach().numpy()}, out64)
    parity_check("cogmen/config1/grads", grads, {k: p.grad.numpy() for k, p in o.named_parameters()}, want64)


@pytest.mark.parametrize("speaker", [0, 1])
def test_cogmen_one_speaker_batch_relation_census_vs_oracle(speaker):
    """MOSEI-shaped batches carry ONE speaker id (mosei_feature.py:211): only 2 of the 8 relation ids occur and RGCNConv
    transforms / gathers only those slots.  Logits, loss and every gradient (absent relations: exactly zero) must match
    the oracle, which evaluates all 8 relations."""
    import erc_b200
    from erc_b200.track_mm.cogmen import COGMENModule
    from erc_b200 import synth
    batch = synth.config1(seed=5, B=8)
    batch["speaker_tensor"] = torch.full_like(batch["speaker_tensor"], speaker)
    torch.manual_seed(0)
    o = om.CogmenOracle(1380, n_classes=4, dropout=0.0)
    o.train()
    ol, of = o(batch["input_tensor"], batch["speaker_tensor"], batch["text_length"])
    oloss = F.cross_entropy(ol, batch["label"])
    oloss.backward()
    m = COGMENModule(1380, 100, 17, 2, 4, build_dead_encoder=False).cuda()
    m.load_state_dict(o.state_dict(), strict=False)
    m.cls[2].p = 0.0
    m.train()
    logits, feats, loss, grads = _run_ours(m, batch["input_tensor"], batch["speaker_tensor"], batch["text_length"],
                                           batch["label"])
    assert rel_err(logits, ol.detach()) < TOL
    assert abs(loss - float(oloss)) < TOL * float(oloss)
    _, want64 = _oracle_fp64(o, batch["input_tensor"], batch["speaker_tensor"], batch["text_length"], batch["label"])
    parity_check("cogmen/one_speaker_%d/grads" % speaker, grads, {k: p.grad.numpy() for k, p in o.named_parameters()}, want64)
    gw = grads["gcn.conv1.weight"]
    live = [speaker * 6, speaker * 6 + 1]                      # ((s*2 + s)*2 + dir)
    assert all(np.abs(gw[r]).max() > 0 for r in live)
    assert all(np.abs(gw[r]).max() == 0 for r in range(8) if r not in live)


def test_cogmen_packed_entry_point_equals_padded():
    import erc_b200
    from erc_b200.track_mm.cogmen import COGMENModule
    from erc_b200 import synth, ops
    batch = synth.config1(seed=3, B=6)
    torch.manual_seed(1)
    m = COGMENModule(1380, 100, 17, 2, 4, build_dead_encoder=False).cuda()
    m.eval()
    x, spk, lens = batch["input_tensor"], batch["speaker_tensor"], batch["text_length"]
    mask = torch.arange(x.size(1))[None, :] < lens[:, None]
    with torch.no_grad():
        a = m(x.cuda(), spk.cuda(), lens)[0]
        b = m.forward_packed(x[mask].contiguous().cuda(), spk[mask].contiguous().cuda(), lens)[0]
    assert rel_err(b, a) < 1e-5       # padded path gathers rows in the SIMT GEMM, packed path runs the tcgen05 GEMM


def test_cogmen_generic_edge_index_path_matches_attached_graph():
    """Layers handed a plain edge_index (no attached CSR, shuffled like the reference's set order)."""
    import erc_b200
    from erc_b200.track_mm.cogmen import GNN
    from erc_b200.graph import build_graph
    torch.manual_seed(2)
    lens = torch.tensor([9, 4, 17])
    spk = torch.randint(0, 2, (3, 17))
    g = build_graph(lens, spk.cuda(), 5, 5, 2)
    gnn = GNN(100, 100, 100).cuda().eval()
    x = torch.randn(g.N, 100).cuda()
    ei = g.edge_index.clone()
    ei._ercg_graph = g
    with torch.no_grad():
        a = gnn(x, ei, g.edge_type)
        perm = torch.randperm(g.E).cuda()
        b = gnn(x, g.edge_index[:, perm].contiguous(), g.edge_type[perm].contiguous())
    assert rel_err(b.cpu(), a.cpu()) < 1e-6


def test_cogmen_large_padded_batch_uses_tensor_core_projection_and_matches_packed():
    """Reference-layout (padded) input big enough for the tensor-core projection over the padded rows + row packing."""
    import erc_b200
    from erc_b200 import _lib, ops
    from erc_b200.track_mm.cogmen import COGMENModule
    from erc_b200 import synth
    gen = torch.Generator().manual_seed(7)
    lens = torch.randint(60, 111, (96,), generator=gen)
    batch = synth.padded_batch(lens, 1380, 2, 4, gen)
    torch.manual_seed(1)
    m = COGMENModule(1380, 100, 17, 2, 4, build_dead_encoder=False).cuda()
    m.cls[2].p = 0.0
    m.train()
    x, spk = batch["input_tensor"].cuda(), batch["speaker_tensor"].cuda()
    with _lib.KernelTimer() as kt:
        logits, _ = m(x, spk, lens)
        loss = ops.cross_entropy(logits, batch["label"].cuda())
        loss.backward()
    assert "ercg_gemm_nn_tc" in kt.summary() and "ercg_pack_rows" in kt.summary()
    ga = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
    m.zero_grad()
    mask = torch.arange(x.size(1))[None, :] < lens[:, None]
    lg2, _ = m.forward_packed(batch["input_tensor"][mask].contiguous().cuda(), batch["speaker_tensor"][mask].contiguous().cuda(), lens)
    ops.cross_entropy(lg2, batch["label"].cuda()).backward()
    assert rel_err(logits, lg2) < 1e-5
    scale = max(float(p.grad.abs().max()) for p in m.parameters() if p.grad is not None)
    for k, p in m.named_parameters():
        if p.grad is not None:
            assert rel_err(ga[k], p.grad, floor=1e-2 * scale) < 1e-5, k
