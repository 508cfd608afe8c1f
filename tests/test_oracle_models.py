"""oracle/modules.py (vectorised restatement) against fixtures produced by the reference's own module code."""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import modules as om
from conftest import rel_err, check_grads

TOL = 1e-5   # fp32 relative, max-norm per tensor (BASELINE.json north_star)


def _params(fx):
    return {k[len("param/"):]: v for k, v in fx.items() if k.startswith("param/")}


def _grads(fx):
    return {k[len("grad/"):]: v for k, v in fx.items() if k.startswith("grad/")}


def test_cogmen_oracle_matches_reference_fixture(golden):
    fx = golden("cogmen_small")
    D = fx["input_tensor"].shape[-1]
    m = om.CogmenOracle(D, n_classes=fx["logits"].shape[1], dropout=0.0)
    skipped = om.load_live(m, _params(fx))
    assert skipped == []
    m.train()
    logits, feats = m(torch.from_numpy(fx["input_tensor"]), torch.from_numpy(fx["speaker_tensor"]),
                      torch.from_numpy(fx["text_length"]))
    loss = F.cross_entropy(logits, torch.from_numpy(fx["label"]))
    loss.backward()
    assert rel_err(logits.detach(), fx["logits"]) < TOL
    assert rel_err(feats.detach(), fx["features"]) < TOL
    assert abs(float(loss.detach()) - float(fx["loss"])) < TOL * abs(float(fx["loss"]))
    got = {k: p.grad.numpy() for k, p in m.named_parameters()}
    want = _grads(fx)
    check_grads(got, want, 5 * TOL)
    assert rel_err(m.gcn.bn.running_mean, fx["bn_running_mean"]) < TOL
    assert rel_err(m.gcn.bn.running_var, fx["bn_running_var"]) < TOL
    m.eval()
    with torch.no_grad():
        le = m(torch.from_numpy(fx["input_tensor"]), torch.from_numpy(fx["speaker_tensor"]),
               torch.from_numpy(fx["text_length"]))[0]
    assert rel_err(le, fx["logits_eval"]) < TOL


def test_dgcn_oracle_matches_reference_fixture(golden):
    fx = golden("dgcn_small")
    D = fx["input_tensor"].shape[-1]
    H = fx["context"].shape[-1]
    m = om.DgcnOracle(2, input_size=D, hidden_size=H, n_classes=6, dropout=0.0)
    skipped = om.load_live(m, _params(fx))
    assert sorted(skipped) == ["clf.emotion_att.lin.bias", "clf.emotion_att.lin.weight"]   # dead in the reference
    m.train()
    x, spk, lens = (torch.from_numpy(fx[k]) for k in ("input_tensor", "speaker_tensor", "text_length"))
    ctx = m.rnn(lens, x)
    assert rel_err(ctx.detach(), fx["context"]) < TOL
    feats = om.pack_nodes(ctx, lens)
    ei, et = torch.from_numpy(fx["edge_index"]), torch.from_numpy(fx["edge_type"])
    norm = m.edge_att(feats, ei)
    assert rel_err(norm.detach(), fx["edge_norm"]) < TOL
    rg = m.gcn.conv1(feats, ei, et, norm)
    assert rel_err(rg.detach(), fx["rgcn_out"]) < TOL
    logits, graph_out = m(x, spk, lens)
    loss = F.cross_entropy(logits, torch.from_numpy(fx["label"]), weight=torch.from_numpy(fx["class_weights"]))
    loss.backward()
    assert rel_err(logits.detach(), fx["logits"]) < TOL
    assert rel_err(graph_out.detach(), fx["graph_out"]) < TOL
    got = {k: p.grad.numpy() for k, p in m.named_parameters()}
    want = _grads(fx)
    check_grads(got, want, 5 * TOL)
