"""DialogueGCN through the drop-in modules: CUDA path vs the reference-generated fixture and the CPU oracle."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import copy

from conftest import rel_err, parity_check
from oracle import modules as om

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _oracle_fp64(o, x, spk, lens, y, w):
    """DgcnOracle in double precision = the truth for conftest.parity_check."""
    o64 = copy.deepcopy(o).double()
    o64.zero_grad()
    o64.train()
    ol, og = o64(x.double(), spk, lens)
    F.cross_entropy(ol, y, weight=w.double()).backward()
    return ({"logits": ol.detach().numpy(), "graph_out": og.detach().numpy()},
            {k: p.grad.numpy() for k, p in o64.named_parameters() if p.grad is not None})


def _params(fx):
    return {k[6:]: torch.from_numpy(v) for k, v in fx.items() if k.startswith("param/")}


def test_lstm_seqcontext_vs_fixture(golden):
    import erc_b200
    from erc_b200.track_mm.dgcn_models import SeqContext
    fx = golden("dgcn_small")
    D, H = fx["input_tensor"].shape[-1], fx["context"].shape[-1]
    m = SeqContext(D, H, dropout=0.0).cuda()
    m.load_state_dict({k[4:]: v for k, v in _params(fx).items() if k.startswith("rnn.")})
    out = m(torch.from_numpy(fx["text_length"]), torch.from_numpy(fx["input_tensor"]).cuda())
    assert out.shape == fx["context"].shape
    assert rel_err(out, fx["context"]) < TOL


def test_edgeatt_and_graphify_vs_fixture(golden):
    import erc_b200
    from erc_b200.track_mm.dgcn_models import EdgeAtt, batch_graphify
    from erc_b200.graph import standard_edge_dict
    fx = golden("dgcn_small")
    H = fx["context"].shape[-1]
    att = EdgeAtt(H, 10, 10).cuda()
    att.load_state_dict({"weight": torch.from_numpy(fx["param/edge_att.weight"])})
    ctx = torch.from_numpy(fx["context"]).cuda()
    lens, spk = torch.from_numpy(fx["text_length"]), torch.from_numpy(fx["speaker_tensor"]).cuda()
    nf, ei, en, et, el = batch_graphify(ctx, lens, spk, 10, 10, standard_edge_dict(2), att)
    assert np.array_equal(ei.cpu().numpy(), fx["edge_index"]) and np.array_equal(et.cpu().numpy(), fx["edge_type"])
    assert np.array_equal(el.cpu().numpy(), fx["edge_index_lengths"])
    assert rel_err(nf, fx["node_features"]) < 1e-7
    assert rel_err(en, fx["edge_norm"]) < TOL
    # reference-signature EdgeAtt.forward: list of [Lmax,110]
    alphas = att(ctx, lens, None)
    assert len(alphas) == ctx.size(0) and alphas[0].shape == (ctx.size(1), 110)
    src, dst = fx["edge_index"]
    off = np.concatenate([[0], np.cumsum(fx["text_length"])])
    d = np.searchsorted(off, src, side="right") - 1
    got = torch.stack(alphas).detach().cpu().numpy()[d, src - off[d], dst - off[d]]
    assert rel_err(got, fx["edge_norm"]) < TOL


def test_vendored_rgcn_vs_fixture(golden):
    import erc_b200
    from erc_b200.models.rgcn import RGCNConv
    fx = golden("dgcn_small")
    H = fx["context"].shape[-1]
    conv = RGCNConv(H, 100, 8, num_bases=30).cuda()
    conv.load_state_dict({k[len("gcn.conv1."):]: v for k, v in _params(fx).items() if k.startswith("gcn.conv1.")})
    x = torch.from_numpy(fx["node_features"]).cuda()
    # a plain (shuffled) edge_index without attached CSR, like a caller outside our graphify
    perm = torch.randperm(fx["edge_type"].shape[0], generator=torch.Generator().manual_seed(0))
    ei = torch.from_numpy(fx["edge_index"])[:, perm].cuda()
    et = torch.from_numpy(fx["edge_type"])[perm].cuda()
    en = torch.from_numpy(fx["edge_norm"])[perm].cuda()
    out = conv(x, ei, et, edge_norm=en)
    assert rel_err(out, fx["rgcn_out"]) < TOL
    with pytest.raises(ValueError):
        conv(x[:-1], ei, et, edge_norm=en)


def test_dgcn_module_vs_reference_fixture(golden):
    import erc_b200
    from erc_b200 import ops
    from erc_b200.track_mm.dgcn import DGCNModule
    fx = golden("dgcn_small")
    D, H = fx["input_tensor"].shape[-1], fx["context"].shape[-1]
    m = DGCNModule(2, input_size=D, hidden_size=H, n_classes=6).cuda()
    m.load_state_dict(_params(fx), strict=True)              # every reference key, dead ones included
    m.rnn.rnn.dropout = 0.0
    m.clf.drop.p = 0.0
    m.train()
    x, spk, lens, y = (torch.from_numpy(fx[k]) for k in ("input_tensor", "speaker_tensor", "text_length", "label"))
    logits, graph_out = m(x.cuda(), spk.cuda(), lens)
    loss = ops.cross_entropy(logits, y.cuda(), torch.from_numpy(fx["class_weights"]).cuda())
    loss.backward()
    assert rel_err(logits, fx["logits"]) < TOL
    assert rel_err(graph_out, fx["graph_out"]) < TOL
    assert abs(float(loss.detach()) - float(fx["loss"])) < TOL * float(fx["loss"])
    grads = {k: p.grad.cpu().numpy() for k, p in m.named_parameters() if p.grad is not None}
    o = om.DgcnOracle(2, input_size=D, hidden_size=H, n_classes=6, dropout=0.0)
    om.load_live(o, _params(fx))
    out64, want64 = _oracle_fp64(o, x, spk, lens, y, torch.from_numpy(fx["class_weights"]))
    parity_check("dgcn/fixture/outputs", {"logits": logits, "graph_out": graph_out},
                 {"logits": fx["logits"], "graph_out": fx["graph_out"]}, out64)
    parity_check("dgcn/fixture/grads", grads, {k[5:]: v for k, v in fx.items() if k.startswith("grad/")}, want64)


def test_dgcn_config2_shape_vs_oracle():
    """BASELINE config 2: DialogueGCN 6-way, IEMOCAP-shaped synthetic batch of 32, window 10/10, class weights."""
    import erc_b200
    from erc_b200 import ops, synth
    from erc_b200.track_mm.dgcn import DGCNModule, LOSS_WEIGHTS
    batch = synth.config2(seed=0)
    torch.manual_seed(0)
    o = om.DgcnOracle(2, input_size=1380, hidden_size=200, n_classes=6, dropout=0.0)
    o.train()
    w = torch.tensor(LOSS_WEIGHTS)
    ol, og = o(batch["input_tensor"], batch["speaker_tensor"], batch["text_length"])
    oloss = F.cross_entropy(ol, batch["label"], weight=w)
    oloss.backward()
    m = DGCNModule(2, input_size=1380, hidden_size=200, n_classes=6).cuda()
    missing, unexpected = m.load_state_dict(o.state_dict(), strict=False)
    assert not unexpected and all(k.startswith("clf.emotion_att") for k in missing)
    m.rnn.rnn.dropout = 0.0
    m.clf.drop.p = 0.0
    m.train()
    logits, graph_out = m(batch["input_tensor"].cuda(), batch["speaker_tensor"].cuda(), batch["text_length"])
    loss = ops.cross_entropy(logits, batch["label"].cuda(), w.cuda())
    loss.backward()
    assert rel_err(graph_out, og.detach()) < TOL
    assert rel_err(logits, ol.detach()) < TOL
    assert abs(float(loss.detach()) - float(oloss.detach())) < TOL * float(oloss.detach())
    grads = {k: p.grad.cpu().numpy() for k, p in m.named_parameters() if p.grad is not None}
    out64, want64 = _oracle_fp64(o, batch["input_tensor"], batch["speaker_tensor"], batch["text_length"], batch["label"], w)
    parity_check("dgcn/config2/outputs", {"logits": logits, "graph_out": graph_out},
                 {"logits": ol.detach(), "graph_out": og.detach()}, out64)
    parity_check("dgcn/config2/grads", grads, {k: p.grad.numpy() for k, p in o.named_parameters()}, want64)


def test_dropout_kernel_is_reproducible_and_unbiased():
    import erc_b200
    from erc_b200 import ops
    x = torch.ones(1 << 20, device="cuda", requires_grad=True)
    a, b = ops.dropout(x, 0.4, 77), ops.dropout(x, 0.4, 77)
    assert torch.equal(a, b)
    assert abs(float(a.mean()) - 1.0) < 5e-3
    assert abs(float((a == 0).float().mean()) - 0.4) < 5e-3
    a.sum().backward()
    assert torch.equal(x.grad, a.detach())
