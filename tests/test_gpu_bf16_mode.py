"""bf16 INPUT-FEATURE MODE (north_star: "stated separately for bf16"; include/ercgraph.h ercg_gemm_*_tc_bf16a): the utterance
features are stored in bf16, everything else stays fp32.

Stated tolerances (max-norm relative, per tensor):
  * kernels vs fp64 on the SAME bf16-rounded features: 2e-5 (the forward product sees the weights with 16 significant bits);
  * COGMEN train step (logits / loss / all live gradients) in bf16-input mode vs the fp32 oracle fed the bf16-rounded features:
    the fp32 bars of conftest.parity_check (1e-5, or within 4x the fp32 oracle's own distance to fp64);
  * the mode itself vs the fp32 path on the ORIGINAL features: 1e-2 on logits (bf16 keeps 8 significant bits of every
    feature; with hidden_all = 1443 terms per dot product the error averages down to a few 1e-3)."""
import copy

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err, parity_check, grads_of
from oracle import modules as om

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,K,N", [(4096, 1443, 100), (70001, 1443, 100), (1500, 100, 100), (2048, 64, 16), (5000, 1380, 100), (1111, 333, 128)])
def test_bf16_input_gemms_vs_fp64(M, K, N):
    import erc_b200  # noqa: F401
    from erc_b200 import ops, synth
    g = torch.Generator().manual_seed(M + K)
    x = torch.randn(M, K, generator=g)
    W = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    dY = torch.randn(M, N, generator=g)
    xb = synth.to_bf16_rows(x.cuda())
    xr = xb.float().cpu().double()                               # the stored values, exactly
    Wc, bc = W.cuda().requires_grad_(), b.cuda().requires_grad_()
    y = ops.linear(xb, Wc, bc)
    y.backward(dY.cuda())
    want = xr @ W.double().t() + b.double()
    assert rel_err(y, want) < 2e-5
    assert rel_err(Wc.grad, dY.double().t() @ xr) < 2e-5
    assert rel_err(bc.grad, dY.double().sum(0)) < 1e-5


def test_cogmen_bf16_input_mode_vs_oracle():
    import erc_b200  # noqa: F401
    from erc_b200 import ops, synth, _lib
    from erc_b200.track_mm.cogmen import COGMENModule
    lengths = synth.config5_lengths(1 << 14, seed=2)
    b = synth.packed_batch(lengths, 1443, 2, 6, torch.Generator().manual_seed(12), one_speaker=True)
    torch.manual_seed(2)
    o = om.CogmenOracle(1443, n_classes=6, dropout=0.0)
    m = COGMENModule(1443, 100, 17, 2, 6, build_dead_encoder=False).cuda()
    m.load_state_dict(o.state_dict(), strict=False)
    m.cls[2].p = 0.0
    m.train()
    xb = synth.to_bf16_rows(b["x_packed"].cuda())
    seen = {}
    hook = m.gcn.register_forward_hook(lambda mod, inp, out: seen.__setitem__("graph_out", out.detach()))
    with _lib.KernelTimer() as kt:
        logits, feats = m.forward_packed(xb, b["speaker_packed"].cuda(), b["text_length"])
        loss = ops.cross_entropy(logits, b["label"].cuda())
        loss.backward()
    hook.remove()
    names = kt.summary()
    assert "ercg_gemm_nn_tc_bf16a" in names and "ercg_gemm_tn_tc_bf16a" in names
    lin0 = m.cls[0]
    hid = ops.linear(seen["graph_out"], lin0.weight.detach(), lin0.bias.detach(), act=ops.ACT_RELU)
    masks = ((seen["graph_out"] > 0).cpu(), (hid > 0).cpu())
    x_rounded = xb.float().cpu()                                 # what the kernels read
    res = []
    for dt in (torch.float32, torch.float64):
        oo = copy.deepcopy(o).to(dt).train()
        lg, ft = oo.forward_packed(x_rounded.to(dt), b["speaker_packed"], b["text_length"], None, masks)
        F.cross_entropy(lg, b["label"]).backward()
        res.append(({"logits": lg.detach().numpy(), "features": ft.detach().numpy()}, grads_of(oo)))
    parity_check("bf16_input/config5/outputs", {"logits": logits.detach(), "features": feats.detach()}, res[0][0], res[1][0])
    parity_check("bf16_input/config5/grads", grads_of(m), res[0][1], res[1][1])
    # the mode against the fp32 path on the ORIGINAL (unrounded) features
    m.zero_grad()
    lg32, _ = m.forward_packed(b["x_storage"].cuda()[:, :1443], b["speaker_packed"].cuda(), b["text_length"])
    e = rel_err(logits.detach(), lg32.detach())
    print("bf16-input mode vs fp32 features: logits rel err %.2e" % e)
    assert e < 1e-2
