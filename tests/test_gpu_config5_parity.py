"""The path bench.py times, checked VALUE BY VALUE against the oracle: BASELINE config-5 shape (hidden_all 1443, 6 classes,
one speaker id, dialogue lengths 1+Geom(1/7) <= 40, synth.config5_lengths) through ``COGMENModule.forward_packed`` --
tcgen05 projection GEMM (NN forward, TN weight gradient), relation census (2 of 8 relation ids), window gather /
attention kernels, fused classifier tail -- at N ~ 16 k and ~ 64 k utterances, logits / loss / all 19 live gradients
(cogmen.py:138-160,179-195).  fp64 oracle = the truth for the tolerance bookkeeping (conftest.parity_check).
Dropout: once off (p = 0), once ON with the kernel's own mask handed to the oracle."""
import copy

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import parity_check, grads_of
from oracle import modules as om

pytestmark = pytest.mark.gpu
HIDDEN, C = 1443, 6


def _batch(total, seed):
    import erc_b200  # noqa: F401
    from erc_b200 import synth
    lengths = synth.config5_lengths(total, seed=seed)
    gen = torch.Generator().manual_seed(100 + seed)
    return synth.packed_batch(lengths, HIDDEN, 2, C, gen, one_speaker=True)


def _oracle_run(o, b, dtype, drop_mask, act_masks):
    o = copy.deepcopy(o).to(dtype)
    o.train()
    logits, feats = o.forward_packed(b["x_packed"].to(dtype), b["speaker_packed"], b["text_length"],
                                     None if drop_mask is None else drop_mask.to(dtype), act_masks)
    loss = F.cross_entropy(logits, b["label"])
    loss.backward()
    out = {"logits": logits.detach().numpy(), "features": feats.detach().numpy(), "loss": np.asarray(float(loss.detach()))}
    # branch pinning (see CogmenOracle.forward_packed) may only ever touch elements within rounding distance of the kink
    for name, (count, far) in o.kinks.items():
        assert count <= 64 and far < 1e-4, (name, count, far)
    return out, grads_of(o), dict(o.kinks)


@pytest.mark.parametrize("total,seed", [(1 << 14, 0), (1 << 16, 1)])
@pytest.mark.parametrize("dropout", [False, True])
def test_config5_forward_packed_vs_oracle(total, seed, dropout, monkeypatch):
    from erc_b200 import ops, _lib
    from erc_b200.track_mm import cogmen as ours_cogmen
    from erc_b200.track_mm.cogmen import COGMENModule
    b = _batch(total, seed)
    N = b["x_packed"].size(0)
    torch.manual_seed(seed)
    o = om.CogmenOracle(HIDDEN, n_classes=C, dropout=0.0)
    m = COGMENModule(HIDDEN, 100, 17, 2, C, build_dead_encoder=False).cuda()
    m.load_state_dict(o.state_dict(), strict=False)
    m.train()
    mask = None
    if dropout:
        SEED = 0x5EED + seed
        monkeypatch.setattr(ours_cogmen, "_fresh_seed", lambda: SEED)   # the classifier's dropout seed, pinned for this test
        eye = torch.eye(100, device="cuda")
        # the epilogue's mask is a function of (seed, row, column, width): read it off a product whose pre-activation is 1
        mask = ops.linear(torch.ones(N, 100, device="cuda"), eye, None, act=ops.ACT_RELU_DROPOUT, drop_p=0.5, seed=SEED).cpu()
        kept = float((mask > 0).float().mean())
        assert set(np.unique(mask.numpy()).tolist()) <= {0.0, 2.0} and 0.49 < kept < 0.51, kept
    else:
        m.cls[2].p = 0.0
    before = _lib.launch_count()
    x = b["x_storage"].cuda()[:, :HIDDEN]                             # row pitch 1444 floats: the resident layout of the bench
    seen = {}
    hook = m.gcn.register_forward_hook(lambda mod, inp, out: seen.__setitem__("graph_out", out.detach()))
    logits, feats = m.forward_packed(x, b["speaker_packed"].cuda(), b["text_length"])
    hook.remove()
    # the branches the kernels took at the two piecewise-linear activations: sign of the LeakyReLU output, and the ReLU
    # pattern of the classifier's hidden layer (recomputed with the same kernel and arguments the fused node uses)
    lin0 = m.cls[0]
    hid = ops.linear(seen["graph_out"], lin0.weight.detach(), lin0.bias.detach(), act=ops.ACT_RELU)
    act_masks = ((seen["graph_out"] > 0).cpu(), (hid > 0).cpu())
    loss = ops.cross_entropy(logits, b["label"].cuda())
    loss.backward()
    torch.cuda.synchronize()
    assert _lib.launch_count() - before >= 30
    got = {"logits": logits.detach().cpu().numpy(), "features": feats.detach().cpu().numpy(),
           "loss": np.asarray(float(loss))}
    ggrads = grads_of(m)
    r32, g32, k32 = _oracle_run(o, b, torch.float32, mask, act_masks)
    r64, g64, k64 = _oracle_run(o, b, torch.float64, mask, act_masks)
    case = "config5/N=%d/dropout=%s" % (N, "on" if dropout else "off")
    print(case, "activations on the other side of a kink (count, max |pre-activation|): fp32 oracle", k32, "fp64 oracle", k64)
    parity_check(case + "/outputs", got, r32, r64)
    assert len(g32) == 19 and set(ggrads) == set(g32)
    parity_check(case + "/grads", ggrads, g32, g64)
    # only relation ids 0 / 1 occur: the other six weight slices get an exactly-zero gradient (as in the reference)
    gw = ggrads["gcn.conv1.weight"]
    assert all(np.abs(gw[r]).max() > 0 for r in (0, 1)) and all(np.abs(gw[r]).max() == 0 for r in range(2, 8))
