"""oracle/ref_port.py (the timed CPU baseline of bench.py's reference arm) pinned against the REAL reference module, and
the cogmen.py:114 quirk (GNN always built with n_speakers = 2) pinned for a 9-speaker (MELD-shaped) batch.
CPU only; needs /root/reference (present in the build container, absent on the GPU box -> skipped there)."""
import numpy as np
import pytest
import torch

from oracle import ref_loader, ref_port, modules as om

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")


def _batch(B, D, n_speakers, n_classes, seed, lo=3, hi=20):
    gen = torch.Generator().manual_seed(seed)
    lens = torch.randint(lo, hi, (B,), generator=gen)
    Lmax = int(lens.max())
    x = torch.randn(B, Lmax, D, generator=gen)
    spk = torch.randint(0, n_speakers, (B, Lmax), generator=gen)
    mask = torch.arange(Lmax)[None, :] < lens[:, None]
    return x * mask[..., None], spk * mask, lens, torch.randint(0, n_classes, (int(lens.sum()),), generator=gen)


def test_ref_port_is_logit_identical_to_the_reference_module():
    ref = ref_loader.load()
    D, C = 36, 4                                              # 36 % 6 == 0 -> 6 heads
    torch.manual_seed(0)
    real = ref.cogmen.COGMENModule(D, 100, 17, 2, C)
    port = ref_port.CogmenRefPort(D, n_classes=C)
    port.load_state_dict(real.state_dict(), strict=True)      # same parameter names and shapes, dead encoder included
    x, spk, lens, y = _batch(5, D, 2, C, seed=1)
    # graph construction: same edges in the same (CPython set) order, same relation ids
    f0, ei0, et0, el0 = ref.cogmen_utils.batch_graphify(x, lens, spk, 5, 5, real.edge_type_to_idx)
    f1, ei1, et1, el1 = ref_port.graphify_loop(x, lens, spk, 5, 5, port.rel_ids)
    assert torch.equal(ei0, ei1) and torch.equal(et0, et1) and torch.equal(el0, el1) and torch.equal(f0, f1)
    real.eval()
    port.eval()
    with torch.no_grad():
        a, fa = real(x, spk, lens)
        b, fb = port(x, spk, lens)
    assert float((a - b).abs().max()) == 0.0 and float((fa - fb).abs().max()) == 0.0
    # one training step each from the same RNG state: same loss, same updated parameters
    outs = []
    for m in (real, port):
        m.train()
        torch.manual_seed(7)
        opt = torch.optim.Adam(m.parameters(), lr=1e-4, weight_decay=1e-8)
        loss = ref_port.train_step(m, opt, dict(input_tensor=x, speaker_tensor=spk, text_length=lens, label=y))
        outs.append((float(loss), {k: v.clone() for k, v in m.state_dict().items()}))
    assert outs[0][0] == outs[1][0]
    for k in outs[0][1]:
        assert torch.equal(outs[0][1][k], outs[1][1][k]), k


def test_reference_builds_gnn_with_default_speaker_count_and_oracle_follows():
    """cogmen.py:114: ``GNN(hidden, hidden, hidden)`` -- 8 relations for every data set.  With 9 speakers (MELD) relation ids
    run to 161; ids >= 8 are never selected by PyG's per-relation loop, i.e. those edges carry no message."""
    ref = ref_loader.load()
    D, C = 36, 7
    torch.manual_seed(3)
    real = ref.cogmen.COGMENModule(D, 100, 17, 9, C)
    assert tuple(real.gcn.conv1.weight.shape) == (8, 100, 100) and len(real.edge_type_to_idx) == 162
    o = om.CogmenOracle(D, n_speakers=9, n_classes=C, dropout=0.0)
    assert om.load_live(o, real.state_dict()) and tuple(o.gcn.conv1.weight.shape) == (8, 100, 100)
    x, spk, lens, y = _batch(6, D, 9, C, seed=4)
    real.eval()
    o.eval()
    with torch.no_grad():
        a, _ = real(x, spk, lens)
        b, _ = o(x, spk, lens)
    assert float((a - b).abs().max()) <= 1e-5 * float(a.abs().max())


def test_oracle_gat_step_equals_the_reference_class():
    """dagerc_oracle._Gat (the checker of the stand-alone CUDA GAT step) against the real GAT_dialoggcn_v1
    (dagerc_models.py:312-365): same state_dict, same outputs."""
    from oracle import dagerc_oracle
    ref = ref_loader.load()
    torch.manual_seed(11)
    real = ref.dagerc_models.GAT_dialoggcn_v1(48)
    o = dagerc_oracle._Gat(48)
    o.load_state_dict(real.state_dict(), strict=True)
    gen = torch.Generator().manual_seed(5)
    Q, K = torch.randn(4, 48, generator=gen), torch.randn(4, 9, 48, generator=gen)
    adj = (torch.rand(4, 9, generator=gen) < 0.5).float()
    sm = (torch.rand(4, 9, generator=gen) < 0.5).long()
    w0, s0 = real(Q, K, K, adj, sm)
    w1, s1 = o(Q, K, K, adj, sm)
    assert torch.equal(w0, w1) and torch.equal(s0, s1)
