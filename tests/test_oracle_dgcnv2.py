"""oracle/dgcnv2_oracle.py against the fixture produced by the REAL dgcnv2.DGCNModule (oracle/make_golden.py dgcnv2)."""
import torch
import torch.nn.functional as F

from conftest import rel_err
from oracle import dgcnv2_oracle, seeded
from oracle.make_golden import DGCNV2_SEED
from test_oracle_mmgcn import check_against_fixture


def test_dgcnv2_oracle_matches_reference_fixture(golden):
    fx = golden("dgcnv2_small")
    D, C = fx["input_tensor"].shape[-1], fx["logits"].shape[1]
    o = dgcnv2_oracle.Dgcnv2Oracle(D, n_classes=C, dropout=0.0)
    seeded.fill_by_name(o, DGCNV2_SEED)
    o.train()
    b = {k: torch.from_numpy(fx[k]) for k in ("input_tensor", "speaker_tensor", "attention_mask", "text_length", "label")}
    logits, feats = o(**{k: v for k, v in b.items() if k != "label"})
    loss = F.cross_entropy(logits, b["label"], weight=torch.from_numpy(fx["class_weights"]))
    loss.backward()
    assert rel_err(logits.detach(), fx["logits"]) < 1e-5 and rel_err(feats.detach(), fx["features"]) < 1e-5
    assert abs(float(loss.detach()) - float(fx["loss"])) < 1e-5 * float(fx["loss"])
    live = set(str(k) for k in fx["live"])
    grads = {k: p.grad.numpy() for k, p in o.named_parameters() if p.grad is not None}
    assert set(grads) == live, sorted(set(grads) ^ live)       # the same parameters are live (dead: att_model.matchatt / att ...)
    check_against_fixture(fx, grads, 5e-5)
