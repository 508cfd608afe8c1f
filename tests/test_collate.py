"""the collate.  CPU: oracle/collate_np.py == the REAL ERCCollate (track_mm/mmbase.py:344-455, cut out of the
reference file) for every modality / layout combination.  GPU: erc_b200.collate.DeviceCollate (packed upload, padded tensors
built by kernels on the device) == the oracle, key by key, bit for bit; and a COGMEN forward fed by it equals the padded one."""
import types

import numpy as np
import pytest
import torch

from oracle import collate_np, ref_loader

CASES = [("atv", True, False, 2), ("tav", False, True, 2), ("av", True, True, 9), ("t", False, False, 2), ("vt", True, False, 3)]


def _same(a, b, k):
    if a is None or b is None:
        assert a is None and b is None, k
    elif isinstance(a, list):
        assert a == b, k
    else:
        a = a.cpu() if hasattr(a, "cpu") else a
        assert a.dtype == b.dtype and a.shape == b.shape and torch.equal(a, b), (k, a.dtype, b.dtype, a.shape, b.shape)


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")
@pytest.mark.parametrize("modality,batch_first,onehot,n_spk", CASES)
def test_oracle_collate_equals_the_reference_class(modality, batch_first, onehot, n_spk):
    ERCCollate = ref_loader.load_collate()
    params = types.SimpleNamespace(batch_first=batch_first, speaker_onehot=onehot, n_classes=6, n_speakers=n_spk, modality=modality)
    samples = collate_np.synthetic_samples([5, 1, 17, 9], (12, 7, 5), n_spk, 6, seed=3, with_sentence=True)
    want = ERCCollate(params)(samples)
    got = collate_np.collate(samples, modality, batch_first, onehot, n_spk)
    assert set(got) == set(want)
    for k in want:
        _same(got[k], want[k], k)


@pytest.mark.gpu
@pytest.mark.parametrize("modality,batch_first,onehot,n_spk", CASES)
def test_device_collate_equals_oracle(modality, batch_first, onehot, n_spk):
    import erc_b200  # noqa: F401
    from erc_b200.collate import DeviceCollate
    samples = collate_np.synthetic_samples([5, 1, 17, 9, 33], (12, 7, 5), n_spk, 6, seed=4, with_sentence=True)
    want = collate_np.collate(samples, modality, batch_first, onehot, n_spk)
    got = DeviceCollate(modality, batch_first, onehot, n_spk, device="cuda")(samples)
    assert set(want) <= set(got) and set(got) - set(want) == {"x_packed", "speaker_packed"}
    for k in want:
        _same(got[k], want[k], k)
    mask = want["attention_mask"].bool()
    it = want["input_tensor"] if batch_first else want["input_tensor"].transpose(0, 1)
    assert torch.equal(got["x_packed"].cpu(), it[mask])


@pytest.mark.gpu
def test_cogmen_fed_by_device_collate_takes_the_packed_path_and_matches_padded():
    import erc_b200  # noqa: F401
    from erc_b200 import _lib
    from erc_b200.collate import DeviceCollate
    from erc_b200.track_mm.cogmen import COGMENModule
    rng = np.random.default_rng(0)
    lengths = [int(v) for v in rng.integers(8, 60, size=24)]
    samples = collate_np.synthetic_samples(lengths, (768, 100, 512), 2, 4, seed=1)
    batch = DeviceCollate("atv", device="cuda")(samples)
    torch.manual_seed(0)
    m = COGMENModule(1380, 100, 17, 2, 4, build_dead_encoder=False).cuda().eval()
    with torch.no_grad():
        a, _ = m(**batch.packed_kwargs())                       # what a trainer does with the batch: model(**batch)
        assert "input_tensor" not in batch._d                   # the padded tensors were never built
        b, _ = m(input_tensor=batch["input_tensor"], speaker_tensor=batch["speaker_tensor"], text_length=batch["text_length"])
        c, _ = m(**batch)                                       # all keys present: still the packed path
    assert float((a - b).abs().max()) <= 1e-5 * float(b.abs().max()) and torch.equal(a, c)
