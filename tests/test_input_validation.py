"""Bad inputs fail loudly: the reference raises IndexError / KeyError in its python loops
(cogmen_utils.py:131-137) and F.cross_entropy device-asserts on labels outside [0, C); the kernels must not read out of
bounds or silently mis-type edges.  Host-resident inputs are validated on the host (CPU tests, no GPU needed);
device-resident ones by K1 itself (flags in rel_info[513]) -- GPU tests."""
import pytest
import torch


def test_host_validation_raises_before_any_device_work():
    import erc_b200  # noqa: F401
    from erc_b200.graph import build_graph
    spk = torch.zeros(3, 7, dtype=torch.int64)
    with pytest.raises(ValueError, match="padded speaker width"):
        build_graph(torch.tensor([3, 9, 2]), spk, 5, 5, 2)
    with pytest.raises(ValueError, match="negative"):
        build_graph(torch.tensor([3, -1, 2]), spk, 5, 5, 2)
    spk[1, 2] = 2
    with pytest.raises(ValueError, match="speaker ids"):
        build_graph(torch.tensor([3, 4, 2]), spk, 5, 5, 2)


@pytest.mark.gpu
def test_kernel_flags_device_resident_bad_inputs():
    import erc_b200  # noqa: F401
    from erc_b200.graph import build_graph
    lens = torch.tensor([3, 9, 2]).cuda()                       # device lengths: only the kernel can see them
    spk = torch.zeros(3, 7, dtype=torch.int64).cuda()
    g = build_graph(lens, spk, 5, 5, 2)
    with pytest.raises(ValueError, match="padded speaker width"):
        g.check_inputs()
    assert int(g.pad_row.max()) < 3 * 7 and int(g.pad_row.min()) >= 0          # clamped: usable as row indices
    spk2 = torch.zeros(3, 9, dtype=torch.int64).cuda()
    spk2[1, 4] = 5
    g = build_graph(lens, spk2, 5, 5, 2)
    with pytest.raises(ValueError, match="speaker id"):
        g.relation_slots()
    assert int(g.etype.max()) < 8
    g = build_graph(lens, torch.zeros(3, 9, dtype=torch.int64).cuda(), 5, 5, 2)
    g.check_inputs()                                             # clean input: no error
    assert g.relation_slots()[0] == [0, 1]


@pytest.mark.gpu
def test_cross_entropy_poisons_the_loss_on_out_of_range_labels():
    import erc_b200  # noqa: F401
    from erc_b200 import ops
    logits = torch.randn(50, 6).cuda()
    y = torch.randint(0, 6, (50,)).cuda()
    assert torch.isfinite(ops.cross_entropy(logits, y))
    y2 = y.clone()
    y2[7] = -100                                                 # ignore_index: weight 0, like F.cross_entropy
    want = torch.nn.functional.cross_entropy(logits, y2)
    assert abs(float(ops.cross_entropy(logits, y2)) - float(want)) < 1e-5
    y2[9] = 6
    assert torch.isnan(ops.cross_entropy(logits, y2))
