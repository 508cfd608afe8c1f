"""INTEGRATION.md Level 1 (import swaps applied to the real reference files) -- see tests/level1_swap_check.py."""
import json
import os
import subprocess
import sys

import pytest

from oracle import ref_loader

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")
def test_level1_import_swaps_construct_and_load_reference_state_dicts():
    p = subprocess.run([sys.executable, os.path.join(HERE, "level1_swap_check.py")], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-3000:]
    out = json.loads(p.stdout.strip().splitlines()[-1])
    assert out["cogmen_keys"] > 40 and out["dgcn_keys"] > 20
    import torch
    if not torch.cuda.is_available():
        assert "cpu_forward_refused" in out, out           # no CPU fallback: a CPU tensor must raise, not compute
