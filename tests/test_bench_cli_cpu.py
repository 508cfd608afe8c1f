"""bench.py pieces that do not need a GPU: the argument parser (its help text once crashed on a bare '%') and the labels the
per-kernel table and profiles/traffic.json are keyed by."""
import importlib.util
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_help_prints_every_option():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-1000:]
    for opt in ("--gpus", "--steps", "--warmup", "--impl", "--scaling", "--transport", "--bn-sync", "--library-yardstick"):
        assert opt in r.stdout, opt


def test_kernel_labels_match_the_traffic_table():
    b = _bench()
    # (A, lda, B, ldb, bias, C, ldc, M, N, K, ...)
    assert b.kernel_label("ercg_gemm_nn_tc", (0, 1444, 0, 100, 0, 0, 100, 1 << 20, 100, 1443)) == "gemm_nn_tc[K=1443,N=100]"
    # (A, lda, B, ldb, C, ldc, M, K1, N1, ...)
    assert b.kernel_label("ercg_gemm_tn_tc", (0, 1444, 0, 100, 0, 100, 1 << 20, 1443, 100)) == "gemm_tn_tc[K1=1443,N1=100]"
    assert b.kernel_label("ercg_attn_window_bwd_src", ()) == "attn_bwd_src"
    assert b.kernel_label("ercg_gather_window_bwd", ()) == "gather_bwd"
    assert b.kernel_label("ercg_p2p_allreduce", ()) == "p2p_allreduce"
    traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    for key in ("gemm_nn_tc[K=1443,N=100]", "gemm_tn_tc[K1=1443,N1=100]"):
        assert key in traffic, key
