"""tcgen05 3xTF32 GEMM (csrc/gemm_tc.cu) against fp64: must hold the same 1e-5 bound as the exact-fp32 SIMT path."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _tc(fn):
    import erc_b200
    from erc_b200 import ops
    old = ops.GEMM_ENGINE
    ops.GEMM_ENGINE = "tc"
    try:
        return fn(ops)
    finally:
        ops.GEMM_ENGINE = old


@pytest.mark.parametrize("M,K,N", [(256, 32, 16), (300, 100, 100), (1000, 1443, 100), (4096, 1380, 100), (777, 100, 900),
                                   (2050, 900, 100), (513, 100, 400), (640, 400, 100), (33000, 1443, 100), (260, 200, 800),
                                   # >= 148 row tiles: the CTA-pair (cta_group::2) kernels -- long K, small K with one N tile,
                                   # A resident over several N tiles (tail tiles of 16 / 44 / 4 columns), odd tile count
                                   (19000, 100, 400), (20000, 100, 300), (33100, 100, 100), (19071, 128, 900), (25000, 64, 132),
                                   (40000, 400, 100), (19200, 300, 36)])
def test_tc_gemm_matches_fp64(M, K, N):
    g = torch.Generator().manual_seed(M + K + N)
    ld = (K + 3) // 4 * 4
    store = torch.randn(M, ld, generator=g)
    A = store[:, :K]
    B, bias = torch.randn(K, N, generator=g), torch.randn(N, generator=g)
    want = A.double() @ B.double() + bias.double()

    def run(ops):
        from erc_b200 import _lib
        Ad = store.cuda()[:, :K]
        assert _lib.lib().ercg_gemm_nn_tc_supported(Ad.data_ptr(), ld, Ad.data_ptr(), (N + 3) // 4 * 4, M, N, K)
        out = ops.gemm_nn(Ad, B.cuda(), bias.cuda())
        out_relu = ops.gemm_nn(Ad, B.cuda(), bias.cuda(), act=ops.ACT_RELU)
        torch.cuda.synchronize()
        return out, out_relu

    out, out_relu = _tc(run)
    assert rel_err(out, want) < TOL
    assert rel_err(out_relu, want.clamp(min=0)) < TOL


def test_tc_gemm_is_actually_used_and_deterministic():
    import erc_b200
    from erc_b200 import _lib, ops
    g = torch.Generator().manual_seed(1)
    A, B = torch.randn(2048, 128, generator=g).cuda(), torch.randn(128, 100, generator=g).cuda()
    with _lib.KernelTimer() as kt:
        r1 = _tc(lambda o: o.gemm_nn(A, B))
        r2 = _tc(lambda o: o.gemm_nn(A, B))
    assert "ercg_gemm_nn_tc" in kt.summary()
    assert torch.equal(r1, r2)


def test_tc_gemm_adversarial_magnitudes():
    """Large dynamic range: the hi/lo split must keep fp32-grade accuracy, plain TF32 would be ~1e-3."""
    g = torch.Generator().manual_seed(2)
    A = torch.randn(1024, 512, generator=g) * torch.logspace(-3, 3, 512)[None, :]
    B = torch.randn(512, 100, generator=g) / torch.logspace(-3, 3, 512)[:, None]
    want = A.double() @ B.double()
    out = _tc(lambda o: o.gemm_nn(A.cuda(), B.cuda()))
    assert rel_err(out, want) < TOL


@pytest.mark.parametrize("M,K1,N1", [(1024, 100, 100), (5000, 1443, 100), (4100, 100, 900), (2048, 100, 400), (40000, 1380, 100),
                                     (3000, 200, 800), (1500, 300, 100), (1111, 32, 32)])
def test_tc_gemm_tn_matches_fp64(M, K1, N1):
    import erc_b200
    from erc_b200 import _lib
    g = torch.Generator().manual_seed(M + K1 + N1)
    lda = (K1 + 3) // 4 * 4
    store = torch.randn(M, lda, generator=g)
    A, dC = store[:, :K1], torch.randn(M, N1, generator=g)
    want = A.double().t() @ dC.double()

    def run(ops):
        Ad, Bd = store.cuda()[:, :K1], dC.cuda()
        assert _lib.lib().ercg_gemm_tn_tc_supported(Ad.data_ptr(), lda, Bd.data_ptr(), N1, M, K1, N1)
        with _lib.KernelTimer() as kt:
            r1 = ops.gemm_tn(Ad, Bd)
            r2 = ops.gemm_tn(Ad, Bd)
        assert "ercg_gemm_tn_tc" in kt.summary()
        return r1, r2

    r1, r2 = _tc(run)
    assert torch.equal(r1, r2)
    assert rel_err(r1, want) < TOL


@pytest.mark.parametrize("M,K,N", [(1376, 800, 200), (8192, 1443, 100), (4096, 100, 900), (3000, 256, 384),
                                   # CTA-pair kernels (>= 148 row tiles): long K, medium K, A resident over N tiles
                                   (19200, 1443, 100), (19100, 400, 100), (19000, 100, 400)])
def test_tc_gemm_race_stress(M, K, N):
    """The same product 25 times with other work in between: bit-identical and correct every time.  (A shared-memory
    stage that is released before its loads have returned shows up here as a rare, small mismatch.)"""
    g = torch.Generator().manual_seed(M + N)
    A, B = torch.randn(M, K, generator=g).cuda(), torch.randn(K, N, generator=g).cuda()
    want = A.double() @ B.double()

    def run(ops):
        outs = []
        for i in range(25):
            outs.append(ops.gemm_nn(A, B))
            ops.gemm_nn(A[: M // 2], B)                     # perturbs the timing of the next launch
        torch.cuda.synchronize()
        return outs

    outs = _tc(run)
    assert rel_err(outs[0], want) < TOL
    for o in outs[1:]:
        assert torch.equal(o, outs[0])


@pytest.mark.parametrize("M,K,N", [(300, 100, 100), (1000, 300, 100), (33000, 400, 100), (2049, 100, 36), (70000, 900, 128)])
def test_tc_gemm_fused_column_sums(M, K, N):
    """want_colsum: the epilogue reduces the column sums of the product from the staged tiles (bias gradient of the
    upstream layer).  Must equal the fp64 column sums of the fp64 product at the same 1e-5 bound, and ops.colsum must
    pick them up from the tensor instead of launching its own reduction."""
    g = torch.Generator().manual_seed(M * 3 + K + N)
    A, B = torch.randn(M, K, generator=g), torch.randn(K, N, generator=g)
    want = (A.double() @ B.double()).sum(0)

    def run(ops):
        from erc_b200 import _lib
        C = ops.gemm_nn(A.cuda(), B.cuda(), want_colsum=True)
        assert ops._tag_get(C, "_ercg_colsum") is not None
        n0 = _lib.launch_count()
        cs = ops.colsum(C)
        assert _lib.launch_count() == n0            # served from the fused result, no kernel
        plain = ops.gemm_nn(A.cuda(), B.cuda())
        return C, plain, cs, ops.colsum(plain)

    C, plain, cs, cs_plain = _tc(run)
    assert torch.equal(C, plain)                    # the product itself is untouched
    scale = float((A.double().abs() @ B.double().abs()).sum(0).max())
    assert float((cs.cpu().double() - want).abs().max()) < TOL * scale
    assert rel_err(cs, cs_plain) < 1e-6


@pytest.mark.parametrize("M,K1,N1", [(40000, 1443, 100), (40000, 100, 400), (20000, 256, 100), (6000, 1443, 100)])
def test_tc_gemm_tn_race_stress(M, K1, N1):
    """Weight-gradient kernels (single CTA and CTA pairs: even numbers of 128-column tiles of the wide operand) 25 times with
    other work in between: bit-identical and correct every time -- the pair protocol crosses CTAs (remote arrivals on the
    leader's barriers, multicast commits, conversion teams that skip every other chunk)."""
    g = torch.Generator().manual_seed(M + K1 + N1)
    A, B = torch.randn(M, (K1 + 3) // 4 * 4, generator=g).cuda()[:, :K1], torch.randn(M, N1, generator=g).cuda()
    want = A.double().t() @ B.double()

    def run(ops):
        outs = []
        for i in range(25):
            outs.append(ops.gemm_tn(A, B))
            ops.gemm_tn(A[: M // 2], B[: M // 2])           # perturbs the timing of the next launch
        torch.cuda.synchronize()
        return outs

    outs = _tc(run)
    assert rel_err(outs[0], want) < TOL
    for o in outs[1:]:
        assert torch.equal(o, outs[0])
