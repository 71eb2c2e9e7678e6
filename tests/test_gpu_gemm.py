"""GPU: the two GEMM engines (SIMT fp32, tcgen05) against an fp64 evaluation of the same operands."""
import pytest
import torch

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _mk(shape, seed, dev="cuda"):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randn(*shape, generator=g).to(dev)


def _ref(A, B, alpha=1.0):
    return alpha * (A.double() @ B.double())


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (300, 180, 200), (1, 7, 5), (257, 33, 1030)])
@pytest.mark.parametrize("ta,tb", [(False, False), (True, False), (False, True), (True, True)])
def test_simt_layouts(M, N, K, ta, tb):
    import sml_b200
    A = _mk((K, M), 1).t() if ta else _mk((M, K), 1)
    B = _mk((N, K), 2).t() if tb else _mk((K, N), 2)
    C = sml_b200.ops.gemm(A, B)
    assert rel_err(C, _ref(A, B)) < 2e-6


def test_simt_batch_epilogue():
    import sml_b200
    A = _mk((2, 3, 70, 40), 3); B = _mk((3, 40, 70), 4); C0 = _mk((2, 3, 70, 70), 5)
    bias = _mk((70,), 6)
    C = C0.clone()
    sml_b200.ops.gemm(A, B, out=C, alpha=-0.5, beta=2.0, diag=7.0, bias=bias, relu=True)
    ref = -0.5 * (A.double() @ B.double()) + 2.0 * C0.double() + 7.0 * torch.eye(70, device="cuda", dtype=torch.float64) + bias.double()
    assert rel_err(C, ref.clamp_min(0)) < 2e-6


def _bf(t):
    return t.to(torch.bfloat16)


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (128, 64, 128), (300, 180, 200), (256, 64, 1000), (130, 250, 72), (64, 40, 24)])
@pytest.mark.parametrize("ta,tb", [(False, True), (False, False), (True, False), (True, True)])
def test_tc_layouts(M, N, K, ta, tb):
    """ta: A stored M-major ([K,M] row-major); tb (here): B stored K-major ([N,K] row-major)."""
    import sml_b200
    S = sml_b200.ops.Split
    # dims padded to multiples of 8 elements for the TMA stride rule
    A = _bf(_mk((K, M + (-M) % 8), 1))[:, :M].t() if ta else _bf(_mk((M, K + (-K) % 8), 1))[:, :K]
    B = _bf(_mk((N, K + (-K) % 8), 2))[:, :K].t() if tb else _bf(_mk((K, N + (-N) % 8), 2))[:, :N]
    C = sml_b200.ops.gemm(S(A), S(B))
    torch.cuda.synchronize()
    assert sml_b200._lib.lib.sml_tc_timeout_flag() == 0
    assert rel_err(C, _ref(A, B)) < 1e-5, (M, N, K, ta, tb)


def test_tc_batch_views_like_nystrom():
    import sml_b200
    S = sml_b200.ops.Split
    b, h, n, d, m = 2, 4, 200, 64, 96
    qkv = _bf(_mk((b, n, 3, h, d), 7))
    kl = _bf(_mk((b, h, m, d), 8))
    q = qkv[:, :, 0].permute(0, 2, 1, 3)
    v = qkv[:, :, 2].permute(0, 2, 1, 3)
    s1 = sml_b200.ops.gemm(S(q), S(kl.transpose(-1, -2)))                       # K-major x K-major, strided batch
    assert rel_err(s1, _ref(q, kl.transpose(-1, -2))) < 1e-5
    a3 = _bf(torch.softmax(_mk((b, h, m, n), 9), -1))
    a3v = sml_b200.ops.gemm(S(a3), S(v))                                         # K-major x N-major (strided rows)
    assert rel_err(a3v, _ref(a3, v)) < 1e-5
    out = torch.zeros(b, n, h * d, device="cuda")
    a1 = _bf(_mk((b, h, n, m), 10)); wm = _bf(_mk((b, h, m, d), 11))
    sml_b200.ops.gemm(S(a1), S(wm), out=out.view(b, n, h, d).permute(0, 2, 1, 3))  # strided output
    assert rel_err(out.view(b, n, h, d).permute(0, 2, 1, 3), _ref(a1, wm)) < 1e-5
    dwm = sml_b200.ops.gemm(S(a1.transpose(-1, -2)), S(v))                       # M-major x N-major, K = n
    assert rel_err(dwm, _ref(a1.transpose(-1, -2), v)) < 1e-5
    wq = _bf(_mk((96, 256), 12))                                                 # broadcast weight over batch
    x = _bf(_mk((b, 150, 256), 13))
    y = sml_b200.ops.gemm(S(x), S(wq.t()))
    assert rel_err(y, _ref(x, wq.t())) < 1e-5


def test_tc_split_is_fp32_grade():
    import sml_b200
    A = _mk((304, 520), 14); B = _mk((520, 264), 15)
    a, b = sml_b200.ops.to_bf16(A, split=True), sml_b200.ops.to_bf16(B, split=True)
    C = sml_b200.ops.gemm(a, b)
    assert rel_err(C, _ref(A, B)) < 2e-5
    C1 = sml_b200.ops.gemm(sml_b200.ops.Split(a.hi), sml_b200.ops.Split(b.hi))
    assert rel_err(C1, _ref(A, B)) < 2e-2          # single pass: bf16-grade only
    assert rel_err(C1, _ref(A, B)) > 1e-4          # (so the split really ran above)
    Bt = B.t().contiguous()                        # K-major B, M-major A with split
    bt = sml_b200.ops.to_bf16(Bt, split=True)
    At = sml_b200.ops.to_bf16(A.t().contiguous(), split=True)
    C2 = sml_b200.ops.gemm(At.transpose(0, 1), bt.transpose(0, 1))
    assert rel_err(C2, _ref(A, B)) < 2e-5


def test_tc_epilogues():
    import sml_b200
    S = sml_b200.ops.Split
    A = _bf(_mk((2, 200, 64), 16)); B = _bf(_mk((2, 64, 256), 17))
    C0 = _mk((2, 200, 256), 18); bias = _mk((256,), 19)
    C = C0.clone()
    hi = torch.empty(2, 200, 256, device="cuda", dtype=torch.bfloat16); lo = torch.empty_like(hi)
    sml_b200.ops.gemm(S(A), S(B), out=C, alpha=-0.5, beta=2.0, diag=7.0, bias=bias, relu=True, out_bf16=hi, out_bf16_lo=lo)
    ref = (-0.5 * (A.double() @ B.double()) + 2.0 * C0.double() + 7.0 * torch.eye(200, 256, device="cuda", dtype=torch.float64)
           + bias.double()).clamp_min(0)
    assert rel_err(C, ref) < 1e-5
    assert rel_err(hi.float() + lo.float(), ref) < 2e-5
    P = sml_b200.ops.gemm(S(A), S(B), alpha=0.3, softmax=True)
    assert rel_err(P, torch.softmax(0.3 * (A.double() @ B.double()), -1)) < 1e-5
    P2 = sml_b200.ops.gemm(S(A), S(B[:, :, :200]), softmax=True)                 # N tail inside the softmax tile
    assert rel_err(P2, torch.softmax(A.double() @ B[:, :, :200].double(), -1)) < 1e-5
    torch.cuda.synchronize()
    assert sml_b200._lib.lib.sml_tc_timeout_flag() == 0
