"""track2d_gemm_tf32x3 (tcgen05 3xTF32) through the C ABI against float64 and against torch's fp32 linear autograd.
Tolerance: |err| <= 2e-6 * (|A| |B|)[m, n] -- fp32 accuracy (torch's own fp32 GEMM lands at 2e-7..4e-7 on the same inputs);
plain TF32 would be ~5e-4."""
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [
    # name, M, N, K, a_mn_major, b_mn_major, bias, relu
    ("one_tile", 128, 128, 32, 0, 0, False, False),
    ("fc_forward", 512, 256, 512, 0, 0, True, True),
    ("ragged_forward", 300, 132, 100, 0, 0, True, False),
    ("single_env_forward", 1, 256, 1024, 0, 0, True, True),
    ("dgrad", 256, 512, 256, 0, 1, False, False),
    ("dgrad_ragged", 260, 132, 36, 0, 1, False, False),
    ("wgrad_splitk", 256, 512, 1024, 1, 1, False, False),
    ("wgrad_ragged_splitk", 256, 1024, 4100, 1, 1, False, False),
    ("wgrad_tiny_batch", 512, 128, 3, 1, 1, False, False),
    ("a_mn_only", 256, 128, 256, 1, 0, False, False),
    ("many_tiles_persistent", 128 * 150, 256, 64, 0, 0, True, False),
]


@pytest.mark.parametrize("name,M,N,K,a_mn,b_mn,bias,relu", SHAPES, ids=[s[0] for s in SHAPES])
def test_gemm_matches_float64(name, M, N, K, a_mn, b_mn, bias, relu):
    from active_tracking_rl_b200 import gemm as G
    dev = "cuda:0"
    g = torch.Generator(device=dev).manual_seed(7)
    A = torch.randn((K, M) if a_mn else (M, K), generator=g, device=dev)
    B = torch.randn((K, N) if b_mn else (N, K), generator=g, device=dev)
    bv = torch.randn(N, generator=g, device=dev) if bias else None
    Am, Bm = (A.t() if a_mn else A), (B.t() if b_mn else B)
    ref = Am.double() @ Bm.double().t()
    if bias:
        ref = ref + bv.double()
    if relu:
        ref = ref.clamp_min(0)
    out = G.gemm(A, a_mn, A.stride(0), B, b_mn, B.stride(0), M, N, K, bias=bv, relu=relu)
    scale = Am.double().abs() @ Bm.double().abs().t() + (bv.double().abs() if bias else 0)
    assert ((out.double() - ref).abs() <= 2e-6 * scale + 1e-30).all()
    # bit-reproducible (fixed split-K summation order)
    out2 = G.gemm(A, a_mn, A.stride(0), B, b_mn, B.stride(0), M, N, K, bias=bv, relu=relu)
    assert torch.equal(out, out2)


def test_linear_autograd_matches_torch():
    from active_tracking_rl_b200 import gemm as G
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = "cuda:0"
    g = torch.Generator(device=dev).manual_seed(3)
    x = torch.randn(640, 2, 128, generator=g, device=dev)[:, 1]  # strided rows, like hx[:, 1]
    x.requires_grad_(True)
    w = torch.randn(512, 128, generator=g, device=dev, requires_grad=True)
    b = torch.randn(512, generator=g, device=dev, requires_grad=True)
    gy = torch.randn(640, 512, generator=g, device=dev)
    y = G.linear(x, w, b, relu=True)
    gx, gw, gb = torch.autograd.grad(y, (x, w, b), gy)
    y_ref = torch.relu(torch.nn.functional.linear(x.double(), w.double(), b.double()))
    rx, rw, rb = torch.autograd.grad(y_ref, (x, w, b), gy.double())
    for got, ref in ((y, y_ref), (gx, rx), (gw, rw), (gb, rb)):
        assert torch.allclose(got.double(), ref.double(), rtol=1e-5, atol=1e-4 * float(ref.abs().max()))


def test_rejects_unaligned():
    from active_tracking_rl_b200 import _lib, gemm as G
    a = torch.randn(8, 6, device="cuda:0")  # K = 6 is not a multiple of 4
    b = torch.randn(8, 6, device="cuda:0")
    with pytest.raises(_lib.Track2DError):
        G.gemm(a, 0, 6, b, 0, 6, 8, 8, 6)


def test_lstm_pointwise_matches_torch_lstmcell():
    """csrc/track2d_lstm.cu against nn.LSTMCell's own fused pointwise op (forward values, all five gradients), with the strided
    cx / dhy / dcy views the batched A3C_Dueling produces (hx, cx are slices of (E, 2, 128) tensors)."""
    from active_tracking_rl_b200 import gemm as G
    dev = "cuda:0"
    g = torch.Generator(device=dev).manual_seed(11)
    E, H = 1000, 128
    ig = torch.randn(E, 4 * H, generator=g, device=dev, requires_grad=True)
    hg = torch.randn(E, 4 * H, generator=g, device=dev, requires_grad=True)
    cx2 = torch.randn(E, 2, H, generator=g, device=dev, requires_grad=True)
    b_ih = torch.randn(4 * H, generator=g, device=dev, requires_grad=True)
    b_hh = torch.randn(4 * H, generator=g, device=dev, requires_grad=True)
    gh = torch.randn(E, 2, H, generator=g, device=dev)
    gc = torch.randn(E, 2, H, generator=g, device=dev)
    cx = cx2[:, 1]
    assert G.lstm_pointwise_supported(ig, cx)
    hy, cy = G.lstm_pointwise(ig, hg, cx, b_ih, b_hh)
    hy_r, cy_r, _ = torch.ops.aten._thnn_fused_lstm_cell(ig, hg, cx, b_ih, b_hh)
    assert torch.allclose(hy, hy_r, rtol=1e-5, atol=1e-6) and torch.allclose(cy, cy_r, rtol=1e-5, atol=1e-6)
    ins = (ig, hg, cx2, b_ih, b_hh)
    got = torch.autograd.grad((hy, cy), ins, (gh[:, 0], gc[:, 1]))
    ref = torch.autograd.grad((hy_r, cy_r), ins, (gh[:, 0], gc[:, 1]))
    for a, b in zip(got, ref):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-5 * float(b.abs().max()) + 1e-6)
    # only hy used downstream (dcy = None inside the Function), and bit-reproducible bias gradients
    got2 = torch.autograd.grad(G.lstm_pointwise(ig, hg, cx, b_ih, b_hh)[0], ins, gh[:, 0])
    ref2 = torch.autograd.grad(torch.ops.aten._thnn_fused_lstm_cell(ig, hg, cx, b_ih, b_hh)[0], ins, gh[:, 0])
    for a, b in zip(got2, ref2):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-5 * float(b.abs().max()) + 1e-6)
    got3 = torch.autograd.grad(G.lstm_pointwise(ig, hg, cx, b_ih, b_hh)[0], ins, gh[:, 0])
    assert torch.equal(got2[3], got3[3])


@pytest.mark.parametrize("M,N", [(65536, 256), (1000, 8), (7, 512), (4096, 1024), (300, 12)])
def test_colsum_matches_torch(M, N):
    from active_tracking_rl_b200 import gemm as G
    g = torch.Generator(device="cuda:0").manual_seed(5)
    x = torch.randn(M, N, generator=g, device="cuda:0")
    out = G.colsum(x)  # N = 12 is not a power of two: torch's own sum
    ref = x.double().sum(0)
    assert torch.allclose(out.double(), ref, rtol=1e-5, atol=1e-4 * float(M) ** 0.5)
    assert torch.equal(out, G.colsum(x))
