"""tcgen05 GEMM (csrc/gemm_tcgen05.cu) against torch.matmul in fp32 on the same (bf16-rounded) operands."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref_act(x, act, slope):
    if act == "quickgelu":
        return x * torch.sigmoid(1.702 * x)
    if act == "gelu":
        return torch.nn.functional.gelu(x)
    if act == "leaky_relu":
        return torch.nn.functional.leaky_relu(x, slope)
    if act == "sigmoid":
        return torch.sigmoid(x)
    if act == "relu":
        return torch.relu(x)
    return x


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (128, 128, 128), (300, 768, 768), (197 * 8, 2304, 768),
                                   (20000, 2304, 768), (4097, 768, 3072), (45, 512, 1536), (9, 6144, 512),
                                   (10001, 768, 768), (47280, 3072, 768), (18945, 768, 3072)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_gemm_plain(M, N, K, dtype):
    from clip_spm_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(M * 31 + N * 7 + K)
    a = torch.randn(M, K, generator=g).cuda().to(dtype)
    b = (torch.randn(N, K, generator=g) / K ** 0.5).cuda().to(dtype)
    out = ops.gemm(a, b)
    torch.cuda.synchronize()
    if dtype == torch.float32:
        # tf32 keeps 10 mantissa bits of each operand
        ref = a.double() @ b.double().t()
        tol = 2e-3
    else:
        ref = a.double() @ b.double().t()
        tol = 1e-4  # operands are exactly bf16; only fp32 accumulation order differs
    err = (out.double() - ref).abs().max().item() / ref.abs().max().item()
    assert err < tol, (M, N, K, err)


@pytest.mark.parametrize("act", ["none", "quickgelu", "gelu", "leaky_relu", "sigmoid", "relu"])
@pytest.mark.parametrize("out_dtype", [torch.float32, torch.bfloat16])
def test_gemm_epilogue(act, out_dtype):
    from clip_spm_b200 import ops
    M, N, K = 1000, 768, 512
    g = torch.Generator(device="cpu").manual_seed(5)
    a = torch.randn(M, K, generator=g).cuda().bfloat16()
    b = (torch.randn(N, K, generator=g) / K ** 0.5).cuda().bfloat16()
    bias = torch.randn(N, generator=g).cuda()
    # the residual stream is fp32: a residual is only accepted together with an fp32 output
    res = torch.randn(M, N, generator=g).cuda() if out_dtype == torch.float32 else None
    out = ops.gemm(a, b, bias=bias, act=act, slope=0.0025, residual=res, out_dtype=out_dtype)
    ref = _ref_act(a.float() @ b.float().t() + bias, act, 0.0025) + (res if res is not None else 0.0)
    tol = 2e-2 if out_dtype == torch.bfloat16 else 2e-4
    err = (out.float() - ref).abs().max().item() / ref.abs().max().item()
    assert err < tol, (act, out_dtype, err)


def test_gemm_inplace_residual_and_row_maps():
    from clip_spm_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(9)
    # in-place residual stream update: x += a @ b.T + bias
    M, N, K = 197 * 5, 768, 3072
    a = torch.randn(M, K, generator=g).cuda().bfloat16()
    b = (torch.randn(N, K, generator=g) / K ** 0.5).cuda().bfloat16()
    bias = torch.randn(N, generator=g).cuda()
    x = torch.randn(M, N, generator=g).cuda()
    ref = x + a.float() @ b.float().t() + bias
    ops.gemm(a, b, bias=bias, residual=x, out=x)
    assert (x - ref).abs().max().item() < 2e-4 * ref.abs().max().item()
    # patch-embedding layout: patch row m of frame f goes to token row f*197 + 1 + m%196, plus pos[1 + m%196]
    F_, P, W = 3, 196, 768
    a = torch.randn(F_ * P, W, generator=g).cuda().bfloat16()
    b = (torch.randn(W, W, generator=g) / W ** 0.5).cuda().bfloat16()
    pos = torch.randn(P + 1, W, generator=g).cuda()
    out = torch.zeros(F_ * (P + 1), W, device="cuda")
    ops.gemm(a, b, residual=pos, res_row_mod=P, res_row_off=1, out_row_group=P, out_group_stride=P + 1, out_row_off=1,
             out=out)
    ref = (a.float() @ b.float().t()).view(F_, P, W) + pos[1:]
    got = out.view(F_, P + 1, W)
    assert (got[:, 1:] - ref).abs().max().item() < 2e-4 * ref.abs().max().item()
    assert got[:, 0].abs().max().item() == 0.0


def test_gemm_strided_rows():
    """A operand taken as every 197th row (class-token rows) without a gather."""
    from clip_spm_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(11)
    x = torch.randn(40 * 197, 768, generator=g).cuda().bfloat16()
    w = (torch.randn(512, 768, generator=g) / 768 ** 0.5).cuda().bfloat16()
    a = x.view(40, 197, 768)[:, 0, :]
    out = ops.gemm(a, w)
    ref = a.float() @ w.float().t()
    assert (out - ref).abs().max().item() < 2e-4 * ref.abs().max().item()


def test_gemm_two_cta_path_with_epilogues():
    """shapes large enough for the cta_group::2 kernel (pairs of CTAs, 256x256 tiles), odd M, every epilogue flavour"""
    from clip_spm_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(21)
    M, N, K = 19999, 768, 768
    a = torch.randn(M, K, generator=g).cuda().bfloat16()
    b = (torch.randn(N, K, generator=g) / K ** 0.5).cuda().bfloat16()
    bias = torch.randn(N, generator=g).cuda()
    x = torch.randn(M, N, generator=g).cuda()
    ref = x + a.float() @ b.float().t() + bias
    ops.gemm(a, b, bias=bias, residual=x, out=x)                      # in-place fp32 residual stream
    assert (x - ref).abs().max().item() < 2e-4 * ref.abs().max().item()
    out = ops.gemm(a, b, bias=bias, act="quickgelu", out_dtype=torch.bfloat16)
    r2 = a.float() @ b.float().t() + bias
    r2 = r2 * torch.sigmoid(1.702 * r2)
    assert (out.float() - r2).abs().max().item() < 2e-2 * r2.abs().max().item()


@pytest.mark.parametrize("M,K", [(19999, 768), (25216, 3072), (18944, 64)])
def test_gemm_two_cta_tma_residual(M, K):
    """2-CTA kernel with the TMA-prefetched residual ring: separate (not in-place) residual with a row stride larger
    than N, no bias, ragged and exact M, short and long reductions; each output element must get its own residual."""
    from clip_spm_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(M + K)
    N = 768
    a = torch.randn(M, K, generator=g).cuda().bfloat16()
    b = (torch.randn(N, K, generator=g) / K ** 0.5).cuda().bfloat16()
    res_full = torch.randn(M, 1024, generator=g).cuda()
    res = res_full[:, 128:128 + N]                       # 16-byte aligned view, row stride 1024
    out = torch.full((M, N), float("nan"), device="cuda")
    ops.gemm(a, b, residual=res, out=out)
    ref = res + a.float() @ b.float().t()
    assert torch.isfinite(out).all()
    assert (out - ref).abs().max().item() < 2e-4 * ref.abs().max().item()
    # a residual that is an exact function of (row, col) makes any misplaced box visible even where the GEMM term is tiny
    rows = torch.arange(M, device="cuda", dtype=torch.float32)[:, None]
    cols = torch.arange(N, device="cuda", dtype=torch.float32)[None, :]
    res2 = (rows * 1024 + cols).contiguous()
    z = torch.zeros(M, K, device="cuda", dtype=torch.bfloat16)
    out2 = ops.gemm(z, b, residual=res2, out=torch.empty(M, N, device="cuda"))
    assert torch.equal(out2, res2)
