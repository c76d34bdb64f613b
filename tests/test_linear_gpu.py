"""GPU: csb200_linear_fwd (tcgen05 Linear with bias / GELU epilogues) — the K = C GEMMs of the CSWinBlock
(qkv C:357-358, proj C:366, Mlp.fc1 + act C:188-196) — against a plain PyTorch fp32 reference of the same op
on the same bf16-rounded operands, through the C ABI.

Tolerance: the kernel accumulates in fp32 and rounds ONCE to bf16, so y must be within one bf16 rounding
(2^-8 relative to the row's scale) of the fp32 result; GELU is evaluated on the rounded pre-activation, so
against GELU(bf16(h_ref)) the same bound holds except where h_ref sits on a rounding boundary (<= 2 ulp)."""
import pytest
import torch

from conftest import rel_err
import cswin_simam_unet_b200 as pkg
from cswin_simam_unet_b200 import capi, functional as csbF

pytestmark = pytest.mark.gpu

# (M, K, N): config-3 widths of the four K = C GEMM families + ragged M, several n-tiles, one partial tile
SHAPES = [(4096, 64, 192), (4096, 64, 64), (4096, 64, 256), (2048, 128, 384), (2048, 128, 512), (1024, 256, 768),
          (1024, 256, 1024), (1000, 64, 256), (77, 128, 128), (300 * 128 + 5, 64, 256), (19000, 256, 256)]


def _operands(M, K, N, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn((M, K), generator=g).to(torch.bfloat16).cuda()
    w = (torch.randn((N, K), generator=g) / K ** 0.5).to(torch.bfloat16).cuda()
    b = (torch.randn((N,), generator=g) * 0.5).cuda()
    return x, w, b


@pytest.mark.parametrize("shape", SHAPES)
def test_bias_epilogue_matches_fp32_reference(shape):
    M, K, N = shape
    x, w, b = _operands(M, K, N, M + N)
    assert capi.lib().csb200_linear_supported(M, N, K, capi.BF16)
    y, _ = csbF._tc_linear(x, w, b, capi.EPI_BIAS)
    ref = x.float() @ w.float().t() + b
    assert rel_err(y.float(), ref) < 2 ** -8
    # and no worse than cuBLAS on the same operands
    cublas = torch.nn.functional.linear(x, w, b.to(torch.bfloat16))
    assert rel_err(y.float(), ref) <= rel_err(cublas.float(), ref) + 2 ** -9


@pytest.mark.parametrize("shape", SHAPES)
def test_gelu_epilogue_and_saved_preactivation(shape):
    M, K, N = shape
    x, w, b = _operands(M, K, N, M + 2 * N)
    a, h = csbF._tc_linear(x, w, b, capi.EPI_GELU_SAVE)
    ref_h = x.float() @ w.float().t() + b
    assert rel_err(h.float(), ref_h) < 2 ** -8
    # GELU acts on the value that was stored (bit-exact contract with the flat csb200_gelu_fwd pass)
    flat = torch.empty_like(h)
    capi.check(capi.lib().csb200_gelu_fwd(csbF._ptr(h), csbF._ptr(flat), M, N, capi.BF16,
                                          csbF._vp(capi.stream_of(h))), "csb200_gelu_fwd")
    assert torch.equal(a, flat)
    ref_a = torch.nn.functional.gelu(h.float())  # exact erf form, nn.GELU() of C:190
    assert rel_err(a.float(), ref_a) < 2 ** -8
    a_only, none = csbF._tc_linear(x, w, b, capi.EPI_GELU)
    assert none is None and torch.equal(a_only, a)


def test_strided_rows_no_bias_and_rejections():
    M, K, N = 640, 64, 128
    x, w, b = _operands(M, 2 * K, N, 3)
    view = x[:, :K]  # row stride 2K: a channel slice of a wider token matrix
    y, _ = csbF._tc_linear(view, w[:, :K].contiguous(), None, capi.EPI_BIAS)
    assert rel_err(y.float(), view.float() @ w[:, :K].float().t()) < 2 ** -8
    lib = capi.lib()
    assert not lib.csb200_linear_supported(128, 128, 96, capi.BF16)   # K not tiled
    assert not lib.csb200_linear_supported(128, 100, 64, capi.BF16)   # N not a multiple of 32
    assert not lib.csb200_linear_supported(128, 128, 64, capi.F32)
    with pytest.raises(RuntimeError, match="K in"):
        csbF._tc_linear(torch.zeros(128, 96, dtype=torch.bfloat16, device="cuda"),
                        torch.zeros(128, 96, dtype=torch.bfloat16, device="cuda"), None, capi.EPI_BIAS)


def test_mlp_through_the_fused_path_matches_the_two_pass_path():
    """Mlp.forward (C:188-196) under bf16 autocast: tcgen05 fc1 + GELU (saving GELU'(h)) vs cuBLAS fc1 + flat GELU
    pass (saving h): outputs and gradients agree to bf16 rounding."""
    torch.manual_seed(0)
    mlp = pkg.modules.Mlp(128, 512).cuda()
    x = torch.randn(4, 1024, 128, device="cuda").to(torch.bfloat16)
    outs = {}
    for fused in (True, False):
        csbF.set_tc_linear(gelu=fused)
        xi = x.clone().requires_grad_(True)
        mlp.zero_grad()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = mlp(xi)
        y.float().square().mean().backward()
        outs[fused] = (y.detach().float(), xi.grad.float(), mlp.fc1.weight.grad.clone(), mlp.fc1.bias.grad.clone())
    csbF.set_tc_linear(gelu=True)
    for a, b in zip(outs[True], outs[False]):
        assert rel_err(a, b) < 2 ** -6


@pytest.mark.parametrize("shape", [(4096, 64, 256), (2048, 128, 512), (1024, 256, 1024), (1000, 64, 128), (77, 128, 256),
                                   (300 * 128 + 5, 64, 256), (19000, 256, 512)])
def test_dgelu_backward_gemm_matches_fp32_reference(shape):
    """csb200_linear_dgelu_bwd: grad_h = (grad_y W2) * GELU'(h) and its column sums against fp32 torch on the same
    bf16 operands (GELU' by autograd of the exact-erf GELU), and against the two-pass path it replaces."""
    M, K, N = shape
    g = torch.Generator().manual_seed(M + 3 * N)
    gy = torch.randn((M, K), generator=g).to(torch.bfloat16).cuda()
    w2 = (torch.randn((K, N), generator=g) / K ** 0.5).to(torch.bfloat16).cuda()   # fc2.weight: (out = K, in = N)
    h = (torch.randn((M, N), generator=g) * 1.5).to(torch.bfloat16).cuda()
    assert capi.lib().csb200_linear_dgelu_supported(M, N, K, capi.BF16)
    dh, gb = csbF._tc_dgelu(gy, w2, h)
    hf = h.float().requires_grad_(True)
    torch.nn.functional.gelu(hf).backward(gy.float() @ w2.float())
    ref = hf.grad
    assert rel_err(dh.float(), ref) < 2 ** -8
    # the bias gradient sums the fp32 values before their bf16 rounding: close to the fp32 reference, and within
    # the rounding noise (2^-9 per term) of a sum over the stored tensor
    assert rel_err(gb, ref.sum(0)) < 1e-3
    assert rel_err(gb, dh.float().sum(0)) < 1e-2
    # the two-pass path: cuBLAS input gradient (rounded to bf16) + flat GELU' pass
    da = torch.mm(gy, w2)
    flat = torch.empty_like(da)
    gb2 = torch.empty(N, dtype=torch.float32, device="cuda")
    lib = capi.lib()
    nws = lib.csb200_gelu_bwd_workspace_bytes(N)
    wsp = torch.empty(nws, dtype=torch.uint8, device="cuda")
    capi.check(lib.csb200_gelu_bwd(csbF._ptr(da), csbF._ptr(h), csbF._ptr(flat), csbF._ptr(gb2), csbF._ptr(wsp), nws, M, N,
                                   capi.BF16, csbF._vp(capi.stream_of(h))), "csb200_gelu_bwd")
    assert rel_err(dh.float(), ref) <= rel_err(flat.float(), ref) + 2 ** -10


@pytest.mark.parametrize("shape", [(4096, 64, 256), (2048, 128, 512), (1024, 256, 1024), (1000, 64, 128),
                                   (300 * 128 + 5, 64, 256)])
def test_saved_derivative_pair_matches_fp32_reference(shape):
    """csb200_linear_fwd with CSB200_EPI_GELU_SAVE_DERIV stores GELU'(h) (exact-erf form, C:190) instead of h, and
    csb200_linear_dact_bwd multiplies by it: same GELU output as the h-saving epilogue bit for bit, derivative and
    input gradient against fp32 torch autograd, and against the erf-recomputing csb200_linear_dgelu_bwd."""
    M, K, N = shape
    x, w, b = _operands(M, K, N, M + 5 * N)
    a, d = csbF._tc_linear(x, w, b, capi.EPI_GELU_SAVE_DERIV)
    a_ref, h = csbF._tc_linear(x, w, b, capi.EPI_GELU_SAVE)
    assert torch.equal(a, a_ref)
    hf = h.float().requires_grad_(True)
    torch.nn.functional.gelu(hf).sum().backward()
    assert (d.float() - hf.grad).abs().max() < 2 ** -7      # |GELU'| <= 1.13: one bf16 rounding (half an ulp of [1, 2) is 2^-8)
    g = torch.Generator().manual_seed(N)
    gy = torch.randn((M, K), generator=g).to(torch.bfloat16).cuda()
    w2 = (torch.randn((K, N), generator=g) / K ** 0.5).to(torch.bfloat16).cuda()
    dh, gb = csbF._tc_dgelu(gy, w2, d, deriv=True)
    ref = (gy.float() @ w2.float()) * hf.grad
    assert rel_err(dh.float(), ref) < 2 ** -7                # bf16 derivative x bf16 output rounding
    # column sums of thousands of signed terms, each carrying the 2^-9 rounding of the stored derivative
    assert rel_err(gb, ref.sum(0)) < 5e-3
    dh2, gb2 = csbF._tc_dgelu(gy, w2, h)
    assert rel_err(dh.float(), dh2.float()) < 2 ** -7 and rel_err(gb, gb2) < 5e-3


# (M, N, K): the four Linear layers of a CSWinBlock at the config-3 widths (stage 1-4 geometry, token counts
# reduced), ragged token counts, N below one tile, K spanning several column tiles
WGRAD_SHAPES = [(8192, 192, 64), (8192, 64, 64), (8192, 256, 64), (8192, 64, 256), (4096, 384, 128), (4096, 128, 512),
                (4096, 768, 256), (2048, 256, 1024), (1024, 1536, 512), (1024, 512, 2048), (1000, 64, 64),
                (77, 128, 128), (300 * 128 + 5, 192, 64), (19000, 1024, 256), (4096, 8, 64)]


@pytest.mark.parametrize("shape", WGRAD_SHAPES)
def test_wgrad_and_bias_gradient_match_fp32_reference(shape):
    """csb200_linear_wgrad: grad_W = g^T x and grad_b = column sums of g, fp32 accumulation of bf16 operands,
    against fp64 torch on the same operands; tolerance 1e-4 of the largest entry (fp32 summation noise over
    up to 38 405 tokens; the operands themselves are exact)."""
    M, N, K = shape
    gen = torch.Generator().manual_seed(M + N + K)
    g = torch.randn((M, N), generator=gen).to(torch.bfloat16).cuda()
    x = torch.randn((M, K), generator=gen).to(torch.bfloat16).cuda()
    assert csbF._tc_wgrad_ok(g, x, torch.float32)
    gw, gb = csbF._tc_wgrad(g, x, True)
    ref_w = g.double().t() @ x.double()
    ref_b = g.double().sum(0)
    assert rel_err(gw, ref_w) < 1e-4
    assert rel_err(gb, ref_b) < 1e-4
    gw2, none = csbF._tc_wgrad(g, x, False)
    assert none is None and rel_err(gw2, ref_w) < 1e-4
    cublas = torch.mm(g.t(), x, out_dtype=torch.float32)
    assert rel_err(gw, ref_w) <= rel_err(cublas, ref_w) + 1e-5


def test_wgrad_on_strided_operands():
    """Channel slices of wider token matrices (row stride > width), as the backward of a packed qkv sees them."""
    gen = torch.Generator().manual_seed(5)
    big_g = torch.randn((3000, 320), generator=gen).to(torch.bfloat16).cuda()
    big_x = torch.randn((3000, 256), generator=gen).to(torch.bfloat16).cuda()
    g, x = big_g[:, 64:256], big_x[:, 128:]
    gw, gb = csbF._tc_wgrad(g, x, True)
    assert rel_err(gw, g.double().t() @ x.double()) < 1e-4 and rel_err(gb, g.double().sum(0)) < 1e-4


def test_wgrad_into_the_zero_arena_of_a_deferred_block():
    """csb200_linear_wgrad_acc: inside ``deferred_sums(zero_arena_numel=...)`` grad_W / grad_b are slices of ONE buffer
    zeroed once (no memset nodes per call); a first block without an arena measures the demand, an exhausted arena
    falls back to the self-zeroing call."""
    gen = torch.Generator().manual_seed(11)
    g = torch.randn((4096, 192), generator=gen).to(torch.bfloat16).cuda()
    x = torch.randn((4096, 64), generator=gen).to(torch.bfloat16).cuda()
    ref_w, ref_b = g.double().t() @ x.double(), g.double().sum(0)
    with csbF.deferred_sums("cuda") as first:
        gw, gb = csbF._tc_wgrad(g, x, True)
    assert first.arena_demand >= 192 * 64 + 192 and rel_err(gw, ref_w) < 1e-4
    n0 = pkg.capi.launch_count()
    with csbF.deferred_sums("cuda", zero_arena_numel=first.arena_demand) as block:
        gw1, gb1 = csbF._tc_wgrad(g, x, True)      # served by the arena
        gw2, gb2 = csbF._tc_wgrad(g, x, True)      # arena exhausted: standalone buffers
    assert block.arena_demand == 2 * first.arena_demand
    assert gw1.untyped_storage().data_ptr() == gb1.untyped_storage().data_ptr()
    assert gw2.untyped_storage().data_ptr() != gb2.untyped_storage().data_ptr()
    for a, b in ((gw1, gb1), (gw2, gb2)):
        assert rel_err(a, ref_w) < 1e-4 and rel_err(b, ref_b) < 1e-4
    assert pkg.capi.launch_count() - n0 == 2
