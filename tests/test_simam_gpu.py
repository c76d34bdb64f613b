"""GPU: SimAM kernels vs the oracle, through the C ABI (csb200_simam_fwd / _bwd).

Tolerances: fp32 <= 1e-5 relative (north_star); bf16 <= 2 bf16 ulps of the largest magnitude
(the kernel computes in fp32 from bf16 inputs and rounds once; the oracle is evaluated in fp64 on the
same bf16 inputs and is NOT rounded, so half an ulp is rounding and the rest is headroom).
"""
import pytest
import torch

from conftest import rel_err
import cswin_simam_unet_b200 as pkg
from oracle import ops

pytestmark = pytest.mark.gpu
TOL = {torch.float32: 1e-5, torch.bfloat16: 2 ** -7}

# every dispatch path: warp-per-plane, CTA-per-plane, clusters of 2/4/8, generic (odd sizes)
NCHW_SHAPES = [(2, 3, 4, 4), (3, 5, 16, 16), (2, 4, 32, 32), (2, 3, 64, 64), (2, 2, 128, 128), (1, 2, 256, 256),
               (1, 1, 512, 512), (2, 3, 7, 7), (1, 2, 224, 224), (1, 3, 14, 14), (2, 2, 56, 56), (1, 1, 600, 600)]
NLC_SHAPES = [(2, 16, 64), (2, 1024, 256), (2, 4096, 128), (1, 16384, 64), (1, 256, 512), (2, 49, 24), (1, 3136, 64),
              (1, 100, 13), (1, 65536, 32),
              # streaming kernels: clusters of 8 / 4 / 1 with more images than co-resident clusters (persistent
              # loop, alternating partial buffers), a ragged last chunk count, 512 channels
              (18, 2048, 64), (37, 1024, 64), (150, 1024, 64), (3, 1536, 128), (2, 512, 512)]


def _check(x, layout, dtype, offset=0.0):
    x = (x + offset).to(dtype)
    gy = torch.randn_like(x, dtype=torch.float32).to(dtype)
    xd = x.cuda().requires_grad_(True)
    y = pkg.simam(xd, 1e-4, layout)
    y.backward(gy.cuda())
    x64 = x.double().requires_grad_(True)
    y64 = ops.simam(x64, 1e-4, layout)
    y64.backward(gy.double())
    assert y.dtype == dtype and y.shape == x.shape
    assert rel_err(y.float().cpu(), y64.detach()) < TOL[dtype]
    assert rel_err(xd.grad.float().cpu(), x64.grad) < TOL[dtype]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", NCHW_SHAPES)
def test_simam_nchw_matches_oracle(shape, dtype):
    torch.manual_seed(sum(shape))
    _check(torch.randn(shape), "NCHW", dtype)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", NLC_SHAPES)
def test_simam_nlc_matches_oracle(shape, dtype):
    torch.manual_seed(sum(shape))
    _check(torch.randn(shape) * 2 + 0.5, "NLC", dtype)


def test_simam_large_mean_keeps_fp32_parity():
    # naive sum(x^2) - sum(x)^2/n would lose every digit here
    torch.manual_seed(0)
    _check(torch.randn(2, 3, 64, 64), "NCHW", torch.float32, offset=300.0)
    _check(torch.randn(2, 1024, 64), "NLC", torch.float32, offset=300.0)
    _check(torch.randn(2, 4096, 64), "NLC", torch.float32, offset=300.0)  # streaming kernel


def test_simam_nlc_streaming_is_deterministic_and_image_independent():
    # config 3 skip shape x2; every image is reduced by its own cluster in a fixed order
    torch.manual_seed(3)
    x = (torch.randn(32, 4096, 128, device="cuda") * 3 - 1).bfloat16()
    y = pkg.simam(x, 1e-4, "NLC")
    assert torch.equal(y, pkg.simam(x, 1e-4, "NLC"))
    perm = torch.randperm(32, device="cuda")
    assert torch.equal(pkg.simam(x[perm].contiguous(), 1e-4, "NLC"), y[perm])
    for b in (0, 13, 31):
        ref = ops.simam(x[b:b + 1].double().cpu(), 1e-4, "NLC")
        assert rel_err(y[b:b + 1].float().cpu(), ref) < TOL[torch.bfloat16]


def test_simam_channels_last_takes_the_token_kernel_in_place():
    torch.manual_seed(1)
    x = torch.randn(2, 64, 32, 32)
    xc = x.cuda().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    y = pkg.simam(xc, 1e-4, "NCHW")
    assert y.is_contiguous(memory_format=torch.channels_last)
    y.sum().backward()
    x64 = x.double().requires_grad_(True)
    y64 = ops.simam(x64)
    y64.sum().backward()
    assert rel_err(y.cpu(), y64.detach()) < 1e-5 and rel_err(xc.grad.cpu(), x64.grad) < 1e-5


def test_simam_module_and_edge_cases():
    m = pkg.SimAM()
    assert list(m.state_dict()) == []
    assert m(torch.empty(0, 4, 8, 8, device="cuda")).shape == (0, 4, 8, 8)
    const = torch.full((1, 2, 8, 8), 3.0, device="cuda")  # zero variance: v = lambda, d = 0 -> sigmoid(0.5)
    assert rel_err(m(const).cpu(), ops.simam(const.cpu().double())) < 1e-6
    one = torch.randn(1, 2, 1, 1, device="cuda")  # H*W = 1: n = 0 -> NaN, exactly like the definition
    assert torch.isnan(m(one)).all() and torch.isnan(ops.simam(one.cpu())).all()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_simam_full_size_config2_plane_samples(dtype):
    # BASELINE config 2 largest call: (16, 64, 256, 256).  Oracle on sampled planes + global invariants.
    torch.manual_seed(2)
    x = torch.randn(16, 64, 256, 256, device="cuda").to(dtype)
    y = pkg.simam(x)
    for (b, c) in [(0, 0), (7, 31), (15, 63)]:
        ref = ops.simam(x[b:b + 1, c:c + 1].double().cpu())
        assert rel_err(y[b:b + 1, c:c + 1].float().cpu(), ref) < TOL[dtype]
    # the gate is a sigmoid of a value >= 0.5: sign preserved, 0.62|x| <= |y| <= |x|
    yf, xf = y.float(), x.float()
    assert bool(((yf.abs() <= xf.abs() * (1 + 2 ** -7)) & (yf.abs() >= xf.abs() * 0.62 * (1 - 2 ** -6))).all())
    assert torch.equal(torch.sign(yf), torch.sign(xf))
    # per-plane independence: permuting planes permutes the output bit-for-bit
    perm = torch.randperm(64, device="cuda")
    assert torch.equal(pkg.simam(x[:, perm].contiguous()), y[:, perm])


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", [(32, 16384, 64), (32, 4096, 128), (32, 1024, 256), (12, 3000, 128)])
def test_simam_grid_resident_kernels(shape, dtype, monkeypatch):
    """The three CSWin skips of BASELINE config 3 (and a ragged shape) through csb200_simam_fwd_ws / _bwd_ws with a
    workspace: the grid-resident kernels (csrc/simam_grid.cuh).  Checked against the fp64 oracle on sampled images,
    against the cluster kernels of the plain entry points on the whole batch, for bit-reproducibility, and for the
    workspace contract (counters left zero, so calls of different shapes can share it)."""
    from cswin_simam_unet_b200 import capi, functional as csbF
    monkeypatch.setattr(csbF, "SIMAM_GRID_KERNELS", True)  # off by default: slower than the cluster kernels so far
    B, L, C = shape
    assert capi.lib().csb200_simam_workspace_bytes(B, C, L, capi.NLC, capi._DTYPES[dtype]) > 0
    torch.manual_seed(L)
    x = (torch.randn(shape, device="cuda") * 2 + 0.5).to(dtype)
    gy = torch.randn(shape, device="cuda").to(dtype)
    xd = x.clone().requires_grad_(True)
    y = pkg.simam(xd, 1e-4, "NLC")
    y.backward(gy)
    # plain entry points (no workspace): the cluster kernels
    lib, vp = capi.lib(), lambda t: __import__("ctypes").c_void_p(t.data_ptr())
    y2, gx2 = torch.empty_like(x), torch.empty_like(x)
    stats = torch.empty(B * C, 2, device="cuda")
    st = __import__("ctypes").c_void_p(torch.cuda.current_stream().cuda_stream)
    code = capi._DTYPES[dtype]
    capi.check(lib.csb200_simam_fwd(vp(x), vp(y2), vp(stats), B, C, L, capi.NLC, code, 1e-4, st), "fwd")
    capi.check(lib.csb200_simam_bwd(vp(x), vp(gy), vp(stats), vp(gx2), B, C, L, capi.NLC, code, 1e-4, st), "bwd")
    tol = 2e-6 if dtype == torch.float32 else 2 ** -8  # two summation orders; bf16: at most one output ulp
    assert rel_err(y.float(), y2.float()) < tol and rel_err(xd.grad.float(), gx2.float()) < tol
    for b in sorted({0, B // 2, B - 1}):
        x64 = x[b:b + 1].double().cpu().requires_grad_(True)
        y64 = ops.simam(x64, 1e-4, "NLC")
        y64.backward(gy[b:b + 1].double().cpu())
        assert rel_err(y[b:b + 1].float().cpu(), y64.detach()) < TOL[dtype]
        assert rel_err(xd.grad[b:b + 1].float().cpu(), x64.grad) < TOL[dtype]
    # deterministic, and independent of an image's place in the batch
    xr = x.clone().requires_grad_(True)
    yr = pkg.simam(xr, 1e-4, "NLC")
    yr.backward(gy)
    assert torch.equal(yr, y) and torch.equal(xr.grad, xd.grad)
    perm = torch.randperm(B, device="cuda")
    assert torch.equal(pkg.simam(x[perm].contiguous(), 1e-4, "NLC"), y[perm])
    # the workspace is left clean: every per-image counter is zero again
    ws = csbF._simam_ws[x.device.index]
    torch.cuda.synchronize()
    assert int(ws[:32768].view(torch.int32).abs().sum()) == 0
