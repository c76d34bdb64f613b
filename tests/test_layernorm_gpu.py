"""GPU: fused token-major LayerNorm (csb200_layernorm_fwd / _bwd) vs torch.nn.functional.layer_norm
evaluated in fp64 on the same inputs.  fp32 <= 1e-5 relative; bf16 output / bf16 incoming gradient
<= 2^-7 (one bf16 rounding of the result)."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err
import cswin_simam_unet_b200 as pkg
from cswin_simam_unet_b200 import functional as csbF, modules

pytestmark = pytest.mark.gpu

# (rows..., C): every tiling (8/16/32/64/128 vectors per row), ragged row counts, CSWin widths
SHAPES = [(2, 49, 64), (3, 100, 32), (2, 3136, 64), (1, 784, 128), (2, 196, 256), (2, 49, 512), (5, 7, 128), (1, 1, 64)]


@pytest.mark.parametrize("x_dtype,out_dtype", [(torch.float32, torch.float32), (torch.float32, torch.bfloat16),
                                               (torch.bfloat16, torch.bfloat16)])
@pytest.mark.parametrize("shape", SHAPES)
def test_layernorm_matches_fp64_reference(shape, x_dtype, out_dtype):
    torch.manual_seed(sum(shape))
    C = shape[-1]
    x = (torch.randn(shape) * 2 + 0.3).to(x_dtype)
    w, b = torch.randn(C) * 0.5 + 1, torch.randn(C) * 0.2
    gy = torch.randn(shape).to(out_dtype)
    xd = x.cuda().requires_grad_(True)
    wd, bd = w.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    if not csbF.layer_norm_supported(xd):
        pytest.skip("width not tiled for this dtype")
    y = csbF.layer_norm(xd, wd, bd, 1e-5, out_dtype)
    y.backward(gy.cuda())
    x64 = x.double().requires_grad_(True)
    w64, b64 = w.double().requires_grad_(True), b.double().requires_grad_(True)
    F.layer_norm(x64, (C,), w64, b64, 1e-5).backward(gy.double())
    tol = 1e-5 if (x_dtype == torch.float32 and out_dtype == torch.float32) else 2 ** -7
    assert y.dtype == out_dtype and xd.grad.dtype == x_dtype
    assert rel_err(y.float().cpu(), F.layer_norm(x.double(), (C,), w.double(), b.double(), 1e-5)) < tol
    assert rel_err(xd.grad.float().cpu(), x64.grad) < tol
    assert rel_err(wd.grad.cpu(), w64.grad) < max(tol, 2e-5)
    assert rel_err(bd.grad.cpu(), b64.grad) < max(tol, 2e-5)


@pytest.mark.parametrize("autocast", [False, True])
@pytest.mark.parametrize("shape", [(2, 49, 64), (1, 784, 128), (4, 196, 256), (2, 49, 512)])
def test_deferred_linear_bias_gradient_rides_in_the_layernorm_backward(shape, autocast, no_tf32):
    """r = Linear(h) with defer_bias_grad, (s, y) = add_layer_norm(x, r, residual_bias=bias): every gradient,
    the Linear's bias included, equals the unfused torch graph; no csb200_colsum launch is needed."""
    torch.manual_seed(sum(shape) + 7)
    C = shape[-1]
    lin, ref = torch.nn.Linear(2 * C, C).cuda(), torch.nn.Linear(2 * C, C).cuda()
    ref.load_state_dict(lin.state_dict())
    norm, nref = torch.nn.LayerNorm(C).cuda(), torch.nn.LayerNorm(C).cuda()
    dt = torch.bfloat16 if autocast else torch.float32
    x = torch.randn(shape, device="cuda").to(dt)
    h = torch.randn(*shape[:-1], 2 * C, device="cuda").to(dt)
    gs, gy = torch.randn(shape, device="cuda").to(dt), torch.randn(shape, device="cuda").to(dt)
    xa, ha, xb, hb = (t.clone().requires_grad_(True) for t in (x, h, x, h))
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        r = modules.apply_linear(lin, ha, defer_bias_grad=True)
        s, y = modules.apply_add_norm(norm, xa, r, feeds_gemm=True, delta_bias=lin.bias)
        sb = xb + ref(hb)
        yb = nref(sb)
    torch.autograd.backward([s, y], [gs, gy.to(y.dtype)])
    torch.autograd.backward([sb, yb], [gs, gy.to(yb.dtype)])
    tol = 2e-2 if autocast else 2e-5
    assert rel_err(lin.bias.grad.cpu(), ref.bias.grad.cpu()) < tol
    assert rel_err(lin.weight.grad.cpu(), ref.weight.grad.cpu()) < tol
    assert rel_err(norm.weight.grad.cpu(), nref.weight.grad.cpu()) < tol
    assert rel_err(norm.bias.grad.cpu(), nref.bias.grad.cpu()) < tol
    assert rel_err(xa.grad.float().cpu(), xb.grad.float().cpu()) < tol
    assert rel_err(ha.grad.float().cpu(), hb.grad.float().cpu()) < tol
    # the landing pad for deltas that do not end in a fused add + LayerNorm
    lin.zero_grad()
    r2 = csbF.route_bias_grad(modules.apply_linear(lin, h, defer_bias_grad=True), lin.bias)
    (r2.float() * gs.float()).sum().backward()
    assert rel_err(lin.bias.grad.cpu(), gs.float().reshape(-1, C).sum(0).cpu()) < tol


def test_apply_norm_routes_and_falls_back():
    ln = torch.nn.LayerNorm(64).cuda()
    x = torch.randn(2, 50, 64, device="cuda")
    n0 = pkg.capi.launch_count()
    y = modules.apply_norm(ln, x)
    assert pkg.capi.launch_count() == n0 + 1 and rel_err(y.cpu(), ln(x).cpu()) < 1e-5
    with torch.autocast("cuda", dtype=torch.bfloat16):
        assert modules.apply_norm(ln, x, feeds_gemm=True).dtype == torch.bfloat16
        assert modules.apply_norm(ln, x).dtype == torch.float32
    odd = torch.nn.LayerNorm(40).cuda()  # 40 fp32 channels = 10 vectors: not tiled -> module itself
    xo = torch.randn(3, 40, device="cuda")
    n0 = pkg.capi.launch_count()
    assert torch.equal(modules.apply_norm(odd, xo), odd(xo)) and pkg.capi.launch_count() == n0
    ident = torch.nn.Identity()
    assert modules.apply_norm(ident, x) is x


def test_layernorm_large_rows_deterministic():
    torch.manual_seed(0)
    x = torch.randn(32 * 4096, 128, device="cuda", requires_grad=True)
    w = torch.randn(128, device="cuda", requires_grad=True)
    b = torch.randn(128, device="cuda", requires_grad=True)
    g = torch.randn_like(x)
    outs = []
    for _ in range(2):
        x.grad = w.grad = b.grad = None
        csbF.layer_norm(x, w, b).backward(g)
        outs.append((x.grad.clone(), w.grad.clone(), b.grad.clone()))
    assert all(torch.equal(a, c) for a, c in zip(*outs))  # fixed-order reductions
    ref = F.layer_norm(x.detach().double(), (128,), w.detach().double(), b.detach().double())
    assert rel_err(csbF.layer_norm(x, w, b).detach().cpu(), ref.cpu()) < 1e-5


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("rows,cols", [(1000, 64), (4097, 192), (333, 768), (50, 1536), (7, 2048), (1, 8), (20000, 256)])
def test_column_sum_matches_fp64(rows, cols, dtype):
    torch.manual_seed(rows + cols)
    x = torch.randn(rows, cols).to(dtype)
    if not pkg.capi.lib().csb200_colsum_supported(cols, pkg.capi.dtype_code(x)):
        with pytest.raises(RuntimeError, match="not tiled"):  # > 256 vectors per row: refused, Linear uses ATen
            csbF.column_sum(x.cuda())
        return
    got = csbF.column_sum(x.cuda())
    assert rel_err(got.cpu(), x.double().sum(0)) < (1e-5 if dtype == torch.float32 else 1e-4)
    assert torch.equal(got, csbF.column_sum(x.cuda()))  # deterministic


@pytest.mark.parametrize("autocast", [False, True])
def test_linear_function_matches_torch(autocast):
    torch.manual_seed(0)
    lin = torch.nn.Linear(64, 192).cuda()
    ref = torch.nn.Linear(64, 192).cuda()
    ref.load_state_dict(lin.state_dict())
    x = torch.randn(2, 300, 64, device="cuda")
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    g = torch.randn(2, 300, 192, device="cuda")
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
        ya = modules.apply_linear(lin, xa)
        yb = ref(xb)
    assert ya.dtype == yb.dtype
    ya.backward(g.to(ya.dtype))
    yb.backward(g.to(yb.dtype))
    tol = 2 ** -7 if autocast else 1e-5
    assert rel_err(ya.float().cpu(), yb.float().cpu()) < tol
    assert rel_err(xa.grad.cpu(), xb.grad.cpu()) < tol
    assert rel_err(lin.weight.grad.cpu(), ref.weight.grad.cpu()) < tol
    assert rel_err(lin.bias.grad.cpu(), ref.bias.grad.cpu()) < tol
    assert lin.weight.grad.dtype == torch.float32 and xa.grad.dtype == torch.float32


@pytest.mark.parametrize("x_dtype,out_dtype", [(torch.float32, torch.float32), (torch.bfloat16, torch.bfloat16)])
@pytest.mark.parametrize("shape", [(2, 49, 64), (3, 100, 32), (1, 784, 128), (2, 196, 256), (2, 49, 512)])
def test_residual_add_fused_into_layernorm(shape, x_dtype, out_dtype):
    """(s, y) = (x + r, LN(x + r)) and its backward (gradient on s folded into the LN backward pass)
    against the unfused torch ops in fp64."""
    torch.manual_seed(sum(shape) + 1)
    C = shape[-1]
    x, r = (torch.randn(shape) * 2).to(x_dtype), torch.randn(shape).to(x_dtype)
    w, b = torch.randn(C) * 0.5 + 1, torch.randn(C) * 0.2
    gs, gy = torch.randn(shape).to(x_dtype), torch.randn(shape).to(out_dtype)
    xd, rd = x.cuda().requires_grad_(True), r.cuda().requires_grad_(True)
    wd, bd = w.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    if not csbF.layer_norm_supported(xd):
        pytest.skip("width not tiled for this dtype")
    s, y = csbF.add_layer_norm(xd, rd, wd, bd, 1e-5, out_dtype)
    torch.autograd.backward([s, y], [gs.cuda(), gy.cuda()])
    x64, r64 = x.double().requires_grad_(True), r.double().requires_grad_(True)
    w64, b64 = w.double().requires_grad_(True), b.double().requires_grad_(True)
    s64 = (x64 + r64)
    s_in = s64 if x_dtype == torch.float32 else s64 + (s.detach().double().cpu() - s64.detach())  # LN sees the ROUNDED sum
    y64 = F.layer_norm(s_in, (C,), w64, b64, 1e-5)
    torch.autograd.backward([s64, y64], [gs.double(), gy.double()])
    tol = 1e-5 if x_dtype == torch.float32 else 2 ** -7
    assert s.dtype == x_dtype and y.dtype == out_dtype
    assert rel_err(s.float().cpu(), s64.detach()) < tol
    assert rel_err(y.float().cpu(), y64.detach()) < tol
    assert rel_err(xd.grad.float().cpu(), x64.grad) < tol and torch.equal(xd.grad, rd.grad)
    assert rel_err(wd.grad.cpu(), w64.grad) < max(tol, 2e-5)
    assert rel_err(bd.grad.cpu(), b64.grad) < max(tol, 2e-5)


def test_stage_runner_equals_block_by_block(no_tf32):
    """run_blocks (adds fused into the next pre-norm) == calling the blocks one after another."""
    torch.manual_seed(5)
    blocks = torch.nn.ModuleList([modules.CSWinBlock(dim=64, reso=14, num_heads=2, split_size=2) for _ in range(3)]).cuda()
    x = torch.randn(2, 196, 64, device="cuda")
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    ya = modules.run_blocks(blocks, xa)
    ya.square().mean().backward()
    ga = [p.grad.clone() for p in blocks.parameters()]
    blocks.zero_grad()
    yb = xb
    for blk in blocks:
        yb = blk(yb)
    yb.square().mean().backward()
    assert rel_err(ya, yb) < 1e-5 and rel_err(xa.grad, xb.grad) < 2e-5
    for g1, p in zip(ga, blocks.parameters()):
        assert rel_err(g1, p.grad) < 5e-5


# (rows, C) of the all-bf16 path of the autocast train step at config-3-like sizes: ragged row counts, fewer row
# groups than resident warps, the stage shapes scaled down in rows
STREAM_SHAPES = [(4096, 64), (4096 + 34, 64), (8192 + 2, 128), (32768, 256), (4100, 256), (65536, 64), (20000, 128)]


@pytest.mark.parametrize("with_residual", [False, True])
@pytest.mark.parametrize("shape", STREAM_SHAPES)
def test_bf16_layernorm_and_fused_add_at_large_row_counts(shape, with_residual):
    """The all-bf16 path of the autocast train step against F.layer_norm in fp64: y, s = x + r, grad_x (with the
    residual-stream gradient summed inside), gamma / beta gradients and the deferred residual bias gradient."""
    rows, C = shape
    torch.manual_seed(rows + C)
    x = (torch.randn(rows, C) * 2 + 0.3).to(torch.bfloat16)
    r = torch.randn(rows, C).to(torch.bfloat16)
    w, b = torch.randn(C) * 0.5 + 1, torch.randn(C) * 0.2
    rb = torch.randn(C)
    gy, gs = torch.randn(rows, C).to(torch.bfloat16), torch.randn(rows, C).to(torch.bfloat16)
    xd, rd = x.cuda().requires_grad_(True), r.cuda().requires_grad_(True)
    wd, bd, rbd = w.cuda().requires_grad_(True), b.cuda().requires_grad_(True), rb.cuda().requires_grad_(True)
    x64, r64 = x.double().requires_grad_(True), r.double().requires_grad_(True)
    w64, b64 = w.double().requires_grad_(True), b.double().requires_grad_(True)
    if with_residual:
        s, y = csbF.add_layer_norm(xd, rd, wd, bd, 1e-5, torch.bfloat16, rbd)
        torch.autograd.backward([s, y], [gs.cuda(), gy.cuda()])
        s64 = (x64 + r64).to(torch.bfloat16).double()  # the statistics are taken from the ROUNDED sum
        s_ref = x64 + r64
        y64 = F.layer_norm(s_ref, (C,), w64, b64, 1e-5)
        torch.autograd.backward([s_ref, y64], [gs.double(), gy.double()])
        assert rel_err(s.float().cpu(), s64) < 2 ** -8
        assert rel_err(rd.grad.float().cpu(), r64.grad) < 2 ** -7
        assert rel_err(rbd.grad.cpu(), r64.grad.sum(0)) < 3e-3  # column sums of the rounded gradient
    else:
        y = csbF.layer_norm(xd, wd, bd, 1e-5, torch.bfloat16)
        y.backward(gy.cuda())
        y64 = F.layer_norm(x64, (C,), w64, b64, 1e-5)
        y64.backward(gy.double())
    assert rel_err(y.float().cpu(), y64.detach()) < 2 ** -7
    assert rel_err(xd.grad.float().cpu(), x64.grad) < 2 ** -7
    # gamma / beta gradients: fp32 sums of bf16-rounded terms over >= 4096 rows
    assert rel_err(wd.grad.cpu(), w64.grad) < 2 ** -7 and rel_err(bd.grad.cpu(), b64.grad) < 2 ** -7


def test_deferred_final_sums_equal_the_immediate_ones_bit_for_bit():
    """csb200_sum_rows_deferred / _flush (one launch per 120 recorded sums) against the immediate final kernels:
    LayerNorm parameter gradients (plain, fused add, with the residual-bias gradient), bias column sums and
    the raw entry point with more than 120 records of ragged sizes."""
    lib = pkg.capi.lib()
    torch.manual_seed(3)
    # raw records: 130 jobs -> two launches
    parts = [torch.randn(r, c, device="cuda") for r, c in [(1 + (7 * i) % 40, 8 * (1 + i % 9)) for i in range(130)]]
    outs = [torch.full((p.shape[1],), float("nan"), device="cuda") for p in parts]
    assert lib.csb200_sum_rows_discard() == 0 and lib.csb200_sum_rows_pending() == 0
    for p, o in zip(parts, outs):
        pkg.capi.check(lib.csb200_sum_rows_deferred(p.data_ptr(), p.shape[0], p.shape[1], p.shape[1], o.data_ptr()), "rec")
    assert lib.csb200_sum_rows_pending() == 130
    n0 = pkg.capi.launch_count()
    pkg.capi.check(lib.csb200_sum_rows_flush(torch.cuda.current_stream().cuda_stream), "flush")
    assert pkg.capi.launch_count() - n0 == 2 and lib.csb200_sum_rows_pending() == 0
    for p, o in zip(parts, outs):
        assert rel_err(o.cpu(), p.double().sum(0).cpu()) < 1e-5
    assert lib.csb200_sum_rows_deferred(0, 1, 8, 8, outs[0].data_ptr()) != 0  # null partials are refused

    # through autograd: the same ops inside / outside a deferred_sums block
    x = torch.randn(4, 3000, 256, device="cuda", dtype=torch.bfloat16)
    r = torch.randn_like(x)
    # (every parameter is used ONCE: a second contribution would be added to a vector that is not filled yet —
    # the documented precondition of deferred_sums)
    ps = [torch.randn(256, device="cuda", requires_grad=True) for _ in range(6)]
    w0, b0, w1, b1, rb, cb = ps
    g = torch.randn_like(x)

    def run(deferred):
        for p in ps:
            p.grad = None
        ctx = csbF.deferred_sums("cuda") if deferred else __import__("contextlib").nullcontext()
        n0 = pkg.capi.launch_count()
        with ctx:
            y0 = csbF.layer_norm(x, w0, b0, out_dtype=torch.bfloat16)
            s, y1 = csbF.add_layer_norm(x, r, w1, b1, out_dtype=torch.bfloat16, residual_bias=rb)
            y2 = csbF.route_bias_grad(x, cb)
            torch.autograd.backward([y0, s, y1, y2], [g, g, g, g])
        return [p.grad.clone() for p in ps], pkg.capi.launch_count() - n0

    now, n_now = run(False)
    later, n_later = run(True)
    assert all(torch.equal(a, c) for a, c in zip(now, later))
    assert n_later == n_now - 2  # three final launches became one
    with pytest.raises(RuntimeError, match="nest"):
        with csbF.deferred_sums("cuda"), csbF.deferred_sums("cuda"):
            pass
