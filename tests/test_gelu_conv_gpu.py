"""GPU: the fused Mlp front half (fc1 -> exact-erf GELU, csb200_gelu_fwd / _bwd, whose backward also
emits the fc1 bias gradient), the conv2d wrapper whose bias gradient is a csb200 column sum, and the
bf16 parameter shadows — each against the same maths in torch fp64 on identical inputs.
fp32 <= 1e-5 relative; bf16 <= 2^-7 (one bf16 rounding of the result)."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from conftest import rel_err
import cswin_simam_unet_b200 as pkg
from cswin_simam_unet_b200 import functional as csbF, modules

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("rows,cin,hidden", [(98, 64, 256), (3136, 64, 256), (49, 512, 2048), (7, 32, 8), (1000, 128, 512)])
def test_linear_gelu_matches_fp64_reference(rows, cin, hidden, dtype, no_tf32):
    torch.manual_seed(rows + hidden)
    x = torch.randn(2, rows, cin) * 1.5
    w, b = torch.randn(hidden, cin) * cin ** -0.5, torch.randn(hidden) * 0.3
    ga = torch.randn(2, rows, hidden)
    xd = x.cuda().to(dtype).requires_grad_(True)
    wd, bd = w.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    if not csbF.linear_gelu_supported(xd, wd, bd):
        pytest.skip("hidden width not tiled for this dtype")
    ctx = torch.autocast("cuda", dtype=torch.bfloat16) if dtype == torch.bfloat16 else torch.autocast("cuda", enabled=False)
    with ctx:
        a = csbF.linear_gelu(xd, wd, bd)
    assert a.dtype == dtype
    a.backward(ga.cuda().to(dtype))
    x64 = xd.detach().double().cpu().requires_grad_(True)
    w64, b64 = w.double().requires_grad_(True), b.double().requires_grad_(True)
    if dtype == torch.bfloat16:  # the GEMM sees the bf16-rounded weight and bias
        w64 = w.bfloat16().double().requires_grad_(True)
        b64 = b.bfloat16().double().requires_grad_(True)
    ref = F.gelu(F.linear(x64, w64, b64))
    ref.backward(ga.to(dtype).double())
    tol = 1e-5 if dtype == torch.float32 else 2 ** -7
    assert rel_err(a.float().cpu(), ref.detach()) < tol
    assert rel_err(xd.grad.float().cpu(), x64.grad) < (tol if dtype == torch.float32 else 3e-2)
    assert rel_err(wd.grad.cpu(), w64.grad) < (2e-5 if dtype == torch.float32 else 3e-2)
    assert rel_err(bd.grad.cpu(), b64.grad) < (2e-5 if dtype == torch.float32 else 3e-2)


def test_gelu_extremes_stay_finite():
    h = torch.tensor([[-40.0, -12.0, -6.0, -1e-3, 0.0, 1e-3, 6.0, 40.0]] * 4).repeat(1, 4).cuda()
    w = torch.eye(32, device="cuda").requires_grad_(True)
    b = torch.zeros(32, device="cuda", requires_grad=True)
    for dt in (torch.float32, torch.bfloat16):
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=dt == torch.bfloat16):
            a = csbF.linear_gelu(h.to(dt), w, b)
        ref = F.gelu(h.to(dt).double())
        assert torch.isfinite(a).all()
        assert (a.double() - ref).abs().max().item() < (1e-5 if dt == torch.float32 else 0.26)


def test_mlp_module_takes_the_fused_path_and_matches_torch(no_tf32):
    torch.manual_seed(3)
    mlp = modules.Mlp(64, 256).cuda()
    ref = nn.Sequential(nn.Linear(64, 256), nn.GELU(), nn.Linear(256, 64)).cuda()
    ref[0].load_state_dict(mlp.fc1.state_dict())
    ref[2].load_state_dict(mlp.fc2.state_dict())
    x = torch.randn(2, 196, 64, device="cuda")
    n0 = pkg.capi.launch_count()
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    ya, yb = mlp(xa), ref(xb)
    ya.sum().backward()
    yb.sum().backward()
    assert pkg.capi.launch_count() - n0 >= 4  # gelu fwd + bwd (2 launches) + fc2 colsum
    assert rel_err(ya, yb) < 1e-5 and rel_err(xa.grad, xb.grad) < 1e-5
    assert rel_err(mlp.fc1.bias.grad, ref[0].bias.grad) < 2e-5
    assert rel_err(mlp.fc1.weight.grad, ref[0].weight.grad) < 2e-5


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("cin,cout,k,stride,pad", [(64, 128, 3, 2, 1), (3, 64, 7, 4, 2), (64, 16, 1, 1, 0), (16, 36, 3, 1, 1)])
def test_conv2d_wrapper_matches_torch(cin, cout, k, stride, pad, dtype, no_tf32):
    torch.manual_seed(cin + cout)
    conv = nn.Conv2d(cin, cout, k, stride, pad).cuda()
    x = torch.randn(2, cin, 28, 28, device="cuda").contiguous(memory_format=torch.channels_last)
    xa, xb = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=dtype == torch.bfloat16):
        ya = modules.apply_conv(conv, xa)
        ga = torch.randn_like(ya)
        ya.backward(ga)
        got = [xa.grad.clone(), conv.weight.grad.clone(), conv.bias.grad.clone()]
        conv.zero_grad()
        yb = conv(xb)
        yb.backward(ga)
    tol = 1e-5 if dtype == torch.float32 else 2 ** -7
    assert ya.dtype == yb.dtype and rel_err(ya.float(), yb.float()) < tol
    assert rel_err(got[0], xb.grad) < tol and rel_err(got[1], conv.weight.grad) < max(tol, 2e-5)
    # the bias gradient is accumulated in fp32 over the rounded gradient: tighter than ATen's bf16 sum
    want = ga.double().sum((0, 2, 3))
    assert rel_err(got[2], want) < (2e-5 if dtype == torch.float32 else 2 ** -7)


def test_parameter_shadows_follow_the_optimizer_and_are_never_stale():
    torch.manual_seed(0)
    lin = nn.Linear(64, 64).cuda()
    opt = torch.optim.AdamW(lin.parameters(), lr=1e-2)
    masters, shadows = csbF.shadow_params(lin.parameters())
    assert csbF.cast_param(lin.weight, torch.bfloat16) is shadows[0]
    lin(torch.randn(4, 64, device="cuda")).sum().backward()
    opt.step()
    # the master changed in place: the shadow is stale and must not be served
    fresh = csbF.cast_param(lin.weight, torch.bfloat16)
    assert fresh is not shadows[0] and torch.equal(fresh, lin.weight.detach().bfloat16())
    csbF.refresh_shadows(masters, shadows)
    assert csbF.cast_param(lin.weight, torch.bfloat16) is shadows[0]
    assert torch.equal(shadows[0], lin.weight.detach().bfloat16())
    with torch.no_grad():
        lin.weight.mul_(2.0)
    assert csbF.cast_param(lin.weight, torch.bfloat16) is not shadows[0]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("C", [16, 64, 144, 36, 1])
def test_row_bias_and_bias_gradient_on_channels_last(C, dtype):
    """x + bias[None, :, None, None] through csb200_add_row_bias (tiled widths) or ATen, and the bias
    gradient through csb200_colsum — directly, by row folding (36, 1) or after compaction (a slice)."""
    torch.manual_seed(C)
    x = torch.randn(3, C, 20, 24, device="cuda").to(dtype).contiguous(memory_format=torch.channels_last)
    b = torch.randn(C, device="cuda", requires_grad=True)
    xa = x.clone().requires_grad_(True)
    y = csbF.add_channel_bias(xa, b)
    want = x.double() + b.detach().double().view(1, -1, 1, 1)
    assert rel_err(y.double(), want) < (1e-6 if dtype == torch.float32 else 2 ** -8)
    g = torch.randn_like(y)
    y.backward(g)
    assert torch.equal(xa.grad, g)
    assert rel_err(b.grad, g.double().sum((0, 2, 3))) < 2e-5
    # a channel slice of a wider channels-last gradient (what torch.cat's backward hands over)
    wide = torch.randn(3, 20, 24, 2 * C + 8, device="cuda").to(dtype).permute(0, 3, 1, 2)
    sl = wide[:, 8:8 + C]
    assert rel_err(csbF.channel_sum(sl), sl.double().sum((0, 2, 3))) < 2e-5


def test_prefetched_steps_equal_direct_steps():
    """TrainStep.prefetch + step() (H2D on a side stream under the previous step) == step(x, y)."""
    import copy
    torch.manual_seed(0)
    net_a = pkg.CSWinTransformer(img_size=64, split_size=[1, 2, 2, 2], simam=True).cuda()
    net_b = copy.deepcopy(net_a)
    batches = [pkg.synthetic_batch(2, 64, "cpu", seed=s, pin=True) for s in range(4)]
    losses = {}
    for name, net in (("direct", net_a), ("prefetch", net_b)):
        step = pkg.TrainStep(net, torch.optim.AdamW(net.parameters(), lr=1e-3), precision="fp32")
        out = []
        if name == "prefetch":
            step.prefetch(*batches[0])
        for i in range(4):
            if name == "direct":
                out.append(step(batches[i][0].cuda(), batches[i][1].cuda()).item())
            else:
                loss = step()
                if i + 1 < 4:
                    step.prefetch(*batches[i + 1])
                out.append(loss.item())
        losses[name] = out
    assert losses["direct"] == losses["prefetch"], losses
    with pytest.raises(RuntimeError):
        pkg.TrainStep(net_a, torch.optim.AdamW(net_a.parameters()), precision="fp32")()
