"""GPU: fused CARAFE reassembly (csb200_carafe_fwd / _bwd) vs the oracle's restatement of
CARAFE.forward (C:391-437) and vs the torch-op path of the same module."""
import pytest
import torch

from conftest import rel_err
import cswin_simam_unet_b200 as pkg
from cswin_simam_unet_b200 import functional as csbF, modules
from oracle import ops

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,C,H,up", [(2, 64, 8, 2), (1, 128, 6, 2), (2, 256, 4, 2), (2, 1, 16, 4), (1, 8, 5, 4),
                                      (1, 64, 16, 4), (3, 16, 7, 2),
                                      # tiled kernels (W % 16 == 0): tile widths 16 / 32 / 64, 1..32 lanes per pixel
                                      (1, 64, 32, 2), (1, 8, 64, 2), (1, 128, 16, 2), (1, 256, 16, 2), (1, 1, 128, 4),
                                      (2, 1, 32, 2), (1, 16, 48, 4), (2, 64, 64, 2)])
def test_reassembly_kernel_matches_torch_ops(B, C, H, up, dtype):
    torch.manual_seed(B + C + H + up)
    low = torch.randn(B, C, H, H).to(dtype)
    enc = (torch.randn(B, 9 * up * up, H, H) * 2).to(dtype)
    gout = torch.randn(B, C, H * up, H * up).to(dtype)
    a, e = low.cuda().requires_grad_(True), enc.cuda().requires_grad_(True)
    out = csbF.carafe_reassemble(a, e, up)
    out.backward(gout.cuda())
    a64, e64 = low.double().requires_grad_(True), enc.double().requires_grad_(True)
    ref = modules.carafe_reassemble(a64, torch.softmax(torch.nn.functional.pixel_shuffle(e64, up), dim=1), up)
    ref.backward(gout.double())
    tol = 1e-5 if dtype == torch.float32 else 2e-2  # bf16: the 9 softmax weights are stored in bf16
    assert out.shape == ref.shape
    assert rel_err(out.float().cpu(), ref.detach()) < tol
    assert rel_err(a.grad.float().cpu(), a64.grad) < tol
    assert rel_err(e.grad.float().cpu(), e64.grad) < tol


@pytest.mark.parametrize("cls,up", [(pkg.CARAFE, 2), (pkg.CARAFE4, 4)])
def test_carafe_module_matches_oracle(no_tf32, cls, up):
    torch.manual_seed(0)
    m = cls(64, 32).cuda()
    x = torch.randn(2, 36, 64)
    p = {("u." + k): v.detach().cpu().double() for k, v in m.state_dict().items()}
    want = ops.carafe(x.double(), p, "u.", up)
    xd = x.cuda().requires_grad_(True)
    got = m(xd)
    assert rel_err(got.detach().cpu(), want) < 1e-5
    got.sum().backward()
    x64 = x.double().requires_grad_(True)
    ops.carafe(x64, p, "u.", up).sum().backward()
    assert rel_err(xd.grad.cpu(), x64.grad) < 1e-5


def test_unsupported_channel_count_uses_torch_path():
    low = torch.randn(1, 12, 4, 4, device="cuda")  # 12 channels: neither 1 nor a multiple of 8
    assert not csbF.carafe_supported(low)
    with pytest.raises(RuntimeError, match="multiple of 8"):
        csbF.carafe_reassemble(low, torch.randn(1, 36, 4, 4, device="cuda"), 2)
