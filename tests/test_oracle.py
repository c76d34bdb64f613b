"""CPU: the oracle against the golden vectors produced by the unmodified reference, plus the
self-consistency of the pieces the reference cannot pin (SimAM)."""
import numpy as np
import pytest
import torch

from conftest import ATTN_GOLDEN, golden, rel_err
from oracle import models as om, ops, reference_shim


@pytest.mark.parametrize("name", ATTN_GOLDEN)
def test_stripe_attention_oracle_matches_reference_golden(name):
    g = golden(f"attn_{name}.npz")
    dim, reso, idx, split, heads, B, hs, ws = [int(v) for v in g["meta"]]
    assert (hs, ws) == ops.branch_geometry(reso, idx, split)
    qkv = torch.tensor(g["qkv"], dtype=torch.float64, requires_grad=True)
    w = torch.tensor(g["lepe_w"], dtype=torch.float64, requires_grad=True)
    b = torch.tensor(g["lepe_b"], dtype=torch.float64, requires_grad=True)
    out = ops.stripe_attention(qkv[0], qkv[1], qkv[2], w, b, reso, reso, hs, ws, heads)
    out.backward(torch.tensor(g["gout"], dtype=torch.float64))
    assert rel_err(out, g["out"]) < 1e-12
    assert rel_err(qkv.grad, g["dqkv"]) < 1e-12
    assert rel_err(w.grad, g["dw"]) < 1e-11
    assert rel_err(b.grad, g["db"]) < 1e-11


def test_block_oracle_matches_reference_golden():
    g = golden("block_dim64_reso8.npz")
    shapes = {k[5:]: v.shape for k, v in g.items() if k.startswith("grad.")}
    p = {k: v.requires_grad_(True) for k, v in om.synth_params(shapes, seed=3, dtype=torch.float64).items()}
    x = torch.tensor(g["x"], requires_grad=True)
    y = om.cswin_block(x, p, "", 8, 2, 2, False)
    y.backward(torch.tensor(g["gout"]))
    assert rel_err(y, g["y"]) < 1e-12
    assert rel_err(x.grad, g["dx"]) < 1e-12
    for k in shapes:
        assert rel_err(p[k].grad, g["grad." + k]) < 1e-10, k


@pytest.mark.parametrize("fname", ["cswin_64.npz", "cswin_224_config1.npz"])
def test_cswin_model_oracle_matches_reference_golden(fname):
    g = golden(fname)
    img, batch, seed = [int(v) for v in g["meta"][:3]]
    cfg = om.CSWinConfig(img_size=img, split_size=[int(v) for v in g["meta"][3:]])
    p = {k: v.requires_grad_(True) for k, v in om.synth_params(om.cswin_param_shapes(cfg), seed).items()}
    x, y = torch.tensor(g["x"]), torch.tensor(g["y"])
    logits = om.cswin_unet_logits(p, x, cfg)
    loss = torch.nn.functional.binary_cross_entropy(torch.sigmoid(logits), y)
    loss.backward()
    assert rel_err(logits, g["logits"]) < 2e-5  # fp32 CPU vs fp32 CPU, different op order
    assert abs(loss.item() - float(g["loss"])) < 1e-6
    names = [str(n) for n in g["grad_names"]]
    norms = np.array([p[n].grad.double().norm().item() for n in names])
    np.testing.assert_allclose(norms, g["grad_norms"], rtol=2e-4, atol=1e-9)
    for k in [k for k in g if k.startswith("grad.")]:
        assert rel_err(p[k[5:]].grad, g[k]) < 1e-4, k


def test_unet_oracle_matches_reference_golden():
    g = golden("unet_64.npz")
    from cswin_simam_unet_b200 import UNet
    shapes = {k: tuple(v.shape) for k, v in UNet().state_dict().items()}
    assert list(shapes) == [str(n) for n in g["shape_names"]]  # state_dict key contract, U:221-237
    p = {k: (v.requires_grad_(True) if v.is_floating_point() else v) for k, v in om.synth_params(shapes, 1).items()}
    logits = om.unet_logits(p, torch.tensor(g["x"]), training=True)
    loss = torch.nn.functional.binary_cross_entropy(torch.sigmoid(logits), torch.tensor(g["y"]))
    loss.backward()
    assert rel_err(logits, g["logits"]) < 1e-5
    names = [str(n) for n in g["grad_names"]]
    norms = np.array([p[n].grad.double().norm().item() for n in names])
    np.testing.assert_allclose(norms, g["grad_norms"], rtol=1e-3, atol=1e-9)


@pytest.mark.skipif(not reference_shim.available(), reason="/root/reference only exists in the build container")
def test_oracle_against_live_reference_module():
    ref = reference_shim.load("cswin")
    torch.manual_seed(3)
    mod = ref.LePEAttention(64, 12, 1, 3, num_heads=2).double()
    qkv = torch.randn(3, 2, 144, 64, dtype=torch.float64)
    out = ops.stripe_attention(qkv[0], qkv[1], qkv[2], mod.get_v.weight, mod.get_v.bias, 12, 12, 3, 12, 2)
    assert rel_err(out, mod(qkv)) < 1e-13


def test_reference_failure_modes_are_mirrored_by_the_oracle():
    # default split [1,2,7,7] does not divide 512/16 = 32: the reference raises RuntimeError (C:204)
    with pytest.raises(RuntimeError):
        ops.tokens_to_stripes(torch.zeros(1, 32 * 32, 32), 32, 32, 32, 7, 1)
    with pytest.raises(AssertionError):
        ops.tokens_to_stripes(torch.zeros(1, 10, 32), 4, 4, 4, 1, 1)  # L != H*W, C:281
    with pytest.raises(ValueError):
        ops.branch_geometry(8, 2, 2)  # "ERROR MODE", C:238-240


def test_stripe_partition_round_trip():
    t = torch.randn(2, 8 * 12, 64)
    for hs, ws in ((8, 3), (2, 12), (8, 12), (1, 1)):
        s = ops.tokens_to_stripes(t, 8, 12, hs, ws, 2)
        assert s.shape == (2 * (8 // hs) * (12 // ws), 2, hs * ws, 32)
        assert torch.equal(ops.stripes_to_tokens(s, 2, 8, 12, hs, ws), t)


@pytest.mark.parametrize("layout,shape", [("NCHW", (2, 3, 5, 7)), ("NLC", (2, 35, 3))])
def test_simam_analytic_backward_matches_autograd_fp64(layout, shape):
    # SimAM is NOT in the reference (parity unpinned): the closed form the kernels implement is
    # cross-checked against autograd of the public definition, including a large-mean input.
    g = torch.Generator().manual_seed(0)
    for offset in (0.0, 1000.0):
        x = (torch.randn(shape, generator=g, dtype=torch.float64) + offset).requires_grad_(True)
        gy = torch.randn(shape, generator=g, dtype=torch.float64)
        ops.simam(x, 1e-4, layout).backward(gy)
        mine = ops.simam_backward_numpy(x.detach().numpy(), gy.numpy(), 1e-4, layout)
        assert rel_err(mine, x.grad) < 1e-9


def test_simam_layouts_agree():
    x = torch.randn(2, 6, 4, 5, dtype=torch.float64)
    a = ops.simam(x, 1e-4, "NCHW")
    b = ops.simam(x.permute(0, 2, 3, 1).reshape(2, 20, 6), 1e-4, "NLC")
    assert rel_err(b.reshape(2, 4, 5, 6).permute(0, 3, 1, 2), a) < 1e-14
