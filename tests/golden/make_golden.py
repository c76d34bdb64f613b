"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference modules.

Run in the build container only (needs /root/reference; the GPU box has none):

    python tests/golden/make_golden.py

The reference ships no tests or fixtures of its own, so these vectors ARE the pin of
the oracle: every case is produced by the reference's own ``LePEAttention`` / ``CSWinBlock`` /
``CSWinTransformer`` / ``UNet`` (fp32 or fp64 on CPU, constructor-default dropout 0) and the oracle is
asserted against it before anything is written.  Model weights are not stored (94 MB): they come
from ``oracle.models.synth_params`` (a pure function of key names, shapes and a seed).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import models as om, ops, reference_shim  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

# (name, dim, resolution, idx, split, heads, batch) — stripe lengths N = 8 ... 256, all three idx modes
ATTN_CASES = [
    ("v_sw1_n8", 32, 8, 0, 1, 1, 2),
    ("h_sw2_n16", 32, 8, 1, 2, 1, 2),
    ("v_sw7_n98", 64, 14, 0, 7, 2, 1),
    ("full_n49", 64, 7, -1, 7, 2, 2),
    ("h_sw8_n128", 64, 16, 1, 8, 2, 1),
    ("full_n256", 64, 16, -1, 16, 2, 1),
    ("v_sw4_n128", 32, 32, 0, 4, 1, 1),
]


def attention_cases(ref):
    for name, dim, reso, idx, split, heads, B in ATTN_CASES:
        g = torch.Generator().manual_seed(sum(map(ord, name)))
        L = reso * reso
        mod = ref.LePEAttention(dim, reso, idx, split, num_heads=heads).double()
        with torch.no_grad():
            mod.get_v.weight.copy_((torch.randn(mod.get_v.weight.shape, generator=g, dtype=torch.float64) * 0.3).float())
            mod.get_v.bias.copy_((torch.randn(mod.get_v.bias.shape, generator=g, dtype=torch.float64) * 0.1).float())
        # (weights above are fp32-exact so the stored fp32 copies are the values the reference used)
        # q, k with enough spread that the softmax is far from uniform; values exactly representable in
        # bf16 so the same vectors drive the bf16 kernels without an input-rounding term
        qkv = torch.randn((3, B, L, dim), generator=g, dtype=torch.float64)
        qkv[:2] *= 1.5
        qkv = qkv.to(torch.bfloat16).to(torch.float64).requires_grad_(True)
        gout = torch.randn((B, L, dim), generator=g, dtype=torch.float64).to(torch.bfloat16).to(torch.float64)
        out = mod(qkv)
        out.backward(gout)
        hs, ws = ops.branch_geometry(reso, idx, split)
        # pin the oracle on this case before saving
        q2 = qkv.detach().clone().requires_grad_(True)
        w2 = mod.get_v.weight.detach().clone().requires_grad_(True)
        b2 = mod.get_v.bias.detach().clone().requires_grad_(True)
        o2 = ops.stripe_attention(q2[0], q2[1], q2[2], w2, b2, reso, reso, hs, ws, heads)
        o2.backward(gout)
        assert torch.allclose(o2, out, atol=1e-12), name
        assert torch.allclose(q2.grad, qkv.grad, atol=1e-12), name
        assert torch.allclose(w2.grad, mod.get_v.weight.grad, atol=1e-11), name
        assert torch.allclose(b2.grad, mod.get_v.bias.grad, atol=1e-11), name
        yield name, dict(
            meta=np.array([dim, reso, idx, split, heads, B, hs, ws], dtype=np.int64),
            qkv=qkv.detach().float().numpy(), lepe_w=mod.get_v.weight.detach().float().numpy(),
            lepe_b=mod.get_v.bias.detach().float().numpy(), gout=gout.float().numpy(),
            out=out.detach().numpy(), dqkv=qkv.grad.numpy(),
            dw=mod.get_v.weight.grad.numpy(), db=mod.get_v.bias.grad.numpy())


def block_case(ref):
    torch.manual_seed(5)
    blk = ref.CSWinBlock(dim=64, reso=8, num_heads=2, split_size=2, qkv_bias=True).double()
    sd = om.synth_params({k: tuple(v.shape) for k, v in blk.state_dict().items()}, seed=3, dtype=torch.float64)
    blk.load_state_dict(sd)
    g = torch.Generator().manual_seed(11)
    x = torch.randn((2, 64, 64), generator=g, dtype=torch.float64).requires_grad_(True)
    gout = torch.randn((2, 64, 64), generator=g, dtype=torch.float64)
    y = blk(x)
    y.backward(gout)
    p2 = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    y2 = om.cswin_block(x.detach(), p2, "", 8, 2, 2, False)
    assert torch.allclose(y2, y, atol=1e-12)
    grads = {("grad." + k): p.grad.numpy() for k, p in blk.named_parameters()}
    return dict(x=x.detach().numpy(), gout=gout.numpy(), y=y.detach().numpy(), dx=x.grad.numpy(), **grads)


def bf16_drift_case(ref, img_size, split, batch, seed):
    """fp32 logits of the reference at ITS OWN initialisation scale, plus how far the reference's own
    bf16 (CPU autocast) forward drifts from them — the yardstick for the bf16 tolerance."""
    cfg = om.CSWinConfig(img_size=img_size, split_size=split)
    net = ref.CSWinTransformer(img_size=img_size, split_size=split)
    params = om.synth_params(om.cswin_param_shapes(cfg), seed, style="init")
    net.load_state_dict(params)
    g = torch.Generator().manual_seed(200 + seed)
    x = torch.rand((batch, 3, img_size, img_size), generator=g)
    grabbed = {}
    net.output.register_forward_hook(lambda m, i, o: grabbed.__setitem__("logits", o.detach().float()))
    with torch.no_grad():
        net(x)
        fp32 = grabbed["logits"].clone()
        with torch.autocast("cpu", dtype=torch.bfloat16):
            net(x)
        bf16 = grabbed["logits"].clone()
    thr = fp32.median()
    # random-init logits are all negative, so "mask agreement at p > 0.5" (C:731) is vacuous.
    # The harness adds an output bias b_q = -quantile_q(logits) (the reference's `output` conv has none, C:603),
    # i.e. a foreground fraction of 1 - q, and records the reference's OWN bf16 agreement at each setting.
    qs = np.array(MASK_QUANTILES)
    biases = np.array([-torch.quantile(fp32.flatten().double(), float(q)).item() for q in qs])
    agree = np.array([((fp32 + b > 0) == (bf16 + b > 0)).float().mean().item() for b in biases])
    return dict(meta=np.array([img_size, batch, seed] + list(split), dtype=np.int64), x=x.numpy(),
                logits=fp32.numpy(), ref_bf16_max_abs=np.array((fp32 - bf16).abs().max().item()),
                ref_bf16_median_mask_agreement=np.array(((fp32 > thr) == (bf16 > thr)).float().mean().item()),
                mask_quantiles=qs, mask_biases=biases, ref_bf16_mask_agreement=agree)


MASK_QUANTILES = (0.5, 0.9, 0.98)


def model_case_512(ref, seed=0, batch=2):
    """SURVEY.md §8(c) 'full-model logits at 512^2 B=2': the reference CSWinTransformer at BASELINE config 3's
    geometry (split [1,2,8,8], stripes of 128 / 128 / 256 / 256 tokens — the shapes the tcgen05 engines tile),
    fp32 on CPU, weights at the reference's init scale.  Inputs are NOT stored (8 MB): the test regenerates them
    from the same seeded CPU generator (`inputs_512`)."""
    img_size, split = 512, [1, 2, 8, 8]
    cfg = om.CSWinConfig(img_size=img_size, split_size=split)
    net = ref.CSWinTransformer(img_size=img_size, split_size=split)
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    assert shapes == om.cswin_param_shapes(cfg)
    params = om.synth_params(shapes, seed, style="init")
    net.load_state_dict(params)
    x, y = inputs_512(seed, batch)
    grabbed = {}
    net.output.register_forward_hook(lambda m, i, o: grabbed.__setitem__("logits", o.detach()))
    loss = torch.nn.BCELoss()(net(x), y)  # C:936
    loss.backward()
    logits = grabbed["logits"]
    with torch.no_grad():
        assert (om.cswin_unet_logits(params, x, cfg) - logits).abs().max() < 2e-5
    named = dict(net.named_parameters())
    names = list(named)
    gnorm = np.array([named[k].grad.double().norm().item() for k in names])
    keep = ["output.weight", "stage1.0.attns.0.get_v.weight", "stage1.0.attns.1.get_v.bias", "stage2.1.qkv.weight",
            "stage3.0.qkv.bias", "stage3.5.attns.1.get_v.weight", "stage4.0.attns.0.get_v.weight",
            "stage_up3.7.mlp.fc1.weight", "stage_up2.1.proj.weight", "upsample1.out.bias", "concat_linear3.weight"]
    full = {("grad." + k): named[k].grad.numpy() for k in keep}
    return dict(meta=np.array([img_size, batch, seed] + split, dtype=np.int64), logits=logits.numpy(),
                loss=np.array(loss.item()), grad_norms=gnorm, grad_names=np.array(names), **full)


def inputs_512(seed, batch):
    g = torch.Generator().manual_seed(300 + seed)
    x = torch.rand((batch, 3, 512, 512), generator=g)
    y = (torch.rand((batch, 1, 512, 512), generator=g) > 0.5).float()
    return x, y


def model_case(ref, img_size, split, batch, seed):
    cfg = om.CSWinConfig(img_size=img_size, split_size=split)
    net = ref.CSWinTransformer(img_size=img_size, split_size=split)
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    assert shapes == om.cswin_param_shapes(cfg)
    params = om.synth_params(shapes, seed)
    net.load_state_dict(params)
    g = torch.Generator().manual_seed(100 + seed)
    x = torch.rand((batch, 3, img_size, img_size), generator=g)
    y = (torch.rand((batch, 1, img_size, img_size), generator=g) > 0.5).float()
    grabbed = {}
    net.output.register_forward_hook(lambda m, i, o: grabbed.__setitem__("logits", o.detach()))
    probs = net(x)  # reference forward, fp32 CPU
    loss = torch.nn.BCELoss()(probs, y)  # C:936
    loss.backward()
    logits = grabbed["logits"]
    with torch.no_grad():
        assert (om.cswin_unet_logits(params, x, cfg) - logits).abs().max() < 2e-5
    names = [k for k, _ in net.named_parameters()]
    gnorm = np.array([p.grad.double().norm().item() for _, p in net.named_parameters()])
    keep = ["output.weight", "stage1.0.attns.0.get_v.weight", "stage1.0.attns.1.get_v.bias", "stage3.0.qkv.bias",
            "stage_up2.1.proj.weight", "upsample1.out.bias", "concat_linear3.weight"]
    full = {("grad." + k): dict(net.named_parameters())[k].grad.numpy() for k in keep}
    return dict(meta=np.array([img_size, batch, seed] + list(split), dtype=np.int64), x=x.numpy(), y=y.numpy(),
                logits=logits.numpy(), loss=np.array(loss.item()), grad_norms=gnorm,
                grad_names=np.array(names), **full)


def unet_case(refu):
    net = refu.UNet()
    shapes = {k: tuple(v.shape) for k, v in net.state_dict().items()}
    params = om.synth_params(shapes, 1)
    net.load_state_dict(params)
    net.train()
    g = torch.Generator().manual_seed(7)
    x = torch.rand((2, 3, 64, 64), generator=g)
    y = (torch.rand((2, 1, 64, 64), generator=g) > 0.5).float()
    grabbed = {}
    net.outc.register_forward_hook(lambda m, i, o: grabbed.__setitem__("logits", o.detach()))
    loss = torch.nn.BCELoss()(net(x), y)
    loss.backward()
    with torch.no_grad():
        assert (om.unet_logits(params, x, True) - grabbed["logits"]).abs().max() < 1e-5
    names = [k for k, _ in net.named_parameters()]
    gnorm = np.array([p.grad.double().norm().item() for _, p in net.named_parameters()])
    return dict(x=x.numpy(), y=y.numpy(), logits=grabbed["logits"].numpy(), loss=np.array(loss.item()),
                grad_norms=gnorm, grad_names=np.array(names),
                shape_names=np.array(list(shapes)), shape_ranks=np.array([len(s) for s in shapes.values()]))


def main():
    assert reference_shim.available(), "needs /root/reference"
    torch.set_num_threads(os.cpu_count())
    ref = reference_shim.load("cswin")
    refu = reference_shim.load("unet")
    for name, blob in attention_cases(ref):
        np.savez(os.path.join(OUT, f"attn_{name}.npz"), **blob)
        print("attn", name, blob["out"].shape)
    np.savez(os.path.join(OUT, "block_dim64_reso8.npz"), **block_case(ref))
    np.savez(os.path.join(OUT, "cswin_64.npz"), **model_case(ref, 64, [1, 2, 2, 2], 2, 0))
    np.savez(os.path.join(OUT, "cswin_224_config1.npz"), **model_case(ref, 224, [1, 2, 7, 7], 2, 0))
    np.savez(os.path.join(OUT, "unet_64.npz"), **unet_case(refu))
    np.savez(os.path.join(OUT, "cswin_224_init_bf16.npz"), **bf16_drift_case(ref, 224, [1, 2, 7, 7], 2, 0))
    np.savez_compressed(os.path.join(OUT, "cswin_512_b2.npz"), **model_case_512(ref))
    # reference failure modes that the drop-in must mirror
    try:
        ref.CSWinTransformer(img_size=512)(torch.rand(1, 3, 512, 512))
        raise SystemExit("expected the default split_size to fail at 512")
    except RuntimeError as e:
        with open(os.path.join(OUT, "reference_512_default_split_error.txt"), "w") as f:
            f.write(str(e) + "\n")
    print("done")


if __name__ == "__main__":
    main()
