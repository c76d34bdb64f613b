"""Model-level CPU oracle: the reference's two networks restated as pure functions of a parameter
dict (the reference's own ``state_dict`` keys), torch CPU, any float dtype, autograd for backward.

Test infrastructure — see oracle/__init__.py.  "C:" = train_cswinunet_segmentation.py,
"U:" = train_unet_segmentation.py.  SimAM placement is a build decision (the reference has no SimAM,
SURVEY.md §0.2): ``simam=False`` reproduces the reference bit-for-bit in structure; ``simam=True``
applies ops.simam to the three CSWin skip tensors / after every UNet DoubleConv, exactly where the
product models do.
"""
import math
from dataclasses import dataclass, field
from typing import Dict, List

import torch
import torch.nn.functional as F

from . import ops

Params = Dict[str, torch.Tensor]


@dataclass
class CSWinConfig:
    """Constructor arguments of CSWinTransformer that change arithmetic (C:493-496)."""
    img_size: int = 224
    in_chans: int = 3
    num_classes: int = 1
    embed_dim: int = 64
    depth: List[int] = field(default_factory=lambda: [1, 2, 9, 1])
    split_size: List[int] = field(default_factory=lambda: [1, 2, 7, 7])
    num_heads: List[int] = field(default_factory=lambda: [2, 4, 8, 16])
    qk_scale: float = None
    simam: bool = False
    e_lambda: float = 1e-4


def _ln(x, p, prefix):
    return F.layer_norm(x, (x.shape[-1],), p[prefix + "weight"], p[prefix + "bias"], 1e-5)


def _lin(x, p, prefix):
    return F.linear(x, p[prefix + "weight"], p.get(prefix + "bias"))


def cswin_block(x, p, prefix, reso, heads, split, last_stage, qk_scale=None):
    """CSWinBlock.forward, C:349-370 (drop / drop_path = 0, the constructor defaults)."""
    B, L, C = x.shape
    if L != reso * reso:
        raise AssertionError("flatten img_tokens has wrong size")  # C:356
    qkv = _lin(_ln(x, p, prefix + "norm1."), p, prefix + "qkv.")  # (B, L, 3C), C:357-358
    q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
    if last_stage or reso == split:  # C:317-318: one full-window branch
        hs, ws = ops.branch_geometry(reso, -1, split)
        att = ops.stripe_attention(q, k, v, p[prefix + "attns.0.get_v.weight"], p[prefix + "attns.0.get_v.bias"],
                                   reso, reso, hs, ws, heads, qk_scale)
    else:  # two branches on the channel halves, C:360-363
        half = C // 2
        outs = []
        for idx in (0, 1):
            sl = slice(idx * half, (idx + 1) * half)
            hs, ws = ops.branch_geometry(reso, idx, split)
            outs.append(ops.stripe_attention(
                q[..., sl], k[..., sl], v[..., sl],
                p[f"{prefix}attns.{idx}.get_v.weight"], p[f"{prefix}attns.{idx}.get_v.bias"],
                reso, reso, hs, ws, heads // 2, qk_scale))
        att = torch.cat(outs, dim=2)
    x = x + _lin(att, p, prefix + "proj.")  # C:366-367
    h = _lin(_ln(x, p, prefix + "norm2."), p, prefix + "mlp.fc1.")
    h = _lin(F.gelu(h), p, prefix + "mlp.fc2.")  # Mlp, C:188-196
    return x + h


def merge_block(x, p, prefix):
    """Merge_Block.forward, C:379-388."""
    B, L, C = x.shape
    H = W = int(math.isqrt(L))
    img = x.transpose(1, 2).reshape(B, C, H, W)
    img = F.conv2d(img, p[prefix + "conv.weight"], p[prefix + "conv.bias"], stride=2, padding=1)
    return _ln(img.flatten(2).transpose(1, 2), p, prefix + "norm.")


def cswin_unet_logits(p: Params, x: torch.Tensor, cfg: CSWinConfig) -> torch.Tensor:
    """Pre-sigmoid output of CSWinTransformer.forward (C:625-688): (B, num_classes, S, S)."""
    S = cfg.img_size
    resos = [S // 4, S // 8, S // 16, S // 32]
    # stage1_conv_embed, C:504-508
    t = F.conv2d(x, p["stage1_conv_embed.0.weight"], p["stage1_conv_embed.0.bias"], stride=4, padding=2)
    t = _ln(t.flatten(2).transpose(1, 2), p, "stage1_conv_embed.2.")
    skips = []
    for s in range(4):  # encoder, C:630-648
        for i in range(cfg.depth[s]):
            t = cswin_block(t, p, f"stage{s + 1}.{i}.", resos[s], cfg.num_heads[s],
                            cfg.split_size[s] if s < 3 else cfg.split_size[-1], s == 3, cfg.qk_scale)
        if s < 3:
            skips.append(ops.simam(t, cfg.e_lambda, "NLC") if cfg.simam else t)
            t = merge_block(t, p, f"merge{s + 1}.")
    t = _ln(t, p, "norm.")
    for s in (3, 2, 1, 0):  # decoder, C:653-672
        for i in range(cfg.depth[s]):
            t = cswin_block(t, p, f"stage_up{s + 1}.{i}.", resos[s], cfg.num_heads[s],
                            cfg.split_size[s] if s < 3 else cfg.split_size[-1], s == 3, cfg.qk_scale)
        if s > 0:
            t = ops.carafe(t, p, f"upsample{s + 1}.", 2)
            t = _lin(torch.cat([skips[s - 1], t], dim=-1), p, f"concat_linear{s + 1}.")
    t = _ln(t, p, "norm_up.")
    t = ops.carafe(t, p, "upsample1.", 4)  # up_x4, C:674-682
    B = x.shape[0]
    img = t.reshape(B, S, S, -1).permute(0, 3, 1, 2)
    return F.conv2d(img, p["output.weight"])


def cswin_unet_forward(p: Params, x: torch.Tensor, cfg: CSWinConfig) -> torch.Tensor:
    return torch.sigmoid(cswin_unet_logits(p, x, cfg))  # C:688


# ---------------------------------------------------------------------------------------------
# plain UNet, U:177-250.  `buffers` holds BatchNorm running stats; training=True uses batch stats
# (and does NOT update the running buffers — the oracle is stateless).
# ---------------------------------------------------------------------------------------------
def _double_conv(x, p, prefix, training, simam, e_lambda):
    for conv, bn in ((0, 1), (3, 4)):  # Sequential indices, U:181-188
        x = F.conv2d(x, p[f"{prefix}{conv}.weight"], p[f"{prefix}{conv}.bias"], padding=1)
        x = F.batch_norm(x, p.get(f"{prefix}{bn}.running_mean"), p.get(f"{prefix}{bn}.running_var"),
                         p[f"{prefix}{bn}.weight"], p[f"{prefix}{bn}.bias"], training=training,
                         momentum=0.0, eps=1e-5)
        x = F.relu(x)
    return ops.simam(x, e_lambda, "NCHW") if simam else x


def unet_logits(p: Params, x: torch.Tensor, training: bool = True, simam: bool = False,
                e_lambda: float = 1e-4) -> torch.Tensor:
    """Pre-sigmoid output of UNet.forward (U:239-249)."""
    if training:  # batch statistics; keep the running buffers untouched
        p = {k: v for k, v in p.items() if not k.endswith(("running_mean", "running_var", "num_batches_tracked"))}
    feats = [_double_conv(x, p, "inc.double_conv.", training, simam, e_lambda)]
    for i in range(1, 5):  # Down, U:194-204
        feats.append(_double_conv(F.max_pool2d(feats[-1], 2), p, f"down{i}.maxpool_conv.1.double_conv.",
                                  training, simam, e_lambda))
    t = feats[4]
    for i in range(1, 5):  # Up, U:207-218
        t = F.conv_transpose2d(t, p[f"up{i}.up.weight"], p[f"up{i}.up.bias"], stride=2)
        t = _double_conv(torch.cat([feats[4 - i], t], dim=1), p, f"up{i}.conv.double_conv.",
                         training, simam, e_lambda)
    return F.conv2d(t, p["outc.weight"], p["outc.bias"])


def unet_forward(p, x, training=True, simam=False, e_lambda=1e-4):
    return torch.sigmoid(unet_logits(p, x, training, simam, e_lambda))  # U:250


# ---------------------------------------------------------------------------------------------
# deterministic synthetic parameters keyed by name (weights cannot travel as fixtures: 94 MB)
# ---------------------------------------------------------------------------------------------
def synth_params(shapes: Dict[str, tuple], seed: int = 0, dtype=torch.float32, style: str = "unit") -> Params:
    """Reproducible parameters from (name, shape) alone — same values on any machine / torch CPU.

    style "unit": Linear / conv weights ~ N(0, 1/fan_in) so activations stay O(1) through 26 blocks
    (a stress regime: every op contributes); norm weights near 1; biases small.
    style "init": the scale of the reference's own initialisation (C:607-614) — Linear ~ N(0, .02)
    clipped at 2 sigma with zero bias, norms 1/0, convolutions U(+-1/sqrt(fan_in)) like torch's default.
    Generated per key, so the reference model (golden generation) and the product model (tests) can
    both be filled without sharing constructor order.
    """
    import zlib
    out = {}
    for name in sorted(shapes):
        shape = tuple(shapes[name])
        gen = torch.Generator().manual_seed((zlib.crc32(name.encode()) + 7919 * seed) % (2 ** 31))
        if name.endswith("num_batches_tracked"):
            out[name] = torch.zeros(shape, dtype=torch.long)
            continue
        r = torch.randn(shape, generator=gen, dtype=torch.float64)
        if style == "init" and not name.endswith(("running_mean", "running_var")):
            if len(shape) == 1:
                is_norm = "norm" in name or name.startswith("stage1_conv_embed.2")
                if is_norm or len(shapes.get(name[:-4] + "weight", ())) == 2:  # norm or Linear bias
                    val = torch.ones(shape, dtype=torch.float64) if (is_norm and name.endswith("weight")) \
                        else torch.zeros(shape, dtype=torch.float64)
                else:  # conv bias
                    wshape = shapes[name[:-4] + "weight"]
                    fan_in = 1
                    for d in wshape[1:]:
                        fan_in *= d
                    val = (torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1) / math.sqrt(fan_in)
            elif len(shape) == 2:
                val = (0.02 * r).clamp(-0.04, 0.04)
            else:
                fan_in = 1
                for d in shape[1:]:
                    fan_in *= d
                val = (torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1) / math.sqrt(fan_in)
            out[name] = val.to(dtype)
            continue
        if name.endswith("running_mean"):
            val = 0.1 * r
        elif name.endswith("running_var"):
            val = 1.0 + 0.1 * r.abs()
        elif len(shape) == 1:  # biases and norm scales
            is_scale = name.endswith("weight")
            val = (1.0 + 0.05 * r) if is_scale else 0.05 * r
        else:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            val = r / math.sqrt(max(fan_in, 1))
        out[name] = val.to(dtype)
    return out


def cswin_param_shapes(cfg: CSWinConfig) -> Dict[str, tuple]:
    """state_dict keys and shapes of CSWinTransformer(cfg) — C:504-603 (463 tensors at the defaults)."""
    E = cfg.embed_dim
    dims = [E, 2 * E, 4 * E, 8 * E]
    sh = {"stage1_conv_embed.0.weight": (E, cfg.in_chans, 7, 7), "stage1_conv_embed.0.bias": (E,),
          "stage1_conv_embed.2.weight": (E,), "stage1_conv_embed.2.bias": (E,)}

    def block(prefix, dim, last):
        for ln in ("norm1", "norm2"):
            sh[f"{prefix}{ln}.weight"] = (dim,)
            sh[f"{prefix}{ln}.bias"] = (dim,)
        sh[prefix + "qkv.weight"] = (3 * dim, dim)
        sh[prefix + "qkv.bias"] = (3 * dim,)
        sh[prefix + "proj.weight"] = (dim, dim)
        sh[prefix + "proj.bias"] = (dim,)
        sh[prefix + "mlp.fc1.weight"] = (4 * dim, dim)
        sh[prefix + "mlp.fc1.bias"] = (4 * dim,)
        sh[prefix + "mlp.fc2.weight"] = (dim, 4 * dim)
        sh[prefix + "mlp.fc2.bias"] = (dim,)
        for b in range(1 if last else 2):
            cb = dim if last else dim // 2
            sh[f"{prefix}attns.{b}.get_v.weight"] = (cb, 1, 3, 3)
            sh[f"{prefix}attns.{b}.get_v.bias"] = (cb,)

    resos = [cfg.img_size // 4, cfg.img_size // 8, cfg.img_size // 16, cfg.img_size // 32]
    for s in range(4):
        split = cfg.split_size[s] if s < 3 else cfg.split_size[-1]
        last = s == 3 or resos[s] == split
        for i in range(cfg.depth[s]):
            block(f"stage{s + 1}.{i}.", dims[s], last)
            block(f"stage_up{s + 1}.{i}.", dims[s], last)
    for s in range(3):
        sh[f"merge{s + 1}.conv.weight"] = (dims[s + 1], dims[s], 3, 3)
        sh[f"merge{s + 1}.conv.bias"] = (dims[s + 1],)
        sh[f"merge{s + 1}.norm.weight"] = (dims[s + 1],)
        sh[f"merge{s + 1}.norm.bias"] = (dims[s + 1],)
    sh["norm.weight"] = (dims[3],)
    sh["norm.bias"] = (dims[3],)
    for s, up in ((4, 2), (3, 2), (2, 2), (1, 4)):
        d = dims[s - 1]
        dout = d // 2 if s > 1 else 64
        pre = f"upsample{s}."
        sh[pre + "down.weight"] = (d // 4, d, 1, 1)
        sh[pre + "down.bias"] = (d // 4,)
        sh[pre + "encoder.weight"] = (up * up * 9, d // 4, 3, 3)
        sh[pre + "encoder.bias"] = (up * up * 9,)
        sh[pre + "out.weight"] = (dout, d, 1, 1)
        sh[pre + "out.bias"] = (dout,)
    for s, (i, o) in ((4, (512, 256)), (3, (256, 128)), (2, (128, 64))):  # hard-coded widths, C:568,581,592
        sh[f"concat_linear{s}.weight"] = (o, i)
        sh[f"concat_linear{s}.bias"] = (o,)
    sh["norm_up.weight"] = (E,)
    sh["norm_up.bias"] = (E,)
    sh["output.weight"] = (cfg.num_classes, E, 1, 1)
    return sh
