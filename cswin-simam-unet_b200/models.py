"""Model classes with the reference's constructor signatures and ``state_dict`` keys.

``CSWinTransformer`` (train_cswinunet_segmentation.py, "C:", lines 489-688) and ``UNet``
(train_unet_segmentation.py, "U:", lines 221-250).  Swapping the constructor is the whole migration:
``forward(x: B x 3 x S x S) -> B x 1 x S x S`` probabilities, the 463 CSWin / 112 UNet state tensors
load unchanged.  The one new keyword is ``simam`` (default False == reference arithmetic): SimAM is a
parameter-free gate, so it never changes the keys.
"""
from typing import List, Sequence

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functional as csbF
from .modules import (CARAFE, CARAFE4, ConvEmbedTokens, CSWinBlock, Merge_Block, SimAM, apply_conv, apply_linear, apply_norm, run_blocks,
                      carafe_upsample, image_as_tokens, tokens_as_image, _side)


def _trunc_normal_linear_init(m: nn.Module):
    """Reference initialisation (C:607-614): Linear ~ trunc_normal(std=.02), zero bias; norms 1/0."""
    if isinstance(m, nn.Linear):
        nn.init.trunc_normal_(m.weight, std=.02)
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)
    elif isinstance(m, (nn.LayerNorm, nn.BatchNorm2d)):
        nn.init.constant_(m.bias, 0)
        nn.init.constant_(m.weight, 1.0)


class CSWinTransformer(nn.Module):
    """CSWin-UNet: conv stem -> 4 encoder stages (Merge_Block between) -> 4 mirrored decoder stages
    (CARAFE x2 between, skip concat + Linear) -> CARAFE x4 -> 1x1 conv -> sigmoid.

    ``patch_size``, ``hybrid_backbone`` and ``use_chk`` are accepted and unused, as in the reference.
    ``simam=True`` gates the three skip tensors with the fused SimAM kernel (token layout) before the
    concat at C:657,662,667.  ``attn_engine`` selects the stripe-attention engine ("auto", "simt",
    "tcgen05").
    """

    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1, embed_dim=64, depth=[1, 2, 9, 1],
                 split_size=[1, 2, 7, 7], num_heads=[2, 4, 8, 16], mlp_ratio=4., qkv_bias=True, qk_scale=None,
                 drop_rate=0., attn_drop_rate=0., drop_path_rate=0., hybrid_backbone=None,
                 norm_layer=nn.LayerNorm, use_chk=False, simam=False, simam_lambda=1e-4, attn_engine="auto"):
        super().__init__()
        self.use_chk = use_chk
        self.num_classes = num_classes
        self.num_features = self.embed_dim = embed_dim
        self.img_size = img_size
        depth, split_size, num_heads = list(depth), list(split_size), list(num_heads)
        resos = [img_size // 4, img_size // 8, img_size // 16, img_size // 32]  # C:518,527,537,548
        dims = [embed_dim * 2 ** i for i in range(4)]
        dpr = [float(r) for r in np.linspace(0, drop_path_rate, int(np.sum(depth)))]  # C:514
        first = np.concatenate([[0], np.cumsum(depth)[:-1]]).astype(int)

        def stage(s: int) -> nn.ModuleList:
            return nn.ModuleList([
                CSWinBlock(dim=dims[s], num_heads=num_heads[s], reso=resos[s], mlp_ratio=mlp_ratio,
                           qkv_bias=qkv_bias, qk_scale=qk_scale,
                           split_size=split_size[s] if s < 3 else split_size[-1],
                           drop=drop_rate, attn_drop=attn_drop_rate, drop_path=dpr[first[s] + i],
                           norm_layer=norm_layer, last_stage=(s == 3))
                for i in range(depth[s])])

        # registration order follows the reference so that parameters() / optimizer state line up
        self.stage1_conv_embed = nn.Sequential(nn.Conv2d(in_chans, embed_dim, 7, 4, 2), ConvEmbedTokens(),
                                               nn.LayerNorm(embed_dim))
        self.pos_drop = nn.Dropout(p=drop_rate)
        self.stage1 = stage(0)
        self.merge1 = Merge_Block(dims[0], dims[1])
        self.stage2 = stage(1)
        self.merge2 = Merge_Block(dims[1], dims[2])
        self.stage3 = stage(2)
        self.merge3 = Merge_Block(dims[2], dims[3])
        self.stage4 = stage(3)
        self.norm = norm_layer(dims[3])
        self.stage_up4 = stage(3)
        self.upsample4 = CARAFE(dims[3], dims[2])
        self.concat_linear4 = nn.Linear(512, 256)  # widths are hard-coded in the reference (C:568)
        self.stage_up3 = stage(2)
        self.upsample3 = CARAFE(dims[2], dims[1])
        self.concat_linear3 = nn.Linear(256, 128)
        self.stage_up2 = stage(1)
        self.upsample2 = CARAFE(dims[1], dims[0])
        self.concat_linear2 = nn.Linear(128, 64)
        self.stage_up1 = stage(0)
        self.upsample1 = CARAFE4(dims[0], 64)
        self.norm_up = norm_layer(embed_dim)
        self.output = nn.Conv2d(in_channels=embed_dim, out_channels=num_classes, kernel_size=1, bias=False)
        self.skip_gate = SimAM(simam_lambda, layout="NLC") if simam else None
        self.apply(_trunc_normal_linear_init)
        self.set_attn_engine(attn_engine)

    def set_attn_engine(self, engine: str):
        for m in self.modules():
            if hasattr(m, "H_sp"):
                m.engine = engine

    @torch.jit.ignore
    def no_weight_decay(self):
        return {'pos_embed', 'cls_token'}

    @torch.jit.ignore
    def no_weight_decay_keywords(self):
        return {'relative_position_bias_table'}

    def _gate(self, skip):
        return skip if self.skip_gate is None else self.skip_gate(skip)

    def forward_features(self, x):
        """Encoder + bottleneck (C:625-650).  Skips are returned, not stashed on the module, so the
        model is re-entrant (CUDA graphs, overlapping micro-batches)."""
        # channels-last stem: the 7x7 conv then emits NHWC, which IS the token layout (no transpose copy)
        stem = self.stage1_conv_embed
        x = self.pos_drop(apply_norm(stem[2], stem[1](apply_conv(stem[0], x.contiguous(memory_format=torch.channels_last)))))
        skips: List[torch.Tensor] = []
        for blocks, merge in ((self.stage1, self.merge1), (self.stage2, self.merge2), (self.stage3, self.merge3)):
            x = run_blocks(blocks, x)
            skips.append(self._gate(x))
            x = merge(x)
        x = run_blocks(self.stage4, x)
        return apply_norm(self.norm, x), skips

    def forward_up_features(self, x, skips: Sequence[torch.Tensor]):
        """Decoder with skip connections (C:653-672)."""
        plan = ((self.stage_up4, self.upsample4, self.concat_linear4, skips[2]),
                (self.stage_up3, self.upsample3, self.concat_linear3, skips[1]),
                (self.stage_up2, self.upsample2, self.concat_linear2, skips[0]))
        for blocks, upsample, fuse, skip in plan:
            x = run_blocks(blocks, x)
            x = apply_linear(fuse, torch.cat([skip, upsample(x)], dim=-1))
        x = run_blocks(self.stage_up1, x)
        return apply_norm(self.norm_up, x, feeds_gemm=True)

    def up_x4(self, x):
        """CARAFE x4 -> 1x1 `out` conv (64->64) -> 1x1 `output` conv (64->classes)  (C:674-682).

        Both 1x1 convs are linear, pointwise and adjacent, and the reassembly weights are shared by
        all channels, so the chain equals  reassemble(W_eff x) + b_eff  with
        W_eff = output.weight @ out.weight (classes x 64) and b_eff = output.weight @ out.bias:
        the two (B, 64, S, S) full-resolution tensors of the reference are never formed.
        """
        up = self.upsample1
        side = _side(x.shape[1])
        img = tokens_as_image(x, side, side)
        w_out = self.output.weight.flatten(1)  # (classes, 64)
        w_eff = (w_out @ up.out.weight.flatten(1)).unsqueeze(-1).unsqueeze(-1)  # (classes, 64, 1, 1)
        b_eff = w_out @ up.out.bias
        low = F.conv2d(img, w_eff)
        logits = carafe_upsample(low, img, up.down, up.encoder, up.up_factor, up.kernel_size)
        return csbF.add_channel_bias(logits, b_eff)

    def forward_logits(self, x):
        feats, skips = self.forward_features(x)
        return self.up_x4(self.forward_up_features(feats, skips))

    def forward(self, x):
        return torch.sigmoid(self.forward_logits(x).float())  # C:688


# ------------------------------------------------------------------------------------------------
# plain UNet (U:177-250)
# ------------------------------------------------------------------------------------------------
class DoubleConv(nn.Module):
    """(Conv3x3 -> BN -> ReLU) x 2 (U:177-191), optionally followed by the fused SimAM gate."""

    def __init__(self, in_channels, out_channels, simam=False, simam_lambda=1e-4):
        super().__init__()
        self.double_conv = nn.Sequential(
            nn.Conv2d(in_channels, out_channels, kernel_size=3, padding=1), nn.BatchNorm2d(out_channels),
            nn.ReLU(inplace=True),
            nn.Conv2d(out_channels, out_channels, kernel_size=3, padding=1), nn.BatchNorm2d(out_channels),
            nn.ReLU(inplace=True))
        self.gate = SimAM(simam_lambda, layout="NCHW") if simam else None

    def forward(self, x):
        x = self.double_conv(x)
        return x if self.gate is None else self.gate(x)


class Down(nn.Module):
    """MaxPool2d(2) then DoubleConv (U:194-204)."""

    def __init__(self, in_channels, out_channels, simam=False, simam_lambda=1e-4):
        super().__init__()
        self.maxpool_conv = nn.Sequential(nn.MaxPool2d(2), DoubleConv(in_channels, out_channels, simam, simam_lambda))

    def forward(self, x):
        return self.maxpool_conv(x)


class Up(nn.Module):
    """ConvTranspose2d x2, concat with the skip, DoubleConv (U:207-218)."""

    def __init__(self, in_channels, out_channels, simam=False, simam_lambda=1e-4):
        super().__init__()
        self.up = nn.ConvTranspose2d(in_channels, in_channels // 2, kernel_size=2, stride=2)
        self.conv = DoubleConv(in_channels, out_channels, simam, simam_lambda)

    def forward(self, x1, x2):
        return self.conv(torch.cat([x2, self.up(x1)], dim=1))


class UNet(nn.Module):
    """5-level UNet, widths 64..1024, sigmoid output (U:221-250).  ``simam=True`` inserts the fused
    SimAM gate after every DoubleConv (9 sites) — BASELINE config 2."""

    def __init__(self, n_channels=3, n_classes=1, simam=False, simam_lambda=1e-4):
        super().__init__()
        self.n_channels, self.n_classes = n_channels, n_classes
        kw = dict(simam=simam, simam_lambda=simam_lambda)
        self.inc = DoubleConv(n_channels, 64, **kw)
        self.down1, self.down2 = Down(64, 128, **kw), Down(128, 256, **kw)
        self.down3, self.down4 = Down(256, 512, **kw), Down(512, 1024, **kw)
        self.up1, self.up2 = Up(1024, 512, **kw), Up(512, 256, **kw)
        self.up3, self.up4 = Up(256, 128, **kw), Up(128, 64, **kw)
        self.outc = nn.Conv2d(64, n_classes, kernel_size=1)
        self.sigmoid = nn.Sigmoid()

    def forward_logits(self, x):
        x1 = self.inc(x)
        x2 = self.down1(x1)
        x3 = self.down2(x2)
        x4 = self.down3(x3)
        x = self.up1(self.down4(x4), x4)
        x = self.up2(x, x3)
        x = self.up3(x, x2)
        x = self.up4(x, x1)
        return self.outc(x)

    def forward(self, x):
        return self.sigmoid(self.forward_logits(x).float())
