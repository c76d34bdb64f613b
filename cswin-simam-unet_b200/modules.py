"""Drop-in ``nn.Module`` building blocks of the CSWin-UNet, re-designed around the csb200 kernels.

Constructor signatures, attribute names and therefore ``state_dict`` keys follow the reference
(train_cswinunet_segmentation.py = "C:") so a constructor swap is all a user needs; the forward
passes are new: token tensors stay (B, L, C) == NHWC end to end, stripe attention runs in one fused
kernel per branch on the packed qkv buffer, and convolutions see channels-last views instead of
transposed copies.
"""
import math
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functional as csbF


def _side(L: int) -> int:
    s = math.isqrt(L)
    if s * s != L:
        raise ValueError(f"token count {L} is not a square grid (the reference assumes H == W, C:250)")
    return s


def tokens_as_image(x: torch.Tensor, H: int, W: int) -> torch.Tensor:
    """(B, L, C) tokens -> logical (B, C, H, W) with channels-last strides; a view, never a copy."""
    B, L, C = x.shape
    return x.reshape(B, H, W, C).permute(0, 3, 1, 2)


def image_as_tokens(img: torch.Tensor) -> torch.Tensor:
    """(B, C, H, W) -> (B, H*W, C); free when `img` is channels-last."""
    B, C, H, W = img.shape
    return img.permute(0, 2, 3, 1).reshape(B, H * W, C)


def apply_norm(norm: nn.Module, x: torch.Tensor, feeds_gemm: bool = False) -> torch.Tensor:
    """Run a ``norm_layer`` instance.  A plain ``nn.LayerNorm`` over the channel dimension goes through
    the fused csb200 kernel (parameters stay in the module: same ``state_dict``); under bf16 autocast a
    norm whose only consumer is a Linear / conv (``feeds_gemm``) emits bf16 directly.  Anything else
    (custom norm_layer, CPU tensors, untiled widths) runs the module as given."""
    if type(norm) is nn.LayerNorm and x.is_cuda and len(norm.normalized_shape) == 1 \
            and norm.elementwise_affine and norm.bias is not None and csbF.layer_norm_supported(x):
        out_dtype = x.dtype
        if feeds_gemm and torch.is_autocast_enabled("cuda"):
            out_dtype = torch.get_autocast_dtype("cuda")
        return csbF.layer_norm(x, norm.weight, norm.bias, norm.eps, out_dtype)
    return norm(x)


def apply_add_norm(norm: nn.Module, x: torch.Tensor, delta: torch.Tensor, feeds_gemm: bool = False,
                   delta_bias: Optional[torch.Tensor] = None):
    """(x + delta, norm(x + delta)).  With a plain LayerNorm on a tiled width the residual add rides in
    the LayerNorm kernel (one pass instead of an add kernel plus a norm kernel, and in backward the
    residual-stream gradient is summed inside the LayerNorm backward pass).  ``delta_bias``: the bias of
    the Linear that produced `delta` with ``defer_bias_grad`` — its gradient is returned by the same
    backward pass (or by a column-sum pass on the fallback path)."""
    if type(norm) is nn.LayerNorm and x.is_cuda and delta.dtype == x.dtype and delta.shape == x.shape \
            and len(norm.normalized_shape) == 1 and norm.elementwise_affine and norm.bias is not None \
            and csbF.layer_norm_supported(x):
        out_dtype = x.dtype
        if feeds_gemm and torch.is_autocast_enabled("cuda"):
            out_dtype = torch.get_autocast_dtype("cuda")
        return csbF.add_layer_norm(x, delta, norm.weight, norm.bias, norm.eps, out_dtype, delta_bias)
    s = x + csbF.route_bias_grad(delta, delta_bias)
    return s, apply_norm(norm, s, feeds_gemm)


def _can_defer_bias(lin: nn.Module, x: torch.Tensor, *between: nn.Module) -> bool:
    """The bias gradient of `lin` can ride in the LayerNorm backward pass that consumes its output iff the
    output reaches the residual add unchanged (every module in between is an identity in this mode)."""
    if type(lin) is not nn.Linear or lin.bias is None or not x.is_cuda or not lin.bias.requires_grad \
            or not torch.is_grad_enabled():
        return False
    for m in between:
        if isinstance(m, nn.Identity) or (isinstance(m, nn.Dropout) and (m.p == 0.0 or not m.training)):
            continue
        return False
    return True


def run_blocks(blocks, x: torch.Tensor) -> torch.Tensor:
    """A stage of CSWinBlocks with every residual add fused into the LayerNorm that follows it: the
    Mlp branch of block i is added inside norm1 of block i + 1 (only the last add stays a plain add)."""
    pending, pending_bias = None, None
    for blk in blocks:
        if isinstance(blk, CSWinBlock):
            x, pending, pending_bias = blk.forward_fused(x, pending, pending_bias)
        else:
            x = blk(x if pending is None else x + csbF.route_bias_grad(pending, pending_bias))
            pending, pending_bias = None, None
    return x if pending is None else x + csbF.route_bias_grad(pending, pending_bias)


def apply_conv(conv: nn.Module, x: torch.Tensor) -> torch.Tensor:
    """A plain ``nn.Conv2d`` (groups 1, dilation 1, zero padding) runs cuDNN through csbF.conv2d, which
    takes the bias gradient as one flat column-sum pass; anything else runs the module as given."""
    if type(conv) is nn.Conv2d and x.is_cuda and conv.groups == 1 and conv.dilation == (1, 1) \
            and conv.padding_mode == "zeros" and not isinstance(conv.padding, str):
        return csbF.conv2d(x, conv.weight, conv.bias, conv.stride, conv.padding)
    return conv(x)


def apply_linear(lin: nn.Module, x: torch.Tensor, defer_bias_grad: bool = False) -> torch.Tensor:
    """A plain ``nn.Linear`` runs through csbF.linear (cuBLAS GEMMs + one-pass bias gradient)."""
    if type(lin) is nn.Linear and x.is_cuda:
        return csbF.linear(x, lin.weight, lin.bias, defer_bias_grad)
    assert not defer_bias_grad
    return lin(x)


class SimAM(nn.Module):
    """Parameter-free SimAM attention (Yang et al., ICML 2021) as ONE fused kernel per direction.

    Not part of the reference checkout; adds no ``state_dict`` keys.
    ``layout="NCHW"`` for (B, C, H, W) feature maps, ``"NLC"`` for (B, L, C) tokens.
    """

    def __init__(self, e_lambda: float = 1e-4, layout: str = "NCHW"):
        super().__init__()
        self.e_lambda = e_lambda
        self.layout = layout

    def extra_repr(self):
        return f"e_lambda={self.e_lambda}, layout={self.layout}"

    def forward(self, x):
        return csbF.simam(x, self.e_lambda, self.layout)


class DropPath(nn.Module):
    """Per-sample stochastic depth (the reference takes it from timm, C:14, C:344)."""

    def __init__(self, drop_prob: float = 0.0):
        super().__init__()
        self.drop_prob = float(drop_prob)

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        mask = torch.empty((x.shape[0],) + (1,) * (x.dim() - 1), dtype=x.dtype, device=x.device).bernoulli_(keep)
        return x * (mask / keep)


class Mlp(nn.Module):
    """Linear -> act -> drop -> Linear -> drop (C:180-196); dense GEMMs stay on cuBLAS."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features or in_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features or in_features, out_features or in_features)
        self.drop = nn.Dropout(drop)

    def forward(self, x, defer_fc2_bias_grad: bool = False):
        exact_gelu = type(self.act) is nn.GELU and self.act.approximate == "none"
        if exact_gelu and type(self.fc1) is nn.Linear and type(self.fc2) is nn.Linear \
                and (self.drop.p == 0.0 or not self.training) \
                and csbF.mlp_fused_supported(x, self.fc1.weight, self.fc1.bias, self.fc2.weight):
            # the activation rides in the epilogues of the tcgen05 GEMMs on both sides of it (C:188-196)
            return csbF.mlp_fused(x, self.fc1.weight, self.fc1.bias, self.fc2.weight, self.fc2.bias,
                                  defer_fc2_bias_grad)
        if type(self.fc1) is nn.Linear and exact_gelu \
                and csbF.linear_gelu_supported(x, self.fc1.weight, self.fc1.bias):
            hidden = csbF.linear_gelu(x, self.fc1.weight, self.fc1.bias)  # fc1 + GELU, fused passes
        else:
            hidden = self.act(apply_linear(self.fc1, x))
        return self.drop(apply_linear(self.fc2, self.drop(hidden), defer_fc2_bias_grad))


class LePEAttention(nn.Module):
    """One stripe-attention branch with the LePE depthwise term (C:220-298), fused.

    ``idx``: -1 full window, 0 vertical stripes (H_sp = resolution, W_sp = split_size), 1 horizontal.
    ``forward(qkv)`` accepts anything indexable as qkv[0], qkv[1], qkv[2] with (B, L, C') entries of
    arbitrary strides, as in the reference; CSWinBlock bypasses it and feeds the packed buffer to
    the kernel directly.  ``get_v`` only stores the depthwise weights: the convolution itself is
    evaluated inside the attention kernel from the V tile it already holds.
    """

    def __init__(self, dim, resolution, idx, split_size, dim_out=None, num_heads=9, attn_drop=0., proj_drop=0.,
                 qk_scale=None):
        super().__init__()
        if idx == -1:
            h_sp, w_sp = resolution, resolution
        elif idx == 0:
            h_sp, w_sp = resolution, split_size
        elif idx == 1:
            h_sp, w_sp = split_size, resolution
        else:  # the reference prints and calls exit(0) here (C:238-240); raising is the library-safe form
            raise ValueError(f"LePEAttention: idx must be -1, 0 or 1, got {idx}")
        self.dim, self.dim_out = dim, dim_out or dim
        self.resolution, self.split_size, self.num_heads = resolution, split_size, num_heads
        self.scale = qk_scale or (dim // num_heads) ** -0.5
        self.H_sp, self.W_sp = h_sp, w_sp
        self.get_v = nn.Conv2d(dim, dim, kernel_size=3, stride=1, padding=1, groups=dim)
        self.attn_drop = nn.Dropout(attn_drop)
        self.engine = "auto"

    def drop_p(self) -> float:
        """attn_drop (C:246, applied at C:290) is evaluated inside the attention kernels (Philox mask on the
        softmax probabilities, same mask in backward); the nn.Dropout module only carries p."""
        return float(self.attn_drop.p) if self.training else 0.0

    def branch(self, chan0: int) -> csbF.Branch:
        return csbF.Branch(self.H_sp, self.W_sp, self.num_heads, chan0, self.dim)

    def forward(self, qkv):
        q, k, v = qkv[0], qkv[1], qkv[2]
        B, L, C = q.shape
        if L != self.resolution * self.resolution:
            raise AssertionError("flatten img_tokens has wrong size")
        return csbF.stripe_attention(q, k, v, self.get_v.weight, self.get_v.bias, self.resolution, self.resolution,
                                     self.H_sp, self.W_sp, self.num_heads, self.scale, self.engine, self.drop_p())


class CSWinBlock(nn.Module):
    """pre-LN -> qkv -> cross-shaped stripe attention (2 branches, or 1 full window) -> proj ->
    residual -> LN -> Mlp -> residual (C:301-370)."""

    def __init__(self, dim, reso, num_heads, split_size, mlp_ratio=4., qkv_bias=False, qk_scale=None,
                 drop=0., attn_drop=0., drop_path=0., act_layer=nn.GELU, norm_layer=nn.LayerNorm,
                 last_stage=False):
        super().__init__()
        self.dim, self.num_heads = dim, num_heads
        self.patches_resolution, self.split_size, self.mlp_ratio = reso, split_size, mlp_ratio
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.norm1 = norm_layer(dim)
        last_stage = last_stage or reso == split_size  # C:317-318
        self.branch_num = 1 if last_stage else 2
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(drop)
        if last_stage:
            branches = [LePEAttention(dim, resolution=reso, idx=-1, split_size=split_size, num_heads=num_heads,
                                      dim_out=dim, qk_scale=qk_scale, attn_drop=attn_drop, proj_drop=drop)]
        else:
            branches = [LePEAttention(dim // 2, resolution=reso, idx=i, split_size=split_size,
                                      num_heads=num_heads // 2, dim_out=dim // 2, qk_scale=qk_scale,
                                      attn_drop=attn_drop, proj_drop=drop) for i in range(2)]
        self.attns = nn.ModuleList(branches)
        self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), out_features=dim,
                       act_layer=act_layer, drop=drop)
        self.norm2 = norm_layer(dim)
        self._scale = self.attns[0].scale

    def attend(self, qkv: torch.Tensor) -> torch.Tensor:
        """qkv: packed (B, L, 3C) -> (B, L, C); no slicing copies and no cat (cf. C:358-363)."""
        reso = self.patches_resolution
        width = self.dim // self.branch_num
        branches, params = [], []
        for i, att in enumerate(self.attns):
            branches.append(att.branch(i * width))
            params += [att.get_v.weight, att.get_v.bias]
        return csbF.cross_stripe_attention(qkv, reso, reso, branches, self._scale, params, self.attns[0].engine,
                                           self.attns[0].drop_p())

    def forward_fused(self, x, pending=None, pending_bias=None):
        """The block with its LAST residual add left pending: returns (x', delta, delta_bias) with
        block(x + pending) == x' + delta, so the caller can fuse `+ delta` into the next pre-norm.
        ``delta_bias`` (and ``pending_bias`` on the way in) is the bias of the Linear that produced the
        delta when its gradient has been deferred to the consumer of the delta (else None): the caller
        MUST pass both on to ``apply_add_norm`` / ``csbF.route_bias_grad``."""
        B, L, C = x.shape
        if L != self.patches_resolution ** 2:
            raise AssertionError("flatten img_tokens has wrong size")
        if pending is None:
            n1 = apply_norm(self.norm1, x, feeds_gemm=True)
        else:
            x, n1 = apply_add_norm(self.norm1, x, pending, feeds_gemm=True, delta_bias=pending_bias)
        # proj_drop exists but is never applied in the reference (C:366-367)
        defer = _can_defer_bias(self.proj, x, self.drop_path)
        attended = apply_linear(self.proj, self.attend(apply_linear(self.qkv, n1)), defer)
        x, n2 = apply_add_norm(self.norm2, x, self.drop_path(attended), feeds_gemm=True,
                               delta_bias=self.proj.bias if defer else None)
        if type(self.mlp) is Mlp and _can_defer_bias(self.mlp.fc2, x, self.mlp.drop, self.drop_path):
            return x, self.drop_path(self.mlp(n2, defer_fc2_bias_grad=True)), self.mlp.fc2.bias
        return x, self.drop_path(self.mlp(n2)), None

    def forward(self, x):
        x, delta, delta_bias = self.forward_fused(x)
        return x + csbF.route_bias_grad(delta, delta_bias)


class Merge_Block(nn.Module):
    """Stride-2 3x3 conv between stages + LayerNorm (C:373-388) on channels-last views."""

    def __init__(self, dim, dim_out, norm_layer=nn.LayerNorm):
        super().__init__()
        self.conv = nn.Conv2d(dim, dim_out, 3, 2, 1)
        self.norm = norm_layer(dim_out)

    def forward(self, x):
        side = _side(x.shape[1])
        return apply_norm(self.norm, image_as_tokens(apply_conv(self.conv, tokens_as_image(x, side, side))))


def carafe_kernels(img: torch.Tensor, down: nn.Conv2d, encoder: nn.Conv2d, up: int) -> torch.Tensor:
    """Kernel-prediction half of CARAFE (C:406-411): softmax over the k*k taps, (B, k*k, H*up, W*up)."""
    return torch.softmax(F.pixel_shuffle(encoder(down(img)), up), dim=1)


def carafe_reassemble(low: torch.Tensor, kern: torch.Tensor, up: int, k: int = 3) -> torch.Tensor:
    """out[b,c,h*up+y,w*up+x] = sum_tap kern[b,tap,h*up+y,w*up+x] * low[b,c,h+ky-k//2,w+kx-k//2]
    with zero padding (the content-aware reassembly, C:413-431).  low: (B, C, H, W)."""
    B, C, H, W = low.shape
    nb = F.unfold(low, k, padding=k // 2).reshape(B, C, k * k, H, W)
    kv = kern.reshape(B, k * k, H, up, W, up)
    out = torch.einsum("bcthw,bthywx->bchywx", nb, kv.to(nb.dtype))
    return out.reshape(B, C, H * up, W * up)


def carafe_upsample(low: torch.Tensor, img: torch.Tensor, down: nn.Conv2d, encoder: nn.Conv2d, up: int,
                    k: int = 3) -> torch.Tensor:
    """Kernel prediction (two convs on cuDNN) + fused softmax / reassembly kernel; on tensors the kernel
    does not take (CPU, channel counts that are neither 1 nor a multiple of 8, k != 3) the same maths
    runs as torch ops."""
    enc = apply_conv(encoder, apply_conv(down, img))
    if k == 3 and low.dtype == enc.dtype and csbF.carafe_supported(low):
        return csbF.carafe_reassemble(low, enc, up)
    return carafe_reassemble(low, torch.softmax(F.pixel_shuffle(enc, up), dim=1), up, k)


class CARAFE(nn.Module):
    """Content-aware upsampling (C:391-437): predict a softmax 3x3 kernel per output pixel, then
    reassemble each output pixel from the 3x3 neighbourhood of its source pixel, then a 1x1 conv.

    Re-ordered, not re-defined: reassembly acts on every channel with the same spatial weights and
    ``out`` is a 1x1 convolution, so the two commute —  out(reassemble(x)) == reassemble(W x) + b.
    Applying ``W`` at LOW resolution cuts its FLOPs by up^2 and (dim_out = dim/2) halves the
    channels that get upsampled; the bias is added after reassembly exactly as in the reference.
    """

    def __init__(self, dim, dim_out, kernel_size=3, up_factor=2):
        super().__init__()
        self.kernel_size, self.up_factor = kernel_size, up_factor
        self.down = nn.Conv2d(dim, dim // 4, 1)
        self.encoder = nn.Conv2d(dim // 4, up_factor ** 2 * kernel_size ** 2, kernel_size, 1, kernel_size // 2)
        self.out = nn.Conv2d(dim, dim_out, 1)

    def forward(self, x):
        side = _side(x.shape[1])
        img = tokens_as_image(x, side, side)
        low = F.conv2d(img, self.out.weight)  # bias deferred past the reassembly
        up = carafe_upsample(low, img, self.down, self.encoder, self.up_factor, self.kernel_size)
        return image_as_tokens(csbF.add_channel_bias(up, self.out.bias))


class CARAFE4(CARAFE):
    """CARAFE with up_factor 4 (C:440-486; identical to CARAFE except for the default)."""

    def __init__(self, dim, dim_out, kernel_size=3, up_factor=4):
        super().__init__(dim, dim_out, kernel_size, up_factor)


class ConvEmbedTokens(nn.Module):
    """'b c h w -> b (h w) c' (the einops Rearrange at C:506), parameter-free."""

    def forward(self, x):
        return image_as_tokens(x)
