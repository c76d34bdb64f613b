"""Op-level CPU oracle (torch, any float dtype incl. float64; differentiable through autograd).

Test infrastructure — see oracle/__init__.py.  Every function cites the reference lines it restates
("C:" = train_cswinunet_segmentation.py).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


# ---------------------------------------------------------------------------------------------
# SimAM — [external] definition; the reference checkout has none (parity unpinned by the reference)
# ---------------------------------------------------------------------------------------------
def simam(x: torch.Tensor, e_lambda: float = 1e-4, layout: str = "NCHW") -> torch.Tensor:
    """y = x * sigmoid(d / (4 (sum(d)/n + lambda)) + 0.5), d = (x - mean)^2, n = spatial - 1.

    layout "NCHW": x is (B, C, H, W), statistics over (H, W).
    layout "NLC" : x is (B, L, C) tokens, statistics over L.
    """
    dims = (2, 3) if layout == "NCHW" else (1,)
    spatial = 1
    for a in dims:
        spatial *= x.shape[a]
    n = spatial - 1
    d = (x - x.mean(dim=dims, keepdim=True)).pow(2)
    v = d.sum(dim=dims, keepdim=True) / n + e_lambda
    return x * torch.sigmoid(d / (4 * v) + 0.5)


def simam_backward_numpy(x: np.ndarray, g: np.ndarray, e_lambda: float = 1e-4, layout: str = "NCHW"):
    """Closed-form gradient, independent of autograd; float64 numpy."""
    x = x.astype(np.float64)
    g = g.astype(np.float64)
    ax = (2, 3) if layout == "NCHW" else (1,)
    S = np.prod([x.shape[a] for a in ax])
    n = S - 1
    mu = x.mean(axis=ax, keepdims=True)
    t = x - mu
    d = t * t
    v = d.sum(axis=ax, keepdims=True) / n + e_lambda
    sig = 1.0 / (1.0 + np.exp(-(d / (4 * v) + 0.5)))
    a = g * x * sig * (1 - sig)
    r1 = (a * d).sum(axis=ax, keepdims=True)
    r2 = (a * t).sum(axis=ax, keepdims=True) / (4 * v)
    dcoef = a / (4 * v) - r1 / (4 * v * v * n)
    return g * sig + 2 * t * dcoef - (2.0 / S) * r2


# ---------------------------------------------------------------------------------------------
# stripe partition — C:199-217 (img2windows / windows2img) and C:248-254 (im2cswin), as one reshape
# ---------------------------------------------------------------------------------------------
def tokens_to_stripes(t: torch.Tensor, H: int, W: int, hs: int, ws: int, heads: int) -> torch.Tensor:
    """(B, H*W, C) -> (B * nW, heads, hs*ws, C/heads); stripes ordered (b, y-block, x-block)."""
    B, L, C = t.shape
    if L != H * W:
        raise AssertionError("flatten img_tokens has wrong size")  # C:281
    if H % hs or W % ws:
        raise RuntimeError(f"token grid {H}x{W} not divisible by stripe {hs}x{ws}")  # C:204 view()
    t = t.reshape(B, H // hs, hs, W // ws, ws, heads, C // heads)
    t = t.permute(0, 1, 3, 5, 2, 4, 6)
    return t.reshape(B * (H // hs) * (W // ws), heads, hs * ws, C // heads)


def stripes_to_tokens(s: torch.Tensor, B: int, H: int, W: int, hs: int, ws: int) -> torch.Tensor:
    """inverse of tokens_to_stripes -> (B, H*W, C)."""
    nW, heads, N, hd = s.shape
    s = s.reshape(B, H // hs, W // ws, heads, hs, ws, hd).permute(0, 1, 4, 2, 5, 3, 6)
    return s.reshape(B, H * W, heads * hd)


def lepe(v: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, H: int, W: int, hs: int, ws: int):
    """Depthwise 3x3 on each stripe with zero padding at the STRIPE border — C:256-269, C:244.

    v: (B, L, C) tokens; weight (C, 1, 3, 3); bias (C,).  Returns (B, L, C).
    """
    B, L, C = v.shape
    img = v.reshape(B, H // hs, hs, W // ws, ws, C).permute(0, 1, 3, 5, 2, 4)  # B, ny, nx, C, hs, ws
    img = img.reshape(-1, C, hs, ws)
    out = F.conv2d(img, weight, bias, stride=1, padding=1, groups=C)
    out = out.reshape(B, H // hs, W // ws, C, hs, ws).permute(0, 1, 4, 2, 5, 3)
    return out.reshape(B, L, C)


def stripe_attention(q, k, v, lepe_w, lepe_b, H, W, hs, ws, heads, scale=None, return_lse=False, prob_mask=None):
    """LePEAttention.forward, C:271-298, on token-major q, k, v of shape (B, L, C').
    ``prob_mask``: (B * nW, heads, N, N) multiplier applied to the softmax probabilities — attn_drop (C:290) with
    a GIVEN mask (keep / (1 - p)), so a kernel's own Philox mask can be replayed here."""
    B, L, C = q.shape
    hd = C // heads
    scale = scale if scale is not None else hd ** -0.5  # C:231
    qs = tokens_to_stripes(q, H, W, hs, ws, heads) * scale  # C:283,287
    ks = tokens_to_stripes(k, H, W, hs, ws, heads)
    vs = tokens_to_stripes(v, H, W, hs, ws, heads)
    scores = qs @ ks.transpose(-2, -1)  # C:288
    prob = torch.softmax(scores, dim=-1)  # C:289
    if prob_mask is not None:
        prob = prob * prob_mask  # C:290
    ctx = stripes_to_tokens(prob @ vs, B, H, W, hs, ws)  # C:292-296
    out = ctx + lepe(v, lepe_w, lepe_b, H, W, hs, ws)
    if return_lse:
        lse = torch.logsumexp(scores, dim=-1)  # (B*nW, heads, N)
        lse = stripes_to_tokens(lse.unsqueeze(-1), B, H, W, hs, ws)  # (B, L, heads)
        return out, lse.permute(0, 2, 1).contiguous()  # (B, heads, L)
    return out


def dense_drop_mask(words: torch.Tensor, H: int, W: int, hs: int, ws: int, keep_scale: float) -> torch.Tensor:
    """The kernels' attention-dropout mask — int32 [B][heads][L (key token)][ceil(N / 32)], bit (query % 32) of
    word (query / 32) — as the dense (B * nW, heads, N query, N key) multiplier `keep * keep_scale`."""
    B, heads, L, NW = words.shape
    N = hs * ws
    w = words.to(torch.int64) & 0xFFFFFFFF
    bits = ((w.unsqueeze(-1) >> torch.arange(32, dtype=torch.int64)) & 1).reshape(B, heads, L, NW * 32)[..., :N]
    as_tokens = bits.permute(0, 2, 1, 3).reshape(B, L, heads * N).to(torch.float64)   # "channels" = (head, query)
    per_key = tokens_to_stripes(as_tokens, H, W, hs, ws, heads)                        # (B nW, heads, N key, N query)
    return per_key.transpose(-2, -1).contiguous() * keep_scale


def branch_geometry(resolution: int, idx: int, split_size: int):
    """(h_sp, w_sp) of a LePEAttention branch — C:232-242."""
    if idx == -1:
        return resolution, resolution
    if idx == 0:
        return resolution, split_size
    if idx == 1:
        return split_size, resolution
    raise ValueError(f"ERROR MODE {idx}")


# ---------------------------------------------------------------------------------------------
# CARAFE — C:391-486 (content-aware upsample; `up` = 2 for CARAFE, 4 for CARAFE4)
# ---------------------------------------------------------------------------------------------
def carafe(x, p, prefix, up, ksize=3):
    """x: (B, L, C) tokens -> (B, L*up*up, C_out) tokens.  p: parameter dict, keys prefix+'down.weight'..."""
    B, L, C = x.shape
    H = W = int(math.isqrt(L))
    img = x.transpose(1, 2).reshape(B, C, H, W)
    kt = F.conv2d(img, p[prefix + "down.weight"], p[prefix + "down.bias"])
    kt = F.conv2d(kt, p[prefix + "encoder.weight"], p[prefix + "encoder.bias"], padding=ksize // 2)
    kt = F.pixel_shuffle(kt, up)  # (B, k*k, H*up, W*up)
    kt = torch.softmax(kt, dim=1)
    # kernels of the up*up sub-pixels of every source pixel: (B, H, W, k*k, up*up)
    kt = kt.reshape(B, ksize * ksize, H, up, W, up).permute(0, 2, 4, 1, 3, 5).reshape(B, H, W, ksize * ksize, up * up)
    # 3x3 neighbourhoods with zero padding: (B, H, W, C, k*k)
    nb = F.unfold(img, ksize, padding=ksize // 2).reshape(B, C, ksize * ksize, H, W).permute(0, 3, 4, 1, 2)
    y = nb @ kt  # (B, H, W, C, up*up)
    y = y.reshape(B, H, W, C * up * up).permute(0, 3, 1, 2)
    y = F.pixel_shuffle(y, up)
    y = F.conv2d(y, p[prefix + "out.weight"], p[prefix + "out.bias"])
    return y.flatten(2).transpose(1, 2)
