"""GPU: attention dropout inside the stripe-attention kernels (`attn = self.attn_drop(attn)`, C:290; the
reference trains with attn_drop_rate = 0.3, C:930-932).

Bitwise parity with ATen's Philox stream is not a goal.  What is checked instead:
  * the mask the forward kernel WRITES is the stated function of (seed, call counter, stripe unit, query, key):
    replayed here with an independent numpy Philox4x32-10;
  * it has the stated statistics (keep probability 1 - round(256 p)/256) and changes from call to call;
  * forward output and all gradients equal the fp64 oracle of LePEAttention.forward evaluated WITH THAT MASK
    (so forward and backward used the same mask), for the CUDA-core engine and for both tcgen05 engines;
  * both engines draw the same mask for the same generator state;
  * a CUDA-graph replay draws a fresh mask each time.
"""
import numpy as np
import pytest
import torch

from conftest import rel_err
import cswin_simam_unet_b200 as pkg
from cswin_simam_unet_b200 import functional as csbF
from oracle import ops

pytestmark = pytest.mark.gpu
P = 0.3
THR = int(P * 256 + 0.5)            # 77
KEEP = 1.0 - THR / 256.0
KEEP_SCALE = 1.0 / KEEP


def philox4x32_10(key, ctr):
    """numpy Philox4x32-10 (Salmon et al. 2011): key (2,), ctr (..., 4) uint32 -> (..., 4) uint32."""
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    c = [ctr[..., i].astype(np.uint64) for i in range(4)]
    k0, k1 = np.uint64(key[0]), np.uint64(key[1])
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = np.uint64(M0) * c[0], np.uint64(M1) * c[2]
        c = [((p1 >> np.uint64(32)) ^ c[1] ^ k0) & mask, p1 & mask, ((p0 >> np.uint64(32)) ^ c[3] ^ k1) & mask, p0 & mask]
        k0, k1 = (k0 + np.uint64(W0)) & mask, (k1 + np.uint64(W1)) & mask
    return np.stack(c, -1).astype(np.uint32)


def _run(reso, dim, heads, split, B, dtype, engine, seed=1234, counter=5, last_stage=False):
    torch.manual_seed(reso + dim)
    blk = pkg.CSWinBlock(dim, reso, heads, split, qkv_bias=True, attn_drop=P, last_stage=last_stage).cuda().train()
    gen = torch.Generator().manual_seed(reso * 7 + dim)
    with torch.no_grad():
        for att in blk.attns:
            att.engine = engine
            att.get_v.weight.copy_(torch.randn(att.get_v.weight.shape, generator=gen) * 0.3)
            att.get_v.bias.copy_(torch.randn(att.get_v.bias.shape, generator=gen) * 0.1)
    L = reso * reso
    qkv = (torch.randn((B, L, 3 * dim), generator=gen) * 1.2).to(torch.bfloat16)
    gout = torch.randn((B, L, dim), generator=gen).to(torch.bfloat16)
    csbF.seed_attention_dropout(seed, counter=counter)
    csbF.KEEP_LAST_DROP_MASKS = True
    try:
        x = qkv.to(dtype).cuda().requires_grad_(True)
        out = blk.attend(x)
        out.backward(gout.to(dtype).cuda())
        masks = [m.cpu() for m in csbF.last_drop_masks]
    finally:
        csbF.KEEP_LAST_DROP_MASKS = False
    return blk, qkv, gout, out.detach().float().cpu(), x.grad.float().cpu(), masks


def _oracle(blk, qkv, gout, masks, reso, dim):
    q64 = qkv.double().requires_grad_(True)
    width = dim // len(blk.attns)
    refs, params = [], []
    for i, att in enumerate(blk.attns):
        cs = slice(i * width, (i + 1) * width)
        w64 = att.get_v.weight.detach().double().cpu().requires_grad_(True)
        b64 = att.get_v.bias.detach().double().cpu().requires_grad_(True)
        params.append((w64, b64))
        dense = ops.dense_drop_mask(masks[i], reso, reso, att.H_sp, att.W_sp, KEEP_SCALE)
        refs.append(ops.stripe_attention(q64[..., :dim][..., cs], q64[..., dim:2 * dim][..., cs], q64[..., 2 * dim:][..., cs],
                                         w64, b64, reso, reso, att.H_sp, att.W_sp, att.num_heads, prob_mask=dense))
    ref = torch.cat(refs, -1)
    ref.backward(gout.double())
    return ref.detach(), q64.grad, params


# (reso, dim, heads, split, B, dtype, engine): N = 49 / 28 on the CUDA-core engine in fp32 (config-1 shapes),
# N = 128 and N = 256 two-branch merged launches on the tcgen05 engines with several groups per CTA, and the
# single-branch full window
CASES = [(14, 64, 2, 2, 2, torch.float32, "simt"), (7, 64, 2, 7, 3, torch.float32, "simt"),
         (32, 64, 2, 4, 12, torch.bfloat16, "tcgen05"), (32, 256, 8, 8, 6, torch.bfloat16, "tcgen05"),
         (16, 64, 2, 16, 5, torch.bfloat16, "tcgen05"), (32, 64, 2, 4, 3, torch.bfloat16, "simt")]


@pytest.mark.parametrize("case", CASES)
def test_forward_and_backward_use_the_mask_the_kernel_wrote(case):
    reso, dim, heads, split, B, dtype, engine = case
    blk, qkv, gout, out, dqkv, masks = _run(reso, dim, heads, split, B, dtype, engine, last_stage=(reso == split))
    ref, dref, params = _oracle(blk, qkv, gout, masks, reso, dim)
    tol_f, tol_b = (1e-5, 1e-5) if dtype == torch.float32 else (2 ** -7, 3e-2)
    assert rel_err(out, ref) < tol_f
    assert rel_err(dqkv, dref) < tol_b
    for att, (w64, b64) in zip(blk.attns, params):
        assert rel_err(att.get_v.weight.grad, w64.grad) < max(tol_b, 2e-5)
        assert rel_err(att.get_v.bias.grad, b64.grad) < max(tol_b, 2e-5)
    # statistics of the valid bits: keep probability 1 - 77/256, within 5 sigma
    for att, m in zip(blk.attns, masks):
        N = att.H_sp * att.W_sp
        dense = ops.dense_drop_mask(m, reso, reso, att.H_sp, att.W_sp, 1.0)
        frac, n = dense.mean().item(), dense.numel()
        assert abs(frac - KEEP) < 5 * (KEEP * (1 - KEEP) / n) ** 0.5, (frac, KEEP, n)
        # rows are not copies of each other: the keep fraction of every query row is also binomial
        row = dense.mean(-1)
        assert (row - KEEP).abs().max().item() < 6 * (KEEP * (1 - KEEP) / N) ** 0.5 + 1e-9


def test_mask_is_philox4x32_10_of_the_stated_counter():
    """counter = (key block j >> 4, query i, unit, call counter), key = seed; unit = (((b nwy + wy) nwx + wx)
    heads + head) * 2 + branch; byte j & 15 of the 16 output bytes; keep iff byte >= round(256 p)."""
    reso, dim, heads, split, B = 14, 64, 2, 2, 2
    seed, counter = 987654321, 41
    blk, qkv, gout, out, dqkv, masks = _run(reso, dim, heads, split, B, torch.float32, "simt", seed, counter)
    for br, (att, m) in enumerate(zip(blk.attns, masks)):
        hs, ws, nh = att.H_sp, att.W_sp, att.num_heads
        N, nwy, nwx = hs * ws, reso // hs, reso // ws
        dense = ops.dense_drop_mask(m, reso, reso, hs, ws, 1.0).numpy()  # (B nW, heads, i, j)
        i, j = np.meshgrid(np.arange(N), np.arange(N), indexing="ij")
        for b in range(B):
            for wy in range(nwy):
                for wx in range(nwx):
                    for h in range(nh):
                        unit = ((((b * nwy + wy) * nwx + wx) * nh + h) << 1) | br
                        ctr = np.stack([j >> 4, i, np.full_like(i, unit), np.full_like(i, counter)], -1).astype(np.uint32)
                        rnd = philox4x32_10((seed & 0xFFFFFFFF, seed >> 32), ctr)
                        byte = (np.take_along_axis(rnd, ((j & 15) >> 2)[..., None], -1)[..., 0] >> ((j & 3) * 8)) & 0xFF
                        want = (byte >= THR).astype(np.float64)
                        got = dense[(b * nwy + wy) * nwx + wx, h]
                        assert np.array_equal(got, want), (br, b, wy, wx, h)


def test_engines_draw_the_same_mask_and_calls_differ():
    a = _run(32, 64, 2, 4, 3, torch.bfloat16, "tcgen05", seed=77, counter=3)[5]
    b = _run(32, 64, 2, 4, 3, torch.bfloat16, "simt", seed=77, counter=3)[5]
    c = _run(32, 64, 2, 4, 3, torch.bfloat16, "tcgen05", seed=77, counter=4)[5]
    for ma, mb, mc in zip(a, b, c):
        assert torch.equal(ma, mb)
        assert not torch.equal(ma, mc)
    assert not torch.equal(a[0], a[1])  # the two branches of a block do not share a mask


def test_eval_mode_and_p_zero_are_the_plain_path():
    torch.manual_seed(0)
    blk = pkg.CSWinBlock(64, 16, 2, 4, qkv_bias=True, attn_drop=P).cuda()
    ref = pkg.CSWinBlock(64, 16, 2, 4, qkv_bias=True, attn_drop=0.0).cuda()
    ref.load_state_dict(blk.state_dict())
    qkv = torch.randn(2, 256, 192, device="cuda")
    blk.eval()
    assert torch.equal(blk.attend(qkv), ref.attend(qkv))
    blk.train()
    assert not torch.equal(blk.attend(qkv), ref.attend(qkv))


def test_training_step_with_the_reference_hyperparameters_and_graph_replays_draw_new_masks():
    """C:930-932: drop_rate = attn_drop_rate = drop_path_rate = 0.3.  Under a captured CUDA graph the call
    counter lives on the device, so every replay draws new masks (the loss changes between replays on the
    SAME batch with lr = 0)."""
    torch.manual_seed(0)
    net = pkg.CSWinTransformer(img_size=64, split_size=[1, 2, 2, 2], simam=True, drop_rate=0.3, attn_drop_rate=0.3,
                               drop_path_rate=0.0).cuda()
    step = pkg.TrainStep(net, pkg.FusedAdamW(net.parameters(), lr=0.0, weight_decay=0.0), precision="bf16",
                         cuda_graph=True)
    x, y = pkg.synthetic_batch(2, 64, "cuda", seed=0)
    c0 = int(csbF.attention_dropout_state(x.device)[1])
    losses = [step(x, y).item() for _ in range(4)]
    assert all(np.isfinite(losses)) and len(set(losses)) == 4, losses
    assert int(csbF.attention_dropout_state(x.device)[1]) > c0
    net2 = pkg.CSWinTransformer(img_size=64, split_size=[1, 2, 2, 2], simam=True, drop_rate=0.3, attn_drop_rate=0.3,
                                drop_path_rate=0.3).cuda()
    step2 = pkg.TrainStep(net2, torch.optim.AdamW(net2.parameters(), lr=1e-3), precision="bf16")
    losses2 = [step2(x, y).item() for _ in range(6)]
    assert all(np.isfinite(losses2))
