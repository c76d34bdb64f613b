#!/usr/bin/env python
"""One profiled call of each hot kernel, bracketed by cudaProfilerStart/Stop, for
    ncu --set full --import-source on --clock-control none --profile-from-start off -o <rep> \\
        python benchmarks/ncu_targets.py attn s3        # or: simam nchw | simam nlc | gelu | layernorm | carafe up3 | carafe x4 | linear s3 | attn_long
Shapes are BASELINE config 3 (512^2, batch 32, bf16) call sites; config 2 for SimAM NCHW."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import cswin_simam_unet_b200 as pkg  # noqa: E402
from cswin_simam_unet_b200 import functional as csbF  # noqa: E402

STAGES = {"s1": (128, 1, 2, 64), "s2": (64, 2, 4, 128), "s3": (32, 8, 8, 256), "s4": (16, 16, 16, 512)}
what = sys.argv[1] if len(sys.argv) > 1 else "attn"
arg = sys.argv[2] if len(sys.argv) > 2 else "s3"
B = 32


def profiled(fn, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if what == "attn":
    reso, split, heads, C = STAGES[arg]
    blk = pkg.CSWinBlock(dim=C, reso=reso, num_heads=heads, split_size=split, last_stage=(arg == "s4")).cuda()
    L = reso * reso
    # two buffer sets larger than L2 together would be ideal; one call per profile range keeps it simple
    q = torch.randn(B, L, 3 * C, device="cuda").bfloat16().requires_grad_(True)
    g = torch.randn(B, L, C, device="cuda").bfloat16()
    params = [p for a in blk.attns for p in (a.get_v.weight, a.get_v.bias)]
    profiled(lambda: torch.autograd.grad(blk.attend(q), [q] + params, g))
elif what == "simam":
    if arg == "nchw":
        x = torch.randn(16, 64, 256, 256, device="cuda").bfloat16().requires_grad_(True)
        profiled(lambda: torch.autograd.grad(pkg.simam(x), [x], torch.ones_like(x)))
    else:
        x = torch.randn(32, 16384, 64, device="cuda").bfloat16().requires_grad_(True)
        profiled(lambda: torch.autograd.grad(pkg.simam(x, layout="NLC"), [x], torch.ones_like(x)))
elif what == "gelu":
    x = torch.randn(B, 1024, 256, device="cuda").bfloat16().requires_grad_(True)
    w = torch.randn(1024, 256, device="cuda").bfloat16().requires_grad_(True)
    b = torch.randn(1024, device="cuda").bfloat16().requires_grad_(True)
    profiled(lambda: torch.autograd.grad(csbF._LinearGeluFn.apply(x, w, b, torch.bfloat16), [x, w, b],
                                         torch.ones(B, 1024, 1024, device="cuda").bfloat16()))
elif what == "layernorm":
    x = torch.randn(B, 1024, 256, device="cuda").bfloat16().requires_grad_(True)
    r = torch.randn(B, 1024, 256, device="cuda").bfloat16().requires_grad_(True)
    w = torch.ones(256, device="cuda", requires_grad=True)
    b = torch.zeros(256, device="cuda", requires_grad=True)

    def f():
        s, y = csbF.add_layer_norm(x, r, w, b, 1e-5, torch.bfloat16)
        torch.autograd.grad([s, y], [x, r, w, b], [torch.ones_like(s), torch.ones_like(y)])
    profiled(f)
elif what == "carafe":
    # decoder call sites of config 3: up3 = upsample3 (64 ch, 64^2 -> 128^2), x4 = the final 1-channel 128^2 -> 512^2
    C, H, up = (64, 64, 2) if arg == "up3" else (1, 128, 4)
    low = torch.randn(B, C, H, H, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    enc = torch.randn(B, 9 * up * up, H, H, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last).requires_grad_(True)
    g = torch.randn(B, C, H * up, H * up, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
    profiled(lambda: torch.autograd.grad(csbF.carafe_reassemble(low, enc, up), [low, enc], g))
elif what == "linear":
    # the tcgen05 token-path GEMMs at a config-3 stage: arg = s1 | s2 | s3 (fc1 + GELU forward, dgelu backward, the
    # qkv Linear, and the qkv weight + bias gradient)
    from cswin_simam_unet_b200 import capi
    M, C = {"s1": (524288, 64), "s2": (131072, 128), "s3": (32768, 256)}[arg]
    x = torch.randn(M, C, device="cuda").bfloat16()
    w1 = (torch.randn(4 * C, C, device="cuda") / C ** 0.5).bfloat16()
    b1 = torch.randn(4 * C, device="cuda")
    w2 = (torch.randn(C, 4 * C, device="cuda") / C ** 0.5).bfloat16()
    wq = (torch.randn(3 * C, C, device="cuda") / C ** 0.5).bfloat16()
    bq = torch.randn(3 * C, device="cuda")
    g3 = torch.randn(M, 3 * C, device="cuda").bfloat16()

    def f():
        a, d = csbF._tc_linear(x, w1, b1, capi.EPI_GELU_SAVE_DERIV)   # what _MlpFn runs: GELU and GELU'(h)
        csbF._tc_dgelu(x, w2, d, deriv=True)
        csbF._tc_linear(x, wq, bq, capi.EPI_BIAS)
        csbF._tc_wgrad(g3, x, True)
    profiled(f)
elif what == "attn_long":
    # BASELINE config 5, split 8, stage 1: N = 2048 -> the key/value-tiled forward kernel (stripe_fwd_tc_kv)
    blk = pkg.CSWinBlock(dim=64, reso=256, num_heads=2, split_size=8).cuda()
    q = torch.randn(8, 256 * 256, 192, device="cuda").bfloat16()

    def f():
        with torch.no_grad():
            blk.attend(q)
    profiled(f)
print("done", what, arg)
