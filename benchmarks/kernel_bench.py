#!/usr/bin/env python
"""Per-kernel roofline microbenchmarks (one JSON line per case) — SimAM against the HBM roofline,
stripe attention against the tensor and HBM rooflines.  Inputs rotate through enough distinct
buffers to exceed the 126 MB L2 between repeats; CUDA events on the launching stream.

    python benchmarks/kernel_bench.py simam        # BASELINE config 2 / 3 SimAM call sites
    python benchmarks/kernel_bench.py attn [--engine simt|tcgen05|auto]   # config 3 attention calls
"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import cswin_simam_unet_b200 as pkg  # noqa: E402
from cswin_simam_unet_b200 import functional as csbF  # noqa: E402

PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.isfile(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0,
                                                      "bf16_tflops_sustained": 1400.0}
L2_BYTES = 126 << 20


def time_ms(fn, nbuf, reps=20, warmup=3):
    """Average device time of fn(i) over `reps` calls that rotate through `nbuf` buffer sets.  The calls
    are captured into ONE CUDA graph and replayed, so the Python / launch overhead of the binding
    (~30 us per call, more than many of these kernels take) is not part of the number."""
    for i in range(max(warmup, nbuf)):
        fn(i % nbuf)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        fn(0)
    torch.cuda.current_stream().wait_stream(side)
    with torch.cuda.graph(g):
        for i in range(reps):
            fn(i % nbuf)
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def bench_simam(only_layout=None, only_dtype=None, first=None, workspace=True):
    # config 2 (UNet 256^2, B=16): DoubleConv outputs, NCHW; config 3 (CSWin 512^2, B=32): skips, NLC
    cases = [("NCHW", (16, 64, 256, 256)), ("NCHW", (16, 128, 128, 128)), ("NCHW", (16, 256, 64, 64)),
             ("NCHW", (16, 512, 32, 32)), ("NCHW", (16, 1024, 16, 16)),
             ("NLC", (32, 16384, 64)), ("NLC", (32, 4096, 128)), ("NLC", (32, 1024, 256))]
    cases = [c for c in cases if only_layout in (None, c[0])][:first]
    for dtype in (torch.bfloat16, torch.float32):
        if only_dtype not in (None, str(dtype)[6:]):
            continue
        for layout, shape in cases:
            numel = 1
            for s in shape:
                numel *= s
            nbytes = numel * (2 if dtype == torch.bfloat16 else 4)
            nbuf = max(2, -(-2 * L2_BYTES // nbytes))
            xs = [torch.randn(shape, device="cuda").to(dtype) for _ in range(nbuf)]
            gs = [torch.randn(shape, device="cuda").to(dtype) for _ in range(nbuf)]
            ys = [torch.empty_like(t) for t in xs]
            B, C, S = (shape[0], shape[1], shape[2] * shape[3]) if layout == "NCHW" else (shape[0], shape[2], shape[1])
            stats = [torch.empty(B * C, 2, device="cuda") for _ in range(nbuf)]
            lib, vp = pkg.capi.lib(), ctypes.c_void_p
            lay = pkg.capi.NCHW if layout == "NCHW" else pkg.capi.NLC
            code = pkg.capi.dtype_code(xs[0])
            # the caller-owned workspace of csb200_simam_*_ws (zeroed once): large NLC tensors then take the
            # grid-resident kernels; --no-workspace times the cluster kernels of the plain entry points
            need = lib.csb200_simam_workspace_bytes(B, C, S, lay, code) if workspace else 0
            ws = torch.zeros(max(need, 16), dtype=torch.uint8, device="cuda")
            wsp, wsn = (vp(ws.data_ptr()), need) if need else (None, 0)

            def fwd(i):  # straight through the C ABI, on torch's current stream
                st = vp(torch.cuda.current_stream().cuda_stream)
                pkg.capi.check(lib.csb200_simam_fwd_ws(vp(xs[i].data_ptr()), vp(ys[i].data_ptr()),
                                                       vp(stats[i].data_ptr()), B, C, S, lay, code, 1e-4, wsp, wsn, st),
                               "fwd")

            def bwd(i):
                st = vp(torch.cuda.current_stream().cuda_stream)
                pkg.capi.check(lib.csb200_simam_bwd_ws(vp(xs[i].data_ptr()), vp(gs[i].data_ptr()),
                                                       vp(stats[i].data_ptr()), vp(ys[i].data_ptr()), B, C, S, lay,
                                                       code, 1e-4, wsp, wsn, st), "bwd")
            ms_f = time_ms(fwd, nbuf)
            ms_b = time_ms(bwd, nbuf)
            for name, ms, mult in (("simam_fwd", ms_f, 2), ("simam_bwd", ms_b, 3)):
                gbs = mult * nbytes / ms / 1e6
                print(json.dumps({"kernel": name, "layout": layout, "shape": shape, "dtype": str(dtype)[6:],
                                  "us": round(ms * 1e3, 2), "algorithmic_GBps": round(gbs, 1),
                                  "frac_of_measured_hbm": round(gbs / PEAKS["hbm_gbs"], 3), "buffers": nbuf,
                                  "grid_resident_kernels": bool(need)}),
                      flush=True)
            del xs, gs, ys, stats


def bench_layernorm():
    # CSWin 512^2, B=32 norm sites: fp32 residual stream -> bf16 GEMM operand (autocast), and back
    for rows, C in ((32 * 16384, 64), (32 * 4096, 128), (32 * 1024, 256), (32 * 256, 512)):
        nbuf = max(2, -(-2 * L2_BYTES // (rows * C * 4)))
        xs = [torch.randn(rows, C, device="cuda", requires_grad=True) for _ in range(nbuf)]
        w = torch.randn(C, device="cuda", requires_grad=True)
        b = torch.randn(C, device="cuda", requires_grad=True)
        g = torch.randn(rows, C, device="cuda").bfloat16()
        def fwd(i):
            with torch.no_grad():
                csbF.layer_norm(xs[i], w, b, 1e-5, torch.bfloat16)

        def fwd_bwd(i):
            torch.autograd.grad(csbF.layer_norm(xs[i], w, b, 1e-5, torch.bfloat16), [xs[i], w, b], g)
        ms_f = time_ms(fwd, nbuf)
        ms_b = time_ms(fwd_bwd, nbuf) - ms_f

        def torch_fwd(i):
            with torch.no_grad():
                torch.nn.functional.layer_norm(xs[i], (C,), w, b).bfloat16()

        def torch_fwd_bwd(i):
            torch.autograd.grad(torch.nn.functional.layer_norm(xs[i], (C,), w, b).bfloat16(), [xs[i], w, b], g)
        ms_tf = time_ms(torch_fwd, nbuf)
        ms_tb = time_ms(torch_fwd_bwd, nbuf) - ms_tf
        for name, ms, bpe, ref in (("layernorm_fwd", ms_f, 6, ms_tf), ("layernorm_bwd", ms_b, 10, ms_tb)):
            gbs = rows * C * bpe / ms / 1e6
            print(json.dumps({"kernel": name, "rows": rows, "C": C, "io": "fp32 -> bf16", "us": round(ms * 1e3, 1),
                              "algorithmic_GBps": round(gbs, 1), "frac_of_measured_hbm": round(gbs / PEAKS["hbm_gbs"], 3),
                              "aten_us": round(ref * 1e3, 1)}), flush=True)
        del xs


def bench_attn(engine):
    # config 3 per-call shapes: (reso, split, heads_total, C)
    stages = [("s1", 128, 1, 2, 64), ("s2", 64, 2, 4, 128), ("s3", 32, 8, 8, 256), ("s4", 16, 16, 16, 512)]
    B = 32
    for dtype in (torch.bfloat16,) if engine == "tcgen05" else (torch.bfloat16, torch.float32):
        for name, reso, split, heads, C in stages:
            blk = pkg.CSWinBlock(dim=C, reso=reso, num_heads=heads, split_size=split, last_stage=(name == "s4")).cuda()
            for a in blk.attns:
                a.engine = engine
            L = reso * reso
            nbytes = B * L * 3 * C * (2 if dtype == torch.bfloat16 else 4)
            nbuf = max(2, -(-2 * L2_BYTES // nbytes))
            qs = [(torch.randn(B, L, 3 * C, device="cuda")).to(dtype).requires_grad_(True) for _ in range(nbuf)]
            g = torch.randn(B, L, C, device="cuda").to(dtype)
            outs = [None] * nbuf
            t = csbF.KernelTimer()

            def fwd(i):
                outs[i] = blk.attend(qs[i])

            def fwd_nograd(i):
                with torch.no_grad():
                    blk.attend(qs[i])
            ms_f = time_ms(fwd_nograd, nbuf)
            params = [p for a in blk.attns for p in (a.get_v.weight, a.get_v.bias)]

            def fwd_bwd(i):  # forward + backward inside one captured region (same stream)
                torch.autograd.grad(blk.attend(qs[i]), [qs[i]] + params, g)
            ms_b = time_ms(fwd_bwd, nbuf) - ms_f
            csbF.set_kernel_timer(t)
            fwd_bwd(0)
            csbF.set_kernel_timer(None)
            work = t.summary()
            for fam, ms in (("attn_fwd", ms_f), ("attn_bwd", ms_b)):
                fl, by = work[fam]["flops"], work[fam]["bytes"]
                N = reso * split if name != "s4" else reso * reso
                print(json.dumps({"kernel": fam, "stage": name, "N": N, "dtype": str(dtype)[6:], "engine": engine,
                                  "us": round(ms * 1e3, 1), "TFLOPs": round(fl / ms / 1e9, 2),
                                  "frac_of_bf16_sustained": round(fl / ms / 1e9 / PEAKS["bf16_tflops_sustained"], 4),
                                  "algorithmic_GBps": round(by / ms / 1e6, 1),
                                  "frac_of_measured_hbm": round(by / ms / 1e6 / PEAKS["hbm_gbs"], 3),
                                  "attainable_tensor_frac": round(min(1.0, (fl / by) / 210.0), 3)}), flush=True)
            del qs, outs


def bench_attn_long(engine):
    """BASELINE config 5 per-call shapes (1024^2 inference, batch 8): forward only, the long stripes of split 8
    and the 32 x 32 full window of the last stage (stripe_fwd_tc_kv), plus split 2 stage 1."""
    # (name, reso, split, heads, C, last_stage)
    stages = [("sw8_s1", 256, 8, 2, 64, False), ("sw8_s2", 128, 8, 4, 128, False), ("sw8_s3", 64, 8, 8, 256, False),
              ("s4_full", 32, 32, 16, 512, True), ("sw2_s1", 256, 2, 2, 64, False), ("sw1_s3", 64, 1, 8, 256, False)]
    B, dtype = 8, torch.bfloat16
    for name, reso, split, heads, C, last in stages:
        blk = pkg.CSWinBlock(dim=C, reso=reso, num_heads=heads, split_size=split, last_stage=last).cuda()
        for a in blk.attns:
            a.engine = engine
        L = reso * reso
        nbytes = B * L * 3 * C * 2
        nbuf = max(2, -(-2 * L2_BYTES // nbytes))
        qs = [(torch.randn(B, L, 3 * C, device="cuda")).to(dtype) for _ in range(nbuf)]

        def fwd_nograd(i):
            with torch.no_grad():
                blk.attend(qs[i])
        ms = time_ms(fwd_nograd, nbuf)
        N = blk.attns[0].H_sp * blk.attns[0].W_sp
        fl = 4.0 * N * 32 * B * L * heads
        by = 4.0 * B * L * C * 2
        print(json.dumps({"kernel": "attn_fwd", "stage": name, "N": N, "B": B, "L": L, "C": C, "engine": engine,
                          "us": round(ms * 1e3, 1), "TFLOPs": round(fl / ms / 1e9, 2),
                          "frac_of_bf16_sustained": round(fl / ms / 1e9 / PEAKS["bf16_tflops_sustained"], 4),
                          "algorithmic_GBps": round(by / ms / 1e6, 1),
                          "attainable_tensor_frac": round(min(1.0, (fl / by) / 210.0), 3)}), flush=True)
        del qs


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["simam", "attn", "attn_long", "layernorm"])
    ap.add_argument("--engine", default="auto")
    ap.add_argument("--layout", default=None, choices=["NCHW", "NLC"])
    ap.add_argument("--dtype", default=None, choices=["bfloat16", "float32"])
    ap.add_argument("--first", type=int, default=None, help="only the first K SimAM shapes")
    ap.add_argument("--no-workspace", action="store_true", help="SimAM: plain entry points (cluster kernels)")
    a = ap.parse_args()
    if a.what == "simam":
        bench_simam(a.layout, a.dtype, a.first, not a.no_workspace)
    elif a.what == "layernorm":
        bench_layernorm()
    elif a.what == "attn_long":
        bench_attn_long(a.engine)
    else:
        bench_attn(a.engine)
