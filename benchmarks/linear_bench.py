#!/usr/bin/env python
"""csb200_linear_fwd (tcgen05) vs cuBLAS (+ the flat GELU pass) on the K = C GEMM shapes of BASELINE config 3
(512^2, batch 32).  Prints one JSON line per (shape, epilogue): microseconds (CUDA events over a CUDA-graph
replay of `reps` back-to-back launches on rotating buffers > L2), algorithmic GB/s and TFLOP/s."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from cswin_simam_unet_b200 import capi, functional as csbF  # noqa: E402

SHAPES = {  # name: (M, K, N)
    "s1.qkv": (524288, 64, 192), "s1.proj": (524288, 64, 64), "s1.fc1": (524288, 64, 256),
    "s2.qkv": (131072, 128, 384), "s2.proj": (131072, 128, 128), "s2.fc1": (131072, 128, 512),
    "s3.qkv": (32768, 256, 768), "s3.proj": (32768, 256, 256), "s3.fc1": (32768, 256, 1024),
}


def timed(fn, nbuf, reps=8):
    for i in range(3):
        fn(i % nbuf)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(reps):
            fn(i % nbuf)
    g.replay()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(5):
        g.replay()
    t1.record()
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) * 1e3 / (5 * reps)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.isfile(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0}
    for name, (M, K, N) in SHAPES.items():
        if args.only and args.only not in name:
            continue
        nbuf = max(2, int(300e6 // (2 * M * (K + 2 * N))) + 1)  # rotate over > 2x L2
        xs = [torch.randn(M, K, device="cuda").to(torch.bfloat16) for _ in range(nbuf)]
        w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
        b = torch.randn(N, device="cuda")
        bb = b.to(torch.bfloat16)
        modes = [("bias", capi.EPI_BIAS)] + ([("gelu+h", capi.EPI_GELU_SAVE), ("gelu+gelu'", capi.EPI_GELU_SAVE_DERIV),
                                              ("gelu", capi.EPI_GELU)] if "fc1" in name else [])
        for label, epi in modes:
            us = timed(lambda i: csbF._tc_linear(xs[i], w, b, epi), nbuf)
            outs = 2 if epi in (capi.EPI_GELU_SAVE, capi.EPI_GELU_SAVE_DERIV) else 1
            if epi == capi.EPI_BIAS:
                base = timed(lambda i: torch.nn.functional.linear(xs[i], w, bb), nbuf)
            else:
                def two_pass(i):
                    h = torch.nn.functional.linear(xs[i], w, bb)
                    a = torch.empty_like(h)
                    capi.check(capi.lib().csb200_gelu_fwd(csbF._ptr(h), csbF._ptr(a), M, N, capi.BF16,
                                                          csbF._vp(capi.stream_of(h))), "gelu")
                    return a
                base = timed(two_pass, nbuf)
            nbytes = 2 * (M * K + N * K + outs * M * N)
            print(json.dumps({"shape": name, "M": M, "K": K, "N": N, "epilogue": label, "csb200_us": round(us, 2),
                              "cublas_path_us": round(base, 2), "speedup": round(base / us, 3),
                              "alg_gbs": round(nbytes / us / 1e3, 1), "hbm_frac": round(nbytes / us / 1e3 / peaks["hbm_gbs"], 3),
                              "tflops": round(2 * M * N * K / us / 1e6, 1)}), flush=True)
        if "fc1" in name:  # backward: grad_h = (grad_y W2) * GELU'(h) — one tcgen05 GEMM vs cuBLAS + flat GELU' pass
            hs = [torch.randn(M, N, device="cuda").to(torch.bfloat16) for _ in range(nbuf)]
            w2 = (torch.randn(K, N, device="cuda") / K ** 0.5).to(torch.bfloat16)
            lib = capi.lib()
            nws = lib.csb200_gelu_bwd_workspace_bytes(N)
            wsp = torch.empty(nws, dtype=torch.uint8, device="cuda")
            gb = torch.empty(N, dtype=torch.float32, device="cuda")

            def two_pass_bwd(i):
                da = torch.mm(xs[i], w2)
                dh = torch.empty_like(da)
                capi.check(lib.csb200_gelu_bwd(csbF._ptr(da), csbF._ptr(hs[i]), csbF._ptr(dh), csbF._ptr(gb), csbF._ptr(wsp),
                                               nws, M, N, capi.BF16, csbF._vp(capi.stream_of(da))), "gelu_bwd")
                return dh
            us = timed(lambda i: csbF._tc_dgelu(xs[i], w2, hs[i]), nbuf)
            base = timed(two_pass_bwd, nbuf)
            nbytes = 2 * (M * K + N * K + 2 * M * N)
            print(json.dumps({"shape": name, "M": M, "K": K, "N": N, "epilogue": "dgelu (backward)", "csb200_us": round(us, 2),
                              "cublas_path_us": round(base, 2), "speedup": round(base / us, 3),
                              "alg_gbs": round(nbytes / us / 1e3, 1), "hbm_frac": round(nbytes / us / 1e3 / peaks["hbm_gbs"], 3),
                              "tflops": round(2 * M * N * K / us / 1e6, 1)}), flush=True)
            # the pair that saved GELU'(h) in forward: the backward epilogue is one multiplication
            us = timed(lambda i: csbF._tc_dgelu(xs[i], w2, hs[i], deriv=True), nbuf)
            print(json.dumps({"shape": name, "M": M, "K": K, "N": N, "epilogue": "dact (backward, saved GELU')", "csb200_us": round(us, 2),
                              "cublas_path_us": round(base, 2), "speedup": round(base / us, 3),
                              "alg_gbs": round(nbytes / us / 1e3, 1), "hbm_frac": round(nbytes / us / 1e3 / peaks["hbm_gbs"], 3),
                              "tflops": round(2 * M * N * K / us / 1e6, 1)}), flush=True)
            del hs
        del xs


def wgrad_bench(only):
    """csb200_linear_wgrad (+ bias gradient) vs cuBLAS split-K wgrad + csb200 column-sum pass."""
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.isfile(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    shapes = {}
    for st, (tok, C) in {"s1": (524288, 64), "s2": (131072, 128), "s3": (32768, 256), "s4": (8192, 512)}.items():
        shapes.update({f"{st}.qkv": (tok, 3 * C, C), f"{st}.proj": (tok, C, C), f"{st}.fc1": (tok, 4 * C, C),
                       f"{st}.fc2": (tok, C, 4 * C)})
    for name, (M, N, K) in shapes.items():
        if only and only not in name:
            continue
        nbuf = max(2, int(300e6 // (2 * M * (K + N))) + 1)
        gs = [torch.randn(M, N, device="cuda").to(torch.bfloat16) for _ in range(nbuf)]
        xs = [torch.randn(M, K, device="cuda").to(torch.bfloat16) for _ in range(nbuf)]
        us = timed(lambda i: csbF._tc_wgrad(gs[i], xs[i], True), nbuf)

        def cublas(i):
            return torch.mm(gs[i].t(), xs[i], out_dtype=torch.float32), csbF._bias_grad(gs[i], N)
        base = timed(cublas, nbuf)
        base_nobias = timed(lambda i: torch.mm(gs[i].t(), xs[i], out_dtype=torch.float32), nbuf)
        nbytes = 2 * M * (N + K) + 4 * N * K
        print(json.dumps({"shape": name + ".wgrad", "M": M, "N": N, "K": K, "csb200_us": round(us, 2),
                          "cublas_plus_colsum_us": round(base, 2), "cublas_wgrad_only_us": round(base_nobias, 2),
                          "speedup": round(base / us, 3), "alg_gbs": round(nbytes / us / 1e3, 1),
                          "hbm_frac": round(nbytes / us / 1e3 / peaks["hbm_gbs"], 3),
                          "tflops": round(2 * M * N * K / us / 1e6, 1)}), flush=True)
        del gs, xs


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "wgrad":
        wgrad_bench(sys.argv[2] if len(sys.argv) > 2 else "")
        sys.exit(0)
    main()
