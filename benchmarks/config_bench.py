#!/usr/bin/env python
"""The other BASELINE configs (parity cases in tests/, timed here once; one JSON line each).

    python benchmarks/config_bench.py 2      # UNet + SimAM, 256^2, batch 16, bf16 train step (U:342-348, Adam U:486)
    python benchmarks/config_bench.py 5      # CSWin-SimAM-UNet 1024^2 inference, batch 8, stripe-width sweep
    python benchmarks/config_bench.py simam  # SimAM roofline on the config-2 / config-3 call sites (csb200 C ABI)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import cswin_simam_unet_b200 as pkg  # noqa: E402
from cswin_simam_unet_b200 import functional as csbF  # noqa: E402

PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))


def timed_ms(fn, steps, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


def config2():
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(0)
    net = pkg.UNet(simam=True).cuda().to(memory_format=torch.channels_last)
    x, y = pkg.synthetic_batch(16, 256, "cuda", seed=0)
    which = sys.argv[2] if len(sys.argv) > 2 else "csb200"
    if which == "csb200":  # U:486-490 Adam(lr 1e-3, weight_decay 1e-4) as one csb200_adam_step launch
        opt = pkg.fused_adam(net.parameters(), lr=1e-3, weight_decay=1e-4)
    else:
        opt = torch.optim.Adam(net.parameters(), lr=1e-3, weight_decay=1e-4, capturable=True)
    step = pkg.TrainStep(net, opt, precision="bf16", cuda_graph=True)
    ms = timed_ms(lambda: step(x, y), 20)
    eager = pkg.TrainStep(net, opt, precision="bf16")
    t = csbF.KernelTimer()
    for _ in range(2):
        eager(x, y)
    csbF.set_kernel_timer(t)
    for _ in range(5):
        torch.cuda._sleep(int(0.02 * 1.9e9))
        eager(x, y)
    csbF.set_kernel_timer(None)
    fams = {k: {"GBps": round(v["bytes"] / v["ms"] / 1e6, 1), "frac_of_hbm": round(v["bytes"] / v["ms"] / 1e6 / PEAKS["hbm_gbs"], 3),
                "ms_per_step": round(v["ms"] / 5, 3), "calls_per_step": v["calls"] // 5} for k, v in t.summary().items()}
    print(json.dumps({"config": "2: UNet+SimAM 256^2 batch 16 bf16 train step (CUDA graph)", "optimizer": which, "ms_per_step": round(ms, 3),
                      "img_per_s": round(16 / ms * 1e3, 1), "csb200_families": fams}), flush=True)


def config5():
    torch.backends.cudnn.benchmark = True
    for size, sw in ((1024, 1), (1024, 2), (1024, 8), (896, 7)):
        torch.manual_seed(0)
        net = pkg.CSWinTransformer(img_size=size, split_size=[sw] * 4, simam=True).cuda().eval()
        x = torch.rand(8, 3, size, size, device="cuda")

        def f():
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                net(x)
        ms = timed_ms(f, 10)
        # the same forward replayed from a CUDA graph (pkg.InferStep): eager issue of the ~1 000 launches costs more
        # than the kernels take at split 1 / 2 (6.8 ms of device time against 13.7 ms per eager batch)
        infer = pkg.InferStep(net, precision="bf16", cuda_graph=True)
        infer(x)
        ms_graph = timed_ms(lambda: infer(x), 10)
        engines = sorted({a.engine for m in net.modules() if isinstance(m, pkg.CSWinBlock) for a in m.attns})
        print(json.dumps({"config": f"5: CSWin-SimAM-UNet {size}^2 inference batch 8 bf16, split {[sw] * 4}",
                          "ms_per_batch": round(ms_graph, 2), "img_per_s": round(8 / ms_graph * 1e3, 1),
                          "eager_ms_per_batch": round(ms, 2), "eager_img_per_s": round(8 / ms * 1e3, 1),
                          "mode": "pkg.InferStep: forward replayed from a CUDA graph (eager numbers beside it)",
                          "engine_setting": engines}), flush=True)
        del infer
        del net, x
        torch.cuda.empty_cache()


def simam():
    import subprocess
    subprocess.run([sys.executable, os.path.join(ROOT, "benchmarks", "kernel_bench.py"), "simam", "--dtype", "bfloat16"], check=True)


if __name__ == "__main__":
    {"2": config2, "5": config5, "simam": simam}[sys.argv[1]]()
