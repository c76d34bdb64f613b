#!/usr/bin/env python
"""Headline benchmark: CSWin-SimAM-UNet training throughput (train img/s) at 512^2, bf16.

    python bench.py --gpus N --steps K --warmup W            # this repo, sm_100a kernels
    python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on host cores

N > 1 is launched by torchrun (one rank per GPU, NCCL); weak scaling with 32 images per GPU (BASELINE
config 3 at N=1, config 4's global batch 256 at N=8).  Prints ONE JSON line on rank 0.

step    = zero_grad -> forward (autocast bf16) -> BCE (fp32) -> backward -> [grad all-reduce] -> AdamW
          (C:780-786, lr 1e-4, wd 1e-4 as C:937-941), dropout 0 (constructor defaults)
value   = images/s with the batch already resident in HBM, K steps between CUDA events, max over ranks
e2e     = same step through the public API with pinned HOST batches: H2D copy of images+masks (started
          with TrainStep.prefetch while the previous step computes) and a D2H read of the loss inside the
          timed region, every step
roofline= the csb200 kernel family that takes the most time inside the timed steps, algorithmic
          FLOPs (attention) or bytes (SimAM) over its CUDA-event time, against MEASURED_PEAKS.json
cpu_baseline / --impl reference = the UNMODIFIED reference (baseline/_ref, installed by __graft_entry__.build();
          kind "reference") on the host cores, fp32, all host threads, bounded sample (B=2 steps); the
          oracle's CPU port (kind "port") only when that install is absent
gpu_eager_baseline = the unmodified reference in stock PyTorch eager on the SAME B200 (fp32 TF32-off / TF32-on /
          bf16 autocast), same batch and geometry: the bar BASELINE.md section 5 names (N = 1 only)
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

IMG, SPLIT, BATCH_PER_GPU = 512, [1, 2, 8, 8], 32
METRIC, UNIT = "train img/s at 512^2 bf16 (CSWin-SimAM-UNet)", "img/s"
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            return json.load(f), "measured (MEASURED_PEAKS.json)"
    return dict(FALLBACK_PEAKS), "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# clocks during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.samples, self.stop_flag, self.thread = index, [], threading.Event(), None

    def _run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [v.strip() for v in out.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append((float(f[0]), float(f[1]), f[2:6]))
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def __enter__(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()
        return self

    def __exit__(self, *exc):
        self.stop_flag.set()
        self.thread.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for _, _, fl in self.samples for n, v in zip(names, fl) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(s[0] for s in self.samples), "sm_max_mhz": self.samples[0][1],
                "reasons": reasons, "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# the reference algorithm on host cores (oracle port) — cpu_baseline and --impl reference
# ------------------------------------------------------------------------------------------------
def _reference_module():
    """The UNMODIFIED reference script (train_cswinunet_segmentation.py) as installed under baseline/_ref by
    __graft_entry__.build(), or None when that install is absent (then the oracle port stands in)."""
    try:
        from oracle import reference_shim
        if reference_shim.available():
            return reference_shim.load("cswin")
    except Exception as e:  # a missing third-party import on the box: report the port instead
        sys.stderr.write(f"reference scripts not loadable ({e}); using the oracle port\n")
    return None


def cpu_reference_steps(steps: int, warmup: int, batch: int = 2):
    """One train step of the reference at 512^2 on the host cores, fp32, bounded sample (batch 2):
    zero_grad -> forward -> BCELoss -> backward -> AdamW (C:780-786, C:936-941).  kind "reference": the
    reference's own CSWinTransformer (no SimAM: the checkout has none); kind "port": oracle/models.py with
    SimAM on the skips, when the reference scripts are not installed."""
    from cswin_simam_unet_b200.train import synthetic_batch
    torch.set_num_threads(os.cpu_count())
    x, y = synthetic_batch(batch, IMG, "cpu", seed=0)
    ref = _reference_module()
    if ref is not None:
        torch.manual_seed(0)
        net = ref.CSWinTransformer(img_size=IMG, split_size=SPLIT)  # ctor defaults: dropout 0 (C:495-496)
        net.train()
        opt = torch.optim.AdamW(net.parameters(), lr=1e-4, weight_decay=1e-4)  # C:937-941
        crit = torch.nn.BCELoss()  # C:936
        kind, what = "reference", "unmodified reference CSWinTransformer (baseline/_ref, C:489-688, no SimAM)"

        def one_step():
            opt.zero_grad()
            loss = crit(net(x), y)
            loss.backward()
            opt.step()
    else:
        from oracle import models as om
        cfg = om.CSWinConfig(img_size=IMG, split_size=SPLIT, simam=True)
        params = {k: v.requires_grad_(True) for k, v in om.synth_params(om.cswin_param_shapes(cfg), 0).items()}
        opt = torch.optim.AdamW(list(params.values()), lr=1e-4, weight_decay=1e-4)
        kind, what = "port", "oracle/models.py (CPU port of C:489-688 + SimAM on skips)"

        def one_step():
            opt.zero_grad()
            loss = torch.nn.functional.binary_cross_entropy(om.cswin_unet_forward(params, x, cfg), y)
            loss.backward()
            opt.step()
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        one_step()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    return {"value": batch * steps / total, "unit": UNIT, "cores": torch.get_num_threads(), "kind": kind,
            "sample": f"{steps} steps of batch {batch} at {IMG}^2 fp32 after {warmup} warm-up, {what}, AdamW",
            "ms_per_step": 1e3 * total / steps}


def gpu_eager_reference(dev, batch: int, steps: int = 3, warmup: int = 2):
    """BASELINE.md section 5 / SURVEY.md 8(d) "end-to-end": the UNMODIFIED reference in stock PyTorch eager on the
    same B200, same step (C:780-786), same batch and geometry: fp32 with TF32 off (the parity anchor), fp32
    with TF32 on, and bf16 autocast with the BCE evaluated in fp32 (CUDA autocast refuses BCELoss).  Returns
    None when baseline/_ref is not installed.  Halves the batch on out-of-memory and says so."""
    ref = _reference_module()
    if ref is None:
        return None
    from cswin_simam_unet_b200.train import synthetic_batch
    out = {"model": "reference CSWinTransformer(img_size=512, split_size=[1,2,8,8]), stock PyTorch eager, no SimAM",
           "steps": steps, "warmup": warmup, "unit": UNIT}
    saved = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    for mode in ("fp32_tf32_off", "fp32_tf32_on", "bf16_autocast"):
        b = batch
        while b >= 1:
            try:
                torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = mode == "fp32_tf32_on"
                torch.manual_seed(0)
                net = ref.CSWinTransformer(img_size=IMG, split_size=SPLIT).to(dev).train()
                opt = torch.optim.AdamW(net.parameters(), lr=1e-4, weight_decay=1e-4)
                crit = torch.nn.BCELoss()
                x, y = synthetic_batch(b, IMG, dev, seed=0)
                t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                for it in range(warmup + steps):
                    if it == warmup:
                        torch.cuda.synchronize(dev)
                        t0.record()
                    opt.zero_grad()
                    if mode == "bf16_autocast":
                        with torch.autocast("cuda", dtype=torch.bfloat16):
                            probs = net(x)
                        loss = crit(probs.float(), y)
                    else:
                        loss = crit(net(x), y)
                    loss.backward()
                    opt.step()
                t1.record()
                torch.cuda.synchronize(dev)
                ms = t0.elapsed_time(t1) / steps
                out[mode] = {"value": b * 1e3 / ms, "ms_per_step": ms, "batch": b,
                             "peak_mem_gib": round(torch.cuda.max_memory_allocated(dev) / 2 ** 30, 1)}
                break
            except torch.OutOfMemoryError:
                b //= 2
            finally:
                net = opt = x = y = loss = probs = None
                torch.cuda.empty_cache()
                torch.cuda.reset_peak_memory_stats(dev)
        else:
            out[mode] = {"value": None, "error": "out of memory at batch 1"}
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = saved
    return out


def run_reference(args, rank):
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 5)), max(1, min(args.warmup, 1))
    r = cpu_reference_steps(steps, warmup)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"CSWin-SimAM-UNet train step {IMG}x{IMG}, split {SPLIT}, CPU sample batch 2",
                       "reference": r["kind"]},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------
def run_ours(args, rank, local_rank, world):
    import cswin_simam_unet_b200 as pkg
    from cswin_simam_unet_b200 import functional as csbF

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cudnn.benchmark = True

    B = args.batch_per_gpu
    torch.manual_seed(0)  # identical replicas on every rank
    net = pkg.CSWinTransformer(img_size=IMG, split_size=SPLIT, simam=True, attn_engine=args.attn_engine).to(dev)
    use_graph = args.cuda_graph
    if args.optimizer == "csb200":  # C:937-941 AdamW(lr 1e-4, wd 1e-4) as one csb200_adam_step launch
        opt = pkg.FusedAdamW(net.parameters(), lr=1e-4, weight_decay=1e-4)
    else:
        opt = torch.optim.AdamW(net.parameters(), lr=1e-4, weight_decay=1e-4, fused=True, capturable=use_graph)
    # Data parallel: ONE flat bucket (94 MB fp32), all-reduced after backward INSIDE the captured graph.  Measured on
    # 2 GPUs (profiles/r2_dp_overlap_experiment.txt): overlapping the all-reduce with backward costs more than it
    # hides here — NCCL's CTAs take SMs away from the persistent one-CTA-per-SM kernels, whose last CTAs then run as
    # a second wave (step +1.08 ms) — while one exposed all-reduce on NVLink / NVSwitch is ~0.3 ms.
    reducer = pkg.GradientAllReducer(net.parameters(), bucket_bytes=1 << 30, overlap=False) if world > 1 else None
    step = pkg.TrainStep(net, opt, precision="bf16", reducer=reducer, cuda_graph=use_graph)
    # the roofline needs CUDA events around individual kernels, which a graph replay cannot give:
    # an eager twin of the step (same model, same optimizer) is timed in a separate pass
    eager = pkg.TrainStep(net, opt, precision="bf16", reducer=reducer) if use_graph else step

    shard = pkg.shard_of_global_batch(B * world, rank, world)
    host = [pkg.synthetic_batch(B, IMG, "cpu", seed=s, first_index=shard.start, pin=True) for s in range(2)]
    resident = [(x.to(dev), y.to(dev)) for x, y in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        h0 = time.perf_counter()
        for i in range(steps):
            fn(i)
        timed.host_ms = (time.perf_counter() - h0) * 1e3 / steps  # CPU time to ISSUE one step
        t1.record()
        barrier()
        ms = torch.tensor([t0.elapsed_time(t1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    def step_resident(i):
        x, y = resident[i % 2]
        step(x, y)

    def step_e2e(i):
        # pinned host batches: every step's inputs cross PCIe inside the timed region.  As in a training
        # loop with a data loader, the copy of batch i+1 is started (TrainStep.prefetch, side stream)
        # before the host blocks on the loss of step i, so the copy engine runs under the SMs.
        loss = step()                       # consumes the batch staged by the previous prefetch
        step.prefetch(*host[(i + 1) % 2])
        return loss.item()  # the D2H read of the step's result

    for i in range(args.warmup):
        step_resident(i)
    # --- throughput with the batch resident in HBM ----------------------------------------------
    timer = csbF.KernelTimer()
    if not use_graph:
        csbF.set_kernel_timer(timer)  # eager: per-family CUDA-event spans inside the timed region
    n0 = pkg.capi.launch_count()
    if os.environ.get("CSB200_PROFILE_RANGE"):  # ncu --profile-from-start off: only the timed steps
        torch.cuda.profiler.start()
    with ClockSampler(local_rank) as clk:
        ms = timed(step_resident, args.steps)
    if os.environ.get("CSB200_PROFILE_RANGE"):
        torch.cuda.profiler.stop()
    launches = pkg.capi.launch_count() - n0
    host_issue_ms = timed.host_ms
    csbF.set_kernel_timer(None)
    # (read now: `step` is released further down, before the other legs allocate)
    reduce_in_graph, defer_sums = bool(getattr(step, "_reduce_in_graph", False)), bool(step.defer_sums)
    roofline_pass = "CUDA events inside the timed steps"
    if use_graph:
        # a replay launches the csb200 kernels recorded at capture: count them in one eager step and
        # time the kernel families there (same shapes, same stream)
        for i in range(2):
            eager(*resident[i % 2])
        n0 = pkg.capi.launch_count()
        csbF.set_kernel_timer(timer)
        for i in range(args.steps):
            # The eager step is host-bound (~50 ms to issue, ~28 ms to run): park the GPU on a spin
            # kernel while the host queues the step, so the kernels then run back to back and the
            # event spans measure kernels, not launch gaps.
            torch.cuda._sleep(int(0.05 * 1.9e9))
            eager(*resident[i % 2])
        csbF.set_kernel_timer(None)
        torch.cuda.synchronize()
        launches = pkg.capi.launch_count() - n0
        roofline_pass = f"separate eager pass of {args.steps} steps (the timed region replays a CUDA graph)"
    fams = timer.summary()
    # --- end to end through the public API, host buffers ---------------------------------------
    step.prefetch(*host[0])
    step_e2e(0)
    ms_e2e = timed(step_e2e, args.steps)
    mem_gb = torch.cuda.max_memory_allocated(dev) / 2 ** 30

    # data parallel: every rank must hold bit-identical weights after all those steps (same averaged
    # gradients, same update) — a step that skipped its all-reduce would show up here, not in img/s
    in_sync = None
    if world > 1:
        with torch.no_grad():
            chk = torch.stack([p.detach().double().sum() for p in net.parameters()]).sum().reshape(1)
        allchk = [torch.empty_like(chk) for _ in range(world)]
        dist.all_gather(allchk, chk)
        in_sync = all(torch.equal(c, allchk[0]) for c in allchk)
        vals = [float(c) for c in allchk]
        spread = (max(vals) - min(vals)) / max(abs(vals[0]), 1e-30)
        if spread > 1e-7:  # a skipped all-reduce drifts by 1e-4 and more; identical updates give exactly 0
            raise RuntimeError(f"replicas diverged: parameter checksums {vals}")

    # BASELINE config 4 is a GLOBAL batch of 256: at N = 8 that is the weak-scaling point above (32 per GPU); at
    # N = 2 / 4 it is 128 / 64 images per GPU — measured here as an extra key (strong-scaling points)
    config4 = None
    if world in (2, 4) and not args.no_config4:
        try:
            b4 = 256 // world
            step = eager = None
            torch.cuda.empty_cache()
            opt4 = pkg.FusedAdamW(net.parameters(), lr=1e-4, weight_decay=1e-4)
            red4 = pkg.GradientAllReducer(net.parameters(), bucket_bytes=1 << 30, overlap=False)
            step4 = pkg.TrainStep(net, opt4, precision="bf16", reducer=red4, cuda_graph=use_graph)
            sh4 = pkg.shard_of_global_batch(256, rank, world)
            x4, y4 = pkg.synthetic_batch(b4, IMG, dev, seed=0, first_index=sh4.start)
            for _ in range(3):
                step4(x4, y4)
            n4 = max(5, args.steps // 2)
            ms4 = timed(lambda i: step4(x4, y4), n4)
            config4 = {"global_batch": 256, "batch_per_gpu": b4, "steps": n4, "ms_per_step": ms4 / n4,
                       "value": 256 * n4 / (ms4 / 1e3), "unit": UNIT, "scaling": "strong"}
        except Exception as e:  # never lose the headline line to the extra measurement
            config4 = {"error": str(e)[:200]}
    if rank != 0:
        finish_distributed(world)
        return
    pk, pk_src = peaks()
    roof_all = {}
    for fam, r in fams.items():
        sec = r["ms"] / 1e3
        if fam.startswith("attn"):
            ach, peak, unit, bound = r["flops"] / sec / 1e12, pk["bf16_tflops_sustained"], "TFLOP/s", "tensor"
        else:
            ach, peak, unit, bound = r["bytes"] / sec / 1e9, pk["hbm_gbs"], "GB/s", "hbm"
        roof_all[fam] = {"bound": bound, "achieved": round(ach, 2), "peak": peak, "unit": unit,
                         "frac": round(ach / peak, 4), "calls": r["calls"], "ms_total": round(r["ms"], 3),
                         "hbm_gbs": round(r["bytes"] / sec / 1e9, 1)}
    # The roofline object is quoted for ONE launch configuration: the (kernel family, shape) with the
    # largest share of the step; per-launch algorithmic work over the average launch duration.
    shapes = timer.summary(by_shape=True)
    top = max(shapes, key=lambda f: shapes[f]["ms"])
    r = shapes[top]
    fam = top.split("|")[0]
    sec_per_launch = r["ms"] / 1e3 / r["calls"]
    if fam.startswith("attn"):
        work, peak, unit, bound = r["flops"] / r["calls"], pk["bf16_tflops_sustained"], "TFLOP/s", "tensor"
        ach = work / sec_per_launch / 1e12
    else:
        work, peak, unit, bound = r["bytes"] / r["calls"], pk["hbm_gbs"], "GB/s", "hbm"
        ach = work / sec_per_launch / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get(top, {}).get("dram_bytes_per_launch")
    roofline = {"bound": bound, "achieved": round(ach, 2), "peak": peak, "unit": unit, "frac": round(ach / peak, 4),
                "traffic": traffic, "kernel": top, "launches": r["calls"],
                "us_per_launch": round(sec_per_launch * 1e6, 2),
                "algorithmic_flops_per_launch" if bound == "tensor" else "algorithmic_bytes_per_launch": work,
                "algorithmic_bytes_per_launch": r["bytes"] / r["calls"],
                "hbm_gbs": round(r["bytes"] / r["calls"] / sec_per_launch / 1e9, 1),
                "attainable_frac_of_tensor_peak_at_this_intensity":
                    (round(min(1.0, (r["flops"] / max(r["bytes"], 1)) / (peak * 1e12 / (pk["hbm_gbs"] * 1e9))), 3)
                     if bound == "tensor" else None),
                "peak_source": pk_src, "timing": roofline_pass,
                "share_of_step": round(r["ms"] / ms, 4),
                "traffic_source": "profiles/traffic.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum "
                                  "of every kernel the C-ABI call launches)" if traffic is not None else None}
    # every (kernel family, shape) of the step, largest first: average launch duration and the fraction of the
    # roofline that bounds it (tensor for attention, HBM for everything else)
    shape_rows = []
    for key, v in sorted(shapes.items(), key=lambda kv: -kv[1]["ms"])[:24]:
        us = v["ms"] * 1e3 / v["calls"]
        row = {"kernel": key, "calls_per_step": v["calls"] // args.steps, "us_per_launch": round(us, 2),
               "hbm_frac": round(v["bytes"] / v["calls"] / us / 1e3 / pk["hbm_gbs"], 3)}
        if v["flops"]:
            row["tensor_frac"] = round(v["flops"] / v["calls"] / us / 1e6 / pk["bf16_tflops_sustained"], 3)
        shape_rows.append(row)
    h2d = sum(t.numel() * t.element_size() for t in host[0])
    cpu = cpu_reference_steps(steps=3, warmup=1) if (world == 1 and not args.no_cpu_baseline) else None
    gpu_ref = None
    if world == 1 and not args.no_gpu_baseline:
        step = eager = net = opt = resident = None  # release the graph pools before the reference allocates
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats(dev)
        gpu_ref = gpu_eager_reference(dev, B)
    line = {
        "metric": METRIC, "value": B * world * args.steps / (ms / 1e3), "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"CSWin-SimAM-UNet train step {IMG}x{IMG} (BASELINE configs[2]; configs[3] at N=8)",
                   "global_batch": B * world, "batch_per_gpu": B, "split_size": SPLIT, "simam": "3 skip tensors (NLC); parity UNPINNED by the reference (its checkout has no SimAM code): checked against the public simam_module restated in oracle/ops.py",
                   "precision": "autocast bf16, fp32 master weights, fp32 sigmoid+BCE", "optimizer": "AdamW, one csb200_adam_step launch" if args.optimizer == "csb200" else "AdamW (torch fused)",
                   "dropout": 0.0, "parallelism": f"dp{world}", "attn_engine": args.attn_engine,
                   "cuda_graph": bool(use_graph),
                   "backward_final_sums": ("deferred: one csb200_sum_rows_flush launch per backward pass, weight-gradient "
                                           "outputs from one zeroed arena" if defer_sums else "immediate"),
                   "dp_allreduce": (None if world == 1 else
                                    ("NCCL AVG, one fp32 bucket, captured in the graph after backward"
                                     if reduce_in_graph else
                                     "NCCL AVG, fp32 buckets, issued eagerly between the two graphs")),
                   "l2": "no explicit flush: one step streams several GB of activations (>> 126 MB L2)",
                   "peak_mem_gib": round(mem_gb, 2)},
        "e2e": {"value": B * world * args.steps / (ms_e2e / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches), "host_issue_ms_per_step": round(host_issue_ms, 2),
        "replicas_in_sync": in_sync,
        "clocks": clk.summary(),
        "roofline": roofline,
        "roofline_all": roof_all,
        "roofline_shapes": shape_rows,
    }
    if cpu is not None:
        line["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
    if gpu_ref is not None:
        line["gpu_eager_baseline"] = gpu_ref
    if config4 is not None:
        line["config4_global_batch_256"] = config4
    print(json.dumps(line), flush=True)
    finish_distributed(world)


def finish_distributed(world):
    """End of a rank.  CUDA graphs that captured NCCL kernels are still alive here and tearing the communicator
    down under them can block (observed: destroy_process_group() never returned after the JSON line was printed),
    so after a last barrier every rank leaves without running destructors."""
    if world <= 1:
        return
    try:
        dist.barrier()
        torch.cuda.synchronize()
    finally:
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch-per-gpu", type=int, default=BATCH_PER_GPU)
    ap.add_argument("--attn-engine", default="auto", choices=["auto", "simt", "tcgen05"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-config4", action="store_true", help="skip the global-batch-256 point at N = 2 / 4")
    ap.add_argument("--no-gpu-baseline", action="store_true",
                    help="skip the reference-in-stock-PyTorch-eager leg on the same GPU (gpu_eager_baseline)")
    ap.add_argument("--optimizer", default="csb200", choices=["csb200", "torch"])
    ap.add_argument("--cuda-graph", dest="cuda_graph", action="store_true", default=True)
    ap.add_argument("--no-cuda-graph", dest="cuda_graph", action="store_false")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch N>1 with torch.distributed.run")
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
