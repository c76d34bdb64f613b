"""Where does the data-parallel step spend its extra time?  torchrun --nproc-per-node 2 benchmarks/debug/dp_profile.py
Profiles 3 replays of the captured DP train step (512^2, 32 images per GPU) and prints, for rank 0, the kernels around
the NCCL all-reduce with their start / end times relative to the step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.distributed as dist
from torch.profiler import ProfilerActivity, profile
import cswin_simam_unet_b200 as pkg

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True
torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
net = pkg.CSWinTransformer(img_size=512, split_size=[1, 2, 8, 8], simam=True).to(dev)
opt = pkg.FusedAdamW(net.parameters(), lr=1e-4, weight_decay=1e-4)
kw = {}
if len(sys.argv) > 1 and sys.argv[1] == "bf16":
    kw["reduce_dtype"] = torch.bfloat16
red = pkg.GradientAllReducer(net.parameters(), bucket_bytes=1 << 30, overlap=False, **kw) if world > 1 else None
step = pkg.TrainStep(net, opt, precision="bf16", reducer=red, cuda_graph=True)
x, y = pkg.synthetic_batch(32, 512, dev, seed=0, first_index=32 * rank)
for _ in range(4):
    step(x, y)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(10):
    step(x, y)
t1.record()
torch.cuda.synchronize()
if rank == 0:
    print(f"world {world}: {t0.elapsed_time(t1) / 10:.3f} ms/step", flush=True)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step(x, y)
    torch.cuda.synchronize()
if rank == 0:
    ev = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start)
    idx = [i for i, e in enumerate(ev) if "nccl" in e.name.lower()]
    print("nccl kernels:", [(ev[i].name[:50], round(ev[i].time_range.elapsed_us(), 1)) for i in idx])
    if idx:
        i0 = idx[-1]
        base = ev[max(0, i0 - 14)].time_range.start
        for e in ev[max(0, i0 - 14): i0 + 8]:
            print(f"{(e.time_range.start - base):9.1f} -> {(e.time_range.end - base):9.1f} us  {e.name[:90]}")
# live CUDA graphs hold NCCL work: tearing the process group down here hangs (it cost two 300-s timeouts on a
# 2-GPU box) — leave the way bench.py does
sys.stdout.flush()
os._exit(0)
