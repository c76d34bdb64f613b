#!/usr/bin/env python
"""Per-kernel GPU time of the 512^2 / batch-32 bf16 train step (torch.profiler / CUPTI, eager mode).
Complements the ncu launch list under profiles/: durations here are warm and overlapped as in a real
step.  Prints the top kernels by total device time over the profiled steps."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402
import cswin_simam_unet_b200 as pkg  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda")
torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True
torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
net = pkg.CSWinTransformer(img_size=512, split_size=[1, 2, 8, 8], simam=True).to(dev)
opt = pkg.FusedAdamW(net.parameters(), lr=1e-4, weight_decay=1e-4)
step = pkg.TrainStep(net, opt, precision="bf16")
x, y = pkg.synthetic_batch(32, 512, dev, seed=0)
for _ in range(3):
    step(x, y)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(steps):
        step(x, y)
    torch.cuda.synchronize()
rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
total = sum(e.device_time_total for e in rows)
print(f"total device time {total / steps / 1e3:.2f} ms/step over {steps} steps")
for e in rows[:70]:
    print(f"{e.device_time_total / steps / 1e3:8.3f} ms {100 * e.device_time_total / total:5.1f}%  n={e.count // steps:4d}  {e.key[:110]}")

if len(sys.argv) > 2 and sys.argv[2] == "ops":
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
        step(x, y)
        torch.cuda.synchronize()
    ops = sorted(prof.key_averages(group_by_input_shape=True), key=lambda e: -e.self_device_time_total)
    print("\n--- aten ops by self device time (one step), with input shapes ---")
    ops = [e for e in ops if not e.key.startswith("void ") and "::" in e.key or e.key.startswith("_")]
    for e in ops[:70]:
        print(f"{e.self_device_time_total / 1e3:8.3f} ms  n={e.count:4d}  {e.key[:40]:40s} {str(e.input_shapes)[:110]}")
