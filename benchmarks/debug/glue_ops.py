"""Which ATen glue ops (copies, adds, cats, casts) are left in the 512^2 train step, and who calls them?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from torch.profiler import ProfilerActivity, profile
import cswin_simam_unet_b200 as pkg

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
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=True) as prof:
    step(x, y)
    torch.cuda.synchronize()
ops = sorted(prof.key_averages(group_by_input_shape=True, group_by_stack_n=8), key=lambda e: -e.self_device_time_total)
want = ("aten::copy_", "aten::add", "aten::cat", "aten::_to_copy", "aten::sum", "aten::mul", "aten::clone", "aten::fill_",
        "aten::zero_", "aten::sigmoid", "aten::binary_cross", "aten::div", "aten::sub", "aten::neg", "aten::where")
tot = 0.0
ALL = len(sys.argv) > 1 and sys.argv[1] == "all"
for e in ops:
    if e.self_device_time_total < 5 or not (e.key.startswith(want) or (ALL and e.key.startswith("aten::"))):
        continue
    tot += e.self_device_time_total
    frames = [f for f in e.stack if "cswin" in f or "bench" in f][:4]
    print(f"{e.self_device_time_total / 1e3:7.3f} ms n={e.count:3d} {e.key[:24]:24s} {str(e.input_shapes)[:70]:70s} {' <- '.join(s.split('/')[-1][:60] for s in frames)}")
print(f"total glue {tot / 1e3:.3f} ms")
