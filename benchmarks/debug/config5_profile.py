import os, sys
sys.path.insert(0, "/root/repo")
import torch
from torch.profiler import ProfilerActivity, profile
import cswin_simam_unet_b200 as pkg
torch.backends.cudnn.benchmark = True
for sw in (1, 8):
    net = pkg.CSWinTransformer(img_size=1024, split_size=[sw] * 4, simam=True).cuda().eval()
    x = torch.rand(8, 3, 1024, 1024, device="cuda")
    def f():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            net(x)
    for _ in range(3): f()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        f(); torch.cuda.synchronize()
    rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
    tot = sum(e.device_time_total for e in rows)
    print(f"split {sw}: total device time {tot/1e3:.2f} ms")
    for e in rows[:14]:
        print(f"  {e.device_time_total/1e3:7.3f} ms n={e.count:3d} {e.key[:100]}")
    del net, x
