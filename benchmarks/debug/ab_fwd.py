"""A/B of forward ring depths: CSB200_LIB=<path to a variant .so> python benchmarks/debug/ab_fwd.py"""
import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import cswin_simam_unet_b200 as pkg
if os.environ.get("CSB200_LIB"):
    pkg.capi.LIB_PATH = os.environ["CSB200_LIB"]
B = 32
for name, reso, split, heads, C in [("s1",128,1,2,64),("s2",64,2,4,128)]:
    blk = pkg.CSWinBlock(dim=C, reso=reso, num_heads=heads, split_size=split).cuda()
    L = reso*reso
    qs = [torch.randn(B, L, 3*C, device="cuda").bfloat16() for _ in range(4)]
    with torch.no_grad():
        for i in range(4): blk.attend(qs[i])
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for rep in range(5):
            a.record()
            for i in range(8): blk.attend(qs[i % 4])
            b.record(); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b) / 8)
    print(os.environ.get("CSB200_LIB", "default")[-16:], name, round(best * 1e3, 1), "us (incl. ~launch gaps)")
