import ctypes, sys, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import cswin_simam_unet_b200 as pkg
lib = pkg.capi.lib()
B=32
for name, reso, split, heads, C in [("s1",128,1,2,64),("s2",64,2,4,128),("s3",32,8,8,256),("s4",16,16,16,512)]:
    blk = pkg.CSWinBlock(dim=C, reso=reso, num_heads=heads, split_size=split, last_stage=(name=="s4")).cuda()
    L = reso*reso
    q = torch.randn(B, L, 3*C, device="cuda").bfloat16()
    with torch.no_grad():
        for _ in range(3): blk.attend(q)
    buf = (ctypes.c_ulonglong*32)()
    lib.csb200_debug_prof_fwd(buf, 1)
    with torch.no_grad(): blk.attend(q)
    lib.csb200_debug_prof_fwd(buf, 1)
    v = list(buf)
    tiles = v[5]/ (3 if name in ("s1","s2") else 2)  # counted once per WG-lane0 per tile -> tiles total = v[5]
    n = max(v[5],1)
    print(name, "tiles(CTA0)", v[5], "kernel cycles", v[11])
    print("  per tile per WG: wait_S %.0f  max %.0f  exp %.0f  wait_O %.0f  epilogue %.0f" % tuple(v[i]/n for i in range(5)))
    print("  epilogue split: tmem+kvwait %.0f  stencil %.0f  stores %.0f" % (v[12]/n, v[13]/n, v[14]/n))
    print("  MMA warp per tile: wait_P %.0f  wait_QKV %.0f  wait_buf %.0f" % (v[8]/n, v[9]/n, v[10]/n))
