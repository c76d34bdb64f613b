"""Phase timings (clock64) of the tcgen05 backward kernel; needs a build with EXTRA=-DCSB_PROF."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import cswin_simam_unet_b200 as pkg
lib = pkg.capi.lib()
B = 32
for name, reso, split, heads, C in [("s1",128,1,2,64),("s2",64,2,4,128),("s3",32,8,8,256),("s4",16,16,16,512)]:
    blk = pkg.CSWinBlock(dim=C, reso=reso, num_heads=heads, split_size=split, last_stage=(name=="s4")).cuda()
    L = reso*reso
    q = torch.randn(B, L, 3*C, device="cuda").bfloat16().requires_grad_(True)
    g = torch.randn(B, L, C, device="cuda").bfloat16()
    params = [p for a in blk.attns for p in (a.get_v.weight, a.get_v.bias)]
    for _ in range(2): torch.autograd.grad(blk.attend(q), [q] + params, g)
    buf = (ctypes.c_ulonglong*32)()
    lib.csb200_debug_prof_bwd(buf, 1)
    torch.autograd.grad(blk.attend(q), [q] + params, g)
    lib.csb200_debug_prof_bwd(buf, 1)
    v = list(buf)
    nc, ng = max(v[3],1), max(v[20],1)
    print(name, "kernel cycles", v[24], "convert iters (WG0)", v[3], "groups", v[20])
    print("  convert WG0 per iter: wait_ds %.0f  wait_sdp %.0f  compute %.0f" % (v[0]/nc, v[1]/nc, v[2]/nc))
    print("  MMA per iter(all): wait_conv %.0f wait_acc %.0f wait_grp %.0f wait_stage %.0f" % tuple(x/(2*nc) for x in v[8:12]))
    print("  MMA issue per iter: sdp %.0f  dependent %.0f" % (v[12]/(2*nc), v[13]/(2*nc)))
    print("  epilogue per group: wait_dvdk %.0f  dvdk_out %.0f  wait_dq %.0f  dq_out %.0f" % tuple(x/ng for x in v[16:20]))
