"""Phase timers of the grid-resident SimAM kernels (build with `make EXTRA=-DCSB_PROF`): cycles per round of
CTA 1 / thread 0, per phase of the main loop."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import cswin_simam_unet_b200 as pkg
lib = pkg.capi.lib()
names = ["early_acquire", "wait_load", "accumulate", "wait_published", "fetch_slots(issue)", "publish", "totals",
         "rescale", "sync+issue_load"]
for shape in [(32, 16384, 64), (32, 4096, 128), (32, 1024, 256)]:
    x = torch.randn(shape, device="cuda").bfloat16().requires_grad_(True)
    g = torch.randn(shape, device="cuda").bfloat16()
    for which in ("fwd", "bwd"):
        buf = (ctypes.c_ulonglong * 16)()
        for _ in range(3):
            y = pkg.simam(x, 1e-4, "NLC")
            if which == "bwd":
                y.backward(g)
        lib.csb200_debug_prof_simam(buf, 1)
        y = pkg.simam(x, 1e-4, "NLC")
        lib.csb200_debug_prof_simam(buf, 1)  # forward of this call only ... reset
        if which == "bwd":
            y.backward(g)
        else:
            y = pkg.simam(x, 1e-4, "NLC")
        lib.csb200_debug_prof_simam(buf, 1)
        v = list(buf)
        n = max(v[9], 1)
        print(shape, which, "iterations", v[9], "total cycles/iter %.0f" % (sum(v[:9]) / n))
        print("   " + "  ".join("%s %.0f" % (names[i], v[i] / n) for i in range(9)))
