"""Weight-gradient GEMMs of stage 3 (K = 32768 tokens): fp32-output (what functional._wgrad uses) vs bf16 output + cast."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
torch.backends.cuda.matmul.allow_tf32 = True


def t(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3


for (M, N, K) in [(1024, 256, 32768), (256, 1024, 32768), (768, 256, 32768), (256, 256, 32768), (512, 128, 131072),
                  (256, 64, 524288)]:
    g = torch.randn(K, M, device="cuda").bfloat16()
    x = torch.randn(K, N, device="cuda").bfloat16()
    f32 = t(lambda: torch.mm(g.t(), x, out_dtype=torch.float32))
    b16 = t(lambda: torch.mm(g.t(), x))
    b16c = t(lambda: torch.mm(g.t(), x).float())
    xt = t(lambda: torch.mm(x.t(), g, out_dtype=torch.float32))  # transposed result
    fl = 2.0 * M * N * K
    print(f"M{M} N{N} K{K}: fp32-out {f32:.1f} us ({fl / f32 / 1e6:.0f} TF/s)  bf16-out {b16:.1f}  bf16+cast {b16c:.1f}  "
          f"x^T g fp32-out {xt:.1f}")
