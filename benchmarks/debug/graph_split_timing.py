"""Where does the captured train step spend its time: graph 1 (forward + backward) vs graph 2 (optimizer)?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import cswin_simam_unet_b200 as pkg

which = sys.argv[1] if len(sys.argv) > 1 else "csb200"
dev = torch.device("cuda")
torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True
torch.backends.cudnn.benchmark = True
torch.manual_seed(0)
net = pkg.CSWinTransformer(img_size=512, split_size=[1, 2, 8, 8], simam=True).to(dev)
if which == "csb200":
    opt = pkg.FusedAdamW(net.parameters(), lr=1e-4, weight_decay=1e-4)
else:
    opt = torch.optim.AdamW(net.parameters(), lr=1e-4, weight_decay=1e-4, fused=True, capturable=True)
step = pkg.TrainStep(net, opt, precision="bf16", cuda_graph=True)
x, y = pkg.synthetic_batch(32, 512, dev, seed=0)
for _ in range(3):
    step(x, y)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
t1 = t2 = 0.0
N = 20
for _ in range(N):
    ev[0].record()
    step._graph.replay()
    ev[1].record()
    if step._graph_opt is not None:
        step._graph_opt.replay()
    ev[2].record()
    torch.cuda.synchronize()
    t1 += ev[0].elapsed_time(ev[1])
    t2 += ev[1].elapsed_time(ev[2])
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(N):
    step(x, y)
b.record()
torch.cuda.synchronize()
print(f"{which}: graph1 {t1 / N:.3f} ms, graph2 {t2 / N:.3f} ms, back-to-back step {a.elapsed_time(b) / N:.3f} ms")
