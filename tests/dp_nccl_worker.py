"""Worker of tests/test_dp_nccl_gpu.py (one process per GPU, NCCL): data-parallel gradients and captured
data-parallel training steps against the single-GPU global-batch run."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import cswin_simam_unet_b200 as pkg  # noqa: E402


def make_net(dev):
    torch.manual_seed(0)  # identical replicas (and the identical single-GPU model)
    return pkg.CSWinTransformer(img_size=64, split_size=[1, 2, 2, 2], simam=True).to(dev)


def rel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


def main():
    out_path = sys.argv[1]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    G = 4 * world  # global batch
    shard = pkg.shard_of_global_batch(G, rank, world)
    res = {}

    # ---- 1. averaged gradients of one backward == gradients of the global batch on one GPU ----
    net = make_net(dev)
    red = pkg.GradientAllReducer(net.parameters())
    step = pkg.TrainStep(net, torch.optim.SGD(net.parameters(), lr=0.0), precision="fp32", reducer=red)
    x, y = pkg.synthetic_batch(len(shard), 64, dev, seed=3, first_index=shard.start)
    red.begin_step()
    step.forward_loss(x, y).backward()
    red.finish_step()
    if rank == 0:
        ref = make_net(dev)
        gx, gy = pkg.synthetic_batch(G, 64, dev, seed=3)
        pkg.TrainStep(ref, torch.optim.SGD(ref.parameters(), lr=0.0), precision="fp32").forward_loss(gx, gy).backward()
        res["grad_rel"] = max(rel(p.grad, q.grad) for p, q in zip(net.parameters(), ref.parameters()))
        num = sum(float((p.grad.double() - q.grad.double()).pow(2).sum()) for p, q in zip(net.parameters(), ref.parameters()))
        den = sum(float(q.grad.double().pow(2).sum()) for q in ref.parameters())
        res["grad_rel_l2"] = (num / den) ** 0.5
    red.close()

    # ---- 1b. the same with a bf16 wire (reduce_dtype): half the bytes, the average rounded to bf16 ----
    net2 = make_net(dev)
    red = pkg.GradientAllReducer(net2.parameters(), reduce_dtype=torch.bfloat16)
    step = pkg.TrainStep(net2, torch.optim.SGD(net2.parameters(), lr=0.0), precision="fp32", reducer=red)
    red.begin_step()
    step.forward_loss(x, y).backward()
    red.finish_step()
    res["bf16_wire_bytes_ratio"] = red.gradient_bytes() / sum(p.numel() * 4 for p in net2.parameters())
    if rank == 0:
        num = sum(float((p.grad.double() - q.grad.double()).pow(2).sum()) for p, q in zip(net2.parameters(), ref.parameters()))
        res["bf16_wire_grad_rel_l2"] = (num / den) ** 0.5
    red.close()

    # ---- 2. three captured data-parallel AdamW steps (all-reduce inside the graph) == global-batch steps ----
    for capture in (True, False):
        net = make_net(dev)
        # captured: ONE bucket all-reduced after backward (what bench.py runs); eager: bucketed + overlapped
        red = pkg.GradientAllReducer(net.parameters(), bucket_bytes=1 << 30, overlap=False) if capture \
            else pkg.GradientAllReducer(net.parameters())
        opt = pkg.FusedAdamW(net.parameters(), lr=1e-3, weight_decay=1e-4)
        step = pkg.TrainStep(net, opt, precision="fp32", reducer=red, cuda_graph=True, capture_collectives=capture)
        for s in range(3):
            x, y = pkg.synthetic_batch(len(shard), 64, dev, seed=10 + s, first_index=shard.start)
            loss = step(x, y)
        torch.cuda.synchronize()
        key = "in_graph" if capture else "between_graphs"
        res[key + "_reduce_in_graph"] = bool(step._reduce_in_graph)
        # replicas identical?
        chk = torch.stack([p.detach().double().sum() for p in net.parameters()]).sum().reshape(1)
        allchk = [torch.empty_like(chk) for _ in range(world)]
        dist.all_gather(allchk, chk)
        res[key + "_replicas_identical"] = all(torch.equal(c, allchk[0]) for c in allchk)
        if rank == 0:
            ref = make_net(dev)
            ropt = pkg.FusedAdamW(ref.parameters(), lr=1e-3, weight_decay=1e-4)
            rstep = pkg.TrainStep(ref, ropt, precision="fp32")
            for s in range(3):
                gx, gy = pkg.synthetic_batch(G, 64, dev, seed=10 + s)
                rstep(gx, gy)
            num = sum(float((p.double() - q.double()).pow(2).sum()) for p, q in zip(net.parameters(), ref.parameters()))
            den = sum(float(q.double().pow(2).sum()) for q in ref.parameters())
            res[key + "_weights_rel_l2"] = (num / den) ** 0.5
        red.close()
    if rank == 0:
        with open(out_path, "w") as f:
            json.dump(res, f)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
