"""GPU: block / model / train-step parity against the reference's golden vectors (fp32) and the
stated bf16 tolerances (north_star: max abs error <= 2e-2 on logits, mask agreement >= 99.9 %)."""
import numpy as np
import pytest
import torch

from conftest import golden, rel_err
import cswin_simam_unet_b200 as pkg
from oracle import models as om

pytestmark = pytest.mark.gpu


def _load(net, seed, cfg=None):
    shapes = om.cswin_param_shapes(cfg) if cfg is not None else {k: tuple(v.shape) for k, v in net.state_dict().items()}
    net.load_state_dict(om.synth_params(shapes, seed))
    return net.cuda()


def test_block_golden_fp32(no_tf32):
    g = golden("block_dim64_reso8.npz")
    blk = pkg.CSWinBlock(dim=64, reso=8, num_heads=2, split_size=2, qkv_bias=True)
    shapes = {k: tuple(v.shape) for k, v in blk.state_dict().items()}
    blk.load_state_dict(om.synth_params(shapes, seed=3))
    blk.cuda()
    x = torch.tensor(g["x"], dtype=torch.float32, device="cuda", requires_grad=True)
    y = blk(x)
    y.backward(torch.tensor(g["gout"], dtype=torch.float32, device="cuda"))
    assert rel_err(y.cpu(), g["y"]) < 1e-5
    assert rel_err(x.grad.cpu(), g["dx"]) < 1e-5
    for k, p in blk.named_parameters():
        assert rel_err(p.grad.cpu(), g["grad." + k]) < 2e-5, k


@pytest.mark.parametrize("fname", ["cswin_64.npz", "cswin_224_config1.npz"])
def test_cswin_model_golden_fp32(no_tf32, fname):
    """BASELINE config 1 (224^2, B=2, fp32, split [1,2,7,7]) end to end on the CUDA kernels."""
    g = golden(fname)
    img, batch, seed = [int(v) for v in g["meta"][:3]]
    split = [int(v) for v in g["meta"][3:]]
    net = _load(pkg.CSWinTransformer(img_size=img, split_size=split), seed, om.CSWinConfig(img_size=img, split_size=split))
    x, y = torch.tensor(g["x"]).cuda(), torch.tensor(g["y"]).cuda()
    step = pkg.TrainStep(net, torch.optim.SGD(net.parameters(), lr=0.0), precision="fp32")
    loss = step.forward_loss(x, y)
    loss.backward()
    with torch.no_grad():
        logits = net.forward_logits(x)
    assert rel_err(logits.cpu(), g["logits"]) < 2e-5
    assert abs(loss.item() - float(g["loss"])) < 2e-6
    grads = dict(net.named_parameters())
    norms = np.array([grads[str(n)].grad.double().norm().item() for n in g["grad_names"]])
    np.testing.assert_allclose(norms, g["grad_norms"], rtol=5e-4, atol=1e-9)
    for k in [k for k in g if k.startswith("grad.")]:
        assert rel_err(grads[k[5:]].grad.cpu(), g[k]) < 2e-4, k


def test_cswin_model_bf16_within_stated_tolerance(no_tf32):
    """north_star: bf16 max abs error <= 2e-2 on the pre-sigmoid logits, mask agreement >= 99.9 %.

    Weights at the reference's initialisation scale (style="init"); the golden file also records the
    reference's OWN bf16 drift on the same input (CPU autocast) as the yardstick.  Random-init logits
    are all negative, so the stated mask criterion (p > 0.5, C:731) would be met vacuously.
    As H6(i) prescribes, the harness adds an output bias b_q = -quantile_q(reference logits) (the `output`
    conv has none, C:603), so that a fraction 1 - q of the pixels is foreground and the logits straddle 0:
      * q = 0.98 (2 % foreground, a small-object segmentation): mask agreement at threshold 0 >= 99.9 %;
      * q = 0.9 and q = 0.5 (the threshold sits in the densest part of the logit histogram, where ANY
        bf16 path flips the pixels inside its error band — the reference's own CPU-bf16 forward agrees
        99.83 % / 99.19 % there): within half a percent of the reference's own agreement (its CPU autocast
        keeps other intermediates in fp32 than the CUDA path does), and every flipped pixel lies inside
        the 2e-2 band."""
    g = golden("cswin_224_init_bf16.npz")
    img, batch, seed = [int(v) for v in g["meta"][:3]]
    split = [int(v) for v in g["meta"][3:]]
    cfg = om.CSWinConfig(img_size=img, split_size=split)
    net = pkg.CSWinTransformer(img_size=img, split_size=split)
    net.load_state_dict(om.synth_params(om.cswin_param_shapes(cfg), seed, style="init"))
    net.cuda()
    x = torch.tensor(g["x"]).cuda()
    with torch.no_grad():
        fp32 = net.forward_logits(x).float().cpu()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            logits = net.forward_logits(x).float().cpu()
    ref = torch.tensor(g["logits"])
    assert rel_err(fp32, ref) < 5e-5
    err = (logits - ref).abs().max().item()
    assert err <= 2e-2, err
    for q, bias, ref_agree in zip(g["mask_quantiles"], g["mask_biases"], g["ref_bf16_mask_agreement"]):
        bias = float(bias)
        want, got = (ref + bias) > 0, (logits + bias) > 0
        assert abs(want.float().mean().item() - (1 - float(q))) < 1e-3  # the masks are not degenerate
        agree = (want == got).float().mean().item()
        if float(q) >= 0.98:
            assert agree >= 0.999, (q, agree)  # the stated criterion, threshold 0 (p > 0.5, C:731)
        assert agree >= float(ref_agree) - 5e-3, (q, agree, float(ref_agree))
        assert ((ref + bias).abs()[want != got] <= 2e-2).all()


def test_cswin_512_golden_through_the_tcgen05_engines():
    """SURVEY.md 8(c) golden 'CSWin 512^2, B=2' (reference CSWinTransformer(img_size=512, split_size=[1,2,8,8]),
    fp32 CPU, init-scale weights): stripes of 128 / 256 tokens, so under bf16 autocast every LePEAttention call
    runs on the tcgen05 forward AND backward engines inside the model.  Tolerances: logits max-abs <= 2e-2
    (north_star), loss <= 2e-3 absolute, kept gradients <= 3e-2 of their max (the stated bf16 gradient
    tolerance; the reference's own CPU-bf16 backward is within 1.5e-2 on the same tensors), gradient norms:
    median deviation <= 1 %, at most 3 % of the 463 tensors off by more than 10 %."""
    g = golden("cswin_512_b2.npz")
    img, batch, seed = [int(v) for v in g["meta"][:3]]
    split = [int(v) for v in g["meta"][3:]]
    cfg = om.CSWinConfig(img_size=img, split_size=split)
    net = pkg.CSWinTransformer(img_size=img, split_size=split, attn_engine="tcgen05")
    net.load_state_dict(om.synth_params(om.cswin_param_shapes(cfg), seed, style="init"))
    net.cuda()
    gen = torch.Generator().manual_seed(300 + seed)  # == make_golden.inputs_512
    x = torch.rand((batch, 3, img, img), generator=gen)
    y = (torch.rand((batch, 1, img, img), generator=gen) > 0.5).float()
    x, y = x.cuda(), y.cuda()
    n0 = pkg.capi.launch_count()
    step = pkg.TrainStep(net, torch.optim.SGD(net.parameters(), lr=0.0), precision="bf16")
    loss = step.forward_loss(x, y)
    loss.backward()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        logits = net.forward_logits(x).float().cpu()
    assert pkg.capi.launch_count() - n0 > 100
    ref = torch.tensor(g["logits"])
    assert (logits - ref).abs().max().item() <= 2e-2
    assert abs(loss.item() - float(g["loss"])) <= 2e-3
    grads = dict(net.named_parameters())
    for k in [k for k in g if k.startswith("grad.")]:
        assert rel_err(grads[k[5:]].grad.cpu(), g[k]) < 3e-2, k
    norms = np.array([grads[str(n)].grad.double().norm().item() for n in g["grad_names"]])
    ratio = norms / g["grad_norms"]
    assert np.median(np.abs(ratio - 1)) <= 1e-2, np.median(np.abs(ratio - 1))
    off = np.abs(ratio - 1) > 0.1
    assert off.mean() <= 0.03, [(str(n), float(r)) for n, r in zip(g["grad_names"][off], ratio[off])]


def test_unet_golden_fp32_and_simam_variant(no_tf32):
    g = golden("unet_64.npz")
    net = _load(pkg.UNet(), 1)
    net.train()
    x, y = torch.tensor(g["x"]).cuda(), torch.tensor(g["y"]).cuda()
    loss = torch.nn.functional.binary_cross_entropy(net(x), y)
    loss.backward()
    with torch.no_grad():
        assert rel_err(net.forward_logits(x).cpu(), g["logits"]) < 2e-5
    grads = dict(net.named_parameters())
    norms = np.array([grads[str(n)].grad.double().norm().item() for n in g["grad_names"]])
    # plain UNet, no csb200 kernel involved: cuDNN vs CPU through 18 BatchNorms at batch 2 (ill-conditioned)
    np.testing.assert_allclose(norms, g["grad_norms"], rtol=1e-2, atol=1e-5 * g["grad_norms"].max())
    # UNet + SimAM (config 2 structure) against the oracle, forward and input gradient
    gated = pkg.UNet(simam=True)
    p = om.synth_params({k: tuple(v.shape) for k, v in gated.state_dict().items()}, 2)
    gated.load_state_dict(p)
    gated.cuda().train()
    xx = torch.rand(2, 3, 64, 64)
    out = gated.forward_logits(xx.cuda())
    ref = om.unet_logits({k: v.double() if v.is_floating_point() else v for k, v in p.items()}, xx.double(), True,
                         simam=True)
    assert rel_err(out.detach().cpu(), ref) < 5e-5


def test_cswin_simam_train_step_tracks_the_oracle(no_tf32):
    """Three AdamW steps (C:937) of CSWin+SimAM at 64^2: the loss trajectory follows the CPU oracle."""
    cfg = om.CSWinConfig(img_size=64, split_size=[1, 2, 2, 2], simam=True)
    p0 = om.synth_params(om.cswin_param_shapes(cfg), 5)
    net = pkg.CSWinTransformer(img_size=64, split_size=[1, 2, 2, 2], simam=True)
    net.load_state_dict(p0)
    net.cuda()
    step = pkg.TrainStep(net, torch.optim.AdamW(net.parameters(), lr=1e-4, weight_decay=1e-4), precision="fp32")
    names = [k for k, _ in net.named_parameters()]
    po = {k: v.clone().requires_grad_(True) for k, v in p0.items()}
    opt_o = torch.optim.AdamW([po[k] for k in names], lr=1e-4, weight_decay=1e-4)
    for it in range(3):
        x, y = pkg.synthetic_batch(2, 64, "cpu", seed=it)
        loss = step(x.cuda(), y.cuda()).item()
        opt_o.zero_grad()
        lo = torch.nn.functional.binary_cross_entropy(om.cswin_unet_forward(po, x, cfg), y)
        lo.backward()
        opt_o.step()
        assert abs(loss - lo.item()) < 2e-5 * max(1.0, abs(lo.item())), (it, loss, lo.item())


def test_bf16_train_step_runs_and_decreases_loss():
    torch.manual_seed(0)
    net = pkg.CSWinTransformer(img_size=128, split_size=[1, 2, 4, 4], simam=True).cuda()
    step = pkg.TrainStep(net, torch.optim.AdamW(net.parameters(), lr=1e-3, weight_decay=1e-4), precision="bf16")
    x, y = pkg.synthetic_batch(4, 128, "cuda", seed=1)
    losses = [step(x, y).item() for _ in range(6)]
    assert all(np.isfinite(losses)) and losses[-1] < losses[0]


def test_cuda_graph_step_matches_eager_step():
    """TrainStep(cuda_graph=True): same loss trajectory as the eager step, from the same start (the
    capture warm-up must leave weights and optimizer state untouched), host batches accepted."""
    losses = {}
    for mode in (False, True):
        torch.manual_seed(0)
        net = pkg.CSWinTransformer(img_size=64, split_size=[1, 2, 2, 2], simam=True).cuda()
        opt = torch.optim.AdamW(net.parameters(), lr=1e-3, weight_decay=1e-4, fused=True, capturable=mode)
        step = pkg.TrainStep(net, opt, precision="bf16", cuda_graph=mode)
        run = []
        for it in range(4):
            x, y = pkg.synthetic_batch(2, 64, "cpu", seed=it, pin=True)
            run.append(step(x if mode else x.cuda(), y if mode else y.cuda()).item())
        losses[mode] = run
    assert np.allclose(losses[True], losses[False], rtol=2e-3), losses
    assert losses[True][-1] != losses[True][0]


def test_train_step_with_deferred_final_sums_equals_the_immediate_step():
    """TrainStep batches the ~100 "sum the per-CTA partials" launches of backward into one
    (functional.deferred_sums): same gradients as the step that launches each sum at once (the fixed-order
    sums to 1e-6, the atomically accumulated ones to 1e-5), and fewer csb200 launches."""
    grads, launches = {}, {}
    for defer in (False, True):
        torch.manual_seed(0)
        net = pkg.CSWinTransformer(img_size=128, split_size=[1, 2, 4, 4], simam=True).cuda()
        step = pkg.TrainStep(net, torch.optim.SGD(net.parameters(), lr=0.0), precision="bf16")
        step.defer_sums = defer
        x, y = pkg.synthetic_batch(2, 128, "cuda", seed=5)
        step(x, y)
        n0 = pkg.capi.launch_count()
        step(x, y)
        launches[defer] = pkg.capi.launch_count() - n0
        grads[defer] = {k: p.grad.clone() for k, p in net.named_parameters()}
        assert pkg.capi.lib().csb200_sum_rows_pending() == 0
        assert (step._zero_arena_numel > 0) == defer  # the second step's weight-gradient outputs came from the arena
    assert launches[True] < launches[False] - 40, launches
    for k, g in grads[False].items():
        assert torch.isfinite(grads[True][k]).all(), k
        # bit for bit where the vector comes out of a deferred sum with a fixed order (LayerNorm parameters, the fc1
        # bias of the fused Mlp); biases that ride in csb200_linear_wgrad are accumulated atomically (order not fixed)
        # (bit-for-bit equality of the sums themselves: tests/test_layernorm_gpu.py; here two whole backward passes)
        if "norm" in k or k.endswith("fc1.bias"):
            assert rel_err(grads[True][k].cpu(), g.cpu()) < 1e-6, k
        else:
            assert rel_err(grads[True][k].cpu(), g.cpu()) < 1e-5, k


def test_cuda_graph_step_with_reducer_single_rank():
    """The data-parallel graph path (two graphs around the bucket all-reduce) on one rank: same
    trajectory as the plain eager step; gradients live in the reducer's flat buckets."""
    losses = {}
    for mode in ("eager", "graph+reducer"):
        torch.manual_seed(0)
        net = pkg.CSWinTransformer(img_size=64, split_size=[1, 2, 2, 2], simam=True).cuda()
        graph = mode != "eager"
        opt = torch.optim.AdamW(net.parameters(), lr=1e-3, weight_decay=1e-4, fused=True, capturable=graph)
        red = pkg.GradientAllReducer(net.parameters()) if graph else None
        step = pkg.TrainStep(net, opt, precision="bf16", reducer=red, cuda_graph=graph)
        run = []
        for it in range(4):
            x, y = pkg.synthetic_batch(2, 64, "cuda", seed=it)
            run.append(step(x, y).item())
        losses[mode] = run
    assert np.allclose(losses["graph+reducer"], losses["eager"], rtol=2e-3), losses


def test_config2_unet_simam_bf16_train_step():
    """BASELINE config 2: plain UNet + SimAM after every DoubleConv, 256^2, batch 16, bf16."""
    torch.manual_seed(0)
    net = pkg.UNet(simam=True).cuda()
    step = pkg.TrainStep(net, torch.optim.Adam(net.parameters(), lr=1e-3, weight_decay=1e-4), precision="bf16")  # U:486
    x, y = pkg.synthetic_batch(16, 256, "cuda", seed=0)
    n0 = pkg.capi.launch_count()
    losses = [step(x, y).item() for _ in range(4)]
    assert pkg.capi.launch_count() - n0 >= 4 * 18  # 9 SimAM sites, forward + backward, every step
    assert all(np.isfinite(losses)) and losses[-1] < losses[0]


@pytest.mark.parametrize("size,sw", [(1024, 1), (1024, 2), (1024, 8), (896, 7)])
def test_config5_high_res_inference_stripe_width_sweep(size, sw):
    """BASELINE config 5: 1024^2 inference with uniform stripe widths.  sw = 7 does not divide the
    1024 grid (the reference raises from view()), so that point runs at 896^2.
    Checked: bf16 logits against this package's own fp32 path (<= 2e-2, the stated bf16 tolerance)
    and one image against the CPU oracle for the cheapest case."""
    torch.manual_seed(0)
    net = pkg.CSWinTransformer(img_size=size, split_size=[sw] * 4, simam=True).cuda().eval()
    x = torch.rand(2, 3, size, size, device="cuda")
    with torch.no_grad():
        ref = net.forward_logits(x).float()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = net.forward_logits(x).float()
            prob = net(x)
    assert out.shape == (2, 1, size, size) and torch.isfinite(out).all()
    assert (out - ref).abs().max().item() <= 2e-2
    assert prob.min() >= 0 and prob.max() <= 1
    if (size, sw) == (1024, 1):
        cfg = om.CSWinConfig(img_size=size, split_size=[sw] * 4, simam=True)
        p = {k: v.detach().cpu() for k, v in net.state_dict().items()}
        with torch.no_grad():
            want = om.cswin_unet_logits(p, x[:1].cpu(), cfg)
        assert rel_err(ref[:1].cpu(), want) < 5e-5
    if sw == 7:
        with pytest.raises(RuntimeError):  # 1024/16 = 64 is not divisible by 7, exactly like the reference
            pkg.CSWinTransformer(img_size=1024, split_size=[7] * 4).cuda()(torch.rand(1, 3, 1024, 1024, device="cuda"))


def test_infer_step_replays_a_cuda_graph_and_matches_eager():
    """pkg.InferStep: eval-mode forward under no_grad, replayed from a CUDA graph (config 5 is host-bound in eager
    mode).  Same bits as the eager forward, a new input really reaches the static buffer, training mode is refused.
    448^2 with stripe width 7: stripes of 784 / 392 / 196 / 196 tokens -> the tiled tcgen05 kernel with masked tiles."""
    torch.manual_seed(0)
    net = pkg.CSWinTransformer(img_size=448, split_size=[7] * 4, simam=True).cuda().eval()
    infer = pkg.InferStep(net, precision="bf16", cuda_graph=True)
    eager = pkg.InferStep(net, precision="bf16", cuda_graph=False)
    x1, x2 = torch.rand(2, 3, 448, 448, device="cuda"), torch.rand(2, 3, 448, 448, device="cuda")
    n0 = pkg.capi.launch_count()
    y1 = infer(x1).clone()
    assert pkg.capi.launch_count() > n0
    y2 = infer(x2).clone()
    assert torch.equal(y1, eager(x1)) and torch.equal(y2, eager(x2)) and not torch.equal(y1, y2)
    with torch.no_grad():
        ref = net(x1)  # fp32, CUDA-core attention
    assert (y1.float() - ref).abs().max().item() <= 2e-2
    net.train()
    with pytest.raises(RuntimeError, match="eval"):
        infer(x1)
