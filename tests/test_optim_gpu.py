"""GPU: csb200_adam_step (FusedAdamW) vs torch.optim.AdamW / Adam — `optimizer.step()` of the reference
train loops (C:786 with C:937-941, U:348 with U:486-490) — eager and captured in the CUDA-graph train step."""
import copy

import pytest
import torch

from conftest import rel_err
import cswin_simam_unet_b200 as pkg

pytestmark = pytest.mark.gpu
SIZES = [(1,), (7,), (64, 3, 7, 7), (8192,), (8193,), (3, 33331), (100003,), (512, 512)]


@pytest.mark.parametrize("decoupled", [True, False])
def test_matches_torch_over_several_steps(decoupled):
    torch.manual_seed(0)
    ours = [torch.nn.Parameter(torch.randn(s, device="cuda")) for s in SIZES]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    kw = dict(lr=3e-3, betas=(0.9, 0.98), eps=1e-8, weight_decay=0.05)
    o = pkg.FusedAdamW(ours, decoupled=decoupled, **kw)
    t = (torch.optim.AdamW if decoupled else torch.optim.Adam)(ref, **kw)
    for it in range(6):
        for a, b in zip(ours, ref):
            g = torch.randn_like(a) * (10.0 if it == 3 else 1.0)
            a.grad, b.grad = g.clone(), g.clone()
        if it == 4:  # a scheduler changes the learning rate between steps
            o.param_groups[0]["lr"] = t.param_groups[0]["lr"] = 1e-3
        o.step()
        t.step()
    for a, b in zip(ours, ref):
        assert rel_err(a.detach().cpu(), b.detach().cpu()) < 2e-6
        assert rel_err(o.state[a]["exp_avg"].cpu(), t.state[b]["exp_avg"].cpu()) < 2e-6
        assert rel_err(o.state[a]["exp_avg_sq"].cpu(), t.state[b]["exp_avg_sq"].cpu()) < 2e-6
    assert float(o.state[ours[0]]["step"]) == 6.0


def test_shadows_written_in_the_same_pass_and_skipped_parameters():
    torch.manual_seed(1)
    ps = [torch.nn.Parameter(torch.randn(s, device="cuda")) for s in [(1000,), (5,), (4097,)]]
    o = pkg.FusedAdamW(ps, lr=1e-2)
    masters, shadows = pkg.functional.shadow_params(ps)
    o.attach_shadows(masters, shadows)
    before = ps[1].detach().clone()
    ps[0].grad, ps[2].grad = torch.randn_like(ps[0]), torch.randn_like(ps[2])  # ps[1] takes no part (grad None)
    o.step()
    assert torch.equal(shadows[0], ps[0].detach().bfloat16()) and torch.equal(shadows[2], ps[2].detach().bfloat16())
    assert torch.equal(ps[1].detach(), before)
    for p, s in zip(ps, shadows):  # still trusted by the Linear / conv layers (no version bump)
        assert pkg.functional.cast_param(p, torch.bfloat16) is s


def test_refuses_what_it_cannot_do():
    p = torch.nn.Parameter(torch.zeros(4, device="cuda", dtype=torch.bfloat16))
    p.grad = torch.ones_like(p)
    with pytest.raises(RuntimeError, match="fp32"):
        pkg.FusedAdamW([p]).step()
    with pytest.raises(ValueError):
        pkg.FusedAdamW([torch.nn.Parameter(torch.zeros(1))], lr=-1.0)


@pytest.mark.parametrize("cuda_graph", [False, True])
def test_train_step_with_fused_adamw_matches_torch_adamw(cuda_graph, no_tf32):
    """Three fp32 steps of the CSWin-UNet (64^2, batch 2): same losses and weights as torch.optim.AdamW."""
    def run(make_opt, graph):
        torch.manual_seed(0)
        net = pkg.CSWinTransformer(img_size=64, split_size=[1, 2, 2, 2], simam=True).cuda()
        step = pkg.TrainStep(net, make_opt(net.parameters()), precision="fp32", cuda_graph=graph)
        losses = []
        for s in range(3):
            x, y = pkg.synthetic_batch(2, 64, "cuda", seed=s)
            losses.append(step(x, y).item())
        return losses, [p.detach().clone() for p in net.parameters()]
    la, wa = run(lambda ps: pkg.FusedAdamW(ps, lr=1e-3, weight_decay=1e-4), cuda_graph)
    lb, wb = run(lambda ps: torch.optim.AdamW(ps, lr=1e-3, weight_decay=1e-4), False)
    assert max(abs(a - b) for a, b in zip(la, lb)) < 1e-4
    # Adam divides by sqrt(v): entries with tiny gradients amplify rounding, so compare the update, not bits
    num = sum(float((a - b).double().pow(2).sum()) for a, b in zip(wa, wb)) ** 0.5
    den = sum(float(b.double().pow(2).sum()) for b in wb) ** 0.5
    assert num / den < 1e-4


def test_bf16_graph_step_keeps_the_shadows_in_sync():
    torch.manual_seed(0)
    net = pkg.CSWinTransformer(img_size=64, split_size=[1, 2, 2, 2], simam=True).cuda()
    step = pkg.TrainStep(net, pkg.FusedAdamW(net.parameters(), lr=1e-3, weight_decay=1e-4), precision="bf16",
                         cuda_graph=True)
    for s in range(3):
        x, y = pkg.synthetic_batch(2, 64, "cuda", seed=s)
        loss = step(x, y)
    assert torch.isfinite(loss)
    for p in net.parameters():
        sh = getattr(p, "_csb_shadow", None)
        assert sh is not None and torch.equal(sh[0], p.detach().bfloat16())


@pytest.mark.parametrize("fused", [True, False])
def test_graph_capture_keeps_restored_optimizer_state(fused, no_tf32):
    """A checkpoint restored with optimizer.load_state_dict (or eager steps taken) BEFORE the first captured
    step must survive the capture's warm-up: the captured run continues from moments / step count of the
    checkpoint exactly like the eager run does (_capture used to zero every state tensor)."""
    def make(net):
        if fused:
            return pkg.FusedAdamW(net.parameters(), lr=1e-3, weight_decay=1e-4)
        return torch.optim.AdamW(net.parameters(), lr=1e-3, weight_decay=1e-4, capturable=True)

    torch.manual_seed(0)
    net0 = pkg.CSWinTransformer(img_size=64, split_size=[1, 2, 2, 2], simam=True).cuda()
    step0 = pkg.TrainStep(net0, make(net0), precision="fp32")
    for s in range(3):
        step0(*pkg.synthetic_batch(2, 64, "cuda", seed=s))
    ckpt_model = {k: v.clone() for k, v in net0.state_dict().items()}
    ckpt_opt = copy.deepcopy(step0.optimizer.state_dict())

    def resume(graph):
        net = pkg.CSWinTransformer(img_size=64, split_size=[1, 2, 2, 2], simam=True).cuda()
        net.load_state_dict(ckpt_model)
        opt = make(net)
        opt.load_state_dict(copy.deepcopy(ckpt_opt))  # load_state_dict aliases tensors of matching dtype / device
        step = pkg.TrainStep(net, opt, precision="fp32", cuda_graph=graph)
        losses = [step(*pkg.synthetic_batch(2, 64, "cuda", seed=10 + s)).item() for s in range(2)]
        p0 = next(iter(opt.state))
        return losses, [p.detach().clone() for p in net.parameters()], float(opt.state[p0]["step"])

    le, we, se = resume(False)
    lg, wg, sg = resume(True)
    assert se == sg == 5.0  # 3 checkpointed steps + 2, not 2 (bias correction did not restart)
    assert max(abs(a - b) for a, b in zip(le, lg)) < 1e-5
    num = sum(float((a - b).double().pow(2).sum()) for a, b in zip(we, wg)) ** 0.5
    den = sum(float(b.double().pow(2).sum()) for b in we) ** 0.5
    assert num / den < 1e-5
