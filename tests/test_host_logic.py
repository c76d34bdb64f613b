"""CPU: host-side logic around the kernels — model assembly, state_dict contract, train step and the
data-parallel reducer (world_size 2, gloo).  The two kernel entry points are replaced by the oracle
ops here ONLY to exercise the host code without a GPU; the product path has no such fallback
(tests/test_capi.py::test_no_cpu_fallback)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import golden, rel_err
import cswin_simam_unet_b200 as pkg
from cswin_simam_unet_b200 import data_parallel, functional as csbF
from oracle import models as om, ops


@pytest.fixture
def oracle_kernels(monkeypatch):
    def cross(qkv, H, W, branches, scale, wb, engine="auto", drop_p=0.0):
        C = qkv.shape[-1] // 3
        outs = []
        for i, br in enumerate(branches):
            cs = slice(br.chan0, br.chan0 + br.chans)
            outs.append(ops.stripe_attention(qkv[..., :C][..., cs], qkv[..., C:2 * C][..., cs],
                                             qkv[..., 2 * C:][..., cs], wb[2 * i], wb[2 * i + 1], H, W,
                                             br.h_sp, br.w_sp, br.heads, scale))
        return torch.cat(outs, -1)
    monkeypatch.setattr(csbF, "cross_stripe_attention", cross)
    monkeypatch.setattr(csbF, "simam", lambda x, lam=1e-4, layout="NCHW": ops.simam(x, lam, layout))


def test_state_dict_contract():
    net = pkg.CSWinTransformer(img_size=224)
    sd = net.state_dict()
    assert len(sd) == 463
    assert {k: tuple(v.shape) for k, v in sd.items()} == om.cswin_param_shapes(om.CSWinConfig())
    assert sum(p.numel() for p in net.parameters()) == 23_567_980
    assert "output.bias" not in sd and "stage4.0.attns.1.get_v.weight" not in sd
    gated = pkg.CSWinTransformer(img_size=224, simam=True)
    assert list(gated.state_dict()) == list(sd)  # SimAM adds no keys
    assert sum(p.numel() for p in pkg.UNet().parameters()) == 31_043_521
    assert list(pkg.UNet(simam=True).state_dict()) == list(pkg.UNet().state_dict())


def test_constructor_signatures_match_reference():
    import inspect
    sig = inspect.signature(pkg.CSWinTransformer.__init__)
    ref_args = ["img_size", "patch_size", "in_chans", "num_classes", "embed_dim", "depth", "split_size", "num_heads",
                "mlp_ratio", "qkv_bias", "qk_scale", "drop_rate", "attn_drop_rate", "drop_path_rate",
                "hybrid_backbone", "norm_layer", "use_chk"]  # C:493-496
    assert list(sig.parameters)[1:1 + len(ref_args)] == ref_args
    assert sig.parameters["split_size"].default == [1, 2, 7, 7] and sig.parameters["img_size"].default == 224
    blk = inspect.signature(pkg.CSWinBlock.__init__)
    assert list(blk.parameters)[1:] == ["dim", "reso", "num_heads", "split_size", "mlp_ratio", "qkv_bias",
                                        "qk_scale", "drop", "attn_drop", "drop_path", "act_layer", "norm_layer",
                                        "last_stage"]  # C:303-307
    att = inspect.signature(pkg.LePEAttention.__init__)
    assert list(att.parameters)[1:] == ["dim", "resolution", "idx", "split_size", "dim_out", "num_heads",
                                        "attn_drop", "proj_drop", "qk_scale"]  # C:221-222
    assert att.parameters["num_heads"].default == 9
    assert list(inspect.signature(pkg.UNet.__init__).parameters)[1:3] == ["n_channels", "n_classes"]
    with pytest.raises(ValueError):
        pkg.LePEAttention(32, 8, 2, 2, num_heads=1)  # bad idx raises instead of exit(0), C:238-240


@pytest.mark.parametrize("fname", ["cswin_64.npz", "cswin_224_config1.npz"])
def test_cswin_host_assembly_matches_reference_golden(oracle_kernels, fname):
    g = golden(fname)
    img, batch, seed = [int(v) for v in g["meta"][:3]]
    split = [int(v) for v in g["meta"][3:]]
    net = pkg.CSWinTransformer(img_size=img, split_size=split)
    net.load_state_dict(om.synth_params(om.cswin_param_shapes(om.CSWinConfig(img_size=img, split_size=split)), seed))
    x, y = torch.tensor(g["x"]), torch.tensor(g["y"])
    step = pkg.TrainStep(net, torch.optim.SGD(net.parameters(), lr=0.0), precision="fp32")
    loss = step.forward_loss(x, y)
    loss.backward()
    with torch.no_grad():
        assert rel_err(net.forward_logits(x), g["logits"]) < 2e-5
    assert abs(loss.item() - float(g["loss"])) < 1e-6
    grads = dict(net.named_parameters())
    norms = np.array([grads[str(n)].grad.double().norm().item() for n in g["grad_names"]])
    np.testing.assert_allclose(norms, g["grad_norms"], rtol=2e-4, atol=1e-9)
    for k in [k for k in g if k.startswith("grad.")]:
        assert rel_err(grads[k[5:]].grad, g[k]) < 1e-4, k


def test_unet_host_assembly_matches_reference_golden(oracle_kernels):
    g = golden("unet_64.npz")
    net = pkg.UNet()
    net.load_state_dict(om.synth_params({k: tuple(v.shape) for k, v in net.state_dict().items()}, 1))
    net.train()
    assert rel_err(net.forward_logits(torch.tensor(g["x"])), g["logits"]) < 1e-5


def test_simam_placement_matches_oracle(oracle_kernels):
    cfg = om.CSWinConfig(img_size=64, split_size=[1, 2, 2, 2], simam=True)
    p = om.synth_params(om.cswin_param_shapes(cfg), 4)
    net = pkg.CSWinTransformer(img_size=64, split_size=[1, 2, 2, 2], simam=True)
    net.load_state_dict(p)
    x = torch.rand(2, 3, 64, 64)
    with torch.no_grad():
        assert rel_err(net.forward_logits(x), om.cswin_unet_logits(p, x, cfg)) < 2e-5
        plain = pkg.CSWinTransformer(img_size=64, split_size=[1, 2, 2, 2])
        plain.load_state_dict(p)
        assert rel_err(plain.forward_logits(x), net.forward_logits(x)) > 1e-3  # the gate does something
    u = pkg.UNet(simam=True)
    up = om.synth_params({k: tuple(v.shape) for k, v in u.state_dict().items()}, 2)
    u.load_state_dict(up)
    u.train()
    xx = torch.rand(2, 3, 32, 32)
    with torch.no_grad():
        assert rel_err(u.forward_logits(xx), om.unet_logits(up, xx, True, simam=True)) < 1e-5


def test_default_split_fails_at_512_like_the_reference(oracle_kernels):
    # reference: RuntimeError "shape '[1, 128, 1, 32, 4, 7]' is invalid ..." (tests/golden/..._error.txt)
    net = pkg.CSWinTransformer(img_size=512, depth=[1, 1, 1, 1])
    with pytest.raises(RuntimeError):
        net(torch.rand(1, 3, 512, 512))


def test_attn_dropout_is_passed_to_the_kernels_in_training_only():
    """attn_drop (C:246, C:290): the module hands p to the fused kernels in training mode and 0 in eval mode
    (nn.Dropout semantics); the reference's training hyper-parameters (C:930-932) construct without error."""
    att = pkg.LePEAttention(32, 8, 0, 2, num_heads=1, attn_drop=0.3)
    assert att.drop_p() == pytest.approx(0.3)
    att.eval()
    assert att.drop_p() == 0.0
    net = pkg.CSWinTransformer(img_size=64, split_size=[1, 2, 2, 2], drop_rate=0.3, attn_drop_rate=0.3,
                               drop_path_rate=0.3)
    assert all(a.drop_p() == pytest.approx(0.3) for m in net.modules() if isinstance(m, pkg.CSWinBlock) for a in m.attns)


def test_synthetic_batch_is_shard_invariant():
    full_x, full_y = pkg.synthetic_batch(8, 16, "cpu", seed=3)
    parts = [pkg.synthetic_batch(4, 16, "cpu", seed=3, first_index=r.start)
             for r in (data_parallel.shard_of_global_batch(8, k, 2) for k in range(2))]
    assert torch.equal(torch.cat([p[0] for p in parts]), full_x)
    assert torch.equal(torch.cat([p[1] for p in parts]), full_y)
    assert set(full_y.unique().tolist()) <= {0.0, 1.0} and 0 <= full_x.min() and full_x.max() < 1
    with pytest.raises(ValueError):
        data_parallel.shard_of_global_batch(10, 0, 4)


# ---------------------------------------------------------------------------------------------
# data parallel, world_size 2 over gloo
# ---------------------------------------------------------------------------------------------
def _tiny_model():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, padding=1), torch.nn.GELU(), torch.nn.Conv2d(8, 1, 1),
                               torch.nn.Sigmoid())


def _dp_worker(rank, world, port, overlap, q, reduce_dtype=None):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    model = _tiny_model()
    red = data_parallel.GradientAllReducer(model.parameters(), bucket_bytes=256, overlap=overlap,
                                           reduce_dtype=reduce_dtype)
    assert len(red.buckets) > 1 and all((b.wire is None) == (reduce_dtype is None) for b in red.buckets)
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    step = pkg.TrainStep(model, opt, precision="fp32", reducer=red)
    shard = data_parallel.shard_of_global_batch(8, rank, world)
    for it in range(2):
        x, y = pkg.synthetic_batch(len(shard), 8, "cpu", seed=it, first_index=shard.start)
        step(x, y)
    params, grads = [p.detach().numpy().copy() for p in model.parameters()], [p.grad.numpy().copy() for p in model.parameters()]
    # what a captured CUDA-graph step does: backward is REPLAYED into the buckets, begin_step() is not called
    # again, finish_step() must still average (it used to wait on the finished all-reduce of the step before)
    replay_ok = True
    for it in range(2):
        for b in red.buckets:  # (with reduce_dtype the replayed pack copy writes the wire image of the bucket)
            (b.flat if b.wire is None else b.wire).fill_(float(rank + 1 + it))
        red.finish_step()
        replay_ok &= all(torch.allclose(b.flat, torch.full_like(b.flat, (1 + world) / 2 + it)) for b in red.buckets)
    q.put((rank, params, grads, replay_ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("overlap,reduce_dtype", [(True, None), (False, None), (True, torch.bfloat16),
                                                  (False, torch.bfloat16)])
def test_gradient_allreduce_matches_global_batch(overlap, reduce_dtype):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = __import__("multiprocessing").get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, overlap, q, reduce_dtype)) for r in range(2)]
    for p in procs:
        p.start()
    results = sorted([q.get(timeout=120) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    # single-process run on the GLOBAL batch
    model = _tiny_model()
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    step = pkg.TrainStep(model, opt, precision="fp32")
    for it in range(2):
        x, y = pkg.synthetic_batch(8, 8, "cpu", seed=it)
        step(x, y)
    for (_, params, grads, replay_ok) in results:
        assert replay_ok
        for p, g, ref in zip(params, grads, model.parameters()):
            # fp32 wire: SURVEY.md 8(e)'s 1e-5; bf16 wire: each rank's gradient and the average are rounded to bf16 (2^-9 each)
            assert rel_err(p, ref.detach()) < (1e-5 if reduce_dtype is None else 1e-2)
            assert rel_err(g, ref.grad) < (1e-5 if reduce_dtype is None else 1e-2)
    for a, b in zip(results[0][1], results[1][1]):
        assert np.array_equal(a, b)  # replicas stay bit-identical


def test_reducer_single_process_manages_flat_grads():
    model = _tiny_model()
    red = data_parallel.GradientAllReducer(model.parameters(), bucket_bytes=1 << 20)
    assert red.world == 1 and len(red.buckets) == 1
    step = pkg.TrainStep(model, torch.optim.SGD(model.parameters(), lr=0.0), precision="fp32", reducer=red)
    x, y = pkg.synthetic_batch(2, 8, "cpu")
    step(x, y)
    flat = red.buckets[0].flat
    assert flat.abs().sum() > 0
    for p in model.parameters():
        assert flat.data_ptr() <= p.grad.data_ptr() < flat.data_ptr() + flat.numel() * 4
    assert red.gradient_bytes() == sum(p.numel() for p in model.parameters()) * 4


def test_train_step_defers_the_final_sums_only_when_it_is_safe(monkeypatch):
    """TrainStep._backward wraps loss.backward() in functional.deferred_sums only for CUDA losses, when enabled, and
    never while a multi-rank reducer launches all-reduces from its gradient hooks (those read the gradients before
    the block would have filled them)."""
    from cswin_simam_unet_b200 import functional as csbF
    used = []

    class FakeBlock:
        arena_demand = 123

        def __init__(self, device, zero_arena_numel=0):
            used.append(zero_arena_numel)

        def __enter__(self):
            return self

        def __exit__(self, *exc):
            return False

    class FakeLoss:
        is_cuda, device = True, "cuda:0"

        def backward(self):
            pass

    class FakeReducer:
        def __init__(self, world, overlap):
            self.world, self.overlap = world, overlap

    monkeypatch.setattr(csbF, "deferred_sums", FakeBlock)
    model = _tiny_model()
    step = pkg.TrainStep(model, torch.optim.SGD(model.parameters(), lr=0.1), precision="fp32")
    step.defer_sums = True
    step._backward(FakeLoss())
    step._backward(FakeLoss())
    assert used == [0, 123]  # the second pass sizes its zero arena with what the first one asked for
    step.reducer = FakeReducer(world=2, overlap=True)
    step._backward(FakeLoss())
    assert used == [0, 123]
    step.reducer = FakeReducer(world=2, overlap=False)  # bench.py: one bucket reduced after backward
    step._backward(FakeLoss())
    step.reducer = FakeReducer(world=1, overlap=True)   # a single rank launches nothing from its hooks
    step._backward(FakeLoss())
    assert len(used) == 4
    step.defer_sums = False
    step._backward(FakeLoss())
    cpu_loss = FakeLoss()
    cpu_loss.is_cuda = False
    step.defer_sums = True
    step._backward(cpu_loss)
    assert len(used) == 4
