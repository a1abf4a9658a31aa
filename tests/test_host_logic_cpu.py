"""CPU: host-side logic of the package (no kernels run): module surface / state-dict parity with the reference layout,
tensor-layout plumbing, optimizer flattening and the data-parallel exchange over gloo (world_size 2)."""
import os
import subprocess
import sys
import textwrap

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_state_dict_names_match_reference_layout():
    from oracle import port
    from structure_knowledge_distillation_b200.networks.pspnet_combine import BasicBlock, Bottleneck, Res_pspnet
    from structure_knowledge_distillation_b200.networks.sagan_models import Discriminator
    for mine, theirs in ((Res_pspnet(BasicBlock, [2, 2, 2, 2], 19), port.PSPNet("resnet18", 19)),
                         (Res_pspnet(Bottleneck, [3, 4, 23, 3], 19), port.PSPNet("resnet101", 19)),
                         (Discriminator(1, 19, 8, 65, 64), port.Discriminator(1, 19, 64))):
        a, b = mine.state_dict(), theirs.state_dict()
        assert set(a) == set(b)
        assert all(a[k].shape == b[k].shape for k in a)
        mine.load_state_dict(b)                                  # reference-layout checkpoints load as they are
    with pytest.raises(ValueError):
        Res_pspnet(BasicBlock, [1, 1, 1, 1], 19)
    with pytest.raises(ValueError):
        Discriminator(7, 19)


def test_conv_weights_live_in_ohwi_storage():
    from structure_knowledge_distillation_b200.networks.pspnet_combine import Conv2d
    from structure_knowledge_distillation_b200 import ops
    c = Conv2d(16, 32, 3, padding=1, bias=False)
    assert c.weight.shape == (32, 16, 3, 3) and c.weight.stride() == (144, 1, 48, 16)
    w = ops.weight_ohwi(c.weight)
    assert w.shape == (32, 3, 3, 16) and w.is_contiguous() and w.data_ptr() == c.weight.data_ptr()


def test_nhwc_meta_and_pitch_views():
    from structure_knowledge_distillation_b200 import ops
    t = ops.empty_nhwc(2, 8, 5, 7, "cpu")
    assert ops.nhwc_meta(t) == (2, 8, 5, 7, 8)
    wide = ops.empty_nhwc(2, 24, 5, 7, "cpu")
    assert ops.nhwc_meta(wide[:, 8:16]) == (2, 8, 5, 7, 24)
    with pytest.raises(ValueError):
        ops.nhwc_meta(torch.zeros(2, 8, 5, 7))                   # NCHW-contiguous is not NHWC
    assert ops.pixel_strides(torch.zeros(2, 8, 5, 7)) == (280, 35, 1)
    assert ops.pixel_strides(t) == (280, 1, 8)
    p = ops.pad_channels(torch.arange(2 * 3 * 2 * 2, dtype=torch.float32).view(2, 3, 2, 2), 4)
    assert ops.nhwc_meta(p) == (2, 4, 2, 2, 4) and float(p[:, 3].abs().max()) == 0 and float(p[1, 2, 1, 1]) == 23


def test_criterion_api_errors_match_reference():
    from structure_knowledge_distillation_b200.utils import criterion as C
    with pytest.raises(ValueError):
        C.CriterionAdv("lsgan")
    with pytest.raises(ValueError):
        C.CriterionAdvForG("lsgan")
    with pytest.raises(AssertionError):
        C.CriterionPixelWise()([torch.zeros(1, 19, 4, 4)], [torch.zeros(1, 19, 4, 5)])
    assert C.CriterionPairWise is C.CriterionPairWiseforWholeFeatAfterPool
    d = [torch.tensor([[0.5, -2.0]])]; t = [torch.tensor([[1.5, 0.25]])]
    assert abs(float(C.CriterionAdv("hinge")(d, t)) - (0.375 + 0.75)) < 1e-6       # relu(1-T).mean + relu(1+S).mean
    assert abs(float(C.CriterionAdv("wgan-gp")(d, t)) - (-0.875 + -0.75)) < 1e-6
    assert abs(float(C.CriterionAdvForG("hinge")(d, d)) - 0.75) < 1e-6


def test_args_and_lr_poly():
    from structure_knowledge_distillation_b200.utils.train_options import make_args
    a = make_args(batch_size=4)
    assert a.lambda_pi == 10.0 and a.lambda_pa == 0.5 and a.lambda_d == 0.1 and a.pool_scale == 0.5 and a.adv_loss_type == "wgan-gp"
    from structure_knowledge_distillation_b200.networks.kd_model import NetModel
    assert abs(NetModel.lr_poly(None, 1e-2, 20000, 40000, 0.9) - 1e-2 * 0.5 ** 0.9) < 1e-12


_DDP = textwrap.dedent('''
    import os, sys, torch, torch.distributed as dist
    sys.path.insert(0, %r)
    from structure_knowledge_distillation_b200.optim import FlatSGD
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)                                             # identical replicas
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3, padding=1), torch.nn.ReLU(), torch.nn.Conv2d(4, 2, 1))
    net[0].weight.data = net[0].weight.data.contiguous(memory_format=torch.channels_last)
    opt = FlatSGD(list(net.parameters()), 0.1, momentum=0.9, weight_decay=1e-4)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(4, 3, 5, 5, generator=g); y = torch.randn(4, 2, 5, 5, generator=g)
    shard = slice(rank * 2, rank * 2 + 2)
    opt.zero_grad()
    # per-rank loss on the local shard (sum over the local batch, like Pi) -> gradients averaged over ranks
    ((net(x[shard]) - y[shard]) ** 2).sum().backward()
    assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(net.parameters(), opt._views))   # autograd wrote into the flat buffer
    opt.all_reduce_grads(world)
    mine = opt.flat_g.clone() * opt.grad_scale
    # single-process statement of utils/parallel.py:155: mean over "GPUs" of the per-GPU gradients
    ref = torch.zeros_like(mine)
    for r in range(world):
        net2 = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3, padding=1), torch.nn.ReLU(), torch.nn.Conv2d(4, 2, 1))
        net2.load_state_dict(net.state_dict())
        ((net2(x[r * 2:r * 2 + 2]) - y[r * 2:r * 2 + 2]) ** 2).sum().backward()
        off = 0
        for p in net2.parameters():
            n = p.numel(); ref[off:off + n] += p.grad.permute(*([0, 2, 3, 1] if p.dim() == 4 else range(p.dim()))).reshape(-1) / world
            off += (n + 3) // 4 * 4
    assert torch.allclose(mine, ref, rtol=1e-5, atol=1e-6), (mine - ref).abs().max()
    dist.barrier(); dist.destroy_process_group()
    print("rank", rank, "ok")
''')


def test_flat_gradient_allreduce_gloo_world2(tmp_path):
    script = tmp_path / "ddp.py"
    script.write_text(_DDP % ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2
