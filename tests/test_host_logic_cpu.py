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
    with pytest.raises(AssertionError):
        C.CriterionAdv("hinge")(d, [torch.zeros(1, 3)])
    # the adversarial criteria are CUDA kernels (values checked on the GPU: tests/test_discriminator_gpu.py); no CPU fallback
    with pytest.raises(RuntimeError, match="no CPU path"):
        C.CriterionAdv("hinge")(d, t)
    with pytest.raises(RuntimeError, match="no CPU path"):
        C.CriterionAdvForG("hinge")(d, d)


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


def test_checkpoint_loaders_follow_reference_semantics(tmp_path):
    """utils/utils.py:73-151: teacher key remap, ImageNet key intersection, S/D resume from <dir>/model_best.pth.tar with the
    DataParallel 'module.' prefix stripped and last_step / start_epoch / best_mean_IU written back into args."""
    import argparse
    import numpy as np
    from structure_knowledge_distillation_b200.networks.pspnet_combine import BasicBlock, Res_pspnet
    from structure_knowledge_distillation_b200.networks.sagan_models import Discriminator
    from structure_knowledge_distillation_b200.utils import utils as U
    torch.manual_seed(0)
    src = Res_pspnet(BasicBlock, [2, 2, 2, 2], 19)
    sd = {"module." + k: v.clone() for k, v in src.state_dict().items()}
    sdir = tmp_path / "S"; sdir.mkdir()
    torch.save(dict(step=1234, epoch=5, best_mean_IU=0.61, IU_array=np.arange(19, dtype=np.float64), state_dict=sd), str(sdir / "model_best.pth.tar"))
    args = argparse.Namespace(S_ckpt_path=str(sdir), S_resume=True, is_student_load_imgnet=False, student_pretrain_model_imgnet="None",
                              D_ckpt_path=str(tmp_path / "D"), D_resume=True, last_step=0, start_epoch=0, best_mean_IU=0.0)
    dst = Res_pspnet(BasicBlock, [2, 2, 2, 2], 19)
    assert U.load_S_model(args, dst, False) == "resume"
    assert (args.last_step, args.start_epoch, args.best_mean_IU) == (1234, 5, 0.61)
    assert all(torch.equal(a, b) for a, b in zip(dst.state_dict().values(), src.state_dict().values()))
    # ImageNet initialisation: only the intersecting keys are taken, the flag wins over S_resume
    inet = {k: v.clone() + 1 for k, v in src.state_dict().items() if k.startswith("layer1.")}
    inet["fc.weight"] = torch.zeros(3)
    torch.save(inet, str(tmp_path / "inet.pth"))
    args.is_student_load_imgnet, args.student_pretrain_model_imgnet = True, str(tmp_path / "inet.pth")
    dst2 = Res_pspnet(BasicBlock, [2, 2, 2, 2], 19)
    before = dst2.state_dict()["conv1.weight"].clone()
    assert U.load_S_model(args, dst2, False) == "imagenet"
    assert torch.equal(dst2.state_dict()["layer1.0.conv1.weight"], inet["layer1.0.conv1.weight"]) and torch.equal(dst2.state_dict()["conv1.weight"], before)
    # discriminator resume (creates the directory, no file yet -> nothing loaded; then with a file)
    D = Discriminator(1, 19, 8, 65, 64)
    assert U.load_D_model(args, D, False) is None and os.path.isdir(args.D_ckpt_path)
    torch.save(dict(epoch=7, best_mean_IU=0.5, state_dict={"module." + k: v for k, v in D.state_dict().items()}), os.path.join(args.D_ckpt_path, "model_best.pth.tar"))
    D2 = Discriminator(1, 19, 8, 65, 64)
    assert U.load_D_model(args, D2, False) == "resume" and args.start_epoch == 7
    assert torch.equal(D2.l1[0].module.weight_u, D.l1[0].module.weight_u)
    # teacher remap (utils.py:78-87)
    t_sd = {}
    for k, v in src.state_dict().items():
        if k.startswith("pspmodule."):
            t_sd["head.0." + k[len("pspmodule."):]] = v
        elif k.startswith("head."):
            t_sd["head.1." + k[len("head."):]] = v
        else:
            t_sd[k] = v
    t_sd["fc.bias"] = torch.zeros(2)
    torch.save(t_sd, str(tmp_path / "teacher.pth"))
    dst3 = Res_pspnet(BasicBlock, [2, 2, 2, 2], 19)
    assert U.load_T_model(dst3, str(tmp_path / "teacher.pth"))
    assert torch.equal(dst3.state_dict()["pspmodule.bottleneck.0.weight"], src.state_dict()["pspmodule.bottleneck.0.weight"])
    assert not U.load_T_model(dst3, str(tmp_path / "missing.pth"))


def test_nvls_owned_ranges_tile_the_flat_buffer():
    """FlatSGD.step_nvls(): every float of the flat buffer is reduced / updated / multicast by exactly one rank, on float4 boundaries."""
    from structure_knowledge_distillation_b200.optim import owned_range
    for total in (4, 8, 52, 13_071_236, 3_200_004):
        for world in (1, 2, 3, 4, 8):
            edges = [owned_range(total, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == total
            for (lo, hi), (lo2, _) in zip(edges, edges[1:] + [(total, total)]):
                assert lo % 4 == 0 and hi % 4 == 0 and lo <= hi and hi == lo2
