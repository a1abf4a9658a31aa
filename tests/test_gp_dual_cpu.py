"""CPU: the analytic "reverse over forward" WGAN-GP gradient (oracle/gp_dual.py -- the exact operation sequence the
CUDA discriminator engine performs) against torch's double backward of oracle/port.py, which tests/golden/discriminator.pt
pins to the unmodified reference (utils/criterion.py:98-120, networks/sagan_models.py:105-168)."""
import copy

import pytest
import torch

from oracle import gp_dual, port


@pytest.mark.parametrize("shape", [(2, 65, 65), (2, 65, 129)])
def test_dual_pass_equals_double_backward(shape):
    B, H, W = shape
    torch.manual_seed(3)
    D = port.Discriminator(1, 19, 64).double()
    with torch.no_grad():
        D.attn1.gamma.fill_(0.3); D.attn2.gamma.fill_(-0.2)          # gamma = 0 would hide the attention path
        D.preprocess_additional.weight.mul_(1.3); D.preprocess_additional.bias.add_(0.2)
    D.train()
    D2 = copy.deepcopy(D)
    xs = torch.randn(B, 19, H, W, dtype=torch.float64) * 3
    xt = torch.randn(B, 19, H, W, dtype=torch.float64) * 3
    alpha = torch.rand(B, 1, 1, 1, dtype=torch.float64)
    gp = port.gradient_penalty(D, xs, xt, alpha, 10.0)
    gp.backward()
    gp2, grads = gp_dual.gp_and_param_grads(D2, alpha * xt + (1 - alpha) * xs, 10.0)
    assert abs(float(gp.detach()) - float(gp2)) < 1e-8 * abs(float(gp2))
    assert torch.allclose(D.l1[0].module.weight_u, D2.l1[0].module.weight_u)              # both advanced u,v once
    for name, p in D.named_parameters():
        if p.grad is None:
            continue
        scale = float(p.grad.norm())
        assert float((p.grad - grads[name]).norm()) <= 1e-7 * scale + 1e-15, name
