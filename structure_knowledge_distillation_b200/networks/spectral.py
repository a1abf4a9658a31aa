"""Spectral normalisation wrapper with the reference's state-dict layout (networks/spectral.py): the wrapped module
keeps `<name>_bar`, `<name>_u`, `<name>_v`; one power iteration per forward; sigma differentiates w.r.t. w_bar only."""
import torch
from torch import nn
from torch.nn import Parameter


def l2normalize(v, eps=1e-12):
    return v / (v.norm() + eps)


class SpectralNorm(nn.Module):
    def __init__(self, module, name='weight', power_iterations=1):
        super().__init__()
        self.module, self.name, self.power_iterations = module, name, power_iterations
        if not hasattr(module, name + "_u"):
            w = getattr(module, name)
            height = w.shape[0]
            width = w.view(height, -1).shape[1]
            u = Parameter(l2normalize(w.data.new(height).normal_(0, 1)), requires_grad=False)
            v = Parameter(l2normalize(w.data.new(width).normal_(0, 1)), requires_grad=False)
            w_bar = Parameter(w.data)
            del module._parameters[name]
            module.register_parameter(name + "_u", u)
            module.register_parameter(name + "_v", v)
            module.register_parameter(name + "_bar", w_bar)

    def _update_u_v(self):
        m, n = self.module, self.name
        u, v, w = getattr(m, n + "_u"), getattr(m, n + "_v"), getattr(m, n + "_bar")
        w2 = w.view(w.shape[0], -1)
        with torch.no_grad():
            nu, nv = u, v
            for _ in range(self.power_iterations):
                nv = l2normalize(torch.mv(w2.t(), nu))
                nu = l2normalize(torch.mv(w2, nv))
            # persistent state advances in place (CUDA-graph safe); autograd saves the fresh nu/nv, so several forwards
            # before one backward (kd_model.py:156-161) do not trip the version counter
            v.copy_(nv); u.copy_(nu)
        sigma = nu.dot(w2.mv(nv))
        setattr(m, n, w / sigma.expand_as(w))

    def forward(self, *args):
        self._update_u_v()
        return self.module.forward(*args)
