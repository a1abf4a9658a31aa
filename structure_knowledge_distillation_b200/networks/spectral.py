"""Spectral normalisation wrapper with the reference's state-dict layout (networks/spectral.py): the wrapped module keeps
`<name>_bar`, `<name>_u`, `<name>_v`; one power iteration per forward (csrc/disc.cu: sn_power_iter_kernel, one 8-CTA
cluster combining W^T u through distributed shared memory); sigma differentiates w.r.t. w_bar only (u, v are `.data`)."""
import torch
from torch import nn
from torch.nn import Parameter


def l2normalize(v, eps=1e-12):
    return v / (v.norm() + eps)


class SpectralNorm(nn.Module):
    def __init__(self, module, name='weight', power_iterations=1):
        super().__init__()
        self.module, self.name, self.power_iterations = module, name, power_iterations
        if power_iterations != 1:
            raise ValueError("one power iteration per forward (the only setting the reference uses, sagan_models.py:117-133)")
        if not hasattr(module, name + "_u"):
            w = getattr(module, name)
            height = w.shape[0]
            width = w.numel() // height
            u = Parameter(l2normalize(w.data.new(height).normal_(0, 1)), requires_grad=False)
            v = Parameter(l2normalize(w.data.new(width).normal_(0, 1)), requires_grad=False)
            w_bar = Parameter(w.data)                      # keeps the channels-last (OHWI) storage of the wrapped conv
            del module._parameters[name]
            module.register_parameter(name + "_u", u)
            module.register_parameter(name + "_v", v)
            module.register_parameter(name + "_bar", w_bar)

    def forward(self, x):
        """Stand-alone use: y = conv(x, w_bar / sigma) + bias after one power iteration."""
        from .sagan_engine import SNConvFn
        m = self.module
        return SNConvFn.apply(m, x, m.weight_bar, m.bias)
