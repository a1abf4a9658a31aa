"""SAGAN discriminator of the holistic (Ho) loss -- same classes, arguments, state-dict names and [out, p1, p2] output
as networks/sagan_models.py:9-41,105-168.

Scope note: D is ~0.26 GMAC per image (0.03 % of the step) and its WGAN-GP penalty needs a double
backward (utils/criterion.py:105-116), which hand-written once-differentiable kernels cannot provide.  It therefore
runs on torch's differentiable CUDA operators in this round; it is listed as the next kernel target in DESIGN.md.
"""
import torch
import torch.nn as nn

from .spectral import SpectralNorm


class Self_Attn(nn.Module):
    """Self attention layer: softmax(Q^T K) without 1/sqrt(d), out = gamma * (V A^T) + x, gamma initialised to 0."""

    def __init__(self, in_dim, activation=None):
        super().__init__()
        self.chanel_in, self.activation = in_dim, activation
        self.query_conv = nn.Conv2d(in_dim, in_dim // 8, 1)
        self.key_conv = nn.Conv2d(in_dim, in_dim // 8, 1)
        self.value_conv = nn.Conv2d(in_dim, in_dim, 1)
        self.gamma = nn.Parameter(torch.zeros(1))
        self.softmax = nn.Softmax(dim=-1)

    def forward(self, x):
        b, c, w, h = x.size()
        q = self.query_conv(x).reshape(b, -1, w * h).permute(0, 2, 1)
        k = self.key_conv(x).reshape(b, -1, w * h)
        attention = self.softmax(torch.bmm(q, k))
        v = self.value_conv(x).reshape(b, -1, w * h)
        out = torch.bmm(v, attention.permute(0, 2, 1)).reshape(b, c, w, h)
        return self.gamma * out + x, attention


class Discriminator(nn.Module):
    """Discriminator, Auxiliary Classifier."""

    def __init__(self, preprocess_GAN_mode, input_channel, batch_size=64, image_size=64, conv_dim=64):
        super().__init__()
        self.imsize = image_size
        d = conv_dim
        self.l1 = nn.Sequential(SpectralNorm(nn.Conv2d(input_channel, d, 4, 2, 1)), nn.LeakyReLU(0.1))
        self.l2 = nn.Sequential(SpectralNorm(nn.Conv2d(d, d * 2, 4, 2, 1)), nn.LeakyReLU(0.1))
        self.l3 = nn.Sequential(SpectralNorm(nn.Conv2d(d * 2, d * 4, 4, 2, 1)), nn.LeakyReLU(0.1))
        curr = d * 4
        if self.imsize == 65:                              # sagan_models.py:131-136
            self.l4 = nn.Sequential(SpectralNorm(nn.Conv2d(curr, curr * 2, 4, 2, 1)), nn.LeakyReLU(0.1))
            curr *= 2
        self.last = nn.Sequential(nn.Conv2d(curr, 1, 4))
        self.attn1 = Self_Attn(256, 'relu')
        self.attn2 = Self_Attn(512, 'relu')
        if preprocess_GAN_mode == 1:
            self.preprocess_additional = nn.BatchNorm2d(input_channel)
        elif preprocess_GAN_mode == 2:
            self.preprocess_additional = nn.Tanh()
        elif preprocess_GAN_mode == 3:
            self.preprocess_additional = lambda x: 2 * (x / 255 - 0.5)
        else:
            raise ValueError('preprocess_GAN_mode should be 1:bn or 2:tanh or 3:-1 - 1')

    def forward(self, x):
        x = self.preprocess_additional(x)
        out = self.l3(self.l2(self.l1(x)))
        out, p1 = self.attn1(out)
        out = self.l4(out)
        out, p2 = self.attn2(out)
        return [self.last(out), p1, p2]
