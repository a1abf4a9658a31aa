"""CPU restatement of the reference's distillation step (TEST INFRASTRUCTURE ONLY).

Plain torch (CPU, fp32 by default, fp64 on request) re-derivation of the
algorithms on the hot path of irfanICMLL/structure_knowledge_distillation.
Each function cites the reference file:line it restates.  Pinned against the
reference's own Python by oracle/make_golden.py -> tests/golden/*.pt and
tests/test_oracle_golden.py.  Never imported by the product package.

Parity quirks deliberately kept:
  * ABN affine gamma = |weight| + eps, sign-corrected dweight  (libs/src/bn.cu:153,217-223)
  * biased batch variance for normalisation, n/(n-1) for running_var (libs/functions.py:90-91)
  * Pi loss is batch-summed and divided by W*H only           (utils/criterion.py:222-225)
  * Pa norm is detached, eps added outside the sqrt            (utils/utils.py:170-176)
  * Self-attention has no 1/sqrt(d); gamma init 0              (networks/sagan_models.py:19,31-40)
  * SpectralNorm: one power iteration per forward, u,v persistent (networks/spectral.py:23-35)
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# InPlace-ABN  (libs/src/bn.cu:125-232, libs/functions.py:70-162, libs/bn.py:48-105)
# --------------------------------------------------------------------------------------
ACT_NONE, ACT_LEAKY, ACT_ELU = "none", "leaky_relu", "elu"


def _bshape(x):
    return [1, x.shape[1]] + [1] * (x.dim() - 2)


def abn_batch_stats(x):
    """mean / biased variance over every axis but the channel one (bn.cu:125-138)."""
    dims = [d for d in range(x.dim()) if d != 1]
    mean = x.mean(dim=dims)
    var = ((x - mean.view(_bshape(x))) ** 2).mean(dim=dims)
    return mean, var


def abn_normalise(x, mean, var, weight, bias, eps):
    """z = (x-mean)*rsqrt(var+eps) * (|w|+eps) + b   (bn.cu:140-165)."""
    invstd = torch.rsqrt(var + eps)
    gamma = weight.abs() + eps if weight is not None else torch.ones_like(mean)
    beta = bias if bias is not None else torch.zeros_like(mean)
    y = (x - mean.view(_bshape(x))) * invstd.view(_bshape(x))
    return y * gamma.view(_bshape(x)) + beta.view(_bshape(x))


def act_forward(z, activation, slope):
    """libs/functions.py:45-51, bn.cu:302-315,333-346."""
    if activation == ACT_LEAKY:
        return torch.where(z < 0, z * slope, z)
    if activation == ACT_ELU:
        return torch.where(z < 0, torch.expm1(z), z)
    return z


class _ABNFn(torch.autograd.Function):
    """Backward written out exactly as the reference's kernels do it (z-only, y recovered from z)."""

    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, training, momentum, eps, activation, slope):
        n = x.numel() // x.shape[1]
        if training:
            mean, var = abn_batch_stats(x)
            running_mean.mul_(1 - momentum).add_(momentum * mean)                       # functions.py:90
            running_var.mul_(1 - momentum).add_(momentum * var * n / (n - 1))           # functions.py:91
        else:
            mean, var = running_mean.clone(), running_var.clone()
        z = act_forward(abn_normalise(x, mean, var, weight, bias, eps), activation, slope)
        ctx.save_for_backward(z, weight, bias, var)
        ctx.cfg = (training, eps, activation, slope, n)
        return z

    @staticmethod
    def backward(ctx, dz):
        z, weight, bias, var = ctx.saved_tensors
        training, eps, activation, slope, n = ctx.cfg
        bs = _bshape(z)
        # undo the activation (functions.py:54-62, bn.cu:317-331,348-377)
        if activation == ACT_LEAKY:
            dz = torch.where(z < 0, dz * slope, dz)
            z = torch.where(z < 0, z / slope, z)
        elif activation == ACT_ELU:
            dz = torch.where(z < 0, dz * (z + 1.0), dz)
            z = torch.where(z < 0, torch.log1p(z), z)
        gamma = weight.abs() + eps if weight is not None else torch.ones_like(var)
        beta = bias if bias is not None else torch.zeros_like(var)
        y = (z - beta.view(bs)) / gamma.view(bs)
        dims = [d for d in range(z.dim()) if d != 1]
        if training:                                                                     # bn.cu:167-184
            edz = dz.mean(dim=dims)
            eydz = (y * dz).mean(dim=dims)
        else:                                                                            # functions.py:144-147
            edz = torch.zeros_like(var)
            eydz = torch.zeros_like(var)
        mul = gamma * torch.rsqrt(var + eps)
        dx = (dz - edz.view(bs) - y * eydz.view(bs)) * mul.view(bs)                     # bn.cu:186-212
        dweight = dbias = None
        if weight is not None:
            dweight = torch.sign(weight) * eydz * n                                      # bn.cu:214-223
            dbias = edz * n                                                              # bn.cu:225-229
        return dx, dweight, dbias, None, None, None, None, None, None, None


class ABN(nn.Module):
    """Stand-in for libs.InPlaceABN / InPlaceABNSync on one device (libs/bn.py:48-105,108-193)."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, activation=ACT_LEAKY, slope=0.01):
        super().__init__()
        self.num_features, self.eps, self.momentum = num_features, eps, momentum
        self.affine, self.activation, self.slope = affine, activation, slope
        if affine:
            self.weight = nn.Parameter(torch.ones(num_features))
            self.bias = nn.Parameter(torch.zeros(num_features))
        else:
            self.register_parameter("weight", None)
            self.register_parameter("bias", None)
        self.register_buffer("running_mean", torch.zeros(num_features))
        self.register_buffer("running_var", torch.ones(num_features))

    def forward(self, x):
        return _ABNFn.apply(x, self.weight, self.bias, self.running_mean, self.running_var,
                            self.training, self.momentum, self.eps, self.activation, self.slope)


def abn_autograd_equivalent(x, weight, bias, eps, activation, slope):
    """Same forward in differentiable torch ops (training mode); used to cross-check _ABNFn.backward."""
    mean, var = abn_batch_stats(x)
    return act_forward(abn_normalise(x, mean, var, weight, bias, eps), activation, slope)


# --------------------------------------------------------------------------------------
# PSPNet / ResNet  (networks/pspnet_combine.py)
# --------------------------------------------------------------------------------------
class ChannelDropout(nn.Module):
    """nn.Dropout2d(p) with an optionally injected (N,C) keep-mask (SURVEY.md §8c RNG note)."""

    def __init__(self, p):
        super().__init__()
        self.p = p
        self.injected = None      # tensor (N, C) of {0,1}; scaled by 1/(1-p) here

    def forward(self, x):
        if not self.training:
            return x
        if self.injected is not None:
            m = self.injected.to(x.dtype).view(x.shape[0], x.shape[1], 1, 1) / (1.0 - self.p)
            return x * m
        return F.dropout2d(x, self.p, True)


def _bn_plain(c):
    return ABN(c, activation=ACT_NONE)                      # pspnet_combine.py:12


class _Basic(nn.Module):                                    # pspnet_combine.py:19-45
    expansion = 1

    def __init__(self, cin, planes, stride, dilation, downsample):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, planes, 3, stride, dilation, dilation, bias=False)
        self.bn1 = _bn_plain(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, dilation, dilation, bias=False)
        self.bn2 = _bn_plain(planes)
        self.downsample = downsample

    def forward(self, x):
        r = x if self.downsample is None else self.downsample(x)
        o = F.relu(self.bn1(self.conv1(x)))
        o = self.bn2(self.conv2(o))
        return F.relu(o + r)


class _Bottle(nn.Module):                                   # pspnet_combine.py:47-84
    expansion = 4

    def __init__(self, cin, planes, stride, dilation, downsample):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, planes, 1, bias=False)
        self.bn1 = _bn_plain(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride, dilation, dilation, bias=False)
        self.bn2 = _bn_plain(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = _bn_plain(planes * 4)
        self.downsample = downsample

    def forward(self, x):
        r = x if self.downsample is None else self.downsample(x)
        o = F.relu(self.bn1(self.conv1(x)))
        o = F.relu(self.bn2(self.conv2(o)))
        o = self.bn3(self.conv3(o))
        return F.relu(o + r)


class _PSP(nn.Module):                                      # pspnet_combine.py:86-112
    def __init__(self, features, out_features, sizes=(1, 2, 3, 6)):
        super().__init__()
        self.stages = nn.ModuleList([
            nn.Sequential(nn.AdaptiveAvgPool2d((s, s)), nn.Conv2d(features, out_features, 1, bias=False),
                          ABN(out_features)) for s in sizes])
        self.bottleneck = nn.Sequential(
            nn.Conv2d(features + len(sizes) * out_features, out_features, 3, padding=1, bias=False),
            ABN(out_features), ChannelDropout(0.1))

    def forward(self, feats):
        h, w = feats.shape[2:]
        pri = [F.interpolate(st(feats), size=(h, w), mode="bilinear", align_corners=True) for st in self.stages]
        return self.bottleneck(torch.cat(pri + [feats], 1))


class PSPNet(nn.Module):                                    # pspnet_combine.py:114-189
    CONFIGS = {"resnet18": (_Basic, (2, 2, 2, 2), 512, 128, 256),
               "resnet101": (_Bottle, (3, 4, 23, 3), 2048, 512, 1024)}

    def __init__(self, arch, num_classes):
        super().__init__()
        block, depths, c4, cpsp, c3 = self.CONFIGS[arch]
        self.conv1 = nn.Conv2d(3, 64, 3, 2, 1, bias=False); self.bn1 = _bn_plain(64)
        self.conv2 = nn.Conv2d(64, 64, 3, 1, 1, bias=False); self.bn2 = _bn_plain(64)
        self.conv3 = nn.Conv2d(64, 128, 3, 1, 1, bias=False); self.bn3 = _bn_plain(128)
        self.maxpool = nn.MaxPool2d(3, 2, 1, ceil_mode=True)                            # :130
        self._cin = 128
        self.layer1 = self._stage(block, 64, depths[0], 1, 1)
        self.layer2 = self._stage(block, 128, depths[1], 2, 1)
        self.layer3 = self._stage(block, 256, depths[2], 1, 2)
        self.layer4 = self._stage(block, 512, depths[3], 1, 4)
        self.pspmodule = _PSP(c4, cpsp)
        self.head = nn.Conv2d(cpsp, num_classes, 1, bias=True)
        self.dsn = nn.Sequential(nn.Conv2d(c3, cpsp, 3, 1, 1), ABN(cpsp), ChannelDropout(0.1),
                                 nn.Conv2d(cpsp, num_classes, 1, bias=True))

    def _stage(self, block, planes, n, stride, dilation):                                # :157-174
        ds = None
        if stride != 1 or self._cin != planes * block.expansion:
            ds = nn.Sequential(nn.Conv2d(self._cin, planes * block.expansion, 1, stride, bias=False),
                               _bn_plain(planes * block.expansion))
        blocks = [block(self._cin, planes, stride, dilation, ds)]
        self._cin = planes * block.expansion
        blocks += [block(self._cin, planes, 1, dilation, None) for _ in range(1, n)]
        return nn.Sequential(*blocks)

    def dropouts(self):
        return [self.pspmodule.bottleneck[2], self.dsn[2]]

    def forward(self, x):                                                                # :176-189
        x = F.relu(self.bn1(self.conv1(x)))
        x = F.relu(self.bn2(self.conv2(x)))
        x = F.relu(self.bn3(self.conv3(x)))
        x = self.maxpool(x)
        x1 = self.layer1(x); x2 = self.layer2(x1); x3 = self.layer3(x2)
        x_dsn = self.dsn(x3)
        x4 = self.layer4(x3)
        feat = self.pspmodule(x4)
        return [self.head(feat), x_dsn, feat, x4, x3, x2, x1]


# --------------------------------------------------------------------------------------
# SAGAN discriminator  (networks/sagan_models.py:9-41,105-168, networks/spectral.py)
# --------------------------------------------------------------------------------------
class SNConv(nn.Module):
    """Conv2d whose weight is w_bar / sigma with one power iteration per forward (spectral.py:23-35)."""

    def __init__(self, cin, cout, k, s, p):
        super().__init__()
        conv = nn.Conv2d(cin, cout, k, s, p)
        self.stride, self.padding = s, p
        self.module = nn.Module()                       # keeps state-dict names  lN.0.module.weight_bar/_u/_v/bias
        w = conv.weight.data
        self.module.bias = nn.Parameter(conv.bias.data)
        u = F.normalize(torch.randn(cout), dim=0, eps=1e-12)
        v = F.normalize(torch.randn(w[0].numel()), dim=0, eps=1e-12)
        self.module.weight_u = nn.Parameter(u, requires_grad=False)
        self.module.weight_v = nn.Parameter(v, requires_grad=False)
        self.module.weight_bar = nn.Parameter(w)

    def forward(self, x):
        m = self.module
        w2 = m.weight_bar.view(m.weight_bar.shape[0], -1)
        with torch.no_grad():
            v = w2.t().mv(m.weight_u); v = v / (v.norm() + 1e-12)
            u = w2.mv(v); u = u / (u.norm() + 1e-12)
            m.weight_v.data = v; m.weight_u.data = u          # rebinding, like spectral.py:30-31 (`.data =`)
        sigma = m.weight_u.dot(w2.mv(m.weight_v))
        return F.conv2d(x, m.weight_bar / sigma, m.bias, self.stride, self.padding)


class SelfAttn(nn.Module):                                  # sagan_models.py:9-41
    def __init__(self, c):
        super().__init__()
        self.query_conv = nn.Conv2d(c, c // 8, 1)
        self.key_conv = nn.Conv2d(c, c // 8, 1)
        self.value_conv = nn.Conv2d(c, c, 1)
        self.gamma = nn.Parameter(torch.zeros(1))

    def forward(self, x):
        b, c, h, w = x.shape
        q = self.query_conv(x).flatten(2).transpose(1, 2)
        k = self.key_conv(x).flatten(2)
        att = torch.softmax(torch.bmm(q, k), dim=-1)
        v = self.value_conv(x).flatten(2)
        o = torch.bmm(v, att.transpose(1, 2)).view(b, c, h, w)
        return self.gamma * o + x, att


class Discriminator(nn.Module):                             # sagan_models.py:105-168 (imsize 65 branch)
    def __init__(self, preprocess_mode=1, in_ch=19, conv_dim=64):
        super().__init__()
        d = conv_dim
        self.l1 = nn.Sequential(SNConv(in_ch, d, 4, 2, 1), nn.LeakyReLU(0.1))
        self.l2 = nn.Sequential(SNConv(d, 2 * d, 4, 2, 1), nn.LeakyReLU(0.1))
        self.l3 = nn.Sequential(SNConv(2 * d, 4 * d, 4, 2, 1), nn.LeakyReLU(0.1))
        self.l4 = nn.Sequential(SNConv(4 * d, 8 * d, 4, 2, 1), nn.LeakyReLU(0.1))
        self.last = nn.Sequential(nn.Conv2d(8 * d, 1, 4))
        self.attn1 = SelfAttn(4 * d)
        self.attn2 = SelfAttn(8 * d)
        self.mode = preprocess_mode
        if preprocess_mode == 1:
            self.preprocess_additional = nn.BatchNorm2d(in_ch)
        elif preprocess_mode not in (2, 3):
            raise ValueError("preprocess_GAN_mode should be 1:bn or 2:tanh or 3:-1 - 1")

    def forward(self, x):
        if self.mode == 1:
            x = self.preprocess_additional(x)
        elif self.mode == 2:
            x = torch.tanh(x)
        else:
            x = 2 * (x / 255 - 0.5)
        o = self.l3(self.l2(self.l1(x)))
        o, p1 = self.attn1(o)
        o = self.l4(o)
        o, p2 = self.attn2(o)
        return [self.last(o), p1, p2]


# --------------------------------------------------------------------------------------
# Losses  (utils/criterion.py, utils/utils.py:170-183)
# --------------------------------------------------------------------------------------
def pixelwise_loss(logits_S, logits_T):
    """criterion.py:219-226: sum_{n,h,w} -softmax(T).log_softmax(S) / (W*H)."""
    assert logits_S.shape == logits_T.shape, "the output dim of teacher and student differ"
    n, c, w, h = logits_S.shape
    p_t = torch.softmax(logits_T.detach(), dim=1)
    return -(p_t * torch.log_softmax(logits_S, dim=1)).sum() / w / h


def pool_patch(h, w, scale):
    return int(h * scale), int(w * scale)                                               # criterion.py:241-242


def affinity(feat):
    """utils.py:170-178: cosine affinity with the norm detached and eps outside the sqrt."""
    norm = (feat.detach() ** 2).sum(1, keepdim=True).sqrt() + 1e-8
    f = (feat / norm).flatten(2)
    return torch.bmm(f.transpose(1, 2), f)


def pairwise_loss(feat_S, feat_T, scale):
    """criterion.py:236-245 + utils.py:180-183."""
    ph, pw = pool_patch(feat_T.shape[2], feat_T.shape[3], scale)
    ps = F.max_pool2d(feat_S, (ph, pw), (ph, pw), 0, ceil_mode=True)
    pt = F.max_pool2d(feat_T.detach(), (ph, pw), (ph, pw), 0, ceil_mode=True)
    nodes = pt.shape[2] * pt.shape[3]
    return ((affinity(pt) - affinity(ps)) ** 2).sum() / (nodes ** 2) / pt.shape[0]


def dsn_ce_loss(preds, target, ignore_index=255):
    """criterion.py:179-188: CE(up(logits)) + 0.4*CE(up(dsn)), bilinear align_corners, mean over valid."""
    h, w = target.shape[1:]
    out = 0.0
    for p, wt in ((preds[0], 1.0), (preds[1], 0.4)):
        up = F.interpolate(p, size=(h, w), mode="bilinear", align_corners=True)
        out = out + wt * F.cross_entropy(up, target, ignore_index=ignore_index)
    return out


def adv_loss_g(d_out_S):
    return -d_out_S[0].mean()                                                            # criterion.py:129-137


def adv_loss_d(d_out_S, d_out_T, adv_type):
    """criterion.py:146-166."""
    assert d_out_S[0].shape == d_out_T[0].shape
    if adv_type == "wgan-gp":
        return -d_out_T[0].mean() + d_out_S[0].mean()
    if adv_type == "hinge":
        return F.relu(1.0 - d_out_T[0]).mean() + F.relu(1.0 + d_out_S[0]).mean()
    raise ValueError("adv_type should be wgan-gp or hinge")


def gradient_penalty(D, logits_S, logits_T, alpha, lambda_gp):
    """criterion.py:98-120 with the random alpha (N,1,1,1) injected."""
    x = (alpha * logits_T.detach() + (1 - alpha) * logits_S.detach()).requires_grad_(True)
    out = D(x)[0]
    (g,) = torch.autograd.grad(out, x, torch.ones_like(out), create_graph=True, retain_graph=True)
    gn = g.flatten(1).pow(2).sum(1).sqrt()
    return lambda_gp * ((gn - 1) ** 2).mean()


# --------------------------------------------------------------------------------------
# One distillation step in the order of networks/kd_model.py:119-173
# --------------------------------------------------------------------------------------
class StepConfig:
    def __init__(self, pi=True, pa=True, ho=True, lambda_pi=10.0, lambda_pa=0.5, lambda_d=0.1, lambda_gp=10.0,
                 pool_scale=0.5, adv_type="wgan-gp", lr_g=1e-2, lr_d=4e-4, momentum=0.9, weight_decay=5e-4):
        self.__dict__.update(locals()); del self.__dict__["self"]


def make_optimizers(student, D, cfg):
    """kd_model.py:74-75."""
    g = torch.optim.SGD(student.parameters(), cfg.lr_g, momentum=cfg.momentum, weight_decay=cfg.weight_decay)
    d = None
    if D is not None:
        d = torch.optim.SGD([p for p in D.parameters() if p.requires_grad], cfg.lr_d, momentum=cfg.momentum,
                            weight_decay=cfg.weight_decay)
    return g, d


def distill_step(teacher, student, D, images, labels, cfg, g_opt=None, d_opt=None, gp_alpha=None):
    """forward -> student_backward -> G step -> discriminator_backward -> D step.  Returns a dict of floats
    plus preds; gradients stay on the modules' .grad."""
    teacher.eval(); student.train()
    with torch.no_grad():
        preds_T = teacher(images)                                                        # kd_model.py:121-122
    preds_S = student(images)                                                            # :123
    if g_opt is not None:
        g_opt.zero_grad()
    out = {}
    ce = dsn_ce_loss(preds_S, labels); out["ce"] = float(ce)                            # :128
    G = ce
    if cfg.pi:
        t = cfg.lambda_pi * pixelwise_loss(preds_S[0], preds_T[0]); out["pi"] = float(t); G = G + t   # :131-135
    if cfg.pa:
        t = pairwise_loss(preds_S[2], preds_T[2], cfg.pool_scale); out["pa"] = float(t)
        G = G + cfg.lambda_pa * t                                                        # :143-146
    if cfg.ho:
        D.train()
        t = cfg.lambda_d * adv_loss_g(D(preds_S[0])); out["adv_g"] = float(t); G = G + t  # :147-149
    G.backward(); out["G"] = float(G)                                                    # :150-151
    out["preds_S"], out["preds_T"] = preds_S, preds_T
    if g_opt is not None:
        g_opt.step()                                                                     # :171
    if cfg.ho:
        if d_opt is not None:
            d_opt.zero_grad()                                                            # :154
        else:
            for p in D.parameters():
                p.grad = None
        dT = D(preds_T[0].detach()); dS = D(preds_S[0].detach())                        # :156-157
        dl = cfg.lambda_d * adv_loss_d(dS, dT, cfg.adv_type)
        if cfg.adv_type == "wgan-gp":
            if gp_alpha is None:
                gp_alpha = torch.rand(images.shape[0], 1, 1, 1)
            dl = dl + cfg.lambda_d * gradient_penalty(D, preds_S[0], preds_T[0], gp_alpha, cfg.lambda_gp)
        dl.backward(); out["D"] = float(dl)                                              # :163-164
        if d_opt is not None:
            d_opt.step()
    return out


def synthetic_batch(batch, h, w, seed=0, classes=19, ignore_frac=0.05):
    """SURVEY.md §8d synthetic inputs: unit-normal images, random labels with 5% ignore(255)."""
    g = torch.Generator().manual_seed(seed)
    images = torch.randn(batch, 3, h, w, generator=g)
    labels = torch.randint(0, classes, (batch, h, w), generator=g)
    labels[torch.rand(batch, h, w, generator=g) < ignore_frac] = 255
    return images, labels


def perturb_bn_stats(model, seed=0):
    """Make eval-mode BN non-trivial: running_mean~N(0,.1), running_var~U(.5,1.5)."""
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, ABN):
            m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.num_features, generator=g) + 0.5)
            with torch.no_grad():
                m.weight.copy_(1.0 + 0.2 * torch.randn(m.num_features, generator=g))
                m.bias.copy_(0.1 * torch.randn(m.num_features, generator=g))
