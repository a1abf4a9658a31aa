"""Analytic restatement of the WGAN-GP penalty's parameter gradient (TEST INFRASTRUCTURE ONLY).

The reference obtains d(GP)/d(theta) by double backward through torch autograd (utils/criterion.py:98-120:
`torch.autograd.grad(..., create_graph=True)` followed by `d_loss.backward()`, networks/kd_model.py:161-163).  The CUDA
path (csrc/disc.cu + networks/sagan_engine.py) computes the same gradient WITHOUT autograd, as "reverse over forward":

    GP(theta)   = lambda * mean_n (|g_n| - 1)^2,         g = d(sum D(x))/dx          (first-order chain to the input)
    dGP/dtheta  = d/dtheta <v, g(theta)>,  v_n = 2 lambda/N (|g_n| - 1)/|g_n| g_n   (v held constant)
                = d/dtheta  sum JVP_x[D](x; v)                                        (<v, J^T 1> = <J v, 1>)

i.e. one tangent (forward-mode) pass of D along v, then ONE reverse pass over the joint (primal, tangent) graph.  Every
layer of the SAGAN discriminator (networks/sagan_models.py:105-168) is bilinear or has a piecewise-constant mask, except
the softmax attention (:31-40) and the batch-statistics BatchNorm2d (:147), whose tangent / joint-adjoint formulas are
written out below.  This file is the CPU (fp64-capable) statement of exactly the sequence of operations the kernels
perform; tests/test_gp_dual_cpu.py pins it against torch's double backward of oracle/port.py (itself pinned to the
reference by tests/golden/discriminator.pt).
"""
import torch
import torch.nn.functional as F

LEAK = 0.1          # nn.LeakyReLU(0.1), sagan_models.py:117-130
BN_EPS = 1e-5       # nn.BatchNorm2d default


def _sn_state(sn):
    """SNConv (oracle/port.py) -> (w_bar, bias, u, v); u,v are the vectors AFTER this call's power iteration."""
    m = sn.module
    return m.weight_bar.detach(), m.bias.detach(), m.weight_u.detach(), m.weight_v.detach()


def power_iteration(w_bar, u, v, eps=1e-12):
    """networks/spectral.py:28-35: v <- normalize(W^T u), u <- normalize(W v), sigma = u.(W v)."""
    w2 = w_bar.reshape(w_bar.shape[0], -1)
    v = w2.t().mv(u); v = v / (v.norm() + eps)
    wv = w2.mv(v)
    u = wv / (wv.norm() + eps)
    return u, v, u.dot(wv)


def sn_weight_grad(d_wn, w_bar, u, v, sigma):
    """Gradient w.r.t. w_bar of any loss that sees the layer only through Wn = w_bar / sigma(w_bar), sigma = u^T W v with
    u, v constant (spectral.py:34-35: `.data` vectors): dW = dWn/sigma - <dWn, W>/sigma^2 * u v^T."""
    uv = torch.outer(u, v).reshape(w_bar.shape)
    return d_wn / sigma - (d_wn * w_bar).sum() / sigma ** 2 * uv


# ---- attention core on (B, n, C) row-major matrices (sagan_models.py:31-40) -------------------------------------------
def attn_forward(q, k, v, x, gamma):
    a = torch.softmax(q @ k.transpose(1, 2), dim=-1)          # (B, n, n): A[i, j] over keys j
    o = a @ v
    return gamma * o + x, a, o


def attn_tangent(q, k, v, a, dq, dk, dv, dx, gamma):
    """Forward-mode derivative of attn_forward along (dq, dk, dv, dx)."""
    ds = dq @ k.transpose(1, 2) + q @ dk.transpose(1, 2)
    r = (a * ds).sum(-1, keepdim=True)
    da = a * (ds - r)
    do = da @ v + a @ dv
    return gamma * do + dx, ds, da, do


def attn_joint_backward(q, k, v, a, o, gamma, gy, tq=None, tk=None, tv=None, ds=None, da=None, do=None, gty=None):
    """Adjoint of the pair (y, ydot) = (attn_forward, attn_tangent).  gy / gty are the adjoints of y / ydot.  With the
    tangent arguments None this is the ordinary first-order attention backward.
    Returns (gq, gk, gv, gx, ggamma, gtq, gtk, gtv, gtx)."""
    go = gamma * gy
    ggamma = (gy * o).sum()
    gv = a.transpose(1, 2) @ go
    ga = go @ v.transpose(1, 2)
    gtq = gtk = gtv = gtx = None
    gs_t = None
    if gty is not None:
        gto = gamma * gty
        ggamma = ggamma + (gty * do).sum()
        G = gto @ v.transpose(1, 2)                            # adjoint of Adot
        gv = gv + da.transpose(1, 2) @ gto
        gtv = a.transpose(1, 2) @ gto
        ga = ga + gto @ tv.transpose(1, 2)
        r = (a * ds).sum(-1, keepdim=True)
        g = (a * G).sum(-1, keepdim=True)
        gs_t = a * (G - g)                                     # adjoint of Sdot
        ga = ga + G * (ds - r) - g * ds
        gtx = gty
    gs = a * (ga - (a * ga).sum(-1, keepdim=True))
    gq = gs @ k
    gk = gs.transpose(1, 2) @ q
    if gty is not None:
        gq = gq + gs_t @ tk
        gk = gk + gs_t.transpose(1, 2) @ tq
        gtq = gs_t @ k
        gtk = gs_t.transpose(1, 2) @ q
    return gq, gk, gv, gy, ggamma, gtq, gtk, gtv, gtx


def _rows(t):          # (B, C, H, W) -> (B, n, C)
    return t.flatten(2).transpose(1, 2)


def _maps(t, like):    # (B, n, C) -> (B, C, H, W)
    return t.transpose(1, 2).reshape(like.shape[0], -1, like.shape[2], like.shape[3])


class _Attn:
    def __init__(self, mod):
        c = mod.value_conv.weight.shape[0]
        self.d = mod.query_conv.weight.shape[0]
        self.w = torch.cat([mod.query_conv.weight, mod.key_conv.weight, mod.value_conv.weight]).detach().reshape(-1, c)
        self.b = torch.cat([mod.query_conv.bias, mod.key_conv.bias, mod.value_conv.bias]).detach()
        self.gamma = mod.gamma.detach()
        self.c = c

    def split(self, m):
        d = self.d
        return m[..., :d], m[..., d:2 * d], m[..., 2 * d:]


def gp_and_param_grads(D, x, lambda_gp):
    """GP value and d(GP)/d(theta) for oracle/port.Discriminator `D` (preprocess mode 1, imsize-65 branch) at the
    interpolated input x (B,19,H,W), AFTER one power iteration per SN layer (the caller's D(x) would do the same).
    Returns (gp, grads: dict name -> tensor with the state-dict names of the trainable parameters)."""
    dt = x.dtype
    bn = D.preprocess_additional
    B = x.shape[0]
    sn_layers = [D.l1[0], D.l2[0], D.l3[0], D.l4[0]]
    # ---- power iteration (state advances, like a forward call)
    st = []
    for sn in sn_layers:
        m = sn.module
        u, v, sigma = power_iteration(m.weight_bar.detach().to(dt), m.weight_u.detach().to(dt), m.weight_v.detach().to(dt))
        m.weight_u.data = u.to(m.weight_u.dtype); m.weight_v.data = v.to(m.weight_v.dtype)
        st.append(dict(w=m.weight_bar.detach().to(dt), b=m.bias.detach().to(dt), u=u, v=v, sigma=sigma,
                       wn=m.weight_bar.detach().to(dt) / sigma, s=sn.stride, p=sn.padding))
    at = [_Attn(D.attn1), _Attn(D.attn2)]
    for a in at:
        a.w, a.b, a.gamma = a.w.to(dt), a.b.to(dt), a.gamma.to(dt)
    w5, b5 = D.last[0].weight.detach().to(dt), D.last[0].bias.detach().to(dt)
    gam, bet = bn.weight.detach().to(dt), bn.bias.detach().to(dt)

    # ---- 1. primal forward (batch statistics)
    mu = x.mean((0, 2, 3), keepdim=True)
    var = ((x - mu) ** 2).mean((0, 2, 3), keepdim=True)
    rstd = torch.rsqrt(var + BN_EPS)
    xh = (x - mu) * rstd
    h = gam.view(1, -1, 1, 1) * xh + bet.view(1, -1, 1, 1)
    acts = [h]                                                  # input of every conv layer
    zs = []
    att = {}
    for i, L in enumerate(st):
        z = F.conv2d(acts[-1], L["wn"], L["b"], L["s"], L["p"])
        zs.append(z)
        h = F.leaky_relu(z, LEAK)
        if i in (2, 3):
            A = at[i - 2]
            xr = _rows(h)
            qkv = xr @ A.w.t() + A.b
            q, k, v = A.split(qkv)
            y, a, o = attn_forward(q, k, v, xr, A.gamma)
            att[i] = dict(x=xr, q=q, k=k, v=v, a=a, o=o, like=h)
            h = _maps(y, h)
        acts.append(h)
    out = F.conv2d(acts[-1], w5, b5)

    # ---- 2. first-order chain to the input: g = d(sum out)/dx
    def input_grad(g_out):
        gh = F.conv_transpose2d(g_out, w5)
        for i in (3, 2, 1, 0):
            if i in (2, 3):
                A, T = at[i - 2], att[i]
                gq, gk, gv, gx, _, _, _, _, _ = attn_joint_backward(T["q"], T["k"], T["v"], T["a"], T["o"], A.gamma, _rows(gh))
                gh = _maps(torch.cat([gq, gk, gv], -1) @ A.w + gx, T["like"])
            gz = gh * torch.where(zs[i] > 0, 1.0, LEAK)
            L = st[i]
            gh = F.conv_transpose2d(gz, L["wn"], None, L["s"], L["p"],
                                    output_padding=(acts[i].shape[2] - ((gz.shape[2] - 1) * L["s"] - 2 * L["p"] + 4),
                                                    acts[i].shape[3] - ((gz.shape[3] - 1) * L["s"] - 2 * L["p"] + 4)))
        gxh = gh * gam.view(1, -1, 1, 1)
        return rstd * (gxh - gxh.mean((0, 2, 3), keepdim=True) - xh * (gxh * xh).mean((0, 2, 3), keepdim=True))

    g = input_grad(torch.ones_like(out))
    gn = g.flatten(1).norm(dim=1)
    gp = lambda_gp * ((gn - 1) ** 2).mean()
    vdir = (2 * lambda_gp / B * (gn - 1) / gn).view(B, 1, 1, 1) * g            # constant tangent direction

    # ---- 3. tangent forward along v (BN Jacobian is symmetric: same formula as its input gradient)
    t0 = rstd * (vdir - vdir.mean((0, 2, 3), keepdim=True) - xh * (vdir * xh).mean((0, 2, 3), keepdim=True))
    th = gam.view(1, -1, 1, 1) * t0
    tacts = [th]
    tatt = {}
    for i, L in enumerate(st):
        tz = F.conv2d(tacts[-1], L["wn"], None, L["s"], L["p"])
        th = tz * torch.where(zs[i] > 0, 1.0, LEAK)
        if i in (2, 3):
            A, T = at[i - 2], att[i]
            txr = _rows(th)
            tq, tk, tv = A.split(txr @ A.w.t())                                # no bias in the tangent
            ty, ds, da, do = attn_tangent(T["q"], T["k"], T["v"], T["a"], tq, tk, tv, txr, A.gamma)
            tatt[i] = dict(x=txr, q=tq, k=tk, v=tv, ds=ds, da=da, do=do)
            th = _maps(ty, T["like"])
        tacts.append(th)
    # sdot = sum conv(tacts[-1], w5)  (not needed as a value)

    # ---- 4. reverse pass over the joint graph: adjoint of sdot w.r.t. theta
    grads = {}
    ones = torch.ones_like(out)
    grads["last.0.weight"] = torch.nn.grad.conv2d_weight(tacts[-1], w5.shape, ones)
    grads["last.0.bias"] = torch.zeros_like(b5)
    gth = F.conv_transpose2d(ones, w5)                           # adjoint of the tangent activation
    gh = torch.zeros_like(gth)                                   # adjoint of the primal activation
    names = ["l1", "l2", "l3", "l4"]
    for i in (3, 2, 1, 0):
        if i in (2, 3):
            A, T, TT = at[i - 2], att[i], tatt[i]
            gq, gk, gv, gx, gg, gtq, gtk, gtv, gtx = attn_joint_backward(
                T["q"], T["k"], T["v"], T["a"], T["o"], A.gamma, _rows(gh), TT["q"], TT["k"], TT["v"], TT["ds"], TT["da"], TT["do"],
                _rows(gth))
            gqkv, gtqkv = torch.cat([gq, gk, gv], -1), torch.cat([gtq, gtk, gtv], -1)
            gw = gqkv.flatten(0, 1).t() @ T["x"].flatten(0, 1) + gtqkv.flatten(0, 1).t() @ TT["x"].flatten(0, 1)
            gb = gqkv.sum((0, 1))
            pre = "attn1." if i == 2 else "attn2."
            d, c = A.d, A.c
            grads[pre + "query_conv.weight"] = gw[:d].reshape(d, c, 1, 1); grads[pre + "query_conv.bias"] = gb[:d]
            grads[pre + "key_conv.weight"] = gw[d:2 * d].reshape(d, c, 1, 1); grads[pre + "key_conv.bias"] = gb[d:2 * d]
            grads[pre + "value_conv.weight"] = gw[2 * d:].reshape(c, c, 1, 1); grads[pre + "value_conv.bias"] = gb[2 * d:]
            grads[pre + "gamma"] = gg.reshape(1)
            gh = _maps(gqkv @ A.w + gx, T["like"])
            gth = _maps(gtqkv @ A.w + gtx, T["like"])
        mask = torch.where(zs[i] > 0, 1.0, LEAK)
        gz, gtz = gh * mask, gth * mask
        L = st[i]
        d_wn = torch.nn.grad.conv2d_weight(acts[i], L["w"].shape, gz, L["s"], L["p"]) + \
            torch.nn.grad.conv2d_weight(tacts[i], L["w"].shape, gtz, L["s"], L["p"])
        grads[names[i] + ".0.module.weight_bar"] = sn_weight_grad(d_wn, L["w"], L["u"], L["v"], L["sigma"])
        grads[names[i] + ".0.module.bias"] = gz.sum((0, 2, 3))
        op = (acts[i].shape[2] - ((gz.shape[2] - 1) * L["s"] - 2 * L["p"] + 4), acts[i].shape[3] - ((gz.shape[3] - 1) * L["s"] - 2 * L["p"] + 4))
        gh = F.conv_transpose2d(gz, L["wn"], None, L["s"], L["p"], output_padding=op)
        gth = F.conv_transpose2d(gtz, L["wn"], None, L["s"], L["p"], output_padding=op)
    grads["preprocess_additional.weight"] = (gh * xh).sum((0, 2, 3)) + (gth * t0).sum((0, 2, 3))
    grads["preprocess_additional.bias"] = gh.sum((0, 2, 3))
    return gp, grads
