"""torch.autograd.Functions over the sm_100a kernels (the reference's libs/functions.py role, for the whole hot path).

Conventions kept from the reference (libs/functions.py:70-162): ctx.save_for_backward for tensors, non-tensor
arguments get None gradients, backward is once-differentiable (no double backward through hand-written kernels).
"""
import torch
import torch.distributed as dist
from torch.autograd.function import once_differentiable

from . import ops


# ------------------------------------------------------------------------------------------------ losses
class PixelWiseLoss(torch.autograd.Function):
    """CriterionPixelWise math (utils/criterion.py:219-226) in one warp-reduced kernel each way."""

    @staticmethod
    def forward(ctx, logits_S, logits_T):
        ctx.save_for_backward(logits_S, logits_T)
        return ops.pixelwise_fwd(logits_S, logits_T)

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        S, T = ctx.saved_tensors
        return ops.pixelwise_bwd(S, T, g.contiguous()), None


class PairWiseLoss(torch.autograd.Function):
    """CriterionPairWiseforWholeFeatAfterPool math (utils/criterion.py:236-245, utils/utils.py:170-183)."""

    TCGEN05_MIN_NODES = 1024      # below: one SIMT tile kernel (the default pool_scale 0.5 has 9 nodes); above: tcgen05 GEMM

    @staticmethod
    def forward(ctx, feat_S, feat_T, ph, pw):
        pS, arg, rS = ops.pairwise_pool(feat_S, ph, pw, True)
        pT, _, rT = ops.pairwise_pool(feat_T, ph, pw, False)
        ctx.tensor_core = pS.shape[1] >= PairWiseLoss.TCGEN05_MIN_NODES and (pS.shape[2] + pT.shape[2]) % 4 == 0 and pS.shape[2] % 4 == 0
        need_E = feat_S.requires_grad
        if ctx.tensor_core:
            loss, E = ops.pairwise_affinity_sm100(pS, pT, rS, rT, need_E)
        else:
            loss, E = ops.pairwise_gram(pS, pT, rS, rT, True)
        ctx.save_for_backward(E, pS, rS, arg, feat_S)
        return loss

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        E, pS, rS, arg, feat_S = ctx.saved_tensors
        if ctx.tensor_core:
            return ops.pairwise_affinity_bwd_sm100(E, pS, rS, arg, g.contiguous(), feat_S), None, None, None
        return ops.pairwise_bwd(E, pS, rS, arg, g.contiguous(), feat_S), None, None, None


class DsnCrossEntropy(torch.autograd.Function):
    """CriterionDSN math (utils/criterion.py:179-188): fused upsample + log-softmax + NLL for both heads."""

    @staticmethod
    def forward(ctx, l0, l1, labels, ignore_index, w0, w1):
        # with a backward to come, the forward pass over the upsampled pixels already leaves the backward's row-phase result behind
        # (one softmax per pixel and step instead of two); inference keeps the plain forward
        ctx.fused = l0.requires_grad or (l1 is not None and l1.requires_grad)
        if ctx.fused:
            out, rows = ops.dsn_ce_fwd_train(l0, l1, labels, ignore_index, w0, w1)
            ctx.save_for_backward(l0, l1, rows, out)
            ctx.H = labels.shape[1]
        else:
            out = ops.dsn_ce_fwd(l0, l1, labels, ignore_index, w0, w1)
            ctx.save_for_backward(l0, l1, labels, out)
        ctx.cfg = (ignore_index, w0, w1)
        return out[0].clone()

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        ig, w0, w1 = ctx.cfg
        if ctx.fused:
            l0, l1, rows, out = ctx.saved_tensors
            d0, d1 = ops.dsn_ce_bwd_cols(l0, l1, rows, ctx.H, w0, w1, g.contiguous(), out[1])
        else:
            l0, l1, labels, out = ctx.saved_tensors
            d0, d1 = ops.dsn_ce_bwd(l0, l1, labels, ig, w0, w1, g.contiguous(), out[1])
        return d0, d1, None, None, None, None


# ------------------------------------------------------------------------------------------------ layers
def _padded_weight(w_ohwi, cout_p, cin_p):
    cout, kh, kw, cin = w_ohwi.shape
    if cout_p == cout and cin_p == cin:
        return w_ohwi
    wp = torch.zeros((cout_p, kh, kw, cin_p), device=w_ohwi.device, dtype=torch.float32)
    wp[:cout, :, :, :cin] = w_ohwi
    return wp


def conv_forward_padded(x, weight, bias, stride, pad, dil, **epi):
    """tcgen05 forward for any channel count: Cin / Cout that are not multiples of 4 (3-channel image, 19 classes) are
    zero-padded to the next multiple so that every pixel row is 16-byte aligned for TMA.  Returns a (N,Cout,OH,OW) view."""
    w = ops.weight_ohwi(weight)
    cout, kh, kw, cin = w.shape
    cin_p, cout_p = ops.pad4(cin), ops.pad4(cout)
    if x.shape[1] != cin_p:                                   # callers may hand in an already padded image
        x = ops.pad_channels(x, cin_p)
    wp = _padded_weight(w, cout_p, cin_p)
    if cout_p != cout:
        for k in ("scale", "shift"):
            v = epi.get(k)
            if v is not None:
                pv = torch.zeros(cout_p, device=v.device, dtype=torch.float32); pv[:cout] = v; epi[k] = pv
        if bias is not None:
            pb = torch.zeros(cout_p, device=bias.device, dtype=torch.float32); pb[:cout] = bias; bias = pb
    if bias is not None:
        sc = epi.get("scale")
        epi["shift"] = bias if epi.get("shift") is None else epi["shift"] + (bias * sc if sc is not None else bias)
    y = ops.conv2d_fwd(x, wp, stride, pad, dil, **epi)
    return y[:, :cout] if cout_p != cout else y, x


def _small_cin(weight):
    return weight.shape[1] <= 4 and weight.shape[2] * weight.shape[3] > 1


def conv_forward_small_cin(x, weight, stride, pad, dil, precise=False, **epi):
    """Tiny-Cin convolution (3-channel stem) as explicit im2col + one dense K=32 GEMM step on tcgen05."""
    cout, cin, kh, kw = weight.shape
    k = kh * kw * cin
    kp = (k + 31) // 32 * 32
    xin = x[:, :cin] if x.shape[1] != cin else x
    col = ops.im2col_small(xin, kh, kw, stride, pad, dil, kp)
    wp = torch.zeros((cout, 1, 1, kp), device=weight.device, dtype=torch.float32)
    wp[:, 0, 0, :k] = ops.weight_ohwi(weight).reshape(cout, k)
    if precise:
        return ops.conv2d_fwd_3xtf32(col, wp, 1, 0, 1), col
    return ops.conv2d_fwd(col, wp, 1, 0, 1, **epi), col


def _direct_grad(param, shape_ohwi=None):
    """The parameter's view of FlatSGD's gradient buffer when the optimizer asked for direct gradient writes (optim.FlatSGD,
    direct_grads=True) and its memory is the dense OHWI / 1-D layout the kernels produce; else None."""
    g = getattr(param, "_skd_grad", None)
    if g is None:
        return None
    if shape_ohwi is not None:
        cout, kh, kw, cin = shape_ohwi
        if tuple(g.shape) != (cout, cin, kh, kw) or g.stride() != (kh * kw * cin, 1, kw * cin, cin):
            return None
    elif g.dim() != 1 or not g.is_contiguous():
        return None
    return g


def _world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def _sync_mean(t):
    dist.all_reduce(t)
    t /= _world()


class WgradOverlap:
    """Weight-gradient kernels on a side stream.  In the backward pass only the data-gradient chain (dgrad -> ABN backward -> dgrad
    ...) is sequential; a layer's weight gradient needs (x, dy) and feeds nothing but the optimizer.  Issued on a second stream the
    tcgen05 wgrad kernels (12 K registers, no co-residency conflict) overlap the HBM-bound ABN backward passes of the chain instead
    of queueing between them.  NetModel opens a window around G_loss.backward(); tensors a side-stream kernel reads are kept alive
    until the window's join, so the caching allocator cannot hand their memory to a main-stream kernel early (valid in eager
    mode and under CUDA-graph capture alike -- no record_stream)."""
    stream = None
    keep = None

    @classmethod
    def begin(cls):
        if cls.stream is None:
            cls.stream = torch.cuda.Stream()
        cls.keep = []

    @classmethod
    def end(cls):
        if cls.keep is None:
            return
        torch.cuda.current_stream().wait_stream(cls.stream)
        cls.keep = None

    @classmethod
    def run(cls, fn, *tensors):
        """fn() on the side stream after everything enqueued so far on the current stream (no-op wrapper outside a window)."""
        if cls.keep is None:
            return fn()
        cls.stream.wait_stream(torch.cuda.current_stream())
        cls.keep.extend(t for t in tensors if t is not None)
        with torch.cuda.stream(cls.stream):
            out = fn()
        if isinstance(out, torch.Tensor):
            cls.keep.append(out)
        return out


def _grad_written(param):
    cb = getattr(param, "_skd_arrived", None)          # bucketed all-reduce bookkeeping (optim.FlatSGD.enable_overlap)
    if cb is not None:
        cb()


class Conv2d(torch.autograd.Function):
    """nn.Conv2d forward / dgrad / wgrad on tcgen05.  x is NHWC-stored; weight is the (Cout,Cin,KH,KW) parameter held in
    channels-last (OHWI) storage."""

    @staticmethod
    def forward(ctx, x, weight, bias, stride, pad, dil, precise=False):
        ctx.small = _small_cin(weight) and bias is None and not x.requires_grad
        if ctx.small:
            y, xin = conv_forward_small_cin(x, weight, stride, pad, dil, precise=precise)
        elif precise and weight.shape[0] % 4 == 0 and weight.shape[1] % 4 == 0 and ops.nhwc_meta(x)[4] == x.shape[1]:
            # split-precision (3xTF32) forward for the layers whose rounding error train-mode BN amplifies most
            xin = x
            y = ops.conv2d_fwd_3xtf32(x, ops.weight_ohwi(weight), stride, pad, dil, shift=bias)
        else:
            y, xin = conv_forward_padded(x, weight, bias, stride, pad, dil)
        ctx.save_for_backward(xin, weight)
        ctx.params = (weight, bias)                          # the Parameter objects (their flat gradient views hang on them)
        ctx.cfg = (stride, pad, dil, bias is not None, tuple(x.shape))
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        xin, weight = ctx.saved_tensors
        stride, pad, dil, has_bias, xshape = ctx.cfg
        w = ops.weight_ohwi(weight)
        cout, kh, kw, cin = w.shape
        if ctx.small:                                        # xin is the im2col matrix: the weight gradient is a single-tap GEMM
            dwc = ops.conv2d_wgrad(xin, ops.to_nhwc(dy), (1, 1), 1, 0, 1)               # [Cout][1][1][Kp] (main stream: its result is reshaped right here)
            dw = dwc.reshape(cout, -1)[:, :kh * kw * cin].reshape(cout, kh, kw, cin).permute(0, 3, 1, 2)
            return None, dw, None, None, None, None, None
        cin_p, cout_p = ops.pad4(cin), ops.pad4(cout)
        dy = ops.pad_channels(dy, cout_p) if cout_p != cout else ops.to_nhwc(dy)
        if ops.nhwc_meta(dy)[4] % 4:
            dy = ops.pad_channels(dy, cout_p)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            wp = _padded_weight(w, cout_p, cin_p)
            dx = ops.conv2d_dgrad(dy, wp, (xshape[0], cin_p, xshape[2], xshape[3]), stride, pad, dil)
            if cin_p != cin:
                dx = dx[:, :cin]
        wparam, bparam = ctx.params
        if ctx.needs_input_grad[1]:
            gview = _direct_grad(wparam, (cout, kh, kw, cin)) if (cout_p == cout and cin_p == cin) else None
            if gview is not None:                            # wgrad (and its split-K reduction) writes the flat gradient buffer itself
                WgradOverlap.run(lambda: ops.conv2d_wgrad(xin, dy, (kh, kw), stride, pad, dil, out=gview), xin, dy)
                _grad_written(wparam)
            else:
                dw = ops.conv2d_wgrad(xin, dy, (kh, kw), stride, pad, dil)
                dw = dw[:cout, :, :, :cin].permute(0, 3, 1, 2)
        if has_bias and ctx.needs_input_grad[2]:
            gview = _direct_grad(bparam) if cout_p == cout else None
            if gview is not None:
                ops.colsum(dy, out=gview)
                _grad_written(bparam)
            else:
                db = ops.colsum(dy)[:cout]
        return dx, dw, db, None, None, None, None


class ABN(torch.autograd.Function):
    """InPlaceABN(Sync) forward/backward (libs/functions.py:70-162) fused with what follows it in the network:
    the ReLU the backbone applies right after an activation='none' ABN (networks/pspnet_combine.py:36,69,177), the
    residual add + ReLU that closes a block (:42-43,81-82) and the Dropout2d after the PSP / DSN ABN (:99,143)."""

    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, training, momentum, eps, activation, slope, residual,
                chan_mul, sync=False):
        ctx.sync = bool(sync) and training and _world() > 1
        if ctx.sync:
            # InPlaceABNSync across processes (libs/functions.py:177-209): per-rank mean / biased variance, combined as
            # mean = E[mean_r], var = E[var_r + (mean - mean_r)^2] (equal per-rank counts, like the reference), running statistics
            # with the GLOBAL sample count; one small all-reduce per layer and direction
            loc = ops.abn_stats(x, weight, bias, eps, momentum, None, None)
            m = torch.stack([loc[0], loc[1] + loc[0] * loc[0]])
            dist.all_reduce(m)
            m /= _world()
            mean, var = m[0], (m[1] - m[0] * m[0]).clamp_min_(0)
            n = float(x.numel() // x.shape[1]) * _world()
            running_mean.mul_(1 - momentum).add_(mean, alpha=momentum)
            running_var.mul_(1 - momentum).add_(var, alpha=momentum * (n / (n - 1) if n > 1 else 1.0))
            sc, sh = ops.abn_fold(mean, var, weight, bias, eps)
            st = torch.stack([mean, var, sc, sh])
        elif training:
            st = ops.abn_stats(x, weight, bias, eps, momentum, running_mean, running_var)
        else:
            sc, sh = ops.abn_fold(running_mean, running_var, weight, bias, eps)
            st = torch.stack([running_mean, running_var, sc, sh])
        out = ops.abn_apply(x, st[2], st[3], activation, slope, residual=residual, chan_mul=chan_mul)
        # without a fused residual the backward recomputes the activation's sign from x (same FMA as the apply pass): the stored
        # output is not read again -- two HBM passes fewer per layer
        ctx.out_saved = residual is not None or activation == "elu"
        ctx.save_for_backward(x, out if ctx.out_saved else None, st, weight, chan_mul)
        ctx.params = (weight, bias)
        ctx.cfg = (training, eps, activation, slope, residual is not None)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, dout):
        x, out, st, weight, chan_mul = ctx.saved_tensors
        training, eps, activation, slope, has_res = ctx.cfg
        # eval mode (libs/functions.py:144-147): edz = eydz = 0 -> dx = dz * gamma * rsqrt(var + eps), and -- a quirk of the
        # reference kept as is -- dweight = sign(w) * eydz * n = 0, dbias = edz * n = 0
        wparam, bparam = ctx.params
        gw = _direct_grad(wparam) if (training and ctx.needs_input_grad[1]) else None
        gb = _direct_grad(bparam) if (training and ctx.needs_input_grad[2]) else None
        if gw is None or gb is None:
            gw = gb = None
        dx, dres, dw, db = ops.abn_backward(x, out, ops.to_nhwc(dout), st, weight, eps, activation, slope, chan_mul, has_res,
                                            training=training, dweight_out=gw, dbias_out=gb, sync_fn=_sync_mean if ctx.sync else None)
        if gw is not None:                                   # the reduce kernel wrote dweight / dbias into the flat gradient buffer
            _grad_written(wparam); _grad_written(bparam)
            dw = db = None
        return dx, dw, db, None, None, None, None, None, None, None, dres, None, None


class MaxPool3x3s2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        y, arg = ops.maxpool_fwd(x)
        ctx.save_for_backward(arg)
        ctx.shape = x.shape
        return y

    @staticmethod
    @once_differentiable
    def backward(ctx, dy):
        (arg,) = ctx.saved_tensors
        return ops.maxpool_bwd(ops.to_nhwc(dy), arg, ctx.shape)


class PspPool(torch.autograd.Function):
    """AdaptiveAvgPool2d(1/2/3/6) of the PSP module in one launch -> (N, 50, C)."""

    @staticmethod
    def forward(ctx, x, sizes):
        ctx.sizes, ctx.shape = sizes, x.shape
        return ops.psp_pool_fwd(x, list(sizes))

    @staticmethod
    @once_differentiable
    def backward(ctx, dp):
        return ops.psp_pool_bwd(dp.contiguous(), list(ctx.sizes), ctx.shape), None


class PspAssemble(torch.autograd.Function):
    """torch.cat([upsample(stage_i)..., feats], 1) written in place into one NHWC buffer (pspnet_combine.py:110-111)."""

    @staticmethod
    def forward(ctx, feats, sizes, *stages):
        n, c, h, w = feats.shape
        cs = stages[0].shape[2]
        ctot = len(stages) * cs + c
        buf = ops.empty_nhwc(n, ctot, h, w, feats.device)
        for i, (s, stg) in enumerate(zip(sizes, stages)):
            ops.psp_upsample_fwd(stg.contiguous(), s, buf, i * cs)
        ops.slice_copy(feats, 0, buf, len(stages) * cs, c)
        ctx.cfg = (sizes, cs, c, feats.shape)
        return buf

    @staticmethod
    @once_differentiable
    def backward(ctx, dbuf):
        sizes, cs, c, fshape = ctx.cfg
        dbuf = ops.to_nhwc(dbuf)
        dfe = ops.empty_nhwc(fshape[0], c, fshape[2], fshape[3], dbuf.device)
        ops.slice_copy(dbuf, len(sizes) * cs, dfe, 0, c)
        dst = tuple(ops.psp_upsample_bwd(dbuf, s, cs, i * cs) for i, s in enumerate(sizes))
        return (dfe, None) + dst
