"""GPU bring-up probe (not a test): staged checks with flushed prints so a hang is attributable, TF32 operand-mode
diagnosis, and first timings of the tcgen05 kernels at the real 512x1024 layer shapes."""
import os
import sys
import time

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from structure_knowledge_distillation_b200 import ops  # noqa: E402
from structure_knowledge_distillation_b200._cabi import lib  # noqa: E402


def log(*a):
    print(*a, flush=True)


def rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def trunc_tf32(t):
    return (t.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)


def rna_tf32(t):
    out = torch.empty_like(t.contiguous())
    lib().skd_round_tf32(t.numel(), t.contiguous().data_ptr(), out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    return out


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def stage_conv(N, Cin, H, W, Cout, k, s, p, d, tag):
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(N, Cin, H, W, device="cuda", generator=g)
    w = torch.randn(Cout, Cin, k, k, device="cuda", generator=g) / (Cin * k * k) ** 0.5
    xc, wo = ops.to_nhwc(x), ops.weight_ohwi(w)
    log("  [%s] launching fwd" % tag)
    y = ops.conv2d_fwd(xc, wo, s, p, d)
    torch.cuda.synchronize()
    ref = F.conv2d(x.double(), w.double(), None, s, p, d)
    r_t = F.conv2d(trunc_tf32(x).double(), trunc_tf32(w).double(), None, s, p, d)
    r_r = F.conv2d(rna_tf32(x).double(), rna_tf32(w).double(), None, s, p, d)
    log("  [%s] fwd rel err vs exact %.3e | vs truncated-operand model %.3e | vs RNA-rounded model %.3e | mean signed bias %.3e"
        % (tag, rel(y, ref), rel(y, r_t), rel(y, r_r), float(((y.double() - ref) * ref.sign()).mean() / ref.abs().mean())))
    return x, w, xc, wo, ref


def main():
    log(torch.cuda.get_device_name(0), torch.version.cuda)
    L = lib()
    for mode in (1, 0):
        L.skd_set_tf32_tma_type(mode)
        log("TMA data type:", "TFLOAT32" if mode else "FLOAT32")
        stage_conv(1, 32, 8, 16, 32, 1, 1, 0, 1, "1x1 32->32 one tile")
        stage_conv(2, 64, 9, 13, 64, 3, 1, 1, 1, "3x3 64->64 ragged")
        stage_conv(1, 256, 12, 9, 512, 3, 1, 4, 4, "3x3 d4 256->512")
        stage_conv(2, 64, 33, 31, 128, 3, 2, 1, 1, "3x3 s2 (elementStrides)")
    L.skd_set_tf32_tma_type(1)
    # wgrad / dgrad
    for case in [(2, 64, 9, 13, 64, 3, 1, 1, 1), (1, 128, 17, 19, 256, 3, 1, 2, 2), (2, 64, 33, 31, 128, 3, 2, 1, 1)]:
        N, Cin, H, W, Cout, k, s, p, d = case
        g = torch.Generator(device="cuda").manual_seed(2)
        x = torch.randn(N, Cin, H, W, device="cuda", generator=g).double().requires_grad_(True)
        w = (torch.randn(Cout, Cin, k, k, device="cuda", generator=g) / (Cin * k * k) ** 0.5).double().requires_grad_(True)
        y = F.conv2d(x, w, None, s, p, d)
        dy = torch.randn(y.shape, device="cuda", generator=g)
        y.backward(dy.double())
        xc, dyc, wo = ops.to_nhwc(x.detach().float()), ops.to_nhwc(dy), ops.weight_ohwi(w.detach().float())
        log("  wgrad case", case, "launching")
        dw = ops.conv2d_wgrad(xc, dyc, (k, k), s, p, d)
        torch.cuda.synchronize()
        log("  wgrad tcgen05 rel err %.3e ; direct %.3e" % (rel(dw, w.grad.permute(0, 2, 3, 1)),
            rel(ops.conv2d_wgrad(xc, dyc, (k, k), s, p, d, force_direct=True), w.grad.permute(0, 2, 3, 1))))
        if s == 1:
            dx = ops.conv2d_dgrad(dyc, wo, x.shape, s, p, d)
            torch.cuda.synchronize()
            log("  dgrad tcgen05 rel err %.3e" % rel(dx, x.grad))
    # timings at the real shapes (batch 8, 1/8 resolution 65x129) -- cold L2 not enforced here, indicative only
    log("timings (ms, TFLOP/s):")
    shapes = [("S layer4 3x3 d4 512->512", 8, 512, 65, 129, 512, 3, 1, 4, 4),
              ("T psp bottleneck 3x3 4096->512", 8, 4096, 65, 129, 512, 3, 1, 1, 1),
              ("T 1x1 256->1024", 8, 256, 65, 129, 1024, 1, 1, 0, 1),
              ("T 1x1 1024->256", 8, 1024, 65, 129, 256, 1, 1, 0, 1),
              ("T 3x3 d2 256->256", 8, 256, 65, 129, 256, 3, 1, 2, 2),
              ("S stem 3x3 64->128 @256x512", 8, 64, 256, 512, 128, 3, 1, 1, 1),
              ("S layer1 3x3 64->64 @129x257", 8, 64, 129, 257, 64, 3, 1, 1, 1)]
    for name, N, Cin, H, W, Cout, k, s, p, d in shapes:
        x = ops.to_nhwc(torch.randn(N, Cin, H, W, device="cuda"))
        w = torch.randn(Cout, k, k, Cin, device="cuda") / (Cin * k * k) ** 0.5
        y = ops.conv2d_fwd(x, w, s, p, d)
        fl = 2.0 * y.numel() * Cin * k * k
        t = timeit(lambda: ops.conv2d_fwd(x, w, s, p, d, out=y))
        dy = ops.to_nhwc(torch.randn(*y.shape, device="cuda"))
        tw = timeit(lambda: ops.conv2d_wgrad(x, dy, (k, k), s, p, d), iters=5, warm=2)
        td = timeit(lambda: ops.conv2d_dgrad(dy, w, x.shape, s, p, d), iters=5, warm=2)
        torch.backends.cudnn.allow_tf32 = True; torch.backends.cuda.matmul.allow_tf32 = True
        xt = x.contiguous(memory_format=torch.channels_last); wt = w.permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)
        tc = timeit(lambda: F.conv2d(xt, wt, None, s, p, d))
        log("  %-34s fwd %.3f ms %.0f TF | wgrad %.3f ms %.0f TF | dgrad %.3f ms %.0f TF | cuDNN-tf32 fwd %.3f ms %.0f TF" %
            (name, t, fl / t / 1e9, tw, fl / tw / 1e9, td, fl / td / 1e9, tc, fl / tc / 1e9))
    # ABN kernels vs the reference's bn.cu (HBM GB/s)
    import ctypes
    refp = os.path.join(ROOT, "oracle", "_ref", "libbn_ref.so")
    ref = ctypes.CDLL(refp) if os.path.exists(refp) else None
    for (N, C, S) in [(8, 64, 256 * 512), (8, 128, 65 * 129), (8, 512, 65 * 129)]:
        x = torch.randn(N, C, S, device="cuda"); mean = torch.empty(C, device="cuda"); var = torch.empty(C, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        vp = ctypes.c_void_p
        t_m = timeit(lambda: L.skd_bn_mean_var_cuda(N, C, S, x.data_ptr(), mean.data_ptr(), var.data_ptr(), st))
        t_r = timeit(lambda: ref._bn_mean_var_cuda(N, C, S, vp(x.data_ptr()), vp(mean.data_ptr()), vp(var.data_ptr()), vp(st))) if ref else float("nan")
        xh = ops.to_nhwc(x.view(N, C, S, 1))
        t_h = timeit(lambda: ops.abn_stats(xh, None, None, 1e-5, 0.1, None, None))
        gb = x.numel() * 4 / 1e6
        log("  mean_var (N,C,S)=(%d,%d,%d): ours NCHW %.3f ms %.0f GB/s | reference bn.cu %.3f ms %.0f GB/s (2 passes) | ours NHWC %.3f ms %.0f GB/s"
            % (N, C, S, t_m, gb / t_m, t_r, gb / t_r, t_h, gb / t_h))


if __name__ == "__main__":
    main()
