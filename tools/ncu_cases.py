"""One launch (after 2 warm-ups) of named kernel cases at the benchmark shapes, for `ncu --set full -k regex:<kernel>`.
   usage: python tools/ncu_cases.py <case> [<case> ...]     (cases: see CASES)"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from structure_knowledge_distillation_b200 import ops
from structure_knowledge_distillation_b200._cabi import lib

L = lib()
dev = "cuda"


def rn(*s):
    return torch.randn(*s, device=dev)


def conv_case(N, Cin, H, W, Cout, k, s, p, d, halo=0, precise=False):
    x = ops.to_nhwc(rn(N, Cin, H, W)); w = rn(Cout, k, k, Cin) / (Cin * k * k) ** 0.5
    L.skd_set_conv_halo(halo)
    fn = (lambda: ops.conv2d_fwd_3xtf32(x, w, s, p, d)) if precise else (lambda: ops.conv2d_fwd(x, w, s, p, d))
    return fn


def conv1x1_case(N, Cin, H, W, Cout, res):
    """teacher bottleneck 1x1 convolutions: folded BN (scale / shift), optional fused residual add, ReLU -- the HBM-bound launches"""
    x = ops.to_nhwc(rn(N, Cin, H, W)); w = rn(Cout, 1, 1, Cin) / Cin ** 0.5
    sc, sh = torch.rand(Cout, device=dev) + 0.5, rn(Cout) * 0.1
    r = ops.to_nhwc(rn(N, Cout, H, W)) if res else None
    out = ops.empty_nhwc(N, Cout, H, W, dev)
    return lambda: ops.conv2d_fwd(x, w, 1, 0, 1, scale=sc, shift=sh, residual=r, act="relu", out=out)


def dsn_case():
    l0 = ops.to_nhwc(rn(8, 19, 65, 129)); l1 = ops.to_nhwc(rn(8, 19, 65, 129))
    lab = torch.randint(0, 19, (8, 512, 1024), device=dev); lab[torch.rand(8, 512, 1024, device=dev) < 0.05] = 255
    return lambda: ops.dsn_ce_fwd_train(l0, l1, lab, 255, 1.0, 0.4)


def augment_case():
    from structure_knowledge_distillation_b200.dataset.datasets import DeviceAugment, draw_augmentation
    import random, numpy as np
    random.seed(0); np.random.seed(0)
    img = torch.randint(0, 256, (8, 1024, 2048, 3), dtype=torch.uint8, device=dev); lab = torch.randint(0, 34, (8, 1024, 2048), dtype=torch.uint8, device=dev)
    augs = torch.tensor([draw_augmentation(1024, 2048, (512, 1024)) for _ in range(8)], dtype=torch.float64)
    aug = DeviceAugment((512, 1024), (104.00698793, 116.66876762, 122.67891434))
    return lambda: aug(img, lab, augs)


def wgrad_case(N, Cin, H, W, Cout, k, s, p, d):
    x = ops.to_nhwc(rn(N, Cin, H, W))
    oh, ow = ops.conv_out_hw(H, W, (k, k), s, p, d)
    dy = ops.to_nhwc(rn(N, Cout, oh, ow))
    return lambda: ops.conv2d_wgrad(x, dy, (k, k), s, p, d)


def abn_case(N, C, H, W, which, res=False):
    x = ops.to_nhwc(rn(N, C, H, W)); w = torch.rand(C, device=dev) + 0.5; b = rn(C) * 0.1
    rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
    st = ops.abn_stats(x, w, b, 1e-5, 0.1, rm, rv)
    r = ops.to_nhwc(rn(N, C, H, W)) if res else None
    out = ops.abn_apply(x, st[2], st[3], "relu", 0.0, residual=r)
    dout = ops.to_nhwc(rn(N, C, H, W))
    if which == "stats":
        return lambda: ops.abn_stats(x, w, b, 1e-5, 0.1, rm, rv)
    if which == "apply":
        return lambda: ops.abn_apply(x, st[2], st[3], "relu", 0.0, residual=r)
    return lambda: ops.abn_backward(x, out if res else None, dout, st, w, 1e-5, "relu", 0.0, None, res)


CASES = {
    # weight gradients, the Cin = 64 convolutions (general kernel vs halo kernel), ABN passes
    "wgrad_512_d4": lambda: wgrad_case(8, 512, 65, 129, 512, 3, 1, 4, 4),
    "wgrad_stem_64": lambda: wgrad_case(8, 64, 256, 512, 64, 3, 1, 1, 1),
    "conv64_general": lambda: conv_case(8, 64, 256, 512, 64, 3, 1, 1, 1, halo=0),
    "conv64_halo": lambda: conv_case(8, 64, 256, 512, 64, 3, 1, 1, 1, halo=1),
    "conv64_3x_general": lambda: conv_case(8, 64, 256, 512, 64, 3, 1, 1, 1, halo=0, precise=True),
    "conv64_3x_halo": lambda: conv_case(8, 64, 256, 512, 64, 3, 1, 1, 1, halo=1, precise=True),
    "conv128_64_halo": lambda: conv_case(8, 128, 256, 512, 64, 3, 1, 1, 1, halo=1),
    "conv1x1_256_1024_res": lambda: conv1x1_case(8, 256, 65, 129, 1024, True),
    "conv1x1_1024_256": lambda: conv1x1_case(8, 1024, 65, 129, 256, False),
    "conv1x1_512_2048_res": lambda: conv1x1_case(8, 512, 65, 129, 2048, True),
    "dsn_ce_train": dsn_case,
    "cs_augment": augment_case,
    "abn_stats_stem": lambda: abn_case(8, 64, 256, 512, "stats"),
    "abn_apply_stem": lambda: abn_case(8, 64, 256, 512, "apply"),
    "abn_bwd_stem": lambda: abn_case(8, 64, 256, 512, "bwd"),
    "abn_bwd_res_l1": lambda: abn_case(8, 64, 129, 257, "bwd", res=True),
}

def time_case(name, iters=20):
    fn = CASES[name]()
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


if __name__ == "__main__":
    if sys.argv[1] == "--time":                                  # CUDA-event timings (us) instead of a profiler range
        for name in sys.argv[2:]:
            print("%-28s %8.1f us" % (name, time_case(name)))
        sys.exit(0)
    for name in sys.argv[1:]:
        fn = CASES[name]()
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        fn()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        L.skd_set_conv_halo(0)
        print(name, "done")
