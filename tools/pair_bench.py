"""Teacher / student conv shapes of the 512x1024 step: single-CTA tiles vs cta_group::2 pairs (skd_set_conv_cta_pairs)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from structure_knowledge_distillation_b200 import ops, _cabi
L = _cabi.lib()
SHAPES = [  # name, N, Cin, H, W, Cout, k, stride, pad, dil, residual
    ("t.layer3 conv1 1x1 1024->256", 8, 1024, 65, 129, 256, 1, 1, 0, 1, False),
    ("t.layer3 conv2 3x3 256->256 d2", 8, 256, 65, 129, 256, 3, 1, 2, 2, False),
    ("t.layer3 conv3 1x1 256->1024 +res", 8, 256, 65, 129, 1024, 1, 1, 0, 1, True),
    ("t.layer4 conv1 1x1 2048->512", 8, 2048, 65, 129, 512, 1, 1, 0, 1, False),
    ("t.layer4 conv2 3x3 512->512 d4", 8, 512, 65, 129, 512, 3, 1, 4, 4, False),
    ("t.layer4 conv3 1x1 512->2048 +res", 8, 512, 65, 129, 2048, 1, 1, 0, 1, True),
    ("t.layer2 conv2 3x3 128->128", 8, 128, 65, 129, 128, 3, 1, 1, 1, False),
    ("t.layer1 conv3 1x1 64->256 +res", 8, 64, 129, 257, 256, 1, 1, 0, 1, True),
    ("t.psp bottleneck 3x3 4096->512", 8, 4096, 65, 129, 512, 3, 1, 1, 1, False),
    ("s.layer3 3x3 256->256 d2", 8, 256, 65, 129, 256, 3, 1, 2, 2, False),
    ("s.layer4 3x3 512->512 d4", 8, 512, 65, 129, 512, 3, 1, 4, 4, False),
    ("s.stem conv2 3x3 64->64 @256x512", 8, 64, 256, 512, 64, 3, 1, 1, 1, False),
    ("s.stem conv3 3x3 64->128 @256x512", 8, 64, 256, 512, 128, 3, 1, 1, 1, False),
    ("s.stem dgrad 3x3 128->64 @256x512", 8, 128, 256, 512, 64, 3, 1, 1, 1, False),
    ("s.layer1 3x3 64->64 @129x257", 8, 64, 129, 257, 64, 3, 1, 1, 1, False),
    ("s.layer2 3x3 128->128 @65x129", 8, 128, 65, 129, 128, 3, 1, 1, 1, False),
]
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def timeit(fn, reps=10):
    fn(); fn(); torch.cuda.synchronize(); s.record()
    for _ in range(reps): fn()
    e.record(); torch.cuda.synchronize(); return s.elapsed_time(e) / reps
g = torch.Generator(device="cuda").manual_seed(0)
only = sys.argv[1] if len(sys.argv) > 1 else ""
if len(sys.argv) > 2: L.skd_set_conv_im2col(int(sys.argv[2])); print("im2col mode", sys.argv[2])
for name, N, Cin, H, W, Cout, k, st, p, d, with_res in SHAPES:
    if not name.startswith(only): continue
    x = ops.to_nhwc(torch.randn(N, Cin, H, W, device="cuda", generator=g))
    w = ops.weight_ohwi(torch.randn(Cout, Cin, k, k, device="cuda", generator=g) / (Cin * k * k) ** 0.5)
    OH = (H + 2 * p - d * (k - 1) - 1) // st + 1; OW = (W + 2 * p - d * (k - 1) - 1) // st + 1
    res = ops.to_nhwc(torch.randn(N, Cout, OH, OW, device="cuda", generator=g)) if with_res else None
    sc = torch.rand(Cout, device="cuda", generator=g) + 0.5; sh = torch.randn(Cout, device="cuda", generator=g)
    out = ops.to_nhwc(torch.empty(N, Cout, OH, OW, device="cuda"))
    fl = 2.0 * N * OH * OW * Cout * Cin * k * k
    ys, cols = [], []
    for pairs, ring in ((0, 1), (3, 1)) + (((0, 0), (3, 0)) if with_res else ()):
        L.skd_set_conv_cta_pairs(pairs); L.skd_set_conv_res_prefetch(ring)
        t = timeit(lambda: ops.conv2d_fwd(x, w, st, p, d, scale=sc, shift=sh, residual=res, act="relu", out=out))
        ys.append(out.clone())
        cols.append("%s%s %.3f ms %5.1f TF" % ("pair" if pairs else "single", "" if ring else "/noring", t, fl / t / 1e9))
    if name.startswith("s.stem") or name.startswith("s.layer1"):       # split-precision forward of the same shape
        for pairs in (0, 3):
            L.skd_set_conv_cta_pairs(pairs)
            t = timeit(lambda: ops.conv2d_fwd_3xtf32(x, w, st, p, d))
            cols.append("3x %s %.3f ms %5.1f TF" % ("pair" if pairs else "single", t, 3 * fl / t / 1e9))
    L.skd_set_conv_cta_pairs(1); L.skd_set_conv_res_prefetch(1)
    diff = max((ys[0] - y).abs().max().item() for y in ys[1:])
    print("%-34s %s | max diff %.2g" % (name, " | ".join(cols), diff), flush=True)
