"""Launch a few named conv shapes once each (for `ncu -k regex:conv_`)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from structure_knowledge_distillation_b200 import ops

SHAPES = {"t1x1_256_1024_res": (8, 256, 65, 129, 1024, 1, 1, 0, 1), "t1x1_256_1024": (8, 256, 65, 129, 1024, 1, 1, 0, 1), "t1x1_1024_256": (8, 1024, 65, 129, 256, 1, 1, 0, 1),
          "s3x3_512": (8, 512, 65, 129, 512, 3, 1, 4, 4), "stem_64_128": (8, 64, 256, 512, 128, 3, 1, 1, 1)}
which = sys.argv[1:] or list(SHAPES)
for name in which:
    N, Cin, H, W, Cout, k, s, p, d = SHAPES[name]
    x = ops.to_nhwc(torch.randn(N, Cin, H, W, device="cuda"))
    w = torch.randn(Cout, k, k, Cin, device="cuda") / (Cin * k * k) ** 0.5
    res = None; sc = sh = None
    if name.endswith("_res"):
        res = ops.to_nhwc(torch.randn(N, Cout, H, W, device="cuda")); sc = torch.rand(Cout, device="cuda") + 0.5; sh = torch.randn(Cout, device="cuda")
    y = ops.conv2d_fwd(x, w, s, p, d, scale=sc, shift=sh, residual=res, act="relu" if res is not None else "none")
    for _ in range(2):
        ops.conv2d_fwd(x, w, s, p, d, scale=sc, shift=sh, residual=res, act="relu" if res is not None else "none", out=y)
    torch.cuda.synchronize()
    print(name, "done")
