"""A/B: pair-wise affinity GEMM (8 385 nodes, K = 640) with and without cta_group::2 pairs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from structure_knowledge_distillation_b200 import ops
from structure_knowledge_distillation_b200._cabi import lib
N, CS, CT, H, W = 8, 128, 512, 65, 129
fS = ops.to_nhwc(torch.randn(N, CS, H, W, device="cuda") + 0.3); fT = ops.to_nhwc(torch.randn(N, CT, H, W, device="cuda") + 0.3)
pS, arg, rS = ops.pairwise_pool(fS, 1, 1, True); pT, _, rT = ops.pairwise_pool(fT, 1, 1, False)
def timeit(fn, it=5):
    fn(); fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(it): fn()
    e.record(); torch.cuda.synchronize(); return s.elapsed_time(e) / it
fl = 2.0 * N * (H * W) ** 2 * (CS + CT)
for mode in (1, 0, 1, 0):
    lib().skd_set_conv_cta_pairs(mode)
    t = timeit(lambda: ops.pairwise_affinity_sm100(pS, pT, rS, rT, False))
    t2 = timeit(lambda: ops.pairwise_affinity_sm100(pS, pT, rS, rT, True))
    print("pairs=%d  loss-only %.3f ms (%.0f TFLOP/s)   with E store %.3f ms" % (mode, t, fl / t / 1e9, t2))
lib().skd_set_conv_cta_pairs(1)
