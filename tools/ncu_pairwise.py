"""Pair-wise affinity at pool_scale = 1/65 (8 385 nodes per image, batch 8): tcgen05 GEMM timing / ncu target."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from structure_knowledge_distillation_b200 import functions as Fn, ops

N, CS, CT, H, W = 8, 128, 512, 65, 129
fS = ops.to_nhwc(torch.randn(N, CS, H, W, device="cuda") + 0.3).requires_grad_(True)
fT = ops.to_nhwc(torch.randn(N, CT, H, W, device="cuda") + 0.3)
nodes = H * W
def timeit(fn, it=3):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(it): fn()
    e.record(); torch.cuda.synchronize(); return s.elapsed_time(e) / it
pS, arg, rS = ops.pairwise_pool(fS.detach(), 1, 1, True); pT, _, rT = ops.pairwise_pool(fT, 1, 1, False)
t_fwd = timeit(lambda: ops.pairwise_affinity_sm100(pS, pT, rS, rT, False))
fl = 2.0 * N * nodes * nodes * (CS + CT)
print("affinity fwd (loss only, E never written): %.3f ms  %.0f TFLOP/s (2*N*nodes^2*(C_S+C_T) = %.0f GFLOP)" % (t_fwd, fl / t_fwd / 1e9, fl / 1e9))
t_fwdE = timeit(lambda: ops.pairwise_affinity_sm100(pS, pT, rS, rT, True))
print("affinity fwd + E store (training): %.3f ms" % t_fwdE)
def full():
    fS.grad = None
    Fn.PairWiseLoss.apply(fS, fT, 1, 1).backward()
print("criterion fwd+bwd end to end: %.3f ms" % timeit(full, 2))
