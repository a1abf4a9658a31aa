"""Timing of the fused DSN cross-entropy training forward (loss + row phase of the backward) and its column phase at the benchmark shape."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from structure_knowledge_distillation_b200 import ops

torch.manual_seed(0)
l0 = ops.to_nhwc(torch.randn(8, 19, 65, 129, device="cuda")); l1 = ops.to_nhwc(torch.randn(8, 19, 65, 129, device="cuda"))
lab = torch.randint(0, 19, (8, 512, 1024), device="cuda"); lab[torch.rand(8, 512, 1024, device="cuda") < 0.05] = 255
g = torch.ones((), device="cuda")


def t(fn, it=20):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it


out, rows = ops.dsn_ce_fwd_train(l0, l1, lab, 255, 1.0, 0.4)
print("dsn_ce_fwd_train %.3f ms   dsn_ce_bwd_cols %.3f ms   loss %.6f" % (
    t(lambda: ops.dsn_ce_fwd_train(l0, l1, lab, 255, 1.0, 0.4)),
    t(lambda: ops.dsn_ce_bwd_cols(l0, l1, rows, 512, 1.0, 0.4, g, out[1])), float(out[0])))
