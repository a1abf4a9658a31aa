"""Host enqueue time vs GPU time per step, and the cost of the Ho (discriminator) part."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from structure_knowledge_distillation_b200.networks.kd_model import NetModel
from structure_knowledge_distillation_b200.utils.train_options import make_args

for ho in (True, False):
    torch.manual_seed(0)
    m = NetModel(make_args(batch_size=8, pi=True, pa=True, ho=ho))
    images, labels = bench.synthetic(8, 100)
    m.set_input((images, labels, None, None))
    for _ in range(3):
        m.optimize_parameters()
    torch.cuda.synchronize()
    n = 6
    t0 = time.perf_counter()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        m.optimize_parameters()
    t_host = (time.perf_counter() - t0) / n * 1e3
    e.record(); torch.cuda.synchronize()
    print("ho=%s  host enqueue %.1f ms/step   gpu %.1f ms/step" % (ho, t_host, s.elapsed_time(e) / n), flush=True)
    # phases
    def phase(fn, reps=4):
        torch.cuda.synchronize(); s.record()
        for _ in range(reps): fn()
        e.record(); torch.cuda.synchronize(); return s.elapsed_time(e) / reps
    def teacher():
        with torch.no_grad(): m.teacher(m.images)
    print("   teacher fwd %.1f ms" % phase(teacher))
    def fwd(): m.forward()
    print("   teacher+student fwd %.1f ms" % phase(fwd))
    def fb():
        m.forward(); m.G_solver.zero_grad(); m.student_backward()
    print("   fwd + losses + student bwd %.1f ms" % phase(fb))
    if ho:
        m.forward()
        print("   discriminator_backward only %.1f ms" % phase(m.discriminator_backward))
    del m
