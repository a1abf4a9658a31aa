import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from structure_knowledge_distillation_b200 import ops
from structure_knowledge_distillation_b200.networks.kd_model import NetModel
from structure_knowledge_distillation_b200.utils.train_options import make_args
torch.manual_seed(0)
m = NetModel(make_args(batch_size=8, pi=True, pa=True, ho=False))
images, labels = bench.synthetic(8, 100)
m.set_input((images, labels, None, None))
img4 = ops.pad_channels(m.images, 4)
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def phase(fn, reps=4):
    fn(); torch.cuda.synchronize(); s.record()
    for _ in range(reps): fn()
    e.record(); torch.cuda.synchronize(); return s.elapsed_time(e) / reps
for on in (False, True, False, True):
    m.student.set_precise_early_layers(on)
    def sf():
        with torch.no_grad():
            m.student.train()(img4)
    def sfb():
        m.G_solver.zero_grad()
        out = m.student.train()(img4)
        (out[0].sum() + out[1].sum()).backward()
    print("precise=%s  student fwd (no autograd) %.2f ms   fwd+bwd %.2f ms" % (on, phase(sf), phase(sfb)), flush=True)
# single layer
x = ops.to_nhwc(torch.randn(8, 64, 256, 512, device="cuda")); w = torch.randn(64, 3, 3, 64, device="cuda") / 24
print("conv2 tf32 %.3f ms | 3x %.3f ms | lo_tf32(x) %.3f ms" % (phase(lambda: ops.conv2d_fwd(x, w, 1, 1, 1)), phase(lambda: ops.conv2d_fwd_3xtf32(x, w, 1, 1, 1)), phase(lambda: ops.lo_tf32(x))))
