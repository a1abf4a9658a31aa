"""Teacher forward time vs. the conv tile walk order (skd_set_conv_tile_order): 0 front-to-back, 2 alternate per launch."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from structure_knowledge_distillation_b200 import ops, _cabi
from structure_knowledge_distillation_b200.networks.kd_model import NetModel
from structure_knowledge_distillation_b200.utils.train_options import make_args
torch.manual_seed(0)
m = NetModel(make_args(batch_size=8, pi=True, pa=True, ho=False))
images, labels = bench.synthetic(8, 100)
m.set_input((images, labels, None, None))
img4 = ops.pad_channels(m.images, 4)
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
def phase(fn, reps=5):
    fn(); torch.cuda.synchronize(); s.record()
    for _ in range(reps): fn()
    e.record(); torch.cuda.synchronize(); return s.elapsed_time(e) / reps
def tf():
    with torch.no_grad():
        m.teacher(img4)
ref = None
for mode in (0, 2, 1, 0, 2):
    _cabi.lib().skd_set_conv_tile_order(mode)
    t = phase(tf)
    with torch.no_grad():
        out = m.teacher(img4)[0].float().clone()
    if ref is None: ref = out
    print("tile order %d: teacher fwd %.2f ms   max|diff vs order 0| %.3g" % (mode, t, (out - ref).abs().max().item()), flush=True)
