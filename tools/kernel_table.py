"""Per-kernel GPU time of one eager distillation step via torch.profiler (CUPTI, no replay) -- cheap alternative to an
ncu launch list for day-to-day optimisation (the committed evidence under profiles/ is ncu)."""
import os, sys, collections, re
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from structure_knowledge_distillation_b200.networks.kd_model import NetModel
from structure_knowledge_distillation_b200.utils.train_options import make_args
from torch.profiler import profile, ProfilerActivity

torch.manual_seed(0)
m = NetModel(make_args(batch_size=8, pi=True, pa=True, ho=True))
images, labels = bench.synthetic(8, 100)
m.set_input((images, labels, None, None))
for _ in range(3):
    m.optimize_parameters()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    m.optimize_parameters()
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA or getattr(ev, "device_time", 0) > 0:
        name = re.sub(r"\(.*", "", ev.name); name = re.sub(r"void |<unnamed>::|at::native::|\(anonymous namespace\)::", "", name)[:72]
        t = getattr(ev, "device_time", 0) or getattr(ev, "cuda_time", 0)
        if t > 0:
            agg[name][0] += 1; agg[name][1] += t
tot = sum(v[1] for v in agg.values())
print("total kernel time %.2f ms over %d launches" % (tot / 1e3, sum(v[0] for v in agg.values())))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print("%7.2f ms %5.1f%% %5d  %s" % (v[1] / 1e3, 100 * v[1] / tot, v[0], k))
