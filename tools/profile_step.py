"""Run W warm-up steps, then ONE distillation step inside cudaProfilerStart/Stop (for `ncu --profile-from-start off`)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from structure_knowledge_distillation_b200.networks.kd_model import NetModel
from structure_knowledge_distillation_b200.utils.train_options import make_args

warm = int(sys.argv[1]) if len(sys.argv) > 1 else 2
torch.manual_seed(0)
margs = make_args(batch_size=8, pi=True, pa=True, ho=True, adv_loss_type="wgan-gp")
model = NetModel(margs)
images, labels = bench.synthetic(8, 100)
model.set_input((images, labels, None, None))
for i in range(warm):
    model.optimize_parameters()
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
torch.cuda.profiler.start()
model.optimize_parameters()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled step wall ms:", (time.perf_counter() - t0) * 1e3)
