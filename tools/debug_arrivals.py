"""1-GPU debug: which parameters report their gradient as complete, how often and through which path (bucketed all-reduce bookkeeping)."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import port as oport
from structure_knowledge_distillation_b200.networks.kd_model import NetModel
from structure_knowledge_distillation_b200.utils.train_options import make_args
torch.manual_seed(0)
m = NetModel(make_args(batch_size=1, pi=True, pa=True, ho=True, adv_loss_type="hinge"))
images, labels = oport.synthetic_batch(1, 256, 256, seed=7)
m.set_input((images, labels, None, None))
m.G_solver.enable_overlap(4)
names = [n for n, p in m.student.named_parameters() if p.requires_grad]
for it in range(2):
    m.G_solver._arrival_log = []
    m._student_phase()
    torch.cuda.synchronize()
    log = m.G_solver._arrival_log
    cnt = collections.Counter(i for i, _ in log)
    print("pass", it, "arrivals", len(log), "params", len(names))
    print("  never:", [names[i] for i in range(len(names)) if cnt[i] == 0][:20])
    print("  multiple:", [(names[i], cnt[i], [v for j, v in log if j == i]) for i in range(len(names)) if cnt[i] > 1][:20])
    print("  order (first 12):", [(names[i], v) for i, v in log[:12]])
    print("  order (last 12):", [(names[i], v) for i, v in log[-12:]])
