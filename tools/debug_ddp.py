"""2-GPU debug (torchrun --nproc-per-node 2): which parameters differ between the plain and the bucket-overlapped gradient exchange."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from oracle import port as oport
from structure_knowledge_distillation_b200.networks.kd_model import NetModel
from structure_knowledge_distillation_b200.utils.train_options import make_args

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
torch.manual_seed(100 + rank)
m = NetModel(make_args(batch_size=1, pi=True, pa=True, ho=True, adv_loss_type="hinge", gpu_num=world))
images, labels = oport.synthetic_batch(1, 512, 512, seed=7 + rank)
m.set_input((images, labels, None, None))
for drop in m.student.dropouts():
    drop.injected = (torch.rand(1, 128, generator=torch.Generator().manual_seed(5 + rank)) >= 0.1).float()
buckets, m.G_solver._buckets = m.G_solver._buckets, None
d_state = {k: v.clone() for k, v in m.D_model.state_dict().items()}


def restore():
    with torch.no_grad():
        for k, v in m.D_model.state_dict().items():
            v.copy_(d_state[k])


results = {}
for mode in ("plain", "plain2", "overlap", "overlap_nostreams"):
    restore()
    m.G_solver._buckets = buckets if mode.startswith("overlap") else None
    m.overlap_streams = mode != "overlap_nostreams"
    m._student_phase()
    if not mode.startswith("overlap"):
        m.G_solver.all_reduce_grads(world)
    torch.cuda.synchronize()
    results[mode] = m.G_solver.flat_g.clone()
names = [n for n, p in m.student.named_parameters() if p.requires_grad]
offs = m.G_solver._offsets
if rank == 0:
    print("buckets:", [(b["lo"], b["hi"]) for b in buckets])
for mode in ("plain2", "overlap", "overlap_nostreams"):
    bad = []
    for i, n in enumerate(names):
        a, b = results[mode][offs[i]:offs[i + 1]], results["plain"][offs[i]:offs[i + 1]]
        if not torch.equal(a, b):
            bad.append((i, n, float((a - b).abs().max()), float(b.abs().max())))
    print("rank", rank, mode, "mismatching params:", len(bad), bad[:12])
dist.destroy_process_group()
