"""GPU debug: the discriminator phase of a step case against oracle/port.py in float64 ON THE SAME LOGITS (isolates the D phase from the
student's TF32 logit perturbation)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import cases, port
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import test_step_gpu as T


def main(name="pi_pa_ho_hinge_512"):
    m, cfg = T._build(name)
    m.forward(); m.G_solver.zero_grad(); m.student_backward(); m.G_solver.step()
    sd = {k: v.detach().clone() for k, v in m.D_model.state_dict().items()}
    ls, lt = m.preds_S[0].detach().double().contiguous(), m.preds_T[0].detach().double().contiguous()
    m._discriminator_phase()
    Dq = port.Discriminator(1, 19, 64)
    Dq.load_state_dict({k: v.cpu() for k, v in sd.items()})
    Dq = Dq.cuda().double().train()
    torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
    dT, dS = Dq(lt), Dq(ls)
    dl = cfg.lambda_d * port.adv_loss_d(dS, dT, cfg.adv_type)
    if cfg.adv_type == "wgan-gp":
        dl = dl + cfg.lambda_d * port.gradient_penalty(Dq, ls, lt, m.criterion_AdditionalGP.alpha.double(), cfg.lambda_gp)
    dl.backward()
    print(name, "D loss ours %.8f port-on-our-logits %.8f" % (float(m.D_loss), float(dl)))
    refs = dict(Dq.named_parameters())
    for n, p in m.D_model.named_parameters():
        q = refs.get(n)
        if q is None or q.grad is None or p.grad is None or float(q.grad.norm()) == 0:
            continue
        a, b = p.grad.detach().double().flatten(), q.grad.flatten()
        print("   %-34s rel-L2 %.2e   norm ours %.4e ref %.4e" % (n, float((a - b).norm() / b.norm()), float(a.norm()), float(b.norm())))


if __name__ == "__main__":
    for nm in sys.argv[1:] or ["pi_pa_ho_hinge_512"]:
        main(nm)
