"""GPU debug: per-layer error of the discriminator's reverse pass against oracle/port.py in float64."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import port
from structure_knowledge_distillation_b200.networks.sagan_models import Discriminator


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def main():
    B, H, W = 2, 65, 129
    torch.manual_seed(11)
    Dp = port.Discriminator(1, 19, 64)
    with torch.no_grad():
        Dp.attn1.gamma.fill_(-0.4); Dp.attn2.gamma.fill_(0.25)
        Dp.preprocess_additional.weight.mul_(1.2)
    D = Discriminator(1, 19, B, 65, 64).cuda().train()
    D.load_state_dict({k: v.clone() for k, v in Dp.state_dict().items()})
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    xs = torch.randn(B, 19, H, W, device="cuda", generator=g) * 3
    for dtype in (torch.float64, torch.float32):
        Dq = port.Discriminator(1, 19, 64)
        Dq.load_state_dict(Dp.state_dict())
        Dq = Dq.cuda().to(dtype).train()
        zs = {}
        for i, l in enumerate((Dq.l1, Dq.l2, Dq.l3, Dq.l4)):
            def hook(mod, inp, out, i=i):
                out.retain_grad(); zs[i] = out
            l[0].register_forward_hook(hook)
        aq = {}
        for nm, A in (("a2", Dq.attn1), ("a3", Dq.attn2)):
            def pre(mod, inp, nm=nm):
                inp[0].retain_grad(); aq[nm + "in"] = inp[0]
            def post(mod, inp, out, nm=nm):
                out[0].retain_grad(); aq[nm + "out"] = out[0]
            A.register_forward_pre_hook(pre); A.register_forward_hook(post)
            for cn in ("query_conv", "key_conv", "value_conv"):
                def ph(mod, inp, out, key=nm + cn[0]):
                    out.retain_grad(); aq[key] = out
                getattr(A, cn).register_forward_hook(ph)
        torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
        xd = xs.to(dtype).requires_grad_(True)
        out = Dq(xd)[0]
        (-out.mean()).backward()
        if dtype == torch.float64:
            ref_z = {i: z.grad.clone() for i, z in zs.items()}; ref_dx = xd.grad.clone(); ref_out = out.detach().clone()
            ref_zv = {i: z.detach().clone() for i, z in zs.items()}
            ref_a = {k: v.grad.clone() for k, v in aq.items()}
        else:
            print("torch fp32 vs fp64: out %.2e dx %.2e" % (rel(out, ref_out), rel(xd.grad, ref_dx)),
                  {i: "%.2e" % rel(z.grad, ref_z[i]) for i, z in zs.items()})
    for precise in (True, False):
        D.load_state_dict({k: v.clone() for k, v in Dp.state_dict().items()})
        eng = D.engine
        eng.precise = precise
        t = eng.forward(xs)
        t.debug = []
        gout = torch.full_like(t.out, -1.0 / t.out.numel())
        dx = eng.backward(t, gout, {}, True, True)
        print("precise", precise, "out %.2e dx %.2e" % (rel(t.out, ref_out), rel(dx, ref_dx)))
        for i in range(4):
            print("   h%d fwd (post-leaky vs leaky(z_ref)) %.2e" % (i, rel(t.h[i][:B].permute(0, 3, 1, 2), torch.nn.functional.leaky_relu(ref_zv[i], 0.1))))
        for tag, gz in t.debug:
            i = int(tag[-1])
            if tag.startswith("gz"):
                print("   %s %.2e" % (tag, rel(gz[:B].permute(0, 3, 1, 2), ref_z[i])))
            elif tag.startswith("gy"):
                r = ref_a["a%dout" % i]
                print("   %s %.2e" % (tag, rel(gz.view(B, r.shape[2], r.shape[3], -1).permute(0, 3, 1, 2), r)))
            elif tag.startswith("gh"):
                r = ref_a["a%din" % i]
                print("   %s %.2e" % (tag, rel(gz.permute(0, 3, 1, 2), r)))
            elif tag.startswith("gqkv"):
                rq, rk, rv = ref_a["a%dq" % i], ref_a["a%dk" % i], ref_a["a%dv" % i]
                r = torch.cat([rq, rk, rv], 1).flatten(2).transpose(1, 2).reshape(gz.shape)
                d = rq.shape[1]
                print("   %s q %.2e k %.2e v %.2e" % (tag, rel(gz[:, :d], r[:, :d]), rel(gz[:, d:2 * d], r[:, d:2 * d]), rel(gz[:, 2 * d:], r[:, 2 * d:])))


if __name__ == "__main__":
    main()
