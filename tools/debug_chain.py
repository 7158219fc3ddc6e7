"""Error statistics of the tcgen05 chains vs the bf16-rounded torch reference (GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tests.test_chain_gpu as T
from neural_radiance_caching_b200 import mlp_chain as mc
from tests.util import gen, f32

dev = torch.device("cuda:0")
for name in T.SPECS:
    for P in (1000, 4096, 32768):
        g = gen(500 + P + len(name))
        spec = mc.ChainSpec(**T.SPECS[name])
        p = T.make_params(g, spec)
        srcs = [f32(g.normal(size=(P, w))) for w in spec.in_widths]
        gouts = [T.bf(f32(g.normal(size=(P, w)))) for grp in spec.heads for _, w in grp]
        po = {k: {a: b.clone().to(dev).requires_grad_(True) for a, b in v.items()} for k, v in p.items()}
        so = [s.clone().to(dev).requires_grad_(True) for s in srcs]
        oo = T.ref_stack(spec, po, so, True)
        sum((o * go.to(dev)).sum() for o, go in zip(oo, gouts)).backward()
        pn = {k: {a: b.to(dev).requires_grad_(True) for a, b in v.items()} for k, v in p.items()}
        sn = [s.to(dev).requires_grad_(True) for s in srcs]
        on = mc.apply(spec, pn, sn)
        sum((o * go.to(dev)).sum() for o, go in zip(on, gouts)).backward()
        torch.cuda.synchronize()
        msg = [f"{name} P={P}"]
        for a, b in zip(on, oo):
            msg.append(f"fwd {float((a-b).abs().max()/b.abs().max()):.1e}")
        for a, b in zip(sn, so):
            e = (a.grad - b.grad).abs().max(dim=1).values / b.grad.abs().max()
            msg.append(f"dx max {float(e.max()):.1e} rows>1e-3: {int((e > 1e-3).sum())}")
        for k in p:
            for w in ("kernel", "bias"):
                a, b = pn[k][w].grad, po[k][w].grad
                msg.append(f"{k}.{w[0]} {float((a-b).abs().max()/b.abs().max()):.1e}")
        print(" | ".join(msg), flush=True)
