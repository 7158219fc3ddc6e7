"""One forward (and optionally data-gradient) launch of a shader stack on the chain kernel - the command profiled
with `ncu --set full --import-source on -k regex:chain` (profiles/*_ncu_chain_*): python tools/prof_chain.py slf 524288 fwd"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tests.test_chain_gpu as T
from neural_radiance_caching_b200 import mlp_chain as mc
from tests.util import gen, f32

name = sys.argv[1] if len(sys.argv) > 1 else "slf"
P = int(sys.argv[2]) if len(sys.argv) > 2 else 524288
mode = sys.argv[3] if len(sys.argv) > 3 else "fwd"
dev = torch.device("cuda:0")
g = gen(1)
spec = mc.ChainSpec(**T.SPECS[name])
p = {k: {a: b.to(dev) for a, b in v.items()} for k, v in T.make_params(g, spec).items()}
srcs = [f32(g.normal(size=(P, w))).to(dev) for w in spec.in_widths]
packed = mc.pack_weights(spec, p)
for _ in range(2):
    bufs, outs, act = mc.run_forward(spec, p, srcs, packed, save=(mode != "fwd"))
    if mode != "fwd":
        gh = [torch.randn_like(b) for b in bufs]
        d_src = [(torch.empty((P, w), device=dev), False) for w in spec.in_widths]
        mc.run_backward_data(spec, p, gh, act, packed, P, d_src)
torch.cuda.synchronize()
print("ok")
