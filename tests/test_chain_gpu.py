"""tcgen05 MLP chains (csrc/chain.cu, mlp_chain.py) against a PyTorch fp32 reference of the same
Dense stacks.  Two references: (a) plain fp32 (the reference's arithmetic; bf16 tolerance 2e-2 from
BASELINE.json north_star), (b) the same stack with operands rounded to bf16 at exactly the points the
kernel rounds them (tight tolerance: proves the GEMMs, layouts, masks and reductions are exact)."""
import numpy as np
import pytest
import torch

from neural_radiance_caching_b200 import mlp_chain as mc
from tests.util import f32, gen, rel_err

pytestmark = pytest.mark.gpu


def bf(x):
    return x.to(torch.bfloat16).to(torch.float32)


def he(g, fi, fo):
    lim = np.sqrt(6.0 / fi)
    return f32(g.uniform(-lim, lim, size=(fi, fo)))


def make_params(g, spec):
    p, parts = {}, None
    for li, (name, w, _) in enumerate(spec.hidden):
        fi = sum(x[1] for x in spec.x_parts[li])
        p[name] = {"kernel": he(g, fi, w), "bias": f32(g.normal(size=(w,)) * 0.1)}
    fi = sum(x[1] for x in spec.x_last)
    for grp in spec.heads:
        for name, w in grp:
            p[name] = {"kernel": he(g, fi, w), "bias": f32(g.normal(size=(w,)) * 0.1)}
    return p


def ref_stack(spec, p, sources, rounded):
    """Differentiable torch reference.  rounded=True mirrors the kernel's bf16 rounding points with
    straight-through gradients (the kernel's VJP uses the same rounded operands)."""
    class Round(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x):
            return bf(x)

        @staticmethod
        def backward(ctx, g):
            return bf(g)     # the data-gradient chain stores dY as bf16 too

    rnd = (lambda x: Round.apply(x)) if rounded else (lambda x: x)
    wq = (lambda w: bf(w.detach()) + (w - w.detach())) if rounded else (lambda w: w)
    cat = torch.cat(sources, dim=-1)
    inputs = (bf(cat.detach()) + (cat - cat.detach())) if rounded else cat   # d_in stays fp32 in the kernel
    x = inputs
    for (name, w, skip), act in zip(spec.hidden, spec.hidden_act):
        z = x @ wq(p[name]["kernel"]) + p[name]["bias"]
        x = rnd(torch.relu(z) if act == "relu" else z)
        if skip:
            x = torch.cat([x, inputs], dim=-1)
    outs = []
    for grp in spec.heads:
        for name, w in grp:
            outs.append(x @ wq(p[name]["kernel"]) + p[name]["bias"])
    return outs


SPECS = {
    "trunk": dict(in_widths=[64, 32], hidden=[], heads=[[("bottleneck", 128)], [("rough", 1), ("amb", 3), ("irr", 3),
                                                                                 ("tint", 3)]]),
    "int_brdf": dict(in_widths=[128, 1], hidden=[("l0", 64, False), ("l1", 64, False)], heads=[[("out", 1)]]),
    "slf": dict(in_widths=[128, 72], hidden=[("l0", 128, False), ("l1", 128, False), ("l2", 128, True), ("lb", 128, False)],
                heads=[[("rgb", 3)]]),
    "material": dict(in_widths=[32], hidden=[("bottleneck", 128, False, "linear")], heads=[[("pred_brdf", 10)]]),
    "env": dict(in_widths=[38], hidden=[("l0", 128, False), ("l1", 128, False), ("l2", 128, True), ("lb", 128, False)],
                heads=[[("rgb", 3)]]),
}


@pytest.mark.parametrize("name", list(SPECS))
@pytest.mark.parametrize("P", [1000, 4096])
def test_chain_forward_backward(cuda_device, name, P):
    g = gen(500 + P + len(name))
    spec = mc.ChainSpec(**SPECS[name])
    p = make_params(g, spec)
    srcs = [f32(g.normal(size=(P, w))) for w in spec.in_widths]
    n_out = sum(len(grp) for grp in spec.heads)
    # references
    def run_ref(rounded):
        po = {k: {a: b.clone().requires_grad_(True) for a, b in v.items()} for k, v in p.items()}
        so = [s.clone().requires_grad_(True) for s in srcs]
        outs = ref_stack(spec, po, so, rounded)
        return po, so, outs

    gouts = [bf(f32(g.normal(size=(P, w)))) for grp in spec.heads for _, w in grp]   # dY of the heads is bf16
    po, so, oo = run_ref(True)
    sum((o * go).sum() for o, go in zip(oo, gouts)).backward()
    p32, s32, o32 = run_ref(False)
    # kernels
    pn = {k: {a: b.to(cuda_device).requires_grad_(True) for a, b in v.items()} for k, v in p.items()}
    sn = [s.to(cuda_device).requires_grad_(True) for s in srcs]
    on = mc.apply(spec, pn, sn)
    assert len(on) == n_out
    sum((o * go.to(cuda_device)).sum() for o, go in zip(on, gouts)).backward()
    torch.cuda.synchronize()
    for a, b, c in zip(on, oo, o32):
        assert a.shape == b.shape
        assert rel_err(a, b) <= 5e-3, "forward vs bf16-rounded reference"
        assert rel_err(a, c) <= 2e-2, "forward vs fp32 reference (north_star bf16 tolerance)"
    # A hidden pre-activation within an ulp of zero (or of a bf16 rounding boundary) can land on the
    # other side under a different fp32 summation order: a handful of ROWS then differ by O(1e-2).
    # Everything else must agree to accumulation-order noise.
    for a, b in zip(sn, so):
        row_err = (a.grad.cpu() - b.grad).abs().max(dim=1).values / b.grad.abs().max()
        assert float((row_err > 2e-3).float().mean()) <= 1e-3, "input gradient"
        assert float(row_err.max()) <= 5e-2
    for k in p:
        assert rel_err(pn[k]["kernel"].grad, po[k]["kernel"].grad) <= 1e-2, f"kernel gradient {k}"
        assert rel_err(pn[k]["bias"].grad, po[k]["bias"].grad) <= 1e-2, f"bias gradient {k}"


def test_chain_rejects_bad_specs():
    with pytest.raises(ValueError):
        mc.ChainSpec(in_widths=[3, 5], hidden=[], heads=[[("a", 3)]])
    with pytest.raises(ValueError):
        mc.ChainSpec(in_widths=[8], hidden=[("l0", 512, False)], heads=[[("a", 3)]])
    with pytest.raises(ValueError):
        mc.ChainSpec(in_widths=[8], hidden=[], heads=[[("a", 100), ("b", 100)]])


def test_chain_256_wide_forward(cuda_device):
    """Model-level EnvMap shape (configs/nerf_ngp_yobo.gin:253-297): pos_enc(dir, 4) = 27 -> 4 x 256 with the
    input re-concatenated after layer 2 -> rgba (4) + ambient (3).  Forward-only on the chain kernel."""
    g = gen(560)
    spec = mc.ChainSpec(in_widths=[27], hidden=[("l0", 256, False), ("l1", 256, False), ("l2", 256, True), ("lb", 256, False)],
                        heads=[[("rgba", 4), ("amb", 3)]])
    assert not spec.supports_backward
    p = make_params(g, spec)
    P = 3000
    srcs = [f32(g.normal(size=(P, 27)))]
    want = ref_stack(spec, p, srcs, True)
    want32 = ref_stack(spec, p, srcs, False)
    pn = {k: {a: b.to(cuda_device) for a, b in v.items()} for k, v in p.items()}
    with torch.no_grad():
        got = mc.apply(spec, pn, [s.to(cuda_device) for s in srcs])
    for a, b, c in zip(got, want, want32):
        assert rel_err(a, b) <= 5e-3
        assert rel_err(a, c) <= 2e-2
