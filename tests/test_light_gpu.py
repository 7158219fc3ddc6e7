"""Light sampler: nrc_vmf_head_{fwd,bwd}, nrc_vmf_loss and the LightMLP mirror against the oracle's
restatement of internal/light_sampler.py:135-214 and render_utils.vmf_loss_fn (render_utils.py:1493-1550)."""
import numpy as np
import pytest
import torch

from oracle import light_sampler as ols
from neural_radiance_caching_b200 import light_sampler as nls
from tests.util import f32, gen, rel_err

pytestmark = pytest.mark.gpu


def _unit(x):
    return x / np.linalg.norm(x, axis=-1, keepdims=True)


def _loss_inputs(g, P, K, S):
    means = f32(g.normal(size=(P, K, 3)) * 3.0)
    means[0, 0] = 0.0                                    # zero vector: l2_normalize's where() branch
    means[0, 1] = 1e-3                                   # |x|^2 < grad_eps: clamped backward denominator
    kappas = f32(g.uniform(0.0, 50.0, size=(P, K, 1)))
    kappas[0, 2] = 0.0                                   # kappa <= eps: uniform lobe
    logits = f32(g.normal(size=(P, K, 1)) - 4.0)
    normals = f32(_unit(g.normal(size=(P, 3))))
    dirs = f32(_unit(g.normal(size=(P, S, 3))))
    pdf = f32(g.uniform(0.0, 2.0, size=(P, S)))
    weight = f32(g.uniform(-1.0, 12.0, size=(P, S)))
    rad = f32(g.uniform(0.0, 1.5, size=(P, S, 3)))
    rad[0, 0] = 0.0                                      # below the 1e-5 clamp
    return means, kappas, logits, normals, dirs, pdf, weight, rad


@pytest.mark.parametrize("P,K,S,srgb", [(9, 128, 32, True), (5, 128, 40, False), (3, 16, 7, True)])
def test_vmf_loss_matches_oracle(cuda_device, P, K, S, srgb):
    g = gen(1600 + S)
    means, kappas, logits, normals, dirs, pdf, weight, rad = _loss_inputs(g, P, K, S)
    def run(lib, dev, dt=torch.float32):
        M, Kp, Lg = (t.clone().to(dev, dt).requires_grad_(True) for t in (means, kappas, logits))
        c = lambda t: t.to(dev, dt)
        vmfs = dict(vmf_means=M, vmf_kappas=Kp, vmf_logits=Lg, vmf_normals=c(normals)[:, None, :])
        l = lib.light_sampling_loss(vmfs, c(dirs), c(pdf), c(weight), c(rad), srgb)
        l.backward()
        return l.detach().cpu().float(), M.grad.cpu().float(), Kp.grad.cpu().float(), Lg.grad.cpu().float()
    # the oracle in float64 is the truth: kappa / sinh(kappa) * exp(kappa t) and coth(kappa) - 1/kappa both cancel
    # badly in fp32 (the fp32 oracle itself is off by 5e-4 on the kappa gradient)
    lo, gmo, gko, glo = run(ols, "cpu", torch.float64)
    ln, gmn, gkn, gln = run(nls, cuda_device)
    assert abs(float(ln) - float(lo)) <= 1e-4 * abs(float(lo))
    assert rel_err(gln, glo) <= 1e-4
    # at kappa == 0 the where() of eval_vmf differentiates 0/0 in its untaken branch: NaN in the oracle (and in
    # JAX); the kernel writes 0.  kappa = min(softplus(.), 50) never is 0 in the model.
    assert torch.isnan(gko[0, 2]).all() and float(gkn[0, 2].abs().max()) == 0.0
    gko[0, 2] = 0.0
    gmo[0, 2] = torch.nan_to_num(gmo[0, 2])
    assert rel_err(gkn, gko) <= 1e-4
    assert rel_err(gmn, gmo) <= 1e-4


@pytest.mark.parametrize("per_point", [True, False])
def test_vmf_head_matches_oracle(cuda_device, per_point):
    g = gen(1700)
    P, K = 37, 128
    raw = f32(g.normal(size=(P, K * 5)) * 3.0)
    raw[0, 3] = 80.0            # softplus saturates at 50
    raw[0, 4] = -60.0           # logit clamps at -50
    mr = f32(g.normal(size=(P, K, 3) if per_point else (K, 3)) * 10.0)
    pos = f32(g.uniform(-2, 2, size=(P, 3)))
    G = [f32(g.normal(size=s)) for s in ((P, K, 3), (P, K, 1), (P, K, 1))]
    Ro = raw.clone().requires_grad_(True)
    vo = ols.get_vmfs(Ro.reshape(P, K, 5), mr)
    mo = vo["vmf_means"] - pos[:, None, :]
    (mo * G[0]).sum().backward(retain_graph=True)
    (vo["vmf_kappas"] * G[1]).sum().backward(retain_graph=True)
    (vo["vmf_logits"] * G[2]).sum().backward()
    Rn = raw.clone().to(cuda_device).requires_grad_(True)
    mn, kn, ln = nls._VmfHeadFn.apply(Rn, mr.to(cuda_device), pos.to(cuda_device), K, 20.0)
    ((mn * G[0].to(cuda_device)).sum() + (kn * G[1].to(cuda_device)).sum() + (ln * G[2].to(cuda_device)).sum()).backward()
    assert rel_err(mn, mo.detach()) <= 1e-6
    assert rel_err(kn, vo["vmf_kappas"].detach()) <= 1e-6
    assert rel_err(ln, vo["vmf_logits"].detach()) <= 1e-6
    assert rel_err(Rn.grad, Ro.grad) <= 1e-5


@pytest.mark.parametrize("bf16", [False, True])
def test_light_mlp_matches_oracle(cuda_device, bf16):
    """predict_lighting end to end (grid -> two ReLU layers -> output layer -> head) and the gradient of the
    light-sampling loss w.r.t. the output layer and the light grid."""
    g = gen(1800)
    P, S = 300, 8
    o, n = ols.LightMLP(), nls.LightMLP(bf16=bf16)
    po = o.init(g, table_init_range=0.3)
    pn = n.from_oracle(po, cuda_device)
    K = o.num_components
    means = f32(g.uniform(-2.5, 2.5, size=(P, 3)))
    mr = f32(g.normal(size=(K, 3)) * 10.0)
    normals = f32(_unit(g.normal(size=(P, 3))))
    dirs = f32(_unit(g.normal(size=(P, S, 3))))
    pdf, weight = f32(g.uniform(0.05, 2.0, size=(P, S))), f32(g.uniform(0.0, 5.0, size=(P, S)))
    rad = f32(g.uniform(0.0, 1.5, size=(P, S, 3)))
    for v in po.values():
        for t in v.values():
            t.requires_grad_(True)
    vo = o.predict_lighting(po, means, mr, normals=normals)
    lo = ols.light_sampling_loss(vo, dirs, pdf, weight, rad)
    lo.backward()
    arena = pn["light_grid"]["_arena"].requires_grad_(True)
    for k in ("layers_0", "layers_1", "output_layer"):
        for kk in pn[k]:
            pn[k][kk].requires_grad_(True)
    d = lambda t: t.to(cuda_device)
    vn = n.predict_lighting(pn, d(means), d(mr), normals=d(normals))
    ln = nls.light_sampling_loss(vn, d(dirs), d(pdf), d(weight), d(rad))
    ln.backward()
    tol = 2e-2 if bf16 else 1e-4        # bf16 operands on the tensor cores (north-star bf16-MLP tolerance) / fp32 parity
    for k in ("vmf_means", "vmf_kappas", "vmf_logits"):
        assert rel_err(vn[k], vo[k].detach()) <= tol, k
    assert abs(float(ln.detach()) - float(lo.detach())) <= (5e-2 if bf16 else 1e-4) * abs(float(lo.detach()))
    if not bf16:
        for k in ("layers_0", "layers_1", "output_layer"):
            assert rel_err(pn[k]["kernel"].grad, po[k]["kernel"].grad) <= 2e-4, k
            assert rel_err(pn[k]["bias"].grad, po[k]["bias"].grad) <= 2e-4, k
        names = [nm for (nm, _, _, _) in n.grid.level_layout]
        want = torch.cat([po["light_grid"][nm].grad.reshape(-1) for nm in names])
        assert rel_err(arena.grad, want) <= 2e-4
    else:
        from tests.util import rel_l2
        assert rel_l2(pn["output_layer"]["kernel"].grad, po["output_layer"]["kernel"].grad.to(cuda_device)) <= 5e-2


def test_light_mlp_render_path_chain(cuda_device):
    """no-grad bf16 path (one tcgen05 chain program, five head groups) against the fp32 oracle."""
    g = gen(1810)
    P = 1000
    o, n = ols.LightMLP(), nls.LightMLP(bf16=True)
    po = o.init(g, table_init_range=0.3)
    pn = n.from_oracle(po, cuda_device)
    means = f32(g.uniform(-2.5, 2.5, size=(P, 3)))
    mr = f32(g.normal(size=(o.num_components, 3)) * 10.0)
    with torch.no_grad():
        vo = o.predict_lighting(po, means, mr)
        vn = n.predict_lighting(pn, means.to(cuda_device), mr.to(cuda_device))
        vn2 = n.predict_lighting(pn, means.to(cuda_device), mr.to(cuda_device))     # packed weights from the cache
    for k in ("vmf_means", "vmf_kappas", "vmf_logits"):
        assert rel_err(vn[k], vo[k]) <= 2e-2, k
        assert torch.equal(vn[k], vn2[k])
