"""nrc_chain_query (hash-grid gather feeding the tcgen05 density stack, render path) against the mma.sync
fused query kernel, nrc_encode_fwd and the oracle (internal/geometry.py:199-341)."""
import pytest
import torch

from oracle import geometry as ogeo
from neural_radiance_caching_b200 import geometry as ngeo, mlp_chain
from tests.util import f32, gen

pytestmark = pytest.mark.gpu

GRIDS = [
    dict(hash_map_size=524288, max_grid_size=512, num_features=1),     # 6 levels
    dict(hash_map_size=524288, max_grid_size=1024, num_features=1),    # 7 levels
    dict(hash_map_size=524288, max_grid_size=2048, num_features=4),    # 8 levels x 4 features
    dict(hash_map_size=4096, max_grid_size=256, num_features=2),
]


def _pair(g, grid, pred, device, bf16):
    kw = dict(grid_params=grid, enable_pred_normals=pred)
    o = ogeo.DensityMLP(bf16=bf16, **kw)
    n = ngeo.DensityMLP(bf16=True, **kw)
    po = o.init(g, table_init_range=0.5, bias_range=0.1)
    return o, n, po, n.from_oracle(po, device)


def _scaled_err(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-6))


@pytest.mark.parametrize("gi,pred,P", [(0, False, 1000), (1, True, 128 * 301 + 5), (2, True, 128 * 700 + 77), (3, True, 3)])
def test_query_tc_matches_fused_query(cuda_device, gi, pred, P):
    g = gen(40 + gi)
    _, n, _, pn = _pair(g, GRIDS[gi], pred, cuda_device, True)
    if not n.supports_query_tc():
        pytest.skip("more than 32 encoded features")
    means = f32(g.uniform(-3, 3, size=(P, 3))).to(cuda_device)
    ref = n.query(pn, means, want_feat=True)
    density = torch.full((P,), -7.0, device=cuda_device)
    feat = torch.full((P, 64), -7.0, device=cuda_device)
    gp = torch.full((P, 3), -7.0, device=cuda_device) if pred else None
    enc_out = torch.full((P, n.in_dim), -7.0, device=cuda_device)
    cache = mlp_chain.PackCache()
    for _ in range(2):   # the second call takes the packed weights from the cache
        n.query_tc(pn, means, density, feat, gp, enc_out, cache=cache)
    torch.cuda.synchronize()
    # Both kernels round operands to bf16 and accumulate in fp32; only the accumulation order differs.
    # Tolerance: 2e-2 of the output scale (the bound north_star states for the bf16 variant).
    assert _scaled_err(feat, ref["feature"]) < 2e-2
    assert torch.equal(density == 0, ref["density"] == 0)      # bbox mask: exact
    m = ref["density"] > 0
    if m.any():
        assert float(((density - ref["density"]).abs() / ref["density"])[m].max()) < 5e-2
    if pred:
        assert _scaled_err(gp, ref["grad_pred"]) < 2e-2
    # the gathered features are the arithmetic of nrc_encode_fwd: bit-exact
    from neural_radiance_caching_b200 import coord
    enc_ref = n.grid(pn["density_grid"], coord._ContractFn.apply(means, n.warp_c))
    assert torch.equal(enc_out, enc_ref.reshape(P, -1))


def test_query_tc_matches_oracle(cuda_device):
    """Against the oracle with bf16-rounded operands (oracle.geometry.dense_bf16): 2e-2 of the output scale."""
    g = gen(77)
    o, n, po, pn = _pair(g, GRIDS[2], True, cuda_device, True)
    P = 4099
    means = f32(g.uniform(-2.5, 2.5, size=(P, 3)))
    with torch.no_grad():
        raw, feat_o = o.predict_density(po, means)
        dens_o = o.convert_raw_density(raw, means)
        gp_o = o.dense(po["pred_normals_layer"], feat_o)
    density = torch.empty((P,), device=cuda_device)
    feat = torch.empty((P, 64), device=cuda_device)
    gp = torch.empty((P, 3), device=cuda_device)
    n.query_tc(pn, means.to(cuda_device), density, feat, gp)
    torch.cuda.synchronize()
    assert _scaled_err(feat.cpu(), feat_o) < 2e-2
    assert _scaled_err(gp.cpu(), gp_o) < 2e-2
    assert torch.equal(density.cpu() == 0, dens_o == 0)
    m = dens_o > 0
    assert float(((density.cpu() - dens_o).abs() / dens_o)[m].max()) < 5e-2


@pytest.mark.parametrize("secondary", [False, True])
def test_query_tc_in_render_schedule(cuda_device, secondary):
    """FusedCacheQuery with and without the tensor-core density query renders the same rays."""
    import numpy as np
    from neural_radiance_caching_b200 import engine, workload
    step = workload.CacheTrainStep(cuda_device, bf16=True)
    R = 700
    g = np.random.Generator(np.random.PCG64(workload.SEED + 21))
    rn = workload.make_rays_np(g, R, near=0.05, far=2.0, radius=0.7) if secondary else workload.make_rays_np(g, R)
    rays = {k: torch.from_numpy(v).to(cuda_device) for k, v in rn.items()}
    u01 = [torch.from_numpy(g.uniform(size=(R, 1)).astype(np.float32)).to(cuda_device) for _ in range(3)]
    outs = []
    for flag in (True, False):
        q = engine.FusedCacheQuery(step.model)
        q.tensor_core_query = flag
        with torch.no_grad():
            outs.append(q(step.params, rays, u01, is_secondary=secondary))
    torch.cuda.synchronize()
    # proposal densities steer the resampling: sample positions move slightly between the two bf16 paths
    assert _scaled_err(outs[0]["rgb"], outs[1]["rgb"]) < 5e-2
    assert _scaled_err(outs[0]["acc"], outs[1]["acc"]) < 5e-2
