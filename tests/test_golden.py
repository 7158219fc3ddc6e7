"""CPU: the oracle restatement against its frozen outputs (tests/golden/oracle_v1.npz) and against
known answers worked out by hand from the reference's formulas."""
import os

import numpy as np
import torch

from oracle import grid_utils as og, render as orender
from tests.golden import make_golden

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_v1.npz"))


def test_oracle_reproduces_golden_vectors():
    now = make_golden.build()
    assert set(now.keys()) == set(GOLD.files)
    for k in GOLD.files:
        want, got = GOLD[k], now[k]
        assert got.shape == want.shape, k
        if np.issubdtype(want.dtype, np.integer):
            assert np.array_equal(got, want), k                     # bit-exact
        elif want.dtype == np.float64:
            np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-14, err_msg=k)
        else:
            np.testing.assert_allclose(got, want, rtol=2e-6, atol=1e-7, err_msg=k)


def test_hash_known_answers():
    """idx = (x ^ y*19349663 ^ z*83492791) mod T in uint32 (internal/grid_utils.py:101-111), corners in
    the order fff, ffc, fcf, fcc, cff, cfc, ccf, ccc of floor(loc - 0.5) (:61-77)."""
    T = 524288
    loc = np.array([[10.5, 20.5, 30.5], [0.5, 0.5, 0.5], [-0.5, 0.5, 0.5]], np.float32)  # floor -> (10,20,30), (0,0,0), (-1,0,0)
    got = og.hash_corner_indices_np(loc, T)

    def h(x, y, z):
        return ((x & 0xFFFFFFFF) ^ ((y * 19349663) & 0xFFFFFFFF) ^ ((z * 83492791) & 0xFFFFFFFF)) % T

    for row, (x, y, z) in zip(got, [(10, 20, 30), (0, 0, 0), (-1, 0, 0)]):
        want = [h(x + dx, y + dy, z + dz) for dx in (0, 1) for dy in (0, 1) for dz in (0, 1)]
        assert list(row) == want
    assert got[1, 0] == 0 and got[1, 4] == 1                      # h(0,0,0) = 0, h(1,0,0) = 1
    assert got[2, 0] == 0xFFFFFFFF % T                            # int32 -1 wraps to 2^32-1


def test_level_schedule_matches_configs():
    """grid sizes {16,...,Nmax}, ss = 1 (internal/grid_utils.py:772-794)."""
    assert list(og.grid_sizes(16, 512, 1.0)) == [16, 32, 64, 128, 256, 512]
    assert list(og.grid_sizes(16, 2048, 1.0)) == [16, 32, 64, 128, 256, 512, 1024, 2048]
    enc = og.HashEncoding(hash_map_size=524288, num_features=4, scale_supersample=1.0, max_grid_size=2048, bbox_scaling=1.0)
    assert [k for k, _, _ in enc.layout] == ["grid"] * 3 + ["hash"] * 5     # N^3 <= T  <=>  N <= 64


def test_alpha_weights_properties():
    """weights = alpha * trans; sum(weights) = 1 - prod(1 - alpha)  (internal/render.py:134-169)."""
    g = np.random.Generator(np.random.PCG64(5))
    dens = torch.from_numpy(g.gamma(0.5, 2.0, size=(32, 40)).astype(np.float32))
    tdist = torch.from_numpy(np.sort(g.uniform(2, 6, size=(32, 41)).astype(np.float32), -1))
    dirs = torch.from_numpy(g.normal(size=(32, 3)).astype(np.float32))
    w, a, t = orender.compute_alpha_weights(dens, tdist, dirs)
    assert torch.allclose(w, a * t)
    assert torch.allclose(w.sum(-1), 1 - torch.prod(1 - a.double(), -1).float(), atol=2e-6)
    assert float(t[:, 0].min()) == 1.0 and bool((t[:, 1:] <= t[:, :-1] + 1e-7).all())


# ------------------------------------------------------------------ SURVEY 8f-1 pieces (oracle_v2.npz)
def test_oracle_reproduces_golden_vectors_v2():
    from tests.golden import make_golden_v2
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_v2.npz"))
    now = make_golden_v2.build()
    assert set(now.keys()) == set(gold.files)
    for k in gold.files:
        np.testing.assert_allclose(now[k], gold[k], rtol=2e-6, atol=1e-7, err_msg=k)


def test_geometry_and_mask_loss_known_answers():
    """Worked by hand from internal/loss_utils.py:127-199 and internal/train_utils.py:785-836."""
    from oracle import loss_utils as oloss
    rays = {"viewdirs": torch.tensor([[0.0, 0.0, 1.0]])}             # v = -viewdirs = (0, 0, -1)
    n_pred = torch.tensor([[[0.0, 0.8, 0.6], [0.0, 0.0, -1.0]]])     # n.v = -0.6 (back-facing), +1 (front-facing)
    n = torch.tensor([[[0.0, 0.6, 0.8], [0.0, 0.0, -1.0]]])
    res = {"weights": torch.tensor([[0.5, 0.25]]), "normals": n, "normals_pred": n_pred}
    lo = oloss.orientation_loss(rays, res, target="normals_pred", mult=0.01)
    assert abs(float(lo) - 0.01 * (0.5 * 0.36 + 1e-5)) < 1e-9        # only the back-facing sample counts
    beta = torch.ones(1, 2, 1)
    lp = oloss.predicted_normal_loss(res, beta, mult=0.001, gt="normals_pred", pred="normals", stopgrad_weight=0.1)
    assert abs(float(lp) - 0.001 * (0.5 * (1 - 0.96) + 0.0 + 1e-5)) < 1e-9   # n.n_pred = 0.48 + 0.48, and 1
    acc = torch.tensor([1.0, 0.25])
    lm = oloss.compute_mask_loss(acc, None, 0.001, 1.0, 1.0)
    assert abs(float(lm) - 0.5 * (0.001 + np.sqrt(0.75**2 + 1e-6))) < 1e-7
    lb = oloss.compute_mask_loss(acc, torch.zeros(2, 1), 0.001, empty_loss_weight=0.1, backward=True)
    assert abs(float(lb) - 0.5 * 0.1 * (np.sqrt(1 + 1e-6) + np.sqrt(0.0625 + 1e-6))) < 1e-7
    # lossfun_distortion (stepfun.py:253-269): t = [0,1,3], w = [.5,.25]: u = [.5, 2] -> inter .375, intra .125
    ld = oloss.lossfun_distortion(torch.tensor([[0.0, 1.0, 3.0]]), torch.tensor([[0.5, 0.25]]))
    assert abs(float(ld) - 0.5) < 1e-7
    # stopgrad_with_weight: value unchanged, gradient scaled
    x = torch.tensor([2.0], requires_grad=True)
    y = oloss.stopgrad_with_weight(x, 0.1)
    y.backward()
    assert float(y) == 2.0 and abs(float(x.grad) - 0.1) < 1e-7
