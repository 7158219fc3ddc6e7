"""CPU oracle for the radiance-cache query path.  TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (NumPy for the integer/index arithmetic,
PyTorch-CPU fp32 for the floating point arithmetic and its autograd) of the
reference's JAX code for the hot path named in BASELINE.json.  Every function
cites the reference file:line it follows (paths relative to /root/reference).

Rules (see DESIGN.md "Oracle"):
  * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
    --impl reference leg may import this package, and only as the checker.
  * The product package (neural_radiance_caching_b200) never imports it and
    has no CPU fallback: it raises when the CUDA library is missing.

Pinning status: the reference ships no tests, golden vectors or fixtures for
this path and JAX/XLA is not installable here, so the restatement cannot be
checked against XLA itself ("parity unpinned" w.r.t. XLA's own rounding).  It IS
pinned to the reference's own *source*: tests/golden/make_reference_vectors.py
imports the reference's modules unmodified from /root/reference under a small
NumPy stand-in for jax (tests/golden/jax_numpy_shim.py), runs their functions
(internal/math.py, coord.py, stepfun.py, render.py incl. the transient renderer,
grid_utils.py incl. the HashEncoding class, ref_utils.py, image.py,
loss_utils.py, train_utils.compute_mask_loss, Model.maybe_resample,
inverse_render/render_utils.py: GGX lobe + integration, samplers,
importance_sample_rays, vMF pdf / loss, zero_invalid_bins)
on seeded float32 inputs and freezes inputs and outputs into
tests/golden/reference_np.npz; tests/test_reference_vectors.py checks this
oracle against those vectors (corner indices / interpolation / contraction /
cast_rays / l2_normalize bit-exact, the rest to fp32 rounding) and
tests/test_reference_vectors_gpu.py checks the CUDA kernels against them
directly.  Classes whose reference counterpart needs flax parameter scoping (the
MLP classes in geometry.py, nerf.py, material.py, light_sampler.py and the
sampler loop in sampling.py) are pinned only through these building blocks and
say so in their own headers.  The oracle's own outputs are additionally frozen in
tests/golden/oracle_v*.npz (make_golden*.py) so that any later edit of the
restatement is caught.  Assumed XLA semantics (IEEE-754 fp32 round-to-nearest
per op, no FMA contraction, saturating f32->s32 convert, uint32 wrap-around
multiply) are listed in DESIGN.md.
"""
import torch

# Deterministic fp32 CPU arithmetic.
torch.backends.cuda.matmul.allow_tf32 = False

F32 = torch.float32
