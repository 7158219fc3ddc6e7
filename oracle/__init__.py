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
checked against XLA itself ("parity unpinned" w.r.t. XLA).  It IS pinned to
the reference's own *source*: tests/golden/make_golden.py imports the
reference's modules unmodified from /root/reference under a small NumPy
stand-in for `jax.numpy` and freezes their outputs into tests/golden/*.npz;
tests/test_oracle_golden.py checks this oracle against those vectors.
Assumed XLA semantics (IEEE-754 fp32 round-to-nearest per op, no FMA
contraction, saturating f32->s32 convert, uint32 wrap-around multiply) are
listed in DESIGN.md.
"""
import torch

# Deterministic fp32 CPU arithmetic.
torch.backends.cuda.matmul.allow_tf32 = False

F32 = torch.float32
