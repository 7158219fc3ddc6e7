"""CPU checks of the drop-in boundary: the shared library loads and exports every symbol that
include/nrc_b200.h declares; the ctypes prototypes cover exactly that set; argument
validation returns status codes (no compute is launched on this box)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "nrc_b200.h")


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nrc_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__

    __graft_entry__.build()
    from neural_radiance_caching_b200 import _lib

    return _lib.load()


def test_header_and_prototypes_agree(lib):
    from neural_radiance_caching_b200 import _lib

    syms = header_symbols()
    assert len(syms) >= 20
    assert sorted(_lib.PROTOTYPES.keys()) == syms


def test_library_exports_every_declared_symbol(lib):
    for s in header_symbols():
        assert hasattr(lib, s), s
    assert lib.nrc_abi_version() == 1


def test_header_compiles_as_c(tmp_path):
    import subprocess

    src = tmp_path / "t.c"
    src.write_text('#include "nrc_b200.h"\nint main(void){ nrc_encoding_t e; (void)e; return NRC_OK; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c",
                           str(src), "-o", str(tmp_path / "t.o")])


def test_struct_layouts_match_header(lib, tmp_path):
    """sizeof/offsetof of the ABI structs as seen by gcc == the ctypes mirrors."""
    import subprocess

    from neural_radiance_caching_b200 import _lib

    src = tmp_path / "s.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "nrc_b200.h"\n'
        "int main(void){ printf(\"%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n\", sizeof(nrc_level_t), sizeof(nrc_encoding_t),"
        " offsetof(nrc_encoding_t, levels), offsetof(nrc_encoding_t, bbox_span), sizeof(nrc_density_mlp_t),"
        " offsetof(nrc_density_mlp_t, in_dim), sizeof(nrc_density_mlp_grad_t), sizeof(nrc_slf_points_t),"
        " offsetof(nrc_slf_points_t, near), offsetof(nrc_slf_points_t, ref_warp_c)); return 0; }\n")
    exe = tmp_path / "s"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [ctypes.sizeof(_lib.nrc_level_t), ctypes.sizeof(_lib.nrc_encoding_t), _lib.nrc_encoding_t.levels.offset,
            _lib.nrc_encoding_t.bbox_span.offset, ctypes.sizeof(_lib.nrc_density_mlp_t),
            _lib.nrc_density_mlp_t.in_dim.offset, ctypes.sizeof(_lib.nrc_density_mlp_grad_t),
            ctypes.sizeof(_lib.nrc_slf_points_t), _lib.nrc_slf_points_t.near.offset, _lib.nrc_slf_points_t.ref_warp_c.offset]
    assert got == want


def test_error_convention(lib):
    assert lib.nrc_error_string(0) == b"ok"
    assert b"invalid" in lib.nrc_error_string(-1)
    # NULL descriptor / bad sizes are reported as status codes before any launch
    assert lib.nrc_encode_fwd(None, None, None, 16, None) == -1
    assert lib.nrc_ray_alpha_weights_fwd(None, None, None, None, 4, 0, 0, None, None, None) == -1
    assert lib.nrc_ray_alpha_weights_fwd(None, None, None, None, 4, 1000, 0, None, None, None) == -1
    # empty inputs are a successful no-op
    assert lib.nrc_ray_alpha_weights_fwd(None, None, None, None, 0, 32, 0, None, None, None) == 0
    assert lib.nrc_contract_fwd(None, None, 0, 2.0, None) == 0


def test_product_path_has_no_cpu_fallback():
    import torch

    from neural_radiance_caching_b200 import _lib, grid_utils

    enc = grid_utils.HashEncoding(hash_map_size=2**10, num_features=1, scale_supersample=1.0, max_grid_size=32,
                                  bbox_scaling=1.0)
    params = {n: torch.zeros(s) for (n, _, _, s) in enc.level_layout}
    with pytest.raises(_lib.NrcError):
        enc(params, torch.zeros(4, 3))  # CPU tensors: must raise, never compute


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "neural_radiance_caching_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_xla_custom_call_targets(lib, tmp_path):
    """include/nrc_xla.h: every declared custom-call target is exported and listed by nrc_xla_targets(); the descriptor
    structs the JAX side packs into `opaque` (neural_radiance_caching_b200/jax_binding/nrc_jax.py) have the layout gcc
    gives the header's; a descriptor of the wrong size is refused with a status, not a crash."""
    import subprocess

    from neural_radiance_caching_b200.jax_binding import nrc_jax

    src = open(os.path.join(ROOT, "include", "nrc_xla.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    declared = sorted(set(re.findall(r"\bvoid\s+(nrc_xla_[a-z0-9_]+)\s*\(", src)))
    assert declared == sorted(nrc_jax.TARGETS) and len(declared) == 19
    for name in declared:
        assert hasattr(lib, name), name

    class Target(ctypes.Structure):
        _fields_ = [("name", ctypes.c_char_p), ("fn", ctypes.c_void_p)]

    lib.nrc_xla_targets.restype = ctypes.POINTER(Target)
    table, listed = lib.nrc_xla_targets(), []
    while table[len(listed)].name:
        listed.append(table[len(listed)].name.decode())
    assert sorted(listed) == declared

    c = tmp_path / "x.c"
    c.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "nrc_xla.h"\n'
                 'int main(void){ printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(nrc_xla_encode_desc_t),'
                 ' offsetof(nrc_xla_encode_desc_t, enc), sizeof(nrc_xla_contract_desc_t), sizeof(nrc_xla_density_query_desc_t),'
                 ' offsetof(nrc_xla_density_query_desc_t, warp_c), sizeof(nrc_xla_ray_desc_t), offsetof(nrc_xla_ray_desc_t, bias),'
                 ' sizeof(nrc_xla_ggx_desc_t), offsetof(nrc_xla_ggx_desc_t, rgb_max), sizeof(nrc_xla_slf_desc_t),'
                 ' offsetof(nrc_xla_slf_desc_t, cfg)); return 0; }\n')
    exe = tmp_path / "x"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    J = nrc_jax
    want = [ctypes.sizeof(J.nrc_xla_encode_desc_t), J.nrc_xla_encode_desc_t.enc.offset, ctypes.sizeof(J.nrc_xla_contract_desc_t),
            ctypes.sizeof(J.nrc_xla_density_query_desc_t), J.nrc_xla_density_query_desc_t.warp_c.offset,
            ctypes.sizeof(J.nrc_xla_ray_desc_t), J.nrc_xla_ray_desc_t.bias.offset, ctypes.sizeof(J.nrc_xla_ggx_desc_t),
            J.nrc_xla_ggx_desc_t.rgb_max.offset, ctypes.sizeof(J.nrc_xla_slf_desc_t), J.nrc_xla_slf_desc_t.cfg.offset]
    assert got == want

    lib.nrc_xla_last_status.restype = ctypes.c_int32
    lib.nrc_xla_encode_fwd.restype = None
    lib.nrc_xla_encode_fwd(None, None, b"short", ctypes.c_size_t(5))
    assert lib.nrc_xla_last_status() == -1 and lib.nrc_xla_last_status() == 0
