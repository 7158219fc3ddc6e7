"""ctypes binding of include/nrc_b200.h (the C-ABI drop-in boundary).

There is NO CPU fallback: importing the package works anywhere (so that the
build check and host-logic tests run on a CPU box), but every compute call
raises unless libnrc_b200.so is built in-tree and the tensors live on a CUDA
device.
"""
import ctypes as C
import os

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NRC_LIB_PATH") or os.path.join(HERE, "libnrc_b200.so")
NRC_MAX_LEVELS = 16


class NrcError(RuntimeError):
    pass


class nrc_level_t(C.Structure):
    _fields_ = [
        ("d_table", C.c_void_p),
        ("d_grad", C.c_void_p),
        ("grid_size", C.c_int32),
        ("is_hash", C.c_int32),
        ("table_size", C.c_uint32),
        ("reserved", C.c_uint32),
    ]


class nrc_encoding_t(C.Structure):
    _fields_ = [
        ("num_levels", C.c_int32),
        ("num_features", C.c_int32),
        ("bbox_min", C.c_float * 3),
        ("bbox_max", C.c_float * 3),
        ("bbox_span", C.c_float * 3),
        ("precondition_scaling", C.c_float),
        ("levels", nrc_level_t * NRC_MAX_LEVELS),
    ]


class nrc_shader_images_t(C.Structure):
    _fields_ = [("slf_img", C.c_void_p), ("slf_img_atoms", C.c_int32), ("slf_atom0", C.c_int32),
                ("env_img", C.c_void_p), ("env_img_atoms", C.c_int32), ("env_atom0", C.c_int32),
                ("dot_img", C.c_void_p), ("dot_img_atoms", C.c_int32), ("dot_atom0", C.c_int32)]


class nrc_density_mlp_t(C.Structure):
    _fields_ = [
        ("d_w0", C.c_void_p), ("d_b0", C.c_void_p),
        ("d_w1", C.c_void_p), ("d_b1", C.c_void_p),
        ("d_wd", C.c_void_p), ("d_bd", C.c_void_p),
        ("d_wn", C.c_void_p), ("d_bn", C.c_void_p),
        ("in_dim", C.c_int32), ("width", C.c_int32),
    ]


class nrc_density_mlp_grad_t(C.Structure):
    _fields_ = [
        ("d_w0", C.c_void_p), ("d_b0", C.c_void_p),
        ("d_w1", C.c_void_p), ("d_b1", C.c_void_p),
        ("d_wd", C.c_void_p), ("d_bd", C.c_void_p),
        ("d_wn", C.c_void_p), ("d_bn", C.c_void_p),
    ]


class nrc_slf_points_t(C.Structure):
    _fields_ = [
        ("num_distance_samples", C.c_int32), ("warp_kind", C.c_int32),
        ("distance_near", C.c_float), ("distance_far", C.c_float), ("near", C.c_float), ("far", C.c_float),
        ("distance_scale", C.c_float), ("distance_bias", C.c_float),
        ("rgb_premultiplier", C.c_float), ("rgb_bias", C.c_float), ("alpha_bias", C.c_float),
        ("warp_p", C.c_float), ("warp_premult", C.c_float), ("ref_warp_c", C.c_float),
    ]


_P = C.c_void_p
_I32 = C.c_int32
_I64 = C.c_int64
_F = C.c_float

# name -> argtypes (restype is int32 unless listed in _RESTYPES).  Must list every
# symbol include/nrc_b200.h declares (tests/test_abi.py checks this against the header).
PROTOTYPES = {
    "nrc_abi_version": [],
    "nrc_build_digest": [],
    "nrc_error_string": [_I32],
    "nrc_last_cuda_error": [],
    "nrc_encode_fwd": [_P, C.POINTER(nrc_encoding_t), _P, _I64, _P],
    "nrc_encode_indices": [_P, C.POINTER(nrc_encoding_t), _I32, _P, _I64, _P],
    "nrc_encode_bwd": [_P, C.POINTER(nrc_encoding_t), _P, _P, _I64, _P],
    "nrc_encode_bwd_warped": [_P, C.POINTER(nrc_encoding_t), _P, _F, _P, _I64],
    "nrc_contract_fwd": [_P, _P, _I64, _F, _P],
    "nrc_contract_bwd": [_P, _P, _P, _I64, _F, _P],
    "nrc_density_mlp_fwd": [_P, C.POINTER(nrc_density_mlp_t), _P, _I64, _I32, _P, _P, _P],
    "nrc_density_mlp_bwd": [_P, C.POINTER(nrc_density_mlp_t), _P, _P, _P, _P, _P, _I64, _I32, _P,
                            C.POINTER(nrc_density_mlp_grad_t)],
    "nrc_density_query_fwd": [_P, C.POINTER(nrc_encoding_t), C.POINTER(nrc_density_mlp_t), _P, _I64, _F, _F,
                              _I32, _P, _P, _P, _P, _P, _P],
    "nrc_ray_alpha_weights_fwd": [_P, _P, _P, _P, _I64, _I32, _I32, _P, _P, _P],
    "nrc_ray_alpha_weights_bwd": [_P, _P, _P, _P, _P, _P, _P, _I64, _I32, _P],
    "nrc_ray_sample_intervals": [_P, _P, _P, _P, _P, _I64, _I32, _I32, _F, _F, _F, _F, _F, _P, _P],
    "nrc_ray_cast": [_P, _P, _P, _P, _P, _P, _I64, _I32, _I32, _F, _F, _P, _P],
    "nrc_ray_cast_covs": [_P, _P, _P, _P, _P, _I64, _I32, _I32, _I32, _P, _P],
    "nrc_ray_sample_cast": [_P, _P, _P, _P, _P, _I64, _I32, _I32, _F, _F, _F, _F, _F, _P, _P, _P, _P, _I32, _F, _F, _P, _P, _P],
    "nrc_ray_weights_sample_cast": [_P, _P, _P, _P, _I32, _P, _P, _P, _I64, _I32, _I32, _F, _F, _F, _F, _F, _P, _P, _P, _P, _I32,
                                    _F, _F, _P, _P, _P],
    "nrc_ray_composite_fwd": [_P, _P, _P, _I32, _P, _P, _P, _I64, _I32, _I32, _I32, _P, _P, _P],
    "nrc_ray_composite_bwd": [_P, _P, _P, _I32, _P, _P, _P, _P, _I64, _I32, _I32, _I32, _P, _P, _P],
    "nrc_ray_resample": [_P, _P, _P, _I64, _I32, _I32, _F, _F, _P, _P],
    "nrc_ray_resample_gather": [_P, _P, _P, _I64, _I32, _I32, _I32, _P],
    "nrc_dense_fwd": [_P, _P, _I64, _P, _P, _I64, _I32, _I32, _I32, _I32, _P, _I64],
    "nrc_relu_bwd": [_P, _P, _I64, _P, _I64, _I64, _I32, _P, _I64],
    "nrc_dense_bwd": [_P, _P, _I64, _P, _P, _I64, _I64, _I32, _I32, _I32, _P, _I64, _I32, _P, _P],
    "nrc_ide_fwd": [_P, _I32, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_float), _P, _P, _P, _I64,
                    _P, _I64],
    "nrc_ide_bwd": [_P, _I32, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_float), _P, _P, _P, _P,
                    _I64, _I64, _P, _P],
    "nrc_mask_loss": [_P, _P, _I32, _P, _I64, _F, _F, _F, _P, _P],
    "nrc_vmf_head_fwd": [_P, _P, _P, _I32, _P, _I64, _I32, _F, _P, _P, _P],
    "nrc_vmf_head_bwd": [_P, _P, _P, _P, _P, _I64, _I32, _F, _P],
    "nrc_vmf_loss": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I32, _I32, _F, _I32, _P, _P, _P, _P],
    "nrc_grid_regularizer": [_P, _P, _F, _P],
    "nrc_grid_regularizer_init": [_P, _P, _F, _P],
    "nrc_zero_ranges": [_P, _P, C.POINTER(C.c_int64), C.POINTER(C.c_int64), _I32, _I32],
    "nrc_probe_gather": [_P, _P, _I64, _I32, _I64, _I32, _P],
    "nrc_allreduce_mean_multicast": [_P, _P, _I64, _I64, _I32, _I32, _I32],
    "nrc_allreduce_mean_peer": [_P, _P, _I64, _I64, _I32, _I32, _I32],
    "nrc_distortion_loss": [_P, _P, _P, _I32, _I64, _F, _F, _F, _P, _P],
    "nrc_geometry_losses": [_P, _P, _P, _P, _P, _I64, _I32, _F, _F, _F, _F, _P, _P, _P, _P],
    "nrc_density_normals_bwd": [_P, _P, _P, _P, _P, _I64, _F, _P],
    "nrc_encode_tangent_fwd": [_P, _P, _P, _P, _I64, _F, _P],
    "nrc_density_mlp_bwd_tangent": [_P, _P, _P, _P, _I64, _P, _P],
    "nrc_encode_tangent_bwd": [_P, _P, _P, _P, _P, _I64, _F],
    "nrc_transient_head_render_fwd": [_P, _P, _P, _I32, _P, _I64, _P, _P, _I32, _P, _I64, _P, _P, _P, _P, _P, _P, _I64, _I32, _I32,
                                      _I32, _F, _F, _F, _F, _F, _F, _F, _F, _I32, _F, _F, _F, _P, _I32, _P, _P, _P],
    "nrc_transient_filter": [_P, _P, _P, _I32, _I64, _I32, _I32, _P],
    "nrc_ggx_integrate_transient_fwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I32, _I32, _I32, _F, _P, _P, _P],
    "nrc_camera_rays": [_P, _P, _P, _I32, _I32, _I64, _P, _I64, _I64, _F, _F, _P, _P, _P, _P, _P, _P, _P],
    "nrc_chunk_advance": [_P, _P, _I64],
    "nrc_band_store": [_P, _P, _I32, _P, _I64, _I64, _I64, _P],
    "nrc_chain_run": [_P, _P, _P, _I32, _P, _I64],
    "nrc_chain_run_multi": [_P, _I32, _P, _P, _P, _P, _I64],
    "nrc_chain_query": [_P, _P, _P, _I32, _P, _I64, _P, _F],
    "nrc_chain_pack_weights": [_P, _P, _I32, _P, _I32, _P, _I32, _I32],
    "nrc_chain_wgrad": [_P, _P, _I32, _P, _I32, _I64],
    "nrc_shader_mid_fwd": [_P, _I32, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_float), _P, _I32, _P, _I64,
                           _P, _P, _I64, _I32, _F, _P, _P, _P, _P, _P, _P],
    "nrc_shader_mid_bwd": [_P, _I32, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_float), _P, _I32, _P, _I64,
                           _P, _P, _I64, _I32, _F, _P, _I64, _P, _I64, _P, _I64, _P, _I64, _P],
    "nrc_shader_out_fwd": [_P, _P, _I64, _P, _I64, _P, _I64, _P, _I64, _I64, _F, _F, _F, _F, _P, _P],
    "nrc_shader_out_bwd": [_P, _P, _I64, _P, _I64, _P, _I64, _I64, _F, _F, _F, _F, _P, _P, _I64, _P, _I64, _P, _I64],
    "nrc_normals_fwd": [_P, _P, _I64, _P],
    "nrc_normals_bwd": [_P, _P, _P, _I64, _P],
    "nrc_interlevel_loss": [_P, _P, _P, _I32, _P, _P, _I32, _I64, _F, _F, _F, _P, _P, _P],
    "nrc_charb_srgb_loss": [_P, _P, _P, _I64, _F, _P, _P],
    "nrc_render_loss": [_P, _P, _P, _P, _P, _P, _I64, _I32, _F, _I32, _F, _F, _P, _P, _P, _P, _P],
    "nrc_shade_render_loss": [_P, _P, _I64, _P, _I64, _P, _I64, _P, _I64, _F, _F, _F, _F, _P, _P, _P, _P, _I64, _I32, _F, _I32,
                              _F, _F, _P, _P, _P, _P, _P, _P, _I64, _P, _I64, _P, _I64],
    "nrc_cache_loss": [_P, _P, _P, _P, _I32, _P, _I32, _P, _I32, _I64, _F, _F, _P, _P, _P, _P],
    "nrc_pos_enc": [_P, _P, _I64, _I32, _I32, _I32, _I32, _P, _I64],
    "nrc_transient_render_fwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I32, _I32, _I32, _F, _F, _F, _F, _F, _I32, _F,
                                 _F, _F, _P, _P, _P],
    "nrc_transient_render_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I32, _I32, _I32, _F, _F, _F, _F, _F, _I32, _F,
                                 _F, _P, _P, _P, _P, _P, _P, _P],
    "nrc_secondary_sample": [_P, _P, _P, _P, _P, _I64, _I32, _I32, _I32, _P, _P, _P, _P, _I32, _P, _P, _F, _P, _P, _P, _P,
                             _P, _P],
    "nrc_material_head": [_P, _P, _I64, _I64, _F, _F, _P, _P, _P, _P, _P],
    "nrc_ggx_integrate_fwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I32, _I32, _F, _P, _P, _P],
    "nrc_ggx_integrate_bwd": [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I32, _I32, _F, _P],
    "nrc_slf_points_fwd": [_P, C.POINTER(nrc_slf_points_t), _P, _I64, _P, _P, _I64, _P, _P, _P, _P, _P],
    "nrc_slf_points_bwd": [_P, C.POINTER(nrc_slf_points_t), _P, _I64, _P, _P, _I64, _P, _P, _P, _P, _P, _P],
    "nrc_slf_reduce_fwd": [_P, _P, _P, _I64, _I32, _I32, _P],
    "nrc_slf_reduce_bwd": [_P, _P, _P, _P, _I64, _I32, _I32, _P, _P],
}
_RESTYPES = {"nrc_error_string": C.c_char_p, "nrc_build_digest": C.c_char_p}

_lib = None


def load():
    """Load libnrc_b200.so; raise loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NrcError(
            f"{LIB_PATH} is missing: build it with `python -m neural_radiance_caching_b200.build` "
            "(there is no CPU fallback for the radiance-cache query path)."
        )
    lib = C.CDLL(LIB_PATH)
    missing = []
    for name, argtypes in PROTOTYPES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            missing.append(name)
            continue
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int32)
    if missing:
        raise NrcError(f"{LIB_PATH} does not export {missing}: stale build? run the build with --force")
    # the library must have been compiled from the sources beside it (the .so is git-ignored and survives a pull)
    if os.path.isdir(os.path.join(HERE, "csrc")) and os.environ.get("NRC_SKIP_DIGEST_CHECK") != "1":
        from . import build as _build
        have, want = lib.nrc_build_digest().decode(), _build._digest()
        if have != want:
            raise NrcError(f"{LIB_PATH} was built from other sources (digest {have[:12]} != {want[:12]}): "
                           "run `python -m neural_radiance_caching_b200.build`")
    _lib = lib
    return lib


def check(status, what):
    if status != 0:
        lib = load()
        msg = lib.nrc_error_string(status).decode()
        extra = f" (cudaError {lib.nrc_last_cuda_error()})" if status == -3 else ""
        raise NrcError(f"{what}: {msg}{extra}")


def ptr(t):
    """Device pointer of a contiguous fp32/int32 CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not isinstance(t, torch.Tensor):
        raise NrcError(f"expected a torch.Tensor, got {type(t)}")
    if not t.is_cuda:
        raise NrcError("nrc_b200 kernels need CUDA tensors: there is no CPU fallback")
    if not t.is_contiguous():
        raise NrcError("nrc_b200 kernels need contiguous tensors")
    if t.dtype not in (torch.float32, torch.int32, torch.bfloat16):
        raise NrcError(f"unsupported dtype {t.dtype}")
    return C.c_void_p(t.data_ptr())


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


# Count of kernel-launching ABI calls (bench.py reports it as gpu_launches).
launch_count = 0


def call(name, *args):
    global launch_count
    lib = load()
    launch_count += 1
    check(getattr(lib, name)(*args), name)


# ------------------------------------------------------------------------------ gradient sinks
# The C ABI accumulates parameter gradients into caller-owned, caller-zeroed buffers.  A training
# harness can register one persistent buffer per parameter (views of ONE flat arena: a single
# memset per step, a single NCCL all-reduce for data parallelism); the custom VJPs then accumulate
# straight into it and hand autograd `None` for that parameter, instead of allocating and zero-
# filling a fresh gradient per layer per step.
_grad_sinks = {}


def register_grad_sink(param, sink):
    if sink.shape != param.shape or not sink.is_contiguous():
        raise NrcError("gradient sink must be a contiguous tensor of the parameter's shape")
    import weakref
    _grad_sinks[param.data_ptr()] = (weakref.ref(param), sink)


def clear_grad_sinks():
    _grad_sinks.clear()


def grad_sink(param):
    """Sink registered for the parameter whose storage `param` views, or None.  Entries die with the
    registered parameter (the allocator may hand its address to an unrelated tensor afterwards)."""
    ent = _grad_sinks.get(param.data_ptr())
    if ent is None:
        return None
    ref, sink = ent
    owner = ref()
    if owner is None or owner.data_ptr() != param.data_ptr() or owner.shape != param.shape:
        if owner is None:
            del _grad_sinks[param.data_ptr()]
        return None
    return sink
