"""A minimal stand-in for `jax` / `jax.numpy` / `flax` / `gin` built on NumPy, used ONLY by
tests/golden/make_reference_vectors.py to EXECUTE the reference's own function bodies (internal/math.py, stepfun.py,
coord.py, render.py, linspline.py, grid_utils.py, ref_utils.py under /root/reference) in this container, where JAX is
not installable.  It makes the golden vectors "the reference's source run with NumPy float32 semantics" - not XLA: op
order inside a NumPy primitive (cumsum, sum) is NumPy's, transcendental functions are libm's.

Semantics reproduced on purpose:
  * default float dtype float32 (linspace / arange / zeros / array of Python floats);
  * jax.random.uniform(key, shape, minval, maxval): `key` IS a pre-drawn U[0,1) array -> minval + key * (maxval - minval)
    (how the kernels and the oracle take their randomness);
  * custom_jvp / custom_vjp: the primal function; stop_gradient: identity; vmap over the leading axis by a Python loop;
  * uint32 arithmetic wraps (NumPy does for arrays).
Not reproduced: JAX's weak-type promotion of int32 * float32 (NumPy widens to float64; the generator rounds such
results back to float32 when saving, which is exact for the +-1 sign factors where it occurs)."""
import functools
import sys
import types

import numpy as np

f32 = np.float32


def _as(x):
    a = np.asarray(x)
    if a.dtype == np.float64:
        a = a.astype(np.float32)
    if a.dtype == np.int64:
        a = a.astype(np.int32)
    return a


class _JnpModule(types.ModuleType):
    """numpy with float32 / int32 defaults; anything not overridden falls through to numpy."""

    def __getattr__(self, name):
        return getattr(np, name)


jnp = _JnpModule("jax.numpy")
jnp.ndarray = np.ndarray
jnp.float32, jnp.int32, jnp.uint32, jnp.bool_ = np.float32, np.int32, np.uint32, np.bool_
jnp.inf, jnp.pi, jnp.newaxis, jnp.nan = np.inf, np.pi, None, np.nan


def _wrap_f32(fn):
    @functools.wraps(fn)
    def w(*a, **k):
        out = fn(*a, **k)
        if isinstance(out, np.ndarray) and out.dtype == np.float64:
            return out.astype(np.float32)
        if isinstance(out, np.ndarray) and out.dtype == np.int64:
            return out.astype(np.int32)
        if isinstance(out, np.float64):
            return np.float32(out)
        return out
    return w


for _n in ("linspace", "arange", "zeros", "ones", "eye", "array", "asarray", "full", "geomspace", "logspace", "mean", "sum",
           "cumsum", "sqrt", "exp", "log", "sin", "cos", "arccos", "power", "where", "interp", "diff", "floor", "ceil",
           "maximum", "minimum", "clip", "abs", "square", "reciprocal", "cbrt", "log1p", "expm1", "tanh", "arctan2",
           "concatenate", "stack", "matmul", "cross", "prod", "max", "min", "sign", "nan_to_num", "select", "round",
           "meshgrid", "zeros_like", "ones_like", "full_like", "searchsorted", "argsort", "sort", "take_along_axis"):
    if hasattr(np, _n):
        setattr(jnp, _n, _wrap_f32(getattr(np, _n)))


def _matmul(a, b, precision=None, **k):
    out = np.matmul(a, b)
    return out.astype(np.float32) if out.dtype == np.float64 else out


jnp.matmul = _matmul


def _nan_to_num(x, copy=True, nan=0.0, posinf=None, neginf=None):
    # jnp.nan_to_num(x, jnp.inf): the second positional argument is `copy` (as in NumPy), so NaN -> 0
    return np.nan_to_num(x, nan=nan, posinf=posinf, neginf=neginf).astype(np.asarray(x).dtype)


jnp.nan_to_num = _nan_to_num


def _vectorize(fn=None, *, signature=None, excluded=None):
    def deco(f):
        v = np.vectorize(f, signature=signature, excluded=excluded or set())

        @functools.wraps(f)
        def w(*a, **k):
            out = v(*a, **k)
            return tuple(_as(o) for o in out) if isinstance(out, tuple) else _as(out)
        return w
    return deco(fn) if fn is not None else deco


jnp.vectorize = _vectorize


def _finfo(dt):
    return np.finfo(np.float32 if dt in (jnp.float32, np.float32, float) else dt)


jnp.finfo = _finfo
linalg = types.ModuleType("jax.numpy.linalg")
for _n in ("norm", "det", "slogdet", "inv", "eigh", "cholesky"):
    setattr(linalg, _n, _wrap_f32(getattr(np.linalg, _n)))
jnp.linalg = linalg


class _Custom:
    """jax.custom_jvp / jax.custom_vjp: only the primal is executed."""

    def __init__(self, fn, *a, **k):
        self.fn = fn
        functools.update_wrapper(self, fn)

    def __call__(self, *a, **k):
        return self.fn(*a, **k)

    def defjvp(self, f, *a, **k):
        return f

    def defjvps(self, *f):
        return None

    def defvjp(self, fwd, bwd, *a, **k):
        return None


def _custom(fn=None, **kw):
    if fn is None:
        return lambda f: _Custom(f)
    return _Custom(fn)


def _vmap(fn, in_axes=0, out_axes=0):
    def w(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = next(np.asarray(a).shape[ax] for a, ax in zip(args, axes) if ax is not None)
        outs = [fn(*[np.take(a, i, axis=ax) if ax is not None else a for a, ax in zip(args, axes)]) for i in range(n)]
        if isinstance(outs[0], tuple):
            return tuple(np.stack([o[j] for o in outs], axis=out_axes) for j in range(len(outs[0])))
        return np.stack(outs, axis=out_axes)
    return w


random = types.ModuleType("jax.random")
# a key is the pre-drawn array itself, or a dict of pre-drawn arrays by distribution when one key (split = identity)
# reaches several draws: {"uniform": ..., "normal": ..., "gumbel": ...}
_draw = lambda key, name: np.asarray(key[name] if isinstance(key, dict) else key, np.float32)
random.uniform = lambda key, shape=(), dtype=np.float32, minval=0.0, maxval=1.0: (
    np.float32(minval) + _draw(key, "uniform").reshape(shape) * np.float32(maxval - minval)).astype(np.float32)
random.normal = lambda key, shape=(), dtype=np.float32: _draw(key, "normal").reshape(shape)
random.split = lambda key, num=2: [key] * num
# jax.random.categorical(key, logits, axis, shape) = argmax(gumbel(key, <shape with the category axis re-inserted>) + logits,
# axis): `key` IS that pre-drawn Gumbel array (how the kernels and the oracle take the noise)
random.categorical = lambda key, logits, axis=-1, shape=None: np.argmax(_draw(key, "gumbel") + logits, axis=axis).astype(np.int32)
random.PRNGKey = lambda seed: seed

lax = types.ModuleType("jax.lax")
lax.stop_gradient = lambda x: x
lax.Precision = types.SimpleNamespace(HIGHEST=None, DEFAULT=None, HIGH=None)

nn_mod = types.ModuleType("jax.nn")
nn_mod.softmax = lambda x, axis=-1: (lambda e: e / e.sum(axis=axis, keepdims=True))(np.exp(x - x.max(axis=axis, keepdims=True)))
nn_mod.softplus = lambda x: np.logaddexp(x, np.float32(0)).astype(np.float32)
nn_mod.sigmoid = lambda x: (1 / (1 + np.exp(-x))).astype(np.float32)
nn_mod.relu = lambda x: np.maximum(x, 0)
nn_mod.tanh = np.tanh
nn_mod.silu = lambda x: (x / (1 + np.exp(-x))).astype(np.float32)

jax = types.ModuleType("jax")
jax.numpy, jax.random, jax.lax, jax.nn = jnp, random, lax, nn_mod
jax.custom_jvp, jax.custom_vjp = _custom, _custom
jax.vmap = _vmap
jax.jit = lambda f=None, **k: (f if f is not None else (lambda g: g))
jax.named_scope = lambda name: (lambda f: f)
jax.Array = np.ndarray
tree_util = types.ModuleType("jax.tree_util")
tree_util.tree_map = lambda f, t: {k: f(v) for k, v in t.items()} if isinstance(t, dict) else f(t)


class _DictKey:
    """jax.tree_util.DictKey: a path element (NOT a str - `"name" in path` is False, as in jax)."""

    def __init__(self, key):
        self.key = key


tree_util.tree_map_with_path = lambda f, t: {k: f((_DictKey(k),), v) for k, v in t.items()}
jax.tree_util = tree_util
experimental = types.ModuleType("jax.experimental")
checkify = types.ModuleType("jax.experimental.checkify")
checkify.check = lambda cond, msg, **k: None
checkify.checkify = lambda f, **k: f
experimental.checkify = checkify
jax.experimental = experimental
jscipy = types.ModuleType("jax.scipy")
jax.scipy = jscipy


class _AtArray(np.ndarray):
    """`x.at[idx].add(v)` / `.set(v)` of jax arrays: functional update, duplicate indices accumulate, negative indices
    wrap, out-of-bounds updates are dropped (jax's default scatter mode)."""

    @property
    def at(self):
        arr = self

        class _Idx:
            def __getitem__(self, idx):
                class _Op:
                    def _apply(self, v, add):
                        out = np.array(arr, copy=True).view(_AtArray)
                        i = np.asarray(idx).astype(np.int64)
                        i = np.where(i < 0, i + out.shape[0], i)
                        ok = (i >= 0) & (i < out.shape[0])
                        vb = np.broadcast_to(np.asarray(v, out.dtype), i.shape + out.shape[1:])
                        if add:
                            np.add.at(out, i[ok], vb[ok])
                        else:
                            out[i[ok]] = vb[ok]
                        return out

                    def add(self, v):
                        return self._apply(v, True)

                    def set(self, v):
                        return self._apply(v, False)
                return _Op()
        return _Idx()


class F32Array(np.ndarray):
    """jax's type promotion with x64 disabled, for the two cases NumPy gets differently: a float64 operand (a NumPy
    constant such as HashEncoding.bbox) is demoted to float32, and an int32 operand meeting a float32 one (x * grid_size)
    is promoted to float32, not float64.  Inputs wrapped in it keep every intermediate float32, as under jax."""

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kw):
        arrs = [np.asarray(i) if isinstance(i, (np.ndarray, np.generic)) else i for i in inputs]
        has_f = any(isinstance(a, np.ndarray) and a.dtype.kind == "f" for a in arrs)
        conv = []
        for a in arrs:
            if isinstance(a, np.ndarray) and (a.dtype == np.float64 or (has_f and a.dtype.kind in "iu")):
                a = a.astype(np.float32)
            conv.append(a)
        if out is not None:
            kw["out"] = tuple(np.asarray(o) for o in out)
        res = getattr(ufunc, method)(*conv, **kw)
        wrap = lambda r: r.view(F32Array) if isinstance(r, np.ndarray) else r
        return tuple(wrap(r) for r in res) if isinstance(res, tuple) else wrap(res)


_zeros = jnp.zeros
jnp.zeros = lambda *a, **k: _zeros(*a, **k).view(_AtArray)


def _map_coordinates(input, coordinates, order, mode="constant", cval=0.0):
    """jax.scipy.ndimage.map_coordinates: coordinates in float32; an out-of-range TAP contributes cval (jax interpolates
    across the edge - SciPy calls that rule 'grid-constant'; SciPy's own 'constant' cuts at the edge instead)."""
    import scipy.ndimage
    coords = np.stack([np.asarray(c, np.float32) for c in coordinates]).astype(np.float64)
    m = {"constant": "grid-constant"}.get(mode, mode)
    return scipy.ndimage.map_coordinates(np.asarray(input, np.float32), coords, order=order, mode=m, cval=cval).astype(np.float32)


jscipy.ndimage = types.ModuleType("jax.scipy.ndimage")
jscipy.ndimage.map_coordinates = _map_coordinates


class _Anything:
    """Attribute access never fails (tf.io.gfile.GFile ...); called with one callable it returns it (a decorator),
    called with anything else it returns another _Anything (a decorator factory / an opaque object)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __getitem__(self, item):
        return _Anything()

    def __iter__(self):
        return iter(())

    def __call__(self, *a, **k):
        if len(a) == 1 and callable(a[0]) and not k and not isinstance(a[0], _Anything):
            return a[0]
        return _Anything()


class _Passthrough(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()


STUBBED = ("flax", "etils", "third_party", "jmp", "plyfile", "h5py", "scipy_io_stub", "pyexr", "OpenEXR", "Imath", "gdown", "lpips", "jaxlib", "clu", "tensorboardX", "torch_stub", "dm_pix", "chex", "ml_collections", "optax", "tensorflow", "cv2", "absl", "PIL", "mediapy", "rawpy", "scipy_stub",
           "orbax", "jaxcam", "trimesh", "matplotlib", "tqdm", "skimage", "sklearn", "torchvision", "pdb_stub", "jaxopt",
           "tensorflow_graphics", "pycolmap", "camp_zipnerf", "open3d", "plotly", "imageio")


class _StubFinder:
    """Empty pass-through modules for the reference's non-numeric dependencies (decorators return their argument,
    attribute access never fails): nothing numeric is ever taken from them."""

    def find_spec(self, name, path=None, target=None):
        import importlib.machinery
        top = name.split(".")[0]
        if top in STUBBED and name not in sys.modules:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _Passthrough(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


def _fallback(name):
    if name.startswith("__"):
        raise AttributeError(name)
    return _Anything()


def install():
    """Put the stand-ins into sys.modules (idempotent).  Heavy optional imports of the reference that the numeric
    functions never touch (tensorflow, cv2, flax, gin, absl, PIL, ...) become empty pass-through modules."""
    if not hasattr(np, "math"):
        import math as _math
        np.math = _math          # the reference calls np.math.factorial (removed in NumPy 2)
    for m_ in (jax, nn_mod, lax, random, jscipy, experimental):   # non-numeric names only (initializers, tree utilities ...);
        m_.__getattr__ = _fallback                                 # a numeric call through one fails loudly at np.asarray
    mods = {"jax": jax, "jax.numpy": jnp, "jax.random": random, "jax.lax": lax, "jax.nn": nn_mod,
            "jax.tree_util": tree_util, "jax.experimental": experimental, "jax.experimental.checkify": checkify,
            "jax.scipy": jscipy, "jax.scipy.ndimage": jscipy.ndimage, "jax.numpy.linalg": linalg}
    for k, v in mods.items():
        sys.modules[k] = v
    gin = _Passthrough("gin")
    gin.configurable = lambda *a, **k: (a[0] if (len(a) == 1 and callable(a[0]) and not k) else (lambda f: f))
    gin.config = _Passthrough("gin.config")
    sys.modules["gin"] = gin
    sys.modules["gin.config"] = gin.config
    flax = _Passthrough("flax")
    flax.__path__ = []
    linen = _Passthrough("flax.linen")
    linen.__path__ = []

    class Module:
        """flax.linen.Module as a plain class: dataclass-style keyword construction, no parameter scoping - a test drives
        `setup()` itself and assigns parameters to the Dense stand-ins below."""

        def __init__(self, **kw):
            for k_, v_ in kw.items():
                setattr(self, k_, v_)

    class Dense(Module):
        """flax.linen.Dense: x @ kernel + bias in float32; `kernel` [in, features] / `bias` [features] assigned by the caller."""

        def __init__(self, features, use_bias=True, **kw):
            super().__init__(features=features, use_bias=use_bias, **kw)
            self.kernel = self.bias = None

        def __call__(self, x):
            assert self.kernel is not None and self.kernel.shape == (x.shape[-1], self.features), (self.kernel, x.shape)
            y = np.matmul(np.asarray(x, np.float32), self.kernel)
            return (y + self.bias).astype(np.float32) if self.use_bias else y.astype(np.float32)

    linen.Module, linen.Dense = Module, Dense
    global nn_mod_linen
    nn_mod_linen = linen
    linen.compact = lambda f: f
    linen.relu, linen.softplus, linen.sigmoid, linen.tanh = nn_mod.relu, nn_mod.softplus, nn_mod.sigmoid, np.tanh
    flax.linen = linen
    flax.struct = _Passthrough("flax.struct")
    import dataclasses

    def _struct_dataclass(cls):
        cls = dataclasses.dataclass(cls)
        cls.replace = lambda self, **kw: dataclasses.replace(self, **kw)
        return cls

    flax.struct.dataclass = _struct_dataclass
    sys.modules["flax"] = flax
    sys.modules["flax.linen"] = linen
    sys.modules["flax.struct"] = flax.struct
    mlc = _Passthrough("ml_collections")          # class-level defaults such as LightMLP.vmf_activation must survive
    mlc.__path__ = []
    mlc.FrozenConfigDict = mlc.ConfigDict = dict
    sys.modules["ml_collections"] = mlc
    if not any(isinstance(f, _StubFinder) for f in sys.meta_path):
        sys.meta_path.append(_StubFinder())
    return jax
