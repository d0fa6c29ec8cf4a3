"""NumPy-backed stand-ins for `jax`, `jax.numpy`, `haiku` and the small config / dataclass
packages the reference imports, so that the REFERENCE'S OWN MODEL SOURCE (structure_tokenizer/
model/*.py under /root/reference) can be executed unmodified in an image that has neither JAX nor
Haiku.  TEST INFRASTRUCTURE ONLY: used by tests/golden/make_golden_model.py to generate the
committed fixtures; nothing in the product imports it.

What this is and is not
  * The model code that runs is the reference's (Vq3D.encode_and_quantize and everything it calls:
    StructureEncoder, PositionalEncodingLayer, MPNNLayer, MaskedLayerNorm, CrossAttentionScaler,
    layer_stack, mapping.inference_subbatch, FiniteScalarCodebook ...).  Concatenation order, masks,
    parameter names, which LayerNorm is applied where - all of that comes from the reference.
  * The array library underneath is NumPy with JAX's default dtype rules re-imposed (x64 disabled:
    float32 / int32 everywhere; Python scalars are weakly typed), not XLA.  fp32 matmul summation
    order and the last ulp of tanh / exp / sin / cos therefore differ from a real XLA:CPU run; the
    fixtures are compared with a tolerance that covers this (tests/test_golden_model.py).
  * Haiku's module / parameter naming rules are restated from dm-haiku 0.0.10 (`module.py`:
    unique_and_canonical_name; "~" for modules built in __init__, "~method" for other methods).
  * jax.random is NOT threefry: initial weights are never compared, they are inputs.

`install()` registers the stand-ins in sys.modules and returns nothing; call it before importing
`structure_tokenizer`.
"""
from __future__ import annotations

import collections
import contextlib
import dataclasses
import functools
import inspect
import math
import re
import sys
import types
import zlib

import numpy as np

# =====================================================================================
#  JArray: ndarray with JAX's default (x64-disabled) dtype behaviour
# =====================================================================================
_DEMOTE = {np.dtype(np.float64): np.float32, np.dtype(np.int64): np.int32, np.dtype(np.uint64): np.uint32,
           np.dtype(np.complex128): np.complex64}
_FLOAT_UFUNCS = {np.true_divide, np.sqrt, np.exp, np.log, np.sin, np.cos, np.tan, np.tanh, np.arctan2, np.expm1,
                 np.log1p, np.reciprocal, np.arctan, np.arcsin, np.arccos, np.sinh, np.cosh, np.exp2, np.log2, np.cbrt}


def _demote(a):
    if isinstance(a, np.ndarray) or isinstance(a, np.generic):
        a = np.asarray(a)
        t = _DEMOTE.get(a.dtype)
        if t is not None:
            a = a.astype(t)
        return a.view(JArray)
    return a


def _demote_tree(x):
    if isinstance(x, (np.ndarray, np.generic)):
        return _demote(x)
    if isinstance(x, tuple) and not hasattr(x, "_fields"):
        return tuple(_demote_tree(v) for v in x)
    if isinstance(x, list):
        return [_demote_tree(v) for v in x]
    return x


class JArray(np.ndarray):
    __array_priority__ = 100.0

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kwargs):
        any_float = ufunc in _FLOAT_UFUNCS
        for x in inputs:
            if isinstance(x, float) or (isinstance(x, (np.ndarray, np.generic)) and np.asarray(x).dtype.kind == "f"):
                any_float = True
        prepped = []
        for x in inputs:
            if isinstance(x, (bool, np.bool_)):
                prepped.append(x)
            elif isinstance(x, (int, float)):
                # weakly typed Python scalar: takes the (x64-disabled) dtype of the other operand's kind
                prepped.append(np.float32(x) if any_float else (np.int32(x) if -2**31 <= x < 2**31 else x))
            elif isinstance(x, (np.ndarray, np.generic)):
                a = np.asarray(x)
                if any_float and a.dtype.kind in "iub":
                    a = a.astype(np.float32)
                else:
                    t = _DEMOTE.get(a.dtype)
                    if t is not None:
                        a = a.astype(t)
                prepped.append(a.view(np.ndarray))
            elif isinstance(x, (list, tuple)):
                prepped.append(np.asarray(_demote(np.asarray(x))))
            else:
                prepped.append(x)
        if out is not None:
            kwargs["out"] = tuple(np.asarray(o).view(np.ndarray) if isinstance(o, np.ndarray) else o for o in out)
        res = getattr(ufunc, method)(*prepped, **kwargs)
        if out is not None:
            return out[0] if len(out) == 1 else out
        if isinstance(res, tuple):
            return tuple(_demote(r) for r in res)
        if res is None:
            return None
        return _demote(res)

    def __getitem__(self, idx):
        r = np.ndarray.__getitem__(self, idx)
        if isinstance(r, np.generic):
            return np.asarray(r).view(JArray)
        return r

    def astype(self, dtype, *a, **k):
        dt = np.dtype(dtype)
        dt = np.dtype(_DEMOTE.get(dt, dt))
        return np.ndarray.astype(self, dt, *a, **k)

    def block_until_ready(self):
        return self

    def __hash__(self):  # jax arrays of size 1 are not hashable either; identity hash keeps dict/set use working
        return id(self)


def _wrap_np(fn):
    @functools.wraps(fn)
    def w(*args, **kwargs):
        kwargs.pop("precision", None)
        if "dtype" in kwargs and kwargs["dtype"] is not None:
            dt = np.dtype(kwargs["dtype"])
            kwargs["dtype"] = np.dtype(_DEMOTE.get(dt, dt))
        return _demote_tree(fn(*args, **kwargs))

    return w


def _mk(name, **kw):
    m = types.ModuleType(name)
    m.__dict__.update(kw)
    sys.modules[name] = m
    return m


# =====================================================================================
#  pytrees
# =====================================================================================
_PYTREE_REG = {}


def register_pytree_node(cls, flatten, unflatten):
    _PYTREE_REG[cls] = (flatten, unflatten)


class _Leaf:
    def __repr__(self):
        return "*"


_LEAF = _Leaf()


class PyTreeDef:
    def __init__(self, kind, meta, children):
        self.kind, self.meta, self.children = kind, meta, children

    @property
    def num_leaves(self):
        return 1 if self.kind is _LEAF else sum(c.num_leaves for c in self.children)

    def unflatten(self, leaves):
        it = iter(leaves)
        return _unflatten(self, it)


def _flatten(x, leaves, is_leaf=None):
    if is_leaf is not None and is_leaf(x):
        leaves.append(x)
        return PyTreeDef(_LEAF, None, [])
    if x is None:
        return PyTreeDef("none", None, [])
    t = type(x)
    if t in _PYTREE_REG:
        ch, aux = _PYTREE_REG[t][0](x)
        return PyTreeDef(t, aux, [_flatten(c, leaves, is_leaf) for c in ch])
    if isinstance(x, tuple) and hasattr(x, "_fields"):
        return PyTreeDef("namedtuple", t, [_flatten(c, leaves, is_leaf) for c in x])
    if t in (tuple, list):
        return PyTreeDef(t, None, [_flatten(c, leaves, is_leaf) for c in x])
    if isinstance(x, dict):
        keys = sorted(x.keys()) if t is dict or isinstance(x, collections.OrderedDict) is False else list(x.keys())
        return PyTreeDef("dict", (t, keys), [_flatten(x[k], leaves, is_leaf) for k in keys])
    leaves.append(x)
    return PyTreeDef(_LEAF, None, [])


def _unflatten(td, it):
    if td.kind is _LEAF:
        return next(it)
    if td.kind == "none":
        return None
    ch = [_unflatten(c, it) for c in td.children]
    if td.kind == "namedtuple":
        return td.meta(*ch)
    if td.kind in (tuple, list):
        return td.kind(ch)
    if td.kind == "dict":
        t, keys = td.meta
        d = dict(zip(keys, ch))
        try:
            return t(d)
        except Exception:
            return d
    return _PYTREE_REG[td.kind][1](td.meta, ch)


def tree_flatten(x, is_leaf=None):
    leaves = []
    td = _flatten(x, leaves, is_leaf)
    return leaves, td


def tree_unflatten(td, leaves):
    return td.unflatten(leaves)


def tree_leaves(x, is_leaf=None):
    return tree_flatten(x, is_leaf)[0]


def _flatten_up_to(td, x):
    """leaves of x cut at the structure of td (td may be a prefix of x's structure)"""
    if td.kind is _LEAF:
        return [x]
    if td.kind == "none":
        return []
    t = type(x)
    if t in _PYTREE_REG:
        ch = _PYTREE_REG[t][0](x)[0]
    elif isinstance(x, dict):
        ch = [x[k] for k in td.meta[1]]
    else:
        ch = list(x)
    out = []
    for c, ctd in zip(ch, td.children):
        out.extend(_flatten_up_to(ctd, c))
    return out


def tree_map(f, tree, *rest, is_leaf=None):
    leaves, td = tree_flatten(tree, is_leaf)
    others = [_flatten_up_to(td, r) for r in rest]
    return td.unflatten([f(*xs) for xs in zip(leaves, *others)])


def flatten_axes(name, treedef, axis_tree):
    """jax.api_util.flatten_axes: broadcast a prefix tree of axes over `treedef`."""
    dummy = treedef.unflatten([object()] * treedef.num_leaves)
    axes = []

    def add(ax, sub):
        axes.extend([ax] * len(tree_leaves(sub)))

    tree_map(add, axis_tree, dummy, is_leaf=lambda x: x is None)
    return axes


# =====================================================================================
#  jax.numpy / jax.nn / jax.lax / jax.random / jax.ops
# =====================================================================================
def _asj(x, dtype=None):
    a = np.asarray(x, dtype=dtype) if dtype is not None else np.asarray(x)
    return _demote(a)


def _gelu(x, approximate=True):
    # jax.nn.gelu (jax 0.4.23, _src/nn/functions.py): tanh form when approximate
    x = _asj(x)
    if approximate:
        sqrt_2_over_pi = np.sqrt(2 / np.pi).astype(x.dtype)
        cdf = 0.5 * (1.0 + np.tanh(sqrt_2_over_pi * (x + 0.044715 * (x ** 3))))
        return x * cdf
    from scipy.special import erf

    return _asj(x * (erf(x / np.sqrt(2)) + 1) / 2)


def _softmax(x, axis=-1, where=None, initial=None):
    x = _asj(x)
    unnormalized = np.exp(x - np.max(x, axis=axis, keepdims=True))
    return unnormalized / np.sum(unnormalized, axis=axis, keepdims=True)


def _sigmoid(x):
    x = _asj(x)
    return 1 / (1 + np.exp(-x))


def _one_hot(x, num_classes, dtype=np.float32, axis=-1):
    x = np.asarray(x)
    return _asj((x[..., None] == np.arange(num_classes, dtype=x.dtype)).astype(dtype))


def _segment_sum(data, segment_ids, num_segments=None, indices_are_sorted=False, unique_indices=False, bucket_size=None,
                 mode=None):
    data = np.asarray(data)
    out = np.zeros((num_segments,) + data.shape[1:], data.dtype)
    np.add.at(out, np.asarray(segment_ids), data)
    return _asj(out)


def _key_to_seed(key):
    return zlib.crc32(np.asarray(key, np.uint32).tobytes())


def _prngkey(seed):
    return _asj(np.array([0, int(seed) & 0xFFFFFFFF], np.uint32))


def _split(key, num=2):
    ss = np.random.SeedSequence([_key_to_seed(key), 0x5EED])
    return _asj(ss.generate_state(2 * num, np.uint32).reshape(num, 2))


def _fold_in(key, data):
    ss = np.random.SeedSequence([_key_to_seed(key), int(data) & 0xFFFFFFFF, 0xF01D])
    return _asj(ss.generate_state(2, np.uint32))


def _rng(key):
    return np.random.default_rng(_key_to_seed(key))


def _rand_normal(key, shape=(), dtype=np.float32):
    return _asj(_rng(key).standard_normal(shape).astype(dtype))


def _rand_uniform(key, shape=(), dtype=np.float32, minval=0.0, maxval=1.0):
    return _asj((_rng(key).random(shape) * (maxval - minval) + minval).astype(dtype))


def _rand_trunc_normal(key, lower, upper, shape=(), dtype=np.float32):
    r = _rng(key)
    out = r.standard_normal(shape)
    bad = (out < lower) | (out > upper)
    while bad.any():
        out[bad] = r.standard_normal(int(bad.sum()))
        bad = (out < lower) | (out > upper)
    return _asj(out.astype(dtype))


def _rand_bernoulli(key, p=0.5, shape=()):
    return _asj(_rng(key).random(shape) < p)


# ------------------------------------------------------------------------------ vmap
_VEC_OK = {}


def _shapes(x):
    return tuple(np.shape(l) for l in tree_leaves(x) if isinstance(l, (np.ndarray, np.generic)))


def _fn_key(f, args, kwargs):
    """(code object, shapes of every array bound into / passed to the function): the verdict
    'whole-array call == per-row calls' is only reused for calls of identical shape signature."""
    bound = []
    while isinstance(f, functools.partial):
        bound.append((_shapes(f.args), _shapes(f.keywords)))
        f = f.func
    f = getattr(f, "__wrapped__", f)
    f = getattr(f, "__func__", f)
    code = getattr(f, "__code__", None)
    if code is None:
        return None
    return (code, tuple(bound), _shapes(args), _shapes(kwargs))


def _index(a, i, ax):
    return _asj(np.take(np.asarray(a), i, axis=ax))


def _fresh(a):
    """a structurally identical copy of a non-mapped argument (stateful pytrees such as
    prng.SafeKey must not be shared between the per-row calls a real vmap traces once)"""
    leaves, td = tree_flatten(a)
    return td.unflatten(leaves)


def _equal_trees(y, yi, i):
    ly, ty = tree_flatten(y)
    li, ti = tree_flatten(yi)
    if len(ly) != len(li):
        return False
    for a, b in zip(ly, li):
        a, b = np.asarray(a), np.asarray(b)
        if a.ndim != b.ndim + 1 or a.shape[1:] != b.shape or a.dtype != b.dtype:
            return False
        if not np.array_equal(a[i], b, equal_nan=True):
            return False
    return True


def vmap(fun, in_axes=0, out_axes=0, axis_name=None, allow_vectorised=True, **_):
    """Semantics of jax.vmap by looping; as a speed-up, a function whose whole-array call is
    verified (bitwise, on three rows) to equal the per-row calls is evaluated in one call."""

    def mapped(*args, **kwargs):
        axes = tuple(in_axes) if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        assert len(axes) == len(args), (axes, len(args))
        n = None
        for a, ax in list(zip(args, axes)) + [(v, 0) for v in kwargs.values()]:
            if ax is None:
                continue
            for leaf in tree_leaves(a):
                n = np.asarray(leaf).shape[ax]
                break
            if n is not None:
                break
        assert n is not None, "vmap: nothing to map over"

        def take(i):
            a_i = [_fresh(a) if ax is None else tree_map(lambda l: _index(l, i, ax), a) for a, ax in zip(args, axes)]
            k_i = {k: tree_map(lambda l: _index(l, i, 0), v) for k, v in kwargs.items()}
            return a_i, k_i

        key = _fn_key(fun, args, kwargs)
        simple = allow_vectorised and out_axes == 0 and all(ax in (0, None) for ax in axes) and key is not None
        if simple and _VEC_OK.get(key) is not False:
            snap = _hk_snapshot()
            try:
                y = fun(*args, **kwargs)
                ok = _VEC_OK.get(key) is True
                if not ok:
                    ok = True
                    for i in sorted({0, n // 2, n - 1}):
                        a_i, k_i = take(i)
                        if not _equal_trees(y, fun(*a_i, **k_i), i):
                            ok = False
                            break
                if ok:
                    _VEC_OK[key] = True
                    return y
            except Exception:
                pass
            _hk_restore(snap)
            _VEC_OK[key] = False
        outs = []
        for i in range(n):
            a_i, k_i = take(i)
            outs.append(fun(*a_i, **k_i))
        return tree_map(lambda *xs: _asj(np.stack([np.asarray(x) for x in xs], axis=out_axes)), *outs)

    return mapped


def _scan(f, init, xs, length=None, reverse=False, unroll=1):
    leaves = tree_leaves(xs)
    if length is None:
        length = np.asarray(leaves[0]).shape[0]
    carry = init
    ys = []
    for i in range(length):
        x_i = tree_map(lambda l: _index(l, i, 0), xs)
        carry, y = f(carry, x_i)
        ys.append(y)
    if not ys or not tree_leaves(ys[0]):
        return carry, (ys[0] if ys else None)
    return carry, tree_map(lambda *v: _asj(np.stack([np.asarray(a) for a in v])), *ys)


def _dynamic_index_in_dim(a, index, axis=0, keepdims=True):
    a = _asj(a)
    r = np.take(a, int(index), axis=axis)
    if keepdims:
        r = np.expand_dims(r, axis)
    return _asj(r)


def _dynamic_slice_in_dim(a, start, slice_size, axis=0):
    a = _asj(a)
    sl = [slice(None)] * a.ndim
    sl[axis] = slice(int(start), int(start) + int(slice_size))
    return a[tuple(sl)]


def _dynamic_update_slice_in_dim(a, update, start, axis):
    a = np.array(a, copy=True)
    sl = [slice(None)] * a.ndim
    sl[axis] = slice(int(start), int(start) + np.asarray(update).shape[axis])
    a[tuple(sl)] = update
    return _asj(a)


def _install_jax():
    jnp = _mk("jax.numpy")
    for name in dir(np):
        obj = getattr(np, name)
        if name.startswith("_"):
            continue
        if isinstance(obj, np.ufunc):
            # route through JArray.__array_ufunc__ so that weak Python scalars get JAX's dtype
            def mkuf(uf):
                def w(*a, **k):
                    k.pop("precision", None)
                    a = tuple(_asj(x) if isinstance(x, (np.ndarray, np.generic, list, tuple)) else x for x in a)
                    if not any(isinstance(x, JArray) for x in a):
                        a = (_asj(a[0]),) + a[1:]
                    return uf(*a, **k)

                return w

            setattr(jnp, name, mkuf(obj))
        elif callable(obj) and not isinstance(obj, type):
            setattr(jnp, name, _wrap_np(obj))
        else:
            setattr(jnp, name, obj)
    jnp.ndarray = np.ndarray
    jnp.array = lambda x, dtype=None, copy=True: _asj(np.array(x, dtype=dtype))
    jnp.asarray = lambda x, dtype=None: _asj(x, dtype)
    jnp.bfloat16 = np.float16  # placeholder: mixed precision is off on the tokenize path
    jnp.float_ = np.float32
    jnp.int_ = np.int32
    jnp.divide = jnp.true_divide
    jnp.linalg = _mk("jax.numpy.linalg", norm=_wrap_np(np.linalg.norm), svd=_wrap_np(np.linalg.svd),
                     det=_wrap_np(np.linalg.det), inv=_wrap_np(np.linalg.inv))
    jnp.load = np.load

    def _vectorize(f, signature=None):
        return _wrap_np(np.vectorize(f, signature=signature))

    jnp.vectorize = _vectorize

    nn = _mk("jax.nn", relu=lambda x: np.maximum(_asj(x), 0), gelu=_gelu, softmax=_softmax, sigmoid=_sigmoid,
             swish=lambda x: _asj(x) * _sigmoid(x), silu=lambda x: _asj(x) * _sigmoid(x), one_hot=_one_hot,
             softplus=lambda x: np.logaddexp(_asj(x), 0), tanh=lambda x: np.tanh(_asj(x)),
             elu=lambda x: np.where(_asj(x) > 0, x, np.expm1(_asj(x))))
    lax = _mk(
        "jax.lax",
        rsqrt=lambda x: 1 / np.sqrt(_asj(x)),
        stop_gradient=lambda x: x,
        convert_element_type=lambda x, dt: _asj(np.asarray(x, dtype=dt)),
        dynamic_index_in_dim=_dynamic_index_in_dim,
        index_in_dim=_dynamic_index_in_dim,
        dynamic_slice_in_dim=_dynamic_slice_in_dim,
        dynamic_update_slice_in_dim=_dynamic_update_slice_in_dim,
        pmean=lambda x, axis_name=None: x,  # one device per process here
        psum=lambda x, axis_name=None: x,
        scan=_scan,
    )
    rnd = _mk("jax.random", PRNGKey=_prngkey, split=_split, fold_in=_fold_in, normal=_rand_normal,
              uniform=_rand_uniform, truncated_normal=_rand_trunc_normal, bernoulli=_rand_bernoulli)
    tu = _mk("jax.tree_util", tree_map=tree_map, tree_flatten=tree_flatten, tree_unflatten=tree_unflatten,
             tree_leaves=tree_leaves, register_pytree_node=register_pytree_node, PyTreeDef=PyTreeDef)
    ops = _mk("jax.ops", segment_sum=_segment_sum)

    def wraps(fun, docstr=None, **kw):
        def deco(g):
            try:
                g.__name__ = getattr(fun, "__name__", "fn")
            except Exception:
                pass
            return g

        return deco

    util = _mk("jax.util", wraps=wraps)
    api_util = _mk("jax.api_util", flatten_axes=flatten_axes)

    def jit(f=None, **kw):
        return f if f is not None else (lambda g: g)

    def pmap(f, axis_name=None, devices=None, **kw):
        return vmap(f, allow_vectorised=False)

    class _Cfg:
        def update(self, *a, **k):
            pass

    jax = _mk("jax", numpy=jnp, nn=nn, lax=lax, random=rnd, tree_util=tu, ops=ops, util=util, api_util=api_util,
              tree_map=tree_map, vmap=vmap, jit=jit, pmap=pmap, config=_Cfg(), Array=np.ndarray,
              local_device_count=lambda backend=None: 1, local_devices=lambda backend=None: ["cpu:0"],
              devices=lambda backend=None: ["cpu:0"], process_index=lambda: 0,
              device_put=lambda x, d=None: x, block_until_ready=lambda x: x,
              device_put_replicated=lambda x, devs: tree_map(lambda a: _asj(np.asarray(a)[None]), x),
              Device=object, named_call=lambda f, name=None: f)
    return jax


# =====================================================================================
#  haiku
# =====================================================================================
class _Frame:
    def __init__(self, params, is_init, rng):
        self.params = params
        self.is_init = is_init
        self.rng = rng
        self.rng_count = 0
        self.module_stack = []  # (module, method_name)
        self.counter_stack = [collections.Counter()]
        self.used_names_stack = [set()]
        self.creator_stack = []
        self.getter_stack = []


_FRAMES = []


def _frame() -> _Frame:
    if not _FRAMES:
        raise RuntimeError("haiku shim: must be called inside hk.transform")
    return _FRAMES[-1]


def _hk_snapshot():
    if not _FRAMES:
        return None
    f = _frame()
    return {k: set(v.keys()) for k, v in f.params.items()}


def _hk_restore(snap):
    if snap is None:
        return
    f = _frame()
    for k in list(f.params.keys()):
        if k not in snap:
            del f.params[k]
        else:
            for n in list(f.params[k].keys()):
                if n not in snap[k]:
                    del f.params[k][n]


_CAMEL = re.compile(r"((?<=[a-z0-9])[A-Z]|(?!^)[A-Z](?=[a-z]))")


def _camel_to_snake(v):
    return _CAMEL.sub(r"_\1", v).lower()


def _unique_and_canonical_name(name):
    """dm-haiku 0.0.10 _src/module.py: unique_and_canonical_name."""
    fr = _frame()
    if len(fr.module_stack) > 1:
        parent, method_name = fr.module_stack[-2]
        if method_name == "__init__":
            name = "~/" + name
        elif method_name != "__call__":
            name = "~" + method_name + "/" + name
        name = parent.module_name + "/" + name
    splits = re.split(r"_(\d+)$", name, 3)
    if len(splits) > 1:
        name, n, explicit = splits[0], int(splits[1]), True
    else:
        n, explicit = None, False
    counters = fr.counter_stack[-2]
    if n is not None:
        counters[name] = max(counters[name], n + 1)
    else:
        n = counters[name]
        counters[name] += 1
    qualified = f"{name}_{n}" if explicit or n else name
    used = fr.used_names_stack[-2]
    if qualified in used:
        raise ValueError(f"Module name '{qualified}' is not unique.")
    used.add(qualified)
    return qualified


def _wrap_method(name, fn):
    if getattr(fn, "_hk_transparent", False):
        return fn

    @functools.wraps(fn)
    def wrapped(self, *a, **k):
        if not _FRAMES:
            return fn(self, *a, **k)
        fr = _frame()
        fr.module_stack.append((self, name))
        fr.counter_stack.append(collections.Counter())
        fr.used_names_stack.append(set())
        try:
            return fn(self, *a, **k)
        finally:
            fr.module_stack.pop()
            fr.counter_stack.pop()
            fr.used_names_stack.pop()

    wrapped._hk_wrapped = True
    return wrapped


class _ModuleMeta(type):
    def __new__(mcs, cname, bases, d):
        for key, value in list(d.items()):
            if key.startswith("__") and key != "__call__":
                continue
            if inspect.isfunction(value):
                d[key] = _wrap_method(key, value)
        return super().__new__(mcs, cname, bases, d)

    def __call__(cls, *args, **kwargs):
        module = cls.__new__(cls, *args, **kwargs)
        init = _wrap_method("__init__", cls.__init__)
        init(module, *args, **kwargs)
        if not hasattr(module, "module_name"):
            raise ValueError("super().__init__() was not called")
        return module


class Module(metaclass=_ModuleMeta):
    def __init__(self, name=None):
        if name is None:
            name = _camel_to_snake(type(self).__name__)
        self.module_name = _unique_and_canonical_name(name)
        self.name = self.module_name.split("/")[-1]


def transparent(fn):
    fn._hk_transparent = True
    return fn


GetterContext = collections.namedtuple("GetterContext", "full_name module original_dtype original_shape original_init")


def get_parameter(name, shape, dtype=np.float32, init=None):
    fr = _frame()
    module = fr.module_stack[-1][0]
    bundle = module.module_name
    shape = tuple(int(s) for s in shape)
    ctx = GetterContext(bundle + "/" + name, module, dtype, shape, init)
    param = fr.params.get(bundle, {}).get(name)
    if param is None:
        if not fr.is_init:
            raise ValueError(f"parameter {bundle}/{name} missing at apply time")
        if init is None:
            raise ValueError("initializer required")
        creators = list(fr.creator_stack)

        def next_creator(shape, dtype, init):
            if creators:
                return creators.pop(0)(next_creator, shape, dtype, init, ctx)
            return init(shape, dtype)

        param = _asj(next_creator(shape, dtype, init))
        fr.params.setdefault(bundle, {})[name] = param
    param = _asj(param)
    getters = list(fr.getter_stack)

    def next_getter(value):
        if getters:
            return getters.pop(0)(next_getter, value, ctx)
        return value

    param = next_getter(param)
    assert tuple(param.shape) == shape, f"{bundle}/{name}: {param.shape} != {shape}"
    return param


@contextlib.contextmanager
def _push(stack, item):
    stack.append(item)
    try:
        yield
    finally:
        stack.pop()


def custom_creator(c):
    return _push(_frame().creator_stack, c)


def custom_getter(g):
    return _push(_frame().getter_stack, g)


def next_rng_key():
    fr = _frame()
    if fr.rng is None:
        raise ValueError("rng required")
    fr.rng_count += 1
    return _fold_in(fr.rng, fr.rng_count)


def maybe_next_rng_key():
    return next_rng_key() if _frame().rng is not None else None


@contextlib.contextmanager
def with_rng(key):
    fr = _frame()
    old, oldc = fr.rng, fr.rng_count
    fr.rng, fr.rng_count = key, 0
    try:
        yield
    finally:
        fr.rng, fr.rng_count = old, oldc


class Transformed:
    def __init__(self, f):
        self._f = f

    def init(self, rng, *a, **k):
        fr = _Frame({}, True, rng)
        _FRAMES.append(fr)
        try:
            self._f(*a, **k)
        finally:
            _FRAMES.pop()
        return fr.params

    def apply(self, params, rng, *a, **k):
        fr = _Frame(params, False, rng)
        _FRAMES.append(fr)
        try:
            return self._f(*a, **k)
        finally:
            _FRAMES.pop()


class Constant:
    def __init__(self, constant):
        self.constant = constant

    def __call__(self, shape, dtype):
        return _asj(np.broadcast_to(np.asarray(self.constant, dtype), shape).copy())


class TruncatedNormal:
    def __init__(self, stddev=1.0, mean=0.0):
        self.stddev, self.mean = stddev, mean

    def __call__(self, shape, dtype):
        unscaled = _rand_trunc_normal(next_rng_key(), -2.0, 2.0, shape, dtype)
        return _asj((np.asarray(self.stddev, dtype) * np.asarray(unscaled) + np.asarray(self.mean, dtype)).astype(dtype))


class RandomNormal:
    def __init__(self, stddev=1.0, mean=0.0):
        self.stddev, self.mean = stddev, mean

    def __call__(self, shape, dtype):
        return _asj((self.stddev * np.asarray(_rand_normal(next_rng_key(), shape, dtype)) + self.mean).astype(dtype))


def _compute_fans(shape):
    if len(shape) < 1:
        return 1, 1
    if len(shape) == 1:
        return shape[0], shape[0]
    if len(shape) == 2:
        return shape[0], shape[1]
    rf = int(np.prod(shape[:-2]))
    return shape[-2] * rf, shape[-1] * rf


class VarianceScaling:
    """dm-haiku 0.0.10 _src/initializers.py: VarianceScaling."""

    def __init__(self, scale=1.0, mode="fan_in", distribution="truncated_normal", fan_in_axes=None):
        self.scale, self.mode, self.distribution = scale, mode, distribution

    def __call__(self, shape, dtype):
        fan_in, fan_out = _compute_fans(shape)
        s = self.scale / max(1.0, {"fan_in": fan_in, "fan_out": fan_out, "fan_avg": (fan_in + fan_out) / 2.0}[self.mode])
        if self.distribution == "truncated_normal":
            stddev = np.sqrt(s) / 0.87962566103423978
            return TruncatedNormal(stddev=stddev)(shape, dtype)
        if self.distribution == "normal":
            return RandomNormal(stddev=np.sqrt(s))(shape, dtype)
        limit = np.sqrt(3.0 * s)
        return _asj(np.asarray(_rand_uniform(next_rng_key(), shape, dtype, -limit, limit)))


class Linear(Module):
    """dm-haiku 0.0.10 _src/basic.py: Linear."""

    def __init__(self, output_size, with_bias=True, w_init=None, b_init=None, name=None):
        super().__init__(name=name)
        self.input_size = None
        self.output_size = output_size
        self.with_bias = with_bias
        self.w_init = w_init
        self.b_init = b_init or (lambda s, d: _asj(np.zeros(s, d)))

    def __call__(self, inputs, *, precision=None):
        inputs = _asj(inputs)
        input_size = self.input_size = inputs.shape[-1]
        dtype = inputs.dtype
        w_init = self.w_init
        if w_init is None:
            w_init = TruncatedNormal(stddev=1.0 / np.sqrt(input_size))
        w = get_parameter("w", [input_size, self.output_size], dtype, init=w_init)
        out = _asj(np.dot(np.asarray(inputs), np.asarray(w)))
        if self.with_bias:
            b = get_parameter("b", [self.output_size], dtype, init=self.b_init)
            b = np.broadcast_to(b, out.shape)
            out = out + b
        return out


def to_axes_or_slice(axis):
    if isinstance(axis, slice):
        return axis
    if isinstance(axis, int):
        return (axis,)
    return tuple(axis)


def to_abs_axes(axis, ndim):
    if isinstance(axis, slice):
        return tuple(range(ndim)[axis])
    return tuple(sorted({a % ndim for a in axis}))


class LayerNorm(Module):
    """dm-haiku 0.0.10 _src/layer_norm.py: LayerNorm (use_fast_variance=False)."""

    def __init__(self, axis, create_scale, create_offset, eps=1e-5, scale_init=None, offset_init=None,
                 use_fast_variance=False, name=None, *, param_axis=None):
        super().__init__(name=name)
        self.axis = to_axes_or_slice(axis)
        self.eps = eps
        self.create_scale, self.create_offset = create_scale, create_offset
        self.scale_init = scale_init or (lambda s, d: _asj(np.ones(s, d)))
        self.offset_init = offset_init or (lambda s, d: _asj(np.zeros(s, d)))
        self.param_axis = (-1,) if param_axis is None else to_axes_or_slice(param_axis)

    def __call__(self, inputs, scale=None, offset=None):
        inputs = _asj(inputs)
        axis = to_abs_axes(self.axis, inputs.ndim)
        mean = np.mean(inputs, axis=axis, keepdims=True)
        # jnp.var: mean(|x - mean|^2) (biased)
        centered = inputs - mean
        variance = np.mean(centered * centered, axis=axis, keepdims=True)
        param_axis = to_abs_axes(self.param_axis, inputs.ndim)
        if param_axis == (inputs.ndim - 1,):
            param_shape = (inputs.shape[-1],)
        else:
            param_shape = tuple(inputs.shape[i] if i in param_axis else 1 for i in range(inputs.ndim))
        if self.create_scale:
            scale = get_parameter("scale", param_shape, inputs.dtype, init=self.scale_init)
        elif scale is None:
            scale = np.array(1.0, dtype=inputs.dtype)
        if self.create_offset:
            offset = get_parameter("offset", param_shape, inputs.dtype, init=self.offset_init)
        elif offset is None:
            offset = np.array(0.0, dtype=inputs.dtype)
        scale = _asj(np.broadcast_to(scale, inputs.shape))
        offset = _asj(np.broadcast_to(offset, inputs.shape))
        mean = _asj(np.broadcast_to(mean, inputs.shape))
        eps = np.asarray(self.eps, variance.dtype)
        inv = scale * (1 / np.sqrt(variance + eps))
        return inv * (inputs - mean) + offset


class MLP(Module):
    """dm-haiku 0.0.10 _src/nets/mlp.py: MLP."""

    def __init__(self, output_sizes, w_init=None, b_init=None, with_bias=True, activation=None, activate_final=False,
                 name=None):
        super().__init__(name=name)
        self.activation = activation if activation is not None else (lambda x: np.maximum(_asj(x), 0))
        self.activate_final = activate_final
        self.layers = [Linear(o, w_init=w_init, b_init=b_init, with_bias=with_bias, name="linear_%d" % i)
                       for i, o in enumerate(output_sizes)]

    def __call__(self, inputs, dropout_rate=None, rng=None):
        out = inputs
        n = len(self.layers)
        for i, layer in enumerate(self.layers):
            out = layer(out)
            if i < n - 1 or self.activate_final:
                out = self.activation(out)
        return out


class Sequential(Module):
    def __init__(self, layers, name=None):
        super().__init__(name=name)
        self.layers = tuple(layers)

    def __call__(self, inputs, *args, **kwargs):
        out = inputs
        for i, layer in enumerate(self.layers):
            out = layer(out, *args, **kwargs) if i == 0 else layer(out)
        return out


def hk_vmap(f, in_axes=0, out_axes=0, axis_name=None, *, split_rng=False):
    return vmap(f, in_axes, out_axes, allow_vectorised=False)


def _install_haiku():
    init = _mk("haiku.initializers", Constant=Constant, TruncatedNormal=TruncatedNormal, RandomNormal=RandomNormal,
               VarianceScaling=VarianceScaling, Initializer=object,
               Orthogonal=lambda scale=1.0, axis=-1: VarianceScaling(scale))
    nets = _mk("haiku.nets", MLP=MLP)
    exp = _mk("haiku.experimental", custom_creator=custom_creator, custom_getter=custom_getter)

    class _MP:
        @staticmethod
        def set_policy(cls, policy):
            pass

    src = _mk("haiku._src")
    ln = _mk("haiku._src.layer_norm", AxisOrAxes=object, to_abs_axes=to_abs_axes, to_axes_or_slice=to_axes_or_slice)
    src.layer_norm = ln
    hk = _mk("haiku", Module=Module, transparent=transparent, get_parameter=get_parameter, transform=Transformed,
             Transformed=Transformed, Params=dict, initializers=init, nets=nets, experimental=exp, Linear=Linear,
             LayerNorm=LayerNorm, Sequential=Sequential, vmap=hk_vmap, scan=_scan, remat=lambda f, **k: f,
             eval_shape=lambda f, *a, **k: f(*a, **k), running_init=lambda: _frame().is_init,
             next_rng_key=next_rng_key, maybe_next_rng_key=maybe_next_rng_key, with_rng=with_rng,
             mixed_precision=_MP, dropout=lambda rng, rate, x: x, _src=src)
    return hk


# =====================================================================================
#  ml_collections / jax_dataclasses / misc
# =====================================================================================
class ConfigDict(dict):
    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, ConfigDict):
            v = ConfigDict(v)
        super().__setitem__(k, v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


def pytree_dataclass(cls):
    cls = dataclasses.dataclass(cls)
    names = [f.name for f in dataclasses.fields(cls)]
    register_pytree_node(cls, lambda x: ([getattr(x, n) for n in names], None),
                         lambda aux, ch: cls(**dict(zip(names, ch))))
    return cls


def install(reference_root="/root/reference"):
    _install_jax()
    _install_haiku()
    cd = _mk("ml_collections.config_dict", ConfigDict=ConfigDict)
    _mk("ml_collections", ConfigDict=ConfigDict, config_dict=cd)
    _mk("jax_dataclasses", pytree_dataclass=pytree_dataclass)
    _mk("jmp", Policy=lambda **k: None, get_policy=lambda s: None)
    _mk("tree", map_structure=tree_map)
    _mk("biopandas")
    _mk("biopandas.pdb", PandasPdb=object)
    _mk("Bio")
    _mk("Bio.PDB", PDBParser=object)
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
