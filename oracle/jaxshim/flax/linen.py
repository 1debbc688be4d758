"""flax.linen subset (Module/Dense/Sequential/initializers/celu/silu) on torch tensors.

Test infrastructure: restates the *published* flax semantics the SAKE reference relies on
(setup()-style modules, lazily created Dense params, `layers_<i>` naming inside Sequential,
attribute-name scoping, `.init` / `.apply(..., method=...)`).  No SAKE arithmetic lives here.
"""
import copy
import math
import torch as _t
import jax as _jax
from jax import nn as _jnn


def silu(x):
    return _jnn.silu(x)


swish = silu


def celu(x, alpha=1.0):
    return _jnn.celu(x, alpha)


def tanh(x):
    return _t.tanh(x)


class initializers:  # noqa: N801
    @staticmethod
    def constant(value):
        def init(key, shape, dtype=None):
            v = value if isinstance(value, _t.Tensor) else _t.tensor(value, dtype=_jax.get_dtype())
            return _t.broadcast_to(v.to(_jax.get_dtype()), tuple(shape)).clone()
        return init

    @staticmethod
    def lecun_normal():
        def init(key, shape, dtype=None):
            fan_in = shape[0]
            std = math.sqrt(1.0 / fan_in) / 0.87962566103423978
            w = _t.empty(tuple(shape), dtype=_t.float64)
            _t.nn.init.trunc_normal_(w, mean=0.0, std=1.0, a=-2.0, b=2.0, generator=key)
            # parameters live at fp32 precision (flax default param dtype)
            return (w * std).to(_t.float32).to(_jax.get_dtype())
        return init

    @staticmethod
    def zeros(key, shape, dtype=None):
        return _t.zeros(tuple(shape), dtype=_jax.get_dtype())


_MISSING = object()


class Module:
    """Dataclass-like module with flax naming/scoping rules (subset)."""

    _fields = ()

    def __init_subclass__(cls, **kw):
        super().__init_subclass__(**kw)
        fields = []
        for klass in reversed(cls.__mro__):
            ann = klass.__dict__.get("__annotations__", {})
            for name in ann:
                if name.startswith("_"):
                    continue
                default = klass.__dict__.get(name, _MISSING)
                fields = [f for f in fields if f[0] != name]
                fields.append((name, default))
        # inherited defaults: look them up on the class if not set locally
        fixed = []
        for name, default in fields:
            if default is _MISSING and hasattr(cls, name):
                default = getattr(cls, name)
            fixed.append((name, default))
        cls._fields = tuple(fixed)

    def __init__(self, *args, **kwargs):
        object.__setattr__(self, "_parent", None)
        object.__setattr__(self, "_name", None)
        object.__setattr__(self, "_root", None)
        object.__setattr__(self, "_setup_done", False)
        names = [f[0] for f in self._fields]
        if len(args) > len(names):
            raise TypeError("too many positional arguments")
        vals = dict(zip(names, args))
        for k, v in kwargs.items():
            if k not in names:
                raise TypeError(f"unexpected field {k}")
            vals[k] = v
        for name, default in self._fields:
            if name in vals:
                object.__setattr__(self, name, vals[name])
            elif default is not _MISSING:
                object.__setattr__(self, name, default)
            else:
                raise TypeError(f"missing field {name}")

    # ---- binding ------------------------------------------------------------------
    def _bind(self, parent, name, root):
        object.__setattr__(self, "_parent", parent)
        object.__setattr__(self, "_name", name)
        object.__setattr__(self, "_root", root)
        # adopt modules passed as constructor fields (e.g. Sequential.layers)
        for fname, _ in self._fields:
            self._adopt(fname, getattr(self, fname))
        if not self._setup_done:
            object.__setattr__(self, "_setup_done", True)
            if hasattr(self, "setup"):
                self.setup()

    def _adopt(self, attr, value):
        if isinstance(value, Module):
            if value._root is None and self._root is not None:
                value._bind(self, attr, self._root)
        elif isinstance(value, (list, tuple)):
            for i, v in enumerate(value):
                if isinstance(v, Module) and v._root is None and self._root is not None:
                    v._bind(self, f"{attr}_{i}", self._root)

    def __setattr__(self, key, value):
        object.__setattr__(self, key, value)
        if not key.startswith("_") and self._root is not None:
            self._adopt(key, value)

    def _path(self):
        p = []
        m = self
        while m._parent is not None:
            p.append(m._name)
            m = m._parent
        return list(reversed(p))

    def param(self, name, init_fn, *init_args):
        root = self._root
        tree = root["params"]
        path = self._path()
        if root["mode"] == "init":
            for k in path:
                tree = tree.setdefault(k, {})
            if name not in tree:
                tree[name] = init_fn(root["rng"], *init_args)
            return tree[name]
        for k in path:
            tree = tree[k]
        return tree[name]

    def is_initializing(self):
        return self._root["mode"] == "init"

    def _clone(self):
        new = copy.copy(self)
        object.__setattr__(new, "_parent", None)
        object.__setattr__(new, "_name", None)
        object.__setattr__(new, "_root", None)
        object.__setattr__(new, "_setup_done", False)
        for fname, _ in self._fields:
            v = getattr(self, fname)
            if isinstance(v, Module):
                object.__setattr__(new, fname, v._clone())
            elif isinstance(v, (list, tuple)):
                object.__setattr__(new, fname, type(v)(
                    x._clone() if isinstance(x, Module) else x for x in v))
        return new

    def _run(self, root, args, kwargs, method):
        m = self._clone()
        m._bind(None, None, root)
        if method is None:
            fn = m.__call__
        else:
            fname = method if isinstance(method, str) else method.__name__
            fn = getattr(m, fname)
        return fn(*args, **kwargs)

    def init(self, rng, *args, method=None, **kwargs):
        root = {"params": {}, "mode": "init", "rng": rng}
        self._run(root, args, kwargs, method)
        return {"params": root["params"]}

    def apply(self, variables, *args, method=None, **kwargs):
        root = {"params": variables["params"], "mode": "apply", "rng": None}
        return self._run(root, args, kwargs, method)


class Dense(Module):
    features: int
    use_bias: bool = True

    def __call__(self, x):
        kernel = self.param("kernel", initializers.lecun_normal(), (x.shape[-1], self.features))
        y = x @ kernel
        if self.use_bias:
            bias = self.param("bias", initializers.zeros, (self.features,))
            y = y + bias
        return y


class Sequential(Module):
    layers: list

    def __call__(self, x):
        for layer in self.layers:
            x = layer(x)
        return x
