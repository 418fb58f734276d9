"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

A minimal eager stand-in for the ``tensorflow`` module, backed by torch CPU tensors,
so that the reference sources under ``/root/reference/Code`` (model.py, utils.py,
train.py, dataClasses.py) can be imported and *executed unmodified* in a container
that has no TensorFlow.  It only exists to produce golden vectors (see
``oracle/make_golden.py``) and to cross-check ``oracle/closed_form.py``; it cannot
travel to the GPU box (``/root/reference`` is absent there) and nothing in the
``-m gpu`` tests, ``smoke()`` or ``bench.py`` imports it.

Only the graph-building subset the reference's hot path touches is covered
(op census: SURVEY.md App. C).  Session / placeholder / Saver APIs are deliberately
missing: callers invoke the reference's graph-building functions directly on tensors.

Semantics that matter for parity:
  * ``tf.Variable`` / ``tf.random_normal`` creation order is recorded and values are
    drawn from a caller-supplied provider, so the CUDA path can be fed the very same
    weights (order per conv: W0[M,Cout,Cin], b, u, c, v -- reference model.py:430-447).
  * ``tf.gather(x, idx, axis=1)`` with a 3-D index (the dead op at model.py:385, pruned
    by TF graph mode) is not evaluated.
  * ``tf.div`` on integers floors, ``tf.count_nonzero`` returns int64, ``tf.norm`` without
    axis is a global 2-norm, ``tf.squeeze`` drops every unit dimension.
"""
from __future__ import annotations

import contextlib
import sys
import types

import numpy as np
import torch

float32 = torch.float32
float64 = torch.float64
int32 = torch.int32
int64 = torch.int64
bool = torch.bool  # noqa: A001  (mirrors tf.bool)

__version__ = "1.15.0"  # first char '1' => reference skips its compat.v1 branch (model.py:9-13)


class _ShapeList(list):
    def as_list(self):
        return list(self)


def _get_shape(self):
    return _ShapeList(int(s) for s in self.shape)


torch.Tensor.get_shape = _get_shape  # reference calls x.get_shape().as_list() everywhere


def _t(x, dtype=None):
    if isinstance(x, torch.Tensor):
        return x if dtype is None else x.to(dtype)
    a = np.asarray(x)
    if dtype is None:
        if a.dtype == np.float64:
            dtype = torch.float32  # TF default float
        elif a.dtype == np.int64:
            dtype = torch.int32  # TF default int
    out = torch.from_numpy(np.ascontiguousarray(a))
    return out if dtype is None else out.to(dtype)


# ------------------------------------------------------------------ variables
class VariableLog:
    """Records variables in creation order; values come from ``provider``."""

    def __init__(self):
        self.provider = None
        self.created = []  # list of (name, tensor)

    def reset(self, provider=None):
        self.provider = provider
        self.created = []


variables = VariableLog()


class _PendingInit:
    def __init__(self, shape, stddev):
        self.shape = [int(s) for s in shape]
        self.stddev = float(stddev)


def random_normal(shape, mean=0.0, stddev=1.0, dtype=float32, seed=None, name=None):
    return _PendingInit(shape, stddev)


def truncated_normal(shape, mean=0.0, stddev=1.0, dtype=float32, seed=None, name=None):
    return _PendingInit(shape, stddev)


def Variable(initial, name=None, dtype=None, trainable=True):
    if isinstance(initial, _PendingInit):
        if variables.provider is None:
            raise RuntimeError("tf_standin: no variable provider installed")
        val = variables.provider(initial.shape, initial.stddev, name)
        val = _t(val, torch.float32).reshape(initial.shape).clone()
    else:
        val = _t(initial, dtype).clone()
    variables.created.append((name, val))
    return val


@contextlib.contextmanager
def variable_scope(name=None, *a, **k):
    yield


name_scope = variable_scope


# ------------------------------------------------------------------ creation
def constant(value, dtype=None, shape=None, name=None):
    t = _t(value, dtype)
    if shape is not None:
        t = t.reshape([int(s) for s in shape])
    return t


def zeros(shape, dtype=float32, name=None):
    return torch.zeros([int(s) for s in shape], dtype=dtype)


def ones(shape, dtype=float32, name=None):
    return torch.ones([int(s) for s in shape], dtype=dtype)


def zeros_like(x, dtype=None, name=None):
    return torch.zeros_like(_t(x), dtype=dtype)


def ones_like(x, dtype=None, name=None):
    return torch.ones_like(_t(x), dtype=dtype)


def range(*args, **kw):  # noqa: A001
    return torch.arange(*args, dtype=torch.int32)


def cast(x, dtype, name=None):
    return _t(x).to(dtype)


# ------------------------------------------------------------------ shape ops
def reshape(x, shape, name=None):
    return _t(x).reshape([int(s) for s in shape])


def transpose(x, perm=None, name=None):
    x = _t(x)
    if perm is None:
        perm = list(reversed(list(np.arange(x.dim()))))
    return x.permute([int(p) for p in perm])


def expand_dims(x, axis, name=None):
    return _t(x).unsqueeze(int(axis))


def squeeze(x, axis=None, name=None):
    x = _t(x)
    return x.squeeze() if axis is None else x.squeeze(int(axis))


def tile(x, multiples, name=None):
    return _t(x).repeat([int(m) for m in multiples])


def concat(values, axis, name=None):
    values = [_t(v) for v in values]
    return torch.cat(values, dim=int(axis))


def stack(values, axis=0, name=None):
    return torch.stack([_t(v) for v in values], dim=int(axis))


def slice(x, begin, size, name=None):  # noqa: A001
    x = _t(x)
    idx = []
    for d, (b, s) in enumerate(zip(begin, size)):
        b = int(b)
        s = int(s)
        e = x.shape[d] if s == -1 else b + s
        idx.append(builtins_slice(b, e))
    return x[tuple(idx)]


import builtins as _builtins  # noqa: E402

builtins_slice = _builtins.slice


class _DeadGather:
    """Result of the pruned tf.gather(x, adj, axis=1) at reference model.py:385."""


def gather(params, indices, axis=0, name=None):
    params = _t(params)
    indices = _t(indices)
    if int(axis) != 0:
        if indices.dim() >= 3:
            return _DeadGather()
        return torch.index_select(params, int(axis), indices.reshape(-1).long()).reshape(
            list(params.shape[: int(axis)]) + list(indices.shape) + list(params.shape[int(axis) + 1:])
        )
    return params[indices.long()]


# ------------------------------------------------------------------ math
def matmul(a, b, name=None):
    return torch.matmul(_t(a), _t(b))


def multiply(a, b, name=None):
    return _t(a) * _t(b)


def add(a, b, name=None):
    return _t(a) + _t(b)


def subtract(a, b, name=None):
    return _t(a) - _t(b)


def divide(a, b, name=None):
    return _t(a) / _t(b)


def div(a, b, name=None):
    a, b = _t(a), _t(b)
    if a.dtype.is_floating_point or b.dtype.is_floating_point:
        return a / b
    return torch.div(a, b, rounding_mode="floor")


def reciprocal(x, name=None):
    return torch.reciprocal(_t(x))


def square(x, name=None):
    return torch.square(_t(x))


def sqrt(x, name=None):
    return torch.sqrt(_t(x))


def abs(x, name=None):  # noqa: A001
    return torch.abs(_t(x))


def acos(x, name=None):
    return torch.acos(_t(x))


def minimum(a, b, name=None):
    return torch.minimum(_t(a), torch.as_tensor(b, dtype=_t(a).dtype))


def maximum(a, b, name=None):
    return torch.maximum(_t(a), torch.as_tensor(b, dtype=_t(a).dtype))


def cross(a, b, name=None):
    return torch.cross(_t(a), _t(b), dim=-1)


def is_nan(x, name=None):
    return torch.isnan(_t(x))


def map_fn(fn, elems, dtype=None, name=None):
    return torch.stack([fn(e) for e in _t(elems)], dim=0)


def where(cond, x=None, y=None, name=None):
    return torch.where(_t(cond), _t(x), _t(y))


def equal(a, b, name=None):
    return _t(a) == torch.as_tensor(b)


def not_equal(a, b, name=None):
    return _t(a) != torch.as_tensor(b)


def greater(a, b, name=None):
    return _t(a) > torch.as_tensor(b)


def less_equal(a, b, name=None):
    return _t(a) <= torch.as_tensor(b)


def _axes(axis):
    if axis is None:
        return None
    if isinstance(axis, (list, tuple)):
        return [int(a) for a in axis]
    return int(axis)


def reduce_sum(x, axis=None, keepdims=False, name=None, keep_dims=None):
    if keep_dims is not None:
        keepdims = keep_dims
    x = _t(x)
    return x.sum() if axis is None else x.sum(dim=_axes(axis), keepdim=keepdims)


def reduce_mean(x, axis=None, keepdims=False, name=None, keep_dims=None):
    if keep_dims is not None:
        keepdims = keep_dims
    x = _t(x)
    return x.mean() if axis is None else x.mean(dim=_axes(axis), keepdim=keepdims)


def reduce_max(x, axis=None, keepdims=False, name=None):
    x = _t(x)
    return x.max() if axis is None else x.amax(dim=_axes(axis), keepdim=keepdims)


def reduce_min(x, axis=None, keepdims=False, name=None):
    x = _t(x)
    return x.min() if axis is None else x.amin(dim=_axes(axis), keepdim=keepdims)


def reduce_all(x, axis=None, keepdims=False, name=None):
    x = _t(x)
    return x.all() if axis is None else x.all(dim=_axes(axis), keepdim=keepdims)


def reduce_any(x, axis=None, keepdims=False, name=None):
    x = _t(x)
    return x.any() if axis is None else x.any(dim=_axes(axis), keepdim=keepdims)


def count_nonzero(x, axis=None, keepdims=False, dtype=int64, name=None):
    x = _t(x)
    nz = (x != 0).to(torch.int64)
    return nz.sum() if axis is None else nz.sum(dim=_axes(axis), keepdim=keepdims)


def norm(x, ord="euclidean", axis=None, keepdims=False, name=None):
    x = _t(x)
    if axis is None:
        return torch.sqrt((x * x).sum())
    return torch.sqrt((x * x).sum(dim=_axes(axis), keepdim=keepdims))


class _NN(types.SimpleNamespace):
    pass


def _softmax(x, axis=-1, name=None):
    return torch.softmax(_t(x), dim=int(axis))


def _relu(x, name=None):
    return torch.relu(_t(x))


nn = _NN(softmax=_softmax, relu=_relu)


def install():
    """Register this module as ``tensorflow`` (plus the debug stub train.py:17 imports)."""
    me = sys.modules[__name__]
    sys.modules["tensorflow"] = me
    py = types.ModuleType("tensorflow.python")
    dbg = types.ModuleType("tensorflow.python.debug")
    py.debug = dbg
    me.python = py
    sys.modules["tensorflow.python"] = py
    sys.modules["tensorflow.python.debug"] = dbg
    return me
