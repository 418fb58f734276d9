"""TEST INFRASTRUCTURE ONLY.

Imports the reference sources *unmodified, from where they lie* under
``/root/reference/Code`` on top of ``oracle/tf_standin.py`` and exposes helpers to
execute their graph-building functions eagerly with injected weights.  Works only
in the build container (``/root/reference`` does not exist on the GPU box); used by
``oracle/make_golden.py`` and by the ``-m "not gpu"`` cross-checks, which skip when
the reference tree is absent.  Recipe: SURVEY.md App. E.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import time
import types

import numpy as np

REFERENCE_CODE = "/root/reference/Code"

_loaded = None


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_CODE, "model.py"))


def load():
    """Returns a namespace with the reference modules: model, utils, train, dataClasses, coarsening."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_CODE)
    sys.dont_write_bytecode = True  # the reference tree is read-only
    if not hasattr(time, "clock"):
        time.clock = time.perf_counter  # dataClasses.py:39 uses the removed time.clock
    from oracle import tf_standin

    tf_standin.install()
    if REFERENCE_CODE not in sys.path:
        sys.path.insert(0, REFERENCE_CODE)
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with contextlib.redirect_stdout(io.StringIO()):
            import model as ref_model  # noqa
            import utils as ref_utils  # noqa
            import train as ref_train  # noqa
            import dataClasses as ref_data  # noqa
            from lib import coarsening as ref_coarsening  # noqa
    _loaded = types.SimpleNamespace(
        tf=tf_standin, model=ref_model, utils=ref_utils, train=ref_train,
        dataClasses=ref_data, coarsening=ref_coarsening,
    )
    return _loaded


def rng_provider(seed: int):
    """Weights drawn as RandomState(seed).normal(0, stddev) in variable-creation order."""
    rs = np.random.RandomState(seed)

    def provider(shape, stddev, name):
        return rs.normal(0.0, stddev, size=shape).astype(np.float32)

    return provider


def list_provider(tensors):
    """Weights taken from an explicit list, in variable-creation order."""
    it = iter(tensors)

    def provider(shape, stddev, name):
        t = np.asarray(next(it), dtype=np.float32)
        assert list(t.shape) == list(shape), (t.shape, shape, name)
        return t

    return provider


def run(fn, *args, provider=None, quiet=True, **kwargs):
    """Calls a reference graph-building function eagerly.

    Returns (result, variables) where variables is the list of numpy arrays the
    function created through tf.Variable, in creation order.
    """
    ref = load()
    ref.tf.variables.reset(provider)
    ctx = contextlib.redirect_stdout(io.StringIO()) if quiet else contextlib.nullcontext()
    with ctx:
        out = fn(*args, **kwargs)
    created = [v.numpy().copy() for _, v in ref.tf.variables.created]
    return out, created


def to_np(x):
    import torch

    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy()
    if isinstance(x, (list, tuple)):
        return [to_np(v) for v in x]
    return np.asarray(x)
