"""Eager, torch-backed stand-in for the sliver of the TensorFlow 1.x API that the reference's loss path touches.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  TensorFlow cannot be installed in this image, so the reference's
``networks.batch_hard`` / ``networks.lifted_loss`` (src/networks.py:797-870) and ``utils.all_diffs_tf`` / ``utils.cdist_tf``
(src/utils.py:302-311,343-360) cannot run as shipped.  With this module registered as ``tensorflow`` they DO run, unmodified,
line by line: every ``tf.*`` call they make lands on the torch op with the TF 1.x semantics stated next to it, in float32,
and torch autograd differentiates the composition.  ``oracle/make_golden_losses.py`` uses it to produce
``tests/golden/loss_*.npz``; ``tests/test_oracle_losses_golden.py`` checks the restatement ``oracle/losses_torch.py`` (the
oracle every GPU loss test compares against) with those files.

What this pins: the reference's SOURCE -- masks, reductions, weights, constants, the order of its ops -- as executed code
rather than as a reading of it.  What it cannot pin: TensorFlow's own kernels (summation order inside ``reduce_sum``, the
exact ``softplus`` polynomial); those stay at float32 round-off, which is the tolerance the loss tests use anyway.

Semantics that matter, and where TF documents them:
  reduce_max / reduce_min   gradient split EVENLY among ties (tensorflow/python/ops/math_grad.py, _MinOrMaxGrad) == torch
                            amax / amin; of an empty tensor: the reducer's identity (-inf / +inf since TF 1.13)
  reduce_logsumexp          max-shifted; of an empty tensor -inf (math_ops.reduce_logsumexp guards the non-finite max)
  boolean_mask(x, m)        the selected elements as a 1-D tensor; gradient scatters back
  map_fn(fn, elems, dtype)  fn applied to the slices along axis 0 of every tensor in ``elems``, results stacked
  logical_xor, eye(bool)    as named
  cast(bool -> float32)     0.0 / 1.0
  nn.softplus               log(1 + exp(x)), computed without overflow
"""
from __future__ import annotations

import contextlib
import sys
import types

import torch


def _t(x, like=None):
    if torch.is_tensor(x):
        return x
    dtype = like.dtype if (like is not None and isinstance(x, float)) else None
    return torch.as_tensor(x, dtype=dtype)


def _pair(a, b):
    """Python scalars take the tensor operand's dtype (TF converts constants to the other argument's dtype)."""
    if torch.is_tensor(a) and not torch.is_tensor(b):
        return a, torch.as_tensor(b, dtype=a.dtype)
    if torch.is_tensor(b) and not torch.is_tensor(a):
        return torch.as_tensor(a, dtype=b.dtype), b
    return _t(a), _t(b)


def _reduce(fn_all, fn_axis, x, axis):
    return fn_all(x) if axis is None else fn_axis(x, axis)


def _reduce_max(x, axis=None):
    if x.numel() == 0 and axis is None:
        return torch.tensor(float("-inf"), dtype=x.dtype)
    return _reduce(lambda t: t.amax(), lambda t, a: t.amax(dim=a), x, axis)


def _reduce_min(x, axis=None):
    if x.numel() == 0 and axis is None:
        return torch.tensor(float("inf"), dtype=x.dtype)
    return _reduce(lambda t: t.amin(), lambda t, a: t.amin(dim=a), x, axis)


def _reduce_logsumexp(x, axis=None):
    if x.numel() == 0 and axis is None:
        return torch.tensor(float("-inf"), dtype=x.dtype)
    return _reduce(lambda t: torch.logsumexp(t.reshape(-1), 0), lambda t, a: torch.logsumexp(t, dim=a), x, axis)


def _map_fn(fn, elems, dtype=None):
    if isinstance(elems, (tuple, list)):
        n = elems[0].shape[0]
        out = [fn(tuple(e[i] for e in elems)) for i in range(n)]
    else:
        out = [fn(elems[i]) for i in range(elems.shape[0])]
    return torch.stack([o if dtype is None else o.to(dtype) for o in out])


def _softplus(x):
    return torch.clamp(x, min=0) + torch.log1p(torch.exp(-x.abs()))


def _eye(n, dtype=torch.float32):
    return torch.eye(int(n), dtype=dtype)


def _shape(x):
    return torch.tensor(list(x.shape), dtype=torch.int32)


def _binary(op):
    def f(a, b):
        a, b = _pair(a, b)
        return op(a, b)
    return f


def build() -> types.ModuleType:
    """The ``tensorflow`` module object (plus ``tensorflow.python.ops.rnn`` for networks.py's import line)."""
    tf = types.ModuleType("tensorflow")
    tf.__doc__ = __doc__
    tf.float32, tf.float64, tf.int32, tf.int64, tf.bool = torch.float32, torch.float64, torch.int32, torch.int64, torch.bool
    tf.name_scope = lambda *a, **k: contextlib.nullcontext()
    tf.cast = lambda x, dtype: _t(x).to(dtype)
    tf.shape = _shape
    tf.expand_dims = lambda x, axis: x.unsqueeze(axis)
    tf.equal = _binary(torch.eq)
    tf.not_equal = _binary(torch.ne)
    tf.greater = _binary(torch.gt)
    tf.logical_not = torch.logical_not
    tf.logical_xor = torch.logical_xor
    tf.eye = _eye
    tf.boolean_mask = lambda x, mask: x[mask]
    tf.map_fn = _map_fn
    tf.reduce_max = _reduce_max
    tf.reduce_min = _reduce_min
    tf.reduce_logsumexp = _reduce_logsumexp
    tf.reduce_sum = lambda x, axis=None: _reduce(lambda t: t.sum(), lambda t, a: t.sum(dim=a), x, axis)
    tf.multiply = _binary(torch.mul)
    tf.divide = _binary(torch.div)
    tf.maximum = _binary(torch.maximum)
    tf.square = torch.square
    tf.sqrt = torch.sqrt
    tf.abs = torch.abs
    nn = types.ModuleType("tensorflow.nn")
    nn.softplus = _softplus
    tf.nn = nn
    # `from tensorflow.python.ops.rnn import _transpose_batch_time` (src/networks.py:5) -- never called on this path
    python = types.ModuleType("tensorflow.python")
    ops = types.ModuleType("tensorflow.python.ops")
    rnn = types.ModuleType("tensorflow.python.ops.rnn")
    rnn._transpose_batch_time = lambda x: x.transpose(0, 1)
    python.ops, ops.rnn, tf.python = ops, rnn, python
    tf._submodules = {"tensorflow.nn": nn, "tensorflow.python": python, "tensorflow.python.ops": ops,
                      "tensorflow.python.ops.rnn": rnn}
    return tf


@contextlib.contextmanager
def installed():
    """``with installed(): import networks`` -- registers the shim as ``tensorflow`` for the duration and removes every
    module imported from the reference tree afterwards, so nothing leaks into other tests."""
    tf = build()
    names = ["tensorflow", *tf._submodules]
    saved = {n: sys.modules.get(n) for n in names + ["utils", "networks"]}
    sys.modules["tensorflow"] = tf
    sys.modules.update(tf._submodules)
    sys.modules.pop("utils", None)
    sys.modules.pop("networks", None)
    try:
        yield tf
    finally:
        for n, m in saved.items():
            if m is None:
                sys.modules.pop(n, None)
            else:
                sys.modules[n] = m
