"""Variable store with TensorFlow-1.x style auto-generated names.

The reference creates *unnamed* variables (`tf.Variable(initial)`, darknet.py:12,17) inside
`tf.variable_scope`s (darknet.py:144,189-198), so its checkpoints are keyed by TF's auto names:
`darknet19/Variable`, `darknet19/Variable_1`, `darknet19/batch_normalization/gamma`, ...,
`darknet19_detection/conv1/Variable`, ... ("Checkpoint / resume").  This
store reproduces that naming so a converted TF checkpoint (npz keyed by those names) maps onto
the network, and so that `reuse=True` finds the variables created by the first call.

Initialisers follow darknet.py:10-17: W ~ truncated_normal(stddev=0.1) (values beyond 2 sigma
re-drawn), b = 0.1; BN: gamma=1, beta=0, moving_mean=0, moving_variance=1 (TF defaults).
"""
from __future__ import annotations

import collections
import numpy as np


def truncated_normal(shape, stddev, rng):
    """tf.truncated_normal semantics: N(0, stddev) with samples beyond 2 sigma re-drawn.
    `rng` is a numpy RandomState; the draw order is part of the weight-reproducibility contract
    used by tests/golden/make_golden.py."""
    n = int(np.prod(shape))
    out = rng.standard_normal(n)
    bad = np.nonzero(np.abs(out) > 2.0)[0]
    while bad.size:
        out[bad] = rng.standard_normal(bad.size)
        bad = bad[np.abs(out[bad]) > 2.0]
    return (out * stddev).astype(np.float32).reshape(shape)


class VariableStore:
    """Ordered name -> numpy/torch array map plus TF's unique-name bookkeeping."""

    def __init__(self, seed=0):
        self.vars = collections.OrderedDict()
        self.rng = np.random.RandomState(seed)
        self._counters = {}          # (scope, base) -> next suffix
        self._scope = []
        self._bn_scope_counter = {}
        self.version = 0             # bumped on every mutation (engines re-pack lazily)

    # -- scopes ---------------------------------------------------------------------------
    def scope(self, name):
        store = self

        class _Ctx:
            def __enter__(self_inner):
                store._scope.append(name)
                return store

            def __exit__(self_inner, *exc):
                store._scope.pop()
                return False
        return _Ctx()

    def _prefix(self):
        return '/'.join(self._scope)

    def unique_name(self, base):
        """TF name uniquification inside the current scope: base, base_1, base_2, ..."""
        key = (self._prefix(), base)
        k = self._counters.get(key, 0)
        self._counters[key] = k + 1
        leaf = base if k == 0 else '%s_%d' % (base, k)
        p = self._prefix()
        return (p + '/' + leaf) if p else leaf

    def reset_name_counters(self):
        """Start a `reuse=True` pass: the same call sequence regenerates the same names."""
        self._counters = {}
        self._scope_uses = {}        # (parent prefix, scope name) -> times entered without reuse (TF opens name_1, name_2, ...)

    # -- creation -------------------------------------------------------------------------
    def get_or_create(self, name, init_fn):
        if name not in self.vars:
            self.vars[name] = init_fn()
            self.version += 1
        return self.vars[name]

    def weight_variable(self, shape):
        """darknet.py:10-12"""
        name = self.unique_name('Variable')
        return name, self.get_or_create(name, lambda: truncated_normal(shape, 0.1, self.rng))

    def bias_variable(self, shape):
        """darknet.py:15-17"""
        name = self.unique_name('Variable')
        return name, self.get_or_create(name, lambda: np.full(shape, 0.1, dtype=np.float32))

    def batch_norm_variables(self, channels):
        """tf.layers.batch_normalization creates scope batch_normalization[_k] with
        gamma, beta, moving_mean, moving_variance."""
        scope = self.unique_name('batch_normalization')
        mk = lambda v: (lambda: np.full([channels], v, dtype=np.float32))
        names = {}
        for leaf, v in (('gamma', 1.0), ('beta', 0.0), ('moving_mean', 0.0),
                        ('moving_variance', 1.0)):
            names[leaf] = scope + '/' + leaf
            self.get_or_create(names[leaf], mk(v))
        return names

    # -- access ---------------------------------------------------------------------------
    def __getitem__(self, name):
        return self.vars[name]

    def __setitem__(self, name, value):
        self.vars[name] = value
        self.version += 1

    def __contains__(self, name):
        return name in self.vars

    def names(self):
        return list(self.vars.keys())

    def trainable_names(self):
        return [n for n in self.vars if not (n.endswith('moving_mean') or n.endswith('moving_variance'))]

    def num_parameters(self):
        return int(sum(np.prod(np.shape(v)) for n, v in self.vars.items()
                       if n in set(self.trainable_names())))

    # -- persistence (npz keyed by TF names) ----------------------------------------------
    def save_npz(self, path, extra=None):
        arrays = {k: _to_numpy(v) for k, v in self.vars.items()}
        if extra:
            arrays.update(extra)
        np.savez(path, **arrays)

    def load_npz(self, path, strict=False):
        """Restore every variable whose name is in the file (the reference's warm-start rule,
        net_utils.py:85-89); returns (restored, missing)."""
        data = np.load(path)
        restored, missing = [], []
        for k in list(self.vars.keys()):
            if k in data.files:
                cur = self.vars[k]
                arr = data[k]
                if tuple(arr.shape) != tuple(np.shape(cur)):
                    raise ValueError('shape mismatch for %s: %s vs %s' % (k, arr.shape, np.shape(cur)))
                self._assign(k, arr)
                restored.append(k)
            else:
                missing.append(k)
        if strict and missing:
            raise KeyError('variables missing from %s: %s' % (path, missing[:5]))
        self.version += 1
        return restored, missing

    def _assign(self, k, arr):
        cur = self.vars[k]
        if isinstance(cur, np.ndarray):
            self.vars[k] = arr.astype(cur.dtype)
        else:  # torch tensor: keep device/dtype
            import torch
            cur.copy_(torch.as_tensor(arr, dtype=cur.dtype))


def _to_numpy(v):
    if isinstance(v, np.ndarray):
        return v
    return v.detach().cpu().numpy()


_DEFAULT_STORE = None


def default_store():
    global _DEFAULT_STORE
    if _DEFAULT_STORE is None:
        _DEFAULT_STORE = VariableStore(seed=0)
    return _DEFAULT_STORE


def reset_default_store(seed=0):
    """tf.reset_default_graph() analogue."""
    global _DEFAULT_STORE
    _DEFAULT_STORE = VariableStore(seed=seed)
    return _DEFAULT_STORE
