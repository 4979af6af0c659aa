"""Shared test helpers: move variables between the product's VariableStore and the oracle."""
import numpy as np
import torch

from tensorflow_yolo2_b200.engine import create_classifier_variables, create_variables
from tensorflow_yolo2_b200.variables import VariableStore, _to_numpy


def make_store(output_filter, seed=0, tame=False, passthrough=False, classifier=False):
    """Variables in the reference's order/naming.  tame=True rescales W to He-init magnitude
    (sqrt(2/fan_in)) so activations stay O(1) instead of exploding to 1e9+."""
    st = VariableStore(seed=seed)
    layers = (create_classifier_variables(st, output_filter) if classifier      # darknet19 classifier: 18 core layers + logits conv
              else create_variables(st, output_filter, passthrough=passthrough))
    if tame:
        rs = np.random.RandomState(seed + 1)
        for L in layers:
            fan_in = L['k'] * L['k'] * L['cin']
            st[L['W']] = (st[L['W']] * (np.sqrt(2.0 / fan_in) / 0.1)).astype(np.float32)
            st[L['b']] = (rs.randn(L['cout']) * 0.1).astype(np.float32)
            bn = L['bn']
            st[bn['gamma']] = rs.uniform(0.5, 1.5, L['cout']).astype(np.float32) * np.where(rs.rand(L['cout']) < 0.2, -1, 1).astype(np.float32)
            st[bn['beta']] = (rs.randn(L['cout']) * 0.2).astype(np.float32)
            st[bn['moving_mean']] = (rs.randn(L['cout']) * 0.2).astype(np.float32)
            st[bn['moving_variance']] = rs.uniform(0.5, 2.0, L['cout']).astype(np.float32)
    return st, layers


def oracle_params(st, layers, with_passthrough=False):
    """-> (core, head) parameter lists for the oracle; with_passthrough=True -> (core, head, passthrough_params)."""
    core, head, pt = [], [], None
    for L in layers:
        bn = L['bn']
        g = lambda n: torch.tensor(_to_numpy(st[n]))
        p = dict(W=g(L['W']), b=g(L['b']), gamma=g(bn['gamma']), beta=g(bn['beta']), mm=g(bn['moving_mean']),
                 mv=g(bn['moving_variance']))
        if L.get('role') == 'passthrough':
            pt = p
        else:
            (head if L['head'] else core).append(p)
    return (core, head, pt) if with_passthrough else (core, head)


def rel_l2(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))
