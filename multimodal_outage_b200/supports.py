"""Support (transition-matrix) construction - host-side, numpy, bit-exact with the reference.

``asym_adj`` mirrors ``/root/reference/utils.py:152-158`` (row-normalised ``D^-1 A`` as float32);
it is one-off CPU work and stays on the host.  ``load_adj`` mirrors the *interface*
of ``utils.py:161-180`` / ``graph_wavenet.py:13-32`` but, unlike the reference (which returns a lone
identity matrix for "doubletransition"), offers the real transition matrices as well.
"""
from __future__ import annotations

from typing import List

import numpy as np


def asym_adj(adj) -> np.ndarray:
    """Row-normalised transition matrix D^-1 A, float32 (utils.py:152-158).

    Same floating-point order as the reference's sparse arithmetic: row sums in the input dtype,
    reciprocal through ``power(-1)`` (so integer input raises ``ValueError`` exactly like the
    reference), inf -> 0, row scaling, float32 cast last."""
    adj = np.asarray(adj)
    rowsum = adj.sum(axis=1).flatten()
    with np.errstate(divide='ignore'):
        d_inv = np.power(rowsum, -1).flatten()
    d_inv[np.isinf(d_inv)] = 0.0
    return np.asarray((d_inv[:, None] * adj).astype(np.float32))


def double_transition(adj) -> List[np.ndarray]:
    """[asym_adj(A), asym_adj(A^T)] - forward and backward diffusion supports."""
    a = np.asarray(adj, dtype=np.float32)
    return [asym_adj(a), asym_adj(np.ascontiguousarray(a.T))]


def load_adj(adj_mx, adjtype: str = 'doubletransition'):
    """``adj_mx``: a [V,V] array, or a path to a .csv (index column + header, as the reference's
    ``adj_mx_fl.csv``) / .npy file.  Returns ``(None, None, supports)`` like the reference's csv branch.

    adjtype: 'identity'  -> what the reference's ``load_adj(..., "doubletransition")`` actually
                            returns: ``[I]`` (graph_wavenet.py:23-24)
             'transition' -> ``[asym_adj(A)]``
             'doubletransition' -> ``[asym_adj(A), asym_adj(A^T)]`` (the upstream Graph-WaveNet meaning,
                            used by the BASELINE configs)
    """
    if isinstance(adj_mx, str):
        if adj_mx.endswith('.csv'):
            import pandas as pd
            adj_mx = pd.read_csv(adj_mx, index_col=0).values
        else:
            adj_mx = np.load(adj_mx)
    adj_mx = np.asarray(adj_mx)
    if adjtype == 'identity':
        adj = [np.diag(np.ones(adj_mx.shape[0])).astype(np.float32)]
    elif adjtype == 'transition':
        adj = [asym_adj(adj_mx.astype(np.float32))]
    elif adjtype == 'doubletransition':
        adj = double_transition(adj_mx)
    else:
        raise AssertionError('adj type not defined')
    return None, None, adj
