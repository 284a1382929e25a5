"""CPU restatement of the reference's support construction.  TEST INFRASTRUCTURE ONLY.

``asym_adj`` (``/root/reference/utils.py:152-158``) builds the row-normalised
transition matrix ``D^-1 A`` through scipy sparse arithmetic.  This file restates
the same arithmetic densely in numpy, in the *same floating-point order*
(row sums in the input dtype, reciprocal via ``power(-1)``, inf -> 0, scale each
row, cast to float32 last) so the result is bit-identical; pinned by
``tests/golden/asym_adj_*.npz`` (reference-executed).
"""
from __future__ import annotations

import numpy as np


def asym_adj_dense(adj: np.ndarray) -> np.ndarray:
    """Bit-exact dense restatement of utils.py:152-158.  Integer input raises
    ``ValueError`` exactly as the reference does (``np.power(int, -1)``)."""
    adj = np.asarray(adj)
    rowsum = adj.sum(axis=1).flatten()
    with np.errstate(divide='ignore'):
        d_inv = np.power(rowsum, -1).flatten()
    d_inv[np.isinf(d_inv)] = 0.0
    # sp.diags(d_inv).dot(adj): each stored entry a[i,j] becomes d_inv[i]*a[i,j]
    # in the common dtype, explicit zeros stay zero; then cast and densify.
    out = d_inv[:, None] * adj
    return np.asarray(out.astype(np.float32))


def double_transition(adj: np.ndarray):
    """Upstream 'doubletransition' supports used by BASELINE configs 2-5:
    [asym_adj(A), asym_adj(A^T)]."""
    a = np.asarray(adj, dtype=np.float32)
    return [asym_adj_dense(a), asym_adj_dense(a.T.copy())]


def reference_load_adj_identity(n: int):
    """What the reference's own ``load_adj(..., 'doubletransition')`` returns:
    a single identity matrix (graph_wavenet.py:23-24)."""
    return [np.diag(np.ones(n)).astype(np.float32)]


def synthetic_knn_graph(n: int, k: int = 6, seed: int = 42) -> np.ndarray:
    """Seeded county-like graph: n uniform points in the unit square,
    symmetrised k-nearest-neighbour, 0/1 int64, zero diagonal."""
    rng = np.random.default_rng(seed)
    pts = rng.random((n, 2))
    adj = np.zeros((n, n), dtype=np.int64)
    # blockwise distances keep memory bounded at n=3100
    for s in range(0, n, 512):
        d = ((pts[s:s + 512, None, :] - pts[None, :, :]) ** 2).sum(-1)
        d[np.arange(d.shape[0]), np.arange(s, s + d.shape[0])] = np.inf
        nn = np.argpartition(d, k, axis=1)[:, :k]
        rows = np.repeat(np.arange(s, s + d.shape[0]), k)
        adj[rows, nn.reshape(-1)] = 1
    adj = np.maximum(adj, adj.T)
    return adj


def synthetic_directed_graph(n: int, p: float = 0.07, seed: int = 7) -> np.ndarray:
    """Seeded *directed* 0/1 graph so forward != backward transition matrices."""
    rng = np.random.default_rng(seed)
    adj = (rng.random((n, n)) < p).astype(np.int64)
    np.fill_diagonal(adj, 0)
    adj[0, :] = 0          # one isolated-out node: exercises the inf -> 0 branch
    return adj
