"""B200-native (sm_100a) Graph WaveNet block: a drop-in for the `gwnet` module of
aaparcedo/multimodal_outage (models/graph_wavenet.py), backed by hand-written CUDA kernels
behind a plain C ABI (include/gwn.h)."""
from .graph_wavenet import gcn, gwnet, linear, nconv  # noqa: F401
from .supports import asym_adj, double_transition, load_adj  # noqa: F401

__all__ = ['gwnet', 'gcn', 'linear', 'nconv', 'asym_adj', 'double_transition', 'load_adj']
