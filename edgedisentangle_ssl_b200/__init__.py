"""edgedisentangle_ssl_b200 -- B200-native DISGAT message-passing hot path.

Drop-in for the DISGAT layers / model / SSL loss heads of TianxiangZhao/EdgeDisentangle_SSL
behind the reference's own PyTorch module surface.  All sparse work runs in libedis.so
(hand-written sm_100a CUDA behind the C ABI of include/edis.h); importing this package
fails loudly if that library has not been built -- there is no CPU fallback.
"""
from . import _lib  # noqa: F401  (raises ImportError if libedis.so is missing)
from .graph import Graph, build_adjacency, as_graph  # noqa: F401
from .layers import DisGALayer, FuseLayer, SageConv, GraphConvolution, run_channels  # noqa: F401
from .models import DISGAT, MLP  # noqa: F401
from .baselines import GraphAttentionLayer, DisentangleLayer  # noqa: F401

__version__ = "0.1"
