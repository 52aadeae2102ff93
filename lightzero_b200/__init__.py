"""lightzero_b200 -- B200-native batched MuZero MCTS + model inference behind the reference's own
interfaces (lzero.mcts.tree_search.MuZeroMCTSCtree, lzero.mcts.ctree.ctree_muzero.mz_tree,
lzero.model.MuZeroModel).  All compute is hand-written sm_100a CUDA behind the C ABI of
include/lzb200.h; this package is the thin host side."""
from .mcts_ctree import MuZeroMCTSCtree  # noqa: F401
from .muzero_model import MuZeroModel, MZNetworkOutput  # noqa: F401
from .muzero_model_mlp import MuZeroModelMLP  # noqa: F401
from .scaling_transform import DiscreteSupport, InverseScalarTransform  # noqa: F401
from . import mz_tree  # noqa: F401

__all__ = ["MuZeroMCTSCtree", "MuZeroModel", "MuZeroModelMLP", "MZNetworkOutput", "DiscreteSupport",
           "InverseScalarTransform", "mz_tree"]
