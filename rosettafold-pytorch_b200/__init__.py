"""rosettafold-pytorch_b200 — B200-native (sm_100a) three-track trunk of rosettafold-pytorch.

`ops` wraps the C ABI of librfk.so (include/rfk.h); `modules` mirrors the reference's nn.Module
classes for the trunk (same constructors, forward signatures and state_dict keys).
"""
from . import _lib, ops  # noqa: F401

__version__ = "0.1.0"
from . import modules  # noqa: E402,F401
from .modules import (ColWise, EncoderLayer, FeedForward, MsaUpdateUsingSelfAttention,  # noqa: E402,F401
                      MsaUpdateWithPair, MsaUpdateWithPairAndCoord, MsaUpdateWithPairLayer, OuterProductMean,
                      PairUpdateWithAxialAttention, PairUpdateWithAxialAttentionLayer,
                      PairUpdateWithMsa, PerformerSelfAttention, PositionWiseWeightFactor, Residual,
                      RowWise, SoftTiedAttentionOverResidues, Symmetrization, TrunkBlocks,
                      TwoTrackBlock, get_mode, load_reference_weights, set_bounded_operand_dtype, set_mode)
from .embeddings import (MsaEmbedding, PairEmbedding, SinusoidalPositionalEncoding,  # noqa: E402,F401
                         SinusoidalPositionalEncoding2D)
from . import graph  # noqa: E402,F401
from .graph import GraphTransformer, GraphTransformerBlock  # noqa: E402,F401
from . import heads  # noqa: E402,F401
from .heads import PredictionHead, ResBlock2D, ResNet  # noqa: E402,F401
from . import replicas  # noqa: E402,F401
from .integration import accelerate, accelerate_block  # noqa: E402,F401
from .graphs import GraphedModule  # noqa: E402,F401
from . import sharded  # noqa: E402,F401
from .sharded import ShardedPairAxialAttention, ShardedTrunkBlocks, ShardedTwoTrackBlock  # noqa: E402,F401
