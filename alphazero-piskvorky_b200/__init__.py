"""alphazero-piskvorky_b200: B200-native batched MCTS self-play engine behind the
AlphaZero-Piskvorky API (Gomoku / MCTS / SelfPlayManager / ReplayBuffer / GomokuNet).

The directory name is fixed by the build contract and is not a valid Python identifier;
import it as ``alphazero_piskvorky_b200`` (a one-file alias package at the repo root)."""

from . import _lib
from ._lib import (AUGMENT_D4, AUGMENT_NONE, AUGMENT_REFERENCE, DTYPE_BF16, DTYPE_F32, MORE, NET_STEM, NET_TOWER,
                   EngineError)
from .constants import DRAW, O, X
from .controller import NeuralNetworkController, make_policy_value_fn
from .engine import Engine, flatten_state_dict
from .evaluator import ModelEvaluator, should_promote
from .games import Gomoku, GomokuBatch
from .mcts import MCTS
from .net import GomokuNet, GomokuTowerNet
from .replay_buffer import ReplayBuffer
from .self_play import SelfPlayManager, default_temperature_schedule
from .trainer import Trainer, adamw_step

__all__ = [
    "Engine", "EngineError", "Gomoku", "GomokuBatch", "MCTS", "GomokuNet", "GomokuTowerNet", "SelfPlayManager",
    "ReplayBuffer", "ModelEvaluator", "NeuralNetworkController", "make_policy_value_fn",
    "default_temperature_schedule", "flatten_state_dict", "should_promote", "Trainer", "adamw_step", "X", "O", "DRAW",
]
