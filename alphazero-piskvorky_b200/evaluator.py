"""``ModelEvaluator`` (alphazero/evaluator.py:22-122) and the promotion rule
(alphazero/promoter.py:47): the whole match runs as lockstep games on the GPU with two
weight slots."""

from __future__ import annotations

import numpy as np
import torch

from . import engine as _e
from ._lib import DTYPE_F32
from .constants import (EVAL_EXPLORATION_CONSTANT, EVAL_TEMPERATURE, EVAL_TEMPERATURE_SCHEDULE_HALFTIME,
                        NUM_EVAL_SIMULATIONS, PROMOTION_THRESHOLD, WIN_LENGTH)


def temperature_schedule(move: int) -> float:
    """evaluator.py:14-19."""
    return EVAL_TEMPERATURE * np.exp(-move / EVAL_TEMPERATURE_SCHEDULE_HALFTIME)


def should_promote(win_rate: float, threshold: float = PROMOTION_THRESHOLD) -> bool:
    """promoter.py:47 -- strictly greater."""
    return win_rate > threshold


class ModelEvaluator:
    def __init__(self, game_class=None, print_games=False, device=None, num_simulations=NUM_EVAL_SIMULATIONS,
                 c_puct=EVAL_EXPLORATION_CONSTANT, win_length=None, net_dtype=DTYPE_F32, seed=0, rank=0, world=1):
        self.game_class = game_class
        self.print_games = print_games
        self.device = device
        self.num_simulations, self.c_puct = num_simulations, c_puct
        self.win_length = win_length
        self.net_dtype, self.seed, self.rank, self.world = net_dtype, seed, rank, world
        self._first_game = 0                # global game ids already played: every match gets fresh Philox streams

    def evaluate(self, candidate_controller, baseline_controller, num_games=20, debug=False):
        board = candidate_controller.net.board_size
        win = self.win_length if self.win_length is not None else (WIN_LENGTH if board < 9 else 5)
        local = -(-num_games // self.world)
        from .controller import net_kind_of
        from .net import GomokuTowerNet
        cand, base = candidate_controller.net, baseline_controller.net
        kw = {}
        if isinstance(cand, GomokuTowerNet) or isinstance(base, GomokuTowerNet):
            if type(cand) is not type(base) or (cand.blocks, cand.channels) != (base.blocks, base.channels):
                raise ValueError("candidate and baseline must share one architecture (two weight slots of one engine)")
            kw = dict(net_kind=net_kind_of(cand), tower_blocks=cand.blocks, tower_channels=cand.channels)
        if base.board_size != board:
            raise ValueError("candidate and baseline play on different board sizes")
        eng = _e.Engine(board=board, win_length=win, games=max(1, local), sims=self.num_simulations,
                        c_puct=self.c_puct, net_dtype=self.net_dtype, seed=self.seed, rank=self.rank,
                        world=self.world, **kw)
        eng.load_state_dict(cand.state_dict(), slot=0)
        eng.load_state_dict(base.state_dict(), slot=1)
        eng.set_temperature_table([float(temperature_schedule(s)) for s in range(board * board + 1)])
        # game parity decides the first mover (evaluator.py:64-69): keep first_game even so game i of this match
        # starts like game i of the reference's loop
        wins, losses, draws = eng.arena(num_games, first_game=self._first_game)
        self._first_game += num_games + (num_games & 1)
        eng.close()
        if self.world > 1:
            from .dist import all_reduce_wld
            wins, losses, draws = all_reduce_wld(wins, losses, draws)
        total = wins + losses + draws
        win_rate = (wins + 0.5 * draws) / total if total else 0.0
        return win_rate, {"wins": wins, "losses": losses, "draws": draws, "total": total, "win_rate": win_rate}
