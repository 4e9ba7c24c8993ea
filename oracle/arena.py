"""Oracle: arena match between two evaluators (test infrastructure only).
Restates ``alphazero/evaluator.py:14-19, 38-122`` and the promotion rule
``alphazero/promoter.py:47``."""

from __future__ import annotations

import numpy as np

from . import gomoku
from .mcts import Search


def eval_temperature(step: int, t0=0.3, halftime=4):
    """alphazero/evaluator.py:14-19 with constants.py:34-35: 0.3*exp(-step/4)."""
    return t0 * np.exp(-step / halftime)


def step_counters(first_mover_is_candidate: bool, ply: int) -> int:
    """Temperature step used by the side moving at ``ply`` (0-based).

    alphazero/evaluator.py:71-90: both counters start at 0 and, after every
    move, the counter of the side that is ABOUT to move is incremented.  So the
    first mover searches with step 0,1,2,... and the second mover with 1,2,3,...
    (its counter is bumped before its first move)."""
    turn = ply // 2
    return turn if ply % 2 == 0 else turn + 1


def play_match(candidate_eval, baseline_eval, num_games, board, win, num_simulations,
               c_puct, rng=None):
    """Candidate plays X, baseline plays O; X moves first in even games, O in odd
    games (evaluator.py:64-69); no root noise; fresh search per move.  Returns
    (win_rate, dict) with win_rate = (W + 0.5*D) / total (evaluator.py:103-105)."""
    wins = losses = draws = 0
    for i in range(num_games):
        state = gomoku.State(board, win)
        state.to_move = gomoku.X if i % 2 == 0 else gomoku.O
        searches = {
            gomoku.X: Search(candidate_eval, num_simulations, c_puct, rng=rng),
            gomoku.O: Search(baseline_eval, num_simulations, c_puct, rng=rng),
        }
        steps = {gomoku.X: 0, gomoku.O: 0}
        while not state.is_terminal():
            side = state.to_move
            res = searches[side].run(state, eval_temperature(steps[side]), add_root_noise=False)
            state = state.play(res.action)
            steps[state.to_move] += 1
        result = state.result()
        if result == gomoku.X:
            wins += 1
        elif result == gomoku.O:
            losses += 1
        else:
            draws += 1
    total = wins + losses + draws
    win_rate = (wins + 0.5 * draws) / total
    return win_rate, {"wins": wins, "losses": losses, "draws": draws, "total": total,
                      "win_rate": win_rate}


def should_promote(win_rate: float, threshold: float = 0.55) -> bool:
    """alphazero/promoter.py:19,47 -- strictly greater than the threshold."""
    return win_rate > threshold
