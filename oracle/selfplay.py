"""Oracle: self-play episode loop, labelling and augmentation (test
infrastructure only -- see oracle/__init__.py).  Restates
``alphazero/self_play.py``."""

from __future__ import annotations

import numpy as np

from . import gomoku
from .mcts import Search


def default_temperature_schedule(move: int, halftime=100, baseline=0.01):
    """alphazero/self_play.py:24-26 with alphazero/constants.py:20-21:
    (exp(-move/100) + 0.01) / 1.01, an np.float64."""
    return (np.exp(-move / halftime) + baseline) / (1 + baseline)


def label(result: int, player: int) -> int:
    """alphazero/self_play.py:71 (pinned by alphazero/tests/tests.py:11-22):
    0 on a draw, +1 if the player to move at that position won, else -1."""
    return 0 if result == gomoku.DRAW else (1 if player == result else -1)


def augment_reference(planes: np.ndarray, pi: np.ndarray):
    """alphazero/self_play.py:94-108 AS WRITTEN: four examples; the state is
    rotated k times counter-clockwise (k=0..3) but pi is ALWAYS rotated exactly
    once (``np.rot90(policy_matrix)`` ignores k), so only k=1 is self-consistent.
    CCW convention: out[i,j] = in[j, B-1-i]."""
    out = []
    for k in range(4):
        out.append((np.rot90(planes, k, axes=(1, 2)).copy(), np.rot90(pi).copy()))
    return out


def augment_d4(planes: np.ndarray, pi: np.ndarray):
    """The intended 8-fold dihedral augmentation (README.md:61,82 promises ~8x):
    element s = k + 4*f applies k CCW quarter turns, then (f=1) a left-right
    flip, to state and pi alike."""
    out = []
    for f in range(2):
        for k in range(4):
            s = np.rot90(planes, k, axes=(1, 2))
            p = np.rot90(pi, k)
            if f:
                s, p = s[:, :, ::-1], p[:, ::-1]
            out.append((np.ascontiguousarray(s), np.ascontiguousarray(p)))
    return out


def play_episode(evaluator, board, win, num_simulations, c_puct,
                 temperature_schedule=default_temperature_schedule, rng=None,
                 dirichlet_alpha=0.3, dirichlet_weight=0.25, add_root_noise=True,
                 max_plies=None):
    """alphazero/self_play.py:47-72 -- one game: until terminal, search with
    root noise at T(ply), record (planes, pi, player to move), play the sampled
    move; then label every recorded position.  The final (terminal) position is
    not recorded.  Returns (examples, result, moves)."""
    search = Search(evaluator, num_simulations, c_puct, dirichlet_alpha, dirichlet_weight, rng=rng)
    state = gomoku.State(board, win)
    history, moves = [], []
    ply = 0
    while not state.is_terminal():
        if max_plies is not None and ply >= max_plies:
            break
        res = search.run(state, temperature_schedule(ply), add_root_noise=add_root_noise)
        history.append((state.planes(), res.pi, state.to_move))
        moves.append(res.action)
        state = state.play(res.action)
        ply += 1
    result = state.result()
    examples = [(s, p, label(result, player)) for s, p, player in history]
    return examples, result, moves


def generate_self_play(evaluator, num_games, board, win, num_simulations, c_puct,
                       augment="reference", **kw):
    """alphazero/self_play.py:110-159 minus the process fan-out: flat list of
    (planes float32 (4,B,B), pi float32 (B,B), z int), 4 (reference) or 8 (d4)
    per recorded position."""
    aug = augment_reference if augment == "reference" else augment_d4
    out = []
    for _ in range(num_games):
        examples, _, _ = play_episode(evaluator, board, win, num_simulations, c_puct, **kw)
        for planes, pi, z in examples:
            for s, p in aug(planes, pi):
                out.append((s, p, z))
    return out
