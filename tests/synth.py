"""Deterministic synthetic evaluator shared by the golden generator, the oracle
tests and the GPU parity tests.

It maps a position to (policy over ALL cells summing to 1, value in (-1,1))
through numpy's frozen ``RandomState`` stream keyed by a hash of the position,
so the reference, the oracle and the CUDA engine can all be fed bit-identical
evaluator numbers (hard part H2)."""

from __future__ import annotations

import numpy as np

_MASK = (1 << 64) - 1


def board_key(board: np.ndarray, to_move: int, salt: int = 0) -> int:
    """64-bit FNV-1a style hash of the cells (0 empty / 1 X / 2 O), side to move and salt."""
    h = 0xCBF29CE484222325 ^ (salt * 0x9E3779B97F4A7C15 & _MASK)
    for v in board.reshape(-1).tolist():
        h ^= v + 1
        h = (h * 0x100000001B3) & _MASK
    h ^= to_move + 7
    h = (h * 0x100000001B3) & _MASK
    return h


def evaluate_board(board: np.ndarray, to_move: int, salt: int = 0, sharp: float = 3.0):
    """(policy float32 (B,B), value float).  ``sharp`` scales the logits: larger
    => more peaked priors => deeper trees."""
    n = board.shape[0]
    key = board_key(board, to_move, salt)
    rs = np.random.RandomState(key & 0xFFFFFFFF)
    logits = (rs.random_sample(n * n) * 2.0 - 1.0) * sharp
    value = float(np.float32(rs.random_sample() * 1.8 - 0.9))
    e = np.exp(logits - logits.max())
    policy = (e / e.sum()).astype(np.float32).reshape(n, n)
    return policy, value


def _cells_of_reference(game) -> tuple[np.ndarray, int]:
    """Reference ``Gomoku`` (list of lists of 'X'/'O'/None) -> int8 board, side to move."""
    code = {None: 0, "X": 1, "O": 2}
    b = np.array([[code[c] for c in row] for row in game.board], dtype=np.int8)
    return b, code[game.current_player]


def reference_evaluator(salt: int = 0, sharp: float = 3.0, log=None):
    """policy_value_fn for the reference ``MCTS`` (alphazero/mcts.py:88-93)."""
    def fn(game):
        b, tm = _cells_of_reference(game)
        p, v = evaluate_board(b, tm, salt, sharp)
        if log is not None:
            log.append((b.copy(), tm))
        return p.copy(), v
    return fn


def oracle_evaluator(salt: int = 0, sharp: float = 3.0):
    """evaluator for ``oracle.mcts.Search``."""
    def fn(state):
        p, v = evaluate_board(state.board, state.to_move, salt, sharp)
        return p.copy(), v
    return fn


def random_playout(board: int, win: int, plies: int, seed: int, stop_at_terminal=True):
    """Seeded uniformly-random legal playout on an oracle State; returns the move list."""
    from oracle import gomoku
    rs = np.random.RandomState(seed)
    s = gomoku.State(board, win)
    moves = []
    for _ in range(plies):
        if stop_at_terminal and s.is_terminal():
            break
        acts = s.legal_actions()
        if not acts:
            break
        a = acts[rs.randint(len(acts))]
        moves.append(a)
        s = s.play(a)
    return moves
