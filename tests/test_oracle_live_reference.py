"""The oracle against the UNMODIFIED reference, run live (build container only: needs /root/reference).

The committed fixtures pin the oracle on fixed inputs; here the same comparisons run on inputs the fixtures do not
contain (other seeds, a 7x7 board, WIN 3..5), with the reference's own code doing the computing each time.  Skipped
wherever the reference tree is absent (the GPU box), so nothing in the `-m gpu` tests depends on it."""

import contextlib
import io
import os
import sys

import numpy as np
import pytest
import torch

import synth
from oracle import gomoku, mcts as omcts, selfplay as oselfplay

REFERENCE = "/root/reference/alphazero"
pytestmark = pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="the reference tree only exists in the build container")

CODE = {None: 0, "X": 1, "O": 2, "D": 3}
SYM = {0: None, 1: "X", 2: "O"}


@pytest.fixture(scope="module")
def ref():
    """The reference's flat modules (games, mcts, self_play), imported under private names so they cannot shadow ours."""
    import importlib.util
    mods = {}
    saved_path, saved_mods = list(sys.path), {k: sys.modules.get(k) for k in ("games", "mcts", "constants", "net",
                                                                                "controller", "self_play")}
    sys.path.insert(0, REFERENCE)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            for name in ("games", "mcts", "self_play"):
                mods[name] = importlib.import_module(name)
    finally:
        sys.path[:] = saved_path
        for k, v in saved_mods.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v

    class Ref:
        games, mcts, self_play = mods["games"], mods["mcts"], mods["self_play"]
    return Ref


def _ref_game(ref, board, win, to_move=1, last=None):
    g = ref.games.Gomoku(int(board.shape[0]), win)
    g.board = [[SYM[int(v)] for v in row] for row in board]
    g.current_player = SYM[int(to_move)]
    g.last_action = last
    return g


@pytest.mark.parametrize("size,win", [(3, 3), (5, 4), (7, 4), (7, 5), (9, 5), (15, 5)])
def test_rules_on_random_boards(ref, size, win):
    """games.py:133-179, 212-227 on ARBITRARY boards (both colours may hold lines, overlines, full boards): winner,
    draw, legal moves, encoding."""
    rs = np.random.RandomState(100 * size + win)
    n = 250 if size <= 9 else 80
    for i in range(n):
        density = rs.uniform(0.0, 1.05)
        board = ((rs.random_sample((size, size)) < density) * rs.randint(1, 3, size=(size, size))).astype(np.int8)
        tm = int(rs.randint(1, 3))
        occupied = np.argwhere(board != 0)
        last = tuple(int(v) for v in occupied[rs.randint(len(occupied))]) if len(occupied) and rs.rand() < 0.8 else None
        g = _ref_game(ref, board, win, tm, last)
        s = gomoku.State.from_board(board, to_move=tm, win=win, last=last)
        terminal = g.is_terminal()
        assert s.is_terminal() == terminal, i
        assert s.result() == CODE[g.get_game_result()], i
        assert gomoku.result_vectorised(board, win) == CODE[g.get_game_result()], i
        assert s.legal_actions() == g.get_legal_actions(), i
        assert np.array_equal(s.planes(), g.encode("cpu").numpy()), i


@pytest.mark.parametrize("size,win", [(5, 4), (7, 4), (9, 5)])
def test_playouts_move_by_move(ref, size, win):
    """apply_action / is_terminal along whole random games, including the ValueError on an occupied cell."""
    for seed in range(12):
        rs = np.random.RandomState(7000 + 31 * size + seed)
        g, s = ref.games.Gomoku(size, win), gomoku.State(size, win)
        while not g.is_terminal():
            assert not s.is_terminal()
            acts = g.get_legal_actions()
            assert acts == s.legal_actions()
            a = acts[rs.randint(len(acts))]
            g, s = g.apply_action(a), s.play(a)
            assert np.array_equal(s.planes(), g.encode("cpu").numpy())
            assert s.to_move == CODE[g.current_player] and s.last == g.last_action
            with pytest.raises(ValueError):
                g.apply_action(a)
            with pytest.raises(ValueError):
                s.play(a)
        assert s.is_terminal() and s.result() == CODE[g.get_game_result()]


class _RootTap:
    """mcts.Node swapped for a recording subclass while the reference's own MCTS.run executes (as make_golden.py does)."""

    def __init__(self, ref):
        self.ref = ref

    def __enter__(self):
        tap = self
        tap.nodes = []
        self.base = base = self.ref.mcts.Node

        class Tapped(base):
            def __init__(self, *a, **kw):
                super().__init__(*a, **kw)
                tap.nodes.append(self)

        self.ref.mcts.Node = Tapped
        return self

    def __exit__(self, *exc):
        self.ref.mcts.Node = self.base

    def root_visits(self, size):
        v = np.zeros((size, size), dtype=np.int64)
        for (r, c), child in self.nodes[0].children.items():
            v[r, c] = child.N
        return v


@pytest.mark.parametrize("size,win,sims,roots", [(5, 4, 150, 10), (7, 4, 200, 8), (7, 5, 120, 6), (9, 5, 250, 5)])
def test_mcts_visit_counts_on_fresh_roots(ref, size, win, sims, roots):
    """mcts.py:101-183 on roots and evaluator salts the fixtures do not contain: visit counts exact, pi <= 1e-6, and --
    with the global NumPy RNG seeded identically -- the same sampled action."""
    for j in range(roots):
        rs = np.random.RandomState(9000 + 17 * size + j)
        g, s = ref.games.Gomoku(size, win), gomoku.State(size, win)
        for _ in range(rs.randint(0, size * size // 3)):
            acts = g.get_legal_actions()
            a = acts[rs.randint(len(acts))]
            if g.apply_action(a).is_terminal():
                break
            g, s = g.apply_action(a), s.play(a)
        sharp = (2.0, 4.0, 7.0)[j % 3]
        salt = 500 + j
        T = ref.self_play.default_temperature_schedule(3 * j) if j % 2 == 0 else 0.4
        np.random.seed(123 + j)
        with _RootTap(ref) as tap:
            pi, action = ref.mcts.MCTS(synth.reference_evaluator(salt, sharp), sims, 2.0).run(g, T, add_root_noise=False)
            visits = tap.root_visits(size)
        np.random.seed(123 + j)
        res = omcts.Search(synth.oracle_evaluator(salt, sharp), sims, 2.0).run(s, T, add_root_noise=False)
        assert np.array_equal(res.visits, visits), j
        assert np.abs(res.pi - np.asarray(pi)).max() <= 1e-6, j
        assert res.action == tuple(action), j


def test_augment_reference_mode_live(ref):
    """self_play.py:94-108 as written (state rotated k times, pi always rotated once)."""
    rs = np.random.RandomState(5)
    mgr = ref.self_play.SelfPlayManager.__new__(ref.self_play.SelfPlayManager)
    for size in (3, 5, 8, 15):
        planes = rs.random_sample((4, size, size)).astype(np.float32)
        pi = rs.random_sample((size, size)).astype(np.float32)
        got = oselfplay.augment_reference(planes, pi)
        want = mgr._augment_symmetries(torch.from_numpy(planes), pi)
        assert len(got) == len(want) == 4
        for (gs, gp), (ws, wp) in zip(got, want):
            assert np.array_equal(np.asarray(gs), ws.numpy()) and np.array_equal(np.asarray(gp), wp)
