"""``MCTS`` facade (alphazero/mcts.py:86-183) over the CUDA node-table kernels.

``MCTS(policy_value_fn, num_simulations, c_puct, dirichlet_alpha=0.3, dirichlet_weight=0.25)``
and ``.run(root_state, temperature, add_root_noise=False) -> (pi ndarray (B,B) f32, (r,c) | None)``
keep the reference signatures.  Selection, expansion and backup always run on the GPU;
``policy_value_fn`` is the operator-level hook (mcts.py:88-93):

* a function made by ``controller.make_policy_value_fn`` is recognised and the whole search runs
  on the device with the engine's own network kernels;
* any other Python callable is called once per simulation on the leaf the kernels selected
  (exactly the reference's evaluation order), which is how the parity tests inject identical
  evaluator numbers into the reference, the oracle and this engine.

``run_batch`` is the batched addition: many roots at once, and it also returns the raw root
visit counts the reference computes (mcts.py:144) but never exposes."""

from __future__ import annotations

import numpy as np
import torch

from . import engine as _e
from ._lib import DTYPE_F32
from .games import Gomoku, GomokuBatch, _CODE, _SYM


def _state_from_cells(cells, to_move, last, board, win):
    g = Gomoku(board, win)
    g.board = [[_SYM[int(v)] for v in row] for row in cells]
    g.current_player = _SYM[int(to_move)]
    g.last_action = None if last < 0 else (int(last) // board, int(last) % board)
    return g


class MCTS:
    def __init__(self, policy_value_fn, num_simulations, c_puct, dirichlet_alpha=0.3, dirichlet_weight=0.25,
                 seed=None, net_dtype=DTYPE_F32, max_cached_engines=4):
        self.policy_value_fn = policy_value_fn
        self.num_simulations = num_simulations
        self.c_puct = c_puct
        self.dirichlet_alpha = dirichlet_alpha
        self.dirichlet_weight = dirichlet_weight
        # the reference draws its noise / moves from the global NumPy RNG (mcts.py:114, 177): without an explicit seed
        # the Philox key comes from that RNG, so every MCTS object gets its own stream and np.random.seed() reproduces it
        self.seed = int(np.random.randint(0, 2 ** 31 - 1)) if seed is None else seed
        self.net_dtype = net_dtype
        self._engines = {}                  # (board, win, n, architecture) -> Engine, least recently used first
        self._max_engines = max(1, max_cached_engines)
        self._calls = 0

    def _engine(self, board, win, n):
        ctrl = getattr(self.policy_value_fn, "controller", None)
        net = ctrl.net if ctrl is not None else None           # read at call time
        kw = {}
        if net is not None:
            from .controller import net_kind_of
            from .net import GomokuTowerNet
            if isinstance(net, GomokuTowerNet):
                kw = dict(net_kind=net_kind_of(net), tower_blocks=net.blocks, tower_channels=net.channels)
        key = (board, win, n, tuple(sorted(kw.items())))
        eng = self._engines.pop(key, None)
        if eng is None:
            while len(self._engines) >= self._max_engines:     # each engine owns a node table: close the stalest
                self._engines.pop(next(iter(self._engines))).close()
            eng = _e.Engine(board=board, win_length=win, games=n, sims=self.num_simulations, c_puct=self.c_puct,
                            dirichlet_alpha=self.dirichlet_alpha, dirichlet_weight=self.dirichlet_weight,
                            seed=self.seed, net_dtype=self.net_dtype, **kw)
            eng.weights_version = None
        self._engines[key] = eng                               # most recently used last
        if ctrl is not None:
            from .controller import weights_fingerprint
            fp = weights_fingerprint(ctrl)                     # the reference reads controller.net at every evaluation
            if eng.weights_version != fp:
                eng.load_state_dict(net.state_dict())
                eng.weights_version = fp
        return eng

    def close(self):
        for eng in self._engines.values():
            eng.close()
        self._engines.clear()

    # ---- batched search ------------------------------------------------------------------------------
    def run_batch(self, states, temperature, add_root_noise=False, sample=True):
        """states: list of Gomoku (same size).  Returns (pi [n,B,B] f32 ndarray, actions [n,2] int
        ndarray (-1 where no move exists), visits [n,A] int32 ndarray)."""
        n = len(states)
        board, win = states[0].board_size, states[0].win_length
        eng = self._engine(board, win, n)
        dev = eng.device
        cells = torch.tensor([[[_CODE[c] for c in row] for row in s.board] for s in states], dtype=torch.int8)
        to_move = [_CODE[s.current_player] for s in states]
        last = [-1 if s.last_action is None else s.last_action[0] * board + s.last_action[1] for s in states]
        batch = GomokuBatch.from_cells(cells, to_move=to_move, last=last, win_length=win, device=dev)
        self._calls += 1
        ply = torch.full((n,), self._calls, dtype=torch.int32, device=dev)   # distinct noise per call
        fused = getattr(self.policy_value_fn, "controller", None) is not None
        if fused:
            visits = eng.search(batch.x, batch.o, batch.to_move, batch.last, ply=ply, add_noise=add_root_noise)
        else:
            visits = self._search_with_hook(eng, batch, states, ply, add_root_noise)
        legal = batch.legal_mask().reshape(n, -1)
        T = np.broadcast_to(np.asarray(temperature, dtype=np.float64), (n,)).copy()
        pi = _e.extract_policy(visits, legal, torch.from_numpy(T), board).cpu().numpy()
        visits = visits.cpu().numpy()
        legal = legal.cpu().numpy().astype(bool)
        actions = np.full((n, 2), -1, dtype=np.int64)
        if sample:
            for i in range(n):
                idx = np.flatnonzero(legal[i])
                if idx.size == 0:
                    continue
                p = pi[i, idx].astype(np.float64)
                p /= p.sum()
                a = idx[np.random.choice(idx.size, p=p)]        # always sampled, mcts.py:177
                actions[i] = (a // board, a % board)
        return pi.reshape(n, board, board), actions, visits

    def _search_with_hook(self, eng, batch, states, ply, add_root_noise):
        n, board, win = len(states), batch.board_size, batch.win_length
        A = board * board
        dev = eng.device
        eng.search_begin(batch.x, batch.o, batch.to_move, batch.last, ply=ply)
        pol = np.zeros((n, A), dtype=np.float32)
        val = np.zeros((n,), dtype=np.float64)                  # Python floats stay float64 (mcts.py:137-141)
        for i, s in enumerate(states):                          # root evaluation (value discarded)
            p, _ = self.policy_value_fn(s)
            pol[i] = np.asarray(p, dtype=np.float32).reshape(-1)
        eng.root_expand(torch.from_numpy(pol).to(dev), add_noise=add_root_noise)
        for _ in range(self.num_simulations):
            eng.select()
            x, o, tm, last, kind = eng.leaf_states()
            kind_h = kind.cpu().numpy()
            need = np.flatnonzero(kind_h == 0)
            if need.size:
                cells = _e.unpack_boards(x, o, board).view(n, board, board).cpu().numpy()
                tm_h, last_h = tm.cpu().numpy(), last.cpu().numpy()
                for i in need:
                    leaf = _state_from_cells(cells[i], tm_h[i], last_h[i], board, win)
                    p, v = self.policy_value_fn(leaf)
                    pol[i] = np.asarray(p, dtype=np.float32).reshape(-1)
                    val[i] = v
            eng.expand_backup(torch.from_numpy(pol).to(dev), torch.from_numpy(val).to(dev))
        return eng.root_visits()

    # ---- reference signature ----------------------------------------------------------------------------
    def run(self, root_state, temperature, add_root_noise=False):
        pi, actions, _ = self.run_batch([root_state], temperature, add_root_noise)
        if actions[0, 0] < 0:
            b = root_state.board_size                      # no legal move: mcts.py:152-153
            return torch.zeros((b, b), dtype=torch.float32), None
        return pi[0], (int(actions[0, 0]), int(actions[0, 1]))
