"""``SelfPlayManager`` (alphazero/self_play.py:80-159): same constructor and
``generate_self_play(num_games, num_workers=None, flatten=False)``, but the games run as
concurrent lockstep MCTS episodes on the GPU instead of one Python process per game."""

from __future__ import annotations

from collections.abc import Callable

import numpy as np
import torch

from . import engine as _e
from ._lib import AUGMENT_D4, AUGMENT_MULT, AUGMENT_NONE, AUGMENT_REFERENCE, DTYPE_BF16, DTYPE_F32, MORE
from .constants import (SELF_PLAY_EXPLORATION_CONSTANT, TEMPERATURE_BASELINE, TEMPERATURE_SCHEDULE_HALFTIME,
                        WIN_LENGTH)

_AUG = {"reference": AUGMENT_REFERENCE, "d4": AUGMENT_D4, "none": AUGMENT_NONE}


def default_temperature_schedule(move: int) -> float:
    """self_play.py:24-26."""
    norm = 1 + TEMPERATURE_BASELINE
    return (np.exp(-move / TEMPERATURE_SCHEDULE_HALFTIME) + TEMPERATURE_BASELINE) / norm


class SelfPlayManager:
    def __init__(self, controller, device=None, mcts_params: dict = None,
                 temperature_schedule: Callable[[int], float] = default_temperature_schedule,
                 win_length: int = None, concurrent_games: int = 2048, augment: str = "d4",
                 net_dtype: int = DTYPE_F32, seed: int = 0, rank: int = 0, world: int = 1):
        self.controller = controller
        self.mcts_params = mcts_params or {"num_simulations": 100}
        self.temperature_schedule = temperature_schedule
        self.device = device
        self.board_size = controller.net.board_size
        self.win_length = win_length if win_length is not None else (WIN_LENGTH if self.board_size < 9 else 5)
        self.concurrent_games = concurrent_games
        self.augment = _AUG[augment] if isinstance(augment, str) else augment
        self.net_dtype = net_dtype
        self.seed, self.rank, self.world = seed, rank, world
        self._engine = None
        self._first_game = 0

    def engine(self, num_games=None):
        g = min(self.concurrent_games, max(1, -(-(num_games or self.concurrent_games) // self.world)))
        if self._engine is None or self._engine.games != g:
            from .controller import net_kind_of
            from .net import GomokuTowerNet
            net = self.controller.net
            kw = {}
            if isinstance(net, GomokuTowerNet):
                kw = dict(net_kind=net_kind_of(net), tower_blocks=net.blocks, tower_channels=net.channels)
            p = self.mcts_params
            self._engine = _e.Engine(
                board=self.board_size, win_length=self.win_length, games=g,
                sims=p["num_simulations"], c_puct=p.get("c_puct", SELF_PLAY_EXPLORATION_CONSTANT),
                dirichlet_alpha=p.get("dirichlet_alpha", 0.3), dirichlet_weight=p.get("dirichlet_weight", 0.25),
                augment=self.augment, net_dtype=self.net_dtype, seed=self.seed, rank=self.rank, world=self.world,
                **kw)
            a = self.board_size ** 2
            self._engine.set_temperature_table([float(self.temperature_schedule(m)) for m in range(a + 1)])
        return self._engine

    def generate_device(self, num_games: int):
        """Fast path: examples stay on the GPU.  Yields (planes[n,4,B,B], pi[n,B,B], z[n]) slabs."""
        eng = self.engine(num_games)
        # weights travel as the controller's current state_dict (self_play.py:121)
        eng.load_state_dict(self.controller.net.state_dict())
        resume = False
        out = None
        while True:
            st, planes, pi, z, written = eng.selfplay(num_games, first_game=self._first_game, out=out, resume=resume)
            out = (planes, pi, z)
            if written:
                yield planes[:written].clone(), pi[:written].clone(), z[:written].clone()
            if st != MORE:
                break
            resume = True
        self._first_game += num_games

    def generate_to_ring(self, num_games: int, ring, slab_positions: int = None):
        """Multi-GPU data path (replaces the result queue of self_play.py:140-148): this rank
        plays its share of ``num_games``; every move step the RAW positions leave in one fixed-capacity slab that is
        all-gathered over the ranks on a side stream (NCCL; a device copy in a single process), and every rank expands
        everybody's positions by ``self.augment`` straight into ``ring`` (a ``ReplayBuffer``) -- identical rings on all
        ranks, no host synchronisation besides the driver's own.  Returns the number of examples appended."""
        from .dist import ExampleExchange
        eng = self.engine(num_games)
        eng.load_state_dict(self.controller.net.state_dict())
        eng.set_augment(AUGMENT_NONE)                       # raw positions on the wire, augmentation on the receiving side
        if slab_positions:
            eng.set_slab_capacity(slab_positions)
        ex = ExampleExchange(self.board_size, eng.example_capacity_floor(), eng.device)
        before = ring.sync_head()
        step, st = 0, MORE
        with torch.cuda.stream(eng._driver_stream()):
            while st == MORE:
                k = step & 1
                ex.wait_slab(k)
                st, _, _, _, w = eng.selfplay(num_games, first_game=self._first_game, out=ex.out_views(k),
                                              max_move_steps=1, resume=step > 0)
                ex.exchange(k, w, ring=ring, augment=self.augment)
                step += 1
            ex.wait()
        torch.cuda.synchronize(eng.device)
        eng.set_augment(self.augment)
        self._first_game += num_games
        return ring.sync_head() - before

    def generate_self_play(self, num_games: int, num_workers: int = None, flatten=False) -> list:
        """Returns the reference's flat list of ``(state Tensor f32 (4,B,B), pi ndarray f32 (B,B), z int)``.
        ``num_workers`` and ``flatten`` are accepted and ignored (the reference ignores ``flatten`` too)."""
        results = []
        for planes, pi, z in self.generate_device(num_games):
            planes_h, pi_h, z_h = planes.cpu(), pi.cpu().numpy(), z.cpu().numpy()
            for i in range(planes_h.shape[0]):
                results.append((planes_h[i].clone(), pi_h[i].copy(), int(z_h[i])))
        print(f"[SelfPlayManager] Collected {len(results)} games.")
        return results
