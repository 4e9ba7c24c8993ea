"""Python handle on the CUDA engine (thin: allocation with torch, calls through the C ABI)."""

from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import (AUGMENT_D4, AUGMENT_MULT, AUGMENT_NONE, AUGMENT_REFERENCE, DTYPE_BF16, DTYPE_F32, MORE,
                   NET_STEM, NET_TOWER, check, ptr, stream_ptr)

STEM_KEYS = ("conv1", "conv2", "conv3", "policy_conv", "policy_fc", "value_conv", "value_fc1", "value_fc2")


def words(board: int) -> int:
    return (board * board + 63) // 64


def tower_keys(blocks: int):
    keys = ["conv.weight", "conv.bias"] + [f"bn.{f}" for f in ("weight", "bias", "running_mean", "running_var")]
    for i in range(1, blocks + 1):
        for j in (1, 2):
            keys.append(f"res{i}.conv{j}.weight")
            keys += [f"res{i}.bn{j}.{f}" for f in ("weight", "bias", "running_mean", "running_var")]
    keys += ["policy_conv.weight", "policy_conv.bias"]
    keys += [f"policy_bn.{f}" for f in ("weight", "bias", "running_mean", "running_var")]
    keys += ["policy_fc.weight", "policy_fc.bias", "value_conv.weight", "value_conv.bias"]
    keys += [f"value_bn.{f}" for f in ("weight", "bias", "running_mean", "running_var")]
    keys += ["value_fc1.weight", "value_fc1.bias", "value_fc2.weight", "value_fc2.bias"]
    return keys


def flatten_state_dict(state_dict, net_kind=NET_STEM, tower_blocks=0) -> np.ndarray:
    """Reference ``state_dict`` (net.py:37-53 key names, PyTorch layouts) -> the float32 blob
    ``azp_load_weights`` takes: tensors concatenated in key order."""
    if net_kind == NET_STEM:
        keys = [f"{k}.{f}" for k in STEM_KEYS for f in ("weight", "bias")]
    else:
        keys = tower_keys(tower_blocks)
    parts = []
    for k in keys:
        v = state_dict[k]
        if isinstance(v, torch.Tensor):
            v = v.detach().to("cpu", torch.float32).numpy()
        parts.append(np.ascontiguousarray(v, dtype=np.float32).reshape(-1))
    return np.concatenate(parts)


class Engine:
    """One engine per GPU per process (C ABI: azp_create / azp_destroy)."""

    def __init__(self, board=5, win_length=4, games=256, sims=200, c_puct=2.0, dirichlet_alpha=0.3,
                 dirichlet_weight=0.25, net_kind=NET_STEM, net_dtype=DTYPE_F32, tower_blocks=10,
                 tower_channels=128, augment=AUGMENT_D4, max_plies=0, seed=0, rank=0, world=1, device=None,
                 leaves_per_step=1):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("azp_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device("cuda", device if isinstance(device, int) else torch.device(device).index or 0)
        self.cfg = _lib.Config(board=board, win_length=win_length, games=games, sims=sims, c_puct=c_puct,
                               dirichlet_alpha=dirichlet_alpha, dirichlet_weight=dirichlet_weight,
                               net_kind=net_kind, net_dtype=net_dtype, tower_blocks=tower_blocks,
                               tower_channels=tower_channels, augment_mode=augment, max_plies=max_plies,
                               seed=seed, rank=rank, world=world, device=self.device.index,
                               leaves_per_step=leaves_per_step)
        self.board, self.A, self.NW = board, board * board, words(board)
        self.games, self.sims = games, sims
        self.augment = augment
        handle = C.c_void_p()
        check(self.lib.azp_create(C.byref(self.cfg), C.byref(handle)))
        self.h = handle
        self._stream = None

    def _driver_stream(self):
        """The search / self-play / arena drivers replay a captured CUDA graph, which CUDA cannot capture on the
        legacy default stream: when the caller is on the default stream, run the driver on a private stream
        (ordered after the caller's pending work; the drivers are synchronous at return)."""
        cur = torch.cuda.current_stream(self.device)
        if cur.cuda_stream != 0:
            return cur
        if self._stream is None:
            self._stream = torch.cuda.Stream(self.device)
        self._stream.wait_stream(cur)
        return self._stream

    def _sp(self):
        """cudaStream_t of torch's current stream on THIS engine's device (the C side selects the device itself)."""
        return stream_ptr(self.device)

    def close(self):
        if getattr(self, "h", None):
            self.lib.azp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- weights / net -----------------------------------------------------------------------
    def weight_count(self) -> int:
        return int(self.lib.azp_weight_count(self.h))

    def load_state_dict(self, state_dict, slot=0):
        blob = flatten_state_dict(state_dict, self.cfg.net_kind, self.cfg.tower_blocks)
        self.load_blob(blob, slot)

    def load_blob(self, blob, slot=0):
        if isinstance(blob, torch.Tensor) and blob.is_cuda:
            blob = blob.contiguous().float()
            check(self.lib.azp_load_weights_device(self.h, slot, ptr(blob), blob.numel(), self._sp()), self.h)
        else:
            blob = np.ascontiguousarray(blob, dtype=np.float32)
            check(self.lib.azp_load_weights(self.h, slot, blob.ctypes.data_as(C.c_void_p), blob.size), self.h)

    def net_forward(self, planes: torch.Tensor, slot=0):
        """planes float32 [n,4,B,B] on this device -> (logits [n,A], policy [n,A], value [n])."""
        planes = planes.to(self.device, torch.float32).contiguous()
        n = planes.shape[0]
        logits = torch.empty((n, self.A), device=self.device)
        policy = torch.empty((n, self.A), device=self.device)
        value = torch.empty((n,), device=self.device)
        check(self.lib.azp_net_forward(self.h, slot, n, ptr(planes), ptr(logits), ptr(policy), ptr(value),
                                       self._sp()), self.h)
        return logits, policy, value

    def debug_activations(self, layer, n):
        c = {1: 32, 2: 64}.get(layer)
        shape = (n, self.A, c) if c else (n, 6, self.A)
        if layer == 9:                      # bring-up: conv3 clock stamps [64][16] int64
            out = torch.zeros((64, 16), dtype=torch.int64, device=self.device)
            check(self.lib.azp_debug_activations(self.h, 9, 1, ptr(out), self._sp()), self.h)
            return out
        out = torch.empty(shape, device=self.device)
        check(self.lib.azp_debug_activations(self.h, layer, n, ptr(out), self._sp()), self.h)
        return out

    # ---- step-by-step search (host evaluator hook) ------------------------------------------------
    def search_begin(self, x, o, to_move, last, ply=None, game_id=None):
        n = x.shape[0]
        self._n = n
        check(self.lib.azp_search_begin(self.h, n, ptr(x), ptr(o), ptr(to_move), ptr(last), ptr(ply),
                                        ptr(game_id), self._sp()), self.h)

    def root_expand(self, policy, add_noise=False):
        check(self.lib.azp_root_expand(self.h, ptr(policy), int(add_noise), self._sp()), self.h)

    def select(self):
        check(self.lib.azp_select(self.h, self._sp()), self.h)

    def leaf_states(self):
        n = self._n
        dev = self.device
        x = torch.empty((n, self.NW), dtype=torch.int64, device=dev)
        o = torch.empty((n, self.NW), dtype=torch.int64, device=dev)
        tm = torch.empty((n,), dtype=torch.uint8, device=dev)
        last = torch.empty((n,), dtype=torch.int16, device=dev)
        kind = torch.empty((n,), dtype=torch.uint8, device=dev)
        check(self.lib.azp_leaf_states(self.h, ptr(x), ptr(o), ptr(tm), ptr(last), ptr(kind), self._sp()), self.h)
        return x, o, tm, last, kind

    def expand_backup(self, policy, value):
        """value float32 [n], or float64 [n]: a host evaluator's Python floats are backed up unrounded (mcts.py:137-141)."""
        if value.dtype == torch.float64:
            check(self.lib.azp_expand_backup_f64(self.h, ptr(policy), ptr(value), self._sp()), self.h)
        else:
            check(self.lib.azp_expand_backup(self.h, ptr(policy), ptr(value), self._sp()), self.h)

    def root_visits(self, n=None):
        """Root child visit counts [n,A]; after selfplay()/arena_steps() pass n=self.games: the last move step's searches."""
        if n is not None:
            self._n = n
        v = torch.empty((self._n, self.A), dtype=torch.int32, device=self.device)
        check(self.lib.azp_root_visits(self.h, ptr(v), self._sp()), self.h)
        return v

    def root_priors(self):
        """Priors of the root's children [n,A] after root_expand (mcts.py:109-120); 0 on occupied cells."""
        p = torch.empty((self._n, self.A), dtype=torch.float32, device=self.device)
        check(self.lib.azp_root_priors(self.h, ptr(p), self._sp()), self.h)
        return p

    def search(self, x, o, to_move, last, ply=None, game_id=None, add_noise=False):
        """Full search of n roots with the engine's own net; returns root visit counts [n,A]."""
        n = x.shape[0]
        self._n = n
        v = torch.empty((n, self.A), dtype=torch.int32, device=self.device)
        with torch.cuda.stream(self._driver_stream()):
            check(self.lib.azp_search(self.h, 0, n, ptr(x), ptr(o), ptr(to_move), ptr(last), ptr(ply), ptr(game_id),
                                      int(add_noise), ptr(v), self._sp()), self.h)
        return v

    # ---- drivers ------------------------------------------------------------------------------------
    def set_temperature_table(self, table):
        t = np.ascontiguousarray(table, dtype=np.float64)
        check(self.lib.azp_set_temperature_table(self.h, t.ctypes.data_as(C.c_void_p), t.size), self.h)

    def set_augment(self, mode):
        """Augmentation of the next selfplay() calls (AUGMENT_NONE = raw positions, for the cross-rank exchange)."""
        check(self.lib.azp_set_augment_mode(self.h, int(mode)), self.h)
        self.augment = mode

    def set_slab_capacity(self, positions):
        """Bound the raw positions one move step emits (finished games that do not fit are emitted a step later)."""
        check(self.lib.azp_set_slab_capacity(self.h, int(positions)), self.h)
        mp = self.cfg.max_plies if 0 < self.cfg.max_plies < self.A else self.A
        self._slab = max(min(int(positions), min(self.games * mp, 16384)), mp)

    def example_capacity_floor(self) -> int:
        """Smallest output capacity (examples) azp_selfplay_run accepts."""
        mp = self.cfg.max_plies if 0 < self.cfg.max_plies < self.A else self.A
        raw = getattr(self, "_slab", None) or max(min(self.games * mp, 16384), mp)
        return raw * AUGMENT_MULT[self.augment]

    def selfplay(self, n_games, first_game=0, capacity=None, max_move_steps=0, out=None, resume=False):
        """Run self-play; returns (status, planes[cap,4,B,B], pi[cap,B,B], z[cap], written)."""
        if out is None:
            cap = max(capacity or 0, self.example_capacity_floor())
            out = (torch.empty((cap, 4, self.board, self.board), device=self.device),
                   torch.empty((cap, self.board, self.board), device=self.device),
                   torch.empty((cap,), dtype=torch.int32, device=self.device))
        planes, pi, z = out
        written = C.c_int64(0)
        with torch.cuda.stream(self._driver_stream()):
            st = check(self.lib.azp_selfplay_run(self.h, first_game, n_games, int(resume), max_move_steps, ptr(planes),
                                                 ptr(pi), ptr(z), planes.shape[0], C.byref(written), self._sp()),
                       self.h)
        return st, planes, pi, z, int(written.value)

    def selfplay_all(self, n_games, first_game=0, capacity=None):
        """Run to completion, draining the output slab whenever it fills (AZP_MORE).  Returns
        (planes, pi, z) with every example generated, in generation order."""
        chunks, out, resume = [], None, False
        while True:
            st, planes, pi, z, w = self.selfplay(n_games, first_game, capacity, out=out, resume=resume)
            out = (planes, pi, z)
            chunks.append((planes[:w].clone(), pi[:w].clone(), z[:w].clone()))
            if st != MORE:
                break
            resume = True
        return tuple(torch.cat([c[i] for c in chunks]) for i in range(3))

    def arena(self, n_games, first_game=0):
        wld = (C.c_int64 * 3)()
        with torch.cuda.stream(self._driver_stream()):
            check(self.lib.azp_arena_run(self.h, first_game, n_games, wld, self._sp()), self.h)
        return int(wld[0]), int(wld[1]), int(wld[2])

    def arena_steps(self, n_games, first_game=0, max_move_steps=1, resume=False):
        """The match in slices of lockstep moves: (status, (wins, losses, draws) so far); status MORE while games remain."""
        wld = (C.c_int64 * 3)()
        with torch.cuda.stream(self._driver_stream()):
            st = check(self.lib.azp_arena_run_steps(self.h, first_game, n_games, int(resume), max_move_steps, wld,
                                                    self._sp()), self.h)
        return st, (int(wld[0]), int(wld[1]), int(wld[2]))

    def episode_snapshot(self, history=True):
        """Per-slot episode state: dict(ply [G], game_id [G], status [G], moves [G,max_plies], pi [G,max_plies,A])."""
        g, dev = self.games, self.device
        mp = self.cfg.max_plies if 0 < self.cfg.max_plies < self.A else self.A
        out = dict(ply=torch.empty((g,), dtype=torch.int32, device=dev),
                   game_id=torch.empty((g,), dtype=torch.int64, device=dev),
                   status=torch.empty((g,), dtype=torch.uint8, device=dev))
        if history:
            out["moves"] = torch.empty((g, mp), dtype=torch.int16, device=dev)       # uint16 payload
            out["pi"] = torch.empty((g, mp, self.A), dtype=torch.float32, device=dev)
        check(self.lib.azp_episode_snapshot(self.h, ptr(out["ply"]), ptr(out["game_id"]), ptr(out["status"]),
                                            ptr(out.get("moves")), ptr(out.get("pi")), self._sp()), self.h)
        return out

    def counters(self) -> dict:
        c = _lib.Counters()
        check(self.lib.azp_get_counters(self.h, C.byref(c)), self.h)
        return c.as_dict()

    def set_profiling(self, enabled=True):
        check(self.lib.azp_set_profiling(self.h, int(enabled)), self.h)

    def get_profile(self):
        """(ms, launches, algorithmic flops) of the dominant kernel since the last call."""
        ms, n, fl = C.c_double(0), C.c_uint64(0), C.c_double(0)
        check(self.lib.azp_get_profile(self.h, C.byref(ms), C.byref(n), C.byref(fl)), self.h)
        return float(ms.value), int(n.value), float(fl.value)

    def get_profile_kernels(self) -> dict:
        """{kernel class: (ms, launches, units)} of every bracketed kernel class since the last call (azp_b200.h AZP_PROF_*)."""
        out = {}
        for kind, name in enumerate(_lib.PROF_KINDS):
            ms, n, u = C.c_double(0), C.c_uint64(0), C.c_double(0)
            check(self.lib.azp_get_profile_kernel(self.h, kind, C.byref(ms), C.byref(n), C.byref(u)), self.h)
            out[name] = (float(ms.value), int(n.value), float(u.value))
        return out

    def reset_counters(self):
        check(self.lib.azp_reset_counters(self.h), self.h)


# ---- stateless primitives ---------------------------------------------------------------------------
def _dev(t):
    if not t.is_cuda:
        raise RuntimeError("azp_b200 primitives take CUDA tensors (no CPU fallback)")
    return t


def pack_boards(cells: torch.Tensor, board: int):
    """int8 [n,B*B] (0 empty/1 X/2 O) -> (x, o) bitboards int64 [n,NW]."""
    lib = _lib.load()
    cells = _dev(cells).to(torch.int8).contiguous().view(-1, board * board)
    n = cells.shape[0]
    x = torch.zeros((n, words(board)), dtype=torch.int64, device=cells.device)
    o = torch.zeros_like(x)
    check(lib.azp_pack_boards(board, n, ptr(cells), ptr(x), ptr(o), stream_ptr()))
    return x, o


def unpack_boards(x, o, board: int):
    lib = _lib.load()
    n = x.shape[0]
    cells = torch.empty((n, board * board), dtype=torch.int8, device=x.device)
    check(lib.azp_unpack_boards(board, n, ptr(_dev(x)), ptr(o), ptr(cells), stream_ptr()))
    return cells


def legal_mask(x, o, board: int):
    lib = _lib.load()
    n = x.shape[0]
    m = torch.empty((n, board * board), dtype=torch.uint8, device=x.device)
    check(lib.azp_legal_mask(board, n, ptr(_dev(x)), ptr(o), ptr(m), stream_ptr()))
    return m


def apply_moves(x, o, to_move, last, moves, board: int):
    """In place.  Returns the per-board status tensor (0 ok / 3 invalid move)."""
    lib = _lib.load()
    n = x.shape[0]
    status = torch.empty((n,), dtype=torch.int32, device=x.device)
    moves = moves.to(torch.int16).contiguous()
    check(lib.azp_apply_moves(board, n, ptr(_dev(x)), ptr(o), ptr(to_move), ptr(last), ptr(moves), ptr(status),
                              stream_ptr()))
    return status


def terminal_status(x, o, board: int, win_length: int):
    lib = _lib.load()
    n = x.shape[0]
    r = torch.empty((n,), dtype=torch.uint8, device=x.device)
    check(lib.azp_terminal_status(board, win_length, n, ptr(_dev(x)), ptr(o), ptr(r), stream_ptr()))
    return r


def encode(x, o, to_move, last, board: int):
    lib = _lib.load()
    n = x.shape[0]
    planes = torch.empty((n, 4, board, board), dtype=torch.float32, device=x.device)
    check(lib.azp_encode(board, n, ptr(_dev(x)), ptr(o), ptr(to_move), ptr(last), ptr(planes), stream_ptr()))
    return planes


def extract_policy(visits, legal, temperature, board: int):
    lib = _lib.load()
    n = visits.shape[0]
    visits = _dev(visits).to(torch.int32).contiguous()
    legal = legal.to(torch.uint8).contiguous()
    temperature = temperature.to(visits.device, torch.float64).contiguous()
    pi = torch.empty((n, board * board), dtype=torch.float32, device=visits.device)
    check(lib.azp_extract_policy(board, n, ptr(visits), ptr(legal), ptr(temperature), ptr(pi), stream_ptr()))
    return pi


def augment_write(planes, pi, z, board: int, mode=AUGMENT_D4):
    lib = _lib.load()
    n = planes.shape[0]
    mult = AUGMENT_MULT[mode]
    planes = _dev(planes).float().contiguous()
    pi = pi.float().contiguous()
    z = z.to(torch.int32).contiguous()
    op = torch.empty((n * mult, 4, board, board), dtype=torch.float32, device=planes.device)
    opi = torch.empty((n * mult, board, board), dtype=torch.float32, device=planes.device)
    oz = torch.empty((n * mult,), dtype=torch.int32, device=planes.device)
    check(lib.azp_augment_write(board, mode, n, ptr(planes), ptr(pi), ptr(z), ptr(op), ptr(opi), ptr(oz),
                                stream_ptr()))
    return op, opi, oz
