"""``Gomoku`` facade (alphazero/games.py:20-243) over the CUDA bitboard primitives.

A ``Gomoku`` object is a batch-of-one view: its rules questions (legal moves, apply,
terminal / winner, encode) are answered by the same kernels that run thousands of games
(``GomokuBatch``).  Host-side it keeps the reference's public fields: ``board`` (list of
lists of "X"/"O"/None), ``current_player``, ``winner``, ``last_action``, ``board_size``,
``win_length``."""

from __future__ import annotations

import torch

from . import engine as _e
from .constants import BOARD_SIZE, DRAW, O, WIN_LENGTH, X

_SYM = {0: None, 1: X, 2: O, 3: DRAW}
_CODE = {None: 0, X: 1, O: 2, DRAW: 3}


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("azp_b200 needs a CUDA device; there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


class GomokuBatch:
    """n positions resident on the GPU as bitboards."""

    def __init__(self, n, board_size=BOARD_SIZE, win_length=WIN_LENGTH, device=None):
        self.n, self.board_size, self.win_length = n, board_size, win_length
        dev = device or _device()
        nw = _e.words(board_size)
        self.x = torch.zeros((n, nw), dtype=torch.int64, device=dev)
        self.o = torch.zeros((n, nw), dtype=torch.int64, device=dev)
        self.to_move = torch.ones((n,), dtype=torch.uint8, device=dev)
        self.last = torch.full((n,), -1, dtype=torch.int16, device=dev)

    @classmethod
    def from_cells(cls, cells, to_move=None, last=None, win_length=WIN_LENGTH, device=None):
        """cells: int8 [n,B,B] (0 empty / 1 X / 2 O)."""
        dev = device or _device()
        cells = torch.as_tensor(cells, dtype=torch.int8, device=dev)
        n, b = cells.shape[0], cells.shape[-1]
        self = cls(n, b, win_length, dev)
        self.x, self.o = _e.pack_boards(cells.reshape(n, -1), b)
        if to_move is not None:
            self.to_move = torch.as_tensor(to_move, dtype=torch.uint8, device=dev).reshape(n).contiguous()
        if last is not None:
            self.last = torch.as_tensor(last, dtype=torch.int16, device=dev).reshape(n).contiguous()
        return self

    def clone(self):
        c = GomokuBatch.__new__(GomokuBatch)
        c.n, c.board_size, c.win_length = self.n, self.board_size, self.win_length
        c.x, c.o, c.to_move, c.last = self.x.clone(), self.o.clone(), self.to_move.clone(), self.last.clone()
        return c

    def cells(self):
        return _e.unpack_boards(self.x, self.o, self.board_size).view(self.n, self.board_size, self.board_size)

    def legal_mask(self):
        return _e.legal_mask(self.x, self.o, self.board_size).view(self.n, self.board_size, self.board_size)

    def apply(self, moves):
        """moves: flat cell index per board; raises ValueError("Invalid move") like games.py:76-77."""
        moves = torch.as_tensor(moves, dtype=torch.int16, device=self.x.device).reshape(self.n)
        status = _e.apply_moves(self.x, self.o, self.to_move, self.last, moves, self.board_size)
        if bool((status != 0).any()):
            raise ValueError("Invalid move")
        return self

    def results(self):
        """uint8 [n]: 0 not terminal / 1 X / 2 O / 3 draw."""
        return _e.terminal_status(self.x, self.o, self.board_size, self.win_length)

    def encode(self):
        return _e.encode(self.x, self.o, self.to_move, self.last, self.board_size)


class Gomoku:
    def __init__(self, board_size=BOARD_SIZE, win_length=WIN_LENGTH):
        if not isinstance(board_size, int):
            print(f"[Gomoku] Invalid board size: {board_size}. Must be an integer.")
            raise ValueError("Board size must be an integer.")
        self.board_size = board_size
        self.win_length = win_length
        self.board = [[None for _ in range(board_size)] for _ in range(board_size)]
        self.current_player = X
        self.winner = None
        self.last_action = None

    # ---- device view ----------------------------------------------------------------------------
    def _batch(self) -> GomokuBatch:
        cells = [[_CODE[c] for c in row] for row in self.board]
        last = -1 if self.last_action is None else self.last_action[0] * self.board_size + self.last_action[1]
        return GomokuBatch.from_cells([cells], to_move=[_CODE[self.current_player]], last=[last],
                                      win_length=self.win_length)

    # ---- reference API ----------------------------------------------------------------------------
    def get_legal_actions(self):
        mask = self._batch().legal_mask()[0].cpu()
        idx = torch.nonzero(mask, as_tuple=False)          # row-major order (games.py:42-47)
        return [(int(r), int(c)) for r, c in idx.tolist()]

    def get_other_player(self, player):
        if player not in (X, O):
            raise ValueError(f"Invalid player: {player}. Must be 'X' or 'O'.")
        return O if player == X else X

    def apply_action(self, action):
        r, c = action
        if self.current_player not in (X, O):
            raise ValueError(f"Invalid player: {self.current_player}. Must be 'X' or 'O'.")
        b = self._batch()
        b.apply([int(r) * self.board_size + int(c)])      # ValueError("Invalid move") if occupied
        new = self.clone()
        new.board[r][c] = self.current_player
        new.current_player = self.get_other_player(self.current_player)
        new.last_action = action
        return new

    def encode(self, device=None):
        t = self._batch().encode()[0]
        return t if device is None else t.to(device)

    def is_terminal(self):
        if self.winner is not None:
            return True
        res = int(self._batch().results()[0])
        if res:
            self.winner = _SYM[res]
            return True
        return False

    def get_game_result(self):
        if not self.is_terminal():
            return None
        return self.winner if self.winner is not None else DRAW

    def rot90(self):
        """Clockwise board rotation (games.py:183-189); last_action is left as is, like the reference."""
        new = self.clone()
        new.board = [list(row) for row in zip(*new.board[::-1])]
        return new

    def flip(self):
        new = self.clone()
        new.board = [row[::-1] for row in new.board]
        return new

    def clone(self):
        new = Gomoku(self.board_size, self.win_length)
        new.board = [row[:] for row in self.board]
        new.current_player = self.current_player
        new.winner = self.winner
        if self.last_action is not None:
            new.last_action = tuple(self.last_action)
        return new

    def __repr__(self):
        rows = "\n".join(" | ".join(c if c is not None else " " for c in row) for row in self.board)
        return f"Gomoku(\n{rows}\n)"
