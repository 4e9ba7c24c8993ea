"""Oracle: Gomoku rules (test infrastructure only -- see oracle/__init__.py).

Restates ``alphazero/games.py`` on a numpy int8 board.  Cell codes:
``EMPTY=0, X=1, O=2``; result codes add ``DRAW=3``; ``0`` means "not terminal".
"""

from __future__ import annotations

import numpy as np

EMPTY, X, O, DRAW = 0, 1, 2, 3

# direction order of the reference's win scan: alphazero/games.py:150-155
_DIRECTIONS = ((0, 1), (1, 0), (1, 1), (1, -1))


def other(player: int) -> int:
    """alphazero/games.py:49-61 -- the opponent; anything but X/O is an error."""
    if player not in (X, O):
        raise ValueError(f"Invalid player: {player}")
    return O if player == X else X


class State:
    """A position.  Mirrors the fields of ``Gomoku`` (alphazero/games.py:20-33):
    board, side to move (X first), win length, winner cache, last action."""

    __slots__ = ("size", "win", "board", "to_move", "winner", "last")

    def __init__(self, size: int = 5, win: int = 4):
        if not isinstance(size, int):
            raise ValueError("Board size must be an integer.")  # games.py:25-27
        self.size = size
        self.win = win
        self.board = np.zeros((size, size), dtype=np.int8)
        self.to_move = X                       # games.py:29
        self.winner = 0                        # cache, games.py:31
        self.last = None                       # games.py:32

    # ---- construction helpers (oracle-side only) -------------------------
    @classmethod
    def from_board(cls, board, to_move=X, win=4, last=None):
        board = np.asarray(board, dtype=np.int8)
        s = cls(int(board.shape[0]), win)
        s.board = board.copy()
        s.to_move = to_move
        s.last = None if last is None else (int(last[0]), int(last[1]))
        return s

    def copy(self) -> "State":
        """alphazero/games.py:199-209 -- deep copy INCLUDING the winner cache."""
        s = State.__new__(State)
        s.size, s.win = self.size, self.win
        s.board = self.board.copy()
        s.to_move = self.to_move
        s.winner = self.winner
        s.last = self.last
        return s

    # ---- rules ------------------------------------------------------------
    def legal_mask(self) -> np.ndarray:
        """Boolean (B,B) mask of empty cells.  No terminal guard, exactly like
        alphazero/games.py:35-47 (a won board still lists its empties)."""
        return self.board == EMPTY

    def legal_actions(self) -> list[tuple[int, int]]:
        """Row-major list of empties -- the order defines child order and so
        the PUCT tie-break (alphazero/games.py:42-47)."""
        rr, cc = np.nonzero(self.board == EMPTY)   # np.nonzero is row-major
        return list(zip(rr.tolist(), cc.tolist()))

    def play(self, action) -> "State":
        """alphazero/games.py:64-82 -- occupied cell => ValueError("Invalid move");
        otherwise a new state with the mover's stone, side flipped, last action
        recorded.  The winner cache is carried over by the copy."""
        r, c = int(action[0]), int(action[1])
        if self.board[r, c] != EMPTY:
            raise ValueError("Invalid move")
        s = self.copy()
        s.board[r, c] = self.to_move
        s.to_move = other(self.to_move)
        s.last = (r, c)
        return s

    def _line_from(self, r, c, dr, dc, colour) -> bool:
        """alphazero/games.py:212-227 -- ``win`` consecutive stones of ``colour``
        STARTING at (r,c) going (dr,dc); so overlines count as wins."""
        n = self.size
        for i in range(self.win):
            rr, cc = r + i * dr, c + i * dc
            if not (0 <= rr < n and 0 <= cc < n and self.board[rr, cc] == colour):
                return False
        return True

    def result(self) -> int:
        """0 if not terminal else X / O / DRAW.

        alphazero/games.py:133-166: cached winner first; else scan cells
        row-major and, for each stone, the four directions in the order
        (0,1),(1,0),(1,1),(1,-1); the first line start found decides the winner
        (matters only on unreachable two-winner boards); else a full board is a
        draw.  alphazero/games.py:169-179 maps that to the result."""
        if self.winner:
            return self.winner
        n = self.size
        for r in range(n):
            for c in range(n):
                colour = int(self.board[r, c])
                if colour == EMPTY:
                    continue
                for dr, dc in _DIRECTIONS:
                    if self._line_from(r, c, dr, dc, colour):
                        self.winner = colour
                        return colour
        if not (self.board == EMPTY).any():
            self.winner = DRAW
            return DRAW
        return 0

    def is_terminal(self) -> bool:
        return self.result() != 0

    def planes(self) -> np.ndarray:
        """alphazero/games.py:86-129 -- float32 (4,B,B): plane 0 = side-to-move's
        stones, plane 1 = opponent's stones (value +1.0, games.py:120, despite the
        docstring saying -1), plane 2 = one-hot of the last action (all zero at
        ply 0), plane 3 = zeros."""
        n = self.size
        out = np.zeros((4, n, n), dtype=np.float32)
        out[0][self.board == self.to_move] = 1.0
        out[1][self.board == other(self.to_move)] = 1.0
        if self.last is not None:
            out[2, self.last[0], self.last[1]] = 1.0
        return out


def result_vectorised(board: np.ndarray, win: int) -> int:
    """Same answer as ``State.result`` on a bare board, computed with array
    shifts -- used by the fuzz tests to cross-check the scalar scan and as the
    fast checker at full batch sizes.  Finds every line-start cell per colour
    and direction, then takes the row-major-first one (games.py:144-157)."""
    n = board.shape[0]
    best = None  # (flat index, colour)
    for colour in (X, O):
        own = board == colour
        for dr, dc in _DIRECTIONS:
            start = own.copy()
            for i in range(1, win):
                shifted = np.zeros_like(own)
                r0, r1 = max(0, -i * dr), min(n, n - i * dr)
                c0, c1 = max(0, -i * dc), min(n, n - i * dc)
                if r0 < r1 and c0 < c1:
                    shifted[r0:r1, c0:c1] = own[r0 + i * dr:r1 + i * dr, c0 + i * dc:c1 + i * dc]
                start &= shifted
            idx = np.flatnonzero(start)
            if idx.size and (best is None or idx[0] < best[0]):
                best = (int(idx[0]), colour)
    if best is not None:
        return best[1]
    return DRAW if not (board == EMPTY).any() else 0
