"""Oracle: replay buffer semantics (test infrastructure only).  Restates
``alphazero/replay_buffer.py:14-48,74-75``: bounded FIFO; ``extend`` appends in
order and evicts the oldest; ``sample_batch(n)`` draws n distinct examples
uniformly, or returns everything when fewer than n are held."""

from __future__ import annotations

import random


class Fifo:
    def __init__(self, capacity: int):
        self.capacity = capacity
        self.items: list = []

    def extend(self, examples) -> None:
        self.items.extend(examples)
        overflow = len(self.items) - self.capacity
        if overflow > 0:
            del self.items[:overflow]

    def sample_batch(self, n: int, rng=random):
        if len(self.items) < n:
            return list(self.items)
        return rng.sample(self.items, n)

    def all(self) -> list:
        return list(self.items)

    def __len__(self) -> int:
        return len(self.items)
