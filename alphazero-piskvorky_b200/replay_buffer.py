"""``ReplayBuffer`` (alphazero/replay_buffer.py:14-75): bounded FIFO of ``(state, pi, z)``
examples, plus a device-resident ring (``extend_device``) with the same eviction order so
generated slabs never have to leave HBM."""

from __future__ import annotations

import pickle
import random
from collections import deque

import torch


class ReplayBuffer:
    def __init__(self, capacity, device=None, board_size=None):
        self.buffer = deque(maxlen=capacity)
        self.capacity = capacity
        # device ring (allocated on first extend_device)
        self._planes = self._pi = self._z = None
        self._head = 0          # total examples ever written to the ring
        self._device = device

    # ---- reference API ------------------------------------------------------------------------------
    def extend(self, game_examples):
        self.buffer.extend(game_examples)

    def sample_batch(self, batch_size):
        if len(self.buffer) < batch_size:
            return self.buffer
        return random.sample(self.buffer, batch_size)

    def all(self):
        return list(self.buffer)

    def save(self, filename):
        with open(filename, "wb") as f:
            pickle.dump(self.buffer, f)

    def load(self, filename):
        with open(filename, "rb") as f:
            print(f"[ReplayBuffer:] Loading replay buffer from {filename}...")
            self.buffer = deque(pickle.load(f), maxlen=self.buffer.maxlen)

    def __len__(self):
        return len(self.buffer)

    # ---- device ring -----------------------------------------------------------------------------------
    def extend_device(self, planes, pi, z):
        """Append a slab of examples (CUDA tensors [n,4,B,B], [n,B,B], [n]) in order, evicting the
        oldest when full -- deque(maxlen) semantics on the ring."""
        n = planes.shape[0]
        if self._planes is None:
            self._planes = torch.empty((self.capacity,) + tuple(planes.shape[1:]), dtype=planes.dtype, device=planes.device)
            self._pi = torch.empty((self.capacity,) + tuple(pi.shape[1:]), dtype=pi.dtype, device=pi.device)
            self._z = torch.empty((self.capacity,), dtype=z.dtype, device=z.device)
        if n > self.capacity:                      # only the newest `capacity` survive
            planes, pi, z = planes[-self.capacity:], pi[-self.capacity:], z[-self.capacity:]
            self._head += n - self.capacity
            n = self.capacity
        idx = (torch.arange(n, device=planes.device) + self._head) % self.capacity
        self._planes[idx], self._pi[idx], self._z[idx] = planes, pi, z
        self._head += n

    def device_len(self):
        return min(self._head, self.capacity)

    def device_all(self):
        """Ring contents oldest-first (what ``list(deque)`` would give)."""
        n = self.device_len()
        if n == 0:
            return None
        start = self._head - n
        idx = (torch.arange(n, device=self._planes.device) + start) % self.capacity
        return self._planes[idx], self._pi[idx], self._z[idx]

    def sample_device(self, batch_size, generator=None):
        n = self.device_len()
        planes, pi, z = self.device_all()
        if n < batch_size:
            return planes, pi, z
        sel = torch.randperm(n, device=planes.device, generator=generator)[:batch_size]
        return planes[sel], pi[sel], z[sel]
