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
        self._head_dev = None   # the same cursor as an int64 on the device (ring writer kernel)
        self._head_dev_owns = False
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
    def _alloc_ring(self, planes, pi, z):
        if self._planes is None:
            self._planes = torch.empty((self.capacity,) + tuple(planes.shape[1:]), dtype=planes.dtype, device=planes.device)
            self._pi = torch.empty((self.capacity,) + tuple(pi.shape[1:]), dtype=pi.dtype, device=pi.device)
            self._z = torch.empty((self.capacity,), dtype=z.dtype, device=z.device)

    def extend_device(self, planes, pi, z):
        """Append a slab of examples ([n,4,B,B], [n,B,B], [n]) in order, evicting the oldest when full --
        deque(maxlen) semantics on the ring.  CUDA slabs go through the library's ring writer
        (azp_augment_write_ring, one kernel, no host synchronisation); host tensors are indexed in place."""
        n = planes.shape[0]
        self._alloc_ring(planes, pi, z)
        if n > self.capacity:                      # only the newest `capacity` survive
            planes, pi, z = planes[-self.capacity:], pi[-self.capacity:], z[-self.capacity:]
            self._head += n - self.capacity
            n = self.capacity
        if planes.is_cuda and planes.dtype == torch.float32 and z.dtype == torch.int32 and planes.shape[-1] == planes.shape[-2]:
            self.write_slab(planes, pi, z, n_max=n)
        else:
            idx = (torch.arange(n, device=planes.device) + self._head) % self.capacity
            self._planes[idx], self._pi[idx], self._z[idx] = planes, pi, z
        self._head += n

    def write_slab(self, planes, pi, z, n_max, n_dev=None, augment=None):
        """Ring write of up to ``n_max`` CUDA positions through azp_augment_write_ring.  ``n_dev`` (int32 CUDA tensor of
        one element) gives the count on the device -- the header of a slab received by an all-gather -- and ``augment``
        the symmetry expansion applied on the way in; with a device-side count the host-side cursor is NOT advanced
        (call ``sync_head`` before reading the ring)."""
        from . import _lib
        board = planes.shape[-1]
        mode = _lib.AUGMENT_NONE if augment is None else augment
        if self._planes is None:
            self._alloc_ring(planes, pi, z)
        if self._head_dev is None:
            self._head_dev = torch.tensor([self._head], dtype=torch.int64, device=planes.device)
        elif n_dev is None and not self._head_dev_owns:
            self._head_dev.fill_(self._head)       # host cursor is authoritative until a device-side count is used
        if n_dev is not None:
            self._head_dev_owns = True
        lib = _lib.load()
        _lib.check(lib.azp_augment_write_ring(board, mode, int(n_max), _lib.ptr(n_dev), _lib.ptr(planes.contiguous()),
                                              _lib.ptr(pi.contiguous()), _lib.ptr(z.contiguous()), _lib.ptr(self._planes),
                                              _lib.ptr(self._pi), _lib.ptr(self._z), _lib.ptr(self._head_dev),
                                              self.capacity, _lib.stream_ptr(planes.device)))

    def sync_head(self):
        """Adopt the device-side cursor (one 8-byte read-back) after slabs with device-side counts were written."""
        if self._head_dev is not None and self._head_dev_owns:
            self._head = int(self._head_dev.item())
        return self._head

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
