"""Python handle on the CUDA training step (C ABI: azp_trainer_* / azp_train_loss_grad /
azp_adamw_step; reference: NeuralNetworkController.train_step, alphazero/controller.py:82-131).

Parameters, gradients and AdamW moments are flat float32 CUDA tensors in the ``azp_load_weights``
layout (reference ``state_dict`` order); the library owns only workspaces.  No CPU fallback."""

from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import EngineError, ptr, stream_ptr

BETAS = (0.9, 0.999)      # torch.optim.AdamW defaults; the reference passes only weight_decay (controller.py:75-79)
EPS = 1e-8
WEIGHT_DECAY = 1e-4


def _check(status, handle=None):
    if status != _lib.OK:
        msg = _lib.load().azp_trainer_last_error(handle)
        raise EngineError(status, msg.decode() if msg else "")


class Trainer:
    """Workspaces for batches of up to ``max_batch`` positions of a ``board`` x ``board`` stem net."""

    def __init__(self, board: int, max_batch: int, device=None):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise RuntimeError("azp_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device("cuda", device if isinstance(device, int) else torch.device(device).index or 0)
        self.board, self.A, self.max_batch = board, board * board, max_batch
        handle = C.c_void_p()
        _check(self.lib.azp_trainer_create(board, max_batch, self.device.index, C.byref(handle)))
        self.h = handle

    def close(self):
        if getattr(self, "h", None):
            self.lib.azp_trainer_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def weight_count(self) -> int:
        return int(self.lib.azp_trainer_weight_count(self.h))

    def _f32(self, t, shape):
        t = torch.as_tensor(t).to(self.device, torch.float32).contiguous()
        return t.view(shape)

    def loss_grad(self, params: torch.Tensor, planes, pi, z, grads: torch.Tensor | None = None, want_grads=True):
        """Forward + loss (+ backward).  Returns (losses float32[4] CUDA tensor = policy, value, entropy,
        total; grads flat CUDA tensor or None)."""
        n = int(planes.shape[0])
        planes = self._f32(planes, (n, 4, self.board, self.board))
        pi = self._f32(pi, (n, self.A))
        z = self._f32(z, (n,))
        if not (params.is_cuda and params.dtype == torch.float32 and params.is_contiguous()
                and params.numel() == self.weight_count()):
            raise ValueError("params must be a contiguous float32 CUDA blob of weight_count() elements")
        if want_grads and grads is None:
            grads = torch.empty_like(params)
        losses = torch.empty(4, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _check(self.lib.azp_train_loss_grad(self.h, n, ptr(params), ptr(planes), ptr(pi), ptr(z),
                                                ptr(grads) if want_grads else None, ptr(losses), stream_ptr()), self.h)
        return losses, (grads if want_grads else None)


def adamw_step(params, grads, exp_avg, exp_avg_sq, step: int, lr: float, betas=BETAS, eps=EPS,
               weight_decay=WEIGHT_DECAY):
    """One torch.optim.AdamW update on flat float32 CUDA blobs, in place."""
    for t in (params, grads, exp_avg, exp_avg_sq):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.numel() == params.numel()):
            raise ValueError("adamw_step takes contiguous float32 CUDA blobs of equal size (no CPU fallback)")
    with torch.cuda.device(params.device):
        _check(_lib.load().azp_adamw_step(params.numel(), ptr(params), ptr(grads), ptr(exp_avg), ptr(exp_avg_sq),
                                          int(step), float(lr), float(betas[0]), float(betas[1]), float(eps),
                                          float(weight_decay), stream_ptr()))
