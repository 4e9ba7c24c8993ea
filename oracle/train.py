"""Oracle: one training step of the stem net (test infrastructure only).

Restates ``NeuralNetworkController.train_step`` (alphazero/controller.py:82-131) and the
optimizer it constructs (``torch.optim.AdamW(net.parameters(), weight_decay=1e-4)``,
controller.py:75-79; ``lr`` overwritten before every step, controller.py:115) as pure functions
of a ``state_dict``.  As in the reference, the arithmetic of the forward/backward pass is
PyTorch's CPU implementation (autograd over ``conv2d`` / ``linear`` / ``cross_entropy`` /
``mse_loss``; third-party and unpinned there, torch 2.11.0 in this image); the AdamW update is
written out operation by operation (torch/optim/adamw.py single-tensor path) so the CUDA kernel
can be compared against it element-wise.

Pinned by ``tests/golden/train.npz``: losses, first-step gradients and the parameters after
three steps of the UNMODIFIED reference controller on seeded inputs
(``tests/golden/make_golden.py train``).
"""

from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from .net import STEM_KEYS

BETAS = (0.9, 0.999)       # torch.optim.AdamW defaults (the reference passes only weight_decay)
EPS = 1e-8
WEIGHT_DECAY = 1e-4        # controller.py:75

PARAM_KEYS = tuple(f"{k}.{f}" for k in STEM_KEYS for f in ("weight", "bias"))


def _t(x, dtype):
    x = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
    return x.to(dtype)


def forward_train(p: dict, x):
    """net.py:55-72 with autograd enabled.  Returns (logits (n, A), value (n, 1))."""
    h = F.relu(F.conv2d(x, p["conv1.weight"], p["conv1.bias"], padding=1))
    h = F.relu(F.conv2d(h, p["conv2.weight"], p["conv2.bias"], padding=1))
    h = F.relu(F.conv2d(h, p["conv3.weight"], p["conv3.bias"], padding=1))
    n = h.shape[0]
    pol = F.relu(F.conv2d(h, p["policy_conv.weight"], p["policy_conv.bias"]))
    logits = F.linear(pol.reshape(n, -1), p["policy_fc.weight"], p["policy_fc.bias"])
    val = F.relu(F.conv2d(h, p["value_conv.weight"], p["value_conv.bias"]))
    val = F.relu(F.linear(val.reshape(n, -1), p["value_fc1.weight"], p["value_fc1.bias"]))
    val = torch.tanh(F.linear(val, p["value_fc2.weight"], p["value_fc2.bias"]))
    return logits, val


def loss_and_grads(params: dict, planes, pi, z, dtype=torch.float32):
    """controller.py:117-124.  planes (n,4,B,B), pi (n,B,B) or (n,A), z (n,).
    Returns ({"policy", "value", "entropy", "total"} floats, grads dict of numpy arrays).
    ``entropy`` is the mean entropy of the predicted policy (the quantity the reference's
    logging line means to report; as written, controller.py:130 takes log of raw logits)."""
    p = {k: _t(params[k], dtype).clone().requires_grad_(True) for k in PARAM_KEYS}
    x = _t(planes, dtype)
    pi = _t(pi, dtype).flatten(1, -1)
    z = _t(z, dtype)
    logits, value = forward_train(p, x)
    value_loss = F.mse_loss(value.view(-1), z)
    policy_loss = F.cross_entropy(logits, pi)
    loss = policy_loss + value_loss
    loss.backward()
    with torch.no_grad():
        logp = torch.log_softmax(logits, dim=1)
        entropy = -(logp.exp() * logp).sum(dim=1).mean()
    losses = {"policy": policy_loss.item(), "value": value_loss.item(), "entropy": entropy.item(),
              "total": loss.item()}
    return losses, {k: p[k].grad.detach().numpy() for k in PARAM_KEYS}


def adamw_update(param, grad, exp_avg, exp_avg_sq, step: int, lr: float, betas=BETAS, eps=EPS,
                 weight_decay=WEIGHT_DECAY):
    """torch/optim/adamw.py, single-tensor path, float32 arrays updated IN PLACE:
    decoupled decay, lerp of the first moment, second moment, bias corrections as Python doubles."""
    b1, b2 = betas
    f = np.float32
    param *= f(1.0 - lr * weight_decay)
    exp_avg += (grad - exp_avg) * f(1.0 - b1)
    exp_avg_sq *= f(b2)
    exp_avg_sq += f(1.0 - b2) * grad * grad
    bc1 = 1.0 - b1 ** step
    bc2 = 1.0 - b2 ** step
    step_size = lr / bc1
    denom = np.sqrt(exp_avg_sq) / f(bc2 ** 0.5) + f(eps)
    param += f(-step_size) * (exp_avg / denom)


class TrainerState:
    """Parameters + AdamW moments + step counter (what the reference controller holds in
    ``self.net`` and ``self.optimizer``)."""

    def __init__(self, params: dict):
        self.params = {k: np.array(params[k], dtype=np.float32) for k in PARAM_KEYS}
        self.exp_avg = {k: np.zeros_like(v) for k, v in self.params.items()}
        self.exp_avg_sq = {k: np.zeros_like(v) for k, v in self.params.items()}
        self.step = 0

    def train_step(self, planes, pi, z, lr: float):
        """controller.py:82-131.  Returns the losses dict of ``loss_and_grads``."""
        losses, grads = loss_and_grads(self.params, planes, pi, z)
        self.step += 1
        for k in PARAM_KEYS:
            adamw_update(self.params[k], grads[k], self.exp_avg[k], self.exp_avg_sq[k], self.step, lr)
        self.last_grads = grads
        return losses


def synthetic_batch(board: int, n: int, seed: int):
    """Seeded training batch shaped like the self-play output: 0/1 planes with disjoint stones
    and a one-hot last-move plane (games.py:86-129), visit-count-like pi summing to 1 with
    zeros on occupied cells, z in {-1, 0, 1}."""
    rs = np.random.RandomState(seed)
    a = board * board
    planes = np.zeros((n, 4, board, board), dtype=np.float32)
    pi = np.zeros((n, a), dtype=np.float32)
    z = rs.randint(-1, 2, size=n).astype(np.float32)
    for i in range(n):
        stones = rs.randint(0, max(2, a // 2))
        cells = rs.permutation(a)[:stones]
        own, opp = cells[: stones // 2], cells[stones // 2:]
        planes[i, 0].reshape(-1)[own] = 1.0
        planes[i, 1].reshape(-1)[opp] = 1.0
        if stones:
            planes[i, 2].reshape(-1)[cells[-1]] = 1.0
        w = rs.gamma(0.5, size=a).astype(np.float64)
        w[cells] = 0.0
        pi[i] = (w / w.sum()).astype(np.float32)
    return planes, pi.reshape(n, board, board), z
