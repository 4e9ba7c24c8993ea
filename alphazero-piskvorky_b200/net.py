"""``GomokuNet`` (alphazero/net.py:16-72): same parameter names and shapes, so reference
checkpoints load unchanged.  ``forward`` is the plain differentiable PyTorch graph that the
(out-of-scope) trainer needs; self-play never calls it -- inference on the hot path runs in
the engine's CUDA kernels (``Engine.net_forward`` / the search drivers)."""

from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as functional

from .constants import BOARD_SIZE


class GomokuNet(nn.Module):
    def __init__(self, board_size=BOARD_SIZE, device=None):
        super().__init__()
        self.device = device if device else torch.device("cpu")
        self.board_size = board_size
        self.conv1 = nn.Conv2d(4, 32, kernel_size=3, padding=1)
        self.conv2 = nn.Conv2d(32, 64, kernel_size=3, padding=1)
        self.conv3 = nn.Conv2d(64, 128, kernel_size=3, padding=1)
        self.policy_conv = nn.Conv2d(128, 4, kernel_size=1)
        self.policy_fc = nn.Linear(4 * board_size * board_size, board_size * board_size)
        self.value_conv = nn.Conv2d(128, 2, kernel_size=1)
        self.value_fc1 = nn.Linear(2 * board_size * board_size, 64)
        self.value_fc2 = nn.Linear(64, 1)

    def forward(self, x):
        out = functional.relu(self.conv1(x))
        out = functional.relu(self.conv2(out))
        out = functional.relu(self.conv3(out))
        a = self.board_size * self.board_size
        policy = functional.relu(self.policy_conv(out)).view(-1, 4 * a)
        policy_logits = self.policy_fc(policy)
        value = functional.relu(self.value_conv(out))
        value = functional.relu(self.value_fc1(value.view(value.size(0), -1)))
        value = torch.tanh(self.value_fc2(value))
        return policy_logits, value


class _ResBlock(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv1 = nn.Conv2d(c, c, 3, padding=1, bias=False)
        self.bn1 = nn.BatchNorm2d(c)
        self.conv2 = nn.Conv2d(c, c, 3, padding=1, bias=False)
        self.bn2 = nn.BatchNorm2d(c)

    def forward(self, x):
        y = functional.relu(self.bn1(self.conv1(x)))
        return functional.relu(self.bn2(self.conv2(y)) + x)


class GomokuTowerNet(nn.Module):
    """Residual tower of BASELINE config 4, key layout of the author's old checkpoints
    (alphazero/models/old/model_20250729_052241.pt)."""

    def __init__(self, board_size=15, channels=128, blocks=10):
        super().__init__()
        self.board_size, self.channels, self.blocks = board_size, channels, blocks
        a = board_size * board_size
        self.conv = nn.Conv2d(4, channels, 3, padding=1)
        self.bn = nn.BatchNorm2d(channels)
        for i in range(1, blocks + 1):
            setattr(self, f"res{i}", _ResBlock(channels))
        self.policy_conv = nn.Conv2d(channels, 2, 1)
        self.policy_bn = nn.BatchNorm2d(2)
        self.policy_fc = nn.Linear(2 * a, a)
        self.value_conv = nn.Conv2d(channels, 1, 1)
        self.value_bn = nn.BatchNorm2d(1)
        self.value_fc1 = nn.Linear(a, 64)
        self.value_fc2 = nn.Linear(64, 1)

    def forward(self, x):
        h = functional.relu(self.bn(self.conv(x)))
        for i in range(1, self.blocks + 1):
            h = getattr(self, f"res{i}")(h)
        n = h.shape[0]
        p = functional.relu(self.policy_bn(self.policy_conv(h))).reshape(n, -1)
        v = functional.relu(self.value_bn(self.value_conv(h))).reshape(n, -1)
        v = functional.relu(self.value_fc1(v))
        return self.policy_fc(p), torch.tanh(self.value_fc2(v))
