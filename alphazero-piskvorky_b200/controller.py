"""``make_policy_value_fn`` / ``NeuralNetworkController`` (alphazero/controller.py:33-79).

The controller only carries the network and its device for the self-play path; training
(controller.py:82-187) is outside this engine's scope."""

from __future__ import annotations

import torch

from . import engine as _e
from ._lib import DTYPE_F32, NET_STEM, NET_TOWER
from .net import GomokuNet, GomokuTowerNet


class NeuralNetworkController:
    def __init__(self, net, device="cuda", batch_size=1024):
        self.device = device
        self.net = net
        self.batch_size = batch_size
        self.training_history = []

    def save(self, path):
        torch.save(self.net.state_dict(), path)

    def load(self, path):
        self.net.load_state_dict(torch.load(path, map_location="cpu", weights_only=True))

    def eval(self):
        self.net.eval()

    def train_mode(self):
        self.net.train()


def net_kind_of(net):
    return NET_TOWER if isinstance(net, GomokuTowerNet) else NET_STEM


def make_policy_value_fn(controller, dtype=DTYPE_F32):
    """``policy_value_fn(state) -> (policy (B,B) float32 summing to 1 over ALL cells, value float)``
    -- batch of one through the engine's CUDA net (controller.py:39-53)."""
    net = controller.net
    cache = {}

    def policy_value_fn(state):
        b = state.board_size
        eng = cache.get(b)
        if eng is None:
            kw = {}
            if isinstance(net, GomokuTowerNet):
                kw = dict(net_kind=NET_TOWER, tower_blocks=net.blocks, tower_channels=net.channels)
            eng = cache[b] = _e.Engine(board=b, win_length=state.win_length, games=1, sims=1, net_dtype=dtype, **kw)
            eng.load_state_dict(net.state_dict())
        planes = state.encode().unsqueeze(0)
        _, policy, value = eng.net_forward(planes)
        return policy[0].cpu().numpy().reshape(b, b), float(value[0].item())

    policy_value_fn.controller = controller      # lets MCTS run the whole search on the device
    return policy_value_fn
