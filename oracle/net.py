"""Oracle: policy/value network forward (test infrastructure only).

Restates ``alphazero/net.py:28-72`` (``GomokuNet``) and the evaluator wrapper
``alphazero/controller.py:33-55`` as pure functions of a ``state_dict``.  The
arithmetic itself (conv2d / linear / softmax / tanh) is PyTorch's CPU fp32
implementation, exactly as in the reference (third-party, unpinned there;
torch 2.11.0 in this image).

Also defines the residual tower of BASELINE config 4.  The reference ships NO
forward source for it (only old checkpoints with that key layout, SURVEY.md
section 2), so the tower oracle is "parity unpinned" beyond key names/shapes.
"""

from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

STEM_KEYS = ("conv1", "conv2", "conv3", "policy_conv", "policy_fc",
             "value_conv", "value_fc1", "value_fc2")


def stem_shapes(board: int) -> dict:
    """Parameter shapes of GomokuNet(board) -- alphazero/net.py:37-53."""
    a = board * board
    return {
        "conv1.weight": (32, 4, 3, 3), "conv1.bias": (32,),
        "conv2.weight": (64, 32, 3, 3), "conv2.bias": (64,),
        "conv3.weight": (128, 64, 3, 3), "conv3.bias": (128,),
        "policy_conv.weight": (4, 128, 1, 1), "policy_conv.bias": (4,),
        "policy_fc.weight": (a, 4 * a), "policy_fc.bias": (a,),
        "value_conv.weight": (2, 128, 1, 1), "value_conv.bias": (2,),
        "value_fc1.weight": (64, 2 * a), "value_fc1.bias": (64,),
        "value_fc2.weight": (1, 64), "value_fc2.bias": (1,),
    }


def synthetic_stem_params(board: int, seed: int, scale: float = 1.0) -> dict:
    """Deterministic stand-in for "random init" that does not depend on torch's
    RNG stream: PyTorch's default Conv2d/Linear init is U(-1/sqrt(fan_in),
    +1/sqrt(fan_in)) for weights and biases; drawn here from numpy's PCG64 so
    fixtures only need to store (board, seed)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = {}
    for name, shape in stem_shapes(board).items():
        layer = name.split(".")[0]
        wshape = stem_shapes(board)[layer + ".weight"]
        fan_in = int(np.prod(wshape[1:]))
        bound = scale / np.sqrt(fan_in)
        out[name] = rng.uniform(-bound, bound, size=shape).astype(np.float32)
    return out


def _t(x):
    return x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))


def stem_forward(params: dict, x) -> tuple[torch.Tensor, torch.Tensor]:
    """alphazero/net.py:55-72.  x: (N,4,B,B) float32.  Returns raw policy
    logits (N, B*B) and value (N,1) in [-1,1].  Both head flattens are NCHW
    (``view(-1, 4*B*B)`` / ``view(N,-1)``): FC input index = ch*B*B + r*B + c."""
    p = {k: _t(v).float() for k, v in params.items()}
    x = _t(x).float()
    with torch.no_grad():
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


# ---- residual tower (BASELINE config 4; oracle-defined, parity unpinned) ---------
def tower_shapes(board: int, channels: int = 128, blocks: int = 10) -> dict:
    """Key layout of the author's old checkpoints (e.g.
    alphazero/models/old/model_20250729_052241.pt): stem
    conv(4->C,3x3,bias)+bn; res{i}.{conv1,bn1,conv2,bn2} with bias-free 3x3
    convs; policy_conv(C->2,1x1)+policy_bn, policy_fc(2A->A);
    value_conv(C->1,1x1)+value_bn, value_fc1(A->64), value_fc2(64->1)."""
    a, c = board * board, channels
    s = {"conv.weight": (c, 4, 3, 3), "conv.bias": (c,)}
    for bn, ch in (("bn", c), ("policy_bn", 2), ("value_bn", 1)):
        for f in ("weight", "bias", "running_mean", "running_var"):
            s[f"{bn}.{f}"] = (ch,)
    for i in range(1, blocks + 1):
        for j in (1, 2):
            s[f"res{i}.conv{j}.weight"] = (c, c, 3, 3)
            for f in ("weight", "bias", "running_mean", "running_var"):
                s[f"res{i}.bn{j}.{f}"] = (c,)
    s.update({
        "policy_conv.weight": (2, c, 1, 1), "policy_conv.bias": (2,),
        "policy_fc.weight": (a, 2 * a), "policy_fc.bias": (a,),
        "value_conv.weight": (1, c, 1, 1), "value_conv.bias": (1,),
        "value_fc1.weight": (64, a), "value_fc1.bias": (64,),
        "value_fc2.weight": (1, 64), "value_fc2.bias": (1,),
    })
    return s


def synthetic_tower_params(board: int, seed: int, channels: int = 128, blocks: int = 10) -> dict:
    rng = np.random.Generator(np.random.PCG64(seed))
    shapes = tower_shapes(board, channels, blocks)
    out = {}
    for name, shape in shapes.items():
        leaf = name.split(".")[-1]
        if leaf == "running_var":
            out[name] = rng.uniform(0.5, 1.5, size=shape).astype(np.float32)
        elif leaf == "running_mean":
            out[name] = rng.uniform(-0.1, 0.1, size=shape).astype(np.float32)
        elif "bn" in name.split(".")[-2]:
            lo, hi = (0.8, 1.2) if leaf == "weight" else (-0.1, 0.1)
            out[name] = rng.uniform(lo, hi, size=shape).astype(np.float32)
        else:
            wname = name.rsplit(".", 1)[0] + ".weight"
            fan_in = int(np.prod(shapes[wname][1:]))
            # He-style bound keeps activations O(1) through 10 residual blocks
            bound = np.sqrt(3.0 / fan_in) if leaf == "weight" and len(shape) == 4 else 1.0 / np.sqrt(fan_in)
            out[name] = rng.uniform(-bound, bound, size=shape).astype(np.float32)
    return out


def _bn(h, p, name, eps=1e-5):
    return F.batch_norm(h, p[f"{name}.running_mean"], p[f"{name}.running_var"],
                        p[f"{name}.weight"], p[f"{name}.bias"], training=False, eps=eps)


def tower_forward(params: dict, x, blocks: int = 10):
    """Eval-mode forward of the tower.  Block body as legacy/resnet/example.py:23-27:
    relu(bn1(conv1(x))) -> bn2(conv2(.)) -> relu(. + x)."""
    p = {k: _t(v).float() for k, v in params.items()}
    x = _t(x).float()
    with torch.no_grad():
        h = F.relu(_bn(F.conv2d(x, p["conv.weight"], p["conv.bias"], padding=1), p, "bn"))
        for i in range(1, blocks + 1):
            y = F.relu(_bn(F.conv2d(h, p[f"res{i}.conv1.weight"], None, padding=1), p, f"res{i}.bn1"))
            y = _bn(F.conv2d(y, p[f"res{i}.conv2.weight"], None, padding=1), p, f"res{i}.bn2")
            h = F.relu(y + h)
        n = h.shape[0]
        pol = F.relu(_bn(F.conv2d(h, p["policy_conv.weight"], p["policy_conv.bias"]), p, "policy_bn"))
        logits = F.linear(pol.reshape(n, -1), p["policy_fc.weight"], p["policy_fc.bias"])
        val = F.relu(_bn(F.conv2d(h, p["value_conv.weight"], p["value_conv.bias"]), p, "value_bn"))
        val = F.relu(F.linear(val.reshape(n, -1), p["value_fc1.weight"], p["value_fc1.bias"]))
        val = torch.tanh(F.linear(val, p["value_fc2.weight"], p["value_fc2.bias"]))
    return logits, val


def make_evaluator(params: dict, forward=stem_forward):
    """alphazero/controller.py:33-55 -- encode, batch of one, softmax over ALL
    B*B logits (illegal cells included), value as a Python float."""
    tparams = {k: _t(v).float() for k, v in params.items()}

    def evaluator(state):
        x = torch.from_numpy(state.planes()).unsqueeze(0)
        logits, value = forward(tparams, x)
        policy = torch.softmax(logits[0], dim=0).numpy().reshape(state.size, state.size)
        return policy, float(value.item())

    return evaluator
