"""Multi-GPU plumbing: one process per GPU under torchrun, games sharded by global game id.

The self-play path has NO collective inside the simulation loop.
torch.distributed (NCCL over NVLink on the box, gloo in CPU tests) is used for exactly three
things: all-gather of generated example slabs into every rank's replay buffer, broadcast of
promoted weights from rank 0, and a 3-integer all-reduce of the arena tally.  The training step
 adds one more: the all-reduce of the flat gradient blob."""

from __future__ import annotations

import torch
import torch.distributed as dist


def rank_block(first: int, total: int, rank: int, world: int):
    """Contiguous block of global game ids [first, first+total) owned by ``rank`` -- the same
    arithmetic as the engine's rank_block (csrc/azp_engine.cu)."""
    base, rem = divmod(total, world)
    off = rank * base + min(rank, rem)
    return first + off, base + (1 if rank < rem else 0)


def all_gather_examples(planes, pi, z, group=None):
    """Ragged all-gather: every rank contributes its n_r examples and receives the concatenation
    in rank order.  Slabs are padded to the max count for the collective and trimmed after."""
    world = dist.get_world_size(group)
    dev = planes.device
    n = torch.tensor([planes.shape[0]], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    cap = max(max(counts), 1)

    def gather(t):
        pad = torch.zeros((cap,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
        pad[: t.shape[0]] = t
        outs = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(outs, pad, group=group)
        return torch.cat([o[:c] for o, c in zip(outs, counts)], dim=0)

    return gather(planes), gather(pi), gather(z)


def broadcast_weights(blob: torch.Tensor, src: int = 0, group=None):
    """Promoted weights travel as ONE flat float32 buffer (replaces the pickled state_dict of
    self_play.py:121-127)."""
    dist.broadcast(blob, src=src, group=group)
    return blob


def all_reduce_wld(wins: int, losses: int, draws: int, group=None, device=None):
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([wins, losses, draws], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return int(t[0]), int(t[1]), int(t[2])


def all_reduce_mean(blob: torch.Tensor, group=None):
    """Data-parallel training: average the flat gradient blob over ranks in place (equal local batch sizes
    make this the gradient of the global batch mean)."""
    dist.all_reduce(blob, op=dist.ReduceOp.SUM, group=group)
    blob.div_(dist.get_world_size(group))
    return blob
