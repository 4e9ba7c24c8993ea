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


class ExampleExchange:
    """All-gather of generated positions at move-step granularity (replaces the result queue of
    self_play.py:140-148): every rank contributes ONE fixed-capacity slab of RAW positions per move step --

        float32 [4 + cap*(5A+1)]:  header (word 0 = position count as int32) | planes [cap,4,B,B] | pi [cap,B,B] | z [cap] (int32)

    -- the engine writes its positions straight into the slab (``out_views``), the collective runs on a side stream while
    the next move step searches, and the RECEIVING side expands each rank's positions by the symmetry group and appends
    them to its replay ring with one kernel per source rank whose position count is read on the device from the slab
    header (azp_augment_write_ring): no host synchronisation anywhere, 8x fewer bytes on the wire than gathering
    augmented examples, identical rings on every rank."""

    def __init__(self, board, capacity, device, group=None, slabs=2):
        self.board, self.A, self.cap = board, board * board, int(capacity)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.words = 4 + self.cap * (5 * self.A + 1)
        self.device = torch.device(device)
        cuda = self.device.type == "cuda"
        self.send = [torch.zeros(self.words, dtype=torch.float32, device=self.device) for _ in range(slabs)]
        self.recv = [torch.zeros((self.world, self.words), dtype=torch.float32, device=self.device) for _ in range(slabs)]
        self.stream = torch.cuda.Stream(self.device) if cuda else None
        self.done = [None] * slabs             # event: the gather (and ring writes) of this slab pair have finished
        self.bytes_per_step = self.words * 4 * self.world

    def _views(self, flat):
        A, cap, b = self.A, self.cap, self.board
        o = 4
        planes = flat[..., o:o + cap * 4 * A]; o += cap * 4 * A
        pi = flat[..., o:o + cap * A]; o += cap * A
        z = flat[..., o:o + cap].view(torch.int32)
        lead = flat.shape[:-1]
        return (flat[..., 0:1].view(torch.int32), planes.reshape(*lead, cap, 4, b, b), pi.reshape(*lead, cap, b, b), z)

    def out_views(self, k):
        """(planes, pi, z) views of send slab k for ``Engine.selfplay(out=...)`` (engine in AUGMENT_NONE mode)."""
        _, planes, pi, z = self._views(self.send[k])
        return planes, pi, z

    def wait_slab(self, k, stream=None):
        """Before refilling slab k: its previous gather must have read it."""
        if self.done[k] is not None and self.stream is not None:
            (stream or torch.cuda.current_stream(self.device)).wait_event(self.done[k])

    def exchange(self, k, count, ring=None, augment=None):
        """Post slab k holding ``count`` positions (host int: the driver returns it) and append everybody's positions to
        ``ring`` (a ReplayBuffer).  Asynchronous on CUDA: ordered after the caller's current stream, runs on the side stream."""
        hdr, _, _, _ = self._views(self.send[k])
        hdr.fill_(int(count))
        cuda = self.stream is not None
        if cuda:
            self.stream.wait_stream(torch.cuda.current_stream(self.device))
        ctx = torch.cuda.stream(self.stream) if cuda else _null_context()
        with ctx:
            if self.world > 1 and dist.get_backend(self.group) == "gloo":
                dist.all_gather(list(self.recv[k].unbind(0)), self.send[k], group=self.group)
            elif self.world > 1:
                dist.all_gather_into_tensor(self.recv[k].view(-1), self.send[k], group=self.group)
            else:
                self.recv[k][0].copy_(self.send[k])
            if ring is not None:
                hdrs, planes, pi, z = self._views(self.recv[k])
                for r in range(self.world):
                    if cuda:
                        ring.write_slab(planes[r], pi[r], z[r], n_max=self.cap, n_dev=hdrs[r], augment=augment)
                    else:                              # host tensors (gloo tests): the facade's host-side ring
                        n = int(hdrs[r].item())
                        ring.extend_device(planes[r, :n], pi[r, :n], z[r, :n])
            if cuda:
                self.done[k] = torch.cuda.Event()
                self.done[k].record(self.stream)

    def wait(self, stream=None):
        """Order ``stream`` (default: the current one) after every exchange posted so far (before reading the ring)."""
        if self.stream is not None:
            (stream or torch.cuda.current_stream(self.device)).wait_stream(self.stream)

    def counts(self, k):
        """Position counts every rank sent with slab k (host read-back: diagnostics / tests)."""
        if self.stream is not None:
            self.stream.synchronize()
        return [int(v) for v in self._views(self.recv[k])[0].reshape(-1).tolist()]


class _null_context:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False
