"""``make_policy_value_fn`` / ``NeuralNetworkController`` (alphazero/controller.py:33-201).

The controller carries the network for the self-play path and -- SURVEY.md section 8f row 1 -- trains it:
``train_step`` / ``train`` keep the reference's signatures, but forward, loss, backward and the AdamW
update run in the library's CUDA kernels (``trainer.py`` -> azp_train_loss_grad / azp_adamw_step) on ONE
flat float32 parameter blob.  The module's parameters are re-pointed at views of that blob, so
``controller.net.state_dict()`` (what self-play, the evaluator and ``save`` read) always holds the trained
weights without copies.  There is no CPU training fallback."""

from __future__ import annotations

import torch

from . import engine as _e
from ._lib import DTYPE_F32, NET_STEM, NET_TOWER
from .constants import BATCH_SIZE, LEARNING_RATE
from .net import GomokuNet, GomokuTowerNet
from .trainer import BETAS, EPS, WEIGHT_DECAY, Trainer, adamw_step

_PARAM_KEYS = tuple(f"{k}.{f}" for k in _e.STEM_KEYS for f in ("weight", "bias"))


class NeuralNetworkController:
    def __init__(self, net, device="cuda", batch_size=BATCH_SIZE, rank=0, world=1):
        self.device = device
        self.net = net
        self.batch_size = batch_size
        self.training_history = []
        self.weight_decay = WEIGHT_DECAY        # controller.py:75
        self.rank, self.world = rank, world     # data-parallel training: gradients are averaged over ranks
        self._flat = self._exp_avg = self._exp_avg_sq = self._grads = None
        self._step = 0
        self._trainer = None
        self.version = 0                        # bumped whenever the weights change (train_step / load): cached engines reload

    # ---- flat parameter blob ------------------------------------------------------------------------
    def _cuda_device(self):
        d = torch.device(self.device if self.device not in (None, "cpu") else "cuda")
        return torch.device("cuda", d.index if d.index is not None else torch.cuda.current_device())

    def _bound(self) -> bool:
        if self._flat is None:
            return False
        params = dict(self.net.named_parameters())
        off = 0
        for k in _PARAM_KEYS:
            p = params[k]
            if p.data.data_ptr() != self._flat.data_ptr() + 4 * off or p.dtype != torch.float32:
                return False
            off += p.numel()
        return True

    def flat_parameters(self) -> torch.Tensor:
        """The float32 CUDA blob (azp_load_weights layout) the module's parameters are views of."""
        if isinstance(self.net, GomokuTowerNet):
            raise NotImplementedError("the CUDA training step covers the stem net (GomokuNet) only")
        if not torch.cuda.is_available():
            raise RuntimeError("azp_b200 needs a CUDA device (B200, sm_100a); there is no CPU training fallback")
        if not self._bound():
            dev = self._cuda_device()
            params = dict(self.net.named_parameters())
            flat = torch.cat([params[k].detach().to(dev, torch.float32).reshape(-1) for k in _PARAM_KEYS]).contiguous()
            self.net.to(dev)
            off = 0
            for k in _PARAM_KEYS:
                p = params[k]
                p.data = flat[off:off + p.numel()].view(p.shape)
                off += p.numel()
            if self._flat is None or self._flat.numel() != flat.numel() or self._flat.device != flat.device:
                self._exp_avg, self._exp_avg_sq = torch.zeros_like(flat), torch.zeros_like(flat)
                self._step = 0
            self._flat = flat
            self._grads = torch.empty_like(flat)
        return self._flat

    def _get_trainer(self, n):
        b = self.net.board_size
        if self._trainer is None or self._trainer.board != b or self._trainer.max_batch < n:
            if self._trainer is not None:
                self._trainer.close()
            self._trainer = Trainer(b, max(n, self.batch_size), self._flat.device)
        return self._trainer

    # ---- training (controller.py:82-187) ------------------------------------------------------------------
    def loss_and_grads(self, state_batch, mcts_probs, winner_batch):
        """Forward + loss + backward on the device; returns (losses float32[4] CUDA tensor = policy, value,
        entropy, total; flat gradient blob).  With world > 1 the gradients are averaged over ranks (NCCL)."""
        flat = self.flat_parameters()
        n = int(state_batch.shape[0])
        chunk = max(self.batch_size, 1)
        with torch.cuda.device(flat.device):
            if n <= chunk:
                tr = self._get_trainer(n)
                losses, grads = tr.loss_grad(flat, state_batch, mcts_probs, winner_batch, self._grads)
            else:
                # a batch larger than the workspace (the im2col buffer alone is 9*128*4 bytes per board cell) is
                # processed in chunks of batch_size; batch means combine with the weights n_i / n
                tr = self._get_trainer(chunk)
                part = torch.empty_like(self._grads)
                grads, losses = self._grads.zero_(), torch.zeros(4, device=flat.device)
                pi2 = torch.as_tensor(mcts_probs).reshape(n, -1)
                for i in range(0, n, chunk):
                    j = min(n, i + chunk)
                    l, g = tr.loss_grad(flat, state_batch[i:j], pi2[i:j], torch.as_tensor(winner_batch)[i:j], part)
                    grads.add_(g, alpha=(j - i) / n)
                    losses.add_(l, alpha=(j - i) / n)
            if self.world > 1:
                from .dist import all_reduce_mean
                all_reduce_mean(grads)
        return losses, grads

    def train_step(self, state_batch, mcts_probs, winner_batch, lr):
        """controller.py:82-131: one AdamW step on the batch; returns ``(loss, entropy)`` floats.  ``entropy`` is
        the mean entropy of the predicted policy (the reference's logging line, controller.py:130, takes the log
        of raw logits and usually yields NaN -- not replicated)."""
        losses, grads = self.loss_and_grads(state_batch, mcts_probs, winner_batch)
        self._step += 1
        adamw_step(self._flat, grads, self._exp_avg, self._exp_avg_sq, self._step, lr, BETAS, EPS, self.weight_decay)
        self.version += 1
        host = losses.cpu()
        return float(host[3]), float(host[2])

    def evaluate_losses(self, state_batch, mcts_probs, winner_batch):
        """(policy loss, value loss) of the current weights, forward only."""
        flat = self.flat_parameters()
        n = int(state_batch.shape[0])
        chunk = max(self.batch_size, 1)
        tr = self._get_trainer(min(n, chunk))
        pi2 = torch.as_tensor(mcts_probs).reshape(n, -1)
        total = torch.zeros(4, device=flat.device)
        with torch.cuda.device(flat.device):
            for i in range(0, n, chunk):
                j = min(n, i + chunk)
                l, _ = tr.loss_grad(flat, state_batch[i:j], pi2[i:j], torch.as_tensor(winner_batch)[i:j], want_grads=False)
                total.add_(l, alpha=(j - i) / n)
        host = total.cpu()
        return float(host[0]), float(host[1])

    def train(self, examples, epochs=10, lr=LEARNING_RATE, generator=None):
        """controller.py:134-187: ``epochs`` shuffled passes over ``examples`` (list of ``(state (4,B,B), pi (B,B),
        z)``) in batches of ``batch_size``, LEARNING_RATE from constants.py, per-epoch history."""
        b = self.net.board_size
        for example in examples:
            if tuple(example[0].shape) != (4, b, b):
                raise ValueError(f"[Controller] Example state shape {tuple(example[0].shape)} does not match expected "
                                 f"(4, {b}, {b}). Please check your data preparation.")
        states = torch.stack([torch.as_tensor(e[0]).float() for e in examples])
        pis = torch.stack([torch.as_tensor(e[1]).float() for e in examples])
        zs = torch.tensor([float(e[2]) for e in examples], dtype=torch.float32)
        dev = self.flat_parameters().device
        return self.train_tensors(states.to(dev), pis.to(dev), zs.to(dev), epochs, lr, generator)

    def train_tensors(self, states, pis, zs, epochs=10, lr=LEARNING_RATE, generator=None):
        """Same loop on tensors already on the device (e.g. ``ReplayBuffer.sample_device``)."""
        n = states.shape[0]
        self.net.train()
        for _ in range(epochs):
            perm = torch.randperm(n, generator=generator).to(states.device)      # DataLoader(shuffle=True)
            total = pol = val = 0.0
            batches = 0
            for i in range(0, n, self.batch_size):
                sel = perm[i:i + self.batch_size]
                loss, _ = self.train_step(states[sel], pis[sel], zs[sel], lr)
                p, v = self.evaluate_losses(states[sel], pis[sel], zs[sel])
                total, pol, val, batches = total + loss, pol + p, val + v, batches + 1
            self.training_history.append({"loss": total / batches, "policy": pol / batches, "value": val / batches})
        return self.training_history[-1] if self.training_history else None

    # ---- checkpoints (controller.py:196-201) ------------------------------------------------------------
    def save(self, path):
        torch.save({k: v.detach().cpu() for k, v in self.net.state_dict().items()}, path)

    def load(self, path):
        self.net.load_state_dict(torch.load(path, map_location="cpu", weights_only=True))
        self.version += 1

    def eval(self):
        self.net.eval()

    def train_mode(self):
        self.net.train()


def net_kind_of(net):
    return NET_TOWER if isinstance(net, GomokuTowerNet) else NET_STEM


def weights_fingerprint(controller):
    """Cheap identity of the weights ``controller.net`` holds RIGHT NOW.  The reference evaluates ``controller.net``
    live on every call (controller.py:39-53), so an engine that caches uploaded weights must notice every way they can
    change: the library's raw-pointer AdamW kernels and ``load`` (``controller.version``), ``net.load_state_dict`` /
    a torch optimizer / any in-place op (tensor version counters), re-pointed ``.data`` (data pointers) and a
    reassigned ``controller.net`` (object identity)."""
    net = controller.net
    tensors = list(net.parameters()) + list(net.buffers())
    return (id(net), getattr(controller, "version", 0), sum(t._version for t in tensors),
            sum(t.data_ptr() for t in tensors))


def make_policy_value_fn(controller, dtype=DTYPE_F32):
    """``policy_value_fn(state) -> (policy (B,B) float32 summing to 1 over ALL cells, value float)``
    -- batch of one through the engine's CUDA net (controller.py:39-53)."""
    cache = {}

    def policy_value_fn(state):
        net = controller.net                                   # read at call time, as the reference does
        b = state.board_size
        key = (b, type(net), getattr(net, "blocks", 0), getattr(net, "channels", 0))
        eng = cache.get(key)
        if eng is None:
            for old in cache.values():                         # another architecture / size: one engine at a time
                old.close()
            cache.clear()
            kw = {}
            if isinstance(net, GomokuTowerNet):
                kw = dict(net_kind=NET_TOWER, tower_blocks=net.blocks, tower_channels=net.channels)
            eng = cache[key] = _e.Engine(board=b, win_length=state.win_length, games=1, sims=1, net_dtype=dtype, **kw)
            eng.weights_version = None
        fp = weights_fingerprint(controller)
        if eng.weights_version != fp:
            eng.load_state_dict(net.state_dict())
            eng.weights_version = fp
        planes = state.encode().unsqueeze(0)
        _, policy, value = eng.net_forward(planes)
        return policy[0].cpu().numpy().reshape(b, b), float(value[0].item())

    policy_value_fn.controller = controller      # lets MCTS run the whole search on the device
    return policy_value_fn
