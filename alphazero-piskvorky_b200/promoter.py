"""Promotion step of the loop (alphazero/promoter.py:27-62), organised around the two things SURVEY.md section 8 scopes:
the decision ``should_promote(win_rate)`` on the arena's tally (promoter.py:47, strictly greater than 0.55) and the
distribution of a promoted net -- one flat float32 blob broadcast to every rank (``dist.broadcast_weights``), which
replaces the pickled ``state_dict`` the reference hands to its workers (self_play.py:121-127).  The class keeps the
reference's constructor and ``evaluate_and_maybe_promote`` signature; no kernel lives here."""

from __future__ import annotations

import os
import time

import torch

from .constants import PROMOTION_THRESHOLD
from .controller import NeuralNetworkController
from .evaluator import should_promote
from .model_loader import ModelLoader


def checkpoint_name() -> str:
    """``model_<time>.pt`` like the reference's files; microseconds keep two promotions of one second apart."""
    t = time.time()
    return time.strftime("model_%Y%m%d_%H%M%S", time.localtime(t)) + f"_{int((t % 1) * 1e6):06d}.pt"


def publish_weights(controller: NeuralNetworkController, src: int = 0):
    """Every rank adopts rank ``src``'s weights: one NCCL broadcast of the flat parameter blob the module's parameters
    are views of (no pickling, no host copy); cached engines see the change through ``controller.version``."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    from .dist import broadcast_weights
    broadcast_weights(controller.flat_parameters(), src=src)
    controller.version += 1


class ModelPromoter:
    def __init__(self, model_dir, evaluator, net_class, device, threshold=PROMOTION_THRESHOLD, rank=0):
        self.model_dir, self.evaluator, self.net_class = model_dir, evaluator, net_class
        self.device, self.threshold, self.rank = device, threshold, rank
        self.best_path = None
        os.makedirs(model_dir, exist_ok=True)

    def current_best(self, board_size) -> NeuralNetworkController:
        """The net to beat: newest checkpoint of ``model_dir`` (a fresh random init is saved there when it is empty)."""
        return NeuralNetworkController(ModelLoader(self.model_dir, board_size).get_best_model(), device=self.device)

    def save_as_best(self, controller) -> str:
        """Reference-format checkpoint: a plain ``state_dict`` of CPU tensors (controller.py:196-198)."""
        self.best_path = os.path.join(self.model_dir, checkpoint_name())
        torch.save({k: v.detach().cpu() for k, v in controller.net.state_dict().items()}, self.best_path)
        return self.best_path

    def evaluate_and_maybe_promote(self, candidate_controller: NeuralNetworkController, num_games=20, metadata=None,
                                   debug=False):
        """Arena match against the current best; returns ``(win_rate, metrics, was_promoted)``.  The tally is already
        summed over ranks by the evaluator, so every rank takes the same decision; rank 0 alone writes the file."""
        champion = self.current_best(candidate_controller.net.board_size)
        win_rate, metrics = self.evaluator.evaluate(candidate_controller, champion, num_games=num_games, debug=debug)
        promoted = should_promote(win_rate, self.threshold)
        if promoted and self.rank == 0:
            path = self.save_as_best(candidate_controller)
            print(f"[Promoter] win rate {win_rate:.2%} > {self.threshold:.0%}: candidate saved as {path}"
                  + (f" {metadata}" if metadata else ""))
        elif not promoted:
            print(f"[Promoter] win rate {win_rate:.2%}: candidate rejected")
        return win_rate, metrics, promoted
