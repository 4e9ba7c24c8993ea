"""``ModelPromoter`` (alphazero/promoter.py:18-62): play the candidate against the newest saved model on the
GPU arena and save it as the new best iff ``win_rate > threshold`` (strictly; 0.55)."""

from __future__ import annotations

import os
from datetime import datetime

import torch

from .constants import PROMOTION_THRESHOLD
from .controller import NeuralNetworkController
from .model_loader import ModelLoader


class ModelPromoter:
    def __init__(self, model_dir, evaluator, net_class, device, threshold=PROMOTION_THRESHOLD):
        self.model_dir = model_dir
        self.evaluator = evaluator
        self.net_class = net_class
        self.threshold = threshold
        self.device = device
        self.best_path = None
        os.makedirs(model_dir, exist_ok=True)

    def evaluate_and_maybe_promote(self, candidate_controller: NeuralNetworkController, num_games=20, metadata=None,
                                   debug=False):
        baseline_net = ModelLoader(self.model_dir, candidate_controller.net.board_size).get_best_model()
        baseline_controller = NeuralNetworkController(baseline_net, device=self.device)
        win_rate, metrics = self.evaluator.evaluate(candidate_controller, baseline_controller, num_games=num_games,
                                                    debug=debug)
        was_promoted = False
        if win_rate > self.threshold:
            timestamp = datetime.now().strftime("%Y%m%d_%H%M%S_%f")
            model_path = os.path.join(self.model_dir, f"model_{timestamp}.pt")
            torch.save({k: v.detach().cpu() for k, v in candidate_controller.net.state_dict().items()}, model_path)
            print(f"[Promoter]: Promoted new model with win rate {win_rate:.2%}: {model_path}")
            was_promoted = True
            self.best_path = model_path
            if metadata:
                print("Metadata:", metadata)
        else:
            print(f"[Promoter]: Candidate rejected (win rate: {win_rate:.2%})")
        return win_rate, metrics, was_promoted
