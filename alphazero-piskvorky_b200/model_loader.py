"""``ModelLoader`` (alphazero/model_loader.py:9-63): newest ``*.pt`` state_dict in ``MODEL_DIR`` or a fresh
random-init net saved there.  Checkpoints are plain reference ``state_dict`` files."""

from __future__ import annotations

import os
from datetime import datetime

import torch

from .constants import BOARD_SIZE, MODEL_DIR
from .net import GomokuNet


class ModelLoader:
    def __init__(self, model_dir=MODEL_DIR, board_size=BOARD_SIZE):
        self.model_dir = model_dir
        self.board_size = board_size
        self.best_path = self._find_latest_model()

    def get_best_model(self):
        if self.best_path is None:
            print(f"[ModelLoader] {self.model_dir} holds no checkpoint: saving a random init as the first best")
            random_net = GomokuNet(board_size=self.board_size)
            timestamp = datetime.now().strftime("%Y%m%d_%H%M%S")
            torch.save(random_net.state_dict(), os.path.join(self.model_dir, f"model_{timestamp}.pt"))
            return random_net.float()
        print(f"[ModelLoader] best = {self.best_path}")
        model = GomokuNet(board_size=self.board_size)
        model.load_state_dict(torch.load(self.best_path, map_location="cpu", weights_only=True))
        return model.float()

    def _find_latest_model(self):
        # model_loader.py:50 lists the directory unguarded and fails when it is missing; created here instead
        os.makedirs(self.model_dir, exist_ok=True)
        models = [f for f in os.listdir(self.model_dir) if f.endswith(".pt")]
        if not models:
            return None
        latest = max(models, key=lambda f: os.path.getctime(os.path.join(self.model_dir, f)))
        return os.path.join(self.model_dir, latest)
