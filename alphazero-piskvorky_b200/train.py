"""The reference's top-level loop (alphazero/train.py:39-131) on the facade: self-play -> replay buffer ->
``BATCHES_PER_EPISODE`` x ``controller.train(sample, NUM_EPOCHS)`` -> arena vs the best saved model ->
promotion, every stage on the GPU.  ``run_training`` takes the constants as arguments so other board sizes
than constants.py's 5x5 can be run; ``python -m alphazero_piskvorky_b200.train`` runs the reference defaults."""

from __future__ import annotations

import os

from . import constants as K
from .controller import NeuralNetworkController
from .evaluator import ModelEvaluator
from .games import Gomoku
from .model_loader import ModelLoader
from .net import GomokuNet
from .promoter import ModelPromoter
from .replay_buffer import ReplayBuffer
from .self_play import SelfPlayManager


def run_training(num_episodes=K.NUM_EPISODES, board_size=K.BOARD_SIZE, win_length=K.WIN_LENGTH, device=K.DEVICE,
                 model_dir=K.MODEL_DIR, buffer_path="buffer.pkl", num_games=K.NUM_SELF_PLAY_GAMES,
                 num_simulations=K.NUM_SELF_PLAY_SIMULATIONS, c_puct=K.SELF_PLAY_EXPLORATION_CONSTANT,
                 buffer_capacity=K.BUFFER_CAPACITY, batch_size=K.BATCH_SIZE, batches_per_episode=K.BATCHES_PER_EPISODE,
                 num_epochs=K.NUM_EPOCHS, learning_rate=K.LEARNING_RATE, evaluation_games=K.EVALUATION_GAMES,
                 eval_simulations=K.NUM_EVAL_SIMULATIONS, augment="d4", concurrent_games=2048):
    """Returns a list of per-episode summaries ``{"examples", "buffer", "loss", "win_rate", "promoted"}``.
    ``augment="d4"`` (default) emits 8 consistent examples per position; ``"reference"`` reproduces the 4 pairs of
    self_play.py:94-108 as written, three of which pair a rotated state with a mis-rotated pi -- kept for
    parity tests, not for training."""
    print(f"[Main] Using {device}")
    net = ModelLoader(model_dir, board_size).get_best_model()
    evaluator = ModelEvaluator(Gomoku, print_games=False, device=device, num_simulations=eval_simulations,
                               win_length=win_length)
    promoter = ModelPromoter(model_dir=model_dir, evaluator=evaluator, net_class=GomokuNet, device=device)
    controller = NeuralNetworkController(net=net, device=device, batch_size=batch_size)
    self_play_manager = SelfPlayManager(controller, device=device,
                                        mcts_params={"num_simulations": num_simulations, "c_puct": c_puct},
                                        win_length=win_length, augment=augment, concurrent_games=concurrent_games)
    buffer = ReplayBuffer(capacity=buffer_capacity)
    if buffer_path and os.path.exists(buffer_path):
        buffer.load(buffer_path)
    summaries, number_of_promotions = [], 0
    for episode in range(1, num_episodes + 1):
        print(f"----------------------------\nEPISODE: {episode}/{num_episodes} \n---------------------------")
        data = self_play_manager.generate_self_play(num_games=num_games, num_workers=K.NUM_WORKERS)
        buffer.extend(data)
        print(f"[Buffer]: samples added, current size: {len(buffer)}")
        last = None
        for _ in range(batches_per_episode):
            examples = list(buffer.sample_batch(batch_size))
            last = controller.train(examples, epochs=num_epochs, lr=learning_rate)
        win_rate, metrics, was_promoted = promoter.evaluate_and_maybe_promote(
            controller, num_games=evaluation_games, metadata={"episode": episode}, debug=True)
        number_of_promotions += int(was_promoted)
        print(f"[Summary] Win rate: {win_rate:.2%} | Metrics: {metrics}")
        summaries.append({"examples": len(data), "buffer": len(buffer), "loss": last["loss"] if last else None,
                          "win_rate": win_rate, "promoted": was_promoted})
    print(f"[Promoter] Number of promotions during the batch: {number_of_promotions}")
    if buffer_path:
        buffer.save(buffer_path)
    return summaries


if __name__ == "__main__":
    run_training()
