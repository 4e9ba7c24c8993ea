"""``human_vs_ai`` (alphazero/main.py:26-85): the human plays X, the model answers with the compat
``MCTS.run`` (150 simulations, temperature 0) whose search runs on the GPU."""

from __future__ import annotations

import torch

from .constants import BOARD_SIZE, O, WIN_LENGTH, X
from .controller import NeuralNetworkController, make_policy_value_fn
from .games import Gomoku
from .mcts import MCTS
from .model_loader import ModelLoader
from .net import GomokuNet


def human_vs_ai(model_path=None, model=None, board_size=BOARD_SIZE, win_length=WIN_LENGTH, num_simulations=150,
                input_fn=input, output_fn=print):
    if model is not None:
        net = model
    elif model_path is not None:
        net = GomokuNet(board_size=board_size)
        net.load_state_dict(torch.load(model_path, map_location="cpu", weights_only=True))
        net.eval()
    else:
        net = ModelLoader(board_size=board_size).get_best_model()
    game = Gomoku(board_size, win_length)
    mcts = MCTS(policy_value_fn=make_policy_value_fn(NeuralNetworkController(net, device="cuda")), c_puct=2.0,
                num_simulations=num_simulations)
    while not game.is_terminal():
        output_fn(game)
        if game.current_player == X:
            row, col = map(int, input_fn("Your move (row col): ").strip().split())
            if (row, col) not in game.get_legal_actions():
                output_fn("Illegal move, try again.")
                continue
            game = game.apply_action((row, col))
        else:
            _, action = mcts.run(game, temperature=0)
            output_fn(f"AI plays: {action}")
            game = game.apply_action(action)
    output_fn(game)
    winner = game.get_game_result()
    output_fn("You win!" if winner == X else "You lost." if winner == O else "It's a draw.")
    return winner


if __name__ == "__main__":
    human_vs_ai()
