"""BASELINE config 5 on one GPU's share: 19x19, WIN 5, 1600 sims/move, new-vs-best arena, no root noise,
T = 0.3*exp(-step/4); games alternate the first mover (evaluator.py:64-69)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import alphazero_piskvorky_b200 as az
games = int(sys.argv[1]) if len(sys.argv) > 1 else 50
sims = int(sys.argv[2]) if len(sys.argv) > 2 else 1600
torch.manual_seed(0); best = az.GomokuNet(19)
torch.manual_seed(1); new = az.GomokuNet(19)
eng = az.Engine(board=19, win_length=5, games=games, sims=sims, c_puct=2.0, net_dtype=az.DTYPE_BF16, seed=0)
eng.load_state_dict(new.state_dict(), slot=0)      # candidate plays X
eng.load_state_dict(best.state_dict(), slot=1)     # baseline plays O
eng.set_temperature_table([0.3 * np.exp(-s / 4) for s in range(19 * 19 + 1)])
torch.cuda.synchronize(); t0 = time.time()
w, l, d = eng.arena(games)
dt = time.time() - t0
c = eng.counters()
wr = (w + 0.5 * d) / games
print(json.dumps({"config": f"19x19 WIN5 {sims} sims arena, {games} games on 1 GPU (bf16)", "seconds": dt, "wld": [w, l, d],
                  "win_rate": wr, "promote": wr > 0.55, "sims_per_sec": c["simulations"] / dt, "plies": c["plies"],
                  "mean_game_plies": c["plies"] / games, "leaf_evals": c["leaf_evals"]}))
