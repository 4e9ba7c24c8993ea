import sys, os, time
sys.path.insert(0, "/root/repo")
import torch, numpy as np
import alphazero_piskvorky_b200 as az
torch.manual_seed(0)
net = az.GomokuNet(15)
eng = az.Engine(board=15, win_length=5, games=2048, sims=800, net_dtype=az.DTYPE_BF16, augment=az.AUGMENT_D4, seed=0)
eng.load_state_dict(net.state_dict())
cap = eng.example_capacity_floor() * 2
out = (torch.empty((cap, 4, 15, 15), device="cuda"), torch.empty((cap, 15, 15), device="cuda"), torch.empty((cap,), dtype=torch.int32, device="cuda"))
t0 = time.time(); first = True
for step in range(int(sys.argv[1])):
    eng.selfplay(1 << 40, out=out, resume=not first, max_move_steps=1); first = False
    if step % 10 == 9:
        c = eng.counters()
        print(step + 1, f"{time.time()-t0:.1f}s games={c['games_finished']} plies={c['plies']} positions={c['positions']} "
              f"mean_len={c['positions']/max(1,c['games_finished']):.1f} term_frac={c['terminal_leaves']/max(1,c['simulations']):.3f} "
              f"depth={c['path_length']/max(1,c['simulations']):.2f} xwin={c['x_wins']} owin={c['o_wins']} draw={c['draws']}", flush=True)
