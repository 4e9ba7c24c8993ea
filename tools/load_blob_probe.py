"""Host time of one weight reload (azp_load_weights: host repack + uploads) on a config-3 engine."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import alphazero_piskvorky_b200 as az
net = az.GomokuNet(15)
blob = az.flatten_state_dict(net.state_dict())
eng = az.Engine(board=15, win_length=5, games=2048, sims=800, net_dtype=az.DTYPE_BF16)
eng.load_blob(blob)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20): eng.load_blob(blob)
torch.cuda.synchronize()
print("load_blob ms", (time.perf_counter() - t0) / 20 * 1e3)
