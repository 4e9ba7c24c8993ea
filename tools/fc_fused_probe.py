"""Bring-up (make BRINGUP=1 build, AZP_CLK=5): wall-clock timeline of the first fused FC job of conv3's LAST CTA."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["AZP_CLK"] = "5"
import numpy as np, torch
import alphazero_piskvorky_b200 as az
n = 2048
eng = az.Engine(board=15, win_length=5, games=n, sims=6, net_dtype=az.DTYPE_BF16)
torch.manual_seed(1); eng.load_state_dict(az.GomokuNet(15).state_dict())
b = az.GomokuBatch(n, 15, 5)
with torch.cuda.stream(torch.cuda.Stream()):
    for _ in range(2):
        eng.search(b.x, b.o, b.to_move, b.last)
torch.cuda.synchronize()
c = eng.debug_activations(9, 1).cpu().numpy()
f = c[48]
names = ["P: conv loads issued, waiting for features", "P: features complete", "P: all FC chunk loads issued", "M: first chunk landed",
         "M: last chunk issued + commit", "E: before heads of the last conv tile", "E: heads written", "E: accumulator ready",
         "E: outputs stored", "CTA done"]
t0 = f[0]
for k, nm in enumerate(names):
    print(f"{nm:45s} {int(f[k] - t0):8d} ns")
w = [int(c[i, 14] - c[0, 14]) for i in range(5)]
print(f"CTA 0 wall ns from kernel entry: barriers+TMEM {w[1]}, predecessor done {w[2]}, weights resident {w[3]}, all roles done {w[4]}; last CTA's first FC job started waiting at {int(f[0] - c[0, 14])}")
