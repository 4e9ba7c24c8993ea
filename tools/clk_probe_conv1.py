"""Per-tile clock stamps of the tensor-core conv1 kernel's CTA 0 (AZP_CLK=3), leaf (bitboard) path."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["AZP_CLK"] = "3"
import numpy as np, torch
import alphazero_piskvorky_b200 as az
n = 2048
eng = az.Engine(board=15, win_length=5, games=n, sims=6, net_dtype=az.DTYPE_BF16)
torch.manual_seed(1); eng.load_state_dict(az.GomokuNet(15).state_dict())
cells = np.zeros((n, 15, 15), dtype=np.int8)
cells[:, 7, 7] = 1
batch = az.GomokuBatch.from_cells(cells, to_move=[2] * n, last=[7 * 15 + 7] * n, win_length=5)
eng.search(batch.x, batch.o, batch.to_move, batch.last)
torch.cuda.synchronize()
c = eng.debug_activations(9, 1).cpu().numpy()
names = {0: "P:start", 1: "P:bar1", 2: "P:table", 3: "P:row built", 4: "P:empty ok", 5: "M:tempty ok", 6: "M:full ok",
         7: "M:issued", 8: "E:start", 9: "E:tfull ok", 10: "E:tmem ld", 11: "E:bar1", 12: "E:bar2", 13: "E:tma issued"}
t0 = c[0, 0]
print("tile | " + " | ".join(f"{names[k]:>12s}" for k in sorted(names)))
for tl in range(0, 28):
    print(f"{tl:4d} | " + " | ".join(f"{int(c[tl, k] - t0):12d}" for k in sorted(names)))
