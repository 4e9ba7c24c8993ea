"""Per-tile clock stamps of conv3's CTA 0 (AZP_CLK=1): where do the cycles of a tile go?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["AZP_CLK"] = sys.argv[1] if len(sys.argv) > 1 else "1"
import numpy as np, torch
import alphazero_piskvorky_b200 as az
n = 2048
eng = az.Engine(board=15, win_length=5, games=n, sims=2, net_dtype=az.DTYPE_BF16)
torch.manual_seed(1); eng.load_state_dict(az.GomokuNet(15).state_dict())
x = torch.zeros((n, 4, 15, 15), device="cuda"); x[:, 0, 7, 7] = 1
if os.environ.get("AZP_PROBE_RANDOM"):          # mid-game-like random positions instead of 2048 copies of one board
    r = torch.rand((n, 15, 15), device="cuda", generator=torch.Generator(device="cuda").manual_seed(0))
    x = torch.stack([(r < 0.15).float(), ((r >= 0.15) & (r < 0.3)).float(), torch.zeros_like(r), torch.zeros_like(r)], dim=1)
for _ in range(3): eng.net_forward(x)
torch.cuda.synchronize()
c = eng.debug_activations(9, 1).cpu().numpy()
names = {0: "P:before empty", 1: "P:after empty", 2: "M:before tempty", 3: "M:after tempty", 4: "M:full landed", 5: "M:conv issued+commit",
         6: "M:head issued", 8: "E:before tfull", 9: "E:after tfull", 10: "E:math+st issued", 11: "E:arrived", 12: "E:heads written"}
if os.environ["AZP_CLK"] == "4":        # FC kernel: wall-clock ns of CTA (0,0) from kernel entry
    f = c.reshape(-1)
    print("fc wall ns from entry: prologue done %d, predecessor done %d, first chunk landed %d, all MMAs issued %d, accumulator ready %d, CTA done %d"
          % tuple(int(f[i] - f[0]) for i in range(1, 7)))
    sys.exit(0)
t0 = c[0, 2]
print("tile | " + " | ".join(f"{names[k]:>20s}" for k in sorted(names)))
for tl in range(0, 28):
    print(f"{tl:4d} | " + " | ".join(f"{int(c[tl, k] - t0):20d}" for k in sorted(names)))
ns = float(c[27, 13] - c[1, 13]); cyc = float(c[27, 5] - c[1, 5])
print(f"SM clock during the launch: {cyc / ns * 1e3:.0f} MHz ({cyc:.0f} cycles in {ns:.0f} ns)")
w = [int(c[i, 14] - c[0, 14]) for i in range(5)]
print(f"wall ns from kernel entry (CTA 0): barriers+TMEM {w[1]}, predecessor done {w[2]}, weights resident {w[3]}, all roles done {w[4]};"
      f" first/last issue complete at {int(c[0, 13] - c[0, 14])} / {int(c[27, 13] - c[0, 14])}")
