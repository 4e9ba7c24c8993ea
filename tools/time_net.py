"""Time the bf16 net forward (2048 positions, 15x15) with CUDA events; used for kernel bring-up."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import alphazero_piskvorky_b200 as az
n = 2048
eng = az.Engine(board=15, win_length=5, games=n, sims=2, net_dtype=az.DTYPE_BF16)
torch.manual_seed(1); eng.load_state_dict(az.GomokuNet(15).state_dict())
x = torch.zeros((n, 4, 15, 15), device="cuda"); x[:, 0, 7, 7] = 1
for _ in range(5): eng.net_forward(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): eng.net_forward(x)
e1.record(); torch.cuda.synchronize()
print("AZP_DEBUG_FC=%s net_forward(2048): %.1f us" % (os.environ.get("AZP_DEBUG_FC", "0"), e0.elapsed_time(e1) / 50 * 1000))
