"""Per-kernel event timings of the bf16 net forward (2048 random mid-game-like positions, 15x15), 60 launches.
With a bring-up build (AZP_B200_LIB=build/libazp_b200_bringup.so) AZP_DEBUG_CONV skips work inside the conv kernels
(1: epilogue, 2: all but one UMMA per tile, 4: the TMA loads) -- outputs are then garbage, only the timing is meaningful."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import alphazero_piskvorky_b200 as az
n = int(os.environ.get("AZP_PROBE_N", "2048"))
eng = az.Engine(board=15, win_length=5, games=n, sims=2, net_dtype=az.DTYPE_BF16)
torch.manual_seed(1); eng.load_state_dict(az.GomokuNet(15).state_dict())
r = torch.rand((n, 15, 15), device="cuda", generator=torch.Generator(device="cuda").manual_seed(0))
x = torch.stack([(r < 0.15).float(), ((r >= 0.15) & (r < 0.3)).float(), torch.zeros_like(r), torch.zeros_like(r)], dim=1)
for _ in range(5): eng.net_forward(x)
torch.cuda.synchronize()
eng.set_profiling(True)
for _ in range(60): eng.net_forward(x)
torch.cuda.synchronize()
prof = eng.get_profile_kernels()
print("AZP_DEBUG_CONV=%s " % os.environ.get("AZP_DEBUG_CONV", "0") + "  ".join(f"{k} {v[0] / max(1, v[1]) * 1e3:.1f}us" for k, v in prof.items() if v[1]))
