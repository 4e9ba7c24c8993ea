"""Bring-up diagnostics for the bf16 tcgen05 path: per-layer errors against torch fp32."""
import sys, os
os.environ.setdefault("AZP_DISCARD", "0")      # keep act1 / act2 readable after the evaluation (they are invalidated in L2 by default)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import torch.nn.functional as F
import alphazero_piskvorky_b200 as az
from oracle import net as onet

def run(size, n):
    params = onet.synthetic_stem_params(size, seed=size)
    eng = az.Engine(board=size, win_length=4 if size == 5 else 5, games=max(n, 8), sims=2, net_dtype=az.DTYPE_BF16)
    eng.load_state_dict(params)
    rs = np.random.RandomState(size)
    x = np.zeros((n, 4, size, size), dtype=np.float32)
    occ = rs.random_sample((n, size, size))
    x[:, 0] = occ < 0.15
    x[:, 1] = (occ >= 0.15) & (occ < 0.3)
    for i in range(n):
        x[i, 2, rs.randint(size), rs.randint(size)] = 1
    xt = torch.from_numpy(x).cuda()
    logits, policy, value = eng.net_forward(xt)
    p = {k: torch.from_numpy(v).cuda() for k, v in params.items()}
    h1 = F.relu(F.conv2d(xt, p["conv1.weight"], p["conv1.bias"], padding=1))
    h2 = F.relu(F.conv2d(h1, p["conv2.weight"], p["conv2.bias"], padding=1))
    h3 = F.relu(F.conv2d(h2, p["conv3.weight"], p["conv3.bias"], padding=1))
    pf = F.relu(F.conv2d(h3, p["policy_conv.weight"], p["policy_conv.bias"]))
    vf = F.relu(F.conv2d(h3, p["value_conv.weight"], p["value_conv.bias"]))
    a1 = eng.debug_activations(1, n).view(n, size, size, 32).permute(0, 3, 1, 2)
    a2 = eng.debug_activations(2, n).view(n, size, size, 64).permute(0, 3, 1, 2)
    ft = eng.debug_activations(3, n).view(n, 6, size, size)
    ol, ov = onet.stem_forward(params, x)
    def err(a, b): return float((a - b).abs().max()), float(b.abs().max())
    print(f"board {size} n={n}: act1 {err(a1, h1)} act2 {err(a2, h2)} feat_p {err(ft[:, :4], pf)} feat_v {err(ft[:, 4:], vf)}")
    print(f"   logits {err(logits.cpu(), ol)} value {err(value.cpu(), ov.reshape(-1))} policy {err(policy.cpu(), torch.softmax(ol, 1))}")
    eng.close()

if __name__ == "__main__":
    for size, n in ((15, 3), (15, 300), (9, 37), (5, 11), (19, 9)):
        run(size, n)
