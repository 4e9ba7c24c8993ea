"""Diagnostic: per-tensor gradient error of the CUDA training step vs a float64 oracle, next to the error of
PyTorch's own fp32 paths (CPU and CUDA) on the same batch."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import alphazero_piskvorky_b200 as az
from oracle import net as onet, train as otrain

size = int(sys.argv[1]) if len(sys.argv) > 1 else 15
for n in [int(a) for a in sys.argv[2:]] or [64, 1024]:
    params = onet.synthetic_stem_params(size, seed=7)
    planes, pi, z = otrain.synthetic_batch(size, n, seed=11)
    tr = az.Trainer(size, n)
    P = torch.from_numpy(az.flatten_state_dict(params)).cuda()
    l, g = tr.loss_grad(P, torch.from_numpy(planes), torch.from_numpy(pi), torch.from_numpy(z))
    g = g.cpu().numpy()
    _, g64 = otrain.loss_and_grads(params, planes, pi, z, dtype=torch.float64)
    _, g32 = otrain.loss_and_grads(params, planes, pi, z, dtype=torch.float32)
    # torch CUDA fp32 (cuDNN / cuBLAS, TF32 off)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    pc = {k: torch.from_numpy(params[k]).cuda().requires_grad_(True) for k in otrain.PARAM_KEYS}
    lg, v = otrain.forward_train(pc, torch.from_numpy(planes).cuda())
    loss = torch.nn.functional.cross_entropy(lg, torch.from_numpy(pi).cuda().flatten(1, -1)) + \
        torch.nn.functional.mse_loss(v.view(-1), torch.from_numpy(z).cuda())
    loss.backward()
    off = 0
    print(f"n={n} loss mine {float(l[3]):.7f} torch-cuda {float(loss):.7f}")
    for k in otrain.PARAM_KEYS:
        m = g64[k].size
        mine = g[off:off + m].reshape(g64[k].shape); off += m
        sc = np.abs(g64[k]).max()
        e = lambda a: np.abs(a - g64[k]).max() / sc
        print(f"  {k:20s} max|g| {sc:.3e}  mine {e(mine):.2e}  torch-cpu32 {e(g32[k]):.2e}  torch-cuda32 {e(pc[k].grad.cpu().numpy()):.2e}")
    tr.close()
