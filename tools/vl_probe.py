"""Probe: how close is the virtual-loss mode (K=4) to the exact search, as a function of how peaked the net is?
Prints mean L1 distance of the root visit shares and top-1 agreement; used to set the bound of
tests/test_gpu_parity.py::test_virtual_loss_throughput_mode."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import numpy as np, torch
import alphazero_piskvorky_b200 as az
import synth
from oracle import gomoku, net as onet

size, win, sims, n = 15, 5, 400, 64
states = []
for i in range(n):
    s = gomoku.State(size, win)
    for a in synth.random_playout(size, win, plies=(i * 3) % 60, seed=150 + i):
        if s.play(a).is_terminal():
            break
        s = s.play(a)
    states.append(s)
batch = az.GomokuBatch.from_cells(np.stack([s.board for s in states]), to_move=[s.to_move for s in states],
                                  last=[-1 if s.last is None else s.last[0] * size + s.last[1] for s in states], win_length=win)
for scale in (2.0, 4.0, 8.0):
    params = onet.synthetic_stem_params(size, seed=100 + size, scale=scale)
    out = {}
    for K in (1, 4):
        eng = az.Engine(board=size, win_length=win, games=n, sims=sims, c_puct=2.0, net_dtype=az.DTYPE_BF16, leaves_per_step=K)
        eng.load_state_dict(params)
        out[K] = eng.search(batch.x, batch.o, batch.to_move, batch.last).cpu().numpy().astype(np.float64)
        eng.close()
    p1, p4 = out[1] / out[1].sum(1, keepdims=True), out[4] / out[4].sum(1, keepdims=True)
    l1 = np.abs(p1 - p4).sum(1)
    top = (p1.argmax(1) == p4.argmax(1)).mean()
    ent = -(p1 * np.log(p1 + 1e-12)).sum(1).mean()
    print(f"scale {scale}: mean L1 {l1.mean():.3f} max L1 {l1.max():.3f} top1 agree {top:.2f} entropy(exact) {ent:.2f} max share {p1.max(1).mean():.2f}")
