#!/usr/bin/env python3
"""Training-step benchmark: NeuralNetworkController.train_step on a
batch of BATCH_SIZE = 1024 positions (constants.py:22) of 15x15, inputs resident in HBM, CUDA events;
the oracle's PyTorch-CPU step (the reference's own arithmetic) timed beside it on all host cores.

    python tools/train_bench.py [--board 15] [--batch 1024] [--steps 10] [--warmup 3] [--no-cpu]
Prints one JSON line."""

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch


def flops_per_position(board):
    a = board * board
    conv = 2 * a * 9 * (4 * 32 + 32 * 64 + 64 * 128)
    heads = 2 * a * 128 * 6
    fc = 2 * (4 * a * a + 2 * a * 64 + 64)
    return conv + heads + fc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--board", type=int, default=15)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()

    import alphazero_piskvorky_b200 as az
    from oracle import net as onet, train as otrain

    torch.cuda.set_device(0)
    params = onet.synthetic_stem_params(args.board, seed=0)
    planes, pi, z = otrain.synthetic_batch(args.board, args.batch, seed=1)
    net = az.GomokuNet(board_size=args.board)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()})
    ctrl = az.NeuralNetworkController(net, device="cuda", batch_size=args.batch)
    X, PI, Z = (torch.from_numpy(t).cuda() for t in (planes, pi, z))
    for _ in range(args.warmup):
        ctrl.train_step(X, PI, Z, 5e-5)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(args.steps):
        losses, grads = ctrl.loss_and_grads(X, PI, Z)
        ctrl._step += 1
        az.adamw_step(ctrl._flat, grads, ctrl._exp_avg, ctrl._exp_avg_sq, ctrl._step, 5e-5)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / args.steps
    fl = 3.0 * flops_per_position(args.board) * args.batch          # forward + input grads + weight grads
    out = {"metric": "train_step_ms", "board": args.board, "batch": args.batch, "steps": args.steps,
           "ms_per_step": ms, "positions_per_sec": args.batch / ms * 1e3, "dtype": "f32",
           "algorithmic_tflops": fl / ms / 1e9, "loss": float(losses[3])}
    if not args.no_cpu:
        st = otrain.TrainerState(params)
        st.train_step(planes, pi, z, 5e-5)
        t0 = time.perf_counter()
        reps = 2
        for _ in range(reps):
            st.train_step(planes, pi, z, 5e-5)
        cpu_ms = (time.perf_counter() - t0) / reps * 1e3
        out["cpu_baseline"] = {"ms_per_step": cpu_ms, "kind": "port", "cores": torch.get_num_threads(),
                               "sample": f"{reps} steps of oracle.train (PyTorch CPU autograd + AdamW) on the same batch"}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
