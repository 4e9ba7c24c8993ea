#!/usr/bin/env python3
"""bench.py -- MCTS simulations/sec and self-play games/sec on 15x15 Gomoku (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # this engine (one rank per GPU)
    python bench.py --impl reference ...                      # reference-side CPU arm (oracle port)

A "step" is one MOVE STEP of the hot path: every concurrent game searches its current position
with `sims` MCTS simulations (root noise on), picks a move, plays it; finished games emit their
(state, pi, z) examples and are re-seated.  Workload = BASELINE config 3 per GPU: 15x15, WIN 5,
800 sims/move, 2048 concurrent games per GPU (weak scaling: 16384 games on 8 GPUs), random-init
GomokuNet(15) (torch.manual_seed(0)), synthetic games from the empty board.
Prints ONE JSON line (rank 0).
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BOARD, WIN, SIMS, GAMES_PER_GPU, C_PUCT = 15, 5, 800, 2048, 2.0
METRIC = "mcts_simulations_per_sec"
UNIT = "sims/s"
MEAN_GAME_PLIES = 57.3             # measured at this config: profiles/r1_game_length_probe.txt
# DRAM bytes (read + write) of ONE conv3 launch on 2048 leaves, from `ncu --set full` (cold caches):
# profiles/r1c_ncu_full_step_kernels.txt -> 67.32 MB + 1.09 MB; algorithmic = act2 read once (67.1 MB)
CONV3_DRAM_TRAFFIC_PER_2048_LEAVES = 68.41e6
FLOPS_PER_LEAF = 42.799e6          # 2*MAC of GomokuNet(15) per position


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "50"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# =================================================================================================
# CPU arm: the oracle port of the reference's per-process inner loop (BASELINE.md section 3, C3 method)
def _cpu_worker(args):
    seed, budget_s, board, win, sims = args
    import numpy as np
    import torch
    torch.set_num_threads(1)                       # as the reference worker, self_play.py:32
    from oracle import gomoku, mcts, net, selfplay
    torch.manual_seed(0)
    sys.path.insert(0, ROOT)
    from alphazero_piskvorky_b200.net import GomokuNet
    params = {k: v.numpy() for k, v in GomokuNet(board).state_dict().items()}
    ev = net.make_evaluator(params)
    np.random.seed(seed)
    search = mcts.Search(ev, sims, C_PUCT)
    state, ply, runs = gomoku.State(board, win), 0, 0
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < budget_s:
        res = search.run(state, selfplay.default_temperature_schedule(ply), add_root_noise=True)
        runs += 1
        state = state.play(res.action)
        ply += 1
        if state.is_terminal():
            state, ply = gomoku.State(board, win), 0
    return runs, time.perf_counter() - t0


def cpu_port_throughput(budget_s=15.0, procs=None, board=BOARD, win=WIN, nsims=SIMS):
    import multiprocessing as mp
    cores = procs or os.cpu_count() or 1
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        out = pool.map(_cpu_worker, [(1000 + i, budget_s, board, win, nsims) for i in range(cores)])
    sims = sum(r for r, _ in out) * nsims
    wall = max(t for _, t in out)
    return sims / wall, cores, (f"{cores} processes x {budget_s:.0f}s of MCTS.run on {board}x{board}/{nsims} sims, "
                                "1 torch thread each")


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_step = []
    budget = max(4.0, min(20.0, 60.0 / max(1, args.steps + args.warmup)))
    for i in range(args.warmup + args.steps):
        v, cores, sample = cpu_port_throughput(budget_s=budget)
        if i >= args.warmup:
            per_step.append(v)
    value = sum(per_step) / len(per_step)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * GAMES_PER_GPU * SIMS / value, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "config 3: 15x15 Gomoku, WIN 5, 800 sims/move, self-play from the empty board "
                               "(CPU arm: bounded sample per step)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# =================================================================================================
def run_engine_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import alphazero_piskvorky_b200 as az

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dtype = az.DTYPE_BF16 if args.dtype == "bf16" else az.DTYPE_F32
    games = args.games
    BOARD, WIN = args.board, args.win          # default: config 3 (15x15 / 5)
    torch.manual_seed(0)
    tower = args.net == "tower"
    net = az.GomokuTowerNet(BOARD, 128, 10) if tower else az.GomokuNet(BOARD)
    net_kw = dict(net_kind=az.NET_TOWER, tower_blocks=10, tower_channels=128) if tower else {}
    blob_host = torch.from_numpy(az.flatten_state_dict(net.state_dict(), az.NET_TOWER if tower else az.NET_STEM, 10)).pin_memory()

    import threading
    E = max(1, args.engines)
    assert games % E == 0
    # E engines per GPU = E independent groups of games/E concurrent games, each driven by its own host
    # thread on its own CUDA stream, so one group's latency-bound tree kernels overlap the other's convs
    engines, streams, outs = [], [], []
    for k in range(E):
        eng = az.Engine(board=BOARD, win_length=WIN, games=games // E, sims=args.sims, c_puct=C_PUCT, net_dtype=dtype,
                        augment=az.AUGMENT_D4, seed=0, rank=rank * E + k, world=world * E, device=local,
                        leaves_per_step=args.leaves_per_step, **net_kw)
        eng.load_blob(blob_host.numpy())
        engines.append(eng)
        streams.append(torch.cuda.Stream())
        cap = eng.example_capacity_floor() * 2
        # two output slabs per engine: step i+1 fills one while step i's examples drain to the host from the other
        outs.append([(torch.empty((cap, 4, BOARD, BOARD), device="cuda"), torch.empty((cap, BOARD, BOARD), device="cuda"),
                      torch.empty((cap,), dtype=torch.int32, device="cuda")) for _ in range(2)])
    cap = outs[0][0][0].shape[0]
    host_out = [(torch.empty((cap, 4, BOARD, BOARD)).pin_memory(), torch.empty((cap, BOARD, BOARD)).pin_memory(),
                 torch.empty((cap,), dtype=torch.int32).pin_memory()) for _ in range(2)]
    copy_streams = [torch.cuda.Stream() for _ in range(E)]
    blob_dev = torch.empty(blob_host.numel(), device="cuda")

    def upload_weights(k):
        """e2e arm: the step's weights travel pinned host -> device on the engine's stream and are repacked there
        (azp_load_weights_device: in place, no host synchronisation)."""
        blob_dev.copy_(blob_host, non_blocking=True)
        engines[k].load_blob(blob_dev)
    total_games = 1 << 40                    # endless stream of games: finished slots are re-seated
    first_flag = [True] * E

    def move_steps(k, n, start_event, done_events, e2e, d2h_acc, profile_last=False):
        """n move steps of engine k on its stream (runs in its own host thread)."""
        torch.cuda.set_device(local)
        with torch.cuda.stream(streams[k]):
            if start_event is not None:
                streams[k].wait_event(start_event)
            drained = [None, None]                                     # per slab: event "its examples are on the host"
            for i in range(n):
                slab = i & 1
                if drained[slab] is not None:
                    streams[k].wait_event(drained[slab])               # never overwrite a slab that is still draining
                if profile_last and i == n - 1:
                    # the dominant kernel is bracketed with CUDA events on the LAST timed step only: the event nodes
                    # cut the programmatic-dependent-launch edges around it, so the other steps run unperturbed
                    engines[k].set_profiling(True)
                if e2e:
                    upload_weights(k)                                  # weights: host -> device, every step
                st, _, _, _, w = engines[k].selfplay(total_games, first_game=0, out=outs[k][slab],
                                                     resume=not first_flag[k], max_move_steps=1)
                first_flag[k] = False
                if e2e and w:
                    # examples: device -> pinned host on a copy stream, overlapping the next move step (the driver is
                    # synchronous at return, so the slab is complete); everything drains inside the timed region
                    ready = torch.cuda.Event()
                    ready.record(streams[k])
                    copy_streams[k].wait_event(ready)
                    with torch.cuda.stream(copy_streams[k]):
                        for src, dst in zip(outs[k][slab], host_out[slab]):
                            dst[:w].copy_(src[:w], non_blocking=True)
                        drained[slab] = torch.cuda.Event()
                        drained[slab].record(copy_streams[k])
                    d2h_acc[k] += w * (5 * BOARD * BOARD + 1) * 4
            for ev in drained:
                if ev is not None:
                    streams[k].wait_event(ev)
            done_events[k].record(streams[k])

    def run_all(n, e2e=False, profile_last=False):
        """All engines advance n move steps concurrently; returns device milliseconds (start -> all done)."""
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        done = [torch.cuda.Event() for _ in range(E)]
        d2h_acc = [0] * E
        ev0.record()
        threads = [threading.Thread(target=move_steps, args=(k, n, ev0, done, e2e, d2h_acc, profile_last)) for k in range(E)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        for d in done:
            torch.cuda.current_stream().wait_event(d)
        ev1.record()
        torch.cuda.synchronize()
        return ev0.elapsed_time(ev1), sum(d2h_acc)

    def counters():
        tot = {}
        for eng in engines:
            for key, v in eng.counters().items():
                tot[key] = tot.get(key, 0) + v
        tot["kernel_launches"] = engines[0].counters()["kernel_launches"]      # process-wide counter
        return tot

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- untimed pre-roll to the steady state of the game population (games of every age, finished games
    #      re-seated), then the W warm-up steps; without it every game would be at ply W..W+K and none could end
    if args.preroll > 0:
        run_all(args.preroll)
    run_all(max(args.warmup, 3))

    # ---- kernel-only arm: inputs resident in HBM --------------------------------------------------
    profile_inline = os.environ.get("AZP_BENCH_NOPROF", "0") != "1"
    if profile_inline:               # instantiate the profiled step graph outside the timed region
        for eng in engines:
            eng.set_profiling(True)
        run_all(1)
        for eng in engines:
            eng.get_profile()
            eng.set_profiling(False)
    c0 = counters()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    ms, _ = run_all(args.steps, profile_last=profile_inline)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    c1 = counters()
    kms = klaunches = kflops = 0
    for eng in engines:
        a, b, c = eng.get_profile()
        kms += a; klaunches += b; kflops += c
        eng.set_profiling(False)
    sims_done = c1["simulations"] - c0["simulations"]
    games_done = c1["games_finished"] - c0["games_finished"]
    plies_done = c1["plies"] - c0["plies"]
    evals_done = c1["leaf_evals"] - c0["leaf_evals"]
    launches = c1["kernel_launches"] - c0["kernel_launches"]

    # ---- end-to-end arm: the same move steps through the public call with HOST buffers -----------
    barrier()
    ce0 = counters()
    e2e_ms, d2h = run_all(args.steps, e2e=True)
    barrier()
    ce1 = counters()
    e2e_sims = ce1["simulations"] - ce0["simulations"]

    # ---- side measurement: virtual-loss throughput mode (K leaves per game per step), same workload ------
    vl = None
    if args.leaves_per_step <= 1 and args.vl_leaves > 1 and E == 1 and not args.no_vl:
        for eng in engines:
            eng.close()
        engines.clear()
        torch.cuda.empty_cache()
        eng = az.Engine(board=BOARD, win_length=WIN, games=games, sims=args.sims, c_puct=C_PUCT, net_dtype=dtype,
                        augment=az.AUGMENT_D4, seed=0, rank=rank, world=world, device=local,
                        leaves_per_step=args.vl_leaves, **net_kw)
        eng.load_blob(blob_host.numpy())
        engines.append(eng)
        first_flag[0] = True
        run_all(3)
        barrier()
        v0 = eng.counters()["simulations"]
        vms, _ = run_all(max(2, min(args.steps, 5)))
        barrier()
        vl = {"leaves_per_step": args.vl_leaves, "sims": eng.counters()["simulations"] - v0, "ms": vms}

    # ---- side measurement (rank 0, N = 1, stem net): the consumer of the examples -- one training step of
    #      NeuralNetworkController.train_step on BATCH_SIZE = 1024 positions, device resident
    train = None
    if world == 1 and not tower and not args.no_train:
        for eng in engines:
            eng.close()
        engines.clear()
        torch.cuda.empty_cache()
        tb = 1024
        ctrl = az.NeuralNetworkController(az.GomokuNet(BOARD), device=f"cuda:{local}", batch_size=tb)
        gen = torch.Generator(device="cuda").manual_seed(0)
        X = (torch.rand((tb, 4, BOARD, BOARD), device="cuda", generator=gen) < 0.15).float()
        PI = torch.softmax(torch.randn((tb, BOARD * BOARD), device="cuda", generator=gen), dim=1)
        Z = torch.randint(-1, 2, (tb,), device="cuda", generator=gen).float()
        for _ in range(3):
            ctrl.train_step(X, PI, Z, 5e-5)
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(10):
            ctrl.train_step(X, PI, Z, 5e-5)
        t1.record()
        torch.cuda.synchronize()
        tms = t0.elapsed_time(t1) / 10
        train = {"what": "NeuralNetworkController.train_step (fp32 forward + loss + backward + AdamW, hand-written CUDA), "
                         "1024 positions resident in HBM, loss read back every step",
                 "ms_per_step": tms, "positions_per_sec": tb / tms * 1e3, "batch": tb, "dtype": "f32"}

    # ---- reduce over ranks: totals summed, time = max ----------------------------------------------
    stats = torch.tensor([ms, e2e_ms, float(sims_done), float(e2e_sims), float(games_done), float(plies_done),
                          float(launches), float(evals_done)], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = stats.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, e2e_ms = float(mx[0]), float(mx[1])
        sims_done, e2e_sims, games_done, plies_done, launches, evals_done = (float(v) for v in sm[2:8])
    if vl is not None and world > 1:
        t = torch.tensor([vl["sims"], vl["ms"]], dtype=torch.float64, device="cuda")
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        vl["sims"], vl["ms"] = float(tsum[0]), float(tmax[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    config3 = (BOARD, WIN, args.sims) == (15, 5, 800)
    flops_per_leaf = {5: 4.72e6, 9: 15.314e6, 15: FLOPS_PER_LEAF, 19: 69.061e6}.get(BOARD, FLOPS_PER_LEAF)
    if tower:
        flops_per_leaf = 1329.6e6 * (BOARD * BOARD / 225.0)      # 10x128 tower at 15x15
    value = sims_done / (ms / 1e3)
    achieved_tflops = (kflops / 1e12) / (kms / 1e3) if kms > 0 else 0.0
    peak_tf = peaks["bf16_tflops_sustained"]
    mean_len = None
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, cores, sample = cpu_port_throughput(budget_s=args.cpu_budget, board=BOARD, win=WIN, nsims=args.sims)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic",
        "config": {
            "workload": f"{'config 3: ' if (BOARD, WIN, args.sims, games) == (15, 5, 800, 2048) else ''}"
                        f"{BOARD}x{BOARD} Gomoku, WIN {WIN}, {args.sims} sims/move, {games} concurrent games per GPU "
                        f"({games * world} total), {'residual tower 10x128' if tower else f'GomokuNet({BOARD})'} "
                        f"random init seed 0, Dirichlet root noise on",
            "step": "one move step = every game searches and plays one move",
            "l2": "inputs larger than L2: node table working set is GBs per GPU (no flush needed)",
            "games_per_gpu": games, "sims_per_move": args.sims, "c_puct": C_PUCT, "augment": "d4",
            "engines_per_gpu": E, "preroll_move_steps": args.preroll,
            "leaves_per_step": args.leaves_per_step,
            "search_mode": "exact (visit counts identical to the reference)" if args.leaves_per_step <= 1
                           else f"virtual loss, {args.leaves_per_step} leaves per game per step (throughput mode)",
        },
        "games_per_sec": games_done / (ms / 1e3),
        "games_per_sec_steady_state": (plies_done / (ms / 1e3)) / MEAN_GAME_PLIES if config3 else None,
        "mean_game_plies": MEAN_GAME_PLIES if config3 else None,
        "plies_per_sec": plies_done / (ms / 1e3),
        "leaf_evals_per_sec": evals_done / (ms / 1e3),
        "clocks": clocks,
        "e2e": {"value": e2e_sims / (e2e_ms / 1e3), "unit": UNIT,
                "h2d_bytes_per_step": int(blob_host.numel() * 4), "d2h_bytes_per_step": int(d2h / max(1, args.steps))},
        "gpu_launches": int(launches),
        "roofline": {
            "kernel": ("tower conv 128->128 3x3, streamed weights (20 launches per eval, ~99% of net FLOPs)" if tower
                       else "conv3 64->128 3x3 (dominant: ~77% of net FLOPs)"),
            "bound": "tensor", "achieved": achieved_tflops, "peak": peak_tf, "unit": "TFLOP/s",
            "frac": achieved_tflops / peak_tf if peak_tf else None,
            "traffic": (CONV3_DRAM_TRAFFIC_PER_2048_LEAVES * (games / 2048.0) / max(1, E)
                        if BOARD == 15 and not tower else None),
            "traffic_note": "dram read+write bytes per launch from ncu --set full (profiles/r1c_ncu_full_step_kernels.txt)",
            "peak_source": f"{peaks['source']} sustained bf16 (kernel timed inside a long step)",
            "kernel_ms_per_launch": kms / klaunches if klaunches else None, "kernel_launches": klaunches,
            "kernel_share_of_step": ((kms / klaunches) * (-(-args.sims // max(1, args.leaves_per_step)) + 1) / (ms / args.steps)
                                     if klaunches and ms else None),
            "kernel_timing": "CUDA events around the kernel on the last step of the timed region (event-record nodes in "
                             "the replayed 2-step CUDA graph, sampled once per graph slot, plus the plain launches of that "
                             "step); the other timed steps run without the event nodes",
            "net_tensor_frac_end_to_end": (evals_done / (ms / 1e3)) * flops_per_leaf / 1e12 / world / peak_tf,
        },
        "virtual_loss_mode": None if vl is None else {
            "note": "same workload with K descents per game per step kept apart by virtual loss (throughput mode; "
                    "visit counts are NOT reference-identical, the headline value above is the exact mode)",
            "leaves_per_step": vl["leaves_per_step"], "value": vl["sims"] / (vl["ms"] / 1e3), "unit": UNIT},
        "cpu_baseline": cpu,
        "train_step": train,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--dtype", default=os.environ.get("AZP_BENCH_DTYPE", "bf16"), choices=["bf16", "f32"])
    ap.add_argument("--net", default="stem", choices=["stem", "tower"], help="tower = BASELINE config 4 (10 blocks x 128)")
    ap.add_argument("--board", type=int, default=BOARD)
    ap.add_argument("--win", type=int, default=WIN)
    ap.add_argument("--games", type=int, default=GAMES_PER_GPU)
    ap.add_argument("--engines", type=int, default=int(os.environ.get("AZP_BENCH_ENGINES", "1")),
                    help="engines per GPU, each with games/engines slots on its own CUDA stream (host thread each)")
    ap.add_argument("--sims", type=int, default=SIMS)
    ap.add_argument("--leaves-per-step", type=int, default=int(os.environ.get("AZP_BENCH_LEAVES", "1")),
                    help="1 = exact search (reference-identical visit counts); K > 1 = virtual-loss throughput mode")
    ap.add_argument("--vl-leaves", type=int, default=4, help="K of the virtual-loss side measurement")
    ap.add_argument("--no-vl", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the training-step side measurement")
    ap.add_argument("--preroll", type=int, default=int(os.environ.get("AZP_BENCH_PREROLL", "64")),
                    help="untimed move steps before warm-up: brings the 2048 games to their steady-state age mix")
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_engine_arm(args)


if __name__ == "__main__":
    main()
