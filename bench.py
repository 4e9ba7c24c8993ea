#!/usr/bin/env python3
"""bench.py -- MCTS simulations/sec and self-play games/sec (BASELINE.json), one JSON line on rank 0.

    python bench.py --gpus N --steps K --warmup W              # this engine, BASELINE config 3 per GPU
    python bench.py --config {1,2,3,4,5} ...                   # one flag = one BASELINE config at its stated size
    python bench.py --impl reference ...                       # the reference's own CPU implementation of the path

A "step" is one MOVE STEP of the hot path: every concurrent game searches its current position with `sims` MCTS
simulations, picks a move and plays it; finished games emit their (state, pi, z) positions and are re-seated.

  config 1   5x5 / WIN 4 / 200 sims, 100 concurrent games (the reference's own CPU-runnable constants.py case)
  config 2   9x9 / WIN 5 / 400 sims, 4096 concurrent games on one GPU
  config 3   15x15 / WIN 5 / 800 sims, 2048 concurrent games per GPU (16384 over 8 GPUs)      <- default, the headline
  config 4   15x15 residual tower 10 x 128, bf16 tcgen05, 800 sims, 2048 games x 4 leaves per step = leaf batch 8192
             (virtual loss: the only way to form that batch from 2048 games; stated in `config`)
  config 5   19x19 / WIN 5 / 1600 sims, eval-and-promote arena: 400 games new-vs-best over the GPUs of the job
             (step = one lockstep move of every arena game; plus the whole match timed once: `arena_match`)

In the timed loop the engine emits RAW positions into a fixed-capacity slab; the slab is all-gathered over the ranks on a
side stream (NCCL; a device copy at N = 1) and every rank expands everybody's positions by the 8 symmetries straight into
its device replay ring; with N > 1 rank 0's weight blob is broadcast and re-packed on the device every few steps.  The
same steps are timed once more without any of that (`collectives.value_without`)."""

from __future__ import annotations

import argparse
import hashlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

C_PUCT = 2.0
METRIC = "mcts_simulations_per_sec"
UNIT = "sims/s"
REF_DIR = os.path.join(ROOT, "baseline", "_ref", "alphazero")       # unmodified reference sources (gitignored, built by build())

CONFIGS = {
    1: dict(board=5, win=4, sims=200, games=100, net="stem", mode="selfplay", leaves=1,
            label="config 1: 5x5, WIN 4, 200 sims/move, 100 concurrent games (reference constants.py)"),
    2: dict(board=9, win=5, sims=400, games=4096, net="stem", mode="selfplay", leaves=1,
            label="config 2: 9x9, WIN 5, 400 sims/move, 4096 concurrent games per GPU"),
    3: dict(board=15, win=5, sims=800, games=2048, net="stem", mode="selfplay", leaves=1,
            label="config 3: 15x15 Gomoku, WIN 5, 800 sims/move, 2048 concurrent games per GPU"),
    4: dict(board=15, win=5, sims=800, games=2048, net="tower", mode="selfplay", leaves=4,
            label="config 4: 15x15 Gomoku, residual tower 10x128, 800 sims/move, 2048 games x 4 leaves/step = leaf batch 8192"),
    5: dict(board=19, win=5, sims=1600, games=400, net="stem", mode="arena", leaves=1,
            label="config 5: 19x19, WIN 5, 1600 sims/move, 400-game arena new-vs-best over the job's GPUs"),
}
# 2*MAC of GomokuNet(B) per position; the 10x128 tower at 15x15
FLOPS_PER_LEAF = {5: 4.72e6, 9: 15.314e6, 15: 42.799e6, 19: 69.061e6}
TOWER_FLOPS_15 = 1329.6e6


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def imported_dram_traffic(board, leaves):
    """dram read+write bytes of ONE launch of the dominant kernel, imported from the committed `ncu --set full` capture
    (profiles/r2_dominant_kernel_dram.json; ncu cannot run inside a timed bench).  Scaled to this launch's leaf count."""
    path = os.path.join(ROOT, "profiles", "r2_dominant_kernel_dram.json")
    if not os.path.exists(path):
        return None, None
    raw = open(path, "rb").read()
    d = json.loads(raw)
    e = d.get(str(board))
    if not e:
        return None, None
    note = (f"imported, not measured in this run: dram__bytes_read.sum + dram__bytes_write.sum of one launch on {e['leaves']} leaves "
            f"from {e['source']} (sha256 of the import file {hashlib.sha256(raw).hexdigest()[:12]}), scaled to {leaves} leaves")
    return e["bytes_per_launch"] * leaves / e["leaves"], note


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
# CPU arms.  Both run the reference's per-process inner loop (one process per core, one torch
# thread each, MCTS.run on successive plies of self-play games for a wall budget):
#   kind "reference": the UNMODIFIED reference sources (baseline/_ref/alphazero, copied there by build());
#   kind "port":      the oracle restatement (~15x faster per core: it does not clone a board per legal child).
class _SimCounter:
    """Wraps an evaluator callable: the reference calls it exactly once per simulation on a non-terminal leaf plus once
    per MCTS.run for the root (mcts.py:109, 137), so progress INSIDE a run is visible at simulation granularity.
    slot[0] = simulations of completed runs, slot[1] = evaluator calls of the run in flight (root eval excluded)."""

    def __init__(self, fn, shared, index):
        self.fn, self.shared, self.i, self.calls = fn, shared, index, 0

    def __call__(self, state):
        self.calls += 1
        self.shared[2 * self.i + 1] = max(0, self.calls - 1)
        return self.fn(state)

    def run_done(self, nsims):
        self.calls = 0
        self.shared[2 * self.i] += nsims
        self.shared[2 * self.i + 1] = 0


def _cpu_arm_worker(kind, index, shared, stop, ready, seed, board, win, sims, arena):
    import contextlib
    import io
    import numpy as np
    import torch
    torch.set_num_threads(1)                       # as the reference worker, self_play.py:32
    counters = []
    if kind == "reference":
        sys.path.insert(0, REF_DIR)                # the reference uses flat imports
        with contextlib.redirect_stdout(io.StringIO()):
            from controller import NeuralNetworkController, make_policy_value_fn     # noqa: the reference's own modules
            from evaluator import temperature_schedule as eval_schedule
            from games import Gomoku
            from mcts import MCTS
            from net import GomokuNet
            from self_play import default_temperature_schedule
            for s in ((1, 0) if arena else (0,)):          # arena: "new" (seed 1) plays X, "best" (seed 0) plays O
                torch.manual_seed(s)
                fn = make_policy_value_fn(NeuralNetworkController(GomokuNet(board_size=board), device="cpu"))
                counters.append(_SimCounter(fn, shared, index))
        searches = [MCTS(c, sims, c_puct=C_PUCT) for c in counters]
        new_game = lambda: Gomoku(board, win)
        play = lambda st, a: st.apply_action(a)
        run = lambda srch, st, T, noise: srch.run(st, T, add_root_noise=noise)[1]
    else:
        from oracle import arena as oarena, gomoku, mcts, net, selfplay
        from alphazero_piskvorky_b200.net import GomokuNet
        for s in ((1, 0) if arena else (0,)):
            torch.manual_seed(s)
            ev = net.make_evaluator({k: v.numpy() for k, v in GomokuNet(board).state_dict().items()})
            counters.append(_SimCounter(ev, shared, index))
        searches = [mcts.Search(c, sims, C_PUCT) for c in counters]
        new_game = lambda: gomoku.State(board, win)
        play = lambda st, a: st.play(a)
        run = lambda srch, st, T, noise: srch.run(st, T, add_root_noise=noise).action
        eval_schedule = oarena.eval_temperature
        default_temperature_schedule = selfplay.default_temperature_schedule
    np.random.seed(seed)
    state, ply = new_game(), 0
    ready.set()
    while not stop.is_set():
        if arena:      # evaluator.py:64-90: the side to move searches with its own net, no noise, T = 0.3 exp(-step/4)
            side = ply & 1
            action = run(searches[side], state, eval_schedule(ply // 2 if ply % 2 == 0 else ply // 2 + 1), False)
        else:
            side = 0
            action = run(searches[0], state, default_temperature_schedule(ply), True)
        counters[side].run_done(sims)
        state, ply = play(state, action), ply + 1
        if state.is_terminal():
            state, ply = new_game(), 0


class CpuArm:
    """`cores` persistent single-threaded processes, each playing plies with MCTS.run forever; sample(d) = simulations
    finished by all of them in d seconds of wall clock."""

    def __init__(self, kind, procs=None, board=15, win=5, nsims=800, arena=False):
        import multiprocessing as mp
        self.kind, self.cores = kind, procs or os.cpu_count() or 1
        self.board, self.nsims, self.arena = board, nsims, arena
        ctx = mp.get_context("spawn")
        self.shared = ctx.Array("q", 2 * self.cores, lock=False)
        self.stop = ctx.Event()
        ready = [ctx.Event() for _ in range(self.cores)]
        self.procs = [ctx.Process(target=_cpu_arm_worker, daemon=True,
                                  args=(kind, i, self.shared, self.stop, ready[i], 1000 + i, board, win, nsims, arena))
                      for i in range(self.cores)]
        for p in self.procs:
            p.start()
        for r, p in zip(ready, self.procs):          # process start + `import torch` are not part of any sample
            t0 = time.perf_counter()
            while not r.wait(0.5):
                if not p.is_alive() or time.perf_counter() - t0 > 300:
                    self.close()
                    raise RuntimeError(f"CPU arm worker ({kind}) did not start")

    def _total(self):
        return sum(self.shared[:])

    def sample(self, seconds):
        a, t0 = self._total(), time.perf_counter()
        time.sleep(seconds)
        b, t1 = self._total(), time.perf_counter()
        return (b - a) / (t1 - t0)

    def describe(self, seconds):
        what = "the UNMODIFIED reference (baseline/_ref/alphazero)" if self.kind == "reference" else "the oracle port"
        loop = "arena plies (two nets, no noise)" if self.arena else "self-play plies (root noise on)"
        return (f"{self.cores} persistent processes, 1 torch thread each, MCTS.run of {what} on successive {loop}, "
                f"{self.board}x{self.board}/{self.nsims} sims; simulations finished in {seconds:.0f}s of wall clock, counted per "
                "evaluator call (one per non-terminal simulation)")

    def close(self):
        self.stop.set()
        for p in self.procs:                     # a worker only looks at `stop` between runs (one run = seconds to minutes)
            if p.is_alive():
                p.terminate()
        for p in self.procs:
            p.join(timeout=5)


def cpu_throughput(kind, budget_s=15.0, procs=None, board=15, win=5, nsims=800, arena=False):
    arm = CpuArm(kind, procs, board, win, nsims, arena)
    try:
        arm.sample(min(2.0, budget_s / 4))           # ramp: every process is inside its first run
        v = arm.sample(budget_s)
        return v, arm.cores, arm.describe(budget_s)
    finally:
        arm.close()


def cpu_port_throughput(budget_s=15.0, procs=None, board=15, win=5, nsims=800):
    return cpu_throughput("port", budget_s, procs, board, win, nsims)


def reference_available():
    return os.path.exists(os.path.join(REF_DIR, "mcts.py"))


def resolve_config(args):
    c = dict(CONFIGS[args.config])
    for k, a in (("board", args.board), ("win", args.win), ("sims", args.sims), ("games", args.games), ("net", args.net),
                 ("leaves", args.leaves_per_step)):
        if a is not None:
            if a != c[k]:
                c["label"] = None                  # not a BASELINE config any more
            c[k] = a
    return c


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    c = resolve_config(args)
    kind = "reference" if reference_available() else "port"
    arena = c["mode"] == "arena"
    # every step is a bounded sample of the workload: `budget` seconds of all cores playing plies (the workers persist
    # across steps, so no step pays process start or `import torch`); warmup + steps stay within a few minutes
    budget = max(3.0, min(20.0, 150.0 / max(1, args.steps + args.warmup)))
    arm = CpuArm(kind, None, c["board"], c["win"], c["sims"], arena)
    per_step, cores, sample = [], arm.cores, arm.describe(budget)
    try:
        for i in range(args.warmup + args.steps):
            v = arm.sample(budget)
            if i >= args.warmup:
                per_step.append(v)
    finally:
        arm.close()
    value = sum(per_step) / len(per_step)
    workload = c["label"] or f"{c['board']}x{c['board']}, WIN {c['win']}, {c['sims']} sims/move"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000.0 * c["games"] * c["sims"] / value, "higher_is_better": True,
        "ms_per_step_note": "extrapolated: time this CPU arm would need for one move step of the GPU arm's game count",
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload + " (CPU arm: bounded sample per step)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# =================================================================================================
def hbm_rooflines(prof, counters, A, NW, mult, peak_gbs):
    """algorithmic bytes of the tree / episode kernels from the engine's own counters over the
    timed region, divided by the CUDA-event time of the bracketed launches (both per launch), against the measured HBM
    copy peak.  Edge record here = N int32 + W float64 + P float32 + child int32 = 20 B (SURVEY's formula assumed 16)."""
    out = {}

    def entry(name, what, bytes_per_launch, ms, launches, formula):
        if launches and ms > 0 and bytes_per_launch:
            gbs = bytes_per_launch / (ms / launches * 1e-3) / 1e9
            out[name] = {"kernel": what, "bound": "hbm", "achieved": gbs, "peak": peak_gbs, "unit": "GB/s",
                         "frac": gbs / peak_gbs, "bytes_per_launch": bytes_per_launch, "us_per_launch": 1e3 * ms / launches,
                         "launches_timed": launches, "bytes_formula": formula}

    sims = max(1, counters["simulations"])
    steps_launches = max(1.0, counters["advance_launches"])
    # advance = expand+backup of step i, descent of step i+1: per simulation 20 B per child slot scanned, 24 B RMW of
    # (N, W) per path edge, 20 B per edge slot of the new node + its 4A logits and value, the leaf position written
    adv_bytes = (20.0 * counters["children_scanned"] + 24.0 * counters["path_length"]
                 + (20.0 * A + 4.0 * A + 4.0 + 16.0 * NW + 8.0) * counters["expansions"]) / steps_launches
    ms, n, _ = prof["advance"]
    n_c1 = prof["conv1"][1]
    fused_tail = n > 4 and n_c1 * 2 < n       # the descent kernels ran conv1 of their leaves as their tail (azp_conv1_tail.cuh)
    if fused_tail:
        c1_bytes = (16.0 * NW + 3.0 + 64.0 * A) * counters["leaf_evals"] / steps_launches
        entry("advance", "descend_conv1_kernel: Node.expand + Node.backup + Node.select (mcts.py:50-82) + conv1 of the emitted leaves "
              "(games.py:86-129 encode -> act1 bf16 [A][32]) as the kernel's tail", adv_bytes + c1_bytes, ms, n,
              "20*children_scanned + 24*path_length + (24*A + 16*NW + 12)*expansions + (16*NW + 3 + 64*A)*leaf_evals, per launch")
    else:
        entry("advance", "advance_kernel: Node.expand + Node.backup + Node.select (mcts.py:50-82)", adv_bytes, ms, n,
              "20*children_scanned + 24*path_length + (24*A + 16*NW + 12)*expansions, per launch")
    ms, n, units = prof["conv1"]
    if n and not fused_tail:
        entry("leaf_gather_conv1", "conv1: leaf bitboards -> encode (games.py:86-129) -> act1 bf16 [A][32]",
              (16.0 * NW + 3.0 + 64.0 * A) * units / n, ms, n, "(16*NW + 3 + 64*A) per leaf")
    ms, n, units = prof["end_move"]
    if n:
        entry("end_move", "end_move_kernel: pi from root visits at T(ply), move sampling, apply (mcts.py:143-183)",
              (4.0 * A + 4.0 * A + 32.0 * NW + 16.0) * units / n, ms, n, "(8*A + 32*NW + 16) per playing game")
    ms, n, units = prof["augment"]
    if n:
        entry("augment_write", "augment_write_kernel: symmetry expansion + (state, pi, z) writer (self_play.py:94-108)",
              (5.0 * A * 4 + 4.0) * (1 + mult) * units / n, ms, n, "(20*A + 4) * (1 + mult) per raw position")
    return out


def run_engine_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import alphazero_piskvorky_b200 as az
    from alphazero_piskvorky_b200 import dist as azd

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    c = resolve_config(args)
    BOARD, WIN, SIMS, K = c["board"], c["win"], c["sims"], max(1, c["leaves"])
    arena = c["mode"] == "arena"
    tower = c["net"] == "tower"
    games = -(-c["games"] // world) if arena else c["games"]           # arena: 400 games over the job's GPUs
    dtype = az.DTYPE_BF16 if args.dtype == "bf16" else az.DTYPE_F32
    A, NW = BOARD * BOARD, (BOARD * BOARD + 63) // 64
    kind = az.NET_TOWER if tower else az.NET_STEM
    net_kw = dict(net_kind=az.NET_TOWER, tower_blocks=10, tower_channels=128) if tower else {}

    def make_blob(seed):
        torch.manual_seed(seed)
        net = az.GomokuTowerNet(BOARD, 128, 10) if tower else az.GomokuNet(BOARD)
        return torch.from_numpy(az.flatten_state_dict(net.state_dict(), kind, 10)).pin_memory()

    blobs_host = [make_blob(0)] + ([make_blob(1)] if arena else [])     # arena: seed 0 = "best", seed 1 = "new"
    eng = az.Engine(board=BOARD, win_length=WIN, games=games, sims=SIMS, c_puct=C_PUCT, net_dtype=dtype,
                    augment=az.AUGMENT_NONE, seed=0, rank=rank, world=world, device=local, leaves_per_step=K, **net_kw)
    if arena:
        eng.load_blob(blobs_host[1].numpy(), slot=0)                    # candidate ("new") plays X
        eng.load_blob(blobs_host[0].numpy(), slot=1)                    # baseline ("best") plays O
        eng.set_temperature_table([0.3 * np.exp(-s / 4.0) for s in range(A + 1)])        # evaluator.py:14-19
    else:
        eng.load_blob(blobs_host[0].numpy())
    stream = torch.cuda.Stream()
    copy_stream = torch.cuda.Stream()
    blob_dev = [torch.empty(b.numel(), device="cuda") for b in blobs_host]
    total_games = 1 << 40                    # endless stream of games: finished slots are re-seated
    mult = 8                                 # D4 on the receiving side
    slab_rows = args.slab if args.slab > 0 else max(4096, 2 * games)     # steady state emits ~games positions per move step
    if not arena:
        eng.set_slab_capacity(slab_rows)     # positions per rank per move step on the wire; games that do not fit wait a step
    raw_cap = eng.example_capacity_floor()   # raw positions one move step can emit at most (AUGMENT_NONE: mult 1)
    exch = ring = None
    if not arena:
        exch = azd.ExampleExchange(BOARD, raw_cap, torch.device("cuda", local))
        ring = az.ReplayBuffer(capacity=max(40_000, 2 * raw_cap * mult))          # BUFFER_CAPACITY, constants.py:19
        # e2e arm: augmented examples leave through the public call into double-buffered slabs, then to pinned host memory
        cap8 = raw_cap * mult
        outs8 = [(torch.empty((cap8, 4, BOARD, BOARD), device="cuda"), torch.empty((cap8, BOARD, BOARD), device="cuda"),
                  torch.empty((cap8,), dtype=torch.int32, device="cuda")) for _ in range(2)]
        host8 = [(torch.empty((cap8, 4, BOARD, BOARD)).pin_memory(), torch.empty((cap8, BOARD, BOARD)).pin_memory(),
                  torch.empty((cap8,), dtype=torch.int32).pin_memory()) for _ in range(2)]
    state = {"first": True}
    wld_host = torch.zeros(3, dtype=torch.int64).pin_memory()

    def one_step(out=None):
        if arena:
            st, wld = eng.arena_steps(total_games, max_move_steps=1, resume=not state["first"])
            state["first"] = False
            return 0, wld
        st, _, _, _, w = eng.selfplay(total_games, first_game=0, out=out, resume=not state["first"], max_move_steps=1)
        state["first"] = False
        return w, None

    def run_steps(n, mode, profile_last=False):
        """n move steps on the engine's stream; mode: "plain" (no exchange), "exchange" (the multi-GPU design: raw slabs
        all-gathered + receiver-side augmentation into the ring + periodic weight broadcast), "e2e" (host buffers).
        Returns (device ms, d2h bytes)."""
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d2h = 0
        with torch.cuda.stream(stream):
            ev0.record(stream)
            drained = [None, None]
            for i in range(n):
                k = i & 1
                if profile_last and i == n - 1:
                    eng.set_profiling(True)      # last 40 simulation steps of this move step: one event pair per kernel
                if mode == "e2e":
                    for s_, (bh, bd) in enumerate(zip(blobs_host, blob_dev)):      # weights: pinned host -> device, every step
                        bd.copy_(bh, non_blocking=True)
                        eng.load_blob(bd, slot=(1 - s_) if arena else 0)
                if arena:
                    _, wld = one_step()
                    if mode == "e2e":
                        wld_host.copy_(torch.tensor(wld, dtype=torch.int64))       # the step's result, read on the host
                        d2h += 24
                    continue
                if mode == "exchange":
                    exch.wait_slab(k, stream)
                    w, _ = one_step(exch.out_views(k))
                    exch.exchange(k, w, ring=ring, augment=az.AUGMENT_D4)
                    if world > 1 and args.bcast_every > 0 and (i + 1) % args.bcast_every == 0:
                        # promoted weights: rank 0's blob -> everyone (self_play.py:121-127), repacked on the device
                        with torch.cuda.stream(exch.stream):
                            if rank == 0:
                                blob_dev[0].copy_(blobs_host[0], non_blocking=True)
                            azd.broadcast_weights(blob_dev[0], src=0)
                        stream.wait_stream(exch.stream)
                        eng.load_blob(blob_dev[0])
                elif mode == "e2e":
                    if drained[k] is not None:
                        stream.wait_event(drained[k])          # never overwrite a slab that is still draining
                    w, _ = one_step(outs8[k])
                    if w:
                        ready = torch.cuda.Event()
                        ready.record(stream)
                        copy_stream.wait_event(ready)
                        with torch.cuda.stream(copy_stream):
                            for src, dst in zip(outs8[k], host8[k]):
                                dst[:w].copy_(src[:w], non_blocking=True)
                            drained[k] = torch.cuda.Event()
                            drained[k].record(copy_stream)
                        d2h += w * (5 * A + 1) * 4
                else:
                    one_step(exch.out_views(k))
            for ev in drained:
                if ev is not None:
                    stream.wait_event(ev)
            if mode == "exchange":
                exch.wait(stream)                              # every gather / ring write is inside the timed region
            ev1.record(stream)
        torch.cuda.synchronize()
        return ev0.elapsed_time(ev1), d2h

    def counters():
        return eng.counters()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    main_mode = "plain" if arena else "exchange"
    # ---- untimed pre-roll to the steady state of the game population (games of every age, finished games re-seated),
    #      then the warm-up steps; without it every game would be at ply W..W+K and none could end
    if args.preroll > 0:
        run_steps(args.preroll, main_mode)
    warm = max(args.warmup, 3)
    run_steps(warm, main_mode)
    eng.set_profiling(True); run_steps(1, main_mode); eng.get_profile_kernels(); eng.set_profiling(False)   # event pool warm

    # ---- timed region: the multi-GPU design (value) --------------------------------------------------------------
    c0 = counters()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    ms, _ = run_steps(args.steps, main_mode, profile_last=True)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    c1 = counters()
    prof = eng.get_profile_kernels()
    eng.set_profiling(False)
    delta = {k: c1[k] - c0[k] for k in c1}
    # launches of the advance kernel in the timed region: sims/K per move step (the last one is expand_backup alone)
    delta["advance_launches"] = args.steps * max(1, -(-SIMS // K) - 1)

    # ---- the same steps without exchange / collectives ---------------------------------------------------------------
    plain = None
    if not arena:
        barrier()
        p0 = counters()["simulations"]
        pms, _ = run_steps(args.steps, "plain")
        barrier()
        plain = (counters()["simulations"] - p0, pms)

    # ---- end-to-end arm: the public call with HOST buffers --------------------------------------------------------------
    if not arena:
        eng.set_augment(az.AUGMENT_D4)
    barrier()
    ce0 = counters()["simulations"]
    e2e_ms, d2h = run_steps(args.steps, "e2e")
    barrier()
    e2e_sims = counters()["simulations"] - ce0
    if not arena:
        eng.set_augment(az.AUGMENT_NONE)

    # ---- config 5: the whole 400-game match, W/L/D all-reduced, winner's weights broadcast -----------------------------
    match = None
    if arena and not args.no_match:
        eng.close()
        torch.cuda.empty_cache()
        eng = az.Engine(board=BOARD, win_length=WIN, games=games, sims=SIMS, c_puct=C_PUCT, net_dtype=dtype, seed=1,
                        rank=rank, world=world, device=local, **net_kw)
        eng.load_blob(blobs_host[1].numpy(), slot=0)
        eng.load_blob(blobs_host[0].numpy(), slot=1)
        eng.set_temperature_table([0.3 * np.exp(-s / 4.0) for s in range(A + 1)])
        barrier()
        m0 = eng.counters()
        t0 = time.perf_counter()
        w_, l_, d_ = eng.arena(c["games"])
        if world > 1:
            w_, l_, d_ = azd.all_reduce_wld(w_, l_, d_)
        win_rate = (w_ + 0.5 * d_) / max(1, w_ + l_ + d_)
        promoted = az.should_promote(win_rate)
        bsrc = blobs_host[1] if promoted else blobs_host[0]
        tb0 = time.perf_counter()
        blob_dev[0].copy_(bsrc, non_blocking=True)
        if world > 1:
            azd.broadcast_weights(blob_dev[0], src=0)
        eng.load_blob(blob_dev[0], slot=1)
        torch.cuda.synchronize()
        barrier()
        t1 = time.perf_counter()
        m1 = eng.counters()
        match = {"games": c["games"], "seconds": t1 - t0, "wins": w_, "losses": l_, "draws": d_, "win_rate": win_rate,
                 "promoted": promoted, "broadcast_and_reload_ms": 1e3 * (t1 - tb0),
                 "sims_this_rank": m1["simulations"] - m0["simulations"], "plies_this_rank": m1["plies"] - m0["plies"],
                 "timing": "host wall clock around azp_arena_run + all-reduce of {W,L,D} + weight broadcast + device repack, "
                           "barrier on both sides"}

    # ---- side measurement: virtual-loss throughput mode, same workload ------------------------------------------------
    vl = None
    if not arena and K <= 1 and args.vl_leaves > 1 and not args.no_vl:
        eng.close()
        torch.cuda.empty_cache()
        eng = az.Engine(board=BOARD, win_length=WIN, games=games, sims=SIMS, c_puct=C_PUCT, net_dtype=dtype,
                        augment=az.AUGMENT_NONE, seed=0, rank=rank, world=world, device=local,
                        leaves_per_step=args.vl_leaves, **net_kw)
        eng.load_blob(blobs_host[0].numpy())
        eng.set_slab_capacity(slab_rows)
        state["first"] = True
        run_steps(3, "plain")
        barrier()
        v0 = eng.counters()["simulations"]
        vms, _ = run_steps(max(2, min(args.steps, 5)), "plain")
        barrier()
        vl = {"leaves_per_step": args.vl_leaves, "sims": eng.counters()["simulations"] - v0, "ms": vms}

    # ---- side measurement (rank 0, N = 1, stem net): one training step on 1024 positions ----
    train = None
    if world == 1 and not tower and not arena and not args.no_train:
        eng.close()
        torch.cuda.empty_cache()
        tb = 1024
        ctrl = az.NeuralNetworkController(az.GomokuNet(BOARD), device=f"cuda:{local}", batch_size=tb)
        gen = torch.Generator(device="cuda").manual_seed(0)
        X = (torch.rand((tb, 4, BOARD, BOARD), device="cuda", generator=gen) < 0.15).float()
        PI = torch.softmax(torch.randn((tb, A), device="cuda", generator=gen), dim=1)
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

    # ---- reduce over ranks: totals summed, time = max --------------------------------------------------------------
    keys = ("simulations", "games_finished", "plies", "kernel_launches", "leaf_evals", "positions")
    stats = torch.tensor([ms, e2e_ms, float(e2e_sims), plain[1] if plain else 0.0, float(plain[0]) if plain else 0.0]
                         + [float(delta[k]) for k in keys], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = stats.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        stats = torch.cat([mx[0:2], sm[2:3], mx[3:4], sm[4:]])
        if vl is not None:
            t = torch.tensor([vl["sims"], vl["ms"]], dtype=torch.float64, device="cuda")
            tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
            tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            vl["sims"], vl["ms"] = float(tsum[0]), float(tmax[1])
    stats = stats.tolist()
    ms, e2e_ms, e2e_sims, plain_ms, plain_sims = stats[:5]
    tot = dict(zip(keys, stats[5:]))
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    flops_per_leaf = TOWER_FLOPS_15 * (A / 225.0) if tower else FLOPS_PER_LEAF.get(BOARD, 42.799e6 * A / 225.0)
    value = tot["simulations"] / (ms / 1e3)
    kms, klaunches, kunits = prof["dominant"]
    dom_flops_per_leaf = 2.0 * 9.0 * (128 * 128 if tower else 64 * 128) * A
    achieved_tflops = (kunits * dom_flops_per_leaf / 1e12) / (kms / 1e3) if kms > 0 else 0.0
    peak_tf = peaks["bf16_tflops_sustained"]
    per_move_launches = (20 if tower else 1) * (-(-SIMS // K) + 1) * (2 if arena else 1)
    traffic, traffic_note = (None, None)
    if klaunches and not tower:
        traffic, traffic_note = imported_dram_traffic(BOARD, kunits / klaunches)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, cores, sample = cpu_throughput("port", budget_s=args.cpu_budget, board=BOARD, win=WIN, nsims=SIMS, arena=arena)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
               "note": "the oracle port is ~15x faster per core than the reference itself (no board clone per legal child)"}
        if reference_available():
            rv, rcores, rsample = cpu_throughput("reference", budget_s=max(args.cpu_budget, 20.0), board=BOARD, win=WIN,
                                                 nsims=SIMS, arena=arena)
            cpu["reference"] = {"value": rv, "unit": UNIT, "cores": rcores, "kind": "reference", "sample": rsample}
    mean_len = tot["positions"] / tot["games_finished"] if (tot["games_finished"] and not arena) else None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if arena else "weak",
        "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {
            "workload": (f"{c['label']} ({games * world} total), " if c["label"] else
                         f"{BOARD}x{BOARD} Gomoku, WIN {WIN}, {SIMS} sims/move, {games} concurrent games per GPU ({games * world} total), ")
                        + f"{'residual tower 10x128' if tower else f'GomokuNet({BOARD})'} random init seed 0"
                        + (" (best) vs seed 1 (new), no root noise, T = 0.3*exp(-step/4)" if arena else ", Dirichlet root noise on"),
            "baseline_config": args.config if c["label"] else None,
            "step": "one move step = every game searches and plays one move",
            "l2": "inputs larger than L2: node table working set is GBs per GPU (no flush needed)",
            "games_per_gpu": games, "sims_per_move": SIMS, "c_puct": C_PUCT, "preroll_move_steps": args.preroll,
            "leaves_per_step": K,
            "search_mode": "exact (visit counts identical to the reference)" if K <= 1
                           else f"virtual loss, {K} leaves per game per step = leaf batch {games * K} (throughput mode)",
            "examples": None if arena else "raw positions -> fixed-capacity slab -> all-gather (side stream) -> D4 augmentation into "
                                           "every rank's device replay ring, inside the timed region",
        },
        "games_per_sec": tot["games_finished"] / (ms / 1e3),
        "mean_game_plies": mean_len,
        "mean_game_plies_note": "positions emitted / games finished inside the timed region (in-run estimate)",
        "plies_per_sec": tot["plies"] / (ms / 1e3),
        "leaf_evals_per_sec": tot["leaf_evals"] / (ms / 1e3),
        "clocks": clocks,
        "e2e": {"value": e2e_sims / (e2e_ms / 1e3), "unit": UNIT,
                "h2d_bytes_per_step": int(sum(b.numel() for b in blobs_host) * 4), "d2h_bytes_per_step": int(d2h / max(1, args.steps)),
                "what": "weights pinned host -> device + device repack every step; " +
                        ("W/L/D tally read back every step" if arena else
                         "the step's D4-augmented examples device -> pinned host (copy stream, double-buffered slabs)")},
        "gpu_launches": int(tot["kernel_launches"]),
        "roofline": {
            "kernel": ("tower conv 128->128 3x3, streamed weights (20 launches per eval, ~99% of net FLOPs)" if tower
                       else "conv3 64->128 3x3 + 1x1 heads (dominant: ~77% of net FLOPs)"),
            "bound": "tensor", "achieved": achieved_tflops, "peak": peak_tf, "unit": "TFLOP/s",
            "frac": achieved_tflops / peak_tf if peak_tf else None,
            "traffic": traffic, "traffic_note": traffic_note,
            "peak_source": f"{peaks['source']} sustained bf16 (kernel timed inside a long step)",
            "kernel_ms_per_launch": kms / klaunches if klaunches else None, "kernel_launches": klaunches,
            "kernel_share_of_step": ((kms / klaunches) * per_move_launches / (ms / args.steps) if klaunches and ms else None),
            "kernel_timing": "one CUDA event pair per launch on the launching stream over the last 40 simulation steps of the "
                             "last timed move step (plain launches; the steps before them replay the captured graph)",
            "net_tensor_frac_end_to_end": (tot["leaf_evals"] / (ms / 1e3)) * flops_per_leaf / 1e12 / world / peak_tf,
            "other_kernels_us_per_launch": {k: 1e3 * v[0] / v[1] for k, v in prof.items() if v[1]},
            "hbm": hbm_rooflines(prof, delta, A, NW, mult, peaks["hbm_gbs"]),
        },
        "collectives": None if arena else {
            "in_timed_loop": ["all_gather of raw position slabs (side stream)", "receiver-side D4 augmentation into the replay ring"]
                             + ([f"broadcast of rank 0's weight blob + device repack every {args.bcast_every} steps"] if world > 1 else []),
            "all_gather_bytes_per_step_per_rank": exch.bytes_per_step, "slab_capacity_positions": raw_cap,
            "weight_blob_bytes": int(blobs_host[0].numel() * 4),
            "value_without": plain_sims / (plain_ms / 1e3) if plain_ms else None,
            "overhead_frac": (plain_sims / (plain_ms / 1e3)) / value - 1.0 if plain_ms and value else None,
        },
        "arena_match": match,
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
    ap.add_argument("--config", type=int, default=3, choices=sorted(CONFIGS), help="BASELINE.json config (default 3, the headline)")
    ap.add_argument("--dtype", default=os.environ.get("AZP_BENCH_DTYPE", "bf16"), choices=["bf16", "f32"])
    ap.add_argument("--net", default=None, choices=["stem", "tower"])
    ap.add_argument("--board", type=int, default=None)
    ap.add_argument("--win", type=int, default=None)
    ap.add_argument("--games", type=int, default=None)
    ap.add_argument("--sims", type=int, default=None)
    ap.add_argument("--leaves-per-step", type=int, default=None,
                    help="1 = exact search (reference-identical visit counts); K > 1 = virtual-loss throughput mode")
    ap.add_argument("--vl-leaves", type=int, default=4, help="K of the virtual-loss side measurement")
    ap.add_argument("--bcast-every", type=int, default=4, help="N > 1: broadcast rank 0's weights every this many steps")
    ap.add_argument("--slab", type=int, default=0,
                    help="raw positions per rank per move step in the all-gathered slab (default max(4096, 2 x games): the "
                         "steady state emits ~games positions per step)")
    ap.add_argument("--no-vl", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the training-step side measurement")
    ap.add_argument("--no-match", action="store_true", help="config 5: skip the whole-match measurement")
    ap.add_argument("--preroll", type=int, default=int(os.environ.get("AZP_BENCH_PREROLL", "64")),
                    help="untimed move steps before warm-up: brings the games to their steady-state age mix")
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_engine_arm(args)


if __name__ == "__main__":
    main()
