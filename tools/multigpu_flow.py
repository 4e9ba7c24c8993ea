"""Multi-GPU flow of SURVEY section 8(e) under torchrun (one process per GPU, NCCL over NVLink):
   1. games sharded by global id, each rank self-plays its block (no collective in the simulation loop);
   2. ragged all-gather of the generated example slabs into every rank's replay buffer;
   3. arena new-vs-best sharded the same way, W/L/D all-reduced, promotion rule;
   4. broadcast of the promoted weights as one flat fp32 blob and reload on every rank;
   5. data-parallel training steps: every rank takes its shard of one global batch drawn from the
      gathered replay ring, gradients are all-reduced (NCCL) as ONE flat blob, AdamW runs on every rank; rank 0 also
      trains on the whole batch alone -- the weights must agree.
Checks shard invariance against a single-rank run of the same global game ids on rank 0."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import alphazero_piskvorky_b200 as az
from alphazero_piskvorky_b200 import dist as azd

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
B, WIN, SIMS, GAMES = 9, 5, 64, 64
torch.manual_seed(0); best = az.GomokuNet(B)
torch.manual_seed(1); new = az.GomokuNet(B)

eng = az.Engine(board=B, win_length=WIN, games=GAMES // world, sims=SIMS, net_dtype=az.DTYPE_BF16,
                augment=az.AUGMENT_NONE, seed=3, rank=rank, world=world, device=local)
eng.load_state_dict(best.state_dict())
t0 = time.time()
planes, pi, z = eng.selfplay_all(GAMES)                       # this rank's block of the 64 global games
gp, gpi, gz = azd.all_gather_examples(planes, pi, z)          # every rank now holds all positions
buf = az.ReplayBuffer(capacity=40_000)
buf.extend_device(gp, gpi, gz)
t_selfplay = time.time() - t0

ok_invariance = None
if rank == 0:                                                 # same global ids on ONE rank: identical multiset
    solo = az.Engine(board=B, win_length=WIN, games=GAMES, sims=SIMS, net_dtype=az.DTYPE_BF16,
                     augment=az.AUGMENT_NONE, seed=3, rank=0, world=1, device=local)
    solo.load_state_dict(best.state_dict())
    sp, spi, sz = solo.selfplay_all(GAMES)
    key = lambda P, Q, Z: sorted((P[i].cpu().numpy().tobytes(), Q[i].cpu().numpy().tobytes(), int(Z[i])) for i in range(len(Z)))
    ok_invariance = key(sp, spi, sz) == key(gp, gpi, gz)
    solo.close()

# arena: candidate (slot 0, X) vs best (slot 1, O), 32 games sharded over the ranks
arena = az.Engine(board=B, win_length=WIN, games=max(1, 32 // world), sims=SIMS, net_dtype=az.DTYPE_BF16, seed=5,
                  rank=rank, world=world, device=local)
arena.load_state_dict(new.state_dict(), slot=0)
arena.load_state_dict(best.state_dict(), slot=1)
arena.set_temperature_table([0.3 * np.exp(-s / 4) for s in range(B * B + 1)])
w, l, d = arena.arena(32)
W, L, D = azd.all_reduce_wld(w, l, d)
win_rate = (W + 0.5 * D) / (W + L + D)
promote = az.should_promote(win_rate)

# broadcast the winner's weights from rank 0 and reload them on the device everywhere
winner = new if promote else best
blob = torch.from_numpy(az.flatten_state_dict(winner.state_dict())).cuda() if rank == 0 else \
    torch.zeros(eng.weight_count(), device="cuda")
azd.broadcast_weights(blob, src=0)
eng.load_blob(blob)
x = torch.zeros((1, 4, B, B), device="cuda"); x[0, 0, 4, 4] = 1
_, _, v = eng.net_forward(x)
vs = [torch.zeros_like(v) for _ in range(world)]
dist.all_gather(vs, v)
same_value = all(torch.equal(vs[0], q) for q in vs)
# ---- data-parallel training on the gathered examples ----
torch.manual_seed(2)
GB, STEPS, LR = 256, 3, 1e-3
P, PI, Z = buf.device_all()
sel = torch.randperm(P.shape[0], generator=torch.Generator().manual_seed(7))[:GB].cuda()     # same global batch on every rank
gP, gPI, gZ = P[sel], PI[sel], Z[sel].float()
lo, n_loc = azd.rank_block(0, GB, rank, world)
init = {k: v.clone() for k, v in best.state_dict().items()}
net_dp = az.GomokuNet(B); net_dp.load_state_dict(init)
ctrl = az.NeuralNetworkController(net_dp, device=f"cuda:{local}", batch_size=GB, rank=rank, world=world)
t1 = time.time()
dp_losses = [ctrl.train_step(gP[lo:lo + n_loc], gPI[lo:lo + n_loc], gZ[lo:lo + n_loc], LR)[0] for _ in range(STEPS)]
torch.cuda.synchronize()
t_train = time.time() - t1
flat = ctrl.flat_parameters()
flats = [torch.zeros_like(flat) for _ in range(world)]
dist.all_gather(flats, flat)
dp_identical = all(torch.equal(flats[0], f) for f in flats)
dp_vs_single = None
if rank == 0:
    net_1 = az.GomokuNet(B); net_1.load_state_dict(init)
    solo_ctrl = az.NeuralNetworkController(net_1, device=f"cuda:{local}", batch_size=GB)
    for _ in range(STEPS):
        solo_ctrl.train_step(gP, gPI, gZ, LR)
    d = (solo_ctrl.flat_parameters() - flat).abs()
    dp_vs_single = {"quantile_999": float(torch.quantile(d[::7], 0.999)), "max": float(d.max())}
if rank == 0:
    print(json.dumps({"world": world, "dp_train": {"global_batch": GB, "steps": STEPS, "local_losses": dp_losses,
                      "weights_identical_across_ranks": bool(dp_identical), "abs_diff_vs_single_rank_full_batch": dp_vs_single,
                      "seconds": round(t_train, 3)}, "positions_total": int(gz.numel()), "positions_this_rank": int(z.numel()),
                      "replay_len": buf.device_len(), "shard_invariant": ok_invariance, "arena_wld": [W, L, D],
                      "win_rate": win_rate, "promote": bool(promote), "weights_identical_after_broadcast": bool(same_value),
                      "selfplay_seconds": round(t_selfplay, 2)}))
dist.destroy_process_group()
