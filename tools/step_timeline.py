"""Timeline of the simulation step INSIDE the captured graph (bring-up build: make BRINGUP=1, AZP_TIMELINE=1).

Every kernel of the chain (advance, conv1, conv2, conv3, fc) records per launch the earliest CTA entry, the earliest
moment a CTA got past griddepcontrol.wait and the latest CTA exit (%globaltimer ns).  This prints, over the steady-state
simulation steps of a few move steps at BASELINE config 3 (or --board/--games/--sims), the mean duration of each kernel
(start -> end), the gap to the next kernel's start, how early the next kernel's CTAs became resident, and the period.

usage: AZP_B200_LIB=build/libazp_b200_bringup.so AZP_TIMELINE=1 python tools/step_timeline.py [--moves 3]"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("AZP_TIMELINE", "1")
import numpy as np, torch
import alphazero_piskvorky_b200 as az

ap = argparse.ArgumentParser()
ap.add_argument("--board", type=int, default=15)
ap.add_argument("--win", type=int, default=5)
ap.add_argument("--games", type=int, default=2048)
ap.add_argument("--sims", type=int, default=800)
ap.add_argument("--moves", type=int, default=3)
ap.add_argument("--preroll", type=int, default=12)
ap.add_argument("--dump", default=None)
a = ap.parse_args()

KN = ["advance", "conv1", "conv2", "conv3", "fc"]
RING = 4096
eng = az.Engine(board=a.board, win_length=a.win, games=a.games, sims=a.sims, c_puct=2.0, net_dtype=az.DTYPE_BF16,
                augment=az.AUGMENT_NONE, seed=0)
torch.manual_seed(0)
eng.load_state_dict(az.GomokuNet(a.board).state_dict())
eng.set_slab_capacity(max(4096, 2 * a.games))
cap = eng.example_capacity_floor()
A = a.board * a.board
out = (torch.empty((cap, 4, a.board, a.board), device="cuda"), torch.empty((cap, a.board, a.board), device="cuda"),
       torch.empty((cap,), dtype=torch.int32, device="cuda"))
stream = torch.cuda.Stream()


def read_ring():
    buf = torch.zeros((2 * 1024 * 1024 // 8,), dtype=torch.int64, device="cuda")
    az._lib.check(eng.lib.azp_debug_activations(eng.h, 10, 2, az._lib.ptr(buf), eng._sp()), eng.h)
    torch.cuda.synchronize()
    raw = buf.cpu().numpy().view(np.uint64)
    rec = raw[: 8 * RING * 4].reshape(8, RING, 4)
    done = raw[8 * RING * 4 + 4: 8 * RING * 4 + 8].view(np.uint32)[:8]
    global cta
    cta = raw[8 * RING * 4 + 8: 8 * RING * 4 + 8 + 8 * 320 * 4].reshape(8, 320, 4).astype(np.int64)
    return rec, done


with torch.cuda.stream(stream):
    first = True
    for i in range(a.preroll):
        eng.selfplay(1 << 40, first_game=0, out=out, resume=not first, max_move_steps=1)
        first = False
    torch.cuda.synchronize()
    scratch = torch.zeros(8, dtype=torch.int64, device="cuda")
    az._lib.check(eng.lib.azp_debug_activations(eng.h, 11, 1, az._lib.ptr(scratch), eng._sp()), eng.h)     # clear the ring
    torch.cuda.synchronize()
    for i in range(a.moves):
        eng.selfplay(1 << 40, first_game=0, out=out, resume=True, max_move_steps=1)
    torch.cuda.synchronize()
    rec, done = read_ring()

inv = np.uint64(0xFFFFFFFFFFFFFFFF)
events = []
for k, name in enumerate(KN):
    used = np.nonzero(rec[k, :, 2])[0]          # the ring was cleared before the window: every non-zero slot is one launch
    if len(used) == 0 or len(used) >= RING:
        print(name, "launches in the window:", len(used), "(ring too short or kernel not stamped)")
        continue
    for j in used:
        r = rec[k, j]
        events.append((int(inv - r[1]), int(r[2]), int(inv - r[0]), k))
events.sort()
t0 = events[0][0]
if a.dump:
    with open(a.dump, "w") as f:
        for s, e, en, k in events[:400]:
            f.write(f"{KN[k]:8s} entry {en - t0:10d} start {s - t0:10d} end {e - t0:10d} dur {e - s:7d}\n")
dur = {k: [] for k in range(5)}
gap = {k: [] for k in range(5)}      # this kernel's end -> next kernel's start
early = {k: [] for k in range(5)}    # next kernel's earliest CTA entry relative to this kernel's end (negative = resident early)
period = []
last_adv = None
for i, (s, e, en, k) in enumerate(events):
    if e - s > 2_000_000:             # move boundary / first launch
        continue
    dur[k].append(e - s)
    if i + 1 < len(events):
        s2, e2, en2, k2 = events[i + 1]
        if k2 == (k + 1) % 5 and s2 - e < 200_000:
            gap[k].append(s2 - e)
            early[k].append(en2 - e)
    if k == 0:
        if last_adv is not None and s - last_adv < 1_000_000:
            period.append(s - last_adv)
        last_adv = s
print(f"{'kernel':8s} {'launches':>8s} {'start->end us':>14s} {'end->next start us':>19s} {'next entry - end us':>20s}")
tot = 0.0
for k, name in enumerate(KN):
    if not dur[k]:
        continue
    d, g_, e_ = np.median(dur[k]) / 1e3, np.median(gap[k]) / 1e3 if gap[k] else float("nan"), np.median(early[k]) / 1e3 if early[k] else float("nan")
    tot += d + (g_ if g_ == g_ else 0)
    print(f"{name:8s} {len(dur[k]):8d} {d:14.2f} {g_:19.2f} {e_:20.2f}")
print(f"sum of medians {tot:.2f} us; median period (advance start -> next advance start) {np.median(period) / 1e3:.2f} us over {len(period)} steps")

# per-CTA spread of the LAST launch of every kernel (start = past griddepcontrol.wait)
for k, name in enumerate(KN):
    c = cta[k]
    used = c[:, 1] > 0
    if not used.any():
        continue
    st, en = c[used, 0], c[used, 1]
    d = (en - st) / 1e3
    e0 = (en - en.min()) / 1e3
    print(f"{name:8s} CTAs {used.sum():4d}  duration us: min {d.min():6.2f} median {np.median(d):6.2f} max {d.max():6.2f}   "
          f"end spread us: median {np.median(e0):5.2f} p90 {np.percentile(e0, 90):5.2f} max {e0.max():5.2f}")
    if name == "conv2" and a.dump:
        with open(a.dump + ".conv2_ctas", "w") as f:
            for i in np.nonzero(used)[0]:
                f.write(f"{i} {(c[i,0]-st.min())/1e3:.2f} {(c[i,1]-st.min())/1e3:.2f} sm {c[i,2]} mean_dur {c[i,3]/1e3/max(1,len(dur[k])):.2f}\n")
