"""GPU parity tests of the DRIVER-level semantics: what the episode loops of the reference do between searches.

* self-play (alphazero/self_play.py:47-58): the pi a worker records at ply p is ``MCTS.run(state, T(p))`` -- the root
  visit counts of THAT search pushed through mcts.py:143-170 at the temperature of THAT ply;
* arena (alphazero/evaluator.py:64-90): X (candidate) moves first in even games, O in odd games; the side to move
  searches with ITS OWN net, without root noise, at ``T = table[step]`` where ``step`` is that side's own counter
  (0,1,2,.. for the first mover, 1,2,3,.. for the second); the tally counts X wins as candidate wins;
* the evaluator hook keeps the value a host callable returns as a Python float (float64), mcts.py:137-141.

The drivers run in slices of ONE lockstep move (``max_move_steps=1``); between slices the tests read the visit counts
of the searches that have just ended (azp_root_visits) and the episode bookkeeping (azp_episode_snapshot), and replay
everything with the oracle."""

import os
import numpy as np
import pytest
import torch

import synth
from oracle import arena as oarena, gomoku, mcts as omcts, net as onet, selfplay as oselfplay

pytestmark = pytest.mark.gpu

IDLE, PLAYING, PENDING, DELAYED = 0, 1, 2, 3


@pytest.fixture(scope="module")
def az():
    import alphazero_piskvorky_b200 as az
    return az


def _pi_reference(visits_row, legal, T):
    """mcts.py:143-170 on the root's children in row-major order; T arrives as np.float64 from the schedules."""
    idx = np.flatnonzero(legal)
    return idx, omcts.policy_from_visits(visits_row[idx].astype(np.float32), np.float64(T))


class _Tracker:
    """Host-side mirror of one slot: the oracle position of the game the slot is playing."""

    def __init__(self, size, win):
        self.size, self.win = size, win
        self.gid, self.state, self.moves, self.pis = None, None, [], []

    def seat(self, gid, first=gomoku.X):
        self.gid = int(gid)
        self.state = gomoku.State(self.size, self.win)
        self.state.to_move = first
        self.moves, self.pis = [], []


@pytest.mark.parametrize("size,win,sims,slots,ngames,steps,dtype", [
    (15, 8, 12, 6, 6, 130, "bf16"),      # WIN 8: nobody wins early, so plies 0, 1, 10, 100 are all reached
    (5, 4, 40, 8, 20, 100, "f32"),        # short games: slots are re-seated many times, every game ends inside the run
])
def test_selfplay_pi_is_the_searchs_visit_counts_at_the_plys_temperature(az, size, win, sims, slots, ngames, steps, dtype):
    A = size * size
    eng = az.Engine(board=size, win_length=win, games=slots, sims=sims, augment=az.AUGMENT_NONE, seed=13,
                    net_dtype=az.DTYPE_BF16 if dtype == "bf16" else az.DTYPE_F32)
    eng.load_state_dict(onet.synthetic_stem_params(size, seed=5, scale=2.0))
    # the reference's own schedule (self_play.py:24-26): every ply has a different temperature
    table = np.array([oselfplay.default_temperature_schedule(m) for m in range(A + 1)], dtype=np.float64)
    eng.set_temperature_table(table)
    trackers = [_Tracker(size, win) for _ in range(slots)]
    finished = {}                         # tuple(moves) -> [pi rows] of every finished game
    emitted = []
    plies_checked = set()
    out, status_before = None, None
    for step in range(steps):
        st, planes, pi, z, written = eng.selfplay(ngames, max_move_steps=1, out=out, resume=step > 0)
        out = (planes, pi, z)
        if written:
            emitted.append((planes[:written].cpu().numpy(), pi[:written].cpu().numpy(), z[:written].cpu().numpy()))
        visits = eng.root_visits(slots).cpu().numpy()
        snap = {k: v.cpu().numpy() for k, v in eng.episode_snapshot().items()}
        for g, tr in enumerate(trackers):
            if step == 0:
                # every slot got a game (ngames >= slots) and has played exactly one move of it
                assert snap["ply"][g] == 1 and snap["status"][g] == PLAYING
                tr.seat(snap["game_id"][g])
            elif status_before[g] != PLAYING:                      # parked slot: no search happened here
                if snap["status"][g] == PLAYING:                   # (a game waiting for slab space got flushed)
                    assert snap["ply"][g] == 0
                    tr.seat(snap["game_id"][g])
                continue
            p = len(tr.moves)                                      # the ply this step's search was run at
            board_flat = tr.state.board.reshape(-1)
            legal = board_flat == 0
            row_pi = snap["pi"][g, p]
            move = int(snap["moves"][g, p].astype(np.uint16))
            assert visits[g].sum() == sims and not visits[g][~legal].any()          # root.N == S (mcts.py:123-141)
            idx, ref = _pi_reference(visits[g], legal, table[p])                     # T(ply), self_play.py:53
            assert np.abs(row_pi[idx] - ref).max() <= 1e-6, (step, g, p)            # a12 tolerance
            assert not row_pi[~legal].any()
            assert legal[move] and row_pi[move] > 0                                 # a draw from pi (mcts.py:177)
            plies_checked.add(p)
            tr.pis.append(row_pi.copy())
            tr.moves.append(move)
            tr.state = tr.state.play((move // size, move % size))
            if tr.state.is_terminal():
                finished[tuple(tr.moves)] = (tr.pis, tr.state.result())
                # the slot was flushed and handed its next game (or parked) inside the same driver call
                assert snap["game_id"][g] != tr.gid or snap["status"][g] != PLAYING
                if snap["status"][g] == PLAYING:
                    assert snap["ply"][g] == 0
                    tr.seat(snap["game_id"][g])
            else:
                assert snap["game_id"][g] == tr.gid and snap["ply"][g] == p + 1 and snap["status"][g] == PLAYING
        status_before = snap["status"].copy()
        if st != az.MORE:
            break
    if size == 15:
        assert {0, 1, 10, 100} <= plies_checked
    else:
        assert st != az.MORE and len(finished) >= ngames - 2       # (two games sharing every move would collide)
    # the examples the writer emitted are exactly the recorded rows: game-major, pi bit-identical, z from the result
    n_rows = 0
    for planes, pi, z in emitted:
        i = 0
        while i < len(z):
            assert planes[i].sum() == 0                            # a game starts on the empty board
            j = i + 1
            while j < len(z) and planes[j].sum() != 0:
                j += 1
            moves = []
            for k in range(i, j - 1):                              # move k = the last-move plane of position k+1
                moves.append(int(np.flatnonzero(planes[k + 1][2].reshape(-1))[0]))
            match = [m for m in finished if len(m) == j - i and list(m[:-1]) == moves]
            assert match, "emitted game not among the tracked finished games"
            pis, result = finished[match[0]]
            for k in range(i, j):
                assert np.array_equal(pi[k].reshape(-1), pis[k - i])
                player = gomoku.X if (k - i) % 2 == 0 else gomoku.O
                assert z[k] == oselfplay.label(result, player)     # self_play.py:71
            n_rows += j - i
            i = j
    assert n_rows == sum(len(m) for m in finished) or size == 15   # 15x15: games still running when the test stops


@pytest.mark.parametrize("table_kind", ["reference", "index_revealing"])
def test_arena_first_mover_step_counters_nets_and_tally_match_the_oracle(az, table_kind):
    size, win, sims, slots, ngames = 5, 4, 30, 6, 14
    A = size * size
    params = [onet.synthetic_stem_params(size, seed=1, scale=2.0), onet.synthetic_stem_params(size, seed=2, scale=2.0)]
    eng = az.Engine(board=size, win_length=win, games=slots, sims=sims, c_puct=1.5, seed=5)
    eng.load_state_dict(params[0], slot=0)              # candidate, plays X
    eng.load_state_dict(params[1], slot=1)              # baseline, plays O
    if table_kind == "reference":
        table = np.array([oarena.eval_temperature(s) for s in range(A + 1)], dtype=np.float64)   # evaluator.py:14-19
    else:
        table = np.array([0.5 + 0.13 * s for s in range(A + 1)], dtype=np.float64)     # every index gives another pi
    eng.set_temperature_table(table)

    def evaluator(slot):
        cache = {}

        def fn(state):
            key = (state.board.tobytes(), state.to_move, state.last)
            if key not in cache:
                _, pol, val = eng.net_forward(torch.from_numpy(state.planes()).unsqueeze(0).cuda(), slot=slot)
                cache[key] = (pol[0].cpu().numpy().reshape(size, size), float(val[0].item()))
            p, v = cache[key]
            return p.copy(), v
        return fn

    evals = [evaluator(0), evaluator(1)]
    trackers = [_Tracker(size, win) for _ in range(slots)]
    tally = {gomoku.X: 0, gomoku.O: 0, gomoku.DRAW: 0}
    seen_first = set()
    status_before, done, step, distinguishing = None, 0, 0, 0
    movers = {}                              # move step -> sides that searched in it
    while True:
        st, wld = eng.arena_steps(ngames, max_move_steps=1, resume=step > 0)
        visits = eng.root_visits(slots).cpu().numpy()
        snap = {k: v.cpu().numpy() for k, v in eng.episode_snapshot().items()}
        for g, tr in enumerate(trackers):
            seated_now = snap["status"][g] in (PLAYING, DELAYED)
            if tr.state is None or tr.state.is_terminal():             # no live game tracked in this slot
                if not seated_now:
                    continue                                           # parked: the match has no more games
                gid = int(snap["game_id"][g])
                tr.seat(gid, gomoku.X if gid % 2 == 0 else gomoku.O)   # evaluator.py:64-69
            if seated_now and snap["game_id"][g] == tr.gid:
                # the tracked game is still in the slot: it either searched and moved in this step, or -- phase-aligned
                # arena -- it is waiting for the step whose phase matches its first mover (only ever before its first move)
                if snap["ply"][g] == len(tr.moves):
                    assert len(tr.moves) == 0
                    continue
                assert snap["ply"][g] == len(tr.moves) + 1
            # (otherwise the tracked game made its LAST move in this step and the slot already holds its successor)
            p = len(tr.moves)
            side = tr.state.to_move
            first = gomoku.X if tr.gid % 2 == 0 else gomoku.O
            assert side == (first if p % 2 == 0 else gomoku.other(first))
            seen_first.add(first)
            # the side to move searches with its own net (evaluator.py:71-84): candidate = X = slot 0
            ref = omcts.Search(evals[0 if side == gomoku.X else 1], sims, 1.5).run(
                tr.state, 1.0, add_root_noise=False, forced_action=tr.state.legal_actions()[0])
            assert np.array_equal(visits[g].reshape(size, size), ref.visits), (step, g, p)
            # ... at the temperature of ITS OWN step counter (evaluator.py:87-90)
            tidx = oarena.step_counters(first == gomoku.X, p)
            assert tidx == (p // 2 if p % 2 == 0 else p // 2 + 1)
            legal = tr.state.board.reshape(-1) == 0
            idx, want = _pi_reference(visits[g], legal, table[tidx])
            row_pi = snap["pi"][g, p]
            assert np.abs(row_pi[idx] - want).max() <= 1e-6, (step, g, p, tidx)
            # power of the check: how often would a neighbouring index (the ply itself, the other side's counter) have
            # given a measurably different pi?  (matching `want` to 1e-6 then rules that index out)
            for wrong in {p, tidx - 1, tidx + 1} - {tidx}:
                if 0 <= wrong <= A and np.abs(_pi_reference(visits[g], legal, table[wrong])[1] - want).max() > 1e-4:
                    distinguishing += 1
            move = int(snap["moves"][g, p].astype(np.uint16))
            assert legal[move] and row_pi[move] > 0
            tr.moves.append(move)
            tr.state = tr.state.play((move // size, move % size))
            if tr.state.is_terminal():
                tally[tr.state.result()] += 1
                done += 1
                if snap["status"][g] in (PLAYING, DELAYED):                       # the slot's next game is seated already
                    assert snap["ply"][g] == 0 and snap["game_id"][g] != tr.gid
                    gid = int(snap["game_id"][g])
                    tr.seat(gid, gomoku.X if gid % 2 == 0 else gomoku.O)
                    # all games of a move step are evaluated by one net: X moves in even steps, O in odd ones
                    assert (snap["status"][g] == PLAYING) == ((tr.state.to_move == gomoku.X) == ((step + 1) % 2 == 0))
            else:
                assert snap["game_id"][g] == tr.gid and snap["ply"][g] == p + 1
            movers.setdefault(step, set()).add(side)
        status_before = snap["status"].copy()
        step += 1
        assert wld == (tally[gomoku.X], tally[gomoku.O], tally[gomoku.DRAW])       # candidate wins == X wins
        if st != az.MORE:
            break
        assert step < 40 * ngames
    assert done == ngames and seen_first == {gomoku.X, gomoku.O}
    assert all(m == {gomoku.X if k % 2 == 0 else gomoku.O} for k, m in movers.items())     # one net per move step
    assert distinguishing >= (100 if table_kind == "index_revealing" else 20)      # a wrong index would have been seen
    win_rate = (wld[0] + 0.5 * wld[2]) / ngames                                    # evaluator.py:103-105
    assert oarena.should_promote(win_rate) == az.should_promote(win_rate)
    eng.close()


def test_arena_run_equals_its_sliced_form(az):
    """azp_arena_run == azp_arena_run_steps run to the end: same Philox streams, same tally."""
    size, win = 5, 4
    out = []
    for sliced in (False, True):
        eng = az.Engine(board=size, win_length=win, games=8, sims=24, seed=9)
        eng.load_state_dict(onet.synthetic_stem_params(size, seed=1, scale=2.0), slot=0)
        eng.load_state_dict(onet.synthetic_stem_params(size, seed=2, scale=2.0), slot=1)
        eng.set_temperature_table([oarena.eval_temperature(s) for s in range(26)])
        if not sliced:
            out.append(eng.arena(20))
        else:
            st, wld, k = az.MORE, None, 0
            while st == az.MORE:
                st, wld = eng.arena_steps(20, max_move_steps=3, resume=k > 0)
                k += 1
            out.append(wld)
        eng.close()
    assert out[0] == out[1] and sum(out[0]) == 20


def test_hook_backs_up_the_evaluators_python_float_unrounded(az):
    """mcts.py:137-141: ``value`` is a Python float (float64) and is backed up as such.  With uniform priors every PUCT
    comparison is decided by Q, so rounding the evaluator's value to float32 changes visit counts: the float64 entry
    point must reproduce the oracle fed with the doubles, the float32 one the oracle fed with the rounded values."""
    from alphazero_piskvorky_b200 import engine as _e
    size, win, sims, n = 5, 4, 300, 12
    A = size * size

    def make_eval(rounded):
        def fn(state):
            key = synth.board_key(state.board, state.to_move, 3)
            v = ((key % 13) - 6) * 0.05                                    # decimals: sums tie in float64, not after float32 rounding
            if rounded:
                v = float(np.float32(v))
            return np.full((size, size), 1.0 / A, dtype=np.float32), v
        return fn

    states = []
    for i in range(n):
        s = gomoku.State(size, win)
        for a in synth.random_playout(size, win, plies=i % 6, seed=300 + i):
            s = s.play(a)
        states.append(s)
    batch = az.GomokuBatch.from_cells(np.stack([s.board for s in states]), to_move=[s.to_move for s in states],
                                      last=[-1 if s.last is None else s.last[0] * size + s.last[1] for s in states],
                                      win_length=win)
    eng = az.Engine(board=size, win_length=win, games=n, sims=sims, c_puct=2.0)
    results = {}
    for rounded, dt in ((False, torch.float64), (True, torch.float32)):
        fn = make_eval(rounded)
        eng.search_begin(batch.x, batch.o, batch.to_move, batch.last)
        pol = torch.full((n, A), 1.0 / A, device="cuda")
        eng.root_expand(pol)
        for _ in range(sims):
            eng.select()
            x, o, tm, last, kind = eng.leaf_states()
            cells = _e.unpack_boards(x, o, size).view(n, size, size).cpu().numpy()
            tm_h, last_h, kind_h = tm.cpu().numpy(), last.cpu().numpy(), kind.cpu().numpy()
            val = np.zeros(n, dtype=np.float64)
            for i in np.flatnonzero(kind_h == 0):
                la = None if last_h[i] < 0 else (int(last_h[i]) // size, int(last_h[i]) % size)
                val[i] = fn(gomoku.State.from_board(cells[i], int(tm_h[i]), win, la))[1]
            eng.expand_backup(pol, torch.from_numpy(val).to("cuda", dt))
        got = eng.root_visits().cpu().numpy()
        for i, s in enumerate(states):
            ref = omcts.Search(fn, sims, 2.0).run(s, 1.0, forced_action=s.legal_actions()[0])
            assert np.array_equal(got[i].reshape(size, size), ref.visits), (rounded, i)
        results[rounded] = got
    assert not np.array_equal(results[False], results[True])       # the test is sensitive to the rounding
    # and the facade's hook path takes the float64 route
    sym = {0: None, 1: "X", 2: "O"}
    gs = []
    for s in states:
        g = az.Gomoku(size, win)
        g.board = [[sym[int(v)] for v in row] for row in s.board]
        g.current_player = sym[s.to_move]
        g.last_action = s.last
        gs.append(g)

    def ref_fn(game):
        b, tm = synth._cells_of_reference(game)
        return make_eval(False)(gomoku.State.from_board(b, tm, win))
    _, _, visits = az.MCTS(ref_fn, sims, 2.0).run_batch(gs, 1.0)
    assert np.array_equal(visits, results[False])
    eng.close()


# ---- replay ring writer, raw-slab exchange, per-kernel profiling ---------------------------------------------------------
def test_ring_writer_kernel_has_deque_semantics_and_augments(az):
    """azp_augment_write_ring == ReplayBuffer.extend (replay_buffer.py:27-29: append in order, evict the oldest) fused with
    the augmentation; the position count may live on the device (slab header)."""
    from alphazero_piskvorky_b200 import engine as e
    from oracle import replay
    size, A = 5, 25
    rs = np.random.RandomState(4)
    ring, ref = az.ReplayBuffer(capacity=40), replay.Fifo(40)
    k = 0
    for n in (7, 9, 30, 55, 3):                                   # 55 > capacity: only the newest 40 survive
        planes = torch.from_numpy(rs.random_sample((n, 4, size, size)).astype(np.float32)).cuda()
        planes[:, 0, 0, 0] = torch.arange(k, k + n, dtype=torch.float32)
        pi = torch.from_numpy(rs.random_sample((n, size, size)).astype(np.float32)).cuda()
        z = torch.arange(k, k + n, dtype=torch.int32).cuda()
        ring.extend_device(planes, pi, z)                         # CUDA tensors: the kernel path
        ref.extend(list(range(k, k + n)))
        k += n
        got = ring.device_all()
        assert got[2].cpu().tolist() == ref.all()
        assert got[0][:, 0, 0, 0].cpu().tolist() == [float(v) for v in ref.all()]
    # device-side count + D4 augmentation on the way in: equals azp_augment_write of the first `cnt` positions
    cap, cnt = 12, 5
    planes = torch.from_numpy((rs.random_sample((cap, 4, size, size)) < 0.3).astype(np.float32)).cuda()
    pi = torch.from_numpy(rs.random_sample((cap, size, size)).astype(np.float32)).cuda()
    z = torch.from_numpy(rs.randint(-1, 2, size=cap).astype(np.int32)).cuda()
    n_dev = torch.tensor([cnt], dtype=torch.int32).cuda()
    ring2 = az.ReplayBuffer(capacity=64)
    ring2.write_slab(planes, pi, z, n_max=cap, n_dev=n_dev, augment=az.AUGMENT_D4)
    assert ring2.sync_head() == cnt * 8
    op, opi, oz = e.augment_write(planes[:cnt], pi[:cnt], z[:cnt], size, az.AUGMENT_D4)
    got = ring2.device_all()
    assert torch.equal(got[0], op) and torch.equal(got[1], opi) and torch.equal(got[2], oz)


def test_example_exchange_on_one_gpu_matches_engine_side_augmentation(az):
    """The multi-GPU example path at world size 1: raw positions -> slab -> (all-gather) -> D4 augmentation into the ring
    on the receiving side == the engine's own D4 writer on the same games (same Philox streams)."""
    from alphazero_piskvorky_b200 import dist as azd
    size, win, sims, slots, ngames = 5, 4, 24, 8, 12
    params = onet.synthetic_stem_params(size, seed=3, scale=2.0)
    out = {}
    for mode in ("exchange", "engine"):
        eng = az.Engine(board=size, win_length=win, games=slots, sims=sims, seed=17,
                        augment=az.AUGMENT_NONE if mode == "exchange" else az.AUGMENT_D4)
        eng.load_state_dict(params)
        if mode == "exchange":
            ex = azd.ExampleExchange(size, eng.example_capacity_floor(), torch.device("cuda", 0))
            ring = az.ReplayBuffer(capacity=8 * 25 * ngames)
            step, st = 0, az.MORE
            with torch.cuda.stream(torch.cuda.Stream()):
                while st == az.MORE:
                    k = step & 1
                    ex.wait_slab(k)
                    st, _, _, _, w = eng.selfplay(ngames, max_move_steps=1, out=ex.out_views(k), resume=step > 0)
                    ex.exchange(k, w, ring=ring, augment=az.AUGMENT_D4)
                    step += 1
                ex.wait()
            torch.cuda.synchronize()
            assert ring.sync_head() == eng.counters()["positions"] * 8
            out[mode] = tuple(t.cpu() for t in ring.device_all())
        else:
            out[mode] = tuple(t.cpu() for t in eng.selfplay_all(ngames))
        eng.close()
    # games that finish in the same move step reach the slab in the order of an atomic counter: compare as multisets
    def keyed(t):
        P, Q, Z = (x.numpy() for x in t)
        return sorted((P[i].tobytes(), Q[i].tobytes(), int(Z[i])) for i in range(len(Z)))
    assert keyed(out["exchange"]) == keyed(out["engine"]) and len(out["engine"][2]) > 0


def test_per_kernel_profiling_brackets_every_step_kernel(az):
    """azp_get_profile_kernel: with profiling on, the last simulation steps of a move run as plain launches, each kernel
    between its own CUDA events; units = leaves (net kernels) / game slots (tree kernels)."""
    size, win, sims, n = 9, 5, 64, 32
    eng = az.Engine(board=size, win_length=win, games=n, sims=sims, net_dtype=az.DTYPE_BF16, augment=az.AUGMENT_NONE)
    eng.load_state_dict(onet.synthetic_stem_params(size, seed=1, scale=2.0))
    with torch.cuda.stream(torch.cuda.Stream()):
        ref = eng.selfplay(n, max_move_steps=2)[4]
        eng.set_profiling(True)
        eng.selfplay(n, max_move_steps=1, resume=True)
        prof = eng.get_profile_kernels()
        eng.set_profiling(False)
    for name in ("dominant", "advance", "conv1", "conv2", "fc"):
        ms, launches, units = prof[name]
        # AZP_PROF_TAIL = 40 plain steps (advance: 39 + root).  (32 games: too few for the descent kernels' conv1 tail, which
        # would leave the root evaluation's as the only conv1 LAUNCH -- test_descent_conv1_tail_matches_the_conv1_kernel)
        assert launches >= 39 and ms > 0, name
        assert units <= launches * n and units > 0
    assert prof["end_move"][1] == 1 and prof["root_expand"][1] == 1 and prof["end_move"][2] == n
    assert all(v == (0.0, 0, 0.0) for v in eng.get_profile_kernels().values())      # read-and-reset
    eng.close()


def test_selfplay_manager_ring_path_and_live_weights(az):
    """Facade level: (i) SelfPlayManager.generate_to_ring -- the multi-GPU example path at world size 1 -- fills a device
    ring with 8 examples per position; (ii) engines cached by MCTS / make_policy_value_fn notice every way the
    controller's weights can change (the reference evaluates controller.net live, controller.py:39-53): in-place
    load_state_dict, a torch optimizer step, a re-assigned net."""
    net = az.GomokuNet(board_size=5)
    ctrl = az.NeuralNetworkController(net, device="cuda")
    mgr = az.SelfPlayManager(ctrl, device="cuda", mcts_params={"num_simulations": 16, "c_puct": 2.0}, concurrent_games=8)
    ring = az.ReplayBuffer(capacity=4096)
    n = mgr.generate_to_ring(10, ring, slab_positions=64)
    assert n > 0 and n % 8 == 0 and ring.device_len() == n
    planes, pi, z = ring.device_all()
    assert planes.shape[1:] == (4, 5, 5) and float(pi.sum(dim=(1, 2)).sub(1).abs().max()) < 1e-4
    assert set(z.cpu().tolist()) <= {-1, 0, 1}

    fn = az.make_policy_value_fn(ctrl)
    state = az.Gomoku(5, 4)
    p0, v0 = fn(state)
    torch.manual_seed(7)
    other = az.GomokuNet(board_size=5)
    ctrl.net.load_state_dict(other.state_dict())                 # as the reference's worker does (self_play.py:36)
    p1, v1 = fn(state)
    assert not np.allclose(p0, p1)
    with torch.no_grad():
        ref_logits, ref_v = other(state.encode().unsqueeze(0).cpu())
    assert np.abs(p1.reshape(-1) - torch.softmax(ref_logits[0], 0).numpy()).max() <= 1e-5 and abs(v1 - float(ref_v)) <= 1e-5
    opt = torch.optim.SGD(ctrl.net.parameters(), lr=0.5)         # a torch optimizer on net.parameters()
    for q in ctrl.net.parameters():
        q.grad = torch.ones_like(q)
    opt.step()
    p2, _ = fn(state)
    assert not np.allclose(p1, p2)
    torch.manual_seed(9)
    ctrl.net = az.GomokuNet(board_size=5)                        # re-assigned module
    p3, _ = fn(state)
    assert not np.allclose(p2, p3)
    m = az.MCTS(fn, 24, 2.0, seed=1)
    _, _, va = m.run_batch([state], 1.0)
    ctrl.net.load_state_dict(other.state_dict())
    _, _, vb = m.run_batch([state], 1.0)
    assert va.sum() == vb.sum() == 24 and not np.array_equal(va, vb)          # the cached search engine reloaded too
    m.close()


@pytest.mark.parametrize("size,win,games", [(9, 5, 0), (15, 5, 0), (9, 5, 400), (15, 5, 900), (19, 5, 96), (6, 4, 200)])
def test_descent_conv1_tail_matches_the_conv1_kernel(az, size, win, games):
    """Exact mode on the bf16 stem with enough games to fill the machine: the descent kernels run conv1 of the leaves they emit as
    their own tail (descend_conv1_kernel, azp_conv1_tail.cuh) instead of a conv1_taps_kernel launch.  Same operands, same UMMA
    sequence per row -> bit-identical activations -> the SAME self-play: visit counts of every root, moves, pi and examples equal
    those of an engine created with AZP_FUSE_DESCENT=0, and the per-kernel profile shows which path ran.  Board sizes with one,
    two and four 128-row tiles per board, pitch 16 (6x6) and B+1."""
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    n, sims = games or 8 * sms + 16, 12          # 4, 8 or 16 games per CTA: the fewest that seats every game in one CTA per SM
    outs = {}
    for fused in ("1", "0"):
        os.environ["AZP_FUSE_DESCENT"] = fused
        os.environ["AZP_FUSE_DESCENT_MIN_GAMES"] = "64"      # by default the tail needs half a machine-full of games (8 per SM)
        try:
            eng = az.Engine(board=size, win_length=win, games=n, sims=sims, net_dtype=az.DTYPE_BF16, augment=az.AUGMENT_NONE, seed=3)
        finally:
            os.environ.pop("AZP_FUSE_DESCENT", None)
            os.environ.pop("AZP_FUSE_DESCENT_MIN_GAMES", None)
        eng.load_state_dict(onet.synthetic_stem_params(size, seed=2, scale=2.0))
        with torch.cuda.stream(torch.cuda.Stream()):
            eng.set_profiling(True)
            st, planes, pi, z, w = eng.selfplay(n, max_move_steps=3)
            visits = eng.root_visits(n).cpu().numpy().copy()
            prof = eng.get_profile_kernels()
            eng.set_profiling(False)
            snap = eng.episode_snapshot()
        # slots grab their game ids with an atomic counter: the slot -> game mapping differs from run to run, the games do not
        order = np.argsort(snap["game_id"].cpu().numpy(), kind="stable")
        assert np.array_equal(np.sort(snap["game_id"].cpu().numpy()), np.arange(n))
        outs[fused] = (visits[order], [t.cpu().numpy()[order].copy() for t in (snap["moves"], snap["pi"], snap["ply"])],
                       prof["conv1"][1], prof["advance"][1])
        eng.close()
    (v1, s1, c1_fused, adv_fused), (v0, s0, c1_plain, adv_plain) = outs["1"], outs["0"]
    assert adv_fused >= adv_plain >= sims - 1         # the fused path also brackets the descent of step 0
    assert c1_plain > adv_plain and c1_fused < c1_plain and c1_fused <= 3        # fused: only the root evaluations launch conv1
    assert v1.sum() > 0 and np.array_equal(v1, v0)
    for a, b in zip(s1, s0):
        assert np.array_equal(a, b)
