#!/usr/bin/env python3
"""Generate the golden fixtures in this directory by RUNNING THE UNMODIFIED
REFERENCE (imported from /root/reference/alphazero).  Only runs in the build
container; the fixtures it writes are committed so the tests never need the
reference tree.

    python tests/golden/make_golden.py

Nothing from the reference is copied: fixtures hold inputs (seeds, move lists,
boards) and the reference's OUTPUTS on them.  The one exception is
``ref_ckpt_5x5.npz``: the float32 tensors of the reference's shipped checkpoint
``alphazero/models/saved/5x5_4_in_a_row.pt`` (data, not source), used as golden
weights.
"""

from __future__ import annotations

import contextlib
import io
import os
import queue
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/alphazero"
OUT = os.environ.get("AZP_GOLDEN_OUT", HERE)     # tests regenerate into a scratch dir and compare
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)
sys.path.insert(2, os.path.join(ROOT, "tests"))

with contextlib.redirect_stdout(io.StringIO()):
    import games as ref_games            # noqa: E402
    import mcts as ref_mcts              # noqa: E402
    import net as ref_net                # noqa: E402
    import controller as ref_controller  # noqa: E402
    import self_play as ref_self_play    # noqa: E402
    import evaluator as ref_evaluator    # noqa: E402
    import replay_buffer as ref_replay   # noqa: E402

import synth                             # noqa: E402
from oracle import net as onet           # noqa: E402  (only for the synthetic weight generator)
from oracle import train as otrain       # noqa: E402  (only for the synthetic batch generator)

CODE = {None: 0, "X": 1, "O": 2, "D": 3}
SIZES = ((5, 4), (9, 5), (15, 5), (19, 5))


def quiet():
    return contextlib.redirect_stdout(io.StringIO())


def cells(game) -> np.ndarray:
    return np.array([[CODE[c] for c in row] for row in game.board], dtype=np.int8)


def game_from_cells(board: np.ndarray, win: int, to_move: int = 1, last=None):
    g = ref_games.Gomoku(int(board.shape[0]), win)
    sym = {0: None, 1: "X", 2: "O"}
    g.board = [[sym[int(v)] for v in row] for row in board]
    g.current_player = sym[to_move]
    g.last_action = last
    return g


# --------------------------------------------------------------------------------
def make_rules():
    """Legal mask / apply / terminal / winner / encode on random playouts and on
    arbitrary (unreachable) boards, per board size."""
    for size, win in SIZES:
        rs = np.random.RandomState(1000 + size)
        n_play, n_fuzz = 160, 240
        max_len = size * size
        moves = np.full((n_play, max_len, 2), -1, dtype=np.int16)
        lens = np.zeros(n_play, dtype=np.int32)
        boards = np.zeros((n_play, size, size), dtype=np.int8)
        to_move = np.zeros(n_play, dtype=np.int8)
        legal = np.zeros((n_play, size, size), dtype=np.uint8)
        result = np.zeros(n_play, dtype=np.int8)
        planes = np.zeros((n_play, 4, size, size), dtype=np.uint8)
        term_ply = np.zeros(n_play, dtype=np.int32)        # first ply at which is_terminal() held
        for i in range(n_play):
            g = ref_games.Gomoku(size, win)
            # a third of the playouts run to the end of the game, the rest stop early
            target = max_len if i % 3 == 0 else rs.randint(0, max_len // 2 + 1)
            k = 0
            while k < target and not g.is_terminal():
                acts = g.get_legal_actions()
                a = acts[rs.randint(len(acts))]
                moves[i, k] = a
                g = g.apply_action(a)
                k += 1
            lens[i] = k
            term_ply[i] = k if g.is_terminal() else -1
            boards[i] = cells(g)
            to_move[i] = CODE[g.current_player]
            for r, c in g.get_legal_actions():
                legal[i, r, c] = 1
            result[i] = CODE[g.get_game_result()]
            planes[i] = g.encode("cpu").numpy().astype(np.uint8)

        fuzz = np.zeros((n_fuzz, size, size), dtype=np.int8)
        fuzz_result = np.zeros(n_fuzz, dtype=np.int8)
        fuzz_legal_count = np.zeros(n_fuzz, dtype=np.int32)
        for i in range(n_fuzz):
            density = rs.uniform(0.2, 1.0)
            if i % 8 == 0:
                density = 1.0                                # full boards: draw vs win
            b = (rs.random_sample((size, size)) < density) * rs.randint(1, 3, size=(size, size))
            if i % 8 == 0 and i % 16 == 0:
                # break up lines so that some full boards are genuine draws
                b = ((np.add.outer(np.arange(size), 2 * np.arange(size)) // 2) % 2 + 1)
                b = np.where(rs.random_sample((size, size)) < 0.03, 3 - b, b)
            fuzz[i] = b
            g = game_from_cells(fuzz[i], win)
            fuzz_result[i] = CODE[g.get_game_result()]
            fuzz_legal_count[i] = len(g.get_legal_actions())
        np.savez_compressed(os.path.join(OUT, f"rules_{size}x{size}_w{win}.npz"),
                            size=size, win=win, moves=moves, lens=lens, boards=boards,
                            to_move=to_move, legal=legal, result=result, planes=planes,
                            term_ply=term_ply, fuzz=fuzz, fuzz_result=fuzz_result,
                            fuzz_legal_count=fuzz_legal_count)
        print(f"rules {size}x{size}: results {np.bincount(result, minlength=4)}, "
              f"fuzz {np.bincount(fuzz_result, minlength=4)}")

    # occupied-cell error behaviour (games.py:76-77)
    g = ref_games.Gomoku(5, 4).apply_action((2, 2))
    try:
        g.apply_action((2, 2))
        raise SystemExit("reference accepted an occupied cell?")
    except ValueError as e:
        assert str(e) == "Invalid move"


# --------------------------------------------------------------------------------
def random_positions(size, win, count, seed, max_frac=0.5):
    rs = np.random.RandomState(seed)
    out = []
    while len(out) < count:
        g = ref_games.Gomoku(size, win)
        for _ in range(rs.randint(0, int(size * size * max_frac) + 1)):
            if g.is_terminal():
                break
            acts = g.get_legal_actions()
            g = g.apply_action(acts[rs.randint(len(acts))])
        if not g.is_terminal():
            out.append(g)
    return out


def make_net():
    """GomokuNet.forward and policy_value_fn on synthetic weights (all sizes) and
    on the shipped 5x5 checkpoint."""
    sd = torch.load(os.path.join(REF, "models/saved/5x5_4_in_a_row.pt"), weights_only=True)
    np.savez_compressed(os.path.join(OUT, "ref_ckpt_5x5.npz"),
                        **{k: v.numpy().astype(np.float32) for k, v in sd.items()})

    out = {}
    for size, win in SIZES:
        params = onet.synthetic_stem_params(size, seed=size)
        model = ref_net.GomokuNet(board_size=size)
        model.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()})
        model.eval()
        pos = random_positions(size, win, 12, seed=2000 + size)
        x = torch.stack([g.encode("cpu") for g in pos])
        with torch.no_grad():
            logits, value = model(x)
        with quiet():
            ctrl = ref_controller.NeuralNetworkController(model, device="cpu")
        fn = ref_controller.make_policy_value_fn(ctrl)
        pol = np.stack([fn(g)[0] for g in pos])
        val1 = np.array([fn(g)[1] for g in pos], dtype=np.float64)
        out[f"x_{size}"] = x.numpy().astype(np.uint8)
        out[f"logits_{size}"] = logits.numpy()
        out[f"value_{size}"] = value.numpy()
        out[f"policy_{size}"] = pol
        out[f"value1_{size}"] = val1
    # shipped checkpoint
    model = ref_net.GomokuNet(board_size=5)
    model.load_state_dict(sd)
    model.eval()
    pos = random_positions(5, 4, 16, seed=77)
    x = torch.stack([g.encode("cpu") for g in pos])
    with torch.no_grad():
        logits, value = model(x)
    out["x_ckpt"] = x.numpy().astype(np.uint8)
    out["logits_ckpt"] = logits.numpy()
    out["value_ckpt"] = value.numpy()
    np.savez_compressed(os.path.join(OUT, "net_stem.npz"), **out)
    print("net: wrote", sorted(out))


# --------------------------------------------------------------------------------
class _RootTap:
    """Make the UNMODIFIED ``MCTS.run`` reveal its root: ``mcts.Node`` is
    replaced by a subclass that remembers every instance; the first one created
    in a run is the root.  The reference's own code computes everything."""

    def __enter__(self):
        tap = self
        tap.nodes = []
        base = ref_mcts.Node
        self.base = base

        class Tapped(base):
            def __init__(self, *a, **kw):
                super().__init__(*a, **kw)
                tap.nodes.append(self)

        ref_mcts.Node = Tapped
        return self

    def __exit__(self, *exc):
        ref_mcts.Node = self.base

    def root_visits(self, size):
        root = self.nodes[0]
        v = np.zeros((size, size), dtype=np.int64)
        for (r, c), child in root.children.items():
            v[r, c] = child.N
        return v, root.N

    def depth_stats(self):
        expanded = sum(1 for n in self.nodes if n.children)
        return expanded


def near_win_position(size, win, seed):
    """A position where both sides have open (win-1)-lines, so searches hit terminal leaves."""
    rs = np.random.RandomState(seed)
    g = ref_games.Gomoku(size, win)
    r0 = rs.randint(0, size - 1)
    c0 = rs.randint(0, size - win + 1)
    r1 = (r0 + 1 + rs.randint(0, size - 1)) % size
    if r1 == r0:
        r1 = (r0 + 1) % size
    for i in range(win - 1):
        g = g.apply_action((r0, c0 + i))          # X
        g = g.apply_action((r1, c0 + i))          # O
    assert not g.is_terminal()
    return g


def make_mcts():
    out = {}
    cases = []
    plan = {5: (200, 8), 9: (400, 5), 15: (300, 4), 19: (200, 3)}
    for size, win in SIZES:
        sims, count = plan[size]
        pos = random_positions(size, win, count, seed=3000 + size, max_frac=0.3)
        pos[0] = ref_games.Gomoku(size, win)                  # the empty board
        pos.append(near_win_position(size, win, seed=size))
        for j, g in enumerate(pos):
            for sharp in (3.0, 6.0):
                T = ref_self_play.default_temperature_schedule(j * 7) if sharp == 3.0 else 0.25
                fn = synth.reference_evaluator(salt=size, sharp=sharp)
                np.random.seed(4000 + 10 * size + j)
                with _RootTap() as tap:
                    m = ref_mcts.MCTS(fn, num_simulations=sims, c_puct=2.0)
                    pi, action = m.run(g, temperature=T, add_root_noise=False)
                    visits, root_n = tap.root_visits(size)
                assert root_n == sims and visits.sum() == sims
                tag = f"{size}_{j}_{int(sharp)}"
                out[f"board_{tag}"] = cells(g)
                out[f"meta_{tag}"] = np.array([size, win, CODE[g.current_player], sims,
                                               -1 if g.last_action is None else g.last_action[0],
                                               -1 if g.last_action is None else g.last_action[1]],
                                              dtype=np.int64)
                out[f"temp_{tag}"] = np.array([float(T), 1.0 if isinstance(T, np.floating) else 0.0])
                out[f"visits_{tag}"] = visits
                out[f"pi_{tag}"] = np.asarray(pi, dtype=np.float32)
                out[f"action_{tag}"] = np.array(action, dtype=np.int64)
                out[f"seed_{tag}"] = np.array(4000 + 10 * size + j)
                cases.append(tag)
    out["cases"] = np.array(cases)

    # root noise on: seeded global NumPy RNG (mcts.py:114,177) -- sample-level only
    # reproducible with the same numpy version; the test says so.
    ncases = []
    for size, win in ((5, 4), (9, 5)):
        g = random_positions(size, win, 1, seed=3500 + size, max_frac=0.2)[0]
        fn = synth.reference_evaluator(salt=size + 100, sharp=3.0)
        np.random.seed(777 + size)
        with _RootTap() as tap:
            m = ref_mcts.MCTS(fn, num_simulations=120, c_puct=2.0)
            pi, action = m.run(g, temperature=ref_self_play.default_temperature_schedule(3),
                               add_root_noise=True)
            visits, _ = tap.root_visits(size)
            priors = np.zeros((size, size), dtype=np.float64)
            for (r, c), ch in tap.nodes[0].children.items():
                priors[r, c] = ch.prior
        tag = f"noise_{size}"
        out[f"board_{tag}"] = cells(g)
        out[f"meta_{tag}"] = np.array([size, win, CODE[g.current_player], 120, 777 + size], dtype=np.int64)
        out[f"visits_{tag}"] = visits
        out[f"pi_{tag}"] = np.asarray(pi, dtype=np.float32)
        out[f"action_{tag}"] = np.array(action, dtype=np.int64)
        out[f"priors_{tag}"] = priors
        ncases.append(tag)
    out["noise_cases"] = np.array(ncases)
    out["numpy_version"] = np.array(np.__version__)

    # edge cases: terminal root (won board), full board
    g = ref_games.Gomoku(5, 4)
    for a in ((0, 0), (1, 0), (0, 1), (1, 1), (0, 2), (1, 2), (0, 3)):
        g = g.apply_action(a)
    assert g.is_terminal()
    np.random.seed(5)
    with _RootTap() as tap:
        pi, action = ref_mcts.MCTS(synth.reference_evaluator(1), 20, 2.0).run(g, 1.0)
        visits, root_n = tap.root_visits(5)
    out["edge_won_board"] = cells(g)
    out["edge_won_pi"] = np.asarray(pi, dtype=np.float32)
    out["edge_won_visits"] = visits
    out["edge_won_rootn"] = np.array(root_n)
    full = ((np.add.outer(np.arange(3), 2 * np.arange(3)) // 2) % 2 + 1).astype(np.int8)
    g = game_from_cells(np.array([[1, 2, 1], [1, 2, 2], [2, 1, 1]], dtype=np.int8), 3)
    pi, action = ref_mcts.MCTS(synth.reference_evaluator(1), 5, 2.0).run(g, 1.0)
    assert action is None
    out["edge_full_pi"] = np.asarray(pi, dtype=np.float32)
    np.savez_compressed(os.path.join(OUT, "mcts.npz"), **out)
    print(f"mcts: {len(cases)} deterministic cases, {len(ncases)} noise cases")


# --------------------------------------------------------------------------------
def make_mcts_ckpt():
    """Search driven by the shipped 5x5 checkpoint; the evaluator outputs seen
    during each search are recorded so the test replays them bit-for-bit
    independent of the CPU's conv kernels."""
    sd = torch.load(os.path.join(REF, "models/saved/5x5_4_in_a_row.pt"), weights_only=True)
    model = ref_net.GomokuNet(board_size=5)
    model.load_state_dict(sd)
    model.eval()
    with quiet():
        ctrl = ref_controller.NeuralNetworkController(model, device="cpu")
    base_fn = ref_controller.make_policy_value_fn(ctrl)
    out = {}
    # roots: positions along one greedy game
    g = ref_games.Gomoku(5, 4)
    roots = []
    np.random.seed(11)
    while not g.is_terminal():
        roots.append(g)
        _, a = ref_mcts.MCTS(base_fn, 100, 2.0).run(g, temperature=0.0)
        g = g.apply_action(a)
    roots = roots[:2] + roots[-4:]          # two openings + the four positions before the end
    term_total = 0
    for j, g in enumerate(roots):
        rec_b, rec_tm, rec_p, rec_v, rec_l = [], [], [], [], []

        def fn(game, _b=rec_b, _tm=rec_tm, _p=rec_p, _v=rec_v, _l=rec_l):
            p, v = base_fn(game)
            _b.append(cells(game)); _tm.append(CODE[game.current_player])
            _p.append(p.copy()); _v.append(v)
            _l.append((-1, -1) if game.last_action is None else tuple(game.last_action))
            return p, v

        np.random.seed(100 + j)
        with _RootTap() as tap:
            pi, action = ref_mcts.MCTS(fn, 200, 2.0).run(g, temperature=1.0)
            visits, _ = tap.root_visits(5)
        term_total += 201 - len(rec_v)
        out[f"board_{j}"] = cells(g)
        out[f"tomove_{j}"] = np.array(CODE[g.current_player])
        out[f"visits_{j}"] = visits
        out[f"pi_{j}"] = np.asarray(pi, dtype=np.float32)
        out[f"rec_boards_{j}"] = np.stack(rec_b)
        out[f"rec_tomove_{j}"] = np.array(rec_tm, dtype=np.int8)
        out[f"rec_last_{j}"] = np.array(rec_l, dtype=np.int8)
        out[f"last_{j}"] = np.array((-1, -1) if g.last_action is None else g.last_action, dtype=np.int8)
        out[f"rec_policy_{j}"] = np.stack(rec_p).astype(np.float32)
        out[f"rec_value_{j}"] = np.array(rec_v, dtype=np.float64)
    out["count"] = np.array(len(roots))
    np.savez_compressed(os.path.join(OUT, "mcts_ckpt5.npz"), **out)
    print(f"mcts_ckpt5: {len(roots)} roots, {term_total} terminal-leaf simulations")


# --------------------------------------------------------------------------------
def make_selfplay():
    out = {}
    # temperature schedule (self_play.py:24-26)
    plies = np.array([0, 1, 10, 25, 100, 225, 360])
    out["sched_plies"] = plies
    out["sched_T"] = np.array([ref_self_play.default_temperature_schedule(int(p)) for p in plies])
    out["eval_T"] = np.array([ref_evaluator.temperature_schedule(int(p)) for p in range(12)])

    # augmentation as written (self_play.py:94-108)
    mgr = ref_self_play.SelfPlayManager(controller=None, device="cpu")
    rs = np.random.RandomState(9)
    for size in (3, 5, 15):
        st = torch.from_numpy(rs.random_sample((4, size, size)).astype(np.float32))
        pi = rs.random_sample((size, size)).astype(np.float32)
        sym = mgr._augment_symmetries(st, pi)
        out[f"aug_state_in_{size}"] = st.numpy()
        out[f"aug_pi_in_{size}"] = pi
        out[f"aug_state_out_{size}"] = np.stack([s.numpy() for s, _ in sym])
        out[f"aug_pi_out_{size}"] = np.stack([p for _, p in sym])

    # the reference's own episode loop (_worker, self_play.py:29-77) run in-process
    # on plain queues, with the evaluator factory swapped for the synthetic one.
    orig = ref_self_play.make_policy_value_fn
    ref_self_play.make_policy_value_fn = lambda controller: synth.reference_evaluator(salt=55, sharp=3.0)
    try:
        tq, rq = queue.Queue(), queue.Queue()
        for i in range(3):
            tq.put(i)
        sd = ref_net.GomokuNet().state_dict()
        np.random.seed(2024)
        with quiet():
            ref_self_play._worker(tq, rq, sd, {"num_simulations": 60, "c_puct": 2.0},
                                  ref_self_play.default_temperature_schedule)
    finally:
        ref_self_play.make_policy_value_fn = orig
    games = []
    while not rq.empty():
        games.append(rq.get())
    assert len(games) == 3
    for gi, data in enumerate(games):
        out[f"ep_states_{gi}"] = np.stack([s.numpy() for s, _, _ in data]).astype(np.uint8)
        out[f"ep_pi_{gi}"] = np.stack([np.asarray(p, dtype=np.float32) for _, p, _ in data])
        out[f"ep_z_{gi}"] = np.array([z for _, _, z in data], dtype=np.int8)
    out["ep_count"] = np.array(len(games))
    out["ep_seed"] = np.array(2024)
    out["numpy_version"] = np.array(np.__version__)
    np.savez_compressed(os.path.join(OUT, "selfplay.npz"), **out)
    print("selfplay: episodes of length", [len(g) for g in games])


def make_arena():
    """ModelEvaluator.evaluate (evaluator.py:38-122) with synthetic evaluators; the
    temperature steps it uses are recorded through its own schedule function."""
    out = {}
    steps = []
    orig_fn, orig_T, orig_sims = (ref_evaluator.make_policy_value_fn,
                                  ref_evaluator.temperature_schedule,
                                  ref_evaluator.NUM_EVAL_SIMULATIONS)

    class Ctl:                                        # stands in for a controller
        def __init__(self, salt):
            self.salt = salt
            self.net = self

        def eval(self):
            pass

    ref_evaluator.make_policy_value_fn = lambda c: synth.reference_evaluator(salt=c.salt, sharp=3.0)

    def tapped_T(step):
        steps.append(step)
        return orig_T(step)

    ref_evaluator.temperature_schedule = tapped_T
    ref_evaluator.NUM_EVAL_SIMULATIONS = 40
    try:
        np.random.seed(31337)
        ev = ref_evaluator.ModelEvaluator(ref_games.Gomoku, print_games=False, device="cpu")
        with quiet():
            win_rate, metrics = ev.evaluate(Ctl(1), Ctl(2), num_games=6)
    finally:
        (ref_evaluator.make_policy_value_fn, ref_evaluator.temperature_schedule,
         ref_evaluator.NUM_EVAL_SIMULATIONS) = orig_fn, orig_T, orig_sims
    out["wld"] = np.array([metrics["wins"], metrics["losses"], metrics["draws"]])
    out["win_rate"] = np.array(win_rate)
    out["steps"] = np.array(steps)
    out["seed"] = np.array(31337)
    out["sims"] = np.array(40)
    np.savez_compressed(os.path.join(OUT, "arena.npz"), **out)
    print("arena:", metrics, "temperature steps:", len(steps))


def make_replay():
    buf = ref_replay.ReplayBuffer(capacity=10)
    buf.extend(list(range(7)))
    buf.extend(list(range(100, 108)))
    out = {"final": np.array(buf.all()), "len": np.array(len(buf))}
    small = ref_replay.ReplayBuffer(capacity=10)
    small.extend([1, 2, 3])
    out["undersized_sample"] = np.array(list(small.sample_batch(5)))
    np.savez_compressed(os.path.join(OUT, "replay.npz"), **out)


TRAIN_CASES = ((5, 24, 3, 1e-3), (9, 16, 2, 5e-5))     # board, batch, steps, lr
TRAIN_STRIDE = 4


def train_sample(a: np.ndarray) -> np.ndarray:
    """Fixture size control: big tensors are stored as every TRAIN_STRIDE-th element."""
    a = np.asarray(a, dtype=np.float32).reshape(-1)
    return a if a.size <= 4096 else a[::TRAIN_STRIDE].copy()


def make_train():
    """NeuralNetworkController.train_step (controller.py:82-131) of the UNMODIFIED reference on
    seeded weights and batches: total loss per step, gradients left in ``param.grad`` by the first
    step, parameters after the last step (AdamW, weight_decay 1e-4)."""
    out = {}
    for size, n, steps, lr in TRAIN_CASES:
        params = onet.synthetic_stem_params(size, seed=100 + size)
        model = ref_net.GomokuNet(board_size=size)
        model.load_state_dict({k: torch.from_numpy(v) for k, v in params.items()})
        with quiet():
            ctrl = ref_controller.NeuralNetworkController(model, device="cpu")
        planes, pi, z = otrain.synthetic_batch(size, n, seed=300 + size)
        losses = []
        for step in range(steps):
            with quiet():
                loss, _ = ctrl.train_step(torch.from_numpy(planes), torch.from_numpy(pi), torch.from_numpy(z), lr)
            losses.append(loss)
            if step == 0:
                for k, p in ctrl.net.named_parameters():
                    out[f"grad1_{size}_{k}"] = train_sample(p.grad.detach().numpy())
        for k, p in ctrl.net.named_parameters():
            out[f"final_{size}_{k}"] = train_sample(p.detach().numpy())
        out[f"planes_{size}"] = planes.astype(np.uint8)
        out[f"pi_{size}"] = pi
        out[f"z_{size}"] = z
        out[f"losses_{size}"] = np.array(losses, dtype=np.float64)
        out[f"lr_{size}"] = np.array(lr)
    np.savez_compressed(os.path.join(OUT, "train.npz"), **out)
    print("train: losses", {s: out[f"losses_{s}"].tolist() for s, *_ in TRAIN_CASES})


if __name__ == "__main__":
    torch.set_num_threads(1)
    only = set(sys.argv[1:])
    for name, fn in (("rules", make_rules), ("net", make_net), ("mcts", make_mcts),
                     ("mcts_ckpt", make_mcts_ckpt), ("selfplay", make_selfplay),
                     ("arena", make_arena), ("replay", make_replay), ("train", make_train)):
        if not only or name in only:
            fn()
    sizes = {f: os.path.getsize(os.path.join(HERE, f)) for f in sorted(os.listdir(HERE)) if f.endswith(".npz")}
    print(sizes, "total", sum(sizes.values()))
