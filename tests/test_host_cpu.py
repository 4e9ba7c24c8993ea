"""CPU-side tests (no GPU): the C-ABI library loads and exports every symbol the header
declares, the engine refuses to run without a device, host logic, and the world_size-2
gloo path of the multi-GPU plumbing."""

import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def az():
    import alphazero_piskvorky_b200 as az
    return az


def header_symbols():
    text = open(os.path.join(ROOT, "include", "azp_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(azp_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_header_symbol(az):
    lib = az._lib.load()
    names = header_symbols()
    assert len(names) >= 28
    assert sorted(az._lib.SIGNATURES) == names, "ctypes table and header disagree"
    for n in names:
        assert hasattr(lib, n), n
    out = subprocess.run(["nm", "-D", az._lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (azp_[a-z_0-9]+)", out))
    assert set(names) <= exported
    assert lib.azp_version() >= 100


def _header_text():
    text = open(os.path.join(ROOT, "include", "azp_b200.h")).read()
    return re.sub(r"/\*.*?\*/", "", text, flags=re.S)


def _header_struct(name):
    """[(ctype, field)] of a struct in include/azp_b200.h, in declaration order."""
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), _header_text(), flags=re.S).group(1)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if decl:
            ctype, names = decl.split(None, 1)
            fields += [(ctype, n.strip()) for n in names.split(",")]
    return fields


def test_ctypes_structs_match_the_header(az):
    """ABI drift guard: field order and types of azp_config / azp_counters in the header == the ctypes structures the
    facade passes == the stub INTEGRATION.md tells a maintainer to write."""
    ctype_of = {"int32_t": C.c_int32, "uint64_t": C.c_uint64, "double": C.c_double}
    for cname, struct in (("azp_config", az._lib.Config), ("azp_counters", az._lib.Counters)):
        want = [(n, ctype_of[t]) for t, n in _header_struct(cname)]
        assert [(n, t) for n, t in struct._fields_] == want, cname
    assert C.sizeof(az._lib.Config) == 88 and C.sizeof(az._lib.Counters) == 14 * 8
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    stub = re.search(r"class Config\(C\.Structure\):.*?_fields_ = \[(.*?)\]\n", doc, flags=re.S).group(1)
    stub_fields = re.findall(r'\("([a-z_]+)", C\.(c_[a-z0-9]+)\)', stub)
    assert [(n, getattr(C, t)) for n, t in stub_fields] == list(az._lib.Config._fields_)


def test_ctypes_argument_counts_match_the_header(az):
    """Every prototype in the header has as many parameters as the ctypes argtypes list that binds it."""
    protos = re.findall(r"\b(azp_[a-z_0-9]+)\s*\(([^)]*)\)\s*;", _header_text())
    assert len(protos) == len(az._lib.SIGNATURES)
    for name, params in protos:
        params = params.strip()
        n = 0 if params in ("", "void") else len(params.split(","))
        assert n == len(az._lib.SIGNATURES[name][1]), name


def test_library_is_sm100a_native(az):
    """The shipped cubin targets sm_100a (no PTX-JIT fallback for other archs is relied on)."""
    out = subprocess.run(["cuobjdump", "-lelf", az._lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_engine_fails_loudly_without_gpu(az):
    lib = az._lib.load()
    cfg = az._lib.Config(board=5, win_length=4, games=4, sims=10, c_puct=2.0, dirichlet_alpha=0.3,
                         dirichlet_weight=0.25, net_kind=0, net_dtype=0, tower_blocks=0, tower_channels=0,
                         augment_mode=1, max_plies=0, seed=0, rank=0, world=1, device=0, leaves_per_step=1)
    h = C.c_void_p()
    st = lib.azp_create(C.byref(cfg), C.byref(h))
    assert st == az._lib.ERR_CUDA and not h.value
    assert b"no CPU fallback" in lib.azp_last_error(None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        az.Engine(board=5, win_length=4, games=4, sims=10)
    with pytest.raises(RuntimeError):
        az.Gomoku(5, 4).get_legal_actions()


def test_trainer_fails_loudly_without_gpu(az):
    """The training step has no CPU fallback either: the C ABI reports the missing device, the facade raises."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import ctypes as C
    lib = az._lib.load()
    h = C.c_void_p()
    assert lib.azp_trainer_create(5, 8, 0, C.byref(h)) == az._lib.ERR_CUDA
    assert b"no CPU fallback" in lib.azp_trainer_last_error(None)
    assert lib.azp_trainer_create(2, 8, 0, C.byref(h)) == az._lib.ERR_INVALID_ARG
    ctrl = az.NeuralNetworkController(az.GomokuNet(5), device="cuda")
    with pytest.raises(RuntimeError, match="no CPU"):
        ctrl.train_step(torch.zeros(2, 4, 5, 5), torch.zeros(2, 5, 5), torch.zeros(2), 1e-3)
    with pytest.raises(RuntimeError, match="no CPU"):
        az.Trainer(5, 8)


def test_create_rejects_bad_config(az):
    lib = az._lib.load()
    h = C.c_void_p()
    for kw in (dict(board=2), dict(board=20), dict(win_length=9, board=5), dict(games=0), dict(sims=0),
               dict(world=2, rank=2), dict(net_dtype=7)):
        base = dict(board=5, win_length=4, games=4, sims=10, c_puct=2.0, dirichlet_alpha=0.3, dirichlet_weight=0.25,
                    net_kind=0, net_dtype=0, tower_blocks=0, tower_channels=0, augment_mode=1, max_plies=0, seed=0,
                    rank=0, world=1, device=0, leaves_per_step=1)
        base.update(kw)
        cfg = az._lib.Config(**base)
        assert lib.azp_create(C.byref(cfg), C.byref(h)) == az._lib.ERR_INVALID_ARG, kw


def test_flatten_state_dict_matches_reference_layout(az):
    net = az.GomokuNet(board_size=9)
    sd = net.state_dict()
    assert list(sd.keys()) == [f"{k}.{f}" for k in ("conv1", "conv2", "conv3", "policy_conv", "policy_fc",
                                                     "value_conv", "value_fc1", "value_fc2")
                               for f in ("weight", "bias")]
    blob = az.flatten_state_dict(sd)
    assert blob.dtype == np.float32 and blob.size == sum(v.numel() for v in sd.values()) == 131132
    assert np.array_equal(blob[:32 * 36], sd["conv1.weight"].numpy().reshape(-1))
    tower = az.GomokuTowerNet(board_size=9, channels=32, blocks=2)
    tsd = {k: v for k, v in tower.state_dict().items() if not k.endswith("num_batches_tracked")}
    tblob = az.flatten_state_dict(tower.state_dict(), az.NET_TOWER, 2)
    assert tblob.size == sum(v.numel() for v in tsd.values())
    from oracle import net as onet
    assert set(onet.tower_shapes(9, 32, 2)) == set(tsd)


def test_reference_shapes_and_schedule(az):
    from oracle import net as onet, selfplay
    for b, count in ((5, 100164), (9, 131132), (15, 325964), (19, 662292)):
        assert sum(int(np.prod(s)) for s in onet.stem_shapes(b).values()) == count
    for m in (0, 10, 100, 225):
        assert az.default_temperature_schedule(m) == selfplay.default_temperature_schedule(m)


def test_replay_buffer_host_and_ring_semantics(az):
    from oracle import replay
    buf, ref = az.ReplayBuffer(10), replay.Fifo(10)
    for chunk in (list(range(7)), list(range(100, 108)), [5]):
        buf.extend(chunk); ref.extend(chunk)
    assert buf.all() == ref.all() and len(buf) == len(ref) == 10
    assert list(az.ReplayBuffer(10).sample_batch(5)) == []
    # ring: same eviction order as the deque
    ring, ref = az.ReplayBuffer(10), replay.Fifo(10)
    k = 0
    for n in (4, 5, 3, 12, 1):
        planes = torch.arange(k, k + n, dtype=torch.float32).view(n, 1, 1, 1).expand(n, 4, 2, 2).contiguous()
        pi = torch.arange(k, k + n, dtype=torch.float32).view(n, 1, 1).expand(n, 2, 2).contiguous()
        z = torch.arange(k, k + n, dtype=torch.int32)
        ring.extend_device(planes, pi, z)
        ref.extend(list(range(k, k + n)))
        k += n
        got = ring.device_all()
        assert got[2].tolist() == ref.all()
        assert got[0][:, 0, 0, 0].tolist() == [float(v) for v in ref.all()]
    assert ring.device_len() == 10
    s = ring.sample_device(4)
    assert len(set(s[2].tolist())) == 4


def test_rank_block_partitions_games(az):
    from alphazero_piskvorky_b200 import dist as d
    for total, world in ((16384, 8), (400, 8), (7, 3), (2, 4), (0, 2)):
        blocks = [d.rank_block(100, total, r, world) for r in range(world)]
        assert sum(c for _, c in blocks) == total
        nxt = 100
        for first, c in blocks:
            assert first == nxt
            nxt += c
    assert d.rank_block(0, 16384, 3, 8) == (6144, 2048)           # config 3: 2048 games per GPU


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from alphazero_piskvorky_b200 import dist as d
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    n = 3 + 2 * rank                                      # ragged contributions
    planes = torch.full((n, 4, 3, 3), float(rank))
    pi = torch.full((n, 3, 3), float(rank) + 0.5)
    z = torch.full((n,), rank, dtype=torch.int32)
    gp, gpi, gz = d.all_gather_examples(planes, pi, z)
    blob = torch.arange(10, dtype=torch.float32) if rank == 0 else torch.zeros(10)
    d.broadcast_weights(blob, src=0)
    wld = d.all_reduce_wld(rank + 1, 2, 0, device="cpu")
    grads = d.all_reduce_mean(torch.full((6,), float(rank + 1)))       # data-parallel training: mean of the rank gradients
    q.put((rank, gz.tolist(), gp[:, 0, 0, 0].tolist(), gpi.shape, blob.tolist(), wld, grads.tolist()))
    dist.destroy_process_group()


def test_multi_gpu_plumbing_on_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, gz, gp0, pishape, blob, wld, grads in results:
        assert grads == [1.5] * 6
        assert gz == [0, 0, 0, 1, 1, 1, 1, 1]              # rank order, trimmed to the true counts
        assert gp0 == [0.0] * 3 + [1.0] * 5 and tuple(pishape) == (8, 3, 3)
        assert blob == [float(i) for i in range(10)]
        assert wld == (3, 4, 0)


def _gloo_exchange_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import alphazero_piskvorky_b200 as az
    from alphazero_piskvorky_b200 import dist as d
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    board, cap = 3, 6
    ex = d.ExampleExchange(board, cap, "cpu")
    ring = az.ReplayBuffer(capacity=12)
    log = []
    for step in range(3):
        k = step & 1
        ex.wait_slab(k)
        planes, pi, z = ex.out_views(k)                     # the engine writes its raw positions straight into the slab
        n = (2 + rank + step) % (cap + 1)                   # ragged counts, different on every rank and step
        planes[:n] = 100.0 * step + 10.0 * rank + torch.arange(n, dtype=torch.float32).view(n, 1, 1, 1)
        pi[:n] = 0.5 + planes[:n, 0]
        z[:n] = rank - step
        ex.exchange(k, n, ring=ring)
        log.append(ex.counts(k))
    got = ring.device_all()
    q.put((rank, log, got[0][:, 0, 0, 0].tolist(), got[1][:, 1, 1].tolist(), got[2].tolist(), ex.bytes_per_step))
    dist.destroy_process_group()


def test_example_exchange_slabs_on_gloo_world2():
    """fixed-capacity slabs of RAW positions with the count in the header, all-gathered every move step,
    appended to every rank's ring in rank order -- identical rings everywhere, deque eviction order."""
    import torch.multiprocessing as mp
    from oracle import replay
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_exchange_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref = replay.Fifo(12)
    counts = []
    for step in range(3):
        row = []
        for rank in range(2):
            n = (2 + rank + step) % 7
            row.append(n)
            ref.extend([(100.0 * step + 10.0 * rank + i, rank - step) for i in range(n)])
        counts.append(row)
    for rank, log, p0, pi11, z, nbytes in results:
        assert log == counts
        assert p0 == [v for v, _ in ref.all()] and z == [zz for _, zz in ref.all()]
        assert pi11 == [v + 0.5 for v, _ in ref.all()]
        assert nbytes == 2 * 4 * (4 + 6 * (5 * 9 + 1))


def test_checkpoint_and_buffer_formats(az, tmp_path, golden):
    """reference-format files.  A reference ``state_dict`` (.pt, key names of net.py:37-53)
    loads into the facade net unchanged; ModelLoader picks the newest file or creates a random-init one;
    ``buffer.pkl`` is a pickled deque of (Tensor (4,B,B), ndarray (B,B), int) (replay_buffer.py:50-71)."""
    import pickle
    from collections import deque
    from alphazero_piskvorky_b200.model_loader import ModelLoader
    ck = golden("ref_ckpt_5x5.npz")                                       # tensors of the reference's shipped checkpoint
    path = tmp_path / "models" / "5x5.pt"
    os.makedirs(path.parent)
    torch.save({k: torch.from_numpy(ck[k]) for k in ck.files}, path)
    ctrl = az.NeuralNetworkController(az.GomokuNet(board_size=5), device="cuda")
    v0 = ctrl.version
    ctrl.load(str(path))
    assert ctrl.version == v0 + 1
    assert all(np.array_equal(ctrl.net.state_dict()[k].numpy(), ck[k]) for k in ck.files)
    ctrl.save(str(tmp_path / "out.pt"))
    again = torch.load(tmp_path / "out.pt", weights_only=True)
    assert list(again) == list(ck.files) and all(np.array_equal(again[k].numpy(), ck[k]) for k in ck.files)
    loader = ModelLoader(str(path.parent), board_size=5)
    assert loader.best_path == str(path)
    assert np.array_equal(loader.get_best_model().state_dict()["conv1.weight"].numpy(), ck["conv1.weight"])
    fresh = ModelLoader(str(tmp_path / "empty"), board_size=5)
    assert fresh.best_path is None and fresh.get_best_model().board_size == 5
    assert len(os.listdir(tmp_path / "empty")) == 1                       # the random-init net was saved (model_loader.py:33-37)
    buf = az.ReplayBuffer(capacity=5)
    buf.extend([(torch.full((4, 5, 5), float(i)), np.full((5, 5), i, np.float32), i % 3 - 1) for i in range(7)])
    buf.save(str(tmp_path / "buffer.pkl"))
    with open(tmp_path / "buffer.pkl", "rb") as f:
        raw = pickle.load(f)
    assert isinstance(raw, deque) and len(raw) == 5 and float(raw[0][0][0, 0, 0]) == 2.0
    other = az.ReplayBuffer(capacity=4)
    other.load(str(tmp_path / "buffer.pkl"))
    assert len(other) == 4 and float(other.all()[0][0][0, 0, 0]) == 3.0  # deque(maxlen) keeps the newest


def test_promoter_rule_and_files_with_stub_evaluator(az, tmp_path):
    """promoter.py:27-62 host logic: the candidate is saved as a new reference-format checkpoint iff win_rate > threshold
    (strictly), the baseline is the newest file of the model directory (created random-init when empty)."""
    from alphazero_piskvorky_b200.promoter import ModelPromoter

    class StubEvaluator:
        def __init__(self, rate):
            self.rate, self.calls = rate, []

        def evaluate(self, candidate, baseline, num_games=20, debug=False):
            self.calls.append((candidate, baseline, num_games))
            return self.rate, {"win_rate": self.rate, "total": num_games}

    cand = az.NeuralNetworkController(az.GomokuNet(board_size=5), device="cuda")
    model_dir = str(tmp_path / "models")
    for rate, promoted in ((0.55, False), (0.5501, True), (0.2, False)):
        ev = StubEvaluator(rate)
        before = set(os.listdir(model_dir)) if os.path.isdir(model_dir) else set()
        pr = ModelPromoter(model_dir, ev, az.GomokuNet, device="cuda")
        wr, metrics, was = pr.evaluate_and_maybe_promote(cand, num_games=7, metadata={"episode": 1})
        assert (wr, was) == (rate, promoted) and metrics["total"] == 7
        assert ev.calls[0][1].net.board_size == 5 and ev.calls[0][2] == 7      # baseline = loaded best model, same board
        new = set(os.listdir(model_dir)) - before
        assert len(new) == (1 if not before else 0) + (1 if promoted else 0)    # first call also creates the random-init best
        if promoted:
            sd = torch.load(pr.best_path, weights_only=True)
            assert all(torch.equal(sd[k], v.cpu()) for k, v in cand.net.state_dict().items())
    assert az.should_promote(0.55) is False and az.should_promote(0.56) is True


def test_bench_reference_arm_sampler_runs_on_cpu():
    """bench.py's CPU leg (the oracle port of the reference's per-process inner loop, timed as the baseline): a tiny
    bounded sample on two processes returns a positive simulations/s figure and says what the sample was."""
    sys.path.insert(0, ROOT)
    import bench
    v, cores, sample = bench.cpu_port_throughput(budget_s=1.0, procs=2, board=5, win=4, nsims=20)
    assert v > 0 and cores == 2 and "2 persistent processes" in sample and "5x5/20 sims" in sample
    if bench.reference_available():            # the unmodified reference, staged by __graft_entry__.build()
        rv, rcores, rsample = bench.cpu_throughput("reference", budget_s=1.0, procs=2, board=5, win=4, nsims=20)
        assert rv > 0 and rcores == 2 and "UNMODIFIED reference" in rsample
        assert v > rv                          # the port does not clone a board per legal child


def test_bench_presets_are_the_baseline_configs():
    """`bench.py --config N` must reproduce BASELINE.json's configs[N-1] at its STATED size: board, WIN_LENGTH, simulations
    per move, concurrent games (config 3: 16384 over 8 GPUs = 2048 per GPU, weak scaling), leaf batch (config 4), match size
    (config 5) are parsed out of the config strings and compared with the presets; the default is the headline config."""
    import json
    import re
    sys.path.insert(0, ROOT)
    import bench
    cfgs = json.load(open(os.path.join(ROOT, "BASELINE.json")))["configs"]
    assert len(cfgs) == 5 and sorted(bench.CONFIGS) == [1, 2, 3, 4, 5]
    for i, text in enumerate(cfgs, 1):
        p = bench.CONFIGS[i]
        m = re.search(r"(\d+)\s*[x×]\s*(\d+)", text)
        assert int(m.group(1)) == int(m.group(2)) == p["board"], (i, text)
        w = re.search(r"WIN_LENGTH=(\d+)", text)
        assert p["win"] == (int(w.group(1)) if w else 5), (i, text)
        s = re.search(r"(\d+) sims/move", text)
        if s:
            assert p["sims"] == int(s.group(1)), (i, text)
    assert bench.CONFIGS[1]["sims"] == 200                      # constants.py:NUM_SIMULATIONS of the reference (config 1)
    assert bench.CONFIGS[2]["games"] == 4096
    assert bench.CONFIGS[3]["games"] * 8 == 16384
    assert bench.CONFIGS[4]["net"] == "tower" and bench.CONFIGS[4]["games"] * bench.CONFIGS[4]["leaves"] == 8192
    assert bench.CONFIGS[4]["sims"] == 800
    assert bench.CONFIGS[5]["mode"] == "arena" and bench.CONFIGS[5]["games"] == 400
    assert all(bench.CONFIGS[i]["leaves"] == 1 for i in (1, 2, 3, 5))     # exact mode: reference-identical visit counts
    assert all(c in bench.FLOPS_PER_LEAF for c in {p["board"] for p in bench.CONFIGS.values()})
