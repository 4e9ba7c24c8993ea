// Engine object, C ABI and the lockstep drivers (search / self-play / arena).
#include "azp_net.cuh"
#include <atomic>
#include <mutex>
#include <string.h>
#include <stdlib.h>

// ---- process-wide bookkeeping ------------------------------------------------------------------------
static std::atomic<uint64_t> g_launches{0};
static std::mutex g_err_mutex;
static std::string g_error;

void azp_count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }
uint64_t azp_launch_count() { return g_launches.load(std::memory_order_relaxed); }
void azp_set_global_error(const std::string& msg) { std::lock_guard<std::mutex> l(g_err_mutex); g_error = msg; }
const char* azp_global_error() { return g_error.c_str(); }
int azp_check_launch() {
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) {
    azp_set_global_error(std::string("kernel launch: ") + cudaGetErrorString(err));
    return AZP_ERR_CUDA;
  }
  return AZP_OK;
}

bool azp_pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char* q = getenv("AZP_PDL"); v = q ? (atoi(q) != 0) : 1; }
  return v != 0;
}

#if defined(AZP_BRINGUP) || defined(AZP_TIMELINE_BUILD)
__device__ AzpTimeline* g_azp_tl = nullptr;
static AzpTimeline* g_tl_host_ptr = nullptr;
// AZP_TIMELINE=1 in a bring-up build: allocate the ring once per process and publish it to the kernels
static void azp_timeline_setup() {
  if (g_tl_host_ptr || !getenv("AZP_TIMELINE")) return;
  if (cudaMalloc(&g_tl_host_ptr, sizeof(AzpTimeline)) != cudaSuccess) { g_tl_host_ptr = nullptr; return; }
  cudaMemset(g_tl_host_ptr, 0, sizeof(AzpTimeline));
  cudaMemcpyToSymbol(g_azp_tl, &g_tl_host_ptr, sizeof(g_tl_host_ptr));
}
int azp_timeline_read(void* out_dev, size_t bytes, cudaStream_t s) {
  if (g_tl_host_ptr && !out_dev) return cudaMemsetAsync(g_tl_host_ptr, 0, sizeof(AzpTimeline), s) == cudaSuccess ? AZP_OK : AZP_ERR_CUDA;   // reset
  if (!g_tl_host_ptr || bytes < sizeof(AzpTimeline)) return AZP_ERR_UNSUPPORTED;
  if (cudaMemcpyAsync(out_dev, g_tl_host_ptr, sizeof(AzpTimeline), cudaMemcpyDeviceToDevice, s) != cudaSuccess) return AZP_ERR_CUDA;
  return AZP_OK;
}
#else
static void azp_timeline_setup() {}
int azp_timeline_read(void*, size_t, cudaStream_t) { return AZP_ERR_UNSUPPORTED; }
#endif

RulesParams azp_make_rules(int B, int WL) {
  RulesParams rp;
  memset(&rp, 0, sizeof(rp));
  rp.B = B; rp.A = B * B; rp.NW = AZP_WORDS(B); rp.WL = WL;
  rp.dshift[0] = 1; rp.dshift[1] = B; rp.dshift[2] = B + 1; rp.dshift[3] = B - 1;
  const int dr[4] = {0, 1, 1, 1}, dc[4] = {1, 0, 1, -1};
  for (int r = 0; r < B; r++)
    for (int c = 0; c < B; c++) {
      int i = r * B + c;
      rp.board_mask[i >> 6] |= 1ull << (i & 63);
      for (int d = 0; d < 4; d++) {
        int er = r + (WL - 1) * dr[d], ec = c + (WL - 1) * dc[d];
        if (er >= 0 && er < B && ec >= 0 && ec < B) rp.valid[d][i >> 6] |= 1ull << (i & 63);
      }
    }
  return rp;
}

extern "C" int azp_version(void) { return 100; }

extern "C" const char* azp_last_error(const azp_engine* e) {
  if (e && !e->error.empty()) return e->error.c_str();
  return azp_global_error();
}

// ---- create / destroy ------------------------------------------------------------------------------------
static int create_impl(azp_engine* e) {
  const azp_config& c = e->cfg;
  const size_t G = c.games, A = e->A, NW = e->NW;
  const size_t GK = (size_t)c.games * (c.leaves_per_step > 1 ? c.leaves_per_step : 1);
  TreeView& tv = e->tv;
  e->K = c.leaves_per_step > 1 ? c.leaves_per_step : 1;
  e->slots = c.games * e->K;
  tv.G = c.games; tv.A = e->A; tv.NW = e->NW; tv.S = c.sims; tv.K = e->K; tv.slots = e->slots;
  tv.cap = c.sims + e->K + 2;
  const int empties = e->A;
  tv.maxd = (c.sims + 1 < empties ? c.sims + 1 : empties) + 1;
  int st;
  const size_t edges = G * (size_t)tv.cap * A;
  if ((st = e->alloc(&tv.N, edges))) return st;
  if ((st = e->alloc(&tv.W, edges))) return st;
  if ((st = e->alloc(&tv.P, edges))) return st;
  if ((st = e->alloc(&tv.child, edges))) return st;
  if ((st = e->alloc(&tv.node_count, G))) return st;
  if ((st = e->alloc(&tv.root_n, G))) return st;
  if ((st = e->alloc(&tv.root_x, G * NW))) return st;
  if ((st = e->alloc(&tv.root_o, G * NW))) return st;
  if ((st = e->alloc(&tv.root_tm, G))) return st;
  if ((st = e->alloc(&tv.root_last, G))) return st;
  if ((st = e->alloc(&tv.ply, G))) return st;
  if ((st = e->alloc(&tv.game_id, G))) return st;
  if ((st = e->alloc(&tv.root_result, G))) return st;
  if ((st = e->alloc(&tv.active, G))) return st;
  if ((st = e->alloc(&tv.net_of_game, G))) return st;
  if ((st = e->alloc(&tv.path, GK * (size_t)tv.maxd))) return st;
  if ((st = e->alloc(&tv.path_len, GK))) return st;
  if ((st = e->alloc(&tv.leaf_kind, GK))) return st;
  if ((st = e->alloc(&tv.leaf_tvalue, GK))) return st;
  if ((st = e->alloc(&tv.leaf_slot, GK))) return st;
  if ((st = e->alloc(&tv.slot_game, GK))) return st;
  if ((st = e->alloc(&tv.leaf_x, GK * NW))) return st;
  if ((st = e->alloc(&tv.leaf_o, GK * NW))) return st;
  if ((st = e->alloc(&tv.leaf_tm, GK))) return st;
  if ((st = e->alloc(&tv.leaf_last, GK))) return st;
  if ((st = e->alloc(&tv.n_leaves, 4))) return st;
  if ((st = e->alloc(&tv.error_flag, 1))) return st;
  if ((st = e->alloc(&tv.counters, (size_t)CNT_COUNT))) return st;
  if ((st = e->alloc(&e->d_policy, GK * A))) return st;
  if ((st = e->alloc(&e->d_logits, GK * A))) return st;
  if ((st = e->alloc(&e->d_value, GK))) return st;
  e->max_plies = (c.max_plies > 0 && c.max_plies < e->A) ? c.max_plies : e->A;
  const size_t MP = e->max_plies;
  if ((st = e->alloc(&e->d_temperature, MP + 2))) return st;
  if ((st = e->alloc(&e->d_moves, G * MP))) return st;
  if ((st = e->alloc(&e->d_pi_hist, G * MP * A))) return st;
  if ((st = e->alloc(&e->d_result, G))) return st;
  if ((st = e->alloc(&e->d_status, G))) return st;
  e->raw_capacity = (long long)(G * MP < 16384 ? G * MP : 16384);
  if (e->raw_capacity < (long long)MP) e->raw_capacity = MP;
  e->raw_allocated = e->raw_capacity;
  if ((st = e->alloc(&e->d_raw_planes, (size_t)e->raw_capacity * 4 * A))) return st;
  if ((st = e->alloc(&e->d_raw_pi, (size_t)e->raw_capacity * A))) return st;
  if ((st = e->alloc(&e->d_raw_z, (size_t)e->raw_capacity))) return st;
  if ((st = e->alloc(&e->d_raw_count, 2))) return st;
  if ((st = e->alloc(&e->d_sched, 8))) return st;
  if ((st = e->alloc(&e->d_prof_counts, (size_t)AZP_PROF_MAX))) return st;
  AZP_CUDA(e, cudaMemset(tv.n_leaves, 0, 4 * sizeof(int32_t)));
  AZP_CUDA(e, cudaMemset(tv.error_flag, 0, sizeof(int32_t)));
  AZP_CUDA(e, cudaMemset(tv.counters, 0, CNT_COUNT * sizeof(unsigned long long)));
  AZP_CUDA(e, cudaMemset(tv.active, 0, G));
  AZP_CUDA(e, cudaMemset(tv.net_of_game, 0, G));
  AZP_CUDA(e, cudaMemset(tv.game_id, 0, G * sizeof(int64_t)));
  AZP_CUDA(e, cudaMemset(tv.ply, 0, G * sizeof(int32_t)));
  AZP_CUDA(e, cudaMemset(e->d_status, 0, G));
  AZP_CUDA(e, cudaMemset(e->d_raw_count, 0, 2 * sizeof(int32_t)));
  AZP_CUDA(e, cudaMemset(e->d_sched, 0, 8 * sizeof(long long)));
  // default temperature: the reference's self-play schedule (self_play.py:24-26)
  std::vector<double> T(MP + 1);
  for (size_t m = 0; m <= MP; m++) T[m] = (exp(-(double)m / 100.0) + 0.01) / 1.01;
  AZP_CUDA(e, cudaMemcpy(e->d_temperature, T.data(), T.size() * sizeof(double), cudaMemcpyHostToDevice));
  e->n_temperature = (int)T.size();
  if ((st = azp_net_create(e))) return st;
  { const char* v = getenv("AZP_GRAPH"); e->use_graph = v ? atoi(v) != 0 : true; }
  { const char* v = getenv("AZP_ARENA_ALIGNED"); e->arena_aligned = v ? atoi(v) != 0 : 1; }
  { const char* v = getenv("AZP_GRAPH_ITERS"); int k = v ? atoi(v) : 8; e->graph_iters = k < 2 ? 2 : (k > 62 ? 62 : k) & ~1; }
  AZP_CUDA(e, cudaDeviceSynchronize());
  return AZP_OK;
}

extern "C" int azp_create(const azp_config* cfg, azp_engine** out) {
  if (!cfg || !out) return AZP_ERR_INVALID_ARG;
  *out = nullptr;
  auto bad = [](const char* m) { azp_set_global_error(m); return AZP_ERR_INVALID_ARG; };
  if (cfg->board < 3 || cfg->board > AZP_MAX_BOARD) return bad("board must be in [3,19]");
  if (cfg->win_length < 2 || cfg->win_length > cfg->board) return bad("win_length must be in [2,board]");
  if (cfg->games < 1 || cfg->games > 65535) return bad("games must be in [1,65535]");
  if (cfg->sims < 1 || cfg->sims > 1000000) return bad("sims must be >= 1");
  if (cfg->net_kind != AZP_NET_STEM && cfg->net_kind != AZP_NET_TOWER) return bad("unknown net_kind");
  if (cfg->net_dtype != AZP_DTYPE_F32 && cfg->net_dtype != AZP_DTYPE_BF16) return bad("unknown net_dtype");
  if (cfg->net_kind == AZP_NET_TOWER && (cfg->tower_blocks < 1 || cfg->tower_channels < 32 || cfg->tower_channels % 32))
    return bad("tower needs blocks >= 1 and channels a multiple of 32");
  if (cfg->augment_mode < 0 || cfg->augment_mode > 2) return bad("unknown augment_mode");
  if (cfg->world < 1 || cfg->rank < 0 || cfg->rank >= cfg->world) return bad("rank/world invalid");
  if (cfg->leaves_per_step < 0 || cfg->leaves_per_step > 16 || (long long)cfg->games * (cfg->leaves_per_step > 1 ? cfg->leaves_per_step : 1) > 65535)
    return bad("leaves_per_step must be in [0,16] and games*leaves_per_step <= 65535");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    azp_set_global_error("no CUDA device: this engine has no CPU fallback");
    return AZP_ERR_CUDA;
  }
  if (cfg->device < 0 || cfg->device >= ndev) return bad("device ordinal out of range");
  AzpDeviceGuard dev_guard(cfg->device);
  { int cur = -1; if (cudaGetDevice(&cur) != cudaSuccess || cur != cfg->device) { azp_set_global_error("cudaSetDevice failed"); return AZP_ERR_CUDA; } }
  azp_timeline_setup();
  azp_engine* e = new azp_engine();
  e->cfg = *cfg;
  e->A = cfg->board * cfg->board;
  e->NW = AZP_WORDS(cfg->board);
  e->rp = azp_make_rules(cfg->board, cfg->win_length);
  e->policy_logits = cfg->net_dtype == AZP_DTYPE_BF16 ? 1 : 0;
  int st = create_impl(e);
  if (st) {
    azp_set_global_error(e->error.empty() ? "engine creation failed" : e->error);
    azp_destroy(e);
    return st;
  }
  *out = e;
  return AZP_OK;
}

extern "C" void azp_destroy(azp_engine* e) {
  if (!e) return;
  AzpDeviceGuard dev_guard(e->cfg.device);
  cudaSetDevice(e->cfg.device);
  cudaDeviceSynchronize();
  for (cudaEvent_t ev : e->prof_events) cudaEventDestroy(ev);
  if (e->reload_ev) cudaEventDestroy(e->reload_ev);
  for (cudaGraphExec_t ge : e->graph_exec) if (ge) cudaGraphExecDestroy(ge);
  if (e->cfg.net_dtype == AZP_DTYPE_BF16) azp_umma_destroy(e);
  azp_net_destroy(e);
  for (void* p : e->allocs) cudaFree(p);
  delete e;
}

// ---- network ---------------------------------------------------------------------------------------------
extern "C" size_t azp_weight_count(const azp_engine* e) { return e ? azp_net_weight_count(e) : 0; }
extern "C" int azp_load_weights(azp_engine* e, int slot, const float* host_blob, size_t n) {
  if (!e || !host_blob) return AZP_ERR_INVALID_ARG;
  AzpDeviceGuard dev_guard(e->cfg.device);
  return azp_net_load(e, slot, host_blob, n, false, 0);
}
extern "C" int azp_load_weights_device(azp_engine* e, int slot, const float* dev_blob, size_t n, void* stream) {
  if (!e || !dev_blob) return AZP_ERR_INVALID_ARG;
  AzpDeviceGuard dev_guard(e->cfg.device);
  return azp_net_load(e, slot, dev_blob, n, true, (cudaStream_t)stream);
}
extern "C" int azp_net_forward(azp_engine* e, int slot, int n, const float* planes, float* logits, float* policy,
                               float* value, void* stream) {
  if (!e || !planes || n < 0 || slot < 0 || slot > 1) return AZP_ERR_INVALID_ARG;
  AzpDeviceGuard dev_guard(e->cfg.device);
  if (n == 0) return AZP_OK;
  return azp_net_eval_planes(e, slot, n, planes, logits, policy, value, (cudaStream_t)stream);
}

extern "C" int azp_debug_activations(azp_engine* e, int layer, int n, float* out, void* stream) {
  if (!e || !out || n < 1 || n > e->cfg.games) return AZP_ERR_INVALID_ARG;
  AzpDeviceGuard dev_guard(e->cfg.device);
  if (layer == 11) return azp_timeline_read(nullptr, 0, (cudaStream_t)stream);                      // bring-up: clear the ring
  if (layer == 10) return azp_timeline_read(out, (size_t)n * 1024 * 1024, (cudaStream_t)stream);   // bring-up: step timeline ring, n = MiB of `out`
  if (e->cfg.net_dtype != AZP_DTYPE_BF16) { e->set_error("debug activations exist on the bf16 path only"); return AZP_ERR_UNSUPPORTED; }
  return azp_umma_debug(e, layer, n, out, (cudaStream_t)stream);
}

// ---- step-by-step search -----------------------------------------------------------------------------------
__global__ void set_roots_kernel(TreeView tv, int n, const uint64_t* x, const uint64_t* o, const uint8_t* tm,
                                 const int16_t* last, const int32_t* ply, const int64_t* game_id) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= tv.G) return;
  if (g < n) {
    for (int k = 0; k < tv.NW; k++) {
      tv.root_x[(size_t)g * tv.NW + k] = x[(size_t)g * tv.NW + k];
      tv.root_o[(size_t)g * tv.NW + k] = o[(size_t)g * tv.NW + k];
    }
    tv.root_tm[g] = tm[g];
    tv.root_last[g] = last ? last[g] : (int16_t)-1;
    tv.ply[g] = ply ? ply[g] : 0;
    tv.game_id[g] = game_id ? game_id[g] : (int64_t)g;
    tv.active[g] = 1;
    tv.net_of_game[g] = 0;
  } else {
    tv.active[g] = 0;
  }
}

static int check_device_error(azp_engine* e, cudaStream_t s) {
  int32_t flag = 0;
  AZP_CUDA(e, cudaMemcpyAsync(&flag, e->tv.error_flag, sizeof(flag), cudaMemcpyDeviceToHost, s));
  AZP_CUDA(e, cudaStreamSynchronize(s));
  if (flag) {
    e->set_error(flag == AZP_ERR_INVALID_MOVE ? "device reported an illegal move" : "device reported a capacity overflow");
    cudaMemsetAsync(e->tv.error_flag, 0, sizeof(int32_t), s);
    return flag;
  }
  return AZP_OK;
}

extern "C" int azp_search_begin(azp_engine* e, int n, const uint64_t* x, const uint64_t* o, const uint8_t* to_move,
                                const int16_t* last, const int32_t* ply, const int64_t* game_id, void* stream) {
  if (!e || n < 1 || n > e->cfg.games || !x || !o || !to_move) return AZP_ERR_INVALID_ARG;
  AzpDeviceGuard dev_guard(e->cfg.device);
  if (e->K != 1) { e->set_error("the step-by-step search API needs leaves_per_step == 1"); return AZP_ERR_UNSUPPORTED; }
  cudaStream_t s = (cudaStream_t)stream;
  set_roots_kernel<<<(e->cfg.games + 127) / 128, 128, 0, s>>>(e->tv, n, x, o, to_move, last, ply, game_id);
  azp_count_launch();
  e->search_n = n;
  e->parity ^= 1;
  return azp_tree_begin_move(e, n, /*emit=*/0, /*compact=*/0, e->parity, s);
}
extern "C" int azp_root_expand(azp_engine* e, const float* policy, int add_noise, void* stream) {
  if (!e || !policy || e->search_n < 1) return AZP_ERR_INVALID_ARG;
  AzpDeviceGuard dev_guard(e->cfg.device);
  return azp_tree_root_expand(e, e->search_n, policy, /*by_slot=*/0, add_noise, /*logits=*/0, (cudaStream_t)stream);
}
extern "C" int azp_select(azp_engine* e, void* stream) {
  if (!e || e->search_n < 1) return AZP_ERR_INVALID_ARG;
  AzpDeviceGuard dev_guard(e->cfg.device);
  e->parity ^= 1;
  return azp_tree_select(e, e->search_n, /*compact=*/0, e->parity, (cudaStream_t)stream);
}
extern "C" int azp_leaf_states(azp_engine* e, uint64_t* x, uint64_t* o, uint8_t* to_move, int16_t* last, uint8_t* kind,
                               void* stream) {
  if (!e || e->search_n < 1 || !x || !o || !to_move || !last || !kind) return AZP_ERR_INVALID_ARG;
  AzpDeviceGuard dev_guard(e->cfg.device);
  return azp_tree_leaf_states(e, e->search_n, x, o, to_move, last, kind, (cudaStream_t)stream);
}
extern "C" int azp_expand_backup(azp_engine* e, const float* policy, const float* value, void* stream) {
  if (!e || e->search_n < 1 || !policy || !value) return AZP_ERR_INVALID_ARG;
  AzpDeviceGuard dev_guard(e->cfg.device);
  return azp_tree_expand_backup(e, e->search_n, policy, value, /*by_slot=*/0, /*logits=*/0, /*value_f64=*/0,
                                (cudaStream_t)stream);
}
extern "C" int azp_expand_backup_f64(azp_engine* e, const float* policy, const double* value, void* stream) {
  if (!e || e->search_n < 1 || !policy || !value) return AZP_ERR_INVALID_ARG;
  AzpDeviceGuard dev_guard(e->cfg.device);
  return azp_tree_expand_backup(e, e->search_n, policy, value, /*by_slot=*/0, /*logits=*/0, /*value_f64=*/1,
                                (cudaStream_t)stream);
}
extern "C" int azp_root_visits(azp_engine* e, int32_t* visits, void* stream) {
  if (!e || e->search_n < 1 || !visits) return AZP_ERR_INVALID_ARG;
  AzpDeviceGuard dev_guard(e->cfg.device);
  return azp_tree_root_visits(e, e->search_n, visits, (cudaStream_t)stream);
}
extern "C" int azp_root_priors(azp_engine* e, float* priors, void* stream) {
  if (!e || e->search_n < 1 || !priors) return AZP_ERR_INVALID_ARG;
  AzpDeviceGuard dev_guard(e->cfg.device);
  return azp_tree_root_priors(e, e->search_n, priors, (cudaStream_t)stream);
}

// evaluate the leaf batch of counter pair `parity` with the engine's own net(s)
static int eval_step(azp_engine* e, int two_nets, int net_slot, int parity, cudaStream_t s) {
  int st;
  TreeView& tv = e->tv;
  if ((st = azp_net_eval_leaves(e, two_nets ? 0 : net_slot, tv.n_leaves + parity * 2 + 0, 0, 0, s))) return st;
  if (two_nets && (st = azp_net_eval_leaves(e, 1, tv.n_leaves + parity * 2 + 1, 0, 1, s))) return st;
  return AZP_OK;
}

// One move worth of search with the engine's own net(s):
//   root eval -> root expand (+noise) -> descent 0 -> steps x (net -> expand+backup [+ descent of the next step])
// The expand+backup of step i and the descent of step i+1 are ONE launch (advance_kernel).
// The leaf counters are zeroed at the start of every move, so step i always uses counter pair i&1 and the
// sequence is identical from move to move: an even number of consecutive iterations (graph_iters, default 8 = 40
// kernels) is captured ONCE into a CUDA graph and replayed (removes per-launch CPU cost and most launch gaps; a graph
// boundary is a full dependency, so the programmatic overlap of the first kernel is lost once per replay).  The kernels of the chain are launched with
// programmatic dependent launch, inside the graph too.
// `net_slot`: the weight slot that evaluates the (single) leaf batch when two_nets == 0
static int search_move(azp_engine* e, int n, int add_noise, int two_nets, int net_slot, cudaStream_t s) {
  int st;
  TreeView& tv = e->tv;
  AZP_CUDA(e, cudaMemsetAsync(tv.n_leaves, 0, 4 * sizeof(int32_t), s));
  if ((st = azp_tree_begin_move(e, n, 1, 1, /*parity=*/1, s))) return st;
  if ((st = eval_step(e, two_nets, net_slot, 1, s))) return st;
  if ((st = azp_tree_root_expand(e, n, e->d_policy, 1, add_noise, e->policy_logits, s))) return st;
  const int steps = (e->cfg.sims + e->K - 1) / e->K;      // K descents per game per step
  // Exact mode on the bf16 stem: the descent kernels also run conv1 of the leaves they emit (azp_conv1_tail.cuh), and the
  // evaluation that follows starts at conv2.  Every descent of the loop does, the first one included, so that all captured
  // iterations are alike.
  Conv1TailArgs tail;
  const bool fused = !two_nets && azp_umma_conv1_tail(e, net_slot, &tail) != 0;
  if (fused) {
    if ((st = azp_tree_descend_conv1(e, n, true, nullptr, nullptr, 1, e->policy_logits, 1, 0, tail, s))) return st;
  } else {
    if ((st = azp_tree_select(e, n, 1, 0, s))) return st;   // descent of step 0
  }
  // iteration i: net(pair i&1), then expand+backup(i) fused with the descent of step i+1 (pair (i+1)&1)
  auto iteration = [&](int i, bool last) -> int {
    e->conv1_done = fused;
    int r = eval_step(e, two_nets, net_slot, i & 1, s);
    e->conv1_done = false;
    if (r) return r;
    if (last) return azp_tree_expand_backup(e, n, e->d_policy, e->d_value, 1, e->policy_logits, 0, s);
    if (fused) return azp_tree_descend_conv1(e, n, false, e->d_policy, e->d_value, 1, e->policy_logits, 1, (i + 1) & 1, tail, s);
    return azp_tree_advance(e, n, e->d_policy, e->d_value, 1, e->policy_logits, 1, (i + 1) & 1, s);
  };
  int sim = 0;
  const bool want_graph = e->use_graph && steps >= 5;
  // Profiling (azp_set_profiling): CUDA event records cannot bracket kernels inside a replayed graph one launch at a time,
  // so the last AZP_PROF_TAIL iterations of the move run as plain launches, each profiled kernel between its own pair of
  // events; the iterations before them (and every move step with profiling off) replay the captured graph.
  const int plain_tail = e->profiling ? (steps < AZP_PROF_TAIL ? steps : AZP_PROF_TAIL) : 1;
  if (want_graph) {
    const int key = (n * 2 + two_nets) * 64 + e->graph_iters;
    const int gi = two_nets ? 0 : (net_slot & 1);
    if (e->graph_exec[gi] && (e->graph_key[gi] != key || e->graph_epoch[gi] != e->weights_epoch)) {
      cudaGraphExecDestroy(e->graph_exec[gi]);
      e->graph_exec[gi] = nullptr;
    }
    if (!e->graph_exec[gi]) {
      // capture is best effort: any failure falls back to plain stream launches for this engine
      cudaGraph_t graph = nullptr;
      const uint64_t launches_before = azp_launch_count();
      cudaError_t ce = cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
      if (ce == cudaSuccess) {
        e->capturing = true;
        for (int k = 0; k < e->graph_iters && !st; k++) st = iteration(k, false);
        e->capturing = false;
        ce = cudaStreamEndCapture(s, &graph);
        if (!st && ce == cudaSuccess && graph) ce = cudaGraphInstantiate(&e->graph_exec[gi], graph, 0);
        if (graph) cudaGraphDestroy(graph);
      }
      if (st || ce != cudaSuccess || !e->graph_exec[gi]) {
        if (getenv("AZP_GRAPH_DEBUG")) fprintf(stderr, "[azp] graph capture failed: st=%d cuda=%s\n", st, cudaGetErrorString(ce));
        cudaGetLastError();                     // clear the sticky-free error state
        e->graph_exec[gi] = nullptr;
        e->use_graph = false;
      } else {
        e->graph_key[gi] = key;
        e->graph_epoch[gi] = e->weights_epoch;
        e->graph_launches_per_replay[gi] = (int)(azp_launch_count() - launches_before);   // kernels per replay
      }
    }
    // the last iteration (no descent after it) always runs as plain launches
    e->note_stream(s);
    for (; e->graph_exec[gi] && sim + e->graph_iters <= steps - plain_tail; sim += e->graph_iters) {
      AZP_CUDA(e, cudaGraphLaunch(e->graph_exec[gi], s));
      azp_count_launch(e->graph_launches_per_replay[gi]);
    }
  }
  for (; sim < steps; sim++)
    if ((st = iteration(sim, sim == steps - 1))) return st;
  return AZP_OK;
}

extern "C" int azp_search(azp_engine* e, int slot, int n, const uint64_t* x, const uint64_t* o, const uint8_t* to_move,
                          const int16_t* last, const int32_t* ply, const int64_t* game_id, int add_noise,
                          int32_t* visits, void* stream) {
  if (!e || n < 1 || n > e->cfg.games || !x || !o || !to_move || slot != 0) return AZP_ERR_INVALID_ARG;
  AzpDeviceGuard dev_guard(e->cfg.device);
  cudaStream_t s = (cudaStream_t)stream;
  set_roots_kernel<<<(e->cfg.games + 127) / 128, 128, 0, s>>>(e->tv, n, x, o, to_move, last, ply, game_id);
  azp_count_launch();
  e->search_n = n;
  int st = search_move(e, n, add_noise, 0, 0, s);
  if (st) return st;
  if (visits && (st = azp_tree_root_visits(e, n, visits, s))) return st;
  return check_device_error(e, s);
}

// ---- drivers ------------------------------------------------------------------------------------------------
extern "C" int azp_set_slab_capacity(azp_engine* e, int64_t positions) {
  if (!e || positions < 1) return AZP_ERR_INVALID_ARG;
  long long cap = positions < e->raw_allocated ? positions : e->raw_allocated;
  if (cap < e->max_plies) cap = e->max_plies;            // one whole game must always fit
  e->raw_capacity = cap;
  return AZP_OK;
}
extern "C" int azp_set_augment_mode(azp_engine* e, int mode) {
  if (!e || mode < 0 || mode > 2) return AZP_ERR_INVALID_ARG;
  e->cfg.augment_mode = mode;          // read by the self-play driver at every call
  return AZP_OK;
}
extern "C" int azp_set_temperature_table(azp_engine* e, const double* host_table, int n) {
  if (!e || !host_table || n < 1) return AZP_ERR_INVALID_ARG;
  AzpDeviceGuard dev_guard(e->cfg.device);
  if (n > e->max_plies + 1) n = e->max_plies + 1;
  AZP_CUDA(e, cudaDeviceSynchronize());
  AZP_CUDA(e, cudaMemcpy(e->d_temperature, host_table, (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
  e->n_temperature = n;
  return AZP_OK;
}

// this rank's contiguous block of the global game ids [first, first+total)
static void rank_block(const azp_engine* e, long long first, long long total, long long* lfirst, long long* lcount) {
  long long W = e->cfg.world, r = e->cfg.rank;
  long long base = total / W, rem = total % W;
  long long off = r * base + (r < rem ? r : rem);
  *lfirst = first + off;
  *lcount = base + (r < rem ? 1 : 0);
}

struct MoveStepReport { long long sched[8]; long long out_written; int32_t error; int32_t raw_count; };

static int episodes_run(azp_engine* e, long long first_game, long long n_games, int resume, int max_move_steps,
                        int arena, float* out_planes, float* out_pi, int32_t* out_z, long long capacity,
                        long long* host_written, cudaStream_t s) {
  const int G = e->cfg.games;
  const int mode = e->cfg.augment_mode;
  const int mult = mode == AZP_AUGMENT_D4 ? 8 : (mode == AZP_AUGMENT_REFERENCE ? 4 : 1);
  if (!e->weights[0] || (arena && !e->weights[1])) { e->set_error("weights not loaded"); return AZP_ERR_NO_WEIGHTS; }
  if (!arena && capacity < e->raw_capacity * mult) {
    e->set_error("output capacity must be at least " + std::to_string(e->raw_capacity * mult) + " examples");
    return AZP_ERR_INVALID_ARG;
  }
  long long* d_out_written = e->d_sched + 4;
  int st;
  // arena mode of the kernels: 1 = two leaf batches per step, 2 = phase-aligned (one batch, one net per step; see azp_tree.cu)
  const int amode = arena ? (e->arena_aligned ? 2 : 1) : 0;
  if (!resume) {
    e->arena_step = 0;
    long long lfirst, lcount;
    rank_block(e, first_game, n_games, &lfirst, &lcount);
    long long sched[8] = {0, lcount, 0, lfirst, 0, 0, 0, 0};
    AZP_CUDA(e, cudaMemcpyAsync(e->d_sched, sched, sizeof(sched), cudaMemcpyHostToDevice, s));
    AZP_CUDA(e, cudaMemsetAsync(e->d_raw_count, 0, 2 * sizeof(int32_t), s));
    AZP_CUDA(e, cudaMemsetAsync(e->d_status, 0, G, s));
    AZP_CUDA(e, cudaStreamSynchronize(s));     // `sched` is a stack buffer
    if ((st = azp_episode_seat(e, G, amode, /*next_phase=*/0, s))) return st;
    e->sp_total = lcount;
  } else {
    AZP_CUDA(e, cudaMemsetAsync(d_out_written, 0, sizeof(long long), s));
  }
  long long written = 0;
  *host_written = 0;
  e->search_n = G;                  // azp_root_visits / azp_root_priors read the trees of the last move step
  if (e->sp_total == 0) return AZP_OK;
  for (int step = 0; max_move_steps <= 0 || step < max_move_steps; step++) {
    if (!arena && capacity - written < e->raw_capacity * mult) { *host_written = written; return AZP_MORE; }
    const int phase = (int)(e->arena_step & 1);
    if (amode == 2 && (st = azp_episode_activate(e, G, phase, s))) return st;
    if ((st = search_move(e, G, arena ? 0 : 1, amode == 1, amode == 2 ? phase : 0, s))) return st;
    if ((st = azp_episode_end_move(e, G, amode, s))) return st;
    if ((st = azp_episode_finish(e, G, amode, /*next_phase=*/phase ^ 1, s))) return st;
    e->arena_step++;
    if (!arena) {
      azp_prof_begin(e, AZP_PROF_AUGMENT, e->d_raw_count, 0, s);        // units = raw positions of this move step
      if ((st = azp_augment_launch(e->cfg.board, mode, (int)e->raw_capacity, e->d_raw_count, e->d_raw_planes,
                                   e->d_raw_pi, e->d_raw_z, out_planes, out_pi, out_z, 0, d_out_written, capacity, s)))
        return st;
      azp_prof_end(e, AZP_PROF_AUGMENT, s);
      if ((st = azp_episode_drain(e, d_out_written, mult, s))) return st;
    }
    // one small read-back per move step (a move step is `sims` network evaluations long)
    MoveStepReport rep;
    AZP_CUDA(e, cudaMemcpyAsync(rep.sched, e->d_sched, sizeof(rep.sched), cudaMemcpyDeviceToHost, s));
    AZP_CUDA(e, cudaMemcpyAsync(&rep.error, e->tv.error_flag, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    AZP_CUDA(e, cudaStreamSynchronize(s));
    written = rep.sched[4];
    *host_written = written;
    if (rep.error) {
      e->set_error(rep.error == AZP_ERR_INVALID_MOVE ? "device reported an illegal move" : "device reported a capacity overflow");
      cudaMemsetAsync(e->tv.error_flag, 0, sizeof(int32_t), s);
      return rep.error;
    }
    if (rep.sched[2] >= rep.sched[1]) return AZP_OK;      // every game of this rank is finished and flushed
  }
  return AZP_MORE;
}

extern "C" int azp_selfplay_run(azp_engine* e, int64_t first_game, int64_t n_games, int resume, int max_move_steps,
                                float* out_planes, float* out_pi, int32_t* out_z, int64_t capacity,
                                int64_t* host_written, void* stream) {
  if (!e || n_games < 0 || !out_planes || !out_pi || !out_z || !host_written) return AZP_ERR_INVALID_ARG;
  AzpDeviceGuard dev_guard(e->cfg.device);
  long long w = 0;
  int st = episodes_run(e, first_game, n_games, resume, max_move_steps, 0, out_planes, out_pi, out_z, capacity, &w,
                        (cudaStream_t)stream);
  *host_written = w;
  return st;
}

extern "C" int azp_arena_run_steps(azp_engine* e, int64_t first_game, int64_t n_games, int resume, int max_move_steps,
                                   int64_t* host_wld, void* stream) {
  if (!e || n_games < 0 || !host_wld) return AZP_ERR_INVALID_ARG;
  AzpDeviceGuard dev_guard(e->cfg.device);
  cudaStream_t s = (cudaStream_t)stream;
  unsigned long long after[CNT_COUNT];
  if (!resume) {       // the tally is the counters' growth since the match started
    AZP_CUDA(e, cudaMemcpyAsync(e->arena_before, e->tv.counters, sizeof(e->arena_before), cudaMemcpyDeviceToHost, s));
    AZP_CUDA(e, cudaStreamSynchronize(s));
  }
  long long w = 0;
  int st = episodes_run(e, first_game, n_games, resume, max_move_steps, 1, nullptr, nullptr, nullptr, 0, &w, s);
  if (st != AZP_OK && st != AZP_MORE) return st;
  AZP_CUDA(e, cudaMemcpyAsync(after, e->tv.counters, sizeof(after), cudaMemcpyDeviceToHost, s));
  AZP_CUDA(e, cudaStreamSynchronize(s));
  host_wld[0] = (int64_t)(after[CNT_XWINS] - e->arena_before[CNT_XWINS]);     // candidate plays X
  host_wld[1] = (int64_t)(after[CNT_OWINS] - e->arena_before[CNT_OWINS]);
  host_wld[2] = (int64_t)(after[CNT_DRAWS] - e->arena_before[CNT_DRAWS]);
  return st;
}
extern "C" int azp_arena_run(azp_engine* e, int64_t first_game, int64_t n_games, int64_t* host_wld, void* stream) {
  AzpDeviceGuard dev_guard(e->cfg.device);
  return azp_arena_run_steps(e, first_game, n_games, 0, 0, host_wld, stream);
}

// Episode state of every slot as the drivers hold it (plain device-to-device copies; any pointer may be null).
extern "C" int azp_episode_snapshot(azp_engine* e, int32_t* ply, int64_t* game_id, uint8_t* status, uint16_t* moves,
                                    float* pi_hist, void* stream) {
  if (!e) return AZP_ERR_INVALID_ARG;
  AzpDeviceGuard dev_guard(e->cfg.device);
  cudaStream_t s = (cudaStream_t)stream;
  const size_t G = e->cfg.games, MP = e->max_plies, A = e->A;
  if (ply) AZP_CUDA(e, cudaMemcpyAsync(ply, e->tv.ply, G * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
  if (game_id) AZP_CUDA(e, cudaMemcpyAsync(game_id, e->tv.game_id, G * sizeof(int64_t), cudaMemcpyDeviceToDevice, s));
  if (status) AZP_CUDA(e, cudaMemcpyAsync(status, e->d_status, G, cudaMemcpyDeviceToDevice, s));
  if (moves) AZP_CUDA(e, cudaMemcpyAsync(moves, e->d_moves, G * MP * sizeof(uint16_t), cudaMemcpyDeviceToDevice, s));
  if (pi_hist) AZP_CUDA(e, cudaMemcpyAsync(pi_hist, e->d_pi_hist, G * MP * A * sizeof(float), cudaMemcpyDeviceToDevice, s));
  return AZP_OK;
}

extern "C" int azp_get_counters(azp_engine* e, azp_counters* out) {
  if (!e || !out) return AZP_ERR_INVALID_ARG;
  AzpDeviceGuard dev_guard(e->cfg.device);
  unsigned long long c[CNT_COUNT];
  AZP_CUDA(e, cudaDeviceSynchronize());
  AZP_CUDA(e, cudaMemcpy(c, e->tv.counters, sizeof(c), cudaMemcpyDeviceToHost));
  out->simulations = c[CNT_SIMS]; out->leaf_evals = c[CNT_EVALS]; out->terminal_leaves = c[CNT_TERMINAL];
  out->children_scanned = c[CNT_SCANNED]; out->path_length = c[CNT_PATH]; out->expansions = c[CNT_EXPANSIONS];
  out->games_finished = c[CNT_GAMES]; out->plies = c[CNT_PLIES]; out->positions = c[CNT_POSITIONS];
  out->examples = c[CNT_EXAMPLES]; out->kernel_launches = azp_launch_count();
  out->x_wins = c[CNT_XWINS]; out->o_wins = c[CNT_OWINS]; out->draws = c[CNT_DRAWS];
  return AZP_OK;
}
// ---- per-kernel profiling ---------------------------------------------------------------------------------
// Every launch of a profiled kernel class (AZP_PROF_*) is bracketed by its own pair of CUDA events on the launching
// stream while profiling is on (never inside a graph capture).  `units` = what the launch processed (leaves for the net
// kernels, game slots for the tree / episode kernels), read on the device at launch time when it is only known there.
__global__ void prof_copy_count_kernel(const int32_t* src, int32_t* dst) { *dst = *src; }

void azp_prof_begin(azp_engine* e, int kind, const int32_t* count_dev, int count_host, cudaStream_t s) {
  if (!e->profiling || e->capturing) return;
  if (e->prof_used >= AZP_PROF_MAX - 2) azp_prof_collect(e);
  const size_t i = e->prof_used;
  while (e->prof_events.size() < 2 * (i + 1)) {
    cudaEvent_t ev;
    cudaEventCreate(&ev);
    e->prof_events.push_back(ev);
  }
  if (count_dev) { prof_copy_count_kernel<<<1, 1, 0, s>>>(count_dev, e->d_prof_counts + i); azp_count_launch(); }
  e->prof_kind.push_back(kind);
  e->prof_count_dev.push_back(count_dev ? e->d_prof_counts + i : nullptr);
  e->prof_count_host.push_back(count_host);
  cudaEventRecord(e->prof_events[2 * i], s);
}
void azp_prof_end(azp_engine* e, int kind, cudaStream_t s) {
  (void)kind;
  if (!e->profiling || e->capturing) return;
  cudaEventRecord(e->prof_events[2 * e->prof_used + 1], s);
  e->prof_used++;
}
int azp_prof_collect(azp_engine* e) {
  if (e->prof_used == 0) return AZP_OK;
  AZP_CUDA(e, cudaDeviceSynchronize());
  std::vector<int32_t> counts(AZP_PROF_MAX);
  AZP_CUDA(e, cudaMemcpy(counts.data(), e->d_prof_counts, AZP_PROF_MAX * sizeof(int32_t), cudaMemcpyDeviceToHost));
  for (size_t i = 0; i < e->prof_used; i++) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, e->prof_events[2 * i], e->prof_events[2 * i + 1]) != cudaSuccess) { cudaGetLastError(); continue; }
    const int units = e->prof_count_dev[i] ? counts[i] : e->prof_count_host[i];
    if (units <= 0) continue;
    const int k = e->prof_kind[i];
    e->prof_ms[k] += ms;
    e->prof_launches[k]++;
    e->prof_units[k] += units;
  }
  e->prof_used = 0;
  e->prof_kind.clear();
  e->prof_count_dev.clear();
  e->prof_count_host.clear();
  return AZP_OK;
}
extern "C" int azp_set_profiling(azp_engine* e, int enabled) {
  if (!e) return AZP_ERR_INVALID_ARG;
  AzpDeviceGuard dev_guard(e->cfg.device);
  int st = azp_prof_collect(e);
  e->profiling = enabled != 0;
  return st;
}
extern "C" int azp_get_profile_kernel(azp_engine* e, int kind, double* ms, uint64_t* launches, double* units) {
  if (!e || !ms || !launches || !units || kind < 0 || kind >= AZP_PROF_KINDS) return AZP_ERR_INVALID_ARG;
  AzpDeviceGuard dev_guard(e->cfg.device);
  int st = azp_prof_collect(e);
  *ms = e->prof_ms[kind]; *launches = e->prof_launches[kind]; *units = e->prof_units[kind];
  e->prof_ms[kind] = 0.0; e->prof_units[kind] = 0.0; e->prof_launches[kind] = 0;
  return st;
}
extern "C" int azp_get_profile(azp_engine* e, double* ms, uint64_t* launches, double* flops) {
  double units = 0.0;
  int st = azp_get_profile_kernel(e, AZP_PROF_DOMINANT, ms, launches, &units);
  if (st == AZP_OK && flops) *flops = units * e->prof_flops_per_board;
  return st;
}

extern "C" int azp_reset_counters(azp_engine* e) {
  if (!e) return AZP_ERR_INVALID_ARG;
  AzpDeviceGuard dev_guard(e->cfg.device);
  AZP_CUDA(e, cudaDeviceSynchronize());
  AZP_CUDA(e, cudaMemset(e->tv.counters, 0, CNT_COUNT * sizeof(unsigned long long)));
  return AZP_OK;
}
