// Internal declarations shared by the translation units of libazp_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <string.h>
#include <utility>
#include <vector>
#include "azp_common.cuh"

// ---- launch bookkeeping / error plumbing ------------------------------------------------
void azp_count_launch(int n = 1);
uint64_t azp_launch_count();
int azp_check_launch();                       // cudaGetLastError -> status, records message
void azp_set_global_error(const std::string& msg);
const char* azp_global_error();
RulesParams azp_make_rules(int board, int win_length);

#define AZP_DISPATCH_NW(nw, kernel, grid, block, smem, stream, ...)                         \
  do {                                                                                      \
    switch (nw) {                                                                           \
      case 1: kernel<1><<<grid, block, smem, stream>>>(__VA_ARGS__); break;                 \
      case 2: kernel<2><<<grid, block, smem, stream>>>(__VA_ARGS__); break;                 \
      case 3: kernel<3><<<grid, block, smem, stream>>>(__VA_ARGS__); break;                 \
      case 4: kernel<4><<<grid, block, smem, stream>>>(__VA_ARGS__); break;                 \
      case 5: kernel<5><<<grid, block, smem, stream>>>(__VA_ARGS__); break;                 \
      default: kernel<6><<<grid, block, smem, stream>>>(__VA_ARGS__); break;                \
    }                                                                                       \
  } while (0)

// ---- programmatic dependent launch (PDL) -------------------------------------------------------------
// The kernels of a simulation step form a strict chain (select -> conv1 -> conv2 -> conv3 -> fc ->
// expand/backup -> select ...).  Launched with the programmatic-stream-serialization attribute, kernel
// N+1 may become resident and run its prologue (barrier init, TMEM allocation, weight loads into shared
// memory -- nothing that kernel N writes) while kernel N drains; `pdl_wait()` then blocks until kernel N
// has completed and its writes are visible.  Every kernel launched through azp_launch(..., pdl=true) MUST
// call pdl_wait() before its first access to data produced on the stream.  Both instructions are no-ops
// in a launch without the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool azp_pdl_enabled();                        // AZP_PDL (default 1)

// ---- step timeline (bring-up builds only: make BRINGUP=1, AZP_TIMELINE=1) ------------------------------------
// Every kernel of the simulation step records, per launch, the earliest CTA entry, the earliest moment a CTA got past
// pdl_wait() and the latest CTA exit as %globaltimer nanoseconds: the real overlap / gaps of the chain inside the
// captured graph, which neither ncu (serialised) nor event brackets (they cut the dependent-launch edges) can show.
// tools/step_timeline.py reads the ring (done[k] / participants = launches so far).
#define AZP_TL_KERNELS 8
#define AZP_TL_RING 4096
enum { AZP_TL_ADVANCE = 0, AZP_TL_CONV1, AZP_TL_CONV2, AZP_TL_CONV3, AZP_TL_FC };
struct AzpTimeline {
  unsigned long long rec[AZP_TL_KERNELS][AZP_TL_RING][4];     // ~entry (min), ~start (min), end (max), spare
  unsigned int launch[AZP_TL_KERNELS], done[AZP_TL_KERNELS];
  unsigned long long cta[AZP_TL_KERNELS][320][4];             // last launch: per-CTA start (past pdl_wait) / end / SM id / sum of durations
};
#if defined(AZP_BRINGUP) || defined(AZP_TIMELINE_BUILD)
extern __device__ AzpTimeline* g_azp_tl;
__device__ __forceinline__ unsigned long long azp_gtime() {
  unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t;
}
// One caller per participant (a CTA's thread 0, or lane 0 of every warp).  Nothing here waits for memory: the stamps
// and the participant count are fire-and-forget reductions, so the instrumented chain keeps its timing.  The launch's
// ring slot = participants that have finished so far / participants per launch: the previous launch of the SAME kernel
// ended a whole simulation step ago, so the count is an exact multiple when a launch starts.
__device__ __forceinline__ unsigned tl_entry(int kid, unsigned participants) {
  AzpTimeline* t = g_azp_tl;
  if (!t) return 0;
  const unsigned idx = (*reinterpret_cast<volatile unsigned*>(&t->done[kid]) / participants) % AZP_TL_RING;
  atomicMax(&t->rec[kid][idx][0], ~azp_gtime());
  return idx;
}
__device__ __forceinline__ void tl_start(int kid, unsigned idx) {
  AzpTimeline* t = g_azp_tl;
  if (t) {
    const unsigned long long now = azp_gtime();
    atomicMax(&t->rec[kid][idx][1], ~now);
    const unsigned b = blockIdx.x + blockIdx.y * gridDim.x;
    if (b < 320u && (threadIdx.x & 127) == 0) t->cta[kid][b][0] = now;
  }
}
__device__ __forceinline__ void tl_end(int kid, unsigned idx, unsigned participants) {
  AzpTimeline* t = g_azp_tl;
  if (!t) return;
  const unsigned long long now = azp_gtime();
  atomicMax(&t->rec[kid][idx][2], now);
  atomicAdd(&t->done[kid], 1u);
  const unsigned b = blockIdx.x + blockIdx.y * gridDim.x;
  if (b < 320u && (threadIdx.x & 127) == 0) {
    unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    t->cta[kid][b][1] = now; t->cta[kid][b][2] = smid; t->cta[kid][b][3] += now - t->cta[kid][b][0];
  }
  if (idx == 0xFFFFFFFFu) t->launch[kid] = participants;      // never true: keeps the field referenced
}
#define AZP_TL_ENTRY(kid, n) tl_entry(kid, n)
#define AZP_TL_START(kid, idx) tl_start(kid, idx)
#define AZP_TL_END(kid, idx, n) tl_end(kid, idx, n)
#else
#define AZP_TL_ENTRY(kid, n) 0u
#define AZP_TL_START(kid, idx) do { (void)(idx); } while (0)
#define AZP_TL_END(kid, idx, n) do { (void)(idx); } while (0)
#endif
int azp_timeline_read(void* out_dev, size_t bytes, cudaStream_t s);      // bring-up builds: copy the ring; else AZP_ERR_UNSUPPORTED

template <typename... P, typename... A>
inline cudaError_t azp_launch(void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, bool pdl,
                              A&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (pdl && azp_pdl_enabled()) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<A>(args)...);
}

#define AZP_DISPATCH_NW_PDL(nw, kernel, grid, block, smem, stream, ...)                     \
  do {                                                                                      \
    switch (nw) {                                                                           \
      case 1: azp_launch(kernel<1>, grid, block, smem, stream, true, __VA_ARGS__); break;   \
      case 2: azp_launch(kernel<2>, grid, block, smem, stream, true, __VA_ARGS__); break;   \
      case 3: azp_launch(kernel<3>, grid, block, smem, stream, true, __VA_ARGS__); break;   \
      case 4: azp_launch(kernel<4>, grid, block, smem, stream, true, __VA_ARGS__); break;   \
      case 5: azp_launch(kernel<5>, grid, block, smem, stream, true, __VA_ARGS__); break;   \
      default: azp_launch(kernel<6>, grid, block, smem, stream, true, __VA_ARGS__); break;  \
    }                                                                                       \
  } while (0)

#define AZP_CUDA(e, call)                                                                   \
  do {                                                                                      \
    cudaError_t err__ = (call);                                                             \
    if (err__ != cudaSuccess) {                                                             \
      (e)->set_error(std::string(#call) + ": " + cudaGetErrorString(err__));                \
      return AZP_ERR_CUDA;                                                                  \
    }                                                                                       \
  } while (0)

// ---- device selection -------------------------------------------------------------------------------------------
// Every entry point that takes an engine / trainer runs on THAT object's device whatever the caller's current device
// is, and leaves the caller's current device as it found it (a process may hold engines on several GPUs).
struct AzpDeviceGuard {
  int prev = -1; bool changed = false;
  explicit AzpDeviceGuard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) changed = cudaSetDevice(dev) == cudaSuccess;
  }
  ~AzpDeviceGuard() { if (changed) cudaSetDevice(prev); }
  AzpDeviceGuard(const AzpDeviceGuard&) = delete;
  AzpDeviceGuard& operator=(const AzpDeviceGuard&) = delete;
};

// ---- device-side views ------------------------------------------------------------------
// Structure-of-arrays node table.  Node `nd` of game `g` owns A edge slots, one per CELL
// (slot index == r*B+c, so the reference's row-major child order is the slot order and
// occupied cells are simply skipped).  Edge statistics are the child NODE's N/W of
// alphazero/mcts.py:25-48; the root's own N lives in root_n.
struct TreeView {
  int G, A, NW, cap, maxd, S;
  int K, slots;      // descents per game per step (virtual loss if > 1); leaf-batch rows = G*K
  int32_t* N;        // [G][cap][A] visit count of the child reached through the edge
  double* W;         // [G][cap][A] total value (float64, as the reference's Python floats)
  float* P;          // [G][cap][A] prior (float32 widened at use, mcts.py:63)
  int32_t* child;    // [G][cap][A] node index of the expanded child, -1 if not expanded
  int32_t* node_count;   // [G]
  int32_t* root_n;       // [G]  root.N (== simulations done this move)
  // per-game root position
  uint64_t* root_x; uint64_t* root_o;      // [G][NW]
  uint8_t* root_tm; int16_t* root_last;    // [G]
  int32_t* ply; int64_t* game_id;          // [G]
  uint8_t* root_result;                    // [G] result of the ROOT position (0 if playable)
  uint8_t* active;                         // [G] 1 = this slot searches
  uint8_t* net_of_game;                    // [G] weight slot that evaluates this game's leaves
  // per-step leaf records
  // per-step leaf records, one per (game, descent k): index g*K + k
  int32_t* path;         // [G*K][maxd] (node*A + cell) of every edge on the descent
  int32_t* path_len;     // [G*K]
  uint8_t* leaf_kind;    // [G*K] 0 = evaluate+expand, 1 = terminal, 2 = idle
  float* leaf_tvalue;    // [G*K] value of a terminal leaf from its side to move's view
  int32_t* leaf_slot;    // [G*K] row of the leaf in the contiguous evaluation batch (-1 if none)
  int32_t* slot_game;    // [slots] inverse
  uint64_t* leaf_x; uint64_t* leaf_o;      // [slots][NW] indexed by SLOT (contiguous batch)
  uint8_t* leaf_tm; int16_t* leaf_last;    // [slots]     indexed by SLOT
  int32_t* n_leaves;     // [4]: {front count, back count} x step parity
  int32_t* error_flag;   // [1] device-side error (capacity overflow, illegal move)
  unsigned long long* counters;  // azp_counters as 64-bit words
};

enum CounterIndex {
  CNT_SIMS = 0, CNT_EVALS, CNT_TERMINAL, CNT_SCANNED, CNT_PATH, CNT_EXPANSIONS, CNT_GAMES, CNT_PLIES,
  CNT_POSITIONS, CNT_EXAMPLES, CNT_LAUNCHES, CNT_XWINS, CNT_OWINS, CNT_DRAWS, CNT_COUNT
};

#define AZP_PROF_MAX 4096      // bracketed launches buffered between two collects
#define AZP_PROF_TAIL 40       // iterations of a move step that run as plain (bracketable) launches while profiling

// ---- network ------------------------------------------------------------------------------
struct NetWeights;       // defined in azp_net.cu
struct NetWorkspace;

struct azp_engine {
  azp_config cfg;
  RulesParams rp;
  TreeView tv;
  int A, NW;
  int K = 1, slots = 0;            // leaves per game per step; rows of the evaluation batch (games * K)
  std::string error;
  std::vector<void*> allocs;
  std::vector<void*>* alloc_pool = nullptr;   // when set, alloc() charges this list (weights: freed on reload)
  bool alloc_reuse = false; size_t alloc_cursor = 0;   // reload: hand out the pool's existing buffers in order
  NetWeights* weights[2] = {nullptr, nullptr};
  NetWorkspace* ws = nullptr;
  // evaluation outputs indexed by slot
  float* d_policy = nullptr;   // [G][A]  fp32 net: probabilities; bf16 net: RAW LOGITS (policy_logits = 1), softmax in the consumer
  int policy_logits = 0;
  float* d_logits = nullptr;   // [G][A]
  float* d_value = nullptr;    // [G]
  // self-play state
  double* d_temperature = nullptr; int n_temperature = 0;
  uint16_t* d_moves = nullptr;     // [G][max_plies]
  float* d_pi_hist = nullptr;      // [G][max_plies][A]
  uint8_t* d_result = nullptr;     // [G] result of a finished game awaiting its flush
  uint8_t* d_status = nullptr;     // [G] slot status (see azp_tree.cu)
  float* d_raw_planes = nullptr; float* d_raw_pi = nullptr; int32_t* d_raw_z = nullptr;
  int32_t* d_raw_count = nullptr;  // [1]
  long long raw_capacity = 0, raw_allocated = 0;     // positions one move step may emit (azp_set_slab_capacity) / rows allocated
  long long* d_sched = nullptr;    // [4]: next local game index, games to play, finished, (spare)
  int max_plies = 0;
  int search_n = 0;                // number of game slots in the current step-by-step search
  int parity = 0;                  // which pair of leaf counters the current step uses (step-by-step API)
  // CUDA graph of two consecutive simulation steps (see search_move)
  bool use_graph = true;
  cudaGraphExec_t graph_exec[2] = {nullptr, nullptr};    // one per weight slot that evaluates the front batch
  int graph_key[2] = {-1, -1}, graph_launches_per_replay[2] = {0, 0};
  uint64_t weights_epoch = 0, graph_epoch[2] = {0, 0};
  int graph_iters = 8;             // simulation steps per captured graph (even; AZP_GRAPH_ITERS)
  bool conv1_done = false;         // the descent kernel that emitted the current leaf batch has already written its act1 (azp_conv1_tail.cuh)
  int arena_aligned = 1;           // phase-aligned arena (AZP_ARENA_ALIGNED=0: two leaf batches per step)
  long long arena_step = 0;        // move steps since the current match started (its parity is the phase)
  bool capturing = false;          // inside cudaStreamBeginCapture/EndCapture of the step graph
  // per-kernel profiling of the dominant kernel (roofline.achieved in bench.py)
  bool profiling = false;
  std::vector<cudaEvent_t> prof_events;      // pairs (start, stop), recycled after each collect
  std::vector<const int32_t*> prof_count_dev; std::vector<int> prof_count_host, prof_kind;
  size_t prof_used = 0;
  double prof_ms[AZP_PROF_KINDS] = {0}, prof_units[AZP_PROF_KINDS] = {0}; uint64_t prof_launches[AZP_PROF_KINDS] = {0};
  double prof_flops_per_board = 0.0;    // algorithmic FLOPs of the dominant kernel per leaf
  int32_t* d_prof_counts = nullptr;          // [PROF_MAX] leaf counts copied at launch time
  long long sp_first = 0, sp_total = 0;  // current self-play job
  unsigned long long arena_before[CNT_COUNT] = {0};   // counters when the current arena match started
  // streams this engine has launched network kernels on: a weight reload waits for THESE (not for the whole device,
  // which would also wait for a caller's copy stream draining examples to the host and serialise it with the search)
  std::vector<cudaStream_t> net_streams;
  cudaEvent_t reload_ev = nullptr; cudaStream_t reload_stream = nullptr;   // last device-side weight refresh (see azp_net_load)
  // called wherever network kernels are about to be launched on `s`
  void note_stream(cudaStream_t s) {
    if (reload_ev && s != reload_stream) cudaStreamWaitEvent(s, reload_ev, 0);   // the refresh ran on another stream
    for (cudaStream_t q : net_streams) if (q == s) return;
    net_streams.push_back(s);
  }
  void set_error(const std::string& m) { error = m; azp_set_global_error(m); }
  template <typename T> int alloc(T** p, size_t count) {
    if (alloc_reuse && alloc_pool && alloc_cursor < alloc_pool->size()) { *p = (T*)(*alloc_pool)[alloc_cursor++]; return AZP_OK; }
    void* q = nullptr;
    cudaError_t err = cudaMalloc(&q, count * sizeof(T) + 256);
    if (err != cudaSuccess) { set_error(std::string("cudaMalloc: ") + cudaGetErrorString(err)); return AZP_ERR_CUDA; }
    (alloc_pool ? *alloc_pool : allocs).push_back(q);
    *p = (T*)q;
    return AZP_OK;
  }
};

// geometry of the padded NHWC activation layout (see azp_net_umma.cu)
struct ConvGeom {
  int B, A, P, stride;        // board side, cells, row pitch, rows per board = (B+1)*P
  int G;                      // slots
};
// conv1 of freshly emitted leaves as the tail of the descent kernels (azp_conv1_tail.cuh); filled in by azp_umma_conv1_tail
struct Conv1TailArgs {
  ConvGeom g;
  const __nv_bfloat16* w;          // [10][64][8] (UmmaWeights::w1_taps of the slot that evaluates these leaves)
  __nv_bfloat16* out;              // act1 [rows][32]
  int enabled;                     // 0: no tail; else the warp groups per CTA (1, 2 or 4: four games each)
};
int azp_umma_conv1_tail(azp_engine* e, int slot, Conv1TailArgs* out);      // 1: the bf16 stem of `slot` can take the tail, 0: no

// azp_rules.cu
int azp_augment_launch(int board, int mode, int n_max, const int* n_dev, const float* planes, const float* pi,
                       const int32_t* z, float* out_planes, float* out_pi, int32_t* out_z,
                       long long out_offset, const long long* out_offset_dev, long long out_capacity,
                       cudaStream_t stream);

// azp_tree.cu
int azp_tree_begin_move(azp_engine* e, int n, int emit, int compact, int parity, cudaStream_t s);  // reset trees, root result, roots -> eval batch
// `logits` != 0: `policy` holds raw logits and the softmax over all cells happens in the consumer (bf16 net path)
int azp_tree_root_expand(azp_engine* e, int n, const float* policy, int by_slot, int add_noise, int logits, cudaStream_t s);
int azp_tree_select(azp_engine* e, int n, int compact, int parity, cudaStream_t s);
// `value`: float32 [n] or, with value_f64 != 0, float64 [n] (a host evaluator's Python float, mcts.py:137-141)
int azp_tree_expand_backup(azp_engine* e, int n, const float* policy, const void* value, int by_slot, int logits,
                           int value_f64, cudaStream_t s);
int azp_softmax_rows(int n, int A, const float* logits, float* policy, cudaStream_t s);
// expand+backup of the step just evaluated, then the descent of the next step (counter pair `parity`), one launch
int azp_tree_advance(azp_engine* e, int n, const float* policy, const float* value, int by_slot, int logits, int compact,
                     int parity, cudaStream_t s);
// the same two descents with conv1 of the emitted leaves as the kernel's tail (K == 1 only); `first`: the descent of step 0
int azp_tree_descend_conv1(azp_engine* e, int n, bool first, const float* policy, const float* value, int by_slot, int logits,
                           int compact, int parity, const Conv1TailArgs& ct, cudaStream_t s);
int azp_tree_root_visits(azp_engine* e, int n, int32_t* visits, cudaStream_t s);
int azp_tree_root_priors(azp_engine* e, int n, float* priors, cudaStream_t s);
int azp_tree_leaf_states(azp_engine* e, int n, uint64_t* x, uint64_t* o, uint8_t* tm, int16_t* last, uint8_t* kind, cudaStream_t s);
int azp_episode_seat(azp_engine* e, int n, int arena, int next_phase, cudaStream_t s);
int azp_episode_activate(azp_engine* e, int n, int phase, cudaStream_t s);
int azp_episode_end_move(azp_engine* e, int n, int arena, cudaStream_t s);
int azp_episode_finish(azp_engine* e, int n, int arena, int next_phase, cudaStream_t s);
int azp_episode_drain(azp_engine* e, long long* out_written_dev, int mult, cudaStream_t s);

// profiling helpers (azp_engine.cu)
void azp_prof_begin(azp_engine* e, int kind, const int32_t* count_dev, int count_host, cudaStream_t s);
void azp_prof_end(azp_engine* e, int kind, cudaStream_t s);
int azp_prof_collect(azp_engine* e);

// azp_net.cu
int azp_net_create(azp_engine* e);
void azp_net_destroy(azp_engine* e);
int azp_net_load(azp_engine* e, int slot, const float* blob, size_t n, bool on_device, cudaStream_t s);
size_t azp_net_weight_count(const azp_engine* e);
// evaluate the contiguous leaf batch: rows [0,count) (front) or [G-count, G) read backwards
// (back=1); count is read on the device from *count_dev (or count_host if count_dev == null)
int azp_net_eval_leaves(azp_engine* e, int slot, const int32_t* count_dev, int count_host, int back, cudaStream_t s);
int azp_net_eval_planes(azp_engine* e, int slot, int n, const float* planes, float* logits, float* policy,
                        float* value, cudaStream_t s);
