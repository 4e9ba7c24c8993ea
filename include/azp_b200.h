/* azp_b200.h -- C ABI of the B200-native batched self-play engine.
 *
 * This is the drop-in boundary for the ONE hot path of AlphaZero-Piskvorky
 * (SURVEY.md section 8b).  The reference has no FFI layer -- its seams are plain
 * Python call signatures -- so every entry point below names the reference
 * callable it replaces (file:line relative to /root/reference/).  The Python
 * facade in alphazero-piskvorky_b200/ binds these with ctypes and mirrors the
 * reference's class/method names on top.
 *
 * Conventions
 *   - every function returns an int status (AZP_OK == 0); no exceptions or aborts
 *     cross the ABI; azp_last_error() gives the text of the last failure;
 *   - all buffers are caller-allocated DEVICE memory unless the name says host;
 *     the engine owns only its node table and workspaces;
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous w.r.t.
 *     the host unless documented otherwise;
 *   - one engine per device per process; not thread-safe;
 *   - boards are bitboards: bit (r*B + c) of a little-endian array of
 *     AZP_WORDS(B) uint64 words per colour; to_move is 1 (X) or 2 (O);
 *     last is r*B+c of the last move or -1; results are 0 none / 1 X / 2 O / 3 draw.
 */
#ifndef AZP_B200_H
#define AZP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AZP_OK 0
#define AZP_ERR_INVALID_ARG 1
#define AZP_ERR_CUDA 2
#define AZP_ERR_INVALID_MOVE 3   /* analogue of ValueError("Invalid move"), games.py:76-77 */
#define AZP_ERR_CAPACITY 4       /* node table / slab overflow reported by the device flag */
#define AZP_ERR_UNSUPPORTED 5
#define AZP_ERR_NO_WEIGHTS 6
#define AZP_MORE 100             /* driver stopped because the output slab is full; call again */

#define AZP_MAX_BOARD 19
#define AZP_MAX_WORDS 6
#define AZP_WORDS(B) (((B) * (B) + 63) / 64)

#define AZP_NET_STEM 0           /* GomokuNet, alphazero/net.py:28-72 */
#define AZP_NET_TOWER 1          /* residual tower of BASELINE config 4 */
#define AZP_DTYPE_F32 0          /* SIMT fp32 path: <=1e-5 vs the reference net */
#define AZP_DTYPE_BF16 1         /* tcgen05 implicit-GEMM path: <=2e-2 */
#define AZP_AUGMENT_REFERENCE 0  /* 4 outputs exactly as self_play.py:94-108 is written */
#define AZP_AUGMENT_D4 1         /* 8 outputs, consistent dihedral group action */
#define AZP_AUGMENT_NONE 2       /* 1 output (raw positions, used before an all-gather) */

typedef struct azp_engine azp_engine;

typedef struct azp_config {
  int32_t board;            /* BOARD_SIZE, constants.py:2 */
  int32_t win_length;       /* WIN_LENGTH, constants.py:3 */
  int32_t games;            /* concurrent game slots on this GPU */
  int32_t sims;             /* MCTS simulations per move (MCTS.num_simulations, mcts.py:94) */
  double c_puct;            /* mcts.py:95 */
  double dirichlet_alpha;   /* mcts.py:91 */
  double dirichlet_weight;  /* mcts.py:92 */
  int32_t net_kind;         /* AZP_NET_* */
  int32_t net_dtype;        /* AZP_DTYPE_* */
  int32_t tower_blocks;     /* tower only */
  int32_t tower_channels;   /* tower only */
  int32_t augment_mode;     /* AZP_AUGMENT_* */
  int32_t max_plies;        /* 0 = board*board */
  uint64_t seed;            /* Philox key; streams are keyed by GLOBAL game id */
  int32_t rank, world;      /* game sharding: this engine plays global ids == rank (mod world) blocks */
  int32_t device;           /* CUDA device ordinal */
  int32_t leaves_per_step;  /* 0/1 = exact search (one leaf per game per step, visit counts == reference);
                             * K > 1 = throughput mode: K descents per game per step kept apart by virtual loss */
} azp_config;

typedef struct azp_counters {
  uint64_t simulations;     /* iterations of the loop at mcts.py:123-141, all games */
  uint64_t leaf_evals;      /* network evaluations (non-terminal leaves + roots) */
  uint64_t terminal_leaves; /* simulations that ended on a terminal position */
  uint64_t children_scanned;/* sum over select steps of slots scanned */
  uint64_t path_length;     /* sum over simulations of select steps */
  uint64_t expansions;      /* nodes expanded */
  uint64_t games_finished;
  uint64_t plies;           /* moves played */
  uint64_t positions;       /* raw positions emitted */
  uint64_t examples;        /* examples written after augmentation */
  uint64_t kernel_launches; /* launches of this library's kernels */
  uint64_t x_wins, o_wins, draws;
} azp_counters;

int azp_version(void);
int azp_create(const azp_config* cfg, azp_engine** out);
void azp_destroy(azp_engine* e);
const char* azp_last_error(const azp_engine* e);   /* e may be NULL: error of the last failed create */

/* ---- stateless batched rules primitives (each parity-testable alone) -------- */
/* cells: int8 [n][B*B], 0 empty / 1 X / 2 O  <->  bitboards x,o: uint64 [n][AZP_WORDS(B)] */
int azp_pack_boards(int board, int n, const int8_t* cells, uint64_t* x, uint64_t* o, void* stream);
int azp_unpack_boards(int board, int n, const uint64_t* x, const uint64_t* o, int8_t* cells, void* stream);
/* Gomoku.get_legal_actions, games.py:35-47: mask uint8 [n][B*B] of empties (no terminal guard) */
int azp_legal_mask(int board, int n, const uint64_t* x, const uint64_t* o, uint8_t* mask, void* stream);
/* Gomoku.apply_action, games.py:64-82: in place; status[i]=AZP_ERR_INVALID_MOVE and board i
 * untouched when the cell is occupied or off-board */
int azp_apply_moves(int board, int n, uint64_t* x, uint64_t* o, uint8_t* to_move, int16_t* last,
                    const int16_t* moves, int32_t* status, void* stream);
/* Gomoku.is_terminal/_check_line/get_game_result, games.py:133-179, 212-227 (full-board rule
 * incl. overlines and the row-major-first line start on two-winner boards) */
int azp_terminal_status(int board, int win_length, int n, const uint64_t* x, const uint64_t* o,
                        uint8_t* result, void* stream);
/* Gomoku.encode, games.py:86-129: float32 [n][4][B][B] */
int azp_encode(int board, int n, const uint64_t* x, const uint64_t* o, const uint8_t* to_move,
               const int16_t* last, float* planes, void* stream);
/* MCTS.run epilogue, mcts.py:143-170: visit counts int32 [n][B*B] (+ legal mask) and a
 * temperature per row -> pi float32 [n][B*B] */
int azp_extract_policy(int board, int n, const int32_t* visits, const uint8_t* legal,
                       const double* temperature, float* pi, void* stream);
/* SelfPlayManager._augment_symmetries, self_play.py:94-108 (mode REFERENCE) or the 8-fold D4
 * action (mode D4): n positions -> n*mult examples, position-major */
int azp_augment_write(int board, int mode, int n, const float* planes, const float* pi,
                      const int32_t* z, float* out_planes, float* out_pi, int32_t* out_z, void* stream);

/* ReplayBuffer.extend on the device (replay_buffer.py:27-29: append in order, evict the oldest) fused with the
 * augmentation: up to n_max positions (the count is read on the device from *n_dev when n_dev != NULL -- e.g. the
 * header of a slab received by an all-gather -- else it is n_max) are expanded by `mode` and written into a ring of
 * ring_capacity examples at slot (*ring_head_dev + i) % ring_capacity; *ring_head_dev (int64 on the device = examples
 * ever written) advances by count * mult.  A slab larger than the ring keeps its newest ring_capacity examples
 * (deque(maxlen) semantics).  No host synchronisation. */
int azp_augment_write_ring(int board, int mode, int n_max, const int32_t* n_dev, const float* planes, const float* pi,
                           const int32_t* z, float* ring_planes, float* ring_pi, int32_t* ring_z,
                           int64_t* ring_head_dev, int64_t ring_capacity, void* stream);

/* ---- network ------------------------------------------------------------------ */
/* weights: float32 blob on the HOST holding the reference state_dict tensors concatenated in
 * key order (stem: conv1.weight, conv1.bias, conv2.*, conv3.*, policy_conv.*, policy_fc.*,
 * value_conv.*, value_fc1.*, value_fc2.*; net.py:37-53), PyTorch layouts.  slot 0 = the
 * self-play / candidate net, slot 1 = the arena baseline.  Synchronous. */
int azp_load_weights(azp_engine* e, int slot, const float* host_blob, size_t n_floats);
/* same, blob already on this device (after an NCCL broadcast of promoted weights, or a per-step refresh).  On an
 * already loaded bf16 stem the derived buffers are rebuilt IN PLACE by one kernel on `stream` (bit-identical to the host
 * builders; asynchronous, ordered after the engine's earlier work on any stream); first loads, the fp32 path and the
 * tower go through the host and are synchronous. */
int azp_load_weights_device(azp_engine* e, int slot, const float* dev_blob, size_t n_floats, void* stream);
size_t azp_weight_count(const azp_engine* e);
/* GomokuNet.forward + controller.py:33-55: planes float32 [n][4][B][B] ->
 * logits [n][B*B] (raw), policy [n][B*B] (softmax over ALL cells), value [n]; any output may be NULL */
int azp_net_forward(azp_engine* e, int slot, int n, const float* planes, float* logits,
                    float* policy, float* value, void* stream);

/* diagnostics (bf16 path): intermediate activations of the LAST evaluation of slots [0,n) as dense
 * float32: layer 1 -> [n][B*B][32], layer 2 -> [n][B*B][64], layer 3 -> head features [n][6][B*B] */
int azp_debug_activations(azp_engine* e, int layer, int n, float* out, void* stream);

/* ---- search, step by step (lets a host evaluator be injected: the operator-level hook
 *      policy_value_fn of mcts.py:88-93) --------------------------------------------- */
/* MCTS.run prologue, mcts.py:106: new trees rooted at n positions (n <= games) */
int azp_search_begin(azp_engine* e, int n, const uint64_t* x, const uint64_t* o,
                     const uint8_t* to_move, const int16_t* last, const int32_t* ply,
                     const int64_t* game_id, void* stream);
/* mcts.py:109-120: root priors from policy float32 [n][B*B]; Dirichlet mix if add_noise */
int azp_root_expand(azp_engine* e, const float* policy, int add_noise, void* stream);
/* mcts.py:126-129: one PUCT descent per game; leaves are left in the engine */
int azp_select(azp_engine* e, void* stream);
/* the leaves of the last azp_select: kind uint8 [n] (0 = needs evaluation, 1 = terminal,
 * 2 = idle), boards as for azp_encode */
int azp_leaf_states(azp_engine* e, uint64_t* x, uint64_t* o, uint8_t* to_move, int16_t* last,
                    uint8_t* kind, void* stream);
/* mcts.py:131-141: expand the evaluated leaves with policy float32 [n][B*B] / value float32 [n]
 * (indexed by game), terminal leaves use their game result; back up along each path */
int azp_expand_backup(azp_engine* e, const float* policy, const float* value, void* stream);
/* same with value float64 [n]: a host evaluator returns a Python float and the reference backs up exactly that
 * double (value = float(v.item()), controller.py:50; node.backup(-value), mcts.py:137-141) */
int azp_expand_backup_f64(azp_engine* e, const float* policy, const double* value, void* stream);
/* root child visit counts int32 [n][B*B] (the reference computes them at mcts.py:144); after a driver call
 * (azp_selfplay_run / azp_arena_run_steps) these are the counts of the LAST move step's searches, one row per slot */
int azp_root_visits(azp_engine* e, int32_t* visits, void* stream);
/* root priors float32 [n][B*B] as the search holds them after azp_root_expand: policy[r,c] on legal cells,
 * (1-w)*policy + w*Dirichlet(alpha) if add_noise was set (mcts.py:109-120, Node.expand mcts.py:50-64); 0 on
 * occupied cells */
int azp_root_priors(azp_engine* e, float* priors, void* stream);
/* whole search with the engine's own net: begin + root eval + S x (select, net, expand_backup) */
int azp_search(azp_engine* e, int slot, int n, const uint64_t* x, const uint64_t* o,
               const uint8_t* to_move, const int16_t* last, const int32_t* ply,
               const int64_t* game_id, int add_noise, int32_t* visits, void* stream);

/* ---- drivers ------------------------------------------------------------------- */
/* temperature per ply, host doubles (default_temperature_schedule, self_play.py:24-26, is
 * evaluated by the facade so user-supplied schedules keep working) */
int azp_set_temperature_table(azp_engine* e, const double* host_table, int n);
/* augmentation applied by the next azp_selfplay_run calls (AZP_AUGMENT_*; initial value: azp_config.augment_mode).
 * AZP_AUGMENT_NONE emits raw positions -- the form that is all-gathered across ranks, the receiving side augments
 * (self_play.py:94-108 runs in the parent process in the reference too). */
int azp_set_augment_mode(azp_engine* e, int mode);
/* Upper bound on the raw positions one move step of azp_selfplay_run emits (default and maximum: min(games * max_plies,
 * 16384); never below one whole game).  Finished games that do not fit wait -- their slot stays occupied -- and are
 * emitted by a later step, so a small fixed-capacity slab can be exchanged between ranks every step.
 * The output capacity azp_selfplay_run asks for shrinks with it. */
int azp_set_slab_capacity(azp_engine* e, int64_t positions);
/* SelfPlayManager.generate_self_play, self_play.py:110-159 + _worker, :29-77.
 * Plays global game ids [first_game, first_game + n_games) that belong to this rank, S sims per
 * move with root noise, writes augmented examples to out_* (capacity in examples).  Returns
 * AZP_OK when every game is finished, AZP_MORE when out_* ran full (drain, then call again with
 * resume != 0).  *host_written = examples written by this call.  Synchronous at return. */
int azp_selfplay_run(azp_engine* e, int64_t first_game, int64_t n_games, int resume,
                     int max_move_steps, float* out_planes, float* out_pi, int32_t* out_z,
                     int64_t capacity, int64_t* host_written, void* stream);
/* ModelEvaluator.evaluate, evaluator.py:38-122: candidate (slot 0, plays X) vs baseline (slot 1,
 * plays O); game i has X to move first iff i is even; no root noise; T = table[step].
 * host_wld = {candidate wins, baseline wins, draws} for this rank's games.
 * The lockstep is PHASE-ALIGNED: X moves in the even move steps of the match and O in the odd ones in every game (a game
 * whose first mover does not match the phase of the step it is seated for starts one step later), so each step
 * evaluates one leaf batch with one net.  Games are independent and their random streams are keyed by (game id, ply):
 * the match is the one the two-batch schedule plays (AZP_ARENA_ALIGNED=0), with half the network launches. */
int azp_arena_run(azp_engine* e, int64_t first_game, int64_t n_games, int64_t* host_wld, void* stream);
/* the same match in slices of at most max_move_steps lockstep moves (0 = run to the end): returns AZP_MORE while games
 * remain (call again with resume != 0); host_wld is the tally since the match started.  Lets a caller watch the loop
 * of evaluator.py:64-90 move by move (azp_root_visits + azp_episode_snapshot between slices). */
int azp_arena_run_steps(azp_engine* e, int64_t first_game, int64_t n_games, int resume, int max_move_steps,
                        int64_t* host_wld, void* stream);
/* Episode state of every slot as the drivers hold it (self_play.py:47-58 / evaluator.py:64-90 bookkeeping), copied
 * device-to-device; any pointer may be NULL.  ply int32 [games] = moves played in the slot's current game; game_id
 * int64 [games]; status uint8 [games] (0 idle / 1 playing / 2 finished, not yet emitted / 3 seated, waiting for its
 * phase of the aligned arena); moves uint16 [games][max_plies]
 * and pi_hist float32 [games][max_plies][B*B] = the move played and the pi it was drawn from at every ply of the slot's
 * game (rows of a finished game stay readable until the slot's next game overwrites them). */
int azp_episode_snapshot(azp_engine* e, int32_t* ply, int64_t* game_id, uint8_t* status, uint16_t* moves,
                         float* pi_hist, void* stream);
int azp_get_counters(azp_engine* e, azp_counters* host_out);
/* Per-kernel timing for the roofline: when enabled, every launch of the DOMINANT kernel (the
 * largest convolution of the net) is bracketed by CUDA events on its own stream.
 * azp_get_profile synchronises, returns the summed duration (ms), the number of launches and the
 * algorithmic FLOPs those launches performed, and resets the accumulators. */
int azp_set_profiling(azp_engine* e, int enabled);
int azp_get_profile(azp_engine* e, double* host_ms, uint64_t* host_launches, double* host_flops);
/* The other kernels of the path, for the HBM side of the roofline.  While profiling is on, the
 * last 40 simulation steps of every move run as plain launches with each kernel between its own pair of CUDA events
 * (the steps before them still replay the captured graph, unbracketed).  Returns and resets, per kernel class: summed
 * duration (ms), bracketed launches, and the units those launches processed (leaves of the evaluated batch for the net
 * kernels, game slots for the tree / episode kernels). */
#define AZP_PROF_DOMINANT 0     /* largest convolution of the net (conv3 / a tower conv): GomokuNet.forward, net.py:57-59 */
#define AZP_PROF_ADVANCE 1      /* expand + backup + next descent: Node.expand/backup/select, mcts.py:50-82, 126-141 */
#define AZP_PROF_CONV1 2        /* leaf gather + encode + first conv: Gomoku.encode, games.py:86-129; net.py:57 */
#define AZP_PROF_CONV2 3        /* second convolution, net.py:58 */
#define AZP_PROF_FC 4           /* policy_fc / value_fc1 / value_fc2, net.py:64-70 */
#define AZP_PROF_ROOT_EXPAND 5  /* root priors + Dirichlet noise, mcts.py:109-120 */
#define AZP_PROF_END_MOVE 6     /* pi from visits, move sampling, apply: mcts.py:143-183, self_play.py:53-58 */
#define AZP_PROF_FINISH 7       /* labelled raw positions of finished games: self_play.py:60-72 */
#define AZP_PROF_AUGMENT 8      /* symmetry augmentation + example writer: self_play.py:94-108 */
#define AZP_PROF_KINDS 9
int azp_get_profile_kernel(azp_engine* e, int kind, double* host_ms, uint64_t* host_launches, double* host_units);
int azp_reset_counters(azp_engine* e);

/* ---- training step (the immediate consumer of the examples) ----------- */
/* Stem net only (AZP_NET_STEM).  The trainer object owns activation / im2col / split-K workspaces for
 * batches of up to max_batch positions; parameters, gradients and optimizer moments are caller-owned
 * float32 device blobs in the azp_load_weights layout (reference state_dict order, PyTorch layouts). */
typedef struct azp_trainer azp_trainer;
int azp_trainer_create(int board, int max_batch, int device, azp_trainer** out);
void azp_trainer_destroy(azp_trainer* t);
const char* azp_trainer_last_error(const azp_trainer* t);     /* t may be NULL */
size_t azp_trainer_weight_count(const azp_trainer* t);
/* NeuralNetworkController.train_step up to loss.backward(), controller.py:117-124: forward of
 * GomokuNet (net.py:55-72) on planes float32 [n][4][B][B], policy loss = F.cross_entropy(logits, pi)
 * with probability targets pi float32 [n][B*B], value loss = F.mse_loss(value, z) with z float32 [n],
 * gradients of (policy + value loss) w.r.t. every parameter -> grads (same layout as params; batch-mean
 * reduced, deterministic).  losses float32 [4] on the device = {policy loss, value loss, mean entropy of
 * the predicted policy, policy + value}.  grads may be NULL (forward + losses only). */
int azp_train_loss_grad(azp_trainer* t, int n, const float* params, const float* planes, const float* pi,
                        const float* z, float* grads, float* losses, void* stream);
/* torch.optim.AdamW.step() (controller.py:75-79, 114-115, 125) on flat blobs, stateless: `step` is the
 * 1-based step count of the bias corrections; the reference uses betas (0.9, 0.999), eps 1e-8,
 * weight_decay 1e-4 and sets lr before every step.  In place on params / exp_avg / exp_avg_sq. */
int azp_adamw_step(size_t n, float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int step,
                   double lr, double beta1, double beta2, double eps, double weight_decay, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AZP_B200_H */
