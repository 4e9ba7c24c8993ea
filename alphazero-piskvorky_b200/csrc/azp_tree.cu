// Node-table kernels: PUCT selection, expansion, backup, root noise, move selection and the
// (state, pi, z) position writer.  One warp per game everywhere (SURVEY.md section 8 rows
// a6-a12, a16).  All PUCT arithmetic is float64 in the reference's operation order, so visit
// counts are identical to alphazero/mcts.py whenever the evaluator numbers are identical.
#include "azp_internal.cuh"
#include "azp_conv1_tail.cuh"

#define SLOT_IDLE 0
#define SLOT_PLAYING 1
#define SLOT_PENDING 2     // game over, positions not yet emitted (raw slab was full)
#define SLOT_DELAYED 3     // phase-aligned arena: seated, waits for the move step whose phase matches its first mover

__device__ __forceinline__ void count(const TreeView& tv, int idx, unsigned long long v) {
  if (v) atomicAdd(tv.counters + idx, v);
}

// Claim a row of the contiguous evaluation batch and store the position there.  Games evaluated
// by weight slot 0 fill rows from the front, slot 1 (arena baseline) from the back.
// `rec` = leaf record index (g*K + k).
template <int NW>
__device__ __forceinline__ int emit_leaf(const TreeView& tv, int g, int rec, const BB<NW>& x, const BB<NW>& o, int tm,
                                         int last, int compact, int parity, int lane) {
  int slot = rec;
  if (compact) {
    if (lane == 0) {
      int side = tv.net_of_game[g];
      int idx = atomicAdd(&tv.n_leaves[parity * 2 + side], 1);
      slot = side ? tv.slots - 1 - idx : idx;
    }
    slot = __shfl_sync(AZP_FULL, slot, 0);
  }
  if (lane == 0) {
    tv.leaf_slot[rec] = slot;
    tv.slot_game[slot] = g;
    bb_store<NW>(tv.leaf_x + (size_t)slot * NW, x);
    bb_store<NW>(tv.leaf_o + (size_t)slot * NW, o);
    tv.leaf_tm[slot] = (uint8_t)tm;
    tv.leaf_last[slot] = (int16_t)last;
  }
  return slot;
}

// ---- MCTS.run prologue (mcts.py:106): fresh tree per move --------------------------------
template <int NW>
__global__ void begin_move_kernel(RulesParams rp, TreeView tv, int n, int emit, int compact, int parity) {
  int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    tv.n_leaves[(parity ^ 1) * 2 + 0] = 0;
    tv.n_leaves[(parity ^ 1) * 2 + 1] = 0;
  }
  if (g >= n) return;
  const int rec0 = g * tv.K;
  if (lane < tv.K) {                      // every descent record of this game starts idle
    tv.leaf_kind[rec0 + lane] = 2; tv.leaf_slot[rec0 + lane] = -1; tv.path_len[rec0 + lane] = 0;
  }
  __syncwarp();
  if (!tv.active[g]) return;
  BB<NW> x = bb_load<NW>(tv.root_x + (size_t)g * NW), o = bb_load<NW>(tv.root_o + (size_t)g * NW);
  int res = warp_result<NW>(rp, x, o, lane);
  if (lane == 0) {
    tv.node_count[g] = 1;
    tv.root_n[g] = 0;
    tv.root_result[g] = (uint8_t)res;
    tv.leaf_kind[rec0] = 0;
  }
  // the root is evaluated unconditionally (mcts.py:109), even when it is terminal
  if (emit) {
    emit_leaf<NW>(tv, g, rec0, x, o, tv.root_tm[g], tv.root_last[g], compact, parity, lane);
    if (lane == 0) count(tv, CNT_EVALS, 1);
  }
}

// ---- root priors (+ Dirichlet noise), mcts.py:109-120 ---------------------------------------
template <int NW>
__global__ void root_expand_kernel(RulesParams rp, TreeView tv, int n, const float* __restrict__ policy,
                                   int by_slot, int add_noise, double alpha, double weight, uint64_t seed, int logits) {
  int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (g >= n || !tv.active[g]) return;
  const int A = rp.A;
  BB<NW> x = bb_load<NW>(tv.root_x + (size_t)g * NW), o = bb_load<NW>(tv.root_o + (size_t)g * NW);
  BB<NW> occ;
#pragma unroll
  for (int k = 0; k < NW; k++) occ.w[k] = x.w[k] | o.w[k];
  const float* prow = policy + (size_t)(by_slot ? tv.leaf_slot[g * tv.K] : g) * A;
  float sm_m = 0.f, sm_inv = 1.f;               // `logits`: the row holds raw logits, softmax happens here
  if (logits) warp_softmax_stats(prow, A, lane, sm_m, sm_inv);
  const size_t base = (size_t)g * tv.cap * A;   // node 0
  double gsum = 0.0;
  Philox rng = philox_for_game(seed, tv.game_id[g]);
  const uint32_t ply = (uint32_t)tv.ply[g];
  if (add_noise) {
    for (int a = lane; a < A; a += 32)
      if (!bb_test<NW>(occ, a)) gsum += gamma_variate(rng, alpha, (uint32_t)a, ply);
    gsum = warp_sum_d(gsum);
  }
  for (int a = lane; a < A; a += 32) {
    bool legal = !bb_test<NW>(occ, a);
    float p = 0.f;
    if (legal) {
      p = logits ? __expf(prow[a] - sm_m) * sm_inv : prow[a];
      if (add_noise) {
        // policy[r,c] = (1-w)*policy[r,c] + w*eta with a float32 policy array (mcts.py:116):
        // float32 product, float64 sum, rounded back to float32 on store
        double eta = gamma_variate(rng, alpha, (uint32_t)a, ply) / gsum;
        float keep = __fmul_rn((float)(1.0 - weight), p);
        p = (float)__dadd_rn((double)keep, __dmul_rn(weight, eta));
      }
    }
    tv.P[base + a] = p;
    tv.N[base + a] = 0;
    tv.W[base + a] = 0.0;
    tv.child[base + a] = -1;
  }
  if (lane == 0) count(tv, CNT_EXPANSIONS, 1);
}

#define CHILD_NONE (-1)
#define CHILD_PENDING (-2)      // claimed by an earlier descent of this step (throughput mode only)

// ---- one node's PUCT scan (mcts.py:66-74) ----------------------------------------------------------------------
// score = Q + c_puct * P * sqrt(N_parent + 1e-8) / (1 + N_child), float64, left to right; argmax with the lowest cell winning
// ties; `ex` (or -1): a cell left out of the scan (root_prescan).  Returns the winner's N / child pointer too.
template <int NW>
__device__ __forceinline__ void scan_node(const TreeView& tv, size_t nb, int A, const uint64_t (&occw)[NW], int nparent,
                                          double c_puct, int lane, int ex, double& best, int& besta, int& cn, int& ch) {
  const double sq = __dsqrt_rn(__dadd_rn((double)nparent, 1e-8));
  best = -1.0e308;
  besta = 0x7fffffff;
  // all of a node's edge statistics are fetched before any arithmetic: one memory round trip per level
  constexpr int MAXI = (2 * NW < (AZP_MAX_BOARD * AZP_MAX_BOARD + 31) / 32) ? 2 * NW   // A <= 64*NW cells = 2*NW per lane
                                                                         : (AZP_MAX_BOARD * AZP_MAX_BOARD + 31) / 32;
  int nn_[MAXI], ch_[MAXI];
  double ww_[MAXI];
  float pp_[MAXI];
#pragma unroll
  for (int i = 0; i < MAXI; i++) {
    const int a = lane + 32 * i;
    nn_[i] = 0; ww_[i] = 0.0; pp_[i] = 0.f; ch_[i] = CHILD_NONE;
    if (a < A) { nn_[i] = tv.N[nb + a]; ww_[i] = tv.W[nb + a]; pp_[i] = tv.P[nb + a]; ch_[i] = tv.child[nb + a]; }
  }
#pragma unroll
  for (int i = 0; i < MAXI; i++) {
    const int a = lane + 32 * i;
    const bool occupied = a >= A || a == ex || ((occw[i >> 1] >> (lane + 32 * (i & 1))) & 1ull);
    if (occupied) continue;
    const int nn = nn_[i];
    // unvisited children (the vast majority of slots) need no fp64 division: Q = 0 and x / 1.0 == x exactly
    const double num = __dmul_rn(__dmul_rn(c_puct, (double)pp_[i]), sq);
    double q = 0.0, u = num;
    if (nn) {
      q = __ddiv_rn(ww_[i], (double)nn);
      u = __ddiv_rn(num, (double)(1 + nn));
    }
    const double s = nn ? __dadd_rn(q, u) : u;   // 0.0 + u == u exactly (u >= 0)
    if (s > best) { best = s; besta = a; }    // ascending a: the FIRST maximum is kept
  }
#pragma unroll
  for (int off = 16; off; off >>= 1) {       // argmax with lowest-cell tie-break
    double ob = __shfl_xor_sync(AZP_FULL, best, off);
    int oa = __shfl_xor_sync(AZP_FULL, besta, off);
    if (ob > best || (ob == best && oa < besta)) { best = ob; besta = oa; }
  }
  // the winning edge's N / child pointer are already in the owning lane's registers
  cn = 0; ch = CHILD_NONE;
#pragma unroll
  for (int i = 0; i < MAXI; i++) if (i == (besta >> 5)) { cn = nn_[i]; ch = ch_[i]; }
  cn = __shfl_sync(AZP_FULL, cn, besta & 31);
  ch = __shfl_sync(AZP_FULL, ch, besta & 31);
}
struct RootScan { double best; int besta, cn, ch; };      // argmax over the root's children (see root_prescan)

// ---- K PUCT descents per game (mcts.py:126-129 + 66-74) ---------------------------------------
// K == 1 is the exact mode: the tree is only read here.  K > 1 (throughput mode): after each descent its
// path takes a VIRTUAL LOSS (N += 1, W -= 1, undone by expand_backup) so the next descent of the same step
// goes elsewhere; an unexpanded edge reached twice in one step is a collision and the later descent idles.
template <int NW>
__device__ __forceinline__ void select_body(const RulesParams& rp, const TreeView& tv, int g, int lane, double c_puct,
                                            int compact, int parity, int* slot_out = nullptr, uint64_t* leaf_out = nullptr,
                                            const RootScan* rs = nullptr) {
  const int K = tv.K;
  const int rec0 = g * K;
  if (!tv.active[g]) {
    if (lane < K) { tv.leaf_kind[rec0 + lane] = 2; tv.leaf_slot[rec0 + lane] = -1; tv.path_len[rec0 + lane] = 0; }
    return;
  }
  const int A = rp.A;
  const BB<NW> rx = bb_load<NW>(tv.root_x + (size_t)g * NW), ro = bb_load<NW>(tv.root_o + (size_t)g * NW);
  const int root_tm = tv.root_tm[g], root_last = tv.root_last[g];
  const size_t gbase = (size_t)g * tv.cap * A;
  const int root_result = tv.root_result[g];
  int root_n = tv.root_n[g];
  unsigned long long scanned = 0, path_total = 0, n_eval = 0, n_term = 0;
  for (int kd = 0; kd < K; kd++) {
    const int rec = rec0 + kd;
    int32_t* path = tv.path + (size_t)rec * tv.maxd;
    if (K > 1 && root_n >= tv.S) {       // this move's simulation budget is used up
      if (lane == 0) { tv.leaf_kind[rec] = 2; tv.leaf_slot[rec] = -1; tv.path_len[rec] = 0; }
      continue;
    }
    BB<NW> x = rx, o = ro;
    int tm = root_tm, last = root_last;
    int node = 0, nparent = root_n, depth = 0;
    int result = root_result;
    int kind;
    float tvalue = 0.f;
    while (true) {
      if (result != 0) {   // terminal position: value from its side to move's view (mcts.py:132-134)
        kind = 1;
        tvalue = result == 3 ? 0.f : (result == tm ? 1.f : -1.f);
        break;
      }
      const size_t nb = gbase + (size_t)node * A;
      int stones = 0;
      uint64_t occw[NW];                         // occupancy: cell a = lane + 32i is bit lane + 32(i&1) of word i>>1
#pragma unroll
      for (int k = 0; k < NW; k++) { occw[k] = x.w[k] | o.w[k]; stones += __popcll(occw[k]); }
      double best;
      int besta, cn, ch;
      if (rs && node == 0 && depth == 0) {       // the root's scan was done before the dependency wait (root_prescan)
        best = rs->best; besta = rs->besta; cn = rs->cn; ch = rs->ch;
      } else {
        scan_node<NW>(tv, nb, A, occw, nparent, c_puct, lane, -1, best, besta, cn, ch);
      }
      scanned += (unsigned long long)(A - stones);
      const int a = besta;
      if (a >= A || depth >= tv.maxd) {   // cannot happen on a non-terminal position; never loop forever
        if (lane == 0) atomicExch(tv.error_flag, AZP_ERR_CAPACITY);
        kind = 2;
        break;
      }
      if (lane == 0) path[depth] = node * A + a;
      depth++;
      if (tm == 1) bb_set<NW>(x, a); else bb_set<NW>(o, a);
      if (rp.WL <= 8) {
        // the parent was not terminal: only the mover can have completed a line, and only through `a`
        const bool won = tm == 1 ? warp_line_through<NW>(rp, x, a, lane) : warp_line_through<NW>(rp, o, a, lane);
        result = won ? tm : (stones + 1 == A ? 3 : 0);
      } else {
        result = warp_result<NW>(rp, x, o, lane);
      }
      tm = 3 - tm;
      last = a;
      if (result != 0) continue;
      if (ch == CHILD_PENDING) { kind = 2; break; }          // collision with an earlier descent of this step
      if (ch < 0) {                                          // reached a child that has not been expanded
        kind = 0;
        if (K > 1 && lane == 0) tv.child[nb + a] = CHILD_PENDING;
        break;
      }
      node = ch;
      nparent = cn;
    }
    if (kind == 0) {
      const int slot = emit_leaf<NW>(tv, g, rec, x, o, tm, last, compact, parity, lane);
      if (slot_out) *slot_out = slot;                 // K == 1 callers: the row of this game's leaf in the evaluation batch
      if (leaf_out && lane == 0) {                    // ... and the leaf itself (x words, o words, to-move | last << 8), e.g. in shared memory
#pragma unroll
        for (int k = 0; k < NW; k++) { leaf_out[k] = x.w[k]; leaf_out[NW + k] = o.w[k]; }
        leaf_out[2 * NW] = (uint64_t)(uint32_t)(tm | (last * 256));
      }
    }
    if (lane == 0) {
      tv.path_len[rec] = kind == 2 ? 0 : depth;
      tv.leaf_kind[rec] = (uint8_t)kind;
      tv.leaf_tvalue[rec] = tvalue;
      if (kind != 0) tv.leaf_slot[rec] = -1;
    }
    if (kind != 2) {
      path_total += (unsigned long long)depth;
      n_eval += kind == 0;
      n_term += kind == 1;
      if (K > 1) {
        // virtual loss along the path; visible to this warp's next descent after the __syncwarp
        __syncwarp();
        for (int i = lane; i < depth; i += 32) {
          const size_t e = gbase + path[i];
          tv.N[e] += 1;
          tv.W[e] -= 1.0;
        }
        root_n++;
        __syncwarp();
      }
    }
  }
  if (lane == 0) {
    if (K > 1) tv.root_n[g] = root_n;
    count(tv, CNT_SCANNED, scanned);
    count(tv, CNT_PATH, path_total);
    count(tv, CNT_EVALS, n_eval);
    count(tv, CNT_TERMINAL, n_term);
  }
}

// ---- expand the evaluated leaves and back the values up (mcts.py:136-141, 50-64, 76-82) --------
// returns false on a capacity overflow (error flag set)
// Everything expand+backup needs that the NETWORK kernels do not write -- the leaf record of the previous descent, its
// bitboards, the path -- can be loaded before griddepcontrol.wait (exact mode, K == 1): two dependent memory round trips
// (record -> slot -> bitboards) leave the critical path of the step.
template <int NW>
struct ExpandPre {
  int active, kind, depth, slot, nid, edge;      // edge = path[lane] (lane < depth)
  int en; double ew;                             // that edge's N and W: the backup's read-modify-write starts from registers
  BB<NW> lx, lo;
};
template <int NW>
__device__ __forceinline__ ExpandPre<NW> expand_preload(const TreeView& tv, int A, int g, int lane) {
  ExpandPre<NW> p;
  p.kind = 2; p.depth = 0; p.slot = -1; p.nid = 0; p.edge = 0; p.en = 0; p.ew = 0.0;
  p.active = tv.active[g];
  if (!p.active) return p;
  p.kind = tv.leaf_kind[g];
  p.depth = tv.path_len[g];
  p.slot = tv.leaf_slot[g];
  p.nid = tv.node_count[g];
  if (lane < p.depth) {
    p.edge = tv.path[(size_t)g * tv.maxd + lane];
    const size_t e = (size_t)g * tv.cap * (size_t)A + (size_t)p.edge;
    p.en = tv.N[e]; p.ew = tv.W[e];
  }
  if (p.kind == 0 && p.slot >= 0) {
    p.lx = bb_load<NW>(tv.leaf_x + (size_t)p.slot * NW);
    p.lo = bb_load<NW>(tv.leaf_o + (size_t)p.slot * NW);
  }
  return p;
}

template <int NW>
__device__ __forceinline__ bool expand_backup_body(const RulesParams& rp, const TreeView& tv, int g, int lane,
                                                   const float* __restrict__ policy, const void* __restrict__ value,
                                                   int by_slot, int logits, int value_f64 = 0, ExpandPre<NW>* pre = nullptr) {
  if (!(pre ? pre->active : tv.active[g])) return true;
  const int A = rp.A, K = tv.K;
  const size_t gbase = (size_t)g * tv.cap * A;
  unsigned long long n_exp = 0, n_sims = 0;
  for (int kd = 0; kd < K; kd++) {
    const int rec = g * K + kd;
    const int kind = pre ? pre->kind : tv.leaf_kind[rec];
    if (kind == 2) continue;
    const int32_t* path = tv.path + (size_t)rec * tv.maxd;
    const int depth = pre ? pre->depth : tv.path_len[rec];
    double v;                                     // the reference backs up a Python float (float64), mcts.py:137-141
    if (kind == 0) {
      const int slot = pre ? pre->slot : tv.leaf_slot[rec];
      const size_t row = (size_t)(by_slot ? slot : rec);
      v = value_f64 ? static_cast<const double*>(value)[row] : (double)static_cast<const float*>(value)[row];
      int nid = pre ? pre->nid : tv.node_count[g];
      if (nid >= tv.cap) {
        if (lane == 0) atomicExch(tv.error_flag, AZP_ERR_CAPACITY);
        return false;
      }
      BB<NW> lx, lo;
      if (pre) { lx = pre->lx; lo = pre->lo; }
      else { lx = bb_load<NW>(tv.leaf_x + (size_t)slot * NW); lo = bb_load<NW>(tv.leaf_o + (size_t)slot * NW); }
      const size_t nb = gbase + (size_t)nid * A;
      const float* prow = policy + row * A;
      float sm_m = 0.f, sm_inv = 1.f;             // `logits`: raw logits in, softmax over all cells here (controller.py:49)
      if (logits) warp_softmax_stats(prow, A, lane, sm_m, sm_inv);
#pragma unroll
      for (int i = 0; i < 2 * NW; i++) {          // cell a = lane + 32i is bit lane + 32(i&1) of word i>>1
        const int a = lane + 32 * i;
        if (a < A) {
          const bool legal = !(((lx.w[i >> 1] | lo.w[i >> 1]) >> (lane + 32 * (i & 1))) & 1ull);
          const float pa = logits ? __expf(prow[a] - sm_m) * sm_inv : prow[a];
          tv.P[nb + a] = legal ? pa : 0.f;        // prior = policy[r,c], not renormalised (mcts.py:63)
          tv.N[nb + a] = 0;
          tv.W[nb + a] = 0.0;
          tv.child[nb + a] = CHILD_NONE;
        }
      }
      __syncwarp();
      if (lane == 0) {
        tv.node_count[g] = nid + 1;
        if (depth > 0) tv.child[gbase + path[depth - 1]] = nid;
      }
      __syncwarp();
      n_exp++;
    } else {
      v = (double)tv.leaf_tvalue[rec];
    }
    // node.backup(-value): the leaf's own statistics (last edge) take -v, signs alternate upwards.
    // Throughput mode: the edge already carries this descent's virtual visit (N+1, W-1): swap it for the value.
    for (int i = lane; i < depth; i += 32) {
      const size_t e = gbase + ((pre && i < 32) ? pre->edge : path[i]);
      const double val = ((depth - 1 - i) & 1) ? v : -v;
      if (K == 1) {
        const bool have = pre && i < 32;
        const int n_new = (have ? pre->en : tv.N[e]) + 1;
        const double w_new = __dadd_rn(have ? pre->ew : tv.W[e], val);
        tv.N[e] = n_new;
        tv.W[e] = w_new;
        if (have) { pre->en = n_new; pre->ew = w_new; }     // lane 0: the ROOT edge's new statistics (root_merge)
      } else {
        tv.W[e] += 1.0 + val;
      }
    }
    __syncwarp();
    n_sims++;
  }
  if (lane == 0) {
    if (K == 1) tv.root_n[g] += (int)n_sims;
    count(tv, CNT_SIMS, n_sims);
    count(tv, CNT_EXPANSIONS, n_exp);
  }
  return true;
}

// ---- the root's PUCT scan, before the dependency wait (exact mode, K == 1) -------------------------------------------
// The descent of step i+1 starts with a scan of the root's children.  The backup of step i changes exactly ONE of them (the
// root edge of the pending path: N + 1, W +- v, and its child pointer if the new node hangs right under the root), and the
// root's own visit count becomes root_n + 1 whatever the network says.  So the scan of every OTHER child runs before
// griddepcontrol.wait with the post-backup parent count, and after the backup the one changed child is scored with the same
// arithmetic and merged in with the same tie-break (lowest cell).  Decisions are identical to scanning after the backup
// (parity tests green with it on).  OPT-IN (AZP_ROOT_PRESCAN=1): measured 15.19 M against 15.54 M simulations/s at config 3 --
// unlike the handful of loads of expand_preload, a whole scan before the wait is not hidden: on the 80 SMs the FC kernel
// occupies, the descent's CTA only becomes resident when FC leaves, so its pre-wait work IS the critical path there.
template <int NW>
struct RootPre {
  int have, ex, ex_ch, np;       // ex = root cell of the pending path (-1: none), its child pointer before the expansion, parent count
  float ex_p;
  RootScan rs;
};
template <int NW>
__device__ __forceinline__ RootPre<NW> root_prescan(const RulesParams& rp, const TreeView& tv, int g, int lane, double c_puct,
                                                    const ExpandPre<NW>& pre) {
  RootPre<NW> r;
  r.have = 0; r.ex = -1; r.ex_ch = CHILD_NONE; r.np = 0; r.ex_p = 0.f;
  if (!pre.active || tv.root_result[g] != 0) return r;
  const int A = rp.A;
  const size_t gbase = (size_t)g * tv.cap * A;
  const BB<NW> rx = bb_load<NW>(tv.root_x + (size_t)g * NW), ro = bb_load<NW>(tv.root_o + (size_t)g * NW);
  uint64_t occw[NW];
#pragma unroll
  for (int k = 0; k < NW; k++) occw[k] = rx.w[k] | ro.w[k];
  const bool pending = pre.kind != 2;                                 // a simulation is about to be backed up
  const int e0 = __shfl_sync(AZP_FULL, pre.edge, 0);                  // path[0] = the root cell (node 0)
  r.ex = (pending && pre.depth >= 1) ? e0 : -1;
  r.np = tv.root_n[g] + (pending ? 1 : 0);
  if (r.ex >= 0) { r.ex_p = tv.P[gbase + r.ex]; r.ex_ch = tv.child[gbase + r.ex]; }
  scan_node<NW>(tv, gbase, A, occw, r.np, c_puct, lane, r.ex, r.rs.best, r.rs.besta, r.rs.cn, r.rs.ch);
  r.have = 1;
  return r;
}
// after expand_backup_body: score the one changed root child and merge it into the pre-scan
template <int NW>
__device__ __forceinline__ void root_merge(RootPre<NW>& r, const ExpandPre<NW>& pre, double c_puct) {
  if (!r.have || r.ex < 0) return;
  const int n0 = __shfl_sync(AZP_FULL, pre.en, 0);
  const double w0 = __shfl_sync(AZP_FULL, pre.ew, 0);
  const double sq = __dsqrt_rn(__dadd_rn((double)r.np, 1e-8));
  const double num = __dmul_rn(__dmul_rn(c_puct, (double)r.ex_p), sq);
  const double s = __dadd_rn(__ddiv_rn(w0, (double)n0), __ddiv_rn(num, (double)(1 + n0)));      // n0 >= 1
  const int chx = (pre.kind == 0 && pre.depth == 1) ? pre.nid : r.ex_ch;    // the new node hangs under this very edge
  if (s > r.rs.best || (s == r.rs.best && r.ex < r.rs.besta)) { r.rs.best = s; r.rs.besta = r.ex; r.rs.cn = n0; r.rs.ch = chx; }
}

// ---- kernels: the two halves alone (step-by-step API) and fused ------------------------------------------------
// 2048 games = 512 CTAs of four warps: at >128 registers only 3 CTAs fit an SM (444 per wave) and the latency-bound
// descent ran as TWO waves; capped so that one wave holds them all (4 CTAs x 148 SMs = 592)
// 9x9 boards (NW <= 2) run with up to 4096 concurrent games: 1024 CTAs of four warps need SEVEN resident CTAs per SM to
// stay one wave (the latency-bound descent does not care about the few spilled registers, it cares about waves)
template <int NW>
__global__ void __launch_bounds__(128, NW <= 2 ? 7 : 4) select_kernel(RulesParams rp, TreeView tv, int n, double c_puct, int compact, int parity) {
  pdl_wait();
  pdl_launch_dependents();
  int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    tv.n_leaves[(parity ^ 1) * 2 + 0] = 0;
    tv.n_leaves[(parity ^ 1) * 2 + 1] = 0;
  }
  if (g >= n) return;
  select_body<NW>(rp, tv, g, lane, c_puct, compact, parity);
}

template <int NW>
__global__ void expand_backup_kernel(RulesParams rp, TreeView tv, int n, const float* __restrict__ policy,
                                     const void* __restrict__ value, int by_slot, int logits, int value_f64) {
  pdl_wait();
  pdl_launch_dependents();
  int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (g >= n) return;
  expand_backup_body<NW>(rp, tv, g, lane, policy, value, by_slot, logits, value_f64);
}

// The descent starts with the ROOT's edge arrays (20 B x A, evicted from L2 by ~100 MB of activations per step): ask L2 for
// them ahead of use.  They are tree data only this kernel family writes, so the request may even be issued BEFORE
// griddepcontrol.wait (`early`, AZP_ROOT_PREFETCH=2), while the FC kernel is still running.
template <int NW>
__device__ __forceinline__ void prefetch_root(const RulesParams& rp, const TreeView& tv, int g, int lane) {
  const size_t gb = (size_t)g * tv.cap * rp.A;
  const char* pn = reinterpret_cast<const char*>(tv.N + gb);
  const char* pw = reinterpret_cast<const char*>(tv.W + gb);
  const char* pp = reinterpret_cast<const char*>(tv.P + gb);
  const char* pc = reinterpret_cast<const char*>(tv.child + gb);
  for (int o = lane * 128; o < rp.A * 8; o += 32 * 128) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(pw + o));
    if (o < rp.A * 4) {
      asm volatile("prefetch.global.L2 [%0];" ::"l"(pn + o));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(pp + o));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(pc + o));
    }
  }
}

// expand+backup of simulation step i followed at once by the descent of step i+1: the same warp owns the game in
// both halves, so the two launches of the lockstep loop collapse into one (the tree lines just written are still
// in L1/L2 when the descent re-reads them)
template <int NW>
__global__ void __launch_bounds__(128, NW <= 2 ? 7 : 4) advance_kernel(RulesParams rp, TreeView tv, int n, const float* __restrict__ policy,
                               const float* __restrict__ value, int by_slot, int logits, double c_puct, int compact,
                               int parity) {
  const unsigned tl_idx = (threadIdx.x & 31) == 0 ? AZP_TL_ENTRY(AZP_TL_ADVANCE, gridDim.x * (blockDim.x >> 5)) : 0u;
  int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (tv.K == 1 && ((compact >> 5) & 1) && g < n) prefetch_root<NW>(rp, tv, g, lane);      // bit 5: before the wait (see prefetch_root)
  ExpandPre<NW> pre;
  pre.active = 0;
  const bool use_pre = tv.K == 1 && !((compact >> 6) & 1) && g < n;                       // bit 6: AZP_EXPAND_PRELOAD=0
  RootPre<NW> rpre;
  rpre.have = 0;
  if (use_pre) {
    pre = expand_preload<NW>(tv, rp.A, g, lane);
    if (!((compact >> 7) & 1)) rpre = root_prescan<NW>(rp, tv, g, lane, c_puct, pre);      // bit 7: AZP_ROOT_PRESCAN=0
  }
  pdl_wait();
  pdl_launch_dependents();
  if ((threadIdx.x & 31) == 0) AZP_TL_START(AZP_TL_ADVANCE, tl_idx);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    tv.n_leaves[(parity ^ 1) * 2 + 0] = 0;
    tv.n_leaves[(parity ^ 1) * 2 + 1] = 0;
  }
  if (g >= n) { if (lane == 0) AZP_TL_END(AZP_TL_ADVANCE, tl_idx, gridDim.x * (blockDim.x >> 5)); return; }
  const bool root_prefetch = ((compact >> 4) & 1) && !((compact >> 5) & 1);      // bit 4 of `compact`: AZP_ROOT_PREFETCH
  compact &= 15;
  if (tv.K == 1 && root_prefetch) {
    // The descent below starts with the ROOT's edge arrays (20 B x A, evicted from L2 by ~100 MB of activations per
    // step): ask L2 for them now, so that they arrive while this warp expands and backs up the previous leaf.
    const size_t gb = (size_t)g * tv.cap * rp.A;
    const char* pn = reinterpret_cast<const char*>(tv.N + gb);
    const char* pw = reinterpret_cast<const char*>(tv.W + gb);
    const char* pp = reinterpret_cast<const char*>(tv.P + gb);
    const char* pc = reinterpret_cast<const char*>(tv.child + gb);
    for (int o = lane * 128; o < rp.A * 8; o += 32 * 128) {
      asm volatile("prefetch.global.L2 [%0];" ::"l"(pw + o));
      if (o < rp.A * 4) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(pn + o));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(pp + o));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(pc + o));
      }
    }
  }
  if (expand_backup_body<NW>(rp, tv, g, lane, policy, value, by_slot, logits, 0, use_pre ? &pre : nullptr)) {
    __syncwarp();             // the descent's lanes read edges other lanes of this warp have just written
    root_merge<NW>(rpre, pre, c_puct);
    select_body<NW>(rp, tv, g, lane, c_puct, compact, parity, nullptr, nullptr, rpre.have ? &rpre.rs : nullptr);
  }
  if (lane == 0) AZP_TL_END(AZP_TL_ADVANCE, tl_idx, gridDim.x * (blockDim.x >> 5));
}

// ---- descent kernels with the conv1 tail (exact mode, bf16 stem; azp_conv1_tail.cuh) ----------------------------------
// FIRST = the descent of step 0 (select_kernel), else expand+backup of step i fused with the descent of step i+1
// (advance_kernel).  No early exits: every thread reaches the CTA barriers of the tail.
template <int NW, bool FIRST, int GROUPS>
__global__ void __launch_bounds__(c1t_threads(GROUPS), 1) descend_conv1_kernel(RulesParams rp, TreeView tv, int n, const float* __restrict__ policy,
                                     const float* __restrict__ value, int by_slot, int logits, double c_puct, int compact,
                                     int parity, Conv1TailArgs ct) {
  extern __shared__ uint8_t dc_smem_raw[];
  uint8_t* smem = dc_smem_raw + ((128u - (smem_u32(dc_smem_raw) & 127u)) & 127u);
  __shared__ int s_slot[4 * GROUPS];
  __shared__ int s_rank[4];     // warps of each lane quarter that have finished their descent
  __shared__ uint64_t s_leaf[4 * GROUPS][2 * NW + 1];      // the leaves, by warp: the tail reads them here, not from global memory
  if (threadIdx.x < 4) s_rank[threadIdx.x] = 0;
  const unsigned tl_idx = (threadIdx.x & 31) == 0 ? AZP_TL_ENTRY(AZP_TL_ADVANCE, gridDim.x * (blockDim.x >> 5)) : 0u;
  conv1_tail_prologue<GROUPS>(ct, smem);
  ExpandPre<NW> pre;
  pre.active = 0;
  RootPre<NW> rpre;
  rpre.have = 0;
  {
    const int g0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (!FIRST && g0 < n) {
      if ((compact >> 5) & 1) prefetch_root<NW>(rp, tv, g0, threadIdx.x & 31);
      if (!((compact >> 6) & 1)) {
        pre = expand_preload<NW>(tv, rp.A, g0, threadIdx.x & 31);
        if (!((compact >> 7) & 1)) rpre = root_prescan<NW>(rp, tv, g0, threadIdx.x & 31, c_puct, pre);     // bit 7: AZP_ROOT_PRESCAN=0
      }
    }
  }
  const bool use_pre = !FIRST && !((compact >> 6) & 1);      // bit 6: AZP_EXPAND_PRELOAD=0
  pdl_wait();
  pdl_launch_dependents();
  if ((threadIdx.x & 31) == 0) AZP_TL_START(AZP_TL_ADVANCE, tl_idx);
  const uint32_t tmem = conv1_tail_alloc<GROUPS>(smem);
  const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    tv.n_leaves[(parity ^ 1) * 2 + 0] = 0;
    tv.n_leaves[(parity ^ 1) * 2 + 1] = 0;
  }
  const bool root_prefetch = ((compact >> 4) & 1) && !((compact >> 5) & 1);
  compact &= 15;
  int slot = -1;
  if (g < n) {
    if (!FIRST && root_prefetch) {
      const size_t gb = (size_t)g * tv.cap * rp.A;
      const char* pn = reinterpret_cast<const char*>(tv.N + gb);
      const char* pw = reinterpret_cast<const char*>(tv.W + gb);
      const char* pp = reinterpret_cast<const char*>(tv.P + gb);
      const char* pc = reinterpret_cast<const char*>(tv.child + gb);
      for (int o = lane * 128; o < rp.A * 8; o += 32 * 128) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(pw + o));
        if (o < rp.A * 4) {
          asm volatile("prefetch.global.L2 [%0];" ::"l"(pn + o));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(pp + o));
          asm volatile("prefetch.global.L2 [%0];" ::"l"(pc + o));
        }
      }
    }
    bool ok = true;
    if (!FIRST) {
      ok = expand_backup_body<NW>(rp, tv, g, lane, policy, value, by_slot, logits, 0, use_pre ? &pre : nullptr);
      __syncwarp();             // the descent's lanes read edges other lanes of this warp have just written
    }
    if (ok) {
      if (!FIRST) root_merge<NW>(rpre, pre, c_puct);
      select_body<NW>(rp, tv, g, lane, c_puct, compact, parity, &slot, s_leaf[threadIdx.x >> 5], (!FIRST && rpre.have) ? &rpre.rs : nullptr);
    }
  }
  // team = how many warps of this lane quarter finished before this one (the team's barrier inside makes its members' leaf
  // records and slots -- written by lane 0 of each warp -- visible to the team)
  int team = 0;
  if (lane == 0) {
    team = atomicAdd(&s_rank[(threadIdx.x >> 5) & 3], 1);
    s_slot[team * 4 + ((threadIdx.x >> 5) & 3)] = slot < 0 ? -1 : (slot | ((int)(threadIdx.x >> 5) << 20));     // slot | warp << 20
  }
  team = __shfl_sync(AZP_FULL, team, 0);
  conv1_tail<GROUPS>(ct, smem, tmem, team, s_slot, &s_leaf[0][0], 2 * NW + 1, NW);
  conv1_tail_free<GROUPS>(tmem);
  if (lane == 0) AZP_TL_END(AZP_TL_ADVANCE, tl_idx, gridDim.x * (blockDim.x >> 5));
}

// ---- root visit counts (mcts.py:144) -----------------------------------------------------------
__global__ void root_visits_kernel(TreeView tv, int n, int32_t* __restrict__ visits) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n * tv.A) return;
  int g = i / tv.A, a = i % tv.A;
  // occupied cells were given N = 0 by root_expand and are never selected; no occupancy mask here, so the counts of a
  // search stay readable after the episode driver has played its move into the root position
  visits[i] = tv.N[(size_t)g * tv.cap * tv.A + a];
}

// root priors as stored by root_expand (after the Dirichlet mix of mcts.py:114-116); 0 on occupied cells
__global__ void root_priors_kernel(TreeView tv, int n, float* __restrict__ priors) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)n * tv.A) return;
  int g = i / tv.A, a = i % tv.A;
  priors[i] = tv.P[(size_t)g * tv.cap * tv.A + a];
}

__global__ void leaf_states_kernel(TreeView tv, int n, uint64_t* x, uint64_t* o, uint8_t* tm, int16_t* last,
                                   uint8_t* kind) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  int k = tv.leaf_kind[g];
  kind[g] = (uint8_t)k;
  int slot = tv.leaf_slot[g];
  for (int w = 0; w < tv.NW; w++) {
    x[(size_t)g * tv.NW + w] = (k == 0) ? tv.leaf_x[(size_t)slot * tv.NW + w] : 0;
    o[(size_t)g * tv.NW + w] = (k == 0) ? tv.leaf_o[(size_t)slot * tv.NW + w] : 0;
  }
  tm[g] = (k == 0) ? tv.leaf_tm[slot] : 0;
  last[g] = (k == 0) ? tv.leaf_last[slot] : (int16_t)-1;
}

// ================= self-play / arena episode machinery (self_play.py:47-72) ====================
struct EpisodeView {
  uint8_t* status;          // [G]
  uint16_t* moves;          // [G][max_plies]
  float* pi_hist;           // [G][max_plies][A]
  uint8_t* result;          // [G] result of a finished game
  const double* temperature; int n_temperature;
  long long* sched;         // {next local index, n_local, finished, local_first}
  float* raw_planes; float* raw_pi; int32_t* raw_z; int32_t* raw_count; long long raw_capacity;
  int max_plies;
  uint64_t seed;
};

// Hand the slot its next game (or park it).  Lane 0 only.
// `arena`: 0 self-play; 1 arena with two leaf batches per step (the games of a step are split between the two nets);
// 2 PHASE-ALIGNED arena: X (candidate, weight slot 0) moves in the even move steps of the match and O (baseline, slot 1)
// in the odd ones, in EVERY game -- a game whose first mover does not match the phase of the next step simply starts one
// step later -- so a step evaluates ONE leaf batch with ONE net instead of two half batches with two nets.  Games are
// independent and their random streams are keyed by (game id, ply), so the match is the same; only its kernel count halves.
__device__ __forceinline__ void seat_next_game(const TreeView& tv, const EpisodeView& ev, int g, int NW, int arena,
                                               int next_phase) {
  long long idx = atomicAdd((unsigned long long*)&ev.sched[0], 1ull);
  if (idx < ev.sched[1]) {
    long long gid = ev.sched[3] + idx;
    for (int k = 0; k < NW; k++) { tv.root_x[(size_t)g * NW + k] = 0; tv.root_o[(size_t)g * NW + k] = 0; }
    // arena: X (candidate) moves first in even games, O in odd games (evaluator.py:64-69)
    int first = (arena && (gid & 1)) ? 2 : 1;
    tv.root_tm[g] = (uint8_t)first;
    tv.root_last[g] = -1;
    tv.ply[g] = 0;
    tv.game_id[g] = gid;
    tv.active[g] = 1;
    tv.net_of_game[g] = (uint8_t)(arena == 1 ? (first == 1 ? 0 : 1) : 0);
    ev.status[g] = SLOT_PLAYING;
    if (arena == 2 && (first == 1) != (next_phase == 0)) {
      tv.active[g] = 0;
      ev.status[g] = SLOT_DELAYED;
    }
  } else {
    tv.active[g] = 0;
    ev.status[g] = SLOT_IDLE;
  }
}

__global__ void seat_games_kernel(TreeView tv, EpisodeView ev, int n, int arena, int next_phase) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n) return;
  seat_next_game(tv, ev, g, tv.NW, arena, next_phase);
}

// phase-aligned arena, start of a move step: the waiting games whose first mover matches this step's phase begin
__global__ void arena_activate_kernel(TreeView tv, EpisodeView ev, int n, int phase) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n || ev.status[g] != SLOT_DELAYED) return;
  if ((tv.root_tm[g] == 1) == (phase == 0)) { ev.status[g] = SLOT_PLAYING; tv.active[g] = 1; }
}

// MCTS.run epilogue + the worker's bookkeeping: pi from root visits at T(ply), sample the move,
// record (pi, move), play it, detect the end of the game.
template <int NW>
__global__ void end_move_kernel(RulesParams rp, TreeView tv, EpisodeView ev, int n, int arena) {
  int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (g >= n || ev.status[g] != SLOT_PLAYING) return;
  const int A = rp.A;
  BB<NW> x = bb_load<NW>(tv.root_x + (size_t)g * NW), o = bb_load<NW>(tv.root_o + (size_t)g * NW);
  const int ply = tv.ply[g];
  int tm = tv.root_tm[g];
  // temperature index: the ply in self-play (self_play.py:53); in the arena the mover's own step
  // counter, 0,1,2.. for the first mover and 1,2,3.. for the second (evaluator.py:71-90)
  int tidx = arena ? ((ply & 1) ? ply / 2 + 1 : ply / 2) : ply;
  if (tidx >= ev.n_temperature) tidx = ev.n_temperature - 1;
  double T = ev.temperature[tidx];
  if (T <= 1e-7) T = 1e-7;
  const int32_t* Nroot = tv.N + (size_t)g * tv.cap * A;
  float* pi = ev.pi_hist + ((size_t)g * ev.max_plies + ply) * A;

  double mx = -1.0e300;
  int k = 0;
  for (int a = lane; a < A; a += 32) {
    if (bb_test<NW>(x, a) || bb_test<NW>(o, a)) continue;
    mx = fmax(mx, (double)logf((float)Nroot[a] + 1e-8f) / T);
    k++;
  }
  mx = warp_max_d(mx);
  k = warp_sum_i(k);
  double sum = 0.0;
  for (int a = lane; a < A; a += 32) {
    if (bb_test<NW>(x, a) || bb_test<NW>(o, a)) continue;
    sum += exp((double)logf((float)Nroot[a] + 1e-8f) / T - mx);
  }
  sum = warp_sum_d(sum);
  const bool uniform = !(sum >= 1e-8);

  // action ~ Categorical(probs) by inverse CDF over the children in row-major order (mcts.py:177)
  Philox rng = philox_for_game(ev.seed, tv.game_id[g]);
  uint32_t r[4];
  rng(0u, (uint32_t)ply, AZP_RNG_ACTION, 0u, r);
  const double u = u01_open(r[0], r[1]);
  double run = 0.0;        // probability mass of all earlier chunks
  int action = -1, lastlegal = -1;
  for (int a0 = 0; a0 < A; a0 += 32) {
    int a = a0 + lane;
    bool legal = a < A && !(bb_test<NW>(x, a) || bb_test<NW>(o, a));
    double p = 0.0;
    if (legal) p = uniform ? 1.0 / k : exp((double)logf((float)Nroot[a] + 1e-8f) / T - mx) / sum;
    if (a < A) pi[a] = (float)p;          // the arena keeps it for azp_episode_snapshot only (no examples are emitted)
    double c = p;            // inclusive scan inside the chunk
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      double t = __shfl_up_sync(AZP_FULL, c, off);
      if (lane >= off) c += t;
    }
    c += run;
    unsigned hit = __ballot_sync(AZP_FULL, legal && c > u);
    unsigned leg = __ballot_sync(AZP_FULL, legal);
    if (leg) lastlegal = a0 + 31 - __clz(leg);
    if (action < 0 && hit) action = a0 + __ffs(hit) - 1;
    run = __shfl_sync(AZP_FULL, c, 31);
  }
  if (action < 0) action = lastlegal;   // u beyond the rounded total mass
  if (action < 0) {                     // no legal move on a playing game: cannot happen
    if (lane == 0) atomicExch(tv.error_flag, AZP_ERR_INVALID_MOVE);
    return;
  }
  if (tm == 1) bb_set<NW>(x, action); else bb_set<NW>(o, action);
  int res = warp_result<NW>(rp, x, o, lane);
  if (res == 0 && ply + 1 >= ev.max_plies) res = 3;     // ply cap (only when max_plies < B*B): scored as a draw
  if (lane == 0) {
    ev.moves[(size_t)g * ev.max_plies + ply] = (uint16_t)action;
    bb_store<NW>(tv.root_x + (size_t)g * NW, x);
    bb_store<NW>(tv.root_o + (size_t)g * NW, o);
    tv.root_tm[g] = (uint8_t)(3 - tm);
    tv.root_last[g] = (int16_t)action;
    tv.ply[g] = ply + 1;
    if (arena == 1) tv.net_of_game[g] = (uint8_t)((3 - tm) == 1 ? 0 : 1);
    count(tv, CNT_PLIES, 1);
    if (res) {
      ev.result[g] = (uint8_t)res;
      ev.status[g] = SLOT_PENDING;
      tv.active[g] = 0;
      count(tv, CNT_GAMES, 1);
      count(tv, res == 1 ? CNT_XWINS : (res == 2 ? CNT_OWINS : CNT_DRAWS), 1);
    }
  }
}

// Emit the finished game's positions as raw (planes, pi, z) rows and seat the next game.
// z = 0 on a draw, +1 if the player to move at that position won, else -1 (self_play.py:71).
template <int NW>
__global__ void finish_games_kernel(RulesParams rp, TreeView tv, EpisodeView ev, int n, int arena, int next_phase) {
  int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (g >= n || ev.status[g] != SLOT_PENDING) return;
  const int A = rp.A;
  const int L = tv.ply[g];
  if (!arena) {
    int base = 0;
    if (lane == 0) {
      base = atomicAdd(ev.raw_count, L);
      if ((long long)base + L > ev.raw_capacity) { atomicSub(ev.raw_count, L); base = -1; }
    }
    base = __shfl_sync(AZP_FULL, base, 0);
    if (base < 0) return;      // slab full: stay pending, retry after the next drain
    const int res = ev.result[g];
    BB<NW> x, o;
#pragma unroll
    for (int k = 0; k < NW; k++) { x.w[k] = 0; o.w[k] = 0; }
    int tm = 1, last = -1;     // self-play games start with X on the empty board (games.py:29)
    for (int p = 0; p < L; p++) {
      float* op = ev.raw_planes + (size_t)(base + p) * 4 * A;
      for (int a = lane; a < A; a += 32) {
        bool isx = bb_test<NW>(x, a), iso = bb_test<NW>(o, a);
        op[a] = ((tm == 1) ? isx : iso) ? 1.f : 0.f;
        op[A + a] = ((tm == 1) ? iso : isx) ? 1.f : 0.f;
        op[2 * A + a] = (a == last) ? 1.f : 0.f;
        op[3 * A + a] = 0.f;
      }
      const float* ip = ev.pi_hist + ((size_t)g * ev.max_plies + p) * A;
      float* opi = ev.raw_pi + (size_t)(base + p) * A;
      for (int a = lane; a < A; a += 32) opi[a] = ip[a];
      if (lane == 0) ev.raw_z[base + p] = res == 3 ? 0 : (res == tm ? 1 : -1);
      int mv = ev.moves[(size_t)g * ev.max_plies + p];
      if (tm == 1) bb_set<NW>(x, mv); else bb_set<NW>(o, mv);
      tm = 3 - tm;
      last = mv;
    }
    if (lane == 0) count(tv, CNT_POSITIONS, (unsigned long long)L);
  }
  if (lane == 0) {
    atomicAdd((unsigned long long*)&ev.sched[2], 1ull);
    seat_next_game(tv, ev, g, NW, arena, next_phase);
  }
}

// after the augmenting writer: advance the output cursor and empty the raw slab
__global__ void drain_raw_kernel(int32_t* raw_count, long long* out_written, int mult, unsigned long long* counters) {
  long long nraw = *raw_count;
  *out_written += nraw * mult;
  atomicAdd(counters + CNT_EXAMPLES, (unsigned long long)(nraw * mult));
  *raw_count = 0;
}

// rows of raw logits -> probabilities (azp_net_forward on the bf16 path); one warp per row, same arithmetic as the
// consumers inside the search
__global__ void softmax_rows_kernel(int n, int A, const float* __restrict__ logits, float* __restrict__ policy) {
  int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (r >= n) return;
  const float* row = logits + (size_t)r * A;
  float m, inv;
  warp_softmax_stats(row, A, lane, m, inv);
  for (int a = lane; a < A; a += 32) policy[(size_t)r * A + a] = __expf(row[a] - m) * inv;
}

// ================= host side ====================================================================
static inline int warp_grid(int n) { return (n * 32 + 127) / 128; }
int azp_softmax_rows(int n, int A, const float* logits, float* policy, cudaStream_t s) {
  softmax_rows_kernel<<<warp_grid(n), 128, 0, s>>>(n, A, logits, policy);
  azp_count_launch();
  return azp_check_launch();
}

static EpisodeView episode_view(azp_engine* e) {
  EpisodeView ev;
  ev.status = e->d_status;
  ev.moves = e->d_moves;
  ev.pi_hist = e->d_pi_hist;
  ev.result = e->d_result;
  ev.temperature = e->d_temperature;
  ev.n_temperature = e->n_temperature;
  ev.sched = e->d_sched;
  ev.raw_planes = e->d_raw_planes; ev.raw_pi = e->d_raw_pi; ev.raw_z = e->d_raw_z;
  ev.raw_count = e->d_raw_count; ev.raw_capacity = e->raw_capacity;
  ev.max_plies = e->max_plies;
  ev.seed = e->cfg.seed;
  return ev;
}

int azp_tree_begin_move(azp_engine* e, int n, int emit, int compact, int parity, cudaStream_t s) {
  AZP_DISPATCH_NW(e->NW, begin_move_kernel, warp_grid(n), 128, 0, s, e->rp, e->tv, n, emit, compact, parity);
  azp_count_launch();
  return azp_check_launch();
}
int azp_tree_root_expand(azp_engine* e, int n, const float* policy, int by_slot, int add_noise, int logits, cudaStream_t s) {
  azp_prof_begin(e, AZP_PROF_ROOT_EXPAND, nullptr, n, s);
  AZP_DISPATCH_NW(e->NW, root_expand_kernel, warp_grid(n), 128, 0, s, e->rp, e->tv, n, policy, by_slot, add_noise,
                  e->cfg.dirichlet_alpha, e->cfg.dirichlet_weight, e->cfg.seed, logits);
  azp_count_launch();
  azp_prof_end(e, AZP_PROF_ROOT_EXPAND, s);
  return azp_check_launch();
}
int azp_tree_select(azp_engine* e, int n, int compact, int parity, cudaStream_t s) {
  AZP_DISPATCH_NW_PDL(e->NW, select_kernel, warp_grid(n), 128, 0, s, e->rp, e->tv, n, e->cfg.c_puct, compact, parity);
  azp_count_launch();
  return azp_check_launch();
}
int azp_tree_expand_backup(azp_engine* e, int n, const float* policy, const void* value, int by_slot, int logits,
                           int value_f64, cudaStream_t s) {
  AZP_DISPATCH_NW_PDL(e->NW, expand_backup_kernel, warp_grid(n), 128, 0, s, e->rp, e->tv, n, policy, value, by_slot, logits,
                      value_f64);
  azp_count_launch();
  return azp_check_launch();
}
int azp_tree_advance(azp_engine* e, int n, const float* policy, const float* value, int by_slot, int logits, int compact,
                     int parity, cudaStream_t s) {
  static int prefetch = -1;
  if (prefetch < 0) { const char* v = getenv("AZP_ROOT_PREFETCH"); prefetch = v ? atoi(v) : 2; }     // 0 off, 1 after the wait, 2 before it
  static int preload = -1;
  if (preload < 0) { const char* v = getenv("AZP_EXPAND_PRELOAD"); preload = v ? atoi(v) : 1; }
  static int prescan = -1;
  if (prescan < 0) { const char* v = getenv("AZP_ROOT_PRESCAN"); prescan = v ? atoi(v) : 0; }      // measured: -2.3 % (the scan is not free before the wait: most CTAs only become resident when the FC kernel leaves)
  azp_prof_begin(e, AZP_PROF_ADVANCE, nullptr, n, s);
  AZP_DISPATCH_NW_PDL(e->NW, advance_kernel, warp_grid(n), 128, 0, s, e->rp, e->tv, n, policy, value, by_slot, logits,
                      e->cfg.c_puct, compact | ((prefetch != 0) << 4) | ((prefetch == 2) << 5) | ((preload == 0) << 6) | ((prescan == 0) << 7), parity);
  azp_count_launch();
  azp_prof_end(e, AZP_PROF_ADVANCE, s);
  return azp_check_launch();
}
template <int NW, int GROUPS>
static void descend_conv1_launch(azp_engine* e, int n, bool first, const float* policy, const float* value, int by_slot, int logits,
                                 int cp, int parity, const Conv1TailArgs& ct, cudaStream_t s) {
  // per instantiation AND device (function attributes belong to the device's context): static + dynamic shared memory may
  // exceed the 48 KB a kernel gets without opting in
  static bool configured[64] = {false};
  const int dev = e->cfg.device & 63;
  if (!configured[dev]) {
    configured[dev] = true;
    cudaFuncSetAttribute(descend_conv1_kernel<NW, true, GROUPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, c1t_smem(GROUPS));
    cudaFuncSetAttribute(descend_conv1_kernel<NW, false, GROUPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, c1t_smem(GROUPS));
    cudaGetLastError();
  }
  const int grid = (n * 32 + c1t_threads(GROUPS) - 1) / c1t_threads(GROUPS);
  if (first)
    azp_launch(descend_conv1_kernel<NW, true, GROUPS>, grid, c1t_threads(GROUPS), c1t_smem(GROUPS), s, true, e->rp, e->tv, n, policy, value,
               by_slot, logits, e->cfg.c_puct, cp, parity, ct);
  else
    azp_launch(descend_conv1_kernel<NW, false, GROUPS>, grid, c1t_threads(GROUPS), c1t_smem(GROUPS), s, true, e->rp, e->tv, n, policy, value,
               by_slot, logits, e->cfg.c_puct, cp, parity, ct);
}
template <int NW>
static void descend_conv1_groups(azp_engine* e, int groups, int n, bool first, const float* policy, const float* value, int by_slot,
                                 int logits, int cp, int parity, const Conv1TailArgs& ct, cudaStream_t s) {
  if (groups == 1) descend_conv1_launch<NW, 1>(e, n, first, policy, value, by_slot, logits, cp, parity, ct, s);
  else if (groups == 2) descend_conv1_launch<NW, 2>(e, n, first, policy, value, by_slot, logits, cp, parity, ct, s);
  else descend_conv1_launch<NW, 4>(e, n, first, policy, value, by_slot, logits, cp, parity, ct, s);
}
// `groups` (1, 2 or 4 warp groups of four games per CTA) comes from azp_umma_conv1_tail: the smallest that seats every game in one
// CTA per SM
int azp_tree_descend_conv1(azp_engine* e, int n, bool first, const float* policy, const float* value, int by_slot, int logits,
                           int compact, int parity, const Conv1TailArgs& ct, cudaStream_t s) {
  static int prefetch = -1;
  if (prefetch < 0) { const char* v = getenv("AZP_ROOT_PREFETCH"); prefetch = v ? atoi(v) : 2; }     // 0 off, 1 after the wait, 2 before it
  static int preload = -1;
  if (preload < 0) { const char* v = getenv("AZP_EXPAND_PRELOAD"); preload = v ? atoi(v) : 1; }
  static int prescan = -1;
  if (prescan < 0) { const char* v = getenv("AZP_ROOT_PRESCAN"); prescan = v ? atoi(v) : 0; }      // measured: -2.3 % (the scan is not free before the wait: most CTAs only become resident when the FC kernel leaves)
  const int cp = compact | ((prefetch != 0) << 4) | ((prefetch == 2) << 5) | ((preload == 0) << 6) | ((prescan == 0) << 7);
  azp_prof_begin(e, AZP_PROF_ADVANCE, nullptr, n, s);
  switch (e->NW) {
    case 1: descend_conv1_groups<1>(e, ct.enabled, n, first, policy, value, by_slot, logits, cp, parity, ct, s); break;
    case 2: descend_conv1_groups<2>(e, ct.enabled, n, first, policy, value, by_slot, logits, cp, parity, ct, s); break;
    case 3: descend_conv1_groups<3>(e, ct.enabled, n, first, policy, value, by_slot, logits, cp, parity, ct, s); break;
    case 4: descend_conv1_groups<4>(e, ct.enabled, n, first, policy, value, by_slot, logits, cp, parity, ct, s); break;
    case 5: descend_conv1_groups<5>(e, ct.enabled, n, first, policy, value, by_slot, logits, cp, parity, ct, s); break;
    default: descend_conv1_groups<6>(e, ct.enabled, n, first, policy, value, by_slot, logits, cp, parity, ct, s); break;
  }
  azp_count_launch();
  azp_prof_end(e, AZP_PROF_ADVANCE, s);
  return azp_check_launch();
}
int azp_tree_root_visits(azp_engine* e, int n, int32_t* visits, cudaStream_t s) {
  root_visits_kernel<<<(int)(((size_t)n * e->A + 255) / 256), 256, 0, s>>>(e->tv, n, visits);
  azp_count_launch();
  return azp_check_launch();
}
int azp_tree_root_priors(azp_engine* e, int n, float* priors, cudaStream_t s) {
  root_priors_kernel<<<(int)(((size_t)n * e->A + 255) / 256), 256, 0, s>>>(e->tv, n, priors);
  azp_count_launch();
  return azp_check_launch();
}
int azp_tree_leaf_states(azp_engine* e, int n, uint64_t* x, uint64_t* o, uint8_t* tm, int16_t* last, uint8_t* kind,
                         cudaStream_t s) {
  leaf_states_kernel<<<(n + 127) / 128, 128, 0, s>>>(e->tv, n, x, o, tm, last, kind);
  azp_count_launch();
  return azp_check_launch();
}
int azp_episode_seat(azp_engine* e, int n, int arena, int next_phase, cudaStream_t s) {
  seat_games_kernel<<<(n + 127) / 128, 128, 0, s>>>(e->tv, episode_view(e), n, arena, next_phase);
  azp_count_launch();
  return azp_check_launch();
}
int azp_episode_end_move(azp_engine* e, int n, int arena, cudaStream_t s) {
  EpisodeView ev = episode_view(e);
  azp_prof_begin(e, AZP_PROF_END_MOVE, nullptr, n, s);
  AZP_DISPATCH_NW(e->NW, end_move_kernel, warp_grid(n), 128, 0, s, e->rp, e->tv, ev, n, arena);
  azp_count_launch();
  azp_prof_end(e, AZP_PROF_END_MOVE, s);
  return azp_check_launch();
}
int azp_episode_activate(azp_engine* e, int n, int phase, cudaStream_t s) {
  arena_activate_kernel<<<(n + 127) / 128, 128, 0, s>>>(e->tv, episode_view(e), n, phase);
  azp_count_launch();
  return azp_check_launch();
}
int azp_episode_finish(azp_engine* e, int n, int arena, int next_phase, cudaStream_t s) {
  EpisodeView ev = episode_view(e);
  azp_prof_begin(e, AZP_PROF_FINISH, nullptr, n, s);
  AZP_DISPATCH_NW(e->NW, finish_games_kernel, warp_grid(n), 128, 0, s, e->rp, e->tv, ev, n, arena, next_phase);
  azp_count_launch();
  azp_prof_end(e, AZP_PROF_FINISH, s);
  return azp_check_launch();
}
int azp_episode_drain(azp_engine* e, long long* out_written_dev, int mult, cudaStream_t s) {
  drain_raw_kernel<<<1, 1, 0, s>>>(e->d_raw_count, out_written_dev, mult, e->tv.counters);
  azp_count_launch();
  return azp_check_launch();
}
