// Elite selection (top-k by ascending (cost, index)) and distribution refit.
//
// Replaces (paths relative to /root/reference/icem/):
//   controllers/icem.py:199       elite_idxs = argsort(costs)[:num_elites]
//   controllers/icem.py:201-211   elite_samples, mean/std(ddof=0) of elite actions, momentum update
//   controllers/icem.py:142-145   kept elites of the previous iteration join the population (not re-simulated)
//   controllers/icem.py:149,163   best trajectory = elite 0; executed action = its first action
//   controllers/icem.py:167-175   mean time-shift, std reset (last iteration)
//
// select_kernel: every CTA extracts the k smallest keys of its chunk with k rounds of a
// warp-shuffle / block-reduce arg-min over order-preserving 64-bit (cost, global index) keys; the
// last CTA to finish (atomic ticket) merges the per-CTA candidates.  On one GPU it continues
// straight into merge_refit; with R ranks it writes this rank's k elite records {key, actions}
// for the NCCL all-gather and merge_refit_kernel runs after it.
#pragma once
#include "common.cuh"
#include "rollout.cuh"

namespace icem {

constexpr int kSelectThreads = 256;
constexpr unsigned long long kKeyMax = ~0ull;

struct SelectArgs {
  // population of this rank (same conventions as RolloutArgs)
  int n_fresh_local, n_shift_local, global_offset, n_fresh_global, iteration;
  int k;                    // num_elites
  int stride;               // floats per action row
  const float* costs;       // [rows]
  const float* actions;     // [rows][stride]
  const StepState* ss;
  unsigned long long* cand; // [gridDim.x][k] scratch
  unsigned int* ticket;     // zero-initialised, self-resetting
  // multi-rank: where this rank's record goes ([k] keys then [k][stride] actions); null on one GPU
  unsigned long long* send_keys;
  float* send_actions;
  // several problems in one launch (blockIdx.y = problem): element strides between problems (0 for one problem)
  unsigned long long prob_actions, prob_costs;
  int prob_cand;
};

struct RefitArgs {
  int h, d, k, stride, iteration, last_iteration;
  int world;                       // ranks contributing records
  int n_keep;                      // kept elites of the previous iteration joining (0 at iteration 0 / flag off)
  int n_fresh_global;              // N_i: global index of kept elite 0
  float alpha, one_minus_alpha;
  // gathered records: rank r at rec_keys + r*rec_rank_stride_keys, rec_actions + r*rec_rank_stride_floats.
  // On one GPU rec_keys points at the k merged local keys and rows are resolved into `actions`.
  const unsigned long long* rec_keys;
  const float* rec_actions;
  size_t rec_rank_stride_bytes;    // bytes between consecutive ranks' records (keys and actions share a buffer)
  const float* local_actions;      // single-GPU: population buffer rows (else null)
  int local_n_fresh, local_offset; // single-GPU row resolution
  // elites (double-buffered): previous iteration's (read) and this iteration's (written)
  const float* prev_elite_actions; const float* prev_elite_costs;
  float* new_elite_actions; float* new_elite_costs; int32_t* new_elite_idx;
  float* mean; float* std;         // planner distribution [h*d], updated in place
  const float* init_std;           // [h*d] reset value (icem.py:175)
  // per-iteration record of this plan step
  float* trace_mean; float* trace_std; float* trace_costs; int32_t* trace_idx;
  // MpcCemStd switches (controllers/mpc.py:237-248, 290-301); all 0 / null for MpcICem
  int cem_std, execute_mean, mean_to_zero, levine;
  const float* low; const float* high;   // [d] action bounds (levine std clamp)
  float* out_action;               // [d]  (last iteration)
  float* out_best_cost;            // [1]  min(costs) of the last iteration (icem.py:177)
  // several problems in one launch: element strides between problems (0 for one problem)
  unsigned long long prob_actions;
  int prob_dist, prob_elites, prob_k, prob_trace_hd, prob_trace_k, prob_out;
};

// -------------------------------------------------------------------------------------------------
// The k smallest of `count` unique 64-bit keys, ascending, into out[0..k) (smem or global).  KeyFn(i) -> key of
// element i.  All threads of the CTA must call.  Keys are unique, so "smallest key greater than the last extracted
// one" needs no marking.
//
// Fast path (count <= 8 keys per thread, i.e. every select_kernel chunk and every merge): each thread reads its keys
// ONCE into registers, each WARP extracts the k smallest of its 256 keys with shuffles only (the owner of an
// extracted minimum advances over its own registers), and one warp merges the 8 x k warp candidates the same way:
// two CTA barriers in total.  The general path below does k block-wide rounds with two barriers each and lets the
// owner rescan global memory (a chain of L2 latencies per round): ~40 us per launch against ~15.
constexpr int kSelPerThread = 8;
constexpr int kSelMaxK = 64;

template <class KeyFn>
__device__ void block_topk(KeyFn key, int count, int k, unsigned long long* out) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nwarps = blockDim.x >> 5;
  if (count <= (int)blockDim.x * kSelPerThread && k <= kSelMaxK && nwarps <= 8) {
    __shared__ unsigned long long s_cand[8 * kSelMaxK];
    unsigned long long v[kSelPerThread];
    unsigned long long mine = kKeyMax;
#pragma unroll
    for (int j = 0; j < kSelPerThread; ++j) {
      const int i = tid + j * (int)blockDim.x;
      v[j] = i < count ? key(i) : kKeyMax;
      mine = v[j] < mine ? v[j] : mine;
    }
    for (int r = 0; r < k; ++r) {                     // this warp's k smallest
      const unsigned long long w = warp_min_u64(mine);
      if (lane == 0) s_cand[warp * k + r] = w;
      if (mine == w && w != kKeyMax) {
        unsigned long long nxt = kKeyMax;
#pragma unroll
        for (int j = 0; j < kSelPerThread; ++j) nxt = (v[j] > w && v[j] < nxt) ? v[j] : nxt;
        mine = nxt;
      }
    }
    __syncthreads();
    if (warp == 0) {                                  // merge nwarps * k candidates (<= 512: 16 per lane)
      const int total = nwarps * k;
      unsigned long long c[16];
      unsigned long long m2 = kKeyMax;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int i = lane + 32 * j;
        c[j] = i < total ? s_cand[i] : kKeyMax;
        m2 = c[j] < m2 ? c[j] : m2;
      }
      for (int r = 0; r < k; ++r) {
        const unsigned long long w = warp_min_u64(m2);
        if (lane == 0) out[r] = w;
        if (m2 == w && w != kKeyMax) {
          unsigned long long nxt = kKeyMax;
#pragma unroll
          for (int j = 0; j < 16; ++j) nxt = (c[j] > w && c[j] < nxt) ? c[j] : nxt;
          m2 = nxt;
        }
      }
    }
    __syncthreads();
    return;
  }
  __shared__ unsigned long long s_part[kSelectThreads / 32];
  __shared__ unsigned long long s_min;
  unsigned long long mine = kKeyMax;
  for (int i = tid; i < count; i += blockDim.x) {
    const unsigned long long v = key(i);
    mine = v < mine ? v : mine;
  }
  for (int r = 0; r < k; ++r) {
    const unsigned long long w = warp_min_u64(mine);
    if (lane == 0) s_part[warp] = w;
    __syncthreads();
    if (warp == 0) {
      unsigned long long v = lane < (blockDim.x >> 5) ? s_part[lane] : kKeyMax;
      v = warp_min_u64(v);
      if (lane == 0) { s_min = v; out[r] = v; }
    }
    __syncthreads();
    const unsigned long long m = s_min;
    if (mine == m && m != kKeyMax) {     // owner rescans for its next candidate
      unsigned long long nxt = kKeyMax;
      for (int i = tid; i < count; i += blockDim.x) {
        const unsigned long long v = key(i);
        if (v > m && v < nxt) nxt = v;
      }
      mine = nxt;
    }
  }
  __syncthreads();
}

// -------------------------------------------------------------------------------------------------
// One CTA: merge candidate records (+ kept elites), gather elite actions, refit, record, shift.
__device__ void merge_refit(const RefitArgs& r, unsigned long long* s_keys /* [k] smem scratch */) {
  const int tid = threadIdx.x;
  const int k = r.k, hd = r.h * r.d;
  const int n_rec = r.world * k;
  const int n_cand = n_rec + r.n_keep;
  const char* rec_base = reinterpret_cast<const char*>(r.rec_keys);
  auto rec_key = [&](int c) -> unsigned long long {
    const int rank = c / k, j = c - rank * k;
    return reinterpret_cast<const unsigned long long*>(rec_base + (size_t)rank * r.rec_rank_stride_bytes)[j];
  };
  auto key = [&](int c) -> unsigned long long {
    if (c < n_rec) return rec_key(c);
    const int j = c - n_rec;     // kept elite j enters with its stored cost at global index N_i + j
    return cost_key(r.prev_elite_costs[j], (uint32_t)(r.n_fresh_global + j));
  };
  block_topk(key, n_cand, k, s_keys);

  // resolve each elite to its action row
  __shared__ const float* s_src[64];
  if (tid < k) {
    const unsigned long long kk = s_keys[tid];
    const uint32_t gidx = key_index(kk);
    const float* src = nullptr;
    for (int c = 0; c < n_cand; ++c) {     // which candidate carries this key (n_cand is tiny)
      if (key(c) != kk) continue;
      if (c >= n_rec) {
        src = r.prev_elite_actions + (size_t)(c - n_rec) * r.stride;
      } else if (r.local_actions) {
        const int row = gidx >= (uint32_t)r.n_fresh_global ? r.local_n_fresh + (int)(gidx - r.n_fresh_global)
                                                           : (int)gidx - r.local_offset;
        src = r.local_actions + (size_t)row * r.stride;
      } else {
        const int rank = c / k, j = c - rank * k;
        src = reinterpret_cast<const float*>(rec_base + (size_t)rank * r.rec_rank_stride_bytes +
                                             (size_t)k * sizeof(unsigned long long)) + (size_t)j * r.stride;
      }
      break;
    }
    s_src[tid] = src;
    // the cost is recoverable from the key (inverse of cost_key)
    uint32_t b = (uint32_t)(kk >> 32);
    b = (b & 0x80000000u) ? (b & 0x7FFFFFFFu) : ~b;
    const float c = __uint_as_float(b);
    r.new_elite_idx[tid] = (int32_t)gidx;
    r.trace_idx[tid] = (int32_t)gidx;
    r.new_elite_costs[tid] = c;
    r.trace_costs[tid] = c;
  }
  __syncthreads();

  const float inv_k = 1.0f / (float)k;
  for (int e = tid; e < hd; e += blockDim.x) {
    // The elite rows are read-only here (population / previous elites / gathered records; the new elite buffer is the
    // other half of a double buffer), so they go through the non-coherent path: with plain loads every load had to
    // wait for the store before it (possible alias), a chain of ~20 L2 latencies per element -- most of the kernel.
    float sum = 0.f;
    for (int j = 0; j < k; ++j) sum += __ldg(s_src[j] + e);
    const float m = sum * inv_k;
    float var = 0.f;
    for (int j = 0; j < k; ++j) {
      const float v = __ldg(s_src[j] + e);
      const float dv = v - m;
      var = fmaf(dv, dv, var);
      r.new_elite_actions[(size_t)j * r.stride + e] = v;
    }
    const float sd = sqrtf(var * inv_k);                       // ddof = 0 (icem.py:208)
    const float nm = r.one_minus_alpha * m + r.alpha * r.mean[e];   // icem.py:210-211
    float ns = r.one_minus_alpha * sd + r.alpha * r.std[e];
    if (r.levine) {            // mpc.py:291-294 (_update_bounds after every refit)
      const int dim = e % r.d;
      ns = fmaxf(1e-8f, fminf(fminf((nm - r.low[dim]) * 0.5f, (r.high[dim] - nm) * 0.5f), ns));
    }
    r.trace_mean[e] = nm;
    r.trace_std[e] = ns;
    if (!r.last_iteration) {
      r.mean[e] = nm;
      r.std[e] = ns;
    }
  }
  __syncthreads();
  if (r.last_iteration) {
    // icem.py:163 / mpc.py:237-240: executed action = first action of the best trajectory of the last population,
    // or (MpcCemStd with execute_best_elite false) the first row of the refitted mean
    for (int e = tid; e < r.d; e += blockDim.x) r.out_action[e] = r.execute_mean ? r.trace_mean[e] : s_src[0][e];
    if (tid == 0) r.out_best_cost[0] = r.new_elite_costs[0];
    // icem.py:167-175: mean[:-1] = mean[1:], last row kept; std reset.  mpc.py:243-252: last row zero with
    // bounds_like_levine, mean = zeros without shift_means; the std reset is followed by _update_bounds
    for (int e = tid; e < hd; e += blockDim.x) {
      const int src = e + r.d < hd ? e + r.d : e;
      float nm = r.trace_mean[src];
      if (r.cem_std && r.levine && e + r.d >= hd) nm = 0.f;
      if (r.mean_to_zero) nm = 0.f;
      float ns = r.init_std[e];
      if (r.levine) {
        const int dim = e % r.d;
        ns = fmaxf(1e-8f, fminf(fminf((nm - r.low[dim]) * 0.5f, (r.high[dim] - nm) * 0.5f), ns));
      }
      r.mean[e] = nm;
      r.std[e] = ns;
    }
  }
}

__global__ void __launch_bounds__(kSelectThreads) select_kernel(SelectArgs a, RefitArgs r, int fuse_refit) {
  extern __shared__ unsigned long long s_keys[];   // [k]
  __shared__ bool s_last;
  if (blockIdx.y) {          // this CTA works on problem blockIdx.y
    const unsigned long long pr = blockIdx.y;
    a.costs += pr * a.prob_costs;
    a.actions += pr * a.prob_actions;
    a.cand += pr * (unsigned)a.prob_cand;
    a.ticket += pr;
    r.local_actions += pr * r.prob_actions;
    r.prev_elite_actions += pr * (unsigned)r.prob_elites;
    r.new_elite_actions += pr * (unsigned)r.prob_elites;
    r.prev_elite_costs += pr * (unsigned)r.prob_k;
    r.new_elite_costs += pr * (unsigned)r.prob_k;
    r.new_elite_idx += pr * (unsigned)r.prob_k;
    r.mean += pr * (unsigned)r.prob_dist;
    r.std += pr * (unsigned)r.prob_dist;
    r.trace_mean += pr * (unsigned)r.prob_trace_hd;
    r.trace_std += pr * (unsigned)r.prob_trace_hd;
    r.trace_costs += pr * (unsigned)r.prob_trace_k;
    r.trace_idx += pr * (unsigned)r.prob_trace_k;
    r.out_action += pr * (unsigned)r.prob_out;
    r.out_best_cost += pr * (unsigned)r.prob_out;
  }
  const StepState ss = *a.ss;
  const int n_rows = a.n_fresh_local + ((a.iteration == 0 && ss.has_prev_elites) ? a.n_shift_local : 0);
  const int chunk = (n_rows + gridDim.x - 1) / gridDim.x;
  const int lo = min(n_rows, (int)blockIdx.x * chunk);
  const int cnt = min(n_rows, lo + chunk) - lo;
  auto key = [&](int i) -> unsigned long long {
    const int row = lo + i;
    const uint32_t gidx = row >= a.n_fresh_local ? (uint32_t)(a.n_fresh_global + (row - a.n_fresh_local))
                                                 : (uint32_t)(a.global_offset + row);
    return cost_key(a.costs[row], gidx);
  };
  block_topk(key, cnt, a.k, s_keys);
  if (threadIdx.x < a.k) a.cand[(size_t)blockIdx.x * a.k + threadIdx.x] = s_keys[threadIdx.x];
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int t = atomicAdd(a.ticket, 1u);
    s_last = (t == gridDim.x - 1);
    if (s_last) *a.ticket = 0u;      // self-reset for the next launch
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const volatile unsigned long long* cand = a.cand;
  auto key2 = [&](int i) -> unsigned long long { return cand[i]; };
  if (gridDim.x > 1) block_topk(key2, (int)gridDim.x * a.k, a.k, s_keys);

  if (fuse_refit) {
    // single GPU: the k merged keys are "rank 0's record"; rows resolve into the population buffer
    __shared__ unsigned long long s_rec[64];
    if (threadIdx.x < a.k) s_rec[threadIdx.x] = s_keys[threadIdx.x];
    __syncthreads();
    r.rec_keys = s_rec;
    merge_refit(r, s_keys);
  } else {
    // multi GPU: publish this rank's record for the all-gather
    if (threadIdx.x < a.k) a.send_keys[threadIdx.x] = s_keys[threadIdx.x];
    for (int j = 0; j < a.k; ++j) {
      const unsigned long long kk = s_keys[j];
      if (kk == kKeyMax) continue;
      const uint32_t gidx = key_index(kk);
      const int row = gidx >= (uint32_t)a.n_fresh_global ? a.n_fresh_local + (int)(gidx - a.n_fresh_global)
                                                         : (int)gidx - a.global_offset;
      for (int e = threadIdx.x; e < a.stride; e += blockDim.x)
        a.send_actions[(size_t)j * a.stride + e] = a.actions[(size_t)row * a.stride + e];
    }
  }
}

__global__ void __launch_bounds__(kSelectThreads) merge_refit_kernel(RefitArgs r) {
  extern __shared__ unsigned long long s_keys[];
  merge_refit(r, s_keys);
}

// plain top-k operator (icem_op_topk): indices + costs of the k smallest, ascending (cost, index)
__global__ void __launch_bounds__(kSelectThreads)
topk_finish_kernel(const unsigned long long* keys, int k, const float* costs, int32_t* idx_out, float* cost_out) {
  if (threadIdx.x < k) {
    const uint32_t i = key_index(keys[threadIdx.x]);
    idx_out[threadIdx.x] = (int32_t)i;
    cost_out[threadIdx.x] = costs[i];
  }
}

}  // namespace icem
