// Fused sample -> rollout -> cost kernel for the branch-parallel articulated engine (dyn_chain.cuh):
// a group of G lanes per candidate trajectory, 32 / G trajectories per warp, persistent CTAs.
//
// Replaces, for one CEM iteration (paths relative to /root/reference/icem/), the same reference code as rollout.cuh:
//   controllers/icem.py:61-104   colored-noise sampling, clip, mean injection, shifted elites
//   controllers/mpc.py:56-67     simulate_trajectories -> forward_model.predict_n_steps  (ground-truth model)
//   controllers/abstract_controller.py:74-91  trajectory_cost_fn (sum / best / final)
//
// Per warp and per GROUP OF ROWS (32 / G trajectories):
//   1. the whole warp samples the rows one after the other: unit normals -> colored-noise synthesis -> affine + clip
//      into an [h][d] tile in shared memory (two tiles alternate), each shipped to HBM with ONE 1-D TMA bulk store
//      (cp.async.bulk.global.shared::cta); the tiles alias the dynamics scratch, which is not live yet;
//   2. the lane groups roll their trajectories out side by side: the d controls of step t + 1 are fetched (L2 hits:
//      the rows were written a moment ago; `ld.global.cg`) while step t's substeps run;
//   3. lane 0 of every group writes its trajectory's cost.
// The variant without sampling (given action tiles) skips 1.
#pragma once
#include "common.cuh"
#include "dyn_chain.cuh"
#include "rollout.cuh"

namespace icem {

template <int G>
struct WarpGroupCtx {
  template <int N>
  __device__ __forceinline__ void group_sum(float (&x)[N]) {
    if (G >= 2) {
#pragma unroll
      for (int e = 0; e < N; ++e) x[e] += __shfl_xor_sync(0xffffffffu, x[e], 1);
    }
    if (G >= 4) {
#pragma unroll
      for (int e = 0; e < N; ++e) x[e] += __shfl_xor_sync(0xffffffffu, x[e], 2);
    }
  }
  __device__ __forceinline__ void group_sync() { __syncwarp(); }
};

struct ChainParams {
  const ChainModel* model;   // device global memory
  int act_dim;
  int warp_floats;           // chain_warp_floats(model)
  int rows_per_warp;         // trajectories a warp rolls out side by side: <= 32 / G (fewer = more warps for small N)
};

constexpr int kChainPrefetch = 5;     // controls per lane fetched one step ahead (covers d <= 5 G)

template <int G, bool PLANAR = false>
using ChainWarpLane = ChainLane<WarpGroupCtx<G>, 32 / G, PLANAR>;

// lane-private scratch of lane g of a group: region g behind the group-shared region (see chain_warp_floats)
template <int G, bool PLANAR>
__device__ __forceinline__ void chain_bind_lane(ChainWarpLane<G, PLANAR>& L, const ChainModel& m, float* w_base, int lane,
                                                WarpGroupCtx<G>* ctx) {
  constexpr int NG = 32 / G;
  const int grp = lane / G;
  L.M = &m;
  L.g = lane % G;
  L.sh = w_base + grp;
  L.pr = w_base + m.s_end * NG + L.g * chain_private_region_floats(m.p_end, G) + grp;
  L.ctx = ctx;
}

template <int G, bool PLANAR>
__device__ __forceinline__ float chain_step_cost_pre(const CostConst& cc, const ChainWarpLane<G, PLANAR>& L,
                                                     const typename ChainWarpLane<G, PLANAR>::ChRef& act, int d,
                                                     bool next_obs) {
  const ChainModel& m = *L.M;
  float a2 = 0.f;
  for (int k = 0; k < d; ++k) a2 = fmaf(act[k], act[k], a2);
  const int oo = m.obs_offset;
  if (next_obs) {
    // the part of the locomotion cost known BEFORE the step (rollout.cuh: step_cost)
    const float z = L.state(cc.idx_a + oo);
    const bool z_ok = cc.z_strict ? (z > cc.z_lo && z < cc.z_hi) : (z >= cc.z_lo && z <= cc.z_hi);
    bool ok = true;
    const int n = m.nq + m.nv;
    for (int i = 0; i < n; ++i) {
      const float v = L.state(i);
      ok = ok && (v - v == 0.f);
      if (cc.state_bound > 0.f && i >= cc.idx_b + oo) ok = ok && (i < m.nq ? fabsf(v) : fminf(fabsf(v), 10.f)) < cc.state_bound;
    }
    const float vel = cc.vel_index >= 0 ? cc.w_fwd * L.state(cc.vel_index + oo) : 0.f;
    return ((z_ok && ok) ? 0.f : cc.w_unhealthy) + cc.w_ctrl * a2 - vel;
  }
  if (cc.kind == 3) return reacher_distance(cc, L.state(0), L.state(1), L.state(2), L.state(3));
  if (cc.kind == 4) return 0.f;      // goal-space costs read observations of the batched models only (refused at icem_create)
  if (cc.kind == 0) {   // environments/mujoco.py:67-99
    const float ang = L.state(cc.idx_a + oo), vel = L.state(cc.idx_b + oo);
    float c = 0.f;
    if (cc.penalise_flipping) {
      c += (ang > 1.5707963267948966f) ? 10.f : 0.f;
      c += (ang < -1.5707963267948966f) ? 10.f : 0.f;
    }
    return c + 0.1f * a2 - vel;
  }
  return -L.state(cc.idx_a + oo) + 0.1f * a2;   // environments/mujoco.py:259-277
}

constexpr int kChainMaxWarps = 12;    // per CTA (one CTA per SM): 384 threads x 170 registers fill the register file

template <int G, bool kSample, bool kRollout, bool kNextObs, bool PLANAR = false>
__global__ void __launch_bounds__(kChainMaxWarps * 32, 1)
chain_rollout_kernel(RolloutArgs a, SamplerConst sc, CostConst cc, ChainParams dp) {
  extern __shared__ __align__(128) float smem[];
  constexpr int NG = 32 / G;
  const unsigned long long pr = blockIdx.y;
  const int warps = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int h = sc.h, d = sc.d, hd = h * d;
  const int K2 = 2 * sc.K;
  const int gs = K2 + 1;

  // ---- CTA-shared constants ----
  float* s_model = smem;
  float* s_G = s_model + (((int)(sizeof(ChainModel) / 4) + 3) & ~3);
  float* s_mean = s_G + (kSample ? ((h * gs + 3) & ~3) : 0);
  float* s_std = s_mean + (kSample ? ((hd + 3) & ~3) : 0);
  float* s_low = s_std + (kSample ? ((hd + 3) & ~3) : 0);
  float* s_high = s_low + (kSample ? ((d + 3) & ~3) : 0);
  float* s_warp0 = s_high + (kSample ? ((d + 3) & ~3) : 0);
  const int tile_floats = a.stride;
  const int z_floats = (kSample && !sc.white) ? ((d * gs + 3) & ~3) : 0;
  const int sample_floats = kSample ? 2 * tile_floats + z_floats : 0;
  const int dyn_floats = (dp.warp_floats + 3) & ~3;
  const int warp_floats = sample_floats > dyn_floats ? sample_floats : dyn_floats;
  float* w_base = s_warp0 + (size_t)warp * warp_floats;

  {
    const float* src = reinterpret_cast<const float*>(dp.model);
    for (int i = threadIdx.x; i < (int)(sizeof(ChainModel) / 4); i += blockDim.x) s_model[i] = src[i];
  }
  if (kSample) {
    if (!sc.white)
      for (int i = threadIdx.x; i < h * K2; i += blockDim.x) s_G[(i / K2) * gs + (i % K2)] = sc.G[i];
    const float* gm = a.mean + pr * (unsigned)a.prob_dist;
    const float* gsd = a.std + pr * (unsigned)a.prob_dist;
    for (int i = threadIdx.x; i < hd; i += blockDim.x) { s_mean[i] = gm[i]; s_std[i] = gsd[i]; }
    for (int i = threadIdx.x; i < d; i += blockDim.x) { s_low[i] = sc.low[i]; s_high[i] = sc.high[i]; }
  }
  __syncthreads();
  const ChainModel& m = *reinterpret_cast<const ChainModel*>(s_model);

  const int n_rows = a.n_fresh_local + ((a.iteration == 0 && a.ss->has_prev_elites) ? a.n_shift_local : 0);
  const int rpw = dp.rows_per_warp;                      // rows per warp and trip
  const int n_trips = (n_rows + rpw - 1) / rpw;
  const int wg = blockIdx.x * warps + warp, n_wg = gridDim.x * warps;
  float* actions = a.actions + pr * a.prob_actions;
  const uint32_t tile_bytes = (uint32_t)tile_floats * 4u;

  WarpGroupCtx<G> ctx;
  ChainWarpLane<G, PLANAR> L;
  typedef typename ChainWarpLane<G, PLANAR>::ChRef ChRef;
  const int grp = lane / G;
  chain_bind_lane<G, PLANAR>(L, m, w_base, lane, &ctx);

  for (int trip = wg; trip < n_trips; trip += n_wg) {
    const int row0 = trip * rpw;
    const int rows_here = min(rpw, n_rows - row0);
    if (kSample) {
      // ---- 1. sample the rows of this trip, one bulk store each ----
      float* w_z = w_base + 2 * tile_floats;
      for (int r = 0; r < rows_here; ++r) {
        float* tile = w_base + (r & 1) * tile_floats;
        if (r >= 2) {               // the store that last read this tile buffer (row r - 2) must be done reading
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          __syncwarp();
        }
        sample_row_tile(a, sc, pr, row0 + r, tile, w_z, s_G, s_mean, s_std, s_low, s_high);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_1d(actions + (size_t)(row0 + r) * a.stride, tile, tile_bytes);
          tma_store_commit();
        }
      }
      // all stores complete: the tiles may be overwritten (they alias the dynamics scratch) and the rows are in L2
      if (lane == 0) tma_store_wait_all();
      __syncwarp();
    }
    if (!kRollout) continue;

    // ---- 2. roll the rows out side by side ----
    const bool valid = grp < rows_here;
    const int row = row0 + (valid ? grp : 0);
    const float* arow = actions + (size_t)row * a.stride;
    const float* start = a.start_state + pr * (unsigned)a.prob_state;
    const int ns = m.nq + m.nv;
    for (int i = L.g; i < ns; i += G) L.state(i) = start[i];
    for (int k = L.g; k < d; k += G) L.sh[(m.s_ctrl + k) * NG] = __ldcg(arow + k);
    __syncwarp();
    float total = (cc.reduce == 1) ? INFINITY : 0.f;
    const bool direct = d > kChainPrefetch * G;         // wide action vectors: no register prefetch
    for (int t = 0; t < h; ++t) {
      const ChRef act{L.sh + m.s_ctrl * NG};
      float pre[kChainPrefetch];
      const bool more = t + 1 < h;
      if (more && !direct) {          // the next step's controls travel from L2 while this step's substeps run
#pragma unroll
        for (int u = 0; u < kChainPrefetch; ++u) {
          const int k = L.g + u * G;
          pre[u] = k < d ? __ldcg(arow + (t + 1) * d + k) : 0.f;
        }
      }
      const float c = chain_step_cost_pre<G, PLANAR>(cc, L, act, d, kNextObs);
      if constexpr (!kNextObs) {
        if (cc.reduce == 0) total += c;
        else if (cc.reduce == 1) total = fminf(total, c);
        else total = c;
        if (more) L.step(act);
      } else {
        const float x0 = L.state(m.obs_offset);
        L.step(act);
        const float cf = c - (L.state(m.obs_offset) - x0) * cc.inv_dt;
        if (cc.reduce == 0) total += cf;
        else if (cc.reduce == 1) total = fminf(total, cf);
        else total = cf;
      }
      if (more) {                     // every lane is past its last read of this step's controls (group_sync in step)
        if (!direct) {
#pragma unroll
          for (int u = 0; u < kChainPrefetch; ++u) {
            const int k = L.g + u * G;
            if (k < d) L.sh[(m.s_ctrl + k) * NG] = pre[u];
          }
        } else {
          for (int k = L.g; k < d; k += G) L.sh[(m.s_ctrl + k) * NG] = __ldcg(arow + (t + 1) * d + k);
        }
      }
      __syncwarp();
    }
    if (valid && L.g == 0) (a.costs + pr * a.prob_costs)[row] = total;
    __syncwarp();
  }
}

// shared-memory footprint of chain_rollout_kernel for `warps` warps per CTA
template <bool kSample>
inline size_t chain_rollout_smem_bytes(const SamplerConst& sc, const ChainParams& dp, int stride, int warps) {
  const int h = sc.h, d = sc.d, hd = h * d, K2 = 2 * sc.K, gs = K2 + 1;
  size_t f = (((int)(sizeof(ChainModel) / 4) + 3) & ~3);
  if (kSample) f += ((h * gs + 3) & ~3) + 2 * ((hd + 3) & ~3) + 2 * ((d + 3) & ~3);
  const int z_floats = (kSample && !sc.white) ? ((d * gs + 3) & ~3) : 0;
  const int sample_floats = kSample ? 2 * stride + z_floats : 0;
  const int dyn_floats = (dp.warp_floats + 3) & ~3;
  return (f + (size_t)warps * (sample_floats > dyn_floats ? sample_floats : dyn_floats)) * sizeof(float);
}

// One transition (or just the observation) of a single state: env.step on the device model with the chain engine.
// blockIdx.x = instance; one warp, every group computes the same thing, group 0 reports.
template <int G, bool PLANAR = false>
__global__ void chain_advance_kernel(ChainParams dp, float* state, const float* action, float* next_state,
                                     float* obs_out, int obs_dim, int state_stride, int action_stride, int obs_stride) {
  extern __shared__ __align__(128) float smem[];
  constexpr int NG = 32 / G;
  state += (size_t)blockIdx.x * state_stride;
  if (action) action += (size_t)blockIdx.x * action_stride;
  if (next_state) next_state += (size_t)blockIdx.x * state_stride;
  if (obs_out) obs_out += (size_t)blockIdx.x * obs_stride;
  float* s_model = smem;
  float* w_base = s_model + (((int)(sizeof(ChainModel) / 4) + 3) & ~3);
  {
    const float* src = reinterpret_cast<const float*>(dp.model);
    for (int i = threadIdx.x; i < (int)(sizeof(ChainModel) / 4); i += blockDim.x) s_model[i] = src[i];
  }
  __syncthreads();
  const ChainModel& m = *reinterpret_cast<const ChainModel*>(s_model);
  const int lane = threadIdx.x & 31, grp = lane / G;
  WarpGroupCtx<G> ctx;
  ChainWarpLane<G, PLANAR> L;
  typedef typename ChainWarpLane<G, PLANAR>::ChRef ChRef;
  chain_bind_lane<G, PLANAR>(L, m, w_base, lane, &ctx);
  const int ns = m.nq + m.nv;
  for (int i = L.g; i < ns; i += G) L.state(i) = state[i];
  if (action)
    for (int k = L.g; k < dp.act_dim; k += G) L.sh[(m.s_ctrl + k) * NG] = action[k];
  __syncwarp();
  if (action) L.step(ChRef{L.sh + m.s_ctrl * NG});
  __syncwarp();
  if (grp == 0) {
    if (next_state)
      for (int i = L.g; i < ns; i += G) next_state[i] = L.state(i);
    if (obs_out)
      for (int i = L.g; i < obs_dim; i += G) obs_out[i] = L.state(i + m.obs_offset);
  }
}

}  // namespace icem
