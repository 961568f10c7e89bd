// Fused sample -> rollout -> cost kernel (one warp per candidate trajectory, persistent CTAs).
//
// Replaces, for one CEM iteration (paths relative to /root/reference/icem/):
//   controllers/icem.py:61-104   colored-noise sampling, clip, mean injection, shifted elites
//   controllers/mpc.py:56-67     simulate_trajectories -> forward_model.predict_n_steps
//   controllers/abstract_controller.py:74-91  trajectory_cost_fn (sum / best / final)
//
// Data movement: each warp builds its trajectory's action tile [h][d] in shared memory, rolls the
// forward model out of it in place, and ships the tile to HBM with ONE 1-D TMA bulk store
// (cp.async.bulk.global.shared::cta) that overlaps the rollout.  The `kSample == false` variant
// pulls given action tiles from HBM with TMA bulk loads (mbarrier complete_tx), double-buffered.
#pragma once
#include "common.cuh"

namespace icem {

// dynamic, per-plan-step values; lives in device memory so one CUDA graph serves every step
struct StepState {
  uint32_t step;            // plan-step counter since beginning_of_rollout (Philox counter word)
  int32_t has_prev_elites;  // elite_samples non-empty -> shifted elites join iteration 0
  int32_t inject;           // parity mode: read unit normals from HBM instead of Philox
  uint32_t plans_total;     // plan steps since icem_create (MpcRandom's sample() counter never resets)
};

struct SamplerConst {
  int h, d, K;          // horizon, action dim, rFFT bins (h/2+1)
  int white;            // noise_beta == 0: iid normal, z laid out [row][h][d]
  int trunc;            // MpcCemStd: draws are signed tail probabilities [row][h][d] (truncnorm_ppf), action = mean + std * ppf
  int levine;           // MpcCemStd bounds_like_levine: lower/upper = -2/+2 (else the action bounds in std units)
  int rnd_freq;         // MpcRandom: >= 0 = action_change_frequency (uniform piecewise-constant actions), -1 = off
  int n_global;         // MpcRandom: population size (sample() calls per plan step = n_global * h)
  uint32_t magic_d, magic_K;   // ceil(2^32 / d), ceil(2^32 / K): x / d == __umulhi(x, magic_d) for x < 2^16
  const float* G;       // [h][2K] synthesis matrix (colored): y[t] = sum_j G[t][j] * z[j], z = [zr(K), zi(K)]
  const float* low;     // [d]
  const float* high;    // [d]
};

struct CostConst {
  int kind;             // ICEM_COST_*
  int reduce;           // ICEM_REDUCE_*
  int idx_a, idx_b;     // cheetah: (root angle, x velocity) observation indices; humanoid: (root z, -);
                        // locomotion: (z index, first index of the bounded state entries)
  int penalise_flipping;
  // ICEM_COST_LOCOMOTION (Hopper / Ant, environments/mujoco.py:153-176, 196-231):
  //   -(x' - x) / dt + w_unhealthy * unhealthy(obs) + w_ctrl * |a|^2,   x = obs[0] before / after the step
  float inv_dt, w_ctrl, w_unhealthy, z_lo, z_hi, state_bound;
  int z_strict;         // Hopper: z_lo < z < z_hi; Ant: z_lo <= z <= z_hi
  int vel_index;        // >= 0: x velocity = obs[vel_index] (Humanoid, mujoco.py:333; inv_dt is 0 then); -1: finite difference
  float w_fwd;          // weight of the velocity term
  float reach[4];       // ICEM_COST_REACHER: link lengths l1, l2 and the target body's rest position (x, y)
  int goal_idx, ach_idx, goal_sparse, goal_shaped;      // ICEM_COST_GOAL_DISTANCE (see include/icem_b200.h)
  float goal_threshold;
};

// environments/abstract_environments.py:115-123 / environments/robotics.py:150-164 from the two squared distances
__device__ __forceinline__ float goal_distance_cost(const CostConst& cc, float d2, float e2) {
  const float d = sqrtf(d2), e = cc.goal_shaped ? sqrtf(e2) : 0.f;
  if (cc.goal_sparse) return (d > cc.goal_threshold ? 1.f : 0.f) + ((cc.goal_shaped && e > cc.goal_threshold) ? 0.1f : 0.f);
  return d + 0.1f * e;
}

// environments/mujoco.py:366-368 (Reacher): |fingertip - target| from the state (q0, q1 arm hinges about z; q2, q3 the
// target's slide joints): fingertip = l1 (cos q0, sin q0) + l2 (cos(q0 + q1), sin(q0 + q1)), target = rest + (q2, q3);
// both bodies sit at the same height, so the z difference gym's observation carries is zero.
__device__ __forceinline__ float reacher_distance(const CostConst& cc, float q0, float q1, float q2, float q3) {
  float s0, c0, s01, c01;
  sincosf(q0, &s0, &c0);
  sincosf(q0 + q1, &s01, &c01);
  const float dx = fmaf(cc.reach[0], c0, cc.reach[1] * c01) - (cc.reach[2] + q2);
  const float dy = fmaf(cc.reach[0], s0, cc.reach[1] * s01) - (cc.reach[3] + q3);
  return sqrtf(fmaf(dx, dx, dy * dy));
}

struct RolloutArgs {
  int n_fresh_local;        // fresh rows this rank samples
  int n_shift_local;        // shifted-elite rows this rank simulates when StepState.has_prev_elites (iteration 0)
  int global_offset;        // global trajectory index of local row 0
  int n_fresh_global;       // N_i: global index of the first shifted-elite row
  int iteration;
  int inject_mean_row0;     // last iteration && use_mean_actions: global row 0 <- mean (icem.py:87-88)
  int stride;               // floats per trajectory row in `actions` (16-B multiple)
  float* actions;           // [rows][stride]
  float* costs;             // [rows]
  const float* mean;        // [h*d]
  const float* std;         // [h*d]
  const float* prev_elites; // [k][stride] elites of the previous plan step, best first
  const float* start_state; // [state_dim] fp32
  const float* inj_zr;      // parity mode draws for this iteration (device), rows as in icem_inject_noise
  const float* inj_zi;
  const StepState* ss;
  uint32_t seed_lo, seed_hi;
  // several independent MPC problems in one launch (blockIdx.y = problem): element strides between consecutive
  // problems' buffers; problem i draws with seed + i.  All zero / unused for a single problem.
  unsigned long long prob_actions, prob_costs;
  int prob_dist, prob_elites, prob_state;
};

// ---------------------------------------------------------------------------------------------
// cost of one step on the PRE-action observation
template <class Dyn, bool kNextObs>
__device__ __forceinline__ float step_cost(const CostConst& cc, const Dyn& dyn, const float* act, int d) {
  float a2 = 0.f;
  for (int m = 0; m < d; ++m) a2 = fmaf(act[m], act[m], a2);   // smem broadcast reads, all lanes redundantly
  if constexpr (kNextObs) {
    // the part of the locomotion cost known BEFORE the step (the x-velocity term needs next_obs, added by the caller)
    const float z = dyn.obs(cc.idx_a);
    const bool z_ok = cc.z_strict ? (z > cc.z_lo && z < cc.z_hi) : (z >= cc.z_lo && z <= cc.z_hi);
    const bool healthy = z_ok && dyn.state_healthy(cc.idx_b, cc.state_bound);
    const float vel = cc.vel_index >= 0 ? cc.w_fwd * dyn.obs(cc.vel_index) : 0.f;
    return (healthy ? 0.f : cc.w_unhealthy) + cc.w_ctrl * a2 - vel;
  }
  if (cc.kind == 3) return reacher_distance(cc, dyn.obs(0), dyn.obs(1), dyn.obs(2), dyn.obs(3));
  if (cc.kind == 4) {
    float d2 = 0.f, e2 = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float dk = dyn.obs(cc.goal_idx + k) - dyn.obs(cc.ach_idx + k);
      d2 = fmaf(dk, dk, d2);
      if (cc.goal_shaped) { const float ek = dyn.obs(k) - dyn.obs(3 + k); e2 = fmaf(ek, ek, e2); }
    }
    return goal_distance_cost(cc, d2, e2);
  }
  if (cc.kind == 0) {   // environments/mujoco.py:67-99
    const float ang = dyn.obs(cc.idx_a), vel = dyn.obs(cc.idx_b);
    float c = 0.f;
    if (cc.penalise_flipping) {
      c += (ang > 1.5707963267948966f) ? 10.f : 0.f;
      c += (ang < -1.5707963267948966f) ? 10.f : 0.f;
    }
    return c + 0.1f * a2 - vel;
  }
  // environments/mujoco.py:259-277
  return -dyn.obs(cc.idx_a) + 0.1f * a2;
}

// scipy.stats.truncnorm._ppf (what controllers/mpc.py:194-198 draws through) for fp32: the uniform draw arrives as
// a SIGNED TAIL PROBABILITY w -- u = w for u < 1/2, w = -(1 - u) otherwise -- so that the small tail keeps its
// relative precision (a plain fp32 u has 6e-8 absolute resolution near 1, i.e. 2e-4 sigma at +3.8 sigma); the
// quantile is taken from whichever tail of the truncated distribution is lighter, like scipy's left/right cases.
__device__ __forceinline__ float truncnorm_ppf(float w, float a, float b) {
  const float pa = normcdff(a), pnb = normcdff(-b);          // P(X < a), P(X > b)
  const float mass = (a < 0.f && b > 0.f) ? 1.f - pa - pnb : (b <= 0.f ? normcdff(b) - pa : normcdff(-a) - pnb);
  const float ul = w >= 0.f ? w : 1.f + w, ur = w >= 0.f ? 1.f - w : -w;
  const float pl = fmaf(ul, mass, pa), pr = fmaf(ur, mass, pnb);
  const float x = pl <= pr ? normcdfinvf(pl) : -normcdfinvf(pr);
  return fminf(fmaxf(x, a), b);
}

// MpcRandom (mpc.py:95-107): sample() is called row by row, step by step; the action drawn at construction serves
// the first `freq` calls, every later draw serves freq + 1 calls.  Call index -> segment -> Philox counter, so the
// result does not depend on how rows are spread over ranks.  Kept out of line: a rare mode must not perturb the
// register allocation / code layout of the fused kernel's hot rollout loop (measured: 1.3 % when inlined).
__device__ __noinline__ float random_shooting_uniform(uint32_t plans_total, int n_global, uint32_t grow, int h, int t,
                                                      int dim, int freq, uint32_t seed_lo, uint32_t seed_hi,
                                                      uint32_t problem) {
  const unsigned long long seed = (((unsigned long long)seed_hi << 32) | seed_lo) + problem;
  seed_lo = (uint32_t)seed;
  seed_hi = (uint32_t)(seed >> 32);
  const unsigned long long call = ((unsigned long long)plans_total * (unsigned long long)n_global
                                   + (unsigned long long)grow) * (unsigned long long)h + (unsigned long long)t;
  const unsigned long long f = (unsigned long long)freq;
  const unsigned long long seg = call < f ? 0ull : 1ull + (call - f) / (f + 1ull);
  const Philox4 r = philox4x32_10((uint32_t)seg, (uint32_t)(seg >> 32), (uint32_t)(dim >> 2), 0x524E4431u,
                                  seed_lo, seed_hi);
  const uint32_t w = (dim & 3) == 0 ? r.x : (dim & 3) == 1 ? r.y : (dim & 3) == 2 ? r.z : r.w;
  return ((float)w + 0.5f) * 2.3283064365386963e-10f;
}

// ---------------------------------------------------------------------------------------------
// unit normals for one trajectory row -> z[dim][j] (colored, row stride zs) or straight into the
// tile (white).  Philox counter = (global row, block, step, iteration); key = seed.
__device__ __forceinline__ void fill_normals(float* dst, int count, uint32_t grow, const RolloutArgs& a,
                                             uint32_t problem, uint32_t step, int K2, int zs, bool white, bool uniform,
                                             uint32_t magic_K) {
  const int lane = lane_id();
  for (int b = lane; b * 4 < count; b += 32) {
    const unsigned long long seed = (((unsigned long long)a.seed_hi << 32) | a.seed_lo) + problem;   // problem i: seed + i
    Philox4 r = philox4x32_10(grow, (uint32_t)b, step, (uint32_t)a.iteration, (uint32_t)seed, (uint32_t)(seed >> 32));
    float n[4];
    if (uniform) {                       // signed tail probabilities in (-1/2, 1/2) \ {0} (see truncnorm_ppf)
      const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float tail = ((float)(rr[q] & 0x7FFFFFFFu) + 0.5f) * 2.3283064365386963e-10f;   // (0, 1/2)
        n[q] = (rr[q] >> 31) ? -tail : tail;
      }
    } else {
      box_muller(r.x, r.y, n[0], n[1]);
      box_muller(r.z, r.w, n[2], n[3]);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int c = b * 4 + q;
      if (c < count) {
        if (white) {
          dst[c] = n[q];
        } else {           // c = part*(d*K) + dim*K + k  ->  z[dim][part*K + k]
          const int K = K2 >> 1;
          const int dk = count >> 1;            // d*K
          const int part = c >= dk;
          const int r2 = c - part * dk;
          const int dim = (int)__umulhi((uint32_t)r2, magic_K);
          const int k = r2 - dim * K;
          dst[dim * zs + part * K + k] = n[q];
        }
      }
    }
  }
}

// One trajectory row of the population -> its action tile [h][d] in shared memory (the whole warp works on the row):
// unit normals (Philox, or the injected parity draws), colored-noise synthesis + affine + clip, mean row, shifted
// elites; the truncated-normal (MpcCemStd) and uniform (MpcRandom) samplers.  `z` is per-warp scratch of d * (2K + 1)
// floats (unused for white noise); G / mean / std / low / high are the CTA's shared-memory copies.
__device__ __forceinline__ void sample_row_tile(const RolloutArgs& a, const SamplerConst& sc, unsigned long long pr,
                                                int row, float* tile, float* w_z, const float* s_G,
                                                const float* s_mean, const float* s_std, const float* s_low,
                                                const float* s_high) {
  const int lane = threadIdx.x & 31;
  const int h = sc.h, d = sc.d, hd = h * d;
  const int K2 = 2 * sc.K, gs = K2 + 1, zs = K2 + 1;
  const int tile_floats = a.stride;
  // re-read per row (an L1 hit) rather than held in registers across the rollout loop
  const uint32_t ss_step = a.ss->step;
  const bool ss_inject = a.ss->inject != 0;
  const bool shifted = row >= a.n_fresh_local;
  // global trajectory index: fresh rows are contiguous per rank, shifted rows follow N_i
  const uint32_t grow = shifted ? (uint32_t)(a.n_fresh_global + (row - a.n_fresh_local))
                                : (uint32_t)(a.global_offset + row);
        // ---- 1. unit normals ----
        float* zdst = sc.white ? tile : w_z;
        const int count = sc.white ? hd : d * K2;
        if (ss_inject) {
          if (sc.white) {
            const float* src = a.inj_zr + (size_t)row * hd;
            for (int i = lane; i < hd; i += 32) tile[i] = src[i];
          } else {
            const int dK = d * sc.K;
            const float* sr = a.inj_zr + (size_t)row * dK;
            const float* si = a.inj_zi + (size_t)row * dK;
            for (int i = lane; i < dK; i += 32) {
              const int dim = (int)__umulhi((uint32_t)i, sc.magic_K), k = i - dim * sc.K;
              w_z[dim * zs + k] = sr[i];
              w_z[dim * zs + sc.K + k] = si[i];
            }
          }
        } else if (sc.rnd_freq < 0) {
          fill_normals(zdst, count, grow, a, (uint32_t)pr, ss_step, K2, zs, sc.white, sc.trunc != 0, sc.magic_K);
        }
        __syncwarp();
        // ---- 2. synthesis + affine + clip (icem.py:73-79), elite shift (icem.py:91-104), mean row ----
        const bool mean_row = a.inject_mean_row0 && grow == 0u && !shifted;
        const float* elite = shifted ? (a.prev_elites + pr * (unsigned)a.prob_elites) + (size_t)(row - a.n_fresh_local) * a.stride : nullptr;
        for (int o = lane; o < hd; o += 32) {
          const int t = (int)__umulhi((uint32_t)o, sc.magic_d), dim = o - t * d;
          if (sc.rnd_freq >= 0) {      // MpcRandom: piecewise-constant uniform actions
            const float u = ss_inject ? tile[o]
                                      : random_shooting_uniform(a.ss->plans_total, sc.n_global, grow, h, t, dim, sc.rnd_freq,
                                                                a.seed_lo, a.seed_hi, (uint32_t)pr);
            tile[o] = fminf(fmaf(s_high[dim] - s_low[dim], u, s_low[dim]), s_high[dim]);   // Box.sample (gym)
            continue;
          }
          if (sc.trunc) {       // MpcCemStd: mean + std * truncnorm.ppf(u; lower, upper), no clip (mpc.py:194-198, 290-301)
            const float sd = s_std[o], mu = s_mean[o];
            const float lo_ = sc.levine ? -2.f : (s_low[dim] - mu) / (sd + 1e-8f);
            const float hi_ = sc.levine ? 2.f : (s_high[dim] - mu) / (sd + 1e-8f);
            tile[o] = fmaf(sd, truncnorm_ppf(tile[o], lo_, hi_), mu);
            continue;
          }
          float y;
          if (sc.white) {
            y = tile[o];
          } else {
            const float* g = s_G + t * gs;
            const float* z = w_z + dim * zs;
            float acc = 0.f;
  #pragma unroll 8
            for (int j = 0; j < K2; ++j) acc = fmaf(g[j], z[j], acc);
            y = acc;
          }
          float v = fminf(fmaxf(fmaf(y, s_std[o], s_mean[o]), s_low[dim]), s_high[dim]);
          if (mean_row) v = s_mean[o];
          if (shifted && t < h - 1) v = elite[o + d];
          tile[o] = v;
        }
        for (int o = hd + lane; o < tile_floats; o += 32) tile[o] = 0.f;   // row padding
        __syncwarp();
}

// The rollout of ONE trajectory, out of line on purpose: across the call nothing of the kernel's outer loops
// (sampler pointers, row bookkeeping, TMA state) stays in registers, so the dynamics code gets the whole register
// budget of 80.  Measured on B200: HumanoidStandup 34.5 -> 31.6 ms per plan step (+9 %), HalfCheetah +3 %.
template <class Dyn, bool kNextObs>
__device__ __forceinline__ float rollout_one_body(Dyn& dyn, const CostConst& cc, const float* tile, float* w_stash, int h,
                                                  int d, int barrier_threads) {
  const int lane = threadIdx.x & 31;
  float total = (cc.reduce == 1) ? INFINITY : 0.f;
  for (int t = 0; t < h; ++t) {
    if (Dyn::kCtaLockstep) asm volatile("bar.sync 1, %0;" :: "r"(barrier_threads) : "memory");
    const float* act = tile + t * d;
    const float c = step_cost<Dyn, kNextObs>(cc, dyn, act, d);
    if constexpr (!kNextObs) {
      if (cc.reduce == 0) total += c;
      else if (cc.reduce == 1) total = fminf(total, c);
      else total = c;
      if (t + 1 < h) dyn.step(act);
    } else {
      if (lane == 0) { w_stash[0] = dyn.obs(0); w_stash[1] = c; }
      dyn.step(act);
      __syncwarp();
      const float cf = w_stash[1] - (dyn.obs(0) - w_stash[0]) * cc.inv_dt;
      __syncwarp();
      if (cc.reduce == 0) total += cf;
      else if (cc.reduce == 1) total = fminf(total, cf);
      else total = cf;
    }
  }
  return total;
}

template <class Dyn, bool kNextObs>
__device__ __noinline__ float rollout_one(Dyn& dyn, const CostConst& cc, const float* tile, float* w_stash, int h, int d,
                                          int barrier_threads) {
  return rollout_one_body<Dyn, kNextObs>(dyn, cc, tile, w_stash, h, d, barrier_threads);
}

// kNextObs: the cost reads next_obs (ICEM_COST_LOCOMOTION).  A template flag, not a runtime branch: the fused kernel's
// rollout loop is sensitive to anything that changes its register allocation (measured: +2 % per plan step with the
// branch compiled into the common instantiation).
template <class Dyn, bool kSample, bool kRollout, bool kNextObs = false>
__global__ void __launch_bounds__(Dyn::kWarpsPerCta * 32, Dyn::kMinCtasPerSm)
rollout_kernel(RolloutArgs a, SamplerConst sc, CostConst cc, typename Dyn::Params dp) {
  extern __shared__ __align__(128) float smem[];
  // Several problems in one launch: this CTA works on problem blockIdx.y.  The kernel parameters stay untouched
  // in the constant bank and every per-problem pointer is formed where it is used (all outside the rollout loop):
  // rebasing `a` itself would pin the rebased pointers in registers across the hot loop (measured: -1.5 %).
  const unsigned long long pr = blockIdx.y;
#define ICEM_P_ACTIONS (a.actions + pr * a.prob_actions)
#define ICEM_P_COSTS (a.costs + pr * a.prob_costs)
#define ICEM_P_MEAN (a.mean + pr * (unsigned)a.prob_dist)
#define ICEM_P_STD (a.std + pr * (unsigned)a.prob_dist)
#define ICEM_P_ELITES (a.prev_elites + pr * (unsigned)a.prob_elites)
#define ICEM_P_STATE (a.start_state + pr * (unsigned)a.prob_state)
  const int warps = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int h = sc.h, d = sc.d, hd = h * d;
  const int K2 = 2 * sc.K;
  const int gs = K2 + 1;               // padded row strides (bank-conflict-free)
  const int zs = K2 + 1;

  // ---- CTA-shared constants ----
  float* s_G = smem;                                   // [h][gs]
  float* s_mean = s_G + ((h * gs + 3) & ~3);           // [hd]
  float* s_std = s_mean + ((hd + 3) & ~3);             // [hd]
  float* s_low = s_std + ((hd + 3) & ~3);              // [d]
  float* s_high = s_low + ((d + 3) & ~3);              // [d]
  float* s_dyn = s_high + ((d + 3) & ~3);              // Dyn CTA constants
  float* s_warp0 = s_dyn + ((Dyn::cta_floats(dp) + 3) & ~3);
  // ---- per-warp ----
  const int tile_floats = a.stride;                    // 16-B multiple
  // the sampler's normal-draw scratch (z) is dead once the tile is built: it aliases the dynamics scratch
  const int z_floats = kSample ? ((sc.white ? 0 : d * zs) + 3) & ~3 : 0;
  const int dyn_floats = (Dyn::warp_floats(dp) + 3) & ~3;
  const int scratch_floats = z_floats > dyn_floats ? z_floats : dyn_floats;
  const int ntile = kSample ? 1 : 2;                   // loads are double-buffered
  const int warp_floats = ntile * tile_floats + scratch_floats + 8;
  float* w_base = s_warp0 + (size_t)warp * warp_floats;
  float* w_tile = w_base;
  float* w_z = w_tile + ntile * tile_floats;
  float* w_dyn = w_z;
  uint64_t* w_bar = reinterpret_cast<uint64_t*>(w_dyn + scratch_floats);   // 2 x 8 B
  float* w_stash = reinterpret_cast<float*>(w_bar + 2);                    // 2 floats kept across a step (locomotion cost)

  if (kSample && !sc.white)
    for (int i = threadIdx.x; i < h * K2; i += blockDim.x) s_G[(i / K2) * gs + (i % K2)] = sc.G[i];
  for (int i = threadIdx.x; i < hd; i += blockDim.x) {
    s_mean[i] = ICEM_P_MEAN[i];
    s_std[i] = ICEM_P_STD[i];
  }
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    s_low[i] = sc.low[i];
    s_high[i] = sc.high[i];
  }
  Dyn::cta_init(dp, s_dyn);
  if (!kSample && lane == 0) {
    mbar_init(&w_bar[0], 1);
    mbar_init(&w_bar[1], 1);
    mbar_fence_init();
  }
  __syncthreads();

  const int n_rows = a.n_fresh_local + ((a.iteration == 0 && a.ss->has_prev_elites) ? a.n_shift_local : 0);
  // rows are dealt to CTAs in equal contiguous blocks (every SM gets the same number of trajectories +-1), and
  // round-robin to the warps inside a CTA
  const int cta_lo = (int)((long long)n_rows * blockIdx.x / gridDim.x);
  const int cta_hi = (int)((long long)n_rows * (blockIdx.x + 1) / gridDim.x);
  const uint32_t tile_bytes = (uint32_t)tile_floats * 4u;
  Dyn dyn;
  dyn.bind(dp, s_dyn, w_dyn);

  if (!kSample) {   // prologue of the TMA load pipeline
    if (cta_lo + warp < cta_hi && lane == 0) {
      mbar_expect_tx(&w_bar[0], tile_bytes);
      tma_load_1d(w_tile, ICEM_P_ACTIONS + (size_t)(cta_lo + warp) * a.stride, tile_bytes, &w_bar[0]);
    }
  }

  // Dyn::kCtaLockstep: every warp of the CTA makes the same number of trips and meets at a CTA barrier before each
  // control step, so the 8 warps run the (large, fully unrolled) dynamics code at the same time and share
  // instruction-cache lines; warps whose row is past the end only take part in the barriers.
  int it = 0;
  for (int base = cta_lo; base < cta_hi; base += warps, ++it) {
    const int row = base + warp;
    const bool active = row < cta_hi;
    // Warps whose row is past the end skip the round.  The lockstep barrier below is a NAMED barrier counted over
    // the warps that do have a row this round (warp-uniform, the same for every participant), so no thread ever
    // waits at a barrier that others do not reach (compute-sanitizer synccheck clean).
    const int n_active = min(warps, cta_hi - base);
    if (!active) {
      if (!Dyn::kCtaLockstep) break;
      continue;
    }
    float* tile = w_tile;
    if (kSample) {
      // the previous trajectory's bulk store must have finished READING the tile
      if (lane == 0) tma_store_wait_read();
      __syncwarp();
      sample_row_tile(a, sc, pr, row, tile, w_z, s_G, s_mean, s_std, s_low, s_high);
      // ---- 3. ship the tile: one TMA bulk store, overlapped with the rollout ----
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_1d(ICEM_P_ACTIONS + (size_t)row * a.stride, tile, tile_bytes);
        tma_store_commit();
      }
    } else {
      // consume buffer (it&1); prefetch the next row into the other buffer
      const int buf = it & 1;
      tile = w_tile + buf * tile_floats;
      const int nxt = row + warps;
      if (nxt < cta_hi && lane == 0) {
        mbar_expect_tx(&w_bar[buf ^ 1], tile_bytes);
        tma_load_1d(w_tile + (buf ^ 1) * tile_floats, ICEM_P_ACTIONS + (size_t)nxt * a.stride, tile_bytes,
                    &w_bar[buf ^ 1]);
      }
      mbar_wait(&w_bar[buf], (uint32_t)((it >> 1) & 1));
    }

    if (kRollout) {
      // ---- 4. open-loop rollout from the shared start state, cost on the pre-action observation ----
      dyn.reset(ICEM_P_STATE);
      // out of line only where the dynamics is big enough to need the registers (Dyn::kOutlineRollout); a cheap model
      // (dense layer) loses 30 % to the call and the by-reference state
      float total;
      if constexpr (Dyn::kOutlineRollout) total = rollout_one<Dyn, kNextObs>(dyn, cc, tile, w_stash, h, d, n_active * 32);
      else total = rollout_one_body<Dyn, kNextObs>(dyn, cc, tile, w_stash, h, d, n_active * 32);
      if (lane == 0) ICEM_P_COSTS[row] = total;
    }
    __syncwarp();
  }
  if (kSample && lane == 0) tma_store_wait_all();
#undef ICEM_P_ACTIONS
#undef ICEM_P_COSTS
#undef ICEM_P_MEAN
#undef ICEM_P_STD
#undef ICEM_P_ELITES
#undef ICEM_P_STATE
}

// shared-memory footprint of rollout_kernel for `warps` warps per CTA
template <class Dyn, bool kSample>
inline size_t rollout_smem_bytes(const SamplerConst& sc, const typename Dyn::Params& dp, int stride, int warps) {
  const int h = sc.h, d = sc.d, hd = h * d, K2 = 2 * sc.K, gs = K2 + 1;
  size_t f = ((h * gs + 3) & ~3) + 2 * ((hd + 3) & ~3) + 2 * ((d + 3) & ~3) + ((Dyn::cta_floats(dp) + 3) & ~3);
  const int z_floats = kSample ? ((sc.white ? 0 : d * gs) + 3) & ~3 : 0;
  const int dyn_floats = (Dyn::warp_floats(dp) + 3) & ~3;
  const int ntile = kSample ? 1 : 2;
  const size_t warp_floats = (size_t)ntile * stride + (z_floats > dyn_floats ? z_floats : dyn_floats) + 8;
  return (f + warps * warp_floats) * sizeof(float);
}

}  // namespace icem
