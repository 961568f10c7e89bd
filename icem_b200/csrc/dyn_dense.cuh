// Dense single-layer forward model obs' = tanh(W_o obs + W_a act + b), warp-cooperative.
// The batched-model path of the reference: ForwardModelWithDefaults.predict_n_steps
// (icem/models/abstract_models.py:17-53) with a dense `predict`.  Lane j owns output row j.
#pragma once
#include "common.cuh"

namespace icem {

struct DenseTanh {
  static constexpr int kWarpsPerCta = 8;
  static constexpr int kMinCtasPerSm = 2;
  static constexpr bool kCtaLockstep = false;
  static constexpr bool kOutlineRollout = false;
  static constexpr bool kHasHealth = false;
  struct Params {
    int obs_dim, act_dim;
    const float* w_obs;   // [obs_dim][obs_dim] row-major (device)
    const float* w_act;   // [obs_dim][act_dim]
    const float* bias;    // [obs_dim]
  };
  __host__ __device__ static int ld_obs(const Params& p) { return p.obs_dim | 1; }   // odd stride: no bank conflicts
  __host__ __device__ static int ld_act(const Params& p) { return p.act_dim | 1; }
  __host__ __device__ static int cta_floats(const Params& p) {
    return p.obs_dim * ld_obs(p) + p.obs_dim * ld_act(p) + p.obs_dim;
  }
  __host__ __device__ static int warp_floats(const Params& p) { return 2 * p.obs_dim; }
  __host__ __device__ static int state_dim(const Params& p) { return p.obs_dim; }

  __device__ static void cta_init(const Params& p, float* s) {
    const int lo = ld_obs(p), la = ld_act(p), n = p.obs_dim;
    float* wo = s;
    float* wa = wo + n * lo;
    float* b = wa + n * la;
    for (int i = threadIdx.x; i < n * n; i += blockDim.x) wo[(i / n) * lo + (i % n)] = p.w_obs[i];
    for (int i = threadIdx.x; i < n * p.act_dim; i += blockDim.x)
      wa[(i / p.act_dim) * la + (i % p.act_dim)] = p.w_act[i];
    for (int i = threadIdx.x; i < n; i += blockDim.x) b[i] = p.bias[i];
  }

  int n, m, lo, la;
  const float *wo, *wa, *b;
  float *cur, *nxt;

  __device__ void bind(const Params& p, const float* cta, float* warp) {
    n = p.obs_dim; m = p.act_dim; lo = ld_obs(p); la = ld_act(p);
    wo = cta; wa = wo + n * lo; b = wa + n * la;
    cur = warp; nxt = warp + n;
  }
  __device__ void reset(const float* start_state) {
    for (int i = lane_id(); i < n; i += 32) cur[i] = start_state[i];
    __syncwarp();
  }
  __device__ float obs(int i) const { return cur[i]; }
  __device__ bool state_healthy(int, float) const { return true; }
  __device__ void step(const float* act) {
    for (int j = lane_id(); j < n; j += 32) {
      float acc = b[j];
      const float* r = wo + j * lo;
      for (int i = 0; i < n; ++i) acc = fmaf(r[i], cur[i], acc);
      const float* q = wa + j * la;
      for (int i = 0; i < m; ++i) acc = fmaf(q[i], act[i], acc);
      nxt[j] = tanhf(acc);
    }
    __syncwarp();
    float* t = cur; cur = nxt; nxt = t;
  }
  __device__ void export_state(float* out) const {
    for (int i = lane_id(); i < n; i += 32) out[i] = cur[i];
  }
};

}  // namespace icem
