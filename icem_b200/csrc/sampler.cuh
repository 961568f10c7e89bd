// Colored-noise action sampler, one THREAD per (trajectory, action dim) series -- the stand-alone K1 kernel
// (sample -> HBM), used when the rollout runs in its own kernel (tensor-core MLP model) and by icem_op_sample.
//
// Replaces (paths relative to /root/reference/icem/):
//   controllers/icem.py:61-82   colorednoise.powerlaw_psd_gaussian(beta, (N, d, h)).transpose(0, 2, 1) * std + mean,
//                               np.clip(low, high)
//   controllers/icem.py:84-89   mean injection (row 0 of the last iteration)
//   controllers/icem.py:91-104  shifted elites: previous elites moved one step, fresh last action
//
// Why a second sampler next to the warp-per-trajectory one in rollout.cuh: that one feeds a rollout that owns
// the warp anyway (sampling is <1 % of a ground-truth rollout).  Alone, sampling is an HBM-write-bound job
// (4*h*d bytes per trajectory), so the instruction count per output float is what decides how close to the
// roofline it gets.  Here a thread keeps its series' 2K unit normals in registers and the h x 2K inverse-DFT
// synthesis is folded four ways with the symmetries of the real DFT (even h):
//     with  Ce/Co = sum over even/odd k of Gc[t][k] zr[k],  Se/So = same of Gs[t][k] zi[k],   half = h/2,
//     y[t]        = (Ce + Co) + (Se + So)        y[h - t]    = (Ce + Co) - (Se + So)
//     y[half - t] = (Ce - Co) - (Se - So)        y[half + t] = (Ce - Co) + (Se - So)         t = 0 .. half/2
// so one pass of 2K multiply-adds yields four outputs (8 instead of 32 FMAs per output at h = 30); the table
// rows Gc[t][:], Gs[t][:] (t <= half/2) come from the planner's synthesis matrix G and are read as 128-bit
// shared-memory broadcasts.  Outputs are assembled as [row][h][d] tiles in shared memory and shipped with one
// TMA bulk store per batch of rows (rows are contiguous in HBM), double-buffered against the next batch.
#pragma once
#include "rollout.cuh"

namespace icem {

constexpr int kSamplerThreads = 256;

__device__ __forceinline__ void tma_store_wait_read_le1() {
  asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
}

// Box-Muller on the SFU pipe only (lg2 / sqrt / sin / cos approximations): production sampling noise, not a
// parity path (parity mode injects the reference's draws).
__device__ __forceinline__ void box_muller_sfu(uint32_t a, uint32_t b, float& n0, float& n1) {
  const float u1 = fmaf((float)a, 2.3283064365386963e-10f, 1.1641532182693481e-10f);     // (0, 1]
  const float ang = fmaf((float)b, 1.4629180792671596e-9f, 7.314590396335798e-10f);      // 2 pi * (0, 1]
  float l, r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(u1));
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(-1.3862943611198906f * l));            // sqrt(-2 ln u1)
  n0 = r * __cosf(ang);
  n1 = r * __sinf(ang);
}

template <int KPAD>
__global__ void __launch_bounds__(kSamplerThreads, KPAD > 16 ? 2 : 3)
colored_sampler_kernel(RolloutArgs a, SamplerConst sc, int rows_per_batch) {
  extern __shared__ __align__(128) float smem[];
  const int tid = threadIdx.x;
  const int h = sc.h, d = sc.d, hd = h * d, K = sc.K, half = h >> 1, Q = half >> 1;
  const int stride = a.stride, R = rows_per_batch;
  float* s_c = smem;                                // [Q+1][KPAD] cosine-side rows of G (zero padded)
  float* s_s = s_c + (Q + 1) * KPAD;                // [Q+1][KPAD] sine-side rows
  float2* s_ms = reinterpret_cast<float2*>(s_s + (Q + 1) * KPAD);   // [hd] (mean, std) pairs
  float* s_tile = reinterpret_cast<float*>(s_ms + ((hd + 1) & ~1)); // [2][R][stride]
  const int batch_floats = R * stride;

  for (int i = tid; i < (Q + 1) * KPAD; i += kSamplerThreads) {
    const int t = i / KPAD, k = i - t * KPAD;
    s_c[i] = k < K ? sc.G[(size_t)t * 2 * K + k] : 0.f;
    s_s[i] = k < K ? sc.G[(size_t)t * 2 * K + K + k] : 0.f;
  }
  for (int i = tid; i < hd; i += kSamplerThreads) s_ms[i] = make_float2(a.mean[i], a.std[i]);
  for (int i = tid; i < 2 * batch_floats; i += kSamplerThreads) s_tile[i] = 0.f;   // also the row padding

  const StepState ss = *a.ss;
  const int n_rows = a.n_fresh_local + ((a.iteration == 0 && ss.has_prev_elites) ? a.n_shift_local : 0);
  const int cta_lo = (int)((long long)n_rows * blockIdx.x / gridDim.x);
  const int cta_hi = (int)((long long)n_rows * (blockIdx.x + 1) / gridDim.x);

  const int r = (int)__umulhi((uint32_t)tid, sc.magic_d);      // row inside the batch
  const int dim = tid - r * d;
  const bool lane_used = r < R;
  const float lo = sc.low[dim], hi = sc.high[dim];
  __syncthreads();

  int it = 0;
  for (int base = cta_lo; base < cta_hi; base += R, ++it) {
    float* tile = s_tile + (it & 1) * batch_floats;
    // the bulk store that last read this buffer (two batches ago) must be done with it
    if (tid == 0) tma_store_wait_read_le1();
    __syncthreads();
    const int row = base + r;
    if (lane_used && row < cta_hi) {
      const bool shifted = row >= a.n_fresh_local;
      const uint32_t grow = shifted ? (uint32_t)(a.n_fresh_global + (row - a.n_fresh_local))
                                    : (uint32_t)(a.global_offset + row);
      // ---- unit normals of this series: zr[k], zi[k] in registers --------------------------------
      float za[KPAD], zb[KPAD];
      if (ss.inject) {
        const float* sr = a.inj_zr + ((size_t)row * d + dim) * K;
        const float* si = a.inj_zi + ((size_t)row * d + dim) * K;
#pragma unroll
        for (int k = 0; k < KPAD; ++k) {
          za[k] = k < K ? sr[k] : 0.f;
          zb[k] = k < K ? si[k] : 0.f;
        }
      } else {
        const uint32_t c1 = (uint32_t)dim << 8;
#pragma unroll
        for (int j = 0; j < KPAD / 4; ++j) {
          if (4 * j < K) {
            Philox4 p0 = philox4x32_10(grow, c1 + j, ss.step, (uint32_t)a.iteration, a.seed_lo, a.seed_hi);
            box_muller_sfu(p0.x, p0.y, za[4 * j], za[4 * j + 1]);
            box_muller_sfu(p0.z, p0.w, za[4 * j + 2], za[4 * j + 3]);
            Philox4 p1 = philox4x32_10(grow, c1 + 128 + j, ss.step, (uint32_t)a.iteration, a.seed_lo, a.seed_hi);
            box_muller_sfu(p1.x, p1.y, zb[4 * j], zb[4 * j + 1]);
            box_muller_sfu(p1.z, p1.w, zb[4 * j + 2], zb[4 * j + 3]);
          } else {
            za[4 * j] = za[4 * j + 1] = za[4 * j + 2] = za[4 * j + 3] = 0.f;
            zb[4 * j] = zb[4 * j + 1] = zb[4 * j + 2] = zb[4 * j + 3] = 0.f;
          }
        }
      }
      // ---- folded synthesis + affine + clip -> tile ---------------------------------------------
      const float2* pm = s_ms + dim;
      float* po = tile + r * stride + dim;
      auto emit = [&](int o, float y) {          // o = t * d
        const float2 ms = pm[o];
        po[o] = fminf(fmaxf(fmaf(y, ms.y, ms.x), lo), hi);
      };
      int o0 = 0, o1 = h * d, o2 = half * d, o3 = half * d;     // t*d, (h-t)*d, (half-t)*d, (half+t)*d
      for (int t = 0; t <= Q; ++t, o0 += d, o1 -= d, o2 -= d, o3 += d) {
        // keep the four running offsets in registers (ptxas otherwise re-derives each from t with IMADs)
        asm volatile("" : "+r"(o0), "+r"(o1), "+r"(o2), "+r"(o3));
        const float4* c4 = reinterpret_cast<const float4*>(s_c + t * KPAD);
        const float4* s4 = reinterpret_cast<const float4*>(s_s + t * KPAD);
        float ce = 0.f, co = 0.f, se = 0.f, so = 0.f;
#pragma unroll
        for (int q = 0; q < KPAD / 4; ++q) {
          const float4 c = c4[q], s = s4[q];
          ce = fmaf(c.x, za[4 * q], ce);     co = fmaf(c.y, za[4 * q + 1], co);
          ce = fmaf(c.z, za[4 * q + 2], ce); co = fmaf(c.w, za[4 * q + 3], co);
          se = fmaf(s.x, zb[4 * q], se);     so = fmaf(s.y, zb[4 * q + 1], so);
          se = fmaf(s.z, zb[4 * q + 2], se); so = fmaf(s.w, zb[4 * q + 3], so);
        }
        const float cp = ce + co, cm = ce - co, sp = se + so, sm = se - so;
        emit(o0, cp + sp);
        if (t > 0) emit(o1, cp - sp);
        if (o2 > o0) {
          emit(o2, cm - sm);
          if (t > 0) emit(o3, cm + sm);
        }
      }
      // mean row (icem.py:87-88) and shifted elites (icem.py:91-104) are <= 1 + n_keep rows of a population:
      // their threads overwrite what they just sampled (a shifted row keeps its fresh LAST action)
      const bool mean_row = a.inject_mean_row0 && grow == 0u && !shifted;
      if (mean_row | shifted) {
        const float* elite = a.prev_elites + (size_t)(shifted ? row - a.n_fresh_local : 0) * stride + dim + d;
        for (int t = 0; t < h; ++t) {
          const int o = t * d;
          if (mean_row) po[o] = pm[o].x;
          else if (t < h - 1) po[o] = elite[o];
        }
      }
    }
    // ---- ship the batch: rows base .. base+nrows-1 are contiguous in HBM ---------------------------
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      const int nrows = min(R, cta_hi - base);
      tma_store_1d(a.actions + (size_t)base * stride, tile, (uint32_t)(nrows * stride) * 4u);
      tma_store_commit();
    }
  }
  if (tid == 0) tma_store_wait_all();
}

// rows per batch and dynamic shared memory of colored_sampler_kernel
inline int sampler_rows_per_batch(int d, int stride) {
  int R = kSamplerThreads / d;
  const int cap = (64 * 1024) / (2 * stride * 4);     // <= 64 KB of tiles per CTA: 3 CTAs per SM
  if (R > cap) R = cap;
  return R < 1 ? 1 : R;
}
inline size_t sampler_smem_bytes(int h, int d, int kpad, int stride, int R) {
  const int Q = (h / 2) / 2, hd = h * d;
  return (size_t)(2 * (Q + 1) * kpad + 2 * ((hd + 1) & ~1) + 2 * R * stride) * sizeof(float);
}

}  // namespace icem
