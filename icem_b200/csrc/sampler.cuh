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

// One standard-normal PAIR from two random words (Box-Muller on the SFU pipe: lg2 / sqrt / sin / cos approximations;
// production sampling noise, not a parity path -- parity mode injects the reference's draws).  `rad` carries >= 21
// random bits at the top and one forced low bit (never zero, exactly representable in fp32); `ang` >= 21 random top
// bits.  -2 ln(rad / 2^32) = 2 ln2 (32 - lg2 rad); the additive constant is 2e-5 above 64 ln2 so that the approximate
// lg2 can never push the radicand below zero (shifts r^2 by 2e-5: below fp32 resolution of the sampled actions).
__device__ __forceinline__ void normal_pair(uint32_t rad, uint32_t ang, float& n0, float& n1) {
  float l, r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"((float)rad));
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fmaf(l, -1.3862943611198906f, 44.361440f)));
  const float a = (float)ang * 1.4629180792671596e-9f;           // 2 pi / 2^32
  n0 = r * __cosf(a);
  n1 = r * __sinf(a);
}

// Three normal pairs from ONE Philox call: 128 bits = 6 x 21 (+2 spare).  Pairs 0 / 1 take the top 21 bits of
// (x, y) / (z, w); pair 2 the low 11 bits of x and y (radius) and of z and w (angle) -- disjoint bit fields.
__device__ __forceinline__ void normal_pairs3(const Philox4& p, float (&n)[6]) {
  normal_pair((p.x & 0xFFFFF800u) | 0x400u, p.y & 0xFFFFF800u, n[0], n[1]);
  normal_pair((p.z & 0xFFFFF800u) | 0x400u, p.w & 0xFFFFF800u, n[2], n[3]);
  normal_pair(__funnelshift_l(p.y << 21, p.x, 21) | 0x200u, __funnelshift_l(p.w << 21, p.z, 21), n[4], n[5]);
}

// Table rows of the compile-time-shaped kernels travel as a KERNEL PARAMETER: parameters live in constant bank 0, so
// every table entry is an immediate constant operand of its FFMA -- no load instruction at all.  (Read from shared
// memory as 128-bit broadcasts, the 64 loads per series saturated the shared-memory return path -- a broadcast still
// returns 512 B per warp -- and capped the kernel at 39 % of HBM: ncu mio_throttle 1.6 per issue.)
constexpr int sampler_table_floats(int kpad, int h, int d) { return (h > 0 && d > 0) ? 2 * ((h / 2) / 2 + 1) * kpad : 1; }
template <int N>
struct SamplerTable {          // entry i = the float pair (2i, 2i+1) of the row-major table
  unsigned long long v2[(N + 1) / 2];
};

__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

// KPAD: register slots for the K = h/2 + 1 frequency bins.  H, D > 0: horizon and action dim fixed at compile time
// (the headline shapes: every tile / table offset becomes an immediate and the fold loop is unrolled); 0 = run time.
// CTA size: 256 threads, or 192 for the compile-time-shaped 16-bin kernels -- three resident CTAs (18 warps, at most 5
// per SM sub-partition) then leave each thread 96
// registers instead of 80, which is what keeps ptxas from spilling the unrolled fold.
constexpr int sampler_threads(int kpad, int h, int d) { return (kpad == 16 && h > 0 && d > 0) ? 192 : kSamplerThreads; }

template <int KPAD, int H, int D>
__global__ void __launch_bounds__(sampler_threads(KPAD, H, D), KPAD > 16 ? 2 : 3)
colored_sampler_kernel(RolloutArgs a, SamplerConst sc, int rows_per_batch,
                       const __grid_constant__ SamplerTable<sampler_table_floats(KPAD, H, D)> tab) {
  extern __shared__ __align__(128) float smem[];
  constexpr bool kStatic = H > 0 && D > 0;
  constexpr int kThreads = sampler_threads(KPAD, H, D);
  const int tid = threadIdx.x;
  const int h = kStatic ? H : sc.h, d = kStatic ? D : sc.d, hd = h * d, K = kStatic ? H / 2 + 1 : sc.K;
  const int half = h >> 1, Q = half >> 1;
  const int stride = kStatic ? ((H * D + 3) & ~3) : a.stride, R = rows_per_batch;
  float* s_c = smem;                                // [Q+1][KPAD] cosine-side rows of G (zero padded)
  float* s_s = s_c + (Q + 1) * KPAD;                // [Q+1][KPAD] sine-side rows
  float2* s_ms = reinterpret_cast<float2*>(s_s + (Q + 1) * KPAD);   // [hd] (mean, std) pairs
  float* s_tile = reinterpret_cast<float*>(s_ms + ((hd + 1) & ~1)); // [2][R][stride]
  const int batch_floats = R * stride;

  if constexpr (!kStatic) {       // the compile-time-shaped kernels take their table rows from the kernel parameters
    for (int i = tid; i < (Q + 1) * KPAD; i += kThreads) {
      const int t = i / KPAD, k = i - t * KPAD;
      s_c[i] = k < K ? sc.G[(size_t)t * 2 * K + k] : 0.f;
      s_s[i] = k < K ? sc.G[(size_t)t * 2 * K + K + k] : 0.f;
    }
  }
  for (int i = tid; i < hd; i += kThreads) s_ms[i] = make_float2(a.mean[i], a.std[i]);
  for (int i = tid; i < 2 * batch_floats; i += kThreads) s_tile[i] = 0.f;   // also the row padding

  const StepState ss = *a.ss;
  const int n_rows = a.n_fresh_local + ((a.iteration == 0 && ss.has_prev_elites) ? a.n_shift_local : 0);
  const int cta_lo = (int)((long long)n_rows * blockIdx.x / gridDim.x);
  const int cta_hi = (int)((long long)n_rows * (blockIdx.x + 1) / gridDim.x);

  const int r = kStatic ? tid / D : (int)__umulhi((uint32_t)tid, sc.magic_d);      // row inside the batch
  const int dim = tid - r * d;
  const bool lane_used = r < R;
  const float lo = sc.low[dim], hi = sc.high[dim];
  __syncthreads();

  int it = 0;
  for (int base = cta_lo; base < cta_hi; base += R, ++it) {
    float* tile = s_tile + (it & 1) * batch_floats;
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
        // An even-h series uses h normals: (zr, zi) of the bins 1 .. K-2 and the real parts of DC and Nyquist
        // (their imaginary parts do not enter a real signal).  Pair q < K-2 is bin q+1's (real, imaginary) part --
        // Box-Muller's (r cos, r sin) IS that bin's Rayleigh amplitude and uniform phase -- and pair K-2 is
        // (DC, Nyquist).  Philox4x32-7 keyed by (seed; trajectory, dim * 256 + call, plan step, iteration), three
        // pairs per call: 5 calls for h = 30, 2 for h = 12.
        const uint32_t c1 = (uint32_t)dim << 8;
        float nyq = 0.f;
#pragma unroll
        for (int k = 0; k < KPAD; ++k) za[k] = zb[k] = 0.f;
#pragma unroll
        for (int c = 0; c < (KPAD + 1) / 3; ++c) {
          if (3 * c < K - 1) {
            float n[6];
            normal_pairs3(philox4x32<7>(grow, c1 + c, ss.step, (uint32_t)a.iteration, a.seed_lo, a.seed_hi), n);
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              const int q = 3 * c + j;
              if (q + 1 < KPAD && q < K - 2) { za[q + 1] = n[2 * j]; zb[q + 1] = n[2 * j + 1]; }
              if (q == K - 2) { za[0] = n[2 * j]; nyq = n[2 * j + 1]; }
            }
          }
        }
#pragma unroll
        for (int k = 1; k < KPAD; ++k) za[k] = (k == K - 1) ? nyq : za[k];     // selects: static register indices
      }
      // ---- folded synthesis + affine + clip -> tile ---------------------------------------------
      const float2* pm = s_ms + dim;
      float* po = tile + r * stride + dim;
      auto emit = [&](int o, float y) {          // o = t * d
        const float2 ms = pm[o];
        po[o] = fminf(fmaxf(fmaf(y, ms.y, ms.x), lo), hi);
      };
      auto combine = [&](int t, float ce, float co, float se, float so) {   // outputs t, h - t, half - t, half + t
        const float cp = ce + co, cm = ce - co, sp = se + so, sm = se - so;
        emit(t * d, cp + sp);
        if (t > 0) emit((h - t) * d, cp - sp);
        if (half - t > t) {
          emit((half - t) * d, cm - sm);
          if (t > 0) emit((half + t) * d, cm + sm);
        }
      };
      if constexpr (kStatic) {
        // packed fp32 (fma.rn.f32x2, SASS FFMA2): (ce, co) += (G[t][k], G[t][k+1]) * (zr[k], zr[k+1]) is ONE issue slot,
        // the table pair a 64-bit uniform-register operand
        constexpr int rows = (H / 2) / 2 + 1;
        unsigned long long zap[KPAD / 2], zbp[KPAD / 2];
#pragma unroll
        for (int k = 0; k < KPAD; k += 2) { zap[k / 2] = pack2(za[k], za[k + 1]); zbp[k / 2] = pack2(zb[k], zb[k + 1]); }
#pragma unroll
        for (int t = 0; t < rows; ++t) {
          unsigned long long ac = 0ull, as = 0ull;       // (+0.f, +0.f)
#pragma unroll
          for (int k = 0; k < KPAD; k += 2) {
            if (k < H / 2 + 1) {
              ac = fma2(zap[k / 2], tab.v2[(t * KPAD + k) / 2], ac);
              as = fma2(zbp[k / 2], tab.v2[((rows + t) * KPAD + k) / 2], as);
            }
          }
          float ce, co, se, so;
          unpack2(ac, ce, co);
          unpack2(as, se, so);
          combine(t, ce, co, se, so);
        }
      } else {
        for (int t = 0; t <= Q; ++t) {
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
          combine(t, ce, co, se, so);
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
    // ONE barrier per batch: before it, thread 0 makes sure the previous batch's bulk store has finished READING its
    // buffer (it had this whole batch to do so) -- the buffer the next batch writes into.
    fence_proxy_async_smem();
    if (tid == 0) tma_store_wait_read();
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
inline int sampler_rows_per_batch(int d, int stride, int threads = kSamplerThreads) {
  int R = threads / d;
  const int cap = (64 * 1024) / (2 * stride * 4);     // <= 64 KB of tiles per CTA: 3 CTAs per SM
  if (R > cap) R = cap;
  return R < 1 ? 1 : R;
}
inline size_t sampler_smem_bytes(int h, int d, int kpad, int stride, int R) {
  const int Q = (h / 2) / 2, hd = h * d;
  return (size_t)(2 * (Q + 1) * kpad + 2 * ((hd + 1) & ~1) + 2 * R * stride) * sizeof(float);
}

}  // namespace icem
