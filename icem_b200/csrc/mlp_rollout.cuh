// Tensor-core rollout of a dense MLP forward model (BASELINE configs[3]): tcgen05.mma + TMEM.
//
// The batched-model path of the reference: `ForwardModelWithDefaults.predict_n_steps`
// (icem/models/abstract_models.py:17-53) calling a dense `predict` h times on [p, obs+act] -- the reference ships
// no such model (icem/models/__init__.py:5-8), so the architecture is this repo's:
//     obs' = obs + W3 tanh(W2 tanh(W1 [obs, act] + b1) + b2) + b3            (2 hidden layers of width H)
// followed by the per-step cost on the PRE-action observation (controllers/abstract_controller.py:74-91).
//
// One CTA = one tile of 128 trajectories; 512 threads: thread (q, lane, g) with warp = 4 g + q owns trajectory row
// r = 32 q + lane (TMEM lane r; a warp may only touch the lane quarter warp % 4) and every 4th 32-column chunk
// of the hidden layer in the epilogues; the fp32 observation is replicated in the 4 threads of a row.  The three weight
// matrices stay RESIDENT in shared memory as bf16 in the tcgen05 K-major no-swizzle core-matrix layout
// ([n/8][k/8][n%8][k%8], 128-byte core matrices) for the whole kernel; per control step:
//     X[128 x 32] (bf16, smem)  --tcgen05.mma M128 N=H K=16 x2-->  D[128 x H] fp32 in TMEM
//     epilogue: tcgen05.ld 32x32b.x32 -> +bias, tanh.approx -> bf16 -> smem (the next MMA's A operand)
//     H1[128 x H] --16 MMAs--> D ; epilogue ; H2 --16 MMAs (N=32)--> D[128 x 32] ; obs += D + b3 (fp32 registers)
// MMAs are issued by ONE thread and tracked with tcgen05.commit on an mbarrier; accumulators never leave TMEM
// except through the epilogue loads.  Actions come from the sampler kernel's HBM/L2 tiles.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"
#include "rollout.cuh"

namespace icem {

constexpr int kMlpTile = 128;     // trajectories per CTA tile (UMMA M)
constexpr int kMlpColGroups = 4;  // threads per trajectory row: each takes every 4th 32-column chunk in the epilogues
constexpr int kMlpThreads = kMlpTile * kMlpColGroups;   // 16 warps: 4 per scheduler hide the TMEM / MUFU latency
constexpr int kMlpInPad = 32;     // padded input width (obs + act <= 32)
constexpr int kMlpOutPad = 32;    // padded output width (obs <= 32)

struct MlpParams {
  int obs_dim, act_dim, hidden;               // hidden in {64, 128, 256}
  const __nv_bfloat16* w1;                    // packed [hidden x kMlpInPad]
  const __nv_bfloat16* w2;                    // packed [hidden x hidden]
  const __nv_bfloat16* w3;                    // packed [kMlpOutPad x hidden]
  const float* bias;                          // [hidden + hidden + kMlpOutPad]
};

// ---- tcgen05 wrappers (PTX ISA 8.6+, sm_100a) -------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
               :: "r"(smem_addr(smem_dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
               :: "r"(smem_addr(bar)) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, bf16 x bf16 -> fp32
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// 32 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// K-major, no swizzle: core matrix (8 rows x 16 B) = 128 contiguous bytes; LBO = distance between the two K chunks
// of one MMA (128 B), SBO = distance between 8-row groups.  Descriptor version 1 (Blackwell).
__device__ __forceinline__ uint64_t umma_desc(const void* smem, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  const uint32_t addr = smem_addr(smem);
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;                  // version = 1
  return d;                                // base_offset 0, lbo_mode 0, layout_type 0 (SWIZZLE_NONE)
}
// instruction descriptor: D fp32, A/B bf16, both K-major, dense
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// packed core-matrix index of element (row, k) of a [rows x K] K-major operand (bf16 elements)
__host__ __device__ inline size_t umma_pack_index(int row, int k, int K) {
  return (size_t)(row >> 3) * (size_t)(K >> 3) * 64 + (size_t)(k >> 3) * 64 + (size_t)(row & 7) * 8 + (size_t)(k & 7);
}

inline size_t mlp_smem_bytes(int hidden) {
  const size_t w = ((size_t)hidden * kMlpInPad + (size_t)hidden * hidden + (size_t)kMlpOutPad * hidden) * 2;
  const size_t a = (size_t)kMlpTile * hidden * 2;
  const size_t b = ((size_t)2 * hidden + kMlpOutPad) * 4;
  return w + a + b + 64;
}

// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kMlpThreads, 1)
mlp_rollout_kernel(RolloutArgs a, SamplerConst sc, CostConst cc, MlpParams mp) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const int H = mp.hidden;
  __nv_bfloat16* sW1 = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  __nv_bfloat16* sW2 = sW1 + (size_t)H * kMlpInPad;
  __nv_bfloat16* sW3 = sW2 + (size_t)H * H;
  __nv_bfloat16* sA = sW3 + (size_t)kMlpOutPad * H;                   // [128 x max(H, 32)] activations
  float* sBias = reinterpret_cast<float*>(sA + (size_t)kMlpTile * H);
  uint64_t* bar = reinterpret_cast<uint64_t*>(sBias + 2 * H + kMlpOutPad);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int q = warp & 3, g = warp >> 2;             // TMEM lane quarter, column group
  const int r = q * 32 + (tid & 31);                 // trajectory row inside the tile
  const int h = sc.h, d = sc.d, od = mp.obs_dim;

  // ---- one-time: resident weights, barrier, TMEM ----
  {
    const uint4* src1 = reinterpret_cast<const uint4*>(mp.w1);
    uint4* dst1 = reinterpret_cast<uint4*>(sW1);
    for (int i = tid; i < H * kMlpInPad / 8; i += blockDim.x) dst1[i] = src1[i];
    const uint4* src2 = reinterpret_cast<const uint4*>(mp.w2);
    uint4* dst2 = reinterpret_cast<uint4*>(sW2);
    for (int i = tid; i < H * H / 8; i += blockDim.x) dst2[i] = src2[i];
    const uint4* src3 = reinterpret_cast<const uint4*>(mp.w3);
    uint4* dst3 = reinterpret_cast<uint4*>(sW3);
    for (int i = tid; i < kMlpOutPad * H / 8; i += blockDim.x) dst3[i] = src3[i];
    for (int i = tid; i < 2 * H + kMlpOutPad; i += blockDim.x) sBias[i] = mp.bias[i];
  }
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
  }
  const uint32_t tmem_cols = H <= 64 ? 64u : (H <= 128 ? 128u : 256u);
  if (warp == 0) tmem_alloc(tmem_slot, tmem_cols);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);      // this warp's 32 lanes

  const uint32_t idesc_h = umma_idesc_bf16(kMlpTile, H);
  const uint32_t idesc_o = umma_idesc_bf16(kMlpTile, kMlpOutPad);
  const uint32_t sbo_in = (kMlpInPad / 8) * 128, sbo_h = (uint32_t)(H / 8) * 128;
  uint32_t phase = 0;

  const StepState ss = *a.ss;
  const int n_rows = a.n_fresh_local + ((a.iteration == 0 && ss.has_prev_elites) ? a.n_shift_local : 0);
  unsigned char* rowA = reinterpret_cast<unsigned char*>(sA) + (size_t)(r & 7) * 16;     // + (r/8)*SBO + chunk*128

  for (int tile0 = blockIdx.x * kMlpTile; tile0 < n_rows; tile0 += gridDim.x * kMlpTile) {
    const int row = tile0 + r;
    const bool valid = row < n_rows;
    float x[kMlpInPad];                      // [obs(od), act(d), 0...]: the fp32 state of this trajectory
#pragma unroll
    for (int i = 0; i < kMlpInPad; ++i) x[i] = (valid && i < od) ? a.start_state[i] : 0.f;
    const float* acts = a.actions + (size_t)(valid ? row : 0) * a.stride;
    float total = (cc.reduce == 1) ? INFINITY : 0.f;

    for (int t = 0; t < h; ++t) {
      // ---- action of this step, per-step cost on the pre-action observation ----
      float a2 = 0.f, oa = 0.f, ob = 0.f;
#pragma unroll
      for (int i = 0; i < kMlpInPad; ++i) {
        if (i >= od && i < od + d) {
          x[i] = valid ? acts[t * d + (i - od)] : 0.f;
          a2 = fmaf(x[i], x[i], a2);
        }
        oa = (i == cc.idx_a) ? x[i] : oa;
        ob = (i == cc.idx_b) ? x[i] : ob;
      }
      float c;
      if (cc.kind == 0) {
        c = 0.1f * a2 - ob;
        if (cc.penalise_flipping) c += (oa > 1.5707963267948966f ? 10.f : 0.f) + (oa < -1.5707963267948966f ? 10.f : 0.f);
      } else {
        c = -oa + 0.1f * a2;
      }
      if (cc.reduce == 0) total += c;
      else if (cc.reduce == 1) total = fminf(total, c);
      else total = c;
      if (t + 1 == h) break;                 // the final predicted state is never scored

      // ---- X row -> smem (bf16, K-major core matrices), layer 1: thread g writes 16-byte chunk g ----
      {
        unsigned char* dst = rowA + (size_t)(r >> 3) * sbo_in;
#pragma unroll
        for (int ch = 0; ch < kMlpInPad / 8; ++ch) {
          if (ch == g) {
            __nv_bfloat162 p0 = __floats2bfloat162_rn(x[8 * ch], x[8 * ch + 1]);
            __nv_bfloat162 p1 = __floats2bfloat162_rn(x[8 * ch + 2], x[8 * ch + 3]);
            __nv_bfloat162 p2 = __floats2bfloat162_rn(x[8 * ch + 4], x[8 * ch + 5]);
            __nv_bfloat162 p3 = __floats2bfloat162_rn(x[8 * ch + 6], x[8 * ch + 7]);
            uint4 v;
            v.x = *reinterpret_cast<uint32_t*>(&p0); v.y = *reinterpret_cast<uint32_t*>(&p1);
            v.z = *reinterpret_cast<uint32_t*>(&p2); v.w = *reinterpret_cast<uint32_t*>(&p3);
            *reinterpret_cast<uint4*>(dst + ch * 128) = v;
          }
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < kMlpInPad / 16; ++ks)
          tc_mma_bf16(tmem_base, umma_desc(reinterpret_cast<unsigned char*>(sA) + ks * 256, 128, sbo_in),
                      umma_desc(reinterpret_cast<unsigned char*>(sW1) + ks * 256, 128, sbo_in), idesc_h, ks > 0);
        tc_commit(bar);
      }
      // ---- hidden layers: epilogue (bias + tanh -> bf16 A operand), next MMA ----
      for (int layer = 0; layer < 2; ++layer) {
        mbar_wait(bar, phase);
        phase ^= 1;
        tc_fence_after();
        const float* bs = sBias + layer * H;
        unsigned char* dst = rowA + (size_t)(r >> 3) * sbo_h;
        for (int c0 = g * 32; c0 < H; c0 += 32 * kMlpColGroups) {      // 32-column chunks dealt round-robin
          float v[32];
          tmem_ld32(taddr + (uint32_t)c0, v);
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            uint32_t pk[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int k = 8 * ch + 2 * q;
              __nv_bfloat162 pr = __floats2bfloat162_rn(tanh_approx(v[k] + bs[c0 + k]),
                                                        tanh_approx(v[k + 1] + bs[c0 + k + 1]));
              pk[q] = *reinterpret_cast<uint32_t*>(&pr);
            }
            *reinterpret_cast<uint4*>(dst + (size_t)(c0 / 8 + ch) * 128) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
          tc_fence_after();
          const unsigned char* wB = reinterpret_cast<const unsigned char*>(layer == 0 ? sW2 : sW3);
          const uint32_t id = layer == 0 ? idesc_h : idesc_o;
          for (int ks = 0; ks < H / 16; ++ks)
            tc_mma_bf16(tmem_base, umma_desc(reinterpret_cast<unsigned char*>(sA) + ks * 256, 128, sbo_h),
                        umma_desc(wB + ks * 256, 128, sbo_h), id, ks > 0);
          tc_commit(bar);
        }
      }
      // ---- output layer epilogue: obs += W3 h2 + b3 (fp32 state in registers) ----
      mbar_wait(bar, phase);
      phase ^= 1;
      tc_fence_after();
      {
        float v[32];
        tmem_ld32(taddr, v);
        const float* b3 = sBias + 2 * H;
#pragma unroll
        for (int i = 0; i < kMlpOutPad; ++i)
          if (i < od) x[i] += v[i] + b3[i];
      }
      tc_fence_before();      // the next step's MMA overwrites these TMEM columns after the CTA barrier above it
    }
    if (valid && g == 0) a.costs[row] = total;
    __syncthreads();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, tmem_cols);
}

// one transition of the same model for a single state (env.step on the device model / closed-loop bench):
// same bf16 roundings as the tensor-core path, CUDA-core arithmetic
__global__ void mlp_advance_kernel(MlpParams mp, const float* state, const float* action, float* next_state,
                                   float* obs_out, int obs_dim_out) {
  __shared__ float xin[kMlpInPad];
  __shared__ float h1[256], h2[256];
  const int H = mp.hidden, od = mp.obs_dim, d = mp.act_dim, tid = threadIdx.x;
  auto bf = [](float v) { return __bfloat162float(__float2bfloat16_rn(v)); };
  if (tid < kMlpInPad) xin[tid] = tid < od ? state[tid] : (action && tid < od + d ? action[tid - od] : 0.f);
  __syncthreads();
  if (action) {
    for (int n = tid; n < H; n += blockDim.x) {
      float acc = 0.f;
      for (int k = 0; k < kMlpInPad; ++k) acc = fmaf(bf(xin[k]), __bfloat162float(mp.w1[umma_pack_index(n, k, kMlpInPad)]), acc);
      h1[n] = bf(tanh_approx(acc + mp.bias[n]));
    }
    __syncthreads();
    for (int n = tid; n < H; n += blockDim.x) {
      float acc = 0.f;
      for (int k = 0; k < H; ++k) acc = fmaf(h1[k], __bfloat162float(mp.w2[umma_pack_index(n, k, H)]), acc);
      h2[n] = bf(tanh_approx(acc + mp.bias[H + n]));
    }
    __syncthreads();
    if (tid < od) {
      float acc = 0.f;
      for (int k = 0; k < H; ++k) acc = fmaf(h2[k], __bfloat162float(mp.w3[umma_pack_index(tid, k, H)]), acc);
      xin[tid] = xin[tid] + acc + mp.bias[2 * H + tid];
    }
    __syncthreads();
  }
  if (next_state && tid < od) next_state[tid] = xin[tid];
  if (obs_out && tid < obs_dim_out) obs_out[tid] = xin[tid];
}

}  // namespace icem
