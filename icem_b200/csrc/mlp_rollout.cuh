// Tensor-core rollout of a dense MLP forward model (BASELINE configs[3]): tcgen05.mma + TMEM.
//
// The batched-model path of the reference: `ForwardModelWithDefaults.predict_n_steps`
// (icem/models/abstract_models.py:17-53) calling a dense `predict` h times on [p, obs+act] -- the reference ships
// no such model (icem/models/__init__.py:5-8), so the architecture is this repo's:
//     obs' = obs + W3 tanh(W2 tanh(W1 [obs, act] + b1) + b2) + b3            (2 hidden layers of width H)
// followed by the per-step cost on the PRE-action observation (controllers/abstract_controller.py:74-91).
//
// One CTA = one tile of 128 trajectories.  16 epilogue warps + 1 MMA-issuer warp (warp specialisation, no CTA-wide
// barrier inside the rollout): epilogue thread (q, lane, g), warp = 4 g + q, owns trajectory row r = 32 q + lane
// (TMEM lane r; a warp may only touch the lane quarter warp % 4) and the 16 columns g*16.. of every 64-column
// STAGE of a hidden layer; the fp32 observation is replicated in the 4 threads of a row.  The three weight matrices
// stay RESIDENT in shared memory as fp16 in the tcgen05 K-major no-swizzle core-matrix layout
// ([n/8][k/8][n%8][k%8], 128-byte core matrices) for the whole kernel.  Per control step:
//     X[128 x 32] (fp16, smem) --2 MMAs M128 N=H K=16--> D1[128 x H] fp32 in TMEM (columns 0..H)
//     epilogue 1, stage by stage: tcgen05.ld 32x32b.x16 -> +bias, tanh.approx -> fp16 -> smem; as soon as a stage
//       (64 columns = 4 K-steps of the next layer's A operand) is in shared memory its warps arrive on that stage's
//       mbarrier and the issuer accumulates those K-steps into D2 (TMEM columns 256..256+H): the layer-2 MMAs run
//       UNDER the layer-1 epilogue instead of after it
//     epilogue 2 (from D2), same stage pipeline into the N=32 output MMAs -> D3 (TMEM columns 0..32)
//     obs += D3 + b3 (fp32 registers)
// MMAs are issued by ONE thread and tracked with tcgen05.commit on mbarriers; accumulators never leave TMEM except
// through the epilogue loads.  Actions come from the sampler kernel's HBM/L2 tiles, prefetched one step ahead.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"
#include "rollout.cuh"

namespace icem {

// Tensor-core operand type: IEEE half (11 significand bits), fp32 accumulation in TMEM.  tcgen05 `kind::f16` runs fp16
// and bf16 operands at the same rate; against the fp32 model (oracle MlpModelF32 == torch fp32 nn.Sequential) bf16's 8
// bits cost the planner half of its elite set at N = 65536 (profiles/r2_mlp_bf16_vs_fp32.json), fp16 keeps it.  The
// observations / actions / activations of these models are O(1..100): far inside the fp16 range.
typedef __half mlp_op_t;
__host__ __device__ inline mlp_op_t mlp_op_from_float(float v) { return __float2half_rn(v); }

constexpr int kMlpTile = 128;     // trajectories per CTA tile (UMMA M)
constexpr int kMlpColGroups = 4;  // threads per trajectory row: each takes 16 columns of every 64-column stage
constexpr int kMlpEpiThreads = kMlpTile * kMlpColGroups;   // 16 epilogue warps: 4 per scheduler hide TMEM / MUFU latency
constexpr int kMlpThreads = kMlpEpiThreads + 32;           // + the MMA-issuer warp
constexpr int kMlpStageCols = 64; // hidden columns per pipeline stage (= 4 K-steps of the next layer's MMA)
constexpr int kMlpMaxStages = 4;  // hidden <= 256
constexpr uint32_t kMlpTmemCols = 512;   // D1 / D3 at column 0, D2 at column 256
constexpr int kMlpInPad = 32;     // padded input width (obs + act <= 32)
constexpr int kMlpOutPad = 32;    // padded output width (obs <= 32)

struct MlpParams {
  int obs_dim, act_dim, hidden;               // hidden in {64, 128, 256}
  int act_off;                                // input column of action 0: obs_dim rounded up to 8 (act_off + act_dim <= 32)
  const mlp_op_t* w1;                    // packed [hidden x kMlpInPad], columns [obs | 0 | act | 0]
  const mlp_op_t* w2;                    // packed [hidden x hidden]
  const mlp_op_t* w3;                    // packed [kMlpOutPad x hidden]
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
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, fp16 x fp16 -> fp32
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
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
// 16 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// 8 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
// split form: issue now, tcgen05.wait::ld later (the registers are undefined until the wait)
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_addr(bar)) : "memory");
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
// instruction descriptor: D fp32 (bits 4-5 = 1), A / B format fp16 (bits 7-9 / 10-12 = 0; 1 would be bf16), both
// K-major, dense
__host__ __device__ constexpr uint32_t umma_idesc_f16(int m, int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// packed core-matrix index of element (row, k) of a [rows x K] K-major operand (2-byte elements)
__host__ __device__ inline size_t umma_pack_index(int row, int k, int K) {
  return (size_t)(row >> 3) * (size_t)(K >> 3) * 64 + (size_t)(k >> 3) * 64 + (size_t)(row & 7) * 8 + (size_t)(k & 7);
}

inline size_t mlp_smem_bytes(int hidden) {
  const size_t w = ((size_t)hidden * kMlpInPad + (size_t)hidden * hidden + (size_t)kMlpOutPad * hidden) * 2;
  const size_t a = (size_t)kMlpTile * hidden * 2;
  const size_t b = ((size_t)2 * hidden + kMlpOutPad) * 4;
  return w + a + b + 128;     // + mbarriers (1 + 4 + 3) and the TMEM slot
}

#ifdef ICEM_MLP_TRACE
// debug build only (python -m icem_b200.build with defines=["ICEM_MLP_TRACE"]): cycle stamps of CTA 0's first steps
__device__ long long g_mlp_trace[64 * 32];
#define MLP_TRACE(step, slot) do { if (blockIdx.x == 0 && (step) < 64u && lane == 0 && q == 0 && (g == 0 || g == kMlpColGroups)) \
    g_mlp_trace[(step) * 32 + (slot)] = clock64(); } while (0)
#else
#define MLP_TRACE(step, slot) do { } while (0)
#endif

// ---------------------------------------------------------------------------------------------------------------
// kGoalCost: the goal-distance cost family (ICEM_COST_GOAL_DISTANCE) as its own instantiation -- it gathers up to six
// observation entries at run-time indices from the register-resident state, which must not touch the register
// allocation of the default kernel.
template <bool kGoalCost>
__global__ void __launch_bounds__(kMlpThreads, 1)
mlp_rollout_kernel(RolloutArgs a, SamplerConst sc, CostConst cc, MlpParams mp) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const int H = mp.hidden;
  mlp_op_t* sW1 = reinterpret_cast<mlp_op_t*>(smem_raw);
  mlp_op_t* sW2 = sW1 + (size_t)H * kMlpInPad;
  mlp_op_t* sW3 = sW2 + (size_t)H * H;
  mlp_op_t* sA = sW3 + (size_t)kMlpOutPad * H;                   // [128 x max(H, 32)] activations
  float* sBias = reinterpret_cast<float*>(sA + (size_t)kMlpTile * H);
  uint64_t* bar_x = reinterpret_cast<uint64_t*>(sBias + 2 * H + kMlpOutPad);   // X operand in smem      (16 arrivals)
  uint64_t* bar_a = bar_x + 1;                // [4] activation stage s in smem: layer 1 / layer 2 alternate (16 arrivals)
  uint64_t* bar_d = bar_a + kMlpMaxStages;    // [3] accumulators D1, D2, D3 complete (tcgen05.commit)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_d + 3);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, g = warp >> 2;             // TMEM lane quarter, column group (g == 4: the issuer warp)
  const int r = q * 32 + lane;                       // trajectory row inside the tile
  const int h = sc.h, d = sc.d, od = mp.obs_dim, act_off = mp.act_off;
  const int S = H / kMlpStageCols;                   // pipeline stages per hidden layer (H in {64, 128, 256})

  // ---- one-time: resident weights, barriers, TMEM ----
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
    mbar_init(bar_x, kMlpEpiThreads / 32);
    for (int s = 0; s < kMlpMaxStages; ++s) mbar_init(&bar_a[s], kMlpEpiThreads / 32);
    for (int i = 0; i < 3; ++i) mbar_init(&bar_d[i], 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, kMlpTmemCols);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_d2 = tmem_base + 256u;

  const uint32_t idesc_h = umma_idesc_f16(kMlpTile, H);
  const uint32_t idesc_o = umma_idesc_f16(kMlpTile, kMlpOutPad);
  const uint32_t sbo_in = (kMlpInPad / 8) * 128, sbo_h = (uint32_t)(H / 8) * 128;

  const StepState ss = *a.ss;
  const int n_rows = a.n_fresh_local + ((a.iteration == 0 && ss.has_prev_elites) ? a.n_shift_local : 0);
  uint32_t n = 0;      // control steps done by this CTA: bar_x / bar_d[*] complete once per step, bar_a[*] twice

  if (g == kMlpColGroups) {
    // ================= MMA issuer: one thread =================
    if (lane == 0) {
      const unsigned char* pA = reinterpret_cast<const unsigned char*>(sA);
      const unsigned char* pW1 = reinterpret_cast<const unsigned char*>(sW1);
      const unsigned char* pW2 = reinterpret_cast<const unsigned char*>(sW2);
      const unsigned char* pW3 = reinterpret_cast<const unsigned char*>(sW3);
      for (int tile0 = blockIdx.x * kMlpTile; tile0 < n_rows; tile0 += gridDim.x * kMlpTile) {
        for (int t = 0; t + 1 < h; ++t, ++n) {
          // layer 1: X -> D1
          mbar_wait(bar_x, n & 1);
          MLP_TRACE(n, 16);
          tc_fence_after();
#pragma unroll
          for (int ks = 0; ks < kMlpInPad / 16; ++ks)
            tc_mma_f16(tmem_base, umma_desc(pA + ks * 256, 128, sbo_in), umma_desc(pW1 + ks * 256, 128, sbo_in),
                        idesc_h, ks > 0);
          tc_commit(&bar_d[0]);
          // layer 2: H1 stages -> D2, accumulated as the stages land in shared memory
          for (int s = 0; s < S; ++s) {
            mbar_wait(&bar_a[s], 0);
            MLP_TRACE(n, 17 + s);
            tc_fence_after();
#pragma unroll
            for (int k4 = 0; k4 < kMlpStageCols / 16; ++k4) {
              const int ks = s * (kMlpStageCols / 16) + k4;
              tc_mma_f16(tmem_d2, umma_desc(pA + ks * 256, 128, sbo_h), umma_desc(pW2 + ks * 256, 128, sbo_h),
                          idesc_h, ks > 0);
            }
          }
          tc_commit(&bar_d[1]);
          // output layer: H2 stages -> D3 (over D1's first columns: D1 was consumed before any H2 stage existed)
          for (int s = 0; s < S; ++s) {
            mbar_wait(&bar_a[s], 1);
            MLP_TRACE(n, 21 + s);
            tc_fence_after();
#pragma unroll
            for (int k4 = 0; k4 < kMlpStageCols / 16; ++k4) {
              const int ks = s * (kMlpStageCols / 16) + k4;
              tc_mma_f16(tmem_base, umma_desc(pA + ks * 256, 128, sbo_h), umma_desc(pW3 + ks * 256, 128, sbo_h),
                          idesc_o, ks > 0);
            }
          }
          tc_commit(&bar_d[2]);
          MLP_TRACE(n, 25);
        }
      }
    }
    __syncwarp();
  } else {
    // ================= epilogue warps =================
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;               // this warp's 32 TMEM lanes
    unsigned char* rowA = reinterpret_cast<unsigned char*>(sA) + (size_t)(r & 7) * 16;   // + (r/8)*SBO + chunk*128
    unsigned char* dstX = rowA + (size_t)(r >> 3) * sbo_in;
    unsigned char* dstH = rowA + (size_t)(r >> 3) * sbo_h;

    for (int tile0 = blockIdx.x * kMlpTile; tile0 < n_rows; tile0 += gridDim.x * kMlpTile) {
      const int row = tile0 + r;
      const bool valid = row < n_rows;
      // fp32 state of this trajectory in input-column order [obs(od) | 0 | act(d) at act_off | 0].  Thread g keeps
      // only what it uses current: its own 8-column chunk (the X operand chunk it packs); g == 0 keeps the whole
      // observation (it evaluates the cost).  Action columns are (re)loaded by the threads whose chunk holds them.
      float x[kMlpInPad];
#pragma unroll
      for (int i = 0; i < kMlpInPad; ++i) x[i] = (valid && i < od) ? a.start_state[i] : 0.f;
      const float* acts = a.actions + (size_t)(valid ? row : 0) * a.stride;
      float total = (cc.reduce == 1) ? INFINITY : 0.f;
      const bool holds_act = g >= (act_off >> 3);        // warp-uniform: chunk g overlaps [act_off, act_off + d)
      // action of step t -> x[act_off .. act_off + d): act_off is a multiple of 8, so each case is static indexing.
      // Loads only: nothing consumes them before the next step's cost / pack, so their latency is never waited on.
#define ICEM_FOR_ACT(OFF, BODY) \
      _Pragma("unroll") for (int m = 0; m < kMlpInPad - OFF; ++m) if (m < d) { float& xa = x[OFF + m]; BODY; }
#define ICEM_ACT_CASES(BODY)                          \
      if (act_off == 24) { ICEM_FOR_ACT(24, BODY) }   \
      else if (act_off == 16) { ICEM_FOR_ACT(16, BODY) } \
      else { ICEM_FOR_ACT(8, BODY) }
      auto load_action = [&](int t) {
        if (!(holds_act || g == 0)) return;
        const float* at = acts + t * d;
        ICEM_ACT_CASES(xa = valid ? at[m] : 0.f)
      };
      load_action(0);

      for (int t = 0; t < h; ++t) {

        // ---- X row -> smem (fp16, K-major core matrices), layer 1: thread g writes 16-byte chunk g ----
        const bool last = t + 1 == h;          // the final predicted state is never scored: no transition after it
#pragma unroll
        for (int ch = 0; ch < kMlpInPad / 8; ++ch) {
          if (ch == g && !last) {
            __half2 p0 = __floats2half2_rn(x[8 * ch], x[8 * ch + 1]);
            __half2 p1 = __floats2half2_rn(x[8 * ch + 2], x[8 * ch + 3]);
            __half2 p2 = __floats2half2_rn(x[8 * ch + 4], x[8 * ch + 5]);
            __half2 p3 = __floats2half2_rn(x[8 * ch + 6], x[8 * ch + 7]);
            uint4 v;
            v.x = *reinterpret_cast<uint32_t*>(&p0); v.y = *reinterpret_cast<uint32_t*>(&p1);
            v.z = *reinterpret_cast<uint32_t*>(&p2); v.w = *reinterpret_cast<uint32_t*>(&p3);
            *reinterpret_cast<uint4*>(dstX + ch * 128) = v;
          }
        }
        if (!last) {
          fence_proxy_async_smem();
          tc_fence_before();                   // the out-epilogue's TMEM loads precede the next layer-1 MMA
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_x);
          MLP_TRACE(n, 0);
        }
        // ---- per-step cost on the pre-action observation: one thread per row, AFTER the X operand
        // is handed to the MMA issuer so that it is off the step's critical path ----
        if (g == 0) {
          float oa = 0.f, ob = 0.f, a2 = 0.f;
          ICEM_ACT_CASES(a2 = fmaf(xa, xa, a2))
#pragma unroll
          for (int i = 0; i < kMlpInPad; ++i) {
            oa = (i == cc.idx_a) ? x[i] : oa;
            ob = (i == cc.idx_b) ? x[i] : ob;
          }
          float c;
          if constexpr (kGoalCost) {
            float d2 = 0.f, e2 = 0.f;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              float gk = 0.f, ak = 0.f;
#pragma unroll
              for (int i = 0; i < kMlpInPad; ++i) {
                gk = (i == cc.goal_idx + k) ? x[i] : gk;
                ak = (i == cc.ach_idx + k) ? x[i] : ak;
              }
              d2 = fmaf(gk - ak, gk - ak, d2);
              e2 = fmaf(x[k] - x[3 + k], x[k] - x[3 + k], e2);
            }
            c = goal_distance_cost(cc, d2, e2);
          } else if (cc.kind == 0) {
            c = 0.1f * a2 - ob;
            if (cc.penalise_flipping)
              c += (oa > 1.5707963267948966f ? 10.f : 0.f) + (oa < -1.5707963267948966f ? 10.f : 0.f);
          } else {
            c = -oa + 0.1f * a2;
          }
          if (cc.reduce == 0) total += c;
          else if (cc.reduce == 1) total = fminf(total, c);
          else total = c;
        }
        if (last) break;

        // ---- hidden layers: epilogue stage by stage (bias + tanh -> fp16 A operand of the next MMA) ----
#pragma unroll 1
        for (int layer = 0; layer < 2; ++layer) {
          mbar_wait(&bar_d[layer], n & 1);
          MLP_TRACE(n, 1 + 6 * layer);
          tc_fence_after();
          const float* bs = sBias + layer * H;
          const uint32_t dcol = lane_addr + (layer == 0 ? tmem_base : tmem_d2);
          // TMEM loads run one stage ahead of the arithmetic: stage s+1's tcgen05.ld is in flight while stage s
          // goes through the MUFU / pack / store sequence (tcgen05.wait::ld waits for everything issued so far)
          uint32_t raw[2][16];
          tmem_ld16_issue(dcol + (uint32_t)(g * 16), raw[0]);
#pragma unroll
          for (int s = 0; s < kMlpMaxStages; ++s) {
            if (s < S) {
              const int c0 = s * kMlpStageCols + g * 16;
              tmem_ld_wait();
              if (s + 1 < S) tmem_ld16_issue(dcol + (uint32_t)(c0 + kMlpStageCols), raw[(s + 1) & 1]);
              const uint32_t* rv = raw[s & 1];
              const float4* b4 = reinterpret_cast<const float4*>(bs + c0);
              uint32_t pk[8];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float4 b = b4[j];
                __half2 lo2 = __floats2half2_rn(tanh_approx(__uint_as_float(rv[4 * j]) + b.x),
                                                           tanh_approx(__uint_as_float(rv[4 * j + 1]) + b.y));
                __half2 hi2 = __floats2half2_rn(tanh_approx(__uint_as_float(rv[4 * j + 2]) + b.z),
                                                           tanh_approx(__uint_as_float(rv[4 * j + 3]) + b.w));
                pk[2 * j] = *reinterpret_cast<uint32_t*>(&lo2);
                pk[2 * j + 1] = *reinterpret_cast<uint32_t*>(&hi2);
              }
              unsigned char* dst = dstH + (size_t)(c0 / 8) * 128;
              *reinterpret_cast<uint4*>(dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              *reinterpret_cast<uint4*>(dst + 128) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
              fence_proxy_async_smem();
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&bar_a[s]);
              MLP_TRACE(n, 2 + 6 * layer + s);
            }
          }
        }
        // the next step's action.  Issued HERE, after this step's last proxy fence: fence.proxy.async is a
        // MEMBAR that waits for every outstanding global load of the thread (measured: ~1000 cycles per step
        // when the loads were issued before the epilogues); now they land while the output MMAs drain.
        load_action(t + 1);

        // ---- output layer epilogue: obs += W3 h2 + b3 (fp32 state in registers).  W3 rows / b3 entries past
        // obs_dim are zero, so the padding and action columns receive + 0. ----
        mbar_wait(&bar_d[2], n & 1);
        MLP_TRACE(n, 13);
        tc_fence_after();
        if (g == 0) {
          float v[32];
          tmem_ld32(lane_addr + tmem_base, v);
          const float4* b3 = reinterpret_cast<const float4*>(sBias + 2 * H);
#pragma unroll
          for (int j = 0; j < kMlpOutPad / 4; ++j) {
            const float4 b = b3[j];
            x[4 * j] += v[4 * j] + b.x;         x[4 * j + 1] += v[4 * j + 1] + b.y;
            x[4 * j + 2] += v[4 * j + 2] + b.z; x[4 * j + 3] += v[4 * j + 3] + b.w;
          }
        } else {
          float v[8];
          tmem_ld8(lane_addr + tmem_base + (uint32_t)(8 * g), v);
          const float4* b3 = reinterpret_cast<const float4*>(sBias + 2 * H + 8 * g);
          const float4 b0 = b3[0], b1 = b3[1];
#pragma unroll
          for (int ch = 1; ch < kMlpInPad / 8; ++ch) {
            if (ch == g) {
              x[8 * ch] += v[0] + b0.x;     x[8 * ch + 1] += v[1] + b0.y;
              x[8 * ch + 2] += v[2] + b0.z; x[8 * ch + 3] += v[3] + b0.w;
              x[8 * ch + 4] += v[4] + b1.x; x[8 * ch + 5] += v[5] + b1.y;
              x[8 * ch + 6] += v[6] + b1.z; x[8 * ch + 7] += v[7] + b1.w;
            }
          }
        }
        MLP_TRACE(n, 14);
        ++n;
      }
      if (valid && g == 0) a.costs[row] = total;
    }
#undef ICEM_ACT_CASES
#undef ICEM_FOR_ACT
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, kMlpTmemCols);
}

// one transition of the same model for a single state (env.step on the device model / closed-loop bench):
// same fp16 operand roundings as the tensor-core path, CUDA-core arithmetic
__global__ void mlp_advance_kernel(MlpParams mp, const float* state, const float* action, float* next_state,
                                   float* obs_out, int obs_dim_out) {
  __shared__ float xin[kMlpInPad];
  __shared__ float h1[256], h2[256];
  const int H = mp.hidden, od = mp.obs_dim, d = mp.act_dim, tid = threadIdx.x;
  auto bf = [](float v) { return __half2float(__float2half_rn(v)); };
  const int ao = mp.act_off;
  if (tid < kMlpInPad)
    xin[tid] = tid < od ? state[tid] : (action && tid >= ao && tid < ao + d ? action[tid - ao] : 0.f);
  __syncthreads();
  if (action) {
    for (int n = tid; n < H; n += blockDim.x) {
      float acc = 0.f;
      for (int k = 0; k < kMlpInPad; ++k) acc = fmaf(bf(xin[k]), __half2float(mp.w1[umma_pack_index(n, k, kMlpInPad)]), acc);
      h1[n] = bf(tanh_approx(acc + mp.bias[n]));
    }
    __syncthreads();
    for (int n = tid; n < H; n += blockDim.x) {
      float acc = 0.f;
      for (int k = 0; k < H; ++k) acc = fmaf(h1[k], __half2float(mp.w2[umma_pack_index(n, k, H)]), acc);
      h2[n] = bf(tanh_approx(acc + mp.bias[H + n]));
    }
    __syncthreads();
    if (tid < od) {
      float acc = 0.f;
      for (int k = 0; k < H; ++k) acc = fmaf(h2[k], __half2float(mp.w3[umma_pack_index(tid, k, H)]), acc);
      xin[tid] = xin[tid] + acc + mp.bias[2 * H + tid];
    }
    __syncthreads();
  }
  if (next_state && tid < od) next_state[tid] = xin[tid];
  if (obs_out && tid < obs_dim_out) obs_out[tid] = xin[tid];
}

}  // namespace icem
