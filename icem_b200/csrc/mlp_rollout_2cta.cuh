// Tensor-core MLP rollout on a CTA PAIR (tcgen05 cta_group::2): two 128-row tiles per SM in flight.
//
// Same model, same arithmetic and the same roundings as mlp_rollout.cuh (reference path: the batched `predict` loop of
// icem/models/abstract_models.py:17-53 + per-step cost, controllers/abstract_controller.py:74-91).  What changes is
// the schedule.  One 128-row tile per SM (mlp_rollout.cuh) leaves the SM idle along the per-step dependency chain
// MMA -> tanh epilogue -> MMA -> tanh epilogue -> MMA -> state update (measured: 8.9 k cycles per step against 4.1 k
// of MUFU work); a second tile cannot move in because the resident weights take 160 KB.  With `cta_group::2` the two
// CTAs of a cluster issue ONE M=256 MMA whose B operand is split between them: each CTA keeps only HALF of every
// weight matrix (80 KB), which leaves room for the activations of TWO tiles (2 x 64 KB).  The 16 epilogue warps of a
// CTA then alternate between the tiles -- while they run the tanh epilogue of tile A the tensor cores run the MMAs
// of tile B -- so the chain of one tile hides behind the work of the other.
//
//   cluster = 2 CTAs (leader = rank 0);  super-tile = 512 trajectories = 2 slots x 2 CTAs x 128 rows
//   row of (slot s, CTA r, lane-row rr) = base + 256 s + 128 r + rr
//   TMEM per CTA: slot s owns columns [256 s, 256 s + 256): D1 -> D2 -> D3 reuse them in turn (a slot's layers are
//   sequential; the overlap is ACROSS slots)
//   one issuer thread (leader CTA, warp 16) feeds all MMAs; its barriers bar_x[s] / bar_h[s] live in the leader's
//   shared memory and collect 16 warp arrivals from each CTA (remote arrivals through mapa + mbarrier.arrive
//   .shared::cluster); completion comes back to both CTAs with tcgen05.commit ... multicast::cluster.
//
// Per-row state is sliced instead of replicated: thread (row, g) keeps only the 8 input columns 8g..8g+7 of its
// row (the chunk of the X operand it packs) for both slots; the per-step cost is split into the terms each chunk can
// evaluate (action norm, angle / height term, velocity term) and the partial sums are added in a fixed order at the
// end.  That needs the cost to be a SUM over steps: cost_along_trajectory "best" / "final" use mlp_rollout_kernel.
#pragma once
#include "mlp_rollout.cuh"

namespace icem {

constexpr int kMlp2Slots = 2;
constexpr int kMlp2SuperTile = kMlpTile * 2 * kMlp2Slots;      // rows per cluster iteration

struct MlpHalfParams {          // per-CTA halves of the packed weights (rank 0: output rows [0, N/2), rank 1: the rest)
  const __nv_bfloat16* w1[2];   // [hidden/2 x kMlpInPad]
  const __nv_bfloat16* w2[2];   // [hidden/2 x hidden]
  const __nv_bfloat16* w3[2];   // [kMlpOutPad/2 x hidden]
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address of this CTA -> shared::cluster address of the same offset in CTA `rank`
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// Arrive on a barrier of another CTA of the cluster.  Default semantics (release at CTA scope), as CUTLASS's
// ClusterBarrier::arrive(cta_id) does: what the arrival publishes is this CTA's OWN shared memory, already made
// visible to its own async proxy by fence.proxy.async; the consumer is this SM's own tensor core, triggered by the
// leader's MMA.  (`.release.cluster` compiles to MEMBAR.ALL.GPU + ERRBAR per arrival: measured 2x slower kernel.)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" :: "r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAITC_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONEC_%=;\n\t"
      "bra WAITC_%=;\n\t"
      "DONEC_%=:\n\t}"
      :: "r"(smem_addr(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
               :: "r"(smem_addr(smem_dst)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A[each CTA's smem: its 128 rows] * B[each CTA's smem: its half of N]^T   (M = 256)
__device__ __forceinline__ void tc_mma2_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  const uint32_t z = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
      :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(z) : "memory");
}
// completion of all MMAs issued so far -> the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void tc_commit2(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               :: "r"(smem_addr(bar)), "h"(mask) : "memory");
}

inline size_t mlp2_smem_bytes(int hidden) {
  const size_t w = ((size_t)(hidden / 2) * kMlpInPad + (size_t)(hidden / 2) * hidden +
                    (size_t)(kMlpOutPad / 2) * hidden) * 2;
  const size_t a = (size_t)kMlp2Slots * kMlpTile * hidden * 2;
  const size_t b = ((size_t)2 * hidden + kMlpOutPad) * 4;
  const size_t part = (size_t)kMlp2Slots * kMlpTile * kMlpColGroups * 4;      // partial cost sums
  return w + a + b + part + 128;
}

// ---- per-slot pieces of a control step, as force-inlined functions over that slot's own register arrays -------
// chunk g of the row's input vector <- action of step t (columns [act_off, act_off + d))
__device__ __forceinline__ void mlp2_load_action(float (&x)[8], const float* acts, bool valid, int t, int g, int d,
                                                 int act_off) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = 8 * g + i - act_off;
    if (m >= 0 && m < d) x[i] = valid ? acts[t * d + m] : 0.f;
  }
}
// the terms of the step cost that chunk g can evaluate (pre-action observation)
__device__ __forceinline__ float mlp2_cost_terms(const float (&x)[8], const CostConst& cc, int g, int d, int act_off,
                                                 bool has_act) {
  float c = 0.f;
  if (has_act) {
    float a2 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int m = 8 * g + i - act_off;
      if (m >= 0 && m < d) a2 = fmaf(x[i], x[i], a2);
    }
    c = 0.1f * a2;
  }
  if (g == (cc.idx_a >> 3)) {
    float oa = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) oa = (i == (cc.idx_a & 7)) ? x[i] : oa;
    if (cc.kind == 0) {
      if (cc.penalise_flipping)
        c += (oa > 1.5707963267948966f ? 10.f : 0.f) + (oa < -1.5707963267948966f ? 10.f : 0.f);
    } else {
      c -= oa;
    }
  }
  if (cc.kind == 0 && g == (cc.idx_b >> 3)) {
    float ob = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) ob = (i == (cc.idx_b & 7)) ? x[i] : ob;
    c -= ob;
  }
  return c;
}
__device__ __forceinline__ uint4 mlp2_pack8(const float (&x)[8]) {
  __nv_bfloat162 p0 = __floats2bfloat162_rn(x[0], x[1]);
  __nv_bfloat162 p1 = __floats2bfloat162_rn(x[2], x[3]);
  __nv_bfloat162 p2 = __floats2bfloat162_rn(x[4], x[5]);
  __nv_bfloat162 p3 = __floats2bfloat162_rn(x[6], x[7]);
  uint4 v;
  v.x = *reinterpret_cast<uint32_t*>(&p0); v.y = *reinterpret_cast<uint32_t*>(&p1);
  v.z = *reinterpret_cast<uint32_t*>(&p2); v.w = *reinterpret_cast<uint32_t*>(&p3);
  return v;
}

#ifdef ICEM_MLP_TRACE
#define MLP2_TRACE(step, slot) do { if (blockIdx.x == 0 && (step) < 64u && lane == 0 && q == 0 && (g == 0 || g == kMlpColGroups)) \
    g_mlp_trace[(step) * 32 + (slot)] = clock64(); } while (0)
#else
#define MLP2_TRACE(step, slot) do { } while (0)
#endif

// ---------------------------------------------------------------------------------------------------------------
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kMlpThreads, 1)
mlp_rollout_2cta_kernel(RolloutArgs a, SamplerConst sc, CostConst cc, MlpParams mp, MlpHalfParams hp) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const int H = mp.hidden;
  const uint32_t rank = cluster_ctarank();
  __nv_bfloat16* sW1 = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  __nv_bfloat16* sW2 = sW1 + (size_t)(H / 2) * kMlpInPad;
  __nv_bfloat16* sW3 = sW2 + (size_t)(H / 2) * H;
  __nv_bfloat16* sA0 = sW3 + (size_t)(kMlpOutPad / 2) * H;            // [slot][128 x max(H, 32)] activations
  float* sBias = reinterpret_cast<float*>(sA0 + (size_t)kMlp2Slots * kMlpTile * H);
  float* sPart = sBias + 2 * H + kMlpOutPad;                          // [slot][128][4] partial cost sums
  uint64_t* bar_x = reinterpret_cast<uint64_t*>(sPart + kMlp2Slots * kMlpTile * kMlpColGroups);   // [2] leader's
  uint64_t* bar_h = bar_x + kMlp2Slots;                               // [2] leader's: activations of a slot ready
  uint64_t* bar_d = bar_h + kMlp2Slots;                               // [2] per CTA: accumulator of a slot ready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_d + kMlp2Slots);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, g = warp >> 2;             // TMEM lane quarter, column group (g == 4: the issuer warp)
  const int r = q * 32 + lane;                       // trajectory row inside the tile
  const int h = sc.h, d = sc.d, od = mp.obs_dim, act_off = mp.act_off;

  // ---- one-time: this CTA's halves of the weights, barriers, TMEM ----
  {
    const uint4* src1 = reinterpret_cast<const uint4*>(hp.w1[rank]);
    uint4* dst1 = reinterpret_cast<uint4*>(sW1);
    for (int i = tid; i < (H / 2) * kMlpInPad / 8; i += blockDim.x) dst1[i] = src1[i];
    const uint4* src2 = reinterpret_cast<const uint4*>(hp.w2[rank]);
    uint4* dst2 = reinterpret_cast<uint4*>(sW2);
    for (int i = tid; i < (H / 2) * H / 8; i += blockDim.x) dst2[i] = src2[i];
    const uint4* src3 = reinterpret_cast<const uint4*>(hp.w3[rank]);
    uint4* dst3 = reinterpret_cast<uint4*>(sW3);
    for (int i = tid; i < (kMlpOutPad / 2) * H / 8; i += blockDim.x) dst3[i] = src3[i];
    for (int i = tid; i < 2 * H + kMlpOutPad; i += blockDim.x) sBias[i] = mp.bias[i];
  }
  if (tid == 0) {
    for (int s = 0; s < kMlp2Slots; ++s) {
      mbar_init(&bar_x[s], 2 * kMlpEpiThreads / 32);      // 16 warps of each CTA
      mbar_init(&bar_h[s], 2 * kMlpEpiThreads / 32);
      mbar_init(&bar_d[s], 1);
    }
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc2(tmem_slot, kMlpTmemCols);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                  // both CTAs' barriers and weights exist before any arrival / MMA
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const uint32_t idesc_h = umma_idesc_bf16(2 * kMlpTile, H);
  const uint32_t idesc_o = umma_idesc_bf16(2 * kMlpTile, kMlpOutPad);
  const uint32_t sbo_in = (kMlpInPad / 8) * 128, sbo_h = (uint32_t)(H / 8) * 128;
  const size_t slot_bytes = (size_t)kMlpTile * H * 2;

  const StepState ss = *a.ss;
  const int n_rows = a.n_fresh_local + ((a.iteration == 0 && ss.has_prev_elites) ? a.n_shift_local : 0);
  const int n_clusters = gridDim.x >> 1, cluster_id = blockIdx.x >> 1;

  if (g == kMlpColGroups) {
    // ================= MMA issuer: one thread of the leader CTA =================
    if (rank == 0 && lane == 0) {
      const unsigned char* pW1 = reinterpret_cast<const unsigned char*>(sW1);
      const unsigned char* pW2 = reinterpret_cast<const unsigned char*>(sW2);
      const unsigned char* pW3 = reinterpret_cast<const unsigned char*>(sW3);
      uint32_t n = 0;
      for (int base = cluster_id * kMlp2SuperTile; base < n_rows; base += n_clusters * kMlp2SuperTile) {
        for (int t = 0; t + 1 < h; ++t, ++n) {
          for (int s = 0; s < kMlp2Slots; ++s) {       // layer 1: X -> D[s]
            const unsigned char* pA = reinterpret_cast<const unsigned char*>(sA0) + s * slot_bytes;
            mbar_wait_cluster(&bar_x[s], n & 1);
            MLP2_TRACE(n, 16 + s);
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < kMlpInPad / 16; ++ks)
              tc_mma2_bf16(tmem_base + 256u * s, umma_desc(pA + ks * 256, 128, sbo_in),
                           umma_desc(pW1 + ks * 256, 128, sbo_in), idesc_h, ks > 0);
            tc_commit2(&bar_d[s]);
          }
          for (int layer = 0; layer < 2; ++layer) {    // layer 2 (H1 -> D[s]) and output layer (H2 -> D[s][:, :32])
            for (int s = 0; s < kMlp2Slots; ++s) {
              const unsigned char* pA = reinterpret_cast<const unsigned char*>(sA0) + s * slot_bytes;
              mbar_wait_cluster(&bar_h[s], (uint32_t)layer);
              MLP2_TRACE(n, 18 + 2 * layer + s);
              tc_fence_after();
              const unsigned char* wB = layer == 0 ? pW2 : pW3;
              const uint32_t id = layer == 0 ? idesc_h : idesc_o;
              for (int ks = 0; ks < H / 16; ++ks)
                tc_mma2_bf16(tmem_base + 256u * s, umma_desc(pA + ks * 256, 128, sbo_h),
                             umma_desc(wB + ks * 256, 128, sbo_h), id, ks > 0);
              tc_commit2(&bar_d[s]);
            }
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ================= epilogue warps (both CTAs) =================
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const uint32_t leader_x = mapa_shared(smem_addr(&bar_x[0]), 0), leader_h = mapa_shared(smem_addr(&bar_h[0]), 0);
    const int ca = act_off >> 3;                        // first input chunk holding action columns
    const bool has_act = g >= ca;                       // (warp-uniform) chunk g overlaps [act_off, act_off + d)
    uint32_t dph = 0;                                   // bit s: parity of the next completion of bar_d[s]
    uint32_t n = 0;
    unsigned char* const sAbytes = reinterpret_cast<unsigned char*>(sA0);

    // Everything per slot is written once as a lambda over that slot's OWN scalars / arrays (xA, xB, ...): arrays
    // indexed by a slot variable end up in local memory.
    auto pack_x = [&](const uint4 v, int s) {            // X chunk -> smem, hand the slot's X operand to the issuer
      unsigned char* dstX = sAbytes + s * slot_bytes + (size_t)(r & 7) * 16 + (size_t)(r >> 3) * sbo_in;
      *reinterpret_cast<uint4*>(dstX + g * 128) = v;
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(leader_x + 8u * s);
    };
    // tanh epilogue of one slot: D[slot] -> bf16 activations in the slot's smem buffer -> arrive on the leader
    auto hidden = [&](int s, int layer) {
      const float* bs = sBias + layer * H;
      mbar_wait(&bar_d[s], (dph >> s) & 1u);
      dph ^= 1u << s;
      MLP2_TRACE(n, 1 + 4 * layer + 2 * s);
      tc_fence_after();
      const uint32_t dcol = lane_addr + tmem_base + 256u * s;
      unsigned char* dstH = sAbytes + s * slot_bytes + (size_t)(r & 7) * 16 + (size_t)(r >> 3) * sbo_h;
      // column group g owns the CONTIGUOUS hidden columns [g H/4, (g+1) H/4): 16 at a time, the next 16 in flight
      const int cg0 = g * (H / 4), nch = H / 64;          // 16-column chunks per thread: 1, 2 or 4
      uint32_t ra[16], rb[16];
      auto finish16 = [&](const uint32_t* rv, int c0) {
        const float4* b4 = reinterpret_cast<const float4*>(bs + c0);
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 b = b4[j];
          __nv_bfloat162 lo2 = __floats2bfloat162_rn(tanh_approx(__uint_as_float(rv[4 * j]) + b.x),
                                                     tanh_approx(__uint_as_float(rv[4 * j + 1]) + b.y));
          __nv_bfloat162 hi2 = __floats2bfloat162_rn(tanh_approx(__uint_as_float(rv[4 * j + 2]) + b.z),
                                                     tanh_approx(__uint_as_float(rv[4 * j + 3]) + b.w));
          pk[2 * j] = *reinterpret_cast<uint32_t*>(&lo2);
          pk[2 * j + 1] = *reinterpret_cast<uint32_t*>(&hi2);
        }
        unsigned char* dst = dstH + (size_t)(c0 / 8) * 128;
        *reinterpret_cast<uint4*>(dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        *reinterpret_cast<uint4*>(dst + 128) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
      };
      tmem_ld16_issue(dcol + (uint32_t)cg0, ra);
      tmem_ld_wait();
      if (nch > 1) tmem_ld16_issue(dcol + (uint32_t)(cg0 + 16), rb);
      finish16(ra, cg0);
      if (nch > 1) {
        tmem_ld_wait();
        if (nch > 2) tmem_ld16_issue(dcol + (uint32_t)(cg0 + 32), ra);
        finish16(rb, cg0 + 16);
      }
      if (nch > 2) {
        tmem_ld_wait();
        tmem_ld16_issue(dcol + (uint32_t)(cg0 + 48), rb);
        finish16(ra, cg0 + 32);
        tmem_ld_wait();
        finish16(rb, cg0 + 48);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(leader_h + 8u * s);
      MLP2_TRACE(n, 2 + 4 * layer + 2 * s);
    };
    auto output = [&](int s, float (&v)[8]) {            // D3 columns of this thread's chunk, + b3
      mbar_wait(&bar_d[s], (dph >> s) & 1u);
      dph ^= 1u << s;
      MLP2_TRACE(n, 9 + 2 * s);
      tc_fence_after();
      tmem_ld8(lane_addr + tmem_base + 256u * s + (uint32_t)(8 * g), v);
      const float4* b3 = reinterpret_cast<const float4*>(sBias + 2 * H + 8 * g);
      const float4 b0 = b3[0], b1 = b3[1];
      v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
      v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
      MLP2_TRACE(n, 10 + 2 * s);
    };

    for (int base = cluster_id * kMlp2SuperTile; base < n_rows; base += n_clusters * kMlp2SuperTile) {
      float xA[8], xB[8];                               // this thread's 8 input columns of its row, per slot
      float totA = 0.f, totB = 0.f;
      const int rowA = base + 128 * (int)rank + r, rowB = rowA + 256;
      const bool validA = rowA < n_rows, validB = rowB < n_rows;
      const float* actsA = a.actions + (size_t)(validA ? rowA : 0) * a.stride;
      const float* actsB = a.actions + (size_t)(validB ? rowB : 0) * a.stride;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = 8 * g + i;
        const float v0 = c < od ? a.start_state[c] : 0.f;
        xA[i] = validA ? v0 : 0.f;
        xB[i] = validB ? v0 : 0.f;
      }
      if (has_act) {
        mlp2_load_action(xA, actsA, validA, 0, g, d, act_off);
        mlp2_load_action(xB, actsB, validB, 0, g, d, act_off);
      }

      for (int t = 0; t < h; ++t) {
        const bool last = t + 1 == h;                   // the final predicted state is never scored: no transition
        if (!last) {
          pack_x(mlp2_pack8(xA), 0);
          pack_x(mlp2_pack8(xB), 1);
          MLP2_TRACE(n, 0);
        }
        if (last) {                                     // pre-action observation
          totA += mlp2_cost_terms(xA, cc, g, d, act_off, has_act);
          totB += mlp2_cost_terms(xB, cc, g, d, act_off, has_act);
          break;
        }
        // the epilogue of one slot runs while the tensor cores work on the other slot: A1 B1 A2 B2 (ONE copy of the
        // epilogue code, looped: four inlined copies spill)
#pragma unroll 1
        for (int k = 0; k < 4; ++k) {
          hidden(k & 1, k >> 1);
          // After the slot's LAST proxy fence of the step (a MEMBAR that waits for outstanding global loads): its
          // share of the step cost -- x is still the pre-action observation, and this is off the X -> MMA critical
          // path -- then the next action, which lands while the output MMAs run
          if (k == 2) {
            totA += mlp2_cost_terms(xA, cc, g, d, act_off, has_act);
            if (has_act) mlp2_load_action(xA, actsA, validA, t + 1, g, d, act_off);
          }
          if (k == 3) {
            totB += mlp2_cost_terms(xB, cc, g, d, act_off, has_act);
            if (has_act) mlp2_load_action(xB, actsB, validB, t + 1, g, d, act_off);
          }
        }
        {
          float v[8];
          output(0, v);
#pragma unroll
          for (int i = 0; i < 8; ++i) xA[i] += v[i];
          output(1, v);
#pragma unroll
          for (int i = 0; i < 8; ++i) xB[i] += v[i];
        }
        ++n;
      }
      // ---- per-row cost = the chunks' partial sums, added in a fixed order ----
      sPart[(0 * kMlpTile + r) * kMlpColGroups + g] = totA;
      sPart[(1 * kMlpTile + r) * kMlpColGroups + g] = totB;
      asm volatile("bar.sync 1, %0;" :: "r"(kMlpEpiThreads) : "memory");      // the 16 epilogue warps only
      if (g == 0) {
        const float4 pa = *reinterpret_cast<const float4*>(&sPart[(0 * kMlpTile + r) * kMlpColGroups]);
        const float4 pb = *reinterpret_cast<const float4*>(&sPart[(1 * kMlpTile + r) * kMlpColGroups]);
        if (validA) a.costs[rowA] = ((pa.x + pa.y) + pa.z) + pa.w;
        if (validB) a.costs[rowB] = ((pb.x + pb.y) + pb.z) + pb.w;
      }
      asm volatile("bar.sync 1, %0;" :: "r"(kMlpEpiThreads) : "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                                  // the peer may still be reading this CTA's barriers / TMEM
  if (warp == 0) tmem_dealloc2(tmem_base, kMlpTmemCols);
}

}  // namespace icem
