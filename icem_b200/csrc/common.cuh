// Shared device/host helpers for libicem_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <cstdio>
#include <stdexcept>
#include <string>

#define ICEM_WARP 32

namespace icem {

// ---------------------------------------------------------------------------------------------
// host-side error plumbing: CUDA failures become C++ exceptions, translated to status codes at
// the C-ABI boundary (planner.cu).
struct CudaError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

inline void cuda_check(cudaError_t e, const char* what, const char* file, int line) {
  if (e != cudaSuccess) {
    char buf[512];
    snprintf(buf, sizeof buf, "%s failed: %s (%s:%d)", what, cudaGetErrorString(e), file, line);
    throw CudaError(buf);
  }
}
#define ICEM_CUDA(x) ::icem::cuda_check((x), #x, __FILE__, __LINE__)

// ---------------------------------------------------------------------------------------------
// warp helpers
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// minimum of a 64-bit key over the warp with two 32-bit REDUX instructions (high words, then the low words of the
// lanes that hold the minimal high word) instead of a 5-step shuffle tree of 64-bit values
__device__ __forceinline__ unsigned long long warp_min_u64(unsigned long long v) {
  const unsigned hi = (unsigned)(v >> 32), lo = (unsigned)v;
  const unsigned mh = __reduce_min_sync(0xffffffffu, hi);
  const unsigned ml = __reduce_min_sync(0xffffffffu, hi == mh ? lo : 0xffffffffu);
  return ((unsigned long long)mh << 32) | ml;
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-R counter-based RNG (Salmon et al., SC'11): one call -> 4 x uint32, no state.  R = 10 is the
// authors' default; R = 7 is the smallest round count for which they report the generator Crush-resistant
// (passes TestU01 BigCrush) and is what the stand-alone sampler uses.
struct Philox4 {
  uint32_t x, y, z, w;
};

template <int ROUNDS>
__host__ __device__ __forceinline__ Philox4 philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                        uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < ROUNDS; ++r) {
    const uint64_t p0 = (uint64_t)M0 * c0;
    const uint64_t p1 = (uint64_t)M1 * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  return Philox4{c0, c1, c2, c3};
}

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                           uint32_t k0, uint32_t k1) {
  return philox4x32<10>(c0, c1, c2, c3, k0, k1);
}

// two uniforms -> two standard normals (Box-Muller, fast intrinsics: sampling noise, not a parity path)
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& n0, float& n1) {
  const float u1 = ((float)a + 0.5f) * 2.3283064365386963e-10f;   // (0,1]
  const float u2 = ((float)b + 0.5f) * 2.3283064365386963e-10f;
  const float r = sqrtf(-2.0f * __logf(u1));
  float s, c;
  __sincosf(6.283185307179586f * u2, &s, &c);
  n0 = r * c;
  n1 = r * s;
}

// ---------------------------------------------------------------------------------------------
// TMA (bulk async copy) + mbarrier wrappers: 1-D cp.async.bulk, SASS UBLKCP.
__device__ __forceinline__ uint32_t smem_addr(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// shared -> global bulk store of `bytes` (multiple of 16; both addresses 16-B aligned)
__device__ __forceinline__ void tma_store_1d(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               :: "l"(gdst), "r"(smem_addr(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until the smem SOURCE of all committed bulk stores has been read (safe to overwrite it)
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
               :: "r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}"
      :: "r"(smem_addr(bar)), "r"(parity) : "memory");
}
// global -> shared bulk load, completion signalled on `bar` (complete_tx)
__device__ __forceinline__ void tma_load_1d(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(smem_addr(sdst)), "l"(gsrc), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}

// order-preserving key for (cost, index): ascending u64 order == ascending (cost, index) with
// NaN mapped above +inf (np.argsort puts NaN last).  -0.0 is canonicalised to +0.0 so that
// equal costs tie-break on the index like a stable sort.
__host__ __device__ __forceinline__ unsigned long long cost_key(float c, uint32_t idx) {
  uint32_t b;
#ifdef __CUDA_ARCH__
  b = __float_as_uint(c);
#else
  union { float f; uint32_t u; } cv; cv.f = c; b = cv.u;
#endif
  if (c != c) b = 0x7FC00000u;          // canonical positive NaN > +inf
  if (b == 0x80000000u) b = 0u;         // -0 == +0
  b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
  return ((unsigned long long)b << 32) | idx;
}
__host__ __device__ __forceinline__ uint32_t key_index(unsigned long long k) { return (uint32_t)k; }

}  // namespace icem
