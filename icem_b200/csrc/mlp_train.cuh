// Device trainer of the dense MLP forward model that csrc/mlp_rollout.cuh rolls out:
//     next_obs = obs + W3 tanh(W2 tanh(W1 [obs, act] + b1) + b2) + b3
// fitted to (obs, act) -> next_obs - obs transitions by minibatch Adam on the mean squared error.
//
// Replaces the hook the reference calls once per training iteration (paths relative to /root/reference/icem/):
//   main.py:209-210   forward_model.train(rollout_buffer)
// The reference ships no trainable model (models/__init__.py:5-8 registers ground-truth models only), so
// there is no reference arithmetic to match: the oracle is a PyTorch fp32 nn.Sequential + MSELoss + torch.optim.Adam
// on the same minibatches (oracle/mlp_train_torch.py) -- PARITY UNPINNED against the reference by construction.
//
// All arithmetic is fp32 (the trained weights are rounded to fp16 operands only when the planner packs them for the
// tensor cores).  One training step = 3 forward GEMMs (bias + tanh fused), the loss gradient, 2 backward-data GEMMs
// (tanh' fused), 3 weight-gradient GEMMs (split over the batch dimension into partial buffers that the Adam kernel
// sums in a fixed order: deterministic, no atomics), 3 bias-gradient column sums, 6 Adam updates.  The GEMM is a plain
// shared-memory tiled fp32 kernel (64x64 tile, 4x4 per thread): at 24 -> 256 -> 256 -> 18 a step is ~2 GFLOP per
// 4096-row batch and launch-latency bound; it is not on the planner's hot path.
#pragma once
#include "common.cuh"

namespace icem {

enum { kTrEpNone = 0, kTrEpBiasTanh = 1, kTrEpBias = 2, kTrEpTanhGrad = 3 };
constexpr int kTrTile = 64, kTrKT = 16;

// C = op(A) op(B):  TA ? A is [K][M] : [M][K];  TB ? B is [N][K] : [K][N]  (row-major, leading dimensions lda / ldb).
// blockIdx.z = split of the K range (length kchunk); split z writes its partial product to C + z * M * ldc.
template <bool TA, bool TB, int EP>
__global__ void __launch_bounds__(256) train_gemm_kernel(int M, int N, int K, const float* __restrict__ A, int lda,
                                                         const float* __restrict__ B, int ldb, float* __restrict__ C,
                                                         int ldc, const float* __restrict__ bias,
                                                         const float* __restrict__ aux, int ldaux, int kchunk) {
  __shared__ float sA[kTrKT][kTrTile + 4];     // [k][m]
  __shared__ float sB[kTrKT][kTrTile + 4];     // [k][n]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * kTrTile, n0 = blockIdx.x * kTrTile;
  const int k_lo = blockIdx.z * kchunk, k_hi = min(K, k_lo + kchunk);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = k_lo; k0 < k_hi; k0 += kTrKT) {
    // stage 16 x 64 of op(A) and op(B): 1024 elements each, 4 per thread, coalesced along the contiguous dimension
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int i = tid + e * 256;
      int kk, mm;
      if (TA) { kk = i >> 6; mm = i & 63; } else { mm = i >> 4; kk = i & 15; }
      const int gk = k0 + kk, gm = m0 + mm;
      float v = 0.f;
      if (gk < k_hi && gm < M) v = TA ? A[(size_t)gk * lda + gm] : A[(size_t)gm * lda + gk];
      sA[kk][mm] = v;
      int kb, nn;
      if (TB) { nn = i >> 4; kb = i & 15; } else { kb = i >> 6; nn = i & 63; }
      const int gkb = k0 + kb, gn = n0 + nn;
      float w = 0.f;
      if (gkb < k_hi && gn < N) w = TB ? B[(size_t)gn * ldb + gkb] : B[(size_t)gkb * ldb + gn];
      sB[kb][nn] = w;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kTrKT; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&sA[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&sB[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  float* Cz = C + (size_t)blockIdx.z * M * ldc;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = acc[i][j];
      if (EP == kTrEpBiasTanh) v = tanhf(v + bias[n]);
      if (EP == kTrEpBias) v = v + bias[n];
      if (EP == kTrEpTanhGrad) { const float a = aux[(size_t)m * ldaux + n]; v = v * (1.f - a * a); }
      Cz[(size_t)m * ldc + n] = v;
    }
  }
}

// minibatch rows: xb[b][:] = x[idx[b]][:], tb[b][:] = t[idx[b]][:]
__global__ void train_gather_kernel(int batch, int in, int out, const float* __restrict__ x, const float* __restrict__ t,
                                    const int* __restrict__ idx, float* __restrict__ xb, float* __restrict__ tb) {
  const int b = blockIdx.x;
  const size_t r = (size_t)idx[b];
  for (int i = threadIdx.x; i < in; i += blockDim.x) xb[(size_t)b * in + i] = x[r * in + i];
  for (int i = threadIdx.x; i < out; i += blockDim.x) tb[(size_t)b * out + i] = t[r * out + i];
}

// dY = 2 (Y - T) / count and the squared error; block partial sums (fixed order) -> part[blockIdx.x]
__global__ void __launch_bounds__(256) train_loss_grad_kernel(int count, const float* __restrict__ y,
                                                              const float* __restrict__ t, float* __restrict__ dy,
                                                              float* __restrict__ part) {
  __shared__ float red[256];
  const int i = blockIdx.x * 256 + threadIdx.x;
  float e = 0.f;
  if (i < count) {
    const float d = y[i] - t[i];
    dy[i] = 2.f * d / (float)count;
    e = d * d;
  }
  red[threadIdx.x] = e;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[blockIdx.x] = red[0];
}

// loss[0] = sum(part[0..n)) / count, summed in index order by one thread block
__global__ void __launch_bounds__(256) train_loss_reduce_kernel(int n, int count, const float* __restrict__ part,
                                                                float* __restrict__ loss) {
  __shared__ float red[256];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) s += part[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss[0] = red[0] / (float)count;
}

// db[n] = sum over the batch rows of dz[b][n]; block = 32 columns x 8 row lanes
__global__ void __launch_bounds__(256) train_colsum_kernel(int rows, int cols, const float* __restrict__ dz, int ld,
                                                           float* __restrict__ db) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + (threadIdx.x & 31), rl = threadIdx.x >> 5;
  float s = 0.f;
  if (c < cols)
    for (int r = rl; r < rows; r += 8) s += dz[(size_t)r * ld + c];
  red[rl][threadIdx.x & 31] = s;
  __syncthreads();
  if (rl == 0 && c < cols) {
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) v += red[k][threadIdx.x & 31];
    db[c] = v;
  }
}

// torch.optim.Adam (no amsgrad): g = sum of the gradient partials (+ wd * p), m / v moments, bias-corrected step
__global__ void train_adam_kernel(int n, float* __restrict__ p, const float* __restrict__ g, int splits, size_t gstride,
                                  float* __restrict__ m, float* __restrict__ v, float step_size, float inv_sqrt_bc2,
                                  float b1, float b2, float eps, float wd) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float gi = 0.f;
  for (int s = 0; s < splits; ++s) gi += g[(size_t)s * gstride + i];
  const float pi = p[i];
  if (wd != 0.f) gi = fmaf(wd, pi, gi);
  const float mi = b1 * m[i] + (1.f - b1) * gi;
  const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
  m[i] = mi; v[i] = vi;
  p[i] = pi - step_size * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
}

}  // namespace icem
