// libicem_b200: device-resident iCEM planner + C ABI (include/icem_b200.h).
//
// One planner object == one `MpcICem` instance (reference: icem/controllers/icem.py:16-247).  The whole plan
// step (all CEM iterations: fused sample->rollout->cost kernel, select+refit kernel, optional NCCL
// all-gather) is enqueued on one stream without host synchronisation and replayed as a CUDA graph;
// only the start state goes in and action[d] (+ best cost) comes out per step.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <vector>

#include "../../include/icem_b200.h"
#include "comm.cuh"
#include "common.cuh"
#include "dyn_articulated.cuh"
#include "dyn_dense.cuh"
#include "mlp_rollout.cuh"
#include "sampler.cuh"
#include "rollout.cuh"
#include "rollout_chain.cuh"
#include "select_refit.cuh"
#include "mlp_train.cuh"

namespace icem {

static thread_local std::string g_last_error;
static std::atomic<uint64_t> g_launches{0};

struct InvalidArg : std::runtime_error { using std::runtime_error::runtime_error; };
struct StateError : std::runtime_error { using std::runtime_error::runtime_error; };
struct Unsupported : std::runtime_error { using std::runtime_error::runtime_error; };

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  void alloc(size_t count) {
    release();
    n = count;
    if (count) {
      ICEM_CUDA(cudaMalloc(&p, count * sizeof(T)));
      // A device memset is asynchronous to the host and runs on the legacy default stream, which does NOT order
      // against the planner's non-blocking stream: wait for it, so that work enqueued on any stream afterwards
      // sees the zeros and never races them.
      ICEM_CUDA(cudaMemset(p, 0, count * sizeof(T)));
      ICEM_CUDA(cudaStreamSynchronize(cudaStreamLegacy));
    }
  }
  // grow-only scratch: reuse across calls (no cudaMalloc / cudaFree, hence no device synchronisation, per call)
  void reserve(size_t count) {
    if (count > n) alloc(count);
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr; n = 0;
  }
  ~DevBuf() { release(); }
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
};

// sampler-only launches of the fused kernel (ICEM_DYN_MLP samples with it, then rolls out on the tensor cores)
struct NoDyn {
  static constexpr int kWarpsPerCta = 8;
  static constexpr int kMinCtasPerSm = 2;
  static constexpr bool kCtaLockstep = false;
  static constexpr bool kHasHealth = false;
  static constexpr bool kOutlineRollout = false;
  struct Params { int act_dim; };
  __host__ __device__ static int cta_floats(const Params&) { return 0; }
  __host__ __device__ static int warp_floats(const Params&) { return 0; }
  __device__ static void cta_init(const Params&, float*) {}
  __device__ void bind(const Params&, const float*, float*) {}
  __device__ void reset(const float*) {}
  __device__ float obs(int) const { return 0.f; }
  __device__ bool state_healthy(int, float) const { return true; }
  __device__ void step(const float*) {}
  __device__ void export_state(float*) const {}
};

constexpr int kStateSlot = 256;     // floats reserved per problem for the start state in `step_in`

struct IterPlan {
  int n_global;       // N_i
  int chunk;          // ceil(N_i / world)
  int n_fresh_local;  // rows this rank samples
  int global_offset;  // global index of local row 0
  int n_shift_local;  // shifted-elite rows this rank simulates at iteration 0 when elites exist
  int rows_cap;       // n_fresh_local + n_shift_local
  size_t cost_off;    // offset into costs buffer
  size_t act_off;     // offset (rows) into actions buffer
};

}  // namespace icem

using namespace icem;

struct icem_planner {
  icem_config_t cfg{};
  std::vector<float> low, high;
  int h = 0, d = 0, hd = 0, K = 0, k = 0, n_keep = 0, stride = 0, iters = 0;
  int state_dim = 0, obs_dim = 0;
  int B = 1;                          // independent MPC problems batched in this handle (cfg.num_problems)
  int active = 0;                     // problem the observable-state getters read
  size_t actions_per = 0, costs_per = 0;   // per-problem elements of `actions` / `costs`
  int sm_count = 148;
  bool white = false;
  std::vector<float> G_host;         // host copy of G: the compile-time-shaped samplers take their table rows as kernel parameters
  bool force_warp_sampler = false;   // ICEM_B200_WARP_SAMPLER=1 at icem_create: A/B the two samplers (tests)
  std::vector<IterPlan> plan;
  cudaStream_t stream = nullptr;

  // device state
  DevBuf<float> actions, costs, G, d_low, d_high, mean, stdv, init_mean, init_std, reset_std;
  DevBuf<float> elite_actions[2], elite_costs[2];
  DevBuf<int32_t> elite_idx[2];
  DevBuf<float> trace_mean, trace_std, trace_costs;
  DevBuf<int32_t> trace_idx;
  DevBuf<float> out;              // [d] action, [1] best cost
  DevBuf<unsigned char> step_in;  // StepState + start state (fp32)
  DevBuf<unsigned long long> cand;
  DevBuf<unsigned int> ticket;
  DevBuf<float> w_obs, w_act, bias;      // dense model
  DevBuf<unsigned char> flush;           // L2 flush buffer (bench)
  DevBuf<float> step_scratch;            // icem_sim_step / _batch / icem_op_rollout_observations staging (grow-only)
  std::vector<DevBuf<float>> inj_zr, inj_zi;
  std::vector<int> inj_rows;
  // multi-rank
  Comm comm;
  DevBuf<unsigned char> send_rec, recv_rec;
  size_t rec_bytes = 0;

  // pinned host staging
  unsigned char* h_in = nullptr;
  float* h_out = nullptr;
  size_t in_bytes = 0;

  // model params
  DenseTanh::Params dense{};
  bool model_ready = false;
  Articulated<32>::Params art{};      // Params is layout-identical for every NVMAX instantiation
  DevBuf<ArtModel> art_model;
  // branch-parallel engine (dyn_chain.cuh): used whenever the robot decomposes into trunk + limb chains
  bool chain_ok = false;
  bool force_warp_engine = false;     // ICEM_B200_ENGINE=warp at icem_create: A/B the two engines (tests, profiles)
  ChainModel chain_host{};
  DevBuf<ChainModel> chain_model;
  ChainParams chain{};
  MlpParams mlp{};
  DevBuf<mlp_op_t> mlp_w1, mlp_w2, mlp_w3;
  DevBuf<float> mlp_bias;

  // run state
  bool was_reset = false;
  uint32_t step = 0;
  uint32_t plans_total = 0;           // never reset (MpcRandom's call counter runs across rollouts)
  bool has_prev = false;
  bool inject_pending = false;
  bool plan_in_flight = false;        // between icem_plan_async and icem_plan_finish
  cudaGraphExec_t graph_exec = nullptr;
  cudaEvent_t ev_a = nullptr, ev_b = nullptr;
  std::vector<cudaEvent_t> ev_roll;   // pairs around rollout kernels (bench / timing mode)
  float last_total_ms = 0.f, last_rollout_ms = 0.f;

  ~icem_planner() {
    if (graph_exec) cudaGraphExecDestroy(graph_exec);
    for (auto e : ev_roll) cudaEventDestroy(e);
    if (ev_a) cudaEventDestroy(ev_a);
    if (ev_b) cudaEventDestroy(ev_b);
    if (h_in) cudaFreeHost(h_in);
    if (h_out) cudaFreeHost(h_out);
    comm.destroy();
    if (stream) cudaStreamDestroy(stream);
  }
};

namespace icem {

// --------------------------------------------------------------------------------------------------
// Synthesis matrix of the power-law colored-noise sampler (colorednoise.powerlaw_psd_gaussian, call
// site icem/controllers/icem.py:73-75; algorithm restated in oracle/shims/colorednoise.py):
//   y[t] = sum_j G[t][j] z[j],  z = [zr(K), zi(K)] unit normals,
//   y = irfft(s*(zr + i zi), n=h) / sigma,  s_k = f_k^(-beta/2) (DC takes bin 1's scale),
//   sigma = 2 sqrt(sum_{k>=1} w_k^2) / h  (w = s, Nyquist halved for even h).
static std::vector<float> build_synthesis_matrix(int h, double beta, bool v2) {
  const int K = h / 2 + 1;
  std::vector<double> s(K);
  for (int k = 0; k < K; ++k) s[k] = (double)k / h;
  const double fmin = 1.0 / h;
  int ix = 0;
  for (int k = 0; k < K; ++k) ix += s[k] < fmin;
  if (ix && ix < K)
    for (int k = 0; k < ix; ++k) s[k] = s[ix];
  for (int k = 0; k < K; ++k) s[k] = std::pow(s[k], -beta / 2.0);
  double acc = 0;
  for (int k = 1; k < K; ++k) {
    double w = s[k];
    if (k == K - 1) w *= (1 + (h % 2)) / 2.0;
    acc += w * w;
  }
  const double sigma = 2.0 * std::sqrt(acc) / h;
  const bool even = (h % 2) == 0;
  std::vector<float> G((size_t)h * 2 * K, 0.f);
  const double two_pi = 6.283185307179586476925286766559;
  for (int t = 0; t < h; ++t) {
    for (int k = 0; k < K; ++k) {
      double gr, gi;
      const bool dc = k == 0, nyq = even && k == K - 1;
      if (dc) { gr = s[0] / h; gi = 0; }
      else if (nyq) { gr = ((t & 1) ? -1.0 : 1.0) * s[k] / h; gi = 0; }
      else {
        const double ang = two_pi * (double)((long long)k * t % h) / h;
        gr = 2.0 * std::cos(ang) * s[k] / h;
        gi = -2.0 * std::sin(ang) * s[k] / h;
      }
      if (v2 && (dc || nyq)) gr *= std::sqrt(2.0);
      G[(size_t)t * 2 * K + k] = (float)(gr / sigma);
      G[(size_t)t * 2 * K + K + k] = (float)(gi / sigma);
    }
  }
  return G;
}

static void upload(DevBuf<float>& b, const std::vector<float>& v) {
  b.alloc(v.size());
  ICEM_CUDA(cudaMemcpy(b.p, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice));
}

// --------------------------------------------------------------------------------------------------
static SamplerConst sampler_const(icem_planner* p) {
  SamplerConst sc{};
  sc.h = p->h; sc.d = p->d; sc.K = p->K; sc.white = p->white;
  sc.trunc = p->cfg.planner == ICEM_PLANNER_CEM_STD;
  sc.levine = p->cfg.bounds_like_levine;
  sc.rnd_freq = p->cfg.planner == ICEM_PLANNER_RANDOM ? p->cfg.action_change_frequency : -1;
  sc.n_global = p->cfg.num_simulated_trajectories;
  sc.magic_d = (uint32_t)((0x100000000ull + p->d - 1) / p->d);
  sc.magic_K = (uint32_t)((0x100000000ull + p->K - 1) / p->K);
  sc.G = p->G.p; sc.low = p->d_low.p; sc.high = p->d_high.p;
  return sc;
}

static CostConst cost_const(icem_planner* p) {
  CostConst cc{};
  cc.kind = p->cfg.cost;
  cc.reduce = p->cfg.cost_along_trajectory;
  cc.penalise_flipping = p->cfg.cost_penalise_flipping;
  if (p->cfg.cost == ICEM_COST_HALFCHEETAH) {
    // environments/mujoco.py:77-84: (root angle, x velocity) = obs[2], obs[9] (18-dim) or obs[1], obs[8] (17-dim)
    cc.idx_a = p->obs_dim == 18 ? 2 : 1;
    cc.idx_b = p->obs_dim == 18 ? 9 : 8;
  } else if (p->cfg.cost == ICEM_COST_LOCOMOTION) {
    // environments/mujoco.py:153-176 (Ant), 196-231 (Hopper)
    cc.idx_a = p->cfg.cost_z_index;
    cc.idx_b = 2;                        // Hopper bounds states[..., 2:] (mujoco.py:199)
    const double w_fwd = p->cfg.cost_forward_weight != 0.0 ? p->cfg.cost_forward_weight : 1.0;
    cc.vel_index = p->cfg.cost_velocity_index1 - 1;
    cc.w_fwd = (float)w_fwd;
    // finite-difference velocity: the post-step term is -(x' - x) * w_fwd / dt; with an observed velocity it vanishes
    cc.inv_dt = cc.vel_index >= 0 ? 0.f : (float)(w_fwd / p->cfg.cost_dt);
    cc.w_ctrl = (float)p->cfg.cost_ctrl_weight;
    cc.w_unhealthy = (float)p->cfg.cost_unhealthy_weight;
    cc.z_lo = (float)p->cfg.cost_z_lo;
    cc.z_hi = (float)p->cfg.cost_z_hi;
    cc.state_bound = (float)p->cfg.cost_state_bound;
    cc.z_strict = p->cfg.cost_z_strict;
  } else if (p->cfg.cost == ICEM_COST_REACHER) {
    for (int i = 0; i < 4; ++i) cc.reach[i] = (float)p->cfg.cost_reach[i];
  } else if (p->cfg.cost == ICEM_COST_GOAL_DISTANCE) {
    cc.goal_idx = p->cfg.cost_goal_index; cc.ach_idx = p->cfg.cost_achieved_index;
    cc.goal_sparse = p->cfg.cost_goal_sparse; cc.goal_shaped = p->cfg.cost_goal_shaped;
    cc.goal_threshold = (float)p->cfg.cost_goal_threshold;
  } else {
    cc.idx_a = 2;   // environments/mujoco.py:267 root z
    cc.idx_b = 0;
  }
  return cc;
}

static StepState* step_state_dev(icem_planner* p) { return reinterpret_cast<StepState*>(p->step_in.p); }
static float* start_state_dev(icem_planner* p) { return reinterpret_cast<float*>(p->step_in.p + sizeof(StepState)); }

template <int NVMAX>
static typename Articulated<NVMAX>::Params art_params(icem_planner* p) {
  typename Articulated<NVMAX>::Params q{};
  q.model = p->art.model; q.act_dim = p->art.act_dim; q.nq = p->art.nq; q.nv = p->art.nv;
  q.nb = p->art.nb; q.nc = p->art.nc;
  return q;
}

// function attributes are per device: remember per (kernel, device) what was configured
template <class K>
static void ensure_dynamic_smem(K kern, size_t smem, int device) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, size_t> configured;
  if (smem <= 48 * 1024) return;
  std::lock_guard<std::mutex> lock(mu);
  size_t& have = configured[{reinterpret_cast<const void*>(kern), device}];
  if (smem > have) {
    ICEM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    have = smem;
  }
}

template <class Dyn, bool kSample, bool kRollout, bool kNextObs = false>
static void launch_rollout_impl(icem_planner* p, const RolloutArgs& a, const typename Dyn::Params& dp, int rows_max) {
  const SamplerConst sc = sampler_const(p);
  const CostConst cc = cost_const(p);
  const int warps = Dyn::kWarpsPerCta;
  const size_t smem = rollout_smem_bytes<Dyn, kSample>(sc, dp, a.stride, warps);
  auto kern = rollout_kernel<Dyn, kSample, kRollout, kNextObs>;
  ensure_dynamic_smem(kern, smem, p->cfg.device);
  int occ = 0;
  ICEM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, warps * 32, smem));
  if (occ < 1) throw InvalidArg("rollout kernel does not fit on an SM (shared memory)");
  const int ctas_needed = (rows_max + warps - 1) / warps;
  // persistent: <= one resident wave, shared between the problems of a batched handle (grid.y = problem)
  const int nprob = a.prob_actions ? p->B : 1;      // operator launches carry no problem strides
  const int grid = std::max(1, std::min(ctas_needed, std::max(1, p->sm_count * occ / nprob)));
  kern<<<dim3(grid, nprob), warps * 32, smem, p->stream>>>(a, sc, cc, dp);
  ICEM_CUDA(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);
}

template <class Dyn, bool kSample, bool kRollout>
static void launch_rollout(icem_planner* p, const RolloutArgs& a, const typename Dyn::Params& dp, int rows_max) {
  if (p->cfg.cost == ICEM_COST_LOCOMOTION) {
    // next_obs costs exist for the articulated ground-truth models only, and only where a cost is evaluated
    if constexpr (kRollout && Dyn::kHasHealth) launch_rollout_impl<Dyn, kSample, true, true>(p, a, dp, rows_max);
    else if constexpr (!kRollout) launch_rollout_impl<Dyn, kSample, false, false>(p, a, dp, rows_max);
    else throw Unsupported("the locomotion cost needs an articulated ground-truth model");
  } else {
    launch_rollout_impl<Dyn, kSample, kRollout, false>(p, a, dp, rows_max);
  }
}

static void launch_mlp(icem_planner* p, const RolloutArgs& a, int rows_max) {
  const SamplerConst sc = sampler_const(p);
  const CostConst cc = cost_const(p);
  const size_t smem = mlp_smem_bytes(p->mlp.hidden);
  auto kern = p->cfg.cost == ICEM_COST_GOAL_DISTANCE ? mlp_rollout_kernel<true> : mlp_rollout_kernel<false>;
  ensure_dynamic_smem(kern, smem, p->cfg.device);
  const int tiles = (rows_max + kMlpTile - 1) / kMlpTile;
  const int grid = std::max(1, std::min(tiles, p->sm_count));     // one resident CTA per SM (weights fill smem)
  kern<<<grid, kMlpThreads, smem, p->stream>>>(a, sc, cc, p->mlp);
  ICEM_CUDA(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);
}

// stand-alone colored-noise sampler (sampler.cuh): even horizons up to 62, iCEM planner
static bool series_sampler_eligible(icem_planner* p) {
  return !p->white && p->cfg.planner == ICEM_PLANNER_ICEM && (p->h % 2) == 0 && p->K <= 32 &&
         p->d <= kSamplerThreads && !p->force_warp_sampler;
}

template <int KPAD, int H, int D>
static void launch_series_sampler_k(icem_planner* p, const RolloutArgs& a, int rows_max) {
  const SamplerConst sc = sampler_const(p);
  constexpr int threads = sampler_threads(KPAD, H, D);
  const int R = sampler_rows_per_batch(p->d, a.stride, threads);
  const size_t smem = sampler_smem_bytes(p->h, p->d, KPAD, a.stride, R);
  auto kern = colored_sampler_kernel<KPAD, H, D>;
  ensure_dynamic_smem(kern, smem, p->cfg.device);
  int occ = 0;
  ICEM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
  if (occ < 1) throw InvalidArg("sampler kernel does not fit on an SM (shared memory)");
  const int batches = (rows_max + R - 1) / R;
  const int grid = std::max(1, std::min(batches, p->sm_count * occ));
  SamplerTable<sampler_table_floats(KPAD, H, D)> tab{};
  if constexpr (H > 0 && D > 0) {      // rows t = 0 .. h/4 of G: cosine side [t][KPAD], then sine side, zero padded
    constexpr int rows = (H / 2) / 2 + 1;
    const int K = p->K;
    float* v = reinterpret_cast<float*>(tab.v2);
    for (int t = 0; t < rows; ++t)
      for (int k = 0; k < K; ++k) {
        v[t * KPAD + k] = p->G_host[(size_t)t * 2 * K + k];
        v[(rows + t) * KPAD + k] = p->G_host[(size_t)t * 2 * K + K + k];
      }
  }
  kern<<<grid, threads, smem, p->stream>>>(a, sc, R, tab);
  ICEM_CUDA(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);
}

static void launch_series_sampler(icem_planner* p, const RolloutArgs& a, int rows_max) {
  // the shapes of BASELINE's configurations are compiled with horizon and action dim as constants
  const bool packed = a.stride == ((p->h * p->d + 3) & ~3);
  if (packed && p->h == 30 && p->d == 17) launch_series_sampler_k<16, 30, 17>(p, a, rows_max);
  else if (packed && p->h == 30 && p->d == 6) launch_series_sampler_k<16, 30, 6>(p, a, rows_max);
  else if (packed && p->h == 12 && p->d == 6) launch_series_sampler_k<8, 12, 6>(p, a, rows_max);
  else if (p->K <= 8) launch_series_sampler_k<8, 0, 0>(p, a, rows_max);
  else if (p->K <= 16) launch_series_sampler_k<16, 0, 0>(p, a, rows_max);
  else launch_series_sampler_k<32, 0, 0>(p, a, rows_max);
}

constexpr size_t kMaxSmemPerCta = 227 * 1024;

template <int G, bool kSample, bool kRollout, bool kNextObs, bool kPlanar>
static void launch_chain_impl(icem_planner* p, const RolloutArgs& a, int rows_max) {
  const SamplerConst sc = sampler_const(p);
  const CostConst cc = cost_const(p);
  ChainParams dp = p->chain;
  const int nprob = a.prob_actions ? p->B : 1;      // operator launches carry no problem strides
  const int trips = (rows_max + dp.rows_per_warp - 1) / dp.rows_per_warp;
  // warps per CTA: as many as shared memory holds (one CTA per SM), but no more than it takes to give every SM work
  int wmax = kChainMaxWarps;
  while (wmax > 1 && chain_rollout_smem_bytes<kSample>(sc, dp, a.stride, wmax) > kMaxSmemPerCta) --wmax;
  const int sms = std::max(1, p->sm_count / nprob);
  // Every warp makes `rounds` trips of equal length (all trajectories cost the same), so the launch takes
  // rounds x (time of one trip) whatever the fill of the last round: spread the trips evenly over the rounds and
  // keep only the warps that needs -- fewer resident warps contend less for the issue slots.
  const int per_sm = (trips + sms - 1) / sms;
  const int rounds = std::max(1, (per_sm + wmax - 1) / wmax);
  int warps = std::max(1, std::min(wmax, (per_sm + rounds - 1) / rounds));
  { const char* e = getenv("ICEM_B200_CHAIN_WARPS"); if (e && atoi(e) > 0) warps = std::min(wmax, atoi(e)); }
  const size_t smem = chain_rollout_smem_bytes<kSample>(sc, dp, a.stride, warps);
  if (smem > kMaxSmemPerCta) throw InvalidArg("chain rollout kernel does not fit on an SM (shared memory)");
  auto kern = chain_rollout_kernel<G, kSample, kRollout, kNextObs, kPlanar>;
  ensure_dynamic_smem(kern, chain_rollout_smem_bytes<kSample>(sc, dp, a.stride, wmax), p->cfg.device);
  const int grid = std::max(1, std::min((trips + warps - 1) / warps, sms));
  kern<<<dim3(grid, nprob), warps * 32, smem, p->stream>>>(a, sc, cc, dp);
  ICEM_CUDA(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);
}

template <int G, bool kSample, bool kRollout, bool kPlanar>
static void launch_chain_g(icem_planner* p, const RolloutArgs& a, int rows_max) {
  if (p->cfg.cost == ICEM_COST_LOCOMOTION) {
    if constexpr (kRollout) launch_chain_impl<G, kSample, true, true, kPlanar>(p, a, rows_max);
    else launch_chain_impl<G, kSample, false, false, kPlanar>(p, a, rows_max);
  } else {
    launch_chain_impl<G, kSample, kRollout, false, kPlanar>(p, a, rows_max);
  }
}

template <bool kSample, bool kRollout>
static void launch_chain(icem_planner* p, const RolloutArgs& a, int rows_max) {
  // planar robots (HalfCheetah: 2 lanes, Hopper: 1) run the planar instantiation of the engine
  if (p->chain_host.planar && p->chain_host.lanes <= 2) {
    if (p->chain_host.lanes == 1) launch_chain_g<1, kSample, kRollout, true>(p, a, rows_max);
    else launch_chain_g<2, kSample, kRollout, true>(p, a, rows_max);
    return;
  }
  switch (p->chain_host.lanes) {
    case 1: launch_chain_g<1, kSample, kRollout, false>(p, a, rows_max); break;
    case 2: launch_chain_g<2, kSample, kRollout, false>(p, a, rows_max); break;
    default: launch_chain_g<4, kSample, kRollout, false>(p, a, rows_max); break;
  }
}

template <bool kSample, bool kRollout>
static void launch_rollout_dyn(icem_planner* p, const RolloutArgs& a, int rows_max) {
  if (kSample && !kRollout && series_sampler_eligible(p)) {
    launch_series_sampler(p, a, rows_max);
    return;
  }
  switch (p->cfg.dynamics) {
    case ICEM_DYN_MLP: {
      if (kSample) {
        if (series_sampler_eligible(p)) {
          launch_series_sampler(p, a, rows_max);
        } else {
          NoDyn::Params np{p->d};
          launch_rollout<NoDyn, true, false>(p, a, np, rows_max);
        }
      }
      if (kRollout) launch_mlp(p, a, rows_max);
      break;
    }
    case ICEM_DYN_DENSE_TANH:
      launch_rollout<DenseTanh, kSample, kRollout>(p, a, p->dense, rows_max);
      break;
    case ICEM_DYN_HALFCHEETAH:
    case ICEM_DYN_HUMANOID_STANDUP:
    case ICEM_DYN_ARTICULATED:
      if (p->chain_ok)
        launch_chain<kSample, kRollout>(p, a, rows_max);
      else if (Articulated<12>::fits(p->art.nb, p->art.nv, p->art.nc))
        launch_rollout<Articulated<12>, kSample, kRollout>(p, a, art_params<12>(p), rows_max);
      else if (Articulated<24>::fits(p->art.nb, p->art.nv, p->art.nc))
        launch_rollout<Articulated<24>, kSample, kRollout>(p, a, art_params<24>(p), rows_max);
      else
        launch_rollout<Articulated<32>, kSample, kRollout>(p, a, art_params<32>(p), rows_max);
      break;
    default:
      throw Unsupported("dynamics id not supported by this build");
  }
}

static RolloutArgs rollout_args(icem_planner* p, int i) {
  const IterPlan& ip = p->plan[i];
  RolloutArgs a{};
  a.n_fresh_local = ip.n_fresh_local;
  a.n_shift_local = ip.n_shift_local;
  a.global_offset = ip.global_offset;
  a.n_fresh_global = ip.n_global;
  a.iteration = i;
  a.inject_mean_row0 = p->cfg.use_mean_actions && i == p->iters - 1;
  a.stride = p->stride;
  a.actions = p->actions.p + ip.act_off * p->stride;
  a.costs = p->costs.p + ip.cost_off;
  a.mean = p->mean.p;
  a.std = p->stdv.p;
  a.prev_elites = p->elite_actions[(p->iters - 1) & 1].p;
  a.start_state = start_state_dev(p);
  a.inj_zr = (size_t)i < p->inj_zr.size() ? p->inj_zr[i].p : nullptr;
  a.inj_zi = (size_t)i < p->inj_zi.size() ? p->inj_zi[i].p : nullptr;
  a.ss = step_state_dev(p);
  a.seed_lo = (uint32_t)p->cfg.seed;
  a.seed_hi = (uint32_t)(p->cfg.seed >> 32);
  a.prob_actions = p->actions_per; a.prob_costs = p->costs_per;
  a.prob_dist = p->hd; a.prob_elites = p->k * p->stride; a.prob_state = kStateSlot;
  return a;
}

static RefitArgs refit_args(icem_planner* p, int i) {
  const IterPlan& ip = p->plan[i];
  RefitArgs r{};
  r.h = p->h; r.d = p->d; r.k = p->k; r.stride = p->stride; r.iteration = i;
  r.last_iteration = i == p->iters - 1;
  r.world = p->cfg.world_size;
  r.n_keep = (i > 0 && p->cfg.keep_previous_elites) ? p->n_keep : 0;
  r.n_fresh_global = ip.n_global;
  r.alpha = (float)p->cfg.alpha;
  r.one_minus_alpha = (float)(1.0 - p->cfg.alpha);
  if (p->cfg.world_size > 1) {
    r.rec_keys = reinterpret_cast<const unsigned long long*>(p->recv_rec.p);
    r.rec_rank_stride_bytes = p->rec_bytes;
    r.local_actions = nullptr;
  } else {
    r.rec_keys = nullptr;   // set inside select_kernel
    r.rec_rank_stride_bytes = 0;
    r.local_actions = p->actions.p + ip.act_off * p->stride;
  }
  r.local_n_fresh = ip.n_fresh_local;
  r.local_offset = ip.global_offset;
  const int nb = i & 1, pb = (i > 0 ? i - 1 : p->iters - 1) & 1;
  r.prev_elite_actions = p->elite_actions[pb].p;
  r.prev_elite_costs = p->elite_costs[pb].p;
  r.new_elite_actions = p->elite_actions[nb].p;
  r.new_elite_costs = p->elite_costs[nb].p;
  r.new_elite_idx = p->elite_idx[nb].p;
  r.mean = p->mean.p; r.std = p->stdv.p; r.init_std = p->init_std.p;
  r.trace_mean = p->trace_mean.p + (size_t)i * p->hd;
  r.trace_std = p->trace_std.p + (size_t)i * p->hd;
  r.trace_costs = p->trace_costs.p + (size_t)i * p->k;
  r.trace_idx = p->trace_idx.p + (size_t)i * p->k;
  r.cem_std = p->cfg.planner == ICEM_PLANNER_CEM_STD;
  r.execute_mean = r.cem_std && !p->cfg.execute_best_elite;
  r.mean_to_zero = r.cem_std && !p->cfg.shift_means;
  r.levine = r.cem_std && p->cfg.bounds_like_levine;
  r.low = p->d_low.p; r.high = p->d_high.p;
  r.out_action = p->out.p;
  r.out_best_cost = p->out.p + p->d;
  r.prob_actions = p->actions_per; r.prob_dist = p->hd; r.prob_elites = p->k * p->stride; r.prob_k = p->k;
  r.prob_trace_hd = p->iters * p->hd; r.prob_trace_k = p->iters * p->k; r.prob_out = p->d + 1;
  return r;
}

static int select_grid(icem_planner* p, int rows) {
  return std::max(1, std::min(p->sm_count, (rows + 2047) / 2048));
}
static size_t select_smem(int k) { return (size_t)k * sizeof(unsigned long long); }

// enqueue all CEM iterations of one plan step on the planner's stream
static void enqueue_iterations(icem_planner* p, bool time_rollouts) {
  for (int i = 0; i < p->iters; ++i) {
    const IterPlan& ip = p->plan[i];
    RolloutArgs a = rollout_args(p, i);
    if (time_rollouts) ICEM_CUDA(cudaEventRecord(p->ev_roll[2 * i], p->stream));
    launch_rollout_dyn<true, true>(p, a, ip.rows_cap);
    if (time_rollouts) ICEM_CUDA(cudaEventRecord(p->ev_roll[2 * i + 1], p->stream));

    SelectArgs s{};
    s.n_fresh_local = ip.n_fresh_local; s.n_shift_local = ip.n_shift_local;
    s.global_offset = ip.global_offset; s.n_fresh_global = ip.n_global; s.iteration = i;
    s.k = p->k; s.stride = p->stride;
    s.costs = a.costs; s.actions = a.actions; s.ss = a.ss;
    s.cand = p->cand.p; s.ticket = p->ticket.p;
    s.prob_actions = p->actions_per; s.prob_costs = p->costs_per; s.prob_cand = p->sm_count * p->k;
    RefitArgs r = refit_args(p, i);
    const size_t smem = select_smem(p->k);
    if (p->cfg.world_size == 1) {
      select_kernel<<<dim3(select_grid(p, ip.rows_cap), p->B), kSelectThreads, smem, p->stream>>>(s, r, 1);
      ICEM_CUDA(cudaGetLastError());
      g_launches.fetch_add(1, std::memory_order_relaxed);
    } else {
      s.send_keys = reinterpret_cast<unsigned long long*>(p->send_rec.p);
      s.send_actions = reinterpret_cast<float*>(p->send_rec.p + (size_t)p->k * sizeof(unsigned long long));
      select_kernel<<<select_grid(p, ip.rows_cap), kSelectThreads, smem, p->stream>>>(s, r, 0);
      ICEM_CUDA(cudaGetLastError());
      p->comm.all_gather(p->send_rec.p, p->recv_rec.p, p->rec_bytes, p->stream);
      merge_refit_kernel<<<1, kSelectThreads, smem, p->stream>>>(r);
      ICEM_CUDA(cudaGetLastError());
      g_launches.fetch_add(2, std::memory_order_relaxed);
    }
  }
}

// state_dev <- f(state_dev, executed action): closed loop on the device
// blockIdx.x = instance (icem_sim_step_batch): element strides between instances, 0 for a single transition
struct AdvanceBatch { int state_stride, action_stride, obs_stride; };

template <class Dyn>
__global__ void advance_kernel(typename Dyn::Params dp, float* state, const float* action, float* next_state,
                               float* obs_out, int obs_dim, AdvanceBatch ab) {
  extern __shared__ __align__(128) float smem[];
  state += (size_t)blockIdx.x * ab.state_stride;
  if (action) action += (size_t)blockIdx.x * ab.action_stride;
  if (next_state) next_state += (size_t)blockIdx.x * ab.state_stride;
  if (obs_out) obs_out += (size_t)blockIdx.x * ab.obs_stride;
  float* s_dyn = smem;
  float* w_dyn = s_dyn + ((Dyn::cta_floats(dp) + 3) & ~3);
  float* s_act = w_dyn + ((Dyn::warp_floats(dp) + 3) & ~3);
  Dyn::cta_init(dp, s_dyn);
  if (action)
    for (int i = threadIdx.x; i < dp.act_dim; i += blockDim.x) s_act[i] = action[i];
  __syncthreads();
  Dyn dyn;
  dyn.bind(dp, s_dyn, w_dyn);
  dyn.reset(state);
  if (action) dyn.step(s_act);
  __syncwarp();
  if (next_state) dyn.export_state(next_state);
  if (obs_out)
    for (int i = threadIdx.x; i < obs_dim; i += 32) obs_out[i] = dyn.obs(i);
}

template <class Dyn>
static void launch_advance(icem_planner* p, const typename Dyn::Params& dp, float* state, const float* action,
                           float* next_state, float* obs_out, int obs_dim, int batch = 1, AdvanceBatch ab = {0, 0, 0}) {
  const size_t smem = (((Dyn::cta_floats(dp) + 3) & ~3) + ((Dyn::warp_floats(dp) + 3) & ~3) + dp.act_dim + 4) *
                      sizeof(float);
  auto kern = advance_kernel<Dyn>;
  if (smem > 48 * 1024)
    ICEM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<batch, 32, smem, p->stream>>>(dp, state, action, next_state, obs_out, obs_dim, ab);
  ICEM_CUDA(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);
}

static void advance_dyn(icem_planner* p, float* state, const float* action, float* next_state, float* obs_out,
                        int obs_dim, int batch = 1, AdvanceBatch ab = {0, 0, 0}) {
  switch (p->cfg.dynamics) {
    case ICEM_DYN_MLP:
      if (batch != 1) throw Unsupported("batched transitions are not available for the MLP model");
      mlp_advance_kernel<<<1, 256, 0, p->stream>>>(p->mlp, state, action, next_state, obs_out, obs_dim);
      ICEM_CUDA(cudaGetLastError());
      g_launches.fetch_add(1, std::memory_order_relaxed);
      break;
    case ICEM_DYN_DENSE_TANH:
      launch_advance<DenseTanh>(p, p->dense, state, action, next_state, obs_out, obs_dim, batch, ab);
      break;
    case ICEM_DYN_HALFCHEETAH:
    case ICEM_DYN_HUMANOID_STANDUP:
    case ICEM_DYN_ARTICULATED:
      if (p->chain_ok) {
        const size_t smem = ((((sizeof(ChainModel) / 4) + 3) & ~(size_t)3) + p->chain.warp_floats + 4) * sizeof(float);
        auto launch = [&](auto kern) {
          ensure_dynamic_smem(kern, smem, p->cfg.device);
          kern<<<batch, 32, smem, p->stream>>>(p->chain, state, action, next_state, obs_out, obs_dim, ab.state_stride,
                                               ab.action_stride, ab.obs_stride);
        };
        const bool planar = p->chain_host.planar && p->chain_host.lanes <= 2;
        if (planar && p->chain_host.lanes == 1) launch(chain_advance_kernel<1, true>);
        else if (planar) launch(chain_advance_kernel<2, true>);
        else if (p->chain_host.lanes == 1) launch(chain_advance_kernel<1, false>);
        else if (p->chain_host.lanes == 2) launch(chain_advance_kernel<2, false>);
        else launch(chain_advance_kernel<4, false>);
        ICEM_CUDA(cudaGetLastError());
        g_launches.fetch_add(1, std::memory_order_relaxed);
      } else if (Articulated<12>::fits(p->art.nb, p->art.nv, p->art.nc))
        launch_advance<Articulated<12>>(p, art_params<12>(p), state, action, next_state, obs_out, obs_dim, batch, ab);
      else if (Articulated<24>::fits(p->art.nb, p->art.nv, p->art.nc))
        launch_advance<Articulated<24>>(p, art_params<24>(p), state, action, next_state, obs_out, obs_dim, batch, ab);
      else
        launch_advance<Articulated<32>>(p, art_params<32>(p), state, action, next_state, obs_out, obs_dim, batch, ab);
      break;
    default:
      throw Unsupported("dynamics id not supported by this build");
  }
}

static void require_model(icem_planner* p) {
  if (!p->model_ready) throw StateError("forward model parameters not set (icem_set_dense_model / ...)");
}

static void build_plan(icem_planner* p) {
  const icem_config_t& c = p->cfg;
  p->plan.clear();
  int n = c.num_simulated_trajectories;
  size_t cost_off = 0, act_off = 0;
  int rows_max = 0;
  for (int i = 0; i < c.opt_iterations; ++i) {
    if (i > 0) n = std::max(c.elites_size * 2, (int)((double)n / c.factor_decrease_num));   // icem.py:126-127
    IterPlan ip{};
    ip.n_global = n;
    ip.chunk = (n + c.world_size - 1) / c.world_size;
    ip.global_offset = std::min(n, c.rank * ip.chunk);
    ip.n_fresh_local = std::min(n, ip.global_offset + ip.chunk) - ip.global_offset;
    // shifted elites (iteration 0, t>0) are simulated on the LAST rank at global indices N_0..
    ip.n_shift_local = (i == 0 && c.shift_elites_over_time && c.rank == c.world_size - 1) ? p->n_keep : 0;
    ip.rows_cap = ip.n_fresh_local + ip.n_shift_local;
    ip.cost_off = cost_off;
    cost_off += (size_t)ip.rows_cap;
    ip.act_off = c.keep_iteration_actions ? act_off : 0;
    act_off += (size_t)ip.rows_cap;
    rows_max = std::max(rows_max, ip.rows_cap);
    p->plan.push_back(ip);
  }
  p->costs_per = std::max<size_t>(cost_off, 1);
  p->actions_per = (size_t)std::max<size_t>(c.keep_iteration_actions ? act_off : (size_t)rows_max, 1) * p->stride;
  p->costs.alloc(p->costs_per * p->B);
  p->actions.alloc(p->actions_per * p->B);
}

static void reset_distribution(icem_planner* p) {
  for (int b = 0; b < p->B; ++b) {
    ICEM_CUDA(cudaMemcpyAsync(p->mean.p + (size_t)b * p->hd, p->init_mean.p, p->hd * sizeof(float),
                              cudaMemcpyDeviceToDevice, p->stream));
    ICEM_CUDA(cudaMemcpyAsync(p->stdv.p + (size_t)b * p->hd, p->reset_std.p, p->hd * sizeof(float),
                              cudaMemcpyDeviceToDevice, p->stream));
  }
}

static void write_step_in(icem_planner* p, const double* state) {
  StepState ss{};
  ss.step = p->step;
  ss.has_prev_elites = p->has_prev ? 1 : 0;
  ss.inject = p->inject_pending ? 1 : 0;
  ss.plans_total = p->plans_total;
  memcpy(p->h_in, &ss, sizeof ss);
  if (state) {      // [B][state_dim] doubles -> one kStateSlot-float slot per problem
    float* f = reinterpret_cast<float*>(p->h_in + sizeof(StepState));
    for (int b = 0; b < p->B; ++b)
      for (int i = 0; i < p->state_dim; ++i) f[(size_t)b * kStateSlot + i] = (float)state[(size_t)b * p->state_dim + i];
  }
}

static void ensure_graph(icem_planner* p) {
  if (p->graph_exec) return;
  cudaGraph_t g = nullptr;
  ICEM_CUDA(cudaStreamBeginCapture(p->stream, cudaStreamCaptureModeThreadLocal));
  try {
    ICEM_CUDA(cudaMemcpyAsync(p->step_in.p, p->h_in, p->in_bytes, cudaMemcpyHostToDevice, p->stream));
    enqueue_iterations(p, false);
    ICEM_CUDA(cudaMemcpyAsync(p->h_out, p->out.p, (size_t)p->B * (p->d + 1) * sizeof(float), cudaMemcpyDeviceToHost,
                              p->stream));
  } catch (...) {
    cudaStreamEndCapture(p->stream, &g);
    if (g) cudaGraphDestroy(g);
    throw;
  }
  ICEM_CUDA(cudaStreamEndCapture(p->stream, &g));
  cudaError_t e = cudaGraphInstantiate(&p->graph_exec, g, 0);
  cudaGraphDestroy(g);
  ICEM_CUDA(e);
}

static void finish_step(icem_planner* p) {
  p->step += 1;
  p->plans_total += 1;
  p->has_prev = true;
  if (p->inject_pending) {
    p->inject_pending = false;
    p->inj_zr.clear();
    p->inj_zi.clear();
    p->inj_rows.clear();
  }
}

static int rows_of(icem_planner* p, int i, bool first_step) {
  const IterPlan& ip = p->plan[i];
  return ip.n_fresh_local + ((i == 0 && !first_step) ? ip.n_shift_local : 0);
}

}  // namespace icem

// ====================================================================================================
// C ABI
// ====================================================================================================
// ---------------------------------------------------------------------------------------------------------------
// MLP trainer (mlp_train.cuh): host side
namespace icem {

struct MlpTrainer {
  int device = 0, in = 0, H = 0, out = 0;
  cudaStream_t stream = nullptr;
  // parameters: W1 [H][in], b1 [H], W2 [H][H], b2 [H], W3 [out][H], b3 [out]; Adam moments alongside
  DevBuf<float> par[6], mom[6], var[6];
  size_t par_n[6] = {0, 0, 0, 0, 0, 0};
  long long adam_t = 0;
  // data set and minibatch buffers
  DevBuf<float> x, t;
  size_t n_rows = 0;
  DevBuf<int> idx;
  DevBuf<float> xb, tb, a1, a2, y, dy, dz2, dz1, gw[3], gb[3], loss_part, losses;
  int cap_batch = 0;

  ~MlpTrainer() {
    if (stream) cudaStreamDestroy(stream);
  }
};

template <bool TA, bool TB, int EP>
static void train_gemm(MlpTrainer* t, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C,
                       int ldc, const float* bias, const float* aux, int ldaux, int splits = 1) {
  const int kchunk = ((K + splits - 1) / splits + kTrKT - 1) / kTrKT * kTrKT;
  dim3 grid((N + kTrTile - 1) / kTrTile, (M + kTrTile - 1) / kTrTile, splits);
  train_gemm_kernel<TA, TB, EP><<<grid, 256, 0, t->stream>>>(M, N, K, A, lda, B, ldb, C, ldc, bias, aux, ldaux, kchunk);
  ICEM_CUDA(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);
}

static void trainer_reserve_batch(MlpTrainer* t, int batch) {
  if (batch <= t->cap_batch) return;
  const int H = t->H;
  t->xb.alloc((size_t)batch * t->in); t->tb.alloc((size_t)batch * t->out);
  t->a1.alloc((size_t)batch * H); t->a2.alloc((size_t)batch * H);
  t->y.alloc((size_t)batch * t->out); t->dy.alloc((size_t)batch * t->out);
  t->dz2.alloc((size_t)batch * H); t->dz1.alloc((size_t)batch * H);
  for (int l = 0; l < 3; ++l) {
    t->gw[l].alloc(t->par_n[2 * l] * 16);       // up to 16 batch splits
    t->gb[l].alloc(t->par_n[2 * l + 1]);
  }
  t->loss_part.alloc(((size_t)batch * t->out + 255) / 256);
  t->cap_batch = batch;
}

// one Adam step on the minibatch rows idx[0..batch) (device); the step's loss goes to loss_out (device, 1 float)
static void trainer_step(MlpTrainer* t, const int* idx, int batch, float lr, float b1, float b2, float eps, float wd,
                         float* loss_out) {
  const int in = t->in, H = t->H, out = t->out;
  float* W1 = t->par[0].p; float* B1 = t->par[1].p; float* W2 = t->par[2].p; float* B2 = t->par[3].p;
  float* W3 = t->par[4].p; float* B3 = t->par[5].p;
  train_gather_kernel<<<batch, 32, 0, t->stream>>>(batch, in, out, t->x.p, t->t.p, idx, t->xb.p, t->tb.p);
  // forward
  train_gemm<false, true, kTrEpBiasTanh>(t, batch, H, in, t->xb.p, in, W1, in, t->a1.p, H, B1, nullptr, 0);
  train_gemm<false, true, kTrEpBiasTanh>(t, batch, H, H, t->a1.p, H, W2, H, t->a2.p, H, B2, nullptr, 0);
  train_gemm<false, true, kTrEpBias>(t, batch, out, H, t->a2.p, H, W3, H, t->y.p, out, B3, nullptr, 0);
  // loss and its gradient
  const int count = batch * out, nblk = (count + 255) / 256;
  train_loss_grad_kernel<<<nblk, 256, 0, t->stream>>>(count, t->y.p, t->tb.p, t->dy.p, t->loss_part.p);
  train_loss_reduce_kernel<<<1, 256, 0, t->stream>>>(nblk, count, t->loss_part.p, loss_out);
  // backward: data gradients through the tanh layers, weight gradients split over the batch, bias gradients
  const int S = std::max(1, std::min(16, batch / 256));     // a function of the batch alone: results do not depend
                                                            // on what the handle ran before
  train_gemm<false, false, kTrEpTanhGrad>(t, batch, H, out, t->dy.p, out, W3, H, t->dz2.p, H, nullptr, t->a2.p, H);
  train_gemm<false, false, kTrEpTanhGrad>(t, batch, H, H, t->dz2.p, H, W2, H, t->dz1.p, H, nullptr, t->a1.p, H);
  train_gemm<true, false, kTrEpNone>(t, out, H, batch, t->dy.p, out, t->a2.p, H, t->gw[2].p, H, nullptr, nullptr, 0, S);
  train_gemm<true, false, kTrEpNone>(t, H, H, batch, t->dz2.p, H, t->a1.p, H, t->gw[1].p, H, nullptr, nullptr, 0, S);
  train_gemm<true, false, kTrEpNone>(t, H, in, batch, t->dz1.p, H, t->xb.p, in, t->gw[0].p, in, nullptr, nullptr, 0, S);
  train_colsum_kernel<<<(out + 31) / 32, 256, 0, t->stream>>>(batch, out, t->dy.p, out, t->gb[2].p);
  train_colsum_kernel<<<(H + 31) / 32, 256, 0, t->stream>>>(batch, H, t->dz2.p, H, t->gb[1].p);
  train_colsum_kernel<<<(H + 31) / 32, 256, 0, t->stream>>>(batch, H, t->dz1.p, H, t->gb[0].p);
  // Adam (torch.optim.Adam: step_size = lr / (1 - b1^t), denominator sqrt(v) / sqrt(1 - b2^t) + eps)
  t->adam_t += 1;
  const double bc1 = 1.0 - std::pow((double)b1, (double)t->adam_t), bc2 = 1.0 - std::pow((double)b2, (double)t->adam_t);
  const float step_size = (float)(lr / bc1), inv_sqrt_bc2 = (float)(1.0 / std::sqrt(bc2));
  for (int q = 0; q < 6; ++q) {
    const int n = (int)t->par_n[q];
    const bool is_w = (q % 2) == 0;
    const float* g = is_w ? t->gw[q / 2].p : t->gb[q / 2].p;
    train_adam_kernel<<<(n + 255) / 256, 256, 0, t->stream>>>(n, t->par[q].p, g, is_w ? S : 1, t->par_n[q], t->mom[q].p,
                                                              t->var[q].p, step_size, inv_sqrt_bc2, b1, b2, eps, wd);
  }
  ICEM_CUDA(cudaGetLastError());
  g_launches.fetch_add(12, std::memory_order_relaxed);
}

}  // namespace icem

#define ICEM_API_BEGIN try {
#define ICEM_API_END                                                         \
  }                                                                          \
  catch (const icem::InvalidArg& e) { icem::g_last_error = e.what(); return ICEM_ERR_INVALID; }       \
  catch (const icem::StateError& e) { icem::g_last_error = e.what(); return ICEM_ERR_STATE; }         \
  catch (const icem::Unsupported& e) { icem::g_last_error = e.what(); return ICEM_ERR_UNSUPPORTED; }  \
  catch (const icem::CommError& e) { icem::g_last_error = e.what(); return ICEM_ERR_COMM; }           \
  catch (const icem::CudaError& e) { icem::g_last_error = e.what(); return ICEM_ERR_CUDA; }           \
  catch (const std::exception& e) { icem::g_last_error = e.what(); return ICEM_ERR_INVALID; }         \
  return ICEM_OK;

extern "C" {

const char* icem_last_error(void) { return icem::g_last_error.c_str(); }
int icem_abi_version(void) { return ICEM_ABI_VERSION; }
int icem_config_sizeof(void) { return (int)sizeof(icem_config_t); }
int icem_articulated_model_sizeof(void) { return (int)sizeof(icem_articulated_model_t); }
uint64_t icem_kernel_launch_count(void) { return icem::g_launches.load(); }

int icem_create(const icem_config_t* cfg, icem_planner_t** out) {
  ICEM_API_BEGIN
  if (!cfg || !out) throw InvalidArg("null argument");
  if (cfg->abi_version != ICEM_ABI_VERSION) throw InvalidArg("icem_config_t.abi_version mismatch");
  if (cfg->num_simulated_trajectories < 2) throw InvalidArg("At least two trajectories needed!");   // mpc.py:30-31
  if (cfg->horizon < 1 || cfg->act_dim < 1 || cfg->opt_iterations < 1 || cfg->elites_size < 1)
    throw InvalidArg("horizon, act_dim, opt_iterations, elites_size must be positive");
  if (!cfg->action_low || !cfg->action_high) throw InvalidArg("action bounds missing");
  if (cfg->world_size < 1 || cfg->rank < 0 || cfg->rank >= cfg->world_size) throw InvalidArg("bad rank/world_size");
  if (cfg->factor_decrease_num <= 0) throw InvalidArg("factor_decrease_num must be positive");
  if (cfg->cost_along_trajectory < 0 || cfg->cost_along_trajectory > 2)
    throw Unsupported("Implement method to compute cost along trajectory");   // abstract_controller.py:88-91
  int ndev = 0;
  ICEM_CUDA(cudaGetDeviceCount(&ndev));
  if (cfg->device < 0 || cfg->device >= ndev) throw InvalidArg("CUDA device ordinal out of range");
  ICEM_CUDA(cudaSetDevice(cfg->device));
  cudaDeviceProp prop{};
  ICEM_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major < 10) throw Unsupported("libicem_b200 is built for sm_100a (B200) only");

  std::unique_ptr<icem_planner> p(new icem_planner);
  p->cfg = *cfg;
  p->sm_count = prop.multiProcessorCount;
  p->h = cfg->horizon; p->d = cfg->act_dim; p->hd = p->h * p->d; p->K = p->h / 2 + 1;
  p->iters = cfg->opt_iterations;
  if (cfg->planner != ICEM_PLANNER_ICEM && cfg->planner != ICEM_PLANNER_CEM_STD &&
      cfg->planner != ICEM_PLANNER_RANDOM)
    throw Unsupported("unknown planner id");
  const bool cem_std = cfg->planner == ICEM_PLANNER_CEM_STD;
  const bool rnd = cfg->planner == ICEM_PLANNER_RANDOM;
  p->B = cfg->num_problems > 0 ? cfg->num_problems : 1;
  if (p->B > 4096) throw InvalidArg("num_problems must be <= 4096");
  if (cfg->cost == ICEM_COST_LOCOMOTION) {
    if (!(cfg->cost_dt > 0)) throw InvalidArg("cost_dt must be positive for ICEM_COST_LOCOMOTION");
    if (cfg->dynamics == ICEM_DYN_MLP) throw Unsupported("the locomotion cost is not available for the MLP rollout");
  } else if (cfg->cost == ICEM_COST_REACHER) {
    // the distance is formed from the arm's joint angles: only the articulated ground-truth model carries them
    if (cfg->dynamics != ICEM_DYN_ARTICULATED) throw Unsupported("the reacher cost needs ICEM_DYN_ARTICULATED");
  } else if (cfg->cost == ICEM_COST_GOAL_DISTANCE) {
    if (cfg->dynamics != ICEM_DYN_MLP && cfg->dynamics != ICEM_DYN_DENSE_TANH)
      throw Unsupported("the goal-distance cost reads observations of a batched model (ICEM_DYN_MLP / ICEM_DYN_DENSE_TANH)");
    const int od = cfg->obs_dim;
    if (cfg->cost_goal_index < 0 || cfg->cost_goal_index + 3 > od || cfg->cost_achieved_index < 0 ||
        cfg->cost_achieved_index + 3 > od || (cfg->cost_goal_shaped && od < 6))
      throw InvalidArg("goal / achieved-goal indices outside the observation");
    if (!(cfg->cost_goal_threshold >= 0)) throw InvalidArg("cost_goal_threshold must be >= 0");
  } else if (cfg->cost != ICEM_COST_HALFCHEETAH && cfg->cost != ICEM_COST_HUMANOID_STANDUP) {
    throw Unsupported("unknown cost id");
  }
  if (p->B > 1) {
    if (cfg->world_size > 1) throw Unsupported("num_problems > 1 needs world_size == 1 (problems are not sharded)");
    if (cfg->dynamics == ICEM_DYN_MLP) throw Unsupported("num_problems > 1 is not available for the MLP rollout");
  }
  if (rnd) {
    // MpcRandom (controllers/mpc.py:86-138): ONE population per plan step, nothing is refit or reused
    if (cfg->action_change_frequency < 0 || cfg->action_change_frequency >= cfg->horizon)
      throw InvalidArg("action_change_frequency must be in [0, horizon)");       // mpc.py:92
    p->cfg.opt_iterations = 1; p->iters = 1;
    p->cfg.factor_decrease_num = 1.0;
    p->cfg.use_mean_actions = p->cfg.keep_previous_elites = p->cfg.shift_elites_over_time = 0;
    p->cfg.execute_best_elite = 1; p->cfg.shift_means = 1; p->cfg.bounds_like_levine = 0;
  } else if (cem_std) {
    // MpcCemStd has no population decay, elite reuse or mean injection (controllers/mpc.py:212-233)
    p->cfg.factor_decrease_num = 1.0;
    p->cfg.use_mean_actions = p->cfg.keep_previous_elites = p->cfg.shift_elites_over_time = 0;
  } else {
    p->cfg.execute_best_elite = 1; p->cfg.shift_means = 1; p->cfg.bounds_like_levine = 0;
  }
  p->white = cem_std || rnd || !(cfg->noise_beta > 0);
  { const char* e = getenv("ICEM_B200_WARP_SAMPLER"); p->force_warp_sampler = e && e[0] == '1'; }
  { const char* e = getenv("ICEM_B200_ENGINE"); p->force_warp_engine = e && strcmp(e, "warp") == 0; }
  p->low.assign(cfg->action_low, cfg->action_low + p->d);
  p->high.assign(cfg->action_high, cfg->action_high + p->d);
  p->cfg.action_low = p->low.data();
  p->cfg.action_high = p->high.data();
  // icem.py:237-240 (floor of two elites)
  p->k = std::max(2, std::min(cfg->elites_size, cfg->num_simulated_trajectories / 2));
  if (p->k > 64) throw Unsupported("num_elites > 64 not supported");
  // icem.py:99,145: int(len(elites) * fraction_elites_reused) -- same double arithmetic as Python
  p->n_keep = (int)((double)p->k * cfg->fraction_elites_reused);
  if (p->n_keep < 0 || p->n_keep > p->k) throw InvalidArg("fraction_elites_reused out of range");
  p->stride = (p->hd + 3) & ~3;   // 16-B multiple for the 1-D TMA bulk copies
  p->obs_dim = cfg->obs_dim;

  ICEM_CUDA(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
  ICEM_CUDA(cudaEventCreate(&p->ev_a));
  ICEM_CUDA(cudaEventCreate(&p->ev_b));
  p->ev_roll.resize(2 * p->iters);
  for (auto& e : p->ev_roll) ICEM_CUDA(cudaEventCreate(&e));

  p->G_host = p->white ? std::vector<float>(1, 0.f)
                        : build_synthesis_matrix(p->h, cfg->noise_beta, cfg->colorednoise_v2 != 0);
  upload(p->G, p->G_host);
  upload(p->d_low, p->low);
  upload(p->d_high, p->high);
  // icem.py:48-59: the bounds are float32 (gym Box) and the reference does this arithmetic on them
  std::vector<float> m0(p->hd), s0(p->hd);
  for (int t = 0; t < p->h; ++t)
    for (int j = 0; j < p->d; ++j) {
      const float mid = (p->high[j] + p->low[j]) / 2.0f;
      const float half = (p->high[j] - p->low[j]) / 2.0f;
      m0[t * p->d + j] = (float)(0.0 + (double)mid);
      s0[t * p->d + j] = (float)((double)half * cfg->init_std);
    }
  upload(p->init_mean, m0);
  upload(p->init_std, s0);
  // the std a rollout starts from: MpcCemStd with bounds_like_levine clamps it at once (mpc.py:170, 291-294; the mean
  // is the mid point there, so the distance to either bound is half the range)
  std::vector<float> s0r = s0;
  if (cem_std && cfg->bounds_like_levine)
    for (int t = 0; t < p->h; ++t)
      for (int j = 0; j < p->d; ++j) {
        const float half = (p->high[j] - p->low[j]) / 2.0f;
        s0r[t * p->d + j] = std::max(1e-8f, std::min(half * 0.5f, s0[t * p->d + j]));
      }
  upload(p->reset_std, s0r);
  const size_t B = (size_t)p->B;      // every per-problem buffer is [B][...]
  p->mean.alloc(B * p->hd);
  p->stdv.alloc(B * p->hd);
  for (int b = 0; b < 2; ++b) {
    p->elite_actions[b].alloc(B * p->k * p->stride);
    p->elite_costs[b].alloc(B * p->k);
    p->elite_idx[b].alloc(B * p->k);
  }
  p->trace_mean.alloc(B * p->iters * p->hd);
  p->trace_std.alloc(B * p->iters * p->hd);
  p->trace_costs.alloc(B * p->iters * p->k);
  p->trace_idx.alloc(B * p->iters * p->k);
  p->out.alloc(B * (p->d + 1));
  p->cand.alloc(B * p->sm_count * p->k);
  p->ticket.alloc(B);
  build_plan(p.get());

  if (cfg->dynamics == ICEM_DYN_HALFCHEETAH || cfg->dynamics == ICEM_DYN_HUMANOID_STANDUP ||
      cfg->dynamics == ICEM_DYN_ARTICULATED) {
    p->state_dim = 0;   // known once icem_set_articulated_model provides the tables
  } else if (cfg->dynamics == ICEM_DYN_DENSE_TANH || cfg->dynamics == ICEM_DYN_MLP) {
    p->state_dim = 0;   // known once the model is set
  } else {
    throw Unsupported("dynamics id not supported by this build");
  }
  if (cfg->world_size > 1) {
    p->rec_bytes = (size_t)p->k * sizeof(unsigned long long) + (size_t)p->k * p->stride * sizeof(float);
    p->send_rec.alloc(p->rec_bytes);
    p->recv_rec.alloc(p->rec_bytes * cfg->world_size);
  }
  // staging sized for the largest state we support
  p->in_bytes = sizeof(StepState) + B * kStateSlot * sizeof(float);
  p->step_in.alloc(p->in_bytes);
  ICEM_CUDA(cudaMallocHost(&p->h_in, p->in_bytes));
  memset(p->h_in, 0, p->in_bytes);
  ICEM_CUDA(cudaMallocHost(&p->h_out, B * (p->d + 1) * sizeof(float)));
  ICEM_CUDA(cudaStreamSynchronize(p->stream));
  *out = p.release();
  ICEM_API_END
}

int icem_destroy(icem_planner_t* p) {
  ICEM_API_BEGIN
  if (p) {
    cudaSetDevice(p->cfg.device);
    cudaStreamSynchronize(p->stream);
    delete p;
  }
  ICEM_API_END
}

int icem_set_dense_model(icem_planner_t* p, int32_t obs_dim, const float* w_obs, const float* w_act,
                         const float* bias) {
  ICEM_API_BEGIN
  if (!p || !w_obs || !w_act) throw InvalidArg("null argument");
  if (p->cfg.dynamics != ICEM_DYN_DENSE_TANH) throw InvalidArg("planner was not created with ICEM_DYN_DENSE_TANH");
  if (obs_dim < 1 || obs_dim > 256) throw InvalidArg("obs_dim out of range [1,256]");
  ICEM_CUDA(cudaSetDevice(p->cfg.device));
  upload(p->w_obs, std::vector<float>(w_obs, w_obs + (size_t)obs_dim * obs_dim));
  upload(p->w_act, std::vector<float>(w_act, w_act + (size_t)obs_dim * p->d));
  upload(p->bias, bias ? std::vector<float>(bias, bias + obs_dim) : std::vector<float>(obs_dim, 0.f));
  p->dense.obs_dim = obs_dim;
  p->dense.act_dim = p->d;
  p->dense.w_obs = p->w_obs.p;
  p->dense.w_act = p->w_act.p;
  p->dense.bias = p->bias.p;
  p->state_dim = obs_dim;
  if (p->obs_dim <= 0) p->obs_dim = obs_dim;
  p->model_ready = true;
  if (p->graph_exec) { cudaGraphExecDestroy(p->graph_exec); p->graph_exec = nullptr; }
  ICEM_API_END
}

int icem_set_articulated_model(icem_planner_t* p, const icem_articulated_model_t* a) {
  ICEM_API_BEGIN
  if (!p || !a) throw InvalidArg("null argument");
  if (p->cfg.dynamics != ICEM_DYN_HALFCHEETAH && p->cfg.dynamics != ICEM_DYN_HUMANOID_STANDUP &&
      p->cfg.dynamics != ICEM_DYN_ARTICULATED)
    throw InvalidArg("planner was not created with an articulated dynamics id");
  if (a->nb < 1 || a->nb > kArtMaxBodies || a->nv < 1 || a->nv > kArtMaxDofs || a->nc < 0 ||
      a->nc > kArtMaxContacts || a->nq < a->nv || a->nq + a->nv > 64 || a->nsub < 1)
    throw InvalidArg("articulated model dimensions out of range (bodies<=16, dofs<=32, contacts<=32, nq+nv<=64)");
  if (a->nu != p->d) throw InvalidArg("articulated model control dimension does not match act_dim");
  if (a->obs_offset < 0 || a->obs_offset >= a->nq) throw InvalidArg("obs_offset out of range");
  if (!(a->dt > 0)) throw InvalidArg("dt must be positive");
  ICEM_CUDA(cudaSetDevice(p->cfg.device));
  std::unique_ptr<ArtModel> mp(new ArtModel);
  ArtModel& m = *mp;
  memset(&m, 0, sizeof m);
  m.nb = a->nb; m.nq = a->nq; m.nv = a->nv; m.nu = a->nu; m.nc = a->nc; m.nsub = a->nsub;
  m.obs_offset = a->obs_offset;
  m.dt = a->dt; m.gravity = a->gravity; m.ctrl_limit = a->ctrl_limit;
  m.kc = a->contact_stiffness; m.cc = a->contact_damping; m.kv = a->friction_viscous; m.mu = a->friction;
  m.cdmax = a->contact_damping_max;
  for (int b = 0; b < m.nb; ++b) {
    const int par = a->body_parent[b];
    if (par >= b || par < -1) throw InvalidArg("bodies must be listed parent-before-child");
    m.b_parent[b] = par;
    m.b_depth[b] = par < 0 ? 0 : m.b_depth[par] + 1;
    m.max_depth = std::max(m.max_depth, m.b_depth[b]);
    m.b_dof_start[b] = a->body_dof_start[b];
    m.b_dof_count[b] = a->body_dof_count[b];
    if (m.b_dof_start[b] < 0 || m.b_dof_count[b] < 0 || m.b_dof_start[b] + m.b_dof_count[b] > m.nv)
      throw InvalidArg("body dof range out of bounds");
    if (par >= 0) {
      if (m.b_nchild[par] >= kArtMaxChildren) throw InvalidArg("a body has more than 4 children");
      m.b_child[par][m.b_nchild[par]++] = b;
    }
    for (int i = 0; i < 3; ++i) { m.b_pos[b][i] = a->body_pos[3 * b + i]; m.b_com[b][i] = a->body_com[3 * b + i]; }
    for (int i = 0; i < 6; ++i) m.b_inertia[b][i] = a->body_inertia[6 * b + i];
    m.b_mass[b] = a->body_mass[b];
  }
  for (int j = 0; j < m.nv; ++j) {
    m.d_body[j] = a->dof_body[j]; m.d_type[j] = a->dof_type[j]; m.d_qadr[j] = a->dof_qadr[j];
    m.d_limited[j] = a->dof_limited[j]; m.d_act[j] = a->dof_act[j];
    if (m.d_body[j] < 0 || m.d_body[j] >= m.nb || m.d_type[j] < 0 || m.d_type[j] > 3 || m.d_qadr[j] < 0 ||
        m.d_qadr[j] >= m.nq || m.d_act[j] >= m.nu)
      throw InvalidArg("dof table entry out of range");
    const int par = a->dof_parent[j];
    if (par >= j || par < -1) throw InvalidArg("dof_parent must point to an earlier dof");
    m.d_chain[j] = (1u << j) | (par >= 0 ? m.d_chain[par] : 0u);
    for (int i = 0; i < 3; ++i) { m.d_axis[j][i] = a->dof_axis[3 * j + i]; m.d_anchor[j][i] = a->dof_anchor[3 * j + i]; }
    m.d_stiff[j] = a->dof_stiffness[j]; m.d_damp[j] = a->dof_damping[j]; m.d_arm[j] = a->dof_armature[j];
    m.d_lo[j] = a->dof_lo[j]; m.d_hi[j] = a->dof_hi[j]; m.d_klim[j] = a->dof_klim[j]; m.d_blim[j] = a->dof_blim[j];
    m.d_gear[j] = a->dof_gear[j];
  }
  // a free joint must be 3 translations followed by 3 rotations on the same body
  for (int j = 0; j < m.nv; ++j) {
    if (m.d_type[j] == kFreeRot) throw InvalidArg("free-joint rotation dofs must follow three translation dofs");
    if (m.d_type[j] != kFreeTrans) continue;
    if (j + 5 >= m.nv) throw InvalidArg("incomplete free joint");
    for (int k = 0; k < 6; ++k)
      if (m.d_type[j + k] != (k < 3 ? kFreeTrans : kFreeRot) || m.d_body[j + k] != m.d_body[j])
        throw InvalidArg("free joint must be 3 translation dofs followed by 3 rotation dofs");
    j += 5;
  }
  // derived tables: chain ends, per-level parents, pointer-doubling jumps, frame-velocity references
  if (m.max_depth >= kArtMaxDepth) throw InvalidArg("kinematic tree deeper than 8 levels");
  for (int b = 0; b < m.nb; ++b) {
    m.b_last_dof[b] = m.b_dof_count[b] > 0 ? m.b_dof_start[b] + m.b_dof_count[b] - 1
                                            : (m.b_parent[b] >= 0 ? m.b_last_dof[m.b_parent[b]] : -1);
    if (m.b_nchild[b] > 0) {
      int& np = m.lvl_np[m.b_depth[b]];
      if (np >= kArtMaxLevelParents) throw InvalidArg("more than 8 bodies with children on one tree level");
      for (int k = 0; k < kArtMaxChildren; ++k) m.lvl_child[m.b_depth[b]][np][k] = k < m.b_nchild[b] ? m.b_child[b][k] : -1;
      m.lvl_parent[m.b_depth[b]][np++] = b;
    }
  }
  int max_chain = 1;
  for (int j = 0; j < m.nv; ++j) {
    m.d_jump[0][j] = a->dof_parent[j];
    max_chain = std::max(max_chain, __builtin_popcount(m.d_chain[j]));
    if (m.d_type[j] == kFreeRot) {
      int last = j;
      while (last + 1 < m.nv && m.d_type[last + 1] == kFreeRot && m.d_body[last + 1] == m.d_body[j]) ++last;
      m.d_vref[j] = last;
    } else {
      m.d_vref[j] = a->dof_parent[j];
    }
  }
  for (int r = 1; r < kArtScanRounds; ++r)
    for (int j = 0; j < m.nv; ++j) {
      const int mid = m.d_jump[r - 1][j];
      m.d_jump[r][j] = mid >= 0 ? m.d_jump[r - 1][mid] : -1;
    }
  m.scan_rounds = 0;
  while ((1 << m.scan_rounds) < max_chain) ++m.scan_rounds;
  if (m.scan_rounds > kArtScanRounds) throw InvalidArg("dof chain longer than 32");
  int prev_body = -1;
  for (int c = 0; c < m.nc; ++c) {
    const int b = a->con_body[c];
    if (b < 0 || b >= m.nb || b < prev_body) throw InvalidArg("contact spheres must be sorted by body");
    if (b != prev_body) m.b_con_start[b] = c;
    m.b_con_count[b] += 1;
    prev_body = b;
    m.c_body[c] = b;
    for (int i = 0; i < 3; ++i) m.c_pos[c][i] = a->con_pos[3 * c + i];
    m.c_radius[c] = a->con_radius[c];
  }
  {   // padding children point at the all-zero record of the size class that will run this model
    const int zero_rec = Articulated<12>::fits(m.nb, m.nv, m.nc) ? Articulated<12>::kZeroRec
                         : (Articulated<24>::fits(m.nb, m.nv, m.nc) ? Articulated<24>::kZeroRec : Articulated<32>::kZeroRec);
    for (int L = 0; L < kArtMaxDepth; ++L)
      for (int s2 = 0; s2 < kArtMaxLevelParents; ++s2)
        for (int k = 0; k < kArtMaxChildren; ++k)
          if (m.lvl_child[L][s2][k] < 0) m.lvl_child[L][s2][k] = zero_rec;
  }
  p->art_model.alloc(1);
  ICEM_CUDA(cudaMemcpy(p->art_model.p, &m, sizeof m, cudaMemcpyHostToDevice));
  {
    const char* why = "";
    p->chain_ok = !p->force_warp_engine && build_chain_model(chain_source(*a), p->d, p->chain_host, &why);
    if (p->chain_ok) {
      // ICEM_B200_PLANAR=0: run a planar robot through the spatial instantiation of the engine (A/B, tests)
      { const char* e = getenv("ICEM_B200_PLANAR"); if (e && e[0] == '0') p->chain_host.planar = 0; }
      p->chain_model.alloc(1);
      ICEM_CUDA(cudaMemcpy(p->chain_model.p, &p->chain_host, sizeof(ChainModel), cudaMemcpyHostToDevice));
      p->chain.model = p->chain_model.p;
      p->chain.act_dim = p->d;
      p->chain.warp_floats = chain_warp_floats(p->chain_host);
      p->chain.rows_per_warp = 32 / p->chain_host.lanes;
    } else if (a->integrator != ICEM_INTEGRATOR_EULER) {
      throw Unsupported(std::string("the Runge-Kutta integrator needs the branch-parallel engine, which cannot run "
                                    "this robot: ") + why);
    }
  }
  p->art.model = p->art_model.p;
  p->art.act_dim = p->d; p->art.nq = m.nq; p->art.nv = m.nv; p->art.nb = m.nb; p->art.nc = m.nc;
  p->state_dim = m.nq + m.nv;
  if (p->obs_dim <= 0) p->obs_dim = p->state_dim - m.obs_offset;
  p->model_ready = true;
  if (p->graph_exec) { cudaGraphExecDestroy(p->graph_exec); p->graph_exec = nullptr; }
  ICEM_API_END
}

int icem_set_mlp_model(icem_planner_t* p, int32_t n_layers, const int32_t* dims, const float* const* weights,
                       const float* const* biases) {
  ICEM_API_BEGIN
  if (!p || !dims || !weights || !biases) throw InvalidArg("null argument");
  if (p->cfg.dynamics != ICEM_DYN_MLP) throw InvalidArg("planner was not created with ICEM_DYN_MLP");
  if (n_layers != 3) throw Unsupported("the tensor-core rollout supports exactly two hidden layers (n_layers == 3)");
  const int in = dims[0], H = dims[1], out = dims[3];
  if (dims[2] != H || (H != 64 && H != 128 && H != 256))
    throw Unsupported("hidden widths must be equal and one of 64, 128, 256");
  if (out < 1 || out > kMlpOutPad) throw Unsupported("observation width must be in [1, 32]");
  if (in != out + p->d || in > kMlpInPad) throw InvalidArg("input width must be obs_dim + act_dim <= 32");
  for (int l = 0; l < 3; ++l)
    if (!weights[l] || !biases[l]) throw InvalidArg("null layer parameters");
  ICEM_CUDA(cudaSetDevice(p->cfg.device));
  auto pack = [](const float* w, int rows, int cols, int rows_pad, int K, DevBuf<mlp_op_t>& dst) {
    std::vector<mlp_op_t> h((size_t)rows_pad * K, mlp_op_from_float(0.f));
    for (int n = 0; n < rows; ++n)
      for (int k = 0; k < cols; ++k) h[umma_pack_index(n, k, K)] = mlp_op_from_float(w[(size_t)n * cols + k]);
    dst.alloc(h.size());
    ICEM_CUDA(cudaMemcpy(dst.p, h.data(), h.size() * sizeof(mlp_op_t), cudaMemcpyHostToDevice));
  };
  // layer-1 input columns: [obs | 0 | act at act_off | 0], act_off = obs width rounded up to 8 (the kernel
  // indexes its register state statically per 8-column chunk)
  const int act_off = (out + 7) & ~7;
  if (act_off + p->d > kMlpInPad)
    throw Unsupported("obs_dim rounded up to 8, plus act_dim, must be <= 32 for the tensor-core rollout");
  {
    std::vector<float> w1((size_t)H * kMlpInPad, 0.f);
    for (int n = 0; n < H; ++n) {
      for (int k = 0; k < out; ++k) w1[(size_t)n * kMlpInPad + k] = weights[0][(size_t)n * in + k];
      for (int m = 0; m < p->d; ++m) w1[(size_t)n * kMlpInPad + act_off + m] = weights[0][(size_t)n * in + out + m];
    }
    pack(w1.data(), H, kMlpInPad, H, kMlpInPad, p->mlp_w1);
  }
  p->mlp.act_off = act_off;
  pack(weights[1], H, H, H, H, p->mlp_w2);
  pack(weights[2], out, H, kMlpOutPad, H, p->mlp_w3);
  std::vector<float> b((size_t)2 * H + kMlpOutPad, 0.f);
  for (int i = 0; i < H; ++i) { b[i] = biases[0][i]; b[H + i] = biases[1][i]; }
  for (int i = 0; i < out; ++i) b[2 * H + i] = biases[2][i];
  upload(p->mlp_bias, b);
  p->mlp.obs_dim = out; p->mlp.act_dim = p->d; p->mlp.hidden = H;
  p->mlp.w1 = p->mlp_w1.p; p->mlp.w2 = p->mlp_w2.p; p->mlp.w3 = p->mlp_w3.p; p->mlp.bias = p->mlp_bias.p;
  p->state_dim = out;
  if (p->obs_dim <= 0) p->obs_dim = out;
  p->model_ready = true;
  if (p->graph_exec) { cudaGraphExecDestroy(p->graph_exec); p->graph_exec = nullptr; }
  ICEM_API_END
}

int icem_begin_rollout(icem_planner_t* p) {
  ICEM_API_BEGIN
  if (!p) throw InvalidArg("null planner");
  ICEM_CUDA(cudaSetDevice(p->cfg.device));
  reset_distribution(p);
  ICEM_CUDA(cudaStreamSynchronize(p->stream));
  p->was_reset = true;
  p->step = 0;
  p->has_prev = false;
  ICEM_API_END
}

// launch half of a plan step: stage the inputs, enqueue the CEM iterations, no host synchronisation
static void plan_launch(icem_planner* p, const double* state, int32_t state_dim) {
  if (!p->was_reset) throw StateError("beginning_of_rollout() needs to be called before");   // icem.py:109-110
  require_model(p);
  if (state_dim != p->state_dim) throw InvalidArg("state_dim does not match the forward model");
  if (p->plan_in_flight) throw StateError("a plan step is already in flight (icem_plan_finish not called)");
  ICEM_CUDA(cudaSetDevice(p->cfg.device));
  write_step_in(p, state);
  ICEM_CUDA(cudaEventRecord(p->ev_a, p->stream));
  if (p->inject_pending) {
    // parity mode: direct launches (injected buffers are per-call)
    ICEM_CUDA(cudaMemcpyAsync(p->step_in.p, p->h_in, p->in_bytes, cudaMemcpyHostToDevice, p->stream));
    enqueue_iterations(p, false);
    ICEM_CUDA(cudaMemcpyAsync(p->h_out, p->out.p, (size_t)p->B * (p->d + 1) * sizeof(float), cudaMemcpyDeviceToHost,
                              p->stream));
  } else {
    ensure_graph(p);
    ICEM_CUDA(cudaGraphLaunch(p->graph_exec, p->stream));
  }
  ICEM_CUDA(cudaEventRecord(p->ev_b, p->stream));
  p->plan_in_flight = true;
}

static void plan_finish(icem_planner* p, double* action_out) {
  if (!p->plan_in_flight) throw StateError("no plan step in flight (icem_plan_async not called)");
  ICEM_CUDA(cudaSetDevice(p->cfg.device));
  ICEM_CUDA(cudaStreamSynchronize(p->stream));
  p->plan_in_flight = false;
  ICEM_CUDA(cudaEventElapsedTime(&p->last_total_ms, p->ev_a, p->ev_b));
  p->last_rollout_ms = 0.f;
  for (int b = 0; b < p->B; ++b)
    for (int i = 0; i < p->d; ++i) action_out[(size_t)b * p->d + i] = (double)p->h_out[(size_t)b * (p->d + 1) + i];
  finish_step(p);
}

int icem_plan_batch(icem_planner_t* p, const double* states, int32_t state_dim, int32_t num_states,
                    double* actions_out) {
  ICEM_API_BEGIN
  if (!p || !states || !actions_out) throw InvalidArg("null argument");
  if (num_states != p->B) throw InvalidArg("num_states must equal the handle's num_problems");
  plan_launch(p, states, state_dim);
  plan_finish(p, actions_out);
  ICEM_API_END
}

int icem_num_problems(icem_planner_t* p) { return p ? p->B : -1; }

int icem_set_active_problem(icem_planner_t* p, int32_t problem) {
  ICEM_API_BEGIN
  if (!p) throw InvalidArg("null planner");
  if (problem < 0 || problem >= p->B) throw InvalidArg("problem index out of range");
  p->active = problem;
  ICEM_API_END
}

int icem_plan(icem_planner_t* p, const double* state, int32_t state_dim, double* action_out) {
  ICEM_API_BEGIN
  if (!p || !state || !action_out) throw InvalidArg("null argument");
  if (p->B != 1) throw StateError("this handle batches several problems: use icem_plan_batch");
  plan_launch(p, state, state_dim);
  plan_finish(p, action_out);
  ICEM_API_END
}

int icem_plan_async(icem_planner_t* p, const double* state, int32_t state_dim) {
  ICEM_API_BEGIN
  if (!p || !state) throw InvalidArg("null argument");
  if (p->B != 1) throw StateError("this handle batches several problems: use icem_plan_batch");
  plan_launch(p, state, state_dim);
  ICEM_API_END
}

int icem_plan_finish(icem_planner_t* p, double* action_out) {
  ICEM_API_BEGIN
  if (!p || !action_out) throw InvalidArg("null argument");
  plan_finish(p, action_out);
  ICEM_API_END
}

int icem_plan_device(icem_planner_t* p) {
  ICEM_API_BEGIN
  if (p && p->B != 1) throw Unsupported("not available on a handle that batches several problems");
  if (!p) throw InvalidArg("null planner");
  if (!p->was_reset) throw StateError("beginning_of_rollout() needs to be called before");
  require_model(p);
  ICEM_CUDA(cudaSetDevice(p->cfg.device));
  // only the StepState words change; the start state stays resident
  write_step_in(p, nullptr);
  ICEM_CUDA(cudaMemcpyAsync(p->step_in.p, p->h_in, sizeof(StepState), cudaMemcpyHostToDevice, p->stream));
  enqueue_iterations(p, true);
  finish_step(p);
  ICEM_API_END
}

int icem_advance_state_device(icem_planner_t* p) {
  ICEM_API_BEGIN
  if (p && p->B != 1) throw Unsupported("not available on a handle that batches several problems");
  if (!p) throw InvalidArg("null planner");
  require_model(p);
  ICEM_CUDA(cudaSetDevice(p->cfg.device));
  advance_dyn(p, start_state_dev(p), p->out.p, start_state_dev(p), nullptr, 0);
  ICEM_API_END
}

int icem_sync(icem_planner_t* p) {
  ICEM_API_BEGIN
  if (!p) throw InvalidArg("null planner");
  ICEM_CUDA(cudaSetDevice(p->cfg.device));
  ICEM_CUDA(cudaStreamSynchronize(p->stream));
  ICEM_API_END
}

int icem_last_plan_ms(icem_planner_t* p, float* total_ms, float* rollout_ms) {
  ICEM_API_BEGIN
  if (!p) throw InvalidArg("null planner");
  if (total_ms) *total_ms = p->last_total_ms;
  if (rollout_ms) *rollout_ms = p->last_rollout_ms;
  ICEM_API_END
}

int icem_inject_noise(icem_planner_t* p, int32_t iteration, int32_t rows, const float* zr, const float* zi) {
  ICEM_API_BEGIN
  if (p && p->B != 1) throw Unsupported("not available on a handle that batches several problems");
  if (!p || !zr) throw InvalidArg("null argument");
  if (iteration < 0 || iteration >= p->iters) throw InvalidArg("iteration out of range");
  if (!p->white && !zi) throw InvalidArg("zi required for colored noise");
  ICEM_CUDA(cudaSetDevice(p->cfg.device));
  const int need = rows_of(p, iteration, !p->has_prev);
  if (rows != need) throw InvalidArg("injected row count does not match this rank's population");
  if (p->inj_zr.size() != (size_t)p->iters) {
    p->inj_zr = std::vector<DevBuf<float>>(p->iters);
    p->inj_zi = std::vector<DevBuf<float>>(p->iters);
    p->inj_rows.assign(p->iters, 0);
  }
  const size_t per_row = p->white ? (size_t)p->hd : (size_t)p->d * p->K;
  p->inj_zr[iteration].alloc(per_row * rows);
  ICEM_CUDA(cudaMemcpy(p->inj_zr[iteration].p, zr, per_row * rows * sizeof(float), cudaMemcpyHostToDevice));
  if (!p->white) {
    p->inj_zi[iteration].alloc(per_row * rows);
    ICEM_CUDA(cudaMemcpy(p->inj_zi[iteration].p, zi, per_row * rows * sizeof(float), cudaMemcpyHostToDevice));
  }
  p->inj_rows[iteration] = rows;
  bool all = true;
  for (int i = 0; i < p->iters; ++i) all = all && p->inj_rows[i] > 0;
  p->inject_pending = all;
  ICEM_API_END
}

int icem_get_mean(icem_planner_t* p, float* out) {
  ICEM_API_BEGIN
  if (!p || !out) throw InvalidArg("null argument");
  ICEM_CUDA(cudaSetDevice(p->cfg.device));
  ICEM_CUDA(cudaStreamSynchronize(p->stream));
  ICEM_CUDA(cudaMemcpy(out, p->mean.p + (size_t)p->active * p->hd, p->hd * sizeof(float), cudaMemcpyDeviceToHost));
  ICEM_API_END
}

int icem_get_std(icem_planner_t* p, float* out) {
  ICEM_API_BEGIN
  if (!p || !out) throw InvalidArg("null argument");
  ICEM_CUDA(cudaSetDevice(p->cfg.device));
  ICEM_CUDA(cudaStreamSynchronize(p->stream));
  ICEM_CUDA(cudaMemcpy(out, p->stdv.p + (size_t)p->active * p->hd, p->hd * sizeof(float), cudaMemcpyDeviceToHost));
  ICEM_API_END
}

int icem_num_elites(icem_planner_t* p) { return p ? p->k : -1; }
int icem_state_dim(icem_planner_t* p) { return p ? p->state_dim : -1; }

int icem_get_elites(icem_planner_t* p, float* actions_out, float* costs_out, int32_t* idx_out) {
  ICEM_API_BEGIN
  if (!p) throw InvalidArg("null planner");
  if (!p->has_prev) throw StateError("no elites yet (no plan step since beginning_of_rollout)");
  ICEM_CUDA(cudaSetDevice(p->cfg.device));
  ICEM_CUDA(cudaStreamSynchronize(p->stream));
  const int b = (p->iters - 1) & 1;
  if (actions_out)
    ICEM_CUDA(cudaMemcpy2D(actions_out, p->hd * sizeof(float),
                           p->elite_actions[b].p + (size_t)p->active * p->k * p->stride, p->stride * sizeof(float),
                           p->hd * sizeof(float), p->k, cudaMemcpyDeviceToHost));
  const size_t ak = (size_t)p->active * p->k;
  if (costs_out) ICEM_CUDA(cudaMemcpy(costs_out, p->elite_costs[b].p + ak, p->k * sizeof(float), cudaMemcpyDeviceToHost));
  if (idx_out) ICEM_CUDA(cudaMemcpy(idx_out, p->elite_idx[b].p + ak, p->k * sizeof(int32_t), cudaMemcpyDeviceToHost));
  ICEM_API_END
}

int icem_population_size(icem_planner_t* p, int32_t iteration, int32_t first_step, int32_t* global_out,
                         int32_t* local_out) {
  ICEM_API_BEGIN
  if (!p) throw InvalidArg("null planner");
  if (iteration < 0 || iteration >= p->iters) throw InvalidArg("iteration out of range");
  const IterPlan& ip = p->plan[iteration];
  const bool shift = iteration == 0 && !first_step && p->cfg.shift_elites_over_time;
  if (global_out) *global_out = ip.n_global + (shift ? p->n_keep : 0);
  if (local_out) *local_out = ip.n_fresh_local + (shift ? ip.n_shift_local : 0);
  ICEM_API_END
}

int icem_get_iteration(icem_planner_t* p, int32_t i, float* mean_out, float* std_out, float* elite_costs_out,
                       int32_t* elite_idx_out) {
  ICEM_API_BEGIN
  if (!p) throw InvalidArg("null planner");
  if (i < 0 || i >= p->iters) throw InvalidArg("iteration out of range");
  ICEM_CUDA(cudaSetDevice(p->cfg.device));
  ICEM_CUDA(cudaStreamSynchronize(p->stream));
  const size_t th = ((size_t)p->active * p->iters + i) * p->hd, tk = ((size_t)p->active * p->iters + i) * p->k;
  if (mean_out) ICEM_CUDA(cudaMemcpy(mean_out, p->trace_mean.p + th, p->hd * 4, cudaMemcpyDeviceToHost));
  if (std_out) ICEM_CUDA(cudaMemcpy(std_out, p->trace_std.p + th, p->hd * 4, cudaMemcpyDeviceToHost));
  if (elite_costs_out)
    ICEM_CUDA(cudaMemcpy(elite_costs_out, p->trace_costs.p + tk, p->k * 4, cudaMemcpyDeviceToHost));
  if (elite_idx_out)
    ICEM_CUDA(cudaMemcpy(elite_idx_out, p->trace_idx.p + tk, p->k * 4, cudaMemcpyDeviceToHost));
  ICEM_API_END
}

int icem_get_costs(icem_planner_t* p, int32_t i, float* out, int32_t n) {
  ICEM_API_BEGIN
  if (!p || !out) throw InvalidArg("null argument");
  if (i < 0 || i >= p->iters) throw InvalidArg("iteration out of range");
  if (n < 0 || n > p->plan[i].rows_cap) throw InvalidArg("row count out of range");
  ICEM_CUDA(cudaSetDevice(p->cfg.device));
  ICEM_CUDA(cudaStreamSynchronize(p->stream));
  ICEM_CUDA(cudaMemcpy(out, p->costs.p + (size_t)p->active * p->costs_per + p->plan[i].cost_off, (size_t)n * 4,
                       cudaMemcpyDeviceToHost));
  ICEM_API_END
}

int icem_get_actions(icem_planner_t* p, int32_t i, float* out, int32_t n) {
  ICEM_API_BEGIN
  if (!p || !out) throw InvalidArg("null argument");
  if (i < 0 || i >= p->iters) throw InvalidArg("iteration out of range");
  if (!p->cfg.keep_iteration_actions && i != p->iters - 1)
    throw StateError("create the planner with keep_iteration_actions=1 to read earlier populations");
  if (n < 0 || n > p->plan[i].rows_cap) throw InvalidArg("row count out of range");
  ICEM_CUDA(cudaSetDevice(p->cfg.device));
  ICEM_CUDA(cudaStreamSynchronize(p->stream));
  if (n)
    ICEM_CUDA(cudaMemcpy2D(out, p->hd * sizeof(float),
                           p->actions.p + (size_t)p->active * p->actions_per + p->plan[i].act_off * p->stride,
                           p->stride * sizeof(float), p->hd * sizeof(float), n, cudaMemcpyDeviceToHost));
  ICEM_API_END
}

int icem_sim_step(icem_planner_t* p, const double* state, int32_t state_dim, const double* action,
                  double* next_state, double* obs_out, int32_t obs_dim, double* reward_out) {
  ICEM_API_BEGIN
  if (!p || !state) throw InvalidArg("null argument");
  require_model(p);
  if (state_dim != p->state_dim) throw InvalidArg("state_dim does not match the forward model");
  if (obs_dim < 0 || obs_dim > 512) throw InvalidArg("obs_dim out of range");
  ICEM_CUDA(cudaSetDevice(p->cfg.device));
  DevBuf<float>& buf = p->step_scratch;
  const int sd = p->state_dim;
  buf.reserve((size_t)2 * sd + p->d + std::max(obs_dim, 1));
  std::vector<float> h((size_t)sd + p->d);
  for (int i = 0; i < sd; ++i) h[i] = (float)state[i];
  if (action)
    for (int i = 0; i < p->d; ++i) h[sd + i] = (float)action[i];
  ICEM_CUDA(cudaMemcpyAsync(buf.p, h.data(), h.size() * 4, cudaMemcpyHostToDevice, p->stream));
  float* d_state = buf.p;
  float* d_act = buf.p + sd;
  float* d_next = d_act + p->d;
  float* d_obs = d_next + sd;
  advance_dyn(p, d_state, action ? d_act : nullptr, d_next, obs_out ? d_obs : nullptr, obs_dim);
  std::vector<float> r((size_t)sd + obs_dim);
  ICEM_CUDA(cudaMemcpyAsync(r.data(), d_next, r.size() * 4, cudaMemcpyDeviceToHost, p->stream));
  ICEM_CUDA(cudaStreamSynchronize(p->stream));
  if (next_state)
    for (int i = 0; i < sd; ++i) next_state[i] = r[i];
  if (obs_out)
    for (int i = 0; i < obs_dim; ++i) obs_out[i] = r[sd + i];
  if (reward_out) *reward_out = 0.0;
  ICEM_API_END
}

int icem_sim_step_batch(icem_planner_t* p, int32_t n, const double* states, int32_t state_dim, const double* actions,
                        double* next_states) {
  ICEM_API_BEGIN
  if (!p || !states || !actions || !next_states) throw InvalidArg("null argument");
  if (n < 1 || n > 65535) throw InvalidArg("n must be in [1, 65535]");
  require_model(p);
  if (state_dim != p->state_dim) throw InvalidArg("state_dim does not match the forward model");
  ICEM_CUDA(cudaSetDevice(p->cfg.device));
  const int sd = p->state_dim, d = p->d;
  DevBuf<float>& buf = p->step_scratch;
  buf.reserve((size_t)n * (2 * sd + d));
  std::vector<float> h((size_t)n * (sd + d));
  for (size_t i = 0; i < (size_t)n * sd; ++i) h[i] = (float)states[i];
  for (size_t i = 0; i < (size_t)n * d; ++i) h[(size_t)n * sd + i] = (float)actions[i];
  ICEM_CUDA(cudaMemcpyAsync(buf.p, h.data(), h.size() * 4, cudaMemcpyHostToDevice, p->stream));
  float* d_state = buf.p;
  float* d_act = buf.p + (size_t)n * sd;
  float* d_next = d_act + (size_t)n * d;
  advance_dyn(p, d_state, d_act, d_next, nullptr, 0, n, AdvanceBatch{sd, d, 0});
  std::vector<float> r((size_t)n * sd);
  ICEM_CUDA(cudaMemcpyAsync(r.data(), d_next, r.size() * 4, cudaMemcpyDeviceToHost, p->stream));
  ICEM_CUDA(cudaStreamSynchronize(p->stream));
  for (size_t i = 0; i < r.size(); ++i) next_states[i] = r[i];
  ICEM_API_END
}

int icem_observe(icem_planner_t* p, const double* state, int32_t state_dim, double* obs_out, int32_t obs_dim) {
  return icem_sim_step(p, state, state_dim, nullptr, nullptr, obs_out, obs_dim, nullptr);
}

// ---- single operators ------------------------------------------------------------------------------
int icem_op_sample(icem_planner_t* p, int32_t n, const float* zr, const float* zi, const float* mean,
                   const float* std, float* actions_out) {
  ICEM_API_BEGIN
  if (!p || !zr || !mean || !std || !actions_out) throw InvalidArg("null argument");
  if (!p->white && !zi) throw InvalidArg("zi required for colored noise");
  if (n < 1) throw InvalidArg("n must be positive");
  require_model(p);
  ICEM_CUDA(cudaSetDevice(p->cfg.device));
  const size_t per_row = p->white ? (size_t)p->hd : (size_t)p->d * p->K;
  DevBuf<float> d_zr, d_zi, d_mean, d_std, d_act, d_cost;
  DevBuf<unsigned char> d_ss;
  d_zr.alloc(per_row * n);
  ICEM_CUDA(cudaMemcpy(d_zr.p, zr, per_row * n * 4, cudaMemcpyHostToDevice));
  if (!p->white) {
    d_zi.alloc(per_row * n);
    ICEM_CUDA(cudaMemcpy(d_zi.p, zi, per_row * n * 4, cudaMemcpyHostToDevice));
  }
  d_mean.alloc(p->hd); d_std.alloc(p->hd);
  ICEM_CUDA(cudaMemcpy(d_mean.p, mean, p->hd * 4, cudaMemcpyHostToDevice));
  ICEM_CUDA(cudaMemcpy(d_std.p, std, p->hd * 4, cudaMemcpyHostToDevice));
  d_act.alloc((size_t)n * p->stride);
  d_cost.alloc(n);
  d_ss.alloc(sizeof(StepState) + 256 * sizeof(float));
  StepState ss{0, 0, 1, 0};
  ICEM_CUDA(cudaMemcpy(d_ss.p, &ss, sizeof ss, cudaMemcpyHostToDevice));
  RolloutArgs a{};
  a.n_fresh_local = n; a.n_fresh_global = n; a.stride = p->stride;
  a.actions = d_act.p; a.costs = d_cost.p; a.mean = d_mean.p; a.std = d_std.p;
  a.prev_elites = d_act.p;
  a.start_state = reinterpret_cast<float*>(d_ss.p + sizeof(StepState));
  a.inj_zr = d_zr.p; a.inj_zi = d_zi.p;
  a.ss = reinterpret_cast<StepState*>(d_ss.p);
  launch_rollout_dyn<true, false>(p, a, n);
  ICEM_CUDA(cudaStreamSynchronize(p->stream));
  ICEM_CUDA(cudaMemcpy2D(actions_out, p->hd * sizeof(float), d_act.p, p->stride * sizeof(float),
                         p->hd * sizeof(float), n, cudaMemcpyDeviceToHost));
  ICEM_API_END
}

int icem_op_rollout_cost(icem_planner_t* p, int32_t n, const double* state, int32_t state_dim,
                         const float* actions, float* costs_out) {
  ICEM_API_BEGIN
  if (!p || !state || !actions || !costs_out) throw InvalidArg("null argument");
  if (n < 1) throw InvalidArg("n must be positive");
  require_model(p);
  if (state_dim != p->state_dim) throw InvalidArg("state_dim does not match the forward model");
  ICEM_CUDA(cudaSetDevice(p->cfg.device));
  DevBuf<float> d_act, d_cost, d_ms;
  DevBuf<unsigned char> d_ss;
  d_act.alloc((size_t)n * p->stride);
  ICEM_CUDA(cudaMemcpy2D(d_act.p, p->stride * sizeof(float), actions, p->hd * sizeof(float), p->hd * sizeof(float), n,
                         cudaMemcpyHostToDevice));
  d_cost.alloc(n);
  d_ms.alloc(2 * (size_t)p->hd);
  d_ss.alloc(sizeof(StepState) + 256 * sizeof(float));
  std::vector<unsigned char> hs(sizeof(StepState) + 256 * sizeof(float), 0);
  float* f = reinterpret_cast<float*>(hs.data() + sizeof(StepState));
  for (int i = 0; i < state_dim; ++i) f[i] = (float)state[i];
  ICEM_CUDA(cudaMemcpy(d_ss.p, hs.data(), hs.size(), cudaMemcpyHostToDevice));
  RolloutArgs a{};
  a.n_fresh_local = n; a.n_fresh_global = n; a.stride = p->stride;
  a.actions = d_act.p; a.costs = d_cost.p; a.mean = d_ms.p; a.std = d_ms.p + p->hd;
  a.prev_elites = d_act.p;
  a.start_state = reinterpret_cast<float*>(d_ss.p + sizeof(StepState));
  a.ss = reinterpret_cast<StepState*>(d_ss.p);
  launch_rollout_dyn<false, true>(p, a, n);
  ICEM_CUDA(cudaStreamSynchronize(p->stream));
  ICEM_CUDA(cudaMemcpy(costs_out, d_cost.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
  ICEM_API_END
}

int icem_op_rollout_observations(icem_planner_t* p, int32_t n, const double* state, int32_t state_dim,
                                 const float* actions, int32_t obs_dim, double* obs_out) {
  ICEM_API_BEGIN
  if (!p || !state || !actions || !obs_out) throw InvalidArg("null argument");
  if (n < 1 || n > 4096) throw InvalidArg("n must be in [1, 4096] (this operator is for elites, not populations)");
  require_model(p);
  if (state_dim != p->state_dim) throw InvalidArg("state_dim does not match the forward model");
  if (obs_dim < 1 || obs_dim > 512) throw InvalidArg("obs_dim out of range");
  ICEM_CUDA(cudaSetDevice(p->cfg.device));
  const int sd = p->state_dim, h = p->h, d = p->d;
  const size_t n_state = (size_t)2 * sd, n_act = (size_t)n * h * d, n_obs = (size_t)n * (h + 1) * obs_dim;
  p->step_scratch.reserve(n_state + n_act + n_obs);
  float* d_state = p->step_scratch.p;                  // start state, running state
  float* d_act = d_state + n_state;
  float* d_obs = d_act + n_act;
  std::vector<float> hs(sd);
  for (int i = 0; i < sd; ++i) hs[i] = (float)state[i];
  ICEM_CUDA(cudaMemcpyAsync(d_state, hs.data(), sd * sizeof(float), cudaMemcpyHostToDevice, p->stream));
  ICEM_CUDA(cudaMemcpyAsync(d_act, actions, n_act * sizeof(float), cudaMemcpyHostToDevice, p->stream));
  float* run = d_state + sd;
  for (int r = 0; r < n; ++r) {
    float* obs_r = d_obs + (size_t)r * (h + 1) * obs_dim;
    advance_dyn(p, d_state, nullptr, run, obs_r, obs_dim);                       // entry 0: the start observation
    for (int t = 0; t < h; ++t)
      advance_dyn(p, run, d_act + ((size_t)r * h + t) * d, run, obs_r + (size_t)(t + 1) * obs_dim, obs_dim);
  }
  std::vector<float> ho(n_obs);
  ICEM_CUDA(cudaMemcpyAsync(ho.data(), d_obs, ho.size() * sizeof(float), cudaMemcpyDeviceToHost, p->stream));
  ICEM_CUDA(cudaStreamSynchronize(p->stream));
  for (size_t i = 0; i < ho.size(); ++i) obs_out[i] = ho[i];
  ICEM_API_END
}

int icem_op_topk(icem_planner_t* p, int32_t n, const float* costs, int32_t k, int32_t* idx_out, float* costs_out) {
  ICEM_API_BEGIN
  if (!p || !costs || !idx_out || !costs_out) throw InvalidArg("null argument");
  if (n < 1 || k < 1 || k > 64 || k > n) throw InvalidArg("need 1 <= k <= min(n, 64)");
  ICEM_CUDA(cudaSetDevice(p->cfg.device));
  DevBuf<float> d_cost, d_co;
  DevBuf<int32_t> d_idx;
  DevBuf<unsigned long long> d_cand, d_send;
  DevBuf<unsigned int> d_ticket;
  DevBuf<unsigned char> d_ss;
  DevBuf<float> d_dummy;
  d_cost.alloc(n);
  ICEM_CUDA(cudaMemcpy(d_cost.p, costs, (size_t)n * 4, cudaMemcpyHostToDevice));
  d_co.alloc(k); d_idx.alloc(k);
  d_cand.alloc((size_t)p->sm_count * k);
  d_send.alloc(k);
  d_ticket.alloc(1);
  d_ss.alloc(sizeof(StepState));
  d_dummy.alloc(4);
  SelectArgs s{};
  s.n_fresh_local = n; s.n_fresh_global = n; s.k = k; s.stride = 0;
  s.costs = d_cost.p; s.actions = d_dummy.p; s.ss = reinterpret_cast<StepState*>(d_ss.p);
  s.cand = d_cand.p; s.ticket = d_ticket.p; s.send_keys = d_send.p; s.send_actions = d_dummy.p;
  RefitArgs r{};
  select_kernel<<<select_grid(p, n), kSelectThreads, select_smem(k), p->stream>>>(s, r, 0);
  ICEM_CUDA(cudaGetLastError());
  topk_finish_kernel<<<1, kSelectThreads, 0, p->stream>>>(d_send.p, k, d_cost.p, d_idx.p, d_co.p);
  ICEM_CUDA(cudaGetLastError());
  g_launches.fetch_add(2, std::memory_order_relaxed);
  ICEM_CUDA(cudaStreamSynchronize(p->stream));
  ICEM_CUDA(cudaMemcpy(idx_out, d_idx.p, k * 4, cudaMemcpyDeviceToHost));
  ICEM_CUDA(cudaMemcpy(costs_out, d_co.p, k * 4, cudaMemcpyDeviceToHost));
  ICEM_API_END
}

// ---- multi-GPU -------------------------------------------------------------------------------------
int icem_comm_get_unique_id(char id_out[ICEM_UNIQUE_ID_BYTES]) {
  ICEM_API_BEGIN
  if (!id_out) throw InvalidArg("null argument");
  Comm::unique_id(id_out);
  ICEM_API_END
}

int icem_comm_init(icem_planner_t* p, const char id[ICEM_UNIQUE_ID_BYTES]) {
  ICEM_API_BEGIN
  if (!p || !id) throw InvalidArg("null argument");
  if (p->cfg.world_size < 2) throw InvalidArg("planner was created with world_size 1");
  ICEM_CUDA(cudaSetDevice(p->cfg.device));
  p->comm.init(id, p->cfg.world_size, p->cfg.rank);
  ICEM_API_END
}

// ---- bench -----------------------------------------------------------------------------------------
int icem_bench_device(icem_planner_t* p, int32_t steps, int32_t warmup, int32_t flush_l2, float* total_ms,
                      float* rollout_ms, int32_t* rollout_launches) {
  ICEM_API_BEGIN
  if (p && p->B != 1) throw Unsupported("not available on a handle that batches several problems");
  if (!p) throw InvalidArg("null planner");
  if (!p->was_reset) throw StateError("beginning_of_rollout() needs to be called before");
  require_model(p);
  if (p->cfg.world_size > 1 && !p->comm.ready()) throw StateError("icem_comm_init not called");
  ICEM_CUDA(cudaSetDevice(p->cfg.device));
  const size_t flush_bytes = 256u << 20;   // > 126 MB L2
  if (flush_l2 && p->flush.n != flush_bytes) p->flush.alloc(flush_bytes);
  double tot = 0, roll = 0;
  int launches = 0;
  for (int s = 0; s < warmup + steps; ++s) {
    if (flush_l2) ICEM_CUDA(cudaMemsetAsync(p->flush.p, s & 0xff, flush_bytes, p->stream));
    ICEM_CUDA(cudaEventRecord(p->ev_a, p->stream));
    write_step_in(p, nullptr);
    ICEM_CUDA(cudaMemcpyAsync(p->step_in.p, p->h_in, sizeof(StepState), cudaMemcpyHostToDevice, p->stream));
    enqueue_iterations(p, true);
    advance_dyn(p, start_state_dev(p), p->out.p, start_state_dev(p), nullptr, 0);
    ICEM_CUDA(cudaEventRecord(p->ev_b, p->stream));
    ICEM_CUDA(cudaStreamSynchronize(p->stream));   // h_in is rewritten next step
    finish_step(p);
    if (s >= warmup) {
      float ms = 0;
      ICEM_CUDA(cudaEventElapsedTime(&ms, p->ev_a, p->ev_b));
      tot += ms;
      for (int i = 0; i < p->iters; ++i) {
        ICEM_CUDA(cudaEventElapsedTime(&ms, p->ev_roll[2 * i], p->ev_roll[2 * i + 1]));
        roll += ms;
        ++launches;
      }
    }
  }
  if (total_ms) *total_ms = (float)tot;
  if (rollout_ms) *rollout_ms = (float)roll;
  if (rollout_launches) *rollout_launches = launches;
  ICEM_API_END
}

int icem_bench_op(icem_planner_t* p, int32_t op, int32_t n, int32_t reps, int32_t flush_l2, float* ms_avg) {
  ICEM_API_BEGIN
  if (!p || !ms_avg) throw InvalidArg("null argument");
  if (reps < 1 || op < 0 || op > 4 || (op != 4 && n < p->k)) throw InvalidArg("bad n / reps / op");
  require_model(p);
  ICEM_CUDA(cudaSetDevice(p->cfg.device));
  if (op == 4) {
    // the exchange step of a sharded plan step alone: all-gather of the per-rank elite records + the deterministic
    // merge / refit every rank then runs (latency bound -- report microseconds, not GB/s)
    if (p->cfg.world_size < 2 || !p->comm.ready()) throw StateError("op 4 needs a sharded planner (icem_comm_init)");
    if (!p->was_reset) throw StateError("beginning_of_rollout() needs to be called before");
    write_step_in(p, nullptr);
    ICEM_CUDA(cudaMemcpyAsync(p->step_in.p, p->h_in, sizeof(StepState), cudaMemcpyHostToDevice, p->stream));
    enqueue_iterations(p, false);            // fills the send records with real elites
    const RefitArgs r0 = refit_args(p, 0);
    double tot4 = 0;
    for (int it = 0; it < reps + 3; ++it) {
      ICEM_CUDA(cudaEventRecord(p->ev_a, p->stream));
      p->comm.all_gather(p->send_rec.p, p->recv_rec.p, p->rec_bytes, p->stream);
      merge_refit_kernel<<<1, kSelectThreads, select_smem(p->k), p->stream>>>(r0);
      ICEM_CUDA(cudaGetLastError());
      g_launches.fetch_add(1, std::memory_order_relaxed);
      ICEM_CUDA(cudaEventRecord(p->ev_b, p->stream));
      ICEM_CUDA(cudaStreamSynchronize(p->stream));
      float ms = 0;
      ICEM_CUDA(cudaEventElapsedTime(&ms, p->ev_a, p->ev_b));
      if (it >= 3) tot4 += ms;
    }
    *ms_avg = (float)(tot4 / reps);
    reset_distribution(p);
    ICEM_CUDA(cudaStreamSynchronize(p->stream));
    return ICEM_OK;
  }
  DevBuf<float> d_act, d_cost;
  d_act.alloc((size_t)n * p->stride);
  d_cost.alloc(n);
  const size_t flush_bytes = 256u << 20;
  if (flush_l2 && p->flush.n != flush_bytes) p->flush.alloc(flush_bytes);
  reset_distribution(p);
  write_step_in(p, nullptr);
  ICEM_CUDA(cudaMemcpyAsync(p->step_in.p, p->h_in, sizeof(StepState), cudaMemcpyHostToDevice, p->stream));
  RolloutArgs a{};
  a.n_fresh_local = n; a.n_fresh_global = n; a.stride = p->stride;
  a.actions = d_act.p; a.costs = d_cost.p; a.mean = p->mean.p; a.std = p->stdv.p;
  a.prev_elites = p->elite_actions[0].p;
  a.start_state = start_state_dev(p);
  a.ss = step_state_dev(p);
  a.seed_lo = (uint32_t)p->cfg.seed; a.seed_hi = (uint32_t)(p->cfg.seed >> 32);
  SelectArgs s{};
  s.n_fresh_local = n; s.n_fresh_global = n; s.k = p->k; s.stride = p->stride;
  s.costs = d_cost.p; s.actions = d_act.p; s.ss = a.ss; s.cand = p->cand.p; s.ticket = p->ticket.p;
  RefitArgs r = refit_args(p, 0);
  r.n_keep = 0; r.n_fresh_global = n; r.local_actions = d_act.p; r.local_n_fresh = n; r.local_offset = 0;
  r.world = 1; r.rec_keys = nullptr; r.last_iteration = 0;
  if (op == 2 || op == 3) {   // costs / actions the timed kernel consumes
    launch_rollout_dyn<true, true>(p, a, n);
  }
  double tot = 0;
  for (int it = 0; it < reps + 3; ++it) {
    if (flush_l2) ICEM_CUDA(cudaMemsetAsync(p->flush.p, it & 0xff, flush_bytes, p->stream));
    ICEM_CUDA(cudaEventRecord(p->ev_a, p->stream));
    if (op == 0) launch_rollout_dyn<true, false>(p, a, n);
    else if (op == 1) launch_rollout_dyn<true, true>(p, a, n);
    else if (op == 3) launch_rollout_dyn<false, true>(p, a, n);
    else {
      select_kernel<<<select_grid(p, n), kSelectThreads, select_smem(p->k), p->stream>>>(s, r, 1);
      ICEM_CUDA(cudaGetLastError());
      g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    ICEM_CUDA(cudaEventRecord(p->ev_b, p->stream));
    ICEM_CUDA(cudaStreamSynchronize(p->stream));
    float ms = 0;
    ICEM_CUDA(cudaEventElapsedTime(&ms, p->ev_a, p->ev_b));
    if (it >= 3) tot += ms;
  }
  *ms_avg = (float)(tot / reps);
  reset_distribution(p);
  ICEM_CUDA(cudaStreamSynchronize(p->stream));
  ICEM_API_END
}

// ---- MLP trainer ------------------------------------------------------------------------------------------------
int icem_mlp_trainer_create(int32_t device, int32_t in_dim, int32_t hidden, int32_t out_dim, icem_mlp_trainer_t** out) {
  ICEM_API_BEGIN
  if (!out) throw InvalidArg("null argument");
  if (in_dim < 1 || hidden < 1 || out_dim < 1 || in_dim > 4096 || hidden > 4096 || out_dim > 4096)
    throw InvalidArg("layer widths must be in [1, 4096]");
  int count = 0;
  ICEM_CUDA(cudaGetDeviceCount(&count));
  if (device < 0 || device >= count) throw InvalidArg("no such CUDA device");
  cudaDeviceProp prop{};
  ICEM_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) throw Unsupported("icem_b200 needs an sm_100 device (no CPU or other-architecture fallback)");
  ICEM_CUDA(cudaSetDevice(device));
  std::unique_ptr<MlpTrainer> t(new MlpTrainer());
  t->device = device; t->in = in_dim; t->H = hidden; t->out = out_dim;
  ICEM_CUDA(cudaStreamCreateWithFlags(&t->stream, cudaStreamNonBlocking));
  const size_t n[6] = {(size_t)hidden * in_dim, (size_t)hidden, (size_t)hidden * hidden, (size_t)hidden,
                       (size_t)out_dim * hidden, (size_t)out_dim};
  for (int q = 0; q < 6; ++q) {
    t->par_n[q] = n[q];
    t->par[q].alloc(n[q]); t->mom[q].alloc(n[q]); t->var[q].alloc(n[q]);
  }
  t->losses.alloc(1);
  *out = reinterpret_cast<icem_mlp_trainer_t*>(t.release());
  ICEM_API_END
}

int icem_mlp_trainer_destroy(icem_mlp_trainer_t* h) {
  ICEM_API_BEGIN
  MlpTrainer* t = reinterpret_cast<MlpTrainer*>(h);
  if (t) {
    cudaSetDevice(t->device);
    if (t->stream) cudaStreamSynchronize(t->stream);
    delete t;
  }
  ICEM_API_END
}

int icem_mlp_trainer_set_weights(icem_mlp_trainer_t* h, const float* const* weights, const float* const* biases,
                                 int32_t reset_optimizer) {
  ICEM_API_BEGIN
  MlpTrainer* t = reinterpret_cast<MlpTrainer*>(h);
  if (!t || !weights || !biases) throw InvalidArg("null argument");
  ICEM_CUDA(cudaSetDevice(t->device));
  for (int l = 0; l < 3; ++l) {
    if (!weights[l] || !biases[l]) throw InvalidArg("null layer parameters");
    ICEM_CUDA(cudaMemcpyAsync(t->par[2 * l].p, weights[l], t->par_n[2 * l] * sizeof(float), cudaMemcpyHostToDevice, t->stream));
    ICEM_CUDA(cudaMemcpyAsync(t->par[2 * l + 1].p, biases[l], t->par_n[2 * l + 1] * sizeof(float), cudaMemcpyHostToDevice, t->stream));
  }
  if (reset_optimizer) {
    for (int q = 0; q < 6; ++q) {
      ICEM_CUDA(cudaMemsetAsync(t->mom[q].p, 0, t->par_n[q] * sizeof(float), t->stream));
      ICEM_CUDA(cudaMemsetAsync(t->var[q].p, 0, t->par_n[q] * sizeof(float), t->stream));
    }
    t->adam_t = 0;
  }
  ICEM_CUDA(cudaStreamSynchronize(t->stream));
  ICEM_API_END
}

int icem_mlp_trainer_get_weights(icem_mlp_trainer_t* h, float* const* weights, float* const* biases) {
  ICEM_API_BEGIN
  MlpTrainer* t = reinterpret_cast<MlpTrainer*>(h);
  if (!t || !weights || !biases) throw InvalidArg("null argument");
  ICEM_CUDA(cudaSetDevice(t->device));
  for (int l = 0; l < 3; ++l) {
    if (!weights[l] || !biases[l]) throw InvalidArg("null layer parameters");
    ICEM_CUDA(cudaMemcpyAsync(weights[l], t->par[2 * l].p, t->par_n[2 * l] * sizeof(float), cudaMemcpyDeviceToHost, t->stream));
    ICEM_CUDA(cudaMemcpyAsync(biases[l], t->par[2 * l + 1].p, t->par_n[2 * l + 1] * sizeof(float), cudaMemcpyDeviceToHost, t->stream));
  }
  ICEM_CUDA(cudaStreamSynchronize(t->stream));
  ICEM_API_END
}

int icem_mlp_trainer_set_data(icem_mlp_trainer_t* h, int64_t n, const float* inputs, const float* targets) {
  ICEM_API_BEGIN
  MlpTrainer* t = reinterpret_cast<MlpTrainer*>(h);
  if (!t || !inputs || !targets) throw InvalidArg("null argument");
  if (n < 1 || n > 0x7fffffffLL) throw InvalidArg("number of transitions out of range");
  ICEM_CUDA(cudaSetDevice(t->device));
  t->x.reserve((size_t)n * t->in);
  t->t.reserve((size_t)n * t->out);
  ICEM_CUDA(cudaMemcpyAsync(t->x.p, inputs, (size_t)n * t->in * sizeof(float), cudaMemcpyHostToDevice, t->stream));
  ICEM_CUDA(cudaMemcpyAsync(t->t.p, targets, (size_t)n * t->out * sizeof(float), cudaMemcpyHostToDevice, t->stream));
  ICEM_CUDA(cudaStreamSynchronize(t->stream));
  t->n_rows = (size_t)n;
  ICEM_API_END
}

int icem_mlp_trainer_fit(icem_mlp_trainer_t* h, int32_t n_steps, int32_t batch, const int32_t* indices, float lr,
                         float beta1, float beta2, float eps, float weight_decay, float* losses_out) {
  ICEM_API_BEGIN
  MlpTrainer* t = reinterpret_cast<MlpTrainer*>(h);
  if (!t || !indices) throw InvalidArg("null argument");
  if (t->n_rows == 0) throw StateError("icem_mlp_trainer_set_data() needs to be called before");
  if (n_steps < 1 || batch < 1) throw InvalidArg("n_steps and batch must be positive");
  if (!(lr > 0.f) || !(beta1 >= 0.f && beta1 < 1.f) || !(beta2 >= 0.f && beta2 < 1.f) || !(eps > 0.f))
    throw InvalidArg("Adam hyper-parameters out of range");
  const size_t total = (size_t)n_steps * batch;
  for (size_t i = 0; i < total; ++i)
    if (indices[i] < 0 || (size_t)indices[i] >= t->n_rows) throw InvalidArg("minibatch index outside the data set");
  ICEM_CUDA(cudaSetDevice(t->device));
  trainer_reserve_batch(t, batch);
  t->idx.reserve(total);
  t->losses.reserve((size_t)n_steps);
  ICEM_CUDA(cudaMemcpyAsync(t->idx.p, indices, total * sizeof(int32_t), cudaMemcpyHostToDevice, t->stream));
  for (int s = 0; s < n_steps; ++s)
    trainer_step(t, t->idx.p + (size_t)s * batch, batch, lr, beta1, beta2, eps, weight_decay, t->losses.p + s);
  if (losses_out)
    ICEM_CUDA(cudaMemcpyAsync(losses_out, t->losses.p, (size_t)n_steps * sizeof(float), cudaMemcpyDeviceToHost, t->stream));
  ICEM_CUDA(cudaStreamSynchronize(t->stream));
  ICEM_API_END
}

int icem_mlp_trainer_predict(icem_mlp_trainer_t* h, int32_t n, const float* inputs, float* outputs) {
  ICEM_API_BEGIN
  MlpTrainer* t = reinterpret_cast<MlpTrainer*>(h);
  if (!t || !inputs || !outputs) throw InvalidArg("null argument");
  if (n < 1) throw InvalidArg("n must be positive");
  ICEM_CUDA(cudaSetDevice(t->device));
  trainer_reserve_batch(t, n);
  ICEM_CUDA(cudaMemcpyAsync(t->xb.p, inputs, (size_t)n * t->in * sizeof(float), cudaMemcpyHostToDevice, t->stream));
  train_gemm<false, true, kTrEpBiasTanh>(t, n, t->H, t->in, t->xb.p, t->in, t->par[0].p, t->in, t->a1.p, t->H, t->par[1].p, nullptr, 0);
  train_gemm<false, true, kTrEpBiasTanh>(t, n, t->H, t->H, t->a1.p, t->H, t->par[2].p, t->H, t->a2.p, t->H, t->par[3].p, nullptr, 0);
  train_gemm<false, true, kTrEpBias>(t, n, t->out, t->H, t->a2.p, t->H, t->par[4].p, t->H, t->y.p, t->out, t->par[5].p, nullptr, 0);
  ICEM_CUDA(cudaMemcpyAsync(outputs, t->y.p, (size_t)n * t->out * sizeof(float), cudaMemcpyDeviceToHost, t->stream));
  ICEM_CUDA(cudaStreamSynchronize(t->stream));
  ICEM_API_END
}

#ifdef ICEM_MLP_TRACE
int icem_debug_mlp_trace(long long* out, int n) {
  return (int)cudaMemcpyFromSymbol(out, icem::g_mlp_trace, sizeof(long long) * (size_t)n);
}
#endif

}  // extern "C"
