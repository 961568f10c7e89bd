// TEST INFRASTRUCTURE ONLY (never loaded by the product package).
//
// Compiles icem_b200/csrc/dyn_chain.cuh -- the source the CUDA kernels run -- for the HOST, one thread per lane of a
// group, so that the branch-parallel articulated-body engine can be compared with the float64 oracle
// (oracle/articulated_np.py) in the CPU test suite.  The group sum (warp shuffles on the GPU) is a barrier + sum here.
#include <atomic>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>
#include <sched.h>

#include "../../icem_b200/csrc/dyn_chain.cuh"

namespace {

struct SpinBarrier {
  explicit SpinBarrier(int n) : n_(n), count_(0), sense_(0) {}
  void wait() {
    const int s = sense_.load(std::memory_order_acquire);
    if (count_.fetch_add(1, std::memory_order_acq_rel) == n_ - 1) {
      count_.store(0, std::memory_order_relaxed);
      sense_.store(s ^ 1, std::memory_order_release);
    } else {
      int spins = 0;
      while (sense_.load(std::memory_order_acquire) == s)
        if (++spins > 64) sched_yield();
    }
  }
  int n_;
  std::atomic<int> count_, sense_;
};

struct HostCtx {
  SpinBarrier* bar;
  float* xbuf;     // [G][32]
  int g, G;
  template <int N>
  void group_sum(float (&x)[N]) {
    const int n = N;
    memcpy(xbuf + g * 32, x, n * sizeof(float));
    bar->wait();
    for (int e = 0; e < n; ++e) {
      // butterfly order of the GPU: (x0 + x1) + (x2 + x3)
      float s = 0.f;
      if (G == 1) s = xbuf[e];
      else if (G == 2) s = xbuf[e] + xbuf[32 + e];
      else s = (xbuf[e] + xbuf[32 + e]) + (xbuf[64 + e] + xbuf[96 + e]);
      x[e] = s;
    }
    bar->wait();
  }
  void group_sync() { bar->wait(); }
};

}  // namespace

// states_out[n][h + 1][nq + nv]: the state before every action and after the last one (all h steps are simulated).
// planar: 0 = spatial engine, 1 = planar engine (fails when the model is not planar), -1 = what the product picks.
template <bool PLANAR>
static void run_lanes(const icem::ChainModel& m, int act_dim, int n, int h, const double* start, const float* actions,
                      double* states_out) {
  const int G = m.lanes, ns = m.nq + m.nv;
  std::vector<float> shared(m.s_end + 8, 0.f), xbuf(4 * 32, 0.f);
  std::vector<std::vector<float>> priv(G, std::vector<float>(m.p_end + 8, 0.f));
  SpinBarrier bar(G);
  auto lane = [&](int g) {
    HostCtx ctx{&bar, xbuf.data(), g, G};
    icem::ChainLane<HostCtx, 1, PLANAR> L;
    L.M = &m; L.sh = shared.data(); L.pr = priv[g].data(); L.g = g; L.ctx = &ctx;
    const auto ctrl = L.shared_rec(m.s_ctrl);
    for (int r = 0; r < n; ++r) {
      if (g == 0)
        for (int i = 0; i < ns; ++i) shared[m.s_state + i] = (float)start[i];
      bar.wait();
      for (int t = 0; t <= h; ++t) {
        if (g == 0) {
          for (int i = 0; i < ns; ++i) states_out[((size_t)r * (h + 1) + t) * ns + i] = shared[m.s_state + i];
          if (t < h)
            for (int k = 0; k < act_dim; ++k) shared[m.s_ctrl + k] = actions[((size_t)r * h + t) * act_dim + k];
        }
        bar.wait();
        if (t < h) L.step(ctrl);
        bar.wait();
      }
    }
  };
  std::vector<std::thread> th;
  for (int g = 1; g < G; ++g) th.emplace_back(lane, g);
  lane(0);
  for (auto& t : th) t.join();
}

extern "C" {

// decomposition report: lanes, trunk nodes, limbs, scratch floats per warp; returns 0 when eligible
int chain_host_describe(const icem_articulated_model_t* a, int act_dim, int* out /* [8] */, char* why, int why_len) {
  icem::ChainModel m;
  const char* w = "";
  const bool ok = icem::build_chain_model(icem::chain_source(*a), act_dim, m, &w);
  snprintf(why, why_len, "%s", w);
  if (!ok) return 1;
  out[0] = m.lanes; out[1] = m.n_trunk; out[2] = m.n_limbs; out[3] = m.n_nodes; out[4] = m.trunk_dofs;
  out[5] = m.max_limb_dofs; out[6] = icem::chain_warp_floats(m); out[7] = (int)sizeof(icem::ChainModel);
  return 0;
}

int chain_host_rollout_engine(const icem_articulated_model_t* a, int act_dim, int integrator, int planar, int n, int h,
                              const double* start, const float* actions, double* states_out) {
  icem::ChainModel m;
  const char* w = "";
  icem::ChainSource src = icem::chain_source(*a);
  src.integrator = integrator;
  if (!icem::build_chain_model(src, act_dim, m, &w)) return 1;
  if (planar == 1 && !m.planar) return 2;
  if (planar < 0) planar = m.planar;
  if (planar) run_lanes<true>(m, act_dim, n, h, start, actions, states_out);
  else run_lanes<false>(m, act_dim, n, h, start, actions, states_out);
  return 0;
}

int chain_host_rollout(const icem_articulated_model_t* a, int act_dim, int integrator, int n, int h,
                       const double* start, const float* actions, double* states_out) {
  return chain_host_rollout_engine(a, act_dim, integrator, -1, n, h, start, actions, states_out);
}

int chain_host_is_planar(const icem_articulated_model_t* a, int act_dim) {
  icem::ChainModel m;
  const char* w = "";
  if (!icem::build_chain_model(icem::chain_source(*a), act_dim, m, &w)) return -1;
  return m.planar;
}

}  // extern "C"
