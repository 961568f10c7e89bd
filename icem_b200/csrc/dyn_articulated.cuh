// placeholder -- replaced by the articulated rigid-body engine
#pragma once
#include "common.cuh"
namespace icem {
struct Articulated {
  static constexpr int kWarpsPerCta = 8;
  struct Params { int act_dim; int nq, nv; };
  __host__ __device__ static int cta_floats(const Params&) { return 0; }
  __host__ __device__ static int warp_floats(const Params&) { return 0; }
  __host__ __device__ static int state_dim(const Params& p) { return p.nq + p.nv; }
  __device__ static void cta_init(const Params&, float*) {}
  __device__ void bind(const Params&, const float*, float*) {}
  __device__ void reset(const float*) {}
  __device__ float obs(int) const { return 0.f; }
  __device__ void step(const float*) {}
  __device__ void export_state(float*) const {}
};
inline void articulated_setup(int, int, Articulated::Params*) {
  throw std::runtime_error("articulated dynamics not built yet");
}
}  // namespace icem
