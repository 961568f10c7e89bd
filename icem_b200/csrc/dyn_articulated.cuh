// Articulated rigid-body forward model, one warp per trajectory (ground-truth dynamics of the B200 path).
//
// Stands in for what the reference reaches through `GroundTruthModel.predict_n_steps` -> `env.step` -> MuJoCo
// (icem/models/gt_model.py:76-102, icem/environments/mujoco.py:101-131).  MuJoCo is not available;
// this is the repo's own soft-contact engine on the tables of icem_b200/robots.py, checked against the
// independent float64 restatement oracle/articulated_np.py.  PARITY WITH MUJOCO IS UNPINNED.
//
// Per substep (frame_skip substeps per control step, control held), all in one warp, data exchanged through a
// per-warp shared-memory scratch and warp shuffles:
//   1a. lane = body : transform of every body relative to its parent through its own joints (all bodies at once)
//   1b. lane = body : compose parent x local down the tree, one tree level at a time (positions relative to the
//                     root origin O so that fp32 keeps its precision far from the world origin)
//   1c. lane = dof  : motion axes S_j (Plucker [w; v_O]) in world coordinates
//   1d. lane = dof  : velocities and velocity-product accelerations as two pointer-doubling prefix sums along the
//                     dof chains (log2(chain length) rounds instead of one round per tree level)
//   2.  lane = contact sphere : floor contact force (Hunt-Crossley normal, capped viscous Coulomb friction) -> wrench
//   3.  lane = body : spatial inertia about O and Newton-Euler force f_b = I a + v x* I v - f_ext;
//       lane = (parent, component): leaf -> root accumulation of forces and composite inertias
//   4.  lane = dof  : bias_j = S_j . f, applied torques (gear, spring, limit spring-damper), row j of the
//                     composite-rigid-body mass matrix, kept in REGISTERS
//   5.  (M + dt B + dt^2 K) qacc = rhs: in-register Cholesky (column broadcasts by shuffle), forward substitution by
//       shuffles, back substitution on the transposed factor (transposed through shared memory)
//   6.  semi-implicit Euler; unit-quaternion update for a free root joint
#pragma once
#include "common.cuh"

namespace icem {

constexpr int kArtMaxBodies = 16;
constexpr int kArtMaxDofs = 32;
constexpr int kArtMaxContacts = 32;
constexpr int kArtMaxChildren = 4;
constexpr int kArtMaxDepth = 8;
constexpr int kArtMaxLevelParents = 8;
constexpr int kArtScanRounds = 5;
#ifndef ICEM_ART_JOINT_TYPES
#define ICEM_ART_JOINT_TYPES
enum { kSlide = 0, kHinge = 1, kFreeTrans = 2, kFreeRot = 3 };
#endif

// POD model tables; built on the host (planner.cu: icem_set_articulated_model), copied to shared memory per CTA.
struct ArtModel {
  int nb, nq, nv, nu, nc, nsub, max_depth, obs_offset;
  float dt, gravity, ctrl_limit, kc, cc, kv, mu, cdmax;
  int scan_rounds, pad1, pad2, pad3;
  int b_parent[kArtMaxBodies], b_depth[kArtMaxBodies], b_dof_start[kArtMaxBodies], b_dof_count[kArtMaxBodies];
  int b_nchild[kArtMaxBodies], b_child[kArtMaxBodies][kArtMaxChildren];
  int b_con_start[kArtMaxBodies], b_con_count[kArtMaxBodies];
  int b_last_dof[kArtMaxBodies];      // last dof on the path root -> body (-1: none)
  int lvl_np[kArtMaxDepth], lvl_parent[kArtMaxDepth][kArtMaxLevelParents];   // bodies with children, per depth
  int lvl_child[kArtMaxDepth][kArtMaxLevelParents][kArtMaxChildren];         // their children, padded with -1
  float b_pos[kArtMaxBodies][3], b_mass[kArtMaxBodies], b_com[kArtMaxBodies][3], b_inertia[kArtMaxBodies][6];
  int d_body[kArtMaxDofs], d_type[kArtMaxDofs], d_qadr[kArtMaxDofs], d_limited[kArtMaxDofs], d_act[kArtMaxDofs];
  int d_vref[kArtMaxDofs];            // dof whose inclusive velocity sum is the velocity of the frame S_j is fixed in
  int d_jump[kArtScanRounds][kArtMaxDofs];   // 2^r-th ancestor dof (-1: none): pointer-doubling prefix sums
  unsigned d_chain[kArtMaxDofs];      // bit c set: dof c is dof j or one of its ancestors
  float d_axis[kArtMaxDofs][3], d_anchor[kArtMaxDofs][3];
  float d_stiff[kArtMaxDofs], d_damp[kArtMaxDofs], d_arm[kArtMaxDofs], d_lo[kArtMaxDofs], d_hi[kArtMaxDofs];
  float d_klim[kArtMaxDofs], d_blim[kArtMaxDofs], d_gear[kArtMaxDofs];
  int c_body[kArtMaxContacts];
  float c_pos[kArtMaxContacts][3], c_radius[kArtMaxContacts];
};

// ---- small vector helpers (everything stays in registers) ----------------------------------------------------
__device__ __forceinline__ void cross3(const float* a, const float* b, float* o) {
  const float x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z;
}
__device__ __forceinline__ void matvec3(const float* R, const float* v, float* o) {   // R row-major 3x3
  const float x = R[0] * v[0] + R[1] * v[1] + R[2] * v[2];
  const float y = R[3] * v[0] + R[4] * v[1] + R[5] * v[2];
  const float z = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
__device__ __forceinline__ void matmul3(const float* A, const float* B, float* C) {
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) C[r * 3 + c] = A[r * 3] * B[c] + A[r * 3 + 1] * B[3 + c] + A[r * 3 + 2] * B[6 + c];
}

// padded strides of the per-warp arrays (odd strides: conflict-free when lane = row)
constexpr int kSR = 9, kSP = 3, kS6 = 7, kSRec = 17;

template <int NVMAX>
struct Articulated {
  static constexpr int kWarpsPerCta = 8;
#ifndef ICEM_ART_MIN_CTAS
#define ICEM_ART_MIN_CTAS 3
#endif
#ifndef ICEM_ART_LOCKSTEP
#define ICEM_ART_LOCKSTEP 1
#endif
  static constexpr int kMinCtasPerSm = NVMAX <= 24 ? ICEM_ART_MIN_CTAS : 2;   // register cap: 85 / 128 per thread
  static constexpr bool kCtaLockstep = ICEM_ART_LOCKSTEP != 0;
  static constexpr bool kOutlineRollout = true;
  static constexpr bool kHasHealth = true;      // state_healthy(): usable with ICEM_COST_LOCOMOTION
  static constexpr int kLd = NVMAX + 1;           // row stride of the transposition buffer (odd for NVMAX even)
  struct Params {
    const ArtModel* model;   // device global memory
    int act_dim, nq, nv, nb, nc;
  };
  // per-warp scratch layout (floats): compile-time offsets for this size class (NBMAX bodies, NVMAX dofs, NCMAX
  // contact spheres) -- runtime-computed offsets cost ~7 % in address arithmetic
  static constexpr int NBMAX = NVMAX <= 12 ? 8 : kArtMaxBodies;
  static constexpr int NCMAX = NVMAX <= 12 ? 16 : kArtMaxContacts;
  static constexpr int oO = 64;                               // [0,64): qpos, qvel
  static constexpr int oRb = oO + 4;                          // [nb][9]
  static constexpr int oPb = oRb + NBMAX * kSR;               // [nb][3]
  static constexpr int oRec = oPb + NBMAX * kSP;              // [nb][17]  f(6) + composite inertia(10)
  static constexpr int kZeroRec = NBMAX;                      // an all-zero record: padding child in 3b
  static constexpr int oSd = oRec + (NBMAX + 1) * kSRec;      // [nv][7]   S_j
  static constexpr int oX = (oSd + NVMAX * kS6 + 3) & ~3;     // [nv][7]   prefix sums of S qd       (16-B aligned)
  static constexpr int oY = (oX + NVMAX * kS6 + 3) & ~3;      // [nv][7]   prefix sums of (v x S) qd (16-B aligned)
  static constexpr int oCw = oY + NVMAX * kS6;                // [nc][7]   contact wrenches
  static constexpr int oEnd0 = oCw + NCMAX * kS6;
  // the transposition buffer of the Cholesky factor aliases Rb.. (dead by then)
  static constexpr int oEnd = oEnd0 > oRb + NVMAX * kLd ? oEnd0 : oRb + NVMAX * kLd;
  static_assert(oX % 4 == 0 && oY % 4 == 0 && NVMAX % 4 == 0, "Cholesky column buffers must be 16-byte aligned");
  __host__ __device__ static bool fits(int nb, int nv, int nc) { return nb <= NBMAX && nv <= NVMAX && nc <= NCMAX; }
  static constexpr int oState = 0;
  __host__ __device__ static int cta_floats(const Params&) { return (int)((sizeof(ArtModel) + 3) / 4); }
  __host__ __device__ static int warp_floats(const Params&) { return oEnd; }
  __host__ __device__ static int state_dim(const Params& p) { return p.nq + p.nv; }

  __device__ static void cta_init(const Params& p, float* s) {
    const int n = (int)(sizeof(ArtModel) / 4);
    const float* src = reinterpret_cast<const float*>(p.model);
    for (int i = threadIdx.x; i < n; i += blockDim.x) s[i] = src[i];
  }

  const ArtModel* M;
  float* w;     // per-warp scratch

  __device__ void bind(const Params&, const float* cta, float* warp) {
    M = reinterpret_cast<const ArtModel*>(cta);
    w = warp;
  }
  __device__ void reset(const float* start_state) {
    const int n = M->nq + M->nv;
    for (int i = lane_id(); i < n; i += 32) w[oState + i] = start_state[i];
    __syncwarp();
  }
  __device__ float obs(int i) const { return w[oState + i + M->obs_offset]; }
  // warp-uniform: every state entry finite, and |obs[i]| < bound for the observation entries i >= first (bound <= 0:
  // no bound).  The only user is Hopper (mujoco.py:199-203), whose observation carries the velocities clipped to
  // +-10 (gym hopper_v3._get_obs), so a velocity entry counts as min(|v|, 10).
  __device__ bool state_healthy(int first, float bound) const {
    const int nq = M->nq, n = nq + M->nv, lane = lane_id();
    bool ok = true;
    for (int i = lane; i < n; i += 32) {
      const float v = w[oState + i];
      ok = ok && (v - v == 0.f);
      if (bound > 0.f && i >= first + M->obs_offset) ok = ok && (i < nq ? fabsf(v) : fminf(fabsf(v), 10.f)) < bound;
    }
    return __all_sync(0xffffffffu, ok);
  }
  __device__ void export_state(float* out) const {
    const int n = M->nq + M->nv;
    for (int i = lane_id(); i < n; i += 32) out[i] = w[oState + i];
  }
  __device__ void step(const float* ctrl) {
    for (int s = 0; s < M->nsub; ++s) substep(ctrl);
  }

  // inclusive prefix sum of 6-vectors along the dof chains: X[j] <- sum_{i in chain(j)} X[i]   (lane = dof)
  __device__ __forceinline__ void chain_scan(float* X, float* x, int j, bool is_dof) {
    const ArtModel& m = *M;
    for (int r = 0; r < m.scan_rounds; ++r) {
      float t[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      const int src = is_dof ? m.d_jump[r][j] : -1;
      if (src >= 0) {
#pragma unroll
        for (int i = 0; i < 6; ++i) t[i] = X[src * kS6 + i];
      }
      __syncwarp();
      if (src >= 0) {
#pragma unroll
        for (int i = 0; i < 6; ++i) { x[i] += t[i]; X[j * kS6 + i] = x[i]; }
      }
      __syncwarp();
    }
  }

  // ------------------------------------------------------------------------------------------------------------
  __device__ void substep(const float* ctrl) {
    const ArtModel& m = *M;
    const int lane = lane_id();
    const float* q = w + oState;
    const float* qd = w + oState + m.nq;
    float* O = w + oO;
    float* Rb = w + oRb;
    float* pb = w + oPb;
    float* rec = w + oRec;
    float* Sd = w + oSd;
    float* X = w + oX;
    float* Y = w + oY;
    float* cw = w + oCw;
    const float dt = m.dt;
    const bool is_body = lane < m.nb;
    const bool is_dof = lane < m.nv;
    const int depth = is_body ? m.b_depth[lane] : 1 << 20;

    // ---- 1a. lane = body: transform relative to the parent through the body's own joints -----------------------
    float R[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f}, p[3] = {0.f, 0.f, 0.f};
    if (is_body) {
      const int b = lane;
#pragma unroll
      for (int i = 0; i < 3; ++i) p[i] = m.b_pos[b][i];
      const int j0 = m.b_dof_start[b], j1 = j0 + m.b_dof_count[b];
      for (int j = j0; j < j1; ++j) {
        const int t = m.d_type[j];
        if (t == kFreeTrans) {                 // free joint (root): 3 translations + ball, S directly in world
          const int qa = m.d_qadr[j];
          p[0] = q[qa]; p[1] = q[qa + 1]; p[2] = q[qa + 2];
          float qw = q[qa + 3], x = q[qa + 4], y = q[qa + 5], z = q[qa + 6];
          {   // rotation of the NORMALISED quaternion (a start state may carry reset noise on it)
            const float qn = rsqrtf(qw * qw + x * x + y * y + z * z);
            qw *= qn; x *= qn; y *= qn; z *= qn;
          }
          R[0] = 1.f - 2.f * (y * y + z * z); R[1] = 2.f * (x * y - qw * z); R[2] = 2.f * (x * z + qw * y);
          R[3] = 2.f * (x * y + qw * z); R[4] = 1.f - 2.f * (x * x + z * z); R[5] = 2.f * (y * z - qw * x);
          R[6] = 2.f * (x * z - qw * y); R[7] = 2.f * (y * z + qw * x); R[8] = 1.f - 2.f * (x * x + y * y);
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            float* St = Sd + (j + k) * kS6;
            float* Sr = Sd + (j + 3 + k) * kS6;
#pragma unroll
            for (int i = 0; i < 6; ++i) { St[i] = 0.f; Sr[i] = 0.f; }
            St[3 + k] = 1.f;
            Sr[0] = R[k]; Sr[1] = R[3 + k]; Sr[2] = R[6 + k];   // column k of R, anchored at O (= p)
          }
          j += 5;
          continue;
        }
        float ax[3];
        matvec3(R, m.d_axis[j], ax);
        const float qj = q[m.d_qadr[j]];
        float* loc = Sd + j * kS6;             // local record (axis, anchor) in the parent frame, finished in 1c
        loc[0] = ax[0]; loc[1] = ax[1]; loc[2] = ax[2];
        if (t == kSlide) {
          loc[3] = loc[4] = loc[5] = 0.f;
          p[0] += ax[0] * qj; p[1] += ax[1] * qj; p[2] += ax[2] * qj;
        } else {
          float an[3];
          matvec3(R, m.d_anchor[j], an);
          an[0] += p[0]; an[1] += p[1]; an[2] += p[2];
          loc[3] = an[0]; loc[4] = an[1]; loc[5] = an[2];
          float sn, cs;
          sincosf(qj, &sn, &cs);
          const float C = 1.f - cs;
          const float Rj[9] = {cs + ax[0] * ax[0] * C, ax[0] * ax[1] * C - ax[2] * sn, ax[0] * ax[2] * C + ax[1] * sn,
                               ax[1] * ax[0] * C + ax[2] * sn, cs + ax[1] * ax[1] * C, ax[1] * ax[2] * C - ax[0] * sn,
                               ax[2] * ax[0] * C - ax[1] * sn, ax[2] * ax[1] * C + ax[0] * sn, cs + ax[2] * ax[2] * C};
          float Rn[9];
          matmul3(Rj, R, Rn);
#pragma unroll
          for (int i = 0; i < 9; ++i) R[i] = Rn[i];
          float dp[3] = {p[0] - an[0], p[1] - an[1], p[2] - an[2]}, rp[3];
          matvec3(Rj, dp, rp);
          p[0] = an[0] + rp[0]; p[1] = an[1] + rp[1]; p[2] = an[2] + rp[2];
        }
      }
      // world frame; everything downstream is relative to O = the origin of the FIRST root's frame (body 0; bodies are
      // listed parent-before-child, so body 0 is a root).  A robot may have further roots (Reacher's target body):
      // they are placed relative to the same O below, after the warp has seen it.
      if (b == 0) { O[0] = p[0]; O[1] = p[1]; O[2] = p[2]; }
    }
    __syncwarp();
    if (depth == 0) {
      const int b = lane;
      p[0] -= O[0]; p[1] -= O[1]; p[2] -= O[2];
#pragma unroll
      for (int i = 0; i < 9; ++i) Rb[b * kSR + i] = R[i];
#pragma unroll
      for (int i = 0; i < 3; ++i) pb[b * kSP + i] = p[i];
    }
    __syncwarp();

    // ---- 1b. compose down the tree, one level at a time ----------------------------------------------------------
    for (int L = 1; L <= m.max_depth; ++L) {
      if (depth == L) {
        const int b = lane, par = m.b_parent[b];
        float Rp[9], Rn[9], off[3];
#pragma unroll
        for (int i = 0; i < 9; ++i) Rp[i] = Rb[par * kSR + i];
        matmul3(Rp, R, Rn);
        matvec3(Rp, p, off);
#pragma unroll
        for (int i = 0; i < 9; ++i) { R[i] = Rn[i]; Rb[b * kSR + i] = Rn[i]; }
#pragma unroll
        for (int i = 0; i < 3; ++i) { p[i] = pb[par * kSP + i] + off[i]; pb[b * kSP + i] = p[i]; }
      }
      __syncwarp();
    }

    // ---- 1c. lane = dof: motion axis in world coordinates ---------------------------------------------------------
    float S[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float qdj = 0.f;
    int dtype = -1;
    if (is_dof) {
      const int j = lane;
      dtype = m.d_type[j];
      qdj = qd[j];
      if (dtype == kSlide || dtype == kHinge) {
        const int par = m.b_parent[m.d_body[j]];
        float ax[3] = {Sd[j * kS6], Sd[j * kS6 + 1], Sd[j * kS6 + 2]};
        float an[3] = {Sd[j * kS6 + 3], Sd[j * kS6 + 4], Sd[j * kS6 + 5]};
        if (par >= 0) {
          float axw[3], anw[3];
          matvec3(Rb + par * kSR, ax, axw);
          matvec3(Rb + par * kSR, an, anw);
#pragma unroll
          for (int i = 0; i < 3; ++i) { ax[i] = axw[i]; an[i] = anw[i] + pb[par * kSP + i]; }
        } else {
#pragma unroll
          for (int i = 0; i < 3; ++i) an[i] -= O[i];
        }
        if (dtype == kSlide) {
          S[3] = ax[0]; S[4] = ax[1]; S[5] = ax[2];
        } else {
          S[0] = ax[0]; S[1] = ax[1]; S[2] = ax[2];
          cross3(an, ax, S + 3);               // velocity at O of a rotation about the anchored axis
        }
      } else {
#pragma unroll
        for (int i = 0; i < 6; ++i) S[i] = Sd[j * kS6 + i];
      }
    }
    __syncwarp();
    // ---- 1d. velocities and velocity-product accelerations: prefix sums along the dof chains ---------------------
    float xv[6], ya[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) xv[i] = S[i] * qdj;
    if (is_dof) {
#pragma unroll
      for (int i = 0; i < 6; ++i) { Sd[lane * kS6 + i] = S[i]; X[lane * kS6 + i] = xv[i]; }
    }
    __syncwarp();
    chain_scan(X, xv, lane, is_dof);           // X[j] = velocity of the body right after joint j
    {
      float vf[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      const int ref = is_dof ? m.d_vref[lane] : -1;
      if (ref >= 0) {
#pragma unroll
        for (int i = 0; i < 6; ++i) vf[i] = X[ref * kS6 + i];
      }
      // (v xm S) qd      [w1;v1] xm [w2;v2] = [w1 x w2 ; w1 x v2 + v1 x w2]
      float c1[3], c2[3], c3[3];
      cross3(vf, S, c1);
      cross3(vf, S + 3, c2);
      cross3(vf + 3, S, c3);
      ya[0] = c1[0] * qdj; ya[1] = c1[1] * qdj; ya[2] = c1[2] * qdj;
      ya[3] = (c2[0] + c3[0]) * qdj; ya[4] = (c2[1] + c3[1]) * qdj; ya[5] = (c2[2] + c3[2]) * qdj;
      if (is_dof) {
#pragma unroll
        for (int i = 0; i < 6; ++i) Y[lane * kS6 + i] = ya[i];
      }
    }
    __syncwarp();
    chain_scan(Y, ya, lane, is_dof);           // Y[j] = velocity-product acceleration of the body after joint j

    // ---- 2. floor contacts: lane = contact sphere ---------------------------------------------------------------
    if (lane < m.nc) {
      const int b = m.c_body[lane];
      float x[3];
      matvec3(Rb + b * kSR, m.c_pos[lane], x);
      x[0] += pb[b * kSP]; x[1] += pb[b * kSP + 1]; x[2] += pb[b * kSP + 2];
      const float pen = m.c_radius[lane] - (O[2] + x[2]);
      float wr[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (pen > 0.f) {
        const int e = m.b_last_dof[b];
        float vv[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (e >= 0) {
#pragma unroll
          for (int i = 0; i < 6; ++i) vv[i] = X[e * kS6 + i];
        }
        float u[3];
        cross3(vv, x, u);
        u[0] += vv[3]; u[1] += vv[4]; u[2] += vv[5];
        const float spring = m.kc * pen;
        const float damp = fminf(spring * m.cc, m.cdmax);
        const float fn = fminf(fmaxf(spring - damp * u[2], 0.f), 3.f * spring);
        const float speed = sqrtf(u[0] * u[0] + u[1] * u[1]);
        const float coef = fminf(m.kv, m.mu * fn / fmaxf(speed, 1e-6f));
        const float f[3] = {-coef * u[0], -coef * u[1], fn};
        cross3(x, f, wr);
        wr[3] = f[0]; wr[4] = f[1]; wr[5] = f[2];
      }
#pragma unroll
      for (int i = 0; i < 6; ++i) cw[lane * kS6 + i] = wr[i];
    }
    __syncwarp();

    // ---- 3. spatial inertia about O, Newton-Euler force: lane = body ---------------------------------------------
    if (is_body) {
      const int b = lane;
      const int e = m.b_last_dof[b];
      float v[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, a[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (e >= 0) {
#pragma unroll
        for (int i = 0; i < 6; ++i) { v[i] = X[e * kS6 + i]; a[i] = Y[e * kS6 + i]; }
      }
      a[5] += m.gravity;                       // gravity as a fictitious base acceleration
      const float mass = m.b_mass[b];
      float c[3];
      matvec3(R, m.b_com[b], c);
      c[0] += p[0]; c[1] += p[1]; c[2] += p[2];
      // Ic_world = R I R^T  (I symmetric: xx yy zz xy xz yz)
      const float* I6 = m.b_inertia[b];
      const float Im[9] = {I6[0], I6[3], I6[4], I6[3], I6[1], I6[5], I6[4], I6[5], I6[2]};
      float T[9];
      matmul3(R, Im, T);
      float Io[6];   // xx yy zz xy xz yz about O
      const float c2 = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
      Io[0] = T[0] * R[0] + T[1] * R[1] + T[2] * R[2] + mass * (c2 - c[0] * c[0]);
      Io[1] = T[3] * R[3] + T[4] * R[4] + T[5] * R[5] + mass * (c2 - c[1] * c[1]);
      Io[2] = T[6] * R[6] + T[7] * R[7] + T[8] * R[8] + mass * (c2 - c[2] * c[2]);
      Io[3] = T[0] * R[3] + T[1] * R[4] + T[2] * R[5] - mass * c[0] * c[1];
      Io[4] = T[0] * R[6] + T[1] * R[7] + T[2] * R[8] - mass * c[0] * c[2];
      Io[5] = T[3] * R[6] + T[4] * R[7] + T[5] * R[8] - mass * c[1] * c[2];
      const float h[3] = {mass * c[0], mass * c[1], mass * c[2]};
      // I x = [Io w + h x v ; m v - h x w]
      auto apply = [&](const float* xin, float* out) {
        float hv[3], hw[3];
        cross3(h, xin + 3, hv);
        cross3(h, xin, hw);
        out[0] = Io[0] * xin[0] + Io[3] * xin[1] + Io[4] * xin[2] + hv[0];
        out[1] = Io[3] * xin[0] + Io[1] * xin[1] + Io[5] * xin[2] + hv[1];
        out[2] = Io[4] * xin[0] + Io[5] * xin[1] + Io[2] * xin[2] + hv[2];
        out[3] = mass * xin[3] - hw[0];
        out[4] = mass * xin[4] - hw[1];
        out[5] = mass * xin[5] - hw[2];
      };
      float Ia[6], Iv[6], f[6];
      apply(a, Ia);
      apply(v, Iv);
      // v x* Iv = [w x n + v x f ; w x f]
      float t1[3], t2[3], t3[3];
      cross3(v, Iv, t1);
      cross3(v + 3, Iv + 3, t2);
      cross3(v, Iv + 3, t3);
      f[0] = Ia[0] + t1[0] + t2[0]; f[1] = Ia[1] + t1[1] + t2[1]; f[2] = Ia[2] + t1[2] + t2[2];
      f[3] = Ia[3] + t3[0]; f[4] = Ia[4] + t3[1]; f[5] = Ia[5] + t3[2];
      const int c0 = m.b_con_start[b], c1 = c0 + m.b_con_count[b];
      for (int k = c0; k < c1; ++k)
#pragma unroll
        for (int i = 0; i < 6; ++i) f[i] -= cw[k * kS6 + i];
      float* r = rec + b * kSRec;
#pragma unroll
      for (int i = 0; i < 6; ++i) r[i] = f[i];
      r[6] = mass; r[7] = h[0]; r[8] = h[1]; r[9] = h[2];
#pragma unroll
      for (int i = 0; i < 6; ++i) r[10 + i] = Io[i];
    }
    if (lane >= 16) rec[kZeroRec * kSRec + (lane - 16)] = 0.f;     // (the transposition buffer of step 5 aliases it)
    __syncwarp();

    // ---- 3b. leaf -> root accumulation, lane = (parent body, record component) ----------------------------------
    // child lists are padded to 4 with the all-zero record, so every item is 5 loads + 4 adds without branches
    for (int L = m.max_depth - 1; L >= 0; --L) {
      const int items = m.lvl_np[L] * 16;
      for (int it = lane; it < items; it += 32) {
        const int slot = it >> 4, comp = it & 15;
        const int b = m.lvl_parent[L][slot];
        const int4 ch = *reinterpret_cast<const int4*>(m.lvl_child[L][slot]);
        const float acc = rec[b * kSRec + comp] + rec[ch.x * kSRec + comp] + rec[ch.y * kSRec + comp] +
                          rec[ch.z * kSRec + comp] + rec[ch.w * kSRec + comp];
        rec[b * kSRec + comp] = acc;
      }
      __syncwarp();
    }

    // ---- 4. lane = dof: bias, applied torque, mass-matrix row -------------------------------------------------
    const int j = lane;
    float row[NVMAX];
#pragma unroll
    for (int c = 0; c < NVMAX; ++c) row[c] = 0.f;
    float rhs = 0.f;
    if (is_dof) {
      const float* r = rec + m.d_body[j] * kSRec;
      float bias = 0.f;
#pragma unroll
      for (int i = 0; i < 6; ++i) bias += S[i] * r[i];
      float tau = 0.f;
      const int act = m.d_act[j];
      if (act >= 0) tau = m.d_gear[j] * fminf(fmaxf(ctrl[act], -m.ctrl_limit), m.ctrl_limit);
      float keff = 0.f, beff = m.d_damp[j];
      if (dtype == kSlide || dtype == kHinge) {
        const float qj = q[m.d_qadr[j]];
        keff = m.d_stiff[j];
        tau -= keff * qj;
        if (m.d_limited[j]) {
          const bool below = qj < m.d_lo[j], above = qj > m.d_hi[j];
          if (below) tau += m.d_klim[j] * (m.d_lo[j] - qj);
          if (above) tau += m.d_klim[j] * (m.d_hi[j] - qj);
          if (below || above) { keff += m.d_klim[j]; beff += m.d_blim[j]; }
        }
      }
      // springs and dampers implicit: (M + dt B + dt^2 K) qacc = tau - (B + dt K) qd - bias
      rhs = tau - (beff + dt * keff) * qdj - bias;
      const float diag_add = m.d_arm[j] + dt * beff + dt * dt * keff;
      // F = Ic_body(j) S_j
      const float mass = r[6];
      const float h[3] = {r[7], r[8], r[9]};
      float hv[3], hw[3], F[6];
      cross3(h, S + 3, hv);
      cross3(h, S, hw);
      F[0] = r[10] * S[0] + r[13] * S[1] + r[14] * S[2] + hv[0];
      F[1] = r[13] * S[0] + r[11] * S[1] + r[15] * S[2] + hv[1];
      F[2] = r[14] * S[0] + r[15] * S[1] + r[12] * S[2] + hv[2];
      F[3] = mass * S[3] - hw[0];
      F[4] = mass * S[4] - hw[1];
      F[5] = mass * S[5] - hw[2];
      const unsigned chain = m.d_chain[j];
#pragma unroll
      for (int c = 0; c < NVMAX; ++c) {
        if ((chain >> c) & 1u) {
          const float* Sc = Sd + c * kS6;
          row[c] = Sc[0] * F[0] + Sc[1] * F[1] + Sc[2] * F[2] + Sc[3] * F[3] + Sc[4] * F[4] + Sc[5] * F[5];
        }
      }
#pragma unroll
      for (int c = 0; c < NVMAX; ++c) row[c] += (c == j) ? diag_add : 0.f;
    } else {
#pragma unroll
      for (int c = 0; c < NVMAX; ++c) row[c] = (c == j) ? 1.f : 0.f;      // padding rows: identity
    }

    // ---- 5. in-register Cholesky (lane i holds row i, columns 0..i) ---------------------------------------------
    // step k: every lane publishes its (unscaled) column-k entry in shared memory, then all lanes read the whole
    // column with broadcast 128-bit loads: row_i[c] -= A_ik A_ck / A_kk.  Two alternating buffers, one sync per step.
#pragma unroll
    for (int k = 0; k < NVMAX; ++k) {
      float* col = w + ((k & 1) ? oY : oX);
      if (lane < NVMAX) col[lane] = row[k];
      __syncwarp();
      float cv[NVMAX];
#pragma unroll
      for (int m4 = k / 4; m4 < NVMAX / 4; ++m4) {
        const float4 v = *reinterpret_cast<const float4*>(col + 4 * m4);
        cv[4 * m4] = v.x; cv[4 * m4 + 1] = v.y; cv[4 * m4 + 2] = v.z; cv[4 * m4 + 3] = v.w;
      }
      const float inv = rsqrtf(cv[k]);
      const float t = (lane > k) ? row[k] * inv * inv : 0.f;
      row[k] = (lane >= k) ? row[k] * inv : 0.f;      // lane k: sqrt(akk); lanes > k: l_ik
#pragma unroll
      for (int c = k + 1; c < NVMAX; ++c) row[c] = fmaf(-t, cv[c], row[c]);
    }
    __syncwarp();
    float diag = 1.f;
#pragma unroll
    for (int c = 0; c < NVMAX; ++c) diag = (c == lane) ? row[c] : diag;
    const float dinv = 1.f / diag;
    // forward substitution  L y = rhs
    float y = rhs;
#pragma unroll
    for (int k = 0; k < NVMAX; ++k) {
      const float yk = __shfl_sync(0xffffffffu, y * dinv, k);
      y = (lane == k) ? yk : ((lane > k) ? fmaf(-row[k], yk, y) : y);
    }
    // transpose the factor through shared memory: lane i then holds column i of L (= row i of L^T)
    float* Lt = w + oRb;
    __syncwarp();
    if (lane < NVMAX) {
#pragma unroll
      for (int c = 0; c < NVMAX; ++c) Lt[lane * kLd + c] = row[c];
    }
    __syncwarp();
    if (lane < NVMAX) {
#pragma unroll
      for (int c = 0; c < NVMAX; ++c) row[c] = Lt[c * kLd + lane];     // row[c] = L[c][lane], nonzero for c >= lane
    }
    // back substitution  L^T x = y
    float xs = y;
#pragma unroll
    for (int k = NVMAX - 1; k >= 0; --k) {
      const float xk = __shfl_sync(0xffffffffu, xs * dinv, k);
      xs = (lane == k) ? xk : ((lane < k) ? fmaf(-row[k], xk, xs) : xs);
    }

    // ---- 6. semi-implicit Euler -----------------------------------------------------------------------------------
    __syncwarp();
    float* qw = w + oState;
    float* qdw = w + oState + m.nq;
    if (is_dof) {
      const float qdn = qdj + dt * xs;
      qdw[j] = qdn;
      if (dtype != kFreeRot) qw[m.d_qadr[j]] += dt * qdn;
    }
    __syncwarp();
    if (is_dof && dtype == kFreeRot && (j == 0 || m.d_type[j - 1] != kFreeRot)) {
      const int qa = m.d_qadr[j];
      const float w0 = qdw[j], w1 = qdw[j + 1], w2 = qdw[j + 2];
      const float n = sqrtf(w0 * w0 + w1 * w1 + w2 * w2);
      const float half = 0.5f * dt * n;
      float sn, cs;
      sincosf(half, &sn, &cs);
      const float s = n > 1e-8f ? sn / n : 0.5f * dt;
      const float bw = cs, bx = w0 * s, by = w1 * s, bz = w2 * s;
      const float aw = qw[qa], ax = qw[qa + 1], ay = qw[qa + 2], az = qw[qa + 3];
      float rw = aw * bw - ax * bx - ay * by - az * bz;
      float rx = aw * bx + ax * bw + ay * bz - az * by;
      float ry = aw * by - ax * bz + ay * bw + az * bx;
      float rz = aw * bz + ax * by - ay * bx + az * bw;
      const float inv = rsqrtf(rw * rw + rx * rx + ry * ry + rz * rz);
      qw[qa] = rw * inv; qw[qa + 1] = rx * inv; qw[qa + 2] = ry * inv; qw[qa + 3] = rz * inv;
    }
    __syncwarp();
  }
};

}  // namespace icem
