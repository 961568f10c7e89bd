// Articulated rigid-body forward model, one warp per trajectory (ground-truth dynamics of the B200 path).
//
// Stands in for what the reference reaches through `GroundTruthModel.predict_n_steps` -> `env.step` -> MuJoCo
// (icem/models/gt_model.py:76-102, icem/environments/mujoco.py:101-131).  MuJoCo is not available;
// this is the repo's own soft-contact engine on the tables of icem_b200/robots.py, checked against the
// independent float64 restatement oracle/articulated_np.py.  PARITY WITH MUJOCO IS UNPINNED.
//
// Per substep (frame_skip substeps per control step, control held), all in one warp:
//   1. kinematics by tree level (lane = body): world rotation / position relative to the root origin O,
//      motion axes S_j (Plucker [w; v_O]), body velocity v_b and velocity-product acceleration a_b
//   2. floor contacts (lane = contact sphere): Hunt-Crossley normal force + capped viscous friction -> wrench
//   3. recursive Newton-Euler bias: f_b = I_b a_b + v_b x* I_b v_b - f_ext, accumulated leaf -> root together
//      with the composite inertias (lane = body, parents pull from children level by level)
//   4. lane = dof: bias_j = S_j . f, row j of the composite-rigid-body mass matrix (kept in REGISTERS),
//      applied torques (actuator gear, joint spring, limit spring-damper; springs/dampers implicit)
//   5. (M + dt B + dt^2 K) qacc = rhs by an in-register Cholesky factorisation: column broadcasts with warp shuffles,
//      forward substitution by shuffles, back substitution by warp reductions
//   6. semi-implicit Euler; unit-quaternion update for a free root joint
#pragma once
#include "common.cuh"

namespace icem {

constexpr int kArtMaxBodies = 16;
constexpr int kArtMaxDofs = 32;
constexpr int kArtMaxContacts = 32;
constexpr int kArtMaxChildren = 4;
enum { kSlide = 0, kHinge = 1, kFreeTrans = 2, kFreeRot = 3 };

// POD model tables; built on the host (planner.cu: icem_set_articulated_model), copied to shared memory per CTA.
struct ArtModel {
  int nb, nq, nv, nu, nc, nsub, max_depth, obs_offset;
  float dt, gravity, ctrl_limit, kc, cc, kv, mu, cdmax;
  int b_parent[kArtMaxBodies], b_depth[kArtMaxBodies], b_dof_start[kArtMaxBodies], b_dof_count[kArtMaxBodies];
  int b_nchild[kArtMaxBodies], b_child[kArtMaxBodies][kArtMaxChildren];
  int b_con_start[kArtMaxBodies], b_con_count[kArtMaxBodies];
  float b_pos[kArtMaxBodies][3], b_mass[kArtMaxBodies], b_com[kArtMaxBodies][3], b_inertia[kArtMaxBodies][6];
  int d_body[kArtMaxDofs], d_type[kArtMaxDofs], d_qadr[kArtMaxDofs], d_limited[kArtMaxDofs], d_act[kArtMaxDofs];
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

// padded strides of the per-warp arrays (odd strides: conflict-free when lane = row)
constexpr int kSR = 9, kSP = 3, kS6 = 7, kSI = 11;

template <int NVMAX>
struct Articulated {
  static constexpr int kWarpsPerCta = 8;
  struct Params {
    const ArtModel* model;   // device global memory
    int act_dim, nq, nv;
  };
  __host__ __device__ static int cta_floats(const Params&) { return (int)((sizeof(ArtModel) + 3) / 4); }
  __host__ __device__ static int warp_floats(const Params&) {
    return 64 /*state*/ + 4 /*O*/ + kArtMaxBodies * (kSR + kSP + 2 * kS6 /*v,a*/ + kS6 /*f*/ + kSI) +
           kArtMaxDofs * kS6 /*S*/ + kArtMaxContacts * kS6 /*contact wrenches*/ + kArtMaxDofs /*qacc*/;
  }
  __host__ __device__ static int state_dim(const Params& p) { return p.nq + p.nv; }

  __device__ static void cta_init(const Params& p, float* s) {
    const int n = (int)(sizeof(ArtModel) / 4);
    const float* src = reinterpret_cast<const float*>(p.model);
    for (int i = threadIdx.x; i < n; i += blockDim.x) s[i] = src[i];
  }

  const ArtModel* M;
  float *st, *O, *Rb, *pb, *vb, *ab, *fb, *Ib, *Sd, *cw, *acc;

  __device__ void bind(const Params&, const float* cta, float* warp) {
    M = reinterpret_cast<const ArtModel*>(cta);
    st = warp;
    O = st + 64;
    Rb = O + 4;
    pb = Rb + kArtMaxBodies * kSR;
    vb = pb + kArtMaxBodies * kSP;
    ab = vb + kArtMaxBodies * kS6;
    fb = ab + kArtMaxBodies * kS6;
    Ib = fb + kArtMaxBodies * kS6;
    Sd = Ib + kArtMaxBodies * kSI;
    cw = Sd + kArtMaxDofs * kS6;
    acc = cw + kArtMaxContacts * kS6;
  }
  __device__ void reset(const float* start_state) {
    const int n = M->nq + M->nv;
    for (int i = lane_id(); i < n; i += 32) st[i] = start_state[i];
    __syncwarp();
  }
  __device__ float obs(int i) const { return st[i + M->obs_offset]; }
  __device__ void export_state(float* out) const {
    const int n = M->nq + M->nv;
    for (int i = lane_id(); i < n; i += 32) out[i] = st[i];
  }
  __device__ void step(const float* ctrl) {
    for (int s = 0; s < M->nsub; ++s) substep(ctrl);
  }

  // ------------------------------------------------------------------------------------------------------------
  __device__ void substep(const float* ctrl) {
    const ArtModel& m = *M;
    const int lane = lane_id();
    const float* q = st;
    const float* qd = st + m.nq;
    const float dt = m.dt;

    // ---- 1. kinematics, velocities, velocity-product accelerations: lane = body, level by level -------------
    const bool is_body = lane < m.nb;
    const int depth = is_body ? m.b_depth[lane] : 1 << 20;
    float R[9], p[3], v[6], a[6];
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) p[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 6; ++i) { v[i] = 0.f; a[i] = 0.f; }
    for (int L = 0; L <= m.max_depth; ++L) {
      if (depth == L) {
        const int b = lane, par = m.b_parent[b];
        bool rel = true;          // positions already relative to O?
        if (par >= 0) {
          float Rp[9];
#pragma unroll
          for (int i = 0; i < 9; ++i) Rp[i] = Rb[par * kSR + i];
          float off[3];
          matvec3(Rp, m.b_pos[b], off);
#pragma unroll
          for (int i = 0; i < 3; ++i) p[i] = pb[par * kSP + i] + off[i];
#pragma unroll
          for (int i = 0; i < 9; ++i) R[i] = Rp[i];
#pragma unroll
          for (int i = 0; i < 6; ++i) { v[i] = vb[par * kS6 + i]; a[i] = ab[par * kS6 + i]; }
        } else {
          R[0] = R[4] = R[8] = 1.f;
#pragma unroll
          for (int i = 0; i < 3; ++i) p[i] = m.b_pos[b][i];     // absolute until the first rotation
          a[5] = m.gravity;                                      // gravity as a fictitious base acceleration
          rel = false;
        }
        const int j0 = m.b_dof_start[b], j1 = j0 + m.b_dof_count[b];
        for (int j = j0; j < j1; ++j) {
          const int t = m.d_type[j];
          if (t == kFreeTrans) {                 // free joint: 3 translations + ball, one block
            const int qa = m.d_qadr[j];
            if (!rel) { O[0] = q[qa]; O[1] = q[qa + 1]; O[2] = q[qa + 2]; p[0] = p[1] = p[2] = 0.f; rel = true; }
            const float w = q[qa + 3], x = q[qa + 4], y = q[qa + 5], z = q[qa + 6];
            R[0] = 1.f - 2.f * (y * y + z * z); R[1] = 2.f * (x * y - w * z); R[2] = 2.f * (x * z + w * y);
            R[3] = 2.f * (x * y + w * z); R[4] = 1.f - 2.f * (x * x + z * z); R[5] = 2.f * (y * z - w * x);
            R[6] = 2.f * (x * z - w * y); R[7] = 2.f * (y * z + w * x); R[8] = 1.f - 2.f * (x * x + y * y);
            float wl[3] = {qd[j + 3], qd[j + 4], qd[j + 5]}, ww[3];
            matvec3(R, wl, ww);
            const float vl[3] = {qd[j], qd[j + 1], qd[j + 2]};
#pragma unroll
            for (int k = 0; k < 3; ++k) {
              float* St = Sd + (j + k) * kS6;
              float* Sr = Sd + (j + 3 + k) * kS6;
#pragma unroll
              for (int i = 0; i < 6; ++i) { St[i] = 0.f; Sr[i] = 0.f; }
              St[3 + k] = 1.f;
              Sr[0] = R[k]; Sr[1] = R[3 + k]; Sr[2] = R[6 + k];   // column k of R, anchored at O
            }
            // translations: (v xm [0;e_k]) qd_k = [0 ; w x v_lin]; ball: sum_k (v_after xm S_k) qd_k =
            // (v + v_trans) xm [ww;0] = [w x ww ; (v_lin_prev + v_lin) x ww]   (v = 0 for a root joint)
            float c[3], vt[3] = {v[3] + vl[0], v[4] + vl[1], v[5] + vl[2]};
            cross3(v, vl, c);
            a[3] += c[0]; a[4] += c[1]; a[5] += c[2];
            cross3(v, ww, c);
            a[0] += c[0]; a[1] += c[1]; a[2] += c[2];
            cross3(vt, ww, c);
            a[3] += c[0]; a[4] += c[1]; a[5] += c[2];
            v[0] += ww[0]; v[1] += ww[1]; v[2] += ww[2];
            v[3] += vl[0]; v[4] += vl[1]; v[5] += vl[2];
            j += 5;
            continue;
          }
          float ax[3];
          matvec3(R, m.d_axis[j], ax);
          float S[6];
          const float qj = q[m.d_qadr[j]], qdj = qd[j];
          if (t == kSlide) {
            S[0] = S[1] = S[2] = 0.f; S[3] = ax[0]; S[4] = ax[1]; S[5] = ax[2];
            p[0] += ax[0] * qj; p[1] += ax[1] * qj; p[2] += ax[2] * qj;
          } else {
            if (!rel) { O[0] = p[0]; O[1] = p[1]; O[2] = p[2]; p[0] = p[1] = p[2] = 0.f; rel = true; }
            float an[3];
            matvec3(R, m.d_anchor[j], an);
            an[0] += p[0]; an[1] += p[1]; an[2] += p[2];
            // Rodrigues rotation about ax by qj
            float sn, cs;
            sincosf(qj, &sn, &cs);
            const float C = 1.f - cs;
            const float Rj[9] = {cs + ax[0] * ax[0] * C, ax[0] * ax[1] * C - ax[2] * sn, ax[0] * ax[2] * C + ax[1] * sn,
                                 ax[1] * ax[0] * C + ax[2] * sn, cs + ax[1] * ax[1] * C, ax[1] * ax[2] * C - ax[0] * sn,
                                 ax[2] * ax[0] * C - ax[1] * sn, ax[2] * ax[1] * C + ax[0] * sn, cs + ax[2] * ax[2] * C};
            float Rn[9];
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
              for (int c = 0; c < 3; ++c)
                Rn[r * 3 + c] = Rj[r * 3] * R[c] + Rj[r * 3 + 1] * R[3 + c] + Rj[r * 3 + 2] * R[6 + c];
#pragma unroll
            for (int i = 0; i < 9; ++i) R[i] = Rn[i];
            float dp[3] = {p[0] - an[0], p[1] - an[1], p[2] - an[2]}, rp[3];
            matvec3(Rj, dp, rp);
            p[0] = an[0] + rp[0]; p[1] = an[1] + rp[1]; p[2] = an[2] + rp[2];
            S[0] = ax[0]; S[1] = ax[1]; S[2] = ax[2];
            cross3(an, ax, S + 3);                 // velocity at O of a rotation about the anchored axis
          }
          // a += (v xm S) qd ; v += S qd      ([w1;v1] xm [w2;v2] = [w1 x w2 ; w1 x v2 + v1 x w2])
          float c1[3], c2[3], c3[3];
          cross3(v, S, c1);
          cross3(v, S + 3, c2);
          cross3(v + 3, S, c3);
          a[0] += c1[0] * qdj; a[1] += c1[1] * qdj; a[2] += c1[2] * qdj;
          a[3] += (c2[0] + c3[0]) * qdj; a[4] += (c2[1] + c3[1]) * qdj; a[5] += (c2[2] + c3[2]) * qdj;
#pragma unroll
          for (int i = 0; i < 6; ++i) {
            v[i] += S[i] * qdj;
            Sd[j * kS6 + i] = S[i];
          }
        }
        if (!rel) { O[0] = p[0]; O[1] = p[1]; O[2] = p[2]; p[0] = p[1] = p[2] = 0.f; }
#pragma unroll
        for (int i = 0; i < 9; ++i) Rb[b * kSR + i] = R[i];
#pragma unroll
        for (int i = 0; i < 3; ++i) pb[b * kSP + i] = p[i];
#pragma unroll
        for (int i = 0; i < 6; ++i) { vb[b * kS6 + i] = v[i]; ab[b * kS6 + i] = a[i]; }
      }
      __syncwarp();
    }

    // ---- 2. floor contacts: lane = contact sphere ---------------------------------------------------------------
    if (lane < m.nc) {
      const int b = m.c_body[lane];
      float x[3];
      matvec3(Rb + b * kSR, m.c_pos[lane], x);
      x[0] += pb[b * kSP]; x[1] += pb[b * kSP + 1]; x[2] += pb[b * kSP + 2];
      const float pen = m.c_radius[lane] - (O[2] + x[2]);
      float wr[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (pen > 0.f) {
        const float* vv = vb + b * kS6;
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
      const float mass = m.b_mass[b];
      float c[3];
      matvec3(R, m.b_com[b], c);
      c[0] += p[0]; c[1] += p[1]; c[2] += p[2];
      // Ic_world = R I R^T  (I symmetric: xx yy zz xy xz yz)
      const float* I6 = m.b_inertia[b];
      const float Im[9] = {I6[0], I6[3], I6[4], I6[3], I6[1], I6[5], I6[4], I6[5], I6[2]};
      float T[9];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int cc2 = 0; cc2 < 3; ++cc2)
          T[r * 3 + cc2] = R[r * 3] * Im[cc2] + R[r * 3 + 1] * Im[3 + cc2] + R[r * 3 + 2] * Im[6 + cc2];
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
#pragma unroll
      for (int i = 0; i < 6; ++i) fb[b * kS6 + i] = f[i];
      float* Ic = Ib + b * kSI;
      Ic[0] = mass; Ic[1] = h[0]; Ic[2] = h[1]; Ic[3] = h[2];
#pragma unroll
      for (int i = 0; i < 6; ++i) Ic[4 + i] = Io[i];
    }
    __syncwarp();

    // ---- 3b. leaf -> root accumulation of forces and composite inertias (parents pull) -----------------------
    for (int L = m.max_depth - 1; L >= 0; --L) {
      if (depth == L && m.b_nchild[lane] > 0) {
        const int b = lane;
        float f[6], Ic[10];
#pragma unroll
        for (int i = 0; i < 6; ++i) f[i] = fb[b * kS6 + i];
#pragma unroll
        for (int i = 0; i < 10; ++i) Ic[i] = Ib[b * kSI + i];
        for (int k = 0; k < m.b_nchild[b]; ++k) {
          const int ch = m.b_child[b][k];
#pragma unroll
          for (int i = 0; i < 6; ++i) f[i] += fb[ch * kS6 + i];
#pragma unroll
          for (int i = 0; i < 10; ++i) Ic[i] += Ib[ch * kSI + i];
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) fb[b * kS6 + i] = f[i];
#pragma unroll
        for (int i = 0; i < 10; ++i) Ib[b * kSI + i] = Ic[i];
      }
      __syncwarp();
    }

    // ---- 4. lane = dof: bias, applied torque, mass-matrix row -------------------------------------------------
    const int j = lane;
    const bool is_dof = j < m.nv;
    float row[NVMAX];
#pragma unroll
    for (int c = 0; c < NVMAX; ++c) row[c] = 0.f;
    float rhs = 0.f;
    float qdj = 0.f;
    if (is_dof) {
      const int b = m.d_body[j];
      float S[6];
#pragma unroll
      for (int i = 0; i < 6; ++i) S[i] = Sd[j * kS6 + i];
      float bias = 0.f;
#pragma unroll
      for (int i = 0; i < 6; ++i) bias += S[i] * fb[b * kS6 + i];
      qdj = qd[j];
      const int t = m.d_type[j];
      float tau = 0.f;
      const int act = m.d_act[j];
      if (act >= 0) tau = m.d_gear[j] * fminf(fmaxf(ctrl[act], -m.ctrl_limit), m.ctrl_limit);
      float keff = 0.f, beff = m.d_damp[j];
      if (t == kSlide || t == kHinge) {
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
      const float* Ic = Ib + b * kSI;
      const float mass = Ic[0];
      const float h[3] = {Ic[1], Ic[2], Ic[3]};
      float hv[3], hw[3], F[6];
      cross3(h, S + 3, hv);
      cross3(h, S, hw);
      F[0] = Ic[4] * S[0] + Ic[7] * S[1] + Ic[8] * S[2] + hv[0];
      F[1] = Ic[7] * S[0] + Ic[5] * S[1] + Ic[9] * S[2] + hv[1];
      F[2] = Ic[8] * S[0] + Ic[9] * S[1] + Ic[6] * S[2] + hv[2];
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
      for (int c = 0; c < NVMAX; ++c)
        if (c == j) row[c] += diag_add;
    } else {
#pragma unroll
      for (int c = 0; c < NVMAX; ++c)
        if (c == j) row[c] = 1.f;           // padding rows: identity
    }

    // ---- 5. in-register Cholesky (lane i holds row i, columns 0..i) ---------------------------------------------
#pragma unroll
    for (int k = 0; k < NVMAX; ++k) {
      const float akk = __shfl_sync(0xffffffffu, row[k], k);
      const float inv = rsqrtf(akk);
      row[k] = (lane >= k) ? row[k] * inv : 0.f;      // lane k: sqrt(akk); lanes > k: l_ik
#pragma unroll
      for (int c = k + 1; c < NVMAX; ++c) {
        const float lck = __shfl_sync(0xffffffffu, row[k], c);
        row[c] = fmaf(-row[k], lck, row[c]);
      }
    }
    float dinv = 1.f;
#pragma unroll
    for (int c = 0; c < NVMAX; ++c)
      if (c == lane) dinv = 1.f / row[c];
    // forward substitution  L y = rhs
    float y = rhs;
#pragma unroll
    for (int k = 0; k < NVMAX; ++k) {
      const float yk = __shfl_sync(0xffffffffu, y * dinv, k);
      y = (lane == k) ? yk : ((lane > k) ? fmaf(-row[k], yk, y) : y);
    }
    // back substitution  L^T x = y
    float xs = 0.f;
#pragma unroll
    for (int k = NVMAX - 1; k >= 0; --k) {
      const float s = warp_sum((lane > k) ? row[k] * xs : 0.f);
      if (lane == k) xs = (y - s) * dinv;
    }

    // ---- 6. semi-implicit Euler -----------------------------------------------------------------------------------
    __syncwarp();
    float* qw = st;
    float* qdw = st + m.nq;
    if (is_dof) {
      const float qdn = qdj + dt * xs;
      qdw[j] = qdn;
      const int t = m.d_type[j];
      if (t != kFreeRot) qw[m.d_qadr[j]] += dt * qdn;
    }
    __syncwarp();
    if (is_dof && m.d_type[j] == kFreeRot && (j == 0 || m.d_type[j - 1] != kFreeRot)) {
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
