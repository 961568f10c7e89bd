// Branch-parallel articulated-body engine: a GROUP of G lanes per trajectory, 32/G trajectories per warp.
//
// Same physical model and integrator as dyn_articulated.cuh / oracle/articulated_np.py (what the reference reaches
// through `GroundTruthModel.predict_n_steps` -> `env.step` -> MuJoCo, icem/models/gt_model.py:76-102,
// icem/environments/mujoco.py:101-131; PARITY WITH MUJOCO IS UNPINNED), different formulation:
//
//   dyn_articulated.cuh : one WARP per trajectory, lanes = bodies / dofs / contacts, composite-rigid-body mass matrix
//                         + dense in-register Cholesky  (~4100 warp instructions per substep per trajectory)
//   here                : Featherstone's articulated-body algorithm (no mass matrix, no factorisation, O(n)) on a
//                         robot decomposed into a TRUNK chain plus up to G LIMB chains hanging off trunk bodies.
//                         Lane g of a group walks the trunk (all lanes redundantly: the values stay bit-identical, no
//                         exchange needed) and then ITS limb; the only cross-lane traffic per dynamics evaluation is
//                         the sum of the limbs' articulated inertias / bias forces where they join the trunk
//                         (27 values, xor-shuffles inside the group).  Everything is expressed in world-aligned
//                         coordinates about a common origin O (the root position), so that child -> parent
//                         accumulation is a plain addition.  (Per-node reference points at the joints -- shorter
//                         lever arms, one translation of the articulated inertia per node -- were measured: +7 %
//                         time for 10-15 % less float32 error on HumanoidStandup costs, not kept.)
//
// HumanoidStandup: trunk = torso / lwaist / pelvis (9 dofs), limbs = two arms (3 dofs) and two legs (4 dofs, the
// jointless foot is fused into the shin): G = 4, 8 trajectories per warp, 13 sequential dof steps per lane instead
// of 23.  HalfCheetah: trunk = torso (3 dofs), limbs = back / front leg: G = 2, 16 trajectories per warp.
//
// The code is __host__ __device__ and templated on an execution context (group sum + group barrier): the GPU context
// uses warp shuffles; tests/chain_host compiles THE SAME SOURCE for the host with one thread per lane, so the engine
// is checked against the float64 oracle without a GPU (test infrastructure only, never a product path).
//
// Per dynamics evaluation (q, qd, ctrl) -> qacc:
//   pass 1 (root -> leaves): frames through the joints, motion axes S_j, velocity-product terms c_j, rigid inertia
//                            about O, bias force v x* I v minus floor-contact wrenches; joint-space force and the
//                            implicit spring/damper diagonal per dof
//   pass 2 (leaves -> root): U = IA S, D = S.U + diag, u = tau - S.pA, IA -= U U^T / D, pA += IA c + U u / D;
//                            limbs first (lanes in parallel), group sum at the junctions, then the trunk
//   pass 3 (root -> leaves): a += c; qacc = (u - U.a) / D; a += S qacc; semi-implicit Euler on the fly
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../../include/icem_b200.h"

#ifndef __CUDACC__
#define __host__
#define __device__
#define __forceinline__ inline
#endif

namespace icem {

#ifndef ICEM_ART_JOINT_TYPES
#define ICEM_ART_JOINT_TYPES
enum { kSlide = 0, kHinge = 1, kFreeTrans = 2, kFreeRot = 3 };
#endif

constexpr int kChMaxNodes = 16, kChMaxDofs = 32, kChMaxCon = 32;
constexpr int kChMaxTrunk = 4, kChMaxLimbs = 4, kChMaxLimbNodes = 4;
// dof record: S(6) c(6) U/D(6) u/D(1) [rhs diag].  A limb's records belong to one lane, which parks (rhs, diag)
// of pass 1 in the last two slots that pass 2 overwrites.  Trunk records are written by ALL lanes of the group with
// identical values, so nothing a lane still has to read may be overwritten there: (rhs, diag) get slots of their own.
constexpr int kChDofRec = 19, kChTrunkDofRec = 21;
constexpr int kChNodeRec = 16;    // mass, h(3), Io(6), bias force(6)
constexpr int kChFrame = 18;      // R(9) p(3) v(6) of a trunk body
constexpr int kChFreeRec = 24;    // free root: R(9) c_J linear part(3) | root acceleration(6) qacc(6) from pass 2
constexpr int kChJun = 27;        // articulated inertia (6 + 9 + 6) + bias force (6) summed over the limbs of a junction
enum { kChEuler = 0, kChRK4 = 1 };

// POD tables of the decomposed robot; built on the host by build_chain_model(), copied to shared memory per CTA.
// A NODE is a body together with the jointless bodies rigidly attached below it.
struct ChainModel {
  int nq, nv, nu, nsub, obs_offset, integrator;
  int n_nodes, n_trunk, n_limbs, lanes;          // lanes = G (1, 2 or 4)
  int max_limb_nodes, trunk_dofs, max_limb_dofs, n_junctions;
  float dt, gravity, ctrl_limit, kc, cc, kv, mu, cdmax;
  // scratch layout in slots (group-shared region: slot * (32 / G) + group; lane-private region: slot * 32 + lane)
  int s_state, s_free, s_dof, s_node, s_frame, s_acc, s_jun, s_rk, s_ctrl, s_end;
  int root_free;                                 // the root carries a free joint (handled as one 6-dof joint)
  int planar;                                    // every body moves in the x-z plane and turns about y (HalfCheetah, Hopper):
                                                 // the PLANAR instantiation of ChainLane may run this model
  int seq_last[kChMaxLimbs];                     // position of the last node of lane g's walk
  int p_dof, p_node, p_end;
  int trunk_node[kChMaxTrunk];
  int trunk_junction[kChMaxTrunk];               // junction slot of this trunk position, -1 = no limb hangs here
  int limb_attach[kChMaxLimbs];                  // trunk position the limb hangs off
  int limb_nnodes[kChMaxLimbs];
  int limb_node[kChMaxLimbs][kChMaxLimbNodes];
  int seq_node[kChMaxLimbs][kChMaxTrunk + kChMaxLimbNodes];   // node lane g visits at position i of its walk, -1 = none
  int n_dof_start[kChMaxNodes], n_dof_count[kChMaxNodes], n_con_start[kChMaxNodes], n_con_count[kChMaxNodes];
  float n_pos[kChMaxNodes][3], n_mass[kChMaxNodes], n_com[kChMaxNodes][3], n_inertia[kChMaxNodes][6];
  int d_type[kChMaxDofs], d_limited[kChMaxDofs], d_act[kChMaxDofs];
  // A dof's record sits at s_dof + 21 * index (trunk) / p_dof + 19 * index (own limb), index = its rank in walk order.  The
  // records of a node's (non-free) dofs are consecutive, so the walks address them from ONE table read per node:
  int n_slot0[kChMaxNodes];                      // slot of the record of the node's first slide / hinge dof
  // ... and passes 2 / 3 take the dof range of walk position i straight from per-lane tables (no node indirection):
  int seq_j0[kChMaxLimbs][kChMaxTrunk + kChMaxLimbNodes], seq_j1[kChMaxLimbs][kChMaxTrunk + kChMaxLimbNodes];
  int seq_slot0[kChMaxLimbs][kChMaxTrunk + kChMaxLimbNodes];
  float d_axis[kChMaxDofs][3], d_anchor[kChMaxDofs][3];
  float d_stiff[kChMaxDofs], d_damp[kChMaxDofs], d_arm[kChMaxDofs], d_lo[kChMaxDofs], d_hi[kChMaxDofs];
  float d_klim[kChMaxDofs], d_blim[kChMaxDofs], d_gear[kChMaxDofs];
  float c_pos[kChMaxCon][3], c_radius[kChMaxCon];
};

// ---- small vector helpers (registers) --------------------------------------------------------------------------
namespace ch {
__host__ __device__ __forceinline__ void cross(const float* a, const float* b, float* o) {
  const float x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z;
}
__host__ __device__ __forceinline__ void matvec(const float* R, const float* v, float* o) {   // R row-major 3x3
  const float x = R[0] * v[0] + R[1] * v[1] + R[2] * v[2];
  const float y = R[3] * v[0] + R[4] * v[1] + R[5] * v[2];
  const float z = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
  o[0] = x; o[1] = y; o[2] = z;
}
__host__ __device__ __forceinline__ void matmul(const float* A, const float* B, float* C) {
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) C[r * 3 + c] = A[r * 3] * B[c] + A[r * 3 + 1] * B[3 + c] + A[r * 3 + 2] * B[6 + c];
}
__host__ __device__ __forceinline__ float rsqrt_(float x) {
#ifdef __CUDA_ARCH__
  return rsqrtf(x);
#else
  return 1.0f / sqrtf(x);
#endif
}
__host__ __device__ __forceinline__ void sincos_(float x, float* s, float* c) {
#ifdef __CUDA_ARCH__
  sincosf(x, s, c);
#else
  *s = sinf(x); *c = cosf(x);
#endif
}
// spatial motion cross product  [w1; v1] xm [w2; v2] = [w1 x w2 ; w1 x v2 + v1 x w2]
__host__ __device__ __forceinline__ void cross_motion(const float* a, const float* b, float* o) {
  float c1[3], c2[3], c3[3];
  cross(a, b, c1);
  cross(a, b + 3, c2);
  cross(a + 3, b, c3);
  o[0] = c1[0]; o[1] = c1[1]; o[2] = c1[2];
  o[3] = c2[0] + c3[0]; o[4] = c2[1] + c3[1]; o[5] = c2[2] + c3[2];
}
}  // namespace ch

// symmetric 6x6 articulated inertia  [[A, B], [B^T, C]]  (A, C symmetric: xx yy zz xy xz yz; B general row-major)
struct ArtInertia {
  float A[6], B[9], C[6];
  __host__ __device__ __forceinline__ void zero() {
#pragma unroll
    for (int i = 0; i < 6; ++i) { A[i] = 0.f; C[i] = 0.f; }
#pragma unroll
    for (int i = 0; i < 9; ++i) B[i] = 0.f;
  }
  // y = I x,  x = [w; v]
  __host__ __device__ __forceinline__ void apply(const float* x, float* y) const {
    const float* w = x;
    const float* v = x + 3;
    y[0] = A[0] * w[0] + A[3] * w[1] + A[4] * w[2] + B[0] * v[0] + B[1] * v[1] + B[2] * v[2];
    y[1] = A[3] * w[0] + A[1] * w[1] + A[5] * w[2] + B[3] * v[0] + B[4] * v[1] + B[5] * v[2];
    y[2] = A[4] * w[0] + A[5] * w[1] + A[2] * w[2] + B[6] * v[0] + B[7] * v[1] + B[8] * v[2];
    y[3] = B[0] * w[0] + B[3] * w[1] + B[6] * w[2] + C[0] * v[0] + C[3] * v[1] + C[4] * v[2];
    y[4] = B[1] * w[0] + B[4] * w[1] + B[7] * w[2] + C[3] * v[0] + C[1] * v[1] + C[5] * v[2];
    y[5] = B[2] * w[0] + B[5] * w[1] + B[8] * w[2] + C[4] * v[0] + C[5] * v[1] + C[2] * v[2];
  }
  // I -= a b^T   with a b^T symmetric by construction (b = a / D)
  __host__ __device__ __forceinline__ void sub_outer(const float* a, const float* b) {
    A[0] -= a[0] * b[0]; A[1] -= a[1] * b[1]; A[2] -= a[2] * b[2];
    A[3] -= a[0] * b[1]; A[4] -= a[0] * b[2]; A[5] -= a[1] * b[2];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) B[r * 3 + c] -= a[r] * b[3 + c];
    C[0] -= a[3] * b[3]; C[1] -= a[4] * b[4]; C[2] -= a[5] * b[5];
    C[3] -= a[3] * b[4]; C[4] -= a[3] * b[5]; C[5] -= a[4] * b[5];
  }
  // rigid body: mass m, first moment h = m c, rotational inertia Io about O
  __host__ __device__ __forceinline__ void add_rigid(float m, const float* h, const float* Io) {
#pragma unroll
    for (int i = 0; i < 6; ++i) A[i] += Io[i];
    B[1] -= h[2]; B[2] += h[1]; B[3] += h[2]; B[5] -= h[0]; B[6] -= h[1]; B[7] += h[0];
    C[0] += m; C[1] += m; C[2] += m;
  }
};

// strided view of a record in shared memory; the stride (trajectories per warp on the GPU, 1 on the host) is a
// compile-time constant so that record fields are immediate offsets of ONE base address
template <int STRIDE>
struct ChRefT {
  float* p;
  __host__ __device__ __forceinline__ float& operator[](int i) const { return p[i * STRIDE]; }
};

__host__ __device__ __forceinline__ float ch_rcp(float x) {
#ifdef __CUDA_ARCH__
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
#else
  return 1.0f / x;
#endif
}

// x = M^-1 b for the symmetric positive definite 6x6 matrix M = [[A, B], [B^T, C]] (LDL^T, everything in registers)
__host__ __device__ __forceinline__ void ch_solve_spd6(const ArtInertia& I, const float* b, float* x) {
  float L[6][6];
  L[0][0] = I.A[0]; L[1][0] = I.A[3]; L[1][1] = I.A[1]; L[2][0] = I.A[4]; L[2][1] = I.A[5]; L[2][2] = I.A[2];
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) L[3 + r][c] = I.B[c * 3 + r];          // lower-left block = B^T
  L[3][3] = I.C[0]; L[4][3] = I.C[3]; L[4][4] = I.C[1]; L[5][3] = I.C[4]; L[5][4] = I.C[5]; L[5][5] = I.C[2];
  float dinv[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    float w[6];                       // w[k] = L[j][k] * d_k
    float d = L[j][j];
#pragma unroll
    for (int k = 0; k < j; ++k) { w[k] = L[j][k]; d -= w[k] * w[k] * dinv[k]; }
    dinv[j] = ch_rcp(d);
#pragma unroll
    for (int i = j + 1; i < 6; ++i) {
      float v = L[i][j];
#pragma unroll
      for (int k = 0; k < j; ++k) v -= L[i][k] * w[k] * dinv[k];
      L[i][j] = v;                    // unscaled: L[i][j] * d_j
    }
  }
  // (L D^-1) D (L D^-1)^T x = b   with the unscaled columns stored in L
  float z[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    float v = b[i];
#pragma unroll
    for (int k = 0; k < i; ++k) v -= L[i][k] * dinv[k] * z[k];
    z[i] = v;
  }
#pragma unroll
  for (int i = 5; i >= 0; --i) {
    float v = z[i] * dinv[i];
#pragma unroll
    for (int k = i + 1; k < 6; ++k) v -= L[k][i] * dinv[i] * x[k];
    x[i] = v;
  }
}

// One lane of a group.  Ctx provides: group_sum(float (&x)[N]) (sum over the G lanes, identical in all of them)
// and group_sync() (barrier + memory ordering among the lanes of the warp / group).
//
// PLANAR = true is the same algorithm with the structural zeros of a planar robot removed: a motion / force vector
// [w; v] keeps (w_y, v_x, v_z) -- entries 1, 3, 5 of the same arrays and record slots --, a rotation is (cos, sin) in
// R[0], R[2], the articulated inertia keeps A_yy, B_yx, B_yz, C_xx, C_zz, C_xz (A[1], B[3], B[5], C[0], C[2], C[4]).
// A dynamics evaluation then costs about a third of the spatial one; only models flagged ChainModel::planar may use it.
template <class Ctx, int STRIDE, bool PLANAR = false>
struct ChainLane {
  typedef ChRefT<STRIDE> ChRef;
  const ChainModel* M;
  float* sh;     // group-shared region, already offset by the group index; slot i lives at sh[i * STRIDE]
  float* pr;     // lane-private region, already offset for this lane; same stride
  int g;         // lane inside the group == limb index
  Ctx* ctx;

  __host__ __device__ __forceinline__ ChRef shared_rec(int slot) const { return ChRef{sh + slot * STRIDE}; }
  __host__ __device__ __forceinline__ ChRef private_rec(int slot) const { return ChRef{pr + slot * STRIDE}; }
  __host__ __device__ __forceinline__ float& state(int i) const { return sh[(M->s_state + i) * STRIDE]; }
  // first dof record of a node (slot from a per-node / per-walk-position table) and the distance between records
  __host__ __device__ __forceinline__ float* rec_base(bool trunk, int slot0) const {
    return (trunk ? sh : pr) + slot0 * STRIDE;
  }
  static __host__ __device__ __forceinline__ int rec_step(bool trunk) {
    return (trunk ? kChTrunkDofRec : kChDofRec) * STRIDE;
  }
  // Node records: trunk nodes in the shared region; a limb's FIRST node is parked in the junction region (idle
  // between the limb's start and the junction sums: 16 slots per lane of the group), its last node stays in
  // registers, the ones in between (limbs of 3+ bodies) go to the lane-private region.
  __host__ __device__ __forceinline__ ChRef node_rec(bool trunk, int pos) const {
    if (trunk) return shared_rec(M->s_node + kChNodeRec * pos);
    return pos == 0 ? shared_rec(M->s_frame + kChNodeRec * g) : private_rec(M->p_node + kChNodeRec * (pos - 1));
  }
  // node this lane visits at position i of its walk (trunk nodes, then its own limb); false = nothing to do
  __host__ __device__ __forceinline__ bool node_at(int i, int n_trunk, int& node, bool& trunk, int& pos) const {
    const ChainModel& m = *M;
    trunk = i < n_trunk;
    pos = trunk ? i : i - n_trunk;
    node = m.seq_node[g][i];
    return node >= 0;
  }

  // ---- pass 1 ----------------------------------------------------------------------------------------------------
  // Joints of one node: frame (R, p) through the joints, motion axes S_j, velocity-product terms c_j, the joint-space
  // force and the implicit spring / damper diagonal -> dof records; v becomes the node's velocity.
  __host__ __device__ __forceinline__ void walk_joints(int node, bool root, bool trunk, float (&R)[9], float (&p)[3],
                                                       float (&v)[6], float (&O)[3], const ChRef& q, const ChRef& qd,
                                                       const ChRef& ctrl, float dt) {
    const ChainModel& m = *M;
    if constexpr (PLANAR) {
      walk_joints_planar(node, root, trunk, R, p, v, O, q, qd, ctrl, dt);
      return;
    }
    if (root) {
#pragma unroll
      for (int k = 0; k < 3; ++k) O[k] = m.n_pos[node][k];     // p stays 0: everything is relative to O
    } else {
      float off[3];
      ch::matvec(R, m.n_pos[node], off);
      p[0] += off[0]; p[1] += off[1]; p[2] += off[2];
    }
    int j0 = m.n_dof_start[node];
    const int j1 = j0 + m.n_dof_count[node];
    const int rf = m.root_free;          // q index of a slide / hinge dof j: j + 1 behind a free joint, else j
    if (root && rf) {
      // free joint: 3 world translations + 3 body-frame rotation rates, handled as ONE 6-dof joint.  Its motion
      // subspace spans all of R^6, so pass 2 solves the root acceleration directly (floating base) and no per-dof
      // record is needed: only R and the velocity-product term  c_J = sum_k (v xm S_k) qd_k = [0; v_lin x w].
      const int qa = 0;                  // the free joint is the first joint of the root: coordinates 0..6
      O[0] = q[qa]; O[1] = q[qa + 1]; O[2] = q[qa + 2];
      float qw = q[qa + 3], x = q[qa + 4], y = q[qa + 5], z = q[qa + 6];
      {   // rotation of the NORMALISED quaternion (a start state may carry reset noise on it): R must be orthonormal,
          // pass 2 turns the root's angular acceleration into rotation rates with R^T
        const float qn = ch::rsqrt_(qw * qw + x * x + y * y + z * z);
        qw *= qn; x *= qn; y *= qn; z *= qn;
      }
      R[0] = 1.f - 2.f * (y * y + z * z); R[1] = 2.f * (x * y - qw * z); R[2] = 2.f * (x * z + qw * y);
      R[3] = 2.f * (x * y + qw * z); R[4] = 1.f - 2.f * (x * x + z * z); R[5] = 2.f * (y * z - qw * x);
      R[6] = 2.f * (x * z - qw * y); R[7] = 2.f * (y * z + qw * x); R[8] = 1.f - 2.f * (x * x + y * y);
      const float w0 = qd[j0 + 3], w1 = qd[j0 + 4], w2 = qd[j0 + 5];
      v[0] = R[0] * w0 + R[1] * w1 + R[2] * w2;
      v[1] = R[3] * w0 + R[4] * w1 + R[5] * w2;
      v[2] = R[6] * w0 + R[7] * w1 + R[8] * w2;
      v[3] = qd[j0]; v[4] = qd[j0 + 1]; v[5] = qd[j0 + 2];
      const ChRef fr = shared_rec(m.s_free);
#pragma unroll
      for (int k = 0; k < 9; ++k) fr[k] = R[k];
      float cl[3];
      ch::cross(v + 3, v, cl);
      fr[9] = cl[0]; fr[10] = cl[1]; fr[11] = cl[2];
      j0 += 6;
    }
    float* rb = rec_base(trunk, m.n_slot0[node]);
    const int rs = rec_step(trunk);
    for (int j = j0; j < j1; ++j, rb += rs) {
      const int t = m.d_type[j];
      float ax[3], S[6], c[6];
      ch::matvec(R, m.d_axis[j], ax);
      const float qj = q[j + rf], qdj = qd[j];
      if (t == kSlide) {
        S[0] = S[1] = S[2] = 0.f;
        S[3] = ax[0]; S[4] = ax[1]; S[5] = ax[2];
        float* dst = root ? O : p;           // root slides move the origin itself
        dst[0] += ax[0] * qj; dst[1] += ax[1] * qj; dst[2] += ax[2] * qj;
      } else {
        float an[3];
        ch::matvec(R, m.d_anchor[j], an);
        an[0] += p[0]; an[1] += p[1]; an[2] += p[2];
        S[0] = ax[0]; S[1] = ax[1]; S[2] = ax[2];
        ch::cross(an, ax, S + 3);            // velocity at O of a unit rotation about the anchored axis
        float sn, cs;
        ch::sincos_(qj, &sn, &cs);
        const float C = 1.f - cs;
        const float Rj[9] = {cs + ax[0] * ax[0] * C, ax[0] * ax[1] * C - ax[2] * sn, ax[0] * ax[2] * C + ax[1] * sn,
                             ax[1] * ax[0] * C + ax[2] * sn, cs + ax[1] * ax[1] * C, ax[1] * ax[2] * C - ax[0] * sn,
                             ax[2] * ax[0] * C - ax[1] * sn, ax[2] * ax[1] * C + ax[0] * sn, cs + ax[2] * ax[2] * C};
        float Rn[9];
        ch::matmul(Rj, R, Rn);
#pragma unroll
        for (int k = 0; k < 9; ++k) R[k] = Rn[k];
        const float dp[3] = {p[0] - an[0], p[1] - an[1], p[2] - an[2]};
        float rp[3];
        ch::matvec(Rj, dp, rp);
        p[0] = an[0] + rp[0]; p[1] = an[1] + rp[1]; p[2] = an[2] + rp[2];
      }
      ch::cross_motion(v, S, c);             // the axis is fixed in the frame moving with v (before this joint)
      // joint-space force and the implicit spring / damper diagonal (oracle/articulated_np.py:195-229)
      float tau = 0.f;
      const int act = m.d_act[j];
      if (act >= 0) tau = m.d_gear[j] * fminf(fmaxf(ctrl[act], -m.ctrl_limit), m.ctrl_limit);
      float keff = m.d_stiff[j], beff = m.d_damp[j];
      tau -= keff * qj;
      if (m.d_limited[j]) {
        const bool below = qj < m.d_lo[j], above = qj > m.d_hi[j];
        if (below) tau += m.d_klim[j] * (m.d_lo[j] - qj);
        if (above) tau += m.d_klim[j] * (m.d_hi[j] - qj);
        if (below || above) { keff += m.d_klim[j]; beff += m.d_blim[j]; }
      }
      const ChRef r{rb};
#pragma unroll
      for (int e = 0; e < 6; ++e) {
        r[e] = S[e];
        r[6 + e] = c[e] * qdj;
        v[e] += S[e] * qdj;
      }
      r[trunk ? 19 : 17] = tau - (beff + dt * keff) * qdj;
      r[trunk ? 20 : 18] = m.d_arm[j] + dt * beff + dt * dt * keff;
    }
  }

  // ---- planar twins of walk_joints / node_work (same formulas with the zero components dropped) --------------------
  __host__ __device__ __forceinline__ void walk_joints_planar(int node, bool root, bool trunk, float (&R)[9],
                                                              float (&p)[3], float (&v)[6], float (&O)[3],
                                                              const ChRef& q, const ChRef& qd, const ChRef& ctrl,
                                                              float dt) {
    const ChainModel& m = *M;
    if (root) {
      O[0] = m.n_pos[node][0]; O[2] = m.n_pos[node][2];
    } else {
      p[0] += R[0] * m.n_pos[node][0] + R[2] * m.n_pos[node][2];
      p[2] += R[0] * m.n_pos[node][2] - R[2] * m.n_pos[node][0];
    }
    const int j0 = m.n_dof_start[node], j1 = j0 + m.n_dof_count[node];
    float* rb = rec_base(trunk, m.n_slot0[node]);
    const int rs = rec_step(trunk);
    for (int j = j0; j < j1; ++j, rb += rs) {
      const int t = m.d_type[j];
      const float qj = q[j], qdj = qd[j];          // no free joint in a planar model: q index == dof index
      float Sw, Sx, Sz;
      if (t == kSlide) {
        const float ax = R[0] * m.d_axis[j][0] + R[2] * m.d_axis[j][2];
        const float az = R[0] * m.d_axis[j][2] - R[2] * m.d_axis[j][0];
        Sw = 0.f; Sx = ax; Sz = az;
        float* dst = root ? O : p;           // root slides move the origin itself
        dst[0] += ax * qj; dst[2] += az * qj;
      } else {
        const float sg = m.d_axis[j][1];     // +-1: hinge about +y or -y
        const float anx = R[0] * m.d_anchor[j][0] + R[2] * m.d_anchor[j][2] + p[0];
        const float anz = R[0] * m.d_anchor[j][2] - R[2] * m.d_anchor[j][0] + p[2];
        Sw = sg; Sx = -anz * sg; Sz = anx * sg;      // [axis; anchor x axis]
        float sn, cs;
        ch::sincos_(qj, &sn, &cs);
        sn *= sg;
        const float c0 = R[0], s0 = R[2];
        R[0] = cs * c0 - sn * s0;
        R[2] = cs * s0 + sn * c0;
        const float dx = p[0] - anx, dz = p[2] - anz;
        p[0] = anx + (cs * dx + sn * dz);
        p[2] = anz + (cs * dz - sn * dx);
      }
      // c = v xm S = [0; w x S_lin + v_lin x S_ang]
      const float cx = v[1] * Sz - v[5] * Sw, cz = v[3] * Sw - v[1] * Sx;
      float tau = 0.f;
      const int act = m.d_act[j];
      if (act >= 0) tau = m.d_gear[j] * fminf(fmaxf(ctrl[act], -m.ctrl_limit), m.ctrl_limit);
      float keff = m.d_stiff[j], beff = m.d_damp[j];
      tau -= keff * qj;
      if (m.d_limited[j]) {
        const bool below = qj < m.d_lo[j], above = qj > m.d_hi[j];
        if (below) tau += m.d_klim[j] * (m.d_lo[j] - qj);
        if (above) tau += m.d_klim[j] * (m.d_hi[j] - qj);
        if (below || above) { keff += m.d_klim[j]; beff += m.d_blim[j]; }
      }
      const ChRef r{rb};
      r[1] = Sw; r[3] = Sx; r[5] = Sz;
      r[9] = cx * qdj; r[11] = cz * qdj;
      v[1] += Sw * qdj; v[3] += Sx * qdj; v[5] += Sz * qdj;
      r[trunk ? 19 : 17] = tau - (beff + dt * keff) * qdj;
      r[trunk ? 20 : 18] = m.d_arm[j] + dt * beff + dt * dt * keff;
    }
  }

  __host__ __device__ __forceinline__ void node_work_planar(int node, const float (&R)[9], const float (&p)[3],
                                                            const float (&v)[6], float Oz,
                                                            float (&rec)[kChNodeRec]) const {
    const ChainModel& m = *M;
    const float mass = m.n_mass[node];
    const float cx = R[0] * m.n_com[node][0] + R[2] * m.n_com[node][2] + p[0];
    const float cz = R[0] * m.n_com[node][2] - R[2] * m.n_com[node][0] + p[2];
    const float Io = m.n_inertia[node][1] + mass * (cx * cx + cz * cz);     // I_yy is invariant under turns about y
    const float hx = mass * cx, hz = mass * cz;
    const float wy = v[1], vx = v[3], vz = v[5];
    // I v = [Io w + h x vl ; m vl - h x w],  bias = v x* (I v)
    const float Iv1 = Io * wy + (hz * vx - hx * vz);
    const float Iv3 = mass * vx + hz * wy;
    const float Iv5 = mass * vz - hx * wy;
    float f1 = vz * Iv3 - vx * Iv5, f3 = wy * Iv5, f5 = -wy * Iv3;
    const int k0 = m.n_con_start[node], k1 = k0 + m.n_con_count[node];
    for (int k = k0; k < k1; ++k) {
      const float xz = R[0] * m.c_pos[k][2] - R[2] * m.c_pos[k][0] + p[2];
      const float pen = m.c_radius[k] - (Oz + xz);
      if (pen > 0.f) {
        const float xx = R[0] * m.c_pos[k][0] + R[2] * m.c_pos[k][2] + p[0];
        const float ux = wy * xz + vx, uz = vz - wy * xx;
        const float spring = m.kc * pen;
        const float damp = fminf(spring * m.cc, m.cdmax);
        const float fn = fminf(fmaxf(spring - damp * uz, 0.f), 3.f * spring);
        const float coef = fminf(m.kv, m.mu * fn * ch_rcp(fmaxf(fabsf(ux), 1e-6f)));
        const float fcx = -coef * ux;
        f1 -= xz * fcx - xx * fn;
        f3 -= fcx;
        f5 -= fn;
      }
    }
    rec[0] = mass; rec[1] = hx; rec[3] = hz; rec[5] = Io;
    rec[11] = f1; rec[13] = f3; rec[15] = f5;
  }

  // Body-level work of one node given its frame and velocity: rigid inertia about O, bias force v x* I v minus the
  // floor-contact wrenches -> rec = (mass, h, Io, bias force).
  __host__ __device__ __forceinline__ void node_work(int node, const float (&R)[9], const float (&p)[3],
                                                     const float (&v)[6], float Oz, float (&rec)[kChNodeRec]) const {
    if constexpr (PLANAR) {
      node_work_planar(node, R, p, v, Oz, rec);
      return;
    }
    const ChainModel& m = *M;
    const float mass = m.n_mass[node];
    float c[3];
    ch::matvec(R, m.n_com[node], c);
    c[0] += p[0]; c[1] += p[1]; c[2] += p[2];
    const float* I6 = m.n_inertia[node];
    const float Im[9] = {I6[0], I6[3], I6[4], I6[3], I6[1], I6[5], I6[4], I6[5], I6[2]};
    float T[9];
    ch::matmul(R, Im, T);
    float Io[6];
    const float c2 = c[0] * c[0] + c[1] * c[1] + c[2] * c[2];
    Io[0] = T[0] * R[0] + T[1] * R[1] + T[2] * R[2] + mass * (c2 - c[0] * c[0]);
    Io[1] = T[3] * R[3] + T[4] * R[4] + T[5] * R[5] + mass * (c2 - c[1] * c[1]);
    Io[2] = T[6] * R[6] + T[7] * R[7] + T[8] * R[8] + mass * (c2 - c[2] * c[2]);
    Io[3] = T[0] * R[3] + T[1] * R[4] + T[2] * R[5] - mass * c[0] * c[1];
    Io[4] = T[0] * R[6] + T[1] * R[7] + T[2] * R[8] - mass * c[0] * c[2];
    Io[5] = T[3] * R[6] + T[4] * R[7] + T[5] * R[8] - mass * c[1] * c[2];
    const float h[3] = {mass * c[0], mass * c[1], mass * c[2]};
    // I v = [Io w + h x vl ; m vl - h x w],  bias = v x* (I v) = [w x n + vl x f ; w x f]
    float hv[3], hw[3], Iv[6], t1[3], t2[3], t3[3], f[6];
    ch::cross(h, v + 3, hv);
    ch::cross(h, v, hw);
    Iv[0] = Io[0] * v[0] + Io[3] * v[1] + Io[4] * v[2] + hv[0];
    Iv[1] = Io[3] * v[0] + Io[1] * v[1] + Io[5] * v[2] + hv[1];
    Iv[2] = Io[4] * v[0] + Io[5] * v[1] + Io[2] * v[2] + hv[2];
    Iv[3] = mass * v[3] - hw[0];
    Iv[4] = mass * v[4] - hw[1];
    Iv[5] = mass * v[5] - hw[2];
    ch::cross(v, Iv, t1);
    ch::cross(v + 3, Iv + 3, t2);
    ch::cross(v, Iv + 3, t3);
    f[0] = t1[0] + t2[0]; f[1] = t1[1] + t2[1]; f[2] = t1[2] + t2[2];
    f[3] = t3[0]; f[4] = t3[1]; f[5] = t3[2];
    const int k0 = m.n_con_start[node], k1 = k0 + m.n_con_count[node];
    for (int k = k0; k < k1; ++k) {
      // height first: most spheres are above the floor most of the time
      const float xz = R[6] * m.c_pos[k][0] + R[7] * m.c_pos[k][1] + R[8] * m.c_pos[k][2] + p[2];
      const float pen = m.c_radius[k] - (Oz + xz);
      if (pen > 0.f) {
        float x[3], u[3];
        x[0] = R[0] * m.c_pos[k][0] + R[1] * m.c_pos[k][1] + R[2] * m.c_pos[k][2] + p[0];
        x[1] = R[3] * m.c_pos[k][0] + R[4] * m.c_pos[k][1] + R[5] * m.c_pos[k][2] + p[1];
        x[2] = xz;
        ch::cross(v, x, u);
        u[0] += v[3]; u[1] += v[4]; u[2] += v[5];
        const float spring = m.kc * pen;
        const float damp = fminf(spring * m.cc, m.cdmax);
        const float fn = fminf(fmaxf(spring - damp * u[2], 0.f), 3.f * spring);
        const float speed = sqrtf(u[0] * u[0] + u[1] * u[1]);
        const float coef = fminf(m.kv, m.mu * fn * ch_rcp(fmaxf(speed, 1e-6f)));
        const float fc[3] = {-coef * u[0], -coef * u[1], fn};
        float mo[3];
        ch::cross(x, fc, mo);
        f[0] -= mo[0]; f[1] -= mo[1]; f[2] -= mo[2];
        f[3] -= fc[0]; f[4] -= fc[1]; f[5] -= fc[2];
      }
    }
    rec[0] = mass; rec[1] = h[0]; rec[2] = h[1]; rec[3] = h[2];
#pragma unroll
    for (int e = 0; e < 6; ++e) { rec[4 + e] = Io[e]; rec[10 + e] = f[e]; }
  }

  // frame (R, p, v) of a trunk body and node records (mass, h, Io, bias force) in shared memory: the planar engine
  // moves only the entries it uses (same slots)
  __host__ __device__ __forceinline__ void load_frame(const ChRef& f, float (&R)[9], float (&p)[3], float (&v)[6]) const {
    if constexpr (PLANAR) {
      R[0] = f[0]; R[2] = f[2]; p[0] = f[9]; p[2] = f[11]; v[1] = f[13]; v[3] = f[15]; v[5] = f[17];
    } else {
#pragma unroll
      for (int k = 0; k < 9; ++k) R[k] = f[k];
#pragma unroll
      for (int k = 0; k < 3; ++k) p[k] = f[9 + k];
#pragma unroll
      for (int k = 0; k < 6; ++k) v[k] = f[12 + k];
    }
  }
  __host__ __device__ __forceinline__ void store_node_rec(const ChRef& nr, const float (&rec)[kChNodeRec]) const {
    if constexpr (PLANAR) {
      nr[0] = rec[0]; nr[1] = rec[1]; nr[3] = rec[3]; nr[5] = rec[5]; nr[11] = rec[11]; nr[13] = rec[13]; nr[15] = rec[15];
    } else {
#pragma unroll
      for (int e = 0; e < kChNodeRec; ++e) nr[e] = rec[e];
    }
  }

  // `rec` returns the record of the LAST node of this lane's limb: pass 2 starts with that node, so its record never
  // goes through shared memory.  `implicit`: joint springs / dampers enter the system matrix (dt > 0 in the diagonal).
  //   1. every lane walks the trunk's joints (identical values in all lanes; frames -> shared memory);
  //   2. the trunk's BODY-level work (inertia, bias force, contacts) is dealt out: lane g takes trunk node g, g + G, ...
  //      -- the lanes work on different nodes at the same time instead of all repeating all of them;
  //   3. every lane walks its own limb (joints + body-level work).
  __host__ __device__ void pass1(const ChRef& q, const ChRef& qd, const ChRef& ctrl, bool implicit, float (&rec)[kChNodeRec]) {
    const ChainModel& m = *M;
    const float dt = implicit ? m.dt : 0.f;
    float R[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f}, p[3] = {0.f, 0.f, 0.f};
    float v[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float O[3] = {0.f, 0.f, 0.f};
    for (int t = 0; t < m.n_trunk; ++t) {
      walk_joints(m.trunk_node[t], t == 0, true, R, p, v, O, q, qd, ctrl, dt);
      const ChRef f = shared_rec(m.s_frame + kChFrame * t);
      if constexpr (PLANAR) {
        f[0] = R[0]; f[2] = R[2]; f[9] = p[0]; f[11] = p[2]; f[13] = v[1]; f[15] = v[3]; f[17] = v[5];
      } else {
#pragma unroll
        for (int k = 0; k < 9; ++k) f[k] = R[k];
#pragma unroll
        for (int k = 0; k < 3; ++k) f[9 + k] = p[k];
#pragma unroll
        for (int k = 0; k < 6; ++k) f[12 + k] = v[k];
      }
    }
    const float Oz = O[2];
    // every lane wrote every frame (identical values), so a lane could read right behind its own stores; the barrier
    // keeps the access pattern formally race-free (compute-sanitizer racecheck) at the price of one WARPSYNC
    ctx->group_sync();
    for (int t = g; t < m.n_trunk; t += m.lanes) {
      load_frame(shared_rec(m.s_frame + kChFrame * t), R, p, v);
      node_work(m.trunk_node[t], R, p, v, Oz, rec);
      store_node_rec(shared_rec(m.s_node + kChNodeRec * t), rec);
    }
    if (g < m.n_limbs) {
      load_frame(shared_rec(m.s_frame + kChFrame * m.limb_attach[g]), R, p, v);
    }
    ctx->group_sync();      // every lane has its attach frame: the frame region may now park the limbs' first records
    if (g < m.n_limbs) {
      const int nn = m.limb_nnodes[g];
      for (int i = 0; i < nn; ++i) {
        const int node = m.limb_node[g][i];
        walk_joints(node, false, false, R, p, v, O, q, qd, ctrl, dt);
        node_work(node, R, p, v, Oz, rec);
        if (i != nn - 1) store_node_rec(node_rec(false, i), rec);
      }
    }
  }

  // ---- pass 2 ----------------------------------------------------------------------------------------------------
  __host__ __device__ void pass2(const float (&rec)[kChNodeRec]) {
    if constexpr (PLANAR) {
      pass2_planar(rec);
      return;
    }
    const ChainModel& m = *M;
    ArtInertia IA;
    float P[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    IA.zero();
    // loop-invariant scalars of the model are read ONCE: the model lives in shared memory, and the compiler must assume
    // that the record stores of the loop alias it, so every `m.x` inside the loop is a dependent shared-memory read
    const int n_trunk = m.n_trunk, n_junctions = m.n_junctions;
    const int n_seq = n_trunk + m.max_limb_nodes;
    const int last = m.seq_last[g];
    for (int i = n_seq - 1; i >= 0; --i) {
      if (i == n_trunk - 1 && n_junctions > 0) {
        // the limbs are done: sum them where they join the trunk (every lane gets every junction's sum)
        const int mine = g < m.n_limbs ? m.trunk_junction[m.limb_attach[g]] : -1;
        for (int js = 0; js < n_junctions; ++js) {
          const bool w = mine == js;
          float x[kChJun];
#pragma unroll
          for (int e = 0; e < 6; ++e) { x[e] = w ? IA.A[e] : 0.f; x[15 + e] = w ? IA.C[e] : 0.f; x[21 + e] = w ? P[e] : 0.f; }
#pragma unroll
          for (int e = 0; e < 9; ++e) x[6 + e] = w ? IA.B[e] : 0.f;
          ctx->group_sum(x);
          if (js == 0) ctx->group_sync();      // the junction region changes tenant: parked limb records -> sums
          const ChRef jr = shared_rec(m.s_jun + kChJun * js);
#pragma unroll
          for (int e = 0; e < kChJun; ++e) jr[e] = x[e];
        }
        IA.zero();
#pragma unroll
        for (int e = 0; e < 6; ++e) P[e] = 0.f;
      }
      if (i == n_trunk - 1) ctx->group_sync();      // the trunk's node records were written by different lanes
      int node, pos;
      bool trunk;
      if (!node_at(i, n_trunk, node, trunk, pos)) continue;
      if (i == last) {
        const float h[3] = {rec[1], rec[2], rec[3]};
#pragma unroll
        for (int e = 0; e < 6; ++e) P[e] += rec[10 + e];
        IA.add_rigid(rec[0], h, rec + 4);
      } else {
        const ChRef nr = node_rec(trunk, pos);
        const float h[3] = {nr[1], nr[2], nr[3]};
        float Io[6];
#pragma unroll
        for (int e = 0; e < 6; ++e) { Io[e] = nr[4 + e]; P[e] += nr[10 + e]; }
        IA.add_rigid(nr[0], h, Io);
      }
      if (trunk && m.trunk_junction[i] >= 0) {
        const ChRef jr = shared_rec(m.s_jun + kChJun * m.trunk_junction[i]);
#pragma unroll
        for (int e = 0; e < 6; ++e) { IA.A[e] += jr[e]; IA.C[e] += jr[15 + e]; P[e] += jr[21 + e]; }
#pragma unroll
        for (int e = 0; e < 9; ++e) IA.B[e] += jr[6 + e];
      }
      int j0 = m.seq_j0[g][i];
      const int j1 = m.seq_j1[g][i];
      const bool free_root = i == 0 && m.root_free;
      if (free_root) j0 += 6;
      const int rs = rec_step(trunk);
      float* rb = rec_base(trunk, m.seq_slot0[g][i]) + (j1 - 1 - j0) * rs;
      for (int j = j1 - 1; j >= j0; --j, rb -= rs) {
        const ChRef r{rb};
        float S[6], c[6], U[6], Ud[6];
#pragma unroll
        for (int e = 0; e < 6; ++e) { S[e] = r[e]; c[e] = r[6 + e]; }
        IA.apply(S, U);
        float D = r[trunk ? 20 : 18], u = r[trunk ? 19 : 17];
#pragma unroll
        for (int e = 0; e < 6; ++e) { D += S[e] * U[e]; u -= S[e] * P[e]; }
        const float invD = ch_rcp(D);
        const float ud = u * invD;
#pragma unroll
        for (int e = 0; e < 6; ++e) { Ud[e] = U[e] * invD; r[12 + e] = Ud[e]; }
        r[18] = ud;
        IA.sub_outer(U, Ud);
        float Ic[6];
        IA.apply(c, Ic);
#pragma unroll
        for (int e = 0; e < 6; ++e) P[e] += Ic[e] + U[e] * ud;
      }
      if (free_root) {
        // floating base: no joint force on any of the 6 directions, so  IA a + pA = 0
        const ChRef fr = shared_rec(m.s_free);
        float b[6], a[6];
#pragma unroll
        for (int e = 0; e < 6; ++e) b[e] = -P[e];
        ch_solve_spd6(IA, b, a);
        // a = a_base + c_J + S qacc   with S = [[0, R], [1, 0]]:  qacc_trans = a_lin - g - c_J,  qacc_rot = R^T a_ang
#pragma unroll
        for (int e = 0; e < 6; ++e) fr[12 + e] = a[e];
        fr[18] = a[3] - fr[9];
        fr[19] = a[4] - fr[10];
        fr[20] = a[5] - fr[11] - m.gravity;
        fr[21] = fr[0] * a[0] + fr[3] * a[1] + fr[6] * a[2];
        fr[22] = fr[1] * a[0] + fr[4] * a[1] + fr[7] * a[2];
        fr[23] = fr[2] * a[0] + fr[5] * a[1] + fr[8] * a[2];
      }
    }
  }

  // ---- planar pass 2 / pass 3: articulated inertia (Ayy, Byx, Byz, Cxx, Czz, Cxz), forces (n_y, f_x, f_z) -------------
  __host__ __device__ void pass2_planar(const float (&rec)[kChNodeRec]) {
    const ChainModel& m = *M;
    float Ayy = 0.f, Byx = 0.f, Byz = 0.f, Cxx = 0.f, Czz = 0.f, Cxz = 0.f, P1 = 0.f, P3 = 0.f, P5 = 0.f;
    const int n_trunk = m.n_trunk, n_junctions = m.n_junctions;
    const int n_seq = n_trunk + m.max_limb_nodes;
    const int last = m.seq_last[g];
    for (int i = n_seq - 1; i >= 0; --i) {
      if (i == n_trunk - 1 && n_junctions > 0) {
        const int mine = g < m.n_limbs ? m.trunk_junction[m.limb_attach[g]] : -1;
        for (int js = 0; js < n_junctions; ++js) {
          const bool w = mine == js;
          float x[9] = {w ? Ayy : 0.f, w ? Byx : 0.f, w ? Byz : 0.f, w ? Cxx : 0.f, w ? Czz : 0.f, w ? Cxz : 0.f,
                        w ? P1 : 0.f, w ? P3 : 0.f, w ? P5 : 0.f};
          ctx->group_sum(x);
          if (js == 0) ctx->group_sync();      // the junction region changes tenant: parked limb records -> sums
          const ChRef jr = shared_rec(m.s_jun + kChJun * js);
#pragma unroll
          for (int e = 0; e < 9; ++e) jr[e] = x[e];
        }
        Ayy = Byx = Byz = Cxx = Czz = Cxz = P1 = P3 = P5 = 0.f;
      }
      if (i == n_trunk - 1) ctx->group_sync();      // the trunk's node records were written by different lanes
      int node, pos;
      bool trunk;
      if (!node_at(i, n_trunk, node, trunk, pos)) continue;
      {
        float mass, hx, hz, Io, f1, f3, f5;
        if (i == last) {
          mass = rec[0]; hx = rec[1]; hz = rec[3]; Io = rec[5]; f1 = rec[11]; f3 = rec[13]; f5 = rec[15];
        } else {
          const ChRef nr = node_rec(trunk, pos);
          mass = nr[0]; hx = nr[1]; hz = nr[3]; Io = nr[5]; f1 = nr[11]; f3 = nr[13]; f5 = nr[15];
        }
        P1 += f1; P3 += f3; P5 += f5;
        Ayy += Io; Byx += hz; Byz -= hx; Cxx += mass; Czz += mass;       // ArtInertia::add_rigid
      }
      if (trunk && m.trunk_junction[i] >= 0) {
        const ChRef jr = shared_rec(m.s_jun + kChJun * m.trunk_junction[i]);
        Ayy += jr[0]; Byx += jr[1]; Byz += jr[2]; Cxx += jr[3]; Czz += jr[4]; Cxz += jr[5];
        P1 += jr[6]; P3 += jr[7]; P5 += jr[8];
      }
      const int j0 = m.seq_j0[g][i], j1 = m.seq_j1[g][i];
      const int rs = rec_step(trunk);
      float* rb = rec_base(trunk, m.seq_slot0[g][i]) + (j1 - 1 - j0) * rs;
      for (int j = j1 - 1; j >= j0; --j, rb -= rs) {
        const ChRef r{rb};
        const float S1 = r[1], S3 = r[3], S5 = r[5], c3 = r[9], c5 = r[11];
        // U = IA S
        const float U1 = Ayy * S1 + Byx * S3 + Byz * S5;
        const float U3 = Byx * S1 + Cxx * S3 + Cxz * S5;
        const float U5 = Byz * S1 + Cxz * S3 + Czz * S5;
        float D = r[trunk ? 20 : 18], u = r[trunk ? 19 : 17];
        D += S1 * U1 + S3 * U3 + S5 * U5;
        u -= S1 * P1 + S3 * P3 + S5 * P5;
        const float invD = ch_rcp(D);
        const float ud = u * invD;
        const float Ud1 = U1 * invD, Ud3 = U3 * invD, Ud5 = U5 * invD;
        r[13] = Ud1; r[15] = Ud3; r[17] = Ud5;
        r[18] = ud;
        // IA -= U Ud^T
        Ayy -= U1 * Ud1; Byx -= U1 * Ud3; Byz -= U1 * Ud5; Cxx -= U3 * Ud3; Czz -= U5 * Ud5; Cxz -= U3 * Ud5;
        // pA += IA c + U u / D   (c has no angular part)
        P1 += Byx * c3 + Byz * c5 + U1 * ud;
        P3 += Cxx * c3 + Cxz * c5 + U3 * ud;
        P5 += Cxz * c3 + Czz * c5 + U5 * ud;
      }
    }
  }

  template <class Sink>
  __host__ __device__ void pass3_planar(Sink&& sink) {
    const ChainModel& m = *M;
    float a1 = 0.f, a3 = 0.f, a5 = m.gravity;     // gravity as a fictitious base acceleration
    const int n_trunk = m.n_trunk;
    const int n_seq = n_trunk + m.max_limb_nodes;
    for (int i = 0; i < n_seq; ++i) {
      int node, pos;
      bool trunk;
      if (i == n_trunk) ctx->group_sync();     // the junction accelerations are complete in every lane's view
      if (!node_at(i, n_trunk, node, trunk, pos)) continue;
      if (!trunk && pos == 0) {
        const ChRef ar = shared_rec(m.s_acc + 6 * m.trunk_junction[m.limb_attach[g]]);
        a1 = ar[1]; a3 = ar[3]; a5 = ar[5];
      }
      const int j0 = m.seq_j0[g][i], j1 = m.seq_j1[g][i];
      float* rb = rec_base(trunk, m.seq_slot0[g][i]);
      const int rs = rec_step(trunk);
      for (int j = j0; j < j1; ++j, rb += rs) {
        const ChRef r{rb};
        a3 += r[9]; a5 += r[11];
        const float qacc = r[18] - (r[13] * a1 + r[15] * a3 + r[17] * a5);
        a1 += r[1] * qacc; a3 += r[3] * qacc; a5 += r[5] * qacc;
        sink(j, trunk, qacc);
      }
      if (trunk && m.trunk_junction[i] >= 0) {
        const ChRef ar = shared_rec(m.s_acc + 6 * m.trunk_junction[i]);
        ar[1] = a1; ar[3] = a3; ar[5] = a5;
      }
    }
  }

  // ---- pass 3: accelerations; qacc_j is handed to `sink(j, trunk, qacc)` in dof order ------------------------------
  template <class Sink>
  __host__ __device__ void pass3(Sink&& sink) {
    if constexpr (PLANAR) {
      pass3_planar(sink);
      return;
    }
    const ChainModel& m = *M;
    float a[6] = {0.f, 0.f, 0.f, 0.f, 0.f, m.gravity};     // gravity as a fictitious base acceleration
    const int n_trunk = m.n_trunk;
    const int n_seq = n_trunk + m.max_limb_nodes;
    for (int i = 0; i < n_seq; ++i) {
      int node, pos;
      bool trunk;
      if (i == n_trunk) ctx->group_sync();     // the junction accelerations are complete in every lane's view
      if (!node_at(i, n_trunk, node, trunk, pos)) continue;
      if (!trunk && pos == 0) {
        const ChRef ar = shared_rec(m.s_acc + 6 * m.trunk_junction[m.limb_attach[g]]);
#pragma unroll
        for (int e = 0; e < 6; ++e) a[e] = ar[e];
      }
      int j0 = m.seq_j0[g][i];
      const int j1 = m.seq_j1[g][i];
      if (i == 0 && m.root_free) {
        const ChRef fr = shared_rec(m.s_free);
#pragma unroll
        for (int e = 0; e < 6; ++e) a[e] = fr[12 + e];
#pragma unroll
        for (int e = 0; e < 6; ++e) sink(j0 + e, true, fr[18 + e]);
        j0 += 6;
      }
      float* rb = rec_base(trunk, m.seq_slot0[g][i]);
      const int rs = rec_step(trunk);
      for (int j = j0; j < j1; ++j, rb += rs) {
        const ChRef r{rb};
        float qacc = r[18];
#pragma unroll
        for (int e = 0; e < 6; ++e) { a[e] += r[6 + e]; qacc -= r[12 + e] * a[e]; }
#pragma unroll
        for (int e = 0; e < 6; ++e) a[e] += r[e] * qacc;
        sink(j, trunk, qacc);
      }
      if (trunk && m.trunk_junction[i] >= 0) {
        const ChRef ar = shared_rec(m.s_acc + 6 * m.trunk_junction[i]);
#pragma unroll
        for (int e = 0; e < 6; ++e) ar[e] = a[e];
      }
    }
  }

  // q (+)= step * rate for one dof.  The quaternion of a free joint moves when its last rotation dof comes by:
  // the caller collects the three body-frame rates in `w` (dof order) on the way.
  // Coordinate addresses need no table: a free joint, if any, is the first joint of the root (build_chain_model), so
  // dofs 0..2 are its translations (q 0..2), 3..5 its rotation rates (quaternion at q 3..6) and every later dof j has
  // q index j + 1; without a free joint q index == dof index.  (`root_free` is read once per substep by the caller: as
  // table reads, type and address were two dependent shared-memory loads in front of every state update.)
  __host__ __device__ __forceinline__ void advance_position(int j, int root_free, const ChRef& qsrc, const ChRef& qdst,
                                                            float rate, float (&w)[3], float step) const {
    if (!(root_free && j >= 3 && j < 6)) {
      const int qa = j + ((root_free && j >= 6) ? 1 : 0);
      qdst[qa] = qsrc[qa] + step * rate;
      return;
    }
    w[0] = w[1]; w[1] = w[2]; w[2] = rate;
    if (j < 5) return;
    const int qa = 3;
    const float n = sqrtf(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    const float half = 0.5f * step * n;
    float sn, cs;
    ch::sincos_(half, &sn, &cs);
    const float s = n > 1e-8f ? sn * ch_rcp(n) : 0.5f * step;
    const float bw = cs, bx = w[0] * s, by = w[1] * s, bz = w[2] * s;
    const float aw = qsrc[qa], ax = qsrc[qa + 1], ay = qsrc[qa + 2], az = qsrc[qa + 3];
    const float rw = aw * bw - ax * bx - ay * by - az * bz;
    const float rx = aw * bx + ax * bw + ay * bz - az * by;
    const float ry = aw * by - ax * bz + ay * bw + az * bx;
    const float rz = aw * bz + ax * by - ay * bx + az * bw;
    const float inv = ch::rsqrt_(rw * rw + rx * rx + ry * ry + rz * rz);
    qdst[qa] = rw * inv; qdst[qa + 1] = rx * inv; qdst[qa + 2] = ry * inv; qdst[qa + 3] = rz * inv;
  }

  // ---- one substep, semi-implicit Euler (oracle/articulated_np.py: integrate) -------------------------------------
  __host__ __device__ void substep_euler(const ChRef& ctrl) {
    const ChainModel& m = *M;
    const ChRef q = shared_rec(m.s_state), qd = shared_rec(m.s_state + m.nq);
    float rec[kChNodeRec];
    pass1(q, qd, ctrl, true, rec);
    pass2(rec);
    ctx->group_sync();       // the junction accelerations of pass 3 overwrite the junction sums pass 2 was reading
    const float dt = m.dt;
    const int gg = g, rf = m.root_free;
    const ChainLane* self = this;
    float w[3] = {0.f, 0.f, 0.f};
    pass3([&](int j, bool trunk, float qacc) {
      // trunk dofs are computed by every lane of the group (identical values): lane 0 alone updates the state
      if (trunk && gg != 0) return;
      const float v = qd[j] + dt * qacc;
      qd[j] = v;
      self->advance_position(j, rf, q, q, v, w, dt);
    });
    ctx->group_sync();
  }

  // ---- one substep, four-stage Runge-Kutta like mj_RungeKutta (oracle/articulated_np.py: rk4_substep) -------------
  // The substep's start state stays in the state slots; the stage state and the weighted sums of the stage
  // velocities / accelerations live in the s_rk slots: qs(nq) vs(nv) vsum(nv) asum(nv).
  __host__ __device__ void substep_rk4(const ChRef& ctrl) {
    const ChainModel& m = *M;
    const ChRef q0 = shared_rec(m.s_state), v0 = shared_rec(m.s_state + m.nq);
    const ChRef qs = shared_rec(m.s_rk), vs = shared_rec(m.s_rk + m.nq);
    const ChRef vsum = shared_rec(m.s_rk + m.nq + m.nv), asum = shared_rec(m.s_rk + m.nq + 2 * m.nv);
    const float dt = m.dt;
    const int gg = g, rf = m.root_free;
    const ChainLane* self = this;
    for (int stage = 0; stage < 4; ++stage) {
      float rec[kChNodeRec];
      // every stage evaluates the damped acceleration field of the Euler step (implicit joint springs / dampers)
      pass1(stage == 0 ? q0 : qs, stage == 0 ? v0 : vs, ctrl, true, rec);
      pass2(rec);
      ctx->group_sync();
      const float b = (stage == 0 || stage == 3) ? (1.f / 6.f) : (1.f / 3.f);
      const float nxt = dt * (stage == 2 ? 1.f : 0.5f);
      float w[3] = {0.f, 0.f, 0.f};
      pass3([&](int j, bool trunk, float qacc) {
        if (trunk && gg != 0) return;
        const float vcur = stage == 0 ? v0[j] : vs[j];
        const float vw = (stage == 0 ? 0.f : vsum[j]) + b * vcur;
        const float aw = (stage == 0 ? 0.f : asum[j]) + b * qacc;
        if (stage < 3) {
          vsum[j] = vw;
          asum[j] = aw;
          self->advance_position(j, rf, q0, qs, vcur, w, nxt);      // stage positions start from the substep's q0
          vs[j] = v0[j] + nxt * qacc;
        } else {
          self->advance_position(j, rf, q0, q0, vw, w, dt);
          v0[j] = v0[j] + dt * aw;
        }
      });
      ctx->group_sync();
    }
  }

  __host__ __device__ void step(const ChRef& ctrl) {
    if (M->integrator == kChRK4) {
      for (int s = 0; s < M->nsub; ++s) substep_rk4(ctrl);
    } else {
      for (int s = 0; s < M->nsub; ++s) substep_euler(ctrl);
    }
  }
};

// ---------------------------------------------------------------------------------------------------------------
// Host side: decompose a body tree (tables as in icem_articulated_model_t) into trunk + limb chains.
struct ChainSource {
  int nb, nq, nv, nu, nc, nsub, obs_offset, integrator;
  float dt, gravity, ctrl_limit, kc, cc, cdmax, kv, mu;
  const int32_t *body_parent, *body_dof_start, *body_dof_count;
  const float *body_pos, *body_mass, *body_com, *body_inertia;
  const int32_t *dof_type, *dof_qadr, *dof_limited, *dof_act;
  const float *dof_axis, *dof_anchor, *dof_stiffness, *dof_damping, *dof_armature, *dof_lo, *dof_hi, *dof_klim,
      *dof_blim, *dof_gear;
  const int32_t* con_body;
  const float *con_pos, *con_radius;
};

inline ChainSource chain_source(const icem_articulated_model_t& a) {
  ChainSource s{};
  s.nb = a.nb; s.nq = a.nq; s.nv = a.nv; s.nu = a.nu; s.nc = a.nc; s.nsub = a.nsub; s.obs_offset = a.obs_offset;
  s.integrator = a.integrator;
  s.dt = a.dt; s.gravity = a.gravity; s.ctrl_limit = a.ctrl_limit;
  s.kc = a.contact_stiffness; s.cc = a.contact_damping; s.cdmax = a.contact_damping_max;
  s.kv = a.friction_viscous; s.mu = a.friction;
  s.body_parent = a.body_parent; s.body_dof_start = a.body_dof_start; s.body_dof_count = a.body_dof_count;
  s.body_pos = a.body_pos; s.body_mass = a.body_mass; s.body_com = a.body_com; s.body_inertia = a.body_inertia;
  s.dof_type = a.dof_type; s.dof_qadr = a.dof_qadr; s.dof_limited = a.dof_limited; s.dof_act = a.dof_act;
  s.dof_axis = a.dof_axis; s.dof_anchor = a.dof_anchor; s.dof_stiffness = a.dof_stiffness;
  s.dof_damping = a.dof_damping; s.dof_armature = a.dof_armature; s.dof_lo = a.dof_lo; s.dof_hi = a.dof_hi;
  s.dof_klim = a.dof_klim; s.dof_blim = a.dof_blim; s.dof_gear = a.dof_gear;
  s.con_body = a.con_body; s.con_pos = a.con_pos; s.con_radius = a.con_radius;
  return s;
}

// false (with a reason) when the robot is not "one trunk chain + at most 4 limb chains": the caller falls back to the
// warp-per-trajectory engine of dyn_articulated.cuh.
inline bool build_chain_model(const ChainSource& a, int act_dim, ChainModel& m, const char** why) {
  static const char* none = "";
  *why = none;
  memset(&m, 0, sizeof m);
  int d_rec[kChMaxDofs] = {0}, d_slot[kChMaxDofs] = {0};     // record rank / first slot of every dof (host only)
  if (a.nb < 1 || a.nb > kChMaxNodes || a.nv > kChMaxDofs || a.nc > kChMaxCon) { *why = "too large"; return false; }
  if (a.body_dof_count[0] < 1) { *why = "root body without joints"; return false; }
  // ---- fuse jointless bodies into the node of their parent ---------------------------------------------------------
  int node_of[kChMaxNodes], n_nodes = 0, primary[kChMaxNodes];
  double off[kChMaxNodes][3];          // offset of the body frame inside its node's frame
  for (int b = 0; b < a.nb; ++b) {
    const int par = a.body_parent[b];
    if (par >= b || par < -1) { *why = "bodies not parent-before-child"; return false; }
    if (par >= 0 && a.body_dof_count[b] == 0) {
      node_of[b] = node_of[par];
      for (int k = 0; k < 3; ++k) off[b][k] = off[par][k] + a.body_pos[3 * b + k];
    } else {
      node_of[b] = n_nodes;
      primary[n_nodes++] = b;
      for (int k = 0; k < 3; ++k) off[b][k] = 0.0;
    }
  }
  int node_parent[kChMaxNodes], n_children[kChMaxNodes] = {0};
  for (int n = 0; n < n_nodes; ++n) {
    const int b = primary[n], par = a.body_parent[b];
    node_parent[n] = par < 0 ? -1 : node_of[par];
    if (n > 0 && par < 0) { *why = "more than one root"; return false; }
    if (node_parent[n] >= 0) n_children[node_parent[n]]++;
    for (int k = 0; k < 3; ++k) m.n_pos[n][k] = (float)(a.body_pos[3 * b + k] + (par >= 0 ? off[par][k] : 0.0));
    m.n_dof_start[n] = a.body_dof_start[b];
    m.n_dof_count[n] = a.body_dof_count[b];
  }
  // mass properties of a node: its bodies combined about the common centre of mass (no relative rotation)
  int n_con = 0;
  for (int n = 0; n < n_nodes; ++n) {
    double mass = 0, mc[3] = {0, 0, 0};
    for (int b = 0; b < a.nb; ++b)
      if (node_of[b] == n) {
        mass += a.body_mass[b];
        for (int k = 0; k < 3; ++k) mc[k] += a.body_mass[b] * (off[b][k] + a.body_com[3 * b + k]);
      }
    double com[3] = {mc[0] / mass, mc[1] / mass, mc[2] / mass};
    double I[6] = {0, 0, 0, 0, 0, 0};
    for (int b = 0; b < a.nb; ++b)
      if (node_of[b] == n) {
        const double mb = a.body_mass[b];
        double d[3];
        for (int k = 0; k < 3; ++k) d[k] = off[b][k] + a.body_com[3 * b + k] - com[k];
        const double d2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
        const float* Ib = a.body_inertia + 6 * b;
        I[0] += Ib[0] + mb * (d2 - d[0] * d[0]);
        I[1] += Ib[1] + mb * (d2 - d[1] * d[1]);
        I[2] += Ib[2] + mb * (d2 - d[2] * d[2]);
        I[3] += Ib[3] - mb * d[0] * d[1];
        I[4] += Ib[4] - mb * d[0] * d[2];
        I[5] += Ib[5] - mb * d[1] * d[2];
      }
    m.n_mass[n] = (float)mass;
    for (int k = 0; k < 3; ++k) m.n_com[n][k] = (float)com[k];
    for (int k = 0; k < 6; ++k) m.n_inertia[n][k] = (float)I[k];
    m.n_con_start[n] = n_con;
    for (int c = 0; c < a.nc; ++c) {
      const int b = a.con_body[c];
      if (b < 0 || b >= a.nb) { *why = "contact body out of range"; return false; }
      if (node_of[b] != n) continue;
      for (int k = 0; k < 3; ++k) m.c_pos[n_con][k] = (float)(off[b][k] + a.con_pos[3 * c + k]);
      m.c_radius[n_con] = a.con_radius[c];
      ++n_con;
    }
    m.n_con_count[n] = n_con - m.n_con_start[n];
  }
  // ---- trunk: follow the one child whose subtree branches; every other child must be a plain chain = a limb ------
  bool is_chain[kChMaxNodes];
  for (int n = n_nodes - 1; n >= 0; --n) {
    is_chain[n] = n_children[n] <= 1;
    for (int c = n + 1; c < n_nodes; ++c)
      if (node_parent[c] == n && !is_chain[c]) is_chain[n] = false;
  }
  int t = 0;
  m.n_trunk = 0;
  m.n_limbs = 0;
  for (int k = 0; k < kChMaxTrunk; ++k) m.trunk_junction[k] = -1;
  while (true) {
    if (m.n_trunk >= kChMaxTrunk) { *why = "trunk longer than 4 bodies"; return false; }
    const int pos = m.n_trunk;
    m.trunk_node[m.n_trunk++] = t;
    int next = -1;
    for (int c = t + 1; c < n_nodes; ++c) {
      if (node_parent[c] != t) continue;
      if (!is_chain[c]) {
        if (next >= 0) { *why = "two branching subtrees below one body"; return false; }
        next = c;
        continue;
      }
      if (m.n_limbs >= kChMaxLimbs) { *why = "more than 4 limbs"; return false; }
      const int l = m.n_limbs++;
      m.limb_attach[l] = pos;
      int cur = c, cnt = 0;
      while (cur >= 0) {
        if (cnt >= kChMaxLimbNodes) { *why = "limb longer than 4 bodies"; return false; }
        m.limb_node[l][cnt++] = cur;
        int nxt = -1;
        for (int c2 = cur + 1; c2 < n_nodes; ++c2)
          if (node_parent[c2] == cur) nxt = c2;
        cur = nxt;
      }
      m.limb_nnodes[l] = cnt;
      if (m.trunk_junction[pos] < 0) m.trunk_junction[pos] = m.n_junctions++;
    }
    if (next < 0) break;
    t = next;
  }
  m.lanes = m.n_limbs <= 1 ? 1 : (m.n_limbs <= 2 ? 2 : 4);
  for (int g = 0; g < kChMaxLimbs; ++g) m.seq_last[g] = -1;
  for (int g = 0; g < kChMaxLimbs; ++g)
    for (int i = 0; i < kChMaxTrunk + kChMaxLimbNodes; ++i) {
      int node = -1;
      if (i < m.n_trunk) node = m.trunk_node[i];
      else if (g < m.n_limbs && i - m.n_trunk < m.limb_nnodes[g]) node = m.limb_node[g][i - m.n_trunk];
      m.seq_node[g][i] = node;
      if (node >= 0 && i >= m.n_trunk) m.seq_last[g] = i;      // a trunk node never is "last": its record is shared
    }
  // ---- dofs -----------------------------------------------------------------------------------------------------
  for (int j = 0; j < a.nv; ++j) {
    m.d_type[j] = a.dof_type[j]; m.d_limited[j] = a.dof_limited[j];
    m.d_act[j] = a.dof_act[j];
    if (m.d_act[j] >= act_dim) { *why = "actuator index out of range"; return false; }
    for (int k = 0; k < 3; ++k) { m.d_axis[j][k] = a.dof_axis[3 * j + k]; m.d_anchor[j][k] = a.dof_anchor[3 * j + k]; }
    m.d_stiff[j] = a.dof_stiffness[j]; m.d_damp[j] = a.dof_damping[j]; m.d_arm[j] = a.dof_armature[j];
    m.d_lo[j] = a.dof_lo[j]; m.d_hi[j] = a.dof_hi[j]; m.d_klim[j] = a.dof_klim[j]; m.d_blim[j] = a.dof_blim[j];
    m.d_gear[j] = a.dof_gear[j];
  }
  m.trunk_dofs = 0;
  for (int k = 0; k < m.n_trunk; ++k) {
    const int n = m.trunk_node[k];
    bool seen_hinge = false;
    for (int j = m.n_dof_start[n]; j < m.n_dof_start[n] + m.n_dof_count[n]; ++j) {
      const int ty = m.d_type[j];
      if (ty == kFreeTrans || ty == kFreeRot) {
        // the free root is solved as a floating base: nothing may act on its six dofs
        if (m.d_arm[j] != 0.f || m.d_damp[j] != 0.f || m.d_stiff[j] != 0.f || m.d_act[j] >= 0 || m.d_limited[j]) {
          *why = "free joint with armature / damping / actuator";
          return false;
        }
        m.root_free = 1;
        d_rec[j] = 0;
      } else {
        d_rec[j] = m.trunk_dofs++;
      }
      if ((ty == kFreeTrans || ty == kFreeRot) && k != 0) { *why = "free joint below the root"; return false; }
      if (k == 0 && ty == kHinge) seen_hinge = true;
      if (k == 0 && ty == kSlide && seen_hinge) { *why = "root slide after a root hinge"; return false; }
      if (ty == kFreeTrans && j != m.n_dof_start[n] && m.d_type[j - 1] != kFreeTrans) {
        *why = "free joint must come first on the root";
        return false;
      }
      if (ty == kFreeTrans && j == m.n_dof_start[n]) {
        if (j + 5 >= m.n_dof_start[n] + m.n_dof_count[n]) { *why = "incomplete free joint"; return false; }
        for (int e = 0; e < 6; ++e)
          if (m.d_type[j + e] != (e < 3 ? kFreeTrans : kFreeRot)) { *why = "malformed free joint"; return false; }
      }
    }
  }
  m.max_limb_nodes = 0;
  m.max_limb_dofs = 0;
  for (int l = 0; l < m.n_limbs; ++l) {
    int cnt = 0;
    for (int k = 0; k < m.limb_nnodes[l]; ++k) {
      const int n = m.limb_node[l][k];
      for (int j = m.n_dof_start[n]; j < m.n_dof_start[n] + m.n_dof_count[n]; ++j) {
        if (m.d_type[j] != kSlide && m.d_type[j] != kHinge) { *why = "free joint on a limb"; return false; }
        d_rec[j] = cnt++;
      }
    }
    m.max_limb_dofs = cnt > m.max_limb_dofs ? cnt : m.max_limb_dofs;
    m.max_limb_nodes = m.limb_nnodes[l] > m.max_limb_nodes ? m.limb_nnodes[l] : m.max_limb_nodes;
  }
  m.nq = a.nq; m.nv = a.nv; m.nu = a.nu; m.nsub = a.nsub; m.obs_offset = a.obs_offset; m.n_nodes = n_nodes;
  if (a.integrator != kChEuler && a.integrator != kChRK4) { *why = "unknown integrator"; return false; }
  m.integrator = a.integrator;
  m.dt = a.dt; m.gravity = a.gravity; m.ctrl_limit = a.ctrl_limit;
  m.kc = a.kc; m.cc = a.cc; m.kv = a.kv; m.mu = a.mu; m.cdmax = a.cdmax;
  // ---- scratch layout ---------------------------------------------------------------------------------------------
  int s = 0;
  m.s_state = s; s += m.nq + m.nv;
  m.s_ctrl = s; s += act_dim;                      // the controls of the current control step
  m.s_free = s; s += m.root_free ? kChFreeRec : 0;
  m.s_dof = s; s += kChTrunkDofRec * m.trunk_dofs;
  m.s_node = s; s += kChNodeRec * m.n_trunk;
  // one region, four tenants in turn: trunk frames (pass 1) -> the limbs' first node records (until their pass 2) ->
  // junction sums (pass 2) -> junction accelerations (pass 3); the lanes are synchronised between the tenants
  // (group_sync / group sum)
  m.s_frame = s; m.s_jun = s; m.s_acc = s;
  {
    int region = kChJun * m.n_junctions > kChFrame * m.n_trunk ? kChJun * m.n_junctions : kChFrame * m.n_trunk;
    if (m.max_limb_nodes >= 2 && kChNodeRec * m.lanes > region) region = kChNodeRec * m.lanes;   // parked first records
    s += region;
  }
  m.s_rk = s; s += m.integrator == kChRK4 ? m.nq + 3 * m.nv : 0;
  m.s_end = s;
  int p = 0;
  m.p_dof = p; p += kChDofRec * m.max_limb_dofs;
  // a limb's last node keeps its record in registers (pass 1 -> pass 2), its first one is parked in the junction region
  m.p_node = p; p += kChNodeRec * (m.max_limb_nodes > 2 ? m.max_limb_nodes - 2 : 0);
  m.p_end = p;
  // the engine addresses coordinates without a table (advance_position, walk_joints): q index == dof index, shifted by
  // one behind a free joint (4 quaternion entries for 3 rotation dofs)
  for (int j = 0; j < m.nv; ++j) {
    const int want = m.d_type[j] == kFreeRot ? 3 : j + ((m.root_free && j >= 6) ? 1 : 0);
    if (a.dof_qadr[j] != want) { *why = "unexpected coordinate layout"; return false; }
  }
  {
    bool is_trunk_dof[kChMaxDofs] = {false};
    for (int t = 0; t < m.n_trunk; ++t) {
      const int n = m.trunk_node[t];
      for (int j = m.n_dof_start[n]; j < m.n_dof_start[n] + m.n_dof_count[n]; ++j) is_trunk_dof[j] = true;
    }
    for (int j = 0; j < m.nv; ++j)
      d_slot[j] = is_trunk_dof[j] ? m.s_dof + kChTrunkDofRec * d_rec[j] : m.p_dof + kChDofRec * d_rec[j];
    for (int n = 0; n < m.n_nodes; ++n) {
      const int j1 = m.n_dof_start[n] + m.n_dof_count[n];
      const int jf = m.n_dof_start[n] + ((n == m.trunk_node[0] && m.root_free) ? 6 : 0);
      m.n_slot0[n] = jf < j1 ? d_slot[jf] : 0;
      for (int j = jf; j < j1; ++j)
        if (d_slot[j] != m.n_slot0[n] + (j - jf) * (is_trunk_dof[j] ? kChTrunkDofRec : kChDofRec)) {
          *why = "dof records of a node are not consecutive";
          return false;
        }
    }
    for (int g2 = 0; g2 < kChMaxLimbs; ++g2)
      for (int i = 0; i < kChMaxTrunk + kChMaxLimbNodes; ++i) {
        const int n = m.seq_node[g2][i];
        m.seq_j0[g2][i] = n >= 0 ? m.n_dof_start[n] : 0;
        m.seq_j1[g2][i] = n >= 0 ? m.n_dof_start[n] + m.n_dof_count[n] : 0;
        m.seq_slot0[g2][i] = n >= 0 ? m.n_slot0[n] : 0;
      }
  }
  // planar robots (HalfCheetah, Hopper): no free joint, slides inside the x-z plane, hinges about +-y, every offset /
  // centre of mass / anchor / contact point at y = 0, no inertia product with y
  m.planar = m.root_free ? 0 : 1;
  for (int n = 0; n < m.n_nodes && m.planar; ++n)
    if (m.n_pos[n][1] != 0.f || m.n_com[n][1] != 0.f || m.n_inertia[n][3] != 0.f || m.n_inertia[n][5] != 0.f) m.planar = 0;
  for (int j = 0; j < m.nv && m.planar; ++j) {
    if (m.d_type[j] == kSlide) { if (m.d_axis[j][1] != 0.f) m.planar = 0; }
    else if (m.d_type[j] == kHinge) {
      if (m.d_axis[j][0] != 0.f || m.d_axis[j][2] != 0.f || fabsf(m.d_axis[j][1]) != 1.f || m.d_anchor[j][1] != 0.f) m.planar = 0;
    } else m.planar = 0;
  }
  for (int c = 0; c < n_con && m.planar; ++c)
    if (m.c_pos[c][1] != 0.f) m.planar = 0;
  return true;
}

// Scratch of one warp: the group-shared region ([slot][group]) followed by one lane-private region per lane index g
// inside the group ([g][slot][group]); each private region is padded so that the G regions start 32 / G banks apart
// (all 32 lanes of the warp then hit 32 different banks when they touch the same slot).
__host__ __device__ inline int chain_private_region_floats(int p_end, int lanes) {
  const int ng = 32 / lanes;
  int f = p_end * ng;
  while ((f & 31) != (ng & 31)) ++f;
  return f;
}
inline int chain_warp_floats(const ChainModel& m) {
  return m.s_end * (32 / m.lanes) + m.lanes * chain_private_region_floats(m.p_end, m.lanes);
}

}  // namespace icem
