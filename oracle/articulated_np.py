"""Float64 NumPy restatement of the articulated-body ground-truth model.  TEST INFRASTRUCTURE ONLY.

Role: the reference's ground-truth rollouts are `GroundTruthModel.predict_n_steps` -> `env.step` -> MuJoCo
(icem/models/gt_model.py:76-102, icem/environments/mujoco.py:101-131); MuJoCo is unavailable here, so
the "ground truth" of this repo is its own soft-contact rigid-body model on the tables of icem_b200/robots.py.
PARITY WITH MUJOCO IS UNPINNED.  This file is the independent checker of the CUDA engine
(icem_b200/csrc/dyn_articulated.cuh): same physical model and integrator, DIFFERENT formulation --

  CUDA   : spatial vectors about the root position, recursive Newton-Euler bias + composite-rigid-body mass matrix,
           in-warp Cholesky, float32
  here   : per-body Jacobians at the centres of mass, M = sum J^T I J, bias from analytic Jacobian time
           derivatives, LAPACK solve, float64, vectorised over trajectories

so agreement (tests/test_gpu_articulated.py) plus the physical invariants checked on CPU (tests/test_articulated.py:
momentum / energy conservation in free flight, pendulum period, static equilibrium on the floor) validate both.

Model, per substep of length dt (frame_skip substeps per control step, control held):
  tau      = gear * clip(u) - K (q - q_ref)       K = joint stiffness (+ limit spring while a limit is violated)
  f_c      = soft floor contact at sphere points: fn = clamp(k pen - min(k pen c, d_max) vz, 0, 3 k pen), viscous
             friction capped at mu * fn
  (M + dt B + dt^2 K) qacc = tau - (B + dt K) qd - bias(q, qd) + J_c^T f_c
             B = joint damping (+ limit damper while violated): springs and dampers implicit (MuJoCo's Euler is
             implicit in joint damping the same way), contacts explicit
  qd <- qd + dt qacc ;  q <- q (+) dt qd                         (semi-implicit Euler; unit quaternion for the root)

`integrator="rk4"` (what gym's half_cheetah.xml / humanoidstandup.xml ask of MuJoCo): the classic
four-stage Runge-Kutta step laid out like mj_RungeKutta -- FOUR dynamics evaluations per substep, stage positions
integrated from the substep's start position (quaternion-aware), the final step from the weighted stage velocities /
accelerations (1/6, 1/3, 1/3, 1/6).  Every stage evaluates the SAME damped acceleration field as the Euler step,
(M + dt B + dt^2 K)^-1 (tau - (B + dt K) qd - bias): MuJoCo keeps its stiff limit / contact terms stable inside the
constraint solver; this soft-constraint model needs the implicit joint diagonal for that (with purely explicit
springs and dampers the limit spring-dampers of HumanoidStandup blow up within a few control steps).
"""
import numpy as np

from icem_b200.robots import FREE_ROT, FREE_TRANS, HINGE, SLIDE, CompiledModel, get_model   # data tables only


def _cross(a, b):
    return np.cross(a, b)


def _rodrigues(axis, angle):
    """Rotation matrices [P,3,3] about unit `axis` [P,3] by `angle` [P]."""
    x, y, z = axis[..., 0], axis[..., 1], axis[..., 2]
    c, s = np.cos(angle), np.sin(angle)
    C = 1 - c
    R = np.empty(axis.shape[:-1] + (3, 3))
    R[..., 0, 0] = c + x * x * C
    R[..., 0, 1] = x * y * C - z * s
    R[..., 0, 2] = x * z * C + y * s
    R[..., 1, 0] = y * x * C + z * s
    R[..., 1, 1] = c + y * y * C
    R[..., 1, 2] = y * z * C - x * s
    R[..., 2, 0] = z * x * C - y * s
    R[..., 2, 1] = z * y * C + x * s
    R[..., 2, 2] = c + z * z * C
    return R


def quat_to_mat(q):
    """Rotation matrix of the NORMALISED quaternion (MuJoCo normalises quaternions before using them, so a state
    with gym's reset noise on qpos[3:7] is still a valid orientation)."""
    q = q / np.linalg.norm(q, axis=-1, keepdims=True)
    w, x, y, z = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    R = np.empty(q.shape[:-1] + (3, 3))
    R[..., 0, 0] = 1 - 2 * (y * y + z * z)
    R[..., 0, 1] = 2 * (x * y - w * z)
    R[..., 0, 2] = 2 * (x * z + w * y)
    R[..., 1, 0] = 2 * (x * y + w * z)
    R[..., 1, 1] = 1 - 2 * (x * x + z * z)
    R[..., 1, 2] = 2 * (y * z - w * x)
    R[..., 2, 0] = 2 * (x * z - w * y)
    R[..., 2, 1] = 2 * (y * z + w * x)
    R[..., 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def quat_mul(a, b):
    aw, ax, ay, az = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bw, bx, by, bz = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([aw * bw - ax * bx - ay * by - az * bz,
                     aw * bx + ax * bw + ay * bz - az * by,
                     aw * by - ax * bz + ay * bw + az * bx,
                     aw * bz + ax * by - ay * bx + az * bw], axis=-1)


class ArticulatedModel:
    """`rollout(start_state, actions[p,h,d]) -> observations[p,h,obs_dim]` like oracle/dynamics_np.py models."""

    def __init__(self, m: CompiledModel, obs_skip=0, integrator=None):
        self.m = m
        self.integrator = integrator or getattr(m, "integrator", "euler")
        assert self.integrator in ("euler", "rk4")
        self.obs_skip = obs_skip             # HalfCheetah observation drops qpos[0] (exclude_current_positions)
        self.state_dim = m.nq + m.nv
        self.obs_dim = self.state_dim - obs_skip
        self.act_dim = m.nu
        nb, nv = m.nb, m.nv
        # chain[b, j]: dof j moves body b
        chain = np.zeros((nb, nv), bool)
        for b in range(nb):
            a = b
            while a >= 0:
                s, c = m.body_dof_start[a], m.body_dof_count[a]
                chain[b, s:s + c] = True
                a = m.body_parent[a]
        self.chain = chain
        self.isrot = np.isin(m.dof_type, (HINGE, FREE_ROT))
        # before[j, i]: dof i contributes to the velocity of the frame in which axis j is fixed
        before = np.zeros((nv, nv), bool)
        for j in range(nv):
            i = m.dof_parent[j]
            while i >= 0:
                before[j, i] = True
                i = m.dof_parent[i]
            if m.dof_type[j] == FREE_ROT:      # body-fixed axes of the ball part of a free joint
                for i in range(nv):
                    if m.dof_type[i] == FREE_ROT and m.dof_body[i] == m.dof_body[j]:
                        before[j, i] = True
        self.before = before
        I = m.body_inertia
        self.I_body = np.stack([np.stack([I[:, 0], I[:, 3], I[:, 4]], -1),
                                np.stack([I[:, 3], I[:, 1], I[:, 5]], -1),
                                np.stack([I[:, 4], I[:, 5], I[:, 2]], -1)], -2)     # [nb,3,3]
        self.max_chunk = 2048

    # ---- kinematics ------------------------------------------------------------------------------------------
    def kinematics(self, q):
        m = self.m
        P = q.shape[0]
        Rb = np.empty((P, m.nb, 3, 3))
        pb = np.empty((P, m.nb, 3))
        aw = np.zeros((P, m.nv, 3))
        xw = np.zeros((P, m.nv, 3))
        eye = np.broadcast_to(np.eye(3), (P, 3, 3))
        for b in range(m.nb):
            par = m.body_parent[b]
            if par >= 0:
                R = Rb[:, par].copy()
                p = pb[:, par] + np.einsum("pij,j->pi", Rb[:, par], m.body_pos[b])
            else:
                R = eye.copy()
                p = np.broadcast_to(m.body_pos[b], (P, 3)).copy()
            j = m.body_dof_start[b]
            end = j + m.body_dof_count[b]
            while j < end:
                t = m.dof_type[j]
                if t == FREE_TRANS:            # free joint: 3 translations + ball (quaternion), handled as a block
                    qa = m.dof_qadr[j]
                    p = q[:, qa:qa + 3].copy()
                    R = quat_to_mat(q[:, qa + 3:qa + 7])
                    for a in range(3):
                        aw[:, j + a, a] = 1.0
                        xw[:, j + a] = p
                        aw[:, j + 3 + a] = R[:, :, a]
                        xw[:, j + 3 + a] = p
                    j += 6
                    continue
                axis_w = np.einsum("pij,j->pi", R, m.dof_axis[j])
                if t == SLIDE:
                    xw[:, j] = p
                    p = p + axis_w * q[:, m.dof_qadr[j], None]
                else:
                    anchor = p + np.einsum("pij,j->pi", R, m.dof_anchor[j])
                    Rj = _rodrigues(axis_w, q[:, m.dof_qadr[j]])
                    R = np.einsum("pij,pjk->pik", Rj, R)
                    p = anchor + np.einsum("pij,pj->pi", Rj, p - anchor)
                    xw[:, j] = anchor
                aw[:, j] = axis_w
                j += 1
            Rb[:, b], pb[:, b] = R, p
        return Rb, pb, aw, xw

    # ---- one dynamics evaluation ---------------------------------------------------------------------------
    def qacc(self, q, qd, ctrl, return_parts=False, implicit=True):
        m = self.m
        P = q.shape[0]
        Rb, pb, aw, xw = self.kinematics(q)
        cb = pb + np.einsum("pbij,bj->pbi", Rb, m.body_com)                       # com world [P,nb,3]
        Iw = np.einsum("pbij,bjk,pblk->pbil", Rb, self.I_body, Rb)                 # [P,nb,3,3]
        C = self.chain[None, :, :, None]                                             # [1,nb,nv,1]
        rot = self.isrot[None, None, :, None]
        a4 = aw[:, None]                                                             # [P,1,nv,3]
        Jw = np.where(C & rot, a4, 0.0)                                              # [P,nb,nv,3]
        Jv = np.where(C, np.where(rot, _cross(a4, cb[:, :, None] - xw[:, None]), a4), 0.0)
        M = (np.einsum("b,pbia,pbja->pij", m.body_mass, Jv, Jv)
             + np.einsum("pbia,pbac,pbjc->pij", Jw, Iw, Jw))
        M[:, np.arange(m.nv), np.arange(m.nv)] += m.dof_armature
        wb = np.einsum("pbja,pj->pba", Jw, qd)
        vcb = np.einsum("pbja,pj->pba", Jv, qd)
        # velocity of the frame each axis is fixed in, and of each anchor point moving with it
        aq = aw * qd[:, :, None]
        wf = np.einsum("ji,pia->pja", (self.before & self.isrot[None, :]).astype(float), aq)
        colx = np.where(self.isrot[None, None, :, None], _cross(aw[:, None], xw[:, :, None] - xw[:, None]), aw[:, None])
        vx = np.einsum("ji,pjia,pi->pja", self.before.astype(float), colx, qd)       # [P,nv,3]
        adot = _cross(wf, aw)
        Jwdq = np.einsum("bj,pja,pj->pba", (self.chain & self.isrot[None, :]).astype(float), adot, qd)
        term = np.where(rot, _cross(adot[:, None], cb[:, :, None] - xw[:, None])
                        + _cross(a4, vcb[:, :, None] - vx[:, None]), adot[:, None])
        Jvdq = np.einsum("bj,pbja,pj->pba", self.chain.astype(float), term, qd)
        grav = np.array([0.0, 0.0, m.gravity])
        lin = m.body_mass[None, :, None] * (Jvdq + grav)
        ang = np.einsum("pbij,pbj->pbi", Iw, Jwdq) + _cross(wb, np.einsum("pbij,pbj->pbi", Iw, wb))
        bias = np.einsum("pbja,pba->pj", Jv, lin) + np.einsum("pbja,pba->pj", Jw, ang)

        # ---- applied forces ---------------------------------------------------------------------------------
        u = np.clip(ctrl, -m.ctrl_limit, m.ctrl_limit)
        tau = np.zeros((P, m.nv))
        act = m.dof_act >= 0
        tau[:, act] = m.dof_gear[act] * u[:, m.dof_act[act]]
        simple = np.isin(m.dof_type, (SLIDE, HINGE))
        qj = np.zeros((P, m.nv))
        qj[:, simple] = q[:, m.dof_qadr[simple]]
        tau -= m.dof_stiffness * qj
        lim = m.dof_limited.astype(bool)[None, :]
        below = lim & (qj < m.dof_lo)
        above = lim & (qj > m.dof_hi)
        tau += np.where(below, m.dof_klim * (m.dof_lo - qj), 0.0)
        tau += np.where(above, m.dof_klim * (m.dof_hi - qj), 0.0)
        active = below | above
        # springs and dampers (joint + active limit) are integrated implicitly: diagonal terms of the system matrix
        Keff = m.dof_stiffness + np.where(active, m.dof_klim, 0.0)
        Beff = m.dof_damping + np.where(active, m.dof_blim, 0.0)
        # ---- floor contacts ---------------------------------------------------------------------------------
        xc = pb[:, m.con_body] + np.einsum("pcij,cj->pci", Rb[:, m.con_body], m.con_pos)     # [P,nc,3]
        Cc = self.chain[m.con_body][None, :, :, None]
        Jc = np.where(Cc, np.where(rot, _cross(a4, xc[:, :, None] - xw[:, None]), a4), 0.0)  # [P,nc,nv,3]
        uc = np.einsum("pcja,pj->pca", Jc, qd)
        pen = m.con_radius[None, :] - xc[..., 2]
        spring = m.contact_stiffness * pen
        damp = np.minimum(spring * m.contact_damping, m.contact_damping_max)
        fn = np.where(pen > 0, np.clip(spring - damp * uc[..., 2], 0.0, 3.0 * spring), 0.0)
        speed = np.sqrt(uc[..., 0] ** 2 + uc[..., 1] ** 2)
        coef = np.minimum(m.friction_viscous, m.friction * fn / np.maximum(speed, 1e-6))
        coef = np.where(pen > 0, coef, 0.0)
        fc = np.stack([-coef * uc[..., 0], -coef * uc[..., 1], fn], axis=-1)
        tau_c = np.einsum("pcja,pca->pj", Jc, fc)

        hdt = m.dt if implicit else 0.0
        rhs = tau + tau_c - (Beff + hdt * Keff) * qd - bias
        A = M.copy()
        A[:, np.arange(m.nv), np.arange(m.nv)] += hdt * Beff + hdt * hdt * Keff
        acc = np.linalg.solve(A, rhs[..., None])[..., 0]
        if return_parts:
            return dict(qacc=acc, M=M, bias=bias, tau=tau, tau_contact=tau_c, Rb=Rb, pb=pb, cb=cb, wb=wb, vcb=vcb,
                        Iw=Iw, fn=fn, xc=xc)
        return acc

    def integrate_pos(self, q, qd, dt):
        """q (+) dt * qd (unit quaternion for the rotation part of a free joint)."""
        m = self.m
        qn = q.copy()
        for j in range(m.nv):
            t = m.dof_type[j]
            if t in (SLIDE, HINGE, FREE_TRANS):
                qn[:, m.dof_qadr[j]] = q[:, m.dof_qadr[j]] + dt * qd[:, j]
        for j in range(m.nv):
            if m.dof_type[j] == FREE_ROT and (j == 0 or m.dof_type[j - 1] != FREE_ROT):
                qa = m.dof_qadr[j]
                w = qd[:, j:j + 3]
                n = np.linalg.norm(w, axis=-1)
                half = 0.5 * dt * n
                s = np.where(n > 1e-8, np.sin(half) / np.maximum(n, 1e-30), 0.5 * dt)
                dq = np.concatenate([np.cos(half)[:, None], w * s[:, None]], axis=-1)
                qq = quat_mul(q[:, qa:qa + 4], dq)
                qn[:, qa:qa + 4] = qq / np.linalg.norm(qq, axis=-1, keepdims=True)
        return qn

    def rk4_substep(self, q0, v0, ctrl):
        """mj_RungeKutta(N = 4): A = diag(1/2, 1/2, 1), B = (1/6, 1/3, 1/3, 1/6)."""
        dt = self.m.dt
        b = (1.0 / 6.0, 1.0 / 3.0, 1.0 / 3.0, 1.0 / 6.0)
        nxt = (0.5, 0.5, 1.0)
        q, v = q0, v0
        vsum = np.zeros_like(v0)
        asum = np.zeros_like(v0)
        for i in range(4):
            a = self.qacc(q, v, ctrl, implicit=True)
            vsum = vsum + b[i] * v
            asum = asum + b[i] * a
            if i < 3:
                q_next = self.integrate_pos(q0, v, dt * nxt[i])
                v = v0 + dt * nxt[i] * a
                q = q_next
        return self.integrate_pos(q0, vsum, dt), v0 + dt * asum

    def integrate(self, q, qd, acc):
        m = self.m
        dt = m.dt
        qd = qd + dt * acc
        qn = q.copy()
        for j in range(m.nv):
            t = m.dof_type[j]
            if t in (SLIDE, HINGE, FREE_TRANS):
                qn[:, m.dof_qadr[j]] = q[:, m.dof_qadr[j]] + dt * qd[:, j]
        for j in range(m.nv):
            if m.dof_type[j] == FREE_ROT and (j == 0 or m.dof_type[j - 1] != FREE_ROT):
                qa = m.dof_qadr[j]
                w = qd[:, j:j + 3]
                n = np.linalg.norm(w, axis=-1)
                half = 0.5 * dt * n
                s = np.where(n > 1e-8, np.sin(half) / np.maximum(n, 1e-30), 0.5 * dt)
                dq = np.concatenate([np.cos(half)[:, None], w * s[:, None]], axis=-1)
                qq = quat_mul(q[:, qa:qa + 4], dq)
                qn[:, qa:qa + 4] = qq / np.linalg.norm(qq, axis=-1, keepdims=True)
        return qn, qd

    # ---- env.step equivalent: frame_skip substeps with the control held ---------------------------------------
    def step_state(self, state, action):
        """state [P, nq+nv], action [P, nu] -> next state."""
        m = self.m
        q, qd = state[:, :m.nq].copy(), state[:, m.nq:].copy()
        for _ in range(m.nsub):
            if self.integrator == "rk4":
                q, qd = self.rk4_substep(q, qd, action)
            else:
                acc = self.qacc(q, qd, action)
                q, qd = self.integrate(q, qd, acc)
        return np.concatenate([q, qd], axis=-1)

    def observe(self, state):
        return np.asarray(state, np.float64)[..., self.obs_skip:]

    def step(self, obs, act):
        raise NotImplementedError("articulated models step on the full state (step_state), not the observation")

    def rollout(self, start_state, actions, with_final=False):
        """PRE-action observation of every step: [p,h,obs_dim]; `with_final` appends the observation
        after the last action ([p,h+1,obs_dim]) for costs that read next_obs (Hopper / Ant)."""
        p, h, _ = actions.shape
        out = np.empty((p, h + (1 if with_final else 0), self.obs_dim))
        for lo in range(0, p, self.max_chunk):
            hi = min(p, lo + self.max_chunk)
            st = np.broadcast_to(np.asarray(start_state, np.float64), (hi - lo, self.state_dim)).copy()
            for t in range(h):
                out[lo:hi, t] = st[:, self.obs_skip:]
                if t + 1 < h or with_final:
                    st = self.step_state(st, np.asarray(actions[lo:hi, t], np.float64))
            if with_final:
                out[lo:hi, h] = st[:, self.obs_skip:]
        return out

    # ---- diagnostics for the physics tests ---------------------------------------------------------------------
    def energy(self, q, qd):
        parts = self.qacc(q, qd, np.zeros((q.shape[0], self.m.nu)), return_parts=True)
        m = self.m
        kin = 0.5 * np.einsum("pi,pij,pj->p", qd, parts["M"], qd)
        pot = np.einsum("b,pb->p", m.body_mass, parts["cb"][..., 2]) * m.gravity
        simple = np.isin(m.dof_type, (SLIDE, HINGE))
        qj = q[:, m.dof_qadr[simple]]
        pot = pot + 0.5 * np.sum(m.dof_stiffness[simple] * qj ** 2, axis=-1)
        return kin, pot


def make_model(name, obs_skip=None, integrator=None) -> ArticulatedModel:
    if obs_skip is None:      # HalfCheetah's 17-wide observation drops qpos[0]; every other env keeps the full state
        obs_skip = 1 if name == "halfcheetah" else 0
    return ArticulatedModel(get_model(name), obs_skip=obs_skip, integrator=integrator)
