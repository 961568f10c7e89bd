"""Object wrapper over one `icem_planner_t` handle (include/icem_b200.h).  Host buffers are NumPy."""
import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _lib
from ._lib import IcemError, check, dptr, f32, f64, fptr, iptr  # noqa: F401


def articulated_model_struct(m, obs_offset=0):
    """robots.CompiledModel -> (icem_articulated_model_t, arrays that must outlive the call)."""
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    keep = dict(
        body_parent=i32(m.body_parent), body_dof_start=i32(m.body_dof_start), body_dof_count=i32(m.body_dof_count),
        body_pos=f32(m.body_pos), body_mass=f32(m.body_mass), body_com=f32(m.body_com),
        body_inertia=f32(m.body_inertia), dof_body=i32(m.dof_body), dof_type=i32(m.dof_type),
        dof_qadr=i32(m.dof_qadr), dof_parent=i32(m.dof_parent), dof_limited=i32(m.dof_limited),
        dof_act=i32(m.dof_act), dof_axis=f32(m.dof_axis), dof_anchor=f32(m.dof_anchor),
        dof_stiffness=f32(m.dof_stiffness), dof_damping=f32(m.dof_damping), dof_armature=f32(m.dof_armature),
        dof_lo=f32(m.dof_lo), dof_hi=f32(m.dof_hi), dof_klim=f32(m.dof_klim), dof_blim=f32(m.dof_blim),
        dof_gear=f32(m.dof_gear), con_body=i32(m.con_body), con_pos=f32(m.con_pos), con_radius=f32(m.con_radius))
    st = _lib.IcemArticulatedModel(
        nb=m.nb, nq=m.nq, nv=m.nv, nu=m.nu, nc=m.nc, nsub=m.nsub, obs_offset=int(obs_offset), dt=m.dt,
        gravity=m.gravity, ctrl_limit=m.ctrl_limit, contact_stiffness=m.contact_stiffness,
        contact_damping=m.contact_damping, contact_damping_max=m.contact_damping_max,
        friction_viscous=m.friction_viscous, friction=m.friction,
        integrator=_lib.INTEGRATOR[getattr(m, "integrator", "euler")],
        **{k: (iptr(v) if v.dtype == np.int32 else fptr(v)) for k, v in keep.items()})
    return st, keep


@dataclass
class PlannerSettings:
    """Flattened keyword surface of MpcICem (reference: controllers/icem.py:22,213-233;
    controllers/mpc.py:22; controllers/abstract_controller.py:64-65)."""
    horizon: int
    num_simulated_trajectories: int
    action_low: np.ndarray
    action_high: np.ndarray
    dynamics: str = "dense_tanh"
    cost: str = "halfcheetah"
    obs_dim: int = 0
    penalise_flipping: bool = True
    factor_decrease_num: float = 1.0
    cost_along_trajectory: str = "sum"
    alpha: float = 0.1
    elites_size: int = 10
    opt_iterations: int = 3
    init_std: float = 0.5
    use_mean_actions: bool = True
    keep_previous_elites: bool = True
    shift_elites_over_time: bool = True
    fraction_elites_reused: float = 0.3
    noise_beta: float = 1.0
    seed: int = 0
    device: int = 0
    world_size: int = 1
    rank: int = 0
    colorednoise_v2: bool = False
    integrator: Optional[str] = None     # articulated ground-truth models: "euler" | "rk4"; None = the model's own
    keep_iteration_actions: bool = False
    planner: str = "icem"                # "icem" (MpcICem) | "cem_std" (MpcCemStd) | "random" (MpcRandom)
    action_change_frequency: int = 0     # random only (controllers/mpc.py:91)
    num_problems: int = 1                # independent MPC problems batched in one handle (plan_batch)
    cost_params: Optional[dict] = None   # cost="locomotion": dt, ctrl_weight, unhealthy_weight, z_lo, z_hi, state_bound,
                                         # z_index, z_strict, velocity_index, forward_weight
                                         # (environments/mujoco.py:153-176, 196-231, 314-343)
    execute_best_elite: bool = True      # cem_std only (controllers/mpc.py:237-240)
    shift_means: bool = True             # cem_std only (controllers/mpc.py:243-248)
    bounds_like_levine: bool = False     # cem_std only (controllers/mpc.py:290-301)
    articulated_model: object = None     # robots.CompiledModel; None = the built-in tables for `dynamics`
    obs_offset: Optional[int] = None


def _cost_fields(cp):
    cp = cp or {}
    inf = 3.0e38          # float32-representable stand-in for an open range end
    return dict(cost_z_index=int(cp.get("z_index", 0)), cost_z_strict=int(bool(cp.get("z_strict", False))),
                cost_dt=float(cp.get("dt", 0.0)), cost_ctrl_weight=float(cp.get("ctrl_weight", 0.0)),
                cost_unhealthy_weight=float(cp.get("unhealthy_weight", 0.0)),
                cost_z_lo=float(max(cp.get("z_lo", -inf), -inf)), cost_z_hi=float(min(cp.get("z_hi", inf), inf)),
                cost_state_bound=float(cp.get("state_bound", 0.0)),
                cost_velocity_index1=int(cp.get("velocity_index", -1)) + 1, cost_reserved=0,
                cost_forward_weight=float(cp.get("forward_weight", 1.0)),
                cost_goal_index=int(cp.get("goal_index", 0)), cost_achieved_index=int(cp.get("achieved_index", 0)),
                cost_goal_sparse=int(bool(cp.get("sparse", False))), cost_goal_shaped=int(bool(cp.get("shaped", False))),
                cost_goal_threshold=float(cp.get("threshold", 0.0)),
                cost_reach=(C.c_double * 4)(*[float(v) for v in cp.get("reach", (0.1, 0.11, 0.0, 0.0))]))


class Planner:
    def __init__(self, s: PlannerSettings):
        lib = _lib.load()
        self._lib = lib
        self.settings = s
        self._low = f32(s.action_low)
        self._high = f32(s.action_high)
        if self._low.ndim != 1 or self._low.shape != self._high.shape:
            raise ValueError("action bounds must be 1-D arrays of equal length")
        if s.cost_along_trajectory not in _lib.REDUCE:   # abstract_controller.py:88-91
            raise NotImplementedError(
                "Implement method {} to compute cost along trajectory".format(s.cost_along_trajectory))
        self.h, self.d = int(s.horizon), int(self._low.shape[0])
        cfg = _lib.IcemConfig(
            abi_version=_lib.ICEM_ABI_VERSION, device=s.device, horizon=self.h, act_dim=self.d,
            num_simulated_trajectories=int(s.num_simulated_trajectories), opt_iterations=int(s.opt_iterations),
            elites_size=int(s.elites_size), use_mean_actions=int(bool(s.use_mean_actions)),
            keep_previous_elites=int(bool(s.keep_previous_elites)),
            shift_elites_over_time=int(bool(s.shift_elites_over_time)),
            cost_along_trajectory=_lib.REDUCE[s.cost_along_trajectory], dynamics=_lib.DYN[s.dynamics],
            cost=_lib.COST[s.cost], cost_penalise_flipping=int(bool(s.penalise_flipping)), obs_dim=int(s.obs_dim),
            colorednoise_v2=int(bool(s.colorednoise_v2)), keep_iteration_actions=int(bool(s.keep_iteration_actions)),
            world_size=int(s.world_size), rank=int(s.rank), planner=_lib.PLANNER[s.planner],
            execute_best_elite=int(bool(s.execute_best_elite)), shift_means=int(bool(s.shift_means)),
            bounds_like_levine=int(bool(s.bounds_like_levine)),
            action_change_frequency=int(s.action_change_frequency), num_problems=int(s.num_problems),
            **_cost_fields(s.cost_params),
            factor_decrease_num=float(s.factor_decrease_num), alpha=float(s.alpha), init_std=float(s.init_std),
            fraction_elites_reused=float(s.fraction_elites_reused), noise_beta=float(s.noise_beta),
            seed=int(s.seed) & (2 ** 64 - 1), action_low=fptr(self._low), action_high=fptr(self._high))
        self._h = C.c_void_p()
        check(lib.icem_create(C.byref(cfg), C.byref(self._h)))
        self.k = lib.icem_num_elites(self._h)
        self.iters = int(s.opt_iterations)
        self.K = self.h // 2 + 1
        if s.dynamics in ("halfcheetah", "humanoid_standup", "articulated") and s.articulated_model is not False:
            from . import robots
            if s.dynamics == "articulated" and s.articulated_model is None:
                raise ValueError('dynamics="articulated" needs settings.articulated_model (robots.get_model(...))')
            model = s.articulated_model if s.articulated_model is not None else robots.get_model(s.dynamics)
            if s.integrator is not None and s.integrator != model.integrator:
                import dataclasses
                if s.integrator not in _lib.INTEGRATOR:
                    raise ValueError(f"unknown integrator {s.integrator!r}; choose from {sorted(_lib.INTEGRATOR)}")
                model = dataclasses.replace(model, integrator=s.integrator)
            obs_offset = s.obs_offset
            if obs_offset is None:     # HalfCheetah's 17-wide observation drops qpos[0] (environments/mujoco.py:80-82)
                obs_offset = 1 if (s.dynamics == "halfcheetah" and s.obs_dim != 18) else 0
            self.set_articulated_model(model, obs_offset)

    # ---- lifetime -----------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.icem_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- model --------------------------------------------------------------------------------
    def set_dense_model(self, w_obs, w_act, bias=None):
        w_obs, w_act = f32(w_obs), f32(w_act)
        n = w_obs.shape[0]
        if w_obs.shape != (n, n) or w_act.shape != (n, self.d):
            raise ValueError("dense model shapes must be [n,n] and [n,d]")
        b = f32(bias) if bias is not None else None
        check(self._lib.icem_set_dense_model(self._h, n, fptr(w_obs), fptr(w_act), fptr(b) if b is not None else None))

    def set_mlp_model(self, weights, biases):
        """Dense MLP forward model obs' = obs + W3 tanh(W2 tanh(W1 [obs, act] + b1) + b2) + b3 (icem_set_mlp_model);
        `weights[l]` is [out_l, in_l] row-major."""
        ws = [f32(w) for w in weights]
        bs = [f32(b) for b in biases]
        dims = np.ascontiguousarray([ws[0].shape[1]] + [w.shape[0] for w in ws], dtype=np.int32)
        for l, (w, b) in enumerate(zip(ws, bs)):
            if w.ndim != 2 or b.shape != (w.shape[0],) or (l and w.shape[1] != ws[l - 1].shape[0]):
                raise ValueError("inconsistent MLP layer shapes")
        wp = (C.POINTER(C.c_float) * len(ws))(*[fptr(w) for w in ws])
        bp = (C.POINTER(C.c_float) * len(bs))(*[fptr(b) for b in bs])
        check(self._lib.icem_set_mlp_model(self._h, len(ws), iptr(dims), wp, bp))

    def set_articulated_model(self, m, obs_offset=0):
        """Upload robots.CompiledModel tables (icem_set_articulated_model)."""
        st, _keep = articulated_model_struct(m, obs_offset)
        check(self._lib.icem_set_articulated_model(self._h, C.byref(st)))
        self.articulated = m

    @property
    def state_dim(self):
        return self._lib.icem_state_dim(self._h)

    # ---- plan step ----------------------------------------------------------------------------
    def begin_rollout(self):
        check(self._lib.icem_begin_rollout(self._h))

    def plan(self, state) -> np.ndarray:
        st = f64(state).ravel()
        out = np.empty(self.d, dtype=np.float64)
        check(self._lib.icem_plan(self._h, dptr(st), st.shape[0], dptr(out)))
        return out

    def plan_batch(self, states) -> np.ndarray:
        """One plan step of every problem of a `num_problems = B` handle: states [B, state_dim] -> actions [B, d]."""
        st = f64(states)
        if st.ndim != 2:
            raise ValueError("states must be [num_problems, state_dim]")
        out = np.empty((st.shape[0], self.d), dtype=np.float64)
        check(self._lib.icem_plan_batch(self._h, dptr(st), st.shape[1], st.shape[0], dptr(out)))
        return out

    active_problem = 0

    def set_active_problem(self, i):
        """Which problem mean() / std() / elites() / iteration_record() / costs() / actions() read."""
        check(self._lib.icem_set_active_problem(self._h, int(i)))
        self.active_problem = int(i)

    def plan_async(self, state):
        """Launch a plan step without waiting for it (icem_plan_async); pair with plan_finish()."""
        st = f64(state).ravel()
        check(self._lib.icem_plan_async(self._h, dptr(st), st.shape[0]))

    def plan_finish(self) -> np.ndarray:
        out = np.empty(self.d, dtype=np.float64)
        check(self._lib.icem_plan_finish(self._h, dptr(out)))
        return out

    def plan_device(self):
        check(self._lib.icem_plan_device(self._h))

    def advance_state_device(self):
        check(self._lib.icem_advance_state_device(self._h))

    def sync(self):
        check(self._lib.icem_sync(self._h))

    def last_plan_ms(self):
        a, b = C.c_float(), C.c_float()
        check(self._lib.icem_last_plan_ms(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def inject_noise(self, iteration, zr, zi=None):
        """Unit normal draws (iCEM) or, for planner="cem_std", the uniform draws u of truncnorm.rvs (float64);
        the latter are passed on as signed tail probabilities (include/icem_b200.h: icem_inject_noise)."""
        if self.settings.planner == "cem_std":
            u = np.asarray(zr, np.float64)
            zr = np.where(u < 0.5, u, -(1.0 - u))
        zr = f32(zr)
        zi_ = f32(zi) if zi is not None else None
        check(self._lib.icem_inject_noise(self._h, iteration, zr.shape[0], fptr(zr),
                                          fptr(zi_) if zi_ is not None else None))

    # ---- observable state ---------------------------------------------------------------------
    def mean(self):
        out = np.empty((self.h, self.d), dtype=np.float32)
        check(self._lib.icem_get_mean(self._h, fptr(out)))
        return out

    def std(self):
        out = np.empty((self.h, self.d), dtype=np.float32)
        check(self._lib.icem_get_std(self._h, fptr(out)))
        return out

    def elites(self):
        a = np.empty((self.k, self.h, self.d), dtype=np.float32)
        c = np.empty(self.k, dtype=np.float32)
        i = np.empty(self.k, dtype=np.int32)
        check(self._lib.icem_get_elites(self._h, fptr(a), fptr(c), iptr(i)))
        return a, c, i

    def population_size(self, iteration, first_step):
        g, l = C.c_int32(), C.c_int32()
        check(self._lib.icem_population_size(self._h, iteration, int(bool(first_step)), C.byref(g), C.byref(l)))
        return g.value, l.value

    def iteration_record(self, iteration):
        m = np.empty((self.h, self.d), dtype=np.float32)
        s = np.empty((self.h, self.d), dtype=np.float32)
        c = np.empty(self.k, dtype=np.float32)
        i = np.empty(self.k, dtype=np.int32)
        check(self._lib.icem_get_iteration(self._h, iteration, fptr(m), fptr(s), fptr(c), iptr(i)))
        return dict(mean=m, std=s, elite_costs=c, elite_idx=i)

    def costs(self, iteration, n):
        out = np.empty(n, dtype=np.float32)
        check(self._lib.icem_get_costs(self._h, iteration, fptr(out), n))
        return out

    def actions(self, iteration, n):
        out = np.empty((n, self.h, self.d), dtype=np.float32)
        check(self._lib.icem_get_actions(self._h, iteration, fptr(out), n))
        return out

    # ---- device model -------------------------------------------------------------------------
    def sim_step(self, state, action, obs_dim=0):
        st, ac = f64(state).ravel(), f64(action).ravel()
        nxt = np.empty(st.shape[0], dtype=np.float64)
        obs = np.empty(max(obs_dim, 1), dtype=np.float64)
        rew = C.c_double()
        check(self._lib.icem_sim_step(self._h, dptr(st), st.shape[0], dptr(ac), dptr(nxt),
                                      dptr(obs) if obs_dim else None, obs_dim, C.byref(rew)))
        return nxt, (obs[:obs_dim] if obs_dim else None), rew.value

    def rollout_observations(self, state, actions, obs_dim):
        """[n, h+1, obs_dim] observations along given action sequences [n, h, d] (icem_op_rollout_observations)."""
        st = f64(state).ravel()
        acts = f32(actions)
        if acts.ndim != 3 or acts.shape[1:] != (self.h, self.d):
            raise ValueError("actions must be [n, h, d]")
        out = np.empty((acts.shape[0], self.h + 1, obs_dim), dtype=np.float64)
        check(self._lib.icem_op_rollout_observations(self._h, acts.shape[0], dptr(st), st.shape[0], fptr(acts),
                                                     obs_dim, dptr(out)))
        return out

    def sim_step_batch(self, states, actions):
        """n independent transitions in one launch: states [n, state_dim], actions [n, d] -> next states."""
        st, ac = f64(states), f64(actions)
        if st.ndim != 2 or ac.ndim != 2 or st.shape[0] != ac.shape[0] or ac.shape[1] != self.d:
            raise ValueError("states must be [n, state_dim] and actions [n, d]")
        out = np.empty_like(st)
        check(self._lib.icem_sim_step_batch(self._h, st.shape[0], dptr(st), st.shape[1], dptr(ac), dptr(out)))
        return out

    def observe(self, state, obs_dim):
        st = f64(state).ravel()
        obs = np.empty(obs_dim, dtype=np.float64)
        check(self._lib.icem_observe(self._h, dptr(st), st.shape[0], dptr(obs), obs_dim))
        return obs

    # ---- single operators ---------------------------------------------------------------------
    def op_sample(self, zr, zi, mean, std):
        zr = f32(zr)
        zi_ = f32(zi) if zi is not None else None
        n = zr.shape[0]
        out = np.empty((n, self.h, self.d), dtype=np.float32)
        check(self._lib.icem_op_sample(self._h, n, fptr(zr), fptr(zi_) if zi_ is not None else None,
                                       fptr(f32(mean)), fptr(f32(std)), fptr(out)))
        return out

    def op_rollout_cost(self, state, actions):
        st, a = f64(state).ravel(), f32(actions)
        n = a.shape[0]
        out = np.empty(n, dtype=np.float32)
        check(self._lib.icem_op_rollout_cost(self._h, n, dptr(st), st.shape[0], fptr(a), fptr(out)))
        return out

    def op_topk(self, costs, k):
        c = f32(costs).ravel()
        idx = np.empty(k, dtype=np.int32)
        val = np.empty(k, dtype=np.float32)
        check(self._lib.icem_op_topk(self._h, c.shape[0], fptr(c), k, iptr(idx), fptr(val)))
        return idx, val

    # ---- multi-GPU ----------------------------------------------------------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(_lib.UNIQUE_ID_BYTES)
        check(_lib.load().icem_comm_get_unique_id(buf))
        return buf.raw

    def comm_init(self, unique_id: bytes):
        if len(unique_id) != _lib.UNIQUE_ID_BYTES:
            raise ValueError("unique id must be 128 bytes")
        check(self._lib.icem_comm_init(self._h, C.create_string_buffer(unique_id, _lib.UNIQUE_ID_BYTES)))

    # ---- bench --------------------------------------------------------------------------------
    def bench_op(self, op, n, reps=20, flush_l2=True):
        """Average launch duration (ms) of one kernel in isolation: op 0 sampler, 1 fused rollout, 2 select+refit."""
        ms = C.c_float()
        check(self._lib.icem_bench_op(self._h, {"sample": 0, "fused": 1, "select": 2, "rollout": 3, "exchange": 4}.get(op, op), int(n), int(reps),
                                      int(bool(flush_l2)), C.byref(ms)))
        return ms.value

    def bench_device(self, steps, warmup, flush_l2=True):
        tot, roll, n = C.c_float(), C.c_float(), C.c_int32()
        check(self._lib.icem_bench_device(self._h, steps, warmup, int(bool(flush_l2)), C.byref(tot), C.byref(roll),
                                          C.byref(n)))
        return tot.value, roll.value, n.value


def kernel_launch_count() -> int:
    return int(_lib.load().icem_kernel_launch_count())
