#!/usr/bin/env python
"""bench.py -- iCEM plan-step throughput on B200 (contract: see the task statement / DESIGN.md section "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

A "step" is one plan step (`get_action`: all CEM iterations) on one synthetic start state.
  value   candidate trajectories sampled + rolled out + scored per second, whole job, start state resident in HBM,
          closed loop on the device (plan -> advance state), CUDA events on the planner's stream, max over ranks;
  e2e     the same metric through the reference-facing plugin call `MpcICemB200.get_action(obs, state)` with HOST
          NumPy buffers: host->device copy of the state and device->host copy of the action inside the timed region;
  roofline   HBM roofline of the dominant kernel (fused sample->rollout->cost): algorithmic bytes per launch
          (4*h*d + 4 per trajectory) / its CUDA-event duration, against MEASURED_PEAKS.json;
  cpu_baseline  the NumPy oracle port of the reference path timed on this box's host cores (rank 0, N=1 only).
`--impl reference` times that CPU path alone with all host cores (same metric / config).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "candidate trajs rolled-out/sec at h=30"
UNIT = "trajectories/s"
DEFAULT_WORKLOAD = "humanoid_standup_gt_n16384"
DEFAULT_SHARD_WORKLOAD = DEFAULT_WORKLOAD     # weak scaling: the same 16384 trajectories per GPU at every N


# ------------------------------------------------------------------------------------------------------------
def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                p = json.load(f)
            if "hbm_gbs" in p:
                return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
            for k, v in p.items():          # tolerate another spelling of the key
                if "hbm" in str(k).lower() and isinstance(v, (int, float)) and v > 1000:
                    return float(v), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def read_bf16_peak():
    """Dense bf16 TFLOP/s of this pool's B200s from MEASURED_PEAKS.json (the burst figure: the MLP kernel is timed
    alone), whatever the exact key spelling; else the fallback of B200_PROFILING.md."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            p = json.load(open(path))
            flat = {}

            def walk(prefix, v):
                if isinstance(v, dict):
                    for k, x in v.items():
                        walk(prefix + "." + str(k).lower(), x)
                elif isinstance(v, (int, float)):
                    flat[prefix] = float(v)
            walk("", p)
            cands = [(k, v) for k, v in flat.items() if "bf16" in k and v > 100]
            for pref in ("burst", "tflops", ""):
                for k, v in cands:
                    if pref in k and "sustain" not in k:
                        return v
            if cands:
                return cands[0][1]
        except Exception:
            pass
    return 1590.0


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.device)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


def dist_setup(n_gpus):
    """torchrun env -> (rank, world, local_rank); initialises torch.distributed (NCCL) when world > 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local


def barrier(world):
    if world > 1:
        import torch
        import torch.distributed as dist
        dist.barrier()
        torch.cuda.synchronize()


def max_over_ranks(x, world):
    if world == 1:
        return x
    import torch
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ------------------------------------------------------------------------------------------------------------
def build_planner(name, world, rank, device):
    from icem_b200 import workloads
    from icem_b200.planner import Planner
    s = workloads.planner_settings(name, world_size=world, rank=rank, device=device, seed=0)
    p = Planner(s)
    w = workloads.get_workload(name)
    if w.get("dense"):
        p.set_dense_model(*workloads.dense_model_weights(*w["dense"]))
    return p, s


def build_controller(name, world, rank, device):
    """The reference-facing plugin object (MpcICemB200) on a stand-in env + CUDA model, like main.py builds them."""
    from icem_b200 import envs, workloads
    from icem_b200.controller import MpcICemB200
    from icem_b200.models import CudaDenseTanhModel, CudaGroundTruthModel, CudaMlpModel
    w = workloads.get_workload(name)
    st = dict(w["settings"])
    sampler = {k: st[k] for k in ("alpha", "elites_size", "opt_iterations", "init_std", "use_mean_actions",
                                  "keep_previous_elites", "shift_elites_over_time", "fraction_elites_reused",
                                  "noise_beta")}
    if w.get("mlp"):
        ws, bs = workloads.mlp_model_weights(*w["mlp"])
        env = envs.MlpStandInEnv(name="mlp", act_dim=w["act_dim"], bound=w["bound"], cost=st["cost"],
                                 obs_dim=st["obs_dim"], penalise_flipping=st.get("penalise_flipping", False),
                                 mlp=(ws, bs))
        model = CudaMlpModel(env=env, weights=ws, biases=bs)
    elif w["env"] is None:
        wts = workloads.dense_model_weights(*w["dense"])
        env = envs.DenseStandInEnv(name="dense", act_dim=w["act_dim"], bound=w["bound"], cost=st["cost"],
                                   obs_dim=st["obs_dim"], penalise_flipping=st.get("penalise_flipping", False),
                                   weights=wts)
        model = CudaDenseTanhModel(env=env, **dict(zip(("w_obs", "w_act", "bias"), wts)))
    else:
        env = envs.make_env(w["env"], device=device, **w.get("env_kwargs", {}))
        model = CudaGroundTruthModel(env=env)
    ctrl = MpcICemB200(env=env, forward_model=model, horizon=st["horizon"],
                       num_simulated_trajectories=st["num_simulated_trajectories"],
                       factor_decrease_num=st["factor_decrease_num"], cost_along_trajectory=st["cost_along_trajectory"],
                       action_sampler_params=sampler, seed=0, device=device, world_size=world, rank=rank)
    return ctrl, env


def strong_and_digest(world, rank, local, steps):
    """What the weak-scaling line cannot show (BASELINE configs[4]):

    strong       plan-step time at a FIXED global population of 262144 HumanoidStandup trajectories sharded over the
                 ranks (131072 / 65536 / 32768 per GPU at 2 / 4 / 8), device-timed, max over ranks;
    plan_digest  sha256 over the executed actions and the last iteration's elite index lists of 3 closed-loop plan
                 steps at a fixed global N = 16384: identical digests at 1, 2, 4, 8 GPUs = the sharded plan is
                 bit-identical to the single-GPU plan (global trajectory indices key the Philox draws, the merge is
                 a total order on (cost, global index));
    exchange_us  one elite all-gather + merge/refit kernel (the only data-path collective), CUDA events."""
    import hashlib
    from icem_b200 import workloads
    from icem_b200.planner import Planner
    name = DEFAULT_WORKLOAD
    out = {}
    # ---- digest ----
    s = workloads.planner_settings(name, world_size=world, rank=rank, device=local, seed=0)
    p = Planner(s)
    if world > 1:
        from icem_b200.distributed import init_planner_comm
        init_planner_comm(p)
    p.begin_rollout()
    state = workloads.start_state(name, seed=0)
    h = hashlib.sha256()
    for _ in range(3):
        act = p.plan(state)
        rec = p.iteration_record(s.opt_iterations - 1)
        h.update(np.asarray(act, np.float32).tobytes())
        h.update(np.asarray(rec["elite_idx"], np.int32).tobytes())
        state, _, _ = p.sim_step(state, act)
    out["plan_digest"] = {"sha256": h.hexdigest(), "global_population": s.num_simulated_trajectories,
                          "closed_loop_steps": 3, "what": "executed actions (f32) + last-iteration elite indices (i32)"}
    if world > 1:
        ms = p.bench_op("exchange", 0, reps=50, flush_l2=False)
        out["exchange_us"] = {"value": 1e3 * max_over_ranks(ms, world), "what": "ncclAllGather of k elite records per "
                              "rank + merge_refit_kernel, per CEM iteration, CUDA events, max over ranks",
                              "bytes_per_rank": 8 * p.k + 4 * p.k * ((s.horizon * p.d + 3) // 4 * 4)}
    p.close()
    # ---- strong scaling at global N = 262144 ----
    n_global = 262144
    s2 = workloads.planner_settings(name, world_size=world, rank=rank, device=local, seed=0,
                                    scale_population=n_global // s.num_simulated_trajectories)
    p2 = Planner(s2)
    if world > 1:
        from icem_b200.distributed import init_planner_comm
        init_planner_comm(p2)
    p2.begin_rollout()
    p2.plan(workloads.start_state(name, seed=0))
    barrier(world)
    k_steps = max(2, min(int(steps), 5))
    total_ms, _, _ = p2.bench_device(k_steps, 3, flush_l2=True)
    barrier(world)
    total_ms = max_over_ranks(total_ms, world)
    traj = workloads.trajectories_per_step(s2, first_step=False)
    out["strong"] = {"workload": "humanoid_standup_gt global N=262144 (BASELINE configs[4])", "scaling": "strong",
                     "global_population": n_global, "per_gpu_population": -(-n_global // world), "n_gpus": world,
                     "steps": k_steps, "warmup": 3, "ms_per_step": total_ms / k_steps,
                     "value": k_steps * traj / (total_ms * 1e-3), "unit": UNIT,
                     "populations_global": workloads.populations(s2)}
    p2.close()
    return out


def run_ours(args):
    rank, world, local = dist_setup(args.gpus)
    import torch  # noqa: F401  (device memory / streams / torch.distributed are plumbing here)
    from icem_b200 import workloads
    from icem_b200.planner import kernel_launch_count
    name = args.workload or (DEFAULT_WORKLOAD if world == 1 else DEFAULT_SHARD_WORKLOAD)
    scale = world if world > 1 else 1          # weak scaling: N(global) = per-GPU shard * ranks
    w = workloads.get_workload(name)
    peak, peak_src = read_peaks()

    # ---- device-resident closed loop: `value` ----------------------------------------------------------
    from icem_b200.planner import Planner
    s = workloads.planner_settings(name, world_size=world, rank=rank, device=local, seed=0, scale_population=scale)
    planner = Planner(s)
    if w.get("dense"):
        planner.set_dense_model(*workloads.dense_model_weights(*w["dense"]))
    if w.get("mlp"):
        planner.set_mlp_model(*workloads.mlp_model_weights(*w["mlp"]))
    if world > 1:
        from icem_b200.distributed import init_planner_comm
        init_planner_comm(planner)
    start = workloads.start_state(name, seed=0)
    planner.begin_rollout()
    planner.plan(start)                          # uploads the start state; first (cold) step, untimed
    clocks = ClockSampler(local)
    barrier(world)
    launches0 = kernel_launch_count()
    clocks.start()
    total_ms, rollout_ms, n_roll = planner.bench_device(args.steps, args.warmup, flush_l2=True)
    clk = clocks.stop()
    barrier(world)
    launches = kernel_launch_count() - launches0
    total_ms = max_over_ranks(total_ms, world)
    traj_step = workloads.trajectories_per_step(s, first_step=False)       # global, steady state
    value = args.steps * traj_step / (total_ms * 1e-3)
    launches_timed = int(round(launches * args.steps / float(args.steps + args.warmup)))

    # ---- roofline of the dominant kernel (fused sample->rollout->cost) ---------------------------------
    h, d = s.horizon, w["act_dim"]
    bytes_per_traj = 4 * h * d + 4
    local_rows = []
    for i in range(s.opt_iterations):
        _, nl = planner.population_size(i, first_step=False)
        local_rows.append(nl)
    avg_rows = float(np.mean(local_rows))
    avg_kernel_ms = rollout_ms / max(n_roll, 1)
    achieved = bytes_per_traj * avg_rows / (avg_kernel_ms * 1e-3) / 1e9
    traffic = None
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        try:
            traffic = json.load(open(prof)).get(name)
        except Exception:
            traffic = None
    compute = None
    prof_c = os.path.join(ROOT, "profiles", "compute.json")
    if os.path.exists(prof_c):
        try:
            compute = json.load(open(prof_c)).get(name)
        except Exception:
            compute = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src,
                "kernel": ("chain_rollout_kernel<sample,rollout> (fused sample->rollout->cost, branch-parallel "
                           "articulated engine)") if w.get("env") else "rollout_kernel<sample,rollout> (fused)",
                "kernel_ms_avg": avg_kernel_ms, "kernel_share_of_step": rollout_ms / max(total_ms, 1e-9)
                if world == 1 else None,
                "algorithmic_bytes_per_trajectory": bytes_per_traj,
                "note": "compute/latency-bound kernel (fp32 dynamics); HBM fraction reported as the contract asks",
                "compute": compute}
    if w.get("mlp"):
        # tensor-core rollout: FLOP roofline of mlp_rollout_kernel timed alone (sampler excluded)
        od, ad, hid, _ = w["mlp"]
        n0 = local_rows[0]
        ms = planner.bench_op("rollout", n0, reps=10)
        flops = 2.0 * (h - 1) * ((od + ad) * hid + hid * hid + hid * od) * n0
        pk = read_bf16_peak()
        ach = flops / (ms * 1e-3) / 1e12
        roofline = {"bound": "tensor", "achieved": ach, "peak": pk, "unit": "TFLOP/s", "frac": ach / pk,
                    "traffic": None, "peak_source": peak_src + " (bf16 burst: kernel timed alone)",
                    "kernel": "mlp_rollout_kernel (tcgen05.mma, bf16 operands, fp32 accumulate in TMEM)",
                    "kernel_ms_avg": ms, "kernel_rows": n0,
                    "algorithmic_flops_per_trajectory": flops / n0,
                    "note": "latency-chain + MUFU bound: one 128-row tile per SM (resident weights fill shared memory), per step L1 MMA -> tanh epilogue (MUFU 16/clk/SM) -> L2 MMAs pipelined under it -> tanh epilogue -> L3; see DESIGN.md section 9 for the measured per-step timeline"}
    planner.close()

    # ---- end to end through the plugin API with host buffers: `e2e` ------------------------------------
    ctrl, env = build_controller(name, world, rank, local) if world == 1 else (None, None)
    e2e = None
    if ctrl is not None:
        obs = env.reset()
        state = env.get_GT_state()
        import contextlib
        with contextlib.redirect_stdout(sys.stderr):      # keep stdout to the ONE JSON line (the reference prints here)
            ctrl.beginning_of_rollout(observation=obs, state=state, mode="train")
        t_sum = 0.0
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            act = ctrl.get_action(obs, state=state, mode="train")
            t1 = time.perf_counter()
            if i >= args.warmup:
                t_sum += t1 - t0
            obs, _, _, _ = env.step(act)
            state = env.get_GT_state()
        in_bytes = 16 + 256 * 4                      # StepState + start-state staging block (csrc/planner.cu)
        e2e = {"value": args.steps * traj_step / t_sum, "unit": UNIT, "h2d_bytes_per_step": in_bytes,
               "d2h_bytes_per_step": 4 * (d + 1), "ms_per_step": 1e3 * t_sum / args.steps,
               "call": "MpcICemB200.get_action(obs, state) -> ctypes -> icem_plan (CUDA graph)"}
        ctrl.close()
    else:
        # multi-rank e2e: icem_plan with a host state on every rank (same call the controller makes)
        planner2 = Planner(s)
        if w.get("dense"):
            planner2.set_dense_model(*workloads.dense_model_weights(*w["dense"]))
        if w.get("mlp"):
            planner2.set_mlp_model(*workloads.mlp_model_weights(*w["mlp"]))
        from icem_b200.distributed import init_planner_comm
        init_planner_comm(planner2)
        planner2.begin_rollout()
        t_sum = 0.0
        for i in range(args.warmup + args.steps):
            barrier(world)
            t0 = time.perf_counter()
            planner2.plan(start)
            t1 = time.perf_counter()
            if i >= args.warmup:
                t_sum += t1 - t0
        t_sum = max_over_ranks(t_sum, world)
        e2e = {"value": args.steps * traj_step / t_sum, "unit": UNIT, "h2d_bytes_per_step": (16 + 256 * 4) * world,
               "d2h_bytes_per_step": 4 * (d + 1) * world, "ms_per_step": 1e3 * t_sum / args.steps,
               "call": "icem_plan on every rank (host state in, action out)"}
        planner2.close()

    # ---- configs[4]: strong scaling at global N = 262144 + bit-identity digest (default workload only) ----
    extra = {}
    if name == DEFAULT_WORKLOAD and not args.no_strong:
        extra = strong_and_digest(world, rank, local, args.steps)

    # ---- CPU baseline beside it (rank 0, single GPU run only) -------------------------------------------
    cpu = None
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        cpu = cpu_baseline(name, budget_s=15.0, steps=2)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": name, "horizon": h, "act_dim": d,
                       "num_simulated_trajectories_global": s.num_simulated_trajectories,
                       "populations_global": workloads.populations(s), "opt_iterations": s.opt_iterations,
                       "noise_beta": s.noise_beta, "trajectories_per_step": traj_step,
                       "integrator": (s.integrator or "euler") if w.get("env") else None,
                       "sharding": f"num_sim_traj over {world} rank(s), one NCCL all-gather of elites per CEM iteration"
                       if world > 1 else "single GPU",
                       "l2": "256 MiB memset between timed steps (L2 flush)"},
            "clocks": clk, "e2e": e2e, "gpu_launches": launches_timed, "roofline": roofline,
            "cpu_baseline": cpu,
        }
        line.update(extra)
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def cpu_baseline(name, budget_s, steps):
    """The reference's CPU path on this box's host cores: the UNMODIFIED reference `MpcICem` when its sources are
    present (kind "reference": /root/reference in the build container, baseline/_ref/icem staged on a GPU box by
    scripts/stage_reference.sh), else the NumPy oracle port of the same plan step (kind "port")."""
    import contextlib
    from oracle import cpu_bench, ref_loader
    from icem_b200 import workloads
    cores = os.cpu_count() or 1
    with contextlib.redirect_stdout(sys.stderr):          # the reference prints; stdout carries the ONE JSON line
        if ref_loader.reference_available() and workloads.get_workload(name).get("env"):
            try:
                return cpu_bench.run_reference(name, cores=cores, budget_s=budget_s, steps=steps)
            except Exception as e:                            # never lose the bench line over the baseline leg
                sys.stderr.write(f"reference arm failed ({e!r}); falling back to the oracle port\n")
        return cpu_bench.run(name, cores=cores, budget_s=budget_s, steps=steps, warmup=0)


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the path with all host cores (see cpu_baseline).
    Rank 0 only.  The reference is pure Python, so there is no oracle/_ref to compile; it is imported from its sources.
    `steps` / `warmup` in the line are what was actually run (a CPU plan step at the full population takes minutes:
    each step is a bounded sample, population stated in `config`)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    name = args.workload or (DEFAULT_WORKLOAD if world == 1 else DEFAULT_SHARD_WORKLOAD)
    steps = max(1, min(args.steps, 2))
    res = cpu_baseline(name, budget_s=25.0, steps=steps)
    from icem_b200 import workloads
    s = workloads.planner_settings(name)
    line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": world,
            "steps": steps, "warmup": 0, "steps_requested": args.steps, "ms_per_step": res["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": name, "horizon": s.horizon, "act_dim": len(s.action_low),
                       "opt_iterations": s.opt_iterations, "noise_beta": s.noise_beta,
                       "population": res.get("population"), "population_scaled_from": s.num_simulated_trajectories,
                       "same_population_as_gpu_arm": res.get("population") == s.num_simulated_trajectories,
                       "sample": res["sample"]},
            "cpu_baseline": res,
            "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling / digest legs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
