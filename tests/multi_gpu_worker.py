"""Worker for tests/test_gpu_multi.py (launched with torch.distributed.run, one process per GPU).

Every rank plans the SAME problem sharded over the ranks (production Philox noise, keyed by the global trajectory
index) and rank 0 compares against a single-GPU planner of the same global population: the per-iteration elite
index lists, elite costs, refit mean/std and the executed action must be IDENTICAL (T6)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from icem_b200 import workloads
    from icem_b200.distributed import default_placement, init_planner_comm
    from icem_b200.planner import Planner
    dev, world, rank = default_placement()
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
    for name, scale in (("halfcheetah_gt_n4096", 1 / 8), ("humanoid_standup_gt_n16384", 1 / 32),
                        ("dense_tanh_cheetah_n4096", 1 / 4)):
        w = workloads.get_workload(name)
        s = workloads.planner_settings(name, world_size=world, rank=rank, device=dev, seed=3, scale_population=scale)
        p = Planner(s)
        if w.get("dense"):
            p.set_dense_model(*workloads.dense_model_weights(*w["dense"]))
        init_planner_comm(p)
        state = workloads.start_state(name)
        p.begin_rollout()
        recs, acts = [], []
        for step in range(3):
            acts.append(p.plan(state))
            recs.append([p.iteration_record(i) for i in range(s.opt_iterations)])
        mean = p.mean()
        if rank == 0:
            s1 = workloads.planner_settings(name, world_size=1, rank=0, device=dev, seed=3, scale_population=scale)
            q = Planner(s1)
            if w.get("dense"):
                q.set_dense_model(*workloads.dense_model_weights(*w["dense"]))
            q.begin_rollout()
            for step in range(3):
                a1 = q.plan(state)
                np.testing.assert_array_equal(a1, acts[step])
                for i in range(s.opt_iterations):
                    r1 = q.iteration_record(i)
                    for key in ("elite_idx", "elite_costs", "mean", "std"):
                        np.testing.assert_array_equal(r1[key], recs[step][i][key], err_msg=f"{name} {step} {i} {key}")
            np.testing.assert_array_equal(q.mean(), mean)
            q.close()
            print(f"{name}: {world}-rank sharded plan == single-GPU plan (bit-identical)")
        dist.barrier()
        p.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
