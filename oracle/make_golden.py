"""Generate tests/golden/*.npz by running the UNMODIFIED reference controller (imported from
/root/reference/icem) on the seeded cases of oracle/cases.py.  TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_golden          # this container only (needs /root/reference)
"""
import json
import os
import sys

import numpy as np

from oracle import cases, ref_harness

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    os.makedirs(OUT, exist_ok=True)
    todo = [(n, c, "MpcICem") for n, c in cases.CASES.items()] + \
           [(n, c, "MpcCemStd") for n, c in cases.CEM_STD_CASES.items()] + \
           [(n, c, "MpcRandom") for n, c in cases.RANDOM_CASES.items()]
    for name, case, controller in todo:
        model = case["model"]()
        steps, next_randn = ref_harness.run_reference_episode(
            model, case["cost"], case["ctrl"], case["low"], case["high"], case["start_obs"],
            case["seed"], case["steps"], case["penalise_flipping"], controller=controller)
        blob = {"next_randn": np.float64(next_randn), "num_steps": np.int64(len(steps))}
        for s, st in enumerate(steps):
            blob[f"s{s}_action"] = st["action"]
            if st["mean_after_shift"] is not None:
                blob[f"s{s}_mean_after_shift"] = st["mean_after_shift"]
                blob[f"s{s}_std_after_reset"] = st["std_after_reset"]
            blob[f"s{s}_num_iters"] = np.int64(len(st["iterations"]))
            for i, it in enumerate(st["iterations"]):
                blob[f"s{s}_i{i}_costs"] = it["costs"]
                blob[f"s{s}_i{i}_elite_idx"] = it["elite_idx"].astype(np.int64)
                if "mean" in it:
                    blob[f"s{s}_i{i}_mean"] = it["mean"]
                    blob[f"s{s}_i{i}_std"] = it["std"]
                else:       # MpcRandom: the sampled population itself is the thing to pin
                    blob[f"s{s}_i{i}_actions"] = it["actions"].astype(np.float32)
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **blob)
        print(name, "steps", len(steps), "pops", [len(it["costs"]) for it in steps[-1]["iterations"]],
              "action0", np.round(steps[0]["action"][:3], 6))
    with open(os.path.join(OUT, "README.json"), "w") as f:
        json.dump({"generator": "python -m oracle.make_golden",
                   "reference": "/root/reference/icem (martius-lab/iCEM, unmodified, imported)",
                   "colorednoise": "oracle/shims/colorednoise.py restatement, v1.1.1 semantics (parity unpinned)",
                   "numpy": np.__version__, "python": sys.version.split()[0]}, f, indent=1)


if __name__ == "__main__":
    main()
