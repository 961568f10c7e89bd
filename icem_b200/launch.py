"""Run the reference's UNCHANGED `icem/main.py` with the B200 plugin classes registered.

    python -m icem_b200.launch <settings.json> [--reference /path/to/iCEM/icem] [--shims DIR] [--override-mpc-icem]

The reference selects controller / forward model / environment by STRING through three registries
(icem/controllers/__init__.py:12-17, icem/models/__init__.py:5-8, icem/environments/__init__.py:26-55); this
launcher adds entries to them before calling `main.main()` and edits nothing else:

  controller     "mpc-icem-b200"            -> icem_b200.controller.MpcICemB200
                 "mpc-cem-std-b200"         -> icem_b200.controller.MpcCemStdB200 (vanilla CEM baseline)
                 "mpc-random-b200"          -> icem_b200.controller.MpcRandomB200 (random shooting baseline)
                 ("mpc-icem" / "mpc-cem-std" too with --override-mpc-icem, so existing settings files run unchanged)
  forward_model  "CudaGroundTruthModel"     -> icem_b200.models.CudaGroundTruthModel
                 "CudaDenseTanhModel"       -> icem_b200.models.CudaDenseTanhModel
  env            "HalfCheetah" / "HumanoidStandup" resolve to the device-simulated stand-ins of icem_b200.envs
                 (gym / mujoco-py are not installable here; on a machine with MuJoCo pass --keep-mujoco-envs to
                 leave `environments.mujoco` alone -- the CUDA controller then still needs a CUDA forward model)

`--shims DIR` prepends a directory of stand-in third-party packages (allogger, smart_settings, ...) to sys.path; the
test-suite passes oracle/shims, a deployment has the real packages installed.
"""
import argparse
import collections
import collections.abc
import os
import sys
import types

def _default_reference():
    env = os.environ.get("ICEM_REFERENCE_ROOT")
    if env:
        return env
    staged = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref", "icem")
    for cand in ("/root/reference/icem", staged):
        if os.path.isfile(os.path.join(cand, "main.py")):
            return cand
    return "/root/reference/icem"


DEFAULT_REFERENCE = _default_reference()


def prepare_paths(reference_root=DEFAULT_REFERENCE, shims=None):
    if not os.path.isfile(os.path.join(reference_root, "main.py")):
        raise FileNotFoundError(f"reference iCEM sources not found under {reference_root!r}")
    if not hasattr(collections, "Mapping"):          # icem/misc/helpers.py:5 (`from collections import Mapping`)
        collections.Mapping = collections.abc.Mapping
    for p in ([shims] if shims else []) + [reference_root]:
        if p not in sys.path:
            sys.path.insert(0 if p == shims else 1, p)
    repo_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if repo_root not in sys.path:
        sys.path.append(repo_root)


def register(override_mpc_icem=False, standin_envs=True):
    """Add the B200 entries to the reference's registries (the reference must already be importable)."""
    import controllers
    import models
    table = controllers.ControllerFactory.valid_base_controllers
    table["mpc-icem-b200"] = ("icem_b200.controller", "MpcICemB200")
    table["mpc-cem-std-b200"] = ("icem_b200.controller", "MpcCemStdB200")
    table["mpc-random-b200"] = ("icem_b200.controller", "MpcRandomB200")
    if override_mpc_icem:
        table["mpc-random"] = ("icem_b200.controller", "MpcRandomB200")
        table["mpc-icem"] = ("icem_b200.controller", "MpcICemB200")
        table["mpc-cem-std"] = ("icem_b200.controller", "MpcCemStdB200")
    models.models_dict["CudaGroundTruthModel"] = ("icem_b200.models", "CudaGroundTruthModel")
    models.models_dict["CudaDenseTanhModel"] = ("icem_b200.models", "CudaDenseTanhModel")
    models.models_dict["CudaMlpModel"] = ("icem_b200.models", "CudaMlpModel")
    if standin_envs:
        import environments  # noqa: F401  (package import must precede the submodule override)
        from . import envs
        mod = types.ModuleType("environments.mujoco")
        mod.HalfCheetahMaybeWithPosition = envs.HalfCheetahMaybeWithPosition
        mod.HumanoidStandup = envs.HumanoidStandup
        mod.Hopper = envs.Hopper
        mod.Ant = envs.Ant
        mod.Humanoid = envs.Humanoid
        mod.Reacher = envs.Reacher
        mod.__doc__ = "device-simulated stand-ins registered by icem_b200.launch"
        sys.modules["environments.mujoco"] = mod


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("settings")
    ap.add_argument("--reference", default=DEFAULT_REFERENCE)
    ap.add_argument("--shims", default=None)
    ap.add_argument("--override-mpc-icem", action="store_true")
    ap.add_argument("--keep-mujoco-envs", action="store_true")
    args = ap.parse_args(argv)
    prepare_paths(args.reference, args.shims)
    register(override_mpc_icem=args.override_mpc_icem, standin_envs=not args.keep_mujoco_envs)
    import main as ref_main
    sys.argv = [os.path.join(args.reference, "main.py"), os.path.abspath(args.settings)]
    ref_main.main()


if __name__ == "__main__":
    main()
