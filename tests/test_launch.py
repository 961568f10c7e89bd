"""CPU plumbing tests (T7 / BASELINE configs[0]): the reference's UNCHANGED icem/main.py runs in this container

  (a) with its own MpcICem + GroundTruthModel over the oracle-backed stand-in HalfCheetah (proves shims, settings
      hierarchy, registries, checkpoints), and
  (b) through icem_b200.launch with controller "mpc-icem-b200" + forward model "CudaGroundTruthModel": every layer up
      to the C ABI is exercised and, because this container has no GPU, icem_create must fail LOUDLY (no CPU fallback).

  (c) on a GPU (marker `gpu`): the same launcher run finishes -- the unchanged main.py drives MpcICemB200 for a few env
      steps, checkpoints are written -- and `elite_samples` is the reference's own RolloutBuffer.

All need the reference sources: /root/reference in the build container, or the copy staged for GPU boxes by
scripts/stage_reference.sh (baseline/_ref/icem, found by oracle/ref_loader.py); skipped when neither exists."""
import json
import os
import subprocess
import sys

import pytest

from oracle import ref_loader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not ref_loader.reference_available(), reason="reference sources not present")


def _settings(tmp_path, **over):
    s = {
        "inherits_from": [],
        "env": "HalfCheetah",
        "env_params": {"exclude_current_positions_from_observation": True, "penalise_flipping": True},
        "seed": 3,
        "controller": "mpc-icem",
        "controller_params": {
            "horizon": 30, "num_simulated_trajectories": 16, "factor_decrease_num": 1.25,
            "cost_along_trajectory": "sum", "do_visualize_plan": False, "verbose": False,
            "action_sampler_params": {"alpha": 0.1, "elites_size": 4, "fraction_elites_reused": 0.3, "init_std": 0.5,
                                      "keep_previous_elites": True, "shift_elites_over_time": True,
                                      "use_mean_actions": True, "opt_iterations": 3, "noise_beta": 0.25}},
        "forward_model": "GroundTruthModel", "forward_model_params": {},
        "initial_controller": "none", "initial_controller_params": {}, "initial_number_of_rollouts": 0,
        "number_of_rollouts": 1, "append_data": False, "append_data_eval": False, "training_iterations": 1,
        "evaluation_rollouts": 0,
        "rollout_params": {"render": False, "render_initial": False, "render_eval": False, "record": False,
                           "only_final_reward": False, "use_env_states": True, "task_horizon": 2},
        "checkpoints": {"load": False, "save": True, "save_every_n_iter": 1, "restart_every_n_iter": None},
        "model_dir": str(tmp_path / "results"),
    }
    s.update(over)
    path = tmp_path / "settings.json"
    path.write_text(json.dumps(s))
    return str(path)


def test_reference_main_runs_unchanged_on_cpu(tmp_path):
    """(a) reference controller + reference GroundTruthModel + oracle-backed env, 2 env steps."""
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from oracle import ref_loader, envs_np\n"
        "ref_loader.install_reference(); envs_np.install()\n"
        "import main; sys.argv=['main.py', %r]; main.main()\n" % (ROOT, _settings(tmp_path)))
    res = subprocess.run([sys.executable, "-c", code], cwd=str(tmp_path), capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "iCEM using" in res.stdout                                  # controllers/icem.py:42-43
    out = tmp_path / "results"
    assert (out / "settings.json").exists()
    assert (out / "checkpoints_latest").exists()


def test_launcher_reaches_c_abi_and_fails_loudly_without_gpu(tmp_path):
    """(b) the B200 plugin classes are resolved by the reference's registries and constructed by main.get_controllers;
    without a CUDA device the C ABI refuses (never a silent CPU path)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("this leg asserts the no-GPU failure mode")
    path = _settings(tmp_path, controller="mpc-icem-b200", forward_model="CudaGroundTruthModel",
                     forward_model_params={"num_parallel": 8})
    res = subprocess.run([sys.executable, "-m", "icem_b200.launch", path, "--shims",
                          os.path.join(ROOT, "oracle", "shims")], cwd=str(tmp_path), capture_output=True, text=True,
                         timeout=600, env=dict(os.environ, PYTHONPATH=ROOT))
    assert res.returncode != 0
    assert "IcemError" in res.stderr and "cuda" in res.stderr.lower(), res.stderr[-2000:]
    assert "MpcICemB200" in res.stderr or "controller.py" in res.stderr


def test_register_adds_every_plugin_entry():
    """icem_b200.launch.register: controller / model registry entries and the stand-in env module."""
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from icem_b200 import launch\n"
        "launch.prepare_paths(shims=%r); launch.register(override_mpc_icem=True)\n"
        "import controllers, models, environments.mujoco as m\n"
        "t = controllers.ControllerFactory.valid_base_controllers\n"
        "assert t['mpc-icem-b200'][1] == 'MpcICemB200' and t['mpc-cem-std-b200'][1] == 'MpcCemStdB200'\n"
        "assert t['mpc-random-b200'][1] == 'MpcRandomB200' and t['mpc-icem'][0] == 'icem_b200.controller'\n"
        "assert t['mpc-random'][1] == 'MpcRandomB200'\n"
        "assert {'CudaGroundTruthModel', 'CudaDenseTanhModel', 'CudaMlpModel'} <= set(models.models_dict)\n"
        "for n in ('HalfCheetahMaybeWithPosition', 'HumanoidStandup', 'Hopper', 'Ant', 'Humanoid', 'Reacher'):\n"
        "    assert hasattr(m, n), n\n"
        "cls = controllers.controller_from_string('mpc-random-b200')\n"
        "from controllers.abstract_controller import ModelBasedController\n"
        "assert issubclass(cls, ModelBasedController)\n" % (ROOT, os.path.join(ROOT, "oracle", "shims")))
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]


def test_example_settings_resolve_through_the_reference_loader(tmp_path):
    """examples/*.json go through the reference's own settings code path (smart_settings load + hierarchy merge) and
    name only registered plugins."""
    code = (
        "import sys, glob, os; sys.path.insert(0, %r)\n"
        "from icem_b200 import launch\n"
        "launch.prepare_paths(shims=%r); launch.register()\n"
        "import controllers, models\n"
        "from misc.helpers import resolve_params_hierarchy\n"
        "import smart_settings\n"
        "files = sorted(glob.glob(os.path.join(%r, 'examples', '*.json')))\n"
        "assert len(files) >= 2\n"
        "for f in files:\n"
        "    sys.argv = ['main.py', f]\n"          # helpers.py:160 reads the settings path from argv
        "    p = resolve_params_hierarchy(smart_settings.load(f))\n"
        "    assert p['controller'] in controllers.ControllerFactory.valid_base_controllers, f\n"
        "    assert p['forward_model'] in models.models_dict, f\n"
        "    cp = p['controller_params']\n"
        "    assert cp['horizon'] == 30 and cp['action_sampler_params']['elites_size'] == 10\n"
        "    assert p['rollout_params']['use_env_states'] is True\n" % (ROOT, os.path.join(ROOT, "oracle", "shims"), ROOT))
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=str(tmp_path))
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]


@pytest.mark.gpu
def test_unchanged_main_drives_the_b200_controller_on_the_gpu(tmp_path):
    """(c) T7 on the device: icem/main.py (unchanged) -> ControllerFactory -> MpcICemB200 -> C ABI -> CUDA kernels,
    3 env steps, checkpoint written (reference: icem/main.py:82-243, misc/rollout_utils.py:155-227)."""
    path = _settings(tmp_path, controller="mpc-icem-b200", forward_model="CudaGroundTruthModel",
                     forward_model_params={"num_parallel": 8})
    cfg = json.load(open(path))
    cfg["controller_params"]["num_simulated_trajectories"] = 64
    cfg["controller_params"]["action_sampler_params"]["elites_size"] = 10
    cfg["rollout_params"]["task_horizon"] = 3
    json.dump(cfg, open(path, "w"))
    res = subprocess.run([sys.executable, "-m", "icem_b200.launch", path, "--reference", ref_loader.REFERENCE_ROOT,
                          "--shims", os.path.join(ROOT, "oracle", "shims")], cwd=str(tmp_path), capture_output=True,
                         text=True, timeout=900, env=dict(os.environ, PYTHONPATH=ROOT))
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-3000:]
    assert "iCEM using" in res.stdout                                   # banner of MpcICemB200 (controllers/icem.py:42-43)
    out = tmp_path / "results"
    assert (out / "settings.json").exists() and (out / "checkpoints_latest").exists()


@pytest.mark.gpu
def test_elite_samples_are_the_reference_rolloutbuffer_on_the_gpu(tmp_path):
    """With the reference importable the controller's `elite_samples` must be the reference's own RolloutBuffer of
    Rollout objects (misc/rolloutbuffer.py:10-54,125-281), not the stand-in of icem_b200.api."""
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "import numpy as np\n"
        "from icem_b200 import launch\n"
        "launch.prepare_paths(%r, shims=%r); launch.register()\n"
        "from environments import env_from_string\n"
        "from controllers import controller_from_string\n"
        "from models import forward_model_from_string\n"
        "env = env_from_string('HalfCheetah', penalise_flipping=True, exclude_current_positions_from_observation=True)\n"
        "fm = forward_model_from_string('CudaGroundTruthModel')(env=env)\n"
        "cls = controller_from_string('mpc-icem-b200')\n"
        "ctrl = cls(env=env, forward_model=fm, horizon=30, num_simulated_trajectories=64, factor_decrease_num=1.25,\n"
        "           cost_along_trajectory='sum', action_sampler_params=dict(alpha=0.1, elites_size=10,\n"
        "           opt_iterations=3, init_std=0.5, use_mean_actions=True, keep_previous_elites=True,\n"
        "           shift_elites_over_time=True, fraction_elites_reused=0.3, noise_beta=0.25), seed=1)\n"
        "obs = env.reset(); st = env.get_GT_state()\n"
        "ctrl.beginning_of_rollout(observation=obs, state=st, mode='train')\n"
        "a = ctrl.get_action(obs, state=st, mode='train')\n"
        "es = ctrl.elite_samples\n"
        "from misc.rolloutbuffer import RolloutBuffer, Rollout\n"
        "assert type(es) is RolloutBuffer, type(es)\n"
        "assert len(es) == 10 and isinstance(es[0], Rollout)\n"
        "acts = es.as_array('actions'); obs_ = es.as_array('observations')\n"
        "assert acts.shape == (10, 30, 6) and obs_.shape == (10, 30, 17), (acts.shape, obs_.shape)\n"
        "np.testing.assert_allclose(acts[0, 0], a, atol=1e-6)\n"
        "print('reference RolloutBuffer OK')\n" % (ROOT, ref_loader.REFERENCE_ROOT, os.path.join(ROOT, "oracle", "shims")))
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=str(tmp_path))
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-3000:]
    assert "reference RolloutBuffer OK" in res.stdout


@pytest.mark.gpu
def test_unchanged_main_trains_the_mlp_model_on_the_gpu(tmp_path):
    """The reference's model-based loop (icem/main.py:193-210: collect rollouts with the controller, then
    `forward_model.train(rollout_buffer)`) with the learned model on the device: CudaMlpModel starts from a random
    initialisation, MpcICemB200 plans through it on the tensor cores, the REAL RolloutBuffer of the reference is handed
    to train(), the Adam steps run on the GPU, the next iteration plans with the trained weights, and the checkpoint
    holds the model."""
    path = _settings(tmp_path, controller="mpc-icem-b200", forward_model="CudaMlpModel",
                     forward_model_params={"hidden": 64, "train_params": {"epochs": 30, "batch_size": 32, "lr": 2e-3}},
                     training_iterations=3, append_data=True)
    cfg = json.load(open(path))
    cfg["controller_params"]["num_simulated_trajectories"] = 128
    cfg["controller_params"]["horizon"] = 12
    cfg["controller_params"]["action_sampler_params"]["elites_size"] = 10
    cfg["rollout_params"]["task_horizon"] = 40
    json.dump(cfg, open(path, "w"))
    res = subprocess.run([sys.executable, "-m", "icem_b200.launch", path, "--reference", ref_loader.REFERENCE_ROOT,
                          "--shims", os.path.join(ROOT, "oracle", "shims")], cwd=str(tmp_path), capture_output=True,
                         text=True, timeout=900, env=dict(os.environ, PYTHONPATH=ROOT))
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-3000:]
    lines = [l.strip("' ") for l in res.stdout.splitlines() if "CudaMlpModel.train:" in l]
    assert len(lines) == 3, res.stdout[-3000:]
    assert "40 transitions" in lines[0] and "80 transitions" in lines[1] and "120 transitions" in lines[2]
    first, last = (float(lines[0].split("loss ")[1].split(" -> ")[0]), float(lines[-1].split(" -> ")[1]))
    assert last < 0.5 * first, lines
    ck = tmp_path / "results" / "checkpoints_latest"
    assert ck.exists() and any(f.name.startswith("forward_model") for f in ck.iterdir())
