"""icem_b200 -- B200-native iCEM sampling-MPC inner loop behind the reference's Controller plugin API.

Layout (only what the hot path needs):
  csrc/          hand-written sm_100a CUDA kernels + the C ABI (include/icem_b200.h)
  lib/           the built libicem_b200.so (in-tree, git-ignored)
  _lib.py        ctypes binding of the C ABI (fails loudly when the library or a GPU is missing)
  planner.py     thin object wrapper over one icem_planner_t handle
  api.py         stand-alone copies of the reference's ABC signatures (used when the reference
                 package is not importable)
  controller.py  MpcICemB200: drop-in for controllers/icem.py::MpcICem
  models.py      CUDA-capable forward models (registry names for models/__init__.py::models_dict)
  envs.py        stand-in environments (MuJoCo/gym are not available) stepping the device model
  launch.py      registers the plugin names and runs the reference's unchanged main.main()
  distributed.py one-process-per-GPU sharding helpers (NCCL unique-id exchange over torch.distributed)
"""
__version__ = "0.1.0"
