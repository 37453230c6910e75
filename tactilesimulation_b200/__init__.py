"""tactilesimulation_b200: a B200-native batched differentiable tactile simulator that drops in
behind the plugin surface of eanswer/TactileSimulation (redmax_py.Simulation +
envs/redmax_torch_functions.py) for ONE path: the per-timestep implicit simulation step, the dense
tactile force field and the reverse-time adjoint.  CUDA only -- there is no CPU fallback."""
from ._lib import TactileSimError  # noqa: F401
from .scene import Scene, SceneError, compile_scene  # noqa: F401

__all__ = ["TactileSimError", "Scene", "SceneError", "compile_scene", "Simulation", "BatchedSim",
           "StepSimFunction", "EpisodicSimFunction"]


def __getattr__(name):
    # torch-dependent pieces are imported lazily so that the scene compiler works without torch
    if name == "BatchedSim":
        from .sim import BatchedSim
        return BatchedSim
    if name == "Simulation":
        from .redmax import Simulation
        return Simulation
    if name in ("StepSimFunction", "EpisodicSimFunction"):
        from . import torch_functions
        return getattr(torch_functions, name)
    raise AttributeError(name)
