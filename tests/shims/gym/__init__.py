"""Minimal stand-in for the parts of OpenAI gym (0.2x API) that the reference's env files and gd.py import
(R/envs/redmax_torch_env.py:12-14, R/envs/__init__.py, R/algorithms/gd.py:11-12,41,74).  gym is not installed in
this image and there is no network; this shim is TEST infrastructure (tests/shims is put on sys.path by the
tests that run the reference's own callers), not part of the product."""
from . import spaces, utils                    # noqa: F401
from .core import Env, Wrapper                 # noqa: F401
from .envs.registration import make, register, registry, spec   # noqa: F401


class _Logger:
    def set_level(self, level):
        self.level = level


logger = _Logger()
