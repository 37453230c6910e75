"""Batched environment front-ends over the B200 path: all four gym envs of the reference."""
from .dclaw_rotate import BatchedDClawRotateEnv  # noqa: F401
from .stable_grasp import BatchedStableGraspEnv  # noqa: F401
from .tactile_insertion import BatchedTactileInsertionEnv  # noqa: F401
from .tactile_push import BatchedTactilePushEnv, push_observation, push_reward  # noqa: F401
