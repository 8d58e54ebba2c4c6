"""gym_softrobot_b200 — B200-native batched simulator for gym-softrobot's physics step.

Keeps the reference's Gymnasium ids (`/root/reference/gym_softrobot/__init__.py:6-80`)
for the envs whose hot path is built, and adds batched
vector envs.  All physics runs in hand-written sm_100a CUDA kernels behind the
C-ABI in include/softrod.h; importing this package never touches oracle/.
"""
from .compat import HAVE_GYMNASIUM
from .version import VERSION as __version__

# id -> (entry point, kwargs); same ids / kwargs as the reference registry
REGISTRY = {
    "SoftPendulum-v0": ("gym_softrobot_b200.envs.soft_pendulum:SoftPendulumEnv", {}),
    "SoftPendulum3D-v0": ("gym_softrobot_b200.envs.soft_pendulum_3d:SoftPendulum3DEnv", {}),
    "OctoArmSingle-v0": ("gym_softrobot_b200.envs.arm_single:ArmSingleEnv", {}),
    "OctoFlat-v0": ("gym_softrobot_b200.envs.octo_flat:FlatEnv", {}),
    "OctoFlatLite-v0": ("gym_softrobot_b200.envs.octo_flat:FlatEnv", dict(n_arm=1, n_action=8)),
    "ContinuumSnake-v0": ("gym_softrobot_b200.envs.snake:ContinuumSnakeEnv", {}),
    "SoftArmTracking-v0": ("gym_softrobot_b200.envs.soft_arm_tracking:SoftArmTrackingEnv", {}),
    "OctoCrawl-v0": ("gym_softrobot_b200.envs.octo_crawl:CrawlEnv", {}),
    "OctoReach-v0": ("gym_softrobot_b200.envs.octo_reach:ReachEnv", {}),
    "OctoArmTwo-v0": ("gym_softrobot_b200.envs.arm_two:ArmTwoEnv", {}),
    "OctoArmPush-v0": ("gym_softrobot_b200.envs.arm_push:ArmPushEnv", {}),
    "OctoArmPush-v1": ("gym_softrobot_b200.envs.arm_push:ArmPushEnv", dict(mode="continuous")),
    "OctoArmPullWeight-v0": ("gym_softrobot_b200.envs.arm_push:ArmPullWeightEnv", dict(mode="continuous")),
}
VECTOR_REGISTRY = {
    "SoftPendulum-v0": ("gym_softrobot_b200.envs.soft_pendulum:SoftPendulumVectorEnv", {}),
    "SoftPendulum3D-v0": ("gym_softrobot_b200.envs.soft_pendulum_3d:SoftPendulum3DVectorEnv", {}),
    "OctoArmSingle-v0": ("gym_softrobot_b200.envs.arm_single:ArmSingleVectorEnv", {}),
    "OctoFlat-v0": ("gym_softrobot_b200.envs.octo_flat:OctoFlatVectorEnv", {}),
    "OctoFlatLite-v0": ("gym_softrobot_b200.envs.octo_flat:OctoFlatVectorEnv", dict(n_arm=1, n_action=8)),
    "ContinuumSnake-v0": ("gym_softrobot_b200.envs.snake:ContinuumSnakeVectorEnv", {}),
    "SoftArmTracking-v0": ("gym_softrobot_b200.envs.soft_arm_tracking:SoftArmTrackingVectorEnv", {}),
    "OctoCrawl-v0": ("gym_softrobot_b200.envs.octo_crawl:OctoCrawlVectorEnv", {}),
    "OctoReach-v0": ("gym_softrobot_b200.envs.octo_reach:OctoReachVectorEnv", {}),
    "OctoArmTwo-v0": ("gym_softrobot_b200.envs.arm_two:ArmTwoVectorEnv", {}),
    "OctoArmPush-v0": ("gym_softrobot_b200.envs.arm_push:ArmPushVectorEnv", {}),
    "OctoArmPush-v1": ("gym_softrobot_b200.envs.arm_push:ArmPushVectorEnv", dict(mode="continuous")),
    "OctoArmPullWeight-v0": ("gym_softrobot_b200.envs.arm_push:ArmPushVectorEnv",
                             dict(mode="continuous", pull_weight=True, time_step=2.5e-5)),
}


def _load(entry_point):
    import importlib
    mod, attr = entry_point.split(":")
    return getattr(importlib.import_module(mod), attr)


def make(env_id, **kwargs):
    """`gym.make(env_id)` equivalent that works without gymnasium installed."""
    entry, kw = REGISTRY[env_id]
    return _load(entry)(**{**kw, **kwargs})


def make_vec(env_id, n_env, **kwargs):
    """Batched env: N independent copies of `env_id` advanced by one kernel launch per step."""
    entry, kw = VECTOR_REGISTRY[env_id]
    return _load(entry)(n_env, **{**kw, **kwargs})


if HAVE_GYMNASIUM:  # pragma: no cover - gymnasium is not in the build image
    from gymnasium.envs.registration import register, registry as _registry

    for _id, (_entry, _kw) in REGISTRY.items():
        if _id not in _registry:
            register(id=_id, entry_point=_entry, kwargs=_kw)
