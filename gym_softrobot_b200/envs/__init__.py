from .soft_pendulum import SoftPendulumEnv, SoftPendulumVectorEnv, pendulum_init_params
