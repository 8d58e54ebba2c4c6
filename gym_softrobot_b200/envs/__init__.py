from .soft_pendulum import SoftPendulumEnv, SoftPendulumVectorEnv, pendulum_init_params
from .soft_pendulum_3d import SoftPendulum3DEnv, SoftPendulum3DVectorEnv, pendulum3d_init_params
from .arm_single import ArmSingleEnv, ArmSingleVectorEnv, arm_contact_params, curvature_interp_matrix
from .octo_flat import FlatEnv, OctoFlatVectorEnv, octopus_init_params, padded_curvature_interp_matrix, count_crossings
from .snake import ContinuumSnakeEnv, ContinuumSnakeVectorEnv, snake_contact_params, beta_spline_matrix, projected_forward_velocity
from .soft_arm_tracking import SoftArmTrackingEnv, SoftArmTrackingVectorEnv, target_trajectory
from .arm_push import ArmPushEnv, ArmPullWeightEnv, ArmPushVectorEnv, arm_push_node_masses, arm_push_energy_tables
from .octo_crawl import CrawlEnv, OctoCrawlVectorEnv, crawl_init_params
from .octo_reach import ReachEnv, OctoReachVectorEnv, es_longitudinal_positions
from .arm_two import ArmTwoEnv, ArmTwoVectorEnv, two_arm_init_params, activation_interp_matrix
