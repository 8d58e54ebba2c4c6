/* TEST INFRASTRUCTURE (oracle) — not product code.  See rod_oracle.c. */
#ifndef ROD_ORACLE_H
#define ROD_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

/* boundary-condition kinds on node 0 / element 0 */
enum {
  RO_BC_FREE = 0,
  RO_BC_ONE_END_FIXED = 1,   /* PyElastica OneEndFixedBC */
  RO_BC_PENDULUM_SLIDER = 2, /* reference soft_pendulum/build.py:65-85 */
  RO_BC_MOVING_BASE = 3      /* reference soft_pendulum_3d/build.py:23-40 */
};

typedef struct {
  int n_elem;
  double start[3], direction[3], normal[3];
  double base_length, base_radius, density, youngs_modulus;
  double shear_modulus;    /* <=0: PyElastica default (see shear_convention) */
  int shear_convention;    /* 0: E/(2(1+nu)), nu=.5 ; 1: E/(1+nu) (SURVEY B-4) */
  double dt;
  double gravity[3];
  double damping_constant; /* AnalyticalLinearDamper; <0 disables */
  int laplace_filter_order; /* 0 disables */
  int bc_kind;
  int point_force_on_base; /* SoftPendulum: F_ext[0,0] = action (assignment) */
  int damping_before_constraints; /* 1 = [dampen_rates, constrain_rates]; 0 = call order of the envs */
  /* RodPlaneContactWithAnisotropicFriction on Plane(origin, normal); contact_on = 0 disables */
  int contact_on;
  int contact_before_forcing; /* 1 = contact operator runs before gravity/forcing in synchronize; 0 = after (the envs) */
  double plane_origin[3], plane_normal[3];
  double contact_k, contact_nu, slip_velocity_tol, surface_tol;
  double static_mu[3], kinetic_mu[3]; /* forward, backward, sideways */
  /* PyElastica MuscleTorques (travelling wave; [PE-recall], call site
   * /root/reference/gym_softrobot/envs/snake/continuum_snake.py:186-198,325-337); muscle_on = 0 disables.
   * beta(s) and the wave number are set through ro_muscle() (what `set_action` rebuilds). */
  int muscle_on;
  double muscle_period, muscle_ramp_up_time, muscle_phase_shift, muscle_direction[3];
  /* MuscleTorquesWithVaryingBetaSplines (/root/reference/gym_softrobot/utils/custom_elastica/muscle_torque/
   * muscle_torques_with_bspline.py:46-228; call sites envs/soft_arm/soft_arm_tracking.py:366-400): bit d of
   * spline_dir_mask = one forcing instance on material direction d (0 normal, 1 binormal, 2 tangent). */
  int spline_dir_mask, spline_n_ctrl;
  double spline_scale, spline_max_rate;
  /* tapered rod: base_radius array = np.linspace(base_radius, tip_radius, n_elem)
   * (/root/reference/gym_softrobot/envs/octopus/build_muscle_octopus.py:61-63); <= 0: uniform */
  double tip_radius;
  /* 1: the taper is given on the nodes, np.linspace(base, tip, n_elem + 1), and each element takes the mean of its
   * two nodes (/root/reference/gym_softrobot/envs/octopus/arm_push_env.py:161-175) */
  int taper_node_mean;
} ro_config;

typedef struct ro_rod ro_rod;

ro_rod *ro_create(const ro_config *cfg);
void ro_destroy(ro_rod *);
/* advance n substeps of PositionVerlet; `action` is the point force (kind 2)
 * and base_pos/base_vel the commanded base (kind 3; may be NULL otherwise) */
void ro_substeps(ro_rod *, int n_substeps, double action, const double *base_pos,
                 const double *base_vel);
double ro_time(const ro_rod *);
int ro_n_elem(const ro_rod *);
/* pointers into the rod's arrays, reference layout (3,n+1) / (3,3,n) / (3,n) row-major */
double *ro_position(ro_rod *);
double *ro_velocity(ro_rod *);
double *ro_director(ro_rod *);
double *ro_omega(ro_rod *);
double *ro_tangents(ro_rod *);
double *ro_kappa(ro_rod *);
double *ro_sigma(ro_rod *);
double *ro_dilatation(ro_rod *);
double *ro_rest_kappa(ro_rod *);
double *ro_external_forces(ro_rod *); /* (3,n+1) constant extra nodal load added every substep */
double *ro_external_torques(ro_rod *); /* (3,n) constant extra element couple (material frame) added every substep */
/* ControllableFixConstraint (/root/reference/gym_softrobot/envs/octopus/controllable_constraint.py:42-69):
 * after the dynamic step, v[:, index] and omega[:, index] are scaled by (1 - ratio); up to 8 slots */
void ro_set_sucker(ro_rod *, int slot, int index, double ratio);
/* COOMM TransverseMuscle under ApplyMuscles (restated from the published model — the package is not in
 * /root/reference; see rod_oracle.c:apply_tm_muscle): rest_muscle_area_k = (rest radius_k / radius_ref)^2,
 * scalar activation broadcast over the elements (crawl_env.py:242, arm_push_env.py:259,271).  max_stress = 0 disables. */
void ro_set_tm_muscle(ro_rod *, double max_stress, double radius_ref);
void ro_set_tm_activation(ro_rod *, double activation);
/* General muscle layers with per-element activations (OctoReach-v0 / OctoArmTwo-v0): slot 0..2, kind 1 = longitudinal
 * at material-frame offset (px, py) * radius, 2 = transverse (max_stress is handed over SIGNED: the transverse class
 * passes -max_muscle_stress); ro_muscle_activation: (3, n) activations, one row per slot.  Same unpinned restatement. */
void ro_set_muscle_layer(ro_rod *, int slot, int kind, double max_stress, double radius_ref, double px, double py);
double *ro_muscle_activation(ro_rod *);
double *ro_mass(ro_rod *);
double *ro_internal_forces(ro_rod *);
double *ro_internal_torques(ro_rod *);
double *ro_radius(ro_rod *);
double *ro_muscle(ro_rod *);
double *ro_spline_points(ro_rod *); /* [3][2P+1]: P targets (points_func_array), P cached values, initial-call flag */
double *ro_spline_magnitude(ro_rod *); /* [3][n]: torque_magnitude_cache of each instance */ /* [1 + n]: wave_number, then beta(s_k) at s_k = cumsum(rest_lengths)_k / L */

/* SoftPendulum-v0 env step on top of the rod: follows
 * /root/reference/gym_softrobot/envs/soft_pendulum/soft_pendulum.py:176-251 */
void ro_softpendulum_obs(ro_rod *, float prev_action, float obs[4]);
void ro_softpendulum_step(ro_rod *, float action, int step_skip, double final_time,
                          float obs[4], double *reward, int *terminated, int *truncated);

/* batched helper for the CPU baseline: n_env independent SoftPendulum rods,
 * pthread-parallel over contiguous env ranges (n_threads<=0: all cores) */
void ro_softpendulum_step_batch(ro_rod **rods, int n_env, const float *actions, int step_skip,
                                double final_time, float *obs, double *reward, int *terminated,
                                int *truncated, int n_threads);
int ro_max_threads(void);

/* ---- multi-rod assembly (octopus): n_arm rods + rigid Cylinder head + FixedJoint2Rigid joints + BodyBoundaryCondition
 * (/root/reference/gym_softrobot/envs/octopus/build.py:52-217, utils/custom_elastica/joint.py, constraint.py) */
#define RO_MAX_ARMS 16
typedef struct {
  int n_arm, has_head;
  double dt;
  double head_start[3], head_direction[3], head_normal[3];
  double head_length, head_radius, head_density;
  double joint_k, joint_nu, joint_kt, joint_radius;
  double joint_angle_deg[RO_MAX_ARMS];
} ro_asm_config;
typedef struct ro_assembly ro_assembly;
ro_assembly *ro_asm_create(const ro_config *arm_cfgs /* [n_arm] */, const ro_asm_config *cfg);
void ro_asm_destroy(ro_assembly *);
void ro_asm_substeps(ro_assembly *, int n_substeps);
/* OneEndFixedBC(constrained_position_idx=(0,), constrained_director_idx=(0,)) on the rigid head, registered after the
 * BodyBoundaryCondition (/root/reference/gym_softrobot/envs/octopus/reach_env.py:128-132): the head is pinned at the
 * pose it has when this is called, all its rates are zeroed. */
void ro_asm_set_head_fixed(ro_assembly *, int on);
ro_rod *ro_asm_arm(ro_assembly *, int i);   /* owned by the assembly */
double ro_asm_time(const ro_assembly *);
double *ro_asm_head_position(ro_assembly *);  /* (3) */
double *ro_asm_head_velocity(ro_assembly *);  /* (3) */
double *ro_asm_head_director(ro_assembly *);  /* (3,3) rows */
double *ro_asm_head_omega(ro_assembly *);     /* (3) */
/* pthread-parallel batches of independent rods / assemblies (CPU baselines of bench.py --config 3/4/5) */
void ro_substeps_batch(ro_rod **rods, int n, int n_substeps, int n_threads);
void ro_asm_substeps_batch(ro_assembly **asms, int n, int n_substeps, int n_threads);

#ifdef __cplusplus
}
#endif
#endif
