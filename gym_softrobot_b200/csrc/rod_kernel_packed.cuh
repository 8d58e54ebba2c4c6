// rod_kernel_packed.cuh — CTA-packed variant of the fused K-substep kernel (sm_100a).
//
// Mapping: one thread owns node j and element j of a rod (j = 0..n; thread n owns the
// tip node only), and rods are packed back to back across the CTA: a 256-thread CTA
// holds floor(NT/(n+1)) rods (5 rods x 51 threads = 255 lanes for n = 50, NT = 256), so ~99 %
// of the lanes that issue FP64 instructions do useful work (the warp-per-rod kernel in
// rod_kernels.cuh uses 25 of 32).  The whole per-thread state (x, v, Q, w: 18 values)
// stays in registers for all K substeps; the neighbour values needed by the stencils
// (x,v,Q of j+1, x of j+2; stress and couple of j-1) go through shared memory, with two
// CTA barriers per substep.  Same arithmetic as the fast path of rod_kernels.cuh
// (reference boundary
// /root/reference/gym_softrobot/envs/soft_pendulum/soft_pendulum.py:183-184).
#pragma once
#include "rod_kernels.cuh"

namespace sr {

constexpr int PACKED_MAX_THREADS = 1024;

// row stride of the exchange arrays: slot NT is a permanent zero (what the stencils see to
// the left of element 0), so the base thread needs no select when it reads "j-1"
// shared-memory words: x(3) v(3) Q(9) | s(3) N(3), each row NT + 2 wide (+ 6 joint-reaction rows for
// assemblies, + 1 element-length row for the spline-torque forcing of the contact / forcing variant)
constexpr int packed_smem_words(int nt, bool multi, bool torque = false, bool lmus = false) {
  return ((multi ? 27 : 24) + (lmus ? 7 : 0)) * (nt + 2);   // LMUS: + curvature (3), radius (1), muscle couple (3)
}

// LAPLACE / MOVING: compile the LaplaceDissipationFilter passes and the moving-base controller in
// (SoftPendulum3D); kept out of the instantiation used by the other models so they do not pay
// their registers / code size.  TORQUE (with CONTACT, single rod): the muscle-torque forcings
// (travelling wave, spline) and the element-length row they need.  CONTACT: plane contact with anisotropic friction and per-env rest
// curvature (octopus-arm models): two more neighbour exchanges per substep.  MULTI: several rods per
// env plus one rigid head thread, coupled by FixedJoint2Rigid spring/torque joints.
// FASTONLY: no warp votes / libm fallbacks in the substep body (they split it into basic blocks and cost ~5 %);
// a thread that sees an argument outside the polynomials' range raises a flag instead, the env's state is
// then NOT written back (global memory still holds the pre-launch state) and redo[env] is set: the safe
// instantiation, launched right behind with redo_filter = 1, steps exactly those envs.
// VARY (ElemConst / ET_* in rod_kernels.cuh): element / node / Voronoi constants come from a per-handle table in HBM (tapered rods: base_radius an array,
// /root/reference/gym_softrobot/envs/octopus/build_muscle_octopus.py:61-63) into per-thread registers instead of
// the constant bank; only instantiated for the safe contact / multi-rod variants (168 registers, one CTA per SM).

// LMUS (with VARY and MULTI): COOMM's `ApplyMuscles` over the whole layer set of create_es_muscle_layers
// (envs/octopus/build.py:292-338) with per-element activations (OctoReach-v0 / OctoArmTwo-v0): two off-axis longitudinal
// muscles + the transverse muscle, evaluated every substep; one more neighbour exchange (curvature and radius of the
// neighbouring elements) and barrier per substep; optional OneEndFixedBC on the rigid head.

template <typename T, int NT, int MINB, bool LAPLACE, bool MOVING, bool CONTACT, bool MULTI, bool TORQUE = false,
          bool FASTONLY = false, bool VARY = false, bool LMUS = false>
__global__ void __launch_bounds__(NT, MINB)
rod_packed_kernel(const __grid_constant__ RodArgs<T> A, int rods_per_cta) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T *sh = reinterpret_cast<T *>(smem_raw);
  constexpr int RS = NT + 2;
  T *sh_x = sh, *sh_v = sh + 3 * RS, *sh_Q = sh + 6 * RS, *sh_s = sh + 15 * RS, *sh_N = sh + 18 * RS;

  const int tid = threadIdx.x;
  const int n = A.n_elem, stride = A.stride, tpr = n + 1;  // threads per rod
  // env group = n_rod rods of tpr threads (+ one head thread); single-rod models: group = rod
  const int n_rod = MULTI ? A.n_rod : 1, has_head = MULTI ? A.has_head : 0;
  const int G = n_rod * tpr + has_head;
  const int r = tid / G, u = tid - r * G;           // env slot inside the CTA, thread inside the group
  const bool is_head = MULTI && has_head && (u == G - 1);
  const int arm = (MULTI && !is_head) ? u / tpr : 0;
  const int j = is_head ? 0 : u - arm * tpr;
  const int env = blockIdx.x * rods_per_cta + r;    // rods_per_cta counts env groups
  const bool in_grid = (r < rods_per_cta) && (env < A.n_env);
  bool selected = true;
  if (!FASTONLY && A.redo_filter) {   // fallback launch: flagged envs only; most CTAs have none and leave
    selected = in_grid && A.redo[env] != 0;
    if (!__syncthreads_or(selected)) return;
  }
  const bool live = in_grid && selected;
  bool dom_bad = false;               // FASTONLY: an argument left the fast-math domain
  const bool active = live && !is_head;             // a thread that owns a node / element of a rod
  const int rod = env * n_rod + arm;                // global rod slot in the state arrays
  const int t_head = r * G + G - 1;                 // where this env's head publishes its state
  T *sh_J = sh + 21 * RS;                           // MULTI: joint reactions on the head, 6 rows
  const bool elem_ok = active && j < n, vor_ok = active && j < n - 1, first = (j == 0);
  // neighbour slots (clamped so that idle / edge threads read something harmless)
  const int t_next = (active && j < n) ? tid + 1 : tid;
  const int t_next2 = (active && j < n - 1) ? tid + 2 : t_next;
  const int t_prev = (active && j > 0) ? tid - 1 : NT;   // NT = the zero slot
  // Lean FP64 kernels exchange through array-of-structures records instead of the rows above, so that neighbour
  // data moves with 128-bit accesses (24 LDS/STS per substep instead of 45): per thread 18 doubles
  // {x0 x1 | x2 v0 | v1 v2 | Q0 Q1 | Q2 Q3 | Q4 Q5 | Q6 Q7 | Q8 - | - -} (stride 144 B: conflict-free for LDS.128)
  // followed by 6-double records {s0 s1 | s2 N0 | N1 N2} (stride 48 B).  Record NT of the second array is the zero slot.
  constexpr bool AOS = sizeof(T) == 8 && !LAPLACE && !CONTACT && !MULTI;
  T *rec = sh, *sn = sh + 18 * RS;
  if (AOS) { if (tid < 6) sn[6 * NT + tid] = T(0); }
  else if (tid < 6) sh_s[(tid % 3) * RS + NT + (tid / 3) * (3 * RS)] = T(0);  // zero slots of s and N
  T *sh_M = sh + 27 * RS;             // LMUS: rows 0-2 curvature (zero slot NT), 3 element radius, 4-6 muscle couple
  if (LMUS && tid < 3) sh_M[tid * RS + NT] = T(0);
  __syncthreads();
  // Barriers inside the substep loop: data only crosses threads of the same env group (a rod; an assembly's rods +
  // head), so the synchronisations are per group — named barriers over the warps that hold the group's threads; a
  // warp holding the end of one group and the start of the next takes part in both, in group order.  Groups then
  // drift apart and fill each other's pipeline bubbles (lean kernel: +13 % with one 512-thread CTA per SM).  Needs
  // G >= 32 (a warp touches at most two groups) and <= 15 groups per CTA (barrier ids 1..15).
  int bar_id0 = 0, bar_cnt0 = 0, bar_id1 = 0, bar_cnt1 = 0;
  // (one group per CTA: nothing to decouple; the filtered model's 9 synchronisations per substep are cheaper CTA-wide)
  const bool grp_barriers = !LAPLACE && A.sk_rodsync && G >= 32 && rods_per_cta >= 2 && rods_per_cta <= 15;
  if (grp_barriers) {
    const int wp = tid >> 5;
    for (int rr = 0; rr < rods_per_cta; rr++) {
      const int wa = (rr * G) >> 5, wb = (rr * G + G - 1) >> 5;
      if (wp >= wa && wp <= wb) {
        if (bar_cnt0 == 0) { bar_id0 = rr + 1; bar_cnt0 = 32 * (wb - wa + 1); }
        else { bar_id1 = rr + 1; bar_cnt1 = 32 * (wb - wa + 1); }
      }
    }
  }
  auto grp_sync = [&]() {
    if (!grp_barriers) { __syncthreads(); return; }
    if (bar_cnt0) asm volatile("bar.sync %0, %1;" ::"r"(bar_id0), "r"(bar_cnt0) : "memory");
    if (bar_cnt1) asm volatile("bar.sync %0, %1;" ::"r"(bar_id1), "r"(bar_cnt1) : "memory");
  };

  // The contact models work in a frame whose z axis is the plane normal (every dot / cross product with the normal
  // becomes a component pick).  When the configured normal is not +z (ContinuumSnake: +y) lab-frame vectors —
  // positions, velocities, director rows, BC anchors, the stale tangents, observations — are rotated by A.lab2int on
  // load and back on store; for an axis-aligned normal that is a signed permutation, i.e. exact.  Gravity and the
  // plane origin arrive rotated from the host; omega, kappa, sigma live in the material frame and are untouched.
  const bool rot_frame = CONTACT && A.rot_on;
  auto to_int = [&](T (&u)[3]) {
    if (!rot_frame) return;
    const T a = u[0], b = u[1], c = u[2];
#pragma unroll
    for (int i = 0; i < 3; i++) u[i] = fma(A.lab2int[3 * i + 2], c, fma(A.lab2int[3 * i + 1], b, A.lab2int[3 * i] * a));
  };
  auto to_lab = [&](T (&u)[3]) {      // transpose
    if (!rot_frame) return;
    const T a = u[0], b = u[1], c = u[2];
#pragma unroll
    for (int i = 0; i < 3; i++) u[i] = fma(A.lab2int[6 + i], c, fma(A.lab2int[3 + i], b, A.lab2int[i] * a));
  };
  auto rows_to_int = [&](T (&M)[9]) {
#pragma unroll
    for (int i = 0; i < 3; i++) { T u[3] = {M[3 * i], M[3 * i + 1], M[3 * i + 2]}; to_int(u); M[3 * i] = u[0]; M[3 * i + 1] = u[1]; M[3 * i + 2] = u[2]; }
  };
  auto rows_to_lab = [&](T (&M)[9]) {
#pragma unroll
    for (int i = 0; i < 3; i++) { T u[3] = {M[3 * i], M[3 * i + 1], M[3 * i + 2]}; to_lab(u); M[3 * i] = u[0]; M[3 * i + 1] = u[1]; M[3 * i + 2] = u[2]; }
  };

  T x[3] = {T(0), T(0), T(0)}, v[3] = {T(0), T(0), T(0)}, w[3] = {T(0), T(0), T(0)};
  T Q[9] = {T(1), T(0), T(0), T(0), T(1), T(0), T(0), T(0), T(1)};
  T *st = A.state + (size_t)(active ? rod : 0) * N_FIELDS * stride;
  if (active) {
#pragma unroll
    for (int c = 0; c < 3; c++) {
      x[c] = st[(F_POS + c) * stride + j];
      v[c] = st[(F_VEL + c) * stride + j];
      w[c] = st[(F_OMEGA + c) * stride + j];   // slot n holds 0 (never written by anyone)
    }
#pragma unroll
    for (int c = 0; c < 9; c++) Q[c] = st[(F_DIR + c) * stride + j];  // slot n holds I
    to_int(x); to_int(v); rows_to_int(Q);
  }
  // FP32: absolute positions resolve an element's strain only to ulp(x)/dl ~ 3e-6, which excites stiff
  // modes; the edge vector dx = x_{j+1} - x_j is therefore carried as state of its own (ulp 2e-9) and
  // advanced with the same velocities as the nodes.  x is still integrated for outputs, BCs and contact.
  constexpr bool EDGE = sizeof(T) == 4;
  T ed[3] = {T(0), T(0), T(0)};
  T hh_prev = T(0);   // step of the kinematic update whose edge increment is still pending
  if (EDGE && active) {
#pragma unroll
    for (int c = 0; c < 3; c++) ed[c] = st[(F_EDGE + c) * stride + j];
    to_int(ed);
  }
  T *hd = (MULTI && is_head && live) ? A.head + (size_t)env * HEAD_DIM : nullptr;
  if (MULTI && hd) {     // (assemblies stand on a z-normal plane: sr_create checks, no rotation here)
#pragma unroll
    for (int c = 0; c < 3; c++) { x[c] = hd[c]; v[c] = hd[3 + c]; w[c] = hd[15 + c]; }
#pragma unroll
    for (int c = 0; c < 9; c++) Q[c] = hd[6 + c];
  }
  T fj[3] = {T(0), T(0), T(0)}, tj[3] = {T(0), T(0), T(0)};   // joint force on node 0 / torque on element 0
  // time-step multipliers are zero where there is nothing to integrate (tip thread's
  // pseudo-element, idle threads): no selects in the update expressions.
  // (c_v is 1 when the damper is off)
  // K: where the element constants come from — the kernel parameters (uniform rod) or this thread's table row
  ElemConst<T> ec;
  T mass_j = A.mass * ((j == 0 || j == n) ? T(0.5) : T(1)), mass_j1 = A.mass * ((j + 1 == n) ? T(0.5) : T(1));
  if (VARY) {
    const T *tab = A.elem_tab;
    const int je = min(j, n - 1), jv = min(j, n - 2), es = A.stride;
    ec.rest_len = tab[ET_REST_LEN * es + je]; ec.inv_rest_len = T(1) / ec.rest_len;
    ec.rest_vor = tab[ET_REST_VOR * es + jv]; ec.inv_rest_vor = T(1) / ec.rest_vor;
    ec.S[0] = ec.S[1] = tab[ET_S0 * es + je]; ec.S[2] = tab[ET_S2 * es + je];
    ec.B[0] = ec.B[1] = tab[ET_B0 * es + jv]; ec.B[2] = tab[ET_B2 * es + jv];
    ec.J[0] = ec.J[1] = tab[ET_J0 * es + je]; ec.J[2] = tab[ET_J2 * es + je];
    ec.logc_w[0] = ec.logc_w[1] = tab[ET_LOGCW0 * es + je]; ec.logc_w[2] = tab[ET_LOGCW2 * es + je];
    ec.vol_over_pi = tab[ET_VOL_PI * es + je];
    ec.isotropic = 1;
#pragma unroll
    for (int i = 0; i < 3; i++) {
      ec.S_over_l[i] = ec.S[i] * ec.inv_rest_len; ec.Jinv[i] = T(1) / ec.J[i]; ec.c_w[i] = exp_ref<T>(ec.logc_w[i]);
    }
    mass_j = tab[ET_MASS * es + min(j, n)]; mass_j1 = tab[ET_MASS * es + min(j + 1, n)];
  }
  // COOMM TransverseMuscle (see the force block below): activation x (-max_stress x rest area), constant during a launch
  T tm_gain = T(0);
  if constexpr (VARY) {
    if (A.tm_act && elem_ok) tm_gain = A.tm_act[rod] * A.elem_tab[ET_TM * A.stride + min(j, n - 1)];
  }
  // COOMM muscle layers, per-element activations (constant during a launch): gains of the two longitudinal muscles
  // = activation x max_stress x rest area, spacing of the neighbouring element centres, rest Voronoi length to the left
  T lm_g0 = T(0), lm_g1 = T(0), lm_inv_ds = T(0), lm_rvl = T(0);
  if constexpr (LMUS) {
    if (elem_ok) {
      const T *ma = A.mus_act + (size_t)rod * 3 * n;
      const T area = A.elem_tab[ET_TM * A.stride + j];        // -max_stress(TM) x rest_muscle_area
      lm_g0 = ma[j] * area * A.lm_gain;
      lm_g1 = ma[n + j] * area * A.lm_gain;
      tm_gain = ma[2 * n + j] * area;
      const T *rl = A.elem_tab + ET_REST_LEN * A.stride;
      const int ja = max(j - 1, 0), jb = min(j + 1, n - 1);
      // s = cumsum(rest_lengths) - rest_lengths / 2 (muscle.py, restated): s[jb] - s[ja]
      const T ds = T(0.5) * (rl[ja] + rl[jb]) + ((jb - ja == 2) ? rl[j] : T(0));
      lm_inv_ds = T(1) / ds;
      lm_rvl = A.elem_tab[ET_REST_VOR * A.stride + max(j - 1, 0)];
    }
  }
  T lm_cpl[3] = {T(0), T(0), T(0)}, lm_kl[3] = {T(0), T(0), T(0)}, lm_kr[3] = {T(0), T(0), T(0)};
  // ControllableFixConstraint index of this rod: node / element it acts on (python indexing of arrays of n + 1 / n slots)
  int sk_node = A.sucker_index, sk_elem = A.sucker_index;
  if (A.sucker_idx && active) {
    const int si = A.sucker_idx[rod];
    sk_node = si < 0 ? si + n + 1 : si; sk_elem = si < 0 ? si + n : si;
  }
  const auto &K = [&]() -> const auto & { if constexpr (VARY) return ec; else return A; }();
  const T dtim_cv = !active ? T(0) : VARY ? A.dt / mass_j * A.c_v : A.dt_inv_mass * A.c_v * ((j == 0 || j == n) ? T(2) : T(1));
  const T gmask = active ? T(1) : T(0);
  const T dte = elem_ok ? A.dt : T(0);

  const T gam = elem_ok ? st[F_GAMMA * stride + j] : T(1);   // (L/n) / rest_length_j, see F_GAMMA
  T rk[3] = {T(0), T(0), T(0)};   // rest curvature at Voronoi point j (actuation), constant during a launch
  if (CONTACT && A.rest_kappa && vor_ok) {
#pragma unroll
    for (int c = 0; c < 3; c++) rk[c] = A.rest_kappa[((size_t)rod * 3 + c) * stride + j];
  }
  // MuscleTorques (continuum_snake.py:186-198; PyElastica external_forces.MuscleTorques): element k gets
  // Q_k d (m_k [k >= 1] - m_{k+1} [k <= n-2]),  m_k = min(1, t/ramp) beta_{n-1-k} sin(w t - kw s_{n-1-k} + phi)
  // (the reference walks the magnitudes tail-to-head); s_i = (i+1)/n on the uniform rest template.
  const bool muscle = TORQUE && A.muscle_on && elem_ok;
  double mus_t = 0.0, mus_b0 = 0.0, mus_b1 = 0.0, mus_s0 = 0.0, mus_s1 = 0.0;
  if (TORQUE && A.muscle_on && live) {
    const double *mu = A.muscle + (size_t)env * A.muscle_dim;
    mus_t = mu[0];
    if (muscle) {
      const double kw = mu[1], inv_n = 1.0 / (double)n;
      const int i0 = n - 1 - j, i1 = n - 2 - j;
      if (j >= 1) { mus_b0 = mu[2 + i0]; mus_s0 = kw * ((double)(i0 + 1) * inv_n); }
      if (j <= n - 2) { mus_b1 = mu[2 + i1]; mus_s1 = kw * ((double)(i1 + 1) * inv_n); }
    }
  }
  // MuscleTorquesWithVaryingBetaSplines (muscle_torques_with_bspline.py:126-160,181-228): per enabled
  // material direction d, external_torques[d, k] += mag_d[k]; mag is re-evaluated (not-a-knot cubic through
  // the rate-limited control values, at s = cumsum(current lengths)) in every substep that finds the cached
  // values different from the caller's targets, and kept otherwise — across launches too.
  const bool spl = TORQUE && A.spline_mask != 0;
  __shared__ int sh_need[TORQUE ? 128 : 1];
  T *sh_L = sh + 21 * RS;
  double *sp = (spl && live) ? A.spline + (size_t)env * A.spline_dim : nullptr;
  const int sP = A.spline_p, sCH = 2 * sP + 2;
  T smag[3] = {T(0), T(0), T(0)};
  if (spl && elem_ok) {
#pragma unroll
    for (int d = 0; d < 3; d++)
      if (A.spline_mask >> d & 1) smag[d] = (T)sp[3 * sCH + d * n + j];
  }
  T act0 = T(0), base_vx = T(0), base_vy = T(0);
  float act_f0 = 0.0f, act_f1 = 0.0f;
  if (active && A.action_dim > 0) { act_f0 = A.action[(size_t)env * A.action_dim]; act0 = (T)act_f0; }
  if (active && A.action_dim > 1) act_f1 = A.action[(size_t)env * A.action_dim + 1];
  const bool bc_thread = active && first && A.bc_kind != BC_FREE;
  // Boundary conditions of this build pin node 0 / element 0 by OVERWRITING values after every
  // kinematic update (soft_pendulum/build.py:71-74, OneEndFixedBC, soft_pendulum_3d/build.py:32-35).
  // Applying the overwrite once here and then never moving the pinned quantities is the same
  // thing: the base thread integrates its frame with a zero rotation vector (rot_on = 0; for
  // the slider the free row d2 is invariant under R(0,w1,0) anyway), pinned position components
  // have zero velocity after constrain_rates, and the moving base is re-pinned from registers.
  const T rot_on = bc_thread ? T(0) : T(1);
  T pin_x = T(0), pin_y = T(0);
  if (bc_thread) {
    const T *bc = A.bc + (size_t)rod * BC_DIM;
    if (A.bc_kind == BC_PENDULUM_SLIDER) {
      x[1] = bc[1]; x[2] = bc[2];
#pragma unroll
      for (int m = 0; m < 3; m++) { Q[0 + m] = bc[3 + m]; Q[6 + m] = bc[9 + m]; }
    } else {
#pragma unroll
      for (int c = 0; c < 9; c++) Q[c] = bc[3 + c];
      rows_to_int(Q);
      if (!MOVING || A.bc_kind == BC_ONE_END_FIXED) {
#pragma unroll
        for (int c = 0; c < 3; c++) x[c] = bc[c];
        to_int(x);
      } else {
        T *aux = A.aux + (size_t)env * AUX_DIM;
        if (A.model == MODEL_SOFT_PENDULUM_3D && A.n_substeps > 0) {
          // set_action (soft_pendulum_3d.py:106-120): float32 displacement, float64 clipped position,
          // velocity = actual displacement / (step_skip * time_step); the base jumps at once
          T px = aux[0], py = aux[1];
          T nx = px + (T)__fmul_rn(A.base_step_f32, act_f0), ny = py + (T)__fmul_rn(A.base_step_f32, act_f1);
          nx = fmin(fmax(nx, -A.base_limit), A.base_limit);
          ny = fmin(fmax(ny, -A.base_limit), A.base_limit);
          const T nvx = (nx - px) * A.inv_move_period, nvy = (ny - py) * A.inv_move_period;
          if (FASTONLY) {   // keep global memory at the pre-launch state until the epilogue knows the env is not re-run
            pin_x = nx; pin_y = ny; base_vx = nvx; base_vy = nvy;
          } else {
            aux[3] = nvx; aux[4] = nvy; aux[5] = T(0);
            aux[0] = nx; aux[1] = ny;
          }
        }
        if (!(FASTONLY && A.model == MODEL_SOFT_PENDULUM_3D && A.n_substeps > 0)) {
          pin_x = aux[0]; pin_y = aux[1]; base_vx = aux[3]; base_vy = aux[4];
        }
        x[0] = pin_x; x[1] = pin_y; x[2] = bc[2];
        if (EDGE) {   // the base jumped: element 0's edge is re-derived from the stored node 1
#pragma unroll
          for (int c = 0; c < 3; c++) ed[c] = st[(F_POS + c) * stride + 1] - x[c];
        }
      }
    }
  }
  const bool moving = MOVING && bc_thread && A.bc_kind == BC_MOVING_BASE;

  // branch-free form for the fast-only kernels: which rate components the BC thread zeroes, decided once
  const bool pin_slider = bc_thread && A.bc_kind == BC_PENDULUM_SLIDER;
  const bool pin_fixed = bc_thread && A.bc_kind != BC_PENDULUM_SLIDER && (!MOVING || A.bc_kind == BC_ONE_END_FIXED);
  auto constrain_rates = [&]() {
    if (FASTONLY && !MOVING) {
      const bool z12 = pin_slider || pin_fixed;   // v_y, v_z, w_x, w_z are pinned by both; v_x, w_y by the clamp only
      v[0] = pin_fixed ? T(0) : v[0]; v[1] = z12 ? T(0) : v[1]; v[2] = z12 ? T(0) : v[2];
      w[0] = z12 ? T(0) : w[0]; w[1] = pin_fixed ? T(0) : w[1]; w[2] = z12 ? T(0) : w[2];
      return;
    }
    if (bc_thread) {
      if (A.bc_kind == BC_PENDULUM_SLIDER) {
        v[1] = T(0); v[2] = T(0); w[0] = T(0); w[2] = T(0);
      } else if (!MOVING || A.bc_kind == BC_ONE_END_FIXED) {
#pragma unroll
        for (int c = 0; c < 3; c++) { v[c] = T(0); w[c] = T(0); }
      } else {
        // [constrain, dampen] order: the damper rescales the commanded base velocity too
        const T sc = A.damp_first ? T(1) : A.c_v;
        v[0] = base_vx * sc; v[1] = base_vy * sc; v[2] = T(0);
#pragma unroll
        for (int c = 0; c < 3; c++) w[c] = T(0);
      }
    }
  };
  // x += hh v ; Q <- R(hh w) Q (merged half steps, see rod_kernels.cuh)
  auto kinematic = [&](T hh, T eps) {
    const T hw = hh * rot_on;
    T a0 = hw * w[0], a1 = hw * w[1], a2 = hw * w[2];
#pragma unroll
    for (int c = 0; c < 3; c++) x[c] = fma(hh, v[c], x[c]);
    if (moving) { x[0] = pin_x; x[1] = pin_y; }
    T q = fma(a2, a2, fma(a1, a1, a0 * a0));
    // Range handling is per thread, never a warp vote: an env's bits must not depend on which other envs share
    // its warps.  Inside the range the safe variant evaluates exactly what the fast-only variant evaluates.
    const bool rot_out = !(q <= T(kNarrowRotQ));
    if (FASTONLY) {
      dom_bad = dom_bad || rot_out;
      rotate_directors_fast<T, true>(A.poly, a0, a1, a2, q, eps, Q);
    } else if (!rot_out) rotate_directors_fast<T, true>(A.poly, a0, a1, a2, q, eps, Q);
    else rotate_directors_ref<T>(a0, a1, a2, Q);
    hh_prev = hh;
  };

  // BodyBoundaryCondition on the head (utils/custom_elastica/constraint.py:43-58, 62-85)
  auto head_constrain_values = [&]() {
    if (LMUS && A.head_fixed) return;   // OneEndFixedBC on top: zero rates keep the pose bit for bit (x += hh 0, Q <- R(0) Q)
    if (MULTI && hd) {
      x[2] = hd[18];
      Q[6] = T(0); Q[7] = T(0); Q[8] = T(1);
#pragma unroll
      for (int i = 0; i < 2; i++) {
        T il = rsqrt_nr(Q[3 * i] * Q[3 * i] + Q[3 * i + 1] * Q[3 * i + 1]);   // rows stay ~unit: never 0
        Q[3 * i] *= il; Q[3 * i + 1] *= il; Q[3 * i + 2] = T(0);
      }
    }
  };
  const T h = A.half_dt, dt = A.dt;
  if (A.n_substeps > 0) { kinematic(h, T(1e-14)); head_constrain_values(); }

#pragma unroll 1
  for (int s = 0; s < A.n_substeps; s++) {
    const bool last = (s == A.n_substeps - 1);
    T mtq[3] = {T(0), T(0), T(0)};
    if (TORQUE && A.muscle_on) {
      mus_t += (double)h;   // time of the force evaluation: after the first half step
      if (muscle) {
        const double wt = A.mus_omega * mus_t, fct = fmin(1.0, mus_t / A.mus_ramp);
        const double m0 = (fct * mus_b0) * sin((wt - mus_s0) + A.mus_phase);
        const double m1 = (fct * mus_b1) * sin((wt - mus_s1) + A.mus_phase);
        const T cf = (T)(m0 - m1);
#pragma unroll
        for (int i = 0; i < 3; i++)
          mtq[i] = (Q[3 * i] * A.mus_dir[0] + Q[3 * i + 1] * A.mus_dir[1] + Q[3 * i + 2] * A.mus_dir[2]) * cf;
      }
      mus_t += (double)h;
    }
    // ---- publish what the neighbours need ------------------------------------------
#pragma unroll
    for (int c = 0; c < 3; c++) {
      if (!AOS) { sh_x[c * RS + tid] = EDGE ? ed[c] : x[c]; sh_v[c * RS + tid] = v[c]; }
    }
    if (AOS) {
      double2 *o = reinterpret_cast<double2 *>(rec + 18 * tid);
      o[0] = make_double2(x[0], x[1]); o[1] = make_double2(x[2], v[0]); o[2] = make_double2(v[1], v[2]);
      o[3] = make_double2(Q[0], Q[1]); o[4] = make_double2(Q[2], Q[3]); o[5] = make_double2(Q[4], Q[5]);
      o[6] = make_double2(Q[6], Q[7]); rec[18 * tid + 14] = Q[8];
    }
#pragma unroll
    for (int c = 0; c < 9; c++) if (!AOS) sh_Q[c * RS + tid] = Q[c];
    grp_sync();

    // ---- geometry, shear/stretch strain, internal force ------------------------------
    T dx[3], dv[3], dx2[3];
    T nb_x[3], nb_v[3], nb_x2[3], nb_Q[9];   // AOS: the neighbour's record, fetched with 128-bit loads
    if (AOS) {
      const double2 *q = reinterpret_cast<const double2 *>(rec + 18 * t_next);
      const double2 a0 = q[0], a1 = q[1], a2 = q[2], a3 = q[3], a4 = q[4], a5 = q[5], a6 = q[6];
      nb_x[0] = a0.x; nb_x[1] = a0.y; nb_x[2] = a1.x; nb_v[0] = a1.y; nb_v[1] = a2.x; nb_v[2] = a2.y;
      nb_Q[0] = a3.x; nb_Q[1] = a3.y; nb_Q[2] = a4.x; nb_Q[3] = a4.y; nb_Q[4] = a5.x; nb_Q[5] = a5.y;
      nb_Q[6] = a6.x; nb_Q[7] = a6.y; nb_Q[8] = rec[18 * t_next + 14];
      const double2 b0 = *reinterpret_cast<const double2 *>(rec + 18 * t_next2);
      nb_x2[0] = b0.x; nb_x2[1] = b0.y; nb_x2[2] = rec[18 * t_next2 + 2];
    }
#pragma unroll
    for (int c = 0; c < 3; c++) {
      if (AOS) {
        dv[c] = nb_v[c] - v[c];
        dx[c] = nb_x[c] - x[c];
        dx2[c] = nb_x2[c] - nb_x[c];
        continue;
      }
      T vn = sh_v[c * RS + t_next];
      dv[c] = vn - v[c];
      if (EDGE) {
        // apply the pending increment of the last kinematic update (same v's: nothing changed them since);
        // a moving base is re-pinned, so its own velocity does not stretch element 0
        T dvk = (moving && c < 2) ? vn : dv[c];
        ed[c] = fma(hh_prev, dvk, ed[c]);
        dx[c] = ed[c];
        dx2[c] = fma(hh_prev, sh_v[c * RS + t_next2] - vn, sh_x[c * RS + t_next]);   // neighbour's edge, updated alike
      } else {
        T xn = sh_x[c * RS + t_next];
        dx[c] = xn - x[c];
        dx2[c] = sh_x[c * RS + t_next2] - xn;
      }
    }
    if (EDGE) hh_prev = T(0);
    if (!elem_ok) dx[2] = K.rest_len;   // keeps the pseudo-element's quantities finite
    if (!vor_ok) dx2[2] = K.rest_len;
    T l2 = dot3(dx, dx);
    T l2n = dot3(dx2, dx2);
    T il = rsqrt_nr(l2);
    T iln = rsqrt_nr(l2n);
    T lg = fma(l2, il, T(1e-14));                 // |dx| + 1e-14 (reference guard)
    T ilg = fma(T(-1e-14) * il, il, il);          // 1/(l + 1e-14) to first order in 1e-14/l
    T lgn = fma(l2n, iln, T(1e-14));              // length of element j+1, recomputed locally
    if (spl) {
      sh_L[tid] = lg;
      if (active && first) {   // the forcing's own bookkeeping, once per env
        int need = 0;
        for (int d = 0; d < 3; d++) {
          if (!(A.spline_mask >> d & 1)) continue;
          double *ch = sp + d * sCH;
          bool differ = ch[2 * sP] == 0.0;                      // initial_call_flag
          for (int i = 0; i < sP; i++) differ = differ || !(ch[sP + i] == ch[i]);   // not np.array_equal
          if (differ) {
            ch[2 * sP] = 1.0;
            for (int i = 0; i < sP; i++) {                      // filter_activation
              const double dd = ch[i] - ch[sP + i];
              const double sg = (double)((dd > 0.0) - (dd < 0.0));
              ch[sP + i] += sg * fmin(A.spline_rate, fabs(dd));
            }
            need |= 1 << d;
          }
        }
        sh_need[r] = need;
      }
    }
    T e = lg * K.inv_rest_len * gam;
    T inv_e = K.rest_len * ilg;
    T inv_e_s = elem_ok ? inv_e : T(0);           // the tip thread's pseudo-element carries no stress
    T edot = dot3(dx, dv) * (ilg * K.inv_rest_len);
    // sigma = e Q t - z = Q dx / l0 - z exactly (e t = dx / l0): no tangent needed here
    T Qdx[3], nst[3], sfl[3];
#pragma unroll
    for (int i = 0; i < 3; i++) Qdx[i] = Q[3 * i] * dx[0];          // sweeps share dx[k] / nst[k]
#pragma unroll
    for (int i = 0; i < 3; i++) Qdx[i] = fma(Q[3 * i + 1], dx[1], Qdx[i]);
#pragma unroll
    for (int i = 0; i < 3; i++) Qdx[i] = fma(Q[3 * i + 2], dx[2], Qdx[i]) * gam;   // strain against the element's own rest length
#pragma unroll
    for (int i = 0; i < 3; i++)
      nst[i] = (i == 2) ? fma(K.S_over_l[i], Qdx[i], -K.S[i]) : K.S_over_l[i] * Qdx[i];
    if constexpr (VARY) {
      // COOMM `ApplyMuscles` with the TransverseMuscle active (call sites /root/reference/gym_softrobot/envs/octopus/
      // build.py:329-333, build_muscle_octopus.py:165-177, arm_push_env.py:198-209; the package itself is not in the
      // reference tree: published model restated, oracle/rod_oracle.c:apply_tm_muscle).  Radial fibres on the centre
      // line: normalised length 1/sqrt(e), area rest_area / e, force along nu = sigma + e3 = Q dx / l0 (|nu| = e):
      //   n_m = a (-sigma_max) (A0 / e) h(1/sqrt(e)) nu / e,   h(l) = max(3.06 l^3 - 13.64 l^2 + 18.01 l - 6.44, 0).
      // It enters like the rod's own stress resultant (Delta_h(Q^T n) on the nodes, (Q t e) x n l0 on the elements; the
      // latter vanishes for n_m || nu), so e n_m joins nst[] ahead of the common 1/e.
      if (tm_gain != T(0)) {
        const T l = sqrt_(inv_e);
        T hw = fma(fma(fma(T(3.06), l, T(-13.64)), l, T(18.01)), l, T(-6.44));
        hw = hw > T(0) ? hw : T(0);
        const T Fm = tm_gain * inv_e * hw * K.inv_rest_len;
#pragma unroll
        for (int i = 0; i < 3; i++) nst[i] = fma(Fm, Qdx[i], nst[i]);
      }
    }
    auto publish_stress = [&]() {      // lab-frame stress resultant Q^T n / e of this element, for the neighbour to the right
#pragma unroll
      for (int i = 0; i < 3; i++) sfl[i] = Q[i] * nst[0];
#pragma unroll
      for (int i = 0; i < 3; i++) sfl[i] = fma(Q[3 + i], nst[1], sfl[i]);
#pragma unroll
      for (int i = 0; i < 3; i++) sfl[i] = fma(Q[6 + i], nst[2], sfl[i]);
#pragma unroll
      for (int i = 0; i < 3; i++) {
        sfl[i] *= inv_e_s;
        if (!AOS) sh_s[i * RS + tid] = sfl[i];
      }
    };
    if constexpr (!LMUS) publish_stress();   // (LMUS: the longitudinal muscles need the curvature first, see below)

    if (MULTI && active && first && has_head) {
      // FixedJoint2Rigid(head, -1, arm, 0) (utils/custom_elastica/joint.py:48-123 forces, :125-219 torques)
      T hx[3], hv[3], d2[3], Qh[9];
#pragma unroll
      for (int c = 0; c < 3; c++) { hx[c] = sh_x[c * RS + t_head]; hv[c] = sh_v[c * RS + t_head]; d2[c] = sh_Q[(3 + c) * RS + t_head]; }
#pragma unroll
      for (int c = 0; c < 9; c++) Qh[c] = sh_Q[c * RS + t_head];
      const T cs = A.joint_cs[arm][0], sn = A.joint_cs[arm][1];
      T dir[3] = {-(cs * d2[0] - sn * d2[1]), -(sn * d2[0] + cs * d2[1]), -d2[2]};   // -Rz(angle) d2
      T anchor[3] = {hx[0] + A.joint_radius * dir[0], hx[1] + A.joint_radius * dir[1], T(0) + A.joint_radius * dir[2]};
      T dd[3] = {x[0] - anchor[0], x[1] - anchor[1], x[2] - anchor[2]};
      T dist = sqrt_(dot3(dd, dd));
      T inv = (dist <= T(2.220446049250313e-12)) ? T(0) : rcp_nr(dist);   // (the guard discards rcp(0))
      T nh[3] = {dd[0] * inv, dd[1] * inv, dd[2] * inv};
      T rvn = (v[0] - hv[0]) * nh[0] + (v[1] - hv[1]) * nh[1] + (v[2] - hv[2]) * nh[2];
      T cf[3], tgt[3], fd[3], tau[3];
#pragma unroll
      for (int c = 0; c < 3; c++) {
        cf[c] = A.joint_k * dd[c] - A.joint_nu * (rvn * nh[c]);
        fj[c] = -cf[c];
        tgt[c] = anchor[c] + K.rest_len * dir[c];
        fd[c] = -A.joint_kt * ((x[c] + dx[c]) - tgt[c]);       // x + dx = node 1 of the arm
      }
      cross3(dx, fd, tau);                                        // link_direction x force
#pragma unroll
      for (int i = 0; i < 3; i++) {
        tj[i] = Q[3 * i] * tau[0] + Q[3 * i + 1] * tau[1] + Q[3 * i + 2] * tau[2];
        sh_J[i * RS + tid] = cf[i];
        sh_J[(3 + i) * RS + tid] = -(Qh[3 * i] * tau[0] + Qh[3 * i + 1] * tau[1] + Qh[3 * i + 2] * tau[2]);
      }
    }

    // ---- curvature, bending couple ----------------------------------------------------
    T vec[3], tr;
    {
      T Qn[9];
#pragma unroll
      for (int c = 0; c < 9; c++) Qn[c] = AOS ? nb_Q[c] : sh_Q[c * RS + t_next];
      auto rm = [&](int a, int b) {
        return fma(Qn[3 * a + 2], Q[3 * b + 2], fma(Qn[3 * a + 1], Q[3 * b + 1], Qn[3 * a] * Q[3 * b]));
      };
      // axial part of Rm - Rm^T, each component one 6-term FMA chain
      auto rm_diff = [&](int a, int b) {
        T p = Qn[3 * a] * Q[3 * b];
        p = fma(Qn[3 * a + 1], Q[3 * b + 1], p);
        p = fma(Qn[3 * a + 2], Q[3 * b + 2], p);
        p = fma(-Qn[3 * b], Q[3 * a], p);
        p = fma(-Qn[3 * b + 1], Q[3 * a + 1], p);
        return fma(-Qn[3 * b + 2], Q[3 * a + 2], p);
      };
      vec[0] = rm_diff(2, 1); vec[1] = rm_diff(0, 2); vec[2] = rm_diff(1, 0);
      tr = rm(0, 0) + rm(1, 1) + rm(2, 2);
    }
    T u = fma(T(-0.25), tr, T(0.75 + 0.5e-10));   // sin^2(theta_ref/2) with the 1e-10 guard
    if (!vor_ok) u = T(5e-11);
    constexpr bool F64 = sizeof(T) == 8;
    if (!F64) u = fmax(u, T(0));                  // FP32: the guard is below epsilon; keep u >= 0
    T fac;
    // actuated arms bend more per element than free / clamped rods (the contact models use u <= 0.1, 37 degrees); the
    // 10-element arms of the octopus assemblies (up to ~45 degrees per element) keep the full-range map.  One map per
    // instantiation, shared by its fast-only and safe variants; beyond its range the safe variant falls back, per
    // thread, to the reference's acos / sin.
    constexpr bool MID_BEND = CONTACT && !MULTI, WIDE_BEND = MULTI;
    const bool bend_out = !(u <= T(WIDE_BEND ? kSmallBendU : MID_BEND ? kMidBendU : kNarrowBendU));
    if (FASTONLY) dom_bad = dom_bad || bend_out;
    if (FASTONLY || !bend_out) {
      const T g = MID_BEND ? theta_over_sin_mid(A.poly, u) : !WIDE_BEND ? theta_over_sin_narrow(A.poly, u) : theta_over_sin(A.poly, u);
      if (F64) {
        // cot(theta) = (1 - 2u) / (2 sqrt(u (1 - u))); it only scales the 1e-14 guard term, so on the narrower ranges
        // its expansion rsqrt(4u) (1 - 1.5 u) (relative error < 0.63 u^2 <= 6e-3) is more than enough
        T cot = !WIDE_BEND ? rsqrt_approx(T(4.0) * u) * fma(T(-1.5), u, T(1.0))
                           : fma(T(-2.0), u, T(1.0)) * rsqrt_approx(T(4.0) * u * (T(1.0) - u));
        fac = g * fma(T(0.5e-14), cot, T(-0.5));
      } else {
        fac = T(-0.5) * g;   // the 1e-14 cot(theta) term is < 1e-9: invisible in FP32
      }
    } else {
      fac = bend_factor_ref<T>(u);
    }
    T kp[3], tau[3], kxt[3];
    T fs = fac * K.inv_rest_vor;
#pragma unroll
    for (int i = 0; i < 3; i++) { kp[i] = vec[i] * fs; tau[i] = K.B[i] * (CONTACT ? kp[i] - rk[i] : kp[i]); }
    cross3(kp, tau, kxt);
    if constexpr (LMUS) {
      // COOMM LongitudinalMuscle under ApplyMuscles (package outside the reference tree; published model restated in
      // oracle/shims/coomm/actuations/muscles/muscle.py and oracle/rod_oracle.c:apply_muscle_layers): a muscle at
      // material-frame offset x_m = (px, py, 0) r (r = current element radius) has
      //   nu_m = sigma + e3 + kappa_e x x_m + d x_m / ds,   kappa_e = trapezoid of kappa (zero ghosts),
      //   n_m = a sigma_max (A0 / e) h(|nu_m|) nu_m / |nu_m|,   couple x_m x n_m on the element.
      // Its force loads the rod exactly like the rod's own stress resultant (Delta_h(Q^T n_m) on the nodes,
      // (Q t e) x n_m l0 on the element), so e n_m joins nst[] ahead of the common 1 / e; the couple goes to the Voronoi
      // points (average of the two elements) and comes back like the bending couple after the second barrier.
      const T rad2 = K.vol_over_pi * ilg, rad = rad2 * rsqrt_nr(rad2);   // sqrt(V / (pi l))
#pragma unroll
      for (int i = 0; i < 3; i++) { lm_kr[i] = vor_ok ? kp[i] : T(0); sh_M[i * RS + tid] = lm_kr[i]; }
      sh_M[3 * RS + tid] = rad;
      grp_sync();
      T ke[3], nu[3];
#pragma unroll
      for (int i = 0; i < 3; i++) {
        lm_kl[i] = sh_M[i * RS + t_prev];
        ke[i] = T(0.5) * (lm_kl[i] + lm_kr[i]);
        nu[i] = K.inv_rest_len * Qdx[i];
        lm_cpl[i] = T(0);
      }
      const T rad_a = (elem_ok && j > 0) ? sh_M[3 * RS + tid - 1] : rad;
      const T rad_b = (elem_ok && j < n - 1) ? sh_M[3 * RS + tid + 1] : rad;
      const T drad = (rad_b - rad_a) * lm_inv_ds;
#pragma unroll
      for (int m = 0; m < 2; m++) {
        const T g = m ? lm_g1 : lm_g0;
        if (g != T(0)) {
          const T px = A.lm_px[m], py = A.lm_py[m], xm0 = px * rad, xm1 = py * rad;
          T sm[3];
          sm[0] = nu[0] - ke[2] * xm1 + px * drad;
          sm[1] = nu[1] + ke[2] * xm0 + py * drad;
          sm[2] = nu[2] + (ke[0] * xm1 - ke[1] * xm0);
          const T l2m = dot3(sm, sm), ilm = rsqrt_nr(l2m), len = l2m * ilm;   // |nu_m| and its reciprocal (Newton-refined, ~1 ulp)
          T hw = fma(fma(fma(T(3.06), len, T(-13.64)), len, T(18.01)), len, T(-6.44));
          hw = hw > T(0) ? hw : T(0);
          const T Fe = g * hw * ilm;                    // e x |n_m| / |nu_m|
#pragma unroll
          for (int i = 0; i < 3; i++) nst[i] = fma(Fe, sm[i], nst[i]);
          const T Fn = Fe * inv_e;                      // n_m = Fn nu_m
          lm_cpl[0] = fma(xm1, Fn * sm[2], lm_cpl[0]);
          lm_cpl[1] = fma(-xm0, Fn * sm[2], lm_cpl[1]);
          lm_cpl[2] += Fn * (xm0 * sm[1] - xm1 * sm[0]);
        }
      }
#pragma unroll
      for (int i = 0; i < 3; i++) sh_M[(4 + i) * RS + tid] = lm_cpl[i];
      publish_stress();
    }
    T eps = (T(0.5) * (lgn + lg)) * K.inv_rest_vor;
    T ie3 = rcp_nr(eps * eps * eps);
    if (!vor_ok) ie3 = T(0);
    T hc = T(0.5) * K.rest_vor * ie3;
    // local couples share one 1/e factor:  (Qt x n) l0 + (Jw/e) x w + (Jw/e) (de/dt)/e
    //   = [ (Q dx) x n + (Jw) x w + (Jw) (de/dt)/e ] / e      (Qt l0 = Q dx / e)
    // Everything that does not need the left neighbour is folded into tql[] before the second
    // barrier, so only x,v,Q,w + tql + sfl + e stay live across it.
    T pw[3] = {K.J[0] * w[0], K.J[1] * w[1], K.J[2] * w[2]};
    T ede = edot * inv_e;
    T Gc[3];
    Gc[0] = fma(Qdx[1], nst[2], -(Qdx[2] * nst[1]));
    Gc[1] = fma(Qdx[2], nst[0], -(Qdx[0] * nst[2]));
    Gc[2] = fma(Qdx[0], nst[1], -(Qdx[1] * nst[0]));
    Gc[0] = fma(pw[1], w[2], fma(-pw[2], w[1], Gc[0]));
    Gc[1] = fma(pw[2], w[0], fma(-pw[0], w[2], Gc[1]));
    Gc[2] = fma(pw[0], w[1], fma(-pw[1], w[0], Gc[2]));
    T tql[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
      T m = tau[i] * ie3;
      T Pi = fma(kxt[i], hc, m);                 // m_j + c_j/2  (own element)
      const T Nout = fma(kxt[i], hc, -m);        // c_j/2 - m_j  (element j+1)
      if (AOS) nb_x2[i] = Nout;                  // (reused as staging for the 128-bit store below)
      else sh_N[i * RS + tid] = Nout;
      tql[i] = fma(fma(pw[i], ede, Gc[i]), inv_e, Pi);
    }
    if (AOS) {   // {s0 s1 | s2 N0 | N1 N2}
      double2 *o = reinterpret_cast<double2 *>(sn + 6 * tid);
      o[0] = make_double2(sfl[0], sfl[1]); o[1] = make_double2(sfl[2], nb_x2[0]); o[2] = make_double2(nb_x2[1], nb_x2[2]);
    }
    // rotational damper coefficients c_w^e = c_w exp((e-1) ln c_w): only need e, so they are
    // evaluated here, ahead of the barrier, off the critical path of the dynamic step
    T cw0 = T(1), cw1 = T(1), cw2 = T(1);
    if (FASTONLY || A.damping_on) {   // damper off = identity constants (c_w = 1, ln c_w = 0): no branch needed
      T em1 = e - T(1);
      T z0 = em1 * K.logc_w[0], z1 = em1 * K.logc_w[1], z2 = em1 * K.logc_w[2];
      // the contact / filtered models damp harder and stretch more (z up to ~1e-3): they use the wide exp map
      constexpr bool NARROW_EXP = !CONTACT && !LAPLACE && !LMUS;   // (muscles stretch the arms by tens of per cent)
      constexpr double kz = NARROW_EXP ? kNarrowExpZ : kSmallExpZ;
      const bool exp_out = !(fabs_(z0) <= T(kz)) || !(fabs_(z1) <= T(kz)) || !(fabs_(z2) <= T(kz));
      if (FASTONLY) dom_bad = dom_bad || exp_out;
      if (FASTONLY || !exp_out) {
        cw0 = K.c_w[0] * (NARROW_EXP ? exp_narrow(A.poly, z0) : exp_small(A.poly, z0));
        cw2 = K.c_w[2] * (NARROW_EXP ? exp_narrow(A.poly, z2) : exp_small(A.poly, z2));
        cw1 = K.isotropic ? cw0 : K.c_w[1] * (NARROW_EXP ? exp_narrow(A.poly, z1) : exp_small(A.poly, z1));
      } else {
        cw0 = exp_ref<T>(e * K.logc_w[0]);
        cw1 = exp_ref<T>(e * K.logc_w[1]);
        cw2 = exp_ref<T>(e * K.logc_w[2]);
      }
    }
    if (last) {
      // stale observables of the reference: last force evaluation
      if (active) {
#pragma unroll
        T tg_out[3] = {dx[0] * ilg, dx[1] * ilg, dx[2] * ilg};
        to_lab(tg_out);
#pragma unroll
        for (int i = 0; i < 3; i++) {
          st[(F_TAN + i) * stride + j] = tg_out[i];
          st[(F_KAPPA + i) * stride + j] = kp[i];
          st[(F_SIGMA + i) * stride + j] = fma(K.inv_rest_len, Qdx[i], (i == 2) ? T(-1) : T(0));
        }
        st[F_DIL * stride + j] = e;
      }
    }
    grp_sync();

    // ---- add the left neighbour's share, [contact], dynamic step ---------------------------
    T dtee = dte * e;
    T fint[3], tq[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
      if (!AOS) {
        fint[i] = sfl[i] - sh_s[i * RS + t_prev];
        tq[i] = tql[i] + sh_N[i * RS + t_prev];
      }
    }
    if (AOS) {
      const double2 *q = reinterpret_cast<const double2 *>(sn + 6 * t_prev);
      const double2 a0 = q[0], a1 = q[1], a2 = q[2];
      fint[0] = sfl[0] - a0.x; fint[1] = sfl[1] - a0.y; fint[2] = sfl[2] - a1.x;
      tq[0] = tql[0] + a1.y; tq[1] = tql[1] + a2.x; tq[2] = tql[2] + a2.y;
    }
    if constexpr (LMUS) {
      // muscle couple on the Voronoi points m = (c_k + c_k+1) / 2; element k gets Delta_h(m) + A_h(kappa x m D)
      if (elem_ok) {
        T mr[3], ml[3], cr[3], cl[3];
#pragma unroll
        for (int i = 0; i < 3; i++) {
          mr[i] = vor_ok ? T(0.5) * (sh_M[(4 + i) * RS + tid + 1] + lm_cpl[i]) : T(0);
          ml[i] = (j > 0) ? T(0.5) * (lm_cpl[i] + sh_M[(4 + i) * RS + tid - 1]) : T(0);
        }
        cross3(lm_kr, mr, cr);
        cross3(lm_kl, ml, cl);
#pragma unroll
        for (int i = 0; i < 3; i++)
          tq[i] += (mr[i] - ml[i]) + T(0.5) * (cr[i] * K.rest_vor + cl[i] * lm_rvl);
      }
    }
    // generic per-element external loads (a forcing: what COOMM's ApplyMuscles would feed,
    // /root/reference/gym_softrobot/envs/octopus/build_muscle_octopus.py:171-176): nodal forces in the lab frame,
    // element couples in the material frame, constant during a launch, added every substep ahead of the contact
    if (A.ext_force && active) {
      T fe[3];
#pragma unroll
      for (int i = 0; i < 3; i++) fe[i] = A.ext_force[((size_t)rod * 3 + i) * stride + j];
      to_int(fe);
#pragma unroll
      for (int i = 0; i < 3; i++) fint[i] += fe[i];
    }
    if (A.ext_couple && elem_ok) {
#pragma unroll
      for (int i = 0; i < 3; i++) tq[i] += A.ext_couple[((size_t)rod * 3 + i) * stride + j];
    }
    if (spl) {
      const int need = live ? sh_need[r] : 0;
      if (need && elem_ok) {
        double s_k = 0.0;                                        // np.cumsum(system.lengths)[j]
        for (int i = 0; i <= j; i++) s_k += (double)sh_L[tid - j + i];
        const int m = min(max((int)floor(s_k * A.spline_inv_dx), 0), sP);
        const double t = s_k - (double)m / A.spline_inv_dx;
        for (int d = 0; d < 3; d++) {
          if (!(need >> d & 1)) continue;
          const double *ch = sp + d * sCH, *tb = A.spline_tab + (size_t)m * sP * 4;
          double val = 0.0;
          for (int i = 0; i < sP; i++)
            val += ch[sP + i] * (tb[4 * i] + t * (tb[4 * i + 1] + t * (tb[4 * i + 2] + t * tb[4 * i + 3])));
          val *= A.spline_scale;
          sp[3 * sCH + d * n + j] = val;
          smag[d] = (T)val;
        }
      }
#pragma unroll
      for (int i = 0; i < 3; i++) tq[i] += smag[i];
    }
    // forcing registered before the contact: the static-friction torque balance sees the muscle couple
    if (TORQUE && A.muscle_on && !A.contact_before_forcing) {
#pragma unroll
      for (int i = 0; i < 3; i++) tq[i] += mtq[i];
    }
    if (CONTACT && A.contact_on) {
      // RodPlaneContactWithAnisotropicFriction (elastica/_contact_functions.py), per element j
      // between nodes j and j+1, in the frame whose z axis is the plane normal.  Stage 1: normal response + kinetic
      // friction from the nodal forces accumulated so far; stage 2: static friction from the forces INCLUDING
      // stage 1 of both neighbours (nodal forces mix adjacent elements), hence two more exchanges.
      // stage-1 loads travel through the (now free) Q rows; the final loads through the s rows, which
      // nobody overwrites before the next strain phase (the Q rows are re-published right after this step)
      T *sh_c1 = sh_Q, *sh_c12 = sh_s;
      const bool has_left = active && j > 0, has_right = active && j + 1 < n;
      const bool end0 = (j == 0), end1 = (j + 1 == n);
      // mass weights of the element velocity (m_k v_k + m_k+1 v_k+1) / (m_k + m_k+1): end nodes carry half a mass
      const T w0 = VARY ? mass_j / (mass_j + mass_j1) : (end0 == end1) ? T(0.5) : end0 ? T(1.0 / 3.0) : T(2.0 / 3.0), w1 = T(1) - w0;
      const T hm0 = end0 ? T(0.5) : T(1), hm1 = end1 ? T(0.5) : T(1);
      T etf[3], evel[3], t[3];
#pragma unroll
      for (int i = 0; i < 3; i++) {
        T f0 = fint[i], f1 = sh_s[i * RS + t_next] - sfl[i];   // internal force on nodes j, j+1
        if (!A.contact_before_forcing) {                       // external loads so far: gravity, joints (+ base force)
          if (VARY) { f0 = fma(A.g[i], mass_j, f0); f1 = fma(A.g[i], mass_j1, f1); }
          else { f0 = fma(A.gm[i], hm0, f0); f1 = fma(A.gm[i], hm1, f1); }
          if (i == 0 && A.point_force && first) f0 = fint[i] + act0;
          if (MULTI) f0 += fj[i];                              // joint force on node 0 (zero elsewhere)
        }
        // (an external nodal force is a forcing as well; node j's share is already in fint, node j+1's is added below)
        etf[i] = T(0.5) * (f0 + f1) + (end0 ? T(0.5) * f0 : T(0)) + (end1 ? T(0.5) * f1 : T(0));
        evel[i] = fma(w1, sh_v[i * RS + t_next], w0 * v[i]);
        t[i] = dx[i] * ilg;
      }
      if (A.ext_force && elem_ok && !A.contact_before_forcing) {
        T fe[3];
#pragma unroll
        for (int i = 0; i < 3; i++) fe[i] = A.ext_force[((size_t)rod * 3 + i) * stride + j + 1];
        to_int(fe);
#pragma unroll
        for (int i = 0; i < 3; i++) etf[i] += (end1 ? T(1.0) : T(0.5)) * fe[i];
      }
      const T rad2 = K.vol_over_pi * ilg, inv_rad = rsqrt_nr(rad2), rad = rad2 * inv_rad;   // sqrt(V / (pi l))
      const T fn = etf[2], vn = evel[2];
      const T gap = (fma(T(0.5), dx[2], x[2]) - A.plane_z0) - rad, pen = fmin(gap, T(0));
      const bool nocontact = !elem_ok || (gap > A.surface_tol);
      const T resp_mag = (nocontact || fn > T(0)) ? T(0) : fabs_(fn);
      T c1[3];
      c1[2] = nocontact ? T(0) : ((fn > T(0)) ? T(0) : -fn) - A.contact_k * pen - A.contact_nu * vn;
      // axial direction = the tangent's projection on the plane, normalised with the reference's guard
      // 1 / (|tp| + 1e-14) (to first order in 1e-14 / |tp|); rolling direction = axial x normal
      const T tp2 = fma(t[1], t[1], t[0] * t[0]);
      const T rtp = rsqrt_nr(tp2 + T(sizeof(T) == 8 ? 1e-300 : 1e-30));   // (a rod standing on end: tp = 0, axial direction 0)
      const T inv_tp = fma(T(-1e-14) * rtp, rtp, rtp);
      const T ax0 = t[0] * inv_tp, ax1 = t[1] * inv_tp, rl0 = ax1, rl1 = -ax0;
      auto slip_fn = [&](T a) {   // find_slipping_elements on |v| (|axial| = |rolling| = 1 - 1e-14 / |tp|: taken as 1)
        return (a > A.slip_tol) ? fabs_(T(1) - fmin(T(1), a * A.inv_slip_tol - T(1))) : T(1);
      };
      auto sgn = [](T a) { return T((a > T(0)) - (a < T(0))); };
      const T vax = fma(evel[1], ax1, evel[0] * ax0), sg = sgn(vax);
      const T kmu = T(0.5) * (A.kin_mu[0] * (T(1) + sg) + A.kin_mu[1] * (T(1) - sg));
      const T slipa = slip_fn(fabs_(vax));
      // velocity of the contact point relative to the axis: Q^T (w x Q arm), arm = -rad z
      T qa[3], wq[3];
#pragma unroll
      for (int i = 0; i < 3; i++) qa[i] = -(Q[3 * i + 2] * rad);
      cross3(w, qa, wq);
      const T rv0 = fma(Q[6], wq[2], fma(Q[3], wq[1], Q[0] * wq[0])), rv1 = fma(Q[7], wq[2], fma(Q[4], wq[1], Q[1] * wq[0]));
      const T smag = fma(evel[1] + rv1, rl1, (evel[0] + rv0) * rl0);
      const T slipr = slip_fn(fabs_(smag));
      const T ut0 = fma(smag, rl0, vax * ax0), ut1 = fma(smag, rl1, vax * ax1);
      const T ug0 = ut0 + T(1e-14), ug1 = ut1 + T(1e-14);
      const T iun = rsqrt_nr(fma(ug1, ug1, fma(ug0, ug0, T(1e-28))));
      const T uax = fma(ut1, ax1, ut0 * ax0) * iun, url = fma(ut1, rl1, ut0 * rl0) * iun;
      const T ka = nocontact ? T(0) : -((T(1) - slipa) * kmu * resp_mag * uax);
      const T kr = nocontact ? T(0) : -((T(1) - slipr) * A.kin_mu[2] * resp_mag * url);
      T fr0 = kr * rl0, fr1 = kr * rl1;
      c1[0] = fma(ka, ax0, fr0); c1[1] = fma(ka, ax1, fr1);
      // couple of the rolling friction: Q (arm x F) = Q (rad F_y, -rad F_x, 0)
      T cr0 = rad * fr1, cr1 = -(rad * fr0);
      T text[3];
#pragma unroll
      for (int i = 0; i < 3; i++) {
        text[i] = fma(Q[3 * i + 1], cr1, Q[3 * i] * cr0);
        sh_c1[i * RS + tid] = c1[i];
      }
      grp_sync();
      // stage 2: static friction (in-plane components only)
      T e2[2];
#pragma unroll
      for (int i = 0; i < 2; i++) {
        const T cl = has_left ? sh_c1[i * RS + tid - 1] : T(0), crr = has_right ? sh_c1[i * RS + tid + 1] : T(0);
        const T nc0 = T(0.5) * (cl + c1[i]), nc1 = T(0.5) * (c1[i] + crr);   // stage-1 response on nodes j, j+1
        e2[i] = etf[i] + T(0.5) * (nc0 + nc1) + (end0 ? T(0.5) * nc0 : T(0)) + (end1 ? T(0.5) * nc1 : T(0));
      }
      const T fax = fma(e2[1], ax1, e2[0] * ax0), sga = sgn(fax);
      const T smu = T(0.5) * (A.stat_mu[0] * (T(1) + sga) + A.stat_mu[1] * (T(1) - sga));
      const T sa = nocontact ? T(0) : -(fmin(fabs_(fax), slipa * smu * resp_mag) * sga);
      T tsum[3] = {tq[0] + text[0], tq[1] + text[1], tq[2] + text[2]};
      if (MULTI && !A.contact_before_forcing) {   // the joint's couple on element 0 is already registered
#pragma unroll
        for (int i = 0; i < 3; i++) tsum[i] += tj[i];
      }
      const T tt0 = fma(Q[6], tsum[2], fma(Q[3], tsum[1], Q[0] * tsum[0])), tt1 = fma(Q[7], tsum[2], fma(Q[4], tsum[1], Q[1] * tsum[0]));
      const T noslip = -((rad * fma(e2[1], rl1, e2[0] * rl0) - T(2) * fma(tt1, ax1, tt0 * ax0)) * (T(1.0 / 3.0) * inv_rad));
      const T sr_ = nocontact ? T(0) : fmin(fabs_(noslip), slipr * A.stat_mu[2] * resp_mag) * sgn(noslip);
      fr0 = sr_ * rl0; fr1 = sr_ * rl1;
      sh_c12[0 * RS + tid] = c1[0] + fma(sa, ax0, fr0);
      sh_c12[1 * RS + tid] = c1[1] + fma(sa, ax1, fr1);
      sh_c12[2 * RS + tid] = c1[2];
      cr0 = rad * fr1; cr1 = -(rad * fr0);
#pragma unroll
      for (int i = 0; i < 3; i++) tq[i] += text[i] + fma(Q[3 * i + 1], cr1, Q[3 * i] * cr0);
      grp_sync();
#pragma unroll
      for (int i = 0; i < 3; i++)   // node j collects half of the plane's load on elements j-1 and j
        fint[i] += T(0.5) * (sh_c12[i * RS + tid] + (has_left ? sh_c12[i * RS + tid - 1] : T(0)));
    }
    if (TORQUE && A.muscle_on && A.contact_before_forcing) {
#pragma unroll
      for (int i = 0; i < 3; i++) tq[i] += mtq[i];
    }
    if (MULTI && is_head) {
      // rigid head: a = F/m, alpha = J^-1 ((J w) x w + T)  (SURVEY D.1); loads = the joints' reactions,
      // summed in connection order; no gravity, no damper (build.py:141-152); then the rate part of
      // BodyBoundaryCondition: v_z = 0, w_x = w_y = 0
      if (hd) {
        T F[3] = {T(0), T(0), T(0)}, Tq[3] = {T(0), T(0), T(0)};
        for (int a = 0; a < n_rod; a++) {
          const int ta = r * G + a * tpr;
#pragma unroll
          for (int i = 0; i < 3; i++) { F[i] += sh_J[i * RS + ta]; Tq[i] += sh_J[(3 + i) * RS + ta]; }
        }
        T Jw[3] = {A.head_J[0] * w[0], A.head_J[1] * w[1], A.head_J[2] * w[2]}, lt[3];
        cross3(Jw, w, lt);
#pragma unroll
        for (int i = 0; i < 3; i++) {
          v[i] = fma(F[i], A.head_dt_inv_mass, v[i]);
          w[i] = fma(dt, A.head_Jinv[i] * (lt[i] + Tq[i]), w[i]);
        }
        v[2] = T(0); w[0] = T(0); w[1] = T(0);
        if (LMUS && A.head_fixed) {   // OneEndFixedBC.constrain_rates, registered after the BodyBoundaryCondition
#pragma unroll
          for (int i = 0; i < 3; i++) { v[i] = T(0); w[i] = T(0); }
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 3; i++) {
        T fi = fint[i];
        if (MULTI) fi += fj[i];
        T gd = A.gdt_cv[i];
        if (i == 0 && A.point_force && first) { fi += act0; gd = T(0); }
        // v <- c_v (v + dt f/m + dt g): the translational damper folded into the update
        v[i] = fma(fi, dtim_cv, fma(v[i], A.c_v, gmask * gd));
        T ti = tq[i];
        if (MULTI) ti += tj[i];
        w[i] = fma(dtee, K.Jinv[i] * ti, w[i]);
      }
    }

    // ---- rate constraints and dissipation ------------------------------------------------
    auto dampen = [&]() {
      if (FASTONLY || A.damping_on) { w[0] *= cw0; w[1] *= cw1; w[2] *= cw2; }
      if (LAPLACE && A.laplace_order > 0) {
        // LaplaceDissipationFilter (elastica/dissipation.py:nb_filter_rate): p passes of
        // f <- (-f[k+1] - f[k-1] + 2 f[k]) / 4 on interior nodes / elements, ends held at 0; rate -= f.
        // Two smem buffers alternate (x,v rows / first six Q rows), one barrier per pass.
        const bool node_in = active && j > 0 && j < n, elem_in = active && j > 0 && j < n - 1;
        T fv[3] = {node_in ? v[0] : T(0), node_in ? v[1] : T(0), node_in ? v[2] : T(0)};
        T fw[3] = {elem_in ? w[0] : T(0), elem_in ? w[1] : T(0), elem_in ? w[2] : T(0)};
        // pass 0 sees the *unfiltered* ends (w[0], w[-1] are only zeroed after the first update)
        T ev[3] = {v[0], v[1], v[2]}, ew[3] = {w[0], w[1], w[2]};
        // end nodes / elements (and idle threads) take themselves as both neighbours: (-f - f + 2 f) / 4 = 0 exactly,
        // which is the "ends held at 0" of the reference without a select per component and pass
        const int tv_next = node_in ? tid + 1 : tid, tv_left = node_in ? tid - 1 : tid;
        const int tw_next = elem_in ? tid + 1 : tid, tw_left = elem_in ? tid - 1 : tid;
        auto pass = [&](T *buf, bool first_pass) {
#pragma unroll
          for (int c = 0; c < 3; c++) {
            buf[c * RS + tid] = first_pass ? ev[c] : fv[c];
            buf[(3 + c) * RS + tid] = first_pass ? ew[c] : fw[c];
          }
          grp_sync();
#pragma unroll
          for (int c = 0; c < 3; c++) {
            T mv = first_pass ? ev[c] : fv[c], mw = first_pass ? ew[c] : fw[c];
            fv[c] = (-buf[c * RS + tv_next] - buf[c * RS + tv_left] + T(2) * mv) * T(0.25);
            fw[c] = (-buf[(3 + c) * RS + tw_next] - buf[(3 + c) * RS + tw_left] + T(2) * mw) * T(0.25);
          }
        };
        if (A.laplace_order == 7) {   // SoftPendulum3D-v0's order: unrolled, buffers and the first-pass case resolved at compile time
#pragma unroll
          for (int p = 0; p < 7; p++) pass((p & 1) ? sh_Q : sh_x, p == 0);
        } else {
          for (int p = 0; p < A.laplace_order; p++) pass((p & 1) ? sh_Q : sh_x, p == 0);
        }
        grp_sync();   // the last pass's reads finish before x,v,Q are published again
#pragma unroll
        for (int c = 0; c < 3; c++) { v[c] -= fv[c]; w[c] -= fw[c]; }
      }
    };
    if (!(MULTI && is_head)) {
      // BCs that only zero components commute with the (multiplicative / interior-only) dampers: no order branch
      if ((FASTONLY && !MOVING) || A.damp_first) { dampen(); constrain_rates(); }
      else { constrain_rates(); dampen(); }
      // ControllableFixConstraint ("sucker", envs/octopus/controllable_constraint.py:46-69): the rates of one node /
      // element index are scaled by 1 - reduction_ratio (per rod, 0 = released); registered after the dampers
      if constexpr (LMUS) {
        // several ControllableFixConstraints at fixed, distinct indices (arm_two_env.py:133-145): node and element
        // `index` of the rod scaled by 1 - reduction_ratio of that slot
        if (A.msucker && active) {
#pragma unroll
          for (int sl = 0; sl < 3; sl++) {
            if (sl < A.msucker_n && j == A.msucker_loc[sl]) {
              const T f = T(1) - A.msucker[(size_t)rod * 3 + sl];
#pragma unroll
              for (int i = 0; i < 3; i++) { v[i] *= f; w[i] *= f; }
            }
          }
        }
      }
      if (A.sucker && active) {
        const T f = T(1) - A.sucker[rod];
        if (j == sk_node) {
#pragma unroll
          for (int i = 0; i < 3; i++) v[i] *= f;
        }
        if (j == sk_elem) {
#pragma unroll
          for (int i = 0; i < 3; i++) w[i] *= f;
        }
      }
    }

    kinematic(last ? h : dt, last ? T(1e-14) : T(2e-14));
    head_constrain_values();
  }

  // ---- write back, NaN guard, model outputs ------------------------------------------------
  if (EDGE && A.n_substeps > 0) {   // the last kinematic update's edge increment needs the final velocities
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 3; c++) sh_v[c * RS + tid] = v[c];
    __syncthreads();
    if (active) {
#pragma unroll
      for (int c = 0; c < 3; c++) {
        T vn = sh_v[c * RS + t_next];
        ed[c] = fma(hh_prev, (moving && c < 2) ? vn : vn - v[c], ed[c]);   // (stored below, once the env is known to be final)
      }
    }
  }
  __syncthreads();   // all reads of the exchange buffers are done: reuse them below
  bool redo = false;
  if (FASTONLY) {    // env-level OR of the domain flags; a flagged env keeps its pre-launch state in global memory
    int *sh_dom = reinterpret_cast<int *>(sh_N);
    if (tid < rods_per_cta) sh_dom[tid] = 0;   // (rods_per_cta <= NT / 4: one slot per env group of this CTA)
    __syncthreads();
    if (live && dom_bad) atomicOr(&sh_dom[r], 1);
    __syncthreads();
    redo = live && sh_dom[r] != 0;
    if (redo && active && first && arm == 0) {
      A.redo[env] = 1;
      if (A.redo_count) atomicAdd(A.redo_count, 1ULL);
    }
  }
  if (EDGE && A.n_substeps > 0 && active && !redo) {
    to_lab(ed);
#pragma unroll
    for (int c = 0; c < 3; c++) st[(F_EDGE + c) * stride + j] = elem_ok ? ed[c] : T(0);
  }
  bool bad = false;
  if (MULTI && hd && !redo) {
#pragma unroll
    for (int c = 0; c < 3; c++) { hd[c] = x[c]; hd[3 + c] = v[c]; hd[15 + c] = w[c]; }
#pragma unroll
    for (int c = 0; c < 9; c++) hd[6 + c] = Q[c];
    float *o = A.obs + (size_t)env * A.obs_dim;
    for (int c = 0; c < 3; c++) { o[c] = (float)x[c]; o[3 + c] = (float)v[c]; }
  }
  if (active && !redo) {
    to_lab(x); to_lab(v); rows_to_lab(Q);     // (x, v, Q are final: nothing below reads them in the internal frame)
#pragma unroll
    for (int c = 0; c < 3; c++) {
      st[(F_POS + c) * stride + j] = x[c];
      st[(F_VEL + c) * stride + j] = v[c];
      bad = bad || (x[c] != x[c]) || (v[c] != v[c]);
    }
    if (j < n) {
#pragma unroll
      for (int c = 0; c < 3; c++) st[(F_OMEGA + c) * stride + j] = w[c];
#pragma unroll
      for (int c = 0; c < 9; c++) st[(F_DIR + c) * stride + j] = Q[c];
    }
  }
  // per-rod NaN flag and tangents for the observation (rod r occupies tids r*tpr .. r*tpr+n)
  int *sh_flag = reinterpret_cast<int *>(sh_s);
  if (tid < rods_per_cta) sh_flag[tid] = 0;
  if (active && (A.model == MODEL_SOFT_PENDULUM || A.model == MODEL_SOFT_PENDULUM_3D)) {
    // the tangents were stored to global memory at the last substep; stage them for lane j=0
#pragma unroll
    for (int i = 0; i < 3; i++) sh_x[i * RS + tid] = (j < n) ? st[(F_TAN + i) * stride + j] : T(0);
  }
  __syncthreads();
  if (active && bad) atomicOr(&sh_flag[r], 1);
  __syncthreads();
  if (!FASTONLY && A.redo_filter && active && first && arm == 0) A.redo[env] = 0;
  if (active && first && arm == 0 && !redo) {
    const bool invalid = sh_flag[r] != 0;
    if (TORQUE && A.muscle_on) A.muscle[(size_t)env * A.muscle_dim] = mus_t;
    if (A.model == MODEL_SOFT_PENDULUM) {
      soft_pendulum_outputs<T>(sh_x + tid, RS, n, (double)x[0], (double)v[0], (float)act0, invalid,
                               A.obs + (size_t)env * A.obs_dim, A.reward + env, A.terminated + env);
    } else if (MOVING && A.model == MODEL_SOFT_PENDULUM_3D) {
      const double x0[3] = {(double)x[0], (double)x[1], (double)x[2]};
      const double v0[3] = {(double)v[0], (double)v[1], (double)v[2]};
      T *aux = A.aux + (size_t)env * AUX_DIM;
      if (FASTONLY && A.n_substeps > 0) {   // the base controller's state, deferred from the prologue
        aux[0] = pin_x; aux[1] = pin_y; aux[3] = base_vx; aux[4] = base_vy; aux[5] = T(0);
      }
      soft_pendulum_3d_outputs<T>(sh_x + tid, RS, n, x0, v0, act_f0, act_f1, (double)aux[0], (double)aux[1],
                                  invalid, A.obs + (size_t)env * A.obs_dim, A.reward + env,
                                  A.terminated + env, aux + 6);
    } else {
      A.reward[env] = 0.0;
      A.terminated[env] = invalid ? 1 : 0;
    }
  }
  if (active && A.model == MODEL_ROD && j == n && !MULTI && !redo) {
    float *o = A.obs + (size_t)env * A.obs_dim;
    for (int c = 0; c < 3; c++) { o[c] = (float)x[c]; o[3 + c] = (float)v[c]; }
  }
}

}  // namespace sr
