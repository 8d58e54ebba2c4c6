/* TEST INFRASTRUCTURE (oracle) — not product code.
 *
 * Plain-C, FP64, single-rod CPU restatement of the physics step that
 * gym-softrobot runs through PyElastica:
 *   boundary   /root/reference/gym_softrobot/envs/soft_pendulum/soft_pendulum.py:183-184
 *              (`time = PositionVerlet().step(simulator, time, dt)`)
 *   plugins    /root/reference/gym_softrobot/envs/soft_pendulum/build.py:29-115
 *              /root/reference/gym_softrobot/envs/soft_pendulum_3d/build.py:23-86
 *   algorithm  pyelastica==1.0.0 (third-party, pinned in /root/reference/uv.lock:845-857,
 *              NOT installable here) restated from SURVEY.md Appendix A — the
 *              same restatement as oracle/shims/elastica (NumPy), operation
 *              for operation.  PARITY UNPINNED against real PyElastica.
 *
 * Compile with -ffp-contract=off (no FMA contraction: Numba/NumPy do not fuse).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
 * arm may link or call this file.
 */
#include "rod_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define PI 3.141592653589793

struct ro_rod {
  ro_config cfg;
  int n;
  double time;
  /* state */
  double *x, *v, *Q, *w;                 /* (3,n+1) (3,n+1) (3,3,n) (3,n) */
  double *acc, *alpha;
  /* constants (A.1) */
  double *rest_len, *rest_vor, *mass, *volume, *radius;
  double *J, *Jinv, *S, *B;              /* diagonals: (3,n) (3,n) (3,n) (3,n-1) */
  double *rest_sigma, *rest_kappa;
  /* derived (A.3) — refreshed only inside the force evaluation (A.6) */
  double *len, *tang, *dil, *vdil, *dil_rate, *sigma, *kappa;
  double *stress, *couple, *f_int, *t_int, *f_ext, *t_ext, *f_user, *t_user;
  int n_sucker, sucker_idx[8], sucker_node[8]; /* ControllableFixConstraint: element / node indices and ratios */
  double sucker_ratio[8];
  /* COOMM TransverseMuscle under ApplyMuscles (restated, see apply_tm_muscle) */
  int tm_on;
  double tm_max_stress, tm_radius_ref, tm_activation;
  double *rest_radius;
  /* general COOMM muscle layers (apply_muscle_layers): up to 3 per rod, per-element activations */
  int ml_kind[3];                         /* 0 off, 1 longitudinal, 2 transverse */
  double ml_stress[3], ml_rref[3], ml_px[3], ml_py[3];
  double *ml_act;                         /* (3,n) */
  double *ml_tmp;                         /* scratch 16 n */
  /* plugins */
  double fixed_pos[3], fixed_Q[9];
  double c_v, *c_w;                      /* AnalyticalLinearDamper coefficients */
  double *filt;                          /* Laplace filter scratch (3,n+1) */
  double *tmp;                           /* scratch (3,n+1) x 4 */
  double *ctmp;                          /* contact scratch, 26 n */
  double *muscle;                        /* MuscleTorques: wave number, beta (1 + n) */
  double *spl_pts, *spl_mag;             /* spline forcing: [3][2P+1] control data, [3][n] cached magnitudes */
};

static double *zalloc(size_t n) { return (double *)calloc(n ? n : 1, sizeof(double)); }

#define X(i, k) r->x[(i) * (n + 1) + (k)]
#define V(i, k) r->v[(i) * (n + 1) + (k)]
#define QQ(i, j, k) r->Q[((i) * 3 + (j)) * n + (k)]
#define W(i, k) r->w[(i) * n + (k)]

/* ---- A.3 geometry / strains (cosserat_rod.py: _compute_geometry_from_state ..) */
static void compute_shear_stretch_strains(ro_rod *r) {
  const int n = r->n;
  for (int k = 0; k < n; k++) {
    double d0 = X(0, k + 1) - X(0, k), d1 = X(1, k + 1) - X(1, k), d2 = X(2, k + 1) - X(2, k);
    double len = sqrt(d0 * d0 + d1 * d1 + d2 * d2) + 1e-14;
    r->len[k] = len;
    r->tang[0 * n + k] = d0 / len;
    r->tang[1 * n + k] = d1 / len;
    r->tang[2 * n + k] = d2 / len;
    r->radius[k] = sqrt(r->volume[k] / len / PI);
    r->dil[k] = len / r->rest_len[k];
  }
  for (int k = 0; k < n - 1; k++)
    r->vdil[k] = (0.5 * (r->len[k + 1] + r->len[k])) / r->rest_vor[k];
  for (int k = 0; k < n; k++)
    for (int i = 0; i < 3; i++) {
      double qt = 0.0;
      for (int j = 0; j < 3; j++) qt += QQ(i, j, k) * r->tang[j * n + k];
      r->sigma[i * n + k] = r->dil[k] * qt - (i == 2 ? 1.0 : 0.0);
    }
}

static void compute_bending_twist_strains(ro_rod *r) {
  const int n = r->n;
  for (int k = 0; k < n - 1; k++) {
    double Rm[3][3];
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++)
        Rm[i][j] = QQ(i, 0, k + 1) * QQ(j, 0, k) + QQ(i, 1, k + 1) * QQ(j, 1, k) +
                   QQ(i, 2, k + 1) * QQ(j, 2, k);
    double v0 = Rm[2][1] - Rm[1][2], v1 = Rm[0][2] - Rm[2][0], v2 = Rm[1][0] - Rm[0][1];
    double trace = Rm[0][0] + Rm[1][1] + Rm[2][2];
    double theta = acos(0.5 * trace - 0.5 - 1e-10);
    double fac = -0.5 * theta / sin(theta + 1e-14);
    r->kappa[0 * (n - 1) + k] = (v0 * fac) / r->rest_vor[k];
    r->kappa[1 * (n - 1) + k] = (v1 * fac) / r->rest_vor[k];
    r->kappa[2 * (n - 1) + k] = (v2 * fac) / r->rest_vor[k];
  }
}

static void compute_internal_forces_and_torques(ro_rod *r) {
  const int n = r->n, nv = n - 1;
  double *cs = r->tmp; /* (3,n) Q^T n / e */
  compute_shear_stretch_strains(r);
  for (int k = 0; k < n; k++)
    for (int i = 0; i < 3; i++)
      r->stress[i * n + k] = r->S[i * n + k] * (r->sigma[i * n + k] - r->rest_sigma[i * n + k]);
  for (int k = 0; k < n; k++)
    for (int i = 0; i < 3; i++) {
      double s = 0.0;
      for (int j = 0; j < 3; j++) s += QQ(j, i, k) * r->stress[j * n + k];
      cs[i * n + k] = s / r->dil[k];
    }
  for (int i = 0; i < 3; i++) { /* Delta_h */
    r->f_int[i * (n + 1) + 0] = cs[i * n + 0];
    for (int k = 1; k < n; k++) r->f_int[i * (n + 1) + k] = cs[i * n + k] - cs[i * n + k - 1];
    r->f_int[i * (n + 1) + n] = -cs[i * n + n - 1];
  }

  compute_bending_twist_strains(r);
  for (int k = 0; k < nv; k++)
    for (int i = 0; i < 3; i++)
      r->couple[i * nv + k] = r->B[i * nv + k] * (r->kappa[i * nv + k] - r->rest_kappa[i * nv + k]);
  /* dilatation rate */
  for (int k = 0; k < n; k++) {
    double rv0 = 0, rv1 = 0, rp1v = 0, rvp1 = 0;
    for (int i = 0; i < 3; i++) {
      rv0 += X(i, k) * V(i, k);
      rv1 += X(i, k + 1) * V(i, k + 1);
      rp1v += X(i, k + 1) * V(i, k);
      rvp1 += X(i, k) * V(i, k + 1);
    }
    r->dil_rate[k] = (rv0 + rv1 - rvp1 - rp1v) / r->len[k] / r->rest_len[k];
  }
  double *m2 = r->tmp + 3 * (n + 1);     /* tau / eps^3          (3,nv) */
  double *m3 = r->tmp + 6 * (n + 1);     /* (kappa x tau) D / eps^3 (3,nv) */
  for (int k = 0; k < nv; k++) {
    double e3 = 1.0 / (r->vdil[k] * r->vdil[k] * r->vdil[k]);
    double k0 = r->kappa[k], k1 = r->kappa[nv + k], k2 = r->kappa[2 * nv + k];
    double c0 = r->couple[k], c1 = r->couple[nv + k], c2 = r->couple[2 * nv + k];
    m2[k] = c0 * e3; m2[nv + k] = c1 * e3; m2[2 * nv + k] = c2 * e3;
    m3[k] = (k1 * c2 - k2 * c1) * r->rest_vor[k] * e3;
    m3[nv + k] = (k2 * c0 - k0 * c2) * r->rest_vor[k] * e3;
    m3[2 * nv + k] = (k0 * c1 - k1 * c0) * r->rest_vor[k] * e3;
  }
  for (int k = 0; k < n; k++) {
    double bt2[3], bt3[3], qt[3], ssc[3], jw[3], lt[3], ud[3];
    for (int i = 0; i < 3; i++) {
      if (k == 0) { bt2[i] = m2[i * nv]; bt3[i] = 0.5 * m3[i * nv]; }
      else if (k == n - 1) { bt2[i] = -m2[i * nv + nv - 1]; bt3[i] = 0.5 * m3[i * nv + nv - 1]; }
      else { bt2[i] = m2[i * nv + k] - m2[i * nv + k - 1]; bt3[i] = 0.5 * (m3[i * nv + k] + m3[i * nv + k - 1]); }
      double a = 0.0;
      for (int j = 0; j < 3; j++) a += QQ(i, j, k) * r->tang[j * n + k];
      qt[i] = a;
    }
    double n0 = r->stress[k], n1 = r->stress[n + k], n2 = r->stress[2 * n + k];
    ssc[0] = (qt[1] * n2 - qt[2] * n1) * r->rest_len[k];
    ssc[1] = (qt[2] * n0 - qt[0] * n2) * r->rest_len[k];
    ssc[2] = (qt[0] * n1 - qt[1] * n0) * r->rest_len[k];
    for (int i = 0; i < 3; i++) jw[i] = (r->J[i * n + k] * W(i, k)) / r->dil[k];
    lt[0] = jw[1] * W(2, k) - jw[2] * W(1, k);
    lt[1] = jw[2] * W(0, k) - jw[0] * W(2, k);
    lt[2] = jw[0] * W(1, k) - jw[1] * W(0, k);
    for (int i = 0; i < 3; i++) ud[i] = jw[i] * r->dil_rate[k] / r->dil[k];
    for (int i = 0; i < 3; i++)
      r->t_int[i * n + k] = bt2[i] + bt3[i] + ssc[i] + lt[i] + ud[i];
  }
}

/* ---- A.2.1 kinematic half step */
static void rotation_matrix(double v0, double v1, double v2, double R[3][3]) {
  /* _get_rotation_matrix(1.0, axis): axis / (|axis| + 1e-14), Rodrigues, transpose convention */
  double theta = sqrt(v0 * v0 + v1 * v1 + v2 * v2);
  v0 /= theta + 1e-14; v1 /= theta + 1e-14; v2 /= theta + 1e-14;
  theta = theta * 1.0;
  double up = sin(theta), us = 1.0 - cos(theta);
  R[0][0] = 1.0 - us * (v1 * v1 + v2 * v2);
  R[1][1] = 1.0 - us * (v0 * v0 + v2 * v2);
  R[2][2] = 1.0 - us * (v0 * v0 + v1 * v1);
  R[0][1] = up * v2 + us * v0 * v1;
  R[1][0] = -up * v2 + us * v0 * v1;
  R[0][2] = -up * v1 + us * v0 * v2;
  R[2][0] = up * v1 + us * v0 * v2;
  R[1][2] = up * v0 + us * v1 * v2;
  R[2][1] = -up * v0 + us * v1 * v2;
}

static void kinematic_step(ro_rod *r, double prefac) {
  const int n = r->n;
  for (int i = 0; i < 3; i++)
    for (int k = 0; k <= n; k++) X(i, k) += prefac * V(i, k);
  for (int k = 0; k < n; k++) {
    double R[3][3], Qn[3][3];
    rotation_matrix(prefac * W(0, k), prefac * W(1, k), prefac * W(2, k), R);
    for (int i = 0; i < 3; i++)
      for (int m = 0; m < 3; m++) {
        double s = 0.0;
        for (int j = 0; j < 3; j++) s += R[i][j] * QQ(j, m, k);
        Qn[i][m] = s;
      }
    for (int i = 0; i < 3; i++)
      for (int m = 0; m < 3; m++) QQ(i, m, k) = Qn[i][m];
  }
}

/* ---- plugins */
static void constrain_values(ro_rod *r, const double *base_pos) {
  const int n = r->n;
  switch (r->cfg.bc_kind) {
    case RO_BC_ONE_END_FIXED:
      for (int i = 0; i < 3; i++) X(i, 0) = r->fixed_pos[i];
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) QQ(i, j, 0) = r->fixed_Q[i * 3 + j];
      break;
    case RO_BC_PENDULUM_SLIDER: /* build.py:71-74 */
      X(1, 0) = r->fixed_pos[1];
      X(2, 0) = r->fixed_pos[2];
      for (int j = 0; j < 3; j++) { QQ(0, j, 0) = r->fixed_Q[0 * 3 + j]; QQ(2, j, 0) = r->fixed_Q[2 * 3 + j]; }
      break;
    case RO_BC_MOVING_BASE: /* soft_pendulum_3d/build.py:32-35 */
      X(0, 0) = base_pos[0]; X(1, 0) = base_pos[1]; X(2, 0) = r->fixed_pos[2];
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) QQ(i, j, 0) = r->fixed_Q[i * 3 + j];
      break;
    default: break;
  }
}

static void constrain_rates(ro_rod *r, const double *base_vel) {
  const int n = r->n;
  switch (r->cfg.bc_kind) {
    case RO_BC_ONE_END_FIXED:
      for (int i = 0; i < 3; i++) { V(i, 0) = 0.0; W(i, 0) = 0.0; }
      break;
    case RO_BC_PENDULUM_SLIDER: /* build.py:76-79 */
      V(1, 0) = 0.0; V(2, 0) = 0.0; W(0, 0) = 0.0; W(2, 0) = 0.0;
      break;
    case RO_BC_MOVING_BASE: /* soft_pendulum_3d/build.py:37-40 */
      V(0, 0) = base_vel[0]; V(1, 0) = base_vel[1]; V(2, 0) = 0.0;
      for (int i = 0; i < 3; i++) W(i, 0) = 0.0;
      break;
    default: break;
  }
}

static void laplace_filter(double *rate, double *ft, int m, int order) {
  /* A.4: nb_filter_rate on a (3,m) array */
  double *nw = (double *)malloc(sizeof(double) * (size_t)m);
  for (int i = 0; i < 3; i++) {
    double *f = ft + i * m, *q = rate + i * m;
    memcpy(f, q, sizeof(double) * (size_t)m);
    for (int p = 0; p < order; p++) {
      for (int k = 1; k < m - 1; k++) nw[k] = (-f[k + 1] - f[k - 1] + 2.0 * f[k]) / 4.0;
      for (int k = 1; k < m - 1; k++) f[k] = nw[k];
      f[0] = 0.0; f[m - 1] = 0.0;
    }
    for (int k = 0; k < m; k++) q[k] = q[k] - f[k];
  }
  free(nw);
}

static void dampen_rates(ro_rod *r) {
  const int n = r->n;
  if (r->cfg.damping_constant >= 0.0) {
    for (int i = 0; i < 3; i++)
      for (int k = 0; k <= n; k++) V(i, k) = V(i, k) * r->c_v;
    for (int i = 0; i < 3; i++)
      for (int k = 0; k < n; k++) W(i, k) = W(i, k) * pow(r->c_w[i * n + k], r->dil[k]);
  }
  if (r->cfg.laplace_filter_order > 0) {
    laplace_filter(r->v, r->filt, n + 1, r->cfg.laplace_filter_order);
    laplace_filter(r->w, r->filt, n, r->cfg.laplace_filter_order);
  }
}

/* ---- A.5 rod-plane contact with anisotropic friction
 * (elastica/_contact_functions.py: _calculate_contact_forces_rod_plane[_with_anisotropic_friction]) */
static double slip_fn(const double v[3], double tol) {
  double a = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  if (fabs(a) > tol) { double q = a / tol - 1.0; if (q > 1.0) q = 1.0; return fabs(1.0 - q); }
  return 1.0;
}
static double sgn(double x) { return (x > 0) - (x < 0); }

static void node_to_element_force(const ro_rod *r, double *out /* (3,n) */) {
  const int n = r->n;
  for (int i = 0; i < 3; i++) {
    const double *fi = r->f_int + i * (n + 1), *fe = r->f_ext + i * (n + 1);
    for (int k = 0; k < n; k++) out[i * n + k] = 0.0 + 0.5 * ((fi[k] + fe[k]) + (fi[k + 1] + fe[k + 1]));
    out[i * n + 0] += 0.5 * (fi[0] + fe[0]);
    out[i * n + n - 1] += 0.5 * (fi[n] + fe[n]);
  }
}
static void elements_to_nodes(ro_rod *r, const double *fe /* (3,n) */) {
  const int n = r->n;
  for (int i = 0; i < 3; i++)
    for (int k = 0; k < n; k++) {
      r->f_ext[i * (n + 1) + k] += 0.5 * fe[i * n + k];
      r->f_ext[i * (n + 1) + k + 1] += 0.5 * fe[i * n + k];
    }
}

static void apply_contact(ro_rod *r) {
  const int n = r->n;
  const ro_config *c = &r->cfg;
  const double *N = c->plane_normal;
  double *etf = r->ctmp, *resp_mag = r->ctmp + 3 * n, *force = r->ctmp + 4 * n, *etf2 = r->ctmp + 7 * n;
  unsigned char *nocontact = (unsigned char *)(r->ctmp + 10 * n);
  double evel[3], ax[3], rolldir[3];
  /* normal response */
  node_to_element_force(r, etf);
  for (int k = 0; k < n; k++) {
    double fn = N[0] * etf[k] + N[1] * etf[n + k] + N[2] * etf[2 * n + k];
    double resp[3];
    for (int i = 0; i < 3; i++) resp[i] = -(fn > 0 ? 0.0 : N[i] * fn);
    double ep[3];
    for (int i = 0; i < 3; i++) ep[i] = 0.5 * (X(i, k) + X(i, k + 1));
    double dist = N[0] * (ep[0] - c->plane_origin[0]) + N[1] * (ep[1] - c->plane_origin[1]) + N[2] * (ep[2] - c->plane_origin[2]);
    double pen = dist - r->radius[k]; if (pen > 0.0) pen = 0.0;
    double msum = r->mass[k + 1] + r->mass[k];
    for (int i = 0; i < 3; i++) evel[i] = (r->mass[k + 1] * V(i, k + 1) + r->mass[k] * V(i, k)) / msum;
    double vn = N[0] * evel[0] + N[1] * evel[1] + N[2] * evel[2];
    nocontact[k] = (dist - r->radius[k]) > c->surface_tol;
    for (int i = 0; i < 3; i++) {
      double tot = resp[i] + (-c->contact_k * (N[i] * pen)) + (-c->contact_nu * (N[i] * vn));
      if (nocontact[k]) { resp[i] = 0.0; tot = 0.0; }
      force[i * n + k] = tot;
    }
    resp_mag[k] = sqrt(resp[0] * resp[0] + resp[1] * resp[1] + resp[2] * resp[2]);
  }
  elements_to_nodes(r, force);
  /* kinetic friction; per-element directions are kept for the static part */
  double *axd = r->ctmp + 11 * n, *rld = r->ctmp + 14 * n, *slipa = r->ctmp + 17 * n, *slipr = r->ctmp + 18 * n;
  double *fa = r->ctmp + 19 * n, *fr = r->ctmp + 22 * n;
  for (int k = 0; k < n; k++) {
    double t[3] = {r->tang[k], r->tang[n + k], r->tang[2 * n + k]};
    double tn = N[0] * t[0] + N[1] * t[1] + N[2] * t[2];
    double tp[3] = {t[0] - N[0] * tn, t[1] - N[1] * tn, t[2] - N[2] * tn};
    double tpm = sqrt(tp[0] * tp[0] + tp[1] * tp[1] + tp[2] * tp[2]);
    double inv = 1 / (tpm + 1e-14);
    for (int i = 0; i < 3; i++) ax[i] = inv * tp[i];
    double msum = r->mass[k + 1] + r->mass[k];
    for (int i = 0; i < 3; i++) evel[i] = (r->mass[k + 1] * V(i, k + 1) + r->mass[k] * V(i, k)) / msum;
    double vax = evel[0] * ax[0] + evel[1] * ax[1] + evel[2] * ax[2];
    double va[3] = {vax * ax[0], vax * ax[1], vax * ax[2]};
    double sg = sgn(vax);
    double kmu = 0.5 * (c->kinetic_mu[0] * (1 + sg) + c->kinetic_mu[1] * (1 - sg));
    slipa[k] = slip_fn(va, c->slip_velocity_tol);
    rolldir[0] = ax[1] * N[2] - ax[2] * N[1]; rolldir[1] = ax[2] * N[0] - ax[0] * N[2]; rolldir[2] = ax[0] * N[1] - ax[1] * N[0];
    double arm[3] = {-N[0] * r->radius[k], -N[1] * r->radius[k], -N[2] * r->radius[k]};
    double vroll = evel[0] * rolldir[0] + evel[1] * rolldir[1] + evel[2] * rolldir[2];
    double qa[3], wq[3], rv[3];
    for (int i = 0; i < 3; i++) qa[i] = 0.0 + QQ(i, 0, k) * arm[0] + QQ(i, 1, k) * arm[1] + QQ(i, 2, k) * arm[2];
    wq[0] = W(1, k) * qa[2] - W(2, k) * qa[1]; wq[1] = W(2, k) * qa[0] - W(0, k) * qa[2]; wq[2] = W(0, k) * qa[1] - W(1, k) * qa[0];
    for (int i = 0; i < 3; i++) rv[i] = 0.0 + QQ(0, i, k) * wq[0] + QQ(1, i, k) * wq[1] + QQ(2, i, k) * wq[2];
    double rvr = rv[0] * rolldir[0] + rv[1] * rolldir[1] + rv[2] * rolldir[2];
    double smag = vroll + rvr;
    double sv[3] = {smag * rolldir[0], smag * rolldir[1], smag * rolldir[2]};
    slipr[k] = slip_fn(sv, c->slip_velocity_tol);
    double ut[3] = {sv[0] + va[0], sv[1] + va[1], sv[2] + va[2]};
    double un = sqrt((ut[0] + 1e-14) * (ut[0] + 1e-14) + (ut[1] + 1e-14) * (ut[1] + 1e-14) + (ut[2] + 1e-14) * (ut[2] + 1e-14));
    for (int i = 0; i < 3; i++) ut[i] /= un;
    double uax = ut[0] * ax[0] + ut[1] * ax[1] + ut[2] * ax[2];
    double url = ut[0] * rolldir[0] + ut[1] * rolldir[1] + ut[2] * rolldir[2];
    for (int i = 0; i < 3; i++) {
      fa[i * n + k] = nocontact[k] ? 0.0 : -((1.0 - slipa[k]) * kmu * resp_mag[k] * uax * ax[i]);
      fr[i * n + k] = nocontact[k] ? 0.0 : -((1.0 - slipr[k]) * c->kinetic_mu[2] * resp_mag[k] * url * rolldir[i]);
      axd[i * n + k] = ax[i]; rld[i * n + k] = rolldir[i];
    }
    double frk[3] = {fr[k], fr[n + k], fr[2 * n + k]};
    double cr[3] = {arm[1] * frk[2] - arm[2] * frk[1], arm[2] * frk[0] - arm[0] * frk[2], arm[0] * frk[1] - arm[1] * frk[0]};
    for (int i = 0; i < 3; i++) r->t_ext[i * n + k] += 0.0 + QQ(i, 0, k) * cr[0] + QQ(i, 1, k) * cr[1] + QQ(i, 2, k) * cr[2];
  }
  elements_to_nodes(r, fa);
  elements_to_nodes(r, fr);
  /* static friction: forces re-collected with the responses added above */
  node_to_element_force(r, etf2);
  for (int k = 0; k < n; k++) {
    for (int i = 0; i < 3; i++) { ax[i] = axd[i * n + k]; rolldir[i] = rld[i * n + k]; }
    double fax = etf2[k] * ax[0] + etf2[n + k] * ax[1] + etf2[2 * n + k] * ax[2];
    double sg = sgn(fax);
    double smu = 0.5 * (c->static_mu[0] * (1 + sg) + c->static_mu[1] * (1 - sg));
    double maxf = slipa[k] * smu * resp_mag[k];
    double mag = fabs(fax) < maxf ? fabs(fax) : maxf;
    for (int i = 0; i < 3; i++) fa[i * n + k] = nocontact[k] ? 0.0 : -(mag * sg * ax[i]);
    double tt[3];
    for (int i = 0; i < 3; i++)
      tt[i] = 0.0 + QQ(0, i, k) * (r->t_int[k] + r->t_ext[k]) + QQ(1, i, k) * (r->t_int[n + k] + r->t_ext[n + k]) +
              QQ(2, i, k) * (r->t_int[2 * n + k] + r->t_ext[2 * n + k]);
    double tta = tt[0] * ax[0] + tt[1] * ax[1] + tt[2] * ax[2];
    double frl = etf2[k] * rolldir[0] + etf2[n + k] * rolldir[1] + etf2[2 * n + k] * rolldir[2];
    double noslip = -((r->radius[k] * frl - 2.0 * tta) / 3.0 / r->radius[k]);
    double maxr = slipr[k] * c->static_mu[2] * resp_mag[k];
    double magr = fabs(noslip) < maxr ? fabs(noslip) : maxr;
    double sgr = sgn(noslip);
    for (int i = 0; i < 3; i++) fr[i * n + k] = nocontact[k] ? 0.0 : (magr * sgr * rolldir[i]);
  }
  elements_to_nodes(r, fa);
  elements_to_nodes(r, fr);
  for (int k = 0; k < n; k++) {
    double arm[3] = {-N[0] * r->radius[k], -N[1] * r->radius[k], -N[2] * r->radius[k]};
    double frk[3] = {fr[k], fr[n + k], fr[2 * n + k]};
    double cr[3] = {arm[1] * frk[2] - arm[2] * frk[1], arm[2] * frk[0] - arm[0] * frk[2], arm[0] * frk[1] - arm[1] * frk[0]};
    for (int i = 0; i < 3; i++) r->t_ext[i * n + k] += 0.0 + QQ(i, 0, k) * cr[0] + QQ(i, 1, k) * cr[1] + QQ(i, 2, k) * cr[2];
  }
}

/* ---- MuscleTorques.apply_torques (PyElastica external_forces.py [PE-recall]; restated in
 * oracle/shims/elastica/external_forces.py): magnitudes beta(s) sin(w t - k s + phi), ramped, walked tail-to-head,
 * applied as equal and opposite couples on consecutive elements in the material frame. */
static void apply_muscle_torques(ro_rod *r) {
  const int n = r->n;
  const double t = r->time, kw = r->muscle[0], *beta = r->muscle + 1, *d = r->cfg.muscle_direction;
  const double factor = fmin(1.0, t / r->cfg.muscle_ramp_up_time);
  const double omega = 2.0 * M_PI / r->cfg.muscle_period;
  double *mag = r->tmp; /* (n) scratch */
  double total = 0.0, cum = 0.0;
  for (int k = 0; k < n; k++) total += r->rest_len[k];   /* s = cumsum(rest_lengths) (left to right); s /= s[-1] */
  for (int i = 0; i < n; i++) {
    cum += r->rest_len[i];
    const double s = cum / total;
    mag[i] = factor * beta[i] * sin(omega * t - kw * s + r->cfg.muscle_phase_shift);
  }
  for (int k = 0; k < n; k++) {
    const double mk = mag[n - 1 - k];                       /* torque_mag[::-1] */
    const double mk1 = (k + 1 < n) ? mag[n - 2 - k] : 0.0;
    for (int i = 0; i < 3; i++) {
      const double qd = QQ(i, 0, k) * d[0] + QQ(i, 1, k) * d[1] + QQ(i, 2, k) * d[2];
      if (k >= 1) r->t_ext[i * n + k] += qd * mk;
      if (k + 1 < n) r->t_ext[i * n + k] -= qd * mk1;
    }
  }
}

/* ---- MuscleTorquesWithVaryingBetaSplines.apply_torques (muscle_torques_with_bspline.py:126-160, filter :206-228,
 * compute :181-204).  The spline is scipy's make_interp_spline(x, y) = the not-a-knot interpolating cubic; it is
 * solved here for its second derivatives and evaluated in the classical two-sided form. */
static double notaknot_eval(int N, double dx, const double *y, const double *M, double s) {
  int m = (int)floor(s / dx);
  if (m < 0) m = 0;
  if (m > N - 1) m = N - 1;                 /* beyond the last knot: the last piece continues (extrapolate=True) */
  const double a = (m + 1) * dx - s, b = s - m * dx;
  return M[m] * a * a * a / (6.0 * dx) + M[m + 1] * b * b * b / (6.0 * dx) +
         (y[m] / dx - M[m] * dx / 6.0) * a + (y[m + 1] / dx - M[m + 1] * dx / 6.0) * b;
}

static void notaknot_second_derivatives(int N, double dx, const double *y, double *M) {
  /* unknowns M_0..M_N; rows: third-derivative continuity at x_1 and x_{N-1}, C2 conditions in between */
  const int K = N + 1;
  double A[18][19];
  memset(A, 0, sizeof(A));
  A[0][0] = 1.0; A[0][1] = -2.0; A[0][2] = 1.0;
  A[N][N] = 1.0; A[N][N - 1] = -2.0; A[N][N - 2] = 1.0;
  for (int m = 1; m < N; m++) {
    A[m][m - 1] = 1.0; A[m][m] = 4.0; A[m][m + 1] = 1.0;
    A[m][K] = 6.0 * (y[m - 1] - 2.0 * y[m] + y[m + 1]) / (dx * dx);
  }
  for (int c = 0; c < K; c++) {
    int piv = c;
    for (int q = c + 1; q < K; q++) if (fabs(A[q][c]) > fabs(A[piv][c])) piv = q;
    if (piv != c) for (int k = 0; k <= K; k++) { double t = A[c][k]; A[c][k] = A[piv][k]; A[piv][k] = t; }
    for (int q = c + 1; q < K; q++) {
      const double f = A[q][c] / A[c][c];
      for (int k = c; k <= K; k++) A[q][k] -= f * A[c][k];
    }
  }
  for (int c = K - 1; c >= 0; c--) {
    double acc = A[c][K];
    for (int k = c + 1; k < K; k++) acc -= A[c][k] * M[k];
    M[c] = acc / A[c][c];
  }
}

static void apply_spline_torques(ro_rod *r) {
  const int n = r->n, P = r->cfg.spline_n_ctrl, N = P + 1;
  const double dx = r->cfg.base_length / N;
  for (int d = 0; d < 3; d++) {
    if (!(r->cfg.spline_dir_mask >> d & 1)) continue;
    double *tgt = r->spl_pts + d * (2 * P + 1), *cached = tgt + P, *flag = tgt + 2 * P, *mag = r->spl_mag + d * n;
    int differ = (*flag == 0.0);
    for (int i = 0; i < P; i++) differ = differ || !(cached[i] == tgt[i]);      /* not np.array_equal */
    if (differ) {
      *flag = 1.0;
      for (int i = 0; i < P; i++) {                                             /* filter_activation */
        const double diff = tgt[i] - cached[i];
        const double sg = (diff > 0.0) - (diff < 0.0);
        cached[i] += sg * fmin(r->cfg.spline_max_rate, fabs(diff));
      }
      double y[18], M[18];
      y[0] = 0.0; y[N] = 0.0;
      for (int i = 0; i < P; i++) y[1 + i] = cached[i];
      notaknot_second_derivatives(N, dx, y, M);
      double s = 0.0;
      for (int k = 0; k < n; k++) {                                             /* np.cumsum(system.lengths) */
        s += r->len[k];
        mag[k] = r->cfg.spline_scale * notaknot_eval(N, dx, y, M, s);
      }
    }
    for (int k = 0; k < n; k++) r->t_ext[d * n + k] += mag[k];                  /* compute_muscle_torques */
  }
}

/* ---- A.2 one PositionVerlet substep, in the phases the multi-system stepper interleaves */
static void phase_first_half(ro_rod *r, const double *bp) {
  kinematic_step(r, 0.5 * r->cfg.dt);
  r->time += 0.5 * r->cfg.dt;
  constrain_values(r, bp);
}

/* ---- COOMM `ApplyMuscles` with one active `TransverseMuscle` ------------------------------------------
 * Third-party: coomm 0.1.1, git rev d33fa034 of hanson-hschang/COOMM (branch refactor-numba-hotloops),
 * pinned at /root/reference/uv.lock:172-179 — NOT in /root/reference and not obtainable offline.  PARITY
 * UNPINNED: this restates the published model (Chang et al., "Energy-shaping control of a muscular octopus
 * arm moving in three dimensions", Proc. R. Soc. A 2023, and the public COOMM package layout
 * coomm/actuations/muscles/{muscle,transverse_muscle}.py) as recalled; anchored on the reference's call sites:
 *   create_es_muscle_layers   /root/reference/gym_softrobot/envs/octopus/build.py:292-338
 *     TransverseMuscle(rest_muscle_area=(radius/radius_base)**2, max_muscle_stress=1.0)
 *   ApplyMuscles registration build_muscle_octopus.py:165-177, arm_push_env.py:198-209
 *   apply_activation(scalar)  crawl_env.py:242, arm_push_env.py:259,271
 * In the envs built here only the transverse muscle (index 2) is ever activated; the two longitudinal
 * muscles keep zero activation and a muscle's force is proportional to its activation, so they add exact
 * zeros (oracle/shims/coomm restates them too and the fixtures run through them).
 * Model, per element k (material frame):
 *   muscle position  = 0 (the TM acts on the centre line)
 *   muscle strain    nu = sigma + e3            (= e Q t)
 *   muscle tangent   t_m = nu / |nu|
 *   muscle length    l = 1 / sqrt(|nu|)         (radial fibres of an incompressible arm: r / r0)
 *   muscle area      A = rest_area / e
 *   weight           h(l) = max(3.06 l^3 - 13.64 l^2 + 18.01 l - 6.44, 0)
 *   internal force   n_m = activation * (-max_stress) * A * h(l) * t_m      (contraction lengthens the arm)
 *   internal couple  0
 * and, as for any continuous actuation, the loads handed to the rod are
 *   external_forces  (nodes, lab)      += Delta_h( Q^T n_m )
 *   external_torques (elements, mat.)  += (Q t e) x n_m * rest_length       (round-off zero for the TM) */
static void apply_tm_muscle(ro_rod *r) {
  const int n = r->n;
  double *fl = r->tmp + 9 * (n + 1); /* (3,n) lab-frame internal muscle force */
  for (int k = 0; k < n; k++) {
    double nu[3], nm[3], qt[3];
    for (int i = 0; i < 3; i++) nu[i] = r->sigma[i * n + k] + (i == 2 ? 1.0 : 0.0);
    const double len = sqrt(nu[0] * nu[0] + nu[1] * nu[1] + nu[2] * nu[2]);
    const double l = 1.0 / sqrt(len);
    double h = ((3.06 * l - 13.64) * l + 18.01) * l - 6.44;
    if (h < 0.0) h = 0.0;
    const double a0 = (r->rest_radius[k] / r->tm_radius_ref) * (r->rest_radius[k] / r->tm_radius_ref);
    const double F = r->tm_activation * (-r->tm_max_stress) * (a0 / r->dil[k]) * h;
    for (int i = 0; i < 3; i++) nm[i] = F * (nu[i] / len);
    for (int i = 0; i < 3; i++) {
      double s = 0.0, a = 0.0;
      for (int j = 0; j < 3; j++) { s += QQ(j, i, k) * nm[j]; a += QQ(i, j, k) * r->tang[j * n + k]; }
      fl[i * n + k] = s;
      qt[i] = a * r->dil[k];
    }
    r->t_ext[0 * n + k] += (qt[1] * nm[2] - qt[2] * nm[1]) * r->rest_len[k];
    r->t_ext[1 * n + k] += (qt[2] * nm[0] - qt[0] * nm[2]) * r->rest_len[k];
    r->t_ext[2 * n + k] += (qt[0] * nm[1] - qt[1] * nm[0]) * r->rest_len[k];
  }
  for (int i = 0; i < 3; i++) {
    r->f_ext[i * (n + 1) + 0] += fl[i * n + 0];
    for (int k = 1; k < n; k++) r->f_ext[i * (n + 1) + k] += fl[i * n + k] - fl[i * n + k - 1];
    r->f_ext[i * (n + 1) + n] += -fl[i * n + n - 1];
  }
}

/* ---- COOMM `ApplyMuscles` over a list of muscle layers with per-element activations (OctoReach-v0, OctoArmTwo-v0:
 * /root/reference/gym_softrobot/envs/octopus/reach_env.py:214-227, arm_two_env.py:222-247; layers from
 * create_es_muscle_layers, envs/octopus/build.py:292-338).  Same unpinned restatement as apply_tm_muscle, general in
 * the muscle's material-frame offset x_m = ratio * radius (oracle/shims/coomm/actuations/muscles/muscle.py):
 *   nu_m = sigma + e3 + kappa_e x x_m + d x_m / ds      kappa_e = trapezoid of kappa with zero ghosts,
 *                                                        d/ds = central difference over element centres
 *   l_m = |nu_m| (longitudinal) or 1 / sqrt(|nu_m|) (transverse), t_m = nu_m / |nu_m|, A = rest_area / e
 *   n_m = a sigma_max A h(l_m) t_m,   c_m = x_m x n_m (elements),   m_m = Voronoi average of c_m
 *   external_forces += Delta_h(Q^T n_m);  external_torques += Delta_h(m_m) + A_h(kappa x m_m D) + (Q t e) x n_m l0 */
static void apply_muscle_layers(ro_rod *r) {
  const int n = r->n, nv = n - 1;
  double *pos = r->ml_tmp, *frc = pos + 3 * (n + 1), *cpl = frc + 3 * (n + 1), *fl = cpl + 3 * (n + 1), *mv = fl + 3 * (n + 1);
  for (int m = 0; m < 3; m++) {
    if (!r->ml_kind[m]) continue;
    const double ratio[3] = {r->ml_px[m], r->ml_py[m], 0.0};
    for (int k = 0; k < n; k++)
      for (int i = 0; i < 3; i++) pos[i * n + k] = ratio[i] * r->radius[k];
    for (int k = 0; k < n; k++) {
      double nu[3], ke[3], dp[3], st[3], xm[3] = {pos[k], pos[n + k], pos[2 * n + k]};
      for (int i = 0; i < 3; i++) {
        nu[i] = r->sigma[i * n + k] + (i == 2 ? 1.0 : 0.0);
        const double kl = k > 0 ? r->kappa[i * nv + k - 1] : 0.0, kr = k < nv ? r->kappa[i * nv + k] : 0.0;
        ke[i] = (k == 0) ? 0.5 * kr : (k == n - 1) ? 0.5 * kl : 0.5 * (kr + kl);
        dp[i] = 0.0;
      }
      if (n > 2) {   /* element centres s_k = cumsum(rest_len)_k - rest_len_k / 2 */
        const int a = k == 0 ? 0 : k - 1, b = k == n - 1 ? n - 1 : k + 1;
        double sa = 0.0, sb = 0.0, acc = 0.0;
        for (int q = 0; q <= b; q++) { acc += r->rest_len[q]; if (q == a) sa = acc - 0.5 * r->rest_len[q]; if (q == b) sb = acc - 0.5 * r->rest_len[q]; }
        for (int i = 0; i < 3; i++) dp[i] = (pos[i * n + b] - pos[i * n + a]) / (sb - sa);
      }
      st[0] = nu[0] + (ke[1] * xm[2] - ke[2] * xm[1]) + dp[0];
      st[1] = nu[1] + (ke[2] * xm[0] - ke[0] * xm[2]) + dp[1];
      st[2] = nu[2] + (ke[0] * xm[1] - ke[1] * xm[0]) + dp[2];
      const double len = sqrt(st[0] * st[0] + st[1] * st[1] + st[2] * st[2]);
      const double l = r->ml_kind[m] == 2 ? 1.0 / sqrt(len) : len;
      double h = ((3.06 * l - 13.64) * l + 18.01) * l - 6.44;
      if (h < 0.0) h = 0.0;
      const double a0 = (r->rest_radius[k] / r->ml_rref[m]) * (r->rest_radius[k] / r->ml_rref[m]);
      const double F = r->ml_act[m * n + k] * r->ml_stress[m] * (a0 / r->dil[k]) * h;
      double nm[3] = {F * (st[0] / len), F * (st[1] / len), F * (st[2] / len)}, qt[3];
      for (int i = 0; i < 3; i++) frc[i * n + k] = nm[i];
      cpl[0 * n + k] = xm[1] * nm[2] - xm[2] * nm[1];
      cpl[1 * n + k] = xm[2] * nm[0] - xm[0] * nm[2];
      cpl[2 * n + k] = xm[0] * nm[1] - xm[1] * nm[0];
      for (int i = 0; i < 3; i++) {
        double s_ = 0.0, a_ = 0.0;
        for (int j = 0; j < 3; j++) { s_ += QQ(j, i, k) * nm[j]; a_ += QQ(i, j, k) * (r->tang[j * n + k] * r->dil[k]); }
        fl[i * n + k] = s_;
        qt[i] = a_;
      }
      r->t_ext[0 * n + k] += (qt[1] * nm[2] - qt[2] * nm[1]) * r->rest_len[k];
      r->t_ext[1 * n + k] += (qt[2] * nm[0] - qt[0] * nm[2]) * r->rest_len[k];
      r->t_ext[2 * n + k] += (qt[0] * nm[1] - qt[1] * nm[0]) * r->rest_len[k];
    }
    for (int i = 0; i < 3; i++) {
      r->f_ext[i * (n + 1) + 0] += fl[i * n + 0];
      for (int k = 1; k < n; k++) r->f_ext[i * (n + 1) + k] += fl[i * n + k] - fl[i * n + k - 1];
      r->f_ext[i * (n + 1) + n] += -fl[i * n + n - 1];
    }
    for (int k = 0; k < nv; k++)
      for (int i = 0; i < 3; i++) mv[i * nv + k] = 0.5 * (cpl[i * n + k + 1] + cpl[i * n + k]);
    for (int k = 0; k < n; k++) {
      for (int i = 0; i < 3; i++) {
        const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
        /* kappa x m_m * rest_voronoi_length at Voronoi points k-1 and k */
        const double cl = k > 0 ? (r->kappa[i1 * nv + k - 1] * mv[i2 * nv + k - 1] - r->kappa[i2 * nv + k - 1] * mv[i1 * nv + k - 1]) * r->rest_vor[k - 1] : 0.0;
        const double cr = k < nv ? (r->kappa[i1 * nv + k] * mv[i2 * nv + k] - r->kappa[i2 * nv + k] * mv[i1 * nv + k]) * r->rest_vor[k] : 0.0;
        const double ml = k > 0 ? mv[i * nv + k - 1] : 0.0, mr = k < nv ? mv[i * nv + k] : 0.0;
        const double dif = (k == 0) ? mr : (k == n - 1) ? -ml : mr - ml;
        const double quad = (k == 0) ? 0.5 * cr : (k == n - 1) ? 0.5 * cl : 0.5 * (cr + cl);
        r->t_ext[i * n + k] += dif + quad;
      }
    }
  }
}

static void phase_forcing(ro_rod *r, double action) {
  /* forcings in registration order: gravity, then point force (build.py:88-105), muscle / spline torques */
  const int n = r->n;
  for (int i = 0; i < 3; i++)
    for (int k = 0; k <= n; k++)
      r->f_ext[i * (n + 1) + k] += r->cfg.gravity[i] * r->mass[k] + r->f_user[i * (n + 1) + k];
  for (int i = 0; i < 3; i++)
    for (int k = 0; k < n; k++) r->t_ext[i * n + k] += r->t_user[i * n + k];
  if (r->cfg.point_force_on_base) r->f_ext[0] = action; /* assignment (build.py:101) */
  if (r->cfg.muscle_on) apply_muscle_torques(r);         /* a forcing, registered after gravity (continuum_snake.py:325-337) */
  if (r->cfg.spline_dir_mask) apply_spline_torques(r);   /* forcings (soft_arm_tracking.py:366-400) */
  if (r->tm_on) apply_tm_muscle(r);
  if (r->ml_kind[0] || r->ml_kind[1] || r->ml_kind[2]) apply_muscle_layers(r);                      /* ApplyMuscles (build_muscle_octopus.py:165-177, arm_push_env.py:204-209) */
}

static void phase_dynamic(ro_rod *r) {
  const int n = r->n;
  const double dt = r->cfg.dt;
  for (int i = 0; i < 3; i++)
    for (int k = 0; k <= n; k++)
      r->acc[i * (n + 1) + k] = (r->f_int[i * (n + 1) + k] + r->f_ext[i * (n + 1) + k]) / r->mass[k];
  for (int i = 0; i < 3; i++)
    for (int k = 0; k < n; k++)
      r->alpha[i * n + k] = (r->Jinv[i * n + k] * (r->t_int[i * n + k] + r->t_ext[i * n + k])) * r->dil[k];
  for (int i = 0; i < 3; i++)
    for (int k = 0; k <= n; k++) V(i, k) += dt * r->acc[i * (n + 1) + k];
  for (int i = 0; i < 3; i++)
    for (int k = 0; k < n; k++) W(i, k) += dt * r->alpha[i * n + k];
}

/* ControllableFixConstraint.constrain_rates (envs/octopus/controllable_constraint.py:46-69): rates of the
 * listed node / element indices are scaled by (1 - ratio); FreeBC otherwise (no values constraint) */
static void sucker_rates(ro_rod *r) {
  const int n = r->n;
  for (int q = 0; q < r->n_sucker; q++) {
    const int idx = r->sucker_idx[q], nod = r->sucker_node[q];
    const double f = 1.0 - r->sucker_ratio[q];
    for (int i = 0; i < 3; i++) { V(i, nod) *= f; W(i, idx) *= f; }
  }
}

static void phase_rates(ro_rod *r, const double *bv) {
  /* the sucker constraint is registered by the env after the build function registered the dampers
   * (crawl_env.py:146-157 after build_muscle_octopus.py:95-107; arm_push_env.py:182-195): it runs last */
  if (r->cfg.damping_before_constraints) { dampen_rates(r); constrain_rates(r, bv); }
  else { constrain_rates(r, bv); dampen_rates(r); }
  sucker_rates(r);
}

static void phase_second_half(ro_rod *r, const double *bp) {
  const int n = r->n;
  kinematic_step(r, 0.5 * r->cfg.dt);
  r->time += 0.5 * r->cfg.dt;
  constrain_values(r, bp);
  memset(r->f_ext, 0, sizeof(double) * 3 * (size_t)(n + 1));
  memset(r->t_ext, 0, sizeof(double) * 3 * (size_t)n);
}

static void substep(ro_rod *r, double action, const double *bp, const double *bv) {
  phase_first_half(r, bp);
  compute_internal_forces_and_torques(r);
  /* synchronize: [contact] forcings [contact] */
  if (r->cfg.contact_on && r->cfg.contact_before_forcing) apply_contact(r);
  phase_forcing(r, action);
  if (r->cfg.contact_on && !r->cfg.contact_before_forcing) apply_contact(r);
  phase_dynamic(r);
  phase_rates(r, bv);
  phase_second_half(r, bp);
}

void ro_substeps(ro_rod *r, int n_substeps, double action, const double *bp, const double *bv) {
  for (int s = 0; s < n_substeps; s++) substep(r, action, bp, bv);
}

/* ---- A.1 construction (factory_function.py: allocate) */
ro_rod *ro_create(const ro_config *cfg) {
  ro_rod *r = (ro_rod *)calloc(1, sizeof(ro_rod));
  const int n = cfg->n_elem, nv = n - 1;
  r->cfg = *cfg;
  r->n = n;
  r->time = 0.0;
  r->x = zalloc(3 * (n + 1)); r->v = zalloc(3 * (n + 1)); r->Q = zalloc(9 * n); r->w = zalloc(3 * n);
  r->acc = zalloc(3 * (n + 1)); r->alpha = zalloc(3 * n);
  r->rest_len = zalloc(n); r->rest_vor = zalloc(nv); r->mass = zalloc(n + 1); r->volume = zalloc(n);
  r->ml_act = zalloc(3 * n); r->ml_tmp = zalloc(16 * (n + 1));
  r->radius = zalloc(n); r->rest_radius = zalloc(n); r->J = zalloc(3 * n); r->Jinv = zalloc(3 * n); r->S = zalloc(3 * n); r->B = zalloc(3 * nv);
  r->rest_sigma = zalloc(3 * n); r->rest_kappa = zalloc(3 * nv);
  r->muscle = zalloc(1 + n);
  r->spl_pts = zalloc(3 * (2 * (size_t)(cfg->spline_n_ctrl > 0 ? cfg->spline_n_ctrl : 0) + 1));
  r->spl_mag = zalloc(3 * (size_t)n);
  r->len = zalloc(n); r->tang = zalloc(3 * n); r->dil = zalloc(n); r->vdil = zalloc(nv); r->dil_rate = zalloc(n);
  r->sigma = zalloc(3 * n); r->kappa = zalloc(3 * nv); r->stress = zalloc(3 * n); r->couple = zalloc(3 * nv);
  r->f_int = zalloc(3 * (n + 1)); r->t_int = zalloc(3 * n); r->f_ext = zalloc(3 * (n + 1)); r->t_ext = zalloc(3 * n);
  r->f_user = zalloc(3 * (n + 1)); r->t_user = zalloc(3 * n);
  r->c_w = zalloc(3 * n); r->filt = zalloc(3 * (n + 1)); r->tmp = zalloc(12 * (n + 1)); r->ctmp = zalloc(26 * n + 8);

  /* np.linspace(start, end, n+1): arange(n+1)*step + start, last point = end */
  for (int i = 0; i < 3; i++) {
    double start = cfg->start[i], end = cfg->start[i] + cfg->direction[i] * cfg->base_length;
    double step = (end - start) / (double)n;
    for (int k = 0; k <= n; k++) X(i, k) = (double)k * step + start;
    X(i, n) = end;
  }
  double nrm = sqrt(cfg->normal[0] * cfg->normal[0] + cfg->normal[1] * cfg->normal[1] +
                    cfg->normal[2] * cfg->normal[2]);
  double nor[3] = {cfg->normal[0] / nrm, cfg->normal[1] / nrm, cfg->normal[2] / nrm};
  double E = cfg->youngs_modulus, G = cfg->shear_modulus;
  if (!(G > 0.0)) G = cfg->shear_convention == 1 ? E / (0.5 + 1.0) : E / (2.0 * (1.0 + 0.5));
  double *Bel = zalloc(3 * n);
  for (int k = 0; k < n; k++) {
    double d0 = X(0, k + 1) - X(0, k), d1 = X(1, k + 1) - X(1, k), d2 = X(2, k + 1) - X(2, k);
    double rl = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
    double t[3] = {d0 / rl, d1 / rl, d2 / rl};
    r->rest_len[k] = rl;
    for (int j = 0; j < 3; j++) { QQ(0, j, k) = nor[j]; QQ(2, j, k) = t[j]; }
    QQ(1, 0, k) = t[1] * nor[2] - t[2] * nor[1];
    QQ(1, 1, k) = t[2] * nor[0] - t[0] * nor[2];
    QQ(1, 2, k) = t[0] * nor[1] - t[1] * nor[0];
    /* base_radius may be an array: np.linspace(base, tip, n_elem) (build_muscle_octopus.py:61-63) */
    double rad = cfg->base_radius;
    if (cfg->tip_radius > 0.0 && n > 1 && !cfg->taper_node_mean) {
      const double step = (cfg->tip_radius - cfg->base_radius) / (double)(n - 1);
      rad = (k == n - 1) ? cfg->tip_radius : (double)k * step + cfg->base_radius;
    } else if (cfg->tip_radius > 0.0 && cfg->taper_node_mean) {
      /* radius = np.linspace(base, tip, n_elem + 1); radius_mean = (radius[:-1] + radius[1:]) / 2
       * (/root/reference/gym_softrobot/envs/octopus/arm_push_env.py:161-164) */
      const double step = (cfg->tip_radius - cfg->base_radius) / (double)n;
      const double r0 = (double)k * step + cfg->base_radius;
      const double r1 = (k + 1 == n) ? cfg->tip_radius : (double)(k + 1) * step + cfg->base_radius;
      rad = (r0 + r1) / 2.0;
    }
    r->radius[k] = rad;
    r->rest_radius[k] = rad;
    double A0 = PI * rad * rad;
    double I1 = A0 * A0 / (4.0 * PI), I2 = I1, I3 = 2.0 * I2;
    double rho_l = cfg->density * rl;
    r->J[0 * n + k] = I1 * rho_l; r->J[1 * n + k] = I2 * rho_l; r->J[2 * n + k] = I3 * rho_l;
    for (int i = 0; i < 3; i++) r->Jinv[i * n + k] = 1.0 / r->J[i * n + k];
    double ac = 27.0 / 28.0;
    r->S[0 * n + k] = ac * G * A0; r->S[1 * n + k] = ac * G * A0; r->S[2 * n + k] = E * A0;
    Bel[0 * n + k] = E * I1; Bel[1 * n + k] = E * I2; Bel[2 * n + k] = G * I3;
    r->volume[k] = PI * (rad * rad) * rl;
  }
  for (int k = 0; k < nv; k++) {
    r->rest_vor[k] = 0.5 * (r->rest_len[k + 1] + r->rest_len[k]);
    for (int i = 0; i < 3; i++)
      r->B[i * nv + k] = (Bel[i * n + k + 1] * r->rest_len[k + 1] + Bel[i * n + k] * r->rest_len[k]) /
                         (r->rest_len[k + 1] + r->rest_len[k]);
  }
  free(Bel);
  for (int k = 0; k < n; k++) {
    r->mass[k] += 0.5 * cfg->density * r->volume[k];
    r->mass[k + 1] += 0.5 * cfg->density * r->volume[k];
  }
  compute_shear_stretch_strains(r);
  compute_bending_twist_strains(r);
  /* finalize: constraint copies (B-7), damper coefficients (A.4) */
  for (int i = 0; i < 3; i++) r->fixed_pos[i] = X(i, 0);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) r->fixed_Q[i * 3 + j] = QQ(i, j, 0);
  if (cfg->damping_constant >= 0.0) {
    r->c_v = exp(-cfg->damping_constant * cfg->dt);
    for (int k = 0; k < n; k++) {
      double em = 0.5 * (r->mass[k + 1] + r->mass[k]);
      if (k == 0) em += 0.5 * r->mass[0];
      if (k == n - 1) em += 0.5 * r->mass[n];
      for (int i = 0; i < 3; i++)
        r->c_w[i * n + k] = exp(-cfg->damping_constant * cfg->dt * em * r->Jinv[i * n + k]);
    }
  }
  return r;
}

void ro_destroy(ro_rod *r) {
  if (!r) return;
  double *ptrs[] = {r->x, r->v, r->Q, r->w, r->acc, r->alpha, r->rest_len, r->rest_vor, r->mass,
                    r->volume, r->radius, r->J, r->Jinv, r->S, r->B, r->rest_sigma, r->rest_kappa,
                    r->len, r->tang, r->dil, r->vdil, r->dil_rate, r->sigma, r->kappa, r->stress,
                    r->couple, r->f_int, r->t_int, r->f_ext, r->t_ext, r->f_user, r->t_user, r->c_w, r->filt, r->tmp, r->ctmp, r->muscle, r->spl_pts, r->spl_mag, r->rest_radius, r->ml_act, r->ml_tmp};
  for (size_t i = 0; i < sizeof(ptrs) / sizeof(ptrs[0]); i++) free(ptrs[i]);
  free(r);
}

double ro_time(const ro_rod *r) { return r->time; }
int ro_n_elem(const ro_rod *r) { return r->n; }
double *ro_position(ro_rod *r) { return r->x; }
double *ro_velocity(ro_rod *r) { return r->v; }
double *ro_director(ro_rod *r) { return r->Q; }
double *ro_omega(ro_rod *r) { return r->w; }
double *ro_tangents(ro_rod *r) { return r->tang; }
double *ro_kappa(ro_rod *r) { return r->kappa; }
double *ro_sigma(ro_rod *r) { return r->sigma; }
double *ro_dilatation(ro_rod *r) { return r->dil; }
double *ro_rest_kappa(ro_rod *r) { return r->rest_kappa; }
double *ro_external_forces(ro_rod *r) { return r->f_user; }
double *ro_external_torques(ro_rod *r) { return r->t_user; }
void ro_set_sucker(ro_rod *r, int slot, int index, double ratio) {
  if (slot < 0 || slot >= 8) return;
  if (slot >= r->n_sucker) r->n_sucker = slot + 1;
  /* python indexing on arrays of different length: velocity_collection[..., -1] is the LAST NODE (n),
   * omega_collection[..., -1] the last element (n - 1) (controllable_constraint.py:68-69, arm_push_env.py:261) */
  r->sucker_idx[slot] = index < 0 ? index + r->n : index;
  r->sucker_node[slot] = index < 0 ? index + r->n + 1 : index;
  r->sucker_ratio[slot] = ratio;
}
double *ro_mass(ro_rod *r) { return r->mass; }
double *ro_internal_forces(ro_rod *r) { return r->f_int; }
double *ro_internal_torques(ro_rod *r) { return r->t_int; }
double *ro_radius(ro_rod *r) { return r->radius; }
double *ro_muscle(ro_rod *r) { return r->muscle; }
void ro_set_tm_muscle(ro_rod *r, double max_stress, double radius_ref) {
  r->tm_on = max_stress != 0.0; r->tm_max_stress = max_stress; r->tm_radius_ref = radius_ref;
}
void ro_set_tm_activation(ro_rod *r, double activation) { r->tm_activation = activation; }
void ro_set_muscle_layer(ro_rod *r, int slot, int kind, double max_stress, double radius_ref, double px, double py) {
  if (slot < 0 || slot >= 3) return;
  r->ml_kind[slot] = kind; r->ml_stress[slot] = max_stress; r->ml_rref[slot] = radius_ref; r->ml_px[slot] = px; r->ml_py[slot] = py;
}
double *ro_muscle_activation(ro_rod *r) { return r->ml_act; }
double *ro_spline_points(ro_rod *r) { return r->spl_pts; }
double *ro_spline_magnitude(ro_rod *r) { return r->spl_mag; }

/* numpy pairwise summation (np.add.reduce on a contiguous float64 row, n < 128 block) */
static double np_pairwise_sum(const double *a, int n) {
  if (n < 8) {
    double res = 0.0;
    for (int i = 0; i < n; i++) res += a[i];
    return res;
  } else if (n <= 128) {
    double rr[8];
    for (int j = 0; j < 8; j++) rr[j] = a[j];
    int i;
    for (i = 8; i < n - (n % 8); i += 8)
      for (int j = 0; j < 8; j++) rr[j] += a[i + j];
    double res = ((rr[0] + rr[1]) + (rr[2] + rr[3])) + ((rr[4] + rr[5]) + (rr[6] + rr[7]));
    for (; i < n; i++) res += a[i];
    return res;
  } else {
    int n2 = n / 2;
    n2 -= n2 % 8;
    return np_pairwise_sum(a, n2) + np_pairwise_sum(a + n2, n - n2);
  }
}

static double np_mod(double a, double b) { /* numpy float remainder (sign of divisor) */
  double m = fmod(a, b);
  if (m != 0.0) { if ((b < 0) != (m < 0)) m += b; }
  else m = copysign(0.0, b);
  return m;
}

static double softpendulum_theta(ro_rod *r) { /* soft_pendulum.py:154-156 */
  const int n = r->n;
  double mx = np_pairwise_sum(r->tang + 0 * n, n) / (double)n;
  double my = np_pairwise_sum(r->tang + 1 * n, n) / (double)n;
  double theta = atan(mx / my);
  return np_mod(theta + PI, 2 * PI) - PI;
}

void ro_softpendulum_obs(ro_rod *r, float prev_action, float obs[4]) {
  const int n = r->n;
  obs[0] = (float)X(0, 0);
  obs[1] = (float)V(0, 0);
  obs[2] = prev_action;
  obs[3] = (float)softpendulum_theta(r);
}

void ro_softpendulum_step(ro_rod *r, float action, int step_skip, double final_time, float obs[4],
                          double *reward, int *terminated, int *truncated) {
  const int n = r->n;
  ro_substeps(r, step_skip, (double)action, NULL, NULL);
  int invalid = 0;
  for (int i = 0; i < 3 * (n + 1); i++) invalid |= isnan(r->x[i]) || isnan(r->v[i]);
  double survive = 0.0, forward = 0.0;
  *terminated = 0;
  if (invalid) { *terminated = 1; survive = -50.0; }
  else {
    double d = fabs(X(0, 0));
    double th = softpendulum_theta(r);
    forward = d * 10 + th * th;
  }
  *truncated = r->time > final_time;
  *reward = forward - 0.0 + survive;
  ro_softpendulum_obs(r, action, obs);
}

int ro_max_threads(void) {
  long c = sysconf(_SC_NPROCESSORS_ONLN);
  return c > 0 ? (int)c : 1;
}

typedef struct {
  ro_rod **rods; const float *actions; int step_skip; double final_time;
  float *obs; double *reward; int *terminated, *truncated; int lo, hi;
} batch_job;

static void *batch_worker(void *p) {
  batch_job *j = (batch_job *)p;
  for (int e = j->lo; e < j->hi; e++)
    ro_softpendulum_step(j->rods[e], j->actions[e], j->step_skip, j->final_time, j->obs + 4 * e,
                         j->reward + e, j->terminated + e, j->truncated + e);
  return NULL;
}

/* contiguous env ranges, one POSIX thread each (the image has no libgomp) */
void ro_softpendulum_step_batch(ro_rod **rods, int n_env, const float *actions, int step_skip,
                                double final_time, float *obs, double *reward, int *terminated,
                                int *truncated, int n_threads) {
  if (n_threads <= 0) n_threads = ro_max_threads();
  if (n_threads > n_env) n_threads = n_env;
  if (n_threads < 1) n_threads = 1;
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
  batch_job *jobs = (batch_job *)malloc(sizeof(batch_job) * (size_t)n_threads);
  for (int t = 0; t < n_threads; t++) {
    batch_job j = {rods, actions, step_skip, final_time, obs, reward, terminated, truncated,
                   (int)((long)n_env * t / n_threads), (int)((long)n_env * (t + 1) / n_threads)};
    jobs[t] = j;
    if (t > 0) pthread_create(&th[t], NULL, batch_worker, &jobs[t]);
  }
  batch_worker(&jobs[0]);
  for (int t = 1; t < n_threads; t++) pthread_join(th[t], NULL);
  free(th); free(jobs);
}

/* ==== multi-system assembly: n_arm rods + rigid Cylinder head + FixedJoint2Rigid joints ===================
 * Restates what PositionVerlet does over the systems assembled by
 *   /root/reference/gym_softrobot/envs/octopus/build.py:52-217 (build_octopus)
 * with the gym-softrobot plugins
 *   /root/reference/gym_softrobot/utils/custom_elastica/joint.py:48-123 (forces), :125-219 (torques)
 *   /root/reference/gym_softrobot/utils/custom_elastica/constraint.py:43-58 (values), :62-85 (rates)
 * and PyElastica's Cylinder (same recalled inertia as oracle/shims/elastica/rigidbody.py).
 * Operator order = the build code's call order (OperatorGroupFIFO): synchronize = [joint a: forces, torques
 * (a = 0..n_arm-1), gravity per arm, contact per arm]; constrain_rates = [head BC, dampers per arm].
 * Systems are stepped arms first, head last (append order). */
struct ro_assembly {
  ro_asm_config cfg;
  int n_arm;
  ro_rod *arm[RO_MAX_ARMS];
  double time;
  /* head: one node, one "element" */
  double hx[3], hv[3], hQ[9], hw[3], hF[3], hT[3];
  double h_mass, hJ[3], hJinv[3];
  double h_fixed_pos[3];
  int head_fixed; double h_fix_x[3], h_fix_Q[9];   /* OneEndFixedBC on the head (reach_env.py:128-132) */
};

static void head_kinematic(ro_assembly *a, double prefac) {
  for (int i = 0; i < 3; i++) a->hx[i] += prefac * a->hv[i];
  double R[3][3], Qn[9];
  rotation_matrix(prefac * a->hw[0], prefac * a->hw[1], prefac * a->hw[2], R);
  for (int i = 0; i < 3; i++)
    for (int m = 0; m < 3; m++) {
      double s = 0.0;
      for (int j = 0; j < 3; j++) s += R[i][j] * a->hQ[j * 3 + m];
      Qn[i * 3 + m] = s;
    }
  memcpy(a->hQ, Qn, sizeof(Qn));
}

static void head_constrain_values(ro_assembly *a) { /* constraint.py:43-58 */
  a->hx[2] = a->h_fixed_pos[2];
  a->hQ[6] = 0.0; a->hQ[7] = 0.0; a->hQ[8] = 1.0;
  for (int i = 0; i < 2; i++) {
    double len = sqrt(a->hQ[3 * i] * a->hQ[3 * i] + a->hQ[3 * i + 1] * a->hQ[3 * i + 1]);
    for (int j = 0; j < 2; j++) a->hQ[3 * i + j] /= len;
    a->hQ[3 * i + 2] = 0.0;
  }
}

static void apply_joint(ro_assembly *a, int ai) {
  ro_rod *r = a->arm[ai];
  const int n = r->n;
  const ro_asm_config *c = &a->cfg;
  /* apply_forces (joint.py:48-68): rigid_rod_pos = head position with z := 0 */
  double pos[3] = {a->hx[0], a->hx[1], 0.0};
  /* z_rotation(binormal, angle) (joint.py:7-17): theta = angle / 180 * pi */
  const double th = c->joint_angle_deg[ai] / 180.0 * PI;
  const double cs = cos(th), sn = sin(th);
  const double b[3] = {a->hQ[3], a->hQ[4], a->hQ[5]};   /* director_collection[1, :, -1] */
  double dir[3] = {-(cs * b[0] + -sn * b[1] + 0.0 * b[2]), -(sn * b[0] + cs * b[1] + 0.0 * b[2]),
                   -(0.0 * b[0] + 0.0 * b[1] + 1.0 * b[2])};
  for (int i = 0; i < 3; i++) pos[i] += dir[i] * c->joint_radius;   /* rigid_rod_pos += connection point (in place) */
  double d[3] = {X(0, 0) - pos[0], X(1, 0) - pos[1], X(2, 0) - pos[2]};
  double dist = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
  double nh[3] = {0.0, 0.0, 0.0};
  if (!(dist <= 2.220446049250313e-16 * 1e4)) { nh[0] = d[0] / dist; nh[1] = d[1] / dist; nh[2] = d[2] / dist; }
  double rel[3] = {V(0, 0) - a->hv[0], V(1, 0) - a->hv[1], V(2, 0) - a->hv[2]};
  double rn = rel[0] * nh[0] + rel[1] * nh[1] + rel[2] * nh[2];
  for (int i = 0; i < 3; i++) {
    double cf = c->joint_k * d[i] + (-c->joint_nu * (rn * nh[i]));
    a->hF[i] += cf;
    r->f_ext[i * (n + 1) + 0] -= cf;
  }
  /* apply_torques (joint.py:125-219) with the mutated rigid_rod_pos */
  double link[3] = {X(0, 1) - X(0, 0), X(1, 1) - X(1, 0), X(2, 1) - X(2, 0)};
  double fd[3];
  for (int i = 0; i < 3; i++) {
    double tgt = pos[i] + r->rest_len[0] * dir[i];
    fd[i] = -c->joint_kt * (X(i, 1) - tgt);
  }
  double tq[3] = {link[1] * fd[2] - link[2] * fd[1], link[2] * fd[0] - link[0] * fd[2], link[0] * fd[1] - link[1] * fd[0]};
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      a->hT[i] -= a->hQ[3 * i + j] * tq[j];
      r->t_ext[i * n + 0] += QQ(i, j, 0) * tq[j];
    }
}

static void head_fix_values(ro_assembly *a) {   /* OneEndFixedBC.constrain_values: position / directors of "node 0" back to their initial values */
  if (!a->head_fixed) return;
  for (int i = 0; i < 3; i++) a->hx[i] = a->h_fix_x[i];
  for (int i = 0; i < 9; i++) a->hQ[i] = a->h_fix_Q[i];
}

static void asm_substep(ro_assembly *a) {
  const double dt = a->cfg.dt, prefac = 0.5 * dt;
  for (int i = 0; i < a->n_arm; i++) phase_first_half(a->arm[i], NULL);   /* arms carry no BC of their own */
  head_kinematic(a, prefac);
  a->time += prefac;
  if (a->cfg.has_head) head_constrain_values(a);
  head_fix_values(a);
  for (int i = 0; i < a->n_arm; i++) compute_internal_forces_and_torques(a->arm[i]);
  if (a->cfg.has_head)
    for (int i = 0; i < a->n_arm; i++) apply_joint(a, i);
  for (int i = 0; i < a->n_arm; i++) phase_forcing(a->arm[i], 0.0);
  for (int i = 0; i < a->n_arm; i++)
    if (a->arm[i]->cfg.contact_on) apply_contact(a->arm[i]);
  for (int i = 0; i < a->n_arm; i++) phase_dynamic(a->arm[i]);
  {   /* rigid body: a = F/m ; alpha = J^-1 ((J w) x w + T) */
    double Jw[3] = {a->hJ[0] * a->hw[0], a->hJ[1] * a->hw[1], a->hJ[2] * a->hw[2]};
    double lt[3] = {Jw[1] * a->hw[2] - Jw[2] * a->hw[1], Jw[2] * a->hw[0] - Jw[0] * a->hw[2], Jw[0] * a->hw[1] - Jw[1] * a->hw[0]};
    for (int i = 0; i < 3; i++) {
      double acc = a->hF[i] / a->h_mass, alpha = a->hJinv[i] * (lt[i] + a->hT[i]);
      a->hv[i] += dt * acc;
      a->hw[i] += dt * alpha;
    }
  }
  if (a->cfg.has_head) { a->hv[2] = 0.0; a->hw[0] = 0.0; a->hw[1] = 0.0; }   /* constraint.py:62-85 */
  if (a->head_fixed) for (int i = 0; i < 3; i++) { a->hv[i] = 0.0; a->hw[i] = 0.0; }   /* OneEndFixedBC.constrain_rates, registered after */
  for (int i = 0; i < a->n_arm; i++) phase_rates(a->arm[i], NULL);
  for (int i = 0; i < a->n_arm; i++) phase_second_half(a->arm[i], NULL);
  head_kinematic(a, prefac);
  a->time += prefac;
  if (a->cfg.has_head) head_constrain_values(a);
  head_fix_values(a);
  for (int i = 0; i < 3; i++) { a->hF[i] = 0.0; a->hT[i] = 0.0; }
}

ro_assembly *ro_asm_create(const ro_config *arm_cfgs, const ro_asm_config *cfg) {
  if (cfg->n_arm < 1 || cfg->n_arm > RO_MAX_ARMS) return NULL;
  ro_assembly *a = (ro_assembly *)calloc(1, sizeof(ro_assembly));
  a->cfg = *cfg;
  a->n_arm = cfg->n_arm;
  for (int i = 0; i < a->n_arm; i++) a->arm[i] = ro_create(&arm_cfgs[i]);
  /* Cylinder(start, direction, normal, length, radius, density): centre of mass, rows normal / binormal / direction */
  const double *s = cfg->head_start, *d = cfg->head_direction, *nm = cfg->head_normal;
  for (int i = 0; i < 3; i++) a->hx[i] = s[i] + d[i] * cfg->head_length / 2;
  for (int j = 0; j < 3; j++) { a->hQ[j] = nm[j]; a->hQ[6 + j] = d[j]; }
  a->hQ[3] = d[1] * nm[2] - d[2] * nm[1]; a->hQ[4] = d[2] * nm[0] - d[0] * nm[2]; a->hQ[5] = d[0] * nm[1] - d[1] * nm[0];
  const double A0 = PI * cfg->head_radius * cfg->head_radius, I1 = A0 * A0 / (4.0 * PI);
  const double I0[3] = {I1, I1, 2.0 * I1};
  a->h_mass = PI * cfg->head_radius * cfg->head_radius * cfg->head_length * cfg->head_density;
  for (int i = 0; i < 3; i++) { a->hJ[i] = I0[i] * cfg->head_density * cfg->head_length; a->hJinv[i] = 1.0 / a->hJ[i]; }
  for (int i = 0; i < 3; i++) a->h_fixed_pos[i] = a->hx[i];   /* B-7: copy at finalize */
  if (cfg->has_head) { head_constrain_values(a); a->hv[2] = 0.0; a->hw[0] = 0.0; a->hw[1] = 0.0; }
  return a;
}

void ro_asm_destroy(ro_assembly *a) {
  if (!a) return;
  for (int i = 0; i < a->n_arm; i++) ro_destroy(a->arm[i]);
  free(a);
}
void ro_asm_set_head_fixed(ro_assembly *a, int on) {
  a->head_fixed = on;
  for (int i = 0; i < 3; i++) { a->h_fix_x[i] = a->hx[i]; if (on) { a->hv[i] = 0.0; a->hw[i] = 0.0; } }
  for (int i = 0; i < 9; i++) a->h_fix_Q[i] = a->hQ[i];
}
void ro_asm_substeps(ro_assembly *a, int n_substeps) { for (int s = 0; s < n_substeps; s++) asm_substep(a); }
ro_rod *ro_asm_arm(ro_assembly *a, int i) { return (i >= 0 && i < a->n_arm) ? a->arm[i] : NULL; }
double ro_asm_time(const ro_assembly *a) { return a->time; }
double *ro_asm_head_position(ro_assembly *a) { return a->hx; }
double *ro_asm_head_velocity(ro_assembly *a) { return a->hv; }
double *ro_asm_head_director(ro_assembly *a) { return a->hQ; }
double *ro_asm_head_omega(ro_assembly *a) { return a->hw; }

/* generic batch runners for the CPU baselines of the other workloads (bench.py --config 3/4/5) */
typedef struct { ro_rod **rods; ro_assembly **asms; int n_substeps, lo, hi; } generic_job;

static void *generic_worker(void *p) {
  generic_job *j = (generic_job *)p;
  for (int e = j->lo; e < j->hi; e++) {
    if (j->rods) ro_substeps(j->rods[e], j->n_substeps, 0.0, NULL, NULL);
    else ro_asm_substeps(j->asms[e], j->n_substeps);
  }
  return NULL;
}

static void generic_batch(ro_rod **rods, ro_assembly **asms, int n, int n_substeps, int n_threads) {
  if (n_threads <= 0) n_threads = ro_max_threads();
  if (n_threads > n) n_threads = n;
  if (n_threads < 1) n_threads = 1;
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
  generic_job *jobs = (generic_job *)malloc(sizeof(generic_job) * (size_t)n_threads);
  for (int t = 0; t < n_threads; t++) {
    generic_job j = {rods, asms, n_substeps, (int)((long)n * t / n_threads), (int)((long)n * (t + 1) / n_threads)};
    jobs[t] = j;
    if (t > 0) pthread_create(&th[t], NULL, generic_worker, &jobs[t]);
  }
  generic_worker(&jobs[0]);
  for (int t = 1; t < n_threads; t++) pthread_join(th[t], NULL);
  free(th); free(jobs);
}

void ro_substeps_batch(ro_rod **rods, int n, int n_substeps, int n_threads) { generic_batch(rods, NULL, n, n_substeps, n_threads); }
void ro_asm_substeps_batch(ro_assembly **asms, int n, int n_substeps, int n_threads) { generic_batch(NULL, asms, n, n_substeps, n_threads); }
