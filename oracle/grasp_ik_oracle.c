/* CPU ORACLE (plain C, fp64) -- TEST INFRASTRUCTURE ONLY, never on the product path.
 *
 * Restatement of the reference's dual-arm grasp IK loop (/root/reference/inverse_geometry.py:17-100) and of the
 * pinocchio / numpy primitives it calls, table-driven over a generic tree of 1-dof revolute joints (what
 * pinocchio's Model is for this robot).  It mirrors oracle/grasp_ik_np.py function for function and exists so
 * that (a) parity tests can check thousands of problems in seconds and (b) bench.py has a compiled CPU
 * baseline (OpenMP over problems) -- bench.py's `cpu_baseline` / `--impl reference` legs and tests/ are the
 * only callers.
 *
 * Parity status: PINNED -- reproduces the reference's golden outputs (trajectory.json q_control_points[0]/[-1])
 * to ~1e-14 in 740 / 736 iterations and agrees with the numpy oracle (tests/test_oracle.py).
 *
 * Dependencies restated: pinocchio (PyPI `pin`, version unpinned by requirements.txt:1; 2.x per the
 * notebooks) for FK / frame Jacobians / log6, numpy.linalg.pinv (LAPACK gesdd SVD, rcond = 1e-15) restated as a
 * one-sided Jacobi SVD with the same cutoff rule.
 *
 * Build: see oracle/Makefile  (gcc -O3 -march=native -fopenmp -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAX_NQ 32

/* Same field order as gik_table_t (include/gik.h) so tests can pass one ctypes struct to both. */
typedef struct {
  int32_t nq;
  int32_t parent[ORC_MAX_NQ];
  int32_t axis[ORC_MAX_NQ];
  double joint_R[ORC_MAX_NQ][9];
  double joint_p[ORC_MAX_NQ][3];
  double lower[ORC_MAX_NQ];
  double upper[ORC_MAX_NQ];
  int32_t hand_joint[2];
  double hand_R[2][9];
  double hand_p[2][3];
  double hook_R[2][9];
  double hook_p[2][3];
} orc_table_t;

typedef struct { double R[9]; double p[3]; } se3_t;

static void mat_mul(const double* A, const double* B, double* C) { /* 3x3 row-major */
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c)
      C[3 * r + c] = A[3 * r] * B[c] + A[3 * r + 1] * B[3 + c] + A[3 * r + 2] * B[6 + c];
}
static void mat_vec(const double* A, const double* x, double* y) {
  for (int r = 0; r < 3; ++r) y[r] = A[3 * r] * x[0] + A[3 * r + 1] * x[1] + A[3 * r + 2] * x[2];
}
static void mat_tvec(const double* A, const double* x, double* y) { /* A^T x */
  for (int r = 0; r < 3; ++r) y[r] = A[r] * x[0] + A[3 + r] * x[1] + A[6 + r] * x[2];
}
static void se3_mul(const se3_t* A, const se3_t* B, se3_t* C) {
  double t[3];
  mat_mul(A->R, B->R, C->R);
  mat_vec(A->R, B->p, t);
  for (int i = 0; i < 3; ++i) C->p[i] = A->p[i] + t[i];
}
static void rot_axis(int axis, double a, double* R) { /* pinocchio JointModelRX/RY/RZ */
  const double c = cos(a), s = sin(a);
  memset(R, 0, 9 * sizeof(double));
  if (axis == 0) { R[0] = 1; R[4] = c; R[5] = -s; R[7] = s; R[8] = c; }
  else if (axis == 1) { R[0] = c; R[2] = s; R[4] = 1; R[6] = -s; R[8] = c; }
  else { R[0] = c; R[1] = -s; R[3] = s; R[4] = c; R[8] = 1; }
}

/* pin.framesForwardKinematics (inverse_geometry.py:58): oMi[i] = oMi[parent] * jointPlacement[i] * Rot(axis, q_i) */
static void forward_kinematics(const orc_table_t* t, const double* q, se3_t* oMi) {
  for (int i = 0; i < t->nq; ++i) {
    se3_t pl, jr, tmp;
    memcpy(pl.R, t->joint_R[i], sizeof pl.R);
    memcpy(pl.p, t->joint_p[i], sizeof pl.p);
    rot_axis(t->axis[i], q[i], jr.R);
    jr.p[0] = jr.p[1] = jr.p[2] = 0.0;
    if (t->parent[i] < 0) tmp = pl; else se3_mul(&oMi[t->parent[i]], &pl, &tmp);
    se3_mul(&tmp, &jr, &oMi[i]);
  }
}

/* data.oMf[getFrameId(hand)] (inverse_geometry.py:62-63) */
static void hand_placement(const orc_table_t* t, const se3_t* oMi, int h, se3_t* out) {
  se3_t f;
  memcpy(f.R, t->hand_R[h], sizeof f.R);
  memcpy(f.p, t->hand_p[h], sizeof f.p);
  se3_mul(&oMi[t->hand_joint[h]], &f, out);
}

/* pin.computeFrameJacobian, LOCAL frame (inverse_geometry.py:75-76): J is 6 x nq row-major, rows [lin; ang] */
static void frame_jacobian_local(const orc_table_t* t, const se3_t* oMi, const se3_t* oMf, int h, double* J) {
  const int nq = t->nq;
  memset(J, 0, sizeof(double) * 6 * nq);
  for (int k = t->hand_joint[h]; k >= 0; k = t->parent[k]) {
    const int ax = t->axis[k];
    const double a[3] = {oMi[k].R[ax], oMi[k].R[3 + ax], oMi[k].R[6 + ax]};
    const double d[3] = {oMf->p[0] - oMi[k].p[0], oMf->p[1] - oMi[k].p[1], oMf->p[2] - oMi[k].p[2]};
    const double cr[3] = {a[1] * d[2] - a[2] * d[1], a[2] * d[0] - a[0] * d[2], a[0] * d[1] - a[1] * d[0]};
    double lin[3], ang[3];
    mat_tvec(oMf->R, cr, lin);
    mat_tvec(oMf->R, a, ang);
    for (int r = 0; r < 3; ++r) { J[r * nq + k] = lin[r]; J[(3 + r) * nq + k] = ang[r]; }
  }
}

/* pinocchio log3 / log6 (pin.log, inverse_geometry.py:66-67); out = [v; w] */
static const double TS_PREC3 = 1.220703125e-4; /* eps^(1/4) = 2^-13 */
static double log3(const double* R, double* w) {
  double tr = R[0] + R[4] + R[8], theta;
  const double PI = 3.14159265358979323846;
  if (tr >= 3.0) { tr = 3.0; theta = 0.0; }
  else if (tr <= -1.0) { tr = -1.0; theta = PI; }
  else theta = acos((tr - 1.0) / 2.0);
  if (theta >= PI - 1e-2) {
    const double cphi = -(tr - 1.0) / 2.0, beta = theta * theta / (1.0 + cphi);
    const double t0 = (R[0] + cphi) * beta, t1 = (R[4] + cphi) * beta, t2 = (R[8] + cphi) * beta;
    w[0] = (R[7] > R[5] ? 1.0 : -1.0) * (t0 > 0 ? sqrt(t0) : 0.0);
    w[1] = (R[2] > R[6] ? 1.0 : -1.0) * (t1 > 0 ? sqrt(t1) : 0.0);
    w[2] = (R[3] > R[1] ? 1.0 : -1.0) * (t2 > 0 ? sqrt(t2) : 0.0);
  } else {
    const double f = (theta > TS_PREC3 ? theta / sin(theta) : 1.0) / 2.0;
    w[0] = f * (R[7] - R[5]); w[1] = f * (R[2] - R[6]); w[2] = f * (R[3] - R[1]);
  }
  return theta;
}
static void log6(const se3_t* M, double* out) {
  double w[3], alpha, beta;
  const double t = log3(M->R, w), t2 = t * t;
  if (t < TS_PREC3) {
    alpha = 1.0 - t2 / 12.0 - t2 * t2 / 720.0;
    beta = 1.0 / 12.0 + t2 / 720.0;
  } else {
    const double st = sin(t), ct = cos(t);
    alpha = t * st / (2.0 * (1.0 - ct));
    beta = 1.0 / t2 - st / (2.0 * t * (1.0 - ct));
  }
  const double* p = M->p;
  const double wp = beta * (w[0] * p[0] + w[1] * p[1] + w[2] * p[2]);
  out[0] = alpha * p[0] - 0.5 * (w[1] * p[2] - w[2] * p[1]) + wp * w[0];
  out[1] = alpha * p[1] - 0.5 * (w[2] * p[0] - w[0] * p[2]) + wp * w[1];
  out[2] = alpha * p[2] - 0.5 * (w[0] * p[1] - w[1] * p[0]) + wp * w[2];
  out[3] = w[0]; out[4] = w[1]; out[5] = w[2];
}

/* pin.log(oMhand.inverse() * oMcube).vector */
static void hand_error(const se3_t* hand, const se3_t* target, double* e) {
  se3_t M;
  double d[3] = {target->p[0] - hand->p[0], target->p[1] - hand->p[1], target->p[2] - hand->p[2]};
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c)
      M.R[3 * r + c] = hand->R[r] * target->R[c] + hand->R[3 + r] * target->R[3 + c] + hand->R[6 + r] * target->R[6 + c];
  mat_tvec(hand->R, d, M.p);
  log6(&M, e);
}

/* vq = pinv(J) e with numpy.linalg.pinv semantics (rcond = 1e-15 relative to the largest singular value;
 * inverse_geometry.py:83).  J is m x n row-major (m = 12, n = nq).  One-sided Jacobi SVD of W = J^T (n x m):
 * W = U S V^T  =>  pinv(J) = U S^+ V^T.
 * lambda > 0 (not a reference setting -- the reference is undamped; BASELINE config 3 asks for damping on random
 * restarts): the damped least-squares step J^T (J J^T + lambda I)^-1 e = sum_j u_j s_j / (s_j^2 + lambda) (v_j . e),
 * from the same SVD, no cutoff (every term is finite). */
static void pinv_apply(const double* J, int m, int n, const double* e, double lambda, double* vq) {
  double W[ORC_MAX_NQ][12], V[12][12];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < m; ++j) W[i][j] = J[j * n + i];
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < m; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0;
    for (int p = 0; p < m - 1; ++p)
      for (int q = p + 1; q < m; ++q) {
        double a = 0, b = 0, c = 0;
        for (int i = 0; i < n; ++i) { a += W[i][p] * W[i][p]; b += W[i][q] * W[i][q]; c += W[i][p] * W[i][q]; }
        if (c == 0.0 || fabs(c) <= 1e-300) continue;
        const double rel = fabs(c) / sqrt(a * b > 0 ? a * b : 1e-300);
        if (rel > off) off = rel;
        if (rel < 1e-17) continue;
        const double zeta = (b - a) / (2.0 * c);
        const double tt = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double cs = 1.0 / sqrt(1.0 + tt * tt), sn = cs * tt;
        for (int i = 0; i < n; ++i) {
          const double wp = W[i][p], wq = W[i][q];
          W[i][p] = cs * wp - sn * wq; W[i][q] = sn * wp + cs * wq;
        }
        for (int i = 0; i < m; ++i) {
          const double vp = V[i][p], vqq = V[i][q];
          V[i][p] = cs * vp - sn * vqq; V[i][q] = sn * vp + cs * vqq;
        }
      }
    if (off < 1e-16) break;
  }
  double sig[12], smax = 0.0;
  for (int j = 0; j < m; ++j) {
    double s = 0;
    for (int i = 0; i < n; ++i) s += W[i][j] * W[i][j];
    sig[j] = sqrt(s);
    if (sig[j] > smax) smax = sig[j];
  }
  for (int i = 0; i < n; ++i) vq[i] = 0.0;
  for (int j = 0; j < m; ++j) {
    if (lambda <= 0.0 && !(sig[j] > 1e-15 * smax)) continue;
    double ve = 0;
    for (int i = 0; i < m; ++i) ve += V[i][j] * e[i];
    const double f = ve / (sig[j] * sig[j] + (lambda > 0.0 ? lambda : 0.0)); /* U[:,j] = W[:,j] / sig */
    for (int i = 0; i < n; ++i) vq[i] += W[i][j] * f;
  }
}

/* STRONG CPU BASELINE step (bench.py cpu_baseline.strong only; parity tests always use pinv_apply): the same
 * vq = J^T (J J^T + lambda I)^-1 e through a dense 12 x 12 Cholesky factorisation of the normal equations instead of an
 * SVD -- what a performance-minded CPU implementation of the reference's step would do.  Equal to pinv(J) e wherever J
 * has full row rank; pivots are floored.  ~15x fewer flops than the Jacobi SVD above. */
static void normal_eq_apply(const double* J, int m, int n, const double* e, double lambda, double* vq) {
  double G[12][12], z[12];
  for (int i = 0; i < m; ++i)
    for (int j = 0; j <= i; ++j) {
      double s = (i == j) ? lambda : 0.0;
      for (int k = 0; k < n; ++k) s += J[i * n + k] * J[j * n + k];
      G[i][j] = s;
    }
  for (int j = 0; j < m; ++j) {               /* Cholesky, lower */
    double d = G[j][j];
    for (int k = 0; k < j; ++k) d -= G[j][k] * G[j][k];
    d = sqrt(d > 1e-300 ? d : 1e-300);
    G[j][j] = d;
    for (int i = j + 1; i < m; ++i) {
      double v = G[i][j];
      for (int k = 0; k < j; ++k) v -= G[i][k] * G[j][k];
      G[i][j] = v / d;
    }
  }
  for (int i = 0; i < m; ++i) {               /* L y = e */
    double v = e[i];
    for (int k = 0; k < i; ++k) v -= G[i][k] * z[k];
    z[i] = v / G[i][i];
  }
  for (int i = m - 1; i >= 0; --i) {          /* L^T z = y */
    double v = z[i];
    for (int k = i + 1; k < m; ++k) v -= G[k][i] * z[k];
    z[i] = v / G[i][i];
  }
  for (int k = 0; k < n; ++k) {
    double v = 0.0;
    for (int i = 0; i < m; ++i) v += J[i * n + k] * z[i];
    vq[k] = v;
  }
}
static int g_step_method = 0; /* 0 = pinv via SVD (the restatement), 1 = normal equations (strong baseline timing) */
void orc_set_step_method(int m) { g_step_method = m; }

static void hook_targets(const orc_table_t* t, const double* pose12, se3_t* tg) {
  se3_t cube;
  memcpy(cube.R, pose12, sizeof cube.R);
  memcpy(cube.p, pose12 + 9, sizeof cube.p);
  for (int h = 0; h < 2; ++h) {
    se3_t hk;
    memcpy(hk.R, t->hook_R[h], sizeof hk.R);
    memcpy(hk.p, t->hook_p[h], sizeof hk.p);
    se3_mul(&cube, &hk, &tg[h]); /* tools.getcubeplacement, tools.py:54-59 */
  }
}

/* inverse_geometry.computeqgrasppose (inverse_geometry.py:17-100) without the collision term; loop order kept:
 * residual test BEFORE the update, post-update q returned on exhaustion. */
static int solve_one(const orc_table_t* t, const double* q0, const double* pose12, double eps, double dt,
                     int max_iters, double damping, double* q, int* iters_out, double* resid_out) {
  const int nq = t->nq;
  se3_t tg[2], oMi[ORC_MAX_NQ], hand[2];
  double e[12], J[12 * ORC_MAX_NQ], vq[ORC_MAX_NQ];
  hook_targets(t, pose12, tg);
  memcpy(q, q0, sizeof(double) * nq);
  int success = 0, it;
  double nL = NAN, nR = NAN;
  for (it = 0; it < max_iters; ++it) {
    forward_kinematics(t, q, oMi);
    for (int h = 0; h < 2; ++h) { hand_placement(t, oMi, h, &hand[h]); hand_error(&hand[h], &tg[h], e + 6 * h); }
    nL = sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2] + e[3] * e[3] + e[4] * e[4] + e[5] * e[5]);
    nR = sqrt(e[6] * e[6] + e[7] * e[7] + e[8] * e[8] + e[9] * e[9] + e[10] * e[10] + e[11] * e[11]);
    if (nL < eps && nR < eps) { success = 1; break; }
    for (int h = 0; h < 2; ++h) frame_jacobian_local(t, oMi, &hand[h], h, J + 6 * nq * h);
    if (g_step_method == 1) normal_eq_apply(J, 12, nq, e, damping, vq);
    else pinv_apply(J, 12, nq, e, damping, vq);
    for (int i = 0; i < nq; ++i) {
      double v = q[i] + vq[i] * dt;                 /* pin.integrate (:86) */
      v = v < t->lower[i] ? t->lower[i] : v;         /* projecttojointlimits (:89) */
      q[i] = v > t->upper[i] ? t->upper[i] : v;
    }
  }
  if (!success) { /* residual at the returned configuration (not evaluated by the reference; informational) */
    forward_kinematics(t, q, oMi);
    for (int h = 0; h < 2; ++h) { hand_placement(t, oMi, h, &hand[h]); hand_error(&hand[h], &tg[h], e + 6 * h); }
    nL = sqrt(e[0] * e[0] + e[1] * e[1] + e[2] * e[2] + e[3] * e[3] + e[4] * e[4] + e[5] * e[5]);
    nR = sqrt(e[6] * e[6] + e[7] * e[7] + e[8] * e[8] + e[9] * e[9] + e[10] * e[10] + e[11] * e[11]);
  }
  *iters_out = it;
  resid_out[0] = nL; resid_out[1] = nR;
  return success;
}

/* ---------------------------------------------------------------------------------------------------------
 * exported entry points: ROW-MAJOR batches (q [n][nq], pose [n][12]), host pointers
 * ------------------------------------------------------------------------------------------------------- */
int orc_fk(const orc_table_t* t, int64_t n, const double* q, double* frames /* [n][2][12] */) {
  for (int64_t i = 0; i < n; ++i) {
    se3_t oMi[ORC_MAX_NQ], f;
    forward_kinematics(t, q + i * t->nq, oMi);
    for (int h = 0; h < 2; ++h) {
      hand_placement(t, oMi, h, &f);
      memcpy(frames + (i * 2 + h) * 12, f.R, sizeof f.R);
      memcpy(frames + (i * 2 + h) * 12 + 9, f.p, sizeof f.p);
    }
  }
  return 0;
}

int orc_jac(const orc_table_t* t, int64_t n, const double* q, double* jac /* [n][2][6][nq] */) {
  const int nq = t->nq;
  for (int64_t i = 0; i < n; ++i) {
    se3_t oMi[ORC_MAX_NQ], f;
    forward_kinematics(t, q + i * nq, oMi);
    for (int h = 0; h < 2; ++h) {
      hand_placement(t, oMi, h, &f);
      frame_jacobian_local(t, oMi, &f, h, jac + (i * 2 + h) * 6 * nq);
    }
  }
  return 0;
}

int orc_solve(const orc_table_t* t, int64_t n, const double* q_init, const double* pose, double eps, double dt,
              int max_iters, double damping, int threads, double* q_out, uint8_t* conv, int32_t* iters,
              double* resid /* [n][2] */) {
  const int nq = t->nq;
#ifdef _OPENMP
  omp_set_num_threads(threads > 0 ? threads : omp_get_num_procs());
#pragma omp parallel for schedule(dynamic, 1)
#endif
  for (int64_t i = 0; i < n; ++i) {
    int it;
    double r[2];
    conv[i] = (uint8_t)solve_one(t, q_init + i * nq, pose + i * 12, eps, dt, max_iters, damping, q_out + i * nq, &it, r);
    if (iters) iters[i] = it;
    if (resid) { resid[2 * i] = r[0]; resid[2 * i + 1] = r[1]; }
  }
  return 0;
}

/* SE3.Interpolate(A, B, alpha) = A * exp6(alpha * log6(A^-1 B))  (path.py:141) */
static void se3_interpolate(const double* A12, const double* B12, double alpha, double* out12) {
  se3_t A, B, D, E;
  memcpy(A.R, A12, sizeof A.R); memcpy(A.p, A12 + 9, sizeof A.p);
  memcpy(B.R, B12, sizeof B.R); memcpy(B.p, B12 + 9, sizeof B.p);
  double xi[6];
  hand_error(&A, &B, xi); /* log6(A^-1 B) */
  double v[3] = {alpha * xi[0], alpha * xi[1], alpha * xi[2]}, w[3] = {alpha * xi[3], alpha * xi[4], alpha * xi[5]};
  const double t2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2], t = sqrt(t2);
  double a, b, c; /* pinocchio exp6: sin t / t, (1 - cos t)/t^2, (t - sin t)/t^3 */
  if (t < TS_PREC3) { a = 1.0 - t2 / 6.0; b = 0.5 - t2 / 24.0; c = 1.0 / 6.0 - t2 / 120.0; }
  else { a = sin(t) / t; b = (1.0 - cos(t)) / t2; c = (t - sin(t)) / (t2 * t); }
  const double Wm[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
  double W2[9];
  mat_mul(Wm, Wm, W2);
  double Vm[9];
  for (int i = 0; i < 9; ++i) {
    const double id = (i % 4 == 0) ? 1.0 : 0.0;
    D.R[i] = id + a * Wm[i] + b * W2[i];
    Vm[i] = id + b * Wm[i] + c * W2[i];
  }
  mat_vec(Vm, v, D.p);
  se3_mul(&A, &D, &E);
  memcpy(out12, E.R, sizeof E.R); memcpy(out12 + 9, E.p, sizeof E.p);
}

int orc_interpolate(int64_t n, const double* A, const double* B, const double* alpha, double* out) {
  for (int64_t i = 0; i < n; ++i) se3_interpolate(A + 12 * i, B + 12 * i, alpha[i], out + 12 * i);
  return 0;
}

/* path.project_path loop (path.py:137-160) for n edges: q_path [n][max_steps][nq]; returns n_valid per edge. */
int orc_project_edges(const orc_table_t* t, int64_t n, int max_steps, const double* q_start, const double* pose_a,
                      const double* pose_b, const int32_t* num_steps, double eps, double dt, int max_iters,
                      double damping, int threads, double* q_path, int32_t* n_valid, int32_t* iters_total) {
  const int nq = t->nq;
#ifdef _OPENMP
  omp_set_num_threads(threads > 0 ? threads : omp_get_num_procs());
#pragma omp parallel for schedule(dynamic, 1)
#endif
  for (int64_t e = 0; e < n; ++e) {
    double q[ORC_MAX_NQ], qn[ORC_MAX_NQ], pose[12], r[2];
    memcpy(q, q_start + e * nq, sizeof(double) * nq);
    int ok_steps = 0, total = 0;
    for (int step = 1; step <= num_steps[e] && step <= max_steps; ++step) {
      int it;
      se3_interpolate(pose_a + 12 * e, pose_b + 12 * e, (double)step / (double)num_steps[e], pose);
      const int ok = solve_one(t, q, pose, eps, dt, max_iters, damping, qn, &it, r);
      total += it;
      if (!ok) break;
      memcpy(q, qn, sizeof(double) * nq);
      memcpy(q_path + ((e * max_steps) + (step - 1)) * nq, q, sizeof(double) * nq);
      ok_steps = step;
    }
    n_valid[e] = ok_steps;
    if (iters_total) iters_total[e] = total;
  }
  return 0;
}

#include "collision_oracle.inc"

int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
