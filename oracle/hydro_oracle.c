/*
 * hydro_oracle.c -- TEST INFRASTRUCTURE ONLY (parity oracle, never the product path).
 *
 * A plain-C, float64, strict-IEEE restatement of the reference's CPU hydrodynamics
 * path.  Citations are relative to /root/reference/src/scripts/physics/ :
 *
 *   numba_hydrodynamics.py          (the seven @njit force functions)
 *   numba_hydrodynamics_wrapper.py  (constant record + box geometry precompute)
 *   hydrodynamics_behavior.py:194-238 (numeric tail: finite-difference
 *                                    acceleration, lever-arm torques, clamp)
 *
 * It deliberately keeps the reference's *shape* of computation (27 explicit
 * keypoints, 6 explicit faces, dense 6x6 added-mass matvec, world-space lever
 * arms) so that it is a checker for the structured CUDA kernels rather than a
 * twin of them.  Built with -ffp-contract=off: every operation rounds once.
 *
 * Parity pin: oracle/make_golden.py runs the untouched reference Numba code
 * (imported in place from /root/reference) and tests/test_oracle_golden.py
 * checks this file against those vectors and against SURVEY.md Appendix B.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference
 * legs may load this library.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_API __attribute__((visibility("default")))

/* Constant record of one body: numba_hydrodynamics_wrapper.py:9-32 */
typedef struct {
    double width, depth, height, total_volume;
    double water_density, gravity;
    double linear_drag_coefficient, angular_drag_coefficient;
    double linear_damping, angular_damping, lift_coefficient;
    double local_keypoints[27][3];
    double local_face_centers[6][3];
    double face_areas[6];
    double local_face_normals[6][3];
    double added_mass_matrix[6][6];
} oracle_body_t;

/* Result of solve_hydrodynamics: numba_hydrodynamics.py:314 (9-tuple) */
typedef struct {
    double buoyancy_force[3];
    double drag_force[3];
    double lift_force[3];
    double drag_torque[3];
    double added_mass_force[3];
    double added_mass_torque[3];
    double center_of_buoyancy[3];
    double center_of_pressure[3];
    double sub_ratio;
    /* 1 when the unmodified reference raises TypeError for this input
     * (wet body with speed <= 1e-6: the `return` at numba_hydrodynamics.py:143
     * sits inside the `if speed > 1e-6:` of :118).  The numbers above then
     * follow the evident intent of the "Defaults" at :113-115 (cop=cob, area=0). */
    int32_t reference_raises;
    int32_t _pad;
} oracle_out_t;

/* numba_hydrodynamics_wrapper.py:55-81 (_create_cube_keypoints) */
static void create_cube_keypoints(oracle_body_t *b)
{
    const double x = b->width / 2.0, y = b->depth / 2.0, z = b->height / 2.0;
    /* layer order: top (+z), middle (0), bottom (-z); rows +y,0,-y; cols -x,0,+x */
    const double zs[3] = {+z, 0.0, -z};
    const double ys[3] = {+y, 0.0, -y};
    const double xs[3] = {-x, 0.0, +x};
    int n = 0;
    for (int l = 0; l < 3; ++l)
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) {
                b->local_keypoints[n][0] = xs[c];
                b->local_keypoints[n][1] = ys[r];
                b->local_keypoints[n][2] = zs[l];
                ++n;
            }
    const double fc[6][3] = {{x, 0, 0}, {-x, 0, 0}, {0, y, 0}, {0, -y, 0}, {0, 0, z}, {0, 0, -z}};
    memcpy(b->local_face_centers, fc, sizeof fc);
}

/* numba_hydrodynamics_wrapper.py:83-99 (_create_cube_facepoints) */
static void create_cube_facepoints(oracle_body_t *b)
{
    b->face_areas[0] = b->depth * b->height;
    b->face_areas[1] = b->depth * b->height;
    b->face_areas[2] = b->width * b->height;
    b->face_areas[3] = b->width * b->height;
    b->face_areas[4] = b->width * b->depth;
    b->face_areas[5] = b->width * b->depth;
    const double fn[6][3] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
    memcpy(b->local_face_normals, fn, sizeof fn);
}

/* numba_hydrodynamics_wrapper.py:101-112 (_create_added_mass_matrix) */
static void create_added_mass_matrix(oracle_body_t *b, double lin_c, double ang_c)
{
    const double w = b->width, d = b->depth, h = b->height, vol = b->total_volume,
                 rho = b->water_density;
    memset(b->added_mass_matrix, 0, sizeof b->added_mass_matrix);
    b->added_mass_matrix[0][0] = vol * lin_c * rho;
    b->added_mass_matrix[1][1] = vol * lin_c * rho;
    b->added_mass_matrix[2][2] = vol * lin_c * rho;
    b->added_mass_matrix[3][3] = vol * (d * d + h * h) * ang_c * rho;
    b->added_mass_matrix[4][4] = vol * (w * w + h * h) * ang_c * rho;
    b->added_mass_matrix[5][5] = vol * (w * w + d * d) * ang_c * rho;
}

/* Wrapper ctor, argument order of numba_hydrodynamics_wrapper.py:9-10:
 * ctor[12] = width, depth, height, linear_drag_coefficient, angular_drag_coefficient,
 *            linear_damping, angular_damping, water_density, gravity,
 *            linear_mass_coeff, angular_mass_coeff, lift_coefficient */
ORACLE_API void oracle_body_init(oracle_body_t *b, const double *ctor)
{
    b->width = ctor[0];
    b->depth = ctor[1];
    b->height = ctor[2];
    b->total_volume = ctor[0] * ctor[1] * ctor[2];
    b->water_density = ctor[7];
    b->gravity = ctor[8];
    b->linear_drag_coefficient = ctor[3];
    b->angular_drag_coefficient = ctor[4];
    b->linear_damping = ctor[5];
    b->angular_damping = ctor[6];
    b->lift_coefficient = ctor[11];
    create_cube_facepoints(b);
    create_cube_keypoints(b);
    create_added_mass_matrix(b, ctor[9], ctor[10]);
}

/* numba_hydrodynamics.py:8-51 (quaternion_to_matrix); xyzw, no normalisation */
static void quaternion_to_matrix(const double q[4], double m[3][3])
{
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    const double x2 = x + x, y2 = y + y, z2 = z + z;
    const double xx = x * x2, xy = x * y2, xz = x * z2;
    const double yy = y * y2, yz = y * z2, zz = z * z2;
    const double wx = w * x2, wy = w * y2, wz = w * z2;
    m[0][0] = 1.0 - (yy + zz);
    m[0][1] = xy - wz;
    m[0][2] = xz + wy;
    m[1][0] = xy + wz;
    m[1][1] = 1.0 - (xx + zz);
    m[1][2] = yz - wx;
    m[2][0] = xz - wy;
    m[2][1] = yz + wx;
    m[2][2] = 1.0 - (xx + yy);
}

static inline void matvec3(const double m[3][3], const double v[3], double out[3])
{
    for (int r = 0; r < 3; ++r) {
        double acc = m[r][0] * v[0];
        acc += m[r][1] * v[1];
        acc += m[r][2] * v[2];
        out[r] = acc;
    }
}

static inline void matTvec3(const double m[3][3], const double v[3], double out[3])
{
    for (int r = 0; r < 3; ++r) {
        double acc = m[0][r] * v[0];
        acc += m[1][r] * v[1];
        acc += m[2][r] * v[2];
        out[r] = acc;
    }
}

static inline double norm3(const double v[3])
{
    return sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
}

static inline void cross3(const double a[3], const double b[3], double out[3])
{
    out[0] = a[1] * b[2] - a[2] * b[1];
    out[1] = a[2] * b[0] - a[0] * b[2];
    out[2] = a[0] * b[1] - a[1] * b[0];
}

static int g_warp_compat; /* defined below (Warp-twin deviations) */

/* numba_hydrodynamics.py:53-105 (analyze_submersion_and_cob) */
static double analyze_submersion_and_cob(const double wk[27][3], const double position[3],
                                         double cob[3])
{
    double z_min = INFINITY, z_max = -INFINITY;
    double sum_x = 0.0, sum_y = 0.0, sum_z = 0.0;
    int submerged_count = 0;
    for (int i = 0; i < 27; ++i) {
        const double px = wk[i][0], py = wk[i][1], pz = wk[i][2];
        if (pz < z_min) z_min = pz;
        if (pz > z_max) z_max = pz;
        if (pz < 0) {
            sum_x += px;
            sum_y += py;
            sum_z += pz;
            submerged_count += 1;
        }
    }
    memcpy(cob, position, 3 * sizeof(double));
    if (z_min >= 0) return 0.0; /* fully out */
    if (z_max <= 0) {           /* fully in  */
        if (g_warp_compat && submerged_count != 0) { /* C9: warp_hydrodynamics.py:57-58 */
            const double inv_c = 1.0 / submerged_count;
            cob[0] = sum_x * inv_c;
            cob[1] = sum_y * inv_c;
            cob[2] = sum_z * inv_c;
        }
        return 1.0;
    }
    const double total_height = z_max - z_min;
    double ratio;
    if (total_height < 1e-6) {
        ratio = (z_min < 0) ? 1.0 : 0.0;
    } else {
        ratio = -z_min / total_height;
        if (ratio > 1.0) ratio = 1.0;
    }
    if (submerged_count != 0) {
        const double inv_c = 1.0 / submerged_count;
        cob[0] = sum_x * inv_c;
        cob[1] = sum_y * inv_c;
        cob[2] = sum_z * inv_c;
    }
    return ratio;
}

/* numba_hydrodynamics.py:107-143 (calculate_pressure_and_area).
 * Returns 0 when the reference returns a value, 1 when it falls off the end. */
static int calculate_pressure_and_area(double speed, const double vel_dir[3], const double cob[3],
                                       const double rot[3][3], const oracle_body_t *b,
                                       const double position[3], double cop[3], double *area)
{
    double total_projected_area = 0.0;
    memcpy(cop, cob, 3 * sizeof(double));
    *area = 0.0;
    if (!(speed > 1e-6)) return 1;

    double cop_weighted_sum[3] = {0.0, 0.0, 0.0};
    for (int i = 0; i < 6; ++i) {
        double wn[3], wc[3];
        matvec3(rot, b->local_face_normals[i], wn);
        matvec3(rot, b->local_face_centers[i], wc);
        wc[0] += position[0];
        wc[1] += position[1];
        wc[2] += position[2];
        double dotp = wn[0] * vel_dir[0];
        dotp += wn[1] * vel_dir[1];
        dotp += wn[2] * vel_dir[2];
        const double alignment = -dotp;
        if (alignment > 0) {
            if (wc[2] < 0) {
                const double area_val = alignment * b->face_areas[i];
                total_projected_area += area_val;
                cop_weighted_sum[0] += wc[0] * area_val;
                cop_weighted_sum[1] += wc[1] * area_val;
                cop_weighted_sum[2] += wc[2] * area_val;
            }
        }
    }
    if (total_projected_area > 1e-6) {
        cop[0] = cop_weighted_sum[0] / total_projected_area;
        cop[1] = cop_weighted_sum[1] / total_projected_area;
        cop[2] = cop_weighted_sum[2] / total_projected_area;
    }
    *area = total_projected_area;
    return 0;
}

/* numba_hydrodynamics.py:145-182 (calculate_hybrid_drag) */
static void calculate_hybrid_drag(double speed, const double vel_dir[3], double sub_ratio,
                                  double water_density, double area, double volume,
                                  double linear_drag_coeff, double linear_damping,
                                  const double linear_vel[3], double angular_drag_coeff,
                                  double angular_damping, const double angular_vel[3],
                                  double drag_force[3], double drag_torque[3])
{
    const double LOW_SPEED_THRESHOLD = 0.2;
    double quad[3] = {0.0, 0.0, 0.0};
    if (speed > 1e-6) {
        const double drag_mag = 0.5 * water_density * (speed * speed) * linear_drag_coeff * area;
        for (int k = 0; k < 3; ++k) quad[k] = -drag_mag * vel_dir[k];
    }
    double lin_damp_scale = 1.0;
    if (speed < LOW_SPEED_THRESHOLD) lin_damp_scale = speed / LOW_SPEED_THRESHOLD;
    for (int k = 0; k < 3; ++k) {
        const double damp = -linear_damping * linear_vel[k] * lin_damp_scale;
        drag_force[k] = (quad[k] + damp) * sub_ratio;
    }

    const double ang_speed = norm3(angular_vel);
    double quad_t[3] = {0.0, 0.0, 0.0};
    if (ang_speed > 1e-6) {
        const double ang_mag =
            0.5 * water_density * (ang_speed * ang_speed) * angular_drag_coeff * volume;
        for (int k = 0; k < 3; ++k) quad_t[k] = -ang_mag * (angular_vel[k] / ang_speed);
    }
    double ang_damp_scale = 1.0;
    if (ang_speed < LOW_SPEED_THRESHOLD) ang_damp_scale = ang_speed / LOW_SPEED_THRESHOLD;
    for (int k = 0; k < 3; ++k) {
        const double damp = -angular_damping * angular_vel[k] * ang_damp_scale;
        drag_torque[k] = (quad_t[k] + damp) * sub_ratio;
    }
}

/* numba_hydrodynamics.py:184-217 (calculate_lift) */
static void calculate_lift(double speed, const double vel_dir[3], const double rot[3][3],
                           double area, double water_density, double lift_coeff_param,
                           double sub_ratio, double lift[3])
{
    lift[0] = lift[1] = lift[2] = 0.0;
    if (speed < 1e-6 || sub_ratio <= 1e-9) return;
    const double up[3] = {rot[0][2], rot[1][2], rot[2][2]};
    double dotp = -(up[0] * vel_dir[0] + up[1] * vel_dir[1] + up[2] * vel_dir[2]);
    if (dotp > 1.0) dotp = 1.0;
    else if (dotp < -1.0) dotp = -1.0;
    const double angle_of_attack = asin(dotp);
    const double lift_coefficient = sin(2 * angle_of_attack);
    double lift_magnitude = 0.5 * water_density * (speed * speed) * lift_coefficient * area;
    lift_magnitude *= lift_coeff_param;
    double axis[3];
    cross3(vel_dir, up, axis);
    const double n = norm3(axis);
    if (n < 1e-6) return;
    axis[0] /= n;
    axis[1] /= n;
    axis[2] /= n;
    double dir[3];
    cross3(axis, vel_dir, dir);
    for (int k = 0; k < 3; ++k) lift[k] = lift_magnitude * dir[k] * sub_ratio;
}

/* Behavioural deviations of the reference's Warp twin (warp_hydrodynamics.py, SURVEY.md
 * Appendix C), switchable so that the CUDA warp-compat mode has a checker too.  Pinned against the
 * reference's own kernel source executed through oracle/warp_shim (tests/golden/reference_warp_golden.npz):
 *   C1  accelerations are rotated FORWARD (quat_rotate) instead of by R^T (:216-217)
 *   C3  a dry body returns cob = cop = position instead of zeros (:59-61, :290)
 *   C4  flow along the body z axis: lift_dir is read without having been assigned (:196-200) -- reported
 *       through reference_raises like the at-rest defect (A.8); the oracle returns zero lift there
 *   C7  every rotation is wp.quat_rotate, x (2w^2 - 1) + 2 q (q.x) + 2 w (q x x), which is R(q) + 2 (|q|^2 - 1) I:
 *       identical for unit quaternions, different for the un-normalised ones the Numba path accepts
 *   C9  the centre of buoyancy is the mean of the strictly-wet keypoints whenever there is one (:57-58), also for
 *       a fully submerged body (Numba returns p there, :97-98): differs when keypoints lie exactly on z = 0 */
static int g_warp_compat = 0;
ORACLE_API void oracle_set_warp_compat(int enable) { g_warp_compat = enable != 0; }

/* numba_hydrodynamics.py:219-253 (calculate_added_mass); dense 6x6 matvec */
static void calculate_added_mass(double sub_ratio, const double lin_acc[3], const double ang_acc[3],
                                 const double rot[3][3], const double M[6][6], double force[3],
                                 double torque[3])
{
    force[0] = force[1] = force[2] = 0.0;
    torque[0] = torque[1] = torque[2] = 0.0;
    if (sub_ratio <= 1e-9) return;
    double accel_6d[6];
    if (g_warp_compat) {
        matvec3(rot, lin_acc, &accel_6d[0]);
        matvec3(rot, ang_acc, &accel_6d[3]);
    } else {
        matTvec3(rot, lin_acc, &accel_6d[0]);
        matTvec3(rot, ang_acc, &accel_6d[3]);
    }
    double ft[6];
    for (int i = 0; i < 6; ++i) {
        double acc = 0.0;
        for (int j = 0; j < 6; ++j) acc += M[i][j] * accel_6d[j];
        ft[i] = -1.0 * acc;
    }
    double fw[3], tw[3];
    matvec3(rot, &ft[0], fw);
    matvec3(rot, &ft[3], tw);
    for (int k = 0; k < 3; ++k) {
        force[k] = fw[k] * sub_ratio;
        torque[k] = tw[k] * sub_ratio;
    }
}

/* numba_hydrodynamics.py:255-314 (solve_hydrodynamics) */
ORACLE_API void oracle_solve(const oracle_body_t *b, const double position[3],
                             const double quat_xyzw[4], const double linear_vel[3],
                             const double angular_vel[3], const double linear_accel[3],
                             const double angular_accel[3], oracle_out_t *out)
{
    memset(out, 0, sizeof *out);
    double rot[3][3];
    quaternion_to_matrix(quat_xyzw, rot);
    if (g_warp_compat) { /* C7: wp.quat_rotate == R(q) + 2 (|q|^2 - 1) I */
        const double x = quat_xyzw[0], y = quat_xyzw[1], z = quat_xyzw[2], w = quat_xyzw[3];
        const double d2 = 2.0 * (((x * x + y * y) + (z * z + w * w)) - 1.0);
        rot[0][0] += d2;
        rot[1][1] += d2;
        rot[2][2] += d2;
    }
    double wk[27][3];
    for (int i = 0; i < 27; ++i) {
        matvec3(rot, b->local_keypoints[i], wk[i]);
        wk[i][0] += position[0];
        wk[i][1] += position[1];
        wk[i][2] += position[2];
    }
    double cob[3];
    const double sub_ratio = analyze_submersion_and_cob(wk, position, cob);
    if (sub_ratio <= 1e-9) { /* :277-279 -- every output zero, incl. cob/cop */
        if (g_warp_compat) {
            memcpy(out->center_of_buoyancy, position, 3 * sizeof(double));
            memcpy(out->center_of_pressure, position, 3 * sizeof(double));
        }
        return;
    }

    out->sub_ratio = sub_ratio;
    out->buoyancy_force[2] = b->water_density * (sub_ratio * b->total_volume) * b->gravity;

    const double speed = norm3(linear_vel);
    double vel_dir[3] = {0.0, 0.0, 0.0};
    if (speed > 1e-6)
        for (int k = 0; k < 3; ++k) vel_dir[k] = linear_vel[k] / speed;

    double cop[3], area;
    out->reference_raises =
        calculate_pressure_and_area(speed, vel_dir, cob, rot, b, position, cop, &area);

    calculate_hybrid_drag(speed, vel_dir, sub_ratio, b->water_density, area, b->total_volume,
                          b->linear_drag_coefficient, b->linear_damping, linear_vel,
                          b->angular_drag_coefficient, b->angular_damping, angular_vel,
                          out->drag_force, out->drag_torque);
    calculate_lift(speed, vel_dir, rot, area, b->water_density, b->lift_coefficient, sub_ratio,
                   out->lift_force);
    if (g_warp_compat && speed > 1e-6) { /* C4: lift_dir unassigned when |v_hat x up| <= 1e-6 */
        const double up[3] = {rot[0][2], rot[1][2], rot[2][2]};
        double axis[3];
        cross3(vel_dir, up, axis);
        if (!(norm3(axis) > 1e-6)) out->reference_raises = 1;
    }
    calculate_added_mass(sub_ratio, linear_accel, angular_accel, rot, b->added_mass_matrix,
                         out->added_mass_force, out->added_mass_torque);
    memcpy(out->center_of_buoyancy, cob, sizeof cob);
    memcpy(out->center_of_pressure, cop, sizeof cop);
}

/* hydrodynamics_behavior.py:212-226: lever-arm torques in world space (as the
 * reference does, cancellation included), net wrench, safety clamp. */
ORACLE_API void oracle_net_wrench(const oracle_out_t *c, const double position[3], double mass,
                                  double net_force[3], double net_torque[3], double *scale_out)
{
    double arm_b[3], arm_p[3], tb[3], td[3], tl[3];
    for (int k = 0; k < 3; ++k) {
        arm_b[k] = c->center_of_buoyancy[k] - position[k];
        arm_p[k] = c->center_of_pressure[k] - position[k];
    }
    cross3(arm_b, c->buoyancy_force, tb);
    cross3(arm_p, c->drag_force, td);
    cross3(arm_p, c->lift_force, tl);
    for (int k = 0; k < 3; ++k) {
        net_force[k] = c->buoyancy_force[k] + c->drag_force[k] + c->lift_force[k] +
                       c->added_mass_force[k];
        net_torque[k] =
            tb[k] + td[k] + tl[k] + c->drag_torque[k] + c->added_mass_torque[k];
    }
    const double MAX_ACCEL = 500.0;
    const double max_force = mass * MAX_ACCEL;
    const double force_mag = norm3(net_force);
    double scale = max_force / (force_mag + 1e-6);
    if (scale > 1.0) scale = 1.0;
    for (int k = 0; k < 3; ++k) {
        net_force[k] *= scale;
        net_torque[k] *= scale;
    }
    if (scale_out) *scale_out = scale;
}

/* Optional dense added-mass matrices for the batched drivers.  calculate_added_mass
 * (numba_hydrodynamics.py:219-253) multiplies by whatever 6x6 it is handed; the wrapper only
 * builds a diagonal one (numba_hydrodynamics_wrapper.py:101-112).  With n_types > 0, body i of a
 * batch uses M[slot_type[i % n_slots]] instead of that diagonal.  n_types = 0 switches it off. */
#define ORACLE_MAX_DENSE_TYPES 64
#define ORACLE_MAX_DENSE_SLOTS 256
static int g_dense_types = 0, g_dense_slots = 0;
static double g_dense[ORACLE_MAX_DENSE_TYPES][6][6];
static int32_t g_dense_slot[ORACLE_MAX_DENSE_SLOTS];

ORACLE_API int oracle_set_added_mass_dense(int n_types, const double *M, int n_slots,
                                           const int32_t *slot_type)
{
    if (n_types <= 0) {
        g_dense_types = g_dense_slots = 0;
        return 0;
    }
    if (n_types > ORACLE_MAX_DENSE_TYPES || n_slots < 1 || n_slots > ORACLE_MAX_DENSE_SLOTS || !M ||
        !slot_type)
        return 1;
    for (int i = 0; i < n_slots; ++i)
        if (slot_type[i] < 0 || slot_type[i] >= n_types) return 1;
    memcpy(g_dense, M, (size_t)n_types * 36 * sizeof(double));
    memcpy(g_dense_slot, slot_type, (size_t)n_slots * sizeof(int32_t));
    g_dense_types = n_types;
    g_dense_slots = n_slots;
    return 0;
}

static inline void apply_dense_added_mass(oracle_body_t *b, int64_t i)
{
    if (g_dense_types > 0)
        memcpy(b->added_mass_matrix, g_dense[g_dense_slot[i % g_dense_slots]],
               sizeof b->added_mass_matrix);
}

/* ------------------------------------------------------------------------- */
/* Batched drivers (one body per iteration, OpenMP over bodies).              */
/* ------------------------------------------------------------------------- */

/* ctor: (n,12) per-body wrapper-ctor rows, or a single row when ctor_stride==0. */
ORACLE_API void oracle_components_batch(int64_t n, const double *ctor, int64_t ctor_stride,
                                        const double *pos, const double *quat_xyzw,
                                        const double *lin_vel, const double *ang_vel,
                                        const double *lin_acc, const double *ang_acc,
                                        oracle_out_t *out, int n_threads)
{
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
    (void)n_threads;
#pragma omp parallel
    {
        oracle_body_t body;
        if (ctor_stride == 0) oracle_body_init(&body, ctor);
#pragma omp for schedule(static)
        for (int64_t i = 0; i < n; ++i) {
            if (ctor_stride != 0) oracle_body_init(&body, ctor + i * ctor_stride);
            apply_dense_added_mass(&body, i);
            oracle_solve(&body, pos + 3 * i, quat_xyzw + 4 * i, lin_vel + 3 * i, ang_vel + 3 * i,
                         lin_acc + 3 * i, ang_acc + 3 * i, &out[i]);
        }
    }
}

/* One behaviour step for n independent bodies: hydrodynamics_behavior.py:194-238.
 * quat_wxyz != 0 applies the :194 permutation.  prev_lin/prev_ang are updated
 * in place (:237-238).  dt <= 1e-6 skips the step entirely (:139).
 * flags[i] (optional): bit0 = reference_raises, bit1 = clamp active. */
ORACLE_API void oracle_step_batch(int64_t n, const double *ctor, int64_t ctor_stride,
                                  const double *mass, int64_t mass_stride, const double *pos,
                                  const double *quat, int quat_wxyz, const double *lin_vel,
                                  const double *ang_vel, double *prev_lin, double *prev_ang,
                                  double dt, double *net_force, double *net_torque,
                                  oracle_out_t *components, int32_t *flags, int n_threads)
{
    if (dt <= 1e-6) return;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
    (void)n_threads;
#pragma omp parallel
    {
        oracle_body_t body;
        if (ctor_stride == 0) oracle_body_init(&body, ctor);
#pragma omp for schedule(static)
        for (int64_t i = 0; i < n; ++i) {
            if (ctor_stride != 0) oracle_body_init(&body, ctor + i * ctor_stride);
            apply_dense_added_mass(&body, i);
            double q[4], a[3], al[3];
            if (quat_wxyz) {
                q[0] = quat[4 * i + 1];
                q[1] = quat[4 * i + 2];
                q[2] = quat[4 * i + 3];
                q[3] = quat[4 * i + 0];
            } else {
                memcpy(q, quat + 4 * i, sizeof q);
            }
            for (int k = 0; k < 3; ++k) {
                a[k] = (lin_vel[3 * i + k] - prev_lin[3 * i + k]) / dt;
                al[k] = (ang_vel[3 * i + k] - prev_ang[3 * i + k]) / dt;
            }
            oracle_out_t c;
            oracle_solve(&body, pos + 3 * i, q, lin_vel + 3 * i, ang_vel + 3 * i, a, al, &c);
            double scale;
            oracle_net_wrench(&c, pos + 3 * i, mass[i * mass_stride], net_force + 3 * i,
                              net_torque + 3 * i, &scale);
            if (components) components[i] = c;
            if (flags) flags[i] = (c.reference_raises ? 1 : 0) | (scale < 1.0 ? 2 : 0);
            for (int k = 0; k < 3; ++k) {
                prev_lin[3 * i + k] = lin_vel[3 * i + k];
                prev_ang[3 * i + k] = ang_vel[3 * i + k];
            }
        }
    }
}

/* Per-robot wrench about the robot's slot-0 body:
 * F_R = sum F_i ; tau_R = sum (tau_i + (p_i - p_base) x F_i). out is (n_robots,6). */
ORACLE_API void oracle_robot_wrench(int64_t n_robots, int64_t bodies_per_robot, const double *pos,
                                    const double *force, const double *torque, double *out)
{
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n_robots; ++r) {
        const double *pb = pos + 3 * (r * bodies_per_robot);
        double F[3] = {0, 0, 0}, T[3] = {0, 0, 0};
        for (int64_t j = 0; j < bodies_per_robot; ++j) {
            const int64_t i = r * bodies_per_robot + j;
            const double arm[3] = {pos[3 * i] - pb[0], pos[3 * i + 1] - pb[1],
                                   pos[3 * i + 2] - pb[2]};
            double c[3];
            cross3(arm, force + 3 * i, c);
            for (int k = 0; k < 3; ++k) {
                F[k] += force[3 * i + k];
                T[k] += torque[3 * i + k] + c[k];
            }
        }
        for (int k = 0; k < 3; ++k) {
            out[6 * r + k] = F[k];
            out[6 * r + 3 + k] = T[k];
        }
    }
}

ORACLE_API int oracle_sizeof_out(void) { return (int)sizeof(oracle_out_t); }
ORACLE_API int oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
