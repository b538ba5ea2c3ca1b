// Host-side instantiation of silver2_isaacsim_b200/csrc/h2o_model.cuh.
//
// TEST TOOL ONLY: lets the CPU-only test suite exercise the exact per-body
// arithmetic the CUDA kernels inline (same header, same precision policies)
// against the oracle, without a GPU.  It is not part of the package, is not
// loaded by any product module and is not a fallback path.
#include <cstdint>
#include <cstring>

#include "../../silver2_isaacsim_b200/csrc/h2o_model.cuh"

using namespace h2o;

static int g_no_fallback = 0;  // precision study: report what the fast path alone would give
extern "C" __attribute__((visibility("default"))) void emul_set_no_fallback(int v) { g_no_fallback = v; }

template <typename S, typename H, typename L, bool kExactTrig, bool kFast = false>
static void run(int64_t n, const double* pos, const double* quat, const double* v, const double* w,
                const double* pl, const double* pa, const double* coeff, double rho, double g,
                double dt, double* F, double* T, double* comp, uint32_t* masks, const double* dense = nullptr,
                const int32_t* dense_slot = nullptr, int dense_slots = 0)
{
    const L inv_dt = L(1.0 / dt);
    for (int64_t i = 0; i < n; ++i) {
        BodyIn<H, L> in;
        in.pz = H(S(pos[3 * i + 2]));
        in.qx = H(S(quat[4 * i + 0])); in.qy = H(S(quat[4 * i + 1]));
        in.qz = H(S(quat[4 * i + 2])); in.qw = H(S(quat[4 * i + 3]));
        const L vx = L(S(v[3 * i])), vy = L(S(v[3 * i + 1])), vz = L(S(v[3 * i + 2]));
        const L wx = L(S(w[3 * i])), wy = L(S(w[3 * i + 1])), wz = L(S(w[3 * i + 2]));
        in.vx = vx; in.vy = vy; in.vz = vz;
        in.wx = wx; in.wy = wy; in.wz = wz;
        const L sc = kFast ? L(1) : inv_dt;  // the fast path folds 1/dt into the added-mass constants
        in.acc_scale = kFast ? inv_dt : L(1);
        in.ax = (vx - L(S(pl[3 * i]))) * sc;
        in.ay = (vy - L(S(pl[3 * i + 1]))) * sc;
        in.az = (vz - L(S(pl[3 * i + 2]))) * sc;
        in.bx = (wx - L(S(pa[3 * i]))) * sc;
        in.by = (wy - L(S(pa[3 * i + 1]))) * sc;
        in.bz = (wz - L(S(pa[3 * i + 2]))) * sc;
        const double* c = coeff + 11 * i;
        in.dimx = L(S(c[0])); in.dimy = L(S(c[1])); in.dimz = L(S(c[2]));
        in.c_drag = L(S(c[3])); in.c_drag_ang = L(S(c[4]));
        in.k_damp = L(S(c[5])); in.k_damp_ang = L(S(c[6]));
        in.c_am = L(S(c[7])); in.c_am_ang = L(S(c[8])); in.c_lift = L(S(c[9]));
        in.rho_h = H(rho); in.grav_h = H(g); in.rho = L(rho);
        in.warp_compat = false;
        L mloc[36];
        in.am_dense = nullptr;
        if (dense) {
            const double* m = dense + 36 * dense_slot[i % dense_slots];
            for (int k = 0; k < 36; ++k) mloc[k] = L(S(m[k]));
            in.am_dense = mloc;
        }
        Terms<H, L> t;
        L f[3], tq[3];
        bool clamped;
        if (kFast) {  // same dispatch as body_step() in h2o_kernels.cuh: fast path, float64 re-evaluation when flagged
            H ratio;
            bool still, suspect;
            L dg[14] = {0};
            uint32_t kp_mask;
            body_wrench_fast<H, L>(in, L(S(c[10])), f, tq, clamped, ratio, still, suspect, kp_mask, comp ? dg : nullptr);
            if (comp) {  // mode 4 returns the term-group magnitudes where the other modes return components
                for (int k = 0; k < 14; ++k) comp[28 * i + k] = double(dg[k]);
                comp[28 * i + 27] = suspect ? 2.0 : 1.0;
            }
            t.kp_mask = 0;
            if (suspect && !g_no_fallback) {
                BodyIn<double, double> ex;
                ex.pz = double(in.pz); ex.qx = double(in.qx); ex.qy = double(in.qy); ex.qz = double(in.qz); ex.qw = double(in.qw);
                ex.vx = double(in.vx); ex.vy = double(in.vy); ex.vz = double(in.vz);
                ex.wx = double(in.wx); ex.wy = double(in.wy); ex.wz = double(in.wz);
                const double isc = 1.0 / dt;
                ex.ax = (double(in.vx) - double(S(pl[3 * i]))) * isc; ex.ay = (double(in.vy) - double(S(pl[3 * i + 1]))) * isc;
                ex.az = (double(in.vz) - double(S(pl[3 * i + 2]))) * isc;
                ex.bx = (double(in.wx) - double(S(pa[3 * i]))) * isc; ex.by = (double(in.wy) - double(S(pa[3 * i + 1]))) * isc;
                ex.bz = (double(in.wz) - double(S(pa[3 * i + 2]))) * isc;
                ex.acc_scale = 1.0;
                ex.dimx = double(in.dimx); ex.dimy = double(in.dimy); ex.dimz = double(in.dimz);
                ex.c_drag = double(in.c_drag); ex.c_drag_ang = double(in.c_drag_ang); ex.k_damp = double(in.k_damp);
                ex.k_damp_ang = double(in.k_damp_ang); ex.c_am = double(in.c_am); ex.c_am_ang = double(in.c_am_ang);
                ex.c_lift = double(in.c_lift); ex.warp_compat = false;
                ex.rho_h = rho; ex.grav_h = g; ex.rho = rho;
                double md[36];
                ex.am_dense = nullptr;
                if (dense) {
                    for (int k = 0; k < 36; ++k) md[k] = double(mloc[k]);
                    ex.am_dense = md;
                }
                Terms<double, double> te;
                double fe[3], qe[3];
                body_terms<double, double, false>(ex, te, &kp_mask);  // as the tile kernel's deferred roles do
                net_wrench<double, double>(te, double(S(c[10])), fe, qe, clamped);
                for (int k = 0; k < 3; ++k) { f[k] = L(fe[k]); tq[k] = L(qe[k]); }
            }
        } else {
            body_terms<H, L, kExactTrig>(in, t);
            net_wrench<H, L>(t, L(S(c[10])), f, tq, clamped);
        }
        for (int k = 0; k < 3; ++k) {
            F[3 * i + k] = double(S(f[k]));
            T[3 * i + k] = double(S(tq[k]));
        }
        if (comp && !kFast) {
            double* o = comp + 28 * i;
            o[0] = 0; o[1] = 0; o[2] = double(t.fbz);
            for (int k = 0; k < 3; ++k) {
                o[3 + k] = double(t.fd[k]); o[6 + k] = double(t.fl[k]); o[9 + k] = double(t.td[k]);
                o[12 + k] = double(t.fam[k]); o[15 + k] = double(t.tam[k]);
                o[18 + k] = double(t.cob[k]); o[21 + k] = double(t.cop[k]);
            }
            o[24] = double(t.ratio);
            for (int k = 0; k < 3; ++k) o[25 + k] = double(t.tarm[k]);
        }
        if (masks) masks[i] = t.kp_mask;
    }
}

extern "C" __attribute__((visibility("default")))
int emul_step(int mode, int exact_trig, int64_t n, const double* pos, const double* quat,
              const double* v, const double* w, const double* pl, const double* pa,
              const double* coeff, double rho, double g, double dt, double* F, double* T,
              double* comp, uint32_t* masks)
{
#define GO(S, H, L)                                                                             \
    (exact_trig ? run<S, H, L, true>(n, pos, quat, v, w, pl, pa, coeff, rho, g, dt, F, T, comp, masks) \
                : run<S, H, L, false>(n, pos, quat, v, w, pl, pa, coeff, rho, g, dt, F, T, comp, masks))
    switch (mode) {
        case 0: GO(double, double, double); return 0;   // fp64 mode
        case 1: GO(float, double, float); return 0;     // fp32 mode (mixed policy)
        case 2: GO(float, float, float); return 0;      // all-fp32 (for comparison only)
        case 3: GO(float, double, double); return 0;    // fp32 storage, all-fp64 arithmetic
        case 4: run<float, double, float, false, true>(n, pos, quat, v, w, pl, pa, coeff, rho, g, dt, F, T, comp, masks);
                return 0;                               // fp32 mode, fused-step fast path (body frame)
    }
    return 1;
}

// Same, with dense added-mass matrices (n_types,6,6) + slot map; modes 0 and 4 only.
extern "C" __attribute__((visibility("default")))
int emul_step_dense(int mode, int64_t n, const double* pos, const double* quat, const double* v,
                    const double* w, const double* pl, const double* pa, const double* coeff, double rho,
                    double g, double dt, double* F, double* T, const double* dense, const int32_t* dense_slot,
                    int dense_slots)
{
    if (mode == 0) {
        run<double, double, double, false>(n, pos, quat, v, w, pl, pa, coeff, rho, g, dt, F, T, nullptr, nullptr,
                                           dense, dense_slot, dense_slots);
        return 0;
    }
    if (mode == 4) {
        run<float, double, float, false, true>(n, pos, quat, v, w, pl, pa, coeff, rho, g, dt, F, T, nullptr, nullptr,
                                               dense, dense_slot, dense_slots);
        return 0;
    }
    return 1;
}
