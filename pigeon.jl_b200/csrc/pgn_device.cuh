// pgn_device.cuh — device-side vehicle physics for the batched MPC engine (sm_100a).
//
// What it computes follows the reference's Julia (cited per function, paths relative to the reference tree); how it is
// written is B200-first: scalar-generic inlined device functions over either `double` or a small register-resident dual
// number `Dual<NT>` (NT tangent lanes per thread — the linearisation kernel splits the 8/10 tangent directions of one node
// over adjacent lanes instead of carrying them all in one thread), transcendentals shared between the repeated tire-model
// evaluations, everything FP64 (the reference is Float64 throughout).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace pgn {

struct VehParams {   // same field order as the C ABI vector (include/pigeon_b200.h, PGN_VEHICLE_PARAMS_LEN)
    double L, a, b, h, G, m, Izz, mu, Caf, Car, Cd0, Cd1, Cd2;
    double fwd_frac, rwd_frac, fwb_frac, rwb_frac;
    double Fx_max, Fx_min, Px_max, delta_max, kappa_max;
    double inv_fiala_corrected;
};
struct CtrlParams {
    double V_min, V_max, k_V, k_s, ddelta_max, Q_ds, Q_dpsi, Q_e, W_beta, W_r, W_HJI, N_HJI, R_delta, R_ddelta, R_Fx, R_dFx;
};

#define PGN_HD __host__ __device__ __forceinline__

// ---------------------------------------------------------------------------------------------------------------------
// forward-mode dual numbers; branch semantics = derivative of the taken branch (ForwardDiff), vehicle_dynamics.jl:37-47,293-298
template <int NT>
struct Dual {
    double v;
    double d[NT];
    PGN_HD Dual() {}
    PGN_HD Dual(double x) : v(x) {
#pragma unroll
        for (int i = 0; i < NT; i++) d[i] = 0.0;
    }
};
PGN_HD double val(double x) { return x; }
template <int NT> PGN_HD double val(const Dual<NT>& x) { return x.v; }

#define PGN_DUAL_LOOP _Pragma("unroll") for (int i = 0; i < NT; i++)
template <int NT> PGN_HD Dual<NT> operator+(const Dual<NT>& a, const Dual<NT>& b) { Dual<NT> r; r.v = a.v + b.v; PGN_DUAL_LOOP r.d[i] = a.d[i] + b.d[i]; return r; }
template <int NT> PGN_HD Dual<NT> operator-(const Dual<NT>& a, const Dual<NT>& b) { Dual<NT> r; r.v = a.v - b.v; PGN_DUAL_LOOP r.d[i] = a.d[i] - b.d[i]; return r; }
template <int NT> PGN_HD Dual<NT> operator-(const Dual<NT>& a) { Dual<NT> r; r.v = -a.v; PGN_DUAL_LOOP r.d[i] = -a.d[i]; return r; }
template <int NT> PGN_HD Dual<NT> operator*(const Dual<NT>& a, const Dual<NT>& b) { Dual<NT> r; r.v = a.v * b.v; PGN_DUAL_LOOP r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
template <int NT> PGN_HD Dual<NT> operator/(const Dual<NT>& a, const Dual<NT>& b) {
    Dual<NT> r; double inv = 1.0 / b.v; r.v = a.v * inv;
    PGN_DUAL_LOOP r.d[i] = (a.d[i] - r.v * b.d[i]) * inv;
    return r;
}
template <int NT> PGN_HD Dual<NT> operator+(const Dual<NT>& a, double b) { Dual<NT> r = a; r.v += b; return r; }
template <int NT> PGN_HD Dual<NT> operator+(double a, const Dual<NT>& b) { Dual<NT> r = b; r.v += a; return r; }
template <int NT> PGN_HD Dual<NT> operator-(const Dual<NT>& a, double b) { Dual<NT> r = a; r.v -= b; return r; }
template <int NT> PGN_HD Dual<NT> operator-(double a, const Dual<NT>& b) { Dual<NT> r; r.v = a - b.v; PGN_DUAL_LOOP r.d[i] = -b.d[i]; return r; }
template <int NT> PGN_HD Dual<NT> operator*(const Dual<NT>& a, double b) { Dual<NT> r; r.v = a.v * b; PGN_DUAL_LOOP r.d[i] = a.d[i] * b; return r; }
template <int NT> PGN_HD Dual<NT> operator*(double a, const Dual<NT>& b) { return b * a; }
template <int NT> PGN_HD Dual<NT> operator/(const Dual<NT>& a, double b) { return a * (1.0 / b); }
template <int NT> PGN_HD Dual<NT> operator/(double a, const Dual<NT>& b) {
    Dual<NT> r; double inv = 1.0 / b.v; r.v = a * inv; double g = -r.v * inv;
    PGN_DUAL_LOOP r.d[i] = g * b.d[i];
    return r;
}

PGN_HD void sincos_(double x, double& s, double& c) { sincos(x, &s, &c); }
template <int NT> PGN_HD void sincos_(const Dual<NT>& x, Dual<NT>& s, Dual<NT>& c) {
    double sv, cv; sincos(x.v, &sv, &cv);
    s.v = sv; c.v = cv;
    PGN_DUAL_LOOP { s.d[i] = cv * x.d[i]; c.d[i] = -sv * x.d[i]; }
}
PGN_HD double tan_(double x) { return tan(x); }
template <int NT> PGN_HD Dual<NT> tan_(const Dual<NT>& x) { Dual<NT> r; r.v = tan(x.v); double g = 1.0 + r.v * r.v; PGN_DUAL_LOOP r.d[i] = g * x.d[i]; return r; }
PGN_HD double atan2_(double y, double x) { return atan2(y, x); }
template <int NT> PGN_HD Dual<NT> atan2_(const Dual<NT>& y, const Dual<NT>& x) {
    Dual<NT> r; r.v = atan2(y.v, x.v); double h = 1.0 / (x.v * x.v + y.v * y.v);
    PGN_DUAL_LOOP r.d[i] = (x.v * y.d[i] - y.v * x.d[i]) * h;
    return r;
}
PGN_HD double sqrt_(double x) { return sqrt(x); }
template <int NT> PGN_HD Dual<NT> sqrt_(const Dual<NT>& x) { Dual<NT> r; r.v = sqrt(x.v); double g = 0.5 / r.v; PGN_DUAL_LOOP r.d[i] = g * x.d[i]; return r; }
PGN_HD double abs_(double x) { return fabs(x); }
template <int NT> PGN_HD Dual<NT> abs_(const Dual<NT>& x) { return signbit(x.v) ? -x : x; }
template <class T> PGN_HD double sign_(const T& x) { double v = val(x); return (double)((v > 0) - (v < 0)); }
// Base.min / Base.max / Base.clamp on (Dual, constant): ties keep the first argument for min, the second for max
template <class T> PGN_HD T min_c(const T& x, double c) { return (c < val(x)) ? T(c) : x; }
template <class T> PGN_HD T max_c(const T& x, double c) { return (c < val(x)) ? x : T(c); }
template <class T> PGN_HD T clamp_c(const T& x, double lo, double hi) { return (val(x) > hi) ? T(hi) : ((val(x) < lo) ? T(lo) : x); }

// ---------------------------------------------------------------------------------------------------------------------
// Fiala brush tire with friction-circle derating (vehicle_dynamics.jl:35-48); takes tan(alpha) so that the three
// weight-transfer iterations on the front axle share one tangent evaluation.
template <class T>
PGN_HD T fiala_tan(const T& tana, double Ca, double mu, const T& Fx, const T& Fz) {
    T F_max = mu * Fz;
    if (val(abs_(Fx)) >= val(F_max)) return T(0.0);
    T Fy_max = sqrt_(F_max * F_max - Fx * Fx);
    T tana_slide = 3.0 * Fy_max / Ca;
    T ratio = abs_(tana / tana_slide);
    if (val(ratio) <= 1.0) return -Ca * tana * (1.0 - ratio + ratio * ratio / 3.0);
    return -Fy_max * sign_(tana);
}
// _invfialatiremodel (vehicle_dynamics.jl:56-62): literal by default (returns the slip ratio when unsaturated)
PGN_HD double inv_fiala(double Fy, double Ca, double Fy_max, bool corrected) {
    if (fabs(Fy) >= Fy_max) return -(3 * Fy_max / Ca) * sign_(Fy);
    double r = -(1 + cbrt(fabs(Fy) / Fy_max - 1)) * sign_(Fy);
    return corrected ? r * (3 * Fy_max / Ca) : r;
}
// lateral_tire_forces (vehicle_dynamics.jl:64-76): 3 fixed-point iterations of longitudinal weight transfer, then the rear axle
// The tire model only ever uses tan(alpha) (vehicle_dynamics.jl:35-47), so the core takes the tangents of the slip angles.
template <class T>
PGN_HD void lateral_tire_forces_tan(const VehParams& B, const T& tanf, const T& tanr, const T& Fxf, const T& Fxr, const T& sd, const T& cd,
                                    T& Fyf, T& Fyr, int num_iters = 3) {
    Fyf = T(0.0);
    T Fx = Fxf * cd - Fyf * sd + Fxr;
    for (int i = 0; i < num_iters; i++) {
        T Fzf = (B.m * B.G * B.b - B.h * Fx) / B.L;
        Fyf = fiala_tan(tanf, B.Caf, B.mu, Fxf, Fzf);
        Fx = Fxf * cd - Fyf * sd + Fxr;
    }
    T Fzr = (B.m * B.G * B.a + B.h * Fx) / B.L;
    Fyr = fiala_tan(tanr, B.Car, B.mu, Fxr, Fzr);
}
template <class T>
PGN_HD void lateral_tire_forces(const VehParams& B, const T& af, const T& ar, const T& Fxf, const T& Fxr, const T& sd, const T& cd,
                                T& Fyf, T& Fyr, int num_iters = 3) {
    lateral_tire_forces_tan(B, tan_(af), tan_(ar), Fxf, Fxr, sd, cd, Fyf, Fyr, num_iters);
}

#ifdef __CUDACC__
// optimal_control (HJI_computation.jl:133-158), uMode = :max, N = 50 — the callback's "hammer" policy (ros_integration.jl:115-118):
// steering at the limit picked by the sign of B = gV5/m + a gV7/Izz, Fx by a 50-point grid search of A Fx + B Fyf + C Fyr (first
// maximum wins), tire forces from lateral_tire_forces(BM, (0,0,0,Ux,Uy,r), (delta, raw drive/brake split of Fx)).
// The slip angles and the steering sincos do not depend on Fx and are formed once.
__device__ __forceinline__ void hji_optimal_control(const VehParams& P, double Ux, double Uy, double r, const double* g /*gradV[7]*/, double& d_opt, double& Fx_opt) {
    const double A = g[3] / P.m;
    const double Bc = g[4] / P.m + P.a * g[6] / P.Izz;
    const double Cc = g[4] / P.m - P.b * g[6] / P.Izz;
    d_opt = (Bc >= 0) ? P.delta_max : -P.delta_max;
    double sd, cd;
    sincos(d_opt, &sd, &cd);
    const double af = atan2(Uy + P.a * r, Ux) - d_opt, ar = atan2(Uy - P.b * r, Ux);
    double V_opt = -INFINITY;
    Fx_opt = 0.0;
    const int N = 50;
    for (int n = 0; n < N; n++) {
        const double frac = (double)n / (double)(N - 1);
        const double Fx = __dadd_rn(__dmul_rn(frac, P.Fx_max), __dmul_rn(1.0 - frac, P.Fx_min));     // no FMA contraction: Fx_opt is returned verbatim
        double Fxf, Fxr;
        if (Fx > 0) { Fxf = Fx * P.fwd_frac; Fxr = Fx * P.rwd_frac; } else { Fxf = Fx * P.fwb_frac; Fxr = Fx * P.rwb_frac; }
        double Fyf, Fyr;
        lateral_tire_forces<double>(P, af, ar, Fxf, Fxr, sd, cd, Fyf, Fyr);
        const double V = A * Fx + Bc * Fyf + Cc * Fyr;
        if (V > V_opt) { Fx_opt = Fx; V_opt = V; }
    }
}
#endif

enum { MODEL_BICYCLE = 0, MODEL_TRACKING = 1, MODEL_LATERAL = 2 };

// VehicleModel call (vehicle_dynamics.jl:293-316): control limits (Ux de-dualised), drive/brake split, then the bicycle
// model selected at compile time: BicycleModel (:111-134), TrackingBicycleModel (:159-182), LateralTrackingBicycleModel (:205-223).
// q: state (6 or 4), u2 = (delta, Fx), p = (psi_r|V|Ux, kappa, theta, phi); out: time derivative.
template <int KIND, class T>
PGN_HD void vehicle_model(const VehParams& P, const T* q, const T& delta_in, const T& Fx_in, const T& p0, const T& p1, T* out) {
    const T& Ux = (KIND == MODEL_BICYCLE) ? q[3] : (KIND == MODEL_TRACKING ? q[1] : p0);
    const T& Uy = (KIND == MODEL_BICYCLE) ? q[4] : (KIND == MODEL_TRACKING ? q[2] : q[0]);
    const T& r = (KIND == MODEL_BICYCLE) ? q[5] : (KIND == MODEL_TRACKING ? q[3] : q[1]);
    double Uxv = val(Ux);
    T d = clamp_c(delta_in, -P.delta_max, P.delta_max);
    T Fx = max_c(min_c(min_c(Fx_in, P.Fx_max), P.Px_max / Uxv), P.Fx_min);
    T Fxf, Fxr;
    if (val(Fx) > 0) { Fxf = Fx * P.fwd_frac; Fxr = Fx * P.rwd_frac; }
    else             { Fxf = Fx * P.fwb_frac; Fxr = Fx * P.rwb_frac; }
    T sd, cd;
    sincos_(d, sd, cd);
    // slip angles alpha_f = atan(Uy + a r, Ux) - delta, alpha_r = atan(Uy - b r, Ux) (vehicle_dynamics.jl:118-119) enter the tire model only
    // through their tangents: tan(alpha_r) = (Uy - b r) / Ux and tan(alpha_f) = (t - tan(delta)) / (1 + t tan(delta)), t = (Uy + a r) / Ux.
    // Same function (and, through the dual numbers, same derivative) as tan(atan2(.) - delta) without the two atan2 and two tan
    // evaluations, which were 47 % of the linearisation kernel's samples (profiles/r1b_linearize_ncu_full.md).
    T iUx = 1.0 / Ux;
    T tf = (Uy + P.a * r) * iUx, td = sd / cd;
    T tanf = (tf - td) / (1.0 + tf * td);
    T tanr = (Uy - P.b * r) * iUx;
    T Fyf, Fyr;
    lateral_tire_forces_tan(P, tanf, tanr, Fxf, Fxr, sd, cd, Fyf, Fyr);
    T Fyf_t = Fyf * cd + Fxf * sd;
    T dUy = (Fyf_t + Fyr) / P.m - r * Ux;
    T dr = (P.a * Fyf_t - P.b * Fyr) / P.Izz;
    if (KIND == MODEL_LATERAL) {
        T sp, cp;
        sincos_(q[2], sp, cp);
        out[0] = dUy;
        out[1] = dr;
        out[2] = r - Ux * p1;
        out[3] = Ux * sp + Uy * cp;
    } else {
        T Fx_drag = -P.Cd0 - Ux * (P.Cd1 + P.Cd2 * Ux);
        T Fxf_t = Fxf * cd - Fyf * sd;
        T dUx = (Fxf_t + Fxr + Fx_drag) / P.m + r * Uy;
        T sp, cp;
        if (KIND == MODEL_BICYCLE) {
            sincos_(q[2], sp, cp);
            out[0] = -Ux * sp - Uy * cp;
            out[1] = Ux * cp - Uy * sp;
            out[2] = r;
            out[3] = dUx; out[4] = dUy; out[5] = dr;
        } else {
            sincos_(q[4], sp, cp);
            T vs = Ux * cp - Uy * sp;
            out[0] = vs - p0;
            out[1] = dUx; out[2] = dUy; out[3] = dr;
            out[4] = r - vs * p1;
            out[5] = Ux * sp + Uy * cp;
        }
    }
}

// propagate(dynamics, x, StepControl|RampControl) of DifferentialDynamicsModels (not vendored in the reference; call sites
// model_predictive_control.jl:94, coupled_lat_long.jl:253,262): fixed-step RK4, `nsub` sub-steps, control = (delta, Fx, p0, p1)
// ramping linearly from u0 to uf.  NX = 6 (bicycle / tracking) or 4 (lateral).
template <int KIND, int NX, class T>
PGN_HD void flow_rk4(const VehParams& P, T* x, double dt, const T* u0, const T* uf, int nsub) {
    if (!(dt > 0)) return;
    const double h = dt / nsub;
    T du[4];
#pragma unroll
    for (int i = 0; i < 4; i++) du[i] = uf[i] - u0[i];
    for (int s = 0; s < nsub; s++) {
        const double fa = (s * h) / dt, fm = (s * h + h / 2) / dt, fb = (s * h + h) / dt;
        T k[NX], xt[NX], acc[NX], uc[4];
#pragma unroll
        for (int i = 0; i < 4; i++) uc[i] = u0[i] + fa * du[i];
        vehicle_model<KIND>(P, x, uc[0], uc[1], uc[2], uc[3], k);
#pragma unroll
        for (int i = 0; i < NX; i++) { acc[i] = k[i]; xt[i] = x[i] + (h / 2) * k[i]; }
#pragma unroll
        for (int i = 0; i < 4; i++) uc[i] = u0[i] + fm * du[i];
        vehicle_model<KIND>(P, xt, uc[0], uc[1], uc[2], uc[3], k);
#pragma unroll
        for (int i = 0; i < NX; i++) { acc[i] = acc[i] + 2.0 * k[i]; xt[i] = x[i] + (h / 2) * k[i]; }
        vehicle_model<KIND>(P, xt, uc[0], uc[1], uc[2], uc[3], k);
#pragma unroll
        for (int i = 0; i < NX; i++) { acc[i] = acc[i] + 2.0 * k[i]; xt[i] = x[i] + h * k[i]; }
#pragma unroll
        for (int i = 0; i < 4; i++) uc[i] = u0[i] + fb * du[i];
        vehicle_model<KIND>(P, xt, uc[0], uc[1], uc[2], uc[3], k);
#pragma unroll
        for (int i = 0; i < NX; i++) x[i] = x[i] + (h / 6) * (acc[i] + k[i]);
    }
}

// stable_limits (vehicle_dynamics.jl:227-263) -> delta_min, delta_max, H (4x2 row-major), G (4)
PGN_HD void stable_limits(const VehParams& B, double Ux, double Fxf, double Fxr, double& dmin, double& dmax, double* H, double* G) {
    double Fx = Fxf + Fxr;
    double Fzf = (B.m * B.G * B.b - B.h * Fx) / B.L;
    double Fzr = (B.m * B.G * B.a + B.h * Fx) / B.L;
    double Ff_max = B.mu * Fzf, Fr_max = B.mu * Fzr;
    double Fyf_max = fabs(Fxf) > Ff_max ? 0.0 : sqrt(Ff_max * Ff_max - Fxf * Fxf);
    double Fyr_max = fabs(Fxr) > Fr_max ? 0.0 : sqrt(Fr_max * Fr_max - Fxr * Fxr);
    double tf = 3 * Fyf_max / B.Caf, tr = 3 * Fyr_max / B.Car;
    double af = atan(tf), ar = atan(tr);
    double muG = B.mu * B.G;
    dmax = atan(B.L * muG / (Ux * Ux) - tr) + af;
    dmin = atan(B.L * (-muG) / (Ux * Ux) + tr) - af;
    double rC = muG / Ux;
    double UyC = -Ux * tr + B.b * rC;
    double rD = Ux / B.L * (tan(af + dmax) - tr);
    double UyD = Ux * tr + B.b * rD;
    double mCD = (rD - rC) / (UyD - UyC);
    double rE = Ux / B.L * (tan(-af + dmin) + tr);
    double UyE = -Ux * tr + B.b * rE;
    double rF = -muG / Ux;
    double UyF = Ux * tr + B.b * rF;
    double mEF = (rF - rE) / (UyF - UyE);
    H[0] = 1 / Ux;  H[1] = -B.b / Ux;
    H[2] = -1 / Ux; H[3] = B.b / Ux;
    H[4] = -mCD;    H[5] = 1;
    H[6] = mEF;     H[7] = -1;
    G[0] = ar; G[1] = ar; G[2] = rC - UyC * mCD; G[3] = -rF + UyF * mEF;
}

PGN_HD double clampd(double x, double lo, double hi) { return x > hi ? hi : (x < lo ? lo : x); }
PGN_HD double jmin(double a, double b) { return (a != a || b != b) ? NAN : (b < a ? b : a); }

// steady_state_estimates (vehicle_dynamics.jl:319-390)
struct SteadyState { double beta, Ux, Uy, r, A, delta, Fxf, Fxr; };
PGN_HD SteadyState steady_state_estimates(const VehParams& P, double V, double A_tan, double kappa, int num_iters, double r, double beta0,
                                          double delta0, double Fyf0) {
    const double L = P.L, a = P.a, b = P.b, h = P.h, m = P.m, Izz = P.Izz, mu = P.mu, G = P.G;
    const bool fix = P.inv_fiala_corrected != 0.0;
    double A_rad = V * V * kappa;
    double A_mag = hypot(A_tan, A_rad);
    double A_max = mu * G;
    if (A_mag > A_max) {
        if (fabs(A_rad) > A_max) { A_rad = A_max * sign_(A_rad); A_tan = 0.0; }
        else A_tan = sqrt(A_max * A_max - A_rad * A_rad) * sign_(A_tan);
    }
    double rdot = A_tan * kappa;
    double beta = beta0, delta = delta0, Fyf = Fyf0, Fxr = 0, Fxf = 0;
    for (int i = 1;; i++) {
        double sb, cb, sd, cd;
        sincos(beta, &sb, &cb);
        sincos(delta, &sd, &cd);
        double Ux = V * cb, Uy = V * sb;
        double Fx_drag = -P.Cd0 - Ux * (P.Cd1 + P.Cd2 * Ux);
        double Ax = A_tan * cb - A_rad * sb;
        double Ay = A_tan * sb + A_rad * cb;
        double Fx = Ax * m - Fx_drag;
        Fx = jmin(Fx, jmin(P.Fx_max, P.Px_max / Ux) * (P.rwd_frac + P.fwd_frac * cd) - Fyf * sd);
        double Fzr = (m * G * a + h * Fx) / L, Fzf = (m * G * b - h * Fx) / L;
        double Fr_max = mu * Fzr, Ff_max = mu * Fzf;
        Fxr = clampd((Fx + Fyf * sd) * (Fx > 0 ? P.rwd_frac / (P.rwd_frac + P.fwd_frac * cd) : P.rwb_frac / (P.rwb_frac + P.fwb_frac * cd)),
                     -Fr_max, Fr_max);
        double Fyr_max = sqrt(Fr_max * Fr_max - Fxr * Fxr);
        double Fyr = clampd((Ay * m - rdot * Izz / a) / (1 + b / a), -Fyr_max, Fyr_max);
        double tanar = inv_fiala(Fyr, P.Car, Fyr_max, fix);
        double Fxf_t = clampd(Fx - Fxr, -Ff_max, Ff_max);
        double Fyf_tmax = sqrt(Ff_max * Ff_max - Fxf_t * Fxf_t);
        double Fyf_t = clampd((b * Fyr + rdot * Izz) / a, -Fyf_tmax, Fyf_tmax);
        Fxf = Fxf_t * cd + Fyf_t * sd;
        Fyf = Fyf_t * cd - Fxf_t * sd;
        double Fyf_max = sqrt(Ff_max * Ff_max - Fxf * Fxf);
        double af = atan(inv_fiala(Fyf, P.Caf, Fyf_max, fix));
        delta = atan2(Uy + a * r, Ux) - af;
        if (i == num_iters) {
            Ax = (Fxf * cd - Fyf * sd + Fxr + Fx_drag) / m;
            Ay = (Fyf * cd + Fxf * sd + Fyr) / m;
            A_tan = Ax * cb + Ay * sb;
            break;
        }
        beta = atan(tanar + b * r / Ux);
    }
    double sb, cb;
    sincos(beta, &sb, &cb);
    SteadyState S;
    S.beta = beta; S.Ux = V * cb; S.Uy = V * sb; S.r = r; S.A = A_tan; S.delta = delta; S.Fxf = Fxf; S.Fxr = Fxr;
    return S;
}

// adiff(x, y) = wrap(x - y) into (-pi, pi]  (PigeonViz.jl:24-28)
PGN_HD double adiff(double x, double y) {
    const double twopi = 6.283185307179586476925286766559;
    double d = fmod(x - y, twopi);
    if (d == 0) d = 0.0; else if (d < 0) d += twopi;
    return d <= 3.14159265358979323846 ? d : d - twopi;
}

}  // namespace pgn
