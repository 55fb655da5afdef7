// oracle/vehicle.hpp — TEST INFRASTRUCTURE ONLY (CPU oracle). PARITY UNPINNED (reference has no golden vectors).
//
// CPU restatement of /root/reference/src/vehicles.jl:1-59 (X1 parameters) and
// /root/reference/src/vehicle_dynamics.jl:35-390 (Fiala tire, weight-transfer fixed point, the three bicycle
// models, stability envelope, actuation split / limits, VehicleModel wrapper, steady_state_estimates).
// Generic over the scalar type T (double or orc::Dual<N>) so the same code is differentiated exactly like the
// reference's ForwardDiff pass.
#pragma once
#include <cmath>
#include <limits>
#include "dual.hpp"

namespace orc {

// ---- parameter bundles (vehicle_dynamics.jl:7-28, 272-291) -------------------------------------------------
struct VehicleParams {
    // BicycleModelParams
    double L, a, b, h, G, m, Izz, mu, Caf, Car, Cd0, Cd1, Cd2;
    // LongitudinalActuationParams
    double fwd_frac, rwd_frac, fwb_frac, rwb_frac;
    // ControlLimits
    double Fx_max, Fx_min, Px_max, delta_max, kappa_max;
    // 0 (default): _invfialatiremodel restated literally (vehicle_dynamics.jl:56-62 returns the slip *ratio* in its
    // unsaturated branch); 1: multiply by tan(alpha_slide) = 3 Fy_max / C_alpha (the evident intent). See DESIGN.md.
    double inv_fiala_corrected;
};
static const int VEHICLE_PARAMS_LEN = 23;

// vehicles.jl:1-59
inline VehicleParams X1() {
    VehicleParams P;
    P.G = 9.80665;
    double mfl = 484, mfr = 455, mrl = 521, mrr = 504;
    P.m = mfl + mfr + mrl + mrr;
    P.Izz = 2900;
    P.L = 2.87;
    P.a = (mrl + mrr) / P.m * P.L;
    P.b = (mfl + mfr) / P.m * P.L;
    double hf = 0.1, hr = 0.1, h1 = 0.37;
    P.h = hf * P.b / P.L + hr * P.a / P.L + h1;
    P.mu = 0.92;
    P.Caf = 150e3;
    P.Car = 220e3;
    P.Fx_max = 5600;
    P.Px_max = 75e3;
    P.Cd0 = 241.0;
    P.Cd1 = 25.1;
    P.Cd2 = 0.0;
    P.fwd_frac = 0.0;
    P.rwd_frac = 1 - P.fwd_frac;
    P.fwb_frac = 0.6;
    P.rwb_frac = 1 - P.fwb_frac;
    double f1 = -P.m * P.G * P.a * P.mu / (P.L * P.rwb_frac + P.mu * P.h);
    double f2 = -P.m * P.G * P.b * P.mu / (P.L * P.fwb_frac - P.mu * P.h);
    P.Fx_min = f1 > f2 ? f1 : f2;
    P.delta_max = 18 * M_PI / 180;
    P.kappa_max = std::tan(P.delta_max) / P.L;
    P.inv_fiala_corrected = 0.0;
    return P;
}

// ---- Fiala tire (vehicle_dynamics.jl:35-62) ------------------------------------------------------------------
template <class T>
inline T _fialatiremodel(const T& tana, double Ca, const T& Fy_max) {
    T tana_slide = 3.0 * Fy_max / Ca;
    T ratio = absd(tana / tana_slide);
    if (value(ratio) <= 1) {
        return -Ca * tana * (1.0 - ratio + ratio * ratio / 3.0);
    } else {
        return -Fy_max * signd(tana);
    }
}
template <class T>
inline T fialatiremodel(const T& alpha, double Ca, double mu, const T& Fx, const T& Fz) {
    T F_max = mu * Fz;
    if (value(absd(Fx)) >= value(F_max)) return T(0.0);
    return _fialatiremodel(tan(alpha), Ca, sqrt(F_max * F_max - Fx * Fx));
}
inline double _invfialatiremodel(double Fy, double Ca, double Fy_max, bool corrected = false) {  // returns tan(alpha)
    if (std::fabs(Fy) >= Fy_max) return -(3 * Fy_max / Ca) * signd(Fy);
    // NB restated literally: the reference returns the slip *ratio* here (no 3*Fy_max/Ca factor), vehicle_dynamics.jl:60
    double r = -(1 + std::cbrt(std::fabs(Fy) / Fy_max - 1)) * signd(Fy);
    return corrected ? r * (3 * Fy_max / Ca) : r;
}

// lateral_tire_forces (vehicle_dynamics.jl:64-76)
template <class T>
inline void lateral_tire_forces(const VehicleParams& B, const T& af, const T& ar, const T& Fxf, const T& Fxr,
                                const T& sd, const T& cd, T& Fyf, T& Fyr, int num_iters = 3) {
    Fyf = T(0.0);
    T Fx = Fxf * cd - Fyf * sd + Fxr;
    for (int i = 0; i < num_iters; i++) {
        T Fzf = (B.m * B.G * B.b - B.h * Fx) / B.L;
        Fyf = fialatiremodel(af, B.Caf, B.mu, Fxf, Fzf);
        Fx = Fxf * cd - Fyf * sd + Fxr;
    }
    T Fzr = (B.m * B.G * B.a + B.h * Fx) / B.L;
    Fyr = fialatiremodel(ar, B.Car, B.mu, Fxr, Fzr);
}
// lateral_tire_forces(B, q::6, u::3) (vehicle_dynamics.jl:78-87)
inline void lateral_tire_forces_qu(const VehicleParams& B, const double* q6, const double* u3, double& Fyf, double& Fyr,
                                   int num_iters = 3) {
    double Ux = q6[3], Uy = q6[4], r = q6[5];
    double d = u3[0], Fxf = u3[1], Fxr = u3[2];
    double sd = std::sin(d), cd = std::cos(d);
    double af = std::atan2(Uy + B.a * r, Ux) - d;
    double ar = std::atan2(Uy - B.b * r, Ux);
    lateral_tire_forces<double>(B, af, ar, Fxf, Fxr, sd, cd, Fyf, Fyr, num_iters);
}

// ---- bicycle models ---------------------------------------------------------------------------------------------
enum ModelKind { MODEL_BICYCLE = 0, MODEL_TRACKING = 1, MODEL_LATERAL = 2 };

// BicycleModel (vehicle_dynamics.jl:111-134): q=(E,N,psi,Ux,Uy,r), u=(delta,Fxf,Fxr)
template <class T>
inline void bicycle_model(const VehicleParams& B, const T* q, const T* u, T* out) {
    const T &psi = q[2], &Ux = q[3], &Uy = q[4], &r = q[5];
    const T &d = u[0], &Fxf = u[1], &Fxr = u[2];
    T sp = sin(psi), cp = cos(psi), sd = sin(d), cd = cos(d);
    T af = atan2(Uy + B.a * r, Ux) - d;
    T ar = atan2(Uy - B.b * r, Ux);
    T Fyf, Fyr;
    lateral_tire_forces(B, af, ar, Fxf, Fxr, sd, cd, Fyf, Fyr);
    T Fx_drag = -B.Cd0 - Ux * (B.Cd1 + B.Cd2 * Ux);
    T Fxf_t = Fxf * cd - Fyf * sd;
    T Fyf_t = Fyf * cd + Fxf * sd;
    out[0] = -Ux * sp - Uy * cp;
    out[1] = Ux * cp - Uy * sp;
    out[2] = r;
    out[3] = (Fxf_t + Fxr + Fx_drag) / B.m + r * Uy;
    out[4] = (Fyf_t + Fyr) / B.m - r * Ux;
    out[5] = (B.a * Fyf_t - B.b * Fyr) / B.Izz;
}

// TrackingBicycleModel (vehicle_dynamics.jl:159-182): q=(ds,Ux,Uy,r,dpsi,e), u=(delta,Fxf,Fxr), p=(V,kappa,theta,phi)
template <class T>
inline void tracking_model(const VehicleParams& B, const T* q, const T* u, const T* p, T* out) {
    const T &Ux = q[1], &Uy = q[2], &r = q[3], &dpsi = q[4];
    const T &d = u[0], &Fxf = u[1], &Fxr = u[2];
    const T &V = p[0], &kappa = p[1];
    T sp = sin(dpsi), cp = cos(dpsi), sd = sin(d), cd = cos(d);
    T af = atan2(Uy + B.a * r, Ux) - d;
    T ar = atan2(Uy - B.b * r, Ux);
    T Fyf, Fyr;
    lateral_tire_forces(B, af, ar, Fxf, Fxr, sd, cd, Fyf, Fyr);
    T Fx_drag = -B.Cd0 - Ux * (B.Cd1 + B.Cd2 * Ux);
    T Fxf_t = Fxf * cd - Fyf * sd;
    T Fyf_t = Fyf * cd + Fxf * sd;
    out[0] = Ux * cp - Uy * sp - V;
    out[1] = (Fxf_t + Fxr + Fx_drag) / B.m + r * Uy;
    out[2] = (Fyf_t + Fyr) / B.m - r * Ux;
    out[3] = (B.a * Fyf_t - B.b * Fyr) / B.Izz;
    out[4] = r - (Ux * cp - Uy * sp) * kappa;
    out[5] = Ux * sp + Uy * cp;
}

// LateralTrackingBicycleModel (vehicle_dynamics.jl:205-223): q=(Uy,r,dpsi,e), u=(delta,Fxf,Fxr), p=(Ux,kappa,theta,phi)
template <class T>
inline void lateral_model(const VehicleParams& B, const T* q, const T* u, const T* p, T* out) {
    const T &Uy = q[0], &r = q[1], &dpsi = q[2];
    const T &d = u[0], &Fxf = u[1], &Fxr = u[2];
    const T &Ux = p[0], &kappa = p[1];
    T sp = sin(dpsi), cp = cos(dpsi), sd = sin(d), cd = cos(d);
    T af = atan2(Uy + B.a * r, Ux) - d;
    T ar = atan2(Uy - B.b * r, Ux);
    T Fyf, Fyr;
    lateral_tire_forces(B, af, ar, Fxf, Fxr, sd, cd, Fyf, Fyr);
    T Fyf_t = Fyf * cd + Fxf * sd;
    out[0] = (Fyf_t + Fyr) / B.m - r * Ux;
    out[1] = (B.a * Fyf_t - B.b * Fyr) / B.Izz;
    out[2] = r - Ux * kappa;
    out[3] = Ux * sp + Uy * cp;
}

// ---- actuation (vehicle_dynamics.jl:279-298) ---------------------------------------------------------------------
template <class T>
inline void longitudinal_tire_forces(const VehicleParams& P, const T& Fx, T& Fxf, T& Fxr) {
    if (value(Fx) > 0) { Fxf = Fx * P.fwd_frac; Fxr = Fx * P.rwd_frac; }
    else               { Fxf = Fx * P.fwb_frac; Fxr = Fx * P.rwb_frac; }
}
template <class T>
inline void apply_control_limits(const VehicleParams& P, const T& delta, const T& Fx, const T& Ux_in, T& d_out, T& Fx_out) {
    double Ux = value(Ux_in);  // ForwardDiff.value(Ux)  (vehicle_dynamics.jl:295)
    d_out = jl_clamp(delta, -P.delta_max, P.delta_max);
    T t1 = jl_min(Fx, T(P.Fx_max));
    T t2 = jl_min(t1, T(P.Px_max / Ux));
    Fx_out = jl_max(t2, T(P.Fx_min));
}

// VehicleModel call (vehicle_dynamics.jl:307-316): u2=(delta,Fx); state size 6 (bicycle/tracking) or 4 (lateral)
template <class T>
inline void vehicle_model(int kind, const VehicleParams& P, const T* q, const T* u2, const T* p4, T* out) {
    T Ux = (kind == MODEL_BICYCLE) ? q[3] : (kind == MODEL_TRACKING ? q[1] : p4[0]);
    T dl, Fx;
    apply_control_limits(P, u2[0], u2[1], Ux, dl, Fx);
    T u3[3];
    u3[0] = dl;
    longitudinal_tire_forces(P, Fx, u3[1], u3[2]);
    if (kind == MODEL_BICYCLE) bicycle_model(P, q, u3, out);
    else if (kind == MODEL_TRACKING) tracking_model(P, q, u3, p4, out);
    else lateral_model(P, q, u3, p4, out);
}
inline int model_nx(int kind) { return kind == MODEL_LATERAL ? 4 : 6; }

// ---- stability envelope (vehicle_dynamics.jl:227-263) -------------------------------------------------------------
struct StableLimits { double delta_min, delta_max; double H[4][2]; double G[4]; };
inline StableLimits stable_limits(const VehicleParams& B, double Ux, double Fxf, double Fxr) {
    double Fx = Fxf + Fxr;
    double Fy_grade = 0;
    double Fzf = (B.m * B.G * B.b - B.h * Fx) / B.L;
    double Fzr = (B.m * B.G * B.a + B.h * Fx) / B.L;
    double Ff_max = B.mu * Fzf, Fr_max = B.mu * Fzr;
    double Fyf_max = std::fabs(Fxf) > Ff_max ? 0.0 : std::sqrt(Ff_max * Ff_max - Fxf * Fxf);
    double Fyr_max = std::fabs(Fxr) > Fr_max ? 0.0 : std::sqrt(Fr_max * Fr_max - Fxr * Fxr);
    double tanaf_slide = 3 * Fyf_max / B.Caf;
    double tanar_slide = 3 * Fyr_max / B.Car;
    double af_slide = std::atan(tanaf_slide);
    double ar_slide = std::atan(tanar_slide);
    StableLimits S;
    S.delta_max = std::atan(B.L * (B.mu * B.G + Fy_grade / B.m) / (Ux * Ux) - tanar_slide) + af_slide;
    S.delta_min = std::atan(B.L * (-B.mu * B.G + Fy_grade / B.m) / (Ux * Ux) + tanar_slide) - af_slide;
    double rC = (B.mu * B.G + Fy_grade / B.m) / Ux;
    double UyC = -Ux * tanar_slide + B.b * rC;
    double rD = Ux / B.L * (std::tan(af_slide + S.delta_max) - tanar_slide);
    double UyD = Ux * tanar_slide + B.b * rD;
    double mCD = (rD - rC) / (UyD - UyC);
    double rE = Ux / B.L * (std::tan(-af_slide + S.delta_min) + tanar_slide);
    double UyE = -Ux * tanar_slide + B.b * rE;
    double rF = (-B.mu * B.G + Fy_grade / B.m) / Ux;
    double UyF = Ux * tanar_slide + B.b * rF;
    double mEF = (rF - rE) / (UyF - UyE);
    S.H[0][0] = 1 / Ux;  S.H[0][1] = -B.b / Ux;
    S.H[1][0] = -1 / Ux; S.H[1][1] = B.b / Ux;
    S.H[2][0] = -mCD;    S.H[2][1] = 1;
    S.H[3][0] = mEF;     S.H[3][1] = -1;
    S.G[0] = ar_slide; S.G[1] = ar_slide; S.G[2] = rC - UyC * mCD; S.G[3] = -rF + UyF * mEF;
    return S;
}

// ---- steady-state estimates (vehicle_dynamics.jl:319-390) ---------------------------------------------------------
struct SteadyState { double beta, Ux, Uy, r, A, delta, Fxf, Fxr; };
inline double clampd(double x, double lo, double hi) { return x > hi ? hi : (x < lo ? lo : x); }
inline double mind(double a, double b) { return (a != a || b != b) ? std::numeric_limits<double>::quiet_NaN() : (b < a ? b : a); }
inline SteadyState steady_state_estimates(const VehicleParams& P, double V, double A_tan, double kappa, int num_iters,
                                          double r, double beta0, double delta0, double Fyf0) {
    const double L = P.L, a = P.a, b = P.b, h = P.h, m = P.m, Izz = P.Izz, mu = P.mu, G = P.G;
    double A_rad = V * V * kappa;
    double A_mag = std::hypot(A_tan, A_rad);
    double A_max = mu * G;
    if (A_mag > A_max) {
        if (std::fabs(A_rad) > A_max) {
            A_rad = A_max * signd(A_rad);
            A_tan = 0.0;
        } else {
            A_tan = std::sqrt(A_max * A_max - A_rad * A_rad) * signd(A_tan);
        }
    }
    double rdot = A_tan * kappa;
    int i = 1;
    double beta = beta0, delta = delta0, Fyf = Fyf0;
    double Ux = 0, Uy = 0, Fxr = 0, Fxf = 0;
    while (true) {
        double sb = std::sin(beta), cb = std::cos(beta);
        double sd = std::sin(delta), cd = std::cos(delta);
        Ux = V * cb; Uy = V * sb;
        double Fx_drag = -P.Cd0 - Ux * (P.Cd1 + P.Cd2 * Ux);
        double Fx_grade = 0, Fy_grade = 0;
        double Ax = A_tan * cb - A_rad * sb;
        double Ay = A_tan * sb + A_rad * cb;
        double Fx = Ax * m - Fx_drag - Fx_grade;
        Fx = mind(Fx, mind(P.Fx_max, P.Px_max / Ux) * (P.rwd_frac + P.fwd_frac * cd) - Fyf * sd);
        double Fzr = (m * G * a + h * Fx) / L, Fzf = (m * G * b - h * Fx) / L;
        double Fr_max = mu * Fzr, Ff_max = mu * Fzf;
        Fxr = clampd((Fx + Fyf * sd) * (Fx > 0 ? P.rwd_frac / (P.rwd_frac + P.fwd_frac * cd)
                                               : P.rwb_frac / (P.rwb_frac + P.fwb_frac * cd)),
                     -Fr_max, Fr_max);
        double Fyr_max = std::sqrt(Fr_max * Fr_max - Fxr * Fxr);
        double Fyr = (Ay * m - Fy_grade - rdot * Izz / a) / (1 + b / a);
        Fyr = clampd(Fyr, -Fyr_max, Fyr_max);
        double tanar = _invfialatiremodel(Fyr, P.Car, Fyr_max, P.inv_fiala_corrected != 0);
        double Fxf_t = clampd(Fx - Fxr, -Ff_max, Ff_max);
        double Fyf_tmax = std::sqrt(Ff_max * Ff_max - Fxf_t * Fxf_t);
        double Fyf_t = clampd((b * Fyr + rdot * Izz) / a, -Fyf_tmax, Fyf_tmax);
        Fxf = Fxf_t * cd + Fyf_t * sd;
        Fyf = Fyf_t * cd - Fxf_t * sd;
        double Fyf_max = std::sqrt(Ff_max * Ff_max - Fxf * Fxf);
        double af = std::atan(_invfialatiremodel(Fyf, P.Caf, Fyf_max, P.inv_fiala_corrected != 0));
        delta = std::atan2(Uy + a * r, Ux) - af;
        if (i == num_iters) {
            Ax = (Fxf * cd - Fyf * sd + Fxr + Fx_drag + Fx_grade) / m;
            Ay = (Fyf * cd + Fxf * sd + Fyr + Fy_grade) / m;
            A_tan = Ax * cb + Ay * sb;
            break;
        }
        i = i + 1;
        beta = std::atan(tanar + b * r / Ux);
    }
    double sb = std::sin(beta), cb = std::cos(beta);
    SteadyState S;
    S.beta = beta; S.Ux = V * cb; S.Uy = V * sb; S.r = r; S.A = A_tan; S.delta = delta; S.Fxf = Fxf; S.Fxr = Fxr;
    return S;
}

}  // namespace orc
