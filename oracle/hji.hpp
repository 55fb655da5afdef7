// oracle/hji.hpp — TEST INFRASTRUCTURE ONLY (CPU oracle). PARITY UNPINNED (reference has no golden vectors; the real
// grid file deps/BicycleCAvoid.jld2 is a network download, /root/reference/deps/build.jl:1-4, and is absent).
//
// CPU restatement of /root/reference/src/HJI_computation.jl:20-24 (HJIRelativeState), :26-37,66-72 (HJICache lookup),
// :74-88 (relative_dynamics), :90-131 (optimal_disturbance), :133-158 (optimal_control), :160-170 (compute_reachability_constraint).
// Interpolations.jl 0.11.2 `Gridded(Linear())` on Float32 tables with Float64 queries (not vendored) is restated as
// per-dimension knot search + nested (dimension-1-innermost) linear interpolation with Float64 weights.
#pragma once
#include <cmath>
#include <vector>
#include "linearize.hpp"
#include "trajectory.hpp"

namespace orc {

struct HjiCache {
    int dims[7];
    std::vector<float> knots[7];
    std::vector<float> V;      // column-major, dimension 1 fastest
    std::vector<float> gradV;  // 7 components fastest, then the grid (SVector{7,Float32} per node)
    size_t n_nodes() const { size_t s = 1; for (int d = 0; d < 7; d++) s *= dims[d]; return s; }
};

// placeholder_HJICache (HJI_computation.jl:32-37)
inline HjiCache placeholder_hji() {
    HjiCache C;
    for (int d = 0; d < 7; d++) { C.dims[d] = 2; C.knots[d] = {-1000.f, 1000.f}; }
    C.V.assign(128, 0.f); C.gradV.assign(128 * 7, 0.f);
    return C;
}

// HJIRelativeState(us, them) (HJI_computation.jl:20-24): note `cψ, sψ = sincos(-ψ)` binds cψ <- sin(-ψ), sψ <- cos(-ψ)
inline void hji_relative_state(const double* us6, const double* them4, double* x7) {
    double cpsi = std::sin(-us6[2]), spsi = std::cos(-us6[2]);
    double dE = them4[0] - us6[0], dN = them4[1] - us6[1];
    x7[0] = cpsi * dE + spsi * dN;
    x7[1] = -spsi * dE + cpsi * dN;
    x7[2] = adiff(them4[2], us6[2]);
    x7[3] = us6[3]; x7[4] = us6[4]; x7[5] = them4[3]; x7[6] = us6[5];
}

// cache[x] (HJI_computation.jl:66-72): returns false (V=Inf, gradV=0) outside the grid
inline bool hji_lookup(const HjiCache& C, const double* x7, double& V, double* gV7) {
    for (int d = 0; d < 7; d++) {
        if (!((double)C.knots[d].front() <= x7[d] && x7[d] <= (double)C.knots[d].back())) {
            V = INFINITY; for (int k = 0; k < 7; k++) gV7[k] = 0; return false;
        }
    }
    int idx[7]; double w[7];
    for (int d = 0; d < 7; d++) {
        const std::vector<float>& k = C.knots[d];
        int lo = 0, hi = (int)k.size();
        while (lo < hi) { int mid = (lo + hi) / 2; if ((double)k[mid] <= x7[d]) lo = mid + 1; else hi = mid; }
        int i = lo;  // searchsortedlast (1-based)
        if (i < 1) i = 1; if (i > (int)k.size() - 1) i = (int)k.size() - 1;
        idx[d] = i - 1;
        w[d] = (x7[d] - (double)k[i - 1]) / ((double)k[i] - (double)k[i - 1]);
    }
    size_t stride[7]; stride[0] = 1;
    for (int d = 1; d < 7; d++) stride[d] = stride[d - 1] * C.dims[d - 1];
    double acc[128][8];
    for (int cnr = 0; cnr < 128; cnr++) {
        size_t off = 0;
        for (int d = 0; d < 7; d++) off += (size_t)(idx[d] + ((cnr >> d) & 1)) * stride[d];
        acc[cnr][0] = (double)C.V[off];
        for (int k = 0; k < 7; k++) acc[cnr][1 + k] = (double)C.gradV[off * 7 + k];
    }
    int cnt = 128;
    for (int d = 0; d < 7; d++) {     // dimension 1 innermost
        cnt /= 2;
        for (int cnr = 0; cnr < cnt; cnr++)
            for (int k = 0; k < 8; k++) acc[cnr][k] = (1 - w[d]) * acc[2 * cnr][k] + w[d] * acc[2 * cnr + 1][k];
    }
    V = acc[0][0];
    for (int k = 0; k < 7; k++) gV7[k] = acc[0][1 + k];
    return true;
}

// relative_dynamics (HJI_computation.jl:74-88), generic in the robot control for AD
template <class T>
inline void relative_dynamics(const VehicleParams& P, const double* x7, const T* uR2, const double* uH2, T* out7) {
    T q[6] = {T(x7[0]), T(x7[1]), T(x7[2]), T(x7[3]), T(x7[4]), T(x7[6])};
    T p0[4] = {T(0.0), T(0.0), T(0.0), T(0.0)};
    T bd[6];
    vehicle_model<T>(MODEL_BICYCLE, P, q, uR2, p0, bd);
    double s = std::sin(x7[2]), c = std::cos(x7[2]);
    out7[0] = T(x7[5] * c - x7[3] + x7[1] * x7[6]);
    out7[1] = T(x7[5] * s - x7[4] - x7[0] * x7[6]);
    out7[2] = T(uH2[0] - x7[6]);
    out7[3] = bd[3];
    out7[4] = bd[4];
    out7[5] = T(uH2[1]);
    out7[6] = bd[5];
}

// optimal_disturbance (HJI_computation.jl:90-131), dMode = :min.
// Deviation: the reference divides by the other car's speed and yields NaN for V = 0; here V <= 0
// returns (0, 0).
inline void optimal_disturbance(const VehicleParams& P, const double* x7, const double* gV7, double* uH2) {
    double Ax_max = P.Fx_max / P.m, Pmx_max = P.Px_max / P.m, maxA = 0.9 * P.mu * P.G;
    double sgn = -1;
    double V = x7[5];
    if (!(V > 0)) { uH2[0] = 0; uH2[1] = 0; return; }
    double lam_w = gV7[2], lam_Ax = gV7[5];
    double lam_Ay = lam_w / V;
    double lam_norm = std::hypot(lam_Ax, lam_Ay);
    if (lam_norm < 1e-3) { uH2[0] = 0; uH2[1] = 0; return; }
    double desAx = sgn * lam_Ax * maxA / lam_norm;
    double desAy = sgn * lam_Ay * maxA / lam_norm;
    double maxAx = std::min(Ax_max, Pmx_max / V);
    double maxAy = P.kappa_max * V * V;
    if (desAx > maxAx) {
        if (std::fabs(desAy) < maxAy) maxAy = std::min(maxAy, std::sqrt(maxA * maxA - maxAx * maxAx));
        uH2[0] = std::copysign(maxAy, desAy) / V; uH2[1] = maxAx; return;
    } else {
        if (std::fabs(desAy) > maxAy) {
            if (desAx > 0) {
                maxAx = std::min(std::sqrt(maxA * maxA - maxAy * maxAy), maxAx);
                uH2[0] = std::copysign(maxAy, desAy) / V; uH2[1] = maxAx; return;
            } else {
                uH2[0] = std::copysign(maxAy, desAy) / V; uH2[1] = -std::sqrt(maxA * maxA - maxAy * maxAy); return;
            }
        } else {
            uH2[0] = desAy / V; uH2[1] = maxAx; return;
        }
    }
}

// optimal_control (HJI_computation.jl:133-158), uMode = :max, N = 50: the "hammer" policy of the callback (ros_integration.jl:115-118).
// delta at the limit chosen by the sign of B = gV5/m + a gV7/Izz, Fx by a 50-point grid search of A Fx + B Fyf + C Fyr over [Fx_min, Fx_max]
// (first maximum wins: strict >), tire forces from lateral_tire_forces(BM, fake_qR, uR) with the raw drive/brake split (no control limits).
inline void optimal_control(const VehicleParams& P, const double* x7, const double* gV7, double* uR2, int N = 50) {
    double fake_q[6] = {0.0, 0.0, 0.0, x7[3], x7[4], x7[6]};
    double A = gV7[3] / P.m;
    double B = gV7[4] / P.m + P.a * gV7[6] / P.Izz;
    double C = gV7[4] / P.m - P.b * gV7[6] / P.Izz;
    double d_opt = (B >= 0) ? P.delta_max : -P.delta_max;
    double V_opt = -INFINITY, Fx_opt = 0.0;
    for (int n = 0; n < N; n++) {
        double frac = (double)n / (double)(N - 1);
        double Fx = frac * P.Fx_max + (1 - frac) * P.Fx_min;
        double u3[3] = {d_opt, 0, 0};
        longitudinal_tire_forces<double>(P, Fx, u3[1], u3[2]);
        double Fyf, Fyr;
        lateral_tire_forces_qu(P, fake_q, u3, Fyf, Fyr);
        double V = A * Fx + B * Fyf + C * Fyr;
        if (V > V_opt) { Fx_opt = Fx; V_opt = V; }
    }
    uR2[0] = d_opt; uR2[1] = Fx_opt;
}

// compute_reachability_constraint (HJI_computation.jl:160-170) with uR_lin = BicycleControl2(current_control)
inline void reachability_constraint(const VehicleParams& P, const HjiCache& C, const double* x7, double eps, const double* uR2,
                                    double* M2, double& b, double* V_out = nullptr, double* gV_out = nullptr) {
    double V, gV[7];
    hji_lookup(C, x7, V, gV);
    if (V_out) *V_out = V;
    if (gV_out) for (int k = 0; k < 7; k++) gV_out[k] = gV[k];
    if (V > eps) { M2[0] = 0; M2[1] = 0; b = 1.0; return; }
    double uH[2];
    optimal_disturbance(P, x7, gV, uH);
    typedef Dual<2> D;
    D u[2] = {D::seed(uR2[0], 0), D::seed(uR2[1], 1)}, f[7];
    relative_dynamics<D>(P, x7, u, uH, f);
    D H(0.0);
    for (int k = 0; k < 7; k++) H = H + gV[k] * f[k];
    M2[0] = H.d[0]; M2[1] = H.d[1];
    b = H.v - (M2[0] * uR2[0] + M2[1] * uR2[1]);
}

}  // namespace orc
