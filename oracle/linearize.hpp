// oracle/linearize.hpp — TEST INFRASTRUCTURE ONLY (CPU oracle). PARITY UNPINNED (reference has no golden vectors).
//
// Restates what the reference obtains from two un-vendored packages (pins: /root/reference/env/Manifest.toml,
// DifferentialDynamicsModels master@0f646f28, LinearDynamicsModels master@a4aa0511):
//   * propagate(f, x, StepControl|RampControl): fixed-step RK4, PGN_RK4_SUBSTEPS = 10 sub-steps per control interval,
//     ramp control sampled at t, t+h/2, t+h    [pinned choice, see DESIGN.md]
//   * linearize(f, x, u)                      : continuous A=df/dx, B=df/du, c=f-Ax-Bu (forward-mode AD)
//   * linearize(f, x, Step/RampControl; keep) : Jacobians of the flow map (AD through the integrator)
//   * linearize(LinearDynamics, x, Step/Ramp) : exact ZOH / FOH via matrix exponentials
// Call sites: src/coupled_lat_long.jl:253,262,336,348; src/decoupled_lat_long.jl:172,182,245,253;
//             src/model_predictive_control.jl:94.
#pragma once
#include <cstring>
#include "vehicle.hpp"

namespace orc {

static const int RK4_SUBSTEPS = 10;

// Flow of VehicleModel{kind} over one control interval of length dt with control ramping linearly from up0 to upf
// (up = [delta, Fx, p1..p4]); a StepControl is the special case upf == up0.
template <class T>
inline void flow_rk4(int kind, const VehicleParams& P, const T* x0, double dt, const T* up0, const T* upf, T* xout,
                     int nsub = RK4_SUBSTEPS) {
    const int nx = model_nx(kind);
    T x[6];
    for (int i = 0; i < nx; i++) x[i] = x0[i];
    if (!(dt > 0)) { for (int i = 0; i < nx; i++) xout[i] = x[i]; return; }
    double h = dt / nsub;
    for (int s = 0; s < nsub; s++) {
        double ta = s * h, tm = ta + h / 2, tb = ta + h;
        T ua[6], um[6], ub[6];
        for (int i = 0; i < 6; i++) {
            T du = upf[i] - up0[i];
            ua[i] = up0[i] + (ta / dt) * du;
            um[i] = up0[i] + (tm / dt) * du;
            ub[i] = up0[i] + (tb / dt) * du;
        }
        T k1[6], k2[6], k3[6], k4[6], xt[6];
        vehicle_model(kind, P, x, ua, ua + 2, k1);
        for (int i = 0; i < nx; i++) xt[i] = x[i] + (h / 2) * k1[i];
        vehicle_model(kind, P, xt, um, um + 2, k2);
        for (int i = 0; i < nx; i++) xt[i] = x[i] + (h / 2) * k2[i];
        vehicle_model(kind, P, xt, um, um + 2, k3);
        for (int i = 0; i < nx; i++) xt[i] = x[i] + h * k3[i];
        vehicle_model(kind, P, xt, ub, ub + 2, k4);
        for (int i = 0; i < nx; i++) x[i] = x[i] + (h / 6) * (k1[i] + 2.0 * k2[i] + 2.0 * k3[i] + k4[i]);
    }
    for (int i = 0; i < nx; i++) xout[i] = x[i];
}

// Discrete linearization result:  x+ = A x + B0 u0[keep] + Bf uf[keep] + c   (ZOH: Bf = 0, B0 = B). Row-major.
struct DiscreteLin {
    int nx, nk;
    double A[36], B0[12], Bf[12], c[6];
};

// linearize(dyn, x, StepControl(dt,[u;p]) | RampControl(dt,[u0;p0],[uf;pf]); keep_control_dims = 1:nk) by AD through
// the RK4 flow of the nonlinear model (coupled path).
inline DiscreteLin linearize_flow(int kind, const VehicleParams& P, const double* x, double dt, const double* up0,
                                  const double* upf, bool ramp, int nk) {
    const int nx = model_nx(kind);
    typedef Dual<10> D;
    D xd[6], u0d[6], ufd[6], out[6];
    for (int i = 0; i < nx; i++) xd[i] = D::seed(x[i], i);
    for (int i = 0; i < 6; i++) { u0d[i] = D(up0[i]); ufd[i] = D(ramp ? upf[i] : up0[i]); }
    if (ramp) {
        for (int k = 0; k < nk; k++) { u0d[k] = D::seed(up0[k], 6 + k); ufd[k] = D::seed(upf[k], 8 + k); }
    } else {
        // a StepControl holds one control value: u(t) = u for the whole interval
        for (int k = 0; k < nk; k++) { u0d[k] = D::seed(up0[k], 6 + k); ufd[k] = u0d[k]; }
    }
    flow_rk4<D>(kind, P, xd, dt, u0d, ufd, out);
    DiscreteLin R;
    std::memset(&R, 0, sizeof(R));
    R.nx = nx; R.nk = nk;
    for (int i = 0; i < nx; i++) {
        for (int j = 0; j < nx; j++) R.A[i * nx + j] = out[i].d[j];
        for (int k = 0; k < nk; k++) { R.B0[i * nk + k] = out[i].d[6 + k]; R.Bf[i * nk + k] = ramp ? out[i].d[8 + k] : 0.0; }
        double c = out[i].v;
        for (int j = 0; j < nx; j++) c -= R.A[i * nx + j] * x[j];
        for (int k = 0; k < nk; k++) c -= R.B0[i * nk + k] * up0[k];
        if (ramp) for (int k = 0; k < nk; k++) c -= R.Bf[i * nk + k] * upf[k];
        R.c[i] = c;
    }
    return R;
}

// Continuous linearization linearize(dyn, x, [u;p]): A (nx x nx), B (nx x 6), f
struct ContinuousLin { int nx; double A[36], B[36], f[6]; };
inline ContinuousLin linearize_continuous(int kind, const VehicleParams& P, const double* x, const double* up) {
    const int nx = model_nx(kind);
    typedef Dual<12> D;
    D xd[6], ud[6], out[6];
    for (int i = 0; i < nx; i++) xd[i] = D::seed(x[i], i);
    for (int i = 0; i < 6; i++) ud[i] = D::seed(up[i], 6 + i);
    vehicle_model<D>(kind, P, xd, ud, ud + 2, out);
    ContinuousLin R;
    std::memset(&R, 0, sizeof(R));
    R.nx = nx;
    for (int i = 0; i < nx; i++) {
        for (int j = 0; j < nx; j++) R.A[i * nx + j] = out[i].d[j];
        for (int k = 0; k < 6; k++) R.B[i * 6 + k] = out[i].d[6 + k];
        R.f[i] = out[i].v;
    }
    return R;
}

// ---- dense matrix exponential (Pade 13 with scaling and squaring, Higham 2005), n <= 16 -------------------------
inline void mat_mul(int n, const double* A, const double* B, double* C) {
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) { double s = 0; for (int k = 0; k < n; k++) s += A[i * n + k] * B[k * n + j]; C[i * n + j] = s; }
}
inline bool mat_solve(int n, double* A, double* B, int nrhs) {  // Gaussian elimination with partial pivoting, in place
    for (int k = 0; k < n; k++) {
        int piv = k; double best = std::fabs(A[k * n + k]);
        for (int i = k + 1; i < n; i++) if (std::fabs(A[i * n + k]) > best) { best = std::fabs(A[i * n + k]); piv = i; }
        if (best == 0) return false;
        if (piv != k) {
            for (int j = 0; j < n; j++) std::swap(A[k * n + j], A[piv * n + j]);
            for (int j = 0; j < nrhs; j++) std::swap(B[k * nrhs + j], B[piv * nrhs + j]);
        }
        for (int i = k + 1; i < n; i++) {
            double f = A[i * n + k] / A[k * n + k];
            if (f == 0) continue;
            for (int j = k; j < n; j++) A[i * n + j] -= f * A[k * n + j];
            for (int j = 0; j < nrhs; j++) B[i * nrhs + j] -= f * B[k * nrhs + j];
        }
    }
    for (int k = n - 1; k >= 0; k--) {
        for (int j = 0; j < nrhs; j++) {
            double s = B[k * nrhs + j];
            for (int i = k + 1; i < n; i++) s -= A[k * n + i] * B[i * nrhs + j];
            B[k * nrhs + j] = s / A[k * n + k];
        }
    }
    return true;
}
inline void expm(int n, const double* Ain, double* E) {
    static const double b[14] = {64764752532480000., 32382376266240000., 7771770303897600., 1187353796428800.,
                                 129060195264000.,   10559470521600.,    670442572800.,     33522128640.,
                                 1323241920.,        40840800.,          960960.,           16380., 182., 1.};
    const int nn = n * n;
    double A[256], A2[256], A4[256], A6[256], U[256], V[256], T1[256], T2[256];
    double norm1 = 0;
    for (int j = 0; j < n; j++) { double s = 0; for (int i = 0; i < n; i++) s += std::fabs(Ain[i * n + j]); if (s > norm1) norm1 = s; }
    int sq = 0;
    const double theta13 = 5.371920351148152;
    if (norm1 > theta13) { sq = (int)std::ceil(std::log2(norm1 / theta13)); if (sq < 0) sq = 0; }
    double scale = std::ldexp(1.0, -sq);
    for (int i = 0; i < nn; i++) A[i] = Ain[i] * scale;
    mat_mul(n, A, A, A2); mat_mul(n, A2, A2, A4); mat_mul(n, A4, A2, A6);
    // U = A (A6 (b13 A6 + b11 A4 + b9 A2) + b7 A6 + b5 A4 + b3 A2 + b1 I)
    for (int i = 0; i < nn; i++) T1[i] = b[13] * A6[i] + b[11] * A4[i] + b[9] * A2[i];
    mat_mul(n, A6, T1, T2);
    for (int i = 0; i < nn; i++) T2[i] += b[7] * A6[i] + b[5] * A4[i] + b[3] * A2[i];
    for (int i = 0; i < n; i++) T2[i * n + i] += b[1];
    mat_mul(n, A, T2, U);
    // V = A6 (b12 A6 + b10 A4 + b8 A2) + b6 A6 + b4 A4 + b2 A2 + b0 I
    for (int i = 0; i < nn; i++) T1[i] = b[12] * A6[i] + b[10] * A4[i] + b[8] * A2[i];
    mat_mul(n, A6, T1, V);
    for (int i = 0; i < nn; i++) V[i] += b[6] * A6[i] + b[4] * A4[i] + b[2] * A2[i];
    for (int i = 0; i < n; i++) V[i * n + i] += b[0];
    // (V - U) E = (V + U)
    for (int i = 0; i < nn; i++) { T1[i] = V[i] - U[i]; E[i] = V[i] + U[i]; }
    mat_solve(n, T1, E, n);
    for (int s = 0; s < sq; s++) { mat_mul(n, E, E, T1); std::memcpy(E, T1, sizeof(double) * nn); }
}

// Exact discretization of the LinearDynamics  xdot = A x + B up + c0  (c0 = f - A x0 - B up0), linearized about x0
// (decoupled path: linearize(linearize(dyn, x, up), x, Step/RampControl; keep = 1:nk)).
inline DiscreteLin linearize_exact(const ContinuousLin& CL, const double* x, double dt, const double* up0, const double* upf,
                                   bool ramp, int nk) {
    const int nx = CL.nx;
    // M = dt * [[A, I, 0], [0, 0, I/dt... ]]: use the standard triple-block form
    //   M = [[A*dt, I*dt, 0], [0, 0, I], [0, 0, 0]]  =>  expm(M) = [[Phi, G1, G2], ...]
    //   with G1 = int_0^dt e^{As} ds,  G2 = int_0^dt e^{A(dt-s)} (s/dt) ds
    const int n3 = 3 * nx;
    double M[256] = {0}, EM[256];
    for (int i = 0; i < nx; i++) {
        for (int j = 0; j < nx; j++) M[i * n3 + j] = CL.A[i * nx + j] * dt;
        M[i * n3 + nx + i] = dt;
        M[(nx + i) * n3 + 2 * nx + i] = 1.0;
    }
    expm(n3, M, EM);
    double Phi[16 + 20], G1[36], G2[36];
    for (int i = 0; i < nx; i++) for (int j = 0; j < nx; j++) {
        Phi[i * nx + j] = EM[i * n3 + j];
        G1[i * nx + j] = EM[i * n3 + nx + j];
        G2[i * nx + j] = EM[i * n3 + 2 * nx + j];
    }
    DiscreteLin R;
    std::memset(&R, 0, sizeof(R));
    R.nx = nx; R.nk = nk;
    // g0 = B up0 + c0 = f - A x0 ; dB = B (upf - up0)
    double g0[6], dB[6];
    for (int i = 0; i < nx; i++) {
        double s = CL.f[i];
        for (int j = 0; j < nx; j++) s -= CL.A[i * nx + j] * x[j];
        g0[i] = s;
        double d = 0;
        if (ramp) for (int k = 0; k < 6; k++) d += CL.B[i * 6 + k] * (upf[k] - up0[k]);
        dB[i] = d;
    }
    for (int i = 0; i < nx; i++) {
        double xp = 0;
        for (int j = 0; j < nx; j++) { R.A[i * nx + j] = Phi[i * nx + j]; xp += Phi[i * nx + j] * x[j] + G1[i * nx + j] * g0[j] + G2[i * nx + j] * dB[j]; }
        for (int k = 0; k < nk; k++) {
            double g1b = 0, g2b = 0;
            for (int j = 0; j < nx; j++) { g1b += G1[i * nx + j] * CL.B[j * 6 + k]; g2b += G2[i * nx + j] * CL.B[j * 6 + k]; }
            if (ramp) { R.B0[i * nk + k] = g1b - g2b; R.Bf[i * nk + k] = g2b; }
            else      { R.B0[i * nk + k] = g1b;       R.Bf[i * nk + k] = 0; }
        }
        double c = xp;
        for (int j = 0; j < nx; j++) c -= R.A[i * nx + j] * x[j];
        for (int k = 0; k < nk; k++) c -= R.B0[i * nk + k] * up0[k];
        if (ramp) for (int k = 0; k < nk; k++) c -= R.Bf[i * nk + k] * upf[k];
        R.c[i] = c;
    }
    return R;
}

}  // namespace orc
