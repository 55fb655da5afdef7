// pgn_linearize.cu — update_QP!: per-interval discrete linearisation + stability envelope, written into the QP piece record.
//   coupled   (reference src/coupled_lat_long.jl:335-367): linearize(dynamics, q_t, StepControl|RampControl) = Jacobians of the
//             RK4 flow of VehicleModel{TrackingBicycleModel} by forward-mode AD *through the integrator* (LinearDynamicsModels,
//             not vendored).  One node is handled by a group of 4 adjacent lanes; each lane carries the primal and 2 of the 8
//             tangent directions (Ux,Uy,r,dpsi | delta0,Fx0 | deltaf,Fxf) in registers.  The ds and e columns of A are exact unit
//             vectors (pure integrator states), so they need no tangent lane.  c is reduced over the group with shuffles.
//   decoupled (src/decoupled_lat_long.jl:244-272): continuous linearisation of VehicleModel{LateralTrackingBicycleModel} by AD,
//             then the exact ZOH/FOH of that linear system (scaling-and-squaring Taylor series of e^{Ah} and its first two
//             integrals), one node per thread.
//   stable_limits + actuator bounds for node t+1 (src/coupled_lat_long.jl:356-367, src/vehicle_dynamics.jl:227-263).
#include "pgn_internal.h"

namespace pgn {

struct LinArgs {
    int B, N, T, Ns, nsub;
    VehParams P; CtrlParams C; double un0, un1;
    RecLayout R;
    const double *qs, *us, *ps, *dt;
    double* rec;
    const uint8_t* hold;      // simulate loops with deferred solves: the record of a held vehicle belongs to a QP that is still being solved
};

// envelope + limits of node t+1 and the per-vehicle global part of the record
__device__ __forceinline__ void write_envelope(const LinArgs& a, int kind, const double* qs_n, const double* us_n, const double* ps_n, double* piece) {
    const VehParams& P = a.P;
    const double Uxt = kind == PGN_COUPLED ? qs_n[1] : ps_n[0];
    const double Fx = us_n[1];
    double Fxf, Fxr;
    if (Fx > 0) { Fxf = Fx * P.fwd_frac; Fxr = Fx * P.rwd_frac; } else { Fxf = Fx * P.fwb_frac; Fxr = Fx * P.rwb_frac; }
    double dmin, dmax, H[8], G[4];
    stable_limits(P, Uxt, Fxf, Fxr, dmin, dmax, H, G);
#pragma unroll
    for (int k = 0; k < 8; k++) piece[a.R.oH + k] = H[k];
#pragma unroll
    for (int k = 0; k < 4; k++) piece[a.R.oG + k] = G[k];
    const double dn = kind == PGN_COUPLED ? a.un0 : 1.0;
    piece[a.R.odmin] = fmax(dmin, -P.delta_max) / dn;
    piece[a.R.odmax] = fmin(dmax, P.delta_max) / dn;
    piece[a.R.ofxmax] = fmin(P.Px_max / Uxt, P.Fx_max) / a.un1;
}

// ---- coupled: 4 lanes per (vehicle, interval) ------------------------------------------------------------------------------
// PGN_LIN_MINCTAS (build-time experiment): resident CTAs per SM the register allocation is capped for (2: 244 registers, no spills;
// 3: 168 registers, 380 B of spill stores; 4: 128 registers, 724 B)
#ifndef PGN_LIN_MINCTAS
#define PGN_LIN_MINCTAS 2
#endif
__global__ void __launch_bounds__(128, PGN_LIN_MINCTAS) k_linearize_coupled(const LinArgs a) {
    typedef Dual<2> D;
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int g = gid & 3;                  // tangent group of this lane
    const int node = gid >> 2;
    const int total = a.B * a.T;
    const int nd = node < total ? node : total - 1;
    const int v = nd / a.T, t = nd - v * a.T;
    const bool active = node < total && !(a.hold && a.hold[v]);
    const bool ramp = t >= a.Ns;
    const double* q = a.qs + ((size_t)v * a.N + t) * 6;
    const double* u0p = a.us + ((size_t)v * a.N + t) * 2;
    const double* p0p = a.ps + ((size_t)v * a.N + t) * 4;
    const double* ufp = ramp ? u0p + 2 : u0p;
    const double* pfp = ramp ? p0p + 4 : p0p;
    const double dt = a.dt[(size_t)v * a.T + t];
    // tangent columns of this lane: index into (Ux,Uy,r,dpsi | d0,Fx0 | df,Fxf) = state comps 1..4, u0 comps, uf comps
    const int c0 = 2 * g, c1 = 2 * g + 1;
    D x[6], u0[4], uf[4];
#pragma unroll
    for (int i = 0; i < 6; i++) {
        x[i].v = q[i];
        x[i].d[0] = (i >= 1 && i <= 4 && (i - 1) == c0) ? 1.0 : 0.0;
        x[i].d[1] = (i >= 1 && i <= 4 && (i - 1) == c1) ? 1.0 : 0.0;
    }
#pragma unroll
    for (int k = 0; k < 2; k++) {
        u0[k].v = u0p[k]; uf[k].v = ufp[k];
        u0[k].d[0] = (4 + k == c0) ? 1.0 : 0.0; u0[k].d[1] = (4 + k == c1) ? 1.0 : 0.0;
        if (ramp) { uf[k].d[0] = (6 + k == c0) ? 1.0 : 0.0; uf[k].d[1] = (6 + k == c1) ? 1.0 : 0.0; }
        else      { uf[k].d[0] = u0[k].d[0]; uf[k].d[1] = u0[k].d[1]; }      // StepControl: one control value for the interval
        u0[2 + k] = D(p0p[k]); uf[2 + k] = D(pfp[k]);
    }
    flow_rk4<MODEL_TRACKING, 6, D>(a.P, x, dt, u0, uf, a.nsub);

    double* piece = a.rec + (size_t)v * a.R.rec_len + a.R.piece(t);
    // partial sums of c_i = x+_i - sum_j A_ij x_j - sum_k B0_ik u0_k - sum_k Bf_ik uf_k over this lane's columns
    double part[6];
#pragma unroll
    for (int i = 0; i < 6; i++) part[i] = 0.0;
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const int col = 2 * g + k;
        if (col < 4) {                    // A column (state component col+1)
#pragma unroll
            for (int i = 0; i < 6; i++) { const double d = x[i].d[k]; if (active) piece[a.R.oA + i * 6 + col + 1] = d; part[i] += d * q[col + 1]; }
        } else if (col < 6) {             // B0 column, stored pre-multiplied by u_normalization (coupled_lat_long.jl:338)
            const int c = col - 4;
            const double unc = c == 0 ? a.un0 : a.un1;
#pragma unroll
            for (int i = 0; i < 6; i++) { const double d = x[i].d[k]; if (active) piece[a.R.oB0 + i * 2 + c] = d * unc; part[i] += d * u0p[c]; }
        } else {                          // Bf column (zero for a StepControl interval)
            const int c = col - 6;
            const double unc = c == 0 ? a.un0 : a.un1;
#pragma unroll
            for (int i = 0; i < 6; i++) { const double d = ramp ? x[i].d[k] : 0.0; if (active) piece[a.R.oBf + i * 2 + c] = d * unc; part[i] += d * ufp[c]; }
        }
    }
#pragma unroll
    for (int i = 0; i < 6; i++) {
        part[i] += __shfl_xor_sync(0xffffffffu, part[i], 1);
        part[i] += __shfl_xor_sync(0xffffffffu, part[i], 2);
    }
    if (active && g == 0) {
        // unit columns of the pure integrator states ds (col 0) and e (col 5)
#pragma unroll
        for (int i = 0; i < 6; i++) { piece[a.R.oA + i * 6 + 0] = (i == 0) ? 1.0 : 0.0; piece[a.R.oA + i * 6 + 5] = (i == 5) ? 1.0 : 0.0; }
#pragma unroll
        for (int i = 0; i < 6; i++) {
            double c = x[i].v - part[i];
            if (i == 0) c -= q[0];
            if (i == 5) c -= q[5];
            piece[a.R.oc + i] = c;
        }
    }
    if (active && g == 1) write_envelope(a, PGN_COUPLED, q + 6, u0p + 2, p0p + 4, piece);
    if (active && g == 2 && t == 0) {
        double* glob = a.rec + (size_t)v * a.R.rec_len;
#pragma unroll
        for (int i = 0; i < 6; i++) glob[a.R.o_qcurr + i] = q[i];
        glob[a.R.o_ucurr + 0] = u0p[0] / a.un0;
        glob[a.R.o_ucurr + 1] = u0p[1] / a.un1;
    }
    if (active && g == 3) a.rec[(size_t)v * a.R.rec_len + a.R.o_dt + t] = dt;
}

// ---- decoupled: one (vehicle, interval) per thread -----------------------------------------------------------------------
__device__ __forceinline__ void mat4_mul(const double* A, const double* Bm, double* C) {
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            double s = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) s += A[i * 4 + k] * Bm[k * 4 + j];
            C[i * 4 + j] = s;
        }
}

__global__ void __launch_bounds__(128) k_linearize_decoupled(const LinArgs a) {
    const int node = blockIdx.x * blockDim.x + threadIdx.x;
    if (node >= a.B * a.T) return;
    const int v = node / a.T, t = node - v * a.T;
    if (a.hold && a.hold[v]) return;
    const bool ramp = t >= a.Ns;
    const double* q = a.qs + ((size_t)v * a.N + t) * 4;
    const double* u0p = a.us + ((size_t)v * a.N + t) * 2;
    const double* p0p = a.ps + ((size_t)v * a.N + t) * 4;
    const double dt = a.dt[(size_t)v * a.T + t];
    // continuous linearisation: A = df/dx (4x4), Bc = df/d(delta, Fx, Ux, kappa) (4x4), f
    double A[16], Bc[16], f[4];
    {
        typedef Dual<4> D;
        D x[4], out[4];
#pragma unroll
        for (int i = 0; i < 4; i++) { x[i] = D(q[i]); x[i].d[i] = 1.0; }
        vehicle_model<MODEL_LATERAL, D>(a.P, x, D(u0p[0]), D(u0p[1]), D(p0p[0]), D(p0p[1]), out);
#pragma unroll
        for (int i = 0; i < 4; i++) { f[i] = out[i].v;
#pragma unroll
            for (int j = 0; j < 4; j++) A[i * 4 + j] = out[i].d[j]; }
        D uu[4];
#pragma unroll
        for (int k = 0; k < 4; k++) { uu[k] = D(k < 2 ? u0p[k] : p0p[k - 2]); uu[k].d[k] = 1.0; }
#pragma unroll
        for (int i = 0; i < 4; i++) x[i] = D(q[i]);
        vehicle_model<MODEL_LATERAL, D>(a.P, x, uu[0], uu[1], uu[2], uu[3], out);
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) Bc[i * 4 + j] = out[i].d[j];
    }
    // Phi = e^{A dt}, G1 = int_0^dt e^{As} ds, G2 = int_0^dt e^{A(dt-s)} (s/dt) ds  by Taylor series on h = dt / 2^sq and doubling
    double nrm = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) { double s = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) s += fabs(A[i * 4 + j]); nrm = fmax(nrm, s); }
    int sq = 0;
    { double x = nrm * dt; while (x > 0.5 && sq < 40) { x *= 0.5; sq++; } }
    double h = ldexp(dt, -sq);
    double Ah[16], Phi[16], G1[16], G2[16], Tk[16], Tn[16];
#pragma unroll
    for (int i = 0; i < 16; i++) { Ah[i] = A[i] * h; Tk[i] = (i % 5 == 0) ? 1.0 : 0.0; Phi[i] = Tk[i]; G1[i] = Tk[i]; G2[i] = 0.5 * Tk[i]; }
    for (int k = 1; k <= 16; k++) {
        mat4_mul(Tk, Ah, Tn);
        const double ik = 1.0 / k, c1 = 1.0 / (k + 1), c2 = 1.0 / ((k + 1.0) * (k + 2.0));
#pragma unroll
        for (int i = 0; i < 16; i++) { Tk[i] = Tn[i] * ik; Phi[i] += Tk[i]; G1[i] += Tk[i] * c1; G2[i] += Tk[i] * c2; }
    }
#pragma unroll
    for (int i = 0; i < 16; i++) { G1[i] *= h; G2[i] *= h * h; }     // G2 unnormalised: int_0^h e^{A(h-s)} s ds
    for (int s = 0; s < sq; s++) {
        double M1[16], M2[16];
        mat4_mul(Phi, G2, M1);
#pragma unroll
        for (int i = 0; i < 16; i++) G2[i] = M1[i] + G2[i] + h * G1[i];
        mat4_mul(Phi, G1, M1);
#pragma unroll
        for (int i = 0; i < 16; i++) G1[i] += M1[i];
        mat4_mul(Phi, Phi, M2);
#pragma unroll
        for (int i = 0; i < 16; i++) Phi[i] = M2[i];
        h *= 2;
    }
    const double idt = 1.0 / dt;
    // g0 = f - A x (= B u0 + c0), dB = Bc (u_f - u_0) over (delta, Fx, Ux, kappa)
    double g0[4], dB[4], du[4];
#pragma unroll
    for (int k = 0; k < 4; k++) du[k] = ramp ? ((k < 2 ? u0p[2 + k] : p0p[4 + k - 2]) - (k < 2 ? u0p[k] : p0p[k - 2])) : 0.0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        double s = f[i], d = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) { s -= A[i * 4 + j] * q[j]; d += Bc[i * 4 + j] * du[j]; }
        g0[i] = s; dB[i] = d;
    }
    double* piece = a.rec + (size_t)v * a.R.rec_len + a.R.piece(t);
    const double d0 = u0p[0], df = ramp ? u0p[2] : u0p[0];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        double xp = 0, g1b = 0, g2b = 0, ax = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            xp += Phi[i * 4 + j] * q[j] + G1[i * 4 + j] * g0[j] + G2[i * 4 + j] * idt * dB[j];
            g1b += G1[i * 4 + j] * Bc[j * 4 + 0];
            g2b += G2[i * 4 + j] * idt * Bc[j * 4 + 0];
            ax += Phi[i * 4 + j] * q[j];
            piece[a.R.oA + i * 4 + j] = Phi[i * 4 + j];
        }
        const double B0 = ramp ? g1b - g2b : g1b, Bf = ramp ? g2b : 0.0;
        piece[a.R.oB0 + i] = B0;
        piece[a.R.oBf + i] = Bf;
        piece[a.R.oc + i] = xp - ax - B0 * d0 - (ramp ? Bf * df : 0.0);
    }
    write_envelope(a, PGN_DECOUPLED, q + 4, u0p + 2, p0p + 4, piece);
    double* glob = a.rec + (size_t)v * a.R.rec_len;
    glob[a.R.o_dt + t] = dt;
    if (t == 0) {
#pragma unroll
        for (int i = 0; i < 4; i++) glob[a.R.o_qcurr + i] = q[i];
        glob[a.R.o_ucurr] = u0p[0];
        glob[a.R.o_hji + 0] = 0; glob[a.R.o_hji + 1] = 0; glob[a.R.o_hji + 2] = 1.0;
    }
}

void launch_linearize(pgn_handle* h) {
    LinArgs a;
    a.B = h->nv; a.N = h->N; a.T = h->T; a.Ns = h->cfg.N_short; a.nsub = h->cfg.rk4_substeps;     // per-vehicle rows only: a pipeline part is an offset
    a.P = h->veh; a.C = h->ctl; a.un0 = h->un[0]; a.un1 = h->un[1];
    a.R = h->tab.rec;
    const size_t o = (size_t)h->v0;
    a.qs = h->d_qs + o * h->N * h->nx; a.us = h->d_us + o * h->N * 2; a.ps = h->d_ps + o * h->N * 4; a.dt = h->d_dt + o * h->T; a.rec = h->d_rec + o * h->tab.rec.rec_len;
    a.hold = h->hold_on ? h->d_hold + o : nullptr;
    const long long nodes = (long long)h->nv * h->T;
    if (h->cfg.kind == PGN_COUPLED) {
        const long long threads = nodes * 4;
        k_linearize_coupled<<<(unsigned)((threads + 127) / 128), 128, 0, h->stream>>>(a);
    } else {
        k_linearize_decoupled<<<(unsigned)((nodes + 127) / 128), 128, 0, h->stream>>>(a);
    }
    h->launches++;
}

}  // namespace pgn
