// oracle/osqp_port.hpp — TEST INFRASTRUCTURE ONLY (CPU oracle). PARITY UNPINNED (reference has no golden vectors).
//
// The QP solve of the reference happens in a third-party dependency that is NOT vendored under /root/reference:
// OSQP.jl 0.4.0 -> libosqp 0.4.x (C) + QDLDL (pins: /root/reference/env/Manifest.toml `[[OSQP]] version = "0.4.0"`).
// Call sites: src/coupled_lat_long.jl:201-204, src/decoupled_lat_long.jl:137-140 (defaults except verbose/warm_start),
// src/model_predictive_control.jl:76 (solve!).  This file restates OSQP's published algorithm
// (Stellato et al., "OSQP: an operator splitting solver for quadratic programs", 2020, Alg. 1 + Sec. 5):
// modified Ruiz equilibration with cost scaling, per-constraint rho (equalities 1e3*rho), quasi-definite KKT LDL'
// (up-looking, elimination-tree based, as in QDLDL), alpha-relaxed ADMM, residual checks every `check_termination`
// iterations in unscaled norms, primal/dual infeasibility certificates, adaptive rho with refactorisation, warm start
// carrying the internal (scaled) iterates and rho from solve to solve.
// Pinned choice: adaptive_rho_interval is wall-clock dependent in libosqp ("automatic"); fixed here (default 25).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

namespace orc {

struct Csc {  // compressed sparse column
    int nrow = 0, ncol = 0;
    std::vector<int> p, i;
    std::vector<double> x;
};

struct OsqpSettings {
    double rho = 0.1, sigma = 1e-6, alpha = 1.6;
    double eps_abs = 1e-3, eps_rel = 1e-3, eps_prim_inf = 1e-4, eps_dual_inf = 1e-4;
    int max_iter = 4000, scaling = 10, check_termination = 25;
    int adaptive_rho = 1, adaptive_rho_interval = 25;
    double adaptive_rho_tolerance = 5.0;
    int warm_start = 1;
};

enum OsqpStatus {
    OSQP_SOLVED = 1, OSQP_SOLVED_INACCURATE = 2, OSQP_PRIMAL_INFEASIBLE_INACCURATE = 3, OSQP_DUAL_INFEASIBLE_INACCURATE = 4,
    OSQP_MAX_ITER_REACHED = -2, OSQP_PRIMAL_INFEASIBLE = -3, OSQP_DUAL_INFEASIBLE = -4, OSQP_UNSOLVED = -10
};

static const double OSQP_INFTY_ = 1e20;
static const double MIN_SCALING_ = 1e-4, MAX_SCALING_ = 1e4;
static const double RHO_MIN_ = 1e-6, RHO_MAX_ = 1e6, RHO_TOL_ = 1e-4, RHO_EQ_OVER_RHO_INEQ_ = 1e3;

// exact (slow) minimum-degree ordering on the pattern of a symmetric matrix given by its upper triangle
inline std::vector<int> min_degree_order(int n, const std::vector<int>& Kp, const std::vector<int>& Ki) {
    std::vector<std::vector<int>> adj(n);
    for (int j = 0; j < n; j++) for (int p = Kp[j]; p < Kp[j + 1]; p++) { int i = Ki[p]; if (i != j) { adj[i].push_back(j); adj[j].push_back(i); } }
    for (auto& a : adj) { std::sort(a.begin(), a.end()); a.erase(std::unique(a.begin(), a.end()), a.end()); }
    std::vector<char> alive(n, 1);
    std::vector<int> perm;
    perm.reserve(n);
    for (int it = 0; it < n; it++) {
        int best = -1; size_t bd = (size_t)-1;
        for (int v = 0; v < n; v++) if (alive[v] && adj[v].size() < bd) { bd = adj[v].size(); best = v; }
        int v = best;
        alive[v] = 0; perm.push_back(v);
        std::vector<int> nb = adj[v];
        for (int a : nb) { auto& A = adj[a]; A.erase(std::lower_bound(A.begin(), A.end(), v)); }
        for (size_t x = 0; x < nb.size(); x++) for (size_t y = x + 1; y < nb.size(); y++) {
            int a = nb[x], b = nb[y];
            auto& A = adj[a];
            auto itb = std::lower_bound(A.begin(), A.end(), b);
            if (itb == A.end() || *itb != b) { A.insert(itb, b); auto& Bv = adj[b]; Bv.insert(std::lower_bound(Bv.begin(), Bv.end(), a), a); }
        }
        adj[v].clear();
    }
    return perm;
}

// Up-looking sparse LDL' (QDLDL's algorithm) of a symmetric quasi-definite matrix given by its upper triangle (CSC).
struct Ldl {
    int n = 0;
    std::vector<int> etree, Lnz, Lp, Li;
    std::vector<double> Lx, D, Dinv;
    std::vector<int> iwork; std::vector<char> bwork; std::vector<double> fwork;
    bool symbolic(int n_, const std::vector<int>& Ap, const std::vector<int>& Ai) {
        n = n_;
        etree.assign(n, -1); Lnz.assign(n, 0);
        std::vector<int> work(n, 0);
        for (int j = 0; j < n; j++) {
            work[j] = j;
            for (int p = Ap[j]; p < Ap[j + 1]; p++) {
                int i = Ai[p];
                if (i > j) return false;
                while (work[i] != j) {
                    if (etree[i] == -1) etree[i] = j;
                    Lnz[i]++;
                    work[i] = j;
                    i = etree[i];
                }
            }
        }
        Lp.assign(n + 1, 0);
        for (int i = 0; i < n; i++) Lp[i + 1] = Lp[i] + Lnz[i];
        Li.assign(Lp[n], 0); Lx.assign(Lp[n], 0.0); D.assign(n, 0.0); Dinv.assign(n, 0.0);
        iwork.assign(3 * n, 0); bwork.assign(n, 0); fwork.assign(n, 0.0);
        return true;
    }
    int factor(const std::vector<int>& Ap, const std::vector<int>& Ai, const std::vector<double>& Ax) {
        char* yMarkers = bwork.data();
        int* yIdx = iwork.data(); int* elimBuffer = yIdx + n; int* LNext = elimBuffer + n;
        double* yVals = fwork.data();
        for (int i = 0; i < n; i++) { yMarkers[i] = 0; yVals[i] = 0; D[i] = 0; LNext[i] = Lp[i]; }
        int positive = 0;
        for (int k = 0; k < n; k++) {
            int nnzY = 0;
            for (int p = Ap[k]; p < Ap[k + 1]; p++) {
                int bidx = Ai[p];
                if (bidx == k) { D[k] = Ax[p]; continue; }
                yVals[bidx] = Ax[p];
                int next = bidx;
                if (!yMarkers[next]) {
                    yMarkers[next] = 1; elimBuffer[0] = next; int nnzE = 1;
                    next = etree[bidx];
                    while (next != -1 && next < k) {
                        if (yMarkers[next]) break;
                        yMarkers[next] = 1; elimBuffer[nnzE++] = next; next = etree[next];
                    }
                    while (nnzE) yIdx[nnzY++] = elimBuffer[--nnzE];
                }
            }
            for (int i = nnzY - 1; i >= 0; i--) {
                int cidx = yIdx[i];
                int tmp = LNext[cidx];
                double yv = yVals[cidx];
                for (int j = Lp[cidx]; j < tmp; j++) yVals[Li[j]] -= Lx[j] * yv;
                Li[tmp] = k; Lx[tmp] = yv * Dinv[cidx];
                D[k] -= yv * Lx[tmp];
                LNext[cidx]++;
                yVals[cidx] = 0; yMarkers[cidx] = 0;
            }
            if (D[k] == 0) return -1;
            if (D[k] > 0) positive++;
            Dinv[k] = 1.0 / D[k];
        }
        return positive;
    }
    void solve(double* x) const {
        for (int i = 0; i < n; i++) { double xi = x[i]; for (int j = Lp[i]; j < Lp[i + 1]; j++) x[Li[j]] -= Lx[j] * xi; }
        for (int i = 0; i < n; i++) x[i] *= Dinv[i];
        for (int i = n - 1; i >= 0; i--) { double xi = x[i]; for (int j = Lp[i]; j < Lp[i + 1]; j++) xi -= Lx[j] * x[Li[j]]; x[i] = xi; }
    }
};

struct OsqpInfo { int iter = 0; int status = OSQP_UNSOLVED; double pri_res = 0, dua_res = 0, obj_val = 0; int rho_updates = 0; double rho_estimate = 0; };

class OsqpSolver {
public:
    OsqpSettings st;
    OsqpInfo info;
    int n = 0, m = 0;
    // original (unscaled) data
    Csc P0, A0; std::vector<double> q0, l0, u0;
    // scaled data
    Csc P, A; std::vector<double> q, l, u;
    std::vector<double> D, E, Dinv, Einv; double c = 1, cinv = 1;
    // iterates (scaled space)
    std::vector<double> x, z, y, x_prev, z_prev, xz_tilde, delta_x, delta_y, Ax, Px, Aty, Atdy, Pdx, Adx;
    std::vector<double> rho_vec, rho_inv_vec; std::vector<int> constr_type;
    // solution (unscaled)
    std::vector<double> sol_x, sol_y;
    // KKT
    std::vector<int> perm, iperm, Kp, Ki; std::vector<double> Kx;
    std::vector<int> PtoK, AtoK, rhoToK, sigToK;
    Ldl ldl; std::vector<double> rhs_perm;
    long n_factor = 0;

    void setup(const Csc& P_, const std::vector<double>& q_, const Csc& A_, const std::vector<double>& l_,
               const std::vector<double>& u_, const OsqpSettings& s) {
        st = s; n = P_.ncol; m = A_.nrow;
        P0 = P_; A0 = A_; q0 = q_; l0 = l_; u0 = u_;
        for (auto& v : l0) v = std::max(v, -OSQP_INFTY_);
        for (auto& v : u0) v = std::min(v, OSQP_INFTY_);
        x.assign(n, 0); z.assign(m, 0); y.assign(m, 0); x_prev.assign(n, 0); z_prev.assign(m, 0); xz_tilde.assign(n + m, 0);
        delta_x.assign(n, 0); delta_y.assign(m, 0); Ax.assign(m, 0); Px.assign(n, 0); Aty.assign(n, 0); Atdy.assign(n, 0);
        Pdx.assign(n, 0); Adx.assign(m, 0); sol_x.assign(n, 0); sol_y.assign(m, 0);
        D.assign(n, 1); E.assign(m, 1); Dinv.assign(n, 1); Einv.assign(m, 1);
        rho_vec.assign(m, 0); rho_inv_vec.assign(m, 0); constr_type.assign(m, -2);
        build_kkt_pattern();
        rescale_and_refactor();
    }
    // value-only update of P, A, q, l, u (same pattern): osqp_update_{lin_cost,bounds,P_A} in one go
    void update(const double* Px_, const double* Ax_, const double* q_, const double* l_, const double* u_) {
        if (Px_) std::copy(Px_, Px_ + P0.x.size(), P0.x.begin());
        if (Ax_) std::copy(Ax_, Ax_ + A0.x.size(), A0.x.begin());
        if (q_) std::copy(q_, q_ + n, q0.begin());
        if (l_) for (int i = 0; i < m; i++) l0[i] = std::max(l_[i], -OSQP_INFTY_);
        if (u_) for (int i = 0; i < m; i++) u0[i] = std::min(u_[i], OSQP_INFTY_);
        rescale_and_refactor();
    }
    void cold_start() { std::fill(x.begin(), x.end(), 0); std::fill(z.begin(), z.end(), 0); std::fill(y.begin(), y.end(), 0); }
    void reset_rho(double rho) { st.rho = rho; constr_type.assign(m, -2); set_rho_vec(); refactor(); }
    // osqp_warm_start: x, y given unscaled
    void warm_start(const double* xw, const double* yw) {
        for (int i = 0; i < n; i++) x[i] = Dinv[i] * xw[i];
        for (int i = 0; i < m; i++) y[i] = Einv[i] * yw[i] * c;
        mat_vec(A, x.data(), z.data());
    }

    int solve() {
        info.status = OSQP_UNSOLVED; info.rho_updates = 0;
        if (!st.warm_start) cold_start();
        int iter; bool can_check = false;
        for (iter = 1; iter <= st.max_iter; iter++) {
            x.swap(x_prev); z.swap(z_prev);
            // update_xz_tilde
            for (int i = 0; i < n; i++) xz_tilde[i] = st.sigma * x_prev[i] - q[i];
            for (int i = 0; i < m; i++) xz_tilde[n + i] = z_prev[i] - rho_inv_vec[i] * y[i];
            kkt_solve(xz_tilde.data());
            for (int i = 0; i < m; i++) xz_tilde[n + i] = z_prev[i] + rho_inv_vec[i] * (xz_tilde[n + i] - y[i]);
            // update_x
            for (int i = 0; i < n; i++) { x[i] = st.alpha * xz_tilde[i] + (1.0 - st.alpha) * x_prev[i]; delta_x[i] = x[i] - x_prev[i]; }
            // update_z
            for (int i = 0; i < m; i++) {
                double v = st.alpha * xz_tilde[n + i] + (1.0 - st.alpha) * z_prev[i] + rho_inv_vec[i] * y[i];
                z[i] = std::min(std::max(v, l[i]), u[i]);
            }
            // update_y
            for (int i = 0; i < m; i++) {
                delta_y[i] = rho_vec[i] * (st.alpha * xz_tilde[n + i] + (1.0 - st.alpha) * z_prev[i] - z[i]);
                y[i] += delta_y[i];
            }
            can_check = st.check_termination && (iter % st.check_termination == 0);
            if (can_check) {
                update_info(iter);
                if (check_termination(false)) break;
            }
            if (st.adaptive_rho && st.adaptive_rho_interval && (iter % st.adaptive_rho_interval == 0)) {
                if (!can_check) update_info(iter);
                adapt_rho();
            }
        }
        if (iter > st.max_iter) iter = st.max_iter;
        if (!can_check) { update_info(iter); check_termination(false); }
        if (info.status == OSQP_UNSOLVED) { if (!check_termination(true)) info.status = OSQP_MAX_ITER_REACHED; }
        info.iter = iter;
        info.rho_estimate = compute_rho_estimate();
        store_solution();
        return info.status;
    }

    // ---- helpers exposed for tests ----
    static void mat_vec(const Csc& M, const double* v, double* out) {
        std::fill(out, out + M.nrow, 0.0);
        for (int j = 0; j < M.ncol; j++) for (int p = M.p[j]; p < M.p[j + 1]; p++) out[M.i[p]] += M.x[p] * v[j];
    }
    static void mat_tvec(const Csc& M, const double* v, double* out) {
        for (int j = 0; j < M.ncol; j++) { double s = 0; for (int p = M.p[j]; p < M.p[j + 1]; p++) s += M.x[p] * v[M.i[p]]; out[j] = s; }
    }
    static void sym_mat_vec(const Csc& Pu, const double* v, double* out) {  // P stored as upper triangle
        std::fill(out, out + Pu.ncol, 0.0);
        for (int j = 0; j < Pu.ncol; j++) for (int p = Pu.p[j]; p < Pu.p[j + 1]; p++) {
            int i = Pu.i[p];
            out[i] += Pu.x[p] * v[j];
            if (i != j) out[j] += Pu.x[p] * v[i];
        }
    }

private:
    static double norm_inf(const std::vector<double>& v) { double r = 0; for (double a : v) r = std::max(r, std::fabs(a)); return r; }
    static double scaled_norm_inf(const std::vector<double>& S, const std::vector<double>& v) { double r = 0; for (size_t i = 0; i < v.size(); i++) r = std::max(r, std::fabs(S[i] * v[i])); return r; }
    static void limit_scaling(std::vector<double>& v) { for (auto& a : v) { a = a < MIN_SCALING_ ? 1.0 : a; a = a > MAX_SCALING_ ? MAX_SCALING_ : a; } }
    static double limit_scaling1(double a) { a = a < MIN_SCALING_ ? 1.0 : a; return a > MAX_SCALING_ ? MAX_SCALING_ : a; }

    void scale_data() {
        P = P0; A = A0; q = q0; l = l0; u = u0;
        c = 1.0;
        std::fill(D.begin(), D.end(), 1.0); std::fill(E.begin(), E.end(), 1.0);
        std::vector<double> Dt(n), Et(m);
        for (int it = 0; it < st.scaling; it++) {
            // inf-norm of the columns of [P A'; A 0]
            std::fill(Dt.begin(), Dt.end(), 0.0); std::fill(Et.begin(), Et.end(), 0.0);
            for (int j = 0; j < n; j++) for (int p = P.p[j]; p < P.p[j + 1]; p++) {
                double a = std::fabs(P.x[p]); int i = P.i[p];
                Dt[j] = std::max(Dt[j], a); if (i != j) Dt[i] = std::max(Dt[i], a);
            }
            for (int j = 0; j < n; j++) for (int p = A.p[j]; p < A.p[j + 1]; p++) {
                double a = std::fabs(A.x[p]);
                Dt[j] = std::max(Dt[j], a); Et[A.i[p]] = std::max(Et[A.i[p]], a);
            }
            limit_scaling(Dt); limit_scaling(Et);
            for (auto& a : Dt) a = 1.0 / std::sqrt(a);
            for (auto& a : Et) a = 1.0 / std::sqrt(a);
            for (int j = 0; j < n; j++) for (int p = P.p[j]; p < P.p[j + 1]; p++) P.x[p] *= Dt[P.i[p]] * Dt[j];
            for (int j = 0; j < n; j++) for (int p = A.p[j]; p < A.p[j + 1]; p++) A.x[p] *= Et[A.i[p]] * Dt[j];
            for (int i = 0; i < n; i++) { q[i] *= Dt[i]; D[i] *= Dt[i]; }
            for (int i = 0; i < m; i++) E[i] *= Et[i];
            // cost normalisation
            std::fill(Dt.begin(), Dt.end(), 0.0);
            for (int j = 0; j < n; j++) for (int p = P.p[j]; p < P.p[j + 1]; p++) {
                double a = std::fabs(P.x[p]); int i = P.i[p];
                Dt[j] = std::max(Dt[j], a); if (i != j) Dt[i] = std::max(Dt[i], a);
            }
            double c_temp = 0; for (double a : Dt) c_temp += a; c_temp /= n;
            double inf_q = limit_scaling1(norm_inf(q));
            c_temp = std::max(c_temp, inf_q);
            c_temp = limit_scaling1(c_temp);
            c_temp = 1.0 / c_temp;
            for (auto& a : P.x) a *= c_temp;
            for (auto& a : q) a *= c_temp;
            c *= c_temp;
        }
        cinv = 1.0 / c;
        for (int i = 0; i < n; i++) Dinv[i] = 1.0 / D[i];
        for (int i = 0; i < m; i++) { Einv[i] = 1.0 / E[i]; l[i] *= E[i]; u[i] *= E[i]; }
    }
    bool set_rho_vec() {
        bool changed = false;
        for (int i = 0; i < m; i++) {
            int t; double r;
            if (l[i] < -OSQP_INFTY_ * MIN_SCALING_ && u[i] > OSQP_INFTY_ * MIN_SCALING_) { t = -1; r = RHO_MIN_; }
            else if (u[i] - l[i] < RHO_TOL_) { t = 1; r = RHO_EQ_OVER_RHO_INEQ_ * st.rho; }
            else { t = 0; r = st.rho; }
            if (t != constr_type[i] || r != rho_vec[i]) changed = true;
            constr_type[i] = t; rho_vec[i] = r; rho_inv_vec[i] = 1.0 / r;
        }
        return changed;
    }
    void build_kkt_pattern() {
        // upper triangle of K = [P + sigma I, A'; A, -diag(1/rho)] in natural order as triplets, then permute
        const int N = n + m;
        struct Trip { int r, c, src, kind; };  // kind 0: P entry, 1: A entry, 2: sigma diag (no P entry), 3: rho diag
        std::vector<Trip> T;
        std::vector<char> hasdiag(n, 0);
        for (int j = 0; j < n; j++) for (int p = P0.p[j]; p < P0.p[j + 1]; p++) { T.push_back({P0.i[p], j, p, 0}); if (P0.i[p] == j) hasdiag[j] = 1; }
        for (int j = 0; j < n; j++) if (!hasdiag[j]) T.push_back({j, j, j, 2});
        for (int j = 0; j < n; j++) for (int p = A0.p[j]; p < A0.p[j + 1]; p++) T.push_back({j, n + A0.i[p], p, 1});
        for (int i = 0; i < m; i++) T.push_back({n + i, n + i, i, 3});
        // ordering from the natural pattern
        std::vector<int> cnt(N + 1, 0);
        for (auto& t : T) cnt[std::max(t.r, t.c) + 1]++;
        std::vector<int> Np(N + 1, 0);
        for (int j = 0; j < N; j++) Np[j + 1] = Np[j] + cnt[j + 1];
        std::vector<int> Ni(T.size()), fill = Np;
        for (auto& t : T) { int cc = std::max(t.r, t.c), rr = std::min(t.r, t.c); Ni[fill[cc]++] = rr; }
        perm = min_degree_order(N, Np, Ni);
        iperm.assign(N, 0);
        for (int k = 0; k < N; k++) iperm[perm[k]] = k;
        // permuted upper triangle
        std::vector<int> pc(T.size()), pr(T.size());
        std::fill(cnt.begin(), cnt.end(), 0);
        for (size_t e = 0; e < T.size(); e++) {
            int a = iperm[T[e].r], b = iperm[T[e].c];
            pr[e] = std::min(a, b); pc[e] = std::max(a, b);
            cnt[pc[e] + 1]++;
        }
        Kp.assign(N + 1, 0);
        for (int j = 0; j < N; j++) Kp[j + 1] = Kp[j] + cnt[j + 1];
        Ki.assign(T.size(), 0); Kx.assign(T.size(), 0.0);
        // sort entries within columns by row
        std::vector<size_t> order(T.size());
        for (size_t e = 0; e < T.size(); e++) order[e] = e;
        std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return pc[a] != pc[b] ? pc[a] < pc[b] : pr[a] < pr[b]; });
        PtoK.assign(P0.x.size(), -1); AtoK.assign(A0.x.size(), -1); rhoToK.assign(m, -1); sigToK.assign(n, -1);
        for (size_t pos = 0; pos < order.size(); pos++) {
            size_t e = order[pos];
            Ki[pos] = pr[e];
            switch (T[e].kind) {
                case 0: PtoK[T[e].src] = (int)pos; if (T[e].r == T[e].c) sigToK[T[e].c] = (int)pos; break;
                case 1: AtoK[T[e].src] = (int)pos; break;
                case 2: sigToK[T[e].src] = (int)pos; break;
                case 3: rhoToK[T[e].src] = (int)pos; break;
            }
        }
        ldl.symbolic(N, Kp, Ki);
        rhs_perm.assign(N, 0.0);
    }
    void refactor() {
        std::fill(Kx.begin(), Kx.end(), 0.0);
        for (size_t p = 0; p < P.x.size(); p++) Kx[PtoK[p]] += P.x[p];
        for (int j = 0; j < n; j++) Kx[sigToK[j]] += st.sigma;
        for (size_t p = 0; p < A.x.size(); p++) Kx[AtoK[p]] += A.x[p];
        for (int i = 0; i < m; i++) Kx[rhoToK[i]] = -rho_inv_vec[i];
        ldl.factor(Kp, Ki, Kx);
        n_factor++;
    }
    void rescale_and_refactor() {
        scale_data();
        set_rho_vec();
        refactor();
    }
    void kkt_solve(double* b) {
        const int N = n + m;
        for (int k = 0; k < N; k++) rhs_perm[k] = b[perm[k]];
        ldl.solve(rhs_perm.data());
        for (int k = 0; k < N; k++) b[perm[k]] = rhs_perm[k];
    }
    void update_info(int iter) {
        info.iter = iter;
        // primal residual (z_prev <- Ax - z, scaled)
        mat_vec(A, x.data(), Ax.data());
        for (int i = 0; i < m; i++) z_prev[i] = Ax[i] - z[i];
        info.pri_res = st.scaling ? scaled_norm_inf(Einv, z_prev) : norm_inf(z_prev);
        // dual residual (x_prev <- Px + q + A'y, scaled)
        sym_mat_vec(P, x.data(), Px.data());
        mat_tvec(A, y.data(), Aty.data());
        for (int i = 0; i < n; i++) x_prev[i] = q[i] + Px[i] + Aty[i];
        info.dua_res = st.scaling ? cinv * scaled_norm_inf(Dinv, x_prev) : norm_inf(x_prev);
    }
    bool is_primal_infeasible(double eps) {
        for (int i = 0; i < m; i++) {
            if (u[i] > OSQP_INFTY_ * MIN_SCALING_) {
                if (l[i] < -OSQP_INFTY_ * MIN_SCALING_) delta_y[i] = 0.0;
                else delta_y[i] = std::min(delta_y[i], 0.0);
            } else if (l[i] < -OSQP_INFTY_ * MIN_SCALING_) delta_y[i] = std::max(delta_y[i], 0.0);
        }
        double norm_dy = st.scaling ? scaled_norm_inf(E, delta_y) : norm_inf(delta_y);
        if (norm_dy > eps) {
            double lhs = 0;
            for (int i = 0; i < m; i++) lhs += u[i] * std::max(delta_y[i], 0.0) + l[i] * std::min(delta_y[i], 0.0);
            if (lhs < -eps * norm_dy) {
                mat_tvec(A, delta_y.data(), Atdy.data());
                double nrm = st.scaling ? scaled_norm_inf(Dinv, Atdy) : norm_inf(Atdy);
                return nrm < eps * norm_dy;
            }
        }
        return false;
    }
    bool is_dual_infeasible(double eps) {
        double norm_dx = st.scaling ? scaled_norm_inf(D, delta_x) : norm_inf(delta_x);
        double cost_scaling = st.scaling ? c : 1.0;
        if (norm_dx > eps) {
            double qdx = 0; for (int i = 0; i < n; i++) qdx += q[i] * delta_x[i];
            if (qdx < -cost_scaling * eps * norm_dx) {
                sym_mat_vec(P, delta_x.data(), Pdx.data());
                double nrm = st.scaling ? scaled_norm_inf(Dinv, Pdx) : norm_inf(Pdx);
                if (nrm < cost_scaling * eps * norm_dx) {
                    mat_vec(A, delta_x.data(), Adx.data());
                    for (int i = 0; i < m; i++) {
                        double a = st.scaling ? Einv[i] * Adx[i] : Adx[i];
                        if ((u[i] < OSQP_INFTY_ * MIN_SCALING_ && a > eps * norm_dx) || (l[i] > -OSQP_INFTY_ * MIN_SCALING_ && a < -eps * norm_dx)) return false;
                    }
                    return true;
                }
            }
        }
        return false;
    }
    bool check_termination(bool approximate) {
        double eps_abs = st.eps_abs, eps_rel = st.eps_rel, epi = st.eps_prim_inf, edi = st.eps_dual_inf;
        if (approximate) { eps_abs *= 10; eps_rel *= 10; epi *= 10; edi *= 10; }
        bool prim_ok = false, dual_ok = false, prim_inf = false, dual_inf = false;
        if (m == 0) prim_ok = true;
        else {
            double nz = st.scaling ? scaled_norm_inf(Einv, z) : norm_inf(z);
            double nAx = st.scaling ? scaled_norm_inf(Einv, Ax) : norm_inf(Ax);
            double eps_prim = eps_abs + eps_rel * std::max(nz, nAx);
            if (info.pri_res < eps_prim) prim_ok = true;
            else prim_inf = is_primal_infeasible(epi);
        }
        double nq = st.scaling ? scaled_norm_inf(Dinv, q) : norm_inf(q);
        double nAty = st.scaling ? scaled_norm_inf(Dinv, Aty) : norm_inf(Aty);
        double nPx = st.scaling ? scaled_norm_inf(Dinv, Px) : norm_inf(Px);
        double mx = std::max(nq, std::max(nAty, nPx));
        double eps_dual = eps_abs + eps_rel * (st.scaling ? cinv * mx : mx);
        if (info.dua_res < eps_dual) dual_ok = true;
        else dual_inf = is_dual_infeasible(edi);
        if (prim_ok && dual_ok) { info.status = approximate ? OSQP_SOLVED_INACCURATE : OSQP_SOLVED; return true; }
        if (prim_inf) { info.status = approximate ? OSQP_PRIMAL_INFEASIBLE_INACCURATE : OSQP_PRIMAL_INFEASIBLE; return true; }
        if (dual_inf) { info.status = approximate ? OSQP_DUAL_INFEASIBLE_INACCURATE : OSQP_DUAL_INFEASIBLE; return true; }
        return false;
    }
    double compute_rho_estimate() {
        double pri = norm_inf(z_prev), dua = norm_inf(x_prev);
        double pn = std::max(norm_inf(z), norm_inf(Ax));
        pri /= (pn + 1e-10);
        double dn = std::max(norm_inf(q), std::max(norm_inf(Aty), norm_inf(Px)));
        dua /= (dn + 1e-10);
        double r = st.rho * std::sqrt(pri / (dua + 1e-10));
        return std::min(std::max(r, RHO_MIN_), RHO_MAX_);
    }
    void adapt_rho() {
        double rho_new = compute_rho_estimate();
        info.rho_estimate = rho_new;
        if (rho_new > st.rho * st.adaptive_rho_tolerance || rho_new < st.rho / st.adaptive_rho_tolerance) {
            st.rho = std::min(std::max(rho_new, RHO_MIN_), RHO_MAX_);
            for (int i = 0; i < m; i++) {
                if (constr_type[i] == 0) { rho_vec[i] = st.rho; rho_inv_vec[i] = 1.0 / st.rho; }
                else if (constr_type[i] == 1) { rho_vec[i] = RHO_EQ_OVER_RHO_INEQ_ * st.rho; rho_inv_vec[i] = 1.0 / rho_vec[i]; }
            }
            refactor();
            info.rho_updates++;
        }
    }
    void store_solution() {
        bool infeas = (info.status == OSQP_PRIMAL_INFEASIBLE || info.status == OSQP_PRIMAL_INFEASIBLE_INACCURATE ||
                       info.status == OSQP_DUAL_INFEASIBLE || info.status == OSQP_DUAL_INFEASIBLE_INACCURATE);
        if (!infeas) {
            for (int i = 0; i < n; i++) sol_x[i] = D[i] * x[i];
            for (int i = 0; i < m; i++) sol_y[i] = E[i] * y[i] * cinv;
        } else {
            double nan = std::numeric_limits<double>::quiet_NaN();
            std::fill(sol_x.begin(), sol_x.end(), nan); std::fill(sol_y.begin(), sol_y.end(), nan);
            cold_start();
        }
    }
};

}  // namespace orc
