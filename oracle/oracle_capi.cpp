// oracle/oracle_capi.cpp — TEST INFRASTRUCTURE ONLY (CPU oracle). PARITY UNPINNED (reference has no golden vectors).
// extern "C" surface over the CPU restatement (vehicle.hpp, trajectory.hpp, linearize.hpp, osqp_port.hpp, hji.hpp, mpc.hpp)
// so that tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg can drive it through ctypes.
// Nothing under pigeon.jl_b200/ may include, link or load this.
#include <cstring>
#include <vector>
#include <atomic>
#include <thread>
#include "mpc.hpp"

using namespace orc;

static VehicleParams vp_from(const double* v) { VehicleParams P; std::memcpy(&P, v, sizeof(double) * VEHICLE_PARAMS_LEN); return P; }
static ControlParams cp_from(const double* c) {
    ControlParams C;
    C.V_min = c[0]; C.V_max = c[1]; C.k_V = c[2]; C.k_s = c[3]; C.delta_dot_max = c[4]; C.Q_ds = c[5]; C.Q_dpsi = c[6]; C.Q_e = c[7];
    C.W_beta = c[8]; C.W_r = c[9]; C.W_HJI = c[10]; C.N_HJI = (int)c[11]; C.R_delta = c[12]; C.R_ddelta = c[13]; C.R_Fx = c[14]; C.R_dFx = c[15];
    return C;
}
static OsqpSettings st_from(const double* s) {
    OsqpSettings S;
    if (!s) return S;
    S.rho = s[0]; S.sigma = s[1]; S.alpha = s[2]; S.eps_abs = s[3]; S.eps_rel = s[4]; S.eps_prim_inf = s[5]; S.eps_dual_inf = s[6];
    S.max_iter = (int)s[7]; S.scaling = (int)s[8]; S.check_termination = (int)s[9]; S.adaptive_rho = (int)s[10];
    S.adaptive_rho_interval = (int)s[11]; S.adaptive_rho_tolerance = s[12]; S.warm_start = (int)s[13];
    return S;
}

extern "C" {

int orc_sizeof_vehicle_params() { return VEHICLE_PARAMS_LEN; }
void orc_x1(double* out) { VehicleParams P = X1(); std::memcpy(out, &P, sizeof(double) * VEHICLE_PARAMS_LEN); }
void orc_control_params_default(int kind, double* c) {
    ControlParams C = kind == MPC_COUPLED ? ControlParams::coupled_defaults() : ControlParams::decoupled_defaults();
    c[0] = C.V_min; c[1] = C.V_max; c[2] = C.k_V; c[3] = C.k_s; c[4] = C.delta_dot_max; c[5] = C.Q_ds; c[6] = C.Q_dpsi; c[7] = C.Q_e;
    c[8] = C.W_beta; c[9] = C.W_r; c[10] = C.W_HJI; c[11] = C.N_HJI; c[12] = C.R_delta; c[13] = C.R_ddelta; c[14] = C.R_Fx; c[15] = C.R_dFx;
}
void orc_osqp_settings_default(double* s) {
    OsqpSettings S;
    s[0] = S.rho; s[1] = S.sigma; s[2] = S.alpha; s[3] = S.eps_abs; s[4] = S.eps_rel; s[5] = S.eps_prim_inf; s[6] = S.eps_dual_inf;
    s[7] = S.max_iter; s[8] = S.scaling; s[9] = S.check_termination; s[10] = S.adaptive_rho; s[11] = S.adaptive_rho_interval;
    s[12] = S.adaptive_rho_tolerance; s[13] = S.warm_start;
}

// ---- vehicle primitives ----
void orc_vehicle_model(int kind, const double* vp, const double* q, const double* u2, const double* p4, double* out) {
    vehicle_model<double>(kind, vp_from(vp), q, u2, p4, out);
}
void orc_bicycle_model_raw(int kind, const double* vp, const double* q, const double* u3, const double* p4, double* out) {
    VehicleParams P = vp_from(vp);
    if (kind == MODEL_BICYCLE) bicycle_model<double>(P, q, u3, out);
    else if (kind == MODEL_TRACKING) tracking_model<double>(P, q, u3, p4, out);
    else lateral_model<double>(P, q, u3, p4, out);
}
void orc_lateral_tire_forces(const double* vp, const double* q6, const double* u3, int num_iters, double* out2) {
    lateral_tire_forces_qu(vp_from(vp), q6, u3, out2[0], out2[1], num_iters);
}
double orc_fiala(double alpha, double Ca, double mu, double Fx, double Fz) { return fialatiremodel<double>(alpha, Ca, mu, Fx, Fz); }
double orc_invfiala(double Fy, double Ca, double Fy_max) { return _invfialatiremodel(Fy, Ca, Fy_max); }
void orc_stable_limits(const double* vp, double Ux, double Fxf, double Fxr, double* out14) {
    StableLimits S = stable_limits(vp_from(vp), Ux, Fxf, Fxr);
    out14[0] = S.delta_min; out14[1] = S.delta_max;
    for (int k = 0; k < 4; k++) { out14[2 + 2 * k] = S.H[k][0]; out14[3 + 2 * k] = S.H[k][1]; out14[10 + k] = S.G[k]; }
}
void orc_steady_state(const double* vp, double V, double A_tan, double kappa, int num_iters, double r, double beta0, double delta0,
                      double Fyf0, double* out8) {
    SteadyState S = steady_state_estimates(vp_from(vp), V, A_tan, kappa, num_iters, r, beta0, delta0, Fyf0);
    out8[0] = S.beta; out8[1] = S.Ux; out8[2] = S.Uy; out8[3] = S.r; out8[4] = S.A; out8[5] = S.delta; out8[6] = S.Fxf; out8[7] = S.Fxr;
}
double orc_adiff(double x, double y) { return adiff(x, y); }

// ---- flow / linearization ----
void orc_flow(int kind, const double* vp, const double* x, double dt, const double* up0, const double* upf, int nsub, double* out) {
    flow_rk4<double>(kind, vp_from(vp), x, dt, up0, upf, out, nsub);
}
void orc_linearize_flow(int kind, const double* vp, const double* x, double dt, const double* up0, const double* upf, int ramp, int nk,
                        double* A, double* B0, double* Bf, double* c) {
    DiscreteLin R = linearize_flow(kind, vp_from(vp), x, dt, up0, upf, ramp != 0, nk);
    std::memcpy(A, R.A, sizeof(double) * R.nx * R.nx); std::memcpy(B0, R.B0, sizeof(double) * R.nx * nk);
    std::memcpy(Bf, R.Bf, sizeof(double) * R.nx * nk); std::memcpy(c, R.c, sizeof(double) * R.nx);
}
void orc_linearize_continuous(int kind, const double* vp, const double* x, const double* up, double* A, double* B, double* f) {
    ContinuousLin R = linearize_continuous(kind, vp_from(vp), x, up);
    std::memcpy(A, R.A, sizeof(double) * R.nx * R.nx); std::memcpy(B, R.B, sizeof(double) * R.nx * 6); std::memcpy(f, R.f, sizeof(double) * R.nx);
}
void orc_linearize_exact(int kind, const double* vp, const double* x, double dt, const double* up0, const double* upf, int ramp, int nk,
                         double* A, double* B0, double* Bf, double* c) {
    VehicleParams P = vp_from(vp);
    ContinuousLin CL = linearize_continuous(kind, P, x, up0);
    DiscreteLin R = linearize_exact(CL, x, dt, up0, upf, ramp != 0, nk);
    std::memcpy(A, R.A, sizeof(double) * R.nx * R.nx); std::memcpy(B0, R.B0, sizeof(double) * R.nx * nk);
    std::memcpy(Bf, R.Bf, sizeof(double) * R.nx * nk); std::memcpy(c, R.c, sizeof(double) * R.nx);
}
void orc_expm(int n, const double* A, double* E) { expm(n, A, E); }

// ---- trajectories ----
void* orc_traj_create(int n, const double* t, const double* s, const double* V, const double* A, const double* E, const double* N,
                      const double* psi, const double* kappa, const double* theta, const double* phi, const double* eL, const double* eR) {
    TrajectoryTube* T = new TrajectoryTube();
    T->t.assign(t, t + n); T->s.assign(s, s + n); T->V.assign(V, V + n); T->A.assign(A, A + n); T->E.assign(E, E + n); T->N.assign(N, N + n);
    T->psi.assign(psi, psi + n); T->kappa.assign(kappa, kappa + n); T->theta.assign(theta, theta + n); T->phi.assign(phi, phi + n);
    T->edge_L.assign(eL, eL + n); T->edge_R.assign(eR, eR + n);
    return T;
}
void orc_traj_free(void* h) { delete (TrajectoryTube*)h; }
static void node_out(const TrajectoryNode& o, double* out12) {
    out12[0] = o.t; out12[1] = o.s; out12[2] = o.V; out12[3] = o.A; out12[4] = o.E; out12[5] = o.N; out12[6] = o.psi; out12[7] = o.kappa;
    out12[8] = o.theta; out12[9] = o.phi; out12[10] = o.edge_L; out12[11] = o.edge_R;
}
void orc_traj_at_time(void* h, double t, double* out12) { node_out(((TrajectoryTube*)h)->at_time(t), out12); }
void orc_traj_at_s(void* h, double s, double* out12) { node_out(((TrajectoryTube*)h)->at_s(s), out12); }
void orc_traj_path_coordinates(void* h, double x, double y, double* out3) { ((TrajectoryTube*)h)->path_coordinates(x, y, out3[0], out3[1], out3[2]); }
void orc_invcumtrapz(int n, const double* y, const double* x, double* out) {
    std::vector<double> r = invcumtrapz(std::vector<double>(y, y + n), std::vector<double>(x, x + n));
    std::memcpy(out, r.data(), sizeof(double) * n);
}

// ---- HJI ----
void* orc_hji_create(const int* dims, const float* knots_concat, const float* V, const float* gradV) {
    HjiCache* C = new HjiCache();
    size_t off = 0;
    for (int d = 0; d < 7; d++) { C->dims[d] = dims[d]; C->knots[d].assign(knots_concat + off, knots_concat + off + dims[d]); off += dims[d]; }
    size_t nn = C->n_nodes();
    C->V.assign(V, V + nn); C->gradV.assign(gradV, gradV + 7 * nn);
    return C;
}
void* orc_hji_placeholder() { return new HjiCache(placeholder_hji()); }
void orc_hji_free(void* h) { delete (HjiCache*)h; }
void orc_hji_lookup(void* h, int M, const double* x7, double* V, double* gV7) {
    for (int i = 0; i < M; i++) hji_lookup(*(HjiCache*)h, x7 + 7 * i, V[i], gV7 + 7 * i);
}
void orc_hji_relative_state(const double* us6, const double* them4, double* x7) { hji_relative_state(us6, them4, x7); }
void orc_optimal_disturbance(const double* vp, const double* x7, const double* gV7, double* uH2) { optimal_disturbance(vp_from(vp), x7, gV7, uH2); }
void orc_optimal_control(const double* vp, const double* x7, const double* gV7, double* uR2) { optimal_control(vp_from(vp), x7, gV7, uR2); }
void orc_reachability_constraint(const double* vp, void* h, const double* x7, double eps, const double* uR2, double* M2, double* b) {
    reachability_constraint(vp_from(vp), *(HjiCache*)h, x7, eps, uR2, M2, *b);
}

// ---- generic OSQP-style solver ----
void* orc_osqp_create(int n, int m, const int* Pp, const int* Pi, const double* Px, const double* q, const int* Ap, const int* Ai,
                      const double* Ax, const double* l, const double* u, const double* settings) {
    Csc P, A;
    P.nrow = P.ncol = n; P.p.assign(Pp, Pp + n + 1); P.i.assign(Pi, Pi + Pp[n]); P.x.assign(Px, Px + Pp[n]);
    A.nrow = m; A.ncol = n; A.p.assign(Ap, Ap + n + 1); A.i.assign(Ai, Ai + Ap[n]); A.x.assign(Ax, Ax + Ap[n]);
    OsqpSolver* S = new OsqpSolver();
    S->setup(P, std::vector<double>(q, q + n), A, std::vector<double>(l, l + m), std::vector<double>(u, u + m), st_from(settings));
    return S;
}
void orc_osqp_free(void* h) { delete (OsqpSolver*)h; }
void orc_osqp_update(void* h, const double* Px, const double* Ax, const double* q, const double* l, const double* u) { ((OsqpSolver*)h)->update(Px, Ax, q, l, u); }
void orc_osqp_warm_start(void* h, const double* x, const double* y) { ((OsqpSolver*)h)->warm_start(x, y); }
void orc_osqp_cold_start(void* h) { ((OsqpSolver*)h)->cold_start(); }
// info_out: [status, iter, pri_res, dua_res, rho, rho_updates, n_factor]
int orc_osqp_solve(void* h, double* x, double* y, double* info_out) {
    OsqpSolver* S = (OsqpSolver*)h;
    int st = S->solve();
    if (x) std::memcpy(x, S->sol_x.data(), sizeof(double) * S->n);
    if (y) std::memcpy(y, S->sol_y.data(), sizeof(double) * S->m);
    if (info_out) { info_out[0] = st; info_out[1] = S->info.iter; info_out[2] = S->info.pri_res; info_out[3] = S->info.dua_res; info_out[4] = S->st.rho; info_out[5] = S->info.rho_updates; info_out[6] = (double)S->n_factor; }
    return st;
}
void orc_osqp_get_scaling(void* h, double* D, double* E, double* c) {
    OsqpSolver* S = (OsqpSolver*)h;
    std::memcpy(D, S->D.data(), sizeof(double) * S->n); std::memcpy(E, S->E.data(), sizeof(double) * S->m); *c = S->c;
}
void orc_osqp_get_iterates(void* h, double* x, double* z, double* y) {
    OsqpSolver* S = (OsqpSolver*)h;
    std::memcpy(x, S->x.data(), sizeof(double) * S->n); std::memcpy(z, S->z.data(), sizeof(double) * S->m); std::memcpy(y, S->y.data(), sizeof(double) * S->m);
}

// ---- MPC controllers ----
void* orc_mpc_create(int kind, const double* vp, const double* cp, int N_short, int N_long, double dt_short, double dt_long, int corr,
                     const double* settings) {
    return new Mpc(kind, vp_from(vp), cp_from(cp), N_short, N_long, dt_short, dt_long, corr != 0, st_from(settings));
}
void orc_mpc_free(void* h) { delete (Mpc*)h; }
void orc_mpc_dims(void* h, int* out) { Mpc* M = (Mpc*)h; out[0] = M->N; out[1] = M->nx; out[2] = M->nu; out[3] = M->n; out[4] = M->m; out[5] = (int)M->Am.x.size(); }
void orc_mpc_set_trajectory(void* h, void* traj) { ((Mpc*)h)->traj = *(TrajectoryTube*)traj; }
void orc_mpc_set_hji(void* h, void* hji, double eps) { ((Mpc*)h)->hji = *(HjiCache*)hji; ((Mpc*)h)->hji_eps = eps; }
int orc_mpc_from_autobox(void* h, const double* q6, const double* u3, const double* other4, double stamp, double pause_speed, int nan_fallback, double* out5) {
    return ((Mpc*)h)->from_autobox(q6, u3, other4, stamp, pause_speed, nan_fallback != 0, out5) ? 1 : 0;
}
void orc_mpc_set_hji_policy(void* h, int on) { ((Mpc*)h)->use_hji_policy = on != 0; }
void orc_mpc_get_hji_values(void* h, double* V, double* gV7) { Mpc* M = (Mpc*)h; *V = M->hji_V; std::memcpy(gV7, M->hji_gradV, 56); }
void orc_mpc_set_state(void* h, const double* q6, const double* u3, const double* other4, double time_offset) {
    Mpc* M = (Mpc*)h;
    if (q6) std::memcpy(M->state, q6, 48);
    if (u3) std::memcpy(M->control, u3, 24);
    if (other4) std::memcpy(M->other_car, other4, 32);
    M->time_offset = time_offset;
}
void orc_mpc_get_state(void* h, double* q6, double* u3) { Mpc* M = (Mpc*)h; std::memcpy(q6, M->state, 48); std::memcpy(u3, M->control, 24); }
void orc_mpc_set_solved(void* h, int solved) { ((Mpc*)h)->solved = solved != 0; }
void orc_mpc_reset_solver(void* h) { ((Mpc*)h)->reset_solver(); }
void orc_mpc_compute_time_steps(void* h, double t0) { ((Mpc*)h)->compute_time_steps(t0); }
void orc_mpc_get_time_steps(void* h, double* ts, double* dt, double* prev_ts) {
    Mpc* M = (Mpc*)h;
    std::memcpy(ts, M->TS.ts.data(), 8 * M->N); std::memcpy(dt, M->TS.dt.data(), 8 * M->T); std::memcpy(prev_ts, M->TS.prev_ts.data(), 8 * M->N);
}
void orc_mpc_compute_linearization_nodes(void* h) { ((Mpc*)h)->compute_linearization_nodes(); }
void orc_mpc_get_nodes(void* h, double* qs, double* us, double* ps) {
    Mpc* M = (Mpc*)h;
    std::memcpy(qs, M->qs.data(), 8 * M->qs.size()); std::memcpy(us, M->us.data(), 8 * M->us.size()); std::memcpy(ps, M->ps.data(), 8 * M->ps.size());
}
void orc_mpc_set_nodes(void* h, const double* qs, const double* us, const double* ps) {
    Mpc* M = (Mpc*)h;
    std::memcpy(M->qs.data(), qs, 8 * M->qs.size()); std::memcpy(M->us.data(), us, 8 * M->us.size()); std::memcpy(M->ps.data(), ps, 8 * M->ps.size());
}
void orc_mpc_update_qp(void* h) { ((Mpc*)h)->update_qp(); }
// per-interval pieces: A[T*nx*nx], B0[T*nx*nu], Bf[T*nx*nu], c[T*nx], H[T*8], G[T*4], dmin[T], dmax[T], fxmax[T], hji[3] = (M1, M2, b)
void orc_mpc_get_qp_pieces(void* h, double* A, double* B0, double* Bf, double* c, double* H, double* G, double* dmin, double* dmax,
                           double* fxmax, double* hji3) {
    Mpc* M = (Mpc*)h;
    int nx = M->nx, nu = M->nu;
    for (int t = 0; t < M->T; t++) {
        std::memcpy(A + t * nx * nx, M->lin[t].A, 8 * nx * nx);
        std::memcpy(B0 + t * nx * nu, M->lin[t].B0, 8 * nx * nu);
        std::memcpy(Bf + t * nx * nu, M->lin[t].Bf, 8 * nx * nu);
        std::memcpy(c + t * nx, M->lin[t].c, 8 * nx);
        for (int k = 0; k < 4; k++) { H[t * 8 + 2 * k] = M->env[t].H[k][0]; H[t * 8 + 2 * k + 1] = M->env[t].H[k][1]; G[t * 4 + k] = M->env[t].G[k]; }
        dmin[t] = M->dmin[t]; dmax[t] = M->dmax[t]; fxmax[t] = M->fxmax[t];
    }
    hji3[0] = M->M_hji[0]; hji3[1] = M->M_hji[1]; hji3[2] = M->b_hji;
}
// canonical QP: Pdiag[n], q[n], A as CSC (Ap[n+1], Ai[nnz], Ax[nnz]), l[m], u[m]
void orc_mpc_get_qp(void* h, double* Pdiag, double* q, int* Ap, int* Ai, double* Ax, double* l, double* u) {
    Mpc* M = (Mpc*)h;
    std::memcpy(Pdiag, M->Pm.x.data(), 8 * M->n); std::memcpy(q, M->qv.data(), 8 * M->n);
    std::memcpy(Ap, M->Am.p.data(), 4 * (M->n + 1)); std::memcpy(Ai, M->Am.i.data(), 4 * M->Am.i.size()); std::memcpy(Ax, M->Am.x.data(), 8 * M->Am.x.size());
    std::memcpy(l, M->lv.data(), 8 * M->m); std::memcpy(u, M->uv.data(), 8 * M->m);
}
int orc_mpc_solve(void* h) { return ((Mpc*)h)->solve(); }
void orc_mpc_get_solution(void* h, double* x, double* y) {
    Mpc* M = (Mpc*)h;
    if (x) std::memcpy(x, M->solver.sol_x.data(), 8 * M->n);
    if (y) std::memcpy(y, M->solver.sol_y.data(), 8 * M->m);
}
// stats: [status, iter, pri_res, dua_res, rho, rho_updates]
void orc_mpc_get_stats(void* h, double* s) {
    Mpc* M = (Mpc*)h;
    s[0] = M->solver.info.status; s[1] = M->solver.info.iter; s[2] = M->solver.info.pri_res; s[3] = M->solver.info.dua_res; s[4] = M->solver.st.rho; s[5] = M->solver.info.rho_updates;
}
void orc_mpc_get_next_control(void* h, double* out3) { ((Mpc*)h)->get_next_control(out3); }
void orc_mpc_simulate_step(void* h, double t, double dt) { ((Mpc*)h)->simulate_step(t, dt); }
// full step for many independent controllers (CPU baseline): time_steps -> nodes -> update_qp -> solve -> control [-> plant rollout]
void orc_mpc_batch_step(void** hs, int count, const double* t0, double* controls3, int rollout, double dt_sim, int nthreads) {
    if (nthreads <= 0) nthreads = (int)std::thread::hardware_concurrency();
    if (nthreads < 1) nthreads = 1;
    if (nthreads > count) nthreads = count;
    std::atomic<int> next(0);
    auto work = [&]() {
        for (;;) {
            int i = next.fetch_add(1);
            if (i >= count) break;
            Mpc* M = (Mpc*)hs[i];
            if (rollout) M->simulate_step(t0[i], dt_sim);
            else { M->compute_time_steps(t0[i]); M->compute_linearization_nodes(); M->update_qp(); M->solve(); }
            if (controls3) { if (rollout) std::memcpy(controls3 + 3 * i, M->control, 24); else M->get_next_control(controls3 + 3 * i); }
        }
    };
    if (nthreads == 1) { work(); return; }
    std::vector<std::thread> th;
    for (int k = 0; k < nthreads; k++) th.emplace_back(work);
    for (auto& t : th) t.join();
}
int orc_max_threads() { int n = (int)std::thread::hardware_concurrency(); return n < 1 ? 1 : n; }

}  // extern "C"
