// pgn_capi.cu — the extern "C" boundary of libpigeon_b200.so (declared in include/pigeon_b200.h).
// Host arrays in, host arrays out; CUDA errors become return codes; no CPU compute path exists behind these calls.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "pgn_internal.h"

using namespace pgn;

static thread_local char g_err[512] = "";
static int set_err(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
extern "C" void pgn_set_error_(const char* msg) { snprintf(g_err, sizeof(g_err), "%s", msg); }      // for the other translation units
#define CK(call)                                                                                                              \
    do {                                                                                                                      \
        cudaError_t e__ = (call);                                                                                             \
        if (e__ != cudaSuccess) return set_err(PGN_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)
#define REQUIRE(cond, msg)                                   \
    do {                                                     \
        if (!(cond)) return set_err(PGN_EINVAL, "%s", msg);  \
    } while (0)

extern "C" { static int finish_sim(pgn_handle* h); static void drop_round_graph(pgn_handle* h, int p); }      // defined beside the simulate loop, inside the extern "C" block
static void drain_ring(pgn_handle* h) {
    for (int i = 0, sl = h->ring_tail; i < h->ring_count; i++, sl = (sl + 1) % PGN_RING)
        for (int p = 0; p < h->ring_parts[sl]; p++) cudaEventSynchronize(h->ring_done[sl][p]);
}

namespace {

// Every entry point runs on the handle's own device and leaves the caller's current device as it found it, so that ONE host thread can
// drive handles on several GPUs (the Julia deployment: a single process, one handle per GPU).
struct DeviceGuard {
    int prev; bool changed;
    explicit DeviceGuard(int dev) : prev(-1), changed(false) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) changed = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() { if (changed) cudaSetDevice(prev); }
};
// steps submitted with pgn_step_submit run on the part streams: every other entry point first waits for them (their results stay collectable)
// ... and for the vehicles that a simulate loop with deferred solves left behind (finish_sim: catch-up rounds until every vehicle has its steps)
#define ENTER_NODRAIN(h, msg)  REQUIRE(h, msg); DeviceGuard dev_guard__((h)->device)
#define ENTER(h, msg)  ENTER_NODRAIN(h, msg); if ((h)->ring_count) drain_ring(h); if ((h)->sim_open) { int rc__ = finish_sim(h); if (rc__) return rc__; }

// profiling: 1 = stage timers (the stages run serially, one part), 2 = + in-kernel cycle counters of the ADMM kernel, 3 = the cycle counters alone
// (pipeline parts and graphs stay on: the counters then describe the free-running loop)
static inline bool serial_profiling(const pgn_handle* h) { return h->profiling == 1 || h->profiling == 2; }
struct StageTimer {
    pgn_handle* h; int idx;
    StageTimer(pgn_handle* h_, int idx_) : h(h_), idx(idx_) { if (serial_profiling(h)) cudaEventRecord(h->ev[0], h->stream); }
    ~StageTimer() {
        if (serial_profiling(h)) {
            cudaEventRecord(h->ev[1], h->stream);
            cudaEventSynchronize(h->ev[1]);
            float ms = 0;
            cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]);
            h->stage_ms[idx] += ms;
        }
    }
};


template <class T>
int dev_alloc(pgn_handle* h, T** p, size_t count) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, count * sizeof(T) + 16);
    if (e != cudaSuccess) return set_err(PGN_ENOMEM, "cudaMalloc(%zu bytes) failed: %s", count * sizeof(T), cudaGetErrorString(e));
    h->allocs.push_back(q);
    *p = (T*)q;
    return PGN_OK;
}
template <class T>
int dev_upload(pgn_handle* h, const std::vector<T>& v, const T** out) {
    T* p = nullptr;
    int rc = dev_alloc(h, &p, v.size() + 1);
    if (rc) return rc;
    if (!v.empty()) CK(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    *out = p;
    return PGN_OK;
}

int ensure_stage(pgn_handle* h, size_t bytes) {
    if (h->stage_bytes >= bytes) return PGN_OK;
    double* p = nullptr;
    int rc = dev_alloc(h, &p, bytes / 8 + 1);
    if (rc) return rc;
    h->d_stage = p; h->stage_bytes = bytes;
    return PGN_OK;
}

int download_aos(pgn_handle* h, const double* d_soa, double* host, int k) {
    size_t bytes = (size_t)h->B * k * sizeof(double);
    int rc = ensure_stage(h, bytes);
    if (rc) return rc;
    launch_transpose_out(h, d_soa, h->d_stage, k);
    CK(cudaMemcpyAsync(host, h->d_stage, bytes, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return PGN_OK;
}

void refresh_constants(pgn_handle* h, double* ctab, double* wtab) {
    const VehParams& P = h->veh;
    const CtrlParams& C = h->ctl;
    h->un[0] = P.delta_max;
    h->un[1] = fmax(-P.Fx_min, P.Fx_max);
    const bool cpl = h->cfg.kind == PGN_COUPLED;
    ctab[CT_ZERO] = 0.0; ctab[CT_PINF] = INFINITY; ctab[CT_NINF] = -INFINITY;
    ctab[CT_VMIN] = C.V_min; ctab[CT_VMAX] = C.V_max; ctab[CT_FXMIN_N] = P.Fx_min / h->un[1];
    ctab[CT_DDELTA_N] = cpl ? C.ddelta_max / h->un[0] : C.ddelta_max;
    for (int i = 0; i < W_LEN; i++) wtab[i] = 0.0;
    wtab[W_Q_DS] = C.Q_ds; wtab[W_Q_DPSI] = C.Q_dpsi; wtab[W_Q_E] = C.Q_e; wtab[W_R_DELTA] = C.R_delta; wtab[W_R_FX] = C.R_Fx;
    wtab[W_R_DDELTA] = C.R_ddelta; wtab[W_R_DFX] = C.R_dFx; wtab[W_W_BETA] = C.W_beta; wtab[W_W_R] = C.W_r; wtab[W_W_HJI] = C.W_HJI;
}
int upload_constants(pgn_handle* h) {
    double ctab[CT_LEN], wtab[W_LEN];
    refresh_constants(h, ctab, wtab);
    CK(cudaMemcpyAsync((void*)h->qd.ctab, ctab, sizeof(ctab), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync((void*)h->qd.wtab, wtab, sizeof(wtab), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->qd.n_hji = (int)h->ctl.N_HJI;
    return PGN_OK;
}

void x1_params(double* vp) {
    VehParams P;
    P.G = 9.80665;
    const double mfl = 484, mfr = 455, mrl = 521, mrr = 504;
    P.m = mfl + mfr + mrl + mrr; P.Izz = 2900; P.L = 2.87;
    P.a = (mrl + mrr) / P.m * P.L; P.b = (mfl + mfr) / P.m * P.L;
    P.h = 0.1 * P.b / P.L + 0.1 * P.a / P.L + 0.37;
    P.mu = 0.92; P.Caf = 150e3; P.Car = 220e3; P.Fx_max = 5600; P.Px_max = 75e3; P.Cd0 = 241.0; P.Cd1 = 25.1; P.Cd2 = 0.0;
    P.fwd_frac = 0.0; P.rwd_frac = 1 - P.fwd_frac; P.fwb_frac = 0.6; P.rwb_frac = 1 - P.fwb_frac;
    const double f1 = -P.m * P.G * P.a * P.mu / (P.L * P.rwb_frac + P.mu * P.h), f2 = -P.m * P.G * P.b * P.mu / (P.L * P.fwb_frac - P.mu * P.h);
    P.Fx_min = f1 > f2 ? f1 : f2;
    P.delta_max = 18 * M_PI / 180; P.kappa_max = tan(P.delta_max) / P.L; P.inv_fiala_corrected = 0.0;
    memcpy(vp, &P, sizeof(double) * PGN_VEHICLE_PARAMS_LEN);
}
void default_control(int kind, double* c) {
    const double d10 = 10 * M_PI / 180;
    double v[16] = {1.0, 15.0, 10.0 / 4 / 100, 10.0 / 4 / 10000, 0.344, 1.0, 1.0, 1.0, 50 / d10, 50.0, 500.0, 3, 0.0, 0.1, 0.0, 0.5};
    if (kind == PGN_DECOUPLED) { v[6] = 1 / (d10 * d10); v[7] = 1.0; v[12] = 0.0; v[13] = 0.01 / (d10 * d10); }
    memcpy(c, v, sizeof(v));
}

int set_hji_internal(pgn_handle* h, const int32_t dims[7], const float* knots, const float* V, const float* gradV) {
    size_t nn = 1, nk = 0;
    for (int d = 0; d < 7; d++) { REQUIRE(dims[d] >= 2, "HJI grid needs >= 2 knots per dimension"); nn *= dims[d]; nk += dims[d]; }
    std::vector<float> recs(nn * 8);
    for (size_t i = 0; i < nn; i++) {
        for (int k = 0; k < 7; k++) recs[i * 8 + k] = gradV[i * 7 + k];
        recs[i * 8 + 7] = V[i];
    }
    float *d_k = nullptr, *d_r = nullptr;
    CK(cudaStreamSynchronize(h->stream));            // kernels in flight may still read the grid that is replaced below
    if (h->hji.valid) {                              // release the previous grid (319 MB for the real cache)
        for (const void* old : {(const void*)h->hji.knots, (const void*)h->hji.gV})
            for (size_t i = 0; i < h->allocs.size(); i++) if (h->allocs[i] == old) { h->allocs.erase(h->allocs.begin() + i); cudaFree((void*)old); break; }
        h->hji.valid = 0;
    }
    int rc = dev_alloc(h, &d_k, nk + 8); if (rc) return rc;
    rc = dev_alloc(h, &d_r, nn * 8 + 8); if (rc) return rc;
    CK(cudaMemcpy(d_k, knots, nk * sizeof(float), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_r, recs.data(), recs.size() * sizeof(float), cudaMemcpyHostToDevice));
    HjiView& H = h->hji;
    int off = 0;
    long long st = 1;
    for (int d = 0; d < 7; d++) { H.dims[d] = dims[d]; H.kofs[d] = off; off += dims[d]; H.stride[d] = st; st *= dims[d]; }
    H.knots = d_k; H.V = nullptr; H.gV = d_r; H.valid = 1;
    hji_make_tensor_map(h);               // optional: without it the large lookups stay on the plain cell-ordered gather
    return PGN_OK;
}

// Runs `body` once per pipeline part: the part's vehicle range, stream, side stream and fork / join events are swapped into the handle
// (the launchers read them from there), every part stream starts after what is already queued on the caller's stream, and the caller's
// stream continues after all parts.  One part (or stage timers on): `body` runs as is on the caller's stream.
template <class F>
int for_each_part(pgn_handle* h, F body) {
    if (h->parts <= 1 || serial_profiling(h)) return body();
    cudaStream_t s0 = h->stream, side0 = h->side_stream;
    cudaEvent_t f0 = h->ev_fork, j0 = h->ev_join;
    int rc = PGN_OK;
    cudaEventRecord(h->part_begin, s0);
    for (int p = 0; p < h->parts && rc == PGN_OK; p++) {
        cudaStreamWaitEvent(h->part_stream[p], h->part_begin, 0);
        h->stream = h->part_stream[p]; h->side_stream = h->part_side[p]; h->ev_fork = h->part_evf[p]; h->ev_join = h->part_evj[p];
        h->part = p; h->v0 = (int)((long long)h->B * p / h->parts); h->nv = (int)((long long)h->B * (p + 1) / h->parts) - h->v0;
        rc = body();
        cudaEventRecord(h->part_done[p], h->part_stream[p]);
    }
    h->stream = s0; h->side_stream = side0; h->ev_fork = f0; h->ev_join = j0;
    h->part = 0; h->v0 = 0; h->nv = h->B;
    for (int p = 0; p < h->parts; p++) cudaStreamWaitEvent(s0, h->part_done[p], 0);
    return rc;
}

}  // namespace

extern "C" {

const char* pgn_last_error(void) { return g_err; }

int pgn_default_config(pgn_config* c, int32_t kind) {
    REQUIRE(c, "cfg is NULL");
    memset(c, 0, sizeof(*c));
    c->kind = kind; c->batch = 1; c->N_short = 10; c->N_long = 20; c->dt_short = 0.01; c->dt_long = 0.2; c->use_correction_step = 1; c->device = -1;
    c->rho = 0.1; c->sigma = 1e-6; c->alpha = 1.6; c->eps_abs = 1e-3; c->eps_rel = 1e-3; c->eps_prim_inf = 1e-4; c->eps_dual_inf = 1e-4;
    c->max_iter = 4000; c->scaling = 10; c->check_termination = 25; c->adaptive_rho = 1; c->adaptive_rho_interval = 25; c->adaptive_rho_tolerance = 5.0;
    c->warm_start = 1; c->rk4_substeps = 10; c->hji_eps = 0.05; c->kkt_ordering = 0;
    return PGN_OK;
}
int pgn_x1_vehicle_params(double* vp) { REQUIRE(vp, "vp is NULL"); x1_params(vp); return PGN_OK; }
int pgn_default_control_params(int32_t kind, double* cp) { REQUIRE(cp, "cp is NULL"); default_control(kind, cp); return PGN_OK; }

int pgn_create(const pgn_config* cfg, pgn_handle** out) {
    REQUIRE(cfg && out, "NULL argument");
    REQUIRE(cfg->kind == PGN_COUPLED || cfg->kind == PGN_DECOUPLED, "kind must be PGN_COUPLED or PGN_DECOUPLED");
    REQUIRE(cfg->batch >= 1, "batch must be >= 1");
    REQUIRE(cfg->N_short >= 1 && cfg->N_long >= 0, "N_short >= 1 and N_long >= 0 required");
    REQUIRE(cfg->rk4_substeps >= 1, "rk4_substeps must be >= 1");
    int ndev = 0;
    cudaError_t e0 = cudaGetDeviceCount(&ndev);
    if (e0 != cudaSuccess || ndev == 0) return set_err(PGN_ECUDA, "no CUDA device available (%s); libpigeon_b200 has no CPU path", cudaGetErrorString(e0));
    int dev = cfg->device;
    if (dev < 0) CK(cudaGetDevice(&dev));
    REQUIRE(dev < ndev, "device ordinal out of range");
    DeviceGuard dev_guard__(dev);            // the caller's current device is restored on return
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    if (prop.major < 10) return set_err(PGN_ECUDA, "device %d is sm_%d%d; this library is built for sm_100a only", dev, prop.major, prop.minor);
    pgn_handle* h = new (std::nothrow) pgn_handle();
    if (!h) return set_err(PGN_ENOMEM, "out of host memory");
    h->cfg = *cfg; h->device = dev; h->B = cfg->batch; h->N = 1 + cfg->N_short + cfg->N_long; h->T = h->N - 1;
    h->nx = cfg->kind == PGN_COUPLED ? 6 : 4; h->nu = cfg->kind == PGN_COUPLED ? 2 : 1;
    h->num_sms = prop.multiProcessorCount;
    h->profiling = 0; h->launches = 0; memset(h->stage_ms, 0, sizeof(h->stage_ms));
    h->have_traj = false; h->have_assign = false; h->stage_bytes = 0; h->d_stage = nullptr;
    memset(&h->hji, 0, sizeof(h->hji)); memset(&h->traj, 0, sizeof(h->traj));
    char err[256];
    // ADMM build variant.  QP small enough for TWO 256-thread CTAs per SM with everything in shared memory (static tables read through
    // L1): v256.  Otherwise, if the factor fits one CTA's half of TENSOR MEMORY and the rest fits shared memory twice: vtm (the coupled
    // N = 31 QP).  Otherwise one 512-thread CTA per SM with the tables in shared memory: v512.  PGN_ADMM_VARIANT=512|256|tmem forces one.
    {
        const char* force = getenv("PGN_ADMM_VARIANT");
        const int want = !force ? 0 : (!strcmp(force, "tmem") ? 3 : atoi(force));
        h->admm_threads = 512; h->admm_tmem = 0; h->d_admm_scratch = nullptr;
        if (want != 512) {
            if (!build_qp_tables(cfg->kind, cfg->N_short, cfg->N_long, cfg->kkt_ordering, h->tab, err, sizeof(err), 8, want == 3 ? 1 : 0)) { delete h; return set_err(PGN_EINVAL, "QP analysis failed: %s", err); }
            const size_t sm = admm_smem_bytes(h->tab, 256, false);
            if (want != 3 && 2 * (sm + 1024) <= (size_t)227 * 1024) h->admm_threads = 256;
            else if (want == 256) { delete h; return set_err(PGN_EINVAL, "PGN_ADMM_VARIANT=256: two CTAs of %zu bytes do not fit one SM", sm); }
            else {
                const int tm_threads = 256;      // 384 / 512 threads per QP were measured and lose (pgn_admm.cu)
                if ((!h->tab.tmem_layout || tm_threads != 256) && !build_qp_tables(cfg->kind, cfg->N_short, cfg->N_long, cfg->kkt_ordering, h->tab, err, sizeof(err), tm_threads / 32, 1)) { delete h; return set_err(PGN_EINVAL, "QP analysis failed: %s", err); }
                if (admm_tmem_fits(h->tab, tm_threads)) { h->admm_threads = tm_threads; h->admm_tmem = 1; }
                else if (want == 3) { delete h; return set_err(PGN_EINVAL, "PGN_ADMM_VARIANT=tmem: this QP does not fit (TMEM columns %d, shared memory %zu bytes per CTA)", h->tab.tmem_cols, admm_smem_bytes_tmem(h->tab, 256)); }
            }
        }
        if (h->admm_threads == 512 && !h->admm_tmem &&
            !build_qp_tables(cfg->kind, cfg->N_short, cfg->N_long, cfg->kkt_ordering, h->tab, err, sizeof(err), 16)) { delete h; return set_err(PGN_EINVAL, "QP analysis failed: %s", err); }
    }
    if (4 * ((h->tab.tail_dim + 7) & ~7) + 2 > 2 * ((h->tab.Nk + 2) & ~1)) { delete h; return set_err(PGN_EINVAL, "dense tail of dimension %d: the pivot-column buffer of its sweep does not fit two vectors of %d KKT rows", h->tab.tail_dim, h->tab.Nk); }
    if (pgn::tail_segments(h->tab.tail_dim) > h->admm_threads) { delete h; return set_err(PGN_EINVAL, "dense tail of dimension %d needs more than %d ADMM threads", h->tab.tail_dim, h->admm_threads); }
    if (h->tab.Nk > (h->admm_threads >= 512 ? 3 : 5) * h->admm_threads) { delete h; return set_err(PGN_EINVAL, "QP too large: %d KKT rows for %d ADMM threads", h->tab.Nk, h->admm_threads); }
    auto bail = [&](int rc) { pgn_destroy(h); return rc; };
    h->prio_least = 0; h->prio_greatest = 0;
    cudaDeviceGetStreamPriorityRange(&h->prio_least, &h->prio_greatest);
    { const char* ev = getenv("PGN_ADMM_PRIORITY"); h->admm_low_priority = ev ? atoi(ev) != 0 : 1; }
    cudaError_t e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete h; return set_err(PGN_ECUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(e)); }
    h->stream = h->own_stream;
    cudaEventCreate(&h->ev[0]); cudaEventCreate(&h->ev[1]);
    h->side_stream = nullptr;
    h->parts = 1; h->parts_created = 0; h->v0 = 0; h->nv = h->B; h->part = 0;
    cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming); cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming);
    double vp[PGN_VEHICLE_PARAMS_LEN], cp[PGN_CONTROL_PARAMS_LEN];
    x1_params(vp); default_control(cfg->kind, cp);
    memcpy(&h->veh, vp, sizeof(vp)); memcpy(&h->ctl, cp, sizeof(cp));
    h->st.rho = cfg->rho; h->st.sigma = cfg->sigma; h->st.alpha = cfg->alpha; h->st.eps_abs = cfg->eps_abs; h->st.eps_rel = cfg->eps_rel;
    h->st.eps_prim_inf = cfg->eps_prim_inf; h->st.eps_dual_inf = cfg->eps_dual_inf; h->st.adaptive_rho_tolerance = cfg->adaptive_rho_tolerance;
    h->st.max_iter = cfg->max_iter; h->st.scaling = cfg->scaling; h->st.check_termination = cfg->check_termination; h->st.adaptive_rho = cfg->adaptive_rho;
    h->st.adaptive_rho_interval = cfg->adaptive_rho_interval; h->st.warm_start = cfg->warm_start;

    // static tables -> device
    const QpTables& t = h->tab;
    QpDev& q = h->qd;
    memset(&q, 0, sizeof(q));
    q.kind = t.kind; q.N = t.N; q.T = t.T; q.Ns = t.Ns; q.nx = t.nx; q.nu = t.nu; q.n = t.n; q.m = t.m; q.Nk = t.Nk; q.nnzA = t.nnzA; q.nnzL = t.nnzL; q.nlev = t.nlev;
    q.rec_len = t.rec.rec_len; q.o_dt = t.rec.o_dt; q.var_u1_delta = t.var_u1_delta; q.var_u1_fx = t.var_u1_fx;
    int rc, rc_;
#define UP(field) if ((rc = dev_upload(h, t.field, &q.field))) return bail(rc)
    UP(a_src); UP(a_slot); UP(l_type); UP(u_type); UP(l_idx); UP(u_idx); UP(P_mode); UP(q_mode); UP(P_w); UP(q_w); UP(P_t); UP(q_t);
    UP(q_hji_t); UP(pos_var); UP(pos_con); UP(pos2idx); UP(is_con); UP(kadj_ptr); UP(kadj_e); UP(kadj_nb);
    UP(fac_lvl_ptr); UP(fac_tgt); UP(inv_lvl_ptr); UP(inv_tgt);
    { std::vector<uint16_t> k0(t.bwd_k0); k0.resize(k0.size() + 64, 0); if ((rc = dev_upload(h, k0, &q.bwd_k0))) return bail(rc); }
    q.n_bwd_k0 = (int)t.bwd_k0.size(); q.bwd_k0_phase = t.bwd_k0_phase; q.bwd_k0_warp0 = t.bwd_k0_warp0; q.fwd_k0_end = t.fwd_k0_end; q.fac_k0_end = t.fac_k0_end;
    UP(rz_pos); UP(rz_idx); q.rz_prog = (t.rz_prog && h->admm_threads == RZP_NT) ? 1 : 0;
    {   // the solve tables are read in batches of four slot rows with the surplus masked AFTER the load: the device copies carry four slot
        // rows of padding (zero entries) so that the variant that reads them from global memory never leaves its allocations
        std::vector<uint16_t> orow(t.sol_orow), fidx(t.fidx);
        std::vector<uint32_t> bent(t.bent);
        orow.resize(orow.size() + 64, 0); fidx.resize(fidx.size() + 128, (uint16_t)t.Nk);
        {   // tensor-memory variant: per-task TMEM column and the backward source positions (padded like the other solve tables)
            std::vector<uint16_t> tcol(t.sol_tcol), bsrc(t.bsrc);
            tcol.resize(tcol.size() + 8, 0); bsrc.resize(bsrc.size() + 128, (uint16_t)t.Nk);
            if ((rc = dev_upload(h, tcol, &q.sol_tcol))) return bail(rc);
            if ((rc = dev_upload(h, bsrc, &q.bsrc))) return bail(rc);
        }
        bent.resize(bent.size() + 128, (uint32_t)t.zslot | ((uint32_t)t.Nk << 16));
        if ((rc = dev_upload(h, orow, &q.sol_orow))) return bail(rc);
        if ((rc = dev_upload(h, fidx, &q.fidx))) return bail(rc);
        if ((rc = dev_upload(h, bent, &q.bent))) return bail(rc);
    }
    {
        std::vector<uint32_t> rc(t.nnzA);
        for (int e = 0; e < t.nnzA; e++) rc[e] = (uint32_t)t.a_rowpos[e] | ((uint32_t)t.a_colpos[e] << 16);
        if ((rc_ = dev_upload(h, rc, &q.a_rc))) return bail(rc_);
        auto pack = [](const std::vector<uint32_t>& w) {      // 4-word host tasks -> packed 8-byte device descriptors
            std::vector<uint2> o(w.size() / 4);
            for (size_t i = 0; i < o.size(); i++) {
                const uint32_t ebase = w[4 * i], rbase = w[4 * i + 1] & 0xffff, nrows = (w[4 * i + 1] >> 16) & 0xff, sh = w[4 * i + 1] >> 24, K = w[4 * i + 2] & 0xffff, fl = w[4 * i + 2] >> 16;
                o[i] = make_uint2((ebase >> 5) | (rbase << 16), K | (nrows << 8) | (sh << 16) | (fl << 24));
            }
            return o;
        };
        for (size_t i = 0; i < t.fac_task.size() / 4; i++) if ((t.fac_task[4 * i] >> 5) > 0xffff || (t.fac_task[4 * i + 1] & 0xffff) != (t.fac_task[4 * i + 1] & 0xffff)) return bail(set_err(PGN_EINVAL, "factor program too large"));
        if (t.fac_tgt.size() > 0xffff || t.inv_tgt.size() > 0xffff || t.fac_ent.size() / 32 > 0xffff || t.inv_ent.size() / 32 > 0xffff)
            return bail(set_err(PGN_EINVAL, "QP too large for the packed task descriptors"));
        if ((rc_ = dev_upload(h, pack(t.sol_task), &q.sol_task))) return bail(rc_);
        if ((rc_ = dev_upload(h, pack(t.fac_task), &q.fac_task))) return bail(rc_);
        if ((rc_ = dev_upload(h, pack(t.inv_task), &q.inv_task))) return bail(rc_);
        std::vector<unsigned long long> fe(t.fac_ent.begin(), t.fac_ent.end()), ie(t.inv_ent.begin(), t.inv_ent.end());
        const unsigned long long padent = (unsigned long long)t.zslot | ((unsigned long long)t.zslot << 16) | ((unsigned long long)t.Nk << 32);
        fe.resize(fe.size() + 512, padent); ie.resize(ie.size() + 512, padent);      // masked batches read whole groups of four slot rows past a task's end
        if ((rc_ = dev_upload(h, fe, &q.fac_ent))) return bail(rc_);
        if ((rc_ = dev_upload(h, ie, &q.inv_ent))) return bail(rc_);
    }
    q.nslots = t.nslots; q.zslot = t.zslot; q.rhs_tmp_end = t.rhs_tmp_end; q.n_fwd_ph = t.n_fwd_ph; q.n_bwd_ph = t.n_bwd_ph;
    q.n_sol_task = (int)t.sol_task.size() / 4; q.n_fac_task = (int)t.fac_task.size() / 4; q.n_inv_task = (int)t.inv_task.size() / 4;
    q.n_bent = (int)t.bent.size(); q.n_orow = (int)t.sol_orow.size(); q.n_orow_fwd = admm_orow_fwd(t); q.n_fac_lvl = (int)t.fac_lvl_ptr.size() - 1; q.n_inv_levels = (int)t.inv_lvl_ptr.size() - 1;
    q.tail_level = t.tail_level; q.tail_start = t.tail_start; q.tail_dim = t.tail_dim;
    if (t.n_fwd_ph + t.n_bwd_ph > ADMM_MAX_PHASES) return bail(set_err(PGN_EINVAL, "too many solve phases for this KKT ordering"));
#undef UP
    double *ctab = nullptr, *wtab = nullptr;
    if ((rc = dev_alloc(h, &ctab, CT_LEN + 1))) return bail(rc);
    if ((rc = dev_alloc(h, &wtab, W_LEN + 1))) return bail(rc);
    q.ctab = ctab; q.wtab = wtab;
    if ((rc = upload_constants(h))) return bail(rc);

    const size_t B = h->B, N = h->N, T = h->T;
#define AL(ptr, count) if ((rc = dev_alloc(h, &h->ptr, (count)))) return bail(rc)
    AL(d_state, 6 * B); AL(d_control, 3 * B); AL(d_other, 4 * B); AL(d_toff, B); AL(d_solved, B); AL(d_traj_id, B);
    AL(d_ts, B * N); AL(d_dt, B * T); AL(d_prev_ts, B * N);
    AL(d_qs, B * N * h->nx); AL(d_us, B * N * 2); AL(d_ps, B * N * 4);
    AL(d_rec, B * (size_t)t.rec.rec_len);
    AL(d_ws_xz, B * (size_t)t.Nk); AL(d_ws_y, B * (size_t)t.Nk); AL(d_rho, B);
    AL(d_sol_x, B * (size_t)t.n); AL(d_sol_y, B * (size_t)t.m);
    AL(d_iters, B); AL(d_status, B); AL(d_rho_updates, B); AL(d_pri_res, B); AL(d_dua_res, B);
    AL(d_controls, 3 * B); AL(d_t0, B); AL(d_t0_base, B); AL(d_counter, 2 * PGN_MAX_PARTS); AL(d_order, B); AL(d_skip, B); AL(d_cold, B); AL(d_cycles, 512); AL(d_hji_val, 8 * B); AL(d_io, 1 + 19 * B); AL(d_state_next, 6 * B); AL(d_last_seg, B); AL(d_se0, 2 * B); AL(d_se, 2 * B); AL(d_tskip, B); AL(d_in, 18 * B); AL(d_mask, B); AL(d_hold, B); AL(d_kstep, B); AL(d_iters_acc, B); AL(d_lag, 4);
#undef AL
    CK(cudaMemset(h->d_state, 0, 6 * B * 8)); CK(cudaMemset(h->d_control, 0, 3 * B * 8)); CK(cudaMemset(h->d_solved, 0, B)); CK(cudaMemset(h->d_traj_id, 0, B * 4));
    CK(cudaMemset(h->d_ws_xz, 0, B * t.Nk * 8)); CK(cudaMemset(h->d_ws_y, 0, B * t.Nk * 8));
    CK(cudaMemset(h->d_sol_x, 0, B * t.n * 8)); CK(cudaMemset(h->d_sol_y, 0, B * t.m * 8));
    CK(cudaMemset(h->d_cycles, 0, 4096)); CK(cudaMemset(h->d_skip, 0, B)); CK(cudaMemset(h->d_cold, 0, B));
    h->guard_nan = 0; h->guard_pause = 0.0; h->hji_policy = 0;
    h->path_window = 0; CK(cudaMemset(h->d_last_seg, 0xff, B * 4));
    h->in_callback = 0; h->cb_has_exec = 0; h->epoch = 1; h->cb_epoch = 0; h->cb_launches = 0; h->h_io = nullptr;
    CK(cudaMemset(h->d_tskip, 0, B)); CK(cudaMemset(h->d_se, 0, 2 * B * 8));
    h->hji_sort = getenv("PGN_HJI_SORT") ? atoi(getenv("PGN_HJI_SORT")) : -1; h->d_hji_ws = nullptr; h->hji_ws_bytes = 0; h->hji_tma_valid = 0;
    for (int p = 0; p < PGN_MAX_PARTS; p++) { h->rg_exec[p] = nullptr; h->rg_exec2[p] = nullptr; }
    { const char* ev = getenv("PGN_SPLIT_ROUNDS"); h->split_rounds = ev ? atoi(ev) != 0 : 0; }      // measured: no gain (below), off by default
    h->sim_axis_valid = 0; h->catchup_rounds = 0;
    h->hold_on = 0; h->sim_open = 0; h->sim_target = 0; h->sim_dt = 0.0; h->round_cap = 0; h->h_lag = nullptr;
    h->solve_cap = -1; h->sim_cap = 0;      // deferred solves inside the simulate loops: automatic (effective_cap)
    if (getenv("PGN_SOLVE_CAP")) h->solve_cap = atoi(getenv("PGN_SOLVE_CAP"));
    CK(cudaMemset(h->d_hold, 0, B)); CK(cudaMemset(h->d_kstep, 0, B * 4)); CK(cudaMemset(h->d_iters_acc, 0, B * 4));
    if (cudaMallocHost((void**)&h->h_lag, 16) != cudaSuccess) return bail(set_err(PGN_ENOMEM, "cudaMallocHost failed"));
    h->h_ring = nullptr; h->d_ring = nullptr; h->ring_head = h->ring_tail = h->ring_count = h->ring_created = 0;
    h->h_in = nullptr; h->in_pending = 0; h->d_hist = nullptr; h->hist_cap = h->hist_stride = h->hist_n = 0;
    h->comm = nullptr; h->comm_rank = 0; h->comm_size = 1; h->d_gath_c = nullptr; h->d_gath_i = nullptr;
    cudaEventCreateWithFlags(&h->ev_in, cudaEventDisableTiming);
    {
        cudaError_t e = cudaMallocHost((void**)&h->h_io, (1 + 19 * B) * sizeof(double));
        if (e == cudaSuccess) e = cudaMallocHost((void**)&h->h_in, 18 * B * sizeof(double));
        if (e != cudaSuccess) return bail(set_err(PGN_ENOMEM, "cudaMallocHost failed: %s", cudaGetErrorString(e)));
    }
    {   // no step yet: cache[x] = (Inf, 0)
        std::vector<double> hv(8 * B, 0.0);
        for (size_t v = 0; v < B; v++) hv[7 * B + v] = INFINITY;
        CK(cudaMemcpy(h->d_hji_val, hv.data(), 8 * B * 8, cudaMemcpyHostToDevice));
    }
    CK(cudaMemset(h->d_iters, 0, B * 4)); CK(cudaMemset(h->d_status, 0, B * 4)); CK(cudaMemset(h->d_rho_updates, 0, B * 4));
    CK(cudaMemset(h->d_rec, 0, B * (size_t)t.rec.rec_len * 8)); CK(cudaMemset(h->d_controls, 0, 3 * B * 8));
    {   // other car far away, time_offset = NaN (path mode), ts = 1..N (MPCTimeSteps ctor), rho = setting
        std::vector<double> tmp(4 * B, 0.0), nanv(B, NAN), rho(B, cfg->rho), ts(B * N);
        for (size_t v = 0; v < B; v++) for (size_t i = 0; i < N; i++) ts[v * N + i] = (double)(i + 1);
        std::vector<double> dt(B * T, 1.0);
        CK(cudaMemcpy(h->d_other, tmp.data(), 4 * B * 8, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(h->d_toff, nanv.data(), B * 8, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(h->d_rho, rho.data(), B * 8, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(h->d_ts, ts.data(), B * N * 8, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(h->d_prev_ts, ts.data(), B * N * 8, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(h->d_dt, dt.data(), B * T * 8, cudaMemcpyHostToDevice));
    }
    // placeholder_HJICache() (HJI_computation.jl:32-37): 2^7 zero grid on +-1000
    {
        int32_t dims[7]; float knots[14]; std::vector<float> V(128, 0.f), g(128 * 7, 0.f);
        for (int d = 0; d < 7; d++) { dims[d] = 2; knots[2 * d] = -1000.f; knots[2 * d + 1] = 1000.f; }
        if ((rc = set_hji_internal(h, dims, knots, V.data(), g.data()))) return bail(rc);
    }
    // straight_trajectory(30., 5.) as the default trajectory (Pigeon.jl:34-35)
    {
        const double f[12][2] = {{0, 6}, {0, 30}, {5, 5}, {0, 0}, {0, 0}, {0, 30}, {0, 0}, {0, 0}, {0, 0}, {0, 0}, {4, 4}, {-4, -4}};
        const double* fp[12];
        for (int k = 0; k < 12; k++) fp[k] = f[k];
        if ((rc = pgn_set_trajectories(h, 1, 2, fp))) return bail(rc);
    }
    if (h->admm_tmem && (rc = dev_alloc(h, &h->d_admm_scratch, admm_scratch_doubles(h)))) return bail(rc);
    int ce = admm_configure(h);
    if (ce != 0) return bail(set_err(PGN_ECUDA, "ADMM kernel needs %d bytes of shared memory per CTA: %s", h->admm_smem_bytes, cudaGetErrorString((cudaError_t)ce)));
    if ((rc = pgn_set_pipeline_parts(h, 0))) return bail(rc);     // automatic part count (1 for batches below ~1.5 waves of ADMM CTAs)
    CK(cudaDeviceSynchronize());
    *out = h;
    return PGN_OK;
}

int pgn_destroy(pgn_handle* h) {
    if (!h) return PGN_OK;
    DeviceGuard dev_guard__(h->device);
    cudaDeviceSynchronize();
    pgn_comm_destroy(h);
    for (void* p : h->allocs) cudaFree(p);
    if (h->cb_has_exec) { cudaGraphExecDestroy(h->cb_exec); cudaGraphDestroy(h->cb_graph); }
    for (int p = 0; p < PGN_MAX_PARTS; p++) drop_round_graph(h, p);
    if (h->h_io) cudaFreeHost(h->h_io);
    if (h->h_in) cudaFreeHost(h->h_in);
    if (h->h_lag) cudaFreeHost(h->h_lag);
    if (h->d_hji_ws) cudaFree(h->d_hji_ws);
    if (h->d_trace) cudaFree(h->d_trace);
    if (h->d_trace_n) cudaFree(h->d_trace_n);
    if (h->h_ring) cudaFreeHost(h->h_ring);
    if (h->ring_created)
        for (int sl = 0; sl < PGN_RING; sl++) { cudaEventDestroy(h->ring_h2d[sl]); for (int p = 0; p < PGN_MAX_PARTS; p++) cudaEventDestroy(h->ring_done[sl][p]); }
    cudaEventDestroy(h->ev_in);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    if (h->side_stream) cudaStreamDestroy(h->side_stream);
    if (h->parts_created) {
        for (int p = 0; p < PGN_MAX_PARTS; p++) {
            cudaStreamDestroy(h->part_stream[p]); cudaStreamDestroy(h->part_side[p]);
            cudaEventDestroy(h->part_done[p]); cudaEventDestroy(h->part_evf[p]); cudaEventDestroy(h->part_evj[p]);
        }
        cudaEventDestroy(h->part_begin);
    }
    cudaEventDestroy(h->ev_fork); cudaEventDestroy(h->ev_join);
    cudaEventDestroy(h->ev[0]); cudaEventDestroy(h->ev[1]);
    delete h;
    return PGN_OK;
}

int pgn_set_stream(pgn_handle* h, void* s) { ENTER(h, "NULL handle"); h->epoch++; h->stream = s ? (cudaStream_t)s : h->own_stream; return PGN_OK; }
int pgn_synchronize(pgn_handle* h) { ENTER(h, "NULL handle"); CK(cudaStreamSynchronize(h->stream)); return PGN_OK; }

int pgn_set_vehicle_params(pgn_handle* h, const double* vp) {
    ENTER(h, "NULL handle"); REQUIRE(vp, "NULL argument");
    CK(cudaStreamSynchronize(h->stream));
    h->epoch++;
    memcpy(&h->veh, vp, sizeof(double) * PGN_VEHICLE_PARAMS_LEN);
    return upload_constants(h);
}
int pgn_set_control_params(pgn_handle* h, const double* cp) {
    ENTER(h, "NULL handle"); REQUIRE(cp, "NULL argument");
    CtrlParams tmp;
    memcpy(&tmp, cp, sizeof(double) * PGN_CONTROL_PARAMS_LEN);
    REQUIRE(tmp.N_HJI >= 0 && tmp.N_HJI <= h->cfg.N_short, "N_HJI must be within [0, N_short]");      // nothing is committed on failure
    CK(cudaStreamSynchronize(h->stream));
    h->epoch++;
    h->ctl = tmp;
    return upload_constants(h);
}
int pgn_set_trajectories(pgn_handle* h, int32_t n_traj, int32_t n_nodes, const double* const fields[12]) {
    ENTER(h, "NULL handle"); REQUIRE(fields, "NULL argument");
    REQUIRE(n_traj >= 1 && n_nodes >= 2, "need n_traj >= 1 and n_nodes >= 2");
    const size_t cnt = (size_t)n_traj * n_nodes;
    double* base = nullptr;
    h->epoch++;
    CK(cudaStreamSynchronize(h->stream));
    if (h->have_traj && h->traj.f[0]) {      // latest_trajectory[] is replaced at run time (ros_integration.jl:19,53): release the previous tables
        void* old = (void*)h->traj.f[0];
        for (size_t i = 0; i < h->allocs.size(); i++) if (h->allocs[i] == old) { h->allocs.erase(h->allocs.begin() + i); cudaFree(old); break; }
        h->traj.f[0] = nullptr;
    }
    int rc = dev_alloc(h, &base, 12 * cnt);
    if (rc) return rc;
    for (int k = 0; k < 12; k++) {
        REQUIRE(fields[k], "NULL trajectory field");
        CK(cudaMemcpy(base + k * cnt, fields[k], cnt * sizeof(double), cudaMemcpyHostToDevice));
        h->traj.f[k] = base + k * cnt;
    }
    h->traj.n_traj = n_traj; h->traj.n_nodes = n_nodes;
    h->have_traj = true;
    CK(cudaMemset(h->d_last_seg, 0xff, (size_t)h->B * 4));
    CK(cudaMemset(h->d_traj_id, 0, (size_t)h->B * 4));
    return PGN_OK;
}
int pgn_assign_trajectories(pgn_handle* h, const int32_t* traj_id) {
    ENTER(h, "NULL handle"); REQUIRE(traj_id, "NULL argument");
    for (int v = 0; v < h->B; v++) REQUIRE(traj_id[v] >= 0 && traj_id[v] < h->traj.n_traj, "trajectory id out of range");
    CK(cudaStreamSynchronize(h->stream));        // the blocking copies below run on the legacy stream: order them after the handle's in-flight work
    CK(cudaMemcpy(h->d_traj_id, traj_id, (size_t)h->B * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(h->d_last_seg, 0xff, (size_t)h->B * 4));
    h->epoch++;
    return PGN_OK;
}
int pgn_set_hji_cache(pgn_handle* h, const int32_t dims[7], const float* knots, const float* V, const float* gradV) {
    ENTER(h, "NULL handle"); REQUIRE(dims && knots && V && gradV, "NULL argument");
    h->epoch++;
    return set_hji_internal(h, dims, knots, V, gradV);
}
// host arrays -> the packed pinned buffer; returns the unpack flags.  The pinned buffer is reused by every call: wait for the previous copy.
static int stage_inputs(pgn_handle* h, const double* q, const double* u, const double* other, const double* toff, const double* t0, int* flags_out) {
    if (h->in_pending) { CK(cudaEventSynchronize(h->ev_in)); h->in_pending = 0; }
    const size_t B = h->B;
    int flags = 0;
    if (q) { memcpy(h->h_in, q, 6 * B * 8); flags |= 1; }
    if (u) { memcpy(h->h_in + 6 * B, u, 3 * B * 8); flags |= 2; }
    if (other) { memcpy(h->h_in + 9 * B, other, 4 * B * 8); flags |= 4; }
    if (toff) { memcpy(h->h_in + 13 * B, toff, B * 8); flags |= 8; }
    if (t0) { memcpy(h->h_in + 14 * B, t0, B * 8); flags |= 16; }
    *flags_out = flags;
    if (!flags) return PGN_OK;
    // one copy of the span that holds the present fields
    const size_t lo = (flags & 1) ? 0 : (flags & 2) ? 6 * B : (flags & 4) ? 9 * B : (flags & 8) ? 13 * B : 14 * B;
    const size_t hi = (flags & 16) ? 15 * B : (flags & 8) ? 14 * B : (flags & 4) ? 13 * B : (flags & 2) ? 9 * B : 6 * B;
    CK(cudaMemcpyAsync(h->d_in + lo, h->h_in + lo, (hi - lo) * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaEventRecord(h->ev_in, h->stream));
    h->in_pending = 1;
    launch_unpack_state(h, flags);
    return PGN_OK;
}
int pgn_set_state(pgn_handle* h, const double* q, const double* u, const double* other, const double* toff) {
    ENTER(h, "NULL handle");
    h->sim_axis_valid = 0;
    int flags;
    int rc = stage_inputs(h, q, u, other, toff, nullptr, &flags);      // the caller's arrays are copied before the call returns; the rest is stream-ordered
    if (rc) return rc;
    CK(cudaGetLastError());
    return PGN_OK;
}
int pgn_reset_solved(pgn_handle* h, const uint8_t* mask) {
    ENTER(h, "NULL handle");
    if (!mask) { CK(cudaMemsetAsync(h->d_solved, 0, h->B, h->stream)); return PGN_OK; }
    CK(cudaMemcpyAsync(h->d_mask, mask, h->B, cudaMemcpyHostToDevice, h->stream));      // pageable source: staged before the call returns
    launch_masked_reset(h, h->d_mask, 1);
    CK(cudaGetLastError());
    return PGN_OK;
}
int pgn_set_guards(pgn_handle* h, int32_t nan_fallback, double pause_below_speed) {
    ENTER(h, "NULL handle");
    REQUIRE(pause_below_speed >= 0.0, "pause_below_speed must be >= 0");
    h->epoch++;
    h->guard_nan = nan_fallback != 0; h->guard_pause = pause_below_speed;
    if (pause_below_speed == 0.0) CK(cudaMemsetAsync(h->d_skip, 0, h->B, h->stream));
    return PGN_OK;
}
int pgn_reset_solver(pgn_handle* h, const uint8_t* mask) {
    ENTER(h, "NULL handle");
    if (mask) CK(cudaMemcpyAsync(h->d_mask, mask, h->B, cudaMemcpyHostToDevice, h->stream));
    launch_masked_reset(h, mask ? h->d_mask : nullptr, 2);       // one kernel on the handle's stream, ordered with the solves before and after it
    CK(cudaGetLastError());
    return PGN_OK;
}

// ---- the step API -------------------------------------------------------------------------------------------------------------
static int step_time_steps_dev(pgn_handle* h, const double* d_t0) { StageTimer T(h, 0); launch_time_steps(h, d_t0); return PGN_OK; }
static int step_nodes(pgn_handle* h) { StageTimer T(h, 0); launch_nodes(h); return PGN_OK; }
static int step_update(pgn_handle* h) {
    { StageTimer T(h, 1); launch_linearize(h); }
    if (h->cfg.kind == PGN_COUPLED) { StageTimer T(h, 2); launch_hji_constraint(h); }
    return PGN_OK;
}
static int step_solve(pgn_handle* h) { StageTimer T(h, 3); launch_admm(h); return PGN_OK; }
static int step_controls(pgn_handle* h, double* d_out) { StageTimer T(h, 4); launch_controls(h, d_out); return PGN_OK; }

int pgn_compute_time_steps(pgn_handle* h, const double* t0) {
    ENTER(h, "NULL handle"); REQUIRE(t0, "NULL argument");
    int flags;
    int rc = stage_inputs(h, nullptr, nullptr, nullptr, nullptr, t0, &flags);
    if (rc) return rc;
    step_time_steps_dev(h, h->d_t0);
    CK(cudaGetLastError());
    return PGN_OK;
}
int pgn_compute_linearization_nodes(pgn_handle* h) { ENTER(h, "NULL handle"); step_nodes(h); CK(cudaGetLastError()); return PGN_OK; }
int pgn_update_qp(pgn_handle* h) { ENTER(h, "NULL handle"); step_update(h); CK(cudaGetLastError()); return PGN_OK; }
int pgn_solve(pgn_handle* h) { ENTER(h, "NULL handle"); step_solve(h); CK(cudaGetLastError()); return PGN_OK; }
int pgn_get_next_control(pgn_handle* h, double* out) {
    ENTER(h, "NULL handle"); REQUIRE(out, "NULL argument");
    step_controls(h, h->d_controls);
    return download_aos(h, h->d_controls, out, 3);
}
int pgn_step_device(pgn_handle* h, const double* d_t0, double* d_out) {
    ENTER(h, "NULL handle"); REQUIRE(d_t0, "NULL argument");
    int rc = for_each_part(h, [&]() {
        step_time_steps_dev(h, d_t0);
        step_nodes(h);
        step_update(h);
        step_solve(h);
        step_controls(h, h->d_controls);
        return (int)PGN_OK;
    });
    if (rc) return rc;
    if (d_out) CK(cudaMemcpyAsync(d_out, h->d_controls, (size_t)h->B * 3 * 8, cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaGetLastError());
    return PGN_OK;
}
int pgn_step(pgn_handle* h, const double* t0, double* out) {
    ENTER(h, "NULL handle"); REQUIRE(t0, "NULL argument");
    int flags;
    int rc = stage_inputs(h, nullptr, nullptr, nullptr, nullptr, t0, &flags);
    if (rc) return rc;
    rc = pgn_step_device(h, h->d_t0, nullptr);
    if (rc) return rc;
    if (out) {      // [3][B] -> [B][3] on the device, one D2H copy into the pinned tail, one host copy into the caller's array
        const size_t B = h->B;
        launch_transpose_out(h, h->d_controls, h->d_in + 15 * B, 3);
        CK(cudaMemcpyAsync(h->h_in + 15 * B, h->d_in + 15 * B, 3 * B * 8, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        memcpy(out, h->h_in + 15 * B, 3 * B * 8);
        return PGN_OK;
    }
    CK(cudaStreamSynchronize(h->stream));
    return PGN_OK;
}
// ---- pipelined host-buffer stepping ---------------------------------------------------------------------------------------------------------
// The callback feeds MEASURED states (src/ros_integration.jl:50-53): step k+1's inputs do not depend on step k's output, so a host that serves
// many vehicles can keep several steps in flight.  pgn_step_submit copies the inputs into the slot's pinned block and enqueues, per pipeline
// part and on the part's own stream, [unpack -> the five stages -> pack -> D2H of the part's controls]; nothing joins the parts, so — as in the
// simulate loop — the per-vehicle stages of one part run while the ADMM kernel of another drains.  pgn_step_collect waits for the OLDEST
// submitted step and hands its controls over.  Per vehicle the sequence of operations is exactly that of pgn_set_state + pgn_step.
static int ring_create(pgn_handle* h) {
    if (h->ring_created) return PGN_OK;
    const size_t B = h->B;
    cudaError_t e = cudaMallocHost((void**)&h->h_ring, (size_t)PGN_RING * 18 * B * sizeof(double));
    if (e != cudaSuccess) return set_err(PGN_ENOMEM, "cudaMallocHost failed: %s", cudaGetErrorString(e));
    int rc = dev_alloc(h, &h->d_ring, (size_t)PGN_RING * 18 * B);
    if (rc) return rc;
    for (int sl = 0; sl < PGN_RING; sl++) {
        CK(cudaEventCreateWithFlags(&h->ring_h2d[sl], cudaEventDisableTiming));
        for (int p = 0; p < PGN_MAX_PARTS; p++) CK(cudaEventCreateWithFlags(&h->ring_done[sl][p], cudaEventDisableTiming));
    }
    h->ring_created = 1;
    return PGN_OK;
}
int pgn_step_submit(pgn_handle* h, const double* q, const double* u, const double* other, const double* t0) {
    ENTER_NODRAIN(h, "NULL handle"); REQUIRE(t0, "t0 is NULL");
    REQUIRE(h->ring_count < PGN_RING, "too many steps in flight: call pgn_step_collect first");
    int rc = ring_create(h);
    if (rc) return rc;
    const size_t B = h->B;
    const int sl = h->ring_head;
    double* hin = h->h_ring + (size_t)sl * 18 * B;
    double* din = h->d_ring + (size_t)sl * 18 * B;
    int flags = 16;
    if (q) { memcpy(hin, q, 6 * B * 8); flags |= 1; }
    if (u) { memcpy(hin + 6 * B, u, 3 * B * 8); flags |= 2; }
    if (other) { memcpy(hin + 9 * B, other, 4 * B * 8); flags |= 4; }
    memcpy(hin + 14 * B, t0, B * 8);      // slot layout = k_unpack_state's: q [0, 6B) | u | other | (time_offset, not sent) | t0 [14B, 15B) | out [15B, 18B)
    cudaStream_t s0 = h->stream;
    const size_t lo = (flags & 1) ? 0 : (flags & 2) ? 6 * B : (flags & 4) ? 9 * B : 14 * B;
    CK(cudaMemcpyAsync(din + lo, hin + lo, (15 * B - lo) * 8, cudaMemcpyHostToDevice, s0));      // ONE copy of the span that holds the present fields
    CK(cudaEventRecord(h->ring_h2d[sl], s0));
    const int P = (h->parts <= 1 || serial_profiling(h)) ? 1 : h->parts;
    cudaStream_t side0 = h->side_stream;
    cudaEvent_t f0 = h->ev_fork, j0 = h->ev_join;
    for (int p = 0; p < P; p++) {
        if (P > 1) {
            CK(cudaStreamWaitEvent(h->part_stream[p], h->ring_h2d[sl], 0));
            h->stream = h->part_stream[p]; h->side_stream = h->part_side[p]; h->ev_fork = h->part_evf[p]; h->ev_join = h->part_evj[p];
            h->part = p; h->v0 = (int)((long long)h->B * p / P); h->nv = (int)((long long)h->B * (p + 1) / P) - h->v0;
        }
        launch_unpack_range(h, din, flags & ~8);
        step_time_steps_dev(h, h->d_t0);
        step_nodes(h);
        step_update(h);
        step_solve(h);
        step_controls(h, h->d_controls);
        launch_pack_out(h, h->d_controls, din + 15 * B, 3);            // [3][B] -> rows v0 .. v0+nv of [B][3] in the slot's output block
        CK(cudaMemcpyAsync(hin + 15 * B + (size_t)h->v0 * 3, din + 15 * B + (size_t)h->v0 * 3, (size_t)h->nv * 3 * 8, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaEventRecord(h->ring_done[sl][p], h->stream));
    }
    h->stream = s0; h->side_stream = side0; h->ev_fork = f0; h->ev_join = j0;
    h->part = 0; h->v0 = 0; h->nv = h->B;
    h->ring_parts[sl] = P; h->ring_flags[sl] = flags;
    h->ring_head = (sl + 1) % PGN_RING; h->ring_count++;
    CK(cudaGetLastError());
    return PGN_OK;
}
int pgn_step_collect(pgn_handle* h, double* out) {
    ENTER_NODRAIN(h, "NULL handle");
    REQUIRE(h->ring_count > 0, "no step in flight: call pgn_step_submit first");
    const size_t B = h->B;
    const int sl = h->ring_tail;
    for (int p = 0; p < h->ring_parts[sl]; p++) CK(cudaEventSynchronize(h->ring_done[sl][p]));
    if (out) memcpy(out, h->h_ring + (size_t)sl * 18 * B + 15 * B, 3 * B * 8);
    h->ring_tail = (sl + 1) % PGN_RING; h->ring_count--;
    if (h->ring_count == 0 && h->ring_parts[sl] > 1) {      // the caller's stream continues after everything the parts did
        for (int p = 0; p < h->ring_parts[sl]; p++) CK(cudaStreamWaitEvent(h->stream, h->ring_done[sl][p], 0));
    }
    return PGN_OK;
}
int pgn_steps_in_flight(pgn_handle* h, int32_t* n) { ENTER_NODRAIN(h, "NULL handle"); REQUIRE(n, "NULL argument"); *n = h->ring_count; return PGN_OK; }

// from_autobox_callback (ros_integration.jl:48-151): the whole callback for B vehicles as one packed H2D copy, one graph launch
// (unpack + time selection + the five step stages + pack) and one D2H copy.
static int callback_enqueue(pgn_handle* h) {
    const size_t B = h->B;
    CK(cudaMemcpyAsync(h->d_io, h->h_io, (1 + 14 * B) * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    launch_callback_in(h);
    h->in_callback = 1;
    step_time_steps_dev(h, h->d_t0);
    step_nodes(h);
    step_update(h);
    step_solve(h);
    step_controls(h, h->d_controls);
    h->in_callback = 0;
    launch_callback_out(h);
    CK(cudaMemcpyAsync(h->h_io + 1 + 14 * B, h->d_io + 1 + 14 * B, 5 * B * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    return PGN_OK;
}
int pgn_from_autobox(pgn_handle* h, const double* q, const double* u, const double* other, const double* stamp, double* out) {
    ENTER(h, "NULL handle"); REQUIRE(q && u && stamp && out, "NULL argument");
    REQUIRE(h->have_traj, "no trajectory set");
    const size_t B = h->B;
    double* io = h->h_io;
    io[0] = other ? 1.0 : 0.0;
    memcpy(io + 1, q, 6 * B * 8); memcpy(io + 1 + 6 * B, u, 3 * B * 8);
    if (other) memcpy(io + 1 + 9 * B, other, 4 * B * 8);
    memcpy(io + 1 + 13 * B, stamp, B * 8);
    int rc = PGN_OK;
    if (serial_profiling(h)) {
        rc = callback_enqueue(h);          // stage timers synchronise: no capture
    } else {
        if (!h->cb_has_exec || h->cb_epoch != h->epoch || h->cb_stream != h->stream) {
            if (h->cb_has_exec) { cudaGraphExecDestroy(h->cb_exec); cudaGraphDestroy(h->cb_graph); h->cb_has_exec = 0; }
            const long long l0 = h->launches;
            CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
            rc = callback_enqueue(h);
            cudaGraph_t g = nullptr;
            cudaError_t e = cudaStreamEndCapture(h->stream, &g);
            h->in_callback = 0;
            h->cb_launches = h->launches - l0; h->launches = l0;
            if (rc) { if (g) cudaGraphDestroy(g); return rc; }
            if (e != cudaSuccess) return set_err(PGN_ECUDA, "graph capture failed: %s", cudaGetErrorString(e));
            e = cudaGraphInstantiate(&h->cb_exec, g, 0);
            if (e != cudaSuccess) { cudaGraphDestroy(g); return set_err(PGN_ECUDA, "graph instantiation failed: %s", cudaGetErrorString(e)); }
            h->cb_graph = g; h->cb_has_exec = 1; h->cb_epoch = h->epoch; h->cb_stream = h->stream;
        }
        CK(cudaGraphLaunch(h->cb_exec, h->stream));
        h->launches += h->cb_launches;
    }
    if (rc) return rc;
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaGetLastError());
    memcpy(out, io + 1 + 14 * B, 5 * B * 8);
    return PGN_OK;
}
// one closed-loop step of the current vehicle range on the current stream: the step stages, the plant step on the side stream.
// rec_slot >= 0: the history recorder keeps (state, control, node 1, params 1) of this step (pgn_set_history).
// phase 0: the whole round.  Split rounds (simulate loop, graphs): phase 1 = everything before the QP solve, with the plant propagation forked at the
// start and joined at the end (it then runs beside the nodes / linearisation stages instead of beside the solve); phase 2 = everything after it
static int step_rollout_body(pgn_handle* h, const double* d_t0, double dt, int rec_slot, int phase = 0) {
    if (phase == 2) {
        launch_stamp(h, 4);
        step_controls(h, h->d_controls);
        launch_stamp(h, 5);
        launch_commit_rollout(h);
        launch_stamp(h, 6);
        return PGN_OK;
    }
    launch_stamp(h, 0);
    if (phase == 1) {    // fork: the propagation reads state / current control only (nothing on the main stream writes them before the commit)
        CK(cudaEventRecord(h->ev_fork, h->stream));
        CK(cudaStreamWaitEvent(h->side_stream, h->ev_fork, 0));
        launch_propagate_shadow(h, dt, h->side_stream);
        CK(cudaEventRecord(h->ev_join, h->side_stream));
    }
    step_time_steps_dev(h, d_t0);
    launch_stamp(h, 1);
    step_nodes(h);
    if (rec_slot >= 0) launch_record(h, rec_slot);
    launch_stamp(h, 2);
    step_update(h);
    launch_stamp(h, 3);
    if (phase == 1) { CK(cudaStreamWaitEvent(h->stream, h->ev_join, 0)); return PGN_OK; }
    if (serial_profiling(h) || !h->side_stream) {            // stage timers synchronise: serial order
        step_solve(h);
        step_controls(h, h->d_controls);
        { StageTimer T(h, 5); launch_rollout(h, dt); }
        return PGN_OK;
    }
    // fork: the propagation reads state / current control only (nothing on the main stream writes them before the commit)
    CK(cudaEventRecord(h->ev_fork, h->stream));
    CK(cudaStreamWaitEvent(h->side_stream, h->ev_fork, 0));
    launch_propagate_shadow(h, dt, h->side_stream);
    CK(cudaEventRecord(h->ev_join, h->side_stream));
    step_solve(h);
    launch_stamp(h, 4);
    step_controls(h, h->d_controls);
    launch_stamp(h, 5);
    CK(cudaStreamWaitEvent(h->stream, h->ev_join, 0));
    launch_commit_rollout(h);
    launch_stamp(h, 6);
    return PGN_OK;
}
// step + plant rollout of `simulate` (model_predictive_control.jl:87-98) with the plant step beside the QP solve
int pgn_step_rollout_device(pgn_handle* h, const double* d_t0, double* d_out, double dt) {
    ENTER(h, "NULL handle"); REQUIRE(d_t0, "NULL argument");
    int rc = for_each_part(h, [&]() { return step_rollout_body(h, d_t0, dt, -1); });
    if (rc) return rc;
    if (d_out) CK(cudaMemcpyAsync(d_out, h->d_controls, (size_t)h->B * 3 * 8, cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaGetLastError());
    return PGN_OK;
}
int pgn_rollout(pgn_handle* h, double dt) {
    ENTER(h, "NULL handle");
    { StageTimer T(h, 5); launch_rollout(h, dt); }
    CK(cudaGetLastError());
    return PGN_OK;
}
// the `simulate` loop on the device: every pipeline part runs ALL its steps on its own stream (a vehicle's step k+1 depends only on its own
// step k), so the parts drift apart and the small per-vehicle kernels of one part fill the SMs the ADMM kernel of another leaves idle
// effective cap of this handle: the explicit setting, or (automatic, solve_cap < 0) 400 iterations when a range's ADMM launch is at most two
// waves of CTAs — there one long solve IS the launch time — and none for large ranges, whose launches absorb a few-hundred-iteration straggler
// in their many waves while a deferred one would surface in the catch-up rounds (measured: tools/gpu_cap_sweep.sh).  400 and not
// 200: a vehicle whose QP needs 200-400 iterations at EVERY step would fall one round behind per step (tools/gpu_rank2_cap.sh: the batch of rank 2
// runs 442 k steps/s with 200, 472 k with 400, 480 k uncapped; the cold start of rank 0's batch 188 k uncapped, 217 k with 200, ~208 k with 400)
static int effective_cap(const pgn_handle* h) {
    if (h->solve_cap >= 0) return h->solve_cap;
    const int c = h->st.check_termination, a = (h->st.adaptive_rho && h->st.adaptive_rho_interval > 0) ? h->st.adaptive_rho_interval : 1;
    if (!(c > 0 && 400 % c == 0 && 400 % a == 0)) return 0;
    const int per_range = h->B / (h->parts > 0 ? h->parts : 1), resident = h->num_sms * h->admm_ctas_per_sm;
    return per_range <= 2 * resident ? 400 : 0;
}
// One round of one pipeline part as a CUDA graph.  Nothing in a round depends on the host: the step time comes from the per-vehicle step
// counters, the target step count sits in device memory, the recorder files by the counters — so the graph is captured once per part (and
// again only when a setter bumps the epoch, or dt / cap / recorder change) and every further round is ONE graph launch instead of 13 kernel
// launches and 4 event operations.  The loop was host-bound from 4 parts up (tools/gpu_parts_sweep.sh: 5 parts 401 k steps/s against 514 k with 4).
static void drop_round_graph(pgn_handle* h, int p) {
    if (h->rg_exec[p]) { cudaGraphExecDestroy(h->rg_exec[p]); cudaGraphDestroy(h->rg_graph[p]); h->rg_exec[p] = nullptr; }
    if (h->rg_exec2[p]) { cudaGraphExecDestroy(h->rg_exec2[p]); cudaGraphDestroy(h->rg_graph2[p]); h->rg_exec2[p] = nullptr; }
}
// kernel nodes do not inherit the priority of the capturing stream: every per-vehicle kernel gets the highest priority, the ADMM kernel the lowest
static void set_node_priorities(pgn_handle* h, cudaGraph_t g) {
    size_t nn = 0;
    if (cudaGraphGetNodes(g, nullptr, &nn) != cudaSuccess || !nn) { cudaGetLastError(); return; }
    std::vector<cudaGraphNode_t> nodes(nn);
    cudaGraphGetNodes(g, nodes.data(), &nn);
    for (size_t i = 0; i < nn; i++) {
        cudaGraphNodeType ty;
        if (cudaGraphNodeGetType(nodes[i], &ty) != cudaSuccess || ty != cudaGraphNodeTypeKernel) continue;
        cudaKernelNodeParams kp;
        if (cudaGraphKernelNodeGetParams(nodes[i], &kp) != cudaSuccess) { cudaGetLastError(); continue; }
        cudaLaunchAttributeValue v;
        memset(&v, 0, sizeof(v));
        v.priority = pgn::admm_is_kernel(kp.func) ? h->prio_least : h->prio_greatest;
        if (cudaGraphKernelNodeSetAttribute(nodes[i], cudaLaunchAttributePriority, &v) != cudaSuccess) cudaGetLastError();
    }
}
// Split rounds (PGN_SPLIT_ROUNDS=1): the graph of a round is cut in two around the QP solve and the ADMM kernel is launched between them as an
// ordinary low-priority launch.  Measured (tools/gpu_admm_trace.py): inside ONE graph the node priorities are set and read back correctly but do
// not change the order in which the block scheduler serves the kernels — `controls` (3.5 us of work) waits 46 us on average and 313 us at the 90th
// percentile behind the pending ADMM CTAs of the other parts; with split rounds 16 / 50 us (time steps 79 -> 42, nodes 108 -> 89, commit + next round
// 30 -> 21) — and the loop is exactly as fast as before (548 k against 547 k steps/s): what the short kernels gain, the linearisation (424 -> 453 us)
// and the ADMM launches (1166 -> 1210 us) lose, because the step is bound by the SM time of those two kernels, not by the order they are served in.
static int capture_round(pgn_handle* h, double dt, int rec, int phase, cudaGraph_t* g_out, cudaGraphExec_t* x_out, long long* n_launches) {
    const long long l0 = h->launches;
    CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
    if (phase != 2) launch_round_begin(h, dt);
    int rc = step_rollout_body(h, h->d_t0, dt, rec ? 0 : -1, phase);
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamEndCapture(h->stream, &g);
    *n_launches += h->launches - l0; h->launches = l0;
    if (rc) { if (g) cudaGraphDestroy(g); return rc; }
    if (e != cudaSuccess) return set_err(PGN_ECUDA, "graph capture of a simulate round failed: %s", cudaGetErrorString(e));
    if (h->admm_low_priority) set_node_priorities(h, g);
    e = cudaGraphInstantiate(x_out, g, 0);
    if (e != cudaSuccess) { cudaGraphDestroy(g); *x_out = nullptr; return set_err(PGN_ECUDA, "graph instantiation failed: %s", cudaGetErrorString(e)); }
    *g_out = g;
    return PGN_OK;
}
static int ensure_round_graph(pgn_handle* h, int p, double dt, int cap, int rec) {      // called with the part's streams / range swapped into the handle
    if (h->rg_exec[p] && h->rg_epoch[p] == h->epoch && h->rg_dt[p] == dt && h->rg_cap[p] == cap && h->rg_rec[p] == rec && h->rg_v0[p] == h->v0 && h->rg_nv[p] == h->nv &&
        (h->rg_exec2[p] != nullptr) == (h->split_rounds != 0)) return PGN_OK;
    drop_round_graph(h, p);
    h->rg_launches[p] = 0;
    int rc = capture_round(h, dt, rec, h->split_rounds ? 1 : 0, &h->rg_graph[p], &h->rg_exec[p], &h->rg_launches[p]);
    if (rc) return rc;
    if (h->split_rounds && (rc = capture_round(h, dt, rec, 2, &h->rg_graph2[p], &h->rg_exec2[p], &h->rg_launches[p]))) { drop_round_graph(h, p); return rc; }
    h->rg_epoch[p] = h->epoch; h->rg_dt[p] = dt; h->rg_cap[p] = cap; h->rg_rec[p] = rec; h->rg_v0[p] = h->v0; h->rg_nv[p] = h->nv;
    return PGN_OK;
}
// The simulate loop in ROUNDS: a round = one step attempt of every vehicle of the range that is not held.  Every vehicle counts its own steps
// (d_kstep); with deferred solves (cap > 0) a QP that uses up its share of ADMM iterations keeps its vehicle on hold and continues in the next
// round's launch, so a 4000-iteration straggler costs its own vehicle a few rounds instead of costing every vehicle of the range the whole
// solve; the vehicles that fell behind are finished by finish_sim at the next point that needs results.  n_steps rounds are enqueued here.
static int simulate_enqueue(pgn_handle* h, double dt, int k0, int n_steps) {
    const int cap = effective_cap(h);
    if (k0 == 0) {
        CK(cudaMemsetAsync(h->d_kstep, 0, (size_t)h->B * 4, h->stream));
        CK(cudaMemsetAsync(h->d_hold, 0, h->B, h->stream));
    } else if (!(h->sim_axis_valid && k0 == h->sim_target)) {       // not the continuation of the previous call: every vehicle stands at step k0
        launch_fill_i32(h, h->d_kstep, k0, h->B);
        CK(cudaMemsetAsync(h->d_hold, 0, h->B, h->stream));
    }
    h->sim_target = k0 + n_steps; h->sim_dt = dt; h->sim_open = cap > 0; h->sim_axis_valid = 1;
    launch_fill_i32(h, h->d_lag + 1, h->sim_target, 1);           // the target step count the rounds read
    const int rec = h->hist_stride > 0;
    if (rec) { const int nrec = std::min(h->hist_cap, (h->sim_target + h->hist_stride - 1) / h->hist_stride); if (nrec > h->hist_n) h->hist_n = nrec; }
    h->hold_on = 1; h->round_cap = cap; h->sim_cap = cap;
    const bool graphs = h->parts > 1 && !serial_profiling(h) && !getenv("PGN_NO_GRAPHS");
    if (graphs) {
        // every part's graph is made ready BEFORE anything is launched: destroying / instantiating a graph synchronises with the device, and
        // done between the parts' launches it serialised the parts (measured: 101 cold steps 227 ms -> 553 ms after any re-capture)
        cudaStream_t s0 = h->stream, side0 = h->side_stream;
        cudaEvent_t f0 = h->ev_fork, j0 = h->ev_join;
        int rc0 = PGN_OK;
        for (int p = 0; p < h->parts && rc0 == PGN_OK; p++) {
            h->stream = h->part_stream[p]; h->side_stream = h->part_side[p]; h->ev_fork = h->part_evf[p]; h->ev_join = h->part_evj[p];
            h->part = p; h->v0 = (int)((long long)h->B * p / h->parts); h->nv = (int)((long long)h->B * (p + 1) / h->parts) - h->v0;
            rc0 = ensure_round_graph(h, p, dt, cap, rec);
        }
        h->stream = s0; h->side_stream = side0; h->ev_fork = f0; h->ev_join = j0;
        h->part = 0; h->v0 = 0; h->nv = h->B;
        if (rc0) { h->hold_on = 0; return rc0; }
    }
    int rc = for_each_part(h, [&]() {
        if (graphs) {
            if (h->rg_exec2[h->part]) {
                for (int k = 0; k < n_steps; k++) {
                    CK(cudaGraphLaunch(h->rg_exec[h->part], h->stream));
                    step_solve(h);                                   // counts its own launches
                    CK(cudaGraphLaunch(h->rg_exec2[h->part], h->stream));
                }
            } else {
                for (int k = 0; k < n_steps; k++) CK(cudaGraphLaunch(h->rg_exec[h->part], h->stream));
            }
            h->launches += (long long)n_steps * h->rg_launches[h->part];
            return (int)PGN_OK;
        }
        for (int k = 0; k < n_steps; k++) {
            launch_round_begin(h, dt);
            int rc2 = step_rollout_body(h, h->d_t0, dt, rec ? 0 : -1);
            if (rc2) return rc2;
        }
        return (int)PGN_OK;
    });
    h->hold_on = 0;
    return rc;
}
// catch-up rounds of a simulate loop with deferred solves: until every vehicle has completed its steps.  The remaining solves get a larger
// share of iterations per launch (nothing else is waiting for the SMs any more).
static int finish_sim(pgn_handle* h) {
    if (!h->sim_open) return PGN_OK;
    h->sim_open = 0;
    for (int round = 0;; round++) {
        launch_count_lag(h, h->sim_target);
        CK(cudaMemcpyAsync(h->h_lag, h->d_lag, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        if (*h->h_lag == 0) break;
        h->catchup_rounds++;
        if (round > 100000) return set_err(PGN_ECUDA, "simulate: %d vehicles do not reach step %d", *h->h_lag, h->sim_target);
        h->hold_on = 1; h->round_cap = 5 * h->sim_cap;
        int rc = for_each_part(h, [&]() {
            launch_round_begin(h, h->sim_dt);
            return step_rollout_body(h, h->d_t0, h->sim_dt, h->hist_stride > 0 ? 0 : -1);
        });
        h->hold_on = 0;
        if (rc) return rc;
    }
    CK(cudaGetLastError());
    return PGN_OK;
}
// automatic part count: 4 parts from 64 vehicles up.  Measured (simulate loop, coupled N = 31, ms per step with 1 / 2 / 4 / 8 parts): B = 64: 0.75 / 0.69 /
// 0.63 / 1.10; 128: 0.78 / 0.74 / 0.68 / 1.20; 256: 1.12 / 0.91 / 0.86 / 1.34; 512: 1.74 / 1.46 / 1.29 / 1.63; 1024: 2.87 / 2.67 / 2.49 / 2.68 (3: 2.56, 5: 3.28,
// 6: 2.98, 7: 2.79); 2048: 5.40 / 4.97 / 4.91; 4096: 10.31 / 9.74 / 9.69.  Per-step calls (parts joined every call) are never slower with 4 parts.
static int auto_parts(pgn_handle* h) { return h->B >= 64 ? 4 : 1; }
int pgn_simulate(pgn_handle* h, const double* t0, double dt, int32_t n_steps) {
    ENTER(h, "NULL handle"); REQUIRE(t0 && n_steps >= 0, "bad argument");
    CK(cudaMemcpyAsync(h->d_t0_base, t0, (size_t)h->B * 8, cudaMemcpyHostToDevice, h->stream));
    h->hist_n = 0;
    int rc = simulate_enqueue(h, dt, 0, n_steps);
    if (rc) return rc;
    if ((rc = finish_sim(h))) return rc;
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaGetLastError());
    return PGN_OK;
}
int pgn_simulate_device(pgn_handle* h, const double* d_t0, double dt, int32_t k0, int32_t n_steps) {
    ENTER_NODRAIN(h, "NULL handle"); REQUIRE(d_t0 && n_steps >= 0 && k0 >= 0, "bad argument");
    if (h->ring_count) drain_ring(h);
    // a call that continues the time axis of the previous one (k0 = its end) leaves the vehicles that are behind where they are: they go on here
    if (h->sim_open && !(k0 == h->sim_target && dt == h->sim_dt && effective_cap(h) > 0)) { int rc0 = finish_sim(h); if (rc0) return rc0; }
    CK(cudaMemcpyAsync(h->d_t0_base, d_t0, (size_t)h->B * 8, cudaMemcpyDeviceToDevice, h->stream));
    if (k0 == 0) h->hist_n = 0;
    int rc = simulate_enqueue(h, dt, k0, n_steps);
    if (rc) return rc;
    CK(cudaGetLastError());
    return PGN_OK;
}
int pgn_set_pipeline_parts(pgn_handle* h, int32_t parts) {
    ENTER(h, "NULL handle");
    REQUIRE(parts >= 0 && parts <= PGN_MAX_PARTS, "parts must be 0 (automatic) or 1..8");
    if (parts == 0) parts = auto_parts(h);
    if (parts > h->B) parts = h->B;
    if (parts > 1 && !h->parts_created) {
        CK(cudaEventCreateWithFlags(&h->part_begin, cudaEventDisableTiming));
        for (int p = 0; p < PGN_MAX_PARTS; p++) {
            const int prio = h->admm_low_priority ? h->prio_greatest : h->prio_least;
            CK(cudaStreamCreateWithPriority(&h->part_stream[p], cudaStreamNonBlocking, prio)); CK(cudaStreamCreateWithPriority(&h->part_side[p], cudaStreamNonBlocking, prio));
            CK(cudaEventCreateWithFlags(&h->part_done[p], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&h->part_evf[p], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&h->part_evj[p], cudaEventDisableTiming));
        }
        h->parts_created = 1;
    }
    h->epoch++;
    h->parts = parts;
    return PGN_OK;
}
// simulate's return values (model_predictive_control.jl:84-99): qs / us before every recorded step, xs = mpc.qs[1], ps = mpc.ps[1]
int pgn_set_history(pgn_handle* h, int32_t capacity, int32_t stride) {
    ENTER(h, "NULL handle");
    REQUIRE(capacity >= 0 && stride >= 0 && (capacity == 0) == (stride == 0), "capacity and stride must both be positive, or both 0 (off)");
    CK(cudaStreamSynchronize(h->stream));
    if (h->d_hist) {
        for (size_t i = 0; i < h->allocs.size(); i++) if (h->allocs[i] == (void*)h->d_hist) { h->allocs.erase(h->allocs.begin() + i); cudaFree(h->d_hist); break; }
        h->d_hist = nullptr;
    }
    h->hist_cap = h->hist_stride = h->hist_n = 0;
    h->epoch++;                                     // the recorder's arguments are baked into the round graphs
    if (capacity == 0) return PGN_OK;
    int rc = dev_alloc(h, &h->d_hist, (size_t)capacity * (13 + h->nx) * h->B);
    if (rc) return rc;
    h->hist_cap = capacity; h->hist_stride = stride;
    return PGN_OK;
}
int pgn_get_history(pgn_handle* h, int32_t* n_records, double* qs, double* us, double* xs, double* ps) {
    ENTER(h, "NULL handle"); REQUIRE(n_records, "NULL argument");
    *n_records = h->hist_n;
    if (h->hist_n == 0 || (!qs && !us && !xs && !ps)) return PGN_OK;
    const size_t B = h->B, F = 13 + h->nx, nx = h->nx;
    std::vector<double> rec((size_t)h->hist_n * F * B);
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpy(rec.data(), h->d_hist, rec.size() * 8, cudaMemcpyDeviceToHost));
    for (size_t r = 0; r < (size_t)h->hist_n; r++) {
        const double* s = rec.data() + r * F * B;
        for (size_t v = 0; v < B; v++) {
            if (qs) for (size_t f = 0; f < 6; f++) qs[(r * B + v) * 6 + f] = s[f * B + v];
            if (us) for (size_t f = 0; f < 3; f++) us[(r * B + v) * 3 + f] = s[(6 + f) * B + v];
            if (xs) for (size_t f = 0; f < nx; f++) xs[(r * B + v) * nx + f] = s[(9 + f) * B + v];
            if (ps) for (size_t f = 0; f < 4; f++) ps[(r * B + v) * 4 + f] = s[(9 + nx + f) * B + v];
        }
    }
    return PGN_OK;
}
int pgn_set_solve_cap(pgn_handle* h, int32_t iters) {
    ENTER(h, "NULL handle");
    REQUIRE(iters >= -1, "iters must be >= 0, or -1 (automatic)");
    if (iters > 0) {
        const int c = h->st.check_termination, a = (h->st.adaptive_rho && h->st.adaptive_rho_interval > 0) ? h->st.adaptive_rho_interval : 1;
        REQUIRE(c > 0 && iters % c == 0 && iters % a == 0, "the cap must be a multiple of check_termination and of adaptive_rho_interval");
    }
    h->epoch++;
    h->solve_cap = iters;
    return PGN_OK;
}
int pgn_get_pipeline_parts(pgn_handle* h, int32_t* parts) { ENTER(h, "NULL handle"); REQUIRE(parts, "NULL argument"); *parts = h->parts; return PGN_OK; }

// ---- introspection -----------------------------------------------------------------------------------------------------------
int pgn_qp_dims(pgn_handle* h, int32_t* o) {
    ENTER(h, "NULL handle"); REQUIRE(o, "NULL argument");
    o[0] = h->N; o[1] = h->nx; o[2] = h->nu; o[3] = h->tab.n; o[4] = h->tab.m; o[5] = h->tab.nnzA; o[6] = h->tab.nnzL; o[7] = h->tab.nlev;
    o[8] = h->tab.nslots; o[9] = h->tab.n_fwd_ph + h->tab.n_bwd_ph; o[10] = (int)h->tab.fac_ent.size(); o[11] = (int)h->tab.inv_ent.size(); o[12] = h->tab.tail_dim;
    o[13] = (int)h->tab.bent.size(); o[14] = h->admm_smem_bytes; o[15] = h->admm_threads | (h->admm_tmem << 16) | (h->admm_ctas_per_sm << 20);      // threads | tensor-memory variant << 16 | resident CTAs per SM << 20
    return PGN_OK;
}
int pgn_get_state(pgn_handle* h, double* q, double* u) {
    ENTER(h, "NULL handle");
    int rc;
    if (q && (rc = download_aos(h, h->d_state, q, 6))) return rc;
    if (u && (rc = download_aos(h, h->d_control, u, 3))) return rc;
    return PGN_OK;
}
#define D2H(dst, src, count)                                                                             \
    do {                                                                                                 \
        if (dst) CK(cudaMemcpyAsync(dst, src, (size_t)(count) * sizeof(*(dst)), cudaMemcpyDeviceToHost, h->stream)); \
    } while (0)
int pgn_get_time_steps(pgn_handle* h, double* ts, double* dt, double* prev_ts) {
    ENTER(h, "NULL handle");
    D2H(ts, h->d_ts, (size_t)h->B * h->N); D2H(dt, h->d_dt, (size_t)h->B * h->T); D2H(prev_ts, h->d_prev_ts, (size_t)h->B * h->N);
    CK(cudaStreamSynchronize(h->stream));
    return PGN_OK;
}
int pgn_get_nodes(pgn_handle* h, double* qs, double* us, double* ps) {
    ENTER(h, "NULL handle");
    D2H(qs, h->d_qs, (size_t)h->B * h->N * h->nx); D2H(us, h->d_us, (size_t)h->B * h->N * 2); D2H(ps, h->d_ps, (size_t)h->B * h->N * 4);
    CK(cudaStreamSynchronize(h->stream));
    return PGN_OK;
}
int pgn_set_nodes(pgn_handle* h, const double* qs, const double* us, const double* ps) {
    ENTER(h, "NULL handle");
    if (qs) CK(cudaMemcpyAsync(h->d_qs, qs, (size_t)h->B * h->N * h->nx * 8, cudaMemcpyHostToDevice, h->stream));
    if (us) CK(cudaMemcpyAsync(h->d_us, us, (size_t)h->B * h->N * 2 * 8, cudaMemcpyHostToDevice, h->stream));
    if (ps) CK(cudaMemcpyAsync(h->d_ps, ps, (size_t)h->B * h->N * 4 * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return PGN_OK;
}
int pgn_get_qp_data(pgn_handle* h, double* A, double* B0, double* Bf, double* c, double* H, double* G, double* dmin, double* dmax, double* fxmax,
                    double* hji) {
    ENTER(h, "NULL handle");
    const RecLayout& R = h->tab.rec;
    const size_t B = h->B, T = h->T, nx = h->nx, nu = h->nu;
    std::vector<double> rec(B * (size_t)R.rec_len);
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpy(rec.data(), h->d_rec, rec.size() * 8, cudaMemcpyDeviceToHost));
    for (size_t v = 0; v < B; v++) {
        const double* r = rec.data() + v * R.rec_len;
        for (size_t t = 0; t < T; t++) {
            const double* p = r + R.piece((int)t);
            if (A) memcpy(A + (v * T + t) * nx * nx, p + R.oA, nx * nx * 8);
            if (B0) memcpy(B0 + (v * T + t) * nx * nu, p + R.oB0, nx * nu * 8);
            if (Bf) memcpy(Bf + (v * T + t) * nx * nu, p + R.oBf, nx * nu * 8);
            if (c) memcpy(c + (v * T + t) * nx, p + R.oc, nx * 8);
            if (H) memcpy(H + (v * T + t) * 8, p + R.oH, 64);
            if (G) memcpy(G + (v * T + t) * 4, p + R.oG, 32);
            if (dmin) dmin[v * T + t] = p[R.odmin];
            if (dmax) dmax[v * T + t] = p[R.odmax];
            if (fxmax) fxmax[v * T + t] = p[R.ofxmax];
        }
        if (hji) memcpy(hji + v * 3, r + R.o_hji, 24);
    }
    return PGN_OK;
}
int pgn_get_solution(pgn_handle* h, double* x, double* y) {
    ENTER(h, "NULL handle");
    D2H(x, h->d_sol_x, (size_t)h->B * h->tab.n); D2H(y, h->d_sol_y, (size_t)h->B * h->tab.m);
    CK(cudaStreamSynchronize(h->stream));
    return PGN_OK;
}
int pgn_get_stats(pgn_handle* h, int32_t* iters, int32_t* status, double* pri, double* dua, double* rho, int32_t* rho_updates) {
    ENTER(h, "NULL handle");
    D2H(iters, h->d_iters, h->B); D2H(status, h->d_status, h->B); D2H(pri, h->d_pri_res, h->B); D2H(dua, h->d_dua_res, h->B); D2H(rho, h->d_rho, h->B);
    D2H(rho_updates, h->d_rho_updates, h->B);
    CK(cudaStreamSynchronize(h->stream));
    return PGN_OK;
}
int pgn_hji_lookup_device(pgn_handle* h, int32_t M, const double* d_x, double* d_V, double* d_gV) {
    ENTER(h, "NULL handle"); REQUIRE(d_x && d_V && d_gV && M >= 0, "bad argument");
    if (M == 0) return PGN_OK;
    { StageTimer T(h, 2); launch_hji_lookup(h, M, d_x, d_V, d_gV); }
    CK(cudaGetLastError());
    return PGN_OK;
}
int pgn_hji_lookup(pgn_handle* h, int32_t M, const double* x, double* V, double* gradV) {
    ENTER(h, "NULL handle"); REQUIRE(x && V && gradV && M >= 0, "bad argument");
    if (M == 0) return PGN_OK;
    double *dx = nullptr, *dV = nullptr, *dg = nullptr;
    CK(cudaMalloc(&dx, (size_t)M * 7 * 8)); CK(cudaMalloc(&dV, (size_t)M * 8)); CK(cudaMalloc(&dg, (size_t)M * 7 * 8));
    std::vector<double> xt((size_t)M * 7), gt((size_t)M * 7);
    for (int i = 0; i < M; i++) for (int d = 0; d < 7; d++) xt[(size_t)d * M + i] = x[(size_t)i * 7 + d];
    cudaError_t e = cudaMemcpy(dx, xt.data(), xt.size() * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) { launch_hji_lookup(h, M, dx, dV, dg); e = cudaStreamSynchronize(h->stream); }
    if (e == cudaSuccess) e = cudaMemcpy(V, dV, (size_t)M * 8, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(gt.data(), dg, gt.size() * 8, cudaMemcpyDeviceToHost);
    cudaFree(dx); cudaFree(dV); cudaFree(dg);
    if (e != cudaSuccess) return set_err(PGN_ECUDA, "hji lookup failed: %s", cudaGetErrorString(e));
    for (int i = 0; i < M; i++) for (int d = 0; d < 7; d++) gradV[(size_t)i * 7 + d] = gt[(size_t)d * M + i];
    return PGN_OK;
}
int pgn_set_path_search_window(pgn_handle* h, int32_t half_width) {
    ENTER(h, "NULL handle");
    REQUIRE(half_width >= 0, "half_width must be >= 0");
    h->epoch++;
    h->path_window = half_width;
    CK(cudaMemsetAsync(h->d_last_seg, 0xff, (size_t)h->B * 4, h->stream));
    return PGN_OK;
}
int pgn_set_hji_lookup_order(pgn_handle* h, int32_t mode) {
    ENTER(h, "NULL handle");
    REQUIRE(mode >= -1 && mode <= 2, "mode must be -1 (automatic), 0 (input order), 1 (cell order) or 2 (cell order with TMA-staged tiles)");
    REQUIRE(mode != 2 || h->hji_tma_valid, "no tensor map for this grid: the TMA-staged lookup is not available");
    h->hji_sort = mode;
    return PGN_OK;
}
int pgn_set_hji_policy(pgn_handle* h, int32_t on) { ENTER(h, "NULL handle"); h->epoch++; h->hji_policy = on != 0; return PGN_OK; }
int pgn_get_hji_values(pgn_handle* h, double* V, double* gradV) {
    ENTER(h, "NULL handle"); REQUIRE(V && gradV, "NULL argument");
    const size_t B = h->B;
    std::vector<double> hv(8 * B);
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpy(hv.data(), h->d_hji_val, 8 * B * 8, cudaMemcpyDeviceToHost));
    for (size_t v = 0; v < B; v++) {
        V[v] = hv[7 * B + v];
        for (int k = 0; k < 7; k++) gradV[v * 7 + k] = hv[k * B + v];
    }
    return PGN_OK;
}
int pgn_hji_optimal_control(pgn_handle* h, int32_t M, const double* x, const double* gradV, double* out) {
    ENTER(h, "NULL handle"); REQUIRE(x && gradV && out && M >= 0, "bad argument");
    if (M == 0) return PGN_OK;
    double *dx = nullptr, *dg = nullptr, *dout = nullptr;
    CK(cudaMalloc(&dx, (size_t)M * 7 * 8)); CK(cudaMalloc(&dg, (size_t)M * 7 * 8)); CK(cudaMalloc(&dout, (size_t)M * 2 * 8));
    cudaError_t e = cudaMemcpy(dx, x, (size_t)M * 7 * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(dg, gradV, (size_t)M * 7 * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) { launch_hji_optimal_control(h, M, dx, dg, dout); e = cudaStreamSynchronize(h->stream); }
    if (e == cudaSuccess) e = cudaMemcpy(out, dout, (size_t)M * 2 * 8, cudaMemcpyDeviceToHost);
    cudaFree(dx); cudaFree(dg); cudaFree(dout);
    if (e != cudaSuccess) return set_err(PGN_ECUDA, "hji optimal control failed: %s", cudaGetErrorString(e));
    return PGN_OK;
}
int pgn_device_controls(pgn_handle* h, double** d_out) { ENTER(h, "NULL handle"); REQUIRE(d_out, "NULL argument"); *d_out = h->d_controls; return PGN_OK; }
int pgn_device_stats(pgn_handle* h, int32_t** d_iters, int32_t** d_status) {
    ENTER(h, "NULL handle");
    if (d_iters) *d_iters = h->d_iters;
    if (d_status) *d_status = h->d_status;
    return PGN_OK;
}
int pgn_set_profiling(pgn_handle* h, int32_t on) {
    ENTER(h, "NULL handle");
    if (on == 3 && !h->d_trace) {
        h->trace_cap = 1 << 19;
        CK(cudaMalloc(&h->d_trace, (size_t)h->trace_cap * 24)); CK(cudaMalloc(&h->d_trace_n, sizeof(int)));
        CK(cudaMemset(h->d_trace_n, 0, sizeof(int)));
    }
    h->epoch++; h->profiling = on;
    return PGN_OK;
}
// profiling 3: the (start ns, end ns, part | QPs << 8 | SM << 32) records of the ADMM CTAs launched since the last reset; returns the count through *n
int pgn_get_admm_trace(pgn_handle* h, unsigned long long* out, int32_t max_entries, int32_t* n, int32_t reset) {
    ENTER(h, "NULL handle"); REQUIRE(out && n && max_entries >= 0, "bad argument");
    *n = 0;
    if (!h->d_trace) return PGN_OK;
    CK(cudaStreamSynchronize(h->stream));
    int cnt = 0;
    CK(cudaMemcpy(&cnt, h->d_trace_n, sizeof(int), cudaMemcpyDeviceToHost));
    cnt = std::min(std::min(cnt, h->trace_cap), (int)max_entries);
    CK(cudaMemcpy(out, h->d_trace, (size_t)cnt * 24, cudaMemcpyDeviceToHost));
    *n = cnt;
    if (reset) CK(cudaMemset(h->d_trace_n, 0, sizeof(int)));
    return PGN_OK;
}
int pgn_get_admm_cycles(pgn_handle* h, double* out, int32_t reset) {
    ENTER(h, "NULL handle"); REQUIRE(out, "NULL argument");
    unsigned long long c[512];
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpy(c, h->d_cycles, 4096, cudaMemcpyDeviceToHost));
    for (int i = 0; i < 512; i++) out[i] = (double)c[i];
    if (reset) CK(cudaMemset(h->d_cycles, 0, 4096));
    return PGN_OK;
}
int pgn_get_stage_ms(pgn_handle* h, double* out, int32_t reset) {
    ENTER(h, "NULL handle"); REQUIRE(out, "NULL argument");
    for (int i = 0; i < 6; i++) out[i] = h->stage_ms[i];
    out[6] = (double)h->launches; out[7] = (double)h->catchup_rounds;
    if (reset) { memset(h->stage_ms, 0, sizeof(h->stage_ms)); h->launches = 0; h->catchup_rounds = 0; }
    return PGN_OK;
}

}  // extern "C"
