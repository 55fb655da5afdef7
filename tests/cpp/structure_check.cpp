// Host-side check of the static KKT tables (pigeon.jl_b200/csrc/pgn_structure.cpp): emulates, serially, exactly what the ADMM
// kernel does with them (gather-form LDL' by levels, level-scheduled triangular solves) on a random quasi-definite KKT matrix
// with the QP's pattern and compares with a dense solve.  Prints one line of statistics; exit code 0 on success.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
#include "../../pigeon.jl_b200/csrc/pgn_structure.h"

using namespace pgn;

int main(int argc, char** argv) {
    int kind = argc > 1 ? atoi(argv[1]) : 0, Ns = argc > 2 ? atoi(argv[2]) : 10, Nl = argc > 3 ? atoi(argv[3]) : 20, ord = argc > 4 ? atoi(argv[4]) : 0;
    // argv[5]: verbose flag, argv[6] (or the environment variable PGN_CHECK_NWARPS): warps the programs are scheduled for
    const int nwarps = argc > 6 ? atoi(argv[6]) : (getenv("PGN_CHECK_NWARPS") ? atoi(getenv("PGN_CHECK_NWARPS")) : ADMM_THREADS / 32);
    QpTables Q; char err[256];
    if (!build_qp_tables(kind, Ns, Nl, ord, Q, err, 256, nwarps)) { printf("FAIL %s\n", err); return 1; }
    const int Nk = Q.Nk, n = Q.n, m = Q.m;
    std::mt19937_64 rng(7);
    std::uniform_real_distribution<double> U(-1, 1);
    std::vector<double> Aval(Q.nnzA), Pd(n), rhoinv(m);
    for (auto& a : Aval) a = U(rng);
    for (auto& p : Pd) p = std::fabs(U(rng)) + 1e-6;
    for (auto& r : rhoinv) r = (U(rng) > 0 ? 10.0 : 0.01);
    // dense K in position space
    std::vector<double> K((size_t)Nk * Nk, 0.0);
    for (int j = 0; j < n; j++) K[(size_t)Q.pos_var[j] * Nk + Q.pos_var[j]] = Pd[j];
    for (int i = 0; i < m; i++) K[(size_t)Q.pos_con[i] * Nk + Q.pos_con[i]] = -rhoinv[i];
    for (int e = 0; e < Q.nnzA; e++) { int r = Q.a_rowpos[e], c = Q.a_colpos[e]; K[(size_t)r * Nk + c] = Aval[e]; K[(size_t)c * Nk + r] = Aval[e]; }
    // emulate the kernel factorisation (unscaled form W = L D, one gather pass per level, no scaling pass)
    const int NS = Q.nslots;
    const int ts = Q.tail_start, Dm = Q.tail_dim, npk = Dm * (Dm + 1) / 2;
    std::vector<double> L(NS + npk, 0.0), Dinv(Nk + 1, 0.0);                      // L slots followed by the packed lower dense tail block
    for (int e = 0; e < Q.nnzA; e++) L[Q.a_slot[e]] = Aval[e];
    for (int p = 0; p < Nk; p++) {
        const double kpp = Q.is_con[p] ? -rhoinv[Q.pos2idx[p]] : Pd[Q.pos2idx[p]];
        if (p >= ts && Dm > 0) { const int i = p - ts; L[NS + i * (i + 1) / 2 + i] = kpp; Dinv[p] = 0.0; }
        else Dinv[p] = p < Q.fac_k0_end ? 1.0 / kpp : kpp;               // holds K_pp until the pivot is formed (level 0: formed here)
    }
    size_t npairs = 0;
    // a task list executed the way the warps do: every lane accumulates its K slots, the lanes of a row are summed, lane 0 of the row applies the result
    auto run_gather = [&](const std::vector<uint32_t>& tasks, uint32_t t0, uint32_t t1, const std::vector<uint32_t>& tgts, const std::vector<uint64_t>& ents,
                          std::vector<std::pair<uint32_t, double>>& results) {
        for (uint32_t t = t0; t < t1; t++) {
            const uint32_t ebase = tasks[4 * t], w1 = tasks[4 * t + 1], w2 = tasks[4 * t + 2];
            const int rbase = w1 & 0xffff, nrows = (w1 >> 16) & 0xff, sh = w1 >> 24, K = w2 & 0xffff, g = 1 << sh;
            if (nrows > (32 >> sh)) { printf("FAIL task rows\n"); exit(3); }
            for (int rr = 0; rr < nrows; rr++) {
                double acc = 0;
                for (int sub = 0; sub < g; sub++)
                    for (int k = 0; k < K; k++) {
                        const uint64_t e = ents[ebase + k * 32 + (rr << sh) + sub];
                        const int a = e & 0xffff, bb = (e >> 16) & 0xffff, kk = (int)(e >> 32);
                        acc += L[a] * L[bb] * Dinv[kk];
                        if (a != Q.zslot) npairs++;
                    }
                results.push_back({tgts[rbase + rr], acc});
            }
            for (int lane = nrows << sh; lane < 32; lane++) for (int k = 0; k < K; k++) { const uint64_t e = ents[ebase + k * 32 + lane]; if ((int)(e & 0xffff) != Q.zslot) { printf("FAIL pad\n"); exit(3); } }
        }
    };
    for (size_t l = 0; l + 1 < Q.fac_lvl_ptr.size(); l++) {
        std::vector<std::pair<uint32_t, double>> res;
        run_gather(Q.fac_task, Q.fac_lvl_ptr[l], Q.fac_lvl_ptr[l + 1], Q.fac_tgt, Q.fac_ent, res);
        for (auto& r : res) {
            if (r.first & FAC_TGT_PIVOT) { const int j = r.first & 0x7fffffff; Dinv[j] = 1.0 / (Dinv[j] - r.second); }
            else L[r.first & 0xffff] -= r.second;
        }
    }
    // range inverses in place (two-phase per level, as in the kernel)
    for (size_t l = 0; l + 1 < Q.inv_lvl_ptr.size(); l++) {
        std::vector<std::pair<uint32_t, double>> res;
        run_gather(Q.inv_task, Q.inv_lvl_ptr[l], Q.inv_lvl_ptr[l + 1], Q.inv_tgt, Q.inv_ent, res);
        for (auto& r : res) { const int id = r.first & 0xffff, col = r.first >> 16; r.second = -(L[id] * Dinv[col] + r.second); }
        for (auto& r : res) L[r.first & 0xffff] = r.second;
    }
    // dense tail: symmetric sweep of the packed lower Schur complement over all pivots -> -S^-1 (as in the kernel: the pivot column is staged
    // in a small buffer one step ahead)
    double* S = L.data() + NS;
    auto PK = [](int i, int k) { return i >= k ? i * (i + 1) / 2 + k : k * (k + 1) / 2 + i; };
    for (int p = 0; p < Dm; p++) {
        std::vector<double> col(Dm);
        for (int i = 0; i < Dm; i++) col[i] = S[PK(i, p)];
        const double dinv = 1.0 / col[p];
        for (int i = 0; i < Dm; i++)
            for (int k = 0; k <= i; k++) {
                double& v = S[PK(i, k)];
                if (i == p && k == p) v = -dinv;
                else if (i == p) v = col[k] * dinv;
                else if (k == p) v = col[i] * dinv;
                else v -= col[i] * col[k] * dinv;
            }
    }
    if (L[Q.zslot] != 0.0) { printf("FAIL zslot written\n"); return 3; }
    // solve K x = b with the solve programs
    std::vector<double> b(Nk), x(Nk + 1, 0.0), tmp(Nk + 1, 0.0);
    for (auto& v : b) v = U(rng);
    for (int p = 0; p < Nk; p++) (p < Q.rhs_tmp_end ? tmp : x)[p] = b[p];       // the first range reads its right-hand side from the scratch vector
    std::vector<char> written(Nk, 0);
    auto run_phase = [&](int ph, bool bwd) {
        std::vector<std::pair<int, double>> outs;
        int dst_tmp = 0;
        for (int t = Q.sol_ph_ptr[ph]; t < Q.sol_ph_ptr[ph + 1]; t++) {
            const uint32_t ebase = Q.sol_task[4 * t], w1 = Q.sol_task[4 * t + 1], w2 = Q.sol_task[4 * t + 2];
            const int rbase = w1 & 0xffff, nrows = (w1 >> 16) & 0xff, sh = w1 >> 24, K = w2 & 0xffff, fl = w2 >> 16, g = 1 << sh;
            const std::vector<double>& in = (fl & TASK_SRC_TMP) ? tmp : x;
            dst_tmp = fl & TASK_DST_TMP;
            for (int rr = 0; rr < nrows; rr++) {
                double acc = 0;
                for (int sub = 0; sub < g; sub++)
                    for (int k = 0; k < K; k++) {
                        const int e = ebase + k * 32 + (rr << sh) + sub;
                        if (bwd) acc += L[Q.bent[e] & 0xffff] * in[Q.bent[e] >> 16];
                        else acc += L[e] * in[Q.fidx[e]];
                    }
                const int r = Q.sol_orow[rbase + rr];
                if (fl & TASK_SCALE_ACC) acc *= Dinv[r];
                double v = (fl & TASK_ADD) ? in[r] + acc : in[r] - acc;
                if (fl & TASK_SCALE_OUT) v *= Dinv[r];
                outs.push_back({r, v});
            }
        }
        for (auto& o : outs) { (dst_tmp ? tmp : x)[o.first] = o.second; if (!dst_tmp) written[o.first]++; }
    };
    for (int p = 0; p < Q.fwd_k0_end; p++) { x[p] = tmp[p] * Dinv[p]; written[p]++; }      // level 0 is formed with the right-hand side
    if (!Q.bwd_k0.empty() && !(Q.bwd_k0_phase >= 1 && Q.bwd_k0_phase + 1 < Q.n_fwd_ph + Q.n_bwd_ph)) { printf("FAIL host phase of the task-less columns\n"); return 3; }
    for (int ph = 0; ph < Q.n_fwd_ph; ph++) {
        if (ph == Q.bwd_k0_phase) for (uint16_t c : Q.bwd_k0) tmp[c] = x[c];
        run_phase(ph, false);
    }
    for (int i = 0; i < Dm; i++) { double acc = 0; for (int k = 0; k < Dm; k++) acc += S[PK(i, k)] * tmp[ts + k]; x[ts + i] = -acc; }
    for (int ph = Q.n_fwd_ph; ph < Q.n_fwd_ph + Q.n_bwd_ph; ph++) {
        if (ph == Q.bwd_k0_phase) for (uint16_t c : Q.bwd_k0) tmp[c] = x[c];
        run_phase(ph, true);
    }
    if (x[Nk] != 0.0 || tmp[Nk] != 0.0) { printf("FAIL zero element written\n"); return 3; }
    // residual ||K x - b||_inf and the kadj product against the dense one
    double res = 0, kadj_err = 0;
    for (int r = 0; r < Nk; r++) {
        double s = 0, off = 0;
        for (int c = 0; c < Nk; c++) { s += K[(size_t)r * Nk + c] * x[c]; if (c != r) off += K[(size_t)r * Nk + c] * x[c]; }
        res = std::fmax(res, std::fabs(s - b[r]));
        double t = 0;
        for (int e = Q.kadj_ptr[r]; e < Q.kadj_ptr[r + 1]; e++) t += Aval[Q.kadj_e[e]] * x[Q.kadj_nb[e]];
        kadj_err = std::fmax(kadj_err, std::fabs(t - off));
    }
    // Ruiz norm program (256-thread kernels): every position owned exactly once, and the maximum over its slots (pad slots read the always-zero
    // double behind the A values) equals the maximum over its adjacency list
    int rz_bad = 0;
    if (Q.rz_prog) {
        std::vector<double> Apad(Aval); Apad.push_back(0.0);
        std::vector<int> owned(Nk, 0);
        for (int u = 0, s0 = 0; u < RZP_U; s0 += rzp_k(u), u++)
            for (int t = 0; t < RZP_NT; t++) {
                const int p = Q.rz_pos[u * RZP_NT + t];
                if (p == 0xFFFF) continue;
                if (p >= Nk) { rz_bad++; continue; }
                owned[p]++;
                double a = 0, bref = 0;
                for (int k = 0; k < rzp_k(u); k++) { const int sl = s0 + k; const uint32_t e = (Q.rz_idx[(sl >> 1) * RZP_NT + t] >> (16 * (sl & 1))) & 0xffffu; if ((int)e > Q.nnzA) { rz_bad++; continue; } a = std::fmax(a, std::fabs(Apad[e])); }
                for (int e = Q.kadj_ptr[p]; e < Q.kadj_ptr[p + 1]; e++) bref = std::fmax(bref, std::fabs(Aval[Q.kadj_e[e]]));
                if (a != bref) rz_bad++;
            }
        for (int p = 0; p < Nk; p++) if (owned[p] != 1) rz_bad++;
    }
    if (tail_segments(Q.tail_dim) > 32 * nwarps) rz_bad++;      // the dense-tail sweep gives every eight-element row segment its own thread
    int wmax = 0; for (int l = 0; l < Q.nlev; l++) wmax = std::max(wmax, (int)(Q.lvl_ptr[l + 1] - Q.lvl_ptr[l]));
    printf("kind=%d N=%d n=%d m=%d Nk=%d nnzA=%d nnzL=%d nlev=%d maxwidth=%d pairs=%zu rec_len=%d tail_level=%d tail_dim=%d rz_prog=%d rz_bad=%d res=%.3e kadj_err=%.3e\n", kind, Q.N, n, m, Nk, Q.nnzA,
           Q.nnzL, Q.nlev, wmax, npairs, Q.rec.rec_len, Q.tail_level, Q.tail_dim, Q.rz_prog, rz_bad, res, kadj_err);
    printf("  ranges:"); for (size_t k = 0; k + 1 < Q.range_lvl.size(); k += 2) printf(" [%d,%d)", Q.range_lvl[k], Q.range_lvl[k + 1]);
    printf("  nslots %d  solve phases fwd %d bwd %d tasks %zu bent %zu  factor tasks %zu ents %zu  inverse levels %zu tasks %zu ents %zu max tasks/warp %d\n", Q.nslots, Q.n_fwd_ph, Q.n_bwd_ph,
           Q.sol_task.size() / 4, Q.bent.size(), Q.fac_task.size() / 4, Q.fac_ent.size(), Q.inv_lvl_ptr.size() - 1, Q.inv_task.size() / 4, Q.inv_ent.size(), Q.inv_max_tasks_per_warp);
    if (argc > 5) {   // verbose: per-phase task shapes
        for (int ph = 0; ph < Q.n_fwd_ph + Q.n_bwd_ph; ph++) {
            printf("  phase %2d (%s):", ph, ph < Q.n_fwd_ph ? "fwd" : "bwd");
            for (int t = Q.sol_ph_ptr[ph]; t < Q.sol_ph_ptr[ph + 1]; t++) printf(" %ux%d/K%u", (Q.sol_task[4 * t + 1] >> 16) & 0xff, 1 << (Q.sol_task[4 * t + 1] >> 24), Q.sol_task[4 * t + 2] & 0xffff);
            printf("\n");
        }
        for (int l = 0; l + 1 < (int)Q.fac_lvl_ptr.size(); l++) {
            printf("  factor pass %2d:", l);
            for (uint32_t t = Q.fac_lvl_ptr[l]; t < Q.fac_lvl_ptr[l + 1]; t++) printf(" %ux%d/K%u", (Q.fac_task[4 * t + 1] >> 16) & 0xff, 1 << (Q.fac_task[4 * t + 1] >> 24), Q.fac_task[4 * t + 2] & 0xffff);
            printf("\n");
        }
    }
    return (res < 1e-8 && kadj_err < 1e-10 && rz_bad == 0) ? 0 : 2;
}
