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
    QpTables Q; char err[256];
    if (!build_qp_tables(kind, Ns, Nl, ord, Q, err, 256)) { printf("FAIL %s\n", err); return 1; }
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
    // emulate the kernel factorisation
    std::vector<double> L(Q.nnzL, 0.0), D(Nk), Dinv(Nk);
    for (int e = 0; e < Q.nnzA; e++) L[Q.a_lpos[e]] = Aval[e];
    for (int p = 0; p < Nk; p++) D[p] = Q.is_con[p] ? -rhoinv[Q.pos2idx[p]] : Pd[Q.pos2idx[p]];
    size_t npairs = Q.fac_a.size();
    for (int l = 0; l < Q.nlev; l++) {
        for (uint32_t t = Q.ftgt_ptr[l]; t < Q.ftgt_ptr[l + 1]; t++) {
            int id = Q.ftgt_id[t];
            if (id >= Q.nnzL) { int j = id - Q.nnzL; double s = D[j]; for (uint32_t x = Q.fac_ptr[t]; x < Q.fac_ptr[t + 1]; x++) { double v = L[Q.fac_a[x]]; s -= v * v * D[Q.fac_k[x]]; } D[j] = s; Dinv[j] = 1.0 / s; }
            else { double s = L[id]; for (uint32_t x = Q.fac_ptr[t]; x < Q.fac_ptr[t + 1]; x++) s -= L[Q.fac_a[x]] * L[Q.fac_b[x]] * D[Q.fac_k[x]]; L[id] = s; }
        }
        for (uint32_t t = Q.ftgt_ptr[l]; t < Q.ftgt_ptr[l + 1]; t++) { int id = Q.ftgt_id[t]; if (id < Q.nnzL) L[id] *= Dinv[Q.ftgt_col[t]]; }
    }
    // range inverses in place (two-phase per level, as in the kernel)
    for (size_t l = 0; l + 1 < Q.itgt_ptr.size(); l++) {
        std::vector<double> v;
        for (uint32_t t = Q.itgt_ptr[l]; t < Q.itgt_ptr[l + 1]; t++) { double acc = L[Q.itgt_id[t]]; for (uint32_t x = Q.inv_ptr[t]; x < Q.inv_ptr[t + 1]; x++) acc += L[Q.inv_a[x]] * L[Q.inv_b[x]]; v.push_back(-acc); }
        for (uint32_t t = Q.itgt_ptr[l]; t < Q.itgt_ptr[l + 1]; t++) L[Q.itgt_id[t]] = v[t - Q.itgt_ptr[l]];
    }
    // dense tail: packed copy of L[tail, tail] and its explicit inverse (column-wise forward substitution), as in the kernel
    const int ts = Q.tail_start, Dm = Q.tail_dim;
    std::vector<double> Ld(Dm * (Dm - 1) / 2 + 1, 0.0), Ti(Dm * (Dm - 1) / 2 + 1, 0.0);
    for (size_t e = 0; e < Q.tl_src.size(); e++) Ld[Q.tl_dst[e]] = L[Q.tl_src[e]];
    for (int j = 0; j < Dm; j++) for (int r = j + 1; r < Dm; r++) {
        double acc = 0; int rb = r * (r - 1) / 2;
        for (int k = j + 1; k < r; k++) acc += Ld[rb + k] * Ti[k * (k - 1) / 2 + j];
        Ti[rb + j] = -(Ld[rb + j] + acc);
    }
    // solve K x = b with the step programs
    std::vector<double> b(Nk), x(Nk), tmp(Nk, 0.0);
    for (auto& v : b) v = U(rng);
    x = b;
    const std::vector<uint32_t>* segs[4] = {&Q.fwd_ext, &Q.fwd_in, &Q.bwd_in, &Q.bwd_ext};
    auto run = [&](const std::vector<uint32_t>& st) {
        for (size_t i = 0; i < st.size(); i += 2) {
            int r0 = st[i] & 0xffff, nrows = st[i] >> 16, fl = st[i + 1] >> 8, sg = (fl & STEP_SEG_MASK) >> 1;
            std::vector<double>& in = (fl & STEP_SRC_TMP) ? tmp : x;
            std::vector<double>& out = (fl & STEP_DST_TMP) ? tmp : x;
            std::vector<double> res(nrows);
            for (int rr = 0; rr < nrows; rr++) {
                int r = r0 + rr; uint32_t rd = (*segs[sg])[r]; int base = rd & 0xffff, len = rd >> 16; double acc = 0;
                for (int e = base; e < base + len; e++) acc += (sg < 2) ? L[e] * in[Q.lrow_col[e]] : L[Q.lcol_val[e]] * in[Q.lcol_row[e]];
                double xv = in[r]; if (fl & STEP_SCALE) xv *= Dinv[r];
                res[rr] = (fl & STEP_ADD) ? xv + acc : xv - acc;
            }
            for (int rr = 0; rr < nrows; rr++) out[r0 + rr] = res[rr];
        }
    };
    run(Q.step_f);
    std::vector<double> t2(Nk);
    for (int rr = 0; rr < Dm; rr++) { double s = tmp[ts + rr]; for (int k = 0; k < rr; k++) s += Ti[rr * (rr - 1) / 2 + k] * tmp[ts + k]; x[ts + rr] = s * Dinv[ts + rr]; }
    for (int rr = 0; rr < Dm; rr++) { double s = x[ts + rr]; for (int k = rr + 1; k < Dm; k++) s += Ti[k * (k - 1) / 2 + rr] * x[ts + k]; t2[ts + rr] = s; }
    for (int rr = 0; rr < Dm; rr++) x[ts + rr] = t2[ts + rr];
    run(Q.step_b);
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
    int wmax = 0; for (int l = 0; l < Q.nlev; l++) wmax = std::max(wmax, (int)(Q.lvl_ptr[l + 1] - Q.lvl_ptr[l]));
    printf("kind=%d N=%d n=%d m=%d Nk=%d nnzA=%d nnzL=%d nlev=%d maxwidth=%d pairs=%zu rec_len=%d tail_level=%d tail_dim=%d res=%.3e kadj_err=%.3e\n", kind, Q.N, n, m, Nk, Q.nnzA,
           Q.nnzL, Q.nlev, wmax, npairs, Q.rec.rec_len, Q.tail_level, Q.tail_dim, res, kadj_err);
    printf("  ranges:"); for (size_t k = 0; k + 1 < Q.range_lvl.size(); k += 2) printf(" [%d,%d)", Q.range_lvl[k], Q.range_lvl[k + 1]);
    { size_t mt = 0; for (size_t l = 0; l + 1 < Q.itgt_ptr.size(); l++) mt = std::max(mt, (size_t)(Q.itgt_ptr[l + 1] - Q.itgt_ptr[l])); printf("  steps fwd %zu bwd %zu  inverse levels %zu max targets/level %zu inv pairs %zu\n", Q.step_f.size() / 2, Q.step_b.size() / 2, Q.itgt_ptr.size() - 1, mt, Q.inv_a.size()); }
    return (res < 1e-8 && kadj_err < 1e-10) ? 0 : 2;
}
