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
    // dense tail: packed copy of L[tail, tail] and its explicit inverse (column-wise forward substitution), as in the kernel
    const int ts = Q.tail_start, Dm = Q.tail_dim, Lt = Q.tail_level;
    std::vector<double> Ld(Dm * (Dm - 1) / 2 + 1, 0.0), Ti(Dm * (Dm - 1) / 2 + 1, 0.0);
    for (size_t e = 0; e < Q.tl_src.size(); e++) Ld[Q.tl_dst[e]] = L[Q.tl_src[e]];
    for (int j = 0; j < Dm; j++) for (int r = j + 1; r < Dm; r++) {
        double acc = 0; int rb = r * (r - 1) / 2;
        for (int k = j + 1; k < r; k++) acc += Ld[rb + k] * Ti[k * (k - 1) / 2 + j];
        Ti[rb + j] = -(Ld[rb + j] + acc);
    }
    // solve K x = b: sparse levels < tail_level, dense tail with the inverse, mirror image backwards
    std::vector<double> b(Nk), x(Nk), t(Nk);
    for (auto& v : b) v = U(rng);
    x = b;
    for (int l = 1; l < Lt; l++) for (int r = Q.lvl_ptr[l]; r < Q.lvl_ptr[l + 1]; r++) { double s = x[r]; for (int e = Q.lrow_ptr[r]; e < Q.lrow_ptr[r + 1]; e++) s -= L[e] * x[Q.lrow_col[e]]; x[r] = s; }
    for (int rr = 0; rr < Dm; rr++) { int r = ts + rr; double s = x[r]; for (int e = Q.lrow_ptr[r]; e < Q.lrow_split[r]; e++) s -= L[e] * x[Q.lrow_col[e]]; t[r] = s; }
    for (int rr = 0; rr < Dm; rr++) { double s = t[ts + rr]; for (int k = 0; k < rr; k++) s += Ti[rr * (rr - 1) / 2 + k] * t[ts + k]; x[ts + rr] = s * Dinv[ts + rr]; }
    for (int rr = 0; rr < Dm; rr++) { double s = x[ts + rr]; for (int k = rr + 1; k < Dm; k++) s += Ti[k * (k - 1) / 2 + rr] * x[ts + k]; t[ts + rr] = s; }
    for (int rr = 0; rr < Dm; rr++) x[ts + rr] = t[ts + rr];
    for (int l = Lt - 1; l >= 0; l--) for (int r = Q.lvl_ptr[l]; r < Q.lvl_ptr[l + 1]; r++) { double s = x[r] * Dinv[r]; for (int e = Q.lcol_ptr[r]; e < Q.lcol_ptr[r + 1]; e++) s -= L[Q.lcol_val[e]] * x[Q.lcol_row[e]]; x[r] = s; }
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
    return (res < 1e-8 && kadj_err < 1e-10) ? 0 : 2;
}
