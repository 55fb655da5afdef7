// pgn_structure.cpp — host-side static analysis of the tracking QP (see pgn_structure.h).
//
// Canonical QP  min 1/2 x'Px + q'x  s.t.  l <= Ax <= u, variables and rows in the construction order of the reference:
//   coupled   (coupled_lat_long.jl:233-292): x = [q(6xN) u(2xN) sigma(2x(N-1)) sigma_HJI(N_short) ddelta(N-1) dFx(N-1)]
//   decoupled (decoupled_lat_long.jl:162-211): x = [q(4xN) delta(N) sigma(2x(N-1)) ddelta(N-1)]
// with every row written as  l <= a'x <= u  (variables on the left).  Parametron/MOI may permute or negate rows; ADMM iterates
// are invariant to both up to rounding.
#include "pgn_structure.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <set>

namespace pgn {

namespace {

struct Builder {
    QpTables& Q;
    int row = 0;
    explicit Builder(QpTables& q) : Q(q) {}
    void a(int col, int src) { Q.a_row.push_back(row); Q.a_col.push_back(col); Q.a_src.push_back(src); }
    void end_row(uint8_t lt, int li, uint8_t ut, int ui) {
        Q.l_type.push_back(lt); Q.l_idx.push_back(li); Q.u_type.push_back(ut); Q.u_idx.push_back(ui);
        row++;
    }
};

std::vector<int> constrained_min_degree(int Nk, const std::vector<std::vector<int>>& adj0, const std::vector<int>& cls) {
    std::vector<std::vector<int>> adj = adj0;
    std::vector<char> alive(Nk, 1);
    std::vector<int> perm;
    perm.reserve(Nk);
    int maxc = 0;
    for (int c : cls) maxc = std::max(maxc, c);
    for (int c = 0; c <= maxc; c++) {
        for (;;) {
            int best = -1;
            size_t bd = (size_t)-1;
            for (int v = 0; v < Nk; v++)
                if (alive[v] && cls[v] == c && adj[v].size() < bd) { bd = adj[v].size(); best = v; }
            if (best < 0) break;
            int v = best;
            alive[v] = 0;
            perm.push_back(v);
            std::vector<int> nb = adj[v];
            for (int a : nb) { auto& A = adj[a]; A.erase(std::lower_bound(A.begin(), A.end(), v)); }
            for (size_t x = 0; x < nb.size(); x++)
                for (size_t y = x + 1; y < nb.size(); y++) {
                    int a = nb[x], b = nb[y];
                    auto& A = adj[a];
                    auto it = std::lower_bound(A.begin(), A.end(), b);
                    if (it == A.end() || *it != b) {
                        A.insert(it, b);
                        auto& Bv = adj[b];
                        Bv.insert(std::lower_bound(Bv.begin(), Bv.end(), a), a);
                    }
                }
            adj[v].clear();
        }
    }
    return perm;
}

void nd_classes(int lo, int hi, int depth, int leaf, std::vector<int>& sep_depth) {
    if (hi - lo + 1 <= leaf) return;
    int mid = (lo + hi) / 2;
    sep_depth[mid] = depth;
    nd_classes(lo, mid - 1, depth + 1, leaf, sep_depth);
    nd_classes(mid + 1, hi, depth + 1, leaf, sep_depth);
}

}  // namespace

bool build_qp_tables(int kind, int N_short, int N_long, int ordering, QpTables& Q, char* err, int errlen) {
    auto fail = [&](const char* msg) { snprintf(err, errlen, "%s", msg); return false; };
    if (N_short < 1 || N_long < 0) return fail("N_short must be >= 1 and N_long >= 0");
    Q = QpTables();
    const int N = 1 + N_short + N_long, T = N - 1, Ns = N_short;
    const bool cpl = (kind == 0);
    const int nx = cpl ? 6 : 4, nu = cpl ? 2 : 1;
    Q.kind = kind; Q.N = N; Q.T = T; Q.Ns = Ns; Q.nx = nx; Q.nu = nu;
    RecLayout& R = Q.rec;
    R.nx = nx; R.nu = nu; R.T = T;
    R.oA = 0; R.oB0 = R.oA + nx * nx; R.oBf = R.oB0 + nx * nu; R.oc = R.oBf + nx * nu; R.oH = R.oc + nx; R.oG = R.oH + 8;
    R.odmin = R.oG + 4; R.odmax = R.odmin + 1; R.ofxmax = R.odmax + 1; R.piece_len = R.ofxmax + 1;
    R.o_qcurr = T * R.piece_len; R.o_ucurr = R.o_qcurr + nx; R.o_hji = R.o_ucurr + nu; R.o_dt = R.o_hji + 3; R.rec_len = R.o_dt + T;

    auto vq = [&](int i, int t) { return nx * t + i; };
    auto vu = [&](int i, int t) { return nx * N + nu * t + i; };
    auto vsig = [&](int i, int t) { return (nx + nu) * N + 2 * t + i; };
    auto vsh = [&](int t) { return (nx + nu) * N + 2 * T + t; };
    auto vdd = [&](int t) { return (nx + nu) * N + 2 * T + (cpl ? Ns : 0) + t; };
    auto vdf = [&](int t) { return (nx + nu) * N + 2 * T + Ns + T + t; };
    const int n = cpl ? (8 * N + 4 * T + Ns) : (5 * N + 3 * T);
    Q.n = n;
    Builder b(Q);
    const int ONE = -1, MONE = -2;
    auto P = [&](int t) { return R.piece(t); };
    // --- rows ---
    for (int t = 0; t < T; t++) for (int i = 0; i < 2; i++) { b.a(vsig(i, t), ONE); b.end_row(BND_CONST, CT_ZERO, BND_CONST, CT_PINF); }
    if (cpl) for (int t = 0; t < Ns; t++) { b.a(vsh(t), ONE); b.end_row(BND_CONST, CT_ZERO, BND_CONST, CT_PINF); }
    for (int t = 0; t < T; t++) { b.a(vu(0, t + 1), ONE); b.a(vu(0, t), MONE); b.a(vdd(t), MONE); b.end_row(BND_CONST, CT_ZERO, BND_CONST, CT_ZERO); }
    if (cpl) {
        for (int t = 0; t < T; t++) { b.a(vu(1, t + 1), ONE); b.a(vu(1, t), MONE); b.a(vdf(t), MONE); b.end_row(BND_CONST, CT_ZERO, BND_CONST, CT_ZERO); }
        for (int t = 0; t < N; t++) { b.a(vq(1, t), ONE); b.end_row(BND_CONST, CT_VMIN, BND_CONST, CT_PINF); }
        for (int t = 0; t < N; t++) { b.a(vq(1, t), ONE); b.end_row(BND_CONST, CT_NINF, BND_CONST, CT_VMAX); }
        for (int t = 0; t < N; t++) { b.a(vu(1, t), ONE); b.end_row(BND_CONST, CT_FXMIN_N, BND_CONST, CT_PINF); }
    }
    for (int i = 0; i < nx; i++) { b.a(vq(i, 0), ONE); b.end_row(BND_REC, R.o_qcurr + i, BND_REC, R.o_qcurr + i); }
    for (int i = 0; i < nu; i++) { b.a(vu(i, 0), ONE); b.end_row(BND_REC, R.o_ucurr + i, BND_REC, R.o_ucurr + i); }
    auto dyn_rows = [&](int t, bool ramp) {
        for (int i = 0; i < nx; i++) {
            for (int j = 0; j < nx; j++) b.a(vq(j, t), P(t) + R.oA + i * nx + j);
            for (int k = 0; k < nu; k++) b.a(vu(k, t), P(t) + R.oB0 + i * nu + k);
            if (ramp) for (int k = 0; k < nu; k++) b.a(vu(k, t + 1), P(t) + R.oBf + i * nu + k);
            b.a(vq(i, t + 1), MONE);
            b.end_row(BND_NEG_REC, P(t) + R.oc + i, BND_NEG_REC, P(t) + R.oc + i);
        }
    };
    for (int t = 0; t < Ns; t++) dyn_rows(t, false);
    if (cpl) for (int t = 0; t < Ns; t++) {
        b.a(vu(0, t), R.o_hji + 0); b.a(vu(1, t), R.o_hji + 1); b.a(vsh(t), ONE);
        b.end_row(BND_NEG_REC, R.o_hji + 2, BND_CONST, CT_PINF);
    }
    for (int t = Ns; t < T; t++) dyn_rows(t, true);
    for (int t = 0; t < T; t++) {
        b.a(vu(0, t + 1), ONE); b.end_row(BND_CONST, CT_NINF, BND_REC, P(t) + R.odmax);
        b.a(vu(0, t + 1), ONE); b.end_row(BND_REC, P(t) + R.odmin, BND_CONST, CT_PINF);
        if (cpl) { b.a(vu(1, t + 1), ONE); b.end_row(BND_CONST, CT_NINF, BND_REC, P(t) + R.ofxmax); }
        const int iUy = cpl ? 2 : 0, ir = cpl ? 3 : 1;
        for (int k = 0; k < 4; k++) {
            b.a(vq(iUy, t + 1), P(t) + R.oH + 2 * k); b.a(vq(ir, t + 1), P(t) + R.oH + 2 * k + 1); b.a(vsig(k / 2, t), MONE);
            b.end_row(BND_CONST, CT_NINF, BND_REC, P(t) + R.oG + k);
        }
        b.a(vdd(t), ONE); b.end_row(BND_CONST, CT_NINF, BND_DT_SCALED, R.o_dt + t);
        b.a(vdd(t), ONE); b.end_row(BND_NEG_DT_SCALED, R.o_dt + t, BND_CONST, CT_PINF);
    }
    const int m = b.row;
    Q.m = m; Q.nnzA = (int)Q.a_row.size();
    // --- cost tables ---
    Q.P_mode.assign(n, PQ_ZERO); Q.q_mode.assign(n, PQ_ZERO); Q.P_w.assign(n, W_NONE); Q.q_w.assign(n, W_NONE);
    Q.P_t.assign(n, 0); Q.q_t.assign(n, 0); Q.q_hji_t.assign(n, 0xFFFF);
    auto setP = [&](int v, int mode, int w, int t) { Q.P_mode[v] = mode; Q.P_w[v] = w; Q.P_t[v] = (uint16_t)(R.o_dt + t); };
    auto setq = [&](int v, int mode, int w, int t) { Q.q_mode[v] = mode; Q.q_w[v] = w; Q.q_t[v] = (uint16_t)(R.o_dt + t); };
    for (int t = 0; t < T; t++) {
        if (cpl) {
            setP(vq(0, t + 1), PQ_TIMES_DT, W_Q_DS, t); setP(vq(4, t + 1), PQ_TIMES_DT, W_Q_DPSI, t); setP(vq(5, t + 1), PQ_TIMES_DT, W_Q_E, t);
            setP(vu(0, t + 1), PQ_TIMES_DT, W_R_DELTA, t); setP(vu(1, t + 1), PQ_TIMES_DT, W_R_FX, t);
            setP(vdd(t), PQ_OVER_DT, W_R_DDELTA, t); setP(vdf(t), PQ_OVER_DT, W_R_DFX, t);
        } else {
            setP(vq(2, t + 1), PQ_TIMES_DT, W_Q_DPSI, t); setP(vq(3, t + 1), PQ_TIMES_DT, W_Q_E, t);
            setP(vu(0, t + 1), PQ_TIMES_DT, W_R_DELTA, t); setP(vdd(t), PQ_OVER_DT, W_R_DDELTA, t);
        }
        setq(vsig(0, t), PQ_TIMES_DT, W_W_BETA, t); setq(vsig(1, t), PQ_TIMES_DT, W_W_R, t);
    }
    if (cpl) for (int t = 0; t < Ns; t++) { Q.q_mode[vsh(t)] = PQ_CONST; Q.q_w[vsh(t)] = W_W_HJI; Q.q_hji_t[vsh(t)] = (uint16_t)t; }
    Q.var_u1_delta = vu(0, 1);
    Q.var_u1_fx = cpl ? vu(1, 1) : -1;

    // --- KKT graph: nodes 0..n-1 variables, n..n+m-1 constraints ---
    const int Nk = n + m;
    Q.Nk = Nk;
    if (Nk >= 65535) return fail("KKT dimension exceeds the 16-bit index range of the device tables");
    std::vector<std::vector<int>> adj(Nk);
    for (int e = 0; e < Q.nnzA; e++) { adj[Q.a_col[e]].push_back(n + Q.a_row[e]); adj[n + Q.a_row[e]].push_back(Q.a_col[e]); }
    for (auto& a : adj) { std::sort(a.begin(), a.end()); a.erase(std::unique(a.begin(), a.end()), a.end()); }
    // elimination classes: 0 = interior, >0 = stage separators (deeper separators first)
    std::vector<int> cls(Nk, 0);
    if (ordering == 0) {
        std::vector<int> sep_depth(N, -1);
        nd_classes(0, N - 1, 0, 3, sep_depth);
        int md = 0;
        for (int d : sep_depth) md = std::max(md, d);
        for (int t = 0; t < N; t++)
            if (sep_depth[t] >= 0) {
                int c = md - sep_depth[t] + 1;
                for (int i = 0; i < nx; i++) cls[vq(i, t)] = c;
                for (int i = 0; i < nu; i++) cls[vu(i, t)] = c;
            }
    }
    std::vector<int> perm = constrained_min_degree(Nk, adj, cls);
    std::vector<int> pos(Nk);
    for (int k = 0; k < Nk; k++) pos[perm[k]] = k;

    // symbolic factorisation in the elimination order `perm`
    auto symbolic = [&](const std::vector<int>& posv, std::vector<std::vector<int>>& colstruct) {
        colstruct.assign(Nk, {});
        std::vector<std::set<int>> S(Nk);
        for (int v = 0; v < Nk; v++) for (int w : adj[v]) if (posv[w] > posv[v]) S[posv[v]].insert(posv[w]);
        for (int k = 0; k < Nk; k++) {
            if (S[k].empty()) continue;
            int p = *S[k].begin();
            for (int i : S[k]) if (i != p) S[p].insert(i);
            colstruct[k].assign(S[k].begin(), S[k].end());
        }
    };
    std::vector<std::vector<int>> cs;
    symbolic(pos, cs);
    std::vector<int> level(Nk, 0);
    for (int k = 0; k < Nk; k++) for (int i : cs[k]) level[i] = std::max(level[i], level[k] + 1);
    // re-sort by level (any topological order of the elimination DAG gives the same fill); positions then group by level
    std::vector<int> order(Nk);
    for (int k = 0; k < Nk; k++) order[k] = k;
    std::stable_sort(order.begin(), order.end(), [&](int a, int c) { return level[a] < level[c]; });
    std::vector<int> newpos_of_old(Nk);
    for (int k = 0; k < Nk; k++) newpos_of_old[order[k]] = k;
    std::vector<int> pos2(Nk);
    for (int v = 0; v < Nk; v++) pos2[v] = newpos_of_old[pos[v]];
    symbolic(pos2, cs);
    std::fill(level.begin(), level.end(), 0);
    for (int k = 0; k < Nk; k++) for (int i : cs[k]) level[i] = std::max(level[i], level[k] + 1);
    for (int k = 1; k < Nk; k++) if (level[k] < level[k - 1]) return fail("internal: level order not monotone after re-sort");
    const int nlev = level[Nk - 1] + 1;
    Q.nlev = nlev;
    Q.lvl_ptr.assign(nlev + 1, 0);
    for (int k = 0; k < Nk; k++) Q.lvl_ptr[level[k] + 1]++;
    for (int l = 0; l < nlev; l++) Q.lvl_ptr[l + 1] += Q.lvl_ptr[l];

    Q.pos_var.resize(n); Q.pos_con.resize(m); Q.is_con.assign(Nk, 0); Q.pos2idx.assign(Nk, 0);
    for (int j = 0; j < n; j++) { Q.pos_var[j] = (uint16_t)pos2[j]; Q.pos2idx[pos2[j]] = (uint16_t)j; }
    for (int i = 0; i < m; i++) { Q.pos_con[i] = (uint16_t)pos2[n + i]; Q.is_con[pos2[n + i]] = 1; Q.pos2idx[pos2[n + i]] = (uint16_t)i; }

    // dense tail: trailing levels of width <= 4 (the top separators of the nested dissection), at most 64 positions and small enough
    // for its packed dense copy to fit the 2*Nk-double scratch region of the kernel
    {
        int Lt = nlev;
        while (Lt > 1) {
            int w = Q.lvl_ptr[Lt] - Q.lvl_ptr[Lt - 1];
            int D = Nk - Q.lvl_ptr[Lt - 1];
            if (w > 4 || D > 64 || D * (D - 1) / 2 > 2 * Nk - 8) break;
            Lt--;
        }
        if (nlev - Lt < 4) Lt = nlev;      // not worth it
        Q.tail_level = Lt; Q.tail_start = Q.lvl_ptr[Lt]; Q.tail_dim = Nk - Q.tail_start;
    }
    std::vector<std::vector<int>> rows(Nk);
    for (int k = 0; k < Nk; k++) for (int i : cs[k]) rows[i].push_back(k);   // ascending k by construction
    // level ranges of the sparse part [1, tail_level): the unit lower block L[range, range] is replaced, after every numeric
    // factorisation, by its explicit inverse (same pattern once padded with the transitive closure), so a whole range costs two
    // parallel steps in the triangular solves instead of one step per level.  Ranges grow greedily while the closure adds
    // little fill.
    {
        auto closure = [&](int la, int lb, std::vector<std::set<int>>& M, long& snz, long& cnz, int& maxin) {
            const int pa = Q.lvl_ptr[la], pb = Q.lvl_ptr[lb];
            M.assign(pb - pa, {});
            snz = cnz = 0; maxin = 0;
            for (int r = pa; r < pb; r++) {
                for (int c : rows[r]) if (c >= pa) { snz++; M[r - pa].insert(c); for (int k : M[c - pa]) M[r - pa].insert(k); }
                cnz += (long)M[r - pa].size();
                maxin = std::max(maxin, (int)M[r - pa].size());
            }
        };
        int la = 1;
        Q.range_lvl.clear();
        while (la < Q.tail_level) {
            int lb = la + 1;
            std::vector<std::set<int>> M, M2;
            long snz, cnz; int mi;
            closure(la, lb, M, snz, cnz, mi);
            while (lb < Q.tail_level) {
                long s2, c2; int mi2;
                closure(la, lb + 1, M2, s2, c2, mi2);
                if (c2 <= s2 + s2 / 4 + 32 && mi2 <= 96) { lb++; M.swap(M2); snz = s2; cnz = c2; mi = mi2; }
                else break;
            }
            const int pa = Q.lvl_ptr[la];
            for (size_t x = 0; x < M.size(); x++) {            // pad the pattern of L with the closure
                std::set<int> merged(rows[pa + x].begin(), rows[pa + x].end());
                merged.insert(M[x].begin(), M[x].end());
                rows[pa + x].assign(merged.begin(), merged.end());
            }
            Q.range_lvl.push_back(la); Q.range_lvl.push_back(lb);
            la = lb;
        }
        for (auto& c : cs) c.clear();
        for (int i = 0; i < Nk; i++) for (int k : rows[i]) cs[k].push_back(i);       // ascending i
    }
    // L by rows (CSR) and by columns (CSC)
    size_t nnzL = 0;
    for (auto& c : cs) nnzL += c.size();
    if (nnzL + Nk >= 65535) return fail("nnz(L) exceeds the 16-bit index range of the device tables");
    Q.nnzL = (int)nnzL;
    Q.lrow_ptr.assign(Nk + 1, 0);
    for (int i = 0; i < Nk; i++) Q.lrow_ptr[i + 1] = (uint16_t)(Q.lrow_ptr[i] + rows[i].size());
    Q.lrow_col.resize(nnzL);
    for (int i = 0; i < Nk; i++) for (size_t x = 0; x < rows[i].size(); x++) Q.lrow_col[Q.lrow_ptr[i] + x] = (uint16_t)rows[i][x];
    auto lidx = [&](int i, int k) -> int {   // index of L(i,k) in CSR order, -1 if structurally zero
        auto& r = rows[i];
        auto it = std::lower_bound(r.begin(), r.end(), k);
        if (it == r.end() || *it != k) return -1;
        return Q.lrow_ptr[i] + (int)(it - r.begin());
    };
    Q.lcol_ptr.assign(Nk + 1, 0);
    for (int k = 0; k < Nk; k++) Q.lcol_ptr[k + 1] = (uint16_t)(Q.lcol_ptr[k] + cs[k].size());
    Q.lcol_row.resize(nnzL); Q.lcol_val.resize(nnzL);
    for (int k = 0; k < Nk; k++) for (size_t x = 0; x < cs[k].size(); x++) {
        Q.lcol_row[Q.lcol_ptr[k] + x] = (uint16_t)cs[k][x];
        Q.lcol_val[Q.lcol_ptr[k] + x] = (uint16_t)lidx(cs[k][x], k);
    }
    // A entries in position space
    Q.a_rowpos.resize(Q.nnzA); Q.a_colpos.resize(Q.nnzA); Q.a_lpos.resize(Q.nnzA);
    for (int e = 0; e < Q.nnzA; e++) {
        int rp = pos2[n + Q.a_row[e]], cp = pos2[Q.a_col[e]];
        Q.a_rowpos[e] = (uint16_t)rp; Q.a_colpos[e] = (uint16_t)cp;
        int li = lidx(std::max(rp, cp), std::min(rp, cp));
        if (li < 0) return fail("internal: KKT entry missing from the symbolic factor");
        Q.a_lpos[e] = (uint16_t)li;
    }
    {   // duplicate (row, col) pairs would alias one L slot
        std::vector<uint16_t> chk(Q.a_lpos);
        std::sort(chk.begin(), chk.end());
        if (std::adjacent_find(chk.begin(), chk.end()) != chk.end()) return fail("internal: duplicate A entry");
    }
    // off-diagonal KKT adjacency per position
    std::vector<std::vector<std::pair<int, int>>> kadj(Nk);
    for (int e = 0; e < Q.nnzA; e++) {
        kadj[Q.a_rowpos[e]].push_back({e, Q.a_colpos[e]});
        kadj[Q.a_colpos[e]].push_back({e, Q.a_rowpos[e]});
    }
    Q.kadj_ptr.assign(Nk + 1, 0);
    for (int p = 0; p < Nk; p++) Q.kadj_ptr[p + 1] = (uint16_t)(Q.kadj_ptr[p] + kadj[p].size());
    Q.kadj_e.resize(2 * Q.nnzA); Q.kadj_nb.resize(2 * Q.nnzA);
    for (int p = 0; p < Nk; p++) for (size_t x = 0; x < kadj[p].size(); x++) {
        Q.kadj_e[Q.kadj_ptr[p] + x] = (uint16_t)kadj[p][x].first;
        Q.kadj_nb[Q.kadj_ptr[p] + x] = (uint16_t)kadj[p][x].second;
    }
    // dense tail tables: sparse entries of L[tail, tail] -> packed strictly-lower dense index
    for (int i = Q.tail_start; i < Nk; i++)
        for (int x = Q.lrow_ptr[i]; x < Q.lrow_ptr[i + 1]; x++)
            if (Q.lrow_col[x] >= Q.tail_start) {
                int ii = i - Q.tail_start, jj = Q.lrow_col[x] - Q.tail_start;
                Q.tl_src.push_back((uint16_t)x); Q.tl_dst.push_back((uint16_t)(ii * (ii - 1) / 2 + jj));
            }
    // row descriptors (first entry | count << 16) of the four segments used by the range steps:
    //   forward : CSR row r  = [entries left of the range (external) | entries inside the range]
    //   backward: CSC column r = [entries inside the range | entries below the range (external)]
    std::vector<int> range_of(Nk, -1), range_pa, range_pb;
    {
        const int nr = (int)Q.range_lvl.size() / 2;
        for (int k = 0; k < nr; k++) {
            range_pa.push_back(Q.lvl_ptr[Q.range_lvl[2 * k]]); range_pb.push_back(Q.lvl_ptr[Q.range_lvl[2 * k + 1]]);
            for (int p = range_pa[k]; p < range_pb[k]; p++) range_of[p] = k;
        }
        // level 0 and the tail behave as ranges whose in-range part is handled elsewhere (none / dense)
        Q.fwd_ext.assign(Nk, 0); Q.fwd_in.assign(Nk, 0); Q.bwd_in.assign(Nk, 0); Q.bwd_ext.assign(Nk, 0);
        for (int r = 0; r < Nk; r++) {
            int pa, pb;
            if (range_of[r] >= 0) { pa = range_pa[range_of[r]]; pb = range_pb[range_of[r]]; }
            else if (r >= Q.tail_start) { pa = Q.tail_start; pb = Nk; }
            else { pa = 0; pb = Q.lvl_ptr[1]; }
            int e0 = Q.lrow_ptr[r], e1 = Q.lrow_ptr[r + 1], sp = e0;
            while (sp < e1 && Q.lrow_col[sp] < pa) sp++;
            Q.fwd_ext[r] = (uint32_t)e0 | ((uint32_t)(sp - e0) << 16);
            Q.fwd_in[r] = (uint32_t)sp | ((uint32_t)(e1 - sp) << 16);
            int c0 = Q.lcol_ptr[r], c1 = Q.lcol_ptr[r + 1], cp = c0;
            while (cp < c1 && Q.lcol_row[cp] < pb) cp++;
            Q.bwd_in[r] = (uint32_t)c0 | ((uint32_t)(cp - c0) << 16);
            Q.bwd_ext[r] = (uint32_t)cp | ((uint32_t)(c1 - cp) << 16);
            if ((sp - e0) > 255 || (e1 - sp) > 255 || (cp - c0) > 255 || (c1 - cp) > 255) return fail("row segment longer than 255 entries");
        }
    }
    // step programs of the sparse part (256 lanes): a step = rows [r0, r0+rows) x 2^sh lanes, <= 4 entries per lane;
    // word 0 = r0 | rows << 16, word 1 = sh | flags << 8
    {
        const int T = ADMM_THREADS;
        bool too_long = false;
        auto emit = [&](std::vector<uint32_t>& out, int pa, int pb, const std::vector<uint32_t>& desc, uint32_t flags) {
            int w = pb - pa, mx = 0;
            for (int r = pa; r < pb; r++) mx = std::max(mx, (int)(desc[r] >> 16));
            int sh = 0;
            while (sh < 5 && ((mx + (1 << sh) - 1) >> sh) > 4) sh++;
            if (((mx + (1 << sh) - 1) >> sh) > 4) too_long = true;
            while (sh < 5 && ((mx + (1 << sh) - 1) >> sh) > 2 && (w << (sh + 1)) <= T) sh++;
            int rows_per_pass = T >> sh;
            for (int a = 0; a < w; a += rows_per_pass) {
                int nrows = std::min(rows_per_pass, w - a);
                bool last = a + nrows >= w;
                out.push_back((uint32_t)(pa + a) | ((uint32_t)nrows << 16));
                out.push_back((uint32_t)sh | ((flags | (last ? (uint32_t)STEP_LAST : 0u)) << 8));
            }
        };
        const int nr = (int)range_pa.size();
        for (int k = 0; k < nr; k++) {
            emit(Q.step_f, range_pa[k], range_pb[k], Q.fwd_ext, STEP_SEG_FWD_EXT | STEP_DST_TMP);                       // t = b - L_ext y      (sol -> tmp)
            emit(Q.step_f, range_pa[k], range_pb[k], Q.fwd_in, STEP_SEG_FWD_IN | STEP_SRC_TMP | STEP_ADD);              // y = t + Minv t       (tmp -> sol)
        }
        if (Q.tail_dim > 0) emit(Q.step_f, Q.tail_start, Nk, Q.fwd_ext, STEP_SEG_FWD_EXT | STEP_DST_TMP);               // tail stage 1          (sol -> tmp)
        for (int k = nr - 1; k >= 0; k--) {
            emit(Q.step_b, range_pa[k], range_pb[k], Q.bwd_ext, STEP_SEG_BWD_EXT | STEP_DST_TMP | STEP_SCALE);          // u = w/D - L_ext' x    (sol -> tmp)
            emit(Q.step_b, range_pa[k], range_pb[k], Q.bwd_in, STEP_SEG_BWD_IN | STEP_SRC_TMP | STEP_ADD);              // x = u + Minv' u       (tmp -> sol)
        }
        emit(Q.step_b, 0, Q.lvl_ptr[1], Q.bwd_ext, STEP_SEG_BWD_EXT | STEP_SCALE);                                       // level 0: x = w/D - L' x  (sol -> sol)
        if (too_long) return fail("a row segment of L has more than 128 entries (unsupported by the solve step program)");
    }
    // inverse program: for every range, level by level, each in-range entry (i,j) becomes  M_ij = -(S_ij + sum_{j<k<i} S_ik M_kj)
    {
        Q.inv_ptr.push_back(0);
        Q.itgt_ptr.push_back(0);
        const int nr = (int)range_pa.size();
        for (int k = 0; k < nr; k++) {
            const int pa = range_pa[k];
            for (int l = Q.range_lvl[2 * k] + 1; l < Q.range_lvl[2 * k + 1]; l++) {      // the first level of a range has no in-range entries
                for (int i = Q.lvl_ptr[l]; i < Q.lvl_ptr[l + 1]; i++) {
                    const int e0 = (int)(Q.fwd_in[i] & 0xffff), cnt = (int)(Q.fwd_in[i] >> 16);
                    for (int x = e0; x < e0 + cnt; x++) {
                        const int j = Q.lrow_col[x];
                        Q.itgt_id.push_back((uint16_t)x);
                        for (int y = x + 1; y < e0 + cnt; y++) {           // k = column of entry y, j < k < i
                            const int kk = Q.lrow_col[y];
                            const int mkj = lidx(kk, j);
                            if (mkj >= 0 && j >= pa) { Q.inv_a.push_back((uint16_t)y); Q.inv_b.push_back((uint16_t)mkj); }
                        }
                        Q.inv_ptr.push_back((uint32_t)Q.inv_a.size());
                    }
                }
                if (Q.itgt_id.size() - Q.itgt_ptr.back() > 4 * ADMM_THREADS) return fail("range inverse: too many targets in one level");
                Q.itgt_ptr.push_back((uint32_t)Q.itgt_id.size());
            }
        }
    }
    Q.lvl_gf.assign(nlev, 1); Q.lvl_gb.assign(nlev, 1); Q.lvl_gfac.assign(nlev, 1);
    // numeric factorisation program
    Q.ftgt_ptr.assign(nlev + 1, 0);
    Q.fac_ptr.push_back(0);
    for (int l = 0; l < nlev; l++) {
        struct Tgt { int id, col; std::vector<std::pair<int, int>> pairs; std::vector<int> ks; };
        std::vector<Tgt> tg;
        for (int j = Q.lvl_ptr[l]; j < Q.lvl_ptr[l + 1]; j++) {
            Tgt d; d.id = Q.nnzL + j; d.col = j;
            for (int kk : rows[j]) { int a = lidx(j, kk); d.pairs.push_back({a, a}); d.ks.push_back(kk); }
            tg.push_back(std::move(d));
            for (int i : cs[j]) {
                Tgt o; o.id = lidx(i, j); o.col = j;
                const auto &ri = rows[i], &rj = rows[j];
                size_t x = 0, y = 0;
                while (x < ri.size() && y < rj.size()) {
                    if (ri[x] >= j) break;
                    if (ri[x] == rj[y]) { o.pairs.push_back({Q.lrow_ptr[i] + (int)x, Q.lrow_ptr[j] + (int)y}); o.ks.push_back(ri[x]); x++; y++; }
                    else if (ri[x] < rj[y]) x++;
                    else y++;
                }
                tg.push_back(std::move(o));
            }
        }
        std::stable_sort(tg.begin(), tg.end(), [](const Tgt& a, const Tgt& c) { return a.pairs.size() > c.pairs.size(); });
        for (auto& t : tg) {
            Q.ftgt_id.push_back((uint16_t)t.id); Q.ftgt_col.push_back((uint16_t)t.col);
            for (size_t x = 0; x < t.pairs.size(); x++) {
                Q.fac_a.push_back((uint16_t)t.pairs[x].first); Q.fac_b.push_back((uint16_t)t.pairs[x].second); Q.fac_k.push_back((uint16_t)t.ks[x]);
            }
            Q.fac_ptr.push_back((uint32_t)Q.fac_a.size());
        }
        Q.ftgt_ptr[l + 1] = (uint32_t)Q.ftgt_id.size();
        {
            int mp = 0;
            for (auto& t : tg) mp = std::max(mp, (int)t.pairs.size());
            int g = 1;
            while (g < 32 && (mp + g - 1) / g > 4) g *= 2;
            while (g > 1 && (int)tg.size() * g > 1024) g /= 2;
            Q.lvl_gfac[l] = (uint8_t)g;
        }
    }
    return true;
}

}  // namespace pgn
