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
#include <cstdlib>
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

// ---- warp-task scheduling ------------------------------------------------------------------------------------------------------
struct SchedTask { int sh, K; std::vector<int> rows; };   // rows: indices into the phase's row list

// Packs rows (by decreasing length) into warp tasks: 32 >> sh rows per task, 2^sh lanes per row, K = ceil(maxlen / 2^sh) slots per lane.
std::vector<SchedTask> schedule_rows(const std::vector<int>& len, int kmax) {
    std::vector<int> ord(len.size());
    for (size_t i = 0; i < ord.size(); i++) ord[i] = (int)i;
    std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return len[a] > len[b]; });
    std::vector<SchedTask> out;
    size_t i = 0;
    while (i < ord.size()) {
        const int L = len[ord[i]];
        int sh = 0;
        while (sh < 5 && ((L + (1 << sh) - 1) >> sh) > kmax) sh++;
        SchedTask t; t.sh = sh; t.K = (L + (1 << sh) - 1) >> sh;
        const int nr = 32 >> sh;
        for (int k = 0; k < nr && i < ord.size(); k++, i++) t.rows.push_back(ord[i]);
        out.push_back(std::move(t));
    }
    return out;
}
// Picks the slots-per-lane cap that minimises a simple cost model of the phase on `nw` warps (task t runs on warp t % nw):
// the longest warp's latency chain against the issue slots of the whole CTA.
std::vector<SchedTask> schedule_phase(const std::vector<int>& len, int nw) {
    static const int ladder[] = {2, 3, 4, 6, 8, 12, 16, 24, 32, 64, 255};
    // experiments: PGN_SCHED="a,b,c" overrides the latency model lat = a + b K + c sh
    static double ca = 120.0, cb = 22.0, cc = 30.0;
    static bool init = false;
    if (!init) { init = true; if (const char* e = getenv("PGN_SCHED")) sscanf(e, "%lf,%lf,%lf", &ca, &cb, &cc); }
    std::vector<SchedTask> best;
    double best_cost = 1e300;
    for (int kmax : ladder) {
        std::vector<SchedTask> t = schedule_rows(len, kmax);
        bool ok = true;
        std::vector<double> lat(nw, 0.0);
        double issue = 0.0;
        for (size_t i = 0; i < t.size(); i++) {
            if (t[i].K > 255) ok = false;
            lat[i % nw] += ca + cb * t[i].K + cc * t[i].sh;      // measured: ~90 cycles per batch of 4 entries, ~30 per shuffle stage
            issue += 40.0 + 6.0 * t[i].K + 8.0 * t[i].sh;
        }
        if (!ok) continue;
        const double cost = std::max(*std::max_element(lat.begin(), lat.end()), issue / 2.0);
        if (cost < best_cost) { best_cost = cost; best = std::move(t); }
    }
    return best;
}

void nd_classes(int lo, int hi, int depth, int leaf, std::vector<int>& sep_depth) {
    if (hi - lo + 1 <= leaf) return;
    int mid = (lo + hi) / 2;
    sep_depth[mid] = depth;
    nd_classes(lo, mid - 1, depth + 1, leaf, sep_depth);
    nd_classes(mid + 1, hi, depth + 1, leaf, sep_depth);
}

}  // namespace

bool build_qp_tables(int kind, int N_short, int N_long, int ordering, QpTables& Q, char* err, int errlen, int nwarps, int tmem_layout) {
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

    // dense tail: trailing levels of width <= 4 (the top separators of the nested dissection), at most 64 positions.  Its Schur
    // complement is formed by one gather pass and inverted densely (symmetric sweep) on chip
    {
        int Lt = nlev;
        while (Lt > 1) {
            int w = Q.lvl_ptr[Lt] - Q.lvl_ptr[Lt - 1];
            int D = Nk - Q.lvl_ptr[Lt - 1];
            if (w > 4 || D > 64 || tail_segments(D) > 32 * nwarps) break;      // the sweep gives every eight-element row segment its own thread
            Lt--;
        }
        if (nlev - Lt < 4) Lt = nlev;      // not worth it
        Q.tail_level = Lt; Q.tail_start = Q.lvl_ptr[Lt]; Q.tail_dim = Nk - Q.tail_start;
    }
    std::vector<std::vector<int>> rows(Nk);
    for (int k = 0; k < Nk; k++) for (int i : cs[k]) rows[i].push_back(k);   // ascending k by construction
    // level ranges of the sparse part [0, tail_level): the unit lower block L[range, range] is replaced, after every numeric
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
        int la = 0;
        Q.range_lvl.clear();
        while (la < Q.tail_level) {
            int lb = la + 1;
            std::vector<std::set<int>> M, M2;
            long snz, cnz; int mi;
            closure(la, lb, M, snz, cnz, mi);
            while (lb < Q.tail_level) {
                long s2, c2; int mi2;
                closure(la, lb + 1, M2, s2, c2, mi2);
                // the first range (the stage-local eliminations above level 0) may double its entries: every range saved is two phases per solve
                const long budget = la == 0 ? 2 * s2 + 64 : s2 + s2 / 4 + 32;
                if (c2 <= budget && mi2 <= 96) { lb++; M.swap(M2); snz = s2; cnz = c2; mi = mi2; }
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
    // Ruiz norm program (256-thread kernels): positions by decreasing adjacency length, class u = rank / 256
    {
        Q.rz_prog = 0;
        Q.rz_pos.assign(RZP_U * RZP_NT, 0xFFFF);
        Q.rz_idx.assign((RZP_SLOTS / 2) * RZP_NT, 0);
        std::vector<int> byDeg(Nk);
        for (int p = 0; p < Nk; p++) byDeg[p] = p;
        std::stable_sort(byDeg.begin(), byDeg.end(), [&](int a, int c) { return kadj[a].size() > kadj[c].size(); });
        bool fits = Nk <= RZP_U * RZP_NT;
        const uint32_t pad = (uint32_t)Q.nnzA;      // the always-zero double behind the A values (a_pad_len): one address for every lane, no wavefront of its own
        for (int r = 0; r < Nk && fits; r++) {
            const int u = r / RZP_NT, deg = (int)kadj[byDeg[r]].size();
            if (deg < 1 || deg > rzp_k(u)) fits = false;
        }
        // slot order inside a list: the 16 lanes of a half warp read slot k together (one 8-byte shared-memory bank each, 16 banks) — every lane
        // picks, slot by slot, the remaining entry whose bank the lanes before it use least
        long wavefronts = 0;
        for (int u = 0, s0 = 0; u < RZP_U && fits; s0 += rzp_k(u), u++)
            for (int h0 = 0; h0 < RZP_NT; h0 += 16) {
                std::vector<std::vector<int>> left(16);
                for (int l = 0; l < 16; l++) {
                    const int r = u * RZP_NT + h0 + l;
                    if (r >= Nk) continue;
                    Q.rz_pos[u * RZP_NT + h0 + l] = (uint16_t)byDeg[r];
                    for (auto& en : kadj[byDeg[r]]) left[l].push_back(en.first);
                }
                for (int k = 0; k < rzp_k(u); k++) {
                    int use[16] = {0}, worst = 1;
                    for (int l = 0; l < 16; l++) {
                        uint32_t e = pad;
                        if (!left[l].empty()) {
                            size_t best = 0;
                            for (size_t x = 1; x < left[l].size(); x++) if (use[left[l][x] & 15] < use[left[l][best] & 15]) best = x;
                            e = (uint32_t)left[l][best];
                            left[l].erase(left[l].begin() + best);
                            worst = std::max(worst, ++use[e & 15]);
                        }
                        const int sl = s0 + k;
                        Q.rz_idx[(sl >> 1) * RZP_NT + h0 + l] |= e << (16 * (sl & 1));
                    }
                    wavefronts += worst;
                }
            }
        if (getenv("PGN_STRUCT_VERBOSE")) fprintf(stderr, "[pgn] Ruiz norm program: %ld shared-memory wavefronts per pass (%d slots x %d half warps)\n", wavefronts, RZP_SLOTS, RZP_NT / 16);
        Q.rz_prog = fits ? 1 : 0;
    }

    // ---- warp programs -----------------------------------------------------------------------------------------------------------
    const int NWARP = nwarps;
    Q.nwarps = nwarps;
    Q.tmem_layout = tmem_layout; Q.tmem_cols = 0;
    // the tensor-memory build has no shared memory to spare: its solve tasks are always cut by the 8-warp cost model (the most compact slot
    // layout), whatever the number of warps that run them
    const int SCHED_NW = tmem_layout ? 8 : NWARP;
    // TMEM layout: columns used so far per quadrant; a phase's task list is permuted inside every round of NWARP tasks (one task per warp and
    // round either way, so the time balance of the schedule is unchanged) such that the biggest task of the round goes to the emptiest quadrant
    int quad_cols[4] = {0, 0, 0, 0};
    auto place_tasks = [&](std::vector<SchedTask> t) {
        if (!tmem_layout) return t;
        std::vector<SchedTask> out(t.size());
        for (size_t r0 = 0; r0 < t.size(); r0 += NWARP) {
            const size_t n = std::min((size_t)NWARP, t.size() - r0);
            std::vector<int> by_k(n), warps(n);
            for (size_t i = 0; i < n; i++) { by_k[i] = (int)(r0 + i); warps[i] = (int)i; }
            std::stable_sort(by_k.begin(), by_k.end(), [&](int a, int b) { return t[a].K > t[b].K; });
            int q_add[4] = {0, 0, 0, 0};
            std::vector<char> used(n, 0);
            for (size_t i = 0; i < n; i++) {          // biggest first -> the free warp whose quadrant is emptiest
                int best = -1;
                for (size_t w = 0; w < n; w++) if (!used[w] && (best < 0 || quad_cols[w & 3] + q_add[w & 3] < quad_cols[best & 3] + q_add[best & 3])) best = (int)w;
                used[best] = 1;
                q_add[best & 3] += 2 * t[by_k[i]].K;
                out[r0 + best] = t[by_k[i]];
            }
            for (int q = 0; q < 4; q++) quad_cols[q] += q_add[q];
        }
        return out;
    };
    int quad_next[4] = {0, 0, 0, 0};
    auto tmem_place = [&](size_t task_in_phase, int K) {      // called once per task in program order
        if (!tmem_layout) return;
        const int q = (int)(task_in_phase % NWARP) & 3;
        Q.sol_tcol.push_back((uint16_t)quad_next[q]);
        quad_next[q] += 2 * K;
    };
    std::vector<int> range_pa, range_pb;
    for (size_t k = 0; k + 1 < Q.range_lvl.size(); k += 2) { range_pa.push_back(Q.lvl_ptr[Q.range_lvl[k]]); range_pb.push_back(Q.lvl_ptr[Q.range_lvl[k + 1]]); }
    const int nr = (int)range_pa.size();
    // the ranges tile [0, tail_start); the right-hand side of the first range is written to the scratch vector (its only forward phase is
    // the in-range one, scratch -> solution)
    if (nr == 0 || range_pa[0] != 0 || range_pb[nr - 1] != Q.tail_start) return fail("internal: level ranges do not tile the sparse part");
    Q.rhs_tmp_end = range_pb[0];
    std::vector<int> slot(nnzL, -1);          // CSR entry -> L slot
    // -- forward solve: L slots are allocated in the order the forward phases consume them
    Q.sol_ph_ptr.assign(1, 0);
    int nslots = 0;
    auto add_task_words = [&](std::vector<uint32_t>& tasks, int ebase, int rbase, const SchedTask& t, int flags) {
        tasks.push_back((uint32_t)ebase);
        tasks.push_back((uint32_t)rbase | ((uint32_t)t.rows.size() << 16) | ((uint32_t)t.sh << 24));
        tasks.push_back((uint32_t)t.K | ((uint32_t)flags << 16));
        tasks.push_back(0u);
    };
    const bool skip_k0 = getenv("PGN_NO_K0_SKIP") == nullptr;
    Q.fwd_k0_end = 0; Q.bwd_k0_phase = -1; Q.bwd_k0_warp0 = 0; Q.bwd_k0.clear(); Q.fac_k0_end = 0;
    auto fwd_phase = [&](int pa, int pb, int c_lo, int c_hi, int flags, int skip_below = 0) {      // rows [max(pa, skip_below), pb), CSR entries with c_lo <= col < c_hi
        std::vector<std::vector<int>> ents(pb - pa);
        std::vector<int> len, keep;
        for (int r = pa; r < pb; r++) {
            for (int x = Q.lrow_ptr[r]; x < Q.lrow_ptr[r + 1]; x++) if (Q.lrow_col[x] >= c_lo && Q.lrow_col[x] < c_hi) ents[r - pa].push_back(x);
            if (r >= skip_below) { keep.push_back(r - pa); len.push_back((int)ents[r - pa].size()); }
        }
        size_t tip = 0;
        for (const SchedTask& t : place_tasks(schedule_phase(len, SCHED_NW))) {
            tmem_place(tip++, t.K);
            const int g = 1 << t.sh, ebase = nslots, rbase = (int)Q.sol_orow.size();
            Q.fidx.resize(ebase + 32 * t.K, (uint16_t)Nk);
            for (size_t rr = 0; rr < t.rows.size(); rr++) {
                Q.sol_orow.push_back((uint16_t)(pa + keep[t.rows[rr]]));
                const auto& E = ents[keep[t.rows[rr]]];
                for (size_t x = 0; x < E.size(); x++) {
                    const int lane = (int)(rr << t.sh) + (int)(x % g), k = (int)(x / g), sl = ebase + k * 32 + lane;
                    slot[E[x]] = sl;
                    Q.fidx[sl] = Q.lrow_col[E[x]];
                }
            }
            nslots += 32 * t.K;
            add_task_words(Q.sol_task, ebase, rbase, t, flags);
        }
        Q.sol_ph_ptr.push_back((uint16_t)(Q.sol_task.size() / 4));
    };
    for (int k = 0; k < nr; k++) {
        if (k > 0) fwd_phase(range_pa[k], range_pb[k], 0, range_pa[k], TASK_DST_TMP);                                         // t = b - W_ext y^           (sol -> tmp)
        int skip_below = 0;
        if (k == 0 && skip_k0) {      // level 0 of the first range: no predecessors at all, y^ = t / d is formed with the right-hand side
            int e = 0;
            while (e < range_pb[0] && Q.lrow_ptr[e + 1] == Q.lrow_ptr[e]) e++;
            Q.fwd_k0_end = skip_below = e;
        }
        fwd_phase(range_pa[k], range_pb[k], range_pa[k], range_pb[k], TASK_SRC_TMP | TASK_ADD | TASK_SCALE_OUT, skip_below);   // y^ = (t + M t) / d         (tmp -> sol)
    }
    if (Q.tail_dim > 0) fwd_phase(Q.tail_start, Nk, 0, Q.tail_start, TASK_DST_TMP);                               // tail stage 1               (sol -> tmp)
    Q.n_fwd_ph = (int)Q.sol_ph_ptr.size() - 1;
    Q.zslot = nslots++;                       // entries inside the dense tail block get no slot: they live in the dense Schur complement
    nslots = (nslots + 31) & ~31;
    Q.nslots = nslots;
    Q.fidx.resize(nslots, (uint16_t)Nk);
    if (nslots + Nk >= 65535) return fail("L slots exceed the 16-bit index range of the device tables");
    // -- backward solve: (L slot, source) pairs in program order
    // mode 0: every column gets a task row; 1: the columns without entries go to `dropped` instead; 2: columns without entries that are in `drop_if` get none
    auto bwd_phase = [&](int pa, int pb, int r_lo, int r_hi, int flags, int mode = 0, std::vector<uint16_t>* dropped = nullptr, const std::vector<char>* drop_if = nullptr) {
        std::vector<std::vector<int>> ents(pb - pa);       // columns [pa, pb), CSC entries with r_lo <= row < r_hi
        std::vector<int> len, keep;
        for (int c = pa; c < pb; c++) {
            for (int x = Q.lcol_ptr[c]; x < Q.lcol_ptr[c + 1]; x++) if (Q.lcol_row[x] >= r_lo && Q.lcol_row[x] < r_hi) ents[c - pa].push_back(x);
            const bool empty = ents[c - pa].empty();
            if (mode == 1 && empty) { dropped->push_back((uint16_t)c); continue; }
            if (mode == 2 && empty && (*drop_if)[c]) continue;
            keep.push_back(c - pa); len.push_back((int)ents[c - pa].size());
        }
        size_t tip = 0;
        for (const SchedTask& t : place_tasks(schedule_phase(len, SCHED_NW))) {
            tmem_place(tip++, t.K);
            const int g = 1 << t.sh, ebase = (int)Q.bent.size(), rbase = (int)Q.sol_orow.size();
            Q.bent.resize(ebase + 32 * t.K, (uint32_t)Q.zslot | ((uint32_t)Nk << 16));
            for (size_t rr = 0; rr < t.rows.size(); rr++) {
                Q.sol_orow.push_back((uint16_t)(pa + keep[t.rows[rr]]));
                const auto& E = ents[keep[t.rows[rr]]];
                for (size_t x = 0; x < E.size(); x++) {
                    const int lane = (int)(rr << t.sh) + (int)(x % g), k = (int)(x / g);
                    Q.bent[ebase + k * 32 + lane] = (uint32_t)slot[Q.lcol_val[E[x]]] | ((uint32_t)Q.lcol_row[E[x]] << 16);
                }
            }
            add_task_words(Q.sol_task, ebase, rbase, t, flags);
        }
        Q.sol_ph_ptr.push_back((uint16_t)(Q.sol_task.size() / 4));
    };
    // the copies of the task-less columns of the first range run beside the tasks of the phase (after the first forward phase, which still
    // reads the scratch vector as right-hand side, and before the last backward phase, which reads the copies) that leaves the most warps idle
    int host_phase = -1, host_tasks = 1 << 30;
    const int n_ph_total = Q.n_fwd_ph + 2 * nr;
    for (int ph = 1; ph + 1 < n_ph_total && ph < Q.n_fwd_ph; ph++) {        // forward phases are final here; the backward ones are not built yet
        const int nt = Q.sol_ph_ptr[ph + 1] - Q.sol_ph_ptr[ph];
        if (nt < host_tasks) { host_tasks = nt; host_phase = ph; }
    }
    const bool k0_bwd = skip_k0 && host_phase >= 1 && nr >= 1;
    for (int k = nr - 1; k >= 0; k--) {
        if (k == 0 && k0_bwd) {
            bwd_phase(range_pa[k], range_pb[k], range_pb[k], Nk, TASK_DST_TMP | TASK_SCALE_ACC, 1, &Q.bwd_k0);     // v = y^ - (W_below' x) / d  (sol -> tmp)
            std::vector<char> isk0(Nk, 0);
            for (uint16_t c : Q.bwd_k0) isk0[c] = 1;
            bwd_phase(range_pa[k], range_pb[k], range_pa[k], range_pb[k], TASK_SRC_TMP | TASK_ADD, 2, nullptr, &isk0);   // x = v + M' v           (tmp -> sol)
            Q.bwd_k0_phase = host_phase; Q.bwd_k0_warp0 = host_tasks;      // tasks of the host phase: warps >= this count are idle in it (all warps copy if there is none)
            continue;
        }
        bwd_phase(range_pa[k], range_pb[k], range_pb[k], Nk, TASK_DST_TMP | TASK_SCALE_ACC);                       // v = y^ - (W_below' x) / d  (sol -> tmp)
        bwd_phase(range_pa[k], range_pb[k], range_pa[k], range_pb[k], TASK_SRC_TMP | TASK_ADD);                    // x = v + M' v               (tmp -> sol)
    }
    Q.n_bwd_ph = (int)Q.sol_ph_ptr.size() - 1 - Q.n_fwd_ph;
    if (tmem_layout) {
        Q.bsrc.resize(Q.bent.size());
        for (size_t e = 0; e < Q.bent.size(); e++) Q.bsrc[e] = (uint16_t)(Q.bent[e] >> 16);
        // a partial batch reads a whole group of four slot rows (8 columns): 6 columns of slack behind the last task of every quadrant
        Q.tmem_cols = *std::max_element(quad_next, quad_next + 4) + 6;
    }

    // A entries -> L slot, or (both ends in the tail) nslots + packed lower index i (i + 1) / 2 + j of the dense Schur complement
    Q.a_slot.resize(Q.nnzA);
    for (int e = 0; e < Q.nnzA; e++) {
        const int hi = std::max(Q.a_rowpos[e], Q.a_colpos[e]), lo = std::min(Q.a_rowpos[e], Q.a_colpos[e]);
        if (lo >= Q.tail_start) { const int ii = hi - Q.tail_start, jj = lo - Q.tail_start; Q.a_slot[e] = (uint16_t)(nslots + ii * (ii + 1) / 2 + jj); }
        else Q.a_slot[e] = (uint16_t)slot[Q.a_lpos[e]];
    }
    if (nslots + Q.tail_dim * (Q.tail_dim + 1) / 2 >= 32768) return fail("L slots exceed the index range of the gather programs");

    // generic emitter for the gather programs (factorisation, range inverses): targets with (a, b, k) entry lists
    struct Tgt { uint32_t tgt; std::vector<uint64_t> ents; };
    auto emit_level = [&](const std::vector<Tgt>& tg, std::vector<uint32_t>& tasks, std::vector<uint32_t>& lvl_ptr, std::vector<uint32_t>& tgts, std::vector<uint64_t>& ents) {
        std::vector<int> len(tg.size());
        for (size_t i = 0; i < tg.size(); i++) len[i] = (int)tg[i].ents.size();
        const uint64_t padent = (uint64_t)Q.zslot | ((uint64_t)Q.zslot << 16) | ((uint64_t)Nk << 32);      // 0 * 0 / Dinv[Nk]: element Nk of Dinv is never written by the factorisation (kept at 0)
        for (const SchedTask& t : schedule_phase(len, NWARP)) {
            const int g = 1 << t.sh, ebase = (int)ents.size(), rbase = (int)tgts.size();
            ents.resize(ebase + 32 * t.K, padent);
            for (size_t rr = 0; rr < t.rows.size(); rr++) {
                const Tgt& T = tg[t.rows[rr]];
                tgts.push_back(T.tgt);
                for (size_t x = 0; x < T.ents.size(); x++) ents[ebase + (x / g) * 32 + (rr << t.sh) + (x % g)] = T.ents[x];
            }
            add_task_words(tasks, ebase, rbase, t, 0);
        }
        lvl_ptr.push_back((uint32_t)(tasks.size() / 4));
    };
    auto ent3 = [](int a, int b, int k) { return (uint64_t)a | ((uint64_t)b << 16) | ((uint64_t)k << 32); };
    // numeric factorisation program (left-looking gathers, level scheduled), unscaled form W = L D.  Targets: PIVOT | position, or a slot
    // (L slot, or nslots + packed index of the dense tail Schur complement S = K_TT - sum_{k < tail} W_Tk W_Tk' / d_k, gathered by one
    // extra pass after the last sparse level)
    Q.fac_lvl_ptr.assign(1, 0);
    for (int l = 0; l < Q.tail_level; l++) {
        std::vector<Tgt> tg;
        for (int j = Q.lvl_ptr[l]; j < Q.lvl_ptr[l + 1]; j++) {
            Tgt d; d.tgt = FAC_TGT_PIVOT | (uint32_t)j;
            for (int kk : rows[j]) { int a = slot[lidx(j, kk)]; d.ents.push_back(ent3(a, a, kk)); }
            tg.push_back(std::move(d));
            for (int i : cs[j]) {
                Tgt o; o.tgt = (uint32_t)slot[lidx(i, j)];
                const auto &ri = rows[i], &rj = rows[j];
                size_t x = 0, y = 0;
                while (x < ri.size() && y < rj.size()) {
                    if (ri[x] >= j) break;
                    if (ri[x] == rj[y]) { o.ents.push_back(ent3(slot[Q.lrow_ptr[i] + (int)x], slot[Q.lrow_ptr[j] + (int)y], ri[x])); x++; y++; }
                    else if (ri[x] < rj[y]) x++;
                    else y++;
                }
                if (!o.ents.empty()) tg.push_back(std::move(o));      // W_ij = K_ij needs no work
            }
        }
        if (l == 0 && skip_k0 && Q.tail_level > 1) {      // level 0: pivots without a single update, d_j = K_jj — inverted where K_jj is written
            bool all_empty = true;
            for (const Tgt& t : tg) all_empty = all_empty && t.ents.empty() && (t.tgt & FAC_TGT_PIVOT);
            if (all_empty) { Q.fac_k0_end = Q.lvl_ptr[1]; continue; }
        }
        emit_level(tg, Q.fac_task, Q.fac_lvl_ptr, Q.fac_tgt, Q.fac_ent);
    }
    if (Q.tail_dim > 0) {
        std::vector<Tgt> tg;
        for (int i = Q.tail_start; i < Nk; i++)
            for (int j = Q.tail_start; j <= i; j++) {
                const int ii = i - Q.tail_start, jj = j - Q.tail_start;
                Tgt o; o.tgt = (uint32_t)(nslots + ii * (ii + 1) / 2 + jj);
                const auto &ri = rows[i], &rj = rows[j];
                size_t x = 0, y = 0;
                while (x < ri.size() && y < rj.size() && ri[x] < Q.tail_start && rj[y] < Q.tail_start) {
                    if (ri[x] == rj[y]) { o.ents.push_back(ent3(slot[Q.lrow_ptr[i] + (int)x], slot[Q.lrow_ptr[j] + (int)y], ri[x])); x++; y++; }
                    else if (ri[x] < rj[y]) x++;
                    else y++;
                }
                if (!o.ents.empty()) tg.push_back(std::move(o));
            }
        emit_level(tg, Q.fac_task, Q.fac_lvl_ptr, Q.fac_tgt, Q.fac_ent);
    }
    // inverse program: for every range, level by level, each in-range entry (i,j) becomes  M_ij = -(W_ij/d_j + sum_{j<k<i} W_ik/d_k M_kj)
    Q.inv_lvl_ptr.assign(1, 0);
    Q.inv_max_tasks_per_warp = 0;
    for (int k = 0; k < nr; k++) {
        const int pa = range_pa[k];
        for (int l = Q.range_lvl[2 * k] + 1; l < Q.range_lvl[2 * k + 1]; l++) {      // the first level of a range has no in-range entries
            std::vector<Tgt> tg;
            for (int i = Q.lvl_ptr[l]; i < Q.lvl_ptr[l + 1]; i++) {
                for (int x = Q.lrow_ptr[i]; x < Q.lrow_ptr[i + 1]; x++) {
                    const int j = Q.lrow_col[x];
                    if (j < pa) continue;
                    Tgt t; t.tgt = (uint32_t)slot[x] | ((uint32_t)j << 16);
                    for (int y = x + 1; y < Q.lrow_ptr[i + 1]; y++) {           // kk = column of entry y, j < kk < i
                        const int kk = Q.lrow_col[y];
                        const int mkj = lidx(kk, j);
                        if (mkj >= 0) t.ents.push_back(ent3(slot[y], slot[mkj], kk));
                    }
                    tg.push_back(std::move(t));
                }
            }
            const size_t before = Q.inv_task.size() / 4;
            emit_level(tg, Q.inv_task, Q.inv_lvl_ptr, Q.inv_tgt, Q.inv_ent);
            const int ntask = (int)(Q.inv_task.size() / 4 - before);
            Q.inv_max_tasks_per_warp = std::max(Q.inv_max_tasks_per_warp, (ntask + NWARP - 1) / NWARP);
        }
    }
    if (Q.inv_max_tasks_per_warp > INV_MAX_TASKS_PER_WARP) return fail("range inverse: too many targets in one level");
    return true;
}

}  // namespace pgn
