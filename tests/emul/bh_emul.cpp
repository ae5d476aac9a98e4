// bh_emul.cpp — TEST INFRASTRUCTURE: runs the engine's per-element algorithm core
// (barnes-hut-n-body_b200/csrc/bh_core.h, the same inline functions the CUDA kernels call)
// in serial loops on the CPU, so the key / layout / search / climb / criterion logic can be
// checked against the oracle where no GPU exists.  It is NOT a product path: nothing under
// barnes-hut-n-body_b200/ loads it, and the parallel primitives (onesweep sort, look-back
// scan, atomics ordering) are only exercised by the real kernels in the `-m gpu` tests.
#include <algorithm>
#include <cstring>
#include <numeric>
#include <vector>
#include "../../barnes-hut-n-body_b200/csrc/bh_core.h"
#include "../../barnes-hut-n-body_b200/csrc/bh_export.h"

struct Emul {
    BhRoot root{};
    int n = 0, n_in = 0, M = 0;
    std::vector<uint64_t> key_by_body, keys;
    std::vector<int> order, S, arrived;
    std::vector<BhCell> cell;
    std::vector<BhCellD> cd;
    std::vector<BhCellS> sk;
    std::vector<double> xm, ym;     // coordinates after the jitter replay (BH.kt:145-156 mutates bodies)
    std::vector<int> jflag;         // per body
    std::vector<int> jsorted;       // per sorted position (for the export)
    int n_ghost = 0, unsupported = 0;
    int64_t grid_mismatch = 0;
    int grid_exact = 0;
    BhTreeView view() {
        BhTreeView t{};
        t.keys = keys.data(); t.order = order.data(); t.S = S.data();
        t.cell = cell.data(); t.cd = cd.data(); t.sk = sk.data(); t.arrived = arrived.data();
        t.n_in = n_in; t.M = M;
        return t;
    }
};

static void build(Emul& e, int n, const double* x, const double* y, const double* m,
                  double rcx, double rcy, double rhalf) {
    e.root = BhRoot{rcx, rcy, rhalf, bh_key_levels(rhalf)};
    e.n = n;
    e.key_by_body.resize(n);
    const BhGrid grid = bh_make_grid(e.root);
    for (int b = 0; b < n; ++b) {
        e.key_by_body[b] = bh_root_contains(e.root, x[b], y[b]) ? bh_morton_key(e.root, x[b], y[b]) : BH_KEY_NOT_IN_TREE;
        // the closed-form key must agree with the literal descent whenever the grid is exact
        if (grid.exact && e.key_by_body[b] != BH_KEY_NOT_IN_TREE &&
            bh_morton_key_grid(grid, e.root.levels, x[b], y[b]) != e.key_by_body[b]) e.grid_mismatch++;
    }
    e.grid_exact = grid.exact;
    e.order.resize(n);
    std::iota(e.order.begin(), e.order.end(), 0);
    std::stable_sort(e.order.begin(), e.order.end(), [&](int a, int b) { return e.key_by_body[a] < e.key_by_body[b]; });
    e.n_in = 0;
    for (int b = 0; b < n; ++b) e.n_in += e.key_by_body[b] != BH_KEY_NOT_IN_TREE;
    e.keys.resize(e.n_in);
    for (int i = 0; i < e.n_in; ++i) e.keys[i] = e.key_by_body[e.order[i]];
    e.S.assign(e.n_in + 1, 0);
    for (int i = 0; i < e.n_in; ++i) {
        const int dprev = i > 0 ? bh_common_levels(e.keys[i - 1], e.keys[i], e.root.levels) : -1;
        const int dnext = i + 1 < e.n_in ? bh_common_levels(e.keys[i], e.keys[i + 1], e.root.levels) : -1;
        e.S[i + 1] = e.S[i] + (dnext > dprev ? dnext - dprev : 0);
    }
    e.M = e.n_in + e.S[e.n_in];
    // jitter regime: replay every run of equal keys (the engine does this in k_jitter)
    e.xm.assign(x, x + n); e.ym.assign(y, y + n);
    e.jflag.assign(n, 0);
    {
        std::vector<int> ident(n);
        std::iota(ident.begin(), ident.end(), 0);
        for (int i = 0; i + 1 < e.n_in;) {
            int j = i;
            while (j + 1 < e.n_in && e.keys[j + 1] == e.keys[i]) ++j;
            if (j > i) {
                int un = 0, gh = 0;
                bh_jitter_cluster(e.root, e.keys[i], e.order.data() + i, j - i + 1, ident.data(), e.xm.data(), e.ym.data(),
                                  e.jflag.data(), &un, &gh);
                e.n_ghost += gh; e.unsupported |= un;
            }
            i = j + 1;
        }
    }
    x = e.xm.data(); y = e.ym.data();
    e.jsorted.resize(e.n_in);
    for (int i = 0; i < e.n_in; ++i) e.jsorted[i] = e.jflag[e.order[i]];
    e.arrived.assign(e.M, 0);
    e.cell.assign(e.M, BhCell{}); e.cd.assign(e.M, BhCellD{}); e.sk.assign(e.M, BhCellS{});
    BhTreeView t = e.view();
    for (int i = 0; i < e.n_in; ++i) bh_emit_body(t, e.root.levels, i);
    for (int i = 0; i < e.n_in; ++i) { const int b = e.order[i]; bh_climb_body(t, e.root, i, x[b], y[b], (e.jflag[b] & 1) ? 0.0 : m[b]); }
}

extern "C" {

// One buildTree + computeAccelerations through the engine's algorithm core.
// stats[0..5] = n_in, M, n_internal, interactions, opened, retests
int bh_emul_accelerations(int n, const double* x, const double* y, const double* m,
                          double rcx, double rcy, double rhalf, double theta, double soft2, double G,
                          double* ax, double* ay, int32_t* cntI, int32_t* cntO,
                          uint64_t* key_out, int32_t* depth_out, int32_t* order_out, int64_t* stats) {
    Emul e;
    build(e, n, x, y, m, rcx, rcy, rhalf);
    x = e.xm.data(); y = e.ym.data();
    BhTreeView t = e.view();
    const BhWalkParams w = bh_walk_params(theta, soft2, rhalf);
    int64_t tI = 0, tO = 0, tR = 0;
    for (int si = 0; si < n; ++si) {
        const int b = e.order[si];
        const int self = si < e.n_in ? e.S[si + 1] + si : -1;
        BhWalkResult r = bh_walk_body(t, w, x[b], y[b], self, true);
        if (ax) ax[b] = (m[b] == 0.0) ? NAN : G * (double)r.ax;
        if (ay) ay[b] = (m[b] == 0.0) ? NAN : G * (double)r.ay;
        if (cntI) cntI[b] = r.interactions;
        if (cntO) cntO[b] = r.opened;
        tI += r.interactions; tO += r.opened; tR += r.retests;
    }
    for (int b = 0; b < n; ++b) {
        if (key_out) key_out[b] = e.key_by_body[b];
        if (depth_out) depth_out[b] = -1;
    }
    for (int i = 0; i < e.n_in; ++i) if (depth_out && !(e.jflag[e.order[i]] & 1)) depth_out[e.order[i]] = e.sk[e.S[i + 1] + i].level;
    if (order_out) std::memcpy(order_out, e.order.data(), sizeof(int) * (size_t)n);
    if (stats) { stats[0] = e.n_in - e.n_ghost; stats[1] = e.M; stats[2] = e.S[e.n_in]; stats[3] = tI; stats[4] = tO; stats[5] = tR;
                 stats[6] = e.grid_exact; stats[7] = e.grid_mismatch; }
    return 0;
}

// visitQuads-order export of the tree the core builds (same layout as bh_get_tree)
int bh_emul_tree(int n, const double* x, const double* y, const double* m, double rcx, double rcy, double rhalf,
                 int64_t cap, int64_t* n_cells, double* cx, double* cy, double* h, double* mass, double* comx,
                 double* comy, int32_t* body) {
    Emul e;
    build(e, n, x, y, m, rcx, rcy, rhalf);
    BhHostTree t{e.root, e.n_in, e.M, e.keys.data(), e.order.data(), e.S.data(), e.sk.data(), e.cd.data(), e.jsorted.data()};
    BhCellsOut out;
    out.cap = cap; out.cx = cx; out.cy = cy; out.h = h; out.mass = mass; out.comx = comx; out.comy = comy; out.body = body;
    bh_export_cells(t, out);
    if (n_cells) *n_cells = out.count;
    return 0;
}

// coordinates after one buildTree() (the jitter replay mutates bodies, BH.kt:145-156)
int bh_emul_positions_after_build(int n, const double* x, const double* y, const double* m, double rcx, double rcy,
                                  double rhalf, double* xout, double* yout, int64_t* stats /* n_ghost, unsupported */) {
    Emul e;
    build(e, n, x, y, m, rcx, rcy, rhalf);
    std::memcpy(xout, e.xm.data(), sizeof(double) * (size_t)n);
    std::memcpy(yout, e.ym.data(), sizeof(double) * (size_t)n);
    if (stats) { stats[0] = e.n_ghost; stats[1] = e.unsupported; }
    return 0;
}

}  // extern "C"

// ---- design probe (not a test): size of the union of the cells visited by each group of
// `group` consecutive Morton-sorted bodies vs. the per-body visit counts.  Used to choose
// between per-lane and warp-lockstep traversal.
extern "C" int bh_emul_union_stats(int n, const double* x, const double* y, const double* m, double rcx, double rcy,
                                   double rhalf, double theta, double soft2, int group, double* out /*[4]*/) {
    Emul e;
    build(e, n, x, y, m, rcx, rcy, rhalf);
    BhTreeView t = e.view();
    const double theta2 = theta * theta;
    std::vector<int> stamp(e.M, -1);
    double sumU = 0, sumMax = 0, sumMean = 0, sumUI = 0;
    int groups = 0;
    for (int g0 = 0; g0 + group <= e.n_in; g0 += group, ++groups) {
        int U = 0, UI = 0, vmax = 0; double vsum = 0;
        std::vector<int> istamp;  // interactions union
        for (int si = g0; si < g0 + group; ++si) {
            const int b = e.order[si];
            int v = 0, p = 0;
            while (p < e.M) {
                ++v;
                if (stamp[p] != groups) { stamp[p] = groups; ++U; }
                const bool leafish = e.cell[p].s2 < 0;
                bool acc = leafish || bh_exact_accept(e.cd[p].comx, e.cd[p].comy, x[b], y[b], soft2, theta2, rhalf, e.sk[p].level);
                p = acc ? e.sk[p].skip : p + 1;
            }
            vmax = std::max(vmax, v); vsum += v;
        }
        sumU += U; sumMax += vmax; sumMean += vsum / group; (void)UI; (void)sumUI;
    }
    out[0] = sumU / groups; out[1] = sumMax / groups; out[2] = sumMean / groups; out[3] = groups;
    return 0;
}

// ---- design probe (not a test): L1 sectors touched per warp iteration of the per-lane walk when
// lanes may only run ahead of the slowest lane by `window` preorder positions (window <= 0: free
// running).  out = {iterations per warp, distinct 32B sectors per iteration, distinct 128B lines
// per iteration, sum over iterations of max(floor, sectors) per warp, lane-visits per warp}
extern "C" int bh_emul_window_stats(int n, const double* x, const double* y, const double* m, double rcx, double rcy,
                                    double rhalf, double theta, double soft2, int window, int floor_cost, int stride,
                                    double* out /*[5]*/) {
    Emul e;
    build(e, n, x, y, m, rcx, rcy, rhalf);
    const double theta2 = theta * theta;
    const int G = 32;
    double it = 0, sec = 0, lin = 0, cost = 0, vis = 0;
    int groups = 0;
    for (int g0 = 0; g0 + G <= e.n_in; g0 += G * stride, ++groups) {
        int p[G];
        for (int l = 0; l < G; ++l) p[l] = 0;
        for (;;) {
            int pmin = e.M;
            for (int l = 0; l < G; ++l) pmin = std::min(pmin, p[l]);
            if (pmin >= e.M) break;
            int seen[G], ns = 0, lseen[G], nl = 0;
            for (int l = 0; l < G; ++l) {
                if (p[l] >= e.M) continue;
                if (window > 0 && p[l] >= pmin + window) continue;   // waits
                const int q = p[l];
                bool dup = false;
                for (int k = 0; k < ns; ++k) dup |= seen[k] == q;
                if (!dup) seen[ns++] = q;
                dup = false;
                for (int k = 0; k < nl; ++k) dup |= lseen[k] == (q >> 2);
                if (!dup) lseen[nl++] = q >> 2;
                const int b = e.order[g0 + l];
                const bool leafish = e.cell[q].s2 < 0;
                const bool acc = leafish || bh_exact_accept(e.cd[q].comx, e.cd[q].comy, x[b], y[b], soft2, theta2, rhalf, e.sk[q].level);
                p[l] = acc ? e.sk[q].skip : q + 1;
                vis += 1;
            }
            it += 1; sec += ns; lin += nl; cost += std::max(floor_cost, ns);
        }
    }
    out[0] = it / groups; out[1] = sec / it; out[2] = lin / it; out[3] = cost / groups; out[4] = vis / groups;
    return 0;
}

// ---- design probe (not a test): the warp split into groups of `group` lanes that walk in
// lockstep (one shared position per group: a cell is opened when any un-muted lane of the group
// opens it; a lane that accepts a cell its group opens takes the interaction and is muted until the
// walk leaves that subtree).  Per-body decisions are unchanged.  out = {iterations per warp,
// distinct sectors per iteration, sum over iterations of max(floor, sectors) per warp,
// group-visits per warp, useful lane-visits per warp}
extern "C" int bh_emul_group_stats(int n, const double* x, const double* y, const double* m, double rcx, double rcy,
                                   double rhalf, double theta, double soft2, int group, int floor_cost, int stride,
                                   double* out /*[5]*/) {
    Emul e;
    build(e, n, x, y, m, rcx, rcy, rhalf);
    const double theta2 = theta * theta;
    const int G = 32, NG = G / group;
    double it = 0, sec = 0, cost = 0, gvis = 0, lvis = 0;
    int warps = 0;
    for (int g0 = 0; g0 + G <= e.n_in; g0 += G * stride, ++warps) {
        int p[G], mute[G];
        for (int g = 0; g < NG; ++g) p[g] = 0;
        for (int l = 0; l < G; ++l) mute[l] = 0;
        for (;;) {
            bool any = false;
            int seen[G], ns = 0;
            for (int g = 0; g < NG; ++g) {
                const int q = p[g];
                if (q >= e.M) continue;
                any = true;
                bool dup = false;
                for (int k = 0; k < ns; ++k) dup |= seen[k] == q;
                if (!dup) seen[ns++] = q;
                bool open = false;
                bool acc[G];
                for (int j = 0; j < group; ++j) {
                    const int l = g * group + j;
                    const int b = e.order[g0 + l];
                    const bool leafish = e.cell[q].s2 < 0;
                    acc[j] = leafish || bh_exact_accept(e.cd[q].comx, e.cd[q].comy, x[b], y[b], soft2, theta2, rhalf, e.sk[q].level);
                    if (q >= mute[l]) { lvis += 1; if (!acc[j]) open = true; }
                }
                if (open) {
                    for (int j = 0; j < group; ++j) { const int l = g * group + j; if (q >= mute[l] && acc[j]) mute[l] = e.sk[q].skip; }
                    p[g] = q + 1;
                } else p[g] = e.sk[q].skip;
                gvis += 1;
            }
            if (!any) break;
            it += 1; sec += ns; cost += std::max(floor_cost, ns);
        }
    }
    out[0] = it / warps; out[1] = sec / it; out[2] = cost / warps; out[3] = gvis / warps; out[4] = lvis / warps;
    return 0;
}

// ---- design probe (not a test): distinct 128 B lines per warp iteration of the free-running per-lane
// walk under alternative cell layouts.  layout 0 = preorder (the engine's), 1 = level order (BFS:
// all cells of a level contiguous, siblings adjacent), 2 = level order within 3-level blocks
// (van Emde Boas-like: subtrees of height 3 contiguous).  out = {iterations/warp, sectors/iter, lines/iter}
extern "C" int bh_emul_layout_stats(int n, const double* x, const double* y, const double* m, double rcx, double rcy,
                                    double rhalf, double theta, double soft2, int layout, int stride, double* out /*[3]*/) {
    Emul e;
    build(e, n, x, y, m, rcx, rcy, rhalf);
    const double theta2 = theta * theta;
    std::vector<int> pos(e.M);
    std::iota(pos.begin(), pos.end(), 0);
    if (layout == 1) {
        std::vector<int> idx(e.M);
        std::iota(idx.begin(), idx.end(), 0);
        std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return e.sk[a].level < e.sk[b].level; });
        for (int k = 0; k < e.M; ++k) pos[idx[k]] = k;
    } else if (layout == 2) {
        // key = (level / 3, preorder position of the ancestor at level 3*(level/3), level, preorder)
        std::vector<int> anc(e.M, 0);
        std::vector<int> stack;   // ancestors by level
        std::vector<int> top(64, 0);
        for (int p = 0; p < e.M; ++p) {
            const int L = e.sk[p].level;
            top[L] = p;
            anc[p] = top[(L / 3) * 3];
        }
        std::vector<int> idx(e.M);
        std::iota(idx.begin(), idx.end(), 0);
        std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) {
            const int ba = e.sk[a].level / 3, bb = e.sk[b].level / 3;
            if (ba != bb) return ba < bb;
            if (anc[a] != anc[b]) return anc[a] < anc[b];
            return e.sk[a].level < e.sk[b].level;
        });
        for (int k = 0; k < e.M; ++k) pos[idx[k]] = k;
    }
    const int G = 32;
    double it = 0, sec = 0, lin = 0;
    int groups = 0;
    for (int g0 = 0; g0 + G <= e.n_in; g0 += G * stride, ++groups) {
        int p[G];
        for (int l = 0; l < G; ++l) p[l] = 0;
        for (;;) {
            bool any = false;
            int seen[G], ns = 0, lseen[G], nl = 0;
            for (int l = 0; l < G; ++l) {
                if (p[l] >= e.M) continue;
                any = true;
                const int q = p[l], a = pos[q];
                bool dup = false;
                for (int k = 0; k < ns; ++k) dup |= seen[k] == a;
                if (!dup) seen[ns++] = a;
                dup = false;
                for (int k = 0; k < nl; ++k) dup |= lseen[k] == (a >> 2);
                if (!dup) lseen[nl++] = a >> 2;
                const int b = e.order[g0 + l];
                const bool leafish = e.cell[q].s2 < 0;
                const bool acc = leafish || bh_exact_accept(e.cd[q].comx, e.cd[q].comy, x[b], y[b], soft2, theta2, rhalf, e.sk[q].level);
                p[l] = acc ? e.sk[q].skip : q + 1;
            }
            if (!any) break;
            it += 1; sec += ns; lin += nl;
        }
    }
    out[0] = it / groups; out[1] = sec / it; out[2] = lin / it;
    return 0;
}
