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
#include "../../barnes-hut-n-body_b200/csrc/bh_let_core.h"

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
                  double rcx, double rcy, double rhalf, int ell = -1, uint32_t c_lo = 0, uint32_t c_hi = 0,
                  const int* list_pos = nullptr) {
    e.root = BhRoot{rcx, rcy, rhalf, bh_key_levels(rhalf)};
    e.n = n;
    e.key_by_body.resize(n);
    const BhGrid grid = bh_make_grid(e.root);
    for (int b = 0; b < n; ++b) {
        e.key_by_body[b] = bh_root_contains(e.root, x[b], y[b]) ? bh_morton_key(e.root, x[b], y[b]) : BH_KEY_NOT_IN_TREE;
        if (ell >= 0 && e.key_by_body[b] != BH_KEY_NOT_IN_TREE) {   // LET mode: only the bodies of the rank's code range
            const uint32_t c = bh_let_code(e.key_by_body[b], e.root.levels, ell);
            if (c < c_lo || c >= c_hi) e.key_by_body[b] = BH_KEY_NOT_IN_TREE;
        }
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
        if (list_pos) ident.assign(list_pos, list_pos + n);
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

// ---- design probe (not a test): the warp-cooperative walk (k_walk_coop).  A group of `group`
// Morton-adjacent bodies shares one traversal: a work stack of (sibling range [q, e), lane mask)
// entries, `width` entries popped per round (one per lane); the lane that owns an entry classifies
// cell q for the lanes in its mask: Acc = lanes that accept (BH.kt:228), Op = the rest; Acc != 0
// emits one interaction-list entry (cell, Acc); Op != 0 pushes the children range.  Per-body
// decisions are unchanged.  out = {cells processed per group, list entries per group, entries
// whose decision is not uniform over the mask, sum popc(Acc) per group, rounds per group, max stack
// depth, entries the bounding-box test cannot classify, mean popc(mask) of list entries}
extern "C" int bh_emul_coop_stats(int n, const double* x, const double* y, const double* m, double rcx, double rcy,
                                  double rhalf, double theta, double soft2, int group, int width, int stride,
                                  double* out /*[8]*/) {
    Emul e;
    build(e, n, x, y, m, rcx, rcy, rhalf);
    const double theta2 = theta * theta;
    struct Ent { int q, end; uint32_t mask; };
    double cells = 0, entries = 0, mixed = 0, pops = 0, rounds = 0, bbox_fail = 0, msum = 0;
    int groups = 0; size_t maxdepth = 0;
    std::vector<Ent> st, popped;
    for (int g0 = 0; g0 + group <= e.n_in; g0 += group * stride, ++groups) {
        double bx0 = 1e300, bx1 = -1e300, by0 = 1e300, by1 = -1e300;
        for (int l = 0; l < group; ++l) {
            const int b = e.order[g0 + l];
            bx0 = std::min(bx0, x[b]); bx1 = std::max(bx1, x[b]); by0 = std::min(by0, y[b]); by1 = std::max(by1, y[b]);
        }
        st.clear();
        st.push_back(Ent{0, e.M, group >= 32 ? 0xffffffffu : ((1u << group) - 1u)});
        while (!st.empty()) {
            const int k = (int)std::min<size_t>(st.size(), (size_t)width);
            popped.assign(st.end() - k, st.end());
            st.resize(st.size() - k);
            rounds += 1;
            for (int j = k - 1; j >= 0; --j) {
                const Ent en = popped[j];
                const int q = en.q;
                const int skip = e.sk[q].skip;
                cells += 1;
                const bool leafish = e.cell[q].s2 < 0;
                uint32_t acc = 0;
                for (int l = 0; l < group; ++l) {
                    if (!((en.mask >> l) & 1u)) continue;
                    const int b = e.order[g0 + l];
                    const bool a = leafish || bh_exact_accept(e.cd[q].comx, e.cd[q].comy, x[b], y[b], soft2, theta2, rhalf, e.sk[q].level);
                    if (a) acc |= 1u << l;
                }
                const uint32_t op = en.mask & ~acc;
                if (acc) { entries += 1; pops += __builtin_popcount(acc); msum += __builtin_popcount(acc); }
                if (acc && op) mixed += 1;
                if (!leafish) {
                    // bounding-box classification: min / max distance from the COM to the group's box
                    const double cx = e.cd[q].comx, cy = e.cd[q].comy;
                    const double dxn = std::max(0.0, std::max(bx0 - cx, cx - bx1)), dyn = std::max(0.0, std::max(by0 - cy, cy - by1));
                    const double dxf = std::max(cx - bx0, bx1 - cx), dyf = std::max(cy - by0, by1 - cy);
                    double h = rhalf; for (int d = 0; d < e.sk[q].level; ++d) h /= 2.0;
                    const double s2 = 4.0 * h * h;
                    const bool all_acc = s2 < theta2 * (dxn * dxn + dyn * dyn + soft2) * (1 - 1e-4);
                    const bool all_open = s2 > theta2 * (dxf * dxf + dyf * dyf + soft2) * (1 + 1e-4);
                    if (!all_acc && !all_open) bbox_fail += 1;
                }
                if (skip < en.end) st.push_back(Ent{skip, en.end, en.mask});
                if (op && skip > q + 1) st.push_back(Ent{q + 1, skip, op});
            }
            maxdepth = std::max(maxdepth, st.size());
        }
    }
    out[0] = cells / groups; out[1] = entries / groups; out[2] = mixed / groups; out[3] = pops / groups;
    out[4] = rounds / groups; out[5] = (double)maxdepth; out[6] = bbox_fail / groups; out[7] = msum / std::max(1.0, entries);
    return 0;
}

// ---- locally essential trees (bh_let_core.h): P emulated ranks ------------------------------------------
// Home ranks and code ranges are fixed from the positions (x0, y0) of the last re-homing (NULL: the
// current ones); every rank then builds its local tree from the CURRENT positions of the bodies whose
// keys fall into its code range (its own minus the strays, plus the guests), the level-ELL summaries
// are merged into the replicated table, and every rank assembles its LET (top tree + own blocks +
// imported blocks + far roots) and walks its own bodies over it.  Outputs are per body (list order),
// to be compared bit for bit with bh_emul_accelerations.  stats = {sum of LET cells, imported cells,
// strays, a guest sits in a jitter cluster, max LET cells of a rank, ELL, sum of local-tree cells,
// max items of a rank}
extern "C" int bh_emul_let(int n, const double* x0, const double* y0, const double* x, const double* y, const double* m,
                           double rcx, double rcy, double rhalf, double theta, double soft2, double G, int P, int ell,
                           double* ax, double* ay, int32_t* cntI, int32_t* cntO, int64_t* stats) {
    const BhRoot root{rcx, rcy, rhalf, bh_key_levels(rhalf)};
    const int L = root.levels;
    if (ell <= 0) ell = bh_let_choose_ell(n, L);
    if (ell > L - 2 || ell < 1) return -1;
    const uint32_t ncodes = 1u << (2 * ell);
    if (!x0) { x0 = x; y0 = y; }
    // ---- re-homing: global order of (x0, y0), slices cut at code boundaries
    std::vector<uint64_t> k0(n);
    for (int b = 0; b < n; ++b) k0[b] = bh_root_contains(root, x0[b], y0[b]) ? bh_morton_key(root, x0[b], y0[b]) : BH_KEY_NOT_IN_TREE;
    std::vector<int> ord(n);
    std::iota(ord.begin(), ord.end(), 0);
    std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return k0[a] < k0[b]; });
    int n_in0 = 0;
    for (int b = 0; b < n; ++b) n_in0 += k0[b] != BH_KEY_NOT_IN_TREE;
    std::vector<uint32_t> cs(P + 1, ncodes);
    std::vector<int> cut(P + 1, n);
    cs[0] = 0; cut[0] = 0;
    for (int r = 1; r < P; ++r) {
        int i = (int)((int64_t)n_in0 * r / P);
        i = std::max(i, cut[r - 1]);
        while (i > 0 && i < n_in0 && bh_let_code(k0[ord[i]], L, ell) == bh_let_code(k0[ord[i - 1]], L, ell)) ++i;
        cut[r] = std::min(i, n_in0);
        cs[r] = cut[r] < n_in0 ? bh_let_code(k0[ord[cut[r]]], L, ell) : ncodes;
    }
    std::vector<int> home(n);
    for (int r = 0; r < P; ++r) for (int i = cut[r]; i < cut[r + 1]; ++i) home[ord[i]] = r;   // out-of-box bodies: last rank
    // ---- per rank: local bodies = own (strays keep their slot as targets only) + guests
    struct Rank { std::vector<int> body; int n_own = 0; Emul e; BhLetBox box; std::vector<uint32_t> bits; std::vector<double> lx, ly, lm; std::vector<int> lpos; };
    std::vector<Rank> R(P);
    std::vector<uint64_t> kc(n);
    int64_t n_stray = 0;
    for (int b = 0; b < n; ++b) kc[b] = bh_root_contains(root, x[b], y[b]) ? bh_morton_key(root, x[b], y[b]) : BH_KEY_NOT_IN_TREE;
    for (int b = 0; b < n; ++b) R[home[b]].body.push_back(b);
    for (int r = 0; r < P; ++r) R[r].n_own = (int)R[r].body.size();
    for (int b = 0; b < n; ++b) {
        if (kc[b] == BH_KEY_NOT_IN_TREE) continue;
        const uint32_t c = bh_let_code(kc[b], L, ell);
        const int h = home[b];
        if (c >= cs[h] && c < cs[h + 1]) continue;
        ++n_stray;
        for (int r = 0; r < P; ++r) if (c >= cs[r] && c < cs[r + 1]) R[r].body.push_back(b);   // guest of the range's owner
    }
    std::vector<BhLetEntry> table(ncodes, BhLetEntry{0, 0, 0, 0, 0, 0});
    int64_t guest_jitter = 0, sumLocalM = 0;
    for (int r = 0; r < P; ++r) {
        Rank& k = R[r];
        const int nl = (int)k.body.size();
        k.lx.resize(nl); k.ly.resize(nl); k.lm.resize(nl); k.lpos.resize(nl);
        k.box = BhLetBox{1e300, -1e300, 1e300, -1e300};   // own bodies outside the root box
        const int lam = bh_let_lambda(ell);
        k.bits.assign(((size_t)1 << (2 * lam)) / 32 + 1, 0u);
        for (int j = 0; j < nl; ++j) {
            const int b = k.body[j];
            k.lx[j] = x[b]; k.ly[j] = y[b]; k.lm[j] = m[b]; k.lpos[j] = b;
            if (j < k.n_own) {
                if (kc[b] == BH_KEY_NOT_IN_TREE) {
                    k.box.x0 = std::min(k.box.x0, x[b]); k.box.x1 = std::max(k.box.x1, x[b]);
                    k.box.y0 = std::min(k.box.y0, y[b]); k.box.y1 = std::max(k.box.y1, y[b]);
                } else {
                    const uint32_t q = bh_let_code(kc[b], L, lam);
                    k.bits[q >> 5] |= 1u << (q & 31u);
                }
            }
        }
        build(k.e, nl, k.lx.data(), k.ly.data(), k.lm.data(), rcx, rcy, rhalf, ell, cs[r], cs[r + 1], k.lpos.data());
        for (int j = k.n_own; j < nl; ++j) if (k.e.jflag[j] || k.e.xm[j] != k.lx[j] || k.e.ym[j] != k.ly[j]) guest_jitter = 1;
        sumLocalM += k.e.M;
        BhTreeView t = k.e.view();
        for (int i = 0; i < k.e.n_in; ++i) {
            const int j = k.e.order[i];
            bh_let_summary_body(t, L, ell, i, k.e.xm[j], k.e.ym[j], (k.e.jflag[j] & 1) ? 0.0 : k.lm[j], table.data());
        }
    }
    // ---- per rank: LET
    const BhWalkParams w = bh_walk_params(theta, soft2, rhalf);
    const double theta2 = theta * theta;
    int64_t sumM = 0, sumImp = 0, maxM = 0, maxItems = 0;
    for (int r = 0; r < P; ++r) {
        Rank& k = R[r];
        std::vector<int> first(ncodes + 1, 0), blk(ncodes, 0);
        for (uint32_t c = 0; c < ncodes; ++c) {
            int ni, B;
            bh_let_plan_code(table[c], c, cs[r], cs[r + 1], BhLetRegion{k.bits.data(), k.box}, theta2, soft2, root, ell, &ni, &B);
            first[c + 1] = first[c] + ni; blk[c] = B;
        }
        const int nit = first[ncodes];
        maxItems = std::max<int64_t>(maxItems, nit);
        std::vector<uint64_t> ikey(nit); std::vector<int> itype(nit), iw(nit), iS(nit + 1, 0), iW(nit + 1, 0), ilp(nit);
        for (uint32_t c = 0; c < ncodes; ++c) {
            const int j = first[c], ni = first[c + 1] - first[c];
            if (ni == 1) { ikey[j] = bh_let_item_key(c, BH_LET_SINGLE, L, ell); itype[j] = BH_LET_SINGLE; iw[j] = 1; }
            if (ni == 2) {
                ikey[j] = bh_let_item_key(c, BH_LET_TWIN0, L, ell); itype[j] = BH_LET_TWIN0; iw[j] = blk[c] - 1;
                ikey[j + 1] = bh_let_item_key(c, BH_LET_TWIN1, L, ell); itype[j + 1] = BH_LET_TWIN1; iw[j + 1] = 0;
            }
        }
        for (int j = 0; j < nit; ++j) { iS[j + 1] = iS[j] + bh_let_item_cnt(ikey.data(), nit, L, j); iW[j + 1] = iW[j] + iw[j]; }
        const int M = iS[nit] + iW[nit];
        sumM += M; maxM = std::max<int64_t>(maxM, M);
        std::vector<BhCell> cell(M + 1); std::vector<BhCellD> cd(M + 1); std::vector<BhCellS> sk(M + 1); std::vector<int> arrived(M + 1, 0);
        BhTreeView t{};
        t.cell = cell.data(); t.cd = cd.data(); t.sk = sk.data(); t.arrived = arrived.data(); t.n_in = nit; t.M = M;
        const BhLetItems it{ikey.data(), itype.data(), iS.data(), iW.data(), nit};
        for (int j = 0; j < nit; ++j) ilp[j] = bh_let_emit_item(it, sk.data(), L, j);
        // blocks: own (splice) and imported (through the wire format) — every cell but the root
        std::vector<int> dst(ncodes, -1);
        for (uint32_t c = 0; c < ncodes; ++c) {
            const int ni = first[c + 1] - first[c];
            if (ni == 0) continue;
            dst[c] = ni == 1 ? ilp[first[c]] : ilp[first[c]] - 1;
            if (ni == 2 && blk[c] > 1) {
                int owner = 0;
                while (!(c >= cs[owner] && c < cs[owner + 1])) ++owner;
                Emul& oe = R[owner].e;
                const int rp = (int)table[c].pos;
                for (int j = 1; j < blk[c]; ++j) bh_let_place(t, bh_let_wire(oe.cd.data(), oe.sk.data(), rp + j, rp), dst[c], j, rhalf);
                if (owner != r) sumImp += blk[c] - 1;
            }
        }
        for (int j = 0; j < nit; ++j) bh_let_climb_item(t, root, it, table.data(), L, ell, j, ilp[j]);
        bh_write_terminal_cell(t);
        // walk the own bodies (targets), self = LET position of the body's own leaf
        std::vector<int> sorted_of((size_t)k.body.size(), -1);
        for (int i = 0; i < k.e.n_in; ++i) sorted_of[k.e.order[i]] = i;
        for (int j = 0; j < k.n_own; ++j) {
            const int b = k.body[j];
            int self = -1;
            const uint64_t kk = k.e.key_by_body[j];
            if (kk != BH_KEY_NOT_IN_TREE) {
                const int si = sorted_of[j];   // sorted position of local body j
                const uint32_t c = bh_let_code(kk, L, ell);
                self = dst[c] + (k.e.S[si + 1] + si - (int)table[c].pos);
            }
            const BhWalkResult res = bh_walk_body(t, w, k.e.xm[j], k.e.ym[j], self, true);
            if (ax) ax[b] = (m[b] == 0.0) ? NAN : G * (double)res.ax;
            if (ay) ay[b] = (m[b] == 0.0) ? NAN : G * (double)res.ay;
            if (cntI) cntI[b] = res.interactions;
            if (cntO) cntO[b] = res.opened;
        }
    }
    if (stats) { stats[0] = sumM; stats[1] = sumImp; stats[2] = n_stray; stats[3] = guest_jitter; stats[4] = maxM; stats[5] = ell;
                 stats[6] = sumLocalM; stats[7] = maxItems; }
    return 0;
}

// ---- property check of bh_let_near_region: whenever SOME body opens the depth-ELL cell of a candidate
// (exact f64 test, BH.kt:223-228), the rank's region test must say "near".  Candidates are points; the
// cell is the level-ELL cell containing the point and the point plays its centre of mass.  Returns the
// number of violations; out[0] = candidates some body opens, out[1] = candidates reported near.
extern "C" int bh_emul_let_near_check(int n, const double* x, const double* y, double rcx, double rcy, double rhalf,
                                      double theta, double soft2, int ell, int ncand, const double* cx, const double* cy,
                                      int64_t* out) {
    const BhRoot root{rcx, rcy, rhalf, bh_key_levels(rhalf)};
    const int L = root.levels, lam = bh_let_lambda(ell);
    std::vector<uint32_t> bits(((size_t)1 << (2 * lam)) / 32 + 1, 0u);
    BhLetBox oob{1e300, -1e300, 1e300, -1e300};
    for (int b = 0; b < n; ++b) {
        if (!bh_root_contains(root, x[b], y[b])) {
            oob.x0 = std::min(oob.x0, x[b]); oob.x1 = std::max(oob.x1, x[b]);
            oob.y0 = std::min(oob.y0, y[b]); oob.y1 = std::max(oob.y1, y[b]);
        } else {
            const uint32_t q = bh_let_code(bh_morton_key(root, x[b], y[b]), L, lam);
            bits[q >> 5] |= 1u << (q & 31u);
        }
    }
    const BhLetRegion reg{bits.data(), oob};
    const double theta2 = theta * theta;
    int violations = 0;
    int64_t opened = 0, near = 0;
    for (int k = 0; k < ncand; ++k) {
        if (!bh_root_contains(root, cx[k], cy[k])) continue;
        const uint32_t c = bh_let_code(bh_morton_key(root, cx[k], cy[k]), L, ell);
        bool some_opens = false;
        for (int b = 0; b < n && !some_opens; ++b)
            some_opens = !bh_exact_accept(cx[k], cy[k], x[b], y[b], soft2, theta2, rhalf, ell);
        const bool is_near = bh_let_near_region(reg, c, cx[k], cy[k], theta2, soft2, root, ell);
        opened += some_opens; near += is_near;
        if (some_opens && !is_near) ++violations;
    }
    if (out) { out[0] = opened; out[1] = near; }
    return violations;
}
