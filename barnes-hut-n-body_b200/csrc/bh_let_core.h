// bh_let_core.h — per-element core of the LOCALLY ESSENTIAL TREE (multi-GPU "domain" mode).
//
// The replicated-tree mode makes every rank build the tree of ALL bodies (an Amdahl term: the
// build of an N-body tree costs the same on every rank however many ranks there are).  Here the
// key space is cut at level ELL (4^ELL level-ELL cells, "codes"): rank r owns a contiguous range
// of codes and builds — with the unchanged single-GPU build — only the tree of the bodies whose
// keys fall into its range.  Every cell of depth >= ELL of that local tree is a cell of the
// reference's global tree, bit for bit (same bodies, same f64 sums).  What a rank's own bodies
// need beyond that is assembled into ONE preorder array, the rank's LET:
//
//   * the TOP TREE (cells of depth < ELL) — replicated, rebuilt on every rank from the all-reduced
//     table of level-ELL summaries (count, mass, centre of mass) with the reference's f64 child
//     sums (bh_climb_from), so its records are the global tree's records;
//   * per non-empty code one BLOCK: the whole level-ELL subtree (own code: spliced in from the
//     local arrays; remote code that some own body may open: imported from its owner), or just the
//     subtree's root record (remote code that no own body can open: the conservative box test
//     bh_let_near says every own body accepts it).
//
// A body's walk over its rank's LET visits exactly the cells, in exactly the order, of its walk
// over the global tree, so forces are bit-identical to a single-GPU run.
//
// Layout trick: the top tree is emitted by the same delta/prefix-sum machinery as the body tree
// (bh_emit_body) over a list of ITEMS instead of bodies: a code holding one body is one item of
// width 1 (a leaf); a code holding >= 2 bodies is a pair of twin items whose keys first differ at
// depth ELL+1, so that the column of internal cells the first twin owns ends with the depth-ELL
// cell = the block root; the first twin's "leaf" is the rest of the block (width B-1), the second
// twin's has width 0.  Positions use W = exclusive scan of the widths where the body tree uses i.
#ifndef BH_LET_CORE_H
#define BH_LET_CORE_H

#include "bh_core.h"

// level-ELL summary of one code, as doubles so that the table can be all-reduced (sum: exactly one
// rank writes a non-zero entry).  count = bodies in the code; (mass, comx, comy) = the exact record
// of its depth-ELL cell (count >= 2) or of its single body; pos / size = position and number of
// cells of the level-ELL subtree in the OWNER's local preorder arrays.
struct BhLetEntry { double count, mass, comx, comy, pos, size; };
struct BhLetBox { double x0, x1, y0, y1; };
// one cell of a block on the wire / during the splice: exact record + skip relative to the block root
struct alignas(32) BhLetWire { double comx, comy, mass; int skip_rel, level; };

enum { BH_LET_SINGLE = 0, BH_LET_TWIN0 = 1, BH_LET_TWIN1 = 2 };

BH_HD uint32_t bh_let_code(uint64_t key, int levels, int ell) { return (uint32_t)(key >> (2 * (levels - ell))); }

// level of the cut for n bodies in total: >= 64 bodies per code, 4 <= ELL <= 10
BH_HD int bh_let_choose_ell(long long n_total, int levels) {
    int ell = 4;
    while (ell < 10 && (1ll << (2 * (ell + 1))) * 64ll <= n_total) ++ell;
    if (ell > levels - 2) ell = levels - 2;
    return ell;
}

// Summary of the code whose FIRST sorted body is i (no-op for any other body).  Run after the climb.
BH_HD void bh_let_summary_body(const BhTreeView& t, int levels, int ell, int i, double x, double y, double m,
                               BhLetEntry* __restrict__ table) {
    const uint64_t k = t.keys[i];
    const int dprev = (i > 0) ? bh_common_levels(t.keys[i - 1], k, levels) : -1;
    if (dprev >= ell) return;
    const int dnext = (i + 1 < t.n_in) ? bh_common_levels(k, t.keys[i + 1], levels) : -1;
    BhLetEntry e;
    if (dnext < ell) {   // alone in its code: a leaf of the global tree (wherever the other ranks' bodies put it)
        e.count = 1.0; e.mass = m; e.comx = x; e.comy = y; e.pos = (double)(t.S[i + 1] + i); e.size = 1.0;
    } else {             // the depth-ELL internal cell is owned by i: delta(i-1) < ELL <= delta(i)
        const int p = t.S[i] + i + (ell - dprev - 1);
        const BhCellS s = t.sk[p];
        const BhCellD d = t.cd[p];
        e.count = (double)s.cnt; e.mass = d.mass; e.comx = d.comx; e.comy = d.comy; e.pos = (double)p; e.size = (double)(s.skip - p);
    }
    table[bh_let_code(k, levels, ell)] = e;
}

// May a body inside `box` OPEN the depth-ELL cell with centre of mass (cx, cy)?  Conservative form of
// BH.kt:223-228: the cell is accepted by every body of the box when s^2 < theta^2 (dmin^2 + soft2)
// holds with a relative margin far above the rounding of either side.
BH_HD bool bh_let_near(const BhLetBox& box, double cx, double cy, double theta2, double soft2, double half, int ell) {
    if (!(box.x0 <= box.x1)) return false;                       // the rank has no bodies
    const double ax = box.x0 - cx, bx = cx - box.x1, ay = box.y0 - cy, by = cy - box.y1;
    const double dx = ax > 0.0 ? ax : (bx > 0.0 ? bx : 0.0);
    const double dy = ay > 0.0 ? ay : (by > 0.0 ? by : 0.0);
    const double dmin2 = dx * dx + dy * dy + soft2;
    const double s2 = bh_side2(half, ell);
    return !(s2 < theta2 * dmin2 * 0.999999);
}

// A rank's REGION: the level-LAMBDA cells ("coarse cells", LAMBDA = bh_let_lambda(ELL)) that hold at
// least one of its own bodies (current positions, strays included) as a bitmap over the 4^LAMBDA coarse
// Morton codes, plus the bounding box of its own bodies outside the root box (targets only).  A Morton
// range is not compact in space (it may straddle a top-level quadrant boundary), so one bounding box per
// rank would make almost every code "near"; the footprint is tight.
struct BhLetRegion { const uint32_t* bits; BhLetBox oob; };
BH_HD int bh_let_lambda(int ell) { return ell >= 3 ? ell - 2 : 1; }
BH_HD uint32_t bh_compact_bits(uint64_t v) {   // bit 2i -> bit i
    uint64_t x = v & 0x5555555555555555ull;
    x = (x | (x >> 1)) & 0x3333333333333333ull;
    x = (x | (x >> 2)) & 0x0F0F0F0F0F0F0F0Full;
    x = (x | (x >> 4)) & 0x00FF00FF00FF00FFull;
    x = (x | (x >> 8)) & 0x0000FFFF0000FFFFull;
    x = (x | (x >> 16)) & 0x00000000FFFFFFFFull;
    return (uint32_t)x;
}
// May one of the rank's own bodies OPEN the depth-ELL cell of code c (centre of mass (cx, cy))?  Only
// coarse cells within R = floor(dmax / w) + 1 cells of c's coarse cell can hold such a body (dmax = the
// largest distance at which BH.kt:228 can fail); each occupied one is tested with its exact square.
BH_HD bool bh_let_near_region(const BhLetRegion& reg, uint32_t c, double cx, double cy, double theta2, double soft2,
                              const BhRoot& root, int ell) {
    if (bh_let_near(reg.oob, cx, cy, theta2, soft2, root.half, ell)) return true;
    const int lam = bh_let_lambda(ell);
    const int G = 1 << lam;
    const double w = ldexp(root.half, 1 - lam);
    const double s2 = bh_side2(root.half, ell);
    int R = G;
    if (theta2 > 0.0) {
        const double dmax2 = s2 / (theta2 * 0.999999) - soft2;
        if (dmax2 < 0.0) return false;                           // softening alone makes every body accept the cell
        const double r = floor(sqrt(dmax2) / w) + 1.0;
        R = r < (double)G ? (int)r : G;
    }
    const uint64_t cc = (uint64_t)c >> (2 * (ell - lam));
    const int ix = (int)bh_compact_bits(cc), iy = (int)bh_compact_bits(cc >> 1);
    const int jx0 = ix - R < 0 ? 0 : ix - R, jx1 = ix + R > G - 1 ? G - 1 : ix + R;
    const int jy0 = iy - R < 0 ? 0 : iy - R, jy1 = iy + R > G - 1 ? G - 1 : iy + R;
    const double bx = root.cx - root.half, by = root.cy - root.half;
    for (int jy = jy0; jy <= jy1; ++jy)
        for (int jx = jx0; jx <= jx1; ++jx) {
            const uint32_t q = (uint32_t)(bh_spread_bits((uint32_t)jx) | (bh_spread_bits((uint32_t)jy) << 1));
            if (!((reg.bits[q >> 5] >> (q & 31u)) & 1u)) continue;
            const BhLetBox box{bx + jx * w, bx + (jx + 1) * w, by + jy * w, by + (jy + 1) * w};
            if (bh_let_near(box, cx, cy, theta2, soft2, root.half, ell)) return true;
        }
    return false;
}

// items and block size of code c in the LET of the rank that owns [c_lo, c_hi) and whose bodies occupy `reg`
BH_HD void bh_let_plan_code(const BhLetEntry& e, uint32_t c, uint32_t c_lo, uint32_t c_hi, const BhLetRegion& reg, double theta2,
                            double soft2, const BhRoot& root, int ell, int* n_items, int* block) {
    if (e.count == 0.0) { *n_items = 0; *block = 0; return; }
    if (e.count == 1.0) { *n_items = 1; *block = 1; return; }
    *n_items = 2;
    const bool mine = c >= c_lo && c < c_hi;
    *block = (mine || bh_let_near_region(reg, c, e.comx, e.comy, theta2, soft2, root, ell)) ? (int)e.size : 1;
}

BH_HD uint64_t bh_let_item_key(uint32_t c, int type, int levels, int ell) {
    const uint64_t k = (uint64_t)c << (2 * (levels - ell));
    return type == BH_LET_TWIN1 ? (k | (1ull << (2 * (levels - ell - 1)))) : k;
}

// ---- the top tree over the items -------------------------------------------------------------------
struct BhLetItems {
    const uint64_t* key;    // [n]   sorted ascending
    const int* type;        // [n]   BH_LET_SINGLE / TWIN0 / TWIN1
    const int* S;           // [n+1] exclusive scan of cnt(j) = max(0, delta(j) - delta(j-1))
    const int* W;           // [n+1] exclusive scan of the widths
    int n;
};

BH_HD int bh_let_item_cnt(const uint64_t* __restrict__ key, int n, int levels, int j) {
    const int dprev = (j > 0) ? bh_common_levels(key[j - 1], key[j], levels) : -1;
    const int dnext = (j + 1 < n) ? bh_common_levels(key[j], key[j + 1], levels) : -1;
    return dnext > dprev ? dnext - dprev : 0;
}

// bh_emit_body over the items: skeletons of the column of internal cells item j owns (for a first
// twin the last one, at depth ELL, is the block root) and of a single's leaf.  Returns the position
// of the item's leaf region (first twin: block root + 1; single: the leaf).
BH_HD int bh_let_emit_item(const BhLetItems& it, BhCellS* __restrict__ sk, int levels, int j) {
    const int n = it.n;
    const uint64_t k = it.key[j];
    const int dprev = (j > 0) ? bh_common_levels(it.key[j - 1], k, levels) : -1;
    const int dnext = (j + 1 < n) ? bh_common_levels(k, it.key[j + 1], levels) : -1;
    const int base = it.S[j] + it.W[j];
    int headParent = -1;
    if (j > 0) {
        const int sh = bh_prefix_shift(levels, dprev);
        const int il = bh_gallop_left(it.key, j, k >> sh, sh);
        const int dl = (il > 0) ? bh_common_levels(it.key[il - 1], it.key[il], levels) : -1;
        headParent = it.S[il] + it.W[il] + (dprev - dl - 1);
    }
    const int ncol = (dnext > dprev) ? (dnext - dprev) : 0;
    int hi = j;
    for (int d = dnext; d > dprev; --d) {
        const int sh = bh_prefix_shift(levels, d);
        hi = bh_gallop_right(it.key, n, hi, k >> sh, sh);
        const int p = base + (d - dprev - 1);
        BhCellS c;
        c.skip = it.S[hi + 1] + it.W[hi + 1];
        c.parent = (d == dprev + 1) ? headParent : (p - 1);
        c.cnt = hi - j + 1;                 // in items (a twin pair counts 2)
        c.level = d;
        sk[p] = c;
    }
    const int lp = base + ncol;
    if (it.type[j] == BH_LET_SINGLE) {
        BhCellS c;
        c.skip = lp + 1;
        c.parent = (ncol > 0) ? (lp - 1) : headParent;
        c.cnt = 1;
        c.level = ((dprev > dnext) ? dprev : dnext) + 1;
        sk[lp] = c;
    }
    return lp;
}

// computeMass (BH.kt:173-202) of the top tree: a single writes its body's leaf record, a first twin
// the exact record of its block root (both from the replicated table), then both climb with the
// arrive-counter protocol of the body tree — the last arriver sums the children 0..3 in order.
BH_HD void bh_let_climb_item(const BhTreeView& t, const BhRoot& root, const BhLetItems& it, const BhLetEntry* __restrict__ table,
                             int levels, int ell, int j, int lp) {
    const int type = it.type[j];
    if (type == BH_LET_TWIN1) return;
    const BhLetEntry e = table[bh_let_code(it.key[j], levels, ell)];
    if (type == BH_LET_SINGLE) {
        const BhCellS s = t.sk[lp];
        bh_write_cell(t, lp, e.comx, e.comy, e.mass, s.skip, s.level, true, root.half);
        bh_climb_from(t, root, it.key[j], s, 1);
    } else {
        const int p = lp - 1;               // the depth-ELL cell: last of the column
        const BhCellS s = t.sk[p];
        bh_write_cell(t, p, e.comx, e.comy, e.mass, s.skip, s.level, false, root.half);
        bh_climb_from(t, root, it.key[j], s, 2);
    }
}

// ---- blocks ------------------------------------------------------------------------------------------
// cell `p` of the owner's local arrays, `root_pos` = position of its block root there
BH_HD BhLetWire bh_let_wire(const BhCellD* __restrict__ cd, const BhCellS* __restrict__ sk, int p, int root_pos) {
    const BhCellD d = cd[p];
    const BhCellS s = sk[p];
    BhLetWire w;
    w.comx = d.comx; w.comy = d.comy; w.mass = d.mass; w.skip_rel = s.skip - root_pos; w.level = s.level;
    return w;
}
// the j-th cell of a block (j >= 1) whose root sits at `dst` in the LET
BH_HD void bh_let_place(const BhTreeView& t, const BhLetWire& w, int dst, int j, double half) {
    const int p = dst + j;
    const int skip = dst + w.skip_rel;
    BhCellS s; s.skip = skip; s.parent = -1; s.cnt = 0; s.level = w.level;
    t.sk[p] = s;
    bh_write_cell(t, p, w.comx, w.comy, w.mass, skip, w.level, w.skip_rel == j + 1, half);
}

#endif  // BH_LET_CORE_H
