// bh_core.h — per-element algorithm core of the B200 Barnes–Hut engine.
//
// Everything here is `__host__ __device__ inline`: the CUDA kernels in bh_engine.cu call
// these functions per thread, and tests/emul/bh_emul.cpp runs the very same functions in
// serial loops on the CPU (test infrastructure only) so the layout/search/criterion
// logic can be checked against the oracle without a GPU.
//
// Citations "BH.kt:a-b" are /root/reference/src/main/kotlin/BarnesHutAlg.kt.
//
// ---------------------------------------------------------------------------------------
// Tree representation.  Bodies that pass the root contains() test
// (BH.kt:126) are sorted by a 2L-bit Morton key whose 2-bit digits are the child indices
// the reference's insertIntoChild would pick (BH.kt:153-155).  With
//     delta(i) = number of leading digits shared by key[i] and key[i+1]   (delta(-1) = delta(n-1) = -1)
// the reference's (uncompressed, capacity-1) quadtree is fully determined:
//   * cell (prefix of key[i], depth d) is INTERNAL  <=>  delta(i-1) < d <= delta(i)
//     for its leftmost key i   (it holds >= 2 bodies);
//   * body i sits in a leaf of depth max(delta(i-1), delta(i)) + 1.
// Cells are stored in DFS PREORDER (children in digit order = key order) — empty leaves
// are implicit.  With cnt(i) = max(0, delta(i) - delta(i-1)) and S = exclusive scan of cnt:
//     pos(internal (i,d)) = S[i] + i + (d - delta(i-1) - 1)
//     pos(leaf i)         = S[i+1] + i
//     skip(internal (i,d)) = S[hi+1] + hi + 1,  hi = last key sharing d digits with key[i]
// so the whole layout needs one prefix sum and no allocation.  `skip` is the preorder
// position after the cell's subtree: a stackless walk goes to `skip` on accept and to
// `pos+1` on open.
// ---------------------------------------------------------------------------------------
#ifndef BH_CORE_H
#define BH_CORE_H

#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define BH_HD __host__ __device__ __forceinline__
#else
#define BH_HD inline
#endif

#define BH_MAX_LEVELS 31
#define BH_KEY_NOT_IN_TREE 0xFFFFFFFFFFFFFFFFull
// relative half-width of the band in which the FP32 opening test is re-done in FP64
#define BH_GUARD_BAND 1.0e-5f

// f64 arithmetic that must round exactly like the reference's JVM doubles: no FMA
// contraction.  Device: explicit round-to-nearest intrinsics.  Host: plain operators
// (the emulation is compiled with -ffp-contract=off).
#if defined(__CUDA_ARCH__)
#define BH_DMUL(a, b) __dmul_rn((a), (b))
#define BH_DADD(a, b) __dadd_rn((a), (b))
#define BH_DSUB(a, b) __dsub_rn((a), (b))
#define BH_DDIV(a, b) __ddiv_rn((a), (b))
#else
#define BH_DMUL(a, b) ((a) * (b))
#define BH_DADD(a, b) ((a) + (b))
#define BH_DSUB(a, b) ((a) - (b))
#define BH_DDIV(a, b) ((a) / (b))
#endif

struct BhRoot {
    double cx, cy, half;  // BH.kt:360-361
    int levels;           // digits per key = first depth whose half-side is < 1e-3 (BH.kt:146)
};

// First depth d with half/2^d < 1e-3: a cell at that depth jitters its bodies when it
// has to subdivide (BH.kt:146), so a jitter-free tree never has an internal cell there and
// `levels` digits identify every leaf.
BH_HD int bh_key_levels(double half) {
    int d = 0;
    double h = half;
    while (!(h < 1e-3) && d < BH_MAX_LEVELS) { h = h / 2.0; ++d; }
    return d;
}

// BH.kt:61-62
BH_HD bool bh_root_contains(const BhRoot& r, double x, double y) {
    return x >= BH_DSUB(r.cx, r.half) && x < BH_DADD(r.cx, r.half) &&
           y >= BH_DSUB(r.cy, r.half) && y < BH_DADD(r.cy, r.half);
}

// Literal descent: the digits insertIntoChild (BH.kt:153-155) picks with the child
// centres of Quad.child (BH.kt:73-80), `levels` times.  Digit of depth k lands in bits
// [2(levels-1-k)+1 : 2(levels-1-k)].
BH_HD uint64_t bh_morton_key(const BhRoot& r, double x, double y) {
    double cx = r.cx, cy = r.cy, h = r.half;
    uint64_t key = 0;
    for (int l = 0; l < r.levels; ++l) {
        const double hh = h / 2.0;
        const int ix = (x < cx) ? 0 : 1;
        const int iy = (y < cy) ? 0 : 2;
        key = (key << 2) | (uint64_t)(ix + iy);
        cx = ix ? BH_DADD(cx, hh) : BH_DSUB(cx, hh);
        cy = iy ? BH_DADD(cy, hh) : BH_DSUB(cy, hh);
        h = hh;
    }
    return key;
}

// Geometry of the depth-d cell on the path `key` (same arithmetic as Quad.child).
BH_HD void bh_cell_geometry(const BhRoot& r, uint64_t key, int d, double* ocx, double* ocy, double* oh) {
    double cx = r.cx, cy = r.cy, h = r.half;
    for (int l = 0; l < d; ++l) {
        const double hh = h / 2.0;
        const int dig = (int)((key >> (2 * (r.levels - 1 - l))) & 3ull);
        cx = (dig & 1) ? BH_DADD(cx, hh) : BH_DSUB(cx, hh);
        cy = (dig & 2) ? BH_DADD(cy, hh) : BH_DSUB(cy, hh);
        h = hh;
    }
    *ocx = cx; *ocy = cy; *oh = h;
}

BH_HD int bh_clz64(uint64_t v) {
#if defined(__CUDA_ARCH__)
    return __clzll((long long)v);
#else
    return v ? __builtin_clzll(v) : 64;
#endif
}

// number of leading digits (levels) two keys share; == levels when the keys are equal
BH_HD int bh_common_levels(uint64_t a, uint64_t b, int levels) {
    const uint64_t x = a ^ b;
    if (x == 0) return levels;
    const int hb = 63 - bh_clz64(x);          // highest differing bit
    return (2 * levels - 1 - hb) >> 1;
}

// (side of a depth-d cell)^2 exactly as BH.kt:226: h_d = half/2^d (exact), s = h*2, s2 = s*s
BH_HD double bh_side2(double half, int d) {
    const double side = ldexp(half, 1 - d);
    return BH_DMUL(side, side);
}

// f64 -> (hi, lo) float pair: hi = fl32(v), lo = fl32(v - hi).  Differences of two such
// pairs, (ah-bh)+(al-bl), keep ~2^-24 RELATIVE accuracy in the difference itself, which is
// what the force needs (absolute FP32 coordinates do not: SURVEY.md §7.3-H3).
BH_HD void bh_split(double v, float* hi, float* lo) {
    const float h = (float)v;
    *hi = h;
    *lo = (float)(v - (double)h);
}

// last index j >= from (from is known to be inside) whose key shares the prefix
// (key >> sh) == pref; keys sorted ascending, n = number of in-tree keys
BH_HD int bh_gallop_right(const uint64_t* __restrict__ keys, int n, int from, uint64_t pref, int sh) {
    int j = from;
    int step = 1;
    while (j + step < n && (keys[j + step] >> sh) == pref) { j += step; step <<= 1; }
    int out = (j + step < n) ? (j + step) : n;   // first index known to be outside (or n)
    while (out - j > 1) {
        const int mid = j + ((out - j) >> 1);
        if ((keys[mid] >> sh) == pref) j = mid; else out = mid;
    }
    return j;
}

// first index j <= from whose key shares the prefix
BH_HD int bh_gallop_left(const uint64_t* __restrict__ keys, int from, uint64_t pref, int sh) {
    int j = from;
    int step = 1;
    while (j - step >= 0 && (keys[j - step] >> sh) == pref) { j -= step; step <<= 1; }
    int out = (j - step >= 0) ? (j - step) : -1;  // last index known to be outside (or -1)
    while (j - out > 1) {
        const int mid = out + ((j - out) >> 1);
        if ((keys[mid] >> sh) == pref) j = mid; else out = mid;
    }
    return j;
}

// shift that isolates the first d digits of a key
BH_HD int bh_prefix_shift(int levels, int d) { return 2 * (levels - d); }

// ---- hot cell record read by the walk: two 16-byte vectors per cell ---------------------
//   A = (comX_hi, comY_hi, mass, s2)   s2 = (cell side)^2 as float; -1 for a leaf or a
//                                       zero-mass cell (always "accepted": BH.kt:216-221)
//   B = (comX_lo, comY_lo, skip, level) skip/level are int bit patterns
struct BhCellA { float xh, yh, m, s2; };
struct BhCellB { float xl, yl; int skip; int level; };

// The reference's f64 opening test, BH.kt:223-228, bit-for-bit.
BH_HD bool bh_exact_accept(double comx, double comy, double x, double y, double soft2, double theta2,
                           double half, int level) {
    const double dx = BH_DSUB(comx, x);
    const double dy = BH_DSUB(comy, y);
    const double dist2 = BH_DADD(BH_DADD(BH_DMUL(dx, dx), BH_DMUL(dy, dy)), soft2);
    const double s2 = bh_side2(half, level);
    return s2 < BH_DMUL(theta2, dist2);
}

// ---- SoA view of the preorder cell arrays --------------------------------------------------
struct BhTreeView {
    const uint64_t* keys;   // sorted keys of the in-tree bodies            [n_in]
    const int*      order;  // body (home) index at each sorted position     [n]
    const int*      S;      // exclusive scan of cnt(i), S[n_in] = #internal [n_in+1]
    BhCellA* A;             // hot record, first half                        [M]
    BhCellB* B;             // hot record, second half                       [M]
    double*  comx;          // f64 centre of mass / mass, bit-identical to   [M]
    double*  comy;          //   BHTree.computeMass (BH.kt:173-202)
    double*  cmass;
    int*     skip;          // preorder position after the subtree           [M]
    int*     parent;        // preorder position of the parent, -1 for root  [M]
    int*     cnt;           // bodies below the cell                         [M]
    int*     arrived;       // climb counters, zero before the climb         [M]
    signed char* lvl;       // depth of the cell                             [M]
    int n_in;               // bodies in the tree
    int M;                  // cells = n_in + #internal
};

#if defined(__CUDA_ARCH__)
#define BH_ATOMIC_ADD_INT(p, v) atomicAdd((p), (v))
#define BH_FENCE() __threadfence()
#define BH_LD_D(p) __ldcg(p)
#define BH_LD_I(p) __ldcg(p)
#define BH_RSQRTF(v) rsqrtf(v)
#else
static inline int bh_host_fetch_add(int* p, int v) { const int o = *p; *p = o + v; return o; }
#define BH_ATOMIC_ADD_INT(p, v) bh_host_fetch_add((p), (v))
#define BH_FENCE() ((void)0)
#define BH_LD_D(p) (*(p))
#define BH_LD_I(p) (*(p))
#define BH_RSQRTF(v) (1.0f / sqrtf(v))
#endif

// Skeleton of everything body i "owns" in the preorder array: the column of internal
// cells whose leftmost key is key[i], and the leaf of body i.  Pure function of the sorted
// keys and S (no atomics); one thread per in-tree body.
BH_HD void bh_emit_body(const BhTreeView& t, int levels, int i) {
    const int n = t.n_in;
    const uint64_t k = t.keys[i];
    const int dprev = (i > 0) ? bh_common_levels(t.keys[i - 1], k, levels) : -1;
    const int dnext = (i + 1 < n) ? bh_common_levels(k, t.keys[i + 1], levels) : -1;
    const int base = t.S[i] + i;
    // parent of the first entry of this group: the depth-dprev cell holding key[i-1] and key[i]
    int headParent = -1;
    if (i > 0) {
        const int sh = bh_prefix_shift(levels, dprev);
        const int il = bh_gallop_left(t.keys, i, k >> sh, sh);
        const int dl = (il > 0) ? bh_common_levels(t.keys[il - 1], t.keys[il], levels) : -1;
        headParent = t.S[il] + il + (dprev - dl - 1);
    }
    const int ncol = (dnext > dprev) ? (dnext - dprev) : 0;
    int hi = i;
    for (int d = dnext; d > dprev; --d) {   // deepest first: hi only grows
        const int sh = bh_prefix_shift(levels, d);
        hi = bh_gallop_right(t.keys, n, hi, k >> sh, sh);
        const int p = base + (d - dprev - 1);
        t.skip[p] = t.S[hi + 1] + hi + 1;
        t.cnt[p] = hi - i + 1;
        t.parent[p] = (d == dprev + 1) ? headParent : (p - 1);
        t.lvl[p] = (signed char)d;
    }
    const int lp = base + ncol;
    t.skip[lp] = lp + 1;
    t.cnt[lp] = 1;
    t.parent[lp] = (ncol > 0) ? (lp - 1) : headParent;
    t.lvl[lp] = (signed char)(((dprev > dnext) ? dprev : dnext) + 1);
}

BH_HD void bh_write_cell(const BhTreeView& t, int p, double cx, double cy, double m, int level, bool leaf, double half) {
    t.comx[p] = cx; t.comy[p] = cy; t.cmass[p] = m;
    BhCellA a; BhCellB b;
    bh_split(cx, &a.xh, &b.xl);
    bh_split(cy, &a.yh, &b.yl);
    a.m = (float)m;
    // leaves and zero-mass cells are never opened (BH.kt:216-221): s2 = -1 always passes
    a.s2 = (leaf || m == 0.0) ? -1.0f : (float)bh_side2(half, level);
    b.skip = t.skip[p];
    b.level = level;
    t.A[p] = a; t.B[p] = b;
}

// computeMass (BH.kt:173-202) bottom-up: the thread of body i writes its leaf, then climbs;
// at each parent it adds the body count of the finished child and continues only if it
// completed the parent (every child done) — the last arriver sums the children in child
// order 0..3 (= preorder order), with the reference's exact f64 expression order.
BH_HD void bh_climb_body(const BhTreeView& t, const BhRoot& root, int i, double x, double y, double m) {
    int p = t.S[i + 1] + i;
    bh_write_cell(t, p, x, y, m, t.lvl[p], true, root.half);
    int carry = 1;
    for (;;) {
        const int q = t.parent[p];
        if (q < 0) break;
        BH_FENCE();
        const int old = BH_ATOMIC_ADD_INT(&t.arrived[q], carry);
        if (old + carry != t.cnt[q]) break;
        BH_FENCE();
        double mSum = 0.0, sx = 0.0, sy = 0.0;
        const int end = t.skip[q];
        for (int c = q + 1; c < end; c = t.skip[c]) {
            const double mc = BH_LD_D(&t.cmass[c]);
            if (mc > 0.0) {   // BH.kt:189-192
                mSum = BH_DADD(mSum, mc);
                sx = BH_DADD(sx, BH_DMUL(BH_LD_D(&t.comx[c]), mc));
                sy = BH_DADD(sy, BH_DMUL(BH_LD_D(&t.comy[c]), mc));
            }
        }
        const int level = t.lvl[q];
        double cx, cy;
        if (mSum > 0.0) { cx = BH_DDIV(sx, mSum); cy = BH_DDIV(sy, mSum); }   // BH.kt:194-196
        else { double h; bh_cell_geometry(root, t.keys[i], level, &cx, &cy, &h); }  // BH.kt:197-200
        bh_write_cell(t, q, cx, cy, mSum, level, false, root.half);
        carry = t.cnt[q];
        p = q;
    }
}

struct BhWalkParams {
    float  th2f, soft2f;      // FP32 copies for the fast test
    double theta2, soft2;     // BH.kt:378, Config.kt:20
    double half;              // root half-side
};

struct BhWalkResult { double ax, ay; int interactions, opened, retests; };

// accumulateForce (BH.kt:215-239) for one body, stackless over the preorder array.
// `self` = preorder position of the body's own leaf (-1 if it is not in the tree).
// Per-body decisions are the reference's: FP32 test outside the guard band, the exact f64
// expression inside it.  Interaction math is FP32 on (hi,lo)-split coordinate differences;
// the FP32 partial sums are folded into f64 accumulators every 16 visits (all lanes of a
// warp share the visit counter, so the fold is a uniform branch) — this removes the FP32
// accumulation error, which otherwise dominates.
// Returns sum m*d/r^3 (G is applied by the caller).
BH_HD BhWalkResult bh_walk_body(const BhTreeView& t, const BhWalkParams& w, double x, double y, int self) {
    float xh, xl, yh, yl;
    bh_split(x, &xh, &xl);
    bh_split(y, &yh, &yl);
    BhWalkResult r; r.ax = 0.0; r.ay = 0.0; r.interactions = 0; r.opened = 0; r.retests = 0;
    const float ghi = 1.0f + BH_GUARD_BAND, glo = 1.0f - BH_GUARD_BAND;
    float fx = 0.f, fy = 0.f;
    int p = 0, it = 0;
    const int M = t.M;
    while (p < M) {
        const BhCellA a = t.A[p];
        const BhCellB b = t.B[p];
        const float dx = (a.xh - xh) + (b.xl - xl);
        const float dy = (a.yh - yh) + (b.yl - yl);
        const float d2 = fmaf(dx, dx, fmaf(dy, dy, w.soft2f));
        const float tt = w.th2f * d2;
        bool accept = a.s2 * ghi < tt;
        if (!accept && !(a.s2 * glo > tt)) {   // borderline: the reference's f64 test decides
            accept = bh_exact_accept(t.comx[p], t.comy[p], x, y, w.soft2, w.theta2, w.half, b.level);
            r.retests++;
        }
        if (accept) {
            if (p != self) {
                float inv = BH_RSQRTF(d2);
#if defined(__CUDA_ARCH__)
                inv = inv * fmaf(-0.5f * d2, inv * inv, 1.5f);   // one Newton step on MUFU.RSQ
#endif
                const float wgt = a.m * inv * inv * inv;
                fx = fmaf(wgt, dx, fx);
                fy = fmaf(wgt, dy, fy);
                r.interactions += (a.m != 0.0f);
            }
            p = b.skip;
        } else {
            r.opened++;
            p = p + 1;
        }
        if ((++it & 15) == 0) { r.ax += (double)fx; r.ay += (double)fy; fx = 0.f; fy = 0.f; }
    }
    r.ax += (double)fx; r.ay += (double)fy;
    return r;
}

#endif  // BH_CORE_H
