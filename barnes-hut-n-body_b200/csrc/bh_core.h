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
#include <string.h>

#if defined(__CUDACC__)
#define BH_HD __host__ __device__ __forceinline__
#define BH_HD_COLD __host__ __device__ __noinline__
#else
#define BH_HD inline
#define BH_HD_COLD inline
#endif

#if !defined(__CUDACC__)
struct float4 { float x, y, z, w; };   // host stand-in for the CUDA vector type
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

// Closed form of the same key.  When every cell boundary x0 + k*w (w = 2*half/2^levels) of the
// root box is exactly representable in binary64 (bh_grid_is_exact: true for the reference's
// half-integer root boxes), the centres Quad.child builds by repeated +-h/2 are exact, so the
// descent's comparisons locate x on that exact grid: the column index is floor((x-x0)/w) in exact
// arithmetic.  A floating-point estimate is corrected against the exact boundaries (at most one
// step), then the two indices are bit-interleaved (x = low bit, y = high bit of each digit).
struct BhGrid { double x0, y0, w, inv_w; int exact; };

BH_HD bool bh_is_multiple_of_pow2(double v, int e) {   // v is an integer multiple of 2^e
    const double s = ldexp(v, -e);
    return s == floor(s) && fabs(s) < 9007199254740992.0;
}
BH_HD BhGrid bh_make_grid(const BhRoot& r) {
    BhGrid g;
    g.x0 = BH_DSUB(r.cx, r.half); g.y0 = BH_DSUB(r.cy, r.half);
    g.w = ldexp(r.half, 1 - r.levels);
    g.inv_w = 1.0 / g.w;
    // quantum of every centre / boundary: half * 2^-levels; all of cx, cy, half must be multiples of
    // a power of two 2^e with (|coordinate| + 2*half) / 2^(e - levels) < 2^52
    int ok = 0;
    for (int e = 0; e >= -12 && !ok; --e) {
        if (bh_is_multiple_of_pow2(r.cx, e) && bh_is_multiple_of_pow2(r.cy, e) && bh_is_multiple_of_pow2(r.half, e)) {
            const double span = fmax(fabs(r.cx), fabs(r.cy)) + 2.0 * r.half;
            ok = ldexp(span, r.levels - e + 1) < 4503599627370496.0;
            break;
        }
    }
    g.exact = ok && r.levels >= 1 && r.levels <= 31;
    return g;
}
BH_HD uint32_t bh_grid_index(double v, double v0, double w, double inv_w, int levels) {
    const double nmax = (double)((1u << levels) - 1u);
    double k = floor(BH_DMUL(BH_DSUB(v, v0), inv_w));
    k = k < 0.0 ? 0.0 : (k > nmax ? nmax : k);
    // exact boundaries b(k) = v0 + k*w (representable by construction)
    while (k > 0.0 && v < BH_DADD(v0, BH_DMUL(k, w))) k -= 1.0;
    while (k < nmax && v >= BH_DADD(v0, BH_DMUL(k + 1.0, w))) k += 1.0;
    return (uint32_t)k;
}
BH_HD uint64_t bh_spread_bits(uint32_t v) {   // bit i -> bit 2i
    uint64_t x = v;
    x = (x | (x << 16)) & 0x0000FFFF0000FFFFull;
    x = (x | (x << 8)) & 0x00FF00FF00FF00FFull;
    x = (x | (x << 4)) & 0x0F0F0F0F0F0F0F0Full;
    x = (x | (x << 2)) & 0x3333333333333333ull;
    x = (x | (x << 1)) & 0x5555555555555555ull;
    return x;
}
BH_HD uint64_t bh_morton_key_grid(const BhGrid& g, int levels, double x, double y) {
    const uint32_t ix = bh_grid_index(x, g.x0, g.w, g.inv_w, levels);
    const uint32_t iy = bh_grid_index(y, g.y0, g.w, g.inv_w, levels);
    return bh_spread_bits(ix) | (bh_spread_bits(iy) << 1);
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

// Sorted keys / scan values behind an accessor (the searches below are templated on it).
struct BhKeysGlobal {
    const uint64_t* __restrict__ keys;
    const int* __restrict__ S;
    BH_HD uint64_t key(int j) const { return keys[j]; }
    BH_HD int scan(int j) const { return S[j]; }
};
// last index j >= from (from is known to be inside) whose key shares the prefix
// (key >> sh) == pref; keys sorted ascending, n = number of in-tree keys
template <class K>
BH_HD int bh_gallop_right(const K& k, int n, int from, uint64_t pref, int sh) {
    int j = from;
    int step = 1;
    while (j + step < n && (k.key(j + step) >> sh) == pref) { j += step; step <<= 1; }
    int out = (j + step < n) ? (j + step) : n;   // first index known to be outside (or n)
    while (out - j > 1) {
        const int mid = j + ((out - j) >> 1);
        if ((k.key(mid) >> sh) == pref) j = mid; else out = mid;
    }
    return j;
}

// first index j <= from whose key shares the prefix
template <class K>
BH_HD int bh_gallop_left(const K& k, int from, uint64_t pref, int sh) {
    int j = from;
    int step = 1;
    while (j - step >= 0 && (k.key(j - step) >> sh) == pref) { j -= step; step <<= 1; }
    int out = (j - step >= 0) ? (j - step) : -1;  // last index known to be outside (or -1)
    while (j - out > 1) {
        const int mid = out + ((j - out) >> 1);
        if ((k.key(mid) >> sh) == pref) j = mid; else out = mid;
    }
    return j;
}
BH_HD int bh_gallop_right(const uint64_t* __restrict__ keys, int n, int from, uint64_t pref, int sh) {
    return bh_gallop_right(BhKeysGlobal{keys, nullptr}, n, from, pref, sh);
}
BH_HD int bh_gallop_left(const uint64_t* __restrict__ keys, int from, uint64_t pref, int sh) {
    return bh_gallop_left(BhKeysGlobal{keys, nullptr}, from, pref, sh);
}

// shift that isolates the first d digits of a key
BH_HD int bh_prefix_shift(int levels, int d) { return 2 * (levels - d); }

// ---- cell records -------------------------------------------------------------------------
// Hot record read by the walk: ONE 32-byte sector per cell, loaded as two 16-byte vectors.
//   (xh, yh, m, s2)         centre of mass (hi parts), mass, (cell side)^2 as float;
//                           s2 = -1 for a leaf or a zero-mass cell: always "accepted"
//                           (no opening test for leaves, zero mass is pruned: BH.kt:216-221)
//   (xl, yl, skip, band)    lo parts; preorder position after the subtree; half-width of the
//                           band |theta^2 d^2 - s^2| <= band in which the FP32 opening test is
//                           re-done in f64 (BH_GUARD_BAND * s^2; 0 for a leaf / zero-mass cell)
struct alignas(32) BhCell {
    float xh, yh, m, s2;
    float xl, yl;
    int   skip;
    float band;
};
// Exact record: f64 centre of mass and mass, bit-identical to BHTree.computeMass
// (BH.kt:173-202).  Read by the climb, by borderline opening tests and by the export.
struct alignas(32) BhCellD { double comx, comy, mass, pad; };
// Skeleton written by bh_emit_body.
struct alignas(16) BhCellS { int skip, parent, cnt, level; };

// The reference's f64 opening test, BH.kt:223-228, bit-for-bit.
BH_HD bool bh_exact_accept(double comx, double comy, double x, double y, double soft2, double theta2,
                           double half, int level) {
    const double dx = BH_DSUB(comx, x);
    const double dy = BH_DSUB(comy, y);
    const double dist2 = BH_DADD(BH_DADD(BH_DMUL(dx, dx), BH_DMUL(dy, dy)), soft2);
    const double s2 = bh_side2(half, level);
    return s2 < BH_DMUL(theta2, dist2);
}

// ---- view of the preorder cell arrays -------------------------------------------------------
struct BhTreeView {
    const uint64_t* keys;   // sorted keys of the in-tree bodies            [n_in]
    const int*      order;  // body index at each sorted position            [n]
    const int*      S;      // exclusive scan of cnt(i), S[n_in] = #internal [n_in+1]
    BhCell*  cell;          // hot records                                   [M]
    BhCellD* cd;            // exact f64 records                             [M]
    BhCellS* sk;            // skeletons                                     [M]
    int*     arrived;       // climb counters, zero before the climb         [M]
    int n_in;               // bodies in the tree
    int M;                  // cells = n_in + #internal
    // Sync-free builds (small body counts: the cell arrays are sized for the worst case, so the host never reads
    // the counts back): n_in and the number of internal cells are taken from device memory by every kernel.
    const int* dev_n_in;    // nullptr: n_in / M above are valid
    const int* dev_n_int;
};
// kernels call this first: with device-side counts it fills n_in / M of their (by-value) copy of the view
#if defined(__CUDACC__)
__device__ __forceinline__ void bh_view_resolve(BhTreeView& t) {
    if (t.dev_n_in) { t.n_in = *t.dev_n_in; t.M = t.n_in + *t.dev_n_int; }
}
#endif

#if defined(__CUDA_ARCH__)
// acq_rel RMW: releases this thread's child record, acquires the siblings' records
__device__ __forceinline__ int bh_atomic_add_acq_rel(int* p, int v) {
    int old;
    asm volatile("atom.acq_rel.gpu.global.add.s32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}
#define BH_ATOMIC_ADD_INT(p, v) bh_atomic_add_acq_rel((p), (v))
#define BH_LD_D(p) __ldcg(p)
// MUFU.RSQ without the denormal fix-up sequence (d2 >= soft2 is never subnormal)
__device__ __forceinline__ float bh_rsqrt_ftz(float v) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
#define BH_RSQRTF(v) bh_rsqrt_ftz(v)
// an identity shuffle: stops ptxas from re-materialising the f64->f32 conversions in the loop
#define BH_OPAQUE_F(v) (v) = __shfl_sync(0xffffffffu, (v), threadIdx.x & 31)
#define BH_NOINLINE __noinline__
#else
static inline int bh_host_fetch_add(int* p, int v) { const int o = *p; *p = o + v; return o; }
#define BH_ATOMIC_ADD_INT(p, v) bh_host_fetch_add((p), (v))
#define BH_LD_D(p) (*(p))
#define BH_RSQRTF(v) (1.0f / sqrtf(v))
#define BH_OPAQUE_F(v) ((void)0)
#define BH_NOINLINE
#endif

// Skeleton of everything body i "owns" in the preorder array: the column of internal
// cells whose leftmost key is key[i], and the leaf of body i.  Pure function of the sorted
// keys and S (no atomics); one thread per in-tree body.
template <class K>
BH_HD void bh_emit_body(const BhTreeView& t, const K& ks, int levels, int i) {
    const int n = t.n_in;
    const uint64_t k = ks.key(i);
    const int dprev = (i > 0) ? bh_common_levels(ks.key(i - 1), k, levels) : -1;
    const int dnext = (i + 1 < n) ? bh_common_levels(k, ks.key(i + 1), levels) : -1;
    const int base = ks.scan(i) + i;
    // parent of the first entry of this group: the depth-dprev cell holding key[i-1] and key[i]
    int headParent = -1;
    if (i > 0) {
        const int sh = bh_prefix_shift(levels, dprev);
        const int il = bh_gallop_left(ks, i, k >> sh, sh);
        const int dl = (il > 0) ? bh_common_levels(ks.key(il - 1), ks.key(il), levels) : -1;
        headParent = ks.scan(il) + il + (dprev - dl - 1);
    }
    const int ncol = (dnext > dprev) ? (dnext - dprev) : 0;
    int hi = i;
    for (int d = dnext; d > dprev; --d) {   // deepest first: hi only grows
        const int sh = bh_prefix_shift(levels, d);
        hi = bh_gallop_right(ks, n, hi, k >> sh, sh);
        const int p = base + (d - dprev - 1);
        BhCellS c;
        c.skip = ks.scan(hi + 1) + hi + 1;
        c.parent = (d == dprev + 1) ? headParent : (p - 1);
        c.cnt = hi - i + 1;
        c.level = d;
        t.sk[p] = c;
    }
    const int lp = base + ncol;
    BhCellS c;
    c.skip = lp + 1;
    c.parent = (ncol > 0) ? (lp - 1) : headParent;
    c.cnt = 1;
    c.level = ((dprev > dnext) ? dprev : dnext) + 1;
    t.sk[lp] = c;
}
BH_HD void bh_emit_body(const BhTreeView& t, int levels, int i) { bh_emit_body(t, BhKeysGlobal{t.keys, t.S}, levels, i); }

BH_HD void bh_write_cell(const BhTreeView& t, int p, double cx, double cy, double m, int skip, int level, bool leaf,
                         double half) {
    BhCellD d; d.comx = cx; d.comy = cy; d.mass = m; d.pad = 0.0;
    t.cd[p] = d;
    BhCell c;
    bh_split(cx, &c.xh, &c.xl);
    bh_split(cy, &c.yh, &c.yl);
    c.m = (float)m;
    const bool always = leaf || m == 0.0;
    const float s2f = (float)bh_side2(half, level);
    c.s2 = always ? -1.0f : s2f;
    c.skip = skip;
    c.band = always ? 0.0f : BH_GUARD_BAND * s2f;
    t.cell[p] = c;
}

// computeMass (BH.kt:173-202) bottom-up: the thread of body i writes its leaf, then climbs;
// at each parent it adds the body count of the finished child (one acq_rel atomic) and
// continues only if that completed the parent — the last arriver sums the children in
// child order 0..3 (= preorder order) with the reference's exact f64 expression order.
// Continues the climb above a finished cell whose skeleton is `s` and which holds `carry` bodies;
// `key` = Morton key of any body below it (cell geometry of a zero-mass ancestor, BH.kt:197-200).
BH_HD void bh_climb_from(const BhTreeView& t, const BhRoot& root, uint64_t key, BhCellS s, int carry) {
    for (;;) {
        const int q = s.parent;
        if (q < 0) break;
        s = t.sk[q];
        const int old = BH_ATOMIC_ADD_INT(&t.arrived[q], carry);
        if (old + carry != s.cnt) break;
        double mSum = 0.0, sx = 0.0, sy = 0.0;
        for (int c = q + 1; c < s.skip; c = t.sk[c].skip) {
            const double mc = BH_LD_D(&t.cd[c].mass);
            if (mc > 0.0) {   // BH.kt:189-192
                mSum = BH_DADD(mSum, mc);
                sx = BH_DADD(sx, BH_DMUL(BH_LD_D(&t.cd[c].comx), mc));
                sy = BH_DADD(sy, BH_DMUL(BH_LD_D(&t.cd[c].comy), mc));
            }
        }
        double cx, cy;
        if (mSum > 0.0) { cx = BH_DDIV(sx, mSum); cy = BH_DDIV(sy, mSum); }   // BH.kt:194-196
        else { double h; bh_cell_geometry(root, key, s.level, &cx, &cy, &h); }  // BH.kt:197-200
        bh_write_cell(t, q, cx, cy, mSum, s.skip, s.level, false, root.half);
        carry = s.cnt;
    }
}
BH_HD void bh_climb_body(const BhTreeView& t, const BhRoot& root, int i, double x, double y, double m) {
    const int p = t.S[i + 1] + i;
    const BhCellS s = t.sk[p];
    bh_write_cell(t, p, x, y, m, s.skip, s.level, true, root.half);
    bh_climb_from(t, root, t.keys[i], s, 1);
}

// ---- jitter regime (BH.kt:145-156) ---------------------------------------------------------------
// A node with half-side h < 1e-3 (depth >= root.levels) that has to push bodies into its children
// first MUTATES them: x += (LSB(bits(x)) == 0 ? +1e-3 : -1e-3), y += (LSB(bits(y)) == 0 ? -1e-3 :
// +1e-3); a body shifted out of the child picked for it is silently dropped by the next insert()
// (BH.kt:126).  Only bodies that share a depth-`levels` cell C (equal Morton keys) are affected, and
// the outcome depends on their insertion order (list order), so each such cluster is replayed
// sequentially, in f64, exactly as BHTree.insert / insertIntoChild would run (BH.kt:125-156).
// Because C's children are narrower than the 1e-3 shift, nothing survives below depth levels+1:
// a child of C ends up empty, holding one body, or "dead" (subdivided, all four children empty).
//
// ord[0..r) = body slots of the cluster (r >= 2); perm[slot] = position in the reference's list.
// On return x/y hold the mutated coordinates, jflag[slot] bit 0 = 1 for a dropped body (it stays
// in the sorted order as a zero-mass ghost leaf: invisible to masses and forces), ord[] lists the
// surviving bodies first, by child digit (the order computeMass sums them in), and
// jflag[ord[0]] bits 4..7 = mask of dead children (only the debug export shows those).
BH_HD double bh_jitter_shift(double v, bool plus_if_even) {
    long long bits;
#if defined(__CUDA_ARCH__)
    bits = __double_as_longlong(v);
#else
    memcpy(&bits, &v, 8);
#endif
    const bool even = (bits & 1LL) == 0LL;
    return BH_DADD(v, (even == plus_if_even) ? 1e-3 : -1e-3);
}
BH_HD bool bh_quad_contains(double cx, double cy, double h, double x, double y) {   // BH.kt:61-62
    return x >= BH_DSUB(cx, h) && x < BH_DADD(cx, h) && y >= BH_DSUB(cy, h) && y < BH_DADD(cy, h);
}
BH_HD void bh_jitter_cluster(const BhRoot& root, uint64_t key, int* ord, int r, const int* __restrict__ perm,
                             double* x, double* y, int* jflag, int* unsupported, int* n_ghost) {
    // insertion order = list order: heap-sort the members by perm[]
    for (int start = r / 2 - 1; start >= 0; --start) {
        int i = start;
        for (;;) {
            int c = 2 * i + 1;
            if (c >= r) break;
            if (c + 1 < r && perm[ord[c + 1]] > perm[ord[c]]) ++c;
            if (perm[ord[c]] <= perm[ord[i]]) break;
            const int tmp = ord[c]; ord[c] = ord[i]; ord[i] = tmp; i = c;
        }
    }
    for (int end = r - 1; end > 0; --end) {
        int tmp = ord[0]; ord[0] = ord[end]; ord[end] = tmp;
        int i = 0;
        for (;;) {
            int c = 2 * i + 1;
            if (c >= end) break;
            if (c + 1 < end && perm[ord[c + 1]] > perm[ord[c]]) ++c;
            if (perm[ord[c]] <= perm[ord[i]]) break;
            tmp = ord[c]; ord[c] = ord[i]; ord[i] = tmp; i = c;
        }
    }
    double cx, cy, h;
    bh_cell_geometry(root, key, root.levels, &cx, &cy, &h);
    const double hh = h / 2.0;            // Quad.child, BH.kt:73-80
    int kbody[4] = {-1, -1, -1, -1};
    int dead = 0;
    // insertIntoChild of a DEAD child K (or of K while it subdivides): jitter, pick K's child, insert
    // there; its containment test cannot pass (K is narrower than the shift) but is made literally.
    auto sink = [&](int b, double kcx, double kcy) {
        x[b] = bh_jitter_shift(x[b], true);
        y[b] = bh_jitter_shift(y[b], false);
        const double qh = hh / 2.0;
        const double qcx = (x[b] < kcx) ? BH_DSUB(kcx, qh) : BH_DADD(kcx, qh);
        const double qcy = (y[b] < kcy) ? BH_DSUB(kcy, qh) : BH_DADD(kcy, qh);
        if (bh_quad_contains(qcx, qcy, qh, x[b], y[b])) *unsupported = 1;
    };
    // insertIntoChild at C (BH.kt:145-156) followed by the child's insert() (BH.kt:125-137)
    auto place = [&](int b) {
        x[b] = bh_jitter_shift(x[b], true);
        y[b] = bh_jitter_shift(y[b], false);
        const int ix = (x[b] < cx) ? 0 : 1, iy = (y[b] < cy) ? 0 : 2;
        const int c = ix + iy;
        const double kcx = ix ? BH_DADD(cx, hh) : BH_DSUB(cx, hh);
        const double kcy = iy ? BH_DADD(cy, hh) : BH_DSUB(cy, hh);
        if (!bh_quad_contains(kcx, kcy, hh, x[b], y[b])) return;          // dropped, BH.kt:126
        if (dead & (1 << c)) { sink(b, kcx, kcy); return; }
        if (kbody[c] < 0) { kbody[c] = b; return; }                       // empty leaf, BH.kt:127-130
        const int e = kbody[c];                                           // subdivide, BH.kt:131-136
        kbody[c] = -1;
        dead |= 1 << c;
        sink(e, kcx, kcy);
        sink(b, kcx, kcy);
    };
    place(ord[0]);                        // the resident of C is pushed down first (BH.kt:132-135)
    for (int k = 1; k < r; ++k) place(ord[k]);
    // survivors first, by child digit; the dropped bodies keep their relative order behind them
    int nt = 0;
    for (int k = 0; k < r; ++k) jflag[ord[k]] = 1;
    for (int c = 0; c < 4; ++c) if (kbody[c] >= 0) { jflag[kbody[c]] = 0; ++nt; }
    int w = r - 1;
    for (int k = r - 1; k >= 0; --k) if (jflag[ord[k]] & 1) ord[w--] = ord[k];   // compact ghosts to the back (stable)
    w = 0;
    for (int c = 0; c < 4; ++c) if (kbody[c] >= 0) ord[w++] = kbody[c];
    jflag[ord[0]] |= dead << 4;
    *n_ghost = r - nt;
}

struct BhWalkParams {
    float  th2f;              // theta^2 as float: the FP32 opening test is fma(d2, th2f, -s2) > 0
    float  soft2f;
    double theta2, soft2;     // BH.kt:378, Config.kt:20
    double half;              // root half-side
};

BH_HD BhWalkParams bh_walk_params(double theta, double soft2, double half) {
    BhWalkParams w;
    w.theta2 = theta * theta;   // BH.kt:378
    w.soft2 = soft2;
    w.half = half;
    w.th2f = (float)w.theta2;
    w.soft2f = (float)soft2;
    return w;
}

struct BhWalkResult { double ax, ay; int interactions, opened, retests; };

#define BH_WALK_CHUNK 16      // visits between two folds of the FP32 partial sums into f64

// The reference's f64 opening test for a borderline cell (cold path), BH.kt:223-228 bit-for-bit.
BH_HD_COLD bool bh_retest_cell(const BhCellD* __restrict__ cd, const BhCellS* __restrict__ sk, int p, double x, double y,
                               double soft2, double theta2, double half) {
    return bh_exact_accept(cd[p].comx, cd[p].comy, x, y, soft2, theta2, half, sk[p].level);
}

// accumulateForce (BH.kt:215-239) for one body, stackless over the preorder array: the HOST form (CPU
// emulation of the device walk, tests only).  `self` = preorder position of the body's own leaf (-1 if it
// is not in the tree).  Per-body decisions are the reference's: the FP32 test t = theta^2 d^2 - s^2 > 0
// decides unless |t| is inside the cell's guard band, where the exact f64 expression decides.
// Interaction math is FP32 on (hi,lo)-split coordinate differences; FP32 partial sums are folded into
// f64 accumulators every BH_WALK_CHUNK visits.  Returns sum m*d/r^3 (G is applied by the caller).
BH_HD BhWalkResult bh_walk_body(const BhTreeView& t, const BhWalkParams& w, double x, double y, int self, bool active) {
    float xh, xl, yh, yl;
    bh_split(x, &xh, &xl);
    bh_split(y, &yh, &yl);
    BhWalkResult r; r.ax = 0.0; r.ay = 0.0; r.interactions = 0; r.opened = 0; r.retests = 0;
    const float th2 = w.th2f, soft2 = w.soft2f;
    const BhCell* __restrict__ cells = t.cell;
    const int M = t.M;
    int p = active ? 0 : M;
    while (p < M) {
        float fx = 0.f, fy = 0.f;
        for (int k = 0; k < BH_WALK_CHUNK; ++k) {
            const BhCell& c = cells[p];
            const float dx = (c.xh - xh) + (c.xl - xl);
            const float dy = (c.yh - yh) + (c.yl - yl);
            const float d2 = fmaf(dx, dx, fmaf(dy, dy, soft2));
            const float tt = fmaf(d2, th2, -c.s2);
            bool accept = tt > 0.0f;
            if (fabsf(tt) <= c.band) {   // borderline: the reference's f64 test decides
                accept = bh_retest_cell(t.cd, t.sk, p, x, y, w.soft2, w.theta2, w.half);
                r.retests++;
            }
            const float inv = BH_RSQRTF(d2);
            const bool use = accept && (p != self) && (c.m != 0.0f);
            const float wgt = use ? c.m * inv * inv * inv : 0.0f;
            fx = fmaf(wgt, dx, fx);
            fy = fmaf(wgt, dy, fy);
            r.interactions += use;
            r.opened += !accept;
            p = accept ? c.skip : p + 1;
            if (p >= M) break;
        }
        r.ax += (double)fx; r.ay += (double)fy;
    }
    return r;
}

// the terminal record cell[M] finished lanes idle on
BH_HD void bh_write_terminal_cell(const BhTreeView& t) {
    BhCell c;
    c.xh = 0.f; c.yh = 0.f; c.m = 0.f; c.s2 = -1.0f; c.xl = 0.f; c.yl = 0.f; c.skip = t.M; c.band = 0.0f;
    t.cell[t.M] = c;
}

#if defined(__CUDACC__)
// ---- the DEVICE walk: G Morton-consecutive bodies per thread, one shared preorder position ---------
// ncu on the one-body-per-lane walk of round 1 (profiles/r01c): the L1 data stage is the limiter, not
// the issue slots — the 32 lanes of a warp sit at ~15 different cells, and every distinct 32 B sector
// of a warp-wide load costs a data-stage cycle (15 cycles per warp iteration against 35 issue slots
// shared by 4 schedulers).  Here a THREAD walks for G bodies at once: the record is loaded once and
// tested against the G bodies, so a warp-wide load serves 32*G body visits and the G independent
// dependency chains give the scheduler instruction-level parallelism.
//
// Per-body decisions stay the reference's (BH.kt:223-228 is evaluated per body): a cell is opened when
// ANY un-muted body of the group opens it; a body that accepts a cell the group opens takes the
// interaction and is MUTED (mute = skip of that cell) until the walk leaves the subtree.  The group
// visits the union of its bodies' cells (probe bh_emul_group_stats: 1.06x the visits of one body for
// G = 2, 1.14x for G = 4 on the bench cloud).  G = 1 is the plain per-lane walk (no mute logic).
//
// One body's share of a visit: 3 FADD2 (packed f32x2 arithmetic of sm_100 on (x, y): the (hi,lo)-split
// difference of both coordinates), 2 FFMA (d^2); then, packed over PAIRS of bodies (the cell's scalars are
// broadcast operands): the test t = theta^2 d^2 - s^2, and — after one MUFU.RSQ per body — one Newton step
// written ic = inv*(d2*inv^2 - 3) (= -2 x the refined 1/sqrt; the factor (-2)^3 is divided out exactly at
// the end) and m*ic^3; then per body the guard-band compare, ISETP + FSETP (un-muted / accepts / opens),
// five packed predicated instructions of compensated accumulation, two predicated counters, the mute update.
// A zero-mass cell is "always accepted" (BH.kt:216 prunes it) and the body's own leaf is skipped
// (BH.kt:219): one predicate each keeps them out of the sums and the counts.  Finished threads idle on
// the terminal record cell[M] (mass 0, accepted, skip = M), so there is no per-visit exit test.
template <int G>
struct BhMultiResult {
    double ax[G], ay[G];
    int interactions[G], opened[G];
    int retests;
};

// one 256-bit load of a whole 32-byte cell record (sm_100: LDG.E.256): one L1 wavefront per
// distinct line instead of two
__device__ __forceinline__ void bh_load_cell(const BhCell* __restrict__ c, float4* a, float4* b) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a->x), "=f"(a->y), "=f"(a->z), "=f"(a->w), "=f"(b->x), "=f"(b->y), "=f"(b->z), "=f"(b->w)
                 : "l"(c));
}
// MUFU.RSQ without the denormal fix-up sequence (d2 >= soft2 is never subnormal)
__device__ __forceinline__ float bh_mufu_rsq(float v) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}

// How a body's force terms are summed.
//   BH_ACC_FOLD  FP32 partial sums, folded into f64 every BH_WALK_CHUNK iterations of the loop.  With one
//                body per thread an iteration is one of the body's own visits, so the result depends on the
//                body's own walk only; with G > 1 the fold timing follows the GROUP's iterations, i.e. a
//                body's rounding would depend on its group-mates.
//   BH_ACC_F64   every term (an FP32 product) is converted and added to an f64 sum right away: no partial
//                sums, no timing — the result depends on nothing but the body's own sequence of
//                interactions, however its group-mates are chosen, hence across any multi-GPU partition
//                of the targets (the bit-identity tests).  5 more issue slots per visit.
#define BH_ACC_FOLD 0
#define BH_ACC_F64 1

// Last part of one body's share of a visit: decisions, accumulation, counters, next position.
// operands  %0 ax  %1 ay  %2 interactions  %3 opened  %4 mute  %5 pnext  |  %6 tt  %7 wg  %8 m  %9 dx
// %10 dy  %11 skip  %12 p  %13 p+1  %14 own leaf.   Pa accepts, Po opens (while un-muted only), Pc = Pa and
// mass != 0 and not the own leaf (BH.kt:216,219): the interaction proper.  ax/ay are FP32 or f64.
#define BH_MULTI_ACC_FOLD_ASM                                                                    \
        "@Pc fma.rn.f32 %0, %7, %9, %0;\n\t"                                                    \
        "@Pc fma.rn.f32 %1, %7, %10, %1;\n\t"
#define BH_MULTI_ACC_F64_ASM   /* FP64 adds are not predicated by ptxas (it would add selects): a masked weight */ \
        "selp.f32 tx, %7, 0f00000000, Pc;\n\t"           /* makes the term +-0, and x + 0 = x exactly        */ \
        "mul.f32 ty, tx, %10;\n\t"                                                              \
        "mul.f32 tx, tx, %9;\n\t"                                                               \
        "cvt.f64.f32 dx8, tx;\n\t"                                                              \
        "cvt.f64.f32 dy8, ty;\n\t"                                                              \
        "add.rn.f64 %0, %0, dx8;\n\t"                                                           \
        "add.rn.f64 %1, %1, dy8;\n\t"
#define BH_MULTI_TAIL(J, PRED_ASM, ACC_ASM, ACC_C, MOVE_ASM)                                    \
    asm("{\n\t"                                                                                 \
        ".reg .pred Pu, Pa, Po, Pc;\n\t"                                                        \
        ".reg .f32 tx, ty;\n\t"                                                                 \
        ".reg .f64 dx8, dy8;\n\t"                                                               \
        PRED_ASM                                                                                \
        "setp.neu.and.f32 Pc, %8, 0f00000000, Pa;\n\t"    /* mass != 0 (BH.kt:216)            */ \
        "setp.ne.and.s32 Pc, %12, %14, Pc;\n\t"           /* not the body's own leaf (:219)   */ \
        ACC_ASM                                                                                 \
        "@Pc add.s32 %2, %2, 1;\n\t"                                                            \
        "@Po add.s32 %3, %3, 1;\n\t"                                                            \
        MOVE_ASM                                                                                \
        "}"                                                                                     \
        : ACC_C(sx[J]), ACC_C(sy[J]), "+r"(ni[J]), "+r"(no[J]), "+r"(mute[J]), "+r"(pnext)      \
        : "f"(tt[J]), "f"(wg[J]), "f"(c0.z), "f"(dx[J]), "f"(dy[J]), "r"(skip), "r"(p), "r"(p1), "r"(self[J]))
#define BH_ACC_C_F32 "+f"
#define BH_ACC_C_F64 "+d"
// G > 1: the mute logic; the group opens the cell (pnext = p + 1) when any of its bodies does
#define BH_PRED_GROUP                                                                           \
        "setp.ge.s32 Pu, %12, %4;\n\t"                    /* un-muted                         */ \
        "setp.gt.and.f32 Pa|Po, %6, 0f00000000, Pu;\n\t"  /* accepts | opens                  */
#define BH_MOVE_GROUP                                                                           \
        "@Pa mov.s32 %4, %11;\n\t"                        /* muted until the walk leaves it   */ \
        "@Po mov.s32 %5, %13;\n\t"
// G == 1: the plain per-lane walk (accept -> skip, open -> p + 1)
#define BH_PRED_SINGLE "setp.gt.f32 Pa|Po, %6, 0f00000000;\n\t"
#define BH_MOVE_SINGLE "@Po mov.s32 %5, %13;\n\t"

#ifndef BH_WALK_NEWTON
#define BH_WALK_NEWTON 1      // one Newton step on MUFU.RSQ
#endif
// m * (refined 1/sqrt(d2))^3 up to the exact factor BH_WALK_WSCALE, for one body and for a pair of bodies
// (packed f32x2 instructions of sm_100: the two bodies of a pair share every instruction but the MUFU)
#if BH_WALK_NEWTON
#define BH_WALK_WSCALE (-0.125)   // ic = inv*(d2*inv^2 - 3) = -2 x the refined 1/sqrt: (-2)^3 is divided out at the end
__device__ __forceinline__ float bh_weight1(float d2, float m) {
    const float inv = bh_mufu_rsq(d2);
    const float ic = inv * fmaf(d2, inv * inv, -3.0f);
    return (m * ic) * (ic * ic);
}
__device__ __forceinline__ float2 bh_weight2(float2 d2, float m) {
    const float2 inv = make_float2(bh_mufu_rsq(d2.x), bh_mufu_rsq(d2.y));
    const float2 ic = __fmul2_rn(inv, __ffma2_rn(d2, __fmul2_rn(inv, inv), make_float2(-3.0f, -3.0f)));
    return __fmul2_rn(__fmul2_rn(make_float2(m, m), ic), __fmul2_rn(ic, ic));
}
#else
#define BH_WALK_WSCALE 1.0
__device__ __forceinline__ float bh_weight1(float d2, float m) {
    const float inv = bh_mufu_rsq(d2);
    return (m * inv) * (inv * inv);
}
__device__ __forceinline__ float2 bh_weight2(float2 d2, float m) {
    const float2 inv = make_float2(bh_mufu_rsq(d2.x), bh_mufu_rsq(d2.y));
    return __fmul2_rn(__fmul2_rn(make_float2(m, m), inv), __fmul2_rn(inv, inv));
}
#endif

template <int ACC> struct BhAccT { typedef float type; };
template <> struct BhAccT<BH_ACC_F64> { typedef double type; };

// x[j], y[j] for j < nactive are the bodies of this thread (Morton-consecutive); every thread of the warp
// must call (the loop contains full-warp votes).  `zero`: see below.
template <int G, int ACC>
__device__ __forceinline__ void bh_walk_multi(const BhTreeView& t, const BhWalkParams& w, const double* x, const double* y,
                                              const int* self, int nactive, int zero, BhMultiResult<G>* out) {
    typedef typename BhAccT<ACC>::type acc_t;
    float2 nh[G], nl[G];                  // -(hi parts), -(lo parts) of the bodies' (x, y)
    acc_t sx[G], sy[G];                   // BH_ACC_FOLD: FP32 partial sums of the chunk; BH_ACC_F64: the f64 sums
    double fx[G], fy[G];                  // BH_ACC_FOLD: the f64 sums
    int ni[G], no[G], mute[G];
#pragma unroll
    for (int j = 0; j < G; ++j) {
        float xh, xl, yh, yl;
        bh_split(x[j], &xh, &xl);
        bh_split(y[j], &yh, &yl);
        nh[j] = make_float2(-xh, -yh);
        nl[j] = make_float2(-xl, -yl);
        sx[j] = 0; sy[j] = 0; fx[j] = 0.0; fy[j] = 0.0; ni[j] = 0; no[j] = 0;
        mute[j] = (j < nactive) ? 0 : 0x7fffffff;     // surplus slots stay muted for ever
    }
    int retests = 0;
    const float th2 = w.th2f, soft2 = w.soft2f;
    // `zero` is a 0 the compiler cannot see through (loaded from memory): it keeps the base pointer in registers,
    // so that the record address is ONE IMAD.WIDE (p * 32 + base) instead of a constant-bank reload per visit
    const BhCell* __restrict__ cells;
    asm volatile("mad.wide.s32 %0, %1, 32, %2;" : "=l"(cells) : "r"(zero), "l"(t.cell));
    const int M = t.M;
    int p = nactive > 0 ? 0 : M;
    while (__any_sync(0xffffffffu, p < M)) {
#pragma unroll
        for (int k = 0; k < BH_WALK_CHUNK / (G > 2 ? 2 : 1); ++k) {
            float4 c0, c1;
            bh_load_cell(cells + p, &c0, &c1);
            const int skip = __float_as_int(c1.z);
            const int p1 = p + 1;
            int pnext = skip;
            const float2 ch = make_float2(c0.x, c0.y), cl = make_float2(c1.x, c1.y);
            float dx[G], dy[G], d2[G], tt[G], wg[G];
            bool border = false;
#pragma unroll
            for (int j = 0; j < G; ++j) {
                const float2 d = __fadd2_rn(__fadd2_rn(ch, nh[j]), __fadd2_rn(cl, nl[j]));   // (hi - hi) + (lo - lo), both axes
                dx[j] = d.x; dy[j] = d.y;
                d2[j] = fmaf(d.x, d.x, fmaf(d.y, d.y, soft2));
            }
            if constexpr (G == 1) {
                tt[0] = fmaf(d2[0], th2, -c0.w);
                wg[0] = bh_weight1(d2[0], c0.z);
                border = fabsf(tt[0]) <= c1.w;
            } else {
#pragma unroll
                for (int j = 0; j + 1 < G; j += 2) {      // two bodies per packed instruction
                    const float2 d2p = make_float2(d2[j], d2[j + 1]);
                    const float2 tp = __ffma2_rn(d2p, make_float2(th2, th2), make_float2(-c0.w, -c0.w));
                    const float2 wp = bh_weight2(d2p, c0.z);
                    tt[j] = tp.x; tt[j + 1] = tp.y; wg[j] = wp.x; wg[j + 1] = wp.y;
                    border |= fminf(fabsf(tp.x), fabsf(tp.y)) <= c1.w;
                }
            }
            if (border) {   // some body is inside the guard band: the reference's f64 test decides (cold)
#pragma unroll
                for (int j = 0; j < G; ++j)
                    if (fabsf(tt[j]) <= c1.w) {
                        tt[j] = bh_retest_cell(t.cd, t.sk, p, x[j], y[j], w.soft2, w.theta2, w.half) ? 1.0f : -1.0f;
                        retests += (p >= mute[j]);
                    }
            }
#pragma unroll
            for (int j = 0; j < G; ++j) {
                if constexpr (G == 1) {
                    if constexpr (ACC == BH_ACC_F64) { BH_MULTI_TAIL(j, BH_PRED_SINGLE, BH_MULTI_ACC_F64_ASM, BH_ACC_C_F64, BH_MOVE_SINGLE); }
                    else { BH_MULTI_TAIL(j, BH_PRED_SINGLE, BH_MULTI_ACC_FOLD_ASM, BH_ACC_C_F32, BH_MOVE_SINGLE); }
                } else {
                    if constexpr (ACC == BH_ACC_F64) { BH_MULTI_TAIL(j, BH_PRED_GROUP, BH_MULTI_ACC_F64_ASM, BH_ACC_C_F64, BH_MOVE_GROUP); }
                    else { BH_MULTI_TAIL(j, BH_PRED_GROUP, BH_MULTI_ACC_FOLD_ASM, BH_ACC_C_F32, BH_MOVE_GROUP); }
                }
            }
            p = pnext;
        }
        if constexpr (ACC == BH_ACC_FOLD) {
#pragma unroll
            for (int j = 0; j < G; ++j) { fx[j] += (double)sx[j]; fy[j] += (double)sy[j]; sx[j] = 0; sy[j] = 0; }
        }
    }
#pragma unroll
    for (int j = 0; j < G; ++j) {
        out->ax[j] = (ACC == BH_ACC_FOLD ? fx[j] : (double)sx[j]) * BH_WALK_WSCALE;     // the scale is a power of two: exact
        out->ay[j] = (ACC == BH_ACC_FOLD ? fy[j] : (double)sy[j]) * BH_WALK_WSCALE;
        out->interactions[j] = ni[j];
        out->opened[j] = no[j];
    }
    out->retests = retests;
}
#endif

// Potential of one body from the tree: the walk of bh_walk_body with m/sqrt(d^2+soft2) in place of
// the force (same per-body decisions: guarded FP32 test, f64 re-test in the band).  Diagnostics
// path (bh_energy_tree): plain code, FP32 terms folded into f64 every BH_WALK_CHUNK visits.
BH_HD double bh_walk_potential(const BhTreeView& t, const BhWalkParams& w, double x, double y, int self, bool active) {
    float xh, xl, yh, yl;
    bh_split(x, &xh, &xl);
    bh_split(y, &yh, &yl);
    const float th2 = w.th2f, soft2 = w.soft2f;
    const BhCell* __restrict__ cells = t.cell;
    const int M = t.M;
    int p = active ? 0 : M;
    double phi = 0.0;
    while (p < M) {
        float f = 0.f;
        for (int k = 0; k < BH_WALK_CHUNK && p < M; ++k) {
            const BhCell c = cells[p];
            const float dx = (c.xh - xh) + (c.xl - xl);
            const float dy = (c.yh - yh) + (c.yl - yl);
            const float d2 = fmaf(dx, dx, fmaf(dy, dy, soft2));
            const float tt = fmaf(d2, th2, -c.s2);
            bool accept = tt > 0.0f;
            if (fabsf(tt) <= c.band) accept = bh_retest_cell(t.cd, t.sk, p, x, y, w.soft2, w.theta2, w.half);
            if (accept && p != self && c.m != 0.0f) {
                float inv = BH_RSQRTF(d2);
#if defined(__CUDA_ARCH__)
                inv = inv * fmaf(-0.5f * d2, inv * inv, 1.5f);   // one Newton step on MUFU.RSQ
#endif
                f = fmaf(c.m, inv, f);
            }
            p = accept ? c.skip : p + 1;
        }
        phi += (double)f;
    }
    return phi;
}

#endif  // BH_CORE_H
