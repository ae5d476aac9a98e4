// bh_kernels.cuh — the CUDA kernels of the B200 Barnes–Hut step (sm_100a).  Per-thread logic
// lives in bh_core.h (shared with the CPU emulation used by the tests); this file adds the
// parallel structure: launch shapes, warp reductions, the look-back scan, staging.
#ifndef BH_KERNELS_CUH
#define BH_KERNELS_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include "bh_core.h"
#include "bh_sort.cuh"

namespace {

// ---------------------------------------------------------------------------------------
// device-side scalars
// ---------------------------------------------------------------------------------------
struct DevScalars {           // zeroed at the start of every build
    int n_in;                 // bodies that passed the root contains() test
    int n_internal;           // internal cells
    int n_jitter;             // bodies sharing a cell with h < 1e-3
    int max_depth;
    unsigned long long interactions, opened, retests;   // of the evaluation that follows
    unsigned int scan_ticket;
    unsigned int pad;         // always 0: the walk adds it to its base pointer (see bh_walk_multi)
    int n_ghost;              // bodies dropped by the jitter replay (ghost leaves)
    int jitter_unsupported;   // a jittered body survived below depth levels+1 (cannot happen for h < 1e-3)
    // bounding box of all bodies as order-preserving keys (bh_ord_key): max of key(x), max of ~key(x), same for y;
    // all-zero = no body seen (the scalars are zeroed per build)
    unsigned long long bb[4];
};
// order-preserving map double -> uint64 (and back): a < b  <=>  key(a) < key(b); key is never 0 for a finite value
__host__ __device__ inline unsigned long long bh_ord_key(double v) {
    unsigned long long b;
#if defined(__CUDA_ARCH__)
    b = (unsigned long long)__double_as_longlong(v);
#else
    memcpy(&b, &v, 8);
#endif
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
inline double bh_ord_unkey(unsigned long long k) {
    const unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
    double v;
    memcpy(&v, &b, 8);
    return v;
}
struct DevTotals {            // zeroed by bh_reset_counters only
    unsigned long long interactions, opened, retests, evaluations;
};

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_IPT = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_IPT;

// Decoupled look-back of a single-pass scan, executed by ONE warp of the tile: publishes this
// tile's aggregate, sums the aggregates of the tiles before it (32 predecessors per round) and
// publishes the inclusive prefix.  Returns the exclusive prefix of the tile.  Tiles take their
// index from an atomic ticket, so a tile only waits on tiles that are already running.
__device__ __forceinline__ int bh_tile_lookback(uint32_t* __restrict__ status, int tile, int total, int lane) {
    int excl = 0;
    if (tile == 0) {
        if (lane == 0) bhsort::st_volatile_u32(status, bhsort::FLAG_PREFIX | (uint32_t)total);
        return 0;
    }
    if (lane == 0) bhsort::st_volatile_u32(status + tile, bhsort::FLAG_AGG | (uint32_t)total);
    int t = tile - 1;
    for (;;) {
        const int idx = t - lane;
        const uint32_t v = (idx >= 0) ? bhsort::ld_volatile_u32(status + idx) : bhsort::FLAG_PREFIX;
        const uint32_t f = v >> bhsort::FLAG_SHIFT;
        const unsigned pref = __ballot_sync(0xffffffffu, f == 2);
        const unsigned inval = __ballot_sync(0xffffffffu, f == 0);
        const unsigned window = pref ? ((2u << (__ffs(pref) - 1)) - 1u) : 0xffffffffu;
        if (inval & window) continue;   // some needed predecessor has not published yet
        int contrib = ((window >> lane) & 1u) ? (int)(v & bhsort::VALUE_MASK) : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
        excl += contrib;
        if (pref) break;
        t -= 32;
    }
    if (lane == 0) bhsort::st_volatile_u32(status + tile, bhsort::FLAG_PREFIX | (uint32_t)(excl + total));
    return excl;
}

// ---------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------

// Morton keys by literal descent (BH.kt:153-155, :73-80) + root contains() (BH.kt:126).
// HBM-bound: 16 B read + 8 B written per body.
__global__ void __launch_bounds__(256) k_keygen(const double* __restrict__ x, const double* __restrict__ y, int n,
                                                BhRoot root, BhGrid grid, uint64_t sentinel, uint64_t* __restrict__ keys,
                                                DevScalars* __restrict__ sc, int ell = -1, uint32_t c_lo = 0, uint32_t c_hi = 0) {
    // grid-stride: ONE atomic per block for the in-box count (same-address atomics serialise in L2)
    int cnt = 0;
    unsigned long long bb0 = 0, bb1 = 0, bb2 = 0, bb3 = 0;   // running max of key(x), ~key(x), key(y), ~key(y)
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < n; b += gridDim.x * blockDim.x) {
        const double px = x[b], py = y[b];
        if (px == px && py == py) {                          // bounding box of every body, in the box or not
            const unsigned long long kx = bh_ord_key(px), ky = bh_ord_key(py);
            bb0 = kx > bb0 ? kx : bb0; bb1 = ~kx > bb1 ? ~kx : bb1;
            bb2 = ky > bb2 ? ky : bb2; bb3 = ~ky > bb3 ? ~ky : bb3;
        }
        bool in = bh_root_contains(root, px, py);
        // closed form on the exact grid when the root box allows it, else the literal descent
        uint64_t key = in ? (grid.exact ? bh_morton_key_grid(grid, root.levels, px, py) : bh_morton_key(root, px, py)) : sentinel;
        if (ell >= 0 && in) {   // locally essential tree: only the bodies of this rank's code range enter its tree
            const uint32_t c = (uint32_t)(key >> (2 * (root.levels - ell)));
            if (c < c_lo || c >= c_hi) { in = false; key = sentinel; }
        }
        keys[b] = key;
        cnt += in;
    }
    if (sc) {
        __shared__ int s_cnt;
        __shared__ unsigned long long s_bb[4];
        if (threadIdx.x == 0) { s_cnt = 0; s_bb[0] = s_bb[1] = s_bb[2] = s_bb[3] = 0; }
        __syncthreads();
        int v = cnt;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {                   // warp-shuffle reductions: count and the four extrema
            v += __shfl_xor_sync(0xffffffffu, v, o);
            unsigned long long t;
            t = __shfl_xor_sync(0xffffffffu, bb0, o); bb0 = t > bb0 ? t : bb0;
            t = __shfl_xor_sync(0xffffffffu, bb1, o); bb1 = t > bb1 ? t : bb1;
            t = __shfl_xor_sync(0xffffffffu, bb2, o); bb2 = t > bb2 ? t : bb2;
            t = __shfl_xor_sync(0xffffffffu, bb3, o); bb3 = t > bb3 ? t : bb3;
        }
        if ((threadIdx.x & 31) == 0) {
            if (v) atomicAdd(&s_cnt, v);
            atomicMax(&s_bb[0], bb0); atomicMax(&s_bb[1], bb1); atomicMax(&s_bb[2], bb2); atomicMax(&s_bb[3], bb3);
        }
        __syncthreads();
        if (threadIdx.x == 0 && s_cnt) atomicAdd(&sc->n_in, s_cnt);
        if (threadIdx.x < 4 && s_bb[threadIdx.x]) atomicMax(&sc->bb[threadIdx.x], s_bb[threadIdx.x]);   // one atomic per block each
    }
}

// cnt(i) = max(0, delta(i) - delta(i-1)) and its exclusive scan S (single pass, decoupled
// look-back), plus tree statistics.  HBM-bound: 8 B read + 4 B written per in-tree body.
__global__ void __launch_bounds__(SCAN_THREADS)
k_count_scan(const uint64_t* __restrict__ keys, int levels, DevScalars* __restrict__ sc, int* __restrict__ S,
             uint32_t* __restrict__ status) {
    __shared__ uint32_t s_tile;
    __shared__ int s_warp[SCAN_THREADS / 32];
    __shared__ int s_tile_excl;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(&sc->scan_ticket, 1u);
    __syncthreads();
    const int tile = (int)s_tile;
    const int n = sc->n_in;
    const int64_t base = (int64_t)tile * SCAN_TILE;
    if (n == 0) { if (tile == 0 && tid == 0) S[0] = 0; return; }
    if (base >= n) return;

    const int64_t i0 = base + (int64_t)tid * SCAN_IPT;
    uint64_t kk[SCAN_IPT + 2];
#pragma unroll
    for (int j = 0; j < SCAN_IPT + 2; ++j) {
        const int64_t idx = i0 - 1 + j;
        kk[j] = (idx >= 0 && idx < n) ? keys[idx] : 0ull;
    }
    int c[SCAN_IPT];
    int sum = 0, jit = 0, maxd = 0;
    int dprev = (i0 >= 1 && i0 < n) ? bh_common_levels(kk[0], kk[1], levels) : -1;
#pragma unroll
    for (int j = 0; j < SCAN_IPT; ++j) {
        const int64_t i = i0 + j;
        c[j] = 0;
        if (i < n) {
            const int dnext = (i + 1 < n) ? bh_common_levels(kk[j + 1], kk[j + 2], levels) : -1;
            c[j] = dnext > dprev ? dnext - dprev : 0;
            const int dep = (dprev > dnext ? dprev : dnext) + 1;
            maxd = dep > maxd ? dep : maxd;
            jit += (dprev == levels || dnext == levels);
            dprev = dnext;
        }
        sum += c[j];
    }
    // block exclusive scan of the thread sums
    int inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[w] = inc;
    __syncthreads();
    int wbase = 0, total = 0;
#pragma unroll
    for (int k = 0; k < SCAN_THREADS / 32; ++k) { const int v = s_warp[k]; if (k < w) wbase += v; total += v; }
    const int thread_excl = wbase + inc - sum;

    if (w == 0) {
        const int excl = bh_tile_lookback(status, tile, total, lane);
        if (lane == 0) s_tile_excl = excl;
    }
    __syncthreads();
    int run = s_tile_excl + thread_excl;
#pragma unroll
    for (int j = 0; j < SCAN_IPT; ++j) {
        const int64_t i = i0 + j;
        if (i < n) {
            S[i] = run;
            run += c[j];
            if (i == n - 1) { S[n] = run; sc->n_internal = run; }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        jit += __shfl_xor_sync(0xffffffffu, jit, o);
        const int om = __shfl_xor_sync(0xffffffffu, maxd, o);
        maxd = om > maxd ? om : maxd;
    }
    // one atomic per block (same-address atomics serialise in L2)
    __shared__ int s_stat[2];
    if (tid == 0) { s_stat[0] = 0; s_stat[1] = 0; }
    __syncthreads();
    if (lane == 0) {
        if (jit) atomicAdd(&s_stat[0], jit);
        atomicMax(&s_stat[1], maxd);
    }
    __syncthreads();
    if (tid == 0) {
        if (s_stat[0]) atomicAdd(&sc->n_jitter, s_stat[0]);
        atomicMax(&sc->max_depth, s_stat[1]);
    }
}

// cell skeletons (skip / parent / count / level) — bh_emit_body per in-tree body.  (Staging a block's keys and
// scan values in shared memory for the galloping searches was measured: 81 us against 64 us for the plain
// L2-resident arrays at 1M bodies, so the searches read the global arrays.)
__global__ void __launch_bounds__(256) k_emit(BhTreeView t, int levels) {
    bh_view_resolve(t);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < t.n_in) bh_emit_body(t, levels, i);
}

// computeMass (BH.kt:173-202), block-local form.  A block owns CLIMB_B consecutive sorted bodies,
// i.e. the contiguous preorder range [P0, P1) of the cells those bodies own.  A cell of that range
// whose subtree ends inside it (skip <= P1) is LOCAL: ~97 % of the internal cells hold <= 32 bodies.
// Local cells are computed entirely in shared memory with the same arrive-counter protocol as
// bh_climb_from (the last arriving child sums the children 0..3 in order, the reference's f64
// expression order) but with shared-memory atomics and block-scope fences instead of acq_rel
// round trips to L2, and their records are written to HBM coalesced.  Only the roots of the local
// subtrees continue through the global protocol (cells that span blocks).  A block whose range has
// more than CLIMB_CAP cells falls back to bh_climb_body per thread.
constexpr int CLIMB_B = 256;
constexpr int CLIMB_CAP = 1024;
struct alignas(16) BhClimbRoot { BhCellS s; uint64_t key; int carry, pad; };   // a finished local subtree
__global__ void __launch_bounds__(CLIMB_B)
k_climb_block(BhTreeView t, BhRoot root, const double* __restrict__ x, const double* __restrict__ y,
              const double* __restrict__ m, const int* __restrict__ jflag, int* __restrict__ leafpos,
              BhClimbRoot* __restrict__ roots, int* __restrict__ n_roots) {
    __shared__ BhCellS s_sk[CLIMB_CAP];
    __shared__ double s_m[CLIMB_CAP], s_x[CLIMB_CAP], s_y[CLIMB_CAP];
    __shared__ int s_arr[CLIMB_CAP];
    bh_view_resolve(t);
    const int tid = threadIdx.x;
    const int b0 = blockIdx.x * CLIMB_B;
    if (b0 >= t.n_in) return;                    // (sync-free builds launch one block per 256 BODIES, in the tree or not)
    const int b1 = min(b0 + CLIMB_B, t.n_in);
    const int i = b0 + tid;
    const int P0 = t.S[b0] + b0, P1 = t.S[b1] + b1;
    const int nc = P1 - P0;
    if (blockIdx.x == 0 && tid == 0) bh_write_terminal_cell(t);
    int lp = 0;
    double bx = 0.0, by = 0.0, bm = 0.0;
    uint64_t key = 0;
    if (i < b1) {
        const int b = t.order[i];
        lp = t.S[i + 1] + i;
        leafpos[b] = lp;
        bx = x[b]; by = y[b];
        bm = (jflag && (jflag[b] & 1)) ? 0.0 : m[b];    // ghost leaf of a body the jitter replay dropped
        key = t.keys[i];
    }
    if (nc > CLIMB_CAP) {                               // unusually deep range: per-thread global climb
        if (i < b1) bh_climb_body(t, root, i, bx, by, bm);
        return;
    }
    for (int c = tid; c < nc; c += CLIMB_B) { s_sk[c] = t.sk[P0 + c]; s_arr[c] = 0; }
    __syncthreads();
    // local climb; a thread that finishes the root of a local subtree keeps (rs, rcarry) for phase 3
    BhCellS rs; rs.parent = -1; rs.skip = 0; rs.cnt = 0; rs.level = 0;
    int rcarry = 0;
    if (i < b1) {
        int p = lp - P0;
        BhCellS s = s_sk[p];
        s_m[p] = bm; s_x[p] = bx; s_y[p] = by;
        int carry = 1;
        for (;;) {
            const int q = s.parent;
            if (q < 0) break;                                         // the root of the whole tree
            if (q < P0 || s_sk[q - P0].skip > P1) { rs = s; rcarry = carry; break; }   // parent spans blocks
            const BhCellS sq = s_sk[q - P0];
            __threadfence_block();
            const int old = atomicAdd(&s_arr[q - P0], carry);
            if (old + carry != sq.cnt) break;
            __threadfence_block();
            double mSum = 0.0, sx = 0.0, sy = 0.0;
            const int end = sq.skip - P0;
            for (int ch = q - P0 + 1; ch < end; ch = s_sk[ch].skip - P0) {
                const double mc = s_m[ch];
                if (mc > 0.0) {   // BH.kt:189-192
                    mSum = __dadd_rn(mSum, mc);
                    sx = __dadd_rn(sx, __dmul_rn(s_x[ch], mc));
                    sy = __dadd_rn(sy, __dmul_rn(s_y[ch], mc));
                }
            }
            double cx, cy;
            if (mSum > 0.0) { cx = __ddiv_rn(sx, mSum); cy = __ddiv_rn(sy, mSum); }     // BH.kt:194-196
            else { double h; bh_cell_geometry(root, key, sq.level, &cx, &cy, &h); }    // BH.kt:197-200
            s_m[q - P0] = mSum; s_x[q - P0] = cx; s_y[q - P0] = cy;
            carry = sq.cnt;
            s = sq;
        }
    }
    __syncthreads();
    // the finished local cells, coalesced
    for (int c = tid; c < nc; c += CLIMB_B) {
        const BhCellS s = s_sk[c];
        if (s.skip <= P1) bh_write_cell(t, P0 + c, s_x[c], s_y[c], s_m[c], s.skip, s.level, s.skip == P0 + c + 1, root.half);
    }
    // The roots of the local subtrees still have to report to their block-spanning parents.  That
    // is a chain of L2 round trips per level; doing it here would pin the block's shared memory
    // for its whole length, so the roots are queued for k_climb_top instead (one slot reservation
    // per block).
    __shared__ int s_nroot, s_base_root;
    if (tid == 0) s_nroot = 0;
    __syncthreads();
    int slot = -1;
    if (rcarry > 0) slot = atomicAdd(&s_nroot, 1);
    __syncthreads();
    if (tid == 0 && s_nroot) s_base_root = atomicAdd(n_roots, s_nroot);
    __syncthreads();
    if (slot >= 0) {
        BhClimbRoot r; r.s = rs; r.key = key; r.carry = rcarry; r.pad = 0;
        roots[s_base_root + slot] = r;
    }
}

// The block-spanning top of the tree: every queued local root climbs with the global
// arrive-counter protocol (bh_climb_from).  Runs after k_climb_block, so all local cells are visible.
__global__ void __launch_bounds__(128)
k_climb_top(BhTreeView t, BhRoot root, const BhClimbRoot* __restrict__ roots, const int* __restrict__ n_roots) {
    bh_view_resolve(t);
    const int n = *n_roots;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const BhClimbRoot r = roots[k];
        bh_climb_from(t, root, r.key, r.s, r.carry);
    }
}

// accumulateForce (BH.kt:215-239) + ax = fx/m (BH.kt:390-391): one thread per G target bodies
// (bh_walk_multi).  Targets are the bodies [first_target, first_target + n_targets) in HOME order, which is
// the Morton order of the last re-homing, so the bodies of a thread and the lanes of a warp are spatial
// neighbours and the body reads / acceleration writes are coalesced.  Stackless over the preorder cells.
//
// PERSISTENT, SM-AFFINE schedule.  ncu on a plain one-block-per-chunk launch: the walk waits on L1 misses
// (a warp-wide record load touches ~16 sectors; at a 93 % sector hit rate most of those loads contain a
// miss, i.e. an L2 or DRAM round trip on the pointer chase), because the blocks resident on one SM come
// from all over the box.  Here the targets are cut into as many contiguous RANGES as there are SMs and
// every warp takes its work (chunks of 32*G bodies, one atomic each) from the range of the SM it runs on
// (%smid), so that all warps of an SM walk the same neighbourhood of the tree and share its cells in L1.
// A warp whose range is exhausted STEALS: the lanes read all range counters at once, and the warp moves to
// the next range (cyclically) that still has chunks — which also covers SMs that got no block of a small
// grid and %smid values that skip numbers.  Work is handed out per WARP: no block-wide barriers.
struct BhWalkQueue {
    unsigned int* next;     // [n_ranges] counters, zero before the launch
    int n_ranges;           // = SMs of the device (<= 1024)
    int per;                // chunks per range (the last ranges may be short or empty)
    int chunks;             // chunks in total
};
__device__ __forceinline__ int bh_walk_range_len(const BhWalkQueue& q, int r) {
    const int left = q.chunks - r * q.per;
    return left < 0 ? 0 : (left > q.per ? q.per : left);
}

template <int G, int ACC, int MINB>
__global__ void __launch_bounds__(128, MINB)
k_walk(BhTreeView t, BhWalkParams w, int first_target, int n_targets, const double* __restrict__ x,
       const double* __restrict__ y, const double* __restrict__ m, const int* __restrict__ leafpos, double Gc,
       double* __restrict__ ax, double* __restrict__ ay, int* __restrict__ cntI, int* __restrict__ cntO,
       DevScalars* __restrict__ sc, DevTotals* __restrict__ tot, BhWalkQueue q) {
    bh_view_resolve(t);
    const int lane = threadIdx.x & 31;
    unsigned int smid;
    asm("mov.u32 %0, %%smid;" : "=r"(smid));
    int my = (int)(smid % (unsigned int)q.n_ranges);
    const int zero = (int)sc->pad;
    long long ni = 0, no = 0;
    int nr = 0;
    for (;;) {
        int c = 0;
        if (lane == 0) c = (int)atomicAdd(&q.next[my], 1u);
        c = __shfl_sync(0xffffffffu, c, 0);
        if (c >= bh_walk_range_len(q, my)) {
            // this range is exhausted: the next one (cyclically) with unclaimed chunks, or done.  Lane l looks at the
            // ranges my+1+l, my+1+l+32, ...; a range seen as exhausted stays exhausted, one seen as open may be
            // exhausted by the time of the atomic above — then the scan simply runs again.
            int found = -1;
            for (int base = 0; base < q.n_ranges && found < 0; base += 32) {
                const int k = base + lane;
                int r = my + 1 + k;
                if (r >= q.n_ranges) r -= q.n_ranges;
                const bool open = k < q.n_ranges - 1 && *(volatile unsigned int*)&q.next[r] < (unsigned int)bh_walk_range_len(q, r);
                const unsigned int bal = __ballot_sync(0xffffffffu, open);
                if (bal) { found = my + 1 + base + (__ffs(bal) - 1); if (found >= q.n_ranges) found -= q.n_ranges; }
            }
            if (found < 0) break;
            my = found;
            continue;
        }
        const int chunk = my * q.per + c;
        const int rel = (chunk * 32 + lane) * G;
        int nactive = n_targets - rel;
        nactive = nactive < 0 ? 0 : (nactive > G ? G : nactive);
        double bx[G], by[G];
        int self[G];
#pragma unroll
        for (int j = 0; j < G; ++j) {
            const int b = first_target + (j < nactive ? rel + j : 0);
            bx[j] = x[b]; by[j] = y[b]; self[j] = leafpos[b];
        }
        // every lane enters the walk (it contains full-warp votes); surplus lanes idle on the terminal record
        BhMultiResult<G> r;
        bh_walk_multi<G, ACC>(t, w, bx, by, self, nactive, zero, &r);
        nr += r.retests;
#pragma unroll
        for (int j = 0; j < G; ++j) {
            if (j < nactive) {
                const int b = first_target + rel + j;
                const double mb = m[b];
                // BH.kt:390-391 divides the force by b.m: a zero-mass body gets 0/0 = NaN
                ax[b] = (mb == 0.0) ? nan("") : Gc * r.ax[j];
                ay[b] = (mb == 0.0) ? nan("") : Gc * r.ay[j];
                ni += r.interactions[j]; no += r.opened[j];
                if (cntI) { cntI[b] = r.interactions[j]; cntO[b] = r.opened[j]; }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ni += __shfl_xor_sync(0xffffffffu, ni, o);
        no += __shfl_xor_sync(0xffffffffu, no, o);
        nr += __shfl_xor_sync(0xffffffffu, nr, o);
    }
    if (lane == 0 && (ni | no | nr)) {
        atomicAdd(&sc->interactions, (unsigned long long)ni);
        atomicAdd(&sc->opened, (unsigned long long)no);
        atomicAdd(&tot->interactions, (unsigned long long)ni);
        atomicAdd(&tot->opened, (unsigned long long)no);
        if (nr) { atomicAdd(&sc->retests, (unsigned long long)nr); atomicAdd(&tot->retests, (unsigned long long)nr); }
    }
}

// BH.kt:411-422 / :429-432 in f64 with the reference's rounding (no FMA contraction):
//   v += a * dtHalf ; if (drift) x += v * dt
__global__ void __launch_bounds__(256)
k_kick_drift(int lo, int hi, double* __restrict__ x, double* __restrict__ y, double* __restrict__ vx,
             double* __restrict__ vy, const double* __restrict__ ax, const double* __restrict__ ay, double dtHalf,
             double dt, int drift) {
    const int i = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hi) return;
    const double nvx = __dadd_rn(vx[i], __dmul_rn(ax[i], dtHalf));
    const double nvy = __dadd_rn(vy[i], __dmul_rn(ay[i], dtHalf));
    vx[i] = nvx; vy[i] = nvy;
    if (drift) {
        x[i] = __dadd_rn(x[i], __dmul_rn(nvx, dt));
        y[i] = __dadd_rn(y[i], __dmul_rn(nvy, dt));
    }
}

// Tiled all-pairs direct sum (accuracy oracle): FP32 interaction math on (hi,lo) split
// coordinates, per-tile FP32 partial sums folded into f64 accumulators.
constexpr int DS_TILE = 256;
__global__ void __launch_bounds__(DS_TILE)
k_direct(const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ m, int n,
         float soft2f, double G, double* __restrict__ ax, double* __restrict__ ay) {
    __shared__ float4 sA[DS_TILE];   // xh, yh, m, -
    __shared__ float2 sB[DS_TILE];   // xl, yl
    const int i = blockIdx.x * DS_TILE + threadIdx.x;
    float xh = 0.f, xl = 0.f, yh = 0.f, yl = 0.f;
    if (i < n) { bh_split(x[i], &xh, &xl); bh_split(y[i], &yh, &yl); }
    double accx = 0.0, accy = 0.0;
    for (int t0 = 0; t0 < n; t0 += DS_TILE) {
        const int j = t0 + threadIdx.x;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        float2 b = make_float2(0.f, 0.f);
        if (j < n) { bh_split(x[j], &a.x, &b.x); bh_split(y[j], &a.y, &b.y); a.z = (float)m[j]; }
        __syncthreads();
        sA[threadIdx.x] = a; sB[threadIdx.x] = b;
        __syncthreads();
        float fx = 0.f, fy = 0.f;
#pragma unroll 8
        for (int k = 0; k < DS_TILE; ++k) {
            const float4 s = sA[k];
            const float2 l = sB[k];
            const float dx = (s.x - xh) + (l.x - xl);
            const float dy = (s.y - yh) + (l.y - yl);
            const float r2 = fmaf(dx, dx, fmaf(dy, dy, soft2f));
            float inv = rsqrtf(r2);
            inv = inv * fmaf(-0.5f * r2, inv * inv, 1.5f);   // one Newton step
            const float wgt = (r2 > 0.f) ? s.z * inv * inv * inv : 0.f;
            fx = fmaf(wgt, dx, fx);
            fy = fmaf(wgt, dy, fy);
        }
        accx += (double)fx; accy += (double)fy;
    }
    if (i < n) {
        const double mb = m[i];
        ax[i] = (mb == 0.0) ? nan("") : G * accx;
        ay[i] = (mb == 0.0) ? nan("") : G * accy;
    }
}

// energy / momentum diagnostics in f64.  out[0]=KE out[1]=sum m_i u_i out[2]=px out[3]=py
__global__ void __launch_bounds__(DS_TILE)
k_energy(const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ vx,
         const double* __restrict__ vy, const double* __restrict__ m, int n, double soft2, double* __restrict__ out) {
    __shared__ double sx[DS_TILE], sy[DS_TILE], sm[DS_TILE];
    __shared__ double red[4][DS_TILE / 32];
    const int i = blockIdx.x * DS_TILE + threadIdx.x;
    const double xi = i < n ? x[i] : 0.0, yi = i < n ? y[i] : 0.0;
    double u = 0.0;
    for (int t0 = 0; t0 < n; t0 += DS_TILE) {
        const int j = t0 + threadIdx.x;
        __syncthreads();
        sx[threadIdx.x] = j < n ? x[j] : 0.0;
        sy[threadIdx.x] = j < n ? y[j] : 0.0;
        sm[threadIdx.x] = j < n ? m[j] : 0.0;
        __syncthreads();
#pragma unroll 4
        for (int k = 0; k < DS_TILE; ++k) {
            const double dx = sx[k] - xi, dy = sy[k] - yi;
            const double r2 = dx * dx + dy * dy + soft2;
            u += (t0 + k != i && r2 > 0.0) ? sm[k] * rsqrt(r2) : 0.0;
        }
    }
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    if (i < n) {
        const double mi = m[i];
        v[0] = 0.5 * mi * (vx[i] * vx[i] + vy[i] * vy[i]);
        v[1] = mi * u;
        v[2] = mi * vx[i];
        v[3] = mi * vy[i];
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        double t = v[q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) red[q][w] = t;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double t = 0.0;
        for (int k = 0; k < DS_TILE / 32; ++k) t += red[threadIdx.x][k];
        atomicAdd(&out[threadIdx.x], t);
    }
}

// bh_energy_tree: phi_i from the tree, then out[0]=KE out[1]=sum m_i phi_i out[2]=px out[3]=py (f64)
__global__ void __launch_bounds__(128)
k_energy_tree(BhTreeView t, BhWalkParams w, int n, const double* __restrict__ x, const double* __restrict__ y,
              const double* __restrict__ vx, const double* __restrict__ vy, const double* __restrict__ m,
              const int* __restrict__ leafpos, double* __restrict__ out) {
    __shared__ double red[4][4];
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    if (b < n) {
        const double phi = bh_walk_potential(t, w, x[b], y[b], leafpos[b], true);
        const double mi = m[b];
        v[0] = 0.5 * mi * (vx[b] * vx[b] + vy[b] * vy[b]);
        v[1] = mi * phi;
        v[2] = mi * vx[b];
        v[3] = mi * vy[b];
    }
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        double s = v[q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) red[q][wp] = s;
    }
    __syncthreads();
    if (threadIdx.x < 4) atomicAdd(&out[threadIdx.x], red[threadIdx.x][0] + red[threadIdx.x][1] + red[threadIdx.x][2] + red[threadIdx.x][3]);
}

// render read-back in USER order: xy[perm[i]] = (x, y), mf[perm[i]] = m
__global__ void k_positions_f32(const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ m,
                                const int* __restrict__ perm, int n, float2* __restrict__ xy, float* __restrict__ mf) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { const int u = perm[i]; xy[u] = make_float2((float)x[i], (float)y[i]); mf[u] = (float)m[i]; }
}

// depth of each body's leaf (-1: not in the tree), home order
__global__ void k_leaf_depth(BhTreeView t, const int* __restrict__ leafpos, const int* __restrict__ jflag, int n,
                             int* __restrict__ depth) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n) return;
    const int lp = leafpos[b];
    depth[b] = (lp >= 0 && !(jflag && (jflag[b] & 1))) ? t.sk[lp].level : -1;
}

// Jitter regime (BH.kt:145-156): the thread at the first key of every run of equal keys replays
// that cluster sequentially (bh_jitter_cluster).  Launched only when the scan found such runs.
__global__ void __launch_bounds__(128)
k_jitter(const uint64_t* __restrict__ keys, int* __restrict__ order, int n_in, BhRoot root, const int* __restrict__ perm,
         double* __restrict__ x, double* __restrict__ y, int* __restrict__ jflag, DevScalars* __restrict__ sc) {
    if (n_in < 0) {                              // sync-free build: counts on the device; nothing to do without equal keys
        if (sc->n_jitter == 0) return;
        n_in = sc->n_in;
    }
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_in - 1) return;
    const uint64_t k = keys[i];
    if (keys[i + 1] != k || (i > 0 && keys[i - 1] == k)) return;
    int j = i + 1;
    while (j + 1 < n_in && keys[j + 1] == k) ++j;
    int unsupported = 0, ghosts = 0;
    bh_jitter_cluster(root, k, order + i, j - i + 1, perm, x, y, jflag, &unsupported, &ghosts);
    if (ghosts) atomicAdd(&sc->n_ghost, ghosts);
    if (unsupported) atomicExch(&sc->jitter_unsupported, 1);
}

// ---- permutation helpers (home order <-> user order, re-homing) -----------------------------
__global__ void k_iota(int* __restrict__ a, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = i;
}
template <class T>
__global__ void k_gather(T* __restrict__ dst, const T* __restrict__ src, const int* __restrict__ idx, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}
template <class T>
__global__ void k_scatter(T* __restrict__ dst, const T* __restrict__ src, const int* __restrict__ idx, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[idx[i]] = src[i];
}

// ---- generic single-pass exclusive scan (decoupled look-back) --------------------------------
// out[i] = sum of in[0..i), out[n] = total.  `ticket` and `status` (tiles words) must be zero.
__global__ void __launch_bounds__(SCAN_THREADS)
k_excl_scan(const int* __restrict__ in, int n, int* __restrict__ out, uint32_t* __restrict__ ticket,
            uint32_t* __restrict__ status) {
    __shared__ uint32_t s_tile;
    __shared__ int s_warp[SCAN_THREADS / 32];
    __shared__ int s_tile_excl;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    __syncthreads();
    const int tile = (int)s_tile;
    const int64_t base = (int64_t)tile * SCAN_TILE;
    if (base >= n) { if (n == 0 && tile == 0 && tid == 0) out[0] = 0; return; }
    const int64_t i0 = base + (int64_t)tid * SCAN_IPT;
    int c[SCAN_IPT];
    int sum = 0;
#pragma unroll
    for (int j = 0; j < SCAN_IPT; ++j) { c[j] = (i0 + j < n) ? in[i0 + j] : 0; sum += c[j]; }
    int inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[w] = inc;
    __syncthreads();
    int wbase = 0, total = 0;
#pragma unroll
    for (int k = 0; k < SCAN_THREADS / 32; ++k) { const int v = s_warp[k]; if (k < w) wbase += v; total += v; }
    if (w == 0) {
        const int excl = bh_tile_lookback(status, tile, total, lane);
        if (lane == 0) s_tile_excl = excl;
    }
    __syncthreads();
    int run = s_tile_excl + wbase + inc - sum;
#pragma unroll
    for (int j = 0; j < SCAN_IPT; ++j) {
        const int64_t i = i0 + j;
        if (i < n) { out[i] = run; run += c[j]; if (i == n - 1) out[n] = run; }
    }
}

// ---- merge ("devour") rule, BH.kt:463-532 ----------------------------------------------------
// All index arithmetic is in USER order (the order of the reference's `bodies` list): inv[u] is the
// home slot of user index u.
__global__ void k_invert_perm(const int* __restrict__ perm, int n, int* __restrict__ inv) {
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h < n) inv[perm[h]] = h;
}
// flag[u] = bodies[u].m > mergeMaxMass  (BH.kt:472, strict)
__global__ void k_merge_flag_heavy(const double* __restrict__ m, const int* __restrict__ inv, int n, double max_mass,
                                   int* __restrict__ flag) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u < n) flag[u] = m[inv[u]] > max_mass ? 1 : 0;
}
// heavy[k] = home slot of the k-th heavy body in ascending user index
__global__ void k_merge_list_heavy(const int* __restrict__ flag, const int* __restrict__ scan, const int* __restrict__ inv,
                                   int n, int* __restrict__ heavy) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u < n && flag[u]) heavy[scan[u]] = inv[u];
}
// Candidate victims (BH.kt:488-499): every (heavy k, body j != heavy) with
// dx*dx + dy*dy < minD2 in the reference's f64 expression.  key = k<<32 | ~user(j), val = home(j):
// sorting ascending gives, per heavy in processing order, its victims in DESCENDING user index
// (BH.kt:514 sortedDescending).  *count may exceed cap (overflow is detected by the host).
constexpr int MERGE_TILE = 256;
__global__ void __launch_bounds__(MERGE_TILE)
k_merge_candidates(const double* __restrict__ x, const double* __restrict__ y, const int* __restrict__ perm, int n,
                   const int* __restrict__ heavy, int n_heavy, double minD2, uint64_t* __restrict__ keys,
                   uint32_t* __restrict__ vals, int cap, unsigned int* __restrict__ count) {
    __shared__ double hx[MERGE_TILE], hy[MERGE_TILE];
    __shared__ int hh[MERGE_TILE];
    const int j = blockIdx.x * MERGE_TILE + threadIdx.x;
    const double xj = j < n ? x[j] : 0.0, yj = j < n ? y[j] : 0.0;
    for (int k0 = 0; k0 < n_heavy; k0 += MERGE_TILE) {
        const int kk = k0 + threadIdx.x;
        __syncthreads();
        if (kk < n_heavy) { const int h = heavy[kk]; hh[threadIdx.x] = h; hx[threadIdx.x] = x[h]; hy[threadIdx.x] = y[h]; }
        __syncthreads();
        const int lim = min(MERGE_TILE, n_heavy - k0);
        if (j < n) {
            for (int t = 0; t < lim; ++t) {
                if (hh[t] == j) continue;                                   // j != i, BH.kt:491
                const double dx = __dsub_rn(xj, hx[t]), dy = __dsub_rn(yj, hy[t]);
                if (__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)) < minD2) {
                    const unsigned int slot = atomicAdd(count, 1u);
                    if (slot < (unsigned int)cap) {
                        keys[slot] = ((uint64_t)(k0 + t) << 32) | (uint64_t)(0xFFFFFFFFu - (uint32_t)perm[j]);
                        vals[slot] = (uint32_t)j;
                    }
                }
            }
        }
    }
}
// Sequential application (one thread: the f64 sums must run in the reference's order).
// For each heavy in ascending user index that is still alive, absorb the MASS ONLY (BH.kt:518) of
// each still-alive candidate, descending user index; victims are marked dead.
// `slot[c]` is the index of sorted candidate c in the unsorted candidate list `cand_home`.
__global__ void k_merge_apply(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ slot,
                              const uint32_t* __restrict__ cand_home, int n_cand, const int* __restrict__ heavy,
                              double* __restrict__ m, int* __restrict__ dead, int* __restrict__ n_dead) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    int nd = 0;
    int cur_k = -1, cur_h = -1;
    double cur_m = 0.0;
    for (int c = 0; c < n_cand; ++c) {
        const int k = (int)(keys[c] >> 32);
        if (k != cur_k) {
            if (cur_h >= 0) m[cur_h] = cur_m;
            cur_k = k; cur_h = heavy[k];
            if (dead[cur_h]) cur_h = -1; else cur_m = m[cur_h];
        }
        if (cur_h < 0) continue;                      // this heavy was eaten by an earlier one
        const int j = (int)cand_home[slot[c]];
        if (dead[j]) continue;                        // already removed from the list
        cur_m = __dadd_rn(cur_m, m[j]);               // bi.m += bj.m
        dead[j] = 1;
        ++nd;
    }
    if (cur_h >= 0) m[cur_h] = cur_m;
    *n_dead = nd;
}
// alive flags in home order and in user order (for the two stable compactions)
__global__ void k_merge_alive(const int* __restrict__ dead, const int* __restrict__ inv, int n, int* __restrict__ alive_home,
                              int* __restrict__ alive_user) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { alive_home[i] = dead[i] ? 0 : 1; alive_user[i] = dead[inv[i]] ? 0 : 1; }
}
// bodies.removeAt(j) for every victim: stable compaction of the home-ordered state; user
// indices shrink by the number of removed bodies before them (list order is preserved).
__global__ void k_merge_compact(const int* __restrict__ dead, const int* __restrict__ new_home, const int* __restrict__ new_user,
                                int n, const double* __restrict__ x, const double* __restrict__ y,
                                const double* __restrict__ vx, const double* __restrict__ vy, const double* __restrict__ m,
                                const int* __restrict__ perm, double* __restrict__ x2, double* __restrict__ y2,
                                double* __restrict__ vx2, double* __restrict__ vy2, double* __restrict__ m2,
                                int* __restrict__ perm2) {
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= n || dead[h]) return;
    const int d = new_home[h];
    x2[d] = x[h]; y2[d] = y[h]; vx2[d] = vx[h]; vy2[d] = vy[h]; m2[d] = m[h];
    perm2[d] = new_user[perm[h]];
}
__global__ void k_merge_compact_origin(const int* __restrict__ dead, const int* __restrict__ inv, const int* __restrict__ new_user,
                                       int n, const int* __restrict__ origin, int* __restrict__ origin2) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u < n && !dead[inv[u]]) origin2[new_user[u]] = origin[u];
}

// register-only FFMA throughput probe: 8 independent chains per thread
__global__ void __launch_bounds__(256) k_fp32_peak(float* out, int iters, float a, float b) {
    float v0 = threadIdx.x, v1 = v0 + 1.f, v2 = v0 + 2.f, v3 = v0 + 3.f, v4 = v0 + 4.f, v5 = v0 + 5.f, v6 = v0 + 6.f, v7 = v0 + 7.f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            v0 = fmaf(v0, a, b); v1 = fmaf(v1, a, b); v2 = fmaf(v2, a, b); v3 = fmaf(v3, a, b);
            v4 = fmaf(v4, a, b); v5 = fmaf(v5, a, b); v6 = fmaf(v6, a, b); v7 = fmaf(v7, a, b);
        }
    }
    const float s = v0 + v1 + v2 + v3 + v4 + v5 + v6 + v7;
    if (s == 12345.678f) out[0] = s;   // never true; keeps the chains alive
}

}  // namespace

#endif  // BH_KERNELS_CUH
