// bh_engine.cu — B200 (sm_100a) Barnes–Hut physics-step engine behind include/bh_engine.h.
//
// One PhysicsEngine.step() of the reference (BarnesHutAlg.kt:405-439) becomes, per force
// evaluation:  k_keygen -> onesweep radix sort -> k_count_scan -> k_emit -> k_climb -> k_walk,
// then the f64 kick/drift kernels.  The tree is the reference's own quadtree (same cells,
// same f64 centres of mass, same per-body accept/open decisions) stored as a flattened
// DFS-preorder SoA with skip links; see bh_core.h and DESIGN.md.
//
// No CPU fallback: every compute entry point runs CUDA kernels or returns an error.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <new>
#include <string>
#include <vector>

#include "../../include/bh_engine.h"
#include "bh_core.h"
#include "bh_export.h"
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
    unsigned int pad;
};
struct DevTotals {            // zeroed by bh_reset_counters only
    unsigned long long interactions, opened, retests, evaluations;
};

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_IPT = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_IPT;

// ---------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------

// Morton keys by literal descent (BH.kt:153-155, :73-80) + root contains() (BH.kt:126).
// HBM-bound: 16 B read + 8 B written per body.
__global__ void __launch_bounds__(256) k_keygen(const double* __restrict__ x, const double* __restrict__ y, int n,
                                                BhRoot root, uint64_t sentinel, uint64_t* __restrict__ keys,
                                                DevScalars* __restrict__ sc) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    bool in = false;
    if (b < n) {
        const double px = x[b], py = y[b];
        in = bh_root_contains(root, px, py);
        keys[b] = in ? bh_morton_key(root, px, py) : sentinel;
    }
    const unsigned ball = __ballot_sync(0xffffffffu, in);
    if (sc && (threadIdx.x & 31) == 0 && ball) atomicAdd(&sc->n_in, __popc(ball));
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

    // decoupled look-back, one warp, 32 predecessors per round
    if (w == 0) {
        int excl = 0;
        if (tile == 0) {
            if (lane == 0) bhsort::st_volatile_u32(status, bhsort::FLAG_PREFIX | (uint32_t)total);
        } else {
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
        }
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
    if (lane == 0) {
        if (jit) atomicAdd(&sc->n_jitter, jit);
        atomicMax(&sc->max_depth, maxd);
    }
}

// cell skeletons (skip / parent / count / level) — bh_emit_body per in-tree body
__global__ void __launch_bounds__(256) k_emit(BhTreeView t, int levels) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < t.n_in) bh_emit_body(t, levels, i);
}

// computeMass (BH.kt:173-202) bottom-up — bh_climb_body per in-tree body
__global__ void __launch_bounds__(256) k_climb(BhTreeView t, BhRoot root, const double* __restrict__ x,
                                               const double* __restrict__ y, const double* __restrict__ m) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= t.n_in) return;
    const int b = t.order[i];
    bh_climb_body(t, root, i, x[b], y[b], m[b]);
}

// accumulateForce (BH.kt:215-239) + ax = fx/m (BH.kt:390-391): one thread per target body,
// Morton-adjacent bodies in a warp, stackless over the preorder cells.
template <bool ZERO_MASS>
__global__ void __launch_bounds__(128)
k_walk(BhTreeView t, BhWalkParams w, int first_target, int n_targets, const double* __restrict__ x,
       const double* __restrict__ y, const double* __restrict__ m, double G, double* __restrict__ ax,
       double* __restrict__ ay, int* __restrict__ cntI, int* __restrict__ cntO, DevScalars* __restrict__ sc,
       DevTotals* __restrict__ tot) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    int ni = 0, no = 0, nr = 0;
    // every lane enters the walk (it contains full-warp shuffles); surplus lanes see an empty tree
    const bool active = k < n_targets;
    const int si = first_target + (active ? k : 0);
    const int b = t.order[si];
    const int self = (si < t.n_in) ? (t.S[si + 1] + si) : -1;
    BhTreeView tv = t;
    if (!active) tv.M = 0;
    const BhWalkResult r = bh_walk_body<ZERO_MASS>(tv, w, x[b], y[b], self);
    if (active) {
        const double mb = m[b];
        // BH.kt:390-391 divides the force by b.m: a zero-mass body gets 0/0 = NaN
        ax[b] = (mb == 0.0) ? nan("") : G * r.ax;
        ay[b] = (mb == 0.0) ? nan("") : G * r.ay;
        ni = r.interactions; no = r.opened; nr = r.retests;
        if (cntI) { cntI[b] = ni; cntO[b] = no; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ni += __shfl_xor_sync(0xffffffffu, ni, o);
        no += __shfl_xor_sync(0xffffffffu, no, o);
        nr += __shfl_xor_sync(0xffffffffu, nr, o);
    }
    if ((threadIdx.x & 31) == 0) {
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

__global__ void k_positions_f32(const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ m,
                                int n, float2* __restrict__ xy, float* __restrict__ mf) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { xy[i] = make_float2((float)x[i], (float)y[i]); mf[i] = (float)m[i]; }
}

// depth of each body's leaf, scattered to body order (bh_get_morton)
__global__ void k_leaf_depth(BhTreeView t, int n, int* __restrict__ depth) {
    const int si = blockIdx.x * blockDim.x + threadIdx.x;
    if (si >= n) return;
    depth[t.order[si]] = (si < t.n_in) ? t.sk[t.S[si + 1] + si].level : -1;
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

thread_local std::string g_create_err;

template <class T>
cudaError_t dev_alloc(T** p, size_t count) {
    *p = nullptr;
    return cudaMalloc((void**)p, std::max<size_t>(count, 1) * sizeof(T));
}
template <class T>
void dev_free(T*& p) { if (p) cudaFree(p); p = nullptr; }

}  // namespace

// ---------------------------------------------------------------------------------------
// engine
// ---------------------------------------------------------------------------------------
struct bh_engine {
    bh_config cfg{};
    bh_params par{};
    int device = 0, num_sms = 148;
    cudaStream_t st = nullptr;
    cudaEvent_t ev[12]{};   // [0..3] evaluation A, [4..7] evaluation B, [8..9] whole step, [10..11] whole call
    std::string err;

    // body state, f64 SoA, index = position in the reference's `bodies` list
    int64_t n = 0, cap = 0;
    double *x = nullptr, *y = nullptr, *vx = nullptr, *vy = nullptr, *m = nullptr, *ax = nullptr, *ay = nullptr;
    int *cntI = nullptr, *cntO = nullptr;
    std::vector<int32_t> origin;
    bool any_zero_mass = false;   // some body has m == 0 (zero-mass cells are pruned, BH.kt:216)

    // sort buffers + scratch (scalars | sort scratch | scan status) zeroed per build
    uint64_t *keys_a = nullptr, *keys_b = nullptr;
    uint32_t *vals_a = nullptr, *vals_b = nullptr;
    uint32_t* scratch = nullptr;
    size_t scratch_words = 0;
    DevScalars* sc_host = nullptr;   // pinned
    DevTotals* tot = nullptr;
    DevTotals* tot_host = nullptr;   // pinned
    double* red = nullptr;           // 4 doubles for k_energy

    // tree
    int* S = nullptr;
    int64_t cell_cap = 0;
    BhCell* cell = nullptr;      // hot 32 B records
    BhCellD* cd = nullptr;       // exact f64 records
    BhCellS* sk = nullptr;       // skeletons
    int* arrived = nullptr;

    bool tree_valid = false;
    BhRoot root{};
    int n_in = 0, n_internal = 0, M = 0;
    const uint64_t* keys_sorted = nullptr;
    const int* order = nullptr;

    bh_counters ctr{};

    int fail(int code, const char* what) { err = what; return code; }
    int cuda_fail(cudaError_t e, const char* what) {
        err = std::string(what) + ": " + cudaGetErrorString(e);
        return e == cudaErrorMemoryAllocation ? BH_E_OOM : BH_E_CUDA;
    }

    static constexpr size_t SC_WORDS = (sizeof(DevScalars) + 3) / 4;
    DevScalars* sc() const { return reinterpret_cast<DevScalars*>(scratch); }
    uint32_t* sort_scratch() const { return scratch + SC_WORDS; }
    size_t scan_tiles(int64_t nn) const { return (size_t)((nn + SCAN_TILE - 1) / SCAN_TILE) + 1; }

    void free_bodies() {
        dev_free(x); dev_free(y); dev_free(vx); dev_free(vy); dev_free(m); dev_free(ax); dev_free(ay);
        dev_free(cntI); dev_free(cntO);
        dev_free(keys_a); dev_free(keys_b); dev_free(vals_a); dev_free(vals_b); dev_free(scratch); dev_free(S);
        cap = 0;
    }
    void free_cells() {
        dev_free(cell); dev_free(cd); dev_free(sk); dev_free(arrived);
        cell_cap = 0;
    }

#define BH_TRY(expr)                                                        \
    do {                                                                    \
        cudaError_t _e = (expr);                                            \
        if (_e != cudaSuccess) return cuda_fail(_e, #expr);                 \
    } while (0)

    int ensure_bodies(int64_t nn) {
        if (nn <= cap) return BH_OK;
        if (nn >= (int64_t)1 << 30) return fail(BH_E_ARG, "more than 2^30 bodies are not supported");
        const int64_t c = std::max<int64_t>(nn, std::max<int64_t>(1024, cap + cap / 4));
        free_bodies();
        BH_TRY(dev_alloc(&x, c)); BH_TRY(dev_alloc(&y, c)); BH_TRY(dev_alloc(&vx, c)); BH_TRY(dev_alloc(&vy, c));
        BH_TRY(dev_alloc(&m, c)); BH_TRY(dev_alloc(&ax, c)); BH_TRY(dev_alloc(&ay, c));
        if (cfg.flags & BH_FLAG_BODY_COUNTS) { BH_TRY(dev_alloc(&cntI, c)); BH_TRY(dev_alloc(&cntO, c)); }
        BH_TRY(dev_alloc(&keys_a, c)); BH_TRY(dev_alloc(&keys_b, c));
        BH_TRY(dev_alloc(&vals_a, c)); BH_TRY(dev_alloc(&vals_b, c));
        BH_TRY(dev_alloc(&S, c + 1));
        scratch_words = SC_WORDS + bhsort::sort_scratch_words(c, bhsort::MAX_PASSES) + scan_tiles(c);
        BH_TRY(dev_alloc(&scratch, scratch_words));
        cap = c;
        return BH_OK;
    }
    int ensure_cells(int64_t mm) {
        if (mm <= cell_cap) return BH_OK;
        if (mm >= (int64_t)1 << 31) return fail(BH_E_ARG, "tree has more than 2^31 cells");
        const int64_t c = std::max<int64_t>(mm + mm / 8, 2048);
        free_cells();
        BH_TRY(dev_alloc(&cell, c)); BH_TRY(dev_alloc(&cd, c)); BH_TRY(dev_alloc(&sk, c));
        BH_TRY(dev_alloc(&arrived, c));
        cell_cap = c;
        return BH_OK;
    }

    BhTreeView view() const {
        BhTreeView t{};
        t.keys = keys_sorted; t.order = order; t.S = S;
        t.cell = cell; t.cd = cd; t.sk = sk; t.arrived = arrived;
        t.n_in = n_in; t.M = M;
        return t;
    }

    // buildTree(), BH.kt:359-366
    int build(int slot = 0) {
        tree_valid = false;
        root = BhRoot{par.root_cx, par.root_cy, par.root_half, bh_key_levels(par.root_half)};
        const int nn = (int)n;
        BH_TRY(cudaEventRecord(ev[slot + 0], st));
        // zero: scalars | sort scratch (sized for this n) | scan status
        const int key_bits = 2 * root.levels + 1;   // +1: the not-in-tree sentinel 1<<2L sorts last
        const int passes = (key_bits + bhsort::RADIX_BITS - 1) / bhsort::RADIX_BITS;
        const size_t sortw = bhsort::sort_scratch_words(nn, passes);
        uint32_t* scan_status = sort_scratch() + sortw;
        BH_TRY(cudaMemsetAsync(scratch, 0, (SC_WORDS + sortw + scan_tiles(nn)) * sizeof(uint32_t), st));
        n_in = 0; n_internal = 0; M = 0;
        if (nn > 0) {
            const uint64_t sentinel = 1ull << (2 * root.levels);
            k_keygen<<<(nn + 255) / 256, 256, 0, st>>>(x, y, nn, root, sentinel, keys_a, sc());
            const int where = sort_pairs(nn, key_bits, sortw);
            keys_sorted = where ? keys_b : keys_a;
            order = reinterpret_cast<const int*>(where ? vals_b : vals_a);
            k_count_scan<<<(nn + SCAN_TILE - 1) / SCAN_TILE, SCAN_THREADS, 0, st>>>(keys_sorted, root.levels, sc(), S, scan_status);
            ctr.kernel_launches += 4 + passes;   // keygen, histogram, histogram_scan, passes, count_scan
        }
        BH_TRY(cudaMemcpyAsync(sc_host, sc(), sizeof(DevScalars), cudaMemcpyDeviceToHost, st));
        BH_TRY(cudaStreamSynchronize(st));
        BH_TRY(cudaGetLastError());
        n_in = sc_host->n_in; n_internal = sc_host->n_internal; M = n_in + n_internal;
        ctr.n_in_tree = n_in; ctr.n_out_of_box = n - n_in; ctr.n_internal = n_internal; ctr.n_cells = M;
        ctr.n_jitter_bodies = sc_host->n_jitter; ctr.max_depth = sc_host->max_depth; ctr.key_levels = root.levels;
        if (int rc = ensure_cells(M)) return rc;
        if (n_in > 0) {
            BH_TRY(cudaMemsetAsync(arrived, 0, (size_t)M * sizeof(int), st));
            const BhTreeView t = view();
            k_emit<<<(n_in + 255) / 256, 256, 0, st>>>(t, root.levels);
            k_climb<<<(n_in + 255) / 256, 256, 0, st>>>(t, root, x, y, m);
            ctr.kernel_launches += 2;
        }
        BH_TRY(cudaEventRecord(ev[slot + 1], st));
        BH_TRY(cudaGetLastError());
        tree_valid = true;
        return BH_OK;
    }

    int sort_pairs(int nn, int key_bits, size_t sortw) {
        (void)sortw;
        // the sort zeroes nothing itself here: build() already cleared the scratch region
        const int passes = (key_bits + bhsort::RADIX_BITS - 1) / bhsort::RADIX_BITS;
        const int tiles = bhsort::sort_tiles(nn);
        uint32_t* hist = sort_scratch();
        uint32_t* tickets = hist + bhsort::MAX_PASSES * bhsort::RADIX;
        uint32_t* lookback = tickets + bhsort::MAX_PASSES;
        int hb = std::min((nn + 2047) / 2048, num_sms * 8);
        if (hb < 1) hb = 1;
        bhsort::k_histogram<<<hb, 256, 0, st>>>(keys_a, nullptr, nn, passes, hist);
        bhsort::k_histogram_scan<<<passes, bhsort::RADIX, 0, st>>>(hist);
        int cur = 0;
        for (int p = 0; p < passes; ++p) {
            const uint64_t* kin = cur ? keys_b : keys_a;
            const uint32_t* vin = (p == 0) ? nullptr : (cur ? vals_b : vals_a);
            uint64_t* kout = cur ? keys_a : keys_b;
            uint32_t* vout = cur ? vals_a : vals_b;
            bhsort::k_onesweep_pass<<<tiles, bhsort::SORT_THREADS, 0, st>>>(
                kin, vin, kout, vout, nn, p * bhsort::RADIX_BITS, hist + p * bhsort::RADIX, tickets + p,
                lookback + (size_t)p * tiles * bhsort::RADIX);
            cur ^= 1;
        }
        return cur;
    }

    // computeAccelerations(root), BH.kt:374-395, for sorted targets [first, first+count)
    int walk(int first, int count, int slot = 0) {
        BH_TRY(cudaEventRecord(ev[slot + 2], st));
        if (count > 0) {
            const BhWalkParams w = bh_walk_params(par.theta, par.soft2, par.root_half);
            if (any_zero_mass)
                k_walk<true><<<(count + 127) / 128, 128, 0, st>>>(view(), w, first, count, x, y, m, par.G, ax, ay, cntI, cntO, sc(), tot);
            else
                k_walk<false><<<(count + 127) / 128, 128, 0, st>>>(view(), w, first, count, x, y, m, par.G, ax, ay, cntI, cntO, sc(), tot);
            ctr.kernel_launches += 1;
        }
        BH_TRY(cudaEventRecord(ev[slot + 3], st));
        BH_TRY(cudaGetLastError());
        return BH_OK;
    }

    int evaluate(int slot = 0) {
        if (int rc = build(slot)) return rc;
        if (int rc = walk(0, (int)n, slot)) return rc;
        ctr.total_evaluations++;
        return BH_OK;
    }

    // after a stream sync: fold event timings and device counters into ctr
    int finish() {
        BH_TRY(cudaMemcpyAsync(sc_host, sc(), sizeof(DevScalars), cudaMemcpyDeviceToHost, st));
        BH_TRY(cudaMemcpyAsync(tot_host, tot, sizeof(DevTotals), cudaMemcpyDeviceToHost, st));
        BH_TRY(cudaStreamSynchronize(st));
        BH_TRY(cudaGetLastError());
        ctr.interactions = (int64_t)sc_host->interactions;
        ctr.opened = (int64_t)sc_host->opened;
        ctr.exact_retests = (int64_t)sc_host->retests;
        ctr.total_interactions = (int64_t)tot_host->interactions;
        ctr.total_opened = (int64_t)tot_host->opened;
        return BH_OK;
    }
    // events of evaluate(slot); call after a sync.  Returns build+walk ms.
    float add_phase_times(int slot) {
        float a = 0.f, b = 0.f;
        if (cudaEventElapsedTime(&a, ev[slot + 0], ev[slot + 1]) == cudaSuccess) ctr.ms_build += a; else a = 0.f;
        if (cudaEventElapsedTime(&b, ev[slot + 2], ev[slot + 3]) == cudaSuccess) ctr.ms_walk += b; else b = 0.f;
        return a + b;
    }

    int kick(double dtHalf, double dt, int drift) {
        if (n > 0) { k_kick_drift<<<((int)n + 255) / 256, 256, 0, st>>>(0, (int)n, x, y, vx, vy, ax, ay, dtHalf, dt, drift); ctr.kernel_launches += 1; }
        BH_TRY(cudaGetLastError());
        return BH_OK;
    }

    // PhysicsEngine.step(), BH.kt:405-439
    int step_once() {
        const double dt = par.dt, dtHalf = par.dt * 0.5;   // BH.kt:412
        if (int rc = evaluate(0)) return rc;               // a(t)
        if (int rc = kick(dtHalf, dt, 1)) return rc;       // kick + drift
        if (int rc = evaluate(4)) return rc;               // a(t+dt)
        if (int rc = kick(dtHalf, dt, 0)) return rc;       // kick
        ctr.total_steps++;
        return merge_rule();
    }

    bool merge_enabled() const { return par.merge_min_dist > 0.0 && n > 1; }
    std::vector<int32_t> heavies;        // indices with m > merge_max_mass, ascending (host cache)
    bool heavies_valid = false;
    int merge_rule();
};

#include "bh_merge.cuh"

// ---------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------
extern "C" {

int bh_abi_version(void) { return BH_ABI_VERSION; }
const char* bh_backend_name(void) { return "b200-cuda"; }

int bh_default_params(int32_t w, int32_t h, bh_params* p) {
    if (!p) return BH_E_ARG;
    p->G = 80.0; p->dt = 0.005; p->theta = 0.30; p->soft2 = 1.0 * 1.0;   // Config.kt:11,14,23,17,20
    p->root_cx = w / 2.0; p->root_cy = h / 2.0;                          // BH.kt:361
    p->root_half = std::max(w, h) / 2.0 + 2.0;                           // BH.kt:360
    p->merge_max_mass = 4000.0; p->merge_min_dist = 8.0;                 // BH.kt:315,321
    return BH_OK;
}

int bh_create(const bh_config* cfg, bh_engine** out) {
    if (!out) { g_create_err = "bh_create: out is NULL"; return BH_E_ARG; }
    *out = nullptr;
    bh_engine* e = new (std::nothrow) bh_engine();
    if (!e) { g_create_err = "bh_create: out of memory"; return BH_E_OOM; }
    if (cfg) memcpy(&e->cfg, cfg, std::min<size_t>(sizeof(bh_config), cfg->struct_size > 0 ? (size_t)cfg->struct_size : sizeof(bh_config)));
    e->device = e->cfg.device;
    cudaError_t ce = cudaSetDevice(e->device);
    if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&e->st, cudaStreamNonBlocking);
    for (int k = 0; k < 12 && ce == cudaSuccess; ++k) ce = cudaEventCreate(&e->ev[k]);
    if (ce == cudaSuccess) ce = cudaMallocHost((void**)&e->sc_host, sizeof(DevScalars));
    if (ce == cudaSuccess) ce = cudaMallocHost((void**)&e->tot_host, sizeof(DevTotals));
    if (ce == cudaSuccess) ce = dev_alloc(&e->tot, 1);
    if (ce == cudaSuccess) ce = cudaMemset(e->tot, 0, sizeof(DevTotals));
    if (ce == cudaSuccess) ce = dev_alloc(&e->red, 4);
    if (ce == cudaSuccess) ce = cudaDeviceGetAttribute(&e->num_sms, cudaDevAttrMultiProcessorCount, e->device);
    if (ce != cudaSuccess) {
        g_create_err = std::string("bh_create: CUDA device ") + std::to_string(e->device) + " unavailable: " +
                       cudaGetErrorString(ce) + " (this library has no CPU fallback)";
        bh_destroy(e);
        return BH_E_CUDA;
    }
    bh_default_params(2400, 800, &e->par);
    if (e->cfg.capacity_hint > 0) {
        const int rc = e->ensure_bodies(e->cfg.capacity_hint);
        if (rc != BH_OK) { g_create_err = e->err; bh_destroy(e); return rc; }
    }
    *out = e;
    return BH_OK;
}

void bh_destroy(bh_engine* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    if (e->st) cudaStreamSynchronize(e->st);
    e->free_bodies();
    e->free_cells();
    dev_free(e->tot); dev_free(e->red);
    if (e->sc_host) cudaFreeHost(e->sc_host);
    if (e->tot_host) cudaFreeHost(e->tot_host);
    for (auto& ev : e->ev) if (ev) cudaEventDestroy(ev);
    if (e->st) cudaStreamDestroy(e->st);
    delete e;
}

const char* bh_last_error(const bh_engine* e) { return e ? e->err.c_str() : g_create_err.c_str(); }

int bh_set_params(bh_engine* e, const bh_params* p) {
    if (!e || !p) return BH_E_ARG;
    if (!(p->root_half > 0.0)) return e->fail(BH_E_ARG, "bh_set_params: root_half must be > 0");
    e->par = *p;
    return BH_OK;
}
int bh_get_params(const bh_engine* e, bh_params* p) {
    if (!e || !p) return BH_E_ARG;
    *p = e->par;
    return BH_OK;
}

#define E_TRY(expr)                                                         \
    do {                                                                    \
        cudaError_t _e = (expr);                                            \
        if (_e != cudaSuccess) return e->cuda_fail(_e, #expr);              \
    } while (0)

int bh_set_bodies(bh_engine* e, int64_t n, const double* x, const double* y, const double* vx, const double* vy,
                  const double* m) {
    if (!e || n < 0 || (n > 0 && (!x || !y || !vx || !vy || !m))) return e ? e->fail(BH_E_ARG, "bh_set_bodies: bad arguments") : BH_E_ARG;
    E_TRY(cudaSetDevice(e->device));
    if (int rc = e->ensure_bodies(n)) return rc;
    const size_t bytes = (size_t)n * sizeof(double);
    if (n > 0) {
        E_TRY(cudaMemcpyAsync(e->x, x, bytes, cudaMemcpyHostToDevice, e->st));
        E_TRY(cudaMemcpyAsync(e->y, y, bytes, cudaMemcpyHostToDevice, e->st));
        E_TRY(cudaMemcpyAsync(e->vx, vx, bytes, cudaMemcpyHostToDevice, e->st));
        E_TRY(cudaMemcpyAsync(e->vy, vy, bytes, cudaMemcpyHostToDevice, e->st));
        E_TRY(cudaMemcpyAsync(e->m, m, bytes, cudaMemcpyHostToDevice, e->st));
        E_TRY(cudaMemsetAsync(e->ax, 0, bytes, e->st));
        E_TRY(cudaMemsetAsync(e->ay, 0, bytes, e->st));
    }
    E_TRY(cudaStreamSynchronize(e->st));
    e->n = n;
    e->tree_valid = false;
    e->heavies_valid = false;
    e->any_zero_mass = false;
    for (int64_t i = 0; i < n; ++i) if (m[i] == 0.0) { e->any_zero_mass = true; break; }
    try {
        e->origin.resize((size_t)n);
        for (int64_t i = 0; i < n; ++i) e->origin[(size_t)i] = (int32_t)i;
    } catch (const std::bad_alloc&) { return e->fail(BH_E_OOM, "bh_set_bodies: host out of memory"); }
    return BH_OK;
}

int64_t bh_num_bodies(const bh_engine* e) { return e ? e->n : 0; }

int bh_get_bodies(bh_engine* e, int64_t cap, double* x, double* y, double* vx, double* vy, double* m, int64_t* n_out) {
    if (!e) return BH_E_ARG;
    if (n_out) *n_out = e->n;
    if (cap < e->n) return e->fail(BH_E_ARG, "bh_get_bodies: capacity too small");
    E_TRY(cudaSetDevice(e->device));
    const size_t bytes = (size_t)e->n * sizeof(double);
    if (e->n > 0) {
        if (x) E_TRY(cudaMemcpyAsync(x, e->x, bytes, cudaMemcpyDeviceToHost, e->st));
        if (y) E_TRY(cudaMemcpyAsync(y, e->y, bytes, cudaMemcpyDeviceToHost, e->st));
        if (vx) E_TRY(cudaMemcpyAsync(vx, e->vx, bytes, cudaMemcpyDeviceToHost, e->st));
        if (vy) E_TRY(cudaMemcpyAsync(vy, e->vy, bytes, cudaMemcpyDeviceToHost, e->st));
        if (m) E_TRY(cudaMemcpyAsync(m, e->m, bytes, cudaMemcpyDeviceToHost, e->st));
    }
    E_TRY(cudaStreamSynchronize(e->st));
    return BH_OK;
}

int bh_get_origin(bh_engine* e, int64_t cap, int32_t* origin, int64_t* n_out) {
    if (!e) return BH_E_ARG;
    if (n_out) *n_out = e->n;
    if (cap < e->n) return e->fail(BH_E_ARG, "bh_get_origin: capacity too small");
    if (origin && e->n > 0) memcpy(origin, e->origin.data(), (size_t)e->n * sizeof(int32_t));
    return BH_OK;
}

int bh_get_positions_f32(bh_engine* e, int64_t cap, float* xy, float* m, int64_t* n_out) {
    if (!e) return BH_E_ARG;
    if (n_out) *n_out = e->n;
    if (cap < e->n) return e->fail(BH_E_ARG, "bh_get_positions_f32: capacity too small");
    if (e->n == 0) return BH_OK;
    E_TRY(cudaSetDevice(e->device));
    // stage through the (idle) sort buffers: keys_b as float2[n], vals_b as float[n]
    float2* dxy = reinterpret_cast<float2*>(e->keys_b);
    float* dm = reinterpret_cast<float*>(e->vals_b);
    const bool keep_tree = e->tree_valid && e->keys_sorted != e->keys_b;
    if (e->tree_valid && !keep_tree) e->tree_valid = false;   // the staging overwrote the sorted keys
    k_positions_f32<<<((int)e->n + 255) / 256, 256, 0, e->st>>>(e->x, e->y, e->m, (int)e->n, dxy, dm);
    if (xy) E_TRY(cudaMemcpyAsync(xy, dxy, (size_t)e->n * sizeof(float2), cudaMemcpyDeviceToHost, e->st));
    if (m) E_TRY(cudaMemcpyAsync(m, dm, (size_t)e->n * sizeof(float), cudaMemcpyDeviceToHost, e->st));
    E_TRY(cudaStreamSynchronize(e->st));
    return BH_OK;
}

int bh_step(bh_engine* e, int32_t nsteps) {
    if (!e || nsteps < 0) return e ? e->fail(BH_E_ARG, "bh_step: bad arguments") : BH_E_ARG;
    E_TRY(cudaSetDevice(e->device));
    E_TRY(cudaEventRecord(e->ev[10], e->st));
    for (int s = 0; s < nsteps; ++s) {
        E_TRY(cudaEventRecord(e->ev[8], e->st));
        if (int rc = e->step_once()) return rc;
        E_TRY(cudaEventRecord(e->ev[9], e->st));
        if (int rc = e->finish()) return rc;
        const float phases = e->add_phase_times(0) + e->add_phase_times(4);
        float total = 0.f;
        // kick/drift (+ merge) = whole step minus the build and walk phases
        if (cudaEventElapsedTime(&total, e->ev[8], e->ev[9]) == cudaSuccess && total > phases) e->ctr.ms_integrate += total - phases;
    }
    E_TRY(cudaEventRecord(e->ev[11], e->st));
    if (int rc = e->finish()) return rc;
    float call_ms = 0.f;
    if (cudaEventElapsedTime(&call_ms, e->ev[10], e->ev[11]) == cudaSuccess) e->ctr.ms_step_call = call_ms;
    return BH_OK;
}

int bh_build_tree(bh_engine* e) {
    if (!e) return BH_E_ARG;
    E_TRY(cudaSetDevice(e->device));
    if (int rc = e->build()) return rc;
    E_TRY(cudaStreamSynchronize(e->st));
    return BH_OK;
}

int bh_compute_accelerations(bh_engine* e, double* ax, double* ay) {
    if (!e) return BH_E_ARG;
    E_TRY(cudaSetDevice(e->device));
    if (int rc = e->evaluate(0)) return rc;
    if (int rc = e->finish()) return rc;
    e->add_phase_times(0);
    const size_t bytes = (size_t)e->n * sizeof(double);
    if (e->n > 0) {
        if (ax) E_TRY(cudaMemcpy(ax, e->ax, bytes, cudaMemcpyDeviceToHost));
        if (ay) E_TRY(cudaMemcpy(ay, e->ay, bytes, cudaMemcpyDeviceToHost));
    }
    return BH_OK;
}

int bh_direct_sum(bh_engine* e, double* ax, double* ay) {
    if (!e) return BH_E_ARG;
    E_TRY(cudaSetDevice(e->device));
    if (e->n == 0) return BH_OK;
    // results go to the sort buffers (not e->ax/ay, which belong to the integrator)
    double* dax = reinterpret_cast<double*>(e->keys_a);
    double* day = reinterpret_cast<double*>(e->keys_b);
    e->tree_valid = false;
    k_direct<<<((int)e->n + DS_TILE - 1) / DS_TILE, DS_TILE, 0, e->st>>>(e->x, e->y, e->m, (int)e->n, (float)e->par.soft2,
                                                                          e->par.G, dax, day);
    E_TRY(cudaGetLastError());
    const size_t bytes = (size_t)e->n * sizeof(double);
    if (ax) E_TRY(cudaMemcpyAsync(ax, dax, bytes, cudaMemcpyDeviceToHost, e->st));
    if (ay) E_TRY(cudaMemcpyAsync(ay, day, bytes, cudaMemcpyDeviceToHost, e->st));
    E_TRY(cudaStreamSynchronize(e->st));
    return BH_OK;
}

int bh_energy(bh_engine* e, double* ke, double* pe, double* px, double* py) {
    if (!e) return BH_E_ARG;
    E_TRY(cudaSetDevice(e->device));
    double h[4] = {0, 0, 0, 0};
    if (e->n > 0) {
        E_TRY(cudaMemsetAsync(e->red, 0, 4 * sizeof(double), e->st));
        k_energy<<<((int)e->n + DS_TILE - 1) / DS_TILE, DS_TILE, 0, e->st>>>(e->x, e->y, e->vx, e->vy, e->m, (int)e->n,
                                                                              e->par.soft2, e->red);
        E_TRY(cudaGetLastError());
        E_TRY(cudaMemcpyAsync(h, e->red, sizeof(h), cudaMemcpyDeviceToHost, e->st));
        E_TRY(cudaStreamSynchronize(e->st));
    }
    if (ke) *ke = h[0];
    if (pe) *pe = -0.5 * e->par.G * h[1];
    if (px) *px = h[2];
    if (py) *py = h[3];
    return BH_OK;
}

int bh_get_morton(bh_engine* e, uint64_t* key, int32_t* depth, int32_t* order) {
    if (!e) return BH_E_ARG;
    E_TRY(cudaSetDevice(e->device));
    if (!e->tree_valid) { if (int rc = bh_build_tree(e)) return rc; }
    const int n = (int)e->n;
    if (n == 0) return BH_OK;
    uint64_t* dkey = nullptr;
    int* ddepth = nullptr;
    E_TRY(dev_alloc(&dkey, n));
    cudaError_t ce = dev_alloc(&ddepth, n);
    if (ce == cudaSuccess) {
        // sentinel here is the ABI's UINT64_MAX, not the sortable 1<<2L
        k_keygen<<<(n + 255) / 256, 256, 0, e->st>>>(e->x, e->y, n, e->root, BH_KEY_NOT_IN_TREE, dkey, nullptr);
        k_leaf_depth<<<(n + 255) / 256, 256, 0, e->st>>>(e->view(), n, ddepth);
        ce = cudaGetLastError();
        if (ce == cudaSuccess && key) ce = cudaMemcpyAsync(key, dkey, (size_t)n * 8, cudaMemcpyDeviceToHost, e->st);
        if (ce == cudaSuccess && depth) ce = cudaMemcpyAsync(depth, ddepth, (size_t)n * 4, cudaMemcpyDeviceToHost, e->st);
        if (ce == cudaSuccess && order) ce = cudaMemcpyAsync(order, e->order, (size_t)n * 4, cudaMemcpyDeviceToHost, e->st);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->st);
    }
    cudaFree(dkey);
    cudaFree(ddepth);
    if (ce != cudaSuccess) return e->cuda_fail(ce, "bh_get_morton");
    return BH_OK;
}

int bh_get_tree(bh_engine* e, int64_t cap, int64_t* n_cells, double* cx, double* cy, double* h, double* mass,
                double* comx, double* comy, int32_t* body) {
    if (!e) return BH_E_ARG;
    E_TRY(cudaSetDevice(e->device));
    if (!e->tree_valid) { if (int rc = bh_build_tree(e)) return rc; }   // lastTree ?: buildTree(), BH.kt:329-332
    try {
        const size_t M = (size_t)e->M, ni = (size_t)e->n_in;
        std::vector<uint64_t> keys(ni);
        std::vector<int> order(ni), S(ni + 1);
        std::vector<BhCellS> sk(M);
        std::vector<BhCellD> cd(M);
        if (ni) {
            E_TRY(cudaMemcpy(keys.data(), e->keys_sorted, ni * 8, cudaMemcpyDeviceToHost));
            E_TRY(cudaMemcpy(order.data(), e->order, ni * 4, cudaMemcpyDeviceToHost));
            E_TRY(cudaMemcpy(S.data(), e->S, (ni + 1) * 4, cudaMemcpyDeviceToHost));
            E_TRY(cudaMemcpy(sk.data(), e->sk, M * sizeof(BhCellS), cudaMemcpyDeviceToHost));
            E_TRY(cudaMemcpy(cd.data(), e->cd, M * sizeof(BhCellD), cudaMemcpyDeviceToHost));
        }
        BhHostTree t{e->root, e->n_in, e->M, keys.data(), order.data(), S.data(), sk.data(), cd.data()};
        BhCellsOut out;
        out.cap = cap; out.cx = cx; out.cy = cy; out.h = h; out.mass = mass; out.comx = comx; out.comy = comy; out.body = body;
        bh_export_cells(t, out);
        if (n_cells) *n_cells = out.count;
        if (cap != 0 && cap < out.count) return e->fail(BH_E_ARG, "bh_get_tree: capacity too small");
    } catch (const std::bad_alloc&) { return e->fail(BH_E_OOM, "bh_get_tree: host out of memory"); }
    return BH_OK;
}

int bh_get_counters(bh_engine* e, bh_counters* out) {
    if (!e || !out) return BH_E_ARG;
    e->ctr.n_bodies = e->n;
    *out = e->ctr;
    return BH_OK;
}

int bh_reset_counters(bh_engine* e) {
    if (!e) return BH_E_ARG;
    E_TRY(cudaSetDevice(e->device));
    E_TRY(cudaMemsetAsync(e->tot, 0, sizeof(DevTotals), e->st));
    E_TRY(cudaStreamSynchronize(e->st));
    const bh_counters keep = e->ctr;
    e->ctr = bh_counters{};
    e->ctr.n_in_tree = keep.n_in_tree; e->ctr.n_out_of_box = keep.n_out_of_box; e->ctr.n_cells = keep.n_cells;
    e->ctr.n_internal = keep.n_internal; e->ctr.key_levels = keep.key_levels; e->ctr.max_depth = keep.max_depth;
    e->ctr.n_jitter_bodies = keep.n_jitter_bodies;
    e->ctr.ms_step_call = keep.ms_step_call;
    return BH_OK;
}

int bh_get_body_counts(bh_engine* e, int32_t* interactions, int32_t* opened) {
    if (!e) return BH_E_ARG;
    if (!(e->cfg.flags & BH_FLAG_BODY_COUNTS)) return e->fail(BH_E_STATE, "bh_get_body_counts: engine created without BH_FLAG_BODY_COUNTS");
    E_TRY(cudaSetDevice(e->device));
    if (e->n > 0) {
        if (interactions) E_TRY(cudaMemcpy(interactions, e->cntI, (size_t)e->n * 4, cudaMemcpyDeviceToHost));
        if (opened) E_TRY(cudaMemcpy(opened, e->cntO, (size_t)e->n * 4, cudaMemcpyDeviceToHost));
    }
    return BH_OK;
}

int bh_slice_bounds(int64_t n, int32_t world, int32_t rank, int64_t* lo, int64_t* hi) {
    if (n < 0 || world < 1 || rank < 0 || rank >= world || !lo || !hi) return BH_E_ARG;
    const int64_t per = (n + world - 1) / world;
    *lo = std::min<int64_t>(n, per * rank);
    *hi = std::min<int64_t>(n, per * (rank + 1));
    return BH_OK;
}

int bh_measure_fp32_tflops(int32_t device, double* tflops) {
    if (!tflops) return BH_E_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return BH_E_CUDA;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    float* d = nullptr;
    cudaEvent_t a, b;
    if (cudaMalloc(&d, 4) != cudaSuccess || cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return BH_E_CUDA;
    const int blocks = sms * 8, iters = 4096;
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(a);
        k_fp32_peak<<<blocks, 256>>>(d, iters, 1.0001f, 0.5f);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        const double fl = 2.0 * 64.0 * (double)iters * 256.0 * blocks;
        if (ms > 0.f) best = std::max(best, fl / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(d);
    if (cudaGetLastError() != cudaSuccess) return BH_E_CUDA;
    *tflops = best;
    return BH_OK;
}

int bh_comm_unique_id(void*, int32_t) { return BH_E_UNSUPPORTED; }
int bh_comm_init(bh_engine* e, int32_t, int32_t, const void*, int32_t) {
    return e ? e->fail(BH_E_UNSUPPORTED, "bh_comm_init: not implemented yet") : BH_E_ARG;
}

}  // extern "C"
