// bh_sort.cuh — CUB-free onesweep LSD radix sort (64-bit keys, 32-bit values) and a
// single-pass decoupled-look-back exclusive scan, hand-written for sm_100a.
//
// Onesweep (Adinets & Merrill 2022): one upfront kernel builds the digit histograms of
// ALL passes; then each pass is ONE kernel in which every tile ranks its keys locally,
// publishes its per-digit counts, resolves its global offsets by looking back at the
// tiles before it (decoupled look-back), and scatters.  Per pass each key/value is read
// once and written once: 24 B/element/pass + 8 B/element for the histogram read.
//
// Forward progress of the look-back: tiles take their index from an atomic ticket, so a
// tile only ever waits on tiles whose blocks are already running or finished.
#ifndef BH_SORT_CUH
#define BH_SORT_CUH

#include <stdint.h>
#include <cuda_runtime.h>

namespace bhsort {

constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;
constexpr int SORT_THREADS = 256;           // == RADIX: thread d owns digit d in the tile
constexpr int SORT_WARPS = SORT_THREADS / 32;
constexpr int SORT_IPT = 12;                // keys per thread
constexpr int SORT_TILE = SORT_THREADS * SORT_IPT;
constexpr int MAX_PASSES = 8;
constexpr int LOOKBACK_BATCH = 8;

constexpr uint32_t FLAG_SHIFT = 30;
constexpr uint32_t FLAG_AGG = 1u << FLAG_SHIFT;     // tile aggregate available
constexpr uint32_t FLAG_PREFIX = 2u << FLAG_SHIFT;  // inclusive prefix available
constexpr uint32_t VALUE_MASK = (1u << FLAG_SHIFT) - 1u;

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---- upfront histograms of every pass -----------------------------------------------------
// hist[pass][digit]; n is read from device memory (*n_ptr) so no host sync is needed.
__global__ void __launch_bounds__(256) k_histogram(const uint64_t* __restrict__ keys, const int* __restrict__ n_ptr,
                                                    int n_static, int passes, uint32_t* __restrict__ hist) {
    __shared__ uint32_t sh[MAX_PASSES * RADIX];
    for (int k = threadIdx.x; k < passes * RADIX; k += blockDim.x) sh[k] = 0;
    __syncthreads();
    const int n = n_ptr ? *n_ptr : n_static;
    const int stride = gridDim.x * blockDim.x;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t k = keys[i];
#pragma unroll
        for (int p = 0; p < MAX_PASSES; ++p)
            if (p < passes) atomicAdd(&sh[p * RADIX + (int)((k >> (p * RADIX_BITS)) & (RADIX - 1))], 1u);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < passes * RADIX; k += blockDim.x)
        if (sh[k]) atomicAdd(&hist[k], sh[k]);
}

// exclusive scan of each pass's 256-bin histogram; one block (256 threads) per pass
__global__ void __launch_bounds__(RADIX) k_histogram_scan(uint32_t* __restrict__ hist) {
    __shared__ uint32_t warp_tot[RADIX / 32];
    uint32_t* h = hist + blockIdx.x * RADIX;
    const int d = threadIdx.x, lane = d & 31, w = d >> 5;
    const uint32_t v = h[d];
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[w] = inc;
    __syncthreads();
    uint32_t base = 0;
    for (int k = 0; k < w; ++k) base += warp_tot[k];
    h[d] = base + inc - v;
}

// ---- one onesweep pass ----------------------------------------------------------------------
// Sorts by the digit at `shift`.  vals_in == nullptr means "value = element index".
// `ticket` and `lookback` (tiles*RADIX words) must be zero on entry.
// Per tile: warp-striped load -> stable in-warp ranking by digit (match.any multi-split) ->
// per-digit scan over the warps + decoupled look-back -> the tile is re-ordered by digit in
// SHARED MEMORY -> the scatter to HBM writes runs of consecutive addresses (coalesced) instead of
// one 8-byte store per key wherever its digit sends it.
__global__ void __launch_bounds__(SORT_THREADS)
k_onesweep_pass(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, int n, int shift,
                const uint32_t* __restrict__ digit_base /* exclusive-scanned histogram of this pass */,
                uint32_t* __restrict__ ticket, uint32_t* __restrict__ lookback) {
    __shared__ uint64_t s_keys[SORT_TILE];          // re-used for the values (as uint32) afterwards
    __shared__ uint32_t s_hist[SORT_WARPS][RADIX];
    __shared__ uint32_t s_base[RADIX];              // global position of the tile's first key of digit d
    __shared__ uint32_t s_dstart[RADIX];            // position of digit d inside the re-ordered tile
    __shared__ uint32_t s_wsum[SORT_WARPS];
    __shared__ uint32_t s_tile;

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
#pragma unroll
    for (int k = 0; k < RADIX / 32; ++k) s_hist[w][lane + 32 * k] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    const int64_t tile_base = (int64_t)tile * SORT_TILE;
    if (tile_base >= n) return;   // surplus block (whole block exits together)
    const int tile_n = (int)((n - tile_base < SORT_TILE) ? (n - tile_base) : SORT_TILE);

    // warp-striped load: warp w owns SORT_IPT*32 consecutive elements
    uint64_t key[SORT_IPT];
    uint32_t val[SORT_IPT];
    uint32_t rank[SORT_IPT];
    const int64_t warp_base = tile_base + (int64_t)w * (SORT_IPT * 32);
#pragma unroll
    for (int j = 0; j < SORT_IPT; ++j) {
        const int64_t idx = warp_base + j * 32 + lane;
        const bool ok = idx < n;
        key[j] = ok ? keys_in[idx] : ~0ull;
        val[j] = ok ? (vals_in ? vals_in[idx] : (uint32_t)idx) : 0u;
    }

    // stable in-warp ranking by digit (match.any multi-split), counts in the warp's histogram.
    // Padding keys (~0: digit 255) of the last tile rank behind every real key of their digit.
    const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int j = 0; j < SORT_IPT; ++j) {
        const int d = (int)((key[j] >> shift) & (RADIX - 1));
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        const uint32_t prior = s_hist[w][d];
        rank[j] = prior + __popc(peers & lt_mask);
        __syncwarp();
        if (lane == __ffs(peers) - 1) s_hist[w][d] = prior + __popc(peers);
        __syncwarp();
    }
    __syncthreads();

    // thread d: exclusive scan of digit d over the warps, tile count, look-back
    uint32_t dsum;
    {
        const int d = tid;
        uint32_t sum = 0;
#pragma unroll
        for (int k = 0; k < SORT_WARPS; ++k) { const uint32_t c = s_hist[k][d]; s_hist[k][d] = sum; sum += c; }
        dsum = sum;
        // the padding keys of a partial tile were counted under digit 255: not part of the data
        const uint32_t real = (d == RADIX - 1) ? sum - (uint32_t)(SORT_TILE - tile_n) : sum;
        uint32_t* mine = lookback + (size_t)tile * RADIX + d;
        uint32_t excl = 0;
        if (tile == 0) {
            st_volatile_u32(mine, FLAG_PREFIX | real);
        } else {
            st_volatile_u32(mine, FLAG_AGG | real);
            // walk back over the predecessors, LOOKBACK_BATCH independent loads in flight at a time
            // (when a whole wave of tiles starts together the walk is long: one L2 round trip per
            // predecessor would serialise the pass)
            int64_t t = (int64_t)tile - 1;
            bool done = false;
            while (!done) {
                uint32_t v[LOOKBACK_BATCH];
#pragma unroll
                for (int k = 0; k < LOOKBACK_BATCH; ++k)
                    v[k] = (t - k >= 0) ? ld_volatile_u32(lookback + (size_t)(t - k) * RADIX + d) : FLAG_PREFIX;
                int used = 0;
#pragma unroll
                for (int k = 0; k < LOOKBACK_BATCH; ++k) {
                    if (done || used != k) continue;
                    const uint32_t f = v[k] >> FLAG_SHIFT;
                    if (f == 0) continue;            // not published yet: re-read from here
                    excl += v[k] & VALUE_MASK;
                    used = k + 1;
                    if (f == 2) done = true;         // inclusive prefix: done
                }
                t -= used;
            }
            st_volatile_u32(mine, FLAG_PREFIX | ((excl + real) & VALUE_MASK));
        }
        s_base[d] = digit_base[d] + excl;
    }
    // exclusive scan of the tile's digit counts: where digit d starts in the re-ordered tile
    {
        uint32_t inc = dsum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) s_wsum[w] = inc;
        __syncthreads();
        uint32_t wb = 0;
#pragma unroll
        for (int k = 0; k < SORT_WARPS; ++k) if (k < w) wb += s_wsum[k];
        s_dstart[tid] = wb + inc - dsum;
    }
    __syncthreads();

    // re-order the tile by digit in shared memory
    uint32_t lp[SORT_IPT];
#pragma unroll
    for (int j = 0; j < SORT_IPT; ++j) {
        const int d = (int)((key[j] >> shift) & (RADIX - 1));
        lp[j] = s_dstart[d] + s_hist[w][d] + rank[j];
        s_keys[lp[j]] = key[j];
    }
    __syncthreads();
    // coalesced scatter: consecutive threads hold consecutive keys of the same digit
    uint32_t dst[SORT_IPT];
#pragma unroll
    for (int j = 0; j < SORT_IPT; ++j) {
        const int idx = tid + j * SORT_THREADS;
        dst[j] = 0xffffffffu;
        if (idx < tile_n) {
            const uint64_t k = s_keys[idx];
            const int d = (int)((k >> shift) & (RADIX - 1));
            dst[j] = s_base[d] + ((uint32_t)idx - s_dstart[d]);
            keys_out[dst[j]] = k;
        }
    }
    __syncthreads();
    uint32_t* s_vals = reinterpret_cast<uint32_t*>(s_keys);
#pragma unroll
    for (int j = 0; j < SORT_IPT; ++j) s_vals[lp[j]] = val[j];
    __syncthreads();
#pragma unroll
    for (int j = 0; j < SORT_IPT; ++j) {
        const int idx = tid + j * SORT_THREADS;
        if (idx < tile_n) vals_out[dst[j]] = s_vals[idx];
    }
}

inline int sort_tiles(int64_t n) { return (int)((n + SORT_TILE - 1) / SORT_TILE); }
// words of scratch (hist + tickets + look-back) a sort of n keys with `passes` passes needs
inline size_t sort_scratch_words(int64_t n, int passes) {
    return (size_t)MAX_PASSES * RADIX + MAX_PASSES + (size_t)passes * sort_tiles(n) * RADIX;
}

// Sort n (key,index) pairs by the low `key_bits` bits.  Buffers ping-pong a -> b -> a ...;
// returns 0 if the result is in (keys_a, vals_a), 1 if in (keys_b, vals_b).  The first
// pass synthesises values = 0..n-1.  `scratch` must hold sort_scratch_words() words.
inline int onesweep_sort(uint64_t* keys_a, uint32_t* vals_a, uint64_t* keys_b, uint32_t* vals_b, int64_t n,
                         int key_bits, uint32_t* scratch, cudaStream_t st, int num_sms) {
    const int passes = (key_bits + RADIX_BITS - 1) / RADIX_BITS;
    if (n <= 0 || passes <= 0) return 0;
    const int tiles = sort_tiles(n);
    uint32_t* hist = scratch;
    uint32_t* tickets = scratch + MAX_PASSES * RADIX;
    uint32_t* lookback = tickets + MAX_PASSES;
    cudaMemsetAsync(scratch, 0, sort_scratch_words(n, passes) * sizeof(uint32_t), st);
    int hb = (int)((n + 256 * 8 - 1) / (256 * 8));
    if (hb > num_sms * 8) hb = num_sms * 8;
    if (hb < 1) hb = 1;
    k_histogram<<<hb, 256, 0, st>>>(keys_a, nullptr, (int)n, passes, hist);
    k_histogram_scan<<<passes, RADIX, 0, st>>>(hist);
    int cur = 0;
    for (int p = 0; p < passes; ++p) {
        const uint64_t* kin = cur ? keys_b : keys_a;
        const uint32_t* vin = (p == 0) ? nullptr : (cur ? vals_b : vals_a);
        uint64_t* kout = cur ? keys_a : keys_b;
        uint32_t* vout = cur ? vals_a : vals_b;
        k_onesweep_pass<<<tiles, SORT_THREADS, 0, st>>>(kin, vin, kout, vout, (int)n, p * RADIX_BITS,
                                                         hist + p * RADIX, tickets + p,
                                                         lookback + (size_t)p * tiles * RADIX);
        cur ^= 1;
    }
    return cur;
}

}  // namespace bhsort
#endif  // BH_SORT_CUH
