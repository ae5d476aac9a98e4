// bh_let.cuh — multi-GPU "domain" mode: every rank builds only the tree of its own Morton range and
// walks its bodies over a LOCALLY ESSENTIAL TREE (bh_let_core.h) instead of a replicated tree.
//
// Per force evaluation on rank r (all on the engine's stream, NCCL over NVLink):
//   1. k_let_local        own slice -> local source arrays; footprint bitmap; strays (own bodies whose
//                         key left the rank's code range since the last re-homing) -> the rank's segment
//   2. ncclAllGather      segments (stray count, footprint, strays); k_let_guests takes the strays that
//                         fell into this rank's range as extra SOURCES of its local tree
//   3. build()            the unchanged single-GPU build on the local arrays (keys outside the range get
//                         the not-in-tree sentinel: strays stay targets, like bodies outside the root box)
//   4. k_let_summary      level-ELL summaries of the own codes; ncclAllReduce makes the table replicated
//   5. k_let_plan + scans which blocks this rank needs / must send (deterministic: every rank derives the
//                         same plan from the same table and footprints), items of the top tree
//   6. k_let_pack, ncclSend/ncclRecv   blocks to the ranks that may open them
//   7. k_let_emit, k_let_blocks, k_let_climb   top tree, own + imported blocks, exact f64 climb
//   8. k_walk over the LET: the same cells in the same order as over the global tree -> bit-identical
// Every R steps (re-homing) the state is all-gathered once, the replicated build re-sorts it, and the
// slices are re-cut at code boundaries (k_let_cut).
#ifndef BH_LET_CUH
#define BH_LET_CUH

namespace {

__device__ __forceinline__ void let_atomic_min_d(double* p, double v) {
    unsigned long long* a = reinterpret_cast<unsigned long long*>(p);
    unsigned long long old = *a;
    while (__longlong_as_double((long long)old) > v) {
        const unsigned long long was = atomicCAS(a, old, (unsigned long long)__double_as_longlong(v));
        if (was == old) break;
        old = was;
    }
}
__device__ __forceinline__ void let_atomic_max_d(double* p, double v) {
    unsigned long long* a = reinterpret_cast<unsigned long long*>(p);
    unsigned long long old = *a;
    while (__longlong_as_double((long long)old) < v) {
        const unsigned long long was = atomicCAS(a, old, (unsigned long long)__double_as_longlong(v));
        if (was == old) break;
        old = was;
    }
}

// segment of one rank in the all-gather buffer, in doubles:
//   [0] strays in the segment  [1] 1 = more strays than the segment holds  [2..3] -
//   [4..7] bounding box of the own bodies outside the root box  [8 .. 8+bw) footprint bitmap
//   then cap x (x, y, m, user index)
constexpr int LET_SEG_HDR = 8;
enum { LET_D_STRAYS = 0, LET_D_GUESTS = 1, LET_D_FLAG = 2, LET_D_JRET = 3, LET_D_DESCENTS = 4, LET_D_SCANS = 5, LET_D_COUNT = 6 };
// jitter-return slots per host rank behind the level summaries in the all-reduced table (see k_let_jitter_returns)
constexpr int LET_JRET = 256;

__global__ void k_let_seg_init(double* __restrict__ seg, int bw, int* __restrict__ dcnt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < LET_SEG_HDR + bw) {
        double v = 0.0;
        if (i == 4 || i == 6) v = 1e300;
        if (i == 5 || i == 7) v = -1e300;
        seg[i] = v;
    }
    if (i < LET_D_COUNT) dcnt[i] = 0;
}

// own slice -> local arrays, footprint, strays
__global__ void __launch_bounds__(256)
k_let_local(int n_own, const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ m,
            const int* __restrict__ perm, BhRoot root, BhGrid grid, int ell, int lam, uint32_t c_lo, uint32_t c_hi, int cap,
            double* __restrict__ lx, double* __restrict__ ly, double* __restrict__ lm, int* __restrict__ lperm,
            double* __restrict__ seg, int bw, int* __restrict__ dcnt, int* __restrict__ stray_slot) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_own) return;
    const double px = x[j], py = y[j], pm = m[j];
    const int pu = perm[j];
    lx[j] = px; ly[j] = py; lm[j] = pm; lperm[j] = pu;
    if (!bh_root_contains(root, px, py)) {
        let_atomic_min_d(seg + 4, px); let_atomic_max_d(seg + 5, px);
        let_atomic_min_d(seg + 6, py); let_atomic_max_d(seg + 7, py);
        return;
    }
    const uint64_t key = grid.exact ? bh_morton_key_grid(grid, root.levels, px, py) : bh_morton_key(root, px, py);
    // footprint: the lanes of a warp are Morton neighbours, so one lane per distinct coarse cell sets the bit
    const uint32_t q = bh_let_code(key, root.levels, lam);
    uint32_t* bits = reinterpret_cast<uint32_t*>(seg + LET_SEG_HDR);
    const unsigned same = __match_any_sync(__activemask(), q);
    if ((threadIdx.x & 31) == (unsigned)(__ffs(same) - 1)) {
        const uint32_t bit = 1u << (q & 31u);
        if (!(__ldcg(&bits[q >> 5]) & bit)) atomicOr(&bits[q >> 5], bit);
    }
    const uint32_t c = bh_let_code(key, root.levels, ell);
    if (c < c_lo || c >= c_hi) {
        const int k = atomicAdd(&dcnt[LET_D_STRAYS], 1);
        if (k < cap) {
            double* e = seg + LET_SEG_HDR + bw + 4 * (size_t)k;
            e[0] = px; e[1] = py; e[2] = pm; e[3] = (double)pu;
            stray_slot[k] = j;       // segment entry k is own body j (a host may send its position back, k_let_jitter_returns)
        }
    }
}

__global__ void k_let_seg_header(double* __restrict__ seg, const int* __restrict__ dcnt, int cap) {
    const int c = dcnt[LET_D_STRAYS];
    seg[0] = (double)(c < cap ? c : cap);
    seg[1] = c > cap ? 1.0 : 0.0;
}

// the strays of the other ranks that fell into this rank's code range become sources of its tree
__global__ void __launch_bounds__(256)
k_let_guests(const double* __restrict__ segs, int64_t seg_len, int bw, int cap, int world, int me, BhRoot root, BhGrid grid,
             int ell, uint32_t c_lo, uint32_t c_hi, int n_own, int max_guests, double* __restrict__ lx, double* __restrict__ ly,
             double* __restrict__ lm, int* __restrict__ lperm, int* __restrict__ dcnt, int* __restrict__ gsrc) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int q = (int)(t / cap), k = (int)(t % cap);
    if (q >= world || q == me) return;
    const double* seg = segs + (size_t)q * seg_len;
    if (k >= (int)seg[0]) return;
    const double* e = seg + LET_SEG_HDR + bw + 4 * (size_t)k;
    const double px = e[0], py = e[1];
    const uint64_t key = grid.exact ? bh_morton_key_grid(grid, root.levels, px, py) : bh_morton_key(root, px, py);
    const uint32_t c = bh_let_code(key, root.levels, ell);
    if (c < c_lo || c >= c_hi) return;
    const int g = atomicAdd(&dcnt[LET_D_GUESTS], 1);
    if (g >= max_guests) return;   // cannot happen: max_guests = all strays of the other ranks
    lx[n_own + g] = px; ly[n_own + g] = py; lm[n_own + g] = e[2]; lperm[n_own + g] = (int)e[3];
    gsrc[g] = q * cap + k;           // guest g is entry k of rank q's segment
}

// A guest inside a jitter cluster: the replay of BH.kt:145-156 on this (host) rank mutated a body another rank owns.
// The mutated position goes back to its home rank in one of this rank's LET_JRET return slots behind the level
// summaries of the all-reduced table: (count = 1, mass = home rank, comx/comy = new position, pos = entry of the
// body in its home rank's segment).  More such guests than slots: retry flag (the evaluation is redone after a
// re-homing, after which there are no strays).
__global__ void k_let_jitter_returns(const uint64_t* __restrict__ keys, const int* __restrict__ order, int n_in, int n_own,
                                     const double* __restrict__ lx, const double* __restrict__ ly, const int* __restrict__ gsrc,
                                     int cap, BhLetEntry* __restrict__ slots, int* __restrict__ dcnt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_in || order[i] < n_own) return;
    const uint64_t k = keys[i];
    if (!((i > 0 && keys[i - 1] == k) || (i + 1 < n_in && keys[i + 1] == k))) return;
    const int j = atomicAdd(&dcnt[LET_D_JRET], 1);
    if (j >= LET_JRET) { atomicOr(&dcnt[LET_D_FLAG], 1); return; }
    const int b = order[i], src = gsrc[b - n_own];
    BhLetEntry e;
    e.count = 1.0; e.mass = (double)(src / cap); e.comx = lx[b]; e.comy = ly[b]; e.pos = (double)(src % cap); e.size = 0.0;
    slots[j] = e;
}
// home side: take the positions the hosts sent back for this rank's strays (after the table all-reduce)
__global__ void k_let_apply_returns(const BhLetEntry* __restrict__ slots, int n_slots, int me, const int* __restrict__ stray_slot,
                                    double* __restrict__ x, double* __restrict__ y, double* __restrict__ lx, double* __restrict__ ly) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_slots) return;
    const BhLetEntry e = slots[i];
    if (e.count == 0.0 || (int)e.mass != me) return;
    const int j = stray_slot[(int)e.pos];
    x[j] = e.comx; y[j] = e.comy; lx[j] = e.comx; ly[j] = e.comy;
}

__global__ void k_let_summary(BhTreeView t, int levels, int ell, const double* __restrict__ lx, const double* __restrict__ ly,
                              const double* __restrict__ lm, const int* __restrict__ jflag, BhLetEntry* __restrict__ table) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= t.n_in) return;
    const int b = t.order[i];
    bh_let_summary_body(t, levels, ell, i, lx[b], ly[b], (jflag && (jflag[b] & 1)) ? 0.0 : lm[b], table);
}

// this rank's retry flag: a guest in a jitter cluster, or some rank had more strays than a segment holds
__global__ void k_let_flag(BhLetEntry* __restrict__ e, const int* __restrict__ flag, const double* __restrict__ segs,
                           int64_t seg_len, int world) {
    // reason bits: 1 = a guest in a jitter cluster, 2 = stray overflow on some rank, 4 = the local tree outgrew the cell arrays
    int code = ((*flag & 1) ? 1 : 0) | ((*flag & 4) ? 4 : 0);
    for (int q = 0; q < world; ++q) if (segs[(size_t)q * seg_len + 1] != 0.0) code |= 2;
    e->count = (double)code;
}

struct LetSplit { uint32_t cs[17]; int world, me; };
__device__ __forceinline__ int let_owner(const LetSplit& s, uint32_t c) {
    int r = 0;
    while (r + 1 < s.world && c >= s.cs[r + 1]) ++r;
    return r;
}
__device__ __forceinline__ BhLetRegion let_region(const double* __restrict__ segs, int64_t seg_len, int r) {
    const double* seg = segs + (size_t)r * seg_len;
    BhLetRegion g;
    g.bits = reinterpret_cast<const uint32_t*>(seg + LET_SEG_HDR);
    g.oob = BhLetBox{seg[4], seg[5], seg[6], seg[7]};
    return g;
}

// per code: items and block size in THIS rank's LET, cells to receive, cells to send to every peer
__global__ void __launch_bounds__(256)
k_let_plan(const BhLetEntry* __restrict__ table, uint32_t ncodes, const double* __restrict__ segs, int64_t seg_len, LetSplit sp,
           double theta2, double soft2, BhRoot root, int ell, int* __restrict__ nit, int* __restrict__ blk,
           int* __restrict__ recvsz, int* __restrict__ sendsz) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncodes) return;
    const BhLetEntry e = table[c];
    const uint32_t c_lo = sp.cs[sp.me], c_hi = sp.cs[sp.me + 1];
    int ni, B;
    bh_let_plan_code(e, c, c_lo, c_hi, let_region(segs, seg_len, sp.me), theta2, soft2, root, ell, &ni, &B);
    nit[c] = ni; blk[c] = B;
    const bool mine = c >= c_lo && c < c_hi;
    recvsz[c] = (!mine && ni == 2 && B > 1) ? B - 1 : 0;
    if (mine) {
        const uint32_t mylen = c_hi - c_lo;
        for (int r = 0; r < sp.world; ++r) {
            int v = 0;
            if (r != sp.me && e.count >= 2.0 &&
                bh_let_near_region(let_region(segs, seg_len, r), c, e.comx, e.comy, theta2, soft2, root, ell))
                v = (int)e.size - 1;
            sendsz[(size_t)r * mylen + (c - c_lo)] = v;
        }
    }
}

__global__ void __launch_bounds__(256)
k_let_items(uint32_t ncodes, const int* __restrict__ item_first, const int* __restrict__ blk, int levels, int ell,
            uint64_t* __restrict__ ikey, int* __restrict__ itype, int* __restrict__ iw) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncodes) return;
    const int j = item_first[c], ni = item_first[c + 1] - j;
    if (ni == 1) { ikey[j] = bh_let_item_key(c, BH_LET_SINGLE, levels, ell); itype[j] = BH_LET_SINGLE; iw[j] = 1; }
    if (ni == 2) {
        ikey[j] = bh_let_item_key(c, BH_LET_TWIN0, levels, ell); itype[j] = BH_LET_TWIN0; iw[j] = blk[c] - 1;
        ikey[j + 1] = bh_let_item_key(c, BH_LET_TWIN1, levels, ell); itype[j + 1] = BH_LET_TWIN1; iw[j + 1] = 0;
    }
}

// cnt(j) of every item; slots behind the last item are zeroed so that the scans can run over 2*ncodes
__global__ void __launch_bounds__(256)
k_let_item_cnt(const uint64_t* __restrict__ ikey, const int* __restrict__ n_items, int n_slots, int levels,
               int* __restrict__ icnt, int* __restrict__ iw) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_slots) return;
    const int n = *n_items;
    if (j < n) icnt[j] = bh_let_item_cnt(ikey, n, levels, j);
    else { icnt[j] = 0; iw[j] = 0; }
}

// up to 3 independent exclusive scans in one launch: block b scans in[b][0..n[b]) into out[b][0..n[b]] (out[n] = total)
struct LetScanJob { const int* in[3]; int* out[3]; int n[3]; };
__global__ void __launch_bounds__(1024) k_let_scans(LetScanJob job) {
    const int* __restrict__ in = job.in[blockIdx.x];
    int* __restrict__ out = job.out[blockIdx.x];
    const int n = job.n[blockIdx.x];
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    constexpr int IPT = 16;
    for (int base = 0; base < n; base += 1024 * IPT) {
        int v[IPT], sum = 0;
#pragma unroll
        for (int k = 0; k < IPT; ++k) { const int i = base + tid * IPT + k; v[k] = i < n ? in[i] : 0; sum += v[k]; }
        int inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) s_warp[w] = inc;
        __syncthreads();
        if (w == 0) {
            int x = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += t; }
            s_warp[lane] = x;     // inclusive over warps
        }
        __syncthreads();
        const int carry = s_carry;
        int run = carry + (w > 0 ? s_warp[w - 1] : 0) + inc - sum;
#pragma unroll
        for (int k = 0; k < IPT; ++k) { const int i = base + tid * IPT + k; if (i < n) out[i] = run; run += v[k]; }
        __syncthreads();
        if (tid == 1023) s_carry = carry + s_warp[31];
        __syncthreads();
    }
    if (tid == 0) out[n] = s_carry;
}

// the few numbers the host needs: items, LET cells, receive offset of every owner, send offset of every peer
__global__ void k_let_collect(const int* __restrict__ item_first, uint32_t ncodes, const int* __restrict__ iS,
                              const int* __restrict__ iW, const int* __restrict__ recvoff, const int* __restrict__ sendoff,
                              LetSplit sp, const int* __restrict__ dcnt, int* __restrict__ out,
                              const BhLetEntry* __restrict__ jret_slots) {
    const int t = threadIdx.x;
    const int n = item_first[ncodes];
    {   // positions coming back for this rank's strays (k_let_apply_returns will change x / y of own bodies)
        int mine = 0;
        for (int i = t; i < sp.world * LET_JRET; i += 32) mine += (jret_slots[i].count != 0.0 && (int)jret_slots[i].mass == sp.me);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
        if (t == 0) out[37] = mine;
    }
    if (t == 0) { out[0] = n; out[1] = iS[n] + iW[n]; out[36] = dcnt[LET_D_JRET]; out[38] = dcnt[LET_D_STRAYS]; out[39] = dcnt[LET_D_GUESTS]; }
    const uint32_t mylen = sp.cs[sp.me + 1] - sp.cs[sp.me];
    if (t <= sp.world) {
        out[2 + t] = recvoff[sp.cs[t]];
        out[2 + 17 + t] = sendoff[(size_t)t * mylen];
    }
}

// one warp per (peer, own code): the block's cells (all but the root) in wire format
__global__ void __launch_bounds__(256)
k_let_pack(const BhLetEntry* __restrict__ table, LetSplit sp, const int* __restrict__ sendsz, const int* __restrict__ sendoff,
           const BhCellD* __restrict__ cd, const BhCellS* __restrict__ sk, BhLetWire* __restrict__ sendbuf) {
    const uint32_t mylen = sp.cs[sp.me + 1] - sp.cs[sp.me];
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (wid >= (int64_t)sp.world * mylen) return;
    const int sz = sendsz[wid];
    if (sz == 0) return;
    const uint32_t c = sp.cs[sp.me] + (uint32_t)(wid % mylen);
    const int rp = (int)table[c].pos;
    BhLetWire* out = sendbuf + sendoff[wid];
    for (int j = 1 + lane; j <= sz; j += 32) out[j - 1] = bh_let_wire(cd, sk, rp + j, rp);
}

__global__ void __launch_bounds__(256)
k_let_emit(BhLetItems it, BhCellS* __restrict__ sk, int* __restrict__ arrived, int levels, int ell, int* __restrict__ ilp,
           int* __restrict__ dst) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= it.n) return;
    const int lp = bh_let_emit_item(it, sk, levels, j);
    ilp[j] = lp;
    {   // arrive counters of the internal cells this item owns (the column in front of its leaf region)
        const int first = it.S[j] + it.W[j];
        for (int p = first; p < lp; ++p) arrived[p] = 0;
    }
    const int type = it.type[j];
    if (type != BH_LET_TWIN1) dst[bh_let_code(it.key[j], levels, ell)] = (type == BH_LET_SINGLE) ? lp : lp - 1;
}

// the local-tree arrays of every rank, mapped into this process (CUDA IPC over NVLink peer memory)
struct LetPeers { const BhCellD* cd[16]; const BhCellS* sk[16]; };

// one warp per code: own blocks from the local arrays; imported blocks straight from the OWNER's local
// arrays over NVLink peer memory (peers != nullptr), else from the receive buffer of the NCCL exchange
// cells of a block are copied in CHUNKS of LET_BLK_CHUNK by one warp each: a level-ELL code of a dense galactic core
// holds 10^5 bodies, and one warp per block (as the uniform cloud allowed) serialised the whole phase on it
constexpr int LET_BLK_CHUNK = 256;
__global__ void k_let_block_chunks(const int* __restrict__ nit, const int* __restrict__ blk, uint32_t ncodes, int* __restrict__ nchunk) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncodes) return;
    const int B = blk[c];
    nchunk[c] = (nit[c] == 2 && B > 1) ? (B - 1 + LET_BLK_CHUNK - 1) / LET_BLK_CHUNK : 0;
}
__global__ void __launch_bounds__(256)
k_let_blocks(BhTreeView let, const BhLetEntry* __restrict__ table, uint32_t ncodes, LetSplit sp, const int* __restrict__ chunkoff,
             const int* __restrict__ blk, const int* __restrict__ dst, const int* __restrict__ recvoff,
             const BhLetWire* __restrict__ recvbuf, const BhCellD* __restrict__ cd, const BhCellS* __restrict__ sk, double half,
             LetPeers peers, int use_peers) {
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (wid >= chunkoff[ncodes]) return;
    // the code of chunk `wid`: last c with chunkoff[c] <= wid (exclusive scan; codes without chunks repeat the value)
    uint32_t lo = 0, hi = ncodes;
    while (hi - lo > 1) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (chunkoff[mid] <= (int)wid) lo = mid; else hi = mid;
    }
    const uint32_t c = lo;
    const int B = blk[c];
    const int j0 = 1 + ((int)wid - chunkoff[c]) * LET_BLK_CHUNK;
    const int j1 = min(B, j0 + LET_BLK_CHUNK);
    const int d = dst[c];
    if (c >= sp.cs[sp.me] && c < sp.cs[sp.me + 1]) {
        const int rp = (int)table[c].pos;
        for (int j = j0 + lane; j < j1; j += 32) bh_let_place(let, bh_let_wire(cd, sk, rp + j, rp), d, j, half);
    } else if (use_peers) {
        const int o = let_owner(sp, c);
        const BhCellD* __restrict__ pcd = peers.cd[o];
        const BhCellS* __restrict__ psk = peers.sk[o];
        const int rp = (int)table[c].pos;
        for (int j = j0 + lane; j < j1; j += 32) bh_let_place(let, bh_let_wire(pcd, psk, rp + j, rp), d, j, half);
    } else {
        const BhLetWire* in = recvbuf + recvoff[c];
        for (int j = j0 + lane; j < j1; j += 32) bh_let_place(let, in[j - 1], d, j, half);
    }
}

__global__ void __launch_bounds__(128)
k_let_climb(BhTreeView let, BhRoot root, BhLetItems it, const BhLetEntry* __restrict__ table, int levels, int ell,
            const int* __restrict__ ilp) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j == 0) bh_write_terminal_cell(let);
    if (j >= it.n) return;
    bh_let_climb_item(let, root, it, table, levels, ell, j, ilp[j]);
}

// preorder position of each own body's leaf in the LET (-1: outside the root box).  A stray's leaf was
// built by the rank that hosts it: it is found in the imported block by its exact coordinates (unique:
// a stray with a coincident twin is a jitter cluster, which forces a re-homing instead).
__global__ void __launch_bounds__(256)
k_let_leafpos(int n_own, const int* __restrict__ lleaf, const double* __restrict__ lx, const double* __restrict__ ly,
              BhRoot root, BhGrid grid, int ell, const BhLetEntry* __restrict__ table, const int* __restrict__ dst,
              const int* __restrict__ blk, BhTreeView let, int* __restrict__ leafpos, int* __restrict__ dcnt) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool live = j < n_own;
    const int p = live ? lleaf[j] : -1;
    const double px = live ? lx[j] : 0.0, py = live ? ly[j] : 0.0;
    int out = -1, d = -1, B = 0;
    bool search = false;
    if (live && (p >= 0 || bh_root_contains(root, px, py))) {
        const uint64_t key = grid.exact ? bh_morton_key_grid(grid, root.levels, px, py) : bh_morton_key(root, px, py);
        const uint32_t c = bh_let_code(key, root.levels, ell);
        d = dst[c];
        if (p >= 0) out = d + (p - (int)table[c].pos);
        else if (table[c].count == 1.0) out = d;                 // the stray is the only body of its code
        else { B = blk[c]; search = B > 1; }
    }
    // strays: the leaf with the stray's exact coordinates sits in the block its host rank built.  Descend the block by
    // the digits of the stray's key: all children of a cell are one level below it, and the child on the path is the
    // one whose own centre of mass (inside its box) has the wanted digit.  The centre of mass of a deep cell of a
    // dense core can round onto (or past) a box edge and name no child or the wrong one: then the leaf test at the
    // end fails, and the warp scans the subtree of the cell three levels above the point of failure — and only if
    // that fails too the whole block (a level-ELL code of a galactic core holds 10^5 cells: scanning it by default
    // cost 17 ms per evaluation at 100M bodies, for 21 strays).
    const int d_block = d, B_block = B;
    if (search) {
        atomicAdd(&dcnt[LET_D_DESCENTS], 1);
        const uint64_t key = grid.exact ? bh_morton_key_grid(grid, root.levels, px, py) : bh_morton_key(root, px, py);
        int q = d, qa1 = d, qa2 = d, qa3 = d;      // the last three cells of the path above q
        bool ok = true;
        while (ok && let.sk[q].skip != q + 1) {
            const BhCellS sq = let.sk[q];
            int next = -1;
            for (int ch = q + 1; ch < sq.skip; ch = let.sk[ch].skip) {
                const BhCellD cc = let.cd[ch];
                const uint64_t kc = grid.exact ? bh_morton_key_grid(grid, root.levels, cc.comx, cc.comy) : bh_morton_key(root, cc.comx, cc.comy);
                if (((kc ^ key) >> (2 * (root.levels - sq.level - 1))) == 0) { next = ch; break; }   // shares level + 1 digits
            }
            if (next < 0) ok = false; else { qa3 = qa2; qa2 = qa1; qa1 = q; q = next; }
        }
        if (ok && q > d && let.cd[q].comx == px && let.cd[q].comy == py) { out = q; search = false; }
        else { atomicAdd(&dcnt[LET_D_SCANS], 1); d = qa3; B = let.sk[qa3].skip - qa3; }
    }
    for (int stage = 0; stage < 2; ++stage) {
        unsigned todo = __ballot_sync(0xffffffffu, search);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const int sd = __shfl_sync(0xffffffffu, d, src), sB = __shfl_sync(0xffffffffu, B, src);
            const double sx = __shfl_sync(0xffffffffu, px, src), sy = __shfl_sync(0xffffffffu, py, src);
            int found = 0x7fffffff;
            for (int q = sd + 1 + lane; q < sd + sB; q += 32)
                if (let.sk[q].skip == q + 1 && let.cd[q].comx == sx && let.cd[q].comy == sy) { found = q; break; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) found = min(found, __shfl_xor_sync(0xffffffffu, found, o));
            if (lane == src && found != 0x7fffffff) { out = found; search = false; }
        }
        d = d_block; B = B_block;                  // second stage: whatever is still missing, in the whole block
    }
    if (live) leafpos[j] = out;
}

// slices of the freshly re-homed (globally sorted) state, cut at code boundaries
__global__ void k_let_cut(const uint64_t* __restrict__ keys, int n_in, int n, int levels, int ell, int world, int* __restrict__ out) {
    const int r = threadIdx.x;
    if (r > world) return;
    const uint32_t ncodes = 1u << (2 * ell);
    int cut = 0;
    uint32_t cs = 0;
    if (r == world) { cut = n; cs = ncodes; }
    else if (r > 0) {
        const int i = (int)((int64_t)n_in * r / world);
        if (i <= 0) cut = 0;
        else {
            const int sh = bh_prefix_shift(levels, ell);
            cut = bh_gallop_right(keys, n_in, i - 1, keys[i - 1] >> sh, sh) + 1;
        }
        cs = cut < n_in ? bh_let_code(keys[cut], levels, ell) : ncodes;
    }
    out[r] = cut;
    out[17 + r] = (int)cs;
}

int64_t g_let_grows = 0;
template <class T>
cudaError_t let_grow(T*& p, int64_t& cap, int64_t need) {
    if (need <= cap) return cudaSuccess;
    ++g_let_grows;
    if (p) cudaFree(p);
    p = nullptr;
    cap = need + need / 2 + 64;      // generous: cudaFree + cudaMalloc stall the stream (more so with peer mappings)
    return cudaMalloc((void**)&p, (size_t)cap * sizeof(T));
}

}  // namespace

struct bh_let_state {
    bool enabled = false;        // BH_FLAG_LET / BH_LET=1
    int min_world = 4;           // smallest world the domain mode is used for (BH_LET_MIN_WORLD): below it the replicated
                                 // build of the few-times-larger tree is cheaper than the fixed cost of assembling a LET
    bool part_valid = false;     // slices are cut at code boundaries (set by the re-homing build)
    bool pos_valid = true;       // positions of ALL bodies are current on this rank
    int ell = 0, lam = 0, bw = 0;
    uint32_t ncodes = 0;
    int64_t n_part = -1;         // body count the partition refers to
    int64_t n_declined = -1;     // body count for which let_partition declined (too few bodies)
    BhRoot part_root{};          // root box the partition refers to
    bool local_build = false;    // build() is running on the local arrays
    bool local_overflow = false; // ... and its tree did not fit the (peer-mapped) cell arrays
    bool view_valid = false;
    int64_t cut[17] = {0};
    LetSplit split{};
    int64_t stray_cap = 0, seg_len = 0;
    // local source arrays: own slice, then guests
    double *lx = nullptr, *ly = nullptr, *lm = nullptr;
    int *lperm = nullptr, *lleaf = nullptr;
    int *nchunk = nullptr, *chunkoff = nullptr; int64_t nchunk_cap = 0, chunkoff_cap = 0;   // k_let_blocks work list
    int *gsrc = nullptr, *stray_slot = nullptr;        // guest -> (rank, segment entry); own segment entry -> own body
    int64_t gsrc_cap = 0, stray_slot_cap = 0;
    int64_t lx_cap = 0, ly_cap = 0, lm_cap = 0, lperm_cap = 0, lleaf_cap = 0;
    double* segs = nullptr; int64_t segs_cap = 0;
    BhLetEntry* table = nullptr; int64_t table_cap = 0;
    int *nit = nullptr, *blk = nullptr, *recvsz = nullptr, *recvoff = nullptr, *item_first = nullptr, *dst = nullptr;
    int64_t nit_cap = 0, blk_cap = 0, recvsz_cap = 0, recvoff_cap = 0, item_first_cap = 0, dst_cap = 0;
    int *sendsz = nullptr, *sendoff = nullptr; int64_t sendsz_cap = 0, sendoff_cap = 0;
    uint64_t* ikey = nullptr; int64_t ikey_cap = 0;
    int *itype = nullptr, *iw = nullptr, *icnt = nullptr, *iS = nullptr, *iW = nullptr, *ilp = nullptr;
    int64_t itype_cap = 0, iw_cap = 0, icnt_cap = 0, iS_cap = 0, iW_cap = 0, ilp_cap = 0;
    BhLetWire *sendbuf = nullptr, *recvbuf = nullptr; int64_t sendbuf_cap = 0, recvbuf_cap = 0;
    BhCell* cell = nullptr; BhCellD* cd = nullptr; BhCellS* sk = nullptr; int* arrived = nullptr;
    int64_t cell_cap = 0, cd_cap = 0, sk_cap = 0, arrived_cap = 0;
    int* dcnt = nullptr;          // device counters (LET_D_*) + collect output
    int* dcollect = nullptr;
    int* hcollect = nullptr;      // pinned: 2 + 2*17 ints
    double* hhdr = nullptr;       // pinned: world x 2 doubles
    int* hcut = nullptr;          // pinned: 2*17 ints
    int M = 0, n_items = 0;
    // peer memory (CUDA IPC): every rank's local-tree arrays (cd, sk) mapped here
    struct IpcPair { cudaIpcMemHandle_t cd, sk; };
    bool ipc_wanted = true, ipc_ok = false;
    IpcPair ipc_cached[16] = {};
    bool ipc_open[16] = {false};
    void* ipc_ptr_cd[16] = {nullptr};
    void* ipc_ptr_sk[16] = {nullptr};
    IpcPair* ipc_host = nullptr;      // pinned, world entries
    double* ipc_dev = nullptr;        // world x sizeof(IpcPair) bytes (+ 1 double for the agreement flag)
    LetPeers peers{};
    // statistics of the last LET evaluation
    int64_t last_imported = 0, last_sent = 0, last_strays = 0, last_guests_max = 0, evaluations = 0, fallbacks = 0;
    bool returns_applied = false;   // the last evaluation changed positions of own strays (sent back by their host ranks)
    int64_t jret_total = 0;        // guest positions this rank sent back to their home ranks after a jitter replay
    int64_t fb_jitter = 0, fb_strays = 0, fb_cells = 0;   // fallbacks by reason (a fallback may have several)
    // phase timers of let_evaluate (CUDA events, folded at the start of the next evaluation)
    static constexpr int NPH = 14;
    cudaEvent_t pe[NPH + 1] = {};
    bool pe_armed = false;
    int64_t n_folds = 0;
    double ms_phase[NPH] = {0};
    double cpu_us[NPH] = {0};      // host time between the same points
    double cpu_t[NPH + 1] = {0};

    void release() {
        void* ptrs[] = {nchunk, chunkoff, gsrc, stray_slot, lx, ly, lm, lperm, lleaf, segs, table, nit, blk, recvsz, recvoff, item_first, dst, sendsz, sendoff, ikey,
                        itype, iw, icnt, iS, iW, ilp, sendbuf, recvbuf, cell, cd, sk, arrived, dcnt, dcollect};
        for (void* p : ptrs) if (p) cudaFree(p);
        if (hcollect) cudaFreeHost(hcollect);
        if (hhdr) cudaFreeHost(hhdr);
        if (hcut) cudaFreeHost(hcut);
        for (int q = 0; q < 16; ++q) if (ipc_open[q]) { cudaIpcCloseMemHandle(ipc_ptr_cd[q]); cudaIpcCloseMemHandle(ipc_ptr_sk[q]); }
        if (ipc_host) cudaFreeHost(ipc_host);
        if (ipc_dev) cudaFree(ipc_dev);
        for (auto& e : pe) if (e) cudaEventDestroy(e);
    }
};

#endif  // BH_LET_CUH
