// bh_engine.cu — B200 (sm_100a) Barnes–Hut physics-step engine behind include/bh_engine.h.
//
// One PhysicsEngine.step() of the reference (BarnesHutAlg.kt:405-439) becomes, per force
// evaluation:  k_keygen -> onesweep radix sort -> k_count_scan -> k_emit -> k_climb_block/top -> k_walk,
// then the f64 kick/drift kernels and the merge rule.  The tree is the reference's own quadtree
// (same cells, same f64 centres of mass, same per-body accept/open decisions) stored as a
// flattened DFS-preorder array with skip links; see bh_core.h, bh_kernels.cuh and DESIGN.md.
//
// Device-resident state is kept in HOME order: the Morton order of the last re-homing.  perm[h]
// is the position of home slot h in the reference's `bodies` list (USER order); the C ABI speaks
// user order only.  Home order makes the body reads / acceleration writes of the walk coalesced,
// keeps the lanes of a warp spatial neighbours, and gives every rank of a multi-GPU run a
// contiguous slice of targets.
//
// No CPU fallback: every compute entry point runs CUDA kernels or returns an error.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <new>
#include <string>
#include <vector>

#include "../../include/bh_engine.h"
#include "bh_core.h"
#include "bh_export.h"
#include "bh_sort.cuh"
#include "bh_kernels.cuh"
#include "bh_comm.cuh"
#include "bh_let_core.h"
#include "bh_let.cuh"

namespace {

thread_local std::string g_create_err;

template <class T>
cudaError_t dev_alloc(T** p, size_t count) {
    *p = nullptr;
    return cudaMalloc((void**)p, std::max<size_t>(count, 1) * sizeof(T));
}
template <class T>
void dev_free(T*& p) { if (p) cudaFree(p); p = nullptr; }

inline int grid_for(int64_t n, int block) { return (int)((n + block - 1) / block); }

constexpr int PAD = 64;            // slack behind every body array (in-place all-gather of padded slices)
constexpr int MAX_WORLD = 64;
enum HostFlag { HF_SPARE = 0, HF_N_DEAD = 1, HF_N_CAND = 2, HF_N_ROOTS = 3, HF_COUNT = 4 };

}  // namespace

// ---------------------------------------------------------------------------------------
// engine
// ---------------------------------------------------------------------------------------
struct bh_engine {
    bh_config cfg{};
    bh_params par{};
    int device = 0, num_sms = 148;
    cudaStream_t st = nullptr;
    // CUDA-event timing slots: a ring, so that a multi-step bh_step never waits for the device just
    // to read its timers.  Per slot: [0..3] evaluation A, [4..7] evaluation B, [8..9] step,
    // [12..13] exchange, [14..15] merge rule.  `ev` = the slot in use; call_ev = whole call.
    static constexpr int EV_RING = 4;
    struct EvSlot { cudaEvent_t e[16]{}; bool used = false, has_comm = false, has_merge = false, phases = true; };
    EvSlot ring[EV_RING];
    EvSlot* cur = &ring[0];
    cudaEvent_t* ev = ring[0].e;
    cudaEvent_t call_ev[2]{};
    std::string err;

    // body state, f64 SoA in HOME order
    int64_t n = 0, cap = 0;
    double *x = nullptr, *y = nullptr, *vx = nullptr, *vy = nullptr, *m = nullptr, *ax = nullptr, *ay = nullptr;
    double* dtmp = nullptr;          // staging for permutations / user-order transfers
    int *perm = nullptr;             // home slot -> user index
    int *origin = nullptr;           // user index -> index in the list of the last bh_set_bodies
    int *leafpos = nullptr;          // home slot -> preorder position of the body's leaf (-1: not in the tree)
    int *itmp = nullptr, *inv = nullptr, *iscr0 = nullptr, *iscr1 = nullptr, *dead = nullptr;
    int *jflag = nullptr;            // jitter replay flags per home slot (valid iff jitter_active)
    bool jitter_active = false;      // the last build replayed jitter clusters
    int *cntI = nullptr, *cntO = nullptr;
    bool perm_identity = true, origin_identity = true;
    int* dflags = nullptr;           // device flags / small counters (HostFlag)
    int* hflags = nullptr;           // pinned mirror

    // re-homing
    bool rehome_due = true;
    int steps_since_rehome = 0;
    int rehome_interval = 8;

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
    BhClimbRoot* climb_roots = nullptr;   // queue of finished local subtrees (k_climb_block -> k_climb_top)
    int64_t climb_roots_cap = 0;

    bool tree_valid = false;
    // sync-free builds + per-step CUDA graphs (small body counts, one GPU): see build() and run_steps()
    static constexpr int64_t SYNCFREE_MAX_N = 262144;
    bool syncfree_enabled = true;    // BH_SYNCFREE=0 switches both off
    bool graph_enabled = true;       // BH_GRAPH=0: sync-free builds, but every kernel launched on its own
    bool counts_on_host = true;      // n_in / M / jitter_active below describe the last build (false: still on the device)
    bool capturing = false;          // the launches of the current step are being captured into a graph
    cudaGraphExec_t gexec = nullptr;
    int64_t ctr_graph_steps = 0, ctr_graph_instantiations = 0;
    BhRoot root{};
    int n_in = 0, n_internal = 0, M = 0;
    const uint64_t* keys_sorted = nullptr;
    const int* order = nullptr;      // sorted position -> home slot

    // merge rule
    int* heavy = nullptr;            // home slots of the bodies with m > merge_max_mass, ascending user index
    int64_t heavy_cap = 0;
    int n_heavy = 0;
    bool heavies_valid = false;
    double heavies_max_mass = 0.0;

    // multi-process
    enum Transport { T_NONE, T_NCCL, T_EXTERNAL };
    Transport transport = T_NONE;
    int rank = 0, world = 1;
    bhcomm::Comm comm = nullptr;
    bool vel_valid = true;           // velocities of ALL bodies are current on this rank
    bool mass_valid = true;          // masses of ALL bodies are current on this rank (bh_step_io_slice replaces a slice only)
    int phase = 0;                   // 0 idle, 1 after step_begin, 2 after step_end
    bool acc_valid = false;          // ax/ay of this rank's slice = a(current positions, current params)
    bh_params acc_par{};             // parameters acc_valid refers to

    // bh_step_io: transfers overlapped with the compute
    double* io_stage[4] = {nullptr, nullptr, nullptr, nullptr};   // staging for (vx, vy, m) in / (x, y, m, vy) out
    int64_t io_cap = 0;
    cudaEvent_t io_ev[2]{};          // [0] (vx, vy, m) arrived, [1] final positions written
    bool io_wait_in = false;         // the compute stream has not yet waited for (vx, vy, m)
    struct IoOut { double *x = nullptr, *y = nullptr, *m = nullptr; bool armed = false, slice = false; int64_t epoch = -1; } io_out;
    int wait_inputs() {              // before the first kernel that reads vx / vy / m
        if (io_wait_in) {
            const cudaError_t ce = cudaStreamWaitEvent(st, io_ev[0], 0);
            if (ce != cudaSuccess) return cuda_fail(ce, "cudaStreamWaitEvent");
            io_wait_in = false;
        }
        return BH_OK;
    }
    int emit_positions_out();        // after the last drift: (x, y, m) -> host on the copy stream

    // asynchronous render read-back (bh_request_positions_f32): dedicated staging + copy stream
    cudaStream_t copy_st = nullptr;
    cudaEvent_t snap_ev[2]{};        // [0] snapshot written (main stream), [1] copy finished (copy stream)
    float2* snap_xy = nullptr; float* snap_m = nullptr;      // device staging
    float* snap_hxy = nullptr; float* snap_hm = nullptr;     // pinned host
    int64_t snap_cap = 0, snap_n = -1;

    bh_counters ctr{};

    // locally-essential-tree mode (bh_let.cuh)
    bh_let_state let;
    bool let_usable() const;
    bool let_ready() const;
    int sync_positions();
    int let_partition();
    int let_map_peers();
    int let_evaluate(int slot);
    int evaluate_slice(int slot);

    int fail(int code, const char* what) { err = what; return code; }
    int fail(int code, const std::string& what) { err = what; return code; }
    int cuda_fail(cudaError_t e, const char* what) {
        err = std::string(what) + ": " + cudaGetErrorString(e);
        return e == cudaErrorMemoryAllocation ? BH_E_OOM : BH_E_CUDA;
    }

    static constexpr size_t SC_WORDS = (sizeof(DevScalars) + 3) / 4;
    DevScalars* sc() const { return reinterpret_cast<DevScalars*>(scratch); }
    uint32_t* sort_scratch() const { return scratch + SC_WORDS; }
    size_t scan_tiles(int64_t nn) const { return (size_t)((nn + SCAN_TILE - 1) / SCAN_TILE) + 1; }

    void free_bodies() {
        dev_free(x); dev_free(y); dev_free(vx); dev_free(vy); dev_free(m); dev_free(ax); dev_free(ay); dev_free(dtmp);
        dev_free(perm); dev_free(origin); dev_free(leafpos); dev_free(itmp); dev_free(inv); dev_free(iscr0);
        dev_free(iscr1); dev_free(dead); dev_free(jflag);
        dev_free(cntI); dev_free(cntO);
        dev_free(keys_a); dev_free(keys_b); dev_free(vals_a); dev_free(vals_b); dev_free(scratch); dev_free(S);
        cap = 0;
    }
    void free_cells() {
        dev_free(cell); dev_free(cd); dev_free(sk); dev_free(arrived); dev_free(climb_roots);
        cell_cap = 0; climb_roots_cap = 0;
    }

#define BH_TRY(expr)                                                        \
    do {                                                                    \
        cudaError_t _e = (expr);                                            \
        if (_e != cudaSuccess) return cuda_fail(_e, #expr);                 \
    } while (0)
#define BH_RC(expr)                                                         \
    do {                                                                    \
        const int _rc = (expr);                                             \
        if (_rc != BH_OK) return _rc;                                       \
    } while (0)

    // (re)allocate the per-body buffers; existing contents are NOT preserved
    int ensure_bodies(int64_t nn) {
        if (nn <= cap && scratch) return BH_OK;
        if (nn >= (int64_t)1 << 30) return fail(BH_E_ARG, "more than 2^30 bodies are not supported");
        const int64_t c = std::max<int64_t>(nn, std::max<int64_t>(1024, cap + cap / 4));
        const size_t cp = (size_t)c + PAD;
        free_bodies();
        BH_TRY(dev_alloc(&x, cp)); BH_TRY(dev_alloc(&y, cp)); BH_TRY(dev_alloc(&vx, cp)); BH_TRY(dev_alloc(&vy, cp));
        BH_TRY(dev_alloc(&m, cp)); BH_TRY(dev_alloc(&ax, cp)); BH_TRY(dev_alloc(&ay, cp)); BH_TRY(dev_alloc(&dtmp, cp));
        BH_TRY(dev_alloc(&perm, cp)); BH_TRY(dev_alloc(&origin, cp)); BH_TRY(dev_alloc(&leafpos, cp));
        BH_TRY(dev_alloc(&itmp, cp)); BH_TRY(dev_alloc(&inv, cp)); BH_TRY(dev_alloc(&iscr0, cp)); BH_TRY(dev_alloc(&iscr1, cp));
        BH_TRY(dev_alloc(&dead, cp)); BH_TRY(dev_alloc(&jflag, cp));
        if (cfg.flags & BH_FLAG_BODY_COUNTS) { BH_TRY(dev_alloc(&cntI, cp)); BH_TRY(dev_alloc(&cntO, cp)); }
        BH_TRY(dev_alloc(&keys_a, cp)); BH_TRY(dev_alloc(&keys_b, cp));
        BH_TRY(dev_alloc(&vals_a, cp)); BH_TRY(dev_alloc(&vals_b, cp));
        BH_TRY(dev_alloc(&S, cp));
        scratch_words = SC_WORDS + bhsort::sort_scratch_words(c, bhsort::MAX_PASSES) + scan_tiles(c) + 8;
        BH_TRY(dev_alloc(&scratch, scratch_words));
        cap = c;
        return BH_OK;
    }
    int ensure_cells(int64_t mm) {
        if (mm <= cell_cap) return BH_OK;
        if (mm >= (int64_t)1 << 31) return fail(BH_E_ARG, "tree has more than 2^31 cells");
        const int64_t c = std::max<int64_t>(mm + mm / 8, 2048);
        let.part_valid = false;      // peers map these arrays (let_map_peers): re-partition and re-map at the next re-homing
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
        if (!counts_on_host) { t.dev_n_in = &sc()->n_in; t.dev_n_int = &sc()->n_internal; }
        return t;
    }

    // ---- slices ------------------------------------------------------------------------
    int64_t slice_per() const { return (n + world - 1) / world; }
    void my_slice(int64_t* lo, int64_t* hi) const {
        if (world > 1 && let.part_valid && let.n_part == n) { *lo = let.cut[rank]; *hi = let.cut[rank + 1]; }   // cut at code boundaries
        else bh_slice_bounds(n, world, rank, lo, hi);
    }

    // ---- permutation helpers -----------------------------------------------------------
    // arr[h] <- arr[idx[h]] for the n body slots, through dtmp (pointer swap)
    int permute_d(double*& arr, const int* idx) {
        k_gather<double><<<grid_for(n, 256), 256, 0, st>>>(dtmp, arr, idx, (int)n);
        std::swap(arr, dtmp);
        ctr.kernel_launches += 1;
        return BH_OK;
    }
    // host (user order) <- device (home order)
    template <class T>
    int download_user(T* host, const T* dev_home, T* dev_tmp) {
        if (!host || n == 0) return BH_OK;
        const T* src = dev_home;
        if (!perm_identity) {
            k_scatter<T><<<grid_for(n, 256), 256, 0, st>>>(dev_tmp, dev_home, perm, (int)n);
            ctr.kernel_launches += 1;
            src = dev_tmp;
        }
        BH_TRY(cudaMemcpyAsync(host, src, (size_t)n * sizeof(T), cudaMemcpyDeviceToHost, st));
        return BH_OK;
    }
    // device (home order) <- host (user order)
    int upload_user(double* dev_home, const double* host) {
        if (n == 0) return BH_OK;
        if (perm_identity) {
            BH_TRY(cudaMemcpyAsync(dev_home, host, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
        } else {
            BH_TRY(cudaMemcpyAsync(dtmp, host, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, st));
            k_gather<double><<<grid_for(n, 256), 256, 0, st>>>(dev_home, dtmp, perm, (int)n);
            ctr.kernel_launches += 1;
        }
        return BH_OK;
    }

    // ---- communication -----------------------------------------------------------------
    int nccl_fail(int rc, const char* what) {
        const char* s = bhcomm::api().GetErrorString ? bhcomm::api().GetErrorString(rc) : "?";
        err = std::string(what) + ": " + s;
        return BH_E_NCCL;
    }
    // in-place all-gather of the rank slices of a and b (home order); slices are padded to
    // ceil(n/world) elements, the arrays have PAD slack behind n
    int all_gather_pair(double* a, double* b) {
        if (world <= 1) return BH_OK;
        if (transport != T_NCCL) return fail(BH_E_STATE, "host-staged transport: exchange the slices with bh_export_slice / bh_import_slices");
        bhcomm::Api& A = bhcomm::api();
        const size_t per = (size_t)slice_per();
        BH_TRY(cudaEventRecord(ev[12], st));
        int rc = A.GroupStart();
        if (let.part_valid && let.n_part == n) {   // slices of different lengths: one broadcast per rank
            for (int r = 0; r < world && rc == bhcomm::kSuccess; ++r) {
                const size_t cnt = (size_t)(let.cut[r + 1] - let.cut[r]);
                if (cnt == 0) continue;
                rc = A.Broadcast(a + let.cut[r], a + let.cut[r], cnt, bhcomm::kFloat64, r, comm, st);
                if (rc == bhcomm::kSuccess && b) rc = A.Broadcast(b + let.cut[r], b + let.cut[r], cnt, bhcomm::kFloat64, r, comm, st);
            }
        } else {
            if (rc == bhcomm::kSuccess) rc = A.AllGather(a + (size_t)rank * per, a, per, bhcomm::kFloat64, comm, st);
            if (rc == bhcomm::kSuccess && b) rc = A.AllGather(b + (size_t)rank * per, b, per, bhcomm::kFloat64, comm, st);
        }
        const int rc2 = A.GroupEnd();
        if (rc == bhcomm::kSuccess) rc = rc2;
        if (rc != bhcomm::kSuccess) return nccl_fail(rc, "ncclAllGather");
        BH_TRY(cudaEventRecord(ev[13], st));
        cur->has_comm = true;
        return BH_OK;
    }
    // every rank needs every body's velocity (re-homing, removal of merged bodies, read-back)
    int sync_velocities() {
        if (vel_valid || world <= 1) { vel_valid = true; return BH_OK; }
        if (transport != T_NCCL)
            return fail(BH_E_STATE, "velocities are not replicated: exchange BH_FIELD_VEL (bh_export_slice / bh_import_slices) first");
        BH_RC(all_gather_pair(vx, vy));
        vel_valid = true;
        return BH_OK;
    }

    // every rank needs every body's mass (replicated build, re-homing)
    int sync_masses() {
        if (mass_valid || world <= 1) { mass_valid = true; return BH_OK; }
        if (transport != T_NCCL) return fail(BH_E_STATE, "masses are not replicated");
        BH_RC(all_gather_pair(m, nullptr));
        mass_valid = true;
        heavies_valid = false;
        return BH_OK;
    }

    // ---- buildTree(), BH.kt:359-366 ------------------------------------------------------
    int build(int slot = 0, bool timed = true) {
        tree_valid = false;
        let.view_valid = false;
        root = BhRoot{par.root_cx, par.root_cy, par.root_half, bh_key_levels(par.root_half)};
        const int nn = (int)n;
        if (!let.local_build) {   // a replicated build needs every body's position and mass
            BH_RC(sync_positions());
            if (world > 1 && !mass_valid) BH_RC(wait_inputs());   // bh_step_io_slice: this slice's masses may still be in flight on the copy stream
            BH_RC(sync_masses());
        }
        if (rehome_due && nn > 0) { BH_RC(wait_inputs()); BH_RC(sync_velocities()); }   // (uploaded velocities first, then their exchange)
        if (timed && !capturing) BH_TRY(cudaEventRecord(ev[slot + 0], st));
        bool rehomed = false;
        // zero: scalars | sort scratch (sized for this n) | scan status
        const int key_bits = 2 * root.levels + 1;   // +1: the not-in-tree sentinel 1<<2L sorts last
        const int passes = (key_bits + bhsort::RADIX_BITS - 1) / bhsort::RADIX_BITS;
        const size_t sortw = bhsort::sort_scratch_words(nn, passes);
        uint32_t* scan_status = sort_scratch() + sortw;
        BH_TRY(cudaMemsetAsync(scratch, 0, (SC_WORDS + sortw + scan_tiles(nn)) * sizeof(uint32_t), st));
        n_in = 0; n_internal = 0; M = 0;
        if (nn > 0) {
            const uint64_t sentinel = 1ull << (2 * root.levels);
            if (let.local_build)
                k_keygen<<<std::min(grid_for(nn, 256), num_sms * 16), 256, 0, st>>>(x, y, nn, root, bh_make_grid(root), sentinel, keys_a, sc(),
                                                                                      let.ell, let.split.cs[rank], let.split.cs[rank + 1]);
            else
                k_keygen<<<std::min(grid_for(nn, 256), num_sms * 16), 256, 0, st>>>(x, y, nn, root, bh_make_grid(root), sentinel, keys_a, sc());
            const int where = sort_pairs(nn, key_bits);
            keys_sorted = where ? keys_b : keys_a;
            int* ord = reinterpret_cast<int*>(where ? vals_b : vals_a);
            order = ord;
            ctr.kernel_launches += 3 + passes;   // keygen, histogram, histogram_scan, passes
            if (rehome_due) {
                // the sorted order becomes the new home order; `order` becomes the identity
                permute_d(x, ord); permute_d(y, ord); permute_d(vx, ord); permute_d(vy, ord); permute_d(m, ord);
                k_gather<int><<<grid_for(nn, 256), 256, 0, st>>>(itmp, perm, ord, nn);
                std::swap(perm, itmp);
                k_iota<<<grid_for(nn, 256), 256, 0, st>>>(ord, nn);
                ctr.kernel_launches += 2;
                perm_identity = false;
                heavies_valid = false;
                rehome_due = false;
                steps_since_rehome = 0;
                ctr_rehomes++;
                rehomed = true;
                acc_valid = false;   // ax/ay are still in the OLD home order (BH_FLAG_REUSE_ACC must not kick with them)
            }
            k_count_scan<<<grid_for(nn, SCAN_TILE), SCAN_THREADS, 0, st>>>(keys_sorted, root.levels, sc(), S, scan_status);
            ctr.kernel_launches += 1;
        }
        if (syncfree_ok(nn)) {
            // SYNC-FREE build (small body counts, one GPU): the cell arrays hold the worst case, so nothing has to be read
            // back before the rest of the build is launched — n_in and the cell count stay on the device, every kernel
            // takes them from there (BhTreeView::dev_n_in), and the launch shapes cover all n bodies.  The jitter replay
            // is launched unconditionally (it returns at once when no two keys are equal).  No host round trip: the
            // whole step can be captured into one CUDA graph (run_steps).
            BH_RC(prepare_syncfree(nn, root.levels));      // (allocates only when n or the root box grew)
            counts_on_host = false;
            jitter_active = true;                          // unknown until finish(): handled as "maybe" (flags passed, no reuse)
            acc_valid = false;
            BH_TRY(cudaMemsetAsync(jflag, 0, (size_t)nn * sizeof(int), st));
            k_jitter<<<grid_for(nn, 128), 128, 0, st>>>(keys_sorted, const_cast<int*>(order), -1, root, perm, x, y, jflag, sc());
            BH_TRY(cudaMemsetAsync(leafpos, 0xFF, (size_t)nn * sizeof(int), st));   // -1: not in the tree
            BH_TRY(cudaMemsetAsync(arrived, 0, (size_t)cell_cap * sizeof(int), st));
            const BhTreeView t = view();
            k_emit<<<grid_for(nn, 256), 256, 0, st>>>(t, root.levels);
            BH_RC(wait_inputs());   // bh_step_io: the masses may still be in flight
            int* n_roots = dflags + HF_N_ROOTS;
            BH_TRY(cudaMemsetAsync(n_roots, 0, sizeof(int), st));
            k_climb_block<<<grid_for(nn, CLIMB_B), CLIMB_B, 0, st>>>(t, root, x, y, m, jflag, leafpos, climb_roots, n_roots);
            k_climb_top<<<std::min(grid_for(nn, 128 * 8), num_sms * 8), 128, 0, st>>>(t, root, climb_roots, n_roots);
            ctr.kernel_launches += 4;
            if (timed && !capturing) BH_TRY(cudaEventRecord(ev[slot + 1], st));
            BH_TRY(cudaGetLastError());
            tree_valid = true;
            return BH_OK;
        }
        counts_on_host = true;
        BH_TRY(cudaMemcpyAsync(sc_host, sc(), sizeof(DevScalars), cudaMemcpyDeviceToHost, st));
        BH_TRY(cudaStreamSynchronize(st));
        BH_TRY(cudaGetLastError());
        n_in = sc_host->n_in; n_internal = sc_host->n_internal; M = n_in + n_internal;
        ctr.n_in_tree = n_in; ctr.n_out_of_box = n - n_in; ctr.n_internal = n_internal; ctr.n_cells = M;
        ctr.n_jitter_bodies = sc_host->n_jitter; ctr.max_depth = sc_host->max_depth; ctr.key_levels = root.levels;
        if (!let.local_build) {      // (a local build of the domain mode sees this rank's bodies only)
            const bool any = sc_host->bb[0] != 0;
            ctr.bbox_max_x = any ? bh_ord_unkey(sc_host->bb[0]) : nan(""); ctr.bbox_min_x = any ? bh_ord_unkey(~sc_host->bb[1]) : nan("");
            ctr.bbox_max_y = any ? bh_ord_unkey(sc_host->bb[2]) : nan(""); ctr.bbox_min_y = any ? bh_ord_unkey(~sc_host->bb[3]) : nan("");
        }
        if (let.local_build && (int64_t)M + 1 > cell_cap) {
            // Domain mode: the peers map this rank's cell arrays, so a local build must never re-allocate them.
            // (A local tree has no more cells than the global tree the arrays were sized for; this can only
            // trigger if the tree grew by > 12 % since the last re-homing.)  Report it: the evaluation is
            // redone after a re-homing, whose replicated build grows the arrays on every rank alike.
            let.local_overflow = true;
            n_in = 0; n_internal = 0; M = 0;
            jitter_active = false;
            return BH_OK;
        }
        BH_RC(ensure_cells((int64_t)M + 1));   // + the terminal record the walk idles on
        if (rehomed && let_usable()) BH_RC(let_partition());   // slices of the new home order, cut at code boundaries
        jitter_active = false;
        if (sc_host->n_jitter > 0) {
            // jitter regime (BH.kt:145-156): replay each cluster of equal keys sequentially; this
            // MUTATES x/y of its bodies exactly like the reference's buildTree() does
            BH_TRY(cudaMemsetAsync(jflag, 0, (size_t)nn * sizeof(int), st));
            k_jitter<<<grid_for(n_in, 128), 128, 0, st>>>(keys_sorted, const_cast<int*>(order), n_in, root, perm, x, y, jflag, sc());
            ctr.kernel_launches += 1;
            jitter_active = true;
            acc_valid = false;       // positions were mutated: the accelerations on file belong to the old ones
        }
        if (nn > 0) BH_TRY(cudaMemsetAsync(leafpos, 0xFF, (size_t)nn * sizeof(int), st));   // -1: not in the tree
        if (n_in > 0) {
            BH_TRY(cudaMemsetAsync(arrived, 0, (size_t)M * sizeof(int), st));
            const BhTreeView t = view();
            k_emit<<<grid_for(n_in, 256), 256, 0, st>>>(t, root.levels);
            BH_RC(wait_inputs());   // bh_step_io: the masses may still be in flight
            {
                if (n_in > climb_roots_cap) {
                    dev_free(climb_roots);
                    climb_roots_cap = std::max<int64_t>(n_in + n_in / 8, 1024);
                    BH_TRY(dev_alloc(&climb_roots, (size_t)climb_roots_cap));
                }
                int* n_roots = dflags + HF_N_ROOTS;
                BH_TRY(cudaMemsetAsync(n_roots, 0, sizeof(int), st));
                k_climb_block<<<grid_for(n_in, CLIMB_B), CLIMB_B, 0, st>>>(t, root, x, y, m, jitter_active ? jflag : nullptr, leafpos,
                                                                             climb_roots, n_roots);
                k_climb_top<<<std::min(grid_for(n_in, 128 * 8), num_sms * 8), 128, 0, st>>>(t, root, climb_roots, n_roots);
                ctr.kernel_launches += 3;
            }
        }
        if (timed && !capturing) BH_TRY(cudaEventRecord(ev[slot + 1], st));
        BH_TRY(cudaGetLastError());
        tree_valid = true;
        return BH_OK;
    }
    int64_t ctr_rehomes = 0, ctr_reused = 0;
    int io_steps_left = 0;
    bool syncfree_ok(int64_t nn) const {
        return syncfree_enabled && world == 1 && !let.local_build && nn > 0 && nn <= SYNCFREE_MAX_N && !(cfg.flags & BH_FLAG_REUSE_ACC);
    }
    // worst case of the cell count: every body adds at most `levels` internal cells on its path
    int prepare_syncfree(int64_t nn, int levels) {
        BH_RC(ensure_cells(nn * (levels + 1) + 2));
        if (nn > climb_roots_cap) {
            dev_free(climb_roots);
            climb_roots_cap = std::max<int64_t>(nn + nn / 8, 1024);
            BH_TRY(dev_alloc(&climb_roots, (size_t)climb_roots_cap));
        }
        return BH_OK;
    }
    int walk_g = 0;                 // bodies per thread of the walk; 0 = automatic (BH_WALK_G=1|2 pins it)
    bool walk_affine = true;        // BH_WALK_AFFINE=0: every chunk from the global queue (no SM affinity)
    unsigned int* walk_queue = nullptr;   // work counters of the persistent walk kernel
    int walk_acc = -1;              // BH_WALK_ACC=1: f64 summation (BH_ACC_F64) also for one body per lane

    int sort_pairs(int nn, int key_bits) {
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

    // computeAccelerations(root), BH.kt:374-395, for the home slots [first, first+count)
    int walk(int64_t first, int64_t count, int slot = 0, const BhTreeView* over = nullptr) {
        const BhTreeView tv = over ? *over : view();
        if (!capturing) BH_TRY(cudaEventRecord(ev[slot + 2], st));
        if (count > 0) {
            const BhWalkParams w = bh_walk_params(par.theta, par.soft2, par.root_half);
            // Bodies per thread (bh_walk_multi): pairs sharing every record load, each term added to an f64 sum at once
            // (BH_ACC_F64: a body's result does not depend on its partner), when the LIST is long enough to fill the
            // machine with half the threads; else one body per lane folding FP32 partial sums every BH_WALK_CHUNK of
            // its own visits (BH_ACC_FOLD).  The choice is made from the length of the whole list, n — the same number
            // on every rank and for every world size — never from this call's `count`: all ranks and all partitions
            // of the targets then run the same arithmetic and reproduce one GPU bit for bit.  BH_WALK_G pins the shape,
            // BH_WALK_ACC=1 selects the f64 summation for one body per lane as well (tests, measurements).
            int g = walk_g;
            if (g == 0) g = n >= (int64_t)num_sms * 128 * 16 ? 2 : 1;
            g = g >= 2 ? 2 : 1;
            const bool f64acc = g > 1 || walk_acc == BH_ACC_F64;
            // persistent SM-affine schedule (see k_walk): one range of chunks per SM, stealing between ranges
            BhWalkQueue q;
            q.next = walk_queue;
            q.n_ranges = num_sms;
            q.chunks = (int)((count + 32 * g - 1) / (32 * g));
            q.per = (q.chunks + q.n_ranges - 1) / q.n_ranges;
            if (!walk_affine) { q.n_ranges = 1; q.per = q.chunks; }     // one queue: no SM affinity
            BH_TRY(cudaMemsetAsync(walk_queue, 0, (size_t)q.n_ranges * sizeof(unsigned int), st));
#define BH_LAUNCH_WALK(GG, AA, MINB)                                                                                        \
    k_walk<GG, AA, MINB><<<(int)std::min<int64_t>((q.chunks + 3) / 4, (int64_t)num_sms * MINB), 128, 0, st>>>(              \
        tv, w, (int)first, (int)count, x, y, m, leafpos, par.G, ax, ay, cntI, cntO, sc(), tot, q)
            if (g == 2) BH_LAUNCH_WALK(2, BH_ACC_F64, 7);
            else if (f64acc) BH_LAUNCH_WALK(1, BH_ACC_F64, 9);
            else BH_LAUNCH_WALK(1, BH_ACC_FOLD, 9);
#undef BH_LAUNCH_WALK
            ctr.kernel_launches += 1;
        }
        if (!capturing) BH_TRY(cudaEventRecord(ev[slot + 3], st));
        BH_TRY(cudaGetLastError());
        return BH_OK;
    }

    int evaluate(int slot, int64_t lo, int64_t hi) {
        BH_RC(build(slot));
        BH_RC(walk(lo, hi - lo, slot));
        ctr.total_evaluations++;
        return BH_OK;
    }

    // after a stream sync: fold event timings and device counters into ctr
    int finish() {
        BH_TRY(cudaMemcpyAsync(sc_host, sc(), sizeof(DevScalars), cudaMemcpyDeviceToHost, st));
        BH_TRY(cudaMemcpyAsync(tot_host, tot, sizeof(DevTotals), cudaMemcpyDeviceToHost, st));
        BH_TRY(cudaStreamSynchronize(st));
        BH_TRY(cudaGetLastError());
        if (!counts_on_host) {       // the last build was sync-free: its counts arrive only now
            n_in = sc_host->n_in; n_internal = sc_host->n_internal; M = n_in + n_internal;
            ctr.n_in_tree = n_in; ctr.n_out_of_box = n - n_in; ctr.n_internal = n_internal; ctr.n_cells = M;
            ctr.n_jitter_bodies = sc_host->n_jitter; ctr.max_depth = sc_host->max_depth; ctr.key_levels = root.levels;
            const bool any = sc_host->bb[0] != 0;
            ctr.bbox_max_x = any ? bh_ord_unkey(sc_host->bb[0]) : nan(""); ctr.bbox_min_x = any ? bh_ord_unkey(~sc_host->bb[1]) : nan("");
            ctr.bbox_max_y = any ? bh_ord_unkey(sc_host->bb[2]) : nan(""); ctr.bbox_min_y = any ? bh_ord_unkey(~sc_host->bb[3]) : nan("");
            jitter_active = sc_host->n_jitter > 0;
            counts_on_host = true;
        }
        ctr.interactions = (int64_t)sc_host->interactions;
        ctr.opened = (int64_t)sc_host->opened;
        ctr.exact_retests = (int64_t)sc_host->retests;
        if (jitter_active) {
            ctr.n_in_tree = n_in - sc_host->n_ghost;
            if (sc_host->jitter_unsupported) return fail(BH_E_UNSUPPORTED, "jitter regime: a body survived below depth levels+1");
        }
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

    int kick(int64_t lo, int64_t hi, double dtHalf, double dt, int drift) {
        if (hi > lo) {
            k_kick_drift<<<grid_for(hi - lo, 256), 256, 0, st>>>((int)lo, (int)hi, x, y, vx, vy, ax, ay, dtHalf, dt, drift);
            ctr.kernel_launches += 1;
        }
        BH_TRY(cudaGetLastError());
        return BH_OK;
    }

    // ---- PhysicsEngine.step(), BH.kt:405-439, in three phases ------------------------------
    int step_begin() {   // :407-422  a(t), half kick, drift — own slice
        if (phase != 0) return fail(BH_E_STATE, "bh_step_begin: a step is already in progress");
        const double dt = par.dt, dtHalf = par.dt * 0.5;   // BH.kt:412
        int64_t lo, hi;
        my_slice(&lo, &hi);
        const bool reuse = (cfg.flags & BH_FLAG_REUSE_ACC) && acc_valid && !rehome_due && acc_par.theta == par.theta &&
                           acc_par.G == par.G && acc_par.soft2 == par.soft2 && acc_par.root_cx == par.root_cx &&
                           acc_par.root_cy == par.root_cy && acc_par.root_half == par.root_half;
        if (reuse) {   // a(t) is the a(t+dt) the previous step ended with: same positions, masses and parameters
            BH_TRY(cudaEventRecord(ev[0], st)); BH_TRY(cudaEventRecord(ev[1], st));
            BH_TRY(cudaEventRecord(ev[2], st)); BH_TRY(cudaEventRecord(ev[3], st));
            ctr_reused++;
        } else {
            BH_RC(evaluate_slice(0));
            my_slice(&lo, &hi);          // a re-homing build may have re-cut the slices
        }
        acc_valid = false;
        BH_RC(wait_inputs());
        BH_RC(kick(lo, hi, dtHalf, dt, 1));
        if (io_out.armed && io_steps_left == 1) BH_RC(emit_positions_out());
        if (world > 1) {
            vel_valid = false;
            if (transport == T_NCCL) let.pos_valid = false;   // only this rank's slice has drifted here
        }
        phase = 1;
        return BH_OK;
    }
    int step_end() {     // :425-435  a(t+dt), half kick — own slice
        if (phase != 1) return fail(BH_E_STATE, "bh_step_end: call bh_step_begin (and exchange BH_FIELD_POS) first");
        const double dt = par.dt, dtHalf = par.dt * 0.5;
        int64_t lo, hi;
        BH_RC(evaluate_slice(4));
        my_slice(&lo, &hi);
        BH_RC(kick(lo, hi, dtHalf, dt, 0));
        if (world > 1) vel_valid = false;   // (a domain-mode fallback may have re-homed, i.e. replicated the velocities, inside this evaluation)
        acc_valid = !jitter_active;      // a jittering build mutated positions: the next build will again
        acc_par = par;
        phase = 2;
        return BH_OK;
    }
    int step_finish() {  // :438  merge rule
        if (phase != 2) return fail(BH_E_STATE, "bh_step_finish: call bh_step_end first");
        phase = 0;
        ctr.total_steps++;
        if (io_steps_left > 0) --io_steps_left;
        if (++steps_since_rehome >= rehome_interval) rehome_due = true;
        return merge_rule();
    }
    // one whole step with the engine's own transport
    int step_once() {
        BH_RC(step_begin());
        // the one exchange of the step: drifted positions — not in domain mode, where every rank keeps
        // only its own slice current and the next build exchanges strays and tree blocks instead
        if (world > 1 && !let_usable()) BH_RC(sync_positions());
        BH_RC(step_end());
        return step_finish();
    }

    bool merge_enabled() const { return par.merge_min_dist > 0.0 && n > 1; }
    int excl_scan(const int* in, int nn, int* out);
    int merge_rule();
    int merge_done();
    int grow_keep(int64_t extra);
    int finish_append(int64_t added);
    int run_steps(int nsteps);
    void collect(EvSlot& sl);
};

#include "bh_merge.cuh"
#include "bh_scene.cuh"
#include "bh_let_engine.cuh"

// one force evaluation for this rank's slice: over the locally essential tree when the domain mode is
// active, else over the (replicated) tree of all bodies
int bh_engine::evaluate_slice(int slot) {
    if (let_usable()) {
        if (!let_ready()) {
            if (let.n_declined != n) rehome_due = true;       // (re)partition through a re-homing build
        } else {
            const int rc = let_evaluate(slot);
            if (rc == BH_OK) {
                BhTreeView lv{};
                lv.cell = let.cell; lv.cd = let.cd; lv.sk = let.sk; lv.arrived = let.arrived; lv.n_in = let.n_items; lv.M = let.M;
                BH_RC(walk(let.cut[rank], let.cut[rank + 1] - let.cut[rank], slot, &lv));
                ctr.total_evaluations++;
                return BH_OK;
            }
            if (rc != BH_LET_RETRY) return rc;
            let.fallbacks++;
            rehome_due = true;
        }
    }
    BH_RC(build(slot));
    int64_t lo, hi;
    my_slice(&lo, &hi);
    BH_RC(walk(lo, hi - lo, slot));
    ctr.total_evaluations++;
    return BH_OK;
}

// bh_step_io: the positions are final after the last drift — send (x, y, m) to the host on the copy
// stream while the last force evaluation runs on the compute stream
int bh_engine::emit_positions_out() {
    io_out.armed = false;
    BH_TRY(cudaEventRecord(io_ev[1], st));
    BH_TRY(cudaStreamWaitEvent(copy_st, io_ev[1], 0));
    if (io_out.slice) {                      // bh_step_io_slice: this rank's slice as it lies in the home order
        int64_t lo, hi;
        my_slice(&lo, &hi);
        io_out.epoch = ctr_rehomes;          // (a re-homing after this point re-orders the slice: the caller copies again)
        const size_t bytes = (size_t)(hi - lo) * sizeof(double);
        if (bytes) {
            if (io_out.x) BH_TRY(cudaMemcpyAsync(io_out.x, x + lo, bytes, cudaMemcpyDeviceToHost, copy_st));
            if (io_out.y) BH_TRY(cudaMemcpyAsync(io_out.y, y + lo, bytes, cudaMemcpyDeviceToHost, copy_st));
            if (io_out.m) BH_TRY(cudaMemcpyAsync(io_out.m, m + lo, bytes, cudaMemcpyDeviceToHost, copy_st));
        }
        return BH_OK;
    }
    const double* src[3] = {x, y, m};
    double* dst[3] = {io_out.x, io_out.y, io_out.m};
    for (int k = 0; k < 3; ++k) {
        if (!dst[k]) continue;
        const double* from = src[k];
        if (!perm_identity) {
            k_scatter<double><<<grid_for(n, 256), 256, 0, copy_st>>>(io_stage[k], src[k], perm, (int)n);
            ctr.kernel_launches += 1;
            from = io_stage[k];
        }
        BH_TRY(cudaMemcpyAsync(dst[k], from, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, copy_st));
    }
    return BH_OK;
}

// fold the timers of a finished step slot into the counters (waits for that step only)
void bh_engine::collect(EvSlot& sl) {
    if (!sl.used) return;
    cudaEventSynchronize(sl.e[9]);
    cudaEvent_t* keep = ev;
    ev = sl.e;
    const float phases = sl.phases ? add_phase_times(0) + add_phase_times(4) : 0.f;   // (a step replayed from a graph has no phase events)
    float total = 0.f, c = 0.f, mg = 0.f;
    // kick/drift (+ exchange, merge) = whole step minus the build and walk phases
    if (cudaEventElapsedTime(&total, ev[8], ev[9]) == cudaSuccess && total > phases) ctr.ms_integrate += total - phases;
    if (sl.has_comm && cudaEventElapsedTime(&c, ev[12], ev[13]) == cudaSuccess) ctr.ms_comm += c;
    if (sl.has_merge && cudaEventElapsedTime(&mg, ev[14], ev[15]) == cudaSuccess) ctr.ms_merge += mg;
    sl.used = sl.has_comm = sl.has_merge = false;
    sl.phases = true;
    ev = keep;
}

int bh_engine::run_steps(int nsteps) {
    BH_TRY(cudaEventRecord(call_ev[0], st));
    int rc = BH_OK;
    for (int s = 0; s < nsteps && rc == BH_OK; ++s) {
        EvSlot& sl = ring[s % EV_RING];
        collect(sl);
        cur = &sl; ev = sl.e;
        sl.used = true;
        BH_TRY(cudaEventRecord(ev[8], st));
        if (graph_enabled && syncfree_ok(n) && !io_out.armed && !io_wait_in) {
            // One CUDA graph per step: the launches of the two evaluations and the kicks (45 small kernels and memsets at
            // the reference's 12,500 bodies, where the step is launch-latency bound) are CAPTURED instead of issued —
            // the host-side bookkeeping of step_begin / step_end runs as always — and the graph of the previous step is
            // updated in place with this step's arguments (same topology: a cheap parameter patch; a re-homing step has
            // another topology and instantiates anew).  The merge rule reads counts back, so it runs behind the graph.
            BH_RC(prepare_syncfree(n, bh_key_levels(par.root_half)));   // every allocation before the capture
            capturing = true;
            cudaError_t ce = cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
            if (ce == cudaSuccess) {
                rc = step_begin();
                if (rc == BH_OK) rc = step_end();
                cudaGraph_t g = nullptr;
                ce = cudaStreamEndCapture(st, &g);
                capturing = false;
                if (rc == BH_OK && ce == cudaSuccess) {
                    if (gexec) {
                        cudaGraphExecUpdateResultInfo info;
                        if (cudaGraphExecUpdate(gexec, g, &info) != cudaSuccess) { cudaGraphExecDestroy(gexec); gexec = nullptr; cudaGetLastError(); }
                    }
                    if (!gexec) { ce = cudaGraphInstantiate(&gexec, g, 0); ctr_graph_instantiations++; }
                    if (ce == cudaSuccess) ce = cudaGraphLaunch(gexec, st);
                }
                if (g) cudaGraphDestroy(g);
            }
            capturing = false;
            if (rc == BH_OK && ce != cudaSuccess) rc = cuda_fail(ce, "CUDA graph of a step");
            sl.phases = false;
            ctr_graph_steps++;
            if (rc == BH_OK) rc = step_finish();
        } else
            rc = step_once();
        if (rc != BH_OK) { phase = 0; sl.used = false; break; }
        BH_TRY(cudaEventRecord(ev[9], st));
    }
    for (auto& sl : ring) collect(sl);
    cur = &ring[0]; ev = ring[0].e;
    if (rc != BH_OK) return rc;
    BH_TRY(cudaEventRecord(call_ev[1], st));
    BH_RC(finish());
    float call_ms = 0.f;
    if (cudaEventElapsedTime(&call_ms, call_ev[0], call_ev[1]) == cudaSuccess) ctr.ms_step_call = call_ms;
    return BH_OK;
}

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
    e->rehome_interval = e->cfg.rehome_interval > 0 ? e->cfg.rehome_interval : 8;
    if (const char* s = getenv("BH_REHOME_INTERVAL")) { const int v = atoi(s); if (v > 0) e->rehome_interval = v; }
    if (const char* s = getenv("BH_WALK_G")) e->walk_g = atoi(s);
    if (const char* s = getenv("BH_WALK_ACC")) e->walk_acc = atoi(s);
    if (const char* s = getenv("BH_WALK_AFFINE")) e->walk_affine = atoi(s) != 0;
    if (const char* s = getenv("BH_SYNCFREE")) e->syncfree_enabled = atoi(s) != 0;
    if (const char* s = getenv("BH_GRAPH")) e->graph_enabled = atoi(s) != 0;
    cudaError_t ce = cudaSetDevice(e->device);
    if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&e->st, cudaStreamNonBlocking);
    for (auto& sl : e->ring) for (int k = 0; k < 16 && ce == cudaSuccess; ++k) ce = cudaEventCreate(&sl.e[k]);
    for (int k = 0; k < 2 && ce == cudaSuccess; ++k) ce = cudaEventCreate(&e->call_ev[k]);
    if (ce == cudaSuccess) ce = cudaMallocHost((void**)&e->sc_host, sizeof(DevScalars));
    if (ce == cudaSuccess) ce = cudaMallocHost((void**)&e->tot_host, sizeof(DevTotals));
    if (ce == cudaSuccess) ce = cudaMallocHost((void**)&e->hflags, HF_COUNT * sizeof(int));
    if (ce == cudaSuccess) ce = dev_alloc(&e->dflags, HF_COUNT);
    if (ce == cudaSuccess) ce = dev_alloc(&e->tot, 1);
    if (ce == cudaSuccess) ce = cudaMemset(e->tot, 0, sizeof(DevTotals));
    if (ce == cudaSuccess) ce = dev_alloc(&e->red, 4);
    if (ce == cudaSuccess) ce = dev_alloc(&e->walk_queue, 1024);
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
    if (e->copy_st) { cudaStreamSynchronize(e->copy_st); cudaStreamDestroy(e->copy_st); }
    for (auto& ev : e->snap_ev) if (ev) cudaEventDestroy(ev);
    for (auto& q : e->io_stage) dev_free(q);
    for (auto& ev : e->io_ev) if (ev) cudaEventDestroy(ev);
    dev_free(e->snap_xy); dev_free(e->snap_m);
    if (e->snap_hxy) cudaFreeHost(e->snap_hxy);
    if (e->snap_hm) cudaFreeHost(e->snap_hm);
    if (e->gexec) cudaGraphExecDestroy(e->gexec);
    if (e->comm && bhcomm::api().ok) bhcomm::api().CommDestroy(e->comm);
    e->free_bodies();
    e->free_cells();
    e->let.release();
    dev_free(e->tot); dev_free(e->red); dev_free(e->walk_queue); dev_free(e->dflags); dev_free(e->heavy);
    if (e->sc_host) cudaFreeHost(e->sc_host);
    if (e->tot_host) cudaFreeHost(e->tot_host);
    if (e->hflags) cudaFreeHost(e->hflags);
    for (auto& sl : e->ring) for (auto& ev : sl.e) if (ev) cudaEventDestroy(ev);
    for (auto& ev : e->call_ev) if (ev) cudaEventDestroy(ev);
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
#define E_RC(expr)                                                          \
    do {                                                                    \
        const int _rc = (expr);                                             \
        if (_rc != BH_OK) return _rc;                                       \
    } while (0)

// resetBodies(newBodies), BH.kt:342-349.  A list of the same length as the previous one is
// stored through the existing home permutation (any permutation is valid; it only affects
// memory coalescing); otherwise home order starts as user order and the first build re-homes.
int bh_set_bodies(bh_engine* e, int64_t n, const double* x, const double* y, const double* vx, const double* vy,
                  const double* m) {
    if (!e || n < 0 || (n > 0 && (!x || !y || !vx || !vy || !m))) return e ? e->fail(BH_E_ARG, "bh_set_bodies: bad arguments") : BH_E_ARG;
    E_TRY(cudaSetDevice(e->device));
    if (e->phase != 0) return e->fail(BH_E_STATE, "bh_set_bodies: a step is in progress");
    const bool keep_perm = (n == e->n && n > 0 && !e->perm_identity && n <= e->cap);
    E_RC(e->ensure_bodies(n));
    e->n = n;
    if (!keep_perm) {
        e->perm_identity = true;
        e->rehome_due = true;
        if (n > 0) k_iota<<<grid_for(n, 256), 256, 0, e->st>>>(e->perm, (int)n);
    }
    const size_t bytes = (size_t)n * sizeof(double);
    if (n > 0) {
        E_RC(e->upload_user(e->x, x)); E_RC(e->upload_user(e->y, y));
        E_RC(e->upload_user(e->vx, vx)); E_RC(e->upload_user(e->vy, vy)); E_RC(e->upload_user(e->m, m));
        E_TRY(cudaMemsetAsync(e->ax, 0, bytes, e->st));
        E_TRY(cudaMemsetAsync(e->ay, 0, bytes, e->st));
        E_TRY(cudaMemsetAsync(e->dflags, 0, HF_COUNT * sizeof(int), e->st));
    }
    E_TRY(cudaStreamSynchronize(e->st));
    E_TRY(cudaGetLastError());
    e->origin_identity = true;
    e->tree_valid = false;
    e->heavies_valid = false;
    e->vel_valid = true;
    e->mass_valid = true;
    e->acc_valid = false;
    e->let.pos_valid = true;
    e->let.part_valid = false; e->let.n_declined = -1;
    if (e->let.enabled) e->rehome_due = true;
    return BH_OK;
}

int64_t bh_num_bodies(const bh_engine* e) { return e ? e->n : 0; }

int bh_get_bodies(bh_engine* e, int64_t cap, double* x, double* y, double* vx, double* vy, double* m, int64_t* n_out) {
    if (!e) return BH_E_ARG;
    if (n_out) *n_out = e->n;
    if (cap < e->n) return e->fail(BH_E_ARG, "bh_get_bodies: capacity too small");
    E_TRY(cudaSetDevice(e->device));
    if (e->n > 0) {
        E_RC(e->sync_positions());
        if (vx || vy) E_RC(e->sync_velocities());
        E_RC(e->sync_masses());             // (stale only after bh_step_io_slice brought new masses for the slices)
        // one staging buffer: each scatter is stream-ordered behind the previous copy
        E_RC(e->download_user(x, e->x, e->dtmp)); E_RC(e->download_user(y, e->y, e->dtmp));
        E_RC(e->download_user(vx, e->vx, e->dtmp)); E_RC(e->download_user(vy, e->vy, e->dtmp));
        E_RC(e->download_user(m, e->m, e->dtmp));
    }
    E_TRY(cudaStreamSynchronize(e->st));
    return BH_OK;
}

int bh_get_origin(bh_engine* e, int64_t cap, int32_t* origin, int64_t* n_out) {
    if (!e) return BH_E_ARG;
    if (n_out) *n_out = e->n;
    if (cap < e->n) return e->fail(BH_E_ARG, "bh_get_origin: capacity too small");
    if (!origin || e->n == 0) return BH_OK;
    if (e->origin_identity) { for (int64_t i = 0; i < e->n; ++i) origin[i] = (int32_t)i; return BH_OK; }
    E_TRY(cudaSetDevice(e->device));
    E_TRY(cudaMemcpyAsync(origin, e->origin, (size_t)e->n * sizeof(int32_t), cudaMemcpyDeviceToHost, e->st));
    E_TRY(cudaStreamSynchronize(e->st));
    return BH_OK;
}

int bh_rebase_origin(bh_engine* e) {
    if (!e) return BH_E_ARG;
    if (e->phase != 0) return e->fail(BH_E_STATE, "bh_rebase_origin: a step is in progress");
    e->origin_identity = true;     // the merge rule re-initialises origin[] from the identity on demand
    return BH_OK;
}

int bh_get_positions_f32(bh_engine* e, int64_t cap, float* xy, float* m, int64_t* n_out) {
    if (!e) return BH_E_ARG;
    if (n_out) *n_out = e->n;
    if (cap < e->n) return e->fail(BH_E_ARG, "bh_get_positions_f32: capacity too small");
    if (e->n == 0) return BH_OK;
    E_TRY(cudaSetDevice(e->device));
    E_RC(e->sync_positions());
    E_RC(e->sync_masses());
    // staged in user order through dtmp (float2[n]) and itmp (float[n])
    float2* dxy = reinterpret_cast<float2*>(e->dtmp);
    float* dm = reinterpret_cast<float*>(e->itmp);
    k_positions_f32<<<grid_for(e->n, 256), 256, 0, e->st>>>(e->x, e->y, e->m, e->perm, (int)e->n, dxy, dm);
    e->ctr.kernel_launches += 1;
    if (xy) E_TRY(cudaMemcpyAsync(xy, dxy, (size_t)e->n * sizeof(float2), cudaMemcpyDeviceToHost, e->st));
    if (m) E_TRY(cudaMemcpyAsync(m, dm, (size_t)e->n * sizeof(float), cudaMemcpyDeviceToHost, e->st));
    E_TRY(cudaStreamSynchronize(e->st));
    return BH_OK;
}

// ---- scene generators on the device (BodyFactory.kt) -------------------------------------------
int bh_default_disk_params(int32_t w, int32_t h, bh_disk_params* p) {
    if (!p) return BH_E_ARG;
    memset(p, 0, sizeof(*p));
    p->x = w * 0.5; p->y = h * 0.5;                      // BodyFactory.kt:75-76
    p->r = 200.0; p->min_r = 8.0;                        // :77-78, Config.kt:35
    p->central_mass = 50000.0; p->total_satellite_mass = 5000.0;   // Config.kt:32,38
    p->eps_m2 = 0.03; p->phi0 = 0.0; p->bar_taper_r = 0.0; p->radial_scale = 0.0;   // :66-70
    p->speed_jitter = 0.01; p->radial_jitter = 0.0; p->clockwise = 1; p->kepler = 0;   // :71-73
    return BH_OK;
}

int bh_append_disk(bh_engine* e, int64_t n_total, const bh_disk_params* p, uint64_t seed) {
    if (!e || !p) return e ? e->fail(BH_E_ARG, "bh_append_disk: bad arguments") : BH_E_ARG;
    if (e->phase != 0) return e->fail(BH_E_STATE, "bh_append_disk: a step is in progress");
    E_TRY(cudaSetDevice(e->device));
    const int64_t add = std::max<int64_t>(n_total, 1);   // sats = (nTotal - 1).coerceAtLeast(0), plus the centre
    if (e->n + add >= (int64_t)1 << 30) return e->fail(BH_E_ARG, "more than 2^30 bodies are not supported");
    E_RC(e->sync_positions());
    E_RC(e->sync_masses());
    E_RC(e->sync_velocities());
    E_RC(e->grow_keep(add));
    const int64_t b = e->n;
    const int na = (int)add;
    // distances from the centre -> keys_a; sorted by radius with the engine's onesweep (64-bit keys)
    k_gen_disk_positions<<<grid_for(na, 256), 256, 0, e->st>>>(na, *p, seed, e->x + b, e->y + b, e->vx + b, e->vy + b, e->m + b, e->keys_a);
    if (na > 1) {
        const int where = bhsort::onesweep_sort(e->keys_a, e->vals_a, e->keys_b, e->vals_b, na, 64, e->sort_scratch(), e->st, e->num_sms);
        const uint32_t* by_radius = where ? e->vals_b : e->vals_a;
        k_gen_disk_velocities<<<grid_for(na, 256), 256, 0, e->st>>>(na, *p, e->par.G, seed, by_radius, e->x + b, e->y + b, e->vx + b, e->vy + b);
        e->ctr.kernel_launches += 12;
    }
    E_TRY(cudaGetLastError());
    return e->finish_append(add);
}

int bh_append_uniform_random(bh_engine* e, int64_t n, double m, int32_t w, int32_t h, uint64_t seed) {
    if (!e) return BH_E_ARG;
    if (e->phase != 0) return e->fail(BH_E_STATE, "bh_append_uniform_random: a step is in progress");
    if (n <= 0 || !(m > 0.0)) return BH_OK;              // BodyFactory.kt:165: empty list
    E_TRY(cudaSetDevice(e->device));
    if (e->n + n >= (int64_t)1 << 30) return e->fail(BH_E_ARG, "more than 2^30 bodies are not supported");
    E_RC(e->sync_positions());
    E_RC(e->sync_masses());
    E_RC(e->sync_velocities());
    E_RC(e->grow_keep(n));
    const int64_t b = e->n;
    k_gen_uniform<<<grid_for(n, 256), 256, 0, e->st>>>((int)n, (double)w, (double)h, m, seed, e->x + b, e->y + b, e->vx + b, e->vy + b, e->m + b);
    e->ctr.kernel_launches += 1;
    E_TRY(cudaGetLastError());
    return e->finish_append(n);
}

int bh_request_positions_f32(bh_engine* e) {
    if (!e) return BH_E_ARG;
    E_TRY(cudaSetDevice(e->device));
    if (!e->copy_st) {
        E_TRY(cudaStreamCreateWithFlags(&e->copy_st, cudaStreamNonBlocking));
        E_TRY(cudaEventCreateWithFlags(&e->snap_ev[0], cudaEventDisableTiming));
        E_TRY(cudaEventCreateWithFlags(&e->snap_ev[1], cudaEventDisableTiming));
    }
    if (e->n > e->snap_cap) {
        E_TRY(cudaStreamSynchronize(e->copy_st));
        dev_free(e->snap_xy); dev_free(e->snap_m);
        if (e->snap_hxy) cudaFreeHost(e->snap_hxy);
        if (e->snap_hm) cudaFreeHost(e->snap_hm);
        e->snap_hxy = e->snap_hm = nullptr;
        const int64_t c = std::max<int64_t>(e->n + e->n / 8, 1024);
        E_TRY(dev_alloc(&e->snap_xy, (size_t)c)); E_TRY(dev_alloc(&e->snap_m, (size_t)c));
        E_TRY(cudaMallocHost((void**)&e->snap_hxy, (size_t)c * sizeof(float2)));
        E_TRY(cudaMallocHost((void**)&e->snap_hm, (size_t)c * sizeof(float)));
        e->snap_cap = c;
    }
    if (e->snap_n >= 0) E_TRY(cudaStreamWaitEvent(e->st, e->snap_ev[1], 0));   // previous copy still reads the staging
    e->snap_n = e->n;
    E_RC(e->sync_positions());
    E_RC(e->sync_masses());
    if (e->n > 0) {
        k_positions_f32<<<grid_for(e->n, 256), 256, 0, e->st>>>(e->x, e->y, e->m, e->perm, (int)e->n, e->snap_xy, e->snap_m);
        e->ctr.kernel_launches += 1;
    }
    E_TRY(cudaEventRecord(e->snap_ev[0], e->st));
    E_TRY(cudaStreamWaitEvent(e->copy_st, e->snap_ev[0], 0));
    if (e->n > 0) {
        E_TRY(cudaMemcpyAsync(e->snap_hxy, e->snap_xy, (size_t)e->n * sizeof(float2), cudaMemcpyDeviceToHost, e->copy_st));
        E_TRY(cudaMemcpyAsync(e->snap_hm, e->snap_m, (size_t)e->n * sizeof(float), cudaMemcpyDeviceToHost, e->copy_st));
    }
    E_TRY(cudaEventRecord(e->snap_ev[1], e->copy_st));
    return BH_OK;
}

int bh_wait_positions_f32(bh_engine* e, const float** xy, const float** m, int64_t* n) {
    if (!e) return BH_E_ARG;
    if (e->snap_n < 0) return e->fail(BH_E_STATE, "bh_wait_positions_f32: no snapshot was requested");
    E_TRY(cudaSetDevice(e->device));
    E_TRY(cudaEventSynchronize(e->snap_ev[1]));
    if (xy) *xy = e->snap_hxy;
    if (m) *m = e->snap_hm;
    if (n) *n = e->snap_n;
    return BH_OK;
}

int bh_step(bh_engine* e, int32_t nsteps) {
    if (!e || nsteps < 0) return e ? e->fail(BH_E_ARG, "bh_step: bad arguments") : BH_E_ARG;
    E_TRY(cudaSetDevice(e->device));
    if (e->world > 1 && e->transport != bh_engine::T_NCCL)
        return e->fail(BH_E_STATE, "bh_step: host-staged transport — drive the step with bh_step_begin / bh_step_end / bh_step_finish");
    return e->run_steps(nsteps);
}

int bh_step_begin(bh_engine* e) {
    if (!e) return BH_E_ARG;
    E_TRY(cudaSetDevice(e->device));
    E_RC(e->step_begin());
    E_TRY(cudaStreamSynchronize(e->st));
    return BH_OK;
}
int bh_step_end(bh_engine* e) {
    if (!e) return BH_E_ARG;
    E_TRY(cudaSetDevice(e->device));
    E_RC(e->step_end());
    E_RC(e->finish());
    e->add_phase_times(0); e->add_phase_times(4);
    return BH_OK;
}
int bh_step_finish(bh_engine* e) {
    if (!e) return BH_E_ARG;
    E_TRY(cudaSetDevice(e->device));
    E_RC(e->step_finish());
    E_TRY(cudaStreamSynchronize(e->st));
    float mg = 0.f;
    if (e->cur->has_merge && cudaEventElapsedTime(&mg, e->ev[14], e->ev[15]) == cudaSuccess) e->ctr.ms_merge += mg;
    e->cur->has_merge = false;
    return BH_OK;
}

int bh_step_io(bh_engine* e, int32_t nsteps, int64_t n_in, const double* x_in, const double* y_in, const double* vx_in,
               const double* vy_in, const double* m_in, int64_t cap_out, double* x_out, double* y_out, double* vx_out,
               double* vy_out, double* m_out, int64_t* n_out) {
    if (!e || nsteps < 0 || (x_in && (n_in < 0 || (n_in > 0 && (!y_in || !vx_in || !vy_in || !m_in)))))
        return e ? e->fail(BH_E_ARG, "bh_step_io: bad arguments") : BH_E_ARG;
    E_TRY(cudaSetDevice(e->device));
    if (e->phase != 0) return e->fail(BH_E_STATE, "bh_step_io: a step is in progress");
    const int64_t n_new = x_in ? n_in : e->n;
    const bool want_out = x_out || y_out || vx_out || vy_out || m_out;
    const bool merge_possible = e->par.merge_min_dist > 0.0 && n_new > 1;
    if (merge_possible || e->world > 1 || nsteps < 1 || n_new == 0 || (want_out && cap_out < n_new)) {
        // plain sequence (the body count may change, or there is nothing to overlap)
        if (x_in) E_RC(bh_set_bodies(e, n_in, x_in, y_in, vx_in, vy_in, m_in));
        E_RC(bh_step(e, nsteps));
        if (n_out) *n_out = e->n;
        if (want_out) return bh_get_bodies(e, cap_out, x_out, y_out, vx_out, vy_out, m_out, n_out);
        return BH_OK;
    }
    if (!e->copy_st) {
        E_TRY(cudaStreamCreateWithFlags(&e->copy_st, cudaStreamNonBlocking));
        E_TRY(cudaEventCreateWithFlags(&e->snap_ev[0], cudaEventDisableTiming));
        E_TRY(cudaEventCreateWithFlags(&e->snap_ev[1], cudaEventDisableTiming));
    }
    if (!e->io_ev[0]) {
        E_TRY(cudaEventCreateWithFlags(&e->io_ev[0], cudaEventDisableTiming));
        E_TRY(cudaEventCreateWithFlags(&e->io_ev[1], cudaEventDisableTiming));
    }
    const size_t bytes = (size_t)n_new * sizeof(double);
    if (x_in) {
        // ---- resetBodies: (x, y) on the compute stream, (vx, vy, m) on the copy stream
        const bool keep_perm = (n_new == e->n && !e->perm_identity && n_new <= e->cap);
        E_RC(e->ensure_bodies(n_new));
        e->n = n_new;
        if (!keep_perm) {
            e->perm_identity = true;
            e->rehome_due = true;
            k_iota<<<grid_for(n_new, 256), 256, 0, e->st>>>(e->perm, (int)n_new);
        }
    }
    if (n_new > e->io_cap) {
        E_TRY(cudaStreamSynchronize(e->copy_st));
        for (auto& q : e->io_stage) dev_free(q);
        const int64_t c = std::max<int64_t>(n_new + n_new / 8, 1024);
        for (auto& q : e->io_stage) E_TRY(dev_alloc(&q, (size_t)c));
        e->io_cap = c;
    }
    if (x_in) {
        E_RC(e->upload_user(e->x, x_in));
        E_RC(e->upload_user(e->y, y_in));
        E_TRY(cudaMemsetAsync(e->ax, 0, bytes, e->st));
        E_TRY(cudaMemsetAsync(e->ay, 0, bytes, e->st));
        // the copy stream must not start before the compute stream is past whatever used vx/vy/m/perm
        E_TRY(cudaEventRecord(e->io_ev[1], e->st));
        E_TRY(cudaStreamWaitEvent(e->copy_st, e->io_ev[1], 0));
        const double* hin[3] = {vx_in, vy_in, m_in};
        double* dev[3] = {e->vx, e->vy, e->m};
        for (int k = 0; k < 3; ++k) {
            if (e->perm_identity) {
                E_TRY(cudaMemcpyAsync(dev[k], hin[k], bytes, cudaMemcpyHostToDevice, e->copy_st));
            } else {
                E_TRY(cudaMemcpyAsync(e->io_stage[k], hin[k], bytes, cudaMemcpyHostToDevice, e->copy_st));
                k_gather<double><<<grid_for(n_new, 256), 256, 0, e->copy_st>>>(dev[k], e->io_stage[k], e->perm, (int)n_new);
            }
        }
        E_TRY(cudaEventRecord(e->io_ev[0], e->copy_st));
        e->io_wait_in = true;
        e->ctr.kernel_launches += 3;
        e->origin_identity = true;
        e->tree_valid = false; e->heavies_valid = false; e->vel_valid = true; e->acc_valid = false;
    }
    // ---- the steps; the last drift triggers the (x, y, m) read-back
    e->io_out.x = x_out; e->io_out.y = y_out; e->io_out.m = m_out;
    e->io_out.armed = x_out || y_out || m_out;
    e->io_steps_left = nsteps;
    int rc = e->run_steps(nsteps);
    e->io_steps_left = 0;
    e->io_out.armed = false;
    if (rc == BH_OK && e->io_wait_in) rc = e->wait_inputs();
    // the last build of the last step replayed jitter clusters: it mutated x/y AFTER the early read-back
    // (the reference's buildTree() does, BH.kt:146-151) — send the positions again
    if (rc == BH_OK && e->jitter_active && (x_out || y_out)) {
        const cudaError_t cj = cudaStreamSynchronize(e->copy_st);
        if (cj != cudaSuccess) rc = e->cuda_fail(cj, "bh_step_io");
        if (rc == BH_OK && x_out) rc = e->download_user(x_out, (const double*)e->x, e->io_stage[0]);
        if (rc == BH_OK && y_out) rc = e->download_user(y_out, (const double*)e->y, e->io_stage[1]);
    }
    // ---- (vx, vy) after the last kick
    if (rc == BH_OK) {
        cudaError_t ce = cudaSuccess;
        if (vx_out) { const int r2 = e->download_user(vx_out, (const double*)e->vx, e->dtmp); if (r2) rc = r2; }
        if (rc == BH_OK && vy_out) {   // own staging buffer so that the two copies overlap
            const int r2 = e->download_user(vy_out, (const double*)e->vy, e->io_stage[3]);
            if (r2) rc = r2;
        }
        (void)ce;
    }
    cudaError_t c1 = cudaStreamSynchronize(e->st), c2 = cudaStreamSynchronize(e->copy_st);
    if (rc == BH_OK && (c1 != cudaSuccess || c2 != cudaSuccess)) rc = e->cuda_fail(c1 != cudaSuccess ? c1 : c2, "bh_step_io");
    if (n_out) *n_out = e->n;
    return rc;
}

int bh_get_slice_index(bh_engine* e, int64_t cap, int32_t* user_index, int64_t* n_slice) {
    if (!e) return BH_E_ARG;
    E_TRY(cudaSetDevice(e->device));
    int64_t lo, hi;
    e->my_slice(&lo, &hi);
    if (n_slice) *n_slice = hi - lo;
    if (cap < hi - lo) return e->fail(BH_E_ARG, "bh_get_slice_index: capacity too small");
    if (user_index && hi > lo) {
        E_TRY(cudaMemcpyAsync(user_index, e->perm + lo, (size_t)(hi - lo) * sizeof(int32_t), cudaMemcpyDeviceToHost, e->st));
        E_TRY(cudaStreamSynchronize(e->st));
    }
    return BH_OK;
}

int64_t bh_slice_epoch(const bh_engine* e) { return e ? e->ctr_rehomes : 0; }

int bh_step_io_slice(bh_engine* e, int32_t nsteps, int64_t n_in, const double* x_in, const double* y_in, const double* vx_in,
                     const double* vy_in, const double* m_in, int64_t cap_out, double* x_out, double* y_out, double* vx_out,
                     double* vy_out, double* m_out, int64_t* n_out) {
    if (!e || nsteps < 0 || (x_in && (!y_in || !vx_in || !vy_in || !m_in)))
        return e ? e->fail(BH_E_ARG, "bh_step_io_slice: bad arguments") : BH_E_ARG;
    E_TRY(cudaSetDevice(e->device));
    if (e->phase != 0) return e->fail(BH_E_STATE, "bh_step_io_slice: a step is in progress");
    if (e->world > 1 && e->transport != bh_engine::T_NCCL) return e->fail(BH_E_STATE, "bh_step_io_slice: needs the NCCL transport");
    if (e->world > 1 && e->merge_enabled()) return e->fail(BH_E_STATE, "bh_step_io_slice: the merge rule re-indexes the list on every rank; disable it (merge_min_dist <= 0)");
    if (!e->copy_st) {
        E_TRY(cudaStreamCreateWithFlags(&e->copy_st, cudaStreamNonBlocking));
        E_TRY(cudaEventCreateWithFlags(&e->snap_ev[0], cudaEventDisableTiming));
        E_TRY(cudaEventCreateWithFlags(&e->snap_ev[1], cudaEventDisableTiming));
    }
    if (!e->io_ev[0]) {
        E_TRY(cudaEventCreateWithFlags(&e->io_ev[0], cudaEventDisableTiming));
        E_TRY(cudaEventCreateWithFlags(&e->io_ev[1], cudaEventDisableTiming));
    }
    int64_t lo, hi;
    e->my_slice(&lo, &hi);
    if (x_in) {
        if (n_in != hi - lo) return e->fail(BH_E_ARG, "bh_step_io_slice: n_in must be the length of this rank's slice (bh_get_slice_index)");
        const size_t bytes = (size_t)(hi - lo) * sizeof(double);
        if (bytes) {
            // positions on the compute stream (the build starts on them); velocities and masses on the copy stream,
            // behind whatever used the old ones — the compute stream waits for them where it first needs them
            E_TRY(cudaMemcpyAsync(e->x + lo, x_in, bytes, cudaMemcpyHostToDevice, e->st));
            E_TRY(cudaMemcpyAsync(e->y + lo, y_in, bytes, cudaMemcpyHostToDevice, e->st));
            E_TRY(cudaEventRecord(e->io_ev[1], e->st));
            E_TRY(cudaStreamWaitEvent(e->copy_st, e->io_ev[1], 0));
            E_TRY(cudaMemcpyAsync(e->m + lo, m_in, bytes, cudaMemcpyHostToDevice, e->copy_st));
            E_TRY(cudaMemcpyAsync(e->vx + lo, vx_in, bytes, cudaMemcpyHostToDevice, e->copy_st));
            E_TRY(cudaMemcpyAsync(e->vy + lo, vy_in, bytes, cudaMemcpyHostToDevice, e->copy_st));
            E_TRY(cudaEventRecord(e->io_ev[0], e->copy_st));
            e->io_wait_in = true;
        }
        e->tree_valid = false; e->acc_valid = false; e->heavies_valid = false;
    }
    // The call is COLLECTIVE (every rank makes it, with or without inputs): each rank treats the other slices as stale,
    // so that all ranks enter the same exchanges (positions / masses / velocities) at the next replicated build even
    // if only some of them brought new state.
    if (e->world > 1) { e->let.pos_valid = false; e->vel_valid = false; e->mass_valid = false; }
    const bool want_out = x_out || y_out || vx_out || vy_out || m_out;
    if (want_out && cap_out < hi - lo) return e->fail(BH_E_ARG, "bh_step_io_slice: capacity too small");
    // the last drift triggers the read-back of the slice's (x, y, m) on the copy stream, under the last evaluation
    e->io_out.x = x_out; e->io_out.y = y_out; e->io_out.m = m_out;
    e->io_out.slice = true; e->io_out.epoch = -1;
    e->io_out.armed = nsteps >= 1 && (x_out || y_out || m_out);
    e->io_steps_left = nsteps;
    int rc = e->run_steps(nsteps);
    e->io_steps_left = 0;
    e->io_out.armed = false; e->io_out.slice = false;
    if (rc == BH_OK && e->io_wait_in) rc = e->wait_inputs();
    if (rc != BH_OK) { cudaStreamSynchronize(e->copy_st); return rc; }
    e->my_slice(&lo, &hi);                     // a re-homing inside the steps re-cuts the slices (bh_slice_epoch changed)
    if (n_out) *n_out = hi - lo;
    if (want_out && cap_out < hi - lo) { cudaStreamSynchronize(e->copy_st); return e->fail(BH_E_ARG, "bh_step_io_slice: capacity too small"); }
    const size_t ob = (size_t)(hi - lo) * sizeof(double);
    if (want_out && ob) {
        // positions and masses again if they were not sent early, or were re-ordered (re-homing) or mutated (jitter) since
        const bool again = e->io_out.epoch != e->ctr_rehomes || e->jitter_active || e->let.returns_applied;
        if (again) {
            E_TRY(cudaStreamSynchronize(e->copy_st));
            if (x_out) E_TRY(cudaMemcpyAsync(x_out, e->x + lo, ob, cudaMemcpyDeviceToHost, e->st));
            if (y_out) E_TRY(cudaMemcpyAsync(y_out, e->y + lo, ob, cudaMemcpyDeviceToHost, e->st));
            if (m_out) E_TRY(cudaMemcpyAsync(m_out, e->m + lo, ob, cudaMemcpyDeviceToHost, e->st));
        }
        if (vx_out) E_TRY(cudaMemcpyAsync(vx_out, e->vx + lo, ob, cudaMemcpyDeviceToHost, e->st));
        if (vy_out) E_TRY(cudaMemcpyAsync(vy_out, e->vy + lo, ob, cudaMemcpyDeviceToHost, e->st));
    }
    const cudaError_t c1 = cudaStreamSynchronize(e->st), c2 = cudaStreamSynchronize(e->copy_st);
    if (c1 != cudaSuccess || c2 != cudaSuccess) return e->cuda_fail(c1 != cudaSuccess ? c1 : c2, "bh_step_io_slice");
    return BH_OK;
}

int bh_build_tree(bh_engine* e) {
    if (!e) return BH_E_ARG;
    E_TRY(cudaSetDevice(e->device));
    E_RC(e->build());
    E_RC(e->finish());
    return BH_OK;
}

int bh_compute_accelerations(bh_engine* e, double* ax, double* ay) {
    if (!e) return BH_E_ARG;
    E_TRY(cudaSetDevice(e->device));
    e->acc_valid = false;
    E_RC(e->evaluate(0, 0, e->n));
    E_RC(e->finish());
    e->add_phase_times(0);
    E_RC(e->download_user(ax, e->ax, e->dtmp));
    E_RC(e->download_user(ay, e->ay, e->dtmp));
    E_TRY(cudaStreamSynchronize(e->st));
    return BH_OK;
}

int bh_direct_sum(bh_engine* e, double* ax, double* ay) {
    if (!e) return BH_E_ARG;
    E_TRY(cudaSetDevice(e->device));
    if (e->n == 0) return BH_OK;
    E_RC(e->sync_positions());
    E_RC(e->sync_masses());
    // results go to the sort buffers (not e->ax/ay, which belong to the integrator)
    double* dax = reinterpret_cast<double*>(e->keys_a);
    double* day = reinterpret_cast<double*>(e->keys_b);
    e->tree_valid = false;
    E_TRY(cudaEventRecord(e->call_ev[0], e->st));
    k_direct<<<grid_for(e->n, DS_TILE), DS_TILE, 0, e->st>>>(e->x, e->y, e->m, (int)e->n, (float)e->par.soft2, e->par.G, dax, day);
    E_TRY(cudaEventRecord(e->call_ev[1], e->st));
    e->ctr.kernel_launches += 1;
    E_TRY(cudaGetLastError());
    E_TRY(cudaEventSynchronize(e->call_ev[1]));
    float dms = 0.f;
    if (cudaEventElapsedTime(&dms, e->call_ev[0], e->call_ev[1]) == cudaSuccess) e->ctr.ms_direct = dms;
    E_RC(e->download_user(ax, (const double*)dax, e->dtmp));
    E_RC(e->download_user(ay, (const double*)day, e->dtmp));
    E_TRY(cudaStreamSynchronize(e->st));
    return BH_OK;
}

int bh_energy(bh_engine* e, double* ke, double* pe, double* px, double* py) {
    if (!e) return BH_E_ARG;
    E_TRY(cudaSetDevice(e->device));
    double h[4] = {0, 0, 0, 0};
    if (e->n > 0) {
        E_RC(e->sync_positions());
        E_RC(e->sync_velocities());
        E_RC(e->sync_masses());
        E_TRY(cudaMemsetAsync(e->red, 0, 4 * sizeof(double), e->st));
        k_energy<<<grid_for(e->n, DS_TILE), DS_TILE, 0, e->st>>>(e->x, e->y, e->vx, e->vy, e->m, (int)e->n, e->par.soft2, e->red);
        e->ctr.kernel_launches += 1;
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

int bh_energy_tree(bh_engine* e, double theta, double* ke, double* pe, double* px, double* py) {
    if (!e) return BH_E_ARG;
    E_TRY(cudaSetDevice(e->device));
    double h[4] = {0, 0, 0, 0};
    if (e->n > 0) {
        E_RC(e->sync_velocities());
        e->acc_valid = false;
        E_RC(e->build());                               // buildTree() on the current state
        E_RC(e->finish());
        const BhWalkParams w = bh_walk_params(theta > 0.0 ? theta : e->par.theta, e->par.soft2, e->par.root_half);
        E_TRY(cudaMemsetAsync(e->red, 0, 4 * sizeof(double), e->st));
        k_energy_tree<<<grid_for(e->n, 128), 128, 0, e->st>>>(e->view(), w, (int)e->n, e->x, e->y, e->vx, e->vy, e->m, e->leafpos, e->red);
        e->ctr.kernel_launches += 1;
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
    if (!e->tree_valid) E_RC(bh_build_tree(e));
    const int n = (int)e->n;
    if (n == 0) return BH_OK;
    uint64_t *dkey = nullptr, *dkey_u = nullptr;
    int* ddepth = nullptr;
    cudaError_t ce = dev_alloc(&dkey, n);
    if (ce == cudaSuccess) ce = dev_alloc(&dkey_u, n);
    if (ce == cudaSuccess) ce = dev_alloc(&ddepth, n);
    int rc = BH_OK;
    if (ce == cudaSuccess) {
        // sentinel here is the ABI's UINT64_MAX, not the sortable 1<<2L
        k_keygen<<<grid_for(n, 256), 256, 0, e->st>>>(e->x, e->y, n, e->root, bh_make_grid(e->root), BH_KEY_NOT_IN_TREE, dkey, nullptr);
        k_leaf_depth<<<grid_for(n, 256), 256, 0, e->st>>>(e->view(), e->leafpos, e->jitter_active ? e->jflag : nullptr, n, ddepth);
        k_scatter<uint64_t><<<grid_for(n, 256), 256, 0, e->st>>>(dkey_u, dkey, e->perm, n);
        ce = cudaGetLastError();
        if (ce == cudaSuccess && key) ce = cudaMemcpyAsync(key, dkey_u, (size_t)n * 8, cudaMemcpyDeviceToHost, e->st);
        if (ce == cudaSuccess && depth) {
            k_scatter<int><<<grid_for(n, 256), 256, 0, e->st>>>(e->itmp, ddepth, e->perm, n);
            ce = cudaMemcpyAsync(depth, e->itmp, (size_t)n * 4, cudaMemcpyDeviceToHost, e->st);
        }
        if (ce == cudaSuccess && order) {   // sorted position -> USER index
            k_gather<int><<<grid_for(n, 256), 256, 0, e->st>>>(ddepth, e->perm, e->order, n);
            ce = cudaMemcpyAsync(order, ddepth, (size_t)n * 4, cudaMemcpyDeviceToHost, e->st);
        }
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(e->st);
        e->ctr.kernel_launches += 5;
    }
    cudaFree(dkey); cudaFree(dkey_u); cudaFree(ddepth);
    if (ce != cudaSuccess) return e->cuda_fail(ce, "bh_get_morton");
    return rc;
}

int bh_get_tree(bh_engine* e, int64_t cap, int64_t* n_cells, double* cx, double* cy, double* h, double* mass,
                double* comx, double* comy, int32_t* body) {
    if (!e) return BH_E_ARG;
    E_TRY(cudaSetDevice(e->device));
    if (!e->tree_valid) E_RC(bh_build_tree(e));   // lastTree ?: buildTree(), BH.kt:329-332
    try {
        const size_t M = (size_t)e->M, ni = (size_t)e->n_in;
        std::vector<uint64_t> keys(ni);
        std::vector<int> order(ni), S(ni + 1), perm;
        std::vector<BhCellS> sk(M);
        std::vector<BhCellD> cd(M);
        if (ni) {
            E_TRY(cudaMemcpy(keys.data(), e->keys_sorted, ni * 8, cudaMemcpyDeviceToHost));
            E_TRY(cudaMemcpy(order.data(), e->order, ni * 4, cudaMemcpyDeviceToHost));
            E_TRY(cudaMemcpy(S.data(), e->S, (ni + 1) * 4, cudaMemcpyDeviceToHost));
            E_TRY(cudaMemcpy(sk.data(), e->sk, M * sizeof(BhCellS), cudaMemcpyDeviceToHost));
            E_TRY(cudaMemcpy(cd.data(), e->cd, M * sizeof(BhCellD), cudaMemcpyDeviceToHost));
            if (!e->perm_identity) {   // the export names bodies by USER index
                perm.resize((size_t)e->n);
                E_TRY(cudaMemcpy(perm.data(), e->perm, (size_t)e->n * 4, cudaMemcpyDeviceToHost));
                for (size_t i = 0; i < ni; ++i) order[i] = perm[(size_t)order[i]];
            }
        }
        std::vector<int> jf;   // jitter flags per sorted position
        if (e->jitter_active && ni) {
            std::vector<int> jh((size_t)e->n), ord_home(ni);
            E_TRY(cudaMemcpy(jh.data(), e->jflag, (size_t)e->n * 4, cudaMemcpyDeviceToHost));
            E_TRY(cudaMemcpy(ord_home.data(), e->order, ni * 4, cudaMemcpyDeviceToHost));
            jf.resize(ni);
            for (size_t i = 0; i < ni; ++i) jf[i] = jh[(size_t)ord_home[i]];
        }
        BhHostTree t{e->root, e->n_in, e->M, keys.data(), order.data(), S.data(), sk.data(), cd.data(), jf.empty() ? nullptr : jf.data()};
        BhCellsOut out;
        out.cap = cap; out.cx = cx; out.cy = cy; out.h = h; out.mass = mass; out.comx = comx; out.comy = comy; out.body = body;
        bh_export_cells(t, out);
        if (n_cells) *n_cells = out.count;
        if (cap != 0 && cap < out.count) return e->fail(BH_E_ARG, "bh_get_tree: capacity too small");
    } catch (const std::bad_alloc&) { return e->fail(BH_E_OOM, "bh_get_tree: host out of memory"); }
    return BH_OK;
}

int bh_get_tree_root(bh_engine* e, double* mass, double* comx, double* comy, int64_t* n_cells) {
    if (!e) return BH_E_ARG;
    E_TRY(cudaSetDevice(e->device));
    if (!e->tree_valid) E_RC(bh_build_tree(e));   // lastTree ?: buildTree(), BH.kt:329-332
    BhCellD r; r.comx = e->root.cx; r.comy = e->root.cy; r.mass = 0.0; r.pad = 0.0;   // empty tree: BH.kt:179-183
    if (e->M > 0) E_TRY(cudaMemcpy(&r, e->cd, sizeof(BhCellD), cudaMemcpyDeviceToHost));
    if (mass) *mass = r.mass;
    if (comx) *comx = r.comx;
    if (comy) *comy = r.comy;
    if (n_cells) *n_cells = e->M;
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
    for (auto& v : e->let.ms_phase) v = 0.0;
    for (auto& v : e->let.cpu_us) v = 0.0;
    e->let.n_folds = 0; e->let.pe_armed = false;
    e->ctr = bh_counters{};
    e->ctr.n_in_tree = keep.n_in_tree; e->ctr.n_out_of_box = keep.n_out_of_box; e->ctr.n_cells = keep.n_cells;
    e->ctr.n_internal = keep.n_internal; e->ctr.key_levels = keep.key_levels; e->ctr.max_depth = keep.max_depth;
    e->ctr.n_jitter_bodies = keep.n_jitter_bodies;
    e->ctr.ms_step_call = keep.ms_step_call;
    e->ctr.bbox_min_x = keep.bbox_min_x; e->ctr.bbox_max_x = keep.bbox_max_x; e->ctr.bbox_min_y = keep.bbox_min_y; e->ctr.bbox_max_y = keep.bbox_max_y;
    e->ctr.ms_direct = keep.ms_direct;
    return BH_OK;
}

int bh_get_body_counts(bh_engine* e, int32_t* interactions, int32_t* opened) {
    if (!e) return BH_E_ARG;
    if (!(e->cfg.flags & BH_FLAG_BODY_COUNTS)) return e->fail(BH_E_STATE, "bh_get_body_counts: engine created without BH_FLAG_BODY_COUNTS");
    E_TRY(cudaSetDevice(e->device));
    if (e->n > 0) {
        E_RC(e->download_user<int>(interactions, e->cntI, e->itmp));
        E_RC(e->download_user<int>(opened, e->cntO, e->itmp));
        E_TRY(cudaStreamSynchronize(e->st));
    }
    return BH_OK;
}

// ---- multi-process --------------------------------------------------------------------------
int bh_slice_bounds(int64_t n, int32_t world, int32_t rank, int64_t* lo, int64_t* hi) {
    if (n < 0 || world < 1 || rank < 0 || rank >= world || !lo || !hi) return BH_E_ARG;
    const int64_t per = (n + world - 1) / world;
    *lo = std::min<int64_t>(n, per * rank);
    *hi = std::min<int64_t>(n, per * (rank + 1));
    return BH_OK;
}

int bh_comm_unique_id(void* id_out, int32_t id_bytes) {
    if (!id_out || id_bytes < (int32_t)sizeof(bhcomm::UniqueId)) return BH_E_ARG;
    bhcomm::Api& A = bhcomm::api();
    if (!A.ok) { g_create_err = A.err; return BH_E_NCCL; }
    bhcomm::UniqueId id;
    if (A.GetUniqueId(&id) != bhcomm::kSuccess) { g_create_err = "ncclGetUniqueId failed"; return BH_E_NCCL; }
    memcpy(id_out, &id, sizeof(id));
    return BH_OK;
}

int bh_comm_init(bh_engine* e, int32_t rank, int32_t world, const void* id, int32_t id_bytes) {
    if (!e) return BH_E_ARG;
    if (world < 1 || world > MAX_WORLD || rank < 0 || rank >= world || !id || id_bytes < (int32_t)sizeof(bhcomm::UniqueId))
        return e->fail(BH_E_ARG, "bh_comm_init: bad arguments");
    if (e->transport != bh_engine::T_NONE) return e->fail(BH_E_STATE, "bh_comm_init: a transport is already set");
    E_TRY(cudaSetDevice(e->device));
    bhcomm::Api& A = bhcomm::api();
    if (!A.ok) return e->fail(BH_E_NCCL, A.err);
    bhcomm::UniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    const int rc = A.CommInitRank(&e->comm, world, uid, rank);
    if (rc != bhcomm::kSuccess) return e->nccl_fail(rc, "ncclCommInitRank");
    e->rank = rank; e->world = world; e->transport = bh_engine::T_NCCL;
    // domain mode: BH_FLAG_LET, or BH_LET=1/0 in the environment (overrides the flag)
    e->let.enabled = (e->cfg.flags & BH_FLAG_LET) != 0;
    if (const char* s = getenv("BH_LET")) e->let.enabled = atoi(s) != 0;
    if (const char* s = getenv("BH_LET_MIN_WORLD")) e->let.min_world = std::max(2, atoi(s));
    if (world > 16) e->let.enabled = false;
    if (e->let.enabled) e->rehome_due = true;
    return BH_OK;
}

int bh_comm_init_external(bh_engine* e, int32_t rank, int32_t world) {
    if (!e) return BH_E_ARG;
    if (world < 1 || world > MAX_WORLD || rank < 0 || rank >= world) return e->fail(BH_E_ARG, "bh_comm_init_external: bad arguments");
    if (e->transport != bh_engine::T_NONE) return e->fail(BH_E_STATE, "bh_comm_init_external: a transport is already set");
    e->rank = rank; e->world = world; e->transport = bh_engine::T_EXTERNAL;
    return BH_OK;
}

int bh_export_slice(bh_engine* e, int32_t field, int64_t cap, double* a, double* b, int64_t* lo_out, int64_t* hi_out) {
    if (!e || (field != BH_FIELD_POS && field != BH_FIELD_VEL)) return e ? e->fail(BH_E_ARG, "bh_export_slice: bad arguments") : BH_E_ARG;
    E_TRY(cudaSetDevice(e->device));
    int64_t lo, hi;
    e->my_slice(&lo, &hi);
    if (lo_out) *lo_out = lo;
    if (hi_out) *hi_out = hi;
    if (cap < hi - lo) return e->fail(BH_E_ARG, "bh_export_slice: capacity too small");
    const double* sa = field == BH_FIELD_POS ? e->x : e->vx;
    const double* sb = field == BH_FIELD_POS ? e->y : e->vy;
    if (hi > lo) {
        if (a) E_TRY(cudaMemcpyAsync(a, sa + lo, (size_t)(hi - lo) * sizeof(double), cudaMemcpyDeviceToHost, e->st));
        if (b) E_TRY(cudaMemcpyAsync(b, sb + lo, (size_t)(hi - lo) * sizeof(double), cudaMemcpyDeviceToHost, e->st));
    }
    E_TRY(cudaStreamSynchronize(e->st));
    return BH_OK;
}

int bh_import_slices(bh_engine* e, int32_t field, int64_t n, const double* a, const double* b) {
    if (!e || (field != BH_FIELD_POS && field != BH_FIELD_VEL) || !a || !b) return e ? e->fail(BH_E_ARG, "bh_import_slices: bad arguments") : BH_E_ARG;
    if (n != e->n) return e->fail(BH_E_ARG, "bh_import_slices: n must equal bh_num_bodies");
    E_TRY(cudaSetDevice(e->device));
    double* da = field == BH_FIELD_POS ? e->x : e->vx;
    double* db = field == BH_FIELD_POS ? e->y : e->vy;
    if (n > 0) {
        E_TRY(cudaMemcpyAsync(da, a, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, e->st));
        E_TRY(cudaMemcpyAsync(db, b, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, e->st));
    }
    E_TRY(cudaStreamSynchronize(e->st));
    if (field == BH_FIELD_VEL) e->vel_valid = true;
    else { e->tree_valid = false; e->acc_valid = false; }
    return BH_OK;
}

int bh_evaluate_slice(bh_engine* e, int64_t cap, double* ax, double* ay, int32_t* user_index, int64_t* n_slice) {
    if (!e) return BH_E_ARG;
    E_TRY(cudaSetDevice(e->device));
    if (e->phase != 0) return e->fail(BH_E_STATE, "bh_evaluate_slice: a step is in progress");
    e->acc_valid = false;
    E_RC(e->evaluate_slice(0));
    E_RC(e->finish());
    e->add_phase_times(0);
    int64_t lo, hi;
    e->my_slice(&lo, &hi);                 // (a re-homing build inside the evaluation may have re-cut the slices)
    if (n_slice) *n_slice = hi - lo;
    if (cap < hi - lo) return e->fail(BH_E_ARG, "bh_evaluate_slice: capacity too small");
    if (hi > lo) {
        const size_t k = (size_t)(hi - lo);
        if (ax) E_TRY(cudaMemcpyAsync(ax, e->ax + lo, k * sizeof(double), cudaMemcpyDeviceToHost, e->st));
        if (ay) E_TRY(cudaMemcpyAsync(ay, e->ay + lo, k * sizeof(double), cudaMemcpyDeviceToHost, e->st));
        if (user_index) E_TRY(cudaMemcpyAsync(user_index, e->perm + lo, k * sizeof(int32_t), cudaMemcpyDeviceToHost, e->st));
        E_TRY(cudaStreamSynchronize(e->st));
    }
    return BH_OK;
}

int bh_set_domain_mode(bh_engine* e, int32_t enabled) {
    if (!e) return BH_E_ARG;
    if (e->phase != 0) return e->fail(BH_E_STATE, "bh_set_domain_mode: a step is in progress");
    e->let.enabled = enabled != 0 && e->world > 1 && e->world <= 16 && e->transport == bh_engine::T_NCCL;
    return BH_OK;
}

int bh_get_let_stats(bh_engine* e, int64_t* out, int32_t n_out) {
    if (!e || !out || n_out < 1) return BH_E_ARG;
    int64_t v[40] = {e->let.enabled ? 1 : 0, e->let.part_valid ? 1 : 0, e->let.ell, e->let.evaluations, e->let.fallbacks,
                     e->let.M, e->let.last_imported, e->let.last_sent, e->let.last_strays, e->let.n_items};
    for (int k = 0; k < 7; ++k) v[10 + k] = 0;
    v[10] = e->let.fb_jitter; v[11] = e->let.fb_strays; v[12] = e->let.fb_cells; v[13] = e->let.stray_cap; v[14] = e->let.jret_total;
    if (e->let.dcnt) {   // diagnostics of the last evaluation: stray leaf look-ups by descent / by the fall-back scan
        int dc[8] = {0};
        if (cudaMemcpy(dc, e->let.dcnt, 8 * sizeof(int), cudaMemcpyDeviceToHost) == cudaSuccess) { v[15] = dc[4]; v[16] = dc[5]; }
    }
    v[17] = e->let.n_folds;
    if (e->let.ipc_ok) v[0] = 2;   // enabled, blocks imported over peer memory
    for (int k = 0; k < 8; ++k) v[18 + k] = 0;
    v[25] = g_let_grows;
    if (getenv("BH_LET_TIMERS") && atoi(getenv("BH_LET_TIMERS")) != 0 && e->let.n_folds > 0) {   // per-rank phase times (diagnostics)
        char line[1024];
        int off = snprintf(line, sizeof(line), "[let rank %d] evals %lld grows %lld gpu_us/eval:", e->rank, (long long)e->let.n_folds, (long long)g_let_grows);
        for (int k = 0; k < 12; ++k) off += snprintf(line + off, sizeof(line) - off, " %.0f", e->let.ms_phase[k] * 1000.0 / (double)e->let.n_folds);
        off += snprintf(line + off, sizeof(line) - off, "  cpu_us/eval:");
        for (int k = 0; k < 12; ++k) off += snprintf(line + off, sizeof(line) - off, " %.0f", e->let.cpu_us[k] / (double)e->let.n_folds);
        off += snprintf(line + off, sizeof(line) - off, "  M %d imported %lld own %lld walk_ms %.3f\n", e->let.M, (long long)e->let.last_imported,
                        (long long)(e->let.cut[e->rank + 1] - e->let.cut[e->rank]), e->ctr.ms_walk / std::max<double>(1.0, (double)e->ctr.total_evaluations));
        fputs(line, stderr);
    }
    for (int k = 0; k < n_out && k < 26; ++k) out[k] = v[k];
    return BH_OK;
}

// ---- diagnostics ------------------------------------------------------------------------------
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

}  // extern "C"
