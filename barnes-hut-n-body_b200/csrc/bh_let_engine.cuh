// bh_let_engine.cuh — bh_engine methods of the locally-essential-tree mode (see bh_let.cuh).
#ifndef BH_LET_ENGINE_CUH
#define BH_LET_ENGINE_CUH

#define LET_GROW(field, need) BH_TRY(let_grow(let.field, let.field##_cap, (int64_t)(need)))

inline bool bh_engine::let_usable() const {
    // from min_world ranks up (4: below, replicating a few-times-larger tree is cheaper than assembling a LET), or from 2
    // ranks when the replicated build itself is the cost (32 M bodies and more: measured at C4, 100 M bodies on 2 ranks)
    return let.enabled && (world >= let.min_world || (world >= 2 && n >= ((int64_t)32 << 20))) && world <= 16 && transport == T_NCCL &&
           !merge_enabled();
}
inline bool bh_engine::let_ready() const {
    return let_usable() && let.part_valid && let.n_part == n && !rehome_due && let.part_root.cx == par.root_cx &&
           let.part_root.cy == par.root_cy && let.part_root.half == par.root_half;
}

// every rank needs every body's position (replicated build, read-back, diagnostics)
inline int bh_engine::sync_positions() {
    if (let.pos_valid || world <= 1) { let.pos_valid = true; return BH_OK; }
    if (transport != T_NCCL) return fail(BH_E_STATE, "positions are not replicated");
    BH_RC(all_gather_pair(x, y));
    let.pos_valid = true;
    return BH_OK;
}

// After a re-homing build (home order = global Morton order, replicated): cut the slices at code
// boundaries.  Every rank computes the same cuts from the same keys.
inline int bh_engine::let_partition() {
    let.part_valid = false;
    const int ell = bh_let_choose_ell((long long)n, root.levels);
    let.n_declined = n;
    if (ell < 1 || n_in < 2 * world) return BH_OK;      // too small / too shallow: stay replicated
    let.ell = ell;
    let.lam = bh_let_lambda(ell);
    let.ncodes = 1u << (2 * ell);
    let.bw = (int)std::max<int64_t>(1, ((int64_t)1 << (2 * let.lam)) / 64);
    if (!let.dcnt) {
        BH_TRY(cudaMalloc((void**)&let.dcnt, 16 * sizeof(int)));
        BH_TRY(cudaMalloc((void**)&let.dcollect, 64 * sizeof(int)));
        BH_TRY(cudaMallocHost((void**)&let.hcollect, 64 * sizeof(int)));
        BH_TRY(cudaMallocHost((void**)&let.hhdr, 2 * 17 * sizeof(double)));
        BH_TRY(cudaMallocHost((void**)&let.hcut, 64 * sizeof(int)));
    }
    k_let_cut<<<1, 32, 0, st>>>(keys_sorted, n_in, (int)n, root.levels, ell, world, let.dcollect);
    ctr.kernel_launches += 1;
    BH_TRY(cudaMemcpyAsync(let.hcut, let.dcollect, 34 * sizeof(int), cudaMemcpyDeviceToHost, st));
    BH_TRY(cudaStreamSynchronize(st));
    BH_TRY(cudaGetLastError());
    let.split.world = world; let.split.me = rank;
    for (int r = 0; r <= world; ++r) { let.cut[r] = let.hcut[r]; let.split.cs[r] = (uint32_t)let.hcut[17 + r]; }
    for (int r = 0; r < world; ++r) if (let.cut[r + 1] < let.cut[r]) return BH_OK;
    // stray slots per rank: the local build (own slice + (P-1) x stray_cap guest slots) runs in the engine's per-body buffers
    int64_t max_slice = 0;
    for (int r = 0; r < world; ++r) max_slice = std::max(max_slice, let.cut[r + 1] - let.cut[r]);
    const int64_t room = (cap - max_slice) / (world - 1);
    let.stray_cap = std::min<int64_t>(room, std::max<int64_t>(n / world / 64, 1024));
    if (let.stray_cap < 16 || let.stray_cap > room) return BH_OK;
    if (const char* sc_env = getenv("BH_LET_STRAY_CAP")) let.stray_cap = std::max(1, atoi(sc_env));   // test hook: force overflows -> fallbacks
    let.seg_len = LET_SEG_HDR + let.bw + 4 * let.stray_cap;
    BH_RC(let_map_peers());
    let.part_root = root;
    let.n_part = n;
    let.part_valid = true;
    let.n_declined = -1;
    return BH_OK;
}

// Map every rank's local-tree arrays (cd, sk) into this process (CUDA IPC; NVLink peer memory), so that
// k_let_blocks reads the blocks it imports straight from their owner — no packing, no send/recv.  The
// arrays are the engine's own cell arrays: sized by the re-homing build of the GLOBAL tree, so a local
// tree always fits and the pointers only change when a re-homing build grows them, i.e. right before
// this (collective) call.  The existing collectives order the accesses: importers read after the table
// all-reduce (issued behind the owner's climb) and an owner overwrites its arrays only behind the next
// segment all-gather / position gather, which cannot complete before every rank has got there.
// All ranks agree (all-reduce) on whether the mapping worked; if not, blocks go through ncclSend/ncclRecv.
inline int bh_engine::let_map_peers() {
    bhcomm::Api& A = bhcomm::api();
    typedef bh_let_state::IpcPair IpcPair;
    static_assert(sizeof(IpcPair) % sizeof(double) == 0, "IpcPair must be a whole number of doubles");
    const size_t pair_d = sizeof(IpcPair) / sizeof(double);
    if (const char* s = getenv("BH_LET_IPC")) let.ipc_wanted = atoi(s) != 0;
    if (!let.ipc_host) {
        BH_TRY(cudaMallocHost((void**)&let.ipc_host, 17 * sizeof(IpcPair)));
        BH_TRY(cudaMalloc((void**)&let.ipc_dev, 17 * sizeof(IpcPair) + 64));
    }
    double failed = 0.0;
    IpcPair mine;
    memset(&mine, 0, sizeof(mine));
    if (!let.ipc_wanted || cudaIpcGetMemHandle(&mine.cd, cd) != cudaSuccess || cudaIpcGetMemHandle(&mine.sk, sk) != cudaSuccess) failed = 1.0;
    (void)cudaGetLastError();
    let.ipc_host[rank] = mine;
    BH_TRY(cudaMemcpyAsync(reinterpret_cast<char*>(let.ipc_dev) + (size_t)rank * sizeof(IpcPair), &let.ipc_host[rank], sizeof(IpcPair),
                           cudaMemcpyHostToDevice, st));
    int rc = A.AllGather(reinterpret_cast<char*>(let.ipc_dev) + (size_t)rank * sizeof(IpcPair), let.ipc_dev, pair_d, bhcomm::kFloat64, comm, st);
    if (rc != bhcomm::kSuccess) return nccl_fail(rc, "ncclAllGather(ipc handles)");
    BH_TRY(cudaMemcpyAsync(let.ipc_host, let.ipc_dev, (size_t)world * sizeof(IpcPair), cudaMemcpyDeviceToHost, st));
    BH_TRY(cudaStreamSynchronize(st));
    if (failed == 0.0) {
        for (int q = 0; q < world && failed == 0.0; ++q) {
            if (q == rank) { let.peers.cd[q] = cd; let.peers.sk[q] = sk; continue; }
            if (let.ipc_open[q] && memcmp(&let.ipc_cached[q], &let.ipc_host[q], sizeof(IpcPair)) == 0) continue;
            if (let.ipc_open[q]) { cudaIpcCloseMemHandle(let.ipc_ptr_cd[q]); cudaIpcCloseMemHandle(let.ipc_ptr_sk[q]); let.ipc_open[q] = false; }
            void *pc = nullptr, *ps = nullptr;
            if (cudaIpcOpenMemHandle(&pc, let.ipc_host[q].cd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { failed = 1.0; break; }
            if (cudaIpcOpenMemHandle(&ps, let.ipc_host[q].sk, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaIpcCloseMemHandle(pc); failed = 1.0; break;
            }
            let.ipc_ptr_cd[q] = pc; let.ipc_ptr_sk[q] = ps; let.ipc_open[q] = true;
            let.ipc_cached[q] = let.ipc_host[q];
            let.peers.cd[q] = static_cast<const BhCellD*>(pc);
            let.peers.sk[q] = static_cast<const BhCellS*>(ps);
        }
        (void)cudaGetLastError();
    }
    // agreement: peer memory is used only if it works on every rank
    double* flag = let.ipc_dev + 17 * pair_d;
    BH_TRY(cudaMemcpyAsync(flag, &failed, sizeof(double), cudaMemcpyHostToDevice, st));
    rc = A.AllReduce(flag, flag, 1, bhcomm::kFloat64, bhcomm::kSum, comm, st);
    if (rc != bhcomm::kSuccess) return nccl_fail(rc, "ncclAllReduce(ipc agreement)");
    double total = 1.0;
    BH_TRY(cudaMemcpyAsync(&total, flag, sizeof(double), cudaMemcpyDeviceToHost, st));
    BH_TRY(cudaStreamSynchronize(st));
    let.ipc_ok = total == 0.0;
    return BH_OK;
}

// One force evaluation over the locally essential tree.  Returns BH_LET_RETRY when the evaluation
// has to be redone after a re-homing (stray overflow, or a stray inside a jitter cluster).
constexpr int BH_LET_RETRY = 1;
inline int bh_engine::let_evaluate(int slot) {
    bhcomm::Api& A = bhcomm::api();
    const int64_t lo = let.cut[rank], hi = let.cut[rank + 1];
    const int n_own = (int)(hi - lo);
    const int ell = let.ell, P = world;
    const uint32_t ncodes = let.ncodes, c_lo = let.split.cs[rank], c_hi = let.split.cs[rank + 1], mylen = c_hi - c_lo;
    const int cap = (int)let.stray_cap;
    root = BhRoot{par.root_cx, par.root_cy, par.root_half, bh_key_levels(par.root_half)};
    const BhGrid grid = bh_make_grid(root);
    BH_TRY(cudaEventRecord(ev[slot + 0], st));
    // phase timers: fold the previous evaluation's events (complete by now: that evaluation ended in a walk + sync)
    static const bool timers = getenv("BH_LET_TIMERS") && atoi(getenv("BH_LET_TIMERS")) != 0;
    if (!let.pe[0]) for (auto& e : let.pe) BH_TRY(cudaEventCreate(&e));
    if (timers && let.pe_armed && cudaEventSynchronize(let.pe[bh_let_state::NPH]) == cudaSuccess) {
        let.n_folds++;
        for (int k = 0; k < bh_let_state::NPH; ++k) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, let.pe[k], let.pe[k + 1]) == cudaSuccess) let.ms_phase[k] += ms;
        }
    }
    let.pe_armed = false;
#define LET_PHASE(k)                                                                                   \
    do {                                                                                                \
        BH_TRY(cudaEventRecord(let.pe[k], st));                                                         \
        let.cpu_t[k] = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); \
        if ((k) > 0) let.cpu_us[(k) - 1] += let.cpu_t[k] - let.cpu_t[(k) - 1];                            \
    } while (0)
    LET_PHASE(0);

    // ---- 1. local arrays, footprint, strays
    BH_RC(wait_inputs());            // bh_step_io_slice: the slice's masses may still be in flight on the copy stream
    const int64_t nl_max = (int64_t)n_own + (int64_t)(P - 1) * cap;
    LET_GROW(lx, nl_max + PAD); LET_GROW(ly, nl_max + PAD); LET_GROW(lm, nl_max + PAD);
    LET_GROW(lperm, nl_max + PAD); LET_GROW(lleaf, nl_max + PAD);
    LET_GROW(gsrc, (int64_t)(P - 1) * cap + 1); LET_GROW(stray_slot, (int64_t)cap + 1);
    LET_GROW(segs, let.seg_len * P);
    double* myseg = let.segs + (size_t)rank * let.seg_len;
    k_let_seg_init<<<grid_for(LET_SEG_HDR + let.bw, 256), 256, 0, st>>>(myseg, let.bw, let.dcnt);
    if (n_own > 0)
        k_let_local<<<grid_for(n_own, 256), 256, 0, st>>>(n_own, x + lo, y + lo, m + lo, perm + lo, root, grid, ell, let.lam, c_lo, c_hi,
                                                            cap, let.lx, let.ly, let.lm, let.lperm, myseg, let.bw, let.dcnt, let.stray_slot);
    k_let_seg_header<<<1, 1, 0, st>>>(myseg, let.dcnt, cap);
    ctr.kernel_launches += 3;
    LET_PHASE(1);   // 0: own slice -> local arrays, footprint, strays
    int rc = A.AllGather(myseg, let.segs, (size_t)let.seg_len, bhcomm::kFloat64, comm, st);
    if (rc != bhcomm::kSuccess) return nccl_fail(rc, "ncclAllGather(segments)");
    LET_PHASE(2);   // 1: segment all-gather (includes waiting for the slowest rank)
    // ---- 2. guests: at most (P-1)*cap; unused slots stay NaN = outside the root box = not in the tree
    const int64_t others = (int64_t)(P - 1) * cap;
    const int n_loc = n_own + (int)others;
    BH_TRY(cudaMemsetAsync(let.lx + n_own, 0xFF, (size_t)others * sizeof(double), st));
    BH_TRY(cudaMemsetAsync(let.ly + n_own, 0xFF, (size_t)others * sizeof(double), st));
    BH_TRY(cudaMemsetAsync(let.lm + n_own, 0, (size_t)others * sizeof(double), st));
    BH_TRY(cudaMemsetAsync(let.lperm + n_own, 0, (size_t)others * sizeof(int), st));
    k_let_guests<<<grid_for((int64_t)P * cap, 256), 256, 0, st>>>(let.segs, let.seg_len, let.bw, cap, P, rank, root, grid, ell, c_lo, c_hi,
                                                                   n_own, (int)others, let.lx, let.ly, let.lm, let.lperm, let.dcnt, let.gsrc);
    ctr.kernel_launches += 1;
    LET_PHASE(3);   // 2: guests
    // ---- 3. the single-GPU build on the local arrays (keys outside [c_lo, c_hi) -> not in the tree)
    {
        double *sx = x, *sy = y, *sm = m;
        int *sperm = perm, *sleaf = leafpos;
        const int64_t sn = n;
        x = let.lx; y = let.ly; m = let.lm; perm = let.lperm; leafpos = let.lleaf; n = n_loc;
        let.local_build = true;
        let.local_overflow = false;
        const int brc = build(slot, false);
        let.local_build = false;
        x = sx; y = sy; m = sm; perm = sperm; leafpos = sleaf; n = sn;
        BH_RC(brc);
    }
    LET_PHASE(4);   // 3: local build
    tree_valid = false;              // the engine's cell arrays hold the LOCAL tree: exports rebuild the global one
    const bool jit = jitter_active;
    if (let.local_overflow) BH_TRY(cudaMemsetAsync(let.dcnt + LET_D_FLAG, 0x04, sizeof(int), st));   // -> retry flag of this rank (reason 4)
    // ---- 4. level-ELL summaries -> replicated table (P extra entries carry the ranks' retry flags, P x LET_JRET more the
    //         positions of guests that a jitter replay on their host rank mutated)
    const size_t table_entries = (size_t)ncodes + P + (size_t)P * LET_JRET;
    LET_GROW(table, (int64_t)table_entries + 1);
    BH_TRY(cudaMemsetAsync(let.table, 0, table_entries * sizeof(BhLetEntry), st));
    if (jit && n_in > 0) {
        k_let_jitter_returns<<<grid_for(n_in, 256), 256, 0, st>>>(keys_sorted, order, n_in, n_own, let.lx, let.ly, let.gsrc, cap,
                                                                    let.table + ncodes + P + (size_t)rank * LET_JRET, let.dcnt);
        ctr.kernel_launches += 1;
    }
    if (n_in > 0) {
        k_let_summary<<<grid_for(n_in, 256), 256, 0, st>>>(view(), root.levels, ell, let.lx, let.ly, let.lm, jit ? jflag : nullptr, let.table);
        ctr.kernel_launches += 1;
    }
    k_let_flag<<<1, 1, 0, st>>>(let.table + ncodes + rank, let.dcnt + LET_D_FLAG, let.segs, let.seg_len, P);
    LET_PHASE(5);   // 4: summaries
    rc = A.AllReduce(let.table, let.table, table_entries * 6, bhcomm::kFloat64, bhcomm::kSum, comm, st);
    if (rc != bhcomm::kSuccess) return nccl_fail(rc, "ncclAllReduce(table)");
    LET_PHASE(6);   // 5: table all-reduce
    // ---- 5. plan
    const int n_slots = 2 * (int)ncodes;
    LET_GROW(nit, ncodes + 1); LET_GROW(blk, ncodes + 1); LET_GROW(recvsz, ncodes + 1); LET_GROW(recvoff, ncodes + 2);
    LET_GROW(item_first, ncodes + 2); LET_GROW(dst, ncodes + 1);
    LET_GROW(sendsz, (int64_t)P * mylen + 1); LET_GROW(sendoff, (int64_t)P * mylen + 2);
    LET_GROW(ikey, n_slots + 1); LET_GROW(itype, n_slots + 1); LET_GROW(iw, n_slots + 1); LET_GROW(icnt, n_slots + 1);
    LET_GROW(iS, n_slots + 2); LET_GROW(iW, n_slots + 2); LET_GROW(ilp, n_slots + 1);
    const double theta2 = par.theta * par.theta;
    k_let_plan<<<grid_for(ncodes, 256), 256, 0, st>>>(let.table, ncodes, let.segs, let.seg_len, let.split, theta2, par.soft2, root, ell,
                                                        let.nit, let.blk, let.recvsz, let.sendsz);
    // exclusive scans: one fused single-block launch while the arrays are small, the multi-block look-back scan beyond
    const bool small_scans = n_slots <= 32768;
    if (small_scans) {
        LetScanJob job{};
        job.in[0] = let.nit; job.out[0] = let.item_first; job.n[0] = (int)ncodes;
        job.in[1] = let.recvsz; job.out[1] = let.recvoff; job.n[1] = (int)ncodes;
        job.in[2] = let.sendsz; job.out[2] = let.sendoff; job.n[2] = let.ipc_ok ? 0 : (int)(P * mylen);
        k_let_scans<<<3, 1024, 0, st>>>(job);
    } else {
        BH_RC(excl_scan(let.nit, (int)ncodes, let.item_first));
        BH_RC(excl_scan(let.recvsz, (int)ncodes, let.recvoff));
        if (!let.ipc_ok) BH_RC(excl_scan(let.sendsz, (int)(P * mylen), let.sendoff));
    }
    if (let.ipc_ok) BH_TRY(cudaMemsetAsync(let.sendoff, 0, ((size_t)P * mylen + 1) * sizeof(int), st));
    k_let_items<<<grid_for(ncodes, 256), 256, 0, st>>>(ncodes, let.item_first, let.blk, root.levels, ell, let.ikey, let.itype, let.iw);
    k_let_item_cnt<<<grid_for(n_slots, 256), 256, 0, st>>>(let.ikey, let.item_first + ncodes, n_slots, root.levels, let.icnt, let.iw);
    if (small_scans) {
        LetScanJob job{};
        job.in[0] = let.icnt; job.out[0] = let.iS; job.n[0] = n_slots;
        job.in[1] = let.iw; job.out[1] = let.iW; job.n[1] = n_slots;
        k_let_scans<<<2, 1024, 0, st>>>(job);
    } else {
        BH_RC(excl_scan(let.icnt, n_slots, let.iS));
        BH_RC(excl_scan(let.iw, n_slots, let.iW));
    }
    k_let_collect<<<1, 32, 0, st>>>(let.item_first, ncodes, let.iS, let.iW, let.recvoff, let.sendoff, let.split, let.dcnt, let.dcollect,
                                let.table + ncodes + P);
    ctr.kernel_launches += 5;
    BH_TRY(cudaMemcpyAsync(let.hcollect, let.dcollect, 40 * sizeof(int), cudaMemcpyDeviceToHost, st));
    BH_TRY(cudaMemcpy2DAsync(let.hhdr, sizeof(double), let.table + ncodes, sizeof(BhLetEntry), sizeof(double), (size_t)P,
                             cudaMemcpyDeviceToHost, st));
    BH_TRY(cudaStreamSynchronize(st));
    BH_TRY(cudaGetLastError());
    LET_PHASE(7);   // 6: plan + scans + host sync
    // a stray sits in a jitter cluster of its host, or a rank had more strays than its segment holds
    {
        int why = 0;
        for (int q = 0; q < P; ++q) why |= (int)let.hhdr[q];
        if (why) {
            let.fb_jitter += (why & 1) != 0; let.fb_strays += (why & 2) != 0; let.fb_cells += (why & 4) != 0;
            return BH_LET_RETRY;
        }
    }
    let.last_strays = let.hcollect[38];
    // (NOT a place to schedule an early re-homing from: this is THIS rank's count, and every rank must take the same
    // branch into the collectives of the next build — an overflow is agreed on through the all-reduced retry flags)
    let.jret_total += let.hcollect[36];
    let.returns_applied = let.hcollect[37] > 0;
    let.n_items = let.hcollect[0];
    let.M = let.hcollect[1];
    const int* roff = let.hcollect + 2;
    const int* soff = let.hcollect + 2 + 17;
    if (jit && n_own > 0) {          // the jitter replay mutated own bodies (BH.kt:145-156): keep the mutation
        BH_TRY(cudaMemcpyAsync(x + lo, let.lx, (size_t)n_own * sizeof(double), cudaMemcpyDeviceToDevice, st));
        BH_TRY(cudaMemcpyAsync(y + lo, let.ly, (size_t)n_own * sizeof(double), cudaMemcpyDeviceToDevice, st));
        let.pos_valid = false;
    }
    if (n_own > 0) {                 // ... and the mutation of own strays by the replay on THEIR host ranks
        k_let_apply_returns<<<grid_for((int64_t)P * LET_JRET, 256), 256, 0, st>>>(let.table + ncodes + P, P * LET_JRET, rank, let.stray_slot,
                                                                                 x + lo, y + lo, let.lx, let.ly);
        ctr.kernel_launches += 1;
    }
    // ---- 6. blocks to the ranks that may open them
    LET_GROW(cell, (int64_t)let.M + 1); LET_GROW(cd, (int64_t)let.M + 1); LET_GROW(sk, (int64_t)let.M + 1); LET_GROW(arrived, (int64_t)let.M + 1);
    if (!let.ipc_ok) {               // no peer memory: pack, ncclSend / ncclRecv, unpack
        LET_GROW(sendbuf, (std::max(1, soff[P]))); LET_GROW(recvbuf, (std::max(1, roff[P])));
        if (soff[P] > 0) {
            k_let_pack<<<grid_for((int64_t)P * mylen * 32, 256), 256, 0, st>>>(let.table, let.split, let.sendsz, let.sendoff, cd, sk, let.sendbuf);
            ctr.kernel_launches += 1;
        }
        rc = A.GroupStart();
        for (int q = 0; q < P && rc == bhcomm::kSuccess; ++q) {
            if (q == rank) continue;
            const int ns = soff[q + 1] - soff[q], nr = roff[q + 1] - roff[q];
            if (ns > 0) rc = A.Send(let.sendbuf + soff[q], (size_t)ns * 4, bhcomm::kFloat64, q, comm, st);
            if (nr > 0 && rc == bhcomm::kSuccess) rc = A.Recv(let.recvbuf + roff[q], (size_t)nr * 4, bhcomm::kFloat64, q, comm, st);
        }
        const int rc2 = A.GroupEnd();
        if (rc == bhcomm::kSuccess) rc = rc2;
        if (rc != bhcomm::kSuccess) return nccl_fail(rc, "ncclSend/ncclRecv(blocks)");
    }
    LET_PHASE(8);   // 7: host gap (+ pack, send/recv without peer memory)
    let.last_imported = roff[P]; let.last_sent = soff[P];
    // ---- 7. top tree + blocks + exact climb
    BhTreeView lv{};
    lv.cell = let.cell; lv.cd = let.cd; lv.sk = let.sk; lv.arrived = let.arrived; lv.n_in = let.n_items; lv.M = let.M;
    const BhLetItems it{let.ikey, let.itype, let.iS, let.iW, let.n_items};
    if (let.n_items > 0) {
        k_let_emit<<<grid_for(let.n_items, 256), 256, 0, st>>>(it, let.sk, let.arrived, root.levels, ell, let.ilp, let.dst);
        LET_PHASE(9);    // 8: top-tree skeletons
        // one warp per chunk of LET_BLK_CHUNK cells: chunk counts per code, their exclusive scan, then at most
        // M / chunk + ncodes warps (the kernel reads the real total from the scan)
        LET_GROW(nchunk, (int64_t)ncodes + 1); LET_GROW(chunkoff, (int64_t)ncodes + 2);
        k_let_block_chunks<<<grid_for(ncodes, 256), 256, 0, st>>>(let.nit, let.blk, ncodes, let.nchunk);
        BH_RC(excl_scan(let.nchunk, (int)ncodes, let.chunkoff));
        const int64_t max_chunks = (int64_t)let.M / LET_BLK_CHUNK + (int64_t)ncodes + 1;
        k_let_blocks<<<grid_for(max_chunks * 32, 256), 256, 0, st>>>(lv, let.table, ncodes, let.split, let.chunkoff, let.blk, let.dst, let.recvoff,
                                                                     let.recvbuf, cd, sk, root.half, let.peers, let.ipc_ok ? 1 : 0);
        ctr.kernel_launches += 1;
    } else LET_PHASE(9);
    LET_PHASE(10);   // 9: own + imported blocks
    k_let_climb<<<std::max(1, grid_for(let.n_items, 128)), 128, 0, st>>>(lv, root, it, let.table, root.levels, ell, let.ilp);
    LET_PHASE(11);   // 10: top-tree climb
    if (n_own > 0)
        k_let_leafpos<<<grid_for(n_own, 256), 256, 0, st>>>(n_own, let.lleaf, let.lx, let.ly, root, grid, ell, let.table, let.dst, let.blk, lv, leafpos + lo, let.dcnt);
    ctr.kernel_launches += 4;
    LET_PHASE(12);   // 11: leaf positions
    LET_PHASE(13); LET_PHASE(14);
    let.pe_armed = true;
#undef LET_PHASE
    BH_TRY(cudaEventRecord(ev[slot + 1], st));
    BH_TRY(cudaGetLastError());
    ctr.n_cells = let.M;
    let.view_valid = true;
    let.evaluations++;
    return BH_OK;
}

#undef LET_GROW
#endif  // BH_LET_ENGINE_CUH
