// bh_merge.cuh — the merge ("devour") rule of PhysicsEngine.step(), BarnesHutAlg.kt:463-532, on
// the device.  Included by bh_engine.cu after the engine definition.
//
// Reference semantics (BH.kt:463-532): for i ascending over the `bodies` list, if
// bodies[i].m > mergeMaxMass (strict), every j != i with dx*dx + dy*dy < mergeMinDist^2 (strict,
// f64) is a victim; victims are removed in DESCENDING index order, each adding its mass only
// (no momentum) to bodies[i]; later heavy bodies see the shrunk list, so a body inside the
// radius of two heavy bodies goes to the first one, and a heavy body can be eaten by an earlier
// one.  List order of the survivors is preserved (removeAt).
//
// Device form: positions do not change during the rule, so the geometric test is done ONCE for
// all (heavy, body) pairs (k_merge_candidates); the candidates are radix-sorted by
// (heavy rank, descending user index) and applied by a single thread in exactly the
// reference's order (the f64 mass sums are order-dependent); the survivors are then compacted
// with two prefix sums (home order and user order).  Everything is replicated on every rank of a
// multi-GPU run (positions and masses are), only the removal needs all velocities.
#ifndef BH_MERGE_CUH
#define BH_MERGE_CUH

inline int bh_engine::excl_scan(const int* in, int nn, int* out) {
    uint32_t* ticket = sort_scratch();
    uint32_t* status = ticket + 1;
    BH_TRY(cudaMemsetAsync(ticket, 0, (1 + scan_tiles(nn)) * sizeof(uint32_t), st));
    k_excl_scan<<<std::max(1, grid_for(nn, SCAN_TILE)), SCAN_THREADS, 0, st>>>(in, nn, out, ticket, status);
    ctr.kernel_launches += 1;
    return BH_OK;
}

inline int bh_engine::merge_rule() {
    if (!merge_enabled()) return BH_OK;                       // BH.kt:465
    BH_TRY(cudaEventRecord(ev[14], st));
    const int nn = (int)n;
    const int g = grid_for(nn, 256);
    bool inv_valid = false;

    // ---- the heavy bodies, ascending user index (cached: masses only change here / in set_bodies)
    if (!heavies_valid || heavies_max_mass != par.merge_max_mass) {
        k_invert_perm<<<g, 256, 0, st>>>(perm, nn, inv);
        inv_valid = true;
        k_merge_flag_heavy<<<g, 256, 0, st>>>(m, inv, nn, par.merge_max_mass, iscr0);
        ctr.kernel_launches += 2;
        BH_RC(excl_scan(iscr0, nn, iscr1));
        BH_TRY(cudaMemcpyAsync(hflags + HF_N_CAND, iscr1 + nn, sizeof(int), cudaMemcpyDeviceToHost, st));
        BH_TRY(cudaStreamSynchronize(st));
        n_heavy = hflags[HF_N_CAND];
        if (n_heavy > heavy_cap) {
            dev_free(heavy);
            heavy_cap = std::max<int64_t>(n_heavy + n_heavy / 4, 64);
            BH_TRY(dev_alloc(&heavy, (size_t)heavy_cap));
        }
        if (n_heavy > 0) {
            k_merge_list_heavy<<<g, 256, 0, st>>>(iscr0, iscr1, inv, nn, heavy);
            ctr.kernel_launches += 1;
        }
        heavies_valid = true;
        heavies_max_mass = par.merge_max_mass;
    }
    if (n_heavy == 0) return BH_OK;

    // ---- candidate victims of every heavy body (one pass over the bodies)
    const double minD2 = par.merge_min_dist * par.merge_min_dist;   // BH.kt:468
    unsigned int* d_count = reinterpret_cast<unsigned int*>(dflags + HF_N_CAND);
    BH_TRY(cudaMemsetAsync(dflags + HF_N_DEAD, 0, 2 * sizeof(int), st));
    uint32_t* cand_home = reinterpret_cast<uint32_t*>(iscr1);
    tree_valid = false;   // keys_a/keys_b are reused below; bh_get_tree rebuilds the (identical) tree on demand
    k_merge_candidates<<<grid_for(nn, MERGE_TILE), MERGE_TILE, 0, st>>>(x, y, perm, nn, heavy, n_heavy, minD2, keys_a, cand_home,
                                                                         nn, d_count);
    ctr.kernel_launches += 1;
    BH_TRY(cudaMemcpyAsync(hflags + HF_N_CAND, dflags + HF_N_CAND, sizeof(int), cudaMemcpyDeviceToHost, st));
    BH_TRY(cudaStreamSynchronize(st));
    BH_TRY(cudaGetLastError());
    const unsigned int n_cand = (unsigned int)hflags[HF_N_CAND];
    if (n_cand == 0) return merge_done();
    // More (heavy, victim) pairs than bodies (a large part of the system is heavy): the per-body buffers are
    // too small, so the candidates are listed again into buffers of the reported size.  The SET of pairs is
    // deterministic (positions do not change during the rule); their order is fixed by the sort below.
    struct Big {
        uint64_t *ka = nullptr, *kb = nullptr; uint32_t *va = nullptr, *vb = nullptr, *cand = nullptr, *scr = nullptr;
        ~Big() { cudaFree(ka); cudaFree(kb); cudaFree(va); cudaFree(vb); cudaFree(cand); cudaFree(scr); }
    } big;
    uint64_t *ck_a = keys_a, *ck_b = keys_b;
    uint32_t *cv_a = vals_a, *cv_b = vals_b, *cscr = sort_scratch();
    int kb = 1;
    while ((1 << kb) < n_heavy) ++kb;
    if (n_cand > (unsigned int)nn) {
        if (n_cand >= (1u << 30)) return fail(BH_E_ARG, "merge rule: more than 2^30 (heavy, victim) candidate pairs");
        const size_t c = (size_t)n_cand;
        BH_TRY(dev_alloc(&big.ka, c)); BH_TRY(dev_alloc(&big.kb, c)); BH_TRY(dev_alloc(&big.va, c)); BH_TRY(dev_alloc(&big.vb, c));
        BH_TRY(dev_alloc(&big.cand, c));
        BH_TRY(dev_alloc(&big.scr, bhsort::sort_scratch_words((int64_t)c, (32 + kb + 7) / 8)));
        BH_TRY(cudaMemsetAsync(dflags + HF_N_CAND, 0, sizeof(int), st));
        k_merge_candidates<<<grid_for(nn, MERGE_TILE), MERGE_TILE, 0, st>>>(x, y, perm, nn, heavy, n_heavy, minD2, big.ka, big.cand,
                                                                             (int)n_cand, d_count);
        ctr.kernel_launches += 1;
        ck_a = big.ka; ck_b = big.kb; cv_a = big.va; cv_b = big.vb; cscr = big.scr; cand_home = big.cand;
    }

    // ---- order: heavy rank ascending, victim user index descending; apply sequentially
    const int where = bhsort::onesweep_sort(ck_a, cv_a, ck_b, cv_b, (int64_t)n_cand, 32 + kb, cscr, st, num_sms);
    const uint64_t* skeys = where ? ck_b : ck_a;
    const uint32_t* sslot = where ? cv_b : cv_a;
    ctr.kernel_launches += 2 + (32 + kb + 7) / 8;
    BH_TRY(cudaMemsetAsync(dead, 0, (size_t)nn * sizeof(int), st));
    acc_valid = false;   // masses change
    k_merge_apply<<<1, 32, 0, st>>>(skeys, sslot, cand_home, (int)n_cand, heavy, m, dead, dflags + HF_N_DEAD);
    ctr.kernel_launches += 1;
    BH_TRY(cudaMemcpyAsync(hflags + HF_N_DEAD, dflags + HF_N_DEAD, sizeof(int), cudaMemcpyDeviceToHost, st));
    BH_TRY(cudaStreamSynchronize(st));
    BH_TRY(cudaGetLastError());
    const int n_dead = hflags[HF_N_DEAD];
    if (n_dead == 0) return merge_done();

    // ---- bodies.removeAt(j): stable compaction of the state (home order) and of the user indices
    BH_RC(sync_velocities());                                  // the survivors are re-sliced: every rank needs all velocities
    if (!inv_valid) { k_invert_perm<<<g, 256, 0, st>>>(perm, nn, inv); ctr.kernel_launches += 1; }
    if (origin_identity) { k_iota<<<g, 256, 0, st>>>(origin, nn); ctr.kernel_launches += 1; }
    int* alive_home = iscr0;
    int* alive_user = iscr1;
    k_merge_alive<<<g, 256, 0, st>>>(dead, inv, nn, alive_home, alive_user);
    int* new_home = leafpos;                                   // the tree is dropped (BH.kt:526): leafpos and S are free
    int* new_user = S;
    BH_RC(excl_scan(alive_home, nn, new_home));
    BH_RC(excl_scan(alive_user, nn, new_user));
    // x,y,vx,vy,m -> (dtmp, ax, ay, keys_a, keys_b); perm -> itmp; origin -> vals_a.  ax/ay are recomputed by
    // the next evaluation (the reference never reads them between steps either).
    double* x2 = dtmp;
    double* y2 = ax;
    double* vx2 = ay;
    double* vy2 = reinterpret_cast<double*>(keys_a);
    double* m2 = reinterpret_cast<double*>(keys_b);
    k_merge_compact<<<g, 256, 0, st>>>(dead, new_home, new_user, nn, x, y, vx, vy, m, perm, x2, y2, vx2, vy2, m2, itmp);
    k_merge_compact_origin<<<g, 256, 0, st>>>(dead, inv, new_user, nn, origin, reinterpret_cast<int*>(vals_a));
    ctr.kernel_launches += 3;
    const int64_t n2 = n - n_dead;
    const size_t b2 = (size_t)n2 * sizeof(double);
    std::swap(x, dtmp);
    std::swap(y, ax);
    std::swap(vx, ay);
    BH_TRY(cudaMemcpyAsync(vy, vy2, b2, cudaMemcpyDeviceToDevice, st));
    BH_TRY(cudaMemcpyAsync(m, m2, b2, cudaMemcpyDeviceToDevice, st));
    std::swap(perm, itmp);
    BH_TRY(cudaMemcpyAsync(origin, vals_a, (size_t)n2 * sizeof(int), cudaMemcpyDeviceToDevice, st));
    BH_TRY(cudaStreamSynchronize(st));
    BH_TRY(cudaGetLastError());
    n = n2;
    origin_identity = false;
    perm_identity = false;
    heavies_valid = false;            // home slots moved
    tree_valid = false;               // lastTree = null, BH.kt:526
    ctr.total_merged += n_dead;
    return merge_done();
}

inline int bh_engine::merge_done() {
    BH_TRY(cudaEventRecord(ev[15], st));
    cur->has_merge = true;     // folded into ms_merge when the step's timers are collected
    return BH_OK;
}

#endif  // BH_MERGE_CUH
