// bh_scene.cuh — scene generators of BodyFactory.kt on the device.
// Included by bh_engine.cu after the engine definition.  Distributional parity only: same sampling
// laws as the reference, a counter-based generator (splitmix64 of seed, body, draw) instead of
// Kotlin's XorWow stream, so any number of bodies is generated in parallel and reproducibly.
#ifndef BH_SCENE_CUH
#define BH_SCENE_CUH

namespace {

__device__ __forceinline__ double bh_rand01(uint64_t seed, uint64_t body, uint32_t draw) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (body * 8ull + draw + 1ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return (double)(z >> 11) * (1.0 / 9007199254740992.0);   // [0, 1)
}

// makeUniformRandom, BodyFactory.kt:160-177
__global__ void k_gen_uniform(int n, double w, double h, double mass, uint64_t seed, double* __restrict__ x,
                              double* __restrict__ y, double* __restrict__ vx, double* __restrict__ vy, double* __restrict__ m) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    x[i] = bh_rand01(seed, i, 0) * w;
    y[i] = bh_rand01(seed, i, 1) * h;
    vx[i] = 0.0; vy[i] = 0.0; m[i] = mass;
}

// positions of makeGalaxyDisk (BodyFactory.kt:84-116) / makeKeplerDisk (:27-41); body 0 = the centre.
// rkey[i] = bits of the body's distance from the centre (non-negative doubles sort as integers).
__global__ void k_gen_disk_positions(int n_total, bh_disk_params p, uint64_t seed, double* __restrict__ x, double* __restrict__ y,
                                     double* __restrict__ vx, double* __restrict__ vy, double* __restrict__ m,
                                     uint64_t* __restrict__ rkey) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_total) return;
    const int sats = n_total - 1;
    double px = p.x, py = p.y, mass = p.central_mass;
    if (i > 0) {
        mass = sats > 0 ? p.total_satellite_mass / sats : 0.0;
        const double u = bh_rand01(seed, i, 0);
        if (p.kepler) {   // uniform in area on [min_r, r], radius jitter (:35-37)
            const double rr = sqrt(u * (p.r * p.r - p.min_r * p.min_r) + p.min_r * p.min_r);
            const double rj = rr * (1.0 + (bh_rand01(seed, i, 1) - 0.5) * 2.0 * p.radial_jitter);
            const double ang = bh_rand01(seed, i, 2) * 2.0 * 3.141592653589793;
            px = p.x + rj * cos(ang); py = p.y + rj * sin(ang);
        } else {          // truncated exponential radius + m=2 bar (:97-113)
            const double Rd = p.radial_scale > 0.0 ? p.radial_scale : p.r / 3.0;
            const double taperR = p.bar_taper_r > 0.0 ? p.bar_taper_r : p.r * 0.6;
            const double A = exp(-(p.r - p.min_r) / Rd);
            const double R = p.min_r - Rd * log(1.0 - u * (1.0 - A));
            const double theta = bh_rand01(seed, i, 2) * 2.0 * 3.141592653589793;
            const double taper = exp(-(R / taperR) * (R / taperR));
            const double R2 = R * (1.0 + p.eps_m2 * cos(2.0 * (theta - p.phi0)) * taper);
            px = p.x + R2 * cos(theta); py = p.y + R2 * sin(theta);
        }
    }
    x[i] = px; y[i] = py; vx[i] = p.vx; vy[i] = p.vy; m[i] = mass;
    rkey[i] = (uint64_t)__double_as_longlong(hypot(px - p.x, py - p.y));
}

// circular speeds from the exact enclosed mass (BodyFactory.kt:118-147): sorted position k of a
// satellite = number of satellites at or inside its radius (the centre, r = 0, is position 0)
__global__ void k_gen_disk_velocities(int n_total, bh_disk_params p, double G, uint64_t seed, const uint32_t* __restrict__ by_radius,
                                      const double* __restrict__ x, const double* __restrict__ y, double* __restrict__ vx,
                                      double* __restrict__ vy) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_total) return;
    const int i = (int)by_radius[k];
    if (i == 0) return;
    const int sats = n_total - 1;
    const double msat = p.total_satellite_mass / sats;
    const double menc = p.central_mass + (double)k * msat;
    const double dx = x[i] - p.x, dy = y[i] - p.y;
    const double R = fmax(1e-6, hypot(dx, dy));
    const double vcirc = sqrt(G * menc / R);
    const double v = vcirc * (1.0 + (bh_rand01(seed, i, 3) - 0.5) * 2.0 * p.speed_jitter);
    const double tx = p.clockwise ? dy / R : -dy / R, ty = p.clockwise ? -dx / R : dx / R;
    double vx0 = tx * v, vy0 = ty * v;
    if (!p.kepler && p.radial_jitter > 0.0) {
        const double vr = (bh_rand01(seed, i, 4) - 0.5) * 2.0 * p.radial_jitter * vcirc;
        vx0 += dx / R * vr; vy0 += dy / R * vr;
    }
    vx[i] = vx0 + p.vx; vy[i] = vy0 + p.vy;
}

__global__ void k_iota_from(int* __restrict__ a, int first, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[first + i] = first + i;
}

}  // namespace

// room for `extra` more bodies behind the current ones, keeping the state (x, y, vx, vy, m, perm)
inline int bh_engine::grow_keep(int64_t extra) {
    const int64_t nn = n + extra;
    if (nn <= cap && scratch) return BH_OK;
    double* keep[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    int* keep_perm = nullptr;
    double** src[5] = {&x, &y, &vx, &vy, &m};
    if (n > 0) {
        cudaError_t ce = cudaSuccess;
        for (int k = 0; k < 5 && ce == cudaSuccess; ++k) {
            ce = dev_alloc(&keep[k], (size_t)n);
            if (ce == cudaSuccess) ce = cudaMemcpyAsync(keep[k], *src[k], (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, st);
        }
        if (ce == cudaSuccess) ce = dev_alloc(&keep_perm, (size_t)n);
        if (ce == cudaSuccess) ce = cudaMemcpyAsync(keep_perm, perm, (size_t)n * sizeof(int), cudaMemcpyDeviceToDevice, st);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
        if (ce != cudaSuccess) {           // nothing was changed yet: release the copies and report
            for (auto& q : keep) cudaFree(q);
            cudaFree(keep_perm);
            return cuda_fail(ce, "grow_keep");
        }
    }
    const int64_t n_keep = n;
    int rc = ensure_bodies(std::max<int64_t>(nn, cap + cap / 2));
    if (rc == BH_OK && n_keep > 0) {
        cudaError_t ce = cudaSuccess;
        for (int k = 0; k < 5 && ce == cudaSuccess; ++k)
            ce = cudaMemcpyAsync(*src[k], keep[k], (size_t)n_keep * sizeof(double), cudaMemcpyDeviceToDevice, st);
        if (ce == cudaSuccess) ce = cudaMemcpyAsync(perm, keep_perm, (size_t)n_keep * sizeof(int), cudaMemcpyDeviceToDevice, st);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
        if (ce != cudaSuccess) rc = cuda_fail(ce, "grow_keep");
    }
    for (auto& q : keep) cudaFree(q);
    cudaFree(keep_perm);
    return rc;
}

// after new bodies were written to the home slots [n, n + added): list order = old + new
inline int bh_engine::finish_append(int64_t added) {
    if (added <= 0) return BH_OK;
    k_iota_from<<<grid_for(added, 256), 256, 0, st>>>(perm, (int)n, (int)added);
    n += added;
    BH_TRY(cudaMemsetAsync(ax, 0, (size_t)n * sizeof(double), st));
    BH_TRY(cudaMemsetAsync(ay, 0, (size_t)n * sizeof(double), st));
    BH_TRY(cudaMemsetAsync(dflags, 0, HF_COUNT * sizeof(int), st));
    BH_TRY(cudaStreamSynchronize(st));
    BH_TRY(cudaGetLastError());
    ctr.kernel_launches += 1;
    origin_identity = true;
    perm_identity = false;            // (kept simple: the combined permutation is treated as general)
    rehome_due = true;
    tree_valid = false; heavies_valid = false; acc_valid = false;
    return BH_OK;
}

#endif  // BH_SCENE_CUH
