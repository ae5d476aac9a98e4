// bh_ref.cpp — ORACLE: literal double-precision C++ restatement of the reference's
// CPU physics step (/root/reference/src/main/kotlin/BarnesHutAlg.kt).
//
// THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load it.  The product
// (barnes-hut-n-body_b200/) never links, loads or calls anything under oracle/.
//
// PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures
//  and no JVM/Kotlin toolchain exists in the build container, so
// this restatement cannot be checked against a running reference.  It is pinned only
// by hand-derivable known-answer cases (tests/test_oracle_known_answers.py) and is
// kept literal enough to diff by eye against BarnesHutAlg.kt: same recursion, same
// expression order, same jitter, same out-of-box drop, same merge rule.  The path
// uses only IEEE binary64 + - * / sqrt and comparisons, which a strict-FP JVM
// (JDK >= 17) and g++ -O2 -ffp-contract=off evaluate identically.
//
// A second, independent reading of the same source (oracle/bh_ref_second.py: Python objects and
// recursion) must agree with this one bit for bit: tests/test_oracle_second_reading.py.
//
// Every function cites the BarnesHutAlg.kt lines it follows ("BH.kt:a-b").
#include "../include/bh_engine.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <new>
#include <string>
#include <thread>
#include <vector>

namespace {

// ---- BH.kt:21-25  data class Body -------------------------------------------------
struct Body { double x, y, vx, vy, m; };

// ---- BH.kt:33-41  class Acc ---------------------------------------------------------
struct Acc {
    double fx = 0.0, fy = 0.0;
    void reset() { fx = 0.0; fy = 0.0; }
};

// ---- BH.kt:53-82  data class Quad ---------------------------------------------------
struct Quad {
    double cx, cy, h;
    // BH.kt:61-62 — half-open [cx-h, cx+h) x [cy-h, cy+h)
    bool contains(const Body& b) const {
        return b.x >= cx - h && b.x < cx + h && b.y >= cy - h && b.y < cy + h;
    }
    // BH.kt:73-81 — 0 NW, 1 NE, 2 SW, 3 SE; child half-side h/2
    Quad child(int which) const {
        const double hh = h / 2.0;
        switch (which) {
            case 0:  return Quad{cx - hh, cy - hh, hh};
            case 1:  return Quad{cx + hh, cy - hh, hh};
            case 2:  return Quad{cx - hh, cy + hh, hh};
            default: return Quad{cx + hh, cy + hh, hh};
        }
    }
};

struct Params { double G, soft2; };

// per-walk observer counters (do not perturb the arithmetic)
struct WalkStat { int64_t interactions = 0, opened = 0; };

struct BHTree;

// Bump allocator standing in for the JVM's TLAB allocation of BHTree/Quad objects
// (BH.kt:159-166 allocates 4 nodes per subdivide); the whole tree is garbage after
// each build, like in the reference.
struct Arena {
    static constexpr size_t CHUNK = 1 << 16;  // nodes per chunk
    std::vector<BHTree*> chunks;
    size_t used = CHUNK;
    ~Arena();
    BHTree* alloc4();
    void clear();
};

// ---- BH.kt:95-275  class BHTree ------------------------------------------------------
struct BHTree {
    Quad    quad;
    Body*   body;      // BH.kt:97
    BHTree* children;  // BH.kt:100 — array of 4 or null (leaf)
    double  mass, comX, comY;  // BH.kt:103-109

    void init(const Quad& q) { quad = q; body = nullptr; children = nullptr; mass = 0.0; comX = 0.0; comY = 0.0; }
    bool isLeaf() const { return children == nullptr; }  // BH.kt:112

    // BH.kt:125-137
    void insert(Body* b, Arena& arena) {
        if (!quad.contains(*b)) return;
        if (body == nullptr && isLeaf()) { body = b; return; }
        if (isLeaf()) subdivide(arena);
        if (body != nullptr) {
            Body* existing = body;
            body = nullptr;
            insertIntoChild(existing, arena);
        }
        insertIntoChild(b, arena);
    }

    // BH.kt:145-156 — incl. the h < 1e-3 jitter that MUTATES the body
    void insertIntoChild(Body* b, Arena& arena) {
        if (quad.h < 1e-3) {
            const double eps = 1e-3;
            uint64_t xb, yb;
            std::memcpy(&xb, &b->x, 8);
            b->x += ((xb & 1ull) == 0ull) ? +eps : -eps;
            std::memcpy(&yb, &b->y, 8);
            b->y += ((yb & 1ull) == 0ull) ? -eps : +eps;
        }
        const int ix = (b->x < quad.cx) ? 0 : 1;
        const int iy = (b->y < quad.cy) ? 0 : 2;
        children[ix + iy].insert(b, arena);
    }

    // BH.kt:159-166
    void subdivide(Arena& arena) {
        children = arena.alloc4();
        for (int k = 0; k < 4; ++k) children[k].init(quad.child(k));
    }

    // BH.kt:173-202
    void computeMass() {
        if (isLeaf()) {
            if (body != nullptr) { mass = body->m; comX = body->x; comY = body->y; }
            else                 { mass = 0.0;     comX = quad.cx; comY = quad.cy; }
        } else {
            double mSum = 0.0, cx = 0.0, cy = 0.0;
            for (int k = 0; k < 4; ++k) {  // BH.kt:189-192, children 0,1,2,3
                BHTree& c = children[k];
                c.computeMass();
                if (c.mass > 0.0) { mSum += c.mass; cx += c.comX * c.mass; cy += c.comY * c.mass; }
            }
            mass = mSum;
            if (mSum > 0.0) { comX = cx / mSum; comY = cy / mSum; }
            else            { comX = quad.cx;   comY = quad.cy; }
        }
    }

    // BH.kt:250-259
    static inline void pointForceAcc(const Body& b, double px, double py, double m, Acc& acc, const Params& P) {
        const double dx = px - b.x;
        const double dy = py - b.y;
        const double r2 = dx * dx + dy * dy + P.soft2;
        const double invR = 1.0 / std::sqrt(r2);
        const double invR2 = 1.0 / r2;
        const double f = P.G * b.m * m * invR2;
        acc.fx += f * dx * invR;
        acc.fy += f * dy * invR;
    }

    // BH.kt:215-239
    void accumulateForce(const Body* b, double theta2, Acc& acc, const Params& P, WalkStat& st) const {
        if (mass == 0.0) return;
        if (isLeaf()) {
            const Body* single = body;
            if (single == nullptr || single == b) return;  // identity, BH.kt:219
            pointForceAcc(*b, comX, comY, mass, acc, P);
            st.interactions++;
            return;
        }
        const double dx = comX - b->x;
        const double dy = comY - b->y;
        const double dist2 = dx * dx + dy * dy + P.soft2;  // softening inside the criterion
        const double side = quad.h * 2.0;
        const double s2 = side * side;
        if (s2 < theta2 * dist2) {
            pointForceAcc(*b, comX, comY, mass, acc, P);
            st.interactions++;
        } else {
            st.opened++;
            children[0].accumulateForce(b, theta2, acc, P, st);
            children[1].accumulateForce(b, theta2, acc, P, st);
            children[2].accumulateForce(b, theta2, acc, P, st);
            children[3].accumulateForce(b, theta2, acc, P, st);
        }
    }
};

Arena::~Arena() { for (auto* c : chunks) std::free(c); }
BHTree* Arena::alloc4() {
    if (used + 4 > CHUNK) {
        void* p = std::malloc(sizeof(BHTree) * CHUNK);
        if (!p) throw std::bad_alloc();
        chunks.push_back(static_cast<BHTree*>(p));
        used = 0;
    }
    BHTree* r = chunks.back() + used;
    used += 4;
    return r;
}
void Arena::clear() { for (auto* c : chunks) std::free(c); chunks.clear(); used = CHUNK; }

double now_ms() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

}  // namespace

// ---- BH.kt:287-533  class PhysicsEngine ---------------------------------------------
struct bh_engine {
    bh_config cfg{};
    bh_params par{};
    int cores = 1;                 // BH.kt:292
    std::vector<Body> bodies;      // BH.kt:295
    std::vector<double> ax, ay;    // BH.kt:298-301
    Arena arena;                   // storage of lastTree
    BHTree rootStore{};
    BHTree* lastTree = nullptr;    // BH.kt:304
    std::vector<int32_t> cntI, cntO;
    std::vector<int32_t> origin;   // observer: index at bh_set_bodies time of each surviving body
    bh_counters ctr{};
    std::string err;
    std::vector<float> snapXY, snapM;   // bh_request_positions_f32 snapshot
    int64_t snapN = -1;
    // host-staged multi-process protocol of include/bh_engine.h (no reference counterpart): this
    // rank evaluates and integrates only the list positions [lo, hi) of bh_slice_bounds
    int rank = 0, world = 1, phase = 0;
    void mySlice(int64_t* lo, int64_t* hi) const { bh_slice_bounds((int64_t)bodies.size(), world, rank, lo, hi); }

    // BH.kt:359-366
    BHTree* buildTree() {
        const double t0 = now_ms();
        arena.clear();
        rootStore.init(Quad{par.root_cx, par.root_cy, par.root_half});
        BHTree* root = &rootStore;
        // observer (not part of the reference): bounding box of all bodies, as the build finds them
        double lox = std::nan(""), hix = std::nan(""), loy = std::nan(""), hiy = std::nan("");
        for (const auto& b : bodies) {
            if (b.x != b.x || b.y != b.y) continue;
            if (!(lox <= b.x)) lox = b.x;
            if (!(hix >= b.x)) hix = b.x;
            if (!(loy <= b.y)) loy = b.y;
            if (!(hiy >= b.y)) hiy = b.y;
        }
        ctr.bbox_min_x = lox; ctr.bbox_max_x = hix; ctr.bbox_min_y = loy; ctr.bbox_max_y = hiy;
        for (auto& b : bodies) root->insert(&b, arena);
        root->computeMass();
        ctr.ms_build += now_ms() - t0;
        return root;
    }

    // BH.kt:374-395 — min(cores, n) workers pulling indices from one atomic counter
    void computeAccelerations(const BHTree* root, int64_t first = 0, int64_t last = -1) {
        const double t0 = now_ms();
        const int64_t n = last < 0 ? (int64_t)bodies.size() : last;
        const int workers = (int)std::min<int64_t>(cores, std::max<int64_t>(n, 1));
        const double theta2 = par.theta * par.theta;
        const Params P{par.G, par.soft2};
        const bool keep = (cfg.flags & BH_FLAG_BODY_COUNTS) != 0;
        if (keep) { cntI.assign(bodies.size(), 0); cntO.assign(bodies.size(), 0); }
        std::atomic<int64_t> next{first};
        std::atomic<int64_t> totI{0}, totO{0};
        auto work = [&]() {
            Acc acc;
            WalkStat tot;
            for (;;) {
                const int64_t i = next.fetch_add(1, std::memory_order_relaxed);
                if (i >= n) break;
                const Body* b = &bodies[i];
                acc.reset();
                WalkStat st;
                root->accumulateForce(b, theta2, acc, P, st);
                ax[i] = acc.fx / b->m;   // BH.kt:390-391 (0/0 = NaN for m == 0)
                ay[i] = acc.fy / b->m;
                tot.interactions += st.interactions;
                tot.opened += st.opened;
                if (keep) { cntI[i] = (int32_t)st.interactions; cntO[i] = (int32_t)st.opened; }
            }
            totI += tot.interactions;
            totO += tot.opened;
        };
        if (workers <= 1) work();
        else {
            std::vector<std::thread> th;
            th.reserve(workers);
            for (int w = 0; w < workers; ++w) th.emplace_back(work);
            for (auto& t : th) t.join();
        }
        ctr.interactions = totI; ctr.opened = totO; ctr.exact_retests = 0;
        ctr.total_interactions += totI; ctr.total_opened += totO; ctr.total_evaluations++;
        ctr.ms_walk += now_ms() - t0;
    }

    // BH.kt:405-439, in the three phases of the ABI (world == 1: the slice is the whole list)
    void stepBegin() {   // :407-422
        int64_t lo, hi;
        mySlice(&lo, &hi);
        BHTree* root = buildTree();
        computeAccelerations(root, lo, hi);
        const double t0 = now_ms();
        const double dt = par.dt;
        const double dtHalf = dt * 0.5;
        for (int64_t i = lo; i < hi; ++i) { bodies[i].vx += ax[i] * dtHalf; bodies[i].vy += ay[i] * dtHalf; }
        for (int64_t i = lo; i < hi; ++i) { Body& b = bodies[i]; b.x += b.vx * dt; b.y += b.vy * dt; }
        ctr.ms_integrate += now_ms() - t0;
        phase = 1;
    }
    void stepEnd() {     // :425-435
        int64_t lo, hi;
        mySlice(&lo, &hi);
        BHTree* root = buildTree();
        computeAccelerations(root, lo, hi);
        const double t0 = now_ms();
        const double dtHalf = par.dt * 0.5;
        for (int64_t i = lo; i < hi; ++i) { bodies[i].vx += ax[i] * dtHalf; bodies[i].vy += ay[i] * dtHalf; }
        ctr.ms_integrate += now_ms() - t0;
        lastTree = root;
        phase = 2;
    }
    void stepFinish() {  // :438
        const double t0 = now_ms();
        mergeCloseBodiesIfNeeded();
        ctr.ms_merge += now_ms() - t0;
        ctr.total_steps++;
        phase = 0;
    }
    void step() { stepBegin(); stepEnd(); stepFinish(); }

    // data-class equals() used by bodies.indexOf(bi), BH.kt:522 (structural, bitwise on doubles)
    static bool structEq(const Body& a, const Body& b) { return std::memcmp(&a, &b, sizeof(Body)) == 0; }

    // BH.kt:463-532
    void mergeCloseBodiesIfNeeded() {
        if (par.merge_min_dist <= 0.0 || bodies.size() <= 1) return;
        const double minD2 = par.merge_min_dist * par.merge_min_dist;
        size_t i = 0;
        while (i < bodies.size()) {
            if (bodies[i].m > par.merge_max_mass) {
                const size_t n = bodies.size();
                if (n > 1) {
                    // BH.kt:478-510: chunk-parallel scan whose chunks are concatenated in
                    // index order => the victim list is simply ascending in j.
                    std::vector<size_t> victims;
                    const Body bi = bodies[i];
                    for (size_t j = 0; j < n; ++j) {
                        if (j != i) {
                            const double dx = bodies[j].x - bi.x;
                            const double dy = bodies[j].y - bi.y;
                            if (dx * dx + dy * dy < minD2) victims.push_back(j);
                        }
                    }
                    if (!victims.empty()) {
                        size_t cur = i;  // where `bi` (by identity) currently lives
                        for (size_t k = victims.size(); k-- > 0;) {  // sortedDescending, BH.kt:514
                            const size_t j = victims[k];
                            if (j >= bodies.size()) continue;
                            if (j == cur) continue;                  // bodies[j] === bi
                            bodies[cur].m += bodies[j].m;            // mass only, BH.kt:518
                            bodies.erase(bodies.begin() + (std::ptrdiff_t)j);
                            origin.erase(origin.begin() + (std::ptrdiff_t)j);
                            if (j < cur) --cur;
                            ctr.total_merged++;
                        }
                        // BH.kt:522-523: indexOf(bi) is structural equality, first match
                        size_t newIndex = cur;
                        for (size_t k = 0; k < cur; ++k) if (structEq(bodies[k], bodies[cur])) { newIndex = k; break; }
                        i = newIndex;
                        lastTree = nullptr;  // BH.kt:526
                    }
                }
            }
            ++i;
        }
    }

    // BH.kt:329-332
    BHTree* getTreeForDebug() {
        if (lastTree) return lastTree;
        lastTree = buildTree();
        return lastTree;
    }
};

namespace {
thread_local std::string g_create_err;

int fail(bh_engine* e, int code, const char* msg) {
    if (e) e->err = msg; else g_create_err = msg;
    return code;
}

// BH.kt:265-274 preorder over all cells
void visitAll(const BHTree* t, const Body* base, int64_t cap, int64_t& k, double* cx, double* cy, double* h,
              double* mass, double* comx, double* comy, int32_t* body) {
    if (k < cap) {
        if (cx) cx[k] = t->quad.cx;
        if (cy) cy[k] = t->quad.cy;
        if (h) h[k] = t->quad.h;
        if (mass) mass[k] = t->mass;
        if (comx) comx[k] = t->comX;
        if (comy) comy[k] = t->comY;
        if (body) body[k] = t->children ? -2 : (t->body ? (int32_t)(t->body - base) : -1);
    }
    ++k;
    if (t->children) for (int c = 0; c < 4; ++c) visitAll(&t->children[c], base, cap, k, cx, cy, h, mass, comx, comy, body);
}

void leafPaths(const BHTree* t, const Body* base, int depth, uint64_t path, int32_t* depthOut, uint64_t* pathOut,
               int64_t& nInternal, int& maxDepth) {
    if (t->children) {
        ++nInternal;
        for (int c = 0; c < 4; ++c)
            leafPaths(&t->children[c], base, depth + 1, (path << 2) | (uint64_t)c, depthOut, pathOut, nInternal, maxDepth);
    } else if (t->body) {
        const int64_t i = t->body - base;
        if (depthOut) depthOut[i] = depth;
        if (pathOut) pathOut[i] = path;
        if (depth > maxDepth) maxDepth = depth;
    }
}
}  // namespace

extern "C" {

int bh_abi_version(void) { return BH_ABI_VERSION; }
const char* bh_backend_name(void) { return "reference-port"; }

int bh_default_params(int32_t w, int32_t h, bh_params* p) {
    if (!p) return BH_E_ARG;
    p->G = 80.0; p->dt = 0.005; p->theta = 0.30; p->soft2 = 1.0 * 1.0;   // Config.kt:11,14,23,17,20
    p->root_cx = w / 2.0; p->root_cy = h / 2.0;                          // BH.kt:361
    p->root_half = std::max(w, h) / 2.0 + 2.0;                           // BH.kt:360
    p->merge_max_mass = 4000.0; p->merge_min_dist = 8.0;                 // BH.kt:315,321; Config.kt:35
    return BH_OK;
}

int bh_create(const bh_config* cfg, bh_engine** out) {
    if (!out) return fail(nullptr, BH_E_ARG, "bh_create: out is NULL");
    bh_engine* e = new (std::nothrow) bh_engine();
    if (!e) return fail(nullptr, BH_E_OOM, "bh_create: out of memory");
    if (cfg) std::memcpy(&e->cfg, cfg, std::min<size_t>(sizeof(bh_config), cfg->struct_size > 0 ? (size_t)cfg->struct_size : sizeof(bh_config)));
    int hc = (int)std::thread::hardware_concurrency();
    if (hc < 1) hc = 1;
    e->cores = e->cfg.threads > 0 ? e->cfg.threads : hc;
    bh_default_params(2400, 800, &e->par);  // Config.kt:5,8
    *out = e;
    return BH_OK;
}

void bh_destroy(bh_engine* e) { delete e; }

const char* bh_last_error(const bh_engine* e) { return e ? e->err.c_str() : g_create_err.c_str(); }

int bh_set_params(bh_engine* e, const bh_params* p) {
    if (!e || !p) return fail(e, BH_E_ARG, "bh_set_params: NULL");
    e->par = *p;
    return BH_OK;
}
int bh_get_params(const bh_engine* e, bh_params* p) {
    if (!e || !p) return BH_E_ARG;
    *p = e->par;
    return BH_OK;
}

// BH.kt:342-349
int bh_set_bodies(bh_engine* e, int64_t n, const double* x, const double* y, const double* vx, const double* vy, const double* m) {
    if (!e || n < 0 || (n > 0 && (!x || !y || !vx || !vy || !m))) return fail(e, BH_E_ARG, "bh_set_bodies: bad arguments");
    try {
        e->bodies.resize((size_t)n);
        for (int64_t i = 0; i < n; ++i) e->bodies[i] = Body{x[i], y[i], vx[i], vy[i], m[i]};
        if ((int64_t)e->ax.size() != n) { e->ax.assign((size_t)n, 0.0); e->ay.assign((size_t)n, 0.0); }
        e->origin.resize((size_t)n);
        for (int64_t i = 0; i < n; ++i) e->origin[i] = (int32_t)i;
    } catch (const std::bad_alloc&) { return fail(e, BH_E_OOM, "bh_set_bodies: out of memory"); }
    e->lastTree = nullptr;
    return BH_OK;
}

int64_t bh_num_bodies(const bh_engine* e) { return e ? (int64_t)e->bodies.size() : 0; }

int bh_get_bodies(bh_engine* e, int64_t cap, double* x, double* y, double* vx, double* vy, double* m, int64_t* n_out) {
    if (!e) return BH_E_ARG;
    const int64_t n = (int64_t)e->bodies.size();
    if (n_out) *n_out = n;
    if (cap < n) return fail(e, BH_E_ARG, "bh_get_bodies: capacity too small");
    for (int64_t i = 0; i < n; ++i) {
        const Body& b = e->bodies[i];
        if (x) x[i] = b.x;
        if (y) y[i] = b.y;
        if (vx) vx[i] = b.vx;
        if (vy) vy[i] = b.vy;
        if (m) m[i] = b.m;
    }
    return BH_OK;
}

int bh_get_origin(bh_engine* e, int64_t cap, int32_t* origin, int64_t* n_out) {
    if (!e) return BH_E_ARG;
    const int64_t n = (int64_t)e->bodies.size();
    if (n_out) *n_out = n;
    if (cap < n) return fail(e, BH_E_ARG, "bh_get_origin: capacity too small");
    if (origin) std::memcpy(origin, e->origin.data(), (size_t)n * sizeof(int32_t));
    return BH_OK;
}

int bh_rebase_origin(bh_engine* e) {
    if (!e) return BH_E_ARG;
    for (size_t i = 0; i < e->origin.size(); ++i) e->origin[i] = (int32_t)i;
    return BH_OK;
}

int bh_get_positions_f32(bh_engine* e, int64_t cap, float* xy, float* m, int64_t* n_out) {
    if (!e) return BH_E_ARG;
    const int64_t n = (int64_t)e->bodies.size();
    if (n_out) *n_out = n;
    if (cap < n) return fail(e, BH_E_ARG, "bh_get_positions_f32: capacity too small");
    for (int64_t i = 0; i < n; ++i) {
        if (xy) { xy[2 * i] = (float)e->bodies[i].x; xy[2 * i + 1] = (float)e->bodies[i].y; }
        if (m) m[i] = (float)e->bodies[i].m;
    }
    return BH_OK;
}

// The device scene generators have no oracle twin: their parity is distributional and is checked
// against the numpy generators of scenes.py (tests/test_gpu_parity.py).
int bh_default_disk_params(int32_t w, int32_t h, bh_disk_params* p) {
    if (!p) return BH_E_ARG;
    std::memset(p, 0, sizeof(*p));
    p->x = w * 0.5; p->y = h * 0.5; p->r = 200.0; p->min_r = 8.0;
    p->central_mass = 50000.0; p->total_satellite_mass = 5000.0;
    p->eps_m2 = 0.03; p->speed_jitter = 0.01; p->clockwise = 1;
    return BH_OK;
}
int bh_append_disk(bh_engine* e, int64_t, const bh_disk_params*, uint64_t) { return fail(e, BH_E_UNSUPPORTED, "bh_append_disk: device generator (use scenes.py with the reference port)"); }
int bh_append_uniform_random(bh_engine* e, int64_t, double, int32_t, int32_t, uint64_t) { return fail(e, BH_E_UNSUPPORTED, "bh_append_uniform_random: device generator (use scenes.py with the reference port)"); }

int bh_request_positions_f32(bh_engine* e) {
    if (!e) return BH_E_ARG;
    const size_t n = e->bodies.size();
    e->snapXY.resize(2 * n); e->snapM.resize(n);
    for (size_t i = 0; i < n; ++i) {
        e->snapXY[2 * i] = (float)e->bodies[i].x; e->snapXY[2 * i + 1] = (float)e->bodies[i].y;
        e->snapM[i] = (float)e->bodies[i].m;
    }
    e->snapN = (int64_t)n;
    return BH_OK;
}
int bh_wait_positions_f32(bh_engine* e, const float** xy, const float** m, int64_t* n) {
    if (!e) return BH_E_ARG;
    if (e->snapN < 0) return fail(e, BH_E_STATE, "bh_wait_positions_f32: no snapshot was requested");
    if (xy) *xy = e->snapXY.data();
    if (m) *m = e->snapM.data();
    if (n) *n = e->snapN;
    return BH_OK;
}

int bh_step(bh_engine* e, int32_t nsteps) {
    if (!e || nsteps < 0) return fail(e, BH_E_ARG, "bh_step: bad arguments");
    if (e->world > 1) return fail(e, BH_E_STATE, "bh_step: host-staged transport — drive the step with bh_step_begin / bh_step_end / bh_step_finish");
    const double t0 = now_ms();
    try { for (int s = 0; s < nsteps; ++s) e->step(); }
    catch (const std::bad_alloc&) { return fail(e, BH_E_OOM, "bh_step: out of memory"); }
    e->ctr.ms_step_call = now_ms() - t0;
    return BH_OK;
}

int bh_step_io(bh_engine* e, int32_t nsteps, int64_t n_in, const double* x_in, const double* y_in, const double* vx_in,
               const double* vy_in, const double* m_in, int64_t cap_out, double* x_out, double* y_out, double* vx_out,
               double* vy_out, double* m_out, int64_t* n_out) {
    if (!e) return BH_E_ARG;
    if (x_in) { const int rc = bh_set_bodies(e, n_in, x_in, y_in, vx_in, vy_in, m_in); if (rc != BH_OK) return rc; }
    const int rc = bh_step(e, nsteps);
    if (rc != BH_OK) return rc;
    if (n_out) *n_out = (int64_t)e->bodies.size();
    if (x_out) return bh_get_bodies(e, cap_out, x_out, y_out, vx_out, vy_out, m_out, n_out);
    return BH_OK;
}

int bh_build_tree(bh_engine* e) {
    if (!e) return BH_E_ARG;
    try { e->lastTree = e->buildTree(); }
    catch (const std::bad_alloc&) { return fail(e, BH_E_OOM, "bh_build_tree: out of memory"); }
    return BH_OK;
}

int bh_compute_accelerations(bh_engine* e, double* ax, double* ay) {
    if (!e) return BH_E_ARG;
    try {
        BHTree* root = e->buildTree();
        e->computeAccelerations(root);
        e->lastTree = root;
    } catch (const std::bad_alloc&) { return fail(e, BH_E_OOM, "bh_compute_accelerations: out of memory"); }
    const size_t n = e->bodies.size();
    if (ax) std::memcpy(ax, e->ax.data(), n * sizeof(double));
    if (ay) std::memcpy(ay, e->ay.data(), n * sizeof(double));
    return BH_OK;
}

// BH.kt:250-259 applied to every pair j != i (every body is a source), then BH.kt:390-391
int bh_direct_sum(bh_engine* e, double* ax, double* ay) {
    if (!e) return BH_E_ARG;
    const int64_t n = (int64_t)e->bodies.size();
    const Params P{e->par.G, e->par.soft2};
    const int workers = (int)std::min<int64_t>(e->cores, std::max<int64_t>(n, 1));
    std::atomic<int64_t> next{0};
    const std::vector<Body>& bs = e->bodies;
    auto work = [&]() {
        for (;;) {
            const int64_t i0 = next.fetch_add(64, std::memory_order_relaxed);
            if (i0 >= n) break;
            const int64_t i1 = std::min(n, i0 + 64);
            for (int64_t i = i0; i < i1; ++i) {
                Acc acc;
                const Body& b = bs[i];
                for (int64_t j = 0; j < n; ++j)
                    if (j != i) BHTree::pointForceAcc(b, bs[j].x, bs[j].y, bs[j].m, acc, P);
                if (ax) ax[i] = acc.fx / b.m;
                if (ay) ay[i] = acc.fy / b.m;
            }
        }
    };
    std::vector<std::thread> th;
    for (int w = 1; w < workers; ++w) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
    return BH_OK;
}

int bh_energy(bh_engine* e, double* ke, double* pe, double* px, double* py) {
    if (!e) return BH_E_ARG;
    const int64_t n = (int64_t)e->bodies.size();
    const std::vector<Body>& bs = e->bodies;
    double K = 0.0, PX = 0.0, PY = 0.0;
    for (const auto& b : bs) { K += 0.5 * b.m * (b.vx * b.vx + b.vy * b.vy); PX += b.m * b.vx; PY += b.m * b.vy; }
    const int workers = (int)std::min<int64_t>(e->cores, std::max<int64_t>(n, 1));
    std::vector<double> part((size_t)workers, 0.0);
    std::atomic<int64_t> next{0};
    auto work = [&](int w) {
        double u = 0.0;
        for (;;) {
            const int64_t i0 = next.fetch_add(64, std::memory_order_relaxed);
            if (i0 >= n) break;
            const int64_t i1 = std::min(n, i0 + 64);
            for (int64_t i = i0; i < i1; ++i) {
                double ui = 0.0;
                for (int64_t j = 0; j < n; ++j) {
                    if (j == i) continue;
                    const double dx = bs[j].x - bs[i].x, dy = bs[j].y - bs[i].y;
                    ui += bs[j].m / std::sqrt(dx * dx + dy * dy + e->par.soft2);
                }
                u += bs[i].m * ui;
            }
        }
        part[(size_t)w] = u;
    };
    std::vector<std::thread> th;
    for (int w = 1; w < workers; ++w) th.emplace_back(work, w);
    work(0);
    for (auto& t : th) t.join();
    double U = 0.0;
    for (double v : part) U += v;
    if (ke) *ke = K;
    if (pe) *pe = -0.5 * e->par.G * U;
    if (px) *px = PX;
    if (py) *py = PY;
    return BH_OK;
}

// Potential analogue of accumulateForce (BH.kt:215-239): same pruning, self-skip and opening test,
// m / sqrt(d^2 + soft2) in place of pointForceAcc.  (No reference counterpart: diagnostics.)
static double treePotential(const BHTree* t, const Body* b, double theta2, double soft2) {
    if (t->mass == 0.0) return 0.0;
    const double dx = t->comX - b->x, dy = t->comY - b->y;
    const double dist2 = dx * dx + dy * dy + soft2;
    if (t->isLeaf()) {
        if (t->body == nullptr || t->body == b) return 0.0;
        return t->mass / std::sqrt(dist2);
    }
    const double side = t->quad.h * 2.0;
    if (side * side < theta2 * dist2) return t->mass / std::sqrt(dist2);
    double s = 0.0;
    for (int c = 0; c < 4; ++c) s += treePotential(&t->children[c], b, theta2, soft2);
    return s;
}

int bh_energy_tree(bh_engine* e, double theta, double* ke, double* pe, double* px, double* py) {
    if (!e) return BH_E_ARG;
    const double th = theta > 0.0 ? theta : e->par.theta;
    double K = 0.0, U = 0.0, PX = 0.0, PY = 0.0;
    try {
        const BHTree* root = e->buildTree();
        e->lastTree = const_cast<BHTree*>(root);
        const int64_t n = (int64_t)e->bodies.size();
        const int workers = (int)std::min<int64_t>(e->cores, std::max<int64_t>(n, 1));
        std::vector<double> part((size_t)workers, 0.0);
        std::atomic<int64_t> next{0};
        auto work = [&](int w) {
            double u = 0.0;
            for (;;) {
                const int64_t i0 = next.fetch_add(256, std::memory_order_relaxed);
                if (i0 >= n) break;
                const int64_t i1 = std::min(n, i0 + 256);
                for (int64_t i = i0; i < i1; ++i) u += e->bodies[i].m * treePotential(root, &e->bodies[i], th * th, e->par.soft2);
            }
            part[(size_t)w] = u;
        };
        std::vector<std::thread> thr;
        for (int w = 1; w < workers; ++w) thr.emplace_back(work, w);
        work(0);
        for (auto& t : thr) t.join();
        for (double v : part) U += v;
        for (const auto& b : e->bodies) { K += 0.5 * b.m * (b.vx * b.vx + b.vy * b.vy); PX += b.m * b.vx; PY += b.m * b.vy; }
    } catch (const std::bad_alloc&) { return fail(e, BH_E_OOM, "bh_energy_tree: out of memory"); }
    if (ke) *ke = K;
    if (pe) *pe = -0.5 * e->par.G * U;
    if (px) *px = PX;
    if (py) *py = PY;
    return BH_OK;
}

int bh_get_morton(bh_engine* e, uint64_t*, int32_t*, int32_t*) {
    // The oracle has no Morton keys: its observer is bh_ref_get_leaf_paths() below.
    return fail(e, BH_E_UNSUPPORTED, "bh_get_morton: the reference port has no Morton keys; use bh_ref_get_leaf_paths");
}

int bh_get_tree(bh_engine* e, int64_t cap, int64_t* n_cells, double* cx, double* cy, double* h,
                double* mass, double* comx, double* comy, int32_t* body) {
    if (!e) return BH_E_ARG;
    try {
        const BHTree* root = e->getTreeForDebug();
        int64_t k = 0;
        visitAll(root, e->bodies.data(), cap, k, cx, cy, h, mass, comx, comy, body);
        if (n_cells) *n_cells = k;
        if (cap != 0 && cap < k) return fail(e, BH_E_ARG, "bh_get_tree: capacity too small");
    } catch (const std::bad_alloc&) { return fail(e, BH_E_OOM, "bh_get_tree: out of memory"); }
    return BH_OK;
}

int bh_get_tree_root(bh_engine* e, double* mass, double* comx, double* comy, int64_t* n_cells) {
    if (!e) return BH_E_ARG;
    try {
        const BHTree* root = e->getTreeForDebug();      // BH.kt:329-332
        if (mass) *mass = root->mass;                   // BH.kt:103-109
        if (comx) *comx = root->comX;
        if (comy) *comy = root->comY;
        if (n_cells) {                                  // internal + body-leaf cells (empty leaves are implicit on the device)
            int64_t k = 0;
            struct Rec { static void go(const BHTree* t, int64_t& k) {
                if (t->children) { ++k; for (int c = 0; c < 4; ++c) go(&t->children[c], k); }
                else if (t->body) ++k; } };
            Rec::go(root, k);
            *n_cells = k;
        }
    } catch (const std::bad_alloc&) { return fail(e, BH_E_OOM, "bh_get_tree_root: out of memory"); }
    return BH_OK;
}

int bh_get_counters(bh_engine* e, bh_counters* out) {
    if (!e || !out) return BH_E_ARG;
    e->ctr.n_bodies = (int64_t)e->bodies.size();
    *out = e->ctr;
    return BH_OK;
}
int bh_reset_counters(bh_engine* e) {
    if (!e) return BH_E_ARG;
    e->ctr = bh_counters{};
    return BH_OK;
}
int bh_get_body_counts(bh_engine* e, int32_t* interactions, int32_t* opened) {
    if (!e) return BH_E_ARG;
    if (!(e->cfg.flags & BH_FLAG_BODY_COUNTS)) return fail(e, BH_E_STATE, "bh_get_body_counts: engine created without BH_FLAG_BODY_COUNTS");
    if (interactions) std::memcpy(interactions, e->cntI.data(), e->cntI.size() * sizeof(int32_t));
    if (opened) std::memcpy(opened, e->cntO.data(), e->cntO.size() * sizeof(int32_t));
    return BH_OK;
}

int bh_get_let_stats(bh_engine* e, int64_t* out, int32_t n_out) {   // the port has no domain mode: all zero
    if (!e || !out || n_out < 1) return BH_E_ARG;
    for (int k = 0; k < n_out; ++k) out[k] = 0;
    return BH_OK;
}

int bh_comm_unique_id(void*, int32_t) { return BH_E_UNSUPPORTED; }
int bh_comm_init(bh_engine* e, int32_t, int32_t, const void*, int32_t) { return fail(e, BH_E_UNSUPPORTED, "bh_comm_init: the reference port has no NCCL transport"); }
int bh_comm_init_external(bh_engine* e, int32_t rank, int32_t world) {
    if (!e || world < 1 || rank < 0 || rank >= world) return fail(e, BH_E_ARG, "bh_comm_init_external: bad arguments");
    e->rank = rank; e->world = world;
    return BH_OK;
}
int bh_step_begin(bh_engine* e) {
    if (!e) return BH_E_ARG;
    if (e->phase != 0) return fail(e, BH_E_STATE, "bh_step_begin: a step is already in progress");
    try { e->stepBegin(); } catch (const std::bad_alloc&) { return fail(e, BH_E_OOM, "bh_step_begin: out of memory"); }
    return BH_OK;
}
int bh_step_end(bh_engine* e) {
    if (!e) return BH_E_ARG;
    if (e->phase != 1) return fail(e, BH_E_STATE, "bh_step_end: call bh_step_begin first");
    try { e->stepEnd(); } catch (const std::bad_alloc&) { return fail(e, BH_E_OOM, "bh_step_end: out of memory"); }
    return BH_OK;
}
int bh_step_finish(bh_engine* e) {
    if (!e) return BH_E_ARG;
    if (e->phase != 2) return fail(e, BH_E_STATE, "bh_step_finish: call bh_step_end first");
    e->stepFinish();
    return BH_OK;
}
int bh_export_slice(bh_engine* e, int32_t field, int64_t cap, double* a, double* b, int64_t* lo_out, int64_t* hi_out) {
    if (!e || (field != BH_FIELD_POS && field != BH_FIELD_VEL)) return fail(e, BH_E_ARG, "bh_export_slice: bad arguments");
    int64_t lo, hi;
    e->mySlice(&lo, &hi);
    if (lo_out) *lo_out = lo;
    if (hi_out) *hi_out = hi;
    if (cap < hi - lo) return fail(e, BH_E_ARG, "bh_export_slice: capacity too small");
    for (int64_t i = lo; i < hi; ++i) {
        const Body& q = e->bodies[i];
        if (a) a[i - lo] = field == BH_FIELD_POS ? q.x : q.vx;
        if (b) b[i - lo] = field == BH_FIELD_POS ? q.y : q.vy;
    }
    return BH_OK;
}
int bh_import_slices(bh_engine* e, int32_t field, int64_t n, const double* a, const double* b) {
    if (!e || (field != BH_FIELD_POS && field != BH_FIELD_VEL) || !a || !b) return fail(e, BH_E_ARG, "bh_import_slices: bad arguments");
    if (n != (int64_t)e->bodies.size()) return fail(e, BH_E_ARG, "bh_import_slices: n must equal bh_num_bodies");
    for (int64_t i = 0; i < n; ++i) {
        Body& q = e->bodies[i];
        if (field == BH_FIELD_POS) { q.x = a[i]; q.y = b[i]; } else { q.vx = a[i]; q.vy = b[i]; }
    }
    if (field == BH_FIELD_POS) e->lastTree = nullptr;
    return BH_OK;
}
int bh_get_slice_index(bh_engine* e, int64_t cap, int32_t* user_index, int64_t* n_slice) {
    if (!e) return BH_E_ARG;
    int64_t lo, hi;
    e->mySlice(&lo, &hi);
    if (n_slice) *n_slice = hi - lo;
    if (cap < hi - lo) return fail(e, BH_E_ARG, "bh_get_slice_index: capacity too small");
    if (user_index) for (int64_t k = lo; k < hi; ++k) user_index[k - lo] = (int32_t)k;   // the port keeps list order
    return BH_OK;
}
int64_t bh_slice_epoch(const bh_engine*) { return 0; }
int bh_step_io_slice(bh_engine* e, int32_t nsteps, int64_t n_in, const double* x_in, const double* y_in, const double* vx_in,
                     const double* vy_in, const double* m_in, int64_t cap_out, double* x_out, double* y_out, double* vx_out,
                     double* vy_out, double* m_out, int64_t* n_out) {
    if (!e) return BH_E_ARG;
    if (e->world > 1) return fail(e, BH_E_UNSUPPORTED, "bh_step_io_slice: the reference port is single-process here (its slice is the whole list)");
    return bh_step_io(e, nsteps, n_in, x_in, y_in, vx_in, vy_in, m_in, cap_out, x_out, y_out, vx_out, vy_out, m_out, n_out);
}

int bh_evaluate_slice(bh_engine* e, int64_t cap, double* ax, double* ay, int32_t* user_index, int64_t* n_slice) {
    if (!e) return BH_E_ARG;
    int64_t lo, hi;
    e->mySlice(&lo, &hi);
    if (n_slice) *n_slice = hi - lo;
    if (cap < hi - lo) return fail(e, BH_E_ARG, "bh_evaluate_slice: capacity too small");
    const size_t n = e->bodies.size();
    std::vector<double> fx(n), fy(n);
    const int rc = bh_compute_accelerations(e, fx.data(), fy.data());   // the port evaluates every body
    if (rc != BH_OK) return rc;
    for (int64_t k = lo; k < hi; ++k) {
        if (ax) ax[k - lo] = fx[(size_t)k];
        if (ay) ay[k - lo] = fy[(size_t)k];
        if (user_index) user_index[k - lo] = (int32_t)k;
    }
    return BH_OK;
}
int bh_set_domain_mode(bh_engine* e, int32_t) { return e ? BH_OK : BH_E_ARG; }   // the port has no domain mode

int bh_slice_bounds(int64_t n, int32_t world, int32_t rank, int64_t* lo, int64_t* hi) {
    if (n < 0 || world < 1 || rank < 0 || rank >= world || !lo || !hi) return BH_E_ARG;
    const int64_t per = (n + world - 1) / world;
    *lo = std::min<int64_t>(n, per * rank);
    *hi = std::min<int64_t>(n, per * (rank + 1));
    return BH_OK;
}

int bh_measure_fp32_tflops(int32_t, double*) { return BH_E_UNSUPPORTED; }

// ---- oracle-only observers (not part of bh_engine.h) --------------------------------
// For every body: depth of the leaf holding it in the tree of the LAST build (0 = root,
// -1 = not in the tree: rejected at BH.kt:126 or dropped after a jitter) and the 2-bit
// child digits of its root-to-leaf path (BH.kt:153-155), first digit most significant,
// right-aligned (`depth` digits).  Also fills the tree statistics counters.
int bh_ref_get_leaf_paths(bh_engine* e, int32_t* depth, uint64_t* path) {
    if (!e) return BH_E_ARG;
    const BHTree* root = e->getTreeForDebug();
    const int64_t n = (int64_t)e->bodies.size();
    for (int64_t i = 0; i < n; ++i) { if (depth) depth[i] = -1; if (path) path[i] = 0; }
    int64_t nInternal = 0;
    int maxDepth = 0;
    leafPaths(root, e->bodies.data(), 0, 0, depth, path, nInternal, maxDepth);
    int64_t inTree = 0;
    if (depth) for (int64_t i = 0; i < n; ++i) inTree += depth[i] >= 0;
    e->ctr.n_internal = nInternal;
    e->ctr.max_depth = maxDepth;
    e->ctr.n_in_tree = inTree;
    e->ctr.n_cells = nInternal + inTree;
    return BH_OK;
}

int bh_ref_threads(const bh_engine* e) { return e ? e->cores : 0; }

}  // extern "C"
