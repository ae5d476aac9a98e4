/*
 * bh_engine.h — C ABI of the Barnes–Hut physics-step engine (drop-in boundary).
 *
 * The reference (qwertukg/Barnes-Hut-N-Body) has no FFI of its own: the boundary
 * is the Kotlin class API in BarnesHutAlg.kt that NBodyPanel.kt calls.  Every
 * entry point below names the reference interface (file:line, paths relative to
 * /root/reference/src/main/kotlin/) it replaces.  Two shared libraries export
 * this ABI:
 *
 *   libbh_b200.so  — the product: hand-written sm_100a CUDA engine
 *                    (barnes-hut-n-body_b200/csrc/).  No CPU fallback.
 *   libbh_ref.so   — the oracle: literal f64 C++ restatement of BarnesHutAlg.kt
 *                    (oracle/).  Test infrastructure + CPU baseline only.
 *
 * Conventions
 *   - Every function returns an int status (BH_OK == 0, negative = error) unless
 *     stated; nothing throws or aborts across the ABI.  bh_last_error() gives text.
 *   - The caller owns all host arrays; they are only read/written during the call.
 *   - State is SoA f64 (x, y, vx, vy, m), index = position in the reference's
 *     `bodies` list (BarnesHutAlg.kt:295).  An output pointer may be NULL to skip it.
 *   - An engine is single-owner, not re-entrant (the reference is driven from the
 *     Swing EDT only: NBodyPanel.kt:106,290-293).  Calls block until complete.
 */
#ifndef BH_ENGINE_H
#define BH_ENGINE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BH_ABI_VERSION 1

/* status codes */
#define BH_OK             0
#define BH_E_ARG         (-1)  /* bad argument / NULL handle                       */
#define BH_E_CUDA        (-2)  /* CUDA runtime error (text in bh_last_error)       */
#define BH_E_NCCL        (-3)  /* NCCL error / NCCL not loadable                   */
#define BH_E_OOM         (-4)  /* host or device allocation failed                 */
#define BH_E_STATE       (-5)  /* call not valid in the current engine state       */
#define BH_E_UNSUPPORTED (-6)  /* not implemented by this library                  */

/* bh_config.flags */
#define BH_FLAG_BODY_COUNTS  1u  /* keep per-body interaction / opened-cell counts  */
#define BH_FLAG_REUSE_ACC    2u  /* CUDA engine: step n+1 starts from the accelerations a(t+dt)
                                  * step n ended with instead of rebuilding and re-walking the
                                  * unchanged positions (BarnesHutAlg.kt:407-408 recomputes them).
                                  * Result-identical; automatically suspended after anything that
                                  * changes the inputs (set_bodies, Config, merge, jitter).       */
#define BH_FLAG_LET          4u  /* CUDA engine, multi-GPU (bh_comm_init): DOMAIN mode.  Every rank
                                  * builds only the tree of its own Morton range and walks its bodies
                                  * over a locally essential tree (top tree from all-reduced level
                                  * summaries + imported boundary subtrees) instead of replicating the
                                  * whole tree; no per-step all-gather of positions.  Bit-identical to
                                  * a single-GPU run; falls back to the replicated tree while the merge
                                  * rule is enabled and below 4 ranks (BH_LET_MIN_WORLD), where replicating
                                  * the tree is cheaper.  Env BH_LET=0/1 overrides.                   */

typedef struct bh_engine bh_engine;

/* Engine construction parameters (no counterpart in the reference, whose engine
 * is constructed from a body list only: BarnesHutAlg.kt:287). */
typedef struct bh_config {
    int32_t  struct_size;   /* sizeof(bh_config), for forward compatibility        */
    int32_t  device;        /* CUDA device ordinal (ignored by the oracle)         */
    int32_t  threads;       /* oracle: worker threads, 0 = all cores (:292,:377)   */
    uint32_t flags;         /* BH_FLAG_*                                           */
    int64_t  capacity_hint; /* bodies to pre-size buffers for (0 = grow on demand) */
    int32_t  rehome_interval; /* CUDA engine: re-sort the device-resident state into Morton
                               * ("home") order every this many steps; 0 = default (8)      */
    int32_t  reserved;
} bh_config;

/* The values the reference reads live from `object Config` on every use
 * (Config.kt:5-23; read sites BarnesHutAlg.kt:256,360-361,378,412,420) plus the
 * two merge knobs (BarnesHutAlg.kt:315,321). */
typedef struct bh_params {
    double G;               /* Config.G        = 80.0  (Config.kt:11)              */
    double dt;              /* Config.DT       = 0.005 (Config.kt:14)              */
    double theta;           /* Config.theta    = 0.30  (Config.kt:23)              */
    double soft2;           /* Config.SOFT2    = 1.0   (Config.kt:20)              */
    double root_cx;         /* WIDTH_PX / 2.0           (BarnesHutAlg.kt:361)      */
    double root_cy;         /* HEIGHT_PX / 2.0          (BarnesHutAlg.kt:361)      */
    double root_half;       /* max(W,H) / 2.0 + 2.0     (BarnesHutAlg.kt:360)      */
    double merge_max_mass;  /* PhysicsEngine.mergeMaxMass = 4000 (:315)            */
    double merge_min_dist;  /* PhysicsEngine.mergeMinDist = 8.0  (:321); <=0 = off */
} bh_params;

/* Counters of the most recent force evaluation / build, and running totals.
 * interaction = one pointForceAcc call (BarnesHutAlg.kt:220 or :230);
 * opened      = one internal node that failed the test at :228;
 * node visits of the reference walk = n_targets + 4*opened. */
typedef struct bh_counters {
    int64_t n_bodies;           /* bodies in the engine                             */
    int64_t n_in_tree;          /* bodies inserted by the last build (:126 passed)  */
    int64_t n_out_of_box;       /* bodies rejected by the root contains() test      */
    int64_t n_jitter_bodies;    /* bodies that share a cell with h<1e-3 (:146)      */
    int64_t n_cells;            /* internal + body-leaf cells of the last tree      */
    int64_t n_internal;         /* internal cells of the last tree                  */
    int32_t key_levels;         /* levels encoded in a Morton key (first h<1e-3)    */
    int32_t max_depth;          /* deepest leaf of the last tree                    */
    int64_t interactions;       /* last evaluation                                  */
    int64_t opened;             /* last evaluation                                  */
    int64_t exact_retests;      /* last evaluation: f64 re-tests of borderline MACs */
    int64_t total_interactions; /* since bh_reset_counters                          */
    int64_t total_opened;
    int64_t total_evaluations;
    int64_t total_steps;
    int64_t total_merged;       /* bodies absorbed by the merge rule                */
    double  ms_build;           /* device/host time of the phases, accumulated      */
    double  ms_walk;            /*   since bh_reset_counters (CUDA events / clock)  */
    double  ms_integrate;
    double  ms_merge;
    double  ms_comm;
    double  ms_step_call;       /* device time (CUDA events on the engine's stream) of */
                                /*   the most recent bh_step call, all its steps       */
    int64_t kernel_launches;    /* CUDA kernels launched since bh_reset_counters       */
    double  bbox_min_x, bbox_max_x, bbox_min_y, bbox_max_y;
                                /* bounding box of ALL bodies seen by the last build (a min/max
                                 * reduction fused into the key generation).  The reference's root
                                 * box is the window (BarnesHutAlg.kt:360-361), not this box: bodies
                                 * outside the root are dropped by :126, so a box that pokes out of
                                 * the root is the diagnosis behind n_out_of_box > 0.  NaN when empty. */
    double  ms_direct;          /* device time of the most recent bh_direct_sum kernel */
} bh_counters;

/* ---- lifecycle ---------------------------------------------------------- */

/* PhysicsEngine(initialBodies) constructor — BarnesHutAlg.kt:287-301. */
int  bh_create(const bh_config* cfg, bh_engine** out);
void bh_destroy(bh_engine* e);
/* text of the last error on this engine (or of the last failed bh_create if e==NULL) */
const char* bh_last_error(const bh_engine* e);
/* "b200-cuda" or "reference-port"; lets a harness assert which library it loaded */
const char* bh_backend_name(void);
int  bh_abi_version(void);

/* ---- parameters --------------------------------------------------------- */

/* Fill *p with the reference defaults for a W x H window: Config.kt:5-23 and the
 * root-box rule of buildTree(), BarnesHutAlg.kt:360-361. */
int bh_default_params(int32_t width_px, int32_t height_px, bh_params* p);
/* Snapshot of `Config` the next calls will use (the reference re-reads Config at
 * every use; a façade calls this at the top of step()). */
int bh_set_params(bh_engine* e, const bh_params* p);
int bh_get_params(const bh_engine* e, bh_params* p);

/* ---- state -------------------------------------------------------------- */

/* resetBodies(newBodies) — BarnesHutAlg.kt:342-349 (and the constructor).  n may be 0. */
int bh_set_bodies(bh_engine* e, int64_t n, const double* x, const double* y,
                  const double* vx, const double* vy, const double* m);
/* getBodies() — BarnesHutAlg.kt:335.  cap = capacity of the output arrays. */
int bh_get_bodies(bh_engine* e, int64_t cap, double* x, double* y,
                  double* vx, double* vy, double* m, int64_t* n_out);
int64_t bh_num_bodies(const bh_engine* e);
/* origin[k] = index, in the list given to the last bh_set_bodies, of the body now at
 * position k.  Identity until the merge rule (BarnesHutAlg.kt:514-520) removes
 * bodies; lets a façade write state back into the SAME Body objects the UI holds. */
int bh_get_origin(bh_engine* e, int64_t cap, int32_t* origin, int64_t* n_out);
/* Declare the CURRENT list the new reference list: afterwards bh_get_origin is the identity again.
 * A façade calls it after it has removed the merged-away Body objects from its own list
 * (bodies.removeAt, BarnesHutAlg.kt:519), so that the next step's origin[] indexes the shrunk
 * list.  Nothing is copied and the device state is untouched. */
int bh_rebase_origin(bh_engine* e);
/* render read-back used by NBodyPanel.paintComponent (NBodyPanel.kt:302-306):
 * interleaved float (x,y) pairs and float masses. */
int bh_get_positions_f32(bh_engine* e, int64_t cap, float* xy, float* m, int64_t* n_out);

/* The same read-back, overlapped with the next step (the renderer of frame k runs while the
 * engine computes frame k+1): bh_request_positions_f32 snapshots (x, y, m) as floats in list order
 * and starts the device->host copy on a second stream; bh_step may be called right away;
 * bh_wait_positions_f32 blocks until the snapshot is in host memory and returns pointers into
 * engine-owned (pinned) buffers, valid until the next request. */
int bh_request_positions_f32(bh_engine* e);
int bh_wait_positions_f32(bh_engine* e, const float** xy, const float** m, int64_t* n);

/* ---- scene generators on the device (BodyFactory.kt) ------------------------------
 * The reference's UI injects bodies with resetBodies(old + BodyFactory.make...(...))
 * (NBodyPanel.kt:228-234, :282-286).  These entry points do the same on the device, so 10M-100M
 * body scenes never cross PCIe.  Kotlin's XorWow stream and the JDK's libm are not reproducible
 * bit-for-bit (and the reference seeds them from the clock, BodyFactory.kt:74), so parity is
 * DISTRIBUTIONAL: same sampling laws, a counter-based generator keyed by `seed`.
 * New bodies are appended behind the existing ones in list order; afterwards bh_get_origin is
 * the identity over the new list. */
typedef struct bh_disk_params {
    double x, y;                 /* centre                              (BodyFactory.kt:75-76)   */
    double vx, vy;               /* drift velocity of the whole disk    (:75)                    */
    double r;                    /* rMax                                (:77, default 200)       */
    double min_r;                /* Config.MIN_R = 8                    (:78)                    */
    double central_mass;         /* Config.CENTRAL_MASS = 50000         (:79)                    */
    double total_satellite_mass; /* Config.TOTAL_SATELLITE_MASS = 5000  (:80)                    */
    double eps_m2;               /* m=2 bar amplitude, 0.03             (:66)                    */
    double phi0;                 /* bar phase                           (:67)                    */
    double bar_taper_r;          /* <= 0: 0.6 r                         (:68, :93)               */
    double radial_scale;         /* Rd; <= 0: r / 3                     (:70, :92)               */
    double speed_jitter;         /* 0.01                                (:71)                    */
    double radial_jitter;        /* 0                                   (:72)                    */
    int32_t clockwise;           /* 1                                   (:73)                    */
    int32_t kepler;              /* 1: makeKeplerDisk's law instead (uniform in area on
                                  * [min_r, r], radius jittered by radial_jitter, :35-41)         */
} bh_disk_params;
/* defaults of makeGalaxyDisk (BodyFactory.kt:63-81) centred on a W x H window */
int bh_default_disk_params(int32_t width_px, int32_t height_px, bh_disk_params* p);
/* bodies += makeGalaxyDisk / makeKeplerDisk(n_total, ...): 1 central body + (n_total-1) satellites
 * with circular speeds from the exact enclosed mass (:118-147); n_total <= 1: the central body only
 * (the right-mouse-button "black hole", NBodyPanel.kt:171). */
int bh_append_disk(bh_engine* e, int64_t n_total, const bh_disk_params* p, uint64_t seed);
/* bodies += makeUniformRandom(n, m): x ~ U[0,W), y ~ U[0,H), v = 0 (BodyFactory.kt:160-177). */
int bh_append_uniform_random(bh_engine* e, int64_t n, double m, int32_t width_px, int32_t height_px, uint64_t seed);

/* ---- compute ------------------------------------------------------------ */

/* nsteps x PhysicsEngine.step() — BarnesHutAlg.kt:405-439: build+eval, half kick,
 * drift, build+eval, half kick, merge rule (:463-532). */
int bh_step(bh_engine* e, int32_t nsteps);
/* resetBodies(in) ; nsteps x step() ; getBodies(out) in ONE call, so that the transfers overlap
 * the compute: (x, y) go up first and the build starts while (vx, vy, m) are still in flight on a
 * second stream; the final positions and masses come down while the last force evaluation runs.
 * Same results as bh_set_bodies + bh_step + bh_get_bodies (which it falls back to when the merge
 * rule is enabled or a multi-rank transport is set).  x_in == NULL: keep the current bodies;
 * x_out == NULL: no read-back.  Page-locked host arrays make the copies truly asynchronous. */
int bh_step_io(bh_engine* e, int32_t nsteps, int64_t n_in,
               const double* x_in, const double* y_in, const double* vx_in, const double* vy_in, const double* m_in,
               int64_t cap_out, double* x_out, double* y_out, double* vx_out, double* vy_out, double* m_out,
               int64_t* n_out);
/* buildTree() + computeAccelerations() on the current state, no integration —
 * BarnesHutAlg.kt:359-366 + :374-395.  The parity entry point.  ax/ay: n doubles. */
int bh_compute_accelerations(bh_engine* e, double* ax, double* ay);
/* All-pairs direct sum with the same softened kernel (BarnesHutAlg.kt:250-259
 * over every j != i, every body a source incl. out-of-box ones): accuracy oracle. */
int bh_direct_sum(bh_engine* e, double* ax, double* ay);
/* E = sum 1/2 m v^2  -  1/2 G sum_{i!=j} m_i m_j / sqrt(r_ij^2 + soft2); momentum. */
int bh_energy(bh_engine* e, double* kinetic, double* potential, double* px, double* py);
/* The same diagnostics with the potential taken from the TREE (no reference counterpart; the
 * all-pairs sum above is O(N^2)): phi_i is accumulated by the walk of BarnesHutAlg.kt:215-239 with
 * 1/sqrt(d^2+soft2) in place of the force, on a tree built from the current state (buildTree(),
 * :359-366 — including its jitter side effect), with opening angle `theta` (<= 0: the engine's).
 * Sources are the bodies the tree holds (out-of-box bodies attract nobody, :126).  O(N log N):
 * energy-drift tracking at 10M+ bodies. */
int bh_energy_tree(bh_engine* e, double theta, double* kinetic, double* potential, double* px, double* py);

/* ---- introspection (parity artefacts) ----------------------------------- */

/* After a build (bh_compute_accelerations / bh_step / bh_build_tree):
 *   key[i]   Morton key of body i: `key_levels` 2-bit digits, MSB first, digit =
 *            (x<cx?0:1)+(y<cy?0:2) at each level (BarnesHutAlg.kt:153-155);
 *            UINT64_MAX for a body the root rejected (:126).
 *   depth[i] depth of the leaf that holds body i (0 = root), -1 if not in the tree.
 *   order[k] body index at sorted position k (in-tree bodies first, by key). */
int bh_get_morton(bh_engine* e, uint64_t* key, int32_t* depth, int32_t* order);
/* getTreeForDebug().visitQuads — BarnesHutAlg.kt:329-332, :265-274: DFS preorder
 * over ALL cells incl. empty leaves, children in order 0..3.  body[k] = index of
 * the body in a leaf, -1 for an empty leaf, -2 for an internal cell.  Pass cap=0
 * to query *n_cells only. */
int bh_get_tree(bh_engine* e, int64_t cap, int64_t* n_cells,
                double* cx, double* cy, double* h,
                double* mass, double* comx, double* comy, int32_t* body);
/* BHTree.mass / comX / comY of the ROOT (BarnesHutAlg.kt:103-109) without exporting the tree
 * (an empty tree reports mass 0 at the root centre, :179-183); *n_cells = internal + body-leaf
 * cells of the device tree.  Builds the tree if none is cached (:329-332). */
int bh_get_tree_root(bh_engine* e, double* mass, double* comx, double* comy, int64_t* n_cells);
/* buildTree() only (BarnesHutAlg.kt:359-366); what getTreeForDebug() triggers. */
int bh_build_tree(bh_engine* e);
int bh_get_counters(bh_engine* e, bh_counters* out);
int bh_reset_counters(bh_engine* e);
/* domain-mode statistics (BH_FLAG_LET): out[0..9] = mode (0 off, 1 on with blocks over
 * ncclSend/ncclRecv, 2 on with blocks over NVLink peer memory), partition valid, level of the cut,
 * LET evaluations, fallbacks to a re-homing build, cells of this rank's LET, cells imported, cells
 * sent, own strays, items of the top tree (last evaluation).  With BH_LET_TIMERS=1 in the
 * environment the call also prints this rank's per-phase times to stderr. */
int bh_get_let_stats(bh_engine* e, int64_t* out, int32_t n_out);
/* per-body counts of the last evaluation (needs BH_FLAG_BODY_COUNTS) */
int bh_get_body_counts(bh_engine* e, int32_t* interactions, int32_t* opened);

/* ---- multi-GPU / multi-process: replicated tree, Morton-sliced targets (no reference
 *      counterpart; the reference is single-process).
 *
 * Every rank holds the full body list (same bh_set_bodies on every rank) and builds the full
 * tree; rank r walks and integrates only its slice [lo,hi) of the engine's internal ("home",
 * Morton-sorted at the last re-homing) body order, which is identical on every rank.  One
 * exchange of the drifted positions per step.  Results are bit-identical to one process.
 * Two transports:
 *   bh_comm_init           NCCL inside the engine (CUDA library only): bh_step() is complete.
 *   bh_comm_init_external  host-staged: the caller moves the slices between ranks with its own
 *                          transport (MPI, gloo, JVM sockets ...) and drives a step as
 *                              bh_step_begin            a(t), half kick, drift   (own slice)
 *                              exchange BH_FIELD_POS    bh_export_slice -> all-gather -> bh_import_slices
 *                              bh_step_end              a(t+dt), half kick       (own slice)
 *                              exchange BH_FIELD_VEL
 *                              bh_step_finish           merge rule, counters
 * ------------------------------------------------------------------------------------------ */

#define BH_COMM_ID_BYTES 128
#define BH_FIELD_POS 0   /* (x, y)   */
#define BH_FIELD_VEL 1   /* (vx, vy) */
/* rank 0 creates an id and ships it to the other ranks (torch.distributed, a file...) */
int bh_comm_unique_id(void* id_out, int32_t id_bytes);
/* every rank: join a `world`-rank NCCL communicator; afterwards bh_step walks only this
 * rank's slice of targets and all-gathers the drifted positions over NCCL / NVLink. */
int bh_comm_init(bh_engine* e, int32_t rank, int32_t world, const void* id, int32_t id_bytes);
/* every rank: host-staged transport (see above); bh_step() is then refused for world > 1. */
int bh_comm_init_external(bh_engine* e, int32_t rank, int32_t world);
/* the slice [lo,hi) of home-ordered bodies rank `rank` of `world` owns */
int bh_slice_bounds(int64_t n, int32_t world, int32_t rank, int64_t* lo, int64_t* hi);
/* the three phases of PhysicsEngine.step() (BarnesHutAlg.kt:405-439) for the host-staged
 * transport; with world == 1 their sequence equals bh_step(e, 1). */
int bh_step_begin(bh_engine* e);    /* :407-422  build, a(t), v += a dt/2, x += v dt */
int bh_step_end(bh_engine* e);      /* :425-435  build, a(t+dt), v += a dt/2         */
int bh_step_finish(bh_engine* e);   /* :438      mergeCloseBodiesIfNeeded()          */
/* this rank's slice of `field` in home order: a[0..hi-lo), b[0..hi-lo); cap >= hi-lo */
int bh_export_slice(bh_engine* e, int32_t field, int64_t cap, double* a, double* b,
                    int64_t* lo, int64_t* hi);
/* the concatenation of every rank's exported slice (n = bh_num_bodies doubles each) */
int bh_import_slices(bh_engine* e, int32_t field, int64_t n, const double* a, const double* b);

/* Sharded end-to-end I/O: every rank moves ONLY ITS OWN SLICE between host and device.  The slice is the set of
 * bodies this rank walks and integrates, in the engine's home order; bh_get_slice_index names them (positions in
 * the `bodies` list) as the device holds them at the time of the call, and bh_slice_epoch changes whenever the
 * slices were re-cut or re-ordered (re-homing: bodies migrate between ranks), i.e. whenever the index has to be
 * fetched again: arrays that come out of a call are in the order the index shows AFTER that call.
 *   bh_step_io_slice: [slice state in: x, y, vx, vy, m of the slice, home order; NULL = keep] ; nsteps x step() ;
 *   [slice state out].  The rest of the state is exchanged between the ranks by the engine as the mode needs it
 *   (nothing in domain mode; positions / masses / velocities of the other slices before a replicated build).
 * With one rank the slice is the whole list (in home order).  The call is COLLECTIVE: every rank makes it (with or
 * without inputs).  Refused while the merge rule is enabled on more than one rank (removals re-index the list on
 * every rank). */
int bh_get_slice_index(bh_engine* e, int64_t cap, int32_t* user_index, int64_t* n_slice);
int64_t bh_slice_epoch(const bh_engine* e);
int bh_step_io_slice(bh_engine* e, int32_t nsteps, int64_t n_in,
                     const double* x_in, const double* y_in, const double* vx_in, const double* vy_in, const double* m_in,
                     int64_t cap_out, double* x_out, double* y_out, double* vx_out, double* vy_out, double* m_out,
                     int64_t* n_out);

/* One force evaluation (buildTree + computeAccelerations, BarnesHutAlg.kt:359-366 + :374-395) of THIS RANK'S
 * slice of targets, in the multi-GPU mode the engine is in (domain mode / replicated tree), without
 * integrating: ax/ay[k] for the k-th body of the slice, user_index[k] = its position in the `bodies` list.
 * With bh_set_domain_mode this is the self-check of a multi-GPU run: the same state evaluated in both modes
 * must give bit-identical accelerations and equal interaction counts (bench.py `parity_check`). */
int bh_evaluate_slice(bh_engine* e, int64_t cap, double* ax, double* ay, int32_t* user_index, int64_t* n_slice);
/* switch the domain mode (BH_FLAG_LET) on/off at run time; takes effect at the next evaluation */
int bh_set_domain_mode(bh_engine* e, int32_t enabled);

/* ---- diagnostics ---------------------------------------------------------- */

/* Measured FP32 FMA throughput of `device` in TFLOP/s (2 flop per FFMA): the roofline
 * denominator of the force walk and the direct sum, which are FP32-issue bound rather
 * than HBM- or tensor-bound.  Runs a register-only FFMA microbenchmark kernel. */
int bh_measure_fp32_tflops(int32_t device, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* BH_ENGINE_H */
