// Drop-in replacement of /src/main/kotlin/BarnesHutAlg.kt of qwertukg/Barnes-Hut-N-Body: the same
// classes and members (Body, Quad, BHTree, PhysicsEngine), routed to libbh_b200.so through JNA
// direct mapping (C ABI: include/bh_engine.h).  NBodyPanel.kt, BodyFactory.kt, Config.kt and Main.kt
// stay the reference's own files.  COMPILE-UNVERIFIED: no JVM/Kotlin toolchain exists in the image
// this repository was built in; the Python twin barnes-hut-n-body_b200/engine.py implements the
// same mapping and is what the tests exercise.  See INTEGRATION.md.
import com.sun.jna.*
import com.sun.jna.ptr.*

data class Body(var x: Double, var y: Double, var vx: Double, var vy: Double, var m: Double)
data class Quad(val cx: Double, val cy: Double, val h: Double) {
    /** half-open box test, BarnesHutAlg.kt:61-62 */
    fun contains(b: Body): Boolean = b.x >= cx - h && b.x < cx + h && b.y >= cy - h && b.y < cy + h
    /** 0 = NW, 1 = NE, 2 = SW, 3 = SE, BarnesHutAlg.kt:73-81 */
    fun child(which: Int): Quad {
        val hh = h / 2.0
        return when (which) {
            0 -> Quad(cx - hh, cy - hh, hh)
            1 -> Quad(cx + hh, cy - hh, hh)
            2 -> Quad(cx - hh, cy + hh, hh)
            else -> Quad(cx + hh, cy + hh, hh)
        }
    }
}

@Structure.FieldOrder("struct_size", "device", "threads", "flags", "capacity_hint", "rehome_interval", "reserved")
class BhConfig : Structure() {
    @JvmField var struct_size = 32; @JvmField var device = 0; @JvmField var threads = 0
    @JvmField var flags = 0; @JvmField var capacity_hint = 0L
    @JvmField var rehome_interval = 0; @JvmField var reserved = 0
}
@Structure.FieldOrder("G", "dt", "theta", "soft2", "root_cx", "root_cy", "root_half", "merge_max_mass", "merge_min_dist")
class BhParams : Structure() {
    @JvmField var G = 0.0; @JvmField var dt = 0.0; @JvmField var theta = 0.0; @JvmField var soft2 = 0.0
    @JvmField var root_cx = 0.0; @JvmField var root_cy = 0.0; @JvmField var root_half = 0.0
    @JvmField var merge_max_mass = 0.0; @JvmField var merge_min_dist = 0.0
}

object BhNative {
    init { Native.register("bh_b200") }          // libbh_b200.so
    @JvmStatic external fun bh_create(cfg: BhConfig?, out: PointerByReference): Int
    @JvmStatic external fun bh_destroy(e: Pointer)
    @JvmStatic external fun bh_last_error(e: Pointer?): String
    @JvmStatic external fun bh_default_params(w: Int, h: Int, p: BhParams): Int
    @JvmStatic external fun bh_set_params(e: Pointer, p: BhParams): Int
    @JvmStatic external fun bh_set_bodies(e: Pointer, n: Long, x: DoubleArray, y: DoubleArray,
                                          vx: DoubleArray, vy: DoubleArray, m: DoubleArray): Int
    @JvmStatic external fun bh_get_bodies(e: Pointer, cap: Long, x: DoubleArray, y: DoubleArray, vx: DoubleArray,
                                          vy: DoubleArray, m: DoubleArray, nOut: LongByReference): Int
    @JvmStatic external fun bh_num_bodies(e: Pointer): Long
    @JvmStatic external fun bh_get_origin(e: Pointer, cap: Long, origin: IntArray, nOut: LongByReference): Int
    @JvmStatic external fun bh_rebase_origin(e: Pointer): Int
    @JvmStatic external fun bh_step(e: Pointer, nsteps: Int): Int
    /** resetBodies + nsteps x step + getBodies in ONE call, host<->device copies overlapped with the compute
     *  (merge rule off); the arrays may be null (keep the current bodies / no read-back). */
    @JvmStatic external fun bh_step_io(e: Pointer, nsteps: Int, nIn: Long, xIn: DoubleArray?, yIn: DoubleArray?, vxIn: DoubleArray?,
                                       vyIn: DoubleArray?, mIn: DoubleArray?, capOut: Long, xOut: DoubleArray?, yOut: DoubleArray?,
                                       vxOut: DoubleArray?, vyOut: DoubleArray?, mOut: DoubleArray?, nOut: LongByReference?): Int
    @JvmStatic external fun bh_get_tree(e: Pointer, cap: Long, nCells: LongByReference, cx: DoubleArray?, cy: DoubleArray?,
                                        h: DoubleArray?, mass: DoubleArray?, comx: DoubleArray?, comy: DoubleArray?,
                                        body: IntArray?): Int
}

/** Per-worker force accumulator, BarnesHutAlg.kt:33-41. */
class Acc { var fx = 0.0; var fy = 0.0; fun reset() { fx = 0.0; fy = 0.0 } }

/** Host view of the device quadtree: visitQuads order incl. empty leaves (BarnesHutAlg.kt:265-274).
 *  body[k] = index of the body in a leaf, -1 empty leaf, -2 internal cell (which always has 4 children). */
class BHTree(private val cx: DoubleArray, private val cy: DoubleArray, private val h: DoubleArray,
             private val massArr: DoubleArray, private val comxArr: DoubleArray, private val comyArr: DoubleArray,
             private val body: IntArray, private val bodies: List<Body>) {
    val mass = massArr.firstOrNull() ?: 0.0
    val comX = comxArr.firstOrNull() ?: 0.0
    val comY = comyArr.firstOrNull() ?: 0.0
    fun visitQuads(visit: (Quad) -> Unit) { for (i in cx.indices) visit(Quad(cx[i], cy[i], h[i])) }

    /** The tree is built on the device from the engine's body list (BarnesHutAlg.kt:359-366): a host-side insert
     *  into this read-only view would not be seen by step().  Use PhysicsEngine.resetBodies(old + new). */
    fun insert(b: Body): Nothing = throw UnsupportedOperationException("BHTree is a read-only view; use PhysicsEngine.resetBodies")
    /** Masses and centres of mass arrive computed (bit-identical to BarnesHutAlg.kt:173-202). */
    fun computeMass() {}

    /** BarnesHutAlg.kt:215-239 over the exported cells (f64, same expression order): for diagnostics / tests. */
    fun accumulateForce(b: Body, theta2: Double, acc: Acc) { if (cx.isNotEmpty()) walk(0, b, theta2, acc) }
    private fun walk(p: Int, b: Body, theta2: Double, acc: Acc): Int {      // returns the position after the subtree
        val internal = body[p] == -2
        if (massArr[p] == 0.0) return skip(p)                                // :216
        if (!internal) {                                                      // :218-221
            if (body[p] >= 0 && bodies[body[p]] !== b) point(b, comxArr[p], comyArr[p], massArr[p], acc)
            return p + 1
        }
        val dx = comxArr[p] - b.x; val dy = comyArr[p] - b.y
        val dist2 = dx * dx + dy * dy + Config.SOFT2                          // :223-225
        val side = h[p] * 2.0
        if (side * side < theta2 * dist2) { point(b, comxArr[p], comyArr[p], massArr[p], acc); return skip(p) }
        var q = p + 1
        repeat(4) { q = walk(q, b, theta2, acc) }                             // :233-237
        return q
    }
    private fun skip(p: Int): Int { if (body[p] != -2) return p + 1; var q = p + 1; repeat(4) { q = skip(q) }; return q }
    private fun point(b: Body, px: Double, py: Double, m: Double, acc: Acc) { // :250-259
        val dx = px - b.x; val dy = py - b.y
        val r2 = dx * dx + dy * dy + Config.SOFT2
        val invR = 1.0 / kotlin.math.sqrt(r2); val invR2 = 1.0 / r2
        val f = Config.G * b.m * m * invR2
        acc.fx += f * dx * invR; acc.fy += f * dy * invR
    }
}

class PhysicsEngine(initialBodies: MutableList<Body>) {
    private val e: Pointer
    private var bodies: MutableList<Body> = initialBodies
    private var lastTree: BHTree? = null
    var mergeMaxMass: Double = 4_000.0
    var mergeMinDist: Double = Config.MIN_R

    init {
        val out = PointerByReference()
        check(BhNative.bh_create(BhConfig(), out) == 0) { BhNative.bh_last_error(null) }   // no CPU fallback
        e = out.value
        upload()
    }
    private fun ck(rc: Int) = check(rc == 0) { BhNative.bh_last_error(e) }

    private fun pushConfig() {             // Config is re-read at every use, BarnesHutAlg.kt:256,360-361,378,412
        val p = BhParams()
        ck(BhNative.bh_default_params(Config.WIDTH_PX, Config.HEIGHT_PX, p))
        p.G = Config.G; p.dt = Config.DT; p.theta = Config.theta; p.soft2 = Config.SOFT2
        p.merge_max_mass = mergeMaxMass; p.merge_min_dist = mergeMinDist
        ck(BhNative.bh_set_params(e, p))
    }
    private fun upload() {
        val n = bodies.size
        ck(BhNative.bh_set_bodies(e, n.toLong(), DoubleArray(n) { bodies[it].x }, DoubleArray(n) { bodies[it].y },
            DoubleArray(n) { bodies[it].vx }, DoubleArray(n) { bodies[it].vy }, DoubleArray(n) { bodies[it].m }))
    }
    private fun download() {               // write back into the SAME Body objects the UI holds
        val n = BhNative.bh_num_bodies(e).toInt()
        val x = DoubleArray(n); val y = DoubleArray(n); val vx = DoubleArray(n); val vy = DoubleArray(n); val m = DoubleArray(n)
        val origin = IntArray(n); val nOut = LongByReference()
        ck(BhNative.bh_get_bodies(e, n.toLong(), x, y, vx, vy, m, nOut))
        ck(BhNative.bh_get_origin(e, n.toLong(), origin, nOut))
        val shrunk = n != bodies.size
        if (shrunk) {                      // merge rule removed bodies, BarnesHutAlg.kt:514-520: drop the same objects
            val keep = origin.map { bodies[it] }
            bodies.clear(); bodies.addAll(keep)
        }
        // write the t+dt state (and the grown masses) into the surviving objects FIRST ...
        for (i in 0 until n) { val b = bodies[i]; b.x = x[i]; b.y = y[i]; b.vx = vx[i]; b.vy = vy[i]; b.m = m[i] }
        // ... then tell the engine that origin[] now indexes the shrunk list; nothing is re-uploaded
        if (shrunk) ck(BhNative.bh_rebase_origin(e))
    }

    fun getBodies(): List<Body> = bodies
    fun resetBodies(newBodies: MutableList<Body>) { bodies = newBodies; upload(); lastTree = null }
    fun step() { pushConfig(); ck(BhNative.bh_step(e, 1)); download(); lastTree = null }
    fun getTreeForDebug(): BHTree = lastTree ?: run {
        pushConfig()
        val nc = LongByReference()
        ck(BhNative.bh_get_tree(e, 0, nc, null, null, null, null, null, null, null))
        val k = nc.value.toInt()
        val cx = DoubleArray(k); val cy = DoubleArray(k); val h = DoubleArray(k)
        val ms = DoubleArray(k); val qx = DoubleArray(k); val qy = DoubleArray(k); val bi = IntArray(k)
        ck(BhNative.bh_get_tree(e, k.toLong(), nc, cx, cy, h, ms, qx, qy, bi))
        BHTree(cx, cy, h, ms, qx, qy, bi, bodies).also { lastTree = it }
    }
}
