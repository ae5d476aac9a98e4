// HeadlessDump.kt — runs the REFERENCE's own PhysicsEngine (BarnesHutAlg.kt, unmodified) on a
// binary scene and dumps the result, so that the C++ oracle of this repository can be checked
// bit-for-bit against the JVM original (INTEGRATION.md, "Pinning the oracle").  Copy into the
// reference's src/main/kotlin/ next to BarnesHutAlg.kt; uses only the reference's classes.
//
// input  (little endian): int64 n | int32 W | int32 H | f64 theta | f64 G | f64 dt | int32 steps |
//                         int32 merge (0 = mergeMinDist 0.0) | x[n] y[n] vx[n] vy[n] m[n]   (f64)
// output (little endian): int64 n_out | x y vx vy m (f64 each) | int64 n_cells | cx cy h (f64 each)
import java.io.File
import java.nio.ByteBuffer
import java.nio.ByteOrder

fun main(args: Array<String>) {
    val inp = ByteBuffer.wrap(File(args[0]).readBytes()).order(ByteOrder.LITTLE_ENDIAN)
    val n = inp.long.toInt()
    Config.WIDTH_PX = inp.int
    Config.HEIGHT_PX = inp.int
    Config.theta = inp.double
    Config.G = inp.double
    Config.DT = inp.double
    val steps = inp.int
    val merge = inp.int
    val cols = Array(5) { DoubleArray(n) { inp.double } }
    val bodies = MutableList(n) { i -> Body(cols[0][i], cols[1][i], cols[2][i], cols[3][i], cols[4][i]) }
    val engine = PhysicsEngine(bodies)
    if (merge == 0) engine.mergeMinDist = 0.0
    repeat(steps) { engine.step() }
    val out = engine.getBodies()
    val quads = ArrayList<Quad>()
    engine.getTreeForDebug().visitQuads { quads.add(it) }
    val buf = ByteBuffer.allocate(16 + out.size * 40 + quads.size * 24).order(ByteOrder.LITTLE_ENDIAN)
    buf.putLong(out.size.toLong())
    for (b in out) buf.putDouble(b.x)
    for (b in out) buf.putDouble(b.y)
    for (b in out) buf.putDouble(b.vx)
    for (b in out) buf.putDouble(b.vy)
    for (b in out) buf.putDouble(b.m)
    buf.putLong(quads.size.toLong())
    for (q in quads) buf.putDouble(q.cx)
    for (q in quads) buf.putDouble(q.cy)
    for (q in quads) buf.putDouble(q.h)
    File(args[1]).writeBytes(buf.array())
}
