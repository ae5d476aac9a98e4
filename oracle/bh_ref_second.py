"""bh_ref_second.py — a SECOND, independent restatement of the reference's physics step
(/root/reference/src/main/kotlin/BarnesHutAlg.kt), in plain Python objects and loops.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rule as bh_ref.cpp: only tests/ may import it).

Why it exists: the oracle proper (bh_ref.cpp) is PARITY UNPINNED — the reference has no tests or fixtures and no
JVM runs in the build container.  What can be done without a JVM is double-entry bookkeeping: this file was
written from the Kotlin source again, object for object (heap nodes with four child references, recursion,
bodies mutated in place, a list that shrinks in the merge rule), sharing no code and no data layout with the
C++ restatement (arrays, indices, threads).  tests/test_oracle_second_reading.py demands that the two agree
BIT FOR BIT on every f64 of the state, every acceleration and every visitQuads cell, on small scenes that reach
the awkward branches (jitter, out-of-box, zero mass, merges with index shifts).  A transcription slip in either
reading shows up as a difference; an error of understanding shared by both does not — the pin against the JVM
itself stays open (tests/golden/jvm_roundtrip.py + integration/kotlin/HeadlessDump.kt).

Python floats are IEEE binary64, `a*b + c` is never fused, math.sqrt is correctly rounded: the same arithmetic
as a strict-FP JVM.  Only small cases: pure-Python loops.  "BH.kt:a-b" = the lines a function follows."""
import math
import struct


class Config:                      # Config.kt:5-23 (the values the step reads)
    WIDTH_PX = 2400
    HEIGHT_PX = 800
    G = 80.0
    DT = 0.005
    SOFT2 = 1.0 * 1.0
    theta = 0.30
    MIN_R = 8.0


def _bits(v):
    return struct.unpack("<q", struct.pack("<d", v))[0]


def _div(a, b):
    """IEEE division (Python raises on a zero divisor; the JVM returns inf / NaN)."""
    if b != 0.0:
        return a / b
    if a != a or a == 0.0:
        return math.nan
    neg = (math.copysign(1.0, a) < 0.0) != (math.copysign(1.0, b) < 0.0)
    return -math.inf if neg else math.inf


class Body:                        # BH.kt:21-25
    __slots__ = ("x", "y", "vx", "vy", "m")

    def __init__(self, x, y, vx, vy, m):
        self.x, self.y, self.vx, self.vy, self.m = float(x), float(y), float(vx), float(vy), float(m)

    def same_values(self, o):      # data-class equals(): what MutableList.indexOf compares with (BH.kt:522)
        return all(_dcmp(a, b) for a, b in ((self.x, o.x), (self.y, o.y), (self.vx, o.vx), (self.vy, o.vy), (self.m, o.m)))


def _dcmp(a, b):                   # java.lang.Double.compare(a, b) == 0: NaN equals NaN, 0.0 differs from -0.0
    if a != a and b != b:
        return True
    return _bits(a) == _bits(b)


class Acc:                         # BH.kt:33-41
    __slots__ = ("fx", "fy")

    def __init__(self):
        self.fx = 0.0
        self.fy = 0.0

    def reset(self):
        self.fx = 0.0
        self.fy = 0.0


class Quad:                        # BH.kt:53-82
    __slots__ = ("cx", "cy", "h")

    def __init__(self, cx, cy, h):
        self.cx, self.cy, self.h = cx, cy, h

    def contains(self, b):         # BH.kt:61-62
        return b.x >= self.cx - self.h and b.x < self.cx + self.h and b.y >= self.cy - self.h and b.y < self.cy + self.h

    def child(self, which):        # BH.kt:73-81
        hh = self.h / 2.0
        if which == 0:
            return Quad(self.cx - hh, self.cy - hh, hh)
        if which == 1:
            return Quad(self.cx + hh, self.cy - hh, hh)
        if which == 2:
            return Quad(self.cx - hh, self.cy + hh, hh)
        return Quad(self.cx + hh, self.cy + hh, hh)


class BHTree:                      # BH.kt:95-275
    __slots__ = ("quad", "body", "children", "mass", "comX", "comY")

    def __init__(self, quad):
        self.quad = quad
        self.body = None
        self.children = None
        self.mass = 0.0
        self.comX = 0.0
        self.comY = 0.0

    def is_leaf(self):             # BH.kt:112
        return self.children is None

    def insert(self, b):           # BH.kt:125-137
        if not self.quad.contains(b):
            return
        if self.body is None and self.is_leaf():
            self.body = b
            return
        if self.is_leaf():
            self.subdivide()
        existing = self.body
        if existing is not None:
            self.body = None
            self.insert_into_child(existing)
        self.insert_into_child(b)

    def insert_into_child(self, b):   # BH.kt:145-156
        if self.quad.h < 1e-3:
            eps = 1e-3
            b.x += +eps if (_bits(b.x) & 1) == 0 else -eps
            b.y += -eps if (_bits(b.y) & 1) == 0 else +eps
        ix = 0 if b.x < self.quad.cx else 1
        iy = 0 if b.y < self.quad.cy else 2
        self.children[ix + iy].insert(b)

    def subdivide(self):           # BH.kt:159-166
        self.children = [BHTree(self.quad.child(0)), BHTree(self.quad.child(1)), BHTree(self.quad.child(2)), BHTree(self.quad.child(3))]

    def compute_mass(self):        # BH.kt:173-202
        if self.is_leaf():
            if self.body is not None:
                self.mass = self.body.m
                self.comX = self.body.x
                self.comY = self.body.y
            else:
                self.mass = 0.0
                self.comX = self.quad.cx
                self.comY = self.quad.cy
            return
        m_sum = 0.0
        cx = 0.0
        cy = 0.0
        for c in self.children:    # 0, 1, 2, 3 (BH.kt:189-192)
            c.compute_mass()
            if c.mass > 0.0:
                m_sum += c.mass
                cx += c.comX * c.mass
                cy += c.comY * c.mass
        self.mass = m_sum
        if m_sum > 0.0:
            self.comX = cx / m_sum
            self.comY = cy / m_sum
        else:
            self.comX = self.quad.cx
            self.comY = self.quad.cy

    def accumulate_force(self, b, theta2, acc, stats=None):   # BH.kt:215-239
        if self.mass == 0.0:
            return
        if self.is_leaf():
            single = self.body
            if single is None or single is b:
                return
            point_force_acc(b, self.comX, self.comY, self.mass, acc)
            if stats is not None:
                stats[0] += 1
            return
        dx = self.comX - b.x
        dy = self.comY - b.y
        dist2 = dx * dx + dy * dy + Config.SOFT2
        side = self.quad.h * 2.0
        s2 = side * side
        if s2 < theta2 * dist2:
            point_force_acc(b, self.comX, self.comY, self.mass, acc)
            if stats is not None:
                stats[0] += 1
        else:
            if stats is not None:
                stats[1] += 1
            for c in self.children:
                c.accumulate_force(b, theta2, acc, stats)

    def visit_quads(self, visit):  # BH.kt:265-274 (the node itself is passed too: the test reads mass / COM / leaf body)
        visit(self)
        if self.children is not None:
            for c in self.children:
                c.visit_quads(visit)


def point_force_acc(b, px, py, m, acc):   # BH.kt:250-259
    dx = px - b.x
    dy = py - b.y
    r2 = dx * dx + dy * dy + Config.SOFT2
    inv_r = 1.0 / math.sqrt(r2)
    inv_r2 = 1.0 / r2
    f = Config.G * b.m * m * inv_r2
    acc.fx += f * dx * inv_r
    acc.fy += f * dy * inv_r


class PhysicsEngine:               # BH.kt:287-532
    def __init__(self, bodies):
        self.bodies = bodies
        self.ax = [0.0] * len(bodies)
        self.ay = [0.0] * len(bodies)
        self.last_tree = None
        self.mergeMaxMass = 4000.0           # BH.kt:315
        self.mergeMinDist = Config.MIN_R     # BH.kt:321
        self.interactions = 0                # (observers: not in the reference)
        self.opened = 0

    def build_tree(self):          # BH.kt:359-366
        half = max(Config.WIDTH_PX, Config.HEIGHT_PX) / 2.0 + 2.0
        root = BHTree(Quad(Config.WIDTH_PX / 2.0, Config.HEIGHT_PX / 2.0, half))
        for b in self.bodies:
            root.insert(b)
        root.compute_mass()
        return root

    def compute_accelerations(self, root):   # BH.kt:374-395 (worker scheduling does not touch the arithmetic)
        theta2 = Config.theta * Config.theta
        n = len(self.bodies)
        if len(self.ax) < n:
            self.ax = [0.0] * n
            self.ay = [0.0] * n
        acc = Acc()
        stats = [0, 0]
        for i in range(n):
            b = self.bodies[i]
            acc.reset()
            root.accumulate_force(b, theta2, acc, stats)
            self.ax[i] = _div(acc.fx, b.m)
            self.ay[i] = _div(acc.fy, b.m)
        self.interactions, self.opened = stats

    def step(self):                # BH.kt:405-439
        root = self.build_tree()
        self.compute_accelerations(root)
        bs = self.bodies
        dt_half = Config.DT * 0.5
        for i in range(len(bs)):
            bs[i].vx += self.ax[i] * dt_half
            bs[i].vy += self.ay[i] * dt_half
        for b in bs:
            b.x += b.vx * Config.DT
            b.y += b.vy * Config.DT
        root = self.build_tree()
        self.compute_accelerations(root)
        for i in range(len(bs)):
            bs[i].vx += self.ax[i] * dt_half
            bs[i].vy += self.ay[i] * dt_half
        self.last_tree = root
        self.merge_close_bodies_if_needed()

    def merge_close_bodies_if_needed(self):   # BH.kt:463-532
        if self.mergeMinDist <= 0.0 or len(self.bodies) <= 1:
            return
        min_d2 = self.mergeMinDist * self.mergeMinDist
        bodies = self.bodies
        i = 0
        while i < len(bodies):
            bi = bodies[i]
            if bi.m > self.mergeMaxMass:
                n = len(bodies)
                if n > 1:
                    victims = []
                    for j in range(n):                     # the chunks of BH.kt:484-508, concatenated in order
                        if j != i:
                            bj = bodies[j]
                            dx = bj.x - bi.x
                            dy = bj.y - bi.y
                            if dx * dx + dy * dy < min_d2:
                                victims.append(j)
                    if victims:
                        for j in sorted(victims, reverse=True):
                            if j < 0 or j >= len(bodies):
                                continue
                            if bodies[j] is bi:
                                continue
                            bj = bodies[j]
                            bi.m += bj.m
                            del bodies[j]
                        new_index = -1
                        for k, b in enumerate(bodies):     # indexOf: first element EQUAL to bi (data-class equality)
                            if b.same_values(bi):
                                new_index = k
                                break
                        i = new_index if new_index >= 0 else max(i - 1, 0)
                        self.last_tree = None
            i += 1
