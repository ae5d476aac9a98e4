"""Synthetic scenes: the *distributions* of the reference's BodyFactory.kt.

The reference draws from kotlin.random.Random (XorWow) and JDK libm, neither of which
is reproducible here bit-for-bit, so these generators reproduce the
sampling laws with numpy's PCG64 and are exchanged with every engine as explicit SoA
arrays — "identical inputs" is literal.  Paths cited as BF.kt are
/root/reference/src/main/kotlin/BodyFactory.kt; NP.kt is NBodyPanel.kt.
"""
from __future__ import annotations

import numpy as np

G_DEFAULT = 80.0        # Config.kt:11
MIN_R = 8.0             # Config.kt:35
CENTRAL_MASS = 50_000.0
TOTAL_SATELLITE_MASS = 5_000.0


def _stack(*parts):
    return tuple(np.concatenate([p[k] for p in parts]) for k in range(5))


def snap_f32(scene):
    """Round positions to FP32-representable doubles (parity inputs)."""
    x, y, vx, vy, m = scene
    return (x.astype(np.float32).astype(np.float64), y.astype(np.float32).astype(np.float64), vx, vy, m)


def make_uniform_random(n, m, width=2400, height=800, seed=3):
    """BF.kt:160-177 makeUniformRandom: x~U[0,W), y~U[0,H), v=0, equal masses."""
    if n <= 0 or m <= 0.0:
        z = np.zeros(0)
        return (z, z.copy(), z.copy(), z.copy(), z.copy())
    rng = np.random.Generator(np.random.PCG64(seed))
    x = rng.random(n) * float(width)
    y = rng.random(n) * float(height)
    return (x, y, np.zeros(n), np.zeros(n), np.full(n, float(m)))


def _circular_velocities(x, y, m, cx, cy, G, speed_jitter, clockwise, rng, vx0, vy0):
    """BF.kt:118-147: exact enclosed mass by sorted radius, v_circ = sqrt(G M_enc / R)."""
    n = x.shape[0]
    r = np.hypot(x - cx, y - cy)
    order = np.argsort(r, kind="stable")
    menc = np.empty(n)
    menc[order] = np.cumsum(m[order])
    vx = np.zeros(n)
    vy = np.zeros(n)
    dx, dy = x[1:] - cx, y[1:] - cy
    R = np.maximum(1e-6, np.hypot(dx, dy))
    v = np.sqrt(G * menc[1:] / R) * (1.0 + (rng.random(n - 1) - 0.5) * 2.0 * speed_jitter)
    tx, ty = (dy / R, -dx / R) if clockwise else (-dy / R, dx / R)
    vx[1:], vy[1:] = tx * v, ty * v
    vx[1:] += vx0
    vy[1:] += vy0
    vx[0], vy[0] = vx0, vy0
    return vx, vy


def make_galaxy_disk(n_total, x=1200.0, y=400.0, r=200.0, min_r=MIN_R, central_mass=CENTRAL_MASS,
                     total_satellite_mass=TOTAL_SATELLITE_MASS, vx=0.0, vy=0.0, eps_m2=0.03, phi0=0.0,
                     speed_jitter=0.01, clockwise=True, G=G_DEFAULT, seed=1):
    """BF.kt:63-150 makeGalaxyDisk: truncated-exponential radius on [minR, r] with Rd = r/3
    (BF.kt:97-102), uniform angle, m=2 bar tapered by exp(-(R/0.6r)^2) (BF.kt:105-116)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    sats = max(n_total - 1, 0)
    m_sat = total_satellite_mass / sats if sats > 0 else 0.0
    Rd, taper_r = r / 3.0, r * 0.6
    u = rng.random(sats)
    A = np.exp(-(r - min_r) / Rd)
    R = min_r - Rd * np.log(1 - u * (1 - A))
    th = rng.random(sats) * 2.0 * np.pi
    R2 = R * (1.0 + eps_m2 * np.cos(2.0 * (th - phi0)) * np.exp(-(R / taper_r) ** 2))
    px = np.concatenate([[x], x + R2 * np.cos(th)])
    py = np.concatenate([[y], y + R2 * np.sin(th)])
    m = np.concatenate([[central_mass], np.full(sats, m_sat)])
    if sats > 0:
        pvx, pvy = _circular_velocities(px, py, m, x, y, G, speed_jitter, clockwise, rng, vx, vy)
    else:
        pvx, pvy = np.array([vx], float), np.array([vy], float)
    return (px, py, pvx, pvy, m)


def make_kepler_disk(n_total, x=1200.0, y=400.0, r=304.0, vx=0.0, vy=0.0, radial_jitter=0.03,
                     speed_jitter=0.01, clockwise=True, G=G_DEFAULT, seed=3):
    """BF.kt:11-61 makeKeplerDisk: uniform-in-area radius on [MIN_R, r] with +-3% jitter."""
    rng = np.random.Generator(np.random.PCG64(seed))
    sats = max(n_total - 1, 0)
    m_sat = TOTAL_SATELLITE_MASS / sats if sats > 0 else 0.0
    u = rng.random(sats)
    rr = np.sqrt(u * (r * r - MIN_R * MIN_R) + MIN_R * MIN_R)
    rj = rr * (1.0 + (rng.random(sats) - 0.5) * 2.0 * radial_jitter)
    ang = rng.random(sats) * 2.0 * np.pi
    px = np.concatenate([[x], x + rj * np.cos(ang)])
    py = np.concatenate([[y], y + rj * np.sin(ang)])
    m = np.concatenate([[CENTRAL_MASS], np.full(sats, m_sat)])
    if sats > 0:
        pvx, pvy = _circular_velocities(px, py, m, x, y, G, speed_jitter, clockwise, rng, vx, vy)
    else:
        pvx, pvy = np.array([vx], float), np.array([vy], float)
    return (px, py, pvx, pvy, m)


def default_two_disks(width=2400, height=800, n1=10_000, n2=2_500, scale=1.0, seed=1):
    """NP.kt:83-99 defaultBodies(): disk A 10,000 bodies r=300 (50,000 + 5,000) at the window
    centre, disk B 2,500 bodies r=100 (5,000 + 500) at y=0.2H drifting vx=-50.  `scale`
    multiplies the radii (constant surface density when n grows as scale^2)."""
    a = make_galaxy_disk(n1, x=width * 0.5, y=height * 0.5, r=300.0 * scale, central_mass=50_000.0,
                         total_satellite_mass=5_000.0, seed=seed)
    b = make_galaxy_disk(n2, x=width * 0.5, y=height * 0.2, vx=-50.0, r=100.0 * scale, central_mass=5_000.0,
                         total_satellite_mass=500.0, seed=seed + 1)
    return _stack(a, b)


def disk_collision(n_each, width, height, radius, separation, closing_v=50.0, seed=6):
    """C4-style: two equal disks approaching along x."""
    cx, cy = width * 0.5, height * 0.5
    a = make_galaxy_disk(n_each, x=cx - separation * 0.5, y=cy, r=radius, vx=+closing_v, seed=seed)
    b = make_galaxy_disk(n_each, x=cx + separation * 0.5, y=cy, r=radius, vx=-closing_v, seed=seed + 1)
    return _stack(a, b)


def mixed_mass_stress(n_disks=8, n_per_disk=1000, n_black_holes=16, width=65536, height=65536, seed=8):
    """C5-style: several disks plus lone 50,000-mass bodies (RMB 'black hole', NP.kt:171)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    parts = []
    radius = 300.0 * np.sqrt(n_per_disk / 10_000.0)
    for d in range(n_disks):
        cx = (0.15 + 0.7 * rng.random()) * width
        cy = (0.15 + 0.7 * rng.random()) * height
        ang = rng.random() * 2 * np.pi
        v = 50.0 * rng.random()
        parts.append(make_galaxy_disk(n_per_disk, x=cx, y=cy, r=radius, vx=v * np.cos(ang), vy=v * np.sin(ang),
                                      seed=seed * 100 + d))
    bx = (0.1 + 0.8 * rng.random(n_black_holes)) * width
    by = (0.1 + 0.8 * rng.random(n_black_holes)) * height
    z = np.zeros(n_black_holes)
    parts.append((bx, by, z, z.copy(), np.full(n_black_holes, CENTRAL_MASS)))
    return _stack(*parts)
