"""Synthetic RGB-D clouds of the shapes BASELINE.json names.

The reference ships one sample cloud and no generator; these frames stand in for the
organised 640x480 Kinect-style input its CLI loads (src/supervoxel_clustering.cpp:313-340):
a pinhole render of a floor, a back wall and a handful of boxes / spheres / cylinders,
with depth noise, per-object colours, NaN holes and NaN at depth discontinuities.

Output layout = pcl::PointXYZRGBA, 32 bytes per point:
    float x, y, z, 1.0f ; uint8 b, g, r, a ; 12 bytes padding
"""
import numpy as np

POINT_DTYPE = np.dtype({
    "names": ["x", "y", "z", "w", "rgba", "pad0", "pad1", "pad2"],
    "formats": ["<f4", "<f4", "<f4", "<f4", "<u4", "<u4", "<u4", "<u4"],
})
assert POINT_DTYPE.itemsize == 32


def _ray_box(d, lo, hi):
    """Ray (origin 0, direction d[...,3]) against an axis-aligned box; returns t (inf = miss)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        t1 = lo / d
        t2 = hi / d
    tmin = np.minimum(t1, t2).max(axis=-1)
    tmax = np.maximum(t1, t2).min(axis=-1)
    hit = (tmax >= tmin) & (tmax > 0)
    t = np.where(tmin > 0, tmin, tmax)
    return np.where(hit, t, np.inf)


def _ray_sphere(d, c, r):
    a = (d * d).sum(-1)
    b = -2.0 * (d * c).sum(-1)
    cc = (c * c).sum() - r * r
    disc = b * b - 4 * a * cc
    with np.errstate(invalid="ignore"):
        t = (-b - np.sqrt(disc)) / (2 * a)
    return np.where((disc >= 0) & (t > 0), t, np.inf)


def _ray_cyl_y(d, c, r, y0, y1):
    """Cylinder with axis parallel to y through (c[0], *, c[2])."""
    a = d[..., 0] ** 2 + d[..., 2] ** 2
    b = -2.0 * (d[..., 0] * c[0] + d[..., 2] * c[2])
    cc = c[0] ** 2 + c[2] ** 2 - r * r
    disc = b * b - 4 * a * cc
    with np.errstate(invalid="ignore"):
        t = (-b - np.sqrt(disc)) / (2 * a)
    y = t * d[..., 1]
    return np.where((disc >= 0) & (t > 0) & (y >= y0) & (y <= y1), t, np.inf)


def make_frame(seed=20020, width=640, height=480, f=525.0, zmin=0.6, zmax=3.5,
               nan_fraction=0.10, as_struct=True):
    """One organised RGB-D frame.  Returns a (height*width,) array of POINT_DTYPE
    (or, with as_struct=False, (xyz float32 [N,3], rgba uint32 [N]))."""
    rng = np.random.default_rng(seed)
    cx, cy = (width - 1) / 2.0, (height - 1) / 2.0
    scale = 525.0 / f * (width / 640.0)          # keep the field of view when the raster grows
    fx = f * (width / 640.0)
    del scale
    u, v = np.meshgrid(np.arange(width, dtype=np.float64), np.arange(height, dtype=np.float64))
    d = np.stack([(u - cx) / fx, (v - cy) / fx, np.ones_like(u)], axis=-1)

    objs = []   # (t array, colour)
    floor_y = 0.9 + 0.2 * rng.random()
    with np.errstate(divide="ignore"):
        tf = floor_y / d[..., 1]
    tf = np.where(tf > 0, tf, np.inf)
    objs.append((tf, rng.integers(40, 216, 3)))
    wall_z = zmax - 0.3 * rng.random()
    objs.append((np.full(u.shape, wall_z), rng.integers(40, 216, 3)))
    n_obj = int(rng.integers(6, 13))
    for _ in range(n_obj):
        kind = int(rng.integers(0, 3))
        c = np.array([rng.uniform(-1.2, 1.2), 0.0, rng.uniform(zmin + 0.4, wall_z - 0.4)])
        col = rng.integers(20, 236, 3)
        if kind == 0:
            half = rng.uniform(0.08, 0.3, 3)
            c[1] = floor_y - half[1]
            objs.append((_ray_box(d, c - half, c + half), col))
        elif kind == 1:
            r = rng.uniform(0.08, 0.25)
            c[1] = floor_y - r
            objs.append((_ray_sphere(d, c, r), col))
        else:
            r = rng.uniform(0.06, 0.2)
            hgt = rng.uniform(0.2, 0.7)
            objs.append((_ray_cyl_y(d, c, r, floor_y - hgt, floor_y), col))
    tstack = np.stack([o[0] for o in objs], axis=0)
    which = tstack.argmin(axis=0)
    t = tstack.min(axis=0)
    z = t.copy()                                  # d_z == 1
    z = z + rng.normal(0.0, 1.0, z.shape) * 0.001 * z * z
    bad = ~np.isfinite(z) | (z < zmin) | (z > zmax + 0.5)
    # NaN at depth discontinuities (like a structured-light sensor)
    zz = np.where(bad, np.nan, z)
    edge = np.zeros_like(bad)
    with np.errstate(invalid="ignore"):
        edge[:, 1:] |= np.abs(zz[:, 1:] - zz[:, :-1]) > 0.05
        edge[1:, :] |= np.abs(zz[1:, :] - zz[:-1, :]) > 0.05
    holes = rng.random(z.shape) < nan_fraction
    bad |= edge | holes
    x = d[..., 0] * z
    y = d[..., 1] * z
    cols = np.stack([o[1] for o in objs], axis=0)[which].astype(np.float64)
    cols = np.clip(np.rint(cols + rng.normal(0.0, 4.0, cols.shape)), 0, 255).astype(np.uint32)
    rgba = (np.uint32(255) << np.uint32(24)) | (cols[..., 0] << np.uint32(16)) | (cols[..., 1] << np.uint32(8)) | cols[..., 2]
    xyz = np.stack([x, y, z], axis=-1).astype(np.float32)
    xyz[bad] = np.nan
    xyz = xyz.reshape(-1, 3)
    rgba = rgba.reshape(-1).astype(np.uint32)
    if not as_struct:
        return xyz, rgba
    return pack_points(xyz, rgba)


def pack_points(xyz, rgba):
    """(N,3) float32 + (N,) uint32 -> (N,) POINT_DTYPE (pcl::PointXYZRGBA layout)."""
    pts = np.zeros(xyz.shape[0], dtype=POINT_DTYPE)
    pts["x"], pts["y"], pts["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    pts["w"] = 1.0
    pts["rgba"] = rgba
    return pts


def make_dense_scene(seed=40000, width=3652, height=2740):
    """C4: the same scene family at ~10 M rays, z in [0.5, 4] m."""
    return make_frame(seed=seed, width=width, height=height, zmin=0.5, zmax=4.0)


# ---- C5: merged room scan ------------------------------------------------------------
ROOM_SCANS = 40


def _room_surfaces(rng):
    """Axis-aligned rectangles (origin, edge u, edge v, colour) of a 6 x 3 x 8 m room (x, y, z) translated so that
    z stays in [1, 9] m (main() folds z < 0 and the VCCS transform takes ln z), plus box-shaped furniture."""
    x0, x1, y0, y1, z0, z1 = -3.0, 3.0, -1.5, 1.5, 1.0, 9.0
    S = []

    def rect(o, u, v, col):
        S.append((np.array(o, float), np.array(u, float), np.array(v, float), np.array(col, float)))

    rect((x0, y1, z0), (x1 - x0, 0, 0), (0, 0, z1 - z0), rng.integers(60, 200, 3))          # floor
    rect((x0, y0, z0), (x1 - x0, 0, 0), (0, 0, z1 - z0), rng.integers(180, 250, 3))         # ceiling
    rect((x0, y0, z0), (0, y1 - y0, 0), (0, 0, z1 - z0), rng.integers(100, 230, 3))         # walls
    rect((x1, y0, z0), (0, y1 - y0, 0), (0, 0, z1 - z0), rng.integers(100, 230, 3))
    rect((x0, y0, z1), (x1 - x0, 0, 0), (0, y1 - y0, 0), rng.integers(100, 230, 3))
    rect((x0, y0, z0), (x1 - x0, 0, 0), (0, y1 - y0, 0), rng.integers(100, 230, 3))
    for _ in range(18):                                                                     # furniture: boxes on the floor
        hx, hy, hz = rng.uniform(0.2, 0.7), rng.uniform(0.2, 0.9), rng.uniform(0.2, 0.7)
        cx, cz = rng.uniform(x0 + 0.8, x1 - 0.8), rng.uniform(z0 + 0.8, z1 - 0.8)
        lo = np.array([cx - hx, y1 - 2 * hy, cz - hz]); hi = np.array([cx + hx, y1, cz + hz])
        col = rng.integers(20, 236, 3)
        d = hi - lo
        rect(lo, (d[0], 0, 0), (0, 0, d[2]), col)                                           # top
        rect(lo, (d[0], 0, 0), (0, d[1], 0), col); rect((lo[0], lo[1], hi[2]), (d[0], 0, 0), (0, d[1], 0), col)
        rect(lo, (0, d[1], 0), (0, 0, d[2]), col); rect((hi[0], lo[1], lo[2]), (0, d[1], 0), (0, 0, d[2]), col)
    return S


def make_room_scan(seed=50000, n_points=50_000_000, scans=None):
    """C5: a room scanned from ROOM_SCANS positions and merged; `scans` (an iterable of scan numbers) selects the
    share of one rank -- the cloud of all scans in ascending order is the concatenation of the shares.
    Every scan samples the surfaces with a density falling off with the distance from its position, adds 2 mm range
    noise and per-point colour noise.  Returns POINT_DTYPE records (no NaNs: merged scans are unorganised)."""
    rng0 = np.random.default_rng(seed)
    S = _room_surfaces(rng0)
    area = np.array([np.linalg.norm(np.cross(u, v)) for _, u, v, _ in S])
    pos = np.stack([rng0.uniform(-2.2, 2.2, ROOM_SCANS), rng0.uniform(-0.3, 0.6, ROOM_SCANS), rng0.uniform(1.8, 8.2, ROOM_SCANS)], 1)
    per_scan = n_points // ROOM_SCANS
    scans = range(ROOM_SCANS) if scans is None else scans
    out = []
    for s in scans:
        rng = np.random.default_rng(seed + 1 + s)
        n = per_scan + (n_points - per_scan * ROOM_SCANS if s == ROOM_SCANS - 1 else 0)
        # importance: surface area / squared distance of its centre from the scanner
        ctr = np.stack([o + 0.5 * u + 0.5 * v for o, u, v, _ in S])
        w = area / np.maximum(0.25, ((ctr - pos[s]) ** 2).sum(1))
        which = rng.choice(len(S), size=n, p=w / w.sum())
        O = np.stack([q[0] for q in S])[which]; U = np.stack([q[1] for q in S])[which]; Vv = np.stack([q[2] for q in S])[which]
        a = rng.random((n, 1)); b = rng.random((n, 1))
        p = O + a * U + b * Vv
        ray = p - pos[s]
        rl = np.linalg.norm(ray, axis=1, keepdims=True)
        p = p + ray / np.maximum(rl, 1e-6) * rng.normal(0.0, 0.002, (n, 1))
        p[:, 2] = np.clip(p[:, 2], 0.9, 9.2)
        col = np.stack([q[3] for q in S])[which]
        col = np.clip(np.rint(col + rng.normal(0.0, 4.0, col.shape)), 0, 255).astype(np.uint32)
        rgba = (np.uint32(255) << np.uint32(24)) | (col[:, 0] << np.uint32(16)) | (col[:, 1] << np.uint32(8)) | col[:, 2]
        out.append(pack_points(p.astype(np.float32), rgba.astype(np.uint32)))
    if not out:
        return np.zeros(0, POINT_DTYPE)
    return np.concatenate(out)


def room_scan_share(rank, world):
    """Scan numbers of rank `rank`: contiguous, so that the shares concatenate to the full cloud in scan order."""
    return range(ROOM_SCANS * rank // world, ROOM_SCANS * (rank + 1) // world)
