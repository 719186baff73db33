"""Synthetic KITTI-shaped scan pairs: there is no dataset on the GPU box.

A procedural street scene (ground plane, axis-aligned boxes = buildings and cars, vertical cylinders = poles and
trunks) is ray-cast from two sensor poses with an HDL-64-like pattern, range noise and ray dropout are added, and the
returns are voxel-barycentre downsampled at 0.3 m - what preporcess/downsample_pcd_kitti.py:28 does offline for the
reference. Pure numpy, seeded, deterministic: seed = 7351 + pair_id (experiments/config.py:13 is the reference seed).
The second scan is the same scene seen from a pose 8-12 m further along the street (yaw U(-10,10) deg, roll/pitch
N(0,0.5 deg)), as KITTI pairs are >= 10 m apart (rdmnet/datasets/registration/kitti/dataset.py:106); it is the REFERENCE
scan of the pair and the first one the SOURCE (see make_pair for why the direction matters).
"""
import numpy as np

SENSOR_HEIGHT = 1.73


def _rot(yaw, pitch, roll):
    cy, sy, cp, sp, cr, sr = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch), np.cos(roll), np.sin(roll)
    rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
    ry = np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
    rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
    return rz @ ry @ rx


def _lattice(ix, iy, iz, seed):
    h = (ix * 73856093) ^ (iy * 19349663) ^ (iz * 83492791) ^ (seed * 2654435761)
    h = (h ^ (h >> 13)) * 1274126177
    return ((h ^ (h >> 16)) & 0xFFFFFF).astype(np.float64) / float(0xFFFFFF) * 2.0 - 1.0


def value_noise(p, scale, seed):
    """Smooth procedural noise in [-1, 1]: a pure function of the WORLD position (trilinear value noise on a hashed
    lattice), so that both scans of a pair see the same surface relief."""
    q = p / scale
    i = np.floor(q).astype(np.int64)
    f = q - i
    f = f * f * (3.0 - 2.0 * f)
    out = np.zeros(p.shape[0])
    for dx in (0, 1):
        for dy in (0, 1):
            for dz in (0, 1):
                w = (f[:, 0] if dx else 1 - f[:, 0]) * (f[:, 1] if dy else 1 - f[:, 1]) * (f[:, 2] if dz else 1 - f[:, 2])
                out += w * _lattice(i[:, 0] + dx, i[:, 1] + dy, i[:, 2] + dz, seed)
    return out


def make_scene(rng):
    """Boxes (lo, hi), cylinders (cx, cy, r, z0, z1) and spheres (cx, cy, cz, r) in a 160 m x 100 m street corridor;
    world frame = first sensor frame, ground at z = -1.73. Two rows of facades line the street, cars stand on it,
    poles/trunks and tree crowns fill the verges - dense enough that the coarse pyramid levels see KITTI-like
    neighbour counts."""
    lo, hi = [], []
    for side in (-1.0, 1.0):
        for row, (y0lo, y0hi) in enumerate([(9.0, 14.0), (30.0, 38.0)]):
            x = -80.0 + rng.uniform(0, 5)
            while x < 80.0:
                w, dep, hgt = rng.uniform(8, 20), rng.uniform(8, 15), rng.uniform(5, 15)
                y0 = rng.uniform(y0lo, y0hi)
                lo.append([x, min(side * y0, side * (y0 + dep)), -SENSOR_HEIGHT])
                hi.append([x + w, max(side * y0, side * (y0 + dep)), -SENSOR_HEIGHT + hgt])
                x += w + rng.uniform(0, 6) * (1 + row)
    n_car = int(rng.integers(20, 41))
    for _ in range(n_car):
        cx, cy = rng.uniform(-78, 78), rng.choice([-1.0, 1.0]) * rng.uniform(2.5, 7.5)
        if abs(cx) < 16 and abs(cy) < 4:  # keep both sensor poses clear
            cy = 6.0 * np.sign(cy)
        lo.append([cx - 2.0, cy - 0.9, -SENSOR_HEIGHT])
        hi.append([cx + 2.0, cy + 0.9, -SENSOR_HEIGHT + 1.5])
    cyl, sph = [], []
    for i in range(70):
        cx, cy = rng.uniform(-78, 78), rng.choice([-1.0, 1.0]) * rng.uniform(5.0, 8.5)
        h = rng.uniform(3, 8)
        cyl.append([cx, cy, rng.uniform(0.15, 0.4), -SENSOR_HEIGHT, -SENSOR_HEIGHT + h])
        if i % 5 != 0:  # a crown on most trunks
            sph.append([cx, cy, -SENSOR_HEIGHT + h, rng.uniform(1.5, 3.0)])
    return np.array(lo), np.array(hi), np.array(cyl), np.array(sph)


def ray_cast(origin, R, scene, n_elev, n_azim, rng, max_range=80.0, noise=0.02, dropout=0.1, relief=True, seed=0,
             ground_range=None):
    """Returns the hit points in the SENSOR frame (float64, (n,3)). With `relief`, every analytic surface carries a
    view-independent procedural relief (value_noise of the world hit position, displaced along the surface normal):
    facade articulation (balcony / window scale, 1.2 m and 0.4 m octaves), rolling ground with a rougher fine octave,
    lumpy tree crowns and car bodies, so that the scene is not a set of featureless planes."""
    lo, hi, cyl, sph = scene
    el = np.deg2rad(np.linspace(-24.8, 2.0, n_elev))
    az = np.linspace(-np.pi, np.pi, n_azim, endpoint=False)
    ce, se = np.cos(el)[:, None], np.sin(el)[:, None]
    d_s = np.stack([ce * np.cos(az)[None], ce * np.sin(az)[None], np.broadcast_to(se, (n_elev, n_azim))], -1).reshape(-1, 3)
    d = d_s @ R.T  # world-frame directions
    o = origin
    n = d.shape[0]
    t = np.full(n, np.inf)
    kind = np.zeros(n, np.int8)        # 0 ground, 1 box, 2 cylinder, 3 sphere
    nrm = np.zeros((n, 3))
    nrm[:, 2] = 1.0
    amp = np.full(n, 0.12)             # relief amplitude (m)
    # ground plane z = -1.73 (world)
    with np.errstate(divide="ignore", invalid="ignore"):
        tg = (-SENSOR_HEIGHT - o[2]) / d[:, 2]
    tg[~(tg > 0)] = np.inf
    t = np.minimum(t, tg)
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / d
        for b in range(lo.shape[0]):
            t0 = (lo[b] - o) * inv
            t1 = (hi[b] - o) * inv
            tmin = np.minimum(t0, t1)
            tn = tmin.max(1)
            tf = np.maximum(t0, t1).min(1)
            hit = (tf >= tn) & (tn > 0) & (tn < t)
            if hit.any():
                ax = tmin[hit].argmax(1)
                t[hit] = tn[hit]
                kind[hit] = 1
                nb = np.zeros((int(hit.sum()), 3))
                nb[np.arange(nb.shape[0]), ax] = -np.sign(d[hit][np.arange(nb.shape[0]), ax])
                nrm[hit] = nb
                amp[hit] = 0.35 if (hi[b, 2] - lo[b, 2]) > 2.0 else 0.12  # buildings vs cars
        a = d[:, 0] ** 2 + d[:, 1] ** 2
        for c in cyl:
            ox, oy = o[0] - c[0], o[1] - c[1]
            bq = ox * d[:, 0] + oy * d[:, 1]
            cq = ox * ox + oy * oy - c[2] ** 2
            disc = bq * bq - a * cq
            tc = (-bq - np.sqrt(np.maximum(disc, 0))) / a
            z = o[2] + tc * d[:, 2]
            hit = (disc > 0) & (tc > 0) & (z >= c[3]) & (z <= c[4]) & (tc < t)
            t[hit] = tc[hit]
            kind[hit] = 2
            amp[hit] = 0.0
        for c in sph:
            oc = o - c[:3]
            bq = d @ oc
            disc = bq * bq - (oc @ oc - c[3] ** 2)
            tc = -bq - np.sqrt(np.maximum(disc, 0))
            hit = (disc > 0) & (tc > 0) & (tc < t)
            if hit.any():
                t[hit] = tc[hit]
                kind[hit] = 3
                w = o + d[hit] * tc[hit, None] - c[:3]
                nrm[hit] = w / np.linalg.norm(w, axis=1, keepdims=True)
                amp[hit] = 0.5
    keep = np.isfinite(t) & (t < max_range) & (rng.random(n) >= dropout)
    # optional: grazing returns from the road fade with range (a real sensor loses most ground returns beyond ~30-40 m)
    if ground_range is not None:
        p_ground = np.clip(1.5 - t / ground_range, 0.03, 1.0)
        keep &= (kind != 0) | (rng.random(n) < p_ground)
    t = t + rng.normal(0.0, noise, n)
    world = o + d[keep] * t[keep, None]
    if relief:
        h = 0.7 * value_noise(world, 1.2, seed) + 0.3 * value_noise(world, 0.4, seed + 1)
        g = kind[keep] == 0
        h[g] = 0.8 * value_noise(world[g], 6.0, seed + 2) + 0.2 * value_noise(world[g], 0.8, seed + 3)  # rolling ground
        world = world + nrm[keep] * (amp[keep] * h)[:, None]
    pts = (world - o) @ R  # sensor frame: p_s = R^T (w - o)
    return pts


def voxel_downsample(points, voxel):
    """Voxel-barycentre downsample (what open3d.voxel_down_sample computes), output ordered by voxel key."""
    key = np.floor(points / voxel).astype(np.int64)
    key -= key.min(0)
    dims = key.max(0) + 1
    flat = (key[:, 0] * dims[1] + key[:, 1]) * dims[2] + key[:, 2]
    uniq, inv = np.unique(flat, return_inverse=True)
    cnt = np.bincount(inv, minlength=uniq.shape[0]).astype(np.float64)
    out = np.stack([np.bincount(inv, weights=points[:, a], minlength=uniq.shape[0]) / cnt for a in range(3)], 1)
    return out


def make_pair(pair_id=0, n_elev=112, n_azim=2000, voxel=0.3):
    """-> dict(ref_points (Nr,3) f32, src_points (Ns,3) f32, transform (4,4) f32 with ref = T * src).

    Default cast = the "16k" size class: 112 elevation rows x 2000 azimuths. The literal HDL-64 pattern (64 rows) on
    this procedural scene leaves only 11.5k-12.7k points per scan after the 0.3 m voxel filter - the smooth facades and
    the flat ground merge far more returns per voxel than real KITTI geometry does (the bundled KITTI scans keep
    18.6k-20.5k) - so the row count is raised until the post-filter size is the ~16k (+-3k) that BASELINE.json's
    config names: 14.4k-16.7k per scan, per-stage mean valid-neighbour counts 44.8 / 45.5 / 44.7 / 45.2 / 52.0 against
    39.3 / 41.0 / 49.4 / 53.0 / 62.0 for the bundled scans (SURVEY 8(d) gate: within +-20 %)."""
    rng = np.random.default_rng(7351 + pair_id)
    scene = make_scene(rng)
    ref = ray_cast(np.zeros(3), np.eye(3), scene, n_elev, n_azim, rng, seed=7351 + pair_id)
    yaw = np.deg2rad(rng.uniform(-10, 10))
    pitch, roll = np.deg2rad(rng.normal(0, 0.5, 2))
    R = _rot(yaw, pitch, roll)
    tr = np.array([rng.uniform(8, 12), rng.normal(0, 0.5), 0.0])
    src = ray_cast(tr, R, scene, n_elev, n_azim, rng, seed=7351 + pair_id)
    T = np.eye(4)
    T[:3, :3], T[:3, 3] = R, tr  # first scan = R * second scan + tr
    rng2 = np.random.default_rng(99991 + pair_id)
    a_d, b_d = voxel_downsample(ref, voxel), voxel_downsample(src, voxel)
    a_d = a_d[rng2.permutation(a_d.shape[0])]  # sensor files are not voxel-sorted
    b_d = b_d[rng2.permutation(b_d.shape[0])]
    # Pair direction = the reference's KITTI convention: the SOURCE scan is the one taken ~10 m further BACK along the
    # driving direction, i.e. estimated_transform (src -> ref) has t_x ~ -10 m, as on both bundled pairs (-10.28 / -10.24 m,
    # SURVEY 8(c)). The pretrained network has absorbed that prior (absolute-coordinate rotary embedding + vote offsets):
    # with the roles the other way round (t_x ~ +10 m) the coarse matching finds no correct node pair on any two-view
    # pair of this scene (0 % inliers, the round-1 bench regime), while in this direction the CPU oracle registers pairs
    # 0-2 at RRE 0.09-0.21 deg, RTE 2-7 cm with ~3400 correspondences, 66-70 % of them inliers.
    # So: ref = the SECOND scan (sensor 8-12 m ahead), src = the first one, transform = inverse of the cast pose.
    return {"ref_points": b_d.astype(np.float32), "src_points": a_d.astype(np.float32),
            "transform": np.linalg.inv(T).astype(np.float32)}


# config 5 size classes: target post-voxel points per scan -> (n_elev, n_azim)
# (calibrated on pair ids 0-2: mean post-filter points per scan 4.2k / 7.8k / 15.5k / 21.1k. The visible surface of the 160 m
# corridor saturates near 21k points per scan at 0.3 m, so the largest class is what the scene can give, not 32k.)
SIZE_CLASSES = {"4k": (24, 360), "8k": (48, 600), "16k": (112, 2000), "32k": (192, 4000)}
