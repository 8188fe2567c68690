"""Synthetic Waymo-shape point clouds (no dataset is reachable from the build or bench boxes).

``ring_lidar`` is the generator fixed by SURVEY.md 8(d) config 2: 64 beams, elevation
linspace(-17.6 deg, +2.4 deg), sensor height 1.8 m, P/64 azimuth steps per beam with random phase,
ground hit at r = h / tan(-elev) unless a per-sector obstacle range U(5,75) m (180 sectors) is
closer, 0.2 % range noise, z noise sigma 0.02, intensity U(0,1).  The obstacle ranges are drawn per
(beam, sector): that is the reading of the survey's text which reproduces the statistics it quotes for
P = 200 000 (about 30.6 k pillars, 189 k kept points, 1130 windows / 1450 sets at 12x12 and 323 / 1016
at 24x24 shift 6; this generator: 30 781 / 189 692 / 1127 / 1464 / 325 / 1016 for seed 0).  With
``per_beam_obstacles=False`` every beam sees the same 180 ranges (vertical walls: all 64 beams of a sector
fall into the same few pillars) -- round 1's reading, about 16.6 k pillars and 897 / 596 sets; kept for the
sparse end of the density sweep.  ``uniform_disc`` is the density stress of config 5.
"""
import numpy as np


def ring_lidar(n_points: int, seed: int = 0, per_beam_obstacles: bool = True) -> np.ndarray:
    rng = np.random.default_rng(seed)
    beams = 64
    per = n_points // beams
    elev = np.deg2rad(np.linspace(-17.6, 2.4, beams))
    h = 1.8
    all_sectors = rng.uniform(5.0, 75.0, size=(beams if per_beam_obstacles else 1, 180))
    pts = np.empty((beams * per, 4), dtype=np.float32)
    for b in range(beams):
        az = (np.arange(per) + rng.uniform()) * (2 * np.pi / per)
        sec = np.minimum((az / (2 * np.pi) * 180).astype(np.int64), 179)
        obstacle = all_sectors[b if per_beam_obstacles else 0][sec]
        if elev[b] < 0:
            ground = h / np.tan(-elev[b])
            r = np.minimum(ground, obstacle)
        else:
            r = obstacle
        r = r * (1.0 + 0.002 * rng.standard_normal(per))
        x = r * np.cos(az)
        y = r * np.sin(az)
        z = r * np.tan(elev[b])      # sensor frame: the ground plane sits at about -1.8 m
        z = z + 0.02 * rng.standard_normal(per)
        sl = slice(b * per, (b + 1) * per)
        pts[sl, 0] = x
        pts[sl, 1] = y
        pts[sl, 2] = z
        pts[sl, 3] = rng.uniform(0.0, 1.0, per)
    if pts.shape[0] < n_points:   # pad the remainder with repeats of the first points
        pts = np.concatenate([pts, pts[: n_points - pts.shape[0]]], axis=0)
    return np.ascontiguousarray(pts[:n_points])


def uniform_disc(n_points: int, seed: int = 0) -> np.ndarray:
    rng = np.random.default_rng(seed)
    theta = rng.uniform(0, 2 * np.pi, n_points)
    r = 2.0 + 73.0 * rng.uniform(0, 1, n_points) ** 2
    z = np.clip(rng.normal(-1.0, 0.6, n_points), -4.9, 2.9)
    pts = np.stack([r * np.cos(theta), r * np.sin(theta), z, rng.uniform(0, 1, n_points)], axis=1)
    return np.ascontiguousarray(pts.astype(np.float32))


def head_candidates(k: int = 500, seed: int = 0, grid: int = 468):
    """Synthetic CenterHead top-K outputs: the eight inputs of filterBoxByScore."""
    rng = np.random.default_rng(seed)
    scores = np.sort(rng.uniform(0.0, 1.0, k).astype(np.float32))[::-1].copy()
    classes = rng.integers(0, 10, k).astype(np.int32)
    xs = rng.integers(0, grid, k).astype(np.int32)
    ys = rng.integers(0, grid, k).astype(np.int32)
    center = rng.uniform(-1.0, 2.0, (k, 2)).astype(np.float32)     # some land outside the range
    center_z = rng.uniform(-6.0, 4.0, k).astype(np.float32)
    angle = rng.uniform(-1.57, 1.57, k).astype(np.float32)
    dim = np.exp(rng.normal(0.5, 0.4, (k, 3))).astype(np.float32)
    return scores, classes, xs, ys, center, center_z, angle, dim
