"""Synthetic lidar scans shaped like the inputs MoPA feeds to UNetSCN.

No dataset is available offline, so the bench and the tests use seeded synthetic scans that follow
SURVEY.md section 8(d): a spinning lidar over a ground plane with random wall sectors, then the
reference's own point -> voxel-coordinate recipe (`mopa/data/utils/augmentation_3d.py:54-59`,
`mopa/data/nuscenes/nuscenes_dataloader.py:415-427`) and the reference's batch layout
(`mopa/data/collate.py:182-186,233-235`): `coords` int64 (N, 4) with the batch index LAST, on the
host; `feats` float32 (N, 1) of ones.
"""
import numpy as np

# name -> (beams, azimuth steps, elevation range in degrees, sensor height m, max range m)
SENSORS = {
    "nuscenes": (32, 1085, (-30.0, 10.0), 1.84, 70.0),
    "kitti": (64, 2000, (-24.8, 2.0), 1.73, 80.0),
}


def lidar_points(sensor="nuscenes", seed=0, n_azimuth=None, n_sectors=64):
    """One sweep of a synthetic spinning lidar; returns float64 (N, 3) points in metres, firing order."""
    beams, n_az, (e_lo, e_hi), height, max_range = SENSORS[sensor]
    if n_azimuth is not None:
        n_az = int(n_azimuth)
    rng = np.random.default_rng(seed)
    elev = np.deg2rad(np.linspace(e_lo, e_hi, beams))
    az = np.linspace(0.0, 2.0 * np.pi, n_az, endpoint=False)
    az_g, el_g = np.meshgrid(az, elev, indexing="ij")  # azimuth-major = firing order
    az_g, el_g = az_g.ravel(), el_g.ravel()

    # ground plane z = -height
    with np.errstate(divide="ignore", invalid="ignore"):
        t_ground = np.where(el_g < 0, height / np.tan(-el_g), np.inf)  # horizontal distance of the hit

    # wall sectors: an arc at horizontal distance r, from the ground up to h metres
    sec_r = rng.uniform(5.0, 60.0, n_sectors)
    sec_h = rng.uniform(1.5, 12.0, n_sectors)
    sec_a0 = rng.uniform(0.0, 2.0 * np.pi, n_sectors)
    sec_w = rng.uniform(0.05, 0.35, n_sectors)
    t_wall = np.full(az_g.shape, np.inf)
    for r, h, a0, w in zip(sec_r, sec_h, sec_a0, sec_w):
        inside = ((az_g - a0) % (2.0 * np.pi)) < w
        z_hit = r * np.tan(el_g)
        ok = inside & (z_hit >= -height) & (z_hit <= h - height)
        t_wall = np.where(ok & (r < t_wall), r, t_wall)

    t = np.minimum(t_ground, t_wall)
    rng_3d = t / np.cos(el_g)
    keep = np.isfinite(t) & (rng_3d < max_range)
    t, az_k, el_k = t[keep], az_g[keep], el_g[keep]
    pts = np.stack([t * np.cos(az_k), t * np.sin(az_k), t * np.tan(el_k)], 1)
    pts += rng.normal(0.0, 0.02, pts.shape)
    return pts


def voxel_coords(points, scale=20, full_scale=4096, rng=None, transl=True):
    """points (metres) -> int64 voxel coords, restating augmentation_3d.py:54-59 + nuscenes_dataloader.py:419-424."""
    coords = np.round(points * scale)
    coords -= coords.min(0)
    if transl:
        rng = rng or np.random.default_rng(0)
        offset = np.clip(full_scale - coords.max(0) - 0.001, a_min=0, a_max=None) * rng.random(3)
        coords += offset
    coords = coords.astype(np.int64)
    keep = (coords.min(1) >= 0) & (coords.max(1) < full_scale)
    return coords[keep]


def make_scan(sensor="nuscenes", seed=0, n_azimuth=None, scale=20, full_scale=4096):
    """One scan as the dataloader emits it: coords int64 (N, 3), feats float32 (N, 1) = 1."""
    pts = lidar_points(sensor, seed, n_azimuth)
    coords = voxel_coords(pts, scale, full_scale, np.random.default_rng(seed + 7919))
    feats = np.ones((coords.shape[0], 1), np.float32)
    return coords, feats


def make_batch(batch_size=8, sensor="nuscenes", seed=0, n_azimuth=None, scale=20, full_scale=4096):
    """A collated batch, layout of collate_scn_base: coords (N, 4) int64 with batch index last, feats (N, 1)."""
    locs, feats = [], []
    for b in range(batch_size):
        c, f = make_scan(sensor, seed * 1000 + b, n_azimuth, scale, full_scale)
        locs.append(np.concatenate([c, np.full((c.shape[0], 1), b, np.int64)], 1))
        feats.append(f)
    return np.concatenate(locs, 0), np.concatenate(feats, 0)


def azimuth_for_points(target_points, sensor="nuscenes"):
    """Azimuth resolution giving roughly `target_points` returns per scan (SURVEY 8(d) config 5 sweep)."""
    beams, n_az, _, _, _ = SENSORS[sensor]
    base = lidar_points(sensor, 0).shape[0]
    return max(16, int(round(n_az * target_points / base)))
