"""Shared input builders for the parity tests."""
import numpy as np
import torch

from mopa_b200 import synth


def small_batch(n_scans=2, n_azimuth=120, seed=0):
    """A small synthetic batch (a few thousand points) in the collate_scn_base layout."""
    return synth.make_batch(n_scans, "nuscenes", seed, n_azimuth=n_azimuth)


def random_cloud(n, extent, seed, n_batch=2, dup_frac=0.3, lo=0):
    """Dense-ish random cloud with duplicates, to get many neighbours per site."""
    rng = np.random.default_rng(seed)
    c = rng.integers(lo, lo + extent, size=(n, 3))
    b = rng.integers(0, n_batch, size=(n, 1))
    coords = np.concatenate([c, b], 1).astype(np.int64)
    ndup = int(n * dup_frac)
    if ndup:
        coords[rng.integers(0, n, ndup)] = coords[rng.integers(0, n, ndup)]
    return coords


def rel_err(a, b):
    a = torch.as_tensor(a, dtype=torch.float64).cpu()
    b = torch.as_tensor(b, dtype=torch.float64).cpu()
    denom = b.abs().max().clamp_min(1e-30)
    return float((a - b).abs().max() / denom)
