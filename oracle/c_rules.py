"""ctypes binding of oracle/scn_rules.c (CPU ORACLE, test infrastructure; PARITY UNPINNED -- see scn_oracle.py).

`CGeometry` is a drop-in for `scn_oracle.Geometry` whose grids/rulebooks come from the C hash-map restatement;
bench.py's CPU baseline uses it so the timed rulebook leg is the serial hash-map algorithm upstream runs.
"""
import ctypes
import os
import subprocess

import numpy as np

from . import scn_oracle as so

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle_rules.so")
_lib = None


def build(force=False):
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "scn_rules.c")):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        p = ctypes.c_void_p
        L.oracle_input_rules.restype = ctypes.c_int64
        L.oracle_input_rules.argtypes = [p, ctypes.c_int64, ctypes.c_int, p, p]
        L.oracle_subm_rules.restype = ctypes.c_int
        L.oracle_subm_rules.argtypes = [p, ctypes.c_int64, ctypes.c_int64, p]
        L.oracle_strided_rules.restype = ctypes.c_int64
        L.oracle_strided_rules.argtypes = [p, ctypes.c_int64, p, p, p]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def input_rules(coords):
    coords = np.ascontiguousarray(coords, np.int64)
    n, ncols = coords.shape
    p2v = np.empty(n, np.int32)
    vc = np.empty((max(n, 1), 4), np.int64)
    v = lib().oracle_input_rules(_ptr(coords), n, ncols, _ptr(p2v), _ptr(vc))
    assert v >= 0
    return vc[:v].copy(), p2v


def subm_rules(voxel_coords, spatial_size):
    vc = np.ascontiguousarray(voxel_coords, np.int64)
    v = vc.shape[0]
    nbr = np.empty((27, v), np.int32)
    assert lib().oracle_subm_rules(_ptr(vc), v, int(spatial_size), _ptr(nbr)) == 0
    return nbr


def strided_rules(fine_coords):
    fc = np.ascontiguousarray(fine_coords, np.int64)
    vf = fc.shape[0]
    cc = np.empty((max(vf, 1), 4), np.int64)
    parent = np.empty(vf, np.int32)
    kidx = np.empty(vf, np.int32)
    vc = lib().oracle_strided_rules(_ptr(fc), vf, _ptr(cc), _ptr(parent), _ptr(kidx))
    assert vc >= 0
    return cc[:vc].copy(), parent, kidx


class CGeometry(so.Geometry):
    """scn_oracle.Geometry with the integer work done by the C hash-map restatement."""

    def __init__(self, coords, spatial_size=4096):
        self.spatial_size = int(spatial_size)
        coords = so.with_batch_column(coords)
        self.n_points = coords.shape[0]
        vc, self.p2v = input_rules(coords)
        self.csr_rows = np.argsort(self.p2v, kind="stable").astype(np.int32)
        self.csr_off = np.concatenate([[0], np.cumsum(np.bincount(self.p2v, minlength=vc.shape[0]))]).astype(np.int32)
        self.level_coords = [vc]
        self.subm = {}
        self.down = {}

    def subm_table(self, level):
        if level not in self.subm:
            self.subm[level] = subm_rules(self.level_coords[level], self.spatial_size >> level)
        return self.subm[level]

    def down_rules(self, level):
        if level not in self.down:
            cc, parent, kidx = strided_rules(self.level_coords[level])
            self.level_coords.append(cc)
            self.down[level] = (parent, kidx)
        return self.down[level]
