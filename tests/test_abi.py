"""CPU-side checks of the C-ABI boundary: libmopa_scn.so loads without a GPU, exports every symbol include/mopa_scn.h
declares, and the ctypes prototypes in mopa_b200/_lib.py agree with the header's argument lists. No compute calls."""
import ctypes
import os
import re

import pytest

from mopa_b200 import _build, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADERS = [os.path.join(ROOT, "include", h) for h in ("mopa_scn.h", "mopa_xm.h")]


def _declarations():
    """name -> list of parameter strings, parsed from the header (comments stripped)."""
    src = "\n".join(open(h).read() for h in HEADERS)
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    src = re.sub(r"^\s*#.*$", " ", src, flags=re.M)
    decls = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(mopa_(?:scn|xm)_\w+)\s*\(([^;{}]*?)\)\s*;", src, flags=re.S):
        name, args = m.group(2), " ".join(m.group(3).split())
        params = [] if args in ("", "void") else [a.strip() for a in args.split(",")]
        decls[name] = (" ".join(m.group(1).split()), params)
    return decls


def _ctype_of(param):
    p = param.replace("const ", "").strip()
    if "*" in p:
        return "ptr"
    base = p.rsplit(" ", 1)[0].strip() if " " in p else p
    return {"int": "int", "int64_t": "i64", "float": "float", "double": "double", "size_t": "size", "uint64_t": "u64"}[base]


_CT = {ctypes.c_int: "int", ctypes.c_int64: "i64", ctypes.c_float: "float", ctypes.c_double: "double", ctypes.c_size_t: "size",
       ctypes.c_void_p: "ptr", ctypes.c_char_p: "ptr"}


def _kind(ct):
    if ct in _CT:
        return _CT[ct]
    if isinstance(ct, type) and issubclass(ct, ctypes._Pointer):
        return "ptr"
    raise AssertionError("unmapped ctypes type %r" % (ct,))


def test_header_declares_what_python_binds():
    decls = _declarations()
    assert set(decls) == set(_lib.PROTOTYPES), (set(decls) ^ set(_lib.PROTOTYPES))


@pytest.mark.parametrize("name", sorted(_lib.PROTOTYPES))
def test_prototype_matches_header(name):
    ret, params = _declarations()[name]
    _, argtypes = _lib.PROTOTYPES[name]
    assert len(params) == len(argtypes), (name, params)
    for p, ct in zip(params, argtypes):
        want, got = _ctype_of(p), _kind(ct)
        if want == "size" and got == "size":
            continue
        assert want == got, (name, p, ct)


def test_library_builds_loads_and_exports_every_symbol():
    path = _build.build()  # no-op when the in-tree .so is fresh
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)  # must load on a box without a GPU / driver (cudart is linked statically)
    for name in _declarations():
        assert hasattr(lib, name), "libmopa_scn.so does not export %s" % name
    lib.mopa_scn_abi_version.restype = ctypes.c_int
    assert lib.mopa_scn_abi_version() == _lib.ABI_VERSION


def test_metadata_new_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = _lib.load()
    assert not lib.mopa_scn_Metadata_new(3, 0)
    assert b"no CPU path" in lib.mopa_scn_last_error() or b"cuda" in lib.mopa_scn_last_error().lower()
    import mopa_b200.scn as scn
    with pytest.raises(_lib.ScnError):
        scn.InputLayer(3, 4096, mode=4)([torch.zeros(4, 4, dtype=torch.long), torch.ones(4, 1)])


def test_no_product_import_of_oracle():
    """The product package must never import the oracle (a CPU fallback would void parity claims)."""
    pkg = os.path.join(ROOT, "mopa_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f


@pytest.mark.parametrize("volume,subm", [(27, 1), (27, 0), (8, 0)])
def test_dweight_work_plan_covers_every_row_of_every_offset_once(volume, subm):
    """Host logic of the tcgen05 d_weight kernel (csrc/conv_dw_tc.cu::dw_tc_plan): items = (offset, row range). Replays the
    kernel's item -> (k, rows) mapping and checks that each offset's ranges tile [0, n_rows) exactly, that the centre offset
    of a submanifold table is split finer, and that the item count (= partial slices = workspace) stays bounded."""
    lib = _lib.load()
    for n_rows in (0, 1, 511, 512, 513, 3624, 23110, 52447, 106203, 230925, 818488, 2_000_003):
        plan = (ctypes.c_int * 6)()
        assert lib.mopa_scn_debug_dweightPlan(volume, subm, n_rows, plan) == 0
        centre, rpi, n_o, rpi_c, n_c, items = list(plan)
        assert (centre == 13) == bool(subm and volume == 27)
        covered = {k: [] for k in range(volume)}
        for item in range(items):  # same arithmetic as the kernel
            if centre >= 0:
                if item < n_c:
                    k, r0, r1 = centre, item * rpi_c, item * rpi_c + rpi_c
                else:
                    kk, j = divmod(item - n_c, n_o)
                    k, r0, r1 = (kk if kk < centre else kk + 1), j * rpi, j * rpi + rpi
            else:
                k, j = divmod(item, n_o)
                r0, r1 = j * rpi, j * rpi + rpi
            covered[k].append((r0, min(r1, n_rows)))
        for k, ranges in covered.items():
            ranges.sort()
            pos = 0
            for r0, r1 in ranges:
                assert r0 == pos or (r0 >= n_rows and r1 <= r0), (n_rows, k, ranges[:4])
                pos = max(pos, r1)
            assert pos == n_rows, (n_rows, k)
        assert items <= 4 * 148 + 40 * volume  # ~4 items per SM, plus rounding
        if centre >= 0 and n_rows > 4096:
            assert rpi_c < rpi  # centre offset: one rule per row, split finer (floor of 512 rows per item) ...
        if centre >= 0 and n_rows >= 100000:
            assert rpi_c * 4 <= rpi  # ... at least 4x finer on the large levels
