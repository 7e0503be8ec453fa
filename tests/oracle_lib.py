"""ctypes binding of the parity oracle (oracle/ts_oracle.c).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = os.path.join(ROOT, "oracle", "_build", "libts_oracle.so")


def build():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port"], check=True)


def _load():
    src = os.path.join(ROOT, "oracle", "ts_oracle.c")
    if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        build()
    lib = C.CDLL(_LIB)
    u32, f64, vp = C.c_uint32, C.c_double, C.c_void_p
    P = C.POINTER
    lib.tso_create.restype = vp
    lib.tso_create.argtypes = [u32, u32, u32, vp, f64, u32, C.c_int]
    lib.tso_destroy.argtypes = [vp]
    for name in ("tso_nval_loc", "tso_per_loc_h", "tso_iter", "tso_last_rounds", "tso_sample_loc"):
        getattr(lib, name).restype = u32
        getattr(lib, name).argtypes = [vp]
    lib.tso_validation.argtypes = [vp, vp, vp]
    for name in ("tso_get_gamma", "tso_get_theta", "tso_get_elogtheta", "tso_get_lambda",
                 "tso_get_beta", "tso_get_counts", "tso_set_gamma"):
        getattr(lib, name).argtypes = [vp, vp]
    lib.tso_train_loc.restype = u32
    lib.tso_train_loc.argtypes = [vp, u32]
    lib.tso_flush.argtypes = [vp]
    lib.tso_heldout.restype = C.c_int
    lib.tso_heldout.argtypes = [vp, C.c_int, P(f64), P(u32), vp]
    lib.tso_infer.restype = u32
    lib.tso_infer.argtypes = [vp, u32, u32, u32, vp, vp, vp, u32, vp, P(u32), P(C.c_int)]
    lib.tso_compute_all_lambda.argtypes = [vp]
    lib.tso_decode_bed.argtypes = [vp, u32, u32, vp]
    return lib


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = _load()
    return _lib


def decode_bed(bed_bytes, n, l):
    """bed_bytes: the .bed payload WITHOUT the 3-byte header. Returns y[l, n] uint8."""
    bed = np.ascontiguousarray(np.frombuffer(bed_bytes, dtype=np.uint8))
    assert bed.size >= l * ((n + 3) // 4)
    y = np.empty((l, n), dtype=np.uint8)
    lib().tso_decode_bed(bed.ctypes.data, n, l, y.ctypes.data)
    return y


class Oracle:
    """Sequential restatement of SNPSamplingE at -nthreads 1."""

    def __init__(self, y, k, seed, online_iterations=10, compute_beta=False):
        self.y = np.ascontiguousarray(y, dtype=np.uint8)
        self.l, self.n = self.y.shape
        self.k = k
        self._h = lib().tso_create(self.n, self.l, k, self.y.ctypes.data, float(seed),
                                   online_iterations, int(compute_beta))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().tso_destroy(self._h)
            self._h = None

    def _get(self, fn, shape, dtype=np.float64):
        out = np.empty(shape, dtype=dtype)
        getattr(lib(), fn)(self._h, out.ctypes.data)
        return out

    gamma = property(lambda s: s._get("tso_get_gamma", (s.n, s.k)))
    theta = property(lambda s: s._get("tso_get_theta", (s.n, s.k)))
    elogtheta = property(lambda s: s._get("tso_get_elogtheta", (s.n, s.k)))
    lam = property(lambda s: s._get("tso_get_lambda", (s.l, s.k, 2)))
    beta = property(lambda s: s._get("tso_get_beta", (s.l, s.k)))
    counts = property(lambda s: s._get("tso_get_counts", (s.n,), np.uint32))
    iter = property(lambda s: lib().tso_iter(s._h))

    def set_gamma(self, g):
        g = np.ascontiguousarray(g, dtype=np.float64)
        assert g.shape == (self.n, self.k)
        lib().tso_set_gamma(self._h, g.ctypes.data)

    def validation(self):
        """(val_loc[nv] ascending, val_off[nv+1], val_indiv[...]) -- CSR."""
        nv, h = lib().tso_nval_loc(self._h), lib().tso_per_loc_h(self._h)
        loc = np.empty(nv, dtype=np.uint32)
        ind = np.empty(nv * h, dtype=np.uint32)
        lib().tso_validation(self._h, loc.ctypes.data, ind.ctypes.data)
        off = (np.arange(nv + 1, dtype=np.uint64) * h)
        return loc, off, ind

    def sample_loc(self):
        return lib().tso_sample_loc(self._h)

    def train_loc(self, loc):
        return lib().tso_train_loc(self._h, int(loc))

    def flush(self):
        lib().tso_flush(self._h)

    def heldout(self, first=False, per_locus=False):
        a, c = C.c_double(), C.c_uint32()
        pl = np.zeros(lib().tso_nval_loc(self._h)) if per_locus else None
        stop = lib().tso_heldout(self._h, int(first), C.byref(a), C.byref(c),
                                 pl.ctypes.data if per_locus else None)
        return (bool(stop), a.value, c.value, pl) if per_locus else (bool(stop), a.value, c.value)

    def infer(self, rfreq, max_iter, cap=4096, lcap=0):
        ri = np.zeros(cap, np.uint32); rl = np.zeros(cap); rc = np.zeros(cap, np.uint32)
        locs = np.zeros(max(lcap, 1), np.uint32)
        nl, st = C.c_uint32(), C.c_int()
        nrep = lib().tso_infer(self._h, rfreq, max_iter, cap, ri.ctypes.data, rl.ctypes.data,
                               rc.ctypes.data, lcap, locs.ctypes.data if lcap else None,
                               C.byref(nl), C.byref(st))
        nrep = min(nrep, cap)
        return dict(iters=ri[:nrep], ll=rl[:nrep], count=rc[:nrep], locs=locs[:min(nl.value, lcap)],
                    nlocs=nl.value, stopped=bool(st.value))

    def compute_all_lambda(self):
        lib().tso_compute_all_lambda(self._h)
