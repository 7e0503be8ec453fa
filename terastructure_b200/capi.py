"""ctypes binding of libtsgpu.so (include/tsgpu.h).

This is plumbing only: every call goes straight through the C ABI that a maintainer of the
reference would bind.  There is no Python or CPU fallback; when the shared
library is missing or no B200 is visible the calls raise.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TSGPU_LIB") or os.path.join(_HERE, "lib", "libtsgpu.so")  # TSGPU_LIB: A/B builds

TS_MAX_K = 32
TS_COMM_HANDLE_BYTES = 64


class TsError(RuntimeError):
    pass


class TsConfig(C.Structure):
    _fields_ = [
        ("n_total", C.c_uint64), ("n_begin", C.c_uint64), ("n_local", C.c_uint64), ("l", C.c_uint64),
        ("k", C.c_uint32), ("online_iterations", C.c_uint32),
        ("alpha", C.c_double), ("eta0", C.c_double), ("eta1", C.c_double),
        ("nodetau0", C.c_double), ("nodekappa", C.c_double), ("meanchangethresh", C.c_double),
        ("device", C.c_int32), ("rank", C.c_int32), ("nranks", C.c_int32), ("reserved", C.c_int32),
    ]


# name -> (restype, argtypes); the list mirrors include/tsgpu.h one to one
_vp, _u32, _u64, _i, _d = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int, C.c_double
EXPORTS = {
    "ts_config_defaults": (None, [C.POINTER(TsConfig), _u64, _u64, _u32]),
    "ts_last_error": (C.c_char_p, []),
    "ts_abi_version": (_i, []),
    "ts_device_count": (_i, []),
    "ts_create": (_i, [C.POINTER(TsConfig), C.POINTER(_vp)]),
    "ts_destroy": (_i, [_vp]),
    "ts_load_bed": (_i, [_vp, _u64, _u64, _vp, _u64]),
    "ts_load_bed_fanout": (_i, [C.POINTER(_vp), _i, _u64, _u64, _vp, _u64]),
    "ts_synth_bed": (_i, [_vp, _u64, _vp, _vp, _d]),
    "ts_get_bed_row": (_i, [_vp, _u64, _vp]),
    "ts_set_validation": (_i, [_vp, _u64, _vp, _vp, _vp]),
    "ts_set_gamma": (_i, [_vp, _vp]),
    "ts_reset_lambda": (_i, [_vp]),
    "ts_reset_counts": (_i, [_vp]),
    "ts_step": (_i, [_vp, _u32, _i, C.POINTER(_i)]),
    "ts_steps": (_i, [_vp, _vp, _u64, _i, _vp]),
    "ts_heldout_ll": (_i, [_vp, _i, C.POINTER(_d), C.POINTER(_u64), _vp]),
    "ts_get_gamma": (_i, [_vp, _vp]),
    "ts_get_theta": (_i, [_vp, _vp]),
    "ts_get_elogtheta": (_i, [_vp, _vp]),
    "ts_get_counts": (_i, [_vp, _vp]),
    "ts_get_lambda": (_i, [_vp, _u64, _u64, _vp]),
    "ts_get_beta": (_i, [_vp, _u64, _u64, _vp]),
    "ts_sync": (_i, [_vp]),
    "ts_comm_export": (_i, [_vp, _vp]),
    "ts_comm_connect": (_i, [_vp, _vp]),
    "ts_comm_connect_local": (_i, [C.POINTER(_vp), _i]),
    "ts_comm_state_bytes": (_u64, []),
    "ts_comm_mode": (_i, [_vp]),
    "ts_comm_attach_symmetric": (_i, [_vp, C.POINTER(_vp), _vp, _u64, _u32]),
    "ts_plan_shard": (_i, [_u64, _i, _i, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "ts_get_plan": (_i, [_vp, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "ts_plan_tiers": (_i, [_u64, _i, _i, C.POINTER(_i), C.POINTER(_u64)]),
    "ts_get_tiers": (_i, [_vp, C.POINTER(_i), C.POINTER(_u64)]),
    "ts_launch_count": (_u64, [_vp]),
    "ts_timer_start": (_i, [_vp]),
    "ts_timer_stop": (_i, [_vp, C.POINTER(C.c_float)]),
    "ts_debug_trace": (_i, [_vp, _vp]),
    "ts_rng_create": (_vp, [_d]),
    "ts_rng_destroy": (None, [_vp]),
    "ts_rng_get": (_u32, [_vp]),
    "ts_rng_uniform_int": (_u32, [_vp, _u32]),
    "ts_rng_gamma": (_d, [_vp, _d, _d]),
    "ts_rng_sample_locs": (None, [_vp, _u32, _vp, _u64]),
    "ts_sample_validation": (_i, [_vp, _u64, _u64, _vp, _u64, C.POINTER(_u64), C.POINTER(_vp),
                                  C.POINTER(_vp), C.POINTER(_vp)]),
    "ts_init_gamma": (None, [_vp, _u64, _u32, _vp]),
    "ts_free": (None, [_vp]),
}

_lib = None


def lib():
    """Load libtsgpu.so (built in-tree by __graft_entry__.build()).  Raises if absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            # a fresh checkout: build in-tree once (nvcc cross-compiles sm_100a without a GPU)
            import subprocess
            try:
                subprocess.run(["make", "-s", "-j", str(os.cpu_count() or 4), "-C", os.path.join(_HERE, "csrc")],
                               check=True, stdout=subprocess.DEVNULL)
            except (OSError, subprocess.CalledProcessError) as ex:
                raise TsError(f"{LIB_PATH} is not built and building it failed ({ex}); run "
                              "`make -C terastructure_b200/csrc` (there is no fallback path)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        raise TsError(f"tsgpu error {rc}: {lib().ts_last_error().decode()}")


def _ptr(a):
    return a.ctypes.data if a is not None else None


class Rng:
    """The reference's host RNG stream (gsl_rng_mt19937, snpsamplinge.cc:59-63)."""

    def __init__(self, seed):
        self._h = lib().ts_rng_create(float(seed))

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.ts_rng_destroy(self._h)
            self._h = None

    def get(self):
        return lib().ts_rng_get(self._h)

    def uniform_int(self, n):
        return lib().ts_rng_uniform_int(self._h, n)

    def gamma(self, a, b):
        return lib().ts_rng_gamma(self._h, a, b)

    def sample_locs(self, l, n):
        out = np.empty(n, dtype=np.uint32)
        lib().ts_rng_sample_locs(self._h, l, out.ctypes.data, n)
        return out

    def sample_validation(self, n, l, bed_rows):
        """set_validation_sample (snpsamplinge.cc:196-224) -> CSR (val_loc, val_off, val_indiv)."""
        if bed_rows is None:  # no missing genotypes anywhere (device-generated synthetic data)
            bed_ptr, pitch = None, 0
        else:
            bed_rows = np.ascontiguousarray(bed_rows, dtype=np.uint8)
            assert bed_rows.ndim == 2 and bed_rows.shape[0] == l and bed_rows.shape[1] >= (n + 3) // 4
            bed_ptr, pitch = bed_rows.ctypes.data, bed_rows.shape[1]
        nval = _u64()
        pl, po, pi = _vp(), _vp(), _vp()
        check(lib().ts_sample_validation(self._h, n, l, bed_ptr, pitch,
                                         C.byref(nval), C.byref(pl), C.byref(po), C.byref(pi)))
        nv = nval.value
        loc = np.ctypeslib.as_array(C.cast(pl, C.POINTER(_u32)), (max(nv, 1),))[:nv].copy()
        off = np.ctypeslib.as_array(C.cast(po, C.POINTER(_u64)), (nv + 1,)).copy()
        ind = np.ctypeslib.as_array(C.cast(pi, C.POINTER(_u32)), (max(int(off[-1]), 1),))[:int(off[-1])].copy()
        for p in (pl, po, pi):
            lib().ts_free(p)
        return loc, off, ind

    def init_gamma(self, n, k):
        out = np.empty((n, k), dtype=np.float64)
        lib().ts_init_gamma(self._h, n, k, out.ctypes.data)
        return out


class Engine:
    """One shard of individuals resident on one B200 (ts_engine)."""

    def __init__(self, n, l, k, *, device=0, rank=0, nranks=1, n_begin=0, n_local=None,
                 online_iterations=10, eta=None):
        cfg = TsConfig()
        lib().ts_config_defaults(C.byref(cfg), n, l, k)
        if eta is not None:  # prior of lambda (the reference hard-codes 1.0, snpsamplinge.cc:25-27)
            cfg.eta0 = cfg.eta1 = float(eta)
        cfg.device, cfg.rank, cfg.nranks = device, rank, nranks
        cfg.n_begin = n_begin
        cfg.n_local = n if n_local is None else n_local
        cfg.online_iterations = online_iterations
        self.cfg = cfg
        self.n, self.l, self.k = n, l, k
        self.n_local = int(cfg.n_local)
        self._h = _vp()
        check(lib().ts_create(C.byref(cfg), C.byref(self._h)))
        self.nval = 0

    def close(self):
        if getattr(self, "_h", None) and self._h.value and _lib is not None:
            _lib.ts_destroy(self._h)
            self._h = _vp()

    __del__ = close

    # ---- data -----------------------------------------------------------------------------
    def load_bed(self, rows, loc_begin=0):
        rows = np.ascontiguousarray(rows, dtype=np.uint8)
        assert rows.ndim == 2
        check(lib().ts_load_bed(self._h, loc_begin, rows.shape[0], rows.ctypes.data, rows.shape[1]))

    def synth_bed(self, seed, theta, beta, missing_rate=0.0):
        theta = np.ascontiguousarray(theta, dtype=np.float32)
        beta = np.ascontiguousarray(beta, dtype=np.float32)
        assert theta.shape == (self.n_local, self.k) and beta.shape == (self.l, self.k)
        check(lib().ts_synth_bed(self._h, seed, theta.ctypes.data, beta.ctypes.data, missing_rate))

    def get_bed_row(self, loc):
        out = np.empty((self.n_local + 3) // 4, dtype=np.uint8)
        check(lib().ts_get_bed_row(self._h, loc, out.ctypes.data))
        return out

    def set_validation(self, val_loc, val_off, val_indiv):
        val_loc = np.ascontiguousarray(val_loc, dtype=np.uint32)
        val_off = np.ascontiguousarray(val_off, dtype=np.uint64)
        val_indiv = np.ascontiguousarray(val_indiv, dtype=np.uint32)
        check(lib().ts_set_validation(self._h, val_loc.size, _ptr(val_loc), _ptr(val_off), _ptr(val_indiv)))
        self.nval = int(val_loc.size)

    def set_gamma(self, gamma_rows):
        g = np.ascontiguousarray(gamma_rows, dtype=np.float64)
        assert g.shape == (self.n_local, self.k)
        check(lib().ts_set_gamma(self._h, g.ctypes.data))

    def reset_lambda(self):
        check(lib().ts_reset_lambda(self._h))

    def reset_counts(self):
        check(lib().ts_reset_counts(self._h))

    # ---- hot path -------------------------------------------------------------------------
    def step(self, loc, hol_mode=False):
        r = _i()
        check(lib().ts_step(self._h, int(loc), int(hol_mode), C.byref(r)))
        return r.value

    def steps(self, locs, hol_mode=False, want_rounds=False):
        locs = np.ascontiguousarray(locs, dtype=np.uint32)
        rounds = np.empty(locs.size, dtype=np.uint32) if want_rounds else None
        check(lib().ts_steps(self._h, locs.ctypes.data, locs.size, int(hol_mode), _ptr(rounds)))
        return rounds

    def heldout_ll(self, first=False):
        """-> (sum, count, per_locus_sums) for this shard."""
        s, c = _d(), _u64()
        per = np.zeros(max(self.nval, 1), dtype=np.float64)
        check(lib().ts_heldout_ll(self._h, int(first), C.byref(s), C.byref(c), per.ctypes.data))
        return s.value, c.value, per[:self.nval]

    # ---- read-back ------------------------------------------------------------------------
    def _rows(self, fn):
        out = np.empty((self.n_local, self.k), dtype=np.float64)
        check(getattr(lib(), fn)(self._h, out.ctypes.data))
        return out

    gamma = property(lambda s: s._rows("ts_get_gamma"))
    theta = property(lambda s: s._rows("ts_get_theta"))
    elogtheta = property(lambda s: s._rows("ts_get_elogtheta"))

    @property
    def counts(self):
        out = np.empty(self.n_local, dtype=np.uint32)
        check(lib().ts_get_counts(self._h, out.ctypes.data))
        return out

    def get_lambda(self, loc_begin=0, nloc=None):
        nloc = self.l - loc_begin if nloc is None else nloc
        out = np.empty((nloc, self.k, 2), dtype=np.float64)
        check(lib().ts_get_lambda(self._h, loc_begin, nloc, out.ctypes.data))
        return out

    def get_beta(self, loc_begin=0, nloc=None):
        nloc = self.l - loc_begin if nloc is None else nloc
        out = np.empty((nloc, self.k), dtype=np.float64)
        check(lib().ts_get_beta(self._h, loc_begin, nloc, out.ctypes.data))
        return out

    def sync(self):
        check(lib().ts_sync(self._h))

    # ---- exchange / profiling ---------------------------------------------------------------
    def attach_symmetric(self, rank_ptrs, multicast_ptr, nbytes, total_ctas=0):
        """Symmetric exchange buffers (one per rank, mapped on this device) and an optional NVLS
        multicast alias; barrier between this call and the first steps()."""
        arr = (_vp * len(rank_ptrs))(*[int(p) for p in rank_ptrs])
        check(lib().ts_comm_attach_symmetric(self._h, arr, _vp(int(multicast_ptr) or None), nbytes, total_ctas))

    @property
    def comm_mode(self):
        return lib().ts_comm_mode(self._h)

    def comm_export(self):
        buf = C.create_string_buffer(TS_COMM_HANDLE_BYTES)
        check(lib().ts_comm_export(self._h, buf))
        return buf.raw

    def comm_connect(self, handles):
        blob = b"".join(handles)
        assert len(blob) == TS_COMM_HANDLE_BYTES * self.cfg.nranks
        check(lib().ts_comm_connect(self._h, blob))

    @property
    def plan(self):
        """(individuals per thread, CTAs, threads per CTA) of this engine's persistent kernel."""
        ipt, grid, block = C.c_int(), C.c_int(), C.c_int()
        check(lib().ts_get_plan(self._h, C.byref(ipt), C.byref(grid), C.byref(block)))
        return ipt.value, grid.value, block.value

    @property
    def tiers(self):
        """(individuals per thread in shared memory, streamed individuals); (-1, 0) = register-only kernel."""
        j, ns = C.c_int(), _u64()
        check(lib().ts_get_tiers(self._h, C.byref(j), C.byref(ns)))
        return j.value, ns.value

    @property
    def launch_count(self):
        return lib().ts_launch_count(self._h)

    def debug_trace(self):
        out = np.zeros((64, 128), dtype=np.int64)
        check(lib().ts_debug_trace(self._h, out.ctypes.data))
        return out

    def timer_start(self):
        check(lib().ts_timer_start(self._h))

    def timer_stop(self):
        ms = C.c_float()
        check(lib().ts_timer_stop(self._h, C.byref(ms)))
        return ms.value


def load_bed_fanout(engines, rows, loc_begin=0):
    """One pass over `rows` ([nloc, >= ceil(N/4)] packed, e.g. a np.memmap of a .bed) for all engines."""
    assert rows.ndim == 2 and rows.dtype == np.uint8 and rows.strides[1] == 1
    arr = (_vp * len(engines))(*[e._h for e in engines])
    check(lib().ts_load_bed_fanout(arr, len(engines), loc_begin, rows.shape[0], rows.ctypes.data, rows.strides[0]))


def connect_local(engines):
    arr = (_vp * len(engines))(*[e._h for e in engines])
    check(lib().ts_comm_connect_local(arr, len(engines)))


def plan_tiers(n_local, k, num_sms=148):
    """(individuals per thread in shared memory, individuals streamed from L2/HBM every round) the engine
    chooses for a shard of `n_local` individuals; (0, 0) for a register-resident shard."""
    j, ns = C.c_int(), _u64()
    check(lib().ts_plan_tiers(n_local, k, num_sms, C.byref(j), C.byref(ns)))
    return j.value, ns.value


def plan_shard(n_local, k, num_sms=148):
    """(individuals per thread, CTAs, threads per CTA) the engine uses for a shard of `n_local`
    individuals (register tier; plan_tiers gives the shared-memory and streaming tiers).  Host arithmetic only."""
    ipt, grid, block = C.c_int(), C.c_int(), C.c_int()
    check(lib().ts_plan_shard(n_local, k, num_sms, C.byref(ipt), C.byref(grid), C.byref(block)))
    return ipt.value, grid.value, block.value
