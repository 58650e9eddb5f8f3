"""ctypes front-end for oracle/_ref/libref_v<N>.so -- TEST INFRASTRUCTURE ONLY.

The shared objects are the UNMODIFIED reference translation units
(/root/reference/<N>/Fluid.cpp) compiled by oracle/Makefile through
oracle/ref_wrap.cpp.  They exist only where /root/reference was present at
build time (this container); they travel to the GPU box as prebuilt files.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
arm may import this module.
"""
import ctypes
import os
import shutil
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

# buffers the reference allocates with `new T[]` and never initialises
# (v3:94-95, v3:408-415, v5:358-369).  glibc hands back fresh zero pages for
# the >=128 KiB blocks the shipped 128^2 runs use; small test grids would see
# heap garbage instead, so the harness pins the "zero page" behaviour
# (SURVEY.md 3.5 quirk 4).
_UNINIT_SOLVER = ["r", "z", "s", "precon", "aDiag", "aPlusX", "aPlusY", "uDensity", "vDensity",
                  "qs.weight", "qs.counts"]
# (the particle arrays are NOT in this list: the constructor has already filled them (initParticles, v8:877);
# what lies past _particleCount is zero because ref_wrap.cpp zero-fills every allocation of the library)
_UNINIT_QUANTITY = ["dst", "old", "normalX", "normalY", "body", "mask", "phi"]


def available(version):
    """version: 1..8, or "8a8" for chapter 8 built with _AvgPerCell = 8 (BASELINE config 5)."""
    return os.path.exists(os.path.join(REF_DIR, "libref_v%s.so" % version))


def fnv64(arr):
    """FNV-1a-64 with one xor-multiply per 8-byte word."""
    words = np.ascontiguousarray(arr).view(np.uint64).ravel()
    h = 0xcbf29ce484222325
    prime = 0x100000001b3
    mask = (1 << 64) - 1
    for wd in words.tolist():
        h = ((h ^ wd) * prime) & mask
    return "%016x" % h


class Ref:
    """One reference FluidSolver instance of chapter `version` (1..8)."""

    def __init__(self, version, w, h, params, bodies=(), fresh_copy=None, zero_uninit=True, variant=""):
        path = os.path.join(REF_DIR, "libref_v%d%s.so" % (version, variant))
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (run `make -C oracle ref` where /root/reference exists)")
        # v8's frand() keeps a function-static LCG seed (v8:38-46): a fresh
        # dlopen of a private copy is the only way to reset it between solvers.
        if fresh_copy is None:
            fresh_copy = (version == 8)
        self._tmp = None
        if fresh_copy:
            fd, tmp = tempfile.mkstemp(suffix=".so", prefix="libref_v%d_" % version)
            os.close(fd)
            shutil.copyfile(path, tmp)
            self._tmp = tmp
            path = tmp
        lib = ctypes.CDLL(path)
        lib.ref_create.restype = ctypes.c_void_p
        lib.ref_create.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int,
                                   ctypes.c_void_p, ctypes.c_int]
        lib.ref_destroy.argtypes = [ctypes.c_void_p]
        lib.ref_buf.restype = ctypes.c_void_p
        lib.ref_buf.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_long),
                                ctypes.POINTER(ctypes.c_int)]
        lib.ref_call.restype = ctypes.c_int
        lib.ref_call.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        lib.ref_log.restype = ctypes.c_char_p
        self.lib = lib
        self.version = version
        self.w, self.h = w, h
        p = np.asarray(params, dtype=np.float64)
        b = np.asarray(bodies, dtype=np.float64).reshape(-1, 9) if len(bodies) else np.zeros((0, 9))
        self.nbodies = b.shape[0]
        self.ptr = lib.ref_create(w, h, p.ctypes.data, p.size, b.ctypes.data, b.shape[0])
        self._out = np.zeros(16)
        if zero_uninit:
            self._zero_uninit()

    def _zero_uninit(self):
        for n in _UNINIT_SOLVER:
            a = self.buf(n, optional=True)
            if a is not None:
                a[...] = 0
        if self.version >= 3:
            self.buf("p")[...] = 0
        for q in "dtuv":
            for f in _UNINIT_QUANTITY:
                a = self.buf(q + "." + f, optional=True)
                if a is not None:
                    a[...] = 0

    def close(self):
        if self.ptr:
            self.lib.ref_destroy(self.ptr)
            self.ptr = None
        if self._tmp:
            try:
                os.unlink(self._tmp)
            except OSError:
                pass
            self._tmp = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def buf(self, name, optional=False):
        """numpy VIEW (read/write) of one of the reference object's arrays."""
        cnt = ctypes.c_long()
        el = ctypes.c_int()
        p = self.lib.ref_buf(self.ptr, name.encode(), ctypes.byref(cnt), ctypes.byref(el))
        if not p:
            if optional:
                return None
            raise KeyError(name)
        dt = {8: np.float64, 4: np.int32, 1: np.uint8}[el.value]
        cbuf = (ctypes.c_char * (cnt.value * el.value)).from_address(p)
        return np.frombuffer(cbuf, dtype=dt)

    def call(self, op, *args):
        a = np.asarray(args if args else [0.0], dtype=np.float64)
        rc = self.lib.ref_call(self.ptr, op.encode(), a.ctypes.data, a.size, self._out.ctypes.data)
        if rc != 0:
            raise KeyError("reference v%d has no op %r" % (self.version, op))
        return float(self._out[0])

    def body_state(self, i):
        a = np.asarray([float(i)])
        self.lib.ref_call(self.ptr, b"bodyState", a.ctypes.data, 1, self._out.ctypes.data)
        return self._out[:8].copy()

    def log(self, clear=True):
        s = self.lib.ref_log().decode()
        if clear:
            self.lib.ref_log_clear()
        return s
