"""ctypes front-end for oracle/libifl_oracle.so (the C restatement) -- TEST INFRASTRUCTURE ONLY.

`PortSolver` mirrors the reference's FluidSolver (chapters 1-3) on numpy arrays so
that parity tests and bench.py's cpu_baseline can drive the same call sequence on
both sides.  Nothing in the product imports this module.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libifl_oracle.so")


class Grid(ctypes.Structure):
    _fields_ = [("w", ctypes.c_int), ("h", ctypes.c_int), ("ox", ctypes.c_double), ("oy", ctypes.c_double)]


def build(force=False):
    src = os.path.join(HERE, "ifl_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "port"], stdout=subprocess.DEVNULL)
    return LIB


_lib = None
_D = ctypes.c_double
_I = ctypes.c_int
_P = ctypes.c_void_p


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(LIB)
        L.ofl_lerp.restype = _D
        L.ofl_lerp.argtypes = [_P, Grid, _D, _D]
        L.ofl_cerp.restype = _D
        L.ofl_cerp.argtypes = [_P, Grid, _D, _D]
        L.ofl_advect.argtypes = [_I, _P, _P, Grid, _P, Grid, _P, Grid, _D, _D]
        L.ofl_add_inflow.argtypes = [_P, Grid, _D, _D, _D, _D, _D, _D, _I]
        L.ofl_build_rhs.argtypes = [_P, _P, _P, _I, _I, _D]
        L.ofl_build_pressure_matrix.argtypes = [_P, _P, _P, _I, _I, _D, _D, _D]
        L.ofl_build_preconditioner.argtypes = [_P, _P, _P, _P, _I, _I]
        L.ofl_apply_preconditioner.argtypes = [_P, _P, _P, _P, _P, _I, _I]
        L.ofl_dot_product.restype = _D
        L.ofl_dot_product.argtypes = [_P, _P, _I]
        L.ofl_matrix_vector_product.argtypes = [_P, _P, _P, _P, _P, _I, _I]
        L.ofl_scaled_add.argtypes = [_P, _P, _P, _D, _I]
        L.ofl_infinity_norm.restype = _D
        L.ofl_infinity_norm.argtypes = [_P, _I]
        L.ofl_project.restype = _I
        L.ofl_project.argtypes = [_I, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P, _P]
        L.ofl_project_gs.restype = _I
        L.ofl_project_gs.argtypes = [_I, _D, _D, _D, _P, _P, _I, _I, _P, _P]
        L.ofl_apply_pressure.argtypes = [_P, _P, _P, _I, _I, _D, _D, _D]
        _lib = L
    return _lib


def _p(a):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data


class PortSolver:
    """FluidSolver of chapter `version` in {1, 2, 3} (v3:187-466, v2:208-350, v1:157-340)."""

    def __init__(self, version, w, h, density):
        assert version in (1, 2, 3)
        self.L = lib()
        self.version, self.w, self.h, self.density = version, w, h, density
        self.hx = 1.0 / min(w, h)
        self.g = {"d": Grid(w, h, 0.5, 0.5), "u": Grid(w + 1, h, 0.0, 0.5), "v": Grid(w, h + 1, 0.5, 0.0)}
        self.src = {k: np.zeros(g.w * g.h) for k, g in self.g.items()}
        self.dst = {k: np.zeros(g.w * g.h) for k, g in self.g.items()}
        n = w * h
        self.r, self.p = np.zeros(n), np.zeros(n)
        if version >= 3:
            self.z, self.s, self.precon = np.zeros(n), np.zeros(n), np.zeros(n)
            self.aDiag, self.aPlusX, self.aPlusY = np.zeros(n), np.zeros(n), np.zeros(n)
        self.last = None  # (status, iterations, max_error) of the last solve

    # ---- granular ops (same names as the reference's methods)
    def addInflow(self, x, y, w, h, d, u, v):
        for k, val in (("d", d), ("u", u), ("v", v)):
            self.L.ofl_add_inflow(_p(self.src[k]), self.g[k], self.hx, x, y, x + w, y + h, val,
                                  1 if self.version >= 2 else 0)

    def buildRhs(self):
        self.L.ofl_build_rhs(_p(self.r), _p(self.src["u"]), _p(self.src["v"]), self.w, self.h, self.hx)

    def buildPressureMatrix(self, dt):
        self.L.ofl_build_pressure_matrix(_p(self.aDiag), _p(self.aPlusX), _p(self.aPlusY), self.w, self.h, dt,
                                         self.density, self.hx)

    def buildPreconditioner(self):
        self.L.ofl_build_preconditioner(_p(self.precon), _p(self.aDiag), _p(self.aPlusX), _p(self.aPlusY),
                                        self.w, self.h)

    def applyPreconditioner(self, dst, a):
        self.L.ofl_apply_preconditioner(_p(dst), _p(a), _p(self.precon), _p(self.aPlusX), _p(self.aPlusY),
                                        self.w, self.h)

    def matrixVectorProduct(self, dst, b):
        self.L.ofl_matrix_vector_product(_p(dst), _p(b), _p(self.aDiag), _p(self.aPlusX), _p(self.aPlusY),
                                         self.w, self.h)

    def dotProduct(self, a, b):
        return self.L.ofl_dot_product(_p(a), _p(b), a.size)

    def scaledAdd(self, dst, a, b, s):
        self.L.ofl_scaled_add(_p(dst), _p(a), _p(b), s, dst.size)

    def infinityNorm(self, a):
        return self.L.ofl_infinity_norm(_p(a), a.size)

    def project(self, limit, dt=None):
        it, err = ctypes.c_int(0), ctypes.c_double(0.0)
        if self.version >= 3:
            st = self.L.ofl_project(limit, _p(self.p), _p(self.r), _p(self.z), _p(self.s), _p(self.precon),
                                    _p(self.aDiag), _p(self.aPlusX), _p(self.aPlusY), self.w, self.h,
                                    ctypes.addressof(it), ctypes.addressof(err))
        else:
            st = self.L.ofl_project_gs(limit, dt, self.density, self.hx, _p(self.p), _p(self.r), self.w, self.h,
                                       ctypes.addressof(it), ctypes.addressof(err))
        self.last = (st, it.value, err.value)
        return self.last

    def applyPressure(self, dt):
        self.L.ofl_apply_pressure(_p(self.src["u"]), _p(self.src["v"]), _p(self.p), self.w, self.h, dt,
                                  self.density, self.hx)

    def advect(self, k, dt):
        self.L.ofl_advect(1 if self.version >= 2 else 0, _p(self.dst[k]), _p(self.src[k]), self.g[k],
                          _p(self.src["u"]), self.g["u"], _p(self.src["v"]), self.g["v"], dt, self.hx)

    def flip(self, k):
        self.src[k], self.dst[k] = self.dst[k], self.src[k]

    # ---- FluidSolver::update  v3:433-447 / v2:320-332
    def update(self, dt, limit=600):
        self.buildRhs()
        if self.version >= 3:
            self.buildPressureMatrix(dt)
            self.buildPreconditioner()
            self.project(limit)
        else:
            self.project(limit, dt)
        self.applyPressure(dt)
        for k in "duv":
            self.advect(k, dt)
        for k in "duv":
            self.flip(k)
        return self.last
