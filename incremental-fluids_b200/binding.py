"""ctypes binding of libifl_b200.so and a Python mirror of the reference's FluidSolver.

`FluidSolver` keeps the reference's method names and argument meaning
(FluidSolver(w, h, density), addInflow, update, toImage; v3:401-466) and exposes the
private hot-path methods (buildRhs, buildPressureMatrix, buildPreconditioner, project,
applyPressure, ...) so that parity tests read like calls into the reference class.
All compute happens in the CUDA library; there is no CPU fallback -- if the library
or a CUDA device is missing, construction raises.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_NAME = "libifl_b200.so"


class IflError(RuntimeError):
    pass


class SolveInfo(ctypes.Structure):
    _fields_ = [("status", ctypes.c_int), ("iterations", ctypes.c_int), ("max_error", ctypes.c_double)]

    def astuple(self):
        return (self.status, self.iterations, self.max_error)


# mirrors enum ifl_buf / ifl_field in include/ifl_b200.h
BUF = {name: i for i, name in enumerate(
    ["d.src", "d.dst", "u.src", "u.dst", "v.src", "v.dst", "t.src", "t.dst",
     "r", "p", "z", "s", "precon", "aDiag", "aPlusX", "aPlusY"])}
FIELD = {"d": 0, "u": 1, "v": 2, "t": 3}
# mirrors enum ifl_kernel_class
KERNEL_CLASSES = ["matvec", "axpy2_norm", "precon_fwd", "precon_bwd", "xpay", "scalar", "factor", "assembly",
                  "advect", "gs_sweep", "p2g", "g2p"]

EXPORTS = [
    "ifl_create", "ifl_create_dist", "ifl_dist_plan", "ifl_dist_info", "ifl_dist_barrier", "ifl_dist_selftest",
    "ifl_destroy", "ifl_last_error", "ifl_launch_count", "ifl_stream", "ifl_sync",
    "ifl_profile", "ifl_profile_read", "ifl_debug_sweep_times",
    "ifl_buf_elems", "ifl_upload", "ifl_download", "ifl_fill",
    "ifl_quantity_add_inflow", "ifl_advect", "ifl_flip", "ifl_max_timestep",
    "ifl_set_bodies", "ifl_fill_solid_fields", "ifl_set_boundary_condition", "ifl_extrapolate",
    "ifl_aux_elems", "ifl_aux_download", "ifl_aux_upload",
    "ifl_set_fluid_params", "ifl_ambient_t", "ifl_build_heat_matrix", "ifl_add_buoyancy", "ifl_compute_densities",
    "ifl_add_inflow_t",
    "ifl_particles_capacity", "ifl_particles_count", "ifl_particles_init", "ifl_particles_to_grid",
    "ifl_count_particles", "ifl_prune_particles", "ifl_seed_particles", "ifl_particles_peek",
    "ifl_particles_upload", "ifl_particles_download", "ifl_from_particles",
    "ifl_grid_to_particles", "ifl_quantity_copy", "ifl_quantity_diff", "ifl_quantity_undiff", "ifl_particles_advect",
    "ifl_build_rhs", "ifl_build_pressure_matrix", "ifl_build_preconditioner", "ifl_apply_preconditioner",
    "ifl_matrix_vector_product", "ifl_dot_product", "ifl_scaled_add", "ifl_infinity_norm",
    "ifl_project", "ifl_project_gs", "ifl_apply_pressure",
    "ifl_add_inflow", "ifl_update", "ifl_update_host", "ifl_update_host_slab", "ifl_slab_elems", "ifl_upload_slab", "ifl_download_slab",
]


def library_path():
    # IFL_B200_LIB: an alternative build of the same library (profiles/tri_experiments.sh)
    return os.environ.get("IFL_B200_LIB") or os.path.join(HERE, _LIB_NAME)


def build_library(verbose=False):
    """Compile csrc/*.cu for sm_100a into libifl_b200.so (nvcc cross-compiles without a GPU)."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", os.path.join(HERE, "csrc"), "-j8"], stdout=out, stderr=out)
    return library_path()


_lib = None


def load_library():
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise IflError(path + " is missing: run __graft_entry__.build() (there is no CPU fallback)")
    L = ctypes.CDLL(path)
    vp, ci, cd = ctypes.c_void_p, ctypes.c_int, ctypes.c_double
    L.ifl_last_error.restype = ctypes.c_char_p
    L.ifl_create.argtypes = [ctypes.POINTER(vp), ci, ci, ci, ci]
    L.ifl_create_dist.argtypes = [ctypes.POINTER(vp), ci, ci, ci, ci, ci, ci, ctypes.c_char_p]
    L.ifl_dist_plan.argtypes = [ci, ci, ci, ctypes.POINTER(ci), ctypes.POINTER(ci)]
    L.ifl_dist_info.argtypes = [vp, ctypes.POINTER(ci), ctypes.POINTER(ci), ctypes.POINTER(ci), ctypes.POINTER(ci)]
    L.ifl_dist_barrier.argtypes = [vp]
    L.ifl_dist_selftest.argtypes = [ci, ci, ctypes.c_char_p]
    L.ifl_destroy.argtypes = [vp]
    L.ifl_launch_count.restype = ctypes.c_longlong
    L.ifl_launch_count.argtypes = [vp]
    L.ifl_stream.restype = vp
    L.ifl_stream.argtypes = [vp]
    L.ifl_sync.argtypes = [vp]
    L.ifl_profile.argtypes = [vp, ci]
    L.ifl_profile_read.argtypes = [vp, vp, vp]
    L.ifl_debug_sweep_times.argtypes = [vp, ci, vp, ci]
    L.ifl_buf_elems.restype = ctypes.c_size_t
    L.ifl_buf_elems.argtypes = [vp, ci]
    L.ifl_upload.argtypes = [vp, ci, vp]
    L.ifl_download.argtypes = [vp, ci, vp]
    L.ifl_fill.argtypes = [vp, ci, cd]
    L.ifl_quantity_add_inflow.argtypes = [vp, ci, cd, cd, cd, cd, cd]
    L.ifl_advect.argtypes = [vp, ci, cd]
    L.ifl_flip.argtypes = [vp, ci]
    L.ifl_max_timestep.argtypes = [vp, ctypes.POINTER(cd)]
    L.ifl_set_bodies.argtypes = [vp, vp, ci]
    L.ifl_fill_solid_fields.argtypes = [vp, ci]
    L.ifl_set_boundary_condition.argtypes = [vp]
    L.ifl_extrapolate.argtypes = [vp, ci]
    L.ifl_aux_elems.restype = ctypes.c_size_t
    L.ifl_aux_elems.argtypes = [vp, ci, ci]
    L.ifl_aux_download.argtypes = [vp, ci, ci, vp]
    L.ifl_aux_upload.argtypes = [vp, ci, ci, vp]
    L.ifl_set_fluid_params.argtypes = [vp, cd, cd, cd]
    L.ifl_ambient_t.restype = cd
    L.ifl_ambient_t.argtypes = [vp]
    L.ifl_build_heat_matrix.argtypes = [vp, cd]
    L.ifl_add_buoyancy.argtypes = [vp, cd]
    L.ifl_compute_densities.argtypes = [vp]
    L.ifl_add_inflow_t.argtypes = [vp, cd, cd, cd, cd, cd, cd, cd, cd]
    ll = ctypes.c_longlong
    L.ifl_particles_capacity.argtypes = [vp]
    L.ifl_particles_capacity.restype = ll
    L.ifl_particles_count.argtypes = [vp]
    L.ifl_particles_count.restype = ll
    L.ifl_particles_init.argtypes = [vp, ci]
    L.ifl_particles_to_grid.argtypes = [vp, ctypes.POINTER(ll)]
    L.ifl_count_particles.argtypes = [vp]
    L.ifl_prune_particles.argtypes = [vp]
    L.ifl_seed_particles.argtypes = [vp]
    L.ifl_particles_peek.argtypes = [vp, ci, ll, ll, vp]
    L.ifl_particles_upload.argtypes = [vp, ll, vp, vp, vp, vp, vp, vp]
    L.ifl_particles_download.argtypes = [vp, ctypes.POINTER(ll), vp, vp, vp, vp, vp, vp]
    L.ifl_from_particles.argtypes = [vp, ci]
    L.ifl_grid_to_particles.argtypes = [vp, cd]
    L.ifl_quantity_copy.argtypes = [vp, ci]
    L.ifl_quantity_diff.argtypes = [vp, ci, cd]
    L.ifl_quantity_undiff.argtypes = [vp, ci, cd]
    L.ifl_particles_advect.argtypes = [vp, cd]
    L.ifl_build_rhs.argtypes = [vp]
    L.ifl_build_pressure_matrix.argtypes = [vp, cd, cd]
    L.ifl_build_preconditioner.argtypes = [vp]
    L.ifl_apply_preconditioner.argtypes = [vp, ci, ci]
    L.ifl_matrix_vector_product.argtypes = [vp, ci, ci]
    L.ifl_dot_product.argtypes = [vp, ci, ci, ctypes.POINTER(cd)]
    L.ifl_scaled_add.argtypes = [vp, ci, ci, ci, cd]
    L.ifl_infinity_norm.argtypes = [vp, ci, ctypes.POINTER(cd)]
    L.ifl_project.argtypes = [vp, ci, ctypes.POINTER(SolveInfo)]
    L.ifl_project_gs.argtypes = [vp, ci, cd, cd, ctypes.POINTER(SolveInfo)]
    L.ifl_apply_pressure.argtypes = [vp, cd, cd]
    L.ifl_add_inflow.argtypes = [vp, cd, cd, cd, cd, cd, cd, cd]
    L.ifl_update.argtypes = [vp, cd, cd, ctypes.POINTER(SolveInfo)]
    L.ifl_update_host.argtypes = [vp, cd, cd, vp, vp, vp, ctypes.POINTER(SolveInfo)]
    L.ifl_update_host_slab.argtypes = [vp, cd, cd, vp, vp, vp, ctypes.POINTER(SolveInfo)]
    L.ifl_slab_elems.restype = ctypes.c_size_t
    L.ifl_slab_elems.argtypes = [vp, ci]
    L.ifl_upload_slab.argtypes = [vp, ci, vp]
    L.ifl_download_slab.argtypes = [vp, ci, vp]
    _lib = L
    return L


class BodyRecord(ctypes.Structure):  # struct ifl_body
    _fields_ = [("kind", ctypes.c_int)] + [(n, ctypes.c_double) for n in
                ("pos_x", "pos_y", "scale_x", "scale_y", "theta", "vel_x", "vel_y", "vel_theta")]


class SolidBody:
    """SolidBody (v4:79-149): rigid transform + velocities; update() is the reference's Euler step."""
    kind = 0

    def __init__(self, x, y, sx, sy, theta, vx, vy, vtheta):
        self.posX, self.posY, self.scaleX, self.scaleY, self.theta = x, y, sx, sy, theta
        self.velX, self.velY, self.velTheta = vx, vy, vtheta

    def update(self, timestep):  # v4:140-144
        self.posX += self.velX * timestep
        self.posY += self.velY * timestep
        self.theta += self.velTheta * timestep

    def record(self):
        return BodyRecord(self.kind, self.posX, self.posY, self.scaleX, self.scaleY, self.theta, self.velX,
                          self.velY, self.velTheta)

    def as_row(self):
        """[kind, x, y, sx, sy, theta, vx, vy, vtheta] (flat row format used by test harnesses)."""
        return [float(self.kind), self.posX, self.posY, self.scaleX, self.scaleY, self.theta, self.velX, self.velY,
                self.velTheta]


class SolidBox(SolidBody):  # v4:152-157
    kind = 0


class SolidSphere(SolidBody):  # v4:204-209: SolidSphere(x, y, s, t, vx, vy, vt)
    kind = 1

    def __init__(self, x, y, s, theta, vx, vy, vtheta):
        super().__init__(x, y, s, s, theta, vx, vy, vtheta)


AUX = {"volume": 0, "normalX": 1, "normalY": 2, "phi": 3, "cell": 4, "body": 5}
BUF.update({"uDensity": 16, "vDensity": 17})


class FluidSolver:
    """Drop-in mirror of the reference's FluidSolver for chapters 1-5.

    FluidSolver(w, h, density[, bodies])          v3:401, v4:836
    addInflow(x, y, w, h, d, u, v)                v3:449
    update(timestep)                              v3:433, v5:927
    toImage() -> uint8[h*w*4]                     v3:455

    `bodies` is held by reference like in the reference (v4:612): the caller moves the
    bodies between updates and update() re-reads them.
    """

    def __init__(self, w, h, density, version=3, device=0, bodies=None, rho_soot=None, diffusion=None,
                 rank=0, world=1, rendezvous=None, avg_per_cell=4, init_particles=True):
        """Chapters 1-5: FluidSolver(w, h, density[, bodies]).  Chapters 6-7:
        FluidSolver(w, h, rhoAir, rhoSoot, diffusion, bodies) (v6:921) -- pass rhoAir as
        `density` plus rho_soot= and diffusion=.

        world > 1: row-slab multi-GPU, one process per GPU (ifl_create_dist); `rendezvous` is
        a UNIX socket path shared by the ranks.  Every method is then a collective call.

        Chapter 8: like the reference's constructors (v8:877, 1306-1314) this seeds the particle set on the
        jittered grid (`avg_per_cell` = _AvgPerCell, v8:698) and interpolates the initial fields onto it;
        init_particles=False leaves the set empty (tests that upload their own particles)."""
        self.L = load_library()
        self.w, self.h, self.density, self.version = w, h, density, version
        self.hx = 1.0 / min(w, h)
        self.rank, self.world = rank, world
        ctx = ctypes.c_void_p()
        self.ctx = None
        if world > 1:
            self._chk(self.L.ifl_create_dist(ctypes.byref(ctx), w, h, version, device, rank, world,
                                             rendezvous.encode()))
        else:
            self._chk(self.L.ifl_create(ctypes.byref(ctx), w, h, version, device))
        self.ctx = ctx
        self.last = None
        self.messages = []  # the stdout lines the reference would have printed
        self.bodies = bodies if bodies is not None else []
        if version >= 4:
            self.syncBodies()
        if version >= 6:
            self._chk(self.L.ifl_set_fluid_params(self.ctx, density, rho_soot, diffusion))
        self.last_heat = None
        if version >= 8 and init_particles:
            self.initParticles(avg_per_cell)

    # ---- chapter 8 (FLIP): ParticleQuantities' transfers
    def setParticles(self, posX, posY, props):
        arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in [posX, posY] + list(props)]
        self._particles_keepalive = arrs
        self._chk(self.L.ifl_particles_upload(self.ctx, arrs[0].size, *[a.ctypes.data for a in arrs]))

    def initParticles(self, avg_per_cell=4):
        """What the reference's two constructors do (v8:877, 1306-1314): initParticles + gridToParticles(1.0)."""
        self._chk(self.L.ifl_particles_init(self.ctx, avg_per_cell))

    def maxTimestep(self):  # v1:310
        out = ctypes.c_double()
        self._chk(self.L.ifl_max_timestep(self.ctx, ctypes.byref(out)))
        return out.value

    def particleCount(self):
        return int(self.L.ifl_particles_count(self.ctx))

    def particlesToGrid(self):
        n = ctypes.c_longlong()
        self._chk(self.L.ifl_particles_to_grid(self.ctx, ctypes.byref(n)))
        self.messages.append("Particle count: %d" % n.value)  # v8:926
        return n.value

    def countParticles(self):
        self._chk(self.L.ifl_count_particles(self.ctx))

    def pruneParticles(self):
        self._chk(self.L.ifl_prune_particles(self.ctx))

    def seedParticles(self):
        self._chk(self.L.ifl_seed_particles(self.ctx))

    def peekParticles(self, what, first, n):
        """Raw slots incl. the stale tail: what in posX, posY, d, t, u, v, counts."""
        k = ["posX", "posY", "d", "t", "u", "v", "counts"].index(what)
        out = np.empty(n, dtype=np.int32 if k == 6 else np.float64)
        self._chk(self.L.ifl_particles_peek(self.ctx, k, first, n, out.ctypes.data))
        return out

    def getParticles(self):
        n = ctypes.c_longlong()
        self._chk(self.L.ifl_particles_download(self.ctx, ctypes.byref(n), None, None, None, None, None, None))
        out = [np.empty(n.value) for _ in range(6)]
        self._chk(self.L.ifl_particles_download(self.ctx, ctypes.byref(n), *[a.ctypes.data for a in out]))
        return out[0], out[1], out[2:]

    def fromParticles(self, field):
        self._chk(self.L.ifl_from_particles(self.ctx, FIELD[field]))

    def gridToParticles(self, alpha):
        self._chk(self.L.ifl_grid_to_particles(self.ctx, alpha))

    def copy(self, field):
        self._chk(self.L.ifl_quantity_copy(self.ctx, FIELD[field]))

    def diff(self, field, alpha):
        self._chk(self.L.ifl_quantity_diff(self.ctx, FIELD[field], alpha))

    def undiff(self, field, alpha):
        self._chk(self.L.ifl_quantity_undiff(self.ctx, FIELD[field], alpha))

    def particlesAdvect(self, timestep):
        self._chk(self.L.ifl_particles_advect(self.ctx, timestep))

    # ---- chapters 6+
    def ambientT(self):
        return self.L.ifl_ambient_t(self.ctx)

    def buildHeatDiffusionMatrix(self, timestep):
        self._chk(self.L.ifl_build_heat_matrix(self.ctx, timestep))

    def addBuoyancy(self, timestep):
        self._chk(self.L.ifl_add_buoyancy(self.ctx, timestep))

    def computeDensities(self):
        self._chk(self.L.ifl_compute_densities(self.ctx))

    # ---- chapters 4+
    def syncBodies(self):
        arr = (BodyRecord * max(len(self.bodies), 1))(*[b.record() for b in self.bodies])
        self._chk(self.L.ifl_set_bodies(self.ctx, ctypes.addressof(arr), len(self.bodies)))

    def fillSolidFields(self, field):
        self._chk(self.L.ifl_fill_solid_fields(self.ctx, FIELD[field]))

    def setBoundaryCondition(self):
        self._chk(self.L.ifl_set_boundary_condition(self.ctx))

    def extrapolate(self, field):
        self._chk(self.L.ifl_extrapolate(self.ctx, FIELD[field]))

    def get_aux(self, field, which):
        n = self.L.ifl_aux_elems(self.ctx, FIELD[field], AUX[which])
        if n == 0:
            raise KeyError((field, which))
        out = np.empty(n, dtype=np.uint8 if which in ("cell", "body") else np.float64)
        self._chk(self.L.ifl_aux_download(self.ctx, FIELD[field], AUX[which], out.ctypes.data))
        return out

    def set_aux(self, field, which, arr):
        a = np.ascontiguousarray(arr, dtype=np.uint8 if which in ("cell", "body") else np.float64).ravel()
        if a.size != self.L.ifl_aux_elems(self.ctx, FIELD[field], AUX[which]):
            raise ValueError("size mismatch for %s.%s" % (field, which))
        self._chk(self.L.ifl_aux_upload(self.ctx, FIELD[field], AUX[which], a.ctypes.data))

    def rows(self):
        """Cell rows [row0, row1) of this rank's slab (the whole grid on one GPU)."""
        r0, r1 = ctypes.c_int(), ctypes.c_int()
        self._chk(self.L.ifl_dist_info(self.ctx, None, None, ctypes.byref(r0), ctypes.byref(r1)))
        return r0.value, r1.value

    def barrier(self):
        self._chk(self.L.ifl_dist_barrier(self.ctx))

    def _chk(self, rc):
        if rc != 0:
            raise IflError("libifl_b200 error %d: %s" % (rc, self.L.ifl_last_error().decode()))

    def close(self):
        if self.ctx is not None:
            self.L.ifl_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- data access (FluidQuantity::src()/at(), and the solver's private arrays)
    def get(self, name):
        n = self.L.ifl_buf_elems(self.ctx, BUF[name])
        if n == 0:
            raise KeyError(name)
        out = np.empty(n, dtype=np.float64)
        self._chk(self.L.ifl_download(self.ctx, BUF[name], out.ctypes.data))
        return out

    def set(self, name, arr):
        a = np.ascontiguousarray(arr, dtype=np.float64).ravel()
        if a.size != self.L.ifl_buf_elems(self.ctx, BUF[name]):
            raise ValueError("size mismatch for " + name)
        self._chk(self.L.ifl_upload(self.ctx, BUF[name], a.ctypes.data))

    def launches(self):
        return self.L.ifl_launch_count(self.ctx)

    def stream(self):
        return self.L.ifl_stream(self.ctx)

    def sync(self):
        self._chk(self.L.ifl_sync(self.ctx))

    def profile(self, on=True):
        self._chk(self.L.ifl_profile(self.ctx, 1 if on else 0))

    def profile_read(self):
        """{kernel class: (device ms, launches)} accumulated since profile(True)."""
        ms = np.zeros(len(KERNEL_CLASSES))
        n = np.zeros(len(KERNEL_CLASSES), dtype=np.int64)
        self._chk(self.L.ifl_profile_read(self.ctx, ms.ctypes.data, n.ctypes.data))
        return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(KERNEL_CLASSES)}

    def sweep_times(self, fn):
        """Runs fn() with strip timing armed; returns ns[strips, 16] (start, end, 1/8 checkpoints) of the LAST sweep fn launched."""
        self._chk(self.L.ifl_debug_sweep_times(self.ctx, 1, None, 0))
        fn()
        buf = np.zeros(16 * ((self.h + 31) // 32), dtype=np.uint64)
        n = self.L.ifl_debug_sweep_times(self.ctx, 0, buf.ctypes.data, buf.size)
        if n < 0:
            self._chk(n)
        return buf.reshape(-1, 16)[:n]

    # ---- private hot-path methods of the reference class
    def buildRhs(self):
        self._chk(self.L.ifl_build_rhs(self.ctx))

    def buildPressureMatrix(self, timestep):
        self._chk(self.L.ifl_build_pressure_matrix(self.ctx, timestep, self.density))

    def buildPreconditioner(self):
        self._chk(self.L.ifl_build_preconditioner(self.ctx))

    def applyPreconditioner(self, dst, a):
        self._chk(self.L.ifl_apply_preconditioner(self.ctx, BUF[dst], BUF[a]))

    def matrixVectorProduct(self, dst, b):
        self._chk(self.L.ifl_matrix_vector_product(self.ctx, BUF[dst], BUF[b]))

    def dotProduct(self, a, b):
        out = ctypes.c_double()
        self._chk(self.L.ifl_dot_product(self.ctx, BUF[a], BUF[b], ctypes.byref(out)))
        return out.value

    def scaledAdd(self, dst, a, b, s):
        self._chk(self.L.ifl_scaled_add(self.ctx, BUF[dst], BUF[a], BUF[b], s))

    def infinityNorm(self, a):
        out = ctypes.c_double()
        self._chk(self.L.ifl_infinity_norm(self.ctx, BUF[a], ctypes.byref(out)))
        return out.value

    def _message(self, info, gs):
        what = "change" if gs else "error"
        if info.status == 0:
            return "Exiting solver after %d iterations, maximum %s is %f" % (info.iterations, what, info.max_error)
        if info.status == 1:
            return "Exceeded budget of %d iterations, maximum %s was %f" % (info.iterations, what, info.max_error)
        if self.version >= 6:
            return "Initial guess sufficiently small"  # v6:835
        return None  # v3:355-356 returns silently

    def project(self, limit, timestep=None):
        info = SolveInfo()
        if self.version >= 3:
            self._chk(self.L.ifl_project(self.ctx, limit, ctypes.byref(info)))
        else:
            self._chk(self.L.ifl_project_gs(self.ctx, limit, timestep, self.density, ctypes.byref(info)))
        self._record(info)
        return self.last

    def _record(self, info):
        self.last = info.astuple()
        msg = self._message(info, self.version < 3)
        if msg:
            self.messages.append(msg)

    def applyPressure(self, timestep):
        self._chk(self.L.ifl_apply_pressure(self.ctx, timestep, self.density))

    def advect(self, field, timestep):
        self._chk(self.L.ifl_advect(self.ctx, FIELD[field], timestep))

    def flip(self, field):
        self._chk(self.L.ifl_flip(self.ctx, FIELD[field]))

    def quantityAddInflow(self, field, x0, y0, x1, y1, v):
        self._chk(self.L.ifl_quantity_add_inflow(self.ctx, FIELD[field], x0, y0, x1, y1, v))

    # ---- public surface
    def addInflow(self, x, y, w, h, d, *rest):
        if self.version >= 6:  # addInflow(x, y, w, h, d, t, u, v)  v6:1010
            t, u, v = rest
            self._chk(self.L.ifl_add_inflow_t(self.ctx, x, y, w, h, d, t, u, v))
        else:
            u, v = rest
            self._chk(self.L.ifl_add_inflow(self.ctx, x, y, w, h, d, u, v))

    def _record_update(self, infos):
        """infos: (SolveInfo * 2); chapters 6+ fill [0] = heat solve, [1] = pressure solve."""
        if self.version >= 6:
            self._record(infos[0])
            self.last_heat = self.last
            self._record(infos[1])
        else:
            self._record(infos[0])
        return self.last

    def update(self, timestep):
        if self.version >= 4:
            self.syncBodies()
        infos = (SolveInfo * 2)()
        self._chk(self.L.ifl_update(self.ctx, timestep, self.density, infos))
        if self.version >= 8:  # particlesToGrid prints first (v8:926), then the two solves
            self.messages.append("Particle count: %d" % self.particleCount())
        return self._record_update(infos)

    def update_host(self, timestep, d, u, v):
        """update() on HOST arrays (in/out, reference layout): H2D + step + D2H."""
        infos = (SolveInfo * 2)()  # chapters 6+ report two solves (heat, pressure)
        self._chk(self.L.ifl_update_host(self.ctx, timestep, self.density, d.ctypes.data, u.ctypes.data,
                                         v.ctypes.data, infos))
        return self._record_update(infos)

    def slab_elems(self, name):
        return self.L.ifl_slab_elems(self.ctx, BUF[name])

    def set_slab(self, name, arr):
        self._chk(self.L.ifl_upload_slab(self.ctx, BUF[name], arr.ctypes.data))

    def get_slab(self, name, out):
        self._chk(self.L.ifl_download_slab(self.ctx, BUF[name], out.ctypes.data))

    def update_host_slab(self, timestep, d, u, v):
        """update() on HOST arrays holding this rank's slab rows only (in/out)."""
        infos = (SolveInfo * 2)()
        self._chk(self.L.ifl_update_host_slab(self.ctx, timestep, self.density, d.ctypes.data, u.ctypes.data,
                                              v.ctypes.data, infos))
        return self._record_update(infos)

    def toImage(self):
        d = self.get("d.src")
        shade = ((1.0 - d) * 255.0).astype(np.int64)  # (int) truncation, v3:457
        shade = np.maximum(np.minimum(shade, 255), 0).astype(np.uint8)
        rgba = np.empty((d.size, 4), dtype=np.uint8)
        rgba[:, 0] = rgba[:, 1] = rgba[:, 2] = shade
        rgba[:, 3] = 0xFF
        return rgba.ravel()
