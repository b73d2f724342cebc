"""ctypes binding of oracle/_ref/libflipref_{golden,fast}.so (the unmodified reference engine
driven through oracle/ref_shim.cpp).  TEST INFRASTRUCTURE ONLY: imported by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs, never by the
product package."""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

STAGES = dict(obstacles=0, liquid_sdf=1, p2g=2, extrapolate_a=3, save=4, body_force=5, pressure=6,
              extrapolate_b=7, constrain=8, g2p=9, advance=10, tail=11)
ARRAYS = dict(U=0, V=1, W=2, validU=3, validV=4, validW=5, liquid_phi=6, solid_phi=7,
              weightU=8, weightV=9, weightW=10, weightC=11, savedU=12, savedV=13, savedW=14, near_solid=15,
              solidU=16, solidV=17, solidW=18)


def lib_path(kind="golden"):
    return os.path.join(_HERE, "_ref", f"libflipref_{kind}.so")


def available(kind="golden"):
    return os.path.exists(lib_path(kind))


_libs = {}


def _load(kind):
    if kind in _libs:
        return _libs[kind]
    L = C.CDLL(lib_path(kind))
    L.ref_create.restype = C.c_void_p
    L.ref_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double]
    L.ref_destroy.argtypes = [C.c_void_p]
    L.ref_last_error.restype = C.c_char_p
    L.ref_last_error.argtypes = [C.c_void_p]
    L.ref_set_threads.argtypes = [C.c_int]
    L.ref_get_threads.restype = C.c_int
    L.ref_add_body_force.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
    L.ref_set_pressure_tolerance.argtypes = [C.c_void_p, C.c_double]
    L.ref_load_particles.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.ref_initialize.argtypes = [C.c_void_p]
    L.ref_update.argtypes = [C.c_void_p, C.c_double]
    for f in ("ref_num_particles", "ref_current_frame", "ref_last_substeps", "ref_pcg_iterations", "ref_num_fluid_cells"):
        getattr(L, f).argtypes = [C.c_void_p]
        getattr(L, f).restype = C.c_int
    L.ref_pcg_error.argtypes = [C.c_void_p]
    L.ref_pcg_error.restype = C.c_double
    L.ref_liquid_sdf_radius.argtypes = [C.c_void_p]
    L.ref_liquid_sdf_radius.restype = C.c_double
    L.ref_get_particles.argtypes = [C.c_void_p, C.c_void_p]
    L.ref_set_particles.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.ref_begin_frame.argtypes = [C.c_void_p, C.c_double]
    L.ref_begin_substep.argtypes = [C.c_void_p]
    L.ref_begin_substep.restype = C.c_double
    L.ref_end_substep.argtypes = [C.c_void_p]
    L.ref_end_substep.restype = C.c_int
    L.ref_end_frame.argtypes = [C.c_void_p]
    L.ref_stage.argtypes = [C.c_void_p, C.c_int, C.c_double]
    L.ref_stage_time.argtypes = [C.c_void_p, C.c_int]
    L.ref_stage_time.restype = C.c_double
    L.ref_array_bytes.argtypes = [C.c_void_p, C.c_int]
    L.ref_array_bytes.restype = C.c_long
    L.ref_get_array.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.ref_set_array.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.ref_near_solid_dims.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 3
    L.ref_update_weight_grid.argtypes = [C.c_void_p]
    L.ref_invalidate_weight_grid.argtypes = [C.c_void_p]
    L.ref_sample_velocity.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.ref_sample_solid_phi.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.ref_add_mesh_fluid_box.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.ref_add_fluid_source_box.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.ref_enable_fluid_source.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.ref_mesh_level_set.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.ref_add_mesh_fluid_mesh.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_double)]
    L.ref_set_step_settings.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_int]
    L.ref_set_extreme_velocity_removal.argtypes = [C.c_void_p, C.c_int]
    L.ref_set_marker_particle_scale.argtypes = [C.c_void_p, C.c_double]
    L.ref_add_obstacle_box.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.ref_remove_obstacle.argtypes = [C.c_void_p, C.c_int]
    L.ref_add_obstacle_mesh.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    L.ref_animate_obstacle_mesh.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    L.ref_set_boundary_friction.argtypes = [C.c_void_p, C.c_double]
    L.ref_set_obstacle_friction.argtypes = [C.c_void_p, C.c_int, C.c_double]
    L.ref_face_friction.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    L.ref_animate_obstacle_box.argtypes = [C.c_void_p, C.c_int] + [C.POINTER(C.c_double)] * 5
    L.ref_constrain_fluid_source_velocity.argtypes = [C.c_void_p, C.c_int, C.c_int]
    L.ref_isomesh.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.ref_get_isomesh.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.ref_mesher_scalar_field.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    _libs[kind] = L
    return L


class RefEngine:
    """One reference FluidSimulation, created headless with particles injected before initialize()."""

    def __init__(self, dims, dx, pos, vel, gravity=(0.0, -25.0, 0.0), kind="golden", threads=None, tol=None):
        self.L = _load(kind)
        self.dims = tuple(int(d) for d in dims)
        self.dx = float(dx)
        if threads is not None:
            self.L.ref_set_threads(int(threads))
        self.h = self.L.ref_create(self.dims[0], self.dims[1], self.dims[2], self.dx)
        self.L.ref_add_body_force(self.h, *[float(g) for g in gravity])
        if tol is not None:
            self.L.ref_set_pressure_tolerance(self.h, float(tol))
        pos = np.ascontiguousarray(pos, dtype=np.float32)
        vel = np.ascontiguousarray(vel, dtype=np.float32)
        if pos.shape[0] > 0:
            self._check(self.L.ref_load_particles(self.h, pos.shape[0], pos.ctypes.data, vel.ctypes.data))
        self._check(self.L.ref_initialize(self.h))

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(self.L.ref_last_error(self.h).decode())

    def close(self):
        if self.h:
            self.L.ref_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- whole frames through the public API
    def update(self, dt):
        self._check(self.L.ref_update(self.h, float(dt)))

    @property
    def num_particles(self):
        return self.L.ref_num_particles(self.h)

    def particles(self):
        a = np.empty((self.num_particles, 6), dtype=np.float32)
        self.L.ref_get_particles(self.h, a.ctypes.data)
        return a

    def set_particles(self, aos):
        aos = np.ascontiguousarray(aos, dtype=np.float32)
        assert self.L.ref_set_particles(self.h, aos.shape[0], aos.ctypes.data) == 0

    @property
    def pcg_iterations(self):
        return self.L.ref_pcg_iterations(self.h)

    @property
    def pcg_error(self):
        return self.L.ref_pcg_error(self.h)

    @property
    def num_fluid_cells(self):
        return self.L.ref_num_fluid_cells(self.h)

    @property
    def radius(self):
        return self.L.ref_liquid_sdf_radius(self.h)

    @property
    def substeps(self):
        return self.L.ref_last_substeps(self.h)

    # ---- stage-wise stepping
    def begin_frame(self, dt):
        self.L.ref_begin_frame(self.h, float(dt))

    def begin_substep(self):
        return self.L.ref_begin_substep(self.h)

    def end_substep(self):
        return bool(self.L.ref_end_substep(self.h))

    def end_frame(self):
        self.L.ref_end_frame(self.h)

    def stage(self, name, dt):
        self._check(self.L.ref_stage(self.h, STAGES[name], float(dt)))
        return self.L.ref_stage_time(self.h, STAGES[name])

    def shape_of(self, name):
        I, J, K = self.dims
        if name in ("U", "validU", "weightU", "savedU", "solidU"):
            return (K, J, I + 1)
        if name in ("V", "validV", "weightV", "savedV", "solidV"):
            return (K, J + 1, I)
        if name in ("W", "validW", "weightW", "savedW", "solidW"):
            return (K + 1, J, I)
        if name in ("liquid_phi", "weightC"):
            return (K, J, I)
        if name == "solid_phi":
            return (K + 1, J + 1, I + 1)
        if name == "near_solid":
            gi, gj, gk = C.c_int(), C.c_int(), C.c_int()
            self.L.ref_near_solid_dims(self.h, C.byref(gi), C.byref(gj), C.byref(gk))
            return (gk.value, gj.value, gi.value)
        raise KeyError(name)

    def array(self, name):
        """Arrays come back shaped (k, j, i) — i fastest, as in Array3d (array3d.h:425-428)."""
        which = ARRAYS[name]
        dt = np.uint8 if name.startswith("valid") or name == "near_solid" else np.float32
        shape = self.shape_of(name)
        a = np.empty(shape, dtype=dt)
        assert self.L.ref_array_bytes(self.h, which) == a.nbytes, (name, self.L.ref_array_bytes(self.h, which), a.nbytes)
        assert self.L.ref_get_array(self.h, which, a.ctypes.data) == 0
        return a

    def set_array(self, name, a):
        which = ARRAYS[name]
        dt = np.uint8 if name.startswith("valid") or name == "near_solid" else np.float32
        a = np.ascontiguousarray(a, dtype=dt)
        assert self.L.ref_array_bytes(self.h, which) == a.nbytes
        assert self.L.ref_set_array(self.h, which, a.ctypes.data) == 0

    def update_weight_grid(self, force=False):
        if force:       # the solid SDF was overwritten with set_array
            self.L.ref_invalidate_weight_grid(self.h)
        self.L.ref_update_weight_grid(self.h)

    def sample_velocity(self, pos):
        pos = np.ascontiguousarray(pos, dtype=np.float32)
        out = np.empty_like(pos)
        self.L.ref_sample_velocity(self.h, pos.shape[0], pos.ctypes.data, out.ctypes.data)
        return out

    def add_mesh_fluid_box(self, lo, hi, velocity=(0.0, 0.0, 0.0)):
        """FluidSimulation::addMeshFluid(MeshObject) with the box mesh FluidManager builds."""
        a, b, v = (C.c_double * 3)(*lo), (C.c_double * 3)(*hi), (C.c_double * 3)(*velocity)
        self._check(self.L.ref_add_mesh_fluid_box(self.h, a, b, v))

    def add_fluid_source_box(self, lo, hi, velocity=(0.0, 0.0, 0.0), outflow=False):
        """FluidSimulation::addMeshFluidSource with a static box inflow / outflow MeshFluidSource; returns its handle."""
        a, b, v = (C.c_double * 3)(*lo), (C.c_double * 3)(*hi), (C.c_double * 3)(*velocity)
        idx = self.L.ref_add_fluid_source_box(self.h, 1 if outflow else 0, a, b, v)
        assert idx >= 0, self.L.ref_last_error(self.h)
        return idx

    def add_mesh_fluid_mesh(self, vertices, triangles, velocity=(0.0, 0.0, 0.0)):
        """FluidSimulation::addMeshFluid with a static closed triangle mesh."""
        v = np.ascontiguousarray(vertices, dtype=np.float32)
        t = np.ascontiguousarray(triangles, dtype=np.int32)
        self._check(self.L.ref_add_mesh_fluid_mesh(self.h, v.ctypes.data, v.shape[0], t.ctypes.data, t.shape[0], (C.c_double * 3)(*velocity)))

    def set_step_settings(self, cfl=0, picflip=-1.0, min_steps=0, max_steps=0):
        """setCFLConditionNumber / setPICFLIPRatio / setMin-, setMaxTimeStepsPerFrame (0 / negative: left alone)."""
        self._check(self.L.ref_set_step_settings(self.h, int(cfl), float(picflip), int(min_steps), int(max_steps)))

    def set_extreme_velocity_removal(self, on=True):
        self.L.ref_set_extreme_velocity_removal(self.h, 1 if on else 0)

    def set_marker_particle_scale(self, scale):
        self._check(self.L.ref_set_marker_particle_scale(self.h, float(scale)))

    def add_obstacle_box(self, lo, hi):
        """FluidSimulation::addMeshObstacle with a static box MeshObject; returns its handle."""
        a, b = (C.c_double * 3)(*lo), (C.c_double * 3)(*hi)
        idx = self.L.ref_add_obstacle_box(self.h, a, b)
        assert idx >= 0, self.L.ref_last_error(self.h)
        return idx

    def animate_obstacle_box(self, idx, lo, hi, off_prev, off_cur, off_next):
        """MeshObject::updateMeshAnimated with the box [lo,hi] moved by the three offsets (previous / current / next frame)."""
        d3 = C.c_double * 3
        self._check(self.L.ref_animate_obstacle_box(self.h, int(idx), d3(*lo), d3(*hi), d3(*off_prev), d3(*off_cur), d3(*off_next)))

    def add_obstacle_mesh(self, vertices, triangles):
        """FluidSimulation::addMeshObstacle with a closed triangle mesh; returns its index in the shim's list."""
        v = np.ascontiguousarray(vertices, dtype=np.float32)
        t = np.ascontiguousarray(triangles, dtype=np.int32)
        idx = self.L.ref_add_obstacle_mesh(self.h, v.ctypes.data, v.shape[0], t.ctypes.data, t.shape[0])
        if idx < 0:
            raise RuntimeError(self.L.ref_last_error(self.h).decode())
        return idx

    def animate_obstacle_mesh(self, idx, prev, cur, nxt, triangles):
        """MeshObject::updateMeshAnimated with the vertices of the previous / current / next frame."""
        a = [np.ascontiguousarray(x, dtype=np.float32) for x in (prev, cur, nxt)]
        t = np.ascontiguousarray(triangles, dtype=np.int32)
        self._check(self.L.ref_animate_obstacle_mesh(self.h, int(idx), a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data,
                                                     a[0].shape[0], t.ctypes.data, t.shape[0]))

    def set_boundary_friction(self, f):
        self._check(self.L.ref_set_boundary_friction(self.h, float(f)))

    def set_obstacle_friction(self, idx, f):
        self._check(self.L.ref_set_obstacle_friction(self.h, int(idx), float(f)))

    def face_friction(self):
        """dict(U, V, W): _getFaceFrictionU/V/W of every face (after the solid SDF has been built)."""
        out = {}
        for comp, n in enumerate("UVW"):
            a = np.empty(self.shape_of("solid" + n), dtype=np.float32)
            self._check(self.L.ref_face_friction(self.h, comp, a.ctypes.data))
            out[n] = a
        return out

    def remove_obstacle(self, idx):
        self._check(self.L.ref_remove_obstacle(self.h, int(idx)))

    def constrain_fluid_source_velocity(self, idx, on=True):
        self.L.ref_constrain_fluid_source_velocity(self.h, int(idx), 1 if on else 0)

    def enable_fluid_source(self, idx, on=True):
        self.L.ref_enable_fluid_source(self.h, int(idx), 1 if on else 0)

    def isomesh(self, subdivisions=1, smooth_iterations=-1):
        """(vertices, triangles) of the surface getIsomesh() would return for the current particles."""
        nv, nt = C.c_int(), C.c_int()
        self._check(self.L.ref_isomesh(self.h, int(subdivisions), int(smooth_iterations), C.byref(nv), C.byref(nt)))
        v = np.empty((nv.value, 3), dtype=np.float32)
        t = np.empty((nt.value, 3), dtype=np.int32)
        self.L.ref_get_isomesh(self.h, v.ctypes.data, t.ctypes.data)
        return v, t

    def mesher_scalar_field(self, subdivisions=1):
        I, J, K = self.dims
        out = np.empty((K * subdivisions + 1, J * subdivisions + 1, I * subdivisions + 1), dtype=np.float32)
        self._check(self.L.ref_mesher_scalar_field(self.h, int(subdivisions), out.ctypes.data))
        return out

    def sample_solid_phi(self, pos):
        pos = np.ascontiguousarray(pos, dtype=np.float32)
        out = np.empty(pos.shape[0], dtype=np.float32)
        self.L.ref_sample_solid_phi(self.h, pos.shape[0], pos.ctypes.data, out.ctypes.data)
        return out

    def step_stagewise(self, dt_frame, on_stage=None):
        """One frame, stage by stage, equivalent to update(dt_frame). on_stage(name, dt_sub, self)
        is called after every stage. Returns the list of substep lengths."""
        self.begin_frame(dt_frame)
        dts = []
        more = True
        while more:
            dt = self.begin_substep()
            dts.append(dt)
            for name in ("obstacles", "liquid_sdf", "p2g", "extrapolate_a", "save", "body_force", "pressure",
                         "extrapolate_b", "constrain", "g2p", "advance", "tail"):
                self.stage(name, dt)
                if on_stage is not None:
                    on_stage(name, dt, self)
            more = self.end_substep()
        self.end_frame()
        return dts


def mesh_level_set(dims, dx, vertices, triangles, band=3, kind="golden"):
    """MeshLevelSet::fastCalculateSignedDistanceField of a triangle mesh: nodal phi, shape (K+1, J+1, I+1)."""
    L = _load(kind)
    v = np.ascontiguousarray(vertices, dtype=np.float32)
    t = np.ascontiguousarray(triangles, dtype=np.int32)
    I, J, K = (int(d) for d in dims)
    out = np.empty((K + 1, J + 1, I + 1), dtype=np.float32)
    rc = L.ref_mesh_level_set(I, J, K, float(dx), v.ctypes.data, v.shape[0], t.ctypes.data, t.shape[0], int(band), out.ctypes.data)
    assert rc == 0
    return out
