"""ctypes loaders for the checkers under oracle/.  TEST INFRASTRUCTURE ONLY.

Two checkers live here:

* ``Oracle``    -- oracle/_build/libppc_oracle.so, our CPU restatement (oracle/ppc_oracle.c).
* ``Reference`` -- oracle/_ref/libref.so, the UNMODIFIED reference sources compiled by oracle/Makefile,
                   loaded behind oracle/_ref/libref_shim.so (pair-list interposer, oracle/ref_shim.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs import this
module.  The product package (particlephysicsc_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "_build", "libppc_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libref.so")
SHIM_SO = os.path.join(HERE, "_ref", "libref_shim.so")

f32p = C.POINTER(C.c_float)
u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
i32p = C.POINTER(C.c_int32)


def build(targets=("oracle", "ref")) -> None:
    """Compile the checkers (gcc only; seconds)."""
    subprocess.run(["make", "-s", "-C", HERE, *targets], check=True)


class ParticlesT(C.Structure):
    """struct Particles_t, reference includes/structs.h:9-22 (72 bytes)."""

    _fields_ = [
        ("v_x", f32p), ("v_y", f32p), ("p_x", f32p), ("p_y", f32p),
        ("colors", u8p), ("mass", f32p), ("radius", f32p),
        ("length", C.c_size_t), ("size", C.c_size_t),
    ]


class Vector2(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float)]


@dataclass
class State:
    """The seven SoA arrays of struct Particles_t as numpy arrays (valid range = all of them)."""

    px: np.ndarray
    py: np.ndarray
    vx: np.ndarray
    vy: np.ndarray
    mass: np.ndarray
    radius: np.ndarray
    rgba: np.ndarray  # (n, 4) uint8

    @property
    def n(self) -> int:
        return int(self.px.shape[0])

    def copy(self) -> "State":
        return State(*(a.copy() for a in (self.px, self.py, self.vx, self.vy, self.mass, self.radius, self.rgba)))

    @staticmethod
    def empty(n: int) -> "State":
        z = lambda: np.zeros(n, dtype=np.float32)
        return State(z(), z(), z(), z(), z(), z(), np.zeros((n, 4), dtype=np.uint8))

    def as_struct(self) -> ParticlesT:
        for a in (self.px, self.py, self.vx, self.vy, self.mass, self.radius, self.rgba):
            assert a.flags["C_CONTIGUOUS"]
        p = lambda a: a.ctypes.data_as(f32p)
        return ParticlesT(p(self.vx), p(self.vy), p(self.px), p(self.py), self.rgba.ctypes.data_as(u8p),
                          p(self.mass), p(self.radius), self.n, self.n)


def _fp(a):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(f32p)


class Oracle:
    """CPU restatement (oracle/ppc_oracle.c)."""

    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build(("oracle",))
        L = self.lib = C.CDLL(ORACLE_SO)
        L.ppc_oracle_update.argtypes = [f32p, f32p, f32p, f32p, f32p, u8p, C.c_size_t, C.c_float, C.c_int]
        L.ppc_oracle_walls.argtypes = [f32p, f32p, f32p, f32p, f32p, C.c_size_t, C.c_uint16, C.c_uint16]
        L.ppc_oracle_grid_dims.argtypes = [f32p, C.c_size_t, C.c_uint16, C.c_uint16, f32p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.ppc_oracle_grid_build.argtypes = [f32p, f32p, C.c_size_t, C.c_float, C.c_int, C.c_int, u32p, u32p, u32p]
        L.ppc_oracle_pairs.argtypes = [f32p, f32p, f32p, C.c_size_t, C.c_int, C.c_int, u32p, u32p, u32p, C.POINTER(i32p)]
        L.ppc_oracle_pairs.restype = C.c_size_t
        L.ppc_oracle_free.argtypes = [C.c_void_p]
        rs = [f32p, f32p, f32p, f32p, f32p, f32p, C.c_size_t, i32p, C.c_size_t, C.c_float]
        L.ppc_oracle_resolve_sequential.argtypes = rs
        L.ppc_oracle_resolve_jacobi.argtypes = rs
        L.ppc_oracle_collisions_check_grid.argtypes = [f32p, f32p, f32p, f32p, f32p, f32p, C.c_size_t, C.c_uint16, C.c_uint16,
                                                       C.c_float, C.c_int, u32p, u32p, u32p, C.POINTER(i32p), C.POINTER(C.c_size_t)]
        L.ppc_oracle_substep.argtypes = [f32p, f32p, f32p, f32p, f32p, f32p, u8p, C.c_size_t, C.c_uint16, C.c_uint16, C.c_float,
                                         C.c_int, C.c_int]
        L.ppc_oracle_stats.argtypes = [f32p, f32p, f32p, f32p, f32p, f32p, C.c_size_t, C.c_uint16, C.c_uint16, C.POINTER(C.c_double)]
        L.ppc_oracle_set_tilegs.argtypes = [C.c_int, C.c_int, C.c_int]
        L.ppc_oracle_tilegs_order.argtypes = [i32p, C.c_size_t, C.c_size_t, u32p, u32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u32p]
        L.ppc_oracle_resolve_permuted.argtypes = rs + [u32p]

    # -- pieces ---------------------------------------------------------------------------------
    def update(self, s: State, dt: float, colour: bool = True) -> None:
        self.lib.ppc_oracle_update(_fp(s.px), _fp(s.py), _fp(s.vx), _fp(s.vy), _fp(s.mass), s.rgba.ctypes.data_as(u8p), s.n, dt,
                                   int(colour))

    def walls(self, s: State, W: int, H: int) -> None:
        self.lib.ppc_oracle_walls(_fp(s.px), _fp(s.py), _fp(s.vx), _fp(s.vy), _fp(s.radius), s.n, W, H)

    def grid_dims(self, s: State, W: int, H: int):
        cs, cols, rows = C.c_float(), C.c_int(), C.c_int()
        self.lib.ppc_oracle_grid_dims(_fp(s.radius), s.n, W, H, C.byref(cs), C.byref(cols), C.byref(rows))
        return cs.value, cols.value, rows.value

    def grid_build(self, s: State, W: int, H: int):
        cs, cols, rows = self.grid_dims(s, W, H)
        keys = np.zeros(s.n, np.uint32)
        order = np.zeros(s.n, np.uint32)
        start = np.zeros(cols * rows + 1, np.uint32)
        self.lib.ppc_oracle_grid_build(_fp(s.px), _fp(s.py), s.n, cs, cols, rows, keys.ctypes.data_as(u32p),
                                       order.ctypes.data_as(u32p), start.ctypes.data_as(u32p))
        return dict(cell_size=cs, cols=cols, rows=rows, keys=keys, order=order, cell_start=start)

    def pairs(self, s: State, grid) -> np.ndarray:
        out = i32p()
        n = self.lib.ppc_oracle_pairs(_fp(s.px), _fp(s.py), _fp(s.radius), s.n, grid["cols"], grid["rows"],
                                      grid["keys"].ctypes.data_as(u32p), grid["order"].ctypes.data_as(u32p),
                                      grid["cell_start"].ctypes.data_as(u32p), C.byref(out))
        arr = np.ctypeslib.as_array(out, shape=(n, 2)).copy() if n else np.zeros((0, 2), np.int32)
        self.lib.ppc_oracle_free(out)
        return arr

    def resolve(self, s: State, pairs: np.ndarray, dt: float, jacobi: bool = False) -> None:
        pairs = np.ascontiguousarray(pairs, dtype=np.int32)
        fn = self.lib.ppc_oracle_resolve_jacobi if jacobi else self.lib.ppc_oracle_resolve_sequential
        fn(_fp(s.px), _fp(s.py), _fp(s.vx), _fp(s.vy), _fp(s.mass), _fp(s.radius), s.n, pairs.ctypes.data_as(i32p), len(pairs), dt)

    def set_tilegs(self, wc: int, tile_r: int = 8, strip_limit: int = 921) -> None:
        """Tile geometry of the mode-3 model (the device's: ppc_debug_gs_params)."""
        self.lib.ppc_oracle_set_tilegs(wc, tile_r, strip_limit)

    def tilegs_order(self, pairs: np.ndarray, grid, n: int, wc: int, tile_r: int = 8, strip_limit: int = 921, row0: int = 0) -> np.ndarray:
        """The order in which the device's tile Gauss-Seidel resolve applies the velocity impulses of `pairs`."""
        pairs = np.ascontiguousarray(pairs, dtype=np.int32)
        perm = np.zeros(len(pairs), np.uint32)
        self.lib.ppc_oracle_tilegs_order(pairs.ctypes.data_as(i32p), len(pairs), n, grid["keys"].ctypes.data_as(u32p),
                                         grid["cell_start"].ctypes.data_as(u32p), grid["cols"], grid["rows"], row0, wc, tile_r, strip_limit,
                                         perm.ctypes.data_as(u32p))
        return perm

    # -- whole entry points ---------------------------------------------------------------------
    def collisions_check_grid(self, s: State, W: int, H: int, dt: float, mode: int = 0, want_grid: bool = False):
        """mode 0 = sequential (reference), 1 = Jacobi model of the device, 2 = walls + grid only, 3 = tile Gauss-Seidel model."""
        keys = order = start = None
        kp = op = sp = None
        if want_grid and s.n >= 2:
            _, cols, rows = self.grid_dims(s, W, H)
            keys, order, start = np.zeros(s.n, np.uint32), np.zeros(s.n, np.uint32), np.zeros(cols * rows + 1, np.uint32)
            kp, op, sp = (a.ctypes.data_as(u32p) for a in (keys, order, start))
        out, cnt = i32p(), C.c_size_t()
        self.lib.ppc_oracle_collisions_check_grid(_fp(s.px), _fp(s.py), _fp(s.vx), _fp(s.vy), _fp(s.mass), _fp(s.radius), s.n, W, H,
                                                  dt, mode, kp, op, sp, C.byref(out), C.byref(cnt))
        pairs = np.ctypeslib.as_array(out, shape=(cnt.value, 2)).copy() if cnt.value else np.zeros((0, 2), np.int32)
        if out:
            self.lib.ppc_oracle_free(out)
        return dict(keys=keys, order=order, cell_start=start, pairs=pairs)

    def substep(self, s: State, W: int, H: int, dt: float, mode: int = 0, colour: bool = False) -> None:
        self.lib.ppc_oracle_substep(_fp(s.px), _fp(s.py), _fp(s.vx), _fp(s.vy), _fp(s.mass), _fp(s.radius),
                                    s.rgba.ctypes.data_as(u8p), s.n, W, H, dt, mode, int(colour))

    def stats(self, s: State, W: int, H: int):
        out = (C.c_double * 5)()
        self.lib.ppc_oracle_stats(_fp(s.px), _fp(s.py), _fp(s.vx), _fp(s.vy), _fp(s.mass), _fp(s.radius), s.n, W, H, out)
        return dict(ke=out[0], mom_x=out[1], mom_y=out[2], overlap=out[3], contacts=int(out[4]))


def reference_available() -> bool:
    return os.path.exists(REF_SO) and os.path.exists(SHIM_SO)


class Reference:
    """The unmodified reference solver (oracle/_ref/libref.so) behind the pair-list interposer."""

    def __init__(self):
        if not reference_available():
            raise FileNotFoundError("oracle/_ref/libref.so missing: run `make -C oracle ref` where /root/reference exists")
        self.shim = C.CDLL(SHIM_SO, mode=C.RTLD_GLOBAL)  # must precede libref.so: it wins the PLT lookup
        L = self.lib = C.CDLL(REF_SO, mode=C.RTLD_GLOBAL)
        self.shim.ppc_capture_set_real.argtypes = [C.c_void_p]
        self.shim.ppc_capture_set_real(C.cast(L.particlesResolveVelocity, C.c_void_p))
        self.shim.ppc_capture_count.restype = C.c_size_t
        self.shim.ppc_capture_calls.restype = C.c_size_t
        self.shim.ppc_capture_copy.argtypes = [i32p, C.c_size_t]
        self.shim.ppc_capture_copy.restype = C.c_size_t
        self.shim.ppc_capture_set_order.argtypes = [C.c_int, C.c_uint64]
        PP = C.POINTER(ParticlesT)
        L.particlesUpdate.argtypes = [PP, C.c_float]
        L.particlesUpdate.restype = None
        L.particlesCollisionsCheck_Grid.argtypes = [PP, C.c_uint16, C.c_uint16, C.c_float]
        L.particlesCollisionsCheck_Grid.restype = None
        L.particlesResolveVelocity.argtypes = [PP, i32p, C.c_size_t, C.c_float]
        L.particlesResolveVelocity.restype = None
        L.particlesCreate.argtypes = [C.c_size_t]
        L.particlesCreate.restype = ParticlesT
        L.particlesFree.argtypes = [PP]
        L.particlesFree.restype = None
        L.particlesAddParticle.argtypes = [PP, Vector2, C.c_float, C.c_float, C.c_uint8]
        L.particlesAddParticle.restype = None
        L.particlesRemoveParticle.argtypes = [PP]
        L.particlesRemoveParticle.restype = None

    def set_pair_order(self, mode: int, seed: int = 1) -> None:
        """0 = as built, 1 = reversed, 2 = shuffled (noise-floor experiments only)."""
        self.shim.ppc_capture_set_order(mode, seed)

    def update(self, s: State, dt: float) -> None:
        st = s.as_struct()
        self.lib.particlesUpdate(C.byref(st), dt)

    def collisions_check_grid(self, s: State, W: int, H: int, dt: float, capture: bool = False):
        st = s.as_struct()
        self.shim.ppc_capture_reset()
        self.shim.ppc_capture_enable(int(capture))
        self.lib.particlesCollisionsCheck_Grid(C.byref(st), W, H, dt)
        self.shim.ppc_capture_enable(0)
        if not capture:
            return None
        n = self.shim.ppc_capture_count()
        pairs = np.zeros((n, 2), np.int32)
        if n:
            self.shim.ppc_capture_copy(pairs.ctypes.data_as(i32p), n)
        return pairs

    def substep(self, s: State, W: int, H: int, dt: float) -> None:
        """src/main.c:92-98."""
        if s.n == 0:
            return
        st = s.as_struct()
        self.lib.particlesUpdate(C.byref(st), dt)
        self.lib.particlesCollisionsCheck_Grid(C.byref(st), W, H, dt)
