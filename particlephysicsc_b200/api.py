"""ctypes binding of libparticlephysicsc_b200.so -- the host-side mirror of the reference's C interface.

Names, argument meaning and error behaviour follow the reference's headers (includes/particles_arrays.h:16-19,
includes/particles_solver.h:9-10, includes/structs.h:9-22): ``particlesCreate``, ``particlesFree``,
``particlesAddParticle``, ``particlesRemoveParticle``, ``particlesUpdate``, ``particlesCollisionsCheck_Grid`` over
a by-value ``struct Particles_t``.  The ``ppc_*`` functions are the headless-harness extensions declared in
include/ppc_b200.h.

There is no CPU fallback: if the shared library is missing the import fails, and the library itself aborts when
no CUDA device is present.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PPC_LIB") or os.path.join(HERE, "lib", "libparticlephysicsc_b200.so")   # PPC_LIB: kernel-variant experiments

f32p = C.POINTER(C.c_float)
u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
i32p = C.POINTER(C.c_int32)

PPC_SYNC_EVERY_CALL, PPC_SYNC_FRAME, PPC_SYNC_MANUAL = 0, 1, 2
PPC_DL_POS, PPC_DL_VEL, PPC_DL_COLOR, PPC_DL_ALL = 1, 2, 4, 7
PPC_DL_IDS, PPC_DL_MASS_RADIUS = 8, 16


class Particles_t(C.Structure):
    """struct Particles_t (reference includes/structs.h:9-22): 72 bytes, field order is ABI."""

    _fields_ = [
        ("v_x", f32p), ("v_y", f32p), ("p_x", f32p), ("p_y", f32p),
        ("colors", u8p), ("mass", f32p), ("radius", f32p),
        ("length", C.c_size_t), ("size", C.c_size_t),
    ]


class Vector2(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float)]


class GridInfo(C.Structure):
    _fields_ = [("max_radius", C.c_float), ("cell_size", C.c_float), ("cols", C.c_int), ("rows", C.c_int),
                ("cells", C.c_uint64), ("n", C.c_uint64)]


class SlabInfo(C.Structure):
    _fields_ = [("n_owned", C.c_uint64), ("own_first", C.c_uint64), ("n_window", C.c_uint64), ("row_lo", C.c_int), ("row_hi", C.c_int),
                ("own_row_begin", C.c_int), ("own_row_end", C.c_int), ("halo_rows", C.c_int), ("cols", C.c_int), ("rows", C.c_int),
                ("rounds", C.c_uint32), ("taint_rows", C.c_uint32), ("violations", C.c_uint32)]


# every symbol include/ppc_b200.h declares (tests check the library exports all of them)
REFERENCE_SYMBOLS = ["particlesCreate", "particlesFree", "particlesAddParticle", "particlesRemoveParticle", "particlesUpdate",
                     "particlesCollisionsCheck_Grid"]
EXTENSION_SYMBOLS = ["ppc_set_device", "ppc_set_sync_policy", "ppc_set_download_mask", "ppc_request_colors", "ppc_host_modified",
                     "ppc_download", "ppc_download_async", "ppc_download_wait", "ppc_synchronize", "ppc_run_substeps", "ppc_set_option", "ppc_debug_grid_info", "ppc_debug_keys",
                     "ppc_debug_sorted_order", "ppc_debug_cell_start", "ppc_debug_pairs", "ppc_debug_hot_pairs", "ppc_debug_gs_params", "ppc_gs_fallbacks", "ppc_gs_fallback_reasons", "ppc_debug_wave_stats", "ppc_stream", "ppc_kernel_launches",
                     "ppc_profile", "ppc_profile_stages", "ppc_profile_stage_name", "ppc_profile_read", "ppc_profile_reset",
                     "ppc_version"]
SLAB_SYMBOLS = ["ppc_grid_dims", "ppc_slab_gs_halo", "ppc_gs_auto_tile_width", "ppc_slab_init", "ppc_slab_begin", "ppc_slab_pack", "ppc_slab_interior", "ppc_slab_finish", "ppc_slab_download", "ppc_slab_download_async", "ppc_slab_info",
                "ppc_slab_debug_pairs"]
SNAPSHOT_SYMBOLS = ["ppc_snapshot_write", "ppc_snapshot_read"]

_lib = None


def build_library() -> None:
    subprocess.run(["make", "-s", "-C", os.path.join(HERE, "csrc")], check=True)


def load_library() -> C.CDLL:
    """Load the CUDA library.  Raises if it has not been built -- there is nothing to fall back to."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(the solver has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    PP = C.POINTER(Particles_t)
    L.particlesCreate.argtypes, L.particlesCreate.restype = [C.c_size_t], Particles_t
    L.particlesFree.argtypes, L.particlesFree.restype = [PP], None
    L.particlesAddParticle.argtypes, L.particlesAddParticle.restype = [PP, Vector2, C.c_float, C.c_float, C.c_uint8], None
    L.particlesRemoveParticle.argtypes, L.particlesRemoveParticle.restype = [PP], None
    L.particlesUpdate.argtypes, L.particlesUpdate.restype = [PP, C.c_float], None
    L.particlesCollisionsCheck_Grid.argtypes, L.particlesCollisionsCheck_Grid.restype = [PP, C.c_uint16, C.c_uint16, C.c_float], None
    L.ppc_set_device.argtypes, L.ppc_set_device.restype = [C.c_int], None
    L.ppc_set_sync_policy.argtypes, L.ppc_set_sync_policy.restype = [PP, C.c_int], None
    L.ppc_set_download_mask.argtypes, L.ppc_set_download_mask.restype = [PP, C.c_uint], None
    L.ppc_request_colors.argtypes, L.ppc_request_colors.restype = [PP], None
    L.ppc_host_modified.argtypes, L.ppc_host_modified.restype = [PP], None
    L.ppc_download.argtypes, L.ppc_download.restype = [PP, C.c_uint], None
    L.ppc_synchronize.argtypes, L.ppc_synchronize.restype = [PP], None
    L.ppc_download_async.argtypes, L.ppc_download_async.restype = [PP, C.c_uint], None
    L.ppc_download_wait.argtypes, L.ppc_download_wait.restype = [PP], None
    L.ppc_run_substeps.argtypes, L.ppc_run_substeps.restype = [PP, C.c_uint16, C.c_uint16, C.c_float, C.c_int, C.c_int], None
    L.ppc_set_option.argtypes, L.ppc_set_option.restype = [PP, C.c_char_p, C.c_longlong], C.c_int
    L.ppc_debug_grid_info.argtypes, L.ppc_debug_grid_info.restype = [PP, C.POINTER(GridInfo)], C.c_int
    L.ppc_debug_keys.argtypes, L.ppc_debug_keys.restype = [PP, u32p], C.c_int
    L.ppc_debug_sorted_order.argtypes, L.ppc_debug_sorted_order.restype = [PP, u32p], C.c_int
    L.ppc_debug_cell_start.argtypes, L.ppc_debug_cell_start.restype = [PP, u32p], C.c_int
    L.ppc_debug_pairs.argtypes, L.ppc_debug_pairs.restype = [PP, i32p, C.c_size_t], C.c_int64
    L.ppc_debug_hot_pairs.argtypes, L.ppc_debug_hot_pairs.restype = [PP, i32p, C.c_size_t], C.c_int64
    L.ppc_debug_gs_params.argtypes, L.ppc_debug_gs_params.restype = [PP, C.POINTER(C.c_int)], C.c_int
    L.ppc_gs_fallbacks.argtypes, L.ppc_gs_fallbacks.restype = [PP], C.c_ulonglong
    L.ppc_gs_fallback_reasons.argtypes, L.ppc_gs_fallback_reasons.restype = [PP, C.POINTER(C.c_ulonglong)], None
    L.ppc_debug_wave_stats.argtypes = [PP, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]
    L.ppc_debug_wave_stats.restype = C.c_int
    L.ppc_stream.argtypes, L.ppc_stream.restype = [PP], C.c_void_p
    L.ppc_kernel_launches.argtypes, L.ppc_kernel_launches.restype = [], C.c_ulonglong
    L.ppc_profile.argtypes, L.ppc_profile.restype = [PP, C.c_int], None
    L.ppc_profile_stages.argtypes, L.ppc_profile_stages.restype = [], C.c_int
    L.ppc_profile_stage_name.argtypes, L.ppc_profile_stage_name.restype = [C.c_int], C.c_char_p
    L.ppc_profile_read.argtypes, L.ppc_profile_read.restype = [PP, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_ulonglong)], C.c_int
    L.ppc_profile_reset.argtypes, L.ppc_profile_reset.restype = [PP], None
    L.ppc_version.argtypes, L.ppc_version.restype = [], C.c_char_p
    L.ppc_snapshot_write.argtypes, L.ppc_snapshot_write.restype = [PP, C.c_uint16, C.c_uint16, C.c_char_p], C.c_int
    L.ppc_snapshot_read.argtypes = [PP, C.c_char_p, C.POINTER(C.c_uint16), C.POINTER(C.c_uint16)]
    L.ppc_snapshot_read.restype = C.c_int64
    L.ppc_grid_dims.argtypes, L.ppc_grid_dims.restype = [C.c_uint16, C.c_uint16, C.c_float, C.POINTER(GridInfo)], C.c_int
    L.ppc_slab_gs_halo.argtypes, L.ppc_slab_gs_halo.restype = [C.c_int, C.c_int, C.c_int], C.c_int
    L.ppc_gs_auto_tile_width.argtypes, L.ppc_gs_auto_tile_width.restype = [C.c_double, C.c_double], C.c_int
    L.ppc_slab_init.argtypes = [PP, u32p, C.c_uint16, C.c_uint16, C.c_float, C.c_int, C.c_int, C.c_int]
    L.ppc_slab_init.restype = C.c_int
    L.ppc_slab_begin.argtypes, L.ppc_slab_begin.restype = [PP, C.c_float, C.c_int, C.POINTER(C.c_uint64)], C.c_int
    L.ppc_slab_pack.argtypes, L.ppc_slab_pack.restype = [PP, C.c_void_p, C.c_void_p, C.c_int], C.c_int
    L.ppc_slab_interior.argtypes, L.ppc_slab_interior.restype = [PP], C.c_int
    L.ppc_slab_finish.argtypes = [PP, C.c_float, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_int, C.POINTER(C.c_uint64)]
    L.ppc_slab_finish.restype = C.c_int64
    L.ppc_slab_download.argtypes, L.ppc_slab_download.restype = [PP, u32p, C.c_uint], C.c_int
    L.ppc_slab_download_async.argtypes, L.ppc_slab_download_async.restype = [PP, u32p, C.c_uint], C.c_int
    L.ppc_slab_info.argtypes, L.ppc_slab_info.restype = [PP, C.POINTER(SlabInfo)], C.c_int
    L.ppc_slab_debug_pairs.argtypes, L.ppc_slab_debug_pairs.restype = [PP, i32p, C.c_size_t], C.c_int64
    _lib = L
    return L


class Particles:
    """One ``struct Particles_t`` created by the library, with numpy views of its seven host arrays.

    Methods carry the reference's function names without the ``particles`` prefix:
    ``Update`` = particlesUpdate, ``CollisionsCheck_Grid`` = particlesCollisionsCheck_Grid, ``AddParticle``,
    ``RemoveParticle``, ``Free``.
    """

    def __init__(self, num_particles_init: int, device: int | None = None):
        self.lib = load_library()
        if device is not None:
            self.lib.ppc_set_device(device)
        self.s = self.lib.particlesCreate(num_particles_init)
        self._freed = False

    # -- views of the host arrays (re-made on every access: the arrays move when the store grows) ----------
    def _view(self, ptr, dtype, count, width=1):
        if count == 0:
            return np.zeros((0, width) if width > 1 else 0, dtype)
        arr = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(count * width * np.dtype(dtype).itemsize,))
        arr = arr.view(dtype)
        return arr.reshape(count, width) if width > 1 else arr

    def array(self, name: str, full: bool = False) -> np.ndarray:
        """numpy view of host array ``name`` over [0, length) (or the whole capacity with full=True)."""
        count = self.s.size if full else self.s.length
        if name == "colors":
            return self._view(self.s.colors, np.uint8, count, 4)
        return self._view(getattr(self.s, name), np.float32, count)

    @property
    def length(self) -> int:
        return int(self.s.length)

    @property
    def size(self) -> int:
        return int(self.s.size)

    def fill(self, px, py, vx, vy, mass, radius, rgba=None) -> None:
        """Write a whole field straight into the arrays with length = N (the harness path of SURVEY.md section 8d,
        bypassing particlesAddParticle) and tell the library the host arrays changed."""
        n = len(px)
        if n > self.s.size:
            raise ValueError("field larger than the store")
        self.s.length = n
        for name, src in (("p_x", px), ("p_y", py), ("v_x", vx), ("v_y", vy), ("mass", mass), ("radius", radius)):
            self.array(name)[:] = src
        if rgba is not None:
            self.array("colors")[:] = rgba
        self.lib.ppc_host_modified(C.byref(self.s))

    # -- reference entry points ---------------------------------------------------------------------------
    def AddParticle(self, x: float, y: float, mass: float, radius: float, count: int = 1) -> None:
        self.lib.particlesAddParticle(C.byref(self.s), Vector2(x, y), mass, radius, count)

    def RemoveParticle(self) -> None:
        self.lib.particlesRemoveParticle(C.byref(self.s))

    def Update(self, time_step: float) -> None:
        self.lib.particlesUpdate(C.byref(self.s), time_step)

    def CollisionsCheck_Grid(self, window_width: int, window_height: int, time_step: float) -> None:
        self.lib.particlesCollisionsCheck_Grid(C.byref(self.s), window_width, window_height, time_step)

    def Free(self) -> None:
        if not self._freed:
            self.lib.particlesFree(C.byref(self.s))
            self._freed = True

    # -- extensions ---------------------------------------------------------------------------------------
    def set_sync_policy(self, policy: int) -> None:
        self.lib.ppc_set_sync_policy(C.byref(self.s), policy)

    def set_download_mask(self, mask: int) -> None:
        self.lib.ppc_set_download_mask(C.byref(self.s), mask)

    def request_colors(self) -> None:
        self.lib.ppc_request_colors(C.byref(self.s))

    def host_modified(self) -> None:
        self.lib.ppc_host_modified(C.byref(self.s))

    def download(self, mask: int = PPC_DL_ALL) -> None:
        self.lib.ppc_download(C.byref(self.s), mask)

    def download_async(self, mask: int = PPC_DL_ALL) -> None:
        """Start copying the current state to the host arrays; they are valid after download_wait()."""
        self.lib.ppc_download_async(C.byref(self.s), mask)

    def download_wait(self) -> None:
        self.lib.ppc_download_wait(C.byref(self.s))

    def synchronize(self) -> None:
        self.lib.ppc_synchronize(C.byref(self.s))

    def run_substeps(self, W: int, H: int, dt: float, n: int, color_last: bool = False) -> None:
        self.lib.ppc_run_substeps(C.byref(self.s), W, H, dt, n, int(color_last))

    def set_option(self, name: str, value: int) -> None:
        if self.lib.ppc_set_option(C.byref(self.s), name.encode(), value) != 0:
            raise KeyError(name)

    def grid_info(self) -> GridInfo:
        g = GridInfo()
        if self.lib.ppc_debug_grid_info(C.byref(self.s), C.byref(g)) != 0:
            raise RuntimeError("no grid has been built yet")
        return g

    def debug_grid(self) -> dict:
        """Keys by API index, canonical sorted order and cell-start table of the last grid build."""
        g = self.grid_info()
        keys = np.zeros(g.n, np.uint32)
        order = np.zeros(g.n, np.uint32)
        start = np.zeros(g.cells + 1, np.uint32)
        assert self.lib.ppc_debug_keys(C.byref(self.s), keys.ctypes.data_as(u32p)) == 0
        assert self.lib.ppc_debug_sorted_order(C.byref(self.s), order.ctypes.data_as(u32p)) == 0
        assert self.lib.ppc_debug_cell_start(C.byref(self.s), start.ctypes.data_as(u32p)) == 0
        return dict(cell_size=g.cell_size, max_radius=g.max_radius, cols=g.cols, rows=g.rows, keys=keys, order=order, cell_start=start)

    def debug_pairs(self) -> np.ndarray:
        n = self.lib.ppc_debug_pairs(C.byref(self.s), None, 0)
        if n < 0:
            raise RuntimeError("pair list unavailable: no collisions check since the last change")
        out = np.zeros((n, 2), np.int32)
        if n:
            self.lib.ppc_debug_pairs(C.byref(self.s), out.ctypes.data_as(i32p), n)
        return out

    def debug_hot_pairs(self) -> np.ndarray:
        """The pair set as the hot narrow-phase kernel of the last substep found it, sorted by (i, j)."""
        n = self.lib.ppc_debug_hot_pairs(C.byref(self.s), None, 0)
        if n < 0:
            raise RuntimeError(f"hot pair set unavailable ({n})")
        out = np.zeros((n, 2), np.int32)
        if n:
            self.lib.ppc_debug_hot_pairs(C.byref(self.s), out.ctypes.data_as(i32p), n)
        return out[np.lexsort((out[:, 1], out[:, 0]))]

    def gs_params(self) -> dict:
        """Tile geometry of the last "resolve"=3 substep (the parameters of the oracle's model of its pair order)."""
        out = (C.c_int * 4)()
        if self.lib.ppc_debug_gs_params(C.byref(self.s), out) != 0:
            raise RuntimeError("no tile Gauss-Seidel substep has run")
        return dict(wc=out[0], tile_r=out[1], strip_limit=out[2], cap=out[3])

    def gs_fallbacks(self) -> int:
        """Substeps on which "resolve"=3 ran the exact resolve instead because a sub-tile overflowed the tile kernel."""
        return int(self.lib.ppc_gs_fallbacks(C.byref(self.s)))

    def gs_fallback_reasons(self) -> dict:
        """... by cause (a substep may have several)."""
        out = (C.c_ulonglong * 4)()
        self.lib.ppc_gs_fallback_reasons(C.byref(self.s), out)
        return dict(strip_over_capacity=int(out[0]), contact_pool_full=int(out[1]), cell_or_colour_limit=int(out[2]), record_buffer_full=int(out[3]))

    def wave_stats(self):
        """(rounds, particle visits) of the last sequential-equivalent resolve."""
        r, v = C.c_ulonglong(), C.c_ulonglong()
        if self.lib.ppc_debug_wave_stats(C.byref(self.s), C.byref(r), C.byref(v)) != 0:
            return None
        return int(r.value), int(v.value)

    def snapshot_write(self, W: int, H: int, path: str) -> None:
        rc = self.lib.ppc_snapshot_write(C.byref(self.s), W, H, path.encode())
        if rc != 0:
            raise OSError(f"ppc_snapshot_write({path}) failed ({rc})")

    def snapshot_read(self, path: str):
        """Load a snapshot into this store (capacity permitting); returns (n, W, H)."""
        w, h = C.c_uint16(), C.c_uint16()
        n = self.lib.ppc_snapshot_read(C.byref(self.s), path.encode(), C.byref(w), C.byref(h))
        if n < 0:
            raise OSError(f"ppc_snapshot_read({path}) failed ({n})")
        return int(n), int(w.value), int(h.value)

    def stream(self) -> int:
        return int(self.lib.ppc_stream(C.byref(self.s)) or 0)

    def profile(self, enable: bool) -> None:
        self.lib.ppc_profile(C.byref(self.s), int(enable))

    def profile_reset(self) -> None:
        self.lib.ppc_profile_reset(C.byref(self.s))

    def profile_read(self) -> dict:
        out = {}
        for st in range(self.lib.ppc_profile_stages()):
            ms, cnt = C.c_double(), C.c_ulonglong()
            self.lib.ppc_profile_read(C.byref(self.s), st, C.byref(ms), C.byref(cnt))
            out[self.lib.ppc_profile_stage_name(st).decode()] = (ms.value, int(cnt.value))
        return out


def kernel_launches() -> int:
    return int(load_library().ppc_kernel_launches())
