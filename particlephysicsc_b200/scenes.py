"""Synthetic particle fields for the benchmark configurations.

Host-side numpy only.  Every field is generated once on the host from a counter-based splitmix64 stream
(seed 0x5EED0000 + config number; float = top 24 bits / 2**24) and written straight into the seven SoA
arrays of ``struct Particles_t`` (reference includes/structs.h:9-22) with ``length = size = N`` -- i.e.
bypassing particlesAddParticle, so there is no dormant slot 0 (reference src/particles_arrays.c:266-269).
Masses are uniform integers 5..10 as in the reference's debug scene (src/main.c:82).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

_GOLDEN = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def splitmix64(seed: int, start: int, count: int) -> np.ndarray:
    """Outputs start .. start+count-1 of the splitmix64 stream seeded with ``seed`` (vectorised)."""
    with np.errstate(over="ignore"):
        k = np.arange(start + 1, start + count + 1, dtype=np.uint64)
        z = np.uint64(seed) + k * _GOLDEN
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        return z ^ (z >> np.uint64(31))


class Stream:
    """Sequential consumer of one splitmix64 stream."""

    def __init__(self, seed: int):
        self.seed, self.pos = seed, 0

    def u64(self, n: int) -> np.ndarray:
        out = splitmix64(self.seed, self.pos, n)
        self.pos += n
        return out

    def unit(self, n: int) -> np.ndarray:
        """float32 in [0, 1): top 24 bits / 2**24."""
        return ((self.u64(n) >> np.uint64(40)).astype(np.float32)) * np.float32(1.0 / 16777216.0)

    def ints(self, n: int, lo: int, hi: int) -> np.ndarray:
        """uniform integers in [lo, hi] inclusive (like Raylib's GetRandomValue, src/main.c:81-83)."""
        return (self.u64(n) % np.uint64(hi - lo + 1)).astype(np.int64) + lo


@dataclass
class Scene:
    """A particle field plus the box it lives in.  Arrays are float32 / (n,4) uint8, C-contiguous."""

    name: str
    W: int
    H: int
    px: np.ndarray
    py: np.ndarray
    vx: np.ndarray
    vy: np.ndarray
    mass: np.ndarray
    radius: np.ndarray
    rgba: np.ndarray

    @property
    def n(self) -> int:
        return int(self.px.shape[0])

    def area_fraction(self) -> float:
        return float(np.sum(np.pi * self.radius.astype(np.float64) ** 2) / (self.W * self.H))


def _finish(name, W, H, px, py, radius, rng: Stream, vmax: float) -> Scene:
    n = px.shape[0]
    mass = rng.ints(n, 5, 10).astype(np.float32)
    if vmax > 0:
        vx = (rng.unit(n) * np.float32(2) - np.float32(1)) * np.float32(vmax)
        vy = (rng.unit(n) * np.float32(2) - np.float32(1)) * np.float32(vmax)
    else:
        vx = np.zeros(n, np.float32)
        vy = np.zeros(n, np.float32)
    rgba = np.zeros((n, 4), np.uint8)
    rgba[:] = (0, 121, 241, 255)  # BLUE, as particlesAddParticle leaves it (src/particles_arrays.c:272)
    c = np.ascontiguousarray
    return Scene(name, int(W), int(H), c(px, np.float32), c(py, np.float32), vx, vy, mass, c(radius, np.float32), rgba)


def box_for(n: int, mean_disc_area: float, phi: float, aspect: float = 1.0):
    """Box (W, H) whose area holds n discs of the given mean area at area fraction phi."""
    area = n * mean_disc_area / phi
    W = math.sqrt(area * aspect)
    H = area / W
    W, H = int(round(W)), int(round(H))
    if W > 65535 or H > 65535:
        raise ValueError(f"box {W}x{H} exceeds the uint16_t window limit of the reference API")
    return W, H


def uniform_random(n: int, radius: float = 2.0, phi: float = 0.30, seed: int = 0x5EED0002, vmax: float = 0.0,
                   name: str = "uniform") -> Scene:
    """Config 2 / 4 throughput field: equal radii, uniform random centres in [r, W-r] x [r, H-r].

    Overlaps are present (it is a Poisson field); the solver relaxes them during the warm-up substeps.
    """
    W, H = box_for(n, math.pi * radius * radius, phi)
    rng = Stream(seed)
    r = np.float32(radius)
    px = r + rng.unit(n) * np.float32(W - 2 * radius)
    py = r + rng.unit(n) * np.float32(H - 2 * radius)
    return _finish(name, W, H, px, py, np.full(n, r, np.float32), rng, vmax)


def uniform_slab(n: int, radius: float, W: int, H: int, y0: float, y1: float, seed: int, vmax: float = 0.0, name: str = "uniform_slab") -> Scene:
    """One rank's share of the config-4 field (256M particles over 8 slabs, "generated per slab"): n equal-radius particles,
    uniform random in [r, W-r] x ([y0, y1) clipped to [r, H-r]) of a W x H box."""
    rng = Stream(seed)
    r = np.float32(radius)
    lo, hi = max(float(y0), radius), min(float(y1), H - radius)
    px = r + rng.unit(n) * np.float32(W - 2 * radius)
    py = np.float32(lo) + rng.unit(n) * np.float32(hi - lo)
    py = np.minimum(py, np.nextafter(np.float32(hi), np.float32(0)))   # float rounding must not push a particle into the next slab
    return _finish(name, W, H, px, py, np.full(n, r, np.float32), rng, vmax)


def jittered_lattice(n: int, radius: float = 2.0, phi: float = 0.30, seed: int = 0x5EED0102, vmax: float = 20.0,
                     name: str = "lattice") -> Scene:
    """Non-overlapping field for the statistical comparison: a square lattice with
    random jitter small enough that no two discs touch, particle index order shuffled."""
    W, H = box_for(n, math.pi * radius * radius, phi)
    nx = int(math.ceil(math.sqrt(n * W / H)))
    ny = int(math.ceil(n / nx))
    pitch_x, pitch_y = (W - 2 * radius) / nx, (H - 2 * radius) / ny
    slack = min(pitch_x, pitch_y) - 2 * radius
    if slack <= 0:
        raise ValueError("area fraction too high for a square lattice")
    rng = Stream(seed)
    idx = np.arange(nx * ny, dtype=np.int64)
    # Fisher-Yates driven by the stream would be O(n) python; a keyed argsort is equivalent and vectorised
    perm = np.argsort(rng.u64(nx * ny), kind="stable")
    idx = idx[perm][:n]
    jx = (rng.unit(n) - np.float32(0.5)) * np.float32(0.9 * slack)
    jy = (rng.unit(n) - np.float32(0.5)) * np.float32(0.9 * slack)
    px = (radius + (idx % nx + 0.5) * pitch_x).astype(np.float32) + jx
    py = (radius + (idx // nx + 0.5) * pitch_y).astype(np.float32) + jy
    return _finish(name, W, H, px, py, np.full(n, np.float32(radius), np.float32), rng, vmax)


def polydisperse_pile(n: int, rmin: float = 1.5, rmax: float = 6.0, W: int = 31500, H: int = 40000, seed: int = 0x5EED0003,
                      classes: int = 64, compress: float = 0.99, vmax: float = 0.0, name: str = "pile") -> Scene:
    """Config 3 field: radii uniform in [rmin, rmax] (ratio 1:4) packed bottom-up (y grows downward, reference
    includes/constants.h:15) as a dense pile resting on the bottom wall.  Particles are grouped into narrow size
    classes, largest lowest; each class fills jittered hex-lattice rows whose pitch is ``compress`` x the class's
    mean diameter, so neighbours touch or overlap by about a percent -- the contact network of a pile under its own
    weight (area fraction ~0.8).  The solver then relaxes it during the settle/warm-up substeps; the relaxed state
    is what gets benchmarked.

    For n other than 2**24 the default box is scaled at constant aspect so the pile fills the same share of it.
    """
    rng = Stream(seed)
    radius = (np.float32(rmin) + rng.unit(n) * np.float32(rmax - rmin)).astype(np.float32)
    classes = max(4, min(classes, n // 4096))   # small fields: fewer, wider classes so rows are not mostly empty
    scale = math.sqrt(n / float(1 << 24))
    W, H = max(64, int(round(W * scale))), max(64, int(round(H * scale)))
    order = np.argsort(radius, kind="stable")[::-1]  # big ones at the bottom
    edges = np.linspace(0, n, classes + 1).astype(np.int64)
    px = np.zeros(n, np.float32)
    py = np.zeros(n, np.float32)
    y_floor = float(H)
    for c in range(classes):
        members = order[edges[c]:edges[c + 1]]
        if members.size == 0:
            continue
        rc, rmean = float(radius[members].max()), float(radius[members].mean())
        pitch = 2.0 * rmean * compress
        per_row = max(1, int((W - 2 * rc) // pitch))
        rows = int(math.ceil(members.size / per_row))
        row_h = pitch * math.sqrt(3.0) / 2.0
        k = np.arange(members.size)
        row, col = k // per_row, k % per_row
        jit = np.float32(0.05 * rmean)
        x = rc + (col + 0.5 * (row % 2)) * pitch
        x = np.minimum(x, W - rc)
        y = y_floor - rc - row * row_h
        px[members] = x.astype(np.float32) + (rng.unit(members.size) - np.float32(0.5)) * jit
        py[members] = y.astype(np.float32) + (rng.unit(members.size) - np.float32(0.5)) * jit
        y_floor -= (rows - 1) * row_h + 2.0 * rc * compress
    if y_floor < 0:
        raise ValueError("pile does not fit the box")
    return _finish(name, W, H, px, py, radius, rng, vmax)


def debug_scene_frames(frames: int = 4500, per_frame: int = 10, W: int = 1200, H: int = 675, seed: int = 0x5EED0001):
    """Config 1, the reference's own debug scene (src/main.c:36-38,80-85): each frame one random position,
    mass 5..10 and radius 2..8 is drawn and added ``per_frame`` times through particlesAddParticle.
    Yields (x, y, mass, radius, count) per frame; GetRandomValue is replaced by the seeded stream."""
    rng = Stream(seed)
    for _ in range(frames):
        x = int(rng.ints(1, 0, W)[0])
        y = int(rng.ints(1, 0, H)[0])
        m = int(rng.ints(1, 5, 10)[0])
        r = int(rng.ints(1, 2, 8)[0])
        yield float(x), float(y), float(m), float(r), per_frame


# ---- snapshot files (format of ppc_snapshot_write / ppc_snapshot_read, include/ppc_b200.h part 4) ---------------------
SNAPSHOT_MAGIC = 0x3130303242435050  # "PPCB2001"


def write_snapshot(path: str, sc: Scene) -> None:
    """Write a scene in the library's snapshot format (host only; the library reads it with ppc_snapshot_read)."""
    with open(path, "wb") as f:
        np.array([SNAPSHOT_MAGIC, 1, sc.n, sc.W, sc.H, 0, 0, 0], dtype="<u8").tofile(f)
        for a in (sc.px, sc.py, sc.vx, sc.vy, sc.mass, sc.radius):
            np.ascontiguousarray(a, dtype="<f4").tofile(f)
        np.ascontiguousarray(sc.rgba, dtype=np.uint8).tofile(f)


def read_snapshot(path: str, name: str = "snapshot") -> Scene:
    with open(path, "rb") as f:
        h = np.fromfile(f, dtype="<u8", count=8)
        if len(h) != 8 or int(h[0]) != SNAPSHOT_MAGIC or int(h[1]) != 1:
            raise ValueError(f"{path} is not a particle snapshot")
        n = int(h[2])
        arrays = [np.fromfile(f, dtype="<f4", count=n) for _ in range(6)]
        rgba = np.fromfile(f, dtype=np.uint8, count=4 * n).reshape(n, 4)
        if any(len(a) != n for a in arrays) or rgba.shape[0] != n:
            raise ValueError(f"{path} is truncated")
    return Scene(name, int(h[3]), int(h[4]), *arrays, rgba)
