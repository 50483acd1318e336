"""Grid and field containers: the minimum of Oceananigans' `RectilinearGrid` / `Field` that the
biogeochemical hot path reads — sizes, halos, z nodes, and halo'd column-major parent arrays.

Memory layout is exactly the parent array of an Oceananigans field (x fastest, halos on every
non-Flat side) so that the pointers handed to the C ABI here are interchangeable with the
`CuArray` parents the Julia glue passes.  torch is used for device memory only.
"""
from __future__ import annotations

import contextlib
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib

FLAT = "Flat"


class RectilinearGrid:
    """`RectilinearGrid(size=(Nx, Ny, Nz), extent=(Lx, Ly, Lz))` with z ∈ [-Lz, 0], or explicit
    `x=(x0, x1)`, `y=…`, `z=(z0, z1) | array of Nz+1 faces | callable k ↦ z_face(k), k = 1…Nz+1`.
    `topology` entries: "Periodic", "Bounded" or "Flat" (Flat ⇒ N = 1, H = 0).  Default halo 3."""

    def __init__(self, size, extent=None, x=None, y=None, z=None,
                 topology=("Periodic", "Periodic", "Bounded"), halo=None, device="cuda"):
        topology = tuple(topology)
        size = tuple(size) if np.ndim(size) else (int(size),)
        nonflat = [t != FLAT for t in topology]
        if len(size) != sum(nonflat):
            raise ValueError(f"size {size} does not match the {sum(nonflat)} non-Flat dimensions of {topology}")
        it = iter(size)
        N = [next(it) if nf else 1 for nf in nonflat]
        if halo is None:
            halo = tuple(3 for nf in nonflat if nf)
        halo = tuple(halo) if np.ndim(halo) else (int(halo),)
        ih = iter(halo)
        H = [next(ih) if nf else 0 for nf in nonflat]
        self.Nx, self.Ny, self.Nz = (int(n) for n in N)
        self.Hx, self.Hy, self.Hz = (int(h) for h in H)
        self.topology = topology
        self.device = torch.device(device)

        if extent is not None:
            ie = iter(extent if np.ndim(extent) else (extent,))
            L = [next(ie) if nf else None for nf in nonflat]
            x = x if x is not None else ((0.0, L[0]) if L[0] is not None else None)
            y = y if y is not None else ((0.0, L[1]) if L[1] is not None else None)
            z = z if z is not None else ((-L[2], 0.0) if L[2] is not None else None)
        self.x, self.y = x, y
        self.Lx = (x[1] - x[0]) if x is not None else 1.0
        self.Ly = (y[1] - y[0]) if y is not None else 1.0
        self.dx = self.Lx / self.Nx
        self.dy = self.Ly / self.Ny

        # z faces of the interior (Nz+1) …
        if z is None:
            zf = np.array([-1.0, 0.0]) if self.Nz == 1 else None
            if zf is None:
                raise ValueError("z (or extent) is required for a non-Flat vertical dimension")
        elif callable(z):
            zf = np.array([float(z(k)) for k in range(1, self.Nz + 2)])
        else:
            z = np.asarray(z, dtype=np.float64)
            if z.size == 2 and self.Nz != 1:
                zf = z[0] + (z[1] - z[0]) / self.Nz * np.arange(self.Nz + 1)
                zf[-1] = z[1]
            else:
                zf = z.copy()
        if zf.size != self.Nz + 1:
            raise ValueError(f"need {self.Nz + 1} z faces, got {zf.size}")
        if not np.all(np.diff(zf) > 0):  # Oceananigans: "z faces must be strictly increasing" (bottom first)
            raise ValueError("z faces must be strictly increasing, from the bottom face to the surface")
        # … extended into the halos with the end-cell spacing, like Oceananigans does
        Hz = self.Hz
        lo = zf[0] - (zf[1] - zf[0]) * np.arange(Hz, 0, -1)
        hi = zf[-1] + (zf[-1] - zf[-2]) * np.arange(1, Hz + 1)
        zf_parent = np.concatenate([lo, zf, hi])
        zc_parent = 0.5 * (zf_parent[:-1] + zf_parent[1:])
        self.zf_host = zf_parent  # Nz + 1 + 2Hz
        self.zc_host = zc_parent  # Nz + 2Hz
        self.zf_dev = torch.from_numpy(zf_parent).to(self.device)
        self.zc_dev = torch.from_numpy(zc_parent).to(self.device)
        self.Lz = float(zf[-1] - zf[0])

    # interior node / spacing accessors (0-based k)
    @property
    def zc(self):
        return self.zc_host[self.Hz:self.Hz + self.Nz]

    @property
    def zf(self):
        return self.zf_host[self.Hz:self.Hz + self.Nz + 1]

    @property
    def dz(self):
        return np.diff(self.zf)

    @property
    def parent_shape(self):
        """torch (row-major) shape of a 3-D parent array == Julia column-major (x, y, z)."""
        return (self.Nz + 2 * self.Hz, self.Ny + 2 * self.Hy, self.Nx + 2 * self.Hx)

    @property
    def plane_shape(self):
        return (1, self.Ny + 2 * self.Hy, self.Nx + 2 * self.Hx)

    @property
    def ncells(self):
        return self.Nx * self.Ny * self.Nz

    def interior(self, parent: torch.Tensor) -> torch.Tensor:
        if parent.shape[0] == 1:  # 2-D field
            return parent[:, self.Hy:self.Hy + self.Ny, self.Hx:self.Hx + self.Nx]
        return parent[self.Hz:self.Hz + self.Nz, self.Hy:self.Hy + self.Ny, self.Hx:self.Hx + self.Nx]

    def cell_volume(self) -> torch.Tensor:
        """Interior cell volumes, shape (Nz, 1, 1) broadcastable against `interior(...)`."""
        return torch.from_numpy(self.dz * self.dx * self.dy).to(self.device).reshape(-1, 1, 1)

    def c_grid(self, i0=None, i1=None, j0=None, j1=None) -> _lib.obm_grid:
        r = getattr(self, "_subrange", (0, 0, 0, 0))
        i0, i1, j0, j1 = (r[n] if v is None else v for n, v in enumerate((i0, i1, j0, j1)))
        bottom = getattr(self, "bottom_indices", None)
        return _lib.obm_grid(self.Nx, self.Ny, self.Nz, self.Hx, self.Hy, self.Hz, i0, i1, j0, j1,
                             self.zc_dev.data_ptr(), self.zf_dev.data_ptr(), bottom.data_ptr() if bottom is not None else None)

    def immersed(self, bottom_height: "Field") -> "RectilinearGrid":
        """`ImmersedBoundaryGrid(grid, GridFittedBottom(bottom_height))` restricted to what the path reads: the index of
        the bottom-most active cell of every column (`calculate_bottom_indices`, src/Sediments/bottom_indices.jl:19-26).
        Kernels the reference guards with `!immersed_cell(i, j, k, grid)` — ScaleNegativeTracers — then leave the cells
        below it untouched.  Returns a grid sharing everything else with this one."""
        from .sediments import calculate_bottom_indices
        g = type(self).__new__(type(self))
        g.__dict__.update(self.__dict__)
        g.bottom_indices = calculate_bottom_indices(self, bottom_height)
        return g

    @contextlib.contextmanager
    def restrict(self, j0: int, j1: int, i0: int = 0, i1: int = 0):
        """Every kernel launched inside the block processes only the interior sub-range
        i ∈ [i0, i1), j ∈ [j0, j1) (partial launches: slab pipelining, multi-stream execution)."""
        old = getattr(self, "_subrange", (0, 0, 0, 0))
        self._subrange = (i0, i1, j0, j1)
        try:
            yield self
        finally:
            self._subrange = old

    def slab(self, rank: int, world: int) -> "RectilinearGrid":
        """x–y slab decomposition for multi-GPU runs: split y (the slower horizontal axis, so
        x rows stay contiguous) into `world` equal ranges; every hot kernel is pointwise or
        column-local so slabs need no halo exchange."""
        if self.Ny % world:
            raise ValueError(f"Ny = {self.Ny} is not divisible by {world} ranks")
        ny = self.Ny // world
        y0 = (self.y[0] if self.y is not None else 0.0) + rank * ny * self.dy
        g = type(self).__new__(type(self))
        g.__dict__.update(self.__dict__)
        g.Ny = ny
        g.y = (y0, y0 + ny * self.dy)
        g.Ly = ny * self.dy
        return g

    def __repr__(self):
        return (f"RectilinearGrid(size=({self.Nx}, {self.Ny}, {self.Nz}), halo=({self.Hx}, {self.Hy}, {self.Hz}), "
                f"topology={self.topology}, device={self.device})")


class LatitudeLongitudeGrid(RectilinearGrid):
    """`LatitudeLongitudeGrid(size = (Nλ, Nφ, Nz), longitude = (λ₀, λ₁), latitude = (φ₀, φ₁), z = …)` in degrees — the
    second grid of the reference's light tests (test/test_light.jl:113-114).  Every hot kernel is pointwise or local to
    one column, so on the device this is the same arrays and the same launches as a `RectilinearGrid` of that size;
    what the sphere changes is host-side: cell areas (tracer inventories) and the latitude of each row of columns."""

    def __init__(self, size, longitude, latitude, z, topology=("Periodic", "Bounded", "Bounded"), halo=None,
                 radius: float = 6371e3, device="cuda"):
        if not (-90.0 <= latitude[0] < latitude[1] <= 90.0):
            raise ValueError("latitude must satisfy −90 ≤ φ₀ < φ₁ ≤ 90")
        super().__init__(size, x=tuple(longitude), y=tuple(latitude), z=z, topology=topology, halo=halo, device=device)
        self.radius = float(radius)

    @property
    def latitude_faces(self):
        return self.y[0] + self.dy * np.arange(self.Ny + 1)

    @property
    def latitude_centers(self):
        return self.y[0] + self.dy * (np.arange(self.Ny) + 0.5)

    def cell_area(self) -> np.ndarray:
        """Area of the cells of each latitude row, R² Δλ (sin φⱼ₊₁ − sin φⱼ), shape (Ny,)."""
        phi = np.radians(self.latitude_faces)
        return self.radius ** 2 * np.radians(self.dx) * np.diff(np.sin(phi))

    def cell_volume(self) -> torch.Tensor:
        v = self.dz.reshape(-1, 1, 1) * self.cell_area().reshape(1, -1, 1)
        return torch.from_numpy(v).to(self.device)

    def __repr__(self):
        return (f"LatitudeLongitudeGrid(size=({self.Nx}, {self.Ny}, {self.Nz}), longitude={self.x}, latitude={self.y}, "
                f"halo=({self.Hx}, {self.Hy}, {self.Hz}), device={self.device})")


@dataclass
class Field:
    """A halo'd field: `data` is the parent array (torch, float64, on the grid's device)."""
    grid: RectilinearGrid
    data: torch.Tensor
    name: str = ""
    constant: bool = False  # Oceananigans' ConstantField: prescribed by the user, never recomputed by a hook

    @property
    def ptr(self) -> int:
        return self.data.data_ptr()

    @property
    def interior(self) -> torch.Tensor:
        return self.grid.interior(self.data)

    @property
    def is_2d(self) -> bool:
        return self.data.shape[0] == 1

    @property
    def face_interior(self) -> torch.Tensor:
        """Interior of a z-FACE field: Nz + 1 faces (face k, 0-based, at parent level k + Hz)."""
        g = self.grid
        return self.data[g.Hz:g.Hz + g.Nz + 1, g.Hy:g.Hy + g.Ny, g.Hx:g.Hx + g.Nx]

    def set(self, value):
        """`set!(field, value)`: scalar, array of interior shape (Nz, Ny, Nx) or (Ny, Nx), torch or numpy."""
        if not torch.is_tensor(value):
            value = torch.as_tensor(np.asarray(value, dtype=np.float64))
        self.interior.copy_(value.to(self.data.device, torch.float64).expand_as(self.interior))
        return self

    def fill_halos_zero_gradient(self):
        g = self.grid
        d = self.data
        if g.Hx:
            d[..., :g.Hx] = d[..., g.Hx:g.Hx + 1]
            d[..., g.Hx + g.Nx:] = d[..., g.Hx + g.Nx - 1:g.Hx + g.Nx]
        if g.Hy:
            d[:, :g.Hy] = d[:, g.Hy:g.Hy + 1]
            d[:, g.Hy + g.Ny:] = d[:, g.Hy + g.Ny - 1:g.Hy + g.Ny]
        if g.Hz and not self.is_2d:
            d[:g.Hz] = d[g.Hz:g.Hz + 1]
            d[g.Hz + g.Nz:] = d[g.Hz + g.Nz - 1:g.Hz + g.Nz]
        return self


def fill_z_halos(field: Field, j0: int = 0, j1: int = 0):
    """Zero-gradient z-halos of a 3-D field (what Oceananigans' `fill_halo_regions!` leaves for a tracer with the default
    no-flux boundary conditions), optionally for the interior rows j ∈ [j0, j1) only.  The sinking operators read the cell
    below the bottom one (sinking.cu `face_value`, sediments.cu `sinking_flux`), so the host mirror refreshes these planes
    whenever the tracer has changed — Oceananigans does it in `update_state!` before the hooks run."""
    g, d = field.grid, field.data
    if not g.Hz or field.is_2d:
        return field
    rows = slice(g.Hy + j0, g.Hy + (j1 or g.Ny)) if (j0 or j1) else slice(None)
    d[:g.Hz, rows] = d[g.Hz:g.Hz + 1, rows]
    d[g.Hz + g.Nz:, rows] = d[g.Hz + g.Nz - 1:g.Hz + g.Nz, rows]
    return field


def CenterField(grid: RectilinearGrid, name: str = "", fill: float = 0.0) -> Field:
    return Field(grid, torch.full(grid.parent_shape, fill, dtype=torch.float64, device=grid.device), name)


def ZFaceField(grid: RectilinearGrid, name: str = "", fill: float = 0.0) -> Field:
    """`ZFaceField(grid)` — (Center, Center, Face): parent has Nz + 1 + 2Hz levels."""
    shape = (grid.Nz + 1 + 2 * grid.Hz, grid.Ny + 2 * grid.Hy, grid.Nx + 2 * grid.Hx)
    return Field(grid, torch.full(shape, fill, dtype=torch.float64, device=grid.device), name)


def ConstantField(grid: RectilinearGrid, value: float, name: str = "") -> Field:
    """`ConstantField(value)` where the reference prescribes a column quantity (zₘₓₗ, zₑᵤ, κ̄, PAR̄ₘₓₗ of a box model,
    test_PISCES.jl:44-48): an x–y field holding `value` that the state update leaves alone, like the reference's
    `compute_euphotic_depth!(::ConstantField, args...) = nothing` (compute_euphotic_depth.jl:44-46) and
    `compute_mixed_layer_mean!(::ConstantField, …) = nothing` (mean_mixed_layer_properties.jl:21)."""
    return Field(grid, torch.full(grid.plane_shape, float(value), dtype=torch.float64, device=grid.device), name, True)


def Field2D(grid: RectilinearGrid, name: str = "", fill: float = 0.0) -> Field:
    """`Field{Center, Center, Nothing}(grid)` — indexed [i, j, k] with k ignored."""
    return Field(grid, torch.full(grid.plane_shape, fill, dtype=torch.float64, device=grid.device), name)


def current_stream_ptr(device) -> int:
    """The caller's stream (the Julia glue passes `CUDA.stream()`); here torch's current stream."""
    if torch.device(device).type != "cuda":
        raise RuntimeError("oceanbiome.jl_b200 kernels need CUDA tensors: there is no CPU fallback")
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(*fields):
    for f in fields:
        t = f.data if isinstance(f, Field) else f
        if t is not None and not t.is_cuda:
            raise RuntimeError("oceanbiome.jl_b200 kernels need CUDA tensors: there is no CPU fallback")
