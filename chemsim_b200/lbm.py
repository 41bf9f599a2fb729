"""Host-side mirror of the reference's `lbm` module over the C ABI.

Same names, argument meaning and error behaviour as /root/reference/src/lbm.rs
(and the parts of src/matrix.rs its callers use), so that code written against
`chemsim::lbm` — main.rs's `initial_state`, `LBMSim::step/render` — reads the
same here.  All arithmetic happens in the CUDA library behind
include/chemsim_lbm.h; this file only marshals buffers.  (The reference is Rust;
no Rust toolchain exists in this image, so the shim a maintainer would add to
lbm.rs is given as source in INTEGRATION.md and this Python mirror is what the
tests and bench drive.)
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _ffi
from ._ffi import EDGE_PERIODIC, EDGE_ZEROFILL, InvalidSliceSize, LbmError  # noqa: F401

Scalar = np.float32  # `pub type Scalar = f32`, src/lbm.rs:13


def _dtype_code(dtype) -> int:
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return _ffi.F32
    if dtype == np.float64:
        return _ffi.F64
    raise TypeError(f"lattice dtype must be float32 or float64, not {dtype}")


class Matrix:
    """matrix::Matrix (src/matrix.rs:10-13) as a host value: shape (w, h), element
    (y, x) at slice[y*w + x]."""

    def __init__(self, array: np.ndarray):
        assert array.ndim == 2
        self.array = array  # (h, w)

    @staticmethod
    def new(slice_, dims, dtype=Scalar) -> "Matrix":
        """Matrix::new(slice, (w, h)) -> Result (src/matrix.rs:24-30)."""
        w, h = dims
        flat = np.asarray(slice_, dtype=dtype).reshape(-1)
        if flat.size != w * h:
            raise InvalidSliceSize(_ffi.ERR_INVALID_SLICE_SIZE, f"slice has {flat.size} elements, dims are {w}x{h}")
        return Matrix(np.ascontiguousarray(flat.reshape(h, w)))

    @staticmethod
    def new_filled(value, dims, dtype=Scalar) -> "Matrix":
        """Matrix::new_filled (src/matrix.rs:40-44), with the intended (w, h) meaning
        (the reference builds dims [w, h] untransposed and so only works for w == h)."""
        w, h = dims
        return Matrix(np.full((h, w), value, dtype=dtype))

    def get_width(self) -> int:
        return self.array.shape[1]

    def get_height(self) -> int:
        return self.array.shape[0]

    def get_shape(self):
        return (self.get_width(), self.get_height())

    def get_underlying(self) -> np.ndarray:
        """Row-major host copy, index y*w + x (src/matrix.rs:120-126)."""
        return self.array.reshape(-1).copy()


@dataclass(frozen=True)
class Discretization:
    """src/lbm.rs:75-86"""
    delta_x: float = 1.0
    delta_t: float = 1.0

    def isothermal_speed_of_sound(self, dtype=Scalar):
        t = np.dtype(dtype).type
        return t(self.delta_x) / (np.sqrt(t(3.0)) * t(self.delta_t))


@dataclass(frozen=True)
class Direction:
    """src/lbm.rs:90-95"""
    w_scalar: float
    c_vector: tuple
    stencil: tuple


@dataclass(frozen=True)
class BGK:
    """src/lbm.rs:345-370"""
    tau: float

    def kinematic_shear_viscosity(self, disc: Discretization, dtype=Scalar):
        t = np.dtype(dtype).type
        dx, dt, tau = t(disc.delta_x), t(disc.delta_t), t(self.tau)
        return (dx * dx / (t(3.0) * dt * dt)) * (tau - dt / t(2.0))

    def kinematic_bulk_viscosity(self, disc: Discretization, dtype=Scalar):
        t = np.dtype(dtype).type
        return t(2.0) * self.kinematic_shear_viscosity(disc, dtype) / t(3.0)

    def _apply(self, handle, disc=None, dtype=None):
        _ffi.check(_ffi.load().chemsim_lbm_set_bgk(handle, float(self.tau)), handle)


@dataclass(frozen=True)
class TRT:
    """src/lbm.rs:374-451"""
    tau_minus: float
    tau_plus: float

    @staticmethod
    def new(lambda_, ks_viscosity, disc: Discretization, dtype=Scalar) -> "TRT":
        """TRT::new(lambda, ks_viscosity, &disc), src/lbm.rs:380-390 (host arithmetic in Scalar)."""
        t = np.dtype(dtype).type
        dt = t(disc.delta_t)
        cs = disc.isothermal_speed_of_sound(dtype)
        tau_plus = dt * ((t(ks_viscosity) / (cs * cs)) + t(0.5))
        tau_minus = dt * ((t(lambda_) / ((tau_plus / dt) - t(0.5))) + t(0.5))
        return TRT(tau_minus=float(tau_minus), tau_plus=float(tau_plus))

    def lambda_(self, disc: Discretization, dtype=Scalar):
        t = np.dtype(dtype).type
        dt = t(disc.delta_t)
        return t(1.0) * ((t(self.tau_plus) / dt) - t(0.5)) * ((t(self.tau_minus) / dt) - t(0.5))

    def kinematic_shear_viscosity(self, disc: Discretization, dtype=Scalar):
        t = np.dtype(dtype).type
        cs = disc.isothermal_speed_of_sound(dtype)
        return cs * cs * (t(self.tau_plus) / t(disc.delta_t) - t(0.5))

    def _apply(self, handle, disc=None, dtype=None):
        _ffi.check(_ffi.load().chemsim_lbm_set_trt(handle, float(self.tau_plus), float(self.tau_minus)), handle)


@dataclass(frozen=True)
class KBC:
    """src/lbm.rs:455-590"""
    ks_viscosity: float

    @staticmethod
    def new(ks_viscosity) -> "KBC":
        return KBC(ks_viscosity)

    def kinematic_shear_viscosity(self, disc: Discretization, dtype=Scalar):
        return np.dtype(dtype).type(self.ks_viscosity)

    def _apply(self, handle, disc=None, dtype=None):
        _ffi.check(_ffi.load().chemsim_lbm_set_kbc(handle, float(self.ks_viscosity)), handle)


@dataclass(frozen=True)
class Regularized:
    """src/lbm.rs:596-666: wraps an operator but only ever uses its viscosity."""
    underlying: object

    @staticmethod
    def new(underlying) -> "Regularized":
        return Regularized(underlying)

    def kinematic_shear_viscosity(self, disc: Discretization, dtype=Scalar):
        return self.underlying.kinematic_shear_viscosity(disc, dtype)

    def _apply(self, handle, disc: Discretization = Discretization(), dtype=Scalar):
        # Regularized::kinematic_shear_viscosity(disc) = underlying.kinematic_shear_viscosity(disc)
        # (src/lbm.rs:663-665), with the State's discretization and dtype
        nu = float(self.underlying.kinematic_shear_viscosity(disc, dtype))
        _ffi.check(_ffi.load().chemsim_lbm_set_regularized(handle, nu), handle)


class EquilibriumPopulations:
    """Result of compute_equilibrium: the nine equilibrium populations of (rho, u),
    kept as their generating fields so that they are evaluated on the GPU when the
    State is built instead of crossing the bus nine times."""

    def __init__(self, density: Matrix, velocity, discretization: Discretization):
        self.density, self.velocity, self.discretization = density, velocity, discretization

    def __len__(self):
        return 9


def compute_equilibrium(density: Matrix, velocity, directions, discretization: Discretization):
    """lbm::compute_equilibrium (src/lbm.rs:43-71)."""
    vx, vy = velocity
    size = density.get_shape()
    assert size == vx.get_shape()   # src/lbm.rs:51
    assert size == vy.get_shape()   # src/lbm.rs:52
    assert len(directions) == 9
    return EquilibriumPopulations(density, (vx, vy), discretization)


class D2Q9:
    """src/lbm.rs:180-323"""
    W_NUM = (16, 4, 4, 4, 4, 1, 1, 1, 1)
    C = ((0, 0), (1, 0), (0, 1), (-1, 0), (0, -1), (1, 1), (-1, 1), (-1, -1), (1, -1))
    STENCIL_ONE = (4, 3, 7, 5, 1, 6, 8, 2, 0)  # index of the 1 in each row-major 3x3 stencil, :233-269

    def __init__(self, populations):
        self.populations = populations
        self.size = None

    @staticmethod
    def new(populations) -> "D2Q9":
        """D2Q9::new(&[Population; 9]) (src/lbm.rs:187-200): nine Matrix objects, or the
        value of compute_equilibrium."""
        assert len(populations) == 9           # src/lbm.rs:188
        lat = D2Q9(populations)
        if isinstance(populations, EquilibriumPopulations):
            lat.size = populations.density.get_shape()
        else:
            lat.size = populations[0].get_shape()
            for pop in populations:
                assert pop.get_shape() == lat.size   # src/lbm.rs:191
        return lat

    @staticmethod
    def directions():
        out = []
        for i in range(9):
            st = [0] * 9
            st[D2Q9.STENCIL_ONE[i]] = 1
            out.append(Direction(D2Q9.W_NUM[i] / 36.0, D2Q9.C[i], tuple(st)))
        return out


class State:
    """lbm::State<D2Q9> (src/lbm.rs:670-819), device-resident behind the C ABI."""

    def __init__(self, handle, dtype, collision, discretization):
        self._h = handle
        self._lib = _ffi.load()
        self.dtype = np.dtype(dtype)
        self.collision = collision
        self.discretization = discretization
        w, h, hg, r0 = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        self._check(self._lib.chemsim_lbm_shape(handle, C.byref(w), C.byref(h), C.byref(hg), C.byref(r0)))
        self.width, self.local_height, self.global_height, self.row_offset = w.value, h.value, hg.value, r0.value
        self._n = self.width * self.local_height

    # ---- construction ---------------------------------------------------------
    @classmethod
    def create(cls, size, collision, discretization=Discretization(), dtype=Scalar, edge=EDGE_ZEROFILL,
               device=-1, rank=0, nranks=1, nccl_id: bytes | None = None) -> "State":
        lib = _ffi.load()
        w, h = size
        handle = C.c_void_p()
        if nranks > 1:
            idbuf = C.create_string_buffer(nccl_id, _ffi.NCCL_ID_BYTES)
            st = lib.chemsim_lbm_create_slab(w, h, _dtype_code(dtype), edge, device, rank, nranks, idbuf,
                                             C.byref(handle))
        else:
            st = lib.chemsim_lbm_create(w, h, _dtype_code(dtype), edge, device, C.byref(handle))
        _ffi.check(st, None)
        try:
            self = cls(handle, dtype, collision, discretization)
            self._check(lib.chemsim_lbm_set_discretization(handle, float(discretization.delta_x),
                                                           float(discretization.delta_t)))
            collision._apply(handle, discretization, dtype)
        except BaseException:
            lib.chemsim_lbm_destroy(handle)      # do not leak the device-resident State
            raise
        return self

    @classmethod
    def initial(cls, lattice: D2Q9, geometry, collision, discretization: Discretization, edge=EDGE_ZEROFILL,
                device=-1) -> "State":
        """State::initial(Box<L>, Geometry, Box<CollisionOperator>, Discretization)
        (src/lbm.rs:679-692).  `geometry` is a (h, w) / flat y*w+x bool array
        (main.rs:269-312).  `edge` is the one extension: the reference is always
        zero-fill."""
        pops = lattice.populations
        if isinstance(pops, EquilibriumPopulations):
            dtype = pops.density.array.dtype
        else:
            dtype = pops[0].array.dtype
        self = cls.create(lattice.size, collision, discretization, dtype, edge, device)
        if isinstance(pops, EquilibriumPopulations):
            self.init_equilibrium(pops.density.array, pops.velocity[0].array, pops.velocity[1].array)
        else:
            for q, pop in enumerate(pops):
                self.set_population(q, pop.array)
        self.geometry = geometry
        return self

    def close(self):
        if self._h:
            self._lib.chemsim_lbm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, status):
        _ffi.check(status, self._h)

    def _in(self, a, dtype=None):
        a = np.ascontiguousarray(a, dtype=dtype or self.dtype)
        return a, a.ctypes.data_as(C.c_void_p), a.size

    # ---- uploads ----------------------------------------------------------------
    def init_equilibrium(self, rho, vx, vy):
        rho, prho, n = self._in(rho)
        vx, pvx, n1 = self._in(vx)
        vy, pvy, n2 = self._in(vy)
        if not (n == n1 == n2):
            raise InvalidSliceSize(_ffi.ERR_INVALID_SLICE_SIZE, "rho, vx, vy differ in size")
        self._check(self._lib.chemsim_lbm_init_equilibrium(self._h, prho, pvx, pvy, n))

    def init_equilibrium_rows(self, row_begin, rho, vx, vy):
        """init_equilibrium for rows [row_begin, row_begin + rho.shape[0]) only."""
        rho, prho, n = self._in(rho)
        vx, pvx, n1 = self._in(vx)
        vy, pvy, n2 = self._in(vy)
        if not (n == n1 == n2):
            raise InvalidSliceSize(_ffi.ERR_INVALID_SLICE_SIZE, "rho, vx, vy differ in size")
        self._check(self._lib.chemsim_lbm_init_equilibrium_rows(self._h, row_begin, n // self.width, prho, pvx, pvy, n))

    def set_population(self, q, field):
        a, p, n = self._in(field)
        self._check(self._lib.chemsim_lbm_set_population(self._h, q, p, n))

    @property
    def geometry(self) -> np.ndarray:
        out = np.empty((self.local_height, self.width), dtype=np.uint8)
        self._check(self._lib.chemsim_lbm_get_geometry(self._h, out.ctypes.data_as(C.c_void_p), out.size))
        return out.astype(bool)

    @geometry.setter
    def geometry(self, solid):
        a, p, n = self._in(np.asarray(solid).astype(np.uint8, copy=False), np.uint8)
        self._check(self._lib.chemsim_lbm_set_geometry(self._h, p, n))

    def set_geometry_rows(self, row_begin, solid):
        """Rewrite rows [row_begin, row_begin + solid.shape[0]) of the geometry."""
        a, p, n = self._in(np.asarray(solid).astype(np.uint8, copy=False), np.uint8)
        self._check(self._lib.chemsim_lbm_set_geometry_rows(self._h, row_begin, n // self.width, p, n))

    def set_geometry_async(self, pinned_solid_ptr: int, n: int):
        """Asynchronous upload from page-locked memory (pointer + element count)."""
        self._check(self._lib.chemsim_lbm_set_geometry_async(self._h, C.c_void_p(pinned_solid_ptr), n))

    def density_async(self, pinned_dst_ptr: int, n: int):
        """Asynchronous State::density into page-locked memory; valid after synchronize()."""
        self._check(self._lib.chemsim_lbm_get_density_async(self._h, C.c_void_p(pinned_dst_ptr), n))

    def get_async(self, field: int, pinned_dst0: int, n: int, pinned_dst1: int | None = None, q: int = 0):
        """Asynchronous form of any readout (field = _ffi.FIELD_*) into page-locked memory
        (pointers + element count); valid after synchronize().  Two snapshots may be in flight."""
        self._check(self._lib.chemsim_lbm_get_async(self._h, field, q, C.c_void_p(pinned_dst0),
                                                    C.c_void_p(pinned_dst1) if pinned_dst1 else None, n))

    def fill_geometry(self, value: bool):
        """state.geometry = all `value`, on the device."""
        self._check(self._lib.chemsim_lbm_fill_geometry(self._h, int(bool(value))))

    def paint_rect(self, x0: int, y0: int, width: int, height: int, value: bool = True):
        """Set geometry cells [x0, x0+width) x [y0, y0+height) (global rows) on the device."""
        self._check(self._lib.chemsim_lbm_paint_rect(self._h, x0, y0, width, height, int(bool(value))))

    def paint_brush(self, pos):
        """The reference's mouse handler (src/main.rs:71-91) without its host round trip: the geometry
        becomes exactly the 9x9 block around the cursor, row = floor(pos[1]), column = floor(pos[0])."""
        row, col = int(np.floor(pos[1])), int(np.floor(pos[0]))
        if 0 <= row < self.global_height and 0 <= col < self.width:        # main.rs:76
            self.fill_geometry(False)
            self.paint_rect(col - 4, row - 4, 9, 9, True)

    def barrier(self):
        """Device-side barrier over the ranks of a sharded lattice (asynchronous, collective)."""
        self._check(self._lib.chemsim_lbm_barrier(self._h))

    def set_stream_convention(self, mirrored: bool):
        """Select how af::convolve2's flip is read (include/chemsim_lbm.h); before the first upload."""
        self._check(self._lib.chemsim_lbm_set_stream_convention(self._h, int(bool(mirrored))))

    def set_p2p_timeout(self, seconds: float):
        self._check(self._lib.chemsim_lbm_set_p2p_timeout(self._h, float(seconds)))

    def checkpoint(self) -> np.ndarray:
        """This handle's slab (populations, geometry, time, step counter) as bytes."""
        n = C.c_size_t()
        self._check(self._lib.chemsim_lbm_checkpoint_bytes(self._h, C.byref(n)))
        buf = np.empty(n.value, dtype=np.uint8)
        self._check(self._lib.chemsim_lbm_checkpoint(self._h, buf.ctypes.data_as(C.c_void_p), buf.size))
        return buf

    def restore(self, blob):
        blob = np.ascontiguousarray(np.frombuffer(blob, dtype=np.uint8) if not isinstance(blob, np.ndarray) else blob)
        self._check(self._lib.chemsim_lbm_restore(self._h, blob.ctypes.data_as(C.c_void_p), blob.size))

    # ---- the hot path -----------------------------------------------------------
    def step(self, nsteps: int = 1):
        """State::step (src/lbm.rs:694-714), `nsteps` times; asynchronous."""
        self._check(self._lib.chemsim_lbm_step(self._h, nsteps))

    def synchronize(self):
        self._check(self._lib.chemsim_lbm_synchronize(self._h))

    @property
    def time(self) -> float:
        out = C.c_double()
        self._check(self._lib.chemsim_lbm_time(self._h, C.byref(out)))
        return out.value

    # ---- readout ----------------------------------------------------------------
    def _get1(self, fn, *pre):
        out = np.empty((self.local_height, self.width), dtype=self.dtype)
        self._check(fn(self._h, *pre, out.ctypes.data_as(C.c_void_p), out.size))
        return Matrix(out)

    def _get2(self, fn):
        a = np.empty((self.local_height, self.width), dtype=self.dtype)
        b = np.empty_like(a)
        self._check(fn(self._h, a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), a.size))
        return Matrix(a), Matrix(b)

    def size(self):
        return (self.width, self.local_height)

    def delta_x(self):
        return self.discretization.delta_x

    def delta_t(self):
        return self.discretization.delta_t

    def isothermal_speed_of_sound(self):
        return self.discretization.isothermal_speed_of_sound(self.dtype)

    def density(self) -> Matrix:
        return self._get1(self._lib.chemsim_lbm_get_density)

    def pressure(self) -> Matrix:
        return self._get1(self._lib.chemsim_lbm_get_pressure)

    def speed(self) -> Matrix:
        return self._get1(self._lib.chemsim_lbm_get_speed)

    def velocity(self):
        return self._get2(self._lib.chemsim_lbm_get_velocity)

    def momentum_density(self):
        return self._get2(self._lib.chemsim_lbm_get_momentum_density)

    def population(self, q) -> Matrix:
        return self._get1(self._lib.chemsim_lbm_get_population, q)

    def populations(self):
        dirs = D2Q9.directions()
        return [(dirs[q], self.population(q)) for q in range(9)]

    def populations_array(self) -> np.ndarray:
        return np.stack([self.population(q).array for q in range(9)])

    def equilibrium(self):
        dirs = D2Q9.directions()
        return [(dirs[q], self._get1(self._lib.chemsim_lbm_get_equilibrium, q)) for q in range(9)]

    def non_equilibrium(self):
        dirs = D2Q9.directions()
        return [(dirs[q], self._get1(self._lib.chemsim_lbm_get_non_equilibrium, q)) for q in range(9)]

    RENDER_DENSITY, RENDER_SPEED, RENDER_VELOCITY, RENDER_MOMENTUM = range(4)

    def render(self, mode: int = 0, overlay_geometry: bool = True) -> np.ndarray:
        """render_scalar_field / render_vector_field + render_geometry (src/render.rs) on the
        device: (h, w, 4) uint8 RGBA image."""
        out = np.empty((self.local_height, self.width, 4), dtype=np.uint8)
        self._check(self._lib.chemsim_lbm_render(self._h, mode, int(overlay_geometry),
                                                 out.ctypes.data_as(C.c_void_p), self._n))
        return out

    def is_unstable(self) -> bool:
        out = C.c_int()
        self._check(self._lib.chemsim_lbm_is_unstable(self._h, C.byref(out)))
        return bool(out.value)

    def total_mass(self, global_: bool = False) -> float:
        out = C.c_double()
        fn = self._lib.chemsim_lbm_total_mass_global if global_ else self._lib.chemsim_lbm_total_mass
        self._check(fn(self._h, C.byref(out)))
        return out.value

    def enable_p2p_halo(self) -> bool:
        """Collective: switch to the fused peer-memory halo; True if every rank could."""
        self._check(self._lib.chemsim_lbm_enable_p2p_halo(self._h))
        return self.halo_mode() == "p2p"

    def halo_mode(self) -> str:
        out = C.c_int()
        self._check(self._lib.chemsim_lbm_halo_mode(self._h, C.byref(out)))
        return "p2p" if out.value == 1 else "nccl"

    # ---- introspection ----------------------------------------------------------
    def cuda_stream(self) -> int:
        out = C.c_void_p()
        self._check(self._lib.chemsim_lbm_cuda_stream(self._h, C.byref(out)))
        return out.value or 0

    def kernel_launches(self) -> int:
        out = C.c_uint64()
        self._check(self._lib.chemsim_lbm_kernel_launches(self._h, C.byref(out)))
        return out.value

    def step_kernel_name(self) -> str:
        return self._lib.chemsim_lbm_step_kernel_name(self._h).decode()


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(_ffi.NCCL_ID_BYTES)
    _ffi.check(_ffi.load().chemsim_lbm_nccl_unique_id(buf), None)
    return buf.raw


def slab_rows(global_height: int, rank: int, nranks: int):
    """(row_offset, rows) of the y-slab `rank` owns (host-only)."""
    r0, rows = C.c_int(), C.c_int()
    _ffi.check(_ffi.load().chemsim_lbm_slab_rows(global_height, rank, nranks, C.byref(r0), C.byref(rows)), None)
    return r0.value, rows.value


def halo_plan(global_height: int, rank: int, nranks: int, edge: int):
    """The halo messages of one rank per exchange, in issue order (host-only):
    list of (is_send, peer, q, row) with row one of _ffi.ROW_*."""
    buf = (_ffi.HaloMsg * _ffi.HALO_PLAN_MAX)()
    n = C.c_int()
    _ffi.check(_ffi.load().chemsim_lbm_halo_plan(global_height, rank, nranks, edge, buf, C.byref(n)), None)
    return [(bool(m.is_send), m.peer, m.q, m.row) for m in buf[: n.value]]


class MultiState:
    """One lattice sharded over several GPUs of the box, driven from ONE process — what an
    unchanged single-process caller (the reference's main.rs) needs in order to use more than
    one GPU.  It owns one y-slab `State` per device; collective calls (creation, the switch to
    the peer-memory halo, the first exchange inside step, sharded render / global mass) are
    issued from one host thread per slab (ctypes releases the GIL), everything else is a loop.
    Slabs of the same process map each other's memory directly (peer access), not through
    cudaIpc.  Same readout surface as `State`, fields concatenated over the slabs."""

    def __init__(self, size, collision, discretization=Discretization(), dtype=Scalar, edge=EDGE_ZEROFILL,
                 devices=(0,), p2p=True):
        from concurrent.futures import ThreadPoolExecutor
        self.devices = list(devices)
        n = len(self.devices)
        self._pool = ThreadPoolExecutor(max_workers=n)
        self.dtype = np.dtype(dtype)
        self.width, self.global_height = size
        nccl_id = nccl_unique_id() if n > 1 else None
        self.slabs = list(self._pool.map(
            lambda r: State.create(size, collision, discretization, dtype, edge, self.devices[r], r, n, nccl_id),
            range(n)))
        if p2p and n > 1:
            self._each(lambda s: s.enable_p2p_halo())
        self.collision, self.discretization = collision, discretization

    def _each(self, fn):
        return list(self._pool.map(fn, self.slabs))

    def _rows(self, s):
        return slice(s.row_offset, s.row_offset + s.local_height)

    def close(self):
        self._each(lambda s: s.close())
        self._pool.shutdown()

    def halo_mode(self):
        return self.slabs[0].halo_mode() if len(self.slabs) > 1 else "none"

    def init_equilibrium(self, rho, vx, vy):
        for s in self.slabs:
            s.init_equilibrium(rho[self._rows(s)], vx[self._rows(s)], vy[self._rows(s)])

    @property
    def geometry(self):
        return np.concatenate([s.geometry for s in self.slabs], axis=0)

    @geometry.setter
    def geometry(self, solid):
        solid = np.asarray(solid).reshape(self.global_height, self.width)
        for s in self.slabs:
            s.geometry = solid[self._rows(s)]

    def step(self, nsteps: int = 1):
        self._each(lambda s: s.step(nsteps))

    def synchronize(self):
        self._each(lambda s: s.synchronize())

    @property
    def time(self):
        return self.slabs[0].time

    def size(self):
        return (self.width, self.global_height)

    def _cat(self, getter):
        return Matrix(np.concatenate([getter(s).array for s in self.slabs], axis=0))

    def _cat2(self, getter):
        parts = [getter(s) for s in self.slabs]
        return (Matrix(np.concatenate([p[0].array for p in parts], axis=0)),
                Matrix(np.concatenate([p[1].array for p in parts], axis=0)))

    def density(self):
        return self._cat(lambda s: s.density())

    def pressure(self):
        return self._cat(lambda s: s.pressure())

    def speed(self):
        return self._cat(lambda s: s.speed())

    def velocity(self):
        return self._cat2(lambda s: s.velocity())

    def momentum_density(self):
        return self._cat2(lambda s: s.momentum_density())

    def population(self, q):
        return self._cat(lambda s: s.population(q))

    def populations_array(self):
        return np.concatenate([s.populations_array() for s in self.slabs], axis=1)

    def is_unstable(self):
        return any(s.is_unstable() for s in self.slabs)

    def total_mass(self):
        return float(sum(s.total_mass() for s in self.slabs))

    def render(self, mode=0, overlay_geometry=True):
        return np.concatenate(self._each(lambda s: s.render(mode, overlay_geometry)), axis=0)

    def kernel_launches(self):
        return sum(s.kernel_launches() for s in self.slabs)
