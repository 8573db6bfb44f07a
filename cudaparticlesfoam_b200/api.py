"""Python host side above the C ABI (tests, bench.py).

The production host of this path is C++/OpenFOAM: the two glue snippets ``src/initCuda.H`` and
``src/advect.H`` of this repository, which call the same C ABI (include/cpf.h).  This module mirrors
those two snippets one-to-one so the Python tests exercise exactly the calls the glue makes:

* :meth:`ParticleTracker.init_cuda`  == ``#include "initCuda.H"``  (/root/reference/src/initCuda.H)
* :meth:`ParticleTracker.advect`     == ``#include "advect.H"``    (/root/reference/src/advect.H)

Dictionary keys and defaults are those of ``system/cudaParticlesDict`` (initCuda.H:50-57).
numpy arrays in, numpy arrays out; device memory is owned by the library.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import CpfConfig, CpfStats

INTERP_TET, INTERP_VERTEX = 0, 1
LOCATOR_CONVEX, LOCATOR_BARY = 0, 1
EULER, RK2, RK4 = 0, 1, 4
RNG_NONE, RNG_XORWOW, RNG_PHILOX = 0, 1, 2
PATCH_REFLECT, PATCH_ESCAPE = 0, 1
PATH_FILTERED, PATH_EXACT = 0, 1


class CpfError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libcpf error {code}: {msg}")
        self.code = code


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


@dataclass
class ParticlesDict:
    """system/cudaParticlesDict (initCuda.H:50-57 getOrDefault values)."""

    seedingBox: tuple = ((0.0, 0.0, 0.0), (30.0, 30.0, 30.0))
    numParticles: int = 1000
    startTime: float = 0.0
    endTime: float = 1e5
    dt: float = 1e-4
    diffusionCoeff: float = 5.7e-6
    saveInterval: int = 10


class ParticleTracker:
    """One opaque library handle == the ~35 locals initCuda.H leaves in main()'s scope."""

    def __init__(self, **cfg):
        self.lib = _lib.load()
        c = CpfConfig()
        self.lib.cpf_default_config(C.byref(c))
        for k, v in cfg.items():
            if not hasattr(c, k):
                raise TypeError(f"unknown config field {k}")
            setattr(c, k, v)
        self.cfg = c
        self.h = C.c_void_p()
        rc = self.lib.cpf_create(C.byref(c), C.byref(self.h))
        if rc:
            raise CpfError(rc, self.lib.cpf_last_error(None).decode())
        self.step = 0  # initCuda.H:498
        self.dict = ParticlesDict(dt=c.dt, diffusionCoeff=c.diffusion_coeff, saveInterval=c.save_interval)

    # ------------------------------------------------------------------ plumbing
    def _chk(self, rc: int):
        if rc:
            raise CpfError(rc, self.lib.cpf_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.cpf_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_config(self, **cfg):
        for k, v in cfg.items():
            if not hasattr(self.cfg, k):
                raise TypeError(f"unknown config field {k}")
            setattr(self.cfg, k, v)
        self._chk(self.lib.cpf_set_config(self.h, C.byref(self.cfg)))

    # ------------------------------------------------------------------ mesh / field
    def upload_poly(self, pm, patch_kind=None):
        """initCuda.H:76-139: decomposition + topology + upload + locator build."""
        pk = None if patch_kind is None else np.ascontiguousarray(patch_kind, dtype=np.int32)
        npatch = len(pm.patch_starts) - 1
        tb = getattr(pm, "tet_base_pt", None)
        tb = None if tb is None else np.ascontiguousarray(tb, dtype=np.int32)
        self._chk(self.lib.cpf_mesh_upload_poly(
            self.h, pm.n_points, _dp(pm.points), pm.n_faces, _ip(pm.face_offsets), _ip(pm.face_verts), _ip(pm.owner),
            pm.n_internal, _ip(pm.neighbour), pm.n_cells, _dp(pm.cell_centres), _ip(tb) if tb is not None else None, npatch,
            _ip(pm.patch_starts), _ip(pk) if pk is not None else None))

    def set_patch_restitution(self, e):
        """Rebound model: restitution coefficient in (0, 1] per boundary patch (1 = specular, the reference)."""
        e = np.ascontiguousarray(e, dtype=np.float64)
        self._chk(self.lib.cpf_set_patch_restitution(self.h, e.shape[0], _dp(e)))

    def upload_tets(self, pos, tets, tet_cell=None, n_cells=0):
        pos = np.ascontiguousarray(pos, dtype=np.float64)
        tets = np.ascontiguousarray(tets, dtype=np.int32)
        tc = None if tet_cell is None else np.ascontiguousarray(tet_cell, dtype=np.int32)
        self._chk(self.lib.cpf_mesh_upload_tets(self.h, pos.shape[0], _dp(pos), tets.shape[0], _ip(tets),
                                                _ip(tc) if tc is not None else None, int(n_cells)))

    def mesh_info(self):
        v = [C.c_longlong() for _ in range(4)]
        self._chk(self.lib.cpf_mesh_info(self.h, *[C.byref(x) for x in v]))
        return dict(n_verts=v[0].value, n_tets=v[1].value, n_cells=v[2].value, n_boundary_faces=v[3].value)

    def download_tets(self):
        n = self.mesh_info()["n_tets"]
        tv = np.empty((n, 4), dtype=np.int32)
        tc = np.empty(n, dtype=np.int32)
        self._chk(self.lib.cpf_mesh_download_tets(self.h, _ip(tv), _ip(tc)))
        return tv, tc

    def download_neighbours(self):
        n = self.mesh_info()["n_tets"]
        nb = np.empty((n, 4), dtype=np.int32)
        self._chk(self.lib.cpf_mesh_download_neighbours(self.h, _ip(nb)))
        return nb

    def update_velocity(self, U):
        """advect.H:42-84 velocity refresh; U = cell field [nCells,3] (numpy) or a device pointer (int)."""
        if isinstance(U, (int, np.integer)):
            self._chk(self.lib.cpf_update_velocity(self.h, C.c_void_p(int(U)), 1))
        else:
            U = np.ascontiguousarray(U, dtype=np.float64)
            self._chk(self.lib.cpf_update_velocity(self.h, C.c_void_p(U.ctypes.data), 0))
            self._keep_U = U

    def update_velocity_ptr(self, ptr: int, on_device: bool):
        self._chk(self.lib.cpf_update_velocity(self.h, C.c_void_p(int(ptr)), int(on_device)))

    # ------------------------------------------------------------------ particles
    def set_particles(self, xyzw):
        xyzw = np.ascontiguousarray(xyzw, dtype=np.float64)
        assert xyzw.ndim == 2 and xyzw.shape[1] == 4
        self._chk(self.lib.cpf_set_particles(self.h, xyzw.shape[0], _dp(xyzw)))

    def seed_box(self, n, lo, hi, seed=1591593751):
        lo = np.asarray(lo, dtype=np.float64)
        hi = np.asarray(hi, dtype=np.float64)
        self._chk(self.lib.cpf_seed_box(self.h, int(n), _dp(lo), _dp(hi), C.c_ulonglong(seed)))

    def seed_box_slice(self, first, count, lo, hi, seed=1591593751):
        lo = np.asarray(lo, dtype=np.float64)
        hi = np.asarray(hi, dtype=np.float64)
        self._chk(self.lib.cpf_seed_box_slice(self.h, int(first), int(count), _dp(lo), _dp(hi), C.c_ulonglong(seed)))

    def set_tets(self, tet):
        tet = np.ascontiguousarray(tet, dtype=np.int32)
        self._chk(self.lib.cpf_set_tets(self.h, _ip(tet)))

    def locate_initial(self):
        self._chk(self.lib.cpf_locate_initial(self.h))

    def relocate_lost(self):
        self._chk(self.lib.cpf_relocate_lost(self.h))

    def reseed_inactive(self, lo, hi, seed=1591593751, count=False):
        """Continuous injection: every inactive particle goes back into the box [lo, hi]; count=True returns how many (synchronises)."""
        lo = np.asarray(lo, dtype=np.float64)
        hi = np.asarray(hi, dtype=np.float64)
        n = C.c_longlong(0)
        self._chk(self.lib.cpf_reseed_inactive(self.h, _dp(lo), _dp(hi), C.c_ulonglong(seed), C.byref(n) if count else None))
        return n.value if count else None

    def init_rng(self):
        self._chk(self.lib.cpf_init_rng(self.h))

    @property
    def n(self) -> int:
        return int(self.lib.cpf_num_particles(self.h))

    # ------------------------------------------------------------------ the two glue snippets
    def init_cuda(self, pm, U, d: ParticlesDict | None = None, particles=None, patch_kind=None):
        """#include "initCuda.H": upload mesh, seed, locate, initial advect (VTU 0 left to caller)."""
        if d is not None:
            self.dict = d
            self.set_config(dt=d.dt, diffusion_coeff=d.diffusionCoeff, save_interval=d.saveInterval)
        self.upload_poly(pm, patch_kind)
        self.update_velocity(U)
        if particles is not None:
            self.set_particles(particles)
        else:
            lo, hi = self.dict.seedingBox
            self.seed_box(self.dict.numParticles, lo, hi)
        if self.cfg.rng == RNG_XORWOW:
            self.init_rng()
        self.locate_initial()
        self._chk(self.lib.cpf_initial_advect(self.h))
        self.step = 0

    def advect(self, U, deltaT: float, run_time: float | None = None) -> int:
        """#include "advect.H": refresh the velocity, run nCycles fused sub-steps; returns nCycles."""
        if run_time is not None and not (self.dict.startTime <= run_time <= self.dict.endTime):
            return 0
        if U is not None:
            self.update_velocity(U)
        n = C.c_int()
        self._chk(self.lib.cpf_advect(self.h, float(deltaT), C.byref(n)))
        self.step += n.value
        return n.value

    def substeps(self, n: int, dt: float):
        self._chk(self.lib.cpf_substeps(self.h, int(n), float(dt)))
        self.step += n

    def sort(self):
        self._chk(self.lib.cpf_sort_particles(self.h))

    def sync(self):
        self._chk(self.lib.cpf_sync(self.h))

    def last_step_ms(self) -> float:
        ms = C.c_float()
        self._chk(self.lib.cpf_last_step_ms(self.h, C.byref(ms)))
        return ms.value

    # ------------------------------------------------------------------ results
    def download(self, pos=True, vel=True, tet=True):
        n = self.n
        p = np.empty((n, 4)) if pos else None
        v = np.empty((n, 4)) if vel else None
        t = np.empty(n, dtype=np.int32) if tet else None
        self._chk(self.lib.cpf_download(self.h, _dp(p) if pos else None, _dp(v) if vel else None, _ip(t) if tet else None))
        return p, v, t

    def download_cells(self):
        c = np.empty(self.n, dtype=np.int32)
        self._chk(self.lib.cpf_download_cells(self.h, _ip(c)))
        return c

    def stats(self) -> dict:
        s = CpfStats()
        self._chk(self.lib.cpf_stats_get(self.h, C.byref(s)))
        return {k: getattr(s, k) for k, _ in CpfStats._fields_ if k != "reserved"}

    def stats_request(self, full: bool = False):
        """Enqueue a statistics read-back (no host synchronisation); collect it later with stats_collect()."""
        self._chk(self.lib.cpf_stats_request(self.h, int(full)))

    def stats_collect(self) -> dict:
        """Result of the OLDEST outstanding stats_request (blocks only until that one has arrived)."""
        s = CpfStats()
        self._chk(self.lib.cpf_stats_collect(self.h, C.byref(s)))
        d = {k: getattr(s, k) for k, _ in CpfStats._fields_ if k != "reserved"}
        d["full"] = bool(s.reserved[0])
        return d

    # ------------------------------------------------------------------ one rank per GPU (cpf_comm.cu)
    @staticmethod
    def comm_unique_id() -> bytes:
        """rank 0: ncclGetUniqueId; distribute the bytes to the other ranks with the host's own transport."""
        buf = C.create_string_buffer(128)
        rc = _lib.load().cpf_comm_unique_id(buf, 128)
        if rc:
            raise CpfError(rc, "cpf_comm_unique_id failed (libnccl.so.2 not loadable?)")
        return buf.raw

    def comm_init(self, unique_id: bytes | None, rank: int, nranks: int):
        buf = C.create_string_buffer(unique_id, 128) if unique_id is not None else None
        self._chk(self.lib.cpf_comm_init(self.h, buf, 128 if buf is not None else 0, int(rank), int(nranks)))

    def comm_info(self):
        r, n, v = C.c_int(), C.c_int(), C.c_int()
        self._chk(self.lib.cpf_comm_info(self.h, C.byref(r), C.byref(n), C.byref(v)))
        return r.value, n.value, v.value

    def update_velocity_bcast(self, U, root: int = 0, on_device=False):
        """U: numpy cell field / device pointer on `root`, None elsewhere.  on_device: False/0 host, True/1 device memory ready
        in compute-stream order, 2 device memory that is complete already (the exchange overlaps the enqueued sub-steps)."""
        if U is None:
            self._chk(self.lib.cpf_update_velocity_bcast(self.h, None, 0, int(root)))
        elif isinstance(U, (int, np.integer)):
            self._chk(self.lib.cpf_update_velocity_bcast(self.h, C.c_void_p(int(U)), int(on_device), int(root)))
        else:
            U = np.ascontiguousarray(U, dtype=np.float64)
            self._chk(self.lib.cpf_update_velocity_bcast(self.h, C.c_void_p(U.ctypes.data), 0, int(root)))
            self._keep_U = U

    def update_velocity_slices(self, cell_offset: int, U_local):
        U_local = np.ascontiguousarray(U_local, dtype=np.float64)
        self._chk(self.lib.cpf_update_velocity_slices(self.h, int(cell_offset), U_local.shape[0], C.c_void_p(U_local.ctypes.data), 0))
        self._keep_U = U_local

    def set_particle_id_base(self, base: int):
        self._chk(self.lib.cpf_set_particle_id_base(self.h, int(base)))

    def write_vtu(self, directory: str, step: int):
        self._chk(self.lib.cpf_write_vtu(self.h, directory.encode(), int(step)))

    def write_vtu_async(self, directory: str, step: int, stride: int = 1):
        """Binary (raw appended) particle_%04d.vtu written by the library's writer thread; returns once enqueued."""
        self._chk(self.lib.cpf_write_vtu_async(self.h, directory.encode(), int(step), int(stride)))

    def output_wait(self):
        self._chk(self.lib.cpf_output_wait(self.h))

    def checkpoint_save(self, path: str):
        self._chk(self.lib.cpf_checkpoint_save(self.h, str(path).encode()))

    def checkpoint_load(self, path: str):
        """Needs the same mesh uploaded and the same random-walk configuration; continues bit-identically."""
        self._chk(self.lib.cpf_checkpoint_load(self.h, str(path).encode()))
        self.step = int(self.lib.cpf_step_index(self.h))

    def next_normals(self):
        xi = np.empty((self.n, 3))
        self._chk(self.lib.cpf_debug_next_normals(self.h, _dp(xi)))
        return xi

    def normals(self, k: int):
        """Deviates of the next k sub-steps, [k, n, 3] in original particle order (does not advance the stream)."""
        xi = np.empty((int(k), self.n, 3))
        self._chk(self.lib.cpf_debug_normals(self.h, int(k), _dp(xi)))
        return xi

    def device_pointers(self):
        a, b, c = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._chk(self.lib.cpf_device_pointers(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def set_stream(self, cuda_stream: int | None):
        self._chk(self.lib.cpf_set_stream(self.h, C.c_void_p(cuda_stream or 0)))

    def profile(self, enable: bool):
        self._chk(self.lib.cpf_profile_enable(self.h, int(enable)))

    def profile_read(self):
        n, tot, mx = C.c_int(), C.c_double(), C.c_double()
        self._chk(self.lib.cpf_profile_read(self.h, C.byref(n), C.byref(tot), C.byref(mx)))
        return n.value, tot.value, mx.value

    def launch_count(self) -> int:
        return int(self.lib.cpf_launch_count(self.h))
