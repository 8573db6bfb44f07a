"""ctypes doorway onto the CPU oracle (oracle/cpf_oracle.c) and, when present, the literal reference
kernels (oracle/_ref/libref_rtxadvect.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def _has_fma() -> bool:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return " fma " in (line + " ")
    except OSError:
        pass
    return False


def build(force: bool = False) -> None:
    """Compile the oracle (and the reference checker when /root/reference is present)."""
    so = os.path.join(_HERE, "libcpf_oracle.so")
    if force or not os.path.exists(so):
        subprocess.check_call(["make", "-C", _HERE, "libcpf_oracle.so"], stdout=subprocess.DEVNULL)
    ref = os.path.join(_HERE, "_ref", "libref_rtxadvect.so")
    if (force or not os.path.exists(ref)) and os.path.isdir("/root/reference/third_party/RTXAdvect"):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


_lib = None


def lib():
    global _lib
    if _lib is None:
        name = "libcpf_oracle.so" if _has_fma() else "libcpf_oracle_nofma.so"
        path = os.path.join(_HERE, name)
        if not os.path.exists(path):
            build()
        _lib = C.CDLL(path)
        _lib.orc_build_faces.restype = C.c_long
        _lib.orc_decompose_poly.restype = C.c_long
    return _lib


@dataclass
class TetMesh:
    """Reference-layout tet mesh (cuda/HostTetMesh.h:33-41)."""

    pos: np.ndarray  # [nVerts,3] f64   points then cell centres
    idx: np.ndarray  # [nTets,4] i32
    tet_cell: np.ndarray  # [nTets] i32
    tetfacets: np.ndarray  # [nTets,4] i32
    facets: np.ndarray  # [nFaces,4] i32
    finfo: np.ndarray  # [nFaces,2] i32
    n_boundary: int = 0

    @property
    def n_tets(self):
        return int(self.idx.shape[0])

    @property
    def n_faces(self):
        return int(self.facets.shape[0])

    def args(self):
        return (_d(self.pos), _i(self.idx), _i(self.tetfacets), _i(self.facets), _i(self.finfo))


def decompose(pm) -> tuple[np.ndarray, np.ndarray]:
    """A1: polyMesh -> tets (oracle restatement of src/initCuda.H:86-110)."""
    L = lib()
    tb = getattr(pm, "tet_base_pt", None)
    tb = None if tb is None else np.ascontiguousarray(tb, dtype=np.int32)
    a = (C.c_int(pm.n_points), C.c_int(pm.n_cells), C.c_int(pm.n_faces), C.c_int(pm.n_internal),
         _i(pm.face_offsets), _i(pm.face_verts), _i(pm.owner), _i(pm.neighbour), _i(tb) if tb is not None else None)
    n = L.orc_decompose_poly(*a, None, None)
    tets = np.empty((n, 4), dtype=np.int32)
    tet_cell = np.empty(n, dtype=np.int32)
    L.orc_decompose_poly(*a, _i(tets), _i(tet_cell))
    return tets, tet_cell


def build_faces(pos: np.ndarray, idx: np.ndarray):
    """A2: face topology (oracle restatement of HostTetMesh::getBoundaryMesh)."""
    L = lib()
    nT = idx.shape[0]
    tetfacets = np.empty((nT, 4), dtype=np.int32)
    facets = np.empty((4 * nT, 4), dtype=np.int32)
    finfo = np.empty((4 * nT, 2), dtype=np.int32)
    nb = C.c_int(0)
    ns = C.c_int(0)
    nF = L.orc_build_faces(C.c_int(pos.shape[0]), _d(pos), C.c_int(nT), _i(idx), _i(tetfacets), _i(facets),
                           _i(finfo), C.byref(nb), C.byref(ns))
    if nF < 0:
        raise MemoryError("orc_build_faces")
    return tetfacets, np.ascontiguousarray(facets[:nF]), np.ascontiguousarray(finfo[:nF]), nb.value, ns.value


def tet_mesh_from_poly(pm) -> TetMesh:
    tets, tet_cell = decompose(pm)
    pos = np.ascontiguousarray(np.concatenate([pm.points, pm.cell_centres], axis=0))
    tf, fc, fi, nb, _ = build_faces(pos, tets)
    return TetMesh(pos=pos, idx=tets, tet_cell=tet_cell, tetfacets=tf, facets=fc, finfo=fi, n_boundary=nb)


def expand_velocity(mesh: TetMesh, Ucell: np.ndarray) -> np.ndarray:
    Utet = np.empty((mesh.n_tets, 3), dtype=np.float64)
    lib().orc_update_velocity(C.c_long(mesh.n_tets), _i(mesh.tet_cell), _d(np.ascontiguousarray(Ucell)), _d(Utet))
    return Utet


@dataclass
class Cloud:
    p: np.ndarray  # [n,4]
    tet: np.ndarray  # [n] i32
    vel: np.ndarray  # [n,4]
    disp: np.ndarray  # [n,4]

    @staticmethod
    def make(p4: np.ndarray, tet: np.ndarray) -> "Cloud":
        n = p4.shape[0]
        return Cloud(np.ascontiguousarray(p4, dtype=np.float64).copy(), np.ascontiguousarray(tet, dtype=np.int32).copy(),
                     np.zeros((n, 4)), np.zeros((n, 4)))

    @property
    def n(self):
        return int(self.p.shape[0])


def locate_brute(mesh: TetMesh, p4: np.ndarray) -> np.ndarray:
    tet = np.empty(p4.shape[0], dtype=np.int32)
    lib().orc_locate_brute(C.c_long(p4.shape[0]), _d(p4), _i(tet), C.c_long(mesh.n_tets), *mesh.args())
    return tet


def substeps(mesh: TetMesh, cl: Cloud, U: np.ndarray, n_steps: int, dt: float, *, convex=True, reflect=True,
             vertex_velocity=False, xi: np.ndarray | None = None, D: float = 0.0) -> None:
    """n_steps iterations of the src/advect.H:86-184 loop body, in place."""
    U = np.ascontiguousarray(U, dtype=np.float64)
    xp = None
    if xi is not None:
        xi = np.ascontiguousarray(xi, dtype=np.float64)
        assert xi.shape == (n_steps, cl.n, 3)
        xp = _d(xi)
    lib().orc_substeps(C.c_long(cl.n), C.c_int(n_steps), _d(cl.p), _i(cl.tet), _d(cl.vel), _d(cl.disp),
                       C.c_double(dt), *mesh.args(), _d(U), C.c_int(int(vertex_velocity)), C.c_int(int(convex)),
                       C.c_int(int(reflect)), xp, C.c_double(D))


def foamtrack_substeps(mesh: TetMesh, cl: Cloud, Utet: np.ndarray, n_steps: int, dt: float, *, reflect=True, threads=0) -> int:
    """OpenFOAM-style barycentric tracking restated (oracle/cpf_foamtrack.c): the second CPU baseline of SURVEY 8(d).
    In place on cl.p / cl.tet; returns the number of face events."""
    L = lib()
    L.orc_foamtrack_substeps.restype = C.c_long
    Utet = np.ascontiguousarray(Utet, dtype=np.float64)
    return int(L.orc_foamtrack_substeps(C.c_long(cl.n), C.c_int(n_steps), _d(cl.p), _i(cl.tet), C.c_double(dt), *mesh.args(),
                                        _d(Utet), C.c_int(int(reflect)), C.c_int(int(threads))))


class FilterModel:
    """CPU model of the product's fp32 guarded walk (oracle/cpf_filter_model.c) over the oracle's mesh tables."""

    def __init__(self, mesh: TetMesh):
        L = lib()
        L.orc_filter_build.restype = C.c_double
        L.orc_filter_rec_bytes.restype = C.c_int
        self.mesh = mesh
        self.recs = np.zeros(mesh.idx.shape[0] * int(L.orc_filter_rec_bytes()), dtype=np.uint8)
        self.hmin = float(L.orc_filter_build(C.c_long(mesh.idx.shape[0]), *mesh.args(), self.recs.ctypes.data_as(C.c_void_p)))
        g = 1e-11 / self.hmin
        self.guard = g if g > 1e-7 else 1e-7

    def walk(self, p4: np.ndarray, disp4: np.ndarray, tet: np.ndarray, *, guard=None, err_scale=1.0, skip_c1_first=None):
        """-> (final tet or -1 where the filter refused, visits)"""
        n = p4.shape[0]
        out = np.empty(n, dtype=np.int32)
        vis = np.empty(n, dtype=np.int32)
        lib().orc_filter_walk(C.c_long(n), _d(np.ascontiguousarray(p4)), _d(np.ascontiguousarray(disp4)),
                              _i(np.ascontiguousarray(tet, dtype=np.int32)), self.recs.ctypes.data_as(C.c_void_p), _d(self.mesh.pos),
                              C.c_double(self.guard if guard is None else guard), C.c_double(err_scale),
                              None if skip_c1_first is None else np.ascontiguousarray(skip_c1_first, dtype=np.uint8).ctypes.data_as(C.POINTER(C.c_ubyte)),
                              _i(out), _i(vis))
        return out, vis

    def bary_walk(self, p4: np.ndarray, disp4: np.ndarray, tet: np.ndarray, *, guard=None, err_scale=1.0):
        """RTX=true build: -> (tet certified to contain P + disp or -1 where the filter refused, visits)"""
        n = p4.shape[0]
        out = np.empty(n, dtype=np.int32)
        vis = np.empty(n, dtype=np.int32)
        lib().orc_filter_bary_walk(C.c_long(n), _d(np.ascontiguousarray(p4)), _d(np.ascontiguousarray(disp4)),
                                   _i(np.ascontiguousarray(tet, dtype=np.int32)), self.recs.ctypes.data_as(C.c_void_p), _d(self.mesh.pos),
                                   C.c_double(self.guard if guard is None else guard), C.c_double(err_scale), _i(out), _i(vis))
        return out, vis

    def substep(self, cl: "Cloud", disp4: np.ndarray, *, skip_replay=False) -> np.ndarray:
        """One sub-step of every particle as the wall-capable fast pass does it; in place on cl.p / cl.vel / cl.tet where
        certified.  -> status (0 refused, 1 certified, 2 certified with one in-place wall reflection)"""
        st = np.empty(cl.n, dtype=np.int32)
        lib().orc_filter_substep(C.c_long(cl.n), _d(cl.p), _d(np.ascontiguousarray(disp4)), _d(cl.vel), _i(cl.tet),
                                 self.recs.ctypes.data_as(C.c_void_p), *self.mesh.args(), C.c_double(self.guard), C.c_int(int(skip_replay)), _i(st))
        return st


def advect(mesh, cl, U, dt, vertex_velocity=False):
    lib().orc_advect(C.c_long(cl.n), _d(cl.p), _i(cl.tet), _d(cl.vel), _d(cl.disp), C.c_double(dt), *mesh.args(),
                     _d(np.ascontiguousarray(U)), C.c_int(int(vertex_velocity)))


def brownian(cl, xi, D, dt):
    lib().orc_brownian(C.c_long(cl.n), _d(cl.p), _d(cl.disp), _d(np.ascontiguousarray(xi)), C.c_double(D), C.c_double(dt))


def locate_convex(mesh, cl):
    lib().orc_locate_convex(C.c_long(cl.n), _d(cl.p), _d(cl.disp), _i(cl.tet), *mesh.args())


def reflect_convex(mesh, cl):
    lib().orc_reflect_convex(C.c_long(cl.n), _d(cl.p), _d(cl.disp), _d(cl.vel), _i(cl.tet), *mesh.args())


def locate_bary(mesh, cl):
    lib().orc_locate_bary(C.c_long(cl.n), _d(cl.p), _d(cl.disp), _i(cl.tet), *mesh.args())


def reflect_bary(mesh, cl):
    lib().orc_reflect_bary(C.c_long(cl.n), _d(cl.p), _d(cl.disp), _d(cl.vel), _i(cl.tet), *mesh.args())


def bary_query(mesh, cl):
    lib().orc_bary_query(C.c_long(cl.n), _d(cl.p), _i(cl.tet), *mesh.args())


def tet_polyface(pm) -> np.ndarray:
    L = lib()
    L.orc_tet_polyface.restype = C.c_long
    n = 0
    sizes = np.diff(pm.face_offsets) - 2
    n = int(sizes[: pm.n_internal].sum() * 2 + sizes[pm.n_internal:].sum())
    tf = np.empty(n, dtype=np.int32)
    got = L.orc_tet_polyface(C.c_int(pm.n_cells), C.c_int(pm.n_faces), C.c_int(pm.n_internal), _i(pm.face_offsets), _i(pm.owner),
                             _i(pm.neighbour), _i(tf))
    assert got == n
    return tf


def face_kinds(pm, mesh: TetMesh, patch_kind) -> np.ndarray:
    """[nFaces] uint8 action per tet-mesh face (0 reflect, 1 escape) from per-patch kinds: the boundary
    triangle of a tet is the face opposite its centre vertex (slot 0) on the polyMesh face it was fanned from."""
    patch_kind = np.asarray(patch_kind)
    poly_patch = np.full(pm.n_faces, -1, dtype=np.int64)
    for p in range(len(pm.patch_starts) - 1):
        poly_patch[pm.patch_starts[p]:pm.patch_starts[p + 1]] = p
    tf = tet_polyface(pm)
    kinds = np.zeros(mesh.n_faces, dtype=np.uint8)
    bd = poly_patch[tf] >= 0
    kinds[mesh.tetfacets[bd, 0]] = patch_kind[poly_patch[tf[bd]]]
    return kinds


def face_gains(pm, mesh: TetMesh, restitution) -> np.ndarray:
    """[nFaces] float64 gain = 1 + restitution coefficient of the patch a boundary face lies on (2 elsewhere)."""
    restitution = np.asarray(restitution, dtype=np.float64)
    poly_patch = np.full(pm.n_faces, -1, dtype=np.int64)
    for p in range(len(pm.patch_starts) - 1):
        poly_patch[pm.patch_starts[p]:pm.patch_starts[p + 1]] = p
    tf = tet_polyface(pm)
    gains = np.full(mesh.n_faces, 2.0)
    bd = poly_patch[tf] >= 0
    gains[mesh.tetfacets[bd, 0]] = 1.0 + restitution[poly_patch[tf[bd]]]
    return gains


def point_values(pm, Ucell) -> np.ndarray:
    Uv = np.empty((pm.n_points + pm.n_cells, 3))
    lib().orc_point_values(C.c_int(pm.n_points), C.c_int(pm.n_cells), C.c_int(pm.n_faces), C.c_int(pm.n_internal), _i(pm.face_offsets),
                           _i(pm.face_verts), _i(pm.owner), _i(pm.neighbour), _d(pm.points), _d(pm.cell_centres),
                           _d(np.ascontiguousarray(Ucell, dtype=np.float64)), _d(Uv))
    return Uv


def ext_substeps(mesh: TetMesh, cl: Cloud, U, n_steps, dt, *, vertex_velocity=False, integrator=0, face_kind=None, reflect=True,
                 xi=None, D=0.0, face_gain=None) -> int:
    """Generalised loop of oracle/cpf_oracle_ext.c (RK2=1 / RK4=4, vertex interpolation, escape patches)."""
    L = lib()
    L.orc_ext_substeps.restype = C.c_long
    U = np.ascontiguousarray(U, dtype=np.float64)
    xp = None
    if xi is not None:
        xi = np.ascontiguousarray(xi, dtype=np.float64)
        xp = _d(xi)
    fk = None
    if face_kind is not None:
        face_kind = np.ascontiguousarray(face_kind, dtype=np.uint8)
        fk = face_kind.ctypes.data_as(C.POINTER(C.c_ubyte))
    return int(L.orc_ext_substeps(C.c_long(cl.n), C.c_int(n_steps), _d(cl.p), _i(cl.tet), _d(cl.vel), _d(cl.disp), C.c_double(dt),
                                  *mesh.args(), _d(U), C.c_int(int(vertex_velocity)), C.c_int(int(integrator)), fk,
                                  C.c_int(int(reflect)), xp, C.c_double(D),
                                  None if face_gain is None else _d(np.ascontiguousarray(face_gain, dtype=np.float64))))


def move(cl):
    lib().orc_move(C.c_long(cl.n), _d(cl.p), _d(cl.disp))


def bary_of(mesh: TetMesh, p4: np.ndarray, tet: np.ndarray) -> np.ndarray:
    w = np.empty((p4.shape[0], 4))
    lib().orc_bary_of(C.c_long(p4.shape[0]), _d(np.ascontiguousarray(p4)), _i(np.ascontiguousarray(tet, dtype=np.int32)),
                      _d(mesh.pos), _i(mesh.idx), _d(w))
    return w


# ------------------------------------------------------------------------------------------------
# literal reference kernels (GPU only)
# ------------------------------------------------------------------------------------------------
_ref = None


def ref_path() -> str:
    return os.path.join(_HERE, "_ref", "libref_rtxadvect.so")


def ref_available() -> bool:
    return os.path.exists(ref_path())


def ref():
    global _ref
    if _ref is None:
        _ref = C.CDLL(ref_path())
        _ref.ref_mesh_upload.restype = C.c_void_p
        _ref.ref_particles_create.restype = C.c_void_p
        _ref.ref_substeps_timed.restype = C.c_float
    return _ref


def ref_build_faces(pos: np.ndarray, idx: np.ndarray):
    """The reference's own host topology builder (valid for < 2^20 vertices)."""
    R = ref()
    nT = idx.shape[0]
    tetfacets = np.empty((nT, 4), dtype=np.int32)
    facets = np.empty((4 * nT, 4), dtype=np.int32)
    finfo = np.empty((4 * nT, 2), dtype=np.int32)
    nF = R.ref_build_faces(C.c_int(pos.shape[0]), _d(pos), C.c_int(nT), _i(idx), _i(tetfacets), _i(facets), _i(finfo),
                           C.c_int(4 * nT))
    if nF < 0:
        raise RuntimeError(f"reference builder dropped tets ({nF})")
    return tetfacets, np.ascontiguousarray(facets[:nF]), np.ascontiguousarray(finfo[:nF])


class RefRun:
    """Device-side state of the reference library for one mesh + cloud."""

    def __init__(self, mesh: TetMesh, Utet: np.ndarray, p4: np.ndarray, tet: np.ndarray, init_rng: bool = False):
        R = ref()
        self.R = R
        self.n = int(p4.shape[0])
        Utet = np.ascontiguousarray(Utet, dtype=np.float64)
        self.mh = C.c_void_p(R.ref_mesh_upload(C.c_int(mesh.pos.shape[0]), _d(mesh.pos), C.c_int(mesh.n_tets), _i(mesh.idx),
                                               _d(Utet), C.c_int(mesh.n_faces), _i(mesh.facets), _i(mesh.tetfacets),
                                               _i(mesh.finfo)))
        p4 = np.ascontiguousarray(p4, dtype=np.float64)
        tet = np.ascontiguousarray(tet, dtype=np.int32)
        self.ph = C.c_void_p(R.ref_particles_create(C.c_int(self.n), _d(p4), _i(tet), C.c_int(int(init_rng))))

    def substeps(self, n_steps, dt, *, convex=True, brownian=False, D=0.0, reflect=True, vertex_velocity=False):
        self.R.ref_substeps(self.mh, self.ph, C.c_int(n_steps), C.c_double(dt), C.c_int(int(convex)), C.c_int(int(brownian)),
                            C.c_double(D), C.c_int(int(reflect)), C.c_int(int(vertex_velocity)))

    def substeps_timed(self, n_steps, dt, *, convex=True, brownian=False, D=0.0, reflect=True) -> float:
        return float(self.R.ref_substeps_timed(self.mh, self.ph, C.c_int(n_steps), C.c_double(dt), C.c_int(int(convex)),
                                               C.c_int(int(brownian)), C.c_double(D), C.c_int(int(reflect))))

    def update_velocity(self, Utet):
        self.R.ref_update_velocity(self.mh, _d(np.ascontiguousarray(Utet, dtype=np.float64)))

    def bary_query(self):
        self.R.ref_bary_query(self.mh, self.ph)

    def draw_normals(self) -> np.ndarray:
        xi = np.empty((self.n, 3))
        self.R.ref_draw_normals(self.ph, _d(xi))
        return xi

    def set(self, p4, tet):
        self.R.ref_particles_set(self.ph, _d(np.ascontiguousarray(p4, dtype=np.float64)),
                                 _i(np.ascontiguousarray(tet, dtype=np.int32)))

    def download(self) -> Cloud:
        cl = Cloud(np.empty((self.n, 4)), np.empty(self.n, dtype=np.int32), np.empty((self.n, 4)), np.empty((self.n, 4)))
        self.R.ref_download(self.ph, _d(cl.p), _i(cl.tet), _d(cl.disp), _d(cl.vel))
        return cl

    def close(self):
        if self.ph:
            self.R.ref_particles_free(self.ph)
            self.ph = None
        if self.mh:
            self.R.ref_mesh_free(self.mh)
            self.mh = None
