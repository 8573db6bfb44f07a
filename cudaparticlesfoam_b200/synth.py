"""Synthetic OpenFOAM-style inputs for the particle hot path.

These are *input generators* for tests and bench.py -- meshes, frozen velocity fields and seeded
particle clouds of the shapes BASELINE.json names.  They emit the raw polyMesh arrays an OpenFOAM
``fvMesh`` exposes (points, faces, owner, neighbour, patch starts, cell centres), i.e. exactly what
the glue in ``src/initCuda.H`` hands to the library; the tet decomposition itself is done by the
library (``cpf_mesh_upload_poly``) and, independently, by the oracle.

No OpenFOAM is available in this environment, so the face/point numbering follows blockMesh
conventions restated from the OpenFOAM documentation: upper-triangular internal face order, boundary
patches last, face normals pointing owner -> neighbour / out of the domain, hex-model face vertex
cycles.  (reference: /root/reference/tutorials/incompressible/*/system/blockMeshDict)
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

# OpenFOAM hex cell model: local vertices 0..7 = (0,0,0),(1,0,0),(1,1,0),(0,1,0),(0,0,1),(1,0,1),(1,1,1),(0,1,1)
# outward-pointing face cycles of the hex model
_HEX_FACES = {
    "x-": (0, 4, 7, 3),
    "x+": (1, 2, 6, 5),
    "y-": (0, 1, 5, 4),
    "y+": (3, 7, 6, 2),
    "z-": (0, 3, 2, 1),
    "z+": (4, 5, 6, 7),
}
_HEX_OFFS = np.array(
    [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)], dtype=np.int64
)

PATCH_NAMES = ("x-", "x+", "y-", "y+", "z-", "z+")


@dataclass
class PolyMesh:
    """Raw polyMesh arrays (all int32 / float64, C-contiguous)."""

    points: np.ndarray  # [nPoints,3]
    face_offsets: np.ndarray  # [nFaces+1]
    face_verts: np.ndarray  # [sum face sizes]
    owner: np.ndarray  # [nFaces]
    neighbour: np.ndarray  # [nInternal]
    cell_centres: np.ndarray  # [nCells,3]
    patch_starts: np.ndarray  # [nPatches+1] face index ranges of the boundary patches
    patch_names: tuple = PATCH_NAMES
    tet_base_pt: np.ndarray | None = None  # [nFaces] mesh.tetBasePtIs(); None => 0 for every face
    dims: tuple = (0, 0, 0)
    lo: np.ndarray = field(default_factory=lambda: np.zeros(3))
    hi: np.ndarray = field(default_factory=lambda: np.ones(3))

    @property
    def n_points(self) -> int:
        return int(self.points.shape[0])

    @property
    def n_cells(self) -> int:
        return int(self.cell_centres.shape[0])

    @property
    def n_faces(self) -> int:
        return int(self.owner.shape[0])

    @property
    def n_internal(self) -> int:
        return int(self.neighbour.shape[0])


def _splitmix64(x: np.ndarray) -> np.ndarray:
    """Vectorised SplitMix64 (public-domain reference constants)."""
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def uniform01(seed: int, n: int, stream: int = 0) -> np.ndarray:
    """n reproducible doubles in [0,1) from (seed, stream, index) -- explicit host RNG so that the
    oracle, the reference kernels and the product all see the same cloud."""
    with np.errstate(over="ignore"):
        idx = np.arange(n, dtype=np.uint64)
        key = _splitmix64(np.uint64(seed) ^ (np.uint64(stream) * np.uint64(0xD1342543DE82EF95)))
        bits = _splitmix64(idx * np.uint64(0x2545F4914F6CDD1D) + key)
    return (bits >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def box_mesh(nx: int, ny: int, nz: int, lo=(0.0, 0.0, 0.0), hi=(1.0, 1.0, 1.0), jitter: float = 0.0,
             seed: int = 1591593751) -> PolyMesh:
    """nx*ny*nz hex block in blockMesh ordering.  ``jitter`` (fraction of the local spacing, < 0.3)
    displaces interior points pseudo-randomly (boundary points slide within their boundary plane)
    so that tets are irregular and faces non-planar, like a real body-fitted mesh."""
    lo = np.asarray(lo, dtype=np.float64)
    hi = np.asarray(hi, dtype=np.float64)
    npx, npy, npz = nx + 1, ny + 1, nz + 1
    ii, jj, kk = np.meshgrid(np.arange(npx), np.arange(npy), np.arange(npz), indexing="ij")
    # point id = i + npx*(j + npy*k)  -> order arrays as (k,j,i)
    I = ii.transpose(2, 1, 0).reshape(-1)
    J = jj.transpose(2, 1, 0).reshape(-1)
    K = kk.transpose(2, 1, 0).reshape(-1)
    h = (hi - lo) / np.array([nx, ny, nz], dtype=np.float64)
    pts = np.empty((I.size, 3), dtype=np.float64)
    pts[:, 0] = lo[0] + I * h[0]
    pts[:, 1] = lo[1] + J * h[1]
    pts[:, 2] = lo[2] + K * h[2]
    if jitter > 0.0:
        n = I.size
        for ax, (idx, nmax) in enumerate(((I, nx), (J, ny), (K, nz))):
            r = uniform01(seed, n, stream=101 + ax) * 2.0 - 1.0
            interior = (idx > 0) & (idx < nmax)
            pts[:, ax] += np.where(interior, r * jitter * h[ax], 0.0)

    def pid(i, j, k):
        return i + npx * (j + npy * k)

    ci, cj, ck = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    CI = ci.transpose(2, 1, 0).reshape(-1).astype(np.int64)
    CJ = cj.transpose(2, 1, 0).reshape(-1).astype(np.int64)
    CK = ck.transpose(2, 1, 0).reshape(-1).astype(np.int64)
    ncell = CI.size
    cell_id = CI + nx * (CJ + ny * CK)
    assert np.array_equal(cell_id, np.arange(ncell))
    # the 8 corner point ids of every cell
    corners = np.stack([pid(CI + o[0], CJ + o[1], CK + o[2]) for o in _HEX_OFFS], axis=1)  # [ncell,8]

    # internal faces: for each cell ascending, faces towards +x, +y, +z neighbours (ascending nbr id)
    has = [CI < nx - 1, CJ < ny - 1, CK < nz - 1]
    nbr = [cell_id + 1, cell_id + nx, cell_id + nx * ny]
    cyc = [_HEX_FACES["x+"], _HEX_FACES["y+"], _HEX_FACES["z+"]]
    slot = np.stack(has, axis=1)  # [ncell,3]
    order = np.cumsum(slot.reshape(-1)).reshape(ncell, 3) - 1  # face id of (cell, dir) where slot
    n_internal = int(slot.sum())
    int_verts = np.empty((n_internal, 4), dtype=np.int64)
    int_owner = np.empty(n_internal, dtype=np.int64)
    int_nbr = np.empty(n_internal, dtype=np.int64)
    for d in range(3):
        m = slot[:, d]
        fid = order[m, d]
        int_verts[fid] = corners[m][:, cyc[d]]
        int_owner[fid] = cell_id[m]
        int_nbr[fid] = nbr[d][m]

    # boundary patches, each ordered by cell id
    b_verts, b_owner, starts = [], [], [n_internal]
    sel = {
        "x-": CI == 0, "x+": CI == nx - 1, "y-": CJ == 0, "y+": CJ == ny - 1, "z-": CK == 0, "z+": CK == nz - 1,
    }
    for name in PATCH_NAMES:
        m = sel[name]
        b_verts.append(corners[m][:, _HEX_FACES[name]])
        b_owner.append(cell_id[m])
        starts.append(starts[-1] + int(m.sum()))
    verts = np.concatenate([int_verts] + b_verts, axis=0)
    owner = np.concatenate([int_owner] + b_owner, axis=0)
    nfaces = verts.shape[0]
    centres = pts[corners].mean(axis=1)
    if jitter == 0.0:
        centres[:, 0] = lo[0] + (CI + 0.5) * h[0]
        centres[:, 1] = lo[1] + (CJ + 0.5) * h[1]
        centres[:, 2] = lo[2] + (CK + 0.5) * h[2]
    return PolyMesh(
        points=np.ascontiguousarray(pts),
        face_offsets=(np.arange(nfaces + 1, dtype=np.int64) * 4).astype(np.int32),
        face_verts=np.ascontiguousarray(verts.reshape(-1).astype(np.int32)),
        owner=owner.astype(np.int32),
        neighbour=int_nbr.astype(np.int32),
        cell_centres=np.ascontiguousarray(centres),
        patch_starts=np.asarray(starts, dtype=np.int32),
        dims=(nx, ny, nz),
        lo=lo,
        hi=hi,
    )


def polyhex_mesh(nx: int, ny: int, nz: int, lo=(0.0, 0.0, 0.0), hi=(1.0, 1.0, 1.0), jitter: float = 0.1,
                 seed: int = 1591593751) -> PolyMesh:
    """Polyhedral stand-in (BASELINE config 4 code path): the hex block with an extra point in the middle
    of every x-directed edge, so the faces normal to y and z are hexagons, cells have 12 points and 20 tets
    (variable fan sizes, tet != 12*cell + j), and tetBasePtIs is non-trivial (hexagons fan from their
    first mid-edge point, like OpenFOAM picks base points that avoid sliver tets)."""
    base = box_mesh(nx, ny, nz, lo=lo, hi=hi, jitter=jitter, seed=seed)
    npx, npy, npz = nx + 1, ny + 1, nz + 1
    n0 = base.n_points

    def pid(i, j, k):
        return i + npx * (j + npy * k)

    def mid(i, j, k):  # midpoint of the edge (i,j,k)-(i+1,j,k)
        return n0 + i + nx * (j + npy * k)

    I, J, K = np.meshgrid(np.arange(nx), np.arange(npy), np.arange(npz), indexing="ij")
    I = I.transpose(2, 1, 0).reshape(-1); J = J.transpose(2, 1, 0).reshape(-1); K = K.transpose(2, 1, 0).reshape(-1)
    mids = 0.5 * (base.points[pid(I, J, K)] + base.points[pid(I + 1, J, K)])
    h = (np.asarray(hi, float) - np.asarray(lo, float)) / np.array([nx, ny, nz], float)
    # push interior mid-edge points a little off the straight edge (keeps boundary planes planar)
    ry = (uniform01(seed, mids.shape[0], stream=211) * 2 - 1) * 0.08 * h[1]
    rz = (uniform01(seed, mids.shape[0], stream=212) * 2 - 1) * 0.08 * h[2]
    mids[:, 1] += np.where((J > 0) & (J < ny), ry, 0.0)
    mids[:, 2] += np.where((K > 0) & (K < nz), rz, 0.0)
    assert np.array_equal(mid(I, J, K), n0 + np.arange(mids.shape[0]))
    points = np.concatenate([base.points, mids], axis=0)
    # rebuild faces: insert the mid-edge point between two consecutive face points that differ in i only
    inv = {}
    P = np.arange(n0)
    pi_, pj_, pk_ = P % npx, (P // npx) % npy, P // (npx * npy)
    off, verts, tet_base = [0], [], []
    fv = base.face_verts.reshape(-1, 4)
    for f in range(fv.shape[0]):
        q = fv[f]
        out = []
        for a in range(4):
            u, v = int(q[a]), int(q[(a + 1) % 4])
            out.append(u)
            if pj_[u] == pj_[v] and pk_[u] == pk_[v] and abs(int(pi_[u]) - int(pi_[v])) == 1:
                out.append(int(mid(min(pi_[u], pi_[v]), pj_[u], pk_[u])))
        verts.extend(out)
        off.append(len(verts))
        tet_base.append(1 if len(out) == 6 and out[1] >= n0 else (0 if len(out) == 4 else next(k for k, x in enumerate(out) if x >= n0)))
    pm = PolyMesh(points=np.ascontiguousarray(points), face_offsets=np.asarray(off, dtype=np.int32),
                  face_verts=np.asarray(verts, dtype=np.int32), owner=base.owner, neighbour=base.neighbour,
                  cell_centres=base.cell_centres, patch_starts=base.patch_starts, dims=(nx, ny, nz), lo=base.lo, hi=base.hi)
    pm.tet_base_pt = np.asarray(tet_base, dtype=np.int32)
    return pm


def honeycomb_mesh(nx: int, ny: int, nz: int, R: float = 0.06, hz: float = 0.1, jitter: float = 0.08,
                   seed: int = 1591593751) -> PolyMesh:
    """A genuinely polyhedral block (BASELINE config 4 code path): nx x ny hexagonal columns (pointy-top honeycomb of
    circumradius R) extruded in nz layers of height hz.  Cells are hexagonal prisms -- 8 faces, 12 points, 20 tets --
    three cells meet at every vertical edge, and the side boundary is a zigzag, not a plane.  Interior honeycomb vertices
    are displaced by up to ``jitter``*R (all levels alike, so every face stays planar).  Patches: z-, z+, sides."""
    s3 = np.sqrt(3.0)
    ang = np.deg2rad(30.0 + 60.0 * np.arange(6))
    key2id, pts2d, cell_v = {}, [], []
    centres2d = []
    for j in range(ny):
        for i in range(nx):
            cx, cy = s3 * R * (i + 0.5 * (j & 1)), 1.5 * R * j
            centres2d.append((cx, cy))
            ids = []
            for k in range(6):
                x, y = cx + R * np.cos(ang[k]), cy + R * np.sin(ang[k])
                key = (int(round(x / R * 1e6)), int(round(y / R * 1e6)))
                if key not in key2id:
                    key2id[key] = len(pts2d)
                    pts2d.append([x, y])
                ids.append(key2id[key])
            cell_v.append(ids)
    pts2d = np.asarray(pts2d)
    n2 = pts2d.shape[0]
    # interior vertices (shared by three columns) get a reproducible in-plane displacement
    use = np.zeros(n2, dtype=np.int64)
    for ids in cell_v:
        use[ids] += 1
    r = uniform01(seed, n2, stream=301) * jitter * R
    th = uniform01(seed, n2, stream=302) * 2.0 * np.pi
    inner = use == 3
    pts2d[inner, 0] += (r * np.cos(th))[inner]
    pts2d[inner, 1] += (r * np.sin(th))[inner]
    points = np.concatenate([np.column_stack([pts2d, np.full(n2, hz * k)]) for k in range(nz + 1)], axis=0)

    def pid(v, k):
        return v + n2 * k

    ncol = nx * ny
    faces = {}  # sorted vertex tuple -> [loop as seen from the first (lowest-id) cell, owner, neighbour, kind]
    order = []
    for k in range(nz):
        for col in range(ncol):
            c = col + ncol * k
            v = cell_v[col]
            loops = [([pid(x, k) for x in reversed(v)], 0), ([pid(x, k + 1) for x in v], 1)]
            for e in range(6):
                a, b = v[e], v[(e + 1) % 6]
                loops.append(([pid(a, k), pid(b, k), pid(b, k + 1), pid(a, k + 1)], 2))
            for loop, kind in loops:
                key = tuple(sorted(loop))
                if key in faces:
                    faces[key][2] = c
                else:
                    faces[key] = [loop, c, -1, kind]
                    order.append(key)
    internal = sorted((faces[k] for k in order if faces[k][2] >= 0), key=lambda f: (f[1], f[2]))
    boundary = [faces[k] for k in order if faces[k][2] < 0]
    groups = [[f for f in boundary if f[3] == kind] for kind in (0, 1, 2)]
    allf = internal + groups[0] + groups[1] + groups[2]
    off = np.zeros(len(allf) + 1, dtype=np.int32)
    off[1:] = np.cumsum([len(f[0]) for f in allf])
    verts = np.asarray([x for f in allf for x in f[0]], dtype=np.int32)
    owner = np.asarray([f[1] for f in allf], dtype=np.int32)
    neighbour = np.asarray([f[2] for f in internal], dtype=np.int32)
    starts = np.cumsum([len(internal)] + [len(g) for g in groups]).astype(np.int32)
    c2 = np.asarray(centres2d)
    cc = np.concatenate([np.column_stack([c2, np.full(ncol, hz * (k + 0.5))]) for k in range(nz)], axis=0)
    return PolyMesh(points=np.ascontiguousarray(points), face_offsets=off, face_verts=verts, owner=owner, neighbour=neighbour,
                    cell_centres=np.ascontiguousarray(cc), patch_starts=starts, patch_names=("z-", "z+", "sides"), dims=(nx, ny, nz),
                    lo=points.min(axis=0), hi=points.max(axis=0))


def honeycomb_mesh_fast(nx: int, ny: int, nz: int, R: float = 0.06, hz: float = 0.1, jitter: float = 0.08,
                        seed: int = 1591593751) -> PolyMesh:
    """The same honeycomb of hexagonal prisms as honeycomb_mesh, built with array operations only (1e7 cells in well under
    a minute instead of hours): vertices are made unique by sorting quantised coordinates, vertical faces by sorting the
    2-D edges of the columns.  Vertex / face numbering differs from the loop version; geometry and topology do not."""
    s3 = np.sqrt(3.0)
    ang = np.deg2rad(30.0 + 60.0 * np.arange(6))
    jj, ii = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
    cx = (s3 * R * (ii + 0.5 * (jj & 1))).ravel()
    cy = (1.5 * R * jj).ravel().astype(np.float64)
    ncol = nx * ny
    vx = cx[:, None] + R * np.cos(ang)[None, :]
    vy = cy[:, None] + R * np.sin(ang)[None, :]
    kx = np.rint(vx / R * 1e6).astype(np.int64) + (1 << 20)
    ky = np.rint(vy / R * 1e6).astype(np.int64) + (1 << 20)
    key = (kx << 34) | ky
    _, first, inv = np.unique(key.ravel(), return_index=True, return_inverse=True)
    n2 = first.shape[0]
    cell_v = inv.reshape(ncol, 6).astype(np.int64)
    pts2d = np.column_stack([vx.ravel()[first], vy.ravel()[first]])
    use = np.bincount(inv, minlength=n2)
    r = uniform01(seed, n2, stream=301) * jitter * R
    th = uniform01(seed, n2, stream=302) * 2.0 * np.pi
    inner = use == 3  # shared by three columns: reproducible in-plane displacement, all levels alike (faces stay planar)
    pts2d[inner, 0] += (r * np.cos(th))[inner]
    pts2d[inner, 1] += (r * np.sin(th))[inner]
    points = np.empty((n2 * (nz + 1), 3))
    for k in range(nz + 1):
        points[n2 * k:n2 * (k + 1), :2] = pts2d
        points[n2 * k:n2 * (k + 1), 2] = hz * k
    col = np.arange(ncol, dtype=np.int64)
    # ---- 2-D edges of the columns -> vertical quads (shared by two columns, or on the zigzag side boundary)
    ea = cell_v.ravel()
    eb = np.roll(cell_v, -1, axis=1).ravel()
    ecol = np.repeat(col, 6)
    ekey = (np.minimum(ea, eb) << 32) | np.maximum(ea, eb)
    order = np.lexsort((ecol, ekey))           # by edge, then by column: the first entry of an edge is its lower column
    ek, ea_s, eb_s, ec_s = ekey[order], ea[order], eb[order], ecol[order]
    head = np.ones(ek.shape[0], dtype=bool)
    head[1:] = ek[1:] != ek[:-1]
    hidx = np.flatnonzero(head)
    cnt = np.diff(np.append(hidx, ek.shape[0]))
    e_a, e_b, e_own = ea_s[hidx], eb_s[hidx], ec_s[hidx]     # loop orientation as seen from the owner column
    e_nbr = np.where(cnt == 2, ec_s[np.minimum(hidx + 1, ek.shape[0] - 1)], -1)
    vint, vbnd = np.flatnonzero(e_nbr >= 0), np.flatnonzero(e_nbr < 0)

    def quads(sel, k):
        a, b = e_a[sel], e_b[sel]
        q = np.full((sel.shape[0], 6), -1, dtype=np.int64)
        q[:, 0], q[:, 1], q[:, 2], q[:, 3] = a + n2 * k, b + n2 * k, b + n2 * (k + 1), a + n2 * (k + 1)
        return q

    loops, owner, nbr = [], [], []
    for k in range(nz):                         # internal vertical faces of layer k
        loops.append(quads(vint, k)); owner.append(e_own[vint] + ncol * k); nbr.append(e_nbr[vint] + ncol * k)
    for k in range(1, nz):                      # internal horizontal faces: top of (col, k-1) = bottom of (col, k)
        loops.append(cell_v + n2 * k); owner.append(col + ncol * (k - 1)); nbr.append(col + ncol * k)
    loops, owner, nbr = np.concatenate(loops), np.concatenate(owner), np.concatenate(nbr)
    o = np.lexsort((nbr, owner))                # OpenFOAM's upper-triangular order
    loops, owner, nbr = loops[o], owner[o], nbr[o]
    n_int = owner.shape[0]
    b_loops = [cell_v[:, ::-1] + 0, cell_v + n2 * nz] + [quads(vbnd, k) for k in range(nz)]
    b_owner = [col, col + ncol * (nz - 1)] + [e_own[vbnd] + ncol * k for k in range(nz)]
    starts = np.cumsum([n_int, ncol, ncol, vbnd.shape[0] * nz]).astype(np.int32)
    loops = np.concatenate([loops] + b_loops)
    owner = np.concatenate([owner] + b_owner)
    size = (loops >= 0).sum(axis=1)
    off = np.zeros(loops.shape[0] + 1, dtype=np.int32)
    off[1:] = np.cumsum(size)
    verts = loops[loops >= 0].astype(np.int32)  # row-major: the loop order of every face is kept
    cc = np.empty((ncol * nz, 3))
    for k in range(nz):
        cc[ncol * k:ncol * (k + 1), 0], cc[ncol * k:ncol * (k + 1), 1], cc[ncol * k:ncol * (k + 1), 2] = cx, cy, hz * (k + 0.5)
    return PolyMesh(points=np.ascontiguousarray(points), face_offsets=off, face_verts=verts, owner=owner.astype(np.int32),
                    neighbour=nbr.astype(np.int32), cell_centres=np.ascontiguousarray(cc), patch_starts=starts,
                    patch_names=("z-", "z+", "sides"), dims=(nx, ny, nz), lo=points.min(axis=0), hi=points.max(axis=0))


def processor_meshes(pm: PolyMesh, nproc: int) -> list:
    """What decomposePar hands the ranks of a decomposed run: rank r owns the contiguous cell range [c0, c1) and sees a
    self-contained polyMesh of those cells -- local point / face / cell numbering, the original patches restricted to
    its cells, and one extra patch of processor faces (faces whose other cell lives on another rank; on the neighbour
    side they are stored reversed, starting from the same vertex, as OpenFOAM does)."""
    n = pm.n_cells
    off, nint, nf = pm.face_offsets, pm.n_internal, pm.n_faces
    owner, nbr = pm.owner, pm.neighbour
    out = []
    for r in range(nproc):
        c0, c1 = r * n // nproc, (r + 1) * n // nproc
        own_in = (owner >= c0) & (owner < c1)
        nb_in = np.zeros(nf, dtype=bool)
        nb_in[:nint] = (nbr >= c0) & (nbr < c1)
        internal = np.flatnonzero(own_in[:nint] & nb_in[:nint])
        groups = []
        for p0, p1 in zip(pm.patch_starts[:-1], pm.patch_starts[1:]):
            f = np.arange(p0, p1)
            groups.append((f[own_in[p0:p1]], False))
        proc_own = np.flatnonzero(own_in[:nint] & ~nb_in[:nint])
        proc_nbr = np.flatnonzero(~own_in[:nint] & nb_in[:nint])
        loops, fo, fn = [], [], []

        def loop(f, flip):
            v = pm.face_verts[off[f]:off[f + 1]]
            return np.concatenate([v[:1], v[:0:-1]]) if flip else v

        for f in internal:
            loops.append(loop(f, False)); fo.append(owner[f] - c0); fn.append(nbr[f] - c0)
        starts = [len(loops)]
        for fs, _ in groups:
            for f in fs:
                loops.append(loop(f, False)); fo.append(owner[f] - c0)
            starts.append(len(loops))
        for f in proc_own:
            loops.append(loop(f, False)); fo.append(owner[f] - c0)
        for f in proc_nbr:
            loops.append(loop(f, True)); fo.append(nbr[f] - c0)
        starts.append(len(loops))
        used = np.unique(np.concatenate(loops))
        remap = -np.ones(pm.n_points, dtype=np.int64)
        remap[used] = np.arange(used.shape[0])
        fv = remap[np.concatenate(loops)].astype(np.int32)
        fo_ = np.zeros(len(loops) + 1, dtype=np.int32)
        fo_[1:] = np.cumsum([len(l) for l in loops])
        pts = np.ascontiguousarray(pm.points[used])
        out.append(PolyMesh(points=pts, face_offsets=fo_, face_verts=fv, owner=np.asarray(fo, dtype=np.int32),
                            neighbour=np.asarray(fn, dtype=np.int32), cell_centres=np.ascontiguousarray(pm.cell_centres[c0:c1]),
                            patch_starts=np.asarray(starts, dtype=np.int32), patch_names=tuple(pm.patch_names) + ("procBoundary",),
                            dims=pm.dims, lo=pts.min(axis=0), hi=pts.max(axis=0)))
    return out


def channel_mesh(nx=400, ny=50, nz=50, jitter: float = 0.0) -> PolyMesh:
    """BASELINE config 3/5 mesh: 4 x 1 x 1 channel, inlet x-, outlet x+, walls on +-y/+-z."""
    return box_mesh(nx, ny, nz, lo=(0.0, 0.0, 0.0), hi=(4.0, 1.0, 1.0), jitter=jitter)


def backward_step_mesh() -> PolyMesh:
    """C1 stand-in for pitzDaily (12 225 hex cells, one cell thick): a 163 x 75 x 1 block scaled to the
    pitzDaily bounding box; the real multi-block grading needs blockMesh (not available)."""
    return box_mesh(163, 75, 1, lo=(-0.0206, -0.0254, -0.0005), hi=(0.29, 0.0254, 0.0005))


# ------------------------------------------------------------------------------------------------
# frozen velocity fields, evaluated at cell centres (fp64, [nCells,3])
# ------------------------------------------------------------------------------------------------
def field_uniform_vortex(x: np.ndarray, U0=(1.0, 0.0, 0.0), omega=2.0 * np.pi, R=0.2, centre=None) -> np.ndarray:
    """U = U0 + Omega x (x-xc) inside radius R (solid body), decaying as R^2/r^2 outside (Rankine)."""
    x = np.asarray(x, dtype=np.float64)
    if centre is None:
        centre = 0.5 * (x.min(axis=0) + x.max(axis=0))
    dx = x[:, 0] - centre[0]
    dy = x[:, 1] - centre[1]
    r2 = dx * dx + dy * dy
    s = np.where(r2 <= R * R, omega, omega * R * R / np.maximum(r2, 1e-300))
    U = np.empty_like(x)
    U[:, 0] = U0[0] - s * dy
    U[:, 1] = U0[1] + s * dx
    U[:, 2] = U0[2]
    return U


def field_channel(x: np.ndarray, t: float = 0.0, Umax=1.0, eps=0.05, lo=(0, 0, 0), hi=(4, 1, 1)) -> np.ndarray:
    """Plane-Poiseuille-like profile with a small travelling sinusoidal perturbation (C3/C5)."""
    x = np.asarray(x, dtype=np.float64)
    Y = (x[:, 1] - lo[1]) / (hi[1] - lo[1])
    Z = (x[:, 2] - lo[2]) / (hi[2] - lo[2])
    prof = 16.0 * Y * (1 - Y) * Z * (1 - Z)
    ph = 2.0 * np.pi * (x[:, 0] / (hi[0] - lo[0]) - t)
    U = np.empty_like(x)
    U[:, 0] = Umax * prof
    U[:, 1] = eps * Umax * np.sin(ph) * np.sin(np.pi * Y)
    U[:, 2] = eps * Umax * np.cos(ph) * np.sin(np.pi * Z)
    return U


def field_recirculation(x: np.ndarray, t: float = 0.0, Umax=1.4, eps=0.05, lo=(0, 0, 0), hi=(4, 1, 1)) -> np.ndarray:
    """Closed circulation in the x-y plane of a channel-shaped box: stream function psi = A S(X) S(Y), S(s) = sin^2(pi s),
    so that (u, v) = (d psi/dy, -d psi/dx) is divergence free and vanishes on all four x/y walls -- no through-flow, hence
    a statistically steady particle distribution under the reference's all-reflecting boundary (a through-flow parks the
    cloud on the outlet).  A small travelling z-perturbation makes the field differ from step to step."""
    x = np.asarray(x, dtype=np.float64)
    Lx, Ly = hi[0] - lo[0], hi[1] - lo[1]
    X = (x[:, 0] - lo[0]) / Lx
    Y = (x[:, 1] - lo[1]) / Ly
    Z = (x[:, 2] - lo[2]) / (hi[2] - lo[2])
    SX, SY = np.sin(np.pi * X) ** 2, np.sin(np.pi * Y) ** 2
    A = Umax * Ly / np.pi
    U = np.empty_like(x)
    U[:, 0] = A * SX * np.pi * np.sin(2.0 * np.pi * Y) / Ly
    U[:, 1] = -A * np.pi * np.sin(2.0 * np.pi * X) * SY / Lx
    U[:, 2] = eps * Umax * np.sin(2.0 * np.pi * (X - t)) * np.sin(np.pi * Z)
    return U


def seed_box(n: int, lo, hi, seed: int = 1591593751) -> np.ndarray:
    """Particle cloud [n,4] = (x,y,z,w=1) uniform in the seeding box (src/initCuda.H:50-62 keys)."""
    lo = np.asarray(lo, dtype=np.float64)
    hi = np.asarray(hi, dtype=np.float64)
    p = np.empty((n, 4), dtype=np.float64)
    for ax in range(3):
        p[:, ax] = lo[ax] + uniform01(seed, n, stream=ax) * (hi[ax] - lo[ax])
    p[:, 3] = 1.0
    return p
