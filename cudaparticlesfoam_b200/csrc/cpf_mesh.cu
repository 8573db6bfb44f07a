// cpf_mesh.cu -- device mesh builder.
//
// Replaces, for the hot path, the reference's host-side mesh preparation:
//   * tet decomposition in the glue            /root/reference/src/initCuda.H:86-110
//   * HostTetMesh::getBoundaryMesh/add1Facet   third_party/RTXAdvect/cuda/HostTetMesh.h:265-430
//     (std::map with a 3x20-bit key: minutes and GBs at 1M cells, wrong past 2^20 vertices)
//   * DeviceTetMesh::upload                    third_party/RTXAdvect/cuda/DeviceTetMesh.cuh:59-72
// Here the face topology is built ON THE DEVICE: every tet emits its four ascending vertex
// triples, one 64-bit radix sort on the two smallest ids groups candidate partners, and a
// neighbourhood scan matches the third id.  The result is a 32+2 byte tet record
// (sorted vertex ids, neighbour links, orientation code) instead of the reference's
// tetfacets -> facets -> positions -> faceinfos four-level gather (400 B per visited tet).
#include <cub/cub.cuh>

#include <cstdarg>
#include <cstdio>
#include <cstring>

#include <cmath>

#include "cpf_internal.h"

namespace cpf {

enum MeshFlag { MF_REPEATED_VERTEX = 1, MF_BAD_VOLUME = 2, MF_ZERO_DET = 4, MF_NONMANIFOLD = 8, MF_ORIENTATION = 16, MF_OPEN_FACE = 32, MF_BAD_INDEX = 64 };

// scoped device allocation for the builder's temporaries (freed on every return path)
template <typename T> struct DevBuf {
    T *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t n) { return cudaMalloc(&p, sizeof(T) * (n ? n : 1)); }
    operator T *() const { return p; }
};

struct FaceEntry { int c; int tf; }; // third vertex id, (tet<<2 | sorted face slot)

// One thread per tet: validate, sort the vertex ids, derive perm/flip code, emit 4 face keys.
__global__ void k_prepare_tets(long long nTets, int nVerts, const int4 *__restrict__ tetref, const double4 *__restrict__ vpos,
                               int4 *__restrict__ tetv, uint16_t *__restrict__ tetcode, unsigned long long *__restrict__ keys,
                               int *__restrict__ entryIdx, FaceEntry *__restrict__ entries, unsigned *__restrict__ flags,
                               unsigned long long *__restrict__ hminBits, int bits)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nTets) return;
    const int4 r = tetref[t];
    int id[4] = { r.x, r.y, r.z, r.w };
    unsigned f = 0;
    if ((unsigned)id[0] >= (unsigned)nVerts || (unsigned)id[1] >= (unsigned)nVerts || (unsigned)id[2] >= (unsigned)nVerts ||
        (unsigned)id[3] >= (unsigned)nVerts) {
        atomicOr(flags, (unsigned)MF_BAD_INDEX); // never touch positions through a bad index
        tetv[t] = make_int4(0, 0, 0, 0);
        tetcode[t] = 0;
        for (int k = 0; k < 4; ++k) { keys[4 * t + k] = ~0ull; entryIdx[4 * t + k] = (int)(4 * t + k); entries[4 * t + k] = FaceEntry{ -1, (int)((t << 2) | k) }; }
        return;
    }
    if (id[0] == id[1] || id[0] == id[2] || id[0] == id[3] || id[1] == id[2] || id[1] == id[3] || id[2] == id[3]) f |= MF_REPEATED_VERTEX;
    const D3 A = ld_vertex(vpos, id[0]), B = ld_vertex(vpos, id[1]), C = ld_vertex(vpos, id[2]), D = ld_vertex(vpos, id[3]);
    // HostTetMesh.h:334-343 volume sign test: double arithmetic without contraction, narrowed to float
    {
        const D3 e1 = xsub(B, A), e2 = xsub(C, A), e3 = xsub(D, A);
        const double cx = __dsub_rn(__dmul_rn(e1.y, e2.z), __dmul_rn(e2.y, e1.z));
        const double cy = __dsub_rn(__dmul_rn(e1.z, e2.x), __dmul_rn(e2.z, e1.x));
        const double cz = __dsub_rn(__dmul_rn(e1.x, e2.y), __dmul_rn(e2.x, e1.y));
        const float vol = (float)__dadd_rn(__dadd_rn(__dmul_rn(e3.x, cx), __dmul_rn(e3.y, cy)), __dmul_rn(e3.z, cz));
        if (!(vol > 0.f)) f |= MF_BAD_VOLUME; // zero (dropped by the reference) or inverted (swapped by it)
    }
    const double det = xdet(A, B, C, D); // particles.cu:348 den == 0 -> particle frozen
    if (det == 0.0) f |= MF_ZERO_DET;
    // rank of each reference vertex among the four ids
    int rank[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int rk = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) rk += (id[q] < id[k]) ? 1 : 0;
        rank[k] = rk;
    }
    int s[4] = { 0, 0, 0, 0 };
    unsigned code = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (!(f & MF_REPEATED_VERTEX)) s[rank[k]] = id[k];
        code |= (unsigned)rank[k] << (2 * k);
    }
    // Gmsh face order of the reference (HostTetMesh.h:351-358): face k is opposite vertex k.
    // `front` = parity of the sort of (v0,v1,v2) as add1Facet performs it (:272-275);
    // the reference negates the normal when this tet is the face's `back`, i.e. when !front.
    const int fk[4][3] = { { 1, 2, 3 }, { 2, 0, 3 }, { 0, 1, 3 }, { 0, 2, 1 } };
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int a = id[fk[k][0]], b = id[fk[k][1]], c = id[fk[k][2]];
        int front = 0;
        if (a > c) { int x = a; a = c; c = x; front ^= 1; }
        if (b > c) { int x = b; b = c; c = x; front ^= 1; }
        if (a > b) { int x = a; a = b; b = x; front ^= 1; }
        const int j = rank[k];
        if (!front) code |= 1u << (8 + j);
        const long long e = 4 * t + j;
        keys[e] = ((unsigned long long)(unsigned)a << bits) | (unsigned)b;
        entryIdx[e] = (int)e;
        entries[e] = FaceEntry{ c, (int)((t << 2) | j) };
    }
    tetv[t] = make_int4(s[0], s[1], s[2], s[3]);
    tetcode[t] = (uint16_t)code;
    if (f) atomicOr(flags, f);
    else {
        // smallest height = |det| / (largest face normal length): feeds the filter's guard band
        const D3 n0 = xcross(xsub(C, B), xsub(D, B)), n1 = xcross(xsub(C, A), xsub(D, A)), n2 = xcross(xsub(B, A), xsub(D, A)),
                 n3 = xcross(xsub(B, A), xsub(C, A));
        const double m = fmax(fmax(xdot(n0, n0), xdot(n1, n1)), fmax(xdot(n2, n2), xdot(n3, n3)));
        const double h = fabs(det) / sqrt(m);
        atomicMin(hminBits, (unsigned long long)__double_as_longlong(h));
    }
}

// One thread per sorted face entry: find the other tet with the same (a,b,c).
__global__ void k_link_faces(long long nEntries, const unsigned long long *__restrict__ keys, const int *__restrict__ order,
                             const FaceEntry *__restrict__ entries, const uint16_t *__restrict__ tetcode,
                             const int *__restrict__ tetPatch, const int4 *__restrict__ tetv, int *__restrict__ rec,
                             unsigned *__restrict__ flags, unsigned long long *__restrict__ nBoundary)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nEntries) return;
    const unsigned long long key = keys[i];
    const FaceEntry me = entries[order[i]];
    int partner = -1, count = 0;
    for (long long q = i - 1; q >= 0 && keys[q] == key; --q) {
        const FaceEntry o = entries[order[q]];
        if (o.c == me.c) { partner = o.tf; count++; }
    }
    for (long long q = i + 1; q < nEntries && keys[q] == key; ++q) {
        const FaceEntry o = entries[order[q]];
        if (o.c == me.c) { partner = o.tf; count++; }
    }
    const int t = me.tf >> 2, j = me.tf & 3;
    const unsigned code = tetcode[t];
    int link, apex = -1;
    if (count == 0) {
        int patch = 0;
        if (tetPatch) {
            // only the face opposite reference vertex 0 (the cell centre) can lie on a patch
            if ((int)(code & 3u) == j && tetPatch[t] >= 0) patch = tetPatch[t];
            else atomicOr(flags, (unsigned)MF_OPEN_FACE);
        }
        link = -(patch + 1);
        atomicAdd(nBoundary, 1ull);
    } else {
        if (count > 1) atomicOr(flags, (unsigned)MF_NONMANIFOLD);
        const unsigned ocode = tetcode[partner >> 2];
        if (((code >> (8 + j)) & 1u) == ((ocode >> (8 + (partner & 3))) & 1u)) atomicOr(flags, (unsigned)MF_ORIENTATION);
        link = partner;
        // the one vertex of the neighbour that is not on the shared face: lets the walk fetch the
        // next tet's record and its single new vertex in ONE memory round
        const int4 ov = tetv[partner >> 2];
        const int oj = partner & 3;
        apex = oj == 0 ? ov.x : (oj == 1 ? ov.y : (oj == 2 ? ov.z : ov.w));
    }
    rec[8ll * t + j] = link;
    rec[8ll * t + 4 + j] = apex;
}

// fp32 fast record (cpf_geom.cuh Fast32): offsets of the three lower-id vertices from the highest-id
// vertex, computed in fp64 and rounded once.  Stored positively oriented: where 6*volume of
// (s0-O, s1-O, s2-O) is negative, slots 1 and 2 are exchanged (coordinates and links); every link
// carries the STORED slot of the shared face inside the neighbour.
CPF_DEV double sorted_v6(const int4 v, const double4 *__restrict__ vpos)
{
    const D3 O = ld_vertex(vpos, v.w), A = ld_vertex(vpos, v.x), B = ld_vertex(vpos, v.y), C = ld_vertex(vpos, v.z);
    const double X[3][3] = { { A.x - O.x, A.y - O.y, A.z - O.z }, { B.x - O.x, B.y - O.y, B.z - O.z }, { C.x - O.x, C.y - O.y, C.z - O.z } };
    return X[0][0] * (X[1][1] * X[2][2] - X[1][2] * X[2][1]) + X[0][1] * (X[1][2] * X[2][0] - X[1][0] * X[2][2]) +
           X[0][2] * (X[1][0] * X[2][1] - X[1][1] * X[2][0]);
}

__global__ void k_build_fast(long long nTets, const int4 *__restrict__ tetv, const double4 *__restrict__ vpos,
                             const int4 *__restrict__ tetrec, uint4 *__restrict__ out, int cellCentreMesh)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nTets) return;
    const int4 v = tetv[t];
    const int4 l = tetrec[2 * t];
    const double v6 = sorted_v6(v, vpos);
    const bool flip = v6 < 0.0;
    const D3 O = ld_vertex(vpos, v.w);
    const int ids[3] = { v.x, flip ? v.z : v.y, flip ? v.y : v.z };
    int lk[4] = { l.x, flip ? l.z : l.y, flip ? l.y : l.z, l.w };
    // re-express the neighbour-side slot in the neighbour's stored order
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (lk[k] >= 0) {
            const int nt = lk[k] >> 2;
            int ns = lk[k] & 3;
            if ((ns == 1 || ns == 2) && sorted_v6(tetv[nt], vpos) < 0.0) ns = 3 - ns;
            lk[k] = (nt << 2) | ns;
        }
    double X[3][3];
    float E = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const D3 p = ld_vertex(vpos, ids[k]);
        X[k][0] = p.x - O.x; X[k][1] = p.y - O.y; X[k][2] = p.z - O.z;
#pragma unroll
        for (int c = 0; c < 3; ++c) E = fmaxf(E, fabsf((float)X[k][c]));
    }
    E = E * 1.0000002f; // never below the true maximum
    if (flip) E = -E;   // sign bit: stored slots 1 and 2 are exchanged w.r.t. the sorted vertex order (wall handling)
    // inward normals of the faces opposite slots 0,1,2 (the origin is slot 3): N_j = X_{j+1} x X_{j+2}
    float Nf[3][3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const double *u = X[(j + 1) % 3], *w = X[(j + 2) % 3];
        Nf[j][0] = (float)(u[1] * w[2] - u[2] * w[1]);
        Nf[j][1] = (float)(u[2] * w[0] - u[0] * w[2]);
        Nf[j][2] = (float)(u[0] * w[1] - u[1] * w[0]);
    }
    uint4 *o = out + 4 * t;
    o[0] = make_uint4((unsigned)lk[0], (unsigned)lk[1], (unsigned)lk[2], (unsigned)lk[3]);
    o[1] = make_uint4(__float_as_uint(Nf[0][0]), __float_as_uint(Nf[0][1]), __float_as_uint(Nf[0][2]), __float_as_uint(Nf[1][0]));
    o[2] = make_uint4(__float_as_uint(Nf[1][1]), __float_as_uint(Nf[1][2]), __float_as_uint(Nf[2][0]), __float_as_uint(Nf[2][1]));
    // word 13 (Fast32::aux): cell-centre decomposition -> origin id of the tet behind stored slot 3 (the face opposite the
    // centre: the only way into another cell); generic tet mesh -> this tet's own origin id
    int aux = v.w;
    if (cellCentreMesh && lk[3] >= 0) aux = tetv[lk[3] >> 2].w;
    o[3] = make_uint4(__float_as_uint(Nf[2][2]), (unsigned)aux, __float_as_uint((float)fabs(v6)), __float_as_uint(E));
}

// reference face normals, once per (tet, sorted face): see face_normal_exact
__global__ void k_build_normals(long long nTets, const int4 *__restrict__ tetv, const double4 *__restrict__ vpos,
                                const uint16_t *__restrict__ tetcode, double *__restrict__ out)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nTets) return;
    const int4 v = tetv[t];
    Tet T;
    T.P[0] = ld_vertex(vpos, v.x); T.P[1] = ld_vertex(vpos, v.y); T.P[2] = ld_vertex(vpos, v.z); T.P[3] = ld_vertex(vpos, v.w);
    T.code = tetcode[t];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const D3 n = face_normal_build(T, j);
        out[12 * t + 3 * j] = n.x; out[12 * t + 3 * j + 1] = n.y; out[12 * t + 3 * j + 2] = n.z;
    }
}

__global__ void k_pack_positions(long long n, const double *__restrict__ xyz, double4 *__restrict__ out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = make_double4(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], 0.0);
}

static void free_mesh(cpf_context *ctx)
{
    cudaFree(ctx->d_vpos); cudaFree(ctx->d_tetv); cudaFree(ctx->d_tetrec); cudaFree(ctx->d_tetfast); cudaFree(ctx->d_tetnrm); cudaFree(ctx->d_tetcode); cudaFree(ctx->d_tetcell);
    cudaFree(ctx->d_ucell[0]); cudaFree(ctx->d_ucell[1]); cudaFree(ctx->d_ustage); ctx->d_ustage = nullptr; cudaFree(ctx->d_uvert); cudaFree(ctx->d_patch_kind); cudaFree(ctx->d_patch_gain); ctx->d_patch_gain = nullptr; cudaFree(ctx->d_pc_off); cudaFree(ctx->d_pc_cells);
    ctx->d_vpos = nullptr; ctx->d_tetv = nullptr; ctx->d_tetrec = nullptr; ctx->d_tetfast = nullptr; ctx->d_tetnrm = nullptr; ctx->d_tetcode = nullptr; ctx->d_tetcell = nullptr;
    ctx->d_ucell[0] = ctx->d_ucell[1] = nullptr; ctx->d_uvert = nullptr; ctx->d_patch_kind = nullptr; ctx->d_pc_off = ctx->d_pc_cells = nullptr;
    free_bvh(ctx);
    ctx->have_mesh = false;
}

static int mesh_error(cpf_context *ctx, unsigned flags)
{
    return fail(ctx, CPF_ERR_MESH, "invalid tet mesh:%s%s%s%s%s%s%s", (flags & MF_BAD_INDEX) ? " vertex id out of range;" : "",
                (flags & MF_REPEATED_VERTEX) ? " repeated vertex in a tet;" : "",
                (flags & MF_BAD_VOLUME) ? " zero or negative tet volume (orient tets so that det(A,B,C,D) > 0);" : "",
                (flags & MF_ZERO_DET) ? " degenerate tet (det == 0);" : "",
                (flags & MF_NONMANIFOLD) ? " face shared by more than two tets;" : "",
                (flags & MF_ORIENTATION) ? " inconsistent face orientation;" : "",
                (flags & MF_OPEN_FACE) ? " interior tet face without a neighbour;" : "");
}

static int build_device_mesh_impl(cpf_context *ctx, long long nVerts, const double *pos, long long nTets, const int *tetVerts,
                                  const int *tetCell, const int *tetPatch, long long nCells)
{
    cudaStream_t st = ctx->stream;
    // -- raw uploads
    DevBuf<double> d_xyz;
    DevBuf<int4> d_tetref;
    DevBuf<int> d_tetPatch;
    CPF_CUDA(ctx, d_xyz.alloc(3 * (size_t)nVerts));
    CPF_CUDA(ctx, d_tetref.alloc((size_t)nTets));
    CPF_CUDA(ctx, cudaMemcpyAsync(d_xyz, pos, sizeof(double) * 3 * (size_t)nVerts, cudaMemcpyHostToDevice, st));
    CPF_CUDA(ctx, cudaMemcpyAsync(d_tetref, tetVerts, sizeof(int4) * (size_t)nTets, cudaMemcpyHostToDevice, st));
    if (tetPatch) {
        CPF_CUDA(ctx, d_tetPatch.alloc((size_t)nTets));
        CPF_CUDA(ctx, cudaMemcpyAsync(d_tetPatch, tetPatch, sizeof(int) * (size_t)nTets, cudaMemcpyHostToDevice, st));
    }
    if (!ctx->cellFromVertex) {
        CPF_CUDA(ctx, cudaMalloc(&ctx->d_tetcell, sizeof(int) * (size_t)nTets));
        if (tetCell) CPF_CUDA(ctx, cudaMemcpyAsync(ctx->d_tetcell, tetCell, sizeof(int) * (size_t)nTets, cudaMemcpyHostToDevice, st));
        else {
            std::vector<int> iota((size_t)nTets);
            for (long long t = 0; t < nTets; ++t) iota[(size_t)t] = (int)t;
            CPF_CUDA(ctx, cudaMemcpy(ctx->d_tetcell, iota.data(), sizeof(int) * (size_t)nTets, cudaMemcpyHostToDevice));
        }
    }
    CPF_CUDA(ctx, cudaMalloc(&ctx->d_vpos, sizeof(double4) * (size_t)nVerts));
    CPF_CUDA(ctx, cudaMalloc(&ctx->d_tetv, sizeof(int4) * (size_t)nTets));
    CPF_CUDA(ctx, cudaMalloc(&ctx->d_tetrec, sizeof(int4) * 2 * (size_t)nTets));
    CPF_CUDA(ctx, cudaMalloc(&ctx->d_tetcode, sizeof(uint16_t) * (size_t)nTets));
    for (int b = 0; b < 2; ++b) {
        CPF_CUDA(ctx, cudaMalloc(&ctx->d_ucell[b], sizeof(double4) * (size_t)nCells));
        CPF_CUDA(ctx, cudaMemsetAsync(ctx->d_ucell[b], 0, sizeof(double4) * (size_t)nCells, st));
    }
    ctx->ucur = 0;
    k_pack_positions<<<(unsigned)((nVerts + 255) / 256), 256, 0, st>>>(nVerts, d_xyz, ctx->d_vpos);
    ctx->launches++;

    // -- per-tet preparation + face keys
    const long long nE = 4 * nTets;
    DevBuf<unsigned long long> d_keys, d_keys2, d_hmin, d_nb;
    DevBuf<int> d_idx, d_idx2;
    DevBuf<FaceEntry> d_entries;
    DevBuf<unsigned> d_flags;
    CPF_CUDA(ctx, d_keys.alloc((size_t)nE));
    CPF_CUDA(ctx, d_keys2.alloc((size_t)nE));
    CPF_CUDA(ctx, d_idx.alloc((size_t)nE));
    CPF_CUDA(ctx, d_idx2.alloc((size_t)nE));
    CPF_CUDA(ctx, d_entries.alloc((size_t)nE));
    CPF_CUDA(ctx, d_flags.alloc(1));
    CPF_CUDA(ctx, d_hmin.alloc(1));
    CPF_CUDA(ctx, d_nb.alloc(1));
    CPF_CUDA(ctx, cudaMemsetAsync(d_flags, 0, sizeof(unsigned), st));
    CPF_CUDA(ctx, cudaMemsetAsync(d_nb, 0, sizeof(unsigned long long), st));
    CPF_CUDA(ctx, cudaMemsetAsync(d_hmin, 0x7f, sizeof(unsigned long long), st)); // huge positive double
    int bits = 1;
    while ((1ll << bits) < nVerts) ++bits;
    k_prepare_tets<<<(unsigned)((nTets + 127) / 128), 128, 0, st>>>(nTets, (int)nVerts, d_tetref, ctx->d_vpos, ctx->d_tetv, ctx->d_tetcode,
                                                                   d_keys, d_idx, d_entries, d_flags, d_hmin, bits);
    ctx->launches++;
    CPF_CUDA(ctx, cudaGetLastError());
    unsigned flags = 0;
    CPF_CUDA(ctx, cudaMemcpyAsync(&flags, d_flags, sizeof flags, cudaMemcpyDeviceToHost, st));
    CPF_CUDA(ctx, cudaStreamSynchronize(st));
    if (flags) return mesh_error(ctx, flags); // nothing downstream may run on a malformed tet list

    // -- one radix sort on (smallest, second smallest) vertex id, 2*bits significant key bits
    size_t tmpBytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmpBytes, d_keys.p, d_keys2.p, d_idx.p, d_idx2.p, (long long)nE, 0, 2 * bits, st);
    DevBuf<unsigned char> d_tmp;
    CPF_CUDA(ctx, d_tmp.alloc(tmpBytes));
    CPF_CUDA(ctx, cub::DeviceRadixSort::SortPairs(d_tmp.p, tmpBytes, d_keys.p, d_keys2.p, d_idx.p, d_idx2.p, (long long)nE, 0, 2 * bits, st));
    ctx->launches += 4;

    k_link_faces<<<(unsigned)((nE + 127) / 128), 128, 0, st>>>(nE, d_keys2, d_idx2, d_entries, ctx->d_tetcode, d_tetPatch,
                                                              ctx->d_tetv, (int *)ctx->d_tetrec, d_flags, d_nb);
    ctx->launches++;
    CPF_CUDA(ctx, cudaGetLastError());

    CPF_CUDA(ctx, cudaMalloc(&ctx->d_tetnrm, sizeof(double) * 12 * (size_t)nTets));
    k_build_normals<<<(unsigned)((nTets + 127) / 128), 128, 0, st>>>(nTets, ctx->d_tetv, ctx->d_vpos, ctx->d_tetcode, (double *)ctx->d_tetnrm);
    ctx->launches++;
    CPF_CUDA(ctx, cudaMalloc(&ctx->d_tetfast, sizeof(uint4) * 4 * (size_t)nTets));
    k_build_fast<<<(unsigned)((nTets + 127) / 128), 128, 0, st>>>(nTets, ctx->d_tetv, ctx->d_vpos, ctx->d_tetrec, ctx->d_tetfast, ctx->cellFromVertex ? 1 : 0);
    ctx->launches++;
    CPF_CUDA(ctx, cudaGetLastError());

    unsigned long long hbits = 0, nb = 0;
    CPF_CUDA(ctx, cudaMemcpyAsync(&flags, d_flags, sizeof flags, cudaMemcpyDeviceToHost, st));
    CPF_CUDA(ctx, cudaMemcpyAsync(&hbits, d_hmin, sizeof hbits, cudaMemcpyDeviceToHost, st));
    CPF_CUDA(ctx, cudaMemcpyAsync(&nb, d_nb, sizeof nb, cudaMemcpyDeviceToHost, st));
    CPF_CUDA(ctx, cudaStreamSynchronize(st));
    if (flags) return mesh_error(ctx, flags);
    double hmin;
    memcpy(&hmin, &hbits, sizeof hmin);
    ctx->hmin = hmin;
    ctx->nBoundaryFaces = (long long)nb;
    // guard band of the filtered path in barycentric units: at least 100x the reference's absolute
    // 1e-13 tolerance measured against the smallest tet height, never below 1e-7
    const double g = 1e-11 / hmin;
    ctx->guard = g > 1e-7 ? g : 1e-7;
    // The filtered walk checks the start point of a particle once per launch and afterwards relies on "the end point C2
    // certified is the next start point" (cpf_geom.cuh visit_fast32): P + disp in fp64 moves that point by <= 2^-52 |P|,
    // which must stay inside the unused part of the error term (0.34 * 2^-18 E^3 against 12 E^2 |P| 2^-53): |P| < 9e8 E.
    // A mesh whose coordinates exceed 5e8 smallest tet heights (7 decimal digits left inside a tet) runs on the exact path.
    double amax = 0.0;
    for (int k = 0; k < 3; ++k) amax = std::max(amax, std::max(std::fabs(ctx->bbox_lo[k]), std::fabs(ctx->bbox_hi[k])));
    ctx->filter_ok = amax <= 5e8 * hmin;
    ctx->have_mesh = true;
    return build_bvh(ctx);
}

int build_device_mesh(cpf_context *ctx, long long nVerts, const double *pos, long long nTets, const int *tetVerts,
                      const int *tetCell, const int *tetPatch, long long nCells, int nPoints, bool cellFromVertex)
{
    if (nVerts <= 0 || nTets <= 0 || nCells <= 0) return fail(ctx, CPF_ERR_INVALID, "empty mesh");
    if (nTets >= (1ll << 29)) return fail(ctx, CPF_ERR_INVALID, "more than 2^29 tets per GPU are not supported");
    if (nVerts >= (1ll << 31)) return fail(ctx, CPF_ERR_INVALID, "more than 2^31 vertices are not supported");
    free_mesh(ctx);
    ctx->nVerts = nVerts; ctx->nTets = nTets; ctx->nCells = nCells; ctx->nPoints = nPoints;
    ctx->cellFromVertex = cellFromVertex;
    const int rc = build_device_mesh_impl(ctx, nVerts, pos, nTets, tetVerts, tetCell, tetPatch, nCells);
    if (rc != CPF_OK) { // leave the context without a mesh, never with a half-built one
        const std::string keep = ctx->err;
        cudaStreamSynchronize(ctx->stream);
        free_mesh(ctx);
        ctx->err = keep;
    }
    return rc;
}

MeshView mesh_view(const cpf_context *ctx)
{
    MeshView m;
    m.vpos = ctx->d_vpos; m.tetv = ctx->d_tetv; m.tetrec = ctx->d_tetrec; m.tetfast = ctx->d_tetfast; m.tetnrm = ctx->d_tetnrm; m.tetcode = ctx->d_tetcode;
    m.tetcell = ctx->cellFromVertex ? nullptr : ctx->d_tetcell;
    m.ucell = ctx->d_ucell[ctx->ucur];
    m.uvert = ctx->d_uvert;
    m.patch_kind = ctx->d_patch_kind;
    m.patch_gain = ctx->d_patch_gain;
    m.nPoints = ctx->nPoints; m.nTets = ctx->nTets; m.nCells = (int)ctx->nCells;
    m.guard = ctx->guard;
    m.guardf = (float)ctx->guard * 1.0000002f;
    return m;
}

void release_mesh(cpf_context *ctx) { free_mesh(ctx); }

} // namespace cpf
