/*
 * cpf_oracle.c -- CPU restatement of the cudaParticlesFoam / RTXAdvect particle hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (cudaparticlesfoam_b200/, include/, src/)
 * links, imports or calls this file.  It is used by tests/, by __graft_entry__.smoke() and by
 * bench.py's cpu_baseline / --impl reference legs as the *checker* and the CPU baseline.
 *
 * Parity status: PINNED against the reference's own CUDA kernels.  The three reference .cu files
 * compile unmodified for sm_100a (oracle/Makefile -> oracle/_ref/libref_rtxadvect.so); the
 * `-m gpu` tests run those kernels on the B200 and require this file to reproduce their output
 * bit for bit (tests/test_gpu_reference_pin.py), and tests/golden/ref_*.npz hold outputs of the
 * reference kernels captured on a B200 that the CPU-only suite replays through this file.
 * The reference itself ships no golden vectors or tests.
 *
 * Floating point: the reference is compiled by nvcc with default -fmad=true, so the rounding
 * sequence is fixed by where nvcc/ptxas place FMAs.  Every expression below spells that sequence
 * out with fma() and this file MUST be compiled with -ffp-contract=off.  The sequences were read
 * from the sm_100a SASS of the reference kernels (see DESIGN.md "FMA map"):
 *     cross(a,b).x = fma(a.y, b.z, -(b.y*a.z))            (left product fused, right rounded)
 *     dot(a,b)     = fma(a.z, b.z, fma(a.x, b.x, a.y*b.y))
 *     sqrt, div, rcp: IEEE round-to-nearest (sqrt.rn.f64 / div.rn.f64 / rcp.rn.f64)
 *
 * All file:line citations are relative to /root/reference/third_party/RTXAdvect/ unless noted.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#if defined(__FP_FAST_FMA) && 0
#error "unreachable"
#endif

typedef struct { double x, y, z; } v3;

#define ORC_TOL 1e-13 /* query/ConvexQuery.cu:42, :246 */

/* ------------------------------------------------------------------------------------------ */
/* owl/owl/include/owl/common/math/vec.h:317-345 with the nvcc FMA placement                   */
/* ------------------------------------------------------------------------------------------ */
static inline v3 v3_sub(v3 a, v3 b) { v3 r = { a.x - b.x, a.y - b.y, a.z - b.z }; return r; }
static inline v3 v3_add(v3 a, v3 b) { v3 r = { a.x + b.x, a.y + b.y, a.z + b.z }; return r; }
static inline v3 v3_neg(v3 a) { v3 r = { -a.x, -a.y, -a.z }; return r; }

static inline v3 ref_cross(v3 a, v3 b)
{
    v3 r;
    r.x = fma(a.y, b.z, -(b.y * a.z));
    r.y = fma(a.z, b.x, -(b.z * a.x));
    r.z = fma(a.x, b.y, -(b.x * a.y));
    return r;
}

static inline double ref_dot(v3 a, v3 b) { return fma(a.z, b.z, fma(a.x, b.x, a.y * b.y)); }

/* cuda/DeviceTetMesh.cuh:82-88 */
static inline double ref_det(v3 A, v3 B, v3 C, v3 D)
{
    return ref_dot(v3_sub(D, A), ref_cross(v3_sub(B, A), v3_sub(C, A)));
}

/* cuda/DeviceTetMesh.cuh:193-199: norm / length(norm), component-wise IEEE division */
static inline v3 ref_triNorm(v3 A, v3 B, v3 C)
{
    v3 n = ref_cross(v3_sub(B, A), v3_sub(C, A));
    double len = sqrt(ref_dot(n, n));
    v3 r = { n.x / len, n.y / len, n.z / len };
    return r;
}

static inline v3 ld3(const double *p, int i) { v3 r = { p[3 * i], p[3 * i + 1], p[3 * i + 2] }; return r; }

/* ------------------------------------------------------------------------------------------ */
/* Mesh view (reference layout: cuda/HostTetMesh.h:33-41, cuda/DeviceTetMesh.cuh:26-37)        */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    const double *pos;     /* [nVerts][3]  vec3d positions                        */
    const int *idx;        /* [nTets][4]   vec4i indices                          */
    const int *tetfacets;  /* [nTets][4]   face k is opposite vertex k            */
    const int *facets;     /* [nFaces][4]  ascending vertex triple, w = -1        */
    const int *finfo;      /* [nFaces][2]  {front, back}; negative => boundary    */
} orc_mesh;

/* ------------------------------------------------------------------------------------------ */
/* A2: face topology.  cuda/HostTetMesh.h:265-304 (add1Facet) and :307-430 (getBoundaryMesh).  */
/* Deviation (documented): the reference packs the sorted vertex triple in */
/* 3x20 bits (:279) and silently corrupts meshes with >= 2^20 vertices; this restatement keys   */
/* on the full 3x32-bit triple.  Face ids are handed out in first-appearance order like the     */
/* reference's std::map + push_back, so for < 2^20 vertices the tables are identical.           */
/* ------------------------------------------------------------------------------------------ */
typedef struct { int a, b, c, id; } face_slot;

static inline uint64_t face_hash(int a, int b, int c)
{
    uint64_t h = (uint64_t)(uint32_t)a * 0x9E3779B97F4A7C15ull;
    h ^= ((uint64_t)(uint32_t)b + 0x7F4A7C15ull) * 0xC2B2AE3D27D4EB4Full;
    h ^= h >> 29;
    h += (uint64_t)(uint32_t)c * 0x165667B19E3779F9ull;
    h ^= h >> 32;
    return h;
}

/* Returns number of faces, or -1 on allocation failure.  facets/finfo must hold 4*nTets entries
 * of 4 resp. 2 ints (upper bound).  tetfacets has nTets entries; tets skipped by the reference
 * (repeated vertex) are reported through *nSkipped (the reference then shifts every later entry,
 * Appendix A.2 -- callers must not feed such tets; we keep ids aligned and flag it).             */
long orc_build_faces(int nVerts, const double *pos, int nTets, const int *idx,
                     int *tetfacets, int *facets, int *finfo, int *nBoundaryOut, int *nSkipped)
{
    (void)nVerts;
    size_t cap = 1;
    while (cap < (size_t)nTets * 8u + 16u) cap <<= 1;
    face_slot *tab = (face_slot *)malloc(cap * sizeof(face_slot));
    if (!tab) return -1;
    for (size_t i = 0; i < cap; ++i) tab[i].id = -1;
    unsigned char *bmask = (unsigned char *)malloc((size_t)nTets * 4u + 4u);
    if (!bmask) { free(tab); return -1; }
    long nFaces = 0;
    int skipped = 0;

    for (int t = 0; t < nTets; ++t) {
        int ix = idx[4 * t], iy = idx[4 * t + 1], iz = idx[4 * t + 2], iw = idx[4 * t + 3];
        tetfacets[4 * t] = tetfacets[4 * t + 1] = tetfacets[4 * t + 2] = tetfacets[4 * t + 3] = -1;
        if (ix == iy || ix == iz || ix == iw || iy == iz || iy == iw || iz == iw) { skipped++; continue; }
        /* :329-343 volume sign test, evaluated in double then narrowed to float (host code,
         * no FMA contraction on the reference's x86-64 -O3 build) */
        v3 A = ld3(pos, ix), B = ld3(pos, iy), C = ld3(pos, iz), D = ld3(pos, iw);
        v3 e1 = v3_sub(B, A), e2 = v3_sub(C, A), e3 = v3_sub(D, A);
        v3 cr = { e1.y * e2.z - e2.y * e1.z, e1.z * e2.x - e2.z * e1.x, e1.x * e2.y - e2.x * e1.y };
        float volume = (float)(e3.x * cr.x + e3.y * cr.y + e3.z * cr.z);
        if (volume == 0.f) continue;
        if (volume < 0.f) { int tmp = ix; ix = iy; iy = tmp; }
        /* :351-358 Gmsh face order: face k is opposite vertex k */
        int fv[4][3] = { { iy, iz, iw }, { iz, ix, iw }, { ix, iy, iw }, { ix, iz, iy } };
        for (int k = 0; k < 4; ++k) {
            int a = fv[k][0], b = fv[k][1], c = fv[k][2];
            int front = 0; /* :272 */
            if (a > c) { int s = a; a = c; c = s; front = !front; }
            if (b > c) { int s = b; b = c; c = s; front = !front; }
            if (a > b) { int s = a; a = b; b = s; front = !front; }
            size_t h = (size_t)face_hash(a, b, c) & (cap - 1);
            int fid = -1;
            for (;;) {
                if (tab[h].id < 0) break;
                if (tab[h].a == a && tab[h].b == b && tab[h].c == c) { fid = tab[h].id; break; }
                h = (h + 1) & (cap - 1);
            }
            if (fid < 0) {
                fid = (int)nFaces++;
                tab[h].a = a; tab[h].b = b; tab[h].c = c; tab[h].id = fid;
                facets[4 * fid] = a; facets[4 * fid + 1] = b; facets[4 * fid + 2] = c; facets[4 * fid + 3] = -1;
                finfo[2 * fid] = -1; finfo[2 * fid + 1] = -1;
                bmask[fid] = 1;
            } else {
                bmask[fid] = 0;
            }
            tetfacets[4 * t + k] = fid;
            if (front) finfo[2 * fid] = t; else finfo[2 * fid + 1] = t;
        }
    }
    /* :394-411 boundary faces get -(bdCellID+1) on their empty side */
    int bd = 0;
    for (long f = 0; f < nFaces; ++f)
        if (bmask[f]) {
            if (finfo[2 * f] == -1) finfo[2 * f] = -(bd + 1); else finfo[2 * f + 1] = -(bd + 1);
            bd++;
        }
    if (nBoundaryOut) *nBoundaryOut = bd;
    if (nSkipped) *nSkipped = skipped;
    free(bmask);
    free(tab);
    return nFaces;
}

/* ------------------------------------------------------------------------------------------ */
/* A1: tet decomposition done by the glue, /root/reference/src/initCuda.H:86-110, relying on    */
/* OpenFOAM v2106 polyMeshTetDecomposition::cellTetIndices + tetIndices::faceTriIs (not         */
/* vendored; rule restated from SURVEY Appendix B "Tet numbering").  PARITY UNPINNED for this   */
/* function against real OpenFOAM (no OpenFOAM in this environment).                            */
/* Inputs are OpenFOAM polyMesh arrays: faces as CSR (faceOffsets/faceVerts), owner[nFaces],    */
/* neighbour[nInternal].  tetBasePt may be NULL (=> base point 0).                              */
/* Output: tets[nTets][4] = (nPoints+cell, F[b], pA, pB); tetCell[nTets]; returns nTets.        */
/* Pass tets == NULL to only count.                                                             */
/* ------------------------------------------------------------------------------------------ */
long orc_decompose_poly(int nPoints, int nCells, int nFaces, int nInternal,
                        const int *faceOffsets, const int *faceVerts,
                        const int *owner, const int *neighbour, const int *tetBasePt,
                        int *tets, int *tetCell)
{
    /* cell -> faces: owned faces ascending, then neighbour faces ascending (primitiveMesh::calcCells) */
    int *cnt = (int *)calloc((size_t)nCells + 1, sizeof(int));
    if (!cnt) return -1;
    for (int f = 0; f < nFaces; ++f) cnt[owner[f] + 1]++;
    for (int f = 0; f < nInternal; ++f) cnt[neighbour[f] + 1]++;
    for (int c = 0; c < nCells; ++c) cnt[c + 1] += cnt[c];
    int *cf = (int *)malloc((size_t)cnt[nCells] * sizeof(int));
    int *fill = (int *)calloc((size_t)nCells, sizeof(int));
    if (!cf || !fill) { free(cnt); free(cf); free(fill); return -1; }
    for (int f = 0; f < nFaces; ++f) { int c = owner[f]; cf[cnt[c] + fill[c]++] = f; }
    for (int f = 0; f < nInternal; ++f) { int c = neighbour[f]; cf[cnt[c] + fill[c]++] = f; }
    long nT = 0;
    for (int c = 0; c < nCells; ++c)
        for (int j = cnt[c]; j < cnt[c + 1]; ++j) {
            int f = cf[j];
            const int *F = faceVerts + faceOffsets[f];
            int n = faceOffsets[f + 1] - faceOffsets[f];
            int b = tetBasePt ? tetBasePt[f] : 0;
            for (int tetPt = 1; tetPt <= n - 2; ++tetPt) {
                int ia = (tetPt + b) % n;
                int ib = (ia + 1) % n; /* f.fcIndex(facePtI) */
                if (owner[f] != c) { int s = ia; ia = ib; ib = s; }
                if (tets) {
                    tets[4 * nT] = nPoints + c;
                    tets[4 * nT + 1] = F[b];
                    tets[4 * nT + 2] = F[ia];
                    tets[4 * nT + 3] = F[ib];
                    if (tetCell) tetCell[nT] = c;
                }
                nT++;
            }
        }
    free(cnt); free(cf); free(fill);
    return nT;
}

/* /root/reference/src/advect.H:44-57: every tet of a cell receives the cell value (the reference
 * hard-codes 12 tets per cell; tetCell generalises it, identical for hex meshes).               */
void orc_update_velocity(long nTets, const int *tetCell, const double *Ucell, double *Utet)
{
#pragma omp parallel for schedule(static)
    for (long t = 0; t < nTets; ++t) {
        int c = tetCell[t];
        Utet[3 * t] = Ucell[3 * c]; Utet[3 * t + 1] = Ucell[3 * c + 1]; Utet[3 * t + 2] = Ucell[3 * c + 2];
    }
}

/* ------------------------------------------------------------------------------------------ */
/* S1  cuda/particles.cu:316-373 particleAdvectKernelTetVel                                     */
/*     p[i] = (x,y,z,w) ; vel/disp are vec4d with .w = -1                                       */
/* ------------------------------------------------------------------------------------------ */
static inline void s1_advect_tetvel(double *p, int tet, double *vel, double *disp, double dt,
                                    const orc_mesh *m, const double *Utet)
{
    if (p[3] == 0.0) return;
    if (tet < 0) { p[3] = 0.0; return; }
    const int *ix = m->idx + 4 * tet;
    double den = ref_det(ld3(m->pos, ix[0]), ld3(m->pos, ix[1]), ld3(m->pos, ix[2]), ld3(m->pos, ix[3]));
    if (den == 0.0) { p[3] = 0.0; return; }
    v3 v = ld3(Utet, tet);
    /* P_next = P + dt*vel -> fma(dt, vel, P); P_disp = P_next - P */
    disp[0] = fma(dt, v.x, p[0]) - p[0];
    disp[1] = fma(dt, v.y, p[1]) - p[1];
    disp[2] = fma(dt, v.z, p[2]) - p[2];
    disp[3] = -1.0;
    vel[0] = v.x; vel[1] = v.y; vel[2] = v.z; vel[3] = -1.0;
}

/* cuda/particles.cu:244-313 particleAdvectKernel ("VertexVelocity": P1 / cellPoint-style
 * interpolation of per-vertex velocities; unreachable from the reference glue but part of the
 * library).  Uvert is [nVerts][3].  FMA placement follows the same nvcc rules.                   */
static inline void s1_advect_vertvel(double *p, int tet, double *vel, double *disp, double dt,
                                     const orc_mesh *m, const double *Uvert)
{
    if (p[3] == 0.0) return;
    if (tet < 0) { p[3] = 0.0; return; }
    const int *ix = m->idx + 4 * tet;
    v3 A = ld3(m->pos, ix[0]), B = ld3(m->pos, ix[1]), C = ld3(m->pos, ix[2]), D = ld3(m->pos, ix[3]);
    v3 P = { p[0], p[1], p[2] };
    double den = ref_det(A, B, C, D);
    if (den == 0.0) { p[3] = 0.0; return; }
    double rden = 1.0 / den;
    double wA = ref_det(P, B, C, D) * rden;
    double wB = ref_det(A, P, C, D) * rden;
    double wC = ref_det(A, B, P, D) * rden;
    /* det(A,B,C,P): nvcc reuses cross(B-A,C-A) from `den` and emits this dot with the x product
     * rounded first (PTX of particleAdvectKernel): fma(c.z,r.z, fma(c.y,r.y, c.x*r.x)) */
    v3 cr = ref_cross(v3_sub(B, A), v3_sub(C, A));
    v3 rr = v3_sub(P, A);
    double wD = fma(cr.z, rr.z, fma(cr.y, rr.y, cr.x * rr.x)) * rden;
    v3 uA = ld3(Uvert, ix[0]), uB = ld3(Uvert, ix[1]), uC = ld3(Uvert, ix[2]), uD = ld3(Uvert, ix[3]);
    /* wA*velA + wB*velB + wC*velC + wD*velD -> fma(wD,uD, fma(wC,uC, fma(wA,uA, wB*uB))) */
    v3 v;
    v.x = fma(wD, uD.x, fma(wC, uC.x, fma(wA, uA.x, wB * uB.x)));
    v.y = fma(wD, uD.y, fma(wC, uC.y, fma(wA, uA.y, wB * uB.y)));
    v.z = fma(wD, uD.z, fma(wC, uC.z, fma(wA, uA.z, wB * uB.z)));
    disp[0] = fma(dt, v.x, p[0]) - p[0];
    disp[1] = fma(dt, v.y, p[1]) - p[1];
    disp[2] = fma(dt, v.z, p[2]) - p[2];
    disp[3] = -1.0;
    vel[0] = v.x; vel[1] = v.y; vel[2] = v.z; vel[3] = -1.0;
}

/* ------------------------------------------------------------------------------------------ */
/* S2  cuda/particles.cu:551-575 particleBrownianMotion.                                        */
/*     The three normal deviates are INPUTS here (xi[3]); the reference draws them with cuRAND   */
/*     XORWOW + Box-Muller whose log/sincospi are CUDA libm (not reproducible on a CPU).         */
/*     disp += vec4d(xi,0)*randDisp -> disp.x = fma(xi0, randDisp, disp.x); disp.w += 0*randDisp */
/* ------------------------------------------------------------------------------------------ */
static inline void s2_brownian(const double *p, double *disp, const double *xi, double D, double dt)
{
    if (p[3] == 0.0) return;
    double randDisp = sqrt((2.00 * D) * dt);
    disp[0] = fma(xi[0], randDisp, disp[0]);
    disp[1] = fma(xi[1], randDisp, disp[1]);
    disp[2] = fma(xi[2], randDisp, disp[2]);
    disp[3] = fma(0.0, randDisp, disp[3]);
}

/* ------------------------------------------------------------------------------------------ */
/* query/ConvexQuery.cu:32-131 traceIntet                                                       */
/* ------------------------------------------------------------------------------------------ */
static inline int other_tet(const int *finfo, int f, int cur)
{
    /* :99-101 */
    int nxt = finfo[2 * f + 1];
    if (nxt == cur) nxt = finfo[2 * f];
    return nxt;
}

static inline v3 face_inward_normal(const orc_mesh *m, int f, int cur, v3 *Aout)
{
    const int *fc = m->facets + 4 * f;
    v3 A = ld3(m->pos, fc[0]), B = ld3(m->pos, fc[1]), C = ld3(m->pos, fc[2]);
    v3 n = ref_triNorm(A, B, C);             /* :77 */
    if (m->finfo[2 * f + 1] == cur) n = v3_neg(n); /* :78-79 */
    *Aout = A;
    return n;
}

static int trace_in_tet(v3 *P_start, v3 P_end, int cur, const orc_mesh *m, int *outletFace, int inletFace)
{
    const double tol = ORC_TOL;
    int next = cur;
    const v3 P0 = *P_start;
    const v3 d = v3_sub(P_end, P0);
    double dT_min = 1.1;
    for (int i = 0; i < 4; ++i) {
        const int f = m->tetfacets[4 * cur + i];
        v3 A;
        v3 n = face_inward_normal(m, f, cur, &A);
        double face_dist = ref_dot(v3_sub(A, P0), n);  /* :85 */
        double dT = face_dist / ref_dot(d, n);         /* :86 */
        if (isinf(dT)) dT = -1.0;                      /* :89 */
        if (f == inletFace) continue;                  /* :70-71, :94 */
        if (face_dist < tol && dT > tol && dT <= 1.0 && dT < dT_min) { /* :95 */
            dT_min = dT;
            next = other_tet(m->finfo, f, cur);
            P_start->x = fma(d.x, dT, P0.x);           /* :104 Pxf = P_0 + dT*P_disp */
            P_start->y = fma(d.y, dT, P0.y);
            P_start->z = fma(d.z, dT, P0.z);
            *outletFace = f;
        }
    }
    if (next < 0) next = -1; /* :128 */
    return next;
}

/* S3  query/ConvexQuery.cu:135-216 particleLocator */
static inline void s3_locate_convex(const double *p, const double *disp, int *tetIO, const orc_mesh *m)
{
    if (p[3] == 0.0) return;
    v3 P = { p[0], p[1], p[2] };
    v3 dd = { disp[0], disp[1], disp[2] };
    v3 P_end = v3_add(P, dd);
    v3 P_start = P;
    int cur = *tetIO, next = -2, OutFace = -2, InFace = -2;
    for (int i = 0; i < 50; ++i) {
        next = trace_in_tet(&P_start, P_end, cur, m, &OutFace, InFace);
        if (next == cur) break;
        InFace = OutFace;
        if (next == -1) break;
        cur = next;
    }
    if (next == -1) next = -(*tetIO + 1); /* :212 */
    *tetIO = next;
}

/* query/ConvexQuery.cu:239-317 reflectInTet.  If no face satisfies the exit test the source
 * reads uninitialised P_reflect/u_reflect; the compiled sm_100a code leaves P_end and vel
 * unchanged in that case (SASS of convexReflector), which is what we do.                        */
static void reflect_in_tet(v3 Pxf, v3 *P_end, v3 *vel, int tet, const orc_mesh *m)
{
    const double tol = ORC_TOL;
    const v3 d = v3_sub(*P_end, Pxf);
    for (int i = 0; i < 4; ++i) {
        const int f = m->tetfacets[4 * tet + i];
        v3 A;
        v3 n = face_inward_normal(m, f, tet, &A);
        double face_dist = ref_dot(v3_sub(A, Pxf), n);
        double dT = face_dist / ref_dot(d, n);
        if (isinf(dT)) dT = -1.0;
        if (fabs(dT) < tol) dT = tol;
        if (fabs(face_dist) < tol) face_dist = tol;
        if (dT == tol || face_dist == tol) {
            /* :287-295 with nw = -n.  SASS: s = -(fma(rz,nz, fma(ry,ny, rx*nx))); s2 = s+s;
             * P_end = fma(s2, n, P_end); same for the velocity.                                */
            v3 r = v3_sub(*P_end, A);
            double sp = -fma(r.z, n.z, fma(r.y, n.y, r.x * n.x));
            sp = sp + sp;
            double sv = -fma(vel->z, n.z, fma(vel->y, n.y, vel->x * n.x));
            sv = sv + sv;
            P_end->x = fma(sp, n.x, P_end->x);
            P_end->y = fma(sp, n.y, P_end->y);
            P_end->z = fma(sp, n.z, P_end->z);
            vel->x = fma(sv, n.x, vel->x);
            vel->y = fma(sv, n.y, vel->y);
            vel->z = fma(sv, n.z, vel->z);
            return;
        }
    }
}

/* S4  query/ConvexQuery.cu:320-436 convexReflector */
static inline void s4_reflect_convex(double *p, double *disp, double *vel, int *tetIO, const orc_mesh *m)
{
    if (p[3] == 0.0) return;
    int tetID = *tetIO;
    if (tetID >= 0) return;
    v3 P_start = { p[0], p[1], p[2] };
    v3 dd = { disp[0], disp[1], disp[2] };
    v3 P_end = v3_add(P_start, dd);
    v3 u = { vel[0], vel[1], vel[2] };
    int cur = -tetID - 1, next = -2, OutFace = -2, InFace = -2;
    v3 P_hit = { -1.0, -1.0, -1.0 };
    for (int j = 0; j < 5; ++j) {
        for (int i = 0; i < 50; ++i) {
            next = trace_in_tet(&P_start, P_end, cur, m, &OutFace, InFace);
            if (next == cur) break;
            InFace = OutFace;
            if (next == -1) break;
            cur = next;
        }
        if (next == cur && next != -1) break;
        P_hit = P_start;
        reflect_in_tet(P_hit, &P_end, &u, cur, m);
    }
    v3 nd = v3_sub(P_end, P_hit);
    p[0] = P_hit.x; p[1] = P_hit.y; p[2] = P_hit.z;
    disp[0] = nd.x; disp[1] = nd.y; disp[2] = nd.z;
    vel[0] = u.x; vel[1] = u.y; vel[2] = u.z;
    *tetIO = next;
}

/* ------------------------------------------------------------------------------------------ */
/* RTX=true build: cuda/DeviceTetMesh.cuh:108-156 tetBaryCoord, query/RTQuery.cu:35-90          */
/* ------------------------------------------------------------------------------------------ */
static inline void ref_tetBary(v3 P, v3 A, v3 B, v3 C, v3 D, double w[4])
{
    double den = ref_det(A, B, C, D);
    double r = 1.0 / den; /* rcp.rn.f64 */
    w[0] = ref_det(P, B, C, D) * r;
    w[1] = ref_det(A, P, C, D) * r;
    w[2] = ref_det(A, B, P, D) * r;
    w[3] = 1.0 - w[0] - w[1] - w[2];
}

/* CUDA min.f64 semantics: NaN operand is dropped */
static inline double cuda_fmin(double a, double b) { return fmin(a, b); }

static int bary_tet_search(v3 P, int tetStart, const orc_mesh *m, int *faceOut)
{
    int s = tetStart, faceID = -1, prev = s;
    for (int i = 0; i < 50; ++i) {
        const int *ix = m->idx + 4 * s;
        double w[4];
        ref_tetBary(P, ld3(m->pos, ix[0]), ld3(m->pos, ix[1]), ld3(m->pos, ix[2]), ld3(m->pos, ix[3]), w);
        double wmin = cuda_fmin(cuda_fmin(w[0], w[1]), cuda_fmin(w[2], w[3])); /* functors.h:293 */
        if (wmin >= 0.0) break;
        int k = 0; /* functors.h:319-325 arg_min, first minimum wins */
        for (int j = 1; j < 4; ++j) if (w[j] < w[k]) k = j;
        faceID = m->tetfacets[4 * s + k];
        prev = s;
        s = (m->finfo[2 * faceID] == s) ? m->finfo[2 * faceID + 1] : m->finfo[2 * faceID]; /* :62-65 */
        if (s < 0) { s = -(prev + 1); break; }
    }
    *faceOut = faceID;
    return s;
}

/* query/RTQuery.cu:189-218 baryQuery (narrow phase after the OptiX seeding) */
static inline void s0_bary_query(const double *p, int *tetIO, const orc_mesh *m)
{
    if (*tetIO < 0) return;
    v3 P = { p[0], p[1], p[2] };
    int f;
    *tetIO = bary_tet_search(P, *tetIO, m, &f);
}

/* S3' query/RTQuery.cu:221-248 baryQueryDisp (no w / sign checks, Appendix A.8: callers must not
 * pass negative ids; the oracle returns them unchanged instead of reading out of bounds).       */
static inline void s3_locate_bary(const double *p, const double *disp, int *tetIO, const orc_mesh *m)
{
    if (*tetIO < 0) return;
    v3 P = { p[0] + disp[0], p[1] + disp[1], p[2] + disp[2] };
    int f;
    *tetIO = bary_tet_search(P, *tetIO, m, &f);
}

/* query/RTQuery.cu:92-107 specularReflect.  norm is flipped when back == tet; the formula is
 * insensitive to the sign.  Source: P - (1+1)*dot(P-A,norm)*norm.                               */
static inline void specular_reflect(v3 *P, v3 *u, int tet, int f, const orc_mesh *m)
{
    v3 A;
    v3 n = face_inward_normal(m, f, tet, &A);
    v3 r = v3_sub(*P, A);
    double sp = ref_dot(r, n);
    sp = sp + sp;
    double sv = ref_dot(*u, n);
    sv = sv + sv;
    /* P - (2*dot)*n -> fma(-(2dot), n, P) */
    P->x = fma(-sp, n.x, P->x); P->y = fma(-sp, n.y, P->y); P->z = fma(-sp, n.z, P->z);
    u->x = fma(-sv, n.x, u->x); u->y = fma(-sv, n.y, u->y); u->z = fma(-sv, n.z, u->z);
}

/* S4' query/RTQuery.cu:109-186 RTreflection */
static inline void s4_reflect_bary(const double *p, double *disp, double *vel, int *tetIO, const orc_mesh *m)
{
    int tetID = *tetIO;
    if (tetID >= 0) return;
    tetID = -(tetID + 1);
    v3 P = { p[0], p[1], p[2] };
    v3 R = { p[0] + disp[0], p[1] + disp[1], p[2] + disp[2] };
    v3 u = { vel[0], vel[1], vel[2] };
    int bd = tetID;
    for (int i = 0; i < 10; ++i) {
        int f;
        int s = bary_tet_search(R, bd, m, &f);
        if (s >= 0) { bd = s; break; }
        bd = -(s + 1);
        specular_reflect(&R, &u, bd, f, m);
    }
    v3 nd = v3_sub(R, P);
    disp[0] = nd.x; disp[1] = nd.y; disp[2] = nd.z;
    vel[0] = u.x; vel[1] = u.y; vel[2] = u.z;
    *tetIO = bd;
}

/* S5  cuda/particles.cu:659-704 particleMoveKernel(disps) */
static inline void s5_move(double *p, double *disp)
{
    if (p[3] == 0.0) return;
    p[0] += disp[0]; p[1] += disp[1]; p[2] += disp[2];
    disp[0] = 0.0; disp[1] = 0.0; disp[2] = 0.0;
}

/* ------------------------------------------------------------------------------------------ */
/* Flat C entry points (ctypes).  One call == one reference host function over all particles.   */
/* ------------------------------------------------------------------------------------------ */
#define MESH_ARGS const double *pos, const int *idx, const int *tetfacets, const int *facets, const int *finfo
#define MESH_INIT orc_mesh m = { pos, idx, tetfacets, facets, finfo }

void orc_advect(long n, double *p, const int *tet, double *vel, double *disp, double dt,
                MESH_ARGS, const double *U, int vertexVelocity)
{
    MESH_INIT;
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) {
        if (vertexVelocity) s1_advect_vertvel(p + 4 * i, tet[i], vel + 4 * i, disp + 4 * i, dt, &m, U);
        else s1_advect_tetvel(p + 4 * i, tet[i], vel + 4 * i, disp + 4 * i, dt, &m, U);
    }
}

void orc_brownian(long n, const double *p, double *disp, const double *xi, double D, double dt)
{
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) s2_brownian(p + 4 * i, disp + 4 * i, xi + 3 * i, D, dt);
}

void orc_locate_convex(long n, const double *p, const double *disp, int *tet, MESH_ARGS)
{
    MESH_INIT;
#pragma omp parallel for schedule(dynamic, 1024)
    for (long i = 0; i < n; ++i) s3_locate_convex(p + 4 * i, disp + 4 * i, tet + i, &m);
}

void orc_reflect_convex(long n, double *p, double *disp, double *vel, int *tet, MESH_ARGS)
{
    MESH_INIT;
#pragma omp parallel for schedule(dynamic, 1024)
    for (long i = 0; i < n; ++i) s4_reflect_convex(p + 4 * i, disp + 4 * i, vel + 4 * i, tet + i, &m);
}

void orc_bary_query(long n, const double *p, int *tet, MESH_ARGS)
{
    MESH_INIT;
#pragma omp parallel for schedule(dynamic, 1024)
    for (long i = 0; i < n; ++i) s0_bary_query(p + 4 * i, tet + i, &m);
}

void orc_locate_bary(long n, const double *p, const double *disp, int *tet, MESH_ARGS)
{
    MESH_INIT;
#pragma omp parallel for schedule(dynamic, 1024)
    for (long i = 0; i < n; ++i) s3_locate_bary(p + 4 * i, disp + 4 * i, tet + i, &m);
}

void orc_reflect_bary(long n, const double *p, double *disp, double *vel, int *tet, MESH_ARGS)
{
    MESH_INIT;
#pragma omp parallel for schedule(dynamic, 1024)
    for (long i = 0; i < n; ++i) s4_reflect_bary(p + 4 * i, disp + 4 * i, vel + 4 * i, tet + i, &m);
}

void orc_move(long n, double *p, double *disp)
{
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) s5_move(p + 4 * i, disp + 4 * i);
}

/* The sub-step loop of /root/reference/src/advect.H:86-184 (default ConvexPoly build when
 * convex != 0, RTX=true build otherwise), fused per particle: particles never interact, so
 * running S1..S5 particle-by-particle is the same computation as kernel-by-kernel.
 * xi: NULL (no Brownian call at all) or [nSteps][n][3] normal deviates.                         */
void orc_substeps(long n, int nSteps, double *p, int *tet, double *vel, double *disp, double dt,
                  MESH_ARGS, const double *U, int vertexVelocity, int convex, int reflectWall,
                  const double *xi, double D)
{
    MESH_INIT;
#pragma omp parallel for schedule(dynamic, 256)
    for (long i = 0; i < n; ++i) {
        double *pi = p + 4 * i, *vi = vel + 4 * i, *di = disp + 4 * i;
        int *ti = tet + i;
        for (int s = 0; s < nSteps; ++s) {
            if (vertexVelocity) s1_advect_vertvel(pi, *ti, vi, di, dt, &m, U);
            else s1_advect_tetvel(pi, *ti, vi, di, dt, &m, U);
            if (xi) s2_brownian(pi, di, xi + ((size_t)s * (size_t)n + (size_t)i) * 3u, D, dt);
            if (convex) {
                s3_locate_convex(pi, di, ti, &m);
                if (reflectWall) s4_reflect_convex(pi, di, vi, ti, &m);
            } else {
                s3_locate_bary(pi, di, ti, &m);
                if (reflectWall) s4_reflect_bary(pi, di, vi, ti, &m);
            }
            s5_move(pi, di);
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* A4: initial location.  The reference seeds with an fp32 OptiX ray cast (optix/               */
/* optixQueryKernel.cu:63-124, closed-source traversal, not restatable) and then corrects with  */
/* baryQuery (query/RTQuery.cu:295-310).  Parity for seeding is defined as "same    */
/* tet as brute-force fp64 containment (lowest tet id with all bary >= 0), then baryQuery".     */
/* ------------------------------------------------------------------------------------------ */
void orc_locate_brute(long n, const double *p, int *tet, long nTets, MESH_ARGS)
{
    MESH_INIT;
#pragma omp parallel for schedule(dynamic, 16)
    for (long i = 0; i < n; ++i) {
        v3 P = { p[4 * i], p[4 * i + 1], p[4 * i + 2] };
        int found = -1;
        for (long t = 0; t < nTets && found < 0; ++t) {
            const int *ix = idx + 4 * t;
            v3 A = ld3(pos, ix[0]);
            /* cheap reject on the bounding box before the exact test */
            v3 B = ld3(pos, ix[1]), C = ld3(pos, ix[2]), D = ld3(pos, ix[3]);
            double lox = fmin(fmin(A.x, B.x), fmin(C.x, D.x)), hix = fmax(fmax(A.x, B.x), fmax(C.x, D.x));
            if (P.x < lox || P.x > hix) continue;
            double loy = fmin(fmin(A.y, B.y), fmin(C.y, D.y)), hiy = fmax(fmax(A.y, B.y), fmax(C.y, D.y));
            if (P.y < loy || P.y > hiy) continue;
            double loz = fmin(fmin(A.z, B.z), fmin(C.z, D.z)), hiz = fmax(fmax(A.z, B.z), fmax(C.z, D.z));
            if (P.z < loz || P.z > hiz) continue;
            double w[4];
            ref_tetBary(P, A, B, C, D, w);
            if (w[0] >= 0.0 && w[1] >= 0.0 && w[2] >= 0.0 && w[3] >= 0.0) found = (int)t;
        }
        tet[i] = found;
        if (found >= 0) s0_bary_query(p + 4 * i, tet + i, &m);
    }
}

/* Barycentric coordinates of each particle in its tet (diagnostic used by the tests to
 * recognise documented on-face ties).                                                         */
void orc_bary_of(long n, const double *p, const int *tet, const double *pos, const int *idx, double *w)
{
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) {
        if (tet[i] < 0) { w[4 * i] = w[4 * i + 1] = w[4 * i + 2] = w[4 * i + 3] = NAN; continue; }
        const int *ix = idx + 4 * tet[i];
        v3 P = { p[4 * i], p[4 * i + 1], p[4 * i + 2] };
        ref_tetBary(P, ld3(pos, ix[0]), ld3(pos, ix[1]), ld3(pos, ix[2]), ld3(pos, ix[3]), w + 4 * i);
    }
}

int orc_abi_version(void) { return 1; }

#include "cpf_oracle_ext.c"
#include "cpf_foamtrack.c"
#include "cpf_filter_model.c"
