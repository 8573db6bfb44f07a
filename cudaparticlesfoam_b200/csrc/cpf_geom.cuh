// cpf_geom.cuh -- device geometry for the particle hot path (sm_100a).
//
// Two arithmetic policies live here:
//  * EXACT: the reference's expression trees with the FMA placement nvcc/ptxas gives the reference
//    kernels (read from their sm_100a SASS; DESIGN.md "FMA map").  Written with explicit
//    round-to-nearest intrinsics so that no compiler decision can change a bit.  Follows
//    third_party/RTXAdvect/cuda/DeviceTetMesh.cuh:82-156,193-199, query/ConvexQuery.cu:32-131,
//    239-317, query/RTQuery.cu:35-107 and owl/common/math/vec.h:317-345 of the reference.
//  * FILTERED: cheap un-normalised fp32 predicates with a rigorous error bound and guard band;
//    whenever a decision of the reference could depend on rounding (or a wall is touched) the walk
//    reports NEED_EXACT and the sub-step is redone with the EXACT policy, so results stay
//    bit-identical to the reference.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cpf {

struct D3 { double x, y, z; };

#define CPF_DEV __device__ __forceinline__
#define CPF_TOL 1e-13 /* query/ConvexQuery.cu:42 */

// ------------------------------------------------------------------------------------------------
// memory access helpers
// ------------------------------------------------------------------------------------------------
// mesh data: read-only path, allowed to live in L1/L2
CPF_DEV D3 ld_vertex(const double4 *__restrict__ vpos, int v)
{
    double x, y, z, w;
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(x), "=d"(y), "=d"(z), "=d"(w) : "l"(vpos + v));
    (void)w;
    return D3{ x, y, z };
}
CPF_DEV int4 ld_int4(const int4 *__restrict__ p, int i) { return __ldg(p + i); }

// particle state: streamed once per launch -> evict-first, one 256-bit transaction per particle
CPF_DEV double4 ld_stream4(const double4 *p)
{
    double4 r;
    asm volatile("ld.global.cs.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
    return r;
}
CPF_DEV void st_stream4(double4 *p, double4 v)
{
    asm volatile("st.global.cs.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
}
CPF_DEV int ld_stream_i(const int *p)
{
    int r;
    asm volatile("ld.global.cs.s32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
// L1 prefetch as ONE predicated instruction: it stays in the basic block it is written in, and -- unlike a load -- it
// holds no dependency barrier, so the reconvergence points between it and the use of the line do not wait for it
CPF_DEV void prefetch_l1_if(const void *p, bool on)
{
    asm volatile("{ .reg .pred q; setp.ne.s32 q, %1, 0; @q prefetch.global.L1 [%0]; }" ::"l"(p), "r"((int)on));
}
CPF_DEV void st_stream_i(int *p, int v) { asm volatile("st.global.cs.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// ------------------------------------------------------------------------------------------------
// EXACT policy primitives
// ------------------------------------------------------------------------------------------------
CPF_DEV D3 xsub(D3 a, D3 b) { return D3{ __dsub_rn(a.x, b.x), __dsub_rn(a.y, b.y), __dsub_rn(a.z, b.z) }; }
CPF_DEV D3 xadd(D3 a, D3 b) { return D3{ __dadd_rn(a.x, b.x), __dadd_rn(a.y, b.y), __dadd_rn(a.z, b.z) }; }
CPF_DEV D3 xneg(D3 a) { return D3{ -a.x, -a.y, -a.z }; }
// cross(a,b).x = a.y*b.z - b.y*a.z  ->  fma(a.y, b.z, -(b.y*a.z))
CPF_DEV D3 xcross(D3 a, D3 b)
{
    return D3{ __fma_rn(a.y, b.z, -__dmul_rn(b.y, a.z)), __fma_rn(a.z, b.x, -__dmul_rn(b.z, a.x)),
               __fma_rn(a.x, b.y, -__dmul_rn(b.x, a.y)) };
}
// dot(a,b) = a.x*b.x + a.y*b.y + a.z*b.z  ->  fma(a.z,b.z, fma(a.x,b.x, a.y*b.y))
CPF_DEV double xdot(D3 a, D3 b) { return __fma_rn(a.z, b.z, __fma_rn(a.x, b.x, __dmul_rn(a.y, b.y))); }
CPF_DEV double xdet(D3 A, D3 B, D3 C, D3 D) { return xdot(xsub(D, A), xcross(xsub(B, A), xsub(C, A))); }
CPF_DEV D3 xtriNorm(D3 A, D3 B, D3 C)
{
    D3 n = xcross(xsub(B, A), xsub(C, A));
    double len = __dsqrt_rn(xdot(n, n));
    return D3{ __ddiv_rn(n.x, len), __ddiv_rn(n.y, len), __ddiv_rn(n.z, len) };
}

// ------------------------------------------------------------------------------------------------
// Device tet record.  Vertices are stored SORTED BY VERTEX ID (s0<s1<s2<s3): the reference's
// faces are ascending vertex triples (cuda/HostTetMesh.h:273-275), so face j (opposite s_j) is
// simply (s_a,s_b,s_c) with j removed.  `code` keeps what sorting would otherwise lose:
//   bits 0..7  perm[k] (2 bits each): sorted slot of the reference's k-th tet vertex, i.e. the
//              reference's face slot k (tetfacets[t][k], opposite vertex k) is sorted face perm[k]
//   bits 8..11 flip[j]: the reference negates the sorted-triple normal on this side
//              (faceinfos[f].back == t, query/ConvexQuery.cu:78)
// link[j] >= 0: (neighbour tet << 2) | (sorted face slot inside the neighbour); < 0: -(patch+1).
// ------------------------------------------------------------------------------------------------
struct Tet {
    D3 P[4];
    D3 N[4]; // inward unit normals of the sorted faces, exactly as traceIntet would compute them
    int4 link;
    unsigned code;
};

struct MeshView {
    const double4 *__restrict__ vpos; // [nVerts]
    const int4 *__restrict__ tetv;    // [nTets] sorted vertex ids
    const int4 *__restrict__ tetrec;  // [nTets][2] {links, apex ids} (32 B, one 256-bit load)
    const uint4 *__restrict__ tetfast; // [nTets][4] 64-byte fp32 record of the fast walk (see Fast32)
    const double4 *__restrict__ tetnrm; // [nTets][3] = 12 doubles: the four reference face normals (exact path)
    const uint16_t *__restrict__ tetcode;
    const int *__restrict__ tetcell;  // [nTets] or nullptr (cell = max vertex id - nPoints)
    const double4 *__restrict__ ucell; // [nCells] (ux,uy,uz,0): one 256-bit load per velocity fetch
    const double *__restrict__ uvert; // [nVerts][3] (CPF_INTERP_VERTEX)
    const uint8_t *__restrict__ patch_kind; // [nPatches]
    const double *__restrict__ patch_gain;  // [nPatches] 1 + restitution coefficient of the patch (2 = specular, the reference)
    int nPoints;
    long long nTets;
    int nCells;
    double guard; // barycentric guard band of the filtered path
    float guardf; // the same, rounded up to fp32
};

CPF_DEV D3 ld_ucell(const MeshView &m, int cell) { return ld_vertex(m.ucell, cell); }
CPF_DEV int link_at(int4 l, int j) { return j == 0 ? l.x : (j == 1 ? l.y : (j == 2 ? l.z : l.w)); }
CPF_DEV int idx_at(int4 l, int j) { return link_at(l, j); }

CPF_DEV Tet load_tet(const MeshView &m, int t, int4 &vout)
{
    Tet T;
    int4 v = ld_int4(m.tetv, t);
    T.link = ld_int4(m.tetrec, 2 * t);
    T.code = m.tetcode[t];
    T.P[0] = ld_vertex(m.vpos, v.x);
    T.P[1] = ld_vertex(m.vpos, v.y);
    T.P[2] = ld_vertex(m.vpos, v.z);
    T.P[3] = ld_vertex(m.vpos, v.w);
    {   // 96 bytes, three 256-bit loads
        double q[12];
        const double4 *np = m.tetnrm + 3ll * t;
#pragma unroll
        for (int k = 0; k < 3; ++k)
            asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(q[4 * k]), "=d"(q[4 * k + 1]), "=d"(q[4 * k + 2]), "=d"(q[4 * k + 3]) : "l"(np + k));
#pragma unroll
        for (int j = 0; j < 4; ++j) T.N[j] = D3{ q[3 * j], q[3 * j + 1], q[3 * j + 2] };
    }
    vout = v;
    return T;
}

CPF_DEV int tet_cell(const MeshView &m, int t, int4 v) { return m.tetcell ? __ldg(m.tetcell + t) : (v.w - m.nPoints); }

// sorted face j = sorted vertices without slot j
CPF_DEV void face_abc(const Tet &T, int j, D3 &A, D3 &B, D3 &C)
{
    A = (j == 0) ? T.P[1] : T.P[0];
    B = (j <= 1) ? T.P[2] : T.P[1];
    C = (j <= 2) ? T.P[3] : T.P[2];
}

// inward unit normal of sorted face j exactly as traceIntet builds it (ConvexQuery.cu:73-79):
// normalize(cross(B-A, C-A)) of the ascending triple, negated when the tet is the face's `back`.
// It is a pure function of the three vertex positions, so it is evaluated once at mesh build
// (face_normal_build, same expression tree) and read back here: bit-identical, and the exact path
// loses its sqrt + 3 divisions per face.
CPF_DEV D3 face_normal_build(const Tet &T, int j)
{
    D3 A, B, C;
    face_abc(T, j, A, B, C);
    D3 n = xtriNorm(A, B, C);
    if ((T.code >> (8 + j)) & 1u) n = xneg(n);
    return n;
}
CPF_DEV D3 face_normal_exact(const Tet &T, int j, D3 &A)
{
    A = (j == 0) ? T.P[1] : T.P[0];
    return j == 0 ? T.N[0] : (j == 1 ? T.N[1] : (j == 2 ? T.N[2] : T.N[3]));
}

// query/ConvexQuery.cu:32-131 traceIntet.  Returns the sorted face slot the segment leaves through
// (-1: the end point is inside this tet); S is advanced to the exit point.
CPF_DEV int trace_exact(const Tet &T, D3 &S, D3 E, int in_j)
{
    const D3 P0 = S;
    const D3 d = xsub(E, P0);
    double best = 1.1;
    int out_j = -1;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int j = (T.code >> (2 * k)) & 3u;
        if (j == in_j) continue; // inlet face: computed but never accepted by the reference
        D3 A;
        D3 n = face_normal_exact(T, j, A);
        const double fd = xdot(xsub(A, P0), n);
        const double den = xdot(d, n);
        // dT = fd/den is accepted only if fd < tol and tol < dT <= 1.  RN(fd/den) is positive only when fd and
        // den have the same sign, and <= 1 exactly when |fd| <= |den| (distinct doubles differ by >= 2^-53
        // relative, beyond the rounding midpoint), so the fp64 division is needed for those faces only;
        // zero/infinite/NaN quotients (-1 in the reference) are rejected by the same tests.
        if (!(fd < CPF_TOL) || !(fabs(fd) <= fabs(den)) || !((fd < 0.0) == (den < 0.0)) || fd == 0.0) continue;
        const double dT = __ddiv_rn(fd, den);
        if (dT > CPF_TOL && dT < best) {
            best = dT;
            out_j = j;
            S.x = __fma_rn(d.x, dT, P0.x);
            S.y = __fma_rn(d.y, dT, P0.y);
            S.z = __fma_rn(d.z, dT, P0.z);
        }
    }
    return out_j;
}

// query/ConvexQuery.cu:239-317 reflectInTet.  When no face matches, the source reads uninitialised
// P_reflect/u_reflect; the compiled reference leaves E and u untouched, as here.
// Rebound model (extension): the mirrored part is scaled by the restitution coefficient e of the patch the
// matching face lies on, sp = -(1 + e) (E - A).n; e = 1 (gain 2) is the reference's specular reflection bit for bit
// (2 s == s + s).  Faces that are not boundary faces (the reference may match one in degenerate cases) reflect specularly.
CPF_DEV double face_gain(const MeshView &m, const Tet &T, int j)
{
    const int link = link_at(T.link, j);
    return link < 0 ? __ldg(m.patch_gain - link - 1) : 2.0;
}

CPF_DEV void reflect_exact(const MeshView &m, const Tet &T, D3 Pxf, D3 &E, D3 &u)
{
    const D3 d = xsub(E, Pxf);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int j = (T.code >> (2 * k)) & 3u;
        D3 A;
        D3 n = face_normal_exact(T, j, A);
        double fd = xdot(xsub(A, Pxf), n);
        double dT = __ddiv_rn(fd, xdot(d, n));
        if (isinf(dT)) dT = -1.0;
        if (fabs(dT) < CPF_TOL) dT = CPF_TOL;
        if (fabs(fd) < CPF_TOL) fd = CPF_TOL;
        if (dT == CPF_TOL || fd == CPF_TOL) {
            const double gain = face_gain(m, T, j);
            D3 r = xsub(E, A);
            double sp = -__fma_rn(r.z, n.z, __fma_rn(r.y, n.y, __dmul_rn(r.x, n.x)));
            sp = __dmul_rn(gain, sp);
            double sv = -__fma_rn(u.z, n.z, __fma_rn(u.y, n.y, __dmul_rn(u.x, n.x)));
            sv = __dmul_rn(gain, sv);
            E = D3{ __fma_rn(sp, n.x, E.x), __fma_rn(sp, n.y, E.y), __fma_rn(sp, n.z, E.z) };
            u = D3{ __fma_rn(sv, n.x, u.x), __fma_rn(sv, n.y, u.y), __fma_rn(sv, n.z, u.z) };
            return;
        }
    }
}

// cuda/DeviceTetMesh.cuh:108-156 tetBaryCoord in the reference's vertex order
CPF_DEV void bary_exact(const Tet &T, D3 P, double w[4])
{
    const D3 A = T.P[(T.code >> 0) & 3u], B = T.P[(T.code >> 2) & 3u], C = T.P[(T.code >> 4) & 3u],
             D = T.P[(T.code >> 6) & 3u];
    const double r = __drcp_rn(xdet(A, B, C, D));
    w[0] = __dmul_rn(xdet(P, B, C, D), r);
    w[1] = __dmul_rn(xdet(A, P, C, D), r);
    w[2] = __dmul_rn(xdet(A, B, P, D), r);
    w[3] = __dsub_rn(__dsub_rn(__dsub_rn(1.0, w[0]), w[1]), w[2]);
}

#define CPF_NEED_EXACT (-1)

CPF_DEV int sel4(int a, int b, int c, int d, int k)
{
    const int lo = (k & 1) ? b : a, hi = (k & 1) ? d : c;
    return (k & 2) ? hi : lo;
}

// ------------------------------------------------------------------------------------------------
// fp32 fast walk (k_fast).
//
// One self-contained 64-byte record per tet: links, the un-normalised inward normals of the three
// faces through the tet's highest-id vertex (the "origin": for OpenFOAM decompositions the cell
// centre, so all 12 tets of a cell share it) -- computed in fp64 at build time, rounded once --,
// an origin id (below), 6*volume and the largest |vertex offset| E (its sign bit flags records whose
// slots 1 and 2 were exchanged to make the orientation positive).  A hop is ONE 64-byte load;
// the fp64 origin position is fetched only when the walk enters another cell.  In a cell-centre
// decomposition the origin changes exactly when the walk leaves through stored slot 3 (the face opposite
// the centre), so the record carries the origin id of the tet BEHIND that face: the new origin's position
// is requested together with the next record instead of after it (one dependent memory latency less per
// cell change); the walk tracks its current origin id itself (WalkF::org, seeded from tetv[tet].w).  All predicates run
// on the fp32 pipe.  Soundness: every comparison is made against g = G*|V6| + ERR, where
// ERR = 2^-18 * E^2 * (E + 3(R+D)), R = |r|inf, D = |d|inf, bounds the distance between a computed plane
// function and its real value.  With u = 2^-24, |N_c| <= 2E^2, V6 <= 6E^3, one rounding on each of N, r, d, V6:
//   a_j, j<3  (FMUL + 2 FFMA):   inputs 3*4u*R*E^2, operations u*(2+4+6)*R*E^2          -> 24u R E^2
//   a_3 = V6 - a_0 - a_1 - a_2:  inputs 6u E^3 + 72u R E^2, operations u*(18 E^3 + 36 R E^2) -> u(24 E^3 + 108 R E^2)
//   b_j likewise with D (b_3 <= 102u D E^2);  e_j = a_j + b_j, fma(t, b_j, a_j) with 0 <= t <= 1: one more rounding
//   worst case (j = 3): u * E^2 * (30 E + 126 (R + D))  <=  0.66 * 2^-18 * E^2 * (E + 3(R+D)).
// A wrongly chosen exit is caught by C3.  Whatever fails a test is deferred to the exact kernel, so results
// stay bit-identical.
// ------------------------------------------------------------------------------------------------
struct Fast32 {
    int4 link;
    float N[3][3];
    int aux; // cell-centre decompositions (MeshView::tetcell == nullptr): origin id of the tet BEHIND stored slot 3; else: own origin id
    float V6, E;
};

CPF_DEV void f32_load(const MeshView &m, int tet, Fast32 &f)
{
    unsigned w[16];
    const uint4 *p = m.tetfast + 4ll * tet;
    asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(p));
    asm("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(w[8]), "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15]) : "l"(p + 2));
    f.link = make_int4((int)w[0], (int)w[1], (int)w[2], (int)w[3]);
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int c = 0; c < 3; ++c) f.N[k][c] = __uint_as_float(w[4 + 3 * k + c]);
    f.aux = (int)w[13];
    f.V6 = __uint_as_float(w[14]);
    f.E = __uint_as_float(w[15]);
}

// the same, pinned where it is written (volatile): the speculative request of the next record must not sink below the
// checks that follow it
CPF_DEV void f32_load_pinned(const MeshView &m, int tet, Fast32 &f)
{
    unsigned w[16];
    const uint4 *p = m.tetfast + 4ll * tet;
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(p));
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(w[8]), "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15]) : "l"(p + 2));
    f.link = make_int4((int)w[0], (int)w[1], (int)w[2], (int)w[3]);
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int c = 0; c < 3; ++c) f.N[k][c] = __uint_as_float(w[4 + 3 * k + c]);
    f.aux = (int)w[13];
    f.V6 = __uint_as_float(w[14]);
    f.E = __uint_as_float(w[15]);
}

CPF_DEV float rcp_ftz(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// Packed fp32 pairs (sm_100a FFMA2 / FADD2 / FMUL2): the plane functions come in pairs (a_j, b_j) = N_j . (r, d), so
// one packed instruction serves the start-point and the direction term of a face; ptxas folds the {n, n} operand
// into a scalar broadcast (R.F32).  Each half is an IEEE round-to-nearest operation like its scalar form.
typedef unsigned long long F2;
CPF_DEV F2 f2_pack(float lo, float hi) { F2 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
CPF_DEV void f2_unpack(F2 v, float &lo, float &hi) { asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
CPF_DEV F2 f2_mul(float n, F2 x) { F2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(f2_pack(n, n)), "l"(x)); return r; }
CPF_DEV F2 f2_fma(float n, F2 x, F2 c) { F2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(f2_pack(n, n)), "l"(x), "l"(c)); return r; }
CPF_DEV F2 f2_sub(F2 a, F2 b) { F2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

// ------------------------------------------------------------------------------------------------
// The walk, one tet visit at a time.  Lanes of a warp need different numbers of tet visits per
// sub-step; k_fast runs the visits of ALL fused sub-steps of a lane through one loop, which keeps
// the lanes busy until their whole chunk is done instead of idling at every sub-step boundary.
// Records are stored positively oriented (V6 > 0: slots 1 and 2 are exchanged at build time where
// needed, links carry the STORED slot of the entry face), so "inside" is a_j >= 0 without a sign.
// ------------------------------------------------------------------------------------------------
struct WalkF {
    float rx, ry, rz, dx, dy, dz, RD3, Dd, t_in;
    int cur;
    int org;                // origin vertex id of the tet `cur` (the record only names it for generic tet meshes)
    int wall_js, wall_link; // set with CPF_V_WALL
    unsigned path;          // stored exit slot of every hop of this leg, 2 bits each (wall handling)
};
// CPF_V_WALL: every check passed and the certified exit face is a boundary face (callers without wall
// handling treat it like a refusal: oc >= CPF_V_REFUSE)
enum { CPF_V_DONE = 0, CPF_V_HOP = 1, CPF_V_REFUSE = 2, CPF_V_WALL = 3 };
// how a kernel learns that the mesh is a cell-centre decomposition: compile time (k_lean) or from the mesh view
enum { CPF_CFV_NO = 0, CPF_CFV_YES = 1, CPF_CFV_RUNTIME = 2 };

template <int CFV> CPF_DEV bool mesh_is_cfv(const MeshView &m) { return CFV == CPF_CFV_RUNTIME ? (m.tetcell == nullptr) : (CFV == CPF_CFV_YES); }

// origin id of tet `tet` whose record `f` has just been loaded from scratch (particle start, stage-walk restart)
template <int CFV> CPF_DEV int first_origin(const MeshView &m, int tet, const Fast32 &f)
{
    return mesh_is_cfv<CFV>(m) ? __ldg(reinterpret_cast<const int *>(m.tetv) + 4ll * tet + 3) : f.aux;
}

// RD3 < 0 marks a start point that still has to be re-expressed relative to the origin just requested (the
// subtraction is postponed to the next visit so that the warp does not wait for the position in the hop itself)
CPF_DEV void walkf_rebase(WalkF &ws, const D3 &O, const D3 &P0)
{
    ws.rx = (float)(P0.x - O.x); ws.ry = (float)(P0.y - O.y); ws.rz = (float)(P0.z - O.z);
    ws.RD3 = 3.f * (fmaxf(fmaxf(fabsf(ws.rx), fabsf(ws.ry)), fabsf(ws.rz)) + ws.Dd);
}

CPF_DEV void walkf_begin(WalkF &ws, const D3 &O, const D3 &P0, const D3 &disp, int tet, int org)
{
    ws.dx = (float)disp.x; ws.dy = (float)disp.y; ws.dz = (float)disp.z;
    ws.Dd = fmaxf(fmaxf(fabsf(ws.dx), fabsf(ws.dy)), fabsf(ws.dz));
    walkf_rebase(ws, O, P0);
    ws.t_in = 0.f;
    ws.cur = tet;
    ws.org = org;
    ws.path = 0u;
}

// C1 for a start point on its own (the all-particles pass checks it once per launch, with all lanes converged, before
// the visit loop): every plane function of the start tet at r = P - O must clear g0 = G*V6 + 2^-18 E^2 (E + 3R) -- the
// error of a_j does not involve the displacement.
CPF_DEV bool start_point_clear(const MeshView &m, const Fast32 &f, float rx, float ry, float rz)
{
    const float (&N)[3][3] = f.N;
    float a[4];
#pragma unroll
    for (int j = 0; j < 3; ++j) a[j] = rx * N[j][0] + ry * N[j][1] + rz * N[j][2];
    a[3] = f.V6 - a[0] - a[1] - a[2];
    const float E = fabsf(f.E);
    const float g0 = fmaf(m.guardf, f.V6, 3.814697265625e-6f * (E * E) * (E + 3.f * fmaxf(fmaxf(fabsf(rx), fabsf(ry)), fabsf(rz))));
    return fminf(fminf(a[0], a[1]), fminf(a[2], a[3])) >= g0;
}

// One tet visit.  Checks, all against g = G*V6 + ERR (header comment above):
//   C1  start point of the walk clear of every face plane -- NOT here: the kernels check it once per particle and launch,
//       with all lanes converged, before their visit loop (start_point_clear).
//       Entry points of later visits need no test of their own: they are the exit points C3 certified in the tet
//       before, and on the shared face the barycentric coordinates w.r.t. its three vertices are the same in both
//       tets.  Start points of later sub-steps are the end points C2 certified (P + disp in fp64 moves them by
//       <= 2^-52 |P|, far inside the 0.34 ERR of slack as long as |P| < 5e8 h_min, checked at mesh build).
//   C2  end point clear of every face plane; all e_j >= g: the walk ends here.
//   C3  exit point clear of the three other faces (edges/vertices, ties of dT, a wrongly selected exit).
// The entry face needs no special case: after C2 its e_j is certified positive (the plane function of the face a
// segment came in through increases along the segment), so it is never an exit candidate.
// lastVisit: the caller's visit cap is reached -- a hop out of this tet is refused BEFORE the next record is requested
// (keeps every use of the caller's state ahead of the loads: nothing in the hop path then touches a register the
// loads write, which ptxas otherwise resolves with a copy right behind them, i.e. a wait for the record inside the hop)
// PFU (cell-centre decompositions): a lane whose sub-step ends here will want this cell's velocity for its next one; the
// line is requested from the common block, ahead of the divergent sections.
template <int CFV, bool PFU = false>
CPF_DEV int visit_fast32(const MeshView &m, Fast32 &f, D3 &O, const D3 &P0, WalkF &ws, bool lastVisit = false)
{
    const float INF = __int_as_float(0x7f800000);
    if (ws.RD3 < 0.f) walkf_rebase(ws, O, P0); // entered another cell in the hop before: its origin has arrived by now
    const float (&N)[3][3] = f.N;
    const float V = f.V6;
    float a[4], b[4], e[4];
    {
        const F2 X = f2_pack(ws.rx, ws.dx), Y = f2_pack(ws.ry, ws.dy), Z = f2_pack(ws.rz, ws.dz);
        F2 ab[4];
#pragma unroll
        for (int j = 0; j < 3; ++j) ab[j] = f2_fma(N[j][2], Z, f2_fma(N[j][1], Y, f2_mul(N[j][0], X)));
        ab[3] = f2_sub(f2_sub(f2_sub(f2_pack(V, 0.f), ab[0]), ab[1]), ab[2]); // a_3 = V - a_0 - a_1 - a_2, b_3 = -(b_0 + b_1 + b_2)
#pragma unroll
        for (int j = 0; j < 4; ++j) f2_unpack(ab[j], a[j], b[j]);
    }
    const float E = fabsf(f.E);
    const float g = fmaf(m.guardf, V, 3.814697265625e-6f * (E * E) * (E + ws.RD3));
#pragma unroll
    for (int j = 0; j < 4; ++j) e[j] = a[j] + b[j];
    const bool done = fminf(fminf(e[0], e[1]), fminf(e[2], e[3])) >= g; // C2 and "inside" in one
    if (PFU) prefetch_l1_if(m.ucell + (ws.org - m.nPoints), done);
    if (done) return CPF_V_DONE;
    if (!(fminf(fminf(fabsf(e[0]), fabsf(e[1])), fminf(fabsf(e[2]), fabsf(e[3]))) >= g)) return CPF_V_REFUSE;
    // Exit face: smallest crossing parameter among the faces whose plane the end point is behind.  For those the
    // start/entry point is certified in front (true a_j + t_in b_j >= G V6), so b_j < 0 and t_j > t_in >= 0: positive
    // floats order like their bit patterns -- the face index rides in the two lowest mantissa bits through one
    // integer min.  Anything else (flushed denormal b_j -> +inf, a rounding-negative a_j -> sign bit set) fails
    // t_in < t <= 1 below.
    unsigned key = 0xffffffffu;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float tj = (e[j] < 0.f) ? a[j] * rcp_ftz(-b[j]) : INF;
        key = min(key, (__float_as_uint(tj) & ~3u) | (unsigned)j);
    }
    const int js = (int)(key & 3u);
    const float t = __uint_as_float(key & ~3u);
    // C3: the exit point must be clear of every other face (edges/vertices, ties of dT)
    float h[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) h[j] = (j == js) ? INF : fmaf(t, b[j], a[j]);
    const float c3m = fminf(fminf(h[0], h[1]), fminf(h[2], h[3]));
    const int link = sel4(f.link.x, f.link.y, f.link.z, f.link.w, js);
    // The record behind the selected face is requested BEFORE the C3 / range verdict is known (the compare chain runs while
    // the load is in flight).  A lane whose visit is refused never looks at its record registers again (it stops, or
    // reloads after an exact sub-step), and nothing is requested at a wall or at the visit cap.
    if (mesh_is_cfv<CFV>(m) && link >= 0 && !lastVisit) {
        const int aux = f.aux;
        if (js == 3) O = ld_vertex(m.vpos, aux);
        f32_load_pinned(m, link >> 2, f);
        if (!((c3m >= g) && (t > ws.t_in) && (t <= 1.f))) return CPF_V_REFUSE;
        ws.cur = link >> 2;
        ws.t_in = t;
        ws.path = (ws.path << 2) | (unsigned)js;
        if (js == 3) { ws.org = aux; ws.RD3 = -1.f; }
        return CPF_V_HOP;
    }
    if (!((c3m >= g) && (t > ws.t_in) && (t <= 1.f))) return CPF_V_REFUSE; // incl. "no candidate" (t = inf)
    if (link < 0) { ws.wall_js = js; ws.wall_link = link; return CPF_V_WALL; }
    // cell-centre decompositions known at compile time took every hop above: no second load site of the record
    if (lastVisit || CFV == CPF_CFV_YES) return CPF_V_REFUSE;
    ws.cur = link >> 2;
    ws.t_in = t;
    ws.path = (ws.path << 2) | (unsigned)js;
    if (mesh_is_cfv<CFV>(m)) {
        if (js == 3) { // through the face opposite the cell centre: another cell, whose origin id this record already names
            ws.org = f.aux;
            O = ld_vertex(m.vpos, f.aux);
            ws.RD3 = -1.f;
        }
        f32_load(m, ws.cur, f); // after the last use of the old record: the new words land in their loop-carried registers
    } else {
        f32_load(m, ws.cur, f);
        if (f.aux != ws.org) { // generic tet mesh: the record names its own origin
            ws.org = f.aux;
            O = ld_vertex(m.vpos, f.aux);
            walkf_rebase(ws, O, P0);
        }
    }
    return CPF_V_HOP;
}

// ------------------------------------------------------------------------------------------------
// RTX=true build (CPF_LOCATOR_BARY): the reference's baryTetSearch (query/RTQuery.cu:35-90) does not walk the segment, it
// walks towards the END POINT Q = P + disp: barycentric coordinates of Q in the current tet, inside if all >= 0, else
// across the face of the smallest one.  On the fp32 record the four plane functions e_j = N_j . (Q - O) are V6 times those
// coordinates, so the same walk is: all e_j >= g -> Q is certified inside (the reference's w_min >= 0 cannot depend on
// rounding); else the smallest e_j must be clearly negative (<= -g) and clearly the smallest (gap to the runner-up >= 2 g:
// the reference's first-minimum tie break never decides) -> hop across that face.  Boundary faces (the reflection of
// RTreflection), the visit cap and everything unclear are refused: the exact kernel redoes the sub-step.
// ws.rx.. hold Q - O (ws.dx.. unused, Dd = 0), ws.t_in > 0 marks "entered through a hop".
// ------------------------------------------------------------------------------------------------
template <int CFV>
CPF_DEV int visit_bary32(const MeshView &m, Fast32 &f, D3 &O, const D3 &Q, WalkF &ws, bool lastVisit)
{
    if (ws.RD3 < 0.f) walkf_rebase(ws, O, Q);
    const float (&N)[3][3] = f.N;
    float e[4];
#pragma unroll
    for (int j = 0; j < 3; ++j) e[j] = ws.rx * N[j][0] + ws.ry * N[j][1] + ws.rz * N[j][2];
    e[3] = f.V6 - e[0] - e[1] - e[2];
    const float E = fabsf(f.E);
    const float g = fmaf(m.guardf, f.V6, 3.814697265625e-6f * (E * E) * (E + ws.RD3));
    float m1 = e[0];
    int js = 0;
#pragma unroll
    for (int j = 1; j < 4; ++j)
        if (e[j] < m1) { m1 = e[j]; js = j; }
    if (m1 >= g) return CPF_V_DONE;
    const float INF = __int_as_float(0x7f800000);
    float m2 = INF;
#pragma unroll
    for (int j = 0; j < 4; ++j) m2 = fminf(m2, (j == js) ? INF : e[j]);
    if (!(m1 <= -g) || !(m2 - m1 >= 2.f * g)) return CPF_V_REFUSE;
    const int link = sel4(f.link.x, f.link.y, f.link.z, f.link.w, js);
    if (link < 0 || lastVisit) return CPF_V_REFUSE; // boundary: RTreflection runs in the reference's arithmetic
    ws.cur = link >> 2;
    ws.t_in = 1.f;
    if (mesh_is_cfv<CFV>(m)) {
        if (js == 3) {
            ws.org = f.aux;
            O = ld_vertex(m.vpos, f.aux);
            ws.RD3 = -1.f;
        }
        f32_load(m, ws.cur, f);
    } else {
        f32_load(m, ws.cur, f);
        if (f.aux != ws.org) {
            ws.org = f.aux;
            O = ld_vertex(m.vpos, f.aux);
            walkf_rebase(ws, O, Q);
        }
    }
    return CPF_V_HOP;
}

// Wall contact on the first leg of a sub-step, in the reference's arithmetic but only for the faces the filter
// has certified (C1-C3 passed at every visit; `path` holds the stored exit slot of each hop, 2 bits per hop; the
// exit face js of the last tet is a boundary face):
//   traceIntet (ConvexQuery.cu:32-131) accepts the exit face of every tet on the way with fd < tol, tol < dT <= 1
//   and moves the segment start to S = fma(d, dT, S), d = E - S: replayed here face by face, so that the hit
//   point P_hit carries exactly the reference's roundings;
//   reflectInTet (ConvexQuery.cu:239-317) takes the first face in its order with |fd| or |dT| below tol.  C3 puts
//   the hit point >= G*h >= 1e-11 away from every other face plane and |d| < 10 keeps their |dT| above 1e-12, so
//   that face is the wall face, provided |fd| < tol there -- checked exactly.
// Returns false when the reference's own tests on one of those faces do not hold (then the exact path decides).
// On success: Phit = hit point, E = reflected end point, u = reflected velocity; the caller continues the walk
// from Phit in the wall tet.
CPF_DEV bool exact_crossing(const MeshView &m, int tet, int js, bool flipped, const D3 &E, D3 &S, D3 &A, D3 &n)
{
    const int j = (js == 1 || js == 2) ? (flipped ? 3 - js : js) : js; // stored slot -> sorted face
    const int4 v = ld_int4(m.tetv, tet);
    A = ld_vertex(m.vpos, j == 0 ? v.y : v.x);
    const double *np = reinterpret_cast<const double *>(m.tetnrm) + 12ll * tet + 3 * j;
    n = D3{ __ldg(np), __ldg(np + 1), __ldg(np + 2) };
    const D3 d = xsub(E, S);
    const double fd = xdot(xsub(A, S), n);
    const double dT = __ddiv_rn(fd, xdot(d, n));
    if (!(fd < CPF_TOL && dT > CPF_TOL && dT <= 1.0)) return false;
    S = D3{ __fma_rn(d.x, dT, S.x), __fma_rn(d.y, dT, S.y), __fma_rn(d.z, dT, S.z) };
    return true;
}

CPF_DEV bool wall_reflect_on_path(const MeshView &m, int startTet, unsigned path, int nHops, int wallTet, int js, int wallLink, const D3 &P,
                                  const D3 &disp, D3 &Phit, D3 &E, D3 &u)
{
    E = xadd(P, disp);
    D3 S = P, A, n;
    int cur = startTet;
    for (int h = 0; h < nHops; ++h) { // the interior crossings before the wall tet
        const unsigned *rec = reinterpret_cast<const unsigned *>(m.tetfast + 4ll * cur);
        const int hj = (int)((path >> (2 * (nHops - 1 - h))) & 3u); // the latest hop sits in the lowest bits
        const bool flipped = (int)__ldg(rec + 15) < 0; // sign bit of E
        const int link = (int)__ldg(rec + hj);
        if (link < 0 || !exact_crossing(m, cur, hj, flipped, E, S, A, n)) return false;
        cur = link >> 2;
    }
    if (cur != wallTet) return false;
    const bool flipped = (int)__ldg(reinterpret_cast<const unsigned *>(m.tetfast + 4ll * cur) + 15) < 0;
    if (!exact_crossing(m, cur, js, flipped, E, S, A, n)) return false;
    Phit = S;
    if (!(fabs(xdot(xsub(A, Phit), n)) < CPF_TOL)) return false;
    const double gain = __ldg(m.patch_gain - wallLink - 1);
    const D3 r = xsub(E, A);
    double sp = -__fma_rn(r.z, n.z, __fma_rn(r.y, n.y, __dmul_rn(r.x, n.x)));
    sp = __dmul_rn(gain, sp);
    double sv = -__fma_rn(u.z, n.z, __fma_rn(u.y, n.y, __dmul_rn(u.x, n.x)));
    sv = __dmul_rn(gain, sv);
    E = D3{ __fma_rn(sp, n.x, E.x), __fma_rn(sp, n.y, E.y), __fma_rn(sp, n.z, E.z) };
    u = D3{ __fma_rn(sv, n.x, u.x), __fma_rn(sv, n.y, u.y), __fma_rn(sv, n.z, u.z) };
    return true;
}

} // namespace cpf
