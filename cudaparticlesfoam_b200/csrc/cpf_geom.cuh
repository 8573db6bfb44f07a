// cpf_geom.cuh -- device geometry for the particle hot path (sm_100a).
//
// Two arithmetic policies live here:
//  * EXACT: the reference's expression trees with the FMA placement nvcc/ptxas gives the reference
//    kernels (read from their sm_100a SASS; DESIGN.md "FMA map").  Written with explicit
//    round-to-nearest intrinsics so that no compiler decision can change a bit.  Follows
//    third_party/RTXAdvect/cuda/DeviceTetMesh.cuh:82-156,193-199, query/ConvexQuery.cu:32-131,
//    239-317, query/RTQuery.cu:35-107 and owl/common/math/vec.h:317-345 of the reference.
//  * FILTERED: cheap un-normalised predicates with guard bands; whenever a decision of the
//    reference could depend on rounding (or a wall is touched) it reports NEED_EXACT and the
//    caller re-runs the sub-step with the EXACT policy, so results stay bit-identical.
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
    int4 link;
    unsigned code;
};

struct MeshView {
    const double4 *__restrict__ vpos; // [nVerts]
    const int4 *__restrict__ tetv;    // [nTets] sorted vertex ids
    const int4 *__restrict__ tetrec;  // [nTets][2] {links, apex ids} (32 B, one 256-bit load)
    const uint16_t *__restrict__ tetcode;
    const int *__restrict__ tetcell;  // [nTets] or nullptr (cell = max vertex id - nPoints)
    const double *__restrict__ ucell; // [nCells][3]
    const double *__restrict__ uvert; // [nVerts][3] (CPF_INTERP_VERTEX)
    const uint8_t *__restrict__ patch_kind; // [nPatches]
    int nPoints;
    long long nTets;
    int nCells;
    double guard; // barycentric guard band of the filtered path
};

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

// inward unit normal of sorted face j exactly as traceIntet builds it (ConvexQuery.cu:73-79)
CPF_DEV D3 face_normal_exact(const Tet &T, int j, D3 &A)
{
    D3 B, C;
    face_abc(T, j, A, B, C);
    D3 n = xtriNorm(A, B, C);
    if ((T.code >> (8 + j)) & 1u) n = xneg(n);
    return n;
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
        double fd = xdot(xsub(A, P0), n);
        double dT = __ddiv_rn(fd, xdot(d, n));
        if (isinf(dT)) dT = -1.0;
        if (fd < CPF_TOL && dT > CPF_TOL && dT <= 1.0 && dT < best) {
            best = dT;
            out_j = j;
            S.x = __fma_rn(d.x, dT, P0.x);
            S.y = __fma_rn(d.y, dT, P0.y);
            S.z = __fma_rn(d.z, dT, P0.z);
        }
    }
    return out_j;
}

// query/ConvexQuery.cu:239-317 reflectInTet (see oracle/cpf_oracle.c reflect_in_tet for the
// uninitialised-read note: no matching face leaves E and u untouched).
CPF_DEV void reflect_exact(const Tet &T, D3 Pxf, D3 &E, D3 &u)
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
            D3 r = xsub(E, A);
            double sp = -__fma_rn(r.z, n.z, __fma_rn(r.y, n.y, __dmul_rn(r.x, n.x)));
            sp = __dadd_rn(sp, sp);
            double sv = -__fma_rn(u.z, n.z, __fma_rn(u.y, n.y, __dmul_rn(u.x, n.x)));
            sv = __dadd_rn(sv, sv);
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

// ------------------------------------------------------------------------------------------------
// FILTERED policy.
//
// The walk carries the current tet's geometry in registers (WalkState): four vertex ids + positions
// and, per register slot, the link and the "apex" vertex id of the neighbour across the face
// opposite that slot.  Crossing a face replaces exactly one vertex, so a hop costs ONE memory round
// (the neighbour's 32-byte record and its single new vertex, issued together) instead of the
// reference's four dependent gathers, and a particle that stays in its tet costs no mesh load at
// all when sub-steps are fused.
//
// Decisions use un-normalised plane functions a_j + t*b_j (>= 0 inside) of the ORIGINAL segment
// P0 -> P0+d, and the walk refuses (CPF_NEED_EXACT) whenever a decision of the reference could
// depend on rounding:
//   C1 start/entry point within the guard band of another face
//   C2 end point within the guard band of any face plane of the visited tet
//   C3 exit point within the guard band of another face (edge/vertex grazing, ties of dT)
//   a wall is reached, no consistent exit exists, or the reference's 50-tet cap is approached.
// ------------------------------------------------------------------------------------------------
#define CPF_NEED_EXACT (-1)

// Register roles: slot 0 holds the vertex that entered last (the "apex": the entry face is the one
// opposite slot 0), slots 1..3 the three vertices shared with the previous tet.  ord[q] is the slot
// of role q's vertex in the tet's stored record (ascending vertex ids), so links and apex ids are
// picked from the raw record on demand instead of being permuted on every hop.
struct WalkState {
    D3 X[4];
    int4 link, apex; // raw record of the current tet (ascending-id order)
    int ord[4];
    int cell;
};

CPF_DEV void ld_rec(const int4 *__restrict__ rec, int t, int4 &link, int4 &apex)
{
    asm("ld.global.nc.v8.s32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(link.x), "=r"(link.y), "=r"(link.z), "=r"(link.w), "=r"(apex.x), "=r"(apex.y), "=r"(apex.z), "=r"(apex.w)
        : "l"(rec + 2ll * t));
}

CPF_DEV int sel4(int a, int b, int c, int d, int k)
{
    const int lo = (k & 1) ? b : a, hi = (k & 1) ? d : c;
    return (k & 2) ? hi : lo;
}
CPF_DEV double dsel(bool c, double a, double b) { return c ? a : b; }

// cold start: tet id -> full geometry (two dependent rounds: ids, then positions)
CPF_DEV void ws_load(const MeshView &m, int tet, WalkState &ws)
{
    const int4 v = ld_int4(m.tetv, tet);
    ld_rec(m.tetrec, tet, ws.link, ws.apex);
    ws.X[0] = ld_vertex(m.vpos, v.x); ws.X[1] = ld_vertex(m.vpos, v.y);
    ws.X[2] = ld_vertex(m.vpos, v.z); ws.X[3] = ld_vertex(m.vpos, v.w);
    ws.ord[0] = 0; ws.ord[1] = 1; ws.ord[2] = 2; ws.ord[3] = 3;
    ws.cell = m.tetcell ? 0 : v.w - m.nPoints;
}

// cross the face opposite role s (es = its record slot, link >= 0): one memory round
CPF_DEV int ws_hop(const MeshView &m, WalkState &ws, int s, int es, int link)
{
    const int nid = sel4(ws.apex.x, ws.apex.y, ws.apex.z, ws.apex.w, es);
    const int tet = link >> 2;
    const int inj = link & 3; // record slot of the new apex inside the neighbour
    ld_rec(m.tetrec, tet, ws.link, ws.apex);
    const D3 Xn = ld_vertex(m.vpos, nid);
    // kept vertices keep their relative id order: slot o of the old record -> (o minus the removed
    // slot) -> (plus the inserted apex slot) in the neighbour's record
    int no[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int k = ws.ord[q] - (ws.ord[q] > es ? 1 : 0);
        no[q] = k + (k >= inj ? 1 : 0);
    }
    // the previous apex (role 0) becomes one of the shared three: it takes the role that left
#pragma unroll
    for (int q = 1; q < 4; ++q) {
        const bool mv = (q == s);
        ws.X[q].x = dsel(mv, ws.X[0].x, ws.X[q].x);
        ws.X[q].y = dsel(mv, ws.X[0].y, ws.X[q].y);
        ws.X[q].z = dsel(mv, ws.X[0].z, ws.X[q].z);
        ws.ord[q] = mv ? no[0] : no[q];
    }
    ws.X[0] = Xn;
    ws.ord[0] = inj;
    if (!m.tetcell && nid >= m.nPoints) ws.cell = nid - m.nPoints; // a cell-centre vertex entered
    return tet;
}

// ~2^-20 seed + two Newton steps: a few ulp, no branches (the IEEE division's slow path and its
// divergence are not needed for a filtered decision)
CPF_DEV double fast_rcp(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(fma(-x, r, 1.0), r, r);
    r = fma(fma(-x, r, 1.0), r, r);
    return r;
}

// Walks from the tet held in ws (containing P0) along d.  Returns the final tet id (ws then holds
// its geometry) or CPF_NEED_EXACT (ws is left somewhere along the way and must be reloaded).
CPF_DEV int walk_filtered(const MeshView &m, WalkState &ws, int tet0, D3 P0, D3 d, unsigned &hops)
{
    int cur = tet0;
    bool first = true; // the first tet has no entry face; afterwards it is the face opposite role 0
    double t_in = 0.0;
    for (int it = 0; it < 48; ++it) {
        hops++;
        const D3 e1{ ws.X[1].x - ws.X[0].x, ws.X[1].y - ws.X[0].y, ws.X[1].z - ws.X[0].z };
        const D3 e2{ ws.X[2].x - ws.X[0].x, ws.X[2].y - ws.X[0].y, ws.X[2].z - ws.X[0].z };
        const D3 e3{ ws.X[3].x - ws.X[0].x, ws.X[3].y - ws.X[0].y, ws.X[3].z - ws.X[0].z };
        const D3 n1{ e2.y * e3.z - e2.z * e3.y, e2.z * e3.x - e2.x * e3.z, e2.x * e3.y - e2.y * e3.x };
        const D3 n2{ e3.y * e1.z - e3.z * e1.y, e3.z * e1.x - e3.x * e1.z, e3.x * e1.y - e3.y * e1.x };
        const D3 n3{ e1.y * e2.z - e1.z * e2.y, e1.z * e2.x - e1.x * e2.z, e1.x * e2.y - e1.y * e2.x };
        const double V6s = e1.x * n1.x + e1.y * n1.y + e1.z * n1.z;
        const double sg = V6s < 0.0 ? -1.0 : 1.0; // orientation: make "inside" positive
        const D3 r{ P0.x - ws.X[0].x, P0.y - ws.X[0].y, P0.z - ws.X[0].z };
        double a[4], b[4];
        a[1] = sg * (r.x * n1.x + r.y * n1.y + r.z * n1.z);
        a[2] = sg * (r.x * n2.x + r.y * n2.y + r.z * n2.z);
        a[3] = sg * (r.x * n3.x + r.y * n3.y + r.z * n3.z);
        b[1] = sg * (d.x * n1.x + d.y * n1.y + d.z * n1.z);
        b[2] = sg * (d.x * n2.x + d.y * n2.y + d.z * n2.z);
        b[3] = sg * (d.x * n3.x + d.y * n3.y + d.z * n3.z);
        const double V6 = fabs(V6s);
        a[0] = V6 - a[1] - a[2] - a[3];
        b[0] = -(b[1] + b[2] + b[3]);
        const double g = m.guard * V6;
        bool bad = false, allin = true;
        double as = 0.0, nbs = 1.0; // best exit so far: t = as/nbs
        int js = -1;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double e = a[j] + b[j];
            const bool notin = (j != 0) | first;
            bad |= notin & (fma(t_in, b[j], a[j]) < g); // C1
            bad |= fabs(e) < g;                         // C2
            allin &= (e > 0.0);
            const double nb = -b[j];
            const bool cand = notin & (nb > 0.0) & (e < 0.0);
            const bool better = cand & ((js < 0) | (a[j] * nbs < as * nb)); // t_j < t_best, no division
            as = dsel(better, a[j], as);
            nbs = dsel(better, nb, nbs);
            js = better ? j : js;
        }
        if (bad) return CPF_NEED_EXACT;
        if (allin) return cur;
        if (js < 0) return CPF_NEED_EXACT;
        const double t = as * fast_rcp(nbs);
#pragma unroll
        for (int j = 0; j < 4; ++j) bad |= (j != js) & (fma(t, b[j], a[j]) < g); // C3
        const int es = sel4(ws.ord[0], ws.ord[1], ws.ord[2], ws.ord[3], js);
        const int link = sel4(ws.link.x, ws.link.y, ws.link.z, ws.link.w, es);
        if (bad | !(t - t_in >= 1e-9 * (1.0 - t_in)) | !(t <= 1.0) | (link < 0)) return CPF_NEED_EXACT; // incl. walls
        cur = ws_hop(m, ws, js, es, link);
        first = false;
        t_in = t;
    }
    return CPF_NEED_EXACT;
}

} // namespace cpf
