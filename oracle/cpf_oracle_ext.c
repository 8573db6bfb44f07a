/*
 * cpf_oracle_ext.c -- oracle for the features the north star names but the REFERENCE DOES NOT
 * IMPLEMENT: RK2/RK4 integration, per-patch outlet escape, and the cellPoint-style (vertex)
 * interpolation driven through the full sub-step loop.  Included at the end of cpf_oracle.c.
 *
 * PARITY UNPINNED against the reference for everything in this file: there is no reference
 * behaviour to pin to (SURVEY.md "Facts": only Euler + per-tet constant velocity + all-reflecting
 * walls are wired up).  This file DEFINES the semantics; the CUDA path is tested against it bit for
 * bit, and against analytic properties (convergence order, escape counts).  The building blocks
 * (segment walk, reflection, vertex interpolation weights) are the pinned reference arithmetic of
 * cpf_oracle.c.
 */

/* tet -> polyMesh face it was fanned from (same traversal as orc_decompose_poly) */
long orc_tet_polyface(int nCells, int nFaces, int nInternal, const int *faceOffsets, const int *owner,
                      const int *neighbour, int *tetFace)
{
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
            int n = faceOffsets[f + 1] - faceOffsets[f];
            for (int tetPt = 1; tetPt <= n - 2; ++tetPt) tetFace[nT++] = f;
        }
    free(cnt); free(cf); free(fill);
    return nT;
}

typedef struct {
    int integrator;            /* 0 Euler, 1 RK2 (midpoint), 4 RK4 */
    int vertexVelocity;        /* 0: U is [nTets][3], 1: U is [nVerts][3] (P1 / cellPoint-style) */
    const unsigned char *faceKind; /* [nFaces] 0 reflect, 1 escape; NULL = all reflect */
    const double *faceGain;        /* [nFaces] 1 + restitution coefficient of the face's patch; NULL = 2 (specular) everywhere */
    const double *U;
} ext_opts;

/* P1 interpolation with the weights of cuda/particles.cu:281-295 (s1_advect_vertvel arithmetic) */
static v3 ext_vertex_velocity(v3 P, int tet, const orc_mesh *m, const double *Uvert)
{
    const int *ix = m->idx + 4 * tet;
    v3 A = ld3(m->pos, ix[0]), B = ld3(m->pos, ix[1]), C = ld3(m->pos, ix[2]), D = ld3(m->pos, ix[3]);
    double rden = 1.0 / ref_det(A, B, C, D);
    double wA = ref_det(P, B, C, D) * rden;
    double wB = ref_det(A, P, C, D) * rden;
    double wC = ref_det(A, B, P, D) * rden;
    v3 cr = ref_cross(v3_sub(B, A), v3_sub(C, A));
    v3 rr = v3_sub(P, A);
    double wD = fma(cr.z, rr.z, fma(cr.y, rr.y, cr.x * rr.x)) * rden;
    v3 uA = ld3(Uvert, ix[0]), uB = ld3(Uvert, ix[1]), uC = ld3(Uvert, ix[2]), uD = ld3(Uvert, ix[3]);
    v3 v;
    v.x = fma(wD, uD.x, fma(wC, uC.x, fma(wA, uA.x, wB * uB.x)));
    v.y = fma(wD, uD.y, fma(wC, uC.y, fma(wA, uA.y, wB * uB.y)));
    v.z = fma(wD, uD.z, fma(wC, uC.z, fma(wA, uA.z, wB * uB.z)));
    return v;
}

static v3 ext_velocity(v3 P, int tet, const orc_mesh *m, const ext_opts *o)
{
    return o->vertexVelocity ? ext_vertex_velocity(P, tet, m, o->U) : ld3(o->U, tet);
}

/* tet that contains `to` when walking the segment from -> to from tet `t` with the reference's
 * line walk; a stage point beyond a wall is attributed to the last tet before the wall */
static int ext_stage_tet(v3 from, v3 to, int t, const orc_mesh *m)
{
    v3 S = from;
    int cur = t, next, OutFace = -2, InFace = -2;
    for (int i = 0; i < 50; ++i) {
        next = trace_in_tet(&S, to, cur, m, &OutFace, InFace);
        if (next == cur || next == -1) break;
        InFace = OutFace;
        cur = next;
    }
    return cur;
}

static inline v3 ext_axpy(double h, v3 k, v3 P) { v3 r = { fma(h, k.x, P.x), fma(h, k.y, P.y), fma(h, k.z, P.z) }; return r; }

/* S1 generalised: effective velocity of the step by the chosen integrator */
static void ext_s1(double *p, int tet, double *vel, double *disp, double dt, const orc_mesh *m, const ext_opts *o)
{
    if (p[3] == 0.0) return;
    if (tet < 0) { p[3] = 0.0; return; }
    v3 P = { p[0], p[1], p[2] };
    v3 k1 = ext_velocity(P, tet, m, o), v = k1;
    if (o->integrator == 1) {
        v3 Pm = ext_axpy(0.5 * dt, k1, P);
        v = ext_velocity(Pm, ext_stage_tet(P, Pm, tet, m), m, o);
    } else if (o->integrator == 4) {
        const double h = 0.5 * dt;
        v3 P2 = ext_axpy(h, k1, P);
        v3 k2 = ext_velocity(P2, ext_stage_tet(P, P2, tet, m), m, o);
        v3 P3 = ext_axpy(h, k2, P);
        v3 k3 = ext_velocity(P3, ext_stage_tet(P, P3, tet, m), m, o);
        v3 P4 = ext_axpy(dt, k3, P);
        v3 k4 = ext_velocity(P4, ext_stage_tet(P, P4, tet, m), m, o);
        v.x = fma(2.0, k2.x + k3.x, k1.x + k4.x) / 6.0;
        v.y = fma(2.0, k2.y + k3.y, k1.y + k4.y) / 6.0;
        v.z = fma(2.0, k2.z + k3.z, k1.z + k4.z) / 6.0;
    }
    disp[0] = fma(dt, v.x, p[0]) - p[0];
    disp[1] = fma(dt, v.y, p[1]) - p[1];
    disp[2] = fma(dt, v.z, p[2]) - p[2];
    disp[3] = -1.0;
    vel[0] = v.x; vel[1] = v.y; vel[2] = v.z; vel[3] = -1.0;
}

/* reflectInTet (cpf_oracle.c reflect_in_tet) with the rebound model: the mirrored part of the end point and of the
 * velocity is scaled by gain = 1 + e of the matching face's patch; gain = 2 reproduces s + s bit for bit. */
static void ext_reflect_in_tet(v3 Pxf, v3 *P_end, v3 *vel, int tet, const orc_mesh *m, const double *faceGain)
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
            const double gain = faceGain ? faceGain[f] : 2.0;
            v3 r = v3_sub(*P_end, A);
            double sp = -fma(r.z, n.z, fma(r.y, n.y, r.x * n.x));
            sp = gain * sp;
            double sv = -fma(vel->z, n.z, fma(vel->y, n.y, vel->x * n.x));
            sv = gain * sv;
            P_end->x = fma(sp, n.x, P_end->x); P_end->y = fma(sp, n.y, P_end->y); P_end->z = fma(sp, n.z, P_end->z);
            vel->x = fma(sv, n.x, vel->x); vel->y = fma(sv, n.y, vel->y); vel->z = fma(sv, n.z, vel->z);
            return;
        }
    }
}

/* S4 generalised: convexReflector with a per-patch action at every wall contact.  ESCAPE: the
 * particle is parked at the exit point, deactivated (w = 0) and keeps the negative id
 * -(tet at exit + 1); returns 1 when the particle escaped. */
static int ext_s4(double *p, double *disp, double *vel, int *tetIO, const orc_mesh *m, const ext_opts *o)
{
    if (p[3] == 0.0) return 0;
    int tetID = *tetIO;
    if (tetID >= 0) return 0;
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
        if (o->faceKind && o->faceKind[OutFace] == 1) {
            p[0] = P_hit.x; p[1] = P_hit.y; p[2] = P_hit.z; p[3] = 0.0;
            disp[0] = disp[1] = disp[2] = 0.0;
            *tetIO = -(cur + 1);
            return 1;
        }
        ext_reflect_in_tet(P_hit, &P_end, &u, cur, m, o->faceGain);
    }
    v3 nd = v3_sub(P_end, P_hit);
    p[0] = P_hit.x; p[1] = P_hit.y; p[2] = P_hit.z;
    disp[0] = nd.x; disp[1] = nd.y; disp[2] = nd.z;
    vel[0] = u.x; vel[1] = u.y; vel[2] = u.z;
    *tetIO = next;
    return 0;
}

/* generalised sub-step loop (default ConvexPoly build only); returns the number of escapes */
long orc_ext_substeps(long n, int nSteps, double *p, int *tet, double *vel, double *disp, double dt,
                      MESH_ARGS, const double *U, int vertexVelocity, int integrator,
                      const unsigned char *faceKind, int reflectWall, const double *xi, double D, const double *faceGain)
{
    MESH_INIT;
    ext_opts o = { integrator, vertexVelocity, faceKind, faceGain, U };
    long escaped = 0;
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : escaped)
    for (long i = 0; i < n; ++i) {
        double *pi = p + 4 * i, *vi = vel + 4 * i, *di = disp + 4 * i;
        int *ti = tet + i;
        for (int s = 0; s < nSteps; ++s) {
            ext_s1(pi, *ti, vi, di, dt, &m, &o);
            if (xi) s2_brownian(pi, di, xi + ((size_t)s * (size_t)n + (size_t)i) * 3u, D, dt);
            s3_locate_convex(pi, di, ti, &m);
            if (reflectWall) escaped += ext_s4(pi, di, vi, ti, &m, &o);
            s5_move(pi, di);
        }
    }
    return escaped;
}

/* OpenFOAM volPointInterpolation (interior rule): point value = sum_c w_c U_c / sum_c w_c over the
 * cells c that use the point (ascending), w_c = 1/|p - C_c|.  Boundary-condition corrections of
 * OpenFOAM are not restated.  Output Uvert = [point values..., cell values...]. */
void orc_point_values(int nPoints, int nCells, int nFaces, int nInternal, const int *faceOffsets, const int *faceVerts,
                      const int *owner, const int *neighbour, const double *points, const double *centres,
                      const double *Ucell, double *Uvert)
{
    /* point -> cells, ascending and unique */
    long nPairs = 0;
    for (int f = 0; f < nFaces; ++f) nPairs += (long)(faceOffsets[f + 1] - faceOffsets[f]) * (f < nInternal ? 2 : 1);
    long long *pairs = (long long *)malloc(sizeof(long long) * (size_t)(nPairs ? nPairs : 1));
    long k = 0;
    for (int f = 0; f < nFaces; ++f)
        for (int q = faceOffsets[f]; q < faceOffsets[f + 1]; ++q) {
            pairs[k++] = ((long long)faceVerts[q] << 32) | (unsigned)owner[f];
            if (f < nInternal) pairs[k++] = ((long long)faceVerts[q] << 32) | (unsigned)neighbour[f];
        }
    /* simple in-place heap sort free: use qsort */
    int cmp(const void *a, const void *b);
    qsort(pairs, (size_t)nPairs, sizeof(long long), cmp);
    long i = 0;
    for (int p = 0; p < nPoints; ++p) {
        v3 P = ld3(points, p);
        double sumw = 0.0, ax = 0.0, ay = 0.0, az = 0.0;
        long long last = -1;
        while (i < nPairs && (int)(pairs[i] >> 32) == p) {
            if (pairs[i] != last) {
                int c = (int)(pairs[i] & 0xffffffffll);
                v3 d = v3_sub(P, ld3(centres, c));
                double wgt = 1.0 / sqrt(fma(d.z, d.z, fma(d.y, d.y, d.x * d.x)));
                sumw = sumw + wgt;
                ax = fma(wgt, Ucell[3 * c], ax); ay = fma(wgt, Ucell[3 * c + 1], ay); az = fma(wgt, Ucell[3 * c + 2], az);
                last = pairs[i];
            }
            ++i;
        }
        Uvert[3 * p] = ax / sumw; Uvert[3 * p + 1] = ay / sumw; Uvert[3 * p + 2] = az / sumw;
    }
    for (int c = 0; c < nCells; ++c) {
        Uvert[3 * (nPoints + c)] = Ucell[3 * c]; Uvert[3 * (nPoints + c) + 1] = Ucell[3 * c + 1]; Uvert[3 * (nPoints + c) + 2] = Ucell[3 * c + 2];
    }
    free(pairs);
}
int cmp(const void *a, const void *b) { long long x = *(const long long *)a, y = *(const long long *)b; return x < y ? -1 : (x > y ? 1 : 0); }
