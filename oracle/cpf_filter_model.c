/* cpf_filter_model.c -- TEST INFRASTRUCTURE ONLY (included by cpf_oracle.c).
 *
 * A CPU model of the product's fp32 guarded walk (cudaparticlesfoam_b200/csrc/cpf_geom.cuh: f32_load,
 * walkf_begin, visit_fast32; record construction: cpf_mesh.cu k_build_fast), written against the oracle's
 * mesh tables, so that the SOUNDNESS claim of DESIGN.md section 4.1 can be tested on the CPU over millions of
 * random and adversarial segments: whenever the filter certifies a walk (no refusal), the tet it ends in must be
 * the tet the reference's fp64 segment walk (s3_locate_convex) ends in.  The model rounds like the product where
 * the source spells it (fmaf, one rounding per stored quantity); where nvcc is free to contract a*b+c the model
 * does not -- the error term of the filter covers either choice, which is exactly what the test is about. */

typedef struct {
    int link[4];     /* (nbr << 2) | stored slot in nbr, or negative: boundary */
    float N[3][3];   /* inward normals of the faces opposite stored slots 0..2 (slot 3 = origin = highest id) */
    int origin;
    float V6, E;     /* E < 0: stored slots 1/2 exchanged w.r.t. the sorted vertex order */
    int face[4];     /* model only: face id behind each stored slot (the product derives it from tetv / tetnrm) */
} fm_rec;

static void fm_sorted(const orc_mesh *m, int t, int s[4], int perm[4])
{
    for (int k = 0; k < 4; ++k) { s[k] = m->idx[4 * t + k]; perm[k] = k; }
    for (int a = 0; a < 4; ++a)
        for (int b = a + 1; b < 4; ++b)
            if (s[b] < s[a]) { int x = s[a]; s[a] = s[b]; s[b] = x; x = perm[a]; perm[a] = perm[b]; perm[b] = x; }
}

static double fm_sorted_v6(const orc_mesh *m, const int s[4], double X[3][3])
{
    const v3 O = ld3(m->pos, s[3]);
    for (int k = 0; k < 3; ++k) {
        const v3 p = ld3(m->pos, s[k]);
        X[k][0] = p.x - O.x; X[k][1] = p.y - O.y; X[k][2] = p.z - O.z;
    }
    return X[0][0] * (X[1][1] * X[2][2] - X[1][2] * X[2][1]) + X[0][1] * (X[1][2] * X[2][0] - X[1][0] * X[2][2]) +
           X[0][2] * (X[1][0] * X[2][1] - X[1][1] * X[2][0]);
}

/* records for all tets; returns the smallest tet height (for the guard G = max(1e-7, 1e-11 / hmin)) */
double orc_filter_build(long nTets, MESH_ARGS, void *recsOut)
{
    MESH_INIT;
    fm_rec *recs = (fm_rec *)recsOut;
    double hmin = 1e300;
    for (long t = 0; t < nTets; ++t) {
        int s[4], perm[4];
        double X[3][3];
        fm_sorted(&m, (int)t, s, perm);
        const double v6 = fm_sorted_v6(&m, s, X);
        const int flip = v6 < 0.0;
        fm_rec *r = recs + t;
        for (int j = 0; j < 4; ++j) { /* sorted slot j -> stored slot */
            const int st = (flip && (j == 1 || j == 2)) ? 3 - j : j;
            const int f = m.tetfacets[4 * t + perm[j]];
            const int nbr = other_tet(m.finfo, f, (int)t);
            r->face[st] = f;
            if (nbr < 0) { r->link[st] = -1; continue; }
            int s2[4], perm2[4], ns = -1;
            double X2[3][3];
            fm_sorted(&m, nbr, s2, perm2);
            for (int q = 0; q < 4; ++q)
                if (m.tetfacets[4 * nbr + perm2[q]] == f) ns = q;
            if ((ns == 1 || ns == 2) && fm_sorted_v6(&m, s2, X2) < 0.0) ns = 3 - ns;
            r->link[st] = (nbr << 2) | ns;
        }
        if (flip) { for (int c = 0; c < 3; ++c) { const double x = X[1][c]; X[1][c] = X[2][c]; X[2][c] = x; } }
        float E = 0.f;
        for (int k = 0; k < 3; ++k)
            for (int c = 0; c < 3; ++c) E = fmaxf(E, fabsf((float)X[k][c]));
        E = E * 1.0000002f;
        r->E = flip ? -E : E;
        double nmax = 0.0, n3[3] = { 0, 0, 0 };
        for (int j = 0; j < 3; ++j) {
            const double *u = X[(j + 1) % 3], *w = X[(j + 2) % 3];
            const double n[3] = { u[1] * w[2] - u[2] * w[1], u[2] * w[0] - u[0] * w[2], u[0] * w[1] - u[1] * w[0] };
            for (int c = 0; c < 3; ++c) { r->N[j][c] = (float)n[c]; n3[c] -= n[c]; }
            nmax = fmax(nmax, n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
        }
        nmax = fmax(nmax, n3[0] * n3[0] + n3[1] * n3[1] + n3[2] * n3[2]);
        r->origin = s[3];
        r->V6 = (float)fabs(v6);
        const double h = fabs(v6) / sqrt(nmax);
        if (h < hmin) hmin = h;
    }
    return hmin;
}

/* One sub-step walk of every particle: out[i] = final tet if every visit was certified, -1 if the filter refused
 * (guard band, wall, no exit, visit cap); visits[i] = tets visited.  errScale scales the rounding-error term of the
 * guard (1 = the product's; 0 with guard = 0 switches the band off, for the test that shows what it is there for).
 * skipC1First (may be NULL): per particle, drop the C1 test of the first visit -- what the product does for every sub-step
 * whose start point is the end point C2 certified in the sub-step before. */
void orc_filter_walk(long n, const double *p, const double *disp, const int *tet, const void *recsIn, const double *pos,
                     double guard, double errScale, const unsigned char *skipC1First, int *out, int *visits)
{
    const fm_rec *recs = (const fm_rec *)recsIn;
    const float INF = INFINITY, G = (float)guard * 1.0000002f, ES = (float)errScale * 3.814697265625e-6f;
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) {
        out[i] = -1;
        visits[i] = 0;
        int cur = tet[i];
        if (cur < 0 || p[4 * i + 3] == 0.0) continue;
        const double *P0 = p + 4 * i, *d = disp + 4 * i;
        const fm_rec *f = recs + cur;
        const double *O = pos + 3 * (long)f->origin;
        float rx = (float)(P0[0] - O[0]), ry = (float)(P0[1] - O[1]), rz = (float)(P0[2] - O[2]);
        const float dx = (float)d[0], dy = (float)d[1], dz = (float)d[2];
        const float Dd = fmaxf(fmaxf(fabsf(dx), fabsf(dy)), fabsf(dz));
        float RD3 = 3.f * (fmaxf(fmaxf(fabsf(rx), fabsf(ry)), fabsf(rz)) + Dd);
        float t_in = 0.f;
        int result = -1;
        for (int it = 0; it < 48; ++it) {
            visits[i]++;
            float a[4], b[4], e[4];
            for (int j = 0; j < 3; ++j) {
                a[j] = rx * f->N[j][0] + ry * f->N[j][1] + rz * f->N[j][2];
                b[j] = dx * f->N[j][0] + dy * f->N[j][1] + dz * f->N[j][2];
            }
            const float V = f->V6, E = fabsf(f->E);
            a[3] = V - a[0] - a[1] - a[2];
            b[3] = -(b[0] + b[1] + b[2]);
            const float g = fmaf(G, V, ES * (E * E) * (E + RD3));
            float amin = INF, eam = INF, emin = INF;
            for (int j = 0; j < 4; ++j) {
                e[j] = a[j] + b[j];
                amin = fminf(amin, a[j]);
                eam = fminf(eam, fabsf(e[j]));
                emin = fminf(emin, e[j]);
            }
            /* C1: the start point of the sub-step against every face, FIRST visit only (entry points of later visits are
             * the exit points C3 certified in the tet before: same barycentric coordinates on the shared face) */
            if (it == 0 && !(skipC1First && skipC1First[i])) { /* cpf_geom.cuh start_point_clear: the error of a_j does not involve d */
                const float g0 = fmaf(G, V, ES * (E * E) * (E + 3.f * fmaxf(fmaxf(fabsf(rx), fabsf(ry)), fabsf(rz))));
                if (!(amin >= g0)) break;
            }
            if (emin >= g) { result = cur; break; }                         /* C2 + inside: done */
            if (!(eam >= g)) break;                                         /* C2: refuse */
            float t = INF;
            int js = -1;
            for (int j = 0; j < 4; ++j) {
                if (e[j] < 0.f) {
                    const float tj = a[j] * (1.0f / -b[j]);
                    if (tj < t) { t = tj; js = j; }
                }
            }
            if (js < 0) break;
            float c3m = INF;
            for (int j = 0; j < 4; ++j)
                if (j != js) c3m = fminf(c3m, fmaf(t, b[j], a[j]));
            if (!((c3m >= g) && (t > t_in) && (t <= 1.f))) break;
            const int link = f->link[js];
            if (link < 0) break;                                            /* wall: the exact path reflects */
            cur = link >> 2;
            t_in = t;
            const int oldOrigin = f->origin;
            f = recs + cur;
            if (f->origin != oldOrigin) {
                O = pos + 3 * (long)f->origin;
                rx = (float)(P0[0] - O[0]); ry = (float)(P0[1] - O[1]); rz = (float)(P0[2] - O[2]);
                RD3 = 3.f * (fmaxf(fmaxf(fabsf(rx), fabsf(ry)), fabsf(rz)) + Dd);
            }
        }
        out[i] = result;
    }
}


/* ---- one whole sub-step as the wall-capable queue pass does it (k_fast<.., WALL = 1>, wall_reflect_on_path) ---------- */
typedef struct { float rx, ry, rz, dx, dy, dz, RD3, Dd, t_in; int c1, cur; } fm_walk;

static void fm_begin(fm_walk *w, const double *O, v3 P0, v3 d, int tet, int c1)
{
    w->rx = (float)(P0.x - O[0]); w->ry = (float)(P0.y - O[1]); w->rz = (float)(P0.z - O[2]);
    w->dx = (float)d.x; w->dy = (float)d.y; w->dz = (float)d.z;
    w->Dd = fmaxf(fmaxf(fabsf(w->dx), fabsf(w->dy)), fabsf(w->dz));
    w->RD3 = 3.f * (fmaxf(fmaxf(fabsf(w->rx), fabsf(w->ry)), fabsf(w->rz)) + w->Dd);
    w->t_in = 0.f; w->c1 = c1; w->cur = tet;
}

enum { FM_DONE = 0, FM_HOP = 1, FM_REFUSE = 2, FM_WALL = 3 };

static int fm_visit(const fm_rec *recs, const double *pos, v3 P0, fm_walk *w, float G, int *jsOut)
{
    const float INF = INFINITY;
    const fm_rec *f = recs + w->cur;
    float a[4], b[4], e[4];
    for (int j = 0; j < 3; ++j) {
        a[j] = w->rx * f->N[j][0] + w->ry * f->N[j][1] + w->rz * f->N[j][2];
        b[j] = w->dx * f->N[j][0] + w->dy * f->N[j][1] + w->dz * f->N[j][2];
    }
    const float V = f->V6, E = fabsf(f->E);
    a[3] = V - a[0] - a[1] - a[2];
    b[3] = -(b[0] + b[1] + b[2]);
    const float g = fmaf(G, V, 3.814697265625e-6f * (E * E) * (E + w->RD3));
    float amin = INF, eam = INF, emin = INF;
    for (int j = 0; j < 4; ++j) {
        e[j] = a[j] + b[j];
        amin = fminf(amin, a[j]); eam = fminf(eam, fabsf(e[j])); emin = fminf(emin, e[j]);
    }
    if (w->c1) {
        const float g0 = fmaf(G, V, 3.814697265625e-6f * (E * E) * (E + 3.f * fmaxf(fmaxf(fabsf(w->rx), fabsf(w->ry)), fabsf(w->rz))));
        if (!(amin >= g0)) return FM_REFUSE;
        w->c1 = 0;
    }
    if (emin >= g) return FM_DONE;
    if (!(eam >= g)) return FM_REFUSE;
    float t = INF;
    int js = -1;
    for (int j = 0; j < 4; ++j)
        if (e[j] < 0.f) {
            const float tj = a[j] * (1.0f / -b[j]);
            if (tj < t) { t = tj; js = j; }
        }
    if (js < 0) return FM_REFUSE;
    float c3m = INF;
    for (int j = 0; j < 4; ++j)
        if (j != js) c3m = fminf(c3m, fmaf(t, b[j], a[j]));
    if (!((c3m >= g) && (t > w->t_in) && (t <= 1.f))) return FM_REFUSE;
    *jsOut = js;
    const int link = f->link[js];
    if (link < 0) return FM_WALL;
    const int oldOrigin = f->origin;
    w->cur = link >> 2; w->t_in = t;
    f = recs + w->cur;
    if (f->origin != oldOrigin) {
        const double *O = pos + 3 * (long)f->origin;
        w->rx = (float)(P0.x - O[0]); w->ry = (float)(P0.y - O[1]); w->rz = (float)(P0.z - O[2]);
        w->RD3 = 3.f * (fmaxf(fmaxf(fabsf(w->rx), fabsf(w->ry)), fabsf(w->rz)) + w->Dd);
    }
    return FM_HOP;
}

/* cpf_geom.cuh exact_crossing: the reference's arithmetic for ONE certified face */
static int fm_exact_crossing(const orc_mesh *m, int tet, int f, v3 E, v3 *S, v3 *A, v3 *n)
{
    *n = face_inward_normal(m, f, tet, A);
    const v3 d = v3_sub(E, *S);
    const double fd = ref_dot(v3_sub(*A, *S), *n);
    const double dT = fd / ref_dot(d, *n);
    if (!(fd < ORC_TOL && dT > ORC_TOL && dT <= 1.0)) return 0;
    S->x = fma(d.x, dT, S->x); S->y = fma(d.y, dT, S->y); S->z = fma(d.z, dT, S->z);
    return 1;
}

/* status[i]: 0 refused (the exact kernel decides), 1 certified without wall contact, 2 certified with one in-place
 * reflection.  For status > 0: p, vel, tet hold the state after S5, comparable bit for bit with the oracle.
 * skipReplay != 0 (tests only): take the hit point from the sub-step's start instead of replaying the crossed faces --
 * what a wall handler WITHOUT the exact replay would compute. */
void orc_filter_substep(long n, double *p, const double *disp, double *vel, int *tet, const void *recsIn, MESH_ARGS, double guard,
                        int skipReplay, int *status)
{
    MESH_INIT;
    const fm_rec *recs = (const fm_rec *)recsIn;
    const float G = (float)guard * 1.0000002f;
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) {
        status[i] = 0;
        if (tet[i] < 0 || p[4 * i + 3] == 0.0) continue;
        const v3 P = { p[4 * i], p[4 * i + 1], p[4 * i + 2] }, dsp = { disp[4 * i], disp[4 * i + 1], disp[4 * i + 2] };
        v3 u = { vel[4 * i], vel[4 * i + 1], vel[4 * i + 2] };
        fm_walk w;
        fm_begin(&w, m.pos + 3 * (long)recs[tet[i]].origin, P, dsp, tet[i], 1);
        int pathTet[16], pathSlot[16], visits = 0, leg = 0, done = 0, refused = 0;
        v3 Phit = P, Eref = P;
        while (!done && !refused) {
            int js = -1;
            ++visits;
            const int before = w.cur;
            const int oc = fm_visit(recs, m.pos, leg ? Phit : P, &w, G, &js);
            if (oc == FM_DONE) done = 1;
            else if (oc == FM_HOP) {
                if (visits >= 48) refused = 1;
                else if (leg == 0 && visits <= 15) { pathTet[visits - 1] = before; pathSlot[visits - 1] = js; }
            } else if (oc == FM_WALL && leg == 0 && visits <= 15 && w.Dd < 10.f) {
                /* wall_reflect_on_path: exact replay of the crossed faces, then reflectInTet on the wall face */
                v3 E = v3_add(P, dsp), S = P, A, nrm;
                int ok = 1;
                for (int h = 0; h < visits - 1 && ok && !skipReplay; ++h) ok = fm_exact_crossing(&m, pathTet[h], recs[pathTet[h]].face[pathSlot[h]], E, &S, &A, &nrm);
                if (ok) ok = fm_exact_crossing(&m, w.cur, recs[w.cur].face[js], E, &S, &A, &nrm);
                if (ok) ok = fabs(ref_dot(v3_sub(A, S), nrm)) < ORC_TOL;
                if (!ok) { refused = 1; break; }
                Phit = S;
                const v3 r = v3_sub(E, A);
                double sp = -fma(r.z, nrm.z, fma(r.y, nrm.y, r.x * nrm.x));
                sp = sp + sp;
                double sv = -fma(u.z, nrm.z, fma(u.y, nrm.y, u.x * nrm.x));
                sv = sv + sv;
                E.x = fma(sp, nrm.x, E.x); E.y = fma(sp, nrm.y, E.y); E.z = fma(sp, nrm.z, E.z);
                u.x = fma(sv, nrm.x, u.x); u.y = fma(sv, nrm.y, u.y); u.z = fma(sv, nrm.z, u.z);
                Eref = E;
                const int wallTet = w.cur;
                fm_begin(&w, m.pos + 3 * (long)recs[wallTet].origin, Phit, v3_sub(Eref, Phit), wallTet, 0); /* hit point: certified by C3 */
                visits = 0;
                leg = 1;
            } else refused = 1;
        }
        if (refused) continue;
        v3 Pn;
        if (leg) { const v3 nd = v3_sub(Eref, Phit); Pn = v3_add(Phit, nd); }
        else Pn = v3_add(P, dsp);
        p[4 * i] = Pn.x; p[4 * i + 1] = Pn.y; p[4 * i + 2] = Pn.z;
        vel[4 * i] = u.x; vel[4 * i + 1] = u.y; vel[4 * i + 2] = u.z;
        tet[i] = w.cur;
        status[i] = 1 + leg;
    }
}

/* ---- RTX=true build: the fp32 walk towards the END POINT (cpf_geom.cuh visit_bary32) ------------------------------------
 * out[i] = tet certified to contain Q = P + disp, or -1 (refused: unclear minimum, boundary face, visit cap); to be compared
 * with the reference's baryTetSearch (s3_locate_bary). */
void orc_filter_bary_walk(long n, const double *p, const double *disp, const int *tet, const void *recsIn, const double *pos,
                          double guard, double errScale, int *out, int *visits)
{
    const fm_rec *recs = (const fm_rec *)recsIn;
    const float INF = INFINITY, G = (float)guard * 1.0000002f, ES = (float)errScale * 3.814697265625e-6f;
#pragma omp parallel for schedule(static)
    for (long i = 0; i < n; ++i) {
        out[i] = -1;
        visits[i] = 0;
        int cur = tet[i];
        if (cur < 0 || p[4 * i + 3] == 0.0) continue;
        const double Q[3] = { p[4 * i] + disp[4 * i], p[4 * i + 1] + disp[4 * i + 1], p[4 * i + 2] + disp[4 * i + 2] };
        const fm_rec *f = recs + cur;
        const double *O = pos + 3 * (long)f->origin;
        float rx = (float)(Q[0] - O[0]), ry = (float)(Q[1] - O[1]), rz = (float)(Q[2] - O[2]);
        float RD3 = 3.f * fmaxf(fmaxf(fabsf(rx), fabsf(ry)), fabsf(rz));
        for (int it = 0; it < 48; ++it) {
            visits[i]++;
            float e[4];
            for (int j = 0; j < 3; ++j) e[j] = rx * f->N[j][0] + ry * f->N[j][1] + rz * f->N[j][2];
            const float V = f->V6, E = fabsf(f->E);
            e[3] = V - e[0] - e[1] - e[2];
            const float g = fmaf(G, V, ES * (E * E) * (E + RD3));
            float m1 = e[0];
            int js = 0;
            for (int j = 1; j < 4; ++j) if (e[j] < m1) { m1 = e[j]; js = j; }
            if (m1 >= g) { out[i] = cur; break; }
            float m2 = INF;
            for (int j = 0; j < 4; ++j) if (j != js) m2 = fminf(m2, e[j]);
            if (!(m1 <= -g) || !(m2 - m1 >= 2.f * g)) break;
            const int link = f->link[js];
            if (link < 0 || it == 47) break;
            cur = link >> 2;
            const int oldOrigin = f->origin;
            f = recs + cur;
            if (f->origin != oldOrigin) {
                O = pos + 3 * (long)f->origin;
                rx = (float)(Q[0] - O[0]); ry = (float)(Q[1] - O[1]); rz = (float)(Q[2] - O[2]);
                RD3 = 3.f * fmaxf(fmaxf(fabsf(rx), fabsf(ry)), fabsf(rz));
            }
        }
    }
}

int orc_filter_rec_bytes(void) { return (int)sizeof(fm_rec); }
