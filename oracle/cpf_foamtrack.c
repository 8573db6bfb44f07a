/* cpf_foamtrack.c -- TEST / BENCH INFRASTRUCTURE ONLY (included by cpf_oracle.c).
 *
 * "FoamTrack": the CPU tracking OpenFOAM itself would do for this job, restated from the published
 * algorithm.  OpenFOAM v2106 (src/lagrangian/basic/particle/particle.C, trackToTri /
 * trackToStationaryTri) is a dependency of the reference that is NOT under /root/reference and
 * cannot be built here; its idea is: a particle lives in a tet of the cell decomposition with
 * barycentric coordinates; a straight move changes them linearly, the first coordinate to reach
 * zero names the face that is hit; the particle is put on that face, changes tet (or is reflected
 * at a wall patch) and continues with the remaining fraction of the step.
 *
 * This is the second CPU column of SURVEY 8(d) ("OpenFOAM-style CPU tracking (restated; OpenFOAM
 * not available)"), timed on 1 core and on all cores next to the oracle port.  It is NOT a parity
 * oracle -- its arithmetic differs from the reference's -- and parity is unpinned against OpenFOAM;
 * tests check it against the oracle up to rounding (same cells except ties, positions to 1e-9). */
#ifdef _OPENMP
#include <omp.h>
#endif

static inline void ft_bary_frame(const orc_mesh *m, int tet, v3 *A, double inv[9], int *ok)
{
    const int *id = m->idx + 4 * tet;
    const v3 a = ld3(m->pos, id[0]), b = ld3(m->pos, id[1]), c = ld3(m->pos, id[2]), d = ld3(m->pos, id[3]);
    const v3 e1 = v3_sub(b, a), e2 = v3_sub(c, a), e3 = v3_sub(d, a);
    /* inverse of [e1 e2 e3] by cofactors: rows are the reciprocal vectors */
    const v3 r1 = { e2.y * e3.z - e2.z * e3.y, e2.z * e3.x - e2.x * e3.z, e2.x * e3.y - e2.y * e3.x };
    const v3 r2 = { e3.y * e1.z - e3.z * e1.y, e3.z * e1.x - e3.x * e1.z, e3.x * e1.y - e3.y * e1.x };
    const v3 r3 = { e1.y * e2.z - e1.z * e2.y, e1.z * e2.x - e1.x * e2.z, e1.x * e2.y - e1.y * e2.x };
    const double det = e1.x * r1.x + e1.y * r1.y + e1.z * r1.z;
    *ok = det != 0.0;
    const double s = *ok ? 1.0 / det : 0.0;
    inv[0] = r1.x * s; inv[1] = r1.y * s; inv[2] = r1.z * s;
    inv[3] = r2.x * s; inv[4] = r2.y * s; inv[5] = r2.z * s;
    inv[6] = r3.x * s; inv[7] = r3.y * s; inv[8] = r3.z * s;
    *A = a;
}

/* barycentric image of a vector (w[0] belongs to vertex 0): for a point pass x - A and set isPoint */
static inline void ft_to_bary(const double inv[9], v3 r, int isPoint, double w[4])
{
    w[1] = inv[0] * r.x + inv[1] * r.y + inv[2] * r.z;
    w[2] = inv[3] * r.x + inv[4] * r.y + inv[5] * r.z;
    w[3] = inv[6] * r.x + inv[7] * r.y + inv[8] * r.z;
    w[0] = (isPoint ? 1.0 : 0.0) - w[1] - w[2] - w[3];
}

/* returns the number of faces crossed or hit */
static int ft_track(double *p, int *tetIO, v3 d, const orc_mesh *m, int reflectWall)
{
    int tet = *tetIO, events = 0;
    v3 x = { p[0], p[1], p[2] };
    double remaining = 1.0; /* fraction of d still to travel */
    for (int it = 0; it < 200 && remaining > 0.0; ++it) {
        v3 A;
        double inv[9], lam[4], mu[4];
        int ok;
        ft_bary_frame(m, tet, &A, inv, &ok);
        if (!ok) break;
        ft_to_bary(inv, v3_sub(x, A), 1, lam);
        const v3 dr = { d.x * remaining, d.y * remaining, d.z * remaining };
        ft_to_bary(inv, dr, 0, mu);
        double s = 1.0;
        int hit = -1;
        for (int k = 0; k < 4; ++k) {
            if (mu[k] < 0.0) {
                const double lk = lam[k] > 0.0 ? lam[k] : 0.0;
                const double sk = -lk / mu[k];
                if (sk < s) { s = sk; hit = k; }
            }
        }
        x.x += s * dr.x; x.y += s * dr.y; x.z += s * dr.z;
        if (hit < 0) { remaining = 0.0; break; }
        remaining *= (1.0 - s);
        events++;
        const int f = m->tetfacets[4 * tet + hit];
        const int nbr = other_tet(m->finfo, f, tet);
        if (nbr >= 0) { tet = nbr; continue; }
        if (!reflectWall) { tet = -(tet + 1); break; }
        /* wall patch: specular reflection of the direction of travel, the particle stays in this tet */
        v3 Af;
        const v3 n = face_inward_normal(m, f, tet, &Af);
        const double dn = d.x * n.x + d.y * n.y + d.z * n.z;
        d.x -= 2.0 * dn * n.x; d.y -= 2.0 * dn * n.y; d.z -= 2.0 * dn * n.z;
    }
    p[0] = x.x; p[1] = x.y; p[2] = x.z;
    *tetIO = tet;
    return events;
}

/* nSteps Euler sub-steps with the per-tet velocity of the reference (vel = Utet[tet]); nThreads <= 0: all cores.
 * Returns the number of face events (crossings + wall hits). */
long orc_foamtrack_substeps(long n, int nSteps, double *p, int *tet, double dt, MESH_ARGS, const double *Utet,
                            int reflectWall, int nThreads)
{
    MESH_INIT;
    long events = 0;
#ifdef _OPENMP
    if (nThreads <= 0) nThreads = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 256) reduction(+ : events) num_threads(nThreads)
#endif
    for (long i = 0; i < n; ++i) {
        double *pi = p + 4 * i;
        for (int s = 0; s < nSteps; ++s) {
            if (pi[3] == 0.0 || tet[i] < 0) break;
            const double *u = Utet + 3 * (size_t)tet[i];
            const v3 d = { dt * u[0], dt * u[1], dt * u[2] };
            events += ft_track(pi, tet + i, d, &m, reflectWall);
        }
    }
    return events;
}
