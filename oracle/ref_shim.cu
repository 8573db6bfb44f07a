// ref_shim.cu -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// Thin extern "C" doorway onto the UNMODIFIED reference kernels.  The reference's three CUDA
// translation units (cuda/particles.cu, query/ConvexQuery.cu, query/RTQuery.cu) are compiled
// straight from /root/reference by oracle/Makefile into oracle/_ref/libref_rtxadvect.so together
// with this file; this file only (a) supplies link-time stubs for the OptiX-backed OptixQuery
// members that cannot be built without the OptiX SDK, and (b) forwards to the reference's own
// host functions (namespace advect) in exactly the order /root/reference/src/advect.H:96-161
// calls them.  No reference source is copied.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "cuda/common.h"
#include "cuda/DeviceTetMesh.cuh"
#include "cuda/HostTetMesh.h"
#include "query/ConvexQuery.h"
#include "query/RTQuery.h"

namespace advect {
// ---- link-time stubs for optix/OptixTetQuery.cpp / OptixTriQuery.cpp (need the OptiX 7 SDK) ----
// The seeding broad phase is replaced by ids pre-filled by the caller (brute-force fp64
// containment); query_sync therefore leaves out_tetIDs untouched and RTQuery() proceeds to the
// reference's own baryQuery narrow phase.
void OptixQuery::initSystem(const double4 *, int, const int4 *, int) {}
void OptixQuery::initBoundarySystem(const double4 *, int, const int4 *, int) {}
void OptixQuery::query_sync(float4 *, int *, int) {}
void OptixQuery::query_sync(double4 *, int *, int) {}
void OptixQuery::query_disp(double4 *, double4 *, int *, int) {}
void OptixQuery::query_disp_Bd(double4 *, double4 *, int *, int) {}
} // namespace advect

using namespace advect;

struct RefMesh {
    DeviceTetMesh dev;
    int nVerts = 0, nTets = 0, nFaces = 0;
};

struct RefParticles {
    int n = 0;
    Particle *d_particles = nullptr;
    int *d_tetIDs = nullptr;
    vec4d *d_disps = nullptr;
    vec4d *d_vels = nullptr;
    curandState_t *rand_states = nullptr;
};

template <typename T> static void up(T *&d, const void *h, size_t count)
{
    cudaCheck(cudaMalloc(&d, count * sizeof(T)));
    cudaCheck(cudaMemcpy(d, h, count * sizeof(T), cudaMemcpyHostToDevice));
}

extern "C" {

int ref_abi_version() { return 1; }

// The reference's own host face-topology builder (cuda/HostTetMesh.h:307-430).  Valid only for
// nVerts < 2^20 (20-bit key packing).  Prints one line per boundary vertex to stdout like the
// original (HostTetMesh.h:389) -- callers redirect stdout.
int ref_build_faces(int nVerts, const double *pos, int nTets, const int *idx,
                    int *tetfacets, int *facets, int *finfo, int maxFaces)
{
    HostTetMesh m;
    m.positions.resize(nVerts);
    std::memcpy((void *)m.positions.data(), pos, sizeof(double) * 3 * nVerts);
    m.indices.resize(nTets);
    std::memcpy((void *)m.indices.data(), idx, sizeof(int) * 4 * nTets);
    FILE *saved = stdout; (void)saved;
    std::cout.setstate(std::ios_base::failbit); // silence the per-vertex print
    HostTetMesh bd = m.getBoundaryMesh();
    std::cout.clear();
    int nFaces = (int)m.facets.size();
    if (nFaces > maxFaces || (int)m.tetfacets.size() != nTets) return -(int)m.tetfacets.size() - 1;
    std::memcpy(tetfacets, m.tetfacets.data(), sizeof(int) * 4 * nTets);
    std::memcpy(facets, m.facets.data(), sizeof(int) * 4 * nFaces);
    std::memcpy(finfo, m.faceInfos.data(), sizeof(int) * 2 * nFaces);
    return nFaces;
}

// Same as DeviceTetMesh::upload (cuda/DeviceTetMesh.cuh:59-72) but from caller-built tables, so
// meshes past the reference builder's 2^20-vertex limit can still be run through its kernels.
void *ref_mesh_upload(int nVerts, const double *pos, int nTets, const int *idx, const double *Utet,
                      int nFaces, const int *facets, const int *tetfacets, const int *finfo)
{
    RefMesh *m = new RefMesh;
    m->nVerts = nVerts; m->nTets = nTets; m->nFaces = nFaces;
    up(m->dev.d_positions, pos, nVerts);
    up(m->dev.d_velocities, Utet, nTets);
    up(m->dev.d_indices, idx, nTets);
    up(m->dev.d_facets, facets, nFaces);
    up(m->dev.d_tetfacets, tetfacets, nTets);
    up(m->dev.d_faceinfos, finfo, nFaces);
    return m;
}

void ref_mesh_free(void *h)
{
    RefMesh *m = (RefMesh *)h;
    cudaFree(m->dev.d_positions); cudaFree(m->dev.d_velocities); cudaFree(m->dev.d_indices);
    cudaFree(m->dev.d_facets); cudaFree(m->dev.d_tetfacets); cudaFree(m->dev.d_faceinfos);
    delete m;
}

// src/advect.H:44-57 velocity refresh: host vector (one vec3d per tet) passed BY VALUE.
void ref_update_velocity(void *h, const double *Utet)
{
    RefMesh *m = (RefMesh *)h;
    std::vector<vec3d> v(m->nTets);
    std::memcpy((void *)v.data(), Utet, sizeof(double) * 3 * m->nTets);
    cudaUpdateVelocity(v, m->nTets, m->dev.d_indices, m->dev.d_velocities);
}

// src/initCuda.H:141-152 allocations (+ caller-provided positions / seed tets)
void *ref_particles_create(int n, const double *p4, const int *tet, int initRng)
{
    RefParticles *P = new RefParticles;
    P->n = n;
    up(P->d_particles, p4, n);
    up(P->d_tetIDs, tet, n);
    cudaCheck(cudaMalloc(&P->d_disps, n * sizeof(vec4d)));
    cudaCheck(cudaMemset(P->d_disps, 0, n * sizeof(vec4d)));
    cudaCheck(cudaMalloc(&P->d_vels, n * sizeof(vec4d)));
    cudaCheck(cudaMemset(P->d_vels, 0, n * sizeof(vec4d)));
    cudaCheck(cudaMalloc(&P->rand_states, n * sizeof(curandState_t)));
    if (initRng) {
        std::fflush(stdout);
        initRandomGenerator(n, P->rand_states); // cuda/particles.cu:541-548 (seed 1591593751)
    }
    cudaCheck(cudaDeviceSynchronize());
    return P;
}

void ref_particles_free(void *h)
{
    RefParticles *P = (RefParticles *)h;
    cudaFree(P->d_particles); cudaFree(P->d_tetIDs); cudaFree(P->d_disps); cudaFree(P->d_vels);
    cudaFree(P->rand_states);
    delete P;
}

void ref_particles_set(void *h, const double *p4, const int *tet)
{
    RefParticles *P = (RefParticles *)h;
    cudaCheck(cudaMemcpy(P->d_particles, p4, P->n * sizeof(Particle), cudaMemcpyHostToDevice));
    cudaCheck(cudaMemcpy(P->d_tetIDs, tet, P->n * sizeof(int), cudaMemcpyHostToDevice));
    cudaCheck(cudaMemset(P->d_disps, 0, P->n * sizeof(vec4d)));
    cudaCheck(cudaMemset(P->d_vels, 0, P->n * sizeof(vec4d)));
}

void ref_download(void *h, double *p4, int *tet, double *disp4, double *vel4)
{
    RefParticles *P = (RefParticles *)h;
    cudaCheck(cudaDeviceSynchronize());
    if (p4) cudaCheck(cudaMemcpy(p4, P->d_particles, P->n * sizeof(Particle), cudaMemcpyDeviceToHost));
    if (tet) cudaCheck(cudaMemcpy(tet, P->d_tetIDs, P->n * sizeof(int), cudaMemcpyDeviceToHost));
    if (disp4) cudaCheck(cudaMemcpy(disp4, P->d_disps, P->n * sizeof(vec4d), cudaMemcpyDeviceToHost));
    if (vel4) cudaCheck(cudaMemcpy(vel4, P->d_vels, P->n * sizeof(vec4d), cudaMemcpyDeviceToHost));
}

void ref_upload_disp(void *h, const double *disp4)
{
    RefParticles *P = (RefParticles *)h;
    cudaCheck(cudaMemcpy(P->d_disps, disp4, P->n * sizeof(vec4d), cudaMemcpyHostToDevice));
}

// ---- one wrapper per reference host function on the path --------------------------------------
void ref_advect(void *mh, void *ph, double dt, int vertexVelocity)
{
    RefMesh *m = (RefMesh *)mh; RefParticles *P = (RefParticles *)ph;
    cudaAdvect(P->d_particles, P->d_tetIDs, P->d_vels, P->d_disps, dt, P->n,
               m->dev.d_indices, m->dev.d_positions, m->dev.d_velocities,
               vertexVelocity ? "VertexVelocity" : "TetVelocity");
}
void ref_brownian(void *ph, double dt, double D)
{
    RefParticles *P = (RefParticles *)ph;
    cudaBrownianMotion(P->d_particles, P->d_disps, P->rand_states, dt, P->n, D);
}
void ref_locate_convex(void *mh, void *ph)
{
    RefMesh *m = (RefMesh *)mh; RefParticles *P = (RefParticles *)ph;
    convexTetQuery(m->dev, P->d_particles, P->d_disps, P->d_tetIDs, P->n);
}
void ref_reflect_convex(void *mh, void *ph)
{
    RefMesh *m = (RefMesh *)mh; RefParticles *P = (RefParticles *)ph;
    convexWallReflect(m->dev, P->d_tetIDs, P->d_particles, P->d_vels, P->d_disps, P->n);
}
void ref_locate_bary(void *mh, void *ph)
{
    RefMesh *m = (RefMesh *)mh; RefParticles *P = (RefParticles *)ph;
    RTQuery(m->dev, P->d_particles, P->d_disps, P->d_tetIDs, P->n);
}
void ref_reflect_bary(void *mh, void *ph)
{
    RefMesh *m = (RefMesh *)mh; RefParticles *P = (RefParticles *)ph;
    RTWallReflect(m->dev, P->d_tetIDs, P->d_particles, P->d_disps, P->d_vels, P->n);
}
void ref_move(void *ph)
{
    RefParticles *P = (RefParticles *)ph;
    cudaMoveParticles(P->d_particles, P->d_disps, P->n, P->d_tetIDs);
}
// query/RTQuery.cu:295-310 with the OptiX broad phase stubbed: ids must be pre-seeded.
void ref_bary_query(void *mh, void *ph)
{
    RefMesh *m = (RefMesh *)mh; RefParticles *P = (RefParticles *)ph;
    OptixQuery q((const double4 *)nullptr, 0, (const int4 *)nullptr, 0);
    RTQuery(q, m->dev, P->d_particles, P->d_tetIDs, P->n);
}

// The sub-step loop body of src/advect.H:86-184 (master rank), verbatim call order.
void ref_substeps(void *mh, void *ph, int nSteps, double cycleDt, int convexPoly,
                  int usingBrownianMotion, double diffusionCoeff, int reflectWall, int vertexVelocity)
{
    for (int i = 0; i < nSteps; ++i) {
        ref_advect(mh, ph, cycleDt, vertexVelocity);
        if (usingBrownianMotion) ref_brownian(ph, cycleDt, diffusionCoeff);
        if (!convexPoly) {
            ref_locate_bary(mh, ph);
            if (reflectWall) ref_reflect_bary(mh, ph);
        } else {
            ref_locate_convex(mh, ph);
            if (reflectWall) ref_reflect_convex(mh, ph);
        }
        ref_move(ph);
    }
}

// Timed variant for bench.py --impl reference: CUDA events around the same loop.
float ref_substeps_timed(void *mh, void *ph, int nSteps, double cycleDt, int convexPoly,
                         int usingBrownianMotion, double diffusionCoeff, int reflectWall)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a, 0);
    ref_substeps(mh, ph, nSteps, cycleDt, convexPoly, usingBrownianMotion, diffusionCoeff, reflectWall, 0);
    cudaEventRecord(b, 0);
    cudaEventSynchronize(b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    cudaEventDestroy(a); cudaEventDestroy(b);
    return ms;
}

// Normal deviates exactly as particleBrownianMotion draws them (three successive
// curand_normal_double on each particle's XORWOW state); advances the states.
__global__ void ref_draw_normals_kernel(curandState_t *states, double *xi, int n)
{
    int i = threadIdx.x + blockDim.x * blockIdx.x;
    if (i >= n) return;
    xi[3 * i + 0] = curand_normal_double(&states[i]);
    xi[3 * i + 1] = curand_normal_double(&states[i]);
    xi[3 * i + 2] = curand_normal_double(&states[i]);
}
void ref_draw_normals(void *ph, double *xi_host)
{
    RefParticles *P = (RefParticles *)ph;
    double *d = nullptr;
    cudaCheck(cudaMalloc(&d, sizeof(double) * 3 * P->n));
    ref_draw_normals_kernel<<<(P->n + 127) / 128, 128>>>(P->rand_states, d, P->n);
    cudaCheck(cudaDeviceSynchronize());
    cudaCheck(cudaMemcpy(xi_host, d, sizeof(double) * 3 * P->n, cudaMemcpyDeviceToHost));
    cudaFree(d);
}
void ref_rng_state_download(void *ph, void *out) // n * sizeof(curandState_t) (48 B each)
{
    RefParticles *P = (RefParticles *)ph;
    cudaCheck(cudaMemcpy(out, P->rand_states, P->n * sizeof(curandState_t), cudaMemcpyDeviceToHost));
}
void ref_rng_state_upload(void *ph, const void *in)
{
    RefParticles *P = (RefParticles *)ph;
    cudaCheck(cudaMemcpy(P->rand_states, in, P->n * sizeof(curandState_t), cudaMemcpyHostToDevice));
}

} // extern "C"
