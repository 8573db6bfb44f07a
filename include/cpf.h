/*
 * cpf.h -- C ABI of libcpf: B200-native particle advection on OpenFOAM tet-decomposed meshes.
 *
 * This is the drop-in boundary for the hot path of simzero/cudaParticlesFoam: everything the
 * solvers reach through src/initCuda.H and src/advect.H (textually included into main()) and that
 * today crosses into libcudaParticleAdvection.so as C++ free functions in namespace advect
 * (third_party/RTXAdvect/cuda/common.h:32-102, query/ConvexQuery.h:33-46, query/RTQuery.h:34-63).
 * Each entry point below names the reference interface it replaces (paths relative to
 * /root/reference).  extern "C", plain pointers and sizes only; the library owns all device memory
 * behind the handle; host buffers are borrowed for the duration of a call; every call returns an
 * int status (0 = CPF_OK) and never exits or throws (the reference's cudaCheck calls exit(),
 * third_party/RTXAdvect/cuda/cudaHelpers.cuh:32-40).
 */
#ifndef CPF_H
#define CPF_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CPF_ABI_VERSION 3

typedef struct cpf_context cpf_context;

enum cpf_status {
    CPF_OK = 0,
    CPF_ERR_INVALID = 1,      /* bad argument / call order                                      */
    CPF_ERR_CUDA = 2,         /* CUDA runtime error (message via cpf_last_error)                 */
    CPF_ERR_MESH = 3,         /* degenerate / inverted / non-manifold tet mesh                   */
    CPF_ERR_NOMEM = 4,
    CPF_ERR_NO_DEVICE = 5,    /* no CUDA device: there is NO CPU fallback                        */
    CPF_ERR_COMM = 6          /* NCCL not loadable / collective failed                            */
};

/* src/initCuda.H:72 VelocityInterpMethod ("TetVelocity" is what the glue uses;
 * "VertexVelocity" = cuda/particles.cu:244-313, barycentric interpolation of per-vertex values,
 * i.e. OpenFOAM cellPoint-style interpolation over [points..., cell centres...]) */
enum cpf_interp { CPF_INTERP_TET = 0, CPF_INTERP_VERTEX = 1 };
/* -DConvexPoly (default build, etc/bashrc:1 RTX=false) vs RTX=true build */
enum cpf_locator { CPF_LOCATOR_CONVEX = 0, CPF_LOCATOR_BARY = 1 };
/* time integration: the reference wires Euler only (cuda/particles.cu:358); RK2/RK4 are extensions */
enum cpf_integrator { CPF_EULER = 0, CPF_RK2 = 1, CPF_RK4 = 4 };
/* random walk generator: XORWOW = the reference's cuRAND stream (cuda/particles.cu:524-575),
 * PHILOX = stateless counter-based stream (0 B of state traffic), NONE = usingBrownianMotion=false */
enum cpf_rng { CPF_RNG_NONE = 0, CPF_RNG_XORWOW = 1, CPF_RNG_PHILOX = 2 };
/* per-patch boundary action; the reference reflects on every boundary (query/RTQuery.cu:165-166) */
enum cpf_patch_kind { CPF_PATCH_REFLECT = 0, CPF_PATCH_ESCAPE = 1 };
/* arithmetic policy of the locate step */
enum cpf_path {
    CPF_PATH_FILTERED = 0,    /* fast filtered predicates, exact reference arithmetic on demand   */
    CPF_PATH_EXACT = 1        /* reference arithmetic for every particle (cross-check mode)       */
};

/* Mirrors the dictionary keys and hard-coded switches of src/initCuda.H:50-72. */
typedef struct cpf_config {
    int device;               /* CUDA device ordinal                                             */
    int interp;               /* enum cpf_interp        (default CPF_INTERP_TET)                  */
    int locator;              /* enum cpf_locator       (default CPF_LOCATOR_CONVEX)              */
    int integrator;           /* enum cpf_integrator    (default CPF_EULER)                       */
    int rng;                  /* enum cpf_rng           (default CPF_RNG_XORWOW, as the glue)     */
    int reflect_wall;         /* src/initCuda.H:67 reflectWall (default 1)                        */
    int path;                 /* enum cpf_path          (default CPF_PATH_FILTERED)               */
    int sort_interval;        /* re-sort particles by cell every N sub-steps (0 = never; default 50) */
    int fuse_substeps;        /* max sub-steps fused into one launch sequence (0 = library default: 16,
                               * 10 with the XORWOW stream; clamped to 16 / 14)                      */
    double dt;                /* "dt" Lagrangian step (default 1e-4)                              */
    double diffusion_coeff;   /* "diffusionCoeff" (default 5.7e-6)                                */
    unsigned long long seed;  /* RNG seed (default 1591593751, cuda/particles.cu:544)             */
    int save_interval;        /* "saveInterval" (default 10)                                      */
    int reserved[7];
} cpf_config;

/* Counters gathered per GPU ("particle statistics come back by gather"). */
typedef struct cpf_stats {
    long long n_particles;
    long long n_active;       /* w != 0                                                          */
    long long n_negative_tet; /* what cudaReportParticles counts (cuda/particles.cu:763-775)      */
    long long n_escaped;      /* left through an ESCAPE patch (cumulative)                        */
    long long n_reflections;  /* wall hits handled (cumulative)                                   */
    long long n_exact;        /* particle-sub-steps that took the exact path (cumulative)         */
    long long n_hops;         /* tets visited by the locator (cumulative)                         */
    long long n_substeps;     /* particle-sub-steps executed (cumulative)                         */
    double kinetic_energy;    /* sum 0.5*|vel|^2, as printed by writeParticles2VTU (utils.cpp:258) */
    double reserved[3];
} cpf_stats;

/* -- lifecycle ------------------------------------------------------------------------------ */
int cpf_abi_version(void);
/* CUDA devices visible to this process (one rank per GPU: device = rank % count); CPF_ERR_NO_DEVICE if none */
int cpf_device_count(int *n);
void cpf_default_config(cpf_config *cfg);
/* replaces the ~35 locals + cudaMalloc block of src/initCuda.H:33-72, 141-150 */
int cpf_create(const cpf_config *cfg, cpf_context **out);
int cpf_destroy(cpf_context *ctx);
const char *cpf_last_error(const cpf_context *ctx); /* ctx may be NULL: last create() failure */
int cpf_sync(cpf_context *ctx);
/* Run all work of this context on a caller-owned CUDA stream (cudaStream_t), e.g. the stream the
 * host issues its NCCL collectives on, so the velocity broadcast and its first consumer kernel are
 * ordered without host synchronisation.  NULL restores the context's private stream.  The
 * reference runs everything on the default stream with a device-wide sync after every kernel. */
int cpf_set_stream(cpf_context *ctx, void *cuda_stream);
int cpf_set_config(cpf_context *ctx, const cpf_config *cfg); /* run-time switches only */

/* -- mesh upload ---------------------------------------------------------------------------- */
/* Replaces src/initCuda.H:76-130: tet decomposition of the fvMesh (polyMeshTetDecomposition::
 * cellTetIndices / tetIndices::faceTriIs), HostTetMesh::getBoundaryMesh (cuda/HostTetMesh.h:
 * 307-430), DeviceTetMesh::upload (cuda/DeviceTetMesh.cuh:59-72) and the OptiX BVH build.
 * Arrays are OpenFOAM's: points[nPoints][3], faces as CSR, owner[nFaces], neighbour[nInternal],
 * cellCentres[nCells][3] (mesh.C()), tetBasePt[nFaces] (mesh.tetBasePtIs(), NULL => 0),
 * patchStart[nPatches+1] face ranges of the boundary patches, patchKind[nPatches] (NULL => all
 * reflect).  Vertex ids of the device tet mesh are [points..., nPoints + cell]. */
int cpf_mesh_upload_poly(cpf_context *ctx, int nPoints, const double *points, int nFaces,
                         const int *faceOffsets, const int *faceVerts, const int *owner, int nInternal,
                         const int *neighbour, int nCells, const double *cellCentres, const int *tetBasePt,
                         int nPatches, const int *patchStart, const int *patchKind);
/* Per-patch rebound model (extension; the reference reflects specularly everywhere,
 * query/ConvexQuery.cu:287-295, query/RTQuery.cu:92-107, 165-166 TODO): restitution coefficient
 * e[p] in (0, 1] of every boundary patch; at a wall contact the normal part of the remaining
 * displacement and of the reported velocity is returned scaled by e (x' = x - (1 + e)(x.n) n).
 * e = 1 everywhere (the default) is the reference bit for bit.  Call after the mesh upload. */
int cpf_set_patch_restitution(cpf_context *ctx, int nPatches, const double *e);
/* Same for an explicit tet list (HostTetMesh{positions,indices}); tetCell may be NULL (tet == cell). */
int cpf_mesh_upload_tets(cpf_context *ctx, int nVerts, const double *positions, long long nTets,
                         const int *tetVerts, const int *tetCell, int nCells);
int cpf_mesh_info(cpf_context *ctx, long long *nVerts, long long *nTets, long long *nCells,
                  long long *nBoundaryFaces);
/* debug/parity: tet table in upload order */
int cpf_mesh_download_tets(cpf_context *ctx, int *tetVerts /*[nTets][4]*/, int *tetCell /*[nTets]*/);
/* debug/parity: per tet and reference face slot k (opposite vertex k): neighbour tet, or
 * -(patch+1) on the boundary -- the information content of tetfacets+faceInfos */
int cpf_mesh_download_neighbours(cpf_context *ctx, int *nbr /*[nTets][4]*/);

/* -- flow field ----------------------------------------------------------------------------- */
/* Replaces the host 12x expansion + cudaUpdateVelocity of src/advect.H:44-57 and
 * cuda/particles.cu:733-749.  U is the solver's cell field [nCells][3] (fp64).  on_device != 0:
 * U is a device pointer (e.g. the NCCL broadcast buffer); one kernel on the library's stream
 * repacks it into the idle half of the library's double buffer ((ux,uy,uz,0) per cell), after
 * which (in stream order) the caller's buffer may be overwritten.  on_device == 0: U is host
 * memory; upload and repack run on a copy stream, so they overlap sub-steps that are still
 * running, and the sub-steps enqueued afterwards wait for them.  Page-locked host memory is read asynchronously:
 * keep it unchanged until the next synchronising call (cpf_sync, cpf_stats_get, cpf_download). */
int cpf_update_velocity(cpf_context *ctx, const double *U, int on_device);
/* CPF_INTERP_VERTEX: explicit per-vertex field [nVerts][3] (points then centres); if never
 * called the library interpolates point values from the cell field (inverse-distance weights). */
int cpf_update_vertex_velocity(cpf_context *ctx, const double *Uvert, int on_device);

/* -- particles ------------------------------------------------------------------------------ */
/* cudaInitParticles (cuda/particles.cu:100-108) with an explicit, reproducible host-side stream */
int cpf_seed_box(cpf_context *ctx, long long n, const double lo[3], const double hi[3],
                 unsigned long long seed);
/* particles [first, first + count) of that stream (one rank per GPU: the index range this rank
 * tracks); also sets the particle id base to `first` */
int cpf_seed_box_slice(cpf_context *ctx, long long first, long long count, const double lo[3],
                       const double hi[3], unsigned long long seed);
/* cudaInitParticles(fileName) analogue: xyzw[n][4] = (x,y,z,active) */
int cpf_set_particles(cpf_context *ctx, long long n, const double *xyzw);
/* optional: caller-supplied start tets (skips cpf_locate_initial) */
int cpf_set_tets(cpf_context *ctx, const int *tet);
/* replaces RTQuery(OptixQuery&,...) = OptiX ray cast + baryQuery (query/RTQuery.cu:295-310):
 * BVH point location, lowest containing tet id, -1 outside */
int cpf_locate_initial(cpf_context *ctx);
/* Lost-particle fallback (extension): particles that are still active but carry a negative tet id
 * (e.g. after five failed reflections) are re-located with the BVH instead of being frozen on the
 * next sub-step as the reference does (cuda/particles.cu:334-338).  Call between cpf_substeps. */
int cpf_relocate_lost(cpf_context *ctx);
/* Continuous injection (extension; the reference seeds once, src/initCuda.H:141-150): every inactive
 * particle -- escaped through an ESCAPE patch or frozen outside the domain -- is re-seeded uniformly in
 * the box [lo, hi] (e.g. a slab behind the inlet), activated and located with the BVH.  Positions depend
 * on (seed, sub-step index, global particle id) only.  *nReseeded (may be NULL; non-NULL synchronises). */
int cpf_reseed_inactive(cpf_context *ctx, const double lo[3], const double hi[3], unsigned long long seed,
                        long long *nReseeded);
/* initRandomGenerator (cuda/particles.cu:541-548) */
int cpf_init_rng(cpf_context *ctx);

/* -- the hot path --------------------------------------------------------------------------- */
/* One body of src/advect.H:33-184 without the velocity refresh: nCycles = max(ceil(deltaT/dt),1)
 * sub-steps of {cudaAdvect, cudaBrownianMotion, convexTetQuery|RTQuery, convexWallReflect|
 * RTWallReflect, cudaMoveParticles}, fused.  Asynchronous; *nCyclesOut may be NULL. */
int cpf_advect(cpf_context *ctx, double deltaT, int *nCyclesOut);
/* exactly n sub-steps of size dt (the loop body at src/advect.H:86-184) */
int cpf_substeps(cpf_context *ctx, int n, double dt);
/* the initial cudaAdvect of src/initCuda.H:184-199 (deactivates out-of-domain particles) */
int cpf_initial_advect(cpf_context *ctx);
/* force a sort-by-cell now */
int cpf_sort_particles(cpf_context *ctx);
/* last kernel timing: milliseconds of the most recent cpf_advect/cpf_substeps on the device */
int cpf_last_step_ms(cpf_context *ctx, float *ms);

/* -- one rank per GPU ---------------------------------------------------------------- */
/* Replaces the gather-to-master parallel branch of src/initCuda.H:207-270 and src/advect.H:59-89
 * (Pstream::gatherList of points/cells/velocities to rank 0, ONE GPU for the whole MPI job;
 * third_party/RTXAdvect/optix/OptixTetQuery.cpp:154 hard-codes device 0): every rank owns a context on
 * its own GPU, tracks an index range of the particle cloud on a replica of the mesh, and the two
 * per-step exchanges run over NCCL on the context's stream.  NCCL is bound at run time
 * (dlopen libnccl.so.2); a communicator of one rank needs no NCCL.
 * cpf_comm_unique_id: rank 0 creates the id (ncclGetUniqueId) and the HOST distributes its bytes with
 * whatever it has (Pstream::broadcast / MPI_Bcast / a file); CPF_COMM_ID_BYTES bytes. */
#define CPF_COMM_ID_BYTES 128
int cpf_comm_unique_id(void *id, size_t bytes);
int cpf_comm_init(cpf_context *ctx, const void *id, size_t bytes, int rank, int nranks);
int cpf_comm_info(cpf_context *ctx, int *rank, int *nranks, int *ncclVersion);
/* The coupled solver's per-step field: `root` passes the whole cell field (host or device memory), the
 * others pass NULL; upload (root), ncclBroadcast and repack run on the library's copy stream with their
 * own communicator, beside the sub-steps already enqueued (src/advect.H:44-57 + 59-89).  on_device: 0 = host
 * memory, 1 = device memory that is ready in the order of the context's stream (the exchange then starts
 * after the work enqueued so far), 2 = device memory whose contents are complete already (no ordering:
 * the exchange of step k+1 overlaps the sub-steps of step k). */
int cpf_update_velocity_bcast(cpf_context *ctx, const double *U, int on_device, int root);
/* Decomposed solver runs: every rank passes only the nLocal cells it owns, as global
 * cell ids [cellOffset, cellOffset + nLocal) of the replicated mesh; the slices are exchanged
 * rank-to-rank (grouped ncclBroadcast), nothing is gathered to a master. */
int cpf_update_velocity_slices(cpf_context *ctx, long long cellOffset, long long nLocal, const double *Ulocal, int on_device);
/* Global id of this context's particle 0.  The random-walk streams are keyed by GLOBAL particle id
 * (Philox counter, XORWOW subsequence = curand_init(seed, id), cuda/particles.cu:537), so N ranks
 * tracking index ranges of one cloud reproduce the single-GPU run bit for bit.  Call before
 * cpf_init_rng / the first sub-step. */
int cpf_set_particle_id_base(cpf_context *ctx, long long base);

/* -- results -------------------------------------------------------------------------------- */
/* replaces the D2H copies of writeParticles2VTU (cuda/utils.cpp:144-170); any pointer may be
 * NULL; arrays are in ORIGINAL particle order: xyzw[n][4], vel[n][4], tet[n] */
int cpf_download(cpf_context *ctx, double *xyzw, double *vel, int *tet);
int cpf_download_cells(cpf_context *ctx, int *cell);
/* cudaReportParticles + the kinetic-energy print of writeParticles2VTU: full scan, blocks until it has arrived
 * (= cpf_stats_request(ctx, 1) + cpf_stats_collect).  With a communicator: summed over the ranks (collective). */
int cpf_stats_get(cpf_context *ctx, cpf_stats *out);
/* The same without stalling the pipeline: cpf_stats_request enqueues the read-back behind everything
 * submitted so far and returns at once (up to 4 outstanding); cpf_stats_collect waits for the OLDEST
 * outstanding request only -- ask for step k's numbers after submitting step k+1.  full != 0 scans the
 * particle arrays (exact n_active, n_negative_tet, kinetic energy); full == 0 moves the 11 cumulative
 * counters only: n_active is derived from them (every particle that stops being active is counted as an
 * escape or as a freeze), n_negative_tet and kinetic_energy repeat the last full scan.  reserved[0] = 1
 * for a full result. */
int cpf_stats_request(cpf_context *ctx, int full);
int cpf_stats_collect(cpf_context *ctx, cpf_stats *out);
/* writeParticles2VTU (cuda/utils.cpp:144-283): particle_%04d.vtu in `dir` */
int cpf_write_vtu(cpf_context *ctx, const char *dir, unsigned step);
/* The same file name, arrays and array names, but written without stalling the advection
 * (replaces the blocking copies + ASCII printing of cuda/utils.cpp:144-283 and
 * the sync it forces in src/advect.H:163-175): every `stride`-th particle (original ids
 * 0, stride, 2*stride, ...) is packed on the device, copied on the copy stream into one of two
 * page-locked buffers and written by a writer thread as VTK XML with one raw appended-data
 * section (little endian, UInt64 block headers).  Returns once the work is enqueued; blocks only
 * while two earlier files are still being written. */
int cpf_write_vtu_async(cpf_context *ctx, const char *dir, unsigned step, int stride);
/* Blocks until every file handed to cpf_write_vtu_async is on disk and returns the writer's first
 * I/O error, if any (cpf_sync and cpf_destroy drain the writer too). */
int cpf_output_wait(cpf_context *ctx);
/* Checkpoint / restart (the reference has none, src/initCuda.H:498): particle state
 * in original order, the global sub-step index (the Philox counter), cumulative counters and, in
 * XORWOW mode, the generator states.  cpf_checkpoint_load needs the same mesh uploaded and the
 * same random-walk configuration; the run then continues bit-identically. */
int cpf_checkpoint_save(cpf_context *ctx, const char *path);
/* number of sub-steps executed since seeding (restored by cpf_checkpoint_load): the `step` counter of
 * src/initCuda.H:498 and the Philox counter */
unsigned long long cpf_step_index(cpf_context *ctx);
int cpf_checkpoint_load(cpf_context *ctx, const char *path);
long long cpf_num_particles(cpf_context *ctx);
/* raw device pointers (for torch / NCCL plumbing); valid until the next set/seed/sort call;
 * ucell is the library's current field buffer, (ux,uy,uz,0) per cell */
int cpf_device_pointers(cpf_context *ctx, void **pos4, void **tet, void **ucell);
/* parity tooling: the normal deviates the next sub-step will use, [n][3], original order;
 * does not advance the stream */
int cpf_debug_next_normals(cpf_context *ctx, double *xi);
/* the same for the next k sub-steps, xi[k][n][3]: lets a test replay a FUSED chunk through the oracle */
int cpf_debug_normals(cpf_context *ctx, int k, double *xi);
/* Device timing of the fused sub-step kernel: while enabled, a CUDA-event pair brackets every
 * launch on the launching stream; cpf_profile_read returns and resets the totals (replaces the
 * commented-out Adv/Dfs/Qry/Rft/Mov breakdown of src/advect.H:186-203). */
int cpf_profile_enable(cpf_context *ctx, int enable);
int cpf_profile_read(cpf_context *ctx, int *nLaunches, double *total_ms, double *max_ms);
/* number of kernels the library launched since create (bench.py gpu_launches) */
long long cpf_launch_count(cpf_context *ctx);

#ifdef __cplusplus
}
#endif
#endif /* CPF_H */
