// cpf_internal.h -- shared declarations of libcpf's translation units (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <curand_kernel.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/cpf.h"
#include "cpf_geom.cuh"

namespace cpf {

enum Counter { CNT_ESCAPED = 0, CNT_REFLECT, CNT_EXACT, CNT_HOPS, CNT_SUBSTEPS, CNT_LOST, CNT_FROZEN, CNT_COUNT = 8 };

struct ParticleView {
    double4 *pos;        // [n] x,y,z,w (w != 0 => active), reference layout (cuda/common.h:26)
    int *tet;            // [n]
    int *pid;            // [n] original particle index (identity until the first sort)
    double4 *vel;        // [n] vx,vy,vz,-1 (written on the last fused sub-step when requested)
    curandState_t *rng;  // [n] XORWOW states (CPF_RNG_XORWOW)
    long long n;
};

struct StepParams {
    int nSub;
    double dt;
    double randDisp;     // sqrt(2*D*dt), cuda/particles.cu:564
    int reflect;
    int writeVel;
    int integrator, interp;
    unsigned long long seed;
    unsigned long long step0; // global sub-step index of the first fused sub-step (Philox counter)
    unsigned long long idBase; // global id of this context's particle 0 (Philox counter / XORWOW subsequence)
    unsigned long long *counters;
    int2 *queueIn, *queueOut;          // deferral queues: (particle slot, sub-step to resume at)
    unsigned *countIn, *countOut;
};

// BVH over tets for initial / lost-particle location (cpf_locate.cu)
struct BvhLevel { float4 *lo; float4 *hi; long long n; };

} // namespace cpf

namespace cpf { struct OutputState; struct CommState; }

struct cpf_context {
    cpf_config cfg;
    int device = 0;
    cudaStream_t stream = nullptr;      // stream in use (private or caller-owned)
    cudaStream_t ownStream = nullptr;   // the context's private stream
    cudaStream_t copyStream = nullptr;
    bool profiling = false;
    std::vector<cudaEvent_t> profEvents; // pairs
    size_t profUsed = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, evCopy = nullptr;
    cudaEvent_t evRead[2] = { nullptr, nullptr }; // per velocity buffer: end of the kernels that read it
    std::string err;
    float last_ms = 0.f;
    long long launches = 0;

    // mesh
    bool have_mesh = false;
    long long nVerts = 0, nTets = 0, nCells = 0, nBoundaryFaces = 0;
    int nPoints = 0, nPatches = 0;
    bool cellFromVertex = false;
    double4 *d_vpos = nullptr;
    int4 *d_tetv = nullptr;      // sorted ids
    double4 *d_tetnrm = nullptr; // [nTets][3]: the four reference face normals per tet (96 B)
    uint4 *d_tetfast = nullptr;  // [nTets][4]: 64-byte fp32 record of the fast walk
    int4 *d_tetrec = nullptr;    // [nTets][2]: {links, apex vertex id of the neighbour across each face}
    uint16_t *d_tetcode = nullptr;
    int *d_tetcell = nullptr;
    double4 *d_ucell[2] = { nullptr, nullptr }; // double-buffered cell field, (ux,uy,uz,0) per cell
    double *d_ustage = nullptr;                 // host uploads land here in the solver's layout [nCells][3] and are repacked
    int ucur = 0;
    double *d_uvert = nullptr;
    int *d_pc_off = nullptr, *d_pc_cells = nullptr; // point -> cells CSR (vertex interpolation from the cell field)
    uint8_t *d_patch_kind = nullptr;
    double *d_patch_gain = nullptr; // [nPatches] 1 + restitution coefficient (2 = the reference's specular reflection)
    double guard = 1e-7, hmin = 0.0;
    bool filter_ok = true; // coordinates small enough against the smallest tet for the filtered walk (cpf_mesh.cu)
    double bbox_lo[3] = { 0, 0, 0 }, bbox_hi[3] = { 0, 0, 0 };
    double *h_pinned = nullptr;
    size_t pinned_bytes = 0;

    // BVH (sorted tet order + 8-ary implicit tree of float boxes)
    int *d_bvh_tet = nullptr;                 // [nTets] tet ids in Morton order
    float4 *d_bvh_tet_lo = nullptr, *d_bvh_tet_hi = nullptr; // box of every tet, in the order of d_bvh_tet
    std::vector<cpf::BvhLevel> bvh;           // level 0 = groups of 8 tets
    float4 *d_bvh_top_lo = nullptr, *d_bvh_top_hi = nullptr; // concatenated top levels (staged in smem)
    int bvh_top_first_level = 0, bvh_top_nodes = 0;
    std::vector<int> bvh_top_offsets;

    // particles (double-buffered for sort-by-cell)
    long long n = 0;
    double4 *d_pos[2] = { nullptr, nullptr };
    int *d_tet[2] = { nullptr, nullptr };
    int *d_pid[2] = { nullptr, nullptr };
    double4 *d_vel[2] = { nullptr, nullptr };
    curandState_t *d_rng[2] = { nullptr, nullptr };
    int pcur = 0;
    bool permuted = false;
    bool rng_ready = false;
    bool have_tets = false;
    unsigned long long step_index = 0; // global sub-step counter
    unsigned long long id_base = 0;    // global id of particle 0 (cpf_set_particle_id_base): random-walk streams are keyed by global id
    cpf::CommState *comm = nullptr;    // NCCL communicator + staging (cpf_comm.cu), created by cpf_comm_init
    int since_sort = 0;
    int *d_sort_hist = nullptr;
    void *d_scratch = nullptr;
    size_t scratch_bytes = 0;

    unsigned long long *d_counters = nullptr;
    // asynchronous statistics (cpf_stats_request / cpf_stats_collect): ring of page-locked slots, one event each
    static constexpr int STAT_SLOTS = 4, STAT_WORDS = 16;
    unsigned long long *h_stat = nullptr;      // [STAT_SLOTS][STAT_WORDS] page-locked
    double *d_stat = nullptr;                  // [STAT_WORDS] device staging (fp64: one NCCL sum reduces it across ranks)
    cudaEvent_t evStat[STAT_SLOTS] = { nullptr, nullptr, nullptr, nullptr };
    bool statFull[STAT_SLOTS] = { false, false, false, false };
    int statHead = 0, statTail = 0;            // requests [statTail, statHead) are outstanding
    bool statBaseValid = false;                // a full scan has been collected since the particle set last changed
    bool statScanQueued = false;               // ... or at least requested
    long long statBaseActive = 0, statBaseNeg = 0, statBaseEsc = 0, statBaseFrz = 0;
    double statBaseKe = 0.0;
    int2 *d_queue[2] = { nullptr, nullptr }; // [n] ping-pong deferral queues of the filtered policy
    cpf::OutputState *output = nullptr;      // asynchronous VTU writer (cpf_output.cu), created on first use
    unsigned *d_queue_count = nullptr;       // [64]: queue lengths, one per queue of a launch sequence
};

namespace cpf {

int fail(cpf_context *ctx, int code, const char *fmt, ...);
#define CPF_CUDA(ctx, call)                                                                          \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess)                                                                      \
            return cpf::fail(ctx, CPF_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                             __FILE__, __LINE__);                                                    \
    } while (0)

MeshView mesh_view(const cpf_context *ctx);
ParticleView particle_view(const cpf_context *ctx);
int ensure_scratch(cpf_context *ctx, size_t bytes);

// cpf_mesh.cu
int build_device_mesh(cpf_context *ctx, long long nVerts, const double *pos, long long nTets, const int *tetVerts,
                      const int *tetCell, const int *tetPatch, long long nCells, int nPoints, bool cellFromVertex);
// cpf_locate.cu
int build_bvh(cpf_context *ctx);
void free_bvh(cpf_context *ctx);
int locate_particles(cpf_context *ctx, bool lostOnly = false);
// cpf_output.cu
int output_wait(cpf_context *ctx);      // drain the asynchronous writer; reports its first I/O error
void output_shutdown(cpf_context *ctx); // drain, join, free
// cpf_advect.cu
int launch_substeps(cpf_context *ctx, int nSub, double dt, bool writeVel);
int max_fused_substeps(const cpf_context *ctx);
int default_fused_substeps(const cpf_context *ctx);
int launch_initial_advect(cpf_context *ctx, double dt);
int launch_debug_normals(cpf_context *ctx, int k, double *d_xi);
int launch_init_rng(cpf_context *ctx);
int launch_point_interp(cpf_context *ctx);
// cpf_api.cu
int update_velocity_staged(cpf_context *ctx, const double *d_stage);
// cpf_comm.cu
int comm_reduce_stats(cpf_context *ctx, double *d_slot, int words);
void comm_release(cpf_context *ctx);
// cpf_sort.cu
int sort_particles_by_cell(cpf_context *ctx);
int gather_original_order(cpf_context *ctx, double4 *d_pos_out, double4 *d_vel_out, int *d_tet_out);
int stats_request(cpf_context *ctx, bool full);
int stats_collect(cpf_context *ctx, cpf_stats *out);
void stats_release(cpf_context *ctx);

} // namespace cpf
