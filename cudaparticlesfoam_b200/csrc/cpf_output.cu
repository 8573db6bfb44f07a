// cpf_output.cu -- particle output that does not stall the advection, and checkpoint / restart.
//
// Asynchronous binary VTU.  The reference's writeParticles2VTU (cuda/utils.cpp:144-283)
// copies every array to the host with blocking cudaMemcpy and prints it as ASCII between two Eulerian
// steps; with saveInterval 2 that is most of the run.  Here
//   * one kernel packs what the file holds -- every stride-th particle in ORIGINAL order: position
//     (3 x f64), type, id, tet id (i32), velocity (3 x f32), kinetic energy (f32) -- into a device
//     staging block on the compute stream (ordered after the sub-steps already enqueued);
//   * the copy stream moves the block into one of two page-locked host buffers;
//   * a writer thread waits for that copy and writes a VTK XML file with ONE raw appended-data
//     section (same array names as the reference, so existing ParaView states keep working).
// cpf_write_vtu_async returns once the kernel and the copy are enqueued; it only blocks when both
// host buffers are still being written.  cpf_output_wait / cpf_sync / cpf_destroy drain the writer.
//
// Checkpoint / restart.  Particle state in original order + the global sub-step index
// (the Philox counter) + cumulative counters (+ XORWOW states in that mode): a restarted run continues
// bit-identically, whatever the sort state was when the checkpoint was taken.
#include <cuda_runtime.h>
#include <curand_kernel.h>

#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <new>

#include "cpf_internal.h"

using namespace cpf;

namespace cpf {

// layout of one staging block for m output particles (all sections 8-byte aligned)
struct OutLayout {
    size_t pos, type, id, tet, vel, ke, total; // byte offsets; [0,8) holds the kinetic-energy sum (f64)
    explicit OutLayout(size_t m)
    {
        auto al = [](size_t x) { return (x + 7) & ~size_t(7); };
        pos = 8;
        type = pos + al(m * 24);
        id = type + al(m * 4);
        tet = id + al(m * 4);
        vel = tet + al(m * 4);
        ke = vel + al(m * 12);
        total = ke + al(m * 4);
    }
};

__global__ void k_pack_output(long long n, int stride, const int *__restrict__ pid, const double4 *__restrict__ pos,
                              const double4 *__restrict__ vel, const int *__restrict__ tet, unsigned char *__restrict__ out,
                              OutLayout L)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double ke = 0.0;
    if (i < n) {
        const long long o = pid[i];
        if (o % stride == 0) {
            const long long k = o / stride;
            const double4 p = pos[i];
            double4 v = vel[i];
            if (isnan(v.x)) v = make_double4(0.0, 0.0, 0.0, 0.0); // utils.cpp:236-239
            double *op = reinterpret_cast<double *>(out + L.pos) + 3 * k;
            op[0] = p.x; op[1] = p.y; op[2] = p.z;
            reinterpret_cast<int *>(out + L.type)[k] = (int)p.w;
            reinterpret_cast<int *>(out + L.id)[k] = (int)o;
            reinterpret_cast<int *>(out + L.tet)[k] = tet[i];
            float *ov = reinterpret_cast<float *>(out + L.vel) + 3 * k;
            ov[0] = (float)v.x; ov[1] = (float)v.y; ov[2] = (float)v.z;
            ke = 0.5 * (v.x * v.x + v.y * v.y + v.z * v.z); // utils.cpp:254-258
            reinterpret_cast<float *>(out + L.ke)[k] = (float)ke;
        }
    }
    // block sum of the kinetic energy -> one atomic per block
    __shared__ double s_ke[8];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) ke += __shfl_down_sync(0xffffffffu, ke, d);
    if ((threadIdx.x & 31) == 0) s_ke[threadIdx.x >> 5] = ke;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_ke[w];
        if (t != 0.0) atomicAdd(reinterpret_cast<double *>(out), t);
    }
}

__global__ void k_gather_rng(long long n, const int *__restrict__ pid, const curandState_t *__restrict__ in, curandState_t *__restrict__ out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[pid[i]] = in[i];
}

struct OutJob {
    int slot;
    long long m;
    std::string path;
};

struct OutputState {
    cpf_context *ctx = nullptr;
    unsigned char *d_stage[2] = { nullptr, nullptr };
    unsigned char *h_stage[2] = { nullptr, nullptr };
    size_t cap[2] = { 0, 0 };
    cudaEvent_t evPacked[2] = { nullptr, nullptr }, evCopied[2] = { nullptr, nullptr };
    bool busy[2] = { false, false };
    std::mutex mu;
    std::condition_variable cv;
    std::deque<OutJob> jobs;
    bool stop = false;
    std::thread writer;
    std::string error; // first I/O error of the writer
    std::vector<int> iota; // connectivity / offsets source, grown on demand
    std::vector<unsigned char> ones;

    void run();
    bool write_file(const OutJob &job);
};

static void put_block(FILE *fp, const void *data, unsigned long long bytes)
{
    fwrite(&bytes, 8, 1, fp);
    fwrite(data, 1, (size_t)bytes, fp);
}

bool OutputState::write_file(const OutJob &job)
{
    const long long m = job.m;
    const OutLayout L((size_t)m);
    const unsigned char *h = h_stage[job.slot];
    FILE *fp = fopen(job.path.c_str(), "wb");
    if (!fp) return false;
    static const size_t kBuf = 8u << 20;
    std::vector<char> iobuf(kBuf);
    setvbuf(fp, iobuf.data(), _IOFBF, kBuf);
    if ((long long)iota.size() < m + 1) {
        const size_t old = iota.size();
        iota.resize((size_t)m + 1);
        for (size_t k = old; k < iota.size(); ++k) iota[k] = (int)k;
        ones.assign((size_t)m, 1);
    }
    // sizes of the appended blocks, in file order
    const unsigned long long bPos = 24ull * m, bI = 4ull * m, bVel = 12ull * m, bKe = 4ull * m, bTypes = 1ull * m;
    unsigned long long off = 0;
    auto next = [&](unsigned long long bytes) { const unsigned long long o = off; off += 8 + bytes; return o; };
    fprintf(fp, "<VTKFile type='UnstructuredGrid' version='1.0' byte_order='LittleEndian' header_type='UInt64'>\n<UnstructuredGrid>\n");
    fprintf(fp, "<Piece NumberOfCells='%lld' NumberOfPoints='%lld'>\n<Points>\n", m, m);
    fprintf(fp, "<DataArray NumberOfComponents='3' type='Float64' Name='Position' format='appended' offset='%llu'/>\n", next(bPos));
    fprintf(fp, "</Points>\n<PointData>\n");
    fprintf(fp, "<DataArray NumberOfComponents='1' type='Int32' Name='ParticleType' format='appended' offset='%llu'/>\n", next(bI));
    fprintf(fp, "<DataArray NumberOfComponents='1' type='Int32' Name='ParticleID' format='appended' offset='%llu'/>\n", next(bI));
    fprintf(fp, "<DataArray NumberOfComponents='1' type='Int32' Name='ParticleTetID' format='appended' offset='%llu'/>\n", next(bI));
    fprintf(fp, "<DataArray NumberOfComponents='1' type='Int32' Name='ConvexTetID' format='appended' offset='%llu'/>\n", next(bI));
    fprintf(fp, "<DataArray NumberOfComponents='3' type='Float32' Name='vels' format='appended' offset='%llu'/>\n", next(bVel));
    fprintf(fp, "<DataArray NumberOfComponents='1' type='Float32' Name='KEs' format='appended' offset='%llu'/>\n", next(bKe));
    fprintf(fp, "</PointData>\n<Cells>\n");
    fprintf(fp, "<DataArray type='Int32' Name='connectivity' format='appended' offset='%llu'/>\n", next(bI));
    fprintf(fp, "<DataArray type='Int32' Name='offsets' format='appended' offset='%llu'/>\n", next(bI));
    fprintf(fp, "<DataArray type='UInt8' Name='types' format='appended' offset='%llu'/>\n", next(bTypes));
    fprintf(fp, "</Cells>\n</Piece>\n</UnstructuredGrid>\n<AppendedData encoding='raw'>\n_");
    put_block(fp, h + L.pos, bPos);
    put_block(fp, h + L.type, bI);
    put_block(fp, h + L.id, bI);
    put_block(fp, h + L.tet, bI);
    put_block(fp, h + L.tet, bI);
    put_block(fp, h + L.vel, bVel);
    put_block(fp, h + L.ke, bKe);
    put_block(fp, iota.data(), bI);     // connectivity 0..m-1
    put_block(fp, iota.data() + 1, bI); // offsets 1..m
    put_block(fp, ones.data(), bTypes); // VTK_VERTEX
    fprintf(fp, "\n</AppendedData>\n</VTKFile>\n");
    const bool ok = !ferror(fp);
    const bool closed = fclose(fp) == 0;
    double totalKE;
    memcpy(&totalKE, h, 8);
    printf("#adv: System Kinetic Energy=%lf\n", totalKE);
    return ok && closed;
}

void OutputState::run()
{
    cudaSetDevice(ctx->device);
    for (;;) {
        OutJob job;
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return stop || !jobs.empty(); });
            if (jobs.empty()) return; // stop requested and nothing left to write
            job = jobs.front();
            jobs.pop_front();
        }
        std::string err;
        const cudaError_t e = cudaEventSynchronize(evCopied[job.slot]);
        if (e != cudaSuccess) err = std::string("output copy failed: ") + cudaGetErrorString(e);
        else if (!write_file(job)) err = "cannot write " + job.path;
        {
            std::lock_guard<std::mutex> lk(mu);
            if (!err.empty() && error.empty()) error = err;
            busy[job.slot] = false;
        }
        cv.notify_all();
    }
}

static OutputState *output_state(cpf_context *ctx)
{
    if (!ctx->output) {
        OutputState *os = new OutputState;
        os->ctx = ctx;
        for (int s = 0; s < 2; ++s) {
            cudaEventCreateWithFlags(&os->evPacked[s], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&os->evCopied[s], cudaEventDisableTiming);
        }
        os->writer = std::thread([os] { os->run(); });
        ctx->output = os;
    }
    return ctx->output;
}

int output_wait(cpf_context *ctx)
{
    OutputState *os = ctx->output;
    if (!os) return CPF_OK;
    std::unique_lock<std::mutex> lk(os->mu);
    os->cv.wait(lk, [&] { return os->jobs.empty() && !os->busy[0] && !os->busy[1]; });
    if (!os->error.empty()) {
        const std::string e = os->error;
        os->error.clear();
        lk.unlock();
        return fail(ctx, CPF_ERR_INVALID, "%s", e.c_str());
    }
    return CPF_OK;
}

void output_shutdown(cpf_context *ctx)
{
    OutputState *os = ctx->output;
    if (!os) return;
    {
        std::lock_guard<std::mutex> lk(os->mu);
        os->stop = true;
    }
    os->cv.notify_all();
    os->writer.join(); // drains the queue first
    for (int s = 0; s < 2; ++s) {
        cudaFree(os->d_stage[s]);
        if (os->h_stage[s]) cudaFreeHost(os->h_stage[s]);
        cudaEventDestroy(os->evPacked[s]);
        cudaEventDestroy(os->evCopied[s]);
    }
    delete os;
    ctx->output = nullptr;
}

} // namespace cpf

extern "C" {

int cpf_write_vtu_async(cpf_context *ctx, const char *dir, unsigned step, int stride)
{
    if (!ctx) return CPF_ERR_INVALID;
    if (stride < 1) return fail(ctx, CPF_ERR_INVALID, "cpf_write_vtu_async: stride must be >= 1");
    cudaSetDevice(ctx->device);
    OutputState *os = output_state(ctx);
    const long long n = ctx->n;
    const long long m = (n + stride - 1) / stride;
    const OutLayout L((size_t)m);
    int slot;
    {
        std::unique_lock<std::mutex> lk(os->mu);
        os->cv.wait(lk, [&] { return !os->busy[0] || !os->busy[1]; });
        slot = os->busy[0] ? 1 : 0;
        if (!os->error.empty()) {
            const std::string e = os->error;
            os->error.clear();
            lk.unlock();
            return fail(ctx, CPF_ERR_INVALID, "%s", e.c_str());
        }
        os->busy[slot] = true;
    }
    auto release = [&] {
        std::lock_guard<std::mutex> lk(os->mu);
        os->busy[slot] = false;
    };
    if (os->cap[slot] < L.total) {
        cudaFree(os->d_stage[slot]);
        if (os->h_stage[slot]) cudaFreeHost(os->h_stage[slot]);
        os->d_stage[slot] = os->h_stage[slot] = nullptr;
        os->cap[slot] = 0;
        if (cudaMalloc(&os->d_stage[slot], L.total) != cudaSuccess || cudaMallocHost(&os->h_stage[slot], L.total) != cudaSuccess) {
            release();
            os->cv.notify_all();
            return fail(ctx, CPF_ERR_NOMEM, "cpf_write_vtu_async: cannot allocate %zu bytes of staging", L.total);
        }
        os->cap[slot] = L.total;
    }
    const int a = ctx->pcur;
    cudaError_t e = cudaMemsetAsync(os->d_stage[slot], 0, 8, ctx->stream);
    if (e == cudaSuccess && n > 0) {
        k_pack_output<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, stride, ctx->d_pid[a], ctx->d_pos[a], ctx->d_vel[a], ctx->d_tet[a],
                                                                            os->d_stage[slot], L);
        ctx->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaEventRecord(os->evPacked[slot], ctx->stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->copyStream, os->evPacked[slot], 0);
    if (e == cudaSuccess) e = cudaMemcpyAsync(os->h_stage[slot], os->d_stage[slot], L.total, cudaMemcpyDeviceToHost, ctx->copyStream);
    if (e == cudaSuccess) e = cudaEventRecord(os->evCopied[slot], ctx->copyStream);
    if (e != cudaSuccess) {
        release();
        os->cv.notify_all();
        return fail(ctx, CPF_ERR_CUDA, "cpf_write_vtu_async: %s", cudaGetErrorString(e));
    }
    char name[1200];
    snprintf(name, sizeof name, "%s%sparticle_%04u.vtu", dir ? dir : "", (dir && *dir) ? "/" : "", step);
    {
        std::lock_guard<std::mutex> lk(os->mu);
        os->jobs.push_back(OutJob{ slot, m, name });
    }
    os->cv.notify_all();
    return CPF_OK;
}

int cpf_output_wait(cpf_context *ctx)
{
    if (!ctx) return CPF_ERR_INVALID;
    return output_wait(ctx);
}

// ---------------------------------------------------------------------------------------------
// checkpoint / restart
// ---------------------------------------------------------------------------------------------
struct CkptHeader {
    char magic[8]; // "CPFCKPT1"
    long long n, nTets, nVerts;
    unsigned long long step_index;
    unsigned long long counters[8];
    int rng_mode, has_rng_state;
    unsigned long long seed;
    long long reserved[4];
};

int cpf_checkpoint_save(cpf_context *ctx, const char *path)
{
    if (!ctx || !path) return fail(ctx, CPF_ERR_INVALID, "cpf_checkpoint_save: bad arguments");
    if (!ctx->have_mesh) return fail(ctx, CPF_ERR_INVALID, "cpf_checkpoint_save: no mesh");
    cudaSetDevice(ctx->device);
    const size_t n = (size_t)ctx->n;
    std::vector<double> p(n * 4), v(n * 4);
    std::vector<int> tet(n);
    int rc = cpf_download(ctx, p.data(), v.data(), tet.data()); // original order, synchronises
    if (rc) return rc;
    CkptHeader h;
    memset(&h, 0, sizeof h);
    memcpy(h.magic, "CPFCKPT1", 8);
    h.n = ctx->n; h.nTets = ctx->nTets; h.nVerts = ctx->nVerts;
    h.step_index = ctx->step_index;
    h.rng_mode = ctx->cfg.rng;
    h.seed = ctx->cfg.seed;
    CPF_CUDA(ctx, cudaMemcpy(h.counters, ctx->d_counters, sizeof(unsigned long long) * CNT_COUNT, cudaMemcpyDeviceToHost));
    std::vector<curandState_t> states;
    if (ctx->cfg.rng == CPF_RNG_XORWOW && ctx->rng_ready && n) {
        h.has_rng_state = 1;
        rc = ensure_scratch(ctx, sizeof(curandState_t) * n);
        if (rc) return rc;
        const int a = ctx->pcur;
        k_gather_rng<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((long long)n, ctx->d_pid[a], ctx->d_rng[a], (curandState_t *)ctx->d_scratch);
        ctx->launches++;
        states.resize(n);
        CPF_CUDA(ctx, cudaMemcpyAsync(states.data(), ctx->d_scratch, sizeof(curandState_t) * n, cudaMemcpyDeviceToHost, ctx->stream));
        CPF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    FILE *fp = fopen(path, "wb");
    if (!fp) return fail(ctx, CPF_ERR_INVALID, "cpf_checkpoint_save: cannot open %s", path);
    bool ok = fwrite(&h, sizeof h, 1, fp) == 1;
    if (n) {
        ok = ok && fwrite(p.data(), sizeof(double) * 4, n, fp) == n;
        ok = ok && fwrite(v.data(), sizeof(double) * 4, n, fp) == n;
        ok = ok && fwrite(tet.data(), sizeof(int), n, fp) == n;
        if (h.has_rng_state) ok = ok && fwrite(states.data(), sizeof(curandState_t), n, fp) == n;
    }
    ok = (fclose(fp) == 0) && ok;
    if (!ok) return fail(ctx, CPF_ERR_INVALID, "cpf_checkpoint_save: short write to %s", path);
    return CPF_OK;
}

unsigned long long cpf_step_index(cpf_context *ctx) { return ctx ? ctx->step_index : 0ull; }

int cpf_checkpoint_load(cpf_context *ctx, const char *path)
{
    if (!ctx || !path) return fail(ctx, CPF_ERR_INVALID, "cpf_checkpoint_load: bad arguments");
    if (!ctx->have_mesh) return fail(ctx, CPF_ERR_INVALID, "cpf_checkpoint_load: upload the mesh first");
    cudaSetDevice(ctx->device);
    FILE *fp = fopen(path, "rb");
    if (!fp) return fail(ctx, CPF_ERR_INVALID, "cpf_checkpoint_load: cannot open %s", path);
    CkptHeader h;
    if (fread(&h, sizeof h, 1, fp) != 1 || memcmp(h.magic, "CPFCKPT1", 8) != 0) {
        fclose(fp);
        return fail(ctx, CPF_ERR_INVALID, "cpf_checkpoint_load: %s is not a checkpoint", path);
    }
    if (h.nTets != ctx->nTets || h.nVerts != ctx->nVerts || h.n < 0) {
        fclose(fp);
        return fail(ctx, CPF_ERR_INVALID, "cpf_checkpoint_load: checkpoint belongs to another mesh (%lld tets, %lld vertices)", h.nTets, h.nVerts);
    }
    if (h.rng_mode != ctx->cfg.rng || (h.rng_mode != CPF_RNG_NONE && h.seed != ctx->cfg.seed)) {
        fclose(fp);
        return fail(ctx, CPF_ERR_INVALID, "cpf_checkpoint_load: random-walk generator or seed differs from the checkpointed run");
    }
    // the particle count comes from the file: bound it by the file's own size before anything is allocated from it
    const long at = ftell(fp);
    fseek(fp, 0, SEEK_END);
    const long long fileLeft = (long long)ftell(fp) - at;
    fseek(fp, at, SEEK_SET);
    const long long perParticle = (long long)(sizeof(double) * 8 + sizeof(int) + (h.has_rng_state ? sizeof(curandState_t) : 0));
    if (h.n >= (1ll << 31) || h.n * perParticle > fileLeft) {
        fclose(fp);
        return fail(ctx, CPF_ERR_INVALID, "cpf_checkpoint_load: %s claims %lld particles but holds %lld bytes of state", path, h.n, fileLeft);
    }
    const size_t n = (size_t)h.n;
    std::vector<double> p, v;
    std::vector<int> tet;
    std::vector<curandState_t> states;
    try { // the ABI never throws
        p.resize(n * 4); v.resize(n * 4); tet.resize(n);
        if (h.has_rng_state) states.resize(n);
    } catch (const std::bad_alloc &) {
        fclose(fp);
        return fail(ctx, CPF_ERR_NOMEM, "cpf_checkpoint_load: out of host memory for %lld particles", h.n);
    }
    bool ok = true;
    if (n) {
        ok = fread(p.data(), sizeof(double) * 4, n, fp) == n && fread(v.data(), sizeof(double) * 4, n, fp) == n && fread(tet.data(), sizeof(int), n, fp) == n;
        if (ok && h.has_rng_state) ok = fread(states.data(), sizeof(curandState_t), n, fp) == n;
    }
    fclose(fp);
    if (!ok) return fail(ctx, CPF_ERR_INVALID, "cpf_checkpoint_load: %s is truncated", path);
    for (size_t i = 0; i < n; ++i)
        if (tet[i] >= ctx->nTets) return fail(ctx, CPF_ERR_INVALID, "cpf_checkpoint_load: tet id out of range");
    int rc = cpf_set_particles(ctx, h.n, p.data()); // allocates, identity order
    if (rc) return rc;
    if (n) {
        rc = cpf_set_tets(ctx, tet.data());
        if (rc) return rc;
        CPF_CUDA(ctx, cudaMemcpyAsync(ctx->d_vel[ctx->pcur], v.data(), sizeof(double4) * n, cudaMemcpyHostToDevice, ctx->stream));
        if (h.has_rng_state) {
            for (int b = 0; b < 2; ++b)
                if (!ctx->d_rng[b]) CPF_CUDA(ctx, cudaMalloc(&ctx->d_rng[b], sizeof(curandState_t) * n));
            CPF_CUDA(ctx, cudaMemcpyAsync(ctx->d_rng[ctx->pcur], states.data(), sizeof(curandState_t) * n, cudaMemcpyHostToDevice, ctx->stream));
            ctx->rng_ready = true;
        }
    }
    CPF_CUDA(ctx, cudaMemcpyAsync(ctx->d_counters, h.counters, sizeof(unsigned long long) * CNT_COUNT, cudaMemcpyHostToDevice, ctx->stream));
    CPF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->step_index = h.step_index;
    ctx->statBaseValid = ctx->statScanQueued = false; // counters and particle set replaced: the next statistics request scans
    ctx->since_sort = ctx->cfg.sort_interval > 0 ? ctx->cfg.sort_interval : 0; // re-sort by cell before the next sub-step
    return CPF_OK;
}

} // extern "C"
