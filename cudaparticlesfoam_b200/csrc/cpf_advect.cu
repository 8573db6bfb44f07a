// cpf_advect.cu -- the fused particle sub-step kernel (sm_100a).
//
// One launch = nSub iterations of the loop body of /root/reference/src/advect.H:96-161 for every
// particle, i.e. the reference's five kernels
//   particleAdvectKernelTetVel (cuda/particles.cu:316-373)      S1
//   particleBrownianMotion     (cuda/particles.cu:551-575)      S2
//   particleLocator            (query/ConvexQuery.cu:135-216)   S3   | baryQueryDisp (query/RTQuery.cu:221-248)
//   convexReflector            (query/ConvexQuery.cu:320-436)   S4   | RTreflection  (query/RTQuery.cu:109-186)
//   particleMoveKernel         (cuda/particles.cu:659-704)      S5
// fused, with the particle (position, flag, tet id) held in registers across the fused sub-steps:
// one 256-bit load + one 32-bit load per particle per launch, the same back.  disp never touches
// memory; vel is written only when the host can observe it (last sub-step of a call).
#include <algorithm>

#include "cpf_internal.h"

#ifndef CPF_MIN_BLOCKS
#define CPF_MIN_BLOCKS 4
#endif
#ifndef CPF_FAST_MIN_BLOCKS
#define CPF_FAST_MIN_BLOCKS 7
#endif
#ifndef CPF_RK_MIN_BLOCKS
#define CPF_RK_MIN_BLOCKS 6 /* all-particles pass with RK2 stage walks; RK4 (two more velocity accumulators): one less */
#endif
#ifndef CPF_WALL_MIN_BLOCKS
#define CPF_WALL_MIN_BLOCKS 4 /* k_fast with in-place wall reflection (queue passes) */
#endif
#ifndef CPF_FIN_MIN_BLOCKS
#define CPF_FIN_MIN_BLOCKS 3 /* the finishing pass (wall reflection + exact sub-steps in place) */
#endif
#ifndef CPF_WALL_BATCH
#define CPF_WALL_BATCH 32 /* lanes of a warp that must be waiting at a wall before they reflect together (32: all that are left; measured 4 < 8 < 16 < 32) */
#endif
#define CPF_TAIL __device__ __forceinline__

namespace cpf {

// ------------------------------------------------------------------------------------------------
// random walk deviates
// ------------------------------------------------------------------------------------------------
template <int RNG> struct Rng;

// Xi: the type the staged deviates of a chunk are kept in (shared memory); STATEFUL: the stream is a generator state in
// memory that advances by three normals per EXECUTED sub-step (skip(n): n sub-steps that another kernel has executed).
template <> struct Rng<CPF_RNG_NONE> {
    typedef float Xi;
    static constexpr bool STATEFUL = false;
    CPF_DEV void open(const ParticleView &, long long, const StepParams &) {}
    CPF_DEV bool draw(int, double &, double &, double &) { return false; }
    CPF_DEV void skip(int) {}
    CPF_DEV void close(const ParticleView &, long long) {}
    CPF_DEV void stage(Xi *, int, int, int, bool) {}
};

// The reference's stream: cuRAND XORWOW, curand_init(1591593751, particle, 0), three successive
// curand_normal_double per sub-step (cuda/particles.cu:537,565-567).  The state lives in registers
// for the whole launch: 48 B in + 48 B out per particle per launch instead of per sub-step.
template <> struct Rng<CPF_RNG_XORWOW> {
    typedef double Xi; // the reference adds curand_normal_double deviates: every bit counts
    static constexpr bool STATEFUL = true;
    curandState_t st;
    CPF_DEV void open(const ParticleView &pv, long long i, const StepParams &) { st = pv.rng[i]; }
    CPF_DEV void skip(int n)
    {
        for (int k = 0; k < 3 * n; ++k) (void)curand_normal_double(&st);
    }
    CPF_DEV bool draw(int, double &a, double &b, double &c)
    {
        a = curand_normal_double(&st);
        b = curand_normal_double(&st);
        c = curand_normal_double(&st);
        return true;
    }
    CPF_DEV void close(const ParticleView &pv, long long i) { pv.rng[i] = st; }
    // deviates of a whole chunk into xi[(3 q + c) * stride], drawn in sequence from the chunk-start state
    CPF_DEV void stage(Xi *xi, int stride, int nSub, int, bool live)
    {
        for (int q = 0; q < nSub; ++q) {
            double x0 = 0.0, x1 = 0.0, x2 = 0.0;
            if (live) draw(q, x0, x1, x2);
            xi[(q * 3 + 0) * stride] = x0;
            xi[(q * 3 + 1) * stride] = x1;
            xi[(q * 3 + 2) * stride] = x2;
        }
    }
};

// Stateless counter-based stream: Philox4x32-10 keyed by the seed, counter = (particle id,
// sub-step index); fp32 Box-Muller.  0 B of RNG traffic.  Statistical (not bit) parity with XORWOW.
CPF_DEV void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4])
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
// MUFU approximations (abs. error ~1e-6, ample for a random walk), spelled as the .ftz PTX forms so that no
// denormal/IEEE fix-up code is emitted: u1 >= 2^-25 and -2 ln u1 >= 2^-24 are normal numbers.  Every kernel
// draws through this one function, so the deviates of a (particle, sub-step) are the same bits everywhere.
CPF_DEV float mufu_lg2(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
CPF_DEV float mufu_rsqrt(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
CPF_DEV float mufu_sin(float x) { float r; asm("sin.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
CPF_DEV float mufu_cos(float x) { float r; asm("cos.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
CPF_DEV void box_muller_f32(uint32_t x, uint32_t y, double &n0, double &n1)
{
    const float u1 = __fmaf_rn((float)(x >> 8), 1.0f / 16777216.0f, 0.5f / 16777216.0f);
    const float u2 = __fmaf_rn((float)(y >> 8), 1.0f / 16777216.0f, 0.5f / 16777216.0f);
    const float q = __fmul_rn(mufu_lg2(u1), -1.3862943611198906f); // -2 ln u1
    const float r = __fmul_rn(q, mufu_rsqrt(q));
    const float ang = __fmul_rn(6.283185307179586f, u2);
    n0 = (double)__fmul_rn(r, mufu_sin(ang));
    n1 = (double)__fmul_rn(r, mufu_cos(ang));
}
// The stream of a particle is a sequence of N(0,1) deviates z_0, z_1, ...: block b = Philox4x32-10(counter = (global
// particle id, b), key = seed) yields z_{4b..4b+3} through two Box-Muller pairs; sub-step t (global index) consumes
// z_{3t}, z_{3t+1}, z_{3t+2}.  All four outputs of a block are used (3 blocks serve 4 sub-steps).
template <> struct Rng<CPF_RNG_PHILOX> {
    typedef float Xi;
    static constexpr bool STATEFUL = false;
    CPF_DEV void skip(int) {}
    uint32_t id, idhi, k0, k1;
    unsigned long long step0;
    CPF_DEV void open(const ParticleView &pv, long long i, const StepParams &sp)
    {
        // GLOBAL particle id: ranks that track index ranges of one cloud draw disjoint streams (cpf_set_particle_id_base)
        const unsigned long long gid = sp.idBase + (unsigned long long)pv.pid[i];
        id = (uint32_t)gid; idhi = (uint32_t)(gid >> 32);
        k0 = (uint32_t)sp.seed; k1 = (uint32_t)(sp.seed >> 32);
        step0 = sp.step0;
    }
    CPF_DEV void block(unsigned long long b, double z[4])
    {
        uint32_t o[4];
        philox4x32_10(id, idhi, (uint32_t)b, (uint32_t)(b >> 32), k0, k1, o);
        box_muller_f32(o[0], o[1], z[0], z[1]);
        box_muller_f32(o[2], o[3], z[2], z[3]);
    }
    CPF_DEV bool draw(int s, double &a, double &b, double &c)
    {
        const unsigned long long m = 3ull * (step0 + (unsigned long long)s);
        const int r = (int)(m & 3ull);
        double z[8];
        block(m >> 2, z);
        if (r > 1) block((m >> 2) + 1ull, z + 4);
        a = r == 0 ? z[0] : (r == 1 ? z[1] : (r == 2 ? z[2] : z[3]));
        b = r == 0 ? z[1] : (r == 1 ? z[2] : (r == 2 ? z[3] : z[4]));
        c = r == 0 ? z[2] : (r == 1 ? z[3] : (r == 2 ? z[4] : z[5]));
        return true;
    }
    CPF_DEV void close(const ParticleView &, long long) {}
    // whole blocks b0, b0+1, ... (b0 = the block holding the chunk's first deviate) into rows 4k..4k+3: no per-slot bounds
    // (k_lean; the reader starts (3 step0) & 3 rows in)
    CPF_DEV void stage_blocks(Xi *xi, int stride, unsigned nBlocks, bool live)
    {
        if (!live) return;
        const unsigned long long b0 = (3ull * step0) >> 2;
        for (unsigned k = 0; k < nBlocks; ++k) {
            double z[4];
            block(b0 + k, z);
#pragma unroll
            for (int j = 0; j < 4; ++j) xi[(4 * k + j) * stride] = (Xi)z[j];
        }
    }
    // deviates of the sub-steps [sFrom, nSub) of a chunk into xi[(3 q + c) * stride]: block by block, every output used
    CPF_DEV void stage(Xi *xi, int stride, int nSub, int sFrom, bool live)
    {
        const unsigned long long m0 = 3ull * step0;
        const long long lo = 3ll * sFrom, hi = 3ll * nSub;              // wanted slots, relative to m0
        const unsigned long long b0 = (m0 + (unsigned long long)lo) >> 2, b1 = (m0 + (unsigned long long)hi - 1ull) >> 2;
        if (!live || sFrom >= nSub) return;
        for (unsigned long long b = b0; b <= b1; ++b) {
            double z[4];
            block(b, z);
            const long long base = (long long)(4ull * b - m0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const long long slot = base + j;
                if (slot >= lo && slot < hi) xi[slot * stride] = (Xi)z[j];
            }
        }
    }
};

// ------------------------------------------------------------------------------------------------
// exact sub-step tails (S3+S4+S5)
// ------------------------------------------------------------------------------------------------
struct Tally { unsigned hops, exact, refl, esc, frz; };
// deferral queue entry {particle, sub-step | flags}: CPF_Q_WALL = the all-particles pass stopped at a CERTIFIED wall contact
// (the wall-capable pass can reflect it in place); without it the fp32 filter itself refused, and will again
enum { CPF_Q_SUBSTEP = 0xffff, CPF_Q_WALL = 0x10000 };
CPF_DEV double4 vel4(D3 u) { return make_double4(u.x, u.y, u.z, -1.0); }

// Default build: convex line walk + reflector.  The reference's reflector re-walks the segment
// from the start tet with bit-identical arithmetic (ConvexQuery.cu:343-397 vs :165-200), so the
// locator's wall state IS the reflector's state after its first inner loop; we continue from it.
CPF_TAIL void tail_convex_exact(const MeshView &m, D3 &P, D3 disp, D3 &vel, int &tet, double &w, int reflect, Tally &ty)
{
    const D3 E0 = xadd(P, disp);
    D3 E = E0, S = P;
    int cur = tet, in_j = -1, wallLink = -1;
    bool wall = false;
    Tet T;
    int4 v;
    for (int i = 0; i < 50; ++i) {
        T = load_tet(m, cur, v);
        ty.hops++;
        const int out = trace_exact(T, S, E, in_j);
        if (out < 0) break;
        const int link = link_at(T.link, out);
        in_j = out;
        if (link < 0) { wall = true; wallLink = link; break; }
        cur = link >> 2;
        in_j = link & 3;
    }
    if (!wall) {
        tet = cur;
        P = xadd(P, disp); // S5
        return;
    }
    if (!reflect) { // reflectWall == false: id stays -(tet+1), particle is moved, frozen next step
        tet = -(tet + 1);
        P = xadd(P, disp);
        return;
    }
    // S4: convexReflector, up to 5 wall hits
    D3 Phit = S;
    int next = -1;
    for (int j = 0; j < 5; ++j) {
        if (j > 0) {
            next = -1;
            bool hitwall = false;
            int i = 0;
            for (; i < 50; ++i) {
                T = load_tet(m, cur, v);
                ty.hops++;
                const int out = trace_exact(T, S, E, in_j);
                if (out < 0) break;
                const int link = link_at(T.link, out);
                in_j = out;
                if (link < 0) { hitwall = true; wallLink = link; break; }
                cur = link >> 2;
                in_j = link & 3;
            }
            if (!hitwall) { next = cur; break; } // end point found (or 50-tet cap: next == cur)
        }
        Phit = S;
        // per-patch boundary action (extension; the reference reflects everywhere, RTQuery.cu:165-166):
        // an ESCAPE patch parks the particle at the exit point and deactivates it
        if (m.patch_kind[-wallLink - 1] == CPF_PATCH_ESCAPE) {
            P = Phit;
            w = 0.0;
            tet = -(cur + 1);
            ty.esc++;
            return;
        }
        ty.refl++;
        reflect_exact(m, T, Phit, E, vel);
    }
    const D3 nd = xsub(E, Phit);
    tet = next;
    P = xadd(Phit, nd); // p = P_hit (S4) then p += disp (S5)
}

// the same as an out-of-line call: the finishing pass (k_fast<..., FIN = 1>) runs a refused sub-step through it without
// carrying the reflector's registers through its fp32 visit loop
__device__ __noinline__ void exact_substep_convex(const MeshView &m, D3 &P, D3 disp, D3 &vel, int &tet, double &w, int reflect, Tally &ty)
{
    tail_convex_exact(m, P, disp, vel, tet, w, reflect, ty);
}

// RTX=true build: barycentric point walk + RTreflection
CPF_DEV int bary_search(const MeshView &m, D3 Q, int start, Tet &T, int &face_j, Tally &ty)
{
    int s = start;
    face_j = -1;
    int4 v;
    for (int i = 0; i < 50; ++i) {
        T = load_tet(m, s, v);
        ty.hops++;
        double w[4];
        bary_exact(T, Q, w);
        const double wmin = fmin(fmin(w[0], w[1]), fmin(w[2], w[3]));
        if (wmin >= 0.0) break;
        int k = 0;
        if (w[1] < w[0]) k = 1;
        if (w[2] < (k == 0 ? w[0] : w[1])) k = 2;
        if (w[3] < (k == 0 ? w[0] : (k == 1 ? w[1] : w[2]))) k = 3;
        face_j = (T.code >> (2 * k)) & 3u;
        const int link = link_at(T.link, face_j);
        if (link < 0) return -(s + 1);
        s = link >> 2;
    }
    return s;
}

CPF_TAIL void tail_bary_exact(const MeshView &m, D3 &P, D3 disp, D3 &vel, int &tet, int reflect, Tally &ty)
{
    Tet T;
    int fj;
    D3 R = xadd(P, disp);
    int s = bary_search(m, R, tet, T, fj, ty);
    if (s < 0 && reflect) {
        int bd = -(s + 1);
        for (int i = 0; i < 10; ++i) {
            if (i > 0) {
                s = bary_search(m, R, bd, T, fj, ty);
                if (s >= 0) { bd = s; break; }
                bd = -(s + 1);
            }
            // specularReflect (query/RTQuery.cu:92-107) on face fj of tet bd (= T)
            D3 A;
            const D3 n = face_normal_exact(T, fj, A);
            const double gain = face_gain(m, T, fj);
            double sp = xdot(xsub(R, A), n);
            sp = __dmul_rn(gain, sp);
            double sv = xdot(vel, n);
            sv = __dmul_rn(gain, sv);
            R = D3{ __fma_rn(-sp, n.x, R.x), __fma_rn(-sp, n.y, R.y), __fma_rn(-sp, n.z, R.z) };
            vel = D3{ __fma_rn(-sv, n.x, vel.x), __fma_rn(-sv, n.y, vel.y), __fma_rn(-sv, n.z, vel.z) };
            ty.refl++;
        }
        s = bd;
        disp = xsub(R, P);
    }
    tet = s;
    P = xadd(P, disp);
}

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
// S1 + S2 of one sub-step: vel = U[cell]; disp = (P + dt*vel) - P (+ random walk)
template <int RNG>
CPF_DEV D3 displacement(const MeshView &m, const StepParams &sp, Rng<RNG> &rng, int s, int cell, const D3 &P, D3 &vel)
{
    vel = ld_ucell(m, cell);
    D3 disp{ __dsub_rn(__fma_rn(sp.dt, vel.x, P.x), P.x), __dsub_rn(__fma_rn(sp.dt, vel.y, P.y), P.y),
             __dsub_rn(__fma_rn(sp.dt, vel.z, P.z), P.z) };
    double x0, x1, x2;
    if (rng.draw(s, x0, x1, x2)) {
        disp.x = __fma_rn(x0, sp.randDisp, disp.x);
        disp.y = __fma_rn(x1, sp.randDisp, disp.y);
        disp.z = __fma_rn(x2, sp.randDisp, disp.z);
    }
    return disp;
}

// First slot of a thread in a grid-stride loop (stride = gridDim.x * blockDim.x).  Queue passes spread the entries over
// ALL warps of the grid -- lane k of warp g takes entry g + k * (number of warps) -- instead of packing 32 neighbours
// into one warp: the passes are short, divergent and latency-bound, a queue of a few 1e4 entries fills a fraction of the
// resident warps when packed, and every lane less in a warp is one serialised rare-event section less.
CPF_DEV long long first_slot(bool queue, long long total)
{
    // a queue longer than the grid (most particles deferred: a cloud held against a wall) is better off packed (coalesced)
    if (!queue || total >= (long long)gridDim.x * blockDim.x) return (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned wpb = blockDim.x >> 5;
    return (long long)(blockIdx.x * wpb + (threadIdx.x >> 5)) + (long long)(threadIdx.x & 31u) * ((long long)gridDim.x * wpb);
}

// frz: particles this kernel froze (S1: negative tet id -> w := 0); with the escapes it lets the host derive the number of
// active particles from the counters alone (cpf_stats_request light mode)
CPF_DEV void flush_counters(const StepParams &sp, unsigned refl, unsigned exact, unsigned hops, unsigned nsteps, unsigned esc = 0u, unsigned frz = 0u)
{
    const unsigned vals[6] = { esc, refl, exact, hops, nsteps, frz };
    constexpr int slot[6] = { CNT_ESCAPED, CNT_REFLECT, CNT_EXACT, CNT_HOPS, CNT_SUBSTEPS, CNT_FROZEN };
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        unsigned x = __reduce_add_sync(0xffffffffu, vals[c]);
        if ((threadIdx.x & 31) == 0 && x) atomicAdd(sp.counters + slot[c], (unsigned long long)x);
    }
}

// k_exact<LOC,RNG,QMODE>: sub-steps entirely in the reference's arithmetic.
//   QMODE 0: thread i = particle i, all nSub sub-steps (CPF_PATH_EXACT cross-check mode, RTX build)
//   QMODE 1: one sub-step per entry of the deferral queue; the entry is advanced to s+1 in place
//   QMODE 2: entries of the deferral queue, all their remaining sub-steps (last round)
template <int LOC, int RNG, int QMODE>
__global__ void __launch_bounds__(128, CPF_MIN_BLOCKS) k_exact(const MeshView m, const ParticleView pv, const StepParams sp)
{
    Tally ty{ 0u, 0u, 0u, 0u, 0u };
    unsigned nsteps = 0;
    const long long total = QMODE ? (long long)*sp.countIn : pv.n;
    for (long long slot = first_slot(QMODE != 0, total); slot < total; slot += (long long)gridDim.x * blockDim.x) {
        long long i = slot;
        int s0 = 0;
        if (QMODE) { const int2 q = sp.queueIn[slot]; i = q.x; s0 = q.y & CPF_Q_SUBSTEP; }
        const int s1 = (QMODE == 1) ? min(s0 + 1, sp.nSub) : sp.nSub;
        if (s0 >= s1) continue;
        double4 p4 = ld_stream4(pv.pos + i);
        int tet = ld_stream_i(pv.tet + i);
        D3 P{ p4.x, p4.y, p4.z };
        double w = p4.w;
        if (w == 0.0) continue;
        D3 vel{ 0.0, 0.0, 0.0 };
        bool velValid = false;
        Rng<RNG> rng;
        rng.open(pv, i, sp);
        if (QMODE) rng.skip(s0);
        for (int s = s0; s < s1; ++s) {
            if (w == 0.0) break;
            if (tet < 0) { w = 0.0; ty.frz++; break; } // S1: left the domain -> frozen (particles.cu:334-338)
            const int cell = tet_cell(m, tet, ld_int4(m.tetv, tet));
            const D3 disp = displacement<RNG>(m, sp, rng, s, cell, P, vel);
            velValid = true;
            nsteps++;
            ty.exact++;
            if (LOC == CPF_LOCATOR_CONVEX) tail_convex_exact(m, P, disp, vel, tet, w, sp.reflect, ty);
            else tail_bary_exact(m, P, disp, vel, tet, sp.reflect, ty);
        }
        rng.close(pv, i);
        st_stream4(pv.pos + i, make_double4(P.x, P.y, P.z, w));
        st_stream_i(pv.tet + i, tet);
        if (sp.writeVel && velValid && s1 == sp.nSub) st_stream4(pv.vel + i, make_double4(vel.x, vel.y, vel.z, -1.0));
        if (QMODE == 1) sp.queueIn[slot].y = s1;
    }
    flush_counters(sp, ty.refl, ty.exact, ty.hops, nsteps, ty.esc, ty.frz);
}

// k_exact_convex<RNG,QMODE>: k_exact for the default ConvexPoly build as ONE merged loop over tet visits.
// A sub-step in the reference's arithmetic is a chain of visits (load the tet, traceIntet) interleaved with
// rare events (wall reflection, end of the sub-step); lanes need very different numbers of visits, so the
// per-lane loop nest of tail_convex_exact leaves most lanes of a warp idle.  Here every iteration performs
// one visit for every lane that still has work, and the events move a small per-lane state machine:
//   leg      reflections so far in this sub-step (the reflector's j, ConvexQuery.cu:343-397)
//   legHops  visits in the current leg (the 50-tet cap of each walk loop)
// The arithmetic and its order are those of tail_convex_exact; QMODE as for k_exact.
template <int RNG, int QMODE>
__global__ void __launch_bounds__(128, CPF_MIN_BLOCKS) k_exact_convex(const MeshView m, const ParticleView pv, const StepParams sp)
{
    Tally ty{ 0u, 0u, 0u, 0u, 0u };
    unsigned nsteps = 0;
    const long long total = QMODE ? (long long)*sp.countIn : pv.n;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long base = 0; base < total; base += stride) { // grid-uniform trip count
        const long long slot = base + first_slot(QMODE != 0, total);
        long long i = slot;
        int s = 0;
        bool have = slot < total;
        if (QMODE && have) { const int2 q = sp.queueIn[slot]; i = q.x; s = q.y & CPF_Q_SUBSTEP; have = s < sp.nSub; }
        const int sStop = (QMODE == 1) ? min(s + 1, sp.nSub) : sp.nSub;
        double4 p4 = make_double4(0.0, 0.0, 0.0, 0.0);
        int tet = -1;
        if (have) { p4 = ld_stream4(pv.pos + i); tet = ld_stream_i(pv.tet + i); }
        D3 P{ p4.x, p4.y, p4.z };
        double w = p4.w;
        const bool live = have && (w != 0.0);
        bool active = live;
        D3 vel{ 0.0, 0.0, 0.0 };
        bool velValid = false;
        Rng<RNG> rng;
        if (live) { rng.open(pv, i, sp); if (QMODE) rng.skip(s); }
        D3 S{ 0.0, 0.0, 0.0 }, E = S, Phit = S;
        int cur = -1, in_j = -1, leg = 0, legHops = 0;
        bool needPro = true;
        while (__any_sync(0xffffffffu, active)) {
            if (active && needPro) {
                if (tet < 0) { w = 0.0; ty.frz++; active = false; } // S1: left the domain -> frozen (particles.cu:334-338)
                else {
                    const int cell = tet_cell(m, tet, ld_int4(m.tetv, tet));
                    const D3 disp = displacement<RNG>(m, sp, rng, s, cell, P, vel);
                    velValid = true;
                    nsteps++;
                    ty.exact++;
                    E = xadd(P, disp); // also the S5 result of a sub-step without wall contact
                    S = P;
                    cur = tet; in_j = -1; leg = 0; legHops = 0;
                    needPro = false;
                }
            }
            if (active) {
                int4 v;
                const Tet T = load_tet(m, cur, v);
                ty.hops++;
                legHops++;
                const int out = trace_exact(T, S, E, in_j);
                bool subDone = false;
                bool legEnd = out < 0; // the end point lies in this tet
                if (!legEnd) {
                    const int link = link_at(T.link, out);
                    in_j = out;
                    if (link >= 0) {
                        cur = link >> 2;
                        in_j = link & 3;
                        legEnd = legHops >= 50; // the walk loop's cap: it ends on the tet just entered
                    } else if (!sp.reflect) { // reflectWall == false: id -(tet+1), particle is moved, frozen next step
                        tet = -(tet + 1);
                        P = E;
                        subDone = true;
                    } else if (m.patch_kind[-link - 1] == CPF_PATCH_ESCAPE) { // extension, see tail_convex_exact
                        P = S;
                        w = 0.0;
                        tet = -(cur + 1);
                        ty.esc++;
                        subDone = true;
                    } else { // S4: convexReflector, up to 5 wall hits
                        Phit = S;
                        ty.refl++;
                        reflect_exact(m, T, Phit, E, vel);
                        legHops = 0;
                        if (++leg >= 5) { // fifth hit: no further walk, the particle is lost (next == -1)
                            tet = -1;
                            P = xadd(Phit, xsub(E, Phit));
                            subDone = true;
                        }
                    }
                }
                if (legEnd) {
                    tet = cur;
                    P = leg == 0 ? E : xadd(Phit, xsub(E, Phit)); // p = P_hit (S4) then p += disp (S5)
                    subDone = true;
                }
                if (subDone) {
                    needPro = true;
                    if (++s >= sStop || w == 0.0) active = false;
                }
            }
        }
        if (live) {
            rng.close(pv, i);
            st_stream4(pv.pos + i, make_double4(P.x, P.y, P.z, w));
            st_stream_i(pv.tet + i, tet);
            if (sp.writeVel && velValid && (s >= sp.nSub || w == 0.0))
                st_stream4(pv.vel + i, make_double4(vel.x, vel.y, vel.z, -1.0));
            if (QMODE == 1) sp.queueIn[slot].y = (w == 0.0) ? sp.nSub : s;
        }
        if (!QMODE) break;
    }
    flush_counters(sp, ty.refl, ty.exact, ty.hops, nsteps, ty.esc, ty.frz);
}

// ------------------------------------------------------------------------------------------------
// k_general<RNG>: integrators and interpolation the reference does not wire up (RK2 midpoint, RK4,
// vertex / cellPoint-style interpolation), in the reference's arithmetic for every building block;
// semantics as written down in DESIGN.md section 7 (the test oracle restates them).  Default ConvexPoly locator.
// ------------------------------------------------------------------------------------------------
CPF_DEV D3 vertex_velocity_exact(const MeshView &m, int tet, D3 P)
{
    int4 v;
    const Tet T = load_tet(m, tet, v);
    const int sid[4] = { v.x, v.y, v.z, v.w };
    int k[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) k[q] = (T.code >> (2 * q)) & 3u;
    const D3 A = T.P[k[0]], B = T.P[k[1]], C = T.P[k[2]], D = T.P[k[3]];
    const double rden = __drcp_rn(xdet(A, B, C, D));
    const double wA = __dmul_rn(xdet(P, B, C, D), rden);
    const double wB = __dmul_rn(xdet(A, P, C, D), rden);
    const double wC = __dmul_rn(xdet(A, B, P, D), rden);
    const D3 cr = xcross(xsub(B, A), xsub(C, A)), rr = xsub(P, A);
    const double wD = __dmul_rn(__fma_rn(cr.z, rr.z, __fma_rn(cr.y, rr.y, __dmul_rn(cr.x, rr.x))), rden);
    const double *uA = m.uvert + 3ll * sid[k[0]], *uB = m.uvert + 3ll * sid[k[1]], *uC = m.uvert + 3ll * sid[k[2]],
                 *uD = m.uvert + 3ll * sid[k[3]];
    D3 r;
    r.x = __fma_rn(wD, uD[0], __fma_rn(wC, uC[0], __fma_rn(wA, uA[0], __dmul_rn(wB, uB[0]))));
    r.y = __fma_rn(wD, uD[1], __fma_rn(wC, uC[1], __fma_rn(wA, uA[1], __dmul_rn(wB, uB[1]))));
    r.z = __fma_rn(wD, uD[2], __fma_rn(wC, uC[2], __fma_rn(wA, uA[2], __dmul_rn(wB, uB[2]))));
    return r;
}

CPF_DEV D3 velocity_at(const MeshView &m, int interp, int tet, D3 P)
{
    if (interp == CPF_INTERP_VERTEX) return vertex_velocity_exact(m, tet, P);
    const int cell = tet_cell(m, tet, ld_int4(m.tetv, tet));
    return ld_ucell(m, cell);
}

// tet of a stage point: the reference's segment walk from (from, tet); beyond a wall -> last tet
CPF_DEV int stage_tet(const MeshView &m, D3 from, D3 to, int tet, unsigned &hops)
{
    D3 S = from;
    int cur = tet, in_j = -1;
    int4 v;
    for (int i = 0; i < 50; ++i) {
        const Tet T = load_tet(m, cur, v);
        hops++;
        const int out = trace_exact(T, S, to, in_j);
        if (out < 0) break;
        const int link = link_at(T.link, out);
        if (link < 0) break;
        cur = link >> 2;
        in_j = link & 3;
    }
    return cur;
}

CPF_DEV D3 axpy3(double h, D3 k, D3 P) { return D3{ __fma_rn(h, k.x, P.x), __fma_rn(h, k.y, P.y), __fma_rn(h, k.z, P.z) }; }

//   QMODE 0: thread i = particle i, all nSub sub-steps      QMODE 2: entries of the deferral queue, their remaining sub-steps
template <int RNG, int QMODE>
__global__ void __launch_bounds__(128, 3) k_general(const MeshView m, const ParticleView pv, const StepParams sp)
{
    Tally ty{ 0u, 0u, 0u, 0u, 0u };
    unsigned nsteps = 0;
    const long long total = QMODE ? (long long)*sp.countIn : pv.n;
    for (long long slot = first_slot(QMODE != 0, total); slot < total; slot += (long long)gridDim.x * blockDim.x) {
        long long i = slot;
        int s0 = 0;
        if (QMODE) { const int2 q = sp.queueIn[slot]; i = q.x; s0 = q.y & CPF_Q_SUBSTEP; }
        if (s0 >= sp.nSub) continue;
        double4 p4 = ld_stream4(pv.pos + i);
        int tet = ld_stream_i(pv.tet + i);
        D3 P{ p4.x, p4.y, p4.z };
        double w = p4.w;
        D3 vel{ 0.0, 0.0, 0.0 };
        bool velValid = false;
        if (w != 0.0) {
            Rng<RNG> rng;
            rng.open(pv, i, sp);
            if (QMODE) rng.skip(s0);
            for (int s = s0; s < sp.nSub; ++s) {
                if (w == 0.0) break;
                if (tet < 0) { w = 0.0; ty.frz++; break; }
                const D3 k1 = velocity_at(m, sp.interp, tet, P);
                vel = k1;
                if (sp.integrator == CPF_RK2) {
                    const D3 Pm = axpy3(__dmul_rn(0.5, sp.dt), k1, P);
                    vel = velocity_at(m, sp.interp, stage_tet(m, P, Pm, tet, ty.hops), Pm);
                } else if (sp.integrator == CPF_RK4) {
                    const double h = __dmul_rn(0.5, sp.dt);
                    const D3 P2 = axpy3(h, k1, P);
                    const D3 k2 = velocity_at(m, sp.interp, stage_tet(m, P, P2, tet, ty.hops), P2);
                    const D3 P3 = axpy3(h, k2, P);
                    const D3 k3 = velocity_at(m, sp.interp, stage_tet(m, P, P3, tet, ty.hops), P3);
                    const D3 P4 = axpy3(sp.dt, k3, P);
                    const D3 k4 = velocity_at(m, sp.interp, stage_tet(m, P, P4, tet, ty.hops), P4);
                    vel.x = __ddiv_rn(__fma_rn(2.0, __dadd_rn(k2.x, k3.x), __dadd_rn(k1.x, k4.x)), 6.0);
                    vel.y = __ddiv_rn(__fma_rn(2.0, __dadd_rn(k2.y, k3.y), __dadd_rn(k1.y, k4.y)), 6.0);
                    vel.z = __ddiv_rn(__fma_rn(2.0, __dadd_rn(k2.z, k3.z), __dadd_rn(k1.z, k4.z)), 6.0);
                }
                velValid = true;
                D3 disp{ __dsub_rn(__fma_rn(sp.dt, vel.x, P.x), P.x), __dsub_rn(__fma_rn(sp.dt, vel.y, P.y), P.y),
                         __dsub_rn(__fma_rn(sp.dt, vel.z, P.z), P.z) };
                double x0, x1, x2;
                if (rng.draw(s, x0, x1, x2)) {
                    disp.x = __fma_rn(x0, sp.randDisp, disp.x);
                    disp.y = __fma_rn(x1, sp.randDisp, disp.y);
                    disp.z = __fma_rn(x2, sp.randDisp, disp.z);
                }
                nsteps++;
                ty.exact++;
                tail_convex_exact(m, P, disp, vel, tet, w, sp.reflect, ty);
            }
            rng.close(pv, i);
            st_stream4(pv.pos + i, make_double4(P.x, P.y, P.z, w));
            st_stream_i(pv.tet + i, tet);
            if (sp.writeVel && velValid) st_stream4(pv.vel + i, make_double4(vel.x, vel.y, vel.z, -1.0));
        }
    }
    flush_counters(sp, ty.refl, ty.exact, ty.hops, nsteps, ty.esc, ty.frz);
}

// k_fast<RNG,QMODE,WALL,INTEG>: the kernels of the filtered policy (default ConvexPoly build).
// fp32 guarded walk -- in the all-particles pass (WALL = 0) no exact-arithmetic code at all, hence 72
// registers and 7 CTAs per SM.  ONE merged loop runs the tet visits of all fused sub-steps of a lane
// (visit_fast32): lanes need different numbers of visits per sub-step, and a merged loop keeps them busy
// until their whole chunk is done instead of idling at every sub-step boundary.  The random-walk deviates
// of the chunk are drawn up front, with all lanes converged, into shared memory ([sub-step][component]
// [thread], conflict-free), so the per-sub-step prologue inside the divergent loop is only the velocity
// fetch and three fp64 FMAs.  The first sub-step whose walk is refused is NOT executed: the particle is
// written back as it was at the start of that sub-step and (particle, sub-step) is appended to the output
// queue with one warp-aggregated atomic (ballot + popc); the next kernel of the launch sequence
// (launch_filtered) takes it from there.
//   QMODE 0: thread i = particle i from sub-step 0      QMODE 2: entries of the input queue
// WALL = 1 (queue passes): a wall contact on the first leg of a sub-step -- the case of particles that live
// next to a wall: diffusion at a wall, through-flow held against the reflecting outlet -- is reflected in
// place with wall_reflect_on_path (fp64 replay of the certified crossings only) and the walk goes on from the
// hit point; one such contact per sub-step, anything else about a wall is still deferred.  Costs registers
// (4 CTAs/SM), hence not in the all-particles pass, whose refusals land in the first queue pass anyway.
// INTEG = CPF_RK2 / CPF_RK4 (extensions): the stage points P + h*k are located with the same
// guarded walk from (P, tet) -- a stage walk that ends at a certified wall face stays in that tet, like the exact
// stage_tet -- and feed the cell velocities of the stages into v_eff; then the move walk runs as for Euler.  A
// refused stage walk defers the whole sub-step to k_general.
// VERT = 1 (extension): velocities from the vertex (cellPoint-style) interpolation, evaluated in the reference's
// arithmetic (vertex_velocity_exact) at the particle and at the stage points; only the walks are filtered.
// FIN = 1 (Euler, cell value; the LAST kernel of the launch sequence): nothing is deferred.  A sub-step the filter
// refuses is run right here, from its start (P, tet -- both untouched while a walk is under way), in the reference's
// arithmetic (exact_substep_convex), and the lane goes back to the fp32 walk for its next sub-step after a C1 check of
// the new start point (which no C2 has certified).  One kernel, and one exact sub-step per refusal, instead of a
// wall-capable pass followed by a finisher that ran ALL remaining sub-steps of its particles in fp64.
template <int RNG, int QMODE, int WALL, int INTEG, int VERT, int FIN = 0>
__global__ void __launch_bounds__(128, FIN ? CPF_FIN_MIN_BLOCKS : (WALL || VERT) ? CPF_WALL_MIN_BLOCKS : (INTEG == CPF_RK4 ? CPF_RK_MIN_BLOCKS - 1 : INTEG ? CPF_RK_MIN_BLOCKS : CPF_FAST_MIN_BLOCKS))
k_fast(const MeshView m, const ParticleView pv, const StepParams sp)
{
    static_assert(!FIN || (!INTEG && !VERT), "the in-place exact sub-step is the Euler / cell-value one");
    constexpr bool KEEPV = INTEG || VERT; // the reported velocity is not simply U[cell]: carry it
    typedef typename Rng<RNG>::Xi Xi;
    constexpr bool STATEFUL = Rng<RNG>::STATEFUL;
    extern __shared__ double s_dyn[];
    Xi *s_xi = reinterpret_cast<Xi *>(s_dyn);
    // stateful generator: the state after the whole chunk waits here and is committed only if the particle finishes its
    // chunk in this kernel -- a deferred particle keeps its chunk-start state in memory, the next kernel re-draws from it
    curandState_t *stash = reinterpret_cast<curandState_t *>(s_xi + 3 * 128 * (STATEFUL ? sp.nSub : 0)) + threadIdx.x;
    unsigned hops = 0, nsteps = 0, refl = 0, frz = 0;
    Tally ty{ 0u, 0u, 0u, 0u, 0u }; // FIN: what the exact sub-steps count
    const long long total = QMODE ? (long long)*sp.countIn : pv.n;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long base = 0; base < total; base += stride) { // grid-uniform trip count
        const long long slot = base + first_slot(QMODE != 0, total);
        int deferAt = -1;
        bool needExact = false, exactVel = false, frzHit = false; // FIN: sub-step s is to be redone exactly / the velocity to report is velX
        D3 velX{ 0.0, 0.0, 0.0 };
        long long i = slot;
        int s = 0, sBegin = 0;
        bool have = slot < total;
        bool wallHint = true;
        if (QMODE && have) { const int2 q = sp.queueIn[slot]; i = q.x; s = q.y & CPF_Q_SUBSTEP; wallHint = (q.y & CPF_Q_WALL) != 0; have = s < sp.nSub; }
        sBegin = s;
        double4 p4 = make_double4(0.0, 0.0, 0.0, 0.0);
        int tet = -1;
        if (have) { p4 = ld_stream4(pv.pos + i); tet = ld_stream_i(pv.tet + i); }
        D3 P{ p4.x, p4.y, p4.z };
        double w = p4.w;
        const bool live = have && (w != 0.0);
        bool active = live;
        if (RNG != CPF_RNG_NONE) {
            Rng<RNG> rng;
            if (live) rng.open(pv, i, sp);
            rng.stage(s_xi + threadIdx.x, 128, sp.nSub, s, live);
            if constexpr (STATEFUL) { if (live) *stash = rng.st; }
        }
        Fast32 f;
        D3 O{ 0.0, 0.0, 0.0 }, disp{ 0.0, 0.0, 0.0 }; // disp: after an in-place reflection (leg = 1) the reflected END POINT
        D3 Phit{ 0.0, 0.0, 0.0 };
        int leg = 0;
        bool velDone = false;
        int stage = 0;                        // INTEG: 0 = move walk, 1..3 = walk to the stage point of k2..k4
        D3 vel{ 0.0, 0.0, 0.0 }, k1s = vel, k23 = vel, Pst = vel; // KEEPV: v_eff (reported velocity); INTEG: k1, k2 (+ k3), stage point
        bool velValid = false;
        WalkF ws;
        int cell = -1, visits = 0; // cell: the cell whose velocity moved the particle in its latest sub-step
        int org = -1;              // origin vertex id of `tet`
        bool needPro = true;
        if (active && tet >= 0) {
            f32_load(m, tet, f);
            org = first_origin<CPF_CFV_RUNTIME>(m, tet, f);
            O = ld_vertex(m.vpos, org);
            // C1, once per particle and launch, all lanes converged: the start point nothing in this kernel has certified
            // (later sub-steps start where C2 certified the end point, later visits where C3 certified the exit point)
            // FIN: an entry the filter itself refused goes straight to its exact sub-step (same arithmetic, same refusal) --
            // and all such lanes of the warp do so in the same iteration
            if ((FIN && !wallHint) || !start_point_clear(m, f, (float)(P.x - O.x), (float)(P.y - O.y), (float)(P.z - O.z))) {
                if (FIN) needExact = true;
                else { deferAt = s; active = false; }
            }
        }
        bool wallWait = false; // WALL: certified wall contact, waiting for the warp's next batched reflection
        while (__any_sync(0xffffffffu, active)) {
            if (FIN && active && needExact) {
                needExact = false;
                cell = m.tetcell ? __ldg(m.tetcell + tet) : org - m.nPoints;
                velX = ld_ucell(m, cell);
                D3 dX{ __dsub_rn(__fma_rn(sp.dt, velX.x, P.x), P.x), __dsub_rn(__fma_rn(sp.dt, velX.y, P.y), P.y),
                       __dsub_rn(__fma_rn(sp.dt, velX.z, P.z), P.z) };
                if (RNG != CPF_RNG_NONE) {
                    dX.x = __fma_rn((double)s_xi[(s * 3 + 0) * 128 + threadIdx.x], sp.randDisp, dX.x);
                    dX.y = __fma_rn((double)s_xi[(s * 3 + 1) * 128 + threadIdx.x], sp.randDisp, dX.y);
                    dX.z = __fma_rn((double)s_xi[(s * 3 + 2) * 128 + threadIdx.x], sp.randDisp, dX.z);
                }
                exact_substep_convex(m, P, dX, velX, tet, w, sp.reflect, ty);
                ty.exact++;
                exactVel = true;
                velValid = true;
                needPro = true;
                wallWait = false;
                leg = 0;
                if (++s >= sp.nSub || w == 0.0) active = false;
                else if (tet >= 0) { // back to the fp32 walk: C1 for a start point no C2 has certified
                    f32_load(m, tet, f);
                    org = first_origin<CPF_CFV_RUNTIME>(m, tet, f);
                    O = ld_vertex(m.vpos, org);
                    if (!start_point_clear(m, f, (float)(P.x - O.x), (float)(P.y - O.y), (float)(P.z - O.z))) needExact = true;
                } // tet < 0: frozen by the S1 of the next prologue
            }
            if (WALL) {
                // The reflection is long and rare per lane: lanes that reach a wall wait until CPF_WALL_BATCH of them
                // have gathered (or nobody else can move), then reflect together instead of one or two at a time.
                const unsigned waiting = __ballot_sync(0xffffffffu, active && wallWait);
                const unsigned running = __ballot_sync(0xffffffffu, active && !wallWait);
                if (wallWait && (__popc(waiting) >= CPF_WALL_BATCH || running == 0u)) {
                    wallWait = false;
                    D3 Eref, u = ld_ucell(m, cell);
                    if (KEEPV) u = vel;
                    if (wall_reflect_on_path(m, tet, ws.path, visits - 1, ws.cur, ws.wall_js, ws.wall_link, P, disp, Phit, Eref, u)) {
                        disp = Eref;
                        const int wallTet = ws.cur;
                        walkf_begin(ws, O, Phit, xsub(Eref, Phit), wallTet, ws.org); // leg 1: from the hit point (certified by C3), same tet
                        hops += visits;
                        visits = 0;
                        leg = 1;
                        // the velocity a particle leaves the call with is the reflected one (reflectInTet's u)
                        if (KEEPV) vel = u;
                        else if (sp.writeVel && s == sp.nSub - 1) { st_stream4(pv.vel + i, make_double4(u.x, u.y, u.z, -1.0)); velDone = true; }
                    } else {
                        hops += visits;
                        if (FIN) needExact = true;
                        else { deferAt = s; active = false; }
                    }
                }
            }
            if (active && !wallWait && needPro && !(FIN && needExact)) {
                if (tet < 0) { active = false; frzHit = true; } // S1: left the domain -> frozen (particles.cu:334-338), w := 0 below
                else {
                    cell = m.tetcell ? __ldg(m.tetcell + tet) : org - m.nPoints;
                    D3 u0;
                    if (VERT) u0 = vertex_velocity_exact(m, tet, P);
                    else u0 = ld_ucell(m, cell);
                    if (INTEG) { // k1 = v(P, tet); first stage point P + dt/2 * k1 (RK2 midpoint and RK4 alike)
                        k1s = u0;
                        Pst = axpy3(__dmul_rn(0.5, sp.dt), k1s, P);
                        walkf_begin(ws, O, P, xsub(Pst, P), tet, org);
                        stage = 1;
                    } else {
                        if (VERT) vel = u0;
                        disp = D3{ __dsub_rn(__fma_rn(sp.dt, u0.x, P.x), P.x), __dsub_rn(__fma_rn(sp.dt, u0.y, P.y), P.y),
                                   __dsub_rn(__fma_rn(sp.dt, u0.z, P.z), P.z) };
                        if (RNG != CPF_RNG_NONE) {
                            disp.x = __fma_rn((double)s_xi[(s * 3 + 0) * 128 + threadIdx.x], sp.randDisp, disp.x);
                            disp.y = __fma_rn((double)s_xi[(s * 3 + 1) * 128 + threadIdx.x], sp.randDisp, disp.y);
                            disp.z = __fma_rn((double)s_xi[(s * 3 + 2) * 128 + threadIdx.x], sp.randDisp, disp.z);
                        }
                        walkf_begin(ws, O, P, disp, tet, org);
                    }
                    visits = 0;
                    leg = 0;
                    needPro = false;
                }
            }
            if (active && !wallWait && !(FIN && needExact)) {
                ++visits;
                const int oc = visit_fast32<CPF_CFV_RUNTIME>(m, f, O, (WALL && leg) ? Phit : P, ws);
                if (INTEG && stage > 0 && (oc == CPF_V_DONE || oc == CPF_V_WALL)) {
                    // the stage point lies in ws.cur (or beyond a certified wall face of it): take that cell's velocity
                    hops += visits;
                    visits = 0;
                    D3 kx;
                    if (VERT) kx = vertex_velocity_exact(m, ws.cur, Pst);
                    else {
                        const int scell = m.tetcell ? __ldg(m.tetcell + ws.cur) : ws.org - m.nPoints;
                        kx = ld_ucell(m, scell);
                    }
                    if (ws.cur != tet) { // every walk of a sub-step starts from (P, tet)
                        f32_load(m, tet, f);
                        if (ws.org != org) O = ld_vertex(m.vpos, org);
                    }
                    bool last = true;
                    if (INTEG == CPF_RK2) vel = kx;
                    else if (stage == 1) { k23 = kx; Pst = axpy3(__dmul_rn(0.5, sp.dt), kx, P); last = false; }
                    else if (stage == 2) { k23 = xadd(k23, kx); Pst = axpy3(sp.dt, kx, P); last = false; }
                    else {
                        vel.x = __ddiv_rn(__fma_rn(2.0, k23.x, __dadd_rn(k1s.x, kx.x)), 6.0);
                        vel.y = __ddiv_rn(__fma_rn(2.0, k23.y, __dadd_rn(k1s.y, kx.y)), 6.0);
                        vel.z = __ddiv_rn(__fma_rn(2.0, k23.z, __dadd_rn(k1s.z, kx.z)), 6.0);
                    }
                    if (last) {
                        disp = D3{ __dsub_rn(__fma_rn(sp.dt, vel.x, P.x), P.x), __dsub_rn(__fma_rn(sp.dt, vel.y, P.y), P.y),
                                   __dsub_rn(__fma_rn(sp.dt, vel.z, P.z), P.z) };
                        if (RNG != CPF_RNG_NONE) {
                            disp.x = __fma_rn((double)s_xi[(s * 3 + 0) * 128 + threadIdx.x], sp.randDisp, disp.x);
                            disp.y = __fma_rn((double)s_xi[(s * 3 + 1) * 128 + threadIdx.x], sp.randDisp, disp.y);
                            disp.z = __fma_rn((double)s_xi[(s * 3 + 2) * 128 + threadIdx.x], sp.randDisp, disp.z);
                        }
                        walkf_begin(ws, O, P, disp, tet, org);
                        stage = 0;
                    } else {
                        walkf_begin(ws, O, P, xsub(Pst, P), tet, org);
                        ++stage;
                    }
                } else if (oc == CPF_V_DONE) {
                    velValid = true;
                    tet = ws.cur;
                    org = ws.org;
                    if (WALL && leg) { P = xadd(Phit, xsub(disp, Phit)); refl++; } // p = P_hit (S4) then p += E - P_hit (S5)
                    else P = xadd(P, disp);
                    hops += visits;
                    needPro = true;
                    exactVel = false;
                    if (++s >= sp.nSub) active = false;
                } else if (oc != CPF_V_HOP || visits >= 48) {
                    if (WALL && oc == CPF_V_WALL && visits <= 15 && leg == 0 && (!INTEG || stage == 0) && sp.reflect && ws.Dd < 10.f &&
                        m.patch_kind[-ws.wall_link - 1] != CPF_PATCH_ESCAPE) {
                        wallWait = true;
                    } else {
                        hops += visits;
                        if (FIN) needExact = true;
                        else { deferAt = s; active = false; }
                    }
                }
            }
        }
        if (live) {
            if (frzHit) { w = 0.0; frz++; } // frozen in a prologue (tet turns negative through exact sub-steps only)
            nsteps += (unsigned)(s - sBegin);
            if constexpr (STATEFUL) { if (deferAt < 0 && s >= sp.nSub) pv.rng[i] = *stash; }
            st_stream4(pv.pos + i, make_double4(P.x, P.y, P.z, w));
            st_stream_i(pv.tet + i, tet);
            if (KEEPV) {
                if (sp.writeVel && velValid && deferAt < 0) st_stream4(pv.vel + i, make_double4(vel.x, vel.y, vel.z, -1.0));
            } else if (FIN && exactVel) { // the particle's latest sub-step ran exactly: the velocity its reflector left
                if (sp.writeVel) st_stream4(pv.vel + i, make_double4(velX.x, velX.y, velX.z, -1.0));
            } else if (sp.writeVel && cell >= 0 && deferAt < 0 && !velDone) {
                st_stream4(pv.vel + i, vel4(ld_ucell(m, cell)));
            }
        }
        // deferral queue: one atomic per warp
        const unsigned mask = __ballot_sync(0xffffffffu, deferAt >= 0);
        if (mask) {
            const int lane = threadIdx.x & 31;
            int qb = 0;
            if (lane == 0) qb = (int)atomicAdd(sp.countOut, (unsigned)__popc(mask));
            qb = __shfl_sync(0xffffffffu, qb, 0);
            if (deferAt >= 0) sp.queueOut[qb + __popc(mask & ((1u << lane) - 1u))] = make_int2((int)i, deferAt);
        }
        if (!QMODE) break;
    }
    flush_counters(sp, refl + ty.refl, ty.exact, hops + ty.hops, nsteps, ty.esc, frz);
}

// k_lean<RNG,CFV>: the all-particles pass of the filtered policy for the reference's own configuration (Euler, cell
// value) -- k_fast<RNG,0,0,EULER,0> with the per-visit and per-sub-step instruction count cut down, because this is
// the kernel the step time is made of and it is bound by instruction issue and dependent-load latency, not by memory
// bandwidth:
//   * thread i = particle i, no queue input, no wall handling: one loop-carried counter (`left`) says whether a lane
//     is walking, finished or stopped;
//   * C1 once per particle (start_point_clear, all lanes converged), never inside the loop (visit_fast32);
//   * the cell velocity of the NEXT sub-step is requested into L1 (one predicated prefetch, which holds no dependency
//     barrier) from the common block of the visit that ends a sub-step, ahead of the divergent sections: ptxas drains every
//     barrier at a reconvergence point, so a LOAD placed there would be waited for before the exit-face section runs;
//   * CFV (cell id = origin vertex id - nPoints, every OpenFOAM decomposition) is a template parameter.
#ifndef CPF_LEAN_THREADS
#define CPF_LEAN_THREADS 128
#endif
#ifndef CPF_LEAN_BARY_TOP
#define CPF_LEAN_BARY_TOP 1
#endif
#ifndef CPF_LEAN_PFU
#define CPF_LEAN_PFU 1
#endif
#ifndef CPF_LEAN_RK4_BLOCKS
#define CPF_LEAN_RK4_BLOCKS 7
#endif
// dynamic shared memory of k_lean per CTA: the staged deviates [row][thread] (rows: 3 per sub-step; the stateless stream
// stages whole Philox blocks, i.e. up to 3 rows in front of and behind the chunk; no random walk: one row for the
// start-tet slots) and the XORWOW state after the chunk
#define CPF_LEAN_SMEM_BYTES(rows, xiBytes, stateful) ((size_t)CPF_LEAN_THREADS * ((size_t)(xiBytes) * (size_t)((rows) ? (rows) : 1) + ((stateful) ? sizeof(curandState_t) : 0)))
// rows staged for nSub sub-steps starting at global sub-step step0 (host and device agree through this one function)
__host__ __device__ inline unsigned lean_rows(int rng, int nSub, unsigned long long step0)
{
    if (rng == CPF_RNG_NONE) return 0u;
    if (rng != CPF_RNG_PHILOX) return 3u * (unsigned)nSub;
    const unsigned off = (unsigned)((3ull * step0) & 3ull);
    return 4u * ((off + 3u * (unsigned)nSub + 3u) >> 2);
}

// shared-window accessors: the loop keeps ONE running 32-bit address per lane instead of re-deriving
// base + (3 s + c) * NT + tid at every sub-step
template <typename Xi, int OFF> CPF_DEV double lds_xi(unsigned a)
{
    if constexpr (sizeof(Xi) == 4) {
        float r;
        asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(r) : "r"(a), "n"(OFF));
        return (double)r;
    } else {
        double r;
        asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(r) : "r"(a), "n"(OFF));
        return r;
    }
}
CPF_DEV void sts_i32(unsigned a, int v) { asm volatile("st.shared.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
CPF_DEV int lds_i32(unsigned a) { int r; asm volatile("ld.shared.s32 %0, [%1];" : "=r"(r) : "r"(a)); return r; }

//   * LOC = CPF_LOCATOR_BARY (RTX=true build): the same kernel around visit_bary32 -- the walk goes towards the end point
//     Q = P + disp (kept where the convex walk keeps disp), no start-point check, walls always deferred; its loop runs
//     the prologue at the top of the iteration of a sub-step's first visit instead of in place (see there).
// Loop shape: every iteration is ONE tet visit of every lane that still has work; a lane whose visit ends its sub-step
// runs S5 and the S1/S2 prologue of its next sub-step right there (the first prologue runs before the loop with all lanes
// converged), so the loop head carries no mode dispatch.  Per-lane loop state besides the walk: the running deviate
// address xs, `left` (> 0: sub-steps still to do, the lane is walking; 0: chunk finished; < 0: stopped with -left
// sub-steps to do -- refused, frozen or not started), the visit cap as a value of the visit counter, and the cell of the
// last prologue.  The start tet of the current sub-step -- needed only if the sub-step is refused -- waits in shared
// memory, in the slot of the sub-step's first deviate (already consumed by then).
//   * INTEG = CPF_RK2 / CPF_RK4 (extensions; convex locator): a sub-step is a chain of walks from the
//     same start (P, tet) -- one per stage point P + h k, then the move walk -- through the same loop: a stage walk that
//     ends (in a tet, or at a certified wall face of it) yields that cell's velocity and sets up the next walk.  The
//     start tet and its origin id wait in two shared-memory rows of their own, the stage cells of RK4 in three registers
//     (their velocities are re-read when v_eff is formed: no fp64 accumulators are carried through the walks).
template <int RNG, bool CFV, int LOC = CPF_LOCATOR_CONVEX, int INTEG = CPF_EULER>
__global__ void __launch_bounds__(CPF_LEAN_THREADS, (INTEG == CPF_RK4 ? CPF_LEAN_RK4_BLOCKS : CPF_FAST_MIN_BLOCKS) * 128 / CPF_LEAN_THREADS) k_lean(const MeshView m, const ParticleView pv, const StepParams sp)
{
    static_assert(!INTEG || LOC == CPF_LOCATOR_CONVEX, "stage walks are segment walks");
    constexpr bool BARY = LOC == CPF_LOCATOR_BARY;
    constexpr int NT = CPF_LEAN_THREADS;
    typedef typename Rng<RNG>::Xi Xi;
    constexpr bool STATEFUL = Rng<RNG>::STATEFUL;
    constexpr int ROW = NT * (int)sizeof(Xi);                      // bytes between two staged rows of a lane
    constexpr unsigned STEP = RNG == CPF_RNG_NONE ? 0u : 3u * ROW; // no deviates: xs stays on the lane's start-tet slot
    extern __shared__ double s_dyn[];
    const unsigned rows = lean_rows(RNG, sp.nSub, sp.step0);
    Xi *xi = reinterpret_cast<Xi *>(s_dyn) + threadIdx.x;
    curandState_t *stash = reinterpret_cast<curandState_t *>(xi - threadIdx.x + (size_t)NT * (rows + (INTEG ? 2u : 0u))) + threadIdx.x;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool have = i < pv.n;
    double4 p4 = make_double4(0.0, 0.0, 0.0, 0.0);
    int tet = -1;
    if (have) { p4 = ld_stream4(pv.pos + i); tet = ld_stream_i(pv.tet + i); }
    D3 P{ p4.x, p4.y, p4.z };
    double w = p4.w;
    const bool live = have && (w != 0.0);
    unsigned xs = (unsigned)__cvta_generic_to_shared(xi);
    const unsigned t0 = xs + rows * (unsigned)ROW; // INTEG: the rows of the sub-step's start tet and (t0 + ROW) its origin id
    if (RNG != CPF_RNG_NONE) {
        Rng<RNG> rng;
        // the stateless stream does not wait for the particle record: a dead slot's deviates are drawn and never read
        const bool drawFor = STATEFUL ? live : have;
        if (drawFor) rng.open(pv, i, sp);
        if constexpr (RNG == CPF_RNG_PHILOX) {
            xs += (unsigned)((3ull * sp.step0) & 3ull) * (unsigned)ROW; // the chunk's first deviate inside its first block
            rng.stage_blocks(xi, NT, rows >> 2, drawFor);
        } else rng.stage(xi, NT, sp.nSub, 0, live);
        if constexpr (STATEFUL) { if (live) *stash = rng.st; }
    }
    Fast32 f;
    D3 O{ 0.0, 0.0, 0.0 }, disp{ 0.0, 0.0, 0.0 };
    int cell = -1, left = -sp.nSub;
    unsigned hops = 0, cap = 0, frz = 0;
    bool atWall = false;
    constexpr int CF = CFV ? CPF_CFV_YES : CPF_CFV_NO;
    WalkF ws;
    ws.cur = tet;
    ws.org = -1;
    int stage = 0, c1 = -1, c2 = -1, c3 = -1; // INTEG: 0 = move walk, 1.. = walk to the stage point of k2..; RK4: cells of k1..k3
    // a walk along the fp64 displacement d from (P, ws.cur, ws.org, O)
    auto begin_walk = [&](const D3 &d) {
        ws.dx = (float)d.x; ws.dy = (float)d.y; ws.dz = (float)d.z;
        ws.Dd = fmaxf(fmaxf(fabsf(ws.dx), fabsf(ws.dy)), fabsf(ws.dz));
        walkf_rebase(ws, O, P);
        ws.t_in = 0.f;
        cap = hops + 47u; // the 48th visit of a walk must end it
    };
    // S2 with the effective velocity v: disp = (P + dt v) - P (+ random walk), and the move walk
    auto begin_move = [&](const D3 &v) {
        disp = D3{ __dsub_rn(__fma_rn(sp.dt, v.x, P.x), P.x), __dsub_rn(__fma_rn(sp.dt, v.y, P.y), P.y),
                   __dsub_rn(__fma_rn(sp.dt, v.z, P.z), P.z) };
        if (RNG != CPF_RNG_NONE) {
            disp.x = __fma_rn(lds_xi<Xi, 0>(xs), sp.randDisp, disp.x);
            disp.y = __fma_rn(lds_xi<Xi, ROW>(xs), sp.randDisp, disp.y);
            disp.z = __fma_rn(lds_xi<Xi, 2 * ROW>(xs), sp.randDisp, disp.z);
        }
        begin_walk(disp);
    };
    // INTEG: a stage walk has ended in ws.cur -- that cell's velocity, back to the start of the sub-step, next walk
    auto stage_done = [&]() {
        ++hops; // the stage walk's final visit
        const int cs = CFV ? ws.org - m.nPoints : __ldg(m.tetcell + ws.cur);
        const D3 kx = ld_ucell(m, cs);
        const int tet0 = lds_i32(t0), org0 = lds_i32(t0 + ROW);
        if (ws.cur != tet0) { // every walk of a sub-step starts from (P, tet)
            f32_load(m, tet0, f);
            if (ws.org != org0) O = ld_vertex(m.vpos, org0);
        }
        ws.cur = tet0;
        ws.org = org0;
        D3 v = kx;
        bool last = true;
        if (INTEG == CPF_RK4) {
            if (stage == 1) { c2 = cs; last = false; begin_walk(xsub(axpy3(__dmul_rn(0.5, sp.dt), kx, P), P)); }
            else if (stage == 2) { c3 = cs; last = false; begin_walk(xsub(axpy3(sp.dt, kx, P), P)); }
            else {
                const D3 k1 = ld_ucell(m, c1), k23 = xadd(ld_ucell(m, c2), ld_ucell(m, c3));
                v.x = __ddiv_rn(__fma_rn(2.0, k23.x, __dadd_rn(k1.x, kx.x)), 6.0);
                v.y = __ddiv_rn(__fma_rn(2.0, k23.y, __dadd_rn(k1.y, kx.y)), 6.0);
                v.z = __ddiv_rn(__fma_rn(2.0, k23.z, __dadd_rn(k1.z, kx.z)), 6.0);
            }
        }
        if (last) {
            if (sp.writeVel && left == 1) st_stream4(pv.vel + i, vel4(v)); // v_eff of the call's last sub-step (a finisher that redoes it writes again)
            begin_move(v);
            stage = 0;
        } else ++stage;
    };
    // S1 + S2 of the sub-step whose deviates sit at xs, then the walk set up from (P, ws.cur, ws.org, O)
    auto begin_substep = [&]() {
        cell = CFV ? ws.org - m.nPoints : __ldg(m.tetcell + ws.cur);
        const D3 u0 = ld_ucell(m, cell);
        if constexpr (INTEG != CPF_EULER) { // k1 = v(P, tet); first stage point P + dt/2 k1 (RK2 midpoint and RK4 alike)
            sts_i32(t0, ws.cur);
            sts_i32(t0 + ROW, ws.org);
            c1 = cell;
            stage = 1;
            begin_walk(xsub(axpy3(__dmul_rn(0.5, sp.dt), u0, P), P));
            return;
        }
        disp = D3{ __dsub_rn(__fma_rn(sp.dt, u0.x, P.x), P.x), __dsub_rn(__fma_rn(sp.dt, u0.y, P.y), P.y),
                   __dsub_rn(__fma_rn(sp.dt, u0.z, P.z), P.z) };
        if (RNG != CPF_RNG_NONE) {
            disp.x = __fma_rn(lds_xi<Xi, 0>(xs), sp.randDisp, disp.x);
            disp.y = __fma_rn(lds_xi<Xi, ROW>(xs), sp.randDisp, disp.y);
            disp.z = __fma_rn(lds_xi<Xi, 2 * ROW>(xs), sp.randDisp, disp.z);
        }
        sts_i32(xs, ws.cur); // behind the read of the same slot
        if (BARY) {
            disp = xadd(P, disp); // Q: the point the barycentric walk looks for, and the S5 result
            ws.dx = ws.dy = ws.dz = 0.f;
            ws.Dd = 0.f;
            walkf_rebase(ws, O, disp);
        } else {
            ws.dx = (float)disp.x; ws.dy = (float)disp.y; ws.dz = (float)disp.z;
            ws.Dd = fmaxf(fmaxf(fabsf(ws.dx), fabsf(ws.dy)), fabsf(ws.dz));
            walkf_rebase(ws, O, P);
        }
        ws.t_in = 0.f;
        cap = hops + 47u; // the 48th visit of a sub-step must end it
    };
    if (live) {
        if (tet < 0) { w = 0.0; frz = 1u; } // S1: left the domain -> frozen (particles.cu:334-338)
        else {
            f32_load(m, tet, f);
            ws.org = first_origin<CF>(m, tet, f);
            O = ld_vertex(m.vpos, ws.org);
            if (BARY || start_point_clear(m, f, (float)(P.x - O.x), (float)(P.y - O.y), (float)(P.z - O.z))) left = sp.nSub;
            else sts_i32(INTEG ? t0 : xs, tet); // refused before its first sub-step
        }
    }
    constexpr bool PFU = CFV && !BARY && INTEG == CPF_EULER && CPF_LEAN_PFU;
    if constexpr (BARY && CPF_LEAN_BARY_TOP) {
        // The barycentric walk has no exit-face section for the prologue to hide behind, and is faster (measured: +8 %, and 67
        // registers) with the loop of the first RTX=true version: the prologue of a sub-step runs at the TOP of the iteration
        // of its first visit.  (The convex walk is 3 % slower that way.)
        bool due = left > 0;
        while (__any_sync(0xffffffffu, left > 0)) {
            if (left > 0) {
                if (due) { begin_substep(); due = false; }
                const int oc = visit_bary32<CF>(m, f, O, disp, ws, hops >= cap);
                if (oc == CPF_V_HOP) ++hops;
                else if (oc == CPF_V_DONE) { P = disp; xs += STEP; --left; due = true; }
                else { left = -left; atWall = oc == CPF_V_WALL; }
            }
        }
    } else {
    if (left > 0) begin_substep();
    while (__any_sync(0xffffffffu, left > 0)) {
        if (left > 0) {
            const int oc = BARY ? visit_bary32<CF>(m, f, O, disp, ws, hops >= cap) : visit_fast32<CF, PFU>(m, f, O, P, ws, hops >= cap);
            if (oc == CPF_V_HOP) ++hops;
            else if (INTEG != CPF_EULER && stage > 0 && (oc == CPF_V_DONE || oc == CPF_V_WALL)) stage_done();
            else if (oc == CPF_V_DONE) {
                P = BARY ? disp : xadd(P, disp);
                xs += STEP;
                if (--left > 0) begin_substep();
            } else { left = -left; atWall = oc == CPF_V_WALL; }
        }
    }
    }
    const bool deferred = left < 0 && live && frz == 0u;
    const int s = sp.nSub - abs(left); // sub-steps completed here
    if (live && frz == 0u) hops += (unsigned)s + (left < 0 ? 1u : 0u); // tet visits = hops + one final visit per sub-step (+ the refused one)
    if (live) {
        if constexpr (STATEFUL) { if (left == 0) pv.rng[i] = *stash; }
        st_stream4(pv.pos + i, make_double4(P.x, P.y, P.z, w));
        st_stream_i(pv.tet + i, deferred ? lds_i32(INTEG ? t0 : xs) : ws.cur);
        if (INTEG == CPF_EULER && sp.writeVel && cell >= 0 && !deferred) {
            st_stream4(pv.vel + i, vel4(ld_ucell(m, cell)));
        }
    }
    const unsigned mask = __ballot_sync(0xffffffffu, deferred); // deferral queue: one atomic per warp
    if (mask) {
        const int lane = threadIdx.x & 31;
        int qb = 0;
        if (lane == 0) qb = (int)atomicAdd(sp.countOut, (unsigned)__popc(mask));
        qb = __shfl_sync(0xffffffffu, qb, 0);
        if (deferred) sp.queueOut[qb + __popc(mask & ((1u << lane) - 1u))] = make_int2((int)i, s | (atWall ? CPF_Q_WALL : 0));
    }
    flush_counters(sp, 0u, 0u, hops, (unsigned)s, 0u, frz);
}

// Point values from the cell field (OpenFOAM volPointInterpolation, interior weights 1/|p - C|),
// then uvert = [point values..., cell values...] for the vertex (cellPoint-style) interpolation.
__global__ void k_point_interp(int nPoints, int nCells, const int *__restrict__ off, const int *__restrict__ cells,
                               const double4 *__restrict__ vpos, const double4 *__restrict__ ucell, double *__restrict__ uvert)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nPoints) {
        const D3 P = ld_vertex(vpos, i);
        double sumw = 0.0, ax = 0.0, ay = 0.0, az = 0.0;
        for (int q = off[i]; q < off[i + 1]; ++q) {
            const int c = cells[q];
            const D3 d = xsub(P, ld_vertex(vpos, nPoints + c));
            const double wgt = __drcp_rn(__dsqrt_rn(__fma_rn(d.z, d.z, __fma_rn(d.y, d.y, __dmul_rn(d.x, d.x)))));
            sumw = __dadd_rn(sumw, wgt);
            const double4 uc = ucell[c];
            ax = __fma_rn(wgt, uc.x, ax);
            ay = __fma_rn(wgt, uc.y, ay);
            az = __fma_rn(wgt, uc.z, az);
        }
        uvert[3ll * i] = __ddiv_rn(ax, sumw);
        uvert[3ll * i + 1] = __ddiv_rn(ay, sumw);
        uvert[3ll * i + 2] = __ddiv_rn(az, sumw);
    } else if (i < nPoints + nCells) {
        const int c = i - nPoints;
        const double4 uc = ucell[c];
        uvert[3ll * i] = uc.x; uvert[3ll * i + 1] = uc.y; uvert[3ll * i + 2] = uc.z;
    }
}

int launch_point_interp(cpf_context *ctx)
{
    if (!ctx->d_pc_off) return fail(ctx, CPF_ERR_INVALID, "vertex interpolation needs cpf_mesh_upload_poly with cfg.interp = CPF_INTERP_VERTEX, or cpf_update_vertex_velocity");
    if (!ctx->d_uvert) CPF_CUDA(ctx, cudaMalloc(&ctx->d_uvert, sizeof(double) * 3 * (size_t)ctx->nVerts));
    const int n = (int)ctx->nVerts;
    k_point_interp<<<(n + 127) / 128, 128, 0, ctx->stream>>>(ctx->nPoints, (int)ctx->nCells, ctx->d_pc_off, ctx->d_pc_cells, ctx->d_vpos,
                                                           ctx->d_ucell[ctx->ucur], ctx->d_uvert);
    ctx->launches++;
    CPF_CUDA(ctx, cudaGetLastError());
    return CPF_OK;
}

// src/initCuda.H:184-199: the one cudaAdvect right after seeding; its only lasting effect is to
// deactivate particles whose initial location failed (tet < 0) and to fill vel for VTU 0.
__global__ void __launch_bounds__(128) k_initial_advect(const MeshView m, const ParticleView pv, unsigned long long *counters)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pv.n) return;
    double4 p4 = pv.pos[i];
    if (p4.w == 0.0) return;
    const int tet = pv.tet[i];
    if (tet < 0) { p4.w = 0.0; pv.pos[i] = p4; atomicAdd(counters + CNT_FROZEN, 1ull); return; }
    const int4 v = ld_int4(m.tetv, tet);
    const int cell = tet_cell(m, tet, v);
    pv.vel[i] = vel4(ld_ucell(m, cell));
}

__global__ void k_init_rng(curandState_t *st, long long n, unsigned long long seed, unsigned long long idBase)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) curand_init(seed, idBase + (unsigned long long)i, 0ull, &st[i]); // particles.cu:537, subsequence = global particle id
}

// parity tooling: the deviates of the next k sub-steps, [k][n][3] in ORIGINAL particle order; the stream is not advanced
template <int RNG> __global__ void k_debug_normals(const ParticleView pv, const StepParams sp, int k, double *xi)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= pv.n) return;
    Rng<RNG> rng;
    rng.open(pv, i, sp);
    const long long o = pv.pid[i];
    for (int q = 0; q < k; ++q) {
        double a = 0, b = 0, c = 0;
        rng.draw(q, a, b, c);
        double *dst = xi + 3 * ((long long)q * pv.n + o);
        dst[0] = a; dst[1] = b; dst[2] = c;
    }
}

// ------------------------------------------------------------------------------------------------
// host launchers
// ------------------------------------------------------------------------------------------------
#define CPF_RNG_SWITCH(rng, CALL)                                              \
    switch (rng) {                                                             \
    case CPF_RNG_XORWOW: { constexpr int R = CPF_RNG_XORWOW; CALL; } break;     \
    case CPF_RNG_PHILOX: { constexpr int R = CPF_RNG_PHILOX; CALL; } break;     \
    default: { constexpr int R = CPF_RNG_NONE; CALL; } break;                   \
    }

// Filtered policy (ConvexPoly locator), random walk R, integrator I, interpolation V.
//   Euler, cell value (the reference's configuration): k_lean over all particles -> ONE finishing pass over its refusals
//   (k_fast<..., FIN = 1>: fp32 walk with in-place wall reflection, a refused sub-step in the reference's arithmetic in place).
//   RK2 / RK4 / vertex interpolation: k_fast over all particles -> wall-capable k_fast over its refusals -> k_general.
// With the stateful XORWOW stream a kernel commits the generator state only for particles that finish their chunk in it;
// every later kernel re-draws from the chunk-start state (staged deviates / Rng::skip), so the stream stays the reference's.
template <int R, int I, int V>
static int launch_filtered(cpf_context *ctx, const MeshView &m, const ParticleView &pv, const StepParams &sp, dim3 grid, int nSub)
{
    typedef typename Rng<R>::Xi Xi;
    constexpr bool STATEFUL = Rng<R>::STATEFUL;
    cudaStream_t st = ctx->stream;
    CPF_CUDA(ctx, cudaMemsetAsync(ctx->d_queue_count, 0, sizeof(unsigned) * 64, st));
    // queue kernels: one resident wave on the 148 SMs of a B200 (grid-stride loops inside)
    const dim3 wgrid(std::min<unsigned>(grid.x, 148u * ((I == CPF_EULER && !V) ? CPF_FIN_MIN_BLOCKS : CPF_WALL_MIN_BLOCKS)));
    const size_t xiBytes = R == CPF_RNG_NONE ? 0 : sizeof(Xi) * 3 * 128 * (size_t)nSub + (STATEFUL ? sizeof(curandState_t) * 128 : 0);
    StepParams a = sp; // all-particles pass: refusals into queue 0
    a.queueOut = ctx->d_queue[0]; a.countOut = ctx->d_queue_count;
    StepParams b = sp; // first queue pass: queue 0 -> queue 1
    b.queueIn = ctx->d_queue[0]; b.countIn = ctx->d_queue_count;
    if constexpr (I == CPF_EULER && !V) {
        const dim3 lgrid((unsigned)((pv.n + CPF_LEAN_THREADS - 1) / CPF_LEAN_THREADS));
        const size_t lb = CPF_LEAN_SMEM_BYTES(lean_rows(R, nSub, sp.step0), sizeof(Xi), STATEFUL);
        if (m.tetcell == nullptr) k_lean<R, true><<<lgrid, CPF_LEAN_THREADS, lb, st>>>(m, pv, a);
        else k_lean<R, false><<<lgrid, CPF_LEAN_THREADS, lb, st>>>(m, pv, a);
        k_fast<R, 2, 1, CPF_EULER, 0, 1><<<wgrid, 128, xiBytes, st>>>(m, pv, b);
        ctx->launches += 2;
    } else {
        const dim3 egrid(std::min<unsigned>(grid.x, 148u * 8u));
        b.queueOut = ctx->d_queue[1]; b.countOut = ctx->d_queue_count + 1;
        StepParams z = sp; // the rest in the reference's arithmetic
        z.queueIn = ctx->d_queue[1]; z.countIn = ctx->d_queue_count + 1;
        if constexpr (!V) { // cell value: the lean all-particles pass with stage walks
            const dim3 lgrid((unsigned)((pv.n + CPF_LEAN_THREADS - 1) / CPF_LEAN_THREADS));
            const size_t lb = CPF_LEAN_SMEM_BYTES(lean_rows(R, nSub, sp.step0) + 2u, sizeof(Xi), STATEFUL);
            if (m.tetcell == nullptr) k_lean<R, true, CPF_LOCATOR_CONVEX, I><<<lgrid, CPF_LEAN_THREADS, lb, st>>>(m, pv, a);
            else k_lean<R, false, CPF_LOCATOR_CONVEX, I><<<lgrid, CPF_LEAN_THREADS, lb, st>>>(m, pv, a);
        } else
        k_fast<R, 0, 0, I, V><<<grid, 128, xiBytes, st>>>(m, pv, a);
        k_fast<R, 2, 1, I, V><<<wgrid, 128, xiBytes, st>>>(m, pv, b);
        k_general<R, 2><<<egrid, 128, 0, st>>>(m, pv, z);
        ctx->launches += 3;
    }
    return CPF_OK;
}

// RTX=true build on the filtered policy: all-particles pass around visit_bary32, then the exact finisher (k_exact<BARY>)
// for what it refused -- wall contacts included: RTreflection always runs in the reference's arithmetic.
template <int R>
static int launch_filtered_bary(cpf_context *ctx, const MeshView &m, const ParticleView &pv, const StepParams &sp, dim3 grid, int nSub)
{
    typedef typename Rng<R>::Xi Xi;
    constexpr bool STATEFUL = Rng<R>::STATEFUL;
    cudaStream_t st = ctx->stream;
    CPF_CUDA(ctx, cudaMemsetAsync(ctx->d_queue_count, 0, sizeof(unsigned) * 64, st));
    StepParams a = sp;
    a.queueOut = ctx->d_queue[0]; a.countOut = ctx->d_queue_count;
    const dim3 lgrid((unsigned)((pv.n + CPF_LEAN_THREADS - 1) / CPF_LEAN_THREADS));
    const size_t lb = CPF_LEAN_SMEM_BYTES(lean_rows(R, nSub, sp.step0), sizeof(Xi), STATEFUL);
    if (m.tetcell == nullptr) k_lean<R, true, CPF_LOCATOR_BARY><<<lgrid, CPF_LEAN_THREADS, lb, st>>>(m, pv, a);
    else k_lean<R, false, CPF_LOCATOR_BARY><<<lgrid, CPF_LEAN_THREADS, lb, st>>>(m, pv, a);
    StepParams z = sp;
    z.queueIn = ctx->d_queue[0]; z.countIn = ctx->d_queue_count;
    k_exact<CPF_LOCATOR_BARY, R, 2><<<dim3(std::min<unsigned>(grid.x, 148u * 8u)), 128, 0, st>>>(m, pv, z);
    ctx->launches += 2;
    return CPF_OK;
}

template <int R>
static int launch_filtered_rng(cpf_context *ctx, const MeshView &m, const ParticleView &pv, const StepParams &sp, dim3 grid, int nSub)
{
    const bool vert = ctx->cfg.interp == CPF_INTERP_VERTEX;
    if (ctx->cfg.integrator == CPF_RK2) return vert ? launch_filtered<R, CPF_RK2, 1>(ctx, m, pv, sp, grid, nSub) : launch_filtered<R, CPF_RK2, 0>(ctx, m, pv, sp, grid, nSub);
    if (ctx->cfg.integrator == CPF_RK4) return vert ? launch_filtered<R, CPF_RK4, 1>(ctx, m, pv, sp, grid, nSub) : launch_filtered<R, CPF_RK4, 0>(ctx, m, pv, sp, grid, nSub);
    return vert ? launch_filtered<R, CPF_EULER, 1>(ctx, m, pv, sp, grid, nSub) : launch_filtered<R, CPF_EULER, 0>(ctx, m, pv, sp, grid, nSub);
}

// most sub-steps one launch sequence may fuse: the staged deviates of a chunk must fit the 48 KB of shared memory a
// kernel gets without opting in (3 x 4 B per sub-step and thread, stateless streams; 3 x 8 B + the 48-byte state, XORWOW)
int max_fused_substeps(const cpf_context *ctx) { return ctx->cfg.rng == CPF_RNG_XORWOW ? (ctx->cfg.integrator == CPF_EULER ? 14 : 13) : 16; } // RK: two more rows (start tet, origin)
// library default (cfg.fuse_substeps == 0); XORWOW: 10 sub-steps = 36 KB of staged fp64 deviates + states, 6 CTAs per SM
int default_fused_substeps(const cpf_context *ctx) { return ctx->cfg.rng == CPF_RNG_XORWOW ? 10 : 16; }

int launch_substeps(cpf_context *ctx, int nSub, double dt, bool writeVel)
{
    if (ctx->n == 0 || nSub <= 0) return CPF_OK;
    const MeshView m = mesh_view(ctx);
    const ParticleView pv = particle_view(ctx);
    StepParams sp;
    sp.nSub = nSub;
    sp.dt = dt;
    sp.randDisp = sqrt((2.00 * ctx->cfg.diffusion_coeff) * dt);
    sp.reflect = ctx->cfg.reflect_wall;
    sp.writeVel = writeVel ? 1 : 0;
    sp.seed = ctx->cfg.seed;
    sp.step0 = ctx->step_index;
    sp.idBase = ctx->id_base;
    sp.counters = ctx->d_counters;
    sp.queueIn = sp.queueOut = nullptr;
    sp.countIn = sp.countOut = nullptr;
    const int rng = ctx->cfg.rng;
    if (rng == CPF_RNG_XORWOW && !ctx->rng_ready) {
        int rc = launch_init_rng(ctx);
        if (rc) return rc;
    }
    cudaStream_t st = ctx->stream;
    const dim3 grid((unsigned)((ctx->n + 127) / 128));
    if (ctx->profiling) {
        if (ctx->profUsed + 2 > ctx->profEvents.size()) {
            cudaEvent_t a, b;
            CPF_CUDA(ctx, cudaEventCreate(&a));
            CPF_CUDA(ctx, cudaEventCreate(&b));
            ctx->profEvents.push_back(a);
            ctx->profEvents.push_back(b);
        }
        CPF_CUDA(ctx, cudaEventRecord(ctx->profEvents[ctx->profUsed], st));
    }
    sp.integrator = ctx->cfg.integrator;
    sp.interp = ctx->cfg.interp;
    const bool filteredOk = ctx->cfg.locator == CPF_LOCATOR_CONVEX && ctx->cfg.path == CPF_PATH_FILTERED && ctx->filter_ok;
    if (ctx->cfg.locator == CPF_LOCATOR_BARY) {
        if (ctx->cfg.path == CPF_PATH_FILTERED && ctx->filter_ok) {
            int rc = CPF_OK;
            CPF_RNG_SWITCH(rng, (rc = launch_filtered_bary<R>(ctx, m, pv, sp, grid, nSub)));
            if (rc) return rc;
        } else {
            CPF_RNG_SWITCH(rng, (k_exact<CPF_LOCATOR_BARY, R, 0><<<grid, 128, 0, st>>>(m, pv, sp)));
            ctx->launches++;
        }
    } else if (!filteredOk) {
        if (ctx->cfg.interp != CPF_INTERP_TET || ctx->cfg.integrator != CPF_EULER) { CPF_RNG_SWITCH(rng, (k_general<R, 0><<<grid, 128, 0, st>>>(m, pv, sp))); }
        else { CPF_RNG_SWITCH(rng, (k_exact_convex<R, 0><<<grid, 128, 0, st>>>(m, pv, sp))); }
        ctx->launches++;
    } else {
        int rc = CPF_OK;
        CPF_RNG_SWITCH(rng, (rc = launch_filtered_rng<R>(ctx, m, pv, sp, grid, nSub)));
        if (rc) return rc;
    }
    if (ctx->profiling) {
        CPF_CUDA(ctx, cudaEventRecord(ctx->profEvents[ctx->profUsed + 1], st));
        ctx->profUsed += 2;
    }
    ctx->step_index += (unsigned long long)nSub;
    CPF_CUDA(ctx, cudaGetLastError());
    return CPF_OK;
}

int launch_initial_advect(cpf_context *ctx, double)
{
    if (ctx->n == 0) return CPF_OK;
    k_initial_advect<<<(unsigned)((ctx->n + 127) / 128), 128, 0, ctx->stream>>>(mesh_view(ctx), particle_view(ctx), ctx->d_counters);
    ctx->launches++;
    CPF_CUDA(ctx, cudaGetLastError());
    return CPF_OK;
}

int launch_init_rng(cpf_context *ctx)
{
    for (int b = 0; b < 2; ++b)
        if (!ctx->d_rng[b]) CPF_CUDA(ctx, cudaMalloc(&ctx->d_rng[b], sizeof(curandState_t) * (size_t)ctx->n));
    // states are indexed by ORIGINAL particle id; only valid before any sort or via pid scatter
    if (ctx->permuted) return fail(ctx, CPF_ERR_INVALID, "cpf_init_rng must run before particles are sorted");
    k_init_rng<<<(unsigned)((ctx->n + 127) / 128), 128, 0, ctx->stream>>>(ctx->d_rng[ctx->pcur], ctx->n, ctx->cfg.seed, ctx->id_base);
    ctx->launches++;
    CPF_CUDA(ctx, cudaGetLastError());
    ctx->rng_ready = true;
    return CPF_OK;
}

int launch_debug_normals(cpf_context *ctx, int k, double *d_xi)
{
    const ParticleView pv = particle_view(ctx);
    StepParams sp{};
    sp.seed = ctx->cfg.seed;
    sp.step0 = ctx->step_index;
    sp.idBase = ctx->id_base;
    const unsigned grid = (unsigned)((ctx->n + 127) / 128);
    if (ctx->cfg.rng == CPF_RNG_XORWOW) {
        if (!ctx->rng_ready) { int rc = launch_init_rng(ctx); if (rc) return rc; }
        k_debug_normals<CPF_RNG_XORWOW><<<grid, 128, 0, ctx->stream>>>(particle_view(ctx), sp, k, d_xi);
    } else if (ctx->cfg.rng == CPF_RNG_PHILOX) {
        k_debug_normals<CPF_RNG_PHILOX><<<grid, 128, 0, ctx->stream>>>(pv, sp, k, d_xi);
    } else {
        CPF_CUDA(ctx, cudaMemsetAsync(d_xi, 0, sizeof(double) * 3 * (size_t)ctx->n * (size_t)k, ctx->stream));
    }
    ctx->launches++;
    CPF_CUDA(ctx, cudaGetLastError());
    return CPF_OK;
}

} // namespace cpf
