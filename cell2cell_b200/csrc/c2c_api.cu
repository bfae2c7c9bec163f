// libc2c_b200.so -- the C ABI declared in include/c2c_b200.h: argument checks, workspace carving, launch geometry and
// the per-iteration launch sequence of the factorization.  Kernels live in c2c_contract.cuh (the streaming passes over
// the tensor) and c2c_small.cuh (everything that runs out of L2).
//
// One NCP iteration on an N-way tensor with two trailing (in-plane) modes, unmasked:
//   plane_contract  X -> T[P x R]                      (pass 1 over X; serves every leading mode: T depends on A_{N-2},
//   for each leading mode k: lead_mttkrp T -> M_k,      A_{N-1} only, which do not change until the trailing updates)
//                            update A_k, Gram S_k
//   fiber_contract  X -> U[E x R] (partials + reduce)   (pass 2 over X; w = prod of the UPDATED leading factors)
//   trailing mode N-2: trail_mttkrp U -> M, update; trailing mode N-1: trail_mttkrp U (with updated A_{N-2}) -> M, update
//   error scalar + convergence flag on the device.
// Exactly the Gauss-Seidel order of tensorly's non_negative_parafac (mode 0, 1, ..., N-1) -- each M_n is the MTTKRP
// with the factors as they stand when mode n is updated -- with 2 reads of X per iteration instead of N * R.
// Masked: the imputed entries follow the current factors, so T (resp. U) is recomputed before every mode: N passes.
#include "c2c_contract.cuh"
#include "c2c_stream.cuh"
#include "c2c_mma.cuh"
#include "c2c_small.cuh"
#include "c2c_update.cuh"
#include "c2c_sums.cuh"
#include "c2c_corr.cuh"
#include "c2c_plan.cuh"
#include "c2c_multi.cuh"
#include "c2c_xgpu.cuh"
#include "c2c_metrics.cuh"
#include "c2c_admm.cuh"
#include "../../include/c2c_b200.h"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#define C2C_VERSION 100
#ifndef C2C_MMA_PLANE_LEAD
#define C2C_MMA_PLANE_LEAD 2            // plane_mma takes over this many ranks before fiber_mma does
#endif
#ifndef C2C_MMA_DEFAULT_MIN_RANK
#define C2C_MMA_DEFAULT_MIN_RANK 13
#endif

static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
static int cuda_fail(cudaError_t e, const char* what)
{
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    return (int)e;
}
// kernels enqueued by this library (all threads); graph replays add the captured iteration's launch count
static std::atomic<long long> g_launches{0};
#define CK(call, what) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return cuda_fail(e_, what); } while (0)
#define CKL(what) do { g_launches.fetch_add(1, std::memory_order_relaxed); CK(cudaGetLastError(), what); } while (0)

extern "C" int c2c_version(void) { return C2C_VERSION; }
extern "C" const char* c2c_last_error(void) { return g_err; }
extern "C" long long c2c_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

// Executable graphs whose launches may still be in flight when the run that enqueued them returns (a run that is never polled
// -- check_every <= 0 -- does not wait for the device any more: the host part of the NEXT run, its capture and instantiation,
// then overlaps this run's iterations).  An exec object must outlive its launches, so it is parked here with an event
// recorded after its last launch and destroyed by a later call once that event has completed.
struct PendingGraph { cudaGraphExec_t exec; cudaGraph_t graph; cudaEvent_t done; int device; };
static std::mutex g_pending_mu;
static std::vector<PendingGraph> g_pending;

static void reap_pending_graphs(bool wait)
{
    std::lock_guard<std::mutex> lock(g_pending_mu);
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) { (void)cudaGetLastError(); return; }
    size_t keep = 0;
    for (size_t i = 0; i < g_pending.size(); ++i) {
        PendingGraph& p = g_pending[i];
        bool finished = false;
        if (p.device == dev) {
            const cudaError_t q = wait ? cudaEventSynchronize(p.done) : cudaEventQuery(p.done);
            if (q == cudaSuccess) finished = true;
            else if (q != cudaErrorNotReady) { (void)cudaGetLastError(); finished = true; }   // a failed context: nothing left to protect
        }
        if (finished) { cudaGraphExecDestroy(p.exec); cudaGraphDestroy(p.graph); cudaEventDestroy(p.done); }
        else g_pending[keep++] = p;
    }
    g_pending.resize(keep);
}

// waits for and releases every parked graph of the current device (tests / orderly shutdown; never required for correctness)
extern "C" int c2c_release_pending(void) { reap_pending_graphs(true); return 0; }

static int sm_count()
{
    static thread_local int dev_cached = -1, n_cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != dev_cached) {
        if (cudaDeviceGetAttribute(&n_cached, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n_cached = 148;
        dev_cached = dev;
    }
    return n_cached;
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int ldr_of(int rank) { return (rank + 3) / 4 * 4; }

// ---------------------------------------------------------------------------------------------------------------------
// geometry of a tensor seen as P planes of E = I2 x I3 entries
// ---------------------------------------------------------------------------------------------------------------------
struct Geo {
    int ndim; long long shape[C2C_MAX_DIMS];
    long long P, E; int I2, I3; long long maxI;
    int vec_plane, vec_fiber;
    // fiber_contract launch
    int fc_block, fc_nvec, fc_nvec_total, fc_nslots, fc_gx, fc_gy; long long fc_ppc; long long fc_nparts;
    int np2;     // rank pairs per chunk
    int nchunk;  // rank chunks
    // shared-memory-staged (bulk-copy) passes; *_ok = 0 -> the register-staged kernels do the whole pass
    int ps_ok, ps_ns, ps_spw, ps_rs, ps_ring, ps_nwarps, ps_rp, ps_pitch, ps_grid; long long ps_nstage; size_t ps_smem;
    int fs_ok, fs_ns, fs_tp, fs_nthr, fs_nslots, fs_merge, fs_wall, fs_ncw, fs_grid; long long fs_nstage, fs_spc; size_t fs_smem;
    long long max_parts;
    // tensor-core (TMA + tcgen05) passes, c2c_mma.cuh; mma_ok = 0 -> the FFMA kernels do the pass
    int mma_ok, mma_n;         // mma_ok: plane_mma applies;  mma_fiber_ok: fiber_mma applies too
    int pm_n;                  // N of plane_mma: 16, 24 (ranks 17-24) or 32; fiber_mma runs at mma_n (16 / 32)
    int mma_fiber_ok;
    int pm_nstages, pm_grid, pm_nkb; long long pm_ntiles; size_t pm_smem;
    int fm_es, fm_na, fm_na_total, fm_ntile, fm_nstages, fm_grid, fm_nj, fm_stack, fm_in_smem, fm_ring2; long long fm_nst; size_t fm_smem;
};

#define C2C_SMEM_BUDGET (200 * 1024)

static long long gcd_ll(long long a, long long b) { while (b) { long long t = a % b; a = b; b = t; } return a; }

static bool use_stream()
{
    static int v = -1;
    if (v < 0) { const char* e = getenv("C2C_B200_NO_STREAM"); v = (e && e[0] == '1') ? 0 : 1; }
    return v != 0;
}

// launch geometry of plane_stream / fiber_stream (see c2c_stream.cuh); both need the whole rank in one chunk
static void stream_geometry(Geo& g, int rank)
{
    g.ps_ok = g.fs_ok = 0;
    g.max_parts = g.fc_nparts;
    if (g.nchunk != 1 || !use_stream()) return;
    const int nsm = sm_count();
    const long long plane_bytes = g.E * 4;
    const long long align_unit = 4 / gcd_ll(g.E, 4);                 // planes per stage must keep stages 16-byte sized
    {   // ---- plane_stream ----
        const int vec = g.vec_plane;
        const int ns = (vec == 4 && g.np2 <= C2C_NS2_MAXNP2 && (g.I3 / 4) % 2 == 0) ? 2 : 1;
        const int nsets = g.I3 / (vec * ns);
        const int G = nsets < 32 ? nsets : 32;
        const int ppw = 32 / G;
        g.ps_ns = ns;
        g.ps_rp = (2 * g.np2 + 3) / 4 * 4;
        // A2 / A3 copies + barriers, and the per-warp scratch rows of the plane reduction (32 x (NP2 | 1) float2 per warp)
        const size_t red_bytes = (size_t)C2C_ST_MAXWARPS * 32 * (g.np2 | 1) * 8;
        const size_t fixed = align_up((size_t)(g.I2 + g.I3) * g.ps_rp * sizeof(float) + (size_t)C2C_ST_MAXWARPS * C2C_ST_MAXRING * 8, 128);
        const long long target = 10 * 1024;                          // bytes per stage
        long long spw = 0, rs = 1, pitch = g.E;
        if (ppw * plane_bytes > target && nsets <= 32) {
            // large planes: a stage is a row slab of ppw planes; the slab must stay 16-byte sized and aligned
            for (rs = 1; rs <= g.I2; ++rs)
                if (g.I2 % rs == 0 && ((g.I2 / rs) * (long long)g.I3) % 4 == 0 && ppw * (plane_bytes / rs) <= target) break;
            if (rs <= g.I2 && g.E % 4 == 0) {
                spw = ppw;
                pitch = (g.I2 / rs) * (long long)g.I3;
                if (ppw > 1 && (G * vec) % 4 == 0)                   // consecutive planes -> consecutive banks
                    while (pitch % 32 != (G * vec) % 32) pitch += 4;
            } else rs = 1;
        }
        if (spw == 0) {
            rs = 1; pitch = g.E;
            const long long unit = (long long)ppw / gcd_ll(ppw, align_unit) * align_unit;
            long long mult = target / (unit * plane_bytes);
            if (mult < 1) mult = 1;
            spw = unit * mult;
            // small tensors: shrink the stages until every warp of the machine gets at least two plane groups
            while (mult > 1 && g.P / spw < 2LL * nsm * C2C_ST_MAXWARPS) { mult /= 2; spw = unit * mult; }
        }
        const size_t stage_bytes = (size_t)(spw * pitch * 4);
        if (fixed < C2C_SMEM_BUDGET / 2 && stage_bytes <= 24 * 1024) {
            long long nw = C2C_ST_MAXWARPS;
            long long ring = ((long long)C2C_SMEM_BUDGET - (long long)fixed - (long long)red_bytes) / (nw * (long long)stage_bytes);
            if (ring < 2) { nw = 4; ring = ((long long)C2C_SMEM_BUDGET - (long long)fixed - (long long)red_bytes) / (nw * (long long)stage_bytes); }
            if (ring > C2C_ST_MAXRING) ring = C2C_ST_MAXRING;
            const long long ngroups = g.P / spw;
            if (ring >= 2 && ngroups >= 2LL * nsm * nw) {
                g.ps_ok = 1; g.ps_spw = (int)spw; g.ps_rs = (int)rs; g.ps_ring = (int)ring; g.ps_nwarps = (int)nw;
                g.ps_pitch = (int)pitch; g.ps_nstage = ngroups; g.ps_grid = nsm;
                g.ps_smem = fixed + (size_t)nw * ring * stage_bytes + red_bytes;
            }
        }
    }
    {   // ---- fiber_stream ----
        const int vec = g.vec_fiber;
        const long long nvec = g.E / vec;
        const int ns = (vec == 4 && g.np2 <= C2C_NS2_MAXNP2 && nvec % 2 == 0) ? 2 : 1;
        const long long nthr = nvec / ns;
        const int maxthr = (ns == 2 ? g.np2 <= 5 : g.np2 <= 12) ? 512 : 256;
        long long nslots = (maxthr - 32) / nthr;
        if (nslots > 4) nslots = 4;
        if (nthr >= 64 && nslots >= 1) {
            const int ldr = ldr_of(rank);
            const long long per_trip = 2 * nslots;                           // planes consumed per trip of the unrolled loop
            long long unit = align_unit / gcd_ll(align_unit, per_trip) * per_trip;   // ... and stages stay 16-byte sized
            long long mult = (52 * 1024) / (unit * plane_bytes);
            if (mult < 1) mult = 1;
            const long long tp = unit * mult;
            const long long nstage = g.P / tp;
            const long long spc = (nstage + nsm - 1) / nsm;
            // Khatri-Rao rows: all of a CTA's planes up front in shared memory when they fit, else per stage by the producer
            const size_t x_stage = ((size_t)tp * g.E + 31) / 32 * 32;
            const size_t wall_bytes = (size_t)spc * tp * ldr * sizeof(float);
            int w_all = 128 + (size_t)C2C_FS_STAGES * x_stage * sizeof(float) + wall_bytes <= (size_t)C2C_SMEM_BUDGET + 16 * 1024;
            const size_t stage_floats = w_all ? x_stage : ((size_t)tp * g.E + (size_t)tp * ldr + 31) / 32 * 32;
            const size_t smem = 128 + (size_t)C2C_FS_STAGES * stage_floats * sizeof(float) + (w_all ? wall_bytes : 0);
            if (smem <= (size_t)C2C_SMEM_BUDGET + 16 * 1024 && nstage >= 4LL * nsm) {
                g.fs_ok = 1; g.fs_tp = (int)tp; g.fs_ns = ns; g.fs_nthr = (int)nthr; g.fs_nslots = (int)nslots; g.fs_wall = w_all;
                g.fs_ncw = (int)((nthr * nslots + 31) / 32);
                g.fs_nstage = nstage; g.fs_smem = smem;
                g.fs_spc = spc;
                g.fs_grid = (int)((nstage + g.fs_spc - 1) / g.fs_spc);
                const size_t acc_bytes = (size_t)ns * vec * g.np2 * 8;                 // accumulators of one thread
                g.fs_merge = (nslots - 1) * nthr * (long long)acc_bytes <= (long long)((size_t)C2C_FS_STAGES * stage_floats * sizeof(float)) ? 1 : 0;
                const long long parts = (long long)g.fs_grid * (g.fs_merge ? 1 : nslots) + g.fc_nslots;
                if (parts > g.max_parts) g.max_parts = parts;
            }
        }
    }
}


// launch geometry of plane_mma / fiber_mma (c2c_mma.cuh).  Conditions: whole rank in one MMA (R <= 32 here, so that the
// fused updates apply), in-plane size a multiple of 4 floats (TMA needs 16-byte row pitches; a ragged last 32-float atom is
// zero-filled by the TMA unit), enough
// planes to give every SM at least one 128-plane tile.  C2C_B200_NO_MMA=1 turns the path off, C2C_B200_MMA_MIN_RANK sets
// the smallest rank that takes it.
#define C2C_MMA_SMEM_BUDGET (225 * 1024)
static std::atomic<int> g_mma_min_rank{-1};
static int mma_min_rank()
{
    int v = g_mma_min_rank.load(std::memory_order_relaxed);
    if (v < 0) {
        const char* off = getenv("C2C_B200_NO_MMA");
        const char* e = getenv("C2C_B200_MMA_MIN_RANK");
        v = (off && off[0] == '1') ? 1 << 30 : (e ? atoi(e) : C2C_MMA_DEFAULT_MIN_RANK);
        if (v < 1) v = 1;
        g_mma_min_rank.store(v, std::memory_order_relaxed);
    }
    return v;
}
extern "C" int c2c_set_mma_min_rank(int rank)
{
    const int prev = mma_min_rank();
    g_mma_min_rank.store(rank < 1 ? 1 : rank, std::memory_order_relaxed);
    return prev;
}

static int env_int(const char* name, int dflt)
{
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}
static int mma_dbg() { static int v = -1; if (v < 0) v = env_int("C2C_B200_MMA_DBG", 0); return v; }

static int mma_epoch()
{
    static int v = -1;
    if (v < 0) { const char* e = getenv("C2C_B200_MMA_EPOCH"); v = e ? atoi(e) : 0; }
    return v;
}

static void mma_geometry(Geo& g, int rank)
{
    g.mma_ok = 0; g.mma_fiber_ok = 0;
    const int nsm = sm_count();
    // shapes the FFMA kernels only take with narrow (< 128-bit) loads go to the tensor pipe at every rank; the plane pass
    // switches two ranks before the fiber pass (measured at 60x2000x40x40: plane_mma 0.135 ms vs plane_stream 0.145 / 0.155 ms
    // at ranks 11 / 12, fiber_mma 0.167 vs fiber_stream 0.145 at rank 12)
    static const int plane_lead = env_int("C2C_B200_MMA_PLANE_LEAD", C2C_MMA_PLANE_LEAD);
    const int min_rank = g.vec_plane < 4 ? 1 : (mma_min_rank() > plane_lead ? mma_min_rank() - plane_lead : 1);
    if (rank < min_rank || rank > 32 || g.nchunk != 1 || mma_min_rank() > 32 || !c2c_mma_available()) return;
    if (g.E % 4 != 0 || g.E < 128 || g.P < 128LL * nsm || g.P >= (1LL << 31) - 256) return;
    const int n = rank <= 16 ? 16 : 32;
    g.mma_n = n;
    const int ldr = ldr_of(rank);
    static const int plane_n24 = env_int("C2C_B200_NO_MMA_N24", 0) ? 0 : 1;
    const int pn = (rank > 16 && rank <= 24 && plane_n24) ? 24 : n;
    g.pm_n = pn;
    {   // plane_mma: B ring and the factor copies are fixed, the rest of shared memory is the X ring
        const size_t xb = 128 * 128;
        const size_t fixed = c2c_plane_mma_smem(pn, g.I2, g.I3, 0);
        if (fixed + 3 * xb > C2C_MMA_SMEM_BUDGET) return;
        int nx = (int)((C2C_MMA_SMEM_BUDGET - fixed) / (xb + 16));
        if (nx > 12) nx = 12;
        g.pm_nstages = nx;
        g.pm_nkb = (int)((g.E + 31) / 32);
        g.pm_ntiles = (g.P + 127) / 128;
        g.pm_grid = nsm;                                 // P >= 128 * nsm: a CTA's share spans at least one whole tile
        g.pm_smem = c2c_plane_mma_smem(pn, g.I2, g.I3, nx);
    }
    {   // fiber_mma: split the in-plane positions into `es` parts (es divides the SM count when it can) so that the CTA's
        // running sums and >= 5 X slots of 8 planes fit in shared memory and its accumulators + lo(X) ring in TMEM --
        // preferably with the stacked B tile (2N accumulator columns per tile)
        const int na_total = (int)((g.E + 31) / 32);
        int best_es = 0, best_nx = 0, best_stack = 0, best_in_smem = 1;
        // pass 0: stacked B tile with whole-machine splits of <= 2 parts (big stages amortise the per-stage hand-offs);
        // pass 1: unstacked, >= 5 X slots; pass 2: whatever fits (>= 3 X slots)
        for (int pass = 0; pass < 3 && best_es == 0; ++pass) {
            for (int cand = 1; cand <= 8; ++cand) {
                if (nsm % cand != 0) continue;
                const int na = (na_total + cand - 1) / cand;
                if (na < 4) break;
                const int ntile = (na + 3) / 4;
                // TMEM: accumulators (N or 2N columns per tile) + the X ring in TMEM: hi | lo (16 columns per tile and prep slot, all
                // MMAs from TMEM: N = 16) or lo only (8 columns, hi(X) read from shared memory: N = 32)
                const int a_cols = n == 16 ? 16 : 8;
                const int stack = ntile * (2 * n + a_cols * (n == 16 ? 3 : C2C_MMA_RING2)) <= 512;
                if (pass == 0 && (!stack || cand > 2)) continue;
                if (ntile > C2C_FM_MAXTILE || ntile * (n + a_cols * (n == 16 ? 3 : C2C_MMA_RING2)) > 512) continue;
                // running sums in shared memory when that still leaves >= 5 X slots, else in the global partial
                int in_smem = 1;
                size_t fixed = c2c_fiber_mma_smem(n, na, ldr, 0);
                int fit = fixed > C2C_MMA_SMEM_BUDGET ? 0 : (int)((C2C_MMA_SMEM_BUDGET - fixed) / ((size_t)na * 1024 + 16));
                if (fit < 5) {
                    in_smem = 0;
                    fixed = c2c_fiber_mma_smem(n, na, 0, 0);
                    if (fixed > C2C_MMA_SMEM_BUDGET) continue;
                    fit = (int)((C2C_MMA_SMEM_BUDGET - fixed) / ((size_t)na * 1024 + 16));
                }
                if (fit >= 5 || (pass == 2 && fit >= 3)) {
                    best_es = cand; best_nx = fit > 12 ? 12 : fit; best_stack = stack; best_in_smem = in_smem; break;
                }
            }
        }
        // a ragged last atom (in-plane size not a multiple of 32 floats) needs one small TMA box per atom: slower than the
        // FFMA pass below the rank where that pass leaves the HBM roof
        const bool fiber_worth = rank >= mma_min_rank() || (g.E % 32 == 0 && g.vec_plane < 4);
        if (best_es != 0 && nsm / best_es >= 1 && fiber_worth) {
            g.fm_es = best_es; g.fm_na_total = na_total; g.fm_na = (na_total + best_es - 1) / best_es; g.fm_ntile = (g.fm_na + 3) / 4;
            g.fm_nstages = best_nx; g.fm_stack = best_stack; g.fm_in_smem = best_in_smem;
            {
                const int acc = g.fm_ntile * (best_stack ? 2 * n : n);
                int r2 = (512 - acc) / (g.fm_ntile * (n == 16 ? 16 : 8));
                g.fm_ring2 = r2 > C2C_MMA_RING2 ? C2C_MMA_RING2 : r2;
            }
            // three prep groups fill the ring round-robin: with a ring shallower than that a group could run two laps ahead of
            // the MMA warp and pass an mbarrier wait on the parity of an EARLIER phase (the cause of the wrong sums of the
            // all-TMEM N = 32 experiment of round 1, whose ring was 2 deep) -- such a geometry is refused
            if (g.fm_ring2 >= C2C_MMA_PREP_GROUPS) {
                g.fm_nj = nsm / best_es;
                g.fm_grid = g.fm_nj * best_es;
                g.fm_nst = (g.P + 7) / 8;
                g.fm_smem = c2c_fiber_mma_smem(n, g.fm_na, best_in_smem ? ldr : 0, best_nx);
                if ((long long)g.fm_nj > g.max_parts) g.max_parts = g.fm_nj;
                g.mma_fiber_ok = 1;
            }
        }
    }
    g.mma_ok = 1;
}

static int fiber_geometry(Geo& g, int np2_chunk)
{
    const int maxT = C2C_FC_MAXTHREADS(np2_chunk);
    g.fc_nvec_total = (int)(g.E / g.vec_fiber);
    if (g.fc_nvec_total >= maxT) {
        g.fc_nvec = maxT; g.fc_nslots = 1; g.fc_block = maxT;
        g.fc_gy = (g.fc_nvec_total + maxT - 1) / maxT;
    } else {
        g.fc_nvec = g.fc_nvec_total; g.fc_nslots = maxT / g.fc_nvec;
        if ((long long)g.fc_nslots > g.P) g.fc_nslots = (int)g.P;
        g.fc_block = (g.fc_nslots * g.fc_nvec + 31) / 32 * 32;
        g.fc_gy = 1;
    }
    // enough CTAs to fill the machine (occupancy ~ 2048 / block CTAs per SM), each owning >= one tile of planes per slot
    const int per_sm = 2048 / g.fc_block > 0 ? (2048 / g.fc_block > 4 ? 4 : 2048 / g.fc_block) : 1;
    long long want = (long long)sm_count() * per_sm / g.fc_gy;
    if (want < 1) want = 1;
    const long long min_ppc = (long long)g.fc_nslots * C2C_FC_UNROLL;
    long long gx = (g.P + min_ppc - 1) / min_ppc;
    if (gx > want) gx = want;
    if (gx < 1) gx = 1;
    g.fc_ppc = (g.P + gx - 1) / gx;
    g.fc_gx = (int)((g.P + g.fc_ppc - 1) / g.fc_ppc);
    g.fc_nparts = (long long)g.fc_gx * g.fc_nslots;
    return 0;
}

static int make_geo(Geo& g, int ndim, const int64_t* shape, int rank)
{
    if (ndim < 3 || ndim > C2C_MAX_DIMS) return fail(-1, "ndim must be in [3, %d], got %d", C2C_MAX_DIMS, ndim);
    if (rank < 1 || rank > C2C_MAX_RANK) return fail(-2, "rank must be in [1, %d], got %d", C2C_MAX_RANK, rank);
    g.ndim = ndim; g.P = 1; g.maxI = 0;
    for (int i = 0; i < ndim; ++i) {
        if (shape[i] < 1 || shape[i] > 0x7fffffffLL) return fail(-3, "shape[%d] = %lld out of range", i, (long long)shape[i]);
        g.shape[i] = shape[i];
        if (i < ndim - 2) g.P *= shape[i];
        if (shape[i] > g.maxI) g.maxI = shape[i];
    }
    g.I2 = (int)shape[ndim - 2]; g.I3 = (int)shape[ndim - 1];
    g.E = (long long)g.I2 * g.I3;
    g.vec_plane = (g.I3 % 4 == 0) ? 4 : ((g.I3 % 2 == 0) ? 2 : 1);
    g.vec_fiber = (g.E % 4 == 0) ? 4 : ((g.E % 2 == 0) ? 2 : 1);
    const int np2 = (rank + 1) / 2;
    g.nchunk = (np2 + C2C_MAX_NP2 - 1) / C2C_MAX_NP2;
    g.np2 = (np2 + g.nchunk - 1) / g.nchunk;
    // shared memory of plane_contract: A2 as (a,a) pairs + A3
    const size_t smem = ((size_t)g.I2 * 4 * g.np2 + (size_t)g.I3 * 2 * g.np2) * sizeof(float);
    if (smem > 200 * 1024)
        return fail(-4, "trailing modes %d x %d at rank %d need %zu B of shared memory (> 200 KiB): put the two smallest modes last",
                    g.I2, g.I3, rank, smem);
    fiber_geometry(g, g.np2);
    stream_geometry(g, rank);
    mma_geometry(g, rank);
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// workspace
// ---------------------------------------------------------------------------------------------------------------------
struct Ws {
    float* T; float* part; float* U; float* M; float* G; float* Bmat; float* Xh; float* W;
    float* Tw; float* Uw; float* cpart;      // masked runs, dense pass + missing-entry correction (c2c_corr.cu, c2c_plan.cu)
    int cparts;                              // partial U buffers reserved in cpart
    float* Dmat; float* Ymat; float* Ypart;  // restricted Grams of the mask plan's separable model (c2c_plan.cu)
    float* Wmat;                             // Khatri-Rao rows of the leading factors, [P][ldr] (residual lists, c2c_plan.cu)
    double* S; double* gram_part; double* iprod_part; double* normX2; double* sums_part; double* scratch_d;
    int* flags;
    size_t bytes;
};

#define C2C_SUMSQ_BLOCKS 1184     // 148 x 8
#define C2C_SUMS_BLOCKS  1184

static void carve(Ws& w, const Geo& g, int rank, int masked, char* base)
{
    const int ldr = ldr_of(rank);
    size_t off = 0;
    auto take = [&](size_t bytes) { char* p = base ? base + off : nullptr; off += align_up(bytes, 256); return p; };
    const long long nblk = (g.maxI + C2C_UPD_ROWS - 1) / C2C_UPD_ROWS;
    w.T = (float*)take((size_t)g.P * ldr * sizeof(float));
    w.part = (float*)take((size_t)g.max_parts * g.E * ldr * sizeof(float));
    w.W = nullptr;
    w.U = (float*)take((size_t)g.E * ldr * sizeof(float));
    w.M = (float*)take((size_t)g.maxI * ldr * sizeof(float));
    w.G = (float*)take((size_t)ldr * ldr * sizeof(float));
    w.Bmat = (float*)take(masked ? (size_t)g.E * ldr * sizeof(float) : 0);
    // ranks beyond one chunk: the masked passes need the full-rank reconstruction, so the imputed tensor is materialised
    w.Xh = (float*)take((masked && g.nchunk > 1) ? (size_t)g.P * g.E * sizeof(float) : 0);
    // only for shapes the dense + correction form takes (ncp_run_impl makes the same test): otherwise 320 partials of E x ldr
    // floats would be reserved for nothing (2.6 GB at a 255 x 255 plane)
    const bool scan_ok = c2c_fiber_corr_smem(g.E, ldr) <= C2C_CORR_SMEM_MAX;          // mask-scanning corrections (c2c_corr.cu)
    const bool plan_ok = g.E <= C2C_PLAN_MAX_E;                                          // mask-plan corrections (c2c_plan.cu)
    const bool corr_ws = masked && g.nchunk == 1 && rank <= 32 && (scan_ok || plan_ok);
    w.Tw = (float*)take(corr_ws ? (size_t)g.P * ldr * sizeof(float) : 0);
    w.Uw = (float*)take(corr_ws ? (size_t)g.E * ldr * sizeof(float) : 0);
    // partial U corrections: up to C2C_CORR_MAX_PARTS, fewer when a partial is large (at most ~256 MB in all)
    long long cparts = (256LL << 20) / (g.E * ldr * (long long)sizeof(float));
    if (cparts > C2C_CORR_MAX_PARTS) cparts = C2C_CORR_MAX_PARTS;
    if (cparts < 4) cparts = 4;
    if (scan_ok) cparts = C2C_CORR_MAX_PARTS;
    w.cparts = corr_ws ? (int)cparts : 0;
    w.cpart = (float*)take(corr_ws ? (size_t)cparts * g.E * ldr * sizeof(float) : 0);
    const bool plan_ws = corr_ws && plan_ok;
    const int C0 = (int)g.shape[0];
    w.Dmat = (float*)take(plan_ws ? (size_t)(C0 + 1) * ldr * ldr * sizeof(float) : 0);
    w.Ymat = (float*)take(plan_ws ? (size_t)(C0 + 1) * ldr * ldr * sizeof(float) : 0);
    w.Ypart = (float*)take(plan_ws ? (size_t)C0 * c2c_lead_ctx_parts(C0, sm_count()) * 2 * ldr * ldr * sizeof(float) : 0);
    w.Wmat = (float*)take(plan_ws ? (size_t)g.P * ldr * sizeof(float) : 0);
    w.S = (double*)take((size_t)2 * g.ndim * ldr * ldr * sizeof(double));        // two banks (c2c_update.cuh)
    const long long nblk_g = nblk > 512 ? nblk : 512;             // lead_update runs at most 2 CTAs per SM, capped at 512
    w.gram_part = (double*)take((size_t)nblk_g * ldr * ldr * sizeof(double));
    w.iprod_part = (double*)take((size_t)nblk * sizeof(double));
    w.normX2 = (double*)take(sizeof(double));
    w.sums_part = (double*)take((size_t)C2C_SUMS_BLOCKS * 5 * sizeof(double));
    w.scratch_d = (double*)take(16 * sizeof(double));
    w.flags = (int*)take(4 * sizeof(int));       // [0]: lead_update arrival counter (zero between launches)
    w.bytes = off;
}

extern "C" size_t c2c_ncp_workspace_bytes(int ndim, const int64_t* shape, int rank, int masked)
{
    Geo g;
    if (make_geo(g, ndim, shape, rank) != 0) return 0;
    Ws w;
    carve(w, g, rank, masked, nullptr);
    return w.bytes;
}

static int get_ws(Ws& w, const Geo& g, int rank, int masked, void* ws, size_t ws_bytes)
{
    carve(w, g, rank, masked, nullptr);
    if (ws == nullptr || ws_bytes < w.bytes)
        return fail(-5, "workspace too small: need %zu bytes, got %zu (query c2c_ncp_workspace_bytes)", w.bytes, ws_bytes);
    if (((uintptr_t)ws & 255) != 0) return fail(-6, "workspace must be 256-byte aligned");
    carve(w, g, rank, masked, (char*)ws);
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// launch helpers
// ---------------------------------------------------------------------------------------------------------------------
static void fill_lead(LeadInfo& li, const Geo& g, const float* const* A)
{
    li.nl = g.ndim - 2;
    for (int k = 0; k < C2C_MAX_LEAD; ++k) { li.dim[k] = 1; li.fac[k] = nullptr; }
    for (int k = 0; k < li.nl; ++k) { li.dim[k] = (int)g.shape[k]; li.fac[k] = A[k]; }
}

static ContractArgs base_args(const Geo& g, const float* X, const uint8_t* mask, const float* const* A, int ldr, int rank,
                              const int* done)
{
    ContractArgs a;
    memset(&a, 0, sizeof(a));
    a.X = X; a.mask = mask; a.P = g.P; a.p0 = 0; a.I2 = g.I2; a.I3 = g.I3; a.R = rank; a.ldr = ldr; a.r0 = 0;
    a.A2 = A[g.ndim - 2]; a.A3 = A[g.ndim - 1];
    fill_lead(a.lead, g, A);
    a.done = done;
    return a;
}

static int plane_grid(const Geo& g, int vec)
{
    const int nstrips = g.I3 / vec;
    const int G = nstrips < 32 ? nstrips : 32;
    const int ppw = 32 / G;
    const long long warps = (g.P + ppw - 1) / ppw;
    long long blocks = (warps + (C2C_PC_THREADS / 32) - 1) / (C2C_PC_THREADS / 32);
    const long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    return (int)(blocks < 1 ? 1 : blocks);
}

// how a pass is launched inside the captured iteration
struct PassOpts {
    int pdl;                   // programmatic dependent launch: the pass may start under its predecessor (it waits before reading factors)
    bool t_zeroed;             // plane_mma: T was cleared by the previous fiber_mma (else a memset is enqueued)
    float* zero_T;             // fiber_mma: buffer to clear for the next plane_mma, `zero_n` floats
    long long zero_n;
    const XgComm* xg;          // context-sharded run: the partial U is summed over the ranks inside the reduce kernel (c2c_xgpu.cuh)
    double* xg_S; const int* xg_state; int xg_N;
    bool main_only;            // c2c_time_pass_kernel: only the streaming kernel itself (no memset, tail, reduce); -60 if the
                               // shape takes the register-staged kernels
};
static const PassOpts kPlainPass = {0, false, nullptr, 0, nullptr, nullptr, nullptr, 0, false};
static int use_pdl() { static int v = -1; if (v < 0) v = env_int("C2C_B200_NO_PDL", 0) ? 0 : 1; return v; }

// T[P x ldr] = plane contraction of Xh with the two trailing factors
static int launch_plane(const Geo& g, ContractArgs a, bool masked, float* T, cudaStream_t st, const PassOpts& po = kPlainPass)
{
    if (!masked && g.mma_ok && a.p0 == 0 && a.r0 == 0) {
        CUtensorMap map;
        const int rc = c2c_make_plane_map(&map, a.X, g.P, g.E);
        if (rc != 0) return fail(-40, "cuTensorMapEncodeTiled (plane map) failed: %d", rc);
        PlaneMmaArgs pa;
        memset(&pa, 0, sizeof(pa));
        pa.P = g.P; pa.E = (int)g.E; pa.I2 = g.I2; pa.I3 = g.I3; pa.R = a.R; pa.ldr = a.ldr; pa.A2 = a.A2; pa.A3 = a.A3; pa.T = T;
        pa.done = a.done; pa.nkb = g.pm_nkb; pa.ntiles = g.pm_ntiles; pa.nx = g.pm_nstages; pa.dbg = mma_dbg();
        if (!po.t_zeroed && !po.main_only) CK(cudaMemsetAsync(T, 0, (size_t)g.P * a.ldr * sizeof(float), st), "memset T");
        const cudaError_t e = c2c_launch_plane_mma(g.pm_n, map, pa, g.pm_grid, g.pm_smem, st, po.t_zeroed ? po.pdl : 0);
        if (e != cudaSuccess) return cuda_fail(e, "plane_mma launch");
        g_launches.fetch_add(1, std::memory_order_relaxed);
        return 0;
    }
    if (!masked && g.ps_ok) {
        StreamArgs sa;
        memset(&sa, 0, sizeof(sa));
        sa.c = a; sa.c.out = T; sa.W = nullptr; sa.spw = g.ps_spw; sa.rs = g.ps_rs; sa.nring = g.ps_ring; sa.nstage = g.ps_nstage; sa.rp = g.ps_rp; sa.nwarps = g.ps_nwarps;
        const int block = 32 * g.ps_nwarps;
        sa.pitch = g.ps_pitch;
        const cudaError_t e = g.vec_plane == 4 ? (g.ps_ns == 2 ? c2c_launch_plane_stream_v4_s2(g.np2, sa, g.ps_grid, block, g.ps_smem, st, po.pdl)
                                                               : c2c_launch_plane_stream_v4_s1(g.np2, sa, g.ps_grid, block, g.ps_smem, st, po.pdl))
                            : g.vec_plane == 2 ? c2c_launch_plane_stream_v2_s1(g.np2, sa, g.ps_grid, block, g.ps_smem, st, po.pdl)
                                               : c2c_launch_plane_stream_v1_s1(g.np2, sa, g.ps_grid, block, g.ps_smem, st, po.pdl);
        if (e != cudaSuccess) return cuda_fail(e, "plane_stream launch");
        g_launches.fetch_add(1, std::memory_order_relaxed);
        a.p0 = g.ps_nstage * g.ps_spw;                 // the tail (< one stage of planes) goes to the register-staged kernel
        if (a.p0 >= g.P || po.main_only) return 0;
    }
    if (po.main_only) return fail(-60, "this shape / rank takes the register-staged plane kernel");
    const int vec = g.vec_plane;
    const size_t smem = ((size_t)g.I2 * 4 * g.np2 + (size_t)g.I3 * 2 * g.np2) * sizeof(float);
    int grid = plane_grid(g, vec);
    if (a.p0 > 0) { const long long tail = g.P - a.p0; if (grid > tail) grid = (int)tail; }
    a.out = T;
    for (int c = 0; c < g.nchunk; ++c) {
        a.r0 = c * 2 * g.np2;
        if (a.r0 >= a.R) break;
        cudaError_t e;
        if (masked) {
            e = vec == 4 ? c2c_launch_plane_v4_m1(g.np2, a, grid, smem, st)
              : vec == 2 ? c2c_launch_plane_v2_m1(g.np2, a, grid, smem, st) : c2c_launch_plane_v1_m1(g.np2, a, grid, smem, st);
        } else {
            e = vec == 4 ? c2c_launch_plane_v4_m0(g.np2, a, grid, smem, st)
              : vec == 2 ? c2c_launch_plane_v2_m0(g.np2, a, grid, smem, st) : c2c_launch_plane_v1_m0(g.np2, a, grid, smem, st);
        }
        if (e != cudaSuccess) return cuda_fail(e, "plane_contract launch");
        g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    return 0;
}

// U[E x ldr] = fiber contraction of Xh with the Khatri-Rao rows of the leading factors (partials + fixed-order reduce)
static int launch_fiber(const Geo& g, ContractArgs a, bool masked, const Ws& w, cudaStream_t st, const PassOpts& po = kPlainPass)
{
    const int vec = g.vec_fiber;
    a.out = w.part;
    long long nparts = g.fc_nparts;
    dim3 grid(g.fc_gx, g.fc_gy);
    long long ppc = g.fc_ppc;
    bool legacy = true;
    if (!masked && g.mma_fiber_ok) {
        CUtensorMap map;
        const int boxes2d = (g.E % 32) != 0;
        const int rc = boxes2d ? c2c_make_fiber_map2d(&map, a.X, g.P, g.E) : c2c_make_fiber_map(&map, a.X, g.P, g.E, g.fm_na);
        if (rc != 0) return fail(-41, "cuTensorMapEncodeTiled (fiber map) failed: %d", rc);
        FiberMmaArgs fa;
        memset(&fa, 0, sizeof(fa));
        fa.P = g.P; fa.E = (int)g.E; fa.R = a.R; fa.ldr = a.ldr; fa.lead = a.lead; fa.part = w.part; fa.done = a.done;
        fa.es = g.fm_es; fa.na_total = g.fm_na_total; fa.na = g.fm_na; fa.ntile = g.fm_ntile; fa.nx = g.fm_nstages; fa.dbg = mma_dbg(); fa.nst_total = g.fm_nst; fa.epoch = mma_epoch() > 0 ? mma_epoch() : 16;
        fa.boxes2d = boxes2d; fa.sums_in_smem = g.fm_in_smem; fa.ring2 = g.fm_ring2;
        fa.zero_ptr = reinterpret_cast<float4*>(po.zero_T); fa.zero_n4 = po.zero_n / 4;
        const cudaError_t e = c2c_launch_fiber_mma(g.mma_n, g.fm_stack, g.mma_n == 16 ? 1 : 0, map, fa, g.fm_grid, g.fm_smem, st, po.pdl);
        if (e != cudaSuccess) return cuda_fail(e, "fiber_mma launch");
        g_launches.fetch_add(1, std::memory_order_relaxed);
        nparts = g.fm_nj;
        legacy = false;
        if (po.main_only) return 0;
    } else if (!masked && g.fs_ok) {
        StreamArgs sa;
        memset(&sa, 0, sizeof(sa));
        sa.c = a; sa.W = w.W; sa.spw = g.fs_tp; sa.nstage = g.fs_nstage; sa.nwarps = g.fs_ncw; sa.stages_per_cta = g.fs_spc;
        const int block = 32 * (g.fs_ncw + 1);
        const cudaError_t e = vec == 4 ? (g.fs_ns == 2 ? c2c_launch_fiber_stream_v4_s2(g.np2, sa, g.fs_nthr, g.fs_nslots, g.fs_merge, g.fs_wall, g.fs_grid, block, g.fs_smem, st, po.pdl)
                                                       : c2c_launch_fiber_stream_v4_s1(g.np2, sa, g.fs_nthr, g.fs_nslots, g.fs_merge, g.fs_wall, g.fs_grid, block, g.fs_smem, st, po.pdl))
                            : vec == 2 ? c2c_launch_fiber_stream_v2_s1(g.np2, sa, g.fs_nthr, g.fs_nslots, g.fs_merge, g.fs_wall, g.fs_grid, block, g.fs_smem, st, po.pdl)
                                       : c2c_launch_fiber_stream_v1_s1(g.np2, sa, g.fs_nthr, g.fs_nslots, g.fs_merge, g.fs_wall, g.fs_grid, block, g.fs_smem, st, po.pdl);
        if (e != cudaSuccess) return cuda_fail(e, "fiber_stream launch");
        g_launches.fetch_add(1, std::memory_order_relaxed);
        if (po.main_only) return 0;
        nparts = (long long)g.fs_grid * (g.fs_merge ? 1 : g.fs_nslots);
        a.p0 = g.fs_nstage * g.fs_tp;
        legacy = a.p0 < g.P;
        if (legacy) {                                  // tail planes: one CTA row of the register-staged kernel, extra partials
            a.out = w.part + (size_t)nparts * g.E * a.ldr;
            grid = dim3(1, g.fc_gy);
            ppc = g.P - a.p0;
            nparts += g.fc_nslots;
        }
    }
    if (po.main_only) return fail(-60, "this shape / rank takes the register-staged fiber kernel");
    if (legacy) {
        if (masked) {
            const long long n = g.E * a.ldr;
            bmat_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(a.A2, a.A3, g.I2, g.I3, a.ldr, w.Bmat, a.done);
            CKL("bmat launch");
            a.Bmat = w.Bmat;
        }
        for (int c = 0; c < g.nchunk; ++c) {
            a.r0 = c * 2 * g.np2;
            if (a.r0 >= a.R) break;
            cudaError_t e;
            if (masked) {
                e = vec == 4 ? c2c_launch_fiber_v4_m1(g.np2, a, grid, g.fc_block, g.fc_nvec, g.fc_nvec_total, g.fc_nslots, ppc, st)
                  : vec == 2 ? c2c_launch_fiber_v2_m1(g.np2, a, grid, g.fc_block, g.fc_nvec, g.fc_nvec_total, g.fc_nslots, ppc, st)
                             : c2c_launch_fiber_v1_m1(g.np2, a, grid, g.fc_block, g.fc_nvec, g.fc_nvec_total, g.fc_nslots, ppc, st);
            } else {
                e = vec == 4 ? c2c_launch_fiber_v4_m0(g.np2, a, grid, g.fc_block, g.fc_nvec, g.fc_nvec_total, g.fc_nslots, ppc, st)
                  : vec == 2 ? c2c_launch_fiber_v2_m0(g.np2, a, grid, g.fc_block, g.fc_nvec, g.fc_nvec_total, g.fc_nslots, ppc, st)
                             : c2c_launch_fiber_v1_m0(g.np2, a, grid, g.fc_block, g.fc_nvec, g.fc_nvec_total, g.fc_nslots, ppc, st);
            }
            if (e != cudaSuccess) return cuda_fail(e, "fiber_contract launch");
            g_launches.fetch_add(1, std::memory_order_relaxed);
        }
    }
    const long long n4 = g.E * a.ldr / 4;
    if (po.xg != nullptr) {
        FiberXgArgs fx;
        fx.part = w.part; fx.nparts = (int)nparts; fx.n4 = n4; fx.R = a.R; fx.ldr = a.ldr; fx.N = po.xg_N; fx.nl = a.lead.nl;
        fx.U = w.U; fx.S = po.xg_S; fx.state = po.xg_state; fx.done = a.done;
        long long grid_x = (n4 + 31) / 32;
        if (grid_x > 2LL * sm_count()) grid_x = 2LL * sm_count();
        const cudaError_t e = c2c_launch_ex(fiber_reduce_xg_kernel, dim3((unsigned)grid_x), dim3(256), 0, st, legacy ? 0 : po.pdl, fx, *po.xg);
        if (e != cudaSuccess) return cuda_fail(e, "fiber_reduce_xg launch");
        g_launches.fetch_add(1, std::memory_order_relaxed);
        return 0;
    }
    {
        const cudaError_t e = c2c_launch_ex(fiber_reduce2_kernel, dim3((unsigned)((n4 + 31) / 32)), dim3(256), 0, st, legacy ? 0 : po.pdl,
                                            (const float*)w.part, (int)nparts, n4, a.R, a.ldr, w.U, a.done);
        if (e != cudaSuccess) return cuda_fail(e, "fiber_reduce launch");
        g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    return 0;
}

// Xh = mask ? X : reconstruction (used when the rank does not fit one chunk)
__global__ void __launch_bounds__(256)
impute_kernel(const float* __restrict__ X, const uint8_t* __restrict__ mask, long long P, int I2, int I3, LeadInfo lead,
              const float* __restrict__ A2, const float* __restrict__ A3, int R, int ldr, float* __restrict__ Xh,
              const int* done)
{
    if (run_done(done)) return;
    const long long E = (long long)I2 * I3, n = P * E;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float x = X[i];
        if (!mask[i]) {
            const long long p = i / E; const int e = (int)(i - p * E);
            const int s = e / I3, v = e - s * I3;
            float rec = 0.0f;
            for (int q = 0; q < R; ++q)
                rec = fmaf(lead_weight_col(p, lead, ldr, q, -1) * A2[(size_t)s * ldr + q], A3[(size_t)v * ldr + q], rec);
            x = rec;
        }
        Xh[i] = x;
    }
}

// recon_sums for ranks beyond one chunk: one thread per entry, full-rank reconstruction in a loop (the factors are
// L2-resident); per-CTA partial sums in the same layout as the plane_contract<SUMS> kernel.
__global__ void __launch_bounds__(256)
recon_sums_generic_kernel(const float* __restrict__ X, const uint8_t* __restrict__ mask, long long P, int I2, int I3,
                          LeadInfo lead, const float* __restrict__ A2, const float* __restrict__ A3, int R, int ldr,
                          double* __restrict__ sums, int missing)
{
    const long long E = (long long)I2 * I3, n = P * E;
    double v[5] = {0, 0, 0, 0, 0};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if (mask && ((mask[i] != 0) == (missing != 0))) continue;
        const long long p = i / E; const int e = (int)(i - p * E);
        const int s = e / I3, c = e - s * I3;
        float rec = 0.0f;
        for (int q = 0; q < R; ++q)
            rec = fmaf(lead_weight_col(p, lead, ldr, q, -1) * A2[(size_t)s * ldr + q], A3[(size_t)c * ldr + q], rec);
        const double x = missing ? 0.0 : (double)X[i], d = x - (double)rec;
        v[0] += d; v[1] += d * d; v[2] += x; v[3] += x * x; v[4] += 1.0;
    }
    __shared__ double red[5][8];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        const double t = warp_sum(v[k]);
        if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = t;
    }
    __syncthreads();
    if (threadIdx.x < 5) {
        double t = 0;
        for (int i = 0; i < 8; ++i) t += red[threadIdx.x][i];
        sums[(size_t)blockIdx.x * 5 + threadIdx.x] = t;
    }
}

static int maybe_impute(const Geo& g, ContractArgs& a, bool& masked, const Ws& w, cudaStream_t st)
{
    if (masked && g.nchunk > 1) {
        impute_kernel<<<sm_count() * 8, 256, 0, st>>>(a.X, a.mask, g.P, g.I2, g.I3, a.lead, a.A2, a.A3, a.R, a.ldr, w.Xh, a.done);
        CKL("impute launch");
        a.X = w.Xh; a.mask = nullptr; masked = false;
    }
    return 0;
}

static int check_common(const float* X, const float* const* A, int ndim, int ldr, int rank)
{
    if (!X) return fail(-10, "X is NULL");
    if (!A) return fail(-11, "A is NULL");
    for (int i = 0; i < ndim; ++i) if (!A[i]) return fail(-11, "A[%d] is NULL", i);
    if (ldr != ldr_of(rank)) return fail(-12, "ldr must be rank rounded up to a multiple of 4 (%d), got %d", ldr_of(rank), ldr);
    // every entry point dispatches 128-bit loads / bulk copies / TMA on X and float4 loads on the packed factors
    if (((uintptr_t)X & 15) != 0) return fail(-25, "X must be 16-byte aligned");
    for (int i = 0; i < ndim; ++i) if (((uintptr_t)A[i] & 15) != 0) return fail(-25, "A[%d] must be 16-byte aligned", i);
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// K1 / K1b
// ---------------------------------------------------------------------------------------------------------------------
extern "C" int c2c_build_ccc(const float* expr, const int64_t* expr_off, const int32_t* expr_ld, const int32_t* cell_col,
                             const int32_t* a_first, const int32_t* a_count, const int32_t* b_first, const int32_t* b_count,
                             const int32_t* sub_rows, int C, int L, int K, int score_kind, int agg_kind,
                             float* X, uint8_t* mask, void* stream)
{
    if (C < 0 || L < 0 || K < 0) return fail(-1, "negative dimension");
    if ((long long)C * L * K == 0) return 0;
    if (!expr || !expr_off || !expr_ld || !cell_col || !a_first || !a_count || !b_first || !b_count || !sub_rows || !X)
        return fail(-2, "NULL argument");
    if (score_kind < 0 || score_kind > 2) return fail(-3, "score_kind %d not in {product, mean, gmean}", score_kind);
    if (agg_kind < 0 || agg_kind > 2) return fail(-4, "agg_kind %d not in {min, mean, gmean}", agg_kind);
    if (((uintptr_t)X & 15) != 0 || (mask && ((uintptr_t)mask & 3) != 0)) return fail(-5, "X must be 16-byte and mask 4-byte aligned");
    BuildArgs a;
    a.expr = expr; a.expr_off = (const long long*)expr_off; a.expr_ld = expr_ld; a.cell_col = cell_col;
    a.first[0] = a_first; a.count[0] = a_count; a.first[1] = b_first; a.count[1] = b_count; a.sub_rows = sub_rows;
    a.C = C; a.L = L; a.K = K; a.score = score_kind; a.agg = agg_kind; a.X = X; a.mask = mask;
    const long long P = (long long)C * L;
    const long long ntiles = (P + C2C_BUILD_TILE - 1) / C2C_BUILD_TILE;
    const size_t smem = (size_t)C2C_BUILD_TILE * 2 * K * sizeof(double);
    if (smem > 48 * 1024) return fail(-6, "K = %d too large for the build tile", K);
    long long grid = ntiles; const long long cap = (long long)sm_count() * 8;
    if (grid > cap) grid = cap;
    // 'min' aggregates (and plain genes) are exact fp32 values: fp32 staging and scores.  'mean' / 'gmean' keep the fp64 value
    // the reference would carry into the score (only complexes differ, but the flag is per launch).
    a.exact = agg_kind == C2C_AGG_MIN ? 1 : 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (K % 4 == 0 && K <= 64 && a.exact && !env_int("C2C_B200_BUILD_GENERAL", 0)) {
        // the common case: a branch-free quad of scores per thread (c2c_small.cuh)
        const size_t smem_f = (size_t)C2C_BUILD_TILE * 2 * K * sizeof(float);
        if (score_kind == 0)      build_ccc_fast_kernel<0><<<(int)grid, C2C_SMALL_THREADS, smem_f, st>>>(a);
        else if (score_kind == 1) build_ccc_fast_kernel<1><<<(int)grid, C2C_SMALL_THREADS, smem_f, st>>>(a);
        else                      build_ccc_fast_kernel<2><<<(int)grid, C2C_SMALL_THREADS, smem_f, st>>>(a);
    }
    else if (K % 4 == 0) { if (a.exact) build_ccc_kernel<true, true><<<(int)grid, C2C_SMALL_THREADS, smem, st>>>(a);
                      else         build_ccc_kernel<true, false><<<(int)grid, C2C_SMALL_THREADS, smem, st>>>(a); }
    else            { if (a.exact) build_ccc_kernel<false, true><<<(int)grid, C2C_SMALL_THREADS, smem, st>>>(a);
                      else         build_ccc_kernel<false, false><<<(int)grid, C2C_SMALL_THREADS, smem, st>>>(a); }
    CKL("build_ccc launch");
    return 0;
}

extern "C" int c2c_group_aggregate(const float* X_in, const uint8_t* mask_in, const int32_t* grp_off, const int32_t* grp_rows,
                                   int C, int L, int n_groups, int64_t E, int method, float* X_out, uint8_t* mask_out,
                                   void* stream)
{
    if (C < 0 || L < 0 || n_groups < 0 || E < 0) return fail(-1, "negative dimension");
    if (method != C2C_AGG_GMEAN && method != C2C_AGG_SUM && method != C2C_AGG_MEAN) return fail(-3, "method %d not in {gmean, sum, mean}", method);
    const long long n = (long long)C * n_groups * E;
    if (n == 0) return 0;
    if (!X_in || !grp_off || !grp_rows || !X_out) return fail(-2, "NULL argument");
    long long grid = (n + 255) / 256; const long long cap = (long long)sm_count() * 16;
    if (grid > cap) grid = cap;
    group_aggregate_kernel<<<(int)grid, C2C_SMALL_THREADS, 0, (cudaStream_t)stream>>>(X_in, mask_in, grp_off, grp_rows, C, L, n_groups,
                                                                                       E, method, X_out, mask_out);
    CKL("group_aggregate launch");
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// ||X||^2
// ---------------------------------------------------------------------------------------------------------------------
static int sumsq_into(const float* X, long long n, double* out, double* part, cudaStream_t st)
{
    int blocks = C2C_SUMSQ_BLOCKS;
    const long long need = (n / 4 + 255) / 256;
    if (need < blocks) blocks = (int)(need < 1 ? 1 : need);
    sumsq_partial_kernel<<<blocks, C2C_SMALL_THREADS, 0, st>>>(X, n, part);
    CKL("sumsq launch");
    sum_partials_kernel<<<1, 32, 0, st>>>(part, blocks, 1, 1, out);
    CKL("sumsq reduce launch");
    return 0;
}

extern "C" int c2c_sumsq(const float* X, int64_t n, double* out, void* ws, size_t ws_bytes, void* stream)
{
    if (!X || !out) return fail(-1, "NULL argument");
    if (n < 0) return fail(-2, "negative length");
    if (((uintptr_t)X & 15) != 0) return fail(-3, "X must be 16-byte aligned");
    if (!ws || ws_bytes < C2C_SUMSQ_BLOCKS * sizeof(double)) return fail(-5, "workspace too small: need %zu bytes", C2C_SUMSQ_BLOCKS * sizeof(double));
    return sumsq_into(X, n, out, (double*)ws, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------------------------------
// factor-sized steps
// ---------------------------------------------------------------------------------------------------------------------
static int launch_gram_partial(const float* A, const float* M, long long I, int R, int ldr, const Ws& w, bool want_ip,
                               const int* done, cudaStream_t st)
{
    const int nblk = (int)((I + C2C_UPD_ROWS - 1) / C2C_UPD_ROWS);
    const size_t smem = (size_t)C2C_UPD_ROWS * R * sizeof(float);
    gram_partial_kernel<<<nblk, C2C_SMALL_THREADS, smem, st>>>(A, M, I, R, ldr, w.gram_part, want_ip ? w.iprod_part : nullptr, done);
    CKL("gram_partial launch");
    return 0;
}

static int launch_mu(float* A, const float* M, const float* G, long long I, int R, int ldr, float eps, const Ws& w,
                     bool want_gram, bool want_ip, const int* done, cudaStream_t st)
{
    const int nblk = (int)((I + C2C_UPD_ROWS - 1) / C2C_UPD_ROWS);
    const size_t smem = (size_t)2 * C2C_UPD_ROWS * R * sizeof(float);
    mu_update_kernel<<<nblk, C2C_SMALL_THREADS, smem, st>>>(A, M, G, I, R, ldr, eps, want_gram ? w.gram_part : nullptr,
                                                           want_ip ? w.iprod_part : nullptr, done);
    CKL("mu_update launch");
    return 0;
}

static int launch_hals(float* A, const float* M, const float* G, long long I, int R, int ldr, int max_sweeps, float tol,
                       const int* done, cudaStream_t st)
{
    int threads = (int)(I < 1024 ? (I + 31) / 32 * 32 : 1024);
    hals_update_kernel<<<1, threads, (size_t)R * R * sizeof(float), st>>>(A, M, G, I, R, ldr, max_sweeps, tol, done);
    CKL("hals_update launch");
    return 0;
}

static FinalizeArgs fin_args(const Geo& g, int R, int ldr, const Ws& w, const int* done)
{
    FinalizeArgs f;
    memset(&f, 0, sizeof(f));
    f.S = w.S; f.N = g.ndim; f.R = R; f.ldr = ldr; f.k = -1; f.next = -1; f.done = done;
    return f;
}

// M for mode `mode` given fresh T (leading) or U (trailing)
static int launch_mode_mttkrp(const Geo& g, const float* const* A, int mode, int R, int ldr, const Ws& w, float* M,
                              const int* done, cudaStream_t st)
{
    if (mode < g.ndim - 2) {
        LeadInfo li; fill_lead(li, g, A);
        lead_mttkrp_kernel<<<(int)g.shape[mode], C2C_SMALL_THREADS, 0, st>>>(w.T, li, mode, g.P, R, ldr, M, done);
        CKL("lead_mttkrp launch");
    } else {
        const int which = mode - (g.ndim - 2);
        const float* oth = which == 0 ? A[g.ndim - 1] : A[g.ndim - 2];
        const long long n = (long long)(which == 0 ? g.I2 : g.I3) * ldr;
        trail_mttkrp_kernel<<<(int)((n + 127) / 128), 128, 0, st>>>(w.U, oth, g.I2, g.I3, which, R, ldr, M, done);
        CKL("trail_mttkrp launch");
    }
    return 0;
}

extern "C" int c2c_mttkrp(const float* X, const uint8_t* mask, int ndim, const int64_t* shape, const float* const* A,
                          int ldr, int rank, int mode, float* M, void* ws, size_t ws_bytes, void* stream)
{
    Geo g; int rc;
    if (!shape) return fail(-1, "shape is NULL");
    if ((rc = make_geo(g, ndim, shape, rank)) != 0) return rc;
    if ((rc = check_common(X, A, ndim, ldr, rank)) != 0) return rc;
    if (mode < 0 || mode >= ndim) return fail(-13, "mode %d out of range", mode);
    if (!M) return fail(-14, "M is NULL");
    Ws w;
    if ((rc = get_ws(w, g, rank, mask != nullptr, ws, ws_bytes)) != 0) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    ContractArgs a = base_args(g, X, mask, A, ldr, rank, nullptr);
    bool masked = mask != nullptr;
    if ((rc = maybe_impute(g, a, masked, w, st)) != 0) return rc;
    if (mode < ndim - 2) { if ((rc = launch_plane(g, a, masked, w.T, st)) != 0) return rc; }
    else                 { if ((rc = launch_fiber(g, a, masked, w, st)) != 0) return rc; }
    return launch_mode_mttkrp(g, A, mode, rank, ldr, w, M, nullptr, st);
}

extern "C" int c2c_plane_contract(const float* X, const uint8_t* mask, int ndim, const int64_t* shape, const float* const* A,
                                  int ldr, int rank, float* T, void* ws, size_t ws_bytes, void* stream)
{
    Geo g; int rc;
    if (!shape) return fail(-1, "shape is NULL");
    if ((rc = make_geo(g, ndim, shape, rank)) != 0) return rc;
    if ((rc = check_common(X, A, ndim, ldr, rank)) != 0) return rc;
    if (!T) return fail(-14, "T is NULL");
    Ws w;
    if ((rc = get_ws(w, g, rank, mask != nullptr, ws, ws_bytes)) != 0) return rc;
    ContractArgs a = base_args(g, X, mask, A, ldr, rank, nullptr);
    bool masked = mask != nullptr;
    if ((rc = maybe_impute(g, a, masked, w, (cudaStream_t)stream)) != 0) return rc;
    return launch_plane(g, a, masked, T, (cudaStream_t)stream);
}

extern "C" int c2c_fiber_contract(const float* X, const uint8_t* mask, int ndim, const int64_t* shape, const float* const* A,
                                  int ldr, int rank, float* U, void* ws, size_t ws_bytes, void* stream)
{
    Geo g; int rc;
    if (!shape) return fail(-1, "shape is NULL");
    if ((rc = make_geo(g, ndim, shape, rank)) != 0) return rc;
    if ((rc = check_common(X, A, ndim, ldr, rank)) != 0) return rc;
    if (!U) return fail(-14, "U is NULL");
    Ws w;
    if ((rc = get_ws(w, g, rank, mask != nullptr, ws, ws_bytes)) != 0) return rc;
    ContractArgs a = base_args(g, X, mask, A, ldr, rank, nullptr);
    bool masked = mask != nullptr;
    if ((rc = maybe_impute(g, a, masked, w, (cudaStream_t)stream)) != 0) return rc;
    if ((rc = launch_fiber(g, a, masked, w, (cudaStream_t)stream)) != 0) return rc;
    CK(cudaMemcpyAsync(U, w.U, (size_t)g.E * ldr * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream), "copy U");
    return 0;
}

// Measurement aid (bench.py roofline): average duration of the streaming kernel of one pass ALONE -- plane_stream / plane_mma
// (which = 0) or fiber_stream / fiber_mma (which = 1) -- over `reps` back-to-back launches, CUDA events on the launching stream.
// The API calls above add a memset, a tail kernel, the partial reduction and a copy around it.  *kind: 0 FFMA kernel, 1 tensor pipe.
extern "C" int c2c_time_pass_kernel(const float* X, int ndim, const int64_t* shape, const float* const* A, int ldr, int rank,
                                    int which, int reps, float* ms_out, int* kind, void* ws, size_t ws_bytes, void* stream)
{
    Geo g; int rc;
    if (!shape) return fail(-1, "shape is NULL");
    if ((rc = make_geo(g, ndim, shape, rank)) != 0) return rc;
    if ((rc = check_common(X, A, ndim, ldr, rank)) != 0) return rc;
    if (!ms_out || reps < 1 || (which != 0 && which != 1)) return fail(-14, "ms_out is NULL, reps < 1 or which not in {0, 1}");
    Ws w;
    if ((rc = get_ws(w, g, rank, 0, ws, ws_bytes)) != 0) return rc;
    ContractArgs a = base_args(g, X, nullptr, A, ldr, rank, nullptr);
    PassOpts po = kPlainPass;
    po.main_only = true; po.t_zeroed = true;
    cudaStream_t st = (cudaStream_t)stream;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    CK(cudaEventCreate(&e0), "event create");
    if (cudaEventCreate(&e1) != cudaSuccess) { cudaEventDestroy(e0); return cuda_fail(cudaGetLastError(), "event create"); }
    rc = 0;
    for (int i = -1; i < reps && rc == 0; ++i) {                        // i = -1: warm-up launch
        if (i == 0) cudaEventRecord(e0, st);
        rc = which == 0 ? launch_plane(g, a, false, w.T, st, po) : launch_fiber(g, a, false, w, st, po);
    }
    float ms = 0.0f;
    if (rc == 0) {
        cudaEventRecord(e1, st);
        const cudaError_t e = cudaEventSynchronize(e1);
        if (e != cudaSuccess || cudaEventElapsedTime(&ms, e0, e1) != cudaSuccess) rc = cuda_fail(cudaGetLastError(), "time pass kernel");
    } else cudaStreamSynchronize(st);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (rc != 0) return rc;
    *ms_out = ms / (float)reps;
    if (kind) *kind = which == 0 ? (g.mma_ok ? 1 : 0) : (g.mma_fiber_ok ? 1 : 0);
    return 0;
}

// M for one mode from an already computed T (leading modes) or U (trailing modes): the second half of c2c_mttkrp, exposed
// so that a sharded factorization can all-reduce T-derived / U partial sums between the two halves.
extern "C" int c2c_mttkrp_from_contraction(const float* TU, int ndim, const int64_t* shape, const float* const* A, int ldr, int rank,
                                           int mode, float* M, void* stream)
{
    Geo g; int rc;
    if (!shape) return fail(-1, "shape is NULL");
    if ((rc = make_geo(g, ndim, shape, rank)) != 0) return rc;
    if ((rc = check_common(TU, A, ndim, ldr, rank)) != 0) return rc;
    if (mode < 0 || mode >= ndim) return fail(-13, "mode %d out of range", mode);
    if (!M) return fail(-14, "M is NULL");
    Ws w; memset(&w, 0, sizeof(w));
    w.T = const_cast<float*>(TU); w.U = const_cast<float*>(TU);
    return launch_mode_mttkrp(g, A, mode, rank, ldr, w, M, nullptr, (cudaStream_t)stream);
}

static int gram_hadamard_impl(const float* const* A, int ndim, const int64_t* shape, int ldr, int rank, int changed_mode,
                              int skip_mode, float* G, double* rec_norm2, void* ws, size_t ws_bytes, void* stream)
{
    Geo g; int rc;
    if (!shape) return fail(-1, "shape is NULL");
    if ((rc = make_geo(g, ndim, shape, rank)) != 0) return rc;
    if (!A) return fail(-11, "A is NULL");
    for (int i = 0; i < ndim; ++i) if (!A[i]) return fail(-11, "A[%d] is NULL", i);
    if (ldr != ldr_of(rank)) return fail(-12, "ldr must be %d", ldr_of(rank));
    if (skip_mode < -1 || skip_mode >= ndim) return fail(-13, "skip_mode %d out of range", skip_mode);
    if (changed_mode < -1 || changed_mode >= ndim) return fail(-13, "changed_mode %d out of range", changed_mode);
    Ws w;
    if ((rc = get_ws(w, g, rank, 0, ws, ws_bytes)) != 0) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int k0 = changed_mode < 0 ? 0 : changed_mode, k1 = changed_mode < 0 ? ndim : changed_mode + 1;
    for (int k = k0; k < k1; ++k) {
        if ((rc = launch_gram_partial(A[k], nullptr, g.shape[k], rank, ldr, w, false, nullptr, st)) != 0) return rc;
        FinalizeArgs f = fin_args(g, rank, ldr, w, nullptr);
        f.k = k; f.gram_part = w.gram_part; f.nblk = (int)((g.shape[k] + C2C_UPD_ROWS - 1) / C2C_UPD_ROWS);
        if (k == k1 - 1) { f.next = skip_mode < 0 ? ndim : skip_mode; f.G = G; f.rec_norm2 = rec_norm2; }
        gram_finalize_kernel<<<1, C2C_SMALL_THREADS, 0, st>>>(f);
        CKL("gram_finalize launch");
    }
    return 0;
}

extern "C" int c2c_gram_hadamard(const float* const* A, int ndim, const int64_t* shape, int ldr, int rank, int skip_mode,
                                 float* G, double* rec_norm2, void* ws, size_t ws_bytes, void* stream)
{
    return gram_hadamard_impl(A, ndim, shape, ldr, rank, -1, skip_mode, G, rec_norm2, ws, ws_bytes, stream);
}

// only the Gram of factor `changed_mode` is recomputed; the others are the ones an earlier call left in this workspace
extern "C" int c2c_gram_hadamard_update(const float* const* A, int ndim, const int64_t* shape, int ldr, int rank,
                                        int changed_mode, int skip_mode, float* G, double* rec_norm2, void* ws,
                                        size_t ws_bytes, void* stream)
{
    if (changed_mode < 0) return fail(-13, "changed_mode %d out of range", changed_mode);
    return gram_hadamard_impl(A, ndim, shape, ldr, rank, changed_mode, skip_mode, G, rec_norm2, ws, ws_bytes, stream);
}

extern "C" int c2c_mu_update(float* A_mode, const float* M, const float* G, int64_t I, int rank, int ldr, float eps,
                             void* ws, size_t ws_bytes, void* stream)
{
    (void)ws; (void)ws_bytes;
    if (!A_mode || !M || !G) return fail(-1, "NULL argument");
    if (I < 1 || rank < 1 || rank > C2C_MAX_RANK || ldr != ldr_of(rank)) return fail(-2, "bad I / rank / ldr");
    Ws w; memset(&w, 0, sizeof(w));
    return launch_mu(A_mode, M, G, I, rank, ldr, eps, w, false, false, nullptr, (cudaStream_t)stream);
}

extern "C" int c2c_mu_update_coupled(float* A1_mode, float* A2_mode, const float* M1, const float* M2, const float* G1,
                                     const float* G2, int64_t I, int rank, int ldr, float eps, double* iprod, void* stream)
{
    if (!A1_mode || !A2_mode || !M1 || !M2 || !G1 || !G2) return fail(-1, "NULL argument");
    if (I < 1 || rank < 1 || rank > C2C_MAX_RANK || ldr != ldr_of(rank)) return fail(-2, "bad I / rank / ldr");
    if (A1_mode == A2_mode) return fail(-2, "the two tensors keep separate copies of a shared factor");
    cudaStream_t st = (cudaStream_t)stream;
    if (iprod) CK(cudaMemsetAsync(iprod, 0, 2 * sizeof(double), st), "iprod memset");
    const int nblk = (int)((I + C2C_UPD_ROWS - 1) / C2C_UPD_ROWS);
    const size_t smem = (size_t)3 * C2C_UPD_ROWS * rank * sizeof(float);
    mu_update_coupled_kernel<<<nblk, C2C_SMALL_THREADS, smem, st>>>(A1_mode, A2_mode, M1, M2, G1, G2, I, rank, ldr, eps, iprod,
                                                                   nullptr, nullptr);
    CKL("mu_update_coupled launch");
    return 0;
}

extern "C" int c2c_hals_update(float* A_mode, const float* M, const float* G, int64_t I, int rank, int ldr, int max_sweeps,
                               float tol, void* stream)
{
    if (!A_mode || !M || !G) return fail(-1, "NULL argument");
    if (I < 1 || rank < 1 || rank > C2C_MAX_RANK || ldr != ldr_of(rank)) return fail(-2, "bad I / rank / ldr");
    return launch_hals(A_mode, M, G, I, rank, ldr, max_sweeps, tol, nullptr, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------------------------------
// K6 final sums
// ---------------------------------------------------------------------------------------------------------------------
// out[0..4] = the five sums over the observed entries; missing != 0: over the MISSING entries with x taken as 0, so that
// out[1] = ||(1 - mask) * reconstruction||^2
static int recon_sums_impl(const Geo& g, const Ws& w, const float* X, const uint8_t* mask, const float* const* A, int ldr, int rank,
                           int missing, double* out, cudaStream_t st)
{
    ContractArgs a = base_args(g, X, mask, A, ldr, rank, nullptr);
    a.sums = w.sums_part;
    a.sums_missing = missing;
    if (g.nchunk > 1) {
        recon_sums_generic_kernel<<<C2C_SUMS_BLOCKS, 256, 0, st>>>(X, mask, g.P, g.I2, g.I3, a.lead, a.A2, a.A3, rank, ldr, w.sums_part, missing);
        CKL("recon_sums (generic) launch");
        sum_partials_kernel<<<5, 32, 0, st>>>(w.sums_part, C2C_SUMS_BLOCKS, 5, 5, out);
        CKL("recon_sums reduce launch");
        return 0;
    }
    static const int stream_sums = env_int("C2C_B200_NO_STREAM_SUMS", 0) ? 0 : 1;
    if (!mask && stream_sums) {                       // one read of X at streaming speed (c2c_sums.cu) when the shape allows
        int sgrid = 0;
        const cudaError_t e = c2c_launch_sums_stream(a, sm_count(), st, &sgrid);
        if (e == cudaSuccess) {
            g_launches.fetch_add(1, std::memory_order_relaxed);
            sum_partials_kernel<<<5, 32, 0, st>>>(w.sums_part, sgrid, 5, 5, out);
            CKL("recon_sums reduce launch");
            return 0;
        }
        if (e != cudaErrorNotSupported) return cuda_fail(e, "recon_sums (stream) launch");
        (void)cudaGetLastError();
    }
    const int vec = g.vec_plane;
    const size_t smem = ((size_t)g.I2 * 4 * g.np2 + (size_t)g.I3 * 2 * g.np2) * sizeof(float);
    int grid = plane_grid(g, vec);
    if (grid > C2C_SUMS_BLOCKS) grid = C2C_SUMS_BLOCKS;
    cudaError_t e = vec == 4 ? c2c_launch_sums_v4(g.np2, a, grid, smem, st)
                  : vec == 2 ? c2c_launch_sums_v2(g.np2, a, grid, smem, st) : c2c_launch_sums_v1(g.np2, a, grid, smem, st);
    if (e != cudaSuccess) return cuda_fail(e, "recon_sums launch");
    g_launches.fetch_add(1, std::memory_order_relaxed);
    sum_partials_kernel<<<5, 32, 0, st>>>(w.sums_part, grid, 5, 5, out);
    CKL("recon_sums reduce launch");
    return 0;
}

extern "C" int c2c_recon_sums(const float* X, const uint8_t* mask, int ndim, const int64_t* shape, const float* const* A,
                              int ldr, int rank, double* out, void* ws, size_t ws_bytes, void* stream)
{
    Geo g; int rc;
    if (!shape) return fail(-1, "shape is NULL");
    if ((rc = make_geo(g, ndim, shape, rank)) != 0) return rc;
    if ((rc = check_common(X, A, ndim, ldr, rank)) != 0) return rc;
    if (!out) return fail(-14, "out is NULL");
    Ws w;
    if ((rc = get_ws(w, g, rank, mask != nullptr, ws, ws_bytes)) != 0) return rc;
    return recon_sums_impl(g, w, X, mask, A, ldr, rank, 0, out, (cudaStream_t)stream);
}

// out[0] = ||(1 - mask) * reconstruction||^2: the squared norm of what tensorly's imputation `tensor*mask + cp_to_tensor(cp,
// mask=1-mask)` writes into the tensor (coupled_factorization.py:342-343) -- the term by which the direct error of the coupled
// factorization (error_calc without an MTTKRP, :420-431) exceeds the Gram form on the norm of the tensor as passed in
extern "C" int c2c_missing_rec_sumsq(const float* X, const uint8_t* mask, int ndim, const int64_t* shape, const float* const* A,
                                     int ldr, int rank, double* out, void* ws, size_t ws_bytes, void* stream)
{
    Geo g; int rc;
    if (!shape) return fail(-1, "shape is NULL");
    if ((rc = make_geo(g, ndim, shape, rank)) != 0) return rc;
    if ((rc = check_common(X, A, ndim, ldr, rank)) != 0) return rc;
    if (!out || !mask) return fail(-14, "out / mask is NULL");
    Ws w;
    if ((rc = get_ws(w, g, rank, 1, ws, ws_bytes)) != 0) return rc;
    if ((rc = recon_sums_impl(g, w, X, mask, A, ldr, rank, 1, w.scratch_d + 8, (cudaStream_t)stream)) != 0) return rc;
    CK(cudaMemcpyAsync(out, w.scratch_d + 9, sizeof(double), cudaMemcpyDeviceToDevice, (cudaStream_t)stream), "copy sum");
    return 0;
}

extern "C" int c2c_cp_normalize(float* const* A, int ndim, const int64_t* shape, int ldr, int rank, float* weights, void* stream)
{
    if (!A || !shape || !weights) return fail(-1, "NULL argument");
    if (ndim < 1 || ndim > C2C_MAX_DIMS) return fail(-2, "bad ndim");
    if (rank < 1 || rank > C2C_MAX_RANK || ldr != ldr_of(rank)) return fail(-3, "bad rank / ldr");
    for (int k = 0; k < ndim; ++k) {
        if (!A[k]) return fail(-1, "A[%d] is NULL", k);
        cp_normalize_kernel<<<1, C2C_SMALL_THREADS, 0, (cudaStream_t)stream>>>(A[k], shape[k], rank, ldr, weights, k == 0 ? 1 : 0);
        CKL("cp_normalize launch");
    }
    return 0;
}

extern "C" int c2c_admm_update(float* A_mode, double* X64, double* dual, double* aux, const float* M, const float* G, int64_t I,
                               int rank, int ldr, int constraint, double param, int n_inner, double tol, double* out, void* stream)
{
    if (!A_mode || !X64 || !dual || !aux || !M || !G || !out) return fail(-1, "NULL argument");
    if (I < 1 || rank < 1 || rank > C2C_MAX_RANK || ldr != ldr_of(rank)) return fail(-2, "bad I / rank / ldr");
    if (constraint < 0 || constraint > 3) return fail(-3, "constraint must be 0 (none), 1 (non_negative), 2 (l1_reg) or 3 (l2_square_reg)");
    if (n_inner < 1) return fail(-4, "n_inner must be >= 1");
    const cudaError_t e = c2c_launch_admm_update(A_mode, X64, dual, aux, M, G, I, rank, ldr, constraint, param, n_inner, tol, out,
                                                 (cudaStream_t)stream);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (e != cudaSuccess) return cuda_fail(e, "admm_update launch");
    return 0;
}

extern "C" int c2c_corrindex_pairs(const double* F, int n, int64_t rows, int rank, const int64_t* seg_off, int nseg, int method,
                                   double tol, double* out, int* status, void* stream)
{
    if (!F || !seg_off || !out || !status) return fail(-1, "NULL argument");
    if (n < 1 || rows < 1) return fail(-2, "need at least one decomposition and one row");
    if (rank < 1 || rank > C2C_MAX_RANK) return fail(-3, "rank must be in 1..%d", C2C_MAX_RANK);
    if (nseg < 1 || method < 0 || method > 3) return fail(-4, "bad block count / method");
    int launches = 0;
    const cudaError_t e = c2c_launch_corrindex_pairs(F, n, rows, rank, (const long long*)seg_off, nseg, method, tol, out, status,
                                                     (cudaStream_t)stream, &launches);
    g_launches.fetch_add(launches, std::memory_order_relaxed);
    if (e != cudaSuccess) return cuda_fail(e, "corrindex_pairs launch");
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// whole factorization
// ---------------------------------------------------------------------------------------------------------------------
struct RunCtx {
    Geo g; Ws w; const float* X; const uint8_t* mask; float* const* A; int ldr, R, algo, n_iter_max; double tol; int cvg;
    const double* normX2; double* errors; int* state; cudaStream_t st; float eps;
    int pdl;                   // launch the fused unmasked iteration's kernels as programmatic dependents
    int nruns;                 // batched runs: nruns independent factorizations side by side, column c belongs to run run_of[c]
    unsigned char run_of[32];
    int* run_flags; int* run_iters;
    bool mma_zero;             // plane_mma / fiber_mma: T is kept cleared by the fiber pass
    bool corr;                 // masked run as unmasked dense passes + missing-entry corrections (c2c_corr.cu)
    bool plan;                 // ... with the corrections taken from a mask plan (c2c_plan.cu): Gram form + residual lists
    PlanView pv; long long n_resid; bool model_missing;
    const XgComm* xg;          // context-sharded run over peer memory (c2c_xgpu.cuh); null: one GPU
};

static int update_mode(const RunCtx& c, int mode, bool last)
{
    const Geo& g = c.g; int rc;
    const int* done = c.state + 1;
    const float* const* A = (const float* const*)c.A;
    if ((rc = launch_mode_mttkrp(g, A, mode, c.R, c.ldr, c.w, c.w.M, done, c.st)) != 0) return rc;
    if (c.algo == C2C_ALGO_MU) {
        if ((rc = launch_mu(c.A[mode], c.w.M, c.w.G, g.shape[mode], c.R, c.ldr, c.eps, c.w, true, last, done, c.st)) != 0) return rc;
    } else {
        if ((rc = launch_hals(c.A[mode], c.w.M, c.w.G, g.shape[mode], c.R, c.ldr, 100, 1e-8f, done, c.st)) != 0) return rc;
        if ((rc = launch_gram_partial(c.A[mode], c.w.M, g.shape[mode], c.R, c.ldr, c.w, last, done, c.st)) != 0) return rc;
    }
    FinalizeArgs f = fin_args(g, c.R, c.ldr, c.w, done);
    f.k = mode; f.gram_part = c.w.gram_part; f.nblk = (int)((g.shape[mode] + C2C_UPD_ROWS - 1) / C2C_UPD_ROWS);
    f.next = (mode + 1) % g.ndim; f.G = c.w.G;
    if (last) {
        f.iprod_part = c.w.iprod_part; f.nblk_ip = f.nblk; f.normX2 = c.normX2; f.errors = c.errors; f.state = c.state;
        f.n_iter_max = c.n_iter_max; f.tol = c.tol; f.cvg = c.cvg;
    }
    gram_finalize_kernel<<<1, C2C_SMALL_THREADS, 0, c.st>>>(f);
    CKL("gram_finalize launch");
    return 0;
}

// fused MU update of leading mode `mode` from T (c2c_update.cuh)
static int launch_lead_update(const RunCtx& c, int mode)
{
    const Geo& g = c.g;
    LeadUpdArgs u;
    memset(&u, 0, sizeof(u));
    u.T = c.w.T; fill_lead(u.lead, g, (const float* const*)c.A); u.k = mode; u.P = g.P; u.R = c.R; u.ldr = c.ldr; u.N = g.ndim;
    u.A = c.A[mode]; u.S = c.w.S; u.state = c.state; u.eps = c.eps; u.done = c.state + 1;
    memcpy(u.run_of, c.run_of, sizeof(u.run_of)); u.run_flags = c.nruns > 1 ? c.run_flags : nullptr;
    const long long I = g.shape[mode];
    const long long cnt = g.P / I;                                   // planes carrying one index of this mode
    const int block = cnt >= 1024 ? 1024 : 256;
    const int tpr = cnt >= 512 ? block : 32;                         // CTA per row / warp per row
    const int nrc = block / tpr;
    const bool xg = c.xg != nullptr && mode >= 1;                    // replicated mode of a context-sharded run
    // (the exchange inside lead_update_xg has every CTA wait for the peers' CTA of the same index: all of them resident)
    const long long ctas = (xg && block == 1024) ? sm_count() : 2LL * sm_count();
    long long rows = (I + ctas - 1) / ctas;
    rows = (rows + nrc - 1) / nrc * nrc;
    u.rows_per_cta = (int)rows;
    const long long grid = (I + rows - 1) / rows;
    const int njr = tpr / (c.ldr / 4);
    if (njr < 1) return fail(-30, "rank too large for the fused update");
    const size_t smem = (((size_t)(c.R * c.R + nrc * c.ldr) * sizeof(float) + 7) & ~(size_t)7) + (size_t)nrc * njr * c.ldr * sizeof(double);
    if (xg) {
        if (grid > C2C_XG_MAX_CTAS) return fail(-70, "sharded run: too many CTAs in an exchange");
        const cudaError_t e = c2c_launch_ex(lead_update_xg_kernel, dim3((unsigned)grid), dim3(block), smem, c.st, c.pdl, u, tpr, *c.xg);
        if (e != cudaSuccess) return cuda_fail(e, "lead_update_xg launch");
        g_launches.fetch_add(1, std::memory_order_relaxed);
        return 0;
    }
    {
        const cudaError_t e = c2c_launch_ex(lead_update_kernel, dim3((unsigned)grid), dim3(block), smem, c.st, c.pdl, u, tpr);
        if (e != cudaSuccess) return cuda_fail(e, "lead_update launch");
        g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    return 0;
}

static size_t trail_update_smem(const Geo& g, int R, int ldr, int* u_in_smem)
{
    const size_t Imax = g.I2 > g.I3 ? g.I2 : g.I3;
    size_t base = ((size_t)g.ndim * R * R + 32) * sizeof(double) + ((size_t)R * R + (size_t)(g.I2 + g.I3 + Imax) * R) * sizeof(float) + 16;
    const size_t ub = (size_t)g.E * ldr * sizeof(float);
    *u_in_smem = (base + ub <= 200 * 1024) ? 1 : 0;
    return base + (*u_in_smem ? ub : 0);
}

// fused MU update of the trailing modes [which_begin, which_end) from U, + error / stop rule after the last one
static int launch_trail_update(const RunCtx& c, int which_begin, int which_end)
{
    const Geo& g = c.g;
    TrailUpdArgs u;
    memset(&u, 0, sizeof(u));
    u.U = c.w.U; u.A2 = c.A[g.ndim - 2]; u.A3 = c.A[g.ndim - 1]; u.I2 = g.I2; u.I3 = g.I3; u.R = c.R; u.ldr = c.ldr; u.N = g.ndim;
    u.S = c.w.S; u.eps = c.eps; u.which_begin = which_begin; u.which_end = which_end;
    u.normX2 = c.normX2; u.errors = c.errors; u.state = c.state; u.n_iter_max = c.n_iter_max; u.tol = c.tol; u.cvg = c.cvg;
    u.done = c.state + 1;
    memcpy(u.run_of, c.run_of, sizeof(u.run_of)); u.nruns = c.nruns; u.run_flags = c.nruns > 1 ? c.run_flags : nullptr; u.run_iters = c.run_iters;
    const size_t smem = trail_update_smem(g, c.R, c.ldr, &u.u_in_smem);
    if (smem > 48 * 1024) {
        static thread_local size_t set_for = 0;
        if (set_for < smem) {
            CK(cudaFuncSetAttribute(trail_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024), "trail_update smem attribute");
            set_for = 200 * 1024;
        }
    }
    {
        const cudaError_t e = c2c_launch_ex(trail_update_kernel, dim3(1), dim3(C2C_TU_THREADS), smem, c.st, c.pdl, u);
        if (e != cudaSuccess) return cuda_fail(e, "trail_update launch");
        g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    return 0;
}

static bool fused_updates(const RunCtx& c)
{
    static int v = -1;
    if (v < 0) { const char* e = getenv("C2C_B200_NO_FUSED"); v = (e && e[0] == '1') ? 0 : 1; }
    if (!v || c.algo != C2C_ALGO_MU || c.g.nchunk != 1) return false;
    int dummy;
    const size_t Imax = c.g.I2 > c.g.I3 ? c.g.I2 : c.g.I3;
    (void)dummy;
    // the single-CTA trailing update keeps both trailing factors (+ one M) in shared memory, two new entries per thread
    return ((size_t)(c.g.I2 + c.g.I3 + Imax) * c.R * sizeof(float)) <= 96 * 1024 && Imax * c.R <= 2 * C2C_TU_THREADS &&
           c.g.P < (1LL << 31);
}

// Masked iteration, dense + correction form (c2c_corr.cu): ONE unmasked plane pass serves every leading mode and ONE unmasked
// fiber pass both trailing modes (the dense parts do not depend on the factors of their own side); before each mode's fused
// update the contribution of the re-imputed missing entries -- which does -- is added by a kernel that streams the mask only.
static int enqueue_iteration_corr(const RunCtx& c)
{
    const Geo& g = c.g; int rc;
    const int* done = c.state + 1;
    const float* const* A = (const float* const*)c.A;
    const int nlead = g.ndim - 2;
    const long long nB = g.E * c.ldr;
    auto bmat = [&]() -> int {
        bmat_kernel<<<(int)((nB + 255) / 256), 256, 0, c.st>>>(A[g.ndim - 2], A[g.ndim - 1], g.I2, g.I3, c.ldr, c.w.Bmat, done);
        CKL("bmat launch");
        g_launches.fetch_add(1, std::memory_order_relaxed);
        return 0;
    };
    RunCtx c2 = c;
    c2.w.T = c.w.Tw; c2.w.U = c.w.Uw;
    if ((rc = bmat()) != 0) return rc;
    {
        ContractArgs a = base_args(g, c.X, nullptr, A, c.ldr, c.R, done);
        if ((rc = launch_plane(g, a, false, c.w.T, c.st)) != 0) return rc;
    }
    for (int mode = 0; mode < nlead; ++mode) {
        LeadInfo li; fill_lead(li, g, A);
        const cudaError_t e = c2c_launch_plane_corr(c.mask, li, c.w.Bmat, c.w.T, c.w.Tw, g.P, g.E, c.R, c.ldr, done, sm_count(), c.st);
        if (e != cudaSuccess) return cuda_fail(e, "plane_corr launch");
        g_launches.fetch_add(1, std::memory_order_relaxed);
        if ((rc = launch_lead_update(c2, mode)) != 0) return rc;
    }
    {
        ContractArgs a = base_args(g, c.X, nullptr, A, c.ldr, c.R, done);
        if ((rc = launch_fiber(g, a, false, c.w, c.st)) != 0) return rc;
    }
    for (int j = 0; j < 2; ++j) {
        if (j == 1 && (rc = bmat()) != 0) return rc;                 // the first trailing factor has just changed
        LeadInfo li; fill_lead(li, g, A);
        int nparts = 2 * sm_count();
        if (nparts > c.w.cparts) nparts = c.w.cparts;
        const cudaError_t e = c2c_launch_fiber_corr(c.mask, li, c.w.Bmat, c.w.U, c.w.Uw, c.w.cpart, nparts, g.P, g.E, c.R, c.ldr,
                                                    done, c.st);
        if (e != cudaSuccess) return cuda_fail(e, "fiber_corr launch");
        g_launches.fetch_add(2, std::memory_order_relaxed);
        if ((rc = launch_trail_update(c2, j, j + 1)) != 0) return rc;
    }
    return 0;
}

// Masked iteration with a mask plan (c2c_plan.cu): the two unmasked streaming passes, plus -- before every mode's fused update
// -- the contraction of the reconstruction over the missing entries: restricted-Gram products for the separable part of the
// mask (no pass over the mask), list kernels for the residual entries.  Every kernel is a programmatic dependent of its
// predecessor.
static int enqueue_iteration_plan(const RunCtx& c)
{
    const Geo& g = c.g; int rc;
    const int* done = c.state + 1;
    const float* const* A = (const float* const*)c.A;
    const int nlead = g.ndim - 2, nsm = sm_count();
    const bool lists = c.n_resid > 0, gram = c.model_missing;
    PassOpts po = kPlainPass;
    po.pdl = c.pdl;
    RunCtx c2 = c;
    c2.w.T = (gram || lists) ? c.w.Tw : c.w.T;
    c2.w.U = (gram || lists) ? c.w.Uw : c.w.U;
    cudaError_t e;
    {
        ContractArgs a = base_args(g, c.X, nullptr, A, c.ldr, c.R, done);
        if ((rc = launch_plane(g, a, false, c.w.T, c.st, po)) != 0) return rc;
    }
    if (gram) {
        e = c2c_launch_trail_ctx_gram(c.pv, A[g.ndim - 2], A[g.ndim - 1], c.R, c.ldr, c.w.Dmat, done, c.st, c.pdl);
        if (e != cudaSuccess) return cuda_fail(e, "trail_ctx_gram launch");
        g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    if (lists) {
        const long long nB = g.E * c.ldr;
        e = c2c_launch_ex(bmat_kernel, dim3((unsigned)((nB + 255) / 256)), dim3(256), 0, c.st, 0, A[g.ndim - 2], A[g.ndim - 1], g.I2, g.I3,
                          c.ldr, c.w.Bmat, done);
        if (e != cudaSuccess) return cuda_fail(e, "bmat launch");
        g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    for (int mode = 0; mode < nlead; ++mode) {
        LeadInfo li; fill_lead(li, g, A);
        const float* src = c.w.T;
        if (gram) {
            e = c2c_launch_plane_corr_gram(c.pv, li, c.w.Dmat, c.w.T, c.w.Tw, c.R, c.ldr, done, nsm, c.st, c.pdl);
            if (e != cudaSuccess) return cuda_fail(e, "plane_corr_gram launch");
            g_launches.fetch_add(1, std::memory_order_relaxed);
            src = c.w.Tw;
        }
        if (lists) {
            // mode 0 finds the W rows of the previous iteration's fiber corrections (before the first iteration: of the
            // initial factors, ncp_run_impl): the leading factors have not changed since
            e = cudaSuccess;
            if (mode > 0) { e = c2c_launch_wmat(li, g.P, c.ldr, c.w.Wmat, done, c.st); g_launches.fetch_add(1, std::memory_order_relaxed); }
            if (e == cudaSuccess) e = c2c_launch_plane_corr_list(c.pv, c.w.Wmat, c.w.Bmat, src, c.w.Tw, c.ldr, done, nsm, c.st, c.pdl);
            if (e != cudaSuccess) return cuda_fail(e, "plane_corr_list launch");
            g_launches.fetch_add(1, std::memory_order_relaxed);
        }
        if ((rc = launch_lead_update(c2, mode)) != 0) return rc;
    }
    {
        ContractArgs a = base_args(g, c.X, nullptr, A, c.ldr, c.R, done);
        if ((rc = launch_fiber(g, a, false, c.w, c.st, po)) != 0) return rc;
    }
    if (gram) {
        LeadInfo li; fill_lead(li, g, A);
        e = c2c_launch_lead_ctx_gram(c.pv, li, c.R, c.ldr, c.w.Ypart, c.w.Ymat, done, nsm, c.st, c.pdl);
        if (e != cudaSuccess) return cuda_fail(e, "lead_ctx_gram launch");
        g_launches.fetch_add(2, std::memory_order_relaxed);
    }
    if (lists) {                                                       // the leading factors are final for this iteration
        LeadInfo li; fill_lead(li, g, A);
        e = c2c_launch_wmat(li, g.P, c.ldr, c.w.Wmat, done, c.st);
        if (e != cudaSuccess) return cuda_fail(e, "wmat launch");
        g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    for (int j = 0; j < 2; ++j) {
        const float* src = c.w.U;
        if (gram) {
            e = c2c_launch_fiber_corr_gram(c.pv, A[g.ndim - 2], A[g.ndim - 1], c.w.Ymat, c.w.U, c.w.Uw, c.R, c.ldr, done, c.st, c.pdl);
            if (e != cudaSuccess) return cuda_fail(e, "fiber_corr_gram launch");
            g_launches.fetch_add(1, std::memory_order_relaxed);
            src = c.w.Uw;
        }
        if (lists) {
            if (j == 1) {                                              // A_{N-2} has just been updated: B with it
                const long long nB = g.E * c.ldr;
                e = c2c_launch_ex(bmat_kernel, dim3((unsigned)((nB + 255) / 256)), dim3(256), 0, c.st, 0, A[g.ndim - 2], A[g.ndim - 1],
                                  g.I2, g.I3, c.ldr, c.w.Bmat, done);
                if (e != cudaSuccess) return cuda_fail(e, "bmat launch");
                g_launches.fetch_add(1, std::memory_order_relaxed);
            }
            e = c2c_launch_fiber_corr_list(c.pv, c.w.Wmat, c.w.Bmat, src, c.w.Uw, c.w.cpart, c.w.cparts, c.ldr, done, nsm, c.st, c.pdl);
            if (e != cudaSuccess) return cuda_fail(e, "fiber_corr_list launch");
            g_launches.fetch_add(2, std::memory_order_relaxed);
        }
        if ((rc = launch_trail_update(c2, j, j + 1)) != 0) return rc;
    }
    return 0;
}

static int enqueue_iteration(const RunCtx& c)
{
    if (c.plan) return enqueue_iteration_plan(c);
    if (c.corr) return enqueue_iteration_corr(c);
    const Geo& g = c.g; int rc;
    const int* done = c.state + 1;
    const float* const* A = (const float* const*)c.A;
    const int nlead = g.ndim - 2;
    const bool has_mask = c.mask != nullptr;
    for (int mode = 0; mode < g.ndim; ++mode) {
        const bool need_pass = has_mask || mode == 0 || mode == nlead;
        if (need_pass) {
            ContractArgs a = base_args(g, c.X, c.mask, A, c.ldr, c.R, done);
            bool masked = has_mask;
            if ((rc = maybe_impute(g, a, masked, c.w, c.st)) != 0) return rc;
            PassOpts po = kPlainPass;
            po.pdl = c.pdl;
            if (c.mma_zero) { po.t_zeroed = true; po.zero_T = c.w.T; po.zero_n = g.P * c.ldr; }
            if (c.xg != nullptr && mode == nlead) { po.xg = c.xg; po.xg_S = c.w.S; po.xg_state = c.state; po.xg_N = g.ndim; }
            if (mode < nlead) { if ((rc = launch_plane(g, a, masked, c.w.T, c.st, po)) != 0) return rc; }
            else              { if ((rc = launch_fiber(g, a, masked, c.w, c.st, po)) != 0) return rc; }
        }
        if (fused_updates(c)) {
            if (mode < nlead) { if ((rc = launch_lead_update(c, mode)) != 0) return rc; }
            else if (!has_mask) { if (mode == nlead) { if ((rc = launch_trail_update(c, 0, 2)) != 0) return rc; } }
            else { if ((rc = launch_trail_update(c, mode - nlead, mode - nlead + 1)) != 0) return rc; }
        } else if ((rc = update_mode(c, mode, mode == g.ndim - 1)) != 0) return rc;
    }
    return 0;
}

// a caller-built mask plan (c2c_mask_plan_scan / _fill)
struct PlanRef { const void* header; size_t header_bytes; const void* lists; size_t lists_bytes; const int64_t* info; };

static int plan_dims(const Geo& g, long long* P, int* E, int* I2, int* I3, int* C0)
{
    if (g.E > C2C_PLAN_MAX_E) return fail(-50, "mask plan: in-plane size %lld exceeds %d", g.E, C2C_PLAN_MAX_E);
    if (g.P >= (1LL << 31)) return fail(-50, "mask plan: too many planes");
    *P = g.P; *E = (int)g.E; *I2 = g.I2; *I3 = g.I3; *C0 = (int)g.shape[0];
    return 0;
}

static int ncp_run_impl(const float* X, const uint8_t* mask, int ndim, const int64_t* shape, float* const* A, int ldr,
                        int rank, int algo, int n_iter_max, double tol, int cvg_criterion, const double* normX2,
                        double* errors, int* state, int check_every, void* ws, size_t ws_bytes, void* stream,
                        const int* run_ranks, int nruns, const PlanRef* plan = nullptr, const XgComm* xg = nullptr)
{
    RunCtx c; int rc;
    c.xg = xg;
    const bool masked_ws = mask != nullptr;
    if (!shape) return fail(-1, "shape is NULL");
    if ((rc = make_geo(c.g, ndim, shape, rank)) != 0) return rc;
    if ((rc = check_common(X, (const float* const*)A, ndim, ldr, rank)) != 0) return rc;
    if (algo != C2C_ALGO_MU && algo != C2C_ALGO_HALS) return fail(-20, "algo %d not in {MU, HALS}", algo);
    if (algo == C2C_ALGO_HALS && mask) return fail(-21, "HALS does not take a mask (tensorly non_negative_parafac_hals has none)");
    if (cvg_criterion != C2C_CVG_ABS_REC_ERROR && cvg_criterion != C2C_CVG_REC_ERROR) return fail(-22, "unknown cvg_criterion %d", cvg_criterion);
    if (n_iter_max < 0) return fail(-23, "n_iter_max < 0");
    if (!errors || !state) return fail(-24, "errors / state is NULL");
    if (((uintptr_t)X & 15) != 0) return fail(-25, "X must be 16-byte aligned");
    if ((rc = get_ws(c.w, c.g, rank, masked_ws, ws, ws_bytes)) != 0) return rc;
    c.plan = false; c.n_resid = 0; c.model_missing = false;
    if (plan != nullptr) {
        if (!mask) return fail(-51, "mask plan given without a mask");
        if (!plan->header || !plan->info) return fail(-51, "mask plan: header / info is NULL");
        long long P; int E, I2, I3, C0;
        if ((rc = plan_dims(c.g, &P, &E, &I2, &I3, &C0)) != 0) return rc;
        if (plan->header_bytes < c2c_plan_header_bytes(P, E, I2, I3, C0)) return fail(-51, "mask plan: header buffer too small");
        if (plan->info[2] != 0) return fail(-50, "mask plan: the tensor holds non-zero values under zero mask bytes");
        if (plan->info[5] != 0) return fail(-50, "mask plan: too many residual entries");
        if (algo != C2C_ALGO_MU || c.g.nchunk != 1 || rank > 32) return fail(-50, "mask plan: multiplicative updates at rank <= 32 only");
        if (plan->info[1] > 0 && (!plan->lists || plan->lists_bytes < c2c_plan_lists_bytes(plan->info[1], plan->info[3])))
            return fail(-51, "mask plan: lists buffer too small");
        if (plan->info[0] == 0) mask = nullptr;                        // nothing is missing: the unmasked iteration
        else {
            c.plan = true; c.n_resid = plan->info[1]; c.model_missing = plan->info[4] != 0;
            c2c_plan_view(c.pv, P, E, I2, I3, C0, const_cast<void*>(plan->header), const_cast<void*>(plan->lists), c.n_resid);
        }
    }
    c.X = X; c.mask = mask; c.A = A; c.ldr = ldr; c.R = rank; c.algo = algo; c.n_iter_max = n_iter_max; c.tol = tol;
    c.cvg = cvg_criterion; c.errors = errors; c.state = state; c.st = (cudaStream_t)stream;
    c.eps = 1.1920928955078125e-07f;      // finfo(float32).eps: tensorly clips numerator and denominator at the dtype's eps
    cudaStream_t st = c.st;
    const Geo& g = c.g;

    c.nruns = nruns; c.run_flags = state + 2; c.run_iters = state + 2 + nruns;
    memset(c.run_of, 0, sizeof(c.run_of));
    if (nruns > 1) {
        int total = 0;
        for (int b = 0; b < nruns && nruns <= 32; ++b) {
            if (run_ranks[b] < 1) return fail(-26, "batched runs: every run needs rank >= 1");
            for (int q = 0; q < run_ranks[b]; ++q) { if (total < 32) c.run_of[total] = (unsigned char)b; ++total; }
        }
        if (nruns > 32 || total != rank) return fail(-26, "batched runs: the run ranks must add up to the total rank (<= 32), at most 32 runs");
        if (algo != C2C_ALGO_MU || mask != nullptr) return fail(-27, "batched runs: multiplicative updates without a mask only");
    }
    CK(cudaMemsetAsync(state, 0, (size_t)(nruns > 1 ? 2 + 2 * nruns : 2) * sizeof(int), st), "memset state");
    CK(cudaMemsetAsync(c.w.flags, 0, 4 * sizeof(int), st), "memset flags");
    // masked run as dense passes + missing-entry corrections: needs the fused update kernels, one rank chunk, the U correction
    // of a CTA in shared memory, and zeros at the missing entries of X (checked once here; cell2cell's tensors have them:
    // nan_to_num of the NaN the mask was derived from, tensor.py:842-847, 933-946)
    c.corr = false;
    if (c.plan && !fused_updates(c)) return fail(-50, "mask plan: shape outside the fused update kernels");
    if (!c.plan && mask != nullptr && algo == C2C_ALGO_MU && g.nchunk == 1 && rank <= 32 && fused_updates(c) &&
        c2c_fiber_corr_smem(g.E, ldr) <= C2C_CORR_SMEM_MAX && !env_int("C2C_B200_NO_MASK_CORR", 0)) {
        int host_flag = 1;
        const cudaError_t e = c2c_launch_masked_nonzero(X, mask, g.P * g.E, c.w.flags + 2, sm_count(), st);
        if (e != cudaSuccess) return cuda_fail(e, "masked_nonzero launch");
        CK(cudaMemcpyAsync(&host_flag, c.w.flags + 2, sizeof(int), cudaMemcpyDeviceToHost, st), "copy zero-fill flag");
        CK(cudaStreamSynchronize(st), "zero-fill check");
        c.corr = host_flag == 0;
    }
    if (normX2) c.normX2 = normX2;
    else {
        if ((rc = sumsq_into(X, g.P * g.E, c.w.normX2, c.w.sums_part, st)) != 0) return rc;
        c.normX2 = c.w.normX2;
    }
    // Grams of the initial factors (fused updates: into bank 1, bank 0 cleared for iteration 0); G for mode 0
    const bool fused = fused_updates(c);
    if (nruns > 1 && !fused) return fail(-28, "batched runs need the fused update kernels (rank <= 32, small trailing modes)");
    if (xg != nullptr && (!fused || mask != nullptr || nruns > 1 || ndim < 4 || !normX2))
        return fail(-71, "sharded run: unmasked multiplicative updates of a tensor with >= 2 leading modes at rank <= 32, "
                         "with the norm of the WHOLE tensor given");
    const size_t bank = (size_t)ndim * ldr * ldr;
    if (fused) CK(cudaMemsetAsync(c.w.S, 0, bank * sizeof(double), st), "memset Gram bank");
    for (int k = 0; k < ndim; ++k) {
        if ((rc = launch_gram_partial(A[k], nullptr, g.shape[k], rank, ldr, c.w, false, nullptr, st)) != 0) return rc;
        FinalizeArgs f = fin_args(g, rank, ldr, c.w, nullptr);
        if (fused) f.S = c.w.S + bank;
        f.k = k; f.gram_part = c.w.gram_part; f.nblk = (int)((g.shape[k] + C2C_UPD_ROWS - 1) / C2C_UPD_ROWS);
        if (k == ndim - 1) { f.next = 0; f.G = c.w.G; }
        gram_finalize_kernel<<<1, C2C_SMALL_THREADS, 0, st>>>(f);
        CKL("gram_finalize launch");
    }
    if (n_iter_max == 0) return 0;

    // The fused unmasked iteration is a chain of kernels only: they are launched as programmatic dependents (every kernel
    // waits for its predecessor before it touches that predecessor's results, so a streaming pass fills its ring while the
    // update kernel before it is still running).  On the tensor-pipe path T is cleared once here and then by every fiber pass.
    c.pdl = (use_pdl() && fused && (mask == nullptr || c.plan)) ? 1 : 0;
    c.mma_zero = g.mma_ok && g.mma_fiber_ok && mask == nullptr;
    if (c.mma_zero) CK(cudaMemsetAsync(c.w.T, 0, (size_t)g.P * ldr * sizeof(float), st), "memset T");

    if (c.plan && c.n_resid > 0) {                                     // W rows of the initial factors (enqueue_iteration_plan, mode 0)
        LeadInfo li; fill_lead(li, g, (const float* const*)c.A);
        const cudaError_t e = c2c_launch_wmat(li, g.P, ldr, c.w.Wmat, nullptr, st);
        if (e != cudaSuccess) return cuda_fail(e, "wmat launch");
        g_launches.fetch_add(1, std::memory_order_relaxed);
    }

    reap_pending_graphs(false);
    // `unroll` iterations are captured once and replayed: ~6 launches per iteration become a fraction of a graph launch, and
    // the dependent launches also link one iteration to the next
    cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr;
    bool use_graph = n_iter_max >= 3;
    int unroll = 1;
    if (use_graph && n_iter_max >= 10 && (check_every <= 0 || check_every >= 5))
        for (int u = 5; u >= 2; --u) if (n_iter_max % u == 0 && (check_every <= 0 || check_every % u == 0)) { unroll = u; break; }
    long long per_graph = 0;
    if (use_graph) {
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (st != nullptr && st != cudaStreamLegacy && st != cudaStreamPerThread &&
            (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone))
            use_graph = false;                                        // the caller is capturing: just enqueue
    }
    if (!use_graph) unroll = 1;
    if (use_graph) {
        // captured on a private stream (the caller's may be the legacy default stream, which cannot capture) and replayed
        // on the caller's stream
        cudaStream_t cap = nullptr;
        CK(cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking), "create capture stream");
        cudaError_t e = cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal);
        if (e != cudaSuccess) { cudaStreamDestroy(cap); return cuda_fail(e, "begin capture"); }
        const long long before = g_launches.load(std::memory_order_relaxed);
        RunCtx cc = c; cc.st = cap;
        rc = 0;
        for (int u = 0; u < unroll && rc == 0; ++u) rc = enqueue_iteration(cc);
        per_graph = g_launches.load(std::memory_order_relaxed) - before;
        g_launches.fetch_sub(per_graph, std::memory_order_relaxed);    // captured, not executed
        e = cudaStreamEndCapture(cap, &graph);
        cudaStreamDestroy(cap);
        if (rc != 0) { if (graph) cudaGraphDestroy(graph); return rc; }
        if (e != cudaSuccess) return cuda_fail(e, "end capture");
        e = cudaGraphInstantiate(&exec, graph, 0);
        if (e != cudaSuccess) { cudaGraphDestroy(graph); return cuda_fail(e, "graph instantiate"); }
    }
    int host_state[2] = {0, 0};
    rc = 0;
    for (int it = 0; it < n_iter_max; it += unroll) {
        if (use_graph) {
            cudaError_t e = cudaGraphLaunch(exec, st);
            if (e != cudaSuccess) { rc = cuda_fail(e, "graph launch"); break; }
            g_launches.fetch_add(per_graph, std::memory_order_relaxed);
        } else if ((rc = enqueue_iteration(c)) != 0) break;
        const int done_it = it + unroll;
        if (check_every > 0 && done_it % check_every == 0 && done_it < n_iter_max) {
            cudaError_t e = cudaMemcpyAsync(host_state, state, 2 * sizeof(int), cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) { rc = cuda_fail(e, "poll state"); break; }
            if (host_state[1]) break;
        }
    }
    if (use_graph) {
        // the exec object must outlive its in-flight launches: a run that was never polled parks it (see PendingGraph) and
        // returns without waiting for the device; a polled run has synchronised anyway
        bool parked = false;
        if (rc == 0 && check_every <= 0) {
            PendingGraph p; p.exec = exec; p.graph = graph; p.done = nullptr; p.device = -1;
            if (cudaGetDevice(&p.device) == cudaSuccess && cudaEventCreateWithFlags(&p.done, cudaEventDisableTiming) == cudaSuccess) {
                if (cudaEventRecord(p.done, st) == cudaSuccess) {
                    std::lock_guard<std::mutex> lock(g_pending_mu);
                    g_pending.push_back(p);
                    parked = true;
                } else cudaEventDestroy(p.done);
            }
            if (!parked) (void)cudaGetLastError();
        }
        if (!parked) {
            cudaError_t e = cudaStreamSynchronize(st);
            cudaGraphExecDestroy(exec); cudaGraphDestroy(graph);
            if (rc == 0 && e != cudaSuccess) rc = cuda_fail(e, "synchronize after graph replay");
        }
    }
    return rc;
}

extern "C" int c2c_ncp_run(const float* X, const uint8_t* mask, int ndim, const int64_t* shape, float* const* A, int ldr,
                           int rank, int algo, int n_iter_max, double tol, int cvg_criterion, const double* normX2,
                           double* errors, int* state, int check_every, void* ws, size_t ws_bytes, void* stream)
{
    return ncp_run_impl(X, mask, ndim, shape, A, ldr, rank, algo, n_iter_max, tol, cvg_criterion, normX2, errors, state,
                        check_every, ws, ws_bytes, stream, nullptr, 1);
}

extern "C" int c2c_ncp_run_batched(const float* X, int ndim, const int64_t* shape, float* const* A, int ldr, const int* run_ranks,
                                   int nruns, int n_iter_max, double tol, int cvg_criterion, const double* normX2,
                                   double* errors, int* state, int check_every, void* ws, size_t ws_bytes, void* stream)
{
    if (!run_ranks || nruns < 1) return fail(-26, "batched runs: run_ranks is NULL or nruns < 1");
    long long total = 0;
    for (int b = 0; b < nruns; ++b) total += run_ranks[b];
    if (total < 1 || total > 32) return fail(-26, "batched runs: the run ranks must add up to 1..32, got %lld", total);
    return ncp_run_impl(X, nullptr, ndim, shape, A, ldr, (int)total, C2C_ALGO_MU, n_iter_max, tol, cvg_criterion,
                        normX2, errors, state, check_every, ws, ws_bytes, stream, run_ranks, nruns);
}

// ---------------------------------------------------------------------------------------------------------------------
// one factorization sharded along the first mode over the GPUs of a box (c2c_xgpu.cuh)
// ---------------------------------------------------------------------------------------------------------------------
extern "C" size_t c2c_shard_comm_bytes(int ndim, const int64_t* shape, int rank, int world)
{
    if (!shape || ndim < 4 || ndim > C2C_MAX_DIMS || rank < 1 || rank > 32 || world < 1 || world > C2C_XG_MAX_WORLD) return 0;
    XgComm x;
    xg_layout(x, ndim - 2, shape, (long long)shape[ndim - 2] * shape[ndim - 1], ldr_of(rank), world);
    return (size_t)x.total;
}

extern "C" int c2c_enable_peer_access(int peer_device)
{
    int dev = -1;
    CK(cudaGetDevice(&dev), "cudaGetDevice");
    if (peer_device == dev) return 0;
    int can = 0;
    CK(cudaDeviceCanAccessPeer(&can, dev, peer_device), "cudaDeviceCanAccessPeer");
    if (!can) return fail(-72, "device %d cannot access the memory of device %d", dev, peer_device);
    const cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return 0; }
    if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceEnablePeerAccess");
    return 0;
}

// Communication buffers that the library allocates itself (cudaMalloc, zero-filled) so that they can be exported with
// cudaIpcGetMemHandle and opened by the peer processes with THEIR device current (cudaIpcMemLazyEnablePeerAccess maps the
// memory for peer access from the opening device, the way NCCL opens its own P2P buffers).
extern "C" int c2c_comm_alloc(size_t bytes, void** ptr_out, unsigned char* handle_out /*host, 64 bytes*/)
{
    if (!ptr_out || !handle_out || bytes == 0) return fail(-74, "c2c_comm_alloc: bad arguments");
    void* p = nullptr;
    CK(cudaMalloc(&p, bytes), "cudaMalloc (communication buffer)");
    cudaError_t e = cudaMemset(p, 0, bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); return cuda_fail(e, "cudaIpcGetMemHandle"); }
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handle_out, &h, sizeof(h));
    *ptr_out = p;
    return 0;
}

extern "C" int c2c_comm_open(const unsigned char* handle /*host, 64 bytes*/, void** ptr_out)
{
    if (!handle || !ptr_out) return fail(-74, "c2c_comm_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void* p = nullptr;
    CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
    *ptr_out = p;
    return 0;
}

extern "C" int c2c_comm_close(void* ptr)
{
    if (ptr) CK(cudaIpcCloseMemHandle(ptr), "cudaIpcCloseMemHandle");
    return 0;
}

extern "C" int c2c_comm_free(void* ptr)
{
    if (ptr) CK(cudaFree(ptr), "cudaFree (communication buffer)");
    return 0;
}

// One exchange over the buffers (a store into every peer, a flag, a wait, a sum): checks the mapping before a run relies on it.
// out[0] = sum over the ranks of (rank + 1) * value as seen in the local buffer.
__global__ void xg_probe_kernel(XgComm x, unsigned int epoch, float value, float* out)
{
    float* mine = reinterpret_cast<float*>(x.peer[x.me] + x.off_u);
    if ((int)threadIdx.x < x.world) reinterpret_cast<float*>(x.peer[threadIdx.x] + x.off_u)[x.me] = (x.me + 1) * value;
    xg_signal(x, x.off_flags_u, 0, epoch);
    xg_wait(x, x.off_flags_u, 0, -1, epoch);
    if (threadIdx.x == 0) {
        float s = 0.0f;
        for (int g = 0; g < x.world; ++g) s += __ldcg(mine + g);
        out[0] = s;
    }
}

extern "C" int c2c_comm_probe(int world, int me, void* const* comm, size_t comm_bytes, unsigned int epoch, float value, float* out,
                              void* stream)
{
    if (world < 1 || world > C2C_XG_MAX_WORLD || me < 0 || me >= world || !comm || !out) return fail(-73, "c2c_comm_probe: bad arguments");
    XgComm x;
    memset(&x, 0, sizeof(x));
    const int64_t shape[4] = {1, 1, 8, 8};
    xg_layout(x, 2, shape, 64, 4, world);
    if (comm_bytes < (size_t)x.total) return fail(-73, "c2c_comm_probe: buffer too small (%lld bytes needed)", x.total);
    x.me = me;
    for (int g = 0; g < world; ++g) x.peer[g] = (char*)comm[g];
    xg_probe_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(x, epoch, value, out);
    CKL("xg_probe launch");
    return 0;
}

extern "C" int c2c_ncp_run_sharded(const float* X_local, int ndim, const int64_t* local_shape, float* const* A, int ldr, int rank,
                                   int n_iter_max, double tol, int cvg_criterion, const double* normX2, double* errors, int* state,
                                   int check_every, void* ws, size_t ws_bytes, int world, int me, void* const* comm, size_t comm_bytes,
                                   unsigned int epoch_base, void* stream)
{
    if (!local_shape) return fail(-1, "shape is NULL");
    if (world < 1 || world > C2C_XG_MAX_WORLD || me < 0 || me >= world) return fail(-73, "sharded run: world %d / rank %d out of range (<= %d GPUs)", world, me, C2C_XG_MAX_WORLD);
    if (ndim < 4 || ndim > C2C_MAX_DIMS) return fail(-71, "sharded run: the tensor needs at least two leading modes (ndim >= 4)");
    if (!comm) return fail(-73, "sharded run: comm is NULL");
    XgComm x;
    memset(&x, 0, sizeof(x));
    xg_layout(x, ndim - 2, local_shape, (long long)local_shape[ndim - 2] * local_shape[ndim - 1], ldr, world);
    if (comm_bytes < (size_t)x.total) return fail(-73, "sharded run: communication buffer too small: need %lld bytes, got %zu", x.total, comm_bytes);
    x.me = me; x.epoch_base = epoch_base;
    for (int g = 0; g < world; ++g) {
        if (!comm[g]) return fail(-73, "sharded run: comm[%d] is NULL", g);
        if (((uintptr_t)comm[g] & 255) != 0) return fail(-73, "sharded run: communication buffers must be 256-byte aligned");
        x.peer[g] = (char*)comm[g];
    }
    return ncp_run_impl(X_local, nullptr, ndim, local_shape, A, ldr, rank, C2C_ALGO_MU, n_iter_max, tol, cvg_criterion, normX2, errors,
                        state, check_every, ws, ws_bytes, stream, nullptr, 1, nullptr, &x);
}

// ---------------------------------------------------------------------------------------------------------------------
// small tensors: a grid of independent factorizations, one CTA each (c2c_multi.cu)
// ---------------------------------------------------------------------------------------------------------------------
static int multi_check(Geo& g, int ndim, const int64_t* shape, const int* ranks, int nruns, size_t* smem_max, size_t* t_floats)
{
    int rc;
    if (!shape) return fail(-1, "shape is NULL");
    if (!ranks || nruns < 1) return fail(-60, "multi-run: ranks is NULL or nruns < 1");
    if ((rc = make_geo(g, ndim, shape, 1)) != 0) return rc;
    if (g.I2 > 128 || g.I3 > 128) return fail(-61, "multi-run: trailing modes larger than 128 (%d x %d)", g.I2, g.I3);
    if (g.P >= (1LL << 31)) return fail(-61, "multi-run: too many planes");
    *smem_max = 0; *t_floats = 0;
    for (int b = 0; b < nruns; ++b) {
        if (ranks[b] < 1 || ranks[b] > 32) return fail(-61, "multi-run: rank %d of run %d not in 1..32", ranks[b], b);
        const int ldr = ldr_of(ranks[b]);
        const size_t sm = c2c_multi_smem_bytes(ndim, ldr, g.E, g.I2, g.I3);
        if (sm > 200 * 1024) return fail(-61, "multi-run: %zu B of shared memory per run (> 200 KiB): tensor too large for this path", sm);
        if (sm > *smem_max) *smem_max = sm;
        *t_floats += align_up((size_t)g.P * ldr, 64);
    }
    return 0;
}

extern "C" size_t c2c_ncp_multi_workspace_bytes(int ndim, const int64_t* shape, const int* ranks, int nruns)
{
    Geo g; size_t smem, tf;
    if (multi_check(g, ndim, shape, ranks, nruns, &smem, &tf) != 0) return 0;
    return align_up((size_t)nruns * sizeof(MultiRun), 256) + tf * sizeof(float) + 256;
}

extern "C" int c2c_ncp_run_multi(const float* X, int ndim, const int64_t* shape, float* const* A, const int* ranks, int nruns,
                                 int n_iter_max, double tol, int cvg_criterion, const double* normX2, double* errors, int* state,
                                 double* sums, void* ws, size_t ws_bytes, void* stream)
{
    Geo g; size_t smem, tf; int rc;
    if ((rc = multi_check(g, ndim, shape, ranks, nruns, &smem, &tf)) != 0) return rc;
    if (!X || !A || !normX2 || !errors || !state) return fail(-10, "NULL argument");
    if (((uintptr_t)X & 15) != 0) return fail(-25, "X must be 16-byte aligned");
    if (cvg_criterion != C2C_CVG_ABS_REC_ERROR && cvg_criterion != C2C_CVG_REC_ERROR) return fail(-22, "unknown cvg_criterion %d", cvg_criterion);
    if (n_iter_max < 1) return fail(-23, "n_iter_max < 1");
    const size_t need = align_up((size_t)nruns * sizeof(MultiRun), 256) + tf * sizeof(float) + 256;
    if (!ws || ws_bytes < need) return fail(-5, "workspace too small: need %zu bytes, got %zu (query c2c_ncp_multi_workspace_bytes)", need, ws_bytes);
    if (((uintptr_t)ws & 255) != 0) return fail(-6, "workspace must be 256-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    std::vector<MultiRun> runs((size_t)nruns);
    float* tbase = (float*)((char*)ws + align_up((size_t)nruns * sizeof(MultiRun), 256));
    size_t toff = 0;
    for (int b = 0; b < nruns; ++b) {
        MultiRun& r = runs[(size_t)b];
        memset(&r, 0, sizeof(r));
        for (int k = 0; k < ndim; ++k) {
            r.A[k] = A[(size_t)b * ndim + k];
            if (!r.A[k] || ((uintptr_t)r.A[k] & 15) != 0) return fail(-11, "multi-run: factor %d of run %d is NULL or not 16-byte aligned", k, b);
        }
        r.rank = ranks[b]; r.ldr = ldr_of(ranks[b]);
        r.T = tbase + toff; toff += align_up((size_t)g.P * r.ldr, 64);
        r.errors = errors + (size_t)b * n_iter_max;
        r.state = state + 2 * (size_t)b;
        r.sums = sums ? sums + 5 * (size_t)b : nullptr;
    }
    CK(cudaMemcpyAsync(ws, runs.data(), (size_t)nruns * sizeof(MultiRun), cudaMemcpyHostToDevice, st), "copy run table");
    CK(cudaMemsetAsync(state, 0, (size_t)nruns * 2 * sizeof(int), st), "memset state");
    MultiArgs a;
    memset(&a, 0, sizeof(a));
    a.X = X; a.ndim = ndim;
    for (int k = 0; k < ndim; ++k) a.shape[k] = (int)shape[k];
    a.P = g.P; a.E = (int)g.E; a.I2 = g.I2; a.I3 = g.I3; a.runs = (const MultiRun*)ws; a.nruns = nruns;
    a.n_iter_max = n_iter_max; a.tol = tol; a.cvg = cvg_criterion; a.normX2 = normX2;
    a.eps = 1.1920928955078125e-07f;
    const cudaError_t e = c2c_launch_ncp_multi(a, smem, st);
    if (e != cudaSuccess) return cuda_fail(e, "ncp_multi launch");
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// mask plan (c2c_plan.cu): built once per (tensor, mask), used by every masked factorization of that tensor
// ---------------------------------------------------------------------------------------------------------------------
extern "C" size_t c2c_mask_plan_header_bytes(int ndim, const int64_t* shape)
{
    Geo g;
    if (!shape || make_geo(g, ndim, shape, 1) != 0) return 0;
    long long P; int E, I2, I3, C0;
    if (plan_dims(g, &P, &E, &I2, &I3, &C0) != 0) return 0;
    return c2c_plan_header_bytes(P, E, I2, I3, C0);
}

extern "C" int c2c_mask_plan_scan(const float* X, const uint8_t* mask, int ndim, const int64_t* shape, void* header,
                                  size_t header_bytes, int64_t* info, void* stream)
{
    Geo g; int rc;
    if (!shape) return fail(-1, "shape is NULL");
    if ((rc = make_geo(g, ndim, shape, 1)) != 0) return rc;
    if (!mask || !header || !info) return fail(-51, "mask / header / info is NULL");
    long long P; int E, I2, I3, C0;
    if ((rc = plan_dims(g, &P, &E, &I2, &I3, &C0)) != 0) return rc;
    if (header_bytes < c2c_plan_header_bytes(P, E, I2, I3, C0)) return fail(-51, "mask plan: header buffer too small");
    if (((uintptr_t)header & 255) != 0) return fail(-6, "mask plan: header must be 256-byte aligned");
    PlanView v;
    c2c_plan_view(v, P, E, I2, I3, C0, header, nullptr, 0);
    cudaStream_t st = (cudaStream_t)stream;
    const cudaError_t e = c2c_plan_scan(X, mask, v, sm_count(), st);
    if (e != cudaSuccess) return cuda_fail(e, "mask plan scan");
    g_launches.fetch_add(4, std::memory_order_relaxed);
    unsigned long long cnt[C2C_PLAN_NCOUNTERS];
    CK(cudaMemcpyAsync(cnt, v.counters, sizeof(cnt), cudaMemcpyDeviceToHost, st), "copy plan counters");
    CK(cudaStreamSynchronize(st), "mask plan scan");
    for (int i = 0; i < 6; ++i) info[i] = (int64_t)cnt[i];
    info[6] = (int64_t)c2c_plan_lists_bytes((long long)cnt[1], (long long)cnt[3]);
    info[7] = (int64_t)cnt[6];                                         // residual entries (info[1] = list slots incl. padding)
    return 0;
}

extern "C" int c2c_mask_plan_fill(const uint8_t* mask, int ndim, const int64_t* shape, const void* header, size_t header_bytes,
                                  const int64_t* info, void* lists, size_t lists_bytes, void* stream)
{
    Geo g; int rc;
    if (!shape) return fail(-1, "shape is NULL");
    if ((rc = make_geo(g, ndim, shape, 1)) != 0) return rc;
    if (!mask || !header || !info) return fail(-51, "mask / header / info is NULL");
    long long P; int E, I2, I3, C0;
    if ((rc = plan_dims(g, &P, &E, &I2, &I3, &C0)) != 0) return rc;
    if (header_bytes < c2c_plan_header_bytes(P, E, I2, I3, C0)) return fail(-51, "mask plan: header buffer too small");
    if (info[5] != 0) return fail(-50, "mask plan: too many residual entries");
    if (info[1] == 0) return 0;                                        // no residual entries: nothing to list
    if (!lists || lists_bytes < c2c_plan_lists_bytes(info[1], info[3])) return fail(-51, "mask plan: lists buffer too small");
    if (((uintptr_t)lists & 255) != 0) return fail(-6, "mask plan: lists must be 256-byte aligned");
    PlanView v;
    c2c_plan_view(v, P, E, I2, I3, C0, const_cast<void*>(header), lists, info[1]);
    const cudaError_t e = c2c_plan_fill(mask, v, sm_count(), (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "mask plan fill");
    g_launches.fetch_add(2, std::memory_order_relaxed);
    return 0;
}

extern "C" int c2c_ncp_run_planned(const float* X, const uint8_t* mask, int ndim, const int64_t* shape, float* const* A, int ldr,
                                   int rank, int n_iter_max, double tol, int cvg_criterion, const double* normX2, double* errors,
                                   int* state, int check_every, void* ws, size_t ws_bytes, const void* header, size_t header_bytes,
                                   const void* lists, size_t lists_bytes, const int64_t* info, void* stream)
{
    PlanRef pr = {header, header_bytes, lists, lists_bytes, info};
    return ncp_run_impl(X, mask, ndim, shape, A, ldr, rank, C2C_ALGO_MU, n_iter_max, tol, cvg_criterion, normX2, errors, state,
                        check_every, ws, ws_bytes, stream, nullptr, 1, &pr);
}

// T (which = 0) or U (which = 1) of the IMPUTED tensor -- X at the observed entries, the reconstruction of the given factors at
// the missing ones -- as the unmasked pass over X plus the plan's corrections
extern "C" int c2c_masked_contract_planned(const float* X, int ndim, const int64_t* shape, const float* const* A, int ldr, int rank,
                                           int which, float* out, void* ws, size_t ws_bytes, const void* header, size_t header_bytes,
                                           const void* lists, size_t lists_bytes, const int64_t* info, void* stream)
{
    Geo g; int rc;
    if (!shape) return fail(-1, "shape is NULL");
    if ((rc = make_geo(g, ndim, shape, rank)) != 0) return rc;
    if ((rc = check_common(X, A, ndim, ldr, rank)) != 0) return rc;
    if (!out || !header || !info) return fail(-14, "out / header / info is NULL");
    if (which != 0 && which != 1) return fail(-13, "which must be 0 (plane) or 1 (fiber)");
    if (g.nchunk != 1 || rank > 32) return fail(-50, "mask plan: rank <= 32 only");
    long long P; int E, I2, I3, C0;
    if ((rc = plan_dims(g, &P, &E, &I2, &I3, &C0)) != 0) return rc;
    if (header_bytes < c2c_plan_header_bytes(P, E, I2, I3, C0)) return fail(-51, "mask plan: header buffer too small");
    if (info[2] != 0 || info[5] != 0) return fail(-50, "mask plan not usable for this tensor");
    if (info[1] > 0 && (!lists || lists_bytes < c2c_plan_lists_bytes(info[1], info[3]))) return fail(-51, "mask plan: lists buffer too small");
    Ws w;
    if ((rc = get_ws(w, g, rank, 1, ws, ws_bytes)) != 0) return rc;
    PlanView v;
    c2c_plan_view(v, P, E, I2, I3, C0, const_cast<void*>(header), const_cast<void*>(lists), info[1]);
    cudaStream_t st = (cudaStream_t)stream;
    const int nsm = sm_count();
    const bool gram = info[4] != 0, lst = info[1] > 0;
    ContractArgs a = base_args(g, X, nullptr, A, ldr, rank, nullptr);
    LeadInfo li; fill_lead(li, g, A);
    cudaError_t e;
    if (which == 0) {
        if ((rc = launch_plane(g, a, false, w.T, st)) != 0) return rc;
        const float* src = w.T;
        if (gram) {
            e = c2c_launch_trail_ctx_gram(v, A[ndim - 2], A[ndim - 1], rank, ldr, w.Dmat, nullptr, st, 0);
            if (e == cudaSuccess) e = c2c_launch_plane_corr_gram(v, li, w.Dmat, w.T, w.Tw, rank, ldr, nullptr, nsm, st, 0);
            if (e != cudaSuccess) return cuda_fail(e, "plane_corr_gram launch");
            g_launches.fetch_add(2, std::memory_order_relaxed);
            src = w.Tw;
        }
        if (lst) {
            const long long nB = g.E * ldr;
            bmat_kernel<<<(int)((nB + 255) / 256), 256, 0, st>>>(A[ndim - 2], A[ndim - 1], g.I2, g.I3, ldr, w.Bmat, nullptr);
            CKL("bmat launch");
            e = c2c_launch_wmat(li, g.P, ldr, w.Wmat, nullptr, st);
            if (e == cudaSuccess) e = c2c_launch_plane_corr_list(v, w.Wmat, w.Bmat, src, w.Tw, ldr, nullptr, nsm, st, 0);
            if (e != cudaSuccess) return cuda_fail(e, "plane_corr_list launch");
            g_launches.fetch_add(2, std::memory_order_relaxed);
            src = w.Tw;
        }
        CK(cudaMemcpyAsync(out, src, (size_t)g.P * ldr * sizeof(float), cudaMemcpyDeviceToDevice, st), "copy T");
    } else {
        if ((rc = launch_fiber(g, a, false, w, st)) != 0) return rc;
        const float* src = w.U;
        if (gram) {
            e = c2c_launch_lead_ctx_gram(v, li, rank, ldr, w.Ypart, w.Ymat, nullptr, nsm, st, 0);
            if (e == cudaSuccess) e = c2c_launch_fiber_corr_gram(v, A[ndim - 2], A[ndim - 1], w.Ymat, w.U, w.Uw, rank, ldr, nullptr, st, 0);
            if (e != cudaSuccess) return cuda_fail(e, "fiber_corr_gram launch");
            g_launches.fetch_add(3, std::memory_order_relaxed);
            src = w.Uw;
        }
        if (lst) {
            const long long nB = g.E * ldr;
            bmat_kernel<<<(int)((nB + 255) / 256), 256, 0, st>>>(A[ndim - 2], A[ndim - 1], g.I2, g.I3, ldr, w.Bmat, nullptr);
            CKL("bmat launch");
            e = c2c_launch_wmat(li, g.P, ldr, w.Wmat, nullptr, st);
            if (e == cudaSuccess) e = c2c_launch_fiber_corr_list(v, w.Wmat, w.Bmat, src, w.Uw, w.cpart, w.cparts, ldr, nullptr, nsm, st, 0);
            if (e != cudaSuccess) return cuda_fail(e, "fiber_corr_list launch");
            g_launches.fetch_add(4, std::memory_order_relaxed);
            src = w.Uw;
        }
        CK(cudaMemcpyAsync(out, src, (size_t)g.E * ldr * sizeof(float), cudaMemcpyDeviceToDevice, st), "copy U");
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------------
// coupled factorization of two tensors sharing modes (cell2cell/tensor/coupled_factorization.py:123-482)
// ---------------------------------------------------------------------------------------------------------------------
struct CoupledSide {
    Geo g; Ws w; const float* X; const uint8_t* mask; float* const* A;
    bool T_ok, U_ok;           // is the cached plane / fiber contraction still the one of the current factors?
    int seq[C2C_MAX_NDIM]; int nseq;   // this tensor's modes in the order they are updated within an iteration
};

struct CoupledCtx {
    CoupledSide s[2];
    int ldr, R, n_iter_max, cvg; double tol, w1, w2; float eps;
    int kind[2 * C2C_MAX_NDIM], ma[2 * C2C_MAX_NDIM], mb[2 * C2C_MAX_NDIM], nsteps;   // kind 0: tensor 1 only, 1: tensor 2 only, 2: shared
    const double* nx2[2]; double* iprod; double* errors; int* state; cudaStream_t st;
    // tensor 2's chain (passes, MTTKRP tails, Gram finalizes) runs on a stream of its own between the shared updates, so the
    // small kernels of one tensor hide behind the streaming pass of the other; ev12 / ev21 order the two streams
    cudaStream_t st2; cudaEvent_t ev12, ev21;
};

// `to` waits for everything enqueued on `from` so far
static int coupled_order(cudaStream_t from, cudaStream_t to, cudaEvent_t ev)
{
    if (from == to) return 0;
    CK(cudaEventRecord(ev, from), "event record");
    CK(cudaStreamWaitEvent(to, ev, 0), "stream wait event");
    return 0;
}

static int coupled_next_mode(const CoupledSide& sd, int mode)
{
    for (int i = 0; i < sd.nseq; ++i) if (sd.seq[i] == mode) return sd.seq[(i + 1) % sd.nseq];
    return -1;
}

// M (in the side's workspace) of `mode` with the factors as they stand: a streaming pass only if the contraction this mode
// is served from depends on a factor that changed since it was formed (always, with a mask: the imputed entries follow
// every factor)
static int coupled_mttkrp(CoupledSide& sd, int mode, int R, int ldr, const int* done, cudaStream_t st)
{
    const Geo& g = sd.g; int rc;
    const float* const* A = (const float* const*)sd.A;
    const bool lead = mode < g.ndim - 2;
    const bool has_mask = sd.mask != nullptr;
    if (has_mask || (lead ? !sd.T_ok : !sd.U_ok)) {
        ContractArgs a = base_args(g, sd.X, sd.mask, A, ldr, R, done);
        bool masked = has_mask;
        if ((rc = maybe_impute(g, a, masked, sd.w, st)) != 0) return rc;
        if (lead) { if ((rc = launch_plane(g, a, masked, sd.w.T, st)) != 0) return rc; sd.T_ok = true; }
        else      { if ((rc = launch_fiber(g, a, masked, sd.w, st)) != 0) return rc; sd.U_ok = true; }
    }
    return launch_mode_mttkrp(g, A, mode, R, ldr, sd.w, sd.w.M, done, st);
}

static void coupled_touched(CoupledSide& sd, int mode)
{
    if (mode >= sd.g.ndim - 2) sd.T_ok = false; else sd.U_ok = false;
}

// Gram of the just-updated factor `mode` from the update kernel's partials, and the Gram-Hadamard product for the mode this
// tensor updates next
static int coupled_finalize(const CoupledCtx& c, int side, int mode, const double* gram_part, long long I, const int* done,
                            cudaStream_t st)
{
    const CoupledSide& sd = c.s[side];
    FinalizeArgs f = fin_args(sd.g, c.R, c.ldr, sd.w, done);
    f.k = mode; f.gram_part = gram_part; f.nblk = (int)((I + C2C_UPD_ROWS - 1) / C2C_UPD_ROWS);
    f.next = coupled_next_mode(sd, mode); f.G = sd.w.G;
    gram_finalize_kernel<<<1, C2C_SMALL_THREADS, 0, st>>>(f);
    CKL("gram_finalize launch");
    return 0;
}

static int coupled_enqueue_iteration(CoupledCtx& c)
{
    int rc;
    const int* done = c.state + 1;
    cudaStream_t st = c.st, st2 = c.st2;
    if ((rc = coupled_order(st, st2, c.ev12)) != 0) return rc;          // fork: tensor 2's stream starts after the previous iteration
    for (int k = 0; k < c.nsteps; ++k) {
        if (c.kind[k] < 2) {
            const int side = c.kind[k];
            CoupledSide& sd = c.s[side];
            cudaStream_t ss = side == 0 ? st : st2;
            const int mode = side == 0 ? c.ma[k] : c.mb[k];
            const long long I = sd.g.shape[mode];
            if ((rc = coupled_mttkrp(sd, mode, c.R, c.ldr, done, ss)) != 0) return rc;
            if ((rc = launch_mu(sd.A[mode], sd.w.M, sd.w.G, I, c.R, c.ldr, c.eps, sd.w, true, false, done, ss)) != 0) return rc;
            if ((rc = coupled_finalize(c, side, mode, sd.w.gram_part, I, done, ss)) != 0) return rc;
            coupled_touched(sd, mode);
        } else {
            const int a = c.ma[k], b = c.mb[k];
            const long long I = c.s[0].g.shape[a];
            if (k == c.nsteps - 1) {
                // last update of the iteration: a masked tensor is re-imputed here for the last time; the squared norm of the
                // imputed entries enters the direct error (coupled_error_kernel)
                for (int side = 0; side < 2; ++side) {
                    CoupledSide& sd = c.s[side];
                    if (sd.mask == nullptr) continue;
                    if ((rc = recon_sums_impl(sd.g, sd.w, sd.X, sd.mask, (const float* const*)sd.A, c.ldr, c.R, 1, sd.w.scratch_d + 4,
                                              side == 0 ? st : st2)) != 0) return rc;
                }
            }
            if ((rc = coupled_mttkrp(c.s[0], a, c.R, c.ldr, done, st)) != 0) return rc;
            if ((rc = coupled_mttkrp(c.s[1], b, c.R, c.ldr, done, st2)) != 0) return rc;
            if ((rc = coupled_order(st2, st, c.ev21)) != 0) return rc;  // M2 and G2 are ready
            const int nblk = (int)((I + C2C_UPD_ROWS - 1) / C2C_UPD_ROWS);
            const size_t smem = (size_t)3 * C2C_UPD_ROWS * c.R * sizeof(float);
            mu_update_coupled_kernel<<<nblk, C2C_SMALL_THREADS, smem, st>>>(
                c.s[0].A[a], c.s[1].A[b], c.s[0].w.M, c.s[1].w.M, c.s[0].w.G, c.s[1].w.G, I, c.R, c.ldr, c.eps,
                k == c.nsteps - 1 ? c.iprod : nullptr, c.s[0].w.gram_part, done);
            CKL("mu_update_coupled launch");
            if ((rc = coupled_order(st, st2, c.ev12)) != 0) return rc;  // the new factor and its Gram partials are ready
            if ((rc = coupled_finalize(c, 0, a, c.s[0].w.gram_part, I, done, st)) != 0) return rc;
            if ((rc = coupled_finalize(c, 1, b, c.s[0].w.gram_part, I, done, st2)) != 0) return rc;
            coupled_touched(c.s[0], a);
            coupled_touched(c.s[1], b);
        }
    }
    if ((rc = coupled_order(st2, st, c.ev21)) != 0) return rc;          // join
    CoupledErrArgs e;
    e.S1 = c.s[0].w.S; e.S2 = c.s[1].w.S; e.N1 = c.s[0].g.ndim; e.N2 = c.s[1].g.ndim; e.R = c.R; e.ldr = c.ldr;
    e.iprod = c.iprod; e.nx2_1 = c.nx2[0]; e.nx2_2 = c.nx2[1]; e.errors = c.errors; e.state = c.state;
    e.n_iter_max = c.n_iter_max; e.tol = c.tol; e.cvg = c.cvg; e.w1 = c.w1; e.w2 = c.w2;
    e.miss2_1 = c.s[0].mask ? c.s[0].w.scratch_d + 5 : nullptr;
    e.miss2_2 = c.s[1].mask ? c.s[1].w.scratch_d + 5 : nullptr;
    coupled_error_kernel<<<1, C2C_SMALL_THREADS, 0, st>>>(e);
    CKL("coupled_error launch");
    return 0;
}

// Iteration 0 runs eagerly (cold contraction caches).  The cache state at the end of an iteration depends only on the update
// order, so from iteration 1 on every iteration starts and ends in the same state: one is captured and replayed.
static int coupled_run_loop(CoupledCtx& c, int n_iter_max, int check_every)
{
    int rc;
    cudaStream_t st = c.st, st2 = c.st2;
    int* state = c.state;
    if ((rc = coupled_enqueue_iteration(c)) != 0) return rc;
    int it = 1;
    cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr;
    bool use_graph = n_iter_max >= 4;
    if (use_graph) {
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (st != nullptr && st != cudaStreamLegacy && st != cudaStreamPerThread &&
            (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone))
            use_graph = false;
    }
    long long per_graph = 0;
    if (use_graph) {
        const bool before_state[4] = {c.s[0].T_ok, c.s[0].U_ok, c.s[1].T_ok, c.s[1].U_ok};
        cudaStream_t cap = nullptr;
        CK(cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking), "create capture stream");
        cudaError_t e = cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal);
        if (e != cudaSuccess) { cudaStreamDestroy(cap); return cuda_fail(e, "begin capture"); }
        const long long before = g_launches.load(std::memory_order_relaxed);
        cudaStream_t cap2 = cap;
        if (st2 != st) CK(cudaStreamCreateWithFlags(&cap2, cudaStreamNonBlocking), "create capture stream");
        c.st = cap; c.st2 = cap2;                 // cap2 joins the capture through the first event wait and rejoins before the end
        rc = coupled_enqueue_iteration(c);
        c.st = st; c.st2 = st2;
        per_graph = g_launches.load(std::memory_order_relaxed) - before;
        g_launches.fetch_sub(per_graph, std::memory_order_relaxed);
        e = cudaStreamEndCapture(cap, &graph);
        cudaStreamDestroy(cap);
        if (cap2 != cap) cudaStreamDestroy(cap2);
        if (rc != 0) { if (graph) cudaGraphDestroy(graph); return rc; }
        if (e != cudaSuccess) return cuda_fail(e, "end capture");
        if (before_state[0] != c.s[0].T_ok || before_state[1] != c.s[0].U_ok || before_state[2] != c.s[1].T_ok || before_state[3] != c.s[1].U_ok) {
            cudaGraphDestroy(graph);
            return fail(-46, "coupled iteration is not in a steady cache state");
        }
        e = cudaGraphInstantiate(&exec, graph, 0);
        if (e != cudaSuccess) { cudaGraphDestroy(graph); return cuda_fail(e, "graph instantiate"); }
    }
    int host_state[2] = {0, 0};
    rc = 0;
    for (; it < n_iter_max; ++it) {
        if (use_graph) {
            cudaError_t e = cudaGraphLaunch(exec, st);
            if (e != cudaSuccess) { rc = cuda_fail(e, "graph launch"); break; }
            g_launches.fetch_add(per_graph, std::memory_order_relaxed);
        } else if ((rc = coupled_enqueue_iteration(c)) != 0) break;
        if (check_every > 0 && (it + 1) % check_every == 0 && it + 1 < n_iter_max) {
            cudaError_t e = cudaMemcpyAsync(host_state, state, 2 * sizeof(int), cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) { rc = cuda_fail(e, "poll state"); break; }
            if (host_state[1]) break;
        }
    }
    if (use_graph) {
        cudaError_t e = cudaStreamSynchronize(st);
        cudaGraphExecDestroy(exec); cudaGraphDestroy(graph);
        if (rc == 0 && e != cudaSuccess) rc = cuda_fail(e, "synchronize after graph replay");
    }
    return rc;
}

extern "C" int c2c_coupled_ncp_run(const float* X1, const uint8_t* mask1, int ndim1, const int64_t* shape1, float* const* A1,
                                   const float* X2, const uint8_t* mask2, int ndim2, const int64_t* shape2, float* const* A2,
                                   int ldr, int rank, const int* shared_pairs, int npairs, int n_iter_max, double tol,
                                   int cvg_criterion, double w1, double w2, const double* normX2, double* errors, int* state,
                                   int check_every, void* ws1, size_t ws1_bytes, void* ws2, size_t ws2_bytes, void* stream)
{
    static thread_local CoupledCtx c;          // large (two geometries): kept off the stack
    int rc;
    if (!shape1 || !shape2) return fail(-1, "shape is NULL");
    if (!shared_pairs || npairs < 1) return fail(-40, "at least one mode must be shared between the tensors");
    if (!errors || !state) return fail(-24, "errors / state is NULL");
    if (cvg_criterion != C2C_CVG_ABS_REC_ERROR && cvg_criterion != C2C_CVG_REC_ERROR) return fail(-22, "unknown cvg_criterion %d", cvg_criterion);
    if (n_iter_max < 0) return fail(-23, "n_iter_max < 0");
    if (!(w1 > 0.0) || !(w2 > 0.0)) return fail(-41, "error weights must be positive");
    const float* Xs[2] = {X1, X2}; const uint8_t* masks[2] = {mask1, mask2}; const int nds[2] = {ndim1, ndim2};
    const int64_t* shapes[2] = {shape1, shape2}; float* const* As[2] = {A1, A2}; void* wss[2] = {ws1, ws2};
    const size_t wsb[2] = {ws1_bytes, ws2_bytes};
    for (int k = 0; k < 2; ++k) {
        CoupledSide& sd = c.s[k];
        if ((rc = make_geo(sd.g, nds[k], shapes[k], rank)) != 0) return rc;
        if ((rc = check_common(Xs[k], (const float* const*)As[k], nds[k], ldr, rank)) != 0) return rc;
        if (((uintptr_t)Xs[k] & 15) != 0) return fail(-25, "X must be 16-byte aligned");
        if (sd.g.nchunk != 1) return fail(-42, "coupled factorization: rank must fit one chunk (<= 32)");
        if ((rc = get_ws(sd.w, sd.g, rank, masks[k] != nullptr, wss[k], wsb[k])) != 0) return rc;
        sd.X = Xs[k]; sd.mask = masks[k]; sd.A = As[k]; sd.T_ok = sd.U_ok = false; sd.nseq = 0;
    }
    if (ws1 == ws2) return fail(-43, "the two tensors need separate workspaces");
    bool sh1[C2C_MAX_NDIM] = {false}, sh2[C2C_MAX_NDIM] = {false};
    for (int p = 0; p < npairs; ++p) {
        const int a = shared_pairs[2 * p], b = shared_pairs[2 * p + 1];
        if (a < 0 || a >= ndim1 || b < 0 || b >= ndim2) return fail(-44, "shared pair (%d, %d) out of range", a, b);
        if (sh1[a] || sh2[b]) return fail(-44, "a mode can be shared only once");
        if (shape1[a] != shape2[b]) return fail(-45, "shared modes must have the same size: %lld vs %lld", (long long)shape1[a], (long long)shape2[b]);
        sh1[a] = sh2[b] = true;
    }
    // update order (coupled_factorization.py:309-320): tensor-1-only modes, tensor-2-only modes, shared pairs in reverse
    c.nsteps = 0;
    for (int m = 0; m < ndim1; ++m) if (!sh1[m]) { c.kind[c.nsteps] = 0; c.ma[c.nsteps] = m; c.mb[c.nsteps] = -1; ++c.nsteps; c.s[0].seq[c.s[0].nseq++] = m; }
    for (int m = 0; m < ndim2; ++m) if (!sh2[m]) { c.kind[c.nsteps] = 1; c.ma[c.nsteps] = -1; c.mb[c.nsteps] = m; ++c.nsteps; c.s[1].seq[c.s[1].nseq++] = m; }
    for (int p = npairs - 1; p >= 0; --p) {
        c.kind[c.nsteps] = 2; c.ma[c.nsteps] = shared_pairs[2 * p]; c.mb[c.nsteps] = shared_pairs[2 * p + 1]; ++c.nsteps;
        c.s[0].seq[c.s[0].nseq++] = shared_pairs[2 * p]; c.s[1].seq[c.s[1].nseq++] = shared_pairs[2 * p + 1];
    }
    c.ldr = ldr; c.R = rank; c.n_iter_max = n_iter_max; c.tol = tol; c.cvg = cvg_criterion; c.w1 = w1; c.w2 = w2;
    c.eps = 1.1920928955078125e-07f;
    c.errors = errors; c.state = state; c.st = (cudaStream_t)stream;
    cudaStream_t st = c.st;
    c.iprod = c.s[0].w.scratch_d;
    CK(cudaMemsetAsync(state, 0, 2 * sizeof(int), st), "memset state");
    CK(cudaMemsetAsync(c.iprod, 0, 2 * sizeof(double), st), "memset iprod");
    for (int k = 0; k < 2; ++k) {
        CoupledSide& sd = c.s[k];
        if (normX2) c.nx2[k] = normX2 + k;
        else {
            if ((rc = sumsq_into(sd.X, sd.g.P * sd.g.E, sd.w.normX2, sd.w.sums_part, st)) != 0) return rc;
            c.nx2[k] = sd.w.normX2;
        }
        // Grams of the initial factors; G for the first mode this tensor updates
        for (int m = 0; m < sd.g.ndim; ++m) {
            if ((rc = launch_gram_partial(sd.A[m], nullptr, sd.g.shape[m], rank, ldr, sd.w, false, nullptr, st)) != 0) return rc;
            FinalizeArgs f = fin_args(sd.g, rank, ldr, sd.w, nullptr);
            f.k = m; f.gram_part = sd.w.gram_part; f.nblk = (int)((sd.g.shape[m] + C2C_UPD_ROWS - 1) / C2C_UPD_ROWS);
            if (m == sd.g.ndim - 1) { f.next = sd.seq[0]; f.G = sd.w.G; }
            gram_finalize_kernel<<<1, C2C_SMALL_THREADS, 0, st>>>(f);
            CKL("gram_finalize launch");
        }
    }
    if (n_iter_max == 0) return 0;

    // tensor 2's chain gets a stream of its own (C2C_B200_COUPLED_ONE_STREAM=1: everything on the caller's stream)
    static int one_stream = -1;
    if (one_stream < 0) one_stream = env_int("C2C_B200_COUPLED_ONE_STREAM", 0) ? 1 : 0;
    c.st2 = st; c.ev12 = nullptr; c.ev21 = nullptr;
    cudaStream_t own2 = nullptr;
    if (!one_stream) {
        CK(cudaStreamCreateWithFlags(&own2, cudaStreamNonBlocking), "create second stream");
        cudaError_t e = cudaEventCreateWithFlags(&c.ev12, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c.ev21, cudaEventDisableTiming);
        if (e != cudaSuccess) { cudaStreamDestroy(own2); if (c.ev12) cudaEventDestroy(c.ev12); return cuda_fail(e, "create events"); }
        c.st2 = own2;
    }
    rc = coupled_run_loop(c, n_iter_max, check_every);
    if (own2) {
        cudaStreamSynchronize(own2);
        cudaStreamDestroy(own2); cudaEventDestroy(c.ev12); cudaEventDestroy(c.ev21);
    }
    return rc;
}
