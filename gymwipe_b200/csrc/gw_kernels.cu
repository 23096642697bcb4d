// gymwipe_b200 -- sm_100a kernels and the C ABI (include/gymwipe_b200.h).
//
// Kernels
//   K1  tables_kernel        FSPL attenuation + received-power tables from device positions
//   K2  ber_kernel           standalone BPSK BER (numeric parity tests)
//   K3  count_bits_kernel    HBM-streaming popcount of error-mask rows, one warp per descriptor
//   K4  step_kernel          fused event-ordered step: RRM announcement, MAC window, PHY
//                            transmissions, SINR/BER segments, decider, delivery, interpreter
//   K5  (epilogue of K4)     per-block reduction of reward / delivery statistics
//
// Layout of the env-batch state (structure of arrays, one 16-byte chunk per thread and
// field group so that a warp's loads/stores are 512 contiguous bytes, 128-bit per lane):
//   now   [n_envs]                       fp64
//   hot   [HOT_CHUNKS ][n_sims] x 16 B   always read at step start / written at step end
//   cold  [COLD_CHUNKS][n_sims] x 16 B   only touched when a transmission, reception or MAC
//                                        window is in flight across a step boundary
//   ring  [2*100][n_sims] int32          sizes of queued packets that predate a reset()
//   att / srx [16][n_tables] fp64        attenuation (dB) and received power (mW) tables
//
// Compile: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -lineinfo
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <mutex>
#include <new>
#include <vector>

#include "../../include/gymwipe_b200.h"

#include "gw_core.cuh"
#include "gw_pendulum.cuh"
#include "gw_grid.cuh"
#include "gw_band.cuh"

using namespace gw;

// ------------------------------------------------------------------------------------
// error handling
// ------------------------------------------------------------------------------------

static thread_local char g_err[512] = "";
// handle generations (gw_handle::generation) are unique over all handles and calls of the process: a handle
// allocated at the address of a destroyed one never matches a cached launch graph of its predecessor
static std::atomic<unsigned long long> g_generation{1};
static unsigned long long next_generation() { return g_generation.fetch_add(1); }

static int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(expr)                                                                         \
    do {                                                                                       \
        cudaError_t e_ = (expr);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            return fail(GW_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

// ------------------------------------------------------------------------------------
// state layout
// ------------------------------------------------------------------------------------

enum : int {
    H_P01 = 0, H_P23, H_TICK, H_TICKS, H_U0, H_U1, H_U2, H_JAM, H_EP, H_EPC, H_RCV0, H_RCV1, HOT_CHUNKS
};
// cold: per device 5 chunks, then per sender 1
enum : int { C_EV = 0, C_TX, C_RX, C_RT, C_U, C_V, C_PER_DEV };
constexpr int COLD_CHUNKS = C_PER_DEV * kMaxDev + kMaxSend;

struct StatePtrs {
    long long nsim, nenv, ntab;
    int nb;
    int per_env;        // 1: one attenuation / power table per band-sim (ntab == nsim), 0: one shared table
    double *now;
    uint4 *hot;
    uint4 *cold;
    int32_t *ring;
    double *att;
    double *srx;
    double *plant;      // [8][nsim] plant state (plant envs)
    double *pval;       // [2*100][nsim] values of queued packets (plant envs)
};

struct Layout {
    size_t off_now, off_hot, off_cold, off_ring, off_att, off_srx, off_plant, off_pval, total;
};

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static Layout make_layout(long long nenv, int nb, long long ntab, int plant = 0)
{
    const long long nsim = nenv * nb;
    Layout L;
    size_t o = 0;
    L.off_now = o; o = align_up(o + sizeof(double) * nenv, 256);
    L.off_hot = o; o = align_up(o + 16ull * HOT_CHUNKS * nsim, 256);
    L.off_cold = o; o = align_up(o + 16ull * COLD_CHUNKS * nsim, 256);
    L.off_ring = o; o = align_up(o + 4ull * kMaxSend * kRingSlots * nsim, 256);
    L.off_att = o; o = align_up(o + 8ull * 16 * ntab, 256);
    L.off_srx = o; o = align_up(o + 8ull * 16 * ntab, 256);
    L.off_plant = o; if (plant) o = align_up(o + 8ull * 8 * nsim, 256);
    L.off_pval = o; if (plant) o = align_up(o + 8ull * kMaxSend * kRingSlots * nsim, 256);
    L.total = o;
    return L;
}

// ------------------------------------------------------------------------------------
// handle
// ------------------------------------------------------------------------------------

struct gw_handle {
    gw_config cfg;
    int device;
    Params P;
    StatePtrs st;
    Layout layout;
    void *owned_state;
    double *stats;          // device [8]
    double *stats_use;      // accumulators the step kernels add to: `stats`, or another handle's (gw_share_stats)
    int *errflag;           // device [4]: code, sim, fault, -
    unsigned long long *mask_bytes;     // device [2]: mask bytes scanned by the mode-M (fed) step kernels, -
    double power_dbm[kMaxBands][kMaxDev];
    double default_pos[kMaxBands][kMaxDev][2];
    double thermal[kMaxBands];
    double frequency[kMaxBands];
    int D, NS, NJ;
    // staging for gw_step_host
    int32_t *d_dev, *d_dur;
    long long *d_obs;
    double *d_rew;
    unsigned char *d_done;
    // mode M fed masks
    const uint32_t *masks;
    int mask_slots, mask_words;
    // prefix-count index of the fed masks (gw_set_masks; see mask_index_kernel)
    uint16_t *mask_t16;
    uint32_t *mask_s32;
    int mask_gt, mask_nsb;
    size_t mask_idx_bytes;
    // BER memo
    ulonglong2 *memo;
    unsigned memo_entries;
    ulonglong2 *memo_sim;   // per-env geometries: [64][n_sims] per-band-sim BER cache (DevMemo::sim)
    PendulumParams pend;
    int pdl;                // launch the step kernel with programmatic stream serialization
    double *pos_cur;        // per-env geometry: current device positions [ntab][GW_MAX_DEVICES][2]
    int stepped;            // a step has been launched: gw_set_positions moves devices from now on
    // bit v: the dynamic shared-memory limit of step-kernel variant v (0 traced, 1 mode R, 2 Philox, 3 fed)
    // has been raised on THIS handle's device (the attribute is per device, and a handle is driven by
    // one host thread at a time -- include/gymwipe_b200.h)
    unsigned smem_configured;
    int ext;                // some band uses MAC receive mode / finite bursts: the EXT kernels step this handle
    unsigned long long *stamps;         // gw_debug_stamps: device buffer [stamp_cap][4], next slot
    long long stamp_cap, stamp_next;
    unsigned long long generation;      // bumped by every call that changes what a step launch is given (masks, shared
                                        // statistics, positions, stamps): cached launch graphs are keyed by it
};

// ------------------------------------------------------------------------------------
// shared-memory staging of the per-device state of every band-sim of a block: the arrays the
// transition function indexes with run-time device indices (active transmissions, receptions,
// MAC windows) live in dynamic shared memory, laid out [field][index][thread] -- consecutive
// threads hit consecutive banks, and a run-time index is a plain address computation.
// ------------------------------------------------------------------------------------

#ifndef GW_STEP_BLOCK
#define GW_STEP_BLOCK 128
#endif
constexpr int STEP_BLOCK = GW_STEP_BLOCK;

template <class T, int N, int OFF>
struct ShArr {
    using value_type = T;
    static constexpr int size = N;
    static constexpr bool direct = true;
    __host__ __device__ __forceinline__ T &operator[](int i) const
    {
#ifdef __CUDA_ARCH__
        extern __shared__ __align__(16) unsigned char gw_step_smem[];
        return *reinterpret_cast<T *>(gw_step_smem + (size_t)(OFF + i * (int)sizeof(T)) * STEP_BLOCK + threadIdx.x * sizeof(T));
#else
        static T dummy;
        (void)i;
        return dummy;       // never executed: the shared storage exists on the device only
#endif
    }
};

struct ShStore {
    static constexpr bool roll = true;                      // arrays in shared memory: a run-time index is an address
    template <class T, int N, int OFF> using Arr = ShArr<T, N, OFF>;
    template <class T, int N> using Aux = RegArr<T, N>;     // dead (mode R) or few (mode M) registers
};

// ------------------------------------------------------------------------------------
// device helpers: chunk access, pack / unpack
// ------------------------------------------------------------------------------------

__device__ __forceinline__ uint4 ld_chunk(const uint4 *base, long long nsim, int c, long long i)
{
    return base[(long long)c * nsim + i];
}
__device__ __forceinline__ void st_chunk(uint4 *base, long long nsim, int c, long long i, uint4 v)
{
    base[(long long)c * nsim + i] = v;
}
__device__ __forceinline__ uint4 pack_dd(double a, double b)
{
    uint4 v;
    const unsigned long long x = (unsigned long long)__double_as_longlong(a), y = (unsigned long long)__double_as_longlong(b);
    v.x = (unsigned)x; v.y = (unsigned)(x >> 32); v.z = (unsigned)y; v.w = (unsigned)(y >> 32);
    return v;
}
__device__ __forceinline__ double lo_d(uint4 v) { return __longlong_as_double((long long)(((unsigned long long)v.y << 32) | v.x)); }
__device__ __forceinline__ double hi_d(uint4 v) { return __longlong_as_double((long long)(((unsigned long long)v.w << 32) | v.z)); }
__device__ __forceinline__ uint4 pack_qq(unsigned long long x, unsigned long long y)
{
    uint4 v; v.x = (unsigned)x; v.y = (unsigned)(x >> 32); v.z = (unsigned)y; v.w = (unsigned)(y >> 32); return v;
}
__device__ __forceinline__ unsigned long long lo_q(uint4 v) { return ((unsigned long long)v.y << 32) | v.x; }
__device__ __forceinline__ unsigned long long hi_q(uint4 v) { return ((unsigned long long)v.w << 32) | v.z; }

// flagsA: sphase[4] 3b | (rxOf+1)[4] 3b | rxSec[4] 1b | mac[2] 2b
// flagsB: qn[2] 7b | wDone[2] | wPend[2] | jamStage 2b | jamPending 6b | rv0!=0 | rv1!=0 | lastAbs!=0 | done | busy
template <int D, int NS, int NJ, class ST>
__device__ __forceinline__ void unpack_flags(Sim<D, NS, NJ, ST> &s, unsigned a, unsigned b)
{
#pragma unroll (Sim<D, NS, NJ, ST>::kUnrollIO)
    for (int d = 0; d < D; ++d) {
        s.sphase[d] = (a >> (3 * d)) & 7;
        s.rxOf[d] = (int)((a >> (12 + 3 * d)) & 7) - 1;
        s.rxSec[d] = (a >> (24 + d)) & 1;
    }
#pragma unroll
    for (int k = 0; k < NS; ++k) {
        s.mac[k] = (a >> (28 + 2 * k)) & 3;
        s.qn[k] = (b >> (7 * k)) & 127;
        s.wDone[k] = (b >> (14 + k)) & 1;
        s.wPend[k] = (b >> (16 + k)) & 1;
    }
    s.jamStage[0] = (b >> 18) & 3;
    s.jamPending[0] = (b >> 20) & 63;
    s.rv0 = ((b >> 26) & 1) ? kCounterByteLen : 0;
    s.rv1 = ((b >> 27) & 1) ? kCounterByteLen : 0;
    s.latestDiff = s.rv0 - s.rv1;
    s.lastAbsDiff = ((b >> 28) & 1) ? kCounterByteLen : 0;
    s.done = (b >> 29) & 1;
}

template <int D, int NS, int NJ, class ST>
__device__ __forceinline__ bool sim_busy(const Sim<D, NS, NJ, ST> &s)
{
    bool busy = false;
#pragma unroll (Sim<D, NS, NJ, ST>::kUnrollIO)
    for (int d = 0; d < D; ++d) busy |= (s.sphase[d] != S_IDLE) | (s.rxOf[d] >= 0);
#pragma unroll
    for (int k = 0; k < NS; ++k) busy |= (s.mac[k] != MAC_NONE) | (s.wPend[k] != 0);
    return busy;
}

template <int D, int NS, int NJ, class ST>
__device__ __forceinline__ void pack_flags(const Sim<D, NS, NJ, ST> &s, bool busy, unsigned &a, unsigned &b)
{
    a = 0; b = 0;
#pragma unroll (Sim<D, NS, NJ, ST>::kUnrollIO)
    for (int d = 0; d < D; ++d) {
        a |= (unsigned)s.sphase[d] << (3 * d);
        a |= (unsigned)(s.rxOf[d] + 1) << (12 + 3 * d);
        a |= (unsigned)s.rxSec[d] << (24 + d);
    }
#pragma unroll
    for (int k = 0; k < NS; ++k) {
        a |= (unsigned)s.mac[k] << (28 + 2 * k);
        b |= (unsigned)s.qn[k] << (7 * k);
        b |= (unsigned)s.wDone[k] << (14 + k);
        b |= (unsigned)s.wPend[k] << (16 + k);
    }
    b |= (unsigned)s.jamStage[0] << 18;
    b |= (unsigned)s.jamPending[0] << 20;
    b |= (unsigned)(s.rv0 != 0) << 26;
    b |= (unsigned)(s.rv1 != 0) << 27;
    b |= (unsigned)(s.lastAbsDiff != 0) << 28;
    b |= (unsigned)(s.done != 0) << 29;
    b |= (unsigned)busy << 31;
}

// FULL = false: the step kernels skip fields they never read (counter epochs -- fetched on demand
// through DevRing --, mode-M segment starts in mode R, packet values without a plant), which
// makes them dead in registers
template <bool FULL = true, bool WITH_SEG = true, int D, int NS, int NJ, class ST>
__device__ __forceinline__ bool load_sim(Sim<D, NS, NJ, ST> &s, const StatePtrs &st, long long i, double now,
                                         bool with_seq = true)
{
    const long long n = st.nsim;
    s.now = now;
    s.pchg = -1;            // not kept across steps: the first BER of a reception in flight is re-evaluated
    uint4 v = ld_chunk(st.hot, n, H_P01, i);
    s.P[0] = lo_d(v); s.P[1] = hi_d(v);
    v = ld_chunk(st.hot, n, H_P23, i);
    s.P[2] = lo_d(v);
    if (D > 3) s.P[D > 3 ? 3 : 0] = hi_d(v);
    v = ld_chunk(st.hot, n, H_TICK, i);
    s.tTick[0] = lo_d(v); s.tTick[1] = hi_d(v);
    v = ld_chunk(st.hot, n, H_TICKS, i);
    s.ticks[0] = lo_q(v); s.ticks[1] = hi_q(v);
    const uint4 u0 = ld_chunk(st.hot, n, H_U0, i);
    s.seq = u0.x; s.nTx = u0.w;
    unpack_flags(s, u0.y, u0.z);
    v = ld_chunk(st.hot, n, H_U1, i);
    s.sTick[0] = v.x; s.sTick[1] = v.y; s.nDeliv[0] = v.z; s.nDeliv[1] = v.w;
    // per-device transmission numbers: mode-M mask keys and the lazily created attenuation models of
    // moving devices read them; mode R with one shared geometry does not (with_seq = false: not kept)
    v = with_seq ? ld_chunk(st.hot, n, H_U2, i) : make_uint4(0, 0, 0, 0);
    s.txSeq[0] = v.x; s.txSeq[1] = v.y; s.txSeq[2] = v.z;
    if (D > 3) s.txSeq[D > 3 ? 3 : 0] = v.w;
    if (NJ > 0) {
        v = ld_chunk(st.hot, n, H_JAM, i);
        s.tJam[0] = lo_d(v); s.sJam[0] = v.z;
    } else {
        s.tJam[0] = 0; s.sJam[0] = 0;
    }
    if (FULL) {
        // one chunk {epochK[0] | epochC[0] << 63, epochK[1] | epochC[1] << 63}: the counter epoch of each sender
        // (counter value 0 or 1 at tick epochK; packets enqueued before that tick are in the snapshot ring)
        v = ld_chunk(st.hot, n, H_EP, i);
        s.epochK[0] = lo_q(v) & ~(1ull << 63); s.epochC[0] = (int)(lo_q(v) >> 63);
        s.epochK[1] = hi_q(v) & ~(1ull << 63); s.epochC[1] = (int)(hi_q(v) >> 63);
        v = ld_chunk(st.hot, n, H_EPC, i);
        s.fault = (int)v.z; s.ties = v.w;
    } else {
        s.epochK[0] = s.epochK[1] = 0; s.epochC[0] = s.epochC[1] = 0;
        s.fault = 0; s.ties = 0;        // a faulted sim stays flagged in memory (store_sim ORs nothing back)
    }
    const bool busy = (u0.z >> 31) & 1;
    if (busy) {
#pragma unroll (Sim<D, NS, NJ, ST>::kUnrollIO)
        for (int d = 0; d < D; ++d) {
            uint4 c = ld_chunk(st.cold, n, d * C_PER_DEV + C_EV, i);
            s.tEv[d] = lo_d(c); s.tC[d] = hi_d(c);
            c = ld_chunk(st.cold, n, d * C_PER_DEV + C_TX, i);
            s.txStart[d] = lo_d(c); s.tStop[d] = hi_d(c);
            c = ld_chunk(st.cold, n, d * C_PER_DEV + C_RX, i);
            s.ber[d] = lo_d(c); s.err[d] = hi_d(c);
            c = ld_chunk(st.cold, n, d * C_PER_DEV + C_RT, i);
            s.tReset[d] = lo_d(c); set_at(s.segT0, d, WITH_SEG ? hi_d(c) : 0.0);
            c = ld_chunk(st.cold, n, d * C_PER_DEV + C_U, i);
            s.sEv[d] = c.x; s.sC[d] = c.y; s.cmdPay[d] = (int)c.z;
            if (FULL && st.plant) { c = ld_chunk(st.cold, n, d * C_PER_DEV + C_V, i); set_at(s.txVal, d, lo_d(c)); } else set_at(s.txVal, d, 0.0);
        }
#pragma unroll
        for (int k = 0; k < NS; ++k) {
            const uint4 c = ld_chunk(st.cold, n, C_PER_DEV * kMaxDev + k, i);
            s.stopW[k] = lo_d(c); s.sW[k] = c.z;
        }
    } else {
#pragma unroll (Sim<D, NS, NJ, ST>::kUnrollIO)
        for (int d = 0; d < D; ++d) {
            s.tEv[d] = 0; s.tC[d] = 0; s.txStart[d] = 0; s.tStop[d] = 0; s.ber[d] = 0; s.err[d] = 0;
            s.tReset[d] = 0; set_at(s.segT0, d, 0.0); s.sEv[d] = 0; s.sC[d] = 0; s.cmdPay[d] = 0; set_at(s.txVal, d, 0.0);
        }
#pragma unroll
        for (int k = 0; k < NS; ++k) { s.stopW[k] = 0; s.sW[k] = 0; }
    }
    s.annDest = 0; s.annBytes = 0; s.annSlots = 0; s.rrmPend = 0; s.tRrm = 0; s.sRrm = 0; s.assignDone = 0;
    s.trace = nullptr; s.ntrace = 0; s.traceCap = 0;
    return busy;        // something was in flight across the step boundary (sim_busy at the last store)
}

template <bool FULL = true, bool WITH_SEG = true, int D, int NS, int NJ, class ST>
__device__ __forceinline__ void store_sim(const Sim<D, NS, NJ, ST> &s, const StatePtrs &st, long long i, bool epoch_too,
                                          bool with_seq = true, int keep_p = 0)
{
    const long long n = st.nsim;
    // keep_p bit 0 / 1: the received powers of chunk P01 / P23 equal what the step loaded (the usual case:
    // every transmission adds and removes its power, and the rounding residue settles) -- not written back
    if (!(keep_p & 1)) st_chunk(st.hot, n, H_P01, i, pack_dd(s.P[0], s.P[1]));
    if (!(keep_p & 2)) st_chunk(st.hot, n, H_P23, i, pack_dd(s.P[2], D > 3 ? s.P[D > 3 ? 3 : 0] : 0.0));
    st_chunk(st.hot, n, H_TICK, i, pack_dd(s.tTick[0], s.tTick[1]));
    st_chunk(st.hot, n, H_TICKS, i, pack_qq(s.ticks[0], s.ticks[1]));
    const bool busy = sim_busy(s);
    uint4 u0;
    u0.x = s.seq; u0.w = s.nTx;
    pack_flags(s, busy, u0.y, u0.z);
    st_chunk(st.hot, n, H_U0, i, u0);
    uint4 v;
    v.x = s.sTick[0]; v.y = s.sTick[1]; v.z = s.nDeliv[0]; v.w = s.nDeliv[1];
    st_chunk(st.hot, n, H_U1, i, v);
    if (with_seq) {
        v.x = s.txSeq[0]; v.y = s.txSeq[1]; v.z = s.txSeq[2]; v.w = D > 3 ? s.txSeq[D > 3 ? 3 : 0] : 0u;
        st_chunk(st.hot, n, H_U2, i, v);
    }
    if (NJ > 0) {
        v = pack_dd(s.tJam[0], 0.0);
        v.z = s.sJam[0]; v.w = 0;
        st_chunk(st.hot, n, H_JAM, i, v);
    }
    if (epoch_too)
        st_chunk(st.hot, n, H_EP, i, pack_qq(s.epochK[0] | ((unsigned long long)(s.epochC[0] & 1) << 63),
                                             s.epochK[1] | ((unsigned long long)(s.epochC[1] & 1) << 63)));
    if (FULL) {
        if (epoch_too || s.fault || s.ties) {
            v.x = 0; v.y = 0; v.z = (unsigned)s.fault; v.w = s.ties;
            st_chunk(st.hot, n, H_EPC, i, v);
        }
    } else if (s.fault || s.ties) {
        // step kernels: fault / tie counters are merged into the stored words (rare path)
        unsigned *w = reinterpret_cast<unsigned *>(st.hot + (long long)H_EPC * n + i);
        if (s.fault) w[2] = (unsigned)s.fault;
        w[3] += s.ties;
    }
    if (busy) {
#pragma unroll (Sim<D, NS, NJ, ST>::kUnrollIO)
        for (int d = 0; d < D; ++d) {
            st_chunk(st.cold, n, d * C_PER_DEV + C_EV, i, pack_dd(s.tEv[d], s.tC[d]));
            st_chunk(st.cold, n, d * C_PER_DEV + C_TX, i, pack_dd(s.txStart[d], s.tStop[d]));
            st_chunk(st.cold, n, d * C_PER_DEV + C_RX, i, pack_dd(s.ber[d], s.err[d]));
            st_chunk(st.cold, n, d * C_PER_DEV + C_RT, i, pack_dd(s.tReset[d], WITH_SEG ? get_at(s.segT0, d) : 0.0));
            uint4 c; c.x = s.sEv[d]; c.y = s.sC[d]; c.z = (unsigned)s.cmdPay[d]; c.w = 0;
            st_chunk(st.cold, n, d * C_PER_DEV + C_U, i, c);
            if (FULL && st.plant) st_chunk(st.cold, n, d * C_PER_DEV + C_V, i, pack_dd(get_at(s.txVal, d), 0.0));
        }
#pragma unroll
        for (int k = 0; k < NS; ++k) {
            uint4 c = pack_dd(s.stopW[k], 0.0);
            c.z = s.sW[k]; c.w = 0;
            st_chunk(st.cold, n, C_PER_DEV * kMaxDev + k, i, c);
        }
    }
}

// EXT: the band-sims of this handle may use MAC receive mode / finite traffic bursts (gw_device_config.receive /
// max_ticks); the kernels of everybody else are compiled without that code (gw_core.cuh tests `Ring::ext`)
template <bool EXT>
struct DevRingT {
    static constexpr bool ext = EXT;
    int32_t *base;      // ring + sim index
    long long nsim;
    const uint4 *hot;   // hot chunks + sim index
    // counter epochs {epochK | epochC << 63} of the two senders, one chunk: read where head_size() needs
    // them (once per window / packet).  The step kernels prefetch the chunk into L2 together with the
    // state loads, so the first read is an L2 hit and the following ones hit L1 -- without holding it in
    // registers for the whole step.
    __device__ __forceinline__ int operator()(int k, uint32_t slot) const { return base[(long long)(k * kRingSlots + slot) * nsim]; }
    __device__ __forceinline__ void operator()(int k, uint32_t slot, int v) { base[(long long)(k * kRingSlots + slot) * nsim] = v; }
    __device__ __forceinline__ unsigned long long ep(int k) const
    {
        return reinterpret_cast<const unsigned long long *>(hot + (long long)H_EP * nsim)[k];
    }
    template <class S> __device__ __forceinline__ unsigned long long epochK(const S &, int k) const { return ep(k) & ~(1ull << 63); }
    template <class S> __device__ __forceinline__ int epochC(const S &, int k) const { return (int)(ep(k) >> 63); }
    // MAC receive mode: chunk H_RCV0 = {time-out of sender 0's RECEIVE command, of sender 1's}, chunk H_RCV1 =
    // {their creation numbers, packets handed to onReceive per sender}; touched only by bands in receive mode
    __device__ __forceinline__ double rxT(int k) const { return reinterpret_cast<const double *>(hot + (long long)H_RCV0 * nsim)[k]; }
    __device__ __forceinline__ uint32_t rxS(int k) const { return reinterpret_cast<const uint32_t *>(hot + (long long)H_RCV1 * nsim)[k]; }
    __device__ __forceinline__ void set_rx(int k, double t, uint32_t q) const
    {
        const_cast<double *>(reinterpret_cast<const double *>(hot + (long long)H_RCV0 * nsim))[k] = t;
        const_cast<uint32_t *>(reinterpret_cast<const uint32_t *>(hot + (long long)H_RCV1 * nsim))[k] = q;
    }
    __device__ __forceinline__ void add_received(int k) const
    {
        const_cast<uint32_t *>(reinterpret_cast<const uint32_t *>(hot + (long long)H_RCV1 * nsim))[2 + k] += 1u;
    }
    __device__ __forceinline__ uint32_t received(int k) const { return reinterpret_cast<const uint32_t *>(hot + (long long)H_RCV1 * nsim)[2 + k]; }
};
using DevRing = DevRingT<true>;

__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// ------------------------------------------------------------------------------------
// BER memo: direct-mapped table of exact (S, N) -> BER results in device memory (L2 / L1
// resident).  An entry is 32 bytes {S bits, N bits, BER bits, S ^ N ^ BER ^ MAGIC}; it is
// written with two 16-byte stores and validated by the checksum, so a torn or stale read is
// a miss, never a wrong value.  Because BER(S, N) is a pure function evaluated by the same
// device code, a hit returns bit-for-bit what the evaluation would.
// ------------------------------------------------------------------------------------

constexpr int MEMO_L0 = 32;     // entries of the block-local first level of the BER memo

#ifdef GW_MEMO_STATS
// instrumented builds only (-DGW_MEMO_STATS): {evaluations, second-level hits} since load
__device__ unsigned long long g_memo_stats[2];
#define GW_MEMO_COUNT(i) atomicAdd(&g_memo_stats[i], 1ull)
#else
#define GW_MEMO_COUNT(i) ((void)0)
#endif

struct DevMemo {
    ulonglong2 *tab;        // 2 x ulonglong2 per entry
    unsigned mask;          // entries - 1 (power of two)
    // first level (shared geometry only): MEMO_L0 direct-mapped entries that every block copies
    // from `l0g` into shared memory (`l0s`) together with its state loads; a hit costs two
    // shared-memory reads instead of a trip to L2 / DRAM.  Filled from the second level.
    ulonglong2 *l0g;
    const ulonglong2 *l0s;
    // per-env geometries: a cache of the band-sim's own -- [64][n_sims] x 16 B = {S, N} and {BER, -} for each of
    // the 16 links x 2 ways (gw_core.cuh::memo_slot); `sim` points at this thread's band-sim, entries `stride`
    // apart.  Only the owning thread reads and writes it: no validation beyond the exact (S, N) match.
    ulonglong2 *sim;
    long long stride;
    static constexpr unsigned long long MAGIC = 0x9E3779B97F4A7C15ull;
    __device__ __forceinline__ unsigned slot(unsigned long long a, unsigned long long b) const
    {
        unsigned long long h = a * 0x9E3779B97F4A7C15ull ^ (b + 0xC2B2AE3D27D4EB4Full) * 0xD6E8FEB86659FD93ull;
        return (unsigned)(h >> 40) & mask;
    }
    __device__ __forceinline__ bool get(double S, double N, double &ber, int link = 0) const
    {
        if (sim) {
            const ulonglong2 k = sim[(long long)(2 * link) * stride];
            if (k.x == (unsigned long long)__double_as_longlong(S) && k.y == (unsigned long long)__double_as_longlong(N)) {
                ber = __longlong_as_double((long long)sim[(long long)(2 * link + 1) * stride].x);
                return true;
            }
            return false;
        }
        if (!tab) return false;
        const unsigned long long a = (unsigned long long)__double_as_longlong(S), b = (unsigned long long)__double_as_longlong(N);
        const unsigned h = slot(a, b);
        if (l0s) {
            const unsigned h0 = h & (MEMO_L0 - 1);
            const ulonglong2 f0 = l0s[2 * h0], f1 = l0s[2 * h0 + 1];
            if (f0.x == a && f0.y == b && f1.y == (a ^ b ^ f1.x ^ MAGIC)) { ber = __longlong_as_double((long long)f1.x); return true; }
        }
        const ulonglong2 e0 = tab[2 * h], e1 = tab[2 * h + 1];
        if (e0.x == a && e0.y == b && e1.y == (a ^ b ^ e1.x ^ MAGIC)) {
            ber = __longlong_as_double((long long)e1.x);
            GW_MEMO_COUNT(1);
            if (l0g) { const unsigned h0 = h & (MEMO_L0 - 1); l0g[2 * h0] = e0; l0g[2 * h0 + 1] = e1; }
            return true;
        }
        return false;
    }
    __device__ __forceinline__ void put(double S, double N, double ber, int link = 0) const
    {
        if (sim) {
            sim[(long long)(2 * link) * stride] = make_ulonglong2((unsigned long long)__double_as_longlong(S), (unsigned long long)__double_as_longlong(N));
            sim[(long long)(2 * link + 1) * stride] = make_ulonglong2((unsigned long long)__double_as_longlong(ber), 0ull);
            return;
        }
        if (!tab) return;
        const unsigned long long a = (unsigned long long)__double_as_longlong(S), b = (unsigned long long)__double_as_longlong(N);
        const unsigned long long c = (unsigned long long)__double_as_longlong(ber);
        const unsigned h = slot(a, b);
        GW_MEMO_COUNT(0);
        tab[2 * h] = make_ulonglong2(a, b);
        tab[2 * h + 1] = make_ulonglong2(c, a ^ b ^ c ^ MAGIC);
        if (l0g) {
            const unsigned h0 = h & (MEMO_L0 - 1);
            l0g[2 * h0] = make_ulonglong2(a, b);
            l0g[2 * h0 + 1] = make_ulonglong2(c, a ^ b ^ c ^ MAGIC);
        }
    }
};

// ------------------------------------------------------------------------------------
// mode M: warp-cooperative error counting (bit-error masks)
// ------------------------------------------------------------------------------------

struct MaskSource {
    int mode;                   // MODE_M_PHILOX or MODE_M_FED
    unsigned long long seed;
    long long env_offset;
    const uint32_t *words;      // fed masks
    int slots, words_per_row;
    // prefix-count index of the fed masks (mask_index_kernel), or NULL
    const uint16_t *t16;        // [rows][gt]: set bits of groups [first group of g's superblock, g) of the row
    const uint32_t *s32;        // [rows][nsb]: set bits before superblock sb (512 groups); NULL when rows have one superblock
    int gt, nsb;
};

// number of set bits among bits [k0, k1) of a row of 32-bit words; all 32 lanes cooperate,
// 128-bit streaming loads issued four at a time per lane (2 KiB of a row in flight per warp),
// result valid in every lane.  Interior 128-bit groups are counted unmasked (4 POPC); only the
// first and the last group of the range get a prefix / suffix correction.
#ifndef GW_MASK_LOAD
#define GW_MASK_LOAD __ldcs
#endif
#ifndef GW_FED_U
#define GW_FED_U 4
#endif

constexpr int FED_U = GW_FED_U;         // 128-bit loads per lane and mask row in flight in the cooperative scan

__device__ __forceinline__ int popc4(uint4 v) { return __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w); }

// set bits among the first r bits (0 <= r <= 128) of a 128-bit group
__device__ __forceinline__ int popc_prefix(uint4 v, int r)
{
    const unsigned w[4] = {v.x, v.y, v.z, v.w};
    int n = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int b = r - 32 * j;
        const unsigned m = b >= 32 ? 0xFFFFFFFFu : (b <= 0 ? 0u : (0xFFFFFFFFu >> (32 - b)));
        n += __popc(w[j] & m);
    }
    return n;
}

// contribution of group g (content v) to the count of bits [k0, k1); q0 / q1 = first / last group
__device__ __forceinline__ int popc_group(uint4 v, int g, int q0, int q1, int k0, int k1)
{
    int n = popc4(v);
    if (g == q0) n -= popc_prefix(v, k0 - q0 * 128);
    if (g == q1) n -= popc4(v) - popc_prefix(v, k1 - q1 * 128);
    return n;
}

__device__ __forceinline__ int warp_popc_range(const uint32_t *row, int k0, int k1, int lane)
{
    int n = 0;
    if (k1 > k0) {
        const int q0 = k0 >> 7, q1 = (k1 - 1) >> 7;     // first / last 16-byte group
        const uint4 *row4 = reinterpret_cast<const uint4 *>(row);
        for (int q = q0 + lane; q <= q1; q += 128) {
            uint4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                v[u] = (q + 32 * u <= q1) ? GW_MASK_LOAD(row4 + q + 32 * u) : make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (q + 32 * u <= q1) n += popc_group(v[u], q + 32 * u, q0, q1, k0, k1);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xFFFFFFFFu, n, o);
    return n;
}

// the same count by ONE lane (every lane of a warp scanning a range of its own)
__device__ __noinline__ int lane_popc_range(const uint32_t *row, int k0, int k1)
{
    int n = 0;
    const int q0 = k0 >> 7, q1 = (k1 - 1) >> 7;
    const uint4 *row4 = reinterpret_cast<const uint4 *>(row);
    for (int q = q0; q <= q1; ++q) n += popc_group(__ldg(row4 + q), q, q0, q1, k0, k1);
    return n;
}

// set bits of the 128-bit group `v` at positions < r (any int r: <= 0 gives 0, >= 128 gives all): one
// BMSK (clamped bit-mask generate) + AND + POPC per word
__device__ __forceinline__ int popc_below(uint4 v, int r)
{
    const unsigned w[4] = {v.x, v.y, v.z, v.w};
    int n = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int b = max(r - 32 * j, 0);
        unsigned m;
        asm("bmsk.clamp.b32 %0, 0, %1;" : "=r"(m) : "r"(b));         // low min(b, 32) bits set
        n += __popc(w[j] & m);
    }
    return n;
}

// bits of group g (content v) inside the bit range [k0, k1): for the first / last group of a range
__device__ __forceinline__ int popc_edge(uint4 v, int g, int k0, int k1)
{
    return popc_below(v, k1 - g * 128) - popc_below(v, k0 - g * 128);
}

// the same count by ONE lane for a range of at most FED_TINY 16-byte groups (headers, announcement
// payloads: one or two groups): whole groups, minus the bits below k0 in the first group and the bits from
// k1 on in the last one
constexpr int FED_TINY = 4;
__device__ __forceinline__ int lane_popc_small(const uint32_t *row, int k0, int k1)
{
    const int q0 = k0 >> 7, q1 = (k1 - 1) >> 7;
    const uint4 *row4 = reinterpret_cast<const uint4 *>(row);
    const uint4 first = __ldg(row4 + q0);
    uint4 last = first;
    int n = popc4(first);
#pragma unroll 1
    for (int q = q0 + 1; q <= q1; ++q) { last = __ldg(row4 + q); n += popc4(last); }
    n -= popc_below(first, k0 - q0 * 128);
    n -= popc4(last) - popc_below(last, k1 - q1 * 128);
    return n;
}

// fire-and-forget request to bring `bytes` (a multiple of 16) at `p` (16-byte aligned) into L2: one
// instruction per mask row (SASS UBLKPF), no registers or shared memory held while the row is in flight
__device__ __forceinline__ void prefetch_bulk_l2(const void *p, unsigned bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(__cvta_generic_to_global(p)), "r"(bytes) : "memory");
}

// ------------------------------------------------------------------------------------
// K3i: prefix-count index of the fed masks.
//
// A section's error count is the number of set bits among bits [k0, k1) of a mask row (SimplePhy._countBitErrors
// with per-bit masks, simple_stack.py:180-188).  The rows are DATA handed over with gw_set_masks, so the counting
// is hoisted out of the event loop: ONE streaming pass over the whole mask buffer -- this kernel, HBM-bound,
// every mask word read exactly once -- leaves per row the number of set bits in front of every 128-bit group
// (uint16, relative to the group's 512-group superblock; uint32 totals per superblock for rows longer than
// 64 Kibit), and a count in the step kernel is  before(k1) - before(k0)  with
//     before(k) = s32[row][k >> 16] + t16[row][k >> 7] + popc(bits of group k >> 7 below k & 127):
// two index entries and at most two 16-byte groups per row and decision, whatever the length of the section
// and however the section was cut into SINR segments.  One warp per row, four consecutive groups per lane
// (64 contiguous bytes), one warp scan per 128 groups.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
mask_index_kernel(const uint32_t *words, int words_per_row, long long rows, uint16_t *t16, uint32_t *s32, int gt, int nsb)
{
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int G = words_per_row >> 2;
    for (long long r = warp; r < rows; r += nwarps) {
        const uint4 *row4 = reinterpret_cast<const uint4 *>(words + r * words_per_row);
        uint16_t *trow = t16 + r * gt;
        unsigned total = 0, sbBase = 0;
        for (int start = 0; start <= G; start += 128) {
            const int g = start + 4 * lane;
            uint4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = g + u < G ? __ldcs(row4 + g + u) : make_uint4(0, 0, 0, 0);
            const int c0 = popc4(v[0]), c1 = popc4(v[1]), c2 = popc4(v[2]), c3 = popc4(v[3]);
            const int mine = c0 + c1 + c2 + c3;
            int incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if (lane >= o) incl += t;
            }
            if ((start & 511) == 0) {
                sbBase = total;
                if (s32 != nullptr && lane == 0) s32[r * nsb + (start >> 9)] = total;
            }
            const unsigned b0 = total + (unsigned)(incl - mine) - sbBase;      // set bits of this superblock in front of group g
            if (g < gt) {
                const unsigned e0 = b0, e1 = b0 + c0, e2 = e1 + c1, e3 = e2 + c2;
                *reinterpret_cast<uint2 *>(trow + g) = make_uint2(e0 | (e1 << 16), e2 | (e3 << 16));
            }
            total += (unsigned)__shfl_sync(0xFFFFFFFFu, incl, 31);
        }
    }
}

// set bits among bits [k0, k1) of mask row `row` = before(k1) - before(k0): the (up to) two index entries per end
// and the (up to) two edge groups are independent read-only loads, issued together
__device__ __forceinline__ int fed_count(const MaskSource &m, long long row, int k0, int k1)
{
    const int g0 = k0 >> 7, r0 = k0 & 127, g1 = k1 >> 7, r1 = k1 & 127;
    const uint16_t *t = m.t16 + row * m.gt;
    const uint4 *row4 = reinterpret_cast<const uint4 *>(m.words + row * m.words_per_row);
    const int t1 = __ldg(t + g1);
    const int t0 = k0 ? (int)__ldg(t + g0) : 0;
    uint4 v1 = make_uint4(0, 0, 0, 0), v0 = make_uint4(0, 0, 0, 0);
    if (r1) v1 = __ldg(row4 + g1);
    if (r0) v0 = __ldg(row4 + g0);
    int c = t1 - t0;
    if (m.s32 != nullptr) c += (int)(__ldg(m.s32 + row * m.nsb + (g1 >> 9)) - __ldg(m.s32 + row * m.nsb + (g0 >> 9)));
    return c + popc_below(v1, r1) - popc_below(v0, r0);
}

// the index look-up on its own (gw_mask_index_count: numeric tests of the index against plain popcounts)
__global__ void mask_index_count_kernel(MaskSource m, const long long *rows, const int *k0, const int *k1, int *counts, long long n)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    counts[i] = k1[i] > k0[i] ? fed_count(m, rows[i], k0[i], k1[i]) : 0;
}

// Philox-generated masks: bit k is an error iff word (k & 3) of block (k >> 2) < thr
__device__ __forceinline__ int warp_philox_range(unsigned long long seed, long long env, int band, int sender,
                                                 uint32_t txseq, int receiver, int k0, int k1, uint32_t thr, int lane)
{
    int n = 0;
    if (k1 > k0) {
        const int b0 = k0 >> 2, b1 = (k1 - 1) >> 2;
        for (int blk = b0 + lane; blk <= b1; blk += 32) {
            uint32_t w[4];
            mask_words4(seed, env, band, sender, txseq, receiver, (uint32_t)blk, w);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = blk * 4 + j;
                n += (k >= k0 && k < k1 && w[j] < thr) ? 1 : 0;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xFFFFFFFFu, n, o);
    return n;
}

// ------------------------------------------------------------------------------------
// K4: fused event-ordered step kernel (+ K5 epilogue)
// ------------------------------------------------------------------------------------

// kernel-level variant of MODE_M_FED: the handle holds a prefix-count index of its masks (mask_index_kernel)
constexpr int MODE_M_FEDX = 3;

struct StepArgs {
    StatePtrs st;
    const int32_t *device;
    const int32_t *duration;
    long long *obs;
    double *reward;
    unsigned char *done;
    double *stats;
    int *errflag;
    unsigned long long *maskBytes;      // mode M fed: bytes of mask words the step kernels scanned (statistic)
    unsigned long long *stamps;         // diagnostics (gw_debug_stamps): this launch's {first block start, first block
                                        // past the grid dependency, last block end} in globaltimer ns, or NULL
    MaskSource masks;
    DevMemo memo;
    // compact outputs (gw_step_host_packed): used instead of obs / reward when non-NULL
    int *obs32;
    float *reward32;
    // gw_step_host_compact: uint8 actions [n][2] in, one packed uint32 per sim out
    const unsigned char *act8;
    unsigned *res32;
    int tiny;                           // gw_step_host_tiny: act8 = uint8 [n] (device << 7 | duration), res32 = uint16 [n]
    // band-sims [sim_begin, sim_end) are stepped by this launch (a multiple of the block size apart)
    long long sim_begin, sim_end;
    // event trace (gw_step_traced)
    double *trace;
    int *traceCount;
    int traceCap;
};

struct SharedTables {
    double srx[kMaxBands][16];
};

#ifndef GW_STEP_MIN_BLOCKS
#define GW_STEP_MIN_BLOCKS 4
#endif

template <int MODE, int D, int NS, int NJ, bool TRACE = false, bool EXT = false>
__global__ void __launch_bounds__(STEP_BLOCK, GW_STEP_MIN_BLOCKS)
step_kernel(const __grid_constant__ StepArgs A, const __grid_constant__ Params P,
            const __grid_constant__ SharedTables T)
{
    using SimT = Sim<D, NS, NJ, ShStore>;
    const int lane = threadIdx.x & 31;
    const int nb = P.nbands;
    const long long nsim = A.st.nsim;
    int acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0;
    const bool keepSeq = MODE != MODE_R || A.st.per_env != 0;     // see load_sim

    // The first row of this thread is requested into L2 right away: the block-level set-up below (and,
    // under programmatic dependent launch, the tail of the previous kernel) then overlaps the DRAM
    // latency of the state loads.  L2 is the point of coherence, so prefetching before the grid
    // dependency is resolved is safe.
    {
        const long long i0 = A.sim_begin + (long long)blockIdx.x * blockDim.x + threadIdx.x;
        if (i0 < A.sim_end) {
#pragma unroll
            for (int c = 0; c < HOT_CHUNKS; ++c)
                if (c != H_EPC && (NJ > 0 || c != H_JAM) && (keepSeq || c != H_U2)) prefetch_l2(A.st.hot + (long long)c * nsim + i0);
            if (nb == 1) prefetch_l2(A.st.now + i0);
            if (A.act8) prefetch_l2(A.act8 + (A.tiny ? i0 : 2 * i0));
            else { prefetch_l2(A.device + i0); prefetch_l2(A.duration + i0); }
        }
    }
    if (A.stamps != nullptr && threadIdx.x == 0) atomicMin(A.stamps + 0, globaltimer_ns());
    // first level of the BER memo: copied into shared memory while the state loads are in flight.
    // (Read before the grid dependency is resolved: entries are checksum-validated, a stale or
    // torn one is a miss.)
    __shared__ ulonglong2 memo_l0[2 * MEMO_L0];
    DevMemo memo = A.memo;
    memo.sim = nullptr;
    if (memo.l0g != nullptr) {
        for (int t = threadIdx.x; t < 2 * MEMO_L0; t += blockDim.x) memo_l0[t] = memo.l0g[t];
        memo.l0s = memo_l0;
    }
    __shared__ double srx_s[kMaxBands][16];
    __shared__ double p_loaded[NJ == 0 ? 3 : 1][NJ == 0 ? STEP_BLOCK : 1];     // see store_sim / keep_p
    if (threadIdx.x < kMaxBands * 16) srx_s[threadIdx.x >> 4][threadIdx.x & 15] = T.srx[threadIdx.x >> 4][threadIdx.x & 15];
    // Programmatic dependent launch: this grid may have been scheduled while the previous kernel
    // of the stream (the previous step, or whatever produced the actions) was still draining; its
    // memory is visible from here on.  The next kernel may start launching right away -- it waits
    // at the same point for this grid to complete.
    cudaGridDependencySynchronize();
    cudaTriggerProgrammaticLaunchCompletion();
    if (A.stamps != nullptr && threadIdx.x == 0) atomicMin(A.stamps + 1, globaltimer_ns());
    __syncthreads();

    // grid-stride over warps' worth of band-sims; every lane of a warp stays in the loop
    // so that the warp-level operations below are executed by all 32 lanes
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long base = A.sim_begin + (long long)blockIdx.x * blockDim.x; base < A.sim_end; base += stride) {
        const long long i = base + threadIdx.x;
        const bool active = i < A.sim_end;
        // nb is 1, 2 or 4: shifts instead of 64-bit divisions
        const int nbShift = nb == 4 ? 2 : (nb == 2 ? 1 : 0);
        const long long env = active ? (i >> nbShift) : 0;
        const int band = active ? (int)(i & (nb - 1)) : 0;

        SimT s;
        const BandParams &B = P.band[band];
        // received-power table: the block's shared copy (one geometry for all envs) or this
        // band-sim's own table in global memory; no registers are held for it
        const SrxView srx = A.st.per_env ? SrxView{A.st.srx + (active ? i : 0), A.st.ntab}
                                         : SrxView{&srx_s[band][0], 1};
        DevRingT<EXT> ring{A.st.ring + (active ? i : 0), nsim, A.st.hot + (active ? i : 0)};
        // (compile-time off in the kernels of bands without interferers: the default kernel sits at its register
        // limit, and the extra pointer + branch cost its productive regime 7 %)
        if (NJ > 0 && A.memo.sim != nullptr) memo.sim = A.memo.sim + (active ? i : 0);
        int dev = 0, dur = 0;
        bool idle0 = false;
        if (active) {
            idle0 = !load_sim<false, MODE != MODE_R>(s, A.st, i, A.st.now[env], keepSeq);
            if (NJ == 0) {
#pragma unroll
                for (int p = 0; p < 3; ++p) p_loaded[p][threadIdx.x] = s.P[p];
            }
            if (base != A.sim_begin + (long long)blockIdx.x * blockDim.x) {     // later rounds of the grid-stride loop
                prefetch_l2(A.st.hot + (long long)H_EP * nsim + i);
            }
            if (TRACE) { s.trace = A.trace + (long long)i * A.traceCap * 8; s.traceCap = A.traceCap; s.ntrace = 0; }
            if (A.act8) {
                if (A.tiny) {
                    const unsigned a1 = A.act8[i];
                    dev = (int)(a1 >> 7); dur = (int)(a1 & 127u);
                } else {
                    const uchar2 a2 = reinterpret_cast<const uchar2 *>(A.act8)[i];
                    dev = a2.x; dur = a2.y;
                }
            } else {
                dev = A.device[i];
                dur = A.duration[i];
            }
            // assert self.action_space.contains(action)  (counter_traffic.py:147)
            if (dev < 0 || dev >= NS || dur < 0 || dur >= P.maxDuration) {
                if (atomicCAS(A.errflag, 0, GW_E_ACTION) == 0) A.errflag[1] = (int)i;
                dev = dev < 0 ? 0 : (dev >= NS ? NS - 1 : dev);
                dur = dur < 0 ? 0 : (dur >= P.maxDuration ? P.maxDuration - 1 : dur);
            }
            begin_assignment(s, P, dev, dur);
        } else {
            s.assignDone = 1; s.fault = 0; s.now = 0;
            s.nTx = 0;
        }
        uint32_t nTx0 = 0, nD0 = 0, nD1 = 0;
        if (active) { nTx0 = s.nTx; nD0 = s.nDeliv[0]; nD1 = s.nDeliv[1]; }
        const uint32_t ties0 = 0;           // the lean load starts the per-step tie counter at 0

        if (MODE == MODE_R) {
            if (active) run_until_assign<MODE_R>(s, P, B, srx, ring, NoMasks(), memo, idle0 ? 1 : 0);
            if (nb > 1) {
                // SimMan.runSimulation for every band's ASSIGN message: the env's clock ends
                // at the latest band; the other bands keep simulating up to that time
                double Tend = active ? s.now : -INFINITY;
                for (int o = 1; o < nb; o <<= 1) Tend = fmax(Tend, __shfl_xor_sync(0xFFFFFFFFu, Tend, o));
                if (active && s.now < Tend) run_until_time<MODE_R>(s, P, B, srx, ring, NoMasks(), Tend, memo);
            }
        } else if (MODE == MODE_M_FEDX) {
            // Mode M with fed masks and a prefix-count index (mask_index_kernel): the count of a decision is two
            // index look-ups per receiver, done by the deciding lane on the spot -- the per-lane event loop of
            // mode R, no warp-level service of count requests.  (fed_decide_set: only DECIDING events count.)
            double Tend = INFINITY;
            int phase = 0;      // 0: until the own ASSIGN is processed, 1: until Tend
            unsigned maskWords = 0;
            for (;;) {
                if (active) {
                    while (!s.fault) {
                        if (phase == 0 && s.assignDone) break;
                        const Event ev = next_event(s, B, phase == 0 ? (double)INFINITY : Tend, ring);
                        if (phase == 1 && !(ev.t < Tend)) break;
                        const int nd = fed_decide_set(s, ev);
                        s.now = ev.t;
                        if (nd) {
                            // all receivers' look-ups are issued before the first count is used
                            int cnt[D];
#pragma unroll
                            for (int p = 0; p < D; ++p) {
                                cnt[p] = 0;
                                if (!((nd >> p) & 1)) continue;
                                int sender; uint32_t txseq; int64_t a0, a1;
                                mask_range(s, p, P.bitRate, sender, txseq, a0, a1);
                                if (a1 > a0) {
                                    const long long row = ((((env * nb + band) * kMaxDev + sender) * A.masks.slots
                                                            + (long long)(txseq % (uint32_t)A.masks.slots)) * kMaxDev + p);
                                    cnt[p] = fed_count(A.masks, row, (int)a0, (int)a1);
                                    maskWords += (unsigned)((((int)a1 + 31) >> 5) - ((int)a0 >> 5));
                                }
                            }
#pragma unroll
                            for (int p = 0; p < D; ++p)
                                if ((nd >> p) & 1) { s.err[p] += (double)cnt[p]; s.segT0[p] = ev.t; }
                        }
                        const int berMask = apply_event(s, P, B, ev, srx, ring);
                        update_bers(s, P, berMask, srx, memo);
                    }
                }
                if (phase == 0 && nb > 1) {
                    __syncwarp();
                    double t = active ? s.now : -INFINITY;
                    for (int o = 1; o < nb; o <<= 1) t = fmax(t, __shfl_xor_sync(0xFFFFFFFFu, t, o));
                    Tend = t;
                    phase = 1;
                    continue;
                }
                break;
            }
            if (active && nb > 1) s.now = Tend;
            // statistic: the mask words that hold the on-air bits of the decided sections (algorithmic bytes)
            maskWords = __reduce_add_sync(0xFFFFFFFFu, maskWords);
            if (lane == 0 && maskWords != 0 && A.maskBytes != nullptr) atomicAdd(A.maskBytes, 4ull * maskWords);
        } else if (MODE == MODE_M_FED) {
            // Mode M with fed masks.  The count of a section does not depend on the segmentation
            // (gw_core.cuh::fed_decide_set), so only DECIDING events read mask words.  Two alternating parts:
            //  (1) serial: every lane runs ahead, at its own pace, through its events (the cheap per-lane
            //      loop of mode R).  Decisions over a few 16-byte groups (headers, announcements) are
            //      counted on the spot by their own lane; a lane stops in front of a decision over a LONG
            //      range;
            //  (2) warp-uniform: the pending long ranges of all 32 lanes are counted together, four rows
            //      at a time: one row per group of eight lanes, each lane reading every eighth 16-byte
            //      group (whole 128-byte lines per octet and load instruction, FED_U loads per lane in
            //      flight), popc, reduction within the eight lanes.  The event itself is applied at the
            //      head of part (1): the transition function has ONE call site (code size).
            // (Measured alternatives, profiles/README.md: the round-1 lockstep loop, 32 lanes per row, L2
            // prefetch passes -- bulk and per line --, and a vote on the most common event kind were all slower.)
            double Tend = INFINITY;
            int phase = 0;      // 0: until the own ASSIGN is processed, 1: until Tend
            const long long wpr = A.masks.words_per_row;
            unsigned maskWords = 0;
            Event ev; ev.kind = EV_NONE; ev.idx = 0; ev.t = 0; ev.seq = 0;
            int need = 0, k0c = 0, k1c = 0;
            long long rowbase = 0;
            int cnt[D];
#pragma unroll
            for (int p = 0; p < D; ++p) cnt[p] = 0;
            for (;;) {
                if (active) {
                    while (!s.fault) {
                        int nd = need;                  // a pending long decision: its counts have arrived
                        need = 0;
                        if (!nd) {
                            if (phase == 0 && s.assignDone) break;
                            ev = next_event(s, B, phase == 0 ? (double)INFINITY : Tend, ring);
                            if (phase == 1 && !(ev.t < Tend)) break;
                            nd = fed_decide_set(s, ev);
                            if (nd) {
                                s.now = ev.t;
                                int sender = 0; uint32_t txseq = 0;
                                bool first = true, uniform = true;
#pragma unroll
                                for (int p = 0; p < D; ++p) {
                                    if (!((nd >> p) & 1)) continue;
                                    int64_t a0, a1;
                                    mask_range(s, p, P.bitRate, sender, txseq, a0, a1);
                                    if (first) { k0c = (int)a0; k1c = (int)a1; first = false; }
                                    else uniform &= ((int)a0 == k0c) & ((int)a1 == k1c);
                                }
                                rowbase = ((((env * nb + band) * kMaxDev + sender) * A.masks.slots
                                            + (long long)(txseq % (uint32_t)A.masks.slots)) * kMaxDev) * wpr;
#pragma unroll
                                for (int p = 0; p < D; ++p) cnt[p] = 0;
#ifndef GW_PROBE_NOSCAN
                                const int groups = k1c > k0c ? ((k1c - 1) >> 7) - (k0c >> 7) + 1 : 0;
                                if (!uniform) {
                                    // the receivers' segments differ (a position change cut one of them): rare
#pragma unroll 1
                                    for (int p = 0; p < D; ++p) {
                                        if (!((nd >> p) & 1)) continue;
                                        int sd; uint32_t tq; int64_t a0, a1;
                                        mask_range(s, p, P.bitRate, sd, tq, a0, a1);
                                        if (a1 > a0) {
                                            const int c = lane_popc_range(A.masks.words + rowbase + p * wpr, (int)a0, (int)a1);
#pragma unroll
                                            for (int pp = 0; pp < D; ++pp) cnt[pp] = pp == p ? c : cnt[pp];
                                            maskWords += (unsigned)((((int)a1 + 31) >> 5) - ((int)a0 >> 5));
                                        }
                                    }
                                } else if (groups > 0) {
                                    maskWords += (unsigned)__popc(nd) * (unsigned)(((k1c + 31) >> 5) - (k0c >> 5));
                                    if (groups <= FED_TINY) {
#pragma unroll
                                        for (int p = 0; p < D; ++p)
                                            if ((nd >> p) & 1) cnt[p] = lane_popc_small(A.masks.words + rowbase + p * wpr, k0c, k1c);
                                    } else {
                                        need = nd;      // counted by the whole warp below
                                        break;
                                    }
                                }
#endif
                            }
                        }
                        if (nd) {
#pragma unroll
                            for (int p = 0; p < D; ++p)
                                if ((nd >> p) & 1) { s.err[p] += (double)cnt[p]; s.segT0[p] = ev.t; }
                        }
                        s.now = ev.t;
                        const int berMask = apply_event(s, P, B, ev, srx, ring);
                        update_bers(s, P, berMask, srx, memo);
                    }
                    if (s.fault) need = 0;
                }
                const unsigned pend = __ballot_sync(0xFFFFFFFFu, need != 0);
                if (pend == 0) {
                    if (phase == 0 && nb > 1) {
                        double t = active ? s.now : -INFINITY;
                        for (int o = 1; o < nb; o <<= 1) t = fmax(t, __shfl_xor_sync(0xFFFFFFFFu, t, o));
                        Tend = t;
                        phase = 1;
                        continue;
                    }
                    break;
                }
                // ---- (2) the long ranges of the pending decisions, four rows at a time
                const int oct = lane >> 3, sub = lane & 7;
#pragma unroll 1
                for (int p = 0; p < D; ++p) {
                    unsigned M = __ballot_sync(0xFFFFFFFFu, (need >> p) & 1);
                    while (M) {
                        const int s0 = __ffs(M) - 1; M &= M - 1;
                        const int s1 = __ffs(M) - 1; M &= M - 1;        // -1 when M ran empty (0 & x stays 0)
                        const int s2 = __ffs(M) - 1; M &= M - 1;
                        const int s3 = __ffs(M) - 1; M &= M - 1;
                        const int src = oct == 0 ? s0 : (oct == 1 ? s1 : (oct == 2 ? s2 : s3));
                        const int from = src < 0 ? 0 : src;
                        const long long rb = __shfl_sync(0xFFFFFFFFu, rowbase, from);
                        const int a0 = __shfl_sync(0xFFFFFFFFu, k0c, from), a1 = __shfl_sync(0xFFFFFFFFu, k1c, from);
                        const int q0 = a0 >> 7, q1 = src < 0 ? -1 : ((a1 - 1) >> 7);
                        const uint4 *row4 = reinterpret_cast<const uint4 *>(A.masks.words + rb + p * wpr);
                        int tot = 0;
#pragma unroll 1
                        for (int gb = q0 + sub; gb <= q1; gb += 8 * FED_U) {
                            uint4 v[FED_U];
#pragma unroll
                            for (int u = 0; u < FED_U; ++u)
                                v[u] = gb + 8 * u <= q1 ? GW_MASK_LOAD(row4 + gb + 8 * u) : make_uint4(0, 0, 0, 0);
#pragma unroll
                            for (int u = 0; u < FED_U; ++u) {
                                const int g = gb + 8 * u;
                                if (g == q0 || g == q1) {
                                    asm volatile("");           // keep this a (rare, two lanes per row) branch
                                    tot += popc_edge(v[u], g, a0, a1);
                                } else {
                                    tot += popc4(v[u]);
                                }
                            }
                        }
                        tot += __shfl_xor_sync(0xFFFFFFFFu, tot, 4);
                        tot += __shfl_xor_sync(0xFFFFFFFFu, tot, 2);
                        tot += __shfl_xor_sync(0xFFFFFFFFu, tot, 1);
                        const int rank = lane == s0 ? 0 : (lane == s1 ? 1 : (lane == s2 ? 2 : (lane == s3 ? 3 : -1)));
                        const int val = __shfl_sync(0xFFFFFFFFu, tot, rank < 0 ? 0 : 8 * rank);
                        if (rank >= 0) {
#pragma unroll
                            for (int pp = 0; pp < D; ++pp) cnt[pp] = pp == p ? val : cnt[pp];
                        }
                    }
                }
            }
            if (active && nb > 1) s.now = Tend;
            // statistic: mask bytes scanned by this block
            maskWords = __reduce_add_sync(0xFFFFFFFFu, maskWords);
            if (lane == 0 && maskWords != 0 && A.maskBytes != nullptr) atomicAdd(A.maskBytes, 4ull * maskWords);
        } else {
            // warp-synchronous event loop: one timed event per lane and iteration; the error
            // counts of all lanes are serviced cooperatively (ballot / popc / shuffle)
            double Tend = INFINITY;
            int phase = 0;      // 0: until the own ASSIGN is processed, 1: until Tend
            for (;;) {
                Event ev; ev.kind = EV_NONE; ev.idx = 0; ev.t = 0; ev.seq = 0;
                bool run = active && !s.fault;
                if (run) {
                    if (phase == 0) run = !s.assignDone;
                    if (run || phase == 1) {
                        ev = next_event(s, B, phase == 0 ? (double)INFINITY : Tend, ring);
                        run = phase == 0 ? true : (ev.t < Tend);
                    }
                }
                const unsigned running = __ballot_sync(0xFFFFFFFFu, run);
                if (running == 0) {
                    if (phase == 0 && nb > 1) {
                        double t = active ? s.now : -INFINITY;
                        for (int o = 1; o < nb; o <<= 1) t = fmax(t, __shfl_xor_sync(0xFFFFFFFFu, t, o));
                        Tend = t;
                        phase = 1;
                        continue;
                    }
                    break;
                }
                int once = 0, twice = 0;
                if (run) {
                    s.now = ev.t;
                    count_set(s, ev, srx, once, twice);
                }
                // Service the count requests.  Two ways: (a) cooperatively -- for every lane with requests
                // all 32 lanes scan that lane's range (ballot / shuffle / popc, coalesced 512-byte reads):
                // right for few, long ranges; (b) every lane scans its own ranges: right when many lanes
                // have requests at once, which is the rule in this lockstep loop.  The warp picks the
                // cheaper one from an instruction estimate (warp-uniform, so (a) stays converged).
                int costLocal = 0, costCoop = 0;
                if (once != 0) {
                    int req = once;
                    while (req) {
                        const int p = __ffs(req) - 1;
                        req &= req - 1;
                        int sender; uint32_t txseq; int64_t k0, k1;
                        mask_range(s, p, P.bitRate, sender, txseq, k0, k1);
                        if (k1 > k0) {
                            if (MODE == MODE_M_FED) {
                                const int groups = (int)(((k1 - 1) >> 7) - (k0 >> 7)) + 1;       // 16-byte groups
                                costLocal += 14 * groups + 20;
                                costCoop += 45 * ((groups + 127) >> 7) + 50;
                            } else {
                                const int blocks = (int)(((k1 - 1) >> 2) - (k0 >> 2)) + 1;       // Philox blocks
                                costLocal += 120 * blocks + 20;
                                costCoop += 120 * ((blocks + 31) >> 5) + 50;
                            }
                        } else {
                            costLocal += 10; costCoop += 50;
                        }
                    }
                }
                const bool laneLocal = __reduce_max_sync(0xFFFFFFFFu, costLocal) <= __reduce_add_sync(0xFFFFFFFFu, costCoop);
                if (laneLocal) {
                    int req = once;
                    while (req) {
                        const int p = __ffs(req) - 1;
                        req &= req - 1;
                        int sender; uint32_t txseq; int64_t k0, k1;
                        mask_range(s, p, P.bitRate, sender, txseq, k0, k1);
                        long long cnt = 0;
                        if (k1 > k0) {
                            if (MODE == MODE_M_FED) {
                                const long long row = ((((env * nb + band) * kMaxDev + sender) * A.masks.slots
                                                        + (long long)(txseq % (uint32_t)A.masks.slots)) * kMaxDev + p);
#ifdef GW_PROBE_NOSCAN
                                cnt = 0; (void)row;
#else
                                cnt = lane_popc_range(A.masks.words + row * A.masks.words_per_row, (int)k0, (int)k1);
#endif
                            } else {
                                cnt = mask_errors_serial(A.masks.seed, A.masks.env_offset + env, band, sender, txseq, p,
                                                         k0, k1, get_at(s.ber, p));
                            }
                        }
                        set_at(s.err, p, get_at(s.err, p) + (double)cnt);
                        set_at(s.segT0, p, s.now);
                    }
                }
                unsigned pending = laneLocal ? 0u : __ballot_sync(0xFFFFFFFFu, once != 0);
                while (pending) {
                    const int src = __ffs(pending) - 1;
                    pending &= pending - 1;
                    int req = __shfl_sync(0xFFFFFFFFu, once, src);
                    while (req) {
                        const int p = __ffs(req) - 1;
                        req &= req - 1;
                        int sender = 0; uint32_t txseq = 0; int64_t k0 = 0, k1 = 0; double berp = 0;
                        if (lane == src) {
                            mask_range(s, p, P.bitRate, sender, txseq, k0, k1);
                            berp = get_at(s.ber, p);
                        }
                        sender = __shfl_sync(0xFFFFFFFFu, sender, src);
                        txseq = __shfl_sync(0xFFFFFFFFu, txseq, src);
                        const int ik0 = __shfl_sync(0xFFFFFFFFu, (int)k0, src);
                        const int ik1 = __shfl_sync(0xFFFFFFFFu, (int)k1, src);
                        const long long simi = __shfl_sync(0xFFFFFFFFu, i, src);
                        const long long senv = simi / nb;
                        const int sband = (int)(simi - senv * nb);
                        int cnt;
                        if (MODE == MODE_M_FED) {
                            const long long row = ((((senv * nb + sband) * kMaxDev + sender) * A.masks.slots
                                                    + (long long)(txseq % (uint32_t)A.masks.slots)) * kMaxDev + p);
#ifdef GW_PROBE_NOSCAN
                            cnt = 0; (void)row;
#else
                            cnt = warp_popc_range(A.masks.words + row * A.masks.words_per_row, ik0, ik1, lane);
#endif
                        } else {
                            const uint32_t thr = ber_threshold(__shfl_sync(0xFFFFFFFFu, berp, src));
                            cnt = warp_philox_range(A.masks.seed, A.masks.env_offset + senv, sband, sender, txseq, p,
                                                    ik0, ik1, thr, lane);
                        }
                        if (lane == src) {
                            set_at(s.err, p, get_at(s.err, p) + (double)cnt);
                            set_at(s.segT0, p, s.now);
                        }
                    }
                }
                if (run) {
                    const int berMask = apply_event(s, P, B, ev, srx, ring);
                    update_bers(s, P, berMask, srx, memo);
                }
            }
            if (active && nb > 1) s.now = Tend;
        }

        if (active) {
            long long o; double rw; unsigned char dn;
            feedback(s, o, rw, dn);
            if (A.res32) {
                if (A.tiny) {
                    // obs as the signed difference obs - COUNTER_BOUND (the interpreter's latestDifference,
                    // counter_traffic.py:85-94: +-COUNTER_BYTE_LENGTH at most); bit 15: it did not fit 8 bits
                    const int diff = (int)(o - kCounterBound);
                    reinterpret_cast<unsigned short *>(A.res32)[i] = (unsigned short)(
                        ((unsigned)diff & 0xFFu) | ((unsigned)((int)rw + 16) << 8) | ((unsigned)(dn != 0) << 13)
                        | ((diff < -128 || diff > 127) ? 0x8000u : 0u));
                } else {
                    A.res32[i] = ((unsigned)o & 0x1FFFFu) | ((unsigned)((int)rw + 16) << 17) | ((unsigned)(dn != 0) << 22);
                }
            } else {
                if (A.obs32) { A.obs32[i] = (int)o; A.reward32[i] = (float)rw; }
                else { A.obs[i] = o; A.reward[i] = rw; }
                A.done[i] = dn;
            }
            if (band == 0) A.st.now[env] = s.now;
            if (s.fault) { if (atomicCAS(A.errflag, 0, GW_E_SIMFAULT) == 0) { A.errflag[1] = (int)i; A.errflag[2] = s.fault; } }
            int keepP = 0;
            if (NJ == 0) {
                // bitwise comparison with the loaded values (kept in shared memory: no registers held)
                const bool e0 = __double_as_longlong(s.P[0]) == __double_as_longlong(p_loaded[0][threadIdx.x]);
                const bool e1 = __double_as_longlong(s.P[1]) == __double_as_longlong(p_loaded[1][threadIdx.x]);
                const bool e2 = __double_as_longlong(s.P[2]) == __double_as_longlong(p_loaded[2][threadIdx.x]);
                keepP = ((e0 && e1) ? 1 : 0) | (e2 ? 2 : 0);
            }
            store_sim<false, MODE != MODE_R>(s, A.st, i, false, keepSeq, keepP);
            if (TRACE) A.traceCount[i] = s.ntrace;
            acc[0] += (int)rw;                  // rewards are integers in [-10, 10] (counter_traffic.py:96-107)
            acc[1] += (int)(s.nDeliv[0] - nD0);
            acc[2] += (int)(s.nDeliv[1] - nD1);
            acc[3] += (int)dn;
            acc[4] += 1;
            acc[5] += s.latestDiff < 0 ? -s.latestDiff : s.latestDiff;
            acc[6] += (int)(s.nTx - nTx0);
            acc[7] += (int)(s.ties - ties0);
        }
    }

    // K5: warp reduction (REDUX) -> shared memory -> one atomic per block and statistic.  The
    // per-block partial sums are small integers, so the fp64 totals are exact and order-independent.
    __shared__ int red[STEP_BLOCK / 32][8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int v = __reduce_add_sync(0xFFFFFFFFu, acc[k]);
        if (lane == 0) red[threadIdx.x >> 5][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < 8) {
        int v = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[w][threadIdx.x];
        if (v != 0) atomicAdd(A.stats + threadIdx.x, (double)v);
    }
    if (A.stamps != nullptr && threadIdx.x == 0) atomicMax(A.stamps + 2, globaltimer_ns());
}

// ------------------------------------------------------------------------------------
// config 5: networked inverted pendulum -- the same transition function with a plant plugged in
// ------------------------------------------------------------------------------------

#ifndef GW_PEND_MIN_BLOCKS
#define GW_PEND_MIN_BLOCKS 4
#endif

struct DevVals {
    double *base;       // pval + sim index
    long long nsim;
    __device__ __forceinline__ void put(int k, uint32_t slot, double v) { base[(long long)(k * kRingSlots + slot) * nsim] = v; }
    __device__ __forceinline__ double get(int k, uint32_t slot) const { return base[(long long)(k * kRingSlots + slot) * nsim]; }
};
struct DevSrxOut {
    double *base;       // srx table + sim index (per-sim tables)
    long long ntab;
    __device__ __forceinline__ void put(int k, double v) { base[(long long)k * ntab] = v; }
};

__device__ __forceinline__ void load_plant(PendulumState &S, const StatePtrs &st, long long i)
{
    const long long n = st.nsim;
    S.x = st.plant[0 * n + i]; S.v = st.plant[1 * n + i]; S.th = st.plant[2 * n + i]; S.om = st.plant[3 * n + i];
    S.vTarget = st.plant[4 * n + i]; S.tPlant = st.plant[5 * n + i]; S.ctrlAngleDeg = st.plant[6 * n + i];
    S.lastError = st.plant[7 * n + i];
}
__device__ __forceinline__ void store_plant(const PendulumState &S, const StatePtrs &st, long long i)
{
    const long long n = st.nsim;
    st.plant[0 * n + i] = S.x; st.plant[1 * n + i] = S.v; st.plant[2 * n + i] = S.th; st.plant[3 * n + i] = S.om;
    st.plant[4 * n + i] = S.vTarget; st.plant[5 * n + i] = S.tPlant; st.plant[6 * n + i] = S.ctrlAngleDeg;
    st.plant[7 * n + i] = S.lastError;
}

// Same layout as the step kernels: the per-device arrays of the band-sim in dynamic shared memory
// ([field][index][thread], rolled device loops), the received-power table read from / written to the band-sim's
// own table in global memory (the links follow the wagon).
__global__ void __launch_bounds__(STEP_BLOCK, GW_PEND_MIN_BLOCKS)
pendulum_step_kernel(const __grid_constant__ StepArgs A, const __grid_constant__ Params P,
                     const __grid_constant__ PendulumParams Q)
{
    using SimT = Sim<4, 2, 1, ShStore>;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nsim = A.st.nsim;
    if (i >= nsim) return;
    SimT s;
    PendulumState S;
    const SrxView srx{A.st.srx + i, A.st.ntab};
    load_sim(s, A.st, i, A.st.now[i]);
    load_plant(S, A.st, i);
    int dev = A.device[i], dur = A.duration[i];
    if (dev < 0 || dev >= 2 || dur < 0 || dur >= P.maxDuration) {
        if (atomicCAS(A.errflag, 0, GW_E_ACTION) == 0) A.errflag[1] = (int)i;
        dev = dev < 0 ? 0 : (dev >= 2 ? 1 : dev);
        dur = dur < 0 ? 0 : (dur >= P.maxDuration ? P.maxDuration - 1 : dur);
    }
    const uint32_t nTx0 = s.nTx, nD0 = s.nDeliv[0], nD1 = s.nDeliv[1];
    begin_assignment(s, P, dev, dur);
    DevRingT<false> ring{A.st.ring + i, nsim, A.st.hot + i};
    PendulumPlant<DevVals, DevSrxOut> plant(Q, S, DevVals{A.st.pval + i, nsim}, DevSrxOut{A.st.srx + i, A.st.ntab});
    run_until_assign_plant<MODE_R>(s, P, P.band[0], srx, ring, NoMasks(), NoMemo(), plant);
    // InvertedPendulumInterpreter (inverted_pendulum.py:42-56): the angle is read from the plant
    pendulum_advance(Q, S, s.now);
    const double deg = S.th * (180.0 / 3.141592653589793);
    const double rw = fabs(180.0 - deg);
    if (A.obs32) { A.obs32[i] = (int)deg; A.reward32[i] = (float)rw; }
    else { A.obs[i] = (long long)deg; A.reward[i] = rw; }       // int(degrees(angle)): truncation
    A.done[i] = 0;
    A.st.now[i] = s.now;
    if (s.fault) { if (atomicCAS(A.errflag, 0, GW_E_SIMFAULT) == 0) { A.errflag[1] = (int)i; A.errflag[2] = s.fault; } }
    store_sim(s, A.st, i, false);
    store_plant(S, A.st, i);
    // statistics (no block reduction: this env is not the throughput path)
    atomicAdd(A.stats + 0, rw);
    atomicAdd(A.stats + 1, (double)(s.nDeliv[0] - nD0));
    atomicAdd(A.stats + 2, (double)(s.nDeliv[1] - nD1));
    atomicAdd(A.stats + 4, 1.0);
    atomicAdd(A.stats + 5, fabs(deg));
    atomicAdd(A.stats + 6, (double)(s.nTx - nTx0));
}

__global__ void pendulum_init_kernel(StatePtrs st, PendulumParams Q)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= st.nsim) return;
    PendulumState S;
    pendulum_init(Q, S);
    store_plant(S, st, i);
    for (int k = 0; k < kMaxSend * kRingSlots; ++k) st.pval[(long long)k * st.nsim + i] = 0.0;
}

__global__ void plant_read_kernel(StatePtrs st, double *out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= st.nsim) return;
    for (int k = 0; k < 8; ++k) out[(long long)k * st.nsim + i] = st.plant[(long long)k * st.nsim + i];
}

// ------------------------------------------------------------------------------------
// init / reset / tables / read-back kernels
// ------------------------------------------------------------------------------------

template <int D, int NS, int NJ>
__global__ void init_kernel(StatePtrs st, Params P, SharedTables thermal /* srx[b][0] = thermal noise */)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= st.nsim) return;
    const int band = (int)(i % st.nb);
    Sim<D, NS, NJ> s;
    init_sim(s, thermal.srx[band][0]);
    if (band == 0) st.now[i / st.nb] = 0.0;
    // write every chunk once so that the cold part is defined
    const long long n = st.nsim;
    for (int c = 0; c < COLD_CHUNKS; ++c) st_chunk(st.cold, n, c, i, make_uint4(0, 0, 0, 0));
    for (int c = 0; c < HOT_CHUNKS; ++c) st_chunk(st.hot, n, c, i, make_uint4(0, 0, 0, 0));
    DevRing ring{st.ring + i, st.nsim, st.hot + i};
    init_receive(s, P.band[band], ring);            // the receive loops' first time-outs take creation numbers
    store_sim(s, st, i, true);
}

template <int D, int NS, int NJ>
__global__ void reset_kernel(StatePtrs st, Params P, const long long *env_ids, long long n_ids, long long *obs)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (env_ids ? n_ids : st.nenv) * st.nb;
    if (t >= total) return;
    const long long e = env_ids ? env_ids[t / st.nb] : t / st.nb;
    const int band = (int)(t % st.nb);
    if (e < 0 || e >= st.nenv) return;
    const long long i = e * st.nb + band;
    Sim<D, NS, NJ> s;
    load_sim(s, st, i, st.now[e]);
    DevRing ring{st.ring + i, st.nsim, st.hot + i};
    reset_sim(s, P.band[band], ring);
    store_sim(s, st, i, true);
    if (obs) obs[i] = (long long)s.latestDiff + kCounterBound;       // counter_traffic.py:144
}

// K1: attenuation and received-power tables, one thread per (table, receiver, sender)
__global__ void tables_kernel(StatePtrs st, const double *pos /* [ntab][MAXD][2] or NULL */,
                              SharedTables defpos_x, SharedTables defpos_y, SharedTables power, SharedTables freq)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= st.ntab * 16) return;
    const long long tab = t / 16;
    const int k = (int)(t % 16), p = k / kMaxDev, d = k % kMaxDev;
    const int band = (int)((st.per_env ? tab : 0) % st.nb);
    double px, py, dx, dy;
    if (pos) {
        px = pos[(tab * kMaxDev + p) * 2]; py = pos[(tab * kMaxDev + p) * 2 + 1];
        dx = pos[(tab * kMaxDev + d) * 2]; dy = pos[(tab * kMaxDev + d) * 2 + 1];
    } else {
        px = defpos_x.srx[band][p]; py = defpos_y.srx[band][p];
        dx = defpos_x.srx[band][d]; dy = defpos_y.srx[band][d];
    }
    double att = 0.0, rp = 0.0;
    if (p != d) {
        att = fspl_db(px, py, dx, dy, freq.srx[band][0]);
        rp = rx_power_mw(power.srx[band][d], att);
    }
    st.att[(long long)k * st.ntab + tab] = att;
    st.srx[(long long)k * st.ntab + tab] = rp;
}

// tables of one band-sim in global memory, entry (receiver p, sender d) at (p * kMaxDev + d) * stride
struct DevTab {
    double *a, *r;
    long long stride;
    __device__ __forceinline__ double att(int p, int d) const { return a[(long long)(p * kMaxDev + d) * stride]; }
    __device__ __forceinline__ void set_att(int p, int d, double v) { a[(long long)(p * kMaxDev + d) * stride] = v; }
    __device__ __forceinline__ double srx(int p, int d) const { return r[(long long)(p * kMaxDev + d) * stride]; }
    __device__ __forceinline__ void set_srx(int p, int d, double v) { r[(long long)(p * kMaxDev + d) * stride] = v; }
    __device__ __forceinline__ SrxView view() const { return SrxView{r, stride}; }
};

// mode M error counts for one thread (position changes are rare: no warp cooperation)
struct SerialMasks {
    MaskSource m;
    long long env;
    int band, nb;
    __device__ long long operator()(int receiver, int sender, uint32_t txseq, long long k0, long long k1, double ber) const
    {
        if (m.mode != MODE_M_FED)
            return mask_errors_serial(m.seed, m.env_offset + env, band, sender, txseq, receiver, k0, k1, ber);
        const long long row = ((((env * nb + band) * kMaxDev + sender) * m.slots + (long long)(txseq % (uint32_t)m.slots)) * kMaxDev + receiver);
        const uint32_t *w = m.words + row * m.words_per_row;
        long long n = 0;
        for (long long k = k0; k < k1; ++k) n += (w[k >> 5] >> (k & 31)) & 1u;
        return n;
    }
};

// gw_set_positions after the first step: every band-sim moves its devices one after the other
// (gw_core.cuh::move_devices -- the reference's Position.set with SimplePhy._onAttenuationChange for the
// transmissions that are on the air), one thread per band-sim
template <int MODE, int D, int NS, int NJ>
__global__ void move_kernel(StatePtrs st, Params P, MaskSource masks, const double *want /* [ntab][kMaxDev][2] */,
                            double *cur /* [ntab][kMaxDev][2] */, SharedTables power, SharedTables freq, int *errflag)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= st.nsim) return;
    const long long e = i / st.nb;
    const int band = (int)(i % st.nb);
    double c[D * 2], w[D * 2], pw[D];
    bool moved = false;
#pragma unroll
    for (int d = 0; d < D; ++d) {
        c[2 * d] = cur[(i * kMaxDev + d) * 2]; c[2 * d + 1] = cur[(i * kMaxDev + d) * 2 + 1];
        w[2 * d] = want[(i * kMaxDev + d) * 2]; w[2 * d + 1] = want[(i * kMaxDev + d) * 2 + 1];
        pw[d] = power.srx[band][d];
        moved |= (c[2 * d] != w[2 * d]) | (c[2 * d + 1] != w[2 * d + 1]);
    }
    if (!moved) return;
    Sim<D, NS, NJ> s;
    load_sim(s, st, i, st.now[e]);
    DevTab tab{st.att + i, st.srx + i, st.ntab};
    SerialMasks mk{masks, e, band, st.nb};
    move_devices<MODE>(s, P, pw, freq.srx[band][0], c, w, tab, mk, NoMemo());
    if (s.fault) { if (atomicCAS(errflag, 0, GW_E_SIMFAULT) == 0) { errflag[1] = (int)i; errflag[2] = s.fault; } }
    store_sim(s, st, i, false);
#pragma unroll
    for (int d = 0; d < D; ++d) { cur[(i * kMaxDev + d) * 2] = c[2 * d]; cur[(i * kMaxDev + d) * 2 + 1] = c[2 * d + 1]; }
}

template <int D, int NS, int NJ>
__global__ void read_kernel(StatePtrs st, Params P, int field, double *out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= st.nsim) return;
    const long long n = st.nsim;
    Sim<D, NS, NJ> s;
    const long long e = i / st.nb;
    const int band = (int)(i % st.nb);
    load_sim(s, st, i, st.now[e]);
    switch (field) {
    case GW_FIELD_NOW: if (band == 0) out[e] = s.now; break;
    case GW_FIELD_RECEIVED_POWER: for (int d = 0; d < kMaxDev; ++d) out[d * n + i] = d < D ? s.P[d < D ? d : 0] : 0.0; break;
    case GW_FIELD_NEXT_TICK: for (int k = 0; k < NS; ++k) out[k * n + i] = s.tTick[k]; break;
    case GW_FIELD_COUNTER:
        for (int k = 0; k < NS; ++k) {
            const unsigned long long c = (unsigned long long)s.epochC[k] + (s.ticks[k] - s.epochK[k]);
            out[k * n + i] = (double)(c > (unsigned long long)kCounterBound ? (unsigned long long)kCounterBound : c);
        }
        break;
    case GW_FIELD_QUEUE_LEN: for (int k = 0; k < NS; ++k) out[k * n + i] = s.qn[k]; break;
    case GW_FIELD_N_TRANSMISSIONS: out[i] = s.nTx; break;
    case GW_FIELD_N_DELIVERED: for (int k = 0; k < NS; ++k) out[k * n + i] = s.nDeliv[k]; break;
    case GW_FIELD_RECEIVED_VALUES: out[i] = s.rv0; out[n + i] = s.rv1; break;
    case GW_FIELD_ATTENUATION_DB:
        for (int k = 0; k < 16; ++k) out[k * n + i] = st.att[(long long)k * st.ntab + (st.per_env ? i : 0)];
        break;
    case GW_FIELD_RX_POWER_MW:
        for (int k = 0; k < 16; ++k) out[k * n + i] = st.srx[(long long)k * st.ntab + (st.per_env ? i : 0)];
        break;
    case GW_FIELD_FAULT: out[i] = s.fault; break;
    case GW_FIELD_TIES: out[i] = s.ties; break;
    case GW_FIELD_N_RECEIVED: {
        DevRing ring{st.ring + i, st.nsim, st.hot + i};
        for (int k = 0; k < NS; ++k) out[k * n + i] = ring.received(k);
        break;
    }
    case GW_FIELD_TX_SEQ: for (int d = 0; d < kMaxDev; ++d) out[d * n + i] = d < D ? s.txSeq[d < D ? d : 0] : 0.0; break;
    default: break;
    }
}

// ------------------------------------------------------------------------------------
// standalone kernels
// ------------------------------------------------------------------------------------

__global__ void fspl_kernel(const double *ax, const double *ay, const double *bx, const double *by,
                            double f, double *att, long long n)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) att[i] = fspl_db(ax[i], ay[i], bx[i], by[i], f);
}

__global__ void ber_kernel(const double *S, const double *N, double *ber, long long n, double c, double qDen)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) ber[i] = ber_bpsk_mw(S[i], N[i], c, qDen);
}

// K3: one warp per descriptor, grid-stride, software-pipelined: the descriptor of the next
// row is loaded while the current row is scanned; 128-bit streaming loads, popc, shuffle reduction
__global__ void __launch_bounds__(256)
count_bits_kernel(const uint32_t *words, int words_per_row, const long long *rows, const int *k0, const int *k1,
                  int *counts, long long n)
{
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    long long i = warp, r = 0;
    int a0 = 0, a1 = 0;
    if (i < n) { a0 = __ldg(k0 + i); a1 = __ldg(k1 + i); r = __ldg(rows + i); }
    while (i < n) {
        const long long inext = i + nwarps;
        long long rn = 0;
        int b0 = 0, b1 = 0;
        if (inext < n) { b0 = __ldg(k0 + inext); b1 = __ldg(k1 + inext); rn = __ldg(rows + inext); }
        const int c = warp_popc_range(words + r * words_per_row, a0, a1, lane);
        if (lane == 0) counts[i] = c;
        i = inext; a0 = b0; a1 = b1; r = rn;
    }
}

// K3, TMA variant: every warp runs its own 4-deep ring of 2 KiB shared-memory stages that are
// filled by 1-D bulk async copies (cp.async.bulk, SASS UBLKCP) completing on mbarriers, so a warp
// keeps up to 8 KiB of mask rows in flight without holding them in registers; the popcount
// reads shared memory.  Rows longer than one stage are processed in 2 KiB pieces.
namespace tma {
constexpr int STAGES = 4;
constexpr int WARPS = 8;
constexpr int STAGE_BYTES = 2048;

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    unsigned done = 0;
    while (!done) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_async_proxy() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
}  // namespace tma

constexpr int TMA_SMEM_BYTES = tma::WARPS * tma::STAGES * (tma::STAGE_BYTES + 8 + 16);

__global__ void __launch_bounds__(256)
count_bits_tma_kernel(const uint32_t *words, int words_per_row, const long long *rows, const int *k0, const int *k1,
                      int *counts, long long n)
{
    using namespace tma;
    extern __shared__ __align__(128) unsigned char tma_smem[];
    typedef uint4 StageBuf[STAGES][STAGE_BYTES / 16];
    StageBuf *buf = reinterpret_cast<StageBuf *>(tma_smem);                                   // [WARPS]
    typedef unsigned long long BarRow[STAGES];
    BarRow *bars = reinterpret_cast<BarRow *>(tma_smem + WARPS * STAGES * STAGE_BYTES);       // [WARPS]
    typedef int MetaRow[STAGES][4];                 // k0, k1 (relative to the staged span), valid bytes, -
    MetaRow *meta = reinterpret_cast<MetaRow *>(tma_smem + WARPS * STAGES * STAGE_BYTES + WARPS * STAGES * 8);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long warp = (long long)blockIdx.x * WARPS + w;
    const long long nwarps = (long long)gridDim.x * WARPS;
    if (lane == 0)
        for (int s = 0; s < STAGES; ++s) mbar_init(smem_u32(&bars[w][s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();

    // rows of this warp: warp, warp + nwarps, ...  (this kernel handles rows that fit one stage;
    // the host falls back to the register variant otherwise)
    const long long mine = warp < n ? (n - warp + nwarps - 1) / nwarps : 0;

    // lane 0 issues the copies; the descriptor of the NEXT row to issue is prefetched into
    // registers one iteration ahead so that its load latency is off the critical path
    int d0 = 0, d1 = 0;
    long long drow = 0;
    auto fetch = [&](long long j) {
        if (j < mine) {
            const long long i = warp + j * nwarps;
            d0 = __ldg(k0 + i); d1 = __ldg(k1 + i); drow = __ldg(rows + i);
        }
    };
    auto issue = [&](long long j) {                 // uses (d0, d1, drow) fetched for row j
        const int s = (int)(j % STAGES);
        const int a0 = d0, a1 = d1;
        const long long row = drow;
        fetch(j + 1);
        const unsigned bar = smem_u32(&bars[w][s]);
        if (a1 > a0) {
            const int q0 = (a0 >> 5) >> 2, q1 = ((a1 - 1) >> 5) >> 2;
            const unsigned bytes = (unsigned)(q1 - q0 + 1) * 16u;
            const uint32_t *src = words + row * words_per_row + q0 * 4;
            meta[w][s][0] = a0 - q0 * 128; meta[w][s][1] = a1 - q0 * 128; meta[w][s][2] = (int)bytes;
            mbar_expect_tx(bar, bytes);
            bulk_g2s(smem_u32(&buf[w][s][0]), src, bytes, bar);
        } else {
            meta[w][s][0] = 0; meta[w][s][1] = 0; meta[w][s][2] = 0;
            mbar_expect_tx(bar, 0);
        }
    };

    if (lane == 0) {
        fetch(0);
        for (long long j = 0; j < mine && j < STAGES; ++j) issue(j);
    }

    for (long long j = 0; j < mine; ++j) {
        const int s = (int)(j % STAGES);
        const unsigned parity = (unsigned)((j / STAGES) & 1);
        mbar_wait(smem_u32(&bars[w][s]), parity);
        const int r0 = meta[w][s][0], r1 = meta[w][s][1];
        int cnt = 0;
        if (r1 > r0) {
            const int q1 = (r1 - 1) >> 7;               // the staged span starts at group 0
#pragma unroll 4
            for (int q = lane; q <= q1; q += 32) cnt += popc_group(buf[w][s][q], q, 0, q1, r0, r1);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
        if (lane == 0) counts[warp + j * nwarps] = cnt;
        __syncwarp();                               // every lane is done reading this stage
        if (lane == 0 && j + STAGES < mine) {
            fence_async_proxy();                    // generic-proxy reads before the async-proxy refill
            issue(j + STAGES);
        }
    }
}

__global__ void philox_kernel(const uint32_t *ctr, const uint32_t *key, uint32_t *out, long long n)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t o[4];
    philox4x32_10(ctr[4 * i], ctr[4 * i + 1], ctr[4 * i + 2], ctr[4 * i + 3], key[2 * i], key[2 * i + 1], o);
    out[4 * i] = o[0]; out[4 * i + 1] = o[1]; out[4 * i + 2] = o[2]; out[4 * i + 3] = o[3];
}

// Action selection of the DQN learner (agents/dqn_counter_traffic.py:46-63), fused: the 1-16-16-16-A ReLU MLP on the
// centred scalar observation, keras-rl's BoltzmannQPolicy (rl/policy.py: p ~ exp(clip(q / tau, lo, hi)), float64) and
// the draw, one thread per env.  The ~1.3 k weights sit in shared memory; the A q-values of an env in registers.
// The uniform variate of (env, draw counter) is Philox4x32-10 keyed by the seed, so the sample does not depend on
// the batch size or the sharding.  weights: float32, the order of torch's model.parameters() -- W1[16][1], b1[16],
// W2[16][16], b2[16], W3[16][16], b3[16], W4[A][16], b4[A].
constexpr int POLICY_HIDDEN = 16;

template <int A>
__global__ void __launch_bounds__(128)
policy_kernel(const float *weights, const long long *obs, long long n, float obs_center, double tau, double clip_lo,
              double clip_hi, unsigned long long seed, unsigned long long counter, long long env_offset, int n_durations,
              long long *flat, int *device, int *duration, double *probs)
{
    constexpr int H = POLICY_HIDDEN;
    constexpr int NW = H + H + H * H + H + H * H + H + A * H + A;
    __shared__ float w[NW];
    for (int t = threadIdx.x; t < NW; t += blockDim.x) w[t] = weights[t];
    __syncthreads();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *W1 = w, *b1 = W1 + H, *W2 = b1 + H, *b2 = W2 + H * H, *W3 = b2 + H, *b3 = W3 + H * H, *W4 = b3 + H,
                *b4 = W4 + A * H;
    const float x = (float)obs[i] - obs_center;
    float h1[H], h2[H], h3[H];
#pragma unroll
    for (int j = 0; j < H; ++j) h1[j] = fmaxf(__fmaf_rn(W1[j], x, b1[j]), 0.f);
#pragma unroll
    for (int j = 0; j < H; ++j) {
        float a = b2[j];
#pragma unroll
        for (int k = 0; k < H; ++k) a = __fmaf_rn(W2[j * H + k], h1[k], a);
        h2[j] = fmaxf(a, 0.f);
    }
#pragma unroll
    for (int j = 0; j < H; ++j) {
        float a = b3[j];
#pragma unroll
        for (int k = 0; k < H; ++k) a = __fmaf_rn(W3[j * H + k], h2[k], a);
        h3[j] = fmaxf(a, 0.f);
    }
    double e[A];
    double sum = 0.0;
#pragma unroll
    for (int j = 0; j < A; ++j) {
        float q = b4[j];
#pragma unroll
        for (int k = 0; k < H; ++k) q = __fmaf_rn(W4[j * H + k], h3[k], q);
        const double z = fmin(fmax((double)q / tau, clip_lo), clip_hi);
        e[j] = exp(z);
        sum += e[j];
    }
    // inverse CDF with u in (0, 1): the first action whose cumulative probability exceeds u
    uint32_t r[4];
    const unsigned long long env = (unsigned long long)(env_offset + i);
    philox4x32_10((uint32_t)env, (uint32_t)(env >> 32), (uint32_t)counter, (uint32_t)(counter >> 32),
                  (uint32_t)seed, (uint32_t)(seed >> 32), r);
    const double u = ((double)r[0] * 4294967296.0 + (double)r[1] + 0.5) * (1.0 / 18446744073709551616.0);
    const double target = u * sum;
    double acc = 0.0;
    int a = A - 1;
    bool found = false;
#pragma unroll
    for (int j = 0; j < A; ++j) {
        acc += e[j];
        if (!found && acc > target) { a = j; found = true; }
    }
    if (flat) flat[i] = a;
    if (device) { device[i] = a / n_durations; duration[i] = a % n_durations; }       // CounterTrafficProcessor, :25-33
    if (probs) {
#pragma unroll
        for (int j = 0; j < A; ++j) probs[i * A + j] = e[j] / sum;
    }
}

// The same policy for action counts the unrolled kernels are not instantiated for (bands of 3..8 senders: 60..160
// actions): run-time A, the q-values are evaluated twice (normalisation, then the inverse CDF) instead of being kept
// in registers -- same arithmetic per value, same draw.
constexpr int POLICY_MAX_ACTIONS_N = 160;

__global__ void __launch_bounds__(128)
policy_kernel_n(const float *weights, int A, const long long *obs, long long n, float obs_center, double tau, double clip_lo,
                double clip_hi, unsigned long long seed, unsigned long long counter, long long env_offset, int n_durations,
                long long *flat, int *device, int *duration, double *probs)
{
    constexpr int H = POLICY_HIDDEN;
    __shared__ float w[H + H + H * H + H + H * H + H + POLICY_MAX_ACTIONS_N * H + POLICY_MAX_ACTIONS_N];
    const int NW = H + H + H * H + H + H * H + H + A * H + A;
    for (int t = threadIdx.x; t < NW; t += blockDim.x) w[t] = weights[t];
    __syncthreads();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *W1 = w, *b1 = W1 + H, *W2 = b1 + H, *b2 = W2 + H * H, *W3 = b2 + H, *b3 = W3 + H * H, *W4 = b3 + H,
                *b4 = W4 + A * H;
    const float x = (float)obs[i] - obs_center;
    float h1[H], h2[H], h3[H];
#pragma unroll
    for (int j = 0; j < H; ++j) h1[j] = fmaxf(__fmaf_rn(W1[j], x, b1[j]), 0.f);
#pragma unroll
    for (int j = 0; j < H; ++j) {
        float a = b2[j];
#pragma unroll
        for (int k = 0; k < H; ++k) a = __fmaf_rn(W2[j * H + k], h1[k], a);
        h2[j] = fmaxf(a, 0.f);
    }
#pragma unroll
    for (int j = 0; j < H; ++j) {
        float a = b3[j];
#pragma unroll
        for (int k = 0; k < H; ++k) a = __fmaf_rn(W3[j * H + k], h2[k], a);
        h3[j] = fmaxf(a, 0.f);
    }
    auto weight_of = [&](int j) {
        float q = b4[j];
#pragma unroll
        for (int k = 0; k < H; ++k) q = __fmaf_rn(W4[j * H + k], h3[k], q);
        return exp(fmin(fmax((double)q / tau, clip_lo), clip_hi));
    };
    double sum = 0.0;
    for (int j = 0; j < A; ++j) sum += weight_of(j);
    uint32_t r[4];
    const unsigned long long env = (unsigned long long)(env_offset + i);
    philox4x32_10((uint32_t)env, (uint32_t)(env >> 32), (uint32_t)counter, (uint32_t)(counter >> 32),
                  (uint32_t)seed, (uint32_t)(seed >> 32), r);
    const double u = ((double)r[0] * 4294967296.0 + (double)r[1] + 0.5) * (1.0 / 18446744073709551616.0);
    const double target = u * sum;
    double acc = 0.0;
    int a = A - 1;
    bool found = false;
    for (int j = 0; j < A; ++j) {
        const double e = weight_of(j);
        acc += e;
        if (!found && acc > target) { a = j; found = true; }
        if (probs) probs[i * A + j] = e / sum;
    }
    if (flat) flat[i] = a;
    if (device) { device[i] = a / n_durations; duration[i] = a % n_durations; }
}

__global__ void stats_copy_kernel(double *stats, double *out, int clear)
{
    const int k = threadIdx.x;
    if (k < 8) { out[k] = stats[k]; if (clear) stats[k] = 0.0; }
}

// ------------------------------------------------------------------------------------
// grids of PHY-only senders with in-step mobility (gw_grid.cuh): one band-sim per thread, its state block in
// global memory
// ------------------------------------------------------------------------------------

struct GridArgs {
    char *state;                // [n_envs] blocks of `block_bytes`
    size_t block_bytes;
    long long n_envs;
    const double *offsets;      // [n_envs][n_dev][max_moves][2] or NULL
    double *trace;              // [n_envs][cap][8] or NULL
    int *trace_count;
    int trace_cap;
    int *errflag;
};

__global__ void grid_init_kernel(GridArgs A, GridParams G, const double *pos, const double *delays, const double *move_delays)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n_envs) return;
    GridView v = grid_view(A.state + (size_t)i * A.block_bytes, G.ndev);
    grid_init(v, G, pos + i * G.ndev * 2, delays + i * G.ndev, move_delays ? move_delays + i * G.ndev : nullptr);
}

// Work items collected by the lanes of a warp -- lane l holds the set bits of `mask` -- spread over the whole warp:
// f(owner lane, bit) is called once per item, 32 items at a time, each by a different lane.  All 32 lanes must call.
// Used for SimplePhy._updateBitErrorRate in the general engines: one band-sim per thread, the lanes of a warp sit at
// different events of different kinds (ncu: 4 of 32 lanes active per instruction), and what most of them execute
// most of the time is the BER evaluation (2 log10, exp10, sqrt, 2 exp, a division: ~600 fp64-heavy instructions) --
// on a band with a dozen PHYs every transmission that starts or ends makes every receiving PHY re-evaluate.  The
// evaluations of an event read nothing the rest of the event writes, so the transition functions only COLLECT them
// and the warp evaluates the (env, PHY) pairs of all its lanes together: any lane can serve any env because the
// state lives in global memory.
template <class F>
__device__ __forceinline__ void warp_spread(uint32_t mask, unsigned lane, F &&f)
{
    constexpr unsigned FULL = 0xffffffffu;
    const int cnt = __popc(mask);
    int pre = cnt;                                      // inclusive prefix sum of the lanes' counts
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(FULL, pre, o);
        if ((int)lane >= o) pre += t;
    }
    const int total = __shfl_sync(FULL, pre, 31);
    const int excl = pre - cnt;
    for (int t0 = 0; t0 < total; t0 += 32) {
        const int t = t0 + (int)lane;
        int owner = 0;                                  // owner of item t: the number of lanes whose inclusive prefix is <= t
#pragma unroll
        for (int stepw = 16; stepw >= 1; stepw >>= 1) {
            const int q = __shfl_sync(FULL, pre, owner + stepw - 1);
            if (q <= t) owner += stepw;
        }
        const int ol = owner < 32 ? owner : 31;
        const int oExcl = __shfl_sync(FULL, excl, ol);
        const uint32_t oMask = __shfl_sync(FULL, mask, ol);
        if (t < total) f(ol, (int)__fns(oMask, 0u, t - oExcl + 1));
    }
}

__global__ void __launch_bounds__(64)
grid_run_kernel(GridArgs A, GridParams G, double duration)
{
    constexpr unsigned FULL = 0xffffffffu;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < A.n_envs;
    const unsigned lane = threadIdx.x & 31u;
    const long long warpEnv0 = i - lane;
    GridView v = grid_view(A.state + (size_t)(valid ? i : A.n_envs - 1) * A.block_bytes, G.ndev);
    if (A.trace && valid) { v.trace = A.trace + i * A.trace_cap * 8; v.traceCap = A.trace_cap; v.ntrace = 0; }
    const double *off = A.offsets ? A.offsets + (valid ? i : 0) * G.ndev * G.maxMoves * 2 : nullptr;
    const double T = v.h->now + duration;
    const bool serial = A.trace != nullptr;             // the traced variant keeps the BER records in event order
    bool run = valid;
    for (;;) {
        uint32_t mask = 0;
        if (run) run = grid_run_event(v, G, T, off, mask);
        if (serial) {
            if (run) grid_update_bers(v, G, mask);
        } else {
            __syncwarp();                               // the owners' state updates are visible to the helpers
            warp_spread(mask, lane, [&](int ol, int p) {
                GridView w = grid_view(A.state + (size_t)(warpEnv0 + ol) * A.block_bytes, G.ndev);
                grid_update_ber(w, G, p);
            });
            __syncwarp();                               // the evaluated rates are visible to their owners
        }
        if (!__any_sync(FULL, run)) break;
    }
    if (!valid) return;
    v.h->now = T;
    if (A.trace) A.trace_count[i] = v.ntrace;
    if (v.h->fault) { if (atomicCAS(A.errflag, 0, GW_E_SIMFAULT) == 0) { A.errflag[1] = (int)i; A.errflag[2] = v.h->fault; } }
}

__global__ void grid_read_kernel(GridArgs A, GridParams G, int field, double *out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n_envs) return;
    const GridView v = grid_view(A.state + (size_t)i * A.block_bytes, G.ndev);
    const long long n = A.n_envs;
    if (field == GW_GRID_FIELD_NOW) { out[i] = v.h->now; return; }
    if (field == GW_GRID_FIELD_FAULT) { out[i] = v.h->fault; return; }
    for (int d = 0; d < G.ndev; ++d) {
        const GridDev &D = v.dev[d];
        if (field == GW_GRID_FIELD_STATS) {
            const unsigned st[6] = {D.nTx, D.nHdrOk, D.nHdrFail, D.nPayOk, D.nPayFail, D.nBer};
            for (int k = 0; k < 6; ++k) out[((long long)k * G.ndev + d) * n + i] = st[k];
        } else if (field == GW_GRID_FIELD_POSITIONS) {
            out[((long long)0 * G.ndev + d) * n + i] = D.x; out[((long long)1 * G.ndev + d) * n + i] = D.y;
        } else if (field == GW_GRID_FIELD_RECEIVED_POWER) {
            out[(long long)d * n + i] = D.P;
        }
    }
}

// ------------------------------------------------------------------------------------
// general band engine (gw_band.cuh): run-time device counts, one band-sim per thread, state in global memory
// laid out [word][env] so that a warp's accesses to one field are contiguous
// ------------------------------------------------------------------------------------

struct GenArgs {
    double *f64;                // [gen_f64_words][n_envs]
    int32_t *i32;               // [gen_i32_words][n_envs]
    double *srx;                // [nd * nd][n_envs] or [nd * nd]
    double *att;                // [nd * nd][n_envs] attenuation (dB) and
    double *pos;                // [nd * 2][n_envs] current positions: per-env geometries only (devices may move)
    double *mvT;                // [2][nd][n_envs] mobility processes: next wake-up, first delay (gw_genband_set_movers)
    int32_t *mvI;               // [3][nd][n_envs] creation number, stage, jumps done
    const double *offsets;      // [n_envs][nd][max_moves][2]
    long long n_envs;
    int per_env;
    double *trace;              // [n_envs][cap][8] or NULL
    int *trace_count;
    int trace_cap;
    int *errflag;
};

__device__ __forceinline__ GenView gen_view_of(const GenArgs &A, const GenBand &B, long long i, int mode = -1)
{
    GenView v;
    v.mode = mode < 0 ? B.mode : mode;
    v.f = A.f64 + i; v.i = A.i32 + i; v.stride = A.n_envs;
    v.srx = A.per_env ? A.srx + i : A.srx; v.srxStride = A.per_env ? A.n_envs : 1;
    v.att = A.per_env ? A.att + i : nullptr; v.pos = A.per_env ? A.pos + i : nullptr;
    v.mvT = A.mvT ? A.mvT + i : nullptr; v.mvDelay = A.mvT ? A.mvT + (long long)B.nd * A.n_envs + i : nullptr;
    v.mvI = A.mvT ? A.mvI + i : nullptr; v.offsets = A.mvT ? A.offsets + i * B.nd * B.maxMoves * 2 : nullptr;
    v.ns = B.ns; v.nj = B.nj; v.nd = B.nd; v.env = B.envOffset + i;
    v.trace = nullptr; v.ntrace = 0; v.traceCap = 0;
    return v;
}

__global__ void genband_init_kernel(GenArgs A, GenBand B, const double *pos, const double *power, double frequency)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n_envs) return;
    // per-env geometries: the table is evaluated here (libdevice log10 / pow: within 2 ulp of the host's); a geometry
    // common to all envs is evaluated once on the host, with the C library the reference's Python runs on
    if (A.per_env) gen_power_table(B.nd, pos + i * B.nd * 2, power, frequency, A.srx + i, A.n_envs, A.att + i, A.pos + i);
    GenView v = gen_view_of(A, B, i);
    gen_init(v, B);
}

// gw_genband_set_positions: every env moves its devices one after the other (gw_band.cuh::gen_move_devices)
__global__ void genband_movers_kernel(GenArgs A, GenBand B, const double *move_delays)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n_envs) return;
    GenView v = gen_view_of(A, B, i);
    gen_start_movers(v, move_delays + i * B.nd);
}

__global__ void genband_move_kernel(GenArgs A, Params P, GenBand B, const double *want)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n_envs) return;
    GenView v = gen_view_of(A, B, i);
    if (v.sc(GenView::I_fault)) return;
    gen_move_devices(v, P, B, want + i * B.nd * 2);
    const int fault = v.sc(GenView::I_fault);
    if (fault) { if (atomicCAS(A.errflag, 0, GW_E_SIMFAULT) == 0) { A.errflag[1] = (int)i; A.errflag[2] = fault; } }
}

__global__ void genband_reset_kernel(GenArgs A, GenBand B, long long *obs)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n_envs) return;
    GenView v = gen_view_of(A, B, i);
    gen_reset(v, B);
    if (obs) obs[i] = (long long)kCounterBound;
}

// launch configuration (profiles/README.md, r4e): 128 threads x 4 blocks per SM = 128 registers, 16 resident warps per
// SM -- 2.57 ms per step of the 8 + 1 + 4 band against 3.72 ms with 64-thread blocks at 152 registers
#ifndef GW_GEN_BLOCK
#define GW_GEN_BLOCK 128
#endif
#ifndef GW_GEN_MINB
#define GW_GEN_MINB 4
#endif
// One band-sim per thread.  The BER evaluations an event asks for are collected (gen_apply's berMask) and evaluated by
// the warp as a whole (warp_spread); lanes whose step has ended keep helping until the whole warp is done.
template <bool MODE_M>
__global__ void __launch_bounds__(GW_GEN_BLOCK, GW_GEN_MINB)
genband_step_kernel(GenArgs A, Params P, GenBand B, const int32_t *device, const int32_t *duration, long long *obs,
                    double *reward, unsigned char *done)
{
    constexpr unsigned FULL = 0xffffffffu;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < A.n_envs;
    const unsigned lane = threadIdx.x & 31u;
    const long long warpEnv0 = i - lane;
    constexpr int kMode = MODE_M ? MODE_M_PHILOX : MODE_R;
    GenView v = gen_view_of(A, B, valid ? i : A.n_envs - 1, kMode);
    if (A.trace) { v.trace = A.trace + (valid ? i : 0) * A.trace_cap * 8; v.traceCap = A.trace_cap; }
    int before = 0;
    bool run = false;
    if (valid) {
        before = v.sc(GenView::I_fault);
        run = gen_step_begin(v, P, B, device[i], duration[i]);
    }
    const bool serial = A.trace != nullptr;             // the traced variant keeps the BER records in event order
    constexpr bool modeM = MODE_M;
    for (;;) {
        uint32_t mask = 0;
        if (modeM && !serial) {
            // mode M: the event is selected first; the bit ranges its PHYs count (gen_count_set) are counted by the
            // warp, one range after the other, 32 Philox blocks (128 bits) per pass; the transition function then finds
            // them counted.  The ranges of a section are ~10^2..10^4 bits per receiver against ~10^2 instructions per
            // Philox block: counted by the owning lane alone they would dwarf everything else.
            Event ev;
            uint32_t cset = 0;
            ev.kind = EV_NONE; ev.idx = 0; ev.t = 0; ev.seq = 0;
            if (run) {
                if (gen_step_running(v)) { ev = gen_next_event(v, B); v.now() = ev.t; cset = gen_count_set(v, ev); }
                else run = false;
            }
            __syncwarp();
            unsigned owners = __ballot_sync(FULL, cset != 0);
            while (owners) {
                const int ol = __ffs(owners) - 1;
                owners &= owners - 1;
                uint32_t m = __shfl_sync(FULL, cset, ol);
                GenView w = gen_view_of(A, B, warpEnv0 + ol, kMode);
                while (m) {
                    const int p = __ffs(m) - 1;
                    m &= m - 1;
                    int sender; uint32_t txseq; long long k0, k1;
                    gen_mask_range(w, P, p, sender, txseq, k0, k1);
                    const int c = warp_philox_range(B.seed, w.env, 0, sender, txseq, p, (int)k0, (int)k1, ber_threshold(w.ber(p)), (int)lane);
                    if ((int)lane == ol) { w.err(p) += (double)c; w.segT0(p) = w.now(); }
                }
            }
            __syncwarp();
            if (run) mask = gen_apply(v, P, B, ev);
        } else if (run) {
            if (gen_step_running(v)) mask = gen_step_event(v, P, B);
            else run = false;
        }
        if (serial) {
            if (run) gen_update_bers(v, P, mask);
            if (!__any_sync(FULL, run)) break;
            continue;
        }
        __syncwarp();                                   // the owners' state updates are visible to the helpers
        warp_spread(mask, lane, [&](int ol, int p) {
            GenView w = gen_view_of(A, B, warpEnv0 + ol, kMode);
            gen_update_ber(w, P, p);
        });
        __syncwarp();                                   // the evaluated rates are visible to their owners
        if (!__any_sync(FULL, run)) break;
    }
    if (!valid) return;
    long long o; double r; unsigned char d;
    gen_step_end(v, o, r, d);
    obs[i] = o; reward[i] = r; done[i] = d;
    if (A.trace) A.trace_count[i] = v.ntrace;
    const int fault = v.sc(GenView::I_fault);
    if (fault && !before) {
        const int code = fault == FAULT_EMPTY + 1 ? GW_E_ACTION : GW_E_SIMFAULT;
        if (fault == FAULT_EMPTY + 1) v.sc(GenView::I_fault) = 0;      // a rejected action leaves the env as it was
        if (atomicCAS(A.errflag, 0, code) == 0) { A.errflag[1] = (int)i; A.errflag[2] = fault; }
    }
}

__global__ void genband_read_kernel(GenArgs A, GenBand B, int field, double *out)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.n_envs) return;
    const GenView v = gen_view_of(A, B, i);
    const long long n = A.n_envs;
    switch (field) {
    case GW_GENBAND_FIELD_NOW: out[i] = v.now(); break;
    case GW_GENBAND_FIELD_TRANSMISSIONS: out[i] = v.sc(GenView::I_nTx); break;
    case GW_GENBAND_FIELD_FAULT: out[i] = v.sc(GenView::I_fault); break;
    case GW_GENBAND_FIELD_TIES: out[i] = v.sc(GenView::I_ties); break;
    case GW_GENBAND_FIELD_RECEIVED_POWER: for (int p = 0; p < B.nd; ++p) out[p * n + i] = v.P(p); break;
    default:
        for (int k = 0; k < B.ns; ++k) {
            double x = 0;
            if (field == GW_GENBAND_FIELD_DELIVERED) x = v.nDeliv(k);
            else if (field == GW_GENBAND_FIELD_RECEIVED) x = v.nRecv(k);
            else if (field == GW_GENBAND_FIELD_QUEUE_LENGTH) x = v.qn(k);
            else if (field == GW_GENBAND_FIELD_COUNTER) x = gen_counter_at(v, k, v.ticks(k));
            else if (field == GW_GENBAND_FIELD_RECEIVED_VALUES) x = ((v.sc(GenView::I_rvMask) >> k) & 1) ? kCounterByteLen : 0;
            out[k * n + i] = x;
        }
    }
}

struct gw_genband_handle {
    gw_genband_config cfg;
    int device;
    double *dpower;             // transmission powers (dBm) per device
    Params P;
    GenBand B;
    GenArgs A;
    int *errflag;
};

struct gw_grid_handle {
    gw_grid_config cfg;
    int device;
    GridParams G;
    GridArgs A;
    int *errflag;
};

// ------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------

static int grid_for(long long n, int block) { return (int)((n + block - 1) / block); }

double gw_max_correctable_ber(int k, int n)        // Mcs.maxCorrectableBer, physical.py:160-185
{
    const double bound = std::pow(2.0, (double)(n - k));
    double sum = 0;
    int t = 0;
    while (sum <= bound) {
        double c = 1;
        for (int i = 1; i <= t; i++) c = c * (double)(n - t + i) / (double)i;   // binom(n, t)
        sum += c;
        t += 1;
    }
    t -= 1;
    return (double)t / n;
}

static_assert(sizeof(gw_config) <= 2048, "INTEGRATION.md tells binders to reserve 2048 bytes for gw_config");

static int validate(const gw_config *cfg, int &D, int &NS, int &NJ)
{
    if (!cfg) return fail(GW_E_INVALID, "cfg is NULL");
    if (cfg->abi_version != GW_ABI_VERSION) return fail(GW_E_INVALID, "abi_version %d != %d", cfg->abi_version, GW_ABI_VERSION);
    if (cfg->n_envs < 1) return fail(GW_E_INVALID, "n_envs must be >= 1");
    if (cfg->n_bands != 1 && cfg->n_bands != 2 && cfg->n_bands != 4) return fail(GW_E_INVALID, "n_bands must be 1, 2 or 4");
    if (cfg->mode < GW_MODE_REFERENCE || cfg->mode > GW_MODE_MASK_FED) return fail(GW_E_INVALID, "unknown mode %d", cfg->mode);
    if (cfg->assignment_duration_factor < 1 || cfg->max_assign_duration < 1) return fail(GW_E_INVALID, "bad duration constants");
    NS = -1; NJ = -1;
    for (int b = 0; b < cfg->n_bands; ++b) {
        const gw_band_config &B = cfg->band[b];
        if (B.n_devices < 3 || B.n_devices > GW_MAX_DEVICES) return fail(GW_E_INVALID, "band %d: n_devices %d unsupported", b, B.n_devices);
        int ns = 0, nj = 0, stage = 0;
        for (int d = 0; d < B.n_devices; ++d) {
            const int r = B.device[d].role;
            if (r == GW_ROLE_SENDER) { if (stage != 0) return fail(GW_E_INVALID, "device order must be senders, rrm, jammers"); ns++; }
            else if (r == GW_ROLE_RRM) { if (stage != 0) return fail(GW_E_INVALID, "exactly one RRM per band"); stage = 1; }
            else if (r == GW_ROLE_JAMMER) { if (stage != 1) return fail(GW_E_INVALID, "device order must be senders, rrm, jammers"); nj++; }
            else return fail(GW_E_INVALID, "band %d device %d: unknown role %d", b, d, r);
        }
        if (stage != 1) return fail(GW_E_INVALID, "band %d has no RRM", b);
        if (ns != 2 || nj > GW_MAX_JAMMERS) return fail(GW_E_INVALID, "supported bands: 2 senders, 1 RRM, 0-1 jammers");
        if (NS >= 0 && (ns != NS || nj != NJ)) return fail(GW_E_INVALID, "all bands must have the same device roles");
        NS = ns; NJ = nj;
        for (int k = 0; k < ns; ++k) {
            if (B.device[k].multiplicity < 1 || B.device[k].multiplicity > 16) return fail(GW_E_INVALID, "multiplicity out of range");
            if (!(B.device[k].interval > 0)) return fail(GW_E_INVALID, "interval must be > 0");
            if (B.device[k].payload_bytes > 60000) return fail(GW_E_INVALID, "payload_bytes too large");
            if (B.device[k].max_ticks < 0) return fail(GW_E_INVALID, "max_ticks must be >= 0");
            if ((B.device[k].max_ticks != 0 || B.device[k].receive) && cfg->plant != GW_PLANT_NONE)
                return fail(GW_E_INVALID, "plant envs define their own traffic and reception");
        }
        for (int d = ns + 1; d < B.n_devices; ++d) {
            if (!(B.device[d].jam_interval > 0) || B.device[d].jam_delay < 0) return fail(GW_E_INVALID, "bad jammer timing");
            if (B.device[d].jam_header_bytes < 1 || B.device[d].jam_payload_bytes < 1) return fail(GW_E_INVALID, "bad jammer packet");
        }
    }
    D = NS + 1 + NJ;
    if (cfg->plant != GW_PLANT_NONE) {
        if (cfg->plant != GW_PLANT_SLIDING_PENDULUM) return fail(GW_E_INVALID, "unknown plant %d", cfg->plant);
        if (cfg->n_bands != 1 || NS != 2 || NJ != 1) return fail(GW_E_INVALID, "a plant env has one band: sensor, controller, RRM, actuator");
        if (cfg->mode != GW_MODE_REFERENCE) return fail(GW_E_INVALID, "plant envs run in GW_MODE_REFERENCE");
        if (!cfg->per_env_positions) return fail(GW_E_INVALID, "plant envs need per_env_positions = 1 (links follow the wagon)");
        const gw_pendulum_config &pc = cfg->pendulum;
        if (!(pc.cart_mass > 0) || !(pc.pendulum_mass > 0) || !(pc.arm_length > 0) || !(pc.dt_max > 0) || !(pc.motor_kservo >= 0))
            return fail(GW_E_INVALID, "bad pendulum parameters");
    }
    return GW_OK;
}

static void fill_params(const gw_config &cfg, Params &P)
{
    std::memset(&P, 0, sizeof P);
    P.nbands = cfg.n_bands;
    P.factor = cfg.assignment_duration_factor;
    P.maxDuration = cfg.max_assign_duration;
    P.mode = cfg.mode;
    P.bitRate = 133.33333e3;                        // physical.py:196
    P.dataRate = 0.75 * P.bitRate;                  // physical.py:197
    P.maxBer = gw_max_correctable_ber(3, 4);
    P.tenLog10BitRate = 10 * std::log10(P.bitRate);
    P.qDen = 1.135 * std::sqrt(2 * 3.141592653589793);
    P.bitsFactor = 2 - 0.75;
    // GYMWIPE_B200_NO_MACRO=1: every timed event through the generic transition function (A/B tests)
    finish_params(P);
    P.noMacro = std::getenv("GYMWIPE_B200_NO_MACRO") ? std::atoi(std::getenv("GYMWIPE_B200_NO_MACRO")) : 0;
    if (P.noMacro < 0) P.noMacro = 0;
    for (int b = 0; b < cfg.n_bands; ++b) {
        const gw_band_config &cb = cfg.band[b];
        BandParams &B = P.band[b];
        B.ndev = cb.n_devices; B.ns = 0; B.nj = 0;
        for (int d = 0; d < cb.n_devices; ++d) {
            const gw_device_config &dc = cb.device[d];
            if (dc.role == GW_ROLE_SENDER) {
                B.mult[B.ns] = dc.multiplicity; B.payloadRule[B.ns] = dc.payload_bytes < 0 ? -1 : dc.payload_bytes;
                B.interval[B.ns] = dc.interval; B.maxTicks[B.ns] = dc.max_ticks; B.recv[B.ns] = dc.receive ? 1 : 0;
                if (dc.max_ticks != 0 || dc.receive) P.noMacro = 1;      // finite bursts / receive mode: generic path
                B.ns++;
            } else if (dc.role == GW_ROLE_JAMMER) {
                B.jamInterval[B.nj] = dc.jam_interval; B.jamDelay[B.nj] = dc.jam_delay;
                B.jamHdr[B.nj] = dc.jam_header_bytes; B.jamPay[B.nj] = dc.jam_payload_bytes; B.nj++;
            }
        }
    }
}

#define DISPATCH_SHAPE(h, CALL)                                         \
    do {                                                                \
        if ((h)->NJ == 0) { CALL(3, 2, 0); } else { CALL(4, 2, 1); }    \
    } while (0)

extern "C" {

int gw_abi_version(void) { return GW_ABI_VERSION; }
const char *gw_last_error(void) { return g_err; }

int gw_device_count(int *count)
{
    if (!count) return fail(GW_E_INVALID, "count is NULL");
    *count = 0;
    CUDA_TRY(cudaGetDeviceCount(count));
    return GW_OK;
}

int gw_default_config(gw_config *cfg, int64_t n_envs)
{
    if (!cfg) return fail(GW_E_INVALID, "cfg is NULL");
    std::memset(cfg, 0, sizeof *cfg);
    cfg->abi_version = GW_ABI_VERSION;
    cfg->n_envs = n_envs;
    cfg->n_bands = 1;
    cfg->assignment_duration_factor = 1000;         // envs/core.py:27
    cfg->max_assign_duration = 20;                  // envs/core.py:25
    cfg->mode = GW_MODE_REFERENCE;
    gw_band_config &b = cfg->band[0];
    b.n_devices = 3;
    b.frequency_hz = 2.4e9; b.bandwidth_hz = 22e6;  // physical.py:298
    b.device[0].role = GW_ROLE_SENDER; b.device[0].x = 0; b.device[0].y = 2;
    b.device[0].multiplicity = 1; b.device[0].payload_bytes = -1; b.device[0].interval = 0.001;
    b.device[1].role = GW_ROLE_SENDER; b.device[1].x = 0; b.device[1].y = -2;
    b.device[1].multiplicity = 3; b.device[1].payload_bytes = -1; b.device[1].interval = 0.001;
    b.device[2].role = GW_ROLE_RRM; b.device[2].x = 0; b.device[2].y = 0;
    return GW_OK;
}

int gw_state_bytes(const gw_config *cfg, size_t *bytes)
{
    int D, NS, NJ;
    const int rc = validate(cfg, D, NS, NJ);
    if (rc) return rc;
    if (!bytes) return fail(GW_E_INVALID, "bytes is NULL");
    const long long ntab = cfg->per_env_positions ? cfg->n_envs * cfg->n_bands : 1;
    *bytes = make_layout(cfg->n_envs, cfg->n_bands, ntab, cfg->plant).total;
    return GW_OK;
}

static int launch_tables(gw_handle *h, const double *pos, cudaStream_t s)
{
    SharedTables px, py, pw, fr;
    std::memset(&px, 0, sizeof px); std::memset(&py, 0, sizeof py);
    std::memset(&pw, 0, sizeof pw); std::memset(&fr, 0, sizeof fr);
    for (int b = 0; b < h->cfg.n_bands; ++b) {
        for (int d = 0; d < kMaxDev; ++d) {
            px.srx[b][d] = h->default_pos[b][d][0];
            py.srx[b][d] = h->default_pos[b][d][1];
            pw.srx[b][d] = h->power_dbm[b][d];
        }
        fr.srx[b][0] = h->frequency[b];
    }
    const long long n = h->st.ntab * 16;
    tables_kernel<<<grid_for(n, 128), 128, 0, s>>>(h->st, pos, px, py, pw, fr);
    CUDA_TRY(cudaGetLastError());
    return GW_OK;
}

int gw_create(const gw_config *cfg, int device, void *state, size_t state_bytes, void *stream, gw_handle **out)
{
    int D, NS, NJ;
    int rc = validate(cfg, D, NS, NJ);
    if (rc) return rc;
    if (!out) return fail(GW_E_INVALID, "out is NULL");
    int ndev = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (ndev < 1) return fail(GW_E_CUDA, "no CUDA device: gymwipe_b200 has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(GW_E_INVALID, "device %d out of range (%d devices)", device, ndev);
    CUDA_TRY(cudaSetDevice(device));
    gw_handle *h = new (std::nothrow) gw_handle();
    if (!h) return fail(GW_E_INVALID, "out of host memory");
    std::memset(h, 0, sizeof *h);
    h->generation = next_generation();
    h->cfg = *cfg;
    h->device = device;
    h->D = D; h->NS = NS; h->NJ = NJ;
    h->pdl = std::getenv("GYMWIPE_B200_NO_PDL") ? 0 : 1;
    fill_params(*cfg, h->P);
    for (int b = 0; b < cfg->n_bands; ++b)
        for (int d = 0; d < cfg->band[b].n_devices; ++d)
            if (cfg->band[b].device[d].role == GW_ROLE_SENDER && (cfg->band[b].device[d].receive || cfg->band[b].device[d].max_ticks != 0))
                h->ext = 1;
    const long long ntab = cfg->per_env_positions ? cfg->n_envs * cfg->n_bands : 1;
    h->layout = make_layout(cfg->n_envs, cfg->n_bands, ntab, cfg->plant);
    if (state) {
        if (state_bytes < h->layout.total) { delete h; return fail(GW_E_STATE, "state buffer has %zu bytes, %zu needed", state_bytes, h->layout.total); }
        if (((uintptr_t)state & 255) != 0) { delete h; return fail(GW_E_STATE, "state buffer must be 256-byte aligned"); }
    } else {
        cudaError_t e = cudaMalloc(&h->owned_state, h->layout.total);
        if (e != cudaSuccess) { delete h; return fail(GW_E_CUDA, "cudaMalloc(%zu) failed: %s", h->layout.total, cudaGetErrorString(e)); }
        state = h->owned_state;
    }
    char *base = (char *)state;
    h->st.nenv = cfg->n_envs; h->st.nb = cfg->n_bands; h->st.nsim = cfg->n_envs * cfg->n_bands; h->st.ntab = ntab;
    h->st.per_env = cfg->per_env_positions ? 1 : 0;
    h->st.now = (double *)(base + h->layout.off_now);
    h->st.hot = (uint4 *)(base + h->layout.off_hot);
    h->st.cold = (uint4 *)(base + h->layout.off_cold);
    h->st.ring = (int32_t *)(base + h->layout.off_ring);
    h->st.att = (double *)(base + h->layout.off_att);
    h->st.srx = (double *)(base + h->layout.off_srx);
    h->st.plant = cfg->plant ? (double *)(base + h->layout.off_plant) : nullptr;
    h->st.pval = cfg->plant ? (double *)(base + h->layout.off_pval) : nullptr;
    if (cfg->plant) {
        const gw_pendulum_config &pc = cfg->pendulum;
        PendulumParams &Q = h->pend;
        Q.M = pc.cart_mass; Q.m = pc.pendulum_mass; Q.l = pc.arm_length; Q.g = pc.gravity;
        Q.fMax = pc.motor_fmax; Q.kServo = pc.motor_kservo; Q.dtMax = pc.dt_max;
        Q.kp = pc.kp; Q.ki = pc.ki; Q.kd = pc.kd; Q.vInit = pc.motor_v_init;
        Q.frequency = cfg->band[0].frequency_hz;
        Q.sensorY = cfg->band[0].device[0].y;
        Q.ctrlX = cfg->band[0].device[1].x; Q.ctrlY = cfg->band[0].device[1].y;
        Q.rrmX = cfg->band[0].device[2].x; Q.rrmY = cfg->band[0].device[2].y;
        Q.actuatorY = cfg->band[0].device[3].y;
        Q.mobility = pc.mobility;
    }
    for (int b = 0; b < cfg->n_bands; ++b) {
        const gw_band_config &cb = cfg->band[b];
        h->frequency[b] = cb.frequency_hz;
        h->thermal[b] = 1.38e-23 * (20.0 + 273.15) * cb.bandwidth_hz * 1000;     // physical.py:71, simple_stack.py:57,77
        for (int d = 0; d < cb.n_devices; ++d) {
            h->default_pos[b][d][0] = cb.device[d].x;
            h->default_pos[b][d][1] = cb.device[d].y;
            h->power_dbm[b][d] = cb.device[d].role == GW_ROLE_JAMMER ? cb.device[d].jam_power_dbm : 0.0;  // simple_stack.py:364,521
        }
    }
    cudaStream_t s = (cudaStream_t)stream;
    const long long nsim = h->st.nsim;
    const size_t aux = 8 * sizeof(double) + 4 * sizeof(int) + 2 * sizeof(unsigned long long);
    void *stg = nullptr;
    cudaError_t e = cudaMalloc((void **)&h->stats, aux);
    if (e == cudaSuccess) e = cudaMemsetAsync(h->stats, 0, aux, s);
    if (e == cudaSuccess) e = cudaMalloc(&stg, nsim * (4 + 4 + 8 + 8 + 1) + 64);
    if (e != cudaSuccess) { gw_destroy(h); return fail(GW_E_CUDA, "allocation failed: %s", cudaGetErrorString(e)); }
    if (!std::getenv("GYMWIPE_B200_NO_MEMO") && !cfg->per_env_positions) {
        // identical geometry in every env: a small table holds every (S, N) pair that occurs.  Per-env geometries
        // evaluate: a table that holds their ~20 recurring pairs per band-sim has to scale with the batch (512 MB
        // for configs[3]) and its random DRAM probes cost more than the evaluation (configs[3]: 1.56 ms per step
        // with it, 1.51 ms without)
        h->memo_entries = 1u << 14;
        e = cudaMalloc((void **)&h->memo, 32ull * (h->memo_entries + MEMO_L0));      // second + first level
        if (e == cudaSuccess) e = cudaMemsetAsync(h->memo, 0, 32ull * (h->memo_entries + MEMO_L0), s);
        if (e != cudaSuccess) { cudaFree(stg); gw_destroy(h); return fail(GW_E_CUDA, "allocation failed: %s", cudaGetErrorString(e)); }
    }
    if (!std::getenv("GYMWIPE_B200_NO_MEMO") && cfg->per_env_positions && !cfg->plant) {
        // per-env geometries: every band-sim caches the BER of its own links (1 KB per band-sim)
        e = cudaMalloc((void **)&h->memo_sim, 64ull * sizeof(ulonglong2) * (size_t)nsim);
        if (e == cudaSuccess) e = cudaMemsetAsync(h->memo_sim, 0, 64ull * sizeof(ulonglong2) * (size_t)nsim, s);
        if (e != cudaSuccess) { cudaFree(stg); gw_destroy(h); return fail(GW_E_CUDA, "allocation failed: %s", cudaGetErrorString(e)); }
    }
    h->errflag = (int *)(h->stats + 8);
    h->mask_bytes = (unsigned long long *)(h->errflag + 4);
    h->stats_use = h->stats;
    h->d_obs = (long long *)stg;                            // base of the staging allocation
    h->d_rew = (double *)(h->d_obs + nsim);
    int32_t *p32 = (int32_t *)(h->d_rew + nsim);
    h->d_dev = p32; h->d_dur = p32 + nsim;
    h->d_done = (unsigned char *)(p32 + 2 * nsim);
    SharedTables th;
    std::memset(&th, 0, sizeof th);
    for (int b = 0; b < cfg->n_bands; ++b) th.srx[b][0] = h->thermal[b];
#define CALL_INIT(DD, SS, JJ) init_kernel<DD, SS, JJ><<<grid_for(nsim, 128), 128, 0, s>>>(h->st, h->P, th)
    DISPATCH_SHAPE(h, CALL_INIT);
#undef CALL_INIT
    if (cfg->plant) pendulum_init_kernel<<<grid_for(nsim, 128), 128, 0, s>>>(h->st, h->pend);
    e = cudaGetLastError();
    if (e != cudaSuccess) { gw_destroy(h); return fail(GW_E_CUDA, "init kernel: %s", cudaGetErrorString(e)); }
    rc = launch_tables(h, nullptr, s);
    if (rc) { gw_destroy(h); return rc; }
    if (cfg->per_env_positions) {
        // current positions of every band-sim's devices (gw_set_positions moves them from here)
        const size_t cnt = (size_t)ntab * kMaxDev * 2;
        std::vector<double> init(cnt, 0.0);
        for (long long t = 0; t < ntab; ++t)
            for (int d = 0; d < kMaxDev; ++d) {
                init[(t * kMaxDev + d) * 2] = h->default_pos[t % cfg->n_bands][d][0];
                init[(t * kMaxDev + d) * 2 + 1] = h->default_pos[t % cfg->n_bands][d][1];
            }
        e = cudaMalloc((void **)&h->pos_cur, cnt * sizeof(double));
        if (e == cudaSuccess) e = cudaMemcpyAsync(h->pos_cur, init.data(), cnt * sizeof(double), cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);         // `init` is a local buffer
        if (e != cudaSuccess) { gw_destroy(h); return fail(GW_E_CUDA, "allocation failed: %s", cudaGetErrorString(e)); }
    }
    *out = h;
    return GW_OK;
}

void gw_destroy(gw_handle *h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->owned_state) cudaFree(h->owned_state);
    if (h->stats) cudaFree(h->stats);
    if (h->d_obs) cudaFree(h->d_obs);       // base of the staging allocation
    if (h->memo) cudaFree(h->memo);
    if (h->memo_sim) cudaFree(h->memo_sim);
    if (h->pos_cur) cudaFree(h->pos_cur);
    if (h->mask_t16) cudaFree(h->mask_t16);
    delete h;
}

int gw_set_positions(gw_handle *h, const double *positions, void *stream)
{
    if (!h) return fail(GW_E_INVALID, "handle is NULL");
    if (positions && !h->st.per_env) return fail(GW_E_INVALID, "handle was created with per_env_positions = 0");
    h->generation = next_generation();
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    const size_t pos_bytes = sizeof(double) * (size_t)h->st.ntab * kMaxDev * 2;
    if (!positions || h->cfg.plant || !h->stepped) {
        // placement: the devices are CREATED at these positions (fresh attenuation models)
        const int rc = launch_tables(h, positions, s);
        if (rc) return rc;
        if (positions && h->pos_cur) CUDA_TRY(cudaMemcpyAsync(h->pos_cur, positions, pos_bytes, cudaMemcpyDeviceToDevice, s));
        return GW_OK;
    }
    // the envs have been stepped: the devices MOVE (Position.set); transmissions on the air see
    // SimplePhy._onAttenuationChange
    SharedTables pw, fr;
    std::memset(&pw, 0, sizeof pw); std::memset(&fr, 0, sizeof fr);
    for (int b = 0; b < h->cfg.n_bands; ++b) {
        for (int d = 0; d < kMaxDev; ++d) pw.srx[b][d] = h->power_dbm[b][d];
        fr.srx[b][0] = h->frequency[b];
    }
    MaskSource ms;
    ms.mode = h->cfg.mode; ms.seed = h->cfg.seed; ms.env_offset = h->cfg.env_id_offset;
    ms.words = h->masks; ms.slots = h->mask_slots > 0 ? h->mask_slots : 1; ms.words_per_row = h->mask_words;
    ms.t16 = nullptr; ms.s32 = nullptr; ms.gt = 0; ms.nsb = 0;
    if (h->cfg.mode == GW_MODE_MASK_FED && !h->masks) return fail(GW_E_INVALID, "mode MASK_FED: call gw_set_masks first");
    const long long nsim = h->st.nsim;
#define CALL_MOVE(DD, SS, JJ)                                                                                              \
    do {                                                                                                                   \
        if (h->cfg.mode == GW_MODE_REFERENCE)                                                                              \
            move_kernel<MODE_R, DD, SS, JJ><<<grid_for(nsim, 64), 64, 0, s>>>(h->st, h->P, ms, positions, h->pos_cur, pw, fr, h->errflag);        \
        else if (h->cfg.mode == GW_MODE_MASK_PHILOX)                                                                       \
            move_kernel<MODE_M_PHILOX, DD, SS, JJ><<<grid_for(nsim, 64), 64, 0, s>>>(h->st, h->P, ms, positions, h->pos_cur, pw, fr, h->errflag); \
        else                                                                                                               \
            move_kernel<MODE_M_FED, DD, SS, JJ><<<grid_for(nsim, 64), 64, 0, s>>>(h->st, h->P, ms, positions, h->pos_cur, pw, fr, h->errflag);    \
    } while (0)
    DISPATCH_SHAPE(h, CALL_MOVE);
#undef CALL_MOVE
    CUDA_TRY(cudaGetLastError());
    return GW_OK;
}

int gw_reset(gw_handle *h, const int64_t *env_ids, int64_t n, int64_t *obs, void *stream)
{
    if (!h) return fail(GW_E_INVALID, "handle is NULL");
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    const long long total = (env_ids ? n : h->st.nenv) * h->st.nb;
    if (total <= 0) return GW_OK;
#define CALL_RESET(DD, SS, JJ) reset_kernel<DD, SS, JJ><<<grid_for(total, 128), 128, 0, s>>>(h->st, h->P, (const long long *)env_ids, (long long)n, (long long *)obs)
    DISPATCH_SHAPE(h, CALL_RESET);
#undef CALL_RESET
    CUDA_TRY(cudaGetLastError());
    return GW_OK;
}

static int launch_step(gw_handle *h, const int32_t *device, const int32_t *duration, int64_t *obs, double *reward,
                       uint8_t *done, cudaStream_t s, int *obs32 = nullptr, float *reward32 = nullptr,
                       double *trace = nullptr, int *trace_count = nullptr, int trace_cap = 0,
                       const unsigned char *act8 = nullptr, unsigned *res32 = nullptr,
                       long long sim_begin = 0, long long sim_end = -1, int tiny = 0)
{
    if (h->cfg.mode == GW_MODE_MASK_FED && !h->masks) return fail(GW_E_INVALID, "mode MASK_FED: call gw_set_masks first");
    h->stepped = 1;
    StepArgs A;
    A.st = h->st;
    A.device = device; A.duration = duration;
    A.obs = (long long *)obs; A.reward = reward; A.done = done;
    A.stats = h->stats_use; A.errflag = h->errflag; A.maskBytes = h->mask_bytes;
    A.stamps = (h->stamps && h->stamp_next < h->stamp_cap) ? h->stamps + 4 * h->stamp_next++ : nullptr;
    A.obs32 = obs32; A.reward32 = reward32;
    A.act8 = act8; A.res32 = res32; A.tiny = tiny;
    A.sim_begin = sim_begin; A.sim_end = sim_end < 0 ? h->st.nsim : sim_end;
    A.trace = trace; A.traceCount = trace_count; A.traceCap = trace_cap;
    A.masks.mode = h->cfg.mode; A.masks.seed = h->cfg.seed; A.masks.env_offset = h->cfg.env_id_offset;
    A.masks.words = h->masks; A.masks.slots = h->mask_slots > 0 ? h->mask_slots : 1; A.masks.words_per_row = h->mask_words;
    A.masks.t16 = h->mask_t16; A.masks.s32 = h->mask_s32; A.masks.gt = h->mask_gt; A.masks.nsb = h->mask_nsb;
    A.memo.tab = h->memo; A.memo.mask = h->memo_entries ? h->memo_entries - 1 : 0;
    A.memo.l0g = (h->memo && !h->st.per_env) ? h->memo + 2ull * h->memo_entries : nullptr;
    A.memo.l0s = nullptr;
    A.memo.sim = h->memo_sim; A.memo.stride = h->st.nsim;
    SharedTables T;
    std::memset(&T, 0, sizeof T);
    if (!h->st.per_env) {
        for (int b = 0; b < h->cfg.n_bands; ++b)
            for (int p = 0; p < kMaxDev; ++p)
                for (int d = 0; d < kMaxDev; ++d) {
                    if (p == d || p >= h->cfg.band[b].n_devices || d >= h->cfg.band[b].n_devices) continue;
                    // host evaluation of the same formulas; the device table (tables_kernel) is
                    // what per-env positions use
                    const double att = fspl_db(h->default_pos[b][p][0], h->default_pos[b][p][1],
                                               h->default_pos[b][d][0], h->default_pos[b][d][1], h->frequency[b]);
                    T.srx[b][p * kMaxDev + d] = rx_power_mw(h->power_dbm[b][d], att);
                }
    }
    const long long nsim = h->st.nsim;
    if (h->cfg.plant) {
        constexpr int psmem = Sim<4, 2, 1, ShStore>::kDirectBytes * STEP_BLOCK;
        if (!(h->smem_configured & (1u << 10))) {
            CUDA_TRY(cudaFuncSetAttribute(pendulum_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, psmem));
            h->smem_configured |= 1u << 10;
        }
        pendulum_step_kernel<<<grid_for(nsim, STEP_BLOCK), STEP_BLOCK, psmem, s>>>(A, h->P, h->pend);
        CUDA_TRY(cudaGetLastError());
        return GW_OK;
    }
    // one wave: 128-thread blocks, a multiple of the SM count when the batch is large
    int blocks = grid_for(A.sim_end - A.sim_begin, STEP_BLOCK);
    const int cap = 148 * GW_STEP_MIN_BLOCKS * 4;
    if (blocks > cap) blocks = cap;
#define LAUNCH_STEP(KERNEL, VARIANT, DD, SS, JJ)                                                     \
    do {                                                                                             \
        constexpr int smem = Sim<DD, SS, JJ, ShStore>::kDirectBytes * STEP_BLOCK;                     \
        if (!(h->smem_configured & (1u << (VARIANT)))) {                                             \
            CUDA_TRY(cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
            h->smem_configured |= 1u << (VARIANT);                                                   \
        }                                                                                            \
        cudaLaunchConfig_t lc = {};                                                                  \
        lc.gridDim = dim3((unsigned)blocks); lc.blockDim = dim3(STEP_BLOCK);                         \
        lc.dynamicSmemBytes = smem; lc.stream = s;                                                   \
        cudaLaunchAttribute at[1];                                                                   \
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                               \
        at[0].val.programmaticStreamSerializationAllowed = 1;                                        \
        lc.attrs = at; lc.numAttrs = h->pdl ? 1 : 0;                                                 \
        CUDA_TRY(cudaLaunchKernelEx(&lc, KERNEL, A, h->P, T));                                       \
    } while (0)
#define CALL_STEP(DD, SS, JJ)                                                                        \
    do {                                                                                             \
        if (h->ext) {                                                                                \
            if (trace) LAUNCH_STEP((step_kernel<MODE_R, DD, SS, JJ, true, true>), 4, DD, SS, JJ);     \
            else if (h->cfg.mode == GW_MODE_REFERENCE) LAUNCH_STEP((step_kernel<MODE_R, DD, SS, JJ, false, true>), 5, DD, SS, JJ);  \
            else if (h->cfg.mode == GW_MODE_MASK_PHILOX) LAUNCH_STEP((step_kernel<MODE_M_PHILOX, DD, SS, JJ, false, true>), 6, DD, SS, JJ); \
            else if (h->mask_t16) LAUNCH_STEP((step_kernel<MODE_M_FEDX, DD, SS, JJ, false, true>), 9, DD, SS, JJ); \
            else LAUNCH_STEP((step_kernel<MODE_M_FED, DD, SS, JJ, false, true>), 7, DD, SS, JJ);      \
        } else if (trace) LAUNCH_STEP((step_kernel<MODE_R, DD, SS, JJ, true>), 0, DD, SS, JJ);        \
        else if (h->cfg.mode == GW_MODE_REFERENCE) LAUNCH_STEP((step_kernel<MODE_R, DD, SS, JJ>), 1, DD, SS, JJ);  \
        else if (h->cfg.mode == GW_MODE_MASK_PHILOX) LAUNCH_STEP((step_kernel<MODE_M_PHILOX, DD, SS, JJ>), 2, DD, SS, JJ); \
        else if (h->mask_t16) LAUNCH_STEP((step_kernel<MODE_M_FEDX, DD, SS, JJ>), 8, DD, SS, JJ);    \
        else LAUNCH_STEP((step_kernel<MODE_M_FED, DD, SS, JJ>), 3, DD, SS, JJ);                       \
    } while (0)
    DISPATCH_SHAPE(h, CALL_STEP);
#undef CALL_STEP
    CUDA_TRY(cudaGetLastError());
    return GW_OK;
}

int gw_step(gw_handle *h, const int32_t *device, const int32_t *duration, int64_t *obs, double *reward,
            uint8_t *done, void *stream)
{
    if (!h) return fail(GW_E_INVALID, "handle is NULL");
    if (!device || !duration || !obs || !reward || !done) return fail(GW_E_INVALID, "NULL buffer");
    CUDA_TRY(cudaSetDevice(h->device));
    return launch_step(h, device, duration, obs, reward, done, (cudaStream_t)stream);
}

int gw_step_traced(gw_handle *h, const int32_t *device, const int32_t *duration, int64_t *obs, double *reward,
                   uint8_t *done, double *trace, int32_t *trace_count, int32_t cap, void *stream)
{
    if (!h) return fail(GW_E_INVALID, "handle is NULL");
    if (!device || !duration || !obs || !reward || !done || !trace || !trace_count || cap < 1) return fail(GW_E_INVALID, "bad argument");
    if (h->cfg.mode != GW_MODE_REFERENCE || h->cfg.plant) return fail(GW_E_INVALID, "tracing is available in GW_MODE_REFERENCE without a plant");
    CUDA_TRY(cudaSetDevice(h->device));
    return launch_step(h, device, duration, obs, reward, done, (cudaStream_t)stream, nullptr, nullptr, trace, trace_count, cap);
}

int gw_step_host(gw_handle *h, const int32_t *device, const int32_t *duration, int64_t *obs, double *reward,
                 uint8_t *done, void *stream)
{
    if (!h) return fail(GW_E_INVALID, "handle is NULL");
    if (!device || !duration || !obs || !reward || !done) return fail(GW_E_INVALID, "NULL buffer");
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    const long long n = h->st.nsim;
    CUDA_TRY(cudaMemcpyAsync(h->d_dev, device, n * 4, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(h->d_dur, duration, n * 4, cudaMemcpyHostToDevice, s));
    const int rc = launch_step(h, h->d_dev, h->d_dur, (int64_t *)h->d_obs, h->d_rew, h->d_done, s);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(obs, h->d_obs, n * 8, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(reward, h->d_rew, n * 8, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(done, h->d_done, n, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return GW_OK;
}

int gw_step_host_packed(gw_handle *h, const int32_t *actions, void *results, void *stream)
{
    if (!h) return fail(GW_E_INVALID, "handle is NULL");
    if (!actions || !results) return fail(GW_E_INVALID, "NULL buffer");
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    const long long n = h->st.nsim;
    // staging: d_dev and d_dur are adjacent (actions [2][n]); the compact results reuse the
    // obs / reward staging area: int32 obs[n] | float reward[n] | uint8 done[n]
    CUDA_TRY(cudaMemcpyAsync(h->d_dev, actions, n * 8, cudaMemcpyHostToDevice, s));
    int *obs32 = (int *)h->d_obs;
    float *rew32 = (float *)(obs32 + n);
    unsigned char *done8 = (unsigned char *)(rew32 + n);
    const int rc = launch_step(h, h->d_dev, h->d_dur, nullptr, nullptr, done8, s, obs32, rew32);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(results, obs32, n * 9, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return GW_OK;
}

// device-visible alias of a host pointer if it is pinned (cudaHostAlloc / cudaHostRegister) memory
static void *mapped_alias(const void *p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return at.type == cudaMemoryTypeHost ? at.devicePointer : nullptr;
}

static int step_host_compact(gw_handle *h, const uint8_t *actions, void *results, void *stream, bool sync, int tiny = 0);

int gw_step_host_compact(gw_handle *h, const uint8_t *actions, uint32_t *results, void *stream)
{
    return step_host_compact(h, actions, results, stream, true);
}

int gw_step_host_compact_async(gw_handle *h, const uint8_t *actions, uint32_t *results, void *stream)
{
    return step_host_compact(h, actions, results, stream, false);
}

static int step_host_compact(gw_handle *h, const uint8_t *actions, void *results, void *stream, bool sync, int tiny)
{
    if (!h) return fail(GW_E_INVALID, "handle is NULL");
    if (!actions || !results) return fail(GW_E_INVALID, "NULL buffer");
    if (h->cfg.plant) return fail(GW_E_INVALID, "gw_step_host_compact is not available for plant envs");
    if (h->cfg.max_assign_duration > (tiny ? 128 : 256))
        return fail(GW_E_INVALID, "max_assign_duration does not fit the %s action format", tiny ? "1-byte" : "uint8");
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    const long long n = h->st.nsim;
    const size_t abytes = (size_t)(tiny ? n : 2 * n), rbytes = (size_t)(tiny ? 2 * n : 4 * n);
    // Pinned host buffers are mapped into the device's address space (UVA): the step kernel reads the
    // actions and writes the result words over the host link itself -- 2 + 4 (tiny: 1 + 2) bytes per sim in
    // full 64 / 128-byte warp transactions -- and no copy is enqueued at all.  Pageable buffers are staged.
    const unsigned char *m_act = (const unsigned char *)mapped_alias(actions);
    unsigned *m_res = (unsigned *)mapped_alias(results);
    if (m_act && m_res) {
        const int rc = launch_step(h, nullptr, nullptr, nullptr, nullptr, nullptr, s, nullptr, nullptr, nullptr, nullptr, 0,
                                   m_act, m_res, 0, -1, tiny);
        if (rc) return rc;
        if (sync) CUDA_TRY(cudaStreamSynchronize(s));
        return GW_OK;
    }
    if (!sync) return fail(GW_E_INVALID, "the asynchronous / population calls need pinned host buffers");
    // staging: the actions in the action staging area, the result words in the obs area
    unsigned char *d_act = (unsigned char *)h->d_dev;
    unsigned *d_res = (unsigned *)h->d_obs;
    CUDA_TRY(cudaMemcpyAsync(d_act, actions, abytes, cudaMemcpyHostToDevice, s));
    const int rc = launch_step(h, nullptr, nullptr, nullptr, nullptr, nullptr, s, nullptr, nullptr, nullptr, nullptr, 0,
                               d_act, d_res, 0, -1, tiny);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(results, d_res, rbytes, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return GW_OK;
}

// side streams of gw_step_host_compact_many, per device (created on first use, never destroyed: they live
// as long as the library)
constexpr int MANY_STREAMS = 4;
struct ManyGraph {
    unsigned long long key;
    cudaGraphExec_t exec;
    unsigned long long used;
};
struct ManyStreams {
    cudaStream_t s[MANY_STREAMS];
    cudaStream_t main;                  // stands in for the legacy default stream (which cannot be captured)
    cudaEvent_t fork, join[MANY_STREAMS], enter;
    bool ready;
    std::vector<ManyGraph> graphs;      // captured population steps (gw_step_host_compact_many), keyed by the call
    unsigned long long clock;
    std::mutex call;                    // one population call per device at a time: the side streams, the fork / join
                                        // events and the graph cache are per device (concurrent callers would
                                        // contend for the same GPU anyway)
};
static ManyStreams g_many[64];
static std::mutex g_many_mutex;

static int step_host_many(gw_handle *const *handles, int32_t n_handles, const uint8_t *const *actions,
                          void *const *results, void *stream, int tiny);

int gw_step_host_compact_many(gw_handle *const *handles, int32_t n_handles, const uint8_t *const *actions,
                              uint32_t *const *results, void *stream)
{
    return step_host_many(handles, n_handles, actions, (void *const *)results, stream, 0);
}

int gw_step_host_tiny(gw_handle *h, const uint8_t *actions, uint16_t *results, void *stream)
{
    return step_host_compact(h, actions, results, stream, true, 1);
}

int gw_step_host_tiny_many(gw_handle *const *handles, int32_t n_handles, const uint8_t *const *actions,
                           uint16_t *const *results, void *stream)
{
    return step_host_many(handles, n_handles, actions, (void *const *)results, stream, 1);
}

static int step_host_many(gw_handle *const *handles, int32_t n_handles, const uint8_t *const *actions,
                          void *const *results, void *stream, int tiny)
{
    if (!handles || !actions || !results || n_handles < 1) return fail(GW_E_INVALID, "bad argument");
    for (int k = 0; k < n_handles; ++k) {
        if (!handles[k]) return fail(GW_E_INVALID, "handle %d is NULL", k);
        if (handles[k]->device != handles[0]->device) return fail(GW_E_INVALID, "the handles of one call live on one device");
    }
    const int dev = handles[0]->device;
    if (dev < 0 || dev >= 64) return fail(GW_E_INVALID, "device ordinal %d out of range", dev);
    CUDA_TRY(cudaSetDevice(dev));
    cudaStream_t s = (cudaStream_t)stream;
    // The batches are independent (own state, own pinned buffers): their steps are spread over a few side
    // streams, so that one batch's result words drain over the host link while the next batch computes --
    // on one stream every kernel would wait for its predecessor's posted writes.  Fork from / join into
    // the caller's stream with events; ONE synchronisation at the end.
    ManyStreams *ms;
    {
        std::lock_guard<std::mutex> lock(g_many_mutex);
        ms = &g_many[dev];
        if (!ms->ready) {
            CUDA_TRY(cudaEventCreateWithFlags(&ms->fork, cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&ms->enter, cudaEventDisableTiming));
            CUDA_TRY(cudaStreamCreateWithFlags(&ms->main, cudaStreamNonBlocking));
            for (int k = 0; k < MANY_STREAMS; ++k) {
                CUDA_TRY(cudaStreamCreateWithFlags(&ms->s[k], cudaStreamNonBlocking));
                CUDA_TRY(cudaEventCreateWithFlags(&ms->join[k], cudaEventDisableTiming));
            }
            ms->ready = true;
        }
    }
    const int lanes = n_handles < MANY_STREAMS ? n_handles : MANY_STREAMS;
    std::lock_guard<std::mutex> call_lock(ms->call);
    if (s == nullptr) {
        // the legacy default stream: the call runs on a stream of the library's own, ordered behind what the
        // default stream holds; it returns synchronised, so later work on the default stream is ordered behind it
        CUDA_TRY(cudaEventRecord(ms->enter, nullptr));
        CUDA_TRY(cudaStreamWaitEvent(ms->main, ms->enter, 0));
        s = ms->main;
    }
    // enqueues the fork / per-batch launches / join on `s` and the side streams
    auto enqueue = [&]() -> int {
        CUDA_TRY(cudaEventRecord(ms->fork, s));
        for (int k = 0; k < lanes; ++k) CUDA_TRY(cudaStreamWaitEvent(ms->s[k], ms->fork, 0));
        for (int k = 0; k < n_handles; ++k) {
            const int rc = step_host_compact(handles[k], actions[k], results[k], (void *)ms->s[k % lanes], false, tiny);
            if (rc) return rc;
        }
        for (int k = 0; k < lanes; ++k) {
            CUDA_TRY(cudaEventRecord(ms->join[k], ms->s[k]));
            CUDA_TRY(cudaStreamWaitEvent(s, ms->join[k], 0));
        }
        return GW_OK;
    };
    // The same call (same handles, same pinned buffers, nothing about the handles changed) is usually repeated
    // step after step: its ~16 kernel launches and event operations are captured ONCE into a CUDA graph and
    // replayed with one cudaGraphLaunch -- the host-side launch cost per population step drops from ~16 launches
    // to one.  GW_MANY_GRAPH=0 in the environment: plain launches.  Handles with launch stamps are not cached.
    static const bool use_graph = [] { const char *e = std::getenv("GW_MANY_GRAPH"); return !(e && e[0] == '0'); }();
    bool cacheable = use_graph;
    unsigned long long key = 1469598103934665603ull;
    auto mix = [&](unsigned long long v) { key = (key ^ v) * 1099511628211ull; };
    mix((unsigned long long)n_handles); mix((unsigned long long)(uintptr_t)s); mix((unsigned long long)tiny);
    for (int k = 0; k < n_handles; ++k) {
        if (handles[k]->stamps) cacheable = false;
        mix((unsigned long long)(uintptr_t)handles[k]); mix(handles[k]->generation);
        mix((unsigned long long)(uintptr_t)actions[k]); mix((unsigned long long)(uintptr_t)results[k]);
    }
    if (cacheable) {
        cudaGraphExec_t exec = nullptr;
        {
            std::lock_guard<std::mutex> lock(g_many_mutex);
            for (auto &g : ms->graphs) if (g.key == key) { exec = g.exec; g.used = ++ms->clock; break; }
        }
        if (!exec) {
            for (int k = 0; k < n_handles; ++k) {
                // validation and the one-time shared-memory attribute happen outside the capture
                if (!actions[k] || !results[k]) return fail(GW_E_INVALID, "NULL buffer");
                if (!mapped_alias(actions[k]) || !mapped_alias(results[k])) { cacheable = false; break; }
            }
        }
        if (cacheable && !exec) {
            cudaGraph_t graph = nullptr;
            CUDA_TRY(cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed));
            const int rc = enqueue();
            const cudaError_t ce = cudaStreamEndCapture(s, &graph);
            if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
            CUDA_TRY(ce);
            const cudaError_t ci = cudaGraphInstantiate(&exec, graph, 0);
            cudaGraphDestroy(graph);
            CUDA_TRY(ci);
            std::lock_guard<std::mutex> lock(g_many_mutex);
            if (ms->graphs.size() >= 32) {             // evict the least recently used entry
                size_t victim = 0;
                for (size_t q = 1; q < ms->graphs.size(); ++q) if (ms->graphs[q].used < ms->graphs[victim].used) victim = q;
                cudaGraphExecDestroy(ms->graphs[victim].exec);
                ms->graphs.erase(ms->graphs.begin() + (long)victim);
            }
            ms->graphs.push_back(ManyGraph{key, exec, ++ms->clock});
        }
        if (cacheable) {
            for (int k = 0; k < n_handles; ++k) handles[k]->stepped = 1;
            CUDA_TRY(cudaGraphLaunch(exec, s));
            CUDA_TRY(cudaStreamSynchronize(s));
            return GW_OK;
        }
    }
    const int rc = enqueue();
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(s));
    return GW_OK;
}

int gw_check(gw_handle *h, void *stream)
{
    if (!h) return fail(GW_E_INVALID, "handle is NULL");
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    int flag[4] = {0, 0, 0, 0};
    CUDA_TRY(cudaMemcpyAsync(flag, h->errflag, sizeof flag, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    if (flag[0] == 0) return GW_OK;
    CUDA_TRY(cudaMemsetAsync(h->errflag, 0, sizeof flag, s));
    if (flag[0] == GW_E_ACTION)
        return fail(GW_E_ACTION, "action outside the action space (first at sim %d); the reference asserts "
                                 "action_space.contains(action)", flag[1]);
    return fail(GW_E_SIMFAULT, "sim %d hit a condition under which the reference raises (%s)", flag[1],
                flag[2] == FAULT_REF_KEYERROR ? "KeyError in SimplePhy._updateBitErrorRate, SURVEY app. B #12"
                : flag[2] == FAULT_REF_ASSERT ? "assert noisePower >= 0"
                : flag[2] == FAULT_SENDQ ? "SEND queue overflow" : "internal");
}

int gw_stats(gw_handle *h, double *out8, int clear, void *stream)
{
    if (!h || !out8) return fail(GW_E_INVALID, "NULL argument");
    CUDA_TRY(cudaSetDevice(h->device));
    stats_copy_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(h->stats_use, out8, clear);
    CUDA_TRY(cudaGetLastError());
    return GW_OK;
}

int gw_mask_bytes(gw_handle *h, uint64_t *out, int clear, void *stream)
{
    if (!h || !out) return fail(GW_E_INVALID, "NULL argument");
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    CUDA_TRY(cudaMemcpyAsync(out, h->mask_bytes, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    if (clear) CUDA_TRY(cudaMemsetAsync(h->mask_bytes, 0, sizeof(uint64_t), s));
    return GW_OK;
}

int gw_debug_stamps(gw_handle *h, uint64_t *stamps, int64_t capacity)
{
    if (h) h->generation = next_generation();
    if (!h) return fail(GW_E_INVALID, "handle is NULL");
    h->stamps = (unsigned long long *)stamps;
    h->stamp_cap = stamps ? capacity : 0;
    h->stamp_next = 0;
    return GW_OK;
}

int gw_share_stats(gw_handle *h, gw_handle *with)
{
    if (!h) return fail(GW_E_INVALID, "handle is NULL");
    if (with && with->device != h->device) return fail(GW_E_INVALID, "handles on different devices cannot share statistics");
    h->stats_use = with ? with->stats_use : h->stats;
    h->generation = next_generation();
    return GW_OK;
}

int gw_read_state(gw_handle *h, int field, double *out, void *stream)
{
    if (!h || !out) return fail(GW_E_INVALID, "NULL argument");
    if (field < GW_FIELD_NOW || field > GW_FIELD_N_RECEIVED) return fail(GW_E_INVALID, "unknown field %d", field);
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    const long long nsim = h->st.nsim;
    if (field == GW_FIELD_PLANT) {
        if (!h->cfg.plant) return fail(GW_E_INVALID, "not a plant env");
        plant_read_kernel<<<grid_for(nsim, 128), 128, 0, s>>>(h->st, out);
        CUDA_TRY(cudaGetLastError());
        return GW_OK;
    }
#define CALL_READ(DD, SS, JJ) read_kernel<DD, SS, JJ><<<grid_for(nsim, 128), 128, 0, s>>>(h->st, h->P, field, out)
    DISPATCH_SHAPE(h, CALL_READ);
#undef CALL_READ
    CUDA_TRY(cudaGetLastError());
    return GW_OK;
}

int gw_set_masks(gw_handle *h, const uint32_t *mask_words, int32_t slots, int32_t words_per_row, void *stream)
{
    if (!h) return fail(GW_E_INVALID, "handle is NULL");
    if (!mask_words || slots < 1 || words_per_row < 4 || (words_per_row & 3)) return fail(GW_E_INVALID, "bad mask layout (words_per_row must be a multiple of 4)");
    if (((uintptr_t)mask_words & 15) != 0) return fail(GW_E_INVALID, "mask buffer must be 16-byte aligned");
    h->masks = mask_words; h->mask_slots = slots; h->mask_words = words_per_row;
    h->generation = next_generation();
    // prefix-count index (mask_index_kernel): one streaming pass over the buffer on `stream`; GW_FED_INDEX=0 in the
    // environment keeps the step kernels scanning the mask words themselves (step_kernel<MODE_M_FED>)
    const char *no_index = std::getenv("GW_FED_INDEX");
    const bool want_index = !(no_index && no_index[0] == '0');
    CUDA_TRY(cudaSetDevice(h->device));
    if (!want_index || h->cfg.mode != GW_MODE_MASK_FED) {
        if (h->mask_t16) { CUDA_TRY(cudaFree(h->mask_t16)); h->mask_t16 = nullptr; h->mask_s32 = nullptr; h->mask_idx_bytes = 0; }
        return GW_OK;
    }
    const long long rows = h->st.nsim * kMaxDev * (long long)slots * kMaxDev;
    const int G = words_per_row >> 2;
    const int gt = (G + 1 + 7) & ~7;                // entries per row: G + 1, padded to 16 bytes
    const int nsb = (G >> 9) + 1;
    const size_t t16_bytes = align_up((size_t)rows * gt * sizeof(uint16_t), 256);
    const size_t s32_bytes = nsb > 1 ? (size_t)rows * nsb * sizeof(uint32_t) : 0;
    if (h->mask_idx_bytes != t16_bytes + s32_bytes || !h->mask_t16) {
        if (h->mask_t16) { CUDA_TRY(cudaFree(h->mask_t16)); h->mask_t16 = nullptr; h->mask_s32 = nullptr; h->mask_idx_bytes = 0; }
        void *p = nullptr;
        CUDA_TRY(cudaMalloc(&p, t16_bytes + s32_bytes));
        h->mask_t16 = (uint16_t *)p;
        h->mask_idx_bytes = t16_bytes + s32_bytes;
    }
    h->mask_s32 = nsb > 1 ? (uint32_t *)((char *)h->mask_t16 + t16_bytes) : nullptr;
    h->mask_gt = gt; h->mask_nsb = nsb;
    long long blocks = (rows + 7) / 8;              // 8 warps per block, one row per warp and pass
    const long long cap = 148ll * 8 * 4;
    if (blocks > cap) blocks = cap;
    mask_index_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(mask_words, words_per_row, rows, h->mask_t16, h->mask_s32, gt, nsb);
    CUDA_TRY(cudaGetLastError());
    return GW_OK;
}

int gw_mask_index_count(gw_handle *h, const int64_t *rows, const int32_t *k0, const int32_t *k1, int32_t *counts, int64_t n,
                        void *stream)
{
    if (!h) return fail(GW_E_INVALID, "handle is NULL");
    if (!h->mask_t16) return fail(GW_E_INVALID, "no mask index: mode MASK_FED and gw_set_masks first (GW_FED_INDEX=0 disables it)");
    if (n <= 0) return GW_OK;
    if (!rows || !k0 || !k1 || !counts) return fail(GW_E_INVALID, "NULL buffer");
    CUDA_TRY(cudaSetDevice(h->device));
    MaskSource m;
    m.mode = MODE_M_FED; m.seed = 0; m.env_offset = 0;
    m.words = h->masks; m.slots = h->mask_slots; m.words_per_row = h->mask_words;
    m.t16 = h->mask_t16; m.s32 = h->mask_s32; m.gt = h->mask_gt; m.nsb = h->mask_nsb;
    mask_index_count_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(m, (const long long *)rows, k0, k1, counts, n);
    CUDA_TRY(cudaGetLastError());
    return GW_OK;
}

int gw_fspl_attenuation(const double *ax, const double *ay, const double *bx, const double *by, double frequency_hz,
                        double *att_db, int64_t n, void *stream)
{
    if (n <= 0) return GW_OK;
    fspl_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(ax, ay, bx, by, frequency_hz, att_db, n);
    CUDA_TRY(cudaGetLastError());
    return GW_OK;
}

int gw_ber_bpsk(const double *signal_mw, const double *noise_mw, double *ber, int64_t n, void *stream)
{
    if (n <= 0) return GW_OK;
    const double c = 10 * std::log10(133.33333e3), q = 1.135 * std::sqrt(2 * 3.141592653589793);
    ber_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(signal_mw, noise_mw, ber, n, c, q);
    CUDA_TRY(cudaGetLastError());
    return GW_OK;
}

int gw_count_bit_errors(const uint32_t *mask_words, int32_t words_per_row, const int64_t *rows, const int32_t *k0,
                        const int32_t *k1, int32_t *counts, int64_t n, void *stream)
{
    if (n <= 0) return GW_OK;
    if (words_per_row < 4 || (words_per_row & 3) || ((uintptr_t)mask_words & 15)) return fail(GW_E_INVALID, "bad mask layout");
    cudaStream_t st = (cudaStream_t)stream;
    // launch geometry per DEVICE (the shared-memory attribute and the occupancy are per device), guarded:
    // this entry point has no handle, so several host threads may come through here at once
    struct K3Geometry { int sms, occ_reg, occ_tma; };
    static K3Geometry geo[64];
    static std::mutex geo_mutex;
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return fail(GW_E_INVALID, "device ordinal %d out of range", dev);
    int sms, occ_reg, occ_tma;
    {
        std::lock_guard<std::mutex> lock(geo_mutex);
        K3Geometry &g = geo[dev];
        if (g.sms == 0) {
            cudaDeviceGetAttribute(&g.sms, cudaDevAttrMultiProcessorCount, dev);
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g.occ_reg, count_bits_kernel, 256, 0) != cudaSuccess || g.occ_reg < 1) g.occ_reg = 6;
            cudaFuncSetAttribute(count_bits_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TMA_SMEM_BYTES);
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g.occ_tma, count_bits_tma_kernel, 256, TMA_SMEM_BYTES) != cudaSuccess || g.occ_tma < 1) g.occ_tma = 3;
            if (g.sms < 1) g.sms = 148;
        }
        sms = g.sms; occ_reg = g.occ_reg; occ_tma = g.occ_tma;
    }
    // persistent grid: one full wave (SM count x resident blocks per SM), 8 warps per block.
    // Default: the register-staged kernel (measured 5.97 TB/s = 91 % of the measured HBM peak).
    // GYMWIPE_B200_K3=tma selects the cp.async.bulk / mbarrier variant (measured 4.47 TB/s: the
    // single issuing lane and the per-stage barrier round trip cost more than the registers save).
    static const char *force = std::getenv("GYMWIPE_B200_K3");
    const bool use_tma = words_per_row * 4 <= tma::STAGE_BYTES && force && force[0] == 't';
    long long blocks = (n + 7) / 8;
    const long long wave = (long long)sms * (use_tma ? occ_tma : occ_reg);
    if (blocks > wave) blocks = wave;
    if (use_tma)
        count_bits_tma_kernel<<<(int)blocks, 256, TMA_SMEM_BYTES, st>>>(mask_words, words_per_row, (const long long *)rows, k0, k1, counts, n);
    else
        count_bits_kernel<<<(int)blocks, 256, 0, st>>>(mask_words, words_per_row, (const long long *)rows, k0, k1, counts, n);
    CUDA_TRY(cudaGetLastError());
    return GW_OK;
}

#ifdef GW_MEMO_STATS
int gw_debug_memo_stats(unsigned long long *out2)
{
    CUDA_TRY(cudaMemcpyFromSymbol(out2, g_memo_stats, sizeof(unsigned long long) * 2));
    return GW_OK;
}
#endif

int gw_policy_boltzmann(const float *weights, int32_t n_actions, int32_t n_durations, const int64_t *obs, int64_t n,
                        float obs_center, double tau, double clip_lo, double clip_hi, uint64_t seed, uint64_t counter,
                        int64_t env_id_offset, int64_t *flat_action, int32_t *device, int32_t *duration, double *probs,
                        void *stream)
{
    if (n <= 0) return GW_OK;
    if (!weights || !obs) return fail(GW_E_INVALID, "NULL buffer");
    if ((device == nullptr) != (duration == nullptr)) return fail(GW_E_INVALID, "device and duration come together");
    if (n_durations < 1 || !(tau > 0)) return fail(GW_E_INVALID, "bad policy parameters");
    const int grid = grid_for(n, 128);
    cudaStream_t s = (cudaStream_t)stream;
#define CALL_POLICY(AA) policy_kernel<AA><<<grid, 128, 0, s>>>(weights, (const long long *)obs, n, obs_center, tau, clip_lo, clip_hi, \
        seed, counter, env_id_offset, n_durations, (long long *)flat_action, device, duration, probs)
    if (n_actions == 40) CALL_POLICY(40);               // 2 devices x 20 durations (envs/core.py:39-42)
    else if (n_actions == 20) CALL_POLICY(20);
    else if (n_actions == 8) CALL_POLICY(8);
    else if (n_actions >= 1 && n_actions <= POLICY_MAX_ACTIONS_N)
        policy_kernel_n<<<grid, 128, 0, s>>>(weights, n_actions, (const long long *)obs, n, obs_center, tau, clip_lo, clip_hi, seed, counter,
                                             env_id_offset, n_durations, (long long *)flat_action, device, duration, probs);
    else return fail(GW_E_INVALID, "the policy kernels take 1..%d actions, not %d", POLICY_MAX_ACTIONS_N, n_actions);
#undef CALL_POLICY
    CUDA_TRY(cudaGetLastError());
    return GW_OK;
}

int gw_grid_create(const gw_grid_config *cfg, int device, const double *positions, const double *delays,
                   const double *move_delays, const double *offsets, void *stream, gw_grid_handle **out)
{
    if (!cfg || !out || !positions || !delays) return fail(GW_E_INVALID, "NULL argument");
    if (cfg->abi_version != GW_ABI_VERSION) return fail(GW_E_INVALID, "abi_version %d != %d", cfg->abi_version, GW_ABI_VERSION);
    if (cfg->n_envs < 1) return fail(GW_E_INVALID, "n_envs must be >= 1");
    if (cfg->n_devices < 1 || cfg->n_devices > GW_GRID_MAX_DEVICES) return fail(GW_E_INVALID, "n_devices must be in 1..%d", GW_GRID_MAX_DEVICES);
    if (cfg->max_moves < 0 || (cfg->max_moves > 0 && (!move_delays || !offsets || !(cfg->move_interval > 0))))
        return fail(GW_E_INVALID, "mobility needs move_delays, offsets and move_interval > 0");
    for (int d = 0; d < cfg->n_devices; ++d) {
        if (!(cfg->send_interval[d] > 0)) return fail(GW_E_INVALID, "send_interval must be > 0");
        if (cfg->header_bytes[d] < 1 || cfg->payload_bytes[d] < 1) return fail(GW_E_INVALID, "bad packet size");
    }
    int ndev = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (ndev < 1) return fail(GW_E_CUDA, "no CUDA device: gymwipe_b200 has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(GW_E_INVALID, "device %d out of range (%d devices)", device, ndev);
    CUDA_TRY(cudaSetDevice(device));
    gw_grid_handle *h = new (std::nothrow) gw_grid_handle();
    if (!h) return fail(GW_E_INVALID, "out of host memory");
    std::memset(h, 0, sizeof *h);
    h->cfg = *cfg;
    h->device = device;
    GridParams &G = h->G;
    G.ndev = cfg->n_devices; G.maxMoves = cfg->max_moves; G.moveInterval = cfg->move_interval;
    G.bitRate = 133.33333e3; G.dataRate = 0.75 * G.bitRate; G.maxBer = gw_max_correctable_ber(3, 4);
    G.tenLog10BitRate = 10 * std::log10(G.bitRate); G.qDen = 1.135 * std::sqrt(2 * 3.141592653589793);
    G.bitsFactor = 2 - 0.75; G.frequency = cfg->frequency_hz; G.fsplConst = 20 * std::log10(cfg->frequency_hz);
    G.thermal = 1.38e-23 * (20.0 + 273.15) * cfg->bandwidth_hz * 1000;
    for (int d = 0; d < cfg->n_devices; ++d) {
        G.power[d] = cfg->power_dbm[d]; G.interval[d] = cfg->send_interval[d];
        G.hdrBytes[d] = cfg->header_bytes[d]; G.payBytes[d] = cfg->payload_bytes[d];
    }
    h->A.block_bytes = (grid_state_bytes(cfg->n_devices) + 15) / 16 * 16;
    h->A.n_envs = cfg->n_envs;
    h->A.offsets = cfg->max_moves > 0 ? offsets : nullptr;
    cudaError_t e = cudaMalloc((void **)&h->A.state, h->A.block_bytes * (size_t)cfg->n_envs);
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->errflag, 4 * sizeof(int));
    if (e == cudaSuccess) e = cudaMemsetAsync(h->errflag, 0, 4 * sizeof(int), (cudaStream_t)stream);
    if (e != cudaSuccess) { gw_grid_destroy(h); return fail(GW_E_CUDA, "allocation failed: %s", cudaGetErrorString(e)); }
    h->A.errflag = h->errflag;
    grid_init_kernel<<<grid_for(cfg->n_envs, 64), 64, 0, (cudaStream_t)stream>>>(h->A, h->G, positions, delays,
                                                                                 cfg->max_moves > 0 ? move_delays : nullptr);
    e = cudaGetLastError();
    if (e != cudaSuccess) { gw_grid_destroy(h); return fail(GW_E_CUDA, "init kernel: %s", cudaGetErrorString(e)); }
    *out = h;
    return GW_OK;
}

void gw_grid_destroy(gw_grid_handle *h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->A.state) cudaFree(h->A.state);
    if (h->errflag) cudaFree(h->errflag);
    delete h;
}

static int grid_launch(gw_grid_handle *h, double duration, double *trace, int32_t *trace_count, int32_t cap, void *stream)
{
    if (!h) return fail(GW_E_INVALID, "handle is NULL");
    if (!(duration > 0)) return fail(GW_E_INVALID, "duration must be > 0 (simpy: until must be greater than now)");
    CUDA_TRY(cudaSetDevice(h->device));
    GridArgs A = h->A;
    A.trace = trace; A.trace_count = trace_count; A.trace_cap = cap;
    // small batches: one warp per block, so that the warps spread over more SMs (4,096 grids = 128 warps)
    const int block = A.n_envs <= 148LL * 64 ? 32 : 64;
    grid_run_kernel<<<grid_for(A.n_envs, block), block, 0, (cudaStream_t)stream>>>(A, h->G, duration);
    CUDA_TRY(cudaGetLastError());
    return GW_OK;
}

int gw_grid_run(gw_grid_handle *h, double duration, void *stream)
{
    return grid_launch(h, duration, nullptr, nullptr, 0, stream);
}

int gw_grid_run_traced(gw_grid_handle *h, double duration, double *trace, int32_t *trace_count, int32_t cap, void *stream)
{
    if (!trace || !trace_count || cap < 1) return fail(GW_E_INVALID, "bad trace buffer");
    return grid_launch(h, duration, trace, trace_count, cap, stream);
}

int gw_grid_read(gw_grid_handle *h, int field, double *out, void *stream)
{
    if (!h || !out) return fail(GW_E_INVALID, "NULL argument");
    if (field < GW_GRID_FIELD_NOW || field > GW_GRID_FIELD_FAULT) return fail(GW_E_INVALID, "unknown field %d", field);
    CUDA_TRY(cudaSetDevice(h->device));
    grid_read_kernel<<<grid_for(h->A.n_envs, 64), 64, 0, (cudaStream_t)stream>>>(h->A, h->G, field, out);
    CUDA_TRY(cudaGetLastError());
    return GW_OK;
}

int gw_grid_check(gw_grid_handle *h, void *stream)
{
    if (!h) return fail(GW_E_INVALID, "handle is NULL");
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    int flag[4] = {0, 0, 0, 0};
    CUDA_TRY(cudaMemcpyAsync(flag, h->errflag, sizeof flag, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    if (flag[0] == 0) return GW_OK;
    CUDA_TRY(cudaMemsetAsync(h->errflag, 0, sizeof flag, s));
    return fail(GW_E_SIMFAULT, "grid env %d hit a condition under which the reference raises (fault %d)", flag[1], flag[2]);
}

int gw_genband_create(const gw_genband_config *cfg, int device, const double *positions, void *stream, gw_genband_handle **out)
{
    if (!cfg || !out || !positions) return fail(GW_E_INVALID, "NULL argument");
    if (cfg->abi_version != GW_ABI_VERSION) return fail(GW_E_INVALID, "abi_version %d != %d", cfg->abi_version, GW_ABI_VERSION);
    if (cfg->n_envs < 1) return fail(GW_E_INVALID, "n_envs must be >= 1");
    const int ns = cfg->n_senders, nj = cfg->n_phy_senders;
    if (ns < 2 || ns > GW_GENBAND_MAX_SENDERS) return fail(GW_E_INVALID, "n_senders must be in 2..%d", GW_GENBAND_MAX_SENDERS);
    if (nj < 0 || nj > GW_GENBAND_MAX_PHY_SENDERS) return fail(GW_E_INVALID, "n_phy_senders must be in 0..%d", GW_GENBAND_MAX_PHY_SENDERS);
    if (cfg->assignment_duration_factor < 1 || cfg->max_assign_duration < 1) return fail(GW_E_INVALID, "bad duration parameters");
    if (cfg->mode != GW_MODE_REFERENCE && cfg->mode != GW_MODE_MASK_PHILOX)
        return fail(GW_E_INVALID, "the general band engine offers GW_MODE_REFERENCE and GW_MODE_MASK_PHILOX");
    if (!(cfg->frequency_hz > 0) || !(cfg->bandwidth_hz > 0)) return fail(GW_E_INVALID, "bad frequency band");
    for (int k = 0; k < ns; ++k) {
        if (cfg->multiplicity[k] < 1 || cfg->multiplicity[k] > 1000) return fail(GW_E_INVALID, "sender %d: multiplicity must be in 1..1000", k);
        if (!(cfg->interval[k] > 0)) return fail(GW_E_INVALID, "sender %d: interval must be > 0", k);
        if (cfg->payload_bytes[k] > 60000) return fail(GW_E_INVALID, "sender %d: payload_bytes must be <= 60000", k);
        if (cfg->destination[k] < 0 || cfg->destination[k] >= ns || cfg->destination[k] == k)
            return fail(GW_E_INVALID, "sender %d: destination must be another sender", k);
        if (cfg->max_ticks[k] < 0) return fail(GW_E_INVALID, "sender %d: max_ticks must be >= 0", k);
    }
    for (int j = 0; j < nj; ++j) {
        if (!(cfg->phy_interval[j] > 0) || !(cfg->phy_delay[j] >= 0)) return fail(GW_E_INVALID, "PHY-only sender %d: bad interval / delay", j);
        if (cfg->phy_header_bytes[j] < 1 || cfg->phy_payload_bytes[j] < 1 || cfg->phy_header_bytes[j] + cfg->phy_payload_bytes[j] > 60000)
            return fail(GW_E_INVALID, "PHY-only sender %d: bad packet size", j);
    }
    int ndev = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (ndev < 1) return fail(GW_E_CUDA, "no CUDA device: gymwipe_b200 has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(GW_E_INVALID, "device %d out of range (%d devices)", device, ndev);
    CUDA_TRY(cudaSetDevice(device));
    gw_genband_handle *h = new (std::nothrow) gw_genband_handle();
    if (!h) return fail(GW_E_INVALID, "out of host memory");
    std::memset(h, 0, sizeof *h);
    h->cfg = *cfg;
    h->device = device;
    Params &P = h->P;
    P.nbands = 1; P.factor = cfg->assignment_duration_factor; P.maxDuration = cfg->max_assign_duration; P.mode = cfg->mode;
    P.bitRate = 133.33333e3; P.dataRate = 0.75 * P.bitRate; P.maxBer = gw_max_correctable_ber(3, 4);
    P.tenLog10BitRate = 10 * std::log10(P.bitRate); P.qDen = 1.135 * std::sqrt(2 * 3.141592653589793);
    P.bitsFactor = 2 - 0.75;
    finish_params(P);
    GenBand &B = h->B;
    B.ns = ns; B.nj = nj; B.nd = ns + 1 + nj; B.maxDuration = cfg->max_assign_duration;
    B.mode = cfg->mode == GW_MODE_MASK_PHILOX ? MODE_M_PHILOX : MODE_R; B.seed = cfg->seed; B.envOffset = cfg->env_id_offset;
    B.thermal = 1.38e-23 * (20.0 + 273.15) * cfg->bandwidth_hz * 1000;        // simple_stack.py:57, physical.py:61-78
    B.frequency = cfg->frequency_hz; B.maxMoves = 0; B.moveInterval = 0.0;
    double power[GW_GENBAND_MAX_DEVICES];
    for (int d = 0; d < B.nd; ++d) power[d] = 0.0;                              // MACs and the RRM send at 0 dBm
    for (int k = 0; k < ns; ++k) {
        B.mult[k] = cfg->multiplicity[k]; B.payloadRule[k] = cfg->payload_bytes[k] < 0 ? -1 : cfg->payload_bytes[k];
        B.dest[k] = cfg->destination[k]; B.maxTicks[k] = cfg->max_ticks[k]; B.recv[k] = cfg->receive[k] ? 1 : 0;
        B.interval[k] = cfg->interval[k];
    }
    for (int j = 0; j < nj; ++j) {
        B.jamInterval[j] = cfg->phy_interval[j]; B.jamDelay[j] = cfg->phy_delay[j];
        B.jamHdr[j] = cfg->phy_header_bytes[j]; B.jamPay[j] = cfg->phy_payload_bytes[j];
        power[ns + 1 + j] = cfg->phy_power_dbm[j];
    }
    for (int d = 0; d < B.nd; ++d) B.power[d] = power[d];
    GenArgs &A = h->A;
    A.n_envs = cfg->n_envs; A.per_env = cfg->per_env_positions ? 1 : 0;
    const size_t n = (size_t)cfg->n_envs;
    double *&dpower = h->dpower;
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMalloc((void **)&A.f64, sizeof(double) * gen_f64_words(ns, nj) * n);
    if (e == cudaSuccess) e = cudaMalloc((void **)&A.i32, sizeof(int32_t) * gen_i32_words(ns, nj) * n);
    if (e == cudaSuccess) e = cudaMalloc((void **)&A.srx, sizeof(double) * B.nd * B.nd * (A.per_env ? n : 1));
    if (e == cudaSuccess && A.per_env) e = cudaMalloc((void **)&A.att, sizeof(double) * B.nd * B.nd * n);
    if (e == cudaSuccess && A.per_env) e = cudaMalloc((void **)&A.pos, sizeof(double) * 2 * B.nd * n);
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->errflag, 4 * sizeof(int));
    if (e == cudaSuccess) e = cudaMemsetAsync(h->errflag, 0, 4 * sizeof(int), s);
    if (e == cudaSuccess) e = cudaMalloc((void **)&dpower, sizeof power);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dpower, power, sizeof power, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess && !A.per_env) {
        // The incrementally kept power sum of a PHY leaves residues of the size of an ulp of the STRONGEST signal it
        // has seen, so a 1-ulp difference in one table entry moves a later noise power -- and the
        // BER of a weak link -- by up to 1e-7 relative: the shared table is computed with the host's libm.
        double hpos[2 * GW_GENBAND_MAX_DEVICES], tab[GW_GENBAND_MAX_DEVICES * GW_GENBAND_MAX_DEVICES];
        e = cudaMemcpyAsync(hpos, positions, sizeof(double) * 2 * B.nd, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e == cudaSuccess) {
            gen_power_table(B.nd, hpos, power, cfg->frequency_hz, tab, 1);
            e = cudaMemcpyAsync(A.srx, tab, sizeof(double) * B.nd * B.nd, cudaMemcpyHostToDevice, s);
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);                     // `tab` is a stack array
    }
    if (e == cudaSuccess) {
        A.errflag = h->errflag;
        genband_init_kernel<<<grid_for(cfg->n_envs, 64), 64, 0, s>>>(A, B, positions, dpower, cfg->frequency_hz);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);                         // `power` is a stack array
    if (e != cudaSuccess) { gw_genband_destroy(h); return fail(GW_E_CUDA, "general band engine: %s", cudaGetErrorString(e)); }
    *out = h;
    return GW_OK;
}

void gw_genband_destroy(gw_genband_handle *h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->A.f64) cudaFree(h->A.f64);
    if (h->A.i32) cudaFree(h->A.i32);
    if (h->A.srx) cudaFree(h->A.srx);
    if (h->A.att) cudaFree(h->A.att);
    if (h->A.pos) cudaFree(h->A.pos);
    if (h->A.mvT) cudaFree(h->A.mvT);
    if (h->A.mvI) cudaFree(h->A.mvI);
    if (h->dpower) cudaFree(h->dpower);
    if (h->errflag) cudaFree(h->errflag);
    delete h;
}

int gw_genband_set_positions(gw_genband_handle *h, const double *positions, void *stream)
{
    if (!h || !positions) return fail(GW_E_INVALID, "NULL argument");
    if (!h->A.per_env) return fail(GW_E_INVALID, "handle was created with per_env_positions = 0");
    CUDA_TRY(cudaSetDevice(h->device));
    genband_move_kernel<<<grid_for(h->A.n_envs, 64), 64, 0, (cudaStream_t)stream>>>(h->A, h->P, h->B, positions);
    CUDA_TRY(cudaGetLastError());
    return GW_OK;
}

int gw_genband_set_movers(gw_genband_handle *h, const double *move_delays, const double *offsets, int32_t max_moves,
                          double move_interval, void *stream)
{
    if (!h || !move_delays || !offsets) return fail(GW_E_INVALID, "NULL argument");
    if (!h->A.per_env) return fail(GW_E_INVALID, "handle was created with per_env_positions = 0");
    if (max_moves < 1 || !(move_interval > 0)) return fail(GW_E_INVALID, "max_moves must be >= 1 and move_interval > 0");
    if (h->A.mvT) return fail(GW_E_INVALID, "the mobility processes of this handle are running already");
    CUDA_TRY(cudaSetDevice(h->device));
    const size_t n = (size_t)h->A.n_envs;
    cudaError_t e = cudaMalloc((void **)&h->A.mvT, sizeof(double) * 2 * h->B.nd * n);
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->A.mvI, sizeof(int32_t) * 3 * h->B.nd * n);
    if (e != cudaSuccess) return fail(GW_E_CUDA, "allocation failed: %s", cudaGetErrorString(e));
    h->A.offsets = offsets;
    h->B.maxMoves = max_moves; h->B.moveInterval = move_interval;
    genband_movers_kernel<<<grid_for(h->A.n_envs, 64), 64, 0, (cudaStream_t)stream>>>(h->A, h->B, move_delays);
    CUDA_TRY(cudaGetLastError());
    return GW_OK;
}

int gw_genband_reset(gw_genband_handle *h, int64_t *obs, void *stream)
{
    if (!h) return fail(GW_E_INVALID, "handle is NULL");
    CUDA_TRY(cudaSetDevice(h->device));
    genband_reset_kernel<<<grid_for(h->A.n_envs, 64), 64, 0, (cudaStream_t)stream>>>(h->A, h->B, (long long *)obs);
    CUDA_TRY(cudaGetLastError());
    return GW_OK;
}

static int genband_launch(gw_genband_handle *h, const int32_t *device, const int32_t *duration, int64_t *obs, double *reward,
                          uint8_t *done, double *trace, int32_t *trace_count, int32_t cap, void *stream)
{
    if (!h) return fail(GW_E_INVALID, "handle is NULL");
    if (!device || !duration || !obs || !reward || !done) return fail(GW_E_INVALID, "NULL buffer");
    CUDA_TRY(cudaSetDevice(h->device));
    GenArgs A = h->A;
    A.trace = trace; A.trace_count = trace_count; A.trace_cap = cap;
    const int grid = grid_for(A.n_envs, GW_GEN_BLOCK);
    if (h->B.mode == MODE_R)
        genband_step_kernel<false><<<grid, GW_GEN_BLOCK, 0, (cudaStream_t)stream>>>(A, h->P, h->B, device, duration, (long long *)obs, reward, done);
    else
        genband_step_kernel<true><<<grid, GW_GEN_BLOCK, 0, (cudaStream_t)stream>>>(A, h->P, h->B, device, duration, (long long *)obs, reward, done);
    CUDA_TRY(cudaGetLastError());
    return GW_OK;
}

int gw_genband_step(gw_genband_handle *h, const int32_t *device, const int32_t *duration, int64_t *obs, double *reward,
                    uint8_t *done, void *stream)
{
    return genband_launch(h, device, duration, obs, reward, done, nullptr, nullptr, 0, stream);
}

int gw_genband_step_traced(gw_genband_handle *h, const int32_t *device, const int32_t *duration, int64_t *obs, double *reward,
                           uint8_t *done, double *trace, int32_t *trace_count, int32_t cap, void *stream)
{
    if (!trace || !trace_count || cap < 1) return fail(GW_E_INVALID, "bad trace buffer");
    return genband_launch(h, device, duration, obs, reward, done, trace, trace_count, cap, stream);
}

int gw_genband_read(gw_genband_handle *h, int field, double *out, void *stream)
{
    if (!h || !out) return fail(GW_E_INVALID, "NULL argument");
    if (field < GW_GENBAND_FIELD_NOW || field > GW_GENBAND_FIELD_RECEIVED_VALUES) return fail(GW_E_INVALID, "unknown field %d", field);
    CUDA_TRY(cudaSetDevice(h->device));
    genband_read_kernel<<<grid_for(h->A.n_envs, 64), 64, 0, (cudaStream_t)stream>>>(h->A, h->B, field, out);
    CUDA_TRY(cudaGetLastError());
    return GW_OK;
}

int gw_genband_check(gw_genband_handle *h, void *stream)
{
    if (!h) return fail(GW_E_INVALID, "handle is NULL");
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    int flag[4] = {0, 0, 0, 0};
    CUDA_TRY(cudaMemcpyAsync(flag, h->errflag, sizeof flag, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    if (flag[0] == 0) return GW_OK;
    CUDA_TRY(cudaMemsetAsync(h->errflag, 0, sizeof flag, s));
    if (flag[0] == GW_E_ACTION) return fail(GW_E_ACTION, "env %d: action outside the action space", flag[1]);
    return fail(GW_E_SIMFAULT, "env %d hit a condition under which the reference raises (fault %d)", flag[1], flag[2]);
}

int gw_philox4x32(const uint32_t *counter, const uint32_t *key, uint32_t *out, int64_t n, void *stream)
{
    if (n <= 0) return GW_OK;
    philox_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(counter, key, out, n);
    CUDA_TRY(cudaGetLastError());
    return GW_OK;
}

}  // extern "C"
