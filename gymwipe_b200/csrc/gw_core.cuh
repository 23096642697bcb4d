// gymwipe_b200 -- per-env simulation core of the fused, event-ordered step kernel (K4).
//
// One "band-sim" (one frequency band of one env) is advanced by ONE GPU thread; a warp
// holds 32 band-sims.  The SimPy heap of the reference is replaced by a fixed set of
// TIMED event slots per band-sim (traffic ticks, jammer wake-ups, one PHY event per
// device, MAC window time-outs, the RRM time-out), ordered by (time, creation seq) --
// the reference's (time, priority, eid) order restricted to timed events -- while every
// zero-delay event chain of the reference (Initialize / succeed() / process-end events)
// is executed inline, in the order SimPy would pop it.  DESIGN.md section 4 derives the
// reduction; tests/ check it bit-exactly against the literal restatement in oracle/.
//
// Reference semantics implemented here (file:line under /root/reference):
//   SimplePhy  (power bookkeeping, BER accounting, decider)  networking/simple_stack.py:77-286
//   SimpleMac  (queue, assignment-window loop)                networking/simple_stack.py:386-471
//   SimpleRrmMac (announcement, guard slot)                   networking/simple_stack.py:527-561
//   Transmission / FrequencyBand.transmit                     networking/physical.py:224-290,576-608
//   BpskMcs / Eb-N0 / Q-function / dB helpers                 networking/physical.py:25-98,187-212
//   FsplAttenuation / Position.distanceTo                     networking/attenuation_models.py:28-36, devices/core.py:88-95
//   SenderDevice.senderProcess, CounterTrafficInterpreter     envs/counter_traffic.py:53-61,63-112
//   SimMan.nextTimeSlot / timeoutUntil                        simtools.py:44-53,103-116
//
// This header is plain C++ (no CUDA intrinsics): the kernels in gw_kernels.cu include it
// for the device, and tests/hostsim compiles the very same code for the host so that the
// event logic is validated against the oracle without a GPU.  Floating point: every
// operation is IEEE fp64 in the reference's order; compile with -fmad=false (nvcc) /
// -ffp-contract=off (gcc) so that no multiply-add is contracted.
#pragma once

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define GW_HD __host__ __device__ __forceinline__
#define GW_HD_COLD __host__ __device__ __noinline__
#define GW_UNROLL _Pragma("unroll")
// Loops over the devices of a band inside the generic transition function.  Fully unrolled where the per-device
// arrays live in registers (host build, plant kernel: a run-time index would push them to local memory); ROLLED
// (unroll factor 1) in the step kernels, whose per-device arrays are in shared memory.  The kernels of bands with
// interferers diverge into every branch of the transition function, and their unrolled code (8 k SASS instructions
// = 125 KB against a 32 KB instruction cache) left the warps waiting for instruction fetch 65 % of the time (ncu:
// stall_no_instruction 17.6 per issued instruction on configs[3]): rolled, configs[3] runs 1.77x faster
// (2.77 -> 1.57 ms), configs[2] 1.10x, the default kernel 1.02x.  Needs D, NS, NJ, ST in scope.
#define GW_UNROLL_D _Pragma("unroll (Sim<D, NS, NJ, ST>::kUnrollD)")
#else
#define GW_HD inline
#define GW_HD_COLD inline
#define GW_UNROLL
#define GW_UNROLL_D
#endif

// test hook of the host build (tests/hostsim): counts how often each macro event applies
#ifndef GW_STAT_MACRO
#define GW_STAT_MACRO(i) ((void)0)
#endif

namespace gw {

constexpr double kSlot = 1e-6;          // TIME_SLOT_LENGTH, simple_stack.py:27
constexpr int kMacHdr = 13;             // SimpleMacHeader byteSize, messages.py:154
constexpr int kNetHdr = 12;             // SimpleNetworkHeader byteSize, messages.py:180
constexpr int kQueueCap = 100;          // deque(maxlen=100), simple_stack.py:361
constexpr int kRingSlots = 128;         // slots of the snapshot / value rings (power of two >= kQueueCap)
constexpr int kCounterBound = 65536;    // COUNTER_BOUND, counter_traffic.py:35
constexpr int kCounterByteLen = 2;      // COUNTER_BYTE_LENGTH, counter_traffic.py:33
constexpr int kMaxAirBytes = 70000;     // largest packet on the path: 13 + 12 + COUNTER_BOUND bytes (and fixed payloads <= 60000)

enum : int { S_IDLE = 0, S_WAITRX = 1, S_SLOT = 2, S_HDR = 3, S_PAY = 4 };
enum : int { MAC_NONE = 0, MAC_WAIT_COND = 1, MAC_WAIT_TX = 2, MAC_IDLE = 3 };
enum : int { EV_NONE = 0, EV_TICK = 1, EV_JAM = 2, EV_PHY = 3, EV_W = 4, EV_RRM = 5, EV_RXTO = 6 };
constexpr double kReceiveTimeout = 100.0;      // SimpleNetworkDevice.RECEIVE_TIMEOUT (s), devices.py:66
enum : int { MODE_R = 0, MODE_M_PHILOX = 1, MODE_M_FED = 2 };
enum : int { FAULT_NONE = 0, FAULT_REF_KEYERROR = 2, FAULT_REF_ASSERT = 3, FAULT_SENDQ = 4,
             FAULT_EMPTY = 6 };

constexpr int kMaxBands = 4, kMaxDev = 4, kMaxSend = 2, kMaxJam = 1;

// Scenario constants shared by all envs (kernel parameter).
struct BandParams {
    int ndev, ns, nj;
    int mult[kMaxSend];
    int payloadRule[kMaxSend];          // -1: byteSize = counter
    double interval[kMaxSend];
    int maxTicks[kMaxSend];             // 0: the traffic process runs forever (reference); n: a burst of n ticks
    int recv[kMaxSend];                 // 1: MAC receive mode (SimpleNetworkDevice.receiving = True, devices.py:70-97)
    double jamInterval[kMaxJam], jamDelay[kMaxJam];
    int jamHdr[kMaxJam], jamPay[kMaxJam];
};

struct Params {
    int nbands, factor, maxDuration, mode;
    double bitRate;                     // 133.33333e3, physical.py:196
    double dataRate;                    // 0.75 * bitRate, physical.py:197
    double maxBer;                      // Mcs.maxCorrectableBer(), physical.py:160-185
    double tenLog10BitRate;             // 10*log10(bitRate), host libm
    double qDen;                        // 1.135 * sqrt(2*pi), physical.py:44,58
    double bitsFactor;                  // float(2 - codeRate) = 1.25, physical.py:259-263
    int noMacro;                        // macro events: 0 = where they pay (bands without interferers), 1 = never, -1 = always
    double berMult;                     // 1 / maxBer if that is a power of two and bit counts are integers, else 0
    double airtime[32];                 // (k * 8) / dataRate for k < 32 bytes (same IEEE division, done once)
    double rateInv;                     // fl(1 / dataRate) if the multiply-and-correct quotient below was verified, else 0
    BandParams band[kMaxBands];
};

// ---------------------------------------------------------------------------
// physical-layer arithmetic
// ---------------------------------------------------------------------------

// BpskMcs.calculateBitErrorRate on powers in mW (simple_stack.py:166-172, physical.py:208-212,
// 25-42, 46-58).  Operation order follows the Python source literally.
GW_HD double ber_bpsk_mw(double S, double N, double tenLog10BitRate, double qDen)
{
    const double sd = 10 * log10(S);            // milliwattsToDbm, physical.py:82-89
    const double nd = 10 * log10(N);
    if (sd <= nd) return 0.5;
    const double ratio_db = sd - nd - tenLog10BitRate;
#if defined(__CUDA_ARCH__)
    // Device: the three pow() calls are most of the ~10^3 instructions of an evaluation (configs[3] with per-env
    // geometries evaluates one per power change: 39 % of that kernel's instructions).  10**y = exp10(y), and
    // math.e**y with math.e = fl(e) = e (1 + d), d = -5.3e-17, is exp(y) (1 + y d) to first order (|y| < 750:
    // second order < 1e-27) -- the same values within the 1-2 ulp by which CUDA's and the C library's pow()
    // differ anyway (the oracle and the host build keep pow; parity tolerance on BER values: 1e-9).
    const double ratio = exp10(ratio_db / 10);
    const double x = sqrt(2 * ratio);
    const double kD = -5.3182377066058912e-17;          // fl(e) / e - 1 = ln(fl(e)) - 1
    const double y1 = -1.4 * x, y2 = -(x * x / 2);
    return (1 - exp(y1) * fma(y1, kD, 1.0)) * (exp(y2) * fma(y2, kD, 1.0)) / (qDen * x);
#else
    const double ratio = pow(10.0, ratio_db / 10);
    const double x = sqrt(2 * ratio);
    const double e = 2.718281828459045;         // math.e
    return (1 - pow(e, -1.4 * x)) * pow(e, -(x * x / 2)) / (qDen * x);
#endif
}

// Out-of-line copy for the step kernels: with the BER memo the evaluation is a cold path, and
// inlining its ~10^3 instructions at every call site would blow the instruction cache.  A leaf
// function with scalar arguments can be called without forcing the state struct into memory.
GW_HD_COLD double ber_bpsk_mw_cold(double S, double N, double tenLog10BitRate, double qDen)
{
    return ber_bpsk_mw(S, N, tenLog10BitRate, qDen);
}

// FsplAttenuation._update (attenuation_models.py:28-36); equal positions keep 0 dB.
GW_HD double fspl_db(double ax, double ay, double bx, double by, double frequency)
{
    if (ax == bx && ay == by) return 0.0;
    const double dx = ax - bx, dy = ay - by;
    const double d = sqrt(dx * dx + dy * dy);   // Position.distanceTo, devices/core.py:88-95
    return 20 * log10(d) + 20 * log10(frequency) - 147.55;
}

// dbmToMilliwatts(power - attenuation), simple_stack.py:111, physical.py:91-98
GW_HD double rx_power_mw(double power_dbm, double att_db) { return pow(10.0, (power_dbm - att_db) / 10); }

// ---------------------------------------------------------------------------
// Philox4x32-10 (Random123), counter-based; mode M error masks
// ---------------------------------------------------------------------------

GW_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                         uint32_t out[4])
{
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#if defined(__CUDACC__)
#pragma unroll
#endif
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// Key of the error flag of on-air bit k of transmission `txseq` of device `sender` on
// band `band` of GLOBAL env `env`, as seen by `receiver`: the Philox counter is
// (k>>2, txseq, sender | receiver<<8 | band<<16, env_lo), the key (seed_lo ^ env_hi, seed_hi);
// bit k uses output word k&3 and is an error iff word < floor(ber * 2^32).
GW_HD uint32_t ber_threshold(double ber) { return (uint32_t)(ber * 4294967296.0); }

GW_HD void mask_words4(uint64_t seed, int64_t env, int band, int sender, uint32_t txseq, int receiver,
                       uint32_t kq, uint32_t out[4])
{
    philox4x32_10(kq, txseq, (uint32_t)sender | ((uint32_t)receiver << 8) | ((uint32_t)band << 16),
                  (uint32_t)(uint64_t)env, (uint32_t)seed ^ (uint32_t)((uint64_t)env >> 32),
                  (uint32_t)(seed >> 32), out);
}

// serial reference of the per-range error count (host tests, tiny ranges); the kernels
// use the warp-cooperative version in gw_kernels.cu
GW_HD int64_t mask_errors_serial(uint64_t seed, int64_t env, int band, int sender, uint32_t txseq,
                                 int receiver, int64_t k0, int64_t k1, double ber)
{
    const uint32_t thr = ber_threshold(ber);
    int64_t n = 0;
    uint32_t w[4];
    int64_t cur = -1;
    for (int64_t k = k0; k < k1; ++k) {
        if ((k >> 2) != cur) { cur = k >> 2; mask_words4(seed, env, band, sender, txseq, receiver, (uint32_t)cur, w); }
        n += (w[k & 3] < thr) ? 1 : 0;
    }
    return n;
}

// ---------------------------------------------------------------------------
// storage of the per-device / per-sender state arrays
//
// The transition function indexes these arrays with RUN-TIME device indices.  Two storages:
//  * RegArr  -- a plain array that the compiler keeps in registers; run-time indices are resolved
//    by fully unrolled selects (get_at / set_at), because a dynamically indexed member would push
//    the whole state struct into local memory.  Used by the host build and for the fields of the
//    tick loop.
//  * a "direct" storage supplied by the kernels (shared memory, [field][index][thread]) where a
//    run-time index is just an address computation.
// ---------------------------------------------------------------------------

template <class T, int N>
struct RegArr {
    T v[N];
    using value_type = T;
    static constexpr int size = N;
    static constexpr bool direct = false;
    GW_HD T &operator[](int i) { return v[i]; }
    GW_HD const T &operator[](int i) const { return v[i]; }
};

struct RegStore {
    static constexpr bool roll = false;                     // arrays in registers: device loops stay unrolled
    template <class T, int N, int OFF> using Arr = RegArr<T, N>;
    template <class T, int N> using Aux = RegArr<T, N>;     // mode-M / plant-only arrays
};

// ---------------------------------------------------------------------------
// band-sim state
// ---------------------------------------------------------------------------

template <int D, int NS, int NJ, class ST = RegStore>
struct Sim {
    static constexpr int kD = D, kNS = NS, kNJ = NJ, kRrm = NS;
#if !defined(GW_NO_ROLL)
    static constexpr int kUnrollD = ST::roll ? 1 : D;                   // see GW_UNROLL_D
    static constexpr int kUnrollIO = (ST::roll && NJ > 0) ? 1 : D;      // state load / store / flag loops of the kernels
#else
    static constexpr int kUnrollIO = D;
    static constexpr int kUnrollD = D;
#endif
    static constexpr int NJa = NJ > 0 ? NJ : 1;
    // byte offsets (per thread) of the arrays inside a direct storage
    enum : int {
        O_P = 0, O_tEv = O_P + 8 * D, O_txStart = O_tEv + 8 * D, O_tStop = O_txStart + 8 * D, O_tC = O_tStop + 8 * D,
        O_ber = O_tC + 8 * D, O_err = O_ber + 8 * D, O_tReset = O_err + 8 * D, O_stopW = O_tReset + 8 * D,
        O_sphase = O_stopW + 8 * NS, O_sEv = O_sphase + 4 * D, O_sC = O_sEv + 4 * D, O_cmdPay = O_sC + 4 * D,
        O_txSeq = O_cmdPay + 4 * D, O_rxOf = O_txSeq + 4 * D, O_rxSec = O_rxOf + 4 * D,
        O_mac = O_rxSec + 4 * D, O_wDone = O_mac + 4 * NS, O_wPend = O_wDone + 4 * NS, O_sW = O_wPend + 4 * NS,
        O_nDeliv = O_sW + 4 * NS, kDirectBytes = O_nDeliv + 4 * NS
    };
    template <class T, int N, int OFF> using Arr = typename ST::template Arr<T, N, OFF>;
    template <class T, int N> using Aux = typename ST::template Aux<T, N>;

    double now;
    uint32_t seq;                       // creation counter of timed events (the eid order)
    int pchg;                           // bit p: PHY p's power sum changed since its BER was last evaluated
    int fault;
    uint32_t ties;                      // exact-time ties between independent events (diagnostic)

    // PHY, per device
    Arr<double, D, O_P> P;              // SimplePhy._receivedPower (mW), simple_stack.py:80-82
    Arr<int, D, O_sphase> sphase;       // S_*: macInHandler progress (simple_stack.py:192-212)
    Arr<double, D, O_tEv> tEv;          // next PHY event: slot start / header end / completion
    Arr<uint32_t, D, O_sEv> sEv;
    Arr<double, D, O_txStart> txStart;  // Transmission.startTime
    Arr<double, D, O_tStop> tStop;      // Transmission.stopTime
    Arr<double, D, O_tC> tC;            // time of eCompletes
    Arr<uint32_t, D, O_sC> sC;
    Arr<int, D, O_cmdPay> cmdPay;       // payload bytes of the SEND command in flight
    Aux<double, D> txVal;               // payload.value of the packet in flight (plant envs only)
    Arr<uint32_t, D, O_txSeq> txSeq;    // transmissions started by this device (mode M key)
    Arr<int, D, O_rxOf> rxOf;           // device whose transmission is being received, -1 idle
    Arr<int, D, O_rxSec> rxSec;         // 0 header, 1 payload
    Arr<double, D, O_ber> ber;          // _receivedBitErrorRate
    Arr<double, D, O_err> err;          // _receivedBitErrorSum
    Arr<double, D, O_tReset> tReset;    // _lastReceivedErrorCountTime
    Aux<double, D> segT0;               // mode M: start of the current constant-BER segment

    // senders: traffic process (registers: the tick loop indexes them statically) + MAC
    RegArr<double, NS> tTick;
    RegArr<uint32_t, NS> sTick;
    RegArr<uint64_t, NS> ticks;         // ticks fired; packets enqueued = ticks * mult
    RegArr<int, NS> qn;                 // queue length (<= 100)
    RegArr<uint64_t, NS> epochK;        // counter(tick k) = min(65536, epochC + (k - epochK)); packets enqueued before
                                        // tick epochK (index < epochK * mult) read their size from the snapshot ring
    RegArr<int, NS> epochC;
    Arr<int, NS, O_mac> mac;            // MAC_*
    Arr<int, NS, O_wDone> wDone;
    Arr<int, NS, O_wPend> wPend;
    Arr<double, NS, O_stopW> stopW;
    Arr<uint32_t, NS, O_sW> sW;
    // (MAC receive mode, simple_stack.py:436-460 / devices.py:88-97: the current RECEIVE command's timeout and the
    // count of packets handed to onReceive live behind the `ring` accessor -- rxT / rxS / set_rx / add_received --,
    // i.e. in memory, not in this struct: bands without receive mode hold no registers for them.  A MAC in
    // receive mode is receiving at every event boundary: a delivery or a timeout ends the command and the
    // device's receive loop re-issues it within the same instant.)

    // jammers
    RegArr<double, NJa> tJam;
    RegArr<uint32_t, NJa> sJam;
    RegArr<int, NJa> jamStage, jamPending;

    // RRM
    int annDest, annBytes;
    double annSlots;
    int rrmPend;
    double tRrm;
    uint32_t sRrm;
    int assignDone;

    // interpreter (counter_traffic.py:63-112)
    int rv0, rv1, latestDiff, lastAbsDiff, done;

    // statistics
    uint32_t nTx;
    Arr<uint32_t, NS, O_nDeliv> nDeliv;

    // optional event trace (traced step variant only; nullptr in the production kernels, where
    // every recording site folds away): records of 8 doubles {kind, t, dev, x0, x1, x2, x3, -}
    double *trace;
    int ntrace, traceCap;
};

enum : int { REC_TX = 1, REC_BER = 2, REC_DEC = 3, REC_RX = 4, REC_MRX = 5 };

template <int D, int NS, int NJ, class ST>
GW_HD void trace_rec(Sim<D, NS, NJ, ST> &s, int kind, double t, int dev, double x0, double x1, double x2, double x3)
{
    if (s.trace == nullptr) return;
    if (s.ntrace < s.traceCap) {
        double *r = s.trace + (long long)s.ntrace * 8;
        r[0] = kind; r[1] = t; r[2] = dev; r[3] = x0; r[4] = x1; r[5] = x2; r[6] = x3; r[7] = 0;
    }
    s.ntrace += 1;
}

struct Event {
    int kind, idx;
    double t;
    uint32_t seq;
};

// construction-time state (CounterTrafficEnv.__init__, counter_traffic.py:114-133)
template <int D, int NS, int NJ, class ST>
GW_HD void init_sim(Sim<D, NS, NJ, ST> &s, double thermal)
{
    s.now = 0.0; s.seq = 0; s.fault = 0; s.ties = 0; s.pchg = -1;
    GW_UNROLL
    for (int p = 0; p < D; ++p) {
        s.P[p] = thermal; s.sphase[p] = S_IDLE; s.tEv[p] = 0; s.sEv[p] = 0;
        s.txStart[p] = 0; s.tStop[p] = 0; s.tC[p] = 0; s.sC[p] = 0; s.cmdPay[p] = 0; s.txSeq[p] = 0; s.txVal[p] = 0;
        s.rxOf[p] = -1; s.rxSec[p] = 0; s.ber[p] = 0; s.err[p] = 0; s.tReset[p] = 0; s.segT0[p] = 0;
    }
    // process Initialize events in construction order: senders, then jammers (URGENT, t = 0)
    GW_UNROLL
    for (int k = 0; k < NS; ++k) {
        s.tTick[k] = 0.0; s.sTick[k] = s.seq++;
        s.ticks[k] = 0; s.qn[k] = 0; s.epochK[k] = 0; s.epochC[k] = 1;
        s.mac[k] = MAC_NONE; s.wDone[k] = 0; s.wPend[k] = 0; s.stopW[k] = 0; s.sW[k] = 0;
        s.nDeliv[k] = 0;
    }
    GW_UNROLL
    for (int j = 0; j < Sim<D, NS, NJ, ST>::NJa; ++j) {
        s.tJam[j] = 0.0; s.sJam[j] = (j < NJ) ? s.seq++ : 0; s.jamStage[j] = 0; s.jamPending[j] = 0;
    }
    s.annDest = 0; s.annBytes = 0; s.annSlots = 0; s.rrmPend = 0; s.tRrm = 0; s.sRrm = 0; s.assignDone = 0;
    s.rv0 = s.rv1 = 0; s.latestDiff = 0; s.lastAbsDiff = 0; s.done = 0;
    s.nTx = 0;
    s.trace = nullptr; s.ntrace = 0; s.traceCap = 0;
}

// `device.receiving = True` after construction (devices.py:77-84), senders in index order: each starts its receive
// loop as a process (Initialize, t = 0); the loop's first pass issues the RECEIVE command like every later one,
// so it is modelled as a time-out slot at t = 0 (EV_RXTO)
template <int D, int NS, int NJ, class ST, class Ring>
GW_HD void init_receive(Sim<D, NS, NJ, ST> &s, const BandParams &B, Ring &ring)
{
    GW_UNROLL
    for (int k = 0; k < NS; ++k) {
        if (Ring::ext && B.recv[k]) ring.set_rx(k, 0.0, s.seq++);
        else ring.set_rx(k, (double)INFINITY, 0u);
    }
}

GW_HD bool seq_before(uint32_t a, uint32_t b) { return (int32_t)(a - b) < 0; }

// element access with a RUN-TIME index through fully unrolled selects: keeps the per-sim
// state in registers (a dynamically indexed member array would force the whole struct
// into local memory)
template <class A, class V>
GW_HD void set_at(A &a, int i, V v)
{
    using T = typename A::value_type;
    if constexpr (A::direct) {
        a[i] = (T)v;
    } else {
        GW_UNROLL
        for (int q = 0; q < A::size; ++q) a[q] = (q == i) ? (T)v : a[q];
    }
}
template <class A>
GW_HD typename A::value_type get_at(const A &a, int i)
{
    using T = typename A::value_type;
    if constexpr (A::direct) {
        return a[i];
    } else {
        T r = a[0];
        GW_UNROLL
        for (int q = 1; q < A::size; ++q) r = (q == i) ? a[q] : r;
        return r;
    }
}
// Received-power table of a band-sim, entry (receiver p, sender d).  Two representations:
//  * a plain array [D * D] that the compiler keeps in registers (host build, plant envs): a run-time
//    sender index is resolved by selects;
//  * SrxView: a table in memory (the step kernels: shared memory for a geometry common to all
//    envs, global memory for per-env geometries), entry (p, d) at base[(p * kMaxDev + d) * stride]
//    -- no registers are held for it.
struct SrxView {
    const double *base;
    long long stride;
};
template <int D>
GW_HD double srx_at(const SrxView &x, int p, int d) { return x.base[(long long)(p * kMaxDev + d) * x.stride]; }

// received power of receiver p (compile-time after unrolling) from sender d (run time)
template <int D>
GW_HD double srx_at(const double *srx, int p, int d)
{
    double r = srx[p * D];
    GW_UNROLL
    for (int q = 1; q < D; ++q) r = (q == d) ? srx[p * D + q] : r;
    return r;
}

// ---------------------------------------------------------------------------
// event selection: earliest (time, seq) among the timed slots.
//
// Traffic ticks are ~80 % of all timed events and touch nothing but their sender's queue
// counters -- unless that sender's MAC is waiting for a packet.  next_event() therefore
// applies such "silent" ticks in a tight loop, in exact (time, seq) order, and only hands
// PHY / MAC / RRM / jammer events and MAC-waking ticks to the full transition function.
// ---------------------------------------------------------------------------

GW_HD bool before(double ta, uint32_t qa, double tb, uint32_t qb)
{
    return ta < tb || (ta == tb && seq_before(qa, qb));
}

template <int D, int NS, int NJ, class ST, class Ring>
GW_HD Event select_nontick(const Sim<D, NS, NJ, ST> &s, const BandParams &B, const Ring &ring)
{
    Event e;
    e.kind = EV_NONE; e.idx = 0; e.t = INFINITY; e.seq = 0;
#define GW_CONSIDER(T, SQ, K, I)                                                        \
    do {                                                                                \
        const double t_ = (T);                                                          \
        const uint32_t q_ = (SQ);                                                       \
        if (before(t_, q_, e.t, e.seq) || e.kind == EV_NONE) {                          \
            e.t = t_; e.seq = q_; e.kind = (K); e.idx = (I);                            \
        }                                                                               \
    } while (0)
    GW_UNROLL
    for (int j = 0; j < NJ; ++j) GW_CONSIDER(s.tJam[j], s.sJam[j], EV_JAM, j);
    GW_UNROLL_D
    for (int d = 0; d < D; ++d)
        if (s.sphase[d] >= S_SLOT) GW_CONSIDER(s.tEv[d], s.sEv[d], EV_PHY, d);
    GW_UNROLL
    for (int k = 0; k < NS; ++k)
        if (s.wPend[k]) GW_CONSIDER(s.stopW[k], s.sW[k], EV_W, k);
    if (s.rrmPend) GW_CONSIDER(s.tRrm, s.sRrm, EV_RRM, 0);
    GW_UNROLL
    for (int k = 0; k < NS; ++k)
        if (Ring::ext && B.recv[k]) GW_CONSIDER(ring.rxT(k), ring.rxS(k), EV_RXTO, k);
#undef GW_CONSIDER
    return e;
}

// Next event for the full transition function.  Silent ticks that precede it -- and lie
// strictly before `tLimit` -- are applied on the way (SenderDevice.senderProcess,
// counter_traffic.py:53-61: `mult` packets into the drop-oldest queue, counter += 1, next tick).
// Silent ticks of sender K strictly before (tEnd, qEnd) -- and before tLimit -- applied in one
// tight loop: the tick times are accumulated with the reference's fp64 additions, one per tick
// (counter_traffic.py:61), everything else is counted and applied once.  Ticks of different senders
// touch only their own sender's queue counters, so the senders are advanced one after the other;
// the creation numbers they consume only matter for exact-time ties with other events (`ties`).
template <int K, int D, int NS, int NJ, class ST>
GW_HD void silent_ticks(Sim<D, NS, NJ, ST> &s, int mult, double interval, double tEnd, uint32_t qEnd, bool haveEnd,
                        double tLimit)
{
    double t = s.tTick[K];
    if (!(t < tLimit)) return;
    if (haveEnd && !before(t, s.sTick[K], tEnd, qEnd)) return;
    // the first tick is admitted by the full (time, seq) comparison; ticks created from now on
    // have larger creation numbers than the bounding event, so they need t < tEnd strictly
    const double stop = haveEnd ? (tEnd < tLimit ? tEnd : tLimit) : tLimit;
    uint32_t c = 0;
    do {
        t = t + interval;
        ++c;
    } while (t < stop);
    s.ties += (haveEnd && t == tEnd) ? 1u : 0u;     // exact tie of independent events (diagnostic)
    s.tTick[K] = t;
    s.ticks[K] += c;
    const uint32_t n = (uint32_t)s.qn[K] + c * (uint32_t)mult;
    s.qn[K] = n > (uint32_t)kQueueCap ? kQueueCap : (int)n;
    s.seq += c;
    s.sTick[K] = s.seq - 1u;
}

// Both senders' silent ticks up to (tEnd, qEnd): silent_ticks<0> followed by silent_ticks<1>.  When the
// two senders tick in lockstep (same pending tick time, same interval -- the reference's senders both
// start at t = 0 with COUNTER_INTERVAL) the second pass repeats the first one addition for addition, so
// ONE loop serves both; the creation numbers are assigned as the two passes would (sender 0's batch
// first).  A pending tick exactly at tEnd takes the general path.
template <int D, int NS, int NJ, class ST>
GW_HD void silent_ticks_both(Sim<D, NS, NJ, ST> &s, const BandParams &B, double tEnd, uint32_t qEnd, bool haveEnd,
                             double tLimit)
{
    static_assert(NS == 2, "the tick logic is written for two senders per band");
    double t = s.tTick[0];
    const double interval = B.interval[0];
    if (t == s.tTick[1] && interval == B.interval[1] && !(haveEnd && t == tEnd)) {
        if (!(t < tLimit)) return;
        if (haveEnd && !(t < tEnd)) return;
        const double stop = haveEnd ? (tEnd < tLimit ? tEnd : tLimit) : tLimit;
        uint32_t c = 0;
        do {
            t = t + interval;
            ++c;
        } while (t < stop);
        s.ties += (haveEnd && t == tEnd) ? 2u : 0u;
        s.tTick[0] = t; s.tTick[1] = t;
        s.ticks[0] += c; s.ticks[1] += c;
        const uint32_t n0 = (uint32_t)s.qn[0] + c * (uint32_t)B.mult[0];
        const uint32_t n1 = (uint32_t)s.qn[1] + c * (uint32_t)B.mult[1];
        s.qn[0] = n0 > (uint32_t)kQueueCap ? kQueueCap : (int)n0;
        s.qn[1] = n1 > (uint32_t)kQueueCap ? kQueueCap : (int)n1;
        s.seq += c;
        s.sTick[0] = s.seq - 1u;
        s.seq += c;
        s.sTick[1] = s.seq - 1u;
        return;
    }
    silent_ticks<0>(s, B.mult[0], B.interval[0], tEnd, qEnd, haveEnd, tLimit);
    silent_ticks<1>(s, B.mult[1], B.interval[1], tEnd, qEnd, haveEnd, tLimit);
}

template <bool ALL_TICKS = false, int D, int NS, int NJ, class ST, class Ring>
GW_HD Event next_event(Sim<D, NS, NJ, ST> &s, const BandParams &B, double tLimit, const Ring &ring)
{
    static_assert(NS == 2, "the tick logic is written for two senders per band");
    Event ev = select_nontick(s, B, ring);
    // ticks that must go through the transition function: every tick of a plant env, otherwise
    // the ticks of a sender whose MAC waits for a packet
    // (finite traffic bursts: every tick goes through the transition function, which ends the process)
    // (`Ring::ext`: compile-time switch of the kernels for bands with receive mode / bursts -- everybody else
    // carries no code for them)
    const bool wake0 = ALL_TICKS || (Ring::ext && B.maxTicks[0] != 0) || s.mac[0] == MAC_WAIT_COND;
    const bool wake1 = ALL_TICKS || (Ring::ext && B.maxTicks[kMaxSend - 1] != 0) || s.mac[1] == MAC_WAIT_COND;
    if (wake0 && (ev.kind == EV_NONE || before(s.tTick[0], s.sTick[0], ev.t, ev.seq))) {
        ev.kind = EV_TICK; ev.idx = 0; ev.t = s.tTick[0]; ev.seq = s.sTick[0];
    }
    if (wake1 && (ev.kind == EV_NONE || before(s.tTick[1], s.sTick[1], ev.t, ev.seq))) {
        ev.kind = EV_TICK; ev.idx = 1; ev.t = s.tTick[1]; ev.seq = s.sTick[1];
    }
    const bool have = ev.kind != EV_NONE;
    if (!wake0 && !wake1) {
        silent_ticks_both(s, B, ev.t, ev.seq, have, tLimit);
    } else {
        if (!wake0) silent_ticks<0>(s, B.mult[0], B.interval[0], ev.t, ev.seq, have, tLimit);
        if (!wake1) silent_ticks<1>(s, B.mult[1], B.interval[1], ev.t, ev.seq, have, tLimit);
    }
    if (!have) {
        // nothing but silent ticks is pending: report the earliest one (its time is >= tLimit)
        const bool one = before(s.tTick[1], s.sTick[1], s.tTick[0], s.sTick[0]);
        ev.kind = EV_TICK; ev.idx = one ? 1 : 0;
        ev.t = one ? s.tTick[1] : s.tTick[0]; ev.seq = one ? s.sTick[1] : s.sTick[0];
    }
    return ev;
}

// ---------------------------------------------------------------------------
// which PHYs run SimplePhy._countBitErrors for this event (simple_stack.py:180-188)
//   once[p]  : bit p set -> one count
//   twice[p] : bit p set -> a second count (payload end: the completion callback counts,
//              then the receive process counts again -- appendix B #4)
// ---------------------------------------------------------------------------

template <int D, int NS, int NJ, class ST, class SRX>
GW_HD void count_set(const Sim<D, NS, NJ, ST> &s, const Event &ev, const SRX &srx, int &once, int &twice)
{
    once = 0; twice = 0;
    if (ev.kind != EV_PHY) return;
    const int d = ev.idx, ph = get_at(s.sphase, d);
    GW_UNROLL_D
    for (int p = 0; p < D; ++p) {
        const int rx = s.rxOf[p];
        if (rx < 0) continue;
        if (ph == S_HDR) {
            if (rx == d && s.rxSec[p] == 0) once |= 1 << p;
        } else {
            // S_SLOT: the new transmission adds power; S_PAY: the completing one removes it.
            // onReceivedPowerChange counts only for delta != 0 (simple_stack.py:224)
            const bool delta_nz = (p != d) && (srx_at<D>(srx, p, d) != 0.0);
            if (delta_nz) once |= 1 << p;
            if (ph == S_PAY && rx == d && s.rxSec[p] == 1) {
                if (delta_nz) twice |= 1 << p; else once |= 1 << p;
            }
        }
    }
}

// mode R: expected-value accounting, `duration` measured from the last RESET (appendix B #5)
template <int D, int NS, int NJ, class ST>
GW_HD void do_counts_R(Sim<D, NS, NJ, ST> &s, int once, int twice, double bitRate)
{
    GW_UNROLL_D
    for (int p = 0; p < D; ++p) {
        if (!((once >> p) & 1)) continue;
        const double duration = s.now - s.tReset[p];
        const double bitErrors = s.ber[p] * duration * bitRate;
        s.err[p] += bitErrors;
        if ((twice >> p) & 1) s.err[p] += bitErrors;
    }
}

// mode M: on-air bit range [k0, k1) of the segment that ends now, for PHY p
template <int D, int NS, int NJ, class ST>
GW_HD void mask_range(const Sim<D, NS, NJ, ST> &s, int p, double bitRate, int &sender, uint32_t &txseq,
                      int64_t &k0, int64_t &k1)
{
    const int e = get_at(s.rxOf, p);
    const double start = get_at(s.txStart, e);
    const uint32_t sq = get_at(s.txSeq, e) - 1u;
    const double t0 = get_at(s.segT0, p);
    sender = e; txseq = sq;
    k0 = (int64_t)floor((t0 - start) * bitRate);
    k1 = (int64_t)floor((s.now - start) * bitRate);
}

// Mode M with FED masks: the error flags of a row are data, so the count of a section (header or
// payload) does not depend on how the section is cut into constant-SINR segments -- the ranges of
// consecutive segments are contiguous ([k0, k1), [k1, k2), ... with k = floor((t - start) * bitRate) of
// the SAME boundary time t on both sides) and their counts are integers, added exactly in fp64.  Only
// the events that DECIDE on a section (header end, completion) therefore need a count: the bits from
// the start of the running segment (segT0: lock-on, header success, or the last position change) to
// now.  Power changes in between (other transmissions starting / ending) count nothing and leave
// segT0 alone.  Returns the receivers that decide at `ev`.
template <int D, int NS, int NJ, class ST>
GW_HD int fed_decide_set(const Sim<D, NS, NJ, ST> &s, const Event &ev)
{
    if (ev.kind != EV_PHY) return 0;
    const int d = ev.idx, ph = get_at(s.sphase, d);
    if (ph != S_HDR && ph != S_PAY) return 0;
    const int sec = ph == S_HDR ? 0 : 1;
    int need = 0;
    GW_UNROLL_D
    for (int p = 0; p < D; ++p)
        if (s.rxOf[p] == d && s.rxSec[p] == sec) need |= 1 << p;
    return need;
}

// ---------------------------------------------------------------------------
// plant hook: envs whose packets carry values of a simulated plant (the networked inverted
// pendulum, gymwipe/envs/inverted_pendulum.py) plug a plant object into the transition
// function.  NoPlant (CounterTrafficEnv) compiles to nothing.
//   tick_value(k, now)            value the sender's packets of this tick carry
//   delivered(d, p, value, now)   a data packet of sender d was decoded by device p
//   refresh_links(d, now, srx)    received powers from sender d at the start of its transmission
// ---------------------------------------------------------------------------

struct NoPlant {
    static constexpr bool active = false;
    GW_HD double tick_value(int, double) { return 0.0; }
    GW_HD void delivered(int, int, double, double) {}
    template <class SRX> GW_HD void refresh_links(int, double, const SRX &) {}
    GW_HD void put_value(int, uint64_t, double) {}
    GW_HD double get_value(int, uint64_t) { return 0.0; }
};

// ---------------------------------------------------------------------------
// helpers of the transition function
// ---------------------------------------------------------------------------

// fmod(t, kSlot) for t >= 0, bit-identical to the C library's (fmod is exact by definition):
// q = trunc(t * 1e6) estimates the slot number to within +-1 (relative error ~2e-16 of a number
// below 2^52); t - q L is a multiple of ulp(L) smaller than 2 L, hence exactly representable, so one
// explicit fma returns it without rounding and a single correction step finishes.  ~12
// instructions instead of the generic iterative fmod.  Valid while t < 4.5e9 s.
GW_HD double fmod_slot(double t)
{
    const double q = trunc(t * 1e6);
    double r = fma(-q, kSlot, t);
    if (r < 0.0) r += kSlot;
    else if (r >= kSlot) r -= kSlot;
    return r;
}

// derived constants of Params (host side, after bitRate / dataRate / maxBer / bitsFactor are set)
inline void finish_params(Params &P)
{
    P.berMult = 0.0;
    int e = 0;
    const double m = frexp(P.maxBer, &e);
    const double bitsPerByte = 8 * P.bitsFactor;
    if (m == 0.5 && P.maxBer <= 1.0 && bitsPerByte == floor(bitsPerByte)) P.berMult = 1.0 / P.maxBer;
    for (int k = 0; k < 32; ++k) P.airtime[k] = (k * 8) / P.dataRate;
    // x / dataRate through the reciprocal: q0 = x * y, r = fma(-q0, R, x), q = fma(r, y, q0) is the
    // correctly rounded quotient for almost all operands (Markstein); it is USED only if it equals the
    // division on the whole domain of the path, which is finite -- bit counts 8 k, k <= kMaxAirBytes --
    // and checked here exhaustively (~70 k divisions, once per gw_create)
    P.rateInv = 1.0 / P.dataRate;
    for (int k = 0; k <= kMaxAirBytes; ++k) {
        const double x = k * 8;
        const double q0 = x * P.rateInv;
        const double q = fma(fma(-q0, P.dataRate, x), P.rateInv, q0);
        if (q != x / P.dataRate) { P.rateInv = 0.0; break; }
    }
}

// the fp64 divisions that the verified shortcuts replace: out of line, so that their ~60 instructions each do not
// sit between the hot instructions of every call site (the kernels are instruction-fetch bound)
GW_HD_COLD double div_cold(double x, double y) { return x / y; }

// duration of `bytes` bytes on air at the data rate (physical.py:244-250): (bytes * 8) / dataRate
GW_HD double airtime_of(const Params &P, int bytes)
{
    if ((unsigned)bytes < 32u) return P.airtime[bytes];
    const double x = bytes * 8;
    if (P.rateInv != 0.0 && bytes <= kMaxAirBytes) {
        const double q0 = x * P.rateInv;
        return fma(fma(-q0, P.dataRate, x), P.rateInv, q0);     // == x / dataRate, verified in finish_params
    }
    return div_cold(x, P.dataRate);
}

// SimplePhy._decide: round(bitErrorSum) / totalBits <= maxCorrectableBer (simple_stack.py:274-277).
// With maxBer = 2^-k and integer-valued operands (x = rint(errSum), totalBits = bytes * 8 * 1.25 < 2^52)
// the correctly rounded quotient is <= 2^-k exactly when x * 2^k <= totalBits: fl(x / y) <= m  <=>
// x / y <= m (1 + 2^-53)  <=>  x 2^k - y <= y 2^-53 < 1, and the left side is an integer.  Saves an
// fp64 division (~30 dependent instructions) per decision; other code rates take the division.
GW_HD bool within_max_ber(const Params &P, double errSum, double totalBits)
{
    const double x = rint(errSum);
    if (P.berMult != 0.0) return x * P.berMult <= totalBits;
    return div_cold(x, totalBits) <= P.maxBer;
}

template <int D, int NS, int NJ, class ST>
GW_HD void begin_slot_wait(Sim<D, NS, NJ, ST> &s, int d)
{
    // self._transmitting = True; yield SimMan.nextTimeSlot(TIME_SLOT_LENGTH)  (simple_stack.py:202-204)
    const double t = s.now + (kSlot - fmod_slot(s.now));        // simtools.py:53
    const uint32_t q = s.seq++;
    set_at(s.sphase, d, (int)S_SLOT);
    set_at(s.tEv, d, t);
    set_at(s.sEv, d, q);
}

template <int D, int NS, int NJ, class ST>
GW_HD void phy_send_init(Sim<D, NS, NJ, ST> &s, int d)
{
    // SimplePhy.macInHandler start: wait while the receiver is active (simple_stack.py:199-200)
    const bool receiving = get_at(s.rxOf, d) >= 0;
    if (receiving) set_at(s.sphase, d, (int)S_WAITRX);
    else begin_slot_wait(s, d);
}

// size of the head packet of sender k's queue (bytes of the Transmittable):
// packets are enqueued `mult` per tick with byteSize = counter at that tick
template <int D, int NS, int NJ, class ST, class Ring>
GW_HD int head_size(const Sim<D, NS, NJ, ST> &s, const BandParams &B, int k, const Ring &ring)
{
    const int rule = k == 0 ? B.payloadRule[0] : B.payloadRule[kMaxSend - 1];
    if (rule >= 0) return rule;
    const uint32_t m = (uint32_t)(k == 0 ? B.mult[0] : B.mult[kMaxSend - 1]);
    const uint64_t ticks = get_at(s.ticks, k);
    const uint32_t qn = (uint32_t)get_at(s.qn, k);
    const uint64_t enq = ticks * m;
    const uint64_t j = enq - (uint64_t)qn;
    if (j < ring.epochK(s, k) * m) return ring(k, (uint32_t)j & (uint32_t)(kRingSlots - 1));     // predates the last reset()
    // tick of packet j = ticks - ceil(qn / m)
    const uint64_t back = (qn + m - 1u) / m;
    const uint64_t tick = ticks - back;
    const uint64_t c = (uint64_t)ring.epochC(s, k) + (tick - ring.epochK(s, k));
    const int size = c > (uint64_t)kCounterBound ? kCounterBound : (int)c;
    return size;
}

// one pass of the SimpleMac window loop body with a non-empty queue (simple_stack.py:417-434)
template <int D, int NS, int NJ, class ST, class Ring, class Plant>
GW_HD void mac_try_send(Sim<D, NS, NJ, ST> &s, const Params &P, const BandParams &B, int k, const Ring &ring, Plant &plant)
{
    const int size = head_size(s, B, k, ring);
    const double stopW = get_at(s.stopW, k);
    const double timeLeft = stopW - s.now;
    const double txTime = airtime_of(P, kMacHdr + kNetHdr + size);      // bitSize / dataRate
    if (!(timeLeft > txTime)) {
        set_at(s.mac, k, (int)MAC_IDLE);        // yield timeoutEvent
        return;
    }
    if (Plant::active) {
        // tick that enqueued the head packet = ticks - ceil(qn / mult)
        const uint32_t m = (uint32_t)(k == 0 ? B.mult[0] : B.mult[kMaxSend - 1]);
        const uint64_t tick = get_at(s.ticks, k) - (((uint32_t)get_at(s.qn, k) + m - 1u) / m);
        set_at(s.txVal, k, plant.get_value(k, tick));
    }
    set_at(s.qn, k, get_at(s.qn, k) - 1);
    set_at(s.mac, k, (int)MAC_WAIT_TX);
    set_at(s.cmdPay, k, kNetHdr + size);
    phy_send_init(s, k);
}

// loop head of the window loop (simple_stack.py:408-416)
template <int D, int NS, int NJ, class ST, class Ring, class Plant>
GW_HD void mac_loop_head(Sim<D, NS, NJ, ST> &s, const Params &P, const BandParams &B, int k, const Ring &ring, Plant &plant)
{
    const bool done = get_at(s.wDone, k) != 0, empty = get_at(s.qn, k) == 0;
    if (done) { set_at(s.mac, k, (int)MAC_NONE); return; }
    if (empty) { set_at(s.mac, k, (int)MAC_WAIT_COND); return; }
    mac_try_send(s, P, B, k, ring, plant);
}

// end of SimplePhy._receive (simple_stack.py:264-267) without the deferred wake-up
template <int D, int NS, int NJ, class ST>
GW_HD void rx_clear(Sim<D, NS, NJ, ST> &s, int p)
{
    set_at(s.rxOf, p, -1);
    set_at(s.err, p, 0.0);
    set_at(s.ber, p, 0.0);
    set_at(s.tReset, p, s.now);
    set_at(s.segT0, p, s.now);
}

template <int D, int NS, int NJ, class ST>
GW_HD bool decide(const Sim<D, NS, NJ, ST> &s, const Params &P, int p, double totalBits)
{
    // bitErrorSum = round(bitErrorSum); bitErrorSum / totalBits <= maxCorrectableBer  (simple_stack.py:274-277)
    return within_max_ber(P, get_at(s.err, p), totalBits);
}

template <int D, int NS, int NJ, class ST>
GW_HD bool decide_rec(Sim<D, NS, NJ, ST> &s, const Params &P, int p, int section, double totalBits)
{
    const bool ok = decide(s, P, p, totalBits);
    trace_rec(s, REC_DEC, s.now, p, section, get_at(s.err, p), totalBits, ok ? 1.0 : 0.0);
    return ok;
}

// ---------------------------------------------------------------------------
// transition function: applies one timed event (counts already done); returns the set of
// PHYs whose bit error rate must be re-evaluated afterwards (SimplePhy._updateBitErrorRate)
// ---------------------------------------------------------------------------

template <int D, int NS, int NJ, class ST, class SRX, class Ring, class Plant>
GW_HD int apply_event(Sim<D, NS, NJ, ST> &s, const Params &P, const BandParams &B, const Event &ev,
                      SRX srx, const Ring &ring, Plant &plant)
{
    constexpr int RRM = NS;
    int berMask = 0;
    s.now = ev.t;
    switch (ev.kind) {
    case EV_TICK: {
        // SenderDevice.senderProcess (counter_traffic.py:53-61): `mult` packets, counter += 1,
        // next tick; then the MAC wakes if it waits on (_packetAddedEvent | timeoutEvent)
        const int k = ev.idx;
        const int mult = k == 0 ? B.mult[0] : B.mult[kMaxSend - 1];
        const double interval = k == 0 ? B.interval[0] : B.interval[kMaxSend - 1];
        const int maxTicks = k == 0 ? B.maxTicks[0] : B.maxTicks[kMaxSend - 1];
        if (Ring::ext && maxTicks != 0 && get_at(s.ticks, k) >= (uint64_t)maxTicks) {
            // the burst is over: this wake-up only ends the traffic process (its process event takes a number)
            set_at(s.tTick, k, (double)INFINITY);
            s.seq++;
            break;
        }
        const int n = get_at(s.qn, k) + mult;
        set_at(s.qn, k, n > kQueueCap ? kQueueCap : n);             // drop-oldest
        set_at(s.ticks, k, get_at(s.ticks, k) + 1);
        set_at(s.tTick, k, s.now + interval);
        set_at(s.sTick, k, s.seq++);
        if (Plant::active) plant.put_value(k, get_at(s.ticks, k) - 1, plant.tick_value(k, s.now));
        const bool wake = get_at(s.mac, k) == MAC_WAIT_COND;
        if (wake) mac_try_send(s, P, B, k, ring, plant);
        break;
    }
    case EV_JAM: {
        const int d = RRM + 1 + ev.idx;            // kMaxJam == 1: jammer slot 0
        if (NJ > 0) {
            if (s.jamStage[0] == 0) {                   // yield timeout(initialDelay)
                s.jamStage[0] = 1; s.tJam[0] = s.now + B.jamDelay[0]; s.sJam[0] = s.seq++;
            } else if (s.jamStage[0] == 1) {            // first yield timeout(sendInterval)
                s.jamStage[0] = 2; s.tJam[0] = s.now + B.jamInterval[0]; s.sJam[0] = s.seq++;
            } else {
                // macIn.send(SEND) -> queued executor; then yield timeout(sendInterval)
                s.tJam[0] = s.now + B.jamInterval[0]; s.sJam[0] = s.seq++;
                const bool busy = get_at(s.sphase, d) != S_IDLE;
                if (busy) { s.jamPending[0] += 1; if (s.jamPending[0] > 60) s.fault = FAULT_SENDQ; }
                else { set_at(s.cmdPay, d, B.jamPay[0]); phy_send_init(s, d); }
            }
        }
        break;
    }
    case EV_PHY: {
        const int d = ev.idx;
        const int ph = get_at(s.sphase, d);
        if (ph == S_SLOT) {
            // FrequencyBand.transmit -> Transmission.__init__ (physical.py:224-279,596-608)
            const int payBytes = get_at(s.cmdPay, d);
            const int hdrBytes = (NJ > 0 && d > RRM) ? B.jamHdr[0] : kMacHdr;
            const double hd = airtime_of(P, hdrBytes);
            const double pd = airtime_of(P, payBytes);
            const double duration = hd + pd;
            const double stop = s.now + duration;
            const double headerStop = s.now + hd;
            const double tH = s.now + (headerStop > s.now ? headerStop - s.now : 0.0);   // timeoutUntil
            const double tC = s.now + (stop > s.now ? stop - s.now : 0.0);
            const uint32_t qH = s.seq++, qC = s.seq++;
            set_at(s.sphase, d, (int)S_HDR);
            set_at(s.tEv, d, tH);
            set_at(s.sEv, d, qH);
            set_at(s.tC, d, tC);
            set_at(s.sC, d, qC);
            set_at(s.txStart, d, s.now);
            set_at(s.tStop, d, stop);
            set_at(s.txSeq, d, get_at(s.txSeq, d) + 1u);
            s.nTx += 1;
            trace_rec(s, REC_TX, s.now, d, stop, (hdrBytes * 8) * P.bitsFactor, (payBytes * 8) * P.bitsFactor, 0.0);
            if constexpr (Plant::active) plant.refresh_links(d, s.now, srx);
            // zero-delay notification: every other PHY registers the received power
            // (simple_stack.py:130-144); a PHY that is receiving re-evaluates its BER
            GW_UNROLL_D
            for (int p = 0; p < D; ++p) {
                if (p == d) continue;
                const double rp = srx_at<D>(srx, p, d);
                s.P[p] += rp;
                if (rp != 0.0) s.pchg |= 1 << p;
                if (s.rxOf[p] >= 0 && rp != 0.0) {
                    const bool completed = s.now >= get_at(s.tStop, s.rxOf[p]);
                    if (!completed) berMask |= 1 << p;
                }
            }
            // receive processes in PHY construction order: idle, non-transmitting PHYs lock on
            // (simple_stack.py:214-235)
            GW_UNROLL_D
            for (int p = 0; p < D; ++p) {
                if (p == d || s.rxOf[p] >= 0 || s.sphase[p] >= S_SLOT) continue;
                s.rxOf[p] = d; s.rxSec[p] = 0;
                s.err[p] = 0; s.ber[p] = 0.0; s.tReset[p] = s.now; set_at(s.segT0, p, s.now);
                berMask |= 1 << p;
            }
        } else if (ph == S_HDR) {
            // eHeaderCompletes: receivers decide on the header (simple_stack.py:241-251)
            const int hdrBytes = (NJ > 0 && d > RRM) ? B.jamHdr[0] : kMacHdr;
            const double hdrBits = (hdrBytes * 8) * P.bitsFactor;
            int wake = 0;
            GW_UNROLL_D
            for (int p = 0; p < D; ++p) {
                if (s.rxOf[p] != d || s.rxSec[p] != 0) continue;
                if (decide_rec(s, P, p, 0, hdrBits)) {
                    // _resetBitErrorCounter, then _updateBitErrorRate (simple_stack.py:248-250): with the
                    // power sum unchanged since the last evaluation the same (S, N) gives the same value
                    const double same = s.ber[p];
                    s.rxSec[p] = 1; s.err[p] = 0; s.ber[p] = 0.0; s.tReset[p] = s.now; set_at(s.segT0, p, s.now);
                    if ((s.pchg >> p) & 1) {
                        berMask |= 1 << p;
                    } else {
                        s.ber[p] = same;
                        trace_rec(s, REC_BER, s.now, p, same, 0.0, 0.0, 0.0);
                    }
                } else {
                    rx_clear(s, p);
                    if (s.sphase[p] == S_WAITRX) wake |= 1 << p;
                }
            }
            set_at(s.sphase, d, (int)S_PAY);
            set_at(s.tEv, d, get_at(s.tC, d));
            set_at(s.sEv, d, get_at(s.sC, d));
            GW_UNROLL_D
            for (int p = 0; p < D; ++p) if ((wake >> p) & 1) begin_slot_wait(s, p);   // _nReceivingFinished.event
        } else {
            // eCompletes, callbacks in registration order:
            // 1. the sender's macInHandler resumes: _transmitting = False (simple_stack.py:210)
            const int payBytes = get_at(s.cmdPay, d);
            const double stopD = get_at(s.tStop, d);
            set_at(s.sphase, d, (int)S_IDLE);
            const double payBits = (payBytes * 8) * P.bitsFactor;
            // 2. _onCompletingTransmission of every other PHY (simple_stack.py:146-157)
            GW_UNROLL_D
            for (int p = 0; p < D; ++p) {
                if (p == d) continue;
                const double rp = srx_at<D>(srx, p, d);
                s.P[p] += -rp;
                if (rp != 0.0) s.pchg |= 1 << p;
                if (s.rxOf[p] >= 0 && rp != 0.0) {
                    if (s.rxOf[p] == d) {
                        // `if not t.completed: _updateBitErrorRate(t)` with the power entry
                        // already popped: the reference raises KeyError (appendix B #12)
                        if (!(s.now >= stopD)) s.fault = FAULT_REF_KEYERROR;
                    } else {
                        const bool completed = s.now >= get_at(s.tStop, s.rxOf[p]);
                        if (!completed) berMask |= 1 << p;
                    }
                }
            }
            // 3. receivers that passed the header decide on the payload and deliver
            int window = -1, wake = 0, received = 0;
            GW_UNROLL_D
            for (int p = 0; p < D; ++p) {
                if (s.rxOf[p] != d || s.rxSec[p] != 1) continue;
                if (decide_rec(s, P, p, 1, payBits)) {
                    if (Plant::active && d != RRM) plant.delivered(d, p, get_at(s.txVal, d), s.now);
                    if (p < NS) {
                        // SimpleMac.phyInHandler (blocking, not queued): an announcement addressed to an idle
                        // MAC opens its window (simple_stack.py:386-434); a data packet addressed to an idle MAC
                        // in receive mode goes to the network layer (:436-444)
                        const bool idle = get_at(s.mac, p) == MAC_NONE;
                        if (d == RRM && s.annDest == p) {
                            if (idle) window = p;
                        } else if (Ring::ext && d < NS && idle && (p == 0 ? B.recv[0] : B.recv[kMaxSend - 1])) {
                            received |= 1 << p;         // the two senders of a band address each other
                        }
                    } else if (p == RRM) {
                        // SimpleRrmMac.phyInHandler -> interpreter.onPacketReceived
                        // (devices.py:163-168, counter_traffic.py:75-80); payload.value == 2 always
                        if (d < NS) {
                            if (d == 0) s.rv0 = kCounterByteLen;
                            if (d == 1) s.rv1 = kCounterByteLen;
                            s.latestDiff = s.rv0 - s.rv1;
                            set_at(s.nDeliv, d, get_at(s.nDeliv, d) + 1u);
                            trace_rec(s, REC_RX, s.now, d, 0.0, 0.0, 0.0, 0.0);
                        } else {
                            // a PHY-only sender's packet: the interpreter sees it, nothing changes
                            trace_rec(s, REC_RX, s.now, d, 0.0, 0.0, 0.0, 0.0);
                        }
                    }
                }
                rx_clear(s, p);
                if (s.sphase[p] == S_WAITRX) wake |= 1 << p;
            }
            // zero-delay children in SimPy's pop order:
            // a. URGENT: phyInHandler of the grantee opens its window (simple_stack.py:399-406)
            if (window >= 0) {
                const double timeTotal = s.annSlots * kSlot;
                const double stopW = s.now + timeTotal;
                const uint32_t qW = s.seq++;
                set_at(s.stopW, window, stopW);
                set_at(s.sW, window, qW);
                set_at(s.wPend, window, 1);
                set_at(s.wDone, window, 0);
                mac_loop_head(s, P, B, window, ring, plant);
            }
            // b. SEND eProcessed: the sender's upper layer resumes
            if (d < NS) {
                mac_loop_head(s, P, B, d, ring, plant);         // `yield message.eProcessed` returns
            } else if (d == RRM) {
                s.tRrm = s.now + (s.annSlots + 1) * kSlot;      // simple_stack.py:558
                s.sRrm = s.seq++;
                s.rrmPend = 1;
            } else {
                // c. executeNext of the queued macIn executor: a jammer's pending SEND starts
                if (NJ > 0 && s.jamPending[0] > 0) {
                    s.jamPending[0] -= 1;
                    phy_send_init(s, d);
                }
            }
            // d. _nReceivingFinished.event of the receivers that finished
            GW_UNROLL_D
            for (int p = 0; p < D; ++p) if ((wake >> p) & 1) begin_slot_wait(s, p);
            // e. RECEIVE.eProcessed: the device's receive loop hands the packet to onReceive and issues the
            // next RECEIVE command with a fresh timeout (devices.py:88-95, simple_stack.py:452-460)
            GW_UNROLL
            for (int k = 0; k < NS; ++k) {
                if (!Ring::ext || !((received >> k) & 1)) continue;
                ring.add_received(k);
                trace_rec(s, REC_MRX, s.now, k, 0.0, 0.0, 0.0, 0.0);
                ring.set_rx(k, s.now + kReceiveTimeout, s.seq++);
            }
        }
        break;
    }
    case EV_RXTO: {
        // the current RECEIVE command timed out (or the receive loop starts): setProcessed() without a result,
        // the loop issues the next command (simple_stack.py:473-478, devices.py:88-93)
        ring.set_rx(ev.idx, s.now + kReceiveTimeout, s.seq++);
        break;
    }
    case EV_W: {
        // window timeoutEvent processed (simple_stack.py:406-420)
        const int k = ev.idx;
        set_at(s.wPend, k, 0);
        if (get_at(s.mac, k) == MAC_WAIT_TX) set_at(s.wDone, k, 1);
        else set_at(s.mac, k, (int)MAC_NONE);
        break;
    }
    case EV_RRM:
        // assignMessage.setProcessed() (simple_stack.py:561): the step ends here
        s.rrmPend = 0;
        s.assignDone = 1;
        break;
    default:
        s.fault = FAULT_EMPTY;
    }
    return berMask;
}

template <int D, int NS, int NJ, class ST, class SRX, class Ring>
GW_HD int apply_event(Sim<D, NS, NJ, ST> &s, const Params &P, const BandParams &B, const Event &ev,
                      const SRX &srx, const Ring &ring)
{
    NoPlant np;
    return apply_event(s, P, B, ev, srx, ring, np);
}

// BER(S, N) is a pure function and, with static geometry, the same few (S, N) pairs recur in
// every env and step: a memo (exact fp64 results keyed by the exact bit patterns of S and N)
// replaces ~10^3 dependent fp64 instructions by one table probe.  NoMemo = always evaluate.
struct NoMemo {
    GW_HD bool get(double, double, double &, int = 0) const { return false; }
    GW_HD void put(double, double, double, int = 0) const {}
};

// slot of a (receiver p, sender e) link in a per-band-sim BER cache: two ways per link -- with / without the band's
// PHY-only interferer on the air, the two noise levels a link usually sees
template <int D, int NS, int NJ, class ST>
GW_HD int memo_slot(const Sim<D, NS, NJ, ST> &s, int p, int e)
{
    int way = 0;
    if (NJ > 0) way = s.sphase[D - 1] >= S_HDR ? 1 : 0;
    return ((p * kMaxDev + e) << 1) | way;
}

// SimplePhy._updateBitErrorRate for the PHYs in berMask (simple_stack.py:161-173)
template <int D, int NS, int NJ, class ST, class SRX, class Memo>
GW_HD void update_bers(Sim<D, NS, NJ, ST> &s, const Params &P, int berMask, const SRX &srx, const Memo &memo)
{
    GW_UNROLL_D
    for (int p = 0; p < D; ++p) {
        if (!((berMask >> p) & 1)) continue;
        const int e = s.rxOf[p];
        if (e < 0) continue;
        const double S = srx_at<D>(srx, p, e);
        const double N = s.P[p] - S;
        if (!(S >= 0) || !(N >= 0)) { s.fault = FAULT_REF_ASSERT; continue; }   // simple_stack.py:168-169
        double ber;
        const int slot = memo_slot(s, p, e);
        if (!memo.get(S, N, ber, slot)) {
            ber = ber_bpsk_mw_cold(S, N, P.tenLog10BitRate, P.qDen);
            memo.put(S, N, ber, slot);
        }
        s.ber[p] = ber;
        s.pchg &= ~(1 << p);
        trace_rec(s, REC_BER, s.now, p, ber, 0.0, 0.0, 0.0);
    }
}

template <int D, int NS, int NJ, class ST, class SRX>
GW_HD void update_bers(Sim<D, NS, NJ, ST> &s, const Params &P, int berMask, const SRX &srx)
{
    update_bers(s, P, berMask, srx, NoMemo());
}

// SimpleRrmDevice.assignFrequencyBand + SimpleRrmMac._sendAnnouncement start
// (devices.py:178-203, simple_stack.py:536-556): the RRM PHY receives a SEND command
template <int D, int NS, int NJ, class ST>
GW_HD void begin_assignment(Sim<D, NS, NJ, ST> &s, const Params &P, int device, int duration)
{
    const long long slots = (long long)duration * P.factor;         // counter_traffic.py:149
    int nbytes = 1;                                                  // len(str(slots)), messages.py:62-64
    if (slots < 1000000000LL) {
        const int v = (int)slots;
        nbytes += (v >= 10) + (v >= 100) + (v >= 1000) + (v >= 10000) + (v >= 100000) + (v >= 1000000)
                + (v >= 10000000) + (v >= 100000000);
    } else {
        for (long long lim = 10; lim <= slots && nbytes < 18; lim *= 10) ++nbytes;
    }
    s.annDest = device;
    s.annSlots = (double)slots;
    s.annBytes = nbytes;
    s.assignDone = 0;
    // the RRM PHY's queued macIn executor is idle here: its previous SEND completed before
    // the previous assignment's guard time-out (simple_stack.py:557-558)
    if (s.sphase[NS] != S_IDLE) s.fault = FAULT_SENDQ;
    s.cmdPay[NS] = nbytes;
    phy_send_init(s, NS);
}

// ---------------------------------------------------------------------------
// serial drivers (one band-sim at a time): used by the host build and by mode R on the
// device; the mode-M kernels interleave the same pieces with warp-cooperative counting
// ---------------------------------------------------------------------------

struct NoMasks {
    GW_HD int64_t operator()(int, int, uint32_t, int64_t, int64_t, double) const { return 0; }
};

// processes ONE timed event; `masks(receiver, sender, txseq, k0, k1, ber)` supplies mode-M counts
template <int MODE, int D, int NS, int NJ, class ST, class SRX, class Ring, class Masks, class Memo, class Plant>
GW_HD void process_event(Sim<D, NS, NJ, ST> &s, const Params &P, const BandParams &B, const Event &ev,
                         SRX srx, const Ring &ring, const Masks &masks, const Memo &memo, Plant &plant)
{
    int once, twice;
    GW_STAT_MACRO(2 + (ev.kind == EV_TICK ? 1 : 0));
    s.now = ev.t;
    if (MODE == MODE_M_FED) { once = fed_decide_set(s, ev); twice = 0; }
    else count_set(s, ev, srx, once, twice);
    if (MODE == MODE_R) {
        do_counts_R(s, once, twice, P.bitRate);
    } else {
        GW_UNROLL_D
        for (int p = 0; p < D; ++p) {
            if (!((once >> p) & 1)) continue;
            int sender; uint32_t txseq; int64_t k0, k1;
            mask_range(s, p, P.bitRate, sender, txseq, k0, k1);
            if (k1 > k0) s.err[p] += (double)masks(p, sender, txseq, k0, k1, s.ber[p]);
            set_at(s.segT0, p, s.now);
        }
    }
    const int berMask = apply_event(s, P, B, ev, srx, ring, plant);
    update_bers(s, P, berMask, srx, memo);
}

template <int MODE, int D, int NS, int NJ, class ST, class SRX, class Ring, class Masks, class Memo>
GW_HD void process_event(Sim<D, NS, NJ, ST> &s, const Params &P, const BandParams &B, const Event &ev,
                         const SRX &srx, const Ring &ring, const Masks &masks, const Memo &memo)
{
    NoPlant np;
    process_event<MODE>(s, P, B, ev, srx, ring, masks, memo, np);
}

// ---------------------------------------------------------------------------
// macro events (mode R): exact shortcuts through event sequences whose order is known in advance
//
// isolated_tx(): the slot-start event of device d is the next event, every other PHY is idle and
// not receiving, no MAC waits for a packet (ticks stay silent) and every other timed event lies
// strictly after the completion of d's transmission.  Then nothing can interleave with the three
// events of the transmission (slot start, header end, completion), every other PHY locks on at the
// start, and the three handlers of apply_event() + their counts + BER updates collapse into
// straight-line code on registers.  The arithmetic, its order, the creation numbers of the non-tick
// events and the trace records are those of the generic path, operation for operation; the silent
// ticks of the whole transmission are applied in ONE batch bounded by the completion event (nothing
// observes a queue before the completion handler), so a pending tick's creation number may differ
// from the generic path's -- it only breaks exact-time ties between a tick and an unrelated event,
// which are counted in `ties` either way.  Anything unusual (a power table entry that fails the
// reference's assertions, a completion time that rounds below the stop time -- appendix B #12)
// declines, and the generic path handles it.  Transmissions that follow back to back -- the
// grantee's first packet after the announcement, the next packet of the same window -- are CHAINED:
// the follower's slot-start event is the next event by the same test, so no event selection is
// needed in between, and the chain keeps the received powers, the links from the sender and the
// last BER per receiver in registers (the same (S, N) recurs from packet to packet).
//
// quiet_tail(): after the announcement (and the data packets) nothing is on air, no MAC waits for a
// packet or a transmission and only window time-outs and the RRM guard time-out are pending: they
// cannot create events (the time-outs only clear flags), so they are applied without the selection
// machinery and the silent ticks up to the guard time-out in one batch.
// ---------------------------------------------------------------------------

// Macro events are exact wherever they apply, but a warp whose lanes split between a macro event and
// the generic transition function executes both.  With PHY-only interferers on the band a large share
// of the transmissions is not isolated, and the split costs more than the macro events save
// (configs[3]: 2.9 ms per step without, 5.8 ms with); they are therefore used on bands without
// interferers; the host build can force them (P.noMacro = -1: the CPU tests exercise them with
// interferers too).
template <int MODE, int NJ>
GW_HD bool macros_enabled(const Params &P)
{
#ifdef __CUDA_ARCH__
    return MODE == MODE_R && NJ == 0 && P.noMacro == 0;        // kernels with interferers carry no macro code
#else
    return MODE == MODE_R && (P.noMacro < 0 || (P.noMacro == 0 && NJ == 0));
#endif
}

template <int D, int NS, int NJ, class ST, class SRX, class Ring, class Memo>
GW_HD bool isolated_tx(Sim<D, NS, NJ, ST> &s, const Params &P, const BandParams &B, const Event &ev,
                       const SRX &srx, const Ring &ring, const Memo &memo, double tLimit, bool idleStart = false)
{
    static_assert(NS == 2, "the tick logic is written for two senders per band");
    constexpr int RRM = NS;
    int d = ev.idx;
    double tOther = tLimit;
    // idleStart: the caller established that `ev` is the slot start of the RRM's announcement in a
    // band-sim where nothing else is active or pending (run_until_assign): the tests below hold
    if (!idleStart) {
        if (get_at(s.sphase, d) != S_SLOT) return false;
        // structural part of the isolation test; it stays true along the chain (nothing but the chained
        // transmissions happens, and their completion handlers leave every receiver idle)
        bool ok = true;
        GW_UNROLL
        for (int p = 0; p < D; ++p) ok &= (s.rxOf[p] < 0) & ((p == d) | (s.sphase[p] == S_IDLE));
        GW_UNROLL
        for (int k = 0; k < NS; ++k) ok &= s.mac[k] != MAC_WAIT_COND;
        if (!ok) return false;
        // the earliest OTHER timed event (window time-outs, the RRM guard time-out, jammer wake-ups,
        // the caller's horizon): a transmission is isolated if it completes strictly before it
        GW_UNROLL
        for (int k = 0; k < NS; ++k) if (s.wPend[k]) tOther = fmin(tOther, s.stopW[k]);
        if (s.rrmPend) tOther = fmin(tOther, s.tRrm);
        GW_UNROLL
        for (int j = 0; j < NJ; ++j) tOther = fmin(tOther, s.tJam[j]);
    }

    // state of the chain, in registers: received powers (written back at the end), the links from
    // the current sender, the last BER evaluated per receiver (the same (S, N) recurs from packet
    // to packet), transmissions not yet added to txSeq[d]
    double ts = ev.t;
    double Pc[D], rp[D], lastN[D], lastBer[D];
    GW_UNROLL
    for (int p = 0; p < D; ++p) { Pc[p] = s.P[p]; rp[p] = srx_at<D>(srx, p, d); lastN[p] = -1.0; lastBer[p] = 0.0; }
    uint32_t txs = 0;
    bool any = false;
    NoPlant plant;
    for (;;) {
        // Transmission.__init__ (physical.py:224-279), as in apply_event / S_SLOT
        const int payBytes = get_at(s.cmdPay, d);
        const int hdrBytes = (NJ > 0 && d > RRM) ? B.jamHdr[0] : kMacHdr;
        const double hd = airtime_of(P, hdrBytes);
        const double pd = airtime_of(P, payBytes);
        const double duration = hd + pd;
        const double stop = ts + duration;
        const double headerStop = ts + hd;
        const double tH = ts + (headerStop > ts ? headerStop - ts : 0.0);
        const double tC = ts + (stop > ts ? stop - ts : 0.0);
        bool okT = (tH > ts) & (tC > tH) & (tC >= stop) & (tC < tOther);
        // received powers; the reference's assertions on signal / noise power (simple_stack.py:168-169)
        double Pn[D];
        GW_UNROLL
        for (int p = 0; p < D; ++p) {
            Pn[p] = Pc[p] + rp[p];
            const double N = Pn[p] - rp[p];
            okT &= (p == d) | ((rp[p] >= 0) & (N >= 0));
        }
        if (!okT) break;
        any = true;
        GW_STAT_MACRO(0);

        // ---- slot start: the transmission begins, every other PHY registers its power and locks on
        s.seq++;                                    // creation number of the header-end event
        const uint32_t qC = s.seq++;
        ++txs;
        s.nTx += 1;
        const double hdrBits = (hdrBytes * 8) * P.bitsFactor;
        const double payBits = (payBytes * 8) * P.bitsFactor;
        trace_rec(s, REC_TX, ts, d, stop, hdrBits, payBits, 0.0);
        GW_UNROLL
        for (int p = 0; p < D; ++p) {
            if (p == d) continue;
            const double S = rp[p], N = Pn[p] - S;
            if (N != lastN[p]) {
                double b;
                if (!memo.get(S, N, b)) {
                    b = ber_bpsk_mw_cold(S, N, P.tenLog10BitRate, P.qDen);
                    memo.put(S, N, b);
                }
                lastN[p] = N; lastBer[p] = b;
            }
            trace_rec(s, REC_BER, ts, p, lastBer[p], 0.0, 0.0, 0.0);
        }

        // ---- header end (simple_stack.py:241-251).  Ticks are silent here: those up to the
        // completion are applied in one batch below (nothing observes the queues before the
        // completion handler)
        bool locked[D];
        GW_UNROLL
        for (int p = 0; p < D; ++p) {
            locked[p] = false;
            if (p == d) continue;
            const double bitErrors = lastBer[p] * (tH - ts) * P.bitRate;
            const double e = 0.0 + bitErrors;
            locked[p] = within_max_ber(P, e, hdrBits);
            trace_rec(s, REC_DEC, tH, p, 0, e, hdrBits, locked[p] ? 1.0 : 0.0);
        }
        if (s.trace != nullptr) {
            // the BER of a receiver that passed the header is evaluated again: same powers, same value
            GW_UNROLL
            for (int p = 0; p < D; ++p) if (p != d && locked[p]) trace_rec(s, REC_BER, tH, p, lastBer[p], 0.0, 0.0, 0.0);
        }

        // ---- completion (simple_stack.py:146-157, 253-267)
        silent_ticks_both(s, B, tC, qC, true, tLimit);
        s.now = tC;
        set_at(s.sphase, d, (int)S_IDLE);
        int window = -1;
        GW_UNROLL
        for (int p = 0; p < D; ++p) {
            if (p == d) continue;
            double e = 0.0;
            if (locked[p]) {
                // the completion callback counts (if the power changes), then the receive process
                // counts again -- appendix B #4
                const double bitErrors = lastBer[p] * (tC - tH) * P.bitRate;
                e = 0.0 + bitErrors;
                if (rp[p] != 0.0) e += bitErrors;
            }
            Pc[p] = Pn[p] + -rp[p];
            if (locked[p]) {
                const bool okP = within_max_ber(P, e, payBits);
                trace_rec(s, REC_DEC, tC, p, 1, e, payBits, okP ? 1.0 : 0.0);
                if (okP) {
                    if (p < NS) {
                        if (d == RRM && s.annDest == p && s.mac[p] == MAC_NONE) window = p;
                    } else if (p == RRM) {
                        if (d < NS) {
                            if (d == 0) s.rv0 = kCounterByteLen;
                            if (d == 1) s.rv1 = kCounterByteLen;
                            s.latestDiff = s.rv0 - s.rv1;
                            set_at(s.nDeliv, d, get_at(s.nDeliv, d) + 1u);
                        }
                        trace_rec(s, REC_RX, tC, d, 0.0, 0.0, 0.0, 0.0);
                    }
                }
            }
            // (the receivers' reception fields -- error sum, BER, reset time, section -- are dead once
            // rxOf is -1 again: a lock-on rewrites them; they are not maintained along the chain)
        }
        if (window >= 0) {
            const double timeTotal = s.annSlots * kSlot;
            const double stopW = tC + timeTotal;
            const uint32_t qW = s.seq++;
            set_at(s.stopW, window, stopW);
            set_at(s.sW, window, qW);
            set_at(s.wPend, window, 1);
            set_at(s.wDone, window, 0);
            tOther = fmin(tOther, stopW);
            mac_loop_head(s, P, B, window, ring, plant);
        }
        if (d < NS) {
            mac_loop_head(s, P, B, d, ring, plant);
        } else if (d == RRM) {
            s.tRrm = tC + (s.annSlots + 1) * kSlot;
            s.sRrm = s.seq++;
            s.rrmPend = 1;
            tOther = fmin(tOther, s.tRrm);
        } else if (NJ > 0 && s.jamPending[0] > 0) {
            s.jamPending[0] -= 1;
            phy_send_init(s, d);
        }

        // ---- chain: the only device that can have started a SEND in the completion handler; its
        // slot-start event is the next event if its transmission completes before tOther
        const int nx = (d == RRM) ? s.annDest : d;
        if (get_at(s.sphase, nx) != S_SLOT) break;
        if (nx != d) {
            set_at(s.txSeq, d, get_at(s.txSeq, d) + txs);
            txs = 0;
            d = nx;
            GW_UNROLL
            for (int p = 0; p < D; ++p) { rp[p] = srx_at<D>(srx, p, d); lastN[p] = -1.0; }
        }
        ts = get_at(s.tEv, d);
    }
    if (any) {
        set_at(s.txSeq, d, get_at(s.txSeq, d) + txs);
        GW_UNROLL
        for (int p = 0; p < D; ++p) s.P[p] = Pc[p];
        s.pchg = -1;
    }
    return any;
}

template <int D, int NS, int NJ, class ST>
GW_HD bool quiet_tail(Sim<D, NS, NJ, ST> &s, const BandParams &B)
{
    static_assert(NS == 2, "the tick logic is written for two senders per band");
    bool ok = s.rrmPend != 0;
    GW_UNROLL
    for (int p = 0; p < D; ++p) ok &= (s.sphase[p] == S_IDLE) & (s.rxOf[p] < 0);
    GW_UNROLL
    for (int k = 0; k < NS; ++k) ok &= (s.mac[k] == MAC_NONE) | (s.mac[k] == MAC_IDLE);
    GW_UNROLL
    for (int j = 0; j < NJ; ++j) ok &= s.tJam[j] > s.tRrm;
    if (!ok) return false;
    GW_STAT_MACRO(1);
    // window time-outs that precede the guard time-out (simple_stack.py:406-420): they only clear
    // flags of MACs that wait for nothing else, so neither their order nor their position among the
    // silent ticks matters; the ticks up to the guard time-out are applied in one batch
    GW_UNROLL
    for (int k = 0; k < NS; ++k) {
        if (s.wPend[k] && before(s.stopW[k], s.sW[k], s.tRrm, s.sRrm)) {
            s.wPend[k] = 0;
            s.mac[k] = MAC_NONE;
        }
    }
    // assignMessage.setProcessed() (simple_stack.py:561)
    silent_ticks_both(s, B, s.tRrm, s.sRrm, true, INFINITY);
    s.now = s.tRrm;
    s.rrmPend = 0;
    s.assignDone = 1;
    return true;
}

// SimMan.runSimulation(assignSignal.eProcessed) for an env with a plant: every tick is an event
template <int MODE, int D, int NS, int NJ, class ST, class SRX, class Ring, class Masks, class Memo, class Plant>
GW_HD void run_until_assign_plant(Sim<D, NS, NJ, ST> &s, const Params &P, const BandParams &B, const SRX &srx,
                                  const Ring &ring, const Masks &masks, const Memo &memo, Plant &plant)
{
    while (!s.assignDone && !s.fault) {
        const Event ev = next_event<true>(s, B, INFINITY, ring);
        process_event<MODE>(s, P, B, ev, srx, ring, masks, memo, plant);
    }
}

// SimMan.runSimulation(assignSignal.eProcessed) (counter_traffic.py:155)
template <int MODE, int D, int NS, int NJ, class ST, class SRX, class Ring, class Masks, class Memo = NoMemo>
GW_HD void run_until_assign(Sim<D, NS, NJ, ST> &s, const Params &P, const BandParams &B,
                            const SRX &srx, const Ring &ring, const Masks &masks, const Memo &memo = Memo(),
                            int idleAtStart = 0)
{
    const bool macros = macros_enabled<MODE, NJ>(P);
    // The common start of a step: begin_assignment() found the band-sim idle -- no PHY active, no
    // reception, no MAC busy, no window or guard time-out pending (and no interferer on the band).
    // Then the RRM's slot start is the only non-tick event: it needs no selection, and its
    // transmission passes the structural isolation test by construction.  `idleAtStart` > 0: the
    // caller knows (the kernels read it from the packed flags); 0: checked here.
    bool first = false;
    if (macros && NJ == 0) {
        if (idleAtStart > 0) {
            first = true;
        } else {
            bool idle = (s.sphase[NS] == S_SLOT) & (s.rrmPend == 0);
            GW_UNROLL
            for (int p = 0; p < D; ++p) idle &= (s.rxOf[p] < 0) & ((p == NS) | (s.sphase[p] == S_IDLE));
            GW_UNROLL
            for (int k = 0; k < NS; ++k) idle &= (s.mac[k] == MAC_NONE) & (s.wPend[k] == 0);
            first = idle;
        }
    }
    while (!s.assignDone && !s.fault) {
        Event ev;
        const bool idleStart = first;
        if (first) {
            first = false;
            ev.kind = EV_PHY; ev.idx = NS; ev.t = s.tEv[NS]; ev.seq = s.sEv[NS];
            silent_ticks_both(s, B, ev.t, ev.seq, true, INFINITY);
        } else {
            if (macros && quiet_tail(s, B)) break;
            ev = next_event(s, B, INFINITY, ring);
        }
        if (macros && ev.kind == EV_PHY && isolated_tx(s, P, B, ev, srx, ring, memo, INFINITY, idleStart)) continue;
        process_event<MODE>(s, P, B, ev, srx, ring, masks, memo);
    }
}

// another band of the same env ended its assignment later, at time T: events strictly
// before T are processed, then the clock is the env's clock
template <int MODE, int D, int NS, int NJ, class ST, class SRX, class Ring, class Masks, class Memo = NoMemo>
GW_HD void run_until_time(Sim<D, NS, NJ, ST> &s, const Params &P, const BandParams &B, const SRX &srx,
                          const Ring &ring, const Masks &masks, double T, const Memo &memo = Memo())
{
    while (!s.fault) {
        const Event ev = next_event(s, B, T, ring);
        if (!(ev.t < T)) break;
        if (macros_enabled<MODE, NJ>(P) && ev.kind == EV_PHY && isolated_tx(s, P, B, ev, srx, ring, memo, T)) continue;
        process_event<MODE>(s, P, B, ev, srx, ring, masks, memo);
    }
    s.now = T;
}

// ---------------------------------------------------------------------------
// devices moving between steps (the reference's Position.set, devices/core.py:75-84)
//
// Position.set -> nChange -> every attenuation model of the device: FsplAttenuation._update
// (attenuation_models.py:28-36; physical.py:383-386: only below STANDBY_THRESHOLD; equal positions
// keep the previous value) -> _setAttenuation (physical.py:354-362: only a NEW value triggers) ->
// SimplePhy._onAttenuationChange (simple_stack.py:119-128) of the PHYs that registered a transmission
// on that model: the stored received power of the transmission is replaced and the difference goes
// through _nReceivedPowerChanges -- the PHY's power sum changes, and a PHY that is receiving counts
// the errors of the segment that ends and, unless its transmission has completed, re-evaluates its bit
// error rate (simple_stack.py:81-86, 223-233).  Transmissions are registered from the zero-delay
// notification after their creation to their completion: exactly while the sender is in S_HDR / S_PAY.
// The models of a device are notified in Python-set order (simtools.py:255); they are visited by
// ascending partner index here, which only matters for the rounding of the MOVING PHY's own power sum
// when two or more other devices transmit at that instant.  Models are created lazily (see below):
// a pair neither of whose devices has transmitted yet has no model, its table entry follows the
// positions without the threshold / coincidence rules.
//
// `Tab` gives access to the band-sim's attenuation (dB) and received-power (mW) tables:
//   att(p, d), set_att(p, d, v), srx(p, d), set_srx(p, d, v), view() -> object for srx_at<D>()
// ---------------------------------------------------------------------------

template <int MODE, int D, int NS, int NJ, class ST, class SRX, class Masks, class Memo>
GW_HD void received_power_change(Sim<D, NS, NJ, ST> &s, const Params &P, int p, double delta, const SRX &srx,
                                 const Masks &masks, const Memo &memo)
{
    set_at(s.P, p, get_at(s.P, p) + delta);                 // updateReceivedPower (priority 1)
    s.pchg |= 1 << p;
    const int e = get_at(s.rxOf, p);
    if (e < 0 || delta == 0.0) return;                      // onReceivedPowerChange of a running reception
    if (MODE == MODE_R) {
        const double duration = s.now - get_at(s.tReset, p);
        const double bitErrors = get_at(s.ber, p) * duration * P.bitRate;
        set_at(s.err, p, get_at(s.err, p) + bitErrors);
    } else {
        int sender; uint32_t txseq; int64_t k0, k1;
        mask_range(s, p, P.bitRate, sender, txseq, k0, k1);
        if (k1 > k0) set_at(s.err, p, get_at(s.err, p) + (double)masks(p, sender, txseq, k0, k1, get_at(s.ber, p)));
        set_at(s.segT0, p, s.now);
    }
    const bool completed = s.now >= get_at(s.tStop, e);
    if (completed) return;
    const double S = srx_at<D>(srx, p, e);
    const double N = get_at(s.P, p) - S;
    if (!(S >= 0) || !(N >= 0)) { s.fault = FAULT_REF_ASSERT; return; }
    double ber;
    const int slot = memo_slot(s, p, e);
    if (!memo.get(S, N, ber, slot)) {
        ber = ber_bpsk_mw_cold(S, N, P.tenLog10BitRate, P.qDen);
        memo.put(S, N, ber, slot);
    }
    set_at(s.ber, p, ber);
    s.pchg &= ~(1 << p);
    trace_rec(s, REC_BER, s.now, p, ber, 0.0, 0.0, 0.0);
}

// `cur` [D][2]: the positions before the call, updated in place; `want` [D][2]: the requested
// positions.  Devices are moved one after the other by ascending index, like successive
// Position.set calls.  `powerDbm` [D]: transmission power of each device.
template <int MODE, int D, int NS, int NJ, class ST, class Tab, class Masks, class Memo>
GW_HD void move_devices(Sim<D, NS, NJ, ST> &s, const Params &P, const double *powerDbm, double frequency,
                        double *cur, const double *want, Tab &tab, const Masks &masks, const Memo &memo)
{
    for (int m = 0; m < D; ++m) {
        if (want[2 * m] == cur[2 * m] && want[2 * m + 1] == cur[2 * m + 1]) continue;      // Position.set: no trigger
        cur[2 * m] = want[2 * m]; cur[2 * m + 1] = want[2 * m + 1];
        for (int j = 0; j < D; ++j) {
            if (j == m) continue;
            const double dx = cur[2 * m] - cur[2 * j], dy = cur[2 * m + 1] - cur[2 * j + 1];
            const double dist = sqrt(dx * dx + dy * dy);                                    // devices/core.py:88-95
            // The model of a pair is created lazily, at the first transmission one of the two devices
            // sends (SimplePhy._getAttenuationModelByTransmission -> FrequencyBand.getAttenuationModel,
            // physical.py:576-594).  Until then nobody listens to position changes and the model will be
            // computed from the positions of that moment -- without the threshold, 0 dB for coinciding
            // devices: the table simply follows the positions (fspl_db).
            if (get_at(s.txSeq, m) == 0u && get_at(s.txSeq, j) == 0u) {
                const double fresh = (dx == 0.0 && dy == 0.0) ? 0.0 : 20 * log10(dist) + 20 * log10(frequency) - 147.55;
                tab.set_att(m, j, fresh); tab.set_att(j, m, fresh);
                tab.set_srx(j, m, rx_power_mw(powerDbm[m], fresh));
                tab.set_srx(m, j, rx_power_mw(powerDbm[j], fresh));
                continue;
            }
            if (!(dist < 3000.0)) continue;                                                 // STANDBY_THRESHOLD
            if (dx == 0.0 && dy == 0.0) continue;                                           // _update returns early
            const double att = 20 * log10(dist) + 20 * log10(frequency) - 147.55;
            if (att == tab.att(m, j)) continue;
            tab.set_att(m, j, att); tab.set_att(j, m, att);
            // the two directions of the link: (receiver j, sender m) and (receiver m, sender j)
            for (int dir = 0; dir < 2; ++dir) {
                const int p = dir == 0 ? j : m, e = dir == 0 ? m : j;
                const double rp = rx_power_mw(powerDbm[e], att);
                const int ph = get_at(s.sphase, e);
                const bool onAir = ph == S_HDR || ph == S_PAY;
                const double delta = rp - tab.srx(p, e);
                tab.set_srx(p, e, rp);
                if (onAir) received_power_change<MODE>(s, P, p, delta, tab.view(), masks, memo);
            }
        }
    }
}

// CounterTrafficEnv.reset (counter_traffic.py:135-144): sender counters := 0, interpreter
// reset; time, queues and PHY state stay.  Queued packets keep the sizes they were enqueued
// with: they are materialised into the snapshot ring (`ringw(sender, slot, size)`) before
// the counter epoch changes.
template <int D, int NS, int NJ, class ST, class RingW>
GW_HD void reset_sim(Sim<D, NS, NJ, ST> &s, const BandParams &B, RingW &ringw)
{
    GW_UNROLL
    for (int k = 0; k < NS; ++k) {
        const uint64_t m = (uint64_t)B.mult[k];
        const uint64_t enq = s.ticks[k] * m;
        if (B.payloadRule[k] < 0) {
            uint64_t j = enq - (uint64_t)s.qn[k];
            if (j < s.epochK[k] * m) j = s.epochK[k] * m;           // already materialised by an earlier reset()
            for (; j < enq; ++j) {
                const uint64_t tick = j / m;
                const uint64_t c = (uint64_t)s.epochC[k] + (tick - s.epochK[k]);
                ringw(k, (uint32_t)j & (uint32_t)(kRingSlots - 1), c > (uint64_t)kCounterBound ? kCounterBound : (int)c);
            }
        }
        s.epochK[k] = s.ticks[k];
        s.epochC[k] = 0;
    }
    s.latestDiff = 0; s.lastAbsDiff = 0; s.rv0 = 0; s.rv1 = 0; s.done = 0;   // counter_traffic.py:69-73
}

// Interpreter.getFeedback (envs/core.py:142-153, counter_traffic.py:85-107)
template <int D, int NS, int NJ, class ST>
GW_HD void feedback(Sim<D, NS, NJ, ST> &s, long long &obs, double &reward, unsigned char &done)
{
    obs = (long long)s.latestDiff + kCounterBound;
    const int absd = s.latestDiff < 0 ? -s.latestDiff : s.latestDiff;
    int r = s.lastAbsDiff - absd;
    s.lastAbsDiff = absd;
    if (r > 10) r = 10; else if (r < -10) r = -10;
    reward = (double)r;
    done = (unsigned char)s.done;
}

}  // namespace gw
