// gymwipe_b200 -- general band engine: bands beyond CounterTrafficEnv's 2 senders + RRM (+ 1 PHY-only sender)
// template (the generalised scenario compiler).
//
// A band of up to kGenMaxSend MAC senders (SimpleNetworkDevice + a traffic process, any of them in MAC receive
// mode or with a finite burst), ONE RRM (SimpleRrmDevice + CounterTrafficInterpreter) and up to kGenMaxJam PHY-only
// periodic senders, all as RUN-TIME counts, stepped like the env: assignFrequencyBand(device, duration), then
// runSimulation(assignSignal.eProcessed), then the interpreter's feedback (reference accounting, "mode R").
//
// Same method as the step kernel's core (gw_core.cuh): one band-sim per thread, the SimPy heap replaced by timed
// event slots ordered by (time, creation number) -- per sender a traffic tick, a window time-out and a RECEIVE
// time-out, per device one PHY event (slot start -> header end -> completion), per PHY-only sender a wake-up, the
// RRM's guard time-out -- with the zero-delay event chains executed inline in SimPy's pop order.  Unlike the step
// kernel nothing is unrolled or held in registers: every per-device array lives in the band-sim's state in global
// memory ([field][index][sim]: consecutive threads touch consecutive words) and is indexed directly, and the
// received-power changes are counted where they happen.  This is the
// general engine, not the tuned one: the 2-sender + RRM template keeps its own kernels.
//
// Reference semantics (file:line under /root/reference):
//   SimplePhy  (power bookkeeping, BER accounting, decider)  networking/simple_stack.py:77-286
//   SimpleMac  (queue, assignment-window loop, receive mode)  networking/simple_stack.py:386-484
//   SimpleRrmMac (announcement, guard slot)                   networking/simple_stack.py:527-561
//   SimpleNetworkDevice receive loop / SimpleRrmDevice        networking/devices.py:66-111, 113-203
//   Transmission / FrequencyBand.transmit                     networking/physical.py:224-290, 576-608
//   SenderDevice.senderProcess, CounterTrafficInterpreter     envs/counter_traffic.py:53-61, 63-112 (the interpreter
//       keeps one received value per device; its observation is receivedValues[0] - receivedValues[1])
//   PHY-only periodic sender                                  tests/test_benchmark.py:20-50
//
// Plain C++ (no CUDA intrinsics): gw_kernels.cu includes it for the device, tests/hostsim for the host.
#pragma once

#include "gw_core.cuh"

namespace gw {

constexpr int kGenMaxSend = 8, kGenMaxJam = 16, kGenMaxDev = kGenMaxSend + 1 + kGenMaxJam;
constexpr int EV_MOVE = 7;              // a mobility process wakes up (next to gw_core.cuh's EV_* kinds)

// band configuration, common to all band-sims of a handle; device order: senders, RRM, PHY-only senders
struct GenBand {
    int ns, nj, nd;
    int maxDuration;                    // action space: duration in [0, maxDuration)
    int mode;                           // MODE_R (reference accounting) or MODE_M_PHILOX (per-bit Philox error masks)
    unsigned long long seed;            // mode M: Philox seed
    long long envOffset;                // mode M: global id of env 0 (sharding keeps results invariant)
    double thermal;
    double frequency;
    double power[kGenMaxDev];           // transmission power (dBm) per device: 0 for MACs and the RRM
    int maxMoves;                       // mobility processes: jumps per device in the offset tape (0: none)
    double moveInterval;
    int mult[kGenMaxSend];
    int payloadRule[kGenMaxSend];       // -1: byteSize = counter
    int dest[kGenMaxSend];              // device index of the sender its packets are addressed to
    int maxTicks[kGenMaxSend];          // 0: the traffic process runs forever; n: a burst of n ticks
    int recv[kGenMaxSend];              // 1: MAC receive mode (SimpleNetworkDevice.receiving = True)
    double interval[kGenMaxSend];
    double jamInterval[kGenMaxJam], jamDelay[kGenMaxJam];
    int jamHdr[kGenMaxJam], jamPay[kGenMaxJam];
};

// words of one band-sim's state
GW_HD int gen_f64_words(int ns, int nj) { return 4 + 9 * (ns + 1 + nj) + 3 * ns + nj; }
GW_HD int gen_i32_words(int ns, int nj) { return 18 + 7 * (ns + 1 + nj) + 12 * ns + 3 * nj + kQueueCap * ns; }

// view of one band-sim: word w of the fp64 / int32 state at f[w * stride] / i[w * stride]
struct GenView {
    double *f;
    int32_t *i;
    long long stride;
    double *srx;                        // received power (mW), entry (receiver p, sender d) at srx[(p * nd + d) * srxStride]
    long long srxStride;
    double *att;                        // attenuation (dB), same layout, and current positions [nd][2] at pos[(2 * d + c) * srxStride]:
    double *pos;                        //   held for per-env geometries only (devices that move between steps), else null
    // mobility processes (tests/test_benchmark.py:73-85), per-env geometries only; mvT == null: none
    double *mvT, *mvDelay;              // next wake-up, first delay; entry d at [d * srxStride]
    int32_t *mvI;                       // creation number, stage (0 first delay, 1 moving, 2 off), jumps done: [3][nd]
    const double *offsets;              // this env's tape [nd][maxMoves][2] (accumulating jumps)
    int ns, nj, nd;
    long long env;                      // global env id (mode M: key of the error masks)
    int mode;                           // MODE_R / MODE_M_PHILOX: a compile-time constant in the kernels (no mode-M code or stores in mode R)
    double *trace;                      // optional event trace (records of 8 doubles, as gw_core.cuh::trace_rec)
    int ntrace, traceCap;

#define GEN_F(name, base) GW_HD double &name(int k) const { return f[(long long)((base) + k) * stride]; }
#define GEN_I(name, base) GW_HD int32_t &name(int k) const { return i[(long long)((base) + k) * stride]; }
#define GEN_U(name, base) GW_HD uint32_t &name(int k) const { return ((uint32_t *)i)[(long long)((base) + k) * stride]; }
    // scalars
    GW_HD double &now() const { return f[0]; }
    GW_HD double &tRrm() const { return f[stride]; }
    GW_HD double &annSlots() const { return f[2 * stride]; }
    GW_HD double &tTickMin() const { return f[3 * stride]; }    // a lower bound of the senders' next tick times (see gen_next_event)
    // PHY, per device (names as in gw_core.cuh::Sim)
    GEN_F(P, 4) GEN_F(tEv, 4 + nd) GEN_F(tStop, 4 + 2 * nd) GEN_F(tC, 4 + 3 * nd) GEN_F(ber, 4 + 4 * nd)
    GEN_F(err, 4 + 5 * nd) GEN_F(tReset, 4 + 6 * nd)
    GEN_F(txStart, 4 + 7 * nd) GEN_F(segT0, 4 + 8 * nd)        // mode M: Transmission.startTime; start of the running segment
    // senders
    GEN_F(tTick, 4 + 9 * nd) GEN_F(stopW, 4 + 9 * nd + ns) GEN_F(rxT, 4 + 9 * nd + 2 * ns)
    // PHY-only senders
    GEN_F(tJam, 4 + 9 * nd + 3 * ns)

    enum : int { I_seq = 0, I_fault, I_ties, I_annDest, I_rrmPend, I_sRrm, I_assignDone, I_rv0, I_rv1, I_latestDiff,
                 I_lastAbsDiff, I_done, I_nTx,
                 // summaries of the per-device / per-sender arrays, so that the event selection and the receiver loops of
                 // the transition function visit only the entries that matter (the state lives in global memory):
                 I_phyMask,             // bit d: PHY d has a timed event pending (sphase >= S_SLOT)
                 I_rxMask,              // bit p: PHY p is receiving (rxOf >= 0)
                 I_wMask,               // bit k: sender k's window time-out is pending
                 I_condMask,            // bit k: sender k's MAC waits for a packet (MAC_WAIT_COND)
                 I_rvMask,              // bit k: the interpreter's receivedValues[k] is set (payload.value = 2; 0 after reset)
                 kScalars };
    GW_HD int32_t &sc(int w) const { return i[(long long)w * stride]; }
    GW_HD uint32_t &seq() const { return ((uint32_t *)i)[(long long)I_seq * stride]; }
    GEN_I(sphase, 18) GEN_U(sEv, 18 + nd) GEN_U(sC, 18 + 2 * nd) GEN_I(cmdPay, 18 + 3 * nd) GEN_I(rxOf, 18 + 4 * nd)
    GEN_I(rxSec, 18 + 5 * nd) GEN_U(txSeq, 18 + 6 * nd)
    GEN_U(sTick, 18 + 7 * nd) GEN_U(epochK, 18 + 7 * nd + ns) GEN_I(qn, 18 + 7 * nd + 2 * ns) GEN_I(mac, 18 + 7 * nd + 3 * ns)
    GEN_I(wDone, 18 + 7 * nd + 4 * ns) GEN_I(wPend, 18 + 7 * nd + 5 * ns) GEN_U(sW, 18 + 7 * nd + 6 * ns)
    GEN_U(rxS, 18 + 7 * nd + 7 * ns) GEN_U(nDeliv, 18 + 7 * nd + 8 * ns) GEN_U(nRecv, 18 + 7 * nd + 9 * ns)
    GEN_I(epochC, 18 + 7 * nd + 10 * ns) GEN_U(ticks, 18 + 7 * nd + 11 * ns)
    GEN_U(sJam, 18 + 7 * nd + 12 * ns) GEN_I(jamStage, 18 + 7 * nd + 12 * ns + nj) GEN_I(jamPending, 18 + 7 * nd + 12 * ns + 2 * nj)
    GW_HD int32_t &ring(int k, int slot) const { return i[(long long)(18 + 7 * nd + 12 * ns + 3 * nj + k * kQueueCap + slot) * stride]; }
#undef GEN_F
#undef GEN_I
#undef GEN_U
    GW_HD double rp(int p, int d) const { return srx[(long long)(p * nd + d) * srxStride]; }
    GW_HD double &rpw(int p, int d) const { return srx[(long long)(p * nd + d) * srxStride]; }
    GW_HD double &attw(int p, int d) const { return att[(long long)(p * nd + d) * srxStride]; }
    GW_HD double &posw(int d, int c) const { return pos[(long long)(2 * d + c) * srxStride]; }
    GW_HD double &tMove(int d) const { return mvT[(long long)d * srxStride]; }
    GW_HD double &moveDelay(int d) const { return mvDelay[(long long)d * srxStride]; }
    GW_HD uint32_t &sMove(int d) const { return ((uint32_t *)mvI)[(long long)d * srxStride]; }
    GW_HD int32_t &moveStage(int d) const { return mvI[(long long)(nd + d) * srxStride]; }
    GW_HD int32_t &moveK(int d) const { return mvI[(long long)(2 * nd + d) * srxStride]; }
};

static_assert(GenView::kScalars == 18, "scalar block of the int32 state");

GW_HD void gen_rec(GenView &v, int kind, double t, int dev, double x0, double x1, double x2, double x3)
{
    if (v.trace == nullptr) return;
    if (v.ntrace < v.traceCap) {
        double *r = v.trace + (long long)v.ntrace * 8;
        r[0] = kind; r[1] = t; r[2] = dev; r[3] = x0; r[4] = x1; r[5] = x2; r[6] = x3; r[7] = 0;
    }
    v.ntrace += 1;
}

// received-power table of a geometry: pos [nd][2], power [nd] (dBm) -> srx [nd * nd] with the given stride
// (FsplAttenuation._update, attenuation_models.py:28-36; dbmToMilliwatts(power - attenuation), simple_stack.py:111)
GW_HD void gen_power_table(int nd, const double *pos, const double *power, double frequency, double *srx, long long stride,
                           double *att = nullptr, double *posOut = nullptr)
{
    for (int p = 0; p < nd; ++p)
        for (int d = 0; d < nd; ++d) {
            double a = 0.0, rp = 0.0;
            if (p != d) { a = fspl_db(pos[2 * p], pos[2 * p + 1], pos[2 * d], pos[2 * d + 1], frequency); rp = rx_power_mw(power[d], a); }
            srx[(long long)(p * nd + d) * stride] = rp;
            if (att) att[(long long)(p * nd + d) * stride] = a;
        }
    if (posOut) for (int k = 0; k < 2 * nd; ++k) posOut[(long long)k * stride] = pos[k];
}

// construction-time state (counter_traffic.py:114-133; the harness scenario of N senders): process Initialize events
// in construction order -- senders, then PHY-only senders --, then `receiving = True` in sender order (devices.py:77-84)
GW_HD void gen_init(GenView &v, const GenBand &B)
{
    const int ns = v.ns, nj = v.nj, nd = v.nd;
    v.now() = 0.0; v.tRrm() = 0.0; v.annSlots() = 0.0; v.tTickMin() = 0.0;
    for (int w = 0; w < GenView::kScalars; ++w) v.sc(w) = 0;
    for (int p = 0; p < nd; ++p) {
        v.P(p) = B.thermal; v.tEv(p) = 0; v.tStop(p) = 0; v.tC(p) = 0; v.ber(p) = 0; v.err(p) = 0; v.tReset(p) = 0;
        v.txStart(p) = 0; v.segT0(p) = 0;
        v.sphase(p) = S_IDLE; v.sEv(p) = 0; v.sC(p) = 0; v.cmdPay(p) = 0; v.rxOf(p) = -1; v.rxSec(p) = 0; v.txSeq(p) = 0;
    }
    for (int k = 0; k < ns; ++k) {
        v.tTick(k) = 0.0; v.sTick(k) = v.seq()++;
        v.stopW(k) = 0; v.epochK(k) = 0; v.qn(k) = 0; v.mac(k) = MAC_NONE; v.wDone(k) = 0; v.wPend(k) = 0; v.sW(k) = 0;
        v.nDeliv(k) = 0; v.nRecv(k) = 0; v.epochC(k) = 1; v.ticks(k) = 0;
        for (int q = 0; q < kQueueCap; ++q) v.ring(k, q) = 0;
    }
    for (int j = 0; j < nj; ++j) { v.tJam(j) = 0.0; v.sJam(j) = v.seq()++; v.jamStage(j) = 0; v.jamPending(j) = 0; }
    for (int k = 0; k < ns; ++k) {
        if (B.recv[k]) { v.rxT(k) = 0.0; v.rxS(k) = v.seq()++; }
        else { v.rxT(k) = (double)INFINITY; v.rxS(k) = 0; }
    }
}

// The sender queues are not stored: sender k enqueues `mult` packets per tick into a drop-oldest deque(maxlen = 100)
// (SenderDevice.senderProcess, counter_traffic.py:53-61; SimpleMac.networkInHandler, simple_stack.py:463-471), so
// after `ticks` ticks the queue holds the LAST qn of the ticks * mult packets enqueued so far, and packet j was
// enqueued at tick j / mult with byteSize = the counter of that tick: min(COUNTER_BOUND, epochC + (tick - epochK)),
// where (epochK, epochC) is the counter epoch -- (0, 1) after construction, (ticks at reset, 0) after a reset().
// Packets that predate the last reset() keep the sizes they were enqueued with: reset materialises them into the
// ring (slot j % 100: the queue holds at most 100 consecutive packets).  As in gw_core.cuh::head_size.
// `ticks` is a 32-bit count: 2^32 ticks of 1 ms are 49 simulated days.
GW_HD int gen_counter_at(const GenView &v, int k, uint32_t tick)
{
    const long long c = (long long)v.epochC(k) + (long long)(tick - v.epochK(k));
    return c > kCounterBound ? kCounterBound : (int)c;
}

GW_HD int gen_head_size(const GenView &v, const GenBand &B, int k)
{
    const int rule = B.payloadRule[k];
    if (rule >= 0) return rule;
    const unsigned long long m = (unsigned long long)B.mult[k];
    const unsigned long long enq = (unsigned long long)v.ticks(k) * m;
    const unsigned long long j = enq - (unsigned long long)v.qn(k);
    if (j < (unsigned long long)v.epochK(k) * m) return v.ring(k, (int)(j % kQueueCap));      // predates the last reset()
    return gen_counter_at(v, k, (uint32_t)(j / m));
}

// CounterTrafficEnv.reset (counter_traffic.py:135-144): sender counters := 0, interpreter reset; time, queues
// (with the sizes their packets were enqueued with) and PHY state stay
GW_HD void gen_reset(GenView &v, const GenBand &B)
{
    for (int k = 0; k < v.ns; ++k) {
        const unsigned long long m = (unsigned long long)B.mult[k];
        const unsigned long long enq = (unsigned long long)v.ticks(k) * m;
        if (B.payloadRule[k] < 0) {
            unsigned long long j = enq - (unsigned long long)v.qn(k);
            if (j < (unsigned long long)v.epochK(k) * m) j = (unsigned long long)v.epochK(k) * m;   // materialised by an earlier reset()
            for (; j < enq; ++j) v.ring(k, (int)(j % kQueueCap)) = gen_counter_at(v, k, (uint32_t)(j / m));
        }
        v.epochK(k) = v.ticks(k);
        v.epochC(k) = 0;
    }
    v.sc(GenView::I_latestDiff) = 0; v.sc(GenView::I_lastAbsDiff) = 0; v.sc(GenView::I_rv0) = 0; v.sc(GenView::I_rv1) = 0;
    v.sc(GenView::I_rvMask) = 0;
    v.sc(GenView::I_done) = 0;
}

// `c` ticks of sender k's traffic process at once
GW_HD void gen_ticks(GenView &v, const GenBand &B, int k, uint32_t c)
{
    const unsigned long long n = (unsigned long long)v.qn(k) + (unsigned long long)c * (unsigned long long)B.mult[k];
    v.qn(k) = n > (unsigned long long)kQueueCap ? kQueueCap : (int)n;         // drop-oldest
    v.ticks(k) += c;
}

GW_HD int gen_ctz(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return __ffs((int)x) - 1;
#else
    return __builtin_ctz(x);
#endif
}

// state changes that the summary masks follow
GW_HD void gen_set_sphase(GenView &v, int d, int ph)
{
    v.sphase(d) = ph;
    const int bit = 1 << d;
    if (ph >= S_SLOT) v.sc(GenView::I_phyMask) |= bit; else v.sc(GenView::I_phyMask) &= ~bit;
}
GW_HD void gen_set_mac(GenView &v, int k, int m)
{
    v.mac(k) = m;
    const int bit = 1 << k;
    if (m == MAC_WAIT_COND) v.sc(GenView::I_condMask) |= bit; else v.sc(GenView::I_condMask) &= ~bit;
}

// Silent ticks of sender k strictly before (tEnd, qEnd): the tick times are accumulated with the reference's fp64
// additions, one per tick (counter_traffic.py:61); see gw_core.cuh::silent_ticks
GW_HD void gen_silent_ticks(GenView &v, const GenBand &B, int k, double tEnd, uint32_t qEnd)
{
    double t = v.tTick(k);
    if (!before(t, v.sTick(k), tEnd, qEnd)) return;
    const double interval = B.interval[k];
    uint32_t c = 0;
    do {
        t = t + interval;
        ++c;
    } while (t < tEnd);
    v.sc(GenView::I_ties) += (t == tEnd) ? 1 : 0;       // exact tie of independent events (diagnostic)
    v.tTick(k) = t;
    gen_ticks(v, B, k, c);
    v.seq() += c;
    v.sTick(k) = v.seq() - 1u;
}

// earliest (time, creation number) among the timed slots; silent ticks in front of it are applied on the way
GW_HD Event gen_next_event(GenView &v, const GenBand &B)
{
    const int ns = v.ns, nj = v.nj, nd = v.nd;
    Event e;
    e.kind = EV_NONE; e.idx = 0; e.t = INFINITY; e.seq = 0;
#define GEN_CONSIDER(T, SQ, K, I)                                                       \
    do {                                                                                \
        const double t_ = (T);                                                          \
        const uint32_t q_ = (SQ);                                                       \
        if (e.kind == EV_NONE || before(t_, q_, e.t, e.seq)) {                          \
            e.t = t_; e.seq = q_; e.kind = (K); e.idx = (I);                            \
        }                                                                               \
    } while (0)
    (void)nd;
    for (int j = 0; j < nj; ++j) GEN_CONSIDER(v.tJam(j), v.sJam(j), EV_JAM, j);
    for (uint32_t m = (uint32_t)v.sc(GenView::I_phyMask); m; m &= m - 1u) {
        const int d = gen_ctz(m);
        GEN_CONSIDER(v.tEv(d), v.sEv(d), EV_PHY, d);
    }
    const uint32_t wMask = (uint32_t)v.sc(GenView::I_wMask), condMask = (uint32_t)v.sc(GenView::I_condMask);
    for (int k = 0; k < ns; ++k) {
        if ((wMask >> k) & 1u) GEN_CONSIDER(v.stopW(k), v.sW(k), EV_W, k);
        if (B.recv[k]) GEN_CONSIDER(v.rxT(k), v.rxS(k), EV_RXTO, k);
        if ((B.maxTicks[k] != 0 || ((condMask >> k) & 1u)) && v.tTick(k) < (double)INFINITY) GEN_CONSIDER(v.tTick(k), v.sTick(k), EV_TICK, k);
    }
    if (v.sc(GenView::I_rrmPend)) GEN_CONSIDER(v.tRrm(), (uint32_t)v.sc(GenView::I_sRrm), EV_RRM, 0);
    if (v.mvT != nullptr)
        for (int d = 0; d < v.nd; ++d)
            if (v.moveStage(d) < 2) GEN_CONSIDER(v.tMove(d), v.sMove(d), EV_MOVE, d);
#undef GEN_CONSIDER
    if (e.kind == EV_NONE) return e;
    // Silent ticks in front of the event.  Ticks of different senders touch only their own sender's queue, so the
    // senders are advanced one after the other.  tTickMin is a lower bound of all next tick times (tick times only
    // grow): an event strictly before it has no tick in front of it, and the senders are not visited at all.
    if (!(e.t < v.tTickMin())) {
        double lo = INFINITY;
        for (int k = 0; k < ns; ++k) {
            if (!(B.maxTicks[k] != 0 || ((condMask >> k) & 1u))) gen_silent_ticks(v, B, k, e.t, e.seq);
            lo = fmin(lo, v.tTick(k));
        }
        v.tTickMin() = lo;
    }
    return e;
}

// SimplePhy._updateBitErrorRate (simple_stack.py:161-173)
GW_HD void gen_update_ber(GenView &v, const Params &P, int p)
{
    const int e = v.rxOf(p);
    if (e < 0) return;
    const double S = v.rp(p, e);
    const double N = v.P(p) - S;
    if (!(S >= 0) || !(N >= 0)) { v.sc(GenView::I_fault) = FAULT_REF_ASSERT; return; }     // simple_stack.py:168-169
    const double b = ber_bpsk_mw_cold(S, N, P.tenLog10BitRate, P.qDen);
    v.ber(p) = b;
    gen_rec(v, REC_BER, v.now(), p, b, 0.0, 0.0, 0.0);
}

// mode M: the on-air bits [k0, k1) of the segment of PHY p's reception that ends now, and what keys their error flags
GW_HD void gen_mask_range(const GenView &v, const Params &P, int p, int &sender, uint32_t &txseq, long long &k0, long long &k1)
{
    const int e = v.rxOf(p);
    const double start = v.txStart(e);
    sender = e; txseq = v.txSeq(e) - 1u;
    k0 = (long long)floor((v.segT0(p) - start) * P.bitRate);
    k1 = (long long)floor((v.now() - start) * P.bitRate);
}

// SimplePhy._countBitErrors (simple_stack.py:180-188).  Mode R: expected-value accounting, duration from the last
// RESET (appendix B #5).  Mode M: the error flags of the on-air bits since the last count -- bit k of transmission
// `txseq` of device `sender` as seen by `p` in GLOBAL env `env` is an error iff its Philox word is below
// floor(ber * 2^32) (gw_core.cuh::mask_words4) --, counted here bit by bit; the kernel counts the ranges of an event
// with the whole warp BEFORE the transition function runs (gen_count_set), which then finds them empty.
GW_HD void gen_count(GenView &v, const Params &P, const GenBand &B, int p)
{
    if (v.mode == MODE_R) {
        const double duration = v.now() - v.tReset(p);
        const double bitErrors = v.ber(p) * duration * P.bitRate;
        v.err(p) += bitErrors;
        return;
    }
    int sender; uint32_t txseq; long long k0, k1;
    gen_mask_range(v, P, p, sender, txseq, k0, k1);
    if (k1 > k0) v.err(p) += (double)mask_errors_serial(B.seed, v.env, 0, sender, txseq, p, k0, k1, v.ber(p));
    v.segT0(p) = v.now();
}

// _nReceivedPowerChanges.trigger(delta): the power sum, then the running reception (simple_stack.py:81-86, 223-233)
// Returns true if the PHY's bit error rate must be re-evaluated (the caller collects these PHYs: the evaluations
// of an event read nothing that the rest of the event writes, so they are done after it -- by the warp as a whole
// in the kernel, see gw_kernels.cu::genband_step_kernel).
GW_HD bool gen_power_change(GenView &v, const Params &P, const GenBand &B, int p, double delta, bool completingOwn)
{
    v.P(p) += delta;
    if (!(((uint32_t)v.sc(GenView::I_rxMask) >> p) & 1u) || delta == 0.0) return false;
    const int e = v.rxOf(p);
    gen_count(v, P, B, p);
    const bool completed = v.now() >= v.tStop(e);
    if (completed) return false;
    // `if not t.completed: _updateBitErrorRate(t)` with the power entry of its own transmission already popped:
    // the reference raises KeyError (appendix B #12)
    if (completingOwn) { v.sc(GenView::I_fault) = FAULT_REF_KEYERROR; return false; }
    return true;
}

// SimplePhy._updateBitErrorRate for the PHYs in berMask
GW_HD void gen_update_bers(GenView &v, const Params &P, uint32_t berMask)
{
    for (int p = 0; p < v.nd; ++p)
        if ((berMask >> p) & 1u) gen_update_ber(v, P, p);
}

GW_HD void gen_rx_clear(GenView &v, int p)
{
    v.rxOf(p) = -1; v.err(p) = 0.0; v.ber(p) = 0.0; v.tReset(p) = v.now();
    v.sc(GenView::I_rxMask) &= ~(1 << p);
    if (v.mode != MODE_R) v.segT0(p) = v.now();
}

GW_HD bool gen_decide(GenView &v, const Params &P, int p, int section, double totalBits)
{
    const bool ok = within_max_ber(P, v.err(p), totalBits);         // simple_stack.py:274-277
    gen_rec(v, REC_DEC, v.now(), p, section, v.err(p), totalBits, ok ? 1.0 : 0.0);
    return ok;
}

GW_HD void gen_begin_slot_wait(GenView &v, int d)
{
    // self._transmitting = True; yield SimMan.nextTimeSlot(TIME_SLOT_LENGTH)  (simple_stack.py:202-204, simtools.py:53)
    v.tEv(d) = v.now() + (kSlot - fmod_slot(v.now()));
    v.sEv(d) = v.seq()++;
    gen_set_sphase(v, d, S_SLOT);
}

GW_HD void gen_phy_send_init(GenView &v, int d)
{
    if (((uint32_t)v.sc(GenView::I_rxMask) >> d) & 1u) gen_set_sphase(v, d, S_WAITRX);  // yield self._nReceivingFinished.event
    else gen_begin_slot_wait(v, d);
}

// one pass of the SimpleMac window loop body with a non-empty queue (simple_stack.py:417-434)
GW_HD void gen_mac_try_send(GenView &v, const Params &P, const GenBand &B, int k)
{
    const int size = gen_head_size(v, B, k);
    const double timeLeft = v.stopW(k) - v.now();
    const double txTime = airtime_of(P, kMacHdr + kNetHdr + size);
    if (!(timeLeft > txTime)) { gen_set_mac(v, k, MAC_IDLE); return; }     // yield timeoutEvent
    v.qn(k) -= 1;
    gen_set_mac(v, k, MAC_WAIT_TX);
    v.cmdPay(k) = kNetHdr + size;
    gen_phy_send_init(v, k);
}

// loop head of the window loop (simple_stack.py:408-416)
GW_HD void gen_mac_loop_head(GenView &v, const Params &P, const GenBand &B, int k)
{
    if (v.wDone(k)) { gen_set_mac(v, k, MAC_NONE); return; }
    if (v.qn(k) == 0) { gen_set_mac(v, k, MAC_WAIT_COND); return; }
    gen_mac_try_send(v, P, B, k);
}

// Position.set of device m (devices/core.py:75-84) and everything it triggers -- the run-time-count form of
// gw_core.cuh::move_devices, see the comments there: models beyond STANDBY_THRESHOLD or with coinciding devices keep
// their value (physical.py:383-386, attenuation_models.py:31-33), only a NEW value triggers (physical.py:354-362), a
// pair neither of whose devices has transmitted yet has no model (its table entry follows the positions), and
// transmissions that are on the air go through SimplePhy._onAttenuationChange (simple_stack.py:119-128): the PHY's
// power sum changes, a running reception counts the segment that ends and re-evaluates its bit error rate.
GW_HD void gen_move_device(GenView &v, const Params &P, const GenBand &B, int m, double x, double y)
{
    const int nd = v.nd;
    if (x == v.posw(m, 0) && y == v.posw(m, 1)) return;                     // Position.set: no trigger
    v.posw(m, 0) = x; v.posw(m, 1) = y;
    for (int j = 0; j < nd; ++j) {
        if (j == m) continue;
        const double dx = x - v.posw(j, 0), dy = y - v.posw(j, 1);
        const double dist = sqrt(dx * dx + dy * dy);                        // devices/core.py:88-95
        if (v.txSeq(m) == 0u && v.txSeq(j) == 0u) {
            const double fresh = (dx == 0.0 && dy == 0.0) ? 0.0 : 20 * log10(dist) + 20 * log10(B.frequency) - 147.55;
            v.attw(m, j) = fresh; v.attw(j, m) = fresh;
            v.rpw(j, m) = rx_power_mw(B.power[m], fresh);
            v.rpw(m, j) = rx_power_mw(B.power[j], fresh);
            continue;
        }
        if (!(dist < 3000.0)) continue;                                     // STANDBY_THRESHOLD
        if (dx == 0.0 && dy == 0.0) continue;                               // _update returns early
        const double att = 20 * log10(dist) + 20 * log10(B.frequency) - 147.55;
        if (att == v.attw(m, j)) continue;
        v.attw(m, j) = att; v.attw(j, m) = att;
        for (int dir = 0; dir < 2; ++dir) {                                 // (receiver j, sender m), (receiver m, sender j)
            const int p = dir == 0 ? j : m, e = dir == 0 ? m : j;
            const double rp = rx_power_mw(B.power[e], att);
            const int ph = v.sphase(e);
            const double delta = rp - v.rp(p, e);
            v.rpw(p, e) = rp;
            if (ph == S_HDR || ph == S_PAY) {
                if (gen_power_change(v, P, B, p, delta, false)) gen_update_ber(v, P, p);
            }
        }
    }
}

// Mobility processes (the mover of tests/test_benchmark.py:73-85, one per device with moveDelays[d] >= 0): started
// now, in device order -- each an Initialize event that the engine sees as a wake-up at the current time.
GW_HD void gen_start_movers(GenView &v, const double *moveDelays)
{
    for (int d = 0; d < v.nd; ++d) {
        const bool on = moveDelays[d] >= 0.0;
        v.tMove(d) = v.now(); v.moveDelay(d) = on ? moveDelays[d] : 0.0;
        v.sMove(d) = on ? v.seq()++ : 0u; v.moveStage(d) = on ? 0 : 2; v.moveK(d) = 0;
    }
}

GW_HD int gen_hdr_bytes(const GenView &v, const GenBand &B, int d) { return d > v.ns ? B.jamHdr[d - v.ns - 1] : kMacHdr; }

// transition function: one timed event (the structure of gw_core.cuh::apply_event with run-time device counts);
// returns the set of PHYs whose bit error rate must be re-evaluated afterwards (SimplePhy._updateBitErrorRate)
GW_HD uint32_t gen_apply(GenView &v, const Params &P, const GenBand &B, const Event &ev)
{
    const int ns = v.ns, nd = v.nd, RRM = v.ns;
    uint32_t berMask = 0;
    v.now() = ev.t;
    switch (ev.kind) {
    case EV_TICK: {
        const int k = ev.idx;
        if (B.maxTicks[k] != 0 && v.ticks(k) >= (uint32_t)B.maxTicks[k]) {
            // the burst is over: this wake-up only ends the traffic process (its process event takes a number)
            v.tTick(k) = (double)INFINITY;
            v.seq()++;
            break;
        }
        gen_ticks(v, B, k, 1u);
        v.tTick(k) = v.now() + B.interval[k];
        v.sTick(k) = v.seq()++;
        if ((v.sc(GenView::I_condMask) >> k) & 1) gen_mac_try_send(v, P, B, k);     // _packetAddedEvent wakes the window loop
        break;
    }
    case EV_JAM: {
        const int j = ev.idx, d = RRM + 1 + j;
        if (v.jamStage(j) == 0) {                                   // yield timeout(initialDelay)
            v.jamStage(j) = 1; v.tJam(j) = v.now() + B.jamDelay[j]; v.sJam(j) = v.seq()++;
        } else if (v.jamStage(j) == 1) {                            // first yield timeout(sendInterval)
            v.jamStage(j) = 2; v.tJam(j) = v.now() + B.jamInterval[j]; v.sJam(j) = v.seq()++;
        } else {
            // macIn.send(SEND) -> queued executor; then yield timeout(sendInterval)
            v.tJam(j) = v.now() + B.jamInterval[j]; v.sJam(j) = v.seq()++;
            if (v.sphase(d) != S_IDLE) { v.jamPending(j) += 1; if (v.jamPending(j) > 60) v.sc(GenView::I_fault) = FAULT_SENDQ; }
            else { v.cmdPay(d) = B.jamPay[j]; gen_phy_send_init(v, d); }
        }
        break;
    }
    case EV_PHY: {
        const int d = ev.idx;
        const int ph = v.sphase(d);
        if (ph == S_SLOT) {
            // FrequencyBand.transmit -> Transmission.__init__ (physical.py:224-279, 596-608)
            const int payBytes = v.cmdPay(d), hdrBytes = gen_hdr_bytes(v, B, d);
            const double now = v.now();
            const double hd = airtime_of(P, hdrBytes);
            const double pd = airtime_of(P, payBytes);
            const double duration = hd + pd;
            const double stop = now + duration;
            const double headerStop = now + hd;
            const double tH = now + (headerStop > now ? headerStop - now : 0.0);       // timeoutUntil
            const double tC = now + (stop > now ? stop - now : 0.0);
            const uint32_t qH = v.seq()++, qC = v.seq()++;
            gen_set_sphase(v, d, S_HDR); v.tEv(d) = tH; v.sEv(d) = qH; v.tC(d) = tC; v.sC(d) = qC; v.tStop(d) = stop;
            if (v.mode != MODE_R) v.txStart(d) = now;
            v.txSeq(d) += 1u;
            v.sc(GenView::I_nTx) += 1;
            gen_rec(v, REC_TX, now, d, stop, (hdrBytes * 8) * P.bitsFactor, (payBytes * 8) * P.bitsFactor, 0.0);
            // zero-delay notification: every other PHY registers the received power (simple_stack.py:130-144)
            for (int p = 0; p < nd; ++p) {
                if (p == d) continue;
                if (gen_power_change(v, P, B, p, v.rp(p, d), false)) berMask |= 1u << p;
            }
            // receive processes in PHY construction order: idle, non-transmitting PHYs lock on (simple_stack.py:214-235)
            const uint32_t all = nd >= 32 ? 0xffffffffu : (1u << nd) - 1u;
            const uint32_t lock = all & ~(uint32_t)v.sc(GenView::I_rxMask) & ~(uint32_t)v.sc(GenView::I_phyMask) & ~(1u << d);
            v.sc(GenView::I_rxMask) |= (int)lock;
            for (uint32_t m = lock; m; m &= m - 1u) {
                const int p = gen_ctz(m);
                v.rxOf(p) = d; v.rxSec(p) = 0; v.err(p) = 0.0; v.ber(p) = 0.0; v.tReset(p) = now;
                if (v.mode != MODE_R) v.segT0(p) = now;
                berMask |= 1u << p;
            }
        } else if (ph == S_HDR) {
            // eHeaderCompletes: receivers decide on the header (simple_stack.py:241-251)
            const double hdrBits = (gen_hdr_bytes(v, B, d) * 8) * P.bitsFactor;
            uint32_t wake = 0;
            for (uint32_t m = (uint32_t)v.sc(GenView::I_rxMask); m; m &= m - 1u) {
                const int p = gen_ctz(m);
                if (v.rxOf(p) != d || v.rxSec(p) != 0) continue;
                gen_count(v, P, B, p);
                if (gen_decide(v, P, p, 0, hdrBits)) {
                    v.rxSec(p) = 1; v.err(p) = 0.0; v.ber(p) = 0.0; v.tReset(p) = v.now();  // _resetBitErrorCounter
                    if (v.mode != MODE_R) v.segT0(p) = v.now();
                    berMask |= 1u << p;
                } else {
                    gen_rx_clear(v, p);
                    if (v.sphase(p) == S_WAITRX) wake |= 1u << p;
                }
            }
            gen_set_sphase(v, d, S_PAY); v.tEv(d) = v.tC(d); v.sEv(d) = v.sC(d);
            for (int p = 0; p < nd; ++p) if ((wake >> p) & 1u) gen_begin_slot_wait(v, p);   // _nReceivingFinished.event
        } else {
            // eCompletes, callbacks in registration order:
            // 1. the sender's macInHandler resumes: _transmitting = False (simple_stack.py:210)
            const double payBits = (v.cmdPay(d) * 8) * P.bitsFactor;
            gen_set_sphase(v, d, S_IDLE);
            // 2. _onCompletingTransmission of every other PHY (simple_stack.py:146-157)
            for (int p = 0; p < nd; ++p) {
                if (p == d) continue;
                const bool own = (((uint32_t)v.sc(GenView::I_rxMask) >> p) & 1u) && v.rxOf(p) == d;
                if (gen_power_change(v, P, B, p, -v.rp(p, d), own)) berMask |= 1u << p;
            }
            // 3. receivers that passed the header count again (appendix B #4), decide on the payload and deliver
            int window = -1;
            uint32_t wake = 0, received = 0;
            for (uint32_t m = (uint32_t)v.sc(GenView::I_rxMask); m; m &= m - 1u) {
                const int p = gen_ctz(m);
                if (v.rxOf(p) != d || v.rxSec(p) != 1) continue;
                gen_count(v, P, B, p);
                if (gen_decide(v, P, p, 1, payBits)) {
                    if (p < ns) {
                        // SimpleMac.phyInHandler (blocking, not queued): an announcement addressed to an idle MAC
                        // opens its window (simple_stack.py:386-434); a data packet addressed to an idle MAC in
                        // receive mode goes to the network layer (:436-444)
                        const bool idle = v.mac(p) == MAC_NONE;
                        if (d == RRM) {
                            if (idle && v.sc(GenView::I_annDest) == p) window = p;
                        } else if (d < ns && idle && B.recv[p] && B.dest[d] == p) {
                            received |= 1u << p;
                        }
                    } else if (p == RRM) {
                        // SimpleRrmMac.phyInHandler -> interpreter.onPacketReceived (devices.py:163-168,
                        // counter_traffic.py:75-80): receivedValues[sender] = payload.value (= 2, appendix B #1)
                        if (d < ns) {
                            if (d == 0) v.sc(GenView::I_rv0) = kCounterByteLen;
                            if (d == 1) v.sc(GenView::I_rv1) = kCounterByteLen;
                            v.sc(GenView::I_latestDiff) = v.sc(GenView::I_rv0) - v.sc(GenView::I_rv1);
                            v.sc(GenView::I_rvMask) |= 1 << d;
                            v.nDeliv(d) += 1u;
                        }
                        gen_rec(v, REC_RX, v.now(), d, 0.0, 0.0, 0.0, 0.0);
                    }
                }
                gen_rx_clear(v, p);
                if (v.sphase(p) == S_WAITRX) wake |= 1u << p;
            }
            // zero-delay children in SimPy's pop order:
            // a. URGENT: phyInHandler of the grantee opens its window (simple_stack.py:399-406)
            if (window >= 0) {
                const double timeTotal = v.annSlots() * kSlot;
                v.stopW(window) = v.now() + timeTotal;
                v.sW(window) = v.seq()++;
                v.wPend(window) = 1; v.sc(GenView::I_wMask) |= 1 << window;
                v.wDone(window) = 0;
                gen_mac_loop_head(v, P, B, window);
            }
            // b. SEND eProcessed: the sender's upper layer resumes
            if (d < ns) {
                gen_mac_loop_head(v, P, B, d);                       // `yield message.eProcessed` returns
            } else if (d == RRM) {
                v.tRrm() = v.now() + (v.annSlots() + 1) * kSlot;    // simple_stack.py:558
                v.sc(GenView::I_sRrm) = (int32_t)(v.seq()++);
                v.sc(GenView::I_rrmPend) = 1;
            } else {
                // c. executeNext of the queued macIn executor: a PHY-only sender's pending SEND starts
                const int j = d - RRM - 1;
                if (v.jamPending(j) > 0) { v.jamPending(j) -= 1; gen_phy_send_init(v, d); }
            }
            // d. _nReceivingFinished.event of the receivers that finished
            for (int p = 0; p < nd; ++p) if ((wake >> p) & 1u) gen_begin_slot_wait(v, p);
            // e. RECEIVE.eProcessed: the device's receive loop hands the packet to onReceive and issues the next
            // RECEIVE command with a fresh timeout (devices.py:88-95, simple_stack.py:452-460)
            for (int k = 0; k < ns; ++k) {
                if (!((received >> k) & 1u)) continue;
                v.nRecv(k) += 1u;
                gen_rec(v, REC_MRX, v.now(), k, 0.0, 0.0, 0.0, 0.0);
                v.rxT(k) = v.now() + kReceiveTimeout; v.rxS(k) = v.seq()++;
            }
        }
        break;
    }
    case EV_RXTO:
        // the current RECEIVE command timed out (or the receive loop starts): the loop issues the next command
        // (simple_stack.py:473-478, devices.py:88-93)
        v.rxT(ev.idx) = v.now() + kReceiveTimeout; v.rxS(ev.idx) = v.seq()++;
        break;
    case EV_W: {
        // window timeoutEvent processed (simple_stack.py:406-420)
        const int k = ev.idx;
        v.wPend(k) = 0; v.sc(GenView::I_wMask) &= ~(1 << k);
        if (v.mac(k) == MAC_WAIT_TX) v.wDone(k) = 1;
        else gen_set_mac(v, k, MAC_NONE);
        break;
    }
    case EV_MOVE: {
        const int d = ev.idx;
        if (v.moveStage(d) == 0) {                                  // yield SimMan.timeout(first delay)
            v.moveStage(d) = 1; v.tMove(d) = v.now() + v.moveDelay(d); v.sMove(d) = v.seq()++;
        } else if (v.moveK(d) < B.maxMoves) {
            // d.position.set(initialPos.x + xOffset, initialPos.y + yOffset) with `initialPos` the moving Position
            // object itself: the offsets accumulate; then yield SimMan.timeout(MOVE_INTERVAL)
            const double *o = v.offsets + ((long long)d * B.maxMoves + v.moveK(d)) * 2;
            v.moveK(d) += 1;
            gen_move_device(v, P, B, d, v.posw(d, 0) + o[0], v.posw(d, 1) + o[1]);
            v.tMove(d) = v.now() + B.moveInterval; v.sMove(d) = v.seq()++;
        } else {
            v.moveStage(d) = 2;                                     // tape exhausted: the process is not modelled further
        }
        break;
    }
    case EV_RRM:
        // assignMessage.setProcessed() (simple_stack.py:561): the step ends here
        v.sc(GenView::I_rrmPend) = 0;
        v.sc(GenView::I_assignDone) = 1;
        break;
    default:
        v.sc(GenView::I_fault) = FAULT_EMPTY;
    }
    return berMask;
}

// Devices moving between steps: `want` [nd][2] are the requested positions, the devices are moved one after the
// other by ascending index like successive Position.set calls (gen_move_device)
GW_HD void gen_move_devices(GenView &v, const Params &P, const GenBand &B, const double *want)
{
    for (int m = 0; m < v.nd; ++m) gen_move_device(v, P, B, m, want[2 * m], want[2 * m + 1]);
}

// SimpleRrmDevice.assignFrequencyBand + SimpleRrmMac._sendAnnouncement start (devices.py:178-203,
// simple_stack.py:536-556): the RRM PHY receives a SEND command for the announcement
GW_HD void gen_begin_assignment(GenView &v, const Params &P, int device, int duration)
{
    const long long slots = (long long)duration * P.factor;         // counter_traffic.py:149
    int nbytes = 1;                                                  // len(str(slots)), messages.py:62-64
    for (long long lim = 10; lim <= slots && nbytes < 18; lim *= 10) ++nbytes;
    v.sc(GenView::I_annDest) = device;
    v.annSlots() = (double)slots;
    v.sc(GenView::I_assignDone) = 0;
    // the RRM PHY's queued macIn executor is idle here: its previous SEND completed before the previous
    // assignment's guard time-out (simple_stack.py:557-558)
    if (v.sphase(v.ns) != S_IDLE) v.sc(GenView::I_fault) = FAULT_SENDQ;
    v.cmdPay(v.ns) = nbytes;
    gen_phy_send_init(v, v.ns);
}

// Interpreter.getFeedback (envs/core.py:142-153, counter_traffic.py:85-107)
GW_HD void gen_feedback(GenView &v, long long &obs, double &reward, unsigned char &done)
{
    const int diff = v.sc(GenView::I_latestDiff);
    obs = (long long)diff + kCounterBound;
    const int absd = diff < 0 ? -diff : diff;
    int r = v.sc(GenView::I_lastAbsDiff) - absd;
    v.sc(GenView::I_lastAbsDiff) = absd;
    if (r > 10) r = 10; else if (r < -10) r = -10;
    reward = (double)r;
    done = (unsigned char)v.sc(GenView::I_done);
}

// env.step(action) in three pieces (counter_traffic.py:146-158): assign; run until the ASSIGN message is processed --
// one timed event per call, each followed by the BER evaluations it asks for --; the interpreter's feedback.
// An action outside the action space (the reference asserts) leaves the band-sim untouched and reports a fault.
GW_HD bool gen_step_begin(GenView &v, const Params &P, const GenBand &B, int device, int duration)
{
    if (v.sc(GenView::I_fault) != 0) return false;
    if (device < 0 || device >= v.ns || duration < 0 || duration >= B.maxDuration) {
        v.sc(GenView::I_fault) = FAULT_EMPTY + 1;                  // FAULT_ACTION of the C ABI
        return false;
    }
    gen_begin_assignment(v, P, device, duration);
    return v.sc(GenView::I_fault) == 0;
}

GW_HD bool gen_step_running(const GenView &v) { return !v.sc(GenView::I_assignDone) && !v.sc(GenView::I_fault); }

GW_HD uint32_t gen_step_event(GenView &v, const Params &P, const GenBand &B)
{
    const Event ev = gen_next_event(v, B);
    return gen_apply(v, P, B, ev);
}

// Mode M in the kernel: the PHYs that run SimplePhy._countBitErrors at event `ev` (the set of gw_core.cuh::count_set,
// which depends only on the state BEFORE the event): every receiving PHY whose power sum changes when a transmission
// starts or ends (a zero change counts nothing, simple_stack.py:224), the receivers of the header that ends, the
// receivers of the payload that ends (their second count finds an empty range).  The warp counts these ranges
// together, then gen_apply runs and finds them counted (segT0 == now).
GW_HD uint32_t gen_count_set(const GenView &v, const Event &ev)
{
    if (ev.kind != EV_PHY) return 0;
    const int d = ev.idx, ph = v.sphase(d);
    uint32_t set = 0;
    for (uint32_t m = (uint32_t)v.sc(GenView::I_rxMask); m; m &= m - 1u) {
        const int p = gen_ctz(m);
        const int rx = v.rxOf(p);
        if (ph == S_HDR) {
            if (rx == d && v.rxSec(p) == 0) set |= 1u << p;
        } else {
            if (p != d && v.rp(p, d) != 0.0) set |= 1u << p;
            if (ph == S_PAY && rx == d && v.rxSec(p) == 1) set |= 1u << p;
        }
    }
    return set;
}

GW_HD void gen_step_end(GenView &v, long long &obs, double &reward, unsigned char &done)
{
    if (v.sc(GenView::I_fault)) {
        // a rejected action or a faulted band-sim: nothing happened, the interpreter is not consulted
        obs = (long long)v.sc(GenView::I_latestDiff) + kCounterBound; reward = 0.0; done = (unsigned char)v.sc(GenView::I_done);
        return;
    }
    gen_feedback(v, obs, reward, done);
}

// the serial driver (host build, traced kernel)
GW_HD void gen_step(GenView &v, const Params &P, const GenBand &B, int device, int duration, long long &obs,
                    double &reward, unsigned char &done)
{
    if (gen_step_begin(v, P, B, device, duration))
        while (gen_step_running(v)) gen_update_bers(v, P, gen_step_event(v, P, B));
    gen_step_end(v, obs, reward, done);
}

}  // namespace gw
