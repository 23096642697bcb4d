// gymwipe_b200 -- grids of PHY-only senders with in-step mobility.
//
// The reference's own benchmark scenario (tests/test_benchmark.py:20-91): n devices, each a SimplePhy driven by
// a sender process (initial delay, then one SEND of a 13 + 26 byte packet at 40 dBm every SEND_INTERVAL), every
// PHY receiving what the others send; optionally one mobility process per device that moves it every
// MOVE_INTERVAL while transmissions are on the air (SimplePhy._onAttenuationChange, simple_stack.py:119-128);
// the simulation is advanced by SimMan.runSimulation(duration) -- no RRM, no MAC, no gym step.
//
// Same method as the step kernel's core (gw_core.cuh): one band-sim per thread, the SimPy heap replaced by
// timed event slots ordered by (time, creation number) -- per device one sender wake-up, one PHY event (slot
// start -> header end -> completion) and one mobility wake-up -- with the zero-delay event chains executed
// inline in SimPy's pop order.  Unlike the step kernel the number of devices is a RUN-TIME value (up to
// kGridMaxDev), so every per-device array lives in the band-sim's state block in global memory and is indexed
// directly; nothing is unrolled.  This is the general engine, not the tuned one: the 2-sender + RRM template of
// CounterTrafficEnv keeps its own kernel.
//
// Reference semantics (file:line under /root/reference):
//   sender process / macIn queue      tests/test_benchmark.py:31-48, networking/construction.py:290, simtools.py:347-381
//   SimplePhy (send, receive, power)  networking/simple_stack.py:77-286
//   Transmission                      networking/physical.py:224-290, 576-608
//   FSPL model, lazy creation, moves  networking/attenuation_models.py:28-36, physical.py:354-397, 500-528
//   mobility process                  tests/test_benchmark.py:73-85 (`initialPos` is the moving Position object:
//                                     the offsets accumulate)
//   runSimulation(duration)           simtools.py:77-88 (simpy run(until = now + duration))
//
// Plain C++ (no CUDA intrinsics): gw_kernels.cu includes it for the device, tests/hostsim for the host.
#pragma once

#include "gw_core.cuh"

namespace gw {

constexpr int kGridMaxDev = 24;

struct GridParams {
    int ndev;
    int maxMoves;                       // jumps per device in the offset tape (0: no mobility processes)
    double bitRate, dataRate, maxBer, tenLog10BitRate, qDen, bitsFactor;
    double frequency, thermal;
    double fsplConst;                   // 20 * log10(frequency), evaluated once on the host (a move evaluates the FSPL towards
                                        // every partner: one log10 less each)
    double moveInterval;
    double power[kGridMaxDev];          // dBm
    double interval[kGridMaxDev];       // send interval
    int hdrBytes[kGridMaxDev], payBytes[kGridMaxDev];
};

struct GridDev {
    // sender process (tests/test_benchmark.py:31-48) and the PHY's queued macIn executor
    double tJam, delay;
    uint32_t sJam;
    int jamStage, jamPending;
    // transmit side of the PHY (simple_stack.py:192-212): S_* of gw_core.cuh
    int sphase;
    uint32_t sEv, sC, txSeq;
    double tEv, tC, txStart, tStop;
    // receive side (simple_stack.py:214-267)
    int rxOf, rxSec;
    double P, ber, err, tReset;
    // mobility process (tests/test_benchmark.py:73-85)
    double tMove, moveDelay, x, y;
    uint32_t sMove;
    int moveStage, moveK;
    // statistics
    uint32_t nTx, nHdrOk, nHdrFail, nPayOk, nPayFail, nBer;
};

struct GridHead {
    double now;
    uint32_t seq, ties;
    int fault, pad;
};

// view of one band-sim's state block: GridHead | GridDev[n] | att[n*n] | srx[n*n]
struct GridView {
    GridHead *h;
    GridDev *dev;
    double *att, *srx;      // entry (receiver p, sender d) at p * n + d
    int n;
    // optional event trace (records of 8 doubles, as gw_core.cuh::trace_rec)
    double *trace;
    int ntrace, traceCap;
};

GW_HD size_t grid_state_bytes(int n)
{
    return sizeof(GridHead) + sizeof(GridDev) * (size_t)n + 2 * sizeof(double) * (size_t)n * (size_t)n;
}

GW_HD GridView grid_view(void *block, int n)
{
    GridView v;
    char *p = (char *)block;
    v.h = (GridHead *)p; p += sizeof(GridHead);
    v.dev = (GridDev *)p; p += sizeof(GridDev) * (size_t)n;
    v.att = (double *)p; p += sizeof(double) * (size_t)n * (size_t)n;
    v.srx = (double *)p;
    v.n = n;
    v.trace = nullptr; v.ntrace = 0; v.traceCap = 0;
    return v;
}

GW_HD void grid_rec(GridView &v, int kind, double t, int dev, double x0, double x1, double x2, double x3)
{
    if (v.trace == nullptr) return;
    if (v.ntrace < v.traceCap) {
        double *r = v.trace + (long long)v.ntrace * 8;
        r[0] = kind; r[1] = t; r[2] = dev; r[3] = x0; r[4] = x1; r[5] = x2; r[6] = x3; r[7] = 0;
    }
    v.ntrace += 1;
}

// construction (device_grid fixture, tests/test_benchmark.py:52-71; mobile_device_grid :73-85): devices in index
// order, each constructor starts its sender process; then, with mobility, one mover process per device.
// pos [n][2], delays [n], moveDelays [n] (or null)
GW_HD void grid_init(GridView &v, const GridParams &G, const double *pos, const double *delays, const double *moveDelays)
{
    const int n = v.n;
    v.h->now = 0.0; v.h->seq = 0; v.h->ties = 0; v.h->fault = 0; v.h->pad = 0;
    for (int d = 0; d < n; ++d) {
        GridDev &D = v.dev[d];
        D.tJam = 0.0; D.delay = delays[d]; D.sJam = v.h->seq++; D.jamStage = 0; D.jamPending = 0;
        D.sphase = S_IDLE; D.sEv = 0; D.sC = 0; D.txSeq = 0; D.tEv = 0; D.tC = 0; D.txStart = 0; D.tStop = 0;
        D.rxOf = -1; D.rxSec = 0; D.P = G.thermal; D.ber = 0; D.err = 0; D.tReset = 0;
        D.tMove = 0.0; D.moveDelay = moveDelays ? moveDelays[d] : 0.0; D.x = pos[2 * d]; D.y = pos[2 * d + 1];
        D.sMove = 0; D.moveStage = G.maxMoves > 0 ? 0 : 2; D.moveK = 0;
        D.nTx = D.nHdrOk = D.nHdrFail = D.nPayOk = D.nPayFail = D.nBer = 0;
    }
    for (int d = 0; d < n; ++d)
        if (G.maxMoves > 0) v.dev[d].sMove = v.h->seq++;
    for (int p = 0; p < n; ++p)
        for (int d = 0; d < n; ++d) {
            double att = 0.0, rp = 0.0;
            if (p != d) {
                att = fspl_db(pos[2 * p], pos[2 * p + 1], pos[2 * d], pos[2 * d + 1], G.frequency);
                rp = rx_power_mw(G.power[d], att);
            }
            v.att[p * n + d] = att;
            v.srx[p * n + d] = rp;
        }
}

// SimplePhy._updateBitErrorRate (simple_stack.py:161-173)
GW_HD void grid_update_ber(GridView &v, const GridParams &G, int p)
{
    GridDev &R = v.dev[p];
    const int e = R.rxOf;
    if (e < 0) return;
    const double S = v.srx[p * v.n + e];
    const double N = R.P - S;
    if (!(S >= 0) || !(N >= 0)) { v.h->fault = FAULT_REF_ASSERT; return; }
    R.ber = ber_bpsk_mw_cold(S, N, G.tenLog10BitRate, G.qDen);
    R.nBer += 1;
    grid_rec(v, REC_BER, v.h->now, p, R.ber, 0.0, 0.0, 0.0);
}

// SimplePhy._countBitErrors (simple_stack.py:180-188): duration from the last RESET (appendix B #5)
GW_HD void grid_count(GridView &v, const GridParams &G, int p)
{
    GridDev &R = v.dev[p];
    const double duration = v.h->now - R.tReset;
    const double bitErrors = R.ber * duration * G.bitRate;
    R.err += bitErrors;
}

// _nReceivedPowerChanges.trigger(delta): power sum, then the running reception (simple_stack.py:81-86, 223-233)
// `defer`: if given, the PHY is only marked for a BER evaluation after the event (the evaluations of a PHY event
// read nothing the rest of the event writes; the kernel evaluates the marked PHYs of all lanes of a warp together)
GW_HD void grid_power_change(GridView &v, const GridParams &G, int p, double delta, bool completing_own, uint32_t *defer = nullptr)
{
    GridDev &R = v.dev[p];
    R.P += delta;
    if (R.rxOf < 0 || delta == 0.0) return;
    grid_count(v, G, p);
    const bool completed = v.h->now >= v.dev[R.rxOf].tStop;
    if (!completed) {
        // `if not t.completed: _updateBitErrorRate(t)` with the power entry of its own transmission already
        // popped: the reference raises KeyError (appendix B #12)
        if (completing_own) { v.h->fault = FAULT_REF_KEYERROR; return; }
        if (defer) *defer |= 1u << p; else grid_update_ber(v, G, p);
    }
}

GW_HD void grid_update_bers(GridView &v, const GridParams &G, uint32_t mask)
{
    for (int p = 0; p < v.n; ++p)
        if ((mask >> p) & 1u) grid_update_ber(v, G, p);
}

GW_HD void grid_rx_clear(GridView &v, int p)
{
    GridDev &R = v.dev[p];
    R.rxOf = -1; R.err = 0.0; R.ber = 0.0; R.tReset = v.h->now;
}

GW_HD void grid_begin_slot_wait(GridView &v, int d)
{
    GridDev &D = v.dev[d];
    D.tEv = v.h->now + (kSlot - fmod_slot(v.h->now));          // simtools.py:53
    D.sEv = v.h->seq++;
    D.sphase = S_SLOT;
}

GW_HD void grid_phy_send_init(GridView &v, int d)
{
    if (v.dev[d].rxOf >= 0) v.dev[d].sphase = S_WAITRX;         // yield self._nReceivingFinished.event
    else grid_begin_slot_wait(v, d);
}

GW_HD bool grid_decide(GridView &v, const GridParams &G, int p, int section, double totalBits)
{
    const double x = rint(v.dev[p].err);                        // round(): half-even
    const bool ok = x / totalBits <= G.maxBer;
    grid_rec(v, REC_DEC, v.h->now, p, section, v.dev[p].err, totalBits, ok ? 1.0 : 0.0);
    return ok;
}

// Position.set of device m (devices/core.py:75-84) and everything it triggers -- the single-device form of
// gw_core.cuh::move_devices, see the comments there
GW_HD void grid_move(GridView &v, const GridParams &G, int m, double x, double y)
{
    const int n = v.n;
    GridDev &M = v.dev[m];
    if (x == M.x && y == M.y) return;
    M.x = x; M.y = y;
    for (int j = 0; j < n; ++j) {
        if (j == m) continue;
        const double dx = M.x - v.dev[j].x, dy = M.y - v.dev[j].y;
        const double dist = sqrt(dx * dx + dy * dy);
        if (M.txSeq == 0u && v.dev[j].txSeq == 0u) {
            // no model yet (created at the first transmission of either device): the table follows the positions
            const double fresh = (dx == 0.0 && dy == 0.0) ? 0.0 : 20 * log10(dist) + G.fsplConst - 147.55;
            v.att[m * n + j] = fresh; v.att[j * n + m] = fresh;     // (received powers: evaluated at the next transmission start)
            continue;
        }
        if (!(dist < 3000.0)) continue;                         // STANDBY_THRESHOLD, physical.py:371
        if (dx == 0.0 && dy == 0.0) continue;                   // FsplAttenuation._update returns early
        const double att = 20 * log10(dist) + G.fsplConst - 147.55;
        if (att == v.att[m * n + j]) continue;                  // only a NEW value triggers
        v.att[m * n + j] = att; v.att[j * n + m] = att;
        // SimplePhy._onAttenuationChange (simple_stack.py:119-128) exists only for transmissions that are on the air; the
        // received power of a later transmission is evaluated from the attenuation of that moment at its start
        // (simple_stack.py:130-144) -- so a move evaluates dbmToMilliwatts only for the links whose sender is sending
        // (2 pow() per partner and move otherwise: with the FSPL's log10 most of what a mobile grid costs)
        for (int dir = 0; dir < 2; ++dir) {
            const int p = dir == 0 ? j : m, e = dir == 0 ? m : j;
            const int ph = v.dev[e].sphase;
            if (ph != S_HDR && ph != S_PAY) continue;
            const double rp = rx_power_mw(G.power[e], att);
            const double delta = rp - v.srx[p * n + e];
            v.srx[p * n + e] = rp;
            grid_power_change(v, G, p, delta, false);
        }
    }
}

// one timed event; `offsets` [n][maxMoves][2] position offsets of the mobility processes
// `defer`: see grid_power_change (PHY events only: a moving device's PHY can see several power changes in one event,
// each counted with the rate the previous one left)
GW_HD void grid_apply(GridView &v, const GridParams &G, int kind, int d, const double *offsets, uint32_t *defer = nullptr)
{
    const int n = v.n;
    GridDev &D = v.dev[d];
    if (kind == EV_JAM) {
        if (D.jamStage == 0) {                                  // yield SimMan.timeout(initialDelay)
            D.jamStage = 1; D.tJam = v.h->now + D.delay; D.sJam = v.h->seq++;
        } else if (D.jamStage == 1) {                           // first yield SimMan.timeout(sendInterval)
            D.jamStage = 2; D.tJam = v.h->now + G.interval[d]; D.sJam = v.h->seq++;
        } else {
            // macIn.send(SEND) -> queued executor; then the next timeout
            D.tJam = v.h->now + G.interval[d]; D.sJam = v.h->seq++;
            if (D.sphase != S_IDLE) { D.jamPending += 1; if (D.jamPending > 60) v.h->fault = FAULT_SENDQ; }
            else grid_phy_send_init(v, d);
        }
        return;
    }
    if (kind == EV_W) {                                         // mobility wake-up (slot kind reused)
        if (D.moveStage == 0) {                                 // yield SimMan.timeout(random.uniform(0, MOVE_INTERVAL))
            D.moveStage = 1; D.tMove = v.h->now + D.moveDelay; D.sMove = v.h->seq++;
        } else if (D.moveK < G.maxMoves) {
            const double *o = offsets + ((long long)d * G.maxMoves + D.moveK) * 2;
            D.moveK += 1;
            grid_move(v, G, d, D.x + o[0], D.y + o[1]);         // offsets accumulate (the reference's `initialPos`)
            D.tMove = v.h->now + G.moveInterval; D.sMove = v.h->seq++;
        } else {
            D.moveStage = 2;                                    // tape exhausted: the process is not modelled further
        }
        return;
    }
    // EV_PHY
    if (D.sphase == S_SLOT) {
        // FrequencyBand.transmit -> Transmission.__init__ (physical.py:224-279, 596-608)
        const double hd = (G.hdrBytes[d] * 8) / G.dataRate;
        const double pd = (G.payBytes[d] * 8) / G.dataRate;
        const double duration = hd + pd;
        const double now = v.h->now;
        const double stop = now + duration;
        const double headerStop = now + hd;
        const double tH = now + (headerStop > now ? headerStop - now : 0.0);   // timeoutUntil
        const double tC = now + (stop > now ? stop - now : 0.0);
        const uint32_t qH = v.h->seq++, qC = v.h->seq++;
        D.sphase = S_HDR; D.tEv = tH; D.sEv = qH; D.tC = tC; D.sC = qC; D.txStart = now; D.tStop = stop;
        D.txSeq += 1; D.nTx += 1;
        grid_rec(v, REC_TX, now, d, stop, (G.hdrBytes[d] * 8) * G.bitsFactor, (G.payBytes[d] * 8) * G.bitsFactor, 0.0);
        // zero-delay notification: every other PHY registers the received power (simple_stack.py:130-144); with
        // mobility processes it is evaluated here, from the attenuation of this moment (see grid_move)
        for (int p = 0; p < n; ++p) {
            if (p == d) continue;
            if (G.maxMoves > 0) v.srx[p * n + d] = rx_power_mw(G.power[d], v.att[p * n + d]);
            grid_power_change(v, G, p, v.srx[p * n + d], false, defer);
            if (v.h->fault) return;
        }
        // receive processes in construction order: idle, non-transmitting PHYs lock on (simple_stack.py:214-235)
        for (int p = 0; p < n; ++p) {
            GridDev &R = v.dev[p];
            if (p == d || R.rxOf >= 0 || R.sphase >= S_SLOT) continue;
            R.rxOf = d; R.rxSec = 0; R.err = 0.0; R.ber = 0.0; R.tReset = now;
            if (defer) *defer |= 1u << p; else grid_update_ber(v, G, p);
        }
    } else if (D.sphase == S_HDR) {
        // eHeaderCompletes (simple_stack.py:241-251)
        const double hdrBits = (G.hdrBytes[d] * 8) * G.bitsFactor;
        for (int p = 0; p < n; ++p) {
            GridDev &R = v.dev[p];
            if (R.rxOf != d || R.rxSec != 0) continue;
            grid_count(v, G, p);
            if (grid_decide(v, G, p, 0, hdrBits)) {
                R.nHdrOk += 1;
                R.rxSec = 1; R.err = 0.0; R.ber = 0.0; R.tReset = v.h->now;
                if (defer) *defer |= 1u << p; else grid_update_ber(v, G, p);
            } else {
                R.nHdrFail += 1;
                grid_rx_clear(v, p);
                R.rxSec = 2;                                    // marks "finished at this event" for the wake-up pass
            }
            if (v.h->fault) return;
        }
        D.sphase = S_PAY; D.tEv = D.tC; D.sEv = D.sC;
        for (int p = 0; p < n; ++p) {                           // _nReceivingFinished.event of the PHYs that gave up
            GridDev &R = v.dev[p];
            if (R.rxOf < 0 && R.rxSec == 2) { R.rxSec = 0; if (R.sphase == S_WAITRX) grid_begin_slot_wait(v, p); }
        }
    } else {
        // eCompletes, callbacks in registration order:
        // 1. the sender's macInHandler resumes (simple_stack.py:210)
        const double payBits = (G.payBytes[d] * 8) * G.bitsFactor;
        D.sphase = S_IDLE;
        // 2. _onCompletingTransmission of every other PHY (simple_stack.py:146-157)
        for (int p = 0; p < n; ++p) {
            if (p == d) continue;
            grid_power_change(v, G, p, -v.srx[p * n + d], v.dev[p].rxOf == d, defer);
            if (v.h->fault) return;
        }
        // 3. receivers that passed the header decide on the payload (second count: appendix B #4)
        for (int p = 0; p < n; ++p) {
            GridDev &R = v.dev[p];
            if (R.rxOf != d || R.rxSec != 1) continue;
            grid_count(v, G, p);
            if (grid_decide(v, G, p, 1, payBits)) R.nPayOk += 1; else R.nPayFail += 1;
            grid_rx_clear(v, p);
            R.rxSec = 2;
        }
        // zero-delay children in SimPy's pop order: the queued macIn executor starts a pending SEND, then the
        // PHYs that waited for the end of their reception start their slot wait
        if (D.jamPending > 0) { D.jamPending -= 1; grid_phy_send_init(v, d); }
        for (int p = 0; p < n; ++p) {
            GridDev &R = v.dev[p];
            if (R.rxOf < 0 && R.rxSec == 2) { R.rxSec = 0; if (R.sphase == S_WAITRX) grid_begin_slot_wait(v, p); }
        }
    }
}

// SimMan.runSimulation(duration): every event strictly before now + duration, then the clock is set
// one timed event strictly before T; returns false when there is none (or the env has faulted).  `berMask`: the PHYs
// whose bit error rate the caller evaluates afterwards (grid_update_bers)
GW_HD bool grid_run_event(GridView &v, const GridParams &G, double T, const double *offsets, uint32_t &berMask)
{
    const int n = v.n;
    berMask = 0;
    if (v.h->fault) return false;
    int kind = EV_NONE, idx = 0;
    double t = INFINITY;
    uint32_t q = 0;
    for (int d = 0; d < n; ++d) {
        const GridDev &D = v.dev[d];
        if (kind == EV_NONE || before(D.tJam, D.sJam, t, q)) { kind = EV_JAM; idx = d; t = D.tJam; q = D.sJam; }
        if (D.sphase >= S_SLOT && before(D.tEv, D.sEv, t, q)) { kind = EV_PHY; idx = d; t = D.tEv; q = D.sEv; }
        if (D.moveStage < 2 && before(D.tMove, D.sMove, t, q)) { kind = EV_W; idx = d; t = D.tMove; q = D.sMove; }
    }
    if (kind == EV_NONE || !(t < T)) return false;
    v.h->now = t;
    grid_apply(v, G, kind, idx, offsets, kind == EV_PHY ? &berMask : nullptr);
    return true;
}

// SimMan.runSimulation(duration): every event strictly before now + duration, then the clock is set (serial driver:
// host build, traced kernel)
GW_HD void grid_run(GridView &v, const GridParams &G, double duration, const double *offsets)
{
    const double T = v.h->now + duration;
    uint32_t m;
    while (grid_run_event(v, G, T, offsets, m)) grid_update_bers(v, G, m);
    v.h->now = T;
}

}  // namespace gw
