// TEST INFRASTRUCTURE -- host build of the device simulation core (gymwipe_b200/csrc/gw_core.cuh).
//
// The product runs this code only inside CUDA kernels; here the very same header is compiled
// with g++ so that the event logic can be compared with the oracle in the `-m "not gpu"` test
// suite (no GPU in the build container).  Never loaded by the gymwipe_b200 package.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

static long long g_macro_stat[4] = {0, 0, 0, 0};
#define GW_STAT_MACRO(i) (++g_macro_stat[i])
#include "../../gymwipe_b200/csrc/gw_core.cuh"
#include "../../gymwipe_b200/csrc/gw_pendulum.cuh"
#include "../../gymwipe_b200/csrc/gw_grid.cuh"
#include "../../gymwipe_b200/csrc/gw_band.cuh"

using namespace gw;

namespace {

struct HostRing {
    static constexpr bool ext = true;       // receive mode / bursts compiled in
    int32_t *data;      // [NS][100]
    // MAC receive mode: {time-out, creation number, packets handed to onReceive} per sender
    double *rxt; uint32_t *rxs; uint32_t *rxn;
    double rxT(int k) const { return rxt[k]; }
    uint32_t rxS(int k) const { return rxs[k]; }
    void set_rx(int k, double t, uint32_t q) const { rxt[k] = t; rxs[k] = q; }
    void add_received(int k) const { rxn[k] += 1u; }
    int operator()(int k, uint32_t slot) const { return data[k * kRingSlots + slot]; }
    void operator()(int k, uint32_t slot, int v) { data[k * kRingSlots + slot] = v; }
    // counter epochs (the device reads them from global memory on demand)
    template <class S> uint64_t epochK(const S &s, int k) const { return get_at(s.epochK, k); }
    template <class S> int epochC(const S &s, int k) const { return get_at(s.epochC, k); }
};

struct HostMasks {
    uint64_t seed; int64_t env; int band;
    int64_t operator()(int receiver, int sender, uint32_t txseq, int64_t k0, int64_t k1, double ber) const
    {
        return mask_errors_serial(seed, env, band, sender, txseq, receiver, k0, k1, ber);
    }
};

// mode M with FED masks: words laid out [nenv][nbands][4 sender][slots][4 receiver][words_per_row]
struct HostFedMasks {
    const uint32_t *words; int slots, wpr; int64_t env; int band, nb;
    int64_t operator()(int receiver, int sender, uint32_t txseq, int64_t k0, int64_t k1, double) const
    {
        const int64_t row = ((((env * nb + band) * kMaxDev + sender) * slots + (int64_t)(txseq % (uint32_t)slots)) * kMaxDev + receiver);
        const uint32_t *w = words + row * wpr;
        int64_t n = 0;
        for (int64_t k = k0; k < k1; ++k) n += (w[k >> 5] >> (k & 31)) & 1u;
        return n;
    }
};
static const uint32_t *g_fed_words = nullptr;
static int g_fed_slots = 0, g_fed_wpr = 0;

template <int MODE>
struct MasksFor {
    using type = HostMasks;
    static HostMasks make(uint64_t seed, int64_t env, int64_t, int band, int) { return HostMasks{seed, env, band}; }
};
template <>
struct MasksFor<MODE_M_FED> {
    using type = HostFedMasks;
    static HostFedMasks make(uint64_t, int64_t, int64_t local_env, int band, int nb)
    {
        return HostFedMasks{g_fed_words, g_fed_slots, g_fed_wpr, local_env, band, nb};
    }
};

struct HsBand {
    int32_t ns, nj;
    double frequency, bandwidth;
    double x[kMaxDev], y[kMaxDev], power[kMaxDev];
    int32_t mult[kMaxSend], payloadRule[kMaxSend];
    double interval[kMaxSend];
    int32_t maxTicks[kMaxSend], recv[kMaxSend];
    double jamInterval, jamDelay;
    int32_t jamHdr, jamPay;
};

struct HsMove {              // the device `dev` of band `band` jumps to (x, y) before step `step` (all envs)
    int32_t step, band, dev;
    double x, y;
};

template <int D>
struct HostTab {             // tables of one band-sim, [D * D], entry (receiver p, sender d) at p * D + d
    double *a, *r;
    double att(int p, int d) const { return a[p * D + d]; }
    void set_att(int p, int d, double v) { a[p * D + d] = v; }
    double srx(int p, int d) const { return r[p * D + d]; }
    void set_srx(int p, int d, double v) { r[p * D + d] = v; }
    const double *view() const { return r; }
};

struct HsScenario {
    int32_t nbands, factor, mode;
    uint64_t seed;
    HsBand band[kMaxBands];
};

int g_no_macro = -1;    // hs_set_no_macro: -1 macro events wherever they apply (default here), 1 never, 0 as the kernels do

void fill_params(const HsScenario &sc, Params &P)
{
    std::memset(&P, 0, sizeof P);
    P.noMacro = g_no_macro;
    P.nbands = sc.nbands; P.factor = sc.factor; P.maxDuration = 20; P.mode = sc.mode;
    P.bitRate = 133.33333e3;
    P.dataRate = 0.75 * P.bitRate;
    P.maxBer = 0.25;
    P.tenLog10BitRate = 10 * std::log10(P.bitRate);
    P.qDen = 1.135 * std::sqrt(2 * 3.141592653589793);
    P.bitsFactor = 1.25;
    finish_params(P);
    for (int b = 0; b < sc.nbands; ++b) {
        BandParams &B = P.band[b];
        const HsBand &h = sc.band[b];
        B.ns = h.ns; B.nj = h.nj; B.ndev = h.ns + 1 + h.nj;
        for (int k = 0; k < kMaxSend; ++k) {
            B.mult[k] = h.mult[k]; B.payloadRule[k] = h.payloadRule[k]; B.interval[k] = h.interval[k];
            B.maxTicks[k] = h.maxTicks[k]; B.recv[k] = h.recv[k];
            if (h.maxTicks[k] != 0 || h.recv[k]) P.noMacro = 1;       // as gw_kernels.cu::fill_params
        }
        B.jamInterval[0] = h.jamInterval; B.jamDelay[0] = h.jamDelay; B.jamHdr[0] = h.jamHdr; B.jamPay[0] = h.jamPay;
    }
}

template <int D, int NS, int NJ>
struct EnvT {
    Sim<D, NS, NJ> sim[kMaxBands];
    double srx[kMaxBands][D * D];
    double att[kMaxBands][D * D];
    double pos[kMaxBands][D * 2];
    int32_t ring[kMaxBands][NS * kRingSlots];
    double rxt[kMaxBands][NS];
    uint32_t rxs[kMaxBands][NS], rxn[kMaxBands][NS];
};

template <int MODE, int D, int NS, int NJ>
int run_t(const HsScenario &sc, int64_t nenv, int nsteps, int do_reset, const double *pos,
          const int32_t *dev_tape, const int32_t *dur_tape, int64_t *obs, double *reward, uint8_t *done,
          double *now, int64_t *counts, double *power_out, int64_t env_offset,
          const HsMove *moves = nullptr, int nmoves = 0)
{
    Params P;
    fill_params(sc, P);
    const int nb = sc.nbands;
    std::vector<EnvT<D, NS, NJ>> envs(1);
    int fault_any = 0;
    for (int64_t e = 0; e < nenv; ++e) {
        EnvT<D, NS, NJ> &E = envs[0];
        std::memset(&E, 0, sizeof E);
        for (int b = 0; b < nb; ++b) {
            const HsBand &h = sc.band[b];
            const double npd = 1.38e-23 * (20.0 + 273.15);
            const double thermal = npd * h.bandwidth * 1000;
            init_sim(E.sim[b], thermal);
            { HostRing r0{E.ring[b], E.rxt[b], E.rxs[b], E.rxn[b]}; init_receive(E.sim[b], P.band[b], r0); }
            double x[D], y[D];
            for (int d = 0; d < D; ++d) {
                if (pos) { x[d] = pos[((e * nb + b) * kMaxDev + d) * 2]; y[d] = pos[((e * nb + b) * kMaxDev + d) * 2 + 1]; }
                else { x[d] = h.x[d]; y[d] = h.y[d]; }
            }
            for (int d = 0; d < D; ++d) { E.pos[b][2 * d] = x[d]; E.pos[b][2 * d + 1] = y[d]; }
            for (int p = 0; p < D; ++p)
                for (int d = 0; d < D; ++d) {
                    E.att[b][p * D + d] = (p == d) ? 0.0 : fspl_db(x[p], y[p], x[d], y[d], h.frequency);
                    E.srx[b][p * D + d] = (p == d) ? 0.0 : rx_power_mw(h.power[d], E.att[b][p * D + d]);
                }
        }
        if (do_reset)
            for (int b = 0; b < nb; ++b) { HostRing r{E.ring[b], E.rxt[b], E.rxs[b], E.rxn[b]}; reset_sim(E.sim[b], P.band[b], r); }
        for (int t = 0; t < nsteps; ++t) {
            const size_t base = ((size_t)t * nenv + e) * nb;
            double T = 0;
            // devices that move before this step (Position.set between two env.step calls)
            for (int b = 0; b < nb && nmoves > 0; ++b) {
                double want[D * 2];
                bool any = false;
                for (int k = 0; k < D * 2; ++k) want[k] = E.pos[b][k];
                for (int k = 0; k < nmoves; ++k)
                    if (moves[k].step == t && moves[k].band == b && moves[k].dev < D) {
                        want[2 * moves[k].dev] = moves[k].x; want[2 * moves[k].dev + 1] = moves[k].y; any = true;
                    }
                if (!any) continue;
                HostTab<D> tab{E.att[b], E.srx[b]};
                auto mk = MasksFor<MODE>::make(sc.seed, env_offset + e, e, b, nb);
                double power[D];
                for (int d = 0; d < D; ++d) power[d] = sc.band[b].power[d];
                move_devices<MODE>(E.sim[b], P, power, sc.band[b].frequency, E.pos[b], want, tab, mk, NoMemo());
            }
            for (int b = 0; b < nb; ++b) {
                begin_assignment(E.sim[b], P, dev_tape[base + b], dur_tape[base + b]);
            }
            for (int b = 0; b < nb; ++b) {
                HostRing r{E.ring[b], E.rxt[b], E.rxs[b], E.rxn[b]};
                auto mk = MasksFor<MODE>::make(sc.seed, env_offset + e, e, b, nb);
                run_until_assign<MODE>(E.sim[b], P, P.band[b], E.srx[b], r, mk);
                if (E.sim[b].now > T) T = E.sim[b].now;
            }
            for (int b = 0; b < nb; ++b) {
                HostRing r{E.ring[b], E.rxt[b], E.rxs[b], E.rxn[b]};
                auto mk = MasksFor<MODE>::make(sc.seed, env_offset + e, e, b, nb);
                if (E.sim[b].now < T) run_until_time<MODE>(E.sim[b], P, P.band[b], E.srx[b], r, mk, T);
                long long o; double rw; unsigned char dn;
                feedback(E.sim[b], o, rw, dn);
                if (obs) obs[base + b] = o;
                if (reward) reward[base + b] = rw;
                if (done) done[base + b] = dn;
                if (E.sim[b].fault) fault_any = E.sim[b].fault;
            }
            if (now) now[(size_t)t * nenv + e] = T;
            if (fault_any) return fault_any;
        }
        for (int b = 0; b < nb; ++b) {
            if (counts) {
                int64_t *c = counts + ((size_t)e * nb + b) * 9;
                c[0] = E.sim[b].nTx;
                for (int k = 0; k < NS; ++k) c[1 + k] = E.sim[b].nDeliv[k];
                for (int k = 0; k < NS; ++k) c[3 + k] = E.rxn[b][k];      // packets handed to onReceive
                c[8] = E.sim[b].ties;
            }
            if (power_out)
                for (int d = 0; d < D; ++d) power_out[((size_t)e * nb + b) * kMaxDev + d] = E.sim[b].P[d];
        }
    }
    return 0;
}

}  // namespace

extern "C" {

int hs_run_moves(const HsScenario *sc, int64_t nenv, int nsteps, int do_reset, const double *pos,
                 const int32_t *dev_tape, const int32_t *dur_tape, int64_t *obs, double *reward, uint8_t *done,
                 double *now, int64_t *counts, double *power_out, int64_t env_offset, const HsMove *moves, int nmoves);

int hs_run(const HsScenario *sc, int64_t nenv, int nsteps, int do_reset, const double *pos,
           const int32_t *dev_tape, const int32_t *dur_tape, int64_t *obs, double *reward, uint8_t *done,
           double *now, int64_t *counts, double *power_out, int64_t env_offset)
{
    return hs_run_moves(sc, nenv, nsteps, do_reset, pos, dev_tape, dur_tape, obs, reward, done, now, counts, power_out,
                        env_offset, nullptr, 0);
}

int hs_run_moves(const HsScenario *sc, int64_t nenv, int nsteps, int do_reset, const double *pos,
                 const int32_t *dev_tape, const int32_t *dur_tape, int64_t *obs, double *reward, uint8_t *done,
                 double *now, int64_t *counts, double *power_out, int64_t env_offset, const HsMove *moves, int nmoves)
{
    const int ns = sc->band[0].ns, nj = sc->band[0].nj;
    for (int b = 1; b < sc->nbands; ++b)
        if (sc->band[b].ns != ns || sc->band[b].nj != nj) return -1;
#define HS_ARGS *sc, nenv, nsteps, do_reset, pos, dev_tape, dur_tape, obs, reward, done, now, counts, power_out, env_offset, moves, nmoves
    if (sc->mode == MODE_R) {
        if (ns == 2 && nj == 0) return run_t<MODE_R, 3, 2, 0>(HS_ARGS);
        if (ns == 2 && nj == 1) return run_t<MODE_R, 4, 2, 1>(HS_ARGS);
    } else if (sc->mode == MODE_M_FED) {
        if (!g_fed_words) return -2;
        if (ns == 2 && nj == 0) return run_t<MODE_M_FED, 3, 2, 0>(HS_ARGS);
        if (ns == 2 && nj == 1) return run_t<MODE_M_FED, 4, 2, 1>(HS_ARGS);
    } else {
        if (ns == 2 && nj == 0) return run_t<MODE_M_PHILOX, 3, 2, 0>(HS_ARGS);
        if (ns == 2 && nj == 1) return run_t<MODE_M_PHILOX, 4, 2, 1>(HS_ARGS);
    }
#undef HS_ARGS
    return -1;
}

// grid of PHY-only senders (gw_grid.cuh) on the host: one band-sim, `ndur` successive runSimulation(duration)
// calls; per call the clock and the trace records are returned.  Returns the fault code.
int hs_grid_run(int n, double frequency, double bandwidth, const double *power, const double *interval,
                const int32_t *hdr, const int32_t *pay, const double *pos, const double *delays,
                const double *move_delays, const double *offsets, int max_moves, double move_interval,
                const double *durations, int ndur, double *now_out, double *trace, int trace_cap, int32_t *trace_counts,
                uint32_t *stats /* [n][6] */)
{
    GridParams G;
    std::memset(&G, 0, sizeof G);
    G.ndev = n; G.maxMoves = max_moves; G.moveInterval = move_interval;
    G.bitRate = 133.33333e3; G.dataRate = 0.75 * G.bitRate; G.maxBer = 0.25;
    G.tenLog10BitRate = 10 * std::log10(G.bitRate); G.qDen = 1.135 * std::sqrt(2 * 3.141592653589793);
    G.bitsFactor = 1.25; G.frequency = frequency; G.fsplConst = 20 * std::log10(frequency); G.thermal = 1.38e-23 * (20.0 + 273.15) * bandwidth * 1000;
    for (int d = 0; d < n; ++d) { G.power[d] = power[d]; G.interval[d] = interval[d]; G.hdrBytes[d] = hdr[d]; G.payBytes[d] = pay[d]; }
    std::vector<char> block(grid_state_bytes(n));
    GridView v = grid_view(block.data(), n);
    grid_init(v, G, pos, delays, max_moves > 0 ? move_delays : nullptr);
    int used = 0;
    for (int k = 0; k < ndur; ++k) {
        v.trace = trace ? trace + (size_t)used * 8 : nullptr;
        v.traceCap = trace_cap - used; v.ntrace = 0;
        grid_run(v, G, durations[k], offsets);
        now_out[k] = v.h->now;
        trace_counts[k] = v.ntrace;
        used += v.ntrace < v.traceCap ? v.ntrace : v.traceCap;
        if (v.h->fault) return v.h->fault;
    }
    for (int d = 0; d < n; ++d) {
        const GridDev &D = v.dev[d];
        stats[d * 6 + 0] = D.nTx; stats[d * 6 + 1] = D.nHdrOk; stats[d * 6 + 2] = D.nHdrFail;
        stats[d * 6 + 3] = D.nPayOk; stats[d * 6 + 4] = D.nPayFail; stats[d * 6 + 5] = D.nBer;
    }
    return 0;
}

// general band engine (gw_band.cuh) on the host: `nenv` band-sims of one band with ns senders + RRM + nj PHY-only
// senders stepped through an action tape [nsteps][nenv]; state laid out as on the device ([word][sim]).
// cfg_i [6][8]: mult, payloadRule, dest, maxTicks, recv per sender (rows 0-4); cfg_jam_i [2][16]: hdr, payload;
// cfg_d: interval [8], jamInterval [16], jamDelay [16]; pos [nenv or 1][nd][2]; power [nd].
// mode 0: reference accounting; 1: Philox error masks keyed by (seed; env_offset + env, ...).
// Outputs per step and env: obs / reward / done / now; after the last step counts [nenv][1 + 8 + 8]
// (transmissions, deliveries per sender, packets handed to onReceive per sender).  The trace of env 0 is returned
// per step (records of 8 doubles).  `reset_at` >= 0: env.reset() before that step (0: before the first one).
int hs_gen_run(int ns, int nj, int factor, double frequency, double bandwidth, const int32_t *cfg_i, const int32_t *cfg_jam_i,
               const double *cfg_d, const double *pos, int per_env_pos, const double *power, int mode, uint64_t seed,
               int64_t env_offset, int64_t nenv, int nsteps,
               int reset_at, const int32_t *dev_tape, const int32_t *dur_tape, int64_t *obs, double *reward, uint8_t *done,
               double *now_out, int64_t *counts, double *trace, int trace_cap, int32_t *trace_counts,
               const HsMove *moves, int nmoves, const double *move_delays, const double *offsets, int max_moves,
               double move_interval)
{
    Params P;
    std::memset(&P, 0, sizeof P);
    P.nbands = 1; P.factor = factor; P.maxDuration = 20; P.mode = MODE_R;
    P.bitRate = 133.33333e3; P.dataRate = 0.75 * P.bitRate; P.maxBer = 0.25;
    P.tenLog10BitRate = 10 * std::log10(P.bitRate); P.qDen = 1.135 * std::sqrt(2 * 3.141592653589793);
    P.bitsFactor = 1.25;
    finish_params(P);
    GenBand B;
    std::memset(&B, 0, sizeof B);
    B.ns = ns; B.nj = nj; B.nd = ns + 1 + nj; B.maxDuration = 20;
    B.mode = mode; B.seed = seed; B.envOffset = env_offset;
    B.frequency = frequency; B.maxMoves = max_moves; B.moveInterval = move_interval;
    for (int d = 0; d < ns + 1 + nj; ++d) B.power[d] = power[d];
    B.thermal = 1.38e-23 * (20.0 + 273.15) * bandwidth * 1000;
    for (int k = 0; k < ns; ++k) {
        B.mult[k] = cfg_i[0 * 8 + k]; B.payloadRule[k] = cfg_i[1 * 8 + k]; B.dest[k] = cfg_i[2 * 8 + k];
        B.maxTicks[k] = cfg_i[3 * 8 + k]; B.recv[k] = cfg_i[4 * 8 + k]; B.interval[k] = cfg_d[k];
    }
    for (int j = 0; j < nj; ++j) {
        B.jamHdr[j] = cfg_jam_i[j]; B.jamPay[j] = cfg_jam_i[16 + j];
        B.jamInterval[j] = cfg_d[8 + j]; B.jamDelay[j] = cfg_d[24 + j];
    }
    const int nd = B.nd, fw = gen_f64_words(ns, nj), iw = gen_i32_words(ns, nj);
    const bool tables = per_env_pos || nmoves > 0 || max_moves > 0;     // devices that move need the band-sim's own tables
    std::vector<double> f((size_t)fw * nenv), srx((size_t)nd * nd * (tables ? nenv : 1));
    std::vector<double> att(tables ? (size_t)nd * nd * nenv : 1), cur(tables ? (size_t)2 * nd * nenv : 1);
    std::vector<int32_t> iv((size_t)iw * nenv);
    std::vector<double> mvT(max_moves > 0 ? (size_t)2 * nd * nenv : 1);
    std::vector<int32_t> mvI(max_moves > 0 ? (size_t)3 * nd * nenv : 1);
    auto view = [&](int64_t e) {
        GenView v;
        v.f = f.data() + e; v.i = iv.data() + e; v.stride = nenv;
        v.srx = tables ? srx.data() + e : srx.data(); v.srxStride = tables ? nenv : 1;
        v.att = tables ? att.data() + e : nullptr; v.pos = tables ? cur.data() + e : nullptr;
        v.mvT = max_moves > 0 ? mvT.data() + e : nullptr; v.mvDelay = max_moves > 0 ? mvT.data() + (size_t)nd * nenv + e : nullptr;
        v.mvI = max_moves > 0 ? mvI.data() + e : nullptr; v.offsets = offsets;         // (every env the same tape here)
        v.ns = ns; v.nj = nj; v.nd = nd; v.env = env_offset + e; v.mode = mode; v.trace = nullptr; v.ntrace = 0; v.traceCap = 0;
        return v;
    };
    if (tables) for (int64_t e = 0; e < nenv; ++e)
        gen_power_table(nd, pos + (per_env_pos ? (size_t)e * nd * 2 : 0), power, frequency, srx.data() + e, nenv, att.data() + e, cur.data() + e);
    else gen_power_table(nd, pos, power, frequency, srx.data(), 1);
    for (int64_t e = 0; e < nenv; ++e) { GenView v = view(e); gen_init(v, B); if (max_moves > 0) gen_start_movers(v, move_delays); }
    int used = 0, fault = 0;
    for (int t = 0; t < nsteps; ++t) {
        for (int64_t e = 0; e < nenv; ++e) {
            GenView v = view(e);
            if (t == reset_at) gen_reset(v, B);
            if (e == 0 && trace) { v.trace = trace + (size_t)used * 8; v.traceCap = trace_cap - used; }
            {   // devices that jump before this step (all envs alike): successive Position.set calls in the order given
                std::vector<double> want(2 * nd);
                bool any = false;
                for (int k = 0; k < nmoves; ++k) {
                    if (moves[k].step != t) continue;
                    if (!any) { for (int q = 0; q < 2 * nd; ++q) want[q] = cur[(size_t)q * nenv + e]; any = true; }
                    want[2 * moves[k].dev] = moves[k].x; want[2 * moves[k].dev + 1] = moves[k].y;
                }
                if (any) gen_move_devices(v, P, B, want.data());
            }
            long long o; double r; unsigned char d;
            gen_step(v, P, B, dev_tape[(size_t)t * nenv + e], dur_tape[(size_t)t * nenv + e], o, r, d);
            obs[(size_t)t * nenv + e] = o; reward[(size_t)t * nenv + e] = r; done[(size_t)t * nenv + e] = d;
            now_out[(size_t)t * nenv + e] = v.now();
            if (e == 0 && trace) { trace_counts[t] = v.ntrace; used += v.ntrace < v.traceCap ? v.ntrace : v.traceCap; }
            if (v.sc(GenView::I_fault) && !fault) fault = v.sc(GenView::I_fault);
        }
    }
    for (int64_t e = 0; e < nenv; ++e) {
        GenView v = view(e);
        int64_t *c = counts + (size_t)e * 17;
        c[0] = v.sc(GenView::I_nTx);
        for (int k = 0; k < 8; ++k) { c[1 + k] = k < ns ? v.nDeliv(k) : 0; c[9 + k] = k < ns ? v.nRecv(k) : 0; }
    }
    return fault;
}

void hs_set_no_macro(int v) { g_no_macro = v; }

// mode 2 (fed masks): the mask words of the next hs_run* call
void hs_set_fed_masks(const uint32_t *words, int slots, int words_per_row) { g_fed_words = words; g_fed_slots = slots; g_fed_wpr = words_per_row; }

// macro-event statistics since the last call: {isolated transmissions, quiet tails}
void hs_macro_stats(long long *out2) { out2[0] = g_macro_stat[0]; out2[1] = g_macro_stat[1]; g_macro_stat[0] = g_macro_stat[1] = 0; }

// generic-path statistics since the last call: {non-tick events, tick events} through process_event
void hs_generic_stats(long long *out2) { out2[0] = g_macro_stat[2]; out2[1] = g_macro_stat[3]; g_macro_stat[2] = g_macro_stat[3] = 0; }

double hs_ber(double S, double N)
{
    return ber_bpsk_mw(S, N, 10 * std::log10(133.33333e3), 1.135 * std::sqrt(2 * 3.141592653589793));
}

// plant integrator of config 5: advances state[8] = {x, v, theta, omega, vTarget, tPlant, -, -}
void hs_pendulum_advance(const double *params /* M m l g fMax kServo dtMax */, double *state, double now)
{
    PendulumParams Q;
    std::memset(&Q, 0, sizeof Q);
    Q.M = params[0]; Q.m = params[1]; Q.l = params[2]; Q.g = params[3]; Q.fMax = params[4]; Q.kServo = params[5];
    Q.dtMax = params[6];
    PendulumState S;
    S.x = state[0]; S.v = state[1]; S.th = state[2]; S.om = state[3]; S.vTarget = state[4]; S.tPlant = state[5];
    S.ctrlAngleDeg = 0; S.lastError = 0;
    pendulum_advance(Q, S, now);
    state[0] = S.x; state[1] = S.v; state[2] = S.th; state[3] = S.om; state[5] = S.tPlant;
}

// the decider with the division-free shortcut (1) and with the reference's division (0)
int hs_within_max_ber(double err_sum, double total_bits, double max_ber, int shortcut)
{
    Params P;
    std::memset(&P, 0, sizeof P);
    P.bitRate = 133.33333e3; P.dataRate = 0.75 * P.bitRate; P.maxBer = max_ber; P.bitsFactor = 1.25;
    finish_params(P);
    if (!shortcut) P.berMult = 0.0;
    return within_max_ber(P, err_sum, total_bits) ? 1 : 0;
}

double hs_airtime(int bytes)
{
    Params P;
    std::memset(&P, 0, sizeof P);
    P.bitRate = 133.33333e3; P.dataRate = 0.75 * P.bitRate; P.maxBer = 0.25; P.bitsFactor = 1.25;
    finish_params(P);
    return airtime_of(P, bytes);
}

double hs_fmod_slot(double t) { return fmod_slot(t); }

double hs_fspl(double ax, double ay, double bx, double by, double f) { return fspl_db(ax, ay, bx, by, f); }

void hs_philox(const uint32_t *ctr, const uint32_t *key, uint32_t *out)
{
    philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], out);
}

int64_t hs_mask_errors(uint64_t seed, int64_t env, int band, int sender, uint32_t txseq, int receiver,
                       int64_t k0, int64_t k1, double ber)
{
    return mask_errors_serial(seed, env, band, sender, txseq, receiver, k0, k1, ber);
}

}
