/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked, imported or executed by the
 * product (gymwipe_b200/); only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may use it, and only as the checker /
 * the CPU baseline.
 *
 * Plain-C, fp64, single-threaded-per-env restatement of Gym-WiPE's per-step
 * wireless hot path, written as a LITERAL
 * model of the reference's discrete-event structure: every SimPy event the
 * reference creates on this path is an entry of a (time, priority, eid) heap
 * here, created in the same order, and every callback list is replayed in the
 * reference's registration order.  simpy==3.0.11 itself is a third-party
 * dependency absent from /root/reference (Pipfile.lock:176-183); its published
 * scheduling rule (heap key (time, priority, eid), URGENT=0 for process
 * initialisation, NORMAL=1 otherwise) is what the heap below implements.
 *
 * PINNING: this file is checked against the reference ITSELF (run here on the
 * shims of oracle/shims through oracle/ref_harness.py) by
 * oracle/check_restatement.py -- bit-exact obs / reward / done / step end time /
 * every transmission start+stop / every decider input (expected error sums) /
 * every delivery, and BER values -- and against the committed golden traces of
 * tests/golden (tests/test_oracle_golden.py).
 *
 * Reference sites followed (file:line under /root/reference):
 *   step / reset / interpreter ........ gymwipe/envs/counter_traffic.py:53-61,63-112,135-158
 *   feedback order ..................... gymwipe/envs/core.py:142-153
 *   RRM assignment / packet sniffing ... gymwipe/networking/devices.py:84-86,163-203
 *   PHY (power, BER accounting, decider) gymwipe/networking/simple_stack.py:77-286
 *   MAC window loop / queue ............ gymwipe/networking/simple_stack.py:386-471
 *   RRM MAC announcement ............... gymwipe/networking/simple_stack.py:527-561
 *   Transmission / band ................ gymwipe/networking/physical.py:224-290,576-608
 *   Eb/N0, Q approximation, dB helpers . gymwipe/networking/physical.py:25-98
 *   BPSK MCS, max correctable BER ...... gymwipe/networking/physical.py:160-212
 *   FSPL ............................... gymwipe/networking/attenuation_models.py:28-36
 *   distance ........................... gymwipe/devices/core.py:88-95
 *   slot alignment / timeoutUntil ...... gymwipe/simtools.py:44-53,103-116
 *   Notifier executors (blocking/queued) gymwipe/simtools.py:322-412
 *   packet sizes ....................... gymwipe/networking/messages.py:42-75,113-124,154,180
 *
 * Mode R ("reference-exact") reproduces the reference's deterministic expected-
 * value error accounting including its quirks.
 * Mode M ("masked") replaces the accounting by per-bit error masks; see the
 * GWO_MODE_M section.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off, no -ffast-math).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "gw_oracle.h"

/* ---------------------------------------------------------------------- */
/* constants of the reference                                              */
/* ---------------------------------------------------------------------- */

#define SLOT_LEN 1e-6               /* simple_stack.py:27 */
#define MAC_HDR_BYTES 13            /* messages.py:154 */
#define NET_HDR_BYTES 12            /* messages.py:180 */
#define QUEUE_CAP 100               /* simple_stack.py:361 */
#define COUNTER_BOUND 65536         /* counter_traffic.py:35 */
#define COUNTER_BYTE_LENGTH 2       /* counter_traffic.py:33 */

#define PRIO_URGENT 0
#define PRIO_NORMAL 1

enum {
    K_TICK = 1,       /* sender traffic process wakes (counter_traffic.py:53-61) */
    K_JAM_WAKE,       /* PHY-only periodic sender wakes */
    K_PHY_SEND_INIT,  /* Initialize of SimplePhy.macInHandler (simple_stack.py:192) */
    K_RXFIN,          /* nReceivingFinished.event popped (simple_stack.py:200) */
    K_SLOT,           /* nextTimeSlot timeout (simple_stack.py:204) */
    K_Z,              /* zero-delay new-transmission notification (physical.py:601-607) */
    K_RX_INIT,        /* Initialize of SimplePhy._receive */
    K_RX_END,         /* _receive process event -> setRunningFlagToFalse */
    K_H,              /* Transmission.eHeaderCompletes */
    K_C,              /* Transmission.eCompletes */
    K_SEND_DONE,      /* SEND message eProcessed */
    K_PHY_SEND_END,   /* macInHandler process event -> executeNext */
    K_MAC_RX_INIT,    /* Initialize of SimpleMac.phyInHandler */
    K_MAC_RX_END,     /* phyInHandler process event -> setRunningFlagToFalse */
    K_W,              /* MAC window timeoutEvent (simple_stack.py:406) */
    K_PKT_ADDED,      /* SimpleMac._packetAddedEvent (simple_stack.py:470) */
    K_COND,           /* (_packetAddedEvent | timeoutEvent) condition (simple_stack.py:412) */
    K_RRM_ANN_INIT,   /* Initialize of SimpleRrmMac._sendAnnouncement */
    K_RRM_TIMEOUT,    /* (duration+1)*TIME_SLOT_LENGTH timeout (simple_stack.py:558) */
    K_ASSIGN_DONE,    /* ASSIGN message eProcessed (simple_stack.py:561) */
    K_RRM_ANN_END,    /* _sendAnnouncement process event -> executeNext */
    K_STOP,           /* env.run(until=number): StopSimulation event, URGENT (simpy core.py run()) */
    K_MOVE_INIT,      /* Initialize of a mobility process (tests/test_benchmark.py:73-85) */
    K_MOVE,           /* its timeout: Position.set, then the next timeout */
    K_RECV_INIT,      /* Initialize of SimpleNetworkDevice._receiver (devices.py:88-97) */
    K_RECV_TIMEOUT,   /* SimpleMac._receiveTimeout (simple_stack.py:458-459, 473-478) */
    K_RECV_DONE       /* RECEIVE message eProcessed: the receiver loop resumes */
};

#define RECEIVE_TIMEOUT 100.0       /* SimpleNetworkDevice.RECEIVE_TIMEOUT, devices.py:66 */

enum { MAC_NONE = 0, MAC_WAIT_COND, MAC_WAIT_TX, MAC_IDLE };
enum { PKT_ANNOUNCE = 1, PKT_DATA, PKT_JAM };
enum { ORG_MAC = 1, ORG_RRM, ORG_JAM };

typedef struct {
    double t;
    int prio;
    uint64_t eid;
    int kind, band, a, b;
} Ev;

typedef struct {
    int type;          /* PKT_* */
    int src;           /* device index of the MAC-level source */
    int dst;           /* device index of the MAC-level destination (announce: grantee) */
    int hdr_bytes;
    int pay_bytes;
    double slots;      /* announcement payload value */
    uint32_t seq;      /* per-sender transmission sequence number (mode M key) */
} Pkt;

typedef struct {
    Pkt pkt;
    double power;
    int origin;        /* ORG_* : who waits for eProcessed */
} SendCmd;

#define SENDQ_CAP 64

typedef struct {
    int used;
    int sender;
    double power, start, hd, pd, stop, hdrBits, payBits;
    Pkt pkt;
    /* callback lists in registration order */
    int h_n, h_list[GWO_MAXDEV];        /* receivers waiting on eHeaderCompletes */
    int cc_n, cc_list[GWO_MAXDEV];      /* _onCompletingTransmission callbacks */
    int cr_n, cr_list[GWO_MAXDEV];      /* receivers waiting on eCompletes */
    int origin;
} Tx;

typedef struct {
    /* static */
    int role;
    double x, y;
    int mult, payload_rule, dest;
    double interval;
    double jam_interval, jam_delay, jam_power;
    int jam_hdr, jam_payload;
    /* PHY */
    int transmitting, cur_tx;
    int receiving, rx_tx, rx_section;
    double P, S[GWO_MAXTX];
    int hasS[GWO_MAXTX];
    double ber, errSum, tReset;
    int rx_running;
    int send_running, wait_rxfin;
    SendCmd cur_cmd;
    SendCmd sendq[SENDQ_CAP];
    int sendq_h, sendq_n;
    uint32_t tx_seq;
    /* mode M reception bookkeeping */
    double seg_t0;             /* start of the current constant-BER segment */
    int64_t err_int;           /* integer error count of the current section */
    /* MAC (sender) */
    int q_size[QUEUE_CAP];     /* data byteSize of queued packets (ring) */
    int q_h, q_n;
    uint64_t enq_total, pop_total;
    uint32_t pkt_gen;          /* generation of the current _packetAddedEvent */
    int mac_running, mac_state;
    int w_processed, cond_triggered;
    uint32_t cond_gen;
    double stopW;
    Pkt mac_rx_pkt;
    /* traffic */
    int counter;
    int jam_stage;
    /* finite traffic bursts and MAC receive mode */
    int max_ticks, ticks_done;
    int recv_mode, mac_receiving;
    uint32_t recv_gen;
    int64_t n_received;
    /* mobility process (gwo_add_mover) */
    double mv_x0, mv_y0, mv_first, mv_interval;
    const double *mv_offsets;
    int mv_n, mv_k;
} Dev;

typedef struct {
    int ndev, rrm;
    double frequency, bandwidth;
    double thermal;
    double att[GWO_MAXDEV][GWO_MAXDEV];
    Dev dev[GWO_MAXDEV];
    Tx tx[GWO_MAXTX];
    /* RRM MAC announcement executor (queued) */
    int ann_running;
    struct { int dev; double slots; int nbytes; uint32_t seq; } annq[8], ann_cur;
    int annq_h, annq_n;
    uint32_t assign_seq, assign_done_seq;
    /* interpreter (counter_traffic.py:63-112) */
    int rv[2];
    int latestDiff, lastAbsDiff, done;
    /* statistics (oracle-side, not in the reference) */
    int64_t n_tx, n_deliv[GWO_MAXDEV];
} Band;

#define HEAP_CAP 4096

struct gwo_sim {
    int nbands;
    int factor;
    int mode;
    double now;
    uint64_t eid;
    Ev heap[HEAP_CAP];
    int heap_n;
    Band band[GWO_MAXBAND];
    /* mcs (physical.py:192-197) */
    double bitRate, dataRate, codeRate, maxBer, tenLog10BitRate;
    int fault;
    int64_t popped;
    /* trace */
    int trace_on;
    double *rec;
    size_t rec_n, rec_cap;
    /* near-tie diagnostics */
    int64_t near_round_tie, near_ber_tie;
    /* mode M */
    gwo_mask_fn mask_fn;
    void *mask_ctx;
    uint64_t mask_seed;
    int64_t env_id;
};

/* ---------------------------------------------------------------------- */
/* trace                                                                   */
/* ---------------------------------------------------------------------- */

static void rec_push(gwo_sim *s, double kind, double t, double band, double dev,
                     double x0, double x1, double x2, double x3)
{
    if (!s->trace_on) return;
    if (s->rec_n + 8 > s->rec_cap) {
        size_t nc = s->rec_cap ? s->rec_cap * 2 : 4096;
        s->rec = (double *)realloc(s->rec, nc * sizeof(double));
        s->rec_cap = nc;
    }
    double *r = s->rec + s->rec_n;
    r[0] = kind; r[1] = t; r[2] = band; r[3] = dev;
    r[4] = x0; r[5] = x1; r[6] = x2; r[7] = x3;
    s->rec_n += 8;
}

/* ---------------------------------------------------------------------- */
/* heap keyed (time, priority, eid)                                        */
/* ---------------------------------------------------------------------- */

static int ev_less(const Ev *a, const Ev *b)
{
    if (a->t != b->t) return a->t < b->t;
    if (a->prio != b->prio) return a->prio < b->prio;
    return a->eid < b->eid;
}

static void heap_push(gwo_sim *s, Ev e)
{
    if (s->heap_n >= HEAP_CAP) { s->fault = GWO_FAULT_HEAP; return; }
    int i = s->heap_n++;
    s->heap[i] = e;
    while (i > 0) {
        int p = (i - 1) / 2;
        if (!ev_less(&s->heap[i], &s->heap[p])) break;
        Ev tmp = s->heap[i]; s->heap[i] = s->heap[p]; s->heap[p] = tmp;
        i = p;
    }
}

static Ev heap_pop(gwo_sim *s)
{
    Ev top = s->heap[0];
    s->heap[0] = s->heap[--s->heap_n];
    int i = 0;
    for (;;) {
        int l = 2 * i + 1, r = l + 1, m = i;
        if (l < s->heap_n && ev_less(&s->heap[l], &s->heap[m])) m = l;
        if (r < s->heap_n && ev_less(&s->heap[r], &s->heap[m])) m = r;
        if (m == i) break;
        Ev tmp = s->heap[i]; s->heap[i] = s->heap[m]; s->heap[m] = tmp;
        i = m;
    }
    return top;
}

/* env.schedule(event, priority, delay): time = now + delay, eid = next(counter) */
static void schedule(gwo_sim *s, int kind, int prio, double delay, int band, int a, int b)
{
    Ev e;
    e.t = s->now + delay;
    e.prio = prio;
    e.eid = s->eid++;
    e.kind = kind; e.band = band; e.a = a; e.b = b;
    heap_push(s, e);
}

/* an event the reference creates but nobody ever listens to: only the eid advances */
static void schedule_unobserved(gwo_sim *s) { s->eid++; }

/* ---------------------------------------------------------------------- */
/* physical-layer arithmetic (physical.py:25-98,208-212)                    */
/* ---------------------------------------------------------------------- */

/* CPython evaluates `x ** y` on floats through libm pow(); gcc would fold pow(x, 2.0)
 * into x*x (1 ulp away from glibc's pow in rare cases), so every power goes through a
 * volatile function pointer (and the Makefile passes -fno-builtin-pow). */
static double (*volatile libm_pow)(double, double) = pow;
#define pow(a, b) libm_pow((a), (b))

static double mw_to_dbm(double mw) { return 10 * log10(mw); }          /* physical.py:82-89 */
static double dbm_to_mw(double dbm) { return pow(10.0, dbm / 10); }    /* physical.py:91-98 */

double gwo_q_function(double x)                                         /* physical.py:46-58 */
{
    const double e = 2.718281828459045;   /* math.e */
    const double sqrtOfTwoPi = sqrt(2 * 3.141592653589793);
    return (1 - pow(e, -1.4 * x)) * pow(e, -(pow(x, 2.0) / 2)) / (1.135 * sqrtOfTwoPi * x);
}

double gwo_ber_bpsk(double s_dbm, double n_dbm, double bitRate)         /* physical.py:208-212,25-42 */
{
    if (s_dbm <= n_dbm) return 0.5;
    double ratio_db = s_dbm - n_dbm - 10 * log10(bitRate);
    double ratio = pow(10.0, ratio_db / 10);
    return gwo_q_function(sqrt(2 * ratio));
}

double gwo_fspl(double ax, double ay, double bx, double by, double frequency)
{
    /* attenuation_models.py:28-36 ; devices/core.py:88-95 ; equal positions keep 0 dB */
    if (ax == bx && ay == by) return 0.0;
    double d = sqrt(pow(ax - bx, 2.0) + pow(ay - by, 2.0));
    return 20 * log10(d) + 20 * log10(frequency) - 147.55;
}

double gwo_thermal_noise_mw(double bandwidth)                           /* simple_stack.py:57,77 */
{
    double npd = 1.38e-23 * (20.0 + 273.15);                            /* physical.py:71 */
    return npd * bandwidth * 1000;
}

double gwo_max_correctable_ber(int k, int n)                            /* physical.py:160-185 */
{
    double bound = pow(2.0, (double)(n - k));
    double sum = 0;
    int t = 0;
    while (sum <= bound) {
        /* scipy.special.binom(n, t) for small integers */
        double c = 1;
        for (int i = 1; i <= t; i++) c = c * (double)(n - t + i) / (double)i;
        sum += c;
        t += 1;
    }
    t -= 1;
    return (double)t / n;
}

/* ---------------------------------------------------------------------- */
/* PHY bookkeeping                                                          */
/* ---------------------------------------------------------------------- */

static double py_mod_slot(double now)
{
    /* Python float % for positive operands == C fmod (simtools.py:53) */
    return fmod(now, SLOT_LEN);
}

static int tx_completed(gwo_sim *s, Tx *t) { return s->now >= t->stop; }  /* physical.py:284-290 */

static void update_ber(gwo_sim *s, int b, int d, int txi)               /* simple_stack.py:161-173 */
{
    Band *B = &s->band[b];
    Dev *D = &B->dev[d];
    if (!D->hasS[txi]) { s->fault = GWO_FAULT_REF_KEYERROR; return; }
    double S = D->S[txi];
    double N = D->P - S;
    if (!(S >= 0) || !(N >= 0)) { s->fault = GWO_FAULT_REF_ASSERT; return; }
    double sd = mw_to_dbm(S), nd = mw_to_dbm(N);
    D->ber = gwo_ber_bpsk(sd, nd, s->bitRate);
    rec_push(s, GWO_REC_BER, s->now, b, d, D->ber, 0, 0, 0);
}

static void reset_counter(gwo_sim *s, Dev *D)                           /* simple_stack.py:175-178 */
{
    D->errSum = 0;
    D->ber = 0.0;
    D->tReset = s->now;
    D->seg_t0 = s->now;
    D->err_int = 0;
}

/* mode M: integer errors of the on-air bits [floor((t0-start)R), floor((t1-start)R)) */
static void count_masked(gwo_sim *s, int b, int d)
{
    Band *B = &s->band[b];
    Dev *D = &B->dev[d];
    Tx *t = &B->tx[D->rx_tx];
    double t1 = s->now;
    int64_t k0 = (int64_t)floor((D->seg_t0 - t->start) * s->bitRate);
    int64_t k1 = (int64_t)floor((t1 - t->start) * s->bitRate);
    if (k1 > k0 && s->mask_fn) {
        D->err_int += s->mask_fn(s->mask_ctx, s->env_id, b, t->sender, t->pkt.seq, d,
                                 k0, k1, D->ber);
    }
    D->seg_t0 = t1;
}

static void count_errors(gwo_sim *s, int b, int d)                      /* simple_stack.py:180-188 */
{
    Dev *D = &s->band[b].dev[d];
    if (s->mode == GWO_MODE_M) { count_masked(s, b, d); return; }
    double duration = s->now - D->tReset;
    double bitErrors = D->ber * duration * s->bitRate;
    D->errSum += bitErrors;
}

static int decide(gwo_sim *s, int b, int d, int section, double totalBits)  /* simple_stack.py:269-286 */
{
    Dev *D = &s->band[b].dev[d];
    double sum = (s->mode == GWO_MODE_M) ? (double)D->err_int : D->errSum;
    double r = rint(sum);                       /* Python round(): half-even */
    double frac = sum - floor(sum);
    if (fabs(frac - 0.5) < 1e-9) s->near_round_tie++;
    double rate = r / totalBits;
    if (fabs(rate - s->maxBer) < 1e-12) s->near_ber_tie++;
    int ok = rate <= s->maxBer;
    rec_push(s, GWO_REC_DEC, s->now, b, d, section, sum, totalBits, ok);
    return ok;
}

/* _nReceivedPowerChanges.trigger(delta): priority 1 updateReceivedPower, then
 * priority 0 onReceivedPowerChange if a reception is subscribed (simple_stack.py:81-86,223-233) */
static void power_change(gwo_sim *s, int b, int d, double delta)
{
    Band *B = &s->band[b];
    Dev *D = &B->dev[d];
    D->P += delta;
    if (D->receiving) {
        if (delta != 0) {
            count_errors(s, b, d);
            if (!tx_completed(s, &B->tx[D->rx_tx]))
                update_ber(s, b, d, D->rx_tx);
        }
    }
}

/* end of SimplePhy._receive (simple_stack.py:264-267) */
static void rx_finish(gwo_sim *s, int b, int d)
{
    Dev *D = &s->band[b].dev[d];
    reset_counter(s, D);
    D->receiving = 0;
    /* _nReceivingFinished.trigger(): succeeds .event only if somebody asked for it */
    if (D->wait_rxfin) {
        D->wait_rxfin = 0;
        schedule(s, K_RXFIN, PRIO_NORMAL, 0, b, d, 0);
    }
    /* generator ends -> process event (setRunningFlagToFalse) */
    schedule(s, K_RX_END, PRIO_NORMAL, 0, b, d, 0);
}

/* SimplePhy "macIn" gate, queued executor (construction.py:290, simtools.py:347-381) */
static void phy_send(gwo_sim *s, int b, int d, SendCmd cmd)
{
    Dev *D = &s->band[b].dev[d];
    if (D->send_running) {
        if (D->sendq_n >= SENDQ_CAP) { s->fault = GWO_FAULT_SENDQ; return; }
        D->sendq[(D->sendq_h + D->sendq_n) % SENDQ_CAP] = cmd;
        D->sendq_n++;
    } else {
        D->send_running = 1;
        D->cur_cmd = cmd;
        schedule(s, K_PHY_SEND_INIT, PRIO_URGENT, 0, b, d, 0);
    }
}

static void phy_begin_slot_wait(gwo_sim *s, int b, int d)               /* simple_stack.py:202-204 */
{
    Dev *D = &s->band[b].dev[d];
    D->transmitting = 1;
    double delay = SLOT_LEN - py_mod_slot(s->now);
    schedule(s, K_SLOT, PRIO_NORMAL, delay, b, d, 0);
}

/* ---------------------------------------------------------------------- */
/* MAC                                                                      */
/* ---------------------------------------------------------------------- */

/* SimpleMac.networkInHandler for a Packet (simple_stack.py:463-471) */
static void mac_enqueue(gwo_sim *s, int b, int d, int data_bytes)
{
    Dev *D = &s->band[b].dev[d];
    if (D->q_n == QUEUE_CAP) {           /* deque(maxlen=100): drop the oldest */
        D->q_h = (D->q_h + 1) % QUEUE_CAP;
        D->q_n--;
    }
    D->q_size[(D->q_h + D->q_n) % QUEUE_CAP] = data_bytes;
    D->q_n++;
    D->enq_total++;
    /* self._packetAddedEvent.succeed(); self._packetAddedEvent = Event(...) */
    if (D->mac_state == MAC_WAIT_COND && D->cond_gen == D->pkt_gen)
        schedule(s, K_PKT_ADDED, PRIO_NORMAL, 0, b, d, (int)D->pkt_gen);
    else
        schedule_unobserved(s);
    D->pkt_gen++;
}

static void mac_end(gwo_sim *s, int b, int d)
{
    Dev *D = &s->band[b].dev[d];
    D->mac_state = MAC_NONE;
    schedule(s, K_MAC_RX_END, PRIO_NORMAL, 0, b, d, 0);
}

/* the while-loop of SimpleMac.phyInHandler (simple_stack.py:408-434).
 * entry: 0 = loop head, 1 = resumed from the condition */
static void mac_loop(gwo_sim *s, int b, int d, int entry)
{
    Band *B = &s->band[b];
    Dev *D = &B->dev[d];
    int queuedPackets = 1;
    if (entry == 1) {
        /* after `yield self._packetAddedEvent | timeoutEvent` */
        queuedPackets = 0;
        if (!D->w_processed) queuedPackets = 1;
        goto after_empty_check;
    }
    for (;;) {
        if (D->w_processed) { mac_end(s, b, d); return; }
        if (D->q_n == 0) {
            /* Condition(any, [packetAdded, timeout]): both still pending */
            D->mac_state = MAC_WAIT_COND;
            D->cond_gen = D->pkt_gen;
            D->cond_triggered = 0;
            return;
        }
after_empty_check:
        if (queuedPackets) {
            int bitSize = (MAC_HDR_BYTES + NET_HDR_BYTES + D->q_size[D->q_h]) * 8;
            double timeLeft = D->stopW - s->now;
            double txTime = bitSize / s->dataRate;
            if (!(timeLeft > txTime)) {
                D->mac_state = MAC_IDLE;        /* yield timeoutEvent */
                return;
            }
            int data_bytes = D->q_size[D->q_h];
            D->q_h = (D->q_h + 1) % QUEUE_CAP;
            D->q_n--;
            D->pop_total++;
            SendCmd c;
            memset(&c, 0, sizeof c);
            c.pkt.type = PKT_DATA;
            c.pkt.src = d;
            c.pkt.dst = D->dest;
            c.pkt.hdr_bytes = MAC_HDR_BYTES;
            c.pkt.pay_bytes = NET_HDR_BYTES + data_bytes;
            c.power = 0.0;                       /* simple_stack.py:364 */
            c.origin = ORG_MAC;
            D->mac_state = MAC_WAIT_TX;          /* yield message.eProcessed */
            phy_send(s, b, d, c);
            return;
        }
        /* queuedPackets is False only when the timeout was processed: the loop head ends it */
    }
}

/* ---------------------------------------------------------------------- */
/* event handlers                                                           */
/* ---------------------------------------------------------------------- */

static int tx_alloc(Band *B)
{
    for (int i = 0; i < GWO_MAXTX; i++) if (!B->tx[i].used) return i;
    return -1;
}

static void on_slot(gwo_sim *s, int b, int d)       /* simple_stack.py:206-209, physical.py:224-279,596-608 */
{
    Band *B = &s->band[b];
    Dev *D = &B->dev[d];
    int ti = tx_alloc(B);
    if (ti < 0) { s->fault = GWO_FAULT_TXPOOL; return; }
    Tx *t = &B->tx[ti];
    memset(t, 0, sizeof *t);
    t->used = 1;
    t->sender = d;
    t->power = D->cur_cmd.power;
    t->pkt = D->cur_cmd.pkt;
    t->pkt.seq = D->tx_seq++;
    t->origin = D->cur_cmd.origin;
    t->start = s->now;
    int hdrBitSize = t->pkt.hdr_bytes * 8, payBitSize = t->pkt.pay_bytes * 8;
    t->hd = hdrBitSize / s->dataRate;
    t->pd = payBitSize / s->dataRate;
    double duration = t->hd + t->pd;
    t->stop = t->start + duration;
    t->hdrBits = hdrBitSize * (2 - s->codeRate);
    t->payBits = payBitSize * (2 - s->codeRate);
    /* SimMan.timeoutUntil(headerStopTime), timeoutUntil(stopTime), timeout(0) */
    double headerStop = t->start + t->hd;
    double dH = headerStop > s->now ? headerStop - s->now : 0;
    schedule(s, K_H, PRIO_NORMAL, dH, b, ti, 0);
    double dC = t->stop > s->now ? t->stop - s->now : 0;
    schedule(s, K_C, PRIO_NORMAL, dC, b, ti, 0);
    schedule(s, K_Z, PRIO_NORMAL, 0, b, ti, 0);
    D->cur_tx = ti;
    B->n_tx++;
    rec_push(s, GWO_REC_TX, t->start, b, d, t->stop, t->hdrBits, t->payBits, 0);
}

static void on_z(gwo_sim *s, int b, int ti)         /* physical.py:601-602, simple_stack.py:130-144 */
{
    Band *B = &s->band[b];
    Tx *t = &B->tx[ti];
    /* callbacks: every PHY's _onNewTransmission (order irrelevant: each touches its own PHY) */
    for (int p = 0; p < B->ndev; p++) {
        if (p == t->sender) continue;           /* `t is not self._currentTransmission` */
        Dev *P = &B->dev[p];
        double rp = dbm_to_mw(t->power - B->att[p][t->sender]);
        P->S[ti] = rp;
        P->hasS[ti] = 1;
        power_change(s, b, p, rp);
        t->cc_list[t->cc_n++] = p;
    }
    /* executors in subscription (= PHY construction) order: blocking, not queued */
    for (int p = 0; p < B->ndev; p++) {
        Dev *P = &B->dev[p];
        if (P->rx_running) continue;
        P->rx_running = 1;
        schedule(s, K_RX_INIT, PRIO_URGENT, 0, b, p, ti);
    }
}

static void on_rx_init(gwo_sim *s, int b, int p, int ti)   /* simple_stack.py:214-238 */
{
    Band *B = &s->band[b];
    Dev *P = &B->dev[p];
    Tx *t = &B->tx[ti];
    if (!P->transmitting) {
        P->receiving = 1;
        P->rx_tx = ti;
        P->rx_section = 0;
        reset_counter(s, P);
        update_ber(s, b, p, ti);
        t->h_list[t->h_n++] = p;                /* yield t.eHeaderCompletes */
    } else {
        schedule(s, K_RX_END, PRIO_NORMAL, 0, b, p, 0);
    }
}

static void on_h(gwo_sim *s, int b, int ti)         /* simple_stack.py:241-251 */
{
    Band *B = &s->band[b];
    Tx *t = &B->tx[ti];
    for (int i = 0; i < t->h_n; i++) {
        int p = t->h_list[i];
        Dev *P = &B->dev[p];
        count_errors(s, b, p);
        if (decide(s, b, p, 0, t->hdrBits)) {
            P->rx_section = 1;
            reset_counter(s, P);
            update_ber(s, b, p, ti);
            t->cr_list[t->cr_n++] = p;          /* yield t.eCompletes */
        } else {
            rx_finish(s, b, p);
        }
        if (s->fault) return;
    }
}

static void deliver(gwo_sim *s, int b, int p, Tx *t)   /* simple_stack.py:260, devices.py:163-168 */
{
    Band *B = &s->band[b];
    Dev *P = &B->dev[p];
    if (P->role == GWO_ROLE_SENDER) {
        /* SimpleMac "phyIn": generator, blocking, not queued */
        if (!P->mac_running) {
            P->mac_running = 1;
            P->mac_rx_pkt = t->pkt;
            schedule(s, K_MAC_RX_INIT, PRIO_URGENT, 0, b, p, 0);
        }
    } else if (P->role == GWO_ROLE_RRM) {
        /* SimpleRrmMac.phyInHandler -> networkOut -> interpreter.onPacketReceived */
        if (t->pkt.type == PKT_DATA) {
            int k = t->pkt.src;                 /* sender index == device index (senders first) */
            if (k < 2) {
                B->rv[k] = COUNTER_BYTE_LENGTH; /* payload.value, appendix B #1 */
                B->latestDiff = B->rv[0] - B->rv[1];
            }
            B->n_deliv[k]++;
            rec_push(s, GWO_REC_RX, s->now, b, k, 0, 0, 0, 0);
        } else if (t->pkt.type == PKT_JAM) {
            rec_push(s, GWO_REC_RX, s->now, b, t->pkt.src, 0, 0, 0, 0);
        }
    }
}

static void on_c(gwo_sim *s, int b, int ti)         /* callbacks of eCompletes in registration order */
{
    Band *B = &s->band[b];
    Tx *t = &B->tx[ti];
    /* 1. sender PHY macInHandler resumes (simple_stack.py:209-212) */
    {
        Dev *D = &B->dev[t->sender];
        D->transmitting = 0;
        schedule(s, K_SEND_DONE, PRIO_NORMAL, 0, b, t->sender, t->origin);
        schedule(s, K_PHY_SEND_END, PRIO_NORMAL, 0, b, t->sender, 0);
    }
    /* 2. _onCompletingTransmission of the other PHYs (simple_stack.py:146-157) */
    for (int i = 0; i < t->cc_n; i++) {
        int p = t->cc_list[i];
        Dev *P = &B->dev[p];
        double rp = P->S[ti];
        P->hasS[ti] = 0;
        power_change(s, b, p, -rp);
        if (s->fault) return;
    }
    /* 3. receivers that passed the header (simple_stack.py:252-267) */
    for (int i = 0; i < t->cr_n; i++) {
        int p = t->cr_list[i];
        count_errors(s, b, p);
        if (decide(s, b, p, 1, t->payBits))
            deliver(s, b, p, t);
        rx_finish(s, b, p);
    }
    t->used = 0;
}

static void on_mac_rx_init(gwo_sim *s, int b, int d)   /* simple_stack.py:386-448 */
{
    Band *B = &s->band[b];
    Dev *D = &B->dev[d];
    Pkt *p = &D->mac_rx_pkt;
    if (p->type == PKT_ANNOUNCE && p->dst == d) {
        double timeTotal = p->slots * SLOT_LEN;
        D->stopW = s->now + timeTotal;
        D->w_processed = 0;
        schedule(s, K_W, PRIO_NORMAL, timeTotal, b, d, 0);
        mac_loop(s, b, d, 0);
    } else {
        /* a packet from another device, addressed to us (simple_stack.py:436-444): in receive mode its payload goes
         * to the network layer -- setProcessed(payload), _stopReceiving(); otherwise ignored */
        if (p->type == PKT_DATA && p->dst == d && D->mac_receiving) {
            D->mac_receiving = 0;
            schedule(s, K_RECV_DONE, PRIO_NORMAL, 0, b, d, 1);
        }
        mac_end(s, b, d);
    }
}

static void rrm_start_announcement(gwo_sim *s, int b)
{
    schedule(s, K_RRM_ANN_INIT, PRIO_URGENT, 0, b, 0, 0);
}

/* SimpleRrmDevice.assignFrequencyBand (devices.py:178-203) */
static void rrm_assign(gwo_sim *s, int b, int dev, double slots, int nbytes)
{
    Band *B = &s->band[b];
    uint32_t seq = ++B->assign_seq;
    if (B->ann_running) {
        if (B->annq_n >= 8) { s->fault = GWO_FAULT_SENDQ; return; }
        int i = (B->annq_h + B->annq_n) % 8;
        B->annq[i].dev = dev; B->annq[i].slots = slots; B->annq[i].nbytes = nbytes; B->annq[i].seq = seq;
        B->annq_n++;
    } else {
        B->ann_running = 1;
        B->ann_cur.dev = dev; B->ann_cur.slots = slots; B->ann_cur.nbytes = nbytes; B->ann_cur.seq = seq;
        rrm_start_announcement(s, b);
    }
}

/* one pass of `while self._receiving:` in SimpleNetworkDevice._receiver (devices.py:88-93): a RECEIVE message goes
 * to the MAC's networkIn gate (callback): _receiving = True and a fresh timeout (simple_stack.py:452-460) */
static void receiver_issue(gwo_sim *s, int b, int d)
{
    Dev *D = &s->band[b].dev[d];
    D->mac_receiving = 1;
    D->recv_gen++;
    schedule(s, K_RECV_TIMEOUT, PRIO_NORMAL, RECEIVE_TIMEOUT, b, d, (int)D->recv_gen);
}

static void dispatch(gwo_sim *s, Ev e)
{
    int b = e.band;
    Band *B = &s->band[b];
    switch (e.kind) {
    case K_RECV_INIT:
        receiver_issue(s, b, e.a);
        break;
    case K_RECV_TIMEOUT: {
        /* _receiveTimeoutCallback: only the CURRENT receive command's timeout has an effect */
        Dev *D = &B->dev[e.a];
        if (D->mac_receiving && (uint32_t)e.b == D->recv_gen) {
            D->mac_receiving = 0;                                 /* setProcessed() without a result, _stopReceiving() */
            schedule(s, K_RECV_DONE, PRIO_NORMAL, 0, b, e.a, 0);
        }
        break;
    }
    case K_RECV_DONE: {
        Dev *D = &B->dev[e.a];
        if (e.b) {                                                /* `if result: self.onReceive(result)` */
            D->n_received++;
            rec_push(s, GWO_REC_MRX, s->now, b, e.a, 0, 0, 0, 0);
        }
        receiver_issue(s, b, e.a);                                /* the device-level flag stays set: next RECEIVE */
        break;
    }
    case K_TICK: {
        Dev *D = &B->dev[e.a];
        if (D->max_ticks > 0 && D->ticks_done >= D->max_ticks) { schedule_unobserved(s); break; }   /* the process ends */
        D->ticks_done++;
        for (int i = 0; i < D->mult; i++) {
            int bytes = D->payload_rule < 0 ? D->counter : D->payload_rule;
            mac_enqueue(s, b, e.a, bytes);
        }
        if (D->counter < COUNTER_BOUND) D->counter += 1;
        schedule(s, K_TICK, PRIO_NORMAL, D->interval, b, e.a, 0);
        break;
    }
    case K_JAM_WAKE: {
        Dev *D = &B->dev[e.a];
        if (D->jam_stage == 0) {
            D->jam_stage = 1;
            schedule(s, K_JAM_WAKE, PRIO_NORMAL, D->jam_delay, b, e.a, 0);
        } else if (D->jam_stage == 1) {
            D->jam_stage = 2;
            schedule(s, K_JAM_WAKE, PRIO_NORMAL, D->jam_interval, b, e.a, 0);
        } else {
            SendCmd c;
            memset(&c, 0, sizeof c);
            c.pkt.type = PKT_JAM;
            c.pkt.src = e.a;
            c.pkt.dst = e.a;
            c.pkt.hdr_bytes = D->jam_hdr;
            c.pkt.pay_bytes = D->jam_payload;
            c.power = D->jam_power;
            c.origin = ORG_JAM;
            phy_send(s, b, e.a, c);
            schedule(s, K_JAM_WAKE, PRIO_NORMAL, D->jam_interval, b, e.a, 0);
        }
        break;
    }
    case K_PHY_SEND_INIT: {
        Dev *D = &B->dev[e.a];
        if (D->receiving) D->wait_rxfin = 1;      /* yield self._nReceivingFinished.event */
        else phy_begin_slot_wait(s, b, e.a);
        break;
    }
    case K_RXFIN:
        phy_begin_slot_wait(s, b, e.a);
        break;
    case K_SLOT:
        on_slot(s, b, e.a);
        break;
    case K_Z:
        on_z(s, b, e.a);
        break;
    case K_RX_INIT:
        on_rx_init(s, b, e.a, e.b);
        break;
    case K_RX_END:
        B->dev[e.a].rx_running = 0;
        break;
    case K_H:
        on_h(s, b, e.a);
        break;
    case K_C:
        on_c(s, b, e.a);
        break;
    case K_SEND_DONE: {
        if (e.b == ORG_MAC) {
            /* SimpleMac window loop resumes after `yield message.eProcessed` */
            mac_loop(s, b, e.a, 0);
        } else if (e.b == ORG_RRM) {
            /* _sendAnnouncement: yield timeout((duration+1)*TIME_SLOT_LENGTH) */
            double d = (B->ann_cur.slots + 1) * SLOT_LEN;
            schedule(s, K_RRM_TIMEOUT, PRIO_NORMAL, d, b, 0, 0);
        }
        break;
    }
    case K_PHY_SEND_END: {
        Dev *D = &B->dev[e.a];
        if (D->sendq_n > 0) {
            D->cur_cmd = D->sendq[D->sendq_h];
            D->sendq_h = (D->sendq_h + 1) % SENDQ_CAP;
            D->sendq_n--;
            schedule(s, K_PHY_SEND_INIT, PRIO_URGENT, 0, b, e.a, 0);
        } else {
            D->send_running = 0;
        }
        break;
    }
    case K_MAC_RX_INIT:
        on_mac_rx_init(s, b, e.a);
        break;
    case K_MAC_RX_END:
        B->dev[e.a].mac_running = 0;
        break;
    case K_W: {
        Dev *D = &B->dev[e.a];
        D->w_processed = 1;
        if (D->mac_state == MAC_WAIT_COND) {
            if (!D->cond_triggered) {
                D->cond_triggered = 1;
                schedule(s, K_COND, PRIO_NORMAL, 0, b, e.a, 0);
            }
        } else if (D->mac_state == MAC_IDLE) {
            mac_loop(s, b, e.a, 0);               /* loop condition fails -> generator ends */
        }
        break;
    }
    case K_PKT_ADDED: {
        Dev *D = &B->dev[e.a];
        if (D->mac_state == MAC_WAIT_COND && D->cond_gen == (uint32_t)e.b && !D->cond_triggered) {
            D->cond_triggered = 1;
            schedule(s, K_COND, PRIO_NORMAL, 0, b, e.a, 0);
        }
        break;
    }
    case K_COND:
        mac_loop(s, b, e.a, 1);
        break;
    case K_RRM_ANN_INIT: {
        SendCmd c;
        memset(&c, 0, sizeof c);
        c.pkt.type = PKT_ANNOUNCE;
        c.pkt.src = B->rrm;
        c.pkt.dst = B->ann_cur.dev;
        c.pkt.hdr_bytes = MAC_HDR_BYTES;
        c.pkt.pay_bytes = B->ann_cur.nbytes;
        c.pkt.slots = B->ann_cur.slots;
        c.power = 0.0;                             /* simple_stack.py:521 */
        c.origin = ORG_RRM;
        phy_send(s, b, B->rrm, c);
        break;
    }
    case K_RRM_TIMEOUT:
        schedule(s, K_ASSIGN_DONE, PRIO_NORMAL, 0, b, (int)B->ann_cur.seq, 0);
        schedule(s, K_RRM_ANN_END, PRIO_NORMAL, 0, b, 0, 0);
        break;
    case K_ASSIGN_DONE:
        B->assign_done_seq = (uint32_t)e.a;
        break;
    case K_RRM_ANN_END:
        if (B->annq_n > 0) {
            B->ann_cur.dev = B->annq[B->annq_h].dev;
            B->ann_cur.slots = B->annq[B->annq_h].slots;
            B->ann_cur.nbytes = B->annq[B->annq_h].nbytes;
            B->ann_cur.seq = B->annq[B->annq_h].seq;
            B->annq_h = (B->annq_h + 1) % 8;
            B->annq_n--;
            rrm_start_announcement(s, b);
        } else {
            B->ann_running = 0;
        }
        break;
    case K_STOP:
        break;
    case K_MOVE_INIT:
        /* yield SimMan.timeout(random.uniform(0, MOVE_INTERVAL)) */
        schedule(s, K_MOVE, PRIO_NORMAL, B->dev[e.a].mv_first, b, e.a, 0);
        break;
    case K_MOVE: {
        Dev *D = &B->dev[e.a];
        if (D->mv_k < D->mv_n) {
            /* d.position.set(initialPos.x + xOffset, initialPos.y + yOffset); yield SimMan.timeout(MOVE_INTERVAL).
             * `initialPos = d.position` (tests/test_benchmark.py:77) is the Position OBJECT that moves, not a copy:
             * the offsets accumulate -- a random walk, not a jitter around the start position */
            const double x = D->x + D->mv_offsets[2 * D->mv_k], y = D->y + D->mv_offsets[2 * D->mv_k + 1];
            D->mv_k++;
            if (gwo_set_position(s, b, e.a, x, y) < 0 && !s->fault) s->fault = GWO_FAULT_INTERNAL;
            schedule(s, K_MOVE, PRIO_NORMAL, D->mv_interval, b, e.a, 0);
        }
        break;
    }
    default:
        s->fault = GWO_FAULT_INTERNAL;
    }
}

/* env.run(until=event): stops during the pop of the awaited event */
static int run_until_assign(gwo_sim *s, int b, uint32_t seq)
{
    Band *B = &s->band[b];
    if (B->assign_done_seq >= seq) return 0;       /* already processed */
    while (s->heap_n > 0) {
        Ev e = heap_pop(s);
        s->now = e.t;
        s->popped++;
        dispatch(s, e);
        if (s->fault) return s->fault;
        if (e.kind == K_ASSIGN_DONE && e.band == b && (uint32_t)e.a == seq) return 0;
    }
    s->fault = GWO_FAULT_EMPTY;
    return s->fault;
}

/* env.run(until=number): a StopSimulation event is scheduled URGENT at that time (simpy core.py) */
int gwo_run_for(gwo_sim *s, double duration)
{
    if (s->fault) return s->fault;
    if (!(duration > 0)) return -1;                 /* simpy: until must be > now */
    const uint64_t stop_eid = s->eid;
    schedule(s, K_STOP, PRIO_URGENT, duration, 0, 0, 0);
    while (s->heap_n > 0) {
        Ev e = heap_pop(s);
        s->now = e.t;
        s->popped++;
        if (e.kind == K_STOP && e.eid == stop_eid) return 0;
        dispatch(s, e);
        if (s->fault) return s->fault;
    }
    s->fault = GWO_FAULT_EMPTY;
    return s->fault;
}

int gwo_add_mover(gwo_sim *s, int band, int dev, double first_delay, double interval, const double *offsets,
                  int n_offsets)
{
    if (band < 0 || band >= s->nbands || dev < 0 || dev >= s->band[band].ndev) return -1;
    Dev *D = &s->band[band].dev[dev];
    D->mv_x0 = D->x; D->mv_y0 = D->y;
    D->mv_first = first_delay; D->mv_interval = interval;
    D->mv_offsets = offsets; D->mv_n = n_offsets; D->mv_k = 0;
    schedule(s, K_MOVE_INIT, PRIO_URGENT, 0, band, dev, 0);
    return 0;
}

/* ---------------------------------------------------------------------- */
/* public API                                                               */
/* ---------------------------------------------------------------------- */

static int py_str_len_int(long long v)             /* len(str(v)) for v >= 0 */
{
    int n = 1;
    while (v >= 10) { v /= 10; n++; }
    return n;
}

gwo_sim *gwo_create(const gwo_scenario *sc)
{
    if (sc->nbands < 1 || sc->nbands > GWO_MAXBAND) return NULL;
    gwo_sim *s = (gwo_sim *)calloc(1, sizeof(gwo_sim));
    s->nbands = sc->nbands;
    s->factor = sc->factor;
    s->mode = sc->mode;
    s->bitRate = 133.33333e3;                      /* physical.py:196 */
    s->codeRate = 0.75;                            /* float(Fraction(3, 4)) */
    s->dataRate = s->codeRate * s->bitRate;        /* physical.py:197 */
    s->maxBer = gwo_max_correctable_ber(3, 4);
    for (int b = 0; b < sc->nbands; b++) {
        const gwo_band_spec *bs = &sc->band[b];
        Band *B = &s->band[b];
        if (bs->ndev < 1 || bs->ndev > GWO_MAXDEV) { free(s); return NULL; }
        B->ndev = bs->ndev;
        B->frequency = bs->frequency;
        B->bandwidth = bs->bandwidth;
        B->thermal = gwo_thermal_noise_mw(bs->bandwidth);
        B->rrm = -1;
        for (int d = 0; d < bs->ndev; d++) {
            const gwo_dev_spec *ds = &bs->dev[d];
            Dev *D = &B->dev[d];
            D->role = ds->role;
            D->x = ds->x; D->y = ds->y;
            D->mult = ds->mult; D->payload_rule = ds->payload_rule; D->dest = ds->dest;
            D->interval = ds->interval;
            D->jam_interval = ds->jam_interval; D->jam_delay = ds->jam_delay;
            D->jam_power = ds->jam_power; D->jam_hdr = ds->jam_hdr; D->jam_payload = ds->jam_payload;
            D->max_ticks = ds->max_ticks; D->recv_mode = ds->receive;
            D->P = B->thermal;
            D->counter = 1;                        /* counter_traffic.py:48 */
            D->cur_tx = -1;
            if (ds->role == GWO_ROLE_RRM) B->rrm = d;
        }
        for (int i = 0; i < bs->ndev; i++)
            for (int j = 0; j < bs->ndev; j++)
                B->att[i][j] = (i == j) ? 0.0
                    : gwo_fspl(B->dev[i].x, B->dev[i].y, B->dev[j].x, B->dev[j].y, B->frequency);
    }
    /* process Initialize events in construction order: per band senders (device order),
     * then jammers (device order); all URGENT at t = 0 */
    for (int b = 0; b < sc->nbands; b++) {
        Band *B = &s->band[b];
        for (int d = 0; d < B->ndev; d++)
            if (B->dev[d].role == GWO_ROLE_SENDER)
                schedule(s, K_TICK, PRIO_URGENT, 0, b, d, 0);
        for (int d = 0; d < B->ndev; d++)
            if (B->dev[d].role == GWO_ROLE_JAMMER)
                schedule(s, K_JAM_WAKE, PRIO_URGENT, 0, b, d, 0);
    }
    /* receive mode is switched on after construction, band by band in device order (the harness does the same) */
    for (int b = 0; b < sc->nbands; b++) {
        Band *B = &s->band[b];
        for (int d = 0; d < B->ndev; d++)
            if (B->dev[d].role == GWO_ROLE_SENDER && B->dev[d].recv_mode)
                schedule(s, K_RECV_INIT, PRIO_URGENT, 0, b, d, 0);
    }
    return s;
}

void gwo_destroy(gwo_sim *s)
{
    if (!s) return;
    free(s->rec);
    free(s);
}

void gwo_set_trace(gwo_sim *s, int on) { s->trace_on = on; }

void gwo_set_mask_fn(gwo_sim *s, gwo_mask_fn fn, void *ctx, int64_t env_id)
{
    s->mask_fn = fn; s->mask_ctx = ctx; s->env_id = env_id;
}

/* ---------------------------------------------------------------------- */
/* mode M: built-in mask providers                                          */
/* ---------------------------------------------------------------------- */

/* Philox4x32-10 (Random123; Salmon et al., SC'11), restated from the published algorithm */
void gwo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* Error flag of on-air bit k: word (k & 3) of philox(ctr = (k >> 2, txseq,
 * sender | receiver << 8 | band << 16, env_lo), key = (seed_lo ^ env_hi, seed_hi))
 * is below floor(ber * 2^32).  (The keying is this project's definition of mode M.) */
static int64_t philox_mask_fn(void *ctx, int64_t env, int band, int sender, uint32_t seq,
                              int receiver, int64_t k0, int64_t k1, double ber)
{
    uint64_t seed = *(uint64_t *)ctx;
    uint32_t thr = (uint32_t)(ber * 4294967296.0);
    uint32_t key[2] = { (uint32_t)seed ^ (uint32_t)((uint64_t)env >> 32), (uint32_t)(seed >> 32) };
    int64_t n = 0, cur = -1;
    uint32_t w[4] = {0, 0, 0, 0};
    for (int64_t k = k0; k < k1; k++) {
        if ((k >> 2) != cur) {
            cur = k >> 2;
            uint32_t ctr[4] = { (uint32_t)cur, seq,
                                (uint32_t)sender | ((uint32_t)receiver << 8) | ((uint32_t)band << 16),
                                (uint32_t)(uint64_t)env };
            gwo_philox4x32_10(ctr, key, w);
        }
        n += w[k & 3] < thr;
    }
    return n;
}

typedef struct { const uint32_t *words; int slots, words_per_row, nbands; } FedCtx;

/* fed masks: [env][band][GWO_FED_DEV sender][slots][GWO_FED_DEV receiver][words_per_row] */
static int64_t fed_mask_fn(void *ctx, int64_t env, int band, int sender, uint32_t seq,
                           int receiver, int64_t k0, int64_t k1, double ber)
{
    const FedCtx *f = (const FedCtx *)ctx;
    int64_t row = ((((env * f->nbands + band) * GWO_FED_DEV + sender) * f->slots
                    + (int64_t)(seq % (uint32_t)f->slots)) * GWO_FED_DEV + receiver);
    const uint32_t *w = f->words + row * f->words_per_row;
    int64_t n = 0;
    (void)ber;
    for (int64_t k = k0; k < k1; k++) n += (w[k >> 5] >> (k & 31)) & 1u;
    return n;
}

static uint64_t g_seed_store[1024];
static FedCtx g_fed_store[1024];
static int g_store_next = 0;

void gwo_use_philox_masks(gwo_sim *s, uint64_t seed, int64_t env_id)
{
    int i = __sync_fetch_and_add(&g_store_next, 1) & 1023;
    g_seed_store[i] = seed;
    s->mask_fn = philox_mask_fn; s->mask_ctx = &g_seed_store[i]; s->env_id = env_id;
}

void gwo_use_fed_masks(gwo_sim *s, const uint32_t *words, int slots, int words_per_row, int64_t env_index)
{
    int i = __sync_fetch_and_add(&g_store_next, 1) & 1023;
    g_fed_store[i].words = words; g_fed_store[i].slots = slots;
    g_fed_store[i].words_per_row = words_per_row; g_fed_store[i].nbands = s->nbands;
    s->mask_fn = fed_mask_fn; s->mask_ctx = &g_fed_store[i]; s->env_id = env_index;
}

void gwo_reset(gwo_sim *s, int64_t *obs)           /* counter_traffic.py:135-144 */
{
    for (int b = 0; b < s->nbands; b++) {
        Band *B = &s->band[b];
        for (int d = 0; d < B->ndev; d++)
            if (B->dev[d].role == GWO_ROLE_SENDER) B->dev[d].counter = 0;
        B->latestDiff = 0; B->lastAbsDiff = 0; B->rv[0] = B->rv[1] = 0; B->done = 0;
        if (obs) obs[b] = B->latestDiff + COUNTER_BOUND;
    }
}

int gwo_step(gwo_sim *s, const int32_t *device, const int32_t *duration,
             int64_t *obs, double *reward, uint8_t *done)
{
    if (s->fault) return s->fault;
    uint32_t seq[GWO_MAXBAND];
    for (int b = 0; b < s->nbands; b++) {
        long long slots = (long long)duration[b] * s->factor;   /* counter_traffic.py:149 */
        rrm_assign(s, b, device[b], (double)slots, py_str_len_int(slots));
        seq[b] = s->band[b].assign_seq;
    }
    for (int b = 0; b < s->nbands; b++) {
        int rc = run_until_assign(s, b, seq[b]);                 /* counter_traffic.py:155 */
        if (rc) return rc;
    }
    for (int b = 0; b < s->nbands; b++) {                        /* envs/core.py:142-153 */
        Band *B = &s->band[b];
        obs[b] = B->latestDiff + COUNTER_BOUND;
        int absd = abs(B->latestDiff);
        int r = B->lastAbsDiff - absd;
        B->lastAbsDiff = absd;
        if (r > 10) r = 10; else if (r < -10) r = -10;
        reward[b] = (double)r;
        done[b] = (uint8_t)B->done;
    }
    return 0;
}

/* Position.set (devices/core.py:75-84) -> FsplAttenuation._positionChanged -> _update
 * (attenuation_models.py:28-39, physical.py:383-386).  Only valid while no transmission that
 * involves the device is on the air (the restatement does not model _onAttenuationChange). */
/* Position.set (devices/core.py:75-84) -> nChange -> every FsplAttenuation model of the device:
 * _positionChangedCallback (physical.py:383-386) -> _update (attenuation_models.py:28-36) ->
 * _setAttenuation (physical.py:354-362) -> nAttenuationChanges -> SimplePhy._onAttenuationChange
 * (simple_stack.py:119-128) of every PHY that registered a transmission on that model: the stored
 * received power of the transmission is replaced, the difference goes through
 * _nReceivedPowerChanges (power bookkeeping; a PHY that is receiving counts the errors of the
 * segment that ends and re-evaluates its bit error rate).
 * The models of one device are notified in Python-set order (simtools.py:255); they are visited by
 * ascending partner index here.  The order only matters for the rounding of the moving PHY's own
 * received power when two or more OTHER devices are transmitting at that instant. */
int gwo_set_position(gwo_sim *s, int band, int dev, double x, double y)
{
    Band *B = &s->band[band];
    if (x == B->dev[dev].x && y == B->dev[dev].y) return 0;      /* Position.set: no change, no trigger */
    B->dev[dev].x = x; B->dev[dev].y = y;
    for (int j = 0; j < B->ndev; j++) {
        if (j == dev) continue;
        double d = sqrt(pow(B->dev[dev].x - B->dev[j].x, 2.0) + pow(B->dev[dev].y - B->dev[j].y, 2.0));
        /* The model of a pair is created lazily, at the first transmission one of the two devices sends
         * (SimplePhy._getAttenuationModelByTransmission -> FrequencyBand.getAttenuationModel,
         * physical.py:576-594); until then nobody listens to position changes, and the model will be
         * computed from the positions of THAT moment -- without the threshold, and with 0 dB for
         * coinciding devices (AttenuationModel.__init__ + FsplAttenuation._update). */
        if (B->dev[dev].tx_seq == 0 && B->dev[j].tx_seq == 0) {
            double fresh = 0.0;
            if (!(B->dev[dev].x == B->dev[j].x && B->dev[dev].y == B->dev[j].y))
                fresh = 20 * log10(d) + 20 * log10(B->frequency) - 147.55;
            B->att[dev][j] = fresh; B->att[j][dev] = fresh;
            continue;
        }
        if (!(d < 3000.0)) continue;                              /* STANDBY_THRESHOLD, physical.py:371 */
        if (B->dev[dev].x == B->dev[j].x && B->dev[dev].y == B->dev[j].y) continue;   /* _update returns early */
        /* devices[0] / devices[1] order of the model is the frozenset order: the formula is symmetric */
        double att = 20 * log10(d) + 20 * log10(B->frequency) - 147.55;
        if (att == B->att[dev][j]) continue;                      /* _setAttenuation: only a new value triggers */
        B->att[dev][j] = att; B->att[j][dev] = att;
        /* transmissions registered on this model: sent by one of the two devices, received by the other */
        for (int ti = 0; ti < GWO_MAXTX; ti++) {
            Tx *t = &B->tx[ti];
            if (!t->used) continue;
            int p = -1;
            if (t->sender == dev) p = j; else if (t->sender == j) p = dev;
            if (p < 0) continue;
            Dev *P = &B->dev[p];
            if (!P->hasS[ti]) continue;
            double rp = dbm_to_mw(t->power - att);
            double delta = rp - P->S[ti];
            P->S[ti] = rp;
            power_change(s, band, p, delta);
            if (s->fault) return -2;
        }
    }
    return 0;
}

double gwo_now(const gwo_sim *s) { return s->now; }
int64_t gwo_popped(const gwo_sim *s) { return s->popped; }
int gwo_fault(const gwo_sim *s) { return s->fault; }
int64_t gwo_near_ties(const gwo_sim *s) { return s->near_round_tie + s->near_ber_tie; }

void gwo_counts(const gwo_sim *s, int band, int64_t *n_tx, int64_t *n_deliv /* [GWO_MAXDEV] */)
{
    const Band *B = &s->band[band];
    *n_tx = B->n_tx;
    for (int d = 0; d < GWO_MAXDEV; d++) n_deliv[d] = B->n_deliv[d];
}

void gwo_received(const gwo_sim *s, int band, int64_t *n_received)
{
    for (int d = 0; d < GWO_MAXDEV; d++) n_received[d] = s->band[band].dev[d].n_received;
}

double gwo_attenuation(const gwo_sim *s, int band, int i, int j) { return s->band[band].att[i][j]; }

size_t gwo_trace_take(gwo_sim *s, double *out, size_t cap_doubles)
{
    size_t n = s->rec_n < cap_doubles ? s->rec_n : cap_doubles;
    if (out && n) memcpy(out, s->rec, n * sizeof(double));
    size_t total = s->rec_n;
    s->rec_n = 0;
    return total;
}

size_t gwo_trace_size(const gwo_sim *s) { return s->rec_n; }

void gwo_default_scenario(gwo_scenario *sc)        /* counter_traffic.py:114-133 */
{
    memset(sc, 0, sizeof *sc);
    sc->nbands = 1;
    sc->factor = 1000;                             /* envs/core.py:27 */
    sc->mode = GWO_MODE_R;
    gwo_band_spec *b = &sc->band[0];
    b->ndev = 3;
    b->frequency = 2.4e9; b->bandwidth = 22e6;     /* physical.py:298 */
    b->dev[0].role = GWO_ROLE_SENDER; b->dev[0].x = 0; b->dev[0].y = 2;
    b->dev[0].mult = 1; b->dev[0].payload_rule = -1; b->dev[0].dest = 1; b->dev[0].interval = 0.001;
    b->dev[1].role = GWO_ROLE_SENDER; b->dev[1].x = 0; b->dev[1].y = -2;
    b->dev[1].mult = 3; b->dev[1].payload_rule = -1; b->dev[1].dest = 0; b->dev[1].interval = 0.001;
    b->dev[2].role = GWO_ROLE_RRM; b->dev[2].x = 0; b->dev[2].y = 0;
}

/*
 * Batch runner (used by the parity tests and as the CPU baseline): `nenv`
 * independent envs of the same scenario, optional per-env positions
 * pos[env][band][dev][2], action tapes dev_tape/dur_tape[step][env][band].
 * Outputs (any may be NULL): obs/reward/done [step][env][band], now[step][env],
 * counts[env][band][1 + GWO_BATCH_DEV] (n_tx, deliveries per device) after the last step.
 * Envs [env_begin, env_end) are processed -- the caller shards threads.
 */
int gwo_run_batch_m(const gwo_scenario *sc, int64_t nenv, int nsteps, int do_reset,
                    const double *pos, const int32_t *dev_tape, const int32_t *dur_tape,
                    int64_t *obs, double *reward, uint8_t *done, double *now, int64_t *counts,
                    int64_t env_begin, int64_t env_end,
                    uint64_t seed, int64_t env_id_offset, const uint32_t *fed_words, int fed_slots,
                    int fed_words_per_row);

/* Timing support for benchmarks (bench.py): the seconds THIS thread spends in steps t >= time_from of the
 * envs it simulates are accumulated (the envs are simulated one after the other, so the steps before
 * time_from -- a burn-in -- are excluded per env). */
static __thread int g_time_from = -1;
static __thread double g_timed_seconds = 0.0;

static double mono_seconds(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int gwo_run_batch(const gwo_scenario *sc, int64_t nenv, int nsteps, int do_reset,
                  const double *pos, const int32_t *dev_tape, const int32_t *dur_tape,
                  int64_t *obs, double *reward, uint8_t *done, double *now, int64_t *counts,
                  int64_t env_begin, int64_t env_end)
{
    return gwo_run_batch_m(sc, nenv, nsteps, do_reset, pos, dev_tape, dur_tape, obs, reward, done, now,
                           counts, env_begin, env_end, 0, 0, NULL, 0, 0);
}

int gwo_run_batch_m(const gwo_scenario *sc, int64_t nenv, int nsteps, int do_reset,
                    const double *pos, const int32_t *dev_tape, const int32_t *dur_tape,
                    int64_t *obs, double *reward, uint8_t *done, double *now, int64_t *counts,
                    int64_t env_begin, int64_t env_end,
                    uint64_t seed, int64_t env_id_offset, const uint32_t *fed_words, int fed_slots,
                    int fed_words_per_row)
{
    int nb = sc->nbands;
    int rc_all = 0;
    for (int64_t e = env_begin; e < env_end; e++) {
        gwo_scenario local = *sc;
        if (pos) {
            for (int b = 0; b < nb; b++)
                for (int d = 0; d < sc->band[b].ndev; d++) {
                    const double *p = pos + (((size_t)e * nb + b) * GWO_BATCH_DEV + d) * 2;
                    local.band[b].dev[d].x = p[0];
                    local.band[b].dev[d].y = p[1];
                }
        }
        gwo_sim *s = gwo_create(&local);
        if (!s) return GWO_FAULT_INTERNAL;
        if (sc->mode == GWO_MODE_M) {
            uint64_t seed_local = seed;
            FedCtx fed_local;
            if (fed_words) {
                fed_local.words = fed_words; fed_local.slots = fed_slots;
                fed_local.words_per_row = fed_words_per_row; fed_local.nbands = nb;
                s->mask_fn = fed_mask_fn; s->mask_ctx = &fed_local; s->env_id = e;
            } else {
                s->mask_fn = philox_mask_fn; s->mask_ctx = &seed_local; s->env_id = env_id_offset + e;
            }
            /* the contexts live on this stack frame for the lifetime of `s` (destroyed below) */
            if (do_reset) gwo_reset(s, NULL);
            int rc_m = 0;
            int64_t om[GWO_MAXBAND]; double rm[GWO_MAXBAND]; uint8_t dm[GWO_MAXBAND];
            for (int t = 0; t < nsteps; t++) {
                size_t base = ((size_t)t * nenv + e) * nb;
                rc_m = gwo_step(s, dev_tape + base, dur_tape + base, om, rm, dm);
                if (rc_m) break;
                for (int b = 0; b < nb; b++) {
                    if (obs) obs[base + b] = om[b];
                    if (reward) reward[base + b] = rm[b];
                    if (done) done[base + b] = dm[b];
                }
                if (now) now[(size_t)t * nenv + e] = s->now;
            }
            if (counts) {
                for (int b = 0; b < nb; b++) {
                    int64_t *c = counts + ((size_t)e * nb + b) * (1 + GWO_BATCH_DEV);
                    { int64_t nd_[GWO_MAXDEV]; gwo_counts(s, b, c, nd_); for (int d_ = 0; d_ < GWO_BATCH_DEV; d_++) c[1 + d_] = nd_[d_]; }
                }
            }
            gwo_destroy(s);
            if (rc_m) return rc_m;
            continue;
        }
        if (do_reset) gwo_reset(s, NULL);
        int64_t o[GWO_MAXBAND]; double r[GWO_MAXBAND]; uint8_t dn[GWO_MAXBAND];
        double t_begin = 0.0;
        for (int t = 0; t < nsteps; t++) {
            if (t == g_time_from) t_begin = mono_seconds();
            size_t base = ((size_t)t * nenv + e) * nb;
            int rc = gwo_step(s, dev_tape + base, dur_tape + base, o, r, dn);
            if (rc) { rc_all = rc; break; }
            for (int b = 0; b < nb; b++) {
                if (obs) obs[base + b] = o[b];
                if (reward) reward[base + b] = r[b];
                if (done) done[base + b] = dn[b];
            }
            if (now) now[(size_t)t * nenv + e] = s->now;
        }
        if (g_time_from >= 0 && g_time_from < nsteps && !rc_all) g_timed_seconds += mono_seconds() - t_begin;
        if (counts) {
            for (int b = 0; b < nb; b++) {
                int64_t *c = counts + ((size_t)e * nb + b) * (1 + GWO_BATCH_DEV);
                { int64_t nd_[GWO_MAXDEV]; gwo_counts(s, b, c, nd_); for (int d_ = 0; d_ < GWO_BATCH_DEV; d_++) c[1 + d_] = nd_[d_]; }
            }
        }
        gwo_destroy(s);
        if (rc_all) return rc_all;
    }
    return 0;
}

/* gwo_run_batch (mode R) that also reports the seconds this thread spent in steps t >= time_from */
int gwo_run_batch_timed(const gwo_scenario *sc, int64_t nenv, int nsteps, int do_reset,
                        const double *pos, const int32_t *dev_tape, const int32_t *dur_tape,
                        int64_t *obs, double *reward, uint8_t *done, double *now, int64_t *counts,
                        int64_t env_begin, int64_t env_end, int time_from, double *seconds_out)
{
    g_time_from = time_from;
    g_timed_seconds = 0.0;
    int rc = gwo_run_batch_m(sc, nenv, nsteps, do_reset, pos, dev_tape, dur_tape, obs, reward, done, now, counts,
                             env_begin, env_end, 0, 0, NULL, 0, 0);
    if (seconds_out) *seconds_out = g_timed_seconds;
    g_time_from = -1;
    return rc;
}
