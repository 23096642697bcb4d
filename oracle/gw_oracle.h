/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY (see gw_oracle.c).  Interface of the plain-C
 * restatement of the reference's CounterTrafficEnv hot path.
 */
#ifndef GW_ORACLE_H
#define GW_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GWO_MAXDEV 28              /* devices per band held by the model (grids of PHY-only senders: up to 24; bands: 8 senders + RRM + 16 PHY-only senders) */
#define GWO_BATCH_DEV 8            /* device stride of the batch API's pos / counts arrays (kept from round 1) */
#define GWO_MAXBAND 4
#define GWO_MAXTX 28

#define GWO_ROLE_SENDER 1   /* SimpleNetworkDevice + traffic process (counter_traffic.py:37-61) */
#define GWO_ROLE_RRM 2      /* SimpleRrmDevice (devices.py:113) */
#define GWO_ROLE_JAMMER 3   /* PHY-only periodic sender (tests/test_benchmark.py:20-50) */

#define GWO_MODE_R 0        /* reference-exact expected-value accounting (incl. quirks) */
#define GWO_MODE_M 1        /* per-bit error masks */

#define GWO_REC_TX 1        /* t=start, dev=sender, x0=stop, x1=headerBits, x2=payloadBits */
#define GWO_REC_BER 2       /* t, dev=receiver, x0=BER */
#define GWO_REC_DEC 3       /* t, dev=receiver, x0=section(0 header,1 payload), x1=errSum, x2=totalBits, x3=ok */
#define GWO_REC_RX 4        /* t, dev=sender index of a packet the RRM decoded */
#define GWO_REC_MRX 5       /* t, dev=device whose MAC handed a received packet to onReceive (receive mode) */

#define GWO_FAULT_HEAP 1
#define GWO_FAULT_REF_KEYERROR 2   /* the reference would raise KeyError */
#define GWO_FAULT_REF_ASSERT 3     /* the reference would fail an assert (simple_stack.py:168-169) */
#define GWO_FAULT_SENDQ 4
#define GWO_FAULT_TXPOOL 5
#define GWO_FAULT_EMPTY 6
#define GWO_FAULT_INTERNAL 7

typedef struct {
    int32_t role;
    double x, y;
    /* sender */
    int32_t mult;           /* packets per tick */
    int32_t payload_rule;   /* -1: byteSize = counter (reference), else fixed byteSize */
    int32_t dest;           /* destination device index */
    double interval;        /* COUNTER_INTERVAL */
    int32_t max_ticks;      /* 0: the traffic process runs forever (reference); n: a burst of n ticks */
    int32_t receive;        /* 1: MAC receive mode (SimpleNetworkDevice.receiving = True, devices.py:70-97) */
    /* jammer */
    double jam_interval, jam_delay, jam_power;
    int32_t jam_hdr, jam_payload;
} gwo_dev_spec;

typedef struct {
    int32_t ndev;
    double frequency, bandwidth;
    gwo_dev_spec dev[GWO_MAXDEV];   /* canonical order: senders, rrm, jammers */
} gwo_band_spec;

typedef struct {
    int32_t nbands;
    int32_t factor;         /* ASSIGNMENT_DURATION_FACTOR */
    int32_t mode;
    gwo_band_spec band[GWO_MAXBAND];
} gwo_scenario;

typedef struct gwo_sim gwo_sim;

/* mode M: number of bit errors among on-air bits [k0, k1) of transmission `seq` of
 * device `sender` as seen by `receiver`, the segment's BER being `ber` */
typedef int64_t (*gwo_mask_fn)(void *ctx, int64_t env, int band, int sender, uint32_t seq,
                               int receiver, int64_t k0, int64_t k1, double ber);

gwo_sim *gwo_create(const gwo_scenario *sc);
void gwo_destroy(gwo_sim *s);
void gwo_default_scenario(gwo_scenario *sc);
void gwo_set_trace(gwo_sim *s, int on);
void gwo_set_mask_fn(gwo_sim *s, gwo_mask_fn fn, void *ctx, int64_t env_id);
void gwo_reset(gwo_sim *s, int64_t *obs);
int gwo_step(gwo_sim *s, const int32_t *device, const int32_t *duration,
             int64_t *obs, double *reward, uint8_t *done);
int gwo_set_position(gwo_sim *s, int band, int dev, double x, double y);
/* SimMan.runSimulation(duration) (simtools.py:77-88 -> simpy env.run(until = now + duration)): every event
 * strictly before that time and the URGENT ones at it; the clock ends at now + duration. */
int gwo_run_for(gwo_sim *s, double duration);
/* A mobility process of device `dev` as in tests/test_benchmark.py:73-85: created now (its Initialize event is
 * URGENT at the current time), waits `first_delay`, then every `interval` moves the device by
 * offsets[2*k .. 2*k+1] for k = 0, 1, ... -- the reference's `initialPos` is the moving Position object itself
 * (:77), so the offsets ACCUMULATE (a random walk); stops after n_offsets jumps.
 * The offsets array must stay alive. */
int gwo_add_mover(gwo_sim *s, int band, int dev, double first_delay, double interval, const double *offsets,
                  int n_offsets);
double gwo_now(const gwo_sim *s);
int64_t gwo_popped(const gwo_sim *s);
int gwo_fault(const gwo_sim *s);
int64_t gwo_near_ties(const gwo_sim *s);
void gwo_counts(const gwo_sim *s, int band, int64_t *n_tx, int64_t *n_deliv);
/* packets handed to SimpleNetworkDevice.onReceive per device (receive mode) */
void gwo_received(const gwo_sim *s, int band, int64_t *n_received /* [GWO_MAXDEV] */);
double gwo_attenuation(const gwo_sim *s, int band, int i, int j);
size_t gwo_trace_take(gwo_sim *s, double *out, size_t cap_doubles);
size_t gwo_trace_size(const gwo_sim *s);

int gwo_run_batch(const gwo_scenario *sc, int64_t nenv, int nsteps, int do_reset,
                  const double *pos, const int32_t *dev_tape, const int32_t *dur_tape,
                  int64_t *obs, double *reward, uint8_t *done, double *now, int64_t *counts,
                  int64_t env_begin, int64_t env_end);

/* mode M mask providers (GWO_FED_DEV = device stride of the fed-mask layout) */
#define GWO_FED_DEV 4
void gwo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void gwo_use_philox_masks(gwo_sim *s, uint64_t seed, int64_t env_id);
void gwo_use_fed_masks(gwo_sim *s, const uint32_t *words, int slots, int words_per_row, int64_t env_index);
/* gwo_run_batch (mode R) that also reports the seconds THIS thread spent in steps t >= time_from of the
 * envs it simulated (benchmarks: a burn-in is excluded per env). */
int gwo_run_batch_timed(const gwo_scenario *sc, int64_t nenv, int nsteps, int do_reset,
                        const double *pos, const int32_t *dev_tape, const int32_t *dur_tape,
                        int64_t *obs, double *reward, uint8_t *done, double *now, int64_t *counts,
                        int64_t env_begin, int64_t env_end, int time_from, double *seconds_out);

int gwo_run_batch_m(const gwo_scenario *sc, int64_t nenv, int nsteps, int do_reset,
                    const double *pos, const int32_t *dev_tape, const int32_t *dur_tape,
                    int64_t *obs, double *reward, uint8_t *done, double *now, int64_t *counts,
                    int64_t env_begin, int64_t env_end,
                    uint64_t seed, int64_t env_id_offset, const uint32_t *fed_words, int fed_slots,
                    int fed_words_per_row);

/* arithmetic helpers, exported for the numeric parity tests */
double gwo_q_function(double x);
double gwo_ber_bpsk(double s_dbm, double n_dbm, double bitRate);
double gwo_fspl(double ax, double ay, double bx, double by, double frequency);
double gwo_thermal_noise_mw(double bandwidth);
double gwo_max_correctable_ber(int k, int n);

#ifdef __cplusplus
}
#endif
#endif
