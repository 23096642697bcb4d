/*
 * gymwipe_b200 -- C ABI of the B200-native batched simulator for Gym-WiPE's per-step
 * wireless hot path (CounterTrafficEnv.step and the networking stack below it).
 *
 * The reference (Gryph66/gymwipe) is 100 % Python and has NO plugin / FFI layer: the
 * operator API of this path is the gym `Env` object itself.  Each entry point below
 * therefore cites the reference interface it replaces (file:line under
 * /root/reference); INTEGRATION.md shows the ctypes binding a maintainer would add to
 * gymwipe/envs/counter_traffic.py.
 *
 * Conventions
 *  - plain C, no torch / C++ types; all device buffers are raw CUDA device pointers
 *    owned by the caller (e.g. torch.Tensor.data_ptr()), contiguous, on the handle's
 *    device; `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *  - every function returns 0 on success or a negative GW_E_* code; the message is
 *    available from gw_last_error() (thread-local).  Nothing throws across the ABI.
 *  - no allocation and no host synchronisation inside gw_step(); work is enqueued on the
 *    caller's stream.  A handle is bound to one device and driven by one host thread at
 *    a time; several handles (one per GPU / process) are independent.
 *  - there is NO CPU fallback: without a CUDA device every compute entry point fails
 *    with GW_E_CUDA.
 */
#ifndef GYMWIPE_B200_H
#define GYMWIPE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GW_ABI_VERSION 2

#define GW_MAX_BANDS 4
#define GW_MAX_DEVICES 4            /* per frequency band */
#define GW_MAX_SENDERS 2
#define GW_MAX_JAMMERS 1

/* error codes */
#define GW_OK 0
#define GW_E_INVALID (-1)           /* bad argument / unsupported scenario */
#define GW_E_CUDA (-2)              /* CUDA runtime error (message has the detail) */
#define GW_E_STATE (-3)             /* state buffer too small / misaligned */
#define GW_E_ACTION (-4)            /* an action was outside the action space (device flag) */
#define GW_E_SIMFAULT (-5)          /* an env hit a condition under which the reference raises */

/* error-accounting modes */
#define GW_MODE_REFERENCE 0         /* "mode R": the reference's expected-value accounting, quirks included */
#define GW_MODE_MASK_PHILOX 1       /* "mode M": per-bit Philox4x32-10 error masks generated in the kernel */
#define GW_MODE_MASK_FED 2          /* "mode M": per-bit error masks read from HBM (gw_set_masks) */

/* device roles on a band; canonical device order is senders, RRM, jammers */
#define GW_ROLE_SENDER 1            /* SimpleNetworkDevice + traffic process, counter_traffic.py:37-61 */
#define GW_ROLE_RRM 2               /* SimpleRrmDevice, networking/devices.py:113-203 */
#define GW_ROLE_JAMMER 3            /* PHY-only periodic sender, tests/test_benchmark.py:20-50 */

typedef struct {
    int32_t role;
    double x, y;                    /* default position (m), devices/core.py:15-98 */
    /* GW_ROLE_SENDER */
    int32_t multiplicity;           /* packets per tick, counter_traffic.py:44 */
    int32_t payload_bytes;          /* -1: byteSize = counter (reference behaviour, app. B #1); >=0 fixed */
    double interval;                /* COUNTER_INTERVAL, counter_traffic.py:31 */
    int32_t max_ticks;              /* 0: the traffic process runs forever (reference); n: a burst of n ticks (finite
                                       sources as in tests/networking/test_stack.py:186-191) */
    int32_t receive;                /* 1: MAC receive mode -- SimpleNetworkDevice.receiving = True (devices.py:70-97):
                                       the device loops RECEIVE commands (RECEIVE_TIMEOUT 100 s), data packets addressed
                                       to its idle MAC are handed to onReceive (simple_stack.py:436-460); GW_FIELD_N_RECEIVED */
    /* GW_ROLE_JAMMER */
    double jam_interval, jam_delay, jam_power_dbm;
    int32_t jam_header_bytes, jam_payload_bytes;
} gw_device_config;

typedef struct {
    int32_t n_devices;
    double frequency_hz, bandwidth_hz;   /* FrequencyBandSpec, physical.py:293-306 */
    gw_device_config device[GW_MAX_DEVICES];
} gw_band_config;

/* optional plant coupled to the network (config 5: networked inverted pendulum) */
#define GW_PLANT_NONE 0
#define GW_PLANT_SLIDING_PENDULUM 1  /* gymwipe/plants/sliding_pendulum.py:15-114 */

typedef struct {
    double cart_mass, pendulum_mass, arm_length, gravity;   /* sliding_pendulum.py:24-45, plants/core.py:33-35 */
    double motor_fmax, motor_kservo, motor_v_init;          /* slider ParamFMax = 22, ParamVel = 0.1 (:52-53) */
    double dt_max;                                          /* largest integrator sub-step (s) */
    double kp, ki, kd;                                      /* control/inverted_pendulum.py:46-50 */
    int32_t mobility;               /* sensor / actuator x follow the wagon (sliding_pendulum.py:132,149) */
} gw_pendulum_config;

typedef struct {
    int32_t abi_version;            /* GW_ABI_VERSION */
    int64_t n_envs;                 /* envs held by this handle (this GPU's shard) */
    int64_t env_id_offset;          /* global id of env 0: RNG keys use global ids (sharding-invariant) */
    int32_t n_bands;
    int32_t assignment_duration_factor;  /* BaseEnv.ASSIGNMENT_DURATION_FACTOR, envs/core.py:27 */
    int32_t max_assign_duration;         /* BaseEnv.MAX_ASSIGN_DURATION, envs/core.py:25 */
    int32_t mode;                   /* GW_MODE_* */
    uint64_t seed;                  /* Philox key (mode M) */
    int32_t per_env_positions;      /* 0: all envs share the scenario's positions */
    gw_band_config band[GW_MAX_BANDS];
    int32_t plant;                  /* GW_PLANT_*; a plant env has one band with the devices
                                       sensor (sender 0), controller (sender 1), RRM, actuator (receive-only) */
    gw_pendulum_config pendulum;
} gw_config;

typedef struct gw_handle gw_handle;

/* Library / device probing. */
int gw_abi_version(void);
const char *gw_last_error(void);
int gw_device_count(int *count);

/* Fills `cfg` with CounterTrafficEnv's scenario: two senders at (0,+-2) m with
 * multiplicity 1 and 3, the RRM at (0,0), one 2.4 GHz / 22 MHz band
 * (gymwipe/envs/counter_traffic.py:114-133). */
int gw_default_config(gw_config *cfg, int64_t n_envs);

/* Bytes of device memory the env-batch state needs for `cfg`. */
int gw_state_bytes(const gw_config *cfg, size_t *bytes);

/* Replaces CounterTrafficEnv.__init__ (counter_traffic.py:114-133) for a batch of envs.
 * `state` is a caller-owned device buffer of at least gw_state_bytes() bytes, 256-byte
 * aligned (e.g. a torch uint8 tensor), or NULL to let the handle allocate its own.
 * The construction-time state of the reference (time 0, counters 1, empty queues,
 * thermal-noise-only received power) is written on `stream`. */
int gw_create(const gw_config *cfg, int device, void *state, size_t state_bytes, void *stream,
              gw_handle **out);
void gw_destroy(gw_handle *h);

/* Per-env device positions, float64 [n_envs][n_bands][GW_MAX_DEVICES][2] on the device
 * (Device / Position, devices/core.py).
 *  - Before the first step: the devices are CREATED there -- the FSPL attenuation and received-
 *    power tables are computed from scratch (kernel K1; attenuation_models.py:28-36,
 *    simple_stack.py:99-111).  With `positions == NULL` the scenario's default positions are applied.
 *  - Afterwards: the devices MOVE, one after the other by ascending index like successive
 *    Position.set calls (devices/core.py:75-84) between two env.step calls: models beyond
 *    STANDBY_THRESHOLD or with coinciding devices keep their attenuation (physical.py:383-386,
 *    attenuation_models.py:31-33), and transmissions that are on the air at that instant see
 *    SimplePhy._onAttenuationChange (simple_stack.py:119-128): the receivers' power sums change, a
 *    running reception counts the errors of the segment that ends and re-evaluates its bit error
 *    rate (all three accounting modes).  Plant envs re-create the tables instead. */
int gw_set_positions(gw_handle *h, const double *positions, void *stream);

/* Replaces CounterTrafficEnv.reset (counter_traffic.py:135-144): sender counters := 0,
 * interpreter reset; simulated time, queues and PHY state are NOT touched (app. B #10).
 * `env_ids` (device int64[n]) selects envs, NULL = all.  `obs` (device int64
 * [n_envs][n_bands]) may be NULL. */
int gw_reset(gw_handle *h, const int64_t *env_ids, int64_t n, int64_t *obs, void *stream);

/* Plant envs (cfg.plant != 0) use the same entry points: gw_step replaces
 * InvertedPendulumEnv.step (gymwipe/envs/inverted_pendulum.py:79-91) -- obs = int(degrees(angle)),
 * reward = |180 - degrees(angle)| (:42-50) -- and integrates the plant inside the step kernel. */

/* Replaces CounterTrafficEnv.step (counter_traffic.py:146-158): one RRM assignment cycle
 * for every env.  `device`, `duration`: device int32 [n_envs][n_bands] (action["device"],
 * action["duration"]).  Outputs (device): obs int64, reward float64, done uint8, each
 * [n_envs][n_bands].  Out-of-space actions set the handle's error flag (gw_check) and are
 * clamped -- the reference asserts (counter_traffic.py:147). */
int gw_step(gw_handle *h, const int32_t *device, const int32_t *duration,
            int64_t *obs, double *reward, uint8_t *done, void *stream);

/* gw_step with an event trace (debugging / parity checks; mode R): every band-sim appends
 * records of 8 float64 {kind, time, device, x0, x1, x2, x3, 0} to trace[sim][cap][8] and its record
 * count to trace_count[sim] (counts above `cap` mean truncation):
 *   kind 1 transmission start (Transmission, physical.py:224-279): device = sender, x0 = stopTime,
 *          x1 = headerBits, x2 = payloadBits
 *   kind 2 SimplePhy._updateBitErrorRate (simple_stack.py:161-173): device = receiver, x0 = BER
 *   kind 3 SimplePhy._decide (simple_stack.py:269-286): device = receiver, x0 = section (0 header,
 *          1 payload), x1 = bit error sum, x2 = total bits, x3 = verdict
 *   kind 4 interpreter.onPacketReceived (devices.py:163-168): device = sender index
 *   kind 5 SimpleNetworkDevice.onReceive (devices.py:88-97, MAC receive mode): device = receiving device */
int gw_step_traced(gw_handle *h, const int32_t *device, const int32_t *duration,
                   int64_t *obs, double *reward, uint8_t *done,
                   double *trace, int32_t *trace_count, int32_t cap, void *stream);

/* Same call with HOST buffers (pinned or pageable): copies the actions in, steps, copies
 * obs / reward / done out and synchronises the stream.  This is the end-to-end path a
 * gym-style caller with host-resident actions uses. */
int gw_step_host(gw_handle *h, const int32_t *device, const int32_t *duration,
                 int64_t *obs, double *reward, uint8_t *done, void *stream);

/* Packed host-buffer variant for throughput-bound callers: ONE host-to-device copy of
 * `actions` = int32 [2][n_sims] (row 0 action["device"], row 1 action["duration"]) and ONE
 * device-to-host copy of `results` = { int32 obs[n_sims]; float reward[n_sims]; uint8 done[n_sims] }
 * laid out back to back (9 bytes per sim; obs and reward are small integers, exactly representable).
 * Synchronises the stream. */
int gw_step_host_packed(gw_handle *h, const int32_t *actions, void *results, void *stream);

/* Compact host-buffer variant (the transfers, not the kernel, bound the end-to-end rate):
 * `actions` = uint8 [n_sims][2] -- per sim the byte pair { action["device"], action["duration"] }, i.e.
 * actions[2*i] = device and actions[2*i + 1] = duration of sim i (both < 256 by the action space:
 * Discrete(2) x Discrete(MAX_ASSIGN_DURATION), envs/core.py:39-42) -- and `results` =
 * uint32 [n_sims], one word per sim:
 *     bits  0..16  observation (Discrete(131072), counter_traffic.py:120)
 *     bits 17..21  reward + 16  (rewards are integers in [-10, 10], counter_traffic.py:96-107)
 *     bit  22      done
 * (GW_COMPACT_OBS / GW_COMPACT_REWARD / GW_COMPACT_DONE decode it.)  2 bytes in, 4 bytes out per sim
 * and step.  If both buffers are PINNED host memory (cudaHostAlloc / cudaHostRegister, e.g. torch's
 * pin_memory()) the step kernel reads and writes them in place over the host link and no copy is
 * enqueued; pageable buffers are staged through two copies.  Synchronises `stream`.  Not available
 * for plant envs (their observations are not integers of this range). */
int gw_step_host_compact(gw_handle *h, const uint8_t *actions, uint32_t *results, void *stream);
/* Asynchronous form for callers that keep several env batches in flight (one handle each): enqueues
 * the step on `stream` and returns; `results` is valid once the stream (or an event recorded after the
 * call) has completed.  PINNED host buffers only (GW_E_INVALID otherwise). */
int gw_step_host_compact_async(gw_handle *h, const uint8_t *actions, uint32_t *results, void *stream);
/* One env.step of a POPULATION of env batches held by several handles on one device (a vector env whose
 * state is split into independently allocated batches): the steps of all handles are enqueued back to
 * back on `stream` -- each reads its own actions[k] / writes its own results[k], layouts as above, PINNED
 * host buffers -- and the call synchronises ONCE.  The launch and wake-up latency of a synchronous step
 * is paid per population, not per batch; internally the (independent) batches are spread over a few
 * side streams forked from / joined into `stream`, so that one batch's result words cross the host link
 * while the next batch computes.  From the caller's point of view everything is ordered on `stream`. */
int gw_step_host_compact_many(gw_handle *const *handles, int32_t n_handles, const uint8_t *const *actions,
                              uint32_t *const *results, void *stream);
/* The smallest wire format of a CounterTrafficEnv step, for callers bound by the host link (several GPUs of one
 * box stepping from host-resident actors share it): `actions` = uint8 [n_sims], one byte per sim = device << 7 |
 * duration (Discrete(2) x Discrete(MAX_ASSIGN_DURATION <= 128), envs/core.py:39-42); `results` = uint16 [n_sims]:
 *     bits  0..7   observation - COUNTER_BOUND as a signed byte: the interpreter's latestDifference
 *                  (counter_traffic.py:85-94: receivedValues[0] - receivedValues[1], each 0 or COUNTER_BYTE_LENGTH)
 *     bits  8..12  reward + 16
 *     bit  13      done
 *     bit  15      the difference did not fit a signed byte (cannot happen with the reference's value 2)
 * 1 byte in, 2 bytes out per sim and step; same in-place access of pinned buffers, same staging of pageable
 * ones, same restrictions as gw_step_host_compact; gw_step_host_tiny_many = gw_step_host_compact_many with
 * these layouts. */
int gw_step_host_tiny(gw_handle *h, const uint8_t *actions, uint16_t *results, void *stream);
int gw_step_host_tiny_many(gw_handle *const *handles, int32_t n_handles, const uint8_t *const *actions,
                           uint16_t *const *results, void *stream);
#define GW_TINY_ACTION(device, duration) ((uint8_t)(((device) << 7) | (duration)))
#define GW_TINY_OBS(w)    ((int32_t)(int8_t)((w) & 0xFFu) + 65536)
#define GW_TINY_REWARD(w) ((int32_t)(((w) >> 8) & 31u) - 16)
#define GW_TINY_DONE(w)   ((int32_t)(((w) >> 13) & 1u))
#define GW_COMPACT_OBS(w)    ((int32_t)((w) & 0x1FFFFu))
#define GW_COMPACT_REWARD(w) ((int32_t)(((w) >> 17) & 31u) - 16)
#define GW_COMPACT_DONE(w)   ((int32_t)(((w) >> 22) & 1u))

/* Synchronises and reports the error flag the kernels raised since the last call:
 * 0, GW_E_ACTION or GW_E_SIMFAULT (with the first faulting env in the message). */
int gw_check(gw_handle *h, void *stream);

/* Per-step statistics reduced over this handle's envs by the step kernel's epilogue
 * (K5): out[0]=sum reward, [1]=deliveries of sender 0, [2]=of sender 1, [3]=sum done,
 * [4]=env-steps, [5]=sum |latest difference|, [6]=transmissions, [7]=exact-time ties
 * between independent events (diagnostic, expected 0).  `out` is device float64[8]; the
 * accumulators are cleared after the copy when `clear` != 0.  Feeds the learner
 * (agents/dqn_counter_traffic.py:70) -- and the NCCL all-reduce when envs are sharded. */
int gw_stats(gw_handle *h, double *out8, int clear, void *stream);

/* Mode M with fed masks: bytes of mask words the step kernels have read for their error counts since
 * the last clearing call -- per decided section the 32-bit words that hold its on-air bits
 * (SimplePhy._countBitErrors with per-bit masks, simple_stack.py:180-188).  This is the algorithmic
 * traffic of the HBM-bound part of the step (bench.py's roofline of configs[2]).  `out` is a HOST
 * uint64; synchronises `stream`. */
int gw_mask_bytes(gw_handle *h, uint64_t *out, int clear, void *stream);

/* Diagnostics: from now on the k-th step launch of this handle (k = 0, 1, ... < capacity, counted at
 * ENQUEUE / graph-capture time) records into stamps[4*k .. 4*k+2] the globaltimer (ns) at which its first
 * block started, its first block passed the grid dependency (programmatic dependent launch: the previous
 * kernel of the stream has completed), and its last block finished.  `stamps` is a device uint64 buffer of
 * 4 * capacity entries that the caller initialises to {~0, ~0, 0, 0} per launch; NULL switches it off.
 * Shows per-kernel start / end inside a CUDA-graph replay (profiles/scripts/kernel_stamps.py). */
int gw_debug_stamps(gw_handle *h, uint64_t *stamps, int64_t capacity);

/* Several handles on one device (e.g. env batches stepped round-robin) can accumulate into ONE
 * statistics vector: after gw_share_stats(h, with) the step kernels of `h` add to the accumulators of
 * `with`, and gw_stats of either handle reads / clears them -- one copy instead of one per handle in
 * front of the all-reduce.  `with == NULL` restores the handle's own accumulators.  `with` must outlive
 * the sharing. */
int gw_share_stats(gw_handle *h, gw_handle *with);

/* Read-back of the structure-of-arrays state (tests, tracing, render()): converts one
 * field to a dense device float64 array (integers are exact below 2^53).  n_sims =
 * n_envs * n_bands, sim index = env * n_bands + band. */
#define GW_FIELD_NOW 0              /* [n_envs]                simulated time, simtools.py:56-58 */
#define GW_FIELD_RECEIVED_POWER 1   /* [GW_MAX_DEVICES][n_sims] SimplePhy._receivedPower (mW) */
#define GW_FIELD_NEXT_TICK 2        /* [GW_MAX_SENDERS][n_sims] next traffic tick time */
#define GW_FIELD_COUNTER 3          /* [GW_MAX_SENDERS][n_sims] SenderDevice.counter */
#define GW_FIELD_QUEUE_LEN 4        /* [GW_MAX_SENDERS][n_sims] len(SimpleMac._packetQueue) */
#define GW_FIELD_N_TRANSMISSIONS 5  /* [n_sims]                transmissions started on the band */
#define GW_FIELD_N_DELIVERED 6      /* [GW_MAX_SENDERS][n_sims] packets the RRM decoded, per sender */
#define GW_FIELD_RECEIVED_VALUES 7  /* [2][n_sims]             interpreter.receivedValues */
#define GW_FIELD_ATTENUATION_DB 8   /* [GW_MAX_DEVICES^2][n_sims] FSPL table (receiver-major) */
#define GW_FIELD_RX_POWER_MW 9      /* [GW_MAX_DEVICES^2][n_sims] 10**((P_tx - att)/10) */
#define GW_FIELD_FAULT 10           /* [n_sims]                0 or the condition under which the reference raises */
#define GW_FIELD_TIES 11            /* [n_sims]                exact-time ties seen by the event selector */
#define GW_FIELD_TX_SEQ 12          /* [GW_MAX_DEVICES][n_sims] transmissions started per device (kept in mode M and with
                                       per-env positions -- the only users; 0 in mode R with one shared geometry) */
#define GW_FIELD_PLANT 13           /* [8][n_sims] x, v, theta, omega, motor target velocity, plant time,
                                       controller's angle estimate (deg), PID memory (plant envs) */
#define GW_FIELD_N_RECEIVED 14      /* [GW_MAX_SENDERS][n_sims] packets handed to onReceive (MAC receive mode) */
int gw_read_state(gw_handle *h, int field, double *out, void *stream);

/* Mode M, fed masks: `mask_words` is a device uint32 buffer laid out
 * [n_envs][n_bands][GW_MAX_DEVICES sender][slots][GW_MAX_DEVICES receiver][words_per_row];
 * transmission number q of a sender uses slot q % slots; bit k of a row is the error
 * flag of on-air bit k (bit k%32 of word k/32).  The buffer stays caller-owned.
 *
 * The call makes ONE streaming pass over the buffer on `stream` and keeps a prefix-count index of it in the
 * handle (per row the set bits in front of every 128-bit group: 2 bytes per 16 bytes of mask, allocated on the
 * first call / when the layout changes): SimplePhy._countBitErrors over a section (simple_stack.py:180-188)
 * is then two index look-ups per receiver in the step kernel instead of a scan of the section's mask words
 * inside the event loop.  Call it again after changing mask words in place.  GW_FED_INDEX=0 in the environment
 * (read by this call) skips the index; the step kernels then scan the words themselves. */
int gw_set_masks(gw_handle *h, const uint32_t *mask_words, int32_t slots, int32_t words_per_row,
                 void *stream);

/* The index look-up on its own (numeric tests): counts[i] = set bits among bits [k0[i], k1[i]) of mask row
 * rows[i] (row = (((env * n_bands + band) * GW_MAX_DEVICES + sender) * slots + slot) * GW_MAX_DEVICES + receiver),
 * evaluated through the index built by gw_set_masks.  Device arrays: rows int64 [n], k0 / k1 / counts int32 [n]. */
int gw_mask_index_count(gw_handle *h, const int64_t *rows, const int32_t *k0, const int32_t *k1, int32_t *counts,
                        int64_t n, void *stream);

/* ---- grids of PHY-only senders with in-step mobility ---------------------
 *
 * The reference's own benchmark scenario (tests/test_benchmark.py:20-91): n_devices devices on one frequency
 * band, each a SimplePhy driven by a sender process -- timeout(initial delay), then one SEND of a
 * header_bytes + payload_bytes packet at power_dbm every send_interval (:31-48) -- and every PHY receiving what
 * the others send (simple_stack.py:77-286, mode R accounting); optionally one mobility process per device
 * (:73-85): timeout(move delay), then every move_interval the device moves by the next offset of its tape
 * (the reference's `initialPos` is the moving Position object: offsets accumulate), transmissions on the air
 * seeing SimplePhy._onAttenuationChange (simple_stack.py:119-128).  No RRM, no MAC, no gym step: the envs are
 * advanced with gw_grid_run = SimMan.runSimulation(duration) (simtools.py:77-88).  n_envs independent grids,
 * one GPU thread each; the device count is a run-time value. */
#define GW_GRID_MAX_DEVICES 24

typedef struct {
    int32_t abi_version;            /* GW_ABI_VERSION */
    int64_t n_envs;
    int32_t n_devices;
    double frequency_hz, bandwidth_hz;                  /* FrequencyBandSpec, physical.py:293-306 */
    double power_dbm[GW_GRID_MAX_DEVICES];              /* 40.0 in the reference's benchmark (:45) */
    double send_interval[GW_GRID_MAX_DEVICES];          /* SEND_INTERVAL = 1e-2 (:17) */
    int32_t header_bytes[GW_GRID_MAX_DEVICES];          /* SimpleMacHeader: 13 */
    int32_t payload_bytes[GW_GRID_MAX_DEVICES];         /* Transmittable("A message to all my homies"): 26 */
    double move_interval;                               /* MOVE_INTERVAL = 1e-3 (:18) */
    int32_t max_moves;              /* jumps per device in the offset tape; 0: no mobility processes */
} gw_grid_config;

typedef struct gw_grid_handle gw_grid_handle;

/* Replaces the device_grid / mobile_device_grid fixtures (:52-85).  Device buffers (float64): positions
 * [n_envs][n_devices][2]; delays [n_envs][n_devices] (the fixtures draw random.uniform(0, SEND_INTERVAL));
 * with max_moves > 0: move_delays [n_envs][n_devices] (random.uniform(0, MOVE_INTERVAL)) and offsets
 * [n_envs][n_devices][max_moves][2] (random.uniform(-.2, .2)), which must stay alive as long as the handle (a
 * device stops moving when its tape is exhausted).  The other buffers are consumed on `stream`. */
int gw_grid_create(const gw_grid_config *cfg, int device, const double *positions, const double *delays,
                   const double *move_delays, const double *offsets, void *stream, gw_grid_handle **out);
void gw_grid_destroy(gw_grid_handle *h);

/* SimMan.runSimulation(duration) for every env: all events strictly before now + duration. */
int gw_grid_run(gw_grid_handle *h, double duration, void *stream);
/* The same with an event trace (record format of gw_step_traced, kinds 1-3): trace [n_envs][cap][8],
 * trace_count [n_envs]. */
int gw_grid_run_traced(gw_grid_handle *h, double duration, double *trace, int32_t *trace_count, int32_t cap,
                       void *stream);

#define GW_GRID_FIELD_NOW 0             /* [n_envs] */
#define GW_GRID_FIELD_STATS 1           /* [6][n_devices][n_envs]: transmissions started; headers decoded / failed;
                                           payloads decoded / failed (as a receiver); BER evaluations */
#define GW_GRID_FIELD_POSITIONS 2       /* [2][n_devices][n_envs] */
#define GW_GRID_FIELD_RECEIVED_POWER 3  /* [n_devices][n_envs] SimplePhy._receivedPower (mW) */
#define GW_GRID_FIELD_FAULT 4           /* [n_envs] 0, or the condition under which the reference raises (3: `assert
                                           noisePower >= 0`, simple_stack.py:168-169 -- the incrementally kept power
                                           sum of a PHY can round below the signal power; 2: KeyError, app. B #12);
                                           a faulted env stops simulating */
int gw_grid_read(gw_grid_handle *h, int field, double *out, void *stream);
/* Synchronises; GW_E_SIMFAULT if an env hit a condition under which the reference raises. */
int gw_grid_check(gw_grid_handle *h, void *stream);

/* ---- general band engine: more than two MAC senders / one PHY-only sender per band --
 *
 * CounterTrafficEnv wires exactly two SenderDevices and one RRM (counter_traffic.py:114-133) and gw_create's
 * kernels are specialised for that template (+ at most one PHY-only sender).  The reference's building blocks
 * compose to larger bands: any number of SimpleNetworkDevices with a traffic process, one SimpleRrmDevice whose
 * CounterTrafficInterpreter keeps one received value per device (counter_traffic.py:69-80: the observation
 * stays receivedValues[0] - receivedValues[1]) and PHY-only periodic senders (tests/test_benchmark.py:20-50).
 * This engine steps such bands -- n_senders <= 8, n_phy_senders <= 16 as RUN-TIME values -- exactly like the env:
 * assignFrequencyBand(device, duration) (devices.py:178-203), runSimulation(assignSignal.eProcessed), then
 * Interpreter.getFeedback; reference accounting (mode R) or per-bit Philox error masks (mode M).  Device order:
 * senders, RRM, PHY-only senders.
 * One GPU thread per env, state in global memory ([field][index][env]); gymwipe_b200/csrc/gw_band.cuh. */
#define GW_GENBAND_MAX_SENDERS 8
#define GW_GENBAND_MAX_PHY_SENDERS 16
#define GW_GENBAND_MAX_DEVICES (GW_GENBAND_MAX_SENDERS + 1 + GW_GENBAND_MAX_PHY_SENDERS)

typedef struct {
    int32_t abi_version;            /* GW_ABI_VERSION */
    int64_t n_envs;
    int32_t n_senders, n_phy_senders;
    int32_t assignment_duration_factor;     /* ASSIGNMENT_DURATION_FACTOR = 1000, envs/core.py:36 */
    int32_t max_assign_duration;            /* MAX_ASSIGN_DURATION = 20, envs/core.py:31: duration in [0, 20) */
    int32_t per_env_positions;              /* 0: one geometry for all envs; 1: positions per env */
    int32_t mode;                           /* GW_MODE_REFERENCE, or GW_MODE_MASK_PHILOX: per-bit error masks keyed */
    uint64_t seed;                          /*   by (seed; env_id_offset + env, sender, transmission, receiver, bit) */
    int64_t env_id_offset;                  /*   exactly as in gw_config (GW_MODE_MASK_FED is not offered here) */
    double frequency_hz, bandwidth_hz;      /* FrequencyBandSpec, physical.py:293-306 */
    /* senders (SenderDevice, counter_traffic.py:37-61) */
    int32_t multiplicity[GW_GENBAND_MAX_SENDERS];       /* packets per tick */
    int32_t payload_bytes[GW_GENBAND_MAX_SENDERS];      /* < 0: byteSize = counter (the reference's rule) */
    int32_t destination[GW_GENBAND_MAX_SENDERS];        /* sender index the packets are addressed to */
    int32_t max_ticks[GW_GENBAND_MAX_SENDERS];          /* 0: forever; n: a burst of n ticks */
    int32_t receive[GW_GENBAND_MAX_SENDERS];            /* 1: SimpleNetworkDevice.receiving = True */
    double interval[GW_GENBAND_MAX_SENDERS];            /* COUNTER_INTERVAL = 1e-3 */
    /* PHY-only periodic senders (tests/test_benchmark.py:31-48) */
    double phy_interval[GW_GENBAND_MAX_PHY_SENDERS], phy_delay[GW_GENBAND_MAX_PHY_SENDERS];
    double phy_power_dbm[GW_GENBAND_MAX_PHY_SENDERS];
    int32_t phy_header_bytes[GW_GENBAND_MAX_PHY_SENDERS], phy_payload_bytes[GW_GENBAND_MAX_PHY_SENDERS];
} gw_genband_config;

typedef struct gw_genband_handle gw_genband_handle;

/* `positions`: device float64 [n_devices][2], or [n_envs][n_devices][2] with per_env_positions (consumed on
 * `stream`).  The envs start in the state after construction (CounterTrafficEnv.__init__); state is owned by the
 * handle. */
int gw_genband_create(const gw_genband_config *cfg, int device, const double *positions, void *stream,
                      gw_genband_handle **out);
void gw_genband_destroy(gw_genband_handle *h);
/* Devices moving between steps, for handles with per_env_positions: positions device float64 [n_envs][n_devices][2].
 * Every env moves its devices one after the other by ascending index like successive Position.set calls
 * (devices/core.py:75-84) with the semantics of gw_set_positions on a stepped handle: attenuation models beyond
 * STANDBY_THRESHOLD or with coinciding devices keep their value, only a new value triggers, pairs that have not
 * transmitted yet follow the positions, and transmissions that are on the air go through
 * SimplePhy._onAttenuationChange (simple_stack.py:119-128). */
int gw_genband_set_positions(gw_genband_handle *h, const double *positions, void *stream);
/* Mobility processes that move devices DURING the steps (the mover of tests/test_benchmark.py:73-85), for handles with
 * per_env_positions, started at the envs' current time in device order: move_delays device float64
 * [n_envs][n_devices] (the first delay; < 0: the device has no process), offsets [n_envs][n_devices][max_moves][2]
 * (the jumps, which accumulate: the reference's `initialPos` is the moving Position object; a process ends with its
 * tape), one jump every move_interval.  `offsets` must stay alive as long as the handle; `move_delays` is consumed
 * on `stream`.  Once per handle. */
int gw_genband_set_movers(gw_genband_handle *h, const double *move_delays, const double *offsets, int32_t max_moves,
                          double move_interval, void *stream);
/* CounterTrafficEnv.reset (counter_traffic.py:135-144) of every env; obs (device int64 [n_envs]) may be NULL. */
int gw_genband_reset(gw_genband_handle *h, int64_t *obs, void *stream);
/* CounterTrafficEnv.step (counter_traffic.py:146-158).  Device arrays [n_envs]: device in [0, n_senders),
 * duration in [0, max_assign_duration); obs int64, reward float64, done uint8. */
int gw_genband_step(gw_genband_handle *h, const int32_t *device, const int32_t *duration, int64_t *obs, double *reward,
                    uint8_t *done, void *stream);
/* The same with an event trace (record format of gw_step_traced): trace [n_envs][cap][8], trace_count [n_envs]. */
int gw_genband_step_traced(gw_genband_handle *h, const int32_t *device, const int32_t *duration, int64_t *obs,
                           double *reward, uint8_t *done, double *trace, int32_t *trace_count, int32_t cap, void *stream);

#define GW_GENBAND_FIELD_NOW 0              /* [n_envs] */
#define GW_GENBAND_FIELD_DELIVERED 1        /* [n_senders][n_envs] data packets of each sender decoded by the RRM */
#define GW_GENBAND_FIELD_RECEIVED 2         /* [n_senders][n_envs] packets handed to onReceive (receive mode) */
#define GW_GENBAND_FIELD_TRANSMISSIONS 3    /* [n_envs] */
#define GW_GENBAND_FIELD_FAULT 4            /* [n_envs] 0, or the condition under which the reference raises */
#define GW_GENBAND_FIELD_RECEIVED_POWER 5   /* [n_devices][n_envs] SimplePhy._receivedPower (mW) */
#define GW_GENBAND_FIELD_QUEUE_LENGTH 6     /* [n_senders][n_envs] */
#define GW_GENBAND_FIELD_COUNTER 7          /* [n_senders][n_envs] */
#define GW_GENBAND_FIELD_TIES 8             /* [n_envs] exact-time ties between independent events (diagnostic) */
#define GW_GENBAND_FIELD_RECEIVED_VALUES 9  /* [n_senders][n_envs] CounterTrafficInterpreter.receivedValues (counter_traffic.py:69-80) */
int gw_genband_read(gw_genband_handle *h, int field, double *out, void *stream);
/* Synchronises; GW_E_ACTION / GW_E_SIMFAULT like gw_check. */
int gw_genband_check(gw_genband_handle *h, void *stream);

/* ---- standalone kernels (numeric parity tests, roofline measurements) ------------- */

/* K1: FSPL attenuation in dB, FsplAttenuation._update (attenuation_models.py:28-36) with
 * Position.distanceTo (devices/core.py:88-95).  All arrays device float64[n]. */
int gw_fspl_attenuation(const double *ax, const double *ay, const double *bx, const double *by,
                        double frequency_hz, double *att_db, int64_t n, void *stream);

/* K2: BPSK bit error rate from signal / noise power in mW: SimplePhy._updateBitErrorRate
 * (simple_stack.py:161-173) -> BpskMcs.calculateBitErrorRate (physical.py:208-212) ->
 * calculateEbToN0Ratio (:25-42) -> approxQFunction (:46-58). */
int gw_ber_bpsk(const double *signal_mw, const double *noise_mw, double *ber, int64_t n,
                void *stream);

/* K3: bit-error count of mask rows over bit ranges (mode M accounting,
 * SimplePhy._countBitErrors, simple_stack.py:180-188, with per-bit masks):
 * counts[i] = popcount(bits [k0[i], k1[i]) of row rows[i]).  HBM-bound streaming
 * popcount, one warp per descriptor. */
int gw_count_bit_errors(const uint32_t *mask_words, int32_t words_per_row,
                        const int64_t *rows, const int32_t *k0, const int32_t *k1,
                        int32_t *counts, int64_t n, void *stream);

/* Action selection of the learner that consumes the env (agents/dqn_counter_traffic.py:46-63), fused into one
 * kernel over obs[n]: the 1-16-16-16-n_actions ReLU MLP on (obs - obs_center), keras-rl's BoltzmannQPolicy
 * (p ~ exp(clip(q / tau, clip_lo, clip_hi)) in float64; the reference uses tau 1, clip +-500) and the draw -- the
 * uniform variate of (env_id_offset + i, counter) comes from Philox4x32-10 keyed by `seed`, so samples do not
 * depend on batch size or sharding.  `weights`: device float32 in the order of torch's model.parameters()
 * (W1[16][1], b1[16], W2[16][16], b2[16], W3[16][16], b3[16], W4[n_actions][16], b4[n_actions]); `obs`: device
 * int64 [n].  Outputs (device, each may be NULL): flat_action int64 [n]; device / duration int32 [n]
 * (flat // n_durations, flat % n_durations: CounterTrafficProcessor.process_action, :25-33 -- directly usable
 * as gw_step's action arrays); probs float64 [n][n_actions] (tests).  n_actions <= 160 (8 senders x 20 durations:
 * the action space of the largest band of the general engine; 8, 20 and 40 take unrolled kernels). */
int gw_policy_boltzmann(const float *weights, int32_t n_actions, int32_t n_durations, const int64_t *obs, int64_t n,
                        float obs_center, double tau, double clip_lo, double clip_hi, uint64_t seed, uint64_t counter,
                        int64_t env_id_offset, int64_t *flat_action, int32_t *device, int32_t *duration, double *probs,
                        void *stream);

/* Philox4x32-10 (Random123) block function, for known-answer tests: out[4*i..] =
 * philox(counter[4*i..], key[2*i..]).  Device uint32 arrays. */
int gw_philox4x32(const uint32_t *counter, const uint32_t *key, uint32_t *out, int64_t n,
                  void *stream);

/* The decider threshold Mcs.maxCorrectableBer (physical.py:160-185), host-side. */
double gw_max_correctable_ber(int k, int n);

#ifdef __cplusplus
}
#endif
#endif /* GYMWIPE_B200_H */
