#!/usr/bin/env python
"""
bench.py -- env-steps/sec of the batched CounterTrafficEnv hot path (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Own arm: N ranks (one per GPU; torchrun supplies RANK / LOCAL_RANK / WORLD_SIZE for N > 1), each
owning an independent shard of 16 batches of 65,536 envs (BASELINE configs[1], weak scaling).  A
"step" is ONE pass of the hot path over ONE batch (`CounterTrafficEnv.step` of its 65,536 envs = one
launch of the fused step kernel); the batches are stepped round-robin so that every launch finds its
inputs in HBM, not in L2 ("inputs larger than L2", no flush).  W warm-up launches, then EXACTLY K timed
launches replayed from CUDA graphs, one CUDA-event pair around them on the launching stream, barrier +
synchronize on both sides, max over ranks.  `e2e` is the same metric through `gw_step_host_packed`
with pinned HOST buffers (H2D actions + D2H obs/reward/done inside the timed region).

Reference arm (`--impl reference`): the reference's algorithm on the box's host cores -- the
oracle port (plain-C restatement, pinned bit-exactly against the unmodified Python reference;
the reference itself is pure Python on SimPy and cannot travel to the GPU box), all host
threads, same config / metric / unit; rank 0 only.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

ENVS_PER_GPU = 65536
ROTATING_BATCHES = 16                  # independent 65,536-env batches stepped round-robin (inputs > L2)
BURN_IN_STEPS = 128                    # steps of every env before the timed region (steady state, see own_arm)
PRODUCTIVE_STEPS = 32                  # steps of every fresh env timed separately (productive regime)
ALGO_BYTES_PER_ENV_STEP = 193          # SURVEY.md section 8d / DESIGN.md section 6
STATS_EVERY_CHUNKS = 4                 # graph chunks (of 64 launches) between two NCCL reductions of the statistics vector
METRIC = "env-steps/sec CounterTrafficEnv batch"
UNIT = "env-steps/s"


def config_dict(n_envs_total, parallelism, regime):
    return {"workload": "CounterTrafficEnv default scenario (2 counter senders + RRM, 1 FrequencyBand, FSPL, BPSK), "
                        "mode R (reference-exact accounting), batches of %d envs per GPU (BASELINE configs[1]), random "
                        "actions (device~U{0,1}, duration~U{0..19}); one step = one launch of the fused step kernel over "
                        "one batch" % ENVS_PER_GPU,
            "n_envs": n_envs_total * ROTATING_BATCHES, "envs_per_launch": ENVS_PER_GPU,
            "batches_per_gpu": ROTATING_BATCHES, "parallelism": parallelism,
            "regime": regime,
            "l2": "inputs larger than L2: %d independent %d-env batches per GPU are stepped round-robin, so a batch's "
                  "state (~13 MB hot), its fresh action rows and outputs are re-touched only after ~%d MB of other "
                  "traffic (L2 = 126 MB); no flush, launches replayed from CUDA graphs (64 launches each), one "
                  "CUDA-event pair around the K timed launches" % (ROTATING_BATCHES, ENVS_PER_GPU, 13 * (ROTATING_BATCHES - 1))}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic_per_launch():
    """DRAM bytes per step-kernel launch from the committed ncu capture (profiles/), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "step_kernel_traffic.json")) as f:
            d = json.load(f)
        return d.get("dram_bytes_per_launch")
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe).
    The sampler is started before the warm-up (nvidia-smi needs ~0.1 s to come up); samples are
    kept if their timestamp falls inside [mark_begin, mark_end] -- the timed loop."""

    FIELDS = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.file = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        self.t0 = self.t1 = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=self.file, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def stop(self):
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.file.flush()
        self.file.seek(0)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = []
        for line in self.file.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 10:
                continue
            try:
                ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(parts[2]), float(parts[3]),
                             [nm for nm, v in zip(names, parts[6:10]) if v.lower().startswith("active")]))
            except ValueError:
                continue
        try:
            os.unlink(self.file.name)
        except OSError:
            pass
        inside = [r for r in rows if self.t0 is not None and self.t0 - 0.02 <= r[0] <= self.t1 + 0.02]
        window = "timed loop"
        if not inside:
            inside, window = rows, "warm-up + timed loop (the timed loop is shorter than the sampling period)"
        if inside:
            sm = sorted(r[1] for r in inside)
            reasons = sorted({x for r in inside for x in r[3]})
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(r[2] for r in inside), reasons=reasons,
                       samples=len(inside), window=window)
        return out


def mask_scan_roofline(dev_t, peak, order="random"):
    """
    K3, the HBM-bound kernel of the path (mode M accounting over fed masks): streaming popcount of
    1 Mi mask rows of 2 KiB (one 1525-byte packet as seen by one receiver: 16 267 on-air bits),
    2 GiB in total -- far larger than the 126 MB L2, so every launch streams from HBM.
    """
    import torch
    from gymwipe_b200 import _native as N
    rows, words = 1 << 20, 512
    nbits = 16267
    masks = torch.randint(-2 ** 31, 2 ** 31 - 1, (rows, words), dtype=torch.int32, device=dev_t)
    ridx = torch.randperm(rows, device=dev_t).to(torch.int64) if order == "random" else \
        torch.arange(rows, device=dev_t, dtype=torch.int64)
    k0 = torch.zeros(rows, dtype=torch.int32, device=dev_t)
    k1 = torch.full((rows,), nbits, dtype=torch.int32, device=dev_t)
    out = torch.empty(rows, dtype=torch.int32, device=dev_t)
    stream = torch.cuda.current_stream(dev_t)
    lib = N.lib()

    def launch():
        N.check(lib.gw_count_bit_errors(masks.data_ptr(), words, ridx.data_ptr(), k0.data_ptr(), k1.data_ptr(),
                                        out.data_ptr(), rows, stream.cuda_stream))
    for _ in range(3):
        launch()
    torch.cuda.synchronize(dev_t)
    reps = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        launch()
    e1.record(stream)
    torch.cuda.synchronize(dev_t)
    ms = e0.elapsed_time(e1) / reps
    algo = rows * ((nbits + 7) // 8 + 8 + 4 + 4 + 4)       # mask bytes + descriptor (row, k0, k1) + count
    achieved = algo / (ms * 1e-3) / 1e9
    check = int(out[:1024].sum())
    del masks
    return {"kernel": "count_bits_kernel (gw_count_bit_errors)", "bound": "hbm", "achieved": achieved, "peak": peak,
            "unit": "GB/s", "frac": achieved / peak, "rows": rows, "bits_per_row": nbits,
            "algorithmic_bytes_per_launch": algo, "avg_launch_ms": ms, "input": "2 GiB of masks (> L2), %s row order" % order,
            "checksum_first_1024": check}


def cfg3_long_packet(dev_t, steps=24):
    """
    BASELINE configs[2]: 1500-byte payloads, fed per-bit masks, a PHY-only interferer creating
    mid-packet SINR segments; ASSIGNMENT_DURATION_FACTOR = 10000 so that windows fit 122 ms packets.
    """
    import torch
    import gymwipe_b200
    n, slots, words = 65536, 2, 512
    sc = {"assignment_duration_factor": 10000, "bands": [{"frequency": 2.4e9, "bandwidth": 22e6, "devices": [
        {"role": "sender", "x": 0.0, "y": 2.0, "mult": 1, "payload": 1500, "interval": 0.001, "dest": 1},
        {"role": "sender", "x": 0.0, "y": -2.0, "mult": 3, "payload": 1500, "interval": 0.001, "dest": 0},
        {"role": "rrm", "x": 0.0, "y": 0.0},
        {"role": "jammer", "x": 6.0, "y": 0.0, "interval": 0.05, "delay": 0.003, "power": 0.0, "hdr": 13, "payload": 200}]}]}
    env = gymwipe_b200.make('CounterTraffic-v0', num_envs=n, device=dev_t, scenario=sc, mode="mask_fed", strict=False)
    # Bernoulli(1/16) masks: 4 GiB, resident before the timed region
    shape = (n, 1, 4, slots, 4, words)
    masks = torch.randint(-2 ** 31, 2 ** 31 - 1, shape, dtype=torch.int32, device=dev_t)
    for _ in range(3):                                  # AND of 4 random words: bit density 1/16
        masks &= torch.randint(-2 ** 31, 2 ** 31 - 1, shape, dtype=torch.int32, device=dev_t)
    env.set_masks(masks, slots)
    env.reset()
    g = torch.Generator(device=dev_t).manual_seed(7)
    a_dev = torch.randint(0, 2, (steps + 4, n), generator=g, device=dev_t, dtype=torch.int32)
    a_dur = torch.randint(12, 20, (steps + 4, n), generator=g, device=dev_t, dtype=torch.int32)
    stream = torch.cuda.current_stream(dev_t)
    for t in range(4):
        env.step({"device": a_dev[t], "duration": a_dur[t]})
    env.stats()
    torch.cuda.synchronize(dev_t)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for t in range(4, steps + 4):
        env.step({"device": a_dev[t], "duration": a_dur[t]})
    e1.record(stream)
    torch.cuda.synchronize(dev_t)
    env.check()
    st = env.stats().cpu().numpy()
    ms = e0.elapsed_time(e1) / steps
    return {"workload": "configs[2]: 1500-byte payloads, fed per-bit masks (mode M), PHY-only interferer, %d envs" % n,
            "env_steps_per_s": n / (ms * 1e-3), "ms_per_step": ms, "transmissions_per_step": float(st[6]) / steps,
            "deliveries_per_step": float(st[1] + st[2]) / steps, "mask_bytes_resident": int(masks.numel() * 4)}


def cfg4_multiband(dev_t, steps=64):
    """
    BASELINE configs[3] on ONE GPU's share: 16 devices over 4 FrequencyBands (per band: RRM + 2 MAC
    senders + 1 PHY-only interferer), positions ~U(-20, 20) m per env, one action per band; the env's
    clock ends at the latest band (mode R).  131 072 envs x 4 bands = 524 288 band-sims.
    """
    import torch
    import gymwipe_b200
    n = 131072
    bands = []
    for b in range(4):
        bands.append({"frequency": 2.4e9 + b * 25e6, "bandwidth": 22e6, "devices": [
            {"role": "sender", "x": 0.0, "y": 2.0, "mult": 1, "payload": "counter", "interval": 0.001, "dest": 1},
            {"role": "sender", "x": 0.0, "y": -2.0, "mult": 3, "payload": "counter", "interval": 0.001, "dest": 0},
            {"role": "rrm", "x": 0.0, "y": 0.0},
            {"role": "jammer", "x": 5.0, "y": 5.0, "interval": 0.013 + 0.002 * b, "delay": 0.001 * b, "power": 10.0,
             "hdr": 13, "payload": 60}]})
    sc = {"assignment_duration_factor": 1000, "bands": bands}
    g = torch.Generator(device=dev_t).manual_seed(11)
    pos = (torch.rand((n, 4, 4, 2), generator=g, device=dev_t, dtype=torch.float64) * 40.0 - 20.0)
    env = gymwipe_b200.make('CounterTraffic-v0', num_envs=n, device=dev_t, scenario=sc, positions=pos, strict=False)
    env.reset()
    a_dev = torch.randint(0, 2, (steps + 4, n, 4), generator=g, device=dev_t, dtype=torch.int32)
    a_dur = torch.randint(0, 20, (steps + 4, n, 4), generator=g, device=dev_t, dtype=torch.int32)
    stream = torch.cuda.current_stream(dev_t)
    for t in range(4):
        env.step({"device": a_dev[t], "duration": a_dur[t]})
    env.stats()
    torch.cuda.synchronize(dev_t)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for t in range(4, steps + 4):
        env.step({"device": a_dev[t], "duration": a_dur[t]})
    e1.record(stream)
    torch.cuda.synchronize(dev_t)
    env.check()
    st = env.stats().cpu().numpy()
    ms = e0.elapsed_time(e1) / steps
    return {"workload": "configs[3] share of one GPU: %d envs x 4 bands x 4 devices, per-env positions, mode R" % n,
            "env_steps_per_s": n / (ms * 1e-3), "band_steps_per_s": 4 * n / (ms * 1e-3), "ms_per_step": ms,
            "transmissions_per_env_step": float(st[6]) / steps / n, "deliveries_per_env_step": float(st[1] + st[2]) / steps / n}


def cfg5_pendulum(dev_t, steps=32):
    """
    BASELINE configs[4] on ONE GPU's share: the networked inverted-pendulum env (sensor / controller
    band assignment, in-kernel RK4 plant; PARITY UNPINNED), 131 072 envs.
    """
    import torch
    import gymwipe_b200
    n = 131072
    env = gymwipe_b200.make('InvertedPendulum-v0', num_envs=n, device=dev_t, strict=False)
    g = torch.Generator(device=dev_t).manual_seed(5)
    a_dev = torch.randint(0, 2, (steps + 4, n), generator=g, device=dev_t, dtype=torch.int32)
    a_dur = torch.randint(1, 20, (steps + 4, n), generator=g, device=dev_t, dtype=torch.int32)
    stream = torch.cuda.current_stream(dev_t)
    for t in range(4):
        env.step({"device": a_dev[t], "duration": a_dur[t]})
    torch.cuda.synchronize(dev_t)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for t in range(4, steps + 4):
        env.step({"device": a_dev[t], "duration": a_dur[t]})
    e1.record(stream)
    torch.cuda.synchronize(dev_t)
    env.check()
    ms = e0.elapsed_time(e1) / steps
    th = env.plant_state()[2]
    return {"workload": "configs[4] share of one GPU: networked inverted pendulum, %d envs, in-kernel RK4 plant "
                        "(parity unpinned: the reference env is unconstructible)" % n,
            "env_steps_per_s": n / (ms * 1e-3), "ms_per_step": ms,
            "mean_abs_angle_deg": float(torch.rad2deg(th).abs().mean())}


def cpu_baseline_run(target_seconds, threads=None):
    """The oracle port on the host cores, on a bounded sample of the same workload (steady state: every
    env is burnt in for BURN_IN_STEPS steps first; only the steps after that are timed, per thread)."""
    import numpy as np
    import gw_oracle as O
    from gymwipe_b200.scenario import default_scenario_dict
    sc = default_scenario_dict()
    threads = threads or os.cpu_count() or 1
    rs = np.random.RandomState(0)
    T = 256

    def run(nenv):
        steps = BURN_IN_STEPS + T
        dev = rs.randint(0, 2, size=(steps, nenv)).astype(np.int32)
        dur = rs.randint(0, 20, size=(steps, nenv)).astype(np.int32)
        r = O.run_batch(sc, dev, dur, threads=threads, want=("obs", "reward"), time_from=BURN_IN_STEPS)
        return r["seconds"]
    run(threads * 4)                                    # warm-up (page-in, thread start)
    probe_n = threads * 16
    dt = run(probe_n)
    rate = probe_n * T / max(dt, 1e-9)
    nenv = int(max(probe_n, min(rate * target_seconds / (2 * (T + BURN_IN_STEPS)), 200000)))
    nenv = (nenv // threads) * threads
    best = None
    for _ in range(2):
        v = nenv * T / max(run(nenv), 1e-9)
        best = v if best is None else max(best, v)
    return {"value": best, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "%d envs x %d steps after %d burn-in steps (steady state, same action distribution), oracle C "
                      "restatement, %d host threads (the slowest thread's time in the timed steps), best of 2"
                      % (nenv, T, BURN_IN_STEPS, threads)}


def reference_arm(args, rank):
    """--impl reference: the reference's CPU algorithm (oracle port), all host threads, rank 0 only."""
    if rank != 0:
        return 0
    import numpy as np
    import gw_oracle as O
    from gymwipe_b200.scenario import default_scenario_dict
    sc = default_scenario_dict()
    threads = os.cpu_count() or 1
    K, W = args.steps, args.warmup
    # a "step" = one env.step of a bounded sample of the batch
    # sized by the TIMED work: ~100k timed env-steps per thread (a few tenths of a second -- shorter regions
    # are dominated by scheduling noise), within ~40M simulated env-steps in total (burn-in included)
    per_thread = max(8, min(100_000 // max(K, 1), 40_000_000 // ((BURN_IN_STEPS + W + K) * threads)))
    sample = threads * per_thread
    rs = np.random.RandomState(0)
    # every sampled env is simulated from construction; only its steps after the burn-in and the warm-up are
    # timed (per thread, the slowest thread bounds the batch)
    B0 = BURN_IN_STEPS
    dev = rs.randint(0, 2, size=(B0 + W + K, sample)).astype(np.int32)
    dur = rs.randint(0, 20, size=(B0 + W + K, sample)).astype(np.int32)
    r = O.run_batch(sc, dev, dur, threads=threads, want=("obs",), time_from=B0 + W)
    elapsed = max(r["seconds"], 1e-9)
    value = sample * K / elapsed
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": K, "warmup": W, "ms_per_step": 1e3 * elapsed / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(ENVS_PER_GPU * args.gpus, "host threads x%d" % threads,
                                  "steady state: steps %d..%d of every sampled env (the first %d steps are simulated untimed)"
                                  % (BURN_IN_STEPS + W, BURN_IN_STEPS + W + K, BURN_IN_STEPS + W)),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": "%d independent envs of the same workload per step (sized so that the timed region is "
                                       "a few tenths of a second per thread), %d steps, oracle C restatement of the "
                                       "reference's SimPy path" % (sample, K)},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "the reference is pure Python on SimPy (~1e3 env-steps/s/core measured in the build "
                    "container, BASELINE.md); it cannot run on the GPU box, so its algorithm is timed via the "
                    "bit-exact C restatement"}
    print(json.dumps(line))
    return 0


def own_arm(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist
    import gymwipe_b200
    from gymwipe_b200.distributed import StatsReducer

    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.set_num_threads(1)                            # no idle worker threads spinning next to the stepping thread
    torch.cuda.set_device(local_rank)
    dev_t = torch.device("cuda", local_rank)
    K, W = args.steps, args.warmup
    n = ENVS_PER_GPU
    M = ROTATING_BATCHES
    # M independent 65,536-env batches stepped round-robin: one "step" = one launch of the fused
    # step kernel over one batch.  A batch is touched again only after the M - 1 others (M x ~13 MB of
    # hot state + fresh action rows + outputs > the 126 MB L2), so every launch finds its inputs in
    # HBM, not in L2 ("inputs larger than L2"); nothing is flushed and nothing sits between two
    # launches inside the timed region.
    envs = [gymwipe_b200.make('CounterTraffic-v0', num_envs=n, device=dev_t,
                              env_id_offset=(rank * M + b) * n, strict=False) for b in range(M)]
    for e in envs:
        e.reset()
    for e in envs[1:]:
        e.share_stats(envs[0])                          # one statistics vector per GPU (gw_share_stats)

    # synthetic action tapes for every launch, resident in HBM before the timed region
    g = torch.Generator(device=dev_t).manual_seed(1234 + rank)
    total = W + K
    PROD = PRODUCTIVE_STEPS * M                         # launches of the productive-regime measurement
    BURN = BURN_IN_STEPS * M                            # launches up to the steady state
    rows = max(total, PROD)
    a_dev = torch.randint(0, 2, (rows, n), generator=g, device=dev_t, dtype=torch.int32)
    a_dur = torch.randint(0, 20, (rows, n), generator=g, device=dev_t, dtype=torch.int32)
    reducer = StatsReducer(dev_t) if world > 1 else None
    stream = torch.cuda.Stream(device=dev_t)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    counter = [0]                                       # launches so far: launch j steps batch j % M

    def launch(row):
        envs[counter[0] % M].step({"device": a_dev[row], "duration": a_dur[row]})
        counter[0] += 1

    CHUNK = 64

    def capture(first_row, count):
        """CUDA graphs of <= CHUNK launches (the pointers of the action rows are baked in, hence one
        graph per chunk); capturing does not execute.  Returns [(graph, launches)]."""
        out, j, c0 = [], 0, counter[0]
        while j < count:
            cnt = min(CHUNK, count - j)
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=stream):
                for q in range(cnt):
                    launch((first_row + j + q) % rows)
            out.append((gr, cnt))
            j += cnt
        assert counter[0] == c0 + count
        return out

    torch.cuda.synchronize(dev_t)
    with torch.cuda.stream(stream):
        # (1) productive regime (the first ~100 steps after construction / reset(): queues hold packets
        # that fit the windows): PRODUCTIVE_STEPS steps of every fresh env, timed for the record
        gp = capture(0, PROD)
        torch.cuda.synchronize(dev_t)
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record(stream)
        for gr, _ in gp:
            gr.replay()
        p1.record(stream)
        torch.cuda.synchronize(dev_t)
        prod_ms = p0.elapsed_time(p1) / PROD
        del gp
        # (2) burn-in to the steady state of the reference's workload: its training run
        # (agents/dqn_counter_traffic.py: one reset(), dqn.fit(nb_steps=50000), `done` never true) leaves
        # the productive regime after ~100 steps and spends > 99 % of its steps in the regime where the
        # counters are too large for any window (announcements only)
        for j in range(BURN - PROD):
            launch(j % rows)
        # (3) W warm-up launches (+ one untimed replay of the K-launch graphs), then EXACTLY K timed launches
        for j in range(W):
            launch(j)
        torch.cuda.synchronize(dev_t)
        graphs = capture(W, K)
        # one untimed replay of every graph: the first launch of an instantiated graph uploads it to the
        # device (a one-time cost like any warm-up); the envs simply advance K more steps
        for gr, _ in graphs:
            gr.replay()
        torch.cuda.synchronize(dev_t)
        def reduce_stats():
            # K5 partial sums of all batches -> NCCL all-reduce on a side stream
            envs[0].stats(out=reducer.next_slot())
            reducer.submit()
        if reducer is not None:
            # first use loads the small kernels and sets up NCCL's channels: not part of stepping
            reduce_stats()
            reducer.drain()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev_t)
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(len(graphs) + 1)]
        if sampler is not None:
            sampler.mark_begin()
        wall0 = time.perf_counter()
        marks[0].record(stream)
        for c, (gr, cnt) in enumerate(graphs):
            gr.replay()
            if reducer is not None and ((c + 1) % STATS_EVERY_CHUNKS == 0 or c + 1 == len(graphs)):
                reduce_stats()
            marks[c + 1].record(stream)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev_t)
        wall = time.perf_counter() - wall0
    if sampler is not None:
        sampler.mark_end()
    clocks = sampler.stop() if sampler is not None else None
    for e in envs:
        e.check()
    if reducer is not None:
        reducer.drain()
    chunk_ms = np.array([marks[c].elapsed_time(marks[c + 1]) for c in range(len(graphs))])
    chunk_cnt = np.array([cnt for _, cnt in graphs])
    elapsed_ms = float(marks[0].elapsed_time(marks[-1]))
    per_step_ms = chunk_ms / chunk_cnt                  # average launch duration per chunk

    # transparency: (a) one batch stepped back to back from one graph (state stays in L2),
    # (b) the round-1 protocol: per-launch event pairs with a 256 MiB L2-flush memset before each
    env1 = envs[0]
    KB = 64
    with torch.cuda.stream(stream):
        gw = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gw, stream=stream):
            for k in range(KB):
                env1.step({"device": a_dev[k], "duration": a_dur[k]})
        gw.replay()
        torch.cuda.synchronize(dev_t)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        gw.replay()
        e1.record(stream)
        torch.cuda.synchronize(dev_t)
    warm_ms = e0.elapsed_time(e1) / KB
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev_t)
    KF = 64
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(KF)]
    cur = torch.cuda.current_stream(dev_t)
    for k in range(KF):
        flush.zero_()
        evs[k][0].record(cur)
        env1.step({"device": a_dev[k], "duration": a_dur[k]})
        evs[k][1].record(cur)
    torch.cuda.synchronize(dev_t)
    flushed_ms = float(np.mean([a.elapsed_time(b) for a, b in evs]))
    del flush
    for e in envs[1:]:
        e.share_stats(None)
        e.close()
    del envs, graphs, gw
    torch.cuda.empty_cache()

    # e2e: host buffers through gw_step_host (pinned), copies inside the timed region
    env2 = gymwipe_b200.make('CounterTraffic-v0', num_envs=n, device=dev_t, env_id_offset=rank * n, strict=False)
    env2.reset()

    def burn_in(env):
        # to the steady state, through the device-resident step (untimed)
        for t in range(BURN_IN_STEPS):
            env.step({"device": a_dev[t % rows], "duration": a_dur[t % rows]})
        torch.cuda.synchronize(dev_t)
    burn_in(env2)
    KE = min(K, 256)
    h_dev = a_dev[:W + KE].cpu().pin_memory()
    h_dur = a_dur[:W + KE].cpu().pin_memory()
    h_obs = torch.empty(n, dtype=torch.int64).pin_memory()
    h_rew = torch.empty(n, dtype=torch.float64).pin_memory()
    h_done = torch.empty(n, dtype=torch.uint8).pin_memory()
    for t in range(W):
        env2.step_host(h_dev[t], h_dur[t], h_obs, h_rew, h_done)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev_t)
    t0 = time.perf_counter()
    for k in range(KE):
        env2.step_host(h_dev[W + k], h_dur[W + k], h_obs, h_rew, h_done)
    torch.cuda.synchronize(dev_t)
    e2e_wide_s = time.perf_counter() - t0
    # packed variant (gw_step_host_packed): one copy in (8 B/env), one copy out (9 B/env)
    h_act = torch.stack([h_dev, h_dur], dim=1).contiguous().pin_memory()      # [steps, 2, n]
    h_res = torch.empty(9 * n, dtype=torch.uint8).pin_memory()
    env3 = gymwipe_b200.make('CounterTraffic-v0', num_envs=n, device=dev_t, env_id_offset=rank * n, strict=False)
    env3.reset()
    burn_in(env3)
    for t in range(W):
        env3.step_host_packed(h_act[t], h_res)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev_t)
    t0 = time.perf_counter()
    for k in range(KE):
        env3.step_host_packed(h_act[W + k], h_res)
    torch.cuda.synchronize(dev_t)
    e2e_packed_s = time.perf_counter() - t0
    checksum = float(env3.unpack_results(h_res)[1].double().sum())
    assert torch.equal(env3.unpack_results(h_res)[0].to(torch.int64), h_obs)   # both paths agree on the last step
    # compact variant (gw_step_host_compact): uint8 actions [n][2] in (2 B/env), one packed word out (4 B/env)
    h_act8 = torch.stack([h_dev, h_dur], dim=2).to(torch.uint8).contiguous().pin_memory()     # [steps, n, 2]
    h_res32 = torch.empty(n, dtype=torch.int32).pin_memory()
    env4 = gymwipe_b200.make('CounterTraffic-v0', num_envs=n, device=dev_t, env_id_offset=rank * n, strict=False)
    env4.reset()
    burn_in(env4)
    act_rows = [h_act8[t] for t in range(W + KE)]      # views of the pinned action tape, one per step
    for t in range(W):
        env4.step_host_compact(act_rows[t], h_res32)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev_t)
    t0 = time.perf_counter()
    for k in range(KE):
        env4.step_host_compact(act_rows[W + k], h_res32)
    torch.cuda.synchronize(dev_t)
    e2e_s = time.perf_counter() - t0
    env4.check()
    assert torch.equal(env4.unpack_compact(h_res32)[0], h_obs)                  # all three paths agree on the last step
    # pipelined variant: PIPE env batches in flight through gw_step_host_compact_async -- the host waits for
    # a batch's previous results (event), reads them, then submits that batch's next actions; kernels of
    # different batches queue behind each other, so launch and synchronisation latencies overlap with compute
    PIPE = 4
    penvs = [env4] + [gymwipe_b200.make('CounterTraffic-v0', num_envs=n, device=dev_t,
                                        env_id_offset=(rank * PIPE + b) * n, strict=False) for b in range(1, PIPE)]
    for e in penvs[1:]:
        e.reset()
        burn_in(e)
    pres = [torch.empty(n, dtype=torch.int32).pin_memory() for _ in range(PIPE)]
    pres_np = [r.numpy() for r in pres]
    pev = [torch.cuda.Event() for _ in range(PIPE)]
    cur = torch.cuda.current_stream(dev_t)
    nrows = len(act_rows)

    def pipelined(count):
        acc = 0
        for k in range(count):
            b = k % PIPE
            if k >= PIPE:
                pev[b].synchronize()
                acc += int(pres_np[b][k % n])           # the step's results are on the host
            penvs[b].step_host_compact_async(act_rows[k % nrows], pres[b])
            pev[b].record(cur)
        torch.cuda.synchronize(dev_t)
        return acc
    pipelined(4 * PIPE)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev_t)
    t0 = time.perf_counter()
    pipe_checksum = pipelined(KE * PIPE)
    e2e_pipe_s = time.perf_counter() - t0
    for e in penvs:
        e.check()
    del env4, penvs

    # max over ranks
    if world > 1:
        v = torch.tensor([elapsed_ms, e2e_s, warm_ms, wall, e2e_wide_s, flushed_ms, prod_ms, e2e_packed_s, e2e_pipe_s], dtype=torch.float64, device=dev_t)
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        elapsed_ms, e2e_s, warm_ms, wall, e2e_wide_s, flushed_ms, prod_ms, e2e_packed_s, e2e_pipe_s = [float(x) for x in v]
    if rank != 0:
        return 0

    total_envs = n * world
    value = total_envs * K / (elapsed_ms * 1e-3)
    e2e_value = total_envs * KE / e2e_s
    peak, peak_src = measured_peak()
    kernel_ms = elapsed_ms / K                          # average launch duration over the timed region
    achieved = ALGO_BYTES_PER_ENV_STEP * n / (kernel_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": elapsed_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": config_dict(total_envs, "dp%d (independent env shards, no data-path collective; NCCL all-reduce of the 64-byte "
                              "statistics vector every %d launches on a side stream)" % (world, STATS_EVERY_CHUNKS * 64) if world > 1
                              else "single GPU",
                              "steady state of the reference's workload: fresh envs + reset() + %d untimed steps per env, "
                              "then W warm-up and K timed launches (the reference's training run -- one reset(), 50,000 steps, "
                              "`done` never true -- leaves the productive regime after ~100 steps; SURVEY.md section 8d cfg 2: "
                              "'throughput: steady state; report both regimes separately')" % BURN_IN_STEPS),
        "regimes": {"steady_state_env_steps_per_s": value,
                    "productive_env_steps_per_s": total_envs / (prod_ms * 1e-3),
                    "productive_note": "same round-robin protocol, the first %d steps of every fresh env (%d launches): queues "
                                       "hold packets that fit the windows, 1-10 data transmissions per step" % (PRODUCTIVE_STEPS, PRODUCTIVE_STEPS * M),
                    "l2_warm_one_batch_env_steps_per_s": total_envs / (warm_ms * 1e-3),
                    "l2_warm_note": "ONE batch stepped 64x back to back from a CUDA graph (its state stays in L2)",
                    "round1_protocol_env_steps_per_s": total_envs / (flushed_ms * 1e-3),
                    "round1_protocol_note": "one batch, 256 MiB memset before every launch, CUDA-event pair around every launch "
                                            "(includes the launch latency behind the flush)",
                    "per_chunk_ms_per_launch": [float(x) for x in per_step_ms]},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic_per_launch(), "peak_source": peak_src,
                     "kernel": "step_kernel<MODE_R,3,2,0>", "algorithmic_bytes_per_launch": ALGO_BYTES_PER_ENV_STEP * n,
                     "avg_launch_ms": kernel_ms,
                     "note": "mode-R state is ~190 B/env-step: the fused step kernel is latency / fp64-ALU bound, "
                             "not HBM bound (SURVEY.md 8d); the fraction is reported as required"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 2 * n, "d2h_bytes_per_step": 4 * n,
                "steps": KE, "api": "CounterTrafficEnv.step_host_compact -> gw_step_host_compact (pinned host buffers: uint8 "
                                    "actions [n][2] in, one packed uint32 {obs:17, reward+16:5, done:1} per env out; the kernel "
                                    "reads / writes the pinned buffers in place over the host link -- the h2d / d2h bytes are moved "
                                    "by the kernel's own loads and stores, inside the timed region); steady-state envs (burn-in as above)",
                "reward_checksum": checksum,
                "pipelined": {"value": total_envs * KE * PIPE / e2e_pipe_s, "h2d_bytes_per_step": 2 * n, "d2h_bytes_per_step": 4 * n,
                              "batches_in_flight": PIPE, "checksum": pipe_checksum,
                              "api": "gw_step_host_compact_async on %d independent env batches (one handle each): the host waits "
                                     "for a batch's previous results, reads them and submits its next actions while the other "
                                     "batches' steps run -- the throughput form of the same host-buffer path" % PIPE},
                "packed_api": {"value": total_envs * KE / e2e_packed_s, "h2d_bytes_per_step": 8 * n, "d2h_bytes_per_step": 9 * n,
                               "api": "CounterTrafficEnv.step_host_packed -> gw_step_host_packed (int32 actions [2][n] in, "
                                      "int32 obs | float32 reward | uint8 done out)"},
                "wide_api": {"value": total_envs * KE / e2e_wide_s, "h2d_bytes_per_step": 8 * n, "d2h_bytes_per_step": 17 * n,
                             "api": "CounterTrafficEnv.step_host -> gw_step_host (int64 obs, float64 reward, uint8 done)"}},
        "gpu_launches": K + (((len(chunk_cnt) + STATS_EVERY_CHUNKS - 1) // STATS_EVERY_CHUNKS) if world > 1 else 0),   # step kernels (+ statistics copies when sharded)
        "clocks": clocks,
        "wall_s_timed_loop": wall,
    }
    if world == 1 and not args.no_extras:
        del env2, env3
        torch.cuda.empty_cache()
        try:
            line["mask_scan"] = mask_scan_roofline(dev_t, peak, "random")
            torch.cuda.empty_cache()
            line["mask_scan_sequential_rows"] = mask_scan_roofline(dev_t, peak, "sequential")
            torch.cuda.empty_cache()
            line["cfg3_long_packet_mode_m"] = cfg3_long_packet(dev_t)
            torch.cuda.empty_cache()
            line["cfg4_multiband"] = cfg4_multiband(dev_t)
            torch.cuda.empty_cache()
            line["cfg5_pendulum"] = cfg5_pendulum(dev_t)
        except Exception as exc:                      # extras must never take the headline down
            line["extras_error"] = repr(exc)
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_run(args.cpu_seconds)
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=512)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the mask-scan roofline and the cfg-3 run")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return reference_arm(args, rank)
    if world > 1:
        from gymwipe_b200.distributed import init_from_env
        init_from_env("nccl")
    try:
        return own_arm(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            if dist.is_initialized():
                dist.destroy_process_group()


if __name__ == "__main__":
    sys.exit(main())
