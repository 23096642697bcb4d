#!/usr/bin/env python
"""
bench.py -- env-steps/sec of the batched CounterTrafficEnv hot path (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Own arm: N ranks (one per GPU; torchrun supplies RANK / LOCAL_RANK / WORLD_SIZE for N > 1), each owning
an independent POPULATION of `--batches` (default 384) batches of 65,536 envs (BASELINE configs[1]; weak
scaling).  One bench "step" is one `env.step` of EVERY env of the population: one launch of the fused step
kernel per batch, the batches one after the other -- a batch is touched again only after all the others
(gigabytes of other traffic), so every launch finds its inputs in HBM, not in L2 ("inputs larger than L2",
no flush).  A step is ~2.8 ms of device time, so even the driver's `--steps 20` times > 50 ms.  W warm-up
steps, then EXACTLY K timed steps replayed from CUDA graphs (one per step), one CUDA-event pair around them
on the launching stream, barrier + synchronize on both sides, max over ranks.  `e2e` is the same metric
through the C ABI with pinned HOST buffers (`gw_step_host_compact_many`: actions read / results written in
place by the kernels, inside the timed region), timed for >= 60 ms.

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
POPULATION_BATCHES = 384               # 65,536-env batches per GPU: one bench step = one env.step of all of them (~2.5 ms)
E2E_BATCHES = 64                       # batches of the population stepped per host-buffer call (gw_step_host_compact_many)
BURN_IN_STEPS = 128                    # steps of every env before the timed region (steady state, see own_arm)
PRODUCTIVE_STEPS = 32                  # steps of every fresh env timed separately (productive regime)
ALGO_BYTES_PER_ENV_STEP = 193          # SURVEY.md section 8d / DESIGN.md section 6
STATS_EVERY_STEPS = 4                  # bench steps between two NCCL reductions of the statistics vector
METRIC = "env-steps/sec CounterTrafficEnv batch"
UNIT = "env-steps/s"


def config_dict(n_envs_total, batches, parallelism, regime):
    return {"workload": "CounterTrafficEnv default scenario (2 counter senders + RRM, 1 FrequencyBand, FSPL, BPSK), "
                        "mode R (reference-exact accounting), batches of %d envs (BASELINE configs[1]), random "
                        "actions (device~U{0,1}, duration~U{0..19}); one step = one env.step of the GPU's whole env "
                        "population = one launch of the fused step kernel per batch" % ENVS_PER_GPU,
            "n_envs": n_envs_total, "envs_per_launch": ENVS_PER_GPU,
            "batches_per_gpu": batches, "parallelism": parallelism,
            "regime": regime,
            "l2": "inputs larger than L2: the %d batches of a GPU are stepped one after the other, so a batch's state "
                  "(~13 MB hot), its action rows and outputs are re-touched only after ~%.1f GB of other traffic "
                  "(L2 = 126 MB); no flush, steps replayed from CUDA graphs, one CUDA-event pair around the K timed "
                  "steps; EnvPopulation.step spreads the (independent) batches of a step over a few streams" % (batches, 13e-3 * (batches - 1))}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic_per_launch(key="dram_bytes_per_launch"):
    """DRAM bytes per step-kernel launch from the committed ncu capture (profiles/), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "step_kernel_traffic.json")) as f:
            d = json.load(f)
        return d.get(key)
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


def _time_steps(env, a_dev, a_dur, steps, dev_t, warm=4):
    """`steps` device-resident env.step calls behind `warm` untimed ones; returns ms per step (CUDA events)."""
    import torch
    stream = torch.cuda.current_stream(dev_t)
    for t in range(warm):
        env.step({"device": a_dev[t], "duration": a_dur[t]})
    env.stats()
    torch.cuda.synchronize(dev_t)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for t in range(warm, warm + steps):
        env.step({"device": a_dev[t], "duration": a_dur[t]})
    e1.record(stream)
    torch.cuda.synchronize(dev_t)
    env.check()
    return e0.elapsed_time(e1) / steps


def cfg3_long_packet(dev_t, peak, rank=0, world=1, steps=48):
    """
    BASELINE configs[2]: 1500-byte payloads, fed per-bit masks (mode M), a PHY-only interferer creating
    mid-packet SINR segments; ASSIGNMENT_DURATION_FACTOR = 10000 so that windows fit 122 ms packets.
    The HBM-bound part of the path is the popcount over the mask words.  `gw_set_masks` does it in ONE streaming
    pass over the fed buffer (`mask_index_kernel`: per-row prefix counts) -- that kernel's roofline is the
    `roofline` entry; the fused step kernel then takes every section's count from two index entries and at most
    two 16-byte groups per row (`step`), instead of scanning ~6 KB of mask words per env-step inside the event
    loop (`GW_FED_INDEX=0`: the in-step scan of the previous rounds, reported by `profiles/`).
    """
    import torch
    import gymwipe_b200
    n, slots, words = 65536, 2, 512
    sc = {"assignment_duration_factor": 10000, "bands": [{"frequency": 2.4e9, "bandwidth": 22e6, "devices": [
        {"role": "sender", "x": 0.0, "y": 2.0, "mult": 1, "payload": 1500, "interval": 0.001, "dest": 1},
        {"role": "sender", "x": 0.0, "y": -2.0, "mult": 3, "payload": 1500, "interval": 0.001, "dest": 0},
        {"role": "rrm", "x": 0.0, "y": 0.0},
        {"role": "jammer", "x": 6.0, "y": 0.0, "interval": 0.05, "delay": 0.003, "power": 0.0, "hdr": 13, "payload": 200}]}]}
    env = gymwipe_b200.make('CounterTraffic-v0', num_envs=n, device=dev_t, scenario=sc, mode="mask_fed",
                            env_id_offset=rank * n, strict=False)
    # Bernoulli(1/16) masks: 4 GiB, resident before the timed region (>> L2: every pass streams from HBM)
    shape = (n, 1, 4, slots, 4, words)
    g = torch.Generator(device=dev_t).manual_seed(77 + rank)
    masks = torch.randint(-2 ** 31, 2 ** 31 - 1, shape, dtype=torch.int32, device=dev_t, generator=g)
    for _ in range(3):                                  # AND of 4 random words: bit density 1/16
        masks &= torch.randint(-2 ** 31, 2 ** 31 - 1, shape, dtype=torch.int32, device=dev_t, generator=g)
    stream = torch.cuda.current_stream(dev_t)
    env.set_masks(masks, slots)                         # allocates the index; untimed
    torch.cuda.synchronize(dev_t)
    indexed = os.environ.get("GW_FED_INDEX", "1")[:1] != "0"
    reps = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        env.set_masks(masks, slots)                     # one launch of mask_index_kernel over the 4 GiB
    e1.record(stream)
    torch.cuda.synchronize(dev_t)
    idx_ms = e0.elapsed_time(e1) / reps
    rows = n * 4 * slots * 4
    idx_read = int(masks.numel() * 4)
    idx_write = rows * (((words // 4) + 1 + 7) // 8 * 8) * 2
    env.reset()
    a_dev = torch.randint(0, 2, (steps + 4, n), generator=g, device=dev_t, dtype=torch.int32)
    a_dur = torch.randint(12, 20, (steps + 4, n), generator=g, device=dev_t, dtype=torch.int32)
    for t in range(4):
        env.step({"device": a_dev[t], "duration": a_dur[t]})
    env.mask_bytes(clear=True)
    ms = _time_steps(env, a_dev, a_dur, steps, dev_t, warm=0)
    mask_bytes = env.mask_bytes() / steps
    st = env.stats().cpu().numpy()
    algo = mask_bytes + ALGO_BYTES_PER_ENV_STEP * n
    out = {"workload": "configs[2]: 1500-byte payloads, fed per-bit masks (mode M), PHY-only interferer, %d envs per GPU" % n,
           "n_envs": n, "env_steps_per_s": n / (ms * 1e-3), "ms_per_step": ms, "transmissions_per_step": float(st[6]) / steps,
           "deliveries_per_step": float(st[1] + st[2]) / steps, "mask_bytes_resident": idx_read}
    if indexed:
        achieved = (idx_read + idx_write) / (idx_ms * 1e-3) / 1e9
        out["roofline"] = {"bound": "hbm", "kernel": "mask_index_kernel (gw_set_masks): the popcount over the fed mask words, one "
                                                      "streaming pass per feed, per-row prefix counts out",
                           "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                           "algorithmic_bytes_per_launch": idx_read + idx_write, "bytes_read": idx_read,
                           "bytes_written": idx_write, "avg_launch_ms": idx_ms,
                           "traffic": ncu_traffic_per_launch("mask_index_dram_bytes_per_launch"),
                           "note": "every mask word is read once per gw_set_masks; the step kernels read two index "
                                   "entries and at most two 16-byte groups per receiver and decided section"}
        out["step"] = {"kernel": "step_kernel<MODE_M_FEDX,4,2,1> (event loop + index look-ups)", "ms_per_step": ms,
                       "decided_mask_bytes_per_step": mask_bytes,
                       "decided_mask_gbs": algo / (ms * 1e-3) / 1e9,
                       "traffic": ncu_traffic_per_launch("cfg3_dram_bytes_per_launch"),
                       "steps_amortising_one_feed": idx_ms / ms,
                       "note": "decided_mask_bytes = the 32-bit mask words that hold the on-air bits of every decided section "
                               "(what an in-step scan has to read; rounds 1-2 read them at 780 GB/s, 0.54 ms per step); "
                               "decided_mask_gbs is that figure over the step time -- an equivalent rate, not traffic"}
    else:
        achieved = algo / (ms * 1e-3) / 1e9
        out["roofline"] = {"bound": "hbm", "kernel": "step_kernel<MODE_M_FED,4,2,1> (fused step incl. the in-step mask scan)",
                           "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                           "algorithmic_bytes_per_launch": algo, "mask_bytes_per_launch": mask_bytes,
                           "avg_launch_ms": ms, "traffic": ncu_traffic_per_launch("cfg3_scan_dram_bytes_per_launch"),
                           "note": "algorithmic bytes = 32-bit mask words holding the on-air bits of every decided "
                                   "section (counted by the kernel) + 193 B of state per env-step; masks 4 GiB >> L2"}
    return out


def cfg4_multiband(dev_t, rank=0, world=1, steps=64):
    """
    BASELINE configs[3], one GPU's share of the 1 M envs: 16 devices over 4 FrequencyBands (per band: RRM + 2 MAC
    senders + 1 PHY-only interferer), per-env positions (senders and RRM within a few metres of each other so that
    packets are delivered, the interferer anywhere in a 40 m square), one action per band; the env's clock ends at
    the latest band (mode R).  131 072 envs x 4 bands = 524 288 band-sims per GPU.
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
    g = torch.Generator(device=dev_t).manual_seed(11 + rank)
    pos = torch.rand((n, 4, 4, 2), generator=g, device=dev_t, dtype=torch.float64)
    pos[:, :, :3, :] = pos[:, :, :3, :] * 3.0 - 1.5    # senders and RRM: ~U(-1.5, 1.5) m
    pos[:, :, 3, :] = pos[:, :, 3, :] * 40.0 - 20.0    # interferer: ~U(-20, 20) m
    env = gymwipe_b200.make('CounterTraffic-v0', num_envs=n, device=dev_t, scenario=sc, positions=pos,
                            env_id_offset=rank * n, strict=False)
    env.reset()
    a_dev = torch.randint(0, 2, (steps + 4, n, 4), generator=g, device=dev_t, dtype=torch.int32)
    a_dur = torch.randint(0, 20, (steps + 4, n, 4), generator=g, device=dev_t, dtype=torch.int32)
    ms = _time_steps(env, a_dev, a_dur, steps, dev_t)
    st = env.stats().cpu().numpy()
    return {"workload": "configs[3], share of one GPU: %d envs x 4 bands x 4 devices, per-env positions, mode R; the first "
                        "%d steps of fresh envs (productive regime)" % (n, steps + 4),
            "n_envs": n, "env_steps_per_s": n / (ms * 1e-3), "band_steps_per_s": 4 * n / (ms * 1e-3), "ms_per_step": ms,
            "transmissions_per_env_step": float(st[6]) / steps / n, "deliveries_per_env_step": float(st[1] + st[2]) / steps / n}


def cfg5_pendulum(dev_t, rank=0, world=1, steps=32):
    """
    BASELINE configs[4], one GPU's share of the 1 M envs: the networked inverted-pendulum env (sensor / controller
    band assignment, in-kernel RK4 plant; PARITY UNPINNED), 131 072 envs per GPU.
    """
    import torch
    import gymwipe_b200
    n = 131072
    env = gymwipe_b200.make('InvertedPendulum-v0', num_envs=n, device=dev_t, strict=False)
    g = torch.Generator(device=dev_t).manual_seed(5 + rank)
    a_dev = torch.randint(0, 2, (steps + 4, n), generator=g, device=dev_t, dtype=torch.int32)
    a_dur = torch.randint(1, 20, (steps + 4, n), generator=g, device=dev_t, dtype=torch.int32)
    ms = _time_steps(env, a_dev, a_dur, steps, dev_t)
    th = env.plant_state()[2]
    return {"workload": "configs[4], share of one GPU: networked inverted pendulum, %d envs, in-kernel RK4 plant "
                        "(parity unpinned: the reference env is unconstructible)" % n,
            "n_envs": n, "env_steps_per_s": n / (ms * 1e-3), "ms_per_step": ms,
            "mean_abs_angle_deg": float(torch.rad2deg(th).abs().mean())}


def general_band(dev_t, rank=0, world=1, steps=24, n=65536, ns=8, nj=4, mode="reference", with_mode_m=True):
    """
    The general band engine (gw_band.cuh): bands beyond CounterTrafficEnv's template --
    8 MAC senders on a circle of 2 m around the RRM, each addressing its neighbour, 4 PHY-only interferers at 5 m,
    run-time device counts, state in global memory.  65 536 envs per GPU, the first steps of fresh envs.
    """
    import math
    import torch
    import gymwipe_b200
    devs = []
    for k in range(ns):
        a = 2 * math.pi * k / ns
        devs.append({"role": "sender", "x": 2.0 * math.cos(a), "y": 2.0 * math.sin(a), "mult": 1 + k % 3, "payload": "counter",
                     "interval": 0.001, "dest": (k + 1) % ns})
    devs.append({"role": "rrm", "x": 0.0, "y": 0.0})
    for j in range(nj):
        a = 2 * math.pi * (j + 0.5) / nj
        devs.append({"role": "jammer", "x": 5.0 * math.cos(a), "y": 5.0 * math.sin(a), "interval": 0.011 + 0.003 * j,
                     "delay": 0.001 * j, "power": 10.0, "hdr": 13, "payload": 60})
    sc = {"assignment_duration_factor": 1000, "bands": [{"frequency": 2.4e9, "bandwidth": 22e6, "devices": devs}]}
    env = gymwipe_b200.make('CounterTraffic-v0', num_envs=n, device=dev_t, scenario=sc, strict=False, mode=mode, seed=99,
                            env_id_offset=rank * n)
    env.reset()
    g = torch.Generator(device=dev_t).manual_seed(17 + rank)
    a_dev = torch.randint(0, ns, (steps + 4, n), generator=g, device=dev_t, dtype=torch.int32)
    a_dur = torch.randint(0, 20, (steps + 4, n), generator=g, device=dev_t, dtype=torch.int32)
    stream = torch.cuda.current_stream(dev_t)
    for t in range(4):
        env.step({"device": a_dev[t], "duration": a_dur[t]})
    torch.cuda.synchronize(dev_t)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for t in range(4, 4 + steps):
        env.step({"device": a_dev[t], "duration": a_dur[t]})
    e1.record(stream)
    torch.cuda.synchronize(dev_t)
    env.check()
    ms = e0.elapsed_time(e1) / steps
    out = {"workload": "general band engine: %d envs x (%d MAC senders + RRM + %d PHY-only senders), mode %s, the first %d steps of "
                       "fresh envs" % (n, ns, nj, "R" if mode == "reference" else "M (Philox masks)", steps + 4),
           "n_envs": n, "env_steps_per_s": n / (ms * 1e-3), "ms_per_step": ms,
           "transmissions_per_env_step": float(env.transmissions().sum()) / (steps + 4) / n,
           "deliveries_per_env_step": float(env.delivered().sum()) / (steps + 4) / n}
    env.close()
    del env
    torch.cuda.empty_cache()
    if with_mode_m and mode == "reference":
        m = general_band(dev_t, rank, world, steps=8, n=n, ns=ns, nj=nj, mode="mask_philox")
        out["mode_m_philox"] = {k: m[k] for k in ("env_steps_per_s", "ms_per_step", "deliveries_per_env_step")}
    return out


def large_batch(dev_t, rank=0, world=1, n=1048576, batches=8, steps=6):
    """
    The headline workload with 1,048,576 envs per env object / launch instead of 65,536 (the kernel's grid is capped at
    one resident wave and strides over the batch, so a large launch has no partial wave and overlaps its own state
    loads): a population of `batches` such batches, steady state (128 untimed steps), population steps over the
    streams of EnvPopulation.step replayed from CUDA graphs.  Reported next to the headline, which stays on
    BASELINE configs[1]'s 65,536-env batches.
    """
    import torch
    import gymwipe_b200
    from gymwipe_b200.envs import EnvPopulation
    pop = EnvPopulation([gymwipe_b200.make('CounterTraffic-v0', num_envs=n, device=dev_t, env_id_offset=(rank * batches + b) * n,
                                           strict=False) for b in range(batches)])
    pop.reset()
    g = torch.Generator(device=dev_t).manual_seed(77 + rank)
    ROWS = 13
    a_dev = torch.randint(0, 2, (ROWS, n), generator=g, device=dev_t, dtype=torch.int32)
    a_dur = torch.randint(0, 20, (ROWS, n), generator=g, device=dev_t, dtype=torch.int32)
    stream = torch.cuda.Stream(device=dev_t)
    counter = [0]

    def pop_step():
        j = counter[0]
        pop.step([{"device": a_dev[(j + b) % ROWS], "duration": a_dur[(j + b) % ROWS]} for b in range(batches)])
        counter[0] = j + batches

    with torch.cuda.stream(stream):
        for _ in range(128):
            pop_step()
        graphs = []
        for _ in range(steps):
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=stream):
                pop_step()
            graphs.append(gr)
        for gr in graphs:
            gr.replay()                                  # graph upload is warm-up
        torch.cuda.synchronize(dev_t)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(3):
            for gr in graphs:
                gr.replay()
        e1.record(stream)
        torch.cuda.synchronize(dev_t)
    pop.check()
    ms = e0.elapsed_time(e1) / (3 * steps)               # per population step
    peak, _ = measured_peak()
    us_launch = 1e3 * ms / batches
    out = {"workload": "the headline workload with %d envs per launch: population of %d batches, steady state, %d population "
                       "steps from CUDA graphs x 3" % (n, batches, steps),
           "n_envs": n * batches, "envs_per_launch": n, "env_steps_per_s": n * batches / (ms * 1e-3), "ms_per_step": ms,
           "us_per_launch": us_launch,
           "roofline_frac": ALGO_BYTES_PER_ENV_STEP * n / (us_launch * 1e-6) / 1e9 / peak}
    pop.close()
    del pop, graphs
    torch.cuda.empty_cache()
    return out


def grid_benchmark(dev_t, rank=0, world=1, n_envs=4096, n_devices=20):
    """
    The reference's own benchmark (tests/test_benchmark.py:52-91, `make benchmark`): a grid of 20 PHY-only
    SendingDevices, static and with the mobility processes, advanced by ONE simulated second -- here for
    `n_envs` independent grids at once.  BASELINE.md (session-measured, one core): 1.44 s / 4.68 s of wall time
    per simulated second and grid for the static / mobile variant of the Python reference.
    """
    import torch
    from gymwipe_b200.envs import SendingDeviceGrid
    out = {"workload": "reference benchmark grid: %d independent grids x %d PHY-only senders (40 dBm, 10 ms send "
                       "interval), runSimulation(1.0)" % (n_envs, n_devices), "n_envs": n_envs}
    stream = torch.cuda.current_stream(dev_t)
    for label, mobile in (("static", False), ("mobile", True)):
        grid = SendingDeviceGrid(n_envs, n_devices, device=dev_t, mobile=mobile, max_moves=1002, seed=100 + rank)
        grid.runSimulation(0.01)                        # warm-up (module load, first touch)
        torch.cuda.synchronize(dev_t)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        grid.runSimulation(0.99)
        e1.record(stream)
        torch.cuda.synchronize(dev_t)
        faulted = int((grid.faults() != 0).sum())        # grids in which the reference would raise (assert noisePower >= 0)
        ms = e0.elapsed_time(e1)
        st = grid.stats()
        out[label] = {"device_seconds_per_simulated_second": ms * 1e-3 / 0.99,
                      "grid_seconds_per_second": n_envs * 0.99 / (ms * 1e-3),
                      "transmissions_per_grid": float(st[0].sum()) / n_envs,
                      "payloads_decoded_per_grid": float(st[3].sum()) / n_envs, "grids_where_the_reference_raises": faulted,
                      "reference_seconds_per_simulated_second_one_grid": 4.68 if mobile else 1.44}
        grid.close()
        del grid
        torch.cuda.empty_cache()
    out["ms_per_step"] = 1e3 * out["mobile"]["device_seconds_per_simulated_second"]
    return out


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
            "config": config_dict(ENVS_PER_GPU * args.batches * args.gpus, args.batches, "host threads x%d" % threads,
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
    from gymwipe_b200.envs import EnvPopulation

    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.set_num_threads(1)                            # no idle worker threads spinning next to the stepping thread
    torch.cuda.set_device(local_rank)
    dev_t = torch.device("cuda", local_rank)
    K, W = args.steps, args.warmup
    n = ENVS_PER_GPU
    M = args.batches
    # The rank's env POPULATION: M independent batches of 65,536 envs (BASELINE configs[1]).  One bench
    # "step" = one env.step of EVERY env of the population = M launches of the fused step kernel, one per
    # batch.  A batch is touched again only after the M - 1 others (M x ~13 MB of hot state + fresh action
    # rows + outputs >> the 126 MB L2), so every launch finds its inputs in HBM ("inputs larger than L2");
    # nothing is flushed and nothing but step kernels sits inside the timed region.
    pop = EnvPopulation([gymwipe_b200.make('CounterTraffic-v0', num_envs=n, device=dev_t,
                                           env_id_offset=(rank * M + b) * n, strict=False) for b in range(M)])
    envs = pop.envs
    pop.reset()

    # synthetic action rows, resident in HBM before the timed region: a pool of ROWS independent rows;
    # launch j (batch j % M of bench step j // M) reads row j % ROWS
    g = torch.Generator(device=dev_t).manual_seed(1234 + rank)
    ROWS = 1021                                         # prime: successive steps of a batch see different rows
    a_dev = torch.randint(0, 2, (ROWS, n), generator=g, device=dev_t, dtype=torch.int32)
    a_dur = torch.randint(0, 20, (ROWS, n), generator=g, device=dev_t, dtype=torch.int32)
    reducer = StatsReducer(dev_t) if world > 1 else None
    stream = torch.cuda.Stream(device=dev_t)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    counter = [0]                                       # launches so far: launch j steps batch j % M
    pop.streams = args.streams

    def pop_step():
        """One env.step of the whole population through EnvPopulation.step (its batches spread over a few streams)."""
        j = counter[0]
        pop.step([{"device": a_dev[(j + b) % ROWS], "duration": a_dur[(j + b) % ROWS]} for b in range(M)])
        counter[0] = j + M

    def capture_steps(count):
        """One CUDA graph per bench step (M launches; the action-row pointers are baked in); capturing does
        not execute."""
        out = []
        for _ in range(count):
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=stream):
                pop_step()
            out.append(gr)
        return out

    def timed(fn):
        torch.cuda.synchronize(dev_t)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        torch.cuda.synchronize(dev_t)
        return e0.elapsed_time(e1)

    torch.cuda.synchronize(dev_t)
    with torch.cuda.stream(stream):
        # (1) productive regime (the first ~100 steps after construction / reset(): queues hold packets
        # that fit the windows): PRODUCTIVE_STEPS steps of every fresh env, timed for the record
        gp = capture_steps(PRODUCTIVE_STEPS)
        prod_ms = timed(lambda: [gr.replay() for gr in gp]) / (PRODUCTIVE_STEPS * M)
        del gp
        # (2) burn-in to the steady state of the reference's workload: its training run
        # (agents/dqn_counter_traffic.py: one reset(), dqn.fit(nb_steps=50000), `done` never true) leaves
        # the productive regime after ~100 steps and spends > 99 % of its steps in the regime where the
        # counters are too large for any window (announcements only)
        for _ in range(BURN_IN_STEPS - PRODUCTIVE_STEPS):
            pop_step()
        # (3) W warm-up steps, then EXACTLY K timed steps.  The K steps are replayed from G = min(K, 16)
        # step graphs used round-robin (each with its own action rows); every graph is replayed once untimed
        # first (the first launch of an instantiated graph uploads it to the device)
        for _ in range(W):
            pop_step()
        torch.cuda.synchronize(dev_t)
        G = min(K, 16)
        graphs = capture_steps(G)
        for gr in graphs:
            gr.replay()
        torch.cuda.synchronize(dev_t)

        def reduce_stats():
            # K5 partial sums of the whole population -> NCCL all-reduce on a side stream
            pop.stats(out=reducer.next_slot())
            reducer.submit()
        if reducer is not None:
            # first use loads the small kernels and sets up NCCL's channels: not part of stepping
            reduce_stats()
            reducer.drain()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev_t)
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
        n_reduce = 0
        if sampler is not None:
            sampler.mark_begin()
        wall0 = time.perf_counter()
        marks[0].record(stream)
        for k in range(K):
            graphs[k % G].replay()
            marks[k + 1].record(stream)                 # the step's end, BEFORE the host enqueues the reduction
            if reducer is not None and ((k + 1) % STATS_EVERY_STEPS == 0 or k + 1 == K):
                reduce_stats()
                n_reduce += 1
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev_t)
        wall = time.perf_counter() - wall0
    if sampler is not None:
        sampler.mark_end()
    clocks = sampler.stop() if sampler is not None else None
    pop.check()
    if reducer is not None:
        reducer.drain()
    step_ms = np.array([marks[k].elapsed_time(marks[k + 1]) for k in range(K)])
    elapsed_ms = float(marks[0].elapsed_time(marks[-1]))

    # transparency: one batch stepped back to back from one graph (its state stays in L2)
    env1 = envs[0]
    KB = 64
    with torch.cuda.stream(stream):
        gw = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gw, stream=stream):
            for k in range(KB):
                env1.step({"device": a_dev[k], "duration": a_dur[k]})
        gw.replay()
        warm_ms = timed(gw.replay) / KB
    del graphs, gw

    # ---- e2e: the same population step through the C ABI with HOST buffers (gw_step_host_compact_many): every
    # batch reads its pinned uint8 actions and writes its pinned result words in place; one synchronisation per
    # population step.  Steady-state envs (the population above); KE steps so that the region is >= ~60 ms.
    PE = min(M, args.e2e_batches)
    epop = EnvPopulation(envs[:PE])
    EROWS = 8
    h_act = [[torch.stack([a_dev[(r * PE + b) % ROWS], a_dur[(r * PE + b) % ROWS]], dim=1).to(torch.uint8).cpu().pin_memory()
              for b in range(PE)] for r in range(EROWS)]
    h_res = [torch.empty(n, dtype=torch.int32).pin_memory() for _ in range(PE)]
    act_ptrs = [EnvPopulation.pointer_array(h_act[r]) for r in range(EROWS)]
    res_ptrs = EnvPopulation.pointer_array(h_res)
    # the smallest wire format (gw_step_host_tiny_many): one action byte in, one 16-bit result word out per env
    t_act = [[((a_dev[(r * PE + b) % ROWS] << 7) | a_dur[(r * PE + b) % ROWS]).to(torch.uint8).cpu().pin_memory()
              for b in range(PE)] for r in range(EROWS)]
    t_res = [torch.empty(n, dtype=torch.int16).pin_memory() for _ in range(PE)]
    tact_ptrs = [EnvPopulation.pointer_array(t_act[r]) for r in range(EROWS)]
    tres_ptrs = EnvPopulation.pointer_array(t_res)
    res_np = [r.numpy() for r in t_res]

    def e2e_leg(step_fn, a_ptrs, r_ptrs, rnp):
        for r in range(max(W, 3) + EROWS):
            step_fn(a_ptrs[r % EROWS], r_ptrs)
        torch.cuda.synchronize(dev_t)
        t0 = time.perf_counter()
        for r in range(8):
            step_fn(a_ptrs[r % EROWS], r_ptrs)
        est = (time.perf_counter() - t0) / 8
        ke = max(K, int(0.06 / max(est, 1e-6)) + 1)
        if world > 1:
            kt = torch.tensor([ke], dtype=torch.int64, device=dev_t)
            dist.all_reduce(kt, op=dist.ReduceOp.MAX)
            ke = int(kt[0])
            dist.barrier()
        torch.cuda.synchronize(dev_t)
        chk = 0
        t0 = time.perf_counter()
        for r in range(ke):
            step_fn(a_ptrs[r % EROWS], r_ptrs)
            chk += int(rnp[r % PE][r % n])              # the step's results are on the host
        torch.cuda.synchronize(dev_t)
        return ke, time.perf_counter() - t0, chk

    KC, e2e_compact_s, _ = e2e_leg(epop.step_host_compact, act_ptrs, res_ptrs, [r.numpy() for r in h_res])
    KE, e2e_s, checksum = e2e_leg(epop.step_host_tiny, tact_ptrs, tres_ptrs, res_np)
    epop.check()
    reward_checksum = float(sum(env1.unpack_tiny(h)[1].sum() for h in t_res))
    # the single-batch synchronous call (gw_step_host_compact) and the wider-typed variants, for the record
    h_obs = torch.empty(n, dtype=torch.int64).pin_memory()
    h_rew = torch.empty(n, dtype=torch.float64).pin_memory()
    h_done = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_dev = a_dev[:EROWS].cpu().pin_memory()
    h_dur = a_dur[:EROWS].cpu().pin_memory()
    h_pk = torch.stack([h_dev, h_dur], dim=1).contiguous().pin_memory()         # [rows, 2, n]
    h_res9 = torch.empty(9 * n, dtype=torch.uint8).pin_memory()
    KS = max(64, min(KE, 2048))

    def time_single(fn):
        for r in range(3):
            fn(r)
        torch.cuda.synchronize(dev_t)
        t0 = time.perf_counter()
        for r in range(KS):
            fn(r)
        torch.cuda.synchronize(dev_t)
        return time.perf_counter() - t0
    e2e_single_s = time_single(lambda r: env1.step_host_compact(h_act[r % EROWS][0], h_res[0]))
    e2e_packed_s = time_single(lambda r: env1.step_host_packed(h_pk[r % EROWS], h_res9))
    e2e_wide_s = time_single(lambda r: env1.step_host(h_dev[r % EROWS], h_dur[r % EROWS], h_obs, h_rew, h_done))
    assert torch.equal(env1.unpack_results(h_res9)[0].to(torch.int64), h_obs)
    env1.check()

    # max over ranks
    if world > 1:
        v = torch.tensor([elapsed_ms, e2e_s, warm_ms, wall, e2e_wide_s, prod_ms, e2e_packed_s, e2e_single_s, e2e_compact_s],
                         dtype=torch.float64, device=dev_t)
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        elapsed_ms, e2e_s, warm_ms, wall, e2e_wide_s, prod_ms, e2e_packed_s, e2e_single_s, e2e_compact_s = [float(x) for x in v]
    pop.close()
    del pop, epop, envs
    torch.cuda.empty_cache()

    extras = {}
    if not args.no_extras:
        peak, _ = measured_peak()
        for name, fn in (("cfg3_long_packet_mode_m", lambda: cfg3_long_packet(dev_t, peak, rank, world)),
                         ("cfg4_multiband", lambda: cfg4_multiband(dev_t, rank, world)),
                         ("cfg5_pendulum", lambda: cfg5_pendulum(dev_t, rank, world)),
                         ("large_batch_1m_envs", lambda: large_batch(dev_t, rank, world)),
                         ("general_band_8_senders", lambda: general_band(dev_t, rank, world)),
                         ("reference_benchmark_grid", lambda: grid_benchmark(dev_t, rank, world)),
                         ("mask_scan", lambda: mask_scan_roofline(dev_t, peak, "random") if world == 1 else None)):
            try:                                          # extras must never take the headline down
                r, err = fn(), None
            except Exception as exc:
                r, err = None, repr(exc)
            torch.cuda.empty_cache()
            if world > 1:                               # every rank its share; the slowest rank bounds the step
                rl_ms = r["roofline"]["avg_launch_ms"] if r and "roofline" in r else 0.0
                t = torch.tensor([r["ms_per_step"] if r else float("inf"), rl_ms], dtype=torch.float64, device=dev_t)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                if r and float(t[0]) != float("inf"):
                    scale = r["ms_per_step"] / float(t[0])
                    r["ms_per_step"] = float(t[0])
                    for key in ("env_steps_per_s", "band_steps_per_s"):
                        if key in r:
                            r[key] = r[key] * scale * world
                    if "roofline" in r:                  # the roofline kernel's own launch time, slowest rank
                        rs = rl_ms / float(t[1]) if float(t[1]) > 0 else 1.0
                        r["roofline"]["achieved"] *= rs
                        r["roofline"]["frac"] *= rs
                        r["roofline"]["avg_launch_ms"] = float(t[1])
                        r["roofline"]["note"] += "; per GPU, at the slowest rank's launch time"
                    if "step" in r:
                        r["step"]["ms_per_step"] = float(t[0])
                        r["step"]["decided_mask_gbs"] *= scale
                    r["n_envs_total"] = r["n_envs"] * world
                elif r:
                    r, err = None, "another rank failed"
            if r is not None:
                extras[name] = r
            elif err is not None:
                extras[name + "_error"] = err
    if rank != 0:
        return 0

    envs_per_step = n * M * world                       # env-steps of one bench step, all ranks
    value = envs_per_step * K / (elapsed_ms * 1e-3)
    e2e_value = n * PE * world * KE / e2e_s
    peak, peak_src = measured_peak()
    kernel_ms = elapsed_ms / (K * M)                    # average launch duration over the timed region
    achieved = ALGO_BYTES_PER_ENV_STEP * n / (kernel_ms * 1e-3) / 1e9
    traffic = ncu_traffic_per_launch()
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": elapsed_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": config_dict(n * M * world, M, "dp%d (independent env shards, no data-path collective; NCCL all-reduce of the "
                              "64-byte statistics vector every %d steps on a side stream)" % (world, STATS_EVERY_STEPS) if world > 1
                              else "single GPU",
                              "steady state of the reference's workload: fresh envs + reset() + %d untimed steps per env, "
                              "then W warm-up and K timed steps (the reference's training run -- one reset(), 50,000 steps, "
                              "`done` never true -- leaves the productive regime after ~100 steps; SURVEY.md section 8d cfg 2: "
                              "'throughput: steady state; report both regimes separately')" % BURN_IN_STEPS),
        "launches_per_step": M, "us_per_launch": 1e3 * kernel_ms,
        "productive": {"value": n * world / (prod_ms * 1e-3), "unit": UNIT, "us_per_launch": 1e3 * prod_ms,
                       "note": "same population protocol, the first %d steps of every fresh env (%d launches per GPU): queues "
                               "hold packets that fit the windows, 1-10 data transmissions per step" % (PRODUCTIVE_STEPS, PRODUCTIVE_STEPS * M)},
        "regimes": {"steady_state_env_steps_per_s": value,
                    "productive_env_steps_per_s": n * world / (prod_ms * 1e-3),
                    "l2_warm_one_batch_env_steps_per_s": n * world / (warm_ms * 1e-3),
                    "l2_warm_note": "ONE batch stepped 64x back to back from a CUDA graph (its state stays in L2)",
                    "per_step_ms": {"min": float(step_ms.min()), "median": float(np.median(step_ms)), "max": float(step_ms.max())}},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src,
                     "kernel": "step_kernel<MODE_R,3,2,0>", "algorithmic_bytes_per_launch": ALGO_BYTES_PER_ENV_STEP * n,
                     "avg_launch_ms": kernel_ms,
                     "note": "mode-R state is ~190 B/env-step: the fused step kernel is latency / issue bound, "
                             "not HBM bound (SURVEY.md 8d); the HBM-bound kernel of the path is the popcount over the fed "
                             "mode-M mask words (mask_index_kernel, one streaming pass per gw_set_masks) -- see "
                             "cfg3_long_packet_mode_m.roofline"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n * PE, "d2h_bytes_per_step": 2 * n * PE,
                "steps": KE, "envs_per_step_per_gpu": n * PE, "timed_region_s": e2e_s,
                "api": "EnvPopulation.step_host_tiny -> gw_step_host_tiny_many: one env.step of a population of %d "
                       "batches x %d envs per GPU from PINNED HOST buffers (one action byte `device << 7 | duration` in, one "
                       "16-bit result word {obs - COUNTER_BOUND: 8, reward+16: 5, done: 1} per env out -- lossless: the "
                       "interpreter's observation is COUNTER_BOUND + a difference of at most +-COUNTER_BYTE_LENGTH; the kernels "
                       "read / write the pinned buffers in place over the host link -- the h2d / d2h bytes are moved by the "
                       "kernels' own loads and stores, inside the timed region); the call's launches are captured once into a "
                       "CUDA graph and replayed; one stream synchronisation per population step; steady-state envs" % (PE, n),
                "reward_checksum": reward_checksum, "result_checksum": checksum,
                "compact_api": {"value": n * PE * world * KC / e2e_compact_s, "h2d_bytes_per_step": 2 * n * PE,
                                "d2h_bytes_per_step": 4 * n * PE, "steps": KC,
                                "api": "EnvPopulation.step_host_compact -> gw_step_host_compact_many: the same call with uint8 "
                                       "actions [n][2] in and one packed uint32 {obs:17, reward+16:5, done:1} per env out"},
                "single_batch_sync": {"value": n * world * KS / e2e_single_s, "h2d_bytes_per_step": 2 * n, "d2h_bytes_per_step": 4 * n,
                                      "api": "CounterTrafficEnv.step_host_compact -> gw_step_host_compact: ONE batch per call, "
                                             "synchronised every call (launch + kernel + wake-up latency per 65,536 envs)"},
                "packed_api": {"value": n * world * KS / e2e_packed_s, "h2d_bytes_per_step": 8 * n, "d2h_bytes_per_step": 9 * n,
                               "api": "CounterTrafficEnv.step_host_packed -> gw_step_host_packed (int32 actions [2][n] in, "
                                      "int32 obs | float32 reward | uint8 done out)"},
                "wide_api": {"value": n * world * KS / e2e_wide_s, "h2d_bytes_per_step": 8 * n, "d2h_bytes_per_step": 17 * n,
                             "api": "CounterTrafficEnv.step_host -> gw_step_host (int64 obs, float64 reward, uint8 done)"}},
        "gpu_launches": K * M + n_reduce,               # step kernels (+ statistics copies when sharded)
        "clocks": clocks,
        "wall_s_timed_loop": wall,
    }
    line.update(extras)
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_run(args.cpu_seconds)
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--batches", type=int, default=POPULATION_BATCHES,
                    help="65,536-env batches per GPU (one bench step = one env.step of all of them)")
    ap.add_argument("--streams", type=int, default=3,
                    help="side streams EnvPopulation.step spreads the batches of a population step over")
    ap.add_argument("--e2e-batches", type=int, default=E2E_BATCHES,
                    help="batches stepped per host-buffer call of the e2e leg (gw_step_host_compact_many)")
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
