#!/usr/bin/env python
"""Benchmark of the NELE-GAN intelligibility-labelling hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config bench|ganround|sweep]

One "step" = one pass of the hot path (HASPI v2 + SIIB^Gauss + ESTOI, the three
labels train_nele.py:320-322 computes per utterance) over one batch of synthetic
(clean, degraded) pairs: BASELINE.json configs[2], 4096 pairs x 3 s at 16 kHz,
RMS 0.03, speech-shaped, per GPU (weak scaling: every rank scores its own 4096).
Metric: audio-seconds scored per second, whole job.

  value  inputs already resident in HBM (NELE_FLAG_DEVICE_INPUT), CUDA events on
         the launching stream, max over ranks
  e2e    the same batch through the public API with HOST (pinned) buffers: the
         host->device copy of the waveforms and the device->host copy of the
         per-pair records are inside the timed region (for N > 1 also the NCCL
         gather of the records); the upload of step k + 1 is started with
         nele_prefetch before the blocking call of step k
  roofline       dominant kernel: algorithmic bytes per launch / its CUDA-event
                 duration, against MEASURED_PEAKS.json
  cpu_baseline   the CPU oracle (a port of the reference algorithms, oracle/) on
                 the host cores, bounded sample, rank 0 at N = 1 only
  general_case   the same step on 4096 pairs of 47 999 samples.  configs[2]'s 48 000 samples are a multiple
                 of SIIB's 200-sample hop, so the wrapper's tiling (intel.py:71-75) repeats exactly and
                 every pair takes the rank-96 KLT route; any other length -- the toy data, a GAN round, the
                 sweep -- has a full-rank 420 x 420 covariance.  This is the number a user sees.

``--impl reference`` times that CPU path alone, with all host threads, and
prints the same line with "impl": "reference".

``--config toy`` (BASELINE.json configs[1]): the three toy_dataset pairs (arrays of tests/golden/haspi_ref.npz) on one B200
against the CPU oracle, labels compared.
``--config ganround`` (BASELINE.json configs[3]): latency of one GAN sampling round of train_nele.py:36-37,
205-214, 320-335 -- 600 pairs labelled with norm=True (300 generator outputs + 300 pre-enhanced) and 480
validation pairs with norm=False, 2 - 3.5 s each -- from band gains on the device to labels on the host
(nele_resyn + nele_score_batch), pairs dealt over the ranks by length (shard.partition), one gather.
``--config sweep`` (configs[4]): 65 536 pairs of 16000 * U[3, 10] samples (default_rng(12345)), strong scaling over
the ranks through the same partition; the CPU arm scores a 256-pair subsample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FS = 16000
METRIC = "audio_seconds_scored_per_second"
UNIT = "audio-s/s"


# ----------------------------------------------------------------- helpers
def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 6:
                    self.rows.append(parts)
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        reasons = [n for k, n in enumerate(names) if any(r[2 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows)}


def make_workload(pairs, seconds, seed, unique, samples=None):
    from nele_gan_b200.synth import make_batch
    from nele_gan_b200.engine import pack
    L = int(round(seconds * FS)) if samples is None else int(samples)
    refs, degs = make_batch(pairs, L, seed=seed, unique=unique)
    fr, offs, lens = pack(refs)
    fd, _, _ = pack(degs)
    return refs, degs, fr, fd, offs, lens


# algorithmic (compulsory) HBM bytes per pair of each kernel: inputs read once + outputs
# written once with everything between them on chip ("Kernels and rooflines")
def kernel_bytes_per_pair(name, L16, Nf=1950, rank=420):
    n24 = -(-L16 * 3 // 2)
    nsub = -(-n24 // 9)
    n10 = -(-L16 * 5 // 8)
    nfr = max(n10 // 128 - 2, 1)
    F = Nf + 14
    table = {
        "haspi_prep": 2 * (4 * L16 + 4 * n24),              # waveform in, middle-ear output (f32) out
        "haspi_control": 2 * 4 * n24,
        "haspi_ear": 2 * (4 * n24 + 128 * nsub),
        "haspi_cep": 2 * 128 * nsub + 2 * 5 * 4 * nsub,
        "haspi_modcorr": 2 * 5 * 4 * nsub,
        "estoi_resample": 2 * (4 * L16 + 4 * n10),
        "estoi_vad": 4 * n10,
        "estoi_tob": 2 * 4 * n10 + 2 * 60 * nfr,
        "estoi_corr": 2 * 60 * nfr,
        "siib_wrapvad": 4 * L16,
        "siib_vad": 2 * 4 * L16,
        "siib_spec": 2 * 4 * L16 + 2 * 128 * F,
        "siib_mask": 2 * 2 * 128 * F,
        "siib_jacobi_cluster": 2 * rank * 448 * 4,
        "siib_cov": 128 * F + 15 * 8192,            # xx lag blocks, FP64
        "siib_cov32": 2 * 128 * F + 44 * 8192,      # yy, xy, yx lag blocks, FP32 sums
        "siib_expand": 59 * 8192 + 420 * 420 * (8 + 4 + 4),
        "siib_chol": 420 * 420 * 8 + rank * 448 * 4,
        "siib_jacobi": 2 * rank * 448 * 4,
        "siib_quad": rank * 448 * 4 + 2 * 420 * 420 * 4,
        "siib_tridiag": 420 * 420 * 8 + 420 * 448 * 4,   # Sxx in (FP64), reflectors out (FP32); the FP32 work matrix stays in L2
        "siib_trieig": 2 * 420 * 448 * 4,                # eigenvectors of T out, scratch
        "siib_backtf": 3 * 420 * 448 * 4,                # reflectors + eigenvectors of T in, G out
        "siib_projquad": rank * 448 * 4 + 2 * 128 * F,
    }
    return table.get(name)


def pipeline_bytes_per_pair(L16, Nf=1950):
    """125.3 N16 (HASPI) + 12 N16 (ESTOI) + 16 N16 + 6720 Nf (SIIB)."""
    return (125.3 + 12 + 16) * L16 + 6720 * Nf


# ------------------------------------------------------------ CPU reference
def _cpu_score_one(args):
    x, y = args
    from oracle import intel_np
    return intel_np.score_pair(x, y, FS, norm=True, noise="numpy")


def cpu_reference_run(refs, degs, cores, steps, warmup):
    """The oracle (port of the reference's three metrics + intel.py wrappers) under
    joblib.Parallel(n_jobs=cores), the fan-out of audio_util.py:146-202 without disk I/O."""
    from joblib import Parallel, delayed
    n = len(refs)
    sec = sum(len(r) for r in refs) / FS
    with Parallel(n_jobs=cores) as pool:
        for _ in range(max(warmup, 1)):   # per-worker imports + numba cache load
            pool(delayed(_cpu_score_one)((refs[i % n], degs[i % n])) for i in range(cores))
        t0 = time.perf_counter()
        for _ in range(steps):
            pool(delayed(_cpu_score_one)((refs[i], degs[i])) for i in range(n))
        dt = time.perf_counter() - t0
    return sec * steps / dt, dt / steps


# ------------------------------------------------- configs[1]: the toy corpus
def run_toy(a, rank, world, local_rank, cores):
    """BASELINE.json configs[1]: HASPI + SIIB + ESTOI labels of every toy_dataset utterance -- Train (Clean, MultiEnh +
    Noise), Train (Clean, Clean + Noise), Test (Clean, Clean + Noise); the arrays are those of tests/golden/haspi_ref.npz
    (the corpus itself is not on the GPU box) -- one B200 against the CPU oracle on the host, labels compared."""
    if rank != 0:
        return 0
    z = np.load(os.path.join(ROOT, "tests", "golden", "haspi_ref.npz"))
    names = ("toy_train_multienh", "toy_train_clean", "toy_test_clean")
    refs = [z[n + "/x"] for n in names]
    degs = [z[n + "/y"] for n in names]
    audio_s = sum(len(r) for r in refs) / FS
    from joblib import Parallel, delayed
    from oracle import intel_np
    t0 = time.perf_counter()
    want = Parallel(n_jobs=min(cores, 3))(delayed(intel_np.score_pair)(refs[i], degs[i], FS, True, None) for i in range(3))
    t_cpu_first = time.perf_counter() - t0
    t0 = time.perf_counter()
    want = np.array(Parallel(n_jobs=min(cores, 3))(delayed(intel_np.score_pair)(refs[i], degs[i], FS, True, None) for i in range(3)))
    t_cpu = time.perf_counter() - t0
    line = {"metric": "toy_corpus_labelling_latency", "unit": "ms", "higher_is_better": False, "n_gpus": 1, "steps": a.steps,
            "warmup": a.warmup, "config": {"workload": "toy_dataset: 3 (clean, degraded) pairs of 33 536 / 34 048 samples, mapped labels",
                                           "audio_seconds": audio_s},
            "cpu_baseline": {"value": t_cpu * 1e3, "unit": "ms", "cores": min(cores, 3), "kind": "port",
                             "sample": "the three pairs, one per worker process, second pass (first pass incl. imports: %.0f ms)" % (t_cpu_first * 1e3)}}
    if a.impl == "reference":
        line.update({"impl": "reference", "value": t_cpu * 1e3})
        print(json.dumps(line))
        return 0
    from nele_gan_b200.engine import Engine
    eng = Engine(local_rank)
    for _ in range(max(a.warmup, 1)):
        r = eng.score_batch(refs, degs, mapped=True, no_dither=True)
    lat = []
    for _ in range(max(a.steps, 1)):
        t0 = time.perf_counter()
        r = eng.score_batch(refs, degs, mapped=True, no_dither=True)
        lat.append((time.perf_counter() - t0) * 1e3)
    dev = np.abs(r.scores - want)
    line.update({"value": float(np.median(lat)), "ms_per_step": float(np.mean(lat)), "kernel_ms": eng.last_timing()[0],
                 "gpu_launches": int(eng.last_timing()[1]), "labels": {n: [float(v) for v in r.scores[i]] for i, n in enumerate(names)},
                 "max_abs_deviation_from_cpu_oracle": {"siib": float(dev[:, 0].max()), "haspi": float(dev[:, 1].max()),
                                                       "estoi": float(dev[:, 2].max())},
                 "speedup_vs_cpu": t_cpu * 1e3 / float(np.median(lat))})
    print(json.dumps(line))
    return 0


# ------------------------------------------------- configs[3]: GAN sampling round
def _dist_setup(world, local_rank):
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU path")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    return torch, dist, torch.device("cuda", local_rank)


def run_ganround(a, rank, world, local_rank, cores):
    """Latency of one sampling round (train_nele.py:36-37, 205-214, 320-335): 300 generator outputs + 300 pre-enhanced
    utterances labelled with norm=True and 480 validation utterances with norm=False, 2 - 3.5 s each, from the
    generator's band gains on the device to the labels on the host.  The generator itself is out of scope
    (model.py, checkpoint missing): its output is replaced by random band gains in [0.5, 2.5]."""
    from nele_gan_b200 import inloop, shard
    from nele_gan_b200.engine import default_engine
    from nele_gan_b200.synth import make_batch
    if a.impl == "reference":
        if rank != 0:
            return 0
        rng = np.random.default_rng(4242)
        lens = (FS * rng.uniform(2.0, 3.5, size=max(cores, 8))).astype(np.int64)
        refs, degs = make_batch(len(lens), lens, seed=555_000, unique=64)
        v, spstep = cpu_reference_run(refs, degs, cores, a.steps, a.warmup)
        total_s = 1080 * 2.75
        print(json.dumps({"impl": "reference", "metric": "gan_sampling_round_latency", "value": total_s / v * 1e3, "unit": "ms",
                          "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "higher_is_better": False,
                          "note": "extrapolated from %d pairs scored at %.1f audio-s/s on %d host threads (labels only: the "
                                  "reference also writes and re-reads 1080 WAV files per round)" % (len(lens), v, cores),
                          "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": "%d pairs" % len(lens)}}))
        return 0
    torch, dist, dev = _dist_setup(world, local_rank)
    rng = np.random.default_rng(4242)
    sets = []                                              # (pairs, norm)
    for n_pairs, norm, seed in ((600, True, 555_000), (480, False, 556_000)):
        lens = (FS * rng.uniform(2.0, 3.5, size=n_pairs)).astype(np.int64)
        mine = shard.partition(lens, world)[rank]
        refs, degs = make_batch(n_pairs, lens, seed=seed, unique=64)
        L = int(lens.max())
        k = len(mine)
        clean = np.zeros((k, L), np.float32)
        noise = np.zeros((k, L), np.float32)
        for r_, i in enumerate(mine):
            clean[r_, :lens[i]] = refs[i]
            noise[r_, :lens[i]] = degs[i] - refs[i]
        alpha2 = np.random.default_rng(seed + rank).uniform(0.5, 2.5, size=(k, 1 + L // 256, 64)).astype(np.float32)
        sets.append({"n": n_pairs, "norm": norm, "idx": mine, "lens": lens[mine].astype(np.int32),
                     "clean": torch.from_numpy(clean).to(dev), "noise": torch.from_numpy(noise).to(dev),
                     "alpha2": torch.from_numpy(alpha2).to(dev), "audio_s": float(lens.sum()) / FS})
    eng = default_engine(local_rank)

    def barrier():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def one_round():
        kms, nl, out = 0.0, 0, []
        for st in sets:
            if len(st["idx"]):
                sc = inloop.label_sampling_round(st["alpha2"], st["clean"], st["noise"], lengths=st["lens"], norm=st["norm"],
                                                 seed=7).numpy()
                t = eng.last_timing()
                kms, nl = kms + t[0], nl + t[1]
            else:
                sc = np.zeros((0, 3))
            if dist is not None:
                rec = np.zeros((len(st["idx"]), shard.RECORD))
                rec[:, :3] = sc
                sc = shard.gather_records(rec, st["idx"], st["n"], device=dev)[:, :3]
            out.append(sc)
        return out, kms, nl

    for _ in range(a.warmup):
        one_round()
    sampler = ClockSampler(local_rank)
    sampler.start()
    lat, kms_all, launches = [], [], 0
    for _ in range(a.steps):
        barrier()
        t0 = time.perf_counter()
        out, kms, nl = one_round()
        lat.append((time.perf_counter() - t0) * 1e3)
        kms_all.append(kms)
        launches += nl
    barrier()
    clocks = sampler.stop()
    t = torch.tensor([lat, kms_all], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    lat, kms_all = t[0].cpu().numpy(), t[1].cpu().numpy()
    if rank == 0:
        audio_s = sum(st["audio_s"] for st in sets)
        ok = all(np.isfinite(o).all() for o in out)
        print(json.dumps({
            "metric": "gan_sampling_round_latency", "value": float(np.median(lat)), "unit": "ms", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": float(np.mean(lat)), "higher_is_better": False,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32+f64",
            "data": "synthetic speech-shaped pairs; random band gains stand in for the generator's output",
            "config": {"workload": "GAN sampling round: 600 pairs norm=True + 480 pairs norm=False, 2-3.5 s, band gains -> "
                                   "nele_resyn (Resyn, PCM-16, + noise) -> HASPI v2 + SIIB^Gauss + ESTOI -> host",
                       "pairs": 1080, "audio_seconds": audio_s, "pairs_per_rank": [int(len(st["idx"])) for st in sets]},
            "latency_ms": {"min": float(lat.min()), "median": float(np.median(lat)), "max": float(lat.max())},
            "kernel_ms_per_round_max_rank": float(np.median(kms_all)),
            "throughput": {"value": audio_s / (float(np.median(lat)) * 1e-3), "unit": UNIT},
            "gpu_launches": int(launches), "clocks": clocks, "all_labels_finite": bool(ok)}))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


# ------------------------------------------------------ configs[4]: 3 - 10 s sweep
def run_sweep(a, rank, world, local_rank, cores):
    """65 536 pairs of 16000 * U[3, 10] samples (default_rng(12345)), strong scaling: the pairs are dealt over the ranks
    by length (shard.partition), every rank scores its shard, one gather returns the records in input order.  The
    waveforms are 64 distinct synthetic pairs of 10 s; pair i is pair i % 64 cut to its own length -- for the
    device-resident `value` the engine reads all of them from one 82 MB pool through offs / lens (inputs are read
    only), so 65 536 independent pairs cost no 54 GB of host memory; `e2e` uploads real, distinct host buffers for as
    many pairs of the shard as fit the stated budget."""
    from nele_gan_b200 import shard
    from nele_gan_b200.engine import Engine, pack
    from nele_gan_b200.synth import make_batch
    n = a.sweep_pairs
    lens = (FS * np.random.default_rng(12345).uniform(3.0, 10.0, size=n)).astype(np.int64)
    audio_total = float(lens.sum()) / FS
    cfg = {"workload": "sweep %d pairs x U[3,10] s (default_rng(12345)), HASPI v2 + SIIB^Gauss + ESTOI, mapped" % n,
           "pairs": n, "audio_seconds": audio_total, "fs": FS}
    if a.impl == "reference":
        if rank != 0:
            return 0
        sub = 256
        refs, degs = make_batch(sub, lens[:sub], seed=777_000, unique=64)
        v, spstep = cpu_reference_run(refs, degs, cores, max(1, min(a.steps, 1)), 1)
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": 1,
                          "warmup": 1, "ms_per_step": spstep * 1e3, "higher_is_better": True, "scaling": "strong",
                          "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
                          "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                           "sample": "first %d pairs of the sweep (%.0f audio-s), one pass, joblib n_jobs=%d; "
                                                     "the whole sweep extrapolates linearly to %.0f s" % (
                                                         sub, float(lens[:sub].sum()) / FS, cores, audio_total / v)},
                          "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0
    torch, dist, dev = _dist_setup(world, local_rank)
    pool_r, pool_d = make_batch(64, 160000, seed=777_000)
    fr, poffs, _ = pack(pool_r)
    fd, _, _ = pack(pool_d)
    d_ref, d_deg = torch.from_numpy(fr).to(dev), torch.from_numpy(fd).to(dev)
    mine = shard.partition(lens, world)[rank]
    offs = poffs[mine % 64].astype(np.int64)
    mylens = lens[mine].astype(np.int32)
    eng = Engine(local_rank)
    stream = torch.cuda.Stream(device=dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def device_pass():
        return eng.score_packed(d_ref.data_ptr(), d_deg.data_ptr(), offs, mylens, fs=FS, mapped=True, seed=1,
                                device_input=True, stream=stream.cuda_stream)

    # warm-up: one full pass (the grow-only workspace reaches its final size -- cudaMalloc / cudaFree inside a timed pass
    # cost seconds -- and the tables are uploaded), then the timed passes
    device_pass()
    eng.set_profiling(True)
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    steps = max(1, min(a.steps, 2))
    kms, launches, kt_sum = 0.0, 0, {}
    t0 = time.perf_counter()
    for _ in range(steps):
        r = device_pass()
        tm = eng.last_timing()
        kms += tm[0]
        launches += tm[1]
        for k, (ms, nl) in eng.kernel_times().items():
            kt_sum[k] = kt_sum.get(k, 0.0) + ms
    barrier()
    wall_dev = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop()
    eng.set_profiling(False)
    # e2e: real host buffers for (a prefix of) the shard, records gathered over the ranks
    budget_pairs = min(len(mine), 8192)
    sub = mine[:budget_pairs]
    h_refs = [pool_r[i % 64][:lens[i]] for i in sub]
    h_degs = [pool_d[i % 64][:lens[i]] for i in sub]
    hfr, hoffs, hlens = pack(h_refs)
    hfd, _, _ = pack(h_degs)
    del h_refs, h_degs
    h_ref, h_deg = torch.from_numpy(hfr).pin_memory(), torch.from_numpy(hfd).pin_memory()
    barrier()
    t0 = time.perf_counter()
    rh = eng.score_packed(h_ref.data_ptr(), h_deg.data_ptr(), hoffs, hlens, fs=FS, mapped=True, seed=1, stream=stream.cuda_stream)
    if dist is not None:
        cap = budget_pairs * world
        shard.gather_records(shard.pack_records(rh), np.arange(budget_pairs, dtype=np.int64) * world + rank, cap, device=dev)
    barrier()
    wall_e2e = (time.perf_counter() - t0) * 1e3
    e2e_audio = float(hlens.sum()) / FS
    t = torch.tensor([kms / steps, wall_dev / steps, wall_e2e, e2e_audio], dtype=torch.float64, device=dev)
    tmax = t.clone()
    tsum = t.clone()
    if dist is not None:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
    if rank == 0:
        ms_dev = float(tmax[0])
        print(json.dumps({
            "metric": METRIC, "value": audio_total / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": 1, "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic", "config": cfg,
            "wall_ms_per_pass_device_inputs": float(tmax[1]),
            "e2e": {"value": float(tsum[3]) / (float(tmax[2]) * 1e-3), "unit": UNIT,
                    "sample": "%d pairs per rank with distinct pinned host buffers (one blocking call + gather)" % budget_pairs,
                    "h2d_bytes_per_step": int(h_ref.numel() * 8), "d2h_bytes_per_step": int(budget_pairs * 116),
                    "ms": float(tmax[2])},
            "gpu_launches": int(launches), "clocks": clocks, "pairs_ok": int(np.sum(r.ok)), "pairs_this_rank": int(len(mine)),
            "kernels_ms_per_pass_rank0": {k: round(v / steps, 2) for k, v in sorted(kt_sum.items(), key=lambda kv: -kv[1])[:14]}}))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


# -------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=("ours", "reference"))
    ap.add_argument("--pairs", type=int, default=4096, help="pairs per GPU per step")
    ap.add_argument("--seconds", type=float, default=3.0)
    ap.add_argument("--unique", type=int, default=0, help="distinct synthetic pairs per GPU (0 = all)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-general-case", action="store_true", help="skip the 47 999-sample (full-rank SIIB) measurement")
    ap.add_argument("--config", default="bench", choices=("bench", "toy", "ganround", "sweep"),
                    help="bench = BASELINE.json configs[2] (the driver's line); toy = configs[1]; ganround = configs[3]; sweep = configs[4]")
    ap.add_argument("--sweep-pairs", type=int, default=65536)
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = os.cpu_count() or 1
    if a.config == "toy":
        return run_toy(a, rank, world, local_rank, cores)
    if a.config == "ganround":
        return run_ganround(a, rank, world, local_rank, cores)
    if a.config == "sweep":
        return run_sweep(a, rank, world, local_rank, cores)
    workload = "synthetic %d x %.1f s 16 kHz RMS=0.03 speech-shaped noisy pairs per GPU, HASPI v2 + SIIB^Gauss + ESTOI" % (
        a.pairs, a.seconds)
    config = {"workload": workload, "pairs_per_gpu": a.pairs, "seconds_per_pair": a.seconds, "fs": FS,
              "metrics": ["siib", "haspi", "estoi"], "mapped": True,
              "l2": "inputs are %.2f GB per step per GPU, larger than the 126 MB L2; no explicit flush" % (
                  a.pairs * a.seconds * FS * 8 / 1e9)}

    # ------------------------------------------------------ reference arm
    if a.impl == "reference":
        if rank != 0:
            return 0
        from nele_gan_b200.synth import make_batch
        sample = max(cores, 8)
        refs, degs = make_batch(sample, int(round(a.seconds * FS)), seed=666_000)
        v, spstep = cpu_reference_run(refs, degs, cores, a.steps, a.warmup)
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": spstep * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d pairs x %.1f s per step (one per host thread), joblib n_jobs=%d, "
                                       "oracle/ port of pyhaspi2.haspi_v2 + pysiib(gauss) + pystoi(extended)" % (
                                           sample, a.seconds, cores)},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0

    # ----------------------------------------------------------- our arm
    import torch
    from nele_gan_b200 import shard
    from nele_gan_b200.engine import Engine, BatchResult
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU path")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    refs, degs, fr, fd, offs, lens = make_workload(a.pairs, a.seconds, 666_000 + rank * a.pairs,
                                                   a.unique if a.unique > 0 else None)
    audio_s = float(lens.sum()) / FS
    h_ref = torch.from_numpy(fr).pin_memory()
    h_deg = torch.from_numpy(fd).pin_memory()
    d_ref, d_deg = h_ref.to(dev), h_deg.to(dev)
    n = len(lens)
    out = (np.empty((n, 3)), np.empty((n, 10)), np.empty(n, dtype=np.int32))
    eng = Engine(local_rank)
    stream = torch.cuda.Stream(device=dev)
    sptr = stream.cuda_stream

    def barrier():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step_device():
        return eng.score_packed(d_ref.data_ptr(), d_deg.data_ptr(), offs, lens, fs=FS, mapped=True, seed=1,
                                device_input=True, stream=sptr, out=out)

    # host buffers as the reference holds them: 16-bit PCM of the clean, the enhanced and the noise signal (the corpus
    # files and the generator outputs written by sf.write(..., 'PCM_16'), train_nele.py:313); here enhanced = clean
    q16 = lambda v: np.clip(np.round(v * 32768.0), -32768, 32767).astype(np.int16)
    h_c16 = torch.from_numpy(q16(fr)).pin_memory()
    h_n16 = torch.from_numpy(q16(fd - fr)).pin_memory()

    pending = []

    def gather(r):
        # the one collective of the path: the gather of the per-pair records.  It is started here and finished after the
        # next step has been queued (drain() after the last), so every step's records arrive inside the timed region but
        # the ranks do not wait for each other once per step
        if dist is not None:
            rec = shard.pack_records(r)
            pending.append(shard.gather_records_start(rec, np.arange(n, dtype=np.int64) + rank * n, n * world, device=dev))
            while len(pending) > 1:
                shard.gather_records_finish(pending.pop(0))
        return r

    def drain():
        while pending:
            shard.gather_records_finish(pending.pop(0))

    def step_host():           # float32 host buffers (8 bytes per sample pair cross PCIe)
        return gather(eng.score_packed(h_ref.data_ptr(), h_deg.data_ptr(), offs, lens, fs=FS, mapped=True, seed=1,
                                       stream=sptr, out=out))

    trace = os.environ.get("NELE_BENCH_TRACE") == "1"   # per-rank wall time of the pieces of an e2e step, on stderr
    tr_acc = {"score": 0.0, "gather": 0.0, "prefetch": 0.0, "n": 0}

    def step_host_pcm():       # int16 host buffers (6 bytes), enhanced + noise formed on the device
        t_a = time.perf_counter()
        r = eng.score_packed_pcm16(h_c16.data_ptr(), h_c16.data_ptr(), h_n16.data_ptr(), offs, lens, fs=FS,
                                   mapped=True, seed=1, stream=sptr, out=out)
        t_b = time.perf_counter()
        gather(r)
        if trace:
            tr_acc["score"] += t_b - t_a
            tr_acc["gather"] += time.perf_counter() - t_b
            tr_acc["n"] += 1
        return r

    # ---- value: device-resident inputs
    for _ in range(a.warmup):
        step_device()
    eng.set_profiling(True)
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches, kt_sum = 0, {}
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for _ in range(a.steps):
            r = step_device()
            launches += eng.last_timing()[1]
            for k, (ms, nl) in eng.kernel_times().items():
                t = kt_sum.setdefault(k, [0.0, 0])
                t[0] += ms
                t[1] += nl
        ev1.record(stream)
    barrier()
    ms_dev = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    eng.set_profiling(False)
    # ---- e2e: host buffers through the public API
    for _ in range(max(1, a.warmup // 2)):
        step_host()
    drain()
    barrier()
    # every step uploads its own 1.57 GB of waveforms; the upload of step k + 1 is started
    # (nele_prefetch, copy stream, second staging slot) before the blocking call of step k, so it
    # overlaps that step's kernels.  The first upload of the timed region is not hidden.
    t0 = time.perf_counter()
    eng.prefetch(h_ref.data_ptr(), h_deg.data_ptr(), offs, lens)
    for k in range(a.steps):
        if k + 1 < a.steps:
            eng.prefetch(h_ref.data_ptr(), h_deg.data_ptr(), offs, lens)
        r = step_host()
    drain()
    barrier()
    ms_e2e_f32 = (time.perf_counter() - t0) * 1e3
    # the same with the PCM-16 host buffers: this is what the drop-in read_batch_* path uploads (api._score_files)
    step_host_pcm()
    drain()
    barrier()
    t0 = time.perf_counter()
    eng.prefetch_pcm16(h_c16.data_ptr(), h_c16.data_ptr(), h_n16.data_ptr(), offs, lens)
    for k in range(a.steps):
        if k + 1 < a.steps:
            t_p = time.perf_counter()
            eng.prefetch_pcm16(h_c16.data_ptr(), h_c16.data_ptr(), h_n16.data_ptr(), offs, lens)
            tr_acc["prefetch"] += time.perf_counter() - t_p
        r = step_host_pcm()
    drain()
    barrier()
    ms_e2e = (time.perf_counter() - t0) * 1e3
    ok_pcm = int(np.sum(r.ok))
    if trace and tr_acc["n"]:
        sys.stderr.write("[bench trace] rank %d: e2e step %.2f ms = score call %.2f + gather %.2f + prefetch call %.2f (+ loop), kernels %.2f ms\n" % (
            rank, ms_e2e / a.steps, tr_acc["score"] / tr_acc["n"] * 1e3, tr_acc["gather"] / tr_acc["n"] * 1e3,
            tr_acc["prefetch"] / tr_acc["n"] * 1e3, eng.last_timing()[0]))
    # the same loop with plain blocking calls (no nele_prefetch): every upload is exposed
    t0 = time.perf_counter()
    for _ in range(a.steps):
        r = step_host()
    drain()
    barrier()
    ms_e2e_blocking = (time.perf_counter() - t0) * 1e3
    t = torch.tensor([ms_dev, ms_e2e, ms_e2e_blocking, ms_e2e_f32], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e, ms_e2e_blocking, ms_e2e_f32 = float(t[0]), float(t[1]), float(t[2]), float(t[3])
    ok = int(np.sum((r.status & 0xFFFFFF) == 0))
    nulldrop = int(np.sum((r.status & 0x01000000) != 0))   # (the general-case runs below reuse the output arrays)

    # ---- general case: the same step on pairs whose length is not a multiple of SIIB's 200-sample hop
    general = None
    if not a.no_general_case:
        g_samples = int(round(a.seconds * FS)) - 1
        g_steps, g_warm = max(2, min(a.steps, 5)), 2
        del d_ref, d_deg
        _, _, gfr, gfd, goffs, glens = make_workload(a.pairs, a.seconds, 666_000 + rank * a.pairs,
                                                     a.unique if a.unique > 0 else 64, samples=g_samples)
        gh_ref, gh_deg = torch.from_numpy(gfr).pin_memory(), torch.from_numpy(gfd).pin_memory()
        gd_ref, gd_deg = gh_ref.to(dev), gh_deg.to(dev)
        for _ in range(g_warm):
            eng.score_packed(gd_ref.data_ptr(), gd_deg.data_ptr(), goffs, glens, fs=FS, mapped=True, seed=1,
                             device_input=True, stream=sptr, out=out)
        eng.set_profiling(True)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        gkt = {}
        with torch.cuda.stream(stream):
            g0.record(stream)
            for _ in range(g_steps):
                rg = eng.score_packed(gd_ref.data_ptr(), gd_deg.data_ptr(), goffs, glens, fs=FS, mapped=True, seed=1,
                                      device_input=True, stream=sptr, out=out)
                for k, (ms, nl) in eng.kernel_times().items():
                    gkt[k] = gkt.get(k, 0.0) + ms
            g1.record(stream)
        barrier()
        g_ms_dev = g0.elapsed_time(g1)
        eng.set_profiling(False)
        g_ok = int(np.sum(rg.ok))
        g_null = int(np.sum(rg.siib_nullspace_dropped))
        eng.score_packed(gh_ref.data_ptr(), gh_deg.data_ptr(), goffs, glens, fs=FS, mapped=True, seed=1, stream=sptr, out=out)
        barrier()
        t0 = time.perf_counter()
        eng.prefetch(gh_ref.data_ptr(), gh_deg.data_ptr(), goffs, glens)
        for k in range(g_steps):
            if k + 1 < g_steps:
                eng.prefetch(gh_ref.data_ptr(), gh_deg.data_ptr(), goffs, glens)
            rg = eng.score_packed(gh_ref.data_ptr(), gh_deg.data_ptr(), goffs, glens, fs=FS, mapped=True, seed=1,
                                  stream=sptr, out=out)
            if dist is not None:
                shard.gather_records(shard.pack_records(rg), np.arange(n, dtype=np.int64) + rank * n, n * world, device=dev)
        barrier()
        g_ms_e2e = (time.perf_counter() - t0) * 1e3
        tg = torch.tensor([g_ms_dev, g_ms_e2e], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(tg, op=dist.ReduceOp.MAX)
        g_audio = float(glens.sum()) / FS
        general = {
            "workload": "synthetic %d x %d samples (%.5f s) per GPU: SIIB tiling does not repeat, full-rank KLT" % (
                a.pairs, g_samples, g_samples / FS),
            "value": g_audio * world / (float(tg[0]) / g_steps * 1e-3), "unit": UNIT, "ms_per_step": float(tg[0]) / g_steps,
            "e2e": {"value": g_audio * world / (float(tg[1]) / g_steps * 1e-3), "unit": UNIT,
                    "ms_per_step": float(tg[1]) / g_steps, "h2d_bytes_per_step": int(gh_ref.numel() * 4 * 2),
                    "d2h_bytes_per_step": int(n * (14 * 8 + 4))},
            "steps": g_steps, "warmup": g_warm, "pairs_ok": g_ok, "pairs_siib_nullspace_dropped": g_null,
            "kernels_ms_per_step": {k: round(v / g_steps, 3) for k, v in sorted(gkt.items(), key=lambda kv: -kv[1])},
        }

    if rank == 0:
        peak, peak_src = measured_peaks()
        L16 = int(round(a.seconds * FS))
        dom = max(kt_sum.items(), key=lambda kv: kv[1][0])
        name, (kms, knl) = dom[0], dom[1]
        # geometry of this workload for the byte model: KLT frames and rank of pair 0
        eng.score_batch(refs[:1], degs[:1], metrics=("siib",), keep_stages=True)
        tile, rk = eng.stage("siib.tile"), eng.stage("siib.rank")
        Nf, rank_x = max(int(tile[3]) - 14, 1), int(rk[0])
        bpp = kernel_bytes_per_pair(name, L16, Nf, rank_x)
        per_launch_pairs = a.pairs * a.steps / max(knl, 1)
        achieved = bpp * per_launch_pairs / (kms / knl * 1e-3) / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                tj = json.load(f)
            if tj.get("pairs") == a.pairs and abs(tj.get("seconds", 0) - a.seconds) < 1e-9:
                traffic = tj.get("kernels", {}).get(name)
        except Exception:
            pass
        step_s = ms_dev / a.steps * 1e-3
        line = {
            "metric": METRIC, "value": audio_s * world / step_s, "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms_dev / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic", "config": config,
            "e2e": {"value": audio_s * world / (ms_e2e / a.steps * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": int(h_c16.numel() * 2 * 2 + h_n16.numel() * 2), "d2h_bytes_per_step": int(n * (14 * 8 + 4)),
                    "ms_per_step": ms_e2e / a.steps,
                    "host_buffers": "int16 PCM of clean / enhanced / noise (nele_score_batch_pcm16: what the drop-in read_batch_* path "
                                    "uploads for the reference's 16-bit WAV files); enhanced + noise is formed on the device",
                    "pairs_ok": ok_pcm,
                    "float32_host": {"value": audio_s * world / (ms_e2e_f32 / a.steps * 1e-3), "ms_per_step": ms_e2e_f32 / a.steps,
                                     "h2d_bytes_per_step": int(h_ref.numel() * 4 * 2),
                                     "ms_per_step_blocking_calls": ms_e2e_blocking / a.steps},
                    "pipelining": "upload of step k+1 (nele_prefetch[_pcm16]) overlaps the kernels of step k; every step's H2D and D2H copies are inside the timed region"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "kernel_ms_per_launch": kms / knl, "kernel_share_of_step": kms / ms_dev,
                         "algorithmic_bytes_per_pair": bpp,
                         "pipeline_frac": pipeline_bytes_per_pair(L16, Nf) * a.pairs / step_s / 1e9 / peak,
                         "siib_klt_frames": Nf, "siib_rank": rank_x},
            "kernels_ms_per_step": {k: round(v[0] / a.steps, 3) for k, v in sorted(kt_sum.items(), key=lambda kv: -kv[1][0])},
            "pairs_ok": ok,
            "pairs_siib_nullspace_dropped": nulldrop,
        }
        if general is not None:
            line["general_case"] = general
        if world == 1 and not a.no_cpu_baseline:
            sample = 2 * max(cores, 8)      # two pairs per host thread and pass: halves the run-to-run spread of one-pair passes
            v, spstep = cpu_reference_run(refs[:sample], degs[:sample], cores, 2, 1)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "first %d pairs of the batch, 2 timed passes, joblib n_jobs=%d" % (sample, cores)}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
