#!/usr/bin/env python
"""Benchmark of the NELE-GAN intelligibility-labelling hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

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

``--impl reference`` times that CPU path alone, with all host threads, and
prints the same line with "impl": "reference".
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


def make_workload(pairs, seconds, seed, unique):
    from nele_gan_b200.synth import make_batch
    from nele_gan_b200.engine import pack
    L = int(round(seconds * FS))
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
        "haspi_prep": 2 * (4 * L16 + 4 * n24 + 8 * n24),
        "haspi_control": 2 * 8 * n24,
        "haspi_ear": 2 * (8 * n24 + 128 * nsub),
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
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cores = os.cpu_count() or 1
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

    def step_host():
        r = eng.score_packed(h_ref.data_ptr(), h_deg.data_ptr(), offs, lens, fs=FS, mapped=True, seed=1,
                             stream=sptr, out=out)
        if dist is not None:   # the one collective of the path: gather of the per-pair records
            rec = shard.pack_records(r)
            shard.gather_records(rec, np.arange(n, dtype=np.int64) + rank * n, n * world, device=dev)
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
    barrier()
    ms_e2e = (time.perf_counter() - t0) * 1e3
    # the same loop with plain blocking calls (no nele_prefetch): every upload is exposed
    t0 = time.perf_counter()
    for _ in range(a.steps):
        r = step_host()
    barrier()
    ms_e2e_blocking = (time.perf_counter() - t0) * 1e3
    t = torch.tensor([ms_dev, ms_e2e, ms_e2e_blocking], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e, ms_e2e_blocking = float(t[0]), float(t[1]), float(t[2])
    ok = int(np.sum((r.status & 0xFFFFFF) == 0))

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
                    "h2d_bytes_per_step": int(h_ref.numel() * 4 * 2), "d2h_bytes_per_step": int(n * (14 * 8 + 4)),
                    "ms_per_step": ms_e2e / a.steps,
                    "ms_per_step_blocking_calls": ms_e2e_blocking / a.steps,
                    "pipelining": "upload of step k+1 (nele_prefetch) overlaps the kernels of step k; every step's H2D and D2H copies are inside the timed region"},
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
        }
        if world == 1 and not a.no_cpu_baseline:
            sample = max(cores, 8)
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
