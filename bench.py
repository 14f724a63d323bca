#!/usr/bin/env python
"""bench.py — one JSON line per run (driver contract).

Step  = one pass of the hot path over one batch of synthetic input:
        ordered cell-list build from UNSORTED positions (wrap, key, radix sort,
        bucket bounds, reorder of position/id/alive) followed by one sparse
        kernel product y = K b.
Metric = accepted pair interactions per second (BASELINE.json), whole job.

Workload (config.workload "c5-weak"): 3-D periodic unit cube, uniform random,
32M particles per GPU (256M at 8 GPUs), n_particles_in_leaf = 10,
r = bucket side, kernel 1/(|dx|+0.1), fp64.  The particle set (768 MB of
positions per GPU) is far larger than the 126 MB L2, so no L2 flush is needed
between steps.

  python bench.py [--gpus N] [--steps K] [--warmup W]          (our arm)
  python bench.py --impl reference ...                          (CPU reference arm)
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

METRIC = "pair interactions/s (fp64 sparse-kernel matvec, cell-list build included in the step)"
UNIT = "pairs/s"
EPS = 0.1
N_LEAF = 10.0


# The driver reads ONE JSON line from stdout.  Libraries (NCCL's version banner, for one)
# write to file descriptor 1 behind Python's back, so while the bench runs fd 1 points at
# stderr and the result line goes to the saved original.
_REAL_STDOUT = None


def guard_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        os.write(1, data)
    else:
        os.write(_REAL_STDOUT, data)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons during the timed region"""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s, p in zip(sm, power) if p >= 0.5 * max(power)] or sm
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": float(max(power))}


def grid_side(n_total):
    """bucket side of the reference grid for n_total particles in the unit cube
    (src/CellListOrdered.h:140-157)"""
    box_side = (N_LEAF / float(n_total) * 1.0) ** (1.0 / 3.0)
    size = int(np.floor(1.0 / box_side))
    return 1.0 / size, size


# ----------------------------------------------------------------------------
# CPU reference arm: the oracle restatement of the reference's own CPU path
# (the reference itself cannot be compiled here: Boost + Eigen absent)
# ----------------------------------------------------------------------------
def cpu_reference_step(n_sample, nthreads, reps=1, warm=0, inputs=None):
    """build (std::sort mode, position+id+alive reorder) + matvec on the host.
    Returns (pairs, seconds_build, seconds_matvec) of the best of `reps` repetitions
    after `warm` untimed ones."""
    from aboria_b200 import synth
    from oracle import oracle as orc

    orc.set_num_threads(nthreads)  # not OMP_NUM_THREADS: torchrun exports OMP_NUM_THREADS=1
    if inputs is None:
        inputs = (synth.uniform_positions(n_sample, 3), np.arange(n_sample, dtype=np.int64), synth.vector(n_sample))
    pos0, ids, b = inputs
    side, _ = grid_side(n_sample)
    best = None
    for rep in range(warm + reps):
        o = orc.Oracle(3)
        t0 = time.perf_counter()
        o.set_domain(0.0, 1.0, True, N_LEAF)
        pos = pos0.copy()
        out = o.update_positions(pos, None, orc.SORT_STD)
        ps = o.gather(out["order"], pos)
        o.gather(out["order"], ids)
        o.gather(out["order"], out["alive"])
        o.update_iterators(ps)
        t1 = time.perf_counter()
        y, pairs = o.sparse_matvec(ps, orc.K_INV_DIST, [EPS], side, b, nthreads=nthreads)
        t2 = time.perf_counter()
        if rep >= warm and (best is None or (t2 - t0) < best[1] + best[2]):
            best = (pairs, t1 - t0, t2 - t1)
    return best


def run_reference(args):
    """CPU arm: the reference's own algorithm (oracle restatement; the reference itself cannot be
    compiled here) on ALL host cores this process may use, on the same workload as our arm:
    the per-GPU share of c5 (n_per_gpu particles, r = bucket side).  Under torchrun only rank 0
    works; the thread count comes from the affinity mask, not from OMP_NUM_THREADS (torchrun
    sets that to 1)."""
    from aboria_b200 import synth
    from oracle import oracle as orc

    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = orc.host_cores()
    n = args.n_per_gpu
    inputs = (synth.uniform_positions(n, 3), np.arange(n, dtype=np.int64), synth.vector(n))
    times = []
    pairs = 0
    budget_s = float(os.environ.get("ABR_REF_BUDGET_S", 270.0))  # "the whole run ends within a few minutes"
    t_start = time.perf_counter()
    warm_done = 0
    for it in range(args.warmup + args.steps):
        elapsed = time.perf_counter() - t_start
        if it < args.warmup:
            # a CPU step has no launch / allocator warm-up to amortise beyond the first touch of the buffers
            if warm_done >= 1 and elapsed > 0.15 * budget_s:
                continue
            warm_done += 1
        elif times and elapsed + 1.2 * times[-1] > budget_s:
            break
        pairs, tb, tm = cpu_reference_step(n, cores, inputs=inputs)
        if it >= args.warmup:
            times.append(tb + tm)
    sec = float(np.mean(times))
    value = pairs / sec
    sample = (f"c5 workload at N={n} particles (3-D periodic unit cube, r=side, 1/(|dx|+0.1)) = our arm's per-GPU configuration; "
              f"std::sort build + OpenMP matvec over rows on {cores} threads; {len(times)} timed steps of {args.steps} requested "
              f"(wall budget {budget_s:.0f} s), {warm_done} warm-up")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times), "steps_requested": args.steps,
        "warmup": warm_done, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "c5-weak: 3-D periodic unit cube, uniform random, n_leaf=10, r=bucket side, kernel 1/(|dx|+0.1), fp64",
                   "n_particles_per_gpu": n, "n_particles": n, "pairs_per_matvec": int(pairs), "omp_threads": cores,
                   "note": "CPU arm runs ONE per-GPU share of the workload (a rate metric); at N GPUs our arm runs N such shares"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ----------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------
def run_ours(args):
    import torch

    import aboria_b200 as ab
    from aboria_b200 import kernels as K
    from aboria_b200 import synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    if world > 1:
        # keep stdout to the single JSON line: NCCL prints its version banner there at VERSION/INFO level
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        import torch.distributed as dist

        from aboria_b200 import slab

        dist.init_process_group("nccl", device_id=dev)
        return slab.run_bench(args, rank, world, dev, METRIC, UNIT, emit)

    n = args.n_per_gpu
    side, size = grid_side(n)
    radius = side
    hbm_peak, peak_src = measured_peaks()

    # synthetic input, resident in HBM before the timed region
    pos_unsorted = synth.torch_uniform_positions(n, 3, 0.0, 1.0, synth.SEED, 0, dev)
    b = torch.from_numpy(synth.vector(n)).to(dev)
    p = ab.Particles(3, 0)
    op = ab.create_sparse_operator(p, p, radius, K.inv_dist(EPS))
    fp64_peak = p.probe_fp64_peak()

    y_buf = torch.zeros(n, dtype=torch.float64, device=dev)

    def step():
        p.resize_from_positions(pos_unsorted)  # fresh unsorted set (device copy into the container's buffer)
        # asynchronous update: the uniform cloud lies inside the periodic box, nothing dies;
        # verified by p.check_async() after the timed region
        p.init_neighbour_search(0.0, 1.0, True, N_LEAF, assume_all_alive=True)
        return op.matvec(b, out=y_buf)

    ev = lambda: torch.cuda.Event(enable_timing=True)
    # pair count (exact, from the stats kernel) outside the timed region
    step()
    cnt, _ = p.pair_stats(radius)
    pairs = int(cnt.long().sum().item())
    del cnt
    for _ in range(max(0, args.warmup - 1)):
        step()
    torch.cuda.synchronize()

    launches0 = p.last_counters()["total_launches"]
    sampler = ClockSampler(local_rank)
    sampler.start()
    t_build, t_mv = [], []
    e0, e1 = ev(), ev()
    torch.cuda.synchronize()
    e0.record()
    evs = []
    host_t = []
    for _ in range(args.steps):
        th0 = time.perf_counter()
        a, bb_, c = ev(), ev(), ev()
        p.resize_from_positions(pos_unsorted)
        a.record()
        p.init_neighbour_search(0.0, 1.0, True, N_LEAF, assume_all_alive=True)
        bb_.record()
        y = op.matvec(b, out=y_buf)
        c.record()
        evs.append((a, bb_, c))
        host_t.append((time.perf_counter() - th0) * 1e3)
    e1.record()
    torch.cuda.synchronize()
    p.check_async()  # no particle died in the asynchronous updates
    clocks = sampler.stop()
    launches = p.last_counters()["total_launches"] - launches0
    total_ms = e0.elapsed_time(e1)
    for a, bb_, c in evs:
        t_build.append(a.elapsed_time(bb_))
        t_mv.append(bb_.elapsed_time(c))
    ms_per_step = total_ms / args.steps
    value = pairs / (ms_per_step * 1e-3)
    ms_build, ms_mv = float(np.mean(t_build)), float(np.mean(t_mv))
    walk_rows = p.last_counters()["walk_rows"]

    # ---- end-to-end through the public API with HOST (pinned) buffers ----
    pos_host = torch.empty((n, 3), dtype=torch.float64, pin_memory=True)
    pos_host.copy_(pos_unsorted)
    b_host = torch.empty(n, dtype=torch.float64, pin_memory=True)
    b_host.copy_(b)
    y_host = torch.empty(n, dtype=torch.float64, pin_memory=True)
    e2e_steps = max(1, min(args.steps, 3))

    def e2e_step():
        p.resize_from_positions(pos_host)            # H2D positions
        p.init_neighbour_search(0.0, 1.0, True, N_LEAF)
        op.matvec_host(b_host, y_host)               # H2D b, D2H y

    e2e_step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    serial_sec = (time.perf_counter() - t0) / e2e_steps
    y_serial = y_host.clone()

    # the same steps through the streaming driver (aboria_b200/pipeline.py): three containers on
    # three streams, so the PCIe copies of step k+1 overlap the build + product of step k.  Every
    # step uploads its positions and b, builds, multiplies and downloads y.
    from aboria_b200.pipeline import HostPipeline

    pipe = HostPipeline(3, n, 0.0, 1.0, True, radius, K.inv_dist(EPS), N_LEAF)
    y_hosts = [torch.empty(n, dtype=torch.float64, pin_memory=True) for _ in range(3)]
    for k in range(3):
        pipe.submit(pos_host, b_host, y_hosts[k % 3])
    pipe.wait()
    pipe_steps = max(6, min(args.steps, 12))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(pipe_steps):
        pipe.submit(pos_host, b_host, y_hosts[k % 3])
    pipe.wait()
    torch.cuda.synchronize()
    e2e_sec = (time.perf_counter() - t0) / pipe_steps
    for yh in y_hosts:
        rel = float(torch.linalg.norm(yh - y_serial) / torch.linalg.norm(y_serial))
        assert rel <= 1e-12, f"pipelined e2e result differs from the serial one (rel L2 {rel:.3e})"
    del pipe
    e2e = {"value": pairs / e2e_sec, "unit": UNIT, "h2d_bytes_per_step": int(n * 24 + n * 8), "d2h_bytes_per_step": int(n * 8),
           "ms_per_step": e2e_sec * 1e3, "steps": pipe_steps,
           "how": "HostPipeline: pinned host buffers, 3 containers on 3 streams; the PCIe copies of step k+1 overlap the build + product of step k; wall clock over all steps",
           "ms_per_step_unpipelined": serial_sec * 1e3}

    # ---- rooflines ----
    ncells = size ** 3
    mv_bytes = n * (8 * 3 + 8) + n * (8 * 3 + 8) + 8 * ncells          # SURVEY §8d B_mv
    build_bytes = n * (8 * 3 * 2) + 4 * n + 8 * ncells + 2 * n * (8 + 1)  # B_build + id/alive columns
    mv_gbs = mv_bytes / (ms_mv * 1e-3) / 1e9
    build_gbs = build_bytes / (ms_build * 1e-3) / 1e9
    flops_per_pair = 13.0  # 3D-1 distance + sqrt,add,div + 2*BR*BC
    mv_tflops = flops_per_pair * pairs / (ms_mv * 1e-3) / 1e12
    roofline = {"bound": "hbm", "kernel": "abr::tiled_kernel<3, InvDistFast> (sparse matvec, dominant kernel of the step: 80 % of it)",
                "achieved": mv_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": mv_gbs / hbm_peak,
                # dram__bytes_read.sum + dram__bytes_write.sum of one launch at N=32M (profiles/r1y_ncu_tiled_kernel_v9_summary.txt)
                "traffic": 2.345e9 if n == 32_000_000 else None, "algorithmic_bytes": mv_bytes, "peak_source": peak_src,
                "note": "the product is instruction-issue bound, not HBM bound (arithmetic intensity >> 6 flop/B, SURVEY §8d; DESIGN.md §4.2): "
                        "see roofline_fp64 and profiles/r1y_ncu_tiled_kernel_v9_summary.txt. algorithmic bytes = N(8D+8BR)+N(8D+8BC)+8C; "
                        "achieved uses the product time incl. the 0.3 ms record-packing pass"}
    roofline_fp64 = {"bound": "fp64", "achieved": mv_tflops, "peak": fp64_peak, "unit": "TFLOP/s", "frac": mv_tflops / fp64_peak,
                     "peak_source": "measured here (abr_probe_fp64_peak, DFMA loop)", "flops_per_pair": flops_per_pair,
                     "pairs_per_s_matvec_only": pairs / (ms_mv * 1e-3)}
    # what actually binds the product: warp-instruction issue.  Instruction count of one launch
    # from ncu (smsp__inst_executed.sum, profiles/r1y_ncu_tiled_kernel_v9_summary.txt; a property of
    # kernel + workload), issue peak = SMs x 4 schedulers x SM clock
    roofline_issue = None
    if n == 32_000_000 and clocks.get("sm_mhz"):
        inst = 11.514e9
        peak_issue = 148 * 4 * clocks["sm_mhz"] * 1e6
        ms_kernel = ms_mv - 0.40  # product time minus record packing, y zeroing and the exact-walk launch (profiles/r1y_launch_summary.txt)
        roofline_issue = {"bound": "issue", "achieved": inst / (ms_kernel * 1e-3) / 1e12, "peak": peak_issue / 1e12, "unit": "T warp-inst/s",
                          "frac": inst / (ms_kernel * 1e-3) / peak_issue, "warp_inst_per_pair": inst / pairs,
                          "source": "ncu smsp__inst_executed.sum of one launch (static for this workload) / live kernel time"}
    roofline_build = {"bound": "hbm", "kernel": "cell-list build (k_enforce_key + radix sort + bounds + gather of position,id,alive)",
                      "achieved": build_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": build_gbs / hbm_peak,
                      "mparticles_per_s": n / (ms_build * 1e-3) / 1e6, "ms": ms_build}

    # ---- CPU baseline beside it (bounded sample, rank 0, N=1 only) ----
    cpu = None
    if not args.no_cpu_baseline:
        from oracle import oracle as orc

        cores = orc.host_cores()
        cp, tb, tm = cpu_reference_step(args.cpu_sample, cores, reps=5, warm=1)
        cpu = {"value": cp / (tb + tm), "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"same workload at N={args.cpu_sample} (3-D periodic, r=side), best of 5 after 1 warm-up: std::sort build {tb:.2f}s + OpenMP matvec {tm:.2f}s",
               "build_mparticles_per_s": args.cpu_sample / tb / 1e6}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "c5-weak: 3-D periodic unit cube, uniform random, n_leaf=10, r=bucket side, kernel 1/(|dx|+0.1), fp64",
                   "n_particles_per_gpu": n, "n_particles": n, "buckets": ncells, "radius": radius, "pairs_per_matvec": pairs,
                   "l2": "inputs (0.77 GB positions) exceed the 126 MB L2; no flush needed", "rows_recomputed_by_exact_walk": walk_rows},
        "ms_build": ms_build, "ms_matvec": ms_mv, "build_mparticles_per_s": n / (ms_build * 1e-3) / 1e6,
        "ms_build_min_max": [float(np.min(t_build)), float(np.max(t_build))], "host_enqueue_ms_per_step": float(np.median(host_t)),
        "roofline": roofline, "roofline_fp64": roofline_fp64, "roofline_issue": roofline_issue, "roofline_build": roofline_build,
        "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
    }
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-per-gpu", type=int, default=int(os.environ.get("ABR_BENCH_N", 32_000_000)))
    ap.add_argument("--cpu-sample", type=int, default=int(os.environ.get("ABR_BENCH_CPU_N", 2_000_000)))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    guard_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
