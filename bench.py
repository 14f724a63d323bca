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
# workloads
# ----------------------------------------------------------------------------
class Workload:
    """What one bench run multiplies.  uniform = BASELINE config 5 (3-D periodic unit cube, r = bucket
    side, 1/(|dx|+0.1)): weak (n_per_gpu particles per GPU) or strong (n_total fixed, default 256M);
    clustered = BASELINE config 4 (SPH density sum on 64 Gaussian blobs + 10 % background, N = 16M,
    periodic (1,1,0), r = 2h), strong scaling."""

    def __init__(self, args, world):
        from aboria_b200 import kernels as K

        self.cloud, self.world = args.cloud, world
        if args.cloud == "clustered":
            self.scaling = "strong"
            self.n_total = args.n_total or 16_000_000
            h = 1.5 * self.n_total ** (-1.0 / 3.0)
            self.radius = 2.0 * h
            self.kernel = K.sph_density(h, 1.0 / self.n_total, 21.0 / (256.0 * np.pi))
            self.kernel_name = "SphDensity<3>"
            self.periodic = [True, True, False]
            self.flops_per_pair = 8.0 + 9.0 + 2.0  # distance + sqrt, Wendland polynomial + multiply-add
            self.name = (f"c4-strong: SPH density sum (tests/sph.h W), 3-D clustered cloud (64 Gaussian blobs sigma 0.03 + 10 % background), "
                         f"N={self.n_total} total, periodic (1,1,0), n_leaf=10, r=2h, fp64")
        else:
            self.scaling = args.scaling
            self.n_total = (args.n_total or 256_000_000) if args.scaling == "strong" else args.n_per_gpu * world
            self.radius, _ = grid_side(self.n_total)
            self.kernel = K.inv_dist(EPS)
            self.kernel_name = "InvDistFast"
            self.periodic = True
            self.flops_per_pair = 13.0  # 3D-1 distance + sqrt,add,div + 2*BR*BC
            self.name = (f"c5-{self.scaling}: 3-D periodic unit cube, uniform random, n_leaf=10, r=bucket side, kernel 1/(|dx|+0.1), fp64"
                         + (f", N={self.n_total} total" if self.scaling == "strong" else ""))
        base, rem = divmod(self.n_total, world)
        self.shares = [base + (1 if g < rem else 0) for g in range(world)]

    def share(self, rank):
        return sum(self.shares[:rank]), self.shares[rank]

    def positions(self, rank, dev):
        """this rank's share of the GLOBAL cloud (a range of particle ids, anywhere in the domain)"""
        import torch

        from aboria_b200 import synth

        first, n = self.share(rank)
        if self.cloud == "clustered":
            return torch.from_numpy(synth.clustered_positions(n, first_id=first)).to(dev)
        return synth.torch_uniform_positions(n, 3, 0.0, 1.0, synth.SEED, first, dev)

    def expected_pairs(self):
        if self.cloud == "clustered":
            return None
        return self.n_total * (1.0 + 4.0 / 3.0 * np.pi * self.radius**3 * self.n_total)


def extra_configs(dev):
    """BASELINE.json configs c1-c4 (build + one product each, device resident, best of 3) so that the
    driver's record carries them; pair counts come from the stats kernel, whose pair sets the GPU parity
    tests check against the oracle at these sizes (tests/test_gpu_parity.py: c1, c2 in full; c3, c4 sampled)."""
    import torch

    import aboria_b200 as ab
    from aboria_b200 import kernels as K
    from aboria_b200 import synth

    def timed(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        return best

    out = []
    L3 = (4_000_000 / 0.8442) ** (1.0 / 3.0)
    h2 = 0.5 * np.sqrt(30.0 / (np.pi * 1_000_000))
    h4 = 1.5 * 16_000_000 ** (-1.0 / 3.0)
    cfgs = [
        ("c1: 3-D periodic unit cube N=1e5, 1/(r+0.1), r=0.05", 3, lambda: synth.torch_uniform_positions(100_000, 3, 0.0, 1.0, synth.SEED, 0, dev), 0.0, 1.0, True, 0.05,
         K.inv_dist(0.1)),
        ("c2: 2-D RBF Wendland C2 N=1e6, ~30 nbrs", 2, lambda: synth.torch_uniform_positions(1_000_000, 2, 0.0, 1.0, synth.SEED, 0, dev), 0.0, 1.0, False, 2 * h2,
         K.wendland_c2(h2)),
        ("c3: 3-D LJ force N=4M periodic, cutoff 2.5 sigma", 3, lambda: synth.torch_uniform_positions(4_000_000, 3, 0.0, L3, synth.SEED, 0, dev), 0.0, L3, True, 2.5,
         K.lj_force(3, 1.0, 1.0)),
        ("c4: SPH density N=16M clustered, periodic (1,1,0), r=2h", 3, lambda: torch.from_numpy(synth.clustered_positions(16_000_000)).to(dev), 0.0, 1.0,
         [True, True, False], 2 * h4, K.sph_density(h4, 1.0 / 16_000_000, 21.0 / (256.0 * np.pi))),
    ]
    for name, D, make, low, high, periodic, radius, kern in cfgs:
        pos = make()
        n = pos.shape[0]
        p = ab.Particles(D, 0)
        op = ab.create_sparse_operator(p, p, radius, kern)

        def build():
            p.resize_from_positions(pos)
            p.init_neighbour_search(low, high, periodic, N_LEAF)

        build()
        b = torch.ones(n * kern.block_cols, dtype=torch.float64, device=dev)
        cnt, _ = p.pair_stats(radius)
        pairs = int(cnt.long().sum().item())
        del cnt
        ms_build = timed(build)
        y = torch.zeros(n * kern.block_rows, dtype=torch.float64, device=dev)
        ms_mv = timed(lambda: op.matvec(b, out=y))
        out.append({"config": name, "n": n, "buckets": int(p.grid()[2]), "radius": float(radius), "pairs": pairs, "pairs_per_row": pairs / n,
                    "ms_build": ms_build, "ms_matvec": ms_mv, "pairs_per_s": pairs / (ms_mv * 1e-3), "build_mparticles_per_s": n / (ms_build * 1e-3) / 1e6,
                    "rows_recomputed_by_exact_walk": p.last_counters()["walk_rows"]})
        del p, op, pos, b, y
        torch.cuda.empty_cache()
    return out


# ----------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------
def run_ours(args):
    import torch

    import aboria_b200 as ab
    from aboria_b200 import synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    wl = Workload(args, world)
    if world > 1:
        # keep stdout to the single JSON line: NCCL prints its version banner there at VERSION/INFO level
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        import torch.distributed as dist

        from aboria_b200 import slab

        import datetime

        # a rank that fails must not leave the others waiting for the default 10 minutes
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=int(os.environ.get("ABR_NCCL_TIMEOUT_S", 180))))
        return slab.run_bench(args, wl, rank, world, dev, METRIC, UNIT, emit)

    n = wl.n_total
    radius = wl.radius
    hbm_peak, peak_src = measured_peaks()

    # synthetic input, resident in HBM before the timed region
    pos_unsorted = wl.positions(0, dev)
    b = torch.from_numpy(synth.vector(n)).to(dev)
    p = ab.Particles(3, 0)
    op = ab.create_sparse_operator(p, p, radius, wl.kernel)
    fp64_peak = p.probe_fp64_peak()

    y_buf = torch.zeros(n, dtype=torch.float64, device=dev)

    def step():
        p.resize_from_positions(pos_unsorted)  # fresh unsorted set (device copy into the container's buffer)
        # asynchronous update: the cloud lies inside the domain, nothing dies;
        # verified by p.check_async() after the timed region
        p.init_neighbour_search(0.0, 1.0, wl.periodic, N_LEAF, assume_all_alive=True)
        return op.matvec(b, out=y_buf)

    ev = lambda: torch.cuda.Event(enable_timing=True)
    # pair count (exact, from the stats kernel) outside the timed region
    step()
    p.check_async()  # also tells the library what the build saw (heavy buckets of a clustered cloud are split in the product)
    cnt, _ = p.pair_stats(radius)
    pairs = int(cnt.long().sum().item())
    del cnt
    expect = wl.expected_pairs()
    if expect is not None:
        assert abs(pairs - expect) / expect < 2e-3, f"pair count {pairs} differs from the expectation {expect:.6g} of a uniform cloud"
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
        p.init_neighbour_search(0.0, 1.0, wl.periodic, N_LEAF, assume_all_alive=True)
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
    ncells = int(p.grid()[2])

    # ---- end-to-end through the public API with HOST (pinned) buffers ----
    pos_host = torch.empty((n, 3), dtype=torch.float64, pin_memory=True)
    pos_host.copy_(pos_unsorted)
    b_host = torch.empty(n, dtype=torch.float64, pin_memory=True)
    b_host.copy_(b)
    y_host = torch.empty(n, dtype=torch.float64, pin_memory=True)
    e2e_steps = max(1, min(args.steps, 3))

    def e2e_step():
        p.resize_from_positions(pos_host)            # H2D positions
        p.init_neighbour_search(0.0, 1.0, wl.periodic, N_LEAF)
        op.matvec_host(b_host, y_host)               # H2D b, D2H y

    e2e_step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    serial_sec = (time.perf_counter() - t0) / e2e_steps
    y_serial = y_host.clone()

    if n <= 64_000_000:
        # the same steps through the streaming driver (aboria_b200/pipeline.py): three containers on
        # three streams, so the PCIe copies of step k+1 overlap the build + product of step k.  Every
        # step uploads its positions and b, builds, multiplies and downloads y.
        from aboria_b200.pipeline import HostPipeline

        pipe = HostPipeline(3, n, 0.0, 1.0, wl.periodic, radius, wl.kernel, N_LEAF)
        y_hosts = [torch.empty(n, dtype=torch.float64, pin_memory=True) for _ in range(3)]
        for k in range(3):
            pipe.submit(pos_host, b_host, y_hosts[k % 3])
        pipe.wait()
        pipe_steps = max(6, min(args.steps, 50))  # the same K steps as the device-timed region (pipeline fill and drain included in the wall clock)
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
        how = ("HostPipeline: pinned host buffers, 3 containers on 3 streams; the PCIe copies of step k+1 overlap the build + product of step k; "
               "wall clock over all steps")
    else:
        e2e_sec, pipe_steps = serial_sec, e2e_steps
        how = "synchronous drop-in calls (init_neighbour_search + K*b with pinned host buffers), no overlap: three pipelined containers of this size are not kept"
    e2e = {"value": pairs / e2e_sec, "unit": UNIT, "h2d_bytes_per_step": int(n * 24 + n * 8), "d2h_bytes_per_step": int(n * 8),
           "ms_per_step": e2e_sec * 1e3, "steps": pipe_steps, "how": how,
           "ms_per_step_unpipelined": serial_sec * 1e3,
           "unpipelined": "the reference's synchronous call shape: init_neighbour_search + K*b with host buffers, every copy exposed"}

    # ---- rooflines ----
    mv_bytes = n * (8 * 3 + 8) + n * (8 * 3 + 8) + 8 * ncells          # SURVEY §8d B_mv
    build_bytes = n * (8 * 3 * 2) + 4 * n + 8 * ncells + 2 * n * (8 + 1)  # B_build + id/alive columns
    mv_gbs = mv_bytes / (ms_mv * 1e-3) / 1e9
    build_gbs = build_bytes / (ms_build * 1e-3) / 1e9
    mv_tflops = wl.flops_per_pair * pairs / (ms_mv * 1e-3) / 1e12
    share = ms_mv / ms_per_step
    roofline = {"bound": "hbm", "kernel": f"abr::tiled_kernel<3, {wl.kernel_name}> (sparse matvec, dominant kernel of the step: {100 * share:.0f} % of it)",
                "achieved": mv_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": mv_gbs / hbm_peak,
                # dram__bytes_read.sum + dram__bytes_write.sum of one launch at N=32M (profiles/r2z_ncu_tiled_kernel_summary.txt)
                "traffic": 2.345e9 if (n == 32_000_000 and wl.cloud == "uniform") else None, "algorithmic_bytes": mv_bytes, "peak_source": peak_src,
                "note": "the product is instruction-issue / LSU bound, not HBM bound (arithmetic intensity >> 6 flop/B, SURVEY §8d; DESIGN.md §4.2): "
                        "see roofline_fp64 and profiles/r2z_ncu_tiled_kernel_summary.txt (11.5 G warp instructions per launch, issue slots 73 % busy, "
                        "LSU data pipe 74 %). algorithmic bytes = N(8D+8BR)+N(8D+8BC)+8C; achieved uses the product time incl. the 0.3 ms record-packing pass"}
    roofline_fp64 = {"bound": "fp64", "achieved": mv_tflops, "peak": fp64_peak, "unit": "TFLOP/s", "frac": mv_tflops / fp64_peak,
                     "peak_source": "measured here (abr_probe_fp64_peak, DFMA loop)", "flops_per_pair": wl.flops_per_pair,
                     "pairs_per_s_matvec_only": pairs / (ms_mv * 1e-3)}
    roofline_build = {"bound": "hbm", "kernel": "cell-list build (k_enforce_key + two-level radix sort + bounds + gather of position,id,alive)",
                      "achieved": build_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": build_gbs / hbm_peak,
                      "mparticles_per_s": n / (ms_build * 1e-3) / 1e6, "ms": ms_build, "algorithmic_bytes": build_bytes,
                      "floor": "profiles/r2_build_floor.txt: the same byte movement as plain copies (tools/build_floor.cu)"}

    # ---- CPU baseline beside it (bounded sample, rank 0, N=1 only) ----
    cpu = None
    if not args.no_cpu_baseline and wl.cloud == "uniform":
        from oracle import oracle as orc

        cores = orc.host_cores()
        cp, tb, tm = cpu_reference_step(args.cpu_sample, cores, reps=5, warm=1)
        cpu = {"value": cp / (tb + tm), "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"same workload at N={args.cpu_sample} (3-D periodic, r=side), best of 5 after 1 warm-up: std::sort build {tb:.2f}s + OpenMP matvec {tm:.2f}s",
               "build_mparticles_per_s": args.cpu_sample / tb / 1e6}

    # ---- the other BASELINE configs on the same record ----
    extra = None
    if args.extra and wl.cloud == "uniform" and wl.scaling == "weak":
        del p, op, y_buf, pos_unsorted, b, pos_host, b_host, y_host
        torch.cuda.empty_cache()
        extra = {"configs_c1_c4": extra_configs(dev),
                 "note": "BASELINE.json configs 1-4 on this GPU, device resident, best of 3 (the line's value/e2e are config 5)"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": wl.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl.name,
                   "n_particles_per_gpu": n, "n_particles": n, "buckets": ncells, "radius": radius, "pairs_per_matvec": pairs,
                   "l2": f"inputs ({n * 24 / 1e9:.2f} GB positions) exceed the 126 MB L2; no flush needed", "rows_recomputed_by_exact_walk": walk_rows},
        "ms_build": ms_build, "ms_matvec": ms_mv, "build_mparticles_per_s": n / (ms_build * 1e-3) / 1e6,
        "ms_build_min_max": [float(np.min(t_build)), float(np.max(t_build))], "host_enqueue_ms_per_step": float(np.median(host_t)),
        "roofline": roofline, "roofline_fp64": roofline_fp64, "roofline_build": roofline_build,
        "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "extra": extra,
    }
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-per-gpu", type=int, default=int(os.environ.get("ABR_BENCH_N", 32_000_000)))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="strong: --n-total particles (default 256M) split over the GPUs")
    ap.add_argument("--cloud", default="uniform", choices=["uniform", "clustered"], help="clustered: BASELINE config 4 (N=16M SPH density), strong scaling")
    ap.add_argument("--n-total", type=int, default=0)
    ap.add_argument("--cpu-sample", type=int, default=int(os.environ.get("ABR_BENCH_CPU_N", 2_000_000)))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", dest="extra", action="store_false", help="skip the c1-c4 timings appended to the N=1 line")
    args = ap.parse_args()
    guard_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
