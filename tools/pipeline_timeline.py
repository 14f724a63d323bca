"""Debug aid: per-step device timeline of aboria_b200.pipeline.HostPipeline
(H2D / build / product / D2H start and end, ms since the first submit)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import aboria_b200 as ab
from aboria_b200 import kernels as K, synth
from aboria_b200 import pipeline as P
import bench

n = int(os.environ.get("ABR_BENCH_N", 32_000_000))
side, size = bench.grid_side(n)
dev = torch.device("cuda:0")
pos_host = torch.empty((n, 3), dtype=torch.float64, pin_memory=True)
pos_host.copy_(synth.torch_uniform_positions(n, 3, 0.0, 1.0, synth.SEED, 0, dev))
b_host = torch.from_numpy(synth.vector(n)).pin_memory()
ys = [torch.empty(n, dtype=torch.float64, pin_memory=True) for _ in range(3)]

marks = []
ev = lambda s: (lambda e: (e.record(s), e)[1])(torch.cuda.Event(enable_timing=True))

class TimedPipeline(P.HostPipeline):
    def submit(self, pos_host, b_host, y_host):
        s = self.slots[self.k % len(self.slots)]
        k = self.k
        self.k += 1
        if s.pending is not None:
            self._finish(s)
        if s.done is not None:
            s.done.synchronize()
        th = time.perf_counter()
        with torch.cuda.stream(s.stream):
            e0 = ev(s.stream)
            s.p.resize_from_positions(pos_host)
            s.b.copy_(torch.as_tensor(b_host), non_blocking=True)
            e1 = ev(s.stream)
            s.p.init_neighbour_search(self.low, self.high, self.periodic, self.n_leaf, assume_all_alive=True)
            e2 = ev(s.stream)
        s.marks = [k, th, e0, e1, e2]
        s.pending = y_host
        prev, self._last = self._last, s
        if prev is not None and prev is not s and prev.pending is not None:
            self._finish(prev)
        return s

    def _finish(self, s):
        y_host, s.pending = s.pending, None
        with torch.cuda.stream(s.stream):
            s.p.check_async()
            th = time.perf_counter()
            e3 = ev(s.stream)
            s.op.matvec(s.b, out=s.y)
            e4 = ev(s.stream)
            torch.as_tensor(y_host).copy_(s.y, non_blocking=True)
            e5 = ev(s.stream)
            s.done = e5
        marks.append(s.marks + [th, e3, e4, e5])

pipe = TimedPipeline(3, n, 0.0, 1.0, True, side, K.inv_dist(0.1), 10.0)
for k in range(3):
    pipe.submit(pos_host, b_host, ys[k % 3])
pipe.wait()
marks.clear()
torch.cuda.synchronize()
t0 = time.perf_counter()
base = torch.cuda.Event(enable_timing=True); base.record()
for k in range(8):
    pipe.submit(pos_host, b_host, ys[k % 3])
pipe.wait()
torch.cuda.synchronize()
print("ms/step", (time.perf_counter() - t0) / 8 * 1e3)
for k, th, e0, e1, e2, th2, e3, e4, e5 in marks:
    r = lambda e: base.elapsed_time(e)
    print(f"step {k}: host_submit {1e3*(th-t0):7.2f} h2d {r(e0):7.2f}-{r(e1):7.2f} build-{r(e2):7.2f} | host_finish {1e3*(th2-t0):7.2f} mv {r(e3):7.2f}-{r(e4):7.2f} d2h-{r(e5):7.2f}")
