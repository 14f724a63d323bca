"""Host-buffer streaming driver for the hot path: build + product per step with
the particle positions, b and y living in (pinned) HOST memory.

The reference runs this path on host vectors (std::vector / Eigen, `K * b`,
src/Kernels.h:720-751); a caller that keeps its data on the host pays one
host->device copy of positions and b and one device->host copy of y per step.
Those copies (PCIe) cost as much as the device work itself, so this driver
keeps a few particle containers (default 3), each on its own CUDA stream, and
rotates through them: the copies of step k+1 overlap the build/product of step k.  Every step still
does everything — H2D of its inputs, wrap/kill + ordered cell-list build +
reorder, the product, D2H of its result; nothing is cached between steps.

    pipe = HostPipeline(D=3, n=n, low=0.0, high=1.0, periodic=True, radius=r, kernel=K.inv_dist(0.1))
    for k in range(steps):
        pipe.submit(pos_host[k], b_host[k], y_host[k])   # returns after enqueueing
    pipe.wait()                                          # all y_host[...] are valid

torch is used for device memory, streams and events only.
"""
import torch

from .particles import Particles, create_sparse_operator


class _Slot:
    pass


class HostPipeline:
    def __init__(self, D, n, low, high, periodic, radius, kernel, n_particles_in_leaf=10.0, depth=3, device=None):
        self.D, self.n = D, n
        self.low, self.high, self.periodic, self.n_leaf = low, high, periodic, float(n_particles_in_leaf)
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.slots = []
        self.k = 0
        self._last = None
        for _ in range(depth):
            s = _Slot()
            s.stream = torch.cuda.Stream(self.device)
            with torch.cuda.stream(s.stream):
                s.p = Particles(D, n, device=self.device)
                s.op = create_sparse_operator(s.p, s.p, radius, kernel)
                s.b = torch.empty(n * kernel.block_cols, dtype=torch.float64, device=self.device)
                s.y = torch.empty(n * kernel.block_rows, dtype=torch.float64, device=self.device)
            s.done = None
            s.pending = None
            self.slots.append(s)

    def submit(self, pos_host, b_host, y_host):
        """One step: upload positions and b, init_neighbour_search, y = K b, download y.
        Returns after enqueueing the uploads and the build of THIS step and the product
        and download of the PREVIOUS one, so the host never sits between a step's copies
        and the next step's.  The build runs in its asynchronous form; its alive count is
        verified (check_async) before the product of that step is enqueued — a step in
        which a particle left the domain raises instead of multiplying."""
        s = self.slots[self.k % len(self.slots)]
        self.k += 1
        if s.pending is not None:
            self._finish(s)
        if s.done is not None:
            s.done.synchronize()  # the slot's buffers (and the caller's y_host of that step) are free again
        with torch.cuda.stream(s.stream):
            s.p.resize_from_positions(pos_host)                      # H2D positions
            s.b.copy_(torch.as_tensor(b_host), non_blocking=True)     # H2D b
            s.p.init_neighbour_search(self.low, self.high, self.periodic, self.n_leaf, assume_all_alive=True)
        s.pending = y_host
        prev, self._last = self._last, s
        if prev is not None and prev is not s and prev.pending is not None:
            self._finish(prev)
        return s

    def _finish(self, s):
        y_host, s.pending = s.pending, None
        with torch.cuda.stream(s.stream):
            s.p.check_async()                                         # build done, nobody died
            s.op.matvec(s.b, out=s.y)
            torch.as_tensor(y_host).copy_(s.y, non_blocking=True)     # D2H y
            s.done = torch.cuda.Event()
            s.done.record(s.stream)

    def wait(self):
        for s in self.slots:
            if s.pending is not None:
                self._finish(s)
        for s in self.slots:
            if s.done is not None:
                s.done.synchronize()
