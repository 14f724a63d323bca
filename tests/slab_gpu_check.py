"""Run under torchrun (one rank per GPU): the slab-decomposed build + product
against the single-GPU path on the same global particle set.
  - concatenating the ranks' owned ranges reproduces the single-GPU sorted
    order, bucket by bucket (ids, positions bit-exact);
  - y agrees by particle id within 1e-12 relative L2 (tiled kernel on both)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import aboria_b200 as ab  # noqa: E402
from aboria_b200 import kernels as K  # noqa: E402
from aboria_b200 import slab, synth  # noqa: E402


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    dist.init_process_group("nccl", device_id=dev)
    ok_all = True
    for (N, periodic, rfac) in [(200_000, True, 1.0), (200_000, False, 1.0), (60_000, True, 1.7)]:
        pos = synth.uniform_positions(N, 3)
        b = synth.vector(N)
        # single-GPU reference on every rank (cheap at this size)
        p1 = ab.Particles(3, N, device=dev)
        p1.set("position", torch.from_numpy(pos.copy()))
        p1.init_neighbour_search(0.0, 1.0, periodic)
        size, side, _ = p1.grid()
        radius = rfac * float(side[0])
        op1 = ab.create_sparse_operator(p1, p1, radius, K.inv_dist(0.1))
        ids1 = p1.get("id").cpu().numpy()
        y1 = (op1 * torch.from_numpy(b[ids1]).to(dev)).cpu().numpy()
        y1_by_id = np.zeros(N)
        y1_by_id[ids1] = y1
        # slab path
        sp = slab.SlabParticles(3, 0.0, 1.0, periodic, N, 10.0, radius, rank, world, dev)
        assert list(sp.size) == list(size)
        layer = np.floor((pos[:, 0] - 0.0) * (1.0 / side[0])).astype(np.int64)
        mine = np.nonzero((layer >= sp.lo_layer) & (layer < sp.hi_layer))[0]
        ids_t = torch.from_numpy(mine.astype(np.int64)).to(dev)
        sp.build(torch.from_numpy(pos[mine].copy()).to(dev))
        # ids follow the particles: owned sort order, then halo exchange
        ids_local = sp.ex.assemble(ids_t[sp.order_owned.long()])
        op = ab.create_sparse_operator(sp.p, sp.p, radius, K.inv_dist(0.1))
        b_local = torch.from_numpy(b).to(dev)[ids_local]
        b_check = b_local.clone()
        b_local[: sp.ex.own_begin] = 0
        b_local[sp.ex.own_end:] = 0
        y = sp.matvec(op, b_local)  # fills the b halo from the neighbours
        assert torch.equal(b_local, b_check), "b halo exchange"
        own_ids = sp.owned(ids_local).cpu().numpy()
        y_own = sp.owned(y).cpu().numpy()
        # owned range == the single-GPU sorted order restricted to my layers
        sel = np.nonzero((layer[ids1] >= sp.lo_layer) & (layer[ids1] < sp.hi_layer))[0]
        same_order = np.array_equal(ids1[sel], own_ids)
        same_pos = np.array_equal(sp.owned(sp.p.get("position")).cpu().numpy(), p1.get("position").cpu().numpy()[sel])
        err = np.linalg.norm(y_own - y1_by_id[own_ids]) / np.linalg.norm(y1_by_id[own_ids])
        cnt, _ = sp.p.pair_stats(radius)
        cnt1, _ = p1.pair_stats(radius)
        same_cnt = np.array_equal(sp.owned(cnt).cpu().numpy(), cnt1.cpu().numpy()[sel])
        # host-buffer steps through SlabHostPipeline: every step equals the device-resident step
        pos_h = torch.from_numpy(pos[mine].copy()).pin_memory()
        b_owned_sorted = sp.owned(b_check).clone()
        b_h = b_owned_sorted.cpu().pin_memory()  # b is indexed by post-reorder position; the order is the same every step
        y_h = [torch.empty(len(mine), dtype=torch.float64).pin_memory() for _ in range(3)]
        pipe = slab.SlabHostPipeline(sp, op, len(mine), dev)
        for k in range(3):
            pipe.submit(pos_h, b_h, y_h[k])
        pipe.wait()
        same_pipe = all(torch.equal(yh, sp.owned(y).cpu()) for yh in y_h)
        ok = same_order and same_pos and same_cnt and err <= 1e-12 and same_pipe
        print(f"[rank {rank}] N={N} periodic={periodic} r={rfac}*side w={sp.w} layers={sp.lo_layer}..{sp.hi_layer} own={len(own_ids)} "
              f"ghost={sp.ex.n_ghost_lo}+{sp.ex.n_ghost_hi} order={same_order} pos={same_pos} counts={same_cnt} pipeline={same_pipe} rel_l2={err:.2e} -> {'OK' if ok else 'FAIL'}", flush=True)
        ok_all = ok_all and ok
    flag = torch.tensor([1 if ok_all else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
