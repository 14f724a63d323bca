"""Run under torchrun (one rank per GPU): the slab-decomposed build + product against the
single-GPU path on the same global particle set.
  - concatenating the ranks' owned ranges reproduces the single-GPU cell list: same particles
    (ids) and positions bucket by bucket; bit-identical order when the rank received its
    particles in global index order;
  - per-row neighbour counts and pair-set hashes (by particle id) and y by particle id within
    1e-12 relative L2 (tiled kernel on both);
  - uniform, r > bucket side (two ghost layers), non-periodic, and a clustered cloud with the
    layer split balanced by particle count;
  - particle migration: every particle moves by up to 0.4 bucket sides, twice; the slabs
    migrate / rebuild and are compared with a single-GPU rebuild of the moved set.
Prints one line per case and rank ending in OK / FAIL; exit code 1 on any FAIL."""
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


def single_gpu(pos, periodic, dev, rfac, b):
    N = pos.shape[0]
    p1 = ab.Particles(3, N, device=dev)
    p1.set("position", torch.from_numpy(pos.copy()))
    p1.init_neighbour_search(0.0, 1.0, periodic)
    size, side, _ = p1.grid()
    radius = rfac * float(side[0])
    op1 = ab.create_sparse_operator(p1, p1, radius, K.inv_dist(0.1))
    ids1 = p1.get("id").cpu().numpy()
    y1 = (op1 * torch.from_numpy(b[ids1]).to(dev)).cpu().numpy()
    cnt1, hs1 = p1.pair_stats(radius)
    by_id = lambda v: (lambda out: (out.__setitem__(ids1, v), out)[1])(np.zeros(N, dtype=v.dtype))  # noqa: E731
    return p1, size, side, radius, ids1, by_id(y1), by_id(cnt1.cpu().numpy()), p1.get("position").cpu().numpy()


def compare(tag, rank, sp, op, radius, gid_local, b, ids1, y1_by_id, cnt1_by_id, pos1_sorted, layer_of_id, dev, strict_order):
    b_local = torch.from_numpy(b).to(dev)[gid_local]
    b_check = b_local.clone()
    b_local[: sp.ex.own_begin] = 0
    b_local[sp.ex.own_end:] = 0
    y = sp.matvec(op, b_local)  # fills the b halo from the neighbours
    halo_ok = bool(torch.equal(b_local, b_check))
    own_ids = sp.owned(gid_local).cpu().numpy()
    y_own = sp.owned(y).cpu().numpy()
    sel = np.nonzero((layer_of_id[ids1] >= sp.lo_layer) & (layer_of_id[ids1] < sp.hi_layer))[0]
    same_set = np.array_equal(np.sort(ids1[sel]), np.sort(own_ids))
    same_order = np.array_equal(ids1[sel], own_ids)
    # bucket by bucket: the owned range holds the same buckets' particles in bucket order (positions sorted by id inside a bucket agree)
    pos_own = sp.owned(sp.p.get("position")).cpu().numpy()
    pos_by_id = np.zeros((len(layer_of_id), 3))
    pos_by_id[ids1] = pos1_sorted
    same_pos = np.array_equal(pos_own, pos_by_id[own_ids])
    q1_bucket = np.zeros(len(layer_of_id), dtype=np.int64)
    q1_bucket[ids1] = np.arange(len(ids1))  # rank in the single-GPU order: non-decreasing bucket
    err = np.linalg.norm(y_own - y1_by_id[own_ids]) / max(np.linalg.norm(y1_by_id[own_ids]), 1e-300)
    cnt, _ = sp.p.pair_stats(radius)
    same_cnt = np.array_equal(sp.owned(cnt).cpu().numpy(), cnt1_by_id[own_ids])
    ok = halo_ok and same_set and same_pos and same_cnt and err <= 1e-12 and (same_order or not strict_order)
    print(f"[rank {rank}] {tag} w={sp.w} layers={sp.lo_layer}..{sp.hi_layer} own={len(own_ids)} ghost={sp.ex.n_ghost_lo}+{sp.ex.n_ghost_hi} "
          f"set={same_set} order={same_order} pos={same_pos} counts={same_cnt} b_halo={halo_ok} rel_l2={err:.2e} -> {'OK' if ok else 'FAIL'}", flush=True)
    return ok, y


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    import datetime

    dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
    if os.environ.get("ABR_SLAB_TEST_CAP"):
        # force the "halo larger than the reserve" path on some ranks only (rank-dependent reserve)
        slab.SlabParticles._test_cap = lambda self, n_in: (64 if self.rank % 2 == 0 else n_in // 2 + 4096)
    ok_all = True
    cases = [("uniform", 200_000, True, 1.0), ("uniform", 200_000, False, 1.0), ("uniform", 60_000 * max(1, world // 2), True, 1.7),
             ("clustered", 400_000, [True, True, False], 1.0)]
    for kind, N, periodic, rfac in cases:
        pos = synth.uniform_positions(N, 3) if kind == "uniform" else synth.clustered_positions(N)
        b = synth.vector(N)
        p1, size, side, radius, ids1, y1_by_id, cnt1_by_id, pos1_sorted = single_gpu(pos, periodic, dev, rfac, b)
        sp = slab.SlabParticles(3, 0.0, 1.0, periodic, N, 10.0, radius, rank, world, dev)
        assert list(sp.size) == list(size)
        layer = np.floor((pos[:, 0] - 0.0) * (1.0 / side[0])).astype(np.int64)
        if kind == "clustered":
            # split balanced by particle count (the same on every rank: from the global histogram)
            mine0 = np.arange(rank, N, world)
            hist = sp.layer_histogram(torch.from_numpy(pos[mine0]).to(dev))
            assert int(hist.sum()) == N
            sp.set_layers(slab.plan_layers_balanced(hist, world))
            counts = [int(hist[lo:hi].sum()) for lo, hi in sp.layers]
            if rank == 0:
                print(f"[rank 0] clustered: balanced layer split {sp.layers} -> particles per rank {counts} (even split by layers would give "
                      f"{[int(hist[lo:hi].sum()) for lo, hi in slab.plan_layers(int(size[0]), world)]})", flush=True)
        mine = np.nonzero((layer >= sp.lo_layer) & (layer < sp.hi_layer))[0]
        gid = torch.from_numpy(mine.astype(np.int64)).to(dev)
        sp.build(torch.from_numpy(pos[mine].copy()).to(dev), extra_columns={"gid": gid})
        gid_local = sp.p.get("gid")
        op = ab.create_sparse_operator(sp.p, sp.p, radius, K.inv_dist(0.1))
        ok, y = compare(f"{kind} N={N} periodic={periodic} r={rfac}*side", rank, sp, op, radius, gid_local, b, ids1, y1_by_id, cnt1_by_id, pos1_sorted, layer, dev,
                        strict_order=True)
        ok_all = ok_all and ok
        if kind == "uniform" and rfac == 1.0:
            # host-buffer steps through SlabHostPipeline: every step equals the device-resident step
            pos_h = torch.from_numpy(pos[mine].copy()).pin_memory()
            b_owned_sorted = torch.from_numpy(b).to(dev)[sp.owned(gid_local)]
            b_h = b_owned_sorted.cpu().pin_memory()
            y_h = [torch.empty(len(mine), dtype=torch.float64).pin_memory() for _ in range(3)]
            y_ref = sp.owned(y).cpu()
            pipe = slab.SlabHostPipeline(sp, op, len(mine), dev)
            for k in range(3):
                pipe.submit(pos_h, b_h, y_h[k])
            pipe.wait()
            same_pipe = all(torch.equal(yh, y_ref) for yh in y_h)
            print(f"[rank {rank}] {kind} N={N} periodic={periodic}: host pipeline == device-resident step: {same_pipe} -> {'OK' if same_pipe else 'FAIL'}", flush=True)
            ok_all = ok_all and same_pipe
            # ---- migration: move every particle, migrate, rebuild; twice
            if periodic is True:
                cur_pos = torch.from_numpy(pos[mine].copy()).to(dev)
                cur_gid = gid.clone()
                gpos = pos.copy()
                rng = np.random.default_rng(1234)
                for it in range(2):
                    disp = rng.uniform(-0.4, 0.4, size=(N, 3)) * float(side[0])
                    gpos = gpos + disp  # may leave [0, 1): the build wraps it
                    cur_pos = cur_pos + torch.from_numpy(disp).to(dev)[cur_gid]
                    n_before = cur_pos.shape[0]
                    cur_pos, cols = sp.migrate(cur_pos, {"gid": cur_gid})
                    cur_gid = cols["gid"]
                    moved = n_before - int(torch.isin(cur_gid, torch.from_numpy(mine).to(dev)).sum()) if it == 0 else -1
                    p1, size, side, radius, ids1, y1_by_id, cnt1_by_id, pos1_sorted = single_gpu(gpos, periodic, dev, rfac, b)
                    wrapped = gpos - np.floor(gpos)
                    layer_m = np.floor((wrapped[:, 0] - 0.0) * (1.0 / side[0])).astype(np.int64)
                    layer_m = np.clip(layer_m, 0, int(size[0]) - 1)
                    sp.build(cur_pos, extra_columns={"gid": cur_gid})
                    ok, _ = compare(f"uniform N={N} migration step {it + 1} (rank now holds {cur_pos.shape[0]}, first-step arrivals {moved})", rank, sp, op, radius,
                                    sp.p.get("gid"), b, ids1, y1_by_id, cnt1_by_id, pos1_sorted, layer_m, dev, strict_order=False)
                    ok_all = ok_all and ok
    flag = torch.tensor([1 if ok_all else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
