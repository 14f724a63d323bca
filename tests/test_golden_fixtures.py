"""The committed golden fixtures (tests/golden/reference_kats.json: the literal known-answer values
of the reference's own tests) against the CPU oracle, and against the CUDA path through the C-ABI."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as orc

G = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_kats.json")))


def _lattice(D, n):
    idx = np.indices((n,) * D).reshape(D, -1).T[:, ::-1]
    return idx.astype(np.float64) + 0.5


def test_oracle_against_golden_fixtures():
    for k in G["collapse_index_vector"]:
        assert orc.collapse_index_vector(k["size"], k["vindex"]) == k["expect"], k["source"]
    for k in G["point_to_bucket_index"]:
        o = orc.Oracle(len(k["point"]))
        o.force_grid(k["low"], k["high"], False, k["size"])
        idx, v = o.point_to_bucket_index(k["point"])
        assert list(v) == k["vindex"] and idx == k["index"], k["source"]
    for k in G["buckets_near_point"]:
        D = k["D"]
        o = orc.Oracle(D)
        # the reference test fills the unit box with n_particles random particles: n_leaf = 10 gives the grid
        o.init_neighbour_search(np.random.default_rng(0).random((k["n_particles"], D)), 0.0, 1.0, False)
        assert list(o.grid()[0]) == [k["grid"]] * D
        for point, r, expect in k["cases"]:
            assert o.buckets_near_point(point, r)[0] == expect, (k["source"], point, r)
    s = G["single_particle_search"]
    o = orc.Oracle(3)
    o.init_neighbour_search([s["particle"]], s["low"], s["high"], s["periodic"])
    for q, expect in s["queries"]:
        assert o.search_point(q, s["radius"])[0] == expect, s["source"]
    sc = s["scaled"]
    cnt, _ = o.pair_stats_norm(np.array([q for q, _ in sc["queries"]]), sc["radius"], 2, scale=sc["scale"])
    assert cnt.tolist() == [e for _, e in sc["queries"]], s["source"]
    for c in G["lattice_gauss_circle"]["cases"]:
        pos = _lattice(c["D"], c["n"])
        o = orc.Oracle(c["D"])
        out = o.init_neighbour_search(pos, 0.0, float(c["n"]), True)
        cnt, _ = o.pair_stats(out["pos"], c["r"])
        assert np.all(cnt == c["count"]), (G["lattice_gauss_circle"]["source"], c)
    k = G["sparse_operator"]
    o = orc.Oracle(3)
    out = o.init_neighbour_search(np.array(k["positions"]), k["low"], k["high"], k["periodic"])
    assert list(out["order"]) == [0, 1, 2]
    s1, s2, v = np.full(3, k["s1"]), np.full(3, k["s2"]), np.array(k["v"])
    y, npairs = o.sparse_matvec(out["pos"], orc.K_CONST_SUM, [], k["diameter"], v, row_vars=[s1], col_vars=[s2])
    assert list(y) == k["y_const_sum"] and npairs == k["nonzeros"], k["source"]
    y2, np2 = o.sparse_matvec(out["pos"], orc.K_CONST_SUM_DIFF, [], k["diameter"], v, BR=2, BC=1, row_vars=[s1], col_vars=[s2])
    assert list(y2[:3]) == k["y_const_sum_diff_first_n"] and 2 * np2 == k["nonzeros_2x1"]
    ii, jj = np.divmod(np.arange(9), 3)
    dense = o.coeff(out["pos"], out["pos"], ii, jj, orc.K_CONST_SUM, [], k["diameter"], row_vars=[s1], col_vars=[s2]).reshape(3, 3)
    assert dense.tolist() == k["dense"]
    i = G["id_search"]
    ids = np.random.default_rng(0).permutation(i["N"]).astype(np.uint64)
    key, value = orc.id_map_build(ids)
    f = orc.id_find(key, value, [i["find"], i["missing"]])
    assert ids[int(f[0])] == i["find"] and f[1] == i["N"], i["source"]


@pytest.mark.gpu
def test_cuda_path_against_golden_fixtures():
    import torch

    import aboria_b200 as ab
    from aboria_b200 import kernels as K

    for k in G["point_to_bucket_index"]:
        p = ab.Particles(3, 1)
        p.set("position", torch.tensor([k["point"]], dtype=torch.float64))
        p.force_grid(k["low"], k["high"], False, k["size"])
        assert p.get_query().bucket_indices.cpu().tolist() == [k["index"]], k["source"]
    s = G["single_particle_search"]
    p = ab.Particles(3, 1)
    p.set("position", torch.tensor([s["particle"]], dtype=torch.float64))
    p.init_neighbour_search(s["low"], s["high"], s["periodic"])
    q = np.array([q for q, _ in s["queries"]])
    cnt, _ = p.distance_search_stats(s["radius"], 2, queries=q)
    assert cnt.cpu().tolist() == [e for _, e in s["queries"]], s["source"]
    sc = s["scaled"]
    cnt, _ = p.distance_search_stats(sc["radius"], 2, queries=q, scale=sc["scale"])
    assert cnt.cpu().tolist() == [e for _, e in sc["queries"]], s["source"]
    for c in G["lattice_gauss_circle"]["cases"]:
        pos = _lattice(c["D"], c["n"])
        p = ab.Particles(c["D"], len(pos))
        p.set("position", torch.from_numpy(pos))
        p.init_neighbour_search(0.0, float(c["n"]), True)
        for path in (0, 1):
            cnt, _ = p.pair_stats(c["r"], path=path)
            assert bool((cnt == c["count"]).all()), (c, path)
    k = G["sparse_operator"]
    p = ab.Particles(3, 3, variables={"s1": torch.float64, "s2": torch.float64})
    p.set("position", torch.tensor(k["positions"], dtype=torch.float64))
    p.set("s1", torch.full((3,), k["s1"], dtype=torch.float64))
    p.set("s2", torch.full((3,), k["s2"], dtype=torch.float64))
    p.init_neighbour_search(k["low"], k["high"], k["periodic"])
    assert p.get("id").cpu().tolist() == [0, 1, 2]
    v = torch.tensor(k["v"], dtype=torch.float64, device=p.device)
    C = ab.create_sparse_operator(p, p, k["diameter"], K.const_sum("s1", "s2"))
    assert (C * v).cpu().tolist() == k["y_const_sum"], k["source"]
    C2 = ab.create_sparse_operator(p, p, k["diameter"], K.const_sum_diff("s1", "s2"))
    assert (C2 * v).cpu().tolist()[:3] == k["y_const_sum_diff_first_n"]
    rp, col, val = C.assemble()
    assert col.shape[0] == k["nonzeros"]
    rp2, col2, val2 = C2.assemble()
    assert col2.shape[0] * 2 == k["nonzeros_2x1"]
    ii, jj = np.divmod(np.arange(9), 3)
    assert C.coeff(ii, jj).cpu().reshape(3, 3).tolist() == k["dense"]
    i = G["id_search"]
    ids = np.random.default_rng(0).permutation(i["N"]).astype(np.int64)
    p = ab.Particles(3, i["N"])
    p.set("id", torch.from_numpy(ids))
    p.init_id_search()
    f = p.get_query().find([i["find"], i["missing"]]).cpu().tolist()
    assert ids[f[0]] == i["find"] and f[1] == i["N"], i["source"]
