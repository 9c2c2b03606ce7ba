"""world_size-2 gloo tests of the multi-GPU host logic (sctda_b200/parallel.py): gene shards of
the permutation test and patch shards of the Mapper build reassemble to the single-rank result.
The per-rank compute is the CPU oracle here (tests may use it); on the box it is the CUDA path."""
import os
import socket

import numpy
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, ws, port, fn, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=ws)
    try:
        res = fn(rank, ws)
        numpy.save(os.path.join(out_dir, 'r%d.npy' % rank), numpy.asarray(res, dtype=numpy.int64))
    finally:
        dist.destroy_process_group()


def _run(fn, out_dir, ws=2):
    mp.spawn(_worker, args=(ws, _free_port(), fn, str(out_dir)), nprocs=ws, join=True)
    return [numpy.load(os.path.join(str(out_dir), 'r%d.npy' % r)) for r in range(ws)]


def _conn_case():
    from oracle import conn_oracle as co
    from sctda_b200 import synth
    members, adj, Y = synth.connectivity_case(300, 30, 11, seed=21)
    st = co.GraphState([str(k) for k in range(30)], adj.toarray(), {str(k): [int(c) for c in members[k]] for k in range(30)},
                       {'g%d' % i: Y[i] for i in range(11)})
    perms = co.perm_indices(24, 300, numpy.random.RandomState(3))
    return co, st, perms


def _conn_fn(rank, ws):
    from sctda_b200 import parallel
    co, st, perms = _conn_case()

    def compute(lo, hi):
        with numpy.errstate(invalid='ignore', divide='ignore'):
            return torch.tensor([int(round(24 * co.connectivity_pvalue(st, 'g%d' % i, n=24, perms=perms))) for i in range(lo, hi)],
                                dtype=torch.int32)
    return parallel.gene_sharded_counts(11, compute).numpy()


def test_gene_shards_reassemble(tmp_path):
    co, st, perms = _conn_case()
    with numpy.errstate(invalid='ignore', divide='ignore'):
        ref = [int(round(24 * co.connectivity_pvalue(st, 'g%d' % i, n=24, perms=perms))) for i in range(11)]
    outs = _run(_conn_fn, tmp_path)
    for o in outs:
        assert list(o) == ref


def _mapper_case():
    from oracle import mapper_oracle as mo
    from sctda_b200 import synth
    X, _, _ = synth.expression_matrix(260, 30, seed=5)
    Xc = X - X.mean(axis=0)
    u, s, _ = numpy.linalg.svd(Xc, full_matrices=False)
    lens = u[:, :2] * s[:2]
    bins, _ = mo.cover(lens, 5, 0.6, True)
    off = numpy.concatenate([[0], numpy.cumsum([len(b) for b in bins])])
    members = numpy.concatenate(bins)
    Xn, _ = mo.row_normalize(X, 'euclidean')
    return mo, Xn, off, members


def _cluster_sub(mo, Xn, sub_off, sub_members):
    labels, K = [], []
    for i in range(len(sub_off) - 1):
        cells = sub_members[sub_off[i]:sub_off[i + 1]]
        if len(cells) == 0:
            K.append(0)
            continue
        lab, k, _ = mo.cluster_bin(mo.gram(Xn, cells, cells))
        labels += list(lab)
        K.append(k)
    return numpy.array(labels, dtype=numpy.int32), numpy.array(K, dtype=numpy.int32)


def _mapper_fn(rank, ws):
    from sctda_b200 import parallel
    mo, Xn, off, members = _mapper_case()
    labels, K = parallel.patch_sharded_labels(off, members, lambda so, sm: _cluster_sub(mo, Xn, so, sm))
    return numpy.concatenate([labels, K])


def test_patch_shards_reassemble(tmp_path):
    mo, Xn, off, members = _mapper_case()
    labels, K = _cluster_sub(mo, Xn, off, members)
    outs = _run(_mapper_fn, tmp_path)
    for o in outs:
        assert list(o) == list(labels) + list(K)


def test_lpt_assignment_and_shards():
    from sctda_b200 import parallel
    sizes = [5, 0, 100, 7, 7, 50, 0, 49, 1]
    parts = parallel.assign_patches_lpt(sizes, 3)
    assert sorted(numpy.concatenate(parts).tolist()) == list(range(9))
    loads = [sum(sizes[b] ** 2 for b in p) for p in parts]
    assert max(loads) == 100 ** 2                      # the largest patch alone on its rank
    assert parallel.assign_patches_lpt(sizes, 3)[0].tolist() == parts[0].tolist()   # deterministic
    assert [parallel.shard_range(10, r, 4) for r in range(4)] == [(0, 2), (2, 5), (5, 7), (7, 10)]
    off = numpy.array([0, 2, 2, 5])
    so, sm = parallel.sub_csr(off, numpy.array([9, 8, 7, 6, 5]), numpy.array([0, 2]))
    assert so.tolist() == [0, 2, 5] and sm.tolist() == [9, 8, 7, 6, 5]
    so, sm = parallel.sub_csr(off, numpy.array([9, 8, 7, 6, 5]), numpy.array([1]))
    assert so.tolist() == [0, 0] and sm.tolist() == []


@pytest.mark.parametrize('ws', [1, 2, 3, 8])
def test_patch_shard_plan_places_every_fragment(ws):
    """parallel.patch_shard_plan (the host plan of the device-side exchange, sctda_allgather_csr): emulating
    every rank's label fragment and the placement kernel in numpy reassembles the full label array, for
    ragged and empty patches; fragments never exceed max_frag / max_nb."""
    from sctda_b200 import parallel
    rng = numpy.random.RandomState(ws)
    sizes = rng.randint(0, 40, 57)
    sizes[[3, 9]] = 0
    sizes[5] = 400
    bin_off = numpy.concatenate([[0], numpy.cumsum(sizes)])
    truth = rng.randint(0, 5, int(bin_off[-1])).astype(numpy.int32)
    truth_K = rng.randint(0, 6, sizes.size).astype(numpy.int32)
    plans = [parallel.patch_shard_plan(sizes, r, ws) for r in range(ws)]
    max_frag, max_nb = plans[0]['max_frag'], plans[0]['max_nb']
    gathered = numpy.zeros((ws, max(max_frag, 1)), dtype=numpy.int32)
    gathered_K = numpy.zeros((ws, max(max_nb, 1)), dtype=numpy.int32)
    seen = numpy.zeros(sizes.size, dtype=int)
    for r, p in enumerate(plans):
        assert (p['owner'] == plans[0]['owner']).all() and p['max_frag'] == max_frag and p['max_nb'] == max_nb
        assert p['sub_off'][-1] <= max_frag and p['mine'].size <= max_nb
        assert (numpy.diff(p['mine']) > 0).all()
        for i, b in enumerate(p['mine']):                          # the rank's sub-CSR and its fragment
            assert p['owner'][b] == r and p['frag_idx'][b] == i and p['frag_off'][b] == p['sub_off'][i]
            gathered[r, p['sub_off'][i]:p['sub_off'][i + 1]] = truth[bin_off[b]:bin_off[b + 1]]
            gathered_K[r, i] = truth_K[b]
            seen[b] += 1
    assert (seen == 1).all()
    owner, frag_off, frag_idx = plans[0]['owner'], plans[0]['frag_off'], plans[0]['frag_idx']
    labels = numpy.empty_like(truth)
    K = numpy.empty_like(truth_K)
    for b in range(sizes.size):                                    # place_fragments_kernel
        labels[bin_off[b]:bin_off[b + 1]] = gathered[owner[b], frag_off[b]:frag_off[b] + sizes[b]]
        K[b] = gathered_K[owner[b], frag_idx[b]]
    assert (labels == truth).all() and (K == truth_K).all()


def test_tile_shard_positions_partition_the_planned_order():
    """Every position of the planned gene order belongs to exactly one rank; whole 128-gene tiles stay together and
    are dealt round-robin."""
    from sctda_b200 import parallel
    for n in (0, 1, 127, 128, 129, 1000, 20000):
        for ws in (1, 2, 3, 8):
            parts = [parallel.tile_shard_positions(n, r, ws) for r in range(ws)]
            allpos = numpy.sort(numpy.concatenate(parts)) if n else numpy.zeros(0, dtype=numpy.int64)
            assert allpos.tolist() == list(range(n))
            for r, p in enumerate(parts):
                assert ((p // 128) % ws == r).all()
                assert (numpy.diff(p) > 0).all()
            sizes = [p.size for p in parts]
            assert max(sizes) - min(sizes) <= 128


def test_patch_dealer_charges_the_chain_of_the_largest_patch():
    """assign_patches_lpt: every patch goes to exactly one rank; with the chain term the rank that holds the largest
    patch (the longest MST chain) carries fewer Gram-tile entries than the others, without it the m^2 loads are level."""
    from sctda_b200 import parallel
    rs = numpy.random.RandomState(7)
    sizes = numpy.concatenate([[15432, 13027, 12502, 11669, 11490, 11317, 11187, 10546], rs.randint(4200, 10000, 48),
                               rs.randint(50, 4000, 800), [0, 0, 1]])
    for chain in (False, True):
        parts = parallel.assign_patches_lpt(sizes, 8, chain=chain)
        allp = numpy.sort(numpy.concatenate(parts))
        assert allp.tolist() == list(range(sizes.size))
        load = numpy.array([float((sizes[p].astype(numpy.float64) ** 2).sum()) for p in parts])
        top = int(numpy.argmax([sizes[p].max() for p in parts]))
        if chain:
            assert load[top] < numpy.delete(load, top).min()
            assert load.max() / load.min() < 1.25
        else:
            assert load.max() / load.min() < 1.02
    assert [p.tolist() for p in parallel.assign_patches_lpt(sizes, 8)] == [p.tolist() for p in parallel.assign_patches_lpt(sizes, 8)]
