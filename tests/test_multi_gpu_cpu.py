"""World-size-2 tests of the library-sharding plumbing on CPU (gloo): shard bounds, the gather-v
of hit records, and that merging per-shard results equals the single-shard result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from barcoder_b200 import multi_gpu, synth
from oracle import oracle


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 8, 9, 10_000_001):
        for w in (1, 2, 3, 8):
            b = [multi_gpu.shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, tmpdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # every rank searches its shard of the same library with the CPU oracle (stand-in for the
        # per-GPU search; the gather logic is what is under test)
        genome, off = synth.random_genome(40_000, seed=5, n_contigs=2)
        lib = synth.random_library(300, 20, seed=6)
        synth.plant(lib, genome, 0.5, 2, seed=7)
        contigs = [bytes(genome[int(off[i]):int(off[i + 1])]) for i in range(2)]
        lo, hi = multi_gpu.shard_bounds(len(lib), world, rank)
        hits = oracle.search(contigs, synth.rows_to_strings(lib[lo:hi]), 2, pam="NGG", threads=1)
        hits["spacer_id"] += lo  # what BC_PARAM_SPACER_ID_BASE does on the device
        local = torch.from_numpy(hits.view(np.uint32).reshape(-1, 4).astype(np.int64).astype(np.int32)
                                 if False else hits.view(np.int32).reshape(-1, 4).copy())
        merged, counts = multi_gpu.gather_hits(local)
        assert counts[rank] == len(hits) and len(counts) == world
        if rank == 0:
            got = oracle.canonical_sort(multi_gpu.records_from_tensor(merged))
            full = oracle.search(contigs, synth.rows_to_strings(lib), 2, pam="NGG", threads=1)
            assert got.tobytes() == full.tobytes()
            with open(os.path.join(tmpdir, "ok"), "w") as h:
                h.write(str(len(full)))
        else:
            assert merged is None
        # empty shard on one rank
        empty = torch.zeros((0, 4), dtype=torch.int32) if rank == 1 else local
        merged2, counts2 = multi_gpu.gather_hits(empty)
        assert counts2[1] == 0
        if rank == 0:
            assert merged2.shape[0] == counts2[0]
    finally:
        dist.destroy_process_group()


def test_gather_hits_world_size_2(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert int(open(tmp_path / "ok").read()) > 50


def _worker_streamed(rank, world, port, tmpdir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(100 + rank)
        n_pieces = 3 if rank == 0 else 5               # ranks report different numbers of pieces
        sizes = [int(x) for x in rng.integers(0, 700, n_pieces)]
        if rank == 1:
            sizes[2] = 0                               # an empty piece in the middle
        pieces = [torch.from_numpy(rng.integers(0, 2**31 - 1, (sz, 4)).astype(np.int32)) for sz in sizes]
        sg = multi_gpu.StreamedGather(torch.device("cpu"), capacity=64)   # small: forces growth on dst
        for p in pieces:
            assert sg.push(p) is False
        empty = torch.empty((0, 4), dtype=torch.int32)
        while not sg.push(empty, done=True):
            pass
        merged, per_rank = sg.finish()
        mine = torch.cat(pieces)
        totals = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(totals, torch.tensor([mine.shape[0]], dtype=torch.int64))
        assert per_rank == [int(t.item()) for t in totals]
        # rank 0 checks content: the multiset of records equals the union of all ranks' pieces
        gathered = [None] * world
        dist.all_gather_object(gathered, mine.numpy())
        if rank == 0:
            want = np.concatenate(gathered)
            got = merged.numpy()
            assert got.shape == want.shape
            key = lambda a: a[np.lexsort(a.T[::-1])]
            assert (key(got) == key(want)).all()
            with open(os.path.join(tmpdir, "ok2"), "w") as h:
                h.write(str(len(want)))
        else:
            assert merged is None
    finally:
        dist.destroy_process_group()


def test_streamed_gather_world_size_2(tmp_path):
    port = _free_port()
    mp.spawn(_worker_streamed, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert int(open(tmp_path / "ok2").read()) > 100
