"""N>1 host logic on CPU: world_size-2 gloo run of the sharding / counter
all-reduce / ordered gather used by the multi-GPU path.  The per-shard work is
done here by the oracle (no GPU in this container); on the GPU box the same
plumbing carries the CUDA results (bench.py --gpus N)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from centrifuger_b200 import distributed as D

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_exactly_once():
    for n in (0, 1, 2, 7, 100, 1001):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                lo, hi = D.shard_bounds(n, r, world)
                assert 0 <= lo <= hi <= n
                seen += list(range(lo, hi))
            assert seen == list(range(n))
            sizes = [D.shard_bounds(n, r, world)[1] - D.shard_bounds(n, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, idx, fq1, fq2, q):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_binding import Oracle, read_fastx
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ids, r1 = read_fastx(fq1)
    _, r2 = read_fastx(fq2)
    lo, hi = D.shard_bounds(len(ids), rank, world)
    o = Oracle(idx, k=5)
    node_cnt = o.scalar(10)
    counts = torch.zeros(node_cnt + 3, dtype=torch.int64)
    rows = []
    for i in range(lo, hi):
        res = o.query(r1[i], r2[i])
        rows.append(o.format_tsv(ids[i], res))
        counts[node_cnt + 1] += 1
        if res.n > 0:
            counts[node_cnt + 2] += 1
            for j in range(min(res.n, 5)):
                ct = res.ids[j] if res.by_rank else o.L.cfr_oracle_seqid_to_taxid(o.h, res.ids[j])
                counts[min(int(ct), node_cnt)] += 1
    D.allreduce_counts(counts)
    gathered = D.gather_in_order("".join(rows))
    if rank == 0:
        q.put((counts.numpy().copy(), "".join(gathered)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single_process(tiny_dir):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_binding import Oracle, read_fastx
    idx = os.path.join(tiny_dir, "idx")
    fq1, fq2 = os.path.join(tiny_dir, "pe_100_1.fq"), os.path.join(tiny_dir, "pe_100_2.fq")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, idx, fq1, fq2, q)) for r in range(2)]
    for p in procs:
        p.start()
    counts, tsv = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process reference
    ids, r1 = read_fastx(fq1)
    _, r2 = read_fastx(fq2)
    o = Oracle(idx, k=5)
    exp_tsv = o.classify_tsv(ids, r1, r2, header=False)
    assert tsv == exp_tsv
    node_cnt = o.scalar(10)
    assert int(counts[node_cnt + 1]) == len(ids)
    n_rows = len(exp_tsv.strip().split("\n"))
    n_uncls = sum(1 for l in exp_tsv.split("\n") if "\tunclassified\t" in l)
    assert int(counts[node_cnt + 2]) == len(ids) - n_uncls
    assert int(counts[:node_cnt + 1].sum()) == n_rows - n_uncls


def test_shard_packed_rebases_offsets():
    reads = [b"ACGT", b"AA", b"", b"GGGTTT", b"C"]
    off = np.zeros(len(reads) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(r) for r in reads])
    seq = np.frombuffer(b"".join(reads), dtype=np.uint8)
    got = []
    for r in range(2):
        s, o, lo, hi = D.shard_packed(seq, off, r, 2)
        assert o[0] == 0
        got += [s[int(o[i]):int(o[i + 1])].tobytes() for i in range(hi - lo)]
    assert got == reads
