"""world_size-2 test of the N>1 path on CPU (gloo): graphs are sharded by rank, each rank renders its shard, the PCM
is gathered on rank 0 and must equal the single-process render.  The renderer here is the CPU oracle (test
infrastructure) — what is under test is the sharding / gather / max-over-ranks plumbing that bench.py uses with NCCL."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
N_GRAPHS, LENGTH = 7, 128 * 9 + 40


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _render_shard(g0, g1):
    import ctypes
    from conftest import ROOT, load_package
    import graphs as G
    pkg = load_package()
    api = pkg.Api(ctypes.CDLL(os.path.join(ROOT, "oracle", "_build", "liboracle.so")), "wao_")
    be = pkg.context.Backend(api)
    ctxs = [G.c2_buffer_biquad_gain(pkg, be, g, LENGTH) for g in range(g0, g1)]
    return pkg, torch.from_numpy(G.render(pkg, ctxs)) if ctxs else None


def _worker(rank, world, port, out_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from conftest import load_package
    pkg = load_package()
    g0, g1 = pkg.parallel.shard_range(N_GRAPHS, rank, world)
    _, local = _render_shard(g0, g1)
    full = pkg.parallel.gather_pcm(local, N_GRAPHS, dst=0)
    # the grouped all-gather of bench.py's "gather inside the step": every rank contributes the same local graph range per group
    per_rank = N_GRAPHS // world
    mine = local[:per_rank]
    for g0, g1 in pkg.parallel.group_ranges(per_rank, 2):
        full_k = torch.empty((world, g1 - g0) + tuple(mine.shape[1:]), dtype=mine.dtype)
        pkg.parallel.all_gather_group(full_k, mine[g0:g1])
        assert torch.equal(full_k[rank], mine[g0:g1])
        other = torch.tensor([float(full_k[r].abs().sum()) for r in range(world)])
        assert (other > 0).all()
    slowest = pkg.parallel.max_over_ranks(10.0 + rank)
    assert slowest == 10.0 + world - 1
    if rank == 0:
        np.save(out_path, full.numpy())
    else:
        assert full is None
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_cover_everything(pkg):
    for n in (1, 7, 1000):
        for world in (1, 2, 3, 8):
            r = [pkg.parallel.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def test_two_rank_render_and_gather(tmp_path, oracle):
    out = str(tmp_path / "gathered.npy")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    _, want = _render_shard(0, N_GRAPHS)
    assert got.shape == (N_GRAPHS, 2, LENGTH)
    assert np.array_equal(got, want.numpy())
