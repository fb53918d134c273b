"""World-size-2 gloo tests (CPU) of the multi-GPU exchange logic in genomad_b200.dist."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from genomad_b200 import dist as gdist
from oracle import tokenizer as T


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmp, n_windows, offsets):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    info = gdist.init_process_group_if_needed("gloo")
    assert (info.rank, info.world_size) == (rank, world)
    rng = np.random.default_rng(1)
    probs = rng.random((n_windows, 3)).astype(np.float32)             # what one GPU would produce for ALL windows
    s, e = gdist.shard_bounds(n_windows, world, rank)
    local = torch.from_numpy(probs[s:e].copy())
    full = gdist.gather_window_probs(local, n_windows, world)
    assert torch.equal(full, torch.from_numpy(probs))                  # bitwise: same order as single GPU
    # allreduce variant: per-rank partial sums over the local shard
    loc_off = gdist.local_offsets(offsets, s, e)
    part = torch.zeros((len(offsets) - 1, 4))
    for c in range(len(offsets) - 1):
        seg = local[loc_off[c]:loc_off[c + 1]]
        part[c, :3] = seg.sum(0)
        part[c, 3] = seg.shape[0]
    red = gdist.allreduce_partials(part, world)
    mean = gdist.finish_mean(red).numpy()
    ref = T.segment_mean(probs, np.repeat(np.arange(len(offsets) - 1), np.diff(offsets)), len(offsets) - 1)
    assert np.abs(mean - ref).max() < 1e-6
    np.save(os.path.join(tmp, f"ok{rank}.npy"), mean)
    dist.destroy_process_group()


@pytest.mark.parametrize("n_windows,counts", [(11, [1, 5, 2, 3]), (4, [4]), (7, [1, 1, 1, 1, 1, 1, 1]), (3, [2, 0, 1])])
def test_world2_gather_and_allreduce(tmp_path, n_windows, counts):
    offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    assert offsets[-1] == n_windows
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path), n_windows, offsets), nprocs=2, join=True)
    a, b = np.load(tmp_path / "ok0.npy"), np.load(tmp_path / "ok1.npy")
    assert np.array_equal(a, b)


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 8, 167000, 1_000_003):
        for world in (1, 2, 4, 8):
            spans = [gdist.shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1


def test_local_offsets_straddling_contig():
    offsets = np.array([0, 167, 334, 501], np.int32)       # three 1 Mb contigs (167 windows each)
    assert gdist.local_offsets(offsets, 0, 251).tolist() == [0, 167, 251, 251]
    assert gdist.local_offsets(offsets, 251, 501).tolist() == [0, 0, 83, 250]
