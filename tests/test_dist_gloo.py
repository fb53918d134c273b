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


# ------------------------------------------------------------------------------------------ module driver under 2 ranks
class _Stub:
    device = 0
    max_batch = 64


def _stub_classify(clf, parsed, offsets, info, contig_reduce="gather"):
    """Stand-in for the GPU stage that keeps the real sharding + exchange (a collective: ranks that disagree on whether
    to classify would hang here)."""
    n = parsed.n_windows
    s, e = gdist.shard_bounds(n, info.world_size, info.rank)
    buf = np.zeros((max(1, e - s), 6000), np.uint8)
    win = parsed.export_windows(s, e - s, buf)
    x = win.astype(np.float64).sum(1)
    local = torch.from_numpy(np.stack([np.sin(x) ** 2, np.cos(x) ** 2 / 2, np.cos(x) ** 2 / 2], 1).astype(np.float32))
    full = gdist.gather_window_probs(local, n, info.world_size).numpy()
    return T.segment_mean(full, np.repeat(np.arange(len(offsets) - 1), np.diff(offsets)), len(offsets) - 1)


def _driver_worker(rank, world, port, tmp):
    from pathlib import Path
    from genomad_b200 import nn_classification, utils, _paths
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    gdist.init_process_group_if_needed("gloo")
    nn_classification._make_classifier = lambda batch_size, device: _Stub()
    nn_classification._classify_parsed = _stub_classify
    if rank != 0:
        # ranks other than 0 must not base control flow on the file system (skip-decision race)
        def boom(*a, **k):
            raise AssertionError("non-main rank consulted the file system for a control-flow decision")
        utils.check_provirus_execution = boom
        utils.compare_executions = boom
        utils.get_md5 = boom
    tmp = Path(tmp)
    fa, out = tmp / "in.fna", tmp / "out"
    o = _paths.NNOutputs("in", out)
    nn_classification.main(fa, out, False, 128, False, 2, False, False)          # fresh run
    if rank == 0:
        first = np.load(o.nn_classification_npz_output)["predictions"].copy()
        pv = np.load(o.provirus_nn_classification_npz_output)                    # the provirus twin ran on both ranks as well
        assert pv["provirus_names"].tolist() == ["c0|provirus_1_9000", "c3|provirus_101_6500"] and pv["predictions"].shape == (2, 3)
        assert np.load(o.provirus_window_id_output)["provirus_ids"].tolist() == [0, 0, 1]
    nn_classification.main(fa, out, False, 128, False, 2, False, False)          # skip run: nobody may enter the collective
    dist.barrier()
    if rank == 0:
        assert "Skipping sequence classification" in o.nn_classification_log.read_text()
        o.nn_classification_npz_output.unlink()                                   # rank 1 cannot know; rank 0 decides
    dist.barrier()
    nn_classification.main(fa, out, False, 128, False, 2, False, False)          # classification redone on BOTH ranks
    nn_classification.main(fa, out, True, 128, True, 2, False, True)             # restart + parameter change + cleanup
    if rank == 0:
        z = np.load(o.nn_classification_npz_output)
        assert z["predictions"].shape == first.shape and not o.encoded_sequences_dir.exists()
        np.save(tmp / "first.npy", first)
    dist.barrier()
    dist.destroy_process_group()


def test_world2_module_driver_rank0_decides(tmp_path):
    rng = np.random.default_rng(2)
    with open(tmp_path / "in.fna", "w") as fh:
        for i, ln in enumerate([20000, 6100, 3000, 47000, 9000]):
            s = np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, ln)].tobytes().decode()
            fh.write(f">c{i}\n" + "\n".join(s[k:k + 70] for k in range(0, ln, 70)) + "\n")
    # a finished find-proviruses run on the same input: rank 0 alone detects it and broadcasts the decision
    from genomad_b200 import _paths, utils
    out = tmp_path / "out"
    out.mkdir()
    o = _paths.NNOutputs("in", out)
    o.find_proviruses_dir.mkdir()
    utils.write_execution_info("find_proviruses", tmp_path / "in.fna", {}, o.find_proviruses_execution_info)
    pv = {"c0|provirus_1_9000": 9000, "c3|provirus_101_6500": 6400}
    o.find_proviruses_output.write_text("seq_name\tx\n" + "".join(f"{k}\t1\n" for k in pv))
    o.find_proviruses_nucleotide_output.write_text("".join(
        f">{k}\n" + np.frombuffer(b"ACGT", np.uint8)[rng.integers(0, 4, n)].tobytes().decode() + "\n" for k, n in pv.items()))
    o.find_proviruses_proteins_output.write_text("")
    o.find_proviruses_genes_output.write_text("")
    mp.spawn(_driver_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    first = np.load(tmp_path / "first.npy")
    # single-process run of the same stub pipeline gives the same per-contig means (sharding is invisible)
    from genomad_b200 import sequence
    pf = sequence.ParsedFasta(tmp_path / "in.fna")
    idx = pf.index()
    one = _stub_classify(_Stub(), pf, idx.offsets, gdist.DistInfo())
    assert np.array_equal(first, one)
