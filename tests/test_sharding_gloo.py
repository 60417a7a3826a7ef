"""world_size-2 run of the N>1 host path on CPU (gloo): each rank scores its shard of the library with the
oracle standing in for the device, then the single all-gather of the packed score table; every rank must end
up with the table of the unsharded run."""

import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

INT_COLS = ["precursor_idx", "rank", "scan_center", "scan_start", "scan_stop", "frame_center", "frame_start", "frame_stop"]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _score_library(oracle, H, raw, lib):
    from alphadia_b200.sharding import pack_score_table

    p_cfg = H.selection_config(30.0).to_struct()
    cont = oracle.select_candidates(raw, lib, p_cfg, H.default_kernel(raw), n_threads=2)
    m = cont["score"] > 0
    cand = {c: cont[c][m] for c in INT_COLS}
    cin, keep = H.candidates_in_from_arrays(lib, cand)
    out = oracle.score_candidates(raw, lib, H.scoring_config().to_struct(), cin, n_threads=2)
    return pack_score_table(out["features"], keep["precursor_idx"], keep["rank"], out["valid"])


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    import oracle
    from alphadia_b200.library import assemble_library_arrays
    import torch

    from alphadia_b200.sharding import ScoreTableGather, allgather_score_table, shard_library
    from tests import helpers as H

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    raw, pdf, fdf, lib, p = H.workload("config1")
    sp, sf = shard_library(pdf, fdf, rank, world)
    sl = assemble_library_arrays(sp, sf, "rt_library", "mobility_library", "mz_library", "mz_library")
    local = _score_library(oracle, H, raw, sl)
    full = allgather_score_table(local)
    # the preallocated form of the same collective (what bench.py uses with NCCL): capacity = the largest shard's row bound
    g = ScoreTableGather(len(pdf) * 3, torch.device("cpu"))
    g.local[: local.shape[0]] = torch.from_numpy(local.view(np.int32))
    g.n_local.fill_(local.shape[0])
    g.allgather()
    assert np.array_equal(g.compact().numpy().view(np.uint32), full)
    q.put((rank, local.shape[0], full))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_allgather_matches_unsharded(oracle_lib):
    import torch.multiprocessing as mp

    from tests import helpers as H

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    results = [q.get(timeout=500) for _ in procs]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    raw, pdf, fdf, lib, p = H.workload("config1")
    expected = _score_library(oracle_lib, H, raw, lib)
    results.sort(key=lambda r: r[0])
    assert results[0][1] + results[1][1] == expected.shape[0]
    for _, _, full in results:
        assert full.shape == expected.shape
        assert np.array_equal(full, expected)  # bit-identical, NaNs included (uint32 words)
