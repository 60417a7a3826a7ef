"""N > 1 on real devices: two ranks score their shard of the library ON THE GPU (C ABI), the packed score tables are
all-gathered (NCCL when the box has two GPUs, otherwise both ranks share cuda:0 and the collective runs over gloo), and every
rank's gathered table must equal the table of the unsharded single-GPU run bit for bit."""

import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INT_COLS = ["precursor_idx", "rank", "scan_center", "scan_start", "scan_stop", "frame_center", "frame_start", "frame_stop"]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _device_table(engine, H, raw, lib, device):
    """Selection + scoring of `lib` against `raw` on `device` through the C ABI -> packed [n, 48] words."""
    from alphadia_b200.sharding import pack_score_table

    draw, dlib = engine.DeviceRawFile(raw, device=device), engine.DeviceLibrary(lib, device=device)
    cont = engine.select_candidates(draw, dlib, H.selection_config(30.0).to_struct(), H.default_kernel(raw))
    m = cont["score"] > 0
    cin, keep = H.candidates_in_from_arrays(lib, {c: cont[c][m] for c in INT_COLS})
    out = engine.score_candidates(draw, dlib, H.scoring_config().to_struct(), cin)
    dlib.close(); draw.close()
    return pack_score_table(out["features"], keep["precursor_idx"], keep["rank"], out["valid"])


def _worker(rank, world, port, n_dev, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist

    from alphadia_b200 import _lib as engine
    from alphadia_b200.library import assemble_library_arrays
    from alphadia_b200.sharding import ScoreTableGather, allgather_score_table, shard_library
    from tests import helpers as H

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    nccl = n_dev >= world
    device = rank if nccl else 0
    torch.cuda.set_device(device)
    if nccl:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", device))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    raw, pdf, fdf, lib, p = H.workload("parity_small")
    sp, sf = shard_library(pdf, fdf, rank, world)
    sl = assemble_library_arrays(sp, sf, "rt_library", "mobility_library", "mz_library", "mz_library")
    local = _device_table(engine, H, raw, sl, device)
    if nccl:
        t = torch.from_numpy(local.view(np.int32)).cuda(device)
        full = allgather_score_table(t).cpu().numpy().view(np.uint32)
        g = ScoreTableGather(len(pdf) * 3, torch.device("cuda", device))
        g.local[: local.shape[0]] = t
        g.n_local.fill_(local.shape[0])
        g.allgather()
        assert np.array_equal(g.compact().cpu().numpy().view(np.uint32), full)
    else:
        full = allgather_score_table(local)
    q.put((rank, local.shape[0], full, "nccl" if nccl else "gloo"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(900)
def test_two_rank_device_tables_gather_to_the_unsharded_table():
    import torch.multiprocessing as mp

    from alphadia_b200 import _lib as engine
    from tests import helpers as H

    n_dev = engine.require_device()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_dev, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    results = [q.get(timeout=800) for _ in procs]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    raw, pdf, fdf, lib, p = H.workload("parity_small")
    expected = _device_table(engine, H, raw, lib, 0)
    results.sort(key=lambda r: r[0])
    assert results[0][1] + results[1][1] == expected.shape[0] > 1000
    for _, _, full, backend in results:
        assert full.shape == expected.shape
        assert np.array_equal(full, expected), backend  # bit-identical, NaNs included (uint32 words)
