"""Multi-GPU sharding of the hot path: one process per GPU, no data-path collective, ONE all-gather of the
score table at the end.

Precursors (selection) and candidates (scoring) are independent work items that write disjoint output
rows — the reference already runs them under an arbitrary strided thread partition
(alphatims ``pjit``: ``iterable[thread_id::n]``; selection.py:78,656-666, scoring.py:114,633-643) — and
raw files are independent too (search_step.py:455).  So a rank either owns whole raw files (one file per
GPU) or a contiguous block of the library against a replicated raw file; nothing is exchanged until the
fixed-width score table [n, 46 f32 | precursor_idx u32 | rank u8, valid u8] is gathered once.

``torch.distributed`` is plumbing only (NCCL on GPUs, gloo in the CPU tests).
"""

from __future__ import annotations

import numpy as np
import pandas as pd

ROW_WORDS = 48  # 46 feature words + precursor_idx word + (rank | valid << 8) word


def shard_bounds(n_items: int, world_size: int) -> np.ndarray:
    """Contiguous, balanced block boundaries: rank r owns [b[r], b[r+1])."""
    base, rem = divmod(int(n_items), int(world_size))
    sizes = np.full(world_size, base, dtype=np.int64)
    sizes[:rem] += 1
    return np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)


def shard_library(precursor_df: pd.DataFrame, fragment_df: pd.DataFrame, rank: int, world_size: int):
    """Block-partition a flat library by precursor (sorted by precursor_idx).  The fragment table is sliced
    to the block and ``flat_frag_*_idx`` are rebased, so a shard is a self-contained flat library."""
    pdf = precursor_df.sort_values("precursor_idx").reset_index(drop=True)
    b = shard_bounds(len(pdf), world_size)
    sub = pdf.iloc[b[rank]: b[rank + 1]].copy()
    if len(sub) == 0:
        return sub, fragment_df.iloc[0:0].copy()
    starts = sub["flat_frag_start_idx"].values.astype(np.int64)
    stops = sub["flat_frag_stop_idx"].values.astype(np.int64)
    contiguous = np.all(starts[1:] == stops[:-1])
    if contiguous:
        lo, hi = int(starts[0]), int(stops[-1])
        frag = fragment_df.iloc[lo:hi].reset_index(drop=True)
        sub["flat_frag_start_idx"] = (starts - lo).astype(np.uint32)
        sub["flat_frag_stop_idx"] = (stops - lo).astype(np.uint32)
    else:  # general case: gather the referenced ranges
        lens = stops - starts
        idx = np.concatenate([np.arange(s, e) for s, e in zip(starts, stops)]) if len(starts) else np.zeros(0, np.int64)
        frag = fragment_df.iloc[idx].reset_index(drop=True)
        new_stop = np.cumsum(lens)
        sub["flat_frag_start_idx"] = (new_stop - lens).astype(np.uint32)
        sub["flat_frag_stop_idx"] = new_stop.astype(np.uint32)
    return sub.reset_index(drop=True), frag


def pack_score_table(features: np.ndarray, precursor_idx: np.ndarray, rank: np.ndarray, valid: np.ndarray) -> np.ndarray:
    """[n, 48] uint32 words (features bit-cast), the unit that is all-gathered."""
    n = len(precursor_idx)
    out = np.zeros((n, ROW_WORDS), dtype=np.uint32)
    out[:, :46] = np.ascontiguousarray(features, dtype=np.float32).view(np.uint32).reshape(n, 46)
    out[:, 46] = precursor_idx.astype(np.uint32)
    out[:, 47] = rank.astype(np.uint32) | (valid.astype(np.uint32) << 8)
    return out


def unpack_score_table(words: np.ndarray) -> dict:
    words = np.ascontiguousarray(words, dtype=np.uint32)
    return dict(
        features=words[:, :46].copy().view(np.float32),
        precursor_idx=words[:, 46].copy(),
        rank=(words[:, 47] & 0xFF).astype(np.uint8),
        valid=((words[:, 47] >> 8) & 0xFF).astype(np.uint8),
    )


def allgather_score_table(local_words, group=None):
    """All ranks end up with the concatenation (in rank order) of every rank's packed table.

    ``local_words``: [n_local, 48] uint32/int32 — a numpy array (gloo / CPU) or a CUDA ``torch.Tensor``
    (NCCL; stays on the device).  Exactly one collective moves table data (a tiny size exchange precedes it).
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    is_numpy = isinstance(local_words, np.ndarray)
    t = torch.from_numpy(np.ascontiguousarray(local_words).view(np.int32)) if is_numpy else local_words.view(torch.int32)
    dev = t.device
    n_local = torch.tensor([t.shape[0]], dtype=torch.int64, device=dev)
    sizes = torch.zeros(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(sizes, n_local, group=group)
    sizes_h = sizes.cpu().numpy()
    n_max = int(sizes_h.max()) if world else 0
    padded = torch.zeros((n_max, ROW_WORDS), dtype=torch.int32, device=dev)
    padded[: t.shape[0]] = t
    gathered = torch.empty((world * n_max, ROW_WORDS), dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(gathered, padded, group=group)  # THE collective
    gathered = gathered.view(world, n_max, ROW_WORDS)
    parts = [gathered[r, : int(sizes_h[r])] for r in range(world)]
    out = torch.cat(parts, dim=0) if parts else gathered.view(0, ROW_WORDS)
    return out.cpu().numpy().view(np.uint32) if is_numpy else out


class ScoreTableGather:
    """The one collective of the sharded path with preallocated device buffers (NCCL): no per-step allocation, no host
    synchronisation and no concatenation copy.  ``capacity`` is a static upper bound of a rank's rows (precursors x
    candidate_count of the largest shard), so the table collective does not have to wait for the row counts; the counts
    travel in a second, 8-byte collective and stay on the device until someone asks for the compact table."""

    def __init__(self, capacity: int, device, group=None):
        import torch
        import torch.distributed as dist

        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.capacity = int(capacity)
        self.local = torch.zeros((self.capacity, ROW_WORDS), dtype=torch.int32, device=device)
        self.gathered = torch.empty((self.world, self.capacity, ROW_WORDS), dtype=torch.int32, device=device)
        self.n_local = torch.zeros(1, dtype=torch.int64, device=device)
        self.sizes = torch.zeros(self.world, dtype=torch.int64, device=device)

    def pack_resident(self, hot_path, lib_precursor_idx_dev) -> int:
        """Packs the resident score table of ``hot_path`` into ``self.local[:n]`` (device-side, torch ops)."""
        import torch

        p = hot_path.score_table_pointers()
        n = p["n"]
        if n > self.capacity:
            raise ValueError(f"score table has {n} rows, gather capacity is {self.capacity}")
        self.n_local.fill_(n)
        if n == 0:
            return 0
        dev = self.local.device
        feats = torch.as_tensor(_CudaArrayView(p["features"], (n, 46), "<f4"), device=dev)
        valid = torch.as_tensor(_CudaArrayView(p["valid"], (n,), "|u1"), device=dev)
        rank = torch.as_tensor(_CudaArrayView(p["rank"], (n,), "|u1"), device=dev)
        lib_row = torch.as_tensor(_CudaArrayView(p["lib_row"], (n,), "<i8"), device=dev)
        w = self.local[:n]
        w[:, :46] = feats.view(torch.int32)
        w[:, 46] = lib_precursor_idx_dev[lib_row].to(torch.int32)
        w[:, 47] = rank.to(torch.int32) | (valid.to(torch.int32) << 8)
        return n

    def allgather(self):
        """THE collective (+ the 8-byte row-count exchange).  Returns ``(gathered [world, capacity, 48], sizes [world])``, both on
        the device; rows beyond ``sizes[r]`` of block r are padding."""
        import torch.distributed as dist

        if self.world == 1:
            self.gathered[0].copy_(self.local)
            self.sizes.copy_(self.n_local)
        else:
            dist.all_gather_into_tensor(self.sizes, self.n_local, group=self.group)
            dist.all_gather_into_tensor(self.gathered.view(self.world * self.capacity, ROW_WORDS), self.local, group=self.group)
        return self.gathered, self.sizes

    def compact(self):
        """The gathered table without padding, in rank order (synchronises: the row counts come to the host)."""
        import torch

        sizes = self.sizes.cpu().tolist()
        return torch.cat([self.gathered[r, : int(sizes[r])] for r in range(self.world)], dim=0)


class _CudaArrayView:
    """Expose a raw device pointer to torch through __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr: int, shape: tuple, typestr: str):
        self.__cuda_array_interface__ = dict(shape=shape, typestr=typestr, data=(int(ptr), False), version=3, strides=None)


def device_words_from_resident(hot_path, lib_precursor_idx_dev=None):
    """Pack the resident score table of a ``HotPath`` into [n, 48] int32 words ON THE DEVICE (torch ops)."""
    import torch

    p = hot_path.score_table_pointers()
    n = p["n"]
    dev = torch.device("cuda", hot_path.dev_raw.device)
    if n == 0:
        return torch.zeros((0, ROW_WORDS), dtype=torch.int32, device=dev)
    feats = torch.as_tensor(_CudaArrayView(p["features"], (n, 46), "<f4"), device=dev)
    valid = torch.as_tensor(_CudaArrayView(p["valid"], (n,), "|u1"), device=dev)
    rank = torch.as_tensor(_CudaArrayView(p["rank"], (n,), "|u1"), device=dev)
    lib_row = torch.as_tensor(_CudaArrayView(p["lib_row"], (n,), "<i8"), device=dev)
    if lib_precursor_idx_dev is None:
        lib_precursor_idx_dev = torch.from_numpy(hot_path.lib_arrays["precursor_idx"].astype(np.int64)).to(dev)
    words = torch.empty((n, ROW_WORDS), dtype=torch.int32, device=dev)
    words[:, :46] = feats.view(torch.int32)
    words[:, 46] = lib_precursor_idx_dev[lib_row].to(torch.int32)
    words[:, 47] = rank.to(torch.int32) | (valid.to(torch.int32) << 8)
    return words
