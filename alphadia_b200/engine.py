"""One hot-path "step" = candidate selection + candidate scoring of one library batch against one raw file.

Two ways to run it, both through the C ABI:

* ``resident_step``: raw file and library already in HBM, results stay in HBM
  (``adb_select_candidates_resident`` -> device compaction -> ``adb_score_candidates_resident``).
* ``host_step``: what the reference-facing operators do — library upload, selection with the candidate
  container copied back, the ``score > 0`` filter on the host, candidate table upload, scoring, score and
  fragment tables copied back.  Host buffers may be pinned by the caller.
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from alphadia_b200 import _abi, _lib

CAND_COLS = ("scan_start", "scan_stop", "scan_center", "frame_start", "frame_stop", "frame_center")


class HotPath:
    def __init__(self, raw, lib_arrays: dict, sel_cfg, score_cfg, kernel: np.ndarray, device: int | None = None,
                 quad_sigma=(0.2, 0.2), quad_delta_mu=(0.0, 0.0)):
        self.lib_arrays = lib_arrays
        self.sel_struct = sel_cfg.to_struct() if hasattr(sel_cfg, "to_struct") else sel_cfg
        self.score_struct = (score_cfg.to_struct(quad_sigma=quad_sigma, quad_delta_mu=quad_delta_mu)
                             if hasattr(score_cfg, "to_struct") else score_cfg)
        self.kernel = np.ascontiguousarray(kernel, dtype=np.float32)
        self.dev_raw = _lib.DeviceRawFile(raw, device=device)
        self.dev_lib = _lib.DeviceLibrary(lib_arrays, device=self.dev_raw.device)
        self.n_precursors = self.dev_lib.n_precursors
        self.top_k = int(self.score_struct.top_k_fragments)
        self.max_lib_fragments = int(np.max(lib_arrays["frag_stop_idx"].astype(np.int64) - lib_arrays["frag_start_idx"].astype(np.int64))) if len(lib_arrays["frag_start_idx"]) else 1
        self.n_candidates = 0
        self._host_bufs = None

    # ---- resident ---------------------------------------------------------------------------------
    def resident_step(self) -> dict:
        n = _lib.select_candidates_resident(self.dev_raw, self.dev_lib, self.sel_struct, self.kernel)
        t_sel = self.dev_raw.last_timing()
        _lib.score_candidates_resident(self.dev_raw, self.dev_lib, self.score_struct)
        t_sc = self.dev_raw.last_timing()
        self.n_candidates = n
        return dict(n_candidates=n, select_ms=t_sel["kernel_ms"], score_ms=t_sc["kernel_ms"],
                    select_kernel_ms=t_sel["main_kernel_ms"], score_kernel_ms=t_sc["main_kernel_ms"])

    def fetch(self) -> dict:
        return _lib.fetch_scores(self.dev_raw, self.n_candidates, self.top_k)

    def score_table_pointers(self):
        f, v, r, k = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        n = C.c_int64()
        _lib.check(_lib.load().adb_resident_score_table(self.dev_raw.handle, C.byref(f), C.byref(v), C.byref(r),
                                                        C.byref(k), C.byref(n)), "adb_resident_score_table")
        return dict(features=f.value, valid=v.value, lib_row=r.value, rank=k.value, n=int(n.value))

    # ---- host buffers (the operator path) -----------------------------------------------------------
    def host_step(self, alloc=np.zeros, ragged: bool = True, fused: bool = False) -> dict:
        """What the reference-facing operators do, with caller-provided (pinned) host memory:
        library batch H2D -> selection -> compacted candidate table D2H (CandidateSelection's DataFrame columns)
        -> candidate table H2D -> scoring -> score + fragment tables D2H (row blocks overlap the scoring kernel).
        ``ragged`` (default): the result is the device-compacted form of ``adb_score_candidates_ragged`` (valid feature
        rows + kept fragment slots); otherwise the dense ``[n, top_k]`` tables of ``adb_score_candidates``.
        ``fused``: selection and scoring in ONE C-ABI call (``adb_select_score_candidates_ragged``, the workflow's
        select_candidates -> score_and_quantify_candidates pair): the candidate table is copied to the host while the first
        scoring block runs and is not uploaded again.
        ``alloc(shape, dtype) -> ndarray`` lets the caller provide pinned host memory."""
        import time

        lib = _lib.load()
        t_phase = {}
        t_enter = t_last = time.perf_counter()

        def lap(name):
            nonlocal t_last
            now = time.perf_counter()
            t_phase[name] = t_phase.get(name, 0.0) + (now - t_last) * 1e3
            t_last = now

        n_rows = int(self.n_precursors * self.sel_struct.candidate_count)
        if self._host_bufs is None or self._host_bufs["n_rows"] != n_rows:
            lib_host = {}
            by_address = {}
            for k, v in self.lib_arrays.items():  # the library batch as the caller would hold it: (pinned) host arrays
                key = (v.__array_interface__["data"][0], v.shape, v.dtype.str)
                if key in by_address:  # two fields that are views of one column (library m/z == search m/z) stay one buffer
                    lib_host[k] = lib_host[by_address[key]]
                    continue
                buf = alloc(v.shape, v.dtype)
                np.copyto(buf, v)
                lib_host[k] = buf
                by_address[key] = k
            self._host_bufs = dict(n_rows=n_rows, table=_abi.alloc_candidate_table(n_rows, alloc), lib=lib_host,
                                   scores=None, scores_n=0)
        table = self._host_bufs["table"]
        dev_lib = _lib.DeviceLibrary(self._host_bufs["lib"], device=self.dev_raw.device)
        h2d = sum(int(v.nbytes) for v in {id(v): v for v in self._host_bufs["lib"].values()}.values())
        lap("library_upload")
        if fused:
            try:
                per = max(1, min(self.top_k, self.max_lib_fragments))
                if self._host_bufs["scores"] is None or self._host_bufs["scores_n"] < n_rows or not self._host_bufs.get("ragged"):
                    _, sc = _abi.alloc_scores_ragged(n_rows, n_rows * per, alloc)
                    self._host_bufs["scores"], self._host_bufs["scores_n"], self._host_bufs["ragged"] = sc, n_rows, True
                sc = self._host_bufs["scores"]
                lap("host_glue")
                res = _lib.select_score_candidates_ragged(self.dev_raw, dev_lib, self.sel_struct, self.kernel, self.score_struct,
                                                          table, sc)
                lap("select_score_call")
            finally:
                dev_lib.close()
            n = res["n_candidates"]
            self.n_candidates = n
            lap("library_free")
            t_phase["total"] = (time.perf_counter() - t_enter) * 1e3
            d2h = n * (7 * 8 + 1 + 4 + 4) + res["n_rows"] * (_abi.NUM_FEATURES * 4 + 16) + 8 + res["n_fragments"] * (7 * 4 + 5)
            return dict(n_candidates=n, h2d_bytes=h2d, d2h_bytes=d2h, valid=res["n_rows"], n_fragments=res["n_fragments"],
                        phases_ms=t_phase)
        try:
            n = _lib.select_candidates_resident(self.dev_raw, dev_lib, self.sel_struct, self.kernel)
            t_phase["select_call_device"] = dict(self.dev_raw.last_timing())
            lap("select_call")
            _lib.fetch_candidate_table(self.dev_raw, n, table)
            d2h = n * (7 * 8 + 1 + 4 + 4)
            lap("candidate_table_d2h")
            cin = _abi.candidates_in_from_table(table, n)
            h2d += n * (7 * 8 + 1)
            if ragged:
                # what collect_candidates / collect_fragments keep: feature rows of the valid candidates + their fragment
                # slots with mz_library > 0, compacted on the device (adb_score_candidates_ragged)
                per = max(1, min(self.top_k, self.max_lib_fragments))
                if self._host_bufs["scores"] is None or self._host_bufs["scores_n"] < n or not self._host_bufs.get("ragged"):
                    cap = int(n * 1.05) + 16
                    _, sc = _abi.alloc_scores_ragged(cap, cap * per, alloc)
                    self._host_bufs["scores"], self._host_bufs["scores_n"], self._host_bufs["ragged"] = sc, cap, True
                sc = self._host_bufs["scores"]
                lap("host_glue")
                res = _lib.score_candidates_ragged(self.dev_raw, dev_lib, self.score_struct, cin, bufs=sc)
                n_valid, n_frag = res["n_rows"], res["n_fragments"]
                d2h += n_valid * (_abi.NUM_FEATURES * 4 + 16) + 8 + n_frag * (7 * 4 + 5)
            else:
                if self._host_bufs["scores"] is None or self._host_bufs["scores_n"] < n or self._host_bufs.get("ragged"):
                    cap = int(n * 1.05) + 16
                    sc = dict(features=alloc((cap, _abi.NUM_FEATURES), np.float32), valid=alloc(cap, np.uint8))
                    for k in _abi.FRAG_F32:
                        sc[k] = alloc((cap, self.top_k), np.float32)
                    for k in _abi.FRAG_U8:
                        sc[k] = alloc((cap, self.top_k), np.uint8)
                    self._host_bufs["scores"], self._host_bufs["scores_n"], self._host_bufs["ragged"] = sc, cap, False
                sc = self._host_bufs["scores"]
                so = _abi.ScoresOut()
                for k, v in sc.items():
                    setattr(so, k, _abi.ptr(v))
                lap("host_glue")
                _lib.check(lib.adb_score_candidates(self.dev_raw.handle, dev_lib.handle, C.byref(self.score_struct),
                                                    C.byref(cin), C.byref(so)), "adb_score_candidates")
                d2h += n * (_abi.NUM_FEATURES * 4 + 1 + self.top_k * (7 * 4 + 5))
                n_valid, n_frag = int(np.count_nonzero(sc["valid"][:n])), None
            lap("score_call")
            t_phase["score_call_device"] = dict(self.dev_raw.last_timing())
        finally:
            dev_lib.close()
        self.n_candidates = n
        lap("library_free")
        t_phase["total"] = (time.perf_counter() - t_enter) * 1e3
        return dict(n_candidates=n, h2d_bytes=h2d, d2h_bytes=d2h, valid=n_valid, n_fragments=n_frag, phases_ms=t_phase)

    def close(self):
        self.dev_lib.close()
        self.dev_raw.close()
