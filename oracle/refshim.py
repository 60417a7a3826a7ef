"""Import the UNMODIFIED reference numba path (read-only at /root/reference) in this container.

TEST INFRASTRUCTURE ONLY.  Used by ``tests/golden/generate_golden.py`` to produce the
committed golden vectors that pin ``oracle/`` (the C restatement).  Nothing in the product
package, ``bench.py`` or the ``-m gpu`` tests imports this module; ``/root/reference`` does
not exist on the GPU box.

Shims — all outside the arithmetic except the FFT:
  1. ``alphatims.utils.pjit/set_threads``: thread fan-out over ``numba.njit(nogil=True)``,
     strided partition ``iterable[tid::n]`` exactly like alphatims 1.0.9.
  2. ``alphatims.bruker.TimsTOF``, ``alpharaw.*`` readers: empty base classes (import-time only).
  3. ``matplotlib``: permissive stub (debug plotting is imported by candidate.py:18-24).
  4. ``pandas.errors.SettingWithCopyWarning`` (removed in pandas 3; fragcomp.py:10).
  5. ``alphadia.search.selection.fft`` (needs rocket-fft==0.2.5/pocketfft, not installed):
     replaced by the arithmetic definition the oracle uses — direct circular same-size
     convolution, fp64 FMA accumulation over kernel rows then columns ascending, one rounding
     to f32 — see oracle/adb_oracle.c:conv_circular.
     With ``ADB_REFSHIM_FFT=pocketfft`` in the environment the module instead follows fft.py:141-212 step by step
     (rfft2 of the zero-padded kernel, product, irfft2, the four-block roll) on scipy.fft's single-precision pocketfft —
     the library rocket-fft binds.  Only ``tests/golden/measure_fft_disagreement.py`` uses that mode: it measures how far
     the candidate tables move between the two convolutions; no golden vector comes from it.
"""

from __future__ import annotations

import importlib
import sys
import threading
import types

import numpy as np

REFERENCE_ROOT = "/root/reference"
_installed = False
_thread_count = [8]


def _stub_module(name: str, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


class _Anything:
    """Permissive object: any attribute / call returns another _Anything."""

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()

    def __iter__(self):
        return iter(())


class _AnyModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()


def _make_pjit():
    import numba

    def set_threads(n, set_global=True):
        import os

        if n <= 0:
            n = max(1, (os.cpu_count() or 1) + n)
        _thread_count[0] = int(n)
        return _thread_count[0]

    def pjit(_func=None, *, thread_count=None, include_progress_callback=True, cache=None, **kwargs):
        def deco(func):
            inner = numba.njit(nogil=True, **kwargs)(func)

            @numba.njit(nogil=True)
            def run_slice(indices, *args):
                for i in indices:
                    inner(i, *args)

            def wrapper(iterable, *args):
                idx = np.asarray(
                    np.arange(iterable.start, iterable.stop, iterable.step)
                    if isinstance(iterable, range)
                    else iterable,
                    dtype=np.int64,
                )
                n = thread_count if thread_count is not None else _thread_count[0]
                n = max(1, min(int(n), max(1, len(idx))))
                if n == 1:
                    run_slice(idx, *args)
                    return
                errors = []

                def work(t):
                    try:
                        run_slice(np.ascontiguousarray(idx[t::n]), *args)
                    except BaseException as e:  # noqa: BLE001
                        errors.append(e)

                threads = [threading.Thread(target=work, args=(t,)) for t in range(n)]
                for t in threads:
                    t.start()
                for t in threads:
                    t.join()
                if errors:
                    raise errors[0]

            wrapper.__wrapped__ = func
            return wrapper

        if _func is None:
            return deco
        return deco(_func)

    return pjit, set_threads


def _make_pocketfft_module(_MODE="pocketfft"):
    """fft.py:141-212 on scipy.fft (pocketfft, complex64): the measurement mode, see the module docstring.  ``_MODE ==
    "separable"``: the same circular convolution as two 1-D passes (what a separable GPU kernel would compute)."""
    import numba as nb
    import scipy.fft as sf
    from numba.extending import overload

    class NumbaContextOnly(Exception):
        pass

    def convolve_fourier(dense, kernel):
        raise NumbaContextOnly("numba context only")

    def _layer(x, fourier_filter, k0, k1, out):
        delta0, delta1 = -k0 // 2, -k1 // 2
        layer = sf.irfft2(sf.rfft2(x) * fourier_filter, s=x.shape).astype(np.float32, copy=False)
        out[delta0:, delta1:] = layer[:-delta0, :-delta1]
        out[:delta0, delta1:] = layer[-delta0:, :-delta1]
        out[delta0:, :delta1] = layer[:-delta0, -delta1:]
        out[:delta0, :delta1] = layer[-delta0:, -delta1:]

    def _separable_layer(x, kernel):
        """Rank-1 reading of the Gaussian: K ~ outer(u, v) with u = the centre column, v = the centre row / the centre value;
        circular convolution along the cycles with v, then along the scans with u, both in fp64, one rounding to f32."""
        k0, k1 = kernel.shape
        s0, s1 = k0 // 2, k1 // 2
        u = kernel[:, s1].astype(np.float64)
        v = kernel[s0, :].astype(np.float64) / np.float64(kernel[s0, s1])
        xd = x.astype(np.float64)
        tmp = np.zeros_like(xd)
        for b in range(k1):
            tmp += v[b] * np.roll(xd, b - s1, axis=1)
        out = np.zeros_like(xd)
        for a in range(k0):
            out += u[a] * np.roll(tmp, a - s0, axis=0)
        return out.astype(np.float32)

    def _py_conv(dense, kernel):
        dense = np.ascontiguousarray(dense, dtype=np.float32)
        kernel = np.ascontiguousarray(kernel, dtype=np.float32)
        k0, k1 = kernel.shape
        if _MODE == "separable":
            out = np.zeros_like(dense)
            flat_in = dense.reshape((-1,) + dense.shape[-2:])
            flat_out = out.reshape((-1,) + dense.shape[-2:])
            for i in range(flat_in.shape[0]):
                flat_out[i] = _separable_layer(flat_in[i], kernel)
            return out
        ff = sf.rfft2(kernel, s=dense.shape[-2:])
        assert ff.dtype == np.complex64
        out = np.zeros_like(dense)
        flat_in = dense.reshape((-1,) + dense.shape[-2:])
        flat_out = out.reshape((-1,) + dense.shape[-2:])
        for i in range(flat_in.shape[0]):
            _layer(flat_in[i], ff, k0, k1, flat_out[i])
        return out

    @overload(convolve_fourier)
    def _ov(dense, kernel):
        if dense.ndim == 2:
            def impl(dense, kernel):
                with nb.objmode(out="float32[:,::1]"):
                    out = _py_conv(dense, kernel)
                return out
            return impl
        if dense.ndim == 3:
            def impl(dense, kernel):
                with nb.objmode(out="float32[:,:,::1]"):
                    out = _py_conv(dense, kernel)
                return out
            return impl
        if dense.ndim == 4:
            def impl(dense, kernel):
                with nb.objmode(out="float32[:,:,:,::1]"):
                    out = _py_conv(dense, kernel)
                return out
            return impl
        return None

    m = types.ModuleType("alphadia.search.selection.fft")
    m.NumbaContextOnly = NumbaContextOnly
    m.convolve_fourier = convolve_fourier
    m._py_conv = _py_conv
    return m


def _make_fft_module():
    import numba as nb
    from numba.extending import overload

    class NumbaContextOnly(Exception):
        pass

    def convolve_fourier(dense, kernel):
        raise NumbaContextOnly("numba context only")

    @nb.njit(nogil=True)
    def _conv_layer(x, kernel, out):
        n0, n1 = x.shape
        k0, k1 = kernel.shape
        s0 = k0 // 2
        s1 = k1 // 2
        for i in range(n0):
            for j in range(n1):
                acc = 0.0
                for a in range(k0):
                    ii = (i + s0 - a) % n0
                    for b in range(k1):
                        jj = (j + s1 - b) % n1
                        acc = np.float64(kernel[a, b]) * np.float64(x[ii, jj]) + acc
                out[i, j] = np.float32(acc)

    @overload(convolve_fourier)
    def _ov(dense, kernel):
        if dense.ndim == 2:

            def impl(dense, kernel):
                out = np.zeros_like(dense)
                _conv_layer(dense, kernel, out)
                return out

            return impl
        if dense.ndim == 3:

            def impl(dense, kernel):
                out = np.zeros_like(dense)
                for i in range(dense.shape[0]):
                    _conv_layer(dense[i], kernel, out[i])
                return out

            return impl
        if dense.ndim == 4:

            def impl(dense, kernel):
                out = np.zeros_like(dense)
                for i in range(dense.shape[0]):
                    for j in range(dense.shape[1]):
                        _conv_layer(dense[i, j], kernel, out[i, j])
                return out

            return impl
        return None

    m = types.ModuleType("alphadia.search.selection.fft")
    m.NumbaContextOnly = NumbaContextOnly
    m.convolve_fourier = convolve_fourier
    m._conv_layer = _conv_layer
    return m


def install() -> None:
    """Install all shims and put the reference on sys.path (idempotent)."""
    global _installed
    if _installed:
        return
    import pandas.errors

    if not hasattr(pandas.errors, "SettingWithCopyWarning"):
        class SettingWithCopyWarning(Warning):
            pass

        pandas.errors.SettingWithCopyWarning = SettingWithCopyWarning

    pjit, set_threads = _make_pjit()
    at = _stub_module("alphatims")
    at.utils = _stub_module("alphatims.utils", pjit=pjit, set_threads=set_threads)

    class TimsTOF:  # noqa: D401 - import-time base class only
        pass

    at.bruker = _stub_module("alphatims.bruker", TimsTOF=TimsTOF)

    class _Base:
        def __init__(self, *a, **k):
            pass

    ar = _stub_module("alpharaw")
    ar.ms_data_base = _stub_module("alpharaw.ms_data_base", MSData_Base=_Base)
    ar.mzml = _stub_module("alpharaw.mzml", MzMLReader=type("MzMLReader", (_Base,), {}))
    ar.sciex = _stub_module("alpharaw.sciex", SciexWiffData=type("SciexWiffData", (_Base,), {}))
    ar.thermo = _stub_module("alpharaw.thermo", ThermoRawData=type("ThermoRawData", (_Base,), {}))

    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.colors",
                 "matplotlib.figure", "matplotlib.axes", "matplotlib.ticker", "seaborn"):
        if name not in sys.modules:
            sys.modules[name] = _AnyModule(name)

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import os

    mode = os.environ.get("ADB_REFSHIM_FFT", "")
    sys.modules["alphadia.search.selection.fft"] = _make_pocketfft_module(mode) if mode in ("pocketfft", "separable") else _make_fft_module()
    _installed = True


def ref(name: str):
    """Import a reference module by dotted name after installing the shims."""
    install()
    return importlib.import_module(name)


class RefDiaData4D:
    """Duck-typed DiaData for the reference classes, wrapping a RawFile4D (timsTOF layout).

    Only the fields the hot path reads carry data (cycle, dia_mz_cycle, dia_precursor_cycle, mz_values, tof_indptr,
    push_indices, intensity_values, rt_values, mobility_values, zeroth_frame, scan/frame_max_index); the remaining
    constructor arguments of TimsTOFTransposeJIT (bruker_jit.py:56-137) are inert placeholders.
    """

    def __init__(self, raw):
        self._raw = raw
        self.cycle = raw.cycle
        self.rt_values = raw.rt_values
        self.mobility_values = raw.mobility_values
        self.has_mobility = True
        self.has_ms1 = True
        self._jit = None

    def to_jitclass(self):
        if self._jit is None:
            J = ref("alphadia.search.jitclasses.bruker_jit").TimsTOFTransposeJIT
            r = self._raw
            z1f = np.zeros(1, dtype=np.float64)
            z1i = np.zeros(1, dtype=np.int64)
            self._jit = J(
                z1f, r.cycle, r.dia_mz_cycle, r.dia_precursor_cycle, int(r.frame_max_index), z1f, 65535, 0,
                r.intensity_values, 1.0, float(r.mobility_values.max()), float(r.mobility_values.min()),
                r.mobility_values, r.mz_values, z1i, 0, z1i, float(r.cycle.max()), float(r.cycle[r.cycle > 0].min()),
                np.asfortranarray(np.zeros((2, 2), dtype=np.float64)), z1i, r.rt_values, int(r.scan_max_index),
                int(len(r.mz_values)), 0, bool(r.zeroth_frame), r.push_indices, r.tof_indptr,
            )
        return self._jit


class RefDiaData:
    """Duck-typed DiaData for the reference classes, wrapping a RawFile3D."""

    def __init__(self, raw):
        self._raw = raw
        self.cycle = raw.cycle
        self.rt_values = raw.rt_values
        self.mobility_values = raw.mobility_values
        self.has_mobility = bool(raw.has_mobility)
        self.has_ms1 = True
        self._jit = None

    def to_jitclass(self):
        if self._jit is None:
            AlphaRawJIT = ref("alphadia.search.jitclasses.alpharaw_jit").AlphaRawJIT
            r = self._raw
            self._jit = AlphaRawJIT(
                r.cycle, r.rt_values, r.mobility_values, bool(r.zeroth_frame),
                np.float32(r.max_mz_value), np.float32(r.min_mz_value),
                np.float32(r.max_mz_value), np.float32(r.min_mz_value),
                int(r.precursor_cycle_max_index),
                r.peak_start_idx_list, r.peak_stop_idx_list, r.mz_values, r.intensity_values,
                int(r.scan_max_index), int(r.frame_max_index),
            )
        return self._jit
