"""TEST / BASELINE INFRASTRUCTURE — the CPU arm "same table, same algorithm" of SURVEY.md section 8(d).

The per-query device source (binder_b200/csrc/resolve_device.cuh) compiled for the host through
tests/native/cuda_shim.h and driven tile by tile by tests/native/emu_resolve.cpp — the word-wise parse, the
multiply-fold hashes, the one-sector cuckoo probe, the ready-RR copy jobs — over the SAME zone image (table + arena)
the GPU probes, on all host threads, -O3 (x86-64-v2: the library is built where the repo is and runs on the GPU box, whose
CPU may be another model — no -march=native).  It isolates the processor: same layout, same algorithm.
Only tests/ and bench.py's cpu_baseline leg may import this; the product (binder_b200/) never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
_SO = os.path.join(_HERE, 'libsametable.so')
_SRCS = [os.path.join(_ROOT, 'tests', 'native', 'emu_resolve.cpp'), os.path.join(_ROOT, 'binder_b200', 'csrc', 'zone_build.cpp')]
_DEPS = _SRCS + [os.path.join(_ROOT, 'binder_b200', 'csrc', f) for f in ('resolve_device.cuh', 'zone_image.h')] + \
    [os.path.join(_ROOT, 'tests', 'native', 'cuda_shim.h')]


def build(force=False):
    if force or not os.path.exists(_SO) or any(os.path.getmtime(_SO) < os.path.getmtime(d) for d in _DEPS):
        subprocess.check_call(['g++', '-std=c++17', '-O3', '-march=x86-64-v2', '-fPIC', '-shared', '-pthread', '-ftls-model=initial-exec',
                               '-Wno-unknown-pragmas', '-I', os.path.join(_ROOT, 'include'), '-I', os.path.join(_ROOT, 'binder_b200', 'csrc'),
                               '-o', _SO] + _SRCS)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        L.bb_zone_build.restype = ctypes.c_void_p
        L.bb_zone_build.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.POINTER(ctypes.c_int)]
        L.bb_zone_free.argtypes = [ctypes.c_void_p]
        L.bb_zone_image.restype = ctypes.c_void_p
        L.bb_zone_image.argtypes = [ctypes.c_void_p]
        L.bb_emu_timed_resolve.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32,
                                           ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                                           ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64)]
        _lib = L
    return _lib


class SameTable(object):
    """A zone image built by the product's own builder source (compiled into this library) and the timed resolve."""

    def __init__(self, dns_domain, snapshot_jsonl, recursion=False):
        if isinstance(snapshot_jsonl, str):
            snapshot_jsonl = snapshot_jsonl.encode('utf-8')
        err = ctypes.c_int(0)
        self.dom, self.recursion = dns_domain.encode(), int(bool(recursion))
        self._z = lib().bb_zone_build(snapshot_jsonl, len(snapshot_jsonl), self.dom, ctypes.byref(err))
        if not self._z:
            raise RuntimeError('same_table: zone build failed (%d)' % err.value)
        self._img = lib().bb_zone_image(self._z)

    def timed_resolve(self, data, off, seed=0, nthreads=1, repeat=1, resp_cap=1232):
        """-> (best seconds per pass, response bytes, misses) of one batch on `nthreads` host threads."""
        data = np.ascontiguousarray(data, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint32)
        n = len(off) - 1
        secs, tb, tm = ctypes.c_double(0), ctypes.c_uint64(0), ctypes.c_uint64(0)
        rc = lib().bb_emu_timed_resolve(self._img, self.dom, self.recursion, data.ctypes.data, off.ctypes.data, n, seed, nthreads, repeat,
                                        resp_cap, ctypes.byref(secs), ctypes.byref(tb), ctypes.byref(tm))
        if rc != 0:
            raise RuntimeError('same_table: resolve failed (%d)' % rc)
        return secs.value, tb.value, tm.value

    def close(self):
        if getattr(self, '_z', None):
            lib().bb_zone_free(self._z)
            self._z = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
