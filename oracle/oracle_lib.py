"""TEST INFRASTRUCTURE — ctypes binding for oracle/liboracle.so (see oracle.cpp header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this.  The product (binder_b200/) never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, 'liboracle.so')


def build(force=False):
    src = os.path.join(_HERE, 'oracle.cpp')
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(['make', '-C', _HERE, '-B', 'liboracle.so'],
                              stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        L.orc_create.restype = ctypes.c_void_p
        L.orc_create.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int]
        L.orc_destroy.argtypes = [ctypes.c_void_p]
        L.orc_load_snapshot.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t]
        L.orc_set_recursion_filter.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int]
        L.orc_apply_delta.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t]
        L.orc_node_count.restype = ctypes.c_long
        L.orc_node_count.argtypes = [ctypes.c_void_p]
        L.orc_resolve_batch.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint64,
            ctypes.c_uint32, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.orc_resolve_batch_ex.argtypes = L.orc_resolve_batch.argtypes + [ctypes.c_uint32]
        _lib = L
    return _lib


class Oracle(object):
    def __init__(self, dns_domain, datacenter='', recursion=False, snapshot=None):
        self._h = lib().orc_create(dns_domain.encode(), datacenter.encode(), int(recursion))
        if snapshot is not None:
            self.load_snapshot(snapshot)

    def load_snapshot(self, jsonl):
        if isinstance(jsonl, str):
            jsonl = jsonl.encode('utf-8')
        rc = lib().orc_load_snapshot(self._h, jsonl, len(jsonl))
        if rc != 0:
            raise ValueError('oracle: bad snapshot (%d)' % rc)

    def set_recursion_filter(self, region_domain, dcs=(), ptr=False):
        """Misses lib/recursion.js:329-344 would refuse without asking anyone are answered REFUSED."""
        if region_domain is None:
            lib().orc_set_recursion_filter(self._h, None, None, 0, 0)
            return
        arr = (ctypes.c_char_p * max(len(dcs), 1))(*[d.encode('latin-1') for d in dcs])
        lib().orc_set_recursion_filter(self._h, region_domain.encode('latin-1'), arr, len(dcs), int(ptr))

    def apply_delta(self, jsonl):
        """Watch events on the loaded cache (JSON lines; see ZKCache::apply in oracle.cpp)."""
        if isinstance(jsonl, str):
            jsonl = jsonl.encode('utf-8')
        rc = lib().orc_apply_delta(self._h, jsonl, len(jsonl))
        if rc != 0:
            raise ValueError('oracle: bad delta (%d)' % rc)

    def node_count(self):
        return lib().orc_node_count(self._h)

    def resolve_batch(self, data, off, seed=0, qidx_base=0, nthreads=1, out_cap=None, tcp=False):
        """data: uint8 array of packed packets, off: uint32[n+1].
        Returns (out uint8[total], out_off uint32[n+1], out_len uint16[n], status uint8[n], miss uint32[m])."""
        data = np.ascontiguousarray(data, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint32)
        n = len(off) - 1
        if out_cap is None:
            out_cap = max(1, min(n * (65536 if tcp else 1232), 0xFFFFFFF0))
        out = np.empty(out_cap, dtype=np.uint8)
        out_off = np.zeros(n + 1, dtype=np.uint32)
        status = np.zeros(max(n, 1), dtype=np.uint8)
        miss = np.zeros(max(n, 1), dtype=np.uint32)
        n_miss = ctypes.c_uint32(0)
        rc = lib().orc_resolve_batch_ex(self._h, data.ctypes.data, off.ctypes.data, n, seed, qidx_base,
                                        out.ctypes.data, out_cap, out_off.ctypes.data,
                                        status.ctypes.data, miss.ctypes.data, ctypes.byref(n_miss),
                                        nthreads, 1 if tcp else 0)
        if rc != 0:
            raise RuntimeError('oracle resolve failed (%d)' % rc)
        lens = np.diff(out_off.astype(np.int64)).astype(np.uint16)
        return out[:out_off[n]].copy(), out_off, lens, status[:n], miss[:n_miss.value].copy()

    def timed_resolve(self, data, off, seed=0, nthreads=1, repeat=1):
        """Seconds per call of orc_resolve_batch alone (buffers allocated and touched beforehand)."""
        import time
        data = np.ascontiguousarray(data, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.uint32)
        n = len(off) - 1
        out_cap = max(1, min(n * 600, 0xFFFFFFF0))
        out = np.zeros(out_cap, dtype=np.uint8)
        out_off = np.zeros(n + 1, dtype=np.uint32)
        status = np.zeros(max(n, 1), dtype=np.uint8)
        miss = np.zeros(max(n, 1), dtype=np.uint32)
        n_miss = ctypes.c_uint32(0)
        best = []
        for _ in range(repeat):
            t0 = time.perf_counter()
            rc = lib().orc_resolve_batch(self._h, data.ctypes.data, off.ctypes.data, n, seed, 0,
                                         out.ctypes.data, out_cap, out_off.ctypes.data, status.ctypes.data,
                                         miss.ctypes.data, ctypes.byref(n_miss), nthreads)
            best.append(time.perf_counter() - t0)
            if rc != 0:
                raise RuntimeError('oracle resolve failed (%d)' % rc)
        return best

    def resolve_one(self, pkt, seed=0, qidx=0):
        data = np.frombuffer(pkt, dtype=np.uint8)
        off = np.array([0, len(pkt)], dtype=np.uint32)
        out, out_off, lens, status, miss = self.resolve_batch(data, off, seed, qidx)
        return bytes(out), int(status[0])

    def close(self):
        if self._h:
            lib().orc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
