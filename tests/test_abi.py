"""CPU-only checks of the C-ABI library: it loads, exports every symbol include/binder_b200.h
declares, builds zones on the host, and refuses to run without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from binder_b200 import _lib, build, synth
from binder_b200.engine import Zone

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def setup_module(_m):
    build.build()


def test_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, 'include', 'binder_b200.h')).read()
    declared = set(re.findall(r'\b(bb_[a-z_]+)\s*\(', hdr))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    L = ctypes.CDLL(_lib.SO_PATH)
    for name in declared:
        assert hasattr(L, name), name
    assert _lib.lib().bb_abi_version() == 1


def test_zone_build_on_host():
    z = synth.gen_zone(5000, service_frac=0.2)
    st = Zone(z.jsonl, z.dns_domain).stat()
    assert st['nodes'] == z.n_records + 1 - 0          # + root (its own line is part of n_records)
    assert st['forward_keys'] == st['nodes']
    assert st['slots'] >= 2 * (st['forward_keys'] + st['reverse_keys'])
    assert st['slots'] & (st['slots'] - 1) == 0


def test_zone_build_rejects_bad_input():
    for snap, dom in ((b'{"path": 5}\n', 'foo.com'), (b'not json\n', 'foo.com'),
                      (b'{"path":"/com/foo/a","data":null}\n{"path":"/com/foo/a","data":null}\n', 'foo.com'),
                      (b'', 'Foo.com'), (b'', 'a..b'), (b'', '')):
        with pytest.raises(_lib.BinderError):
            Zone(snap, dom)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    from binder_b200.engine import Engine
    with pytest.raises(_lib.BinderError) as ei:
        Engine('foo.com')
    assert ei.value.code == -6


def test_zone_build_survives_hash_collisions():
    """1.5M znodes with services: ~3M keys contain several hundred pairs with identical 32-bit
    hashes.  With the second cuckoo slot derived from the same hash, two such pairs plus one more
    key formed an unsatisfiable cycle (zone build failed after growing 4x); the second slot now
    comes from an independent second hash and the table keeps its intended size."""
    z = synth.gen_zone(1500000, service_frac=0.15)
    st = Zone(z.jsonl, z.dns_domain).stat()
    assert st['slots'] == 8388608 and st['forward_keys'] == st['nodes']


def test_napi_addon_compiles_against_the_header():
    """Node.js is absent here, so the N-API shim (addon/binder_b200_napi.cc) is compiled against a mock <node_api.h>
    carrying the documented N-API signatures: every bb_* call in it must match include/binder_b200.h."""
    import shutil
    import subprocess
    cxx = shutil.which('g++')
    if not cxx:
        pytest.skip('no g++')
    p = subprocess.run([cxx, '-std=c++17', '-fsyntax-only', '-Wall', '-Wno-comment', '-Werror',
                        '-I', os.path.join(ROOT, 'tests', 'native', 'mock_node'), '-I', os.path.join(ROOT, 'include'),
                        os.path.join(ROOT, 'addon', 'binder_b200_napi.cc')], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-3000:]
