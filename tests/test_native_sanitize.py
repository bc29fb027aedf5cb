"""The host-side C++ (zone builder + delta applier, balancer frames) under AddressSanitizer and
UndefinedBehaviorSanitizer, fed garbage: tests/native/fuzz_host.cpp."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_code_is_sanitizer_clean(tmp_path):
    cxx = shutil.which('g++')
    if not cxx:
        pytest.skip('no g++')
    exe = str(tmp_path / 'fuzz_host')
    srcs = [os.path.join(ROOT, 'tests', 'native', f) for f in ('fuzz_host.cpp', 'engine_stubs.cpp')] + \
           [os.path.join(ROOT, 'binder_b200', 'csrc', f) for f in ('zone_build.cpp', 'balancer_frames.cpp')]
    cmd = [cxx, '-std=c++17', '-O1', '-g', '-fsanitize=address,undefined', '-fno-omit-frame-pointer',
           '-I', os.path.join(ROOT, 'include'), '-I', os.path.join(ROOT, 'binder_b200', 'csrc'), '-o', exe] + srcs
    b = subprocess.run(cmd, capture_output=True, text=True)
    if b.returncode != 0 and 'sanitize' in b.stderr:
        pytest.skip('toolchain without sanitizer runtimes')
    assert b.returncode == 0, b.stderr[-2000:]
    p = subprocess.run([exe, '6000', '1200'], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and p.stdout.strip().endswith('OK') and 'runtime error' not in p.stderr, (p.stdout[-500:], p.stderr[-3000:])
