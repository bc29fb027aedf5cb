"""the mname-balancer backend protocol (deps/mname-balancer/backend.c:22-113)
as the batch ingress/egress.  Frames in -> engine batch -> frames out, checked packet by packet."""
import struct

import numpy as np
import pytest

import fuzzgen
import helpers as H
from binder_b200 import balancer, synth


def frame(pkt, ip, port):
    return struct.pack('<IIII', balancer.INBOUND_UDP, ip, port, len(pkt)) + pkt


def split_out(buf):
    pos, frames, ctrl = 0, [], []
    while pos < len(buf):
        t = struct.unpack_from('<I', buf, pos)[0]
        if t in (balancer.SERVER_HELLO, balancer.SERVER_HEARTBEAT):
            ctrl.append(t); pos += 4
        else:
            assert t == balancer.OUTBOUND_UDP
            _, ip, port, ln = struct.unpack_from('<IIII', buf, pos)
            frames.append((ip, port, bytes(buf[pos + 16:pos + 16 + ln]))); pos += 16 + ln
    return ctrl, frames


def _run(kind):
    snap, info = fuzzgen.gen_zone(3, n_top=30)
    impl = H.make_impl(kind, info['dns_domain'], snap, recursion=True)
    pkts = fuzzgen.gen_queries(3, info, n=800) + fuzzgen.malformed_packets()[:6]
    addr = [(0x0A000000 + i, 1024 + (i * 7) % 60000) for i in range(len(pkts))]
    stream = struct.pack('<I', balancer.CLIENT_HELLO) + b''.join(frame(p, *a) for p, a in zip(pkts, addr)) + \
        struct.pack('<I', balancer.CLIENT_HEARTBEAT)
    be = balancer.Backend(impl, shuffle_seed=41)
    # arbitrary chunking of the byte stream, like a socket would deliver it
    outs, pos, rng = [], 0, np.random.default_rng(0)
    while pos < len(stream):
        n = int(rng.integers(1, 4000))
        outs.append(be.feed(stream[pos:pos + n])); pos += n
    ctrl, frames = split_out(b''.join(outs))
    assert ctrl == [balancer.SERVER_HELLO, balancer.SERVER_HEARTBEAT] and be.pending == b''
    assert be.counters['udp'] == len(pkts) and be.counters['dropped'] >= 6
    by_addr = {(ip, port): wire for ip, port, wire in frames}
    assert len(by_addr) == len(frames) == be.counters['answered']
    for p, a in zip(pkts, addr):
        if a in by_addr:                     # the answer goes back to where the query came from, id and question echoed
            w = by_addr[a]
            assert w[:2] == p[:2] and w[2] & 0x80
    assert be.counters['answered'] + be.counters['missed'] + be.counters['dropped'] == len(pkts)
    return by_addr, dict(be.counters)


def test_balancer_frames_cpu_oracle():
    _run('oracle')


@pytest.mark.gpu
def test_balancer_frames_gpu_engine():
    """The same frame stream through an engine-backed and an oracle-backed session: every OUTBOUND_UDP frame — destination
    and payload bytes — and every counter must be identical (the chunking and the shuffle seed are the same on both sides)."""
    g_frames, g_cnt = _run('gpu')
    o_frames, o_cnt = _run('oracle')
    assert g_cnt == o_cnt
    assert g_frames.keys() == o_frames.keys()
    for a in o_frames:
        assert g_frames[a] == o_frames[a], (a, g_frames[a].hex(), o_frames[a].hex())


def test_protocol_errors():
    with pytest.raises(balancer.ProtocolError):
        balancer.parse_frames(struct.pack('<I', 77))
    with pytest.raises(balancer.ProtocolError):
        balancer.parse_frames(struct.pack('<IIII', 2, 1, 2, 5000) + b'x' * 5000)
    d, off, ips, ports, ctrl, used = balancer.parse_frames(struct.pack('<IIII', 2, 1, 2, 40) + b'x' * 10)
    assert used == 0 and len(off) == 1          # partial frame: nothing consumed


# ---- the native implementation (csrc/balancer_frames.cpp) against the numpy one -------------------------
def _stream(seed, n=600):
    snap, info = fuzzgen.gen_zone(seed, n_top=30)
    pkts = fuzzgen.gen_queries(seed, info, n=n) + fuzzgen.malformed_packets()[:6]
    addr = [(0x0A000000 + i, 1024 + (i * 7) % 60000) for i in range(len(pkts))]
    rng = np.random.default_rng(seed)
    parts = [struct.pack('<I', balancer.CLIENT_HELLO)]
    for p, a in zip(pkts, addr):
        parts.append(frame(p, *a))
        if rng.random() < 0.01:
            parts.append(struct.pack('<I', balancer.CLIENT_HEARTBEAT))
    return snap, info, pkts, addr, b''.join(parts)


@pytest.mark.parametrize('seed', range(4))
def test_native_frame_parser_matches(seed):
    _, _, pkts, addr, stream = _stream(seed)
    rng = np.random.default_rng(seed + 9)
    for cut in [len(stream)] + [int(x) for x in rng.integers(0, len(stream), size=40)]:
        d0, o0, i0, p0, c0, u0 = balancer.parse_frames(stream[:cut])
        d1, o1, i1, p1, c1, u1, rc = balancer.parse_frames_native(stream[:cut])
        assert rc == 0 and u0 == u1 and c0 == c1
        assert np.array_equal(o0, o1) and np.array_equal(i0, i1) and np.array_equal(p0, p1)
        assert np.array_equal(d0[:o0[-1]], d1[:o1[-1]])
    # a full batch stops the parser without consuming what does not fit
    d, o, i, p, c, used, rc = balancer.parse_frames_native(stream, cap_n=10)
    assert rc == 0 and len(o) == 11 and stream[used:used + 4] == struct.pack('<I', balancer.INBOUND_UDP)
    d, o, i, p, c, used2, rc = balancer.parse_frames_native(stream, cap_bytes=200)
    assert rc == 0 and o[-1] <= 200 and used2 < len(stream)


def test_native_frame_builder_matches():
    rng = np.random.default_rng(5)
    n = 500
    lens = rng.integers(0, 300, size=n).astype(np.uint16)
    status = rng.choice([0, 0, 0, 1, 2], size=n).astype(np.uint8)
    lens[status != 0] = 0
    off = np.zeros(n + 1, dtype=np.uint32); np.cumsum(lens.astype(np.int64), out=off[1:])
    perm_off = off.copy()
    out = rng.integers(0, 256, size=int(off[-1]), dtype=np.uint8)
    ips = rng.integers(0, 2 ** 32, size=n, dtype=np.uint32); ports = rng.integers(0, 65536, size=n).astype(np.uint32)
    ctrl = [balancer.CLIENT_HELLO, balancer.CLIENT_HEARTBEAT, balancer.CLIENT_HEARTBEAT]
    a = balancer.build_frames(out, perm_off, lens, status, ips, ports, ctrl)
    b = balancer.build_frames_native(out, perm_off, lens, status, ips, ports, ctrl)
    assert a == b and len(a) == 12 + int((status == 0).sum()) * 16 + int(lens.sum())
    assert balancer.build_frames_native([], [0], [], [], [], [], []) == b''
    assert balancer.build_frames_native([], [0], [], [], [], [], [balancer.CLIENT_HELLO]) == struct.pack('<I', balancer.SERVER_HELLO)


def test_native_protocol_errors():
    from binder_b200._lib import BinderError
    assert balancer.parse_frames_native(struct.pack('<I', 77))[-1] == -8
    assert balancer.parse_frames_native(struct.pack('<IIII', 2, 1, 2, 5000) + b'x' * 5000)[-1] == -8
    assert balancer.parse_frames_native(struct.pack('<I', balancer.INBOUND_TCP))[-1] == -8
    good = frame(b'x' * 30, 1, 2)
    d, o, i, p, c, used, rc = balancer.parse_frames_native(good + struct.pack('<I', 77))
    assert rc == -8 and used == len(good) and len(o) == 2          # what preceded the bad frame is still returned
    d, o, i, p, c, used, rc = balancer.parse_frames_native(struct.pack('<IIII', 2, 1, 2, 40) + b'x' * 10)
    assert rc == 0 and used == 0 and len(o) == 1
    with pytest.raises(BinderError):
        rng = np.zeros(4, dtype=np.uint8)
        balancer_build_too_small = __import__('ctypes').c_size_t(0)
        from binder_b200._lib import check, lib
        st = np.zeros(1, dtype=np.uint8); ln = np.full(1, 4, dtype=np.uint16); off = np.zeros(2, dtype=np.uint32)
        ip = np.zeros(1, dtype=np.uint32)
        check(lib().bb_frames_build(rng.ctypes.data, off.ctypes.data, ln.ctypes.data, st.ctypes.data, ip.ctypes.data, ip.ctypes.data, 1,
                                    None, 0, rng.ctypes.data, 4, __import__('ctypes').byref(balancer_build_too_small)))


@pytest.mark.gpu
@pytest.mark.parametrize('seed', range(3))
def test_native_backend_session_matches_python(seed):
    """bb_backend_feed == parse_frames + resolve_batch + build_frames, chunk by chunk, misses included."""
    from binder_b200.engine import Engine
    snap, info, pkts, addr, stream = _stream(seed + 20, n=1500)
    eng = Engine(info['dns_domain'], recursion=True, snapshot=snap, ordered=True)
    nb = balancer.NativeBackend(eng, max_batch=4096)
    rng = np.random.default_rng(seed)
    pending, pos, k = b'', 0, 0
    n_miss = 0
    while pos < len(stream):
        step = int(rng.integers(1, 9000)); chunk = stream[pos:pos + step]; pos += step; k += 1
        got, misses = nb.feed(chunk, seed=1000 + k)
        pending += chunk
        data, off, ips, ports, control, used = balancer.parse_frames(pending)
        pending = pending[used:]
        if len(off) > 1:
            out, ooff, olen, status, miss = eng.resolve_batch(data, off, seed=1000 + k)
            want = balancer.build_frames(out, ooff, olen, status, ips, ports, control)
            want_miss = [(bytes(data[off[i]:off[i + 1]]), int(ips[i]), int(ports[i])) for i in sorted(int(x) for x in miss)]
        else:
            want = balancer.build_frames([], [0], [], [], [], [], control); want_miss = []
        assert got == want and misses == want_miss
        n_miss += len(misses)
    st = nb.stats()
    assert st['udp'] == len(pkts) and st['pending_bytes'] == 0 and st['missed'] == n_miss and st['dropped'] >= 6
    nb.close()
