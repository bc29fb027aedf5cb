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


def test_balancer_frames_cpu_oracle():
    _run('oracle')


@pytest.mark.gpu
def test_balancer_frames_gpu_engine():
    _run('gpu')


def test_protocol_errors():
    with pytest.raises(balancer.ProtocolError):
        balancer.parse_frames(struct.pack('<I', 77))
    with pytest.raises(balancer.ProtocolError):
        balancer.parse_frames(struct.pack('<IIII', 2, 1, 2, 5000) + b'x' * 5000)
    d, off, ips, ports, ctrl, used = balancer.parse_frames(struct.pack('<IIII', 2, 1, 2, 40) + b'x' * 10)
    assert used == 0 and len(off) == 1          # partial frame: nothing consumed
