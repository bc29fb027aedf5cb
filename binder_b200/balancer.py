"""mname-balancer backend protocol adapter (wire format:
deps/mname-balancer/backend.c:22-113, frame types deps/mname-balancer/bbal.h:80-87).

The balancer forwards every UDP packet to a backend over an AF_UNIX stream as an INBOUND_UDP frame
(all integers u32 little-endian):

    0   u32 frame type = 2      4   u32 source IPv4      8   u32 source port
    12  u32 packet length       16  packet bytes

and expects OUTBOUND_UDP frames (type 1002, same layout, destination instead of source) back.
Frames are already "batched raw packets + source address", i.e. exactly the engine's batch
container plus 12 bytes of addressing per packet.  This module turns a buffer of frames into a
batch for bb_resolve_batch and the results back into OUTBOUND_UDP frames; HELLO / HEARTBEAT frames
are answered in place.  numpy only: no per-packet Python on the hot path.
"""
import numpy as np

CLIENT_HELLO, INBOUND_UDP, INBOUND_TCP, CLIENT_HEARTBEAT = 1, 2, 3, 4
SERVER_HELLO, OUTBOUND_UDP, INBOUND_TCP_OK, SERVER_HEARTBEAT = 1001, 1002, 1003, 1004
MAX_UDP = 1500                      # deps/mname-balancer/udp_proxy.c:159-170, backend.c:709-716


class ProtocolError(ValueError):
    pass


def parse_frames(buf):
    """bytes-like stream of balancer->backend frames ->
    (data u8 padded, off u32[n+1], src_ip u32[n], src_port u32[n], control [frame types], consumed)
    Only whole frames are consumed; `consumed` tells the caller how much of buf to drop."""
    b = np.frombuffer(buf, dtype=np.uint8)
    pos, n = 0, len(b)
    starts, lens, ips, ports, control = [], [], [], [], []
    while pos + 4 <= n:
        ftype = int(b[pos:pos + 4].view('<u4')[0])
        if ftype in (CLIENT_HELLO, CLIENT_HEARTBEAT):
            control.append(ftype); pos += 4
        elif ftype == INBOUND_UDP:
            if pos + 16 > n:
                break
            ip, port, ln = (int(x) for x in b[pos + 4:pos + 16].view('<u4'))
            if ln > MAX_UDP:
                raise ProtocolError('INBOUND_UDP frame of %d bytes' % ln)
            if pos + 16 + ln > n:
                break
            starts.append(pos + 16); lens.append(ln); ips.append(ip); ports.append(port)
            pos += 16 + ln
        elif ftype == INBOUND_TCP:
            raise ProtocolError('INBOUND_TCP converts the session to a TCP proxy: not handled by the batch path')
        else:
            raise ProtocolError('unknown frame type %d' % ftype)
    lens_a = np.asarray(lens, dtype=np.int64)
    off = np.zeros(len(lens) + 1, dtype=np.uint32)
    np.cumsum(lens_a, out=off[1:])
    total = int(off[-1])
    data = np.zeros((total + 15) // 16 * 16 + 16, dtype=np.uint8)
    if len(lens):
        # gather the packet bytes of all frames in one indexed copy
        idx = np.repeat(np.asarray(starts, dtype=np.int64) - off[:-1].astype(np.int64), lens_a) + np.arange(total)
        data[:total] = b[idx]
    return data, off, np.asarray(ips, dtype=np.uint32), np.asarray(ports, dtype=np.uint32), control, pos


def build_frames(out, out_off, out_len, status, dst_ip, dst_port, control=()):
    """Engine results -> bytes of backend->balancer frames: SERVER_HELLO / SERVER_HEARTBEAT for each
    control frame received, then one OUTBOUND_UDP frame per answered query (status 0)."""
    ctrl = np.asarray([SERVER_HELLO if c == CLIENT_HELLO else SERVER_HEARTBEAT for c in control], dtype='<u4').tobytes()
    ans = np.nonzero(np.asarray(status) == 0)[0]
    if ans.size == 0:
        return ctrl
    ln = np.asarray(out_len)[ans].astype(np.int64)
    fstart = np.zeros(ans.size + 1, dtype=np.int64)
    np.cumsum(ln + 16, out=fstart[1:])
    buf = np.zeros(int(fstart[-1]), dtype=np.uint8)
    hdr = np.stack([np.full(ans.size, OUTBOUND_UDP, dtype='<u4'), np.asarray(dst_ip)[ans].astype('<u4'),
                    np.asarray(dst_port)[ans].astype('<u4'), ln.astype('<u4')], axis=1).view(np.uint8).reshape(ans.size, 16)
    hidx = (fstart[:-1, None] + np.arange(16)[None, :]).reshape(-1)
    buf[hidx] = hdr.reshape(-1)
    total = int(ln.sum())
    src = np.repeat(np.asarray(out_off)[:-1][ans].astype(np.int64) - (np.cumsum(ln) - ln), ln) + np.arange(total)
    dst = np.repeat(fstart[:-1] + 16 - (np.cumsum(ln) - ln), ln) + np.arange(total)
    buf[dst] = np.asarray(out)[src]
    return ctrl + buf.tobytes()


class Backend(object):
    """One balancer session: feed() bytes read from the AF_UNIX socket, get bytes to write back."""

    def __init__(self, resolver, recursion=None, shuffle_seed=1):
        self.resolver, self.recursion, self.seed = resolver, recursion, shuffle_seed
        self.pending = b''
        self.counters = {'udp': 0, 'answered': 0, 'missed': 0, 'dropped': 0}

    def feed(self, chunk):
        self.pending += bytes(chunk)
        data, off, ips, ports, control, used = parse_frames(self.pending)
        self.pending = self.pending[used:]
        n = len(off) - 1
        if n == 0:
            return build_frames([], [0], [], [], [], [], control)
        self.seed += 1
        out, out_off, out_len, status, miss = self.resolver.resolve_batch(data, off, seed=self.seed, qidx_base=0)
        c = self.counters
        c['udp'] += n; c['answered'] += int((status == 0).sum()); c['missed'] += len(miss); c['dropped'] += int((status == 2).sum())
        if self.recursion is not None:
            for i in miss:
                i = int(i)
                self.recursion.resolve(bytes(data[off[i]:off[i + 1]]), (int(ips[i]), int(ports[i])))
        return build_frames(out, out_off, out_len, status, ips, ports, control)


# ---- the same protocol, natively (csrc/balancer_frames.cpp) ---------------------------------------------
def parse_frames_native(buf, cap_n=1 << 16, cap_bytes=1 << 22, cap_ctrl=64):
    """bb_frames_parse -> (data u8, off u32[n+1], src_ip, src_port, control list, consumed, rc)"""
    import ctypes
    from ._lib import lib
    buf = bytes(buf)
    data = np.zeros(cap_bytes + 32, dtype=np.uint8); off = np.zeros(cap_n + 1, dtype=np.uint32)
    ips = np.zeros(max(cap_n, 1), dtype=np.uint32); ports = np.zeros(max(cap_n, 1), dtype=np.uint32)
    ctrl = np.zeros(max(cap_ctrl, 1), dtype=np.uint32)
    n, nc, used = ctypes.c_uint32(0), ctypes.c_uint32(0), ctypes.c_size_t(0)
    rc = lib().bb_frames_parse(buf, len(buf), data.ctypes.data, cap_bytes, off.ctypes.data, ips.ctypes.data, ports.ctypes.data,
                               cap_n, ctypes.byref(n), ctrl.ctypes.data, cap_ctrl, ctypes.byref(nc), ctypes.byref(used))
    n = n.value
    return data[:(int(off[n]) + 15) // 16 * 16 + 16], off[:n + 1], ips[:n], ports[:n], [int(x) for x in ctrl[:nc.value]], used.value, rc


def build_frames_native(out, out_off, out_len, status, dst_ip, dst_port, control=()):
    """bb_frames_build -> bytes"""
    import ctypes
    from ._lib import check, lib
    n = len(status)
    a = lambda x, dt: np.ascontiguousarray(x, dtype=dt)
    out, out_off, out_len, status = a(out, np.uint8), a(out_off, np.uint32), a(out_len, np.uint16), a(status, np.uint8)
    dst_ip, dst_port, control = a(dst_ip, np.uint32), a(dst_port, np.uint32), a(list(control), np.uint32)
    need = ctypes.c_size_t(0)
    p = lambda x: x.ctypes.data if x.size else None
    args = (p(out), p(out_off), p(out_len), p(status), p(dst_ip), p(dst_port), n, p(control), len(control))
    rc = lib().bb_frames_build(*args, None, 0, ctypes.byref(need))
    if need.value == 0:
        return b''
    buf = np.zeros(need.value, dtype=np.uint8)
    check(lib().bb_frames_build(*args, buf.ctypes.data, buf.size, ctypes.byref(need)))
    return buf.tobytes()


class NativeBackend(object):
    """bb_backend: one balancer session in the C library (frames in -> bb_resolve_batch -> frames out)."""

    def __init__(self, engine, max_batch=1 << 14):
        import ctypes
        from ._lib import BinderError, lib
        err = ctypes.c_int(0)
        self._h = lib().bb_backend_create(engine._h, max_batch, ctypes.byref(err))
        if not self._h:
            raise BinderError(err.value)
        self._engine = engine

    def feed(self, chunk, seed):
        """-> (bytes to write back, [(packet, src_ip, src_port)] handed to recursion)"""
        import ctypes
        from ._lib import check, lib

        class Misses(ctypes.Structure):
            _fields_ = [('n', ctypes.c_uint32), ('pkts', ctypes.POINTER(ctypes.c_uint8)), ('pkt_off', ctypes.POINTER(ctypes.c_uint32)),
                        ('src_ip', ctypes.POINTER(ctypes.c_uint32)), ('src_port', ctypes.POINTER(ctypes.c_uint32))]
        chunk = bytes(chunk)
        out, out_len, m = ctypes.POINTER(ctypes.c_uint8)(), ctypes.c_size_t(0), Misses()
        check(lib().bb_backend_feed(self._h, chunk, len(chunk), seed, ctypes.byref(out), ctypes.byref(out_len), ctypes.byref(m)))
        data = ctypes.string_at(out, out_len.value) if out_len.value else b''
        misses = [(ctypes.string_at(ctypes.addressof(m.pkts.contents) + m.pkt_off[i], m.pkt_off[i + 1] - m.pkt_off[i]),
                   int(m.src_ip[i]), int(m.src_port[i])) for i in range(m.n)]
        return data, misses

    def stats(self):
        from ._lib import lib
        keys = ('udp', 'answered', 'missed', 'dropped', 'pending_bytes')
        return {k: int(lib().bb_backend_stat(self._h, i)) for i, k in enumerate(keys)}

    def close(self):
        from ._lib import lib
        if getattr(self, '_h', None):
            lib().bb_backend_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
