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
