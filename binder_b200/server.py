"""Host-side mirror of binder's server object (lib/server.js createServer / start / stop), in
Python because Node.js is not available in this image.  It keeps binder's shape — options
{host, port, dnsDomain, datacenterName, recursion, zkCache} and a UDP listener — but where mname
emits one 'query' event per packet (lib/server.js:471), packets are collected for a short window
and resolved as ONE batch by `resolver.resolve_batch` (a binder_b200.engine.Engine); the bytes it
returns are sent as they are.  Queries the engine hands back as recursion misses go to
options['recursion'].resolve(packet, addr) (lib/server.js:110-113,222-225) when one is given.

TCP (mname's listenTcp, lib/server.js:643-652): every connection is a stream of messages, each preceded
by its length as two big-endian bytes (RFC 1035 4.2.2; deps/mname-balancer/tcp_proxy.c:10-26 relays the
same framing).  TcpFramer reassembles them; messages from all connections that complete inside one
aggregation window form ONE batch, resolved with tcp=True (no UDP size limit), and each response goes
back on its connection with its own length prefix, in arrival order per connection.

The resolver is injected, so the CPU-only plumbing test (BASELINE config 1) drives this same code
with the CPU oracle, while production passes an Engine.
"""
import select
import socket
import struct
import threading
import time

import numpy as np

MAX_UDP = 1500          # deps/mname-balancer/udp_proxy.c:159-170


class TcpFramer(object):
    """Reassembly of one DNS-over-TCP stream: feed() bytes as they arrive, get back the complete messages.
    A zero-length message is a protocol error (the connection is closed by the caller)."""

    def __init__(self):
        self.buf = bytearray()
        self.bad = False

    def feed(self, chunk):
        self.buf += chunk
        out = []
        while len(self.buf) >= 2:
            n = (self.buf[0] << 8) | self.buf[1]
            if n == 0:
                self.bad = True
                break
            if len(self.buf) < 2 + n:
                break
            out.append(bytes(self.buf[2:2 + n]))
            del self.buf[:2 + n]
        return out

    @staticmethod
    def frame(msg):
        return struct.pack('>H', len(msg)) + msg


def pack_packets(pkts):
    """[bytes] -> (uint8 data padded for 16-byte reads, uint32 offsets[n+1])"""
    lens = np.fromiter((len(p) for p in pkts), dtype=np.int64, count=len(pkts))
    off = np.zeros(len(pkts) + 1, dtype=np.uint32)
    np.cumsum(lens, out=off[1:])
    blob = b''.join(pkts)
    return np.frombuffer(blob + b'\0' * ((-len(blob)) % 16 + 16), dtype=np.uint8), off


class Server(object):
    def __init__(self, options):
        for k in ('dnsDomain', 'resolver'):
            if k not in options:
                raise ValueError('options.%s is required' % k)          # assert.string(options.dnsDomain), lib/server.js:439
        self.options = options
        self.resolver = options['resolver']
        self.host = options.get('host', '127.0.0.1')
        self.port = options.get('port', 0)
        self.window_s = options.get('batch_window_us', 200) * 1e-6
        self.max_batch = options.get('max_batch', 4096)
        self.recursion = options.get('recursion')
        self.counters = {'queries': 0, 'answered': 0, 'missed': 0, 'dropped': 0, 'batches': 0,
                         'tcp_queries': 0, 'tcp_batches': 0, 'tcp_connections': 0}
        self.tcp = options.get('tcp', True)
        self._sock = None
        self._tsock = None
        self._tthread = None
        self._thread = None
        self._stop = threading.Event()
        self._seed = options.get('shuffle_seed', int(time.time() * 1e6))

    # lib/server.js:609-653: UDP and TCP on the same port (the balancer socket stays host JS)
    def start(self, callback=None):
        fam = socket.AF_INET6 if ':' in self.host else socket.AF_INET
        self._sock = socket.socket(fam, socket.SOCK_DGRAM)
        self._sock.bind((self.host, self.port))
        self.port = self._sock.getsockname()[1]
        self._sock.settimeout(0.05)
        self._thread = threading.Thread(target=self._loop, daemon=True)
        self._thread.start()
        if self.tcp:
            self._tsock = socket.socket(fam, socket.SOCK_STREAM)
            self._tsock.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
            self._tsock.bind((self.host, self.port))
            self._tsock.listen(128)
            self._tsock.setblocking(False)
            self._tthread = threading.Thread(target=self._tcp_loop, daemon=True)
            self._tthread.start()
        if callback:
            callback()
        return self

    def stop(self, callback=None):
        self._stop.set()
        if self._thread:
            self._thread.join(2)
        if self._tthread:
            self._tthread.join(2)
        if self._sock:
            self._sock.close()
        if self._tsock:
            self._tsock.close()
        if callback:
            callback()

    def _loop(self):
        while not self._stop.is_set():
            try:
                pkt, addr = self._sock.recvfrom(MAX_UDP + 1)
            except socket.timeout:
                continue
            except OSError:
                break
            batch = [(pkt, addr)]
            deadline = time.perf_counter() + self.window_s
            self._sock.settimeout(max(self.window_s, 1e-4))
            while len(batch) < self.max_batch and time.perf_counter() < deadline:
                try:
                    batch.append(self._sock.recvfrom(MAX_UDP + 1))
                except (socket.timeout, BlockingIOError):
                    break
            self._sock.settimeout(0.05)
            self._resolve_and_send(batch)

    # ---- TCP: accept, reassemble, aggregate, answer ------------------------------------------------
    def _tcp_loop(self):
        conns = {}                                   # socket -> TcpFramer
        pending = []                                 # (message, socket) completed in the current window
        deadline = None
        while not self._stop.is_set():
            timeout = 0.05 if deadline is None else max(deadline - time.perf_counter(), 0.0)
            try:
                ready, _, _ = select.select([self._tsock] + list(conns), [], [], timeout)
            except (OSError, ValueError):
                break
            for s in ready:
                if s is self._tsock:
                    try:
                        c, _ = self._tsock.accept()
                    except OSError:
                        continue
                    c.setblocking(False)
                    conns[c] = TcpFramer()
                    self.counters['tcp_connections'] += 1
                    continue
                try:
                    chunk = s.recv(65536)
                except (BlockingIOError, InterruptedError):
                    continue
                except OSError:
                    chunk = b''
                fr = conns[s]
                msgs = fr.feed(chunk) if chunk else []
                for m in msgs:
                    pending.append((m, s))
                if msgs and deadline is None:
                    deadline = time.perf_counter() + self.window_s
                if not chunk or fr.bad:              # peer closed (its complete messages are still answered) / bad frame
                    del conns[s]
                    if not any(ps is s for _, ps in pending):
                        s.close()
            if pending and (len(pending) >= self.max_batch or time.perf_counter() >= deadline):
                self._resolve_and_send_tcp(pending, conns)
                pending, deadline = [], None
        for s in conns:
            s.close()

    def _resolve_and_send_tcp(self, pending, conns):
        pkts = [m for m, _ in pending]
        data, off = pack_packets(pkts)
        self._seed += 1
        out, out_off, out_len, status, miss = self.resolver.resolve_batch(data, off, seed=self._seed, qidx_base=0, tcp=True)
        c = self.counters
        c['tcp_batches'] += 1
        c['tcp_queries'] += len(pkts)
        replies = {}                                 # per connection, in arrival order
        for i, (_, s) in enumerate(pending):
            if status[i] == 0:
                replies.setdefault(s, []).append(TcpFramer.frame(out[out_off[i]:out_off[i] + out_len[i]].tobytes()))
                c['answered'] += 1
            elif status[i] == 2:
                c['dropped'] += 1
        for s, parts in replies.items():
            try:
                s.setblocking(True)
                s.sendall(b''.join(parts))
                s.setblocking(False)
            except OSError:
                pass
        for i in miss:
            c['missed'] += 1
            if self.recursion is not None:
                self.recursion.resolve(pkts[int(i)], None, pending[int(i)][1])
        for s in set(ps for _, ps in pending):
            if s not in conns:                       # the peer had already closed its side
                s.close()

    def _resolve_and_send(self, batch):
        pkts = [p for p, _ in batch if len(p) <= MAX_UDP]
        addrs = [a for p, a in batch if len(p) <= MAX_UDP]
        data, off = pack_packets(pkts)
        self._seed += 1
        out, out_off, out_len, status, miss = self.resolver.resolve_batch(data, off, seed=self._seed, qidx_base=0)
        c = self.counters
        c['batches'] += 1
        c['queries'] += len(pkts)
        for i, a in enumerate(addrs):
            if status[i] == 0:
                self._sock.sendto(out[out_off[i]:out_off[i] + out_len[i]].tobytes(), a)
                c['answered'] += 1
            elif status[i] == 2:
                c['dropped'] += 1
        for i in miss:
            c['missed'] += 1
            if self.recursion is not None:
                self.recursion.resolve(pkts[int(i)], addrs[int(i)], self._sock)


def createServer(options):
    """core.createServer(options) (lib/server.js:435)."""
    return Server(options)
