"""Host-side mirror of binder's server object (lib/server.js createServer / start / stop), in
Python because Node.js is not available in this image.  It keeps binder's shape — options
{host, port, dnsDomain, datacenterName, recursion, zkCache} and a UDP listener — but where mname
emits one 'query' event per packet (lib/server.js:471), packets are collected for a short window
and resolved as ONE batch by `resolver.resolve_batch` (a binder_b200.engine.Engine); the bytes it
returns are sent as they are.  Queries the engine hands back as recursion misses go to
options['recursion'].resolve(packet, addr) (lib/server.js:110-113,222-225) when one is given.

The resolver is injected, so the CPU-only plumbing test (BASELINE config 1) drives this same code
with the CPU oracle, while production passes an Engine.
"""
import socket
import threading
import time

import numpy as np

MAX_UDP = 1500          # deps/mname-balancer/udp_proxy.c:159-170


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
        self.counters = {'queries': 0, 'answered': 0, 'missed': 0, 'dropped': 0, 'batches': 0}
        self._sock = None
        self._thread = None
        self._stop = threading.Event()
        self._seed = options.get('shuffle_seed', int(time.time() * 1e6))

    # lib/server.js:609-653 (UDP listener only; TCP framing and the balancer socket stay host JS)
    def start(self, callback=None):
        fam = socket.AF_INET6 if ':' in self.host else socket.AF_INET
        self._sock = socket.socket(fam, socket.SOCK_DGRAM)
        self._sock.bind((self.host, self.port))
        self.port = self._sock.getsockname()[1]
        self._sock.settimeout(0.05)
        self._thread = threading.Thread(target=self._loop, daemon=True)
        self._thread.start()
        if callback:
            callback()
        return self

    def stop(self, callback=None):
        self._stop.set()
        if self._thread:
            self._thread.join(2)
        if self._sock:
            self._sock.close()
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

    def _resolve_and_send(self, batch):
        pkts = [p for p, _ in batch if len(p) <= MAX_UDP]
        addrs = [a for p, a in batch if len(p) <= MAX_UDP]
        lens = np.fromiter((len(p) for p in pkts), dtype=np.int64, count=len(pkts))
        off = np.zeros(len(pkts) + 1, dtype=np.uint32)
        np.cumsum(lens, out=off[1:])
        blob = b''.join(pkts)
        data = np.frombuffer(blob + b'\0' * ((-len(blob)) % 16 + 16), dtype=np.uint8)
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
