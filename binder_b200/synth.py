"""Seeded synthetic zones and DNS query batches (SURVEY.md §8d "Zone generator").

A zone snapshot is JSON lines, one znode per line, parents before children, children in
ZooKeeper child-list order:

    {"path": "/com/example/dc1/g0012/h0001234", "data": {"type": "host", "host": {...}}}

`data` is what JSON.parse() of the znode bytes yields (lib/zk.js:139-155); `"raw": "<bytes>"`
may be used instead to carry unparsed znode content.  The record shapes are the ones the
reference serves (test/host.test.js:58-63, test/service.test.js:28-47,
test/database.test.js:28-35).
"""
import struct

import numpy as np

DNS_DOMAIN = 'dc1.example.com'
DATACENTER = 'dc1'
ROOT_PATH = '/com/example/dc1'
N_GROUPS = 1024

QTYPE = {'A': 1, 'NS': 2, 'PTR': 12, 'TXT': 16, 'AAAA': 28, 'SRV': 33, 'ANY': 255}


def host_name(i):
    return 'h%07d.g%04d.%s' % (i, i % N_GROUPS, DNS_DOMAIN)


def host_addr(i):
    return '10.%d.%d.%d' % ((i >> 16) & 255, (i >> 8) & 255, i & 255)


def svc_name(j):
    return 'svc%06d.%s' % (j, DNS_DOMAIN)


def svc_fanout(j):
    """Number of children of service j: 1..8, deterministic."""
    return 1 + (j * 2654435761 >> 7) % 8


class Zone(object):
    """A generated snapshot plus what the query generators need to know about it."""

    def __init__(self, jsonl, n_records, n_hosts, n_services, n_db):
        self.jsonl = jsonl              # bytes
        self.n_records = n_records
        self.n_hosts = n_hosts
        self.n_services = n_services
        self.n_db = n_db
        self.dns_domain = DNS_DOMAIN
        self.datacenter = DATACENTER


def gen_zone(n_records, service_frac=0.0):
    """Zone with ~n_records znodes (root excluded).  service_frac = fraction of the znode
    budget spent on services + their children (0 for config 2; 0.15 for configs 3-5:
    ~300k services x avg 4.5 kids + 1 in a 10M zone)."""
    lines = []
    budget = n_records
    lines.append('{"path":"%s","data":null}' % ROOT_PATH)
    for g in range(N_GROUPS):                     # null-data intermediates -> SERVFAIL when queried
        lines.append('{"path":"%s/g%04d","data":null}' % (ROOT_PATH, g))
    budget -= N_GROUPS
    svc_budget = int(n_records * service_frac)
    n_services = 0
    used = 0
    svc_lines = []
    while used < svc_budget:
        j = n_services
        k = svc_fanout(j)
        sttl = ',"ttl":%d' % (40 + j % 50) if j % 5 == 0 else ''
        rttl = ',"ttl":%d' % (60 + j % 30) if j % 7 == 0 else ''
        base = '%s/svc%06d' % (ROOT_PATH, j)
        svc_lines.append('{"path":"%s","data":{"type":"service","service":{"srvce":"_http",'
                         '"proto":"_tcp","port":%d%s}%s}}' % (base, 80 + j % 3, sttl, rttl))
        for c in range(k):
            sel = (j * 31 + c * 7) % 10
            a = '172.%d.%d.%d' % (16 + (j >> 13) % 16, (j >> 5) & 255, ((j & 31) << 3 | c) & 255)
            if sel < 8:
                kttl = ',"ttl":%d' % (20 + c) if (j + c) % 6 == 0 else ''
                svc_lines.append('{"path":"%s/lb%02d","data":{"type":"load_balancer",'
                                 '"load_balancer":{"address":"%s"%s}}}' % (base, c, a, kttl))
            elif sel == 8:
                svc_lines.append('{"path":"%s/lb%02d","data":{"type":"rr_host","rr_host":'
                                 '{"address":"%s","ports":[%d,%d]},"ttl":%d}}'
                                 % (base, c, a, 8000 + c, 9000 + c, 15 + c))
            else:
                svc_lines.append('{"path":"%s/lb%02d","data":{"type":"host","host":'
                                 '{"address":"%s"}}}' % (base, c, a))
        used += 1 + k
        n_services += 1
    budget -= used
    n_db = min(max(budget // 1000, 0), 4096)
    db_lines = ['{"path":"%s/db%05d","data":{"type":"database","database":{"primary":'
                '"tcp://user@192.168.%d.%d/postgres","standby":"tcp://user@192.168.0.2/postgres"}}}'
                % (ROOT_PATH, d, d >> 8, d & 255) for d in range(n_db)]
    budget -= n_db
    n_hosts = max(budget, 0)
    # hosts: children of the group nodes, grouped so that parents precede children
    host_lines = []
    for i in range(n_hosts):
        m = i % 10
        if m == 3:
            rec = '{"type":"host","host":{"address":"%s"},"ttl":%d}' % (host_addr(i), 120 + i % 7)
        elif m == 7:
            rec = '{"type":"host","host":{"address":"%s","ttl":%d}}' % (host_addr(i), 300 + i % 11)
        else:
            rec = '{"type":"host","host":{"address":"%s"}}' % host_addr(i)
        host_lines.append('{"path":"%s/g%04d/h%07d","data":%s}' % (ROOT_PATH, i % N_GROUPS, i, rec))
    lines.extend(host_lines)
    lines.extend(svc_lines)
    lines.extend(db_lines)
    jsonl = ('\n'.join(lines) + '\n').encode('ascii')
    return Zone(jsonl, len(lines) - 1, n_hosts, n_services, n_db)


# ---------------------------------------------------------------------------------------
# query packets
# ---------------------------------------------------------------------------------------
def encode_name(name):
    """Dotted name (str or bytes) -> wire labels."""
    if isinstance(name, str):
        name = name.encode('latin-1')
    out = bytearray()
    if name:
        for lab in name.split(b'.'):
            out.append(len(lab))
            out += lab
    out.append(0)
    return bytes(out)


def make_query(name, qtype, qid=0x1234, rd=True, edns=None, opcode=0, qclass=1, labels=None):
    """One DNS query packet.  edns = advertised UDP size (adds an OPT RR) or None.
    `labels` (list of bytes) overrides `name` to build names a dotted string cannot express."""
    if isinstance(qtype, str):
        qtype = QTYPE[qtype]
    flags = (opcode << 11) | (0x0100 if rd else 0)
    if labels is not None:
        wire = b''.join(bytes([len(l)]) + l for l in labels) + b'\0'
    else:
        wire = encode_name(name)
    pkt = struct.pack('>HHHHHH', qid, flags, 1, 0, 0, 1 if edns else 0) + wire + \
        struct.pack('>HH', qtype, qclass)
    if edns:
        pkt += b'\0' + struct.pack('>HHIH', 41, edns, 0, 0)
    return pkt


def pack_batch(pkts):
    """list of packet bytes -> (uint8 array, uint32 offsets[n+1]); the byte array is padded
    to a multiple of 16 (the device path reads 16-byte vectors)."""
    lens = np.fromiter((len(p) for p in pkts), dtype=np.int64, count=len(pkts))
    off = np.zeros(len(pkts) + 1, dtype=np.uint32)
    np.cumsum(lens, out=off[1:])
    blob = b''.join(pkts)
    pad = (-len(blob)) % 16
    data = np.frombuffer(blob + b'\0' * pad, dtype=np.uint8).copy()
    return data, off


def batch_host_a(zone, n, seed, miss_frac=0.0, rd=True):
    """Config 2 / 5: A lookups on host names drawn uniformly from the zone; a `miss_frac`
    share uses indices beyond the zone (same shape, absent)."""
    rng = np.random.default_rng(seed)
    idx = rng.integers(0, zone.n_hosts, size=n)
    if miss_frac > 0:
        miss = rng.random(n) < miss_frac
        idx = np.where(miss, idx + zone.n_hosts + 7, idx)
    ids = rng.integers(0, 65536, size=n)
    return [make_query(host_name(int(i)), 1, int(q), rd=rd) for i, q in zip(idx, ids)]


def batch_service(zone, n, seed, srv_frac=0.5):
    """Config 3: srv_frac SRV `_http._tcp.svcN`, rest A on the service name."""
    rng = np.random.default_rng(seed)
    idx = rng.integers(0, zone.n_services, size=n)
    is_srv = rng.random(n) < srv_frac
    ids = rng.integers(0, 65536, size=n)
    out = []
    for j, s, q in zip(idx, is_srv, ids):
        if s:
            out.append(make_query('_http._tcp.' + svc_name(int(j)), 33, int(q)))
        else:
            out.append(make_query(svc_name(int(j)), 1, int(q)))
    return out


def batch_mixed(zone, n, seed, miss_frac=0.0):
    """Config 4: 60 % A(host) / 20 % SRV / 20 % AAAA (-> NOTIMP, lib/server.js:491-506)."""
    rng = np.random.default_rng(seed)
    kind = rng.random(n)
    hidx = rng.integers(0, max(zone.n_hosts, 1), size=n)
    sidx = rng.integers(0, max(zone.n_services, 1), size=n)
    miss = rng.random(n) < miss_frac
    ids = rng.integers(0, 65536, size=n)
    out = []
    for k, h, s, m, q in zip(kind, hidx, sidx, miss, ids):
        h = int(h) + (zone.n_hosts + 7 if m else 0)
        if k < 0.6 or zone.n_services == 0:
            out.append(make_query(host_name(h), 1 if k < 0.8 else 28, int(q)))
        elif k < 0.8:
            out.append(make_query('_http._tcp.' + svc_name(int(s)), 33, int(q)))
        else:
            out.append(make_query(host_name(h), 28, int(q)))
    return out


def batch_host_a_fast(zone, n, seed, miss_frac=0.0, rd=True):
    """Vectorised batch_host_a: -> (uint8 data padded to 16, uint32 off[n+1]).  Every packet is
    the 48-byte query for h%07d.g%04d.dc1.example.com, type A, class IN."""
    rng = np.random.default_rng(seed)
    idx = rng.integers(0, zone.n_hosts, size=n)
    if miss_frac > 0:
        miss = rng.random(n) < miss_frac
        idx = np.where(miss, idx + zone.n_hosts + 7, idx)
    ids = rng.integers(0, 65536, size=n)
    tmpl = np.frombuffer(make_query(host_name(0), 1, 0, rd=rd), dtype=np.uint8)
    assert tmpl.size == 48
    pk = np.tile(tmpl, (n, 1))
    pk[:, 0] = ids >> 8
    pk[:, 1] = ids & 255
    v = idx.copy()
    for k in range(7):
        pk[:, 20 - k] = 48 + v % 10
        v //= 10
    g = idx % N_GROUPS
    for k in range(4):
        pk[:, 26 - k] = 48 + g % 10
        g //= 10
    data = np.zeros((n * 48 + 15) // 16 * 16, dtype=np.uint8)
    data[:n * 48] = pk.reshape(-1)
    off = (np.arange(n + 1, dtype=np.uint64) * 48).astype(np.uint32)
    return data, off


# ---------------------------------------------------------------------------------------
# vectorised workload generators (BASELINE.json configs 2-5 at full size) + SURVEY.md §8(d) byte model
# ---------------------------------------------------------------------------------------
K_HOST_A, K_HOST_AAAA, K_SVC_A, K_SVC_SRV = 0, 1, 2, 3


def _put_digits(pk, col_last, ndig, v):
    v = v.copy()
    for k in range(ndig):
        pk[:, col_last - k] = 48 + v % 10
        v //= 10


def gen_batch(zone, n, seed, mix, miss_frac=0.0, rd=True):
    """n packets drawn from `mix` = {K_*: share}; -> (data uint8 padded to 16, off uint32[n+1], meta).
    meta = dict(kind int8[n], idx int64[n] (host or service index; >= n_hosts: absent name)).
    K_HOST_A/K_HOST_AAAA: h%07d.g%04d.<dom>;  K_SVC_A: svc%06d.<dom>;  K_SVC_SRV: _http._tcp.svc%06d.<dom>."""
    rng = np.random.default_rng(seed)
    kinds = np.array(sorted(mix), dtype=np.int8)
    shares = np.array([mix[k] for k in kinds], dtype=np.float64)
    kind = kinds[np.searchsorted(np.cumsum(shares / shares.sum()), rng.random(n), side='right').clip(0, len(kinds) - 1)]
    is_host = kind <= K_HOST_AAAA
    idx = np.where(is_host, rng.integers(0, max(zone.n_hosts, 1), size=n), rng.integers(0, max(zone.n_services, 1), size=n))
    if miss_frac > 0:
        miss = (rng.random(n) < miss_frac) & is_host
        idx = np.where(miss, idx + zone.n_hosts + 7, idx)
    ids = rng.integers(0, 65536, size=n)
    tm = {K_HOST_A: make_query(host_name(0), 1, 0, rd=rd), K_HOST_AAAA: make_query(host_name(0), 28, 0, rd=rd),
          K_SVC_A: make_query(svc_name(0), 1, 0, rd=rd), K_SVC_SRV: make_query('_http._tcp.' + svc_name(0), 33, 0, rd=rd)}
    W = max(len(t) for t in tm.values())
    pk = np.zeros((n, W), dtype=np.uint8)
    lens = np.zeros(n, dtype=np.int64)
    for k, t in tm.items():
        m = kind == k
        if not m.any():
            continue
        sub = np.tile(np.frombuffer(t, dtype=np.uint8), (int(m.sum()), 1))
        v = idx[m]
        if k <= K_HOST_AAAA:
            _put_digits(sub, 20, 7, v)
            _put_digits(sub, 26, 4, v % N_GROUPS)
        else:
            _put_digits(sub, (21 if k == K_SVC_A else 32), 6, v)
        pk[m, :len(t)] = sub
        lens[m] = len(t)
    pk[:, 0] = ids >> 8
    pk[:, 1] = ids & 255
    off = np.zeros(n + 1, dtype=np.uint32)
    np.cumsum(lens, out=off[1:])
    keep = np.arange(W)[None, :] < lens[:, None]
    flat = pk[keep]
    data = np.zeros((flat.size + 15) // 16 * 16, dtype=np.uint8)
    data[:flat.size] = flat
    return data, off, dict(kind=kind, idx=idx)


WORKLOADS = {
    # name: (description, zone service_frac, mix, miss_frac, recursion)
    'config2': ('config2: 1M-record zone, 65536-query A-record batches (100% hit, RD=1, no OPT)', 0.0, {K_HOST_A: 1.0}, 0.0, False),
    'config3': ('config3: 10M-record zone, 262144-query batches, 50% SRV _http._tcp.svcN / 50% A on service names', 0.15,
                {K_SVC_SRV: 0.5, K_SVC_A: 0.5}, 0.0, False),
    'config4': ('config4: 10M-record zone, 1048576-query global batch, 60% A(host) / 20% SRV(service) / 20% AAAA(->NOTIMP)', 0.15,
                {K_HOST_A: 0.6, K_SVC_SRV: 0.2, K_HOST_AAAA: 0.2}, 0.0, False),
    'config5': ('config5: 10M-record zone, A(host) lookups, 90% of names absent, RD=1, recursion on (misses -> compacted miss list)', 0.15,
                {K_HOST_A: 1.0}, 0.9, True),
    'config5_rd0': ('config5 with RD=0: the same 90%-absent names, recursion not desired -> every miss is answered REFUSED on the device', 0.15,
                    {K_HOST_A: 1.0}, 0.9, True),
}
WORKLOAD_RD = {'config5_rd0': False}     # RD bit of a workload's queries (default set)


def service_payload_bytes(zone):
    """SURVEY.md §8(d) payload(service j) = 16 + sum over member children (len(label)+1 + 4 addr + 4 ttl + 2*nports),
    for every generated service (children typed `host` are not members, lib/server.js:352-360)."""
    j = np.arange(max(zone.n_services, 1), dtype=np.int64)
    k = 1 + (j * 2654435761 >> 7) % 8
    pay = np.full(j.shape, 16, dtype=np.int64)
    for c in range(8):
        sel = (j * 31 + c * 7) % 10
        has = c < k
        pay += np.where(has & (sel < 8), 4 + 1 + 4 + 4 + 2, 0) + np.where(has & (sel == 8), 4 + 1 + 4 + 4 + 4, 0)
    return pay


def algorithmic_bytes(zone, off, meta, out_len):
    """B(q) = len(query)+4 + probe(q) + len(response)+8 summed over a batch -> (read, write).
    probe: hit = len(key)+1 + payload(node) (host 8; service 16 + members); miss = 8; no lookup (AAAA -> NOTIMP) = 0."""
    n = len(off) - 1
    kind, idx = meta['kind'], meta['idx']
    host_key = len(host_name(0)) + 1
    svc_key = len(svc_name(0)) + 1
    probe = np.zeros(n, dtype=np.int64)
    hostq = kind == K_HOST_A
    probe[hostq] = np.where(idx[hostq] < zone.n_hosts, host_key + 8, 8)
    svcq = kind >= K_SVC_A
    if svcq.any():
        probe[svcq] = svc_key + service_payload_bytes(zone)[idx[svcq]]
    read = int(off[n]) + 4 * n + int(probe.sum())
    write = int(np.asarray(out_len, dtype=np.int64).sum()) + 8 * n
    return read, write
