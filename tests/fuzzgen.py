"""Seeded adversarial zones + query sets, shared by the CPU cross-check (binder_ref.py vs
liboracle.so) and the GPU parity tests (CUDA path vs liboracle.so).

The shapes go well beyond the reference's own fixtures (test/*.test.js): every host-like
type, nested service.service, ttl on each level, ports lists, null / malformed / unknown /
non-object records, case-colliding names, duplicate addresses, unencodable labels."""
import json
import random
import struct

from binder_b200 import synth

HOSTLIKE = ['db_host', 'host', 'load_balancer', 'moray_host', 'redis_host', 'ops_host', 'rr_host']
LABELS = ['a', 'b', 'web', 'Web', 'lbA', 'lbB', 'lbC', 'h1', 'h2', 'H2', 'a-b', 'x_y', '_svc', 'moray',
          'ops', 'db', 'z9', 'UPPER', 'mixedCase', 'n0', 'n1', 'n2', 'n3', 'n4', 'n5', 'n6', 'n7',
          'l' * 63, 'k' * 40, '0', '1', '10', 'q.r', 'x' * 64, 'café']


def _ttl_variant(rng):
    r = rng.random()
    if r < 0.62:
        return None
    if r < 0.94:
        return rng.choice([0, 1, 45, 45, 60, 3600, 2 ** 31 - 1, 30.0, 1e2])
    return rng.choice([2 ** 31, -1, 1.5, '30', 'null', True, [30]])


def _addr_variant(rng):
    r = rng.random()
    if r < 0.93:
        return '%d.%d.%d.%d' % (rng.choice([10, 172, 192]), rng.randrange(256), rng.randrange(4), rng.randrange(6))
    return rng.choice(['999.1.1.1', '1.2.3', '01.2.3.4', '1.2.3.4.5', '', 'host.example', 5, 'null',
                       'MISSING', '1.2.3.256', ' 1.2.3.4', '1.2.3.4 ', '::1', True, [1]])


def _apply(d, key, v):
    if v is None:
        return
    d[key] = None if v == 'null' else v


def _hostlike(rng, types=HOSTLIKE):
    t = rng.choice(types)
    sub = {}
    a = _addr_variant(rng)
    if a != 'MISSING':
        sub['address'] = None if a == 'null' else a
    _apply(sub, 'ttl', _ttl_variant(rng))
    if rng.random() < 0.35:
        sub['ports'] = rng.choice([[80], [], [8080, 8081], [1, 2, 3, 4], [443.0], [0], [65535], [8080, 8081]]
                                  if rng.random() < 0.8 else ['80', None, 5, [70000], [1.5], {}, ['80']])
    rec = {'type': t, t: sub}
    _apply(rec, 'ttl', _ttl_variant(rng))
    r = rng.random()
    if r < 0.015:
        rec[t] = None
    elif r < 0.03:
        rec[t] = 'str'
    elif r < 0.04:
        del rec[t]
    elif r < 0.05:
        rec[t] = [1, 2]
    return rec


def _service(rng):
    s = {'srvce': rng.choice(['_http'] * 12 + ['_HTTP', '_pg', 5, None]),
         'proto': rng.choice(['_tcp'] * 12 + ['_udp', None])}
    p = rng.choice([80, 80, 443, 0, 65535, 80.0] * 3 + [65536, -1, 'MISSING', '80', 1.5])
    if p != 'MISSING':
        s['port'] = p
    _apply(s, 'ttl', _ttl_variant(rng))
    r = rng.random()
    if r < 0.15:
        outer = {'service': s}
        _apply(outer, 'ttl', _ttl_variant(rng))
        s = outer
    elif r < 0.165:
        s['service'] = None
    elif r < 0.18:
        s['service'] = 'x'
    rec = {'type': 'service', 'service': s}
    _apply(rec, 'ttl', _ttl_variant(rng))
    if rng.random() < 0.03:
        rec['service'] = None
    return rec


def _record(rng):
    """Returns ('data', value) or ('raw', str)."""
    r = rng.random()
    if r < 0.45:
        return 'data', _hostlike(rng)
    if r < 0.60:
        return 'data', _service(rng)
    if r < 0.66:
        return 'data', {'type': 'database', 'database': {'primary': rng.choice([
            'tcp://user@192.168.0.1/postgres', 'tcp://192.168.7.9:5432/db', 'tcp://u:p@10.1.2.3:5432',
            'TCP://10.9.8.7', 'tcp://user@host.example/postgres', 'tcp:10.0.0.1', '10.0.0.1', '',
            'tcp://[::1]/x', 'tcp://a@b@10.4.4.4/x', 5, None])}}
    if r < 0.72:
        return 'data', None
    if r < 0.75:
        return 'data', {'type': 'weird', 'weird': {'ttl': 5}}
    if r < 0.77:
        return 'data', {'type': 'weird'}
    if r < 0.79:
        return 'data', {'type': 7, 'host': {'address': '1.2.3.4'}}
    if r < 0.81:
        return 'data', [1, 2, 3]
    if r < 0.83:
        return 'data', {'host': {'address': '1.2.3.4'}}
    if r < 0.86:
        return 'raw', rng.choice(['not json', '5', '"str"', 'true', '{"type":"host"', '', '[',
                                  '{"type":"host","host":{"address":"10.77.0.1"},}', 'NaN'])
    if r < 0.90:
        return 'raw', json.dumps(_hostlike(rng))
    if r < 0.92:
        # duplicate keys / escapes: JSON.parse keeps the last duplicate
        return 'raw', '{"type":"host","host":{"address":"1.1.1.1"},"host":{"address":"10.\\u0031.2.3","ttl":1e2}}'
    return 'data', _hostlike(rng, ['load_balancer', 'rr_host', 'moray_host'])


def gen_zone(seed, dns_domain='foo.com', n_top=40):
    """-> (jsonl bytes, info) ; info = {'names': [...], 'addrs': [...], 'services': [...]}"""
    rng = random.Random(seed)
    root = '/' + '/'.join(reversed(dns_domain.split('.')))
    lines = []
    names, addrs, services = [], [], []

    def emit(path, kind, val):
        lines.append(json.dumps({'path': path, kind: val}))
        if kind == 'data' and isinstance(val, dict):
            t = val.get('type')
            if isinstance(t, str) and isinstance(val.get(t), dict):
                a = val[t].get('address')
                if isinstance(a, str):
                    addrs.append(a)

    if rng.random() < 0.8:
        emit(root, 'data', None)
    # something outside the watched subtree: must be ignored
    lines.append(json.dumps({'path': '/other/zone', 'data': {'type': 'host', 'host': {'address': '9.9.9.9'}}}))

    def subtree(path, dom, depth):
        used = set()
        for _ in range(rng.randrange(1, n_top if depth == 0 else 9)):
            lab = rng.choice(LABELS)
            if lab in used:
                continue
            used.add(lab)
            p = path + '/' + lab
            d = lab + '.' + dom
            kind, val = _record(rng)
            is_svc = kind == 'data' and isinstance(val, dict) and val.get('type') == 'service'
            emit(p, kind, val)
            names.append(d)
            if is_svc:
                services.append(d)
                kused = set()
                for _k in range(rng.choice([0, 1, 2, 3, 5, 8, 8, 12, 24])):
                    kl = rng.choice(LABELS)
                    if kl in kused:
                        continue
                    kused.add(kl)
                    kk, kv = _record(rng) if rng.random() < 0.3 else ('data', _hostlike(rng, ['load_balancer', 'rr_host', 'moray_host', 'ops_host', 'redis_host', 'host']))
                    emit(p + '/' + kl, kk, kv)
                    names.append(kl + '.' + d)
            elif depth < 2 and rng.random() < 0.3:
                subtree(p, d, depth + 1)

    subtree(root, dns_domain, 0)
    return ('\n'.join(lines) + '\n').encode('utf-8'), {'names': names, 'addrs': addrs, 'services': services,
                                                       'dns_domain': dns_domain}


def _mutate_case(rng, name):
    return ''.join(c.upper() if rng.random() < 0.3 else c.lower() if rng.random() < 0.3 else c for c in name)


def gen_queries(seed, info, n=400):
    """Well-formed queries around the zone's names."""
    rng = random.Random(seed ^ 0x5EED)
    dom = info['dns_domain']
    out = []
    names = [x for x in info['names'] if all(0 < len(l.encode('utf-8')) < 64 for l in x.split('.'))] or ['x.' + dom]
    for _ in range(n):
        r = rng.random()
        name = rng.choice(names)
        svc_pick = info['services'] and rng.random() < 0.3
        if svc_pick:
            name = rng.choice(info['services'])
        if r < 0.12:
            name = rng.choice(['nope', 'h999', 'a.b.c']) + '.' + dom
        elif r < 0.16:
            name = rng.choice([dom, 'foo.org', 'x.' + dom.upper(), 'x' + dom, '', 'com', 'a.b'])
        elif r < 0.40:
            name = _mutate_case(rng, name)
        t = rng.choice(['A', 'A', 'A', 'A', 'SRV', 'SRV', 'SRV', 'PTR', 'PTR', 'AAAA', 'TXT', 'ANY', 'NS'])
        if svc_pick:
            t = rng.choice(['A', 'SRV', 'SRV'])
        if t == 'SRV':
            q = rng.random()
            if q < 0.8:
                name = rng.choice(['_http._tcp.'] * 8 + ['_http._udp.', '_HTTP._tcp.', '_pg._tcp.']) + name
            elif q < 0.8:
                name = rng.choice(['_http.', '_ht_tp._tcp.', 'http._tcp.', '_http._tcp', '_._.', '_a._b.'])+ name
        if t == 'PTR':
            q = rng.random()
            if q < 0.6 and info['addrs']:
                a = rng.choice(info['addrs'])
                name = '.'.join(reversed(a.split('.'))) + '.in-addr.arpa'
            elif q < 0.75:
                name = rng.choice(['1.2.3.4.in-addr.arpa', 'in-addr.arpa', '1.2.in-addr.arpa', '4.3.2.1.IN-ADDR.ARPA',
                                   'arpa', '1.0.0.10.in-addr.arpa.x', '1.ip6.arpa'])
        try:
            labels = [l.encode('latin-1') for l in name.split('.')] if name else []
        except UnicodeEncodeError:
            labels = [l.encode('utf-8') for l in name.split('.')]
        if any(len(l) == 0 or len(l) > 63 for l in labels) or sum(len(l) + 1 for l in labels) + 1 > 255:
            continue
        if rng.random() < 0.04 and labels:          # in-label dot / control bytes / newline quirk
            i = rng.randrange(len(labels))
            labels[i] = rng.choice([b'a.b', b'x\ny', b'\x00z', b'sp ace', b'\xff\xfe', b'com\nzz', b'x\rfoo'])
        out.append(synth.make_query(None, t, qid=rng.randrange(65536), rd=rng.random() < 0.7,
                                    edns=rng.choice([None, None, None, 512, 1232, 4096, 100, 700]),
                                    opcode=0 if rng.random() < 0.985 else rng.choice([1, 2, 4, 5]),
                                    labels=labels))
    return out


def malformed_packets():
    """Packets the decoder must DROP (status 2, no bytes)."""
    good = synth.make_query('hosta.foo.com', 'A')
    edns = synth.make_query('hosta.foo.com', 'A', edns=4096)
    bad = [
        b'', b'\x00' * 11, good[:12], good[:20], good[:-1], good[:-4],
        good[:2] + bytes([good[2] | 0x80]) + good[3:],                       # QR=1
        good[:4] + b'\x00\x02' + good[6:],                                   # QDCOUNT=2
        good[:4] + b'\x00\x00' + good[6:],                                   # QDCOUNT=0
        good[:6] + b'\x00\x01' + good[8:],                                   # ANCOUNT=1
        good[:8] + b'\x00\x01' + good[10:],                                  # NSCOUNT=1
        good[:10] + b'\x00\x02' + good[12:],                                 # ARCOUNT=2
        good[:10] + b'\x00\x01' + good[12:],                                 # ARCOUNT=1 but no OPT bytes
        good[:12] + b'\xc0\x0c' + good[-4:],                                 # compression pointer as QNAME
        good[:12] + b'\x40' + b'a' * 64 + b'\x00' + good[-4:],               # label length 64
        good[:12] + b'\x05ab',                                               # label runs past the end
        good[:-2] + b'\x00\x03',                                             # class CH
        good[:12] + (b'\x3f' + b'a' * 63) * 4 + b'\x00' + good[-4:],         # name > 255
        edns[:-11] + b'\x01' + edns[-10:],                                   # OPT owner not root
        edns[:-10] + b'\x00\x10' + edns[-8:],                                # additional RR is not OPT
        edns[:-2] + b'\x00\x09',                                             # OPT rdlen beyond packet
    ]
    return bad


def tolerated_packets():
    """Odd-but-accepted packets: trailing junk, OPT with options, max-length name."""
    good = synth.make_query('hosta.foo.com', 'A')
    edns = synth.make_query('hosta.foo.com', 'A', edns=4096)
    longname = synth.make_query(None, 'A', labels=[b'a' * 63, b'b' * 63, b'c' * 63, b'd' * 61])
    return [
        good + b'\xde\xad\xbe\xef',
        edns[:-2] + b'\x00\x0c' + b'\x00\x0a\x00\x08' + b'\x11' * 8,        # OPT carrying a COOKIE option
        longname,
        synth.make_query('', 'A'), synth.make_query('', 'PTR'), synth.make_query('', 'SRV'),
        struct.pack('>HHHHHH', 1, 0x0100, 1, 0, 0, 0) + b'\x00' + struct.pack('>HH', 1, 1),
    ]


# ---------------------------------------------------------------------------------------
# deltas: the watch events of lib/zk.js:120-208 as JSON lines (bb_zone_apply / orc_apply_delta)
# ---------------------------------------------------------------------------------------
def snapshot_paths(snap):
    """Paths of a snapshot's lines, in order."""
    return [json.loads(l)['path'] for l in snap.decode('utf-8').split('\n') if l.strip()]


def gen_delta(seed, paths, info, n_ops=40):
    """A batch of watch events around the zone's znodes -> (jsonl bytes, paths after it).  Updates `info`
    (names / addrs / services) so that gen_queries also asks about what the delta touched.

    Covers: data changes of every record shape (type changes included), address changes that reuse another
    node's address (the reverse map's last-writer / stale-entry behaviour), new children (case twins of
    existing ones included), deletions of leaves and subtrees, re-creation of deleted paths, events outside
    the mirrored subtree, unparsable content (ignored: previous data stays)."""
    rng = random.Random(seed ^ 0xDE17A)
    dom = info['dns_domain']
    root = '/' + '/'.join(reversed(dom.split('.')))
    live = [p for p in paths if p == root or p.startswith(root + '/')]
    lines = []

    def dom_of(path):
        return '.'.join(reversed(path[len(root) + 1:].split('/'))) + '.' + dom if path != root else dom

    def note(path, kind, val):
        if kind == 'data' and isinstance(val, dict):
            t = val.get('type')
            if isinstance(t, str) and isinstance(val.get(t), dict) and isinstance(val[t].get('address'), str):
                info['addrs'].append(val[t]['address'])
            if t == 'service':
                info['services'].append(dom_of(path))
        info['names'].append(dom_of(path))

    for _ in range(n_ops):
        r = rng.random()
        if r < 0.40 and live:                           # dataChanged: any shape
            p = rng.choice(live)
            kind, val = _record(rng)
            lines.append(json.dumps({'path': p, kind: val})); note(p, kind, val)
        elif r < 0.58 and live:                         # address change, often onto an address in use
            p = rng.choice(live)
            val = _hostlike(rng, ['host', 'load_balancer', 'rr_host', 'redis_host'])
            t = val['type']
            if isinstance(val.get(t), dict) and info['addrs'] and rng.random() < 0.6:
                val[t]['address'] = rng.choice(info['addrs'])
            lines.append(json.dumps({'path': p, 'data': val})); note(p, 'data', val)
        elif r < 0.76 and live:                         # childrenChanged: a new child (maybe a case twin)
            parent = rng.choice(live)
            p = parent + '/' + rng.choice(LABELS)
            kind, val = _record(rng) if rng.random() < 0.5 else ('data', _hostlike(rng, ['load_balancer', 'rr_host', 'moray_host', 'host']))
            lines.append(json.dumps({'path': p, kind: val})); note(p, kind, val)
            if p not in live:
                live.append(p)
        elif r < 0.93 and len(live) > 1:                # childrenChanged: a child (and its subtree) is gone
            p = rng.choice([x for x in live if x != root])
            lines.append(json.dumps({'path': p, 'deleted': True}))
            live = [x for x in live if x != p and not x.startswith(p + '/')]
        elif r < 0.96:
            lines.append(json.dumps({'path': '/other/zone/x', 'data': {'type': 'host', 'host': {'address': '8.8.8.8'}}}))
        else:                                           # deleting what is not there; content for a child of nothing
            lines.append(json.dumps({'path': root + '/ghost/child', 'data': None}))
            lines.append(json.dumps({'path': root + '/ghost', 'deleted': True}))
    return ('\n'.join(lines) + '\n').encode('utf-8'), live
