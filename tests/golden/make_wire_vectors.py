"""Generates tests/golden/wire_vectors.json: query -> response bytes (hex) produced by the CPU
oracle for a fixed zone and seed.  Committed so that any later change to the wire spec (oracle or
kernel) shows up as a diff of golden bytes.  Run:  python tests/golden/make_wire_vectors.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')]
import helpers as H
from binder_b200 import synth
from oracle_lib import Oracle

GOLD = json.load(open(os.path.join(HERE, 'reference_cases.json')))
ZONE = [(p, d) for s in GOLD['suites'] for p, d in s['snapshot'] if not (p == '/com/foo' and s is not GOLD['suites'][0])] + [
    ('/com/foo/rr', {'type': 'service', 'service': {'srvce': '_pg', 'proto': '_tcp', 'port': 5432, 'ttl': 45}}),
    ('/com/foo/rr/n1', {'type': 'rr_host', 'rr_host': {'address': '10.9.0.1', 'ports': [5432, 5433]}, 'ttl': 20}),
    ('/com/foo/rr/n2', {'type': 'moray_host', 'moray_host': {'address': '10.9.0.2', 'ttl': 90}}),
    ('/com/foo/rr/dead', {'type': 'redis_host', 'redis_host': {'address': None}}),
    ('/com/foo/null', None),
    ('/com/foo/weird', {'type': 'weird', 'weird': {}}),
]
QUERIES = [(c['name'], c['type'], None) for s in GOLD['suites'] for c in s['cases']] + [
    ('_pg._tcp.rr.foo.com', 'SRV', None), ('_pg._tcp.rr.foo.com', 'SRV', 4096), ('rr.foo.com', 'A', None),
    ('_PG._tcp.rr.foo.com', 'SRV', None), ('_pg._tcp.RR.Foo.com', 'SRV', None), ('HostA.foo.com', 'A', 1232),
    ('null.foo.com', 'A', None), ('weird.foo.com', 'A', None), ('_x._y.weird.foo.com', 'SRV', None),
    ('hosta.foo.com', 'AAAA', None), ('hosta.foo.com', 'TXT', 512), ('foo.com', 'A', None), ('hosta.FOO.com', 'A', None),
]


def main():
    snap = H.snapshot(ZONE)
    orc = Oracle('foo.com', '', False, snapshot=snap)
    vec = []
    for i, (name, qt, edns) in enumerate(QUERIES):
        pkt = synth.make_query(name, qt, qid=0x1000 + i, edns=edns)
        out, st = orc.resolve_one(pkt, seed=0x5EED, qidx=i)
        vec.append({'name': name, 'type': qt, 'edns': edns, 'query': pkt.hex(), 'status': st, 'response': out.hex()})
    json.dump({'dns_domain': 'foo.com', 'seed': 0x5EED, 'zone': ZONE, 'vectors': vec},
              open(os.path.join(HERE, 'wire_vectors.json'), 'w'), indent=1)
    print('wrote %d vectors' % len(vec))


if __name__ == '__main__':
    main()
