"""BASELINE config 1 (plumbing, no GPU): a ~100-record snapshot containing the reference's golden
zones, 1k mixed A/SRV/PTR queries over UDP loopback, client = dnspython.  The reference runs this
with test/dig.js against Node; node/dig/ZooKeeper are absent here, so the same queries go to
binder_b200.server (the host-side mirror) with the CPU oracle injected as its resolver; the GPU
variant injects the real Engine."""
import json
import os
import random

import dns.message
import dns.query
import dns.rcode
import dns.rdatatype
import pytest

import helpers as H
from binder_b200.server import createServer

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_cases.json')))


def _snapshot():
    ent, seen = [], set()
    for suite in GOLD['suites']:
        for path, data in suite['snapshot']:
            if path not in seen:
                seen.add(path); ent.append((path, data))
    for i in range(80):
        ent.append(('/com/foo/h%02d' % i, {'type': 'host', 'host': {'address': '10.1.%d.%d' % (i // 8, i)}}))
    return H.snapshot(ent)


def _run(kind):
    snap = _snapshot()
    impl = H.make_impl(kind, 'foo.com', snap)
    srv = createServer({'dnsDomain': 'foo.com', 'resolver': impl, 'host': '127.0.0.1', 'port': 0}).start()
    try:
        rng = random.Random(1)
        cases = [c for s in GOLD['suites'] for c in s['cases']]
        n_ok = 0
        for i in range(1000):
            if i % 4 == 0:
                c = rng.choice(cases)
                q = dns.message.make_query(c['name'], c['type'], use_edns=rng.random() < 0.5)
                r = dns.query.udp(q, '127.0.0.1', port=srv.port, timeout=2)
                if 'status' in c:
                    assert dns.rcode.to_text(r.rcode()) == c['status']
                assert sum(len(rr) for rr in r.answer) == len(c['answers'])
            else:
                k = rng.randrange(80)
                q = dns.message.make_query('h%02d.foo.com' % k, 'A')
                r = dns.query.udp(q, '127.0.0.1', port=srv.port, timeout=2)
                assert r.rcode() == 0 and r.answer[0][0].address == '10.1.%d.%d' % (k // 8, k) and r.answer[0].ttl == 30
            n_ok += 1
        assert n_ok == 1000 and srv.counters['answered'] == 1000 and srv.counters['dropped'] == 0
    finally:
        srv.stop()


def test_config1_udp_loopback_cpu_oracle():
    _run('oracle')


@pytest.mark.gpu
def test_config1_udp_loopback_gpu_engine():
    _run('gpu')
