"""Two independent restatements of lib/server.js + lib/zk.js must agree: the C++ oracle's
bytes, decoded by dnspython, against oracle/binder_ref.py, over adversarial zones."""
import struct

import pytest

import fuzzgen
import helpers as H
from binder_b200 import synth


def check_against_ref(impl, snap, info, pkts, recursion, seed):
    opts = H.ref_options(snap, info['dns_domain'], recursion=recursion)
    res, miss = H.resolve_list(impl, pkts, seed=seed, qidx_base=11)
    want_miss = []
    for i, (pkt, (st, wire)) in enumerate(zip(pkts, res)):
        ref = H.ref_semantic(opts, pkt, seed=seed, qidx=11 + i)
        labels, qtype, rd, opcode, edns = H.split_query(pkt)
        assert st == ref[0], (i, pkt)
        if st != 0:
            assert wire == b''
            want_miss.append(i)
            continue
        if opcode != 0:
            # dnspython parses non-QUERY opcodes with other grammars (UPDATE): check by hand
            assert (wire[2] >> 3) & 0xF == opcode and wire[3] == 4 and ref[1] == 4
            assert struct.unpack('>HHHH', wire[4:12]) == (1, 0, 0, 1 if edns else 0)
            continue
        rcode, answers, authority, additional, f = H.decode_semantic(wire)
        adv = struct.unpack('>H', pkt[-8:-6])[0] if edns else 0
        maxsz = min(max(adv, 512), 1200) if edns else 512
        assert len(wire) <= maxsz
        assert wire[:2] == pkt[:2] and f['qr'] and f['aa'] and not f['ra'] and f['rd'] == rd
        assert f['edns'] == edns and (not edns or f['payload'] == 1200)
        qend = 12 + sum(len(l) + 1 for l in labels) + 1 + 4
        assert wire[12:qend] == pkt[12:qend]                  # question echoed verbatim
        assert rcode == ref[1], (i, pkt, rcode, ref)
        got = answers + authority + additional
        want = ref[2] + ref[3] + ref[4]
        if f['tc']:
            assert len(got) < len(want) and got == want[:len(got)]
            assert answers == ref[2][:len(answers)]
        else:
            assert (answers, authority, additional) == ref[2:], (i, pkt)
    assert miss == want_miss


@pytest.mark.parametrize('seed', range(24))
def test_cpp_oracle_matches_semantic_restatement(seed):
    snap, info = fuzzgen.gen_zone(seed, n_top=30)
    recursion = seed % 3 == 0
    impl = H.make_impl('oracle', info['dns_domain'], snap, recursion=recursion)
    pkts = fuzzgen.gen_queries(seed, info, n=500)
    check_against_ref(impl, snap, info, pkts, recursion, seed=seed * 7919 + 1)


def test_malformed_packets_are_dropped():
    snap = H.snapshot([('/com/foo', None), ('/com/foo/hosta', {'type': 'host', 'host': {'address': '192.168.0.1'}})])
    impl = H.make_impl('oracle', 'foo.com', snap)
    bad = fuzzgen.malformed_packets()
    res, miss = H.resolve_list(impl, bad)
    assert [r for r in res] == [(2, b'')] * len(bad) and miss == []
    res, miss = H.resolve_list(impl, fuzzgen.tolerated_packets())
    assert all(st == 0 and len(w) >= 17 for st, w in res)


def test_not_ready_is_servfail():
    """lib/server.js:186-192, 86-92: no ZK session yet -> SERVFAIL, but only after the
    suffix / arpa refusals."""
    impl = H.make_impl('oracle', 'foo.com', None)
    pk = [synth.make_query('hosta.foo.com', 'A'), synth.make_query('1.0.0.10.in-addr.arpa', 'PTR'),
          synth.make_query('hosta.bar.org', 'A'), synth.make_query('x.y', 'PTR'),
          synth.make_query('_http._tcp.s.foo.com', 'SRV'), synth.make_query('hosta.foo.com', 'AAAA')]
    res, _ = H.resolve_list(impl, pk)
    assert [H.decode_semantic(w)[0] for _, w in res] == [2, 2, 5, 5, 2, 4]


def test_truncation_keeps_longest_prefix():
    kids = [('/com/foo/svc/lb%02d' % i, {'type': 'load_balancer', 'load_balancer': {'address': '10.0.0.%d' % i}})
            for i in range(40)]
    snap = H.snapshot([('/com/foo', None), ('/com/foo/svc', {'type': 'service', 'service': {
        'srvce': '_http', 'proto': '_tcp', 'port': 80}})] + kids)
    impl = H.make_impl('oracle', 'foo.com', snap)
    opts = H.ref_options(snap, 'foo.com')
    for edns, maxsz in ((None, 512), (4096, 1200), (600, 600), (100, 512)):
        for name, t in (('_http._tcp.svc.foo.com', 'SRV'), ('svc.foo.com', 'A')):
            pkt = synth.make_query(name, t, edns=edns)
            (res,), _ = H.resolve_list(impl, [pkt], seed=5)
            rcode, an, au, ad, f = H.decode_semantic(res[1])
            ref = H.ref_semantic(opts, pkt, seed=5, qidx=0)
            want = ref[2] + ref[3] + ref[4]
            got = an + au + ad
            assert got == want[:len(got)] and len(res[1]) <= maxsz
            assert f['tc'] == (len(got) < len(want))
            if f['tc']:
                # maximal: the next RR would not have fitted (A RR with a pointer owner = 16 B,
                # SRV RR here = 2+10+6+16 = 34 B, additional A with 5-byte literal = 21 B)
                nxt = want[len(got)]
                size = 16 if (nxt[2] == 'A' and nxt[0] == name) else 34 if nxt[2] == 'SRV' else 21
                assert len(res[1]) + size > maxsz
