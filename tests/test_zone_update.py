"""Incremental zone updates (lib/zk.js:120-208 watch events as deltas).
CPU side: the two oracles agree on what a delta does; the product's zone builder re-derives exactly what a
fresh build of the resulting tree holds (bb_zone_probe), including across a forced re-layout."""
import json

import pytest

import fuzzgen
import helpers as H
from test_oracle_cross import check_against_ref


class RefAfterDelta(object):
    """H.ref_options, with deltas applied to its cache."""
    def __init__(self, snap, info, recursion):
        self.opts = H.ref_options(snap, info['dns_domain'], recursion=recursion)

    def apply(self, delta):
        self.opts.zkCache.apply_delta(delta.decode('utf-8').split('\n'))


@pytest.mark.parametrize('seed', range(12))
def test_oracles_agree_after_deltas(seed, monkeypatch):
    snap, info = fuzzgen.gen_zone(seed, n_top=25)
    recursion = seed % 3 == 0
    impl = H.make_impl('oracle', info['dns_domain'], snap, recursion=recursion)
    ref = RefAfterDelta(snap, info, recursion)
    monkeypatch.setattr(H, 'ref_options', lambda *a, **k: ref.opts)
    paths = fuzzgen.snapshot_paths(snap)
    for rnd in range(4):
        delta, paths = fuzzgen.gen_delta(seed * 100 + rnd, paths, info, n_ops=30)
        impl.apply_delta(delta)
        ref.apply(delta)
        pkts = fuzzgen.gen_queries(seed * 31 + rnd, info, n=300)
        check_against_ref(impl, snap, info, pkts, recursion, seed=seed * 7919 + rnd)


ZONE0 = [
    ('/com/foo', None),
    ('/com/foo/hosta', {'type': 'host', 'host': {'address': '192.168.0.1'}}),
    ('/com/foo/hostb', {'type': 'host', 'host': {'address': '192.168.0.2'}, 'ttl': 60}),
    ('/com/foo/svc', {'type': 'service', 'service': {'srvce': '_http', 'proto': '_tcp', 'port': 80}}),
    ('/com/foo/svc/lb0', {'type': 'load_balancer', 'load_balancer': {'address': '10.0.0.1'}}),
    ('/com/foo/svc/lb1', {'type': 'load_balancer', 'load_balancer': {'address': '10.0.0.2'}}),
    ('/com/foo/grp', None),
    ('/com/foo/grp/x', {'type': 'host', 'host': {'address': '172.16.0.1'}}),
    ('/com/foo/grp/y', {'type': 'host', 'host': {'address': '172.16.0.2'}}),
]
DELTA = [
    {'path': '/com/foo/hosta', 'data': {'type': 'host', 'host': {'address': '192.168.0.9'}}},        # address change
    {'path': '/com/foo/hostb', 'data': {'type': 'host', 'host': {'address': '192.168.0.2', 'ttl': 5}}},  # ttl change
    {'path': '/com/foo/svc/lb2', 'data': {'type': 'rr_host', 'rr_host': {'address': '10.0.0.3', 'ports': [8080, 8081]}}},
    {'path': '/com/foo/svc/lb0', 'deleted': True},
    {'path': '/com/foo/newhost', 'data': {'type': 'host', 'host': {'address': '192.168.0.77'}}},
    {'path': '/com/foo/grp', 'deleted': True},
    {'path': '/com/foo/hosta', 'raw': 'not json'},                                                      # ignored
]
ZONE1 = [
    ('/com/foo', None),
    ('/com/foo/hosta', {'type': 'host', 'host': {'address': '192.168.0.9'}}),
    ('/com/foo/hostb', {'type': 'host', 'host': {'address': '192.168.0.2', 'ttl': 5}}),
    ('/com/foo/svc', {'type': 'service', 'service': {'srvce': '_http', 'proto': '_tcp', 'port': 80}}),
    ('/com/foo/svc/lb1', {'type': 'load_balancer', 'load_balancer': {'address': '10.0.0.2'}}),
    ('/com/foo/svc/lb2', {'type': 'rr_host', 'rr_host': {'address': '10.0.0.3', 'ports': [8080, 8081]}}),
    ('/com/foo/newhost', {'type': 'host', 'host': {'address': '192.168.0.77'}}),
]
FWD = ['foo.com', 'hosta.foo.com', 'hostb.foo.com', 'svc.foo.com', 'lb0.svc.foo.com', 'lb1.svc.foo.com', 'lb2.svc.foo.com',
       'grp.foo.com', 'x.grp.foo.com', 'y.grp.foo.com', 'newhost.foo.com', 'nope.foo.com']
# an unbound node keeps its reverse entry (lib/zk.js never removes it), so the deleted znodes' addresses are
# compared separately below
REV = ['192.168.0.1', '192.168.0.9', '192.168.0.2', '10.0.0.2', '10.0.0.3', '192.168.0.77', '1.2.3.4']
REV_OF_DELETED = {'10.0.0.1': 'lb0.svc.foo.com', '172.16.0.1': 'x.grp.foo.com', '172.16.0.2': 'y.grp.foo.com'}


def wire(name):
    return b''.join(bytes([len(l)]) + l.encode() for l in name.split('.')) + b'\0'


def test_incremental_equals_fresh_build():
    from binder_b200.engine import Zone
    z = Zone(H.snapshot(ZONE0), 'foo.com')
    assert z.pending() == (0, True)            # a fresh build has never been uploaded: full upload pending
    z.apply('\n'.join(json.dumps(d) for d in DELTA))
    fresh = Zone(H.snapshot(ZONE1), 'foo.com')
    for k in FWD:
        assert z.probe(k) == fresh.probe(k), k
    for k in REV:
        assert z.probe(k, reverse=True) == fresh.probe(k, reverse=True), k
    for addr, dom in REV_OF_DELETED.items():
        kind, ttl, _, rec = z.probe(addr, reverse=True)
        assert (kind, ttl, rec) == (6, 30, wire(dom)), addr       # K_PTR, still answering
        assert fresh.probe(addr, reverse=True) is None
    assert z.probe('lb0.svc.foo.com') is None and z.probe('grp.foo.com') is None and z.probe('x.grp.foo.com') is None
    st, fs = z.stat(), fresh.stat()
    assert st['forward_keys'] == fs['forward_keys'] and st['reverse_keys'] == fs['reverse_keys'] + 3


def test_growth_forces_relayout_and_keeps_every_key():
    from binder_b200.engine import Zone
    z = Zone(H.snapshot(ZONE0), 'foo.com')
    slots0 = z.stat()['slots']
    new = [{'path': '/com/foo/n%04d' % i, 'data': {'type': 'host', 'host': {'address': '10.9.%d.%d' % (i >> 8, i & 255)}}}
           for i in range(600)]
    z.apply('\n'.join(json.dumps(d) for d in new))
    assert z.stat()['slots'] > slots0 and z.pending()[1]
    fresh = Zone(H.snapshot(ZONE0 + [(d['path'], d['data']) for d in new]), 'foo.com')
    for i in range(600):
        assert z.probe('n%04d.foo.com' % i) == fresh.probe('n%04d.foo.com' % i)
        a = '10.9.%d.%d' % (i >> 8, i & 255)
        assert z.probe(a, reverse=True) == fresh.probe(a, reverse=True)
    for k in FWD:
        assert z.probe(k) == fresh.probe(k), k


def test_bad_delta_line_is_an_error():
    from binder_b200.engine import Zone
    from binder_b200._lib import BinderError
    z = Zone(H.snapshot(ZONE0), 'foo.com')
    with pytest.raises(BinderError):
        z.apply('{"data": null}\n')
    with pytest.raises(BinderError):
        z.apply('{"path": "/com/foo", "deleted": true}\n')
    with pytest.raises(BinderError):
        Zone(H.snapshot(ZONE0) + b'{"path": "/com/foo/hosta", "deleted": true}\n', 'foo.com')


@pytest.mark.parametrize('seed', range(10))
def test_incremental_state_survives_a_relayout(seed):
    """What a sequence of deltas left in the table (incrementally maintained) must equal what a full
    re-derivation from the tree produces: forcing a re-layout must not change any existing key's answer."""
    from binder_b200.engine import Zone
    snap, info = fuzzgen.gen_zone(seed + 50, n_top=25)
    z = Zone(snap, info['dns_domain'])
    paths = fuzzgen.snapshot_paths(snap)
    for rnd in range(5):
        delta, paths = fuzzgen.gen_delta(seed * 100 + rnd, paths, info, n_ops=40)
        z.apply(delta)
    lower = lambda s: ''.join(chr(ord(c) + 32) if 'A' <= c <= 'Z' else c for c in s)
    fkeys = sorted({lower(n).encode('utf-8') for n in info['names']})
    rkeys = sorted({a.encode('utf-8') for a in info['addrs'] if a})
    before = [z.probe(k) for k in fkeys] + [z.probe(k, reverse=True) for k in rkeys]
    assert any(b is not None for b in before)
    slots0 = z.stat()['slots']
    root = '/' + '/'.join(reversed(info['dns_domain'].split('.')))
    grow = [{'path': '%s/grow%05d' % (root, i), 'data': {'type': 'host', 'host': {'address': '203.0.%d.%d' % (i >> 8, i & 255)}}}
            for i in range(max(400, slots0 // 3))]
    z.apply('\n'.join(json.dumps(d) for d in grow))
    assert z.stat()['slots'] > slots0
    after = [z.probe(k) for k in fkeys] + [z.probe(k, reverse=True) for k in rkeys]
    assert before == after


def test_arena_garbage_triggers_compaction():
    """Re-deriving a service record appends to the arena; churn on one service must not grow it without bound."""
    from binder_b200.engine import Zone
    kids = [('/com/foo/svc/k%03d' % i, {'type': 'load_balancer', 'load_balancer': {'address': '10.5.0.%d' % i}}) for i in range(200)]
    z = Zone(H.snapshot(ZONE0 + kids), 'foo.com')
    a0 = z.stat()['arena_bytes']
    relaid = 0
    for rnd in range(6000):                       # each event re-emits the ~7 KB record of /com/foo/svc
        z.apply(json.dumps({'path': '/com/foo/svc/k%03d' % (rnd % 200), 'data': {'type': 'load_balancer', 'load_balancer': {'address': '10.6.%d.%d' % (rnd >> 8, rnd & 255)}}}))
        relaid += z.pending()[1]
    assert z.stat()['arena_bytes'] < 2 * a0 + (17 << 20) and relaid > 0
    kind, ttl, _, rec = z.probe('svc.foo.com')
    fresh = Zone(H.snapshot(ZONE0 + [(p, {'type': 'load_balancer', 'load_balancer': {'address': '10.6.%d.%d' % ((5800 + i) >> 8, (5800 + i) & 255)}})
                                     for i, (p, _) in enumerate(kids)]), 'foo.com')
    assert (kind, ttl, rec) == (fresh.probe('svc.foo.com')[0], fresh.probe('svc.foo.com')[1], fresh.probe('svc.foo.com')[3])


@pytest.mark.parametrize('seed', range(4))
def test_sharded_zones_take_the_same_delta(seed):
    """Every rank applies the same events and keeps only its own keys: after the deltas each key lives on
    exactly one shard, with the payload the unsharded zone holds (arena offsets aside)."""
    from binder_b200.engine import Zone
    from binder_b200.shard import hash_keys, owner_of
    snap, info = fuzzgen.gen_zone(seed + 80, n_top=25)
    dom = info['dns_domain']
    whole = Zone(snap, dom)
    shards = [Zone(snap, dom, nranks=3, rank=r) for r in range(3)]
    paths = fuzzgen.snapshot_paths(snap)
    for rnd in range(4):
        delta, paths = fuzzgen.gen_delta(seed * 100 + rnd, paths, info, n_ops=40)
        for z in [whole] + shards:
            z.apply(delta)
    lower = lambda s: ''.join(chr(ord(c) + 32) if 'A' <= c <= 'Z' else c for c in s)
    fkeys = sorted({lower(n).encode('utf-8') for n in info['names']})
    rkeys = sorted({a.encode('utf-8') for a in info['addrs'] if a})
    n_present = 0
    for keys, rev in ((fkeys, False), (rkeys, True)):
        from binder_b200.shard import canon_forward
        # hash_keys wants equal lengths; a forward name no query can spell is stored nowhere (owner -1)
        owners = [-1 if (not rev and canon_forward(k, dom) is None) else int(owner_of(hash_keys([k], 1 if rev else 0, dom), 3)[0]) for k in keys]
        for k, own in zip(keys, owners):
            want = whole.probe(k, reverse=rev)
            got = [z.probe(k, reverse=rev) for z in shards]
            for r in range(3):
                assert got[r] == (want if r == int(own) else None), (k, rev, r, int(own))
            n_present += want is not None
    assert n_present > 20
    assert sum(z.stat()['forward_keys'] for z in shards) == whole.stat()['forward_keys']
    assert sum(z.stat()['reverse_keys'] for z in shards) == whole.stat()['reverse_keys']


def test_failed_delta_leaves_a_usable_image():
    """A delta whose LAST line is junk returns BB_ERR_SNAPSHOT with the earlier lines applied (include/binder_b200.h).
    Those lines grew the arena (a service record per new child): the image must point at the builder's current arena,
    not at the freed one — probing the zone and building fresh from the same events must agree."""
    from binder_b200.engine import Zone
    from binder_b200._lib import BinderError
    z = Zone(H.snapshot(ZONE0), 'foo.com')
    kids = [('/com/foo/svc/extra%04d' % i, {'type': 'load_balancer', 'load_balancer': {'address': '10.7.%d.%d' % (i >> 8, i & 255)}}) for i in range(2000)]
    delta = H.snapshot(kids) + b'this is not json\n'
    with pytest.raises(BinderError):
        z.apply(delta)
    fresh = Zone(H.snapshot(ZONE0 + kids), 'foo.com')
    assert z.probe('svc.foo.com') == fresh.probe('svc.foo.com')
    assert z.probe('extra1999.svc.foo.com') == fresh.probe('extra1999.svc.foo.com')
    assert z.stat()['arena_bytes'] >= fresh.stat()['arena_bytes']
    z.apply(H.snapshot([('/com/foo/svc/extra0000', {'type': 'load_balancer', 'load_balancer': {'address': '10.9.9.9'}})]))   # and it keeps working
    assert z.probe('10.9.9.9', reverse=True) is not None
