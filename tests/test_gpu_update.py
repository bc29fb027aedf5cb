"""Incremental zone updates on the device: after bb_zone_apply +
bb_engine_apply_update — which ships only the changed slots and the arena tail — the CUDA path must stay
byte-identical to the CPU oracle that applied the same watch events."""
import json

import pytest

import fuzzgen
import helpers as H
from binder_b200 import synth
from test_gpu_parity import assert_same

pytestmark = pytest.mark.gpu


def ptr_queries(info):
    out = []
    for a in sorted(set(x for x in info['addrs'] if x)):
        labels = [l.encode('utf-8') for l in reversed(a.split('.'))] + [b'in-addr', b'arpa']
        if all(0 < len(l) < 64 for l in labels) and sum(len(l) + 1 for l in labels) < 255:
            out.append(synth.make_query(None, 'PTR', labels=labels))
    return out


@pytest.mark.parametrize('seed', range(8))
def test_updates_stay_bit_exact(seed):
    from binder_b200.engine import Engine, Zone
    snap, info = fuzzgen.gen_zone(seed + 300, n_top=30)
    dom = info['dns_domain']
    recursion = seed % 2 == 0
    zone = Zone(snap, dom)
    gpu = Engine(dom, recursion=recursion, ordered=seed % 3 == 0)
    gpu.swap_zone(zone)
    orc = H.make_impl('oracle', dom, snap, recursion=recursion)
    paths = fuzzgen.snapshot_paths(snap)
    incremental = 0
    for rnd in range(6):
        delta, paths = fuzzgen.gen_delta(seed * 100 + rnd, paths, info, n_ops=50)
        zone.apply(delta)
        n_dirty, relaid = zone.pending()
        gpu.apply_update(zone)
        assert zone.pending() == (0, False)
        incremental += (not relaid) and n_dirty > 0
        orc.apply_delta(delta)
        pkts = fuzzgen.gen_queries(seed * 17 + rnd, info, n=2000) + ptr_queries(info)
        data, off = synth.pack_batch(pkts)
        assert_same(gpu, orc, data, off, seed=seed * 7919 + rnd, qidx_base=rnd * 5000)
    assert incremental >= 4          # the slot-patch path, not full swaps, carried the updates


def test_update_through_relayout_and_second_engine():
    """Growth past the table's load limit re-lays the table (full upload); an engine that did not take the
    zone's previous changes gets a full upload too, never a partial patch."""
    from binder_b200.engine import Engine, Zone
    snap, info = fuzzgen.gen_zone(777, n_top=20)
    dom = info['dns_domain']
    zone = Zone(snap, dom)
    a, b = Engine(dom), Engine(dom)
    a.swap_zone(zone)
    orc = H.make_impl('oracle', dom, snap)
    root = '/' + '/'.join(reversed(dom.split('.')))
    grow = '\n'.join(json.dumps({'path': '%s/g%05d' % (root, i), 'data': {'type': 'host', 'host': {'address': '198.51.%d.%d' % (i >> 8, i & 255)}}})
                     for i in range(3000)).encode()
    slots0 = zone.stat()['slots']
    zone.apply(grow); orc.apply_delta(grow)
    assert zone.stat()['slots'] > slots0 and zone.pending()[1]
    a.apply_update(zone)
    small = json.dumps({'path': root + '/g00007', 'data': {'type': 'host', 'host': {'address': '198.51.100.200'}, 'ttl': 7}}).encode()
    zone.apply(small); orc.apply_delta(small)
    b.apply_update(zone)                          # b never saw this zone: full upload
    zone.apply(json.dumps({'path': root + '/g00008', 'deleted': True}).encode()); orc.apply_delta(json.dumps({'path': root + '/g00008', 'deleted': True}).encode())
    a.apply_update(zone)                          # a is one sync behind b: full upload again, not a stale patch
    pkts = [synth.make_query('g%05d.%s' % (i, dom), 'A') for i in range(0, 3000, 7)] + \
           [synth.make_query('g00007.' + dom, 'A'), synth.make_query('g00008.' + dom, 'A'),
            synth.make_query('200.100.51.198.in-addr.arpa', 'PTR'), synth.make_query('7.0.51.198.in-addr.arpa', 'PTR'),
            synth.make_query('8.0.51.198.in-addr.arpa', 'PTR')]
    data, off = synth.pack_batch(pkts)
    assert_same(a, orc, data, off, seed=1)
    b.apply_update(zone)
    assert_same(b, orc, data, off, seed=1)
