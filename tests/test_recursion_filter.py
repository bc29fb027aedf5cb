"""Recursion pre-filter (lib/recursion.js:329-344): misses that Recursion.resolve()
would refuse without asking anyone are answered REFUSED instead of entering the miss list."""
import random

import pytest

import fuzzgen
import helpers as H
from binder_b200 import synth

FILTERS = [
    ('foo.com', ['web', 'h1', 'nope', 'c', 'UPPER', 'db'], True),     # region = dnsDomain (binder's real configuration)
    ('foo.com', [], False),                                            # nowhere to go at all
    ('com', ['foo'], True),                                            # dc is always 'foo'
    ('com', ['bar'], False),
    ('oo.com', ['web', 'f', ''.join(['x'] * 63)], True),              # suffix not on a label boundary: dc = ''
    ('bar.org', ['web'], True),                                        # not our suffix
    ('a.very.long.region.domain.that.is.longer.than.any.query.name.foo.com', ['web'], False),
    ('.foo.com', ['web', 'b'], True),                                  # leading dot: drops one more character
]


def filter_queries(seed, info):
    rng = random.Random(seed)
    dom = info['dns_domain']
    out = fuzzgen.gen_queries(seed, info, n=600)
    tops = ['web', 'Web', 'WEB', 'h1', 'nope', 'c', 'UPPER', 'upper', 'db', 'zz', 'a-b', 'x_y']
    for _ in range(500):
        top = rng.choice(tops)
        pre = rng.choice(['', 'zzz.', 'a.b.', 'Q.', '_http._tcp.', 'h999.n0.'])
        t = 'SRV' if pre.startswith('_') else rng.choice(['A', 'A', 'SRV'])
        name = pre + top + '.' + dom
        if t == 'SRV' and not pre.startswith('_'):
            name = '_http._tcp.' + name
        out.append(synth.make_query(name, t, qid=rng.randrange(65536), rd=rng.random() < 0.85))
    out += [synth.make_query('%d.%d.0.10.in-addr.arpa' % (rng.randrange(256), rng.randrange(256)), 'PTR', rd=rng.random() < 0.8)
            for _ in range(60)]
    return out


@pytest.mark.parametrize('fi', range(len(FILTERS)))
def test_oracles_agree_with_filter(fi):
    region, dcs, ptr = FILTERS[fi]
    snap, info = fuzzgen.gen_zone(900 + fi, n_top=25)
    impl = H.make_impl('oracle', info['dns_domain'], snap, recursion=True)
    impl.set_recursion_filter(region, dcs, ptr)
    opts = H.ref_options(snap, info['dns_domain'], recursion=True)
    opts.recursion_filter = (region, set(dcs), ptr)
    pkts = filter_queries(fi, info)
    res, miss = H.resolve_list(impl, pkts, seed=5)
    n_miss = n_ref = 0
    for i, (pkt, (st, wire)) in enumerate(zip(pkts, res)):
        ref = H.ref_semantic(opts, pkt, seed=5, qidx=i)
        assert st == ref[0], (i, pkt)
        n_miss += st == 1
        if st == 0 and H.split_query(pkt)[3] == 0:
            assert H.decode_semantic(wire)[0] == ref[1], (i, pkt)
    # the filter only ever turns a miss into REFUSED: without it the same queries miss at least as often
    impl.set_recursion_filter(None)
    res0, miss0 = H.resolve_list(impl, pkts, seed=5)
    assert set(miss) <= set(miss0)
    for i in set(miss0) - set(miss):
        assert res[i][0] == 0 and res[i][1][3] & 0xF == 5          # REFUSED
    if fi == 0:
        assert 0 < len(miss) < len(miss0)


@pytest.mark.gpu
@pytest.mark.parametrize('fi', range(len(FILTERS)))
def test_gpu_filter_is_bit_exact(fi):
    from test_gpu_parity import assert_same
    region, dcs, ptr = FILTERS[fi]
    snap, info = fuzzgen.gen_zone(900 + fi, n_top=25)
    gpu = H.make_impl('gpu', info['dns_domain'], snap, recursion=True, ordered=fi % 2 == 1)
    orc = H.make_impl('oracle', info['dns_domain'], snap, recursion=True)
    gpu.set_recursion_filter(region, dcs, ptr)
    orc.set_recursion_filter(region, dcs, ptr)
    data, off = synth.pack_batch(filter_queries(fi, info))
    assert_same(gpu, orc, data, off, seed=5)
    gpu.set_recursion_filter(None); orc.set_recursion_filter(None)
    assert_same(gpu, orc, data, off, seed=5)


@pytest.mark.gpu
def test_filter_needs_recursion():
    from binder_b200._lib import BinderError
    gpu = H.make_impl('gpu', 'foo.com', None, recursion=False)
    with pytest.raises(BinderError):
        gpu.set_recursion_filter('foo.com', ['web'], True)
