"""The reference's own integration-test expectations (test/*.test.js, restated in
tests/golden/reference_cases.json) checked against
  * oracle/binder_ref.py   (semantic restatement),
  * oracle/liboracle.so    (bytes, decoded by dnspython),
  * the CUDA path through the C ABI (gpu marker).
This is what pins the oracle to the reference."""
import json
import os

import dns.rcode
import pytest

import helpers as H
from binder_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, 'golden', 'reference_cases.json')))


def _case_iter():
    for si, suite in enumerate(GOLD['suites']):
        for ci, case in enumerate(suite['cases']):
            yield pytest.param(si, ci, id='%s' % case['source'].replace('test/', ''))


CASES = list(_case_iter())


def _check(case, rcode, answers):
    """What test/dig.js scrapes and the reference test asserts."""
    if 'status' in case:
        assert dns.rcode.to_text(rcode) == case['status']
    assert len(answers) == len(case['answers'])
    got = []
    for a in answers:
        d = {'name': a[0], 'ttl': a[1], 'type': a[2]}
        if a[2] == 'SRV':
            d['port'] = a[3]
            d['target'] = a[4] + '.'
        elif a[2] == 'PTR':
            d['target'] = a[3] + '.'
        else:
            d['target'] = a[3]
        got.append(d)
    want = case['answers']
    if case.get('any_order'):
        key = lambda d: d['target']
        got, want = sorted(got, key=key), sorted(want, key=key)
    for g, w in zip(got, want):
        for k, v in w.items():
            assert g[k] == v, (k, g, w)


@pytest.mark.parametrize("si,ci", CASES)
def test_reference_semantic_restatement(si, ci):
    suite, case = GOLD['suites'][si], GOLD['suites'][si]['cases'][ci]
    opts = H.ref_options(H.snapshot(suite['snapshot']), GOLD['dns_domain'])
    pkt = synth.make_query(case['name'], case['type'])
    status, rcode, answers, authority, additional = H.ref_semantic(opts, pkt, seed=7, qidx=3)
    assert status == 0
    _check(case, rcode, answers)


def _run_impl(kind, si, ci):
    suite, case = GOLD['suites'][si], GOLD['suites'][si]['cases'][ci]
    snap = H.snapshot(suite['snapshot'])
    impl = H.make_impl(kind, GOLD['dns_domain'], snap)
    # dig sends RD=1 and an OPT RR; check both shapes
    for edns in (None, 4096):
        pkt = synth.make_query(case['name'], case['type'], qid=0xBEEF, edns=edns)
        (res,), miss = H.resolve_list(impl, [pkt], seed=7, qidx_base=3)
        st, wire = res
        assert st == 0 and miss == []
        rcode, answers, authority, additional, info = H.decode_semantic(wire)
        _check(case, rcode, answers)
        assert info['id'] == 0xBEEF and info['qr'] and info['rd'] and not info['ra'] and not info['tc']
        assert info['edns'] == (edns is not None)
        assert info['question'][0][0] == case['name']
        # and the independent semantic restatement agrees on every section
        opts = H.ref_options(snap, GOLD['dns_domain'])
        assert H.ref_semantic(opts, pkt, seed=7, qidx=3) == (0, rcode, answers, authority, additional)


@pytest.mark.parametrize("si,ci", CASES)
def test_reference_cases_oracle(si, ci):
    _run_impl('oracle', si, ci)


@pytest.mark.gpu
@pytest.mark.parametrize("si,ci", CASES)
def test_reference_cases_gpu(si, ci):
    _run_impl('gpu', si, ci)
