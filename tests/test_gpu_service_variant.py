"""The service variant of the resolve kernel (bb_engine_set_kernel_profile(e, 2): big tiles of long answers go through
emit rounds, the children's ready RRs as copy jobs run by the whole tile) against the CPU oracle, bit for bit — the
variant changes how answers are assembled, never what they are."""
import numpy as np
import pytest

import fuzzgen
import helpers as H
from binder_b200 import synth
from test_gpu_parity import assert_same, MODES
from test_host_emulation import job_zone

pytestmark = pytest.mark.gpu


def pair(dns_domain, snap, recursion=False, ordered=False):
    gpu = H.make_impl('gpu', dns_domain, snap, recursion=recursion, ordered=ordered)
    gpu.set_kernel_profile('service')
    return gpu, H.make_impl('oracle', dns_domain, snap, recursion=recursion)


@pytest.mark.parametrize('ordered', MODES)
@pytest.mark.parametrize('edns', [0, 1200, 700])
def test_copy_jobs_edges(edns, ordered):
    """More jobs than the list holds, SRV runs cut short by truncation, the OPT as a job, many emit rounds, upper-case
    names (not job mode) in between — the same batch as tests/test_host_emulation.py::test_emulated_copy_jobs_edges."""
    gpu, orc = pair('foo.com', job_zone(), ordered=ordered)
    pk = []
    for i in range(3000):
        s = (i * 7) % 24
        name = 'svc%02d.foo.com' % s
        if i % 11 == 0:
            name = name.upper()[:5] + name[5:]
        pk.append(synth.make_query(name, 'A', i & 0xFFFF, edns=edns) if i % 3 == 2 else
                  synth.make_query('_http._tcp.' + name, 'SRV', i & 0xFFFF, edns=edns))
    data, off = synth.pack_batch(pk)
    for seed in (1, 2):
        assert_same(gpu, orc, data, off, seed=seed)
    assert_same(gpu, orc, data, off, seed=3, tcp=True)


@pytest.mark.parametrize('seed', range(6))
def test_fuzz_zone_parity_service_variant(seed):
    snap, info = fuzzgen.gen_zone(seed, n_top=40)
    gpu, orc = pair(info['dns_domain'], snap, seed % 3 == 0, ordered=seed % 2 == 1)
    pkts = fuzzgen.gen_queries(seed, info, n=3000) + fuzzgen.malformed_packets() + fuzzgen.tolerated_packets()
    data, off = synth.pack_batch(pkts)
    assert_same(gpu, orc, data, off, seed=seed * 1315423911 + 3, qidx_base=seed * 1000)


@pytest.mark.parametrize('ordered', MODES)
def test_config3_and_config4_shapes(ordered):
    """BASELINE configs 3 and 4 in miniature (services SRV + A; hosts / SRV / AAAA) plus 40 % misses with recursion."""
    z = synth.gen_zone(300000, service_frac=0.15)
    gpu, orc = pair(z.dns_domain, z.jsonl, recursion=True, ordered=ordered)
    pk = synth.batch_service(z, 40000, seed=5) + synth.batch_mixed(z, 25536, seed=6, miss_frac=0.4)
    data, off = synth.pack_batch(pk)
    assert_same(gpu, orc, data, off, seed=0xB1DDE5)
    d4, o4, _ = synth.gen_batch(z, 65536, 9, synth.WORKLOADS['config4'][2], 0.0)
    assert_same(gpu, orc, d4, o4, seed=17, qidx_base=123456)


def test_variants_agree_byte_for_byte():
    """Same engine, same batch, both variants: identical responses (arrival packing may place them differently)."""
    from binder_b200.engine import repack
    z = synth.gen_zone(100000, service_frac=0.2)
    gpu = H.make_impl('gpu', z.dns_domain, z.jsonl)
    data, off = synth.pack_batch(synth.batch_service(z, 30000, seed=3))
    res = {}
    for prof in ('small', 'service'):
        gpu.set_kernel_profile(prof)
        out, ooff, olen, status, miss = gpu.resolve_batch(data, off, seed=5)
        res[prof] = (repack(out, ooff, olen), olen, status)
    assert np.array_equal(res['small'][0][0], res['service'][0][0]) and np.array_equal(res['small'][1], res['service'][1])
    assert np.array_equal(res['small'][2], res['service'][2])
