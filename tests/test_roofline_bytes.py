"""The algorithmic-bytes accounting behind `roofline.achieved` is pinned here: the worked example of
config 2 (163 bytes per query) and payload(service) recomputed literally from the generated snapshot."""
import json

import numpy as np

import helpers as H
from binder_b200 import synth


def test_config2_is_163_bytes_per_query():
    z = synth.gen_zone(5000)
    data, off, meta = synth.gen_batch(z, 2000, 1, synth.WORKLOADS['config2'][2], 0.0)
    orc = H.make_impl('oracle', z.dns_domain, z.jsonl)
    out, ooff, olen, status, miss = orc.resolve_batch(data, off)
    rd, wr = synth.algorithmic_bytes(z, off, meta, olen)
    # 48-byte query + 4, probe 30-char key + 1 + 8 -> 91 read; 64-byte answer + 8 -> 72 written
    assert (rd, wr) == (91 * 2000, 72 * 2000)


def test_service_payload_matches_the_snapshot():
    z = synth.gen_zone(30000, service_frac=0.15)
    kids = {}                                   # service path -> [(label, type, record)]
    svc_order = []
    for line in z.jsonl.decode().splitlines():
        o = json.loads(line)
        d = o.get('data')
        if isinstance(d, dict) and d.get('type') == 'service':
            svc_order.append(o['path']); kids[o['path']] = []
        else:
            parent = o['path'].rsplit('/', 1)[0]
            if parent in kids and isinstance(d, dict):
                kids[parent].append((o['path'].rsplit('/', 1)[1], d['type'], d.get(d['type'], {})))
    member = {'load_balancer', 'moray_host', 'redis_host', 'ops_host', 'rr_host'}      # lib/server.js:352-360: `host`, `db_host` are not members
    want = []
    for p in svc_order:
        pay = 16
        for label, typ, rec in kids[p]:
            if typ in member:
                pay += len(label) + 1 + 4 + 4 + 2 * len(rec.get('ports', [0]))
        want.append(pay)
    got = synth.service_payload_bytes(z)
    assert len(want) == z.n_services and np.array_equal(got[:len(want)], np.array(want))
