"""Cost of an incremental zone update on the bench zone (1M records): host-side bb_zone_apply, then
bb_engine_apply_update (changed slots + arena tail) against a full bb_engine_swap_zone."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from binder_b200 import synth, build
from binder_b200.engine import Engine, Zone

build.build()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
zone_src = synth.gen_zone(N, service_frac=0.15)
t0 = time.perf_counter(); zone = Zone(zone_src.jsonl, zone_src.dns_domain); t_build = time.perf_counter() - t0
eng = Engine(zone_src.dns_domain, zone_src.datacenter, device=0, max_batch=65536)
t0 = time.perf_counter(); eng.swap_zone(zone); t_swap = time.perf_counter() - t0
print('zone %d records: build %.2f s, full upload (swap) %.3f s, image %.0f MB' % (N, t_build, t_swap, zone.stat()['image_bytes'] / 1e6))
rng = np.random.default_rng(1)
for n_upd in (1, 100, 10000):
    lines = []
    for i in rng.integers(0, zone_src.n_hosts, size=n_upd):
        i = int(i)
        lines.append(json.dumps({'path': '%s/g%04d/h%07d' % (synth.ROOT_PATH, i % synth.N_GROUPS, i),
                                 'data': {'type': 'host', 'host': {'address': '10.%d.%d.%d' % ((i >> 16) & 255, (i >> 8) & 255, (i + 1) & 255)}, 'ttl': 17}}))
    for j in rng.integers(0, max(zone_src.n_services, 1), size=max(n_upd // 10, 1)):      # a service child changes address
        j = int(j)
        lines.append(json.dumps({'path': '%s/svc%06d/lb00' % (synth.ROOT_PATH, j),
                                 'data': {'type': 'load_balancer', 'load_balancer': {'address': '172.31.%d.%d' % (j >> 8 & 255, j & 255)}}}))
    delta = ('\n'.join(lines) + '\n').encode()
    t0 = time.perf_counter(); zone.apply(delta); t_apply = time.perf_counter() - t0
    dirty, relaid = zone.pending()
    t0 = time.perf_counter(); eng.apply_update(zone); t_up = time.perf_counter() - t0
    print('%6d events: host apply %.3f ms (%.2f us/event), %d slots dirty%s, device update %.3f ms' %
          (len(lines), t_apply * 1e3, t_apply * 1e6 / len(lines), dirty, ' (re-laid)' if relaid else '', t_up * 1e3))
# sanity: an updated host answers with its new address
data, off = synth.pack_batch([synth.make_query(synth.host_name(5), 'A')])
out, ooff, olen, st, miss = eng.resolve_batch(data, off)
print('h0000005 ->', bytes(out[ooff[0]:ooff[0] + olen[0]])[-4:].hex())
