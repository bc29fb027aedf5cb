"""Per-stage timing of the resolve kernel (bb_engine_set_stage_log) on the bench workload."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from binder_b200 import synth, build
from binder_b200.engine import Engine
from binder_b200._lib import lib

# the stamps are compiled in only with -DBB_STAGE_LOG: build such a library for this run, restore the default afterwards
import atexit
os.environ['BB_NVCC_DEFINES'] = (os.environ.get('BB_NVCC_DEFINES', '') + ' -DBB_STAGE_LOG').strip()
build.build(force=True)
atexit.register(lambda: (os.environ.pop('BB_NVCC_DEFINES', None), build.build(force=True)))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
WL = os.environ.get('BB_WL', 'config2')
zone = synth.gen_zone(int(os.environ.get('BB_ZONE', '1000000')), service_frac=synth.WORKLOADS[WL][1])
ORDERED = os.environ.get('BB_ORDERED', '0') == '1'
eng = Engine(zone.dns_domain, zone.datacenter, device=0, max_batch=B, max_batch_bytes=B * 64, snapshot=zone.jsonl, ordered=ORDERED)
dev = torch.device('cuda:0')
bufs = []
for r in range(8):
    data, off = synth.gen_batch(zone, B, r, synth.WORKLOADS[WL][2], synth.WORKLOADS[WL][3])[:2]
    bufs.append((torch.from_numpy(data).to(dev), torch.from_numpy(off.view(np.int32)).to(dev)))
out = torch.empty(B * 512, dtype=torch.uint8, device=dev); oo = torch.empty(B + 1, dtype=torch.int32, device=dev)
st = torch.empty(B, dtype=torch.uint8, device=dev); ms = torch.empty(B, dtype=torch.int32, device=dev)
tot = torch.zeros(4, dtype=torch.int32, device=dev); ol = torch.empty(B, dtype=torch.int16, device=dev)
nt = (B + 127) // 128
NS = 16
log = torch.zeros(nt * NS, dtype=torch.int64, device=dev)
s = torch.cuda.current_stream().cuda_stream
def run(k):
    pk, of = bufs[k % 8]
    eng.resolve_device(pk.data_ptr(), of.data_ptr(), B, 1, 0, out.data_ptr(), out.numel(), oo.data_ptr(), ol.data_ptr(), st.data_ptr(), ms.data_ptr(), tot.data_ptr(), s)
for k in range(5): run(k)
torch.cuda.synchronize()
lib().bb_engine_set_stage_log(eng._h, log.data_ptr())
acc = []
for k in range(10):
    run(5 + k); torch.cuda.synchronize()
    full = log.cpu().numpy().reshape(nt, NS).copy(); acc.append(full[:, :11]); svc = full[:, 11:15]
print('ordered' if ORDERED else 'arrival', 'packing, batch', B)
names = ['start', 'offsets', 'staged', 'decoded', 'hashed', 'probed', 'sized', 'scan', 'placed', 'emitted', 'flushed']
for a in acc[-3:]:
    t0 = a[:, 0].min()
    rel = (a - t0) / 1000.0
    print('kernel span %.2f us (first start -> last flush); tile starts spread %.2f us' % (rel[:, 10].max(), rel[:, 0].max()))
    d = np.diff(a, axis=1) / 1000.0
    print('  stage durations us (mean / max over tiles): ' + ', '.join('%s %.2f/%.2f' % (names[i + 1], d[:, i].mean(), d[:, i].max()) for i in range(10)))
    print('  completion time of each stage, max over tiles: ' + ', '.join('%s %.2f' % (names[i], rel[:, i].max()) for i in range(11)))

