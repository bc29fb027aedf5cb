"""Per-stage timing of route_push_kernel on rank 0 (run under torchrun; bb_engine_set_stage_log)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np, torch
import torch.distributed as dist
from binder_b200 import synth
from binder_b200._lib import lib
from binder_b200.shard import ShardedEngine

rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE']); lr = int(os.environ.get('LOCAL_RANK', rank))
torch.cuda.set_device(lr); dev = torch.device('cuda', lr)
dist.init_process_group('nccl', device_id=dev)
B = 65536
zone = synth.gen_zone(1000000)
se = ShardedEngine(zone.dns_domain, zone.datacenter, zone.jsonl, rank, world, lr, max_batch=B, dist=dist, lanes=1)
data, off = synth.batch_host_a_fast(zone, B, seed=rank)
pk = torch.from_numpy(data).to(dev); of = torch.from_numpy(off.view(np.int32)).to(dev)
st = torch.cuda.current_stream()
nt = (B + 127) // 128; NS = 16
log = torch.zeros((nt + 1) * NS, dtype=torch.int64, device=dev)
rows = []
for it in range(12):
    torch.cuda.synchronize(); dist.barrier()
    if it == 4:
        lib().bb_engine_set_stage_log(se.engine._h, log.data_ptr())
    se.route_push(pk.data_ptr(), of.data_ptr(), B, rank * B, st.cuda_stream)
    torch.cuda.synchronize()
    if it >= 4:
        rows.append(log.cpu().numpy().reshape(nt + 1, NS).copy())
    lib().bb_engine_set_stage_log(se.engine._h, None)
    se.resolve(1, st.cuda_stream)                # keeps the exchange protocol in step (stamps off)
    torch.cuda.synchronize()
    if it >= 4:
        lib().bb_engine_set_stage_log(se.engine._h, log.data_ptr())
if rank == 0:
    names = {0: 'start', 1: 'offsets', 2: 'staged', 4: 'hashed', 6: 'routed', 7: 'claimed', 8: 'grouped', 9: 'stored', 10: 'fenced'}
    idx = sorted(names)
    for a in rows[-3:]:
        t0 = a[:nt, 0].min()
        rel = (a[:nt][:, idx] - t0) / 1000.0
        print('push span %.2f us (first start -> flags published); ' % ((a[nt, 0] - t0) / 1000.0) +
              ', '.join('%s %.2f/%.2f' % (names[i], rel[:, j].mean(), rel[:, j].max()) for j, i in enumerate(idx)) + '  (mean/max completion over tiles, us)', flush=True)
dist.barrier(); dist.destroy_process_group()
