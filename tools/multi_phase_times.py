"""Where a sharded step's time goes (run under torchrun): per-kernel durations via CUDA events."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np, torch, time
import torch.distributed as dist
from binder_b200 import synth
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
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
acc = np.zeros(2); host = np.zeros(2)
for it in range(30):
    torch.cuda.synchronize(); dist.barrier()
    ev[0].record(st)
    t0 = time.perf_counter()
    se.route_push(pk.data_ptr(), of.data_ptr(), B, rank * B, st.cuda_stream)
    t1 = time.perf_counter()
    ev[1].record(st)
    se.resolve(1, st.cuda_stream)
    t2 = time.perf_counter()
    ev[2].record(st)
    torch.cuda.synchronize()
    if it >= 5:
        acc += [ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])]
        host += [t1 - t0, t2 - t1]
if rank == 0:
    print('world %d: route_push %.1f us, wait+resolve %.1f us (device); host issue %.1f / %.1f us' % (world, acc[0] / 25 * 1e3, acc[1] / 25 * 1e3, host[0] / 25 * 1e6, host[1] / 25 * 1e6), flush=True)
dist.barrier(); dist.destroy_process_group()
