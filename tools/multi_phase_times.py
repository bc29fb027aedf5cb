"""Where a sharded step's time goes (run under torchrun): per-kernel durations via CUDA events, and the NVLink bytes a step
actually moves (nvidia-smi nvlink -gt d deltas of GPU 0 over a counted number of steps) next to bench_multi's estimate.
Workload: config 4's mix on a BB_ZONE-record zone (default 3,000,000), BB_BATCH queries per rank (default 1,048,576 / world)."""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np, torch, time
import torch.distributed as dist
from binder_b200 import synth
from binder_b200.shard import ShardedEngine

rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE']); lr = int(os.environ.get('LOCAL_RANK', rank))
torch.cuda.set_device(lr); dev = torch.device('cuda', lr)
dist.init_process_group('nccl', device_id=dev)
B = int(os.environ.get('BB_BATCH', str(1048576 // world)))
desc, service_frac, mix, miss_frac, recursion = synth.WORKLOADS['config4']
zone = synth.gen_zone(int(os.environ.get('BB_ZONE', '3000000')), service_frac=service_frac)
se = ShardedEngine(zone.dns_domain, zone.datacenter, zone.jsonl, rank, world, lr, max_batch=B, dist=dist, lanes=1)
data, off, _ = synth.gen_batch(zone, B, 5000 + rank, mix, miss_frac)
pk = torch.from_numpy(data).to(dev); of = torch.from_numpy(off.view(np.int32)).to(dev)
st = torch.cuda.current_stream()


def nvlink_kib(which):
    """sum over GPU 0's links of the Data Tx / Rx counters (KiB)"""
    try:
        out = subprocess.run(['nvidia-smi', 'nvlink', '-gt', 'd', '-i', '0'], capture_output=True, text=True, timeout=30).stdout
        return sum(int(x) for x in re.findall(r'Data %s:\s*(\d+)\s*KiB' % which, out))
    except Exception:
        return None


ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
acc = np.zeros(2); host = np.zeros(2)
ITERS, SKIP = 40, 5
tx0 = rx0 = None
for it in range(ITERS):
    torch.cuda.synchronize(); dist.barrier()
    if it == SKIP and rank == 0:
        tx0, rx0 = nvlink_kib('Tx'), nvlink_kib('Rx')
    ev[0].record(st)
    t0 = time.perf_counter()
    se.route_push(pk.data_ptr(), of.data_ptr(), B, rank * B, st.cuda_stream)
    t1 = time.perf_counter()
    ev[1].record(st)
    se.resolve(1, st.cuda_stream)
    t2 = time.perf_counter()
    ev[2].record(st)
    torch.cuda.synchronize()
    if it >= SKIP:
        acc += [ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])]
        host += [t1 - t0, t2 - t1]
dist.barrier()
if rank == 0:
    n = ITERS - SKIP
    tx1, rx1 = nvlink_kib('Tx'), nvlink_kib('Rx')
    routed = int(off[B]) + 8 * B            # packets + offset + ingress index of every query, if all were routed
    est = routed * (world - 1) / world * 0.8  # 20 % AAAA stay on the ingress rank; 1/world of the rest is local
    print('world %d, config 4 mix, %d queries per rank per step, %d-record zone: route_push %.1f us, wait+resolve %.1f us (device, one step at a time); '
          'host issue %.1f / %.1f us' % (world, B, zone.n_records, acc[0] / n * 1e3, acc[1] / n * 1e3, host[0] / n * 1e6, host[1] / n * 1e6), flush=True)
    if None not in (tx0, tx1, rx0, rx1):
        print('NVLink, GPU 0, per step (nvidia-smi nvlink -gt d deltas over %d steps): tx %.0f bytes, rx %.0f bytes; bench_multi estimate %.0f bytes pushed'
              % (n, (tx1 - tx0) * 1024.0 / n, (rx1 - rx0) * 1024.0 / n, est), flush=True)
    else:
        print('NVLink counters unavailable (nvidia-smi nvlink -gt d)', flush=True)
dist.barrier(); dist.destroy_process_group()
