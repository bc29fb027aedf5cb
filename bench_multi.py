"""bench.py's N>1 arm: one rank per GPU, zone sharded by key hash, one route+push exchange
over NVLink peer memory per step.  Weak scaling: every rank ingests its own
65,536-query batch per step; value = all ranks' queries / max-over-ranks device time."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))


def main(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from binder_b200 import synth, build
    from binder_b200.shard import ShardedEngine
    B1 = sys.modules['__main__']            # bench.py itself (it owns the real stdout)

    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    dist.init_process_group('nccl', device_id=dev)
    if rank == 0:
        build.build()
    dist.barrier()
    B = args.batch
    zone = synth.gen_zone(args.zone_records)
    LANES = int(os.environ.get('BB_LANES', '8'))
    sync = os.environ.get('BB_SYNC', 'flags')
    se = ShardedEngine(zone.dns_domain, zone.datacenter, zone.jsonl, rank, world, local_rank, max_batch=B,
                       ordered=args.ordered, dist=dist, lanes=LANES, sync=sync)
    RING = 8
    ring = [synth.batch_host_a_fast(zone, B, seed=5000 + 97 * rank + r) for r in range(RING)]
    d = [(torch.from_numpy(x).to(dev), torch.from_numpy(o.view(np.int32)).to(dev)) for x, o in ring]
    stream = torch.cuda.current_stream()

    lane_streams = [torch.cuda.Stream(device=dev) for _ in range(LANES)]
    lane_handles = [st.cuda_stream for st in lane_streams]
    d_ptrs = [(a.data_ptr(), b.data_ptr()) for a, b in d]

    def step(k, lane=None):
        """Step k on lane k % LANES (its own stream and receive regions): LANES steps in flight."""
        lane = k % LANES if lane is None else lane
        pk, off = d_ptrs[k % RING]
        if sync == 'nccl':
            with torch.cuda.stream(lane_streams[lane]):
                se.step(pk, off, B, rank * B, 0xB1DDE5, lane_handles[lane], lane)
        else:
            se.step(pk, off, B, rank * B, 0xB1DDE5, lane_handles[lane], lane)

    def timed(k0, nsteps):
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record(stream)
        for st in lane_streams:
            st.wait_event(ea)
        for k in range(nsteps):
            step(k0 + k)
        for st in lane_streams:
            ev = torch.cuda.Event()
            ev.record(st)
            stream.wait_event(ev)
        eb.record(stream)
        torch.cuda.synchronize()
        return ea.elapsed_time(eb)

    timed(0, max(args.warmup, LANES))
    torch.cuda.synchronize()
    dist.barrier()
    sampler = B1.ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = se.engine.launch_count()
    # all ranks issue the same number of steps per lane (the peer flags count steps)
    nsteps = (args.steps + LANES - 1) // LANES * LANES
    torch.cuda.synchronize()
    dist.barrier()
    elapsed = timed(nsteps, nsteps)
    args.steps = nsteps
    ms = torch.tensor([elapsed], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    launches = se.engine.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None

    # what the last timed step produced: every query answered exactly once somewhere
    owned = 0
    last_lane = (2 * nsteps - 1) % LANES
    for src in range(world):
        reg = se.fetch(src, last_lane, copy=False)
        owned += reg['n']
        assert (reg['status'] == 0).all() and (reg['out_len'] == 64).all(), 'unexpected result in timed batch'
    tot = torch.tensor([owned], device=dev, dtype=torch.int64)
    dist.all_reduce(tot)
    assert int(tot.item()) == B * world, ('coverage', int(tot.item()))

    # ---- e2e: pinned host ingress -> H2D -> route+push -> resolve -> answers in pinned host memory ----
    # LANES steps in flight, each lane on its own stream; the owner's resolve kernel writes its answers
    # straight into the shard's pinned host mirrors (zero-copy), so a step's device-to-host traffic is
    # exactly its results.
    se.set_host_results(True)
    h_ring = [(torch.from_numpy(x).pin_memory(), torch.from_numpy(o.view(np.int32)).pin_memory()) for x, o in ring[:4]]
    d_in = [(torch.empty_like(d[0][0]), torch.empty_like(d[0][1])) for _ in range(LANES)]
    evs = [torch.cuda.Event() for _ in range(LANES)]
    ksteps = (max(8, min(args.steps, 400)) + LANES - 1) // LANES * LANES

    def e2e_issue(k):
        lane = k % LANES
        hp, ho = h_ring[k % len(h_ring)]
        with torch.cuda.stream(lane_streams[lane]):
            d_in[lane][0].copy_(hp, non_blocking=True); d_in[lane][1].copy_(ho, non_blocking=True)
            se.step(d_in[lane][0].data_ptr(), d_in[lane][1].data_ptr(), B, rank * B, 0xB1DDE5, lane_handles[lane], lane)
            evs[lane].record(lane_streams[lane])

    def e2e_collect(k):
        lane = k % LANES
        evs[lane].synchronize()
        nq = nb = 0
        for src in range(world):
            n_src, _, bytes_src = se.totals(src, lane)
            nq += n_src; nb += bytes_src
        return nq, nb

    def e2e_run(nsteps):
        nq = nb = 0
        for k in range(nsteps):
            if k >= LANES:
                a, b2 = e2e_collect(k - LANES); nq += a; nb += b2
            e2e_issue(k)
        for k in range(nsteps - LANES, nsteps):
            a, b2 = e2e_collect(k); nq += a; nb += b2
        return nq, nb

    e2e_run(LANES)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    owned_q, owned_b = e2e_run(ksteps)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], device=dev)
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    cov = torch.tensor([owned_q], device=dev, dtype=torch.int64)
    dist.all_reduce(cov)
    assert int(cov.item()) == B * world * ksteps, ('e2e coverage', int(cov.item()))
    e2e = {'value': world * B * ksteps / float(dt.item()), 'unit': B1.UNIT,
           'h2d_bytes_per_step': int(ring[0][1][B]) + (B + 1) * 4, 'd2h_bytes_per_step': (owned_b + 11 * owned_q) // ksteps,
           'steps': ksteps, 'in_flight': LANES,
           'timing': 'wall clock, max over ranks; %d steps in flight per rank' % LANES,
           'api': 'pinned H2D + bb_shard_route_push + bb_shard_resolve writing into pinned host mirrors (bb_shard_host_results) + bb_shard_results'}

    # step-level roofline per GPU (the kernels of different steps overlap, so there is no per-kernel
    # duration here): algorithmic HBM bytes of one rank's step = its batch parsed twice (route, then
    # resolve: 2 x (packet + 4)) + probe + answers, per SURVEY.md section 8d, over the step time
    peak, peak_src = B1.measured_peaks()
    in_b = int(ring[0][1][B]) + 4 * B
    step_bytes = 2 * in_b + B * (30 + 1 + 8) + B * (64 + 8)
    achieved = step_bytes / (ms / args.steps * 1e-3) / 1e9
    roofline = {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak, 'traffic': None,
                'peak_source': peak_src, 'kernel': 'route_push_kernel + resolve_kernel (per GPU, per step, steps overlapped)',
                'algorithmic_bytes_per_launch': step_bytes,
                'nvlink_bytes_per_step': int((in_b + 8 * B) * (world - 1) / world)}
    if rank == 0:
        line = {'metric': B1.METRIC, 'value': world * B * args.steps / (ms * 1e-3), 'unit': B1.UNIT, 'n_gpus': world,
                'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True,
                'scaling': 'weak', 'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic',
                'config': {'workload': B1.WORKLOAD + '; zone hash-sharded over %d ranks, each rank ingests its own batch' % world,
                           'zone_records': zone.n_records, 'batch_per_rank': B, 'global_batch': B * world,
                           'shard_table_mb': se.zone_stat['image_bytes'] / 1e6,
                           'parallelism': 'shard%d: route+push kernel stores each query into its owner rank over NVLink peer memory; %s; owner resolves; %d steps in flight' % (world, 'per-region epoch flags + device-side wait (no collective)' if sync == 'flags' else '1-element NCCL all-reduce as barrier', LANES),
                           'l2_policy': 'ring of %d distinct batches per rank; shard table %.0f MB' % (RING, se.zone_stat['image_bytes'] / 1e6),
                           'output_packing': 'query order' if args.ordered else 'arrival'},
                'clocks': clocks, 'e2e': e2e, 'gpu_launches': int(launches), 'roofline': roofline, 'cpu_baseline': None}
        B1.emit(line)
    dist.barrier()
    dist.destroy_process_group()
