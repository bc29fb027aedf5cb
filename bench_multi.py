"""bench.py's N>1 arm: one rank per GPU.

Default workload = BASELINE.json configs[3] ("config4"): 10M-record zone, a global batch of 1,048,576 queries per
step split evenly over the ranks' ingress, 60 % A(host) / 20 % SRV(service) / 20 % AAAA (-> NOTIMP on the ingress
rank, lib/server.js:491-506).  Strong scaling in the batch (the global batch is fixed), weak in nothing else;
`scaling` says "strong".

  --mode shard     (default) zone hash-sharded by lookup key; each rank parses its ingress slice once, stores every
                   routed record straight into its owner rank's HBM over NVLink peer memory (route+push, one
                   kernel), owners resolve and answer.  No collective on the data path.
  --mode replicas  every GPU holds the full zone and resolves its own slice: the reference's own scale-out
                   (boot/setup.sh:136-149) and the baseline sharding has to beat.
  --mode nccl      sharded zone, the same routed records exchanged with ONE NCCL all-to-all (ncclSend/ncclRecv
                   grouped) instead of peer stores: the collective baseline of section 8(e).

After the timed region every rank's answers for one more step are compared byte for byte with the CPU oracle
(rank 0 loads it in the background while the GPUs set up; every rank hands its results over through /dev/shm).
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))


def _numa_bind(local_rank):
    """Pin this rank's host threads to the CPUs next to its GPU (nvidia-smi topo's CPU affinity), so that the
    pinned buffers it allocates are first-touched on that NUMA node.  Best effort."""
    try:
        import subprocess
        out = subprocess.run(['nvidia-smi', 'topo', '-m'], capture_output=True, text=True, timeout=20).stdout
        for ln in out.splitlines():
            if ln.startswith('GPU%d\t' % local_rank) or ln.startswith('GPU%d ' % local_rank):
                for tok in ln.split():
                    if '-' in tok and tok.replace('-', '').replace(',', '').isdigit():
                        cpus = set()
                        for part in tok.split(','):
                            a, b = part.split('-') if '-' in part else (part, part)
                            cpus.update(range(int(a), int(b) + 1))
                        cpus &= os.sched_getaffinity(0)
                        if cpus:
                            os.sched_setaffinity(0, cpus)
                            return '%d cpus (%s)' % (len(cpus), tok)
    except Exception as ex:
        return 'unbound (%r)' % (ex,)
    return 'unbound'


def main(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from binder_b200 import synth, build
    from binder_b200.engine import Engine, repack
    from binder_b200.shard import ShardedEngine
    B1 = sys.modules['__main__']            # bench.py itself (it owns the real stdout)

    numa = _numa_bind(local_rank)
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    dist.init_process_group('nccl', device_id=dev)
    if rank == 0:
        build.build()
    dist.barrier()
    zone, desc, mix, miss_frac, recursion, GB, per_q = B1.workload_setup(args)
    # auto: full replicas while one GPU holds the zone many times over (≈ 270 bytes of image per record against 180 GB of
    # HBM), the sharded zone beyond that.  Measured on config 4: replicas 5.7 / 22.9 G q/s at N = 2 / 8,
    # sharded with peer stores 4.8 / 15.3, sharded with an NCCL all-to-all 2.2 at N = 2.
    mode = args.mode
    mode_why = 'requested'
    if mode == 'auto':
        fits = zone.n_records * 270 * 8 < 180e9
        mode = 'replicas' if fits else 'shard'
        mode_why = ('auto: the zone image (~%.1f GB) fits one GPU many times over, and replicas measured faster than the sharded zone at every N '
                    '(profiles/r2_n_*, r2_o_*: N=2 5.74 vs 4.79 (peer stores) vs 2.23 (NCCL all-to-all) G q/s, N=8 22.9 vs 15.3)' % (zone.n_records * 270 / 1e9)
                    if fits else 'auto: the zone image does not fit one GPU several times over')
    B = GB // world                          # this rank's ingress slice of the global batch
    loader = B1.OracleLoader(zone, recursion) if (rank == 0 and not args.no_cpu) else None
    LANES = int(os.environ.get('BB_LANES', '4'))
    K = (args.steps + LANES - 1) // LANES * LANES      # all ranks issue the same number of steps per lane
    stream = torch.cuda.current_stream()
    lane_streams = [torch.cuda.Stream(device=dev) for _ in range(LANES)]
    lane_handles = [st.cuda_stream for st in lane_streams]
    t0 = time.time()

    def batch_of(r, k):                      # rank r's ingress slice of ring entry k (any rank can regenerate it)
        return synth.gen_batch(zone, B, 5000 + 97 * r + k, mix, miss_frac)

    RING = 4
    ring = [batch_of(rank, k) for k in range(RING)]
    d = [(torch.from_numpy(x).to(dev), torch.from_numpy(o.view(np.int32)).to(dev)) for x, o, _ in ring]
    d_ptrs = [(a.data_ptr(), b.data_ptr()) for a, b in d]
    out_cap = B * per_q

    if mode == 'replicas':
        eng = Engine(zone.dns_domain, zone.datacenter, recursion=recursion, device=local_rank, max_batch=B,
                     max_batch_bytes=B * 64, ordered=args.ordered)
        zstat = eng.load_snapshot(zone.jsonl)
        res = [dict(out=torch.empty(out_cap, dtype=torch.uint8, device=dev), oo=torch.empty(B + 1, dtype=torch.int32, device=dev),
                    st=torch.empty(B, dtype=torch.uint8, device=dev), ms=torch.empty(B, dtype=torch.int32, device=dev),
                    ol=torch.empty(B, dtype=torch.int16, device=dev), tot=torch.zeros(4, dtype=torch.int32, device=dev))
               for _ in range(RING)]

        def step(k, lane=None):
            lane = k % LANES if lane is None else lane
            pk, off = d_ptrs[k % RING]
            b = res[k % RING]
            eng.resolve_device(pk, off, B, B1.SEED, rank * B, b['out'].data_ptr(), out_cap, b['oo'].data_ptr(), b['ol'].data_ptr(),
                               b['st'].data_ptr(), b['ms'].data_ptr(), b['tot'].data_ptr(), lane_handles[lane])
        engine = eng
    else:
        se = ShardedEngine(zone.dns_domain, zone.datacenter, zone.jsonl, rank, world, local_rank, max_batch=B, recursion=recursion,
                           ordered=args.ordered, dist=dist, lanes=LANES, sync=('nccl_a2a' if mode == 'nccl' else 'flags'))
        zstat = se.zone_stat
        engine = se.engine

        def step(k, lane=None):
            lane = k % LANES if lane is None else lane
            pk, off = d_ptrs[k % RING]
            if mode == 'nccl':
                with torch.cuda.stream(lane_streams[lane]):
                    se.step(pk, off, B, rank * B, B1.SEED, lane_handles[lane], lane)
            else:
                se.step(pk, off, B, rank * B, B1.SEED, lane_handles[lane], lane)
    B1.log('[rank %d] %s ready in %.1fs (table %.0f MB, numa %s)' % (rank, mode, time.time() - t0, zstat['image_bytes'] / 1e6, numa))

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

    def max_over_ranks(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    timed(0, max((args.warmup + LANES - 1) // LANES * LANES, LANES))
    warm_ms = 0.0
    while warm_ms < 50.0:                    # same number of iterations on every rank: decided on the max
        warm_ms += max_over_ranks(timed(0, K))
    torch.cuda.synchronize()
    dist.barrier()
    sampler = B1.ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = engine.launch_count()
    regions = []
    for r in range(B1.REGIONS):
        torch.cuda.synchronize()
        dist.barrier()
        regions.append(max_over_ranks(timed(r * K, K)))
    ms = float(np.median(regions))
    launches = engine.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None

    # ---- parity: one more step on lane 0, every rank's answers handed to rank 0 and compared with the oracle ----
    tag = '/dev/shm/bb_parity_%s_' % os.environ.get('MASTER_PORT', '0')
    torch.cuda.synchronize()
    dist.barrier()
    step(0, 0)
    torch.cuda.synchronize()
    dist.barrier()
    regs = []
    if mode == 'replicas':
        b = res[0]
        ol = b['ol'].cpu().numpy().view(np.uint16)
        packed, _ = repack(b['out'].cpu().numpy(), b['oo'].cpu().numpy().view(np.uint32), ol)
        nm = int(b['tot'].cpu().numpy()[1])
        regs.append(dict(src=rank, qidx=np.arange(B, dtype=np.uint32) + rank * B, status=b['st'].cpu().numpy(), out_len=ol,
                         packed=packed, miss=np.sort(b['ms'].cpu().numpy().view(np.uint32)[:nm]) + rank * B))
    else:
        for src in range(world):
            reg = se.fetch(src, 0)
            packed, _ = repack(reg['out'], reg['out_off'], reg['out_len'])
            regs.append(dict(src=src, qidx=reg['qidx'], status=reg['status'], out_len=reg['out_len'], packed=packed,
                             miss=reg['qidx'][reg['miss']] if len(reg['miss']) else np.zeros(0, np.uint32)))
    np.savez(tag + '%d.npz' % rank, n=len(regs), **{'%s%d' % (k, i): v for i, g in enumerate(regs) for k, v in g.items()})
    answered_bytes = int(sum(int(g['out_len'].astype(np.int64).sum()) for g in regs))
    dist.barrier()
    parity = 'not checked (--no-cpu)'
    if rank == 0 and loader is not None:
        orc = loader.get()
        B1.log('[cpu] oracle loaded the %d-record zone in %.1fs (background)' % (zone.n_records, loader.secs))
        cores = os.cpu_count() or 1
        want = []
        for r in range(world):
            x, o, _ = batch_of(r, 0)
            want.append(orc.resolve_batch(x, o, seed=B1.SEED, qidx_base=r * B, nthreads=cores))
        seen = np.zeros(world * B, dtype=np.int32)
        nbytes = 0
        for r in range(world):
            z = np.load(tag + '%d.npz' % r)
            for i in range(int(z['n'])):
                src = int(z['src%d' % i])
                o_out, o_off, o_len, o_st, o_miss = want[src]
                qi = z['qidx%d' % i].astype(np.int64)
                local = qi - src * B
                assert local.min(initial=0) >= 0 and local.max(initial=0) < B, ('rank %d region %d: foreign index' % (r, src))
                np.add.at(seen, qi, 1)
                assert np.array_equal(z['status%d' % i], o_st[local]), 'rank %d region %d: statuses differ from the oracle' % (r, src)
                assert np.array_equal(z['out_len%d' % i], o_len[local]), 'rank %d region %d: lengths differ from the oracle' % (r, src)
                exp, _ = repack(o_out, o_off[:-1][local].astype(np.uint32) if len(local) else np.zeros(0, np.uint32), o_len[local])
                assert np.array_equal(z['packed%d' % i], exp), 'rank %d region %d: response bytes differ from the oracle' % (r, src)
                assert np.array_equal(np.sort(z['miss%d' % i].astype(np.int64) - src * B), np.sort(o_miss[np.isin(o_miss, local)])), 'miss list'
                nbytes += len(exp)
        assert (seen == 1).all(), 'coverage: %d queries answered != once' % int((seen != 1).sum())
        parity = 'bit-exact vs oracle, all ranks: %d queries (%d response bytes), each answered exactly once' % (world * B, nbytes)
    if rank == 0:
        for r in range(world):
            try:
                os.unlink(tag + '%d.npz' % r)
            except OSError:
                pass
    dist.barrier()

    # ---- e2e: pinned host ingress -> H2D -> (route+push ->) resolve -> answers in pinned host memory ----
    e2e = None
    if not args.no_e2e:
        ksteps = (max(8, min(K, 200)) + LANES - 1) // LANES * LANES
        h_ring = [(torch.from_numpy(x).pin_memory(), torch.from_numpy(o.view(np.int32)).pin_memory()) for x, o, _ in ring[:2]]
        in_bytes = int(np.mean([int(o[B]) for _, o, _ in ring[:2]])) + (B + 1) * 4
        if mode == 'replicas':
            import ctypes
            from binder_b200._lib import lib, check
            L = lib()
            nslots = L.bb_engine_slots(eng._h)
            hb = []
            for r in range(nslots):
                sizes = dict(out=out_cap, oo=(B + 1) * 4, ol=B * 2, st=B, ms=B * 4)
                hb.append(dict(ptr={k: L.bb_host_alloc(v) for k, v in sizes.items()}, nm=ctypes.c_uint32(0)))

            def e2e_run(nst):
                inflight = [False] * nslots
                nb = 0
                for k in range(nst):
                    slot = k % nslots
                    if inflight[slot]:
                        check(L.bb_resolve_wait(eng._h, slot))
                    hp, ho = h_ring[k % len(h_ring)]
                    h = hb[slot]
                    check(L.bb_resolve_submit(eng._h, slot, hp.data_ptr(), ho.data_ptr(), B, B1.SEED, rank * B, h['ptr']['out'], out_cap,
                                              h['ptr']['oo'], h['ptr']['ol'], h['ptr']['st'], h['ptr']['ms'], ctypes.byref(h['nm'])))
                    inflight[slot] = True
                for slot in range(nslots):
                    if inflight[slot]:
                        check(L.bb_resolve_wait(eng._h, slot))
                return B * nst, answered_bytes * nst
            api = 'bb_resolve_submit/bb_resolve_wait per rank, pinned host buffers'
        else:
            se.set_host_results(True)
            max_bytes = max(int(hp.numel()) for hp, _ in h_ring)              # the slices differ in size (mixed question types)
            d_in = [(torch.empty(max_bytes, dtype=torch.uint8, device=dev), torch.empty_like(d[0][1])) for _ in range(LANES)]
            evs = [torch.cuda.Event() for _ in range(LANES)]

            def e2e_issue(k):
                lane = k % LANES
                hp, ho = h_ring[k % len(h_ring)]
                with torch.cuda.stream(lane_streams[lane]):
                    d_in[lane][0][:hp.numel()].copy_(hp, non_blocking=True); d_in[lane][1].copy_(ho, non_blocking=True)
                    se.step(d_in[lane][0].data_ptr(), d_in[lane][1].data_ptr(), B, rank * B, B1.SEED, lane_handles[lane], lane)
                    evs[lane].record(lane_streams[lane])

            def e2e_collect(k):
                lane = k % LANES
                evs[lane].synchronize()
                nq = nb = 0
                for src in range(world):
                    n_src, _, bytes_src = se.totals(src, lane)
                    nq += n_src; nb += bytes_src
                return nq, nb

            def e2e_run(nst):
                nq = nb = 0
                for k in range(nst):
                    if k >= LANES:
                        a, b2 = e2e_collect(k - LANES); nq += a; nb += b2
                    e2e_issue(k)
                for k in range(nst - LANES, nst):
                    a, b2 = e2e_collect(k); nq += a; nb += b2
                return nq, nb
            api = 'pinned H2D + bb_shard_route_push + bb_shard_resolve writing into pinned host mirrors (bb_shard_host_results) + bb_shard_results'
        e2e_run(LANES * 2)
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        owned_q, owned_b = e2e_run(ksteps)
        torch.cuda.synchronize()
        dt_local = time.perf_counter() - t0
        dt = max_over_ranks(dt_local)
        cov = torch.tensor([owned_q], device=dev, dtype=torch.int64)
        dist.all_reduce(cov)
        assert int(cov.item()) == B * world * ksteps, ('e2e coverage', int(cov.item()))
        d2h = (owned_b + 11 * owned_q) // ksteps
        e2e = {'value': world * B * ksteps / dt, 'unit': B1.UNIT, 'h2d_bytes_per_step': in_bytes, 'd2h_bytes_per_step': int(d2h),
               'steps': ksteps, 'in_flight': LANES, 'per_rank_h2d_gbs': in_bytes * ksteps / dt / 1e9, 'per_rank_d2h_gbs': d2h * ksteps / dt / 1e9,
               'timing': 'wall clock, max over ranks', 'api': api, 'numa': numa}

    # step-level roofline per GPU (kernels of different steps overlap, so there is no single kernel duration here):
    # SURVEY.md section 8(d) bytes of this rank's ingress slice and its answers over the per-step time
    x, o, meta = ring[0]
    ol_est = np.zeros(B, dtype=np.int64)
    rd_b, wr_b = synth.algorithmic_bytes(zone, o, meta, ol_est)
    wr_b += answered_bytes                   # bytes this rank answered in the checked step (its owned share)
    peak, peak_src = B1.measured_peaks()
    step_bytes = rd_b + wr_b + (0 if mode == 'replicas' else int(o[B]) + 4 * B)   # sharded: the slice is read twice (route, then resolve)
    achieved = step_bytes / (ms / K * 1e-3) / 1e9
    in_b = int(o[B]) + 8 * B
    roofline = {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak, 'traffic': None,
                'peak_source': peak_src,
                'kernel': 'resolve_kernel (per GPU, per step)' if mode == 'replicas' else 'route_push_kernel + resolve_kernel (per GPU, per step, steps overlapped)',
                'algorithmic_bytes_per_launch': step_bytes,
                'nvlink_bytes_per_step_estimate': 0 if mode == 'replicas' else int(in_b * (world - 1) / world * 0.8)}
    if rank == 0:
        par = {'shard': 'shard%d: route+push kernel stores each routed record into its owner rank over NVLink peer memory; per-region epoch flags + device-side wait (no collective); owner resolves; %d steps in flight' % (world, LANES),
               'replicas': 'replicas%d: every GPU holds the full zone and resolves its own slice, no exchange; %d steps in flight' % (world, LANES),
               'nccl': 'shard%d: routed records exchanged with one NCCL all-to-all per step; owner resolves; %d steps in flight' % (world, LANES)}[mode]
        line = {'metric': B1.METRIC, 'value': world * B * K / (ms * 1e-3), 'unit': B1.UNIT, 'n_gpus': world,
                'steps': K, 'warmup': args.warmup, 'ms_per_step': ms / K, 'higher_is_better': True,
                'scaling': 'strong', 'vs_baseline': None, 'dtype': 'u8', 'data': 'synthetic',
                'config': {'workload': desc + '; %s' % mode, 'mode': mode, 'mode_choice': mode_why,
                           'zone_records': zone.n_records, 'batch_per_rank': B, 'global_batch': B * world,
                           'table_mb_per_rank': zstat['image_bytes'] / 1e6, 'parallelism': par,
                           'l2_policy': 'ring of %d distinct ingress slices per rank (%.0f MB in + answers) over a %.0f MB table'
                                        % (RING, RING * (x.size + answered_bytes) / 1e6, zstat['image_bytes'] / 1e6),
                           'timing': 'median of %d regions of %d steps (max over ranks each) after %.0f ms of warm-up (regions ms: min %.3f max %.3f)'
                                     % (B1.REGIONS, K, warm_ms, min(regions), max(regions)),
                           'output_packing': 'query order' if args.ordered else 'arrival', 'parity': parity},
                'clocks': clocks, 'e2e': e2e, 'gpu_launches': int(launches), 'roofline': roofline, 'cpu_baseline': None}
        B1.emit(line)
    dist.barrier()
    dist.destroy_process_group()
