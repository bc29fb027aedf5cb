"""World-size-2 CPU test (gloo) of the host-side logic of the sharded path: the shard-aware zone
builder partitions the key space exactly as the host (numpy) owner function predicts, the two
shards together hold every key exactly once, and the handle-exchange plumbing (all_gather_object)
delivers rank-ordered blobs."""
import os

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from binder_b200 import build, synth
from binder_b200.engine import Zone
from binder_b200.shard import hash_keys, owner_of


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    z = synth.gen_zone(30000)
    st = Zone(z.jsonl, z.dns_domain, world, rank).stat()
    # the host-side owner function on the same keys
    names = [synth.host_name(i).encode() for i in range(z.n_hosts)]
    own = owner_of(hash_keys(names, 0, z.dns_domain), world)
    addrs = [synth.host_addr(i).encode() for i in range(z.n_hosts)]
    by_len = {}
    for a in addrs:
        by_len.setdefault(len(a), []).append(a)
    rev_mine = sum(int((owner_of(hash_keys(v, 1), world) == rank).sum()) for v in by_len.values())
    mine = int((own == rank).sum())
    gathered = [None] * world
    dist.all_gather_object(gathered, (rank, st['forward_keys'], st['reverse_keys'], mine, rev_mine, bytes([rank]) * 64))
    q.put((rank, st, mine, rev_mine, gathered))
    dist.destroy_process_group()


def test_shards_partition_the_zone():
    build.build()
    z = synth.gen_zone(30000)
    full = Zone(z.jsonl, z.dns_domain).stat()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29000 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in procs)
    [p.join(60) for p in procs]
    assert sum(r[1]['forward_keys'] for r in res) == full['forward_keys']
    assert sum(r[1]['reverse_keys'] for r in res) == full['reverse_keys']
    others = full['forward_keys'] - z.n_hosts            # root, group nodes, db records: not in `names`
    for rank, st, mine, rev_mine, gathered in res:
        assert st['nodes'] == full['nodes']               # every rank mirrors the whole tree
        assert 0 <= st['forward_keys'] - mine <= others
        assert st['reverse_keys'] == rev_mine
        assert [g[0] for g in gathered] == [0, 1] and gathered[1][5] == b'\x01' * 64
        assert abs(mine - z.n_hosts / 2) < 0.05 * z.n_hosts
