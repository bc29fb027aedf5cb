"""Sharded multi-GPU parity (needs >= 2 GPUs; skipped on a 1-GPU box): tools/multi_check.py under
torchrun compares every routed query's answer with the single-engine CPU oracle."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize('world', [2, 4, 8])
@pytest.mark.parametrize('ordered,host_results', [('0', '0'), ('1', '0'), ('0', '1')])
def test_sharded_ranks(world, ordered, host_results):
    """world ranks, one per GPU (skipped when the box has fewer): every routed query answered exactly once, by its
    owner, with the oracle's bytes.  host_results=1: the owner writes its answers straight into pinned host
    mirrors (bb_shard_host_results)."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip('needs %d GPUs' % world)
    env = dict(os.environ, BB_ORDERED=ordered, BB_HOST_RESULTS=host_results, BB_BATCH='20000', BB_ZONE='100000')
    p = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world),
                        '--master-addr', '127.0.0.1', '--master-port', str(29533 + world), os.path.join(ROOT, 'tools', 'multi_check.py')],
                       env=env, capture_output=True, text=True, timeout=900)
    assert 'MULTI_CHECK_OK world=%d' % world in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]


@pytest.mark.gpu
@pytest.mark.parametrize('host_results', ['0', '1'])
def test_sharded_single_rank(host_results):
    """world=1 on one GPU: the same route+push -> region resolve path (multi-region kernel instantiation,
    device-side batch size and shuffle indices), with device result buffers and with host mirrors."""
    env = dict(os.environ, BB_ORDERED='0', BB_HOST_RESULTS=host_results, BB_BATCH='20000', BB_ZONE='100000')
    p = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '1',
                        '--master-addr', '127.0.0.1', '--master-port', '29534', os.path.join(ROOT, 'tools', 'multi_check.py')],
                       env=env, capture_output=True, text=True, timeout=600)
    assert 'MULTI_CHECK_OK world=1' in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]


@pytest.mark.gpu
@pytest.mark.parametrize('world', [2, 8])
def test_sharded_ranks_collective_exchange(world):
    """The same parity check with the routed records exchanged by ONE grouped ncclSend/ncclRecv all-to-all
    (bb_shard_use_exchange_buffers, sync='nccl_a2a') instead of peer stores."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip('needs %d GPUs' % world)
    env = dict(os.environ, BB_ORDERED='0', BB_HOST_RESULTS='0', BB_BATCH='20000', BB_ZONE='100000', BB_SYNC='nccl_a2a')
    p = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', str(world),
                        '--master-addr', '127.0.0.1', '--master-port', str(29553 + world), os.path.join(ROOT, 'tools', 'multi_check.py')],
                       env=env, capture_output=True, text=True, timeout=900)
    assert 'MULTI_CHECK_OK world=%d' % world in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]
