"""The CPU baseline arm "same table, same algorithm" (oracle/same_table.py: the device source compiled for the host over
the product's own zone image, many threads) must produce what the oracle produces — it is timed beside the GPU."""
import numpy as np
import pytest

import helpers as H
from binder_b200 import synth


@pytest.mark.parametrize('workload', ['config2', 'config3', 'config5'])
def test_same_table_matches_oracle_totals(workload):
    import same_table
    desc, service_frac, mix, miss_frac, recursion = synth.WORKLOADS[workload]
    z = synth.gen_zone(20000, service_frac=service_frac)
    st = same_table.SameTable(z.dns_domain, z.jsonl, recursion)
    orc = H.make_impl('oracle', z.dns_domain, z.jsonl, recursion=recursion)
    data, off, _ = synth.gen_batch(z, 5000, 3, mix, miss_frac)
    out, ooff, olen, status, miss = orc.resolve_batch(data, off, seed=77)
    for nt in (1, 3, 8):
        secs, nbytes, nmiss = st.timed_resolve(data, off, seed=77, nthreads=nt, repeat=1)
        assert secs > 0 and nbytes == len(out) and nmiss == len(miss), (nt, nbytes, len(out), nmiss, len(miss))
