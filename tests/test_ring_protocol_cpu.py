"""The hand-over protocol of query_col_kernel<P = 1> (A ring, accumulators, two issuing threads) as a discrete model:
experiments/ring_sim.py runs every warp and issuing thread as a sequential program of mbarrier waits / arrivals under
random interleavings and fails on a deadlock, an over-arrival or a wait that passes on an aliased phase."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("groups,slots", [(1, 3), (1, 4), (2, 3), (2, 4)])
def test_ring_protocol_model_completes(groups, slots):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "experiments", "ring_sim.py"), str(groups), str(slots), "6", "2"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "all runs completed" in out.stdout
