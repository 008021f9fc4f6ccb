"""CPU suite: the row-distributed PCG algorithm of csrc/pcg2.cuh's multi-rank mode (own-row products, rank-ordered partial sums, parity-buffered z,
exchanged restriction partials, constant and linear prolongation), emulated with 2-8 virtual ranks in numpy (tools/emulate_pcg_dist.py),
reproduces serial two-level PCG: same iteration count up to rounding, same solution.  This checks the algorithm the kernel implements —
every rank taking the same branches from rank-ordered sums, no read of a buffer another rank may still be writing in the emulated
schedule — not the kernel itself; that runs in tools/multirank_check.py and in the parity block of every multi-rank bench line."""
import importlib.util
import os
import warnings

import pytest

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tools", "emulate_pcg_dist.py")


@pytest.fixture(scope="module")
def emu():
    spec = importlib.util.spec_from_file_location("emulate_pcg_dist", _PATH)
    m = importlib.util.module_from_spec(spec)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        spec.loader.exec_module(m)
    return m


@pytest.mark.parametrize("case", range(5))
def test_distributed_pcg_reproduces_serial(emu, case):
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        r = emu.run_case(*emu.CASES[case])
    assert abs(r["dist_its"] - r["serial_its"]) <= 3 and r["rel_err"] < 1e-9 and r["residual"] < 1e-9
