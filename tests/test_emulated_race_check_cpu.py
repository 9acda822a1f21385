"""Race check of the kernels (oracle/emu/race_check.py): the library's own kernels, emulated with one OS thread per CUDA
thread and barrier-based __syncthreads / __syncwarp / __shfl_*_sync, under ThreadSanitizer.  A shared-memory hand-over
between lanes without the barrier the CUDA model requires shows up as a data race even where the hardware would get away
with it.  Skipped where libtsan is not installed."""
import importlib.util
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("race_check", os.path.join(os.path.dirname(HERE), "oracle", "emu", "race_check.py"))
race_check = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(race_check)

pytestmark = pytest.mark.skipif(race_check.libtsan() is None, reason="libtsan.so not available")


def test_kernels_are_race_free_under_thread_sanitizer(oracle):
    """Generic, compile-time specialised and TMA-staged step kernels (16 / 32 lanes, multi-state blocks, the bulk-copy +
    mbarrier pipeline), block-prep, finalize (Gibbs draws, adaptCov), scale
    adaptation, Xbetas, survival-posterior, the accuracy kernels, the posterior-summaries kernel and the survfitJM kernels
    (mode search, sampler, horizon integrals): no report."""
    n, text = race_check.run()
    assert n == 0, text[:4000]


def test_race_check_detects_a_missing_syncwarp(oracle):
    """Negative control: without the __syncwarp between the state load into shared memory and its first use by the other
    lanes of the group, the same run reports races."""
    n, _ = race_check.run(drop_syncwarp=True, names=("config2",), extra=False)
    assert n > 0
