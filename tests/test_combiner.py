"""Host logic of the call-coalescing front end (csrc/combiner.h) against a mock executor: every caller gets
its own answer and return code, batches respect the size cap, the compatibility classes and the in-flight
limit, a lone caller is never delayed, and concurrent callers really are merged.  CPU only."""
import ctypes as C
import os
import subprocess
import time

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(HERE), "c-kzg-4844_b200", "csrc")
OUT = os.path.join(HERE, "hostcheck", "_build", "libcombinercheck.so")


@pytest.fixture(scope="module")
def lib():
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    src = os.path.join(HERE, "hostcheck", "combiner_check.cpp")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-I", CSRC, "-o", OUT, src])
    l = C.CDLL(OUT)
    l.combiner_selftest.restype = C.c_int
    return l


def run(lib, threads, per_thread, max_batch, inflight, exec_us, fail_odd=0, classes=1):
    out = (C.c_uint64 * 4)()
    rc = lib.combiner_selftest(threads, per_thread, max_batch, inflight, exec_us, fail_odd, classes, out)
    return rc, list(out)


def test_single_caller_runs_alone_and_immediately(lib):
    t0 = time.perf_counter()
    rc, (req, batches, largest, peak) = run(lib, 1, 50, 64, 2, 100)
    assert rc == 0 and req == 50 and batches == 50 and largest == 1 and peak == 1
    assert time.perf_counter() - t0 < 1.0  # 50 x 100 us of mock work, no batching delay on top


def test_concurrent_callers_are_merged(lib):
    rc, (req, batches, largest, peak) = run(lib, 32, 20, 64, 2, 2000)
    assert rc == 0 and req == 640
    assert batches < req / 3 and largest >= 8 and peak <= 2


def test_gather_window_fills_batches_under_many_callers(lib):
    """Round 2: 64 callers, two batches in flight: a leader holds its batch open for the callers that usually arrive
    together (bounded by 1/16 of the last batch's duration), so batches average well above the 9 of pure group commit --
    and a lone caller (first test) still never waits."""
    rc, (req, batches, largest, peak) = run(lib, 64, 20, 64, 2, 3000)
    assert rc == 0 and req == 1280 and peak <= 2
    assert req / batches >= 16, (req, batches, largest)


def test_batch_cap_and_inflight_limit(lib):
    rc, (req, batches, largest, peak) = run(lib, 48, 10, 5, 1, 500)
    assert rc == 0 and req == 480 and largest <= 5 and peak == 1


def test_per_request_return_codes_survive_batching(lib):
    rc, (req, batches, largest, peak) = run(lib, 16, 25, 32, 2, 1000, fail_odd=1)
    assert rc == 0 and req == 400 and largest > 1


def test_classes_never_share_a_batch(lib):
    rc, (req, batches, largest, peak) = run(lib, 24, 20, 64, 2, 1000, classes=3)
    assert rc == 0 and req == 480


def test_executor_failure_fails_the_batch_not_the_queue(lib):
    """An executor that throws (e.g. an allocation failure while gathering a batch) must surface as an error code to
    exactly the callers of that batch; later batches keep flowing and nobody gets a silent wrong answer."""
    rc, (req, batches, largest, peak) = run(lib, 16, 30, 8, 2, 300, fail_odd=2)
    assert rc == 0 and req == 480
