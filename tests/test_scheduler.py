"""Host-side tests of the chip-proof lane scheduler (policy of ChipScheduler::execute,
reference ceno_zkvm/src/scheme/scheduler.rs:109-400 and its unit tests :778-830).  No GPU needed: with ctx = NULL the
scheduler runs host-only callbacks (stream = NULL)."""
import threading
import time

import pytest

import ceno_b200 as cb
from ceno_b200 import _lib

MB = 1 << 20


def _sched():
    from ceno_b200 import build as cbuild
    cbuild.build()
    return cb.ChipScheduler(None)


def test_empty_task_list_is_a_noop():
    assert _sched().execute([], lambda t, lane, s: None, mem_budget_bytes=MB) == ([], [])


def test_big_rocks_first_and_results_sorted_by_task_id():
    sizes = [3, 9, 1, 7, 5]
    tasks = [cb.ChipTask(i, s * MB, payload=s) for i, s in enumerate(sizes)]
    order = []
    out, tel = _sched().execute(tasks, lambda t, lane, s: (order.append(t.payload), t.payload * 2)[1], lanes=1, mem_budget_bytes=100 * MB)
    assert order == sorted(sizes, reverse=True)                 # one lane: strictly by memory, descending
    assert out == [2 * s for s in sizes]                        # outputs come back in task_id order
    assert [t["task_id"] for t in tel] == list(range(len(sizes)))
    assert all(t["status"] == 0 and t["lane_id"] == 0 for t in tel)
    assert sorted(t["launch_seq"] for t in tel) == list(range(len(sizes)))


def test_memory_budget_and_lane_limit_are_never_exceeded():
    import random
    rng = random.Random(7)
    sizes = [rng.randint(1, 40) for _ in range(40)]
    tasks = [cb.ChipTask(i, s * MB, payload=s) for i, s in enumerate(sizes)]
    lock = threading.Lock()
    state = {"mem": 0, "run": 0, "max_mem": 0, "max_run": 0}

    def work(t, lane, stream):
        assert stream is None
        with lock:
            state["mem"] += t.payload
            state["run"] += 1
            state["max_mem"] = max(state["max_mem"], state["mem"])
            state["max_run"] = max(state["max_run"], state["run"])
        time.sleep(0.002)
        with lock:
            state["mem"] -= t.payload
            state["run"] -= 1
        return lane

    out, tel = _sched().execute(tasks, work, lanes=3, mem_budget_bytes=64 * MB)
    assert state["max_mem"] <= 64 and 1 < state["max_run"] <= 3
    assert all(0 <= lane < 3 for lane in out)
    assert all(t["booked_total_at_launch"] <= 64 * MB for t in tel)


def test_backfilling_skips_tasks_that_do_not_fit():
    # budget 10: 8 is admitted, 6 and 3 do not fit beside it, 2 does (8 + 2 = 10) -> admitted second
    sizes = {0: 6, 1: 2, 2: 8, 3: 3}
    tasks = [cb.ChipTask(i, s * MB) for i, s in sizes.items()]
    gate = threading.Event()
    started = []

    def work(t, lane, stream):
        started.append(t.task_id)
        if t.task_id in (2, 1):
            gate.wait(5)
        return t.task_id

    th = threading.Timer(0.2, gate.set)
    th.start()
    _, tel = _sched().execute(tasks, work, lanes=4, mem_budget_bytes=10 * MB)
    th.join()
    seq = {t["task_id"]: t["launch_seq"] for t in tel}
    assert seq[2] == 0 and seq[1] == 1            # 8 MB first, then the 2 MB task backfills
    assert {seq[0], seq[3]} == {2, 3}             # 6 MB and 3 MB wait for a completion
    assert set(started[:2]) == {2, 1}


def test_deadlock_when_a_task_can_never_fit():
    tasks = [cb.ChipTask(0, 5 * MB), cb.ChipTask(1, 50 * MB), cb.ChipTask(2, 1 * MB)]
    ran = []
    with pytest.raises(cb.CenoB200Error) as e:
        _sched().execute(tasks, lambda t, lane, s: ran.append(t.task_id), lanes=2, mem_budget_bytes=10 * MB)
    assert e.value.code == 4                      # CG_ERR_OOM: "Deadlock: Remaining tasks are too big for the memory pool"
    assert sorted(ran) == [0, 2]                  # everything that fits has run before the error is reported


def test_lane_configuration_accepts_only_one_through_eight():
    tasks = [cb.ChipTask(0, MB)]
    for lanes in (1, 4, 8):
        _sched().execute(tasks, lambda t, lane, s: None, lanes=lanes, mem_budget_bytes=2 * MB)
    with pytest.raises(cb.CenoB200Error) as e:
        _sched().execute(tasks, lambda t, lane, s: None, lanes=9, mem_budget_bytes=2 * MB)
    assert e.value.code == 2


def test_first_task_error_is_returned_after_inflight_tasks_drain():
    tasks = [cb.ChipTask(i, (10 - i) * MB) for i in range(6)]
    finished = []

    def work(t, lane, stream):
        if t.task_id == 1:
            raise ValueError("chip 1 failed")
        time.sleep(0.05)
        finished.append(t.task_id)

    with pytest.raises(ValueError, match="chip 1 failed"):
        _sched().execute(tasks, work, lanes=2, mem_budget_bytes=100 * MB)
    assert 0 in finished                          # the task running beside the failing one completed
    assert len(finished) < 5                      # nothing new was admitted after the failure


def test_booked_memory_overrides_the_estimate_for_admission():
    # estimates fit together, bookings (with a concurrency margin) do not -> the two tasks never overlap
    tasks = [cb.ChipTask(0, 4 * MB, booked_memory_bytes=7 * MB), cb.ChipTask(1, 3 * MB, booked_memory_bytes=6 * MB)]
    lock = threading.Lock()
    state = {"run": 0, "max_run": 0}

    def work(t, lane, stream):
        with lock:
            state["run"] += 1
            state["max_run"] = max(state["max_run"], state["run"])
        time.sleep(0.01)
        with lock:
            state["run"] -= 1

    _sched().execute(tasks, work, lanes=4, mem_budget_bytes=10 * MB)
    assert state["max_run"] == 1
    assert _lib.CG_OK == 0
