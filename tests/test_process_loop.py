"""the host loop fix-ups (batch dequeue, non-blocking latest-map hand-off), on a fake
processor — no GPU.  The reference loop (process_pc_manager.py:44-54) is restated next to it to show the
stall it removes."""
import multiprocessing
import queue
import threading
import time

import numpy as np

from lidar_slam_b200.plugin import DataTransfer
from lidar_slam_b200.plugin.process_pc_manager import ProcessLoop, drain_pending, publish_latest


class FakeProcessor:
    def __init__(self):
        self.seen = []
        self.exports = 0
        self.resets = 0

    def process_records(self, scan):
        self.seen.append(int(scan[0]))

    @property
    def global_map(self):
        self.exports += 1  # stands for the O(map) device->host copy
        return np.full((3, 3), float(len(self.seen)), np.float32)

    def reset(self):
        self.resets += 1
        self.seen.clear()


def _wait(cond, timeout=10.0):
    t0 = time.time()
    while not cond():
        if time.time() - t0 > timeout:
            return False
        time.sleep(0.005)
    return True


def test_drain_pending_batches_without_blocking():
    q = multiprocessing.Queue()
    assert drain_pending(q, 8, first_timeout=0.01) == []
    for i in range(5):
        q.put(i)
    assert _wait(lambda: not q.empty())
    time.sleep(0.05)  # let the feeder thread flush all five
    assert drain_pending(q, 3, 0.5) == [0, 1, 2]
    assert drain_pending(q, 8, 0.5) == [3, 4]


def test_drain_pending_batches_large_records_still_in_the_feeder_pipe():
    """A scan record (49,152 x 3 float32 = 590 KB) is larger than the queue's pipe, so the records of a backlog
    are never all 'already there': the dequeue must linger for the feeder thread (the round-1 version used
    get_nowait() for items 2..n and never batched on a fast box)."""
    q = multiprocessing.Queue()
    recs = [np.full((49152, 3), float(i), np.float32) for i in range(6)]
    for r in recs:
        q.put(r)
    got = drain_pending(q, 4, first_timeout=2.0, linger=0.5)
    assert [int(g[0, 0]) for g in got] == [0, 1, 2, 3]
    got = drain_pending(q, 4, first_timeout=2.0, linger=0.05)
    assert [int(g[0, 0]) for g in got] == [4, 5]
    t0 = time.time()
    assert drain_pending(q, 4, first_timeout=0.01, linger=0.5) == []  # nothing queued: only first_timeout is spent
    assert time.time() - t0 < 0.3


def test_publish_latest_replaces_stale_map_and_never_blocks():
    dt = DataTransfer()
    t0 = time.time()
    for k in range(20):  # nobody consumes: the reference's put() would block on the second map, lock held
        publish_latest(dt, np.full(2, k))
        time.sleep(0.002)
    assert time.time() - t0 < 2.0
    assert dt.global_map_lock.acquire(timeout=1.0)  # the lock is free for the publisher
    dt.global_map_lock.release()
    got = dt.global_map_queue.get(timeout=1.0)
    assert got[0] >= 12  # a recent map: a refresh is skipped (never waited for) while the feeder thread is mid-flush
    assert dt.global_map_queue.empty()


def test_loop_registers_every_scan_in_order_and_exports_once_per_batch():
    dt = DataTransfer()
    proc = FakeProcessor()
    stop, reset = multiprocessing.Event(), multiprocessing.Event()
    loop = ProcessLoop(proc, dt, stop, reset, max_batch=16, idle_timeout=0.02)
    th = threading.Thread(target=loop.run, daemon=True)
    for i in range(100):  # a backlog, as after a burst from the subscriber
        dt.pixel_depth_map_queue.put(np.array([i, 0, 0], np.float32))
    time.sleep(0.3)  # let the queue's feeder thread flush the backlog into the pipe
    th.start()
    assert _wait(lambda: loop.scans_processed == 100)
    assert proc.seen == list(range(100))
    assert proc.exports == loop.batches <= 100  # one map materialisation per batch, not per scan
    assert loop.batches < 100                   # and the 100 queued scans did arrive in batches
    # reset: processor.reset() is called while the event is up; scans sent afterwards are registered again
    reset.set()
    assert _wait(lambda: proc.resets > 0)
    reset.clear()
    dt.pixel_depth_map_queue.put(np.array([7, 0, 0], np.float32))
    assert _wait(lambda: proc.seen == [7])
    latest = None
    def newest():
        nonlocal latest
        try:
            latest = dt.global_map_queue.get_nowait()
        except queue.Empty:
            pass
        return latest is not None and latest[0, 0] == 1.0
    assert _wait(newest)
    stop.set()
    th.join(timeout=5.0)
    assert not th.is_alive()


def test_reference_loop_shape_stalls_where_the_fixed_one_does_not():
    """process_pc_manager.py:53-54 and output_handler.py:114-122 in shape: the producer put()s on the depth-1
    queue with the lock held; the publisher needs the same lock to take the map out."""
    dt = DataTransfer()
    second_put_ok = threading.Event()
    taken_while_blocked = []

    def reference_producer():
        for k in range(2):
            with dt.global_map_lock:
                try:
                    dt.global_map_queue.put(np.full(2, k), timeout=0.5)  # the reference passes no timeout
                except queue.Full:
                    return
        second_put_ok.set()

    def reference_publisher():
        time.sleep(0.1)  # its 10 Hz timer fires while the producer is inside the second put
        t0 = time.time()
        with dt.global_map_lock:
            waited = time.time() - t0
            if not dt.global_map_queue.empty():
                taken_while_blocked.append((waited, dt.global_map_queue.get_nowait()))

    a = threading.Thread(target=reference_producer, daemon=True)
    b = threading.Thread(target=reference_publisher, daemon=True)
    a.start()
    b.start()
    a.join(timeout=5.0)
    b.join(timeout=5.0)
    assert not second_put_ok.is_set()           # the second map was never handed over
    assert taken_while_blocked[0][0] > 0.2      # and the publisher sat on the lock until the put gave up
    # the fixed hand-off under the same schedule: both maps go through, nobody waits
    dt2 = DataTransfer()
    t0 = time.time()
    publish_latest(dt2, np.full(2, 0))
    time.sleep(0.05)
    publish_latest(dt2, np.full(2, 1))
    assert time.time() - t0 < 0.3
    assert dt2.global_map_queue.get(timeout=1.0)[0] in (0, 1)
