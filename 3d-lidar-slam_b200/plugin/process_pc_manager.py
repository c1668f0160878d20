"""Processor factory, mirroring ProcessPointCloudsHandler.set_algorithm
(src/slam/scripts/process_pc_manager.py:78-98) with the one added branch a maintainer
adds to select the B200 processor.  The reference's process loop
(process_pc_manager.py:35-57) is host control flow and stays as it is."""
from __future__ import annotations

from .algorithm_enum import AlgorithmType


def set_algorithm(algorithm: str, config_path: str, data_transfer, reset_event, **kwargs):
    try:
        algo = AlgorithmType[algorithm.upper()]
    except KeyError as exc:  # the reference lets the KeyError escape (process_pc_manager.py:80)
        raise ValueError(f"Invalid algorithm type: {algorithm}") from exc
    if algo == AlgorithmType.B200_ICP:
        from .b200_icp_processor import B200ICPProcessor

        return B200ICPProcessor(config_path=config_path, data_transfer=data_transfer, reset_event=reset_event,
                                **kwargs)
    if algo == AlgorithmType.ICP:
        raise ValueError("AlgorithmType.ICP is the reference's Open3D processor; it is not part of this package")
    if algo == AlgorithmType.DEEP_LEARNING:
        return None  # vacant slot in the reference too (process_pc_manager.py:91-92)
    raise ValueError(f"Unsupported algorithm: {algo}")


# ----------------------------------------------------------------------------- the process loop (row f4)
import queue as _queue  # noqa: E402
import time as _time  # noqa: E402


def drain_pending(q, max_items: int, first_timeout: float = 0.1, linger: float = 0.005):
    """Batch dequeue: wait up to ``first_timeout`` for one scan, then keep taking scans (at most ``max_items``)
    for as long as the next one shows up within ``linger`` seconds.  Replaces the reference's ``empty()`` poll +
    ``sleep(0.1)`` (process_pc_manager.py:49-51), which adds up to 100 ms of latency per scan and dequeues one
    at a time.

    Why a linger and not ``get_nowait()``: a ``multiprocessing.Queue`` (data_transfer.py:13) hands items over
    through a feeder thread and a 64 KB pipe, so a 590 KB scan record that was ``put`` long ago is still not
    "already there" the instant the previous record has been read — ``get_nowait()`` raises ``Empty`` on a
    backlog and the batch never forms.  A short blocking ``get`` lets the feeder push the next record; the
    cost is at most ``linger`` of added latency at the end of a batch."""
    items = []
    try:
        items.append(q.get(timeout=first_timeout))
    except _queue.Empty:
        return items
    while len(items) < max_items:
        try:
            items.append(q.get_nowait())
            continue
        except _queue.Empty:
            pass
        if linger <= 0.0:
            break
        try:
            items.append(q.get(timeout=linger))
        except _queue.Empty:
            break
    return items


def publish_latest(data_transfer, global_map) -> bool:
    """Hand the newest map to the publisher without ever blocking.  The reference does
    ``with global_map_lock: global_map_queue.put(map)`` on a depth-1 queue (process_pc_manager.py:53-54): when
    the publisher has not collected the previous map, ``put`` blocks WITH THE LOCK HELD, and the publisher —
    which takes the same lock before ``get_nowait`` (output_handler.py:114-122) — can never drain it.  Here the
    stale map is replaced under the lock with non-blocking calls only.  Returns False when the slot could not
    be refreshed this time (the next call carries a newer map anyway)."""
    with data_transfer.global_map_lock:
        try:
            data_transfer.global_map_queue.get_nowait()  # drop the stale map, if any
        except _queue.Empty:
            pass
        try:
            data_transfer.global_map_queue.put_nowait(global_map)
            return True
        except _queue.Full:  # the dropped item is still in the feeder pipe: skip, never wait
            return False


class ProcessLoop:
    """``ProcessPointCloudsHandler.process_loop`` (process_pc_manager.py:35-57) with the two fix-ups a
    processor faster than ~10 scans/s needs: pending scans are dequeued in
    batches and registered back to back, and the map is materialised (a device->host copy of the whole map
    for ``B200ICPProcessor``) and published once per batch through ``publish_latest``.  Same collaborators
    as the reference: a ``DataTransfer``, a stop event, a reset event and a processor with ``process`` /
    ``reset`` / ``global_map``."""

    def __init__(self, processor, data_transfer, stop_event, reset_event, max_batch: int = 64,
                 idle_timeout: float = 0.1, logger=None, linger: float = 0.005):
        self.processor = processor
        self.data_transfer = data_transfer
        self.stop_event = stop_event
        self.reset_event = reset_event
        self.max_batch = int(max_batch)
        self.idle_timeout = float(idle_timeout)
        self.linger = float(linger)
        self.logger = logger
        self.scans_processed = 0
        self.batches = 0
        self.maps_published = 0

    def step(self) -> int:
        """One turn of the loop; returns the number of scans registered."""
        if self.reset_event.is_set():
            if self.logger:
                self.logger.info("Resetting processor")
            self.processor.reset()
            _time.sleep(0.001)  # the reference spins here until the event is cleared (process_pc_manager.py:44-47)
            return 0
        scans = drain_pending(self.data_transfer.pixel_depth_map_queue, self.max_batch, self.idle_timeout, self.linger)
        if not scans:
            return 0
        take = getattr(self.processor, "process_records", None)
        flush = getattr(self.processor, "flush_results", None)
        for scan in scans:
            if take is not None and flush is not None:
                take(scan, defer_result=True)  # enqueue only: the batch's poses are read back once, below
            elif take is not None:
                take(scan)
            else:  # a reference-style processor that dequeues for itself
                self.data_transfer.pixel_depth_map_queue.put(scan)
                self.processor.process()
        if take is not None and flush is not None:
            flush()
        self.scans_processed += len(scans)
        self.batches += 1
        if publish_latest(self.data_transfer, self.processor.global_map):
            self.maps_published += 1
        return len(scans)

    def run(self) -> None:
        try:
            while not self.stop_event.is_set():
                self.step()
        except KeyboardInterrupt:
            if self.logger:
                self.logger.info("Stopped by user")
