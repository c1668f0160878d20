"""IPC container the processor reads from, mirroring src/slam/scripts/data_transfer.py:4-25:
an unbounded input queue of scans, a depth-1 output queue for the map and the lock guarding it."""
import multiprocessing


class DataTransfer:
    def __init__(self):
        self.pixel_depth_map_queue = multiprocessing.Queue()
        self.global_map_queue = multiprocessing.Queue(1)
        self.global_map_lock = multiprocessing.Lock()

    def reset(self):
        while not self.pixel_depth_map_queue.empty():
            self.pixel_depth_map_queue.get()
        with self.global_map_lock:
            if not self.global_map_queue.empty():
                self.global_map_queue.get()
