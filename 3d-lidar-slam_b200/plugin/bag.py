"""rosbag2-shaped storage of PointCloud2 streams without ROS.

The reference records its inputs and saves its maps through ``rosbag2_py.SequentialWriter`` with
``storage_id="sqlite3"`` and ``serialization_format="cdr"`` (src/slam/slam/input_handler.py:80-112,
output_handler.py:76-105, generic_handler.py:33-44).  A rosbag2 sqlite3 bag is a directory holding
``metadata.yaml`` and a ``.db3`` file with two tables — ``topics(id, name, type, serialization_format,
offered_qos_profiles)`` and ``messages(id, topic_id, timestamp, data)`` — each message blob being the CDR
encoding of the ROS message.  ``BagWriter`` / ``BagReader`` restate that layout on Python's ``sqlite3`` and
``pointcloud2.serialize_cdr``, so recorded scans can be replayed into the GPU path (``replay_into``) and a
map exported from the device can be saved, on a machine that has no ROS installation.

rosbag2 itself is not installed here: the layout follows its published schema and is pinned only by this
repo's round-trip tests.
"""
from __future__ import annotations

import os
import sqlite3
from typing import Iterator, Optional, Tuple

from . import pointcloud2 as pc2

POINTCLOUD2_TYPE = "sensor_msgs/msg/PointCloud2"


class BagWriter:
    """Mirror of the calls the reference makes on ``rosbag2_py.SequentialWriter``: ``open`` (constructor),
    ``create_topic``, ``write(topic, serialized, timestamp_ns)``."""

    def __init__(self, uri: str):
        os.makedirs(uri, exist_ok=False)
        self.uri = uri
        self.db_name = os.path.basename(os.path.normpath(uri)) + "_0.db3"
        self.db = sqlite3.connect(os.path.join(uri, self.db_name))
        self.db.execute("CREATE TABLE topics(id INTEGER PRIMARY KEY, name TEXT NOT NULL, type TEXT NOT NULL, "
                        "serialization_format TEXT NOT NULL, offered_qos_profiles TEXT NOT NULL)")
        self.db.execute("CREATE TABLE messages(id INTEGER PRIMARY KEY, topic_id INTEGER NOT NULL, "
                        "timestamp INTEGER NOT NULL, data BLOB NOT NULL)")
        self.db.execute("CREATE INDEX timestamp_idx ON messages (timestamp ASC)")
        self._topics = {}
        self._counts = {}
        self._t0: Optional[int] = None
        self._t1: Optional[int] = None

    def create_topic(self, name: str, type: str = POINTCLOUD2_TYPE, serialization_format: str = "cdr") -> None:
        if name in self._topics:
            return
        cur = self.db.execute("INSERT INTO topics(name, type, serialization_format, offered_qos_profiles) "
                              "VALUES (?, ?, ?, '')", (name, type, serialization_format))
        self._topics[name] = (cur.lastrowid, type, serialization_format)
        self._counts[name] = 0

    def write(self, topic: str, serialized: bytes, timestamp_ns: int) -> None:
        if topic not in self._topics:
            raise KeyError(f"topic {topic!r} was not created")
        self.db.execute("INSERT INTO messages(topic_id, timestamp, data) VALUES (?, ?, ?)",
                        (self._topics[topic][0], int(timestamp_ns), sqlite3.Binary(serialized)))
        self._counts[topic] += 1
        self._t0 = timestamp_ns if self._t0 is None else min(self._t0, timestamp_ns)
        self._t1 = timestamp_ns if self._t1 is None else max(self._t1, timestamp_ns)

    def write_cloud(self, topic: str, cloud: pc2.PointCloud2, timestamp_ns: int) -> None:
        """generic_handler.save_point_cloud: serialise and write one cloud."""
        self.create_topic(topic)
        self.write(topic, pc2.serialize_cdr(cloud), timestamp_ns)

    def close(self) -> None:
        self.db.commit()
        self.db.close()
        t0, t1 = self._t0 or 0, self._t1 or 0
        lines = ["rosbag2_bagfile_information:", "  version: 5", "  storage_identifier: sqlite3",
                 "  duration:", f"    nanoseconds: {t1 - t0}", "  starting_time:",
                 f"    nanoseconds_since_epoch: {t0}", f"  message_count: {sum(self._counts.values())}",
                 "  topics_with_message_count:"]
        for name, (_, type_, fmt) in self._topics.items():
            lines += ["    - topic_metadata:", f"        name: {name}", f"        type: {type_}",
                      f"        serialization_format: {fmt}", "        offered_qos_profiles: \"\"",
                      f"      message_count: {self._counts[name]}"]
        lines += ["  compression_format: \"\"", "  compression_mode: \"\"", "  relative_file_paths:",
                  f"    - {self.db_name}", "  files:", f"    - path: {self.db_name}", "      starting_time:",
                  f"        nanoseconds_since_epoch: {t0}", "      duration:", f"        nanoseconds: {t1 - t0}",
                  f"      message_count: {sum(self._counts.values())}"]
        with open(os.path.join(self.uri, "metadata.yaml"), "w") as f:
            f.write("\n".join(lines) + "\n")

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class BagReader:
    """Sequential reader over every ``.db3`` of a bag directory (or one ``.db3`` file), in timestamp order."""

    def __init__(self, uri: str):
        if os.path.isdir(uri):
            self.files = sorted(os.path.join(uri, f) for f in os.listdir(uri) if f.endswith(".db3"))
        else:
            self.files = [uri]
        if not self.files:
            raise FileNotFoundError(f"no .db3 file under {uri}")

    def topics(self):
        out = {}
        for path in self.files:
            with sqlite3.connect(path) as db:
                for name, type_, fmt in db.execute("SELECT name, type, serialization_format FROM topics"):
                    out[name] = (type_, fmt)
        return out

    def messages(self, topic: Optional[str] = None) -> Iterator[Tuple[str, int, bytes]]:
        for path in self.files:
            db = sqlite3.connect(path)
            try:
                q = ("SELECT topics.name, messages.timestamp, messages.data FROM messages JOIN topics "
                     "ON messages.topic_id = topics.id")
                args = ()
                if topic is not None:
                    q += " WHERE topics.name = ?"
                    args = (topic,)
                for name, ts, data in db.execute(q + " ORDER BY messages.timestamp ASC, messages.id ASC", args):
                    yield name, int(ts), bytes(data)
            finally:
                db.close()

    def clouds(self, topic: str = "/input_pointcloud") -> Iterator[Tuple[int, pc2.PointCloud2]]:
        for _, ts, data in self.messages(topic):
            yield ts, pc2.deserialize_cdr(data)


def replay_into(processor, uri: str, topic: str = "/input_pointcloud", limit: Optional[int] = None) -> int:
    """Feed the recorded scans of a bag to a ``B200ICPProcessor`` without ROS: every PointCloud2 payload goes
    to the GPU as bytes (``process_cloud``), where it is decoded, projected and registered.  Returns the number
    of scans processed."""
    n = 0
    for _, cloud in BagReader(uri).clouds(topic):
        processor.process_cloud(cloud)
        n += 1
        if limit is not None and n >= limit:
            break
    return n
