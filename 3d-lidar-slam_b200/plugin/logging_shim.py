"""rclpy.logging when ROS 2 is present, the standard library otherwise — the plugin
must import without ROS (the reference imports rclpy at module scope,
icp_processor.py:11, which is why it cannot be imported here)."""
import logging


def get_logger(name: str):
    try:
        import rclpy.logging  # type: ignore

        return rclpy.logging.get_logger(name)
    except Exception:
        logger = logging.getLogger(name)
        if not logging.getLogger().handlers:
            logging.basicConfig(level=logging.WARNING)
        return logger
