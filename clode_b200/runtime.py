"""Device selection and logging helpers — the API of the reference's clode/runtime.py:22-91.

`platform_id` 0 is the CUDA driver; `device_id` / `device_ids` are CUDA device ordinals.  Passing
`device_ids=[0, 1, ...]` shards an ensemble over several GPUs of the box (the reference accepts the
argument but only ever uses the first device, clode/cpp/OpenCLResource.hpp:104,110).
"""
from __future__ import annotations

import os

from .cpp.clode_cpp_wrapper import (CLDeviceType, CLVendor, DeviceInfo, LogLevel, OpenCLResource, PlatformInfo,
                                    _print_opencl, get_logger, query_opencl)

# the reference reads its kernel sources from this directory at construction (clode/runtime.py:17);
# here the engine sources are embedded in libclode_rt.so and the argument is accepted and ignored
_clode_root_dir: str = os.path.join(os.path.dirname(__file__), "csrc", "device", "")

DEFAULT_LOG_LEVEL = LogLevel.warn


def initialize_runtime(device_type, vendor, platform_id, device_id, device_ids) -> OpenCLResource:
    """Validate the selection arguments exactly as clode/runtime.py:22-52 does and build the resource."""
    by_id = platform_id is not None
    if by_id:
        if device_type is not None:
            raise ValueError("Cannot specify device_type when platform_id is specified")
        if vendor is not None:
            raise ValueError("Cannot specify vendor when platform_id is specified")
        if device_id is not None and device_ids is not None:
            raise ValueError("Cannot specify both device_id and device_ids")
        if device_id is None and device_ids is None:
            raise ValueError("Must specify one of device_id and device_ids")
        return OpenCLResource(platform_id, device_id) if device_id is not None else OpenCLResource(platform_id, list(device_ids))
    if device_id is not None:
        raise ValueError("Must specify platform_id when specifying device_id")
    if device_ids is not None:
        raise ValueError("Must specify platform_id when specifying device_ids")
    return OpenCLResource(CLDeviceType.DEVICE_TYPE_DEFAULT if device_type is None else device_type,
                          CLVendor.VENDOR_ANY if vendor is None else vendor)


def get_log_level() -> LogLevel:
    return get_logger().get_log_level()


def set_log_level(level: LogLevel) -> None:
    get_logger().set_log_level(level)


def set_log_pattern(pattern: str) -> None:
    get_logger().set_log_pattern(pattern)


set_log_level(DEFAULT_LOG_LEVEL)


def print_opencl():
    previous = get_log_level()
    set_log_level(LogLevel.info)
    _print_opencl()
    set_log_level(previous)


__all__ = ["CLDeviceType", "CLVendor", "DeviceInfo", "PlatformInfo", "OpenCLResource", "initialize_runtime",
           "print_opencl", "query_opencl", "DEFAULT_LOG_LEVEL", "LogLevel", "set_log_level", "set_log_pattern",
           "get_log_level"]
