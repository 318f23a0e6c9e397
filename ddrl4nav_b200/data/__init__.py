from .easybytes import DeviceEasyBytes
from .experience import Experience

__all__ = ["Experience", "DeviceEasyBytes"]
