"""radlite_b200 -- B200-native line ray-tracing hot path of RADLite behind a C ABI.

``Renderer`` (radlite_b200.api) wraps ``libradlite_b200.so`` (CUDA, sm_100a).  There is no CPU
fallback: constructing a Renderer without the built library or without a GPU raises.
"""
from . import synth  # noqa: F401

__all__ = ["synth"]
