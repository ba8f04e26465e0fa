"""permafrost-engine_b200: B200 (sm_100a) implementation of permafrost-engine's navigation +
crowd-movement hot path behind a C ABI (include/pfnav.h).  See DESIGN.md."""
from . import build, capi, synth  # noqa: F401
