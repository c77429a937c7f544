"""Import the real reference ``modules/dsp_core.py`` (read-only, this container
only) so the oracle can be pinned against it.  TEST INFRASTRUCTURE ONLY.

``/root/reference`` does not exist on the GPU box; everything that runs there
uses the committed vectors under ``tests/golden`` instead.  The reference's
``import soundfile`` (dsp_core.py:2) fails here because the wheel is absent, so
a stub module is registered first -- only ``cargar_senal_audio`` touches it.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("DSPB200_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "modules", "dsp_core.py"))


def load_reference_dsp_core():
    """Return the reference's dsp_core module object (not cached in
    ``sys.modules`` under its own name, so it cannot shadow this repo's
    drop-in ``modules.dsp_core``)."""
    if not reference_available():
        raise FileNotFoundError(f"reference not present under {REFERENCE_ROOT}")
    if "soundfile" not in sys.modules:
        stub = types.ModuleType("soundfile")

        def _no_read(*_a, **_k):
            raise RuntimeError("soundfile is not installed (stub)")

        stub.read = _no_read
        sys.modules["soundfile"] = stub
    path = os.path.join(REFERENCE_ROOT, "modules", "dsp_core.py")
    spec = importlib.util.spec_from_file_location("_reference_dsp_core", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
