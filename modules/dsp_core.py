"""``modules.dsp_core`` -- the import path app.py uses (app.py:13-18), served by
the B200 implementation.  Point the reference's app at this repo's root and its
``from modules.dsp_core import ...`` picks up the CUDA path unchanged."""
from dsp_audio_project_b200.dsp_core import *  # noqa: F401,F403
from dsp_audio_project_b200.dsp_core import __all__  # noqa: F401
