"""dsp_audio_project_b200 -- B200-native (sm_100a) implementation of the numeric
hot path of Renatovela-ctrl/dsp-audio-project (modules/dsp_core.py): polyphase
sample-rate conversion, six-band biquad equaliser and radix-2 FFT magnitude
spectrum, batched over channels, behind the reference's own function names.

* ``dsp_core``  -- drop-in module (same 8 names/signatures as the reference)
* ``plans``     -- batched plan API on torch CUDA tensors / numpy host arrays
* ``shard``     -- channel partitioning across GPUs (+ optional spectra gather)
* ``_lib``      -- ctypes binding of ``libdspb200.so`` (include/dspb200.h)

The CUDA library is the only compute path; nothing here falls back to the CPU.
"""
from . import _lib  # noqa: F401
from .plans import (Chain, EqPlan, FftPlan, SrcPlan, WaveScheduler, generate_uniform, mono_normalize,  # noqa: F401
                    select_sections, src_geometry, to_pcm16)
from .shard import channel_block, gather_spectra  # noqa: F401

__version__ = "0.1.0"
