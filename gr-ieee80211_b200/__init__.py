"""gr-ieee80211_b200: B200 (sm_100a) implementation of the gr-ieee80211 20 MHz OFDM receive chain
presiso -> trigger -> sync -> signal -> demod -> decode, behind a C ABI (include/c80211b200.h).

Host side only marshals buffers; all signal processing runs in the CUDA kernels of csrc/."""
from . import _cabi
from . import parallel  # noqa: F401
from . import synth  # noqa: F401
from . import flowgraph  # noqa: F401
from . import blocks  # noqa: F401
from .rx import Receiver, lut_blob  # noqa: F401
from ._cabi import C8bError, FRAME_DTYPE, TXFRAME_DTYPE, TXMU_DTYPE, K_NAMES  # noqa: F401
