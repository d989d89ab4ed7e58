"""B200-native drift-diffusion trial simulator: drop-in for the data-generation hot path of
mdnunez/bayesflow_nddms.  Hand-written sm_100a CUDA behind a C ABI
(include/ddm_b200.h), called through ctypes; no CPU fallback."""
from . import _capi  # noqa: F401
from .simulator import DDMError, DDMSimulator, default_simulator, set_default_simulator  # noqa: F401
from .dlpack import DeviceBatch  # noqa: F401

__version__ = "0.1.0"
