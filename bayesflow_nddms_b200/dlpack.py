"""DLPack hand-off of device-resident batches (ddm_last_output_dlpack).

``DeviceBatch`` implements the Python array-API DLPack protocol (``__dlpack__`` /
``__dlpack_device__``), so ``torch.from_dlpack(batch)`` and
``tf.experimental.dlpack.from_dlpack(batch.__dlpack__())`` both take the buffer with
no copy.  The capsule follows the DLPack convention: named ``"dltensor"``, renamed to
``"used_dltensor"`` by the consumer, which then owns the deleter call.
"""
from __future__ import annotations

import ctypes as C

from . import _capi

_DLTENSOR = b"dltensor"
_USED = b"used_dltensor"

_py = C.pythonapi
_py.PyCapsule_New.restype = C.py_object
_py.PyCapsule_New.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
_py.PyCapsule_IsValid.restype = C.c_int
_py.PyCapsule_IsValid.argtypes = [C.c_void_p, C.c_char_p]
_py.PyCapsule_GetPointer.restype = C.c_void_p
_py.PyCapsule_GetPointer.argtypes = [C.c_void_p, C.c_char_p]


@C.CFUNCTYPE(None, C.c_void_p)
def _capsule_destructor(capsule):
    # Only an unconsumed capsule still owns the tensor.
    if _py.PyCapsule_IsValid(capsule, _DLTENSOR):
        ptr = _py.PyCapsule_GetPointer(capsule, _DLTENSOR)
        managed = C.cast(ptr, C.POINTER(_capi.DLManagedTensor))
        managed.contents.deleter(managed)


class DeviceBatch:
    """A simulated batch living in HBM: shape (B, N, 2) (or (n, 2)), f32 or f64.

    One-shot: the first ``__dlpack__`` call moves ownership into the capsule.
    """

    def __init__(self, managed_ptr, shape, dtype_bits: int, device_id: int):
        self._managed = managed_ptr  # POINTER(DLManagedTensor) or None once exported
        self.shape = tuple(shape)
        self.dtype_bits = dtype_bits
        self.device_id = device_id

    def __dlpack_device__(self):
        return (2, self.device_id)  # kDLCUDA

    def __dlpack__(self, stream=None, **_):
        # The producer stream was synchronised by ddm_last_output_dlpack, so any consumer
        # stream may read the buffer immediately.
        if self._managed is None:
            raise RuntimeError("this DeviceBatch was already exported through DLPack")
        ptr = C.cast(self._managed, C.c_void_p)
        self._managed = None
        return _py.PyCapsule_New(ptr, _DLTENSOR, C.cast(_capsule_destructor, C.c_void_p))

    def to_torch(self):
        import torch

        return torch.from_dlpack(self)

    def to_tensorflow(self):
        """The batch as a TensorFlow tensor on the same GPU, no copy (BayesFlow 1.1 runs on TensorFlow;
        TensorFlow is not part of this image, so this path is exercised only where it is installed)."""
        import tensorflow as tf

        return tf.experimental.dlpack.from_dlpack(self.__dlpack__())

    def __del__(self):
        m, self._managed = getattr(self, "_managed", None), None
        if m is not None:
            try:
                m.contents.deleter(m)
            except Exception:
                pass
