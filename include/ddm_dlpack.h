/*
 * ddm_dlpack.h -- the DLPack v0.8 C ABI structs (layout fixed by the DLPack
 * standard, https://dmlc.github.io/dlpack).  Declared here so that the
 * library has no external header dependency; guarded so that including the
 * upstream dlpack.h first is harmless.
 */
#ifndef DDM_DLPACK_H
#define DDM_DLPACK_H
#ifndef DLPACK_DLPACK_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum { kDLCPU = 1, kDLCUDA = 2, kDLCUDAHost = 3 } DLDeviceType;

typedef struct {
    DLDeviceType device_type;
    int32_t device_id;
} DLDevice;

typedef enum { kDLInt = 0U, kDLUInt = 1U, kDLFloat = 2U } DLDataTypeCode;

typedef struct {
    uint8_t code;
    uint8_t bits;
    uint16_t lanes;
} DLDataType;

typedef struct {
    void *data;
    DLDevice device;
    int32_t ndim;
    DLDataType dtype;
    int64_t *shape;
    int64_t *strides; /* NULL = compact row-major */
    uint64_t byte_offset;
} DLTensor;

typedef struct DLManagedTensor {
    DLTensor dl_tensor;
    void *manager_ctx;
    void (*deleter)(struct DLManagedTensor *self);
} DLManagedTensor;

#ifdef __cplusplus
}
#endif
#endif /* DLPACK_DLPACK_H_ */
#endif /* DDM_DLPACK_H */
