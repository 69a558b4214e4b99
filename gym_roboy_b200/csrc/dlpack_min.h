// Minimal DLPack (v0.x "dltensor" capsule ABI) declarations -- just the structs needed to hand a
// handle-owned HBM buffer to torch.from_dlpack() without copying.  Layout per the public DLPack
// specification (dmlc/dlpack, dlpack.h); declared here because the header is not in this image.
#pragma once
#include <stdint.h>

extern "C" {

typedef enum { kDLCPU = 1, kDLCUDA = 2, kDLCUDAHost = 3 } DLDeviceType;

typedef struct {
    int32_t device_type;  // DLDeviceType
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
    int64_t *strides;  // NULL: compact row-major
    uint64_t byte_offset;
} DLTensor;

typedef struct DLManagedTensor {
    DLTensor dl_tensor;
    void *manager_ctx;
    void (*deleter)(struct DLManagedTensor *self);
} DLManagedTensor;

}  // extern "C"
