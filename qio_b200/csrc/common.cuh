// Shared host/device helpers for the qio_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/qio_b200.h"

namespace qio {

void set_error(const std::string &msg);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);
int arg_fail(const char *msg);

#define QIO_CUDA(call)                                                     \
    do {                                                                   \
        cudaError_t _e = (call);                                           \
        if (_e != cudaSuccess) return ::qio::cuda_fail(_e, #call, __FILE__, __LINE__); \
    } while (0)

#define QIO_LAUNCH_CHECK()                                                 \
    do {                                                                   \
        cudaError_t _e = cudaGetLastError();                               \
        if (_e != cudaSuccess) return ::qio::cuda_fail(_e, "kernel launch", __FILE__, __LINE__); \
    } while (0)

#define QIO_REQUIRE(cond, msg)                                             \
    do {                                                                   \
        if (!(cond)) return ::qio::arg_fail(msg);                          \
    } while (0)

static inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

int sm_count();  // multiprocessor count of the current device (148 on B200); queried per call, nothing cached

}  // namespace qio
