// Error reporting and device queries of the C ABI (include/qio_b200.h).
#include "common.cuh"

namespace qio {

static thread_local std::string g_last_error;

void set_error(const std::string &msg) { g_last_error = msg; }

int cuda_fail(cudaError_t e, const char *what, const char *file, int line) {
    char buf[512];
    snprintf(buf, sizeof(buf), "CUDA error %d (%s) in %s at %s:%d", (int)e, cudaGetErrorString(e), what, file, line);
    g_last_error = buf;
    return QIO_B200_ERR_CUDA;
}

int arg_fail(const char *msg) {
    g_last_error = std::string("invalid argument: ") + msg;
    return QIO_B200_ERR_ARG;
}

int sm_count() {
    // queried per call for the CURRENT device: no process-wide cache (the attribute query costs well under a microsecond)
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0)
        return sms;
    return 148;
}

}  // namespace qio

extern "C" {

const char *qio_b200_last_error(void) { return qio::g_last_error.c_str(); }

int qio_b200_abi_version(void) { return QIO_B200_ABI_VERSION; }

int qio_b200_device_info(int *sm_count, int *cc_major, int *cc_minor) {
    int dev = 0;
    QIO_CUDA(cudaGetDevice(&dev));
    int v = 0;
    if (sm_count) {
        QIO_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
        *sm_count = v;
    }
    if (cc_major) {
        QIO_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev));
        *cc_major = v;
    }
    if (cc_minor) {
        QIO_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev));
        *cc_minor = v;
    }
    return QIO_B200_OK;
}

}  // extern "C"
