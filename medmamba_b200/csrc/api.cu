// Host-only entry points of the C ABI (no CUDA calls).
#include "common.cuh"

#ifndef MMB_SOURCE_DIGEST
#define MMB_SOURCE_DIGEST "unknown"
#endif

extern "C" int mmb_abi_version(void) { return MMB_ABI_VERSION; }

extern "C" const char* mmb_source_digest(void) { return MMB_SOURCE_DIGEST; }

extern "C" const char* mmb_status_string(int status) {
    if (status == MMB_OK) return "ok";
    if (status == MMB_ERR_INVALID_ARG) return "invalid argument (null pointer, non-positive size or inconsistent shape)";
    if (status == MMB_ERR_UNSUPPORTED) return "unsupported shape, stride or dtype for this kernel";
    if (status <= MMB_ERR_CUDA_BASE) return cudaGetErrorString((cudaError_t)(MMB_ERR_CUDA_BASE - status));
    return "unknown status";
}
