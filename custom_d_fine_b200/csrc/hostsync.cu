// Host -> stream hand-off without a host-side wait on the critical path (custom_d_fine_b200/train.py GraphedTrainStep).
//
// A training step has one host-side piece between its two big CUDA graphs: the matcher's index table comes back (D2H), the
// host forms the GO union / normalisers (dfine_criterion.py:570-652) and sends ONE index table down again.  Launching
// graph B only after that planning left the device idle for the planning time PLUS the launch latency of a ~1500-node
// graph.  With a stream memory-wait the host enqueues "wait for flag >= step", the H2D copies of the (pinned) table and
// graph B right behind graph A, then plans, then raises the flag: the device resumes the instant the table is ready.
//
//   dfine_flag_create      a 4-byte flag in mapped pinned host memory (host pointer + the device alias of the same word)
//   dfine_stream_wait_flag cuStreamWaitValue32(stream, flag, value, GEQ) — cyclic comparison, so a step counter never resets
//   dfine_flag_destroy
#include <cuda.h>

#include <mutex>

#include "common.cuh"

namespace {
typedef CUresult (*WaitValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
WaitValue32Fn get_wait_value() {
    static WaitValue32Fn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (WaitValue32Fn)p;
    });
    return fn;
}
}  // namespace

// 1 if the driver exposes stream memory operations (every device of compute capability >= 7.0 under CUDA 12).
DFINE_API int dfine_stream_wait_supported(void) { return get_wait_value() != nullptr ? 1 : 0; }

DFINE_API int dfine_flag_create(void** host_ptr, void** dev_ptr) {
    DFINE_REQUIRE(host_ptr && dev_ptr, "flag_create: null output");
    void* h = nullptr;
    cudaError_t e = cudaHostAlloc(&h, 64, cudaHostAllocMapped | cudaHostAllocPortable);
    if (e != cudaSuccess) { dfine_set_error("flag_create: cudaHostAlloc: %s", cudaGetErrorString(e)); return -2; }
    *reinterpret_cast<volatile unsigned int*>(h) = 0u;
    void* d = nullptr;
    e = cudaHostGetDevicePointer(&d, h, 0);
    if (e != cudaSuccess) { cudaFreeHost(h); dfine_set_error("flag_create: device alias: %s", cudaGetErrorString(e)); return -2; }
    *host_ptr = h;
    *dev_ptr = d;
    return 0;
}

DFINE_API int dfine_flag_destroy(void* host_ptr) {
    if (host_ptr) cudaFreeHost(host_ptr);
    return 0;
}

// Everything enqueued on `stream` after this call starts once the flag word is >= value (cyclic 32-bit comparison).
DFINE_API int dfine_stream_wait_flag(void* dev_ptr, int value, void* stream) {
    WaitValue32Fn fn = get_wait_value();
    DFINE_REQUIRE(fn != nullptr && dev_ptr != nullptr, "stream_wait_flag: cuStreamWaitValue32 unavailable");
    const CUresult r = fn((CUstream)stream, (CUdeviceptr)(uintptr_t)dev_ptr, (cuuint32_t)(unsigned int)value, CU_STREAM_WAIT_VALUE_GEQ);
    if (r != CUDA_SUCCESS) { dfine_set_error("stream_wait_flag: cuStreamWaitValue32 failed (%d)", (int)r); return -2; }
    return 0;
}
