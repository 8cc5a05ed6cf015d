// cabi.cu -- library-level pieces of the C ABI (include/b200dsp.h): version, thread-local
// error string, launch counter, device info.
#include "common.cuh"
#include <stdarg.h>

namespace b200dsp {

static thread_local char g_err[512] = "";
static thread_local int64_t g_launches = 0;

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch() { ++g_launches; }

}  // namespace b200dsp

extern "C" {

int b200dsp_version(void) { return B200DSP_VERSION; }

const char *b200dsp_last_error(void) { return b200dsp::g_err; }

int b200dsp_device_info(int *sm_count, int *cc_major, int *cc_minor)
{
    int dev = 0;
    B200_CHECK_CUDA(cudaGetDevice(&dev));
    int v = 0;
    if (sm_count) {
        B200_CHECK_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
        *sm_count = v;
    }
    if (cc_major) {
        B200_CHECK_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev));
        *cc_major = v;
    }
    if (cc_minor) {
        B200_CHECK_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev));
        *cc_minor = v;
    }
    return B200DSP_OK;
}

int64_t b200dsp_launch_count(void) { return b200dsp::g_launches; }
void b200dsp_launch_count_reset(void) { b200dsp::g_launches = 0; }

}  // extern "C"
