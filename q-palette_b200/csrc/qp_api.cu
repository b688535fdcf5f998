// qp_api.cu -- library-wide state of libqpalette.so: error string, launch counter, device attributes.
#include "qp_common.cuh"

namespace qp {

std::atomic<uint64_t> g_launches{0};


char *last_error_buf() {
    static thread_local char buf[512] = {0};
    return buf;
}

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(last_error_buf(), 512, fmt, ap);
    va_end(ap);
    return code;
}

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

}  // namespace qp

extern "C" int qp_version(void) { return 100; }
extern "C" const char *qp_last_error(void) { return qp::last_error_buf(); }
extern "C" uint64_t qp_launch_count(void) { return qp::g_launches.load(); }
extern "C" int qp_device_sm_count(void) { return qp::sm_count(); }
