// rz_cabi.cu -- ABI version, error string, device query.
#include <string.h>
#include "rz_common.cuh"

namespace rz {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace rz

extern "C" {
int rz_abi_version(void) { return RZ_ABI_VERSION; }
const char* rz_last_error(void) { return rz::g_err; }
int rz_device_count(int* count) {
    RZ_REQUIRE(count, "rz_device_count: null pointer");
    *count = 0;
    RZ_CUDA_TRY(cudaGetDeviceCount(count));
    return RZ_OK;
}
}
