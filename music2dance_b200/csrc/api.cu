// Library-level entry points: error reporting, version, device check.
#include "common.cuh"
#include <cstdlib>
#include <cstring>

namespace m2d {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
// default: 3xTF32 on the tensor cores; the environment variable M2D_GEMM (fp32 | tf32 | tf32bf16 | tf32x3),
// read once at load, lets a whole test / bench run be repeated in another arithmetic
static int initial_gemm_mode() {
    const char* e = getenv("M2D_GEMM");
    if (e && !strcmp(e, "fp32")) return M2D_GEMM_FP32;
    if (e && !strcmp(e, "tf32")) return M2D_GEMM_TF32;
    if (e && !strcmp(e, "tf32bf16")) return M2D_GEMM_TF32_BF16;
    return M2D_GEMM_TF32X3;
}
static int g_gemm_mode = initial_gemm_mode();
int gemm_mode() { return g_gemm_mode; }
}  // namespace m2d

extern "C" int m2d_set_gemm_mode(int mode) {
    if (mode != M2D_GEMM_FP32 && mode != M2D_GEMM_TF32 && mode != M2D_GEMM_TF32_BF16 && mode != M2D_GEMM_TF32X3) {
        m2d::set_error("set_gemm_mode: unknown mode %d", mode);
        return M2D_ERR_BAD_ARG;
    }
    m2d::g_gemm_mode = mode;
    return M2D_OK;
}
extern "C" int m2d_get_gemm_mode(void) { return m2d::g_gemm_mode; }

extern "C" const char* m2d_last_error(void) { return m2d::g_err; }
extern "C" int m2d_version(void) { return 100; }
namespace m2d { long long halo_launch_count(); long long halo_persist_launch_count(); }
extern "C" long long m2d_halo_launch_count(void) { return m2d::halo_launch_count(); }
extern "C" long long m2d_halo_persist_launch_count(void) { return m2d::halo_persist_launch_count(); }

extern "C" int m2d_check_device(int dev) {
    cudaDeviceProp p;
    cudaError_t e = cudaGetDeviceProperties(&p, dev);
    if (e != cudaSuccess) {
        m2d::set_error("check_device: %s", cudaGetErrorString(e));
        return M2D_ERR_CUDA;
    }
    if (p.major != 10) {
        m2d::set_error("libm2d_b200 is built for sm_100a only; device %d is sm_%d%d", dev, p.major, p.minor);
        return M2D_ERR_ARCH;
    }
    return M2D_OK;
}
