// Shared helpers for libm2d_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdarg>
#include <cstdio>
#include <cstdint>
#include "../../include/m2d.h"

namespace m2d {

void set_error(const char* fmt, ...);

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return M2D_ERR_CUDA;
    }
    return M2D_OK;
}

#define M2D_REQUIRE(cond, ...)                      \
    do {                                            \
        if (!(cond)) {                              \
            m2d::set_error(__VA_ARGS__);            \
            return M2D_ERR_BAD_ARG;                 \
        }                                           \
    } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline long long cdiv(long long a, long long b) { return (a + b - 1) / b; }

constexpr int kNumSMs = 148;   // B200

// m2d_rowconv_args.w_tiled layout (include/m2d.h): rows per block and float index of the hi value of
// element (n, tap t, channel ci) of an [N][T x Cc] weight operand; the lo value lives R*32 floats further.
__host__ __device__ inline int tiled_rows(int N) { return N <= 64 ? 64 : 128; }
__host__ __device__ inline long long tiled_block_floats(int R) { return 2LL * R * 32; }
__host__ __device__ inline long long tiled_blocks(int N, int T, int Cc) {
    const int R = tiled_rows(N);
    return (long long)((N + R - 1) / R) * T * ((Cc + 31) / 32);
}
__host__ __device__ inline long long tiled_index(int n, int t, int ci, int T, int Cc, int R) {
    const int nt = n / R, nl = n - nt * R, c = ci >> 5, kk = ci & 31;
    const long long blk = ((long long)nt * T + t) * ((Cc + 31) / 32) + c;
    return blk * tiled_block_floats(R) + nl * 32 + ((((kk >> 2) ^ (nl & 7)) << 2) | (kk & 3));
}

__device__ __forceinline__ float rna_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
// hi / lo values of element `v` into the tiled block that holds float index `pidx` (the hi plane).  TF32X3: lo plane
// = rna_tf32(v - hi) at pidx + R*32.  TF32_BF16 (`mixed`): the lo plane holds, per 128-byte row, the weight side of
// the BF16 cross-term contraction [bf16(lo) x 32 | bf16(hi) x 32] in the same 16-byte-chunk swizzle.
__device__ __forceinline__ void store_tiled_split(float* dst_tiled, long long pidx, int R, float v, bool mixed) {
    const float h = rna_tf32(v);
    dst_tiled[pidx] = h;
    if (!mixed) {
        dst_tiled[pidx + R * 32] = rna_tf32(v - h);
        return;
    }
    const long long blk = pidx / (64LL * R);
    const int f = (int)(pidx - blk * 64LL * R);                 // float index inside the hi plane
    const int nl = f >> 5, j = ((f >> 2) & 7) ^ (nl & 7), kk = 4 * j + (f & 3);
    __nv_bfloat16* plane = reinterpret_cast<__nv_bfloat16*>(dst_tiled + blk * 64LL * R + R * 32);
    const int sw = nl & 7;
    plane[nl * 64 + ((((kk >> 3) ^ sw) << 3) | (kk & 7))] = __float2bfloat16_rn(v - h);
    plane[nl * 64 + (((((32 + kk) >> 3) ^ sw) << 3) | (kk & 7))] = __float2bfloat16_rn(h);
}

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == M2D_ACT_RELU) return v > 0.f ? v : 0.f;
    if (act == M2D_ACT_LEAKY) return v > 0.f ? v : 0.2f * v;
    if (act == M2D_ACT_TANH) return tanhf(v);
    return v;
}
// derivative of the activation evaluated from the stored post-activation value y
__device__ __forceinline__ float act_deriv(float y, int mode) {
    if (mode == M2D_MASK_RELU) return y > 0.f ? 1.f : 0.f;
    if (mode == M2D_MASK_LEAKY) return y > 0.f ? 1.f : 0.2f;
    if (mode == M2D_MASK_TANH) return 1.f - y * y;
    return 1.f;
}

struct AdamConst {
    float step_size, bc2_sqrt, b1, b2, eps, gscale;
};
// torch.optim.Adam element update.  Every rounding is pinned with intrinsics (no compiler-chosen FMA contraction), so
// the flat kernel (m2d_adam) and every path of the fused kernel (m2d_adam_pack) produce bit-identical p / m / v.
__device__ __forceinline__ float adam_update(float p, float g, float& m, float& v, const AdamConst& c) {
    const float gi = __fmul_rn(g, c.gscale);
    const float mi = __fmaf_rn(__fsub_rn(gi, m), 1.f - c.b1, m);
    const float vi = __fmaf_rn(__fmul_rn(1.f - c.b2, gi), gi, __fmul_rn(v, c.b2));
    m = mi;
    v = vi;
    const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(vi), c.bc2_sqrt), c.eps);
    return __fsub_rn(p, __fmul_rn(c.step_size, __fdiv_rn(mi, denom)));
}
__device__ __forceinline__ AdamConst adam_const(int t_step, float lr, float b1, float b2, float eps, float gscale) {
    AdamConst c;
    const double bc1 = 1.0 - pow((double)b1, (double)t_step);
    const double bc2 = 1.0 - pow((double)b2, (double)t_step);
    c.step_size = (float)((double)lr / bc1);
    c.bc2_sqrt = (float)sqrt(bc2);
    c.b1 = b1; c.b2 = b2; c.eps = eps; c.gscale = gscale;
    return c;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace m2d
