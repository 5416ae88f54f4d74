// Single-input-channel strided convolutions with 32 output channels (AudioDiscriminator.l1,
// default.py:298; WaveGAN-style / U-Net encoders' first layers have the same form): forward,
// WGAN-GP tangent pass (forward through the stored ReLU mask) and weight gradient.
//
// These layers carry 0.1 % of the step's FLOPs but touch the largest activation of the critic
// (B x 19200 x 32 floats = 17 MB at B = 7), i.e. they are HBM-bound: per output row 25 FMAs per channel
// against 128 bytes written (forward) or read (weight gradient).  On the tensor-core path they cost
// 42-74 us per launch (a K = 25 contraction gathered one scalar at a time); here lane = output channel,
// the taps of the lane's filter live in registers, the audio segment of the CTA's rows is staged once in
// shared memory (zero padding resolved while staging) and read back as warp-wide broadcasts, and every
// row is one coalesced 128-byte store / load.
#include "common.cuh"

namespace m2d {

constexpr int C1_ROWS = 256;        // output rows per CTA
constexpr int C1_U = 4;             // rows in flight per warp (independent global loads / stores)
constexpr int C1_WARPS = 8;
constexpr int C1_MAXT = 32;

// stage x[b, r0 .. r0 + n) (zero outside [0, x_rows)) into shared memory
__device__ __forceinline__ void c1_stage(float* xs, const float* __restrict__ xb, int r0, int n, int x_rows) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int r = r0 + i;
        xs[i] = (r >= 0 && r < x_rows) ? __ldg(xb + r) : 0.f;
    }
}

template <int T>
__global__ void __launch_bounds__(C1_WARPS * 32)
conv_c1_fwd_kernel(const m2d_rowconv_args a) {
    extern __shared__ float xs[];
    const int tiles = (a.y_rows + C1_ROWS - 1) / C1_ROWS;
    const int b = blockIdx.x / tiles;
    const int i0 = (blockIdx.x - b * tiles) * C1_ROWS;
    const int nrows = min(C1_ROWS, a.y_rows - i0);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int seg = (nrows - 1) * a.sr + T;
    c1_stage(xs, a.x + (long long)b * a.x_bs, i0 * a.sr + a.roff0, seg, a.x_rows);
    float w[T];
#pragma unroll
    for (int t = 0; t < T; ++t) w[t] = __ldg(a.w + (long long)lane * a.w_ld + t);
    const float bz = a.bias ? __ldg(a.bias + lane) : 0.f;
    __syncthreads();
    for (int r0 = warp; r0 < nrows; r0 += C1_WARPS * C1_U) {
        float mk[C1_U], acc[C1_U];
#pragma unroll
        for (int u = 0; u < C1_U; ++u) {                  // mask loads of the group first: C1_U loads in flight
            const int r = r0 + u * C1_WARPS;
            mk[u] = (a.mask_mode && r < nrows)
                        ? __ldg(a.mask + b * a.m_bs + (long long)(i0 + r) * a.m_ld + lane) : 1.f;
            acc[u] = bz;
        }
#pragma unroll
        for (int t = 0; t < T; ++t) {
#pragma unroll
            for (int u = 0; u < C1_U; ++u) acc[u] = fmaf(xs[min(r0 + u * C1_WARPS, nrows - 1) * a.sr + t], w[t], acc[u]);
        }
#pragma unroll
        for (int u = 0; u < C1_U; ++u) {
            const int r = r0 + u * C1_WARPS;
            if (r >= nrows) continue;
            float v = apply_act(acc[u], a.act);
            const long long yo = b * a.y_bs + (long long)(i0 + r) * a.y_ld + lane;
            if (a.y2) a.y2[yo] = v;
            if (a.mask_mode) v *= act_deriv(mk[u], a.mask_mode);
            a.y[yo] = v;
        }
    }
}

// partial[cta][co][t] = sum over the CTA's rows of dy[b,i,co] * x[b, i*sr + roff0 + t]
template <int T>
__global__ void __launch_bounds__(C1_WARPS * 32)
conv_c1_wgrad_kernel(const m2d_wgrad_args a) {
    extern __shared__ float xs[];
    const int tiles = (a.dy_rows + C1_ROWS - 1) / C1_ROWS;
    const int b = blockIdx.x / tiles;
    const int i0 = (blockIdx.x - b * tiles) * C1_ROWS;
    const int nrows = min(C1_ROWS, a.dy_rows - i0);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int seg = (nrows - 1) * a.sr + T;
    c1_stage(xs, a.x + (long long)b * a.x_bs, i0 * a.sr + a.roff0, seg, a.x_rows);
    float* red = xs + ((seg + 3) & ~3);               // [C1_WARPS][T][32]
    float acc[T];
#pragma unroll
    for (int t = 0; t < T; ++t) acc[t] = 0.f;
    __syncthreads();
    const float* dyb = a.dy + (long long)b * a.dy_bs + lane;
    for (int r0 = warp; r0 < nrows; r0 += C1_WARPS * C1_U) {
        float d[C1_U];
#pragma unroll
        for (int u = 0; u < C1_U; ++u) {                  // C1_U independent row loads in flight
            const int r = r0 + u * C1_WARPS;
            d[u] = r < nrows ? __ldg(dyb + (long long)(i0 + r) * a.dy_ld) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < C1_U; ++u) {
            const float* p = xs + min(r0 + u * C1_WARPS, nrows - 1) * a.sr;
#pragma unroll
            for (int t = 0; t < T; ++t) acc[t] = fmaf(d[u], p[t], acc[t]);
        }
    }
#pragma unroll
    for (int t = 0; t < T; ++t) red[(warp * T + t) * 32 + lane] = acc[t];
    __syncthreads();
    float* part = a.ws + (long long)blockIdx.x * 32 * T;
    for (int idx = threadIdx.x; idx < 32 * T; idx += blockDim.x) {
        const int t = idx >> 5, co = idx & 31;
        float s = 0.f;
#pragma unroll
        for (int wv = 0; wv < C1_WARPS; ++wv) s += red[(wv * T + t) * 32 + co];      // fixed order: deterministic
        part[co * T + t] = s;
    }
}

// dw[co, 0, t] = beta*dw + scale * sum over CTAs of partial[cta][co][t].  Block = 32 outputs x 32 part groups: lane =
// output (coalesced 128-byte reads of a partial's row), warp w sums parts w, w+32, ... (about 30 loads in flight per
// thread instead of one 1000-long dependent chain: 80 us -> a few us), then a fixed-order reduction over the warps
// in shared memory (deterministic).
__global__ void __launch_bounds__(1024)
conv_c1_wgrad_finalize(const float* __restrict__ part, int nparts, int n, float* dw, float scale, float beta) {
    __shared__ float red[32][33];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int idx = blockIdx.x * 32 + lane;
    float s = 0.f;
    if (idx < n)
        for (int p = w; p < nparts; p += 32) s += part[(long long)p * n + idx];
    red[w][lane] = s;
    __syncthreads();
    if (w == 0 && idx < n) {
        float v = 0.f;
#pragma unroll
        for (int k = 0; k < 32; ++k) v += red[k][lane];
        v *= scale;
        if (beta != 0.f) v += beta * dw[idx];
        dw[idx] = v;
    }
}

// Backward-data to the single input channel (the WGAN-GP audio gradient dD/d audio, losses.py:40-44 through
// audio_d.l1): dx[b, j] = sum_{co, t : 4 i - PAD + t = j} dy[b, i, co] * w[co, t].  One warp per group of 4
// consecutive samples j = 4m + p: lane = output channel, the lane's 25 taps in registers; the 7 rows i = m - d
// (d = -3..3) that touch the group are read once (coalesced 128-byte rows, L1 reuse between neighbouring groups)
// and every tap is used exactly once: t = 4d + p + PAD.  The four sums are reduced over the 32 channels with a
// folded butterfly (6 shuffles for 4 values) and written as one 16-byte segment.
template <int T, int PAD>
__global__ void __launch_bounds__(256)
conv_c1_dgrad4_kernel(const float* __restrict__ dy, const float* __restrict__ w, int w_ld, float* __restrict__ dx,
                      int Lout, int Lin, long long groups) {
    const int lane = threadIdx.x & 31;
    float wr[T];
#pragma unroll
    for (int t = 0; t < T; ++t) wr[t] = __ldg(w + (long long)lane * w_ld + t);
    const int gpb = Lin / 4;
    constexpr int DLO = -((3 + PAD) / 4), DHI = (T - 1 - PAD) >= 0 ? (T - 1 - PAD) / 4 : -1;
    const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long g = warp0; g < groups; g += nwarps) {
        const int b = (int)(g / gpb), m = (int)(g - (long long)b * gpb);
        const float* dyb = dy + ((long long)b * Lout) * 32 + lane;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        float v[DHI - DLO + 1];
#pragma unroll
        for (int d = DLO; d <= DHI; ++d) {                  // all row loads first: independent, in flight together
            const int i = m - d;
            v[d - DLO] = (i >= 0 && i < Lout) ? __ldg(dyb + (long long)i * 32) : 0.f;
        }
#pragma unroll
        for (int d = DLO; d <= DHI; ++d)
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                constexpr int dummy = 0;
                const int t = 4 * d + p + PAD;              // compile-time after unrolling
                if (t >= 0 && t < T) acc[p] = fmaf(v[d - DLO], wr[t], acc[p]);
                (void)dummy;
            }
        // folded reduction over the 32 channels: halves hold outputs {0,1} / {2,3}, then quarters one output each
        const bool up = lane & 16;
        float s0 = __shfl_xor_sync(0xffffffffu, up ? acc[0] : acc[2], 16);
        float s1 = __shfl_xor_sync(0xffffffffu, up ? acc[1] : acc[3], 16);
        float a = (up ? acc[2] : acc[0]) + s0, c = (up ? acc[3] : acc[1]) + s1;
        const bool q = lane & 8;
        float s2 = __shfl_xor_sync(0xffffffffu, q ? a : c, 8);
        float r = (q ? c : a) + s2;
        r += __shfl_xor_sync(0xffffffffu, r, 4);
        r += __shfl_xor_sync(0xffffffffu, r, 2);
        r += __shfl_xor_sync(0xffffffffu, r, 1);
        if ((lane & 7) == 0) dx[(long long)b * Lin + 4 * m + 2 * (lane >> 4) + ((lane >> 3) & 1)] = r;
    }
}

// Returns 1 if the layer is not of this form.
int conv_c1_dgrad_dispatch(const float* dy, int nb, int Lout, int Cout, const float* w, int k, int stride, int pad,
                           float* dx, int Lin, cudaStream_t st) {
    if (Cout != 32 || k != 25 || stride != 4 || Lin % 4 || (pad != 11 && pad != 0)) return 1;
    const long long groups = (long long)nb * (Lin / 4);
    const int blocks = (int)(cdiv(groups, 8) < 16 * kNumSMs ? cdiv(groups, 8) : 16 * kNumSMs);
    if (pad == 11) conv_c1_dgrad4_kernel<25, 11><<<blocks, 256, 0, st>>>(dy, w, k, dx, Lout, Lin, groups);
    else conv_c1_dgrad4_kernel<25, 0><<<blocks, 256, 0, st>>>(dy, w, k, dx, Lout, Lin, groups);
    return check_launch("conv_c1_dgrad4");
}

template <int T>
static void launch_fwd(const m2d_rowconv_args& a, int grid, int smem, cudaStream_t st) {
    conv_c1_fwd_kernel<T><<<grid, C1_WARPS * 32, smem, st>>>(a);
}
template <int T>
static void launch_wgrad(const m2d_wgrad_args& a, int grid, int smem, cudaStream_t st) {
    conv_c1_wgrad_kernel<T><<<grid, C1_WARPS * 32, smem, st>>>(a);
}

// Returns 1 if the call is not of this form (caller continues with the general kernels).
int conv_c1_fwd_dispatch(const m2d_rowconv_args& a, cudaStream_t st) {
    if (a.Cc != 1 || a.N != 32 || a.win_T > 0 || a.droff != 1 || a.sr < 1 || a.add || a.x_ld != 1) return 1;
    if (a.T != 25) return 1;                          // instantiated filter length (audio_d.l1, WaveGAN-style l1)
    const int tiles = (a.y_rows + C1_ROWS - 1) / C1_ROWS;
    const int smem = ((C1_ROWS - 1) * a.sr + a.T) * 4;
    if (smem > 48 * 1024) return 1;
    const int grid = a.nb * tiles;
    launch_fwd<25>(a, grid, smem, st);
    return check_launch("conv_c1_fwd");
}

int conv_c1_wgrad_dispatch(const m2d_wgrad_args& a, cudaStream_t st) {
    if (a.Cc != 1 || a.Cout != 32 || a.win_T > 0 || a.droff != 1 || a.sr < 1 || a.x_ld != 1 || !a.ws) return 1;
    if (a.T != 25) return 1;
    const int tiles = (a.dy_rows + C1_ROWS - 1) / C1_ROWS;
    const int grid = a.nb * tiles;
    const int seg = (C1_ROWS - 1) * a.sr + a.T;
    const int smem = (((seg + 3) & ~3) + C1_WARPS * a.T * 32) * 4;
    if (smem > 48 * 1024 || a.ws_floats < (long long)grid * 32 * a.T) return 1;
    launch_wgrad<25>(a, grid, smem, st);
    int rc = check_launch("conv_c1_wgrad");
    if (rc) return rc;
    const int n = 32 * a.T;
    conv_c1_wgrad_finalize<<<(n + 31) / 32, 1024, 0, st>>>(a.ws, grid, n, a.dw, a.scale, a.beta);
    return check_launch("conv_c1_wgrad_finalize");
}

}  // namespace m2d
