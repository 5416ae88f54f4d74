// Row-convolution GEMM family (fp32 SIMT path): forward / backward-data / tangent
// passes of Conv1d and Linear (m2d_rowconv), weight gradients (m2d_wgrad), weight
// re-layouts and the single-input-channel backward-data kernel.
//
// Data layout: channels-last row matrices (see include/m2d.h).  Tiling: 256 threads
// per CTA, BM x 64 output tile, K-steps of 16, register micro-tile (BM/16) x 4,
// double-buffered shared memory with register prefetch.
#include <cstdlib>
#include "common.cuh"

namespace m2d {

constexpr int BN = 64;
constexpr int BK = 16;
constexpr int NT = 256;

__device__ __forceinline__ void epi_store(const m2d_rowconv_args& a, int m, int n, float v) {
    int b = m / a.y_rows;
    int i = m - b * a.y_rows;
    if (a.bias) v += __ldg(a.bias + n);
    v = apply_act(v, a.act);
    if (a.add && a.add_before_mask) v += a.add[b * a.a_bs + (long long)i * a.a_ld + n];
    if (a.y2) a.y2[b * a.y_bs + (long long)i * a.y_ld + n] = v;
    if (a.mask_mode) v *= act_deriv(a.mask[b * a.m_bs + (long long)i * a.m_ld + n], a.mask_mode);
    if (a.add && !a.add_before_mask) v += a.add[b * a.a_bs + (long long)i * a.a_ld + n];
    a.y[b * a.y_bs + (long long)i * a.y_ld + n] = v;
}

template <int BM, bool VEC, bool C1>
__global__ void __launch_bounds__(NT)
rowconv_kernel(const m2d_rowconv_args a, const int M, const int nsteps, const int cchunks) {
    constexpr int TM = BM / 16;
    constexpr int A_SLOTS = BM / 64;
    __shared__ __align__(16) float As[2][BK][BM + 4];
    __shared__ __align__(16) float Bs[2][BK][BN + 4];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int per = (nsteps + gridDim.z - 1) / gridDim.z;
    const int s_begin = blockIdx.z * per;
    const int s_end = min(nsteps, s_begin + per);
    const bool win = a.win_T > 0;

    // per-slot row bookkeeping for the A (activation) operand
    long long a_base[A_SLOTS];
    int a_r0[A_SLOTS], a_ab[A_SLOTS];
    bool a_ok[A_SLOTS];
#pragma unroll
    for (int s = 0; s < A_SLOTS; ++s) {
        int idx = tid + s * NT;
        int m = m0 + (idx >> 2);
        a_ok[s] = m < M;
        int mm = a_ok[s] ? m : 0;
        int b = mm / a.y_rows;
        int i = mm - b * a.y_rows;
        a_r0[s] = i * a.sr + a.roff0;
        if (win) {
            int seq = b / a.win_T;
            int f = b - seq * a.win_T;
            a_ab[s] = f * a.win_stride - a.win_pad;
            a_base[s] = (long long)seq * a.win_seq_len;
        } else {
            a_ab[s] = 0;
            a_base[s] = (long long)b * a.x_bs;
        }
    }
    const int b_n = n0 + (tid >> 2);
    const bool b_ok = b_n < a.N;
    const float* b_row = a.w + (long long)(b_ok ? b_n : 0) * a.w_ld;
    const int kq = (tid & 3) * 4;

    float4 ra[A_SLOTS], rb;

    auto load = [&](int s) {
        if (C1) {
            const int k0 = s * BK + kq;
#pragma unroll
            for (int sl = 0; sl < A_SLOTS; ++sl) {
                float v[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    int k = k0 + e;
                    int r = a_r0[sl] + k * a.droff;
                    bool ok = a_ok[sl] && k < a.T && r >= 0 && r < a.x_rows;
                    int pos = a_ab[sl] + r;
                    if (win) ok = ok && pos >= 0 && pos < a.win_seq_len;
                    v[e] = ok ? __ldg(a.x + a_base[sl] + (long long)pos * a.x_ld) : 0.f;
                }
                ra[sl] = make_float4(v[0], v[1], v[2], v[3]);
            }
            float v[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                int k = k0 + e;
                v[e] = (b_ok && k < a.T) ? __ldg(b_row + k) : 0.f;
            }
            rb = make_float4(v[0], v[1], v[2], v[3]);
        } else {
            const int t = s / cchunks;
            const int c = (s - t * cchunks) * BK + kq;
#pragma unroll
            for (int sl = 0; sl < A_SLOTS; ++sl) {
                int r = a_r0[sl] + t * a.droff;
                bool ok = a_ok[sl] && r >= 0 && r < a.x_rows;
                const float* p = a.x + a_base[sl] + (long long)r * a.x_ld + c;
                if (VEC) {
                    ra[sl] = (ok && c < a.Cc) ? __ldg(reinterpret_cast<const float4*>(p))
                                               : make_float4(0.f, 0.f, 0.f, 0.f);
                } else {
                    float v[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) v[e] = (ok && c + e < a.Cc) ? __ldg(p + e) : 0.f;
                    ra[sl] = make_float4(v[0], v[1], v[2], v[3]);
                }
            }
            const float* p = b_row + (long long)t * a.Cc + c;
            if (VEC) {
                rb = (b_ok && c < a.Cc) ? __ldg(reinterpret_cast<const float4*>(p))
                                         : make_float4(0.f, 0.f, 0.f, 0.f);
            } else {
                float v[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) v[e] = (b_ok && c + e < a.Cc) ? __ldg(p + e) : 0.f;
                rb = make_float4(v[0], v[1], v[2], v[3]);
            }
        }
    };
    auto stage = [&](int buf) {
#pragma unroll
        for (int sl = 0; sl < A_SLOTS; ++sl) {
            int row = (tid + sl * NT) >> 2;
            As[buf][kq + 0][row] = ra[sl].x;
            As[buf][kq + 1][row] = ra[sl].y;
            As[buf][kq + 2][row] = ra[sl].z;
            As[buf][kq + 3][row] = ra[sl].w;
        }
        int n = tid >> 2;
        Bs[buf][kq + 0][n] = rb.x;
        Bs[buf][kq + 1][n] = rb.y;
        Bs[buf][kq + 2][n] = rb.z;
        Bs[buf][kq + 3][n] = rb.w;
    };

    float acc[TM][4];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    if (s_begin < s_end) {
        load(s_begin);
        stage(0);
    }
    __syncthreads();
    int cur = 0;
    for (int s = s_begin; s < s_end; ++s) {
        const bool more = s + 1 < s_end;
        if (more) load(s + 1);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float av[TM];
#pragma unroll
            for (int i4 = 0; i4 < TM / 4; ++i4) {
                float4 t4 = *reinterpret_cast<const float4*>(&As[cur][kk][ty * TM + i4 * 4]);
                av[i4 * 4 + 0] = t4.x; av[i4 * 4 + 1] = t4.y; av[i4 * 4 + 2] = t4.z; av[i4 * 4 + 3] = t4.w;
            }
            float4 b4 = *reinterpret_cast<const float4*>(&Bs[cur][kk][tx * 4]);
#pragma unroll
            for (int i = 0; i < TM; ++i) {
                acc[i][0] = fmaf(av[i], b4.x, acc[i][0]);
                acc[i][1] = fmaf(av[i], b4.y, acc[i][1]);
                acc[i][2] = fmaf(av[i], b4.z, acc[i][2]);
                acc[i][3] = fmaf(av[i], b4.w, acc[i][3]);
            }
        }
        if (more) stage(cur ^ 1);
        __syncthreads();
        cur ^= 1;
    }

    if (gridDim.z == 1) {
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            int m = m0 + ty * TM + i;
            if (m >= M) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int n = n0 + tx * 4 + j;
                if (n < a.N) epi_store(a, m, n, acc[i][j]);
            }
        }
    } else {
        float* ws = a.ws + (long long)blockIdx.z * M * a.N;
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            int m = m0 + ty * TM + i;
            if (m >= M) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int n = n0 + tx * 4 + j;
                if (n < a.N) ws[(long long)m * a.N + n] = acc[i][j];
            }
        }
    }
}

__global__ void rowconv_splitk_epilogue(const m2d_rowconv_args a, const int M, const int splits) {
    long long total = (long long)M * a.N;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        // eight independent partial sums: the loads of a thread overlap instead of forming one dependent chain
        // (148 splits x ~0.6 us latency was 20 us for the 7 x 100 outputs of audio_d.l6); fixed order: deterministic
        float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        int z = 0;
        for (; z + 8 <= splits; z += 8) {
#pragma unroll
            for (int u = 0; u < 8; ++u) s[u] += a.ws[(long long)(z + u) * total + idx];
        }
        for (; z < splits; ++z) s[0] += a.ws[(long long)z * total + idx];
        float v = ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
        int m = (int)(idx / a.N);
        int n = (int)(idx - (long long)m * a.N);
        epi_store(a, m, n, v);
    }
}

// split-K epilogue for few outputs and many splits (weight-streaming layers at small batch: 7 x 100 outputs from
// 148 partials): one WARP per output, lanes stride over the splits, fixed-order shuffle tree (deterministic)
__global__ void rowconv_splitk_epilogue_warp(const m2d_rowconv_args a, const int M, const int splits) {
    const long long total = (long long)M * a.N;
    const int lane = threadIdx.x & 31;
    const long long idx = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    if (idx >= total) return;
    float s = 0.f;
    for (int z = lane; z < splits; z += 32) s += a.ws[(long long)z * total + idx];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
        const int m = (int)(idx / a.N);
        epi_store(a, m, (int)(idx - (long long)m * a.N), s);
    }
}

// Skinny GEMM: a handful of rows (M <= MR) against a long contraction whose operands are both K-contiguous
// (audio_d.l6 / stick_d.fconv as Linear over (tap, channel), their tangent passes, at the reference batch):
// pure weight streaming.  grid.x = K splits over the whole chip; a CTA stages its K chunk of the M activation rows
// in shared memory, each warp takes output columns n = warp, warp + 8, ..., its lanes stride over the chunk with
// 16-byte loads of the weight row (coalesced) and keep M partial dot products in registers; warp-shuffle reduce,
// partials [split][m][n] to the workspace, then the warp-per-output epilogue.
template <int MR>
__global__ void __launch_bounds__(NT)
skinny_gemm_kernel(const m2d_rowconv_args a, const int M, const long long K, const int KC) {
    extern __shared__ float As[];                       // [MR][KC]
    const long long k0 = (long long)blockIdx.x * KC;
    const int kc = (int)min((long long)KC, K - k0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int idx = tid; idx < MR * (KC / 4); idx += NT) {
        const int m = idx / (KC / 4), k4 = (idx - m * (KC / 4)) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m < M && k4 < kc) {
            const int b = m / a.y_rows, i = m - b * a.y_rows;
            v = __ldg(reinterpret_cast<const float4*>(a.x + (long long)b * a.x_bs +
                                                      (long long)(i * a.sr + a.roff0) * a.x_ld + k0 + k4));
        }
        *reinterpret_cast<float4*>(As + m * KC + k4) = v;
    }
    __syncthreads();
    float* part = a.ws + (long long)blockIdx.x * M * a.N;
    constexpr int NU = MR <= 8 ? 4 : 2;                 // weight rows in flight per warp (independent 16-byte loads)
    for (int nb = warp * NU; nb < a.N; nb += (NT / 32) * NU) {
        float acc[NU][MR];
#pragma unroll
        for (int u = 0; u < NU; ++u)
#pragma unroll
            for (int m = 0; m < MR; ++m) acc[u][m] = 0.f;
        for (int k4 = lane * 4; k4 < kc; k4 += 128) {
            float4 w4[NU];
#pragma unroll
            for (int u = 0; u < NU; ++u)
                w4[u] = nb + u < a.N ? __ldg(reinterpret_cast<const float4*>(a.w + (long long)(nb + u) * a.w_ld + k0 + k4))
                                     : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int m = 0; m < MR; ++m) {
                const float4 x4 = *reinterpret_cast<const float4*>(As + m * KC + k4);
#pragma unroll
                for (int u = 0; u < NU; ++u)
                    acc[u][m] = fmaf(x4.x, w4[u].x, fmaf(x4.y, w4[u].y, fmaf(x4.z, w4[u].z, fmaf(x4.w, w4[u].w, acc[u][m]))));
            }
        }
#pragma unroll
        for (int u = 0; u < NU; ++u)
#pragma unroll
            for (int m = 0; m < MR; ++m) {
                float s = acc[u][m];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                if (lane == 0 && m < M && nb + u < a.N) part[(long long)m * a.N + nb + u] = s;
            }
    }
}

// Returns 1 if the shape is not a skinny K-contiguous GEMM (caller continues with the tiled kernels).
static int skinny_dispatch(const m2d_rowconv_args& a, int M, cudaStream_t st) {
    static const bool enabled = !(getenv("M2D_SKINNY") && getenv("M2D_SKINNY")[0] == '0');
    if (!enabled) return 1;
    const long long K = (long long)a.T * a.Cc;
    if (M > 16 || K < 2048 || a.win_T > 0 || a.Cc == 1 || !a.ws || a.droff != 1) return 1;
    // both operands contiguous along K: single tap, or a full-length convolution over dense rows
    const bool kcontig = a.T == 1 || (a.x_ld == a.Cc && a.y_rows == 1 && a.roff0 == 0 && a.x_rows >= a.T);
    if (!kcontig || a.T * a.Cc > a.w_ld || K % 4 || a.x_ld % 4 || a.x_bs % 4 || a.w_ld % 4 || !aligned16(a.x) ||
        !aligned16(a.w))
        return 1;
    if (a.T == 1 && (a.roff0 < 0 || (a.y_rows - 1) * a.sr + a.roff0 >= a.x_rows)) return 1;
    int splits = kNumSMs;
    long long KC = (cdiv(K, splits) + 3) / 4 * 4;
    if (KC < 128) KC = 128;
    splits = (int)cdiv(K, KC);
    const int MR = M <= 8 ? 8 : 16;
    if (a.ws_floats < (long long)splits * M * a.N || MR * KC * 4 > 48 * 1024) return 1;
    const int smem = (int)(MR * KC * 4);
    if (MR == 8) skinny_gemm_kernel<8><<<splits, NT, smem, st>>>(a, M, K, (int)KC);
    else skinny_gemm_kernel<16><<<splits, NT, smem, st>>>(a, M, K, (int)KC);
    int rc = check_launch("skinny_gemm");
    if (rc) return rc;
    const long long threads = (long long)M * a.N * 32;
    rowconv_splitk_epilogue_warp<<<(unsigned)cdiv(threads, 256), 256, 0, st>>>(a, M, splits);
    return check_launch("rowconv_splitk_epilogue_warp");
}

template <int BM, bool VEC, bool C1>
static void launch_rowconv(const m2d_rowconv_args& a, int M, int nsteps, int cchunks, int splits,
                           cudaStream_t st) {
    dim3 grid((unsigned)cdiv(M, BM), (unsigned)cdiv(a.N, BN), (unsigned)splits);
    rowconv_kernel<BM, VEC, C1><<<grid, NT, 0, st>>>(a, M, nsteps, cchunks);
}

// ------------------------------------------------------------------ wgrad
// dW[co,(t,c)] = sum_k dy[k,co] * x[row(k,t), c];  M = Cout, N = T*Cc, K = nb*dy_rows.
template <bool VEC>
__global__ void __launch_bounds__(NT)
wgrad_kernel(const m2d_wgrad_args a, const int Ktot, const int Ncols) {
    constexpr int BM = 64, TM = 4;
    __shared__ __align__(16) float As[2][BK][BM + 4];
    __shared__ __align__(16) float Bs[2][BK][BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int nsteps = (Ktot + BK - 1) / BK;
    const int per = (nsteps + gridDim.z - 1) / gridDim.z;
    const int s_begin = blockIdx.z * per;
    const int s_end = min(nsteps, s_begin + per);
    const bool win = a.win_T > 0;

    const int kk = tid >> 4;            // row within the K tile handled by this thread
    const int q4 = (tid & 15) * 4;      // first of 4 columns
    // B operand columns (t, c) are fixed over the K loop
    int bt[4], bc[4];
    bool bok[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        int n = n0 + q4 + e;
        bok[e] = n < Ncols;
        int nn = bok[e] ? n : 0;
        bt[e] = nn / a.Cc;
        bc[e] = nn - bt[e] * a.Cc;
    }
    float4 ra, rb;
    auto load = [&](int s) {
        int k = s * BK + kk;
        bool kok = k < Ktot;
        int kc = kok ? k : 0;
        int b = kc / a.dy_rows;
        int l = kc - b * a.dy_rows;
        // A: dy[k, m0+q4 .. +3]
        {
            const float* p = a.dy + (long long)b * a.dy_bs + (long long)l * a.dy_ld + m0 + q4;
            if (VEC) {
                ra = (kok && m0 + q4 < a.Cout) ? __ldg(reinterpret_cast<const float4*>(p))
                                                : make_float4(0.f, 0.f, 0.f, 0.f);
            } else {
                float v[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) v[e] = (kok && m0 + q4 + e < a.Cout) ? __ldg(p + e) : 0.f;
                ra = make_float4(v[0], v[1], v[2], v[3]);
            }
        }
        long long base;
        int ab = 0;
        if (win) {
            int seq = b / a.win_T;
            int f = b - seq * a.win_T;
            ab = f * a.win_stride - a.win_pad;
            base = (long long)seq * a.win_seq_len;
        } else {
            base = (long long)b * a.x_bs;
        }
        const int rbase = l * a.sr + a.roff0;
        if (VEC) {
            int r = rbase + bt[0] * a.droff;
            bool ok = kok && bok[0] && r >= 0 && r < a.x_rows;
            rb = ok ? __ldg(reinterpret_cast<const float4*>(a.x + base + (long long)r * a.x_ld + bc[0]))
                    : make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
            float v[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                int r = rbase + bt[e] * a.droff;
                bool ok = kok && bok[e] && r >= 0 && r < a.x_rows;
                int pos = ab + r;
                if (win) ok = ok && pos >= 0 && pos < a.win_seq_len;
                v[e] = ok ? __ldg(a.x + base + (long long)pos * a.x_ld + bc[e]) : 0.f;
            }
            rb = make_float4(v[0], v[1], v[2], v[3]);
        }
    };
    auto stage = [&](int buf) {
        *reinterpret_cast<float4*>(&As[buf][kk][q4]) = ra;
        *reinterpret_cast<float4*>(&Bs[buf][kk][q4]) = rb;
    };
    float acc[TM][4];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    if (s_begin < s_end) {
        load(s_begin);
        stage(0);
    }
    __syncthreads();
    int cur = 0;
    for (int s = s_begin; s < s_end; ++s) {
        const bool more = s + 1 < s_end;
        if (more) load(s + 1);
#pragma unroll
        for (int k2 = 0; k2 < BK; ++k2) {
            float4 a4 = *reinterpret_cast<const float4*>(&As[cur][k2][ty * 4]);
            float4 b4 = *reinterpret_cast<const float4*>(&Bs[cur][k2][tx * 4]);
            float av[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
            for (int i = 0; i < TM; ++i) {
                acc[i][0] = fmaf(av[i], b4.x, acc[i][0]);
                acc[i][1] = fmaf(av[i], b4.y, acc[i][1]);
                acc[i][2] = fmaf(av[i], b4.z, acc[i][2]);
                acc[i][3] = fmaf(av[i], b4.w, acc[i][3]);
            }
        }
        if (more) stage(cur ^ 1);
        __syncthreads();
        cur ^= 1;
    }
    float* ws = a.ws + (long long)blockIdx.z * a.Cout * Ncols;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int m = m0 + ty * 4 + i;
        if (m >= a.Cout) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int n = n0 + tx * 4 + j;
            if (n < Ncols) ws[(long long)m * Ncols + n] = acc[i][j];
        }
    }
}

// sum split-K partials and scatter into the PyTorch (Cout, Cc, T) layout
__global__ void wgrad_reduce_kernel(const m2d_wgrad_args a, const int Ncols, const int splits) {
    long long total = (long long)a.Cout * Ncols;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        // idx enumerates the OUTPUT layout (co, c, t) so that stores are coalesced
        int co = (int)(idx / Ncols);
        int rem = (int)(idx - (long long)co * Ncols);
        int c = rem / a.T;
        int t = rem - c * a.T;
        long long src = a.packed ? idx : (long long)co * Ncols + (long long)t * a.Cc + c;
        float v = 0.f;
        for (int z = 0; z < splits; ++z) v += a.ws[(long long)z * total + src];
        v *= a.scale;
        if (a.beta != 0.f) v += a.beta * a.dw[idx];
        a.dw[idx] = v;
    }
}

// ------------------------------------------------------------------ packing
__global__ void pack_fwd_kernel(const float* __restrict__ w, float* __restrict__ wp, int Cout,
                                int Cin, int k) {
    long long total = (long long)Cout * Cin * k;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        // idx over destination (co, t, ci)
        int ci = (int)(idx % Cin);
        long long r = idx / Cin;
        int t = (int)(r % k);
        int co = (int)(r / k);
        wp[idx] = w[((long long)co * Cin + ci) * k + t];
    }
}

__global__ void pack_bwd_kernel(const float* __restrict__ w, float* __restrict__ wd, int Cout,
                                int Cin, int k, int stride) {
    long long total = (long long)Cout * Cin * k;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        // destination: block rho, then (ci, q, co)
        long long off = 0;
        int rho = 0, Trho = 0;
        for (rho = 0; rho < stride; ++rho) {
            Trho = (k - rho + stride - 1) / stride;
            if (Trho < 0) Trho = 0;
            long long sz = (long long)Cin * Cout * Trho;
            if (idx < off + sz) break;
            off += sz;
        }
        long long loc = idx - off;
        int co = (int)(loc % Cout);
        long long r = loc / Cout;
        int q = (int)(r % Trho);
        int ci = (int)(r / Trho);
        wd[idx] = w[((long long)co * Cin + ci) * k + stride * q + rho];
    }
}

// all weight re-layouts of one network in ONE launch: blockIdx.y selects the table entry.  Besides
// the exact fp32 copy (dst), each entry may request the 3xTF32 operand split of the same matrix
// (dst_tiled: pre-tiled in the tensor core's shared-memory image, see m2d_rowconv_args.w_tiled) consumed by
// the tensor-core kernels.
__global__ void __launch_bounds__(256) pack_batch_kernel(const m2d_pack_desc* __restrict__ table, const bool mixed) {
    const m2d_pack_desc d = table[blockIdx.y];
    const float* __restrict__ w = d.w;
    const int Cout = d.Cout, Cin = d.Cin, k = d.k, stride = d.stride;
    const long long total = (long long)Cout * Cin * k;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        float v;
        int R_ = 128;                                 // rows per block of the tiled copy
        long long pidx;                               // index of the hi value in the tiled split copy
        if (d.kind == M2D_UNPACK_GRAD) {              // dst[co, ci, t] = w[co, t*Cin + ci]  (idx walks dst)
            int t = (int)(idx % k);
            long long r = idx / k;
            int ci = (int)(r % Cin);
            int co = (int)(r / Cin);
            d.dst[idx] = w[((long long)co * k + t) * Cin + ci];
            continue;
        }
        if (d.kind == M2D_PACK_BWD_MERGED) {
            // backward-data of a stride-s convolution as ONE stride-1 row convolution that produces the s fine
            // rows of a coarse row at once: output column (r0, ci), unified tap q' with input row m + cmax - q'
            //   dst[(r0*Cin + ci), q'*Cout + co] = w[co, ci, s*q + rho],  rho = (r0+pad) % s, c0 = (r0+pad) / s,
            //   q' = q + cmax - c0;  taps a residue does not have stay zero (buffers are allocated zeroed)
            const int pad = d.reserved;
            const int j = (int)(idx % k);
            long long r = idx / k;
            const int ci = (int)(r % Cin);
            const int co = (int)(r / Cin);
            const int rho = j % stride, q = j / stride;
            int r0 = (rho - pad) % stride;
            if (r0 < 0) r0 += stride;
            const int c0 = (r0 + pad) / stride, cmax = (stride - 1 + pad) / stride;
            int Tm = 0;
            for (int rr = 0; rr < stride; ++rr) {
                const int rh = (rr + pad) % stride, cc = (rr + pad) / stride;
                const int Tr = (k - rh + stride - 1) / stride + cmax - cc;
                Tm = Tr > Tm ? Tr : Tm;
            }
            const int qp = q + cmax - c0;
            const long long row = (long long)r0 * Cin + ci;
            const float v = w[idx];
            if (d.dst) d.dst[row * Tm * Cout + (long long)qp * Cout + co] = v;
            if (d.dst_tiled) {
                const int R = tiled_rows(stride * Cin);
                store_tiled_split(d.dst_tiled, tiled_index((int)row, qp, co, Tm, Cout, R), R, v, mixed);
            }
            continue;
        }
        if (d.kind == M2D_PACK_FWD) {                 // dst[co, t*Cin + ci]
            int ci = (int)(idx % Cin);
            long long r = idx / Cin;
            int t = (int)(r % k);
            int co = (int)(r / k);
            v = w[((long long)co * Cin + ci) * k + t];
            // Cin == 1: the taps are the contraction channels of a single-tap operand
            pidx = Cin == 1 ? tiled_index(co, 0, t, 1, k, tiled_rows(Cout))
                            : tiled_index(co, t, ci, k, Cin, tiled_rows(Cout));
            R_ = tiled_rows(Cout);
        } else if (d.kind == M2D_PACK_FULL_BWD) {     // dst[(t*Cin + ci), co]
            int co = (int)(idx % Cout);
            long long r = idx / Cout;
            int ci = (int)(r % Cin);
            int t = (int)(r / Cin);
            v = w[((long long)co * Cin + ci) * k + t];
            R_ = tiled_rows(k * Cin);
            pidx = tiled_index((int)r, 0, co, 1, Cout, R_);
        } else {                                      // per stride residue rho: dst_rho[ci, q*Cout + co]
            long long off = 0, poff = 0;
            int rho = 0, Trho = 0;
            R_ = tiled_rows(Cin);
            for (rho = 0; rho < stride; ++rho) {
                Trho = (k - rho + stride - 1) / stride;
                if (Trho < 0) Trho = 0;
                long long sz = (long long)Cin * Cout * Trho;
                if (idx < off + sz) break;
                off += sz;
                poff += tiled_blocks(Cin, Trho, Cout) * tiled_block_floats(R_);
            }
            long long loc = idx - off;
            int co = (int)(loc % Cout);
            long long r = loc / Cout;
            int q = (int)(r % Trho);
            int ci = (int)(r / Trho);
            v = w[((long long)co * Cin + ci) * k + stride * q + rho];
            pidx = poff + tiled_index(ci, q, co, Trho, Cout, R_);
        }
        if (d.dst) d.dst[idx] = v;
        if (d.dst_tiled) store_tiled_split(d.dst_tiled, pidx, R_, v, mixed);
    }
}

// ------------------------------------------------------------------ dgrad, Cin == 1
__global__ void __launch_bounds__(256)
conv_dgrad_c1_kernel(const float* __restrict__ dy, int Lout, int Cout, const float* __restrict__ w,
                     int k, int stride, int pad, float* __restrict__ dx, int Lin) {
    extern __shared__ float ws[];   // [k][Cout]
    for (int i = threadIdx.x; i < k * Cout; i += blockDim.x) {
        int j = i / Cout, co = i - j * Cout;
        ws[i] = w[co * k + j];
    }
    __syncthreads();
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Lin) return;
    const float* dyb = dy + (long long)b * Lout * Cout;
    float acc = 0.f;
    int jmin = (i + pad) % stride;
    for (int j = jmin; j < k; j += stride) {
        int l = (i + pad - j) / stride;
        if (i + pad - j < 0) break;
        if (l >= Lout) continue;
        const float* row = dyb + (long long)l * Cout;
        const float* wj = ws + j * Cout;
        float s = 0.f;
        for (int co = 0; co < Cout; ++co) s = fmaf(__ldg(row + co), wj[co], s);
        acc += s;
    }
    dx[(long long)b * Lin + i] = acc;
}

}  // namespace m2d

namespace m2d {
int rowconv_tc_dispatch(const m2d_rowconv_args& a, int M, int mode, cudaStream_t st);
int wgrad_tc_dispatch(const m2d_wgrad_args& a, int Ktot, int Ncols, int mode, cudaStream_t st);
int conv_c1_fwd_dispatch(const m2d_rowconv_args& a, cudaStream_t st);
int conv_c1_wgrad_dispatch(const m2d_wgrad_args& a, cudaStream_t st);
int conv_c1_dgrad_dispatch(const float* dy, int nb, int Lout, int Cout, const float* w, int k, int stride, int pad,
                           float* dx, int Lin, cudaStream_t st);
int gemm_mode();
}  // namespace m2d

using namespace m2d;

static int run_splitk_epilogue(const m2d_rowconv_args& a, int M, int splits, cudaStream_t st) {
    long long total = (long long)M * a.N;
    int blocks = (int)(cdiv(total, 256) < 4 * kNumSMs ? cdiv(total, 256) : 4 * kNumSMs);
    rowconv_splitk_epilogue<<<blocks, 256, 0, st>>>(a, M, splits);
    return check_launch("rowconv_splitk_epilogue");
}

extern "C" int m2d_rowconv(const m2d_rowconv_args* ap, void* stream) {
    const m2d_rowconv_args& a = *ap;
    cudaStream_t st = (cudaStream_t)stream;
    M2D_REQUIRE(a.x && a.w && a.y, "rowconv: null pointer");
    M2D_REQUIRE(a.nb > 0 && a.y_rows > 0 && a.N > 0 && a.T > 0 && a.Cc > 0, "rowconv: bad dims");
    M2D_REQUIRE(a.win_T == 0 || a.Cc == 1, "rowconv: windowed mode needs Cc == 1");
    M2D_REQUIRE(!a.mask_mode || a.mask, "rowconv: mask_mode without mask");
    long long Mll = (long long)a.nb * a.y_rows;
    M2D_REQUIRE(Mll < (1ll << 31), "rowconv: M too large");
    const int M = (int)Mll;
    const int mode = gemm_mode();
    {   // single-input-channel, 32-output-channel layers: HBM-bound, dedicated FFMA kernel in every mode
        int rc = conv_c1_fwd_dispatch(a, st);
        if (rc <= 0) return rc;
    }
    if (mode != M2D_GEMM_FP32) {
        int rc = rowconv_tc_dispatch(a, M, mode, st);
        if (rc <= 0) return rc;
    }
    {   // a handful of rows against a long K-contiguous contraction: weight streaming over the whole chip
        int rc = skinny_dispatch(a, M, st);
        if (rc <= 0) return rc;
    }
    const bool c1 = a.Cc == 1;
    const bool vec = !c1 && a.Cc % 4 == 0 && a.x_ld % 4 == 0 && a.x_bs % 4 == 0 && a.w_ld % 4 == 0 &&
                     aligned16(a.x) && aligned16(a.w);
    const int cchunks = c1 ? 1 : (int)cdiv(a.Cc, BK);
    const int nsteps = c1 ? (int)cdiv(a.T, BK) : a.T * cchunks;
    const bool big = (long long)cdiv(M, 128) * cdiv(a.N, BN) >= 2 * kNumSMs;
    const int bm = big ? 128 : 64;
    long long tiles = cdiv(M, bm) * cdiv(a.N, BN);
    int splits = 1;
    if (a.ws && tiles < kNumSMs && nsteps >= 8) {
        long long want = cdiv(2 * kNumSMs, tiles);
        long long cap = a.ws_floats / ((long long)M * a.N);
        splits = (int)(want < nsteps / 4 ? want : nsteps / 4);
        if (splits > cap) splits = (int)cap;
        if (splits < 1) splits = 1;
    }
    if (bm == 128) {
        if (c1) launch_rowconv<128, false, true>(a, M, nsteps, cchunks, splits, st);
        else if (vec) launch_rowconv<128, true, false>(a, M, nsteps, cchunks, splits, st);
        else launch_rowconv<128, false, false>(a, M, nsteps, cchunks, splits, st);
    } else {
        if (c1) launch_rowconv<64, false, true>(a, M, nsteps, cchunks, splits, st);
        else if (vec) launch_rowconv<64, true, false>(a, M, nsteps, cchunks, splits, st);
        else launch_rowconv<64, false, false>(a, M, nsteps, cchunks, splits, st);
    }
    int rc = check_launch("rowconv");
    if (rc) return rc;
    if (splits > 1) rc = run_splitk_epilogue(a, M, splits, st);
    return rc;
}

extern "C" long long m2d_wgrad_min_ws(int Cout, int T, int Cc) {
    return (long long)Cout * T * Cc;
}

extern "C" int m2d_wgrad(const m2d_wgrad_args* ap, void* stream) {
    const m2d_wgrad_args& a = *ap;
    cudaStream_t st = (cudaStream_t)stream;
    M2D_REQUIRE(a.dy && a.x && a.dw && a.ws, "wgrad: null pointer");
    M2D_REQUIRE(a.nb > 0 && a.dy_rows > 0 && a.Cout > 0 && a.T > 0 && a.Cc > 0, "wgrad: bad dims");
    M2D_REQUIRE(a.win_T == 0 || a.Cc == 1, "wgrad: windowed mode needs Cc == 1");
    const int Ncols = a.T * a.Cc;
    long long per = (long long)a.Cout * Ncols;
    if (a.ws_floats < per) {
        set_error("wgrad: workspace too small (%lld < %lld floats)", a.ws_floats, per);
        return M2D_ERR_WORKSPACE;
    }
    long long Kll = (long long)a.nb * a.dy_rows;
    M2D_REQUIRE(Kll < (1ll << 31), "wgrad: K too large");
    const int Ktot = (int)Kll;
    const int mode = gemm_mode();
    {
        int rc = conv_c1_wgrad_dispatch(a, st);
        if (rc <= 0) return rc;
    }
    if (mode != M2D_GEMM_FP32) {
        int rc = wgrad_tc_dispatch(a, Ktot, Ncols, mode, st);
        if (rc <= 0) return rc;
    }
    const int nsteps = (int)cdiv(Ktot, BK);
    long long tiles = cdiv(a.Cout, 64) * cdiv(Ncols, BN);
    long long want = cdiv(3 * kNumSMs, tiles);
    long long cap = a.ws_floats / per;
    int splits = (int)(want < nsteps / 4 ? want : nsteps / 4);
    if (splits > cap) splits = (int)cap;
    if (splits > 256) splits = 256;
    if (splits < 1) splits = 1;
    const bool vec = a.win_T == 0 && a.Cc % 4 == 0 && a.x_ld % 4 == 0 && a.x_bs % 4 == 0 &&
                     a.dy_ld % 4 == 0 && a.dy_bs % 4 == 0 && a.Cout % 4 == 0 && aligned16(a.x) &&
                     aligned16(a.dy);
    dim3 grid((unsigned)cdiv(a.Cout, 64), (unsigned)cdiv(Ncols, BN), (unsigned)splits);
    if (vec) wgrad_kernel<true><<<grid, NT, 0, st>>>(a, Ktot, Ncols);
    else wgrad_kernel<false><<<grid, NT, 0, st>>>(a, Ktot, Ncols);
    int rc = check_launch("wgrad");
    if (rc) return rc;
    int blocks = (int)(cdiv(per, 256) < 8 * kNumSMs ? cdiv(per, 256) : 8 * kNumSMs);
    wgrad_reduce_kernel<<<blocks, 256, 0, st>>>(a, Ncols, splits);
    return check_launch("wgrad_reduce");
}

extern "C" int m2d_pack_conv_fwd(const float* w, float* wp, int Cout, int Cin, int k, void* stream) {
    M2D_REQUIRE(w && wp && Cout > 0 && Cin > 0 && k > 0, "pack_conv_fwd: bad args");
    long long total = (long long)Cout * Cin * k;
    int blocks = (int)(cdiv(total, 256) < 8 * kNumSMs ? cdiv(total, 256) : 8 * kNumSMs);
    pack_fwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w, wp, Cout, Cin, k);
    return check_launch("pack_conv_fwd");
}

extern "C" int m2d_pack_conv_bwd(const float* w, float* wd, int Cout, int Cin, int k, int stride,
                                 void* stream) {
    M2D_REQUIRE(w && wd && Cout > 0 && Cin > 0 && k > 0 && stride > 0, "pack_conv_bwd: bad args");
    long long total = (long long)Cout * Cin * k;
    int blocks = (int)(cdiv(total, 256) < 8 * kNumSMs ? cdiv(total, 256) : 8 * kNumSMs);
    pack_bwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w, wd, Cout, Cin, k, stride);
    return check_launch("pack_conv_bwd");
}

extern "C" int m2d_pack_batch(const m2d_pack_desc* table, int n, void* stream) {
    M2D_REQUIRE(table && n > 0, "pack_batch: bad args");
    dim3 grid(148, (unsigned)n);
    pack_batch_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(table, gemm_mode() == M2D_GEMM_TF32_BF16);
    return check_launch("pack_batch");
}

extern "C" int m2d_conv_dgrad_c1(const float* dy, int nb, int Lout, int Cout, const float* w, int k,
                                 int stride, int pad, float* dx, int Lin, void* stream) {
    M2D_REQUIRE(dy && w && dx && nb > 0 && Lout > 0 && Cout > 0 && k > 0 && stride > 0 && Lin > 0,
                "conv_dgrad_c1: bad args");
    {   // 32 channels, k = 25, stride 4 (audio_d.l1): one warp per 4 output samples, taps in registers
        int rc = conv_c1_dgrad_dispatch(dy, nb, Lout, Cout, w, k, stride, pad, dx, Lin, (cudaStream_t)stream);
        if (rc <= 0) return rc;
    }
    dim3 grid((unsigned)cdiv(Lin, 256), (unsigned)nb);
    size_t smem = (size_t)k * Cout * sizeof(float);
    M2D_REQUIRE(smem <= 48 * 1024, "conv_dgrad_c1: k*Cout too large");
    conv_dgrad_c1_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(dy, Lout, Cout, w, k, stride, pad,
                                                                    dx, Lin);
    return check_launch("conv_dgrad_c1");
}
