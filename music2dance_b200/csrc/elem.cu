// HBM-bound kernels: train-mode BatchNorm (stats / apply / backward), column sums,
// WGAN-GP reductions, pose losses, pooling / upsampling, Adam.  All operate on
// channels-last row matrices; column reductions accumulate in fp64.
#include "common.cuh"

namespace m2d {

static inline int grid1d(long long n, int threads = 256, int waves = 8) {
    long long b = cdiv(n, threads);
    long long cap = (long long)waves * kNumSMs;
    return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

// ---------------------------------------------------------------- column tiles
// CTA = 32 columns x 8 row lanes; grid.x = column strips, grid.y = row chunks.
struct ColGrid { dim3 grid; long long rows_per; };
static ColGrid col_grid(long long M, int C) {
    int gx = (int)cdiv(C, 32);
    long long want = cdiv(4 * kNumSMs, gx);
    long long gy = cdiv(M, 64);
    if (gy > want) gy = want;
    if (gy < 1) gy = 1;
    ColGrid g;
    g.rows_per = cdiv(M, gy);
    g.grid = dim3((unsigned)gx, (unsigned)cdiv(M, g.rows_per));
    return g;
}

template <int NQ, typename F>
__device__ __forceinline__ void col_reduce(long long M, int C, long long rows_per, double* acc, F f) {
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    const int ry = threadIdx.x >> 5;
    const long long r0 = blockIdx.y * rows_per;
    const long long r1 = r0 + rows_per < M ? r0 + rows_per : M;
    double s[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) s[q] = 0.0;
    if (c < C) {
        for (long long r = r0 + ry; r < r1; r += 8) {
            float v[NQ];
            f(r, c, v);
#pragma unroll
            for (int q = 0; q < NQ; ++q) s[q] += (double)v[q];
        }
    }
    __shared__ double sh[NQ][8][33];
#pragma unroll
    for (int q = 0; q < NQ; ++q) sh[q][ry][threadIdx.x & 31] = s[q];
    __syncthreads();
    if (ry == 0 && c < C) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            double t = 0.0;
#pragma unroll
            for (int j = 0; j < 8; ++j) t += sh[q][j][threadIdx.x];
            atomicAdd(acc + (long long)q * C + c, t);
        }
    }
}

__global__ void __launch_bounds__(256)
colstats_kernel(const float* __restrict__ x, int ld, long long M, int C, long long rows_per, double* acc) {
    // blockIdx.z = group: M rows each, statistics of group z at acc + z*2C (m2d_colstats_groups)
    x += (long long)blockIdx.z * M * ld;
    acc += (long long)blockIdx.z * 2 * C;
    col_reduce<2>(M, C, rows_per, acc, [&](long long r, int c, float* v) {
        float t = x[r * ld + c];
        v[0] = t;
        v[1] = t * t;
    });
}

__global__ void __launch_bounds__(256)
colsum_kernel(const float* __restrict__ x, int ld, long long M, int C, long long rows_per, double* acc) {
    col_reduce<1>(M, C, rows_per, acc, [&](long long r, int c, float* v) { v[0] = x[r * ld + c]; });
}

// column sums of several matrices in ONE launch (the bias gradients of every layer of a network after a backward
// sweep): blockIdx.y = table entry, blockIdx.x = (row chunk, 32-column block).  fp64 partials -> acc[entry.acc_off + c]
__global__ void __launch_bounds__(256)
colsum_batch_kernel(const m2d_colsum_desc* __restrict__ table, double* acc) {
    const m2d_colsum_desc d = table[blockIdx.y];
    const int ncb = (d.C + 31) >> 5;
    const int nrc = (int)gridDim.x / ncb;                  // row chunks (>= 1: the host sizes gridDim.x >= max ncb)
    if (nrc == 0 || (int)blockIdx.x >= nrc * ncb) return;
    const int cb = blockIdx.x % ncb, rc = blockIdx.x / ncb;
    const int c = cb * 32 + (threadIdx.x & 31), ry = threadIdx.x >> 5;
    const long long rows_per = (d.M + nrc - 1) / nrc;
    const long long r0 = rc * rows_per, r1 = r0 + rows_per < d.M ? r0 + rows_per : d.M;
    double s0 = 0.0, s1 = 0.0;
    if (c < d.C) {
        long long r = r0 + ry;
        for (; r + 24 < r1; r += 32) {                     // four independent 128-byte row reads in flight per warp
            const float a0 = d.x[r * d.ld + c], a1 = d.x[(r + 8) * d.ld + c];
            const float a2 = d.x[(r + 16) * d.ld + c], a3 = d.x[(r + 24) * d.ld + c];
            s0 += (double)a0 + (double)a2;
            s1 += (double)a1 + (double)a3;
        }
        for (; r < r1; r += 8) s0 += (double)d.x[r * d.ld + c];
    }
    __shared__ double sh[8][33];
    sh[ry][threadIdx.x & 31] = s0 + s1;
    __syncthreads();
    if (ry == 0 && c < d.C && r0 < r1) {
        double t = 0.0;
#pragma unroll
        for (int j = 0; j < 8; ++j) t += sh[j][threadIdx.x];
        atomicAdd(acc + d.acc_off + c, t);
    }
}
__global__ void colsum_batch_finalize_kernel(const m2d_colsum_desc* __restrict__ table, double* acc) {
    const m2d_colsum_desc d = table[blockIdx.y];
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < d.C) {
        const float v = (float)(acc[d.acc_off + c] * (double)d.scale);
        d.out[c] = d.beta != 0.f ? d.beta * d.out[c] + v : v;
        acc[d.acc_off + c] = 0.0;                          // ready for the next launch (graph replay)
    }
}

__global__ void colsum_finalize_kernel(const double* acc, int C, float* out, float scale, float beta) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C) {
        float v = (float)(acc[c] * (double)scale);
        out[c] = beta != 0.f ? beta * out[c] + v : v;
    }
}

__global__ void __launch_bounds__(256)
bn_apply_kernel(const float* __restrict__ x, int ldx, float* __restrict__ y, int ldy, long long M, int C,
                long long rows_per, const double* __restrict__ acc, const float* __restrict__ gamma,
                const float* __restrict__ beta, float* running_mean, float* running_var, float momentum,
                float eps, float* mr, int act) {
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    const int ry = threadIdx.x >> 5;
    if (c >= C) return;
    // blockIdx.z = group (m2d_bn_apply_groups): M rows each, normalised with the group's own statistics; the running
    // statistics advance group by group in order, exactly as gridDim.z separate calls would
    if (blockIdx.y == 0 && blockIdx.z == 0 && ry == 0 && running_mean) {
        float rm = running_mean[c], rv = running_var[c];
        for (unsigned z = 0; z < gridDim.z; ++z) {
            const double* a = acc + (long long)z * 2 * C;
            const double mean = a[c] / (double)M;
            double var = a[C + c] / (double)M - mean * mean;
            if (var < 0.0) var = 0.0;
            const double unb = M > 1 ? var * (double)M / (double)(M - 1) : var;
            rm = (1.f - momentum) * rm + momentum * (float)mean;
            rv = (1.f - momentum) * rv + momentum * (float)unb;
        }
        running_mean[c] = rm;
        running_var[c] = rv;
    }
    acc += (long long)blockIdx.z * 2 * C;
    x += (long long)blockIdx.z * M * ldx;
    if (y) y += (long long)blockIdx.z * M * ldy;
    const double mean = acc[c] / (double)M;
    double var = acc[C + c] / (double)M - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    const float mu = (float)mean;
    if (blockIdx.y == 0 && ry == 0 && mr) {
        float* o = mr + (long long)blockIdx.z * 2 * C;
        o[c] = mu;
        o[C + c] = rstd;
    }
    if (!y) return;      // statistics-only call (dead LinearBlock branch, Q1)
    const float g = gamma[c], bt = beta[c];
    const long long r0 = blockIdx.y * rows_per;
    const long long r1 = r0 + rows_per < M ? r0 + rows_per : M;
    for (long long r = r0 + ry; r < r1; r += 8) {
        float v = (x[r * ldx + c] - mu) * rstd * g + bt;
        y[r * ldy + c] = apply_act(v, act);
    }
}

// Train-mode BatchNorm in ONE launch: column statistics (fp64 partials, as colstats_kernel), a grid-wide rendezvous, then
// normalisation + activation (as bn_apply_kernel).  The rendezvous is a counter in global memory that the caller zeroes
// with the accumulators (`sync` = the slot behind the 2C sums); the grid (<= 4 x 148 blocks of 256 threads, col_grid) is
// far below what the chip keeps resident, so every block arrives — blocks of OTHER kernels may delay that, never prevent
// it.  A rendezvous that does not complete (a bug) traps instead of hanging the GPU.
__global__ void __launch_bounds__(256)
bn_train_kernel(const float* __restrict__ x, int ldx, float* __restrict__ y, int ldy, long long M, int C,
                long long rows_per, double* acc, const float* __restrict__ gamma, const float* __restrict__ beta,
                float* running_mean, float* running_var, float momentum, float eps, float* mr, int act,
                unsigned int* sync) {
    col_reduce<2>(M, C, rows_per, acc, [&](long long r, int c, float* v) {
        float t = x[r * ldx + c];
        v[0] = t;
        v[1] = t * t;
    });
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(sync, 1u);
        const unsigned int total = gridDim.x * gridDim.y;
        unsigned int spins = 0;
        while (*(volatile unsigned int*)sync < total)
            if (++spins > (1u << 28)) __trap();
        __threadfence();
    }
    __syncthreads();
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    const int ry = threadIdx.x >> 5;
    if (c >= C) return;
    const double mean = __ldcg(acc + c) / (double)M;
    double var = __ldcg(acc + C + c) / (double)M - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = (float)(1.0 / sqrt(var + (double)eps));
    const float mu = (float)mean;
    if (blockIdx.y == 0 && ry == 0) {
        if (mr) { mr[c] = mu; mr[C + c] = rstd; }
        if (running_mean) {
            double unb = M > 1 ? var * (double)M / (double)(M - 1) : var;
            running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mu;
            running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
        }
    }
    if (!y) return;      // statistics-only call (dead LinearBlock branch, Q1)
    const float g = gamma[c], bt = beta[c];
    const long long r0 = blockIdx.y * rows_per;
    const long long r1 = r0 + rows_per < M ? r0 + rows_per : M;
    for (long long r = r0 + ry; r < r1; r += 8) {
        float v = (x[r * ldx + c] - mu) * rstd * g + bt;
        y[r * ldy + c] = apply_act(v, act);
    }
}

__global__ void __launch_bounds__(256)
bn_eval_kernel(const float* __restrict__ x, int ldx, float* __restrict__ y, int ldy, long long M, int C,
               long long rows_per, const float* gamma, const float* beta, const float* rm,
               const float* rv, float eps, int act) {
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    const int ry = threadIdx.x >> 5;
    if (c >= C) return;
    const float rstd = 1.f / sqrtf(rv[c] + eps);
    const float mu = rm[c], g = gamma[c], bt = beta[c];
    const long long r0 = blockIdx.y * rows_per;
    const long long r1 = r0 + rows_per < M ? r0 + rows_per : M;
    for (long long r = r0 + ry; r < r1; r += 8)
        y[r * ldy + c] = apply_act((x[r * ldx + c] - mu) * rstd * g + bt, act);
}

__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(const float* __restrict__ dy, int lddy, const float* __restrict__ y, int ldy,
                     const float* __restrict__ x, int ldx, long long M, int C, long long rows_per,
                     const float* __restrict__ mr, int act, double* acc) {
    col_reduce<2>(M, C, rows_per, acc, [&](long long r, int c, float* v) {
        float d = dy[r * lddy + c] * act_deriv(y[r * ldy + c], act);
        float xh = (x[r * ldx + c] - mr[c]) * mr[C + c];
        v[0] = d;
        v[1] = d * xh;
    });
}

__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const float* __restrict__ dy, int lddy, const float* __restrict__ y, int ldy,
                    const float* __restrict__ x, int ldx, float* __restrict__ dx, int lddx, long long M,
                    int C, long long rows_per, const float* __restrict__ mr,
                    const float* __restrict__ gamma, int act, const double* __restrict__ acc,
                    float* dgamma, float* dbeta) {
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    const int ry = threadIdx.x >> 5;
    if (c >= C) return;
    const float s0 = (float)(acc[c] / (double)M), s1 = (float)(acc[C + c] / (double)M);
    if (blockIdx.y == 0 && ry == 0) {
        if (dbeta) dbeta[c] = (float)acc[c];
        if (dgamma) dgamma[c] = (float)acc[C + c];
    }
    const float mu = mr[c], rstd = mr[C + c];
    const float k = gamma[c] * rstd;
    const long long r0 = blockIdx.y * rows_per;
    const long long r1 = r0 + rows_per < M ? r0 + rows_per : M;
    for (long long r = r0 + ry; r < r1; r += 8) {
        float d = dy[r * lddy + c] * act_deriv(y[r * ldy + c], act);
        float xh = (x[r * ldx + c] - mu) * rstd;
        dx[r * lddx + c] = k * (d - s0 - xh * s1);
    }
}

// ---------------------------------------------------------------- element-wise
__global__ void axpby_kernel(const float* __restrict__ x, const float* __restrict__ z, float* y,
                             long long n, float a, float b) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        y[i] = z ? a * x[i] + b * z[i] : a * x[i];
}

__global__ void fill_kernel(float* y, long long n, float v) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        y[i] = v;
}

__global__ void scale_rows_kernel(const float* __restrict__ x, const float* __restrict__ s, float* y,
                                  long long per) {
    const float k = s[blockIdx.y];
    const long long base = blockIdx.y * per;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per;
         i += (long long)gridDim.x * blockDim.x)
        y[base + i] = k * x[base + i];
}

__global__ void interp_kernel(const float* __restrict__ real, const float* __restrict__ fake,
                              const float* __restrict__ alpha, float* xi, long long per) {
    const float a = alpha[blockIdx.y];
    const long long base = blockIdx.y * per;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per;
         i += (long long)gridDim.x * blockDim.x)
        xi[base + i] = a * real[base + i] + (1.f - a) * fake[base + i];
}

// xi3 = [interpolates; real; fake] (3 x nb entries): the stacked pose input of one critic iteration in one pass
__global__ void interp_stack3_kernel(const float* __restrict__ real, const float* __restrict__ fake,
                                     const float* __restrict__ alpha, float* xi3, int nb, long long per) {
    const float a = alpha[blockIdx.y];
    const long long base = blockIdx.y * per, all = (long long)nb * per;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per;
         i += (long long)gridDim.x * blockDim.x) {
        const float r = real[base + i], f = fake[base + i];
        xi3[base + i] = a * r + (1.f - a) * f;
        xi3[all + base + i] = r;
        xi3[2 * all + base + i] = f;
    }
}

__device__ __forceinline__ void block_atomic_add(double v, double* out) {
    __shared__ double sh[32];
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        double t = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
        t = warp_sum(t);
        if (threadIdx.x == 0) atomicAdd(out, t);
    }
    __syncthreads();
}

__global__ void rows_sumsq_kernel(const float* __restrict__ x, long long per, double* out) {
    const long long base = blockIdx.y * per;
    double s = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < per;
         i += (long long)gridDim.x * blockDim.x) {
        float v = x[base + i];
        s += (double)v * (double)v;
    }
    block_atomic_add(s, out + blockIdx.y);
}

__global__ void sum_kernel(const float* __restrict__ x, long long n, double* out) {
    double s = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        s += (double)x[i];
    block_atomic_add(s, out);
}

__global__ void gp_finalize_kernel(const double* ss0, const double* ss1, int B, float* gp, float* k0,
                                   float* k1, float kscale) {
    // single block; B is a minibatch (<= a few thousand)
    double acc = 0.0;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        float n0 = sqrtf((float)ss0[b] + 1e-12f);
        float d0 = n0 - 1.f;
        acc += (double)(d0 * d0);
        k0[b] = kscale * ((2.f / (float)B) * d0 / n0);
        if (ss1) {
            float n1 = sqrtf((float)ss1[b] + 1e-12f);
            float d1 = n1 - 1.f;
            acc += (double)(d1 * d1);
            k1[b] = kscale * ((2.f / (float)B) * d1 / n1);
        }
    }
    __shared__ double total;
    if (threadIdx.x == 0) total = 0.0;
    __syncthreads();
    block_atomic_add(acc, &total);
    if (threadIdx.x == 0) gp[0] = (float)(total / (double)B);
}

// losses.py:47-50, WGAN-LP: bgrad = max(0, ||g||_2 - 1) (no epsilon under the root), penalty = mean(bgrad^2);
// kappa[b] = d penalty / d g[b,:] / g[b,:] = (2/B) bgrad / ||g||  (0 where the norm is below 1)
__global__ void gp_finalize_lp_kernel(const double* ss0, int B, float* gp, float* k0) {
    double acc = 0.0;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        float n0 = sqrtf((float)ss0[b]);
        float d0 = n0 - 1.f;
        d0 = d0 < 0.f ? 0.f : d0;
        acc += (double)(d0 * d0);
        k0[b] = d0 > 0.f ? (2.f / (float)B) * d0 / n0 : 0.f;
    }
    __shared__ double total;
    if (threadIdx.x == 0) total = 0.0;
    __syncthreads();
    block_atomic_add(acc, &total);
    if (threadIdx.x == 0) gp[0] = (float)(total / (double)B);
}

__device__ __forceinline__ float sgn(float v) { return (v > 0.f) - (v < 0.f); }

__global__ void pose_losses_kernel(const float* __restrict__ real, const float* __restrict__ fake,
                                   float* dfake, int B, int T, int C, float beta, float eta,
                                   int accumulate, double* acc) {
    const long long n = (long long)B * T * C;
    const float kl1 = beta / (float)n;
    const float ktv = T > 1 ? eta / (float)((long long)B * (T - 1) * C) : 0.f;
    double s1 = 0.0, s2 = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        int t = (int)((i / C) % T);
        float f = fake[i];
        float d = real[i] - f;
        s1 += (double)fabsf(d);
        float g = -kl1 * sgn(d);
        if (t + 1 < T) {
            float e = fake[i + C] - f;
            s2 += (double)fabsf(e);
            g -= ktv * sgn(e);
        }
        if (t > 0) g += ktv * sgn(f - fake[i - C]);
        if (dfake) dfake[i] = accumulate ? dfake[i] + g : g;
    }
    block_atomic_add(s1, acc);
    block_atomic_add(s2, acc + 1);
}

// losses.py:85-89 jerkiness on channels-last poses [B,T,C]: acc[0] += sum_{b, t < T-3, c} (third difference)^2
__global__ void jerk_kernel(const float* __restrict__ x, int B, int T, int C, double* acc) {
    const long long n = (long long)B * (T - 3) * C;
    double s = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const long long r = i / C;
        const int t = (int)(r % (T - 3));
        const long long b = r / (T - 3);
        const float* p = x + (b * T + t) * C + c;
        const float d = p[3 * C] - 3.f * p[2 * C] + 3.f * p[C] - p[0];
        s += (double)d * (double)d;
    }
    block_atomic_add(s, acc);
}

__global__ void act_bwd_kernel(float* d, const float* __restrict__ y, long long n, int mode) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        d[i] *= act_deriv(y[i], mode);
}

// out[r, c] = alpha * a[r, c] * b[r, c] * c3[r, c] on strided row matrices (second-order tanh term of the
// gradient penalty: -2 * code * tangent * dD/dcode)
__global__ void mul3_kernel(const float* __restrict__ a, int lda, const float* __restrict__ b, int ldb,
                            const float* __restrict__ c3, int ldc, float* out, int ldo, int C,
                            long long total, float alpha) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        long long r = i / C;
        int c = (int)(i - r * C);
        out[r * ldo + c] = alpha * a[r * lda + c] * b[r * ldb + c] * c3[r * ldc + c];
    }
}

// ---------------------------------------------------------------- label embedding (phase2 conditional)
// y[(b*T + t)*ldy + e] = table[labels[b]*E + e]: the label code of sequence b broadcast over its T rows
// (nn.Embedding lookup + unsqueeze + expand + cat of phase2/archis/conditional.py:19-21,45-46 written straight
// into the concatenated operand's label columns).  err[0] is set if a label is outside [0, n_classes).
__global__ void embed_rows_kernel(const float* __restrict__ table, const long long* __restrict__ labels, float* y,
                                  int ldy, int T, int E, int n_classes, long long total, int* err) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        int e = (int)(i % E);
        long long r = i / E;
        long long lab = labels[r / T];
        if (lab < 0 || lab >= n_classes) {
            if (err) *err = 1;
            y[r * ldy + e] = 0.f;
        } else {
            y[r * ldy + e] = table[lab * E + e];
        }
    }
}

// dtable[c, e] = beta * dtable[c, e] + scale * sum_{b: labels[b] == c} sum_t dy[(b*T + t)*ldd + e]
// (backward of the lookup above).  One CTA per (class, column); fixed-order fp64 tree -> deterministic.
__global__ void embed_grad_kernel(const float* __restrict__ dy, int ldd, const long long* __restrict__ labels,
                                  float* dtable, int B, int T, int E, float scale, float beta) {
    int c = blockIdx.x / E, e = blockIdx.x % E;
    double acc = 0.0;
    for (int b = 0; b < B; ++b) {
        if (labels[b] != c) continue;                      // uniform across the CTA
        const float* p = dy + (long long)b * T * ldd + e;
        for (int t = threadIdx.x; t < T; t += blockDim.x) acc += (double)p[(long long)t * ldd];
    }
    __shared__ double red[256];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = blockDim.x / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        float* o = dtable + c * E + e;
        *o = (beta == 0.f ? 0.f : beta * *o) + scale * (float)red[0];
    }
}

// ---------------------------------------------------------------- pool / upsample
__global__ void maxpool2_kernel(const float* __restrict__ x, int ldx, float* y, int ldy, int Lin,
                                int Lout, int C, long long total) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        int c = (int)(i % C);
        long long r = i / C;
        int l = (int)(r % Lout);
        long long b = r / Lout;
        const float* p = x + (b * Lin + 2 * l) * (long long)ldx + c;
        y[(b * Lout + l) * (long long)ldy + c] = fmaxf(p[0], p[ldx]);
    }
}

__global__ void maxpool2_bwd_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ dy,
                                    int lddy, float* dx, int lddx, int Lin, int Lout, int C,
                                    long long total, int accumulate) {
    // one thread per INPUT element
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        int c = (int)(i % C);
        long long r = i / C;
        int li = (int)(r % Lin);
        long long b = r / Lin;
        int l = li >> 1;
        float g = 0.f;
        if (l < Lout) {
            const float* p = x + (b * Lin + 2 * l) * (long long)ldx + c;
            bool first = p[0] >= p[ldx];   // ties go to the first element (PyTorch argmax)
            bool mine = (li & 1) ? !first : first;
            if (mine) g = dy[(b * Lout + l) * (long long)lddy + c];
        }
        float* o = dx + (b * Lin + li) * (long long)lddx + c;
        *o = accumulate ? *o + g : g;
    }
}

// Upsample x2, linear, align_corners=False: src = (i+0.5)/2 - 0.5 clamped at 0
__global__ void upsample2_kernel(const float* __restrict__ x, int ldx, float* y, int ldy, int Lin, int C,
                                 long long total) {
    const int Lout = 2 * Lin;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        int c = (int)(i % C);
        long long r = i / C;
        int lo = (int)(r % Lout);
        long long b = r / Lout;
        float src = (lo + 0.5f) * 0.5f - 0.5f;
        if (src < 0.f) src = 0.f;
        int i0 = (int)src;
        int i1 = i0 + (i0 < Lin - 1 ? 1 : 0);
        float l1 = src - (float)i0, l0 = 1.f - l1;
        const float* p = x + b * Lin * (long long)ldx + c;
        y[(b * Lout + lo) * (long long)ldy + c] = l0 * p[(long long)i0 * ldx] + l1 * p[(long long)i1 * ldx];
    }
}

__global__ void upsample2_bwd_kernel(const float* __restrict__ dy, int lddy, float* dx, int lddx, int Lin,
                                     int C, long long total, int accumulate) {
    // one thread per INPUT element m: gathers from outputs 2m-2 .. 2m+2
    const int Lout = 2 * Lin;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        int c = (int)(i % C);
        long long r = i / C;
        int m = (int)(r % Lin);
        long long b = r / Lin;
        const float* q = dy + b * Lout * (long long)lddy + c;
        float g = 0.f;
        int lo_begin = 2 * m - 2 < 0 ? 0 : 2 * m - 2;
        int lo_end = 2 * m + 2 >= Lout ? Lout - 1 : 2 * m + 2;
        for (int lo = lo_begin; lo <= lo_end; ++lo) {
            float src = (lo + 0.5f) * 0.5f - 0.5f;
            if (src < 0.f) src = 0.f;
            int i0 = (int)src;
            int i1 = i0 + (i0 < Lin - 1 ? 1 : 0);
            float l1 = src - (float)i0, l0 = 1.f - l1;
            float w = (i0 == m ? l0 : 0.f) + (i1 == m ? l1 : 0.f);
            if (w != 0.f) g += w * q[(long long)lo * lddy];
        }
        float* o = dx + (b * Lin + m) * (long long)lddx + c;
        *o = accumulate ? *o + g : g;
    }
}

__global__ void copy2d_kernel(const float* __restrict__ x, int ldx, float* y, int ldy, long long M, int C,
                              int accumulate) {
    long long total = M * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        int c = (int)(i % C);
        long long r = i / C;
        float v = x[r * ldx + c];
        float* o = y + r * ldy + c;
        *o = accumulate ? *o + v : v;
    }
}

// up to M2D_COPY2D_MAX strided copies in ONE launch, applied in table order (thread t touches the same element index
// of every entry, so an accumulating entry may follow the copy that initialises its destination)
struct Copy2dBatch { m2d_copy2d_desc e[M2D_COPY2D_MAX]; int n; };
__global__ void copy2d_batch_kernel(const Copy2dBatch b) {
    for (int k = 0; k < b.n; ++k) {
        const m2d_copy2d_desc d = b.e[k];
        const long long total = d.M * d.C;
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
             i += (long long)gridDim.x * blockDim.x) {
            const int c = (int)(i % d.C);
            const long long r = i / d.C;
            const float v = d.x[r * d.ldx + c];
            float* o = d.y + r * d.ldy + c;
            *o = d.accumulate ? *o + v : v;
        }
    }
}

// The critic's fusion MLP (default.py:339-345: Linear(F,H) + ReLU, Linear(H,1)) on a handful of rows, forward and —
// when the upstream of the scores is known in advance (dd != null: the WGAN-GP critic step, where it is a constant) —
// its backward-data in the same launch: one block per row, fp32 CUDA-core arithmetic (0.1 MFLOP per row).
//   u = relu(W1 x + b1), d = W2 u + b2;   dh = dd * W2 * [u > 0],  dx = W1^T dh
constexpr int FUSION_MAXH = 256;
__global__ void __launch_bounds__(256)
fusion_mlp_kernel(const float* __restrict__ x, int ldx, int F, int H, const float* __restrict__ w1,
                  const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
                  const float* __restrict__ dd, float* __restrict__ u, float* __restrict__ d, float* __restrict__ dh,
                  float* __restrict__ dx, int lddx) {
    extern __shared__ float fus_x[];                       // F floats
    __shared__ float us[FUSION_MAXH], dhs[FUSION_MAXH];
    const int row = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int f = tid; f < F; f += 256) fus_x[f] = x[(long long)row * ldx + f];
    __syncthreads();
    for (int j = warp; j < H; j += 8) {
        const float* wr = w1 + (long long)j * F;
        float a = 0.f;
        for (int f = lane; f < F; f += 32) a = fmaf(wr[f], fus_x[f], a);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) {
            const float v = fmaxf(a + b1[j], 0.f);
            us[j] = v;
            u[(long long)row * H + j] = v;
        }
    }
    __syncthreads();
    if (warp == 0) {
        float a = 0.f;
        for (int j = lane; j < H; j += 32) a = fmaf(w2[j], us[j], a);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (lane == 0) d[row] = a + b2[0];
    }
    if (!dd) return;
    const float up = dd[row];
    for (int j = tid; j < H; j += 256) {
        const float g = us[j] > 0.f ? up * w2[j] : 0.f;
        dhs[j] = g;
        dh[(long long)row * H + j] = g;
    }
    __syncthreads();
    for (int f = tid; f < F; f += 256) {
        float a = 0.f;
        for (int j = 0; j < H; ++j) a = fmaf(dhs[j], w1[(long long)j * F + f], a);
        dx[(long long)row * lddx + f] = a;
    }
}

__global__ void transpose_kernel(const float* __restrict__ x, float* y, int R, int C) {
    __shared__ float tile[32][33];
    const float* xb = x + (long long)blockIdx.z * R * C;
    float* yb = y + (long long)blockIdx.z * R * C;
    int c = blockIdx.x * 32 + threadIdx.x;
    for (int j = threadIdx.y; j < 32; j += 8) {
        int r = blockIdx.y * 32 + j;
        if (r < R && c < C) tile[j][threadIdx.x] = xb[(long long)r * C + c];
    }
    __syncthreads();
    int r2 = blockIdx.y * 32 + threadIdx.x;
    for (int j = threadIdx.y; j < 32; j += 8) {
        int c2 = blockIdx.x * 32 + j;
        if (r2 < R && c2 < C) yb[(long long)c2 * R + r2] = tile[threadIdx.x][j];
    }
}

// train.py:207-214,226-235: scalar losses of one critic iteration / generator update
__global__ void wgan_scalars_kernel(const double* sums, const float* gp, int B, long long n_l1, long long n_tv,
                                    float c0, float c1, int mode, float* out, const float* d_real,
                                    const float* d_fake) {
    // one warp; d_real / d_fake (B critic scores each) replace sums[0] / sums[1] when given
    double sr = 0.0, sf = 0.0;
    if (d_real) {
        for (int b = threadIdx.x; b < B; b += 32) { sr += (double)d_real[b]; sf += (double)d_fake[b]; }
        sr = warp_sum(sr);
        sf = warp_sum(sf);
    } else {
        sr = sums[0];
        sf = sums[1];
    }
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const float er = (float)(sr / (double)B), ef = (float)(sf / (double)B);
    if (mode == 0) {            // critic: err_fake - err_real + gamma*gp
        const float g = gp[0];
        out[0] = ef - er + c0 * g; out[1] = g; out[2] = ef - er; out[3] = er; out[4] = ef;
    } else {                    // generator: err_real - err_fake + beta*l1 + eta*tv
        const float l1 = (float)(sums[2] / (double)n_l1);
        const float tv = n_tv > 0 ? (float)(sums[3] / (double)n_tv) : 0.f;
        out[0] = er - ef + c0 * l1 + c1 * tv; out[1] = l1; out[2] = tv; out[3] = er; out[4] = ef;
    }
}

// utils.py:329-353 slice_audio_batch as a standalone gather (bit-exact copy)
__global__ void slice_audio_kernel(const float* __restrict__ audio, float* __restrict__ out, int A, int nwin,
                                   int W, int stride, int pad_left) {
    const long long seq = blockIdx.z;
    const int f = blockIdx.y;
    const float* a = audio + seq * A;
    float* o = out + (seq * nwin + f) * (long long)W;
    const int start = f * stride - pad_left;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < W; j += gridDim.x * blockDim.x) {
        int pos = start + j;
        o[j] = (pos >= 0 && pos < A) ? a[pos] : 0.f;
    }
}

// SequenceDataset.__getitem__ + collate_fn (utils.py:91-101,128-144) on a device-resident dataset: batch entry b
// is frames [start, start + T) of sequence seq[b] and audio samples [start*ratio, start*ratio + A) of its music.
// Pure indexing: bit-exact.  blockIdx.y = batch entry; blockIdx.x strides over the entry's T*O + A floats.
__global__ void crop_batch_kernel(const float* __restrict__ poses, const long long* __restrict__ pose_off,
                                  const float* __restrict__ music, const long long* __restrict__ music_off,
                                  const int* __restrict__ seq, const int* __restrict__ start, int TO, int O,
                                  int ratio, int A, float* __restrict__ real, float* __restrict__ audio) {
    const int b = blockIdx.y;
    const int s = seq[b], st = start[b];
    const float* p = poses + pose_off[s] + (long long)st * O;
    const float* m = music + music_off[s] + (long long)st * ratio;
    float* r = real + (long long)b * TO;
    float* a = audio + (long long)b * A;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < TO + A; i += gridDim.x * blockDim.x) {
        if (i < TO) r[i] = p[i];
        else a[i - TO] = m[i - TO];
    }
}

// ---------------------------------------------------------------- Adam
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long long n, const int* step, float lr, float b1,
                            float b2, float eps, float gscale) {
    // step counter was already incremented for this update
    const AdamConst c = adam_const(*step, lr, b1, b2, eps, gscale);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        float mi = m[i], vi = v[i];
        p[i] = adam_update(p[i], g[i], mi, vi, c);
        m[i] = mi;
        v[i] = vi;
    }
}
__global__ void tick_kernel(int* step) { *step += 1; }
__global__ void timestamp_kernel(unsigned long long* slot) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    *slot = t;
}

}  // namespace m2d

using namespace m2d;

extern "C" int m2d_colstats_groups(const float* x, int ld, long long M, int C, int groups, double* acc, void* stream) {
    M2D_REQUIRE(x && acc && M > 0 && C > 0 && groups > 0 && groups <= 65535, "colstats: bad args");
    ColGrid g = col_grid(M, C);
    g.grid.z = (unsigned)groups;
    colstats_kernel<<<g.grid, 256, 0, (cudaStream_t)stream>>>(x, ld, M, C, g.rows_per, acc);
    return check_launch("colstats");
}
extern "C" int m2d_colstats(const float* x, int ld, long long M, int C, double* acc, void* stream) {
    return m2d_colstats_groups(x, ld, M, C, 1, acc, stream);
}

extern "C" int m2d_bn_apply_groups(const float* x, int ldx, float* y, int ldy, long long M, int C, int groups,
                                   const double* acc, const float* gamma, const float* beta,
                                   float* running_mean, float* running_var, float momentum, float eps,
                                   float* mr, int act, void* stream) {
    M2D_REQUIRE(x && acc && gamma && beta && M > 0 && C > 0 && groups > 0 && groups <= 65535, "bn_apply: bad args");
    M2D_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "bn_apply: running_mean / running_var go together");
    ColGrid g = col_grid(M, C);
    g.grid.z = (unsigned)groups;
    bn_apply_kernel<<<g.grid, 256, 0, (cudaStream_t)stream>>>(x, ldx, y, ldy, M, C, g.rows_per, acc, gamma,
                                                            beta, running_mean, running_var, momentum,
                                                            eps, mr, act);
    return check_launch("bn_apply");
}
extern "C" int m2d_bn_apply(const float* x, int ldx, float* y, int ldy, long long M, int C,
                            const double* acc, const float* gamma, const float* beta,
                            float* running_mean, float* running_var, float momentum, float eps,
                            float* mr, int act, void* stream) {
    return m2d_bn_apply_groups(x, ldx, y, ldy, M, C, 1, acc, gamma, beta, running_mean, running_var, momentum, eps,
                               mr, act, stream);
}

extern "C" int m2d_bn_train(const float* x, int ldx, float* y, int ldy, long long M, int C, double* acc,
                            const float* gamma, const float* beta, float* running_mean, float* running_var,
                            float momentum, float eps, float* mr, int act, void* stream) {
    M2D_REQUIRE(x && acc && gamma && beta && M > 0 && C > 0, "bn_train: bad args");
    ColGrid g = col_grid(M, C);
    bn_train_kernel<<<g.grid, 256, 0, (cudaStream_t)stream>>>(x, ldx, y, ldy, M, C, g.rows_per, acc, gamma, beta,
                                                            running_mean, running_var, momentum, eps, mr, act,
                                                            reinterpret_cast<unsigned int*>(acc + 2 * C));
    return check_launch("bn_train");
}

extern "C" int m2d_bn_eval(const float* x, int ldx, float* y, int ldy, long long M, int C,
                           const float* gamma, const float* beta, const float* running_mean,
                           const float* running_var, float eps, int act, void* stream) {
    M2D_REQUIRE(x && y && gamma && beta && running_mean && running_var && M > 0 && C > 0, "bn_eval: bad args");
    ColGrid g = col_grid(M, C);
    bn_eval_kernel<<<g.grid, 256, 0, (cudaStream_t)stream>>>(x, ldx, y, ldy, M, C, g.rows_per, gamma, beta,
                                                           running_mean, running_var, eps, act);
    return check_launch("bn_eval");
}

extern "C" int m2d_bn_bwd_reduce(const float* dy, int lddy, const float* y, int ldy, const float* x,
                                 int ldx, long long M, int C, const float* mr, int act, double* acc,
                                 void* stream) {
    M2D_REQUIRE(dy && y && x && mr && acc && M > 0 && C > 0, "bn_bwd_reduce: bad args");
    ColGrid g = col_grid(M, C);
    bn_bwd_reduce_kernel<<<g.grid, 256, 0, (cudaStream_t)stream>>>(dy, lddy, y, ldy, x, ldx, M, C,
                                                                 g.rows_per, mr, act, acc);
    return check_launch("bn_bwd_reduce");
}

extern "C" int m2d_bn_bwd_apply(const float* dy, int lddy, const float* y, int ldy, const float* x,
                                int ldx, float* dx, int lddx, long long M, int C, const float* mr,
                                const float* gamma, int act, const double* acc, float* dgamma,
                                float* dbeta, void* stream) {
    M2D_REQUIRE(dy && y && x && dx && mr && gamma && acc && M > 0 && C > 0, "bn_bwd_apply: bad args");
    ColGrid g = col_grid(M, C);
    bn_bwd_apply_kernel<<<g.grid, 256, 0, (cudaStream_t)stream>>>(dy, lddy, y, ldy, x, ldx, dx, lddx, M, C,
                                                                g.rows_per, mr, gamma, act, acc, dgamma,
                                                                dbeta);
    return check_launch("bn_bwd_apply");
}

extern "C" int m2d_colsum(const float* x, int ld, long long M, int C, float* out, float scale,
                          float beta, double* acc, void* stream) {
    M2D_REQUIRE(x && out && acc && M > 0 && C > 0, "colsum: bad args");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(acc, 0, sizeof(double) * C, st);
    if (e != cudaSuccess) { set_error("colsum memset: %s", cudaGetErrorString(e)); return M2D_ERR_CUDA; }
    ColGrid g = col_grid(M, C);
    colsum_kernel<<<g.grid, 256, 0, st>>>(x, ld, M, C, g.rows_per, acc);
    colsum_finalize_kernel<<<(C + 255) / 256, 256, 0, st>>>(acc, C, out, scale, beta);
    return check_launch("colsum");
}

extern "C" int m2d_colsum_batch(const m2d_colsum_desc* table, int n, int max_C, double* acc, void* stream) {
    M2D_REQUIRE(table && acc && n > 0 && max_C > 0, "colsum_batch: bad args");
    cudaStream_t st = (cudaStream_t)stream;
    // 4 x 148 blocks per entry: the largest entry (audio_d.l1's deltas, 134 400 rows x 32 columns at batch 7) is then ~16
    // dependent row reads per warp instead of 130 (the first version took 114 us for it: latency-bound)
    const int ncb = (max_C + 31) / 32;
    dim3 grid((unsigned)(ncb > 4 * kNumSMs ? ncb : 4 * kNumSMs), (unsigned)n);
    colsum_batch_kernel<<<grid, 256, 0, st>>>(table, acc);
    dim3 g2((unsigned)((max_C + 255) / 256), (unsigned)n);
    colsum_batch_finalize_kernel<<<g2, 256, 0, st>>>(table, acc);
    return check_launch("colsum_batch");
}

extern "C" int m2d_axpby(const float* x, const float* z, float* y, long long n, float a, float b,
                         void* stream) {
    M2D_REQUIRE(x && y && n > 0, "axpby: bad args");
    axpby_kernel<<<grid1d(n), 256, 0, (cudaStream_t)stream>>>(x, z, y, n, a, b);
    return check_launch("axpby");
}

extern "C" int m2d_fill(float* y, long long n, float v, void* stream) {
    M2D_REQUIRE(y && n > 0, "fill: bad args");
    fill_kernel<<<grid1d(n), 256, 0, (cudaStream_t)stream>>>(y, n, v);
    return check_launch("fill");
}

extern "C" int m2d_scale_rows(const float* x, const float* s, float* y, int nb, long long per,
                              void* stream) {
    M2D_REQUIRE(x && s && y && nb > 0 && per > 0, "scale_rows: bad args");
    dim3 grid((unsigned)grid1d(per, 256, 2), (unsigned)nb);
    scale_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, s, y, per);
    return check_launch("scale_rows");
}

extern "C" int m2d_interp(const float* real, const float* fake, const float* alpha, float* xi, int nb,
                          long long per, void* stream) {
    M2D_REQUIRE(real && fake && alpha && xi && nb > 0 && per > 0, "interp: bad args");
    dim3 grid((unsigned)grid1d(per, 256, 2), (unsigned)nb);
    interp_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(real, fake, alpha, xi, per);
    return check_launch("interp");
}

extern "C" int m2d_interp_stack3(const float* real, const float* fake, const float* alpha, float* xi3, int nb,
                                 long long per, void* stream) {
    M2D_REQUIRE(real && fake && alpha && xi3 && nb > 0 && per > 0, "interp_stack3: bad args");
    dim3 grid((unsigned)grid1d(per, 256, 2), (unsigned)nb);
    interp_stack3_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(real, fake, alpha, xi3, nb, per);
    return check_launch("interp_stack3");
}

extern "C" int m2d_rows_sumsq(const float* x, int nb, long long per, double* out, void* stream) {
    M2D_REQUIRE(x && out && nb > 0 && per > 0, "rows_sumsq: bad args");
    dim3 grid((unsigned)grid1d(per, 256, 1), (unsigned)nb);
    rows_sumsq_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, per, out);
    return check_launch("rows_sumsq");
}

extern "C" int m2d_sum(const float* x, long long n, double* out, void* stream) {
    M2D_REQUIRE(x && out && n > 0, "sum: bad args");
    sum_kernel<<<grid1d(n, 256, 1), 256, 0, (cudaStream_t)stream>>>(x, n, out);
    return check_launch("sum");
}

extern "C" int m2d_gp_finalize(const double* ss0, const double* ss1, int B, float* gp, float* kappa0,
                               float* kappa1, float kscale, void* stream) {
    M2D_REQUIRE(ss0 && gp && kappa0 && B > 0 && (!ss1 || kappa1), "gp_finalize: bad args");
    gp_finalize_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(ss0, ss1, B, gp, kappa0, kappa1, kscale);
    return check_launch("gp_finalize");
}

extern "C" int m2d_gp_finalize_lp(const double* ss0, int B, float* gp, float* kappa0, void* stream) {
    M2D_REQUIRE(ss0 && gp && kappa0 && B > 0, "gp_finalize_lp: bad args");
    gp_finalize_lp_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(ss0, B, gp, kappa0);
    return check_launch("gp_finalize_lp");
}

extern "C" int m2d_pose_losses(const float* real, const float* fake, float* dfake, int B, int T, int C,
                               float beta, float eta, int accumulate, double* acc, void* stream) {
    M2D_REQUIRE(real && fake && acc && B > 0 && T > 0 && C > 0, "pose_losses: bad args");
    long long n = (long long)B * T * C;
    pose_losses_kernel<<<grid1d(n, 256, 2), 256, 0, (cudaStream_t)stream>>>(real, fake, dfake, B, T, C, beta,
                                                                        eta, accumulate, acc);
    return check_launch("pose_losses");
}

extern "C" int m2d_jerkiness(const float* x, int B, int T, int C, double* acc, void* stream) {
    M2D_REQUIRE(x && acc && B > 0 && T > 3 && C > 0, "jerkiness: bad args (needs T > 3)");
    jerk_kernel<<<grid1d((long long)B * (T - 3) * C, 256, 2), 256, 0, (cudaStream_t)stream>>>(x, B, T, C, acc);
    return check_launch("jerkiness");
}

extern "C" int m2d_act_bwd(float* d, const float* y, long long n, int mask_mode, void* stream) {
    M2D_REQUIRE(d && y && n > 0, "act_bwd: bad args");
    act_bwd_kernel<<<grid1d(n), 256, 0, (cudaStream_t)stream>>>(d, y, n, mask_mode);
    return check_launch("act_bwd");
}

extern "C" int m2d_mul3(const float* a, int lda, const float* b, int ldb, const float* c, int ldc, float* out,
                        int ldo, long long M, int C, float alpha, void* stream) {
    M2D_REQUIRE(a && b && c && out && M > 0 && C > 0, "mul3: bad args");
    long long total = M * C;
    mul3_kernel<<<grid1d(total), 256, 0, (cudaStream_t)stream>>>(a, lda, b, ldb, c, ldc, out, ldo, C, total, alpha);
    return check_launch("mul3");
}

extern "C" int m2d_embed_rows(const float* table, const long long* labels, float* y, int ldy, int B, int T, int E,
                              int n_classes, int* err, void* stream) {
    M2D_REQUIRE(table && labels && y && B > 0 && T > 0 && E > 0 && ldy >= E && n_classes > 0, "embed_rows: bad args");
    long long total = (long long)B * T * E;
    embed_rows_kernel<<<grid1d(total), 256, 0, (cudaStream_t)stream>>>(table, labels, y, ldy, T, E, n_classes, total,
                                                                      err);
    return check_launch("embed_rows");
}

extern "C" int m2d_embed_grad(const float* dy, int ldd, const long long* labels, float* dtable, int B, int T, int E,
                              int n_classes, float scale, float beta, void* stream) {
    M2D_REQUIRE(dy && labels && dtable && B > 0 && T > 0 && E > 0 && ldd >= E && n_classes > 0, "embed_grad: bad args");
    embed_grad_kernel<<<n_classes * E, 256, 0, (cudaStream_t)stream>>>(dy, ldd, labels, dtable, B, T, E, scale, beta);
    return check_launch("embed_grad");
}

extern "C" int m2d_maxpool2(const float* x, int ldx, float* y, int ldy, int nb, int Lin, int C,
                            void* stream) {
    M2D_REQUIRE(x && y && nb > 0 && Lin > 1 && C > 0, "maxpool2: bad args");
    int Lout = Lin / 2;
    long long total = (long long)nb * Lout * C;
    maxpool2_kernel<<<grid1d(total), 256, 0, (cudaStream_t)stream>>>(x, ldx, y, ldy, Lin, Lout, C, total);
    return check_launch("maxpool2");
}

extern "C" int m2d_maxpool2_bwd(const float* x, int ldx, const float* dy, int lddy, float* dx, int lddx,
                                int nb, int Lin, int C, int accumulate, void* stream) {
    M2D_REQUIRE(x && dy && dx && nb > 0 && Lin > 1 && C > 0, "maxpool2_bwd: bad args");
    int Lout = Lin / 2;
    long long total = (long long)nb * Lin * C;
    maxpool2_bwd_kernel<<<grid1d(total), 256, 0, (cudaStream_t)stream>>>(x, ldx, dy, lddy, dx, lddx, Lin,
                                                                       Lout, C, total, accumulate);
    return check_launch("maxpool2_bwd");
}

extern "C" int m2d_upsample2(const float* x, int ldx, float* y, int ldy, int nb, int Lin, int C,
                             void* stream) {
    M2D_REQUIRE(x && y && nb > 0 && Lin > 0 && C > 0, "upsample2: bad args");
    long long total = (long long)nb * 2 * Lin * C;
    upsample2_kernel<<<grid1d(total), 256, 0, (cudaStream_t)stream>>>(x, ldx, y, ldy, Lin, C, total);
    return check_launch("upsample2");
}

extern "C" int m2d_upsample2_bwd(const float* dy, int lddy, float* dx, int lddx, int nb, int Lin, int C,
                                 int accumulate, void* stream) {
    M2D_REQUIRE(dy && dx && nb > 0 && Lin > 0 && C > 0, "upsample2_bwd: bad args");
    long long total = (long long)nb * Lin * C;
    upsample2_bwd_kernel<<<grid1d(total), 256, 0, (cudaStream_t)stream>>>(dy, lddy, dx, lddx, Lin, C, total,
                                                                        accumulate);
    return check_launch("upsample2_bwd");
}

extern "C" int m2d_copy2d(const float* x, int ldx, float* y, int ldy, long long M, int C, int accumulate,
                          void* stream) {
    M2D_REQUIRE(x && y && M > 0 && C > 0, "copy2d: bad args");
    copy2d_kernel<<<grid1d(M * C), 256, 0, (cudaStream_t)stream>>>(x, ldx, y, ldy, M, C, accumulate);
    return check_launch("copy2d");
}

extern "C" int m2d_copy2d_batch(const m2d_copy2d_desc* descs, int n, void* stream) {
    M2D_REQUIRE(descs && n > 0 && n <= M2D_COPY2D_MAX, "copy2d_batch: 1..M2D_COPY2D_MAX entries");
    Copy2dBatch b;
    long long most = 0;
    for (int k = 0; k < n; ++k) {
        M2D_REQUIRE(descs[k].x && descs[k].y && descs[k].M > 0 && descs[k].C > 0, "copy2d_batch: bad entry");
        b.e[k] = descs[k];
        const long long t = descs[k].M * descs[k].C;
        most = t > most ? t : most;
    }
    b.n = n;
    copy2d_batch_kernel<<<grid1d(most), 256, 0, (cudaStream_t)stream>>>(b);
    return check_launch("copy2d_batch");
}

extern "C" int m2d_fusion_mlp(const float* x, int ldx, int n, int F, int H, const float* w1, const float* b1,
                              const float* w2, const float* b2, const float* dd, float* u, float* d, float* dh,
                              float* dx, int lddx, void* stream) {
    M2D_REQUIRE(x && w1 && b1 && w2 && b2 && u && d && n > 0 && F > 0 && F <= 8192 && H > 0 && H <= FUSION_MAXH,
                "fusion_mlp: bad args (H <= 256, F <= 8192)");
    M2D_REQUIRE(!dd || (dh && dx), "fusion_mlp: backward needs dh and dx");
    fusion_mlp_kernel<<<n, 256, (size_t)F * sizeof(float), (cudaStream_t)stream>>>(x, ldx, F, H, w1, b1, w2, b2, dd, u, d,
                                                                                 dh, dx, lddx);
    return check_launch("fusion_mlp");
}

extern "C" int m2d_transpose_bcl(const float* x, float* y, int nb, int R, int C, void* stream) {
    M2D_REQUIRE(x && y && nb > 0 && R > 0 && C > 0, "transpose: bad args");
    dim3 grid((unsigned)cdiv(C, 32), (unsigned)cdiv(R, 32), (unsigned)nb);
    transpose_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(x, y, R, C);
    return check_launch("transpose");
}

extern "C" int m2d_wgan_scalars(const double* sums, const float* gp, int B, long long n_l1, long long n_tv,
                                float c0, float c1, int mode, float* out, const float* d_real, const float* d_fake,
                                void* stream) {
    M2D_REQUIRE(out && B > 0 && (mode == 1 || gp) && (sums || (d_real && mode == 0)) && (!d_real == !d_fake),
                "wgan_scalars: bad args");
    wgan_scalars_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(sums, gp, B, n_l1, n_tv, c0, c1, mode, out, d_real, d_fake);
    return check_launch("wgan_scalars");
}

extern "C" int m2d_slice_audio(const float* audio, float* out, int nseq, int A, int nwin, int W, int stride,
                               int pad_left, void* stream) {
    M2D_REQUIRE(audio && out && nseq > 0 && A > 0 && nwin > 0 && W > 0 && stride >= 0, "slice_audio: bad args");
    M2D_REQUIRE(nwin <= 65535 && nseq <= 65535, "slice_audio: too many windows/sequences");
    dim3 grid((unsigned)(cdiv(W, 256) < 8 ? cdiv(W, 256) : 8), (unsigned)nwin, (unsigned)nseq);
    slice_audio_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(audio, out, A, nwin, W, stride, pad_left);
    return check_launch("slice_audio");
}

extern "C" int m2d_crop_batch(const float* poses, const long long* pose_off, const float* music,
                              const long long* music_off, const int* seq, const int* start, int B, int T, int O,
                              int ratio, int A, float* real, float* audio, void* stream) {
    M2D_REQUIRE(poses && pose_off && music && music_off && seq && start && real && audio, "crop_batch: null pointer");
    M2D_REQUIRE(B > 0 && B <= 65535 && T > 0 && O > 0 && ratio > 0 && A > 0, "crop_batch: bad dims");
    const int per = T * O + A;
    dim3 grid((unsigned)((per + 1023) / 1024 < 64 ? (per + 1023) / 1024 : 64), (unsigned)B);
    crop_batch_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(poses, pose_off, music, music_off, seq, start, T * O, O,
                                                             ratio, A, real, audio);
    return check_launch("crop_batch");
}

extern "C" int m2d_adam(float* p, const float* g, float* m, float* v, long long n, int* step, float lr,
                        float beta1, float beta2, float eps, float gscale, void* stream) {
    M2D_REQUIRE(p && g && m && v && step && n > 0, "adam: bad args");
    cudaStream_t st = (cudaStream_t)stream;
    tick_kernel<<<1, 1, 0, st>>>(step);
    adam_kernel<<<grid1d(n), 256, 0, st>>>(p, g, m, v, n, step, lr, beta1, beta2, eps, gscale);
    return check_launch("adam");
}

extern "C" int m2d_timestamp(unsigned long long* slot, void* stream) {
    M2D_REQUIRE(slot, "timestamp: bad args");
    timestamp_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(slot);
    return check_launch("timestamp");
}
