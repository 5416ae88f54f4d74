// Adam fused with the weight re-layouts (m2d_adam_pack, include/m2d.h): one pass over the parameters per optimiser
// step.  Replaces the chain  pack_batch(UNPACK_GRAD) -> adam_kernel -> pack_batch(FWD / BWD / ...)  of round 1.
// HBM-bound: per parameter 4 B gradient + 24 B p/m/v + 8-24 B per packed copy; every array is accessed in its own
// contiguous order through a shared-memory tile.
#include "common.cuh"

namespace m2d {

// destination of weight element (co, ci, t) in the layouts of m2d_pack_desc (same formulas as pack_batch_kernel)
struct PackGeom {
    int kind, Cout, Cin, k, stride, pad;
    int R, Tm, cmax;                 // tiled rows; merged backward: unified taps, cmax
};
__device__ __forceinline__ PackGeom pack_geom(const m2d_adam_pack_out& o, int Cout, int Cin, int k) {
    PackGeom g;
    g.kind = o.kind; g.Cout = Cout; g.Cin = Cin; g.k = k; g.stride = o.stride; g.pad = o.reserved;
    g.Tm = 0; g.cmax = 0;
    if (o.kind == M2D_PACK_FWD) g.R = tiled_rows(Cout);
    else if (o.kind == M2D_PACK_FULL_BWD) g.R = tiled_rows(k * Cin);
    else if (o.kind == M2D_PACK_BWD) g.R = tiled_rows(Cin);
    else {
        g.R = tiled_rows(o.stride * Cin);
        g.cmax = (o.stride - 1 + g.pad) / o.stride;
        for (int rr = 0; rr < o.stride; ++rr) {
            const int rh = (rr + g.pad) % o.stride, cc = (rr + g.pad) / o.stride;
            const int Tr = (k - rh + o.stride - 1) / o.stride + g.cmax - cc;
            g.Tm = Tr > g.Tm ? Tr : g.Tm;
        }
    }
    return g;
}
__device__ __forceinline__ void pack_index(const PackGeom& g, int co, int ci, int t, long long& plain, long long& pidx) {
    if (g.kind == M2D_PACK_FWD) {
        plain = ((long long)co * g.k + t) * g.Cin + ci;
        pidx = g.Cin == 1 ? tiled_index(co, 0, t, 1, g.k, g.R) : tiled_index(co, t, ci, g.k, g.Cin, g.R);
    } else if (g.kind == M2D_PACK_FULL_BWD) {
        const long long r = (long long)t * g.Cin + ci;
        plain = r * g.Cout + co;
        pidx = tiled_index((int)r, 0, co, 1, g.Cout, g.R);
    } else if (g.kind == M2D_PACK_BWD) {
        const int rho = t % g.stride, q = t / g.stride;
        long long off = 0, poff = 0;
        int Trho = 0;
        for (int r = 0; r <= rho; ++r) {
            Trho = (g.k - r + g.stride - 1) / g.stride;
            if (Trho < 0) Trho = 0;
            if (r < rho) {
                off += (long long)g.Cin * g.Cout * Trho;
                poff += tiled_blocks(g.Cin, Trho, g.Cout) * tiled_block_floats(g.R);
            }
        }
        plain = off + ((long long)ci * Trho + q) * g.Cout + co;
        pidx = poff + tiled_index(ci, q, co, Trho, g.Cout, g.R);
    } else {                                     // M2D_PACK_BWD_MERGED
        const int rho = t % g.stride, q = t / g.stride;
        int r0 = (rho - g.pad) % g.stride;
        if (r0 < 0) r0 += g.stride;
        const int c0 = (r0 + g.pad) / g.stride;
        const int qp = q + g.cmax - c0;
        const long long row = (long long)r0 * g.Cin + ci;
        plain = row * g.Tm * g.Cout + (long long)qp * g.Cout + co;
        pidx = tiled_index((int)row, qp, co, g.Tm, g.Cout, g.R);
    }
}

struct AdamConst {
    float step_size, bc2_sqrt, b1, b2, eps, gscale;
};
__device__ __forceinline__ float adam_update(float p, float g, float& m, float& v, const AdamConst& c) {
    const float gi = g * c.gscale;
    const float mi = m + (gi - m) * (1.f - c.b1);
    const float vi = v * c.b2 + (1.f - c.b2) * gi * gi;
    m = mi;
    v = vi;
    const float denom = sqrtf(vi) / c.bc2_sqrt + c.eps;
    return p - c.step_size * (mi / denom);
}

__global__ void __launch_bounds__(256)
adam_pack_kernel(const m2d_adam_item* __restrict__ items, int* counters, float lr, float b1, float b2, float eps,
                 float gscale, const bool mixed) {
    extern __shared__ float tile[];
    const m2d_adam_item& it = items[blockIdx.x];
    const int t_step = counters[0] + 1;
    AdamConst c;
    {
        const double bc1 = 1.0 - pow((double)b1, (double)t_step);
        const double bc2 = 1.0 - pow((double)b2, (double)t_step);
        c.step_size = (float)((double)lr / bc1);
        c.bc2_sqrt = (float)sqrt(bc2);
        c.b1 = b1; c.b2 = b2; c.eps = eps; c.gscale = gscale;
    }
    const int tid = threadIdx.x;
    if (it.flat_n > 0) {
        float* __restrict__ p = it.p;
        float* __restrict__ m = it.m;
        float* __restrict__ v = it.v;
        const float* __restrict__ g = it.g;
        for (long long i = tid; i < it.flat_n; i += 256) {
            float mi = m[i], vi = v[i];
            p[i] = adam_update(p[i], g[i], mi, vi, c);
            m[i] = mi;
            v[i] = vi;
        }
    } else {
        const int Cout = it.Cout, Cin = it.Cin, k = it.k;
        const int nco = it.nco, nci = it.nci, nt = it.nt, co0 = it.co0, ci0 = it.ci0, t0 = it.t0;
        const int ntp = nt | 1;                          // odd tap pitch: conflict-free transposes
        const int Lp = (nci * ntp) | 1;                  // odd row pitch
        const int L = nci * nt;
        const int total = nco * L;
        // 1. gradient -> tile[co][ci][t]
        if (it.g_packed) {
            const float* __restrict__ g = it.g;
            for (int e = tid; e < total; e += 256) {     // (co, t, ci), ci fastest: the tap-major array's order
                const int cl = e / L, r = e - cl * L, tl = r / nci, il = r - tl * nci;
                tile[cl * Lp + il * ntp + tl] = g[((long long)(co0 + cl) * k + t0 + tl) * Cin + ci0 + il];
            }
            __syncthreads();
        }
        // 2. Adam in the parameter layout's order (co, ci, t), t fastest; the new weight replaces the gradient in the tile
        {
            float* __restrict__ p = it.p;
            float* __restrict__ m = it.m;
            float* __restrict__ v = it.v;
            const float* __restrict__ g = it.g;
            for (int e = tid; e < total; e += 256) {
                const int cl = e / L, r = e - cl * L, il = r / nt, tl = r - il * nt;
                const long long gi = ((long long)(co0 + cl) * Cin + ci0 + il) * k + t0 + tl;
                const int si = cl * Lp + il * ntp + tl;
                const float gv = it.g_packed ? tile[si] : g[gi];
                float mi = m[gi], vi = v[gi];
                const float pn = adam_update(p[gi], gv, mi, vi, c);
                p[gi] = pn;
                m[gi] = mi;
                v[gi] = vi;
                tile[si] = pn;
            }
        }
        __syncthreads();
        // 3. re-layouts, each in its destination's contiguous order
        for (int o = 0; o < it.n_pack; ++o) {
            const m2d_adam_pack_out& pk = it.pk[o];
            const PackGeom gm = pack_geom(pk, Cout, Cin, k);
            if (pk.kind == M2D_PACK_FWD) {               // rows of co: (co, t, ci), ci fastest
                for (int e = tid; e < total; e += 256) {
                    const int cl = e / L, r = e - cl * L, tl = r / nci, il = r - tl * nci;
                    const float val = tile[cl * Lp + il * ntp + tl];
                    long long plain, pidx;
                    pack_index(gm, co0 + cl, ci0 + il, t0 + tl, plain, pidx);
                    if (pk.dst) pk.dst[plain] = val;
                    if (pk.dst_tiled) store_tiled_split(pk.dst_tiled, pidx, gm.R, val, mixed);
                }
            } else {                                     // co is the contiguous index: (ci, t, co), co fastest
                for (int e = tid; e < total; e += 256) {
                    const int r = e / nco, cl = e - r * nco, il = r / nt, tl = r - il * nt;
                    const float val = tile[cl * Lp + il * ntp + tl];
                    long long plain, pidx;
                    pack_index(gm, co0 + cl, ci0 + il, t0 + tl, plain, pidx);
                    if (pk.dst) pk.dst[plain] = val;
                    if (pk.dst_tiled) store_tiled_split(pk.dst_tiled, pidx, gm.R, val, mixed);
                }
            }
        }
    }
    // the last block to finish publishes the new step count (every block has read the old one by then)
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(&counters[1], 1) == (int)gridDim.x - 1) {
            counters[0] = t_step;
            counters[1] = 0;
        }
    }
}

int gemm_mode();
}  // namespace m2d

using namespace m2d;

extern "C" int m2d_adam_pack(const m2d_adam_item* items, int n, int smem_floats, int* counters, float lr, float beta1,
                             float beta2, float eps, float gscale, void* stream) {
    M2D_REQUIRE(items && counters && n > 0 && smem_floats >= 0, "adam_pack: bad args");
    const size_t smem = (size_t)smem_floats * sizeof(float);
    M2D_REQUIRE(smem <= 200 * 1024, "adam_pack: tile too large for shared memory");
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(adam_pack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("adam_pack: %s", cudaGetErrorString(e));
            return M2D_ERR_CUDA;
        }
        configured = smem;
    }
    adam_pack_kernel<<<n, 256, smem, (cudaStream_t)stream>>>(items, counters, lr, beta1, beta2, eps, gscale,
                                                            gemm_mode() == M2D_GEMM_TF32_BF16);
    return check_launch("adam_pack");
}
