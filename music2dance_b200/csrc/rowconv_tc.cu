// Row-convolution GEMM on the 5th-generation tensor cores (tcgen05, sm_100a).
//
//   y[m, n] = epi( sum_k A[m, k] * W[n, k] ),   m = (batch, row), k = (tap, channel)
//
// CTA tile 128 x BN (BN <= 128, multiple of 16), K consumed in blocks of 32 floats
// (= one 128-byte swizzle row).  Roles (288 threads):
//   warps 0-7  gather the A rows (strided / windowed / zero padded, exactly the
//              index arithmetic of the fp32 SIMT kernel) and the weight rows from
//              global memory, round them to TF32 with round-to-nearest
//              (cvt.rna.tf32.f32) and store them into shared memory in the
//              canonical K-major SWIZZLE_128B layout; afterwards they run the epilogue
//   warp  8    allocates TMEM and issues tcgen05.mma (kind::tf32, M = 128, N = BN,
//              K = 8 per instruction) from shared-memory descriptors; accumulators
//              live in TMEM (128 lanes x BN fp32 columns)
// Stages are handed over with mbarriers: full[s] (256 producer arrivals after
// fence.proxy.async) and empty[s] (tcgen05.commit).  Precision modes:
//   NS = 1 : one TF32 product per term (10-bit mantissas, fp32 accumulate)
//   NS = 3 : 3xTF32 split  a = a_hi + a_lo :  a_hi*b_hi + a_lo*b_hi + a_hi*b_lo
//            (error ~2^-21 per product: fp32-grade results on the tensor pipe)
// Epilogue: tcgen05.ld (32x32b.x16) -> shared memory -> coalesced global stores with
// the fused bias / activation / mask / residual epilogue of m2d_rowconv, or split-K
// partials into the workspace.
#include <cuda.h>
#include <cstdlib>
#include <map>
#include <tuple>
#include "common.cuh"

#include "tc_ptx.cuh"

namespace m2d {

struct TcRow {          // per output row of the CTA tile
    long long base;     // element offset of the row's batch (or sequence) in x; -1: row beyond M
    int r0;             // i*sr + roff0
    int ab;             // windowed mode: f*win_stride - win_pad
    int b, i;           // batch / row indices for the epilogue
};

// BTMA: the weight operand comes from the pre-split, pre-tiled copy (a.w_tiled: w_hi = rna_tf32(w), w_lo =
// rna_tf32(w - w_hi) in the tensor core's shared-memory image, written by the re-layout kernel after each
// optimizer step): one contiguous bulk copy per stage (warp TC_PW + 1); the producer warps then only stage
// the activation operand.
template <int NS, bool VEC, bool C1, bool BTMA>
__global__ void __launch_bounds__(TC_THREADS, 1)
rowconv_tc_kernel(const m2d_rowconv_args a, const int M, const int nsteps, const int cchunks) {
    constexpr int STAGES = tc_stages(NS);
    constexpr int STAGE_BYTES = tc_stage_bytes(NS);
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bars[2 * 4 + 1];
    __shared__ uint32_t tmem_slot;
    __shared__ TcRow rows[TC_BM];

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // SWIZZLE_128B atoms need 1 KiB alignment
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar_full = smem_u32(&bars[0]);          // + 8*s
    const uint32_t bar_empty = smem_u32(&bars[4]);         // + 8*s
    const uint32_t bar_acc = smem_u32(&bars[8]);

    const int m0 = blockIdx.x * TC_BM, n0 = blockIdx.y * TC_BNMAX;
    int bn = a.N - n0;
    bn = bn > TC_BNMAX ? TC_BNMAX : ((bn + 15) & ~15);
    const int tm_cols = bn <= 32 ? 32 : (bn <= 64 ? 64 : 128);
    const int per = (nsteps + gridDim.z - 1) / gridDim.z;
    const int s_begin = blockIdx.z * per;
    const int s_end = min(nsteps, s_begin + per);
    const int nk = max(0, s_end - s_begin);
    const bool win = a.win_T > 0;

    if (tid < TC_BM) {
        int m = m0 + tid;
        TcRow r;
        if (m < M) {
            r.b = m / a.y_rows;
            r.i = m - r.b * a.y_rows;
            r.r0 = r.i * a.sr + a.roff0;
            if (win) {
                int seq = r.b / a.win_T;
                int f = r.b - seq * a.win_T;
                r.ab = f * a.win_stride - a.win_pad;
                r.base = (long long)seq * a.win_seq_len;
            } else {
                r.ab = 0;
                r.base = (long long)r.b * a.x_bs;
            }
        } else {
            r.b = r.i = r.r0 = r.ab = 0;
            r.base = -1;
        }
        rows[tid] = r;
    }
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar_full + 8 * s, TC_PRODUCERS + (BTMA ? 1 : 0));
            mbar_init(bar_empty + 8 * s, 1);
        }
        mbar_init(bar_acc, 1);
        fence_barrier_init();
    }
    if (warp == TC_PW) tmem_alloc(smem_u32(&tmem_slot), (uint32_t)tm_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    // programmatic dependent launch: everything above overlapped the tail of the previous kernel on the
    // stream; its results (activations, re-packed weights) are visible after this point
    asm volatile("griddepcontrol.wait;" ::: "memory");

    if (warp < TC_PW) {
        // ------------------------------------------------------------------ producers
        const int j = tid & 7;                 // 16-byte chunk of the 128-byte K row
        const int rsub = tid >> 3;             // 0..63
        constexpr int RS = TC_PRODUCERS / 8;    // row stride between the chunks of one thread
        constexpr int NB = BTMA ? 1 : TC_RPT;
        constexpr int D = BTMA ? 4 : 2;         // k-blocks of A (and B) held in registers ahead of the stores
        float4 ra[D][TC_RPT], rb[D][NB];
        // the four A rows of this thread stay in registers (no table lookups in the K loop)
        const float* arow[TC_RPT];
        int ar0[TC_RPT];
#pragma unroll
        for (int q = 0; q < TC_RPT; ++q) {
            const TcRow r = rows[rsub + RS * q];
            ar0[q] = r.base >= 0 ? r.r0 : (1 << 29);          // beyond every x_rows: row reads as zero
            arow[q] = a.x + (r.base >= 0 ? r.base : 0) + (long long)r.r0 * a.x_ld;
        }

        auto load = [&](int s, float4* pa, float4* pb) {
            int t = 0, c = 0, k0 = 0;
            if (C1) {
                k0 = s * TC_BK + 4 * j;
            } else {
                t = s / cchunks;
                c = (s - t * cchunks) * TC_BK + 4 * j;
            }
#pragma unroll
            for (int q = 0; q < TC_RPT; ++q) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (C1) {
                    const TcRow r = rows[rsub + RS * q];
                    if (r.base >= 0) {
                        float e[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            int k = k0 + u;
                            int rr = r.r0 + k * a.droff;
                            bool ok = k < a.T && rr >= 0 && rr < a.x_rows;
                            int pos = r.ab + rr;
                            if (win) ok = ok && pos >= 0 && pos < a.win_seq_len;
                            e[u] = ok ? __ldg(a.x + r.base + (long long)pos * a.x_ld) : 0.f;
                        }
                        v = make_float4(e[0], e[1], e[2], e[3]);
                    }
                } else {
                    const int dr = t * a.droff;
                    if ((unsigned)(ar0[q] + dr) < (unsigned)a.x_rows) {
                        const float* p = arow[q] + (long long)dr * a.x_ld + c;
                        if (VEC) {
                            if (c < a.Cc) v = __ldg(reinterpret_cast<const float4*>(p));
                        } else {
                            float e[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) e[u] = (c + u < a.Cc) ? __ldg(p + u) : 0.f;
                            v = make_float4(e[0], e[1], e[2], e[3]);
                        }
                    }
                }
                pa[q] = v;
            }
            if (!BTMA) {
#pragma unroll
                for (int q = 0; q < TC_RPT; ++q) {
                    const int nl = rsub + RS * q;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (nl < bn && n0 + nl < a.N) {
                        const float* wrow = a.w + (long long)(n0 + nl) * a.w_ld;
                        if (C1) {
                            float e[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) e[u] = (k0 + u < a.T) ? __ldg(wrow + k0 + u) : 0.f;
                            v = make_float4(e[0], e[1], e[2], e[3]);
                        } else {
                            const float* p = wrow + (long long)t * a.Cc + c;
                            if (VEC) {
                                if (c < a.Cc) v = __ldg(reinterpret_cast<const float4*>(p));
                            } else {
                                float e[4];
#pragma unroll
                                for (int u = 0; u < 4; ++u) e[u] = (c + u < a.Cc) ? __ldg(p + u) : 0.f;
                                v = make_float4(e[0], e[1], e[2], e[3]);
                            }
                        }
                    }
                    pb[q] = v;
                }
            }
        };

        auto split_store = [&](uint8_t* hi_tile, uint8_t* lo_tile, int r, float4 v, bool weight) {
            const uint32_t off = sw128_off(r, j);
            float4 h = make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
            *reinterpret_cast<float4*>(hi_tile + off) = h;
            if (NS == 3) {
                float4 l = make_float4(to_tf32(v.x - h.x), to_tf32(v.y - h.y), to_tf32(v.z - h.z), to_tf32(v.w - h.w));
                *reinterpret_cast<float4*>(lo_tile + off) = l;
            }
            if (NS == 2) {      // BF16 plane: [bf16(hi) | bf16(lo)] for the activations, [bf16(lo) | bf16(hi)] for the weights
                const float4 l = make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w);
                store_bf16_pair(lo_tile, r, j, weight ? l : h, weight ? h : l);
            }
        };

        auto store = [&](int it, const float4* pa, const float4* pb) {
            const int st = it % STAGES;
            const uint32_t ph = (uint32_t)((it / STAGES) & 1);
            mbar_wait(bar_empty + 8 * st, ph ^ 1);
            uint8_t* sA = smem + (size_t)st * STAGE_BYTES;
            uint8_t* sB = sA + tc_planes(NS) * TC_A_BYTES;
#pragma unroll
            for (int q = 0; q < TC_RPT; ++q) split_store(sA, sA + TC_A_BYTES, rsub + RS * q, pa[q], false);
            if (!BTMA) {
#pragma unroll
                for (int q = 0; q < TC_RPT; ++q)
                    if (rsub + RS * q < bn) split_store(sB, sB + TC_B_BYTES, rsub + RS * q, pb[q], true);
            }
            fence_proxy_async_smem();          // generic-proxy stores -> visible to the tensor core (async proxy)
            mbar_arrive(bar_full + 8 * st);
        };

#pragma unroll
        for (int p = 0; p < D - 1; ++p)
            if (p < nk) load(s_begin + p, ra[p], rb[p]);
        for (int it = 0; it < nk; it += D) {
#pragma unroll
            for (int u = 0; u < D; ++u) {
                if (it + u < nk) {
                    if (it + u + D - 1 < nk) load(s_begin + it + u + D - 1, ra[(u + D - 1) % D], rb[(u + D - 1) % D]);
                    store(it + u, ra[u], rb[u]);
                }
            }
        }
    } else if (warp == TC_PW) {
        // ------------------------------------------------------------------ MMA issuer (warp-uniform loop, one elected lane issues)
        {
            const uint32_t idesc = tf32_idesc(TC_BM, bn);
            const uint32_t idesc16 = bf16_idesc(TC_BM, bn);
            // the lo plane follows the hi plane: 128 rows apart when staged by threads, R rows in a tiled block
            const uint64_t b_plane = (uint64_t)((BTMA ? (uint32_t)tiled_rows(a.N) * 128u : (uint32_t)TC_B_BYTES) >> 4);
            const uint64_t dA0 = sw128_desc(smem_base);
            const uint64_t offB = (uint64_t)((tc_planes(NS) * TC_A_BYTES) >> 4);
            uint64_t dA = dA0;
            int st = 0;
            uint32_t ph = 0;
            for (int it = 0; it < nk; ++it) {
                mbar_wait(bar_full + 8 * st, ph);
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < TC_BK / 8; ++k) {
                        const uint64_t ah = dA + 2 * k, bh = ah + offB;
                        if (NS == 3) {
                            umma_tf32(tmem, ah + (TC_A_BYTES >> 4), bh, idesc, (it | k) != 0);
                            umma_tf32(tmem, ah, bh + b_plane, idesc, 1);
                            umma_tf32(tmem, ah, bh, idesc, 1);
                        } else {
                            umma_tf32(tmem, ah, bh, idesc, (it | k) != 0);
                        }
                    }
                    if (NS == 2) {      // both cross terms as one BF16 contraction over the 64-element rows of the BF16 planes
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16(tmem, dA + (TC_A_BYTES >> 4) + 2 * k, dA + offB + b_plane + 2 * k, idesc16, 1);
                    }
                    umma_commit(bar_empty + 8 * st);     // stage reusable once these MMAs have read it
                }
                __syncwarp();
                dA += (uint64_t)(STAGE_BYTES >> 4);
                if (++st == STAGES) { st = 0; ph ^= 1; dA = dA0; }
            }
            if (elect_one()) umma_commit(bar_acc);   // accumulator complete
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ weight blocks: one bulk copy per stage
        if (BTMA) {
            const int R = tiled_rows(a.N);
            const uint32_t bytes = (uint32_t)tc_planes(NS) * (uint32_t)R * 128u;   // hi [+ lo / BF16] plane of the block
            // block ((nt*T + t)*cchunks + c); Cc == 1: single tap whose channels are the taps (T = 1, c = K step)
            const long long nt_base = (long long)blockIdx.y * (C1 ? 1 : a.T) * (C1 ? nsteps : cchunks);
            for (int it = 0; it < nk; ++it) {
                const int st = it % STAGES;
                const uint32_t ph = (uint32_t)((it / STAGES) & 1);
                const float* src = a.w_tiled + (nt_base + s_begin + it) * tiled_block_floats(R);
                mbar_wait(bar_empty + 8 * st, ph ^ 1);
                const uint32_t sB = smem_base + (uint32_t)st * STAGE_BYTES + tc_planes(NS) * TC_A_BYTES;
                if (elect_one()) {
                    mbar_arrive_expect_tx(bar_full + 8 * st, bytes);
                    bulk_load(sB, src, bytes, bar_full + 8 * st);
                }
                __syncwarp();
            }
        }
        __syncwarp();
    }

    // ---------------------------------------------------------------------- epilogue
    float* Cs = reinterpret_cast<float*>(smem);      // [128][TC_CLD], reuses the (drained) stages
    if (warp < TC_PW) {
        if (nk > 0) {
            mbar_wait(bar_acc, 0);
            tc_fence_after();
            tmem_to_smem(tmem, Cs, warp, lane, bn);
        } else {
            for (int idx = tid; idx < TC_BM * TC_CLD; idx += TC_PRODUCERS) Cs[idx] = 0.f;
        }
        tc_fence_before();
    }
    __syncthreads();
    const int Z = (int)gridDim.z;                    // == cluster size along z
    if (warp >= TC_PW) {
        if (warp == TC_PW) {
            tc_fence_after();
            tmem_dealloc(tmem, (uint32_t)tm_cols);
        }
        if (Z > 1) {
            cluster_sync_all();                       // partial tiles ready
            cluster_sync_all();                       // peers done reading this CTA's tile
        }
        return;
    }
    const int ncols = min(bn, a.N - n0);
    // thread -> (row group, 4-column chunk): P lanes per row, P = pow2 >= ncols/4
    const int cpr = (ncols + 3) >> 2;
    int P = 1;
    while (P < cpr) P <<= 1;
    const int rpi = TC_PRODUCERS / P;                 // rows per iteration over the CTA
    const int c4 = (tid & (P - 1)) * 4, rsub0 = tid / P;
    const bool cvalid = c4 < ncols;
    const bool vecN = (a.N & 3) == 0 && ncols - c4 >= 4;
    int row_lo = 0, row_hi = TC_BM;
    if (Z > 1) {
        cluster_sync_all();
        const int RB = TC_BM / Z;
        row_lo = (int)cluster_rank() * RB;
        row_hi = row_lo + RB;
    }
    const bool vy = vecN && (a.y_ld & 3) == 0 && (a.y_bs & 3) == 0 && aligned16d(a.y) &&
                    (!a.y2 || aligned16d(a.y2)) &&
                    (!a.mask_mode || ((a.m_ld & 3) == 0 && (a.m_bs & 3) == 0 && aligned16d(a.mask))) &&
                    (!a.add || ((a.a_ld & 3) == 0 && (a.a_bs & 3) == 0 && aligned16d(a.add)));
    for (int rl = row_lo + rsub0; rl < row_hi; rl += rpi) {
        if (!cvalid) continue;
        const TcRow r = rows[rl];
        if (r.base < 0) continue;
        const int n = n0 + c4;
        float v[4];
        if (Z > 1) {
            cluster_reduce4(Cs, rl, c4, Z, v);
        } else {
            const float4 t = *reinterpret_cast<const float4*>(Cs + rl * TC_CLD + c4);
            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        }
        const long long yo = r.b * a.y_bs + (long long)r.i * a.y_ld + n;
        const long long mo = a.mask_mode ? r.b * a.m_bs + (long long)r.i * a.m_ld + n : 0;
        const long long ao = a.add ? r.b * a.a_bs + (long long)r.i * a.a_ld + n : 0;
        if (vy) {
            float ad[4] = {0.f, 0.f, 0.f, 0.f}, mk[4] = {1.f, 1.f, 1.f, 1.f};
            if (a.add) {
                float4 t = *reinterpret_cast<const float4*>(a.add + ao);
                ad[0] = t.x; ad[1] = t.y; ad[2] = t.z; ad[3] = t.w;
            }
            if (a.mask_mode) {
                float4 t = *reinterpret_cast<const float4*>(a.mask + mo);
                mk[0] = act_deriv(t.x, a.mask_mode); mk[1] = act_deriv(t.y, a.mask_mode);
                mk[2] = act_deriv(t.z, a.mask_mode); mk[3] = act_deriv(t.w, a.mask_mode);
            }
            float w2[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float x = v[u];
                if (a.bias) x += __ldg(a.bias + n + u);
                x = apply_act(x, a.act);
                if (a.add && a.add_before_mask) x += ad[u];
                w2[u] = x;
                x *= mk[u];
                if (a.add && !a.add_before_mask) x += ad[u];
                v[u] = x;
            }
            if (a.y2) *reinterpret_cast<float4*>(a.y2 + yo) = make_float4(w2[0], w2[1], w2[2], w2[3]);
            *reinterpret_cast<float4*>(a.y + yo) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
            for (int u = 0; u < 4 && c4 + u < ncols; ++u) {
                float x = v[u];
                if (a.bias) x += __ldg(a.bias + n + u);
                x = apply_act(x, a.act);
                if (a.add && a.add_before_mask) x += a.add[ao + u];
                if (a.y2) a.y2[yo + u] = x;
                if (a.mask_mode) x *= act_deriv(a.mask[mo + u], a.mask_mode);
                if (a.add && !a.add_before_mask) x += a.add[ao + u];
                a.y[yo + u] = x;
            }
        }
    }
    if (Z > 1) cluster_sync_all();                   // keep this CTA's tile alive until every peer has read it
}

// Programmatic dependent launch of the tensor-core kernels (prologue overlaps the previous kernel's
// tail; the kernels call griddepcontrol.wait before touching dependent data).  M2D_PDL=0 disables it.
static bool pdl_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("M2D_PDL");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

// Launch with the split-K CTAs of a tile grouped into a (1,1,Z) thread-block cluster.
template <typename K, typename... Args>
static int launch_clustered(const char* what, K kern, dim3 grid, int smem, int Z, cudaStream_t st, int threads,
                            Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3((unsigned)threads);
    cfg.dynamicSmemBytes = (size_t)smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (pdl_enabled()) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    if (Z > 1) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = 1;
        attr[na].val.clusterDim.y = 1;
        attr[na].val.clusterDim.z = (unsigned)Z;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, args...);
    if (e != cudaSuccess) {
        set_error("%s: launch (grid %u,%u,%u cluster z=%d): %s", what, grid.x, grid.y, grid.z, Z, cudaGetErrorString(e));
        return M2D_ERR_CUDA;
    }
    return M2D_OK;
}

// largest power-of-two split <= want that the device can co-schedule as one cluster of this kernel
template <typename K>
static int max_cluster_z(K kern, int smem, int want, bool allow16, int threads = TC_THREADS) {
    int z = 1;
    while (z * 2 <= want && z * 2 <= (allow16 ? 16 : 8)) z *= 2;
    for (; z > 1; z /= 2) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(1, 1, (unsigned)z);
        cfg.blockDim = dim3((unsigned)threads);
        cfg.dynamicSmemBytes = (size_t)smem;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 1;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = (unsigned)z;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) == cudaSuccess && n > 0) break;
        (void)cudaGetLastError();
    }
    return z;
}

// ---- TMA descriptor encoder (activation halo tiles, rowconv_halo.cuh) -------------
typedef CUresult (*tmap_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static tmap_encode_fn tmap_encoder() {
    static tmap_encode_fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<tmap_encode_fn>(p);
        (void)cudaGetLastError();
    }
    return fn;
}
template <int NS, bool VEC, bool C1, bool BTMA>
static int launch_tc(const m2d_rowconv_args& a, int M, int nsteps, int cchunks, int want_splits, cudaStream_t st) {
    auto kern = rowconv_tc_kernel<NS, VEC, C1, BTMA>;
    static bool configured = false;
    static int zmax = 1;
    const int smem = tc_smem_bytes(NS);
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) {
            set_error("rowconv_tc: smem attribute (%d B): %s", smem, cudaGetErrorString(e));
            return M2D_ERR_CUDA;
        }
        zmax = max_cluster_z(kern, smem, 8, false);
        configured = true;
    }
    int Z = 1;
    while (Z * 2 <= want_splits && Z * 2 <= zmax) Z *= 2;
    dim3 grid((unsigned)cdiv(M, TC_BM), (unsigned)cdiv(a.N, TC_BNMAX), (unsigned)Z);
    return launch_clustered("rowconv_tc", kern, grid, smem, Z, st, TC_THREADS, a, M, nsteps, cchunks);
}

template <int NS, bool BTMA>
static int launch_tc_shape(const m2d_rowconv_args& a, int M, int nsteps, int cchunks, int splits, cudaStream_t st,
                           bool c1, bool vec) {
    if (c1) return launch_tc<NS, false, true, BTMA>(a, M, nsteps, cchunks, splits, st);
    if (vec) return launch_tc<NS, true, false, BTMA>(a, M, nsteps, cchunks, splits, st);
    return launch_tc<NS, false, false, BTMA>(a, M, nsteps, cchunks, splits, st);
}

#include "rowconv_halo.cuh"

long long halo_launch_count() { return g_halo_launches; }
long long halo_persist_launch_count() { return g_halo_persist_launches; }

// Called by m2d_rowconv when the tensor-core path is selected.  Returns 1 if the shape is
// not worth a tensor-core launch (caller falls through to the SIMT kernel), <= 0 otherwise.
int rowconv_tc_dispatch(const m2d_rowconv_args& a, int M, int mode, cudaStream_t st) {
    const bool c1 = a.Cc == 1;
    const long long K = (long long)a.T * a.Cc;
    if (a.N < 4 || (long long)M * a.N * K < (1ll << 18)) return 1;
    // a handful of rows against a long contraction (audio_d.l6, stick_d.fconv and their backward-data
    // passes at small batch) is weight-streaming work: one 128-row tile cannot spread over more than a
    // cluster of CTAs, the FP32 kernel splits K over the whole chip
    if (M <= 64 && K >= 4096 && a.ws) return 1;
    // both operands through TMA, activation halo tile shared by the taps of a stride residue
    if (!c1) {
        const int rc = rowconv_halo_dispatch(a, M, mode, st);
        if (rc <= 0) return rc;
    }
    const bool vec = !c1 && a.Cc % 4 == 0 && a.x_ld % 4 == 0 && a.x_bs % 4 == 0 && a.w_ld % 4 == 0 &&
                     aligned16(a.x) && aligned16(a.w);
    const int cchunks = c1 ? 1 : (int)cdiv(a.Cc, TC_BK);
    const int nsteps = c1 ? (int)cdiv(a.T, TC_BK) : a.T * cchunks;
    const long long tiles = cdiv(M, TC_BM) * cdiv(a.N, TC_BNMAX);
    int splits = 1;
    static const int split_target = getenv("M2D_SPLIT_CTAS") ? atoi(getenv("M2D_SPLIT_CTAS")) : kNumSMs;
    if (tiles < split_target && nsteps >= 8) {          // few tiles, long K: split K over a cluster
        long long want = cdiv(split_target, tiles);
        splits = (int)(want < nsteps / 4 ? want : nsteps / 4);
        if (splits < 1) splits = 1;
    }
    // weight operand by bulk copy when the caller supplies the pre-split, pre-tiled copy
    if (a.w_tiled && aligned16(a.w_tiled)) {
        if (mode == M2D_GEMM_TF32_BF16) return launch_tc_shape<2, true>(a, M, nsteps, cchunks, splits, st, c1, vec);
        return mode == 3 ? launch_tc_shape<3, true>(a, M, nsteps, cchunks, splits, st, c1, vec)
                         : launch_tc_shape<1, true>(a, M, nsteps, cchunks, splits, st, c1, vec);
    }
    if (mode == M2D_GEMM_TF32_BF16) return launch_tc_shape<2, false>(a, M, nsteps, cchunks, splits, st, c1, vec);
    return mode == 3 ? launch_tc_shape<3, false>(a, M, nsteps, cchunks, splits, st, c1, vec)
                     : launch_tc_shape<1, false>(a, M, nsteps, cchunks, splits, st, c1, vec);
}

// ============================================================================ weight gradient
//   dW[co, (t,c)] = sum_{k=(b,l)} dy[k, co] * x[row(k, t), c]
// as a GEMM with M = Cout, N = T*Cc, K = nb*dy_rows.  Both operands are contiguous along
// their M / N index in memory (channels-last), so they are staged MN-major.  For 32-bit
// (TF32) MN-major operands the tensor core accepts exactly one shared-memory layout,
// SWIZZLE_128B_BASE32B: atoms of 4 K-rows x 128 bytes (32 contiguous M/N floats per row),
// the 32-byte units of a row XOR-swizzled with the row index (byte-address Swizzle<2,5,2>).
// A K-block of 32 rows is 8 such K-groups (SBO = 512 B apart); 32-wide M/N blocks are
// LBO = 4096 B apart.  One tcgen05.mma (K = 8) consumes two K-groups.
// Requires Cout % 4 == 0 and Cc % 4 == 0 (16-byte chunks never straddle a tap).
__device__ __forceinline__ uint64_t sw128b32_desc_mn(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)(4096 >> 4) << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;          // SWIZZLE_128B_BASE32B
    return d;
}
// tile [32 k][128 mn]; mn4 = index of the 4-float (16-byte) chunk along M/N (0..31)
__device__ __forceinline__ uint32_t mn_off(int k, int mn4) {
    const int kk = k & 3, j = mn4 & 7;
    return (uint32_t)((mn4 >> 3) * 4096 + (k >> 2) * 512 + kk * 128 + ((((j >> 1) ^ kk) << 5) | ((j & 1) << 4)));
}

template <int NS>
__global__ void __launch_bounds__(TC_THREADS, 1)
wgrad_tc_kernel(const m2d_wgrad_args a, const int Ktot, const int Ncols) {
    constexpr int STAGES = tc_stages(NS);
    constexpr int STAGE_BYTES = tc_stage_bytes(NS);
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bars[2 * 4 + 1];
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // SWIZZLE_128B atoms need 1 KiB alignment
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar_full = smem_u32(&bars[0]);
    const uint32_t bar_empty = smem_u32(&bars[4]);
    const uint32_t bar_acc = smem_u32(&bars[8]);

    const int m0 = blockIdx.x * TC_BM, n0 = blockIdx.y * TC_BNMAX;
    int bn = Ncols - n0;
    bn = bn > TC_BNMAX ? TC_BNMAX : ((bn + 31) & ~31);           // MN-major B: whole 32-wide blocks
    const int tm_cols = bn <= 32 ? 32 : (bn <= 64 ? 64 : 128);
    const int nsteps = (Ktot + TC_BK - 1) / TC_BK;
    const int per = (nsteps + gridDim.z - 1) / gridDim.z;
    const int s_begin = blockIdx.z * per;
    const int s_end = min(nsteps, s_begin + per);
    const int nk = max(0, s_end - s_begin);

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bar_full + 8 * s, TC_PRODUCERS);
            mbar_init(bar_empty + 8 * s, 1);
        }
        mbar_init(bar_acc, 1);
        fence_barrier_init();
    }
    if (warp == TC_PW) tmem_alloc(smem_u32(&tmem_slot), (uint32_t)tm_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    // programmatic dependent launch: everything above overlapped the tail of the previous kernel on the
    // stream; its results (activations, re-packed weights) are visible after this point
    asm volatile("griddepcontrol.wait;" ::: "memory");

    if (warp < TC_PW) {
        const int j = tid & 7;                 // 16-byte chunk within a 32-wide block
        const int kl = (tid >> 3) & 31;        // K row of the block handled by this thread (0..31)
        const int qb = (tid >> 8) * TC_RPT;    // first of the TC_RPT 32-wide M / N blocks of this thread
        // the four (tap, channel) column chunks of the B operand are fixed over the K loop
        int bt[TC_RPT], bc[TC_RPT];
        bool bok[TC_RPT];
#pragma unroll
        for (int q = 0; q < TC_RPT; ++q) {
            int n = n0 + 32 * (qb + q) + 4 * j;
            bok[q] = n < Ncols;
            int nn = bok[q] ? n : 0;
            bt[q] = nn / a.Cc;
            bc[q] = nn - bt[q] * a.Cc;
        }
        float4 ra[2][TC_RPT], rb[2][TC_RPT];
        auto load = [&](int s, float4* pa, float4* pb) {
            const int k = s * TC_BK + kl;
            const bool kok = k < Ktot;
            const int kc = kok ? k : 0;
            const int b = kc / a.dy_rows;
            const int l = kc - b * a.dy_rows;
            const float* dyr = a.dy + (long long)b * a.dy_bs + (long long)l * a.dy_ld;
#pragma unroll
            for (int q = 0; q < TC_RPT; ++q) {
                const int co = m0 + 32 * (qb + q) + 4 * j;
                pa[q] = (kok && co < a.Cout) ? __ldg(reinterpret_cast<const float4*>(dyr + co))
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            const float* xb = a.x + (long long)b * a.x_bs;
            const int rbase = l * a.sr + a.roff0;
#pragma unroll
            for (int q = 0; q < TC_RPT; ++q) {
                const int r = rbase + bt[q] * a.droff;
                const bool ok = kok && bok[q] && r >= 0 && r < a.x_rows;
                pb[q] = ok ? __ldg(reinterpret_cast<const float4*>(xb + (long long)r * a.x_ld + bc[q]))
                           : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };
        auto split_store = [&](uint8_t* hi_tile, uint8_t* lo_tile, uint32_t off, float4 v) {
            float4 h = make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
            *reinterpret_cast<float4*>(hi_tile + off) = h;
            if (NS == 3) {
                float4 lo = make_float4(to_tf32(v.x - h.x), to_tf32(v.y - h.y), to_tf32(v.z - h.z), to_tf32(v.w - h.w));
                *reinterpret_cast<float4*>(lo_tile + off) = lo;
            }
        };
        auto store = [&](int it, const float4* pa, const float4* pb) {
            const int st = it % STAGES;
            const uint32_t ph = (uint32_t)((it / STAGES) & 1);
            mbar_wait(bar_empty + 8 * st, ph ^ 1);
            uint8_t* sA = smem + (size_t)st * STAGE_BYTES;
            uint8_t* sB = sA + (NS == 3 ? 2 : 1) * TC_A_BYTES;
#pragma unroll
            for (int q = 0; q < TC_RPT; ++q) split_store(sA, sA + TC_A_BYTES, mn_off(kl, 8 * (qb + q) + j), pa[q]);
#pragma unroll
            for (int q = 0; q < TC_RPT; ++q)
                if (32 * (qb + q) < bn) split_store(sB, sB + TC_B_BYTES, mn_off(kl, 8 * (qb + q) + j), pb[q]);
            fence_proxy_async_smem();
            mbar_arrive(bar_full + 8 * st);
        };
        if (nk > 0) load(s_begin, ra[0], rb[0]);
        for (int it = 0; it < nk; it += 2) {
            if (it + 1 < nk) load(s_begin + it + 1, ra[1], rb[1]);
            store(it, ra[0], rb[0]);
            if (it + 1 < nk) {
                if (it + 2 < nk) load(s_begin + it + 2, ra[0], rb[0]);
                store(it + 1, ra[1], rb[1]);
            }
        }
    } else if (warp == TC_PW) {
        {
            // both operands MN-major: a_major (bit 15) = b_major (bit 16) = 1
            const uint32_t idesc = tf32_idesc(TC_BM, bn) | (1u << 15) | (1u << 16);
            // descriptors advance incrementally (address field = addr >> 4 in the low 14 bits): no per-instruction
            // 64-bit assembly, no division on the one thread that feeds the tensor pipe
            const uint64_t dA0 = sw128b32_desc_mn(smem_base);
            const uint64_t offB = (uint64_t)(((NS == 3 ? 2 : 1) * TC_A_BYTES) >> 4);
            uint64_t dA = dA0;
            int st = 0;
            uint32_t ph = 0;
            for (int it = 0; it < nk; ++it) {
                mbar_wait(bar_full + 8 * st, ph);
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < TC_BK / 8; ++k) {
                        const uint64_t ah = dA + (1024 >> 4) * k, bh = ah + offB;
                        if (NS == 3) {
                            umma_tf32(tmem, ah + (TC_A_BYTES >> 4), bh, idesc, (it | k) != 0);
                            umma_tf32(tmem, ah, bh + (TC_B_BYTES >> 4), idesc, 1);
                            umma_tf32(tmem, ah, bh, idesc, 1);
                        } else {
                            umma_tf32(tmem, ah, bh, idesc, (it | k) != 0);
                        }
                    }
                    umma_commit(bar_empty + 8 * st);
                }
                __syncwarp();
                dA += (uint64_t)(STAGE_BYTES >> 4);
                if (++st == STAGES) { st = 0; ph ^= 1; dA = dA0; }
            }
            if (elect_one()) umma_commit(bar_acc);
        }
        __syncwarp();
    }

    float* Cs = reinterpret_cast<float*>(smem);
    if (warp < TC_PW) {
        if (nk > 0) {
            mbar_wait(bar_acc, 0);
            tc_fence_after();
            tmem_to_smem(tmem, Cs, warp, lane, bn);
        } else {
            for (int idx = tid; idx < TC_BM * TC_CLD; idx += TC_PRODUCERS) Cs[idx] = 0.f;
        }
        tc_fence_before();
    }
    __syncthreads();
    const int Z = (int)gridDim.z;
    if (warp >= TC_PW) {
        if (warp == TC_PW) {
            tc_fence_after();
            tmem_dealloc(tmem, (uint32_t)tm_cols);
        }
        if (Z > 1) {
            cluster_sync_all();
            cluster_sync_all();
        }
        return;
    }
    // Ncols = T*Cc is a multiple of 4 here (Cc % 4 == 0), so 4-column chunks never straddle the tile edge
    const int ncols = min(bn, Ncols - n0);
    const int cpr = ncols >> 2;
    int P = 1;
    while (P < cpr) P <<= 1;
    const int rpi = TC_PRODUCERS / P;
    const int c4 = (tid & (P - 1)) * 4, rsub0 = tid / P;
    const bool cvalid = c4 < ncols;
    int row_lo = 0, row_hi = TC_BM;
    if (Z > 1) {
        cluster_sync_all();
        const int RB = TC_BM / Z;
        row_lo = (int)cluster_rank() * RB;
        row_hi = row_lo + RB;
    }
    // final values -> PyTorch parameter layout (Cout, Cc, T):  dw = beta*dw + scale*sum
    const int n = n0 + c4;
    const int t = cvalid ? n / a.Cc : 0;
    const int c = n - t * a.Cc;
    for (int rl = row_lo + rsub0; rl < row_hi; rl += rpi) {
        const int co = m0 + rl;
        if (!cvalid || co >= a.Cout) continue;
        float v[4];
        if (Z > 1) {
            cluster_reduce4(Cs, rl, c4, Z, v);
        } else {
            const float4 q = *reinterpret_cast<const float4*>(Cs + rl * TC_CLD + c4);
            v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
        }
        if (a.packed) {                                  // dw[co, t*Cc + c]: one aligned 16-byte store
            float4* d = reinterpret_cast<float4*>(a.dw + (long long)co * Ncols + n);
            float4 o = make_float4(v[0] * a.scale, v[1] * a.scale, v[2] * a.scale, v[3] * a.scale);
            if (a.beta != 0.f) {
                const float4 p = *d;
                o.x += a.beta * p.x; o.y += a.beta * p.y; o.z += a.beta * p.z; o.w += a.beta * p.w;
            }
            *d = o;
            continue;
        }
        float* d = a.dw + ((long long)co * a.Cc + c) * a.T + t;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            float x = v[u] * a.scale;
            float* du = d + (long long)u * a.T;
            if (a.beta != 0.f) x += a.beta * *du;
            *du = x;
        }
    }
    if (Z > 1) cluster_sync_all();
}

template <int NS>
static int launch_wgrad_tc(const m2d_wgrad_args& a, int Ktot, int Ncols, int want_splits, cudaStream_t st) {
    auto kern = wgrad_tc_kernel<NS>;
    static bool configured = false;
    static int zmax = 1;
    const int smem = tc_smem_bytes(NS);
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) {
            set_error("wgrad_tc: smem attribute (%d B): %s", smem, cudaGetErrorString(e));
            return M2D_ERR_CUDA;
        }
        // weight gradients have few output tiles and a long K: allow the non-portable cluster size 16
        bool np = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess;
        (void)cudaGetLastError();
        zmax = max_cluster_z(kern, smem, 16, np);
        configured = true;
    }
    int Z = 1;
    while (Z * 2 <= want_splits && Z * 2 <= zmax) Z *= 2;
    dim3 grid((unsigned)cdiv(a.Cout, TC_BM), (unsigned)cdiv(Ncols, TC_BNMAX), (unsigned)Z);
    return launch_clustered("wgrad_tc", kern, grid, smem, Z, st, TC_THREADS, a, Ktot, Ncols);
}

// Returns 1 when the shape does not qualify (caller uses the SIMT kernel); otherwise dw is final
// (split-K partials are reduced across the cluster inside the kernel).
int wgrad_tc_dispatch(const m2d_wgrad_args& a, int Ktot, int Ncols, int mode, cudaStream_t st) {
    const bool ok = a.win_T == 0 && a.Cc % 4 == 0 && a.Cout % 4 == 0 && a.x_ld % 4 == 0 && a.x_bs % 4 == 0 &&
                    a.dy_ld % 4 == 0 && a.dy_bs % 4 == 0 && aligned16(a.x) && aligned16(a.dy) &&
                    (!a.packed || aligned16(a.dw));
    if (!ok || (long long)a.Cout * Ncols * Ktot < (1ll << 18)) return 1;
    const int nsteps = (int)cdiv(Ktot, TC_BK);
    const long long tiles = cdiv(a.Cout, TC_BM) * cdiv(Ncols, TC_BNMAX);
    // split K until the launch fills ONE wave of SMs: weight gradients are leaves of the step's dependency graph and
    // run next to the tangent chain, so their SM footprint counts, not their latency (measured on B200, batch 7,
    // train steps/s by target CTAs: 74 -> 60.4, 120 -> 62.4, 148 -> 62.6, 185 -> 62.3, 296 -> 61.6, 592 -> 61.0)
    static const int target = getenv("M2D_WGRAD_CTAS") ? atoi(getenv("M2D_WGRAD_CTAS")) : kNumSMs;
    long long want = cdiv(target, tiles);
    int splits = (int)(want < nsteps / 2 ? want : nsteps / 2);
    if (splits < 1) splits = 1;
    // TF32_BF16 applies to the row convolutions only: weight gradients (MN-major operands) stay 3xTF32
    return mode != M2D_GEMM_TF32 ? launch_wgrad_tc<3>(a, Ktot, Ncols, splits, st) : launch_wgrad_tc<1>(a, Ktot, Ncols, splits, st);
}

}  // namespace m2d
