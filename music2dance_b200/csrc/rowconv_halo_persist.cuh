// Persistent variant of rowconv_halo_kernel for launches with at least one wave of output tiles (no split-K):
// one CTA per SM walks a static list of tiles, the accumulator is DOUBLE-BUFFERED in tensor memory and a dedicated
// group of four epilogue warps drains tile i (tcgen05.ld -> bias / activation / mask / residual -> global) while the
// issuer, the loaders and the converter warps already run tile i+1.  The non-persistent kernel pays prologue
// (barrier init, TMEM allocation, pipeline fill) and epilogue once per tile with nothing overlapping them; for the
// short-K layers (audio_d.l2 forward / tangent: K = 800, its backward-data: K = 448) that was more than half of a
// tile's time (ncu r01: tensor pipe active 25 % on those launches, 54-66 % on the long-K ones).
//
// Same operand staging as rowconv_halo_kernel (one TMA halo tile per (residue group, 32-channel block), row-shifted
// descriptors per tap, pre-tiled weight blocks by bulk copy); the pipelines simply keep running across tile
// boundaries.  The epilogue reads its 128 x BN tile straight from tensor memory — thread = output row, 16 columns per
// tcgen05.ld — and writes 64-byte row segments, so no shared-memory staging tile competes with the operand rings.
//
// Warp roles (640 threads): 0-7 converters, 8 MMA issuer, 9 weight loader, 10 activation loader, 11 idle,
// 12-19 epilogue (warp w owns TMEM lanes 32*(w%4).. as tcgen05.ld requires; the two warps of a lane quarter
// take alternate 16-column chunks).

constexpr int HP_THREADS = 640;
constexpr int HP_EPI0 = 12;                     // first epilogue warp
constexpr int HP_EPI_WARPS = 8;
constexpr int HP_TMEM_COLS = 256;               // two 128-column accumulators

__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <int NS>
__global__ void __launch_bounds__(HP_THREADS, 1)
rowconv_halo_persist_kernel(const m2d_rowconv_args a, const HaloPlan plan, const int tpb, const int ntiles_m,
                            const int ntiles, const int NA, const int NB, const int brows,
                            const __grid_constant__ CUtensorMap map_x) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bars[3 * HL_MAX_NA + 2 * HL_MAX_NB + 4];
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar0 = smem_u32(&bars[0]);
    const uint32_t raw_full = bar0, a_full = bar0 + 8 * HL_MAX_NA, a_empty = bar0 + 16 * HL_MAX_NA;
    const uint32_t b_full = bar0 + 24 * HL_MAX_NA, b_empty = b_full + 8 * HL_MAX_NB;
    const uint32_t t_full = b_empty + 8 * HL_MAX_NB, t_empty = t_full + 16;     // two accumulators each
    const uint32_t b_plane = (uint32_t)brows * 128u;
    const uint32_t off_b = (uint32_t)NA * HL_A_STAGE;
    const int cch = plan.cchunks;
    const int nunits = plan.ngroups * cch;
    // tiles of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...; tile id = n_tile * ntiles_m + m_tile so that the
    // CTAs running at the same time share their weight blocks in L2
    const int my_tiles = ((int)blockIdx.x < ntiles) ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (tid == 0) {
        for (int s = 0; s < NA; ++s) {
            mbar_init(raw_full + 8 * s, 1);
            mbar_init(a_full + 8 * s, HL_CW);
            mbar_init(a_empty + 8 * s, 1);
        }
        for (int s = 0; s < NB; ++s) {
            mbar_init(b_full + 8 * s, 1);
            mbar_init(b_empty + 8 * s, 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(t_full + 8 * s, 1);
            mbar_init(t_empty + 8 * s, HP_EPI_WARPS);   // one arrival per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == HL_CW) tmem_alloc(smem_u32(&tmem_slot), (uint32_t)HP_TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (plan.trigger) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (warp < HL_CW) {
        // ------------------------------------------------------------------ converters: raw -> TF32 hi / lo
        const int total = my_tiles * nunits;
        for (int n = 0; n < total; ++n) {
            const int s = n % NA;
            const uint32_t ph = (uint32_t)((n / NA) & 1);
            mbar_wait(raw_full + 8 * s, ph);
            uint8_t* hi = smem + s * HL_A_STAGE;
            uint8_t* lo = hi + HL_TILE;
#pragma unroll
            for (int idx = tid; idx < HL_TILE / 16; idx += HL_CONV) {
                const float4 v = *reinterpret_cast<const float4*>(hi + 16 * idx);
                const float4 h = make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
                *reinterpret_cast<float4*>(hi + 16 * idx) = h;
                if (NS == 3)
                    *reinterpret_cast<float4*>(lo + 16 * idx) =
                        make_float4(to_tf32(v.x - h.x), to_tf32(v.y - h.y), to_tf32(v.z - h.z), to_tf32(v.w - h.w));
                if (NS == 2) {
                    const int r = idx >> 3;
                    store_bf16_pair(lo, r, (idx & 7) ^ (r & 7), h, make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w));
                }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_full + 8 * s);
        }
    } else if (warp == HL_CW) {
        // ------------------------------------------------------------------ MMA issuer
        const uint64_t dB0 = sw128_desc(smem_base + off_b);
        const uint64_t dBstep = (uint64_t)((2u * b_plane) >> 4), dBlo = (uint64_t)(b_plane >> 4);
        uint64_t dB = dB0;
        int st = 0, s = 0;
        uint32_t bph = 0, aph = 0;
        for (int j = 0; j < my_tiles; ++j) {
            const int tile = (int)blockIdx.x + j * (int)gridDim.x;
            const int nt_ = tile / ntiles_m;
            int bn = a.N - nt_ * TC_BNMAX;
            bn = bn > TC_BNMAX ? TC_BNMAX : ((bn + 15) & ~15);
            const uint32_t idesc = tf32_idesc(TC_BM, bn);
            const uint32_t idesc16 = bf16_idesc(TC_BM, bn);
            const int buf = j & 1;
            const uint32_t acc_addr = tmem + (uint32_t)(buf * 128);
            mbar_wait(t_empty + 8 * buf, (uint32_t)(((j >> 1) & 1) ^ 1));     // epilogue of tile j-2 has drained it
            tc_fence_after();
            uint32_t acc = 0;
            for (int n = 0; n < nunits; ++n) {
                const HaloGroup G = plan.g[n / cch];
                mbar_wait(a_full + 8 * s, aph);
                uint64_t dA = sw128_desc(smem_base + s * HL_A_STAGE) + (uint64_t)(8 * G.rowoff0);
                const long long dAstep = 8 * G.drow;
                for (int q = 0; q < G.Q; ++q) {
                    mbar_wait(b_full + 8 * st, bph);
                    tc_fence_after();
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < TC_BK / 8; ++k) {
                            const uint64_t ah = dA + 2 * k, bh = dB + 2 * k;
                            if (NS == 3) {
                                umma_tf32(acc_addr, ah + (HL_TILE >> 4), bh, idesc, k == 0 ? acc : 1u);
                                umma_tf32(acc_addr, ah, bh + dBlo, idesc, 1);
                                umma_tf32(acc_addr, ah, bh, idesc, 1);
                            } else {
                                umma_tf32(acc_addr, ah, bh, idesc, k == 0 ? acc : 1u);
                            }
                        }
                        if (NS == 2) {
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                umma_bf16(acc_addr, dA + (HL_TILE >> 4) + 2 * k, dB + dBlo + 2 * k, idesc16, 1);
                        }
                        umma_commit(b_empty + 8 * st);
                    }
                    __syncwarp();
                    acc = 1;
                    dA += dAstep;
                    dB += dBstep;
                    if (++st == NB) { st = 0; bph ^= 1; dB = dB0; }
                }
                if (elect_one()) umma_commit(a_empty + 8 * s);
                __syncwarp();
                if (++s == NA) { s = 0; aph ^= 1; }
            }
            if (elect_one()) umma_commit(t_full + 8 * buf);
            __syncwarp();
        }
    } else if (warp == HL_CW + 1) {
        // ------------------------------------------------------------------ weight blocks: one bulk copy per tap
        const long long blk = tiled_block_floats(brows);
        const uint32_t bytes = (uint32_t)tc_planes(NS) * b_plane;
        int st = 0;
        uint32_t bph = 1;
        for (int j = 0; j < my_tiles; ++j) {
            const int tile = (int)blockIdx.x + j * (int)gridDim.x;
            const long long nt_base = (long long)(tile / ntiles_m) * a.T * cch;
            for (int n = 0; n < nunits; ++n) {
                const int g = n / cch, c = n - g * cch;
                const HaloGroup G = plan.g[g];
                const float* src = a.w_tiled + (nt_base + (long long)G.t0 * cch + c) * blk;
                const long long sstep = (long long)G.dt * cch * blk;
                for (int q = 0; q < G.Q; ++q) {
                    mbar_wait(b_empty + 8 * st, bph);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(b_full + 8 * st, bytes);
                        bulk_load(smem_base + off_b + st * 2 * b_plane, src, bytes, b_full + 8 * st);
                    }
                    __syncwarp();
                    src += sstep;
                    if (++st == NB) { st = 0; bph ^= 1; }
                }
            }
        }
    } else if (warp == HL_CW + 2) {
        // ------------------------------------------------------------------ activation halo tiles (one per unit)
        int n_tot = 0;
        for (int j = 0; j < my_tiles; ++j) {
            const int tile = (int)blockIdx.x + j * (int)gridDim.x;
            const int mt = tile % ntiles_m;
            const int b = mt / tpb, i0 = (mt - b * tpb) * TC_BM;
            for (int n = 0; n < nunits; ++n, ++n_tot) {
                const int g = n / cch, c = n - g * cch;
                const HaloGroup G = plan.g[g];
                const int s = n_tot % NA;
                mbar_wait(a_empty + 8 * s, (uint32_t)(((n_tot / NA) & 1) ^ 1));
                if (elect_one()) {
                    mbar_arrive_expect_tx(raw_full + 8 * s, HL_TILE);
                    tma_load_4d(smem_base + s * HL_A_STAGE, &map_x, c * TC_BK, G.r, i0 + G.qmin, b, raw_full + 8 * s);
                }
                __syncwarp();
            }
        }
    } else if (warp >= HP_EPI0) {
        // ------------------------------------------------------------------ epilogue: TMEM -> registers -> global
        const int q = warp & 3, half = (warp - HP_EPI0) >> 2;
        const int row = 32 * q + lane;
        const bool vecN = (a.N & 3) == 0;
        const bool vy = vecN && (a.y_ld & 3) == 0 && (a.y_bs & 3) == 0 && aligned16d(a.y) &&
                        (!a.y2 || aligned16d(a.y2)) &&
                        (!a.mask_mode || ((a.m_ld & 3) == 0 && (a.m_bs & 3) == 0 && aligned16d(a.mask))) &&
                        (!a.add || ((a.a_ld & 3) == 0 && (a.a_bs & 3) == 0 && aligned16d(a.add)));
        for (int j = 0; j < my_tiles; ++j) {
            const int tile = (int)blockIdx.x + j * (int)gridDim.x;
            const int nt_ = tile / ntiles_m, mt = tile - nt_ * ntiles_m;
            const int b = mt / tpb, i0 = (mt - b * tpb) * TC_BM;
            const int n0 = nt_ * TC_BNMAX;
            int bn = a.N - n0;
            bn = bn > TC_BNMAX ? TC_BNMAX : ((bn + 15) & ~15);
            const int ncols = min(bn, a.N - n0);
            const int buf = j & 1;
            mbar_wait(t_full + 8 * buf, (uint32_t)((j >> 1) & 1));
            tc_fence_after();
            const int i = i0 + row;
            const bool rvalid = i < a.y_rows;
            const long long yo = b * a.y_bs + (long long)i * a.y_ld + n0;
            const long long mo = a.mask_mode ? b * a.m_bs + (long long)i * a.m_ld + n0 : 0;
            const long long ao = a.add ? b * a.a_bs + (long long)i * a.a_ld + n0 : 0;
            const uint32_t taddr = tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(buf * 128);
            for (int ch = half; ch < bn / 16; ch += HP_EPI_WARPS / 4) {
                uint32_t r[16];
                tmem_ld16_nowait(taddr + (uint32_t)(16 * ch), r);
                // operands of the fused epilogue are fetched while the tensor-memory load is in flight
                float4 mk4[4], ad4[4];
                const int c0 = 16 * ch;
                const bool full = vy && rvalid && c0 + 16 <= ncols;
                if (full) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        // plain loads: the mask may be the output buffer itself (in-place tangent pass; every element is
                        // read and then written by this same thread)
                        if (a.mask_mode) mk4[u] = *(reinterpret_cast<const float4*>(a.mask + mo + c0) + u);
                        if (a.add) ad4[u] = *(reinterpret_cast<const float4*>(a.add + ao + c0) + u);
                    }
                }
                tmem_ld_wait();
                if (full) {
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        float v[4], w2[4];
                        const float mkv[4] = {a.mask_mode ? mk4[u].x : 0.f, a.mask_mode ? mk4[u].y : 0.f,
                                              a.mask_mode ? mk4[u].z : 0.f, a.mask_mode ? mk4[u].w : 0.f};
                        const float adv[4] = {a.add ? ad4[u].x : 0.f, a.add ? ad4[u].y : 0.f, a.add ? ad4[u].z : 0.f,
                                              a.add ? ad4[u].w : 0.f};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            float x = __uint_as_float(r[4 * u + e]);
                            if (a.bias) x += __ldg(a.bias + n0 + c0 + 4 * u + e);
                            x = apply_act(x, a.act);
                            if (a.add && a.add_before_mask) x += adv[e];
                            w2[e] = x;
                            if (a.mask_mode) x *= act_deriv(mkv[e], a.mask_mode);
                            if (a.add && !a.add_before_mask) x += adv[e];
                            v[e] = x;
                        }
                        if (a.y2) *(reinterpret_cast<float4*>(a.y2 + yo + c0) + u) = make_float4(w2[0], w2[1], w2[2], w2[3]);
                        *(reinterpret_cast<float4*>(a.y + yo + c0) + u) = make_float4(v[0], v[1], v[2], v[3]);
                    }
                } else if (rvalid) {
#pragma unroll
                    for (int e = 0; e < 16; ++e) {           // fully unrolled: r[] stays in registers
                        if (c0 + e >= ncols) continue;
                        float x = __uint_as_float(r[e]);
                        if (a.bias) x += __ldg(a.bias + n0 + c0 + e);
                        x = apply_act(x, a.act);
                        if (a.add && a.add_before_mask) x += a.add[ao + c0 + e];
                        if (a.y2) a.y2[yo + c0 + e] = x;
                        if (a.mask_mode) x *= act_deriv(a.mask[mo + c0 + e], a.mask_mode);
                        if (a.add && !a.add_before_mask) x += a.add[ao + c0 + e];
                        a.y[yo + c0 + e] = x;
                    }
                }
            }
            // this warp's quarter of the accumulator is in registers / memory: hand the buffer back to the issuer
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(t_empty + 8 * buf);
        }
    }
    // ---------------------------------------------------------------------- teardown
    tc_fence_before();
    __syncthreads();
    if (warp == HL_CW) {
        tc_fence_after();
        tmem_dealloc(tmem, (uint32_t)HP_TMEM_COLS);
    }
}

template <int NS>
static int launch_halo_persist(const m2d_rowconv_args& a, const HaloPlan& plan, int tpb, cudaStream_t st, int NA, int NB,
                               int brows, const CUtensorMap* mx) {
    auto kern = rowconv_halo_persist_kernel<NS>;
    static bool configured = false;
    const int smem_max = HL_SMEM_BUDGET + 1024;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max);
        if (e != cudaSuccess) {
            set_error("rowconv_halo_persist: smem attribute (%d B): %s", smem_max, cudaGetErrorString(e));
            return M2D_ERR_CUDA;
        }
        configured = true;
    }
    const int smem = NA * HL_A_STAGE + NB * 2 * brows * 128 + 1024;
    const int ntiles_m = a.nb * tpb;
    const int ntiles = ntiles_m * (int)cdiv(a.N, TC_BNMAX);
    // equal tile counts per CTA where possible: ceil(tiles / waves) CTAs instead of a ragged last wave
    const int waves = (int)cdiv(ntiles, kNumSMs);
    int ctas = (int)cdiv(ntiles, waves);
    if (ctas > kNumSMs) ctas = kNumSMs;
    dim3 grid((unsigned)ctas, 1, 1);
    return launch_clustered("rowconv_halo_persist", kern, grid, smem, 1, st, HP_THREADS, a, plan, tpb, ntiles_m, ntiles,
                            NA, NB, brows, *mx);
}
