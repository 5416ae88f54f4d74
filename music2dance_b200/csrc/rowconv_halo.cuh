// Row-convolution GEMM on tcgen05 with BOTH operands fed by TMA and the activation operand staged
// ONCE per (stride residue, channel block) as a halo tile shared by all taps of the CTA.
// Included by rowconv_tc.cu (same translation unit: shares its host-side TMA / cluster helpers).
//
//   y[b,i,n] = epi( sum_t sum_c x[b, i*sr + roff0 + t*droff, c] * w[n, t*Cc + c] )
//
// Write the row index as (i + q_t)*sr + r_t  (q_t = floor((roff0 + t*droff)/sr), r_t the residue).  The
// channels-last activation x[b, l, c] viewed as a 4-D tensor (c, r, l' = l/sr, b) makes "all rows of
// residue r" a plane with contiguous rows, so the operand of tap t for output rows i0..i0+127 is the
// 128-row slab of plane r_t starting at row i0 + q_t — and the taps of one residue differ only by a
// shift of +-1 row.  Per (residue group, 32-channel block) — a "unit" — the CTA therefore
//   * TMA-loads ONE halo tile [128 + Q - 1 rows][32 floats] (SWIZZLE_128B; zero fill outside the
//     tensor = the convolution's zero padding, no predicates anywhere),
//   * splits it once into TF32 hi / lo tiles (8 converter warps, chunk-for-chunk copy: raw and split
//     tiles share the swizzled layout, so the conversion is a linear 16-byte sweep),
//   * issues the MMAs of all Q taps of the unit from that tile: the shared-memory descriptor of tap j
//     starts (q_j - q_min) rows = multiples of 128 bytes into the tile (the 128-byte swizzle is a
//     function of the address bits, so a row-shifted start reads the rows the TMA wrote),
// while the weight tiles of the taps stream through their own TMA ring.  Against the SIMT-staged
// kernel (rowconv_tc_kernel) the activation operand is read from L2 and converted Q times less
// (audio_d: 25 taps / stride 4 -> 6-7x; k = 7 pose blocks: 7x) and no thread computes an address.
// One CTA tile = 128 rows of ONE batch entry x BN columns; split-K over a (1,1,Z) cluster as before.

constexpr int HL_CONV = 256;                    // converter / epilogue threads
constexpr int HL_CW = HL_CONV / 32;             // warp HL_CW: MMA issuer, +1: weight TMA, +2: activation TMA
constexpr int HL_THREADS = HL_CONV + 96;
constexpr int HL_QMAX = 9;                      // taps served by one halo tile
constexpr int HL_ROWS = TC_BM + HL_QMAX - 1;    // 136 rows (multiple of 8: whole swizzle atoms)
constexpr int HL_TILE = HL_ROWS * 128;          // 17408 B = 17 KiB
constexpr int HL_MAXG = 16;
// Shared memory: NA activation stages {hi (the TMA lands the raw fp32 tile here; converted in place), lo}
// followed by NB weight stages {hi, lo} of `brows` rows each; NA / NB / brows are chosen per launch
// (deep weight ring for many-tap units, more activation stages for 1-2 tap units, 64-row weight boxes for N <= 64).
constexpr int HL_A_STAGE = 2 * HL_TILE;         // 34 KiB
constexpr int HL_MAX_NA = 4, HL_MAX_NB = 8;
constexpr int HL_SMEM_BUDGET = 227 * 1024 - 1024 - 512;    // dynamic bytes after alignment slack and static barriers
static_assert(TC_BM * TC_CLD * 4 <= 2 * HL_A_STAGE, "C staging tile must fit in two activation stages");

struct HaloGroup {          // taps t0, t0+dt, ... (Q of them) of one residue plane r
    short r, qmin, t0, dt, Q, rowoff0, drow, pad;   // tap j reads halo rows rowoff0 + j*drow .. +127
};
struct HaloPlan {
    int ngroups, cchunks;
    int trigger;            // 1: griddepcontrol.launch_dependents right after this grid's own dependency wait (M2D_PDL_TRIGGER)
    HaloGroup g[HL_MAXG];
};

__device__ __forceinline__ void tma_load_4d(uint32_t dst_smem, const CUtensorMap* map, int c0, int c1, int c2, int c3,
                                            uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
        ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
        : "memory");
}

template <int NS, bool TRACE>
__global__ void __launch_bounds__(HL_THREADS, 1)
rowconv_halo_kernel(const m2d_rowconv_args a, const HaloPlan plan, const int tpb, const int NA, const int NB,
                    const int brows, long long* __restrict__ trace, const __grid_constant__ CUtensorMap map_x) {
    // trace (M2D_HALO_TRACE=1, bring-up only): per CTA 16 clock64 stamps / accumulated wait times
    long long* tr = (TRACE && trace) ? trace + 16ll * (blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z)) : nullptr;
    const long long t_entry = TRACE ? clock64() : 0;
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bars[3 * HL_MAX_NA + 2 * HL_MAX_NB + 1];
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar0 = smem_u32(&bars[0]);
    const uint32_t raw_full = bar0, a_full = bar0 + 8 * HL_MAX_NA, a_empty = bar0 + 16 * HL_MAX_NA;
    const uint32_t b_full = bar0 + 24 * HL_MAX_NA, b_empty = b_full + 8 * HL_MAX_NB, bar_acc = b_empty + 8 * HL_MAX_NB;
    const uint32_t b_plane = (uint32_t)brows * 128u;                 // bytes of one weight plane (hi or lo)
    const uint32_t off_b = (uint32_t)NA * HL_A_STAGE;

    const int b = blockIdx.x / tpb;
    const int i0 = (blockIdx.x - b * tpb) * TC_BM;
    const int n0 = blockIdx.y * TC_BNMAX;
    int bn = a.N - n0;
    bn = bn > TC_BNMAX ? TC_BNMAX : ((bn + 15) & ~15);
    const int tm_cols = bn <= 32 ? 32 : (bn <= 64 ? 64 : 128);
    const int cch = plan.cchunks;
    const int nunits = plan.ngroups * cch;
    const int per = (nunits + gridDim.z - 1) / gridDim.z;
    const int u_begin = blockIdx.z * per;
    const int nu = max(0, min(nunits, u_begin + per) - u_begin);

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
        mbar_init(bar_acc, 1);
        fence_barrier_init();
    }
    if (warp == HL_CW) tmem_alloc(smem_u32(&tmem_slot), (uint32_t)tm_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    asm volatile("griddepcontrol.wait;" ::: "memory");
    // Let the next kernel of the stream start its prologue (barrier init, TMEM allocation) on idle SMs now; its own
    // griddepcontrol.wait still holds it until this grid has completed and flushed.
    if (plan.trigger) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const long long t_start = TRACE ? clock64() : 0;

    if (warp < HL_CW) {
        // ------------------------------------------------------------------ converters: raw -> TF32 hi / lo
        long long w_raw = 0, t_conv = 0;
        for (int n = 0; n < nu; ++n) {
            const int s = n % NA;
            const uint32_t ph = (uint32_t)((n / NA) & 1);
            const long long c0 = TRACE ? clock64() : 0;
            mbar_wait(raw_full + 8 * s, ph);
            const long long c1 = TRACE ? clock64() : 0;
            w_raw += c1 - c0;
            uint8_t* hi = smem + s * HL_A_STAGE;                 // raw fp32 tile -> TF32 hi, in place
            uint8_t* lo = hi + HL_TILE;
#pragma unroll
            for (int idx = tid; idx < HL_TILE / 16; idx += HL_CONV) {
                const float4 v = *reinterpret_cast<const float4*>(hi + 16 * idx);
                const float4 h = make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
                *reinterpret_cast<float4*>(hi + 16 * idx) = h;
                if (NS == 3)
                    *reinterpret_cast<float4*>(lo + 16 * idx) =
                        make_float4(to_tf32(v.x - h.x), to_tf32(v.y - h.y), to_tf32(v.z - h.z), to_tf32(v.w - h.w));
                if (NS == 2) {      // BF16 plane [bf16(hi) | bf16(lo)]: the sweep is physical, recover the logical chunk
                    const int r = idx >> 3;
                    store_bf16_pair(lo, r, (idx & 7) ^ (r & 7), h, make_float4(v.x - h.x, v.y - h.y, v.z - h.z, v.w - h.w));
                }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_full + 8 * s);
            if (TRACE) t_conv += clock64() - c1;
        }
        if (TRACE && tr && tid == 0) { tr[6] = w_raw; tr[7] = t_conv; }
    } else if (warp == HL_CW) {
        // ------------------------------------------------------------------ MMA issuer
        {
            // This loop is the critical path of the kernel (one thread feeds the tensor pipe): slots, phases and
            // descriptors advance incrementally, no division, no 64-bit descriptor assembly per instruction.  The
            // address field of a descriptor is (addr >> 4) in its low 14 bits, so "+ 32*k bytes" is "+ 2*k" and a
            // row shift of the halo tile is "+ 8 per row" on the 64-bit value.
            const uint32_t idesc = tf32_idesc(TC_BM, bn);
            const uint32_t idesc16 = bf16_idesc(TC_BM, bn);
            const uint64_t dB0 = sw128_desc(smem_base + off_b);
            const uint64_t dBstep = (uint64_t)((2u * b_plane) >> 4), dBlo = (uint64_t)(b_plane >> 4);
            uint64_t dB = dB0;
            int st = 0, s = 0;
            uint32_t bph = 0, aph = 0, acc = 0;
            long long w_a = 0, w_b = 0;
            for (int n = 0; n < nu; ++n) {
                const HaloGroup G = plan.g[(u_begin + n) / cch];
                const long long c0 = TRACE ? clock64() : 0;
                mbar_wait(a_full + 8 * s, aph);
                if (TRACE) w_a += clock64() - c0;
                // tap j: descriptor starts (q_j - q_min) rows into the halo tile; the 128-byte swizzle is a function
                // of the shared-memory address bits, so the base-offset field stays 0 (verified on B200)
                uint64_t dA = sw128_desc(smem_base + s * HL_A_STAGE) + (uint64_t)(8 * G.rowoff0);
                const long long dAstep = 8 * G.drow;
                for (int j = 0; j < G.Q; ++j) {
                    const long long c2 = TRACE ? clock64() : 0;
                    mbar_wait(b_full + 8 * st, bph);
                    if (TRACE) w_b += clock64() - c2;
                    tc_fence_after();
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < TC_BK / 8; ++k) {
                            const uint64_t ah = dA + 2 * k, bh = dB + 2 * k;
                            if (NS == 3) {
                                umma_tf32(tmem, ah + (HL_TILE >> 4), bh, idesc, k == 0 ? acc : 1u);
                                umma_tf32(tmem, ah, bh + dBlo, idesc, 1);
                                umma_tf32(tmem, ah, bh, idesc, 1);
                            } else {
                                umma_tf32(tmem, ah, bh, idesc, k == 0 ? acc : 1u);
                            }
                        }
                        if (NS == 2) {      // both cross terms: [bf16(a_hi) | bf16(a_lo)] x [bf16(b_lo) | bf16(b_hi)], K = 64
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                umma_bf16(tmem, dA + (HL_TILE >> 4) + 2 * k, dB + dBlo + 2 * k, idesc16, 1);
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
            if (elect_one()) umma_commit(bar_acc);
            if (TRACE && tr && lane == 0) { tr[0] = t_start - t_entry; tr[1] = clock64() - t_start; tr[2] = w_a; tr[3] = w_b; }
        }
        __syncwarp();
    } else if (warp == HL_CW + 1) {
        // ------------------------------------------------------------------ weight blocks: one bulk copy per tap
        {
            const long long nt_base = (long long)blockIdx.y * a.T * cch;
            const long long blk = tiled_block_floats(brows);
            const uint32_t bytes = (uint32_t)tc_planes(NS) * b_plane;
            int st = 0;
            uint32_t bph = 1;                                   // fresh "empty" barriers pass the first round
            for (int n = 0; n < nu; ++n) {
                const int u = u_begin + n;
                const int g = u / cch, c = u - g * cch;
                const HaloGroup G = plan.g[g];
                const float* src = a.w_tiled + (nt_base + (long long)G.t0 * cch + c) * blk;
                const long long sstep = (long long)G.dt * cch * blk;
                for (int j = 0; j < G.Q; ++j) {
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
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ activation halo tiles (one per unit)
        {
            for (int n = 0; n < nu; ++n) {
                const int u = u_begin + n;
                const int g = u / cch, c = u - g * cch;
                const HaloGroup G = plan.g[g];
                const int s = n % NA;
                mbar_wait(a_empty + 8 * s, (uint32_t)(((n / NA) & 1) ^ 1));      // the MMAs of unit n - NA are done with the stage
                if (elect_one()) {
                    mbar_arrive_expect_tx(raw_full + 8 * s, HL_TILE);
                    tma_load_4d(smem_base + s * HL_A_STAGE, &map_x, c * TC_BK, G.r, i0 + G.qmin, b, raw_full + 8 * s);
                }
                __syncwarp();
            }
        }
        __syncwarp();
    }

    // ---------------------------------------------------------------------- epilogue
    float* Cs = reinterpret_cast<float*>(smem);      // [128][TC_CLD] over the (drained) activation stages
    if (warp < HL_CW) {
        if (nu > 0) {
            mbar_wait(bar_acc, 0);
            if (TRACE && tr && tid == 0) tr[4] = clock64() - t_start;
            tc_fence_after();
            // The staging tile reuses the activation stages.  Their last writers (these same 256 threads, as converters)
            // are ordered before this point through a_full -> tcgen05.mma -> tcgen05.commit -> bar_acc; a plain barrier
            // among the 256 states the same order in a form compute-sanitizer's racecheck can follow.
            epi_bar();
            const int q = warp & 3, part = warp >> 2;
            const int row = 32 * q + lane;
            for (int ch = part; ch < bn / 16; ch += HL_CW / 4) {
                float v[16];
                tmem_ld16(tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(16 * ch), v);
                float4* dst = reinterpret_cast<float4*>(Cs + row * TC_CLD + 16 * ch);
#pragma unroll
                for (int u = 0; u < 4; ++u) dst[u] = make_float4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
            }
        } else {
            for (int idx = tid; idx < TC_BM * TC_CLD; idx += HL_CONV) Cs[idx] = 0.f;
        }
        tc_fence_before();
    }
    __syncthreads();
    const int Z = (int)gridDim.z;
    if (warp >= HL_CW) {
        if (warp == HL_CW) {
            tc_fence_after();
            tmem_dealloc(tmem, (uint32_t)tm_cols);
        }
        if (Z > 1) {
            cluster_sync_all();
            cluster_sync_all();
        }
        return;
    }
    const int ncols = min(bn, a.N - n0);
    const int cpr = (ncols + 3) >> 2;
    int P = 1;
    while (P < cpr) P <<= 1;
    const int rpi = HL_CONV / P;
    const int c4 = (tid & (P - 1)) * 4, rsub0 = tid / P;
    const bool cvalid = c4 < ncols;
    const bool vecN = (a.N & 3) == 0 && ncols - c4 >= 4;
    int row_lo = 0, row_hi = TC_BM;
    if (Z > 1) {
        cluster_sync_all();
        const int RB = (TC_BM + Z - 1) / Z;              // any cluster size 2..8, not only powers of two
        row_lo = (int)cluster_rank() * RB;
        row_hi = min(TC_BM, row_lo + RB);
    }
    const bool vy = vecN && (a.y_ld & 3) == 0 && (a.y_bs & 3) == 0 && aligned16d(a.y) &&
                    (!a.y2 || aligned16d(a.y2)) &&
                    (!a.mask_mode || ((a.m_ld & 3) == 0 && (a.m_bs & 3) == 0 && aligned16d(a.mask))) &&
                    (!a.add || ((a.a_ld & 3) == 0 && (a.a_bs & 3) == 0 && aligned16d(a.add)));
    // every thread owns ONE 4-column chunk: its bias values are loop invariants
    const int n = n0 + c4;
    float bz[4] = {0.f, 0.f, 0.f, 0.f};
    if (a.bias && cvalid) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (c4 + u < ncols) bz[u] = __ldg(a.bias + n + u);
    }
    for (int rl = row_lo + rsub0; rl < row_hi; rl += rpi) {
        const int i = i0 + rl;
        if (!cvalid || i >= a.y_rows) continue;
        float v[4];
        if (Z > 1) {
            cluster_reduce4(Cs, rl, c4, Z, v);
        } else {
            const float4 t = *reinterpret_cast<const float4*>(Cs + rl * TC_CLD + c4);
            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        }
        const long long yo = b * a.y_bs + (long long)i * a.y_ld + n;
        const long long mo = a.mask_mode ? b * a.m_bs + (long long)i * a.m_ld + n : 0;
        const long long ao = a.add ? b * a.a_bs + (long long)i * a.a_ld + n : 0;
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
                float x = v[u] + bz[u];
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
                float x = v[u] + bz[u];
                x = apply_act(x, a.act);
                if (a.add && a.add_before_mask) x += a.add[ao + u];
                if (a.y2) a.y2[yo + u] = x;
                if (a.mask_mode) x *= act_deriv(a.mask[mo + u], a.mask_mode);
                if (a.add && !a.add_before_mask) x += a.add[ao + u];
                a.y[yo + u] = x;
            }
        }
    }
    if (Z > 1) cluster_sync_all();
    if (TRACE && tr && tid == 0) tr[5] = clock64() - t_start;
}

// ---- host side -------------------------------------------------------------------------------
static int halo_env(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

// 4-D view (c, r, l' = l / sr, b) of a channels-last activation; box = 32 channels x 1 x HL_ROWS x 1
static const CUtensorMap* act_tmap(const float* x, int Cc, int sr, int x_rows, int nb, int x_ld, long long x_bs) {
    typedef std::tuple<const void*, int, int, int, int, int, long long> Key;
    static std::map<Key, CUtensorMap> cache;
    Key key = std::make_tuple((const void*)x, Cc, sr, x_rows, nb, x_ld, x_bs);
    auto it = cache.find(key);
    if (it != cache.end()) return &it->second;
    tmap_encode_fn enc = tmap_encoder();
    if (!enc) return nullptr;
    CUtensorMap m;
    const long long bs = nb > 1 ? x_bs : (long long)x_rows * x_ld;
    cuuint64_t dims[4] = {(cuuint64_t)Cc, (cuuint64_t)sr, (cuuint64_t)(x_rows / sr), (cuuint64_t)nb};
    cuuint64_t strides[3] = {(cuuint64_t)x_ld * 4, (cuuint64_t)sr * x_ld * 4, (cuuint64_t)bs * 4};
    cuuint32_t box[4] = {TC_BK, 1, HL_ROWS, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return nullptr;
    return &cache.emplace(key, m).first->second;
}

static int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }

// taps grouped by stride residue, at most HL_QMAX per group; false if the shape needs more than HL_MAXG groups
static bool halo_plan(const m2d_rowconv_args& a, HaloPlan& p) {
    p.ngroups = 0;
    p.cchunks = (int)cdiv(a.Cc, TC_BK);
    static const int trig = halo_env("M2D_PDL_TRIGGER", 0);
    p.trigger = trig;
    const int sr = a.sr;
    for (int r = 0; r < sr; ++r) {
        // taps of residue r are sr apart in t; the first one:
        int t = -1;
        for (int tt = 0; tt < a.T && tt < sr; ++tt) {
            const int off = a.roff0 + tt * a.droff;
            if (off - floordiv(off, sr) * sr == r) { t = tt; break; }
        }
        if (t < 0) continue;
        while (t < a.T) {
            int Q = 0;
            for (int tt = t; tt < a.T && Q < HL_QMAX; tt += sr) ++Q;
            if (p.ngroups >= HL_MAXG) return false;
            const int q0 = floordiv(a.roff0 + t * a.droff, sr);
            const int q1 = q0 + (Q - 1) * a.droff;          // consecutive taps of a residue: q moves by droff
            HaloGroup g;
            g.r = (short)r; g.t0 = (short)t; g.dt = (short)sr; g.Q = (short)Q;
            g.qmin = (short)(q0 < q1 ? q0 : q1);
            g.rowoff0 = (short)(q0 - g.qmin);
            g.drow = (short)a.droff;
            g.pad = 0;
            p.g[p.ngroups++] = g;
            t += Q * sr;
        }
    }
    return p.ngroups > 0;
}

template <int NS, bool TRACE>
static int launch_halo(const m2d_rowconv_args& a, const HaloPlan& plan, int tpb, int want_splits, cudaStream_t st,
                       int NA, int NB, int brows, const CUtensorMap* mx) {
    auto kern = rowconv_halo_kernel<NS, TRACE>;
    static bool configured = false;
    static int zok[9] = {0};                            // max co-resident clusters of size z (queried once; 0: not schedulable)
    const int smem_max = HL_SMEM_BUDGET + 1024;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max);
        if (e != cudaSuccess) {
            set_error("rowconv_halo: smem attribute (%d B): %s", smem_max, cudaGetErrorString(e));
            return M2D_ERR_CUDA;
        }
        zok[1] = kNumSMs;
        for (int z = 2; z <= 8; ++z) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(1, 1, (unsigned)z);
            cfg.blockDim = dim3(HL_THREADS);
            cfg.dynamicSmemBytes = (size_t)smem_max;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = 1;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = (unsigned)z;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            int n = 0;
            zok[z] = (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) == cudaSuccess && n > 0) ? n : 0;
            (void)cudaGetLastError();
        }
        configured = true;
    }
    const int smem = NA * HL_A_STAGE + NB * 2 * brows * 128 + 1024;
    // largest cluster size <= want_splits whose clusters are all co-resident (odd sizes fragment the GPCs: fewer
    // clusters fit than SMs / z); if none fits in one wave, the largest schedulable size
    const long long clusters = (long long)a.nb * tpb * cdiv(a.N, TC_BNMAX);
    int Z = want_splits < 8 ? want_splits : 8;
    while (Z > 1 && !(zok[Z] > 0 && clusters <= zok[Z])) --Z;
    if (Z < 1) Z = 1;
    if (Z == 1 && want_splits > 1) {
        Z = want_splits < 8 ? want_splits : 8;
        while (Z > 1 && !zok[Z]) --Z;
    }
    dim3 grid((unsigned)(a.nb * tpb), (unsigned)cdiv(a.N, TC_BNMAX), (unsigned)Z);
    // bring-up trace: 16 counters per CTA into the caller's workspace (see tools/halo_probe.py trace)
    long long* trace = nullptr;
    if (TRACE && a.ws && a.ws_floats >= 32ll * grid.x * grid.y * grid.z) trace = reinterpret_cast<long long*>(a.ws);
    return launch_clustered("rowconv_halo", kern, grid, smem, Z, st, HL_THREADS, a, plan, tpb, NA, NB, brows, trace, *mx);
}

#include "rowconv_halo_persist.cuh"

static long long g_halo_launches = 0;
static long long g_halo_persist_launches = 0;

// Returns 1 when the shape is not served by the halo kernel (caller continues with rowconv_tc_kernel).
static int rowconv_halo_dispatch(const m2d_rowconv_args& a, int M, int mode, cudaStream_t st) {
    static const int enabled = halo_env("M2D_HALO", 1);
    if (!enabled) return 1;
    if (a.win_T > 0 || a.Cc < 4 || a.Cc % 4 || a.x_ld % 4 || (a.nb > 1 && a.x_bs % 4) || !aligned16(a.x)) return 1;
    if (a.sr < 1 || (a.droff != 1 && a.droff != -1) || a.x_rows % a.sr) return 1;
    if (!a.w_tiled || !aligned16(a.w_tiled)) return 1;
    const int tpb = (int)cdiv(a.y_rows, TC_BM);
    const long long tiles_m = (long long)a.nb * tpb;
    if (tiles_m > 2 * cdiv(M, TC_BM)) return 1;              // short rows per batch entry: too many idle MMA rows
    HaloPlan plan;
    if (!halo_plan(a, plan)) return 1;
    const CUtensorMap* mx = act_tmap(a.x, a.Cc, a.sr, a.x_rows, a.nb, a.x_ld, a.x_bs);
    const int brows = tiled_rows(a.N);                       // rows of one weight block
    if (!mx) return 1;
    // ring depths: units with many taps keep the tensor core busy for a long time per activation tile (2 stages are
    // enough) and want the deepest weight ring; 1-3 tap units (Linear layers, k <= 4 convolutions) turn over quickly
    int taps = 0;
    for (int g = 0; g < plan.ngroups; ++g) taps += plan.g[g].Q;
    int NA = (taps >= 4 * plan.ngroups) ? 2 : 3;
    NA = halo_env("M2D_HALO_NA", NA);
    int NB = (HL_SMEM_BUDGET - NA * HL_A_STAGE) / (2 * brows * 128);
    if (NB > HL_MAX_NB) NB = HL_MAX_NB;
    static const int nb_cap = halo_env("M2D_HALO_NB", HL_MAX_NB);
    if (NB > nb_cap) NB = nb_cap;
    if (NB < 2 || NA < 2 || NA > HL_MAX_NA) return 1;
    const int nunits = plan.ngroups * plan.cchunks;
    const long long tiles = tiles_m * cdiv(a.N, TC_BNMAX);
    int splits = 1;
    static const int split_target = halo_env("M2D_SPLIT_CTAS", kNumSMs);
    if (tiles < split_target && nunits >= 2) {
        long long want = split_target / tiles;               // keep the launch inside ONE wave of CTAs
        splits = (int)(want < nunits ? want : nunits);
        if (splits < 1) splits = 1;
    }
    ++g_halo_launches;
    static const int trace_on = halo_env("M2D_HALO_TRACE", 0);
    // several waves of tiles and no split-K: persistent CTAs with double-buffered accumulators (the epilogue
    // of a tile overlaps the next tile's MMAs); M2D_HALO_PERSIST=0 keeps the one-tile-per-CTA kernel
    static const int persist_on = halo_env("M2D_HALO_PERSIST", 1);
    // (measured on B200: with only two tiles per CTA the un-overlapped last epilogue eats the gain; three waves and more win)
    static const int persist_min = halo_env("M2D_HALO_PERSIST_MIN", 3 * kNumSMs);
    if (persist_on && !trace_on && splits == 1 && tiles >= persist_min) {
        ++g_halo_persist_launches;
        if (mode == M2D_GEMM_TF32_BF16) return launch_halo_persist<2>(a, plan, tpb, st, NA, NB, brows, mx);
        return mode == 3 ? launch_halo_persist<3>(a, plan, tpb, st, NA, NB, brows, mx)
                         : launch_halo_persist<1>(a, plan, tpb, st, NA, NB, brows, mx);
    }
    if (mode == M2D_GEMM_TF32_BF16) return launch_halo<2, false>(a, plan, tpb, splits, st, NA, NB, brows, mx);
    if (trace_on)
        return mode == 3 ? launch_halo<3, true>(a, plan, tpb, splits, st, NA, NB, brows, mx)
                         : launch_halo<1, true>(a, plan, tpb, splits, st, NA, NB, brows, mx);
    return mode == 3 ? launch_halo<3, false>(a, plan, tpb, splits, st, NA, NB, brows, mx)
                     : launch_halo<1, false>(a, plan, tpb, splits, st, NA, NB, brows, mx);
}
