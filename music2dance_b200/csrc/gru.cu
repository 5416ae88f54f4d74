// Persistent GRU recurrence (forward and BPTT) for sm_100a.
//
// One thread-block cluster per batch group.  The cluster's CTAs shard the hidden
// units: CTA `rank` keeps the three gate rows of W_hh (forward) or the matching
// columns (backward) of its units resident in shared memory for the whole
// sequence, so W_hh is read from HBM exactly once per launch.  Each step every
// CTA computes its slice, pushes it into every peer's shared memory through
// DSMEM, and one cluster barrier publishes the step (buffers are double-buffered,
// which makes a single barrier per step sufficient).
#include <cooperative_groups.h>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace m2d {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

struct GruPlan {
    int NC;    // CTAs per cluster
    int nu;    // hidden units per CTA
    int BG;    // sequences per cluster
    int HP;    // H padded (multiple of 4, +4 to rotate banks)
    int HP3;   // 3H padded likewise
};

static GruPlan make_plan(int H) {
    GruPlan p;
    p.NC = H >= 128 ? 8 : (H >= 64 ? 4 : (H >= 32 ? 2 : 1));
    p.nu = (H + p.NC - 1) / p.NC;
    p.BG = 256 / p.nu;
    if (p.BG > 8) p.BG = 8;
    if (p.BG < 1) p.BG = 1;
    p.HP = ((H + 3) / 4) * 4 + 4;
    p.HP3 = ((3 * H + 3) / 4) * 4 + 4;
    return p;
}

template <bool SAVE>
__global__ void __launch_bounds__(256)
gru_fwd_kernel(const float* __restrict__ gi, const float* __restrict__ w_hh,
               const float* __restrict__ b_hh, float* __restrict__ h_out, int ldh,
               float* __restrict__ save, int B, int T, int H, int nu, int BG, int HP) {
    cg::cluster_group cluster = cg::this_cluster();
    const int NC = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    const int b0 = (blockIdx.x / NC) * BG;
    const int u0 = rank * nu;
    const int nown = max(0, min(nu, H - u0));
    extern __shared__ __align__(16) float smem[];
    float* Ws = smem;                       // [3][nu][HP]
    float* hs = smem + 3 * nu * HP;         // [2][BG][HP]
    const int tid = threadIdx.x;
    for (int idx = tid; idx < 3 * nu * HP; idx += blockDim.x) {
        int k = idx % HP;
        int gu = idx / HP;
        int u = gu % nu, g = gu / nu;
        Ws[idx] = (u < nown && k < H) ? w_hh[((long long)g * H + u0 + u) * H + k] : 0.f;
    }
    for (int idx = tid; idx < 2 * BG * HP; idx += blockDim.x) hs[idx] = 0.f;
    __syncthreads();
    cluster.sync();

    const int u = tid / BG, b = tid - u * BG;
    const bool active = u < nown && (b0 + b) < B && tid < nu * BG;
    const int uu = u0 + u;
    float bhr = 0.f, bhz = 0.f, bhn = 0.f;
    if (active) {
        bhr = b_hh[uu]; bhz = b_hh[H + uu]; bhn = b_hh[2 * H + uu];
    }
    const long long row0 = (long long)(b0 + b) * T;
    float hprev = 0.f;
    float gir = 0.f, giz = 0.f, gin = 0.f;
    if (active) {
        const float* g0 = gi + row0 * 3 * H;
        gir = g0[uu]; giz = g0[H + uu]; gin = g0[2 * H + uu];
    }
    for (int t = 0; t < T; ++t) {
        const int cur = t & 1;
        if (active) {
            float nr = 0.f, nz = 0.f, nn = 0.f;
            if (t + 1 < T) {       // prefetch next step's input projection
                const float* g1 = gi + (row0 + t + 1) * 3 * H;
                nr = g1[uu]; nz = g1[H + uu]; nn = g1[2 * H + uu];
            }
            const float4* wr = reinterpret_cast<const float4*>(Ws + (0 * nu + u) * HP);
            const float4* wz = reinterpret_cast<const float4*>(Ws + (1 * nu + u) * HP);
            const float4* wn = reinterpret_cast<const float4*>(Ws + (2 * nu + u) * HP);
            const float4* hv = reinterpret_cast<const float4*>(hs + (cur * BG + b) * HP);
            float ar = 0.f, az = 0.f, an = 0.f;
            const int n4 = (H + 3) / 4;
#pragma unroll 4
            for (int k = 0; k < n4; ++k) {
                float4 h4 = hv[k];
                float4 a4 = wr[k], b4 = wz[k], c4 = wn[k];
                ar = fmaf(a4.x, h4.x, ar); ar = fmaf(a4.y, h4.y, ar); ar = fmaf(a4.z, h4.z, ar); ar = fmaf(a4.w, h4.w, ar);
                az = fmaf(b4.x, h4.x, az); az = fmaf(b4.y, h4.y, az); az = fmaf(b4.z, h4.z, az); az = fmaf(b4.w, h4.w, az);
                an = fmaf(c4.x, h4.x, an); an = fmaf(c4.y, h4.y, an); an = fmaf(c4.z, h4.z, an); an = fmaf(c4.w, h4.w, an);
            }
            float r = sigmoidf_(gir + (ar + bhr));
            float z = sigmoidf_(giz + (az + bhz));
            float ghn = an + bhn;
            float n = tanhf(gin + r * ghn);
            float h = (1.f - z) * n + z * hprev;
            hprev = h;
            const long long row = row0 + t;
            h_out[row * ldh + uu] = h;
            if (SAVE) {
                float* sv = save + row * 4 * H;
                sv[uu] = r; sv[H + uu] = z; sv[2 * H + uu] = n; sv[3 * H + uu] = ghn;
            }
            const int off = ((cur ^ 1) * BG + b) * HP + uu;
            for (int rk = 0; rk < NC; ++rk) cluster.map_shared_rank(hs, rk)[off] = h;
            gir = nr; giz = nz; gin = nn;
        }
        cluster.sync();
    }
}

__global__ void __launch_bounds__(256)
gru_bwd_kernel(const float* __restrict__ dh_out, int ldd, const float* __restrict__ h_out, int ldh,
               const float* __restrict__ save, const float* __restrict__ w_hh,
               float* __restrict__ dgi, float* __restrict__ dgh, int B, int T, int H, int nu, int BG,
               int HP3) {
    cg::cluster_group cluster = cg::this_cluster();
    const int NC = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    const int b0 = (blockIdx.x / NC) * BG;
    const int u0 = rank * nu;
    const int nown = max(0, min(nu, H - u0));
    extern __shared__ __align__(16) float smem[];
    float* WT = smem;                    // [nu][HP3]: WT[u][j] = w_hh[j][u0+u]
    float* ds = smem + nu * HP3;         // [2][BG][HP3]
    const int tid = threadIdx.x;
    const int H3 = 3 * H;
    // coalesced read of w_hh rows, transposed store
    for (int idx = tid; idx < nu * HP3; idx += blockDim.x) WT[idx] = 0.f;
    for (int idx = tid; idx < 2 * BG * HP3; idx += blockDim.x) ds[idx] = 0.f;
    __syncthreads();
    for (int idx = tid; idx < H3 * nu; idx += blockDim.x) {
        int u = idx % nu, j = idx / nu;
        if (u < nown) WT[u * HP3 + j] = w_hh[(long long)j * H + u0 + u];
    }
    __syncthreads();
    cluster.sync();

    const int u = tid / BG, b = tid - u * BG;
    const bool active = u < nown && (b0 + b) < B && tid < nu * BG;
    const int uu = u0 + u;
    const long long row0 = (long long)(b0 + b) * T;
    float dh_rec = 0.f;
    // prefetched operands of the current step
    float p_dh = 0.f, p_r = 0.f, p_z = 0.f, p_n = 0.f, p_g = 0.f, p_hp = 0.f;
    auto fetch = [&](int t) {
        const long long row = row0 + t;
        p_dh = dh_out[row * ldd + uu];
        const float* sv = save + row * 4 * H;
        p_r = sv[uu]; p_z = sv[H + uu]; p_n = sv[2 * H + uu]; p_g = sv[3 * H + uu];
        p_hp = t > 0 ? h_out[(row - 1) * ldh + uu] : 0.f;
    };
    if (active) fetch(T - 1);
    for (int t = T - 1; t >= 0; --t) {
        const int cur = (T - 1 - t) & 1;
        float dh_direct = 0.f;
        if (active) {
            const float dh = p_dh + dh_rec;
            const float r = p_r, z = p_z, n = p_n, ghn = p_g, hp = p_hp;
            if (t > 0) fetch(t - 1);
            const float dn = dh * (1.f - z);
            const float dz = dh * (hp - n);
            const float dnp = dn * (1.f - n * n);
            const float dzp = dz * z * (1.f - z);
            const float drp = dnp * ghn * r * (1.f - r);
            const float dghn = dnp * r;
            const long long row = row0 + t;
            float* gi_ = dgi + row * H3;
            float* gh_ = dgh + row * H3;
            gi_[uu] = drp; gi_[H + uu] = dzp; gi_[2 * H + uu] = dnp;
            gh_[uu] = drp; gh_[H + uu] = dzp; gh_[2 * H + uu] = dghn;
            const int off = (cur * BG + b) * HP3;
            for (int rk = 0; rk < NC; ++rk) {
                float* d = cluster.map_shared_rank(ds, rk) + off;
                d[uu] = drp; d[H + uu] = dzp; d[2 * H + uu] = dghn;
            }
            dh_direct = dh * z;
        }
        cluster.sync();
        if (active && t > 0) {
            const float4* wv = reinterpret_cast<const float4*>(WT + u * HP3);
            const float4* dv = reinterpret_cast<const float4*>(ds + (cur * BG + b) * HP3);
            float a0 = 0.f, a1 = 0.f;
            const int n4 = (H3 + 3) / 4;
#pragma unroll 4
            for (int k = 0; k < n4; ++k) {
                float4 w4 = wv[k], d4 = dv[k];
                a0 = fmaf(w4.x, d4.x, a0); a1 = fmaf(w4.y, d4.y, a1);
                a0 = fmaf(w4.z, d4.z, a0); a1 = fmaf(w4.w, d4.w, a1);
            }
            dh_rec = dh_direct + (a0 + a1);
        }
    }
}

// ---------------------------------------------------------------------------- forward, v2
// Register-resident recurrence.  Cluster of NC CTAs per batch group; CTA `rank` owns
// nu <= 32 hidden units.  Thread (unit ul = tid/8, k-part kp = tid%8) keeps the three gate
// rows of its unit restricted to k in [kp*KS, (kp+1)*KS) in REGISTERS for the whole
// sequence (3*KS floats), so a step reads only h_{t-1} from shared memory.  Partial dot
// products are reduced over the 8 k-parts with warp shuffles; lane kp == b then finishes
// batch entry b of its unit (gates, h_t), stores it, and pushes h_t into every CTA of the
// cluster with st.async (DSMEM store that completes transaction bytes on the destination's
// mbarrier).  A CTA starts step t as soon as its own mbarrier has received all H*nb floats
// of h_{t-1}: one mbarrier wait per step, no cluster barrier, no __syncthreads.
__device__ __forceinline__ uint32_t gru_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t gru_mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void gru_st_async(uint32_t raddr, float v, uint32_t rbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(raddr),
                 "r"(__float_as_uint(v)), "r"(rbar)
                 : "memory");
}
__device__ __forceinline__ void gru_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void gru_mbar_expect(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Default semantics (.acquire at CTA scope): everything the waiter reads after the phase completes is SHARED memory
// written by st.async / st.shared + complete_tx on this very barrier.  The cluster-scope acquire used before made ptxas
// emit CCTL.IVALL (invalidate the whole L1) after every successful wait — 10 % of the kernel's stall samples (ncu source
// page, round 2) and a cold L1 for the input-projection loads of every step.
__device__ __forceinline__ void gru_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0, spins = 0;
    while (true) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) break;
        if (++spins > (1u << 22)) __trap();      // never hang the GPU on a protocol bug
    }
}

constexpr int GRU2_BGMAX = 8;

__device__ __forceinline__ void gru_st_async4(uint32_t raddr, float4 v, uint32_t rbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(raddr),
                 "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)),
                 "r"(__float_as_uint(v.w)), "r"(rbar)
                 : "memory");
}
// The same message inside a ONE-block "cluster" (hidden sizes <= 32: the noise GRU): an ordinary shared-memory store
// followed by the transaction-count update of the CTA's own mbarrier.  st.async is specified for the shared memory of a
// peer CTA of a real cluster; compute-sanitizer's memcheck rejects it in a 1-block launch.
__device__ __forceinline__ void gru_st_local4(uint32_t addr, float4 v, uint32_t bar) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(__float_as_uint(v.x)),
                 "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)), "r"(__float_as_uint(v.w))
                 : "memory");
    asm volatile("fence.acq_rel.cta;" ::: "memory");
    asm volatile("mbarrier.complete_tx.relaxed.cta.shared::cta.b64 [%0], 16;" ::"r"(bar) : "memory");
}

// Layout shared by the forward and backward recurrences.  A vector of length V (V = H forward,
// 3H backward) lives in shared memory as V floats padded to 32*KC; thread (unit ul = tid/8,
// k-part kp = tid%8) owns the 16-byte chunks c = 8*i + kp, i < KC, i.e. elements 4c..4c+3: the eight
// k-parts of a unit read 128 contiguous bytes per i (conflict-free, no padding), and the matching
// weight elements w[4*i + e] <-> k = 4*(8*i + kp) + e stay in registers for the whole sequence.
// Hidden units are dealt to the cluster's CTAs in blocks of nu (a multiple of 4): warp w of CTA
// `rank` owns units rank*nu + 4w .. +3, so the four results of a warp form one aligned 16-byte
// st.async per destination CTA.

template <int KC, bool SAVE>
__global__ void __launch_bounds__(256)
gru_fwd2_kernel(const float* __restrict__ gi, const float* __restrict__ w_hh, const float* __restrict__ b_hh,
                float* __restrict__ h_out, int ldh, float* __restrict__ save, int B, int T, int H, int nu,
                int BG) {
    constexpr int VP = 32 * KC;
    cg::cluster_group cluster = cg::this_cluster();
    const int NC = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    const int b0 = (blockIdx.x / NC) * BG;
    const int nb = min(BG, B - b0);
    const int u0 = rank * nu;
    __shared__ __align__(16) float hs[2][GRU2_BGMAX][VP];
    __shared__ __align__(8) uint64_t bars[2];
    const int tid = threadIdx.x, lane = tid & 31;
    const int kp = tid & 7, ul = tid >> 3;
    const int uu = u0 + ul;
    const bool active = ul < nu && uu < H;

    for (int i = tid; i < 2 * GRU2_BGMAX * VP; i += 256) (&hs[0][0][0])[i] = 0.f;
    const uint32_t bar0 = gru_smem_u32(&bars[0]);
    const uint32_t tx_bytes = 4u * (uint32_t)((H + 3) & ~3) * (uint32_t)nb;      // whole 4-unit groups are sent
    if (tid == 0) {
        gru_mbar_init(bar0, 1);
        gru_mbar_init(bar0 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (T > 1) gru_mbar_expect(bar0 + 8, tx_bytes);      // filled by the sends of step 0
        if (T > 2) gru_mbar_expect(bar0, tx_bytes);          // filled by the sends of step 1
    }
    float w[3][4 * KC];
#pragma unroll
    for (int g = 0; g < 3; ++g)
#pragma unroll
        for (int i = 0; i < KC; ++i)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int k = 4 * (8 * i + kp) + e;
                w[g][4 * i + e] = (active && k < H) ? __ldg(w_hh + ((long long)g * H + uu) * H + k) : 0.f;
            }
    const bool fin = active && kp < nb;                  // this lane finishes batch entry kp of unit uu
    // lane b (< nb) of a warp whose first unit exists sends the warp's four h values of batch entry b
    const bool sender = lane < nb && (u0 + 4 * (tid >> 5)) < H && 4 * (tid >> 5) < nu;
    float bhr = 0.f, bhz = 0.f, bhn = 0.f;
    long long row0 = 0;
    float hprev = 0.f;
    // input projections of the next GI_D steps live in registers: one step ahead is not enough — the load (L2 / HBM
    // latency, 700+ cycles) would be waited for at the end of EVERY step of a chain whose own work is ~500 cycles
    constexpr int GI_D = 4;
    float pr[GI_D], pz[GI_D], pn[GI_D];
#pragma unroll
    for (int d = 0; d < GI_D; ++d) { pr[d] = 0.f; pz[d] = 0.f; pn[d] = 0.f; }
    if (fin) {
        bhr = b_hh[uu]; bhz = b_hh[H + uu]; bhn = b_hh[2 * H + uu];
        row0 = (long long)(b0 + kp) * T;
#pragma unroll
        for (int d = 0; d < GI_D; ++d) {
            if (d < T) {
                const float* g0 = gi + (row0 + d) * 3 * H;
                pr[d] = g0[uu]; pz[d] = g0[H + uu]; pn[d] = g0[2 * H + uu];
            }
        }
    }
    const uint32_t hs_addr = gru_smem_u32(&hs[0][0][0]);
    __syncthreads();
    cluster.sync();                                      // every CTA's buffers / barriers are ready

    for (int t = 0; t < T; ++t) {
        const int cur = t & 1;
        if (t > 0) {
            const uint32_t parity = (cur ? (uint32_t)(t >> 1) : (uint32_t)((t >> 1) - 1)) & 1u;
            gru_mbar_wait(bar0 + 8 * cur, parity);
            if (tid == 0 && t + 2 <= T - 1) gru_mbar_expect(bar0 + 8 * cur, tx_bytes);
        }
        const float gir = pr[0], giz = pz[0], gin = pn[0];
#pragma unroll
        for (int d = 0; d + 1 < GI_D; ++d) { pr[d] = pr[d + 1]; pz[d] = pz[d + 1]; pn[d] = pn[d + 1]; }
        if (fin && t + GI_D < T) {                       // input projection of step t + GI_D
            const float* g1 = gi + (row0 + t + GI_D) * 3 * H;
            pr[GI_D - 1] = g1[uu]; pz[GI_D - 1] = g1[H + uu]; pn[GI_D - 1] = g1[2 * H + uu];
        }
        float ar = 0.f, az = 0.f, an = 0.f;
        for (int b = 0; b < nb; ++b) {                   // every lane takes part (full-mask shuffles)
            const float4* hv = reinterpret_cast<const float4*>(&hs[cur][b][0]) + kp;
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, q0 = 0.f, q1 = 0.f, q2 = 0.f;
#pragma unroll
            for (int i = 0; i < KC; ++i) {
                const float4 h4 = hv[8 * i];
                s0 = fmaf(w[0][4 * i], h4.x, s0); s1 = fmaf(w[1][4 * i], h4.x, s1); s2 = fmaf(w[2][4 * i], h4.x, s2);
                q0 = fmaf(w[0][4 * i + 1], h4.y, q0); q1 = fmaf(w[1][4 * i + 1], h4.y, q1); q2 = fmaf(w[2][4 * i + 1], h4.y, q2);
                s0 = fmaf(w[0][4 * i + 2], h4.z, s0); s1 = fmaf(w[1][4 * i + 2], h4.z, s1); s2 = fmaf(w[2][4 * i + 2], h4.z, s2);
                q0 = fmaf(w[0][4 * i + 3], h4.w, q0); q1 = fmaf(w[1][4 * i + 3], h4.w, q1); q2 = fmaf(w[2][4 * i + 3], h4.w, q2);
            }
            s0 += q0; s1 += q1; s2 += q2;
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) {
                s0 += __shfl_xor_sync(0xffffffffu, s0, o);
                s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                s2 += __shfl_xor_sync(0xffffffffu, s2, o);
            }
            if (kp == b) { ar = s0; az = s1; an = s2; }
        }
        float h = 0.f;
        if (fin) {
            const float r = sigmoidf_(gir + (ar + bhr));
            const float z = sigmoidf_(giz + (az + bhz));
            const float ghn = an + bhn;
            const float n = tanhf(gin + r * ghn);
            h = (1.f - z) * n + z * hprev;
            hprev = h;
            const long long row = row0 + t;
            h_out[row * ldh + uu] = h;
            if (SAVE) {
                float* sv = save + row * 4 * H;
                sv[uu] = r; sv[H + uu] = z; sv[2 * H + uu] = n; sv[3 * H + uu] = ghn;
            }
        }
        // gather the warp's four units of batch entry (lane & 7) into one 16-byte message
        float4 h4;
        h4.x = __shfl_sync(0xffffffffu, h, (lane & 7));
        h4.y = __shfl_sync(0xffffffffu, h, (lane & 7) + 8);
        h4.z = __shfl_sync(0xffffffffu, h, (lane & 7) + 16);
        h4.w = __shfl_sync(0xffffffffu, h, (lane & 7) + 24);
        if (sender && t + 1 < T) {
            const uint32_t off = (uint32_t)((((cur ^ 1) * GRU2_BGMAX + lane) * VP + u0 + 4 * (tid >> 5)) * 4);
            if (NC == 1) gru_st_local4(hs_addr + off, h4, bar0 + 8 * (cur ^ 1));
            else
                for (int rk = 0; rk < NC; ++rk)
                    gru_st_async4(gru_mapa(hs_addr + off, (uint32_t)rk), h4, gru_mapa(bar0 + 8 * (cur ^ 1), (uint32_t)rk));
        }
    }
    cluster.sync();      // no CTA leaves while a peer could still address its shared memory
}

// ---------------------------------------------------------------------------- backward, v2
// Same organisation for BPTT.  Per step (t descending) lane (unit, batch) turns dh = dh_out[t] + dh_rec
// into the three gate gradients, stores dgi / dgh, and the warp pushes them (three 16-byte messages:
// the r, z and n thirds of the 3H-vector) to every CTA; after the mbarrier wait each thread
// accumulates its slice of  dh_rec[u] = sum_j W_hh[j][u] * dgh[j]  (W_hh^T slice in registers).
template <int KC>
__global__ void __launch_bounds__(256)
gru_bwd2_kernel(const float* __restrict__ dh_out, int ldd, const float* __restrict__ h_out, int ldh,
                const float* __restrict__ save, const float* __restrict__ w_hh, float* __restrict__ dgi,
                float* __restrict__ dgh, int B, int T, int H, int nu, int BG) {
    constexpr int VP = 32 * KC;                          // padded length of ONE third (r | z | n) of the vector
    cg::cluster_group cluster = cg::this_cluster();
    const int NC = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    const int b0 = (blockIdx.x / NC) * BG;
    const int nb = min(BG, B - b0);
    const int u0 = rank * nu;
    extern __shared__ __align__(16) float ds_dyn[];      // [2][BG][3][VP]
    __shared__ __align__(8) uint64_t bars[2];
    const int tid = threadIdx.x, lane = tid & 31;
    const int kp = tid & 7, ul = tid >> 3;
    const int uu = u0 + ul;
    const bool active = ul < nu && uu < H;
    const int H3 = 3 * H;

    for (int i = tid; i < 2 * BG * 3 * VP; i += 256) ds_dyn[i] = 0.f;
    const uint32_t bar0 = gru_smem_u32(&bars[0]);
    const uint32_t tx_bytes = 3u * 4u * (uint32_t)((H + 3) & ~3) * (uint32_t)nb;
    if (tid == 0) {
        gru_mbar_init(bar0, 1);
        gru_mbar_init(bar0 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (T > 1) gru_mbar_expect(bar0, tx_bytes);          // sends of the first processed step (index 0)
        if (T > 2) gru_mbar_expect(bar0 + 8, tx_bytes);
    }
    // W_hh^T slice: wt[g][4i+e] = w_hh[g*H + j][uu],  j = 4*(8i+kp)+e
    float wt[3][4 * KC];
#pragma unroll
    for (int g = 0; g < 3; ++g)
#pragma unroll
        for (int i = 0; i < KC; ++i)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int j = 4 * (8 * i + kp) + e;
                wt[g][4 * i + e] = (active && j < H) ? __ldg(w_hh + ((long long)g * H + j) * H + uu) : 0.f;
            }
    const bool fin = active && kp < nb;
    const bool sender = lane < nb && (u0 + 4 * (tid >> 5)) < H && 4 * (tid >> 5) < nu;
    const long long row0 = (long long)(b0 + (fin ? kp : 0)) * T;
    float dh_rec = 0.f;
    float p_dh = 0.f, p_r = 0.f, p_z = 0.f, p_n = 0.f, p_g = 0.f, p_hp = 0.f;
    auto fetch = [&](int t) {
        const long long row = row0 + t;
        p_dh = dh_out[row * ldd + uu];
        const float* sv = save + row * 4 * H;
        p_r = sv[uu]; p_z = sv[H + uu]; p_n = sv[2 * H + uu]; p_g = sv[3 * H + uu];
        p_hp = t > 0 ? h_out[(row - 1) * ldh + uu] : 0.f;
    };
    if (fin) fetch(T - 1);
    const uint32_t ds_addr = gru_smem_u32(ds_dyn);
    __syncthreads();
    cluster.sync();

    for (int s = 0; s < T; ++s) {                        // s-th processed step, t = T-1-s
        const int t = T - 1 - s;
        const int cur = s & 1;                           // buffer / barrier written by this step's sends
        float drp = 0.f, dzp = 0.f, dghn = 0.f, dh_direct = 0.f;
        if (fin) {
            const float dh = p_dh + dh_rec;
            const float r = p_r, z = p_z, n = p_n, ghn = p_g, hp = p_hp;
            if (t > 0) fetch(t - 1);
            const float dn = dh * (1.f - z);
            const float dz = dh * (hp - n);
            const float dnp = dn * (1.f - n * n);
            dzp = dz * z * (1.f - z);
            drp = dnp * ghn * r * (1.f - r);
            dghn = dnp * r;
            const long long row = row0 + t;
            float* gi_ = dgi + row * H3;
            float* gh_ = dgh + row * H3;
            gi_[uu] = drp; gi_[H + uu] = dzp; gi_[2 * H + uu] = dnp;
            gh_[uu] = drp; gh_[H + uu] = dzp; gh_[2 * H + uu] = dghn;
            dh_direct = dh * z;
        }
        if (t == 0) break;                               // nothing flows further back
        float4 m0, m1, m2;
        const int src = lane & 7;
        m0.x = __shfl_sync(0xffffffffu, drp, src);      m0.y = __shfl_sync(0xffffffffu, drp, src + 8);
        m0.z = __shfl_sync(0xffffffffu, drp, src + 16); m0.w = __shfl_sync(0xffffffffu, drp, src + 24);
        m1.x = __shfl_sync(0xffffffffu, dzp, src);      m1.y = __shfl_sync(0xffffffffu, dzp, src + 8);
        m1.z = __shfl_sync(0xffffffffu, dzp, src + 16); m1.w = __shfl_sync(0xffffffffu, dzp, src + 24);
        m2.x = __shfl_sync(0xffffffffu, dghn, src);      m2.y = __shfl_sync(0xffffffffu, dghn, src + 8);
        m2.z = __shfl_sync(0xffffffffu, dghn, src + 16); m2.w = __shfl_sync(0xffffffffu, dghn, src + 24);
        if (sender) {
            const uint32_t off = (uint32_t)((((cur * BG + lane) * 3) * VP + u0 + 4 * (tid >> 5)) * 4);
            if (NC == 1) {
                gru_st_local4(ds_addr + off, m0, bar0 + 8 * cur);
                gru_st_local4(ds_addr + off + 4 * VP, m1, bar0 + 8 * cur);
                gru_st_local4(ds_addr + off + 8 * VP, m2, bar0 + 8 * cur);
            } else
            for (int rk = 0; rk < NC; ++rk) {
                const uint32_t ra = gru_mapa(ds_addr + off, (uint32_t)rk), rb = gru_mapa(bar0 + 8 * cur, (uint32_t)rk);
                gru_st_async4(ra, m0, rb);
                gru_st_async4(ra + 4 * VP, m1, rb);
                gru_st_async4(ra + 8 * VP, m2, rb);
            }
        }
        gru_mbar_wait(bar0 + 8 * cur, (uint32_t)(s >> 1) & 1u);
        if (tid == 0 && s + 2 <= T - 2) gru_mbar_expect(bar0 + 8 * cur, tx_bytes);
        float acc = 0.f;
        for (int b = 0; b < nb; ++b) {
            const float4* dv = reinterpret_cast<const float4*>(ds_dyn + ((cur * BG + b) * 3) * VP) + kp;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
            for (int i = 0; i < KC; ++i) {
                const float4 d0 = dv[8 * i], d1 = dv[8 * i + VP / 4], d2 = dv[8 * i + 2 * (VP / 4)];
                a0 = fmaf(wt[0][4 * i], d0.x, a0); a0 = fmaf(wt[0][4 * i + 1], d0.y, a0);
                a0 = fmaf(wt[0][4 * i + 2], d0.z, a0); a0 = fmaf(wt[0][4 * i + 3], d0.w, a0);
                a1 = fmaf(wt[1][4 * i], d1.x, a1); a1 = fmaf(wt[1][4 * i + 1], d1.y, a1);
                a1 = fmaf(wt[1][4 * i + 2], d1.z, a1); a1 = fmaf(wt[1][4 * i + 3], d1.w, a1);
                a2 = fmaf(wt[2][4 * i], d2.x, a2); a2 = fmaf(wt[2][4 * i + 1], d2.y, a2);
                a2 = fmaf(wt[2][4 * i + 2], d2.z, a2); a2 = fmaf(wt[2][4 * i + 3], d2.w, a2);
            }
            float sum = a0 + a1 + a2;
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            if (kp == b) acc = sum;
        }
        dh_rec = dh_direct + acc;
    }
    cluster.sync();
}

template <typename K, typename... Args>
static int launch_cluster(K kernel, int nblocks, int NC, size_t smem, cudaStream_t st, Args... args) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        set_error("gru: smem attribute (%zu B): %s", smem, cudaGetErrorString(e));
        return M2D_ERR_CUDA;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)nblocks);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)NC;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    e = cudaLaunchKernelEx(&cfg, kernel, args...);
    if (e != cudaSuccess) {
        set_error("gru: launch: %s", cudaGetErrorString(e));
        return M2D_ERR_CUDA;
    }
    return M2D_OK;
}

}  // namespace m2d

using namespace m2d;

struct Gru2Plan { int NC, nu, BG, KC; };
static Gru2Plan gru2_plan(int B, int H) {
    Gru2Plan p;
    p.NC = H <= 32 ? 1 : (H <= 64 ? 2 : (H <= 128 ? 4 : 8));
    p.nu = (((H + p.NC - 1) / p.NC) + 3) & ~3;       // units per CTA, multiple of 4 (<= 32)
    p.BG = B <= 18 ? 1 : (B + 17) / 18;
    if (p.BG > GRU2_BGMAX) p.BG = GRU2_BGMAX;
    p.KC = (H + 31) / 32;
    return p;
}

template <int KC>
static int launch_gru2(const float* gi, const float* w_hh, const float* b_hh, float* h_out, int ldh, float* save,
                       int B, int T, int H, const Gru2Plan& p, cudaStream_t st) {
    int groups = (B + p.BG - 1) / p.BG;
    if (save)
        return launch_cluster(gru_fwd2_kernel<KC, true>, groups * p.NC, p.NC, 0, st, gi, w_hh, b_hh, h_out, ldh, save,
                              B, T, H, p.nu, p.BG);
    return launch_cluster(gru_fwd2_kernel<KC, false>, groups * p.NC, p.NC, 0, st, gi, w_hh, b_hh, h_out, ldh, save,
                          B, T, H, p.nu, p.BG);
}

template <int KC>
static int launch_gru2_bwd(const float* dh_out, int ldd, const float* h_out, int ldh, const float* save,
                           const float* w_hh, float* dgi, float* dgh, int B, int T, int H, const Gru2Plan& p,
                           cudaStream_t st) {
    int groups = (B + p.BG - 1) / p.BG;
    size_t smem = (size_t)2 * p.BG * 3 * 32 * KC * sizeof(float);
    return launch_cluster(gru_bwd2_kernel<KC>, groups * p.NC, p.NC, smem, st, dh_out, ldd, h_out, ldh, save, w_hh,
                          dgi, dgh, B, T, H, p.nu, p.BG);
}

namespace m2d { int g_gru_impl = 2; int g_gru_fwd_bg = 0; }
extern "C" int m2d_set_gru_impl(int v) { m2d::g_gru_impl = v; return M2D_OK; }
extern "C" int m2d_set_gru_forward_batch_group(int bg) {
    if (bg < 0 || bg > GRU2_BGMAX) {
        m2d::set_error("set_gru_forward_batch_group: 0 (automatic) .. %d", GRU2_BGMAX);
        return M2D_ERR_BAD_ARG;
    }
    m2d::g_gru_fwd_bg = bg;
    return M2D_OK;
}

extern "C" int m2d_gru_forward(const float* gi, const float* w_hh, const float* b_hh, float* h_out,
                               int ldh, float* save, int B, int T, int H, void* stream) {
    M2D_REQUIRE(gi && w_hh && b_hh && h_out && B > 0 && T > 0 && H > 0 && ldh >= H, "gru_forward: bad args");
    if (m2d::g_gru_impl == 2 && H <= 256) {
        Gru2Plan p = gru2_plan(B, H);
        if (m2d::g_gru_fwd_bg > 0) p.BG = m2d::g_gru_fwd_bg < B ? m2d::g_gru_fwd_bg : B;
        cudaStream_t st = (cudaStream_t)stream;
        switch (p.KC) {
            case 1: return launch_gru2<1>(gi, w_hh, b_hh, h_out, ldh, save, B, T, H, p, st);
            case 2: return launch_gru2<2>(gi, w_hh, b_hh, h_out, ldh, save, B, T, H, p, st);
            case 3: return launch_gru2<3>(gi, w_hh, b_hh, h_out, ldh, save, B, T, H, p, st);
            case 4: return launch_gru2<4>(gi, w_hh, b_hh, h_out, ldh, save, B, T, H, p, st);
            case 5: return launch_gru2<5>(gi, w_hh, b_hh, h_out, ldh, save, B, T, H, p, st);
            case 6: return launch_gru2<6>(gi, w_hh, b_hh, h_out, ldh, save, B, T, H, p, st);
            case 7: return launch_gru2<7>(gi, w_hh, b_hh, h_out, ldh, save, B, T, H, p, st);
            default: return launch_gru2<8>(gi, w_hh, b_hh, h_out, ldh, save, B, T, H, p, st);
        }
    }
    GruPlan p = make_plan(H);
    size_t smem = (size_t)(3 * p.nu * p.HP + 2 * p.BG * p.HP) * sizeof(float);
    M2D_REQUIRE(smem <= 220 * 1024, "gru_forward: hidden size %d needs %zu B of shared memory", H, smem);
    int groups = (B + p.BG - 1) / p.BG;
    if (save)
        return launch_cluster(gru_fwd_kernel<true>, groups * p.NC, p.NC, smem, (cudaStream_t)stream, gi,
                              w_hh, b_hh, h_out, ldh, save, B, T, H, p.nu, p.BG, p.HP);
    return launch_cluster(gru_fwd_kernel<false>, groups * p.NC, p.NC, smem, (cudaStream_t)stream, gi,
                          w_hh, b_hh, h_out, ldh, save, B, T, H, p.nu, p.BG, p.HP);
}

extern "C" int m2d_gru_backward(const float* dh_out, int ldd, const float* h_out, int ldh,
                                const float* save, const float* w_hh, float* dgi, float* dgh, int B,
                                int T, int H, void* stream) {
    M2D_REQUIRE(dh_out && h_out && save && w_hh && dgi && dgh && B > 0 && T > 0 && H > 0,
                "gru_backward: bad args");
    if (m2d::g_gru_impl == 2 && H <= 256) {
        const Gru2Plan p = gru2_plan(B, H);
        cudaStream_t st = (cudaStream_t)stream;
        switch (p.KC) {
            case 1: return launch_gru2_bwd<1>(dh_out, ldd, h_out, ldh, save, w_hh, dgi, dgh, B, T, H, p, st);
            case 2: return launch_gru2_bwd<2>(dh_out, ldd, h_out, ldh, save, w_hh, dgi, dgh, B, T, H, p, st);
            case 3: return launch_gru2_bwd<3>(dh_out, ldd, h_out, ldh, save, w_hh, dgi, dgh, B, T, H, p, st);
            case 4: return launch_gru2_bwd<4>(dh_out, ldd, h_out, ldh, save, w_hh, dgi, dgh, B, T, H, p, st);
            case 5: return launch_gru2_bwd<5>(dh_out, ldd, h_out, ldh, save, w_hh, dgi, dgh, B, T, H, p, st);
            case 6: return launch_gru2_bwd<6>(dh_out, ldd, h_out, ldh, save, w_hh, dgi, dgh, B, T, H, p, st);
            case 7: return launch_gru2_bwd<7>(dh_out, ldd, h_out, ldh, save, w_hh, dgi, dgh, B, T, H, p, st);
            default: return launch_gru2_bwd<8>(dh_out, ldd, h_out, ldh, save, w_hh, dgi, dgh, B, T, H, p, st);
        }
    }
    GruPlan p = make_plan(H);
    size_t smem = (size_t)(p.nu * p.HP3 + 2 * p.BG * p.HP3) * sizeof(float);
    M2D_REQUIRE(smem <= 220 * 1024, "gru_backward: hidden size %d needs %zu B of shared memory", H, smem);
    int groups = (B + p.BG - 1) / p.BG;
    return launch_cluster(gru_bwd_kernel, groups * p.NC, p.NC, smem, (cudaStream_t)stream, dh_out, ldd,
                          h_out, ldh, save, w_hh, dgi, dgh, B, T, H, p.nu, p.BG, p.HP3);
}
