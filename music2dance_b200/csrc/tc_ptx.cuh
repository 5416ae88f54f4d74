// tcgen05 / TMA / mbarrier / cluster PTX wrappers and tile constants shared by the tensor-core kernels
// (rowconv_tc.cu: SIMT-staged operands; rowconv_halo.cu: TMA halo tiles).
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace m2d {

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;
constexpr int TC_BNMAX = 128;
constexpr int TC_PRODUCERS = 512;       // 16 producer / epilogue warps
constexpr int TC_PW = TC_PRODUCERS / 32;   // index of the MMA-issuer warp; the TMA issuer is TC_PW + 1
constexpr int TC_RPT = 128 * 8 / TC_PRODUCERS;   // 16-byte chunks of a 128-row x 128-byte tile per producer thread
constexpr int TC_THREADS = TC_PRODUCERS + 64;   // + MMA-issuer warp + TMA-issuer warp
constexpr int TC_A_BYTES = TC_BM * TC_BK * 4;        // 16 KiB
constexpr int TC_B_BYTES = TC_BNMAX * TC_BK * 4;     // 16 KiB

// NS = operand-split scheme of a kernel instance: 1 = one TF32 product (hi plane only), 3 = 3xTF32 (hi + TF32 lo
// plane), 2 = TF32 hi*hi + ONE BF16 contraction for both cross terms (hi plane + a BF16 plane of the same size:
// [bf16(hi) | bf16(lo)] per 128-byte row for the activation operand, [bf16(lo) | bf16(hi)] for the weights)
__host__ __device__ constexpr int tc_planes(int ns) { return ns >= 2 ? 2 : 1; }
__host__ __device__ constexpr int tc_stage_bytes(int ns) { return tc_planes(ns) * (TC_A_BYTES + TC_B_BYTES); }
__host__ __device__ constexpr int tc_stages(int ns) { return ns >= 2 ? 3 : 4; }
// stages + epilogue staging tile never coexist: the C tile (128 x 129 floats) reuses the stages
__host__ __device__ constexpr int tc_smem_bytes(int ns) { return tc_stages(ns) * tc_stage_bytes(ns) + 1024; }

// ---------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// Spin on the barrier; a pipeline bug must surface as a launch failure, never as a hung GPU:
// after 2^22 failed polls (each poll suspends up to the hardware time limit: seconds) the kernel traps.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try(bar, parity)) {
        if (++spins > (1u << 22)) __trap();
    }
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], kind::tf32, issued by ONE thread
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// same, kind::f16 (BF16 operands, K = 16 per instruction, fp32 accumulate into the same TMEM columns)
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive on an mbarrier once every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ bool aligned16d(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
// round to TF32 (10 explicit mantissa bits), nearest with ties away from zero — the result of
// cvt.rna.tf32.f32, computed on the integer pipe (the conversion pipe is a quarter-rate unit)
__device__ __forceinline__ float to_tf32(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}

// K-major operand tile, SWIZZLE_128B: row r (128 bytes = 32 floats) lives at
// (r/8)*1024 + (r%8)*128, its 16-byte chunk j at chunk position j ^ (r%8).
__device__ __forceinline__ uint32_t sw128_off(int r, int j) {
    return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((j ^ (r & 7)) << 4));
}
// shared-memory matrix descriptor (tcgen05): start >> 4 | LBO(16 B, unused for swizzled K-major) |
// SBO = 1024 B between 8-row groups | version 1 | layout SWIZZLE_128B (2)
__device__ __forceinline__ uint64_t sw128_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// instruction descriptor: D = F32 (1 @ bit 4), A = B = TF32 (2 @ bits 7, 10), both K-major,
// N >> 3 @ bit 17, M >> 4 @ bit 24
__device__ __forceinline__ uint32_t tf32_idesc(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// kind::f16 with D = F32, A = B = BF16 (format 1), both K-major
__device__ __forceinline__ uint32_t bf16_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// 4 consecutive channels (16-byte chunk j of row r of an fp32 K-major tile) -> the BF16 plane of the mixed
// scheme: 4 x bf16(first) into row positions 4j.., 4 x bf16(second) into 32 + 4j.. (8 bytes each), same swizzle.
// Activation operand: first = hi, second = lo; weight operand: first = lo, second = hi.
__device__ __forceinline__ void store_bf16_pair(uint8_t* plane, int r, int j, const float4 first, const float4 second) {
    const int sw = r & 7;
    uint8_t* row = plane + (r >> 3) * 1024 + sw * 128 + 8 * (j & 1);
    __nv_bfloat162 f0 = __floats2bfloat162_rn(first.x, first.y), f1 = __floats2bfloat162_rn(first.z, first.w);
    __nv_bfloat162 s0 = __floats2bfloat162_rn(second.x, second.y), s1 = __floats2bfloat162_rn(second.z, second.w);
    uint2 fv, sv;
    fv.x = *reinterpret_cast<uint32_t*>(&f0); fv.y = *reinterpret_cast<uint32_t*>(&f1);
    sv.x = *reinterpret_cast<uint32_t*>(&s0); sv.y = *reinterpret_cast<uint32_t*>(&s1);
    *reinterpret_cast<uint2*>(row + ((((j >> 1)) ^ sw) << 4)) = fv;
    *reinterpret_cast<uint2*>(row + (((4 + (j >> 1)) ^ sw) << 4)) = sv;
}

constexpr int TC_CLD = TC_BNMAX + 4;      // C staging tile row stride (floats): 16-byte aligned rows, conflict-free float4 access

// TMEM accumulator (128 lanes x bn columns) -> shared C tile.  Warp w reads lanes 32*(w%4)..+31
// (the tcgen05.ld lane-quarter rule); the 16-column chunks are dealt round-robin to the warps of a quarter.
__device__ __forceinline__ void tmem_to_smem(uint32_t tmem, float* Cs, int warp, int lane, int bn) {
    const int q = warp & 3, part = warp >> 2;
    const int row = 32 * q + lane;
    const int chunks = bn / 16;
    for (int ch = part; ch < chunks; ch += TC_PW / 4) {
        float v[16];
        tmem_ld16(tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(16 * ch), v);
        float4* dst = reinterpret_cast<float4*>(Cs + row * TC_CLD + 16 * ch);
#pragma unroll
        for (int u = 0; u < 4; ++u) dst[u] = make_float4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
    }
}
// named barrier over the 256 producer / epilogue threads (warp 8 does not take part)
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// Split-K over a thread-block cluster: the gridDim.z CTAs of one (m,n) tile form a cluster (1,1,Z).
// Each keeps its partial C tile in shared memory; after a cluster barrier CTA `rank` sums rows
// [rank*128/Z, (rank+1)*128/Z) over all Z tiles through distributed shared memory (fixed order:
// deterministic), applies the epilogue and stores.  No global workspace, no second launch.
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ float4 dsmem_ld4(uint32_t local_addr, uint32_t rank) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_addr), "r"(rank));
    float4 v;
    asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(ra) : "memory");
    return v;
}
// sum of the 4-float chunk at (row, c4) of the C tiles of all Z CTAs of the cluster
__device__ __forceinline__ void cluster_reduce4(const float* Cs, int rl, int c4, int Z, float* v) {
    const uint32_t addr = smem_u32(Cs + rl * TC_CLD + c4);
    v[0] = v[1] = v[2] = v[3] = 0.f;
    for (int z = 0; z < Z; ++z) {
        const float4 t = dsmem_ld4(addr, (uint32_t)z);
        v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w;
    }
}

// TMA: one box of [128 weight rows][32 floats] lands in shared memory already in the SWIZZLE_128B
// K-major layout the tensor core reads; completion is signalled on the stage's mbarrier.
__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
// One elected lane of a CONVERGED warp.  The issuer warps keep their control flow warp-uniform and predicate only
// the tcgen05 / TMA instructions with this: inside an `if (lane == 0)` region the compiler must assume divergent
// uniform-register operands and wraps every UTCHMMA / UTMALDG in an ELECT / BRA.U.ANY serialisation loop
// (seen in the SASS of the first version: ~100 issue slots per MMA on the single issuing thread).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(pred));
    return pred != 0;
}
// bulk copy of a contiguous global block (16-byte aligned, multiple of 16 bytes) into shared memory;
// completion is signalled on the mbarrier like a TMA tensor load
__device__ __forceinline__ void bulk_load(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

}  // namespace m2d
