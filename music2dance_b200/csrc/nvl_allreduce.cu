// In-place sum all-reduce of a float buffer over the GPUs of one NVSwitch domain, written against peer memory: no
// NCCL call, one kernel, graph-capturable like any other launch (m2d_nvl_allreduce, include/m2d.h).
//
// Two-shot: rank r owns slice r of the buffer.  Every thread block first meets its same-index blocks on all peers
// (release / acquire flags in the peers' signal pads: "my gradients are complete"), then reduces its part of the
// rank's slice over all GPUs — with NVLink SHARP when a multicast mapping exists (multimem.ld_reduce: the switch
// adds the eight copies, one 16-byte load per 4 results) or by peer loads in rank order otherwise — and broadcasts
// the sums to every GPU (multimem.st / peer stores), then meets the peers again ("every slice is written").
// Bytes over NVLink per GPU: 2 * (W-1)/W * n * 4, the all-reduce minimum.
#include "common.cuh"

namespace m2d {

__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ uint32_t cas_release_sys(uint32_t* addr, uint32_t cmp, uint32_t val) {
    uint32_t old;
    asm volatile("atom.global.release.sys.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(addr), "r"(cmp), "r"(val) : "memory");
    return old;
}
__device__ __forceinline__ uint32_t cas_acquire_sys(uint32_t* addr, uint32_t cmp, uint32_t val) {
    uint32_t old;
    asm volatile("atom.global.acquire.sys.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(addr), "r"(cmp), "r"(val) : "memory");
    return old;
}
// Meet block `blockIdx.x` of every peer.  Slot (block, sender) of a GPU's signal pad is raised 0 -> 1 by the sender and
// lowered 1 -> 0 by the owner, so the pads need no epoch and return to zero.  A peer that never arrives (a bug or a
// dead rank) must not hang the GPU: after `timeout_ns` the block records the failure in *status and moves on.
__device__ __forceinline__ void meet_peers(uint32_t* const* pads, int rank, int world, int slot0, int* status,
                                           unsigned long long timeout_ns) {
    __syncthreads();
    if ((int)threadIdx.x < world) {
        const int peer = threadIdx.x;
        uint32_t* put = pads[peer] + slot0 + (int)blockIdx.x * world + rank;
        uint32_t* wait = pads[rank] + slot0 + (int)blockIdx.x * world + peer;
        const unsigned long long t0 = gtimer();
        bool ok = true;
        while (cas_release_sys(put, 0u, 1u) != 0u)
            if (gtimer() - t0 > timeout_ns) { ok = false; break; }
        while (ok && cas_acquire_sys(wait, 1u, 0u) != 1u)
            if (gtimer() - t0 > timeout_ns) { ok = false; break; }
        if (!ok) atomicExch(status, 1);
    }
    __syncthreads();
}

__device__ __forceinline__ float4 mc_ld_reduce(const float* mc) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
    return v;
}
__device__ __forceinline__ void mc_st(float* mc, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

constexpr int NVL_THREADS = 512;
constexpr int NVL_MAX_WORLD = 16;

struct NvlPtrs {
    float* buf[NVL_MAX_WORLD];           // segment 0: every rank's mapping of the buffer
    float* buf2[NVL_MAX_WORLD];          // optional segment 1 (another symmetric buffer), same handshakes
    uint32_t* pad[NVL_MAX_WORLD];
};

// reduce + broadcast this block's share of rank `rank`'s slice of one segment
__device__ __forceinline__ void nvl_segment(float* const* buf, float* mc, int rank, int world, long long off, long long n4,
                                            bool unroll) {
    const long long per_rank = (n4 + world - 1) / world;
    const long long r0 = rank * per_rank, r1 = r0 + per_rank < n4 ? r0 + per_rank : n4;
    const long long base4 = off >> 2;
    // four independent 16-byte transactions in flight per thread: one ld_reduce -> st round trip over the switch takes
    // a few microseconds, so the kernel is bound by bytes in flight, not by the links (one vector per thread and 32
    // blocks = 256 KiB in flight moved ~150 GB/s)
    const long long stride = (long long)gridDim.x * NVL_THREADS;
    long long i = r0 + (long long)blockIdx.x * NVL_THREADS + threadIdx.x;
    if (mc) {
        float4* m4 = reinterpret_cast<float4*>(mc) + base4;
        for (; unroll && i + 3 * stride < r1; i += 4 * stride) {
            const float4 v0 = mc_ld_reduce(reinterpret_cast<const float*>(m4 + i));
            const float4 v1 = mc_ld_reduce(reinterpret_cast<const float*>(m4 + i + stride));
            const float4 v2 = mc_ld_reduce(reinterpret_cast<const float*>(m4 + i + 2 * stride));
            const float4 v3 = mc_ld_reduce(reinterpret_cast<const float*>(m4 + i + 3 * stride));
            mc_st(reinterpret_cast<float*>(m4 + i), v0);
            mc_st(reinterpret_cast<float*>(m4 + i + stride), v1);
            mc_st(reinterpret_cast<float*>(m4 + i + 2 * stride), v2);
            mc_st(reinterpret_cast<float*>(m4 + i + 3 * stride), v3);
        }
        for (; i < r1; i += stride) {
            const float4 v = mc_ld_reduce(reinterpret_cast<const float*>(m4 + i));
            mc_st(reinterpret_cast<float*>(m4 + i), v);
        }
    } else {
        for (; unroll && i + stride < r1; i += 2 * stride) {
            float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
            for (int p = 0; p < world; ++p) {                    // fixed order: identical sums on every rank
                const float4 a = __ldcg(reinterpret_cast<const float4*>(buf[p]) + base4 + i);
                const float4 b = __ldcg(reinterpret_cast<const float4*>(buf[p]) + base4 + i + stride);
                s0.x += a.x; s0.y += a.y; s0.z += a.z; s0.w += a.w;
                s1.x += b.x; s1.y += b.y; s1.z += b.z; s1.w += b.w;
            }
            for (int p = 0; p < world; ++p) {
                __stcg(reinterpret_cast<float4*>(buf[p]) + base4 + i, s0);
                __stcg(reinterpret_cast<float4*>(buf[p]) + base4 + i + stride, s1);
            }
        }
        for (; i < r1; i += stride) {
            float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int p = 0; p < world; ++p) {
                const float4 v = __ldcg(reinterpret_cast<const float4*>(buf[p]) + base4 + i);
                s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            }
            for (int p = 0; p < world; ++p) __stcg(reinterpret_cast<float4*>(buf[p]) + base4 + i, s);
        }
    }
}

__global__ void __launch_bounds__(NVL_THREADS)
nvl_allreduce_kernel(const NvlPtrs P, float* mc, const int rank, const int world, const long long off,
                     const long long n4 /* 16-byte vectors */, float* mc2, const long long off2, const long long n4_2,
                     const int slot0, int* status, const unsigned long long timeout_ns, const int unroll) {
    meet_peers(P.pad, rank, world, slot0, status, timeout_ns);
    nvl_segment(P.buf, mc, rank, world, off, n4, unroll != 0);
    if (n4_2 > 0) nvl_segment(P.buf2, mc2, rank, world, off2, n4_2, unroll != 0);
    __threadfence_system();
    meet_peers(P.pad, rank, world, slot0, status, timeout_ns);
}

}  // namespace m2d

using namespace m2d;

static int nvl_launch(float* const* bufs, float* mc, long long off, long long n, float* const* bufs2, float* mc2,
                      long long off2, long long n2, unsigned int* const* signal_pads, int rank, int world, int blocks,
                      int slot0, int* status, void* stream) {
    M2D_REQUIRE(bufs && signal_pads && status && world >= 2 && world <= NVL_MAX_WORLD && rank >= 0 && rank < world,
                "nvl_allreduce: bad args");
    M2D_REQUIRE(n > 0 && (n & 3) == 0 && (off & 3) == 0 && blocks > 0 && blocks <= 64 && slot0 >= 0,
                "nvl_allreduce: n and off must be multiples of 4 floats, 1 <= blocks <= 64");
    M2D_REQUIRE(n2 == 0 || (bufs2 && n2 > 0 && (n2 & 3) == 0 && (off2 & 3) == 0), "nvl_allreduce: bad second segment");
    NvlPtrs P;
    for (int r = 0; r < world; ++r) {
        M2D_REQUIRE(bufs[r] && signal_pads[r] && aligned16(bufs[r]), "nvl_allreduce: null / unaligned peer pointer");
        P.buf[r] = bufs[r];
        P.buf2[r] = n2 ? bufs2[r] : nullptr;
        M2D_REQUIRE(!n2 || (bufs2[r] && aligned16(bufs2[r])), "nvl_allreduce: null / unaligned peer pointer (segment 1)");
        P.pad[r] = signal_pads[r];
    }
    M2D_REQUIRE((!mc || aligned16(mc)) && (!mc2 || aligned16(mc2)), "nvl_allreduce: unaligned multicast pointer");
    static const int unroll = getenv("M2D_NVL_UNROLL") ? atoi(getenv("M2D_NVL_UNROLL")) : 1;   // 0: one vector in flight per thread
    nvl_allreduce_kernel<<<blocks, NVL_THREADS, 0, (cudaStream_t)stream>>>(P, mc, rank, world, off, n >> 2, mc2, off2,
                                                                          n2 >> 2, slot0, status, 2000000000ull /* 2 s */,
                                                                          unroll);
    return check_launch("nvl_allreduce");
}

extern "C" int m2d_nvl_allreduce(float* const* bufs, float* mc, unsigned int* const* signal_pads, int rank, int world,
                                 long long off, long long n, int blocks, int slot0, int* status, void* stream) {
    return nvl_launch(bufs, mc, off, n, nullptr, nullptr, 0, 0, signal_pads, rank, world, blocks, slot0, status, stream);
}

extern "C" int m2d_nvl_allreduce2(float* const* bufs, float* mc, long long off, long long n, float* const* bufs2,
                                  float* mc2, long long off2, long long n2, unsigned int* const* signal_pads, int rank,
                                  int world, int blocks, int slot0, int* status, void* stream) {
    return nvl_launch(bufs, mc, off, n, bufs2, mc2, off2, n2, signal_pads, rank, world, blocks, slot0, status, stream);
}
