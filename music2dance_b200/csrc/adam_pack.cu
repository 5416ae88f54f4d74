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

// ---- fast index helpers (tile-local ranges: j < 2^16, divisors < 2^10 -> the float reciprocal is exact) ----------
__device__ __forceinline__ int fdiv(int j, float inv) { return (int)(((float)j + 0.5f) * inv); }
__device__ __forceinline__ long long tiled_fast(int n, int t, int ci, int T, int cch, int lgR) {
    const int R = 1 << lgR;
    const int nt_ = n >> lgR, nl = n & (R - 1), c = ci >> 5, kk = ci & 31;
    const long long blk = ((long long)nt_ * T + t) * cch + c;
    return blk * (64LL << lgR) + nl * 32 + ((((kk >> 2) ^ (nl & 7)) << 2) | (kk & 3));
}
struct TapInfo {            // per tap of the tile and pack output: the stride / residue arithmetic of pack_index
    int rowbase, q, T;      // MERGED: r0*Cin, q', Tm;  BWD: 0, q, Trho
    long long off, poff;    // BWD: offsets of the residue's block (plain / tiled copy)
};

// 4 consecutive contraction-channel values (kk0 = 4*chunk .. +3 of the block row `nl`) of the tiled copy: one 16-byte
// chunk of the hi plane and one of the lo plane (3xTF32 split; the BF16 cross-term format keeps the scalar path)
__device__ __forceinline__ void store_tiled_split4(float* dst_tiled, long long blk_base, int nl, int chunk, int R, float4 v) {
    float* row = dst_tiled + blk_base + nl * 32 + ((chunk ^ (nl & 7)) << 2);
    const float4 h = make_float4(rna_tf32(v.x), rna_tf32(v.y), rna_tf32(v.z), rna_tf32(v.w));
    *reinterpret_cast<float4*>(row) = h;
    *reinterpret_cast<float4*>(row + R * 32) =
        make_float4(rna_tf32(v.x - h.x), rna_tf32(v.y - h.y), rna_tf32(v.z - h.z), rna_tf32(v.w - h.w));
}
// base (floats) of the tiled block holding row n, tap t, channel chunk c of an operand with T taps and cch chunks
__device__ __forceinline__ long long tiled_blk(int n, int t, int c, int T, int cch, int lgR) {
    return (((long long)(n >> lgR) * T + t) * cch + c) * (64LL << lgR);
}

constexpr int AP_MAXT = 32;       // taps per tile

// AP_THREADS = 512 with tiles of up to ~100 KiB (2 CTAs / SM) or 256 with tiles of up to ~50 KiB (4 CTAs / SM: same
// number of warps per SM, finer-grained work items)
template <int AP_THREADS>
__global__ void __launch_bounds__(AP_THREADS, 1024 / AP_THREADS)
adam_pack_kernel(const m2d_adam_item* __restrict__ items, int* counters, float lr, float b1, float b2, float eps,
                 float gscale, const bool mixed) {
    extern __shared__ float tile[];
    __shared__ TapInfo taps[3][AP_MAXT];
    const m2d_adam_item& it = items[blockIdx.x];
    const int t_step = counters[0] + 1;
    const AdamConst c = adam_const(t_step, lr, b1, b2, eps, gscale);
    const int tid = threadIdx.x;
    if (it.flat_n > 0) {
        // plain range: 16-byte vectors, two vectors per thread in flight
        float4* __restrict__ p = reinterpret_cast<float4*>(it.p);
        float4* __restrict__ m = reinterpret_cast<float4*>(it.m);
        float4* __restrict__ v = reinterpret_cast<float4*>(it.v);
        const float4* __restrict__ g = reinterpret_cast<const float4*>(it.g);
        const int n4 = (int)(it.flat_n >> 2);
        for (int i = tid; i < n4; i += 2 * AP_THREADS) {
            const int i2 = i + AP_THREADS;
            const bool two = i2 < n4;
            float4 P0 = p[i], M0 = m[i], V0 = v[i], G0 = g[i];
            float4 P1 = P0, M1 = M0, V1 = V0, G1 = G0;
            if (two) { P1 = p[i2]; M1 = m[i2]; V1 = v[i2]; G1 = g[i2]; }
            P0.x = adam_update(P0.x, G0.x, M0.x, V0.x, c); P0.y = adam_update(P0.y, G0.y, M0.y, V0.y, c);
            P0.z = adam_update(P0.z, G0.z, M0.z, V0.z, c); P0.w = adam_update(P0.w, G0.w, M0.w, V0.w, c);
            p[i] = P0; m[i] = M0; v[i] = V0;
            if (two) {
                P1.x = adam_update(P1.x, G1.x, M1.x, V1.x, c); P1.y = adam_update(P1.y, G1.y, M1.y, V1.y, c);
                P1.z = adam_update(P1.z, G1.z, M1.z, V1.z, c); P1.w = adam_update(P1.w, G1.w, M1.w, V1.w, c);
                p[i2] = P1; m[i2] = M1; v[i2] = V1;
            }
        }
        for (long long i = ((long long)n4 << 2) + tid; i < it.flat_n; i += AP_THREADS) {      // tail (< 4 floats)
            float mi = it.m[i], vi = it.v[i];
            it.p[i] = adam_update(it.p[i], it.g[i], mi, vi, c);
            it.m[i] = mi;
            it.v[i] = vi;
        }
    } else {
        const int Cout = it.Cout, Cin = it.Cin, k = it.k;
        const int nco = it.nco, nci = it.nci, nt = it.nt, co0 = it.co0, ci0 = it.ci0, t0 = it.t0;
        const int L = nci * nt;                          // one output row's slice, parameter order (ci, t)
        const int Lp = L | 1;                            // odd row pitch: conflict-free column reads
        const int total = nco * L;
        const float invL = 1.0f / (float)L, inv_nci = 1.0f / (float)nci, inv_nt = 1.0f / (float)nt,
                    inv_nco = 1.0f / (float)nco;
        // per-tap constants of the strided backward layouts
        if (tid < 3 * AP_MAXT) {
            const int o = tid / AP_MAXT, tl = tid - o * AP_MAXT;
            if (o < it.n_pack && tl < nt) {
                const m2d_adam_pack_out& pk = it.pk[o];
                TapInfo ti;
                ti.rowbase = 0; ti.q = 0; ti.T = 0; ti.off = 0; ti.poff = 0;
                const int t = t0 + tl;
                if (pk.kind == M2D_PACK_BWD_MERGED) {
                    const PackGeom gm = pack_geom(pk, Cout, Cin, k);
                    const int rho = t % pk.stride, q = t / pk.stride;
                    int r0 = (rho - gm.pad) % pk.stride;
                    if (r0 < 0) r0 += pk.stride;
                    const int c0 = (r0 + gm.pad) / pk.stride;
                    ti.rowbase = r0 * Cin; ti.q = q + gm.cmax - c0; ti.T = gm.Tm;
                } else if (pk.kind == M2D_PACK_BWD) {
                    const int R = tiled_rows(Cin);
                    const int rho = t % pk.stride;
                    ti.q = t / pk.stride;
                    for (int r = 0; r <= rho; ++r) {
                        int Trho = (k - r + pk.stride - 1) / pk.stride;
                        if (Trho < 0) Trho = 0;
                        ti.T = Trho;
                        if (r < rho) {
                            ti.off += (long long)Cin * Cout * Trho;
                            ti.poff += tiled_blocks(Cin, Trho, Cout) * tiled_block_floats(R);
                        }
                    }
                }
                taps[o][tl] = ti;
            }
        }
        // 1. tap-major gradient -> tile[co][ci*nt + t]   (source order (co, t, ci), ci fastest; 4 loads in flight)
        const bool vec = it.pad_ != 0;                   // host: every extent / offset of the tile is a multiple of 4
        const int nci4 = nci >> 2, nco4 = nco >> 2;
        const float inv_nci4 = 1.0f / (float)(nci4 > 0 ? nci4 : 1), inv_nco4 = 1.0f / (float)(nco4 > 0 ? nco4 : 1);
        if (it.g_packed && vec) {
            const float* __restrict__ g = it.g;
            const int units = nco * nt * nci4, per_co = nt * nci4;
            const float inv_per = 1.0f / (float)per_co;
            for (int u0 = tid; u0 < units; u0 += 2 * AP_THREADS) {
                float4 val[2];
                int si[2];
#pragma unroll
                for (int w = 0; w < 2; ++w) {
                    const int u = u0 + w * AP_THREADS;
                    si[w] = -1;
                    if (u < units) {
                        const int cl = fdiv(u, inv_per), r = u - cl * per_co, tl = fdiv(r, inv_nci4), i4 = r - tl * nci4;
                        val[w] = *reinterpret_cast<const float4*>(g + ((long long)(co0 + cl) * k + t0 + tl) * Cin + ci0 + 4 * i4);
                        si[w] = cl * Lp + 4 * i4 * nt + tl;
                    }
                }
#pragma unroll
                for (int w = 0; w < 2; ++w)
                    if (si[w] >= 0) {
                        tile[si[w]] = val[w].x; tile[si[w] + nt] = val[w].y;
                        tile[si[w] + 2 * nt] = val[w].z; tile[si[w] + 3 * nt] = val[w].w;
                    }
            }
        } else if (it.g_packed) {
            const float* __restrict__ g = it.g;
            for (int e0 = tid; e0 < total; e0 += 4 * AP_THREADS) {
                float val[4];
                int si[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int e = e0 + u * AP_THREADS;
                    si[u] = -1;
                    if (e < total) {
                        const int cl = fdiv(e, invL), r = e - cl * L, tl = fdiv(r, inv_nci), il = r - tl * nci;
                        val[u] = g[((long long)(co0 + cl) * k + t0 + tl) * Cin + ci0 + il];
                        si[u] = cl * Lp + il * nt + tl;
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (si[u] >= 0) tile[si[u]] = val[u];
            }
        }
        __syncthreads();
        // 2. Adam in the parameter layout's order (co, ci, t); the new weight replaces the gradient in the tile
        {
            float* __restrict__ p = it.p;
            float* __restrict__ m = it.m;
            float* __restrict__ v = it.v;
            const float* __restrict__ g = it.g;
            const bool gp = it.g_packed != 0;
            const bool vec2 = vec && nt == k;            // a row's slice is one contiguous, 16-byte aligned run
            if (vec2) {
                const int L4 = L >> 2, units = nco * L4;
                const float invL4 = 1.0f / (float)L4;
                for (int u0 = tid; u0 < units; u0 += 2 * AP_THREADS) {
                    float4 P[2], M[2], V[2], G[2];
                    long long gi[2];
                    int si[2];
#pragma unroll
                    for (int w = 0; w < 2; ++w) {
                        const int u = u0 + w * AP_THREADS;
                        si[w] = -1;
                        if (u < units) {
                            const int cl = fdiv(u, invL4), r4 = u - cl * L4;
                            gi[w] = ((long long)(co0 + cl) * Cin + ci0) * k + 4 * r4;
                            si[w] = cl * Lp + 4 * r4;
                            P[w] = *reinterpret_cast<const float4*>(p + gi[w]);
                            M[w] = *reinterpret_cast<const float4*>(m + gi[w]);
                            V[w] = *reinterpret_cast<const float4*>(v + gi[w]);
                            if (gp) G[w] = make_float4(tile[si[w]], tile[si[w] + 1], tile[si[w] + 2], tile[si[w] + 3]);
                            else G[w] = *reinterpret_cast<const float4*>(g + gi[w]);
                        }
                    }
#pragma unroll
                    for (int w = 0; w < 2; ++w) {
                        if (si[w] >= 0) {
                            float4 N;
                            N.x = adam_update(P[w].x, G[w].x, M[w].x, V[w].x, c); N.y = adam_update(P[w].y, G[w].y, M[w].y, V[w].y, c);
                            N.z = adam_update(P[w].z, G[w].z, M[w].z, V[w].z, c); N.w = adam_update(P[w].w, G[w].w, M[w].w, V[w].w, c);
                            *reinterpret_cast<float4*>(p + gi[w]) = N;
                            *reinterpret_cast<float4*>(m + gi[w]) = M[w];
                            *reinterpret_cast<float4*>(v + gi[w]) = V[w];
                            tile[si[w]] = N.x; tile[si[w] + 1] = N.y; tile[si[w] + 2] = N.z; tile[si[w] + 3] = N.w;
                        }
                    }
                }
            } else
            for (int e0 = tid; e0 < total; e0 += 4 * AP_THREADS) {
                float P[4], M[4], V[4], G[4];
                long long gi[4];
                int si[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int e = e0 + u * AP_THREADS;
                    si[u] = -1;
                    if (e < total) {
                        const int cl = fdiv(e, invL), r = e - cl * L, il = fdiv(r, inv_nt), tl = r - il * nt;
                        gi[u] = ((long long)(co0 + cl) * Cin + ci0 + il) * k + t0 + tl;
                        si[u] = cl * Lp + r;
                        P[u] = p[gi[u]]; M[u] = m[gi[u]]; V[u] = v[gi[u]];
                        G[u] = gp ? tile[si[u]] : g[gi[u]];
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (si[u] >= 0) {
                        const float pn = adam_update(P[u], G[u], M[u], V[u], c);
                        p[gi[u]] = pn; m[gi[u]] = M[u]; v[gi[u]] = V[u];
                        tile[si[u]] = pn;
                    }
                }
            }
        }
        __syncthreads();
        // 3. re-layouts, each in its destination's contiguous order
        for (int o = 0; o < it.n_pack; ++o) {
            const m2d_adam_pack_out& pk = it.pk[o];
            const int kind = pk.kind;
            float* __restrict__ dst = pk.dst;
            float* __restrict__ dtl = pk.dst_tiled;
            if (kind == M2D_PACK_FWD && vec && !mixed) { // (co, t, ci): 4 channels per thread, 16-byte stores
                const int R = tiled_rows(Cout), lgR = R == 64 ? 6 : 7, cch = (Cin + 31) >> 5;
                const int units = nco * nt * nci4, per_co = nt * nci4;
                const float inv_per = 1.0f / (float)per_co;
                for (int u = tid; u < units; u += AP_THREADS) {
                    const int cl = fdiv(u, inv_per), r = u - cl * per_co, tl = fdiv(r, inv_nci4), i4 = r - tl * nci4;
                    const int s0 = cl * Lp + 4 * i4 * nt + tl;
                    const float4 val = make_float4(tile[s0], tile[s0 + nt], tile[s0 + 2 * nt], tile[s0 + 3 * nt]);
                    const int co = co0 + cl, ci = ci0 + 4 * i4, t = t0 + tl;
                    if (dst) *reinterpret_cast<float4*>(dst + ((long long)co * k + t) * Cin + ci) = val;
                    if (dtl) store_tiled_split4(dtl, tiled_blk(co, t, ci >> 5, k, cch, lgR), co & (R - 1), (ci >> 2) & 7, R, val);
                }
            } else if (kind != M2D_PACK_FWD && vec && !mixed) {   // (ci, t, co): 4 output channels per thread
                const int cch = (Cout + 31) >> 5;
                int R;
                if (kind == M2D_PACK_FULL_BWD) R = tiled_rows(k * Cin);
                else if (kind == M2D_PACK_BWD) R = tiled_rows(Cin);
                else R = tiled_rows(pk.stride * Cin);
                const int lgR = R == 64 ? 6 : 7;
                const int units = L * nco4;
                for (int u = tid; u < units; u += AP_THREADS) {
                    const int r = fdiv(u, inv_nco4), c4 = u - r * nco4, il = fdiv(r, inv_nt), tl = r - il * nt;
                    const int s0 = 4 * c4 * Lp + r;
                    const float4 val = make_float4(tile[s0], tile[s0 + Lp], tile[s0 + 2 * Lp], tile[s0 + 3 * Lp]);
                    const int co = co0 + 4 * c4, ci = ci0 + il, t = t0 + tl;
                    long long plain, blk;
                    int row;
                    if (kind == M2D_PACK_FULL_BWD) {
                        row = t * Cin + ci;
                        plain = (long long)row * Cout + co;
                        blk = tiled_blk(row, 0, co >> 5, 1, cch, lgR);
                    } else if (kind == M2D_PACK_BWD) {
                        const TapInfo ti = taps[o][tl];
                        row = ci;
                        plain = ti.off + ((long long)ci * ti.T + ti.q) * Cout + co;
                        blk = ti.poff + tiled_blk(row, ti.q, co >> 5, ti.T, cch, lgR);
                    } else {
                        const TapInfo ti = taps[o][tl];
                        row = ti.rowbase + ci;
                        plain = ((long long)row * ti.T + ti.q) * Cout + co;
                        blk = tiled_blk(row, ti.q, co >> 5, ti.T, cch, lgR);
                    }
                    if (dst) *reinterpret_cast<float4*>(dst + plain) = val;
                    if (dtl) store_tiled_split4(dtl, blk, row & (R - 1), (co >> 2) & 7, R, val);
                }
            } else if (kind == M2D_PACK_FWD) {           // (co, t, ci), ci fastest
                const int R = tiled_rows(Cout), lgR = R == 64 ? 6 : 7;
                const int cch = Cin == 1 ? (k + 31) >> 5 : (Cin + 31) >> 5;
                for (int e = tid; e < total; e += AP_THREADS) {
                    const int cl = fdiv(e, invL), r = e - cl * L, tl = fdiv(r, inv_nci), il = r - tl * nci;
                    const float val = tile[cl * Lp + il * nt + tl];
                    const int co = co0 + cl, ci = ci0 + il, t = t0 + tl;
                    if (dst) dst[((long long)co * k + t) * Cin + ci] = val;
                    if (dtl) {
                        const long long pidx = Cin == 1 ? tiled_fast(co, 0, t, 1, cch, lgR) : tiled_fast(co, t, ci, k, cch, lgR);
                        store_tiled_split(dtl, pidx, R, val, mixed);
                    }
                }
            } else {                                     // (ci, t, co), co fastest
                const int cch = (Cout + 31) >> 5;
                int R;
                if (kind == M2D_PACK_FULL_BWD) R = tiled_rows(k * Cin);
                else if (kind == M2D_PACK_BWD) R = tiled_rows(Cin);
                else R = tiled_rows(pk.stride * Cin);
                const int lgR = R == 64 ? 6 : 7;
                for (int e = tid; e < total; e += AP_THREADS) {
                    const int r = fdiv(e, inv_nco), cl = e - r * nco, il = fdiv(r, inv_nt), tl = r - il * nt;
                    const float val = tile[cl * Lp + r];
                    const int co = co0 + cl, ci = ci0 + il, t = t0 + tl;
                    long long plain, pidx;
                    if (kind == M2D_PACK_FULL_BWD) {
                        const int row = t * Cin + ci;
                        plain = (long long)row * Cout + co;
                        pidx = tiled_fast(row, 0, co, 1, cch, lgR);
                    } else if (kind == M2D_PACK_BWD) {
                        const TapInfo ti = taps[o][tl];
                        plain = ti.off + ((long long)ci * ti.T + ti.q) * Cout + co;
                        pidx = ti.poff + tiled_fast(ci, ti.q, co, ti.T, cch, lgR);
                    } else {
                        const TapInfo ti = taps[o][tl];
                        const int row = ti.rowbase + ci;
                        plain = ((long long)row * ti.T + ti.q) * Cout + co;
                        pidx = tiled_fast(row, ti.q, co, ti.T, cch, lgR);
                    }
                    if (dst) dst[plain] = val;
                    if (dtl) store_tiled_split(dtl, pidx, R, val, mixed);
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
    const bool small = smem <= 52 * 1024;
    static size_t configured[2] = {0, 0};
    if (smem > configured[small]) {
        cudaError_t e = small ? cudaFuncSetAttribute(adam_pack_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                              : cudaFuncSetAttribute(adam_pack_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("adam_pack: %s", cudaGetErrorString(e));
            return M2D_ERR_CUDA;
        }
        configured[small] = smem;
    }
    const bool mixed = gemm_mode() == M2D_GEMM_TF32_BF16;
    if (small)
        adam_pack_kernel<256><<<n, 256, smem, (cudaStream_t)stream>>>(items, counters, lr, beta1, beta2, eps, gscale, mixed);
    else
        adam_pack_kernel<512><<<n, 512, smem, (cudaStream_t)stream>>>(items, counters, lr, beta1, beta2, eps, gscale, mixed);
    return check_launch("adam_pack");
}
