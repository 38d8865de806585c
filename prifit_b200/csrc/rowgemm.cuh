// fp32 CUDA-core micro-kernels shared by the "few rows x many keys" kernels (K selected seeds,
// cluster centres, NMS row groups).  All assume a 256-thread CTA and shared-memory tiles with a
// row stride of D + 4 floats (16-byte aligned rows, conflict-free 128-bit accesses).
//
//   rows tile  : RG_ROWS (32) rows  x D     "ys"
//   keys tile  : RG_KEYS (128) keys x D     "xs"
//   S micro-tile per thread: rows  ty + 8a (a < 4), keys tx + 32c (c < 4); ty = warp, tx = lane
#pragma once
#include <cooperative_groups.h>
#include "common.cuh"

namespace cg = cooperative_groups;

constexpr int RG_THREADS = 256;
constexpr int RG_ROWS = 32;
constexpr int RG_KEYS = 128;
constexpr int RG_LDP = RG_KEYS + 4;   // row stride of a [32][128] coefficient tile

// load `nrows` rows of width D (global row r -> src(r), or zero when src < 0) into smem [nrows][D+4]
template <int D, typename SrcFn>
__device__ __forceinline__ void rg_load_rows(float* dst, int nrows, const float* __restrict__ base, SrcFn src) {
    constexpr int NV = D / 4;
    for (int e = threadIdx.x; e < nrows * NV; e += RG_THREADS) {
        const int r = e / NV, c = e - r * NV;
        const long long g = src(r);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (g >= 0) v = reinterpret_cast<const float4*>(base + (size_t)g * D)[c];
        *reinterpret_cast<float4*>(dst + r * (D + 4) + 4 * c) = v;
    }
}

// acc[a][c] = <ys[ty + 8a], xs[tx + 32c]>.  `nrows` = live rows of the rows tile (rows >= nrows are zero): row group a
// (rows 8a .. 8a + 7) is skipped when it holds no live row -- with K = 16 clusters in a 32-row tile that halves the work.
template <int D>
__device__ __forceinline__ void rg_dot_32x128(const float* __restrict__ ys, const float* __restrict__ xs, float (&acc)[4][4],
                                              int nrows = RG_ROWS) {
    constexpr int LD = D + 4;
    const int ty = threadIdx.x >> 5, tx = threadIdx.x & 31;
    const int na = (nrows + 7) >> 3;                  // live row groups (uniform over the CTA)
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][c] = 0.f;
#pragma unroll 4
    for (int i = 0; i < D; i += 4) {
        float4 x[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) x[c] = *reinterpret_cast<const float4*>(xs + (tx + 32 * c) * LD + i);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            if (a < na) {
                const float4 y = *reinterpret_cast<const float4*>(ys + (ty + 8 * a) * LD + i);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    acc[a][c] = fmaf(y.x, x[c].x, acc[a][c]);
                    acc[a][c] = fmaf(y.y, x[c].y, acc[a][c]);
                    acc[a][c] = fmaf(y.z, x[c].z, acc[a][c]);
                    acc[a][c] = fmaf(y.w, x[c].w, acc[a][c]);
                }
            }
        }
    }
}

// two row tiles against the same key tile in one sweep (shares the key loads)
template <int D>
__device__ __forceinline__ void rg_dot2_32x128(const float* __restrict__ y1s, const float* __restrict__ y2s,
                                               const float* __restrict__ xs, float (&a1)[4][4], float (&a2)[4][4]) {
    constexpr int LD = D + 4;
    const int ty = threadIdx.x >> 5, tx = threadIdx.x & 31;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) { a1[a][c] = 0.f; a2[a][c] = 0.f; }
#pragma unroll 2
    for (int i = 0; i < D; i += 4) {
        float4 y[4], g[4], x[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            y[a] = *reinterpret_cast<const float4*>(y1s + (ty + 8 * a) * LD + i);
            g[a] = *reinterpret_cast<const float4*>(y2s + (ty + 8 * a) * LD + i);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) x[c] = *reinterpret_cast<const float4*>(xs + (tx + 32 * c) * LD + i);
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                a1[a][c] = fmaf(y[a].x, x[c].x, a1[a][c]);
                a1[a][c] = fmaf(y[a].y, x[c].y, a1[a][c]);
                a1[a][c] = fmaf(y[a].z, x[c].z, a1[a][c]);
                a1[a][c] = fmaf(y[a].w, x[c].w, a1[a][c]);
                a2[a][c] = fmaf(g[a].x, x[c].x, a2[a][c]);
                a2[a][c] = fmaf(g[a].y, x[c].y, a2[a][c]);
                a2[a][c] = fmaf(g[a].z, x[c].z, a2[a][c]);
                a2[a][c] = fmaf(g[a].w, x[c].w, a2[a][c]);
            }
    }
}

// o[a][4h+q] += sum_key ps[ty + 8a][key] * xs[key][4 tx + 128 h + q]   (32 x D += [32 x 128] . [128 x D]); row groups
// without a live row (see rg_dot_32x128) are skipped
template <int D>
__device__ __forceinline__ void rg_accum_rows(const float* __restrict__ ps, const float* __restrict__ xs,
                                              float (&o)[4][4 * ((D + 127) / 128)], int nrows = RG_ROWS) {
    constexpr int LD = D + 4;
    constexpr int NH = (D + 127) / 128;
    const int ty = threadIdx.x >> 5, tx = threadIdx.x & 31;
    const int na = (nrows + 7) >> 3;
    for (int c0 = 0; c0 < RG_KEYS; c0 += 4) {
        float4 p[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) p[a] = a < na ? *reinterpret_cast<const float4*>(ps + (ty + 8 * a) * RG_LDP + c0) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
#pragma unroll
            for (int h = 0; h < NH; ++h) {
                if (4 * tx + 128 * h < D) {
                    const float4 x = *reinterpret_cast<const float4*>(xs + (c0 + cc) * LD + 4 * tx + 128 * h);
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        if (a < na) {
                            const float pv = cc == 0 ? p[a].x : cc == 1 ? p[a].y : cc == 2 ? p[a].z : p[a].w;
                            o[a][4 * h + 0] = fmaf(pv, x.x, o[a][4 * h + 0]);
                            o[a][4 * h + 1] = fmaf(pv, x.y, o[a][4 * h + 1]);
                            o[a][4 * h + 2] = fmaf(pv, x.z, o[a][4 * h + 2]);
                            o[a][4 * h + 3] = fmaf(pv, x.w, o[a][4 * h + 3]);
                        }
                    }
                }
            }
        }
    }
}

// Key-side accumulation: for the 128 keys of the tile,
//   gX[key][col] += sum_r  p1[r][key] * y1[r][col]  (+ p2[r][key] * y2[r][col] when TWO)
// thread micro-tile: keys 8 tj .. 8 tj + 7 (tj = tid / 16), columns 4 tc + 64 h (tc = tid % 16).
// The result is added to global memory (rows exclusively owned by this CTA), keys >= nvalid skipped.
template <int D, bool TWO>
__device__ __forceinline__ void rg_accum_keys(const float* __restrict__ p1, const float* __restrict__ y1,
                                              const float* __restrict__ p2, const float* __restrict__ y2,
                                              float* __restrict__ gX_tile /* &gX[key0][0] */, int nvalid, int nrows = RG_ROWS) {
    constexpr int LD = D + 4;
    constexpr int NH = D / 64;
    static_assert(D % 64 == 0, "D must be a multiple of 64");
    const int tj = threadIdx.x >> 4, tc = threadIdx.x & 15;
    float acc[8][4 * NH];
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int c = 0; c < 4 * NH; ++c) acc[k][c] = 0.f;
    for (int r = 0; r < nrows; ++r) {                  // rows >= nrows carry zero coefficients
        const float4 pa = *reinterpret_cast<const float4*>(p1 + r * RG_LDP + 8 * tj);
        const float4 pb = *reinterpret_cast<const float4*>(p1 + r * RG_LDP + 8 * tj + 4);
        const float pk[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
#pragma unroll
        for (int h = 0; h < NH; ++h) {
            const float4 y = *reinterpret_cast<const float4*>(y1 + r * LD + 4 * tc + 64 * h);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                acc[k][4 * h + 0] = fmaf(pk[k], y.x, acc[k][4 * h + 0]);
                acc[k][4 * h + 1] = fmaf(pk[k], y.y, acc[k][4 * h + 1]);
                acc[k][4 * h + 2] = fmaf(pk[k], y.z, acc[k][4 * h + 2]);
                acc[k][4 * h + 3] = fmaf(pk[k], y.w, acc[k][4 * h + 3]);
            }
        }
        if (TWO) {
            const float4 qa = *reinterpret_cast<const float4*>(p2 + r * RG_LDP + 8 * tj);
            const float4 qb = *reinterpret_cast<const float4*>(p2 + r * RG_LDP + 8 * tj + 4);
            const float qk[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
#pragma unroll
            for (int h = 0; h < NH; ++h) {
                const float4 y = *reinterpret_cast<const float4*>(y2 + r * LD + 4 * tc + 64 * h);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    acc[k][4 * h + 0] = fmaf(qk[k], y.x, acc[k][4 * h + 0]);
                    acc[k][4 * h + 1] = fmaf(qk[k], y.y, acc[k][4 * h + 1]);
                    acc[k][4 * h + 2] = fmaf(qk[k], y.z, acc[k][4 * h + 2]);
                    acc[k][4 * h + 3] = fmaf(qk[k], y.w, acc[k][4 * h + 3]);
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int key = 8 * tj + k;
        if (key < nvalid) {
#pragma unroll
            for (int h = 0; h < NH; ++h) {
                float4* g = reinterpret_cast<float4*>(gX_tile + (size_t)key * D + 4 * tc + 64 * h);
                float4 v = __ldcg(g);   // L2 only: other CTAs of the cluster may have updated this row
                v.x += acc[k][4 * h + 0]; v.y += acc[k][4 * h + 1]; v.z += acc[k][4 * h + 2]; v.w += acc[k][4 * h + 3];
                *g = v;
            }
        }
    }
}

// Cluster reduce-scatter + all-gather of a [32][D] partial tile held by every CTA of the cluster:
// CTA `rank` sums rows [rank*RPC, rank*RPC + RPC) over all ranks in fixed rank order (deterministic)
// and hands each owned (row, column-pair) to `fn(row, col, sum)`.  RPC = 32 / cluster size.
// Every CTA must have written its partial tile to `part` ([32][D], no padding) and be past a
// cluster.sync() before calling; the caller syncs the cluster again before `part` is rewritten.
template <int D, typename Fn>
__device__ __forceinline__ void rg_cluster_reduce_rows(cg::cluster_group& cluster, float* part, int csize, Fn fn) {
    const int rank = (int)cluster.block_rank();
    const int rpc = RG_ROWS / csize;                 // rows per CTA
    const int tpr = RG_THREADS / rpc;                // threads per row
    const int rr = threadIdx.x / tpr, tc = threadIdx.x - rr * tpr;
    const int row = rank * rpc + rr;
    for (int col = tc; col < D; col += tpr) {
        float s = 0.f;
        for (int q = 0; q < csize; ++q) {
            const float* remote = cluster.map_shared_rank(part, q);
            s += remote[row * D + col];
        }
        fn(row, col, s);
    }
}
