// Tensor-core Gram engine (tcgen05 / TMEM / TMA) with fused row-wise epilogues:
//
//   NEAREST  nearest[i] = argmin_j (2 - 2 <x_i, x_j>)                       NMS step 1, src/mean_shift.py:169-172
//   BEST     best[i]    = argmax_j ([2 - 2 <x_i, x_j> < bw] * votes[j])     NMS step 3, src/mean_shift.py:187-194
//   HIST     per-row 512-bin histogram of 2 - 2 <x_i, x_j>  -> bin holding the k-th smallest    } bandwidth,
//   COLLECT  candidates inside that bin (+- a rigorous fp16 error margin), EXACT fp32 recompute  } src/mean_shift.py:153-158
//            of just those, exact k-th order statistic, sqrt(max(., 1e-6))
//
// (the Gram matrix is symmetric and the products commute bit-for-bit, so the reference's column-wise
// arg-reductions equal these row-wise ones, lowest index on ties.)
//
// Same skeleton as meanshift_tc.cu: a CTA owns 128 rows (A operand = its rows as packed f16 in TMEM),
// streams every 128-key tile of the shape through a TMA ring (B operand, K-major, SWIZZLE_128B),
// 8 x tcgen05.mma.kind::f16 (M128 N128 K16) per tile into a double-buffered S accumulator, and the
// four epilogue warps (thread = row = TMEM lane) consume S straight from tensor memory.  The n x n
// matrix never exists outside TMEM.
//
// Exactness of the bandwidth: fp16 operands bound the error of every distance by eps = 2^-9
// (|d a.b| <= 2^-10 ||a|| ||b||).  HIST finds the bin of the k-th smallest fp16-distance; the true k-th
// smallest lies within eps of it; COLLECT keeps every element within 2 eps of the bin, counts the
// elements below, recomputes the kept ones in fp32 and ranks them -- so the result is the exact fp32
// order statistic, independent of the fp16 rounding.
#include <cudaTypedefs.h>
#include <cuda_fp16.h>
#include "common.cuh"
#include "sm100_ptx.cuh"

using namespace sm100;

int prifit_tc_convert_to_half(const float* X, __half* Xh, size_t n, cudaStream_t st);
int prifit_tc_make_tile_map(CUtensorMap* map, const __half* X, int B, int N);

namespace {

constexpr int G_D = 128, G_BM = 128, G_BN = 128, G_THREADS = 256;
constexpr uint32_t G_TILE_BYTES = G_BN * G_D * 2;        // 32 KB
constexpr uint32_t G_KBLOCK = G_BN * 128;
constexpr uint32_t GCOL_Q = 256;
enum { GM_NEAREST = 0, GM_BEST = 1, GM_HIST = 2, GM_COLLECT = 3 };

constexpr int HIST_BINS = 512;                 // over [0, 4]: 128 bins per unit
constexpr float HIST_SCALE = 128.0f;
constexpr int HIST_STRIDE = HIST_BINS * 2 + 4; // bytes per row (odd number of words: conflict-free)
constexpr float BW_MARGIN = 2.0f * 0.001953125f;   // 2 eps, eps = 2^-9
constexpr int CAND_CAP = 256;

template <int MODE> struct GCfg {
    static constexpr int stages = MODE == GM_HIST ? 2 : 3;
    static constexpr size_t scratch = MODE == GM_BEST ? 2 * 128 * sizeof(float)
                                    : MODE == GM_HIST ? (size_t)G_BM * HIST_STRIDE
                                    : MODE == GM_COLLECT ? (size_t)G_BM * CAND_CAP * 2 + 4 * CAND_CAP * sizeof(float) + 2 * G_BM * sizeof(int)
                                    : 16;
    static constexpr size_t smem = 1024 + (size_t)stages * G_TILE_BYTES + 256 + scratch;
};

struct GBars {
    uint64_t x_full[3], x_empty[3], s_full[2], s_free[2], q_full;
    uint32_t tmem_base;
};

struct GramArgs {
    const __half* Xh;        // [B,N,128] fp16 rows (A and B operands)
    const float* X32;        // [B,N,128] fp32 rows (COLLECT: exact recompute)
    const float* bw;         // [B]       (BEST)
    const int32_t* votes;    // [B,N]     (BEST)
    const int32_t* kth;      // [B]       (HIST / COLLECT), 1-based rank
    int32_t* out_idx;        // [B,N]     (NEAREST / BEST)
    int2* rowinfo;           // [B,N]     (HIST out, COLLECT in): (bin, count below the bin)
    float* rowval;           // [B,N]     (COLLECT out)
    int32_t* overflow;       // [1]       (COLLECT out): candidate list overflow / window miss
    int N;
};

__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

template <int MODE>
__global__ void __launch_bounds__(G_THREADS, 1) gram_tc_kernel(const __grid_constant__ CUtensorMap tmap, const GramArgs a) {
    constexpr int STAGES = GCfg<MODE>::stages;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* tiles = smem;
    GBars* bars = reinterpret_cast<GBars*>(smem + (size_t)STAGES * G_TILE_BYTES);
    uint8_t* scratch = smem + (size_t)STAGES * G_TILE_BYTES + 256;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y, r0 = blockIdx.x * G_BM, N = a.N;
    const int nt = (N + G_BN - 1) / G_BN;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&bars->x_full[s], 1); mbar_init(&bars->x_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&bars->s_full[s], 1); mbar_init(&bars->s_free[s], 128); }
        mbar_init(&bars->q_full, 128);
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) prefetch_tensormap(&tmap);
    if (warp == 2) { tmem_alloc(&bars->tmem_base, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;

    if (warp == 0) {
        if (lane == 0) {
            for (int j = 0; j < nt; ++j) {
                const uint32_t st = j % STAGES, ph = (j / STAGES) & 1;
                mbar_wait(&bars->x_empty[st], ph ^ 1);
                mbar_arrive_expect_tx(&bars->x_full[st], G_TILE_BYTES);
                const uint32_t dst = smem_u32(tiles + (size_t)st * G_TILE_BYTES);
                tma_load_3d(dst, &tmap, &bars->x_full[st], 0, j * G_BN, b);
                tma_load_3d(dst + G_KBLOCK, &tmap, &bars->x_full[st], 64, j * G_BN, b);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = idesc_f16(G_BM, G_BN, false);
            mbar_wait(&bars->q_full, 0);
            tc_fence_after();
            for (int j = 0; j < nt; ++j) {
                const uint32_t st = j % STAGES, xph = (j / STAGES) & 1, buf = j & 1;
                mbar_wait(&bars->x_full[st], xph);
                mbar_wait(&bars->s_free[buf], ((j >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t base = smem_u32(tiles + (size_t)st * G_TILE_BYTES);
#pragma unroll
                for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        mma_f16_ts(tmem + buf * 128, tmem + GCOL_Q + kb * 32 + ks * 8,
                                   smem_desc_sw128(base + kb * G_KBLOCK + ks * 32, 16, 1024), idesc, (kb | ks) != 0);
                mma_commit(&bars->s_full[buf]);
                mma_commit(&bars->x_empty[st]);
            }
        }
    } else if (warp >= 4) {
        const int row = threadIdx.x - 128, ew = warp - 4;
        const uint32_t lane_base = (uint32_t)(32 * ew) << 16;
        const bool row_ok = r0 + row < N;
        const size_t grow = (size_t)b * N + (row_ok ? r0 + row : 0);
        uint32_t v[32], h[16];
        const uint4* xrow = reinterpret_cast<const uint4*>(a.Xh + grow * G_D);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const uint4 f = row_ok ? xrow[c * 4 + e] : make_uint4(0u, 0u, 0u, 0u);
                h[4 * e] = f.x; h[4 * e + 1] = f.y; h[4 * e + 2] = f.z; h[4 * e + 3] = f.w;
            }
            tmem_st16(tmem + lane_base + GCOL_Q + 16 * c, h);
        }
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(&bars->q_full);

        // ---- per-mode state
        float best = MODE == GM_NEAREST ? INFINITY : -1.0f;
        int besti = 0;
        float bwv = 0.f;
        float* vt = reinterpret_cast<float*>(scratch);                                   // BEST: [2][128]
        uint16_t* hist = reinterpret_cast<uint16_t*>(scratch + (size_t)row * HIST_STRIDE);   // HIST: own row
        uint16_t* cand = reinterpret_cast<uint16_t*>(scratch) + (size_t)row * CAND_CAP;  // COLLECT: own row
        float win_lo = 0.f, win_hi = 0.f;
        int below = 0, ncand = 0;
        float vnext = 0.f;
        if (MODE == GM_BEST) {
            bwv = a.bw[b];
            vnext = row < N ? (float)a.votes[(size_t)b * N + row] : 0.f;
        }
        if (MODE == GM_HIST) {
            uint32_t* hw = reinterpret_cast<uint32_t*>(hist);
            for (int q = 0; q < HIST_BINS / 2; ++q) hw[q] = 0u;
        }
        if (MODE == GM_COLLECT) {
            const int2 ri = a.rowinfo[grow];
            win_lo = (float)ri.x / HIST_SCALE - BW_MARGIN;
            win_hi = (float)(ri.x + 1) / HIST_SCALE + BW_MARGIN;
        }

        for (int j = 0; j < nt; ++j) {
            const uint32_t buf = j & 1, ph = (j >> 1) & 1;
            const int key0 = j * G_BN;
            if (MODE == GM_BEST) {
                vt[buf * 128 + row] = vnext;                  // votes of this tile's columns (prefetched)
                epi_barrier();
                const int col = key0 + G_BN + row;            // prefetch the next tile's votes
                vnext = col < N ? (float)a.votes[(size_t)b * N + col] : 0.f;
            }
            mbar_wait(&bars->s_full[buf], ph);
            tc_fence_after();
            const int ncols = min(G_BN, N - key0);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                tmem_ld32(tmem + lane_base + buf * 128 + 32 * c, v);
                tmem_wait_ld();
#pragma unroll
                for (int e = 0; e < 32; ++e) {
                    const int cl = 32 * c + e;
                    const float dist = fmaf(-2.0f, __uint_as_float(v[e]), 2.0f);        // 2.0 - 2.0 * s
                    if (MODE == GM_HIST) {
                        // groups of four counter updates with their loads in flight together; equal bins
                        // inside a group are forwarded in registers (stores stay in program order)
                        if ((e & 3) == 3) {
                            int bn[4], inc[4], cnt[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const float du = fmaf(-2.0f, __uint_as_float(v[e - 3 + u]), 2.0f);
                                bn[u] = min(max((int)(du * HIST_SCALE), 0), HIST_BINS - 1);
                                inc[u] = (cl - 3 + u) < ncols ? 1 : 0;
                            }
#pragma unroll
                            for (int u = 0; u < 4; ++u) cnt[u] = hist[bn[u]];
                            cnt[0] += inc[0];
                            cnt[1] = (bn[1] == bn[0] ? cnt[0] : cnt[1]) + inc[1];
                            cnt[2] = (bn[2] == bn[1] ? cnt[1] : bn[2] == bn[0] ? cnt[0] : cnt[2]) + inc[2];
                            cnt[3] = (bn[3] == bn[2] ? cnt[2] : bn[3] == bn[1] ? cnt[1] : bn[3] == bn[0] ? cnt[0] : cnt[3]) + inc[3];
#pragma unroll
                            for (int u = 0; u < 4; ++u) hist[bn[u]] = (uint16_t)cnt[u];
                        }
                    } else if (cl < ncols) {
                        if (MODE == GM_NEAREST) {
                            if (dist < best) { best = dist; besti = key0 + cl; }
                        } else if (MODE == GM_BEST) {
                            const float val = dist < bwv ? vt[buf * 128 + cl] : 0.f;
                            if (val > best) { best = val; besti = key0 + cl; }
                        } else {
                            below += dist < win_lo ? 1 : 0;
                            if (dist >= win_lo && dist <= win_hi) {
                                if (ncand < CAND_CAP) cand[ncand] = (uint16_t)(key0 + cl);
                                ++ncand;
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&bars->s_free[buf]);
        }

        // ---- per-mode finalisation
        if (MODE == GM_NEAREST || MODE == GM_BEST) {
            if (row_ok) a.out_idx[grow] = besti;
        } else if (MODE == GM_HIST) {
            if (row_ok) {
                const int k = max(1, min(a.kth[b], N));
                int cum = 0, bin = HIST_BINS - 1, before = 0;
                for (int q = 0; q < HIST_BINS; ++q) {
                    const int cnt = hist[q];
                    if (cum + cnt >= k) { bin = q; before = cum; break; }
                    cum += cnt;
                }
                a.rowinfo[grow] = make_int2(bin, before);
            }
        } else {
            // exact fp32 recompute of the candidates, one warp per row (lanes split the candidates)
            float* vals = reinterpret_cast<float*>(scratch + (size_t)G_BM * CAND_CAP * 2) + ew * CAND_CAP;
            int* below_s = reinterpret_cast<int*>(scratch + (size_t)G_BM * CAND_CAP * 2 + 4 * CAND_CAP * sizeof(float));
            int* ncand_s = below_s + G_BM;
            below_s[row] = below;
            ncand_s[row] = ncand;
            __syncwarp();
            const int k = max(1, min(a.kth[b], N));
            for (int rr = 0; rr < 32; ++rr) {
                const int r = 32 * ew + rr;
                if (r0 + r >= N) break;                                         // warp-uniform
                const int nc = ncand_s[r], m = k - below_s[r];                 // m-th smallest candidate (1-based)
                if (nc > CAND_CAP || m < 1 || m > nc) {
                    if (lane == 0) { atomicExch(a.overflow, 1); a.rowval[(size_t)b * N + r0 + r] = 0.f; }
                    continue;
                }
                // the whole warp works on one candidate at a time: lane l owns dims 4l..4l+3, so every
                // candidate row is one coalesced 512-byte request (4 candidates in flight per step)
                const float4 xr = __ldg(reinterpret_cast<const float4*>(a.X32 + ((size_t)b * N + r0 + r) * G_D) + lane);
                const uint16_t* cr = reinterpret_cast<const uint16_t*>(scratch) + (size_t)r * CAND_CAP;
                for (int c0 = 0; c0 < nc; c0 += 4) {
                    float acc[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int ci = min(c0 + u, nc - 1);
                        const float4 w = __ldg(reinterpret_cast<const float4*>(a.X32 + ((size_t)b * N + cr[ci]) * G_D) + lane);
                        acc[u] = fmaf(xr.x, w.x, fmaf(xr.y, w.y, fmaf(xr.z, w.z, xr.w * w.w)));
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) acc[u] = warp_sum(acc[u]);
                    if (lane < 4 && c0 + lane < nc)
                        vals[c0 + lane] = 2.0f - 2.0f * (lane == 0 ? acc[0] : lane == 1 ? acc[1] : lane == 2 ? acc[2] : acc[3]);
                }
                __syncwarp();
                float found = 0.f;
                bool have = false;
                for (int ci = lane; ci < nc; ci += 32) {
                    const float vi = vals[ci];
                    int rank = 0;
                    for (int cj = 0; cj < nc; ++cj) {
                        const float vj = vals[cj];
                        rank += (vj < vi || (vj == vi && cj < ci)) ? 1 : 0;
                    }
                    if (rank == m - 1) { found = vi; have = true; }
                }
                const unsigned ball = __ballot_sync(0xffffffffu, have);
                if (ball) {
                    const float res = __shfl_sync(0xffffffffu, found, __ffs(ball) - 1);
                    if (lane == 0) a.rowval[(size_t)b * N + r0 + r] = sqrtf(fmaxf(res, 1e-6f));   // guard_sqrt(., 1e-6)
                } else if (lane == 0) {
                    atomicExch(a.overflow, 1);
                }
                __syncwarp();
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int MODE>
int launch_gram(const CUtensorMap& map, const GramArgs& a, int B, cudaStream_t st) {
    PF_CUDA(cudaFuncSetAttribute(gram_tc_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GCfg<MODE>::smem));
    dim3 grid((a.N + G_BM - 1) / G_BM, B);
    gram_tc_kernel<MODE><<<grid, G_THREADS, GCfg<MODE>::smem, st>>>(map, a);
    PF_LAUNCH_CHECK();
    return 0;
}

}  // namespace

// ---- NMS steps 1 and 3 on the tensor cores (called from nms.cu).  Xh_ws: B*N*128 halves of scratch.
int prifit_tc_nms_nearest(const float* newX, int B, int N, __half* Xh_ws, CUtensorMap* map_out, int32_t* nearest, cudaStream_t st) {
    int rc = prifit_tc_convert_to_half(newX, Xh_ws, (size_t)B * N * G_D, st);
    if (rc) return rc;
    rc = prifit_tc_make_tile_map(map_out, Xh_ws, B, N);
    if (rc) return rc;
    GramArgs a = {};
    a.Xh = Xh_ws; a.N = N; a.out_idx = nearest;
    return launch_gram<GM_NEAREST>(*map_out, a, B, st);
}

int prifit_tc_nms_best(const CUtensorMap* map, const __half* Xh, const float* bw, const int32_t* votes, int B, int N,
                       int32_t* best, cudaStream_t st) {
    GramArgs a = {};
    a.Xh = Xh; a.N = N; a.bw = bw; a.votes = votes; a.out_idx = best;
    return launch_gram<GM_BEST>(*map, a, B, st);
}

// ---- bandwidth order statistic on the tensor cores (called from bandwidth.cu)
int prifit_tc_bandwidth_rows(const float* X, int B, int N, const int32_t* kth, __half* Xh_ws, int2* rowinfo_ws,
                             float* rowval, int32_t* overflow, cudaStream_t st) {
    int rc = prifit_tc_convert_to_half(X, Xh_ws, (size_t)B * N * G_D, st);
    if (rc) return rc;
    CUtensorMap map;
    rc = prifit_tc_make_tile_map(&map, Xh_ws, B, N);
    if (rc) return rc;
    GramArgs a = {};
    a.Xh = Xh_ws; a.X32 = X; a.N = N; a.kth = kth; a.rowinfo = rowinfo_ws; a.rowval = rowval; a.overflow = overflow;
    rc = launch_gram<GM_HIST>(map, a, B, st);
    if (rc) return rc;
    return launch_gram<GM_COLLECT>(map, a, B, st);
}
