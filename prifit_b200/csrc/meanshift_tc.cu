// k2 (tensor-core engine) -- T mean-shift iterations of all N seeds on tcgen05 / TMEM / TMA.
// reference src/mean_shift.py:50-84 (mean_shift_, gaussian branch).
//
// One CTA owns 128 seed rows of one shape for ALL T iterations (rows are independent given X), so a
// single launch covers the whole stage and the N x N kernel matrix only ever exists as 128 x 128
// tiles in tensor memory.  Per key tile (128 keys x 128 d fp16 = 32 KB, one TMA transaction, ring):
//
//   GEMM1  S[128 x 128]  = Q[128 x d] . Xtile^T      A = Q in smem (K-major, rewritten once per iteration), B = Xtile in smem, K-major
//   softmax warps        : S -> P = 2^10 exp2((S - 1) log2e / bw^2), clamped at e^-13, rounded to f16 and
//                          written back over S in TMEM (thread = row, no shared-memory round trip)
//   GEMM2  O[128 x d]   += P[128 x 128] . Xtile      A = P in TMEM, B = the SAME smem tile, MN-major
//
// and per iteration the epilogue renormalises O row-wise (y' = O / ||O||: the 1/rowsum factor of the
// reference and the 2^10 scale both cancel in the normalisation) and stores it as the next Q (fp16,
// SWIZZLE_128B, K-major) in shared memory.  TMEM columns: S0|P0 [0,128)  S1|P1 [128,256)  S2|P2 [256,384)  O [384,512).
//
// What bounds it (in-kernel phase clocks, PRIFIT_MS_DBG=1 + scripts/ms_phases.py; profiles/r02_ms_phases.txt).
// tcgen05.mma issue BLOCKS: the queue behind the issuing thread is only an instruction or two deep, so the
// thread spends ~83 clk per MMA inside the 16 issue slots of a tile (64 clk nominal for M128 N128 K16; an SS MMA
// of this shape pulls 8 KB out of shared memory = the SM's whole 128 B/clk, next to the TMA fill) and whatever
// else it does -- barrier waits, commits -- is time the tensor pipe runs dry.  Hence (round 2):
//   * three S|P buffers and the issue order G1(j+2) . G2(j): the wait for the softmax of tile j has two GEMMs
//     queued behind it instead of none;
//   * the two softmax warps of an SM sub-partition work on DIFFERENT tiles (even / odd) and prefetch the next
//     32-column chunk of S during the exponentials of the current one: a tile's S -> P latency no longer sits
//     between two GEMMs of the same buffer, and each softmax warp is busy 1750 of the 3900 clk between its tiles;
//   * one mbarrier arrival per warp instead of per thread; six ring stages.
// 551 -> 520 us at cfg2.  Measured and rejected: a CTA-pair version (cta_group::2, M = 256, each CTA supplying
// half of B's N extent: 96 instead of 128 B/clk of operand traffic) -- bit-identical results, 757 us: every
// S -> P hand-off then crosses the pair and both tensor pipes run in lock-step with the slower softmax.
//
// Why kind::f16 and not kind::tf32: fp16 carries the same 10-bit mantissa as tf32 and every operand
// here lives in [6e-5, 2^10] (unit vectors, weights pre-scaled by 2^10), so the rounding is the same;
// but MN-major tf32 operands only exist in the SW128_32B shared-memory layout, which K-major operands
// cannot use, so one tf32 tile could not feed both GEMMs.  fp16 uses SWIZZLE_128B in both majors,
// halves the tile bytes and doubles the MMA rate.
//
// Warp roles (384 threads): warp 0 = TMA producer, warp 1 = MMA issuer, warp 2 = TMEM allocator,
// warps 4-11 = softmax (warps 4-7 even key tiles, 8-11 odd ones; thread = seed row) / epilogue (thread = seed row x 64-column half).
//
// This pass only feeds the NMS (discrete outcome); the K centres that carry gradient are recomputed in
// fp32 (meanshift_rows.cu).
#include <cudaTypedefs.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include "common.cuh"
#include "sm100_ptx.cuh"

using namespace sm100;

namespace {

constexpr int TC_D = 128;
constexpr int TC_BM = 128;                 // seed rows per CTA
constexpr int TC_BN = 128;                 // keys per tile
constexpr int TC_STAGES = 6;
constexpr uint32_t TC_TILE_BYTES = TC_BN * TC_D * 2;      // 32768
constexpr uint32_t TC_KBLOCK_BYTES = TC_BN * 128;          // one 64-column (128 B) block of the tile
constexpr int TC_THREADS = 384;             // warps 0-2: TMA / MMA / TMEM alloc, warps 4-11: softmax
constexpr int TC_SOFTMAX = 256;             // softmax threads: (seed row, 64-key half of every tile)
constexpr int TC_SBUF = 3;                 // S|P accumulators in flight
constexpr uint32_t COL_S0 = 0, COL_O = 128 * TC_SBUF;      // 3 x 128 + 128 = all 512 columns
constexpr uint32_t TC_Q_BYTES = TC_BM * TC_D * 2;            // Q tile: two 64-column blocks of [128 rows][128 B]
constexpr float LOG2E = 1.4426950408889634f;
constexpr float P_SCALE_LOG2 = 10.0f;      // weights are stored as 2^10 * kappa (keeps e^-13 a normal f16)

struct TcBarriers {
    uint64_t x_full[TC_STAGES];
    uint64_t x_empty[TC_STAGES];
    uint64_t s_full[TC_SBUF];
    uint64_t p_full[TC_SBUF];
    uint64_t o_full;
    uint64_t q_full;
    uint32_t tmem_base;
    float ssum[2][TC_BM];        // [column half][row] partial ||O||^2 (rewritten only after every warp's q_full arrival of the iteration)
};

constexpr size_t TC_SMEM_BYTES = 1024 /*alignment slack*/ + (size_t)TC_STAGES * TC_TILE_BYTES + TC_Q_BYTES + sizeof(TcBarriers);

__global__ void to_half_kernel(const float4* __restrict__ in, uint2* __restrict__ out, size_t n4) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = in[i];
        out[i] = make_uint2(pack_f16x2(v.x, v.y), pack_f16x2(v.z, v.w));
    }
}

template <int DBG>
__global__ void __launch_bounds__(TC_THREADS, 1) meanshift_tc_kernel(
    const __grid_constant__ CUtensorMap tmap, const __half* __restrict__ Xh, const float* __restrict__ bw,
    int N, int T, float* __restrict__ newX) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // offset on the __shared__ symbol: accesses stay LDS / STS
    uint8_t* tiles = smem;
    uint8_t* qtile = smem + (size_t)TC_STAGES * TC_TILE_BYTES;
    TcBarriers* bars = reinterpret_cast<TcBarriers*>(qtile + TC_Q_BYTES);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y, r0 = blockIdx.x * TC_BM;
    const int nt = (N + TC_BN - 1) / TC_BN;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&bars->x_full[s], 1); mbar_init(&bars->x_empty[s], 1); }
        for (int s = 0; s < TC_SBUF; ++s) { mbar_init(&bars->s_full[s], 1); mbar_init(&bars->p_full[s], TC_SOFTMAX / 64); }
        mbar_init(&bars->o_full, 1);
        mbar_init(&bars->q_full, TC_SOFTMAX / 32);      // one arrival per softmax warp
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) prefetch_tensormap(&tmap);
    if (warp == 2) { tmem_alloc(&bars->tmem_base, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;

    if (warp == 0) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            uint32_t it = 0;
            for (int t = 0; t < T; ++t)
                for (int j = 0; j < nt; ++j, ++it) {
                    const uint32_t st = it % TC_STAGES, ph = (it / TC_STAGES) & 1;
                    mbar_wait(&bars->x_empty[st], ph ^ 1);
                    mbar_arrive_expect_tx(&bars->x_full[st], TC_TILE_BYTES);
                    const uint32_t dst = smem_u32(tiles + (size_t)st * TC_TILE_BYTES);
                    tma_load_3d(dst, &tmap, &bars->x_full[st], 0, j * TC_BN, b);
                    tma_load_3d(dst + TC_KBLOCK_BYTES, &tmap, &bars->x_full[st], 64, j * TC_BN, b);
                }
        }
    } else if (warp == 1) {
        // ================================= MMA issuer =================================
        if (lane == 0) {
            constexpr uint32_t idesc1 = idesc_f16(TC_BM, TC_BN, false);   // S = Q . X^T   (B K-major)
            constexpr uint32_t idesc2 = idesc_f16(TC_BM, TC_D, true);     // O += P . X    (B MN-major)
            // tile i (counted over all iterations) uses ring stage i % STAGES and S|P buffer i % 3
            long long m_p = 0, m_x = 0, m_q = 0, m_g1 = 0, m_g2 = 0, m_tot = clock64();
            auto gemm1 = [&](uint32_t i) {
                const uint32_t st = i % TC_STAGES, xph = (i / TC_STAGES) & 1, buf = i % TC_SBUF;
                long long k0 = 0;
                if (DBG) k0 = clock64();
                mbar_wait(&bars->x_full[st], xph);
                tc_fence_after();
                if (DBG) { const long long n = clock64(); m_x += n - k0; k0 = n; }
                const uint32_t base = smem_u32(tiles + (size_t)st * TC_TILE_BYTES), qbase = smem_u32(qtile);
#pragma unroll
                for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {      // 16 d-elements (32 B) per MMA
                        const uint64_t ad = smem_desc_sw128(qbase + kb * TC_KBLOCK_BYTES + ks * 32, 16, 1024);
                        const uint64_t bd = smem_desc_sw128(base + kb * TC_KBLOCK_BYTES + ks * 32, 16, 1024);
                        mma_f16_ss(tmem + COL_S0 + buf * 128, ad, bd, idesc1, (kb | ks) != 0);
                    }
                mma_commit(&bars->s_full[buf]);
                if (DBG) m_g1 += clock64() - k0;
            };
            auto gemm2 = [&](uint32_t i, bool first_of_iter) {
                const uint32_t st = i % TC_STAGES, buf = i % TC_SBUF, ph = (i / TC_SBUF) & 1;
                long long k0 = 0;
                if (DBG) k0 = clock64();
                mbar_wait(&bars->p_full[buf], ph);
                tc_fence_after();
                if (DBG) { const long long n = clock64(); m_p += n - k0; k0 = n; }
                const uint32_t base = smem_u32(tiles + (size_t)st * TC_TILE_BYTES);
#pragma unroll
                for (int kk = 0; kk < TC_BN / 16; ++kk) {      // 16 keys per MMA = two 8-row groups, 1024 B apart
                    const uint64_t bd = smem_desc_sw128(base + kk * 2048, TC_KBLOCK_BYTES, 1024);
                    // P of keys [64h, 64h+64) sits packed in columns [64h, 64h+32) of the S buffer it replaces
                    mma_f16_ts(tmem + COL_O, tmem + COL_S0 + buf * 128 + (kk >> 2) * 64 + (kk & 3) * 8, bd, idesc2,
                               !(first_of_iter && kk == 0));
                }
                mma_commit(&bars->x_empty[st]);
                if (DBG) m_g2 += clock64() - k0;
            };
            // Issue order inside an iteration: G1(0) G1(1) | G1(j+2) G2(j) ...: the S tile of key tile j+2 is queued BEFORE
            // the issuer waits for the softmax of tile j, so the tensor pipe (whose queue is a few MMAs deep: issue
            // blocks for about one MMA time per instruction) always has work behind the wait and the softmax warps
            // find their next S tile finished.  G1(j+2) overwrites the buffer of tile j-1, whose GEMM2 was issued before.
            uint32_t it = 0;
            for (int t = 0; t < T; ++t) {
                long long q0 = 0;
                if (DBG) q0 = clock64();
                mbar_wait(&bars->q_full, t & 1);
                tc_fence_after();
                if (DBG) m_q += clock64() - q0;
                gemm1(it);
                if (nt > 1) gemm1(it + 1);
                for (int j = 0; j < nt; ++j, ++it) {
                    if (j + 2 < nt) gemm1(it + 2);
                    gemm2(it, j == 0);
                }
                mma_commit(&bars->o_full);
            }
            if (DBG && blockIdx.x == 0 && blockIdx.y == 0)
                printf("mma thread: tiles %u  per tile: wait_x %lld  issue_g1+commit %lld  wait_p %lld  issue_g2+commit %lld | per iteration wait_q %lld | total %lld\n",
                       it, m_x / it, m_g1 / it, m_p / it, m_g2 / it, m_q / T, clock64() - m_tot);
        }
    } else if (warp >= 4) {
        // ============================= softmax / epilogue ==============================
        const int ew = warp - 4, half = ew >> 2;
        const int row = 32 * (ew & 3) + lane;                    // TMEM lane == seed row within the tile
        const uint32_t lane_base = (uint32_t)(32 * (ew & 3)) << 16;
        const bool row_ok = r0 + row < N;
        const float bwv = bw[b];
        const float c1 = LOG2E / (bwv * bwv), c0 = P_SCALE_LOG2 - c1, lo2 = PRIFIT_LO * LOG2E + P_SCALE_LOG2;
        uint32_t v[2][32], h[16];
        // Q^0 = this tile's own rows of X (already fp16: 256 B per row = 64 packed columns, 32 per half)
        // (row, half) owns row `row` of 64-column block `half` of the Q tile: eight 16-byte chunks, chunk c at c ^ (row & 7)
        uint8_t* qrow = qtile + (size_t)half * TC_KBLOCK_BYTES + (size_t)row * 128;
        const uint4* xrow = reinterpret_cast<const uint4*>(Xh + ((size_t)b * N + (row_ok ? r0 + row : 0)) * TC_D) + 8 * half;
#pragma unroll
        for (int c = 0; c < 8; ++c)
            *reinterpret_cast<uint4*>(qrow + ((c ^ (row & 7)) << 4)) = row_ok ? xrow[c] : make_uint4(0u, 0u, 0u, 0u);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->q_full);

        // Softmax: warps 4-7 take the even key tiles, warps 8-11 the odd ones (thread = one seed row, all 128 keys of the
        // tile), so the two warps that share an SM sub-partition are out of phase: while one waits for its S tile, loads
        // or stores, the other keeps the MUFU pipe busy.  Inside a tile the row is processed in four 32-column chunks with
        // the tcgen05.ld of the next chunk in flight during the exponentials of the current one.
        auto to_p = [&](const uint32_t (&sv)[32], int key0, uint32_t dst) {
            // P = 2^10 exp(clamp((s-1)/bw^2, -13, .)) as f16 pairs; padded keys weigh nothing
            if (key0 + 32 <= N) {
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const float x0 = fmaxf(fmaf(__uint_as_float(sv[2 * e]), c1, c0), lo2);
                    const float x1 = fmaxf(fmaf(__uint_as_float(sv[2 * e + 1]), c1, c0), lo2);
                    h[e] = pack_f16x2(ex2_approx(x0), ex2_approx(x1));
                }
            } else {
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    float p0 = ex2_approx(fmaxf(fmaf(__uint_as_float(sv[2 * e]), c1, c0), lo2));
                    float p1 = ex2_approx(fmaxf(fmaf(__uint_as_float(sv[2 * e + 1]), c1, c0), lo2));
                    if (key0 + 2 * e >= N) p0 = 0.f;
                    if (key0 + 2 * e + 1 >= N) p1 = 0.f;
                    h[e] = pack_f16x2(p0, p1);
                }
            }
            tmem_st16(dst, h);
        };
        uint32_t it = 0, mine = 0;
        long long c_wait = 0, c_busy = 0, c_epi = 0, c_tot = clock64();
        for (int t = 0; t < T; ++t) {
            for (int j = 0; j < nt; ++j, ++it) {
                if ((int)(it & 1) != half) continue;
                const uint32_t buf = it % TC_SBUF, ph = (it / TC_SBUF) & 1;
                long long k0 = 0, k1 = 0;
                if (DBG) k0 = clock64();
                mbar_wait(&bars->s_full[buf], ph);
                tc_fence_after();
                if (DBG) k1 = clock64();
                // S columns [32 c, 32 c + 32) -> P (packed pairs) in columns 64 (c / 2) + 16 (c % 2) .. + 16 of the same buffer:
                // every store lands on columns this thread has already loaded
                const uint32_t sb = tmem + lane_base + COL_S0 + buf * 128;
                const int key0 = j * TC_BN;
                tmem_ld32(sb, v[0]);
                tmem_wait_ld();
                tmem_ld32(sb + 32, v[1]);
                to_p(v[0], key0, sb);
                tmem_wait_ld();
                tmem_ld32(sb + 64, v[0]);
                to_p(v[1], key0 + 32, sb + 16);
                tmem_wait_ld();
                tmem_ld32(sb + 96, v[1]);
                to_p(v[0], key0 + 64, sb + 64);
                tmem_wait_ld();
                to_p(v[1], key0 + 96, sb + 80);
                tmem_wait_st();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->p_full[buf]);
                if (DBG) { c_wait += k1 - k0; c_busy += clock64() - k1; ++mine; }
            }
            long long e0 = 0;
            if (DBG) e0 = clock64();
            // epilogue of iteration t: y' = O / ||O||  (each thread owns 64 of the row's 128 columns)
            mbar_wait(&bars->o_full, t & 1);
            tc_fence_after();
            tmem_ld32(tmem + lane_base + COL_O + 64 * half, v[0]);
            tmem_ld32(tmem + lane_base + COL_O + 64 * half + 32, v[1]);
            tmem_wait_ld();
            float ss = 0.f;
#pragma unroll
            for (int c = 0; c < 2; ++c)
#pragma unroll
                for (int e = 0; e < 32; ++e) ss = fmaf(__uint_as_float(v[c][e]), __uint_as_float(v[c][e]), ss);
            bars->ssum[half][row] = ss;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const float inv = 1.0f / sqrtf(bars->ssum[0][row] + bars->ssum[1][row]);
            if (t == T - 1) {
                if (row_ok) {
                    float4* orow = reinterpret_cast<float4*>(newX + ((size_t)b * N + r0 + row) * TC_D) + 16 * half;
#pragma unroll
                    for (int c = 0; c < 2; ++c)
#pragma unroll
                        for (int e = 0; e < 8; ++e)
                            orow[c * 8 + e] = make_float4(__uint_as_float(v[c][4 * e]) * inv, __uint_as_float(v[c][4 * e + 1]) * inv,
                                                          __uint_as_float(v[c][4 * e + 2]) * inv, __uint_as_float(v[c][4 * e + 3]) * inv);
                }
            } else {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
#pragma unroll
                    for (int e = 0; e < 16; ++e)
                        h[e] = pack_f16x2(__uint_as_float(v[c][2 * e]) * inv, __uint_as_float(v[c][2 * e + 1]) * inv);
#pragma unroll
                    for (int q = 0; q < 4; ++q)        // columns [32 c + 8 q, +8) of this half = chunk 4 c + q
                        *reinterpret_cast<uint4*>(qrow + (((4 * c + q) ^ (row & 7)) << 4)) = make_uint4(h[4 * q], h[4 * q + 1], h[4 * q + 2], h[4 * q + 3]);
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->q_full);
            }
            if (DBG) c_epi += clock64() - e0;
        }
        if (DBG && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0 && (warp == 4 || warp == 8))
            printf("softmax warp %d: own tiles %u  per own tile: wait_s %lld  busy %lld | per iteration epilogue (incl. wait) %lld | total %lld\n",
                   warp, mine, c_wait / mine, c_busy / mine, c_epi / T, clock64() - c_tot);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// --------------------------------------------------------------------------------------------------
// Descriptor self-test: D[128 x 128] = A . B^T (mode 0, B K-major) or A . B (mode 1, B MN-major) with
// A staged in TMEM (packed f16 pairs) and B fetched by TMA exactly like the kernel above.  lbo/sbo are
// runtime values so the encodings can be verified on hardware (tests/test_gpu_tcgen05_probe.py).
// --------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) tc_probe_kernel(const __grid_constant__ CUtensorMap tmap,
                                                           const float* __restrict__ A, int mode, uint32_t lbo, uint32_t sbo,
                                                           float* __restrict__ Dout) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // offset on the __shared__ symbol: accesses stay LDS / STS
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + TC_TILE_BYTES);      // [0] tile full, [1] mma done
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
    const int warp = threadIdx.x >> 5, row = threadIdx.x;
    if (threadIdx.x == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_barrier_init(); }
    if (warp == 0) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t lane_base = (uint32_t)(32 * warp) << 16;
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bars[0], TC_TILE_BYTES);
        tma_load_3d(smem_u32(smem), &tmap, &bars[0], 0, 0, 0);
        tma_load_3d(smem_u32(smem) + TC_KBLOCK_BYTES, &tmap, &bars[0], 64, 0, 0);
    }
    uint32_t h[16], v[32];
    for (int c = 0; c < 4; ++c) {
        for (int e = 0; e < 16; ++e) h[e] = pack_f16x2(A[row * 128 + 32 * c + 2 * e], A[row * 128 + 32 * c + 2 * e + 1]);
        tmem_st16(tmem + lane_base + 128 + 16 * c, h);         // A (packed) at columns [128,192)
    }
    tmem_wait_st();
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
        tc_fence_after();
        mbar_wait(&bars[0], 0);
        tc_fence_after();
        const uint32_t base = smem_u32(smem);
        const uint32_t idesc = idesc_f16(128, 128, mode == 1);
        for (int kk = 0; kk < 8; ++kk) {
            const uint32_t addr = mode == 0 ? base + (kk >> 2) * TC_KBLOCK_BYTES + (kk & 3) * 32 : base + kk * 2048;
            mma_f16_ts(tmem, tmem + 128 + kk * 8, smem_desc_sw128(addr, lbo, sbo), idesc, kk != 0);
        }
        mma_commit(&bars[1]);
    }
    mbar_wait(&bars[1], 0);
    tc_fence_after();
    for (int c = 0; c < 4; ++c) {
        tmem_ld32(tmem + lane_base + 32 * c, v);
        tmem_wait_ld();
        for (int e = 0; e < 32; ++e) Dout[row * 128 + 32 * c + e] = __uint_as_float(v[e]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }
    return fn;
}

// 3-D map over Xh[B][N][128] fp16: box = 64 columns (128 B, SWIZZLE_128B) x 128 rows x 1 shape;
// rows beyond N are zero-filled by the TMA unit.
int make_tile_map(CUtensorMap* map, const __half* X, int B, int N) {
    PFN_cuTensorMapEncodeTiled_v12000 enc = get_encode_fn();
    if (!enc) { prifit_set_error("cuTensorMapEncodeTiled driver entry point not available"); return PRIFIT_E_NODEVICE; }
    cuuint64_t dims[3] = {(cuuint64_t)TC_D, (cuuint64_t)N, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)TC_D * 2, (cuuint64_t)N * TC_D * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)TC_BN, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<__half*>(X), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { prifit_set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return PRIFIT_E_BADARG; }
    return 0;
}

int convert_to_half(const float* X, __half* Xh, size_t n, cudaStream_t st) {
    const size_t n4 = n / 4;
    to_half_kernel<<<(unsigned)min((size_t)148 * 8, (n4 + 255) / 256), 256, 0, st>>>(
        reinterpret_cast<const float4*>(X), reinterpret_cast<uint2*>(Xh), n4);
    PF_LAUNCH_CHECK();
    return 0;
}

}  // namespace

// shared with gram_tc.cu
int prifit_tc_convert_to_half(const float* X, __half* Xh, size_t n, cudaStream_t st) { return convert_to_half(X, Xh, n, st); }
int prifit_tc_make_tile_map(CUtensorMap* map, const __half* X, int B, int N) { return make_tile_map(map, X, B, N); }

size_t prifit_meanshift_tc_workspace_bytes(int B, int N) {
    return (size_t)B * N * TC_D * sizeof(__half) + 256;
}

int prifit_meanshift_fwd_tc(const float* X, const float* bw, int B, int N, int T, float* newX,
                            void* ws, size_t ws_bytes, cudaStream_t st) {
    (void)ws_bytes;
    if (T == 0) {
        PF_CUDA(cudaMemcpyAsync(newX, X, (size_t)B * N * TC_D * sizeof(float), cudaMemcpyDeviceToDevice, st));
        return 0;
    }
    __half* Xh = reinterpret_cast<__half*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
    int rc = convert_to_half(X, Xh, (size_t)B * N * TC_D, st);
    if (rc) return rc;
    CUtensorMap map;
    rc = make_tile_map(&map, Xh, B, N);
    if (rc) return rc;
    dim3 grid((N + TC_BM - 1) / TC_BM, B);
    static const bool dbg = [] { const char* e = getenv("PRIFIT_MS_DBG"); return e && atoi(e) != 0; }();   // in-kernel phase clocks (scripts/ms_phases.py)
    auto kern = dbg ? meanshift_tc_kernel<1> : meanshift_tc_kernel<0>;
    PF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES));
    kern<<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(map, Xh, bw, N, T, newX);
    PF_LAUNCH_CHECK();
    return 0;
}

// diagnostics: see tc_probe_kernel.  A[128,128], Bm[128,128], D[128,128] device fp32; ws >= 64 KB + 256.
extern "C" int prifit_debug_tc_probe(const float* A, const float* Bm, int mode, int lbo_bytes, int sbo_bytes,
                                     float* D, void* ws, void* stream) {
    PF_CHECK_ARG(A && Bm && D && ws, PRIFIT_E_BADARG, "null pointer");
    __half* Bh = reinterpret_cast<__half*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
    int rc = convert_to_half(Bm, Bh, 128 * 128, pf_stream(stream));
    if (rc) return rc;
    CUtensorMap map;
    rc = make_tile_map(&map, Bh, 1, 128);
    if (rc) return rc;
    const size_t smem = 1024 + TC_TILE_BYTES + 64;
    PF_CUDA(cudaFuncSetAttribute(tc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc_probe_kernel<<<1, 128, smem, pf_stream(stream)>>>(map, A, mode, (uint32_t)lbo_bytes, (uint32_t)sbo_bytes, D);
    PF_LAUNCH_CHECK();
    return 0;
}
