// Tensor-core Gram engine (tcgen05 / TMEM / TMA) with fused row-wise epilogues:
//
//   NEAREST  nearest[i] = argmin_j (2 - 2 <x_i, x_j>)                       NMS step 1, src/mean_shift.py:169-172
//   BEST     best[i]    = argmax_j ([2 - 2 <x_i, x_j> < bw] * votes[j])     NMS step 3, src/mean_shift.py:187-194
//            (only for the rows i that received votes -- the reference indexes nbrs[uniques] too)
//   HIST     per-row 256-bin histogram of the distances                                       } bandwidth,
//            level 0: LOGARITHMIC bins (16 per binade over 16 binades: 6 % wide) -> the bin    } src/mean_shift.py:153-158
//            holding the k-th smallest; level 1 (only launched to work when some row's bin     }
//            holds more candidates than COLLECT's row list takes): 256 uniform sub-bins        }
//   COLLECT  candidates inside the final window (+- margin) with their tensor-core values;    }
//            the k-th of those is located on the tensor-core values, and only the candidates  }
//            within the rounding margin of it are recomputed EXACTLY in fp32 and ranked:      }
//            exact k-th order statistic, sqrt(max(., 1e-6))                                   }
//   DUMP     the distance matrix itself (tests only)
//
// (the Gram matrix is symmetric and the products commute bit-for-bit, so the reference's column-wise
// arg-reductions equal these row-wise ones, lowest index on ties.)
//
// Precision.  fp16 operands alone would put an error of 2^-9 on every distance, which is what decides
// threshold tests and the order statistic.  Every row is therefore split into two fp16 vectors,
//      x * 2^8 = hi + lo,   hi = fp16(x * 2^8),   lo = fp16(x * 2^8 - hi)          (22 significant bits)
// and the dot product is accumulated in fp32 by THREE tcgen05.mma.kind::f16 per 16 d-elements:
//      2^16 <a, b>  ~=  lo_a.hi_b + hi_a.lo_b + hi_a.hi_b                           (|error| < ~1e-6 on <a,b>)
// The 2^8 pre-scale keeps `lo` a normal fp16 for every element that matters.  The tensor pipe has so
// much headroom here (the epilogues bound these kernels) that the 3x MMA work is free.
//
// Skeleton: a CTA owns 128 rows (A operands = its rows' hi | lo halves as packed f16 in TMEM), streams
// every 128-key tile of the shape through a TMA ring (B operands hi + lo, K-major, SWIZZLE_128B, 64 KB
// per stage), 24 MMAs (M128 N128 K16) per tile into a double-buffered S accumulator, and SIXTEEN epilogue
// warps (four per SM sub-partition: thread = (row, 32-column quarter)) consume S straight from tensor
// memory.  The n x n matrix never exists outside TMEM.
//
// Exactness of the bandwidth.  The histogram bins partition the tensor-core distances exactly (the MMA
// sequence is deterministic and every pass evaluates the same expression; log bins are bit fields of
// the value, uniform sub-bins have power-of-two widths), so after HIST the k-th smallest tensor-core
// distance t_k is known to lie in one bin.  COLLECT keeps every element within BW_MARGIN of that bin
// with its tensor-core value and counts the elements below; t_k is then the (k - below)-th kept value.
// An element's fp32 distance lies within BW_MARGIN/2 of its tensor-core distance, so the order of two
// elements whose tensor-core values differ by more than BW_MARGIN is already the fp32 order: only the
// kept elements within BW_MARGIN of t_k are recomputed with fp32 FMAs and ranked -- the result is the
// exact fp32 order statistic, independent of the tensor-core rounding, at a tail cost that does not
// grow with the bin population.  Round 1 used two uniform levels (1/64, then 2^-14) = three full Gram
// passes per bandwidth; the logarithmic level makes it two (98 + 114 + 107 us -> see DESIGN.md).
// Rows with too many candidates (massive duplicates) raise the overflow flag and the exact CUDA-core
// kernel (bandwidth.cu) redoes the batch.
#include <cudaTypedefs.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include <type_traits>
#include "common.cuh"
#include "sm100_ptx.cuh"

using namespace sm100;

int prifit_tc_make_tile_map(CUtensorMap* map, const __half* X, int B, int N);

namespace {

constexpr int G_D = 128, G_BM = 128, G_BN = 128;
constexpr int G_THREADS = 640;                           // warps 0-2: TMA / MMA / TMEM alloc; warps 4-19: epilogue
constexpr int G_EPI = 512;                               // epilogue threads: (row, 32-column quarter of every tile)
constexpr int G_STAGES = 2;
constexpr uint32_t G_HALF_BYTES = G_BN * G_D * 2;        // one fp16 tile (hi or lo): 32 KB
constexpr uint32_t G_STAGE_BYTES = 2 * G_HALF_BYTES;     // hi + lo
constexpr uint32_t G_KBLOCK = G_BN * 128;                // one 64-column (128 B) block of a tile
constexpr uint32_t GCOL_QHI = 256, GCOL_QLO = 320;
constexpr float G_PRESCALE = 256.0f;                     // operands carry x * 2^8  ->  S = 2^16 <a, b>
constexpr float G_DIST_MUL = -2.0f / 65536.0f;
enum { GM_NEAREST = 0, GM_BEST = 1, GM_HIST = 2, GM_COLLECT = 3, GM_DUMP = 4 };

constexpr int HIST_BINS = 256;
constexpr int HIST_WORDS = HIST_BINS / 2 + 1;            // packed uint16 pairs; odd word stride: conflict-free rows
// bandwidth passes work on u = dist / 4 + 2^-15 in [~2^-15, 1 + 2^-15]: one FFMA, always positive and normal, so the bit
// pattern orders like the value and its top bits are a logarithmic bin index
constexpr float BW_UOFF = 1.0f / 32768.0f;
constexpr uint32_t LOG_SHIFT = 19;                       // 4 mantissa bits: 16 bins per binade, 6.25 % wide
constexpr uint32_t LOG_BASE = 111u << 4;                 // first bin starts at u = 2^-16 (biased exponent 111): 16 binades up to u = 1
constexpr uint32_t LOG_LO_BITS = LOG_BASE << LOG_SHIFT;  // bit pattern of 2^-16
constexpr uint32_t LOG_PAD = (uint32_t)(HIST_BINS / 2) << (LOG_SHIFT + 1);   // (bits - LOG_LO_BITS) of u = 1: word 128 = the pad word
constexpr float BW_MARGIN = 2.0e-5f;                     // >= 2 x the bound on |tensor-core - fp32| distance
constexpr float BW_MARGIN_U = 0.25f * BW_MARGIN;         // the same in u units
constexpr int CAND_ROW = 64;                             // candidates per row (COLLECT)
constexpr int REFINE_ABOVE = 48;                         // level-0 bin population above which level 1 refines the row

template <int MODE> struct GCfg {
    static constexpr size_t scratch =
        MODE == GM_NEAREST ? (size_t)3 * G_BM * 8
      : MODE == GM_BEST ? (size_t)3 * G_BM * 8 + 2 * G_BN * sizeof(float)
      : MODE == GM_HIST ? (size_t)G_BM * HIST_WORDS * 4
      : MODE == GM_COLLECT ? (size_t)G_BM * CAND_ROW * (4 + 2) + (size_t)G_BM * sizeof(int) + (size_t)G_EPI * sizeof(int)
                             + (size_t)16 * 4 * CAND_ROW * sizeof(float)
      : 16;
    static constexpr size_t smem = 1024 + (size_t)G_STAGES * G_STAGE_BYTES + 256 + scratch;
};

struct GBars {
    uint64_t x_full[G_STAGES], x_empty[G_STAGES], s_full[2], s_free[2], q_full;
    uint32_t tmem_base;
};

struct GramArgs {
    const __half* Xs;        // [2B,N,128] split rows: shapes [0,B) = hi, [B,2B) = lo   (A and B operands)
    const float* X32;        // [B,N,128] fp32 rows (COLLECT: exact recompute)
    const float* bw;         // [B]       (BEST)
    const int32_t* votes;    // [B,N]     (BEST)
    const int32_t* rowsel;   // [B,N]     (BEST) compacted list of the rows to process, ascending
    const int32_t* nrows;    // [B]       (BEST) length of that list
    int small_rows;          //           (BEST) shapes with at most this many voted rows are skipped (done by the small kernel)
    const int32_t* kth;      // [B]       (HIST / COLLECT), 1-based rank
    int32_t* out_idx;        // [B,N]     (NEAREST / BEST)
    int2* rowinfo;           // [B,N]     (HIST in/out, COLLECT in): x = window start (bits of u), y = remaining rank (24 bits)
                             //           | width code c << 24 (window width 2^-c in u units) | bit 31: level 1 must refine the row
    float* rowval;           // [B,N]     (COLLECT out)
    int32_t* overflow;       // [2]       [0] (COLLECT out): candidate list overflow / window miss; [1] (HIST level 0 out): some
                             //           row needs level 1
    float* dump;             // [B,N,N]   (DUMP out)
    int N, B, level;
    int dbg;                 // timing experiments only (PRIFIT_GRAM_DEBUG): 1 = hi.hi product only
};

__device__ __forceinline__ void epi_barrier() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

// The epilogues work on q = dist / 4 = saturate(0.5 - <a, b> / 2) in [0, 1]: one FFMA.SAT on the FMA pipe instead of an
// FFMA and two FMNMX on the half-rate ALU pipe, which is what bounds these kernels (ncu: ALU pipe, not issue slots, not
// the tensor pipe).  Scaling by a power of two commutes with every rounding involved, so q * 4 is bit-identical to
// clamp(2 - 2 <a, b>, 0, 4) and every comparison / bin below equals the one in distance units.  Same expression in every
// pass, so the bins are consistent.
__device__ __forceinline__ float tc_q(uint32_t sbits) {
    return __saturatef(fmaf(__uint_as_float(sbits), 0.25f * G_DIST_MUL, 0.5f));
}
// bandwidth passes: u = dist / 4 + 2^-15, NOT saturated (a dot product that rounds to 1 + 1e-6 still gives u > 2^-16)
__device__ __forceinline__ float tc_u(uint32_t sbits) {
    return fmaf(__uint_as_float(sbits), 0.25f * G_DIST_MUL, 0.5f + BW_UOFF);
}

template <int MODE>
__global__ void __launch_bounds__(G_THREADS, 1) gram_tc_kernel(const __grid_constant__ CUtensorMap tmap, const GramArgs a) {
    const int b = blockIdx.y, r0 = blockIdx.x * G_BM, N = a.N;
    int nsel = N;                                              // rows this shape contributes (BEST: voted rows only)
    if (MODE == GM_BEST) nsel = a.nrows[b];
    if (r0 >= nsel) return;                                    // uniform over the CTA, before any barrier / allocation
    if (MODE == GM_BEST && nsel <= a.small_rows) return;       // few voted rows: nms_best_small_kernel (nms.cu) has done this shape
    if (MODE == GM_HIST && a.level > 0 && a.overflow[1] == 0) return;   // no row of the batch asked for the refinement level

    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment as an offset on the __shared__ symbol (not through an integer round trip), so that every scratch
    // access below is known to be in shared memory: LDS / STS / ATOMS instead of generic LD.E / ST.E / ATOM.E
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* tiles = smem;
    GBars* bars = reinterpret_cast<GBars*>(smem + (size_t)G_STAGES * G_STAGE_BYTES);
    uint8_t* scratch = smem + (size_t)G_STAGES * G_STAGE_BYTES + 256;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nt = (N + G_BN - 1) / G_BN;

    if (threadIdx.x == 0) {
        for (int s = 0; s < G_STAGES; ++s) { mbar_init(&bars->x_full[s], 1); mbar_init(&bars->x_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&bars->s_full[s], 1); mbar_init(&bars->s_free[s], G_EPI); }
        mbar_init(&bars->q_full, G_EPI);
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
            for (int j = 0; j < nt; ++j) {
                const uint32_t st = j % G_STAGES, ph = (j / G_STAGES) & 1;
                mbar_wait(&bars->x_empty[st], ph ^ 1);
                mbar_arrive_expect_tx(&bars->x_full[st], G_STAGE_BYTES);
                const uint32_t dst = smem_u32(tiles + (size_t)st * G_STAGE_BYTES);
                tma_load_3d(dst, &tmap, &bars->x_full[st], 0, j * G_BN, b);
                tma_load_3d(dst + G_KBLOCK, &tmap, &bars->x_full[st], 64, j * G_BN, b);
                tma_load_3d(dst + G_HALF_BYTES, &tmap, &bars->x_full[st], 0, j * G_BN, a.B + b);
                tma_load_3d(dst + G_HALF_BYTES + G_KBLOCK, &tmap, &bars->x_full[st], 64, j * G_BN, a.B + b);
            }
        }
    } else if (warp == 1) {
        // ================================= MMA issuer =================================
        if (lane == 0) {
            constexpr uint32_t idesc = idesc_f16(G_BM, G_BN, false);
            mbar_wait(&bars->q_full, 0);
            tc_fence_after();
            for (int j = 0; j < nt; ++j) {
                const uint32_t st = j % G_STAGES, xph = (j / G_STAGES) & 1, buf = j & 1;
                mbar_wait(&bars->x_full[st], xph);
                mbar_wait(&bars->s_free[buf], ((j >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t base = smem_u32(tiles + (size_t)st * G_STAGE_BYTES);
                const uint32_t acc = tmem + buf * 128;
#pragma unroll
                for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint32_t qo = kb * 32 + ks * 8;
                        const uint64_t bhi = smem_desc_sw128(base + kb * G_KBLOCK + ks * 32, 16, 1024);
                        const uint64_t blo = smem_desc_sw128(base + G_HALF_BYTES + kb * G_KBLOCK + ks * 32, 16, 1024);
                        if (!(a.dbg & 1)) {
                            mma_f16_ts(acc, tmem + GCOL_QLO + qo, bhi, idesc, (kb | ks) != 0);
                            mma_f16_ts(acc, tmem + GCOL_QHI + qo, blo, idesc, true);
                        }
                        mma_f16_ts(acc, tmem + GCOL_QHI + qo, bhi, idesc, (a.dbg & 1) ? (kb | ks) != 0 : true);
                    }
                mma_commit(&bars->s_full[buf]);
                mma_commit(&bars->x_empty[st]);
            }
        }
    } else if (warp >= 4) {
        // ================================== epilogue ==================================
        // Four warps per SM sub-partition (thread = row x 32-column quarter): the per-element work is a short
        // dependent ALU chain, so latency hiding comes from warps, not from ILP.
        const int et = threadIdx.x - 128;                        // 0..511
        const int ew = warp - 4;                                 // 0..15
        const int qt = ew >> 2;                                  // column quarter of every tile this thread owns
        const int row = 32 * (ew & 3) + lane;                    // TMEM lane == row within the CTA tile
        const uint32_t lane_base = (uint32_t)(32 * (ew & 3)) << 16;
        const bool row_ok = r0 + row < nsel;
        int grow_i = row_ok ? r0 + row : 0;                      // row index inside the shape
        if (MODE == GM_BEST) grow_i = row_ok ? a.rowsel[(size_t)b * N + r0 + row] : 0;
        const size_t grow = (size_t)b * N + grow_i;

        // ---- A operands: this thread's 32-element quarter of the row, hi and lo, into tensor memory
        {
            uint32_t h[16];
            const uint4* xhi = reinterpret_cast<const uint4*>(a.Xs + grow * G_D) + 4 * qt;
            const uint4* xlo = reinterpret_cast<const uint4*>(a.Xs + ((size_t)a.B * N + grow) * G_D) + 4 * qt;
#pragma unroll
            for (int part = 0; part < 2; ++part) {
                const uint4* src = part == 0 ? xhi : xlo;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const uint4 f = row_ok ? src[e] : make_uint4(0u, 0u, 0u, 0u);
                    h[4 * e] = f.x; h[4 * e + 1] = f.y; h[4 * e + 2] = f.z; h[4 * e + 3] = f.w;
                }
                tmem_st16(tmem + lane_base + (part == 0 ? GCOL_QHI : GCOL_QLO) + 16 * qt, h);
            }
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive(&bars->q_full);
        }

        // ---- per-mode state
        float best = MODE == GM_NEAREST ? INFINITY : -1.0f;
        int besti = 0x7fffffff;
        float bwv = 0.f;
        float* cmb_v = reinterpret_cast<float*>(scratch);                                  // NEAREST/BEST: [3][128]
        int* cmb_i = reinterpret_cast<int*>(scratch) + 3 * G_BM;                           //               [3][128]
        float* vt = reinterpret_cast<float*>(scratch + (size_t)3 * G_BM * 8);              // BEST: [2][128]
        uint32_t* hist = reinterpret_cast<uint32_t*>(scratch) + (size_t)row * HIST_WORDS;  // HIST: own row
        const uint32_t hist_s = smem_u32(hist);
        // COLLECT scratch: values [128][64] u32 | columns [128][64] u16 | per-row counters [128] | per-thread below [512] | tail values
        uint32_t* cval = reinterpret_cast<uint32_t*>(scratch) + (size_t)row * CAND_ROW;
        uint16_t* ccol = reinterpret_cast<uint16_t*>(scratch + (size_t)G_BM * CAND_ROW * 4) + (size_t)row * CAND_ROW;
        int* ccnt_all = reinterpret_cast<int*>(scratch + (size_t)G_BM * CAND_ROW * 6);
        int* below_s = ccnt_all + G_BM;
        float win_lo = 0.f, win_hi = 0.f, hscale4 = 0.f;
        int below = 0, krem = 1, wcode = 0;
        bool refine_row = false;
        float vnext = 0.f;
        if (MODE == GM_BEST) {
            bwv = 0.25f * a.bw[b];                                 // threshold in q units
            vnext = et < G_BN && et < N ? (float)a.votes[(size_t)b * N + et] : 0.f;
        }
        if (MODE == GM_HIST) {
            for (int q = qt; q < HIST_WORDS; q += 4) hist[q] = 0u;
            krem = max(1, min(a.kth[b], N));
            if (a.level > 0 && row_ok) {
                const int2 ri = a.rowinfo[grow];
                refine_row = ri.y < 0;                                // bit 31
                if (refine_row) {
                    win_lo = __int_as_float(ri.x);                    // window start (u units), width 2^-wcode
                    wcode = (ri.y >> 24) & 0x7f;
                    krem = ri.y & 0xffffff;
                    hscale4 = ldexpf(256.0f, wcode);                  // 256 uniform sub-bins
                }
            }
            epi_barrier();
        }
        if (MODE == GM_COLLECT) {
            if (qt == 0) ccnt_all[row] = 0;
            if (row_ok) {
                const int2 ri = a.rowinfo[grow];
                const float lo_u = __int_as_float(ri.x);
                win_lo = lo_u - BW_MARGIN_U;
                win_hi = lo_u + ldexpf(1.0f, -((ri.y >> 24) & 0x7f)) + BW_MARGIN_U;
            }
            epi_barrier();
        }

        const uint32_t lo_bits = __float_as_uint(fmaxf(win_lo, 0.f));                 // COLLECT: window as bit patterns of u
        const uint32_t win_bits = __float_as_uint(fmaxf(win_hi, 0.f)) - lo_bits;
        const bool level0 = a.level == 0;
        int* ccnt = ccnt_all + row;
        for (int j = 0; j < nt; ++j) {
            const uint32_t buf = j & 1, ph = (j >> 1) & 1;
            const int key0 = j * G_BN;
            if (MODE == GM_BEST) {
                if (et < G_BN) vt[buf * G_BN + et] = vnext;          // votes of this tile's columns (prefetched)
                epi_barrier();
                const int col = key0 + G_BN + et;                    // prefetch the next tile's votes
                vnext = et < G_BN && col < N ? (float)a.votes[(size_t)b * N + col] : 0.f;
            }
            mbar_wait(&bars->s_full[buf], ph);
            tc_fence_after();
            uint32_t v[32];
            tmem_ld32(tmem + lane_base + buf * 128 + 32 * qt, v);
            tmem_wait_ld();
            tc_fence_before();
            mbar_arrive(&bars->s_free[buf]);                         // S is in registers: the MMA warp may refill it
            const int ncols = min(G_BN, N - key0) - 32 * qt;         // valid columns among this thread's 32
            // per-element epilogue; CHK = false on full tiles (every tile but possibly the last) drops the column test
            auto consume = [&](auto chk, auto lv0) {
                constexpr bool CHK = decltype(chk)::value;
                constexpr bool LEVEL0 = decltype(lv0)::value;            // HIST only: logarithmic (true) / uniform sub-bins (false)
#pragma unroll
                for (int e = 0; e < 32; ++e) {
                    if (CHK && e >= ncols) break;
                    const float dist = (MODE == GM_HIST || MODE == GM_COLLECT) ? tc_u(v[e]) : tc_q(v[e]);   // distance / 4 (+ 2^-15)
                    const int col = key0 + 32 * qt + e;
                    if (MODE == GM_NEAREST) {
                        if (dist < best) { best = dist; besti = col; }
                    } else if (MODE == GM_BEST) {
                        const float val = dist < bwv ? vt[buf * G_BN + 32 * qt + e] : 0.f;
                        if (val > best) { best = val; besti = col; }
                    } else if (MODE == GM_HIST) {
                        // One predicated shared-memory atomic per element, no divergent region (a conditional update compiles to a
                        // BSSY / BRA / BSYNC region per element and the C++ atomicAdd on a generic pointer to a generic ATOM: 19.5
                        // instructions per element).  Elements outside [0, 256) -- u >= 1 at level 0, anything outside the
                        // row's window at level 1 (negative values wrap to huge unsigned ones) -- update nothing.
                        //   level 0: bin = top bits of u (logarithmic, 16 bins per binade)
                        //   level 1: bin = floor((u - lo) * 256 / width) through a round-down FMA onto 2^23 (FMA pipe; F2I is an
                        //            8-cycle XU instruction): the low mantissa bits of floor(y) + 2^23 are floor(y) for 0 <= y < 2^22
                        if (LEVEL0) {
                            // t = (bits of u) - (bits of 2^-16), clamped to [0, 128 << 20]: word t >> 20 of the OWN row -- 0..127 are the
                            // 256 bins (16 per binade, half-word picked by bit 19), 128 is the row's pad word, which takes u >= 1 and
                            // whatever a non-unit input row could produce (the wrap of u < 2^-16 included).  The atomic is therefore
                            // UNCONDITIONAL: a predicated one compiles to a BSSY / BRA / BSYNC region per element around the address
                            // arithmetic (12.5 instructions per element, no overlap between elements; now 9 and straight-line).
                            const uint32_t t = min(__float_as_uint(dist) - LOG_LO_BITS, LOG_PAD);
                            asm volatile("red.shared.add.u32 [%0], %1;" :: "r"(hist_s + ((t >> 20) << 2)), "r"(1u << ((t >> 15) & 16u)) : "memory");
                        } else {
                            const float biased = __fmaf_rd(dist - win_lo, hscale4, 8388608.0f);
                            const uint32_t ub = refine_row ? __float_as_uint(biased) - 0x4b000000u : 0xffffffffu;
                            const uint32_t sh = (ub & 1u) << 4;
                            asm volatile("{\n\t.reg .pred p;\n\tsetp.lt.u32 p, %0, 256;\n\t@p red.shared.add.u32 [%1], %2;\n\t}"
                                         :: "r"(ub), "r"(hist_s + ((ub >> 1) << 2)), "r"(1u << sh) : "memory");
                        }
                    } else if (MODE == GM_COLLECT) {
                        // q >= 0, so its bit pattern orders like the value: two integer compares instead of three NaN-aware
                        // float compares on the (half-rate) ALU pipe
                        // (both patterns are below 2^31, so the wrapped difference has its top bit set exactly when qb < lo_bits:
                        // one subtraction serves the count and the window test, and the count is an add, not a select chain)
                        const uint32_t qb = __float_as_uint(dist);
                        const uint32_t dlo = qb - lo_bits;
                        below += (int)(dlo >> 31);
                        if (dlo <= win_bits) {                           // rare: one list per row, slots handed out atomically
                            const int slot = atomicAdd(ccnt, 1);
                            if (slot < CAND_ROW) { cval[slot] = qb; ccol[slot] = (uint16_t)col; }
                        }
                    } else if (row_ok) {
                        a.dump[grow * N + col] = 4.0f * dist;
                    }
                }
            };
            if (MODE == GM_HIST && level0) {
                if (ncols >= 32) consume(std::false_type{}, std::true_type{}); else consume(std::true_type{}, std::true_type{});
            } else {
                if (ncols >= 32) consume(std::false_type{}, std::false_type{}); else consume(std::true_type{}, std::false_type{});
            }
        }

        // ---- per-mode finalisation
        if (MODE == GM_NEAREST || MODE == GM_BEST) {
            // combine the four column quarters of every row: better value, then lower index
            if (qt > 0) { cmb_v[(qt - 1) * G_BM + row] = best; cmb_i[(qt - 1) * G_BM + row] = besti; }
            epi_barrier();
            if (qt == 0 && row_ok) {
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    const float ov = cmb_v[q * G_BM + row];
                    const int oi = cmb_i[q * G_BM + row];
                    const bool take = MODE == GM_NEAREST ? (ov < best || (ov == best && oi < besti))
                                                         : (ov > best || (ov == best && oi < besti));
                    if (take) { best = ov; besti = oi; }
                }
                a.out_idx[grow] = besti;
            }
        } else if (MODE == GM_HIST) {
            epi_barrier();
            if (qt == 0 && row_ok && (level0 || refine_row)) {
                int cum = 0, bin = HIST_BINS - 1, before = 0, cnt = 0;
                bool found = false;
                for (int q = 0; q < HIST_BINS / 2 && !found; ++q) {
                    const uint32_t w = hist[q];
                    const int c0 = (int)(w & 0xffffu), c1 = (int)(w >> 16);
                    if (cum + c0 >= krem) { bin = 2 * q; before = cum; cnt = c0; found = true; }
                    else if (cum + c0 + c1 >= krem) { bin = 2 * q + 1; before = cum + c0; cnt = c1; found = true; }
                    cum += c0 + c1;
                }
                // not found: the k-th smallest lies among the elements with u >= 1 (distance 4, antipodal rows) at level 0;
                // COLLECT then finds no consistent rank in the last bin's window and raises the overflow flag
                if (!found) before = max(0, krem - 1);
                if (level0) {
                    const uint32_t lo_u = (LOG_BASE + (uint32_t)bin) << LOG_SHIFT;            // bits of the bin's lower edge
                    const int code = 131 - (int)(lo_u >> 23);                                  // bin width 2^(E - 127 - 4) = 2^-code
                    const bool mark = found && cnt > REFINE_ABOVE;
                    if (mark) atomicOr(a.overflow + 1, 1);
                    a.rowinfo[grow] = make_int2((int)lo_u, (krem - before) | (code << 24) | (mark ? (int)0x80000000u : 0));
                } else {
                    const float lo_new = win_lo + ldexpf((float)bin, -(wcode + 8));           // exact: a multiple of the sub-bin width
                    a.rowinfo[grow] = make_int2(__float_as_int(lo_new), (krem - before) | ((wcode + 8) << 24));
                }
            }
        } else if (MODE == GM_COLLECT) {
            // Tail.  Every epilogue warp owns 8 rows and works on 4 of them at a time, 8 lanes per row.
            //  (1) t_k = the (k - below)-th smallest kept tensor-core value (rank counting over <= 64 values);
            //  (2) the kept elements within BW_MARGIN_U of t_k are the only ones whose fp32 order can differ from their
            //      tensor-core order: L = number of kept elements below that band, the band's columns are compacted;
            //  (3) exact fp32 recompute of the band (lane l8 of a group owns dims 16 l8 .. 16 l8 + 15, 4 candidates in
            //      flight) and the (k - below - L)-th smallest of those = the exact fp32 order statistic.
            below_s[row * 4 + qt] = below;
            epi_barrier();
            float* vals_w = reinterpret_cast<float*>(scratch + (size_t)G_BM * CAND_ROW * 6 + (size_t)(G_BM + G_EPI) * sizeof(int)) + ew * 4 * CAND_ROW;
            const int k = max(1, min(a.kth[b], N));
            const int grp = lane >> 3, l8 = lane & 7;
            float* vals = vals_w + grp * CAND_ROW;
            for (int rr = 0; rr < 2; ++rr) {
                const int r = 8 * ew + 4 * rr + grp;
                const bool in_range = r0 + r < N;
                uint32_t* rv = reinterpret_cast<uint32_t*>(scratch) + (size_t)r * CAND_ROW;
                uint16_t* rc = reinterpret_cast<uint16_t*>(scratch + (size_t)G_BM * CAND_ROW * 4) + (size_t)r * CAND_ROW;
                int nc = 0, bel = 0;
                bool over = false;
                if (in_range) {
                    nc = ccnt_all[r];
                    over = nc > CAND_ROW;
                    bel = (below_s[4 * r] + below_s[4 * r + 1]) + (below_s[4 * r + 2] + below_s[4 * r + 3]);
                }
                const int m = k - bel;                                          // m-th smallest kept element (1-based)
                const bool bad = in_range && (over || m < 1 || m > nc);
                if (bad && l8 == 0) { atomicExch(a.overflow, 1); a.rowval[(size_t)b * N + r0 + r] = 0.f; }
                if (!in_range || bad) nc = 0;
                // (1) t_k on the tensor-core values (bit patterns order like the values; ties broken by list position)
                uint32_t tk = 0u;
                bool have = false;
                for (int ci = l8; ci < nc; ci += 8) {
                    const uint32_t vi = rv[ci];
                    int rank = 0;
                    for (int cj = 0; cj < nc; ++cj) {
                        const uint32_t vj = rv[cj];
                        rank += (vj < vi || (vj == vi && cj < ci)) ? 1 : 0;
                    }
                    if (rank == m - 1) { tk = vi; have = true; }
                }
                unsigned ball = (__ballot_sync(0xffffffffu, have) >> (8 * grp)) & 0xffu;
                tk = __shfl_sync(0xffffffffu, tk, ball ? 8 * grp + __ffs(ball) - 1 : lane);       // every lane takes part
                // (2) band around t_k: lane 0 of the group compacts its columns to the front of the row's column list
                int ns = 0, L = 0;
                if (l8 == 0 && nc > 0) {
                    const float tkf = __uint_as_float(tk);
                    const uint32_t band_lo = __float_as_uint(fmaxf(tkf - BW_MARGIN_U, 0.f)), band_hi = __float_as_uint(tkf + BW_MARGIN_U);
                    for (int ci = 0; ci < nc; ++ci) {
                        const uint32_t vi = rv[ci];
                        const uint16_t cc = rc[ci];
                        if (vi < band_lo) ++L;
                        else if (vi <= band_hi) rc[ns++] = cc;                  // ns <= ci: never overwrites an unread entry
                    }
                }
                ns = __shfl_sync(0xffffffffu, ns, 8 * grp);
                L = __shfl_sync(0xffffffffu, L, 8 * grp);
                __syncwarp();
                const int mm = m - L;                                           // rank inside the band (1-based)
                // (3) exact fp32 distances of the band
                float4 xr[4];
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    xr[e] = ns > 0 ? __ldg(reinterpret_cast<const float4*>(a.X32 + ((size_t)b * N + r0 + r) * G_D) + 4 * l8 + e)
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
                int nsmax = ns;                                                 // warp-uniform trip count
#pragma unroll
                for (int o = 8; o < 32; o <<= 1) nsmax = max(nsmax, __shfl_xor_sync(0xffffffffu, nsmax, o));
                for (int c0 = 0; c0 < nsmax; c0 += 4) {
                    float acc[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        acc[u] = 0.f;
                        if (c0 < ns) {
                            const float4* wrow = reinterpret_cast<const float4*>(a.X32 + ((size_t)b * N + rc[min(c0 + u, ns - 1)]) * G_D) + 4 * l8;
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float4 w = __ldg(wrow + e);
                                acc[u] = fmaf(xr[e].x, w.x, fmaf(xr[e].y, w.y, fmaf(xr[e].z, w.z, fmaf(xr[e].w, w.w, acc[u]))));
                            }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u)
#pragma unroll
                        for (int o = 4; o > 0; o >>= 1) acc[u] += __shfl_xor_sync(0xffffffffu, acc[u], o);
                    if (l8 < 4 && c0 + l8 < ns)
                        vals[c0 + l8] = 2.0f - 2.0f * (l8 == 0 ? acc[0] : l8 == 1 ? acc[1] : l8 == 2 ? acc[2] : acc[3]);
                }
                __syncwarp();
                float found = 0.f;
                have = false;
                for (int ci = l8; ci < ns; ci += 8) {
                    const float vi = vals[ci];
                    int rank = 0;
                    for (int cj = 0; cj < ns; ++cj) {
                        const float vj = vals[cj];
                        rank += (vj < vi || (vj == vi && cj < ci)) ? 1 : 0;
                    }
                    if (rank == mm - 1) { found = vi; have = true; }
                }
                ball = (__ballot_sync(0xffffffffu, have) >> (8 * grp)) & 0xffu;
                const float res = __shfl_sync(0xffffffffu, found, ball ? 8 * grp + __ffs(ball) - 1 : lane);   // every lane takes part
                if (nc > 0 && l8 == 0) {
                    if (ball) a.rowval[(size_t)b * N + r0 + r] = sqrtf(fmaxf(res, 1e-6f));           // guard_sqrt(., 1e-6)
                    else atomicExch(a.overflow, 1);
                }
                __syncwarp();
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// x * 2^8 = hi + lo (both fp16, round to nearest); hi -> Xs[0 .. n), lo -> Xs[n .. 2n)
__global__ void split_half_kernel(const float4* __restrict__ in, uint2* __restrict__ hi, uint2* __restrict__ lo, size_t n4) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = in[i];
        const float s[4] = {v.x * G_PRESCALE, v.y * G_PRESCALE, v.z * G_PRESCALE, v.w * G_PRESCALE};
        float r[4];
        uint32_t hw[2], lw[2];
#pragma unroll
        for (int e = 0; e < 4; ++e) r[e] = s[e] - __half2float(__float2half_rn(s[e]));
        hw[0] = pack_f16x2(s[0], s[1]); hw[1] = pack_f16x2(s[2], s[3]);
        lw[0] = pack_f16x2(r[0], r[1]); lw[1] = pack_f16x2(r[2], r[3]);
        hi[i] = make_uint2(hw[0], hw[1]);
        lo[i] = make_uint2(lw[0], lw[1]);
    }
}

template <int MODE>
int launch_gram(const CUtensorMap& map, const GramArgs& a, cudaStream_t st) {
    PF_CUDA(cudaFuncSetAttribute(gram_tc_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GCfg<MODE>::smem));
    dim3 grid((a.N + G_BM - 1) / G_BM, a.B);
    GramArgs aa = a;
    if (const char* e = getenv("PRIFIT_GRAM_DEBUG")) aa.dbg = atoi(e);
    gram_tc_kernel<MODE><<<grid, G_THREADS, GCfg<MODE>::smem, st>>>(map, aa);
    PF_LAUNCH_CHECK();
    return 0;
}

int split_rows(const float* X, __half* Xs, int B, int N, CUtensorMap* map, cudaStream_t st) {
    const size_t n = (size_t)B * N * G_D, n4 = n / 4;
    split_half_kernel<<<(unsigned)min((size_t)148 * 8, (n4 + 255) / 256), 256, 0, st>>>(
        reinterpret_cast<const float4*>(X), reinterpret_cast<uint2*>(Xs), reinterpret_cast<uint2*>(Xs + n), n4);
    PF_LAUNCH_CHECK();
    return prifit_tc_make_tile_map(map, Xs, 2 * B, N);
}

}  // namespace

// shared with meanshift_rows_tc.cu: x * 2^8 = hi + lo rows and the [2B][N][128] tile map over them
int prifit_tc_split_rows(const float* X, __half* Xs, int B, int N, CUtensorMap* map, cudaStream_t st) {
    return split_rows(X, Xs, B, N, map, st);
}

// Xs_ws scratch for every entry point below: 2 * B * N * 128 halves.
size_t prifit_tc_gram_split_bytes(int B, int N) { return (size_t)2 * B * N * G_D * sizeof(__half); }

// ---- NMS steps 1 and 3 on the tensor cores (called from nms.cu)
int prifit_tc_nms_nearest(const float* newX, int B, int N, __half* Xs_ws, CUtensorMap* map_out, int32_t* nearest, cudaStream_t st) {
    int rc = split_rows(newX, Xs_ws, B, N, map_out, st);
    if (rc) return rc;
    GramArgs a = {};
    a.Xs = Xs_ws; a.N = N; a.B = B; a.out_idx = nearest;
    return launch_gram<GM_NEAREST>(*map_out, a, st);
}

// best[b, i] is written only for the rows listed in rowsel[b, 0 .. nrows[b])
int prifit_tc_nms_best(const CUtensorMap* map, const __half* Xs, const float* bw, const int32_t* votes,
                       const int32_t* rowsel, const int32_t* nrows, int B, int N, int small_rows, int32_t* best, cudaStream_t st) {
    GramArgs a = {};
    a.Xs = Xs; a.N = N; a.B = B; a.bw = bw; a.votes = votes; a.rowsel = rowsel; a.nrows = nrows; a.out_idx = best;
    a.small_rows = small_rows;
    return launch_gram<GM_BEST>(*map, a, st);
}

// ---- bandwidth order statistic on the tensor cores (called from bandwidth.cu)
int prifit_tc_bandwidth_rows(const float* X, int B, int N, const int32_t* kth, __half* Xs_ws, int2* rowinfo_ws,
                             float* rowval, int32_t* overflow, cudaStream_t st) {
    CUtensorMap map;
    int rc = split_rows(X, Xs_ws, B, N, &map, st);
    if (rc) return rc;
    GramArgs a = {};
    a.Xs = Xs_ws; a.X32 = X; a.N = N; a.B = B; a.kth = kth; a.rowinfo = rowinfo_ws; a.rowval = rowval; a.overflow = overflow;
    a.level = 0;                                   // logarithmic bins: one pass locates the k-th to 6 %
    rc = launch_gram<GM_HIST>(map, a, st);
    if (rc) return rc;
    a.level = 1;                                   // every CTA returns at once unless level 0 set overflow[1] (a crowded bin)
    rc = launch_gram<GM_HIST>(map, a, st);
    if (rc) return rc;
    return launch_gram<GM_COLLECT>(map, a, st);
}

// diagnostics (tests): dist[B,N,N] = the tensor-core distance matrix.  ws >= prifit_tc_gram_split_bytes(B, N) + 256.
extern "C" int prifit_debug_tc_gram(const float* X, int B, int N, float* dist_out, void* ws, void* stream) {
    PF_CHECK_ARG(X && dist_out && ws, PRIFIT_E_BADARG, "null pointer");
    PF_CHECK_ARG(B > 0 && N > 0, PRIFIT_E_BADARG, "B, N > 0 required");
    __half* Xs = reinterpret_cast<__half*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
    CUtensorMap map;
    int rc = split_rows(X, Xs, B, N, &map, pf_stream(stream));
    if (rc) return rc;
    GramArgs a = {};
    a.Xs = Xs; a.N = N; a.B = B; a.dump = dist_out;
    return launch_gram<GM_DUMP>(map, a, pf_stream(stream));
}
