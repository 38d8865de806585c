// k2 rows, tensor-core engine -- trajectories of the K selected seeds (forward) and the reverse sweep
// through them (backward) on tcgen05 / TMEM / TMA with fp32-class operands.
//
// reference: center = new_X[indices] (src/mean_shift.py:46) over mean_shift_ (src/mean_shift.py:50-84).
// Same mathematics and the same saved tensors (traj, stat) as the CUDA-core kernels of
// meanshift_rows.cu, which remain as the cross-check engine.
//
// Layout of the problem.  K <= 64 seeds against N keys is a "few rows x many keys" contraction, so
// the MMAs run TRANSPOSED: keys are the M dimension (128 per tile = TMEM lanes), a group of 32 seeds
// is the N dimension, and the KEYS of a shape are split over a thread-block cluster (<= 8 CTAs);
// per-iteration partial sums (32 x d) are reduced through distributed shared memory in fixed rank
// order.  One epilogue thread owns one key (and 16 of the 32 seeds).
//
//   forward, per key tile     S^T[128 keys x 32]  = Xtile . Y^T              (A smem K-major, B smem K-major)
//                             P^T = exp(clamp(-(2 - 2 S)/b^2/2))             -> smem tile [key][seed]
//                             O^T[d x 32]        += Xtile^T . P^T            (A smem MN-major, B smem MN-major)
//   backward, per key tile    S1^T = Xtile . Y^T ,  S2^T = Xtile . Gm^T
//                             ds, e1 (see meanshift_rows.cu)                  -> smem tile [key][ds | e1]
//                             gy^T[d x 32]       += Xtile^T . ds^T
//                             gX[128 keys x d]    = [ds | e1] . [Y ; Gm]     (A smem K-major, B smem MN-major)
//
// Precision.  Every operand is split into two fp16 numbers (hi + lo, 22 significant bits) after a
// power-of-two pre-scale that keeps both halves normal, and every product is the three partial products
// lo.hi + hi.lo + hi.hi accumulated in fp32 -- the scheme of the Gram engine (gram_tc.cu), |error| ~ 3e-7
// on a dot product of unit vectors, i.e. fp32-class.  The MMAs are bound by the shared-memory fetch of
// the 128-row key tile (~80 clk per instruction whatever N is), so the hi and lo halves of the SMALL
// operand are stacked along N: one MMA of the key tile's hi half against [b_hi ; b_lo] (two accumulator
// column groups, summed by the epilogue) plus one of its lo half against b_hi -- two key-tile fetches
// per K step instead of three.  Scales:
//   x, y      * 2^8                       (unit vectors)
//   p         * 2^14                      (p in [e^-13, 1])
//   gm        * SG  (power of two, max |gm| of the shape and iteration -> [2^7, 2^8))
//   ds        * SD  (power of two from the bound |ds| <= 2 ||gm|| D / b^2 -> < 2^14)
//   e1        * SE = SD 2^8 / SG          (so both halves of the K = 64 contraction share one scale)
#include <cudaTypedefs.h>
#include <cuda_fp16.h>
#include <cooperative_groups.h>
#include <stdlib.h>
#include "common.cuh"
#include "sm100_ptx.cuh"

namespace cg = cooperative_groups;
using namespace sm100;

int prifit_tc_make_tile_map(CUtensorMap* map, const __half* X, int B, int N);
int prifit_tc_split_rows(const float* X, __half* Xs, int B, int N, CUtensorMap* map, cudaStream_t st);

namespace {

constexpr int RT_D = 128, RT_KEYS = 128, RT_SEEDS = 32;
constexpr int RT_THREADS = 384;                            // warps 0-3: TMA / MMA / TMEM alloc / spare, warps 4-11: epilogue
constexpr int RT_EPI = 256;
constexpr int RT_STAGES = 2;
constexpr uint32_t RT_HALF_BYTES = RT_KEYS * RT_D * 2;     // one fp16 key tile (hi or lo): 32 KB
constexpr uint32_t RT_STAGE_BYTES = 2 * RT_HALF_BYTES;     // hi + lo
constexpr uint32_t RT_KBLOCK = RT_KEYS * 128;              // one 64-column (128 B) block of a key tile
constexpr uint32_t RT_PPART = RT_KEYS * 128;               // coefficient tile [128 keys][128 B], one part (hi or lo)
constexpr float RT_XSCALE = 256.0f;                        // x, y pre-scale (matches split_half_kernel of gram_tc.cu)
constexpr float RT_PSCALE = 16384.0f;                      // p pre-scale
constexpr int RT_MAXC = 8;
constexpr float RT_LOG2E = 1.4426950408889634f;

// ---- shared-memory plans (offsets from a 1024-aligned base; every MMA tile is 1024-aligned)
struct FwdPlan {
    // Y tile: per 64-column block, rows [y_hi 0-31 ; y_lo 32-63] x 128 B -- one N = 64 B operand for the key tile's
    // hi half (x_hi.y_hi | x_hi.y_lo in separate accumulator columns) and an N = 32 one for its lo half
    static constexpr uint32_t Y_KB = 2 * RT_SEEDS * 128;                     // 8 KB
    static constexpr uint32_t Y_LO = RT_SEEDS * 128;                         // row 32 of a block
    static constexpr uint32_t x = 0;
    static constexpr uint32_t y = x + RT_STAGES * RT_STAGE_BYTES;            // Y tile, two 64-column blocks
    static constexpr uint32_t p = y + 2 * Y_KB;                              // P tile [128 keys][p_hi 64 B | p_lo 64 B]
    static constexpr uint32_t ys = p + RT_PPART;                             // fp32 [32][128] current seeds
    static constexpr uint32_t misc = ys + RT_SEEDS * RT_D * 4;
    static constexpr uint32_t total = misc + 4096;
    // aliases inside operand tiles that are dead between the tile loop and the next one
    static constexpr uint32_t part_o = p;                                    // fp32 [32][128] partial numerators
};
struct BwdPlan {
    // stacked operand tile: per 64-column block, rows [y_hi 0-31 ; gm_hi 32-63 ; y_lo 64-95 ; gm_lo 96-127] x 128 B
    static constexpr uint32_t YG_KB = 4 * RT_SEEDS * 128;                    // 16 KB
    static constexpr uint32_t YG_ROWS = RT_SEEDS * 128;                      // 32 rows
    static constexpr uint32_t x = 0;
    static constexpr uint32_t yg = x + RT_STAGES * RT_STAGE_BYTES;
    static constexpr uint32_t de = yg + 2 * YG_KB;                           // [128 keys][ds_hi | ds_lo] , [128 keys][e1_hi | e1_lo]
    static constexpr uint32_t gy = de + 2 * RT_PPART;                        // fp32 [32][128] dL/dy^{t+1}
    static constexpr uint32_t misc = gy + RT_SEEDS * RT_D * 4;
    static constexpr uint32_t total = misc + 4096;
    // aliases inside the coefficient tile (dead between the tile loop and the next one)
    static constexpr uint32_t part_o = de;                                   // fp32 [32][128]
    static constexpr uint32_t gms = de + RT_PPART;                           // fp32 [32][128]
};

struct RtMisc {
    uint64_t x_full[RT_STAGES], x_empty[RT_STAGES];
    uint64_t s_full[2], s_free[2];
    uint64_t c_full, c_free;             // coefficient tile (P / ds|e1) written / consumed
    uint64_t gx_full, gx_free;           // backward: gX accumulator of a tile complete / read out
    uint64_t o_full, y_full;
    uint32_t tmem_base;
    float part_z[RT_SEEDS];
    float zwarp[2][8][16];               // [unit][epilogue warp][seed]
    float red[RT_EPI];
    float gmm[RT_SEEDS], dinv[RT_SEEDS], gmax[RT_SEEDS], gbound[RT_SEEDS];
    float scal[4];                        // backward: SG, SD, SE
};
static_assert(sizeof(RtMisc) <= 4096, "misc block");

__host__ __device__ constexpr uint32_t idesc_f16_mm(int M, int N, bool a_mn, bool b_mn) {
    return (1u << 4) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
// 16 lanes x 64 consecutive 32-bit columns in the accumulator-fragment layout: thread t of the warp holds, for every group
// k = 0..7 of 8 columns, columns 8k + 2(t % 4) + {0, 1} of lane (t / 4) in r[4k], r[4k+1] and of lane (t / 4) + 8 in r[4k+2], r[4k+3]
__device__ __forceinline__ void tmem_ld_16x256b_x8(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void red_add_v4(float* p, float x, float y, float z, float w) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// (a, b) -> packed f16 pairs hi, lo with a = hi.x + lo.x (22 significant bits), likewise b
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 f = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - f.x, b - f.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

// Sum 16 per-lane values over the 32 lanes of a warp; lane l ends up with the total of value index (l >> 1) & 15.
__device__ __forceinline__ float warp_transpose_sum16(const float (&v)[16], int lane) {
    float w8[8], w4[4], w2[2];
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float send = b4 ? v[i] : v[i + 8], keep = b4 ? v[i + 8] : v[i];
        w8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float send = b3 ? w8[i] : w8[i + 4], keep = b3 ? w8[i + 4] : w8[i];
        w4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float send = b2 ? w4[i] : w4[i + 2], keep = b2 ? w4[i + 2] : w4[i];
        w2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    const float send = b1 ? w2[0] : w2[1], keep = b1 ? w2[1] : w2[0];
    float r = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    r += __shfl_xor_sync(0xffffffffu, r, 1);
    return r;
}

// write 8 consecutive fp32 (pre-scaled) of row `r`, 16-byte chunk `chunk` (0..7) of a [rows][128 B] SWIZZLE_128B
// block as hi / lo fp16
__device__ __forceinline__ void store_chunk_hilo(uint8_t* hi_blk, uint8_t* lo_blk, int r, int chunk, const float (&f)[8]) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) split2(f[2 * e], f[2 * e + 1], h[e], l[e]);
    const uint32_t off = (uint32_t)r * 128u + (uint32_t)((chunk ^ (r & 7)) << 4);
    *reinterpret_cast<uint4*>(hi_blk + off) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(lo_blk + off) = make_uint4(l[0], l[1], l[2], l[3]);
}

__device__ __forceinline__ void warp_sum4(float (&v)[4]) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
}
__device__ __forceinline__ void warp_max4(float (&v)[4]) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = fmaxf(v[q], __shfl_xor_sync(0xffffffffu, v[q], o));
}

// power of two 2^e with v * 2^e in [2^(top-1), 2^top)   (v > 0, finite)
__device__ __forceinline__ float pow2_scale_to(float v, int top) {
    int ex;
    frexpf(v, &ex);                   // v = m 2^ex, m in [0.5, 1)
    return ldexpf(1.0f, top - ex);
}

// Optional phase timeline (diagnostics): epilogue thread 0 of CTA (0, 0, 0) records (phase id, clock64).
__device__ long long g_rt_dbg[8192];
__device__ int g_rt_dbg_n;
#define RT_MARK(id)                                                                        \
    do { if ((a.dbg & 1) && blockIdx.x == 0 && blockIdx.z == 0 && threadIdx.x == 128) {          \
             const int n__ = g_rt_dbg_n; if (n__ < 4090) { g_rt_dbg[2 * n__] = (id); g_rt_dbg[2 * n__ + 1] = clock64(); g_rt_dbg_n = n__ + 1; } } } while (0)

struct RowsArgs {
    const float* X; const float* bw; const int32_t* idx; const int32_t* K;
    int N, B, T, Kcap;
    // forward
    float* traj; float* stat; float* C_out;
    // backward
    const float* traj_in; const float* stat_in; const float* gC; float* gX;
    int dbg;
};

// The key tiles of a shape form RT_UNITS = 8 fixed UNITS of G = ceil(nt / 8) consecutive tiles, whatever the cluster size.
// A CTA of a c-CTA cluster (c = 4 or 8) owns 8 / c consecutive units = tiles [j0, j0 + ntl); its first unit ends at local
// tile G.  Partial sums over keys are formed PER UNIT (one accumulator per unit), units 2r and 2r+1 are added first and the
// four pair sums then in order -- the same floating-point expression whether a pair is added inside a CTA (c = 4) or by the
// reducing CTA (c = 8), so a shape's result does not depend on the cluster size its launch was given.
constexpr int RT_UNITS = 8;
__device__ __forceinline__ void my_tiles(int N, int csize, int rank, int& j0, int& ntl, int& G) {
    const int nt = (N + RT_KEYS - 1) / RT_KEYS;
    G = max(1, (nt + RT_UNITS - 1) / RT_UNITS);
    const int tpc = G * (RT_UNITS / csize);
    j0 = min(nt, rank * tpc);
    ntl = max(0, min(nt, rank * tpc + tpc) - j0);
}
// pair sums, then the four pairs in order (see my_tiles)
__device__ __forceinline__ float unit_sum(const float (&q)[RT_MAXC], int csize) {
    if (csize == RT_UNITS) return (((q[0] + q[1]) + (q[2] + q[3])) + (q[4] + q[5])) + (q[6] + q[7]);
    return ((q[0] + q[1]) + q[2]) + q[3];
}

__device__ __forceinline__ void produce_tile(uint8_t* smem, RtMisc* m, const CUtensorMap* tmap, uint32_t st, int row0, int b, int B) {
    mbar_arrive_expect_tx(&m->x_full[st], RT_STAGE_BYTES);
    const uint32_t dst = smem_u32(smem + (size_t)st * RT_STAGE_BYTES);
    tma_load_3d(dst, tmap, &m->x_full[st], 0, row0, b);
    tma_load_3d(dst + RT_KBLOCK, tmap, &m->x_full[st], 64, row0, b);
    tma_load_3d(dst + RT_HALF_BYTES, tmap, &m->x_full[st], 0, row0, B + b);
    tma_load_3d(dst + RT_HALF_BYTES + RT_KBLOCK, tmap, &m->x_full[st], 64, row0, B + b);
}

// =========================================================================================== forward
__global__ void __launch_bounds__(RT_THREADS, 1) rows_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmap, const RowsArgs a) {
    cg::cluster_group cluster = cg::this_cluster();
    const int csize = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    const int b = blockIdx.z, k0 = blockIdx.y * RT_SEEDS;
    const int N = a.N, T = a.T, Kcap = a.Kcap;
    const int Kb = min(a.K[b], Kcap);
    const int nrows = max(0, min(RT_SEEDS, Kb - k0));
    const int krows = min(RT_SEEDS, Kcap - k0);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    float* traj_b = a.traj + (size_t)b * (T + 1) * Kcap * RT_D;
    float* stat_b = a.stat + (size_t)b * T * Kcap * 2;

    // padded rows of this seed group are defined as zero
    for (int t = rank; t <= T; t += csize) {
        for (int e = tid; e < (krows - nrows) * RT_D; e += RT_THREADS) {
            const int r = nrows + e / RT_D, c = e % RT_D;
            traj_b[((size_t)t * Kcap + k0 + r) * RT_D + c] = 0.f;
            if (t == T) a.C_out[((size_t)b * Kcap + k0 + r) * RT_D + c] = 0.f;
        }
        if (t < T)
            for (int e = tid; e < (krows - nrows) * 2; e += RT_THREADS) stat_b[((size_t)t * Kcap + k0 + nrows) * 2 + e] = 0.f;
    }
    if (nrows == 0) return;   // uniform over the cluster, before any barrier / allocation

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // offset on the __shared__ symbol: accesses stay LDS / STS
    RtMisc* m = reinterpret_cast<RtMisc*>(smem + FwdPlan::misc);
    float* ys = reinterpret_cast<float*>(smem + FwdPlan::ys);
    float* part_o = reinterpret_cast<float*>(smem + FwdPlan::part_o);

    int j0, ntl, G;
    my_tiles(N, csize, rank, j0, ntl, G);
    const bool resident = ntl <= RT_STAGES;       // the CTA's key slice stays in shared memory for all T iterations

    if (tid == 0) {
        for (int s = 0; s < RT_STAGES; ++s) { mbar_init(&m->x_full[s], 1); mbar_init(&m->x_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&m->s_full[s], 1); mbar_init(&m->s_free[s], RT_EPI); }
        mbar_init(&m->c_full, RT_EPI); mbar_init(&m->c_free, 1);
        mbar_init(&m->o_full, 1); mbar_init(&m->y_full, RT_EPI);
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) prefetch_tensormap(&tmap);
    if (warp == 2) { tmem_alloc(&m->tmem_base, 256); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = m->tmem_base;
    constexpr uint32_t COL_S = 0 /* [buf][x_hi.y_hi + x_lo.y_hi | x_hi.y_lo] */, COL_O = 128 /* [unit][same split] */;

    const float* Xb = a.X + (size_t)b * N * RT_D;
    const int32_t* idx_b = a.idx + (size_t)b * Kcap + k0;
    const float bwv = a.bw[b];
    const float b2 = bwv * bwv;
    const float a_mul = (1.0f / b2) * (1.0f / (RT_XSCALE * RT_XSCALE));

    // epilogue-thread coordinates (valid for warps >= 4)
    const int et = tid - 128, ew = warp - 4, half = (ew >> 2) & 1;
    const bool seeds_live = 16 * half < nrows;                     // any real seed among this thread's 16 columns
    const int row = 32 * (ew & 3) + lane;                          // TMEM lane: key within the tile / column d of O^T
    const uint32_t lane_base = (uint32_t)(32 * (ew & 3)) << 16;
    // reduce-phase ownership: this CTA finishes seeds [rank * rpc, rank * rpc + rpc)
    const int rpc = (RT_SEEDS + csize - 1) / csize;

    auto write_y_tile = [&]() {      // ys (fp32) -> Y tile hi | lo, K-major SWIZZLE_128B, rows = seeds
        for (int q = et; q < RT_SEEDS * 16; q += RT_EPI) {
            const int r = q >> 4, c8 = q & 15;                 // 8 columns [8 c8, 8 c8 + 8)
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = ys[r * RT_D + 8 * c8 + e] * RT_XSCALE;
            uint8_t* hi_blk = smem + FwdPlan::y + (c8 >> 3) * FwdPlan::Y_KB;
            store_chunk_hilo(hi_blk, hi_blk + FwdPlan::Y_LO, r, c8 & 7, f);
        }
        fence_proxy_async();
        mbar_arrive(&m->y_full);
    };

    if (warp >= 4) {
        for (int q = et; q < RT_SEEDS * (RT_D / 4); q += RT_EPI) {
            const int r = q / (RT_D / 4), c = q % (RT_D / 4);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < nrows) v = reinterpret_cast<const float4*>(Xb + (size_t)idx_b[r] * RT_D)[c];
            reinterpret_cast<float4*>(ys + r * RT_D)[c] = v;
            if (rank == 0 && r < nrows) reinterpret_cast<float4*>(traj_b + ((size_t)k0 + r) * RT_D)[c] = v;
            if (T == 0 && rank == 0 && r < nrows) reinterpret_cast<float4*>(a.C_out + ((size_t)b * Kcap + k0 + r) * RT_D)[c] = v;
        }
        epi_bar();
        write_y_tile();
    }

    uint32_t it = 0;                       // tiles processed so far (every role counts alike)
    for (int t = 0; t < T; ++t) {
        if (warp == 0) {
            // ================================ TMA producer ================================
            if (lane == 0 && (!resident || t == 0)) {
                uint32_t pit = it;
                for (int jl = 0; jl < ntl; ++jl, ++pit) {
                    const uint32_t st = resident ? (uint32_t)jl : pit % RT_STAGES, ph = (pit / RT_STAGES) & 1;
                    if (!resident) mbar_wait(&m->x_empty[st], ph ^ 1);
                    produce_tile(smem, m, &tmap, st, (j0 + jl) * RT_KEYS, b, a.B);
                }
            }
            __syncwarp();
        } else if (warp == 1) {
            // ================================= MMA issuer =================================
            if (lane == 0) {
                constexpr uint32_t id1w = idesc_f16_mm(RT_KEYS, 2 * RT_SEEDS, false, false);   // x_hi . [y_hi ; y_lo]^T
                constexpr uint32_t id1n = idesc_f16_mm(RT_KEYS, RT_SEEDS, false, false);       // x_lo . y_hi^T
                constexpr uint32_t id2w = idesc_f16_mm(RT_D, 2 * RT_SEEDS, true, true);        // x_hi^T . [p_hi | p_lo]
                constexpr uint32_t id2n = idesc_f16_mm(RT_D, RT_SEEDS, true, true);            // x_lo^T . p_hi
                const uint32_t ybase = smem_u32(smem + FwdPlan::y), pbase = smem_u32(smem + FwdPlan::p);
                auto gemm2 = [&](uint32_t i, int jl) {
                    const uint32_t st = resident ? (uint32_t)jl : i % RT_STAGES;
                    const bool first = jl == 0 || jl == G;             // first tile of a unit: fresh accumulator
                    const uint32_t col_o = COL_O + (jl >= G ? 64u : 0u);
                    mbar_wait(&m->c_full, i & 1);
                    tc_fence_after();
                    const uint32_t xb = smem_u32(smem + (size_t)st * RT_STAGE_BYTES);
#pragma unroll
                    for (int kk = 0; kk < RT_KEYS / 16; ++kk) {        // 16 keys per MMA
                        const uint64_t a_hi = smem_desc_sw128(xb + kk * 2048, RT_KBLOCK, 1024);
                        const uint64_t a_lo = smem_desc_sw128(xb + RT_HALF_BYTES + kk * 2048, RT_KBLOCK, 1024);
                        const uint64_t bp = smem_desc_sw128(pbase + kk * 2048, RT_KBLOCK, 1024);
                        mma_f16_ss(tmem + col_o, a_hi, bp, id2w, !(first && kk == 0));
                        mma_f16_ss(tmem + col_o, a_lo, bp, id2n, true);
                    }
                    if (!resident) mma_commit(&m->x_empty[st]);
                    mma_commit(&m->c_free);
                };
                mbar_wait(&m->y_full, t & 1);
                tc_fence_after();
                uint32_t mit = it;
                for (int jl = 0; jl < ntl; ++jl, ++mit) {
                    const uint32_t st = resident ? (uint32_t)jl : mit % RT_STAGES, buf = mit & 1;
                    if (!resident) mbar_wait(&m->x_full[st], (mit / RT_STAGES) & 1);
                    else if (t == 0) mbar_wait(&m->x_full[st], 0);
                    mbar_wait(&m->s_free[buf], ((mit >> 1) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t xb = smem_u32(smem + (size_t)st * RT_STAGE_BYTES);
#pragma unroll
                    for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {               // 16 d-elements (32 B) per MMA
                            const uint32_t xo = kb * RT_KBLOCK + ks * 32, yo = kb * FwdPlan::Y_KB + ks * 32;
                            const uint64_t a_hi = smem_desc_sw128(xb + xo, 16, 1024);
                            const uint64_t a_lo = smem_desc_sw128(xb + RT_HALF_BYTES + xo, 16, 1024);
                            const uint64_t by = smem_desc_sw128(ybase + yo, 16, 1024);
                            mma_f16_ss(tmem + COL_S + buf * 64, a_hi, by, id1w, (kb | ks) != 0);
                            mma_f16_ss(tmem + COL_S + buf * 64, a_lo, by, id1n, true);
                        }
                    mma_commit(&m->s_full[buf]);
                    if (jl > 0) gemm2(mit - 1, jl - 1);
                }
                if (ntl > 0) { gemm2(mit - 1, ntl - 1); mma_commit(&m->o_full); }
            }
            __syncwarp();
        } else if (warp >= 4) {
            // ============================= coefficients (thread = key) ==============================
            float zlane[2] = {0.f, 0.f};                            // per unit
            uint32_t eit = it;
            for (int jl = 0; jl < ntl; ++jl, ++eit) {
                const uint32_t buf = eit & 1;
                mbar_wait(&m->s_full[buf], (eit >> 1) & 1);
                RT_MARK(4);
                tc_fence_after();
                uint32_t v[16], w[16];
                tmem_ld16(tmem + lane_base + COL_S + buf * 64 + 16 * half, v);
                tmem_ld16(tmem + lane_base + COL_S + buf * 64 + 32 + 16 * half, w);
                tmem_wait_ld();
                tc_fence_before();
                mbar_arrive(&m->s_free[buf]);
                const bool kvalid = (j0 + jl) * RT_KEYS + row < N;
                float p[16];
                if (kvalid && seeds_live) {
#pragma unroll
                    for (int s = 0; s < 16; ++s) {
                        // a = -(2 - 2 dot) / b^2 / 2 = (dot - 1) / b^2 with dot = S 2^-16   (src/mean_shift.py:65,68)
                        const float av = ((__uint_as_float(v[s]) + __uint_as_float(w[s])) - RT_XSCALE * RT_XSCALE) * a_mul;
                        p[s] = ex2_approx(fminf(fmaxf(av, PRIFIT_LO), PRIFIT_HI) * RT_LOG2E);
                    }
                } else {
#pragma unroll
                    for (int s = 0; s < 16; ++s) p[s] = 0.f;
                }
                const float zt = warp_transpose_sum16(p, lane);
                if (jl < G) zlane[0] += zt; else zlane[1] += zt;
                uint32_t h[8], l[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) split2(p[2 * e] * RT_PSCALE, p[2 * e + 1] * RT_PSCALE, h[e], l[e]);
                RT_MARK(5);
                mbar_wait(&m->c_free, (eit & 1) ^ 1);                           // GEMM2 of the previous tile is done with the tile
                RT_MARK(6);
                uint8_t* prow = smem + FwdPlan::p + (uint32_t)row * 128u;
#pragma unroll
                for (int c = 0; c < 2; ++c) {          // p_hi: chunks 0-3 of the row, p_lo: chunks 4-7
                    const uint32_t oh = (uint32_t)(((2 * half + c) ^ (row & 7)) << 4);
                    const uint32_t ol = (uint32_t)(((4 + 2 * half + c) ^ (row & 7)) << 4);
                    *reinterpret_cast<uint4*>(prow + oh) = make_uint4(h[4 * c], h[4 * c + 1], h[4 * c + 2], h[4 * c + 3]);
                    *reinterpret_cast<uint4*>(prow + ol) = make_uint4(l[4 * c], l[4 * c + 1], l[4 * c + 2], l[4 * c + 3]);
                }
                fence_proxy_async();
                mbar_arrive(&m->c_full);
                RT_MARK(7);
            }
            // partial numerators O^T (lane = column d) and denominators of this CTA's key slice: one per unit, added here
            float po[16];
#pragma unroll
            for (int s = 0; s < 16; ++s) po[s] = 0.f;
            if (ntl > 0) {
                uint32_t v[16], w[16];
                mbar_wait(&m->o_full, t & 1);
                tc_fence_after();
                tmem_ld16(tmem + lane_base + COL_O + 16 * half, v);
                tmem_ld16(tmem + lane_base + COL_O + 32 + 16 * half, w);
                tmem_wait_ld();
#pragma unroll
                for (int s = 0; s < 16; ++s) po[s] = (__uint_as_float(v[s]) + __uint_as_float(w[s])) * (1.0f / (RT_XSCALE * RT_PSCALE));
                if (ntl > G) {                                         // second unit of this CTA
                    tmem_ld16(tmem + lane_base + COL_O + 64 + 16 * half, v);
                    tmem_ld16(tmem + lane_base + COL_O + 64 + 32 + 16 * half, w);
                    tmem_wait_ld();
#pragma unroll
                    for (int s = 0; s < 16; ++s) po[s] += (__uint_as_float(v[s]) + __uint_as_float(w[s])) * (1.0f / (RT_XSCALE * RT_PSCALE));
                }
                tc_fence_before();
            }
#pragma unroll
            for (int s = 0; s < 16; ++s) part_o[(16 * half + s) * RT_D + row] = po[s];
            if ((lane & 1) == 0) { m->zwarp[0][ew][lane >> 1] = zlane[0]; m->zwarp[1][ew][lane >> 1] = zlane[1]; }
            epi_bar();
            if (et < RT_SEEDS) {
                const int hz = et >> 4, sz = et & 15;
                float zu[2];
#pragma unroll
                for (int u = 0; u < 2; ++u)
                    zu[u] = (m->zwarp[u][4 * hz][sz] + m->zwarp[u][4 * hz + 1][sz]) + (m->zwarp[u][4 * hz + 2][sz] + m->zwarp[u][4 * hz + 3][sz]);
                m->part_z[et] = zu[0] + zu[1];
            }
            RT_MARK(11);
        }
        it += ntl;
        cluster.sync();
        RT_MARK(12);
        // ---- reduce over the cluster in fixed rank order; finish the owned seeds (src/mean_shift.py:75-82).
        //      One warp per owned seed row, one float4 per lane: every remote load of the step is independent.
        if (warp >= 4) {
            for (int f = et; f < rpc * 32; f += RT_EPI) {
                const int orow = rank * rpc + (f >> 5), c4 = f & 31;       // warp-uniform row
                if (orow >= RT_SEEDS) break;
                float zq[RT_MAXC];
                float4 oq[RT_MAXC];
#pragma unroll
                for (int q = 0; q < RT_MAXC; ++q) {
                    zq[q] = q < csize ? cluster.map_shared_rank(m->part_z, q)[orow] : 0.f;
                    oq[q] = q < csize ? reinterpret_cast<const float4*>(cluster.map_shared_rank(part_o, q) + orow * RT_D)[c4]
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                float ox[RT_MAXC], oy[RT_MAXC], oz[RT_MAXC], ow[RT_MAXC];
#pragma unroll
                for (int q = 0; q < RT_MAXC; ++q) { ox[q] = oq[q].x; oy[q] = oq[q].y; oz[q] = oq[q].z; ow[q] = oq[q].w; }
                const float zsum = unit_sum(zq, csize);
                const float4 o = make_float4(unit_sum(ox, csize), unit_sum(oy, csize), unit_sum(oz, csize), unit_sum(ow, csize));
                const float dinv = 1.0f / zsum;
                const float4 y = reinterpret_cast<const float4*>(ys + orow * RT_D)[c4];
                float4 u;                                                  // new = y + ((K X) D - y)
                u.x = y.x + (o.x * dinv - y.x); u.y = y.y + (o.y * dinv - y.y);
                u.z = y.z + (o.z * dinv - y.z); u.w = y.w + (o.w * dinv - y.w);
                const float nrm = sqrtf(warp_sum(fmaf(u.x, u.x, fmaf(u.y, u.y, fmaf(u.z, u.z, u.w * u.w)))));
                const float4 yn = make_float4(u.x / nrm, u.y / nrm, u.z / nrm, u.w / nrm);   // new /= ||new||
                for (int q = 0; q < csize; ++q) reinterpret_cast<float4*>(cluster.map_shared_rank(ys, q) + orow * RT_D)[c4] = yn;
                if (orow < nrows) {
                    reinterpret_cast<float4*>(traj_b + ((size_t)(t + 1) * Kcap + k0 + orow) * RT_D)[c4] = yn;
                    if (t == T - 1) reinterpret_cast<float4*>(a.C_out + ((size_t)b * Kcap + k0 + orow) * RT_D)[c4] = yn;
                    if (c4 == 0) {
                        stat_b[((size_t)t * Kcap + k0 + orow) * 2 + 0] = zsum;
                        stat_b[((size_t)t * Kcap + k0 + orow) * 2 + 1] = nrm;
                    }
                }
            }
        }
        RT_MARK(13);
        cluster.sync();
        RT_MARK(14);
        if (warp >= 4 && t + 1 < T) write_y_tile();
        RT_MARK(3);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

// ========================================================================================== backward
// One reverse step t for seed r with g = dL/dy^{t+1} -- the formulas of meanshift_rows.cu:
//   g_m = (g - (g.y^{t+1}) y^{t+1}) / ||u||,  dkappa_j = (g_m.x_j - g_m.m) D,  ds_j = kappa_j dkappa_j [lo<=a_j<=hi] / b^2,
//   dL/dy^t = sum_j ds_j x_j,   dL/dx_j += ds_j y^t + kappa_j D g_m
__global__ void __launch_bounds__(RT_THREADS, 1) rows_tc_bwd_kernel(const __grid_constant__ CUtensorMap tmap, const RowsArgs a) {
    cg::cluster_group cluster = cg::this_cluster();
    const int csize = (int)cluster.num_blocks(), rank = (int)cluster.block_rank();
    const int b = blockIdx.z;
    const int N = a.N, T = a.T, Kcap = a.Kcap;
    const int Kb = min(a.K[b], Kcap);
    if (Kb <= 0 || T <= 0) {
        // T == 0: center = X[idx]; the gradient lands on the seeds' own rows
        if (T <= 0 && Kb > 0 && rank == 0)
            for (int e = threadIdx.x; e < Kb * RT_D; e += RT_THREADS) {
                const int r = e / RT_D, c = e % RT_D;
                atomicAdd(a.gX + ((size_t)b * N + a.idx[(size_t)b * Kcap + r]) * RT_D + c, a.gC[((size_t)b * Kcap + r) * RT_D + c]);
            }
        return;       // uniform over the cluster
    }
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // offset on the __shared__ symbol: accesses stay LDS / STS
    RtMisc* m = reinterpret_cast<RtMisc*>(smem + BwdPlan::misc);
    float* gy = reinterpret_cast<float*>(smem + BwdPlan::gy);
    float* part_o = reinterpret_cast<float*>(smem + BwdPlan::part_o);
    float* gms = reinterpret_cast<float*>(smem + BwdPlan::gms);

    int j0, ntl, G;
    my_tiles(N, csize, rank, j0, ntl, G);
    const bool resident = ntl <= RT_STAGES;

    if (tid == 0) {
        for (int s = 0; s < RT_STAGES; ++s) { mbar_init(&m->x_full[s], 1); mbar_init(&m->x_empty[s], 1); }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&m->s_full[s], 1); mbar_init(&m->s_free[s], RT_EPI);
        }
        mbar_init(&m->gx_full, 1); mbar_init(&m->gx_free, RT_EPI);
        mbar_init(&m->c_full, RT_EPI); mbar_init(&m->c_free, 1);
        mbar_init(&m->o_full, 1); mbar_init(&m->y_full, RT_EPI);
        fence_barrier_init();
    }
    if (warp == 0 && lane == 0) prefetch_tensormap(&tmap);
    if (warp == 2) { tmem_alloc(&m->tmem_base, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = m->tmem_base;
    // S[buf]: [x.y_hi | x.gm_hi | x_hi.y_lo | x_hi.gm_lo] (32 columns each); GY[unit]: [x.ds_hi | x_hi.ds_lo]; GX: one 128-column accumulator
    constexpr uint32_t COL_S = 0, COL_GY = 256, COL_GX = 384;

    float* gXb = a.gX + (size_t)b * N * RT_D;
    const float* traj_b = a.traj_in + (size_t)b * (T + 1) * Kcap * RT_D;
    const float* stat_b = a.stat_in + (size_t)b * T * Kcap * 2;
    const float bwv = a.bw[b];
    const float b2 = bwv * bwv;
    const float a_mul = (1.0f / b2) * (1.0f / (RT_XSCALE * RT_XSCALE));

    const int et = tid - 128, ew = warp - 4, half = (ew >> 2) & 1;
    const int row = 32 * (ew & 3) + lane;
    const uint32_t lane_base = (uint32_t)(32 * (ew & 3)) << 16;
    const int rpc = (RT_SEEDS + csize - 1) / csize, tpr = RT_EPI / rpc;
    const int rr = et / tpr, tc = et - rr * tpr;
    const int myrow = rank * rpc + rr;
    const bool owner = warp >= 4 && rr < rpc && myrow < RT_SEEDS;

    uint32_t it = 0, itn = 0;              // tiles / iterations processed so far
    bool x_loaded = false;
    for (int k0 = 0; k0 < Kb; k0 += RT_SEEDS) {
        const int nrows = min(RT_SEEDS, Kb - k0);
        if (warp >= 4) {
            const float* gC_b = a.gC + ((size_t)b * Kcap + k0) * RT_D;
            for (int q = et; q < RT_SEEDS * (RT_D / 4); q += RT_EPI) {
                const int r = q / (RT_D / 4), c = q % (RT_D / 4);
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (r < nrows) v = reinterpret_cast<const float4*>(gC_b + (size_t)r * RT_D)[c];
                reinterpret_cast<float4*>(gy + r * RT_D)[c] = v;
            }
            epi_bar();
        }
        for (int t = T - 1; t >= 0; --t, ++itn) {
            float sg = 1.f, sd = 1.f, se = 1.f;
            if (warp == 0) {
                if (lane == 0 && (!resident || !x_loaded)) {
                    uint32_t pit = it;
                    for (int jl = 0; jl < ntl; ++jl, ++pit) {
                        const uint32_t st = resident ? (uint32_t)jl : pit % RT_STAGES, ph = (pit / RT_STAGES) & 1;
                        if (!resident) mbar_wait(&m->x_empty[st], ph ^ 1);
                        produce_tile(smem, m, &tmap, st, (j0 + jl) * RT_KEYS, b, a.B);
                    }
                }
                __syncwarp();
            } else if (warp == 1) {
                if (lane == 0) {
                    constexpr uint32_t id1w = idesc_f16_mm(RT_KEYS, 4 * RT_SEEDS, false, false);  // x_hi . [y_hi; gm_hi; y_lo; gm_lo]^T
                    constexpr uint32_t id1n = idesc_f16_mm(RT_KEYS, 2 * RT_SEEDS, false, false);  // x_lo . [y_hi; gm_hi]^T
                    constexpr uint32_t id2w = idesc_f16_mm(RT_D, 2 * RT_SEEDS, true, true);       // x_hi^T . [ds_hi | ds_lo]
                    constexpr uint32_t id2n = idesc_f16_mm(RT_D, RT_SEEDS, true, true);           // x_lo^T . ds_hi
                    constexpr uint32_t id3 = idesc_f16_mm(RT_KEYS, RT_D, false, true);            // gX = [ds | e1] . [Y ; Gm]
                    const uint32_t ygbase = smem_u32(smem + BwdPlan::yg), cbase = smem_u32(smem + BwdPlan::de);
                    auto gemm_g = [&](uint32_t i, int jl) {
                        const uint32_t st = resident ? (uint32_t)jl : i % RT_STAGES;
                        const bool first = jl == 0 || jl == G;         // first tile of a unit: fresh dL/dy accumulator
                        const uint32_t col_gy = COL_GY + (jl >= G ? 64u : 0u);
                        mbar_wait(&m->c_full, i & 1);
                        tc_fence_after();
                        const uint32_t xb = smem_u32(smem + (size_t)st * RT_STAGE_BYTES);
#pragma unroll
                        for (int kk = 0; kk < RT_KEYS / 16; ++kk) {
                            const uint64_t a_hi = smem_desc_sw128(xb + kk * 2048, RT_KBLOCK, 1024);
                            const uint64_t a_lo = smem_desc_sw128(xb + RT_HALF_BYTES + kk * 2048, RT_KBLOCK, 1024);
                            const uint64_t bd = smem_desc_sw128(cbase + kk * 2048, RT_KBLOCK, 1024);
                            mma_f16_ss(tmem + col_gy, a_hi, bd, id2w, !(first && kk == 0));
                            mma_f16_ss(tmem + col_gy, a_lo, bd, id2n, true);
                        }
                        mbar_wait(&m->gx_free, (i & 1) ^ 1);           // the previous tile's gX has been read out
                        tc_fence_after();
                        // K pairs (coefficient columns, operand rows): (ds_hi, y_hi) (ds_hi, y_lo) (ds_lo, y_hi)
                        //                                              (e1_hi, gm_hi) (e1_hi, gm_lo) (e1_lo, gm_hi)
                        constexpr uint32_t a_off[6] = {0, 0, 64, RT_PPART, RT_PPART, RT_PPART + 64};
                        constexpr uint32_t b_row[6] = {0, 64, 0, 32, 96, 32};
#pragma unroll
                        for (int pr = 0; pr < 6; ++pr)
#pragma unroll
                            for (int ks = 0; ks < 2; ++ks) {           // 16 seeds per MMA
                                const uint64_t ad = smem_desc_sw128(cbase + a_off[pr] + ks * 32, 16, 1024);
                                const uint64_t bd = smem_desc_sw128(ygbase + (b_row[pr] + 16 * ks) * 128, BwdPlan::YG_KB, 1024);
                                mma_f16_ss(tmem + COL_GX, ad, bd, id3, (pr | ks) != 0);
                            }
                        if (!resident) mma_commit(&m->x_empty[st]);
                        mma_commit(&m->c_free);
                        mma_commit(&m->gx_full);
                    };
                    mbar_wait(&m->y_full, itn & 1);
                    tc_fence_after();
                    uint32_t mit = it;
                    for (int jl = 0; jl < ntl; ++jl, ++mit) {
                        const uint32_t st = resident ? (uint32_t)jl : mit % RT_STAGES, buf = mit & 1;
                        if (!resident) mbar_wait(&m->x_full[st], (mit / RT_STAGES) & 1);
                        else if (!x_loaded) mbar_wait(&m->x_full[st], 0);
                        mbar_wait(&m->s_free[buf], ((mit >> 1) & 1) ^ 1);
                        tc_fence_after();
                        const uint32_t xb = smem_u32(smem + (size_t)st * RT_STAGE_BYTES);
#pragma unroll
                        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) {           // 16 d-elements (32 B) per MMA
                                const uint32_t xo = kb * RT_KBLOCK + ks * 32;
                                const uint64_t a_hi = smem_desc_sw128(xb + xo, 16, 1024);
                                const uint64_t a_lo = smem_desc_sw128(xb + RT_HALF_BYTES + xo, 16, 1024);
                                const uint64_t bd = smem_desc_sw128(ygbase + kb * BwdPlan::YG_KB + ks * 32, 16, 1024);
                                mma_f16_ss(tmem + COL_S + buf * 128, a_hi, bd, id1w, (kb | ks) != 0);
                                mma_f16_ss(tmem + COL_S + buf * 128, a_lo, bd, id1n, true);
                            }
                        mma_commit(&m->s_full[buf]);
                        if (jl > 0) gemm_g(mit - 1, jl - 1);
                    }
                    if (ntl > 0) { gemm_g(mit - 1, ntl - 1); mma_commit(&m->o_full); }
                }
                __syncwarp();
            } else if (warp >= 4) {
                RT_MARK(1);
                // ---- per-seed preparation: one warp per seed, 4 seeds per warp; every global load of the step is
                //      issued before the first dependent instruction
                float4 yn[4], g4[4], gm[4];
                float zt[4], nr[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int r = ew * 4 + q;
                    yn[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                    zt[q] = 1.f; nr[q] = 1.f;
                    if (r < nrows) {
                        yn[q] = __ldg(reinterpret_cast<const float4*>(traj_b + ((size_t)(t + 1) * Kcap + k0 + r) * RT_D) + lane);
                        const float2 st2 = __ldg(reinterpret_cast<const float2*>(stat_b + ((size_t)t * Kcap + k0 + r) * 2));
                        zt[q] = st2.x; nr[q] = st2.y;
                    }
                }
                float4 ytile[2][2];                                    // this thread's two chunks of the y^t rows
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int q = et + RT_EPI * i, r = (q >> 4) & 31, c8 = q & 15;
                    ytile[i][0] = ytile[i][1] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (r < nrows) {
                        const float4* src = reinterpret_cast<const float4*>(traj_b + ((size_t)t * Kcap + k0 + r) * RT_D + 8 * c8);
                        ytile[i][0] = __ldg(src); ytile[i][1] = __ldg(src + 1);
                    }
                }
                float red4[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    g4[q] = reinterpret_cast<const float4*>(gy + (ew * 4 + q) * RT_D)[lane];
                    red4[q] = g4[q].x * yn[q].x + g4[q].y * yn[q].y + g4[q].z * yn[q].z + g4[q].w * yn[q].w;
                }
                warp_sum4(red4);
                float gmm4[4], g24[4], gmx4[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float gdot = red4[q], nrm = nr[q];
                    gm[q].x = (g4[q].x - gdot * yn[q].x) / nrm; gm[q].y = (g4[q].y - gdot * yn[q].y) / nrm;
                    gm[q].z = (g4[q].z - gdot * yn[q].z) / nrm; gm[q].w = (g4[q].w - gdot * yn[q].w) / nrm;
                    gmm4[q] = gm[q].x * (yn[q].x * nrm) + gm[q].y * (yn[q].y * nrm) + gm[q].z * (yn[q].z * nrm) + gm[q].w * (yn[q].w * nrm);
                    g24[q] = gm[q].x * gm[q].x + gm[q].y * gm[q].y + gm[q].z * gm[q].z + gm[q].w * gm[q].w;
                    gmx4[q] = fmaxf(fmaxf(fabsf(gm[q].x), fabsf(gm[q].y)), fmaxf(fabsf(gm[q].z), fabsf(gm[q].w)));
                }
                warp_sum4(gmm4);
                warp_sum4(g24);
                warp_max4(gmx4);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int r = ew * 4 + q;
                    const bool rl = r < nrows;
                    reinterpret_cast<float4*>(gms + r * RT_D)[lane] = rl ? gm[q] : make_float4(0.f, 0.f, 0.f, 0.f);
                    if (lane == 0) {
                        const float dinv = rl ? 1.0f / zt[q] : 0.f;
                        m->gmm[r] = rl ? gmm4[q] : 0.f; m->dinv[r] = dinv; m->gmax[r] = rl ? gmx4[q] : 0.f;
                        m->gbound[r] = rl ? 2.0f * sqrtf(g24[q]) * dinv / b2 : 0.f;
                    }
                }
                epi_bar();
                RT_MARK(2);
                // ---- scales of this step (identical in every thread and every CTA of the cluster)
                {
                    float gmx = 0.f, bnd = 0.f, dmx = 0.f;
                    for (int r = 0; r < RT_SEEDS; ++r) { gmx = fmaxf(gmx, m->gmax[r]); bnd = fmaxf(bnd, m->gbound[r]); dmx = fmaxf(dmx, m->dinv[r]); }
                    const bool ok = gmx > 0.f && bnd > 0.f && isfinite(gmx) && isfinite(bnd);
                    sg = ok ? pow2_scale_to(gmx, 8) : 1.0f;
                    sd = ok ? pow2_scale_to(bnd, 14) : 1.0f;
                    se = sd * RT_XSCALE / sg;
                    if (ok && dmx * se > 32768.0f) {                   // keep e1 inside fp16: give up bits of ds instead
                        const float shrink = pow2_scale_to(dmx * se, 15);
                        sd *= shrink;
                        se *= shrink;
                    }
                }
                // ---- stacked operand tile [y^t rows | g_m rows], hi | lo
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int q = et + RT_EPI * i;
                    const int which = q >> 9, r = (q >> 4) & 31, c8 = q & 15;
                    float f[8];
                    if (i < 2) {
                        const float4 u = ytile[i & 1][0], w = ytile[i & 1][1];
                        f[0] = u.x * RT_XSCALE; f[1] = u.y * RT_XSCALE; f[2] = u.z * RT_XSCALE; f[3] = u.w * RT_XSCALE;
                        f[4] = w.x * RT_XSCALE; f[5] = w.y * RT_XSCALE; f[6] = w.z * RT_XSCALE; f[7] = w.w * RT_XSCALE;
                    } else {
#pragma unroll
                        for (int e = 0; e < 8; ++e) f[e] = gms[r * RT_D + 8 * c8 + e] * sg;
                    }
                    uint8_t* hi_blk = smem + BwdPlan::yg + (c8 >> 3) * BwdPlan::YG_KB + which * BwdPlan::YG_ROWS;
                    store_chunk_hilo(hi_blk, hi_blk + 2 * BwdPlan::YG_ROWS, r, c8 & 7, f);
                }
                float gmm_r[16], dinv_r[16];
#pragma unroll
                for (int s = 0; s < 16; ++s) { gmm_r[s] = m->gmm[16 * half + s]; dinv_r[s] = m->dinv[16 * half + s]; }
                epi_bar();                                             // gms (aliases the coefficient tile) is consumed
                fence_proxy_async();
                mbar_arrive(&m->y_full);
                RT_MARK(3);

                const float s2_mul = 1.0f / (RT_XSCALE * sg), ds_mul = sd / b2;
                const float gx_mul = 1.0f / (sd * RT_XSCALE);
                auto flush = [&](uint32_t i, int jl) {                 // gX accumulator of tile i -> global (rows owned by this CTA)
                    mbar_wait(&m->gx_full, i & 1);
                    tc_fence_after();
                    // This warp's 32 lanes x 64 columns in the fragment layout (two 16-lane loads): a quad of threads holds 8
                    // consecutive columns of one row.  One exchange inside the quad turns two such groups into four consecutive
                    // columns per thread, so that every red.global.add.v4 of the warp covers 8 rows x 64 contiguous bytes --
                    // full 32-byte sectors (one row per lane touched 32 half-used sectors per instruction).
                    uint32_t g0[32], g1[32];
                    tmem_ld_16x256b_x8(tmem + lane_base + COL_GX + 64 * half, g0);
                    tmem_ld_16x256b_x8(tmem + lane_base + (16u << 16) + COL_GX + 64 * half, g1);
                    tmem_wait_ld();
                    tc_fence_before();
                    mbar_arrive(&m->gx_free);
                    if (a.dbg & 2) return;
                    const int qc = lane & 3, odd = qc & 1;
                    const int key0 = (j0 + jl) * RT_KEYS + 32 * (ew & 3) + (lane >> 2);
                    const int cbase = 64 * half + 4 * (qc >> 1) + 8 * odd;         // even threads: group k, odd threads: group k + 1
                    auto emit = [&](const uint32_t (&g)[32], int rbase) {
#pragma unroll
                        for (int kp = 0; kp < 4; ++kp)                             // groups 2 kp, 2 kp + 1
#pragma unroll
                            for (int rh = 0; rh < 2; ++rh) {                       // lane t/4, lane t/4 + 8
                                const uint32_t a0 = g[8 * kp + 2 * rh], a1 = g[8 * kp + 2 * rh + 1];          // group 2 kp
                                const uint32_t b0 = g[8 * kp + 4 + 2 * rh], b1 = g[8 * kp + 4 + 2 * rh + 1];  // group 2 kp + 1
                                const uint32_t s0 = odd ? a0 : b0, s1 = odd ? a1 : b1;                          // what the partner wants
                                const uint32_t r0 = __shfl_xor_sync(0xffffffffu, s0, 1), r1 = __shfl_xor_sync(0xffffffffu, s1, 1);
                                const float x0 = __uint_as_float(odd ? r0 : a0), x1 = __uint_as_float(odd ? r1 : a1);
                                const float x2 = __uint_as_float(odd ? b0 : r0), x3 = __uint_as_float(odd ? b1 : r1);
                                const int key = key0 + rbase + 8 * rh;
                                if (key < N)
                                    red_add_v4(gXb + (size_t)key * RT_D + cbase + 16 * kp, x0 * gx_mul, x1 * gx_mul, x2 * gx_mul, x3 * gx_mul);
                            }
                    };
                    emit(g0, 0);
                    emit(g1, 16);
                };

                uint32_t eit = it;
                for (int jl = 0; jl < ntl; ++jl, ++eit) {
                    const uint32_t buf = eit & 1;
                    mbar_wait(&m->s_full[buf], (eit >> 1) & 1);
                    RT_MARK(4);
                    tc_fence_after();
                    uint32_t v1[16], v2[16], w1[16], w2[16];
                    tmem_ld16(tmem + lane_base + COL_S + buf * 128 + 16 * half, v1);
                    tmem_ld16(tmem + lane_base + COL_S + buf * 128 + 32 + 16 * half, v2);
                    tmem_ld16(tmem + lane_base + COL_S + buf * 128 + 64 + 16 * half, w1);
                    tmem_ld16(tmem + lane_base + COL_S + buf * 128 + 96 + 16 * half, w2);
                    tmem_wait_ld();
                    tc_fence_before();
                    mbar_arrive(&m->s_free[buf]);
                    const bool live = (j0 + jl) * RT_KEYS + row < N && 16 * half < nrows;
                    uint32_t dh[8], dl[8], eh[8], el[8];
#pragma unroll
                    for (int s2i = 0; s2i < 8; ++s2i) {
                        float dsv[2], e1v[2];
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            const int s = 2 * s2i + u;
                            const float av = ((__uint_as_float(v1[s]) + __uint_as_float(w1[s])) - RT_XSCALE * RT_XSCALE) * a_mul;    // (dot - 1) / b^2
                            const float kap = ex2_approx(fminf(fmaxf(av, PRIFIT_LO), PRIFIT_HI) * RT_LOG2E);
                            const bool inr = (av >= PRIFIT_LO) && (av <= PRIFIT_HI);
                            const float dk = ((__uint_as_float(v2[s]) + __uint_as_float(w2[s])) * s2_mul - gmm_r[s]) * dinv_r[s];
                            dsv[u] = (inr && live) ? (kap * dk) * ds_mul : 0.f;
                            e1v[u] = live ? (kap * dinv_r[s]) * se : 0.f;
                        }
                        split2(dsv[0], dsv[1], dh[s2i], dl[s2i]);
                        split2(e1v[0], e1v[1], eh[s2i], el[s2i]);
                    }
                    RT_MARK(5);
                    mbar_wait(&m->c_free, (eit & 1) ^ 1);
                    RT_MARK(6);
                    uint8_t* crow = smem + BwdPlan::de + (uint32_t)row * 128u;
#pragma unroll
                    for (int c = 0; c < 2; ++c) {          // block 0: [ds_hi | ds_lo], block 1: [e1_hi | e1_lo]; hi = chunks 0-3, lo = chunks 4-7
                        const uint32_t oh = (uint32_t)(((2 * half + c) ^ (row & 7)) << 4);
                        const uint32_t ol = (uint32_t)(((4 + 2 * half + c) ^ (row & 7)) << 4);
                        *reinterpret_cast<uint4*>(crow + oh) = make_uint4(dh[4 * c], dh[4 * c + 1], dh[4 * c + 2], dh[4 * c + 3]);
                        *reinterpret_cast<uint4*>(crow + ol) = make_uint4(dl[4 * c], dl[4 * c + 1], dl[4 * c + 2], dl[4 * c + 3]);
                        *reinterpret_cast<uint4*>(crow + RT_PPART + oh) = make_uint4(eh[4 * c], eh[4 * c + 1], eh[4 * c + 2], eh[4 * c + 3]);
                        *reinterpret_cast<uint4*>(crow + RT_PPART + ol) = make_uint4(el[4 * c], el[4 * c + 1], el[4 * c + 2], el[4 * c + 3]);
                    }
                    fence_proxy_async();
                    mbar_arrive(&m->c_full);
                    RT_MARK(7);
                    if (jl > 0) flush(eit - 1, jl - 1);
                    RT_MARK(8);
                }
                if (ntl > 0) flush(eit - 1, ntl - 1);
                RT_MARK(9);
                // partial dL/dy^t of this CTA's keys (lane = column d), one per unit, added here; the coefficient tile is dead now
                float po[16];
#pragma unroll
                for (int s = 0; s < 16; ++s) po[s] = 0.f;
                if (ntl > 0) {
                    uint32_t v[16], w[16];
                    mbar_wait(&m->o_full, itn & 1);
                    tc_fence_after();
                    tmem_ld16(tmem + lane_base + COL_GY + 16 * half, v);
                    tmem_ld16(tmem + lane_base + COL_GY + 32 + 16 * half, w);
                    tmem_wait_ld();
#pragma unroll
                    for (int s = 0; s < 16; ++s) po[s] = (__uint_as_float(v[s]) + __uint_as_float(w[s])) * gx_mul;
                    if (ntl > G) {                                     // second unit of this CTA
                        tmem_ld16(tmem + lane_base + COL_GY + 64 + 16 * half, v);
                        tmem_ld16(tmem + lane_base + COL_GY + 64 + 32 + 16 * half, w);
                        tmem_wait_ld();
#pragma unroll
                        for (int s = 0; s < 16; ++s) po[s] += (__uint_as_float(v[s]) + __uint_as_float(w[s])) * gx_mul;
                    }
                    tc_fence_before();
                }
                RT_MARK(10);
                epi_bar();                                             // all flush() loads of this CTA are done with TMEM
#pragma unroll
                for (int s = 0; s < 16; ++s) part_o[(16 * half + s) * RT_D + row] = po[s];
                RT_MARK(11);
            }
            it += ntl;
            x_loaded = true;
            cluster.sync();
            RT_MARK(12);
            if (warp >= 4) {
                for (int f = et; f < rpc * 32; f += RT_EPI) {
                    const int orow = rank * rpc + (f >> 5), c4 = f & 31;
                    if (orow >= RT_SEEDS) break;
                    float4 oq[RT_MAXC];
#pragma unroll
                    for (int q = 0; q < RT_MAXC; ++q)
                        oq[q] = q < csize ? reinterpret_cast<const float4*>(cluster.map_shared_rank(part_o, q) + orow * RT_D)[c4]
                                          : make_float4(0.f, 0.f, 0.f, 0.f);
                    float ox[RT_MAXC], oy[RT_MAXC], oz[RT_MAXC], ow[RT_MAXC];
#pragma unroll
                    for (int q = 0; q < RT_MAXC; ++q) { ox[q] = oq[q].x; oy[q] = oq[q].y; oz[q] = oq[q].z; ow[q] = oq[q].w; }
                    const float4 o = make_float4(unit_sum(ox, csize), unit_sum(oy, csize), unit_sum(oz, csize), unit_sum(ow, csize));
                    for (int q = 0; q < csize; ++q) reinterpret_cast<float4*>(cluster.map_shared_rank(gy, q) + orow * RT_D)[c4] = o;
                }
            }
            RT_MARK(13);
            cluster.sync();
            RT_MARK(14);
        }
        // dL/dy^0 lands on the seed's own row of X (new_X = X.clone(), gather by idx)
        __threadfence();
        cluster.sync();
        if (owner && myrow < nrows) {
            const int src = a.idx[(size_t)b * Kcap + k0 + myrow];
            for (int col = tc; col < RT_D; col += tpr) atomicAdd(gXb + (size_t)src * RT_D + col, gy[myrow * RT_D + col]);
        }
        __threadfence();
        cluster.sync();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

// Keys of a shape are split over a cluster of 4 or 8 CTAs (one CTA per SM: shared memory).  The result does not depend on
// the choice (units, see my_tiles), so it is a pure scheduling decision.  8 has 2/3 of the per-iteration latency (2 resident
// key tiles per CTA at N = 2048; 8 shapes: forward 124 -> 79 us, backward 257 -> 185 us) as long as all clusters are
// co-resident: a 24-shape batch would be 192 CTAs in two rounds (375 us backward), and three graph branches of 8 shapes
// launching 8-CTA clusters side by side lose too (a cluster needs 8 free SMs inside one GPC).  Hence: 8 when the launch is
// at most half the GPU, else 4; `wide` (PRIFIT_ROWS_WIDE / _NARROW, 1 / 0, -1 = automatic) and PRIFIT_ROWS_CLUSTER=4|8
// (diagnostics) override.
int pick_cluster(int wide, int n_clusters) {
    int csize = wide < 0 ? (8 * n_clusters <= 74 ? 8 : 4) : (wide ? 8 : 4);
    if (const char* e = getenv("PRIFIT_ROWS_CLUSTER")) {
        const int v = atoi(e);
        if (v == 4 || v == 8) csize = v;
    }
    return csize;
}

template <typename Kern>
int launch_rows_tc(Kern kern, size_t smem, int csize, dim3 grid, const CUtensorMap& map, const RowsArgs& a, cudaStream_t st) {
    PF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(RT_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    PF_CUDA(cudaLaunchKernelEx(&cfg, kern, map, a));
    return 0;
}

}  // namespace

// diagnostics: copy the phase timeline recorded by the last launch made with PRIFIT_ROWS_TIMELINE=1
extern "C" int prifit_debug_rows_timeline(long long* host_pairs, int max_pairs) {
    int n = 0;
    if (cudaMemcpyFromSymbol(&n, g_rt_dbg_n, sizeof(int)) != cudaSuccess) return -1;
    n = n < max_pairs ? n : max_pairs;
    if (n > 0 && cudaMemcpyFromSymbol(host_pairs, g_rt_dbg, (size_t)n * 2 * sizeof(long long)) != cudaSuccess) return -1;
    return n;
}

static int rows_dbg_begin() {
    const char* e = getenv("PRIFIT_ROWS_TIMELINE");        // bit 0: record the phase timeline; bit 1: skip the gX flush (timing experiment)
    if (!e || atoi(e) == 0) return 0;
    int zero = 0;
    if (atoi(e) & 1) cudaMemcpyToSymbol(g_rt_dbg_n, &zero, sizeof(int));
    return atoi(e);
}

size_t prifit_rows_tc_workspace_bytes(int B, int N) { return (size_t)2 * B * N * RT_D * sizeof(__half) + 256; }

int prifit_rows_tc_fwd(const float* X, const float* bw, const int32_t* idx, const int32_t* K, int B, int N, int T, int Kcap,
                       float* traj, float* stat, float* C_out, void* ws, bool ws_holds_split, int wide, cudaStream_t st) {
    __half* Xs = reinterpret_cast<__half*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
    CUtensorMap map;
    int rc = ws_holds_split ? prifit_tc_make_tile_map(&map, Xs, 2 * B, N) : prifit_tc_split_rows(X, Xs, B, N, &map, st);
    if (rc) return rc;
    RowsArgs a = {};
    a.X = X; a.bw = bw; a.idx = idx; a.K = K; a.N = N; a.B = B; a.T = T; a.Kcap = Kcap;
    a.traj = traj; a.stat = stat; a.C_out = C_out;
    a.dbg = rows_dbg_begin();
    const int csize = pick_cluster(wide, B * ((Kcap + RT_SEEDS - 1) / RT_SEEDS));
    return launch_rows_tc(rows_tc_fwd_kernel, 1024 + FwdPlan::total, csize, dim3(csize, (Kcap + RT_SEEDS - 1) / RT_SEEDS, B), map, a, st);
}

// the split fp16 rows of X into the workspace, ahead of the forward / backward calls that then pass PRIFIT_ROWS_WS_HOLDS_SPLIT
int prifit_rows_tc_prepare(const float* X, int B, int N, void* ws, cudaStream_t st) {
    __half* Xs = reinterpret_cast<__half*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
    CUtensorMap map;
    return prifit_tc_split_rows(X, Xs, B, N, &map, st);
}

int prifit_rows_tc_bwd(const float* X, const float* bw, const int32_t* idx, const int32_t* K, const float* traj,
                       const float* stat, const float* gC, int B, int N, int T, int Kcap, float* gX, void* ws, bool ws_holds_split,
                       int wide, cudaStream_t st) {
    __half* Xs = reinterpret_cast<__half*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
    CUtensorMap map;
    // the forward call left the split rows of the same X in this workspace: only the tile map is rebuilt (host side)
    int rc = ws_holds_split ? prifit_tc_make_tile_map(&map, Xs, 2 * B, N) : prifit_tc_split_rows(X, Xs, B, N, &map, st);
    if (rc) return rc;
    RowsArgs a = {};
    a.X = X; a.bw = bw; a.idx = idx; a.K = K; a.N = N; a.B = B; a.T = T; a.Kcap = Kcap;
    a.traj_in = traj; a.stat_in = stat; a.gC = gC; a.gX = gX;
    a.dbg = rows_dbg_begin();
    const int csize = pick_cluster(wide, B);
    return launch_rows_tc(rows_tc_bwd_kernel, 1024 + BwdPlan::total, csize, dim3(csize, 1, B), map, a, st);
}
