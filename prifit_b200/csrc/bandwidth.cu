// k1 -- mean-shift bandwidth: mean over sample rows of sqrt(k-th smallest (2 - 2 x_i.x_j)).
// reference src/mean_shift.py:138-160 (compute_bandwidth); guard_sqrt src/guard.py:13-18.
//
// fp32 CUDA-core version.  A CTA owns R sample rows; their full distance rows (R x n_s fp32) live
// in shared memory while the key rows stream through a 64-key tile, so the n_s x n_s matrix never
// reaches HBM.  The k-th order statistic is an exact 4 x 8-bit radix select on order-preserving
// integer keys (one warp per row), so the result does not depend on any sort's tie handling.
#include <cuda_fp16.h>
#include "common.cuh"

int prifit_tc_bandwidth_rows(const float* X, int B, int N, const int32_t* kth, __half* Xs_ws, int2* rowinfo_ws,
                             float* rowval, int32_t* overflow, cudaStream_t st);
size_t prifit_tc_gram_split_bytes(int B, int N);
int prifit_gram_engine();

namespace {

constexpr int BW_THREADS = 256;
constexpr int BW_TILE = 64;  // keys per tile

template <int R>
__global__ void __launch_bounds__(BW_THREADS) bandwidth_rows_kernel(
    const float* __restrict__ X, int N, int d, const int32_t* __restrict__ rows, int n_s,
    const int32_t* __restrict__ kth, float* __restrict__ rowval /*[B,n_s]*/, const int32_t* __restrict__ only_if,
    int tiles_x, int items) {
    extern __shared__ __align__(16) float smem[];
    if (only_if && *only_if == 0) return;     // exact fallback of the tensor-core path: runs only on overflow
    // work item = (shape, block of R rows); the fallback launch uses one CTA per SM looping over the items, so that
    // the normal case (flag clear) costs 148 empty CTAs instead of one empty CTA per item
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
    __syncthreads();
    const int ld = d + 4;
    float* dist = smem;                       // [R][n_s]
    float* ys = dist + (size_t)R * n_s;       // [R][ld]
    float* xs = ys + R * ld;                  // [BW_TILE][ld]
    uint32_t* hist = reinterpret_cast<uint32_t*>(xs + BW_TILE * ld);  // [8][256]

    const int b = item / tiles_x;
    const int r0 = (item - b * tiles_x) * R;
    const float* Xb = X + (size_t)b * N * d;
    const int32_t* rb = rows ? rows + (size_t)b * n_s : nullptr;
    const int tid = threadIdx.x;
    const int nv = d >> 2;

    // stage the R query rows
    for (int e = tid; e < R * nv; e += BW_THREADS) {
        const int r = e / nv, c = e - r * nv;
        const int i = r0 + r;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < n_s) {
            const int src = rb ? rb[i] : i;
            v = reinterpret_cast<const float4*>(Xb + (size_t)src * d)[c];
        }
        *reinterpret_cast<float4*>(ys + r * ld + 4 * c) = v;
    }

    constexpr int CPT = R / 4;            // columns per thread (R=16:4, 8:2, 4:1)
    constexpr int TXN = BW_TILE / CPT;    // threads per row
    const int ty = tid / TXN, tx = tid - ty * TXN;

    for (int j0 = 0; j0 < n_s; j0 += BW_TILE) {
        __syncthreads();
        for (int e = tid; e < BW_TILE * nv; e += BW_THREADS) {
            const int r = e / nv, c = e - r * nv;
            const int j = j0 + r;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (j < n_s) {
                const int src = rb ? rb[j] : j;
                v = reinterpret_cast<const float4*>(Xb + (size_t)src * d)[c];
            }
            *reinterpret_cast<float4*>(xs + r * ld + 4 * c) = v;
        }
        __syncthreads();
        float acc[CPT];
#pragma unroll
        for (int a = 0; a < CPT; ++a) acc[a] = 0.f;
        const float* yrow = ys + ty * ld;
        for (int i = 0; i < d; i += 4) {
            const float4 y = *reinterpret_cast<const float4*>(yrow + i);
#pragma unroll
            for (int a = 0; a < CPT; ++a) {
                const float4 x = *reinterpret_cast<const float4*>(xs + (tx + TXN * a) * ld + i);
                acc[a] = fmaf(y.x, x.x, acc[a]);
                acc[a] = fmaf(y.y, x.y, acc[a]);
                acc[a] = fmaf(y.z, x.z, acc[a]);
                acc[a] = fmaf(y.w, x.w, acc[a]);
            }
        }
#pragma unroll
        for (int a = 0; a < CPT; ++a) {
            const int j = j0 + tx + TXN * a;
            if (j < n_s) dist[(size_t)ty * n_s + j] = 2.0f - 2.0f * acc[a];   // line 153: 2 - 2 X X^T
        }
    }
    __syncthreads();

    // exact k-th smallest per row: one warp per row, 4 passes of 8 bits
    const int warp = tid >> 5, lane = tid & 31;
    uint32_t* h = hist + warp * 256;
    const int k_target = max(1, min(kth[b], n_s));   // 1-based rank
    for (int r = warp; r < R; r += BW_THREADS / 32) {
        const int i = r0 + r;
        if (i >= n_s) continue;
        const float* drow = dist + (size_t)r * n_s;
        uint32_t prefix = 0, mask = 0;
        int remaining = k_target;
        for (int shift = 24; shift >= 0; shift -= 8) {
            for (int q = lane; q < 256; q += 32) h[q] = 0;
            __syncwarp();
            for (int j = lane; j < n_s; j += 32) {
                const uint32_t key = float_to_ordered(drow[j]);
                if ((key & mask) == prefix) atomicAdd(&h[(key >> shift) & 255u], 1u);
            }
            __syncwarp();
            // each lane owns 8 consecutive bins
            uint32_t c[8], tot = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) { c[q] = h[lane * 8 + q]; tot += c[q]; }
            uint32_t incl = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += n;
            }
            const uint32_t excl = incl - tot;
            int found_bin = -1, found_before = 0;
            if ((uint32_t)remaining > excl && (uint32_t)remaining <= incl) {
                uint32_t run = excl;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    if (found_bin < 0 && (uint32_t)remaining <= run + c[q]) { found_bin = lane * 8 + q; found_before = (int)run; }
                    run += c[q];
                }
            }
            const uint32_t ball = __ballot_sync(0xffffffffu, found_bin >= 0);
            const int src = __ffs(ball) - 1;
            found_bin = __shfl_sync(0xffffffffu, found_bin, src);
            found_before = __shfl_sync(0xffffffffu, found_before, src);
            prefix |= ((uint32_t)found_bin) << shift;
            mask |= 255u << shift;
            remaining -= found_before;
            __syncwarp();
        }
        if (lane == 0) {
            const float v = ordered_to_float(prefix);
            rowval[(size_t)b * n_s + i] = sqrtf(fmaxf(v, 1e-6f));     // guard_sqrt(., 1e-6), line 158
        }
    }
    }
}

__global__ void __launch_bounds__(256) bandwidth_mean_kernel(const float* __restrict__ rowval, int n_s,
                                                             float* __restrict__ bw_out) {
    __shared__ float red[32];
    const int b = blockIdx.x;
    float v[1] = {0.f};
    for (int i = threadIdx.x; i < n_s; i += blockDim.x) v[0] += rowval[(size_t)b * n_s + i];
    block_sum<1>(v, red);
    if (threadIdx.x == 0) bw_out[b] = v[0] / (float)n_s;               // torch.mean, line 160
}

size_t bw_smem_bytes(int R, int n_s, int d) {
    return ((size_t)R * n_s + (size_t)(R + BW_TILE) * (d + 4)) * sizeof(float) + 8 * 256 * sizeof(uint32_t);
}

template <int R>
int launch_rows(const float* X, int B, int N, int d, const int32_t* rows, int n_s, const int32_t* kth,
                float* rowval, const int32_t* only_if, cudaStream_t st) {
    const size_t smem = bw_smem_bytes(R, n_s, d);
    PF_CUDA(cudaFuncSetAttribute(bandwidth_rows_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int tiles_x = (n_s + R - 1) / R, items = tiles_x * B;
    const int grid = only_if ? (items < 148 ? items : 148) : items;
    bandwidth_rows_kernel<R><<<grid, BW_THREADS, smem, st>>>(X, N, d, rows, n_s, kth, rowval, only_if, tiles_x, items);
    PF_LAUNCH_CHECK();
    return 0;
}

}  // namespace

extern "C" size_t prifit_bandwidth_workspace_bytes(int B, int N, int d, int n_s) {
    // row values | overflow flag | (tensor-core path) per-row (window, rank) | split fp16 rows (hi + lo)
    return (size_t)B * (size_t)n_s * sizeof(float) + 512 + (size_t)B * N * sizeof(int2) + 512 +
           (d == 128 ? prifit_tc_gram_split_bytes(B, N) : 0);
}

extern "C" int prifit_bandwidth_fwd(const float* X, int B, int N, int d, const int32_t* rows, int n_s,
                                    const int32_t* kth, float* bw_out, void* ws, size_t ws_bytes, void* stream) {
    PF_CHECK_ARG(X && kth && bw_out && ws, PRIFIT_E_BADARG, "null pointer");
    PF_CHECK_ARG(B > 0 && N > 0 && n_s > 0 && n_s <= N, PRIFIT_E_BADARG, "need 0 < n_s <= N");
    PF_CHECK_ARG(rows || n_s == N, PRIFIT_E_BADARG, "rows == NULL requires n_s == N");
    PF_CHECK_ARG(d % 4 == 0 && d >= 4 && d <= 512, PRIFIT_E_SHAPE, "d must be a multiple of 4, <= 512");
    PF_CHECK_ARG(ws_bytes >= prifit_bandwidth_workspace_bytes(B, N, d, n_s), PRIFIT_E_WS, "workspace too small");
    const size_t limit = 227 * 1024;
    float* rowval = static_cast<float*>(ws);
    cudaStream_t st = pf_stream(stream);
    int rc;
    const int32_t* only_if = nullptr;
    if (!rows && d == 128 && N < 65536 && prifit_gram_engine() == 0) {
        // tensor-core path (gram_tc.cu): two histogram passes + candidate pass with exact fp32 refinement
        uint8_t* p = static_cast<uint8_t*>(ws) + (size_t)B * n_s * sizeof(float);
        int32_t* overflow = reinterpret_cast<int32_t*>((reinterpret_cast<uintptr_t>(p) + 15) & ~(uintptr_t)15);
        int2* rowinfo = reinterpret_cast<int2*>((reinterpret_cast<uintptr_t>(overflow) + 16 + 255) & ~(uintptr_t)255);
        __half* Xh = reinterpret_cast<__half*>((reinterpret_cast<uintptr_t>(rowinfo + (size_t)B * N) + 255) & ~(uintptr_t)255);
        PF_CUDA(cudaMemsetAsync(overflow, 0, 4 * sizeof(int32_t), st));      // [0] overflow, [1] level-1 request
        rc = prifit_tc_bandwidth_rows(X, B, N, kth, Xh, rowinfo, rowval, overflow, st);
        if (rc) return rc;
        only_if = overflow;      // the exact CUDA-core kernel below re-does the batch only if a candidate list overflowed
    }
    if (bw_smem_bytes(16, n_s, d) <= limit) rc = launch_rows<16>(X, B, N, d, rows, n_s, kth, rowval, only_if, st);
    else if (bw_smem_bytes(8, n_s, d) <= limit) rc = launch_rows<8>(X, B, N, d, rows, n_s, kth, rowval, only_if, st);
    else if (bw_smem_bytes(4, n_s, d) <= limit) rc = launch_rows<4>(X, B, N, d, rows, n_s, kth, rowval, only_if, st);
    else if (only_if) rc = 0;
    else { prifit_set_error("prifit_bandwidth_fwd: n_s=%d too large for the shared-memory row block", n_s); return PRIFIT_E_SHAPE; }
    if (rc) return rc;
    bandwidth_mean_kernel<<<B, 256, 0, st>>>(rowval, n_s, bw_out);
    PF_LAUNCH_CHECK();
    return 0;
}
