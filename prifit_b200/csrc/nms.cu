// k3 -- non-maximum suppression of the shifted seeds + hard labels, entirely on the device.
// reference src/mean_shift.py:162-202, called as nms(new_X, new_X, bw) (src/mean_shift.py:44):
//   1. nearest[j] = argmin_i (2 - 2 <c_i, x_j>)                         (:169-172)
//   2. votes[i]   = #{j : nearest[j] == i}            (np.unique on the host in the reference, :175-184)
//   3. for every i with votes[i] > 0:  best[i] = argmax_j ([2 - 2 <c_i, c_j> < bw] * votes[j])   (:187-194)
//   4. ids = sorted unique best[i]; centres = c[ids]                     (:194-196)
//   5. labels[j] = argmax_k <centres_k, x_j>                             (:200-201)
// All arg-reductions return the LOWEST index among equal values (torch CPU semantics; SURVEY B).
// The reference needs a D2H copy + numpy for step 2; here nothing leaves the GPU.
//
// fp32 CUDA-core Gram tiles (row groups of 32 against a resident 128-key tile).
#include <cuda.h>
#include <cuda_fp16.h>
#include "rowgemm.cuh"

int prifit_tc_nms_nearest(const float* newX, int B, int N, __half* Xh_ws, CUtensorMap* map_out, int32_t* nearest, cudaStream_t st);
int prifit_tc_nms_best(const CUtensorMap* map, const __half* Xs, const float* bw, const int32_t* votes,
                       const int32_t* rowsel, const int32_t* nrows, int B, int N, int small_rows, int32_t* best, cudaStream_t st);
size_t prifit_tc_gram_split_bytes(int B, int N);
int prifit_gram_engine();

namespace {

// lexicographic "better" for arg-min: smaller value, then smaller index
__device__ __forceinline__ bool lt_min(float v, int i, float bv, int bi) { return v < bv || (v == bv && i < bi); }
// for arg-max: larger value, then smaller index
__device__ __forceinline__ bool gt_max(float v, int i, float bv, int bi) { return v > bv || (v == bv && i < bi); }

// MODE 0: nearest[j] = argmin_i dist(i, j)                         over all rows i
// MODE 1: best[j]    = argmax_i ((dist(i, j) < bw) ? votes[i] : 0) over all rows i   (j plays the role of "u")
template <int D, int MODE>
__global__ void __launch_bounds__(RG_THREADS) nms_gram_kernel(
    const float* __restrict__ Xn, const float* __restrict__ bw, int N, const int32_t* __restrict__ votes,
    int32_t* __restrict__ out) {
    constexpr int LD = D + 4;
    extern __shared__ __align__(16) float smem[];
    float* ys = smem;                      // 32-row group
    float* xs = ys + RG_ROWS * LD;         // resident 128 keys (columns)
    float* cv = xs + RG_KEYS * LD;         // [8][128] cross-warp candidates
    int* ci = reinterpret_cast<int*>(cv + 8 * RG_KEYS);
    const int b = blockIdx.y, j0 = blockIdx.x * RG_KEYS;
    const float* Xb = Xn + (size_t)b * N * D;
    const int tid = threadIdx.x, ty = tid >> 5, tx = tid & 31;
    const float bwv = bw[b];
    const int32_t* vb = votes ? votes + (size_t)b * N : nullptr;

    rg_load_rows<D>(xs, RG_KEYS, Xb, [&](int r) -> long long { return j0 + r < N ? (long long)(j0 + r) : -1; });
    float bestv[4];
    int besti[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) { bestv[c] = MODE == 0 ? INFINITY : -1.0f; besti[c] = 0; }

    for (int i0 = 0; i0 < N; i0 += RG_ROWS) {
        __syncthreads();
        rg_load_rows<D>(ys, RG_ROWS, Xb, [&](int r) -> long long { return i0 + r < N ? (long long)(i0 + r) : -1; });
        __syncthreads();
        float acc[4][4];
        rg_dot_32x128<D>(ys, xs, acc);
#pragma unroll
        for (int a = 0; a < 4; ++a) {       // rows ascending within the thread
            const int i = i0 + ty + 8 * a;
            if (i < N) {
                float vote = 0.f;
                if (MODE == 1) vote = (float)vb[i];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float dist = 2.0f - 2.0f * acc[a][c];
                    if (MODE == 0) {
                        if (dist < bestv[c]) { bestv[c] = dist; besti[c] = i; }
                    } else {
                        const float v = dist < bwv ? vote : 0.f;
                        if (v > bestv[c]) { bestv[c] = v; besti[c] = i; }
                    }
                }
            }
        }
    }
    // combine the 8 warps (row residues) per column
#pragma unroll
    for (int c = 0; c < 4; ++c) { cv[ty * RG_KEYS + tx + 32 * c] = bestv[c]; ci[ty * RG_KEYS + tx + 32 * c] = besti[c]; }
    __syncthreads();
    if (tid < RG_KEYS && j0 + tid < N) {
        float bv = cv[tid];
        int bi = ci[tid];
        for (int w = 1; w < 8; ++w) {
            const float v = cv[w * RG_KEYS + tid];
            const int i = ci[w * RG_KEYS + tid];
            if (MODE == 0 ? lt_min(v, i, bv, bi) : gt_max(v, i, bv, bi)) { bv = v; bi = i; }
        }
        out[(size_t)b * N + j0 + tid] = bi;
    }
}

__global__ void nms_vote_kernel(const int32_t* __restrict__ nearest, int N, int32_t* __restrict__ votes) {
    const int b = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < N) atomicAdd(&votes[(size_t)b * N + nearest[(size_t)b * N + j]], 1);
}

__global__ void nms_flag_kernel(const int32_t* __restrict__ votes, const int32_t* __restrict__ best, int N,
                                int32_t* __restrict__ flags) {
    const int b = blockIdx.y, u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u < N && votes[(size_t)b * N + u] > 0) flags[(size_t)b * N + best[(size_t)b * N + u]] = 1;
}

// ascending compaction of the flagged indices (== torch.unique ordering); one CTA per shape
__global__ void __launch_bounds__(1024) nms_compact_kernel(const int32_t* __restrict__ flags, int N, int Kcap,
                                                           int32_t* __restrict__ idx_out, int32_t* __restrict__ K_out) {
    __shared__ int wsum[32];
    __shared__ int base_s;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) base_s = 0;
    for (int k = tid; k < Kcap; k += blockDim.x) idx_out[(size_t)b * Kcap + k] = -1;
    __syncthreads();
    for (int c0 = 0; c0 < N; c0 += blockDim.x) {
        const int i = c0 + tid;
        const int f = (i < N && flags[(size_t)b * N + i]) ? 1 : 0;
        const unsigned ball = __ballot_sync(0xffffffffu, f);
        const int wpre = __popc(ball & ((1u << lane) - 1u));
        if (lane == 0) wsum[warp] = __popc(ball);
        __syncthreads();
        int before = base_s;
        for (int w = 0; w < warp; ++w) before += wsum[w];
        const int pos = before + wpre;
        if (f && pos < Kcap) idx_out[(size_t)b * Kcap + pos] = i;
        __syncthreads();
        if (tid == 0) {
            int tot = 0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += wsum[w];
            base_s += tot;
        }
        __syncthreads();
    }
    if (tid == 0) K_out[b] = base_s;
}

// labels[j] = argmax_k <newX[idx[k]], newX[j]>, lowest k on ties; used[b][k] = 1 if some point took label k.
// idx is the FULL ascending list of centres (stride N): a shape may have more centres than the padded capacity of the
// differentiable stages and still pass the guard, which counts distinct labels (src/ellipsoid_utils.py:23); the label pass
// therefore runs over every centre, like the reference's (src/mean_shift.py:200-201).
template <int D>
__global__ void __launch_bounds__(RG_THREADS) nms_label_kernel(
    const float* __restrict__ Xn, int N, const int32_t* __restrict__ idx, const int32_t* __restrict__ Kfound,
    int32_t* __restrict__ labels, int32_t* __restrict__ used) {
    constexpr int LD = D + 4;
    extern __shared__ __align__(16) float smem[];
    float* ys = smem;
    float* xs = ys + RG_ROWS * LD;
    float* cv = xs + RG_KEYS * LD;
    int* ci = reinterpret_cast<int*>(cv + 8 * RG_KEYS);
    const int b = blockIdx.y, j0 = blockIdx.x * RG_KEYS;
    const float* Xb = Xn + (size_t)b * N * D;
    const int tid = threadIdx.x, ty = tid >> 5, tx = tid & 31;
    const int Kb = Kfound[b];
    const int32_t* idx_b = idx + (size_t)b * N;

    rg_load_rows<D>(xs, RG_KEYS, Xb, [&](int r) -> long long { return j0 + r < N ? (long long)(j0 + r) : -1; });
    float bestv[4];
    int besti[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) { bestv[c] = -INFINITY; besti[c] = 0; }
    for (int k0 = 0; k0 < Kb; k0 += RG_ROWS) {
        __syncthreads();
        rg_load_rows<D>(ys, RG_ROWS, Xb, [&](int r) -> long long { return k0 + r < Kb ? (long long)idx_b[k0 + r] : -1; });
        __syncthreads();
        float acc[4][4];
        rg_dot_32x128<D>(ys, xs, acc, min(RG_ROWS, Kb - k0));
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int k = k0 + ty + 8 * a;
            if (k < Kb) {
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (acc[a][c] > bestv[c]) { bestv[c] = acc[a][c]; besti[c] = k; }
            }
        }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) { cv[ty * RG_KEYS + tx + 32 * c] = bestv[c]; ci[ty * RG_KEYS + tx + 32 * c] = besti[c]; }
    __syncthreads();
    if (tid < RG_KEYS && j0 + tid < N) {
        float bv = cv[tid];
        int bi = ci[tid];
        for (int w = 1; w < 8; ++w) {
            const float v = cv[w * RG_KEYS + tid];
            const int i = ci[w * RG_KEYS + tid];
            if (gt_max(v, i, bv, bi)) { bv = v; bi = i; }
        }
        labels[(size_t)b * N + j0 + tid] = bi;
        used[(size_t)b * N + bi] = 1;
    }
}

// n_labels[b] = number of distinct labels (exact, also when there are more centres than Kcap); idx_out[b][:Kcap] = the
// first Kcap centres of the full list, -1 padded
__global__ void __launch_bounds__(256) nms_nlabels_kernel(const int32_t* __restrict__ used, const int32_t* __restrict__ Kfound,
                                                          const int32_t* __restrict__ idx_full, int N, int Kcap,
                                                          int32_t* __restrict__ idx_out, int32_t* __restrict__ n_labels) {
    const int b = blockIdx.x;
    const int Kf = Kfound[b];
    int v = 0;
    for (int k = threadIdx.x; k < Kf; k += blockDim.x) v += used[(size_t)b * N + k] ? 1 : 0;
    for (int k = threadIdx.x; k < Kcap; k += blockDim.x) idx_out[(size_t)b * Kcap + k] = k < Kf ? idx_full[(size_t)b * N + k] : -1;
    __shared__ int s[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 8; ++w) t += s[w];
        n_labels[b] = t;
    }
}

// idx_out[b][:Kcap] = the first Kcap centres of the full list, -1 padded (the centres-only call: nms_nlabels_kernel writes the
// same values when the labels are computed in the same call)
__global__ void nms_idx_kernel(const int32_t* __restrict__ Kfound, const int32_t* __restrict__ idx_full, int N, int Kcap,
                               int32_t* __restrict__ idx_out) {
    const int b = blockIdx.x, Kf = Kfound[b];
    for (int k = threadIdx.x; k < Kcap; k += blockDim.x) idx_out[(size_t)b * Kcap + k] = k < Kf ? idx_full[(size_t)b * N + k] : -1;
}

// NMS step 3 when few rows received votes (the usual case: one row per mode): votes are zero outside the voted rows and a
// voted row is within the threshold of itself, so the arg-max over ALL columns of [dist < bw] * votes is attained on a voted
// column -- an n x n problem over the n voted rows (n <= NMS_SMALL) instead of a Gram pass of n rows against all N columns
// through 128-row tensor-core tiles (27 us of latency for 16 real rows).  One CTA per shape, fp32 dot products, lowest index
// among equal values.  Shapes with more voted rows are left to the Gram pass (which skips the shapes done here).
constexpr int NMS_SMALL = 64;
template <int D>
__global__ void __launch_bounds__(256) nms_best_small_kernel(const float* __restrict__ Xn, const float* __restrict__ bw,
                                                             const int32_t* __restrict__ votes, const int32_t* __restrict__ rowsel,
                                                             const int32_t* __restrict__ nrows, int N, int32_t* __restrict__ best) {
    constexpr int LD = D + 1;                                   // odd stride: lane i reads row i conflict-free
    extern __shared__ float rows[];                             // [n][LD]
    __shared__ int sel[NMS_SMALL];
    __shared__ float vote[NMS_SMALL];
    const int b = blockIdx.x, n = nrows[b];
    if (n > NMS_SMALL || n <= 0) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float* Xb = Xn + (size_t)b * N * D;
    if (tid < n) { sel[tid] = rowsel[(size_t)b * N + tid]; vote[tid] = (float)votes[(size_t)b * N + sel[tid]]; }
    __syncthreads();
    for (int e = tid; e < n * (D / 4); e += 256) {
        const int r = e / (D / 4), c = e % (D / 4);
        const float4 v = reinterpret_cast<const float4*>(Xb + (size_t)sel[r] * D)[c];
        float* dst = rows + r * LD + 4 * c;
        dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
    }
    __syncthreads();
    const float bwv = bw[b];
    for (int u = warp; u < n; u += 8) {
        float bv = -1.0f;
        int bi = 0x7fffffff;
        for (int i = lane; i < n; i += 32) {                    // i ascending per lane: '>' keeps the lowest index
            float dot = 0.f;
#pragma unroll 8
            for (int c = 0; c < D; ++c) dot = fmaf(rows[u * LD + c], rows[i * LD + c], dot);
            const float v = (2.0f - 2.0f * dot) < bwv ? vote[i] : 0.f;
            if (v > bv) { bv = v; bi = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (gt_max(ov, oi, bv, bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) best[(size_t)b * N + sel[u]] = sel[bi];
    }
}

struct NmsWs {
    int32_t *nearest, *votes, *best, *flags, *used, *rowsel, *nrows, *idx_full;
};

NmsWs carve(void* ws, int B, int N) {
    NmsWs w;
    int32_t* p = static_cast<int32_t*>(ws);
    const size_t bn = (size_t)B * N;
    w.votes = p;                 // votes, flags, used are contiguous: one memset
    w.flags = p + bn;
    w.used = p + 2 * bn;
    w.nearest = p + 3 * bn;
    w.best = w.nearest + bn;
    w.rowsel = w.best + bn;
    w.idx_full = w.rowsel + bn;
    w.nrows = w.idx_full + bn;
    return w;
}

template <int D>
int launch_nms(const float* newX, const float* bw, int B, int N, int Kcap, int32_t* idx_out, int32_t* K_out,
               int32_t* labels_out, int32_t* n_labels_out, void* ws, cudaStream_t st) {
    NmsWs w = carve(ws, B, N);
    const size_t smem = ((size_t)(RG_ROWS + RG_KEYS) * (D + 4) + 16 * RG_KEYS) * sizeof(float);
    PF_CUDA(cudaFuncSetAttribute(nms_gram_kernel<D, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PF_CUDA(cudaFuncSetAttribute(nms_gram_kernel<D, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PF_CUDA(cudaFuncSetAttribute(nms_label_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PF_CUDA(cudaMemsetAsync(w.votes, 0, 3 * (size_t)B * N * sizeof(int32_t), st));
    dim3 gt((N + RG_KEYS - 1) / RG_KEYS, B), ge((N + 255) / 256, B);
    const bool tc = D == 128 && prifit_gram_engine() == 0;      // Gram passes on the tensor cores (gram_tc.cu)
    __half* Xh = reinterpret_cast<__half*>((reinterpret_cast<uintptr_t>(w.nrows + B) + 255) & ~(uintptr_t)255);
    CUtensorMap map;
    if (tc) {
        int rc = prifit_tc_nms_nearest(newX, B, N, Xh, &map, w.nearest, st);
        if (rc) return rc;
    } else {
        nms_gram_kernel<D, 0><<<gt, RG_THREADS, smem, st>>>(newX, bw, N, nullptr, w.nearest);
        PF_LAUNCH_CHECK();
    }
    nms_vote_kernel<<<ge, 256, 0, st>>>(w.nearest, N, w.votes);
    PF_LAUNCH_CHECK();
    if (tc) {
        // step 3 only concerns the rows that received votes (nbrs[uniques], src/mean_shift.py:192-194):
        // compact them (ascending) and run the Gram pass over those rows alone
        nms_compact_kernel<<<B, 1024, 0, st>>>(w.votes, N, N, w.rowsel, w.nrows);
        PF_LAUNCH_CHECK();
        const size_t sm_small = (size_t)NMS_SMALL * (D + 1) * sizeof(float);
        PF_CUDA(cudaFuncSetAttribute(nms_best_small_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_small));
        nms_best_small_kernel<D><<<B, 256, sm_small, st>>>(newX, bw, w.votes, w.rowsel, w.nrows, N, w.best);
        PF_LAUNCH_CHECK();
        int rc = prifit_tc_nms_best(&map, Xh, bw, w.votes, w.rowsel, w.nrows, B, N, NMS_SMALL, w.best, st);
        if (rc) return rc;
    } else {
        nms_gram_kernel<D, 1><<<gt, RG_THREADS, smem, st>>>(newX, bw, N, w.votes, w.best);
        PF_LAUNCH_CHECK();
    }
    nms_flag_kernel<<<ge, 256, 0, st>>>(w.votes, w.best, N, w.flags);
    PF_LAUNCH_CHECK();
    nms_compact_kernel<<<B, 1024, 0, st>>>(w.flags, N, N, w.idx_full, K_out);
    PF_LAUNCH_CHECK();
    if (!labels_out) {          // centres only: the label pass follows as prifit_nms_labels (possibly on another stream)
        nms_idx_kernel<<<B, 64, 0, st>>>(K_out, w.idx_full, N, Kcap, idx_out);
        PF_LAUNCH_CHECK();
        return 0;
    }
    nms_label_kernel<D><<<gt, RG_THREADS, smem, st>>>(newX, N, w.idx_full, K_out, labels_out, w.used);
    PF_LAUNCH_CHECK();
    nms_nlabels_kernel<<<B, 256, 0, st>>>(w.used, K_out, w.idx_full, N, Kcap, idx_out, n_labels_out);
    PF_LAUNCH_CHECK();
    return 0;
}

// step 5 alone, after a centres-only launch_nms on the same workspace (full centre list and zeroed `used` flags are there)
template <int D>
int launch_nms_labels(const float* newX, int B, int N, int Kcap, const int32_t* K, int32_t* idx_out, int32_t* labels_out,
                      int32_t* n_labels_out, void* ws, cudaStream_t st) {
    NmsWs w = carve(ws, B, N);
    const size_t smem = ((size_t)(RG_ROWS + RG_KEYS) * (D + 4) + 16 * RG_KEYS) * sizeof(float);
    PF_CUDA(cudaFuncSetAttribute(nms_label_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 gt((N + RG_KEYS - 1) / RG_KEYS, B);
    nms_label_kernel<D><<<gt, RG_THREADS, smem, st>>>(newX, N, w.idx_full, K, labels_out, w.used);
    PF_LAUNCH_CHECK();
    nms_nlabels_kernel<<<B, 256, 0, st>>>(w.used, K, w.idx_full, N, Kcap, idx_out, n_labels_out);
    PF_LAUNCH_CHECK();
    return 0;
}

}  // namespace

extern "C" size_t prifit_nms_workspace_bytes(int B, int N, int d) {
    // votes, flags, used, nearest, best, rowsel, idx_full, nrows | split fp16 rows (hi + lo) for the tensor-core Gram
    return (7 * (size_t)B * N + B) * sizeof(int32_t) + 256 + (d == 128 ? prifit_tc_gram_split_bytes(B, N) : 0);
}

extern "C" int prifit_nms_fwd(const float* newX, const float* bw, int B, int N, int d, int Kcap,
                              int32_t* idx_out, int32_t* K_out, int32_t* labels_out, int32_t* n_labels_out,
                              void* ws, size_t ws_bytes, void* stream) {
    PF_CHECK_ARG(newX && bw && idx_out && K_out && ws, PRIFIT_E_BADARG, "null pointer");
    PF_CHECK_ARG((labels_out == nullptr) == (n_labels_out == nullptr), PRIFIT_E_BADARG, "labels_out and n_labels_out go together");
    PF_CHECK_ARG(B > 0 && N > 0, PRIFIT_E_BADARG, "B, N > 0 required");
    PF_CHECK_ARG(Kcap > 0 && Kcap % 4 == 0 && Kcap <= 64, PRIFIT_E_SHAPE, "Kcap must be a multiple of 4, <= 64");
    PF_CHECK_ARG(ws_bytes >= prifit_nms_workspace_bytes(B, N, d), PRIFIT_E_WS, "workspace too small");
    switch (d) {
        case 64: return launch_nms<64>(newX, bw, B, N, Kcap, idx_out, K_out, labels_out, n_labels_out, ws, pf_stream(stream));
        case 128: return launch_nms<128>(newX, bw, B, N, Kcap, idx_out, K_out, labels_out, n_labels_out, ws, pf_stream(stream));
        case 256: return launch_nms<256>(newX, bw, B, N, Kcap, idx_out, K_out, labels_out, n_labels_out, ws, pf_stream(stream));
        default: prifit_set_error("prifit_nms_fwd: d must be 64, 128 or 256 (got %d)", d); return PRIFIT_E_SHAPE;
    }
}

extern "C" int prifit_nms_labels(const float* newX, int B, int N, int d, int Kcap, const int32_t* K,
                                 int32_t* idx_out, int32_t* labels_out, int32_t* n_labels_out,
                                 void* ws, size_t ws_bytes, void* stream) {
    PF_CHECK_ARG(newX && K && idx_out && labels_out && n_labels_out && ws, PRIFIT_E_BADARG, "null pointer");
    PF_CHECK_ARG(B > 0 && N > 0, PRIFIT_E_BADARG, "B, N > 0 required");
    PF_CHECK_ARG(Kcap > 0 && Kcap % 4 == 0 && Kcap <= 64, PRIFIT_E_SHAPE, "Kcap must be a multiple of 4, <= 64");
    PF_CHECK_ARG(ws_bytes >= prifit_nms_workspace_bytes(B, N, d), PRIFIT_E_WS, "workspace too small");
    switch (d) {
        case 64: return launch_nms_labels<64>(newX, B, N, Kcap, K, idx_out, labels_out, n_labels_out, ws, pf_stream(stream));
        case 128: return launch_nms_labels<128>(newX, B, N, Kcap, K, idx_out, labels_out, n_labels_out, ws, pf_stream(stream));
        case 256: return launch_nms_labels<256>(newX, B, N, Kcap, K, idx_out, labels_out, n_labels_out, ws, pf_stream(stream));
        default: prifit_set_error("prifit_nms_labels: d must be 64, 128 or 256 (got %d)", d); return PRIFIT_E_SHAPE;
    }
}
