// f3 -- the entropy regulariser of convex_loss.py:209-225 (used with --include_entropy_loss, convex_loss.py:59-62):
//     l_b = sum_ij (1 + <x_i, x_j>)^2 / n^2     over a sub-sample of n points of shape b,     loss = relu(mean_b l_b - 1.8)
// The reference materialises the n x n matrix.  Expanding the square,
//     sum_ij (1 + s_ij)^2 = n^2 + 2 |m|^2 + |M|_F^2,     m = sum_i x_i  (d),   M = sum_i x_i x_i^T  (d x d),
// so the forward is one pass of second moments (n d^2 instead of n^2 d MACs, nothing of size n x n) and
//     d l_b / d x_i = (4 / n^2) (m + M x_i).
// Kernels: moments (rows split over ENT_SPLIT CTAs per shape, 8x8 register blocks of M), finalize (fixed-order sum of the
// partials, l_b), backward (32 sample rows per CTA, M in shared memory).  relu / mean / margin stay with the caller.
#include "common.cuh"

namespace {

constexpr int ENT_SPLIT = 8;
constexpr int ENT_THREADS = 256;
constexpr int ENT_ROWS = 32;            // sample rows staged per step

template <int D>
__global__ void __launch_bounds__(ENT_THREADS) entropy_moments_kernel(
    const float* __restrict__ X, const int32_t* __restrict__ idx, int N, int n, float* __restrict__ partial) {
    constexpr int BLK = D / 16;                       // each thread owns BLK x BLK entries of M: rows ty + 16 i, columns tx + 16 j
    __shared__ __align__(16) float xs[ENT_ROWS][D];
    const int b = blockIdx.y, part = blockIdx.x, tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;
    const int per = (n + ENT_SPLIT - 1) / ENT_SPLIT;
    const int i0 = part * per, i1 = min(n, i0 + per);
    float acc[BLK][BLK], msum = 0.f;
#pragma unroll
    for (int i = 0; i < BLK; ++i)
#pragma unroll
        for (int j = 0; j < BLK; ++j) acc[i][j] = 0.f;
    for (int r0 = i0; r0 < i1; r0 += ENT_ROWS) {
        const int nr = min(ENT_ROWS, i1 - r0);
        __syncthreads();
        for (int e = tid; e < nr * (D / 4); e += ENT_THREADS) {
            const int r = e / (D / 4), c = e - r * (D / 4);
            const int src = idx ? idx[r0 + r] : r0 + r;
            reinterpret_cast<float4*>(&xs[r][0])[c] = reinterpret_cast<const float4*>(X + ((size_t)b * N + src) * D)[c];
        }
        __syncthreads();
        for (int r = 0; r < nr; ++r) {
            float xa[BLK], xb[BLK];
#pragma unroll
            for (int i = 0; i < BLK; ++i) { xa[i] = xs[r][ty + 16 * i]; xb[i] = xs[r][tx + 16 * i]; }   // interleaved: conflict-free
#pragma unroll
            for (int i = 0; i < BLK; ++i)
#pragma unroll
                for (int j = 0; j < BLK; ++j) acc[i][j] = fmaf(xa[i], xb[j], acc[i][j]);
            if (tid < D) msum += xs[r][tid];
        }
    }
    float* out = partial + ((size_t)b * ENT_SPLIT + part) * (D * D + D);
#pragma unroll
    for (int i = 0; i < BLK; ++i)
#pragma unroll
        for (int j = 0; j < BLK; ++j) out[(ty + 16 * i) * D + tx + 16 * j] = acc[i][j];
    if (tid < D) out[D * D + tid] = msum;
}

__global__ void __launch_bounds__(ENT_THREADS) entropy_finalize_kernel(const float* __restrict__ partial, int D, int n,
                                                                       float* __restrict__ moments, float* __restrict__ loss_b) {
    __shared__ float red[2 * 32];
    const int b = blockIdx.x, tid = threadIdx.x, W = D * D + D;
    float acc[2] = {0.f, 0.f};                       // |M|_F^2, |m|^2
    for (int e = tid; e < W; e += ENT_THREADS) {
        float v = 0.f;
        for (int p = 0; p < ENT_SPLIT; ++p) v += partial[((size_t)b * ENT_SPLIT + p) * W + e];
        moments[(size_t)b * W + e] = v;
        if (e < D * D) acc[0] = fmaf(v, v, acc[0]); else acc[1] = fmaf(v, v, acc[1]);
    }
    block_sum<2>(acc, red);
    if (tid == 0) {
        const float n2 = (float)n * (float)n;
        loss_b[b] = (n2 + 2.0f * acc[1] + acc[0]) / n2;
    }
}

// gX[b, idx[i], :] += gl[b] * (4 / n^2) * (m + M x_i)
template <int D>
__global__ void __launch_bounds__(ENT_THREADS) entropy_bwd_kernel(
    const float* __restrict__ X, const int32_t* __restrict__ idx, const float* __restrict__ gl, int N, int n,
    const float* __restrict__ moments, float* __restrict__ gX) {
    extern __shared__ __align__(16) float smem[];
    float* Ms = smem;                                 // [D][D + 1]
    float* xs = Ms + D * (D + 1);                     // [ENT_ROWS][D + 4]  (D (D + 1) is a multiple of 4: 16-byte aligned rows)
    const int b = blockIdx.y, r0 = blockIdx.x * ENT_ROWS, tid = threadIdx.x;
    const int nr = min(ENT_ROWS, n - r0);
    const float* Mb = moments + (size_t)b * (D * D + D);
    for (int e = tid; e < D * D; e += ENT_THREADS) Ms[(e / D) * (D + 1) + (e % D)] = Mb[e];
    for (int e = tid; e < nr * (D / 4); e += ENT_THREADS) {
        const int r = e / (D / 4), c = e - r * (D / 4);
        const int src = idx ? idx[r0 + r] : r0 + r;
        reinterpret_cast<float4*>(xs + r * (D + 4))[c] = reinterpret_cast<const float4*>(X + ((size_t)b * N + src) * D)[c];
    }
    __syncthreads();
    const float scale = gl[b] * 4.0f / ((float)n * (float)n);
    // thread -> (row r = tid / 8, columns c = tid % 8 + 8 q): consecutive lanes read consecutive rows of Ms (conflict-free)
    const int r = tid >> 3, cg = tid & 7;
    if (r < nr) {
        const int dst = idx ? idx[r0 + r] : r0 + r;
        float* g = gX + ((size_t)b * N + dst) * D;
        for (int c = cg; c < D; c += 8) {
            float v = Mb[D * D + c];
            for (int a = 0; a < D; ++a) v = fmaf(Ms[c * (D + 1) + a], xs[r * (D + 4) + a], v);
            g[c] += scale * v;
        }
    }
}

template <int D>
int entropy_fwd(const float* X, const int32_t* idx, int B, int N, int n, float* loss_b, float* partial, float* moments, cudaStream_t st) {
    entropy_moments_kernel<D><<<dim3(ENT_SPLIT, B), ENT_THREADS, 0, st>>>(X, idx, N, n, partial);
    PF_LAUNCH_CHECK();
    entropy_finalize_kernel<<<B, ENT_THREADS, 0, st>>>(partial, D, n, moments, loss_b);
    PF_LAUNCH_CHECK();
    return 0;
}

template <int D>
int entropy_bwd(const float* X, const int32_t* idx, const float* gl, int B, int N, int n, const float* moments, float* gX, cudaStream_t st) {
    const size_t smem = ((size_t)D * (D + 1) + (size_t)ENT_ROWS * (D + 4)) * sizeof(float);
    PF_CUDA(cudaFuncSetAttribute(entropy_bwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    entropy_bwd_kernel<D><<<dim3((n + ENT_ROWS - 1) / ENT_ROWS, B), ENT_THREADS, smem, st>>>(X, idx, gl, N, n, moments, gX);
    PF_LAUNCH_CHECK();
    return 0;
}

}  // namespace

extern "C" size_t prifit_entropy_workspace_bytes(int B, int d) {
    return (size_t)B * (ENT_SPLIT + 1) * ((size_t)d * d + d) * sizeof(float);
}

extern "C" int prifit_entropy_fwd(const float* X, const int32_t* idx, int B, int N, int d, int n, float* loss_b_out,
                                  void* ws, size_t ws_bytes, void* stream) {
    PF_CHECK_ARG(X && loss_b_out && ws, PRIFIT_E_BADARG, "null pointer");
    PF_CHECK_ARG(B > 0 && N > 0 && n > 0 && n <= N, PRIFIT_E_BADARG, "need B > 0 and 0 < n <= N");
    PF_CHECK_ARG(idx || n == N, PRIFIT_E_BADARG, "idx == NULL requires n == N");
    PF_CHECK_ARG(ws_bytes >= prifit_entropy_workspace_bytes(B, d), PRIFIT_E_WS, "workspace too small");
    float* partial = static_cast<float*>(ws);
    float* moments = partial + (size_t)B * ENT_SPLIT * ((size_t)d * d + d);
    switch (d) {
        case 64: return entropy_fwd<64>(X, idx, B, N, n, loss_b_out, partial, moments, pf_stream(stream));
        case 128: return entropy_fwd<128>(X, idx, B, N, n, loss_b_out, partial, moments, pf_stream(stream));
        default: prifit_set_error("prifit_entropy_fwd: d must be 64 or 128 (got %d)", d); return PRIFIT_E_SHAPE;
    }
}

extern "C" int prifit_entropy_bwd(const float* X, const int32_t* idx, const float* gloss_b, int B, int N, int d, int n,
                                  const void* ws, float* gX_inout, void* stream) {
    PF_CHECK_ARG(X && gloss_b && ws && gX_inout, PRIFIT_E_BADARG, "null pointer");
    PF_CHECK_ARG(B > 0 && N > 0 && n > 0 && n <= N, PRIFIT_E_BADARG, "need B > 0 and 0 < n <= N");
    PF_CHECK_ARG(idx || n == N, PRIFIT_E_BADARG, "idx == NULL requires n == N");
    const float* moments = static_cast<const float*>(ws) + (size_t)B * ENT_SPLIT * ((size_t)d * d + d);
    switch (d) {
        case 64: return entropy_bwd<64>(X, idx, gloss_b, B, N, n, moments, gX_inout, pf_stream(stream));
        case 128: return entropy_bwd<128>(X, idx, gloss_b, B, N, n, moments, gX_inout, pf_stream(stream));
        default: prifit_set_error("prifit_entropy_bwd: d must be 64 or 128 (got %d)", d); return PRIFIT_E_SHAPE;
    }
}
