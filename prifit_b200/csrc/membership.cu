// k4 -- soft membership of every point in every cluster, forward and backward.
// reference src/mean_shift.py:230-247 (membership):
//     sim = C X^T / bw^2 ; sim -= sim.max().detach() ; e = guard_exp(sim) ; mem = e / sum_k e
// Output layout is the reference's [K, N] (cluster-major), padded to [B, Kcap, N].
#include <stdlib.h>
#include "rowgemm.cuh"

namespace {

// ---- forward pass 1: raw similarities into W, per-CTA maxima into blockmax[B][ntile*nrg]
template <int D>
__global__ void __launch_bounds__(RG_THREADS) membership_sim_kernel(
    const float* __restrict__ C, const float* __restrict__ X, const float* __restrict__ bw,
    const int32_t* __restrict__ K, int N, int Kcap, float* __restrict__ W, float* __restrict__ blockmax) {
    constexpr int LD = D + 4;
    extern __shared__ __align__(16) float smem[];
    float* ys = smem;
    float* xs = ys + RG_ROWS * LD;
    __shared__ float red[8];
    const int b = blockIdx.z, k0 = blockIdx.y * RG_ROWS, j0 = blockIdx.x * RG_KEYS;
    const int Kb = min(K[b], Kcap);
    const int nrows = max(0, min(RG_ROWS, Kb - k0));
    const int tid = threadIdx.x, ty = tid >> 5, tx = tid & 31;
    float* bm = blockmax + ((size_t)b * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    if (nrows == 0) { if (tid == 0) *bm = -INFINITY; return; }
    const float bwv = bw[b];
    const float b2 = bwv * bwv;
    rg_load_rows<D>(ys, RG_ROWS, C + ((size_t)b * Kcap + k0) * D, [&](int r) -> long long { return r < nrows ? (long long)r : -1; });
    rg_load_rows<D>(xs, RG_KEYS, X + (size_t)b * N * D, [&](int r) -> long long { return j0 + r < N ? (long long)(j0 + r) : -1; });
    __syncthreads();
    float acc[4][4];
    rg_dot_32x128<D>(ys, xs, acc, nrows);
    float mx = -INFINITY;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int r = ty + 8 * a;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int j = j0 + tx + 32 * c;
            if (r < nrows && j < N) {
                const float s = acc[a][c] / b2;
                W[((size_t)b * Kcap + k0 + r) * N + j] = s;
                mx = fmaxf(mx, s);
            }
        }
    }
    mx = warp_max(mx);
    if (tx == 0) red[ty] = mx;
    __syncthreads();
    if (tid == 0) {
        float m = red[0];
        for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]);
        *bm = m;
    }
}

// ---- forward pass 2: W = exp(clamp(sim - max)) / sum_k, one thread per point
__global__ void __launch_bounds__(256) membership_softmax_kernel(
    const int32_t* __restrict__ K, int N, int Kcap, int nbm, const float* __restrict__ blockmax,
    float* __restrict__ W, float* __restrict__ smax_out) {
    const int b = blockIdx.y;
    const int Kb = min(K[b], Kcap);
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    float smax = -INFINITY;
    for (int q = 0; q < nbm; ++q) smax = fmaxf(smax, blockmax[(size_t)b * nbm + q]);
    if (blockIdx.x == 0 && threadIdx.x == 0) smax_out[b] = smax;
    if (j >= N) return;
    float* Wb = W + (size_t)b * Kcap * N + j;
    float sum = 0.f;
    for (int k = 0; k < Kb; ++k) sum += guard_expf(Wb[(size_t)k * N] - smax);
    for (int k = 0; k < Kb; ++k) Wb[(size_t)k * N] = guard_expf(Wb[(size_t)k * N] - smax) / sum;
    for (int k = Kb; k < Kcap; ++k) Wb[(size_t)k * N] = 0.f;
}

// ---- backward
//   t_j   = sum_k gW_kj w_kj
//   dsim_kj = w_kj (gW_kj - t_j) [lo <= sim_kj - max <= hi] / bw^2        (sim recomputed)
//   gC_k  = sum_j dsim_kj x_j         (partial sums over MB_SPLIT key slices, reduced in fixed order by a second kernel)
//   gX_j += sum_k dsim_kj c_k
// The keys of a shape are cut into MB_SPLIT slices (a number that depends on N only, so a shape's result does not depend on
// the batch it is launched in); with 102 KB of shared memory two CTAs share an SM.  16 slices (one 128-key tile per CTA at
// N = 2048): the 8 shapes of a graph branch are 128 CTAs, 24.4 us (8 slices: 64 CTAs, 37.8 us); a 24-shape batch 50.6 (54.8) us.
// (The first version used a 4-CTA cluster per shape and a DSMEM reduction: 96 CTAs, 82 us.)
constexpr int MB_SPLIT = 16;

template <int D>
__global__ void __launch_bounds__(RG_THREADS, 2) membership_bwd_kernel(
    const float* __restrict__ C, const float* __restrict__ X, const float* __restrict__ bw,
    const int32_t* __restrict__ K, const float* __restrict__ W, const float* __restrict__ smax,
    const float* __restrict__ gW, int N, int Kcap, int nsplit, float* __restrict__ part, float* __restrict__ gX) {
    constexpr int LD = D + 4;
    constexpr int NH = (D + 127) / 128;
    const int rank = blockIdx.x, b = blockIdx.y;
    const int Kb = min(K[b], Kcap);
    const int tid = threadIdx.x, ty = tid >> 5, tx = tid & 31;
    float* part_b = part + ((size_t)b * nsplit + rank) * Kcap * D;
    if (Kb <= 0) return;     // the reduce kernel writes zeros

    extern __shared__ __align__(16) float smem[];
    float* ys = smem;
    float* xs = ys + RG_ROWS * LD;
    float* ps = xs + RG_KEYS * LD;
    float* tj = ps + RG_ROWS * RG_LDP;     // [128]

    const float* Xb = X + (size_t)b * N * D;
    float* gXb = gX + (size_t)b * N * D;
    const float* Wb = W + (size_t)b * Kcap * N;
    const float* gWb = gW + (size_t)b * Kcap * N;
    const float bwv = bw[b];
    const float b2 = bwv * bwv;
    const float mxv = smax[b];
    const int ntile = (N + RG_KEYS - 1) / RG_KEYS;
    const int tpr = (ntile + nsplit - 1) / nsplit;                      // key tiles per slice
    const int jbeg = min(N, rank * tpr * RG_KEYS), jend = min(N, (rank + 1) * tpr * RG_KEYS);

    for (int k0 = 0; k0 < Kb; k0 += RG_ROWS) {
        const int nrows = min(RG_ROWS, Kb - k0);
        __syncthreads();
        rg_load_rows<D>(ys, RG_ROWS, C + ((size_t)b * Kcap + k0) * D, [&](int r) -> long long { return r < nrows ? (long long)r : -1; });
        float o[4][4 * NH];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 4 * NH; ++c) o[a][c] = 0.f;
        for (int j0 = jbeg; j0 < jend; j0 += RG_KEYS) {
            __syncthreads();
            rg_load_rows<D>(xs, RG_KEYS, Xb, [&](int r) -> long long { return j0 + r < jend ? (long long)(j0 + r) : -1; });
            if (tid < RG_KEYS) {
                const int j = j0 + tid;
                float t = 0.f;
                if (j < jend)
                    for (int k = 0; k < Kb; ++k) t = fmaf(gWb[(size_t)k * N + j], Wb[(size_t)k * N + j], t);
                tj[tid] = t;
            }
            __syncthreads();
            float acc[4][4];
            rg_dot_32x128<D>(ys, xs, acc, nrows);
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const int r = ty + 8 * a;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int jl = tx + 32 * c, j = j0 + jl;
                    float ds = 0.f;
                    if (r < nrows && j < jend) {
                        const float sh = acc[a][c] / b2 - mxv;
                        if (sh >= PRIFIT_LO && sh <= PRIFIT_HI) {
                            const float w = Wb[(size_t)(k0 + r) * N + j];
                            ds = (w * (gWb[(size_t)(k0 + r) * N + j] - tj[jl])) / b2;
                        }
                    }
                    ps[r * RG_LDP + jl] = ds;
                }
            }
            __syncthreads();
            rg_accum_rows<D>(ps, xs, o, nrows);
            rg_accum_keys<D, false>(ps, ys, nullptr, nullptr, gXb + (size_t)j0 * D, min(RG_KEYS, jend - j0), nrows);
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int h = 0; h < NH; ++h)
                if (4 * tx + 128 * h < D && ty + 8 * a < nrows)
                    *reinterpret_cast<float4*>(part_b + (size_t)(k0 + ty + 8 * a) * D + 4 * tx + 128 * h) =
                        make_float4(o[a][4 * h], o[a][4 * h + 1], o[a][4 * h + 2], o[a][4 * h + 3]);
    }
}

// gC[b, k, :] = sum over the key slices in slice order (rows k >= K[b] are zero)
__global__ void membership_gc_reduce_kernel(const float* __restrict__ part, const int32_t* __restrict__ K, int Kcap, int D, int nsplit,
                                            float* __restrict__ gC) {
    const int b = blockIdx.y, k = blockIdx.x;
    const int Kb = min(K[b], Kcap);
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
        float v = 0.f;
        if (k < Kb)
            for (int q = 0; q < nsplit; ++q) v += part[(((size_t)b * nsplit + q) * Kcap + k) * D + c];
        gC[((size_t)b * Kcap + k) * D + c] = v;
    }
}

template <int D>
int launch_fwd(const float* C, const float* X, const float* bw, const int32_t* K, int B, int N, int Kcap,
               float* W, float* smax, float* blockmax, cudaStream_t st) {
    const size_t smem = (size_t)(RG_ROWS + RG_KEYS) * (D + 4) * sizeof(float);
    PF_CUDA(cudaFuncSetAttribute(membership_sim_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((N + RG_KEYS - 1) / RG_KEYS, (Kcap + RG_ROWS - 1) / RG_ROWS, B);
    membership_sim_kernel<D><<<grid, RG_THREADS, smem, st>>>(C, X, bw, K, N, Kcap, W, blockmax);
    PF_LAUNCH_CHECK();
    dim3 g2((N + 255) / 256, B);
    membership_softmax_kernel<<<g2, 256, 0, st>>>(K, N, Kcap, (int)(grid.x * grid.y), blockmax, W, smax);
    PF_LAUNCH_CHECK();
    return 0;
}

template <int D>
int launch_bwd(const float* C, const float* X, const float* bw, const int32_t* K, const float* W, const float* smax,
               const float* gW, int B, int N, int Kcap, float* gC, float* gX, float* part, cudaStream_t st) {
    const size_t smem = ((size_t)(RG_ROWS + RG_KEYS) * (D + 4) + (size_t)RG_ROWS * RG_LDP + RG_KEYS) * sizeof(float);
    PF_CUDA(cudaFuncSetAttribute(membership_bwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int ntile = (N + RG_KEYS - 1) / RG_KEYS;
    const int nsplit = ntile < MB_SPLIT ? ntile : MB_SPLIT;          // depends on N only
    membership_bwd_kernel<D><<<dim3(nsplit, B), RG_THREADS, smem, st>>>(C, X, bw, K, W, smax, gW, N, Kcap, nsplit, part, gX);
    PF_LAUNCH_CHECK();
    membership_gc_reduce_kernel<<<dim3(Kcap, B), 128, 0, st>>>(part, K, Kcap, D, nsplit, gC);
    PF_LAUNCH_CHECK();
    return 0;
}

}  // namespace

extern "C" size_t prifit_membership_workspace_bytes(int B, int N, int Kcap) {
    const size_t nbm = (size_t)((N + RG_KEYS - 1) / RG_KEYS) * ((Kcap + RG_ROWS - 1) / RG_ROWS);
    return (size_t)B * nbm * sizeof(float);
}

extern "C" int prifit_membership_fwd(const float* C, const float* X, const float* bw, const int32_t* K,
                                     int B, int N, int d, int Kcap, float* W_out, float* smax_out,
                                     void* ws, size_t ws_bytes, void* stream) {
    PF_CHECK_ARG(C && X && bw && K && W_out && smax_out && ws, PRIFIT_E_BADARG, "null pointer");
    PF_CHECK_ARG(B > 0 && N > 0, PRIFIT_E_BADARG, "B, N > 0 required");
    PF_CHECK_ARG(Kcap > 0 && Kcap % 4 == 0 && Kcap <= 64, PRIFIT_E_SHAPE, "Kcap must be a multiple of 4, <= 64");
    PF_CHECK_ARG(ws_bytes >= prifit_membership_workspace_bytes(B, N, Kcap), PRIFIT_E_WS, "workspace too small");
    float* bm = static_cast<float*>(ws);
    switch (d) {
        case 64: return launch_fwd<64>(C, X, bw, K, B, N, Kcap, W_out, smax_out, bm, pf_stream(stream));
        case 128: return launch_fwd<128>(C, X, bw, K, B, N, Kcap, W_out, smax_out, bm, pf_stream(stream));
        case 256: return launch_fwd<256>(C, X, bw, K, B, N, Kcap, W_out, smax_out, bm, pf_stream(stream));
        default: prifit_set_error("prifit_membership_fwd: d must be 64, 128 or 256 (got %d)", d); return PRIFIT_E_SHAPE;
    }
}

extern "C" size_t prifit_membership_bwd_workspace_bytes(int B, int Kcap, int d) {
    return (size_t)B * MB_SPLIT * Kcap * d * sizeof(float);
}

extern "C" int prifit_membership_bwd(const float* C, const float* X, const float* bw, const int32_t* K,
                                     const float* W, const float* smax, const float* gW,
                                     int B, int N, int d, int Kcap, float* gC_out, float* gX_inout,
                                     void* ws, size_t ws_bytes, void* stream) {
    PF_CHECK_ARG(C && X && bw && K && W && smax && gW && gC_out && gX_inout && ws, PRIFIT_E_BADARG, "null pointer");
    PF_CHECK_ARG(B > 0 && N > 0, PRIFIT_E_BADARG, "B, N > 0 required");
    PF_CHECK_ARG(Kcap > 0 && Kcap % 4 == 0 && Kcap <= 64, PRIFIT_E_SHAPE, "Kcap must be a multiple of 4, <= 64");
    PF_CHECK_ARG(ws_bytes >= prifit_membership_bwd_workspace_bytes(B, Kcap, d), PRIFIT_E_WS, "workspace too small");
    float* part = static_cast<float*>(ws);
    switch (d) {
        case 64: return launch_bwd<64>(C, X, bw, K, W, smax, gW, B, N, Kcap, gC_out, gX_inout, part, pf_stream(stream));
        case 128: return launch_bwd<128>(C, X, bw, K, W, smax, gW, B, N, Kcap, gC_out, gX_inout, part, pf_stream(stream));
        case 256: return launch_bwd<256>(C, X, bw, K, W, smax, gW, B, N, Kcap, gC_out, gX_inout, part, pf_stream(stream));
        default: prifit_set_error("prifit_membership_bwd: d must be 64, 128 or 256 (got %d)", d); return PRIFIT_E_SHAPE;
    }
}
