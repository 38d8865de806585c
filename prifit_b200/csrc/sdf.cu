// k8 -- SDF half of the fitting loss, forward + backward.
// reference convex_loss.py:313-343 (compute_sdf_ellipsoid[s][_batch]) and the reduction in
// src/utils.py:407-411 (abs, min over ellipsoids, square, mean, /2):
//   z = V^T (p - c) ; k0 = |z / (s + 1e-6)| ; k1 = |z / (s^2 + 1e-6)| ; sdf = k0 (k0 - 1) / (k1 + 1e-6)
//   loss_b = 0.5 * mean_j (min_k |sdf_kj|)^2
#include "common.cuh"
#include "sdf_math.cuh"

namespace {

__global__ void __launch_bounds__(SDF_THREADS) sdf_fwd_kernel(
    const float* __restrict__ Q, const float* __restrict__ s, const float* __restrict__ V, const float* __restrict__ c,
    const uint8_t* __restrict__ valid, const int32_t* __restrict__ K, int M, int Kcap,
    int32_t* __restrict__ argmin_out, float* __restrict__ sdf_out, float* __restrict__ partial) {
    __shared__ float prm[SDF_MAXK][15];
    __shared__ int list[SDF_MAXK];
    __shared__ int nlist;
    __shared__ float red[32];
    const int b = blockIdx.y, tid = threadIdx.x;
    const int Kb = min(K[b], Kcap);
    if (tid == 0) {
        int n = 0;
        for (int k = 0; k < Kb; ++k)
            if (valid[(size_t)b * Kcap + k]) list[n++] = k;
        nlist = n;
    }
    for (int e = tid; e < Kb * 15; e += SDF_THREADS) {
        const int k = e / 15, f = e - 15 * k;
        const size_t bk = (size_t)b * Kcap + k;
        prm[k][f] = f < 3 ? s[bk * 3 + f] : (f < 12 ? V[bk * 9 + f - 3] : c[bk * 3 + f - 12]);
    }
    __syncthreads();
    const int j = blockIdx.x * SDF_THREADS + tid;
    float sq[1] = {0.f};
    if (j < M) {
        const float* q = Q + ((size_t)b * M + j) * 3;
        const float px = q[0], py = q[1], pz = q[2];
        float best = INFINITY, bests = 0.f;
        int bi = -1;
        for (int e = 0; e < nlist; ++e) {
            const float v = sdf_eval(prm[list[e]], px, py, pz);
            if (bi < 0 || fabsf(v) < best) { best = fabsf(v); bests = v; bi = list[e]; }
        }
        argmin_out[(size_t)b * M + j] = bi;
        sdf_out[(size_t)b * M + j] = bests;
        if (bi >= 0) sq[0] = best * best;
    }
    block_sum<1>(sq, red);
    if (tid == 0) partial[(size_t)b * gridDim.x + blockIdx.x] = sq[0];
}

__global__ void sdf_finalize_kernel(const float* __restrict__ partial, int nblk, int M, float* __restrict__ loss_out) {
    const int b = blockIdx.x;
    if (threadIdx.x == 0) {
        float acc = 0.f;
        for (int q = 0; q < nblk; ++q) acc += partial[(size_t)b * nblk + q];
        loss_out[b] = (acc / (float)M) / 2.0f;
    }
}

// one CTA per (shape, ellipsoid): deterministic reduction over the points that chose it
__global__ void __launch_bounds__(SDF_THREADS) sdf_bwd_params_kernel(
    const float* __restrict__ Q, const float* __restrict__ s, const float* __restrict__ V, const float* __restrict__ c,
    const uint8_t* __restrict__ valid, const int32_t* __restrict__ K, const int32_t* __restrict__ argmin,
    const float* __restrict__ gloss, int M, int Kcap,
    float* __restrict__ gs, float* __restrict__ gV, float* __restrict__ gc) {
    __shared__ float prm[15];
    __shared__ float red[15 * 32];
    const int k = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const size_t bk = (size_t)b * Kcap + k;
    if (k >= min(K[b], Kcap) || !valid[bk]) {
        if (tid < 3) { gs[bk * 3 + tid] = 0.f; gc[bk * 3 + tid] = 0.f; }
        if (tid < 9) gV[bk * 9 + tid] = 0.f;
        return;
    }
    if (tid < 15) prm[tid] = tid < 3 ? s[bk * 3 + tid] : (tid < 12 ? V[bk * 9 + tid - 3] : c[bk * 3 + tid - 12]);
    __syncthreads();
    const float scale = gloss[b] / (float)M;
    float acc[15];
#pragma unroll
    for (int i = 0; i < 15; ++i) acc[i] = 0.f;
    // 8 points per thread and round: all arg-min loads of a round are issued together and only the points that chose this
    // ellipsoid (1 in K) go through the gradient arithmetic -- in ascending order, so the sums are those of the plain loop
    // (which paid an L2 round trip and, in nearly every warp, a divergent pass through the arithmetic per point: 21 us)
    constexpr int U = 8;
    const int32_t* am = argmin + (size_t)b * M;
    for (int j0 = 0; j0 < M; j0 += U * SDF_THREADS) {
        uint32_t hit = 0;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int j = j0 + u * SDF_THREADS + tid;
            if (j < M && am[j] == k) hit |= 1u << u;
        }
        while (hit) {
            const int u = __ffs(hit) - 1;
            hit &= hit - 1;
            const float* q = Q + ((size_t)b * M + j0 + u * SDF_THREADS + tid) * 3;
            const float v = sdf_eval(prm, q[0], q[1], q[2]);
            float dV[9], dsv[3], dpt[3];
            sdf_point_grad(prm, q[0], q[1], q[2], scale * v, dV, dsv, dpt);
#pragma unroll
            for (int i = 0; i < 3; ++i) { acc[i] += dsv[i]; acc[12 + i] -= dpt[i]; }
#pragma unroll
            for (int i = 0; i < 9; ++i) acc[3 + i] += dV[i];
        }
    }
    block_sum<15>(acc, red);
    if (tid < 3) { gs[bk * 3 + tid] = acc[tid]; gc[bk * 3 + tid] = acc[12 + tid]; }
    if (tid < 9) gV[bk * 9 + tid] = acc[3 + tid];
}

__global__ void __launch_bounds__(SDF_THREADS) sdf_bwd_points_kernel(
    const float* __restrict__ Q, const float* __restrict__ s, const float* __restrict__ V, const float* __restrict__ c,
    const int32_t* __restrict__ argmin, const float* __restrict__ gloss, int M, int Kcap, float* __restrict__ gQ) {
    const int b = blockIdx.y, j = blockIdx.x * SDF_THREADS + threadIdx.x;
    if (j >= M) return;
    float* out = gQ + ((size_t)b * M + j) * 3;
    const int k = argmin[(size_t)b * M + j];
    if (k < 0) { out[0] = out[1] = out[2] = 0.f; return; }
    const size_t bk = (size_t)b * Kcap + k;
    float prm[15];
    for (int f = 0; f < 15; ++f) prm[f] = f < 3 ? s[bk * 3 + f] : (f < 12 ? V[bk * 9 + f - 3] : c[bk * 3 + f - 12]);
    const float* q = Q + ((size_t)b * M + j) * 3;
    const float v = sdf_eval(prm, q[0], q[1], q[2]);
    float dV[9], dsv[3], dpt[3];
    sdf_point_grad(prm, q[0], q[1], q[2], gloss[b] / (float)M * v, dV, dsv, dpt);
    out[0] = dpt[0]; out[1] = dpt[1]; out[2] = dpt[2];
}

// ---- batch mean over the shapes that kept at least one ellipsoid (src/utils.py:418,425)
__global__ void __launch_bounds__(256) masked_mean_fwd_kernel(const float* __restrict__ loss_b, const uint8_t* __restrict__ valid,
                                                              int B, int Kcap, float* __restrict__ has_out, float* __restrict__ stats) {
    __shared__ float red[2 * 32];
    float acc[2] = {0.f, 0.f};
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        int any = 0;
        for (int k = 0; k < Kcap; ++k) any |= valid[(size_t)b * Kcap + k];
        const float h = any ? 1.0f : 0.0f;
        has_out[b] = h;
        acc[0] += any ? loss_b[b] : 0.f;
        acc[1] += h;
    }
    block_sum<2>(acc, red);
    if (threadIdx.x == 0) {
        stats[0] = acc[0];
        stats[1] = acc[1];
        stats[2] = acc[0] / fmaxf(acc[1], 1.0f);
    }
}

__global__ void masked_mean_bwd_kernel(const float* __restrict__ g_sum, const float* __restrict__ g_mean,
                                       const float* __restrict__ has, const float* __restrict__ stats, int B,
                                       float* __restrict__ gloss) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float g = (g_sum ? g_sum[0] : 0.f) + (g_mean ? g_mean[0] / fmaxf(stats[1], 1.0f) : 0.f);
    gloss[b] = has[b] * g;
}

}  // namespace

extern "C" size_t prifit_sdf_workspace_bytes(int B, int M) {
    return (size_t)B * ((M + SDF_THREADS - 1) / SDF_THREADS) * sizeof(float);
}

extern "C" int prifit_sdf_loss_fwd(const float* Q, const float* s, const float* V, const float* c, const uint8_t* valid,
                                   const int32_t* K, int B, int M, int Kcap, float* loss_out, int32_t* argmin_out,
                                   float* sdf_out, void* ws, size_t ws_bytes, void* stream) {
    PF_CHECK_ARG(Q && s && V && c && valid && K && loss_out && argmin_out && sdf_out && ws, PRIFIT_E_BADARG, "null pointer");
    PF_CHECK_ARG(B > 0 && M > 0, PRIFIT_E_BADARG, "B, M > 0 required");
    PF_CHECK_ARG(Kcap > 0 && Kcap <= SDF_MAXK, PRIFIT_E_SHAPE, "Kcap must be in 1..64");
    PF_CHECK_ARG(ws_bytes >= prifit_sdf_workspace_bytes(B, M), PRIFIT_E_WS, "workspace too small");
    const int nblk = (M + SDF_THREADS - 1) / SDF_THREADS;
    float* partial = static_cast<float*>(ws);
    sdf_fwd_kernel<<<dim3(nblk, B), SDF_THREADS, 0, pf_stream(stream)>>>(Q, s, V, c, valid, K, M, Kcap, argmin_out, sdf_out, partial);
    PF_LAUNCH_CHECK();
    sdf_finalize_kernel<<<B, 32, 0, pf_stream(stream)>>>(partial, nblk, M, loss_out);
    PF_LAUNCH_CHECK();
    return 0;
}

extern "C" int prifit_sdf_loss_bwd(const float* Q, const float* s, const float* V, const float* c, const uint8_t* valid,
                                   const int32_t* K, const int32_t* argmin, const float* gloss, int B, int M, int Kcap,
                                   float* gs_out, float* gV_out, float* gc_out, float* gQ_out, void* stream) {
    PF_CHECK_ARG(Q && s && V && c && valid && K && argmin && gloss && gs_out && gV_out && gc_out, PRIFIT_E_BADARG, "null pointer");
    PF_CHECK_ARG(B > 0 && M > 0, PRIFIT_E_BADARG, "B, M > 0 required");
    PF_CHECK_ARG(Kcap > 0 && Kcap <= SDF_MAXK, PRIFIT_E_SHAPE, "Kcap must be in 1..64");
    sdf_bwd_params_kernel<<<dim3(Kcap, B), SDF_THREADS, 0, pf_stream(stream)>>>(Q, s, V, c, valid, K, argmin, gloss, M, Kcap, gs_out, gV_out, gc_out);
    PF_LAUNCH_CHECK();
    if (gQ_out) {
        sdf_bwd_points_kernel<<<dim3((M + SDF_THREADS - 1) / SDF_THREADS, B), SDF_THREADS, 0, pf_stream(stream)>>>(Q, s, V, c, argmin, gloss, M, Kcap, gQ_out);
        PF_LAUNCH_CHECK();
    }
    return 0;
}

extern "C" int prifit_masked_mean_fwd(const float* loss_b, const uint8_t* valid, int B, int Kcap, float* has_out,
                                      float* stats_out, void* stream) {
    PF_CHECK_ARG(loss_b && valid && has_out && stats_out, PRIFIT_E_BADARG, "null pointer");
    PF_CHECK_ARG(B > 0 && Kcap > 0, PRIFIT_E_BADARG, "B, Kcap > 0 required");
    masked_mean_fwd_kernel<<<1, 256, 0, pf_stream(stream)>>>(loss_b, valid, B, Kcap, has_out, stats_out);
    PF_LAUNCH_CHECK();
    return 0;
}

extern "C" int prifit_masked_mean_bwd(const float* g_sum, const float* g_mean, const float* has, const float* stats, int B,
                                      float* gloss_out, void* stream) {
    PF_CHECK_ARG(has && stats && gloss_out, PRIFIT_E_BADARG, "null pointer");
    PF_CHECK_ARG(B > 0, PRIFIT_E_BADARG, "B > 0 required");
    masked_mean_bwd_kernel<<<(B + 127) / 128, 128, 0, pf_stream(stream)>>>(g_sum, g_mean, has, stats, B, gloss_out);
    PF_LAUNCH_CHECK();
    return 0;
}
