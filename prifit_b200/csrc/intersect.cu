// Intersection penalty between the fitted ellipsoids of a shape, forward + backward.
// reference convex_loss.py:346-441 -- compute_intersection_loss_volume_3 (what convex_loss calls at :97, through
// torch_scatter.scatter_mean whose import is commented out at :17, so the shipped file raises NameError there) and
// compute_intersection_loss_volume_4.  Both work on g_kj = min(sdf_k(p_j), -1e-3) over the probe points p_j (the chamfer
// cloud minus U[0, 0.2) jitter, :97), i.e. only points INSIDE an ellipsoid (sdf < -1e-3) carry gradient:
//   version 3   k*_j = argmin_k g_kj ; loss_b = mean_j ( mean_{k != k*_j} g_kj )^2          (:377-410, scatter_mean over index 0)
//   version 4   loss_b = mean_j ( sum_k g_kj^2 - (min_k g_kj)^2 )                            (:413-441)
// Shapes with fewer than two ellipsoids are skipped; the batch value is the mean over the remaining shapes (0 if none).
// One thread per probe point in the forward; the backward is one CTA per (shape, ellipsoid) reducing over the points in
// a fixed order (deterministic), with the per-point quantities it needs (k*_j and mean / min) saved by the forward.
#include "sdf_math.cuh"

namespace {

constexpr float IX_CLAMP = -1e-3f;

__device__ __forceinline__ void load_params(float (*prm)[15], int* list, int* nlist, const float* s, const float* V, const float* c,
                                            const uint8_t* valid, int b, int Kb, int Kcap) {
    if (threadIdx.x == 0) {
        int n = 0;
        for (int k = 0; k < Kb; ++k)
            if (valid[(size_t)b * Kcap + k]) list[n++] = k;
        *nlist = n;
    }
    for (int e = threadIdx.x; e < Kb * 15; e += blockDim.x) {
        const int k = e / 15, f = e - 15 * k;
        const size_t bk = (size_t)b * Kcap + k;
        prm[k][f] = f < 3 ? s[bk * 3 + f] : (f < 12 ? V[bk * 9 + f - 3] : c[bk * 3 + f - 12]);
    }
    __syncthreads();
}

template <int VERSION>
__global__ void __launch_bounds__(SDF_THREADS) intersect_fwd_kernel(
    const float* __restrict__ Q, const float* __restrict__ s, const float* __restrict__ V, const float* __restrict__ c,
    const uint8_t* __restrict__ valid, const int32_t* __restrict__ K, int M, int Kcap,
    int32_t* __restrict__ kstar_out, float* __restrict__ aux_out, float* __restrict__ partial) {
    __shared__ float prm[SDF_MAXK][15];
    __shared__ int list[SDF_MAXK];
    __shared__ int nlist;
    __shared__ float red[32];
    const int b = blockIdx.y, tid = threadIdx.x;
    load_params(prm, list, &nlist, s, V, c, valid, b, min(K[b], Kcap), Kcap);
    const int n = nlist;
    const int j = blockIdx.x * SDF_THREADS + tid;
    float term[1] = {0.f};
    if (j < M && n >= 2) {
        const float* q = Q + ((size_t)b * M + j) * 3;
        const float px = q[0], py = q[1], pz = q[2];
        float best = INFINITY, sum = 0.f, sumsq = 0.f;
        int bi = -1;
        for (int e = 0; e < n; ++e) {                         // ascending k: torch.min returns the first index on ties
            const float g = fminf(sdf_eval(prm[list[e]], px, py, pz), IX_CLAMP);
            if (g < best) { best = g; bi = list[e]; }
            sum += g; sumsq += g * g;
        }
        float aux;
        if (VERSION == 3) {
            aux = (sum - best) / (float)(n - 1);              // mean over the other ellipsoids
            term[0] = aux * aux;
        } else {
            aux = best;
            term[0] = sumsq - best * best;
        }
        kstar_out[(size_t)b * M + j] = bi;
        aux_out[(size_t)b * M + j] = aux;
    } else if (j < M) {
        kstar_out[(size_t)b * M + j] = -1;
        aux_out[(size_t)b * M + j] = 0.f;
    }
    block_sum<1>(term, red);
    if (tid == 0) partial[(size_t)b * gridDim.x + blockIdx.x] = term[0];
}

__global__ void intersect_finalize_kernel(const float* __restrict__ partial, const uint8_t* __restrict__ valid,
                                          const int32_t* __restrict__ K, int nblk, int M, int Kcap,
                                          float* __restrict__ loss_out, float* __restrict__ counted_out) {
    const int b = blockIdx.x;
    if (threadIdx.x == 0) {
        int n = 0;
        for (int k = 0; k < min(K[b], Kcap); ++k) n += valid[(size_t)b * Kcap + k] ? 1 : 0;
        float acc = 0.f;
        for (int q = 0; q < nblk; ++q) acc += partial[(size_t)b * nblk + q];
        loss_out[b] = n >= 2 ? acc / (float)M : 0.f;
        counted_out[b] = n >= 2 ? 1.f : 0.f;
    }
}

// d loss_b / d g_kj (before the clamp mask):  version 3: 2 aux_j / (n - 1) for k != k*_j, 0 for k*_j
//                                             version 4: 2 g_kj           for k != k*_j, 0 for k*_j
// clamp_max passes the gradient where sdf <= -1e-3.
template <int VERSION>
__global__ void __launch_bounds__(SDF_THREADS) intersect_bwd_kernel(
    const float* __restrict__ Q, const float* __restrict__ s, const float* __restrict__ V, const float* __restrict__ c,
    const uint8_t* __restrict__ valid, const int32_t* __restrict__ K, const int32_t* __restrict__ kstar,
    const float* __restrict__ aux, const float* __restrict__ gloss, int M, int Kcap,
    float* __restrict__ gs, float* __restrict__ gV, float* __restrict__ gc) {
    __shared__ float prm[15];
    __shared__ float red[15 * 32];
    __shared__ int n_s;
    const int k = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const size_t bk = (size_t)b * Kcap + k;
    const int Kb = min(K[b], Kcap);
    if (tid == 0) {
        int n = 0;
        for (int q = 0; q < Kb; ++q) n += valid[(size_t)b * Kcap + q] ? 1 : 0;
        n_s = n;
    }
    __syncthreads();
    const int n = n_s;
    if (k >= Kb || !valid[bk] || n < 2) {
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
    for (int j = tid; j < M; j += SDF_THREADS) {
        if (kstar[(size_t)b * M + j] == k) continue;          // the closest ellipsoid is excluded (its terms cancel)
        const float* q = Q + ((size_t)b * M + j) * 3;
        const float v = sdf_eval(prm, q[0], q[1], q[2]);
        if (!(v <= IX_CLAMP)) continue;                        // clamped: constant, no gradient
        const float gg = VERSION == 3 ? 2.0f * aux[(size_t)b * M + j] / (float)(n - 1) : 2.0f * v;
        float dV[9], dsv[3], dpt[3];
        sdf_point_grad(prm, q[0], q[1], q[2], scale * gg, dV, dsv, dpt);
#pragma unroll
        for (int i = 0; i < 3; ++i) { acc[i] += dsv[i]; acc[12 + i] -= dpt[i]; }
#pragma unroll
        for (int i = 0; i < 9; ++i) acc[3 + i] += dV[i];
    }
    block_sum<15>(acc, red);
    if (tid < 3) { gs[bk * 3 + tid] = acc[tid]; gc[bk * 3 + tid] = acc[12 + tid]; }
    if (tid < 9) gV[bk * 9 + tid] = acc[3 + tid];
}

}  // namespace

extern "C" size_t prifit_intersect_workspace_bytes(int B, int M) {
    return (size_t)B * ((M + SDF_THREADS - 1) / SDF_THREADS) * sizeof(float);
}

extern "C" int prifit_intersect_fwd(const float* Q, const float* s, const float* V, const float* c, const uint8_t* valid,
                                    const int32_t* K, int B, int M, int Kcap, int version, float* loss_out, float* counted_out,
                                    int32_t* kstar_out, float* aux_out, void* ws, size_t ws_bytes, void* stream) {
    PF_CHECK_ARG(Q && s && V && c && valid && K && loss_out && counted_out && kstar_out && aux_out && ws, PRIFIT_E_BADARG, "null pointer");
    PF_CHECK_ARG(B > 0 && M > 0, PRIFIT_E_BADARG, "B, M > 0 required");
    PF_CHECK_ARG(Kcap > 0 && Kcap <= SDF_MAXK, PRIFIT_E_SHAPE, "Kcap must be in 1..64");
    PF_CHECK_ARG(version == 3 || version == 4, PRIFIT_E_BADARG, "version must be 3 or 4");
    PF_CHECK_ARG(ws_bytes >= prifit_intersect_workspace_bytes(B, M), PRIFIT_E_WS, "workspace too small");
    const int nblk = (M + SDF_THREADS - 1) / SDF_THREADS;
    float* partial = static_cast<float*>(ws);
    if (version == 3)
        intersect_fwd_kernel<3><<<dim3(nblk, B), SDF_THREADS, 0, pf_stream(stream)>>>(Q, s, V, c, valid, K, M, Kcap, kstar_out, aux_out, partial);
    else
        intersect_fwd_kernel<4><<<dim3(nblk, B), SDF_THREADS, 0, pf_stream(stream)>>>(Q, s, V, c, valid, K, M, Kcap, kstar_out, aux_out, partial);
    PF_LAUNCH_CHECK();
    intersect_finalize_kernel<<<B, 32, 0, pf_stream(stream)>>>(partial, valid, K, nblk, M, Kcap, loss_out, counted_out);
    PF_LAUNCH_CHECK();
    return 0;
}

extern "C" int prifit_intersect_bwd(const float* Q, const float* s, const float* V, const float* c, const uint8_t* valid,
                                    const int32_t* K, const int32_t* kstar, const float* aux, const float* gloss, int B, int M,
                                    int Kcap, int version, float* gs_out, float* gV_out, float* gc_out, void* stream) {
    PF_CHECK_ARG(Q && s && V && c && valid && K && kstar && aux && gloss && gs_out && gV_out && gc_out, PRIFIT_E_BADARG, "null pointer");
    PF_CHECK_ARG(B > 0 && M > 0, PRIFIT_E_BADARG, "B, M > 0 required");
    PF_CHECK_ARG(Kcap > 0 && Kcap <= SDF_MAXK, PRIFIT_E_SHAPE, "Kcap must be in 1..64");
    PF_CHECK_ARG(version == 3 || version == 4, PRIFIT_E_BADARG, "version must be 3 or 4");
    if (version == 3)
        intersect_bwd_kernel<3><<<dim3(Kcap, B), SDF_THREADS, 0, pf_stream(stream)>>>(Q, s, V, c, valid, K, kstar, aux, gloss, M, Kcap, gs_out, gV_out, gc_out);
    else
        intersect_bwd_kernel<4><<<dim3(Kcap, B), SDF_THREADS, 0, pf_stream(stream)>>>(Q, s, V, c, valid, K, kstar, aux, gloss, M, Kcap, gs_out, gV_out, gc_out);
    PF_LAUNCH_CHECK();
    return 0;
}
