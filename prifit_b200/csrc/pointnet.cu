// f4 -- the PointNet++ geometric operators in front of the fitting path (models/pointnet_util.py):
//   farthest_point_sample  :63-84    npoint sequential arg-max steps over a running min-distance
//   query_ball_point       :87-107   first `nsample` point indices (ascending) within `radius` of each query, padded
//                                    with the first hit (the reference sorts an N-long index row per query)
//   3-NN interpolation     :287-295  three nearest sampled points per point (the reference sorts all S distances),
//                                    inverse-distance weights, weighted gather of the features; backward to the features
// Index kernels reproduce the reference's arithmetic order where it decides an index: FPS uses (dx^2 + dy^2) + dz^2 without
// FMA contraction, the two others the expanded form |a|^2 + |b|^2 - 2 <a, b> of square_distance (:18-41).
#include "common.cuh"

namespace {

constexpr int FPS_THREADS = 512;
constexpr int FPS_MAX_PER_THREAD = 32;       // N <= 16384

__global__ void __launch_bounds__(FPS_THREADS) fps_kernel(const float* __restrict__ xyz, const int64_t* __restrict__ start,
                                                          int N, int npoint, int64_t* __restrict__ out) {
    extern __shared__ float pts[];           // [N][3]
    __shared__ float wv[FPS_THREADS / 32];
    __shared__ int wi[FPS_THREADS / 32];
    __shared__ int far_s;
    const int b = blockIdx.x, tid = threadIdx.x;
    const float* p = xyz + (size_t)b * N * 3;
    for (int e = tid; e < N * 3; e += FPS_THREADS) pts[e] = p[e];
    float dist[FPS_MAX_PER_THREAD];
#pragma unroll
    for (int i = 0; i < FPS_MAX_PER_THREAD; ++i) dist[i] = 1e10f;
    if (tid == 0) far_s = (int)min(max(start[b], (int64_t)0), (int64_t)(N - 1));
    __syncthreads();
    for (int it = 0; it < npoint; ++it) {
        const int far = far_s;
        if (tid == 0) out[(size_t)b * npoint + it] = far;
        const float cx = pts[3 * far], cy = pts[3 * far + 1], cz = pts[3 * far + 2];
        float bv = -1.f;
        int bi = 0x7fffffff;
#pragma unroll
        for (int i = 0; i < FPS_MAX_PER_THREAD; ++i) {
            const int j = tid + FPS_THREADS * i;
            if (j < N) {
                const float dx = pts[3 * j] - cx, dy = pts[3 * j + 1] - cy, dz = pts[3 * j + 2] - cz;
                const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                if (d < dist[i]) dist[i] = d;                       // distance[mask] = dist[mask]
                if (dist[i] > bv) { bv = dist[i]; bi = j; }         // ascending j within the thread: first index on ties
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        __syncthreads();                                             // far_s consumed by everyone
        if ((tid & 31) == 0) { wv[tid >> 5] = bv; wi[tid >> 5] = bi; }
        __syncthreads();
        if (tid == 0) {
            float v = wv[0];
            int i0 = wi[0];
            for (int w = 1; w < FPS_THREADS / 32; ++w)
                if (wv[w] > v || (wv[w] == v && wi[w] < i0)) { v = wv[w]; i0 = wi[w]; }
            far_s = i0;                                              // torch.max(distance, -1)[1]: first maximal index
        }
        __syncthreads();
    }
}

__device__ __forceinline__ float sq_expanded(float ax, float ay, float az, float a2, float bx, float by, float bz, float b2) {
    // square_distance (:36-40): -2 <a, b>, then + |a|^2, then + |b|^2
    const float dot = __fadd_rn(__fadd_rn(__fmul_rn(ax, bx), __fmul_rn(ay, by)), __fmul_rn(az, bz));
    return __fadd_rn(__fadd_rn(__fmul_rn(-2.0f, dot), a2), b2);
}
__device__ __forceinline__ float sq_norm(float x, float y, float z) {
    return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
}

constexpr int PN_THREADS = 128;
constexpr int PN_TILE = 1024;

// one thread per query point; xyz staged through shared memory in tiles, scanned in ascending index order
__global__ void __launch_bounds__(PN_THREADS) ball_query_kernel(const float* __restrict__ xyz, const float* __restrict__ new_xyz,
                                                                int N, int S, float r2, int nsample, int64_t* __restrict__ out) {
    __shared__ float4 ts[PN_TILE];
    const int b = blockIdx.y, s = blockIdx.x * PN_THREADS + threadIdx.x;
    const bool live = s < S;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (live) { const float* q = new_xyz + ((size_t)b * S + s) * 3; qx = q[0]; qy = q[1]; qz = q[2]; }
    const float q2 = sq_norm(qx, qy, qz);
    int64_t* o = out + ((size_t)b * S + (live ? s : 0)) * nsample;
    int cnt = 0;
    int64_t first = N;
    for (int j0 = 0; j0 < N; j0 += PN_TILE) {
        const int nj = min(PN_TILE, N - j0);
        __syncthreads();
        for (int e = threadIdx.x; e < nj; e += PN_THREADS) {
            const float* t = xyz + ((size_t)b * N + j0 + e) * 3;
            ts[e] = make_float4(t[0], t[1], t[2], sq_norm(t[0], t[1], t[2]));
        }
        __syncthreads();
        if (live && cnt < nsample) {
            for (int j = 0; j < nj && cnt < nsample; ++j) {
                const float4 t = ts[j];
                const float d = sq_expanded(qx, qy, qz, q2, t.x, t.y, t.z, t.w);
                if (!(d > r2)) {                                     // group_idx[sqrdists > radius ** 2] = N
                    if (cnt == 0) first = j0 + j;
                    o[cnt++] = j0 + j;
                }
            }
        }
    }
    if (live)
        for (int k = cnt; k < nsample; ++k) o[k] = first;           // pad with the first hit (N when there is none, like the reference)
}

// three nearest of xyz2[B,S,3] for every point of xyz1[B,N,3]; ties keep the lower index first
__global__ void __launch_bounds__(PN_THREADS) three_nn_kernel(const float* __restrict__ xyz1, const float* __restrict__ xyz2,
                                                              int N, int S, int32_t* __restrict__ idx, float* __restrict__ weight) {
    __shared__ float4 ts[PN_TILE];
    const int b = blockIdx.y, n = blockIdx.x * PN_THREADS + threadIdx.x;
    const bool live = n < N;
    float px = 0.f, py = 0.f, pz = 0.f;
    if (live) { const float* p = xyz1 + ((size_t)b * N + n) * 3; px = p[0]; py = p[1]; pz = p[2]; }
    const float p2 = sq_norm(px, py, pz);
    float d0 = INFINITY, d1 = INFINITY, d2 = INFINITY;
    int i0 = 0, i1 = 0, i2 = 0;
    for (int j0 = 0; j0 < S; j0 += PN_TILE) {
        const int nj = min(PN_TILE, S - j0);
        __syncthreads();
        for (int e = threadIdx.x; e < nj; e += PN_THREADS) {
            const float* t = xyz2 + ((size_t)b * S + j0 + e) * 3;
            ts[e] = make_float4(t[0], t[1], t[2], sq_norm(t[0], t[1], t[2]));
        }
        __syncthreads();
        if (live) {
            for (int j = 0; j < nj; ++j) {
                const float4 t = ts[j];
                const float d = sq_expanded(px, py, pz, p2, t.x, t.y, t.z, t.w);
                if (d < d2) {
                    if (d < d1) {
                        d2 = d1; i2 = i1;
                        if (d < d0) { d1 = d0; i1 = i0; d0 = d; i0 = j0 + j; } else { d1 = d; i1 = j0 + j; }
                    } else { d2 = d; i2 = j0 + j; }
                }
            }
        }
    }
    if (live) {
        const float r0 = 1.0f / (d0 + 1e-8f), r1 = 1.0f / (d1 + 1e-8f), r2 = 1.0f / (d2 + 1e-8f);   // :291-293
        const float norm = (r0 + r1) + r2;
        int32_t* io = idx + ((size_t)b * N + n) * 3;
        float* wo = weight + ((size_t)b * N + n) * 3;
        io[0] = i0; io[1] = i1; io[2] = i2;
        wo[0] = r0 / norm; wo[1] = r1 / norm; wo[2] = r2 / norm;
    }
}

// out[b, n, c] = sum_k weight[b, n, k] * points2[b, idx[b, n, k], c]          (:294)
__global__ void interpolate_fwd_kernel(const float* __restrict__ points2, const int32_t* __restrict__ idx, const float* __restrict__ weight,
                                       int N, int S, int D, float* __restrict__ out) {
    const int b = blockIdx.z, n = blockIdx.y;
    const int32_t* io = idx + ((size_t)b * N + n) * 3;
    const float* wo = weight + ((size_t)b * N + n) * 3;
    const float* f0 = points2 + ((size_t)b * S + io[0]) * D;
    const float* f1 = points2 + ((size_t)b * S + io[1]) * D;
    const float* f2 = points2 + ((size_t)b * S + io[2]) * D;
    const float w0 = wo[0], w1 = wo[1], w2 = wo[2];
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < D; c += gridDim.x * blockDim.x)
        out[((size_t)b * N + n) * D + c] = (f0[c] * w0 + f1[c] * w1) + f2[c] * w2;
}

__global__ void interpolate_bwd_kernel(const float* __restrict__ gout, const int32_t* __restrict__ idx, const float* __restrict__ weight,
                                       int N, int S, int D, float* __restrict__ gpoints2) {
    const int b = blockIdx.z, n = blockIdx.y;
    const int32_t* io = idx + ((size_t)b * N + n) * 3;
    const float* wo = weight + ((size_t)b * N + n) * 3;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < D; c += gridDim.x * blockDim.x) {
        const float g = gout[((size_t)b * N + n) * D + c];
#pragma unroll
        for (int k = 0; k < 3; ++k) atomicAdd(gpoints2 + ((size_t)b * S + io[k]) * D + c, g * wo[k]);
    }
}

}  // namespace

extern "C" int prifit_fps(const float* xyz, const int64_t* start, int B, int N, int npoint, int64_t* idx_out, void* stream) {
    PF_CHECK_ARG(xyz && start && idx_out, PRIFIT_E_BADARG, "null pointer");
    PF_CHECK_ARG(B > 0 && N > 0 && npoint > 0, PRIFIT_E_BADARG, "B, N, npoint > 0 required");
    PF_CHECK_ARG(N <= FPS_THREADS * FPS_MAX_PER_THREAD && (size_t)N * 12 <= 200 * 1024, PRIFIT_E_SHAPE, "N must be <= 16384");
    const size_t smem = (size_t)N * 3 * sizeof(float);
    PF_CUDA(cudaFuncSetAttribute(fps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fps_kernel<<<B, FPS_THREADS, smem, pf_stream(stream)>>>(xyz, start, N, npoint, idx_out);
    PF_LAUNCH_CHECK();
    return 0;
}

extern "C" int prifit_ball_query(const float* xyz, const float* new_xyz, int B, int N, int S, float radius, int nsample,
                                 int64_t* idx_out, void* stream) {
    PF_CHECK_ARG(xyz && new_xyz && idx_out, PRIFIT_E_BADARG, "null pointer");
    PF_CHECK_ARG(B > 0 && N > 0 && S > 0 && nsample > 0, PRIFIT_E_BADARG, "B, N, S, nsample > 0 required");
    ball_query_kernel<<<dim3((S + PN_THREADS - 1) / PN_THREADS, B), PN_THREADS, 0, pf_stream(stream)>>>(xyz, new_xyz, N, S, radius * radius, nsample, idx_out);
    PF_LAUNCH_CHECK();
    return 0;
}

extern "C" int prifit_three_nn(const float* xyz1, const float* xyz2, int B, int N, int S, int32_t* idx_out, float* weight_out, void* stream) {
    PF_CHECK_ARG(xyz1 && xyz2 && idx_out && weight_out, PRIFIT_E_BADARG, "null pointer");
    PF_CHECK_ARG(B > 0 && N > 0 && S >= 3, PRIFIT_E_BADARG, "B, N > 0 and S >= 3 required");
    three_nn_kernel<<<dim3((N + PN_THREADS - 1) / PN_THREADS, B), PN_THREADS, 0, pf_stream(stream)>>>(xyz1, xyz2, N, S, idx_out, weight_out);
    PF_LAUNCH_CHECK();
    return 0;
}

extern "C" int prifit_interpolate_fwd(const float* points2, const int32_t* idx, const float* weight, int B, int N, int S, int D,
                                      float* out, void* stream) {
    PF_CHECK_ARG(points2 && idx && weight && out, PRIFIT_E_BADARG, "null pointer");
    PF_CHECK_ARG(B > 0 && N > 0 && S > 0 && D > 0, PRIFIT_E_BADARG, "B, N, S, D > 0 required");
    interpolate_fwd_kernel<<<dim3((D + 127) / 128, N, B), 128, 0, pf_stream(stream)>>>(points2, idx, weight, N, S, D, out);
    PF_LAUNCH_CHECK();
    return 0;
}

extern "C" int prifit_interpolate_bwd(const float* gout, const int32_t* idx, const float* weight, int B, int N, int S, int D,
                                      float* gpoints2_inout, void* stream) {
    PF_CHECK_ARG(gout && idx && weight && gpoints2_inout, PRIFIT_E_BADARG, "null pointer");
    PF_CHECK_ARG(B > 0 && N > 0 && S > 0 && D > 0, PRIFIT_E_BADARG, "B, N, S, D > 0 required");
    interpolate_bwd_kernel<<<dim3((D + 127) / 128, N, B), 128, 0, pf_stream(stream)>>>(gout, idx, weight, N, S, D, gpoints2_inout);
    PF_LAUNCH_CHECK();
    return 0;
}
