// k2 (fp32 CUDA-core engine) -- T mean-shift iterations of all N seeds of every shape.
// reference src/mean_shift.py:50-84 (mean_shift_, gaussian branch), guard_exp src/guard.py:6-11.
//
// Flash-attention shaped: a CTA owns 64 seed rows and keeps them in shared memory across all T
// iterations (rows are independent given X); the key rows stream through a 64-key tile; the
// 64 x 64 kernel tile lives in registers/shared memory only.  This engine is the fp32 cross-check
// of the tcgen05 kernel (meanshift_tc.cu) and serves embedding widths other than 128.
#include "common.cuh"

namespace {

constexpr int MS_THREADS = 256;
constexpr int MS_BM = 64;   // seed rows per CTA
constexpr int MS_BN = 64;   // keys per tile

template <int D>
__global__ void __launch_bounds__(MS_THREADS) meanshift_simt_kernel(
    const float* __restrict__ X, const float* __restrict__ bw, int N, int T, float* __restrict__ newX) {
    constexpr int LD = D + 4;
    constexpr int LDP = MS_BN + 4;
    constexpr int NH = D / 64;          // float4 column groups per thread in the O micro-tile
    static_assert(D % 64 == 0, "D must be a multiple of 64");
    extern __shared__ __align__(16) float smem[];
    float* ys = smem;                   // [64][LD] current seeds
    float* xs = ys + MS_BM * LD;        // [64][LD] key tile
    float* ps = xs + MS_BN * LD;        // [64][LDP] kernel tile

    const int b = blockIdx.y, r0 = blockIdx.x * MS_BM;
    const float* Xb = X + (size_t)b * N * D;
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const float bwv = bw[b];
    const float b2 = bwv * bwv;         // b ** 2

    for (int e = tid; e < MS_BM * (D / 4); e += MS_THREADS) {
        const int r = e / (D / 4), c = e - r * (D / 4);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r0 + r < N) v = reinterpret_cast<const float4*>(Xb + (size_t)(r0 + r) * D)[c];
        *reinterpret_cast<float4*>(ys + r * LD + 4 * c) = v;          // new_X = X.clone(), line 60
    }

    float o[4][NH * 4];
    for (int t = 0; t < T; ++t) {
        float z[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < NH * 4; ++c) o[a][c] = 0.f;

        for (int j0 = 0; j0 < N; j0 += MS_BN) {
            __syncthreads();
            for (int e = tid; e < MS_BN * (D / 4); e += MS_THREADS) {
                const int r = e / (D / 4), c = e - r * (D / 4);
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (j0 + r < N) v = reinterpret_cast<const float4*>(Xb + (size_t)(j0 + r) * D)[c];
                *reinterpret_cast<float4*>(xs + r * LD + 4 * c) = v;
            }
            __syncthreads();
            // S = Y X^T on a 4x4 micro-tile: rows ty + 16a, keys tx + 16c
            float s[4][4];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int c = 0; c < 4; ++c) s[a][c] = 0.f;
#pragma unroll 4
            for (int i = 0; i < D; i += 4) {
                float4 y[4], x[4];
#pragma unroll
                for (int a = 0; a < 4; ++a) y[a] = *reinterpret_cast<const float4*>(ys + (ty + 16 * a) * LD + i);
#pragma unroll
                for (int c = 0; c < 4; ++c) x[c] = *reinterpret_cast<const float4*>(xs + (tx + 16 * c) * LD + i);
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        s[a][c] = fmaf(y[a].x, x[c].x, s[a][c]);
                        s[a][c] = fmaf(y[a].y, x[c].y, s[a][c]);
                        s[a][c] = fmaf(y[a].z, x[c].z, s[a][c]);
                        s[a][c] = fmaf(y[a].w, x[c].w, s[a][c]);
                    }
            }
            // K = guard_exp(-dist / b^2 / 2), dist = 2 - 2 s  (lines 65-68); keys beyond N weigh 0
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                float part = 0.f;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float dist = 2.0f - 2.0f * s[a][c];
                    float p = guard_expf((-dist / b2) * 0.5f);
                    if (j0 + tx + 16 * c >= N) p = 0.f;
                    ps[(ty + 16 * a) * LDP + tx + 16 * c] = p;
                    part += p;
                }
#pragma unroll
                for (int off = 8; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
                z[a] += part;
            }
            __syncthreads();
            // O += P X_tile on a 4 x (4*NH) micro-tile: rows ty + 16a, columns 4*tx + 64h
            for (int c0 = 0; c0 < MS_BN; c0 += 4) {
                float4 p[4];
#pragma unroll
                for (int a = 0; a < 4; ++a) p[a] = *reinterpret_cast<const float4*>(ps + (ty + 16 * a) * LDP + c0);
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
#pragma unroll
                    for (int h = 0; h < NH; ++h) {
                        const float4 x = *reinterpret_cast<const float4*>(xs + (c0 + cc) * LD + 4 * tx + 64 * h);
#pragma unroll
                        for (int a = 0; a < 4; ++a) {
                            const float pv = cc == 0 ? p[a].x : cc == 1 ? p[a].y : cc == 2 ? p[a].z : p[a].w;
                            o[a][4 * h + 0] = fmaf(pv, x.x, o[a][4 * h + 0]);
                            o[a][4 * h + 1] = fmaf(pv, x.y, o[a][4 * h + 1]);
                            o[a][4 * h + 2] = fmaf(pv, x.z, o[a][4 * h + 2]);
                            o[a][4 * h + 3] = fmaf(pv, x.w, o[a][4 * h + 3]);
                        }
                    }
                }
            }
        }
        // new_X = y + ((K X) D - y); new_X /= ||new_X||   (lines 75-82)
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const float dinv = 1.0f / z[a];
            float* yr = ys + (ty + 16 * a) * LD;
            float nrm = 0.f;
#pragma unroll
            for (int h = 0; h < NH; ++h) {
                const float4 y = *reinterpret_cast<const float4*>(yr + 4 * tx + 64 * h);
                const float yv[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float m = o[a][4 * h + q] * dinv - yv[q];
                    const float u = yv[q] + m;
                    o[a][4 * h + q] = u;
                    nrm = fmaf(u, u, nrm);
                }
            }
#pragma unroll
            for (int off = 8; off > 0; off >>= 1) nrm += __shfl_xor_sync(0xffffffffu, nrm, off);
            nrm = sqrtf(nrm);
#pragma unroll
            for (int h = 0; h < NH; ++h) {
                float4 y;
                y.x = o[a][4 * h + 0] / nrm; y.y = o[a][4 * h + 1] / nrm;
                y.z = o[a][4 * h + 2] / nrm; y.w = o[a][4 * h + 3] / nrm;
                *reinterpret_cast<float4*>(yr + 4 * tx + 64 * h) = y;
            }
        }
    }
    __syncthreads();
    for (int e = tid; e < MS_BM * (D / 4); e += MS_THREADS) {
        const int r = e / (D / 4), c = e - r * (D / 4);
        if (r0 + r < N)
            reinterpret_cast<float4*>(newX + ((size_t)b * N + r0 + r) * D)[c] = *reinterpret_cast<const float4*>(ys + r * LD + 4 * c);
    }
}

template <int D>
int launch_simt(const float* X, const float* bw, int B, int N, int T, float* newX, cudaStream_t st) {
    const size_t smem = ((size_t)(MS_BM + MS_BN) * (D + 4) + (size_t)MS_BM * (MS_BN + 4)) * sizeof(float);
    PF_CUDA(cudaFuncSetAttribute(meanshift_simt_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((N + MS_BM - 1) / MS_BM, B);
    meanshift_simt_kernel<D><<<grid, MS_THREADS, smem, st>>>(X, bw, N, T, newX);
    PF_LAUNCH_CHECK();
    return 0;
}

}  // namespace

int prifit_meanshift_fwd_simt(const float* X, const float* bw, int B, int N, int d, int T, float* newX, cudaStream_t st) {
    switch (d) {
        case 64: return launch_simt<64>(X, bw, B, N, T, newX, st);
        case 128: return launch_simt<128>(X, bw, B, N, T, newX, st);
        case 256: return launch_simt<256>(X, bw, B, N, T, newX, st);
        default:
            prifit_set_error("prifit_meanshift_fwd: fp32 engine supports d in {64,128,256}, got %d", d);
            return PRIFIT_E_SHAPE;
    }
}
