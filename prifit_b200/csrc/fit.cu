// k5-k7 -- soft-membership weighted ellipsoid fit, one CTA per (shape, cluster), forward + backward.
// reference src/ellipsoid_fitting.py:19-69 (weighted_ellipsoid_fitting), :119-141
// (principal_axis_ellipsoid, mode "slow"), src/fitting_utils.py:67-139 (CustomSVD and its backward).
//
//   W = sum w ; c = sum w p / W ; q = p - c ; cov = sum w q q^T / W            (10 weighted moments)
//   A = cov + 1e-4 mean(cov) R                                                  (R = host U[0,1) draw, :37-38)
//   drop if sigma_0 / sigma_2 > 1e5 or anything is non-finite                   (:41-47, :52-69)
//   (U, S, V) = svd(A);  if det(V) < 0: V[:,2] = -V[:,2]                        (:48, :133-135)
//   t = (w q) V ; s_a = |max_j t_ja - min_j t_ja| / 2                           (:136-140)
//
// Data movement (HBM/L2-latency bound: 8 KB of W and 24 KB of P per CTA at N = 2048).  A CTA reads its membership row
// W[b,k,:] and the shape's points ONCE, as float4 (four weights / four points = three float4 per thread and step, fully
// coalesced), and parks them as (w, x, y, z) records in shared memory; the three passes the algorithm needs (centre;
// covariance about that centre -- two-pass like the reference, so no cancellation for off-centre clusters; extents
// along the principal axes) then run out of shared memory with warp-shuffle reductions.  The 3x3 SVD between passes 2
// and 3 is a one-sided Jacobi iteration in fp64 held entirely in registers (every loop unrolled, columns addressed by
// compile-time indices): the round-1 version indexed its 3x3 arrays by run-time (p, q) and lived in local memory, which
// made the single-thread SVD the longest phase of the kernel.  Thread 0 runs it; its results reach the CTA through shared memory
// and leave as coalesced stores.
#include "common.cuh"

namespace {

constexpr int FIT_THREADS = 256;
constexpr int FIT_SMEM_MAX_N = 12288;       // records kept in shared memory up to this N (192 KB); beyond: re-read from L2

// ctx layout (floats)
constexpr int CX_U = 0, CX_S = 9, CX_V = 12, CX_FLIP = 21, CX_W = 22, CX_C = 23, CX_COV = 26,
              CX_AMAX = 35, CX_AMIN = 38, CX_SGN = 41;
static_assert(CX_SGN + 3 <= PRIFIT_FIT_CTX, "ctx too small");

// One Jacobi rotation of columns (P, Q) of G (and of V), everything in registers.
template <int P, int Q>
__device__ __forceinline__ bool jacobi_rotate(double (&G)[3][3], double (&Vm)[3][3], double tol) {
    const double alpha = G[0][P] * G[0][P] + G[1][P] * G[1][P] + G[2][P] * G[2][P];
    const double beta = G[0][Q] * G[0][Q] + G[1][Q] * G[1][Q] + G[2][Q] * G[2][Q];
    const double gamma = G[0][P] * G[0][Q] + G[1][P] * G[1][Q] + G[2][P] * G[2][Q];
    if (!(fabs(gamma) > tol * sqrt(alpha * beta)) || gamma == 0.0) return false;
    const double zeta = (beta - alpha) / (2.0 * gamma);
    const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
    const double cs = rsqrt(1.0 + t * t), sn = cs * t;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double gp = G[i][P], gq = G[i][Q];
        G[i][P] = cs * gp - sn * gq; G[i][Q] = sn * gp + cs * gq;
        const double vp = Vm[i][P], vq = Vm[i][Q];
        Vm[i][P] = cs * vp - sn * vq; Vm[i][Q] = sn * vp + cs * vq;
    }
    return true;
}

// One-sided Jacobi SVD of a 3x3 matrix (row-major A): A = U diag(S) V^T, S descending.  Converged when every pair of
// columns of G = A V is orthogonal to 1e-15 relative (fp64 round-off); the input itself is fp32.
__device__ __forceinline__ void svd3(const double* A, double* U, double* S, double* V) {
    double G[3][3], Vm[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) { G[i][j] = A[3 * i + j]; Vm[i][j] = i == j ? 1.0 : 0.0; }
    for (int sweep = 0; sweep < 40; ++sweep) {
        bool rotated = jacobi_rotate<0, 1>(G, Vm, 1e-15);
        rotated |= jacobi_rotate<0, 2>(G, Vm, 1e-15);
        rotated |= jacobi_rotate<1, 2>(G, Vm, 1e-15);
        if (!rotated) break;
    }
    double sv[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) sv[j] = sqrt(G[0][j] * G[0][j] + G[1][j] * G[1][j] + G[2][j] * G[2][j]);
    // descending order by a 3-element sorting network on (value, column) held in registers
#define FIT_CSWAP(a, b)                                                                      \
    if (sv[b] > sv[a]) {                                                                     \
        double t_ = sv[a]; sv[a] = sv[b]; sv[b] = t_;                                        \
        _Pragma("unroll") for (int i = 0; i < 3; ++i) {                                      \
            t_ = G[i][a]; G[i][a] = G[i][b]; G[i][b] = t_;                                   \
            t_ = Vm[i][a]; Vm[i][a] = Vm[i][b]; Vm[i][b] = t_;                               \
        }                                                                                    \
    }
    FIT_CSWAP(0, 1) FIT_CSWAP(0, 2) FIT_CSWAP(1, 2)
#undef FIT_CSWAP
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        S[j] = sv[j];
        const double inv = sv[j] > 0 ? 1.0 / sv[j] : 0.0;
#pragma unroll
        for (int i = 0; i < 3; ++i) { U[3 * i + j] = G[i][j] * inv; V[3 * i + j] = Vm[i][j]; }
    }
}

__device__ __forceinline__ double det3(const double* M) {
    return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
}

struct ArgVal { float v; int i; };
__device__ __forceinline__ ArgVal arg_max2(ArgVal a, ArgVal b) { return (b.v > a.v || (b.v == a.v && b.i < a.i)) ? b : a; }
__device__ __forceinline__ ArgVal arg_min2(ArgVal a, ArgVal b) { return (b.v < a.v || (b.v == a.v && b.i < a.i)) ? b : a; }

// (w, x, y, z) of point j: from the shared-memory records, or (N too large for them) straight from global memory / L2
template <bool SMEM>
__device__ __forceinline__ float4 fit_rec(const float4* recs, const float* __restrict__ w, const float* __restrict__ p, int j) {
    if (SMEM) return recs[j];
    return make_float4(w[j], p[3 * j], p[3 * j + 1], p[3 * j + 2]);
}

// Stage the CTA's membership row and the shape's points as (w, x, y, z) records in shared memory, reading global memory
// once with float4 loads (rows are 16-byte aligned exactly when N % 4 == 0), and return this thread's partial
// (sum w, sum w x, sum w y, sum w z) over the points it staged.
__device__ __forceinline__ void fit_stage(float4* recs, const float* __restrict__ w, const float* __restrict__ p, int N, int tid,
                                          float (&m)[4]) {
    if ((N & 3) == 0) {
        const float4* w4 = reinterpret_cast<const float4*>(w);
        const float4* p4 = reinterpret_cast<const float4*>(p);
        for (int g = tid; g < (N >> 2); g += FIT_THREADS) {
            const float4 ww = __ldg(w4 + g), a = __ldg(p4 + 3 * g), b4 = __ldg(p4 + 3 * g + 1), c = __ldg(p4 + 3 * g + 2);
            const float4 r0 = make_float4(ww.x, a.x, a.y, a.z), r1 = make_float4(ww.y, a.w, b4.x, b4.y),
                         r2 = make_float4(ww.z, b4.z, b4.w, c.x), r3 = make_float4(ww.w, c.y, c.z, c.w);
            recs[4 * g] = r0; recs[4 * g + 1] = r1; recs[4 * g + 2] = r2; recs[4 * g + 3] = r3;
            m[0] += (r0.x + r1.x) + (r2.x + r3.x);
            m[1] += (r0.x * r0.y + r1.x * r1.y) + (r2.x * r2.y + r3.x * r3.y);
            m[2] += (r0.x * r0.z + r1.x * r1.z) + (r2.x * r2.z + r3.x * r3.z);
            m[3] += (r0.x * r0.w + r1.x * r1.w) + (r2.x * r2.w + r3.x * r3.w);
        }
    } else {
        for (int j = tid; j < N; j += FIT_THREADS) {
            const float4 r = make_float4(w[j], p[3 * j], p[3 * j + 1], p[3 * j + 2]);
            recs[j] = r;
            m[0] += r.x; m[1] += r.x * r.y; m[2] += r.x * r.z; m[3] += r.x * r.w;
        }
    }
}

template <bool SMEM>
__global__ void __launch_bounds__(FIT_THREADS) fit_fwd_kernel(
    const float* __restrict__ P, const float* __restrict__ Wt, const int32_t* __restrict__ K,
    const float* __restrict__ noise, int N, int Kcap,
    float* __restrict__ s_out, float* __restrict__ V_out, float* __restrict__ c_out,
    uint8_t* __restrict__ valid_out, float* __restrict__ ctx_out) {
    extern __shared__ float4 fit_recs[];
    __shared__ float red[10 * 32];
    __shared__ float vs[9];
    __shared__ float ctx_s[PRIFIT_FIT_CTX];
    __shared__ int ok_s;
    __shared__ float av[FIT_THREADS / 32][6];
    __shared__ int ai[FIT_THREADS / 32][6];
    const int k = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const size_t bk = (size_t)b * Kcap + k;
    float* ctx = ctx_out + bk * PRIFIT_FIT_CTX;
    if (k >= min(K[b], Kcap)) {
        if (tid < 3) { s_out[bk * 3 + tid] = 0.f; c_out[bk * 3 + tid] = 0.f; }
        if (tid < 9) V_out[bk * 9 + tid] = 0.f;
        if (tid < PRIFIT_FIT_CTX) ctx[tid] = 0.f;
        if (tid == 0) valid_out[bk] = 0;
        return;
    }
    const float* w = Wt + bk * N;
    const float* p = P + (size_t)b * N * 3;

    // pass 1: W, sum w p  (while staging the records)
    float m[4] = {0.f, 0.f, 0.f, 0.f};
    if (SMEM) {
        fit_stage(fit_recs, w, p, N, tid, m);
    } else {
        for (int j = tid; j < N; j += FIT_THREADS) {
            const float wj = w[j];
            m[0] += wj; m[1] += wj * p[3 * j]; m[2] += wj * p[3 * j + 1]; m[3] += wj * p[3 * j + 2];
        }
    }
    block_sum<4>(m, red);               // its barriers also publish the records
    const float Wsum = m[0];
    const float cx = m[1] / Wsum, cy = m[2] / Wsum, cz = m[3] / Wsum;

    // pass 2: sum w q q^T
    float cv[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int j = tid; j < N; j += FIT_THREADS) {
        const float4 r = fit_rec<SMEM>(fit_recs, w, p, j);
        const float qx = r.y - cx, qy = r.z - cy, qz = r.w - cz;
        const float wx = r.x * qx, wy = r.x * qy, wz = r.x * qz;
        cv[0] += wx * qx; cv[1] += wx * qy; cv[2] += wx * qz;
        cv[3] += wy * qy; cv[4] += wy * qz; cv[5] += wz * qz;
    }
    block_sum<6>(cv, red);

    if (tid == 0) {
        float cov[9] = {cv[0] / Wsum, cv[1] / Wsum, cv[2] / Wsum, cv[1] / Wsum, cv[3] / Wsum, cv[4] / Wsum,
                        cv[2] / Wsum, cv[4] / Wsum, cv[5] / Wsum};
        float mean = 0.f;
#pragma unroll
        for (int i = 0; i < 9; ++i) mean += cov[i];
        mean /= 9.0f;
        const float* R = noise + bk * 9;
        double A[9], U[9], S[3], V[9];
        bool finite = true;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            const float a = cov[i] + (1e-4f * mean) * R[i];
            finite = finite && isfinite(a);
            A[i] = (double)a;
        }
        int ok = 0;
        if (finite) {
            svd3(A, U, S, V);
            ok = !((float)S[0] / (float)S[2] > 1e5f) && S[2] > 0.0;
        }
        ok_s = ok;
        if (ok) {
            const bool flip = det3(V) < 0.0;
#pragma unroll
            for (int i = 0; i < 9; ++i) { ctx_s[CX_U + i] = (float)U[i]; ctx_s[CX_V + i] = (float)V[i]; ctx_s[CX_COV + i] = cov[i]; }
#pragma unroll
            for (int i = 0; i < 3; ++i) ctx_s[CX_S + i] = (float)S[i];
            ctx_s[CX_FLIP] = flip ? 1.f : 0.f;
            ctx_s[CX_W] = Wsum;
            ctx_s[CX_C] = cx; ctx_s[CX_C + 1] = cy; ctx_s[CX_C + 2] = cz;
#pragma unroll
            for (int i = 0; i < 9; ++i) vs[i] = (float)((flip && (i % 3) == 2) ? -V[i] : V[i]);
        }
    }
    __syncthreads();
    if (!ok_s) {
        if (tid < 3) { s_out[bk * 3 + tid] = 0.f; c_out[bk * 3 + tid] = 0.f; }
        if (tid < 9) V_out[bk * 9 + tid] = 0.f;
        if (tid < PRIFIT_FIT_CTX) ctx[tid] = 0.f;
        if (tid == 0) valid_out[bk] = 0;
        return;
    }
    if (tid < 9) V_out[bk * 9 + tid] = vs[tid];
    if (tid == 0) { valid_out[bk] = 1; c_out[bk * 3] = cx; c_out[bk * 3 + 1] = cy; c_out[bk * 3 + 2] = cz; }

    // pass 3: extents of the weighted, rotated points
    ArgVal hi[3], lo[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) { hi[a].v = -INFINITY; hi[a].i = 0x7fffffff; lo[a].v = INFINITY; lo[a].i = 0x7fffffff; }
    for (int j = tid; j < N; j += FIT_THREADS) {
        const float4 r = fit_rec<SMEM>(fit_recs, w, p, j);
        const float rx = (r.y - cx) * r.x, ry = (r.z - cy) * r.x, rz = (r.w - cz) * r.x;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float t = rx * vs[a] + ry * vs[3 + a] + rz * vs[6 + a];
            if (t > hi[a].v) { hi[a].v = t; hi[a].i = j; }
            if (t < lo[a].v) { lo[a].v = t; lo[a].i = j; }
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            ArgVal x, y;
            x.v = __shfl_xor_sync(0xffffffffu, hi[a].v, o); x.i = __shfl_xor_sync(0xffffffffu, hi[a].i, o);
            y.v = __shfl_xor_sync(0xffffffffu, lo[a].v, o); y.i = __shfl_xor_sync(0xffffffffu, lo[a].i, o);
            hi[a] = arg_max2(hi[a], x);
            lo[a] = arg_min2(lo[a], y);
        }
    }
    if ((tid & 31) == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            av[tid >> 5][a] = hi[a].v; ai[tid >> 5][a] = hi[a].i;
            av[tid >> 5][3 + a] = lo[a].v; ai[tid >> 5][3 + a] = lo[a].i;
        }
    }
    __syncthreads();
    if (tid < 3) {
        ArgVal h, l;
        h.v = av[0][tid]; h.i = ai[0][tid]; l.v = av[0][3 + tid]; l.i = ai[0][3 + tid];
        for (int wq = 1; wq < FIT_THREADS / 32; ++wq) {
            ArgVal x, y;
            x.v = av[wq][tid]; x.i = ai[wq][tid]; y.v = av[wq][3 + tid]; y.i = ai[wq][3 + tid];
            h = arg_max2(h, x);
            l = arg_min2(l, y);
        }
        const float diff = h.v - l.v;
        s_out[bk * 3 + tid] = fabsf(diff) / 2.0f;
        ctx_s[CX_AMAX + tid] = __int_as_float(h.i);
        ctx_s[CX_AMIN + tid] = __int_as_float(l.i);
        ctx_s[CX_SGN + tid] = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
    }
    __syncthreads();
    if (tid < PRIFIT_FIT_CTX) ctx[tid] = tid < CX_SGN + 3 ? ctx_s[tid] : 0.f;
}

// ------------------------------------------------------------------------------------------ backward
// Same data movement as the forward: records staged once in shared memory, the 3x3 algebra (extent terms, custom SVD
// backward, noise term) in fp64 registers on thread 0, then one reduction pass (sum_j dL/dq_j -> dL/dc) and the output
// pass that writes dL/dw_j.
template <bool SMEM>
__global__ void __launch_bounds__(FIT_THREADS) fit_bwd_kernel(
    const float* __restrict__ P, const float* __restrict__ Wt, const int32_t* __restrict__ K,
    const float* __restrict__ noise, const float* __restrict__ ctx_in, const uint8_t* __restrict__ valid,
    const float* __restrict__ gs, const float* __restrict__ gV, const float* __restrict__ gc,
    int N, int Kcap, float* __restrict__ gW, float* __restrict__ gP) {
    extern __shared__ float4 fit_recs[];
    __shared__ float red[4 * 32];
    __shared__ float dcov_s[9];     // dL/dcov
    __shared__ float sym_s[9];      // dcov + dcov^T
    __shared__ float sp_dr[6][3];   // sparse dL/dr at the arg-extreme points
    __shared__ int sp_j[6];
    __shared__ float misc[8];       // W, cx, cy, cz, cov:dcov
    const int k = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const size_t bk = (size_t)b * Kcap + k;
    float* gw = gW + bk * N;
    if (k >= min(K[b], Kcap) || !valid[bk]) {
        for (int j = tid; j < N; j += FIT_THREADS) gw[j] = 0.f;
        return;
    }
    const float* ctx = ctx_in + bk * PRIFIT_FIT_CTX;
    const float* w = Wt + bk * N;
    const float* p = P + (size_t)b * N * 3;
    if (SMEM) {
        float unused[4] = {0.f, 0.f, 0.f, 0.f};
        fit_stage(fit_recs, w, p, N, tid, unused);
        __syncthreads();
    }

    if (tid == 0) {
        double U[9], S[3], V[9], Vo[9], cov[9];
        for (int i = 0; i < 9; ++i) { U[i] = ctx[CX_U + i]; V[i] = ctx[CX_V + i]; cov[i] = ctx[CX_COV + i]; }
        for (int i = 0; i < 3; ++i) S[i] = ctx[CX_S + i];
        const bool flip = ctx[CX_FLIP] != 0.f;
        const double Wsum = ctx[CX_W];
        const double c[3] = {ctx[CX_C], ctx[CX_C + 1], ctx[CX_C + 2]};
        for (int i = 0; i < 9; ++i) Vo[i] = (flip && (i % 3) == 2) ? -V[i] : V[i];
        double gVt[9];
        for (int i = 0; i < 9; ++i) gVt[i] = gV[bk * 9 + i];
        // extents: s_a = |hi_a - lo_a| / 2, hi/lo = t at the arg-extreme points, t_ja = r_j . Vo[:,a], r = w q
        for (int a = 0; a < 3; ++a) {
            const double ghi = 0.5 * (double)ctx[CX_SGN + a] * (double)gs[bk * 3 + a];
            const int jj[2] = {__float_as_int(ctx[CX_AMAX + a]), __float_as_int(ctx[CX_AMIN + a])};
            const double gt[2] = {ghi, -ghi};
            for (int e = 0; e < 2; ++e) {
                const int j = jj[e];
                const float4 rec = fit_rec<SMEM>(fit_recs, w, p, j);
                const double wj = rec.x;
                const double r[3] = {((double)rec.y - c[0]) * wj, ((double)rec.z - c[1]) * wj, ((double)rec.w - c[2]) * wj};
                for (int i = 0; i < 3; ++i) {
                    gVt[3 * i + a] += gt[e] * r[i];
                    sp_dr[2 * a + e][i] = (float)(gt[e] * Vo[3 * i + a]);
                }
                sp_j[2 * a + e] = j;
            }
        }
        if (flip) for (int i = 0; i < 3; ++i) gVt[3 * i + 2] = -gVt[3 * i + 2];
        // CustomSVD backward (src/fitting_utils.py:67-105), grad_S = 0 in "slow" mode
        double Kt[9];   // K^T
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                double kij = 0.0;       // K[i][j]
                if (i != j) {
                    const double diff = S[i] - S[j];
                    const double sg = diff > 0 ? 1.0 : (diff < 0 ? -1.0 : 0.0);
                    const double kneg = sg * fmax(fabs(diff), 1e-6);
                    kij = (1.0 / kneg) * (1.0 / (S[i] + S[j]));
                }
                Kt[3 * j + i] = kij;
            }
        double inner[9];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                double vtg = 0.0;
                for (int q = 0; q < 3; ++q) vtg += V[3 * q + i] * gVt[3 * q + j];
                inner[3 * i + j] = Kt[3 * i + j] * vtg;
            }
        double symi[9];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) symi[3 * i + j] = 0.5 * (inner[3 * i + j] + inner[3 * j + i]);
        double tmp[9], dA[9];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                double acc = 0.0;
                for (int q = 0; q < 3; ++q) acc += U[3 * i + q] * S[q] * symi[3 * q + j];
                tmp[3 * i + j] = acc;
            }
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                double acc = 0.0;
                for (int q = 0; q < 3; ++q) acc += tmp[3 * i + q] * V[3 * j + q];
                dA[3 * i + j] = 2.0 * acc;
            }
        // A = cov + 1e-4 mean(cov) R
        const float* R = noise + bk * 9;
        double rdot = 0.0;
        for (int i = 0; i < 9; ++i) rdot += (double)R[i] * dA[i];
        double cdot = 0.0;
        for (int i = 0; i < 9; ++i) {
            const double dc = dA[i] + 1e-4 * rdot / 9.0;
            dcov_s[i] = (float)dc;
            cdot += cov[i] * dc;
        }
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) sym_s[3 * i + j] = dcov_s[3 * i + j] + dcov_s[3 * j + i];
        misc[0] = (float)Wsum; misc[1] = (float)c[0]; misc[2] = (float)c[1]; misc[3] = (float)c[2];
        misc[4] = (float)cdot;
    }
    __syncthreads();
    const float Wsum = misc[0], cx = misc[1], cy = misc[2], cz = misc[3], cdot = misc[4];
    const float invW = 1.0f / Wsum;

    // pass A: G = sum_j dL/dq_j
    float G[3] = {0.f, 0.f, 0.f};
    for (int j = tid; j < N; j += FIT_THREADS) {
        const float4 rec = fit_rec<SMEM>(fit_recs, w, p, j);
        const float wj = rec.x;
        const float q[3] = {rec.y - cx, rec.z - cy, rec.w - cz};
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            float dq = wj * (sym_s[3 * i] * q[0] + sym_s[3 * i + 1] * q[1] + sym_s[3 * i + 2] * q[2]) * invW;
#pragma unroll
            for (int e = 0; e < 6; ++e)
                if (sp_j[e] == j) dq += wj * sp_dr[e][i];
            G[i] += dq;
        }
    }
    block_sum<3>(G, red);
    const float dc[3] = {gc[bk * 3] - G[0], gc[bk * 3 + 1] - G[1], gc[bk * 3 + 2] - G[2]};

    // pass B: dL/dw_j (and optionally dL/dp_j)
    for (int j = tid; j < N; j += FIT_THREADS) {
        const float4 rec = fit_rec<SMEM>(fit_recs, w, p, j);
        const float wj = rec.x;
        const float q[3] = {rec.y - cx, rec.z - cy, rec.w - cz};
        float quad = 0.f;
#pragma unroll
        for (int i = 0; i < 3; ++i)
            quad += q[i] * (dcov_s[3 * i] * q[0] + dcov_s[3 * i + 1] * q[1] + dcov_s[3 * i + 2] * q[2]);
        float g = (quad - cdot) * invW + (q[0] * dc[0] + q[1] * dc[1] + q[2] * dc[2]) * invW;
#pragma unroll
        for (int e = 0; e < 6; ++e)
            if (sp_j[e] == j) g += sp_dr[e][0] * q[0] + sp_dr[e][1] * q[1] + sp_dr[e][2] * q[2];
        gw[j] = g;
        if (gP) {
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                float dq = wj * (sym_s[3 * i] * q[0] + sym_s[3 * i + 1] * q[1] + sym_s[3 * i + 2] * q[2]) * invW;
#pragma unroll
                for (int e = 0; e < 6; ++e)
                    if (sp_j[e] == j) dq += wj * sp_dr[e][i];
                atomicAdd(gP + ((size_t)b * N + j) * 3 + i, dq + wj * invW * dc[i]);
            }
        }
    }
}

}  // namespace

// noise[b, k] = flat[(sum_{b' < b} K[b']) + k] for k < K[b], 0 beyond: the reference draws one torch.rand(3, 3) per
// attempted cluster, shapes outer / clusters inner (src/ellipsoid_fitting.py:38), so cluster (b, k) owns draw number
// prefix(b) + k of the host stream.  Done on the device so that the host need not know K to stage the draws.
__global__ void noise_scatter_kernel(const float* __restrict__ flat, const int32_t* __restrict__ K, int Kcap,
                                     const int32_t* __restrict__ direct, float* __restrict__ noise, int b0) {
    const int b = b0 + blockIdx.x;
    int off = 0;
    if (direct && *direct) {
        off = b * Kcap;                                  // flat is already laid out [B, Kcap, 3, 3]
    } else {
        for (int q = 0; q < b; ++q) off += min(max(K[q], 0), Kcap);
    }
    const int Kb = min(max(K[b], 0), Kcap);
    for (int e = threadIdx.x; e < Kcap * 9; e += blockDim.x)
        noise[(size_t)b * Kcap * 9 + e] = e < Kb * 9 ? flat[(size_t)off * 9 + e] : 0.f;
}

extern "C" int prifit_noise_scatter(const float* flat, const int32_t* K, int B, int Kcap, const int32_t* direct,
                                    float* noise_out, void* stream) {
    PF_CHECK_ARG(flat && K && noise_out, PRIFIT_E_BADARG, "null pointer");
    PF_CHECK_ARG(B > 0 && Kcap > 0, PRIFIT_E_BADARG, "B, Kcap > 0 required");
    noise_scatter_kernel<<<B, 128, 0, pf_stream(stream)>>>(flat, K, Kcap, direct, noise_out, 0);
    PF_LAUNCH_CHECK();
    return 0;
}

// the same for the shapes [b0, b0 + Bb) of the batch only (pointers are those of the whole batch): cluster (b, k) needs the
// counts of the shapes in front of it, so a group of shapes can be served as soon as the groups before it have their counts
extern "C" int prifit_noise_scatter_range(const float* flat, const int32_t* K, int b0, int Bb, int Kcap, const int32_t* direct,
                                          float* noise_out, void* stream) {
    PF_CHECK_ARG(flat && K && noise_out, PRIFIT_E_BADARG, "null pointer");
    PF_CHECK_ARG(b0 >= 0 && Bb > 0 && Kcap > 0, PRIFIT_E_BADARG, "b0 >= 0, Bb, Kcap > 0 required");
    noise_scatter_kernel<<<Bb, 128, 0, pf_stream(stream)>>>(flat, K, Kcap, direct, noise_out, b0);
    PF_LAUNCH_CHECK();
    return 0;
}

// Device-side gate: one thread spins (with back-off) until *flag >= want, then the stream it was launched on goes on.  The
// graph step bumps a device counter when every branch has left the throughput-bound cluster stage; a caller's prefetch stream
// waits here so that its host->device copy (which goes through L2) lands behind the all-seed kernel whatever the host's timing.
// Bounded: gives up after ~50 ms (a gate must never hang a stream).
__global__ void spin_until_ge_kernel(const volatile int32_t* flag, int32_t want) {
    const long long t0 = clock64();
    while (*flag < want) {
        __nanosleep(500);
        if (clock64() - t0 > 100000000LL) break;
    }
}

extern "C" int prifit_spin_until_ge(const int32_t* flag, int32_t want, void* stream) {
    PF_CHECK_ARG(flag, PRIFIT_E_BADARG, "null pointer");
    spin_until_ge_kernel<<<1, 1, 0, pf_stream(stream)>>>(flag, want);
    PF_LAUNCH_CHECK();
    return 0;
}

// [K[0..B) | n_labels[0..B) | ++serial] -> out[2B + 1]: what the host's guard decision needs, in one buffer for one D2H copy;
// the serial number tells the polling host that the copy it sees belongs to this replay
__global__ void pack_counts_kernel(const int32_t* __restrict__ K, const int32_t* __restrict__ nlab, int B,
                                   int32_t* __restrict__ serial, int32_t* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) { out[i] = K[i]; out[B + i] = nlab[i]; }
    if (i == 0) { const int32_t s = *serial + 1; *serial = s; out[2 * B] = s; }
}

extern "C" int prifit_pack_counts(const int32_t* K, const int32_t* nlab, int B, int32_t* serial_inout, int32_t* out, void* stream) {
    PF_CHECK_ARG(K && nlab && serial_inout && out, PRIFIT_E_BADARG, "null pointer");
    PF_CHECK_ARG(B > 0, PRIFIT_E_BADARG, "B > 0 required");
    pack_counts_kernel<<<(B + 127) / 128, 128, 0, pf_stream(stream)>>>(K, nlab, B, serial_inout, out);
    PF_LAUNCH_CHECK();
    return 0;
}

extern "C" int prifit_fit_fwd(const float* P, const float* W, const int32_t* K, const float* noise,
                              int B, int N, int Kcap, float* s_out, float* V_out, float* c_out,
                              uint8_t* valid_out, float* ctx_out, void* stream) {
    PF_CHECK_ARG(P && W && K && noise && s_out && V_out && c_out && valid_out && ctx_out, PRIFIT_E_BADARG, "null pointer");
    PF_CHECK_ARG(B > 0 && N > 0 && Kcap > 0, PRIFIT_E_BADARG, "B, N, Kcap > 0 required");
    if (N <= FIT_SMEM_MAX_N) {
        const size_t smem = (size_t)N * sizeof(float4);
        if (smem > 32 * 1024)
            PF_CUDA(cudaFuncSetAttribute(fit_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        fit_fwd_kernel<true><<<dim3(Kcap, B), FIT_THREADS, smem, pf_stream(stream)>>>(P, W, K, noise, N, Kcap, s_out, V_out, c_out, valid_out, ctx_out);
    } else {
        fit_fwd_kernel<false><<<dim3(Kcap, B), FIT_THREADS, 0, pf_stream(stream)>>>(P, W, K, noise, N, Kcap, s_out, V_out, c_out, valid_out, ctx_out);
    }
    PF_LAUNCH_CHECK();
    return 0;
}

extern "C" int prifit_fit_bwd(const float* P, const float* W, const int32_t* K, const float* noise,
                              const float* ctx, const uint8_t* valid, const float* gs, const float* gV, const float* gc,
                              int B, int N, int Kcap, float* gW_out, float* gP_inout, void* stream) {
    PF_CHECK_ARG(P && W && K && noise && ctx && valid && gs && gV && gc && gW_out, PRIFIT_E_BADARG, "null pointer");
    PF_CHECK_ARG(B > 0 && N > 0 && Kcap > 0, PRIFIT_E_BADARG, "B, N, Kcap > 0 required");
    if (N <= FIT_SMEM_MAX_N) {
        const size_t smem = (size_t)N * sizeof(float4);
        if (smem > 32 * 1024)
            PF_CUDA(cudaFuncSetAttribute(fit_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        fit_bwd_kernel<true><<<dim3(Kcap, B), FIT_THREADS, smem, pf_stream(stream)>>>(P, W, K, noise, ctx, valid, gs, gV, gc, N, Kcap, gW_out, gP_inout);
    } else {
        fit_bwd_kernel<false><<<dim3(Kcap, B), FIT_THREADS, 0, pf_stream(stream)>>>(P, W, K, noise, ctx, valid, gs, gV, gc, N, Kcap, gW_out, gP_inout);
    }
    PF_LAUNCH_CHECK();
    return 0;
}
