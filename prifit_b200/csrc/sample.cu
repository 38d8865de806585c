// f2 -- surface sampling of the fitted ellipsoids on the device (reference: src/ellipsoid_utils.py:76-130 +
// src/sample_ellipsoid.py:17-63, trimesh on the CPU with one host round trip per ellipsoid).
//   counts : per shape, round(total * area_k / sum area) points per ellipsoid (np.round = half to even), 100 where that is
//            <= 0 (:104-107); area = 4 * 3.142 * ((ab)^p + (bc)^p + (ca)^p)^(1/p), p = 1.585 (:157-159)
//   sample : points uniformly distributed over each ellipsoid's surface, returned as the (U, V) parameters the reference
//            extracts from its mesh samples (sample_ellipsoid.py:45-46); the differentiable map (U, V, a, b, c, V, centre)
//            -> points (:50-53, :56-63) stays with the caller (plain tensor expressions).
// The reference samples a subdivided icosphere mesh "evenly" (blue noise) with NumPy's generator; this sampler draws i.i.d.
// uniform surface points from Philox by rejection from the sphere (acceptance = local area stretch of the map sphere ->
// ellipsoid).  Same distribution, different stream: parity is statistical, and documented as such.
#include <curand_kernel.h>
#include "common.cuh"

namespace {

__global__ void sample_counts_kernel(const float* __restrict__ s, const uint8_t* __restrict__ valid, const int32_t* __restrict__ K,
                                     int Kcap, int total_points, int min_points, int32_t* __restrict__ counts,
                                     int32_t* __restrict__ offsets) {
    const int b = blockIdx.x;
    if (threadIdx.x != 0) return;
    const int Kb = min(max(K[b], 0), Kcap);
    double sum = 0.0;
    for (int k = 0; k < Kb; ++k) {
        if (!valid[(size_t)b * Kcap + k]) continue;
        const float* q = s + ((size_t)b * Kcap + k) * 3;
        const float p = 1.585f;
        const float area = 4.0f * 3.142f * powf(powf(q[0] * q[1], p) + powf(q[1] * q[2], p) + powf(q[2] * q[0], p), 1.0f / p);
        sum += (double)area;
    }
    int off = 0;
    for (int k = 0; k < Kcap; ++k) {
        int n = 0;
        if (k < Kb && valid[(size_t)b * Kcap + k]) {
            const float* q = s + ((size_t)b * Kcap + k) * 3;
            const float p = 1.585f;
            const float area = 4.0f * 3.142f * powf(powf(q[0] * q[1], p) + powf(q[1] * q[2], p) + powf(q[2] * q[0], p), 1.0f / p);
            n = (int)rint((double)total_points * ((double)area / sum));        // np.round: half to even
            if (n <= 0 || !isfinite(area)) n = min_points;
        }
        counts[(size_t)b * Kcap + k] = n;
        offsets[(size_t)b * (Kcap + 1) + k] = off;
        off += n;
    }
    offsets[(size_t)b * (Kcap + 1) + Kcap] = off;
}

__global__ void __launch_bounds__(256) sample_surface_kernel(const float* __restrict__ s, const int32_t* __restrict__ offsets,
                                                             int Kcap, int Smax, unsigned long long seed,
                                                             float* __restrict__ U, float* __restrict__ Vang, int32_t* __restrict__ owner) {
    const int b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Smax) return;
    const int32_t* off = offsets + (size_t)b * (Kcap + 1);
    const size_t o = (size_t)b * Smax + i;
    if (i >= off[Kcap]) { U[o] = 0.f; Vang[o] = 0.f; owner[o] = -1; return; }
    int k = 0;
    while (k + 1 < Kcap && off[k + 1] <= i) ++k;                             // Kcap <= 64: linear scan
    const float* q = s + ((size_t)b * Kcap + k) * 3;
    const float a = q[0], bb = q[1], c = q[2];
    const float gbc = bb * c, gac = a * c, gab = a * bb;
    const float gmax = fmaxf(gbc, fmaxf(gac, gab));
    curandStatePhilox4_32_10_t st;
    curand_init(seed, (unsigned long long)o, 0, &st);
    float x = 0.f, y = 0.f, z = 1.f;
    for (int attempt = 0; attempt < 64; ++attempt) {
        const float4 r = curand_uniform4(&st);                               // (0, 1]
        z = 1.0f - 2.0f * r.x;
        const float rad = sqrtf(fmaxf(1.0f - z * z, 0.f));
        float sn, cs;
        sincospif(2.0f * r.y, &sn, &cs);
        x = rad * cs; y = rad * sn;
        // area stretch of (x, y, z) on the unit sphere -> (a x, b y, c z) on the ellipsoid
        const float g = sqrtf((gbc * x) * (gbc * x) + (gac * y) * (gac * y) + (gab * z) * (gab * z));
        if (r.z * gmax <= g) break;
    }
    const float px = a * x, py = bb * y, pz = c * z;
    Vang[o] = acosf(fminf(fmaxf(pz / (c + 1e-6f), -1.0f), 1.0f));            // guard_acos(points[:, 2] / (c + 1e-6))
    U[o] = atan2f(py / (bb + 1e-6f), px / (a + 1e-6f));
    owner[o] = k;
}

// pts = (a cos U sin V, b sin U sin V, c cos V) . Vmat^T + centre      (src/sample_ellipsoid.py:50-53, 56-63)
__global__ void __launch_bounds__(256) surface_points_fwd_kernel(
    const float* __restrict__ s, const float* __restrict__ Vm, const float* __restrict__ c, const float* __restrict__ U,
    const float* __restrict__ Vang, const int32_t* __restrict__ owner, int Kcap, int Smax, float* __restrict__ pts) {
    const int b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Smax) return;
    const size_t o = (size_t)b * Smax + i;
    const int k = owner[o];
    float* p = pts + o * 3;
    if (k < 0) { p[0] = p[1] = p[2] = 0.f; return; }
    const size_t bk = (size_t)b * Kcap + k;
    float su, cu, sv, cv;
    sincosf(U[o], &su, &cu);
    sincosf(Vang[o], &sv, &cv);
    const float l0 = s[bk * 3] * cu * sv, l1 = s[bk * 3 + 1] * su * sv, l2 = s[bk * 3 + 2] * cv;
    const float* R = Vm + bk * 9;
#pragma unroll
    for (int j = 0; j < 3; ++j) p[j] = (l0 * R[3 * j] + l1 * R[3 * j + 1] + l2 * R[3 * j + 2]) + c[bk * 3 + j];
}

// one CTA per (shape, ellipsoid): its points are the contiguous slots [offsets[k], offsets[k + 1]) -> deterministic reduction
__global__ void __launch_bounds__(256) surface_points_bwd_kernel(
    const float* __restrict__ s, const float* __restrict__ Vm, const float* __restrict__ U, const float* __restrict__ Vang,
    const int32_t* __restrict__ offsets, const float* __restrict__ gpts, int Kcap, int Smax,
    float* __restrict__ gs, float* __restrict__ gV, float* __restrict__ gc) {
    __shared__ float red[15 * 32];
    const int k = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const size_t bk = (size_t)b * Kcap + k;
    const int lo = offsets[(size_t)b * (Kcap + 1) + k], hi = min(offsets[(size_t)b * (Kcap + 1) + k + 1], Smax);
    float acc[15];
#pragma unroll
    for (int q = 0; q < 15; ++q) acc[q] = 0.f;
    const float a = s[bk * 3], bb = s[bk * 3 + 1], cc = s[bk * 3 + 2];
    float R[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) R[q] = Vm[bk * 9 + q];
    for (int i = lo + tid; i < hi; i += 256) {
        const size_t o = (size_t)b * Smax + i;
        float su, cu, sv, cv;
        sincosf(U[o], &su, &cu);
        sincosf(Vang[o], &sv, &cv);
        const float e0 = cu * sv, e1 = su * sv, e2 = cv;                    // d local / d (a, b, c)
        const float l[3] = {a * e0, bb * e1, cc * e2};
        const float g[3] = {gpts[o * 3], gpts[o * 3 + 1], gpts[o * 3 + 2]};
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            acc[12 + j] += g[j];                                            // centre
#pragma unroll
            for (int q = 0; q < 3; ++q) acc[3 + 3 * j + q] += g[j] * l[q];  // V[j][q]
        }
        acc[0] += (g[0] * R[0] + g[1] * R[3] + g[2] * R[6]) * e0;
        acc[1] += (g[0] * R[1] + g[1] * R[4] + g[2] * R[7]) * e1;
        acc[2] += (g[0] * R[2] + g[1] * R[5] + g[2] * R[8]) * e2;
    }
    block_sum<15>(acc, red);
    if (tid < 3) { gs[bk * 3 + tid] = acc[tid]; gc[bk * 3 + tid] = acc[12 + tid]; }
    if (tid < 9) gV[bk * 9 + tid] = acc[3 + tid];
}

}  // namespace

extern "C" int prifit_surface_points_fwd(const float* s, const float* V, const float* c, const float* U, const float* Vang,
                                         const int32_t* owner, int B, int Kcap, int Smax, float* pts_out, void* stream) {
    PF_CHECK_ARG(s && V && c && U && Vang && owner && pts_out, PRIFIT_E_BADARG, "null pointer");
    PF_CHECK_ARG(B > 0 && Kcap > 0 && Smax > 0, PRIFIT_E_BADARG, "bad sizes");
    surface_points_fwd_kernel<<<dim3((Smax + 255) / 256, B), 256, 0, pf_stream(stream)>>>(s, V, c, U, Vang, owner, Kcap, Smax, pts_out);
    PF_LAUNCH_CHECK();
    return 0;
}

extern "C" int prifit_surface_points_bwd(const float* s, const float* V, const float* U, const float* Vang, const int32_t* offsets,
                                         const float* gpts, int B, int Kcap, int Smax, float* gs_out, float* gV_out, float* gc_out,
                                         void* stream) {
    PF_CHECK_ARG(s && V && U && Vang && offsets && gpts && gs_out && gV_out && gc_out, PRIFIT_E_BADARG, "null pointer");
    PF_CHECK_ARG(B > 0 && Kcap > 0 && Smax > 0, PRIFIT_E_BADARG, "bad sizes");
    surface_points_bwd_kernel<<<dim3(Kcap, B), 256, 0, pf_stream(stream)>>>(s, V, U, Vang, offsets, gpts, Kcap, Smax, gs_out, gV_out, gc_out);
    PF_LAUNCH_CHECK();
    return 0;
}

extern "C" int prifit_sample_counts(const float* s, const uint8_t* valid, const int32_t* K, int B, int Kcap, int total_points,
                                    int min_points, int32_t* counts_out, int32_t* offsets_out, void* stream) {
    PF_CHECK_ARG(s && valid && K && counts_out && offsets_out, PRIFIT_E_BADARG, "null pointer");
    PF_CHECK_ARG(B > 0 && Kcap > 0 && Kcap <= 64 && total_points > 0 && min_points > 0, PRIFIT_E_BADARG, "bad sizes");
    sample_counts_kernel<<<B, 32, 0, pf_stream(stream)>>>(s, valid, K, Kcap, total_points, min_points, counts_out, offsets_out);
    PF_LAUNCH_CHECK();
    return 0;
}

extern "C" int prifit_sample_surface(const float* s, const int32_t* offsets, int B, int Kcap, int Smax, uint64_t seed,
                                     float* U_out, float* V_out, int32_t* owner_out, void* stream) {
    PF_CHECK_ARG(s && offsets && U_out && V_out && owner_out, PRIFIT_E_BADARG, "null pointer");
    PF_CHECK_ARG(B > 0 && Kcap > 0 && Kcap <= 64 && Smax > 0, PRIFIT_E_BADARG, "bad sizes");
    sample_surface_kernel<<<dim3((Smax + 255) / 256, B), 256, 0, pf_stream(stream)>>>(s, offsets, Kcap, Smax, (unsigned long long)seed, U_out, V_out, owner_out);
    PF_LAUNCH_CHECK();
    return 0;
}
