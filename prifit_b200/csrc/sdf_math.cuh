// Approximate ellipsoid SDF of the reference (convex_loss.py:313-328) and its local gradient, shared by the SDF-loss
// kernels (sdf.cu) and the intersection-loss kernels (intersect.cu).
//   z = V^T (p - c) ; k0 = |z / (s + 1e-6)| ; k1 = |z / (s^2 + 1e-6)| ; sdf = k0 (k0 - 1) / (k1 + 1e-6)
#pragma once
#include "common.cuh"

constexpr int SDF_THREADS = 256;
constexpr int SDF_MAXK = 64;

static __device__ __forceinline__ float sdf_eval(const float* __restrict__ prm, float px, float py, float pz) {
    // prm: s[3], V[9] (row-major, columns = axes), c[3]
    const float dx = px - prm[12], dy = py - prm[13], dz = pz - prm[14];
    float k0 = 0.f, k1 = 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float z = prm[3 + a] * dx + prm[6 + a] * dy + prm[9 + a] * dz;
        const float s = prm[a];
        const float v = z / (s + 1e-6f), u = z / (s * s + 1e-6f);
        k0 += v * v; k1 += u * u;
    }
    k0 = sqrtf(k0); k1 = sqrtf(k1);
    return k0 * (k0 - 1.0f) / (k1 + 1e-6f);
}

// per-point local gradient: returns d sdf-loss / d(z, s) pieces for the arg-min ellipsoid
struct SdfGrad { float dz[3]; float ds[3]; float dvec[3]; };

static __device__ __forceinline__ void sdf_point_grad(const float* __restrict__ prm, float px, float py, float pz, float gsdf,
                                               float (&dV)[9], float (&dsv)[3], float (&dpt)[3]) {
    const float d[3] = {px - prm[12], py - prm[13], pz - prm[14]};
    float z[3], Aa[3], Ba[3], k0 = 0.f, k1 = 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        z[a] = prm[3 + a] * d[0] + prm[6 + a] * d[1] + prm[9 + a] * d[2];
        Aa[a] = prm[a] + 1e-6f; Ba[a] = prm[a] * prm[a] + 1e-6f;
        const float v = z[a] / Aa[a], u = z[a] / Ba[a];
        k0 += v * v; k1 += u * u;
    }
    k0 = sqrtf(k0); k1 = sqrtf(k1);
    const float den = k1 + 1e-6f;
    const float gk0 = gsdf * (2.0f * k0 - 1.0f) / den;
    const float gk1 = -gsdf * k0 * (k0 - 1.0f) / (den * den);
    float dz[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float v = z[a] / Aa[a], u = z[a] / Ba[a];
        const float gv = k0 > 0.f ? gk0 * v / k0 : 0.f;
        const float gu = k1 > 0.f ? gk1 * u / k1 : 0.f;
        dz[a] = gv / Aa[a] + gu / Ba[a];
        dsv[a] = -gv * z[a] / (Aa[a] * Aa[a]) - gu * z[a] / (Ba[a] * Ba[a]) * 2.0f * prm[a];
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int a = 0; a < 3; ++a) dV[3 * i + a] = d[i] * dz[a];
        dpt[i] = prm[3 + 3 * i] * dz[0] + prm[4 + 3 * i] * dz[1] + prm[5 + 3 * i] * dz[2];
    }
}

