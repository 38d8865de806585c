// k2 rows -- fp32 mean-shift trajectories of the K selected seeds, forward and backward.
//
// reference: center = new_X[indices] (src/mean_shift.py:46) is the only place the mean-shift
// iterations (src/mean_shift.py:50-84) feed the differentiable graph, and every seed row evolves
// independently of the others given X.  So the gradient of the whole N-row dense autograd graph
// equals the gradient of the K selected rows.  These kernels recompute those K rows in full fp32
// (they are the `center` the membership / fit stages consume) and back-propagate through them.
//
// Parallelisation: K <= 64 rows is far too little row parallelism for 148 SMs, so the KEYS are
// split over a thread-block cluster (up to 8 CTAs per shape).  Every CTA owns a contiguous key
// slice, holds all 32 rows of a row group, and the per-iteration partial sums (32 x d numerators
// + 32 denominators) are reduced through distributed shared memory in fixed rank order.
#include "rowgemm.cuh"

namespace {

__host__ __device__ inline int slice_len(int N, int csize) { return (N + csize - 1) / csize; }

template <int D>
struct RowsSmem {
    static constexpr int LD = D + 4;
    static constexpr size_t fwd_floats = (size_t)RG_ROWS * LD        // ys
                                       + (size_t)RG_KEYS * LD        // xs
                                       + (size_t)RG_ROWS * RG_LDP    // ps   (part_o and urow alias xs)
                                       + RG_ROWS                     // part_z
                                       + RG_THREADS;                 // red
    static constexpr size_t bwd_floats = (size_t)3 * RG_ROWS * LD    // ys, gms, gy
                                       + (size_t)RG_KEYS * LD        // xs
                                       + (size_t)2 * RG_ROWS * RG_LDP  // ps1, ps2
                                       + (size_t)RG_ROWS * D         // part_o
                                       + 2 * RG_ROWS;                // gmm, dinv
};

// ------------------------------------------------------------------------------------------ forward
template <int D>
__global__ void __launch_bounds__(RG_THREADS, 2) ms_rows_fwd_kernel(
    const float* __restrict__ X, const float* __restrict__ bw, const int32_t* __restrict__ idx,
    const int32_t* __restrict__ K, int N, int T, int Kcap,
    float* __restrict__ traj, float* __restrict__ stat, float* __restrict__ C_out) {
    constexpr int LD = D + 4;
    constexpr int NH = (D + 127) / 128;
    cg::cluster_group cluster = cg::this_cluster();
    const int csize = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    const int b = blockIdx.z, k0 = blockIdx.y * RG_ROWS;
    const int Kb = min(K[b], Kcap);
    const int nrows = max(0, min(RG_ROWS, Kb - k0));
    const int krows = min(RG_ROWS, Kcap - k0);       // rows of this group that exist in the padded layout
    const int tid = threadIdx.x, ty = tid >> 5, tx = tid & 31;

    float* traj_b = traj + (size_t)b * (T + 1) * Kcap * D;
    float* stat_b = stat + (size_t)b * T * Kcap * 2;

    // padded rows of this row group are defined as zero
    for (int t = rank; t <= T; t += csize) {
        for (int e = tid; e < (krows - nrows) * D; e += RG_THREADS) {
            const int r = nrows + e / D, c = e % D;
            traj_b[((size_t)t * Kcap + k0 + r) * D + c] = 0.f;
            if (t == T) C_out[((size_t)b * Kcap + k0 + r) * D + c] = 0.f;
        }
        if (t < T)
            for (int e = tid; e < (krows - nrows) * 2; e += RG_THREADS)
                stat_b[((size_t)t * Kcap + k0 + nrows) * 2 + e] = 0.f;
    }
    if (nrows == 0) return;   // uniform over the cluster

    extern __shared__ __align__(16) float smem[];
    float* ys = smem;
    float* xs = ys + RG_ROWS * LD;
    float* ps = xs + RG_KEYS * LD;
    float* part_o = xs;                       // the key tile is dead once the tile loop is done (2 CTAs / SM)
    float* urow = xs + RG_ROWS * D;
    float* part_z = ps + RG_ROWS * RG_LDP;
    float* red = part_z + RG_ROWS;

    const float* Xb = X + (size_t)b * N * D;
    const int32_t* idx_b = idx + (size_t)b * Kcap + k0;
    const float bwv = bw[b];
    const float b2 = bwv * bwv;
    const int sl = slice_len(N, csize);
    const int jbeg = rank * sl, jend = min(N, jbeg + sl);

    rg_load_rows<D>(ys, RG_ROWS, Xb, [&](int r) -> long long { return r < nrows ? (long long)idx_b[r] : -1; });
    __syncthreads();
    if (rank == 0)
        for (int e = tid; e < nrows * D; e += RG_THREADS) {
            const int r = e / D, c = e % D;
            traj_b[((size_t)k0 + r) * D + c] = ys[r * LD + c];
        }

    const int rpc = RG_ROWS / csize, tpr = RG_THREADS / rpc;
    const int rr = tid / tpr, tc = tid - rr * tpr;
    const int myrow = rank * rpc + rr;

    for (int t = 0; t < T; ++t) {
        float z[4] = {0.f, 0.f, 0.f, 0.f};
        float o[4][4 * NH];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 4 * NH; ++c) o[a][c] = 0.f;

        for (int j0 = jbeg; j0 < jend; j0 += RG_KEYS) {
            __syncthreads();
            rg_load_rows<D>(xs, RG_KEYS, Xb, [&](int r) -> long long { return j0 + r < jend ? (long long)(j0 + r) : -1; });
            __syncthreads();
            float acc[4][4];
            rg_dot_32x128<D>(ys, xs, acc);
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                float part = 0.f;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float dist = 2.0f - 2.0f * acc[a][c];                      // src/mean_shift.py:65
                    float p = guard_expf((-dist / b2) * 0.5f);                       // :68
                    if (j0 + tx + 32 * c >= jend) p = 0.f;
                    ps[(ty + 8 * a) * RG_LDP + tx + 32 * c] = p;
                    part += p;
                }
                z[a] += warp_sum(part);
            }
            __syncthreads();
            rg_accum_rows<D>(ps, xs, o);
        }
        __syncthreads();                              // every warp is done with xs before part_o overwrites it
#pragma unroll
        for (int a = 0; a < 4; ++a) {
#pragma unroll
            for (int h = 0; h < NH; ++h)
                if (4 * tx + 128 * h < D)
                    *reinterpret_cast<float4*>(part_o + (ty + 8 * a) * D + 4 * tx + 128 * h) =
                        make_float4(o[a][4 * h], o[a][4 * h + 1], o[a][4 * h + 2], o[a][4 * h + 3]);
            if (tx == 0) part_z[ty + 8 * a] = z[a];
        }
        cluster.sync();
        rg_cluster_reduce_rows<D>(cluster, part_o, csize, [&](int row, int col, float s) { urow[(row - rank * rpc) * D + col] = s; });
        float zsum = 0.f;
        for (int q = 0; q < csize; ++q) zsum += cluster.map_shared_rank(part_z, q)[myrow];
        // new = y + ((K X) D - y);  new /= ||new||      (src/mean_shift.py:75-82)
        const float dinv = 1.0f / zsum;
        float n2 = 0.f;
        for (int col = tc; col < D; col += tpr) {
            const float y = ys[myrow * LD + col];
            const float m = urow[rr * D + col] * dinv - y;
            const float u = y + m;
            urow[rr * D + col] = u;
            n2 = fmaf(u, u, n2);
        }
        red[tid] = n2;
        __syncthreads();
        float nrm = 0.f;
        for (int q = 0; q < tpr; ++q) nrm += red[rr * tpr + q];
        nrm = sqrtf(nrm);
        const bool live = myrow < nrows;
        for (int col = tc; col < D; col += tpr) {
            const float ynew = urow[rr * D + col] / nrm;
            for (int q = 0; q < csize; ++q) cluster.map_shared_rank(ys, q)[myrow * LD + col] = ynew;
            if (live) {
                traj_b[((size_t)(t + 1) * Kcap + k0 + myrow) * D + col] = ynew;
                if (t == T - 1) C_out[((size_t)b * Kcap + k0 + myrow) * D + col] = ynew;
            }
        }
        if (live && tc == 0) {
            stat_b[((size_t)t * Kcap + k0 + myrow) * 2 + 0] = zsum;
            stat_b[((size_t)t * Kcap + k0 + myrow) * 2 + 1] = nrm;
        }
        cluster.sync();
    }
    if (T == 0 && rank == 0)
        for (int e = tid; e < nrows * D; e += RG_THREADS) {
            const int r = e / D, c = e % D;
            C_out[((size_t)b * Kcap + k0 + r) * D + c] = ys[r * LD + c];
        }
}

// ----------------------------------------------------------------------------------------- backward
// One reverse step t, for seed row r with g = dL/dy^{t+1}  (autograd of src/mean_shift.py:65-82):
//   g_u = (g - (g.y^{t+1}) y^{t+1}) / ||u||          (normalisation)
//   g_m = g_u                                          (the +y / -y paths cancel exactly)
//   dL/dkappa_j = (g_m.x_j - g_m.m) D                  m = u = y^{t+1} ||u||,  D = 1/Z
//   dL/da_j     = kappa_j dL/dkappa_j [lo <= a_j <= hi]      (guard_exp clamp)
//   dL/ds_j     = dL/da_j / b^2
//   dL/dy^t     = sum_j dL/ds_j x_j ;   dL/dx_j += dL/ds_j y^t + kappa_j D g_m
template <int D>
__global__ void __launch_bounds__(RG_THREADS) ms_rows_bwd_kernel(
    const float* __restrict__ X, const float* __restrict__ bw, const int32_t* __restrict__ idx,
    const int32_t* __restrict__ K, const float* __restrict__ traj, const float* __restrict__ stat,
    const float* __restrict__ gC, int N, int T, int Kcap, float* __restrict__ gX) {
    constexpr int LD = D + 4;
    constexpr int NH = (D + 127) / 128;
    cg::cluster_group cluster = cg::this_cluster();
    const int csize = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    const int b = blockIdx.z;
    const int Kb = min(K[b], Kcap);
    if (Kb <= 0) return;      // uniform over the cluster
    const int tid = threadIdx.x, ty = tid >> 5, tx = tid & 31;

    extern __shared__ __align__(16) float smem[];
    float* ys = smem;
    float* gms = ys + RG_ROWS * LD;
    float* gy = gms + RG_ROWS * LD;
    float* xs = gy + RG_ROWS * LD;
    float* ps1 = xs + RG_KEYS * LD;
    float* ps2 = ps1 + RG_ROWS * RG_LDP;
    float* part_o = ps2 + RG_ROWS * RG_LDP;
    float* gmm_s = part_o + RG_ROWS * D;
    float* dinv_s = gmm_s + RG_ROWS;

    const float* Xb = X + (size_t)b * N * D;
    float* gXb = gX + (size_t)b * N * D;
    const float* traj_b = traj + (size_t)b * (T + 1) * Kcap * D;
    const float* stat_b = stat + (size_t)b * T * Kcap * 2;
    const float bwv = bw[b];
    const float b2 = bwv * bwv;
    const int sl = slice_len(N, csize);
    const int jbeg = rank * sl, jend = min(N, jbeg + sl);
    const int rpc = RG_ROWS / csize, tpr = RG_THREADS / rpc;
    const int rr = tid / tpr, tc = tid - rr * tpr;
    const int myrow = rank * rpc + rr;

    for (int k0 = 0; k0 < Kb; k0 += RG_ROWS) {
        const int nrows = min(RG_ROWS, Kb - k0);
        const float* gC_b = gC + ((size_t)b * Kcap + k0) * D;
        rg_load_rows<D>(gy, RG_ROWS, gC_b, [&](int r) -> long long { return r < nrows ? (long long)r : -1; });
        __syncthreads();

        for (int t = T - 1; t >= 0; --t) {
            // per-row preparation, one warp per row (rows ty + 8a)
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const int r = ty + 8 * a;
                if (r < nrows) {
                    const float* yn_g = traj_b + ((size_t)(t + 1) * Kcap + k0 + r) * D;
                    const float* yt_g = traj_b + ((size_t)t * Kcap + k0 + r) * D;
                    const float zt = stat_b[((size_t)t * Kcap + k0 + r) * 2 + 0];
                    const float nrm = stat_b[((size_t)t * Kcap + k0 + r) * 2 + 1];
                    float gdot = 0.f;
                    for (int c = tx; c < D / 4; c += 32) {
                        const float4 g = *reinterpret_cast<const float4*>(gy + r * LD + 4 * c);
                        const float4 yn = reinterpret_cast<const float4*>(yn_g)[c];
                        gdot += g.x * yn.x + g.y * yn.y + g.z * yn.z + g.w * yn.w;
                    }
                    gdot = warp_sum(gdot);
                    float gmm = 0.f;
                    for (int c = tx; c < D / 4; c += 32) {
                        const float4 g = *reinterpret_cast<const float4*>(gy + r * LD + 4 * c);
                        const float4 yn = reinterpret_cast<const float4*>(yn_g)[c];
                        float4 gm;
                        gm.x = (g.x - gdot * yn.x) / nrm; gm.y = (g.y - gdot * yn.y) / nrm;
                        gm.z = (g.z - gdot * yn.z) / nrm; gm.w = (g.w - gdot * yn.w) / nrm;
                        *reinterpret_cast<float4*>(gms + r * LD + 4 * c) = gm;
                        gmm += gm.x * (yn.x * nrm) + gm.y * (yn.y * nrm) + gm.z * (yn.z * nrm) + gm.w * (yn.w * nrm);
                        *reinterpret_cast<float4*>(ys + r * LD + 4 * c) = reinterpret_cast<const float4*>(yt_g)[c];
                    }
                    gmm = warp_sum(gmm);
                    if (tx == 0) { gmm_s[r] = gmm; dinv_s[r] = 1.0f / zt; }
                } else {
                    for (int c = tx; c < D / 4; c += 32) {
                        *reinterpret_cast<float4*>(gms + r * LD + 4 * c) = make_float4(0.f, 0.f, 0.f, 0.f);
                        *reinterpret_cast<float4*>(ys + r * LD + 4 * c) = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    if (tx == 0) { gmm_s[r] = 0.f; dinv_s[r] = 0.f; }
                }
            }

            float o[4][4 * NH];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int c = 0; c < 4 * NH; ++c) o[a][c] = 0.f;

            for (int j0 = jbeg; j0 < jend; j0 += RG_KEYS) {
                __syncthreads();
                rg_load_rows<D>(xs, RG_KEYS, Xb, [&](int r) -> long long { return j0 + r < jend ? (long long)(j0 + r) : -1; });
                __syncthreads();
                float s1[4][4], s2[4][4];
                rg_dot2_32x128<D>(ys, gms, xs, s1, s2);
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    const int r = ty + 8 * a;
                    const float gmm = gmm_s[r], dinv = dinv_s[r];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float dist = 2.0f - 2.0f * s1[a][c];
                        const float av = (-dist / b2) * 0.5f;
                        const float kap = guard_expf(av);
                        const bool inr = (av >= PRIFIT_LO) && (av <= PRIFIT_HI);
                        const float dk = (s2[a][c] - gmm) * dinv;
                        float ds = inr ? (kap * dk) / b2 : 0.f;
                        float e1 = kap * dinv;
                        if (j0 + tx + 32 * c >= jend) { ds = 0.f; e1 = 0.f; }
                        ps1[r * RG_LDP + tx + 32 * c] = ds;
                        ps2[r * RG_LDP + tx + 32 * c] = e1;
                    }
                }
                __syncthreads();
                rg_accum_rows<D>(ps1, xs, o);
                rg_accum_keys<D, true>(ps1, ys, ps2, gms, gXb + (size_t)j0 * D, min(RG_KEYS, jend - j0));
            }
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int h = 0; h < NH; ++h)
                    if (4 * tx + 128 * h < D)
                        *reinterpret_cast<float4*>(part_o + (ty + 8 * a) * D + 4 * tx + 128 * h) =
                            make_float4(o[a][4 * h], o[a][4 * h + 1], o[a][4 * h + 2], o[a][4 * h + 3]);
            cluster.sync();
            rg_cluster_reduce_rows<D>(cluster, part_o, csize, [&](int row, int col, float s) {
                for (int q = 0; q < csize; ++q) cluster.map_shared_rank(gy, q)[row * LD + col] = s;
            });
            cluster.sync();
        }
        // dL/dy^0 lands on the seed's own row of X (new_X = X.clone(), gather by idx)
        if (myrow < nrows) {
            const int src = idx[(size_t)b * Kcap + k0 + myrow];
            for (int col = tc; col < D; col += tpr) atomicAdd(gXb + (size_t)src * D + col, gy[myrow * LD + col]);
        }
        __threadfence();
        cluster.sync();
    }
}

template <int D>
int launch_rows_fwd(const float* X, const float* bw, const int32_t* idx, const int32_t* K, int B, int N, int T,
                    int Kcap, float* traj, float* stat, float* C_out, cudaStream_t st) {
    const size_t smem = RowsSmem<D>::fwd_floats * sizeof(float);
    PF_CUDA(cudaFuncSetAttribute(ms_rows_fwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int csize = 8;
    while (csize > 1 && (N + RG_KEYS - 1) / RG_KEYS < csize) csize >>= 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(csize, (Kcap + RG_ROWS - 1) / RG_ROWS, B);
    cfg.blockDim = dim3(RG_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    PF_CUDA(cudaLaunchKernelEx(&cfg, ms_rows_fwd_kernel<D>, X, bw, idx, K, N, T, Kcap, traj, stat, C_out));
    return 0;
}

template <int D>
int launch_rows_bwd(const float* X, const float* bw, const int32_t* idx, const int32_t* K, const float* traj,
                    const float* stat, const float* gC, int B, int N, int T, int Kcap, float* gX, cudaStream_t st) {
    const size_t smem = RowsSmem<D>::bwd_floats * sizeof(float);
    PF_CUDA(cudaFuncSetAttribute(ms_rows_bwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int csize = 8;
    while (csize > 1 && (N + RG_KEYS - 1) / RG_KEYS < csize) csize >>= 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(csize, 1, B);
    cfg.blockDim = dim3(RG_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = csize; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    PF_CUDA(cudaLaunchKernelEx(&cfg, ms_rows_bwd_kernel<D>, X, bw, idx, K, traj, stat, gC, N, T, Kcap, gX));
    return 0;
}

}  // namespace

size_t prifit_rows_tc_workspace_bytes(int B, int N);
int prifit_rows_tc_fwd(const float* X, const float* bw, const int32_t* idx, const int32_t* K, int B, int N, int T, int Kcap,
                       float* traj, float* stat, float* C_out, void* ws, bool ws_holds_split, int wide, cudaStream_t st);
int prifit_rows_tc_prepare(const float* X, int B, int N, void* ws, cudaStream_t st);
int prifit_rows_tc_bwd(const float* X, const float* bw, const int32_t* idx, const int32_t* K, const float* traj,
                       const float* stat, const float* gC, int B, int N, int T, int Kcap, float* gX, void* ws, bool ws_holds_split,
                       int wide, cudaStream_t st);

extern "C" size_t prifit_meanshift_rows_workspace_bytes(int B, int N, int d, int engine) {
    engine &= 0xff;
    if (engine == PRIFIT_ROWS_SPLIT_TCGEN05 && d == 128) return prifit_rows_tc_workspace_bytes(B, N);
    return 16;
}

extern "C" int prifit_meanshift_rows_prepare(const float* X, int B, int N, int d, int engine, void* ws, size_t ws_bytes, void* stream) {
    PF_CHECK_ARG(X && B > 0 && N > 0, PRIFIT_E_BADARG, "X, B, N > 0 required");
    engine &= 0xff;
    if (engine != PRIFIT_ROWS_SPLIT_TCGEN05 || d != 128) return 0;          // the fp32 engine reads X as it is
    PF_CHECK_ARG(ws && ws_bytes >= prifit_meanshift_rows_workspace_bytes(B, N, d, engine), PRIFIT_E_WS, "workspace too small");
    return prifit_rows_tc_prepare(X, B, N, ws, pf_stream(stream));
}

extern "C" int prifit_meanshift_rows_fwd(const float* X, const float* bw, const int32_t* idx, const int32_t* K,
                                         int B, int N, int d, int T, int Kcap,
                                         float* traj_out, float* stat_out, float* C_out,
                                         int engine, void* ws, size_t ws_bytes, void* stream) {
    PF_CHECK_ARG(X && bw && idx && K && traj_out && C_out && (stat_out || T == 0), PRIFIT_E_BADARG, "null pointer");
    PF_CHECK_ARG(B > 0 && N > 0 && T >= 0, PRIFIT_E_BADARG, "B, N > 0 and T >= 0 required");
    PF_CHECK_ARG(Kcap > 0 && Kcap % 4 == 0 && Kcap <= 64, PRIFIT_E_SHAPE, "Kcap must be a multiple of 4, <= 64");
    const int wide = (engine & PRIFIT_ROWS_WIDE) ? 1 : (engine & PRIFIT_ROWS_NARROW) ? 0 : -1;
    const bool ws_holds_split = (engine & PRIFIT_ROWS_WS_HOLDS_SPLIT) != 0;
    engine &= 0xff;
    if (engine == PRIFIT_ROWS_SPLIT_TCGEN05) {
        PF_CHECK_ARG(d == 128, PRIFIT_E_SHAPE, "the tcgen05 engine is specialised for d == 128");
        PF_CHECK_ARG(ws && ws_bytes >= prifit_meanshift_rows_workspace_bytes(B, N, d, engine), PRIFIT_E_WS, "workspace too small");
        return prifit_rows_tc_fwd(X, bw, idx, K, B, N, T, Kcap, traj_out, stat_out, C_out, ws, ws_holds_split, wide, pf_stream(stream));
    }
    PF_CHECK_ARG(engine == PRIFIT_ROWS_FP32_SIMT, PRIFIT_E_BADARG, "unknown engine");
    switch (d) {
        case 64: return launch_rows_fwd<64>(X, bw, idx, K, B, N, T, Kcap, traj_out, stat_out, C_out, pf_stream(stream));
        case 128: return launch_rows_fwd<128>(X, bw, idx, K, B, N, T, Kcap, traj_out, stat_out, C_out, pf_stream(stream));
        case 256: return launch_rows_fwd<256>(X, bw, idx, K, B, N, T, Kcap, traj_out, stat_out, C_out, pf_stream(stream));
        default: prifit_set_error("prifit_meanshift_rows_fwd: d must be 64, 128 or 256 (got %d)", d); return PRIFIT_E_SHAPE;
    }
}

extern "C" int prifit_meanshift_rows_bwd(const float* X, const float* bw, const int32_t* idx, const int32_t* K,
                                         const float* traj, const float* stat, const float* gC,
                                         int B, int N, int d, int T, int Kcap, float* gX_inout,
                                         int engine, void* ws, size_t ws_bytes, void* stream) {
    PF_CHECK_ARG(X && bw && idx && K && traj && gC && gX_inout && (stat || T == 0), PRIFIT_E_BADARG, "null pointer");
    PF_CHECK_ARG(B > 0 && N > 0 && T >= 0, PRIFIT_E_BADARG, "B, N > 0 and T >= 0 required");
    PF_CHECK_ARG(Kcap > 0 && Kcap % 4 == 0 && Kcap <= 64, PRIFIT_E_SHAPE, "Kcap must be a multiple of 4, <= 64");
    const bool ws_holds_split = (engine & PRIFIT_ROWS_WS_HOLDS_SPLIT) != 0;
    const int wide = (engine & PRIFIT_ROWS_WIDE) ? 1 : (engine & PRIFIT_ROWS_NARROW) ? 0 : -1;
    engine &= 0xff;
    if (engine == PRIFIT_ROWS_SPLIT_TCGEN05) {
        PF_CHECK_ARG(d == 128, PRIFIT_E_SHAPE, "the tcgen05 engine is specialised for d == 128");
        PF_CHECK_ARG(ws && ws_bytes >= prifit_meanshift_rows_workspace_bytes(B, N, d, engine), PRIFIT_E_WS, "workspace too small");
        return prifit_rows_tc_bwd(X, bw, idx, K, traj, stat, gC, B, N, T, Kcap, gX_inout, ws, ws_holds_split, wide, pf_stream(stream));
    }
    PF_CHECK_ARG(engine == PRIFIT_ROWS_FP32_SIMT, PRIFIT_E_BADARG, "unknown engine");
    switch (d) {
        case 64: return launch_rows_bwd<64>(X, bw, idx, K, traj, stat, gC, B, N, T, Kcap, gX_inout, pf_stream(stream));
        case 128: return launch_rows_bwd<128>(X, bw, idx, K, traj, stat, gC, B, N, T, Kcap, gX_inout, pf_stream(stream));
        case 256: return launch_rows_bwd<256>(X, bw, idx, K, traj, stat, gC, B, N, T, Kcap, gX_inout, pf_stream(stream));
        default: prifit_set_error("prifit_meanshift_rows_bwd: d must be 64, 128 or 256 (got %d)", d); return PRIFIT_E_SHAPE;
    }
}
