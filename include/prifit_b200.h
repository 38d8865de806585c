/*
 * prifit_b200 -- C ABI of the B200-native mean-shift + ellipsoid-fit path.
 *
 * The reference (Hippogriff/prifit) is pure Python/eager-torch and has no FFI of its own; the
 * boundary it exposes is the Python call surface of src/mean_shift.py, src/ellipsoid_utils.py,
 * src/ellipsoid_fitting.py, src/fitting_utils.py and convex_loss.py.  `prifit_b200/*.py` keeps that
 * surface and calls the entry points below through ctypes.  Each entry point cites the reference
 * code it replaces (paths relative to the reference root).
 *
 * Common rules
 *   - extern "C"; every function returns int: 0 = ok, <0 = bad argument (PRIFIT_E_*),
 *     >0 = cudaError_t of the failed runtime call.  prifit_last_error_string() describes it.
 *   - never throws, never allocates device memory, never synchronises the host; work is enqueued
 *     on `stream` (a cudaStream_t passed as void*).  Scratch memory is caller-provided
 *     (`ws`, sized by the matching *_workspace_bytes()).  Re-entrant across streams as long as
 *     workspaces are distinct.
 *   - all pointers are DEVICE pointers to contiguous row-major fp32 / int32 / uint8 arrays.
 *   - padded layouts: per-shape cluster lists are [B, Kcap, ...] with int32 K[B] valid entries.
 *   - B shapes, N points per shape, d embedding width (multiple of 4; tcgen05 path: d == 128),
 *     T mean-shift iterations, Kcap cluster capacity (multiple of 4, <= 64).
 */
#ifndef PRIFIT_B200_H
#define PRIFIT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PRIFIT_VERSION 100

#define PRIFIT_E_BADARG   (-1)   /* null pointer / non-positive size */
#define PRIFIT_E_SHAPE    (-2)   /* unsupported shape (d, Kcap, N limits) */
#define PRIFIT_E_WS       (-3)   /* workspace too small */
#define PRIFIT_E_NODEVICE (-4)   /* no sm_100 device / driver entry point missing */

/* mean-shift forward engines */
#define PRIFIT_MS_F16_TCGEN05 0 /* tensor cores: tcgen05.mma kind::f16 (10-bit mantissa operands like tf32), TMEM, TMA */
#define PRIFIT_MS_FP32_SIMT    1 /* CUDA-core fp32, used to cross-check the tensor-core kernel */

int prifit_version(void);
/* Engine of the Gram-matrix passes inside prifit_bandwidth_fwd and prifit_nms_fwd (process-global):
 * 0 = tensor cores (tcgen05, d == 128), 1 = fp32 CUDA cores (cross-check).  Returns the previous value. */
int prifit_set_gram_engine(int engine);
const char* prifit_last_error_string(void);
/* 1 if the current device is compute capability 10.x, else 0 (or <0 on error) */
int prifit_device_ok(void);

/* A0 -- X = normalize(normalize(E)) row-wise, eps 1e-12.  convex_loss.py:41,57 */
int prifit_normalize_fwd(const float* E, int64_t rows, int d, float* X, void* stream);
/* backward of the two stacked F.normalize nodes; gE may alias gX */
int prifit_normalize_bwd(const float* E, const float* gX, int64_t rows, int d, float* gE, void* stream);
/* channel-first variants (the layout convex_loss receives, convex_loss.py:27,37): Ecf[B,d,N] -> X[B,N,d] and
 *   (Ecf[B,d,N], gX[B,N,d]) -> gEcf[B,d,N]; d = 128; same arithmetic as the row-major pair (bit-identical X). */
int prifit_normalize_fwd_cf(const float* Ecf, int B, int N, int d, float* X, void* stream);
int prifit_normalize_bwd_cf(const float* Ecf, const float* gX, int B, int N, int d, float* gEcf, void* stream);
/* Backward of the two normalisations with an upstream scale read from device memory:
 *   gE = normalize_bwd(E, gX * (g_sum[0] + g_mean[0] / max(stats[1], 1)))        (g_sum / g_mean may be NULL = 0)
 * The gradient of the whole path is linear in dL/d(loss): the graph-replayed step computes d(sum_b has_b loss_b)/dX ahead
 * of time and this kernel applies dL/d(loss_sum), dL/d(loss_mean) (stats = [loss_sum, n_valid, loss_mean] of
 * prifit_masked_mean_fwd) when autograd delivers them -- the autograd edge of src/utils.py:425 + convex_loss.py:41,57.
 * E, gE: [B,N,d] (channel_first = 0) or [B,d,N] (channel_first = 1, d == 128); gX: [B,N,d]. */
int prifit_normalize_bwd_scaled(const float* E, const float* gX, int B, int N, int d, int channel_first,
                                const float* g_sum, const float* g_mean, const float* stats, float* gE, void* stream);

/* k1 -- bandwidth.  src/mean_shift.py:138-160 (compute_bandwidth)
 *   rows: optional [B, n_s] int32 subset (the first num_samples entries of the host shuffle,
 *   line 150); NULL = all N rows (n_s must equal N).  kth[B] = int(quantile * n_s) per shape
 *   (1-based rank of the order statistic).  bw_out[B] = mean_i sqrt(max(kth-smallest_i, 1e-6)). */
size_t prifit_bandwidth_workspace_bytes(int B, int N, int d, int n_s);
int prifit_bandwidth_fwd(const float* X, int B, int N, int d, const int32_t* rows, int n_s,
                         const int32_t* kth, float* bw_out, void* ws, size_t ws_bytes, void* stream);

/* k2 -- T mean-shift iterations of all N seeds.  src/mean_shift.py:50-84 (mean_shift_, gaussian)
 *   newX_out[B,N,d].  engine = PRIFIT_MS_*.  Never materialises the N x N kernel matrix. */
size_t prifit_meanshift_workspace_bytes(int B, int N, int d, int engine);
int prifit_meanshift_fwd(const float* X, const float* bw, int B, int N, int d, int T, float* newX_out,
                         int engine, void* ws, size_t ws_bytes, void* stream);

/* k3 -- NMS mode pruning + hard labels.  src/mean_shift.py:162-202 (nms(new_X, new_X, bw)) and the
 *   guard predicate of src/ellipsoid_utils.py:19-26.
 *   idx_out[B,Kcap]  representative point index per cluster, ascending, -1 padded
 *   K_out[B]         number of representatives (may exceed Kcap; only the first Kcap are stored in idx_out)
 *   labels_out[B,N]  argmax_k <newX[rep k], newX[j]> over ALL K_out[b] representatives (also beyond Kcap), lowest k on ties
 *   n_labels_out[B]  number of distinct labels (== torch.unique(labels).shape[0]), exact for any K: the guard criterion of
 *                    src/ellipsoid_utils.py:23.  A shape with K_out > Kcap that passes the guard needs a larger Kcap
 *                    for the differentiable stages (the host re-runs it with Kcap = 64). */
size_t prifit_nms_workspace_bytes(int B, int N, int d);
int prifit_nms_fwd(const float* newX, const float* bw, int B, int N, int d, int Kcap,
                   int32_t* idx_out, int32_t* K_out, int32_t* labels_out, int32_t* n_labels_out,
                   void* ws, size_t ws_bytes, void* stream);

/* step 5 of the NMS alone (src/mean_shift.py:200-201): prifit_nms_fwd may be called with labels_out = n_labels_out = NULL -- it
 *   then stops at the centres (idx_out, K_out) -- and this call, on the same workspace, computes the hard labels and their count
 *   afterwards, e.g. on another stream beside the consumers of the centres. */
int prifit_nms_labels(const float* newX, int B, int N, int d, int Kcap, const int32_t* K,
                      int32_t* idx_out, int32_t* labels_out, int32_t* n_labels_out,
                      void* ws, size_t ws_bytes, void* stream);

/* engines of the K-seed trajectory kernels */
#define PRIFIT_ROWS_SPLIT_TCGEN05 0 /* tensor cores, split-fp16 (hi + lo) operands, 3 tcgen05.mma per product: fp32-class; d == 128 */
#define PRIFIT_ROWS_FP32_SIMT     1 /* CUDA-core fp32 (cross-check; d in {64, 128, 256}) */
/* flag OR-ed into `engine` of prifit_meanshift_rows_fwd / _bwd: the workspace still holds the split fp16 rows of this X, left
 * there by prifit_meanshift_rows_prepare or by the prifit_meanshift_rows_fwd call with the same (X, B, N, ws) -- the call then
 * skips its own split pass */
#define PRIFIT_ROWS_WS_HOLDS_SPLIT 0x100
/* flags OR-ed into `engine` of prifit_meanshift_rows_fwd / _bwd (tcgen05 engine): split every shape's keys over 8 CTAs (WIDE) or
 * 4 (NARROW).  Bit-identical results (the key partial sums are formed per fixed unit of tiles, not per CTA): a scheduling choice.
 * 8 has 2/3 of the latency when the launch has its 8 B SMs to itself and loses when such launches compete (one CTA per SM).
 * Neither flag: 8 when 8 B x ceil(Kcap / 32) <= 74 CTAs, else 4. */
#define PRIFIT_ROWS_WIDE 0x200
#define PRIFIT_ROWS_NARROW 0x400

/* k2 rows -- fp32 trajectories of the K selected seeds (center = new_X[indices],
 *   src/mean_shift.py:46): traj_out[B, T+1, Kcap, d] (y^0..y^T), stat_out[B, T, Kcap, 2] =
 *   (Z_t, ||u_t||), C_out[B,Kcap,d] = y^T.  Rows k >= K[b] are zero.  engine = PRIFIT_ROWS_*. */
size_t prifit_meanshift_rows_workspace_bytes(int B, int N, int d, int engine);
/* optional: the operand preparation of the tensor-core engine (split fp16 rows of X into ws) ahead of time; X is all it needs */
int prifit_meanshift_rows_prepare(const float* X, int B, int N, int d, int engine, void* ws, size_t ws_bytes, void* stream);
int prifit_meanshift_rows_fwd(const float* X, const float* bw, const int32_t* idx, const int32_t* K,
                              int B, int N, int d, int T, int Kcap,
                              float* traj_out, float* stat_out, float* C_out,
                              int engine, void* ws, size_t ws_bytes, void* stream);
/* k2 backward -- autograd of mean_shift_ restricted to the K seeds that carry gradient (rows are
 *   independent given X).  gC[B,Kcap,d] = dL/d center; accumulates dL/dX into gX_inout[B,N,d].
 *   traj / stat may come from either forward engine. */
int prifit_meanshift_rows_bwd(const float* X, const float* bw, const int32_t* idx, const int32_t* K,
                              const float* traj, const float* stat, const float* gC,
                              int B, int N, int d, int T, int Kcap, float* gX_inout,
                              int engine, void* ws, size_t ws_bytes, void* stream);

/* k4 -- soft membership.  src/mean_shift.py:230-247 (membership); W_out[B,Kcap,N] (cluster-major,
 *   i.e. the reference's [K,N] before the caller's transpose), smax_out[B] = the detached global max. */
size_t prifit_membership_workspace_bytes(int B, int N, int Kcap);
int prifit_membership_fwd(const float* C, const float* X, const float* bw, const int32_t* K,
                          int B, int N, int d, int Kcap, float* W_out, float* smax_out,
                          void* ws, size_t ws_bytes, void* stream);
/* gC_out[B,Kcap,d] is overwritten; dL/dX is accumulated into gX_inout[B,N,d]; ws: partial centre gradients of the key slices */
size_t prifit_membership_bwd_workspace_bytes(int B, int Kcap, int d);
int prifit_membership_bwd(const float* C, const float* X, const float* bw, const int32_t* K,
                          const float* W, const float* smax, const float* gW,
                          int B, int N, int d, int Kcap, float* gC_out, float* gX_inout,
                          void* ws, size_t ws_bytes, void* stream);

/* k5-k7 -- weighted ellipsoid fit.  src/ellipsoid_fitting.py:19-69,119-141 and the SVD of
 *   src/fitting_utils.py:108-139.
 *   P[B,N,3], W[B,Kcap,N], noise[B,Kcap,3,3] = the U[0,1) draw of line 38.
 *   s_out[B,Kcap,3] half extents, V_out[B,Kcap,3,3] principal axes (columns), c_out[B,Kcap,3],
 *   valid_out[B,Kcap] uint8: 0 = dropped (the reference's `return -1`: cond > 1e5, or non-finite),
 *   ctx_out[B,Kcap,PRIFIT_FIT_CTX] saved for backward. */
#define PRIFIT_FIT_CTX 48
int prifit_fit_fwd(const float* P, const float* W, const int32_t* K, const float* noise,
                   int B, int N, int Kcap, float* s_out, float* V_out, float* c_out,
                   uint8_t* valid_out, float* ctx_out, void* stream);
/* gs/gV/gc: dL/d(s,V,c) (entries of dropped clusters are ignored).  gW_out[B,Kcap,N] overwritten
 *   (zero for dropped / padded clusters); gP_inout[B,N,3] optional accumulation (may be NULL).
 *   Includes the custom SVD backward of src/fitting_utils.py:67-105. */
int prifit_fit_bwd(const float* P, const float* W, const int32_t* K, const float* noise,
                   const float* ctx, const uint8_t* valid, const float* gs, const float* gV, const float* gc,
                   int B, int N, int Kcap, float* gW_out, float* gP_inout, void* stream);

/* noise_out[B,Kcap,3,3] from the flat host stream of U[0,1) draws: cluster (b, k) takes draw number
 *   (sum_{b' < b} K[b']) + k  (src/ellipsoid_fitting.py:38: one torch.rand(3,3) per attempted cluster, shapes outer,
 *   clusters inner); entries k >= K[b] are zero.  flat[B*Kcap,3,3].
 *   direct: optional device flag; when *direct != 0, flat is taken as already laid out [B,Kcap,3,3] (caller-supplied
 *   noise) and only the k >= K[b] entries are zeroed -- lets one captured launch sequence serve both cases. */
int prifit_noise_scatter(const float* flat, const int32_t* K, int B, int Kcap, const int32_t* direct,
                         float* noise_out, void* stream);
/* The same for the shapes [b0, b0 + Bb) only (all pointers are those of the whole batch). */
int prifit_noise_scatter_range(const float* flat, const int32_t* K, int b0, int Bb, int Kcap, const int32_t* direct,
                               float* noise_out, void* stream);
/* out[2B + 1] = [K | n_labels | ++*serial_inout]: the guard predicate's inputs (src/ellipsoid_utils.py:23) packed for ONE
 * device -> host copy; the serial number lets a host that polls pinned memory recognise the copy of the current step. */
int prifit_pack_counts(const int32_t* K, const int32_t* n_labels, int B, int32_t* serial_inout, int32_t* out, void* stream);
/* Device-side gate: a one-thread kernel on `stream` that returns once *flag >= want (device memory; bounded spin, ~50 ms).  The
 * host-side graph step bumps such a counter (prifit_pack_counts) when every branch has left the cluster stage; a prefetch stream
 * gated on it starts its host->device copy behind the all-seed kernel whatever the host's timing. */
int prifit_spin_until_ge(const int32_t* flag, int32_t want, void* stream);

/* k8 -- SDF half of the fitting loss.  convex_loss.py:313-343 + src/utils.py:407-411.
 *   Q[B,M,3]; loss_out[B] = 0.5 * mean_j (min_k |sdf_kj|)^2 over valid ellipsoids (0 if none);
 *   argmin_out[B,M] = arg of the min (-1 if none); sdf_out[B,M] = signed sdf of that ellipsoid. */
size_t prifit_sdf_workspace_bytes(int B, int M);
int prifit_sdf_loss_fwd(const float* Q, const float* s, const float* V, const float* c, const uint8_t* valid,
                        const int32_t* K, int B, int M, int Kcap, float* loss_out, int32_t* argmin_out,
                        float* sdf_out, void* ws, size_t ws_bytes, void* stream);
/* gloss[B] = dL/d loss_b.  gs/gV/gc [B,Kcap,...] overwritten; gQ_out[B,M,3] optional (may be NULL) */
int prifit_sdf_loss_bwd(const float* Q, const float* s, const float* V, const float* c, const uint8_t* valid,
                        const int32_t* K, const int32_t* argmin, const float* gloss, int B, int M, int Kcap,
                        float* gs_out, float* gV_out, float* gc_out, float* gQ_out, void* stream);

/* Intersection penalty between the ellipsoids of a shape.  convex_loss.py:346-441: version 3 =
 * compute_intersection_loss_volume_3 (called at :97; as written with torch_scatter.scatter_mean), version 4 =
 * compute_intersection_loss_volume_4.  With g_kj = min(sdf_k(q_j), -1e-3) over the valid ellipsoids of shape b:
 *   version 3: loss_b = mean_j (mean_{k != argmin_k g_kj} g_kj)^2        version 4: loss_b = mean_j (sum_k g_kj^2 - (min_k g_kj)^2)
 *   Q[B,M,3] probe points; loss_out[B] (0 for shapes with fewer than two ellipsoids); counted_out[B] = 1 where the shape has
 *   at least two; kstar_out[B,M] / aux_out[B,M] = arg-min ellipsoid and mean (v3) / min (v4), saved for the backward.
 *   Backward: gloss[B] -> gs, gV, gc (the probe points carry no gradient in the reference's use). */
size_t prifit_intersect_workspace_bytes(int B, int M);
int prifit_intersect_fwd(const float* Q, const float* s, const float* V, const float* c, const uint8_t* valid,
                         const int32_t* K, int B, int M, int Kcap, int version, float* loss_out, float* counted_out,
                         int32_t* kstar_out, float* aux_out, void* ws, size_t ws_bytes, void* stream);
int prifit_intersect_bwd(const float* Q, const float* s, const float* V, const float* c, const uint8_t* valid,
                         const int32_t* K, const int32_t* kstar, const float* aux, const float* gloss, int B, int M,
                         int Kcap, int version, float* gs_out, float* gV_out, float* gc_out, void* stream);

/* batch mean of the per-shape losses.  src/utils.py:418,425 (mean over the shapes that kept at least one
 *   ellipsoid) and train_partseg_shapenet.py:445 (mean over the replicas).
 *   has_out[B] = 1 if any valid[b,:] else 0;  stats_out[3] = { sum_b loss_b has_b, sum_b has_b, sum / max(n, 1) }.
 *   One CTA; deterministic summation order. */
int prifit_masked_mean_fwd(const float* loss_b, const uint8_t* valid, int B, int Kcap, float* has_out,
                           float* stats_out, void* stream);
/* gloss_out[b] = has[b] * (g_sum + g_mean / max(n, 1));  g_sum / g_mean: device scalars, either may be NULL;
 *   n = stats[1] of the forward call. */
int prifit_masked_mean_bwd(const float* g_sum, const float* g_mean, const float* has, const float* stats, int B,
                           float* gloss_out, void* stream);

/* f1 -- sampled-surface half of analytic_chamfer_distance.  src/utils.py:413-418 (KD-tree query on the host in the
 *   reference): S[B,Smax,3] source points (nS[b] valid rows per shape, NULL = all), T[B,M,3] target cloud.
 *   idx_out[B,Smax] = nearest target of every source point (lowest index on ties, -1 for padding),
 *   loss_out[b] = mean_i |s_i - t_idx(i)|^2 (0 when nS[b] == 0).
 *   Backward: gS_out[B,Smax,3] = 2 gloss[b] (s - t_idx) / nS[b]; gT_inout[B,M,3] optional accumulation of the opposite. */
size_t prifit_nn_workspace_bytes(int B, int Smax);
int prifit_nn_loss_fwd(const float* S, const int32_t* nS, const float* T, int B, int Smax, int M,
                       int32_t* idx_out, float* loss_out, void* ws, size_t ws_bytes, void* stream);
int prifit_nn_loss_bwd(const float* S, const int32_t* nS, const float* T, const int32_t* idx, const float* gloss,
                       int B, int Smax, int M, float* gS_out, float* gT_inout, void* stream);

/* f2 -- surface sampling of the fitted ellipsoids.  src/ellipsoid_utils.py:76-130, src/sample_ellipsoid.py:17-63.
 *   prifit_sample_counts : counts_out[B,Kcap] = round(total_points * area_k / sum area) (half to even; min_points where
 *                          <= 0; 0 for dropped / padded clusters), offsets_out[B,Kcap+1] = exclusive prefix sums.
 *   prifit_sample_surface: for slot i < offsets[b,Kcap] of shape b: owner_out = its ellipsoid, (U_out, V_out) = parameters of
 *                          a point drawn uniformly over that ellipsoid's surface (Philox stream `seed`, rejection from the
 *                          sphere); padding slots get owner -1.  x = a cos U sin V, y = b sin U sin V, z = c cos V. */
int prifit_sample_counts(const float* s, const uint8_t* valid, const int32_t* K, int B, int Kcap, int total_points,
                         int min_points, int32_t* counts_out, int32_t* offsets_out, void* stream);
int prifit_sample_surface(const float* s, const int32_t* offsets, int B, int Kcap, int Smax, uint64_t seed,
                          float* U_out, float* V_out, int32_t* owner_out, void* stream);
/* the differentiable map of src/sample_ellipsoid.py:50-63: pts_out[B,Smax,3] = (a cos U sin V, b sin U sin V, c cos V) V^T + centre
 *   (zero for padding slots), and its backward: gs/gV/gc [B,Kcap,...] overwritten (one CTA per ellipsoid, fixed order). */
int prifit_surface_points_fwd(const float* s, const float* V, const float* c, const float* U, const float* Vang,
                              const int32_t* owner, int B, int Kcap, int Smax, float* pts_out, void* stream);
int prifit_surface_points_bwd(const float* s, const float* V, const float* U, const float* Vang, const int32_t* offsets,
                              const float* gpts, int B, int Kcap, int Smax, float* gs_out, float* gV_out, float* gc_out,
                              void* stream);

/* f3 -- entropy regulariser.  convex_loss.py:209-225 (entropy) on the sub-sample of convex_loss.py:59-62.
 *   X[B,N,d] unit rows (d = 64 or 128), idx[n] int32 = the sampled point indices (shared by all shapes, unique; NULL = all
 *   N points), loss_b_out[B] = sum_ij (1 + <x_i, x_j>)^2 / n^2 over the sample, computed from the second moments
 *   (n^2 + 2 |sum x|^2 + |sum x x^T|_F^2) -- the caller applies mean over shapes, margin 1.8 and relu.
 *   The moments stay in `ws` for the backward call: gX_inout[b, idx[i], :] += gloss_b[b] * (4 / n^2) (m + M x_i). */
size_t prifit_entropy_workspace_bytes(int B, int d);
int prifit_entropy_fwd(const float* X, const int32_t* idx, int B, int N, int d, int n, float* loss_b_out,
                       void* ws, size_t ws_bytes, void* stream);
int prifit_entropy_bwd(const float* X, const int32_t* idx, const float* gloss_b, int B, int N, int d, int n,
                       const void* ws, float* gX_inout, void* stream);

/* f4 -- PointNet++ geometric operators (models/pointnet_util.py), index semantics of the reference (first index on ties).
 *   prifit_fps            :63-84   xyz[B,N,3], start[B] int64 (the reference's torch.randint draw) -> idx_out[B,npoint] int64; N <= 16384
 *   prifit_ball_query     :87-107  first nsample indices (ascending) with square_distance <= radius^2, padded with the first hit
 *                                  (N when a query has none) -> idx_out[B,S,nsample] int64
 *   prifit_three_nn       :287-293 three nearest xyz2 points of every xyz1 point and their normalised inverse-distance weights
 *   prifit_interpolate_*  :294     out[B,N,D] = sum_k weight_k points2[idx_k]; backward accumulates into gpoints2_inout[B,S,D] */
int prifit_fps(const float* xyz, const int64_t* start, int B, int N, int npoint, int64_t* idx_out, void* stream);
int prifit_ball_query(const float* xyz, const float* new_xyz, int B, int N, int S, float radius, int nsample,
                      int64_t* idx_out, void* stream);
int prifit_three_nn(const float* xyz1, const float* xyz2, int B, int N, int S, int32_t* idx_out, float* weight_out, void* stream);
int prifit_interpolate_fwd(const float* points2, const int32_t* idx, const float* weight, int B, int N, int S, int D,
                           float* out, void* stream);
int prifit_interpolate_bwd(const float* gout, const int32_t* idx, const float* weight, int B, int N, int S, int D,
                           float* gpoints2_inout, void* stream);

/* diagnostics -- hardware self-test of the tcgen05 / TMA descriptor encodings the tensor-core engine
 *   uses: D[128,128] = A . B^T (mode 0: B K-major in shared memory) or A . B (mode 1: B MN-major), A staged
 *   in tensor memory, B fetched by TMA with SWIZZLE_128B; lbo/sbo = descriptor byte offsets under test.
 *   Operands are rounded to fp16 like the engine's; ws >= 33 KB of device scratch. */
int prifit_debug_tc_probe(const float* A, const float* Bm, int mode, int lbo_bytes, int sbo_bytes,
                          float* D, void* ws, void* stream);

/* diagnostics -- the distance matrix 2 - 2 X X^T exactly as the tensor-core Gram engine sees it
 *   (split-fp16 operands, three tcgen05.mma per 16 elements, fp32 accumulate, clamped to [0, 4)):
 *   dist_out[B,N,N]; d = 128; ws >= 4 * B * N * 128 + 256 bytes.  Used by the tests to bound the
 *   engine's error against the margin of the bandwidth candidate pass. */
int prifit_debug_tc_gram(const float* X, int B, int N, float* dist_out, void* ws, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PRIFIT_B200_H */
