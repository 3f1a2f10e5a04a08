// fcpp_fleet.cu — multi-vehicle split ("mvp" = multi_vehicle_planner.py).
//
// mvp:186-209 _cluster_fields hands the field centroids to sklearn.cluster.KMeans(n_clusters=V,
// random_state=42).fit_predict — a third-party dependency that is not part of the reference tree.  Its published
// algorithm (scikit-learn 1.x, `_kmeans_single_lloyd`): centre the data on its mean, then Lloyd iterations
//     E: label_i = argmin_j ||c_j||^2 - 2 x_i . c_j          (first minimum; the ||x_i||^2 term is dropped)
//     M: c_j = sum_{label_i = j} x_i / count_j; an EMPTY cluster takes the point farthest from its centre
//     stop: labels unchanged ("strict convergence"), or sum_j ||c_j_new - c_j_old||^2 <= tol * mean(var(X)),
//           or max_iter; without strict convergence one more E step so that the labels match the centres.
// The k-means++ seeding draws from numpy's RandomState on the host (multi_vehicle.py); this kernel is the
// iteration: ONE CTA per problem (a farm), the whole loop in one launch, every reduction in a fixed order
// (deterministic for a given problem size).  Batched: grid = number of problems (farms, or restarts of one).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include "fcpp_internal.cuh"

namespace {

int fcpp_fail(fcpp_handle *h, int code, const char *msg)
{
    snprintf(h->err, sizeof(h->err), "%s", msg);
    return code;
}
int fcpp_cuda_fail(fcpp_handle *h, cudaError_t e, const char *what)
{
    snprintf(h->err, sizeof(h->err), "%s: %s", what, cudaGetErrorString(e));
    return FCPP_ERR_CUDA;
}

constexpr int KM_THREADS = 256;
constexpr int KM_WARPS = KM_THREADS / 32;

// fixed-shape block reduction of two doubles (every thread gets the sums); scratch = 2 * KM_WARPS doubles
__device__ __forceinline__ void block_sum2(double &a, double &b, double *scratch)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    const int w = threadIdx.x >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
        scratch[2 * w] = a;
        scratch[2 * w + 1] = b;
    }
    __syncthreads();
    a = b = 0.0;
#pragma unroll
    for (int k = 0; k < KM_WARPS; ++k) {
        a += scratch[2 * k];
        b += scratch[2 * k + 1];
    }
}

struct KMeansArgs {
    const int64_t *pt_start;      // [P+1] first point of every problem in xy
    const double *xy;             // [sum n][2]
    const int64_t *center_start;  // [P+1] first centre of every problem in centers
    double *centers;              // [sum k][2]: initial centres in, final centres out
    int32_t *labels;              // [sum n]
    int32_t *n_iter;              // [P]
    double *inertia;              // [P]
    int max_iter;
    double tol;                   // relative: the stop threshold is tol * mean(var(X, axis 0))
};

// E step for the points of this thread: labels, "changed" flag
__device__ __forceinline__ int assign_points(const double *__restrict__ pts, int64_t n, double mx, double my, int k,
                                              const double *c, const double *cn, int32_t *labels)
{
    int changed = 0;
    for (int64_t i = threadIdx.x; i < n; i += KM_THREADS) {
        const double x = pts[2 * i] - mx, y = pts[2 * i + 1] - my;
        int best = 0;
        double bd = cn[0] + -2.0 * (x * c[0] + y * c[1]);
        for (int j = 1; j < k; ++j) {
            const double d = cn[j] + -2.0 * (x * c[2 * j] + y * c[2 * j + 1]);
            if (d < bd) {
                bd = d;
                best = j;
            }
        }
        changed |= (labels[i] != best);
        labels[i] = best;
    }
    return changed;
}

__global__ void __launch_bounds__(KM_THREADS) kmeans_lloyd_kernel(const KMeansArgs a)
{
    extern __shared__ double km_smem[];
    const int p = blockIdx.x;
    const int64_t p0 = a.pt_start[p], n = a.pt_start[p + 1] - p0;
    const int64_t c0 = a.center_start[p];
    const int k = (int)(a.center_start[p + 1] - c0);
    const double *pts = a.xy + 2 * p0;
    int32_t *labels = a.labels + p0;
    double *cg = a.centers + 2 * c0;
    double *c = km_smem;           // [2k] current centres (mean-centred)
    double *cnew = c + 2 * k;      // [2k] sums, then new centres
    double *cn = cnew + 2 * k;     // [k]  squared norms
    double *cw = cn + k;           // [k]  member counts
    double *scratch = cw + k;      // [2 * KM_WARPS]
    __shared__ int s_flag;
    __shared__ double s_shift;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (n <= 0 || k <= 0) {
        if (tid == 0) {
            a.n_iter[p] = 0;
            a.inertia[p] = 0.0;
        }
        return;
    }
    // mean and variance of the points (KMeans.fit centres X; _tolerance = tol * mean of the per-axis variances)
    double sx = 0.0, sy = 0.0;
    for (int64_t i = tid; i < n; i += KM_THREADS) {
        sx += pts[2 * i];
        sy += pts[2 * i + 1];
    }
    block_sum2(sx, sy, scratch);
    const double mx = sx / (double)n, my = sy / (double)n;
    double vx = 0.0, vy = 0.0;
    for (int64_t i = tid; i < n; i += KM_THREADS) {
        const double dx = pts[2 * i] - mx, dy = pts[2 * i + 1] - my;
        vx += dx * dx;
        vy += dy * dy;
    }
    block_sum2(vx, vy, scratch);
    const double tol = a.tol * 0.5 * (vx / (double)n + vy / (double)n);
    for (int j = tid; j < k; j += KM_THREADS) {
        c[2 * j] = cg[2 * j] - mx;
        c[2 * j + 1] = cg[2 * j + 1] - my;
    }
    for (int64_t i = tid; i < n; i += KM_THREADS) labels[i] = -1;
    __syncthreads();

    int it = 0;
    bool strict = false;
    for (; it < a.max_iter;) {
        for (int j = tid; j < k; j += KM_THREADS) cn[j] = c[2 * j] * c[2 * j] + c[2 * j + 1] * c[2 * j + 1];
        if (tid == 0) s_flag = 0;
        __syncthreads();
        const int changed = assign_points(pts, n, mx, my, k, c, cn, labels);
        if (changed) s_flag = 1;  // benign race: every writer stores 1
        __syncthreads();          // labels (global) and s_flag visible to the block
        // M step: cluster j is summed by warp j % KM_WARPS, lanes stride over the points, fixed shuffle tree
        for (int j = warp; j < k; j += KM_WARPS) {
            double ax = 0.0, ay = 0.0, cnt = 0.0;
            for (int64_t i = lane; i < n; i += 32)
                if (labels[i] == j) {
                    ax += pts[2 * i] - mx;
                    ay += pts[2 * i + 1] - my;
                    cnt += 1.0;
                }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                ax += __shfl_xor_sync(0xffffffffu, ax, o);
                ay += __shfl_xor_sync(0xffffffffu, ay, o);
                cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
            }
            if (lane == 0) {
                cnew[2 * j] = ax;
                cnew[2 * j + 1] = ay;
                cw[j] = cnt;
            }
        }
        __syncthreads();
        if (tid == 0) {
            // empty clusters (rare after k-means++): the points farthest from their own centre become the new
            // centres, farthest first, for the empty clusters in ascending order (_relocate_empty_clusters_dense)
            int n_empty = 0;
            for (int j = 0; j < k; ++j) n_empty += (cw[j] == 0.0);
            double last_d = INFINITY;
            int64_t last_i = -1;
            for (int j = 0; j < k && n_empty > 0; ++j) {
                if (cw[j] != 0.0) continue;
                // the farthest point not taken yet: largest (distance, then lowest index) below the last pick
                double bd = -1.0;
                int64_t bi = -1;
                for (int64_t i = 0; i < n; ++i) {
                    const int l = labels[i];
                    const double dx = pts[2 * i] - mx - c[2 * l], dy = pts[2 * i + 1] - my - c[2 * l + 1];
                    const double d = dx * dx + dy * dy;
                    const bool below = d < last_d || (d == last_d && i > last_i);
                    if (below && d > bd) {
                        bd = d;
                        bi = i;
                    }
                }
                if (bi < 0) break;
                last_d = bd;
                last_i = bi;
                const int l = labels[bi];
                const double x = pts[2 * bi] - mx, y = pts[2 * bi + 1] - my;
                cnew[2 * l] -= x;
                cnew[2 * l + 1] -= y;
                cw[l] -= 1.0;
                cnew[2 * j] = x;
                cnew[2 * j + 1] = y;
                cw[j] = 1.0;
            }
            double shift = 0.0;
            for (int j = 0; j < k; ++j) {
                if (cw[j] > 0.0) {
                    cnew[2 * j] /= cw[j];
                    cnew[2 * j + 1] /= cw[j];
                }
                const double dx = cnew[2 * j] - c[2 * j], dy = cnew[2 * j + 1] - c[2 * j + 1];
                const double s = sqrt(dx * dx + dy * dy);   // _center_shift: the norm, squared again below
                shift += s * s;
            }
            s_shift = shift;
        }
        __syncthreads();
        for (int j = tid; j < 2 * k; j += KM_THREADS) c[j] = cnew[j];   // centers, centers_new = centers_new, centers
        ++it;
        const bool same = (s_flag == 0);
        const bool small = (s_shift <= tol);
        __syncthreads();
        if (same) {
            strict = true;
            break;
        }
        if (small) break;
    }
    if (!strict) {  // one more E step so that the labels match the final centres
        for (int j = tid; j < k; j += KM_THREADS) cn[j] = c[2 * j] * c[2 * j] + c[2 * j + 1] * c[2 * j + 1];
        __syncthreads();
        assign_points(pts, n, mx, my, k, c, cn, labels);
        __syncthreads();
    }
    double in = 0.0, unused = 0.0;
    for (int64_t i = tid; i < n; i += KM_THREADS) {
        const int l = labels[i];
        const double dx = pts[2 * i] - mx - c[2 * l], dy = pts[2 * i + 1] - my - c[2 * l + 1];
        in += dx * dx + dy * dy;
    }
    block_sum2(in, unused, scratch);
    for (int j = tid; j < k; j += KM_THREADS) {
        cg[2 * j] = c[2 * j] + mx;
        cg[2 * j + 1] = c[2 * j + 1] + my;
    }
    if (tid == 0) {
        a.n_iter[p] = it;
        a.inertia[p] = in;
    }
}

// ---------------------------------------------------------------------------------------------
// 2-opt tour improvement — the `TSPSolver.solve(distance_matrix)` that multi_field_planner.py:176 and
// multi_vehicle_planner.py:131 import from `multi_field_planner_v37`, a module the reference does not ship.
// BUILD-DEFINED (no reference source; restated in oracle/tsp.py): nearest-neighbour tour from node 0 (first minimum),
// then best-improvement 2-opt on the closed tour — every pair of tour edges (i, i+1), (j, j+1) is evaluated in
// parallel, delta = (D[a][c] + D[b][d]) - (D[a][b] + D[c][d]); the move with the lowest delta below -1e-9 is applied
// (ties: the lowest (i, j)) by reversing tour[i+1 .. j]; node 0 stays first.  D must be symmetric.
// One CTA per problem, tour in shared memory, deterministic.
// ---------------------------------------------------------------------------------------------
constexpr int TSP_THREADS = 256;
constexpr double TSP_EPS = 1e-9;

struct TspArgs {
    const int64_t *mat_start;  // [P+1] offset of every problem's n x n matrix in D (doubles)
    const double *D;
    const int64_t *node_start; // [P+1] first node of every problem in tours
    int32_t *tours;            // [sum n]
    double *lengths;           // [P]
    int32_t *iters;            // [P]
    int max_iter;
};

// block-wide argmin of (value, index), ties to the lowest index; every thread gets the result
__device__ __forceinline__ void block_argmin(double &v, long long &idx, double *sv, long long *si)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, v, o);
        const long long oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (ov < v || (ov == v && oi < idx)) {
            v = ov;
            idx = oi;
        }
    }
    const int w = threadIdx.x >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
        sv[w] = v;
        si[w] = idx;
    }
    __syncthreads();
    v = sv[0];
    idx = si[0];
#pragma unroll
    for (int k = 1; k < TSP_THREADS / 32; ++k)
        if (sv[k] < v || (sv[k] == v && si[k] < idx)) {
            v = sv[k];
            idx = si[k];
        }
}

__global__ void __launch_bounds__(TSP_THREADS) two_opt_kernel(const TspArgs a)
{
    extern __shared__ int tsp_smem[];
    __shared__ double sv[TSP_THREADS / 32];
    __shared__ long long si[TSP_THREADS / 32];
    const int p = blockIdx.x, tid = threadIdx.x;
    const int n = (int)(a.node_start[p + 1] - a.node_start[p]);
    const double *D = a.D + a.mat_start[p];
    int32_t *out = a.tours + a.node_start[p];
    int *tour = tsp_smem;       // [n]
    int *used = tsp_smem + n;   // [n]
    if (n <= 0) {
        if (tid == 0) {
            a.lengths[p] = 0.0;
            a.iters[p] = 0;
        }
        return;
    }
    // nearest-neighbour tour from node 0
    for (int k = tid; k < n; k += TSP_THREADS) used[k] = (k == 0);
    if (tid == 0) tour[0] = 0;
    __syncthreads();
    for (int step = 1; step < n; ++step) {
        const int cur = tour[step - 1];
        double bv = INFINITY;
        long long bi = 0x7fffffffffffffffll;
        for (int k = tid; k < n; k += TSP_THREADS)
            if (!used[k]) {
                const double d = D[(int64_t)cur * n + k];
                if (d < bv || (d == bv && k < bi)) {
                    bv = d;
                    bi = k;
                }
            }
        block_argmin(bv, bi, sv, si);
        if (tid == 0) {
            tour[step] = (int)bi;
            used[bi] = 1;
        }
        __syncthreads();
    }
    // best-improvement 2-opt
    int it = 0;
    const long long n_pairs = (long long)n * n;  // (i, j) as i * n + j, only 0 <= i < j <= n - 1 with j - i >= 2 count
    for (; it < a.max_iter; ++it) {
        double bv = INFINITY;
        long long bi = 0x7fffffffffffffffll;
        for (long long q = tid; q < n_pairs; q += TSP_THREADS) {
            const int i = (int)(q / n), j = (int)(q - (long long)i * n);
            if (j < i + 2 || (i == 0 && j == n - 1)) continue;  // adjacent edges (and the same pair seen cyclically)
            const int ta = tour[i], tb = tour[i + 1], tc = tour[j], td = tour[j + 1 == n ? 0 : j + 1];
            const double delta = (D[(int64_t)ta * n + tc] + D[(int64_t)tb * n + td]) -
                                 (D[(int64_t)ta * n + tb] + D[(int64_t)tc * n + td]);
            if (delta < bv || (delta == bv && q < bi)) {
                bv = delta;
                bi = q;
            }
        }
        block_argmin(bv, bi, sv, si);
        if (!(bv < -TSP_EPS)) break;
        const int i = (int)(bi / n), j = (int)(bi - (long long)i * n);
        __syncthreads();
        for (int k = tid; k < (j - i) / 2; k += TSP_THREADS) {  // reverse tour[i+1 .. j]
            const int x = tour[i + 1 + k];
            tour[i + 1 + k] = tour[j - k];
            tour[j - k] = x;
        }
        __syncthreads();
    }
    for (int k = tid; k < n; k += TSP_THREADS) out[k] = tour[k];
    if (tid == 0) {
        double len = 0.0;  // closed tour, left to right (ga:174-181)
        for (int k = 0; k < n; ++k) len += D[(int64_t)tour[k] * n + tour[k + 1 == n ? 0 : k + 1]];
        a.lengths[p] = len;
        a.iters[p] = it;
    }
}

}  // namespace

extern "C" int fcpp_kmeans_lloyd(fcpp_handle *h, int32_t n_problems, const int64_t *d_pt_start, const double *d_xy,
                                 const int64_t *d_center_start, int32_t max_clusters, double *d_centers,
                                 int32_t *d_labels, int32_t max_iter, double tol, int32_t *d_n_iter,
                                 double *d_inertia, void *stream)
{
    if (!h) return FCPP_ERR_INVALID;
    if (n_problems < 0 || max_clusters < 0 || max_iter < 0 || !(tol >= 0.0))
        return fcpp_fail(h, FCPP_ERR_INVALID, "fcpp_kmeans_lloyd: bad argument");
    if (n_problems == 0) return FCPP_OK;
    if (!d_pt_start || !d_xy || !d_center_start || !d_centers || !d_labels || !d_n_iter || !d_inertia)
        return fcpp_fail(h, FCPP_ERR_INVALID, "fcpp_kmeans_lloyd: NULL pointer");
    if (max_clusters > 4096) return fcpp_fail(h, FCPP_ERR_INVALID, "fcpp_kmeans_lloyd: more than 4096 clusters per problem");
    cudaSetDevice(h->device);
    KMeansArgs a{d_pt_start, d_xy, d_center_start, d_centers, d_labels, d_n_iter, d_inertia, max_iter, tol};
    const size_t bytes = sizeof(double) * ((size_t)6 * max_clusters + 2 * KM_WARPS);
    if (bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kmeans_lloyd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return fcpp_cuda_fail(h, e, "fcpp_kmeans_lloyd");
    }
    kmeans_lloyd_kernel<<<(unsigned)n_problems, KM_THREADS, bytes, (cudaStream_t)stream>>>(a);
    h->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fcpp_cuda_fail(h, e, "fcpp_kmeans_lloyd");
    return FCPP_OK;
}

extern "C" int fcpp_tsp_two_opt(fcpp_handle *h, int32_t n_problems, const int64_t *d_mat_start, const double *d_D,
                                const int64_t *d_node_start, int32_t max_nodes, int32_t *d_tours, double *d_lengths,
                                int32_t *d_iters, int32_t max_iter, void *stream)
{
    if (!h) return FCPP_ERR_INVALID;
    if (n_problems < 0 || max_nodes < 0 || max_iter < 0) return fcpp_fail(h, FCPP_ERR_INVALID, "fcpp_tsp_two_opt: bad argument");
    if (n_problems == 0) return FCPP_OK;
    if (!d_mat_start || !d_D || !d_node_start || !d_tours || !d_lengths || !d_iters)
        return fcpp_fail(h, FCPP_ERR_INVALID, "fcpp_tsp_two_opt: NULL pointer");
    if (max_nodes > 16384) return fcpp_fail(h, FCPP_ERR_INVALID, "fcpp_tsp_two_opt: more than 16384 nodes per problem");
    cudaSetDevice(h->device);
    TspArgs a{d_mat_start, d_D, d_node_start, d_tours, d_lengths, d_iters, max_iter};
    const size_t bytes = sizeof(int) * 2 * (size_t)(max_nodes > 0 ? max_nodes : 1);
    if (bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(two_opt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) return fcpp_cuda_fail(h, e, "fcpp_tsp_two_opt");
    }
    two_opt_kernel<<<(unsigned)n_problems, TSP_THREADS, bytes, (cudaStream_t)stream>>>(a);
    h->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fcpp_cuda_fail(h, e, "fcpp_tsp_two_opt");
    return FCPP_OK;
}
