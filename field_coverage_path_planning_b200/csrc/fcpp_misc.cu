// fcpp_misc.cu — batched GA tour-length fitness (A12) and the per-field argmin of plan costs.
//
// Reference code replaced:
//   genetic_algorithm_solver.py:168-181  _calculate_fitness / _calculate_distance
//   (distance-matrix layout: multi_field_planner.py:263-288, node 0 = depot)
// The per-field argmin has no reference counterpart (build-defined cost
// len_main + len_head, ties to the lowest candidate index).
#include "fcpp_internal.cuh"

namespace {

constexpr int TOUR_THREADS = 128;
constexpr int TOUR_CHUNK = 32;

// One thread per tour: the FP64 sum runs left to right exactly like the reference's Python
// loop, so tour lengths are bit-identical to genetic_algorithm_solver.py:174-181.  Tour rows
// are staged through shared memory in 32-column chunks (coalesced 128-B row segments in,
// conflict-free padded columns out); D (n*n FP64, 323 KB at n=201) is gathered through the
// read-only path and stays L1/L2 resident.
__global__ void __launch_bounds__(TOUR_THREADS) tour_kernel(const double *__restrict__ D, int n,
                                                            const int32_t *__restrict__ pop, int64_t pop_size,
                                                            double *__restrict__ out, double *__restrict__ fit)
{
    __shared__ int32_t tile[TOUR_THREADS][TOUR_CHUNK + 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t t0 = (int64_t)blockIdx.x * TOUR_THREADS;
    const int64_t me = t0 + tid;
    double s = 0.0;
    int first = 0, prev = 0;
    for (int c0 = 0; c0 < n; c0 += TOUR_CHUNK) {
        const int nc = min(TOUR_CHUNK, n - c0);
        for (int rr = warp; rr < TOUR_THREADS; rr += TOUR_THREADS / 32) {
            const int64_t t = t0 + rr;
            if (t < pop_size && lane < nc) tile[rr][lane] = pop[t * n + c0 + lane];
        }
        __syncthreads();
        if (me < pop_size) {
            int k = 0;
            if (c0 == 0) {
                first = prev = tile[tid][0];
                k = 1;
            }
            for (; k < nc; ++k) {
                const int cur = tile[tid][k];
                s += __ldg(&D[(int64_t)prev * n + cur]);
                prev = cur;
            }
        }
        __syncthreads();
    }
    if (me < pop_size) {
        s += __ldg(&D[(int64_t)prev * n + first]);  // closing edge (ga:178 `(i + 1) % len(route)`)
        out[me] = s;
        if (fit) fit[me] = 1.0 / (s + 1e-6);  // ga:172
    }
}

__device__ __forceinline__ unsigned long long order_bits(double x)
{
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double unorder_bits(unsigned long long o)
{
    const unsigned long long b = (o >> 63) ? (o & 0x7fffffffffffffffull) : ~o;
    return __longlong_as_double((long long)b);
}

__device__ __forceinline__ double cand_cost(const fcpp_summary &s, int kind)
{
    return kind == 0 ? s.len_main + s.len_head : s.time_main + s.time_head;
}

__global__ void argmin_init(unsigned long long *key, int64_t *cand, int n_fields)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f < n_fields) {
        key[f] = ~0ull;
        cand[f] = 0x7fffffffffffffffll;
    }
}
__global__ void argmin_cost(const fcpp_summary *__restrict__ sm, const int32_t *__restrict__ cf, int64_t n,
                            int kind, unsigned long long *key)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n && sm[c].status == 0) atomicMin(&key[cf[c]], order_bits(cand_cost(sm[c], kind)));
}
__global__ void argmin_index(const fcpp_summary *__restrict__ sm, const int32_t *__restrict__ cf, int64_t n,
                             int kind, int64_t base, const unsigned long long *__restrict__ key, int64_t *cand)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n && sm[c].status == 0 && order_bits(cand_cost(sm[c], kind)) == key[cf[c]])
        atomicMin((long long *)&cand[cf[c]], (long long)(base + c));
}
__global__ void argmin_final(unsigned long long *key, int64_t *cand, int n_fields)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f < n_fields) {
        const unsigned long long k = key[f];
        double *out = reinterpret_cast<double *>(key);
        if (k == ~0ull) {
            out[f] = INFINITY;
            cand[f] = -1;
        } else {
            out[f] = unorder_bits(k);
        }
    }
}

}  // namespace

cudaError_t fcpp_launch_tours(fcpp_handle *h, const double *d_D, int32_t n, const int32_t *d_pop,
                              int64_t pop_size, double *d_out, double *d_fit, cudaStream_t st)
{
    if (pop_size == 0) return cudaSuccess;
    const unsigned blocks = (unsigned)((pop_size + TOUR_THREADS - 1) / TOUR_THREADS);
    tour_kernel<<<blocks, TOUR_THREADS, 0, st>>>(d_D, n, d_pop, pop_size, d_out, d_fit);
    h->launches++;
    return cudaGetLastError();
}

cudaError_t fcpp_launch_argmin(fcpp_handle *h, const fcpp_summary *d_summary, const int32_t *d_cand_field,
                               int64_t n_cand, int32_t n_fields, int cost_kind, int64_t cand_base,
                               double *d_best_cost, int64_t *d_best_cand, cudaStream_t st)
{
    if (n_fields == 0) return cudaSuccess;
    unsigned long long *key = reinterpret_cast<unsigned long long *>(d_best_cost);
    const int th = 256;
    const unsigned fb = (unsigned)((n_fields + th - 1) / th);
    argmin_init<<<fb, th, 0, st>>>(key, d_best_cand, n_fields);
    h->launches++;
    if (n_cand > 0) {
        const unsigned cb = (unsigned)((n_cand + th - 1) / th);
        argmin_cost<<<cb, th, 0, st>>>(d_summary, d_cand_field, n_cand, cost_kind, key);
        argmin_index<<<cb, th, 0, st>>>(d_summary, d_cand_field, n_cand, cost_kind, cand_base, key, d_best_cand);
        h->launches += 2;
    }
    argmin_final<<<fb, th, 0, st>>>(key, d_best_cand, n_fields);
    h->launches++;
    return cudaGetLastError();
}
