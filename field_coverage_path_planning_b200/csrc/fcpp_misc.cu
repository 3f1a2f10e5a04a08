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
// field of candidate c: the caller's array, or (factored candidate sets) the library's candidate record
__device__ __forceinline__ int cand_field_of(const int32_t *__restrict__ cf, const CandRec *__restrict__ recs, int64_t c)
{
    return cf ? cf[c] : recs[c].field;
}

__global__ void argmin_cost(const fcpp_summary *__restrict__ sm, const int32_t *__restrict__ cf,
                            const CandRec *__restrict__ recs, int64_t n, int kind, unsigned long long *key)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n && sm[c].status == 0) atomicMin(&key[cand_field_of(cf, recs, c)], order_bits(cand_cost(sm[c], kind)));
}
__global__ void argmin_index(const fcpp_summary *__restrict__ sm, const int32_t *__restrict__ cf,
                             const CandRec *__restrict__ recs, int64_t n, int kind, int64_t base,
                             const unsigned long long *__restrict__ key, int64_t *cand)
{
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n || sm[c].status != 0) return;
    const int f = cand_field_of(cf, recs, c);
    if (order_bits(cand_cost(sm[c], kind)) == key[f]) atomicMin((long long *)&cand[f], (long long)(base + c));
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

// atomicMin(&table[f], v) for the active lanes of a warp (all 32 lanes call it): one atomic per warp when its active
// lanes share the field, else one per lane.  Values are unsigned-ordered (candidate indices are >= 0).
__device__ __forceinline__ void warp_field_min(unsigned long long *table, int f, unsigned long long v, bool active)
{
    const unsigned m = __ballot_sync(0xffffffffu, active);
    if (m == 0) return;
    const int f0 = __shfl_sync(0xffffffffu, f, __ffs(m) - 1);
    if (__all_sync(0xffffffffu, !active || f == f0)) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const unsigned long long o = __shfl_xor_sync(0xffffffffu, v, d);
            v = o < v ? o : v;
        }
        if ((threadIdx.x & 31) == 0) atomicMin(&table[f0], v);
    } else if (active) {
        atomicMin(&table[f], v);
    }
}

// the four phases in ONE CTA (small batches: four dependent launches cost more than the work)
constexpr int ARGMIN_SMALL_THREADS = 1024;
__global__ void __launch_bounds__(ARGMIN_SMALL_THREADS) argmin_small(const fcpp_summary *__restrict__ sm,
                                                                      const int32_t *__restrict__ cf,
                                                                      const CandRec *__restrict__ recs, int64_t n, int kind,
                                                                      int64_t base, unsigned long long *key, int64_t *cand,
                                                                      int n_fields)
{
    for (int f = threadIdx.x; f < n_fields; f += ARGMIN_SMALL_THREADS) {
        key[f] = ~0ull;
        cand[f] = 0x7fffffffffffffffll;
    }
    __threadfence();
    __syncthreads();
    // candidates of one field are contiguous (field-major order), so a warp usually holds ONE field: its minimum is
    // reduced by shuffles and costs one atomic instead of 32 on the same address (config 2: 4096 candidates, 1 field)
    for (int64_t c0 = 0; c0 < n; c0 += ARGMIN_SMALL_THREADS) {
        const int64_t c = c0 + threadIdx.x;
        const bool on = c < n && sm[c].status == 0;
        const int f = on ? cand_field_of(cf, recs, c) : -1;
        warp_field_min(key, f, on ? order_bits(cand_cost(sm[c], kind)) : ~0ull, on);
    }
    __threadfence();
    __syncthreads();
    for (int64_t c0 = 0; c0 < n; c0 += ARGMIN_SMALL_THREADS) {
        const int64_t c = c0 + threadIdx.x;
        bool on = c < n && sm[c].status == 0;
        const int f = on ? cand_field_of(cf, recs, c) : -1;
        on = on && order_bits(cand_cost(sm[c], kind)) == __ldcg(&key[f]);
        warp_field_min(reinterpret_cast<unsigned long long *>(cand), on ? f : -1,
                       on ? (unsigned long long)(base + c) : ~0ull, on);
    }
    __threadfence();
    __syncthreads();
    for (int f = threadIdx.x; f < n_fields; f += ARGMIN_SMALL_THREADS) {
        const unsigned long long k = __ldcg(&key[f]);
        double *out = reinterpret_cast<double *>(key);
        if (k == ~0ull) {
            out[f] = INFINITY;
            cand[f] = -1;
        } else {
            out[f] = unorder_bits(k);
        }
    }
}

// multi-GPU: per-field merge of the ranks' local bests after ONE all-gather of (cost, candidate)
// words; lowest cost wins, ties go to the lowest global candidate index, -1 = no candidate
__global__ void argmin_merge_kernel(const long long *__restrict__ g, int world, int n_fields,
                                    double *__restrict__ best_cost, long long *__restrict__ best_cand)
{
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n_fields) return;
    double bc = INFINITY;
    long long bi = -1;
    for (int r = 0; r < world; ++r) {
        const long long *row = g + (int64_t)r * 2 * n_fields;
        const double c = __longlong_as_double(row[f]);
        const long long i = row[n_fields + f];
        if (i >= 0 && (bi < 0 || c < bc || (c == bc && i < bi))) {
            bc = c;
            bi = i;
        }
    }
    best_cost[f] = bc;
    best_cand[f] = bi;
}

// multi-GPU, fused: exchange of the ranks' (cost, candidate) words over PEER MEMORY (NVLink P2P stores
// into every rank's symmetric buffer + a release/acquire flag per rank) and the merge, in ONE kernel —
// no NCCL call, no separate merge launch.  slots = [2 parities][world][2 F] words on every rank; a rank
// can run at most one call ahead of a peer (it needs the peer's flag of the current call to finish),
// so two parities are enough.  Flags hold the epoch (monotonic) of the rank's last completed write.
struct ExchangePeers {
    unsigned long long buf[FCPP_MAX_PEERS];   // peer p's slot array, mapped into this process
    unsigned long long flag[FCPP_MAX_PEERS];  // peer p's flag array (uint32 [world])
};

__device__ __forceinline__ void st_release_sys(unsigned int *p, unsigned int v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

constexpr int EXCH_THREADS = 512;
__global__ void __launch_bounds__(EXCH_THREADS) argmin_exchange_kernel(const ExchangePeers peers, int world, int rank,
                                                                        int n_fields, unsigned int epoch,
                                                                        double *best_cost, long long *best_cand,
                                                                        long long timeout_cycles)
{
    const int tid = threadIdx.x;
    const int64_t words = 2 * (int64_t)n_fields;                    // cost bits [F] then candidates [F]
    const int64_t par = (int64_t)(epoch & 1u) * world * words;
    // 1. my words into slot `rank` of every rank (peer stores; own copy too)
    for (int p = 0; p < world; ++p) {
        long long *dst = reinterpret_cast<long long *>(peers.buf[p]) + par + (int64_t)rank * words;
        for (int64_t i = tid; i < words; i += EXCH_THREADS)
            dst[i] = (i < n_fields) ? __double_as_longlong(best_cost[i]) : best_cand[i - n_fields];
    }
    __threadfence_system();
    __syncthreads();
    // 2. publish: my flag on every rank; 3. wait for every rank's flag here
    __shared__ int timed_out;
    if (tid == 0) timed_out = 0;
    __syncthreads();
    if (tid < world) {
        st_release_sys(reinterpret_cast<unsigned int *>(peers.flag[tid]) + rank, epoch);
        const unsigned int *mine = reinterpret_cast<const unsigned int *>(peers.flag[rank]) + tid;
        // gentle polling: relaxed loads with a back-off (a tight acquire loop on the line the peers are
        // writing to slows their stores down), one acquire at the end
        const long long t0 = clock64();
        while ((int)(*reinterpret_cast<const volatile unsigned int *>(mine) - epoch) < 0) {
            __nanosleep(256);
            if (clock64() - t0 > timeout_cycles) {
                timed_out = 1;
                break;
            }
        }
        (void)ld_acquire_sys(mine);
    }
    __syncthreads();
    // 4. merge (the rule of argmin_merge_kernel); a peer that never arrived poisons the result
    const volatile long long *g = reinterpret_cast<const volatile long long *>(peers.buf[rank]) + par;
    for (int f = tid; f < n_fields; f += EXCH_THREADS) {
        double bc = INFINITY;
        long long bi = -1;
        for (int r = 0; r < world; ++r) {
            const double c = __longlong_as_double(g[(int64_t)r * words + f]);
            const long long i = g[(int64_t)r * words + n_fields + f];
            if (i >= 0 && (bi < 0 || c < bc || (c == bc && i < bi))) {
                bc = c;
                bi = i;
            }
        }
        best_cost[f] = timed_out ? __longlong_as_double(0x7ff8000000000000ll) : bc;
        best_cand[f] = timed_out ? -2 : bi;
    }
}

// multi-GPU: this rank's contribution to the all-gather of the winners' summary records — the 176-byte record of
// every field whose global winner lies in [lo, hi), zeros elsewhere (a SUM over the ranks then is the record)
__global__ void winner_records_kernel(const fcpp_summary *__restrict__ sm, int64_t lo, int64_t hi,
                                      const long long *__restrict__ best_cand, int n_fields, int4 *__restrict__ out)
{
    constexpr int W = sizeof(fcpp_summary) / 16;  // 11 int4 words per record
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)n_fields * W) return;
    const int f = (int)(t / W), k = (int)(t - (int64_t)f * W);
    const long long b = best_cand[f];
    out[t] = (b >= lo && b < hi) ? reinterpret_cast<const int4 *>(sm + (b - lo))[k] : make_int4(0, 0, 0, 0);
}

// multi_field_planner.py:263-288 ("mfp"): D[i][j] = ||pos_i - pos_j||, 0 on the diagonal
__global__ void distance_matrix_kernel(const double *__restrict__ pos, int n, double *__restrict__ D)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y;
    if (j >= n) return;
    const double dx = pos[2 * i] - pos[2 * j], dy = pos[2 * i + 1] - pos[2 * j + 1];
    D[(int64_t)i * n + j] = (i == j) ? 0.0 : sqrt(dx * dx + dy * dy);
}

// mfp:290-320 for every ordered pair of nodes at once.  Node 0 is the depot (one candidate point),
// node f+1 is field f (its four vertices are both exit and entry points, mfp:140-141).
// C[a][b] = min over (from point of a) x (to point of b) of the distance, from points in the outer
// loop, strict '<' so the FIRST minimum wins; arg[a][b] = from_index * 4 + to_index.
__global__ void connection_matrix_kernel(const double *__restrict__ verts, int n_fields, double depot_x,
                                         double depot_y, double *__restrict__ Cm, int32_t *__restrict__ arg)
{
    const int n = n_fields + 1;
    const int b = blockIdx.x * blockDim.x + threadIdx.x, a = blockIdx.y;
    if (b >= n) return;
    double fx[4], fy[4], tx[4], ty[4];
    const int nf = a == 0 ? 1 : 4, nt = b == 0 ? 1 : 4;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        fx[k] = a == 0 ? depot_x : verts[(int64_t)(a - 1) * 8 + 2 * k];
        fy[k] = a == 0 ? depot_y : verts[(int64_t)(a - 1) * 8 + 2 * k + 1];
        tx[k] = b == 0 ? depot_x : verts[(int64_t)(b - 1) * 8 + 2 * k];
        ty[k] = b == 0 ? depot_y : verts[(int64_t)(b - 1) * 8 + 2 * k + 1];
    }
    double best = INFINITY;
    int bi = 0;
    for (int p = 0; p < nf; ++p)
        for (int q = 0; q < nt; ++q) {
            const double dx = fx[p] - tx[q], dy = fy[p] - ty[q];
            const double d = sqrt(dx * dx + dy * dy);
            if (d < best) {
                best = d;
                bi = p * 4 + q;
            }
        }
    Cm[(int64_t)a * n + b] = best;
    arg[(int64_t)a * n + b] = bi;
}

}  // namespace

cudaError_t fcpp_launch_argmin_merge(fcpp_handle *h, const int64_t *d_gathered, int world, int32_t n_fields,
                                     double *d_best_cost, int64_t *d_best_cand, cudaStream_t st)
{
    if (n_fields == 0) return cudaSuccess;
    argmin_merge_kernel<<<(n_fields + 255) / 256, 256, 0, st>>>((const long long *)d_gathered, world, n_fields,
                                                                d_best_cost, (long long *)d_best_cand);
    h->launches++;
    return cudaGetLastError();
}

// number of candidates whose status has one of the bits of `mask`
__global__ void status_count_kernel(const fcpp_summary *__restrict__ sm, int64_t n, int mask, int32_t *__restrict__ count)
{
    int hits = 0;
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (int64_t)gridDim.x * blockDim.x)
        hits += (sm[c].status & mask) != 0;
    hits = __reduce_add_sync(0xffffffffu, hits);
    if ((threadIdx.x & 31) == 0 && hits) atomicAdd(count, hits);
}

cudaError_t fcpp_launch_status_count(fcpp_handle *h, const fcpp_summary *d_summary, int64_t n, int mask, int32_t *d_count,
                                     cudaStream_t st)
{
    cudaError_t e = cudaMemsetAsync(d_count, 0, sizeof(int32_t), st);
    if (e != cudaSuccess || n == 0) return e;
    int64_t blocks = (n + 255) / 256;
    if (blocks > 4 * (int64_t)h->sm_count) blocks = 4 * (int64_t)h->sm_count;
    status_count_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_summary, n, mask, d_count);
    h->launches++;
    return cudaGetLastError();
}

cudaError_t fcpp_launch_winner_records(fcpp_handle *h, const fcpp_summary *d_summary, int64_t lo, int64_t hi,
                                       const int64_t *d_best_cand, int32_t n_fields, void *d_out, cudaStream_t st)
{
    if (n_fields == 0) return cudaSuccess;
    static_assert(sizeof(fcpp_summary) % 16 == 0, "records are copied as int4 words");
    const int64_t n = (int64_t)n_fields * (sizeof(fcpp_summary) / 16);
    winner_records_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_summary, lo, hi, (const long long *)d_best_cand,
                                                                       n_fields, (int4 *)d_out);
    h->launches++;
    return cudaGetLastError();
}

cudaError_t fcpp_launch_argmin_exchange(fcpp_handle *h, int32_t world, int32_t rank, int32_t n_fields, uint32_t epoch,
                                        const uint64_t *peer_bufs, const uint64_t *peer_flags, double *d_best_cost,
                                        int64_t *d_best_cand, cudaStream_t st)
{
    if (n_fields == 0) return cudaSuccess;
    ExchangePeers pp{};
    for (int p = 0; p < world; ++p) {
        pp.buf[p] = peer_bufs[p];
        pp.flag[p] = peer_flags[p];
    }
    argmin_exchange_kernel<<<1, EXCH_THREADS, 0, st>>>(pp, world, rank, n_fields, epoch, d_best_cost,
                                                        (long long *)d_best_cand, 20000000000ll /* ~10 s */);
    h->launches++;
    return cudaGetLastError();
}

cudaError_t fcpp_launch_distance_matrix(fcpp_handle *h, const double *d_pos, int32_t n, double *d_D, cudaStream_t st)
{
    if (n == 0) return cudaSuccess;
    distance_matrix_kernel<<<dim3((n + 127) / 128, n), 128, 0, st>>>(d_pos, n, d_D);
    h->launches++;
    return cudaGetLastError();
}

cudaError_t fcpp_launch_connection_matrix(fcpp_handle *h, const double *d_verts, int32_t n_fields, double depot_x,
                                          double depot_y, double *d_C, int32_t *d_arg, cudaStream_t st)
{
    const int n = n_fields + 1;
    connection_matrix_kernel<<<dim3((n + 127) / 128, n), 128, 0, st>>>(d_verts, n_fields, depot_x, depot_y, d_C, d_arg);
    h->launches++;
    return cudaGetLastError();
}

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
    if (n_cand <= 16384 && n_fields <= 16384) {
        argmin_small<<<1, ARGMIN_SMALL_THREADS, 0, st>>>(d_summary, d_cand_field, h->d_rec, n_cand, cost_kind, cand_base,
                                                           key, d_best_cand, n_fields);
        h->launches++;
        return cudaGetLastError();
    }
    const int th = 256;
    const unsigned fb = (unsigned)((n_fields + th - 1) / th);
    argmin_init<<<fb, th, 0, st>>>(key, d_best_cand, n_fields);
    h->launches++;
    if (n_cand > 0) {
        const unsigned cb = (unsigned)((n_cand + th - 1) / th);
        argmin_cost<<<cb, th, 0, st>>>(d_summary, d_cand_field, h->d_rec, n_cand, cost_kind, key);
        argmin_index<<<cb, th, 0, st>>>(d_summary, d_cand_field, h->d_rec, n_cand, cost_kind, cand_base, key, d_best_cand);
        h->launches += 2;
    }
    argmin_final<<<fb, th, 0, st>>>(key, d_best_cand, n_fields);
    h->launches++;
    return cudaGetLastError();
}
