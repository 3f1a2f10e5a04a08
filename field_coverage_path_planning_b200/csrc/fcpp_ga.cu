// fcpp_ga.cu — the GA evolution operators of the multi-field TSP ordering, on the device
// (SURVEY.md §8(f) N1).  "ga" = genetic_algorithm_solver.py of the reference.
//
//   ga:137-166  _initialize_population / _greedy_init   -> ga_init_kernel
//   ga:183-196  _selection (tournament, first max wins) \
//   ga:198-242  _crossover / _ox_crossover               > ga_evolve_kernel (one warp per pair)
//   ga:244-252  _mutation (swap)                         /
//   ga:254-268  _elitism (overwrites the LAST children)  -> ga_rank_kernel + ga_elite_kernel
//   ga:68-116   best tracking / history / convergence    -> ga_track_kernel
//   ga:44-135   solve()                                  -> fcpp_ga_solve (CUDA graph of 2 generations)
//
// The reference draws from Python's unseeded global `random`, so parity for whole runs is
// statistical.  Parity for the OPERATORS is exact: every random decision of a generation
// (tournament draws, crossover flag and cut points, mutation flag and positions) can be written
// to a trace, and oracle/ga_ops.py — pinned against the unmodified reference operators driven
// through a scripted `random` — replays the same decisions on the host; the populations must be
// identical.  Random numbers are Philox4x32-10, counter = (index, stream, generation, block),
// key = seed: a decision depends only on (seed, generation, slot), never on scheduling.
#include "fcpp_internal.cuh"

namespace {

constexpr int GA_WARPS = 8;  // warps (= offspring pairs) per CTA at small n

struct GaState {
    double best_fit, best_len, mean_fit;
    int gen;       // generations completed (index of the generation being produced)
    int stagnant;  // ga:74 generations_without_improvement
    int done;      // convergence reached (ga:113-116): later kernels are no-ops
    int last_gen;  // value of `generation` when the loop ended
    int best_idx, pad;
};

// ------------------------------------------------------------------ Philox4x32-10
__device__ __forceinline__ uint4 philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint64_t seed)
{
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        c0 = h1 ^ c1 ^ k0;
        c1 = l1;
        c2 = h0 ^ c3 ^ k1;
        c3 = l0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}
enum : uint32_t { RS_INIT = 1, RS_TOURN = 2, RS_CROSS = 3, RS_MUT = 4 };

// uniform integer in [0, m) (multiply-shift; bias < m / 2^32)
__device__ __forceinline__ int below(uint32_t w, int m) { return (int)__umulhi(w, (uint32_t)m); }
// uniform double in [0, 1) with 53 random bits, like Python's random.random()
__device__ __forceinline__ double unit53(uint32_t a, uint32_t b)
{
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}

// ------------------------------------------------------------------ ga:137-166
// First half: uniformly random permutations; second half: start node i % n followed by the
// remaining nodes in uniformly random order (the reference's "greedy" initialisation picks the
// next node with random.choice, ga:161-163).  One thread per individual, Fisher-Yates in place.
__global__ void ga_init_kernel(fcpp_ga_config cfg, int n, int32_t *__restrict__ pop)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int half = cfg.population_size / 2;
    if (i >= 2 * half) return;
    int32_t *row = pop + (int64_t)i * n;
    for (int k = 0; k < n; ++k) row[k] = k;
    int first = 0;
    if (i >= half) {  // fix the start node, shuffle the rest
        const int s = (i - half) % n;
        row[0] = s;
        row[s] = 0;
        first = 1;
    }
    uint4 r = make_uint4(0, 0, 0, 0);
    for (int k = n - 1; k > first; --k) {
        const int t = n - 1 - k;
        if ((t & 3) == 0) r = philox((uint32_t)i, RS_INIT, 0u, (uint32_t)(t >> 2), cfg.seed);
        const uint32_t w = (t & 3) == 0 ? r.x : ((t & 3) == 1 ? r.y : ((t & 3) == 2 ? r.z : r.w));
        const int j = first + below(w, k - first + 1);
        const int32_t tmp = row[k];
        row[k] = row[j];
        row[j] = tmp;
    }
}

// ------------------------------------------------------------------ ga:254-268 (ranks)
// rank[i] = number of individuals that sort AFTER i in a stable ascending argsort of the fitness
// (greater fitness, or equal fitness and greater index): rank 0 = the best elite.
__global__ void ga_rank_kernel(const double *__restrict__ fit, int m, int *__restrict__ rank,
                               const GaState *__restrict__ state)
{
    if (state && state->done) return;
    __shared__ double tile[1024];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j0 = blockIdx.y * 1024;
    for (int k = threadIdx.x; k < 1024; k += blockDim.x) tile[k] = (j0 + k < m) ? fit[j0 + k] : -INFINITY;
    __syncthreads();
    if (i >= m) return;
    const double f = fit[i];
    const int nj = min(1024, m - j0);
    int c = 0;
    for (int k = 0; k < nj; ++k) {
        const double g = tile[k];
        c += (g > f) || (g == f && j0 + k > i);
    }
    if (c) atomicAdd(&rank[i], c);
}

// elites of the OLD population go to the tail of the new one, in ascending order of fitness
// (ga:262-266: `new_population[:-elite_size] + elites`)
__global__ void ga_elite_kernel(const int32_t *__restrict__ pop_in, const int *__restrict__ rank, int m_in, int n,
                                int e_take, int m_out, int32_t *__restrict__ pop_out,
                                const GaState *__restrict__ state)
{
    if (state && state->done) return;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= m_in) return;
    const int r = rank[i];
    if (r >= e_take) return;
    const int32_t *src = pop_in + (int64_t)i * n;
    int32_t *dst = pop_out + (int64_t)(m_out - 1 - r) * n;
    for (int k = threadIdx.x & 31; k < n; k += 32) dst[k] = src[k];
}

// ------------------------------------------------------------------ ga:183-252
// k distinct indices of [0, m) in uniformly random order (random.sample(range(m), k), ga:189) and
// the first one of maximal fitness (ga:190-194)
__device__ int tournament(const double *__restrict__ fit, int m, int k, uint32_t slot, uint32_t gen, uint64_t seed,
                          int32_t *trace /* [k] or null */)
{
    int chosen[FCPP_GA_MAX_TOURNAMENT];  // ascending
    int best = -1;
    double best_f = 0.0;
    uint4 r = make_uint4(0, 0, 0, 0);
    for (int t = 0; t < k; ++t) {
        if ((t & 3) == 0) r = philox(slot, RS_TOURN, gen, (uint32_t)(t >> 2), seed);
        const uint32_t w = (t & 3) == 0 ? r.x : ((t & 3) == 1 ? r.y : ((t & 3) == 2 ? r.z : r.w));
        int x = below(w, m - t);
        int pos = 0;
        while (pos < t && x >= chosen[pos]) {  // skip the indices already drawn
            ++x;
            ++pos;
        }
        for (int q = t; q > pos; --q) chosen[q] = chosen[q - 1];
        chosen[pos] = x;
        if (trace) trace[t] = x;
        const double f = fit[x];
        if (best < 0 || f > best_f) {
            best = x;
            best_f = f;
        }
    }
    return best;
}

// child = OX(seg parent S, fill parent F): child[a:b] = S[a:b]; the other genes in the order of
// F[b:] + F[:b], written from position b on, wrapping to 0 (ga:214-242).  Whole warp.
__device__ void ox_child(const int32_t *S, const int32_t *F, int32_t *child, uint32_t *used, int n, int a, int b,
                         int lane)
{
    const int nw = (n + 31) >> 5;
    for (int w = lane; w < nw; w += 32) used[w] = 0u;
    __syncwarp();
    for (int i = a + lane; i < b; i += 32) {
        const int g = S[i];
        child[i] = g;
        atomicOr(&used[g >> 5], 1u << (g & 31));
    }
    __syncwarp();
    const int tail = n - b;  // slots b .. n-1 are filled first
    int base = 0;
    for (int t0 = 0; t0 < n; t0 += 32) {
        const int t = t0 + lane;
        int idx = b + t;
        if (idx >= n) idx -= n;
        const bool valid = t < n;
        const int g = valid ? F[idx] : 0;
        const bool keep = valid && !((used[g >> 5] >> (g & 31)) & 1u);
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const int q = base + __popc(bal & ((1u << lane) - 1u));
            child[q < tail ? b + q : q - tail] = g;
        }
        base += __popc(bal);
    }
    __syncwarp();
}

__global__ void __launch_bounds__(GA_WARPS * 32)
    ga_evolve_kernel(fcpp_ga_config cfg, int gen_arg, int n, const int32_t *__restrict__ pop_in,
                     const double *__restrict__ fit, int m_in, int keep, int32_t *__restrict__ pop_out,
                     int32_t *__restrict__ trace, const GaState *__restrict__ state, int warps_per_cta)
{
    if (state && state->done) return;
    extern __shared__ __align__(16) unsigned char ga_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp >= warps_per_cta) return;
    const int pair = blockIdx.x * warps_per_cta + warp;
    const int pairs = (m_in + 1) >> 1;
    if (pair >= pairs) return;
    const uint32_t gen = (uint32_t)(gen_arg >= 0 ? gen_arg : state->gen);
    const int nw = (n + 31) >> 5;
    int32_t *pa = reinterpret_cast<int32_t *>(ga_smem) + (size_t)warp * (4 * n + 2 * nw);
    int32_t *pb = pa + n, *c1 = pb + n, *c2 = c1 + n;
    uint32_t *used1 = reinterpret_cast<uint32_t *>(c2 + n), *used2 = used1 + nw;
    int32_t *tr = trace ? trace + (int64_t)pair * FCPP_GA_TRACE_INTS : nullptr;

    // ---- selection: slot 2p and slot 2p+1 (slot 0 again for the odd one out, ga:205) ----
    const int slotA = 2 * pair, slotB = (2 * pair + 1 < m_in) ? 2 * pair + 1 : 0;
    int win = 0;
    if (lane < 2)
        win = tournament(fit, m_in, cfg.tournament_size, (uint32_t)(lane == 0 ? slotA : slotB), gen, cfg.seed,
                         tr ? tr + 12 + lane * FCPP_GA_MAX_TOURNAMENT : nullptr);
    const int wA = __shfl_sync(0xffffffffu, win, 0), wB = __shfl_sync(0xffffffffu, win, 1);
    for (int i = lane; i < n; i += 32) {
        pa[i] = pop_in[(int64_t)wA * n + i];
        pb[i] = pop_in[(int64_t)wB * n + i];
    }
    // ---- crossover decision and cut points (ga:207, :219) ----
    const uint4 rc = philox((uint32_t)pair, RS_CROSS, gen, 0u, cfg.seed);
    const bool cross = unit53(rc.x, rc.y) < cfg.crossover_rate;
    int a = 0, b = 0;
    if (cross && n >= 2) {
        int i = below(rc.z, n), j = below(rc.w, n - 1);
        if (j >= i) ++j;
        a = min(i, j);
        b = max(i, j);
    }
    __syncwarp();
    if (cross && n >= 2) {
        ox_child(pa, pb, c1, used1, n, a, b, lane);
        ox_child(pb, pa, c2, used2, n, a, b, lane);
    } else {
        for (int i = lane; i < n; i += 32) {
            c1[i] = pa[i];
            c2[i] = pb[i];
        }
        __syncwarp();
    }
    // ---- swap mutation, one decision per child (ga:246-250) ----
    int mut = 0, mi = 0, mj = 0;
    if (lane < 2) {
        const uint4 rm = philox((uint32_t)(2 * pair + lane), RS_MUT, gen, 0u, cfg.seed);
        if (unit53(rm.x, rm.y) < cfg.mutation_rate && n >= 2) {
            mut = 1;
            mi = below(rm.z, n);
            mj = below(rm.w, n - 1);
            if (mj >= mi) ++mj;
            int32_t *c = lane == 0 ? c1 : c2;
            const int32_t tmp = c[mi];
            c[mi] = c[mj];
            c[mj] = tmp;
        }
        if (tr) {
            tr[5 + 3 * lane] = mut;
            tr[6 + 3 * lane] = mi;
            tr[7 + 3 * lane] = mj;
        }
    }
    if (tr && lane == 0) {
        tr[0] = wA;
        tr[1] = wB;
        tr[2] = (cross && n >= 2) ? 1 : 0;
        tr[3] = a;
        tr[4] = b;
        tr[11] = cfg.tournament_size;
    }
    __syncwarp();
    // ---- children 2p, 2p+1; the last elite_size slots belong to the elites (ga:266) ----
    const int ca = 2 * pair, cb = 2 * pair + 1;
    for (int i = lane; i < n; i += 32) {
        if (ca < keep) pop_out[(int64_t)ca * n + i] = c1[i];
        if (cb < keep) pop_out[(int64_t)cb * n + i] = c2[i];
    }
}

// ------------------------------------------------------------------ ga:68-116 (one CTA)
// gen_arg = -1: initial population (ga:68-72); otherwise the bookkeeping of one generation.
__global__ void __launch_bounds__(1024)
    ga_track_kernel(const double *__restrict__ fit, const double *__restrict__ len, const int32_t *__restrict__ pop,
                    int m, int n, int threshold, int initial, GaState *state, int32_t *__restrict__ best_route,
                    double *__restrict__ history)
{
    if (state->done && !initial) return;
    __shared__ double sv[32], ss[32];
    __shared__ int si[32];
    __shared__ int upd;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double v = -INFINITY, s = 0.0;
    int idx = 0x7fffffff;
    for (int i = tid; i < m; i += blockDim.x) {
        const double f = fit[i];
        s += f;
        if (f > v || (f == v && i < idx)) {
            v = f;
            idx = i;
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, v, d);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, d);
        s += __shfl_xor_sync(0xffffffffu, s, d);
        if (ov > v || (ov == v && oi < idx)) {
            v = ov;
            idx = oi;
        }
    }
    if (lane == 0) {
        sv[warp] = v;
        si[warp] = idx;
        ss[warp] = s;
    }
    __syncthreads();
    if (tid == 0) {
        const int nwp = blockDim.x >> 5;
        for (int w = 1; w < nwp; ++w) {
            if (sv[w] > v || (sv[w] == v && si[w] < idx)) {
                v = sv[w];
                idx = si[w];
            }
            s += ss[w];
        }
        // np.argmax = first maximum (ga:70, :94); improvement is strict (ga:97)
        int u = 0;
        if (initial) {
            state->best_fit = v;
            state->best_len = len[idx];
            state->best_idx = idx;
            state->gen = 0;
            state->stagnant = 0;
            state->done = 0;
            state->last_gen = -1;
            u = 1;
        } else {
            const int g = state->gen;
            if (v > state->best_fit) {
                state->best_fit = v;
                state->best_len = len[idx];
                state->best_idx = idx;
                state->stagnant = 0;
                u = 1;
            } else {
                state->stagnant += 1;
            }
            state->mean_fit = s / (double)m;
            if (history) {
                history[2 * g] = state->best_fit;        // ga:107
                history[2 * g + 1] = state->mean_fit;    // ga:108
            }
            state->last_gen = g;
            state->gen = g + 1;
            if (state->stagnant >= threshold) state->done = 1;  // ga:113-116
        }
        upd = u;
    }
    __syncthreads();
    if (upd) {
        const int bi = state->best_idx;
        for (int k = tid; k < n; k += blockDim.x) best_route[k] = pop[(int64_t)bi * n + k];  // ga:99
    }
}

// rotate the best route so that node 0 (the depot) comes first (ga:119-120)
__global__ void ga_rotate_kernel(const int32_t *__restrict__ route, int n, int32_t *__restrict__ out)
{
    __shared__ int z;
    for (int k = threadIdx.x; k < n; k += blockDim.x)
        if (route[k] == 0) z = k;
    __syncthreads();
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        int s = k + z;
        if (s >= n) s -= n;
        out[k] = route[s];
    }
}

size_t evolve_smem(int n, int warps) { return (size_t)warps * (4 * (size_t)n + 2 * ((n + 31) / 32)) * 4; }

}  // namespace

// sizes of one generation (python slicing semantics of ga:262-266 included: elite_size 0 keeps
// NO child — `new[:-0]` is empty and `argsort[-0:]` is everything)
void fcpp_ga_sizes(const fcpp_ga_config &cfg, int m_in, int &keep, int &e_take, int &m_out)
{
    const int children = 2 * ((m_in + 1) / 2);
    if (cfg.elite_size <= 0) {
        keep = 0;
        e_take = m_in;
    } else {
        keep = children - cfg.elite_size > 0 ? children - cfg.elite_size : 0;
        e_take = cfg.elite_size < m_in ? cfg.elite_size : m_in;
    }
    m_out = keep + e_take;
}

cudaError_t fcpp_launch_ga_init(fcpp_handle *h, const fcpp_ga_config &cfg, int n, int32_t *d_pop, cudaStream_t st)
{
    const int m = 2 * (cfg.population_size / 2);
    if (m == 0) return cudaSuccess;
    ga_init_kernel<<<(m + 127) / 128, 128, 0, st>>>(cfg, n, d_pop);
    h->launches++;
    return cudaGetLastError();
}

// one generation: pop_in/fit (m_in individuals) -> pop_out (m_out individuals); d_rank is an
// int[m_in] workspace.  state == nullptr: stand-alone call with an explicit generation index.
cudaError_t fcpp_launch_ga_generation(fcpp_handle *h, const fcpp_ga_config &cfg, int gen, int n,
                                      const int32_t *d_pop_in, const double *d_fit, int m_in, int32_t *d_pop_out,
                                      int *d_rank, int32_t *d_trace, const void *d_state, cudaStream_t st)
{
    int keep, e_take, m_out;
    fcpp_ga_sizes(cfg, m_in, keep, e_take, m_out);
    const GaState *state = static_cast<const GaState *>(d_state);
    cudaError_t e = cudaMemsetAsync(d_rank, 0, sizeof(int) * (size_t)m_in, st);
    if (e != cudaSuccess) return e;
    const dim3 rg((m_in + 255) / 256, (m_in + 1023) / 1024);
    ga_rank_kernel<<<rg, 256, 0, st>>>(d_fit, m_in, d_rank, state);
    ga_elite_kernel<<<(m_in + 7) / 8, 256, 0, st>>>(d_pop_in, d_rank, m_in, n, e_take, m_out, d_pop_out, state);
    h->launches += 2;
    int warps = GA_WARPS;
    while (warps > 1 && evolve_smem(n, warps) > (size_t)h->max_smem_optin) warps >>= 1;
    const size_t smem = evolve_smem(n, warps);
    if (smem > (size_t)h->max_smem_optin) return cudaErrorInvalidValue;
    e = cudaFuncSetAttribute(ga_evolve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int pairs = (m_in + 1) / 2;
    ga_evolve_kernel<<<(pairs + warps - 1) / warps, GA_WARPS * 32, smem, st>>>(cfg, gen, n, d_pop_in, d_fit, m_in,
                                                                               keep, d_pop_out, d_trace, state, warps);
    h->launches++;
    return cudaGetLastError();
}

cudaError_t fcpp_launch_ga_track(fcpp_handle *h, const double *d_fit, const double *d_len, const int32_t *d_pop,
                                 int m, int n, int threshold, int initial, void *d_state, int32_t *d_best_route,
                                 double *d_history, cudaStream_t st)
{
    ga_track_kernel<<<1, 1024, 0, st>>>(d_fit, d_len, d_pop, m, n, threshold, initial, static_cast<GaState *>(d_state),
                                        d_best_route, d_history);
    h->launches++;
    return cudaGetLastError();
}

cudaError_t fcpp_launch_ga_rotate(fcpp_handle *h, const int32_t *d_route, int n, int32_t *d_out, cudaStream_t st)
{
    ga_rotate_kernel<<<1, 256, 0, st>>>(d_route, n, d_out);
    h->launches++;
    return cudaGetLastError();
}

size_t fcpp_ga_state_bytes() { return sizeof(GaState); }
void fcpp_ga_read_state(const void *host_copy, int &gen, int &stagnant, int &done, int &last_gen, double &best_fit,
                        double &best_len)
{
    const GaState *s = static_cast<const GaState *>(host_copy);
    gen = s->gen;
    stagnant = s->stagnant;
    done = s->done;
    last_gen = s->last_gen;
    best_fit = s->best_fit;
    best_len = s->best_len;
}
