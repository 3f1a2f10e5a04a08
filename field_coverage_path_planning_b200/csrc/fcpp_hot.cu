// fcpp_hot.cu — the two hot kernels of the batch path in ONE translation unit, plus their fusion.
//
// plan_gen_kernel (fcpp_plan.cu) and cover_kernel (fcpp_cover.cu) read the same candidate records and are
// independent of each other; both are instruction-issue / latency bound (ncu: ~50 % and ~58 % of the issue slots),
// and at two 512-thread coverage CTAs per SM the register file is full, so launching them on two streams gives no
// co-residency.  plan_cover_kernel runs both in one grid with CTA ROLES: every fifth CTA plans four candidates (one
// per 128-thread quarter, named barriers 1-4), the other CTAs rasterise one candidate's coverage each.  An SM then
// holds a mix of both roles.  Results are bit-identical to the separate launches (same bodies, same per-plan thread
// counts; tests/test_gpu_parity.py).  MEASURED (B200, round 2): config 2 fused 1.060 ms vs 0.216 + 0.776 ms apart,
// config 5 2.89 vs 2.97 ms, config 3 49.4 vs 58.1 ms — the fused grid does not put MORE warps on an SM (registers:
// 2 x 512 threads x 64), it only substitutes one role's warps for the other's, and both bodies stall alike, so the
// issue slots stay as idle as before; config 3 gains because its 737 280 mostly skipped coverage CTAs ride along.
// It is therefore OPT-IN (fcpp_set_cover_mode bit 2); the default is two launches.
#include "fcpp_plan.cu"
#include "fcpp_cover.cu"

namespace {

static_assert(FCPP_COVER_THREADS == 4 * FCPP_PLAN_GEN_THREADS, "four plans per coverage-sized CTA");

struct FusedArgs {
    PlanArgs p;  // p.b is the batch of both roles
    int pc, mode;
    const int32_t *rep;
    int64_t q_inter;      // plan quads interleaved with the coverage CTAs (one per five CTAs)
    int64_t n_quads;      // all plan quads = ceil(n_cand / 4)
    uint32_t plan_bytes;  // shared memory of one plan
};

__global__ void __launch_bounds__(FCPP_COVER_THREADS, FCPP_COVER_MINBLOCKS) plan_cover_kernel(const FusedArgs a)
{
    const int64_t n = a.p.b.n_cand;
    const int64_t bid = blockIdx.x;
    // CTA -> role.  The first 5 q_inter CTAs interleave four coverage CTAs and one plan quad; then the remaining
    // coverage CTAs, then the remaining plan quads.
    int64_t quad = -1, cover = -1;
    if (bid < 5 * a.q_inter) {
        if (bid % 5 == 4)
            quad = bid / 5;
        else
            cover = bid - bid / 5;
    } else {
        const int64_t rest = bid - 5 * a.q_inter;
        const int64_t covers_left = n - 4 * a.q_inter;
        if (rest < covers_left)
            cover = 4 * a.q_inter + rest;
        else
            quad = a.q_inter + (rest - covers_left);
    }
    if (quad >= 0) {
        const int sub = threadIdx.x >> 7;
        const int64_t cand = 4 * quad + sub;
        if (cand >= n) return;
        plan_gen_body(a.p, cover_smem + (size_t)sub * a.plan_bytes, threadIdx.x & 127, cand, QuarterSync{1 + sub});
    } else {
        cover_body(a.p.b, a.p.recs, a.p.trig, a.p.out.summary, a.pc, a.mode, a.rep, a.p.out.corner_bits,
                   a.p.out.corner_bits_stride, cover);
    }
}

}  // namespace

// plan + coverage of one batch: fused when four plans fit the shared memory of a coverage CTA, else two launches.
// *fused_out = 1 when the fused kernel ran.
cudaError_t fcpp_launch_plan_cover(fcpp_handle *h, const fcpp_batch &b, const fcpp_outputs &o, cudaStream_t st,
                                   int *fused_out)
{
    if (fused_out) *fused_out = 0;
    if (b.n_cand == 0) return cudaSuccess;
    FusedArgs a{};
    const size_t plan_bytes = (plan_gen_args(h, b, o, a.p) + 127) & ~size_t(127);
    CoverLaunch L;
    // two CTAs per SM: the limit a CTA's shared memory must stay under
    const size_t limit = ((size_t)h->max_smem_sm - 2 * 1024) / 2;
    int pc = cover_point_capacity(h->cover_pcap);
    const bool fuse = (h->cover_mode & 4) && b.turn_model != FCPP_TURN_OMEGA && 4 * plan_bytes <= limit && cover_smem_bytes(pc) <= limit &&
                      4 * plan_bytes <= (size_t)h->max_smem_optin;
    if (!fuse) {
        cudaError_t e = fcpp_launch_plan(h, b, o, st, nullptr);
        if (e != cudaSuccess) return e;
        return fcpp_launch_cover(h, b, o, st);
    }
    cudaError_t e = cover_prepare(h, b, st, L);
    if (e != cudaSuccess) return e;
    a.pc = L.pc;
    a.mode = h->cover_mode;
    a.rep = L.d_rep;
    a.n_quads = (b.n_cand + 3) / 4;
    a.q_inter = b.n_cand / 4;
    a.plan_bytes = (uint32_t)plan_bytes;
    const size_t bytes = L.bytes > 4 * plan_bytes ? L.bytes : 4 * plan_bytes;
    e = cudaFuncSetAttribute(plan_cover_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return e;
    plan_cover_kernel<<<(unsigned)(b.n_cand + a.n_quads), FCPP_COVER_THREADS, bytes, st>>>(a);
    h->launches++;
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (fused_out) *fused_out = 1;
    return cover_finish(h, b, o, st, L);
}
