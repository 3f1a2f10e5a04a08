// fcpp_api.cu — the extern "C" boundary of libfcpp.so (declared in include/fcpp.h).
// No exceptions cross this boundary; every CUDA error is turned into a negative status and a
// message retrievable with fcpp_last_error().
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "fcpp_internal.cuh"

namespace {

int fail(fcpp_handle *h, int code, const char *fmt, ...)
{
    if (h) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(h->err, sizeof(h->err), fmt, ap);
        va_end(ap);
    }
    return code;
}

int cuda_fail(fcpp_handle *h, cudaError_t e, const char *what)
{
    return fail(h, FCPP_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
}

// numpy.linspace(0, stop, n): i*step for i < n-1, exactly `stop` at the end
void host_tables(TrigTables &t)
{
    const double pi = 3.141592653589793;
    for (int i = 0; i < FCPP_UTURN_POINTS; ++i) {
        const double a = (i == FCPP_UTURN_POINTS - 1) ? pi : i * (pi / (FCPP_UTURN_POINTS - 1));
        t.cos20[i] = cos(a);
        t.sin20[i] = sin(a);
    }
    for (int i = 0; i < FCPP_CORNER_POINTS; ++i) {
        const double a = (i == FCPP_CORNER_POINTS - 1) ? pi / 2 : i * ((pi / 2) / (FCPP_CORNER_POINTS - 1));
        t.cos15[i] = cos(a);
        t.sin15[i] = sin(a);
    }
}

int ensure_workspace(fcpp_handle *h, int64_t n_cand)
{
    if (n_cand > h->rec_cap) {
        if (h->d_rec) cudaFree(h->d_rec);
        h->d_rec = nullptr;
        h->rec_cap = 0;
        int64_t cap = n_cand + n_cand / 8 + 1024;
        cudaError_t e = cudaMalloc((void **)&h->d_rec, (size_t)cap * sizeof(CandRec));
        if (e != cudaSuccess) return cuda_fail(h, e, "cudaMalloc(candidate records)");
        h->rec_cap = cap;
    }
    const int64_t tiles = n_cand / 4096 + 2;
    if (tiles > h->scan_tmp_cap) {
        if (h->d_scan_tmp) cudaFree(h->d_scan_tmp);
    if (h->d_big) cudaFree(h->d_big);
        h->d_scan_tmp = nullptr;
        h->scan_tmp_cap = 0;
        cudaError_t e = cudaMalloc(&h->d_scan_tmp, (size_t)(tiles + 1024) * sizeof(int64_t));
        if (e != cudaSuccess) return cuda_fail(h, e, "cudaMalloc(scan workspace)");
        h->scan_tmp_cap = tiles + 1024;
    }
    return FCPP_OK;
}

int check_batch(fcpp_handle *h, const fcpp_batch *b)
{
    if (!h) return FCPP_ERR_INVALID;
    if (!b) return fail(h, FCPP_ERR_INVALID, "batch is NULL");
    if (b->n_cand < 0 || b->n_fields < 0) return fail(h, FCPP_ERR_INVALID, "negative sizes");
    if (b->n_cand > 0 &&
        (!b->field_verts || !b->field_extent || !b->field_flags || !b->cand_field || !b->cand_R || !b->cand_rot ||
         !b->cand_flags))
        return fail(h, FCPP_ERR_INVALID, "a required batch pointer is NULL");
    if (b->obs_poly_start && (!b->obs_vert_start || !b->obs_verts || !b->obs_moments))
        return fail(h, FCPP_ERR_INVALID, "obstacle tables are incomplete");
    if (!(b->vehicle.working_width > 0.0)) return fail(h, FCPP_ERR_INVALID, "working_width must be > 0");
    if (b->turn_model != FCPP_TURN_ARC && b->turn_model != FCPP_TURN_CLOTHOID)
        return fail(h, FCPP_ERR_INVALID, "turn_model must be 0 (arc) or 1 (clothoid)");
    if (b->turn_model == FCPP_TURN_CLOTHOID && !(b->clothoid_share > 0.0 && b->clothoid_share <= 1.0))
        return fail(h, FCPP_ERR_INVALID, "clothoid_share must be in (0, 1]");
    if (b->do_coverage) {
        const double hq = b->grid_h * FCPP_FIXED_UNIT;
        const long long H = llrint(hq);
        if (H < 2 || (H & 1) || fabs(hq - (double)H) > 1e-6)
            return fail(h, FCPP_ERR_INVALID, "grid_h must be an even multiple of 1e-4 m (got %g)", b->grid_h);
    }
    return FCPP_OK;
}

}  // namespace

extern "C" {

int fcpp_abi_version(void) { return FCPP_ABI_VERSION; }

int fcpp_create(int device, fcpp_handle **out)
{
    if (!out) return FCPP_ERR_INVALID;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) {
        cudaGetLastError();
        return FCPP_ERR_NO_DEVICE;
    }
    if (cudaSetDevice(device) != cudaSuccess) return FCPP_ERR_NO_DEVICE;
    fcpp_handle *h = (fcpp_handle *)calloc(1, sizeof(fcpp_handle));
    if (!h) return FCPP_ERR_INVALID;
    h->device = device;
    cudaDeviceGetAttribute(&h->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device);
    TrigTables t;
    host_tables(t);
    cudaError_t e = cudaMalloc((void **)&h->d_trig, sizeof(TrigTables));
    if (e == cudaSuccess) e = cudaMemcpy(h->d_trig, &t, sizeof(t), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->d_maxn, 2 * sizeof(int));
    if (e == cudaSuccess) e = cudaMallocHost((void **)&h->h_maxn, 2 * sizeof(int));
    if (e != cudaSuccess) {
        fcpp_destroy(h);
        return FCPP_ERR_CUDA;
    }
    *out = h;
    return FCPP_OK;
}

void fcpp_destroy(fcpp_handle *h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->d_trig) cudaFree(h->d_trig);
    if (h->d_rec) cudaFree(h->d_rec);
    if (h->d_scan_tmp) cudaFree(h->d_scan_tmp);
    if (h->d_big) cudaFree(h->d_big);
    if (h->d_maxn) cudaFree(h->d_maxn);
    if (h->h_maxn) cudaFreeHost(h->h_maxn);
    for (int k = 0; k < 4; ++k)
        if (h->ev[k]) cudaEventDestroy(h->ev[k]);
    free(h);
}

const char *fcpp_last_error(const fcpp_handle *h) { return h ? h->err : "invalid handle"; }

int64_t fcpp_launch_count(const fcpp_handle *h) { return h ? h->launches : 0; }

int32_t fcpp_last_max_points(const fcpp_handle *h) { return h ? h->last_maxn : 0; }
int32_t fcpp_last_max_head_points(const fcpp_handle *h) { return h ? h->last_maxhead : 0; }

int fcpp_set_profiling(fcpp_handle *h, int on)
{
    if (!h) return FCPP_ERR_INVALID;
    cudaSetDevice(h->device);
    if (on && !h->ev[0]) {
        for (int k = 0; k < 4; ++k)
            if (cudaEventCreate(&h->ev[k]) != cudaSuccess) return cuda_fail(h, cudaGetLastError(), "cudaEventCreate");
    }
    h->profiling = on != 0;
    return FCPP_OK;
}

int fcpp_kernel_times(fcpp_handle *h, float *ms3)
{
    if (!h || !ms3 || !h->ev[0]) return fail(h, FCPP_ERR_INVALID, "profiling is off");
    for (int k = 0; k < 3; ++k) {
        cudaError_t e = cudaEventElapsedTime(&ms3[k], h->ev[k], h->ev[k + 1]);
        if (e != cudaSuccess) return cuda_fail(h, e, "cudaEventElapsedTime");
    }
    return FCPP_OK;
}

int fcpp_set_trig_tables(fcpp_handle *h, const double *cos20, const double *sin20, const double *cos15,
                         const double *sin15)
{
    if (!h || !cos20 || !sin20 || !cos15 || !sin15) return fail(h, FCPP_ERR_INVALID, "NULL table");
    TrigTables t;
    memcpy(t.cos20, cos20, sizeof(t.cos20));
    memcpy(t.sin20, sin20, sizeof(t.sin20));
    memcpy(t.cos15, cos15, sizeof(t.cos15));
    memcpy(t.sin15, sin15, sizeof(t.sin15));
    cudaSetDevice(h->device);
    cudaError_t e = cudaMemcpy(h->d_trig, &t, sizeof(t), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return cuda_fail(h, e, "fcpp_set_trig_tables");
    return FCPP_OK;
}

int fcpp_layout(fcpp_handle *h, const fcpp_batch *batch, int32_t *d_n_pts, int64_t *d_offsets, void *stream)
{
    int rc = check_batch(h, batch);
    if (rc) return rc;
    cudaSetDevice(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    rc = ensure_workspace(h, batch->n_cand);
    if (rc) return rc;
    cudaError_t e = fcpp_launch_layout(h, *batch, d_n_pts, d_offsets, st);
    if (e != cudaSuccess) return cuda_fail(h, e, "layout kernel");
    // the plan kernel sizes its shared-memory staging by the longest plan of the batch
    int maxn = batch->max_points_hint, maxhead = batch->max_head_points_hint;
    if (maxn <= 0 || maxhead <= 0) {
        h->h_maxn[0] = h->h_maxn[1] = 0;
        if (batch->n_cand > 0) {
            e = cudaMemcpyAsync(h->h_maxn, h->d_maxn, 2 * sizeof(int), cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess) e = cudaStreamSynchronize(st);
            if (e != cudaSuccess) return cuda_fail(h, e, "layout readback");
        }
        maxn = h->h_maxn[0];
        maxhead = h->h_maxn[1];
        h->last_maxn = maxn;
        h->last_maxhead = maxhead;
    }
    h->cover_pcap = maxhead;
    int want = (maxn + 63) / 64 * 64;
    if (want < 256) want = 256;
    h->plan_ncap_hint = want;
    h->layout_valid = true;
    h->layout_ncand = batch->n_cand;
    return FCPP_OK;
}

int fcpp_plan_batch(fcpp_handle *h, const fcpp_batch *batch, const fcpp_outputs *out, void *stream)
{
    int rc = check_batch(h, batch);
    if (rc) return rc;
    if (!out || !out->summary) return fail(h, FCPP_ERR_INVALID, "outputs.summary is required");
    const bool paths = out->path_xy || out->speeds_kmh || out->curvature;
    if (paths && !out->offsets)
        return fail(h, FCPP_ERR_INVALID, "materialised paths need outputs.offsets from fcpp_layout");
    cudaSetDevice(h->device);
    cudaStream_t st = (cudaStream_t)stream;
    const bool prof = h->profiling;
    if (prof) cudaEventRecord(h->ev[0], st);
    if (out->offsets) {
        if (!h->layout_valid || h->layout_ncand != batch->n_cand)
            return fail(h, FCPP_ERR_INVALID, "fcpp_layout must be called for this batch before fcpp_plan_batch");
    } else {
        rc = fcpp_layout(h, batch, nullptr, nullptr, stream);
        if (rc) return rc;
    }
    if (prof) cudaEventRecord(h->ev[1], st);
    int ncap = 0;
    cudaError_t e = fcpp_launch_plan(h, *batch, *out, st, &ncap);
    if (e != cudaSuccess) return cuda_fail(h, e, "plan kernel");
    if (prof) cudaEventRecord(h->ev[2], st);
    if (batch->do_coverage) {
        e = fcpp_launch_cover(h, *batch, *out, st);
        if (e != cudaSuccess) return cuda_fail(h, e, "coverage kernel");
    }
    if (prof) cudaEventRecord(h->ev[3], st);
    h->layout_valid = false;  // one layout per plan call
    return FCPP_OK;
}

int fcpp_field_argmin(fcpp_handle *h, const fcpp_summary *d_summary, const int32_t *d_cand_field, int64_t n_cand,
                      int32_t n_fields, int cost_kind, int64_t cand_base, double *d_best_cost,
                      int64_t *d_best_cand, void *stream)
{
    if (!h) return FCPP_ERR_INVALID;
    if (n_cand < 0 || n_fields < 0 || !d_best_cost || !d_best_cand || (n_cand > 0 && (!d_summary || !d_cand_field)))
        return fail(h, FCPP_ERR_INVALID, "fcpp_field_argmin: bad argument");
    cudaSetDevice(h->device);
    cudaError_t e = fcpp_launch_argmin(h, d_summary, d_cand_field, n_cand, n_fields, cost_kind, cand_base,
                                       d_best_cost, d_best_cand, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(h, e, "argmin kernels");
    return FCPP_OK;
}

int fcpp_speed_verify(fcpp_handle *h, const fcpp_vehicle *veh, const double *d_path_xy, const double *d_speeds_in,
                      const int64_t *d_offsets, int64_t n_paths, int64_t max_path_len, int do_speed_plan,
                      double *d_speeds_out, double *d_curvature, fcpp_summary *d_summary, void *stream)
{
    if (!h) return FCPP_ERR_INVALID;
    if (!veh || n_paths < 0 || (n_paths > 0 && (!d_path_xy || !d_speeds_in || !d_offsets)))
        return fail(h, FCPP_ERR_INVALID, "fcpp_speed_verify: bad argument");
    cudaSetDevice(h->device);
    cudaError_t e = fcpp_launch_speed_verify(h, *veh, d_path_xy, d_speeds_in, d_offsets, n_paths, max_path_len,
                                             do_speed_plan, d_speeds_out, d_curvature, d_summary,
                                             (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(h, e, "speed/verify kernel");
    return FCPP_OK;
}

int fcpp_raster_window(fcpp_handle *h, const double *d_path_xy, int32_t n_pts, double radius, double origin_x,
                       double origin_y, double h_cell, int32_t g, uint32_t *d_bits, int64_t *d_count, void *stream)
{
    if (!h) return FCPP_ERR_INVALID;
    if (!d_path_xy || n_pts < 0 || g < 1 || g > 8192 || !d_bits || !d_count || !(h_cell > 0) || !(radius >= 0))
        return fail(h, FCPP_ERR_INVALID, "fcpp_raster_window: bad argument");
    if (llrint(h_cell * FCPP_FIXED_UNIT) < 1) return fail(h, FCPP_ERR_INVALID, "h_cell below the 1e-4 m lattice");
    cudaSetDevice(h->device);
    cudaError_t e = fcpp_launch_raster_window(h, d_path_xy, n_pts, radius, origin_x, origin_y, h_cell, g, d_bits,
                                              d_count, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(h, e, "window raster kernel");
    return FCPP_OK;
}

int fcpp_tour_lengths(fcpp_handle *h, const double *d_D, int32_t n, const int32_t *d_pop, int64_t pop_size,
                      double *d_out, double *d_fitness, void *stream)
{
    if (!h) return FCPP_ERR_INVALID;
    if (n < 1 || pop_size < 0 || !d_out || (pop_size > 0 && (!d_D || !d_pop)))
        return fail(h, FCPP_ERR_INVALID, "fcpp_tour_lengths: bad argument");
    cudaSetDevice(h->device);
    cudaError_t e = fcpp_launch_tours(h, d_D, n, d_pop, pop_size, d_out, d_fitness, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(h, e, "tour kernel");
    return FCPP_OK;
}

}  // extern "C"
