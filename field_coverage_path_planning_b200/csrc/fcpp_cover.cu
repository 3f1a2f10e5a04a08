// fcpp_cover.cu — coverage rasterisation: the W/2 round buffer of a polyline is scatter-written
// (atomicOr) into a 1-bit occupancy grid held in SHARED MEMORY tiles and popcounted; the grid
// never touches HBM.  One CTA per candidate plan.
//
// Reference code replaced ("mlp3" = multi_layer_planner_v3.py):
//   A10 corner-grid verification   mlp3:1426-1510 (per corner) and :1512-1578 (four corners)
//   A11 headland coverage rate     mlp3:1357-1371, recast as integer cell counts (north_star)
//
// Decision D5: all coordinates are snapped to the 1e-4 m lattice and a lattice point
// is covered iff dist²(point, segment) < r² EXACTLY (128-bit integers).
//
// Rasteriser (v3):
//  * per segment ("entry") the offset vector of the capsule's tangent lines (r*n) and the slope
//    dx/dy are computed once per plan and kept in shared memory;
//  * work items are (entry, 32-row chunk) pairs; a prefix sum over the entries and an
//    item -> entry table distribute them evenly over the 8 warps; a lane owns one grid row;
//  * the capsule's left and right boundaries are piecewise (arc of end A | offset line | arc of
//    end B): per row ONE sqrt or ONE multiply-add per side in FP64;
//  * the FP64 boundary is certified: only when it falls within 1e-5 cell of a lattice point is
//    that point decided by the exact integer predicate, so the result equals the per-cell brute
//    force of oracle/raster_oracle.c bit for bit.
#include "fcpp_internal.cuh"

namespace {

constexpr int T = FCPP_COVER_THREADS;
constexpr int NWARP = T / 32;
constexpr int TW = 8192;        // occupancy tile, 32-bit words (32 KB)
constexpr int ROWCAP = 1024;    // grid rows per tile (4 per thread)
constexpr int VPOLY_CAP = 256;  // verification polyline (15-pt arc + reverse fill), per corner
constexpr int ICAP = 4096;      // item -> entry table
constexpr double AMBIG = 1e-5;  // in lattice-index units (= 1e-6 m at h = 0.1 m)

struct Lattice {
    int64_t X0, Y0, H;  // lattice point (i, j) = (X0 + i*H, Y0 + j*H)
    int nx, ny;
    double X0d, Y0d, Hd, invH;
};

__device__ __forceinline__ void lattice_finish(Lattice &L)
{
    L.X0d = (double)L.X0;
    L.Y0d = (double)L.Y0;
    L.Hd = (double)L.H;
    L.invH = 1.0 / L.Hd;
}

// one raster target: a lattice, the window of its rows held in the tile, and where they sit
struct Target {
    Lattice L;
    int j0, nrows;  // lattice rows [j0, j0 + nrows) are resident
    int koff;       // tile-row index of lattice row j0
};

struct CoverFixed {
    CandRec rec;
    TrigTables tt;
    uint32_t tile[TW];
    int rbase[ROWCAP], rfa[ROWCAP], rfb[ROWCAP], rma[ROWCAP], rmb[ROWCAP];
    int scan[T];
    double qedge[2][4][3];  // field / main quad edges: ax, ay, k = ex/ey  (k unused when ey == 0)
    Target tg[4];
    int nrows, total_words;
    int cnt[8];
    unsigned long long acc[2];
    uint64_t bar;
};

// dynamic part, sized by the point capacity pc (>= longest polyline staged)
struct CoverDyn {
    int2 *pts;          // [pc]
    double *egeo;       // [pc][3]  ox, oy, k of entry e = segment pts[e] -> pts[e+1]
    int *estart;        // [pc] first resident row of the entry
    int *apre;          // [1024 + 1] exclusive (entry,row)-pair prefix over the ACTIVE entries of a batch
    uint16_t *act;      // [1024] active entries (batch-local index)
    uint8_t *ekind;     // [pc]
    uint16_t *item_first;  // [ICAP] active index holding the first pair of an item
};

__host__ __device__ inline size_t a16(size_t x) { return (x + 15) & ~size_t(15); }
__host__ __device__ inline size_t cover_smem_bytes(int pc)
{
    return a16(sizeof(CoverFixed)) + a16(sizeof(int2) * pc) + a16(sizeof(double) * 3 * pc) + a16(sizeof(int) * pc) +
           a16(sizeof(int) * 1025) + a16(sizeof(uint16_t) * 1024) + a16(pc) + a16(sizeof(uint16_t) * ICAP);
}
__device__ inline CoverDyn carve_dyn(unsigned char *base, int pc)
{
    CoverDyn d;
    size_t o = a16(sizeof(CoverFixed));
    d.pts = (int2 *)(base + o);
    o += a16(sizeof(int2) * pc);
    d.egeo = (double *)(base + o);
    o += a16(sizeof(double) * 3 * pc);
    d.estart = (int *)(base + o);
    o += a16(sizeof(int) * pc);
    d.apre = (int *)(base + o);
    o += a16(sizeof(int) * 1025);
    d.act = (uint16_t *)(base + o);
    o += a16(sizeof(uint16_t) * 1024);
    d.ekind = (uint8_t *)(base + o);
    o += a16(pc);
    d.item_first = (uint16_t *)(base + o);
    return d;
}

// exact floor(a / H) for |a| < 2^40, 0 < H <= 2^20 through FP64 + integer correction
__device__ __forceinline__ int64_t floor_div_fast(int64_t a, const Lattice &L)
{
    int64_t q = __double2ll_rd((double)a * L.invH);
    const int64_t rem = a - q * L.H;
    if (rem < 0) --q;
    if (rem >= L.H) ++q;
    return q;
}
__device__ __forceinline__ int64_t floor_div(int64_t a, int64_t b)  // b > 0
{
    int64_t q = a / b;
    if ((a % b != 0) && (a < 0)) --q;
    return q;
}
__device__ __forceinline__ int64_t ceil_div(int64_t a, int64_t b) { return -floor_div(-a, b); }

// exact: dist²((px,py), segment a-b) < r2 (oracle/raster_oracle.c near_segment)
__device__ __noinline__ bool near_seg(int64_t px, int64_t py, int64_t ax, int64_t ay, int64_t bx, int64_t by,
                                      int64_t r2)
{
    const int64_t dx = bx - ax, dy = by - ay;
    const int64_t wx = px - ax, wy = py - ay;
    const int64_t t = wx * dx + wy * dy;
    const int64_t dd = dx * dx + dy * dy;
    if (t <= 0) return wx * wx + wy * wy < r2;
    if (t >= dd) {
        const int64_t ux = px - bx, uy = py - by;
        return ux * ux + uy * uy < r2;
    }
    const int64_t cr = wx * dy - wy * dx;
    const uint64_t acr = cr < 0 ? (uint64_t)(-cr) : (uint64_t)cr;
    const uint64_t lhs_hi = __umul64hi(acr, acr), lhs_lo = acr * acr;
    const uint64_t rhs_hi = __umul64hi((uint64_t)r2, (uint64_t)dd), rhs_lo = (uint64_t)r2 * (uint64_t)dd;
    return lhs_hi < rhs_hi || (lhs_hi == rhs_hi && lhs_lo < rhs_lo);
}

// closed containment in a convex CCW quad with integer vertices
__device__ __noinline__ bool in_quad(const int64_t (*q)[2], int64_t px, int64_t py)
{
    bool in = true;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int k1 = (k + 1) & 3;
        const int64_t cr = (q[k1][0] - q[k][0]) * (py - q[k][1]) - (q[k1][1] - q[k][1]) * (px - q[k][0]);
        in = in && (cr >= 0);
    }
    return in;
}

// per-edge constants of a convex quad for the row-interval evaluation
__device__ void quad_edges_setup(const int64_t (*q)[2], double (*e)[3], int k)
{
    const int k1 = (k + 1) & 3;
    const double ex = (double)(q[k1][0] - q[k][0]), ey = (double)(q[k1][1] - q[k][1]);
    e[k][0] = (double)q[k][0];
    e[k][1] = (double)q[k][1];
    e[k][2] = (ey != 0.0) ? ex / ey : 0.0;
}

// inclusive lattice-index interval [a, b] of row cy inside the CLOSED convex quad (empty: a > b).
// FP64 boundary + exact test of a lattice point only when the boundary is within AMBIG of it.
__device__ void quad_row_interval(const int64_t (*q)[2], const double (*e)[3], int64_t cy, const Lattice &L,
                                  int &a, int &b)
{
    double lo = -1e300, hi = 1e300;
    bool empty = false;
    const double y = (double)cy;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int k1 = (k + 1) & 3;
        const int64_t eyi = q[k1][1] - q[k][1];
        const double x = e[k][0] + e[k][2] * (y - e[k][1]);
        if (eyi > 0)
            hi = fmin(hi, x);
        else if (eyi < 0)
            lo = fmax(lo, x);
        else if ((q[k1][0] - q[k][0]) * (cy - q[k][1]) < 0)
            empty = true;
    }
    if (empty || !(lo <= hi + 2.0)) {
        a = 0;
        b = -1;
        return;
    }
    const double tl = (lo - L.X0d) * L.invH, th = (hi - L.X0d) * L.invH;
    double fl = ceil(tl), fh = floor(th);  // closed interval: i >= tl, i <= th
    const double rl = rint(tl), rh = rint(th);
    if (fabs(tl - rl) < AMBIG) fl = in_quad(q, L.X0 + (int64_t)rl * L.H, cy) ? rl : rl + 1.0;
    if (fabs(th - rh) < AMBIG) fh = in_quad(q, L.X0 + (int64_t)rh * L.H, cy) ? rh : rh - 1.0;
    fl = fmax(fl, 0.0);
    fh = fmin(fh, (double)(L.nx - 1));
    if (!(fl <= fh)) {
        a = 0;
        b = -1;
        return;
    }
    a = (int)fl;
    b = (int)fh;
}

// Offset vector r*n (n = (dy,-dx)/len pointing to +x) and slope dx/dy of segment p->q, with the
// ends ordered so that the lower one comes first.  kind 0: general, 1: horizontal, 2: point.
__device__ __forceinline__ void entry_setup(int2 p, int2 q, double r, double *geo, uint8_t &kind)
{
    if (q.y < p.y || (q.y == p.y && q.x < p.x)) {
        const int2 t = p;
        p = q;
        q = t;
    }
    const double dx = (double)(q.x - p.x), dy = (double)(q.y - p.y);
    if (dy == 0.0) {
        kind = (dx == 0.0) ? 2 : 1;
        geo[0] = geo[1] = geo[2] = 0.0;
        return;
    }
    kind = 0;
    const double len = sqrt(dx * dx + dy * dy);
    geo[0] = r * dy / len;
    geo[1] = r * dx / len;
    geo[2] = dx / dy;
}

__device__ __forceinline__ void or_span(uint32_t *win, int w_first, int ia, int ib)
{
    const int w0 = ia >> 5, w1 = ib >> 5;
    for (int w = w0; w <= w1; ++w) {
        uint32_t m = 0xffffffffu;
        if (w == w0) m &= 0xffffffffu << (ia & 31);
        if (w == w1) m &= 0xffffffffu >> (31 - (ib & 31));
        atomicOr(&win[w - w_first], m);
    }
}

// words a row needs: left window [fa, ma-1] + right window [mb+1, fb] (or one window when the
// main interval is empty)
__device__ __forceinline__ int row_words(int fa, int fb, int ma, int mb, int &nl)
{
    nl = 0;
    if (fa > fb) return 0;
    if (ma > mb) {
        nl = (fb >> 5) - (fa >> 5) + 1;
        return nl;
    }
    int n = 0;
    if (fa <= ma - 1) {
        nl = ((ma - 1) >> 5) - (fa >> 5) + 1;
        n += nl;
    }
    if (mb + 1 <= fb) n += (fb >> 5) - ((mb + 1) >> 5) + 1;
    return n;
}

// block-wide inclusive scan of 4 ints per thread (entries 4*tid .. 4*tid+3); returns the total
__device__ int block_scan4(int (&v)[4], int *sh /*[T]*/)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v[1] += v[0];
    v[2] += v[1];
    v[3] += v[2];
    int inc = v[3];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += o;
    }
    __syncthreads();
    if (lane == 31) sh[warp] = inc;
    __syncthreads();
    int woff = 0, total = 0;
#pragma unroll
    for (int w = 0; w < NWARP; ++w) {
        const int x = sh[w];
        if (w < warp) woff += x;
        total += x;
    }
    const int off = woff + inc - v[3];
#pragma unroll
    for (int q = 0; q < 4; ++q) v[q] += off;
    return total;
}

// geometry of entries [e0, e0 + n): once per polyline
__device__ void setup_entries(const CoverDyn &d, int e0, int n, double rd)
{
    for (int e = e0 + threadIdx.x; e < e0 + n; e += T) {
        uint8_t kind;
        entry_setup(d.pts[e], d.pts[e + 1], rd, d.egeo + 3 * e, kind);
        d.ekind[e] = kind;
    }
}

// Rasterise the segments ("entries") e0 .. e0+n_ent-1 (entry e = pts[e] -> pts[e+1]; entries that
// join two different polylines must be masked by `tgt(e) < 0`) into the resident tile rows.
// Work = all (entry, row) pairs of the resident rows, flattened: lane p of item i owns pair
// 32*i + p, so no lane idles on short segments.
template <class TgFn>
__device__ void raster_entries(CoverFixed &s, const CoverDyn &d, int e0, int n_ent, TgFn tgt, int64_t r)
{
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const double rd = (double)r;
    const int64_t r2 = r * r;
    for (int eb = 0; eb < n_ent; eb += 4 * T) {  // batches of 1024 entries
        const int nb = min(4 * T, n_ent - eb);
        // ---- rows per entry (resident rows the capsule can touch); packed (active << 21 | rows) ----
        int rows[4], inc[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int le = tid * 4 + q;
            rows[q] = 0;
            if (le < nb) {
                const int e = e0 + eb + le;
                const int ti = tgt(e);
                if (ti >= 0) {
                    const int2 a = d.pts[e], b = d.pts[e + 1];
                    const Target &t = s.tg[ti];
                    const int64_t ylo = (int64_t)min(a.y, b.y) - r, yhi = (int64_t)max(a.y, b.y) + r;
                    int64_t jlo = floor_div_fast(ylo - t.L.Y0, t.L) + 1;      // cy > ylo
                    int64_t jhi = -floor_div_fast(-(yhi - t.L.Y0), t.L) - 1;  // cy < yhi
                    if (jlo < t.j0) jlo = t.j0;
                    if (jhi > t.j0 + t.nrows - 1) jhi = t.j0 + t.nrows - 1;
                    if (jhi >= jlo) rows[q] = (int)(jhi - jlo + 1);
                    d.estart[e] = (int)jlo;
                }
            }
            inc[q] = rows[q] ? ((1 << 21) | rows[q]) : 0;
        }
        const unsigned tot = (unsigned)block_scan4(inc, s.scan);
        const int n_act = (int)(tot >> 21), n_pairs = (int)(tot & 0x1fffffu);
        const int n_items = (n_pairs + 31) >> 5;
        const bool table = n_items <= ICAP;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (rows[q]) {
                const int a = (int)((unsigned)inc[q] >> 21) - 1;
                const int first = (int)((unsigned)inc[q] & 0x1fffffu) - rows[q];
                d.act[a] = (uint16_t)(tid * 4 + q);
                d.apre[a] = first;
                if (table)
                    for (int i = (first + 31) >> 5; i <= (first + rows[q] - 1) >> 5; ++i) d.item_first[i] = (uint16_t)a;
            }
        }
        if (tid == 0) d.apre[n_act] = n_pairs;
        __syncthreads();
        for (int item = warp; item < n_items; item += NWARP) {
            const int p = 32 * item + lane;
            if (p >= n_pairs) continue;
            int ai;
            if (table) {
                ai = d.item_first[item];
                while (p >= d.apre[ai + 1]) ++ai;
            } else {  // upper_bound over the pair prefix
                int lo = 0, hi = n_act - 1;
                while (lo < hi) {
                    const int mid = (lo + hi + 1) >> 1;
                    if (d.apre[mid] <= p)
                        lo = mid;
                    else
                        hi = mid - 1;
                }
                ai = lo;
            }
            const int e = e0 + eb + d.act[ai];
            const int j = d.estart[e] + (p - d.apre[ai]);
            const Target &t = s.tg[tgt(e)];
            const int k = (j - t.j0) + t.koff;
            const int fa = s.rfa[k], fb = s.rfb[k];
            if (fa > fb) continue;
            int2 a = d.pts[e], b = d.pts[e + 1];
            if (b.y < a.y || (b.y == a.y && b.x < a.x)) {
                const int2 tmp = a;
                a = b;
                b = tmp;
            }
            const double ax = (double)a.x, ay = (double)a.y, bx = (double)b.x, by = (double)b.y;
            const int64_t cy = t.L.Y0 + (int64_t)j * t.L.H;
            const double y = (double)cy;
            const double wa = y - ay, wb = y - by;
            const double r2d = rd * rd;
            double xl, xr;
            if (d.ekind[e] == 0) {
                const double ox = d.egeo[3 * e], oy = d.egeo[3 * e + 1], kk = d.egeo[3 * e + 2];
                const double yL0 = ay + oy, yL1 = by + oy, yR0 = ay - oy, yR1 = by - oy;
                if (y < yL0)
                    xl = ax - sqrt(fmax(r2d - wa * wa, 0.0));
                else if (y <= yL1)
                    xl = (ax - ox) + (y - yL0) * kk;
                else
                    xl = bx - sqrt(fmax(r2d - wb * wb, 0.0));
                if (y < yR0)
                    xr = ax + sqrt(fmax(r2d - wa * wa, 0.0));
                else if (y <= yR1)
                    xr = (ax + ox) + (y - yR0) * kk;
                else
                    xr = bx + sqrt(fmax(r2d - wb * wb, 0.0));
            } else {
                const double h = sqrt(fmax(r2d - wa * wa, 0.0));
                xl = ax - h;
                xr = bx + h;
            }
            // lattice indices strictly inside (xl, xr); ambiguous ends decided exactly
            const double tl = (xl - t.L.X0d) * t.L.invH, th = (xr - t.L.X0d) * t.L.invH;
            double fl = floor(tl) + 1.0, fh = ceil(th) - 1.0;
            const double rl = rint(tl), rh = rint(th);
            if (fabs(tl - rl) < AMBIG)
                fl = near_seg(t.L.X0 + (int64_t)rl * t.L.H, cy, a.x, a.y, b.x, b.y, r2) ? rl : rl + 1.0;
            if (fabs(th - rh) < AMBIG)
                fh = near_seg(t.L.X0 + (int64_t)rh * t.L.H, cy, a.x, a.y, b.x, b.y, r2) ? rh : rh - 1.0;
            fl = fmax(fl, (double)fa);
            fh = fmin(fh, (double)fb);
            if (!(fl <= fh)) continue;
            const int ia = (int)fl, ib = (int)fh;
            const int ma = s.rma[k], mb = s.rmb[k];
            uint32_t *base = s.tile + s.rbase[k];
            if (ma > mb) {
                or_span(base, fa >> 5, ia, ib);
            } else {
                int nl;
                row_words(fa, fb, ma, mb, nl);
                const int lh = ib < ma - 1 ? ib : ma - 1;
                if (ia <= lh) or_span(base, fa >> 5, ia, lh);
                const int rlo = ia > mb + 1 ? ia : mb + 1;
                if (rlo <= ib) or_span(base + nl, (mb + 1) >> 5, rlo, ib);
            }
        }
        __syncthreads();
    }
}

__device__ __forceinline__ int count_words(const uint32_t *w, int n)
{
    int c = 0;
    for (int i = threadIdx.x; i < n; i += T) c += __popc(w[i]);
    return c;
}

__global__ void __launch_bounds__(T, 2) cover_kernel(const fcpp_batch b, const CandRec *__restrict__ recs,
                                                     const TrigTables *__restrict__ trig,
                                                     fcpp_summary *__restrict__ summary, int pc)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CoverFixed &s = *reinterpret_cast<CoverFixed *>(smem_raw);
    const CoverDyn d = carve_dyn(smem_raw, pc);
    const int tid = threadIdx.x;
    const int64_t cand = blockIdx.x;
    fcpp_summary *sum = summary + cand;

    if (tid == 0) mbar_init(&s.bar, 1);
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&s.bar, sizeof(CandRec));
        bulk_g2s(&s.rec, recs + cand, sizeof(CandRec), &s.bar);
        s.acc[0] = s.acc[1] = 0ull;
    }
    if (tid < 8) s.cnt[tid] = 0;
    for (int k = tid; k < (int)(sizeof(TrigTables) / sizeof(double)); k += T)
        ((double *)&s.tt)[k] = ((const double *)trig)[k];
    mbar_wait(&s.bar, 0);
    __syncthreads();
    const CandRec &r = s.rec;
    if (r.status != 0 || r.n_total == 0) {
        if (tid == 0) {
            sum->cov_cells = sum->cov_total = 0;
            for (int k = 0; k < 4; ++k) sum->corner_before[k] = sum->corner_after[k] = 0;
        }
        return;
    }
    const double W = b.vehicle.working_width;
    const int64_t rq = qfix(W / 2);
    const double rd = (double)rq;
    int grid_err = 0;

    // =====================================================================================
    // A10: four verification corner windows (lattice POINTS, h = 0.1 m)
    // =====================================================================================
    {
        const double fl = b.field_extent[2 * r.field], fw = b.field_extent[2 * r.field + 1];
        const int g = r.corner_g;
        const int rw = (g + 31) >> 5;
        const bool okc = (g >= 1) && (rw <= TW) && (4 * VPOLY_CAP <= pc);
        if (!okc) grid_err = 1;
        const int rpt = okc ? min(ROWCAP, TW / rw) : 1;   // rows per tile
        const int group = (okc && 4 * g <= rpt) ? 4 : 1;  // corners per pass
        // snapped polylines of the four corners: 15-pt arc + reverse fill (mlp3:1531-1554);
        // corner ci occupies pts[ci*VPOLY_CAP ...]
        int nv[4];
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
            nv[ci] = min(r.vn_rev[ci], VPOLY_CAP - FCPP_CORNER_POINTS);
            if (r.vn_rev[ci] > VPOLY_CAP - FCPP_CORNER_POINTS) grid_err = 1;
            if (!okc) continue;
            const double qx = (ci == 0 || ci == 3) ? r.R : fl - r.R;  // mlp3:1531-1536
            const double qy = (ci == 0 || ci == 1) ? r.R : fw - r.R;
            for (int k = tid; k < FCPP_CORNER_POINTS + nv[ci]; k += T) {
                double x, y;
                if (k < FCPP_CORNER_POINTS) {
                    corner_arc_pt(s.tt, qx, qy, r.R, ci, k, x, y);
                } else {
                    const int m = k - FCPP_CORNER_POINTS;
                    const double len = r.vrev[ci][4];
                    const double tt_ = (m == r.vn_rev[ci] - 1) ? len : m * (len / (r.vn_rev[ci] - 1));
                    x = r.vrev[ci][0] + tt_ * r.vrev[ci][2];
                    y = r.vrev[ci][1] + tt_ * r.vrev[ci][3];
                }
                d.pts[ci * VPOLY_CAP + k] = make_int2((int)qfix(x), (int)qfix(y));
            }
        }
        __syncthreads();
        if (okc) {
#pragma unroll
            for (int ci = 0; ci < 4; ++ci) setup_entries(d, ci * VPOLY_CAP, FCPP_CORNER_POINTS + nv[ci] - 1, rd);
        }
        int before[4] = {0, 0, 0, 0}, after[4] = {0, 0, 0, 0};
        for (int c0 = 0; c0 < 4 && okc; c0 += group) {
            for (int j0 = 0; j0 < g; j0 += rpt) {
                const int nrows = (group == 4) ? g : min(rpt, g - j0);
                if (tid < group) {
                    const int ci = c0 + tid;
                    const double qx = (ci == 0 || ci == 3) ? r.R : fl - r.R;
                    const double qy = (ci == 0 || ci == 1) ? r.R : fw - r.R;
                    const double ox = (ci == 0 || ci == 3) ? qx : qx - 2 * r.R;  // mlp3:1461-1468
                    const double oy = (ci == 0 || ci == 1) ? qy : qy - 2 * r.R;
                    Target &t = s.tg[tid];
                    t.L.X0 = qfix(ox);
                    t.L.Y0 = qfix(oy);
                    t.L.H = qfix(FCPP_CORNER_GRID_H);
                    t.L.nx = g;
                    t.L.ny = g;
                    lattice_finish(t.L);
                    t.j0 = j0;
                    t.nrows = nrows;
                    t.koff = tid * nrows;
                }
                for (int k = tid; k < group * nrows; k += T) {
                    s.rbase[k] = k * rw;
                    s.rfa[k] = 0;
                    s.rfb[k] = g - 1;
                    s.rma[k] = 1;
                    s.rmb[k] = 0;
                }
                for (int w = tid; w < group * nrows * rw; w += T) s.tile[w] = 0u;
                if (tid < 8) s.cnt[tid] = 0;
                __syncthreads();
                // pass 1: the turn arcs (entries 0..13 of every corner of the group)
                {
                    auto tgt = [&](int e) {
                        const int c = e / VPOLY_CAP, k = e - c * VPOLY_CAP;
                        return (k < FCPP_CORNER_POINTS - 1) ? c - c0 : -1;
                    };
                    raster_entries(s, d, c0 * VPOLY_CAP, (group - 1) * VPOLY_CAP + FCPP_CORNER_POINTS - 1, tgt, rq);
                }
                for (int c = 0; c < group; ++c) {
                    const int n = count_words(s.tile + c * nrows * rw, nrows * rw);
                    if (n) atomicAdd(&s.cnt[c], n);
                }
                __syncthreads();
                // pass 2: the reverse fills (entries 15 .. 15+nv-2; entry 14 joins arc and fill and
                // is NOT part of either LineString, mlp3:1471 / :1487)
                {
                    auto tgt = [&](int e) {
                        const int c = e / VPOLY_CAP, k = e - c * VPOLY_CAP;
                        return (k >= FCPP_CORNER_POINTS && k < FCPP_CORNER_POINTS + nv[c] - 1) ? c - c0 : -1;
                    };
                    raster_entries(s, d, c0 * VPOLY_CAP, group * VPOLY_CAP - 1, tgt, rq);
                }
                for (int c = 0; c < group; ++c) {
                    const int n = count_words(s.tile + c * nrows * rw, nrows * rw);
                    if (n) atomicAdd(&s.cnt[4 + c], n);
                }
                __syncthreads();
                for (int c = 0; c < group; ++c) {
                    before[c0 + c] += s.cnt[c];
                    after[c0 + c] += s.cnt[4 + c];
                }
                __syncthreads();
                if (group == 4) break;
            }
        }
        if (tid == 0)
            for (int k = 0; k < 4; ++k) {
                sum->corner_before[k] = okc ? before[k] : 0;
                sum->corner_after[k] = okc ? after[k] : 0;
            }
    }

    // =====================================================================================
    // A11: headland band (lattice of cell CENTRES anchored at the field bbox minimum)
    // =====================================================================================
    {
        int64_t fq[4][2], mq[4][2];
        double bx0 = 1e300, by0 = 1e300, bx1 = -1e300, by1 = -1e300;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const double x = b.field_verts[(int64_t)r.field * 8 + 2 * k];
            const double y = b.field_verts[(int64_t)r.field * 8 + 2 * k + 1];
            fq[k][0] = qfix(x);
            fq[k][1] = qfix(y);
            mq[k][0] = qfix(r.main_quad[k][0]);
            mq[k][1] = qfix(r.main_quad[k][1]);
            bx0 = fmin(bx0, x);
            by0 = fmin(by0, y);
            bx1 = fmax(bx1, x);
            by1 = fmax(by1, y);
        }
        __syncthreads();
        if (tid < 4) quad_edges_setup(fq, s.qedge[0], tid);
        if (tid >= 4 && tid < 8) quad_edges_setup(mq, s.qedge[1], tid - 4);
        Lattice L;
        L.H = qfix(b.grid_h);
        const int64_t X0 = qfix(bx0), Y0 = qfix(by0);
        const int64_t nx64 = ceil_div(qfix(bx1) - X0, L.H), ny64 = ceil_div(qfix(by1) - Y0, L.H);
        L.X0 = X0 + L.H / 2;
        L.Y0 = Y0 + L.H / 2;
        L.nx = (int)nx64;
        L.ny = (int)ny64;
        lattice_finish(L);
        const int nh = r.n_head;
        const bool ok = (nh <= pc) && nx64 > 0 && ny64 > 0 && nx64 < (1ll << 30) && ny64 < (1ll << 30);
        if (!ok) grid_err = 1;
        if (ok) {
            for (int k = tid; k < nh; k += T) {
                double x, y;
                uint8_t c;
                gen_point(r, s.tt, W, r.n_main + k, x, y, c);
                d.pts[k] = make_int2((int)qfix(x), (int)qfix(y));
            }
            __syncthreads();
            setup_entries(d, 0, nh - 1, rd);
            unsigned long long my_total = 0ull, my_cov = 0ull;
            int j0 = 0;
            while (j0 < L.ny) {
                // --- how many rows to try: from the word count of the first row (uniform) ---
                int a0, b0, a1, b1, nl0;
                quad_row_interval(fq, s.qedge[0], L.Y0 + (int64_t)j0 * L.H, L, a0, b0);
                quad_row_interval(mq, s.qedge[1], L.Y0 + (int64_t)j0 * L.H, L, a1, b1);
                const int w0 = row_words(a0, b0, a1, b1, nl0);
                int rows_try = 2 * TW / (w0 > 0 ? w0 : 1);
                if (rows_try < 8) rows_try = 8;
                if (rows_try > ROWCAP) rows_try = ROWCAP;
                if (rows_try > L.ny - j0) rows_try = L.ny - j0;
                // --- row intervals + word counts: thread t owns rows 4t .. 4t+3 ---
                int wcnt[4], wsum[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int k = tid * 4 + q;
                    wcnt[q] = 0;
                    if (k < rows_try) {
                        int fa, fb, ma, mb, nl;
                        const int64_t cy = L.Y0 + (int64_t)(j0 + k) * L.H;
                        quad_row_interval(fq, s.qedge[0], cy, L, fa, fb);
                        quad_row_interval(mq, s.qedge[1], cy, L, ma, mb);
                        if (ma <= mb) {  // the inset lies inside the field: clamp defensively
                            if (ma < fa) ma = fa;
                            if (mb > fb) mb = fb;
                        }
                        s.rfa[k] = fa;
                        s.rfb[k] = fb;
                        s.rma[k] = ma;
                        s.rmb[k] = mb;
                        wcnt[q] = row_words(fa, fb, ma, mb, nl);
                    }
                    wsum[q] = wcnt[q];
                }
                if (tid == 0) {
                    s.nrows = 0;
                    s.total_words = 0;
                }
                block_scan4(wsum, s.scan);  // inclusive prefix of the row word counts (syncs inside)
                int fit_rows = 0, fit_words = 0;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int k = tid * 4 + q;
                    if (k < rows_try) {
                        s.rbase[k] = wsum[q] - wcnt[q];
                        if (wsum[q] <= TW) {
                            fit_rows = k + 1;
                            fit_words = wsum[q];
                        }
                    }
                }
                if (fit_rows) {
                    atomicMax(&s.nrows, fit_rows);
                    atomicMax(&s.total_words, fit_words);
                }
                __syncthreads();
                const int nrows = s.nrows, nwords = s.total_words;
                if (nrows == 0) {  // a single row does not fit the tile
                    grid_err = 1;
                    break;
                }
                for (int w = tid; w < nwords; w += T) s.tile[w] = 0u;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int k = tid * 4 + q;
                    if (k < nrows) {
                        const int fa = s.rfa[k], fb = s.rfb[k], ma = s.rma[k], mb = s.rmb[k];
                        if (fa <= fb) my_total += (unsigned long long)((fb - fa + 1) - (ma <= mb ? (mb - ma + 1) : 0));
                    }
                }
                if (tid == 0) {
                    Target &t = s.tg[0];
                    t.L = L;
                    t.j0 = j0;
                    t.nrows = nrows;
                    t.koff = 0;
                }
                __syncthreads();
                {
                    auto tgt = [&](int) { return 0; };
                    raster_entries(s, d, 0, nh - 1, tgt, rq);
                }
                my_cov += (unsigned long long)count_words(s.tile, nwords);
                __syncthreads();
                j0 += nrows;
            }
#pragma unroll
            for (int dd = 16; dd > 0; dd >>= 1) {
                my_total += __shfl_xor_sync(0xffffffffu, my_total, dd);
                my_cov += __shfl_xor_sync(0xffffffffu, my_cov, dd);
            }
            if ((tid & 31) == 0) {
                atomicAdd(&s.acc[0], my_total);
                atomicAdd(&s.acc[1], my_cov);
            }
            __syncthreads();
        }
        if (tid == 0) {
            sum->cov_total = (ok && !grid_err) ? (int64_t)s.acc[0] : 0;
            sum->cov_cells = (ok && !grid_err) ? (int64_t)s.acc[1] : 0;
            if (grid_err) sum->status |= FCPP_CAND_GRID_TOO_LARGE;
        }
    }
}

int cover_point_capacity(int max_head)
{
    int pc = max_head > 4 * VPOLY_CAP ? max_head : 4 * VPOLY_CAP;
    return (pc + 255) / 256 * 256 + 16;
}

}  // namespace

cudaError_t fcpp_launch_cover(fcpp_handle *h, const fcpp_batch &b, const fcpp_outputs &o, cudaStream_t st)
{
    if (b.n_cand == 0) return cudaSuccess;
    int pc = cover_point_capacity(h->cover_pcap);
    while (cover_smem_bytes(pc) > (size_t)h->max_smem_optin && pc > 4 * VPOLY_CAP + 16) pc -= 256;
    const size_t bytes = cover_smem_bytes(pc);
    cudaError_t e = cudaFuncSetAttribute(cover_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return e;
    cover_kernel<<<(unsigned)b.n_cand, T, bytes, st>>>(b, h->d_rec, h->d_trig, o.summary, pc);
    h->launches++;
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// generic window raster for the drop-in verify_corner_coverage_grid_based (mlp3:1426-1510)
// ---------------------------------------------------------------------------------------------
namespace {

constexpr int WIN_PC = 1024 + 16;

__global__ void __launch_bounds__(T) window_kernel(const double *__restrict__ path, int n_pts, double radius,
                                                   double ox, double oy, double hc, int g,
                                                   uint32_t *__restrict__ bits, int64_t *__restrict__ count)
{
    // one CTA; rows are processed in tiles of the shared occupancy buffer, then merged into the
    // caller's row-major g x g bit grid (bit j*g+i)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    CoverFixed &s = *reinterpret_cast<CoverFixed *>(smem_raw);
    const CoverDyn d = carve_dyn(smem_raw, WIN_PC);
    const int tid = threadIdx.x;
    Lattice L;
    L.X0 = qfix(ox);
    L.Y0 = qfix(oy);
    L.H = qfix(hc);
    L.nx = g;
    L.ny = g;
    lattice_finish(L);
    const int64_t rq = qfix(radius);
    const int rw = (g + 31) >> 5;
    const int rows_per_tile = min(ROWCAP, TW / rw);
    for (int j0 = 0; j0 < g; j0 += rows_per_tile) {
        const int nrows = min(rows_per_tile, g - j0);
        for (int k = tid; k < nrows; k += T) {
            s.rbase[k] = k * rw;
            s.rfa[k] = 0;
            s.rfb[k] = g - 1;
            s.rma[k] = 1;
            s.rmb[k] = 0;
        }
        for (int w = tid; w < nrows * rw; w += T) s.tile[w] = 0u;
        if (tid == 0) {
            Target &t = s.tg[0];
            t.L = L;
            t.j0 = j0;
            t.nrows = nrows;
            t.koff = 0;
        }
        __syncthreads();
        // polyline in chunks of 1024 points (chunks overlap by one point)
        for (int p0 = 0; p0 + 1 < n_pts; p0 += 1023) {
            const int np = min(1024, n_pts - p0);
            for (int k = tid; k < np; k += T)
                d.pts[k] = make_int2((int)qfix(path[2 * (p0 + k)]), (int)qfix(path[2 * (p0 + k) + 1]));
            __syncthreads();
            setup_entries(d, 0, np - 1, (double)rq);
            auto tgt = [&](int) { return 0; };
            raster_entries(s, d, 0, np - 1, tgt, rq);
        }
        // merge into the global bit grid
        for (int c = tid; c < nrows * g; c += T) {
            const int k = c / g, i = c - k * g;
            if ((s.tile[k * rw + (i >> 5)] >> (i & 31)) & 1u) {
                const int64_t bit = (int64_t)(j0 + k) * g + i;
                atomicOr(&bits[bit >> 5], 1u << (bit & 31));
            }
        }
        __syncthreads();
    }
    __threadfence();
    __syncthreads();
    // count all set bits of the (merged) grid
    if (tid == 0) s.acc[0] = 0ull;
    __syncthreads();
    const int64_t nw = ((int64_t)g * g + 31) >> 5;
    unsigned long long c = 0;
    for (int64_t w = tid; w < nw; w += T) c += __popc(bits[w]);
    atomicAdd(&s.acc[0], c);
    __syncthreads();
    if (tid == 0) *count = (int64_t)s.acc[0];
}

}  // namespace

cudaError_t fcpp_launch_raster_window(fcpp_handle *h, const double *d_path, int32_t n_pts, double radius,
                                      double ox, double oy, double hc, int32_t g, uint32_t *d_bits,
                                      int64_t *d_count, cudaStream_t st)
{
    const size_t bytes = cover_smem_bytes(WIN_PC);
    cudaError_t e = cudaFuncSetAttribute(window_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return e;
    window_kernel<<<1, T, bytes, st>>>(d_path, n_pts, radius, ox, oy, hc, g, d_bits, d_count);
    h->launches++;
    return cudaGetLastError();
}
