/* Brute-force integer rasterisation oracle (TEST INFRASTRUCTURE ONLY — never linked into
 * the product library; see oracle/README.md).
 *
 * Restates, in exact integer arithmetic, the two coverage measures of the reference:
 *   A10  verify_corner_coverage_grid_based   multi_layer_planner_v3.py:1426-1510
 *        (lattice points origin + (i*h, j*h), `buffer(path, W/2).contains(point)`)
 *   A11  _calculate_coverage_rate            multi_layer_planner_v3.py:1357-1371
 *        (recast by north_star as covered cells / cells of the headland band)
 * under decision D5 of SURVEY.md §8(c): every coordinate is snapped to the 1e-4 m lattice
 * (q(x) = llrint(x*1e4), done by the caller) and "inside the round buffer" is the exact
 * predicate dist²(point, segment) < r² evaluated with 128-bit integers (D2, strict).
 *
 * The method is deliberately the dumbest correct one (every cell of every segment's
 * bounding box is tested) so that it is independent of the span-based CUDA rasteriser.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef __int128 i128;

/* exact: dist²((px,py), segment a-b) < r2 ? */
static int near_segment(int64_t px, int64_t py, int64_t ax, int64_t ay, int64_t bx, int64_t by,
                        int64_t r2)
{
    int64_t dx = bx - ax, dy = by - ay;
    int64_t wx = px - ax, wy = py - ay;
    int64_t t = wx * dx + wy * dy;
    int64_t dd = dx * dx + dy * dy;
    if (t <= 0) return wx * wx + wy * wy < r2;
    if (t >= dd) {
        int64_t ux = px - bx, uy = py - by;
        return ux * ux + uy * uy < r2;
    }
    int64_t cr = wx * dy - wy * dx;
    return (i128)cr * cr < (i128)r2 * dd;
}

static int64_t floor_div(int64_t a, int64_t b)
{
    int64_t q = a / b, r = a % b;
    return (r != 0 && ((r < 0) != (b < 0))) ? q - 1 : q;
}

/* OR the cover of polyline pts[0..n) (radius r) into `bits` (nx*ny lattice, row-major,
 * lattice point (i,j) = (X0 + i*H, Y0 + j*H)).  Returns the number of set bits afterwards. */
int64_t fcpo_raster(const int64_t *pts, int n, int64_t r, int64_t X0, int64_t Y0, int64_t H,
                    int64_t nx, int64_t ny, uint8_t *bits)
{
    int64_t r2 = r * r;
    for (int s = 0; s + 1 < n; ++s) {
        int64_t ax = pts[2 * s], ay = pts[2 * s + 1];
        int64_t bx = pts[2 * s + 2], by = pts[2 * s + 3];
        int64_t lox = (ax < bx ? ax : bx) - r, hix = (ax > bx ? ax : bx) + r;
        int64_t loy = (ay < by ? ay : by) - r, hiy = (ay > by ? ay : by) + r;
        int64_t i0 = floor_div(lox - X0, H), i1 = floor_div(hix - X0, H) + 1;
        int64_t j0 = floor_div(loy - Y0, H), j1 = floor_div(hiy - Y0, H) + 1;
        if (i0 < 0) i0 = 0;
        if (j0 < 0) j0 = 0;
        if (i1 > nx - 1) i1 = nx - 1;
        if (j1 > ny - 1) j1 = ny - 1;
        for (int64_t j = j0; j <= j1; ++j)
            for (int64_t i = i0; i <= i1; ++i)
                if (near_segment(X0 + i * H, Y0 + j * H, ax, ay, bx, by, r2)) {
                    int64_t c = j * nx + i;
                    bits[c >> 3] |= (uint8_t)(1u << (c & 7));
                }
    }
    int64_t cnt = 0, nb = (nx * ny + 7) >> 3;
    for (int64_t k = 0; k < nb; ++k) cnt += __builtin_popcount(bits[k]);
    return cnt;
}

static int in_convex(const int64_t *poly, int nv, int64_t px, int64_t py)
{
    for (int k = 0; k < nv; ++k) {
        int64_t ax = poly[2 * k], ay = poly[2 * k + 1];
        int64_t bx = poly[2 * ((k + 1) % nv)], by = poly[2 * ((k + 1) % nv) + 1];
        int64_t cr = (bx - ax) * (py - ay) - (by - ay) * (px - ax);
        if (cr < 0) return 0;
    }
    return 1;
}

/* Headland-band coverage: lattice of cell CENTRES (Xc0 + i*H, Yc0 + j*H); a cell belongs to
 * the band iff its centre is inside-or-on the convex CCW `field` and NOT inside-or-on the
 * convex CCW `main_` polygon (main_ == NULL: band = whole field, mlp3:873-875).
 * out[0] = band cells, out[1] = band cells covered by the polyline buffer. */
int fcpo_band(const int64_t *field, int nf, const int64_t *main_, int nm, const int64_t *pts,
              int n, int64_t r, int64_t Xc0, int64_t Yc0, int64_t H, int64_t nx, int64_t ny,
              int64_t *out)
{
    int64_t nb = (nx * ny + 7) >> 3;
    uint8_t *bits = (uint8_t *)calloc((size_t)nb, 1);
    if (!bits) return -1;
    fcpo_raster(pts, n, r, Xc0, Yc0, H, nx, ny, bits);
    int64_t total = 0, cov = 0;
    for (int64_t j = 0; j < ny; ++j)
        for (int64_t i = 0; i < nx; ++i) {
            int64_t px = Xc0 + i * H, py = Yc0 + j * H;
            if (!in_convex(field, nf, px, py)) continue;
            if (main_ && in_convex(main_, nm, px, py)) continue;
            ++total;
            int64_t c = j * nx + i;
            cov += (bits[c >> 3] >> (c & 7)) & 1;
        }
    free(bits);
    out[0] = total;
    out[1] = cov;
    return 0;
}

/* Closed-tour length, sequential left-to-right FP64 sum (genetic_algorithm_solver.py:174-181). */
void fcpo_tour_lengths(const double *D, int n, const int32_t *pop, int64_t pop_size, double *out)
{
    for (int64_t p = 0; p < pop_size; ++p) {
        const int32_t *r = pop + p * n;
        double s = 0.0;
        for (int i = 0; i < n; ++i) s += D[(int64_t)r[i] * n + r[(i + 1) % n]];
        out[p] = s;
    }
}
