/* TEST INFRASTRUCTURE ONLY (oracle/): CPU restatement of what vessel_graph_generation/tree2img.py:12-114
 * (rasterize_forest) makes matplotlib's Agg backend compute for one LineCollection of round-capped,
 * anti-aliased white strokes on an opaque black canvas.
 *
 * The arithmetic lives in third-party code that is NOT under /root/reference (matplotlib >= 3.10.3 per the reference's
 * pyproject.toml:10, which vendors Anti-Grain Geometry 2.4; neither is installed in the build container).  This file restates
 * the published algorithms of the stages a 2-vertex path of a LineCollection passes through
 *   matplotlib src/_backend_agg.h   RendererAgg::_draw_path_collection_generic / _draw_path
 *   matplotlib src/path_converters.h PathClipper (centre line clipped to [-1, W+1] x [-1, H+1]), PathSnapper (auto)
 *   agg  agg_vcgen_stroke / agg_math_stroke.h   calc_cap (round caps = inscribed polygon, 1/8 px tolerance)
 *   agg  agg_rasterizer_sl_clip.h               rasterizer_sl_clip<ras_conv_dbl> (clip box = canvas)
 *   agg  agg_rasterizer_cells_aa.h              line / render_hline (24.8 fixed point, exact cover/area cells)
 *   agg  agg_rasterizer_scanline_aa.h           sweep_scanline / calculate_alpha (non-zero winding, 8-bit coverage)
 *   matplotlib src/agg_workaround.h             fixed_blender_rgba_plain::blend_pix (8-bit "over" on an opaque pixel)
 * from memory of those sources; PARITY UNPINNED against matplotlib itself.  What pins it: the 500 csv -> label pairs the
 * reference ships (datasets/vessel_graphs/X.csv, datasets/labels/X.png = this gray image through PIL's Floyd-Steinberg
 * convert("1"), visualize_vessel_graphs.py:95-101) -- dithering is chaotic in the gray values, so the fraction of label pixels
 * reproduced is a sharp detector of any deviation (tests/test_raster2d_labels.py records it).
 *
 * Pixel (row, col) <-> device coordinates: x = pos[ax1] * W, y = pos[ax0] * H (y axis inverted by ax.invert_yaxis(), then
 * flipped again by the renderer: rows grow with pos[ax0]). */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SUBPIXEL_SHIFT 8
#define SUBPIXEL_SCALE 256
#define SUBPIXEL_MASK 255

typedef struct {
    int W, H;
    int bx0, by0, bw, bh; /* cell window of the current stroke (columns bx0 .. bx0+bw-1 may include column W) */
    int *cover, *area;    /* bw * bh */
    int cx, cy;           /* current cell */
    double clip_x1, clip_y1, clip_x2, clip_y2;
    double m_x1, m_y1;
    unsigned m_f1;
    double start_x, start_y;
    int status; /* 0 initial, 1 move_to, 2 line_to, 3 closed */
    int minx, miny, maxx, maxy;
} Ras;

static int iround(double v) { return (int)((v < 0.0) ? v - 0.5 : v + 0.5); }

static void cell_add(Ras* r, int ex, int ey, int cover, int area) {
    if (ey < r->by0 || ey >= r->by0 + r->bh || ex < r->bx0 || ex >= r->bx0 + r->bw) return; /* outside the window: never for clipped input */
    const size_t k = (size_t)(ey - r->by0) * r->bw + (ex - r->bx0);
    r->cover[k] += cover;
    r->area[k] += area;
    if (ex < r->minx) r->minx = ex;
    if (ex > r->maxx) r->maxx = ex;
    if (ey < r->miny) r->miny = ey;
    if (ey > r->maxy) r->maxy = ey;
}

/* agg_rasterizer_cells_aa.h render_hline: cells are (cover, area) accumulators, so adding to the dense window replaces
 * the m_curr_cell bookkeeping */
static void render_hline(Ras* r, int ey, int x1, int y1, int x2, int y2) {
    int ex1 = x1 >> SUBPIXEL_SHIFT, ex2 = x2 >> SUBPIXEL_SHIFT;
    int fx1 = x1 & SUBPIXEL_MASK, fx2 = x2 & SUBPIXEL_MASK;
    int delta, p, first, dx, incr, lift, mod, rem;
    if (y1 == y2) return;
    if (ex1 == ex2) {
        delta = y2 - y1;
        cell_add(r, ex1, ey, delta, (fx1 + fx2) * delta);
        return;
    }
    p = (SUBPIXEL_SCALE - fx1) * (y2 - y1);
    first = SUBPIXEL_SCALE;
    incr = 1;
    dx = x2 - x1;
    if (dx < 0) { p = fx1 * (y2 - y1); first = 0; incr = -1; dx = -dx; }
    delta = p / dx;
    mod = p % dx;
    if (mod < 0) { delta--; mod += dx; }
    cell_add(r, ex1, ey, delta, (fx1 + first) * delta);
    ex1 += incr;
    y1 += delta;
    if (ex1 != ex2) {
        p = SUBPIXEL_SCALE * (y2 - y1 + delta);
        lift = p / dx;
        rem = p % dx;
        if (rem < 0) { lift--; rem += dx; }
        mod -= dx;
        while (ex1 != ex2) {
            delta = lift;
            mod += rem;
            if (mod >= 0) { mod -= dx; delta++; }
            cell_add(r, ex1, ey, delta, SUBPIXEL_SCALE * delta);
            y1 += delta;
            ex1 += incr;
        }
    }
    delta = y2 - y1;
    cell_add(r, ex1, ey, delta, (fx2 + SUBPIXEL_SCALE - first) * delta);
}

/* agg_rasterizer_cells_aa.h line */
static void ras_line(Ras* r, int x1, int y1, int x2, int y2) {
    const int dx_limit = 16384 << SUBPIXEL_SHIFT;
    int dx = x2 - x1;
    if (dx >= dx_limit || dx <= -dx_limit) {
        int cx = (x1 + x2) >> 1, cy = (y1 + y2) >> 1;
        ras_line(r, x1, y1, cx, cy);
        ras_line(r, cx, cy, x2, y2);
        return;
    }
    int dy = y2 - y1;
    int ey1 = y1 >> SUBPIXEL_SHIFT, ey2 = y2 >> SUBPIXEL_SHIFT;
    int fy1 = y1 & SUBPIXEL_MASK, fy2 = y2 & SUBPIXEL_MASK;
    int x_from, x_to, p, rem, mod, lift, delta, first, incr;
    if (ey1 == ey2) { render_hline(r, ey1, x1, fy1, x2, fy2); return; }
    incr = 1;
    if (dx == 0) {
        int ex = x1 >> SUBPIXEL_SHIFT;
        int two_fx = (x1 - (ex << SUBPIXEL_SHIFT)) << 1;
        first = SUBPIXEL_SCALE;
        if (dy < 0) { first = 0; incr = -1; }
        delta = first - fy1;
        cell_add(r, ex, ey1, delta, two_fx * delta);
        ey1 += incr;
        delta = first + first - SUBPIXEL_SCALE;
        while (ey1 != ey2) {
            cell_add(r, ex, ey1, delta, two_fx * delta);
            ey1 += incr;
        }
        delta = fy2 - SUBPIXEL_SCALE + first;
        cell_add(r, ex, ey1, delta, two_fx * delta);
        return;
    }
    p = (SUBPIXEL_SCALE - fy1) * dx;
    first = SUBPIXEL_SCALE;
    if (dy < 0) { p = fy1 * dx; first = 0; incr = -1; dy = -dy; }
    delta = p / dy;
    mod = p % dy;
    if (mod < 0) { delta--; mod += dy; }
    x_from = x1 + delta;
    render_hline(r, ey1, x1, fy1, x_from, first);
    ey1 += incr;
    if (ey1 != ey2) {
        p = SUBPIXEL_SCALE * dx;
        lift = p / dy;
        rem = p % dy;
        if (rem < 0) { lift--; rem += dy; }
        mod -= dy;
        while (ey1 != ey2) {
            delta = lift;
            mod += rem;
            if (mod >= 0) { mod -= dy; delta++; }
            x_to = x_from + delta;
            render_hline(r, ey1, x_from, SUBPIXEL_SCALE - first, x_to, first);
            x_from = x_to;
            ey1 += incr;
        }
    }
    render_hline(r, ey1, x_from, SUBPIXEL_SCALE - first, x2, fy2);
}

/* agg_rasterizer_sl_clip.h, rasterizer_sl_clip<ras_conv_dbl> */
static unsigned clip_flags(const Ras* r, double x, double y) {
    return (unsigned)(x > r->clip_x2) | ((unsigned)(y > r->clip_y2) << 1) | ((unsigned)(x < r->clip_x1) << 2) | ((unsigned)(y < r->clip_y1) << 3);
}
static unsigned clip_flags_y(const Ras* r, double y) { return ((unsigned)(y > r->clip_y2) << 1) | ((unsigned)(y < r->clip_y1) << 3); }
static int xi(double v) { return iround(v * SUBPIXEL_SCALE); }

static void line_clip_y(Ras* r, double x1, double y1, double x2, double y2, unsigned f1, unsigned f2) {
    f1 &= 10; f2 &= 10;
    if ((f1 | f2) == 0) { ras_line(r, xi(x1), xi(y1), xi(x2), xi(y2)); return; }
    if (f1 == f2) return;
    double tx1 = x1, ty1 = y1, tx2 = x2, ty2 = y2;
    if (f1 & 8) { tx1 = x1 + (r->clip_y1 - y1) * (x2 - x1) / (y2 - y1); ty1 = r->clip_y1; }
    if (f1 & 2) { tx1 = x1 + (r->clip_y2 - y1) * (x2 - x1) / (y2 - y1); ty1 = r->clip_y2; }
    if (f2 & 8) { tx2 = x1 + (r->clip_y1 - y1) * (x2 - x1) / (y2 - y1); ty2 = r->clip_y1; }
    if (f2 & 2) { tx2 = x1 + (r->clip_y2 - y1) * (x2 - x1) / (y2 - y1); ty2 = r->clip_y2; }
    ras_line(r, xi(tx1), xi(ty1), xi(tx2), xi(ty2));
}

static void clip_move_to(Ras* r, double x, double y) { r->m_x1 = x; r->m_y1 = y; r->m_f1 = clip_flags(r, x, y); }

static void clip_line_to(Ras* r, double x2, double y2) {
    unsigned f2 = clip_flags(r, x2, y2);
    if ((r->m_f1 & 10) == (f2 & 10) && (r->m_f1 & 10) != 0) { r->m_x1 = x2; r->m_y1 = y2; r->m_f1 = f2; return; }
    double x1 = r->m_x1, y1 = r->m_y1, y3, y4;
    unsigned f1 = r->m_f1, f3, f4;
    const double cx1 = r->clip_x1, cx2 = r->clip_x2;
    switch (((f1 & 5) << 1) | (f2 & 5)) {
    case 0: line_clip_y(r, x1, y1, x2, y2, f1, f2); break;
    case 1:
        y3 = y1 + (cx2 - x1) * (y2 - y1) / (x2 - x1); f3 = clip_flags_y(r, y3);
        line_clip_y(r, x1, y1, cx2, y3, f1, f3); line_clip_y(r, cx2, y3, cx2, y2, f3, f2); break;
    case 2:
        y3 = y1 + (cx2 - x1) * (y2 - y1) / (x2 - x1); f3 = clip_flags_y(r, y3);
        line_clip_y(r, cx2, y1, cx2, y3, f1, f3); line_clip_y(r, cx2, y3, x2, y2, f3, f2); break;
    case 3: line_clip_y(r, cx2, y1, cx2, y2, f1, f2); break;
    case 4:
        y3 = y1 + (cx1 - x1) * (y2 - y1) / (x2 - x1); f3 = clip_flags_y(r, y3);
        line_clip_y(r, x1, y1, cx1, y3, f1, f3); line_clip_y(r, cx1, y3, cx1, y2, f3, f2); break;
    case 6:
        y3 = y1 + (cx2 - x1) * (y2 - y1) / (x2 - x1); y4 = y1 + (cx1 - x1) * (y2 - y1) / (x2 - x1);
        f3 = clip_flags_y(r, y3); f4 = clip_flags_y(r, y4);
        line_clip_y(r, cx2, y1, cx2, y3, f1, f3); line_clip_y(r, cx2, y3, cx1, y4, f3, f4); line_clip_y(r, cx1, y4, cx1, y2, f4, f2); break;
    case 8:
        y3 = y1 + (cx1 - x1) * (y2 - y1) / (x2 - x1); f3 = clip_flags_y(r, y3);
        line_clip_y(r, cx1, y1, cx1, y3, f1, f3); line_clip_y(r, cx1, y3, x2, y2, f3, f2); break;
    case 9:
        y3 = y1 + (cx1 - x1) * (y2 - y1) / (x2 - x1); y4 = y1 + (cx2 - x1) * (y2 - y1) / (x2 - x1);
        f3 = clip_flags_y(r, y3); f4 = clip_flags_y(r, y4);
        line_clip_y(r, cx1, y1, cx1, y3, f1, f3); line_clip_y(r, cx1, y3, cx2, y4, f3, f4); line_clip_y(r, cx2, y4, cx2, y2, f4, f2); break;
    case 12: line_clip_y(r, cx1, y1, cx1, y2, f1, f2); break;
    }
    r->m_f1 = f2;
    r->m_x1 = x2; r->m_y1 = y2;
}

/* agg_clip_liang_barsky.h clip_line_segment (used by matplotlib's PathClipper on the centre line) */
static unsigned lb_flags(double x, double y, const double* b) {
    return (unsigned)(x > b[2]) | ((unsigned)(y > b[3]) << 1) | ((unsigned)(x < b[0]) << 2) | ((unsigned)(y < b[1]) << 3);
}
static int lb_move_point(double x1, double y1, double x2, double y2, const double* b, double* x, double* y, unsigned flags) {
    double bound;
    if (flags & 5) {
        if (x1 == x2) return 0;
        bound = (flags & 4) ? b[0] : b[2];
        *y = (bound - x1) * (y2 - y1) / (x2 - x1) + y1;
        *x = bound;
    }
    flags = ((unsigned)(*y > b[3]) << 1) | ((unsigned)(*y < b[1]) << 3);
    if (flags & 10) {
        if (y1 == y2) return 0;
        bound = (flags & 8) ? b[1] : b[3];
        *x = (bound - y1) * (x2 - x1) / (y2 - y1) + x1;
        *y = bound;
    }
    return 1;
}
static unsigned lb_clip_segment(double* x1, double* y1, double* x2, double* y2, const double* b) {
    unsigned f1 = lb_flags(*x1, *y1, b), f2 = lb_flags(*x2, *y2, b), ret = 0;
    if ((f2 | f1) == 0) return 0;
    if ((f1 & 5) != 0 && (f1 & 5) == (f2 & 5)) return 4;
    if ((f1 & 10) != 0 && (f1 & 10) == (f2 & 10)) return 4;
    double tx1 = *x1, ty1 = *y1, tx2 = *x2, ty2 = *y2;
    if (f1) {
        if (!lb_move_point(tx1, ty1, tx2, ty2, b, x1, y1, f1)) return 4;
        if (*x1 == *x2 && *y1 == *y2) return 4;
        ret |= 1;
    }
    if (f2) {
        if (!lb_move_point(tx1, ty1, tx2, ty2, b, x2, y2, f2)) return 4;
        if (*x1 == *x2 && *y1 == *y2) return 4;
        ret |= 2;
    }
    return ret;
}

/* agg_math_stroke.h calc_cap, round cap, positive width.  Appends the cap's vertices around v0 (the segment goes to v1). */
static int calc_cap(double* vx, double* vy, int n0, double x0, double y0, double x1, double y1, double len, double w) {
    const double pi = 3.14159265358979323846;
    double dx1 = (y1 - y0) / len, dy1 = (x1 - x0) / len;
    dx1 *= w; dy1 *= w;
    double da = acos(w / (w + 0.125 / 1.0)) * 2;
    int n = (int)(pi / da);
    da = pi / (n + 1);
    int k = n0;
    vx[k] = x0 - dx1; vy[k] = y0 + dy1; ++k;
    double a1 = atan2(dy1, -dx1);
    a1 += da;
    for (int i = 0; i < n; i++) {
        vx[k] = x0 + cos(a1) * w; vy[k] = y0 + sin(a1) * w; ++k;
        a1 += da;
    }
    vx[k] = x0 + dx1; vy[k] = y0 - dy1; ++k;
    return k;
}

/* number of vertices one cap can have: n <= pi / da + 2 */
#define MAX_CAP 520

/* fixed_blender_rgba_plain::blend_pix on an opaque gray pixel v with white of alpha a (a in 1..254) */
static uint8_t blend_white(uint8_t v, unsigned a) {
    const unsigned A = 255;
    unsigned r = (unsigned)v * A;
    unsigned na = ((a + A) << 8) - a * A;
    return (uint8_t)(((((255u << 8) - r) * a) + (r << 8)) / na);
}

/* seg: n x 4 DATA-space segments (x0, y0, x1, y1): x runs along the image width, y along the rows (y inverted by
 * ax.invert_yaxis() and flipped back by the renderer), both scaled by the canvas size by transData; lw_points: n line widths in
 * points (LineCollection linewidths, tree2img.py:84-86,103).  out: H x W gray.  variant bits (exploration of details this
 * restatement could not pin a priori; 0 = what reproduces the shipped labels): 1 = no centre-line PathClipper, 2 = no snapping,
 * 8 = coverage rounds down instead of up (opposite polygon orientation), 16 = snap parity from ceil(width) instead of round. */
long agg_rasterize_segments(const double* seg, const double* lw_points, long n, int H, int W, int variant, uint8_t* out) {
    Ras r;
    memset(&r, 0, sizeof r);
    r.W = W; r.H = H;
    r.clip_x1 = 0; r.clip_y1 = 0; r.clip_x2 = W; r.clip_y2 = H;
    memset(out, 0, (size_t)H * W);
    const double path_clip[4] = {-1.0, -1.0, W + 1.0, H + 1.0};
    static double vx[2 * MAX_CAP + 8], vy[2 * MAX_CAP + 8];
    int* cov = NULL; int* are = NULL; size_t cap_cells = 0;
    long drawn = 0;
    for (long e = 0; e < n; ++e) {
        const double* q = seg + 4 * e;
        const double width_px = lw_points[e] * 100.0 / 72.0;     /* points_to_pixels, dpi = 100 (tree2img.py:51) */
        double x0 = q[0] * W, y0 = q[1] * H, x1 = q[2] * W, y1 = q[3] * H;
        if (!(variant & 1)) {
            if (lb_clip_segment(&x0, &y0, &x1, &y1, path_clip) >= 4) continue;
        }
        if (!(variant & 2)) {
            if (fabs(x0 - x1) < 1e-4 || fabs(y0 - y1) < 1e-4) {      /* PathSnapper, SNAP_AUTO */
                const double sv = (((variant & 16) ? (int)ceil(width_px) : (int)floor(width_px + 0.5)) % 2) ? 0.5 : 0.0;   /* mpl_round_to_int(stroke_width) % 2 */
                x0 = floor(x0 + 0.5) + sv; y0 = floor(y0 + 0.5) + sv; x1 = floor(x1 + 0.5) + sv; y1 = floor(y1 + 0.5) + sv;
            }
        }
        const double len = sqrt((x1 - x0) * (x1 - x0) + (y1 - y0) * (y1 - y0));
        if (!(len > 1e-14)) continue;                              /* vertex_dist: coincident vertices collapse, nothing is stroked */
        const double w = width_px * 0.5;
        if (!(w > 0)) continue;
        int nv = calc_cap(vx, vy, 0, x0, y0, x1, y1, len, w);
        nv = calc_cap(vx, vy, nv, x1, y1, x0, y0, len, w);
        /* cell window */
        double bx0 = vx[0], bx1 = vx[0], by0 = vy[0], by1 = vy[0];
        for (int k = 1; k < nv; ++k) { if (vx[k] < bx0) bx0 = vx[k]; if (vx[k] > bx1) bx1 = vx[k]; if (vy[k] < by0) by0 = vy[k]; if (vy[k] > by1) by1 = vy[k]; }
        int ix0 = (int)floor(bx0) - 1, ix1 = (int)floor(bx1) + 2, iy0 = (int)floor(by0) - 1, iy1 = (int)floor(by1) + 2;
        if (ix0 < 0) ix0 = 0;
        if (iy0 < 0) iy0 = 0;
        if (ix1 > W) ix1 = W;      /* column W / row H hold clipped edges */
        if (iy1 > H) iy1 = H;
        if (ix1 < ix0 || iy1 < iy0) continue;
        r.bx0 = ix0; r.by0 = iy0; r.bw = ix1 - ix0 + 1; r.bh = iy1 - iy0 + 1;
        const size_t cells = (size_t)r.bw * r.bh;
        if (cells > cap_cells) { free(cov); free(are); cov = (int*)malloc(cells * sizeof(int)); are = (int*)malloc(cells * sizeof(int)); cap_cells = cells; }
        memset(cov, 0, cells * sizeof(int)); memset(are, 0, cells * sizeof(int));
        r.cover = cov; r.area = are;
        r.minx = r.miny = 0x7fffffff; r.maxx = r.maxy = -0x7fffffff;
        /* rasterizer_scanline_aa::add_path: move_to, line_to ..., close_polygon */
        clip_move_to(&r, vx[0], vy[0]);
        for (int k = 1; k < nv; ++k) clip_line_to(&r, vx[k], vy[k]);
        clip_line_to(&r, vx[0], vy[0]);
        ++drawn;
        if (r.maxy < r.miny) continue;
        /* sweep_scanline + renderer_scanline_aa_solid over the canvas */
        for (int y = r.miny; y <= r.maxy && y < H; ++y) {
            if (y < 0) continue;
            const int* cr = cov + (size_t)(y - r.by0) * r.bw;
            const int* ar = are + (size_t)(y - r.by0) * r.bw;
            int cover = 0;
            uint8_t* row = out + (size_t)y * W;
            for (int x = r.bx0; x < r.bx0 + r.bw; ++x) {
                cover += cr[x - r.bx0];
                if (x >= W) break;
                int a = ((cover << (SUBPIXEL_SHIFT + 1)) - ar[x - r.bx0]);
                if (variant & 8) a = -a;
                int c = a >> (SUBPIXEL_SHIFT * 2 + 1 - 8);
                if (c < 0) c = -c;
                if (c > 255) c = 255;
                if (!c) continue;
                const unsigned alpha = (255u * ((unsigned)c + 1)) >> 8;
                if (alpha == 255) row[x] = 255;
                else if (alpha) row[x] = blend_white(row[x], alpha);
            }
        }
    }
    free(cov); free(are);
    return drawn;
}

/* edges7: E x 7 (node1 xyz, node2 xyz, radius): tree2img.py:66-68,82-86 for every edge, then the collection is drawn */
long agg_rasterize(const double* edges7, long E, int H, int W, int ax0, int ax1, double min_radius, double max_radius,
                   int variant, uint8_t* out) {
    double* seg = (double*)malloc(sizeof(double) * 4 * (size_t)(E > 0 ? E : 1));
    double* lw = (double*)malloc(sizeof(double) * (size_t)(E > 0 ? E : 1));
    const double scale = (double)(W > H ? W : H);
    long n = 0;
    for (long e = 0; e < E; ++e) {
        const double* q = edges7 + 7 * e;
        double radius = q[6];
        if (radius < min_radius || radius > max_radius) continue;
        radius *= 1.3;
        lw[n] = radius * scale;
        seg[4 * n] = q[ax1]; seg[4 * n + 1] = q[ax0]; seg[4 * n + 2] = q[3 + ax1]; seg[4 * n + 3] = q[3 + ax0];
        ++n;
    }
    long d = agg_rasterize_segments(seg, lw, n, H, W, variant, out);
    free(seg); free(lw);
    return d;
}
