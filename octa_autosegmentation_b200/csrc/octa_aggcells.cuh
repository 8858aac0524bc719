// Scanline anti-aliasing arithmetic of the 2-D label path, host/device shared.
//
// rasterize_forest (vessel_graph_generation/tree2img.py:12-114) hands one LineCollection of round-capped, anti-aliased white
// strokes to matplotlib's Agg backend; every number of the result comes out of that third-party code (matplotlib
// src/_backend_agg.h, src/path_converters.h, src/agg_workaround.h and the Anti-Grain Geometry 2.4 headers it vendors:
// agg_math_stroke.h, agg_rasterizer_sl_clip.h, agg_rasterizer_cells_aa.h, agg_rasterizer_scanline_aa.h).  This header restates
// the stages a 2-vertex path goes through, in a form that can be evaluated ROW BY ROW in parallel:
//   prepare_stroke     tree2img.py:66-86 + PathClipper (centre line vs [-1, W+1] x [-1, H+1]) + PathSnapper (auto)
//   stroke_vertex      math_stroke::calc_cap, round caps: the outline is a polygon inscribed in the capsule (1/8 px tolerance)
//   clip_edge          rasterizer_sl_clip<ras_conv_dbl>::line_to for one outline edge (clip box = canvas) -> 24.8 integer lines
//   line_row           rasterizer_cells_aa::line restricted to ONE scanline: Agg walks a line scanline by scanline with an
//                      integer DDA whose state after k steps is x1 + floor(k-dependent product / dy); the closed form gives the
//                      piece of the line inside row `ey` directly, so rows can be processed independently
//   render_hline       rasterizer_cells_aa::render_hline, verbatim arithmetic: (cover, area) contributions of that piece to
//                      the cells of the row
//   calc_alpha, blend_white   rasterizer_scanline_aa::calculate_alpha (non-zero winding, 8-bit) and
//                      fixed_blender_rgba_plain::blend_pix for white over an opaque gray pixel
// The test tree holds a sequential CPU restatement of the same pipeline (agg_oracle.c; it reproduces all 500 label PNGs the reference
// ships bit for bit); tests compare this code with it cell by cell and image by image.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define OCTA_AGG_HD __host__ __device__ __forceinline__
#else
#define OCTA_AGG_HD inline
#endif

namespace octa {
namespace agg {

constexpr int SUB_SHIFT = 8, SUB_SCALE = 256, SUB_MASK = 255;
constexpr int MAX_CAP_SEG = 62;                 // arc steps of one round cap (half width up to ~400 px)
constexpr int MAX_VERT = 2 * (MAX_CAP_SEG + 2); // outline vertices of one stroke
constexpr int MAX_LINES = MAX_VERT + 16;        // integer lines of one stroke after clipping

struct Stroke {     // clipped / snapped centre line in device pixels (y grows with the row index) and half width
    double x0, y0, x1, y1, w;
};
struct Line { int x1, y1, x2, y2; };

OCTA_AGG_HD int iround(double v) { return (int)((v < 0.0) ? v - 0.5 : v + 0.5); }
OCTA_AGG_HD long long floordiv(long long a, long long b) {      // b > 0; agg: delta = p / dy; if (p % dy < 0) delta--
    if (a > -2147483647LL && a < 2147483647LL && b < 2147483647LL) {       // the usual case: 32-bit division
        const int ai = (int)a, bi = (int)b;
        int q = ai / bi;
        if (ai % bi < 0) --q;
        return q;
    }
    long long q = a / b;
    if (a % b < 0) --q;
    return q;
}

// ---- centre line: agg_clip_liang_barsky.h clip_line_segment, as PathClipper::draw_clipped_line calls it
OCTA_AGG_HD unsigned lb_flags(double x, double y, const double* b) {
    return (unsigned)(x > b[2]) | ((unsigned)(y > b[3]) << 1) | ((unsigned)(x < b[0]) << 2) | ((unsigned)(y < b[1]) << 3);
}
OCTA_AGG_HD bool lb_move_point(double x1, double y1, double x2, double y2, const double* b, double* x, double* y, unsigned flags) {
    double bound;
    if (flags & 5) {
        if (x1 == x2) return false;
        bound = (flags & 4) ? b[0] : b[2];
        *y = (bound - x1) * (y2 - y1) / (x2 - x1) + y1;
        *x = bound;
    }
    flags = ((unsigned)(*y > b[3]) << 1) | ((unsigned)(*y < b[1]) << 3);
    if (flags & 10) {
        if (y1 == y2) return false;
        bound = (flags & 8) ? b[1] : b[3];
        *x = (bound - y1) * (x2 - x1) / (y2 - y1) + x1;
        *y = bound;
    }
    return true;
}
OCTA_AGG_HD unsigned lb_clip_segment(double* x1, double* y1, double* x2, double* y2, const double* b) {
    const unsigned f1 = lb_flags(*x1, *y1, b), f2 = lb_flags(*x2, *y2, b);
    unsigned ret = 0;
    if ((f2 | f1) == 0) return 0;
    if ((f1 & 5) != 0 && (f1 & 5) == (f2 & 5)) return 4;
    if ((f1 & 10) != 0 && (f1 & 10) == (f2 & 10)) return 4;
    const double tx1 = *x1, ty1 = *y1, tx2 = *x2, ty2 = *y2;
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

// tree2img.py:66-86 for one kept edge + PathClipper + PathSnapper.  Returns false when nothing is drawn.
OCTA_AGG_HD bool prepare_stroke(const double* e7, int ax0, int ax1, int H, int W, double scale, double min_radius,
                                double max_radius, Stroke* s) {
    double radius = e7[6];
    if (radius < min_radius || radius > max_radius) return false;            // :67
    radius *= 1.3;                                                            // :82
    const double thickness = radius * scale;                                  // :84 (points)
    const double width_px = thickness * 100.0 / 72.0;                         // points_to_pixels, dpi = 100 (:51)
    double x0 = e7[ax1] * W, y0 = e7[ax0] * H, x1 = e7[3 + ax1] * W, y1 = e7[3 + ax0] * H;   // :85, transData, y flipped twice
    const double path_clip[4] = {-1.0, -1.0, W + 1.0, H + 1.0};
    if (lb_clip_segment(&x0, &y0, &x1, &y1, path_clip) >= 4) return false;
    if (fabs(x0 - x1) < 1e-4 || fabs(y0 - y1) < 1e-4) {                       // PathSnapper, SNAP_AUTO: rectilinear path
        const double sv = ((int)floor(width_px + 0.5) % 2) ? 0.5 : 0.0;       // mpl_round_to_int(stroke_width) % 2
        x0 = floor(x0 + 0.5) + sv; y0 = floor(y0 + 0.5) + sv; x1 = floor(x1 + 0.5) + sv; y1 = floor(y1 + 0.5) + sv;
    }
    const double len = sqrt((x1 - x0) * (x1 - x0) + (y1 - y0) * (y1 - y0));
    if (!(len > 1e-14)) return false;                                         // vertex_dist: coincident vertices collapse
    const double w = width_px * 0.5;
    if (!(w > 0)) return false;
    s->x0 = x0; s->y0 = y0; s->x1 = x1; s->y1 = y1; s->w = w;
    return true;
}

// math_stroke::calc_cap (round): number of arc steps of one cap
OCTA_AGG_HD int cap_steps(double w) {
    const double pi = 3.14159265358979323846;
    const double da = acos(w / (w + 0.125 / 1.0)) * 2;
    return (int)(pi / da);
}

// vertex k (0 .. 2(n+2)-1) of the stroke outline: cap around (x0,y0) first, then the cap around (x1,y1)
OCTA_AGG_HD void stroke_vertex(const Stroke& s, int n, int k, double* vx, double* vy) {
    const double pi = 3.14159265358979323846;
    double cx = s.x0, cy = s.y0, ox = s.x1, oy = s.y1;
    if (k >= n + 2) { k -= n + 2; cx = s.x1; cy = s.y1; ox = s.x0; oy = s.y0; }
    const double len = sqrt((s.x1 - s.x0) * (s.x1 - s.x0) + (s.y1 - s.y0) * (s.y1 - s.y0));   // vertex_dist::dist (same for both ends)
    double dx1 = (oy - cy) / len, dy1 = (ox - cx) / len;
    dx1 *= s.w; dy1 *= s.w;
    if (k == 0) { *vx = cx - dx1; *vy = cy + dy1; return; }
    if (k == n + 1) { *vx = cx + dx1; *vy = cy - dy1; return; }
    const double da = pi / (n + 1);
    double a1 = atan2(dy1, -dx1);
    for (int i = 0; i < k; ++i) a1 += da;
    *vx = cx + cos(a1) * s.w;
    *vy = cy + sin(a1) * s.w;
}

// ---- rasterizer_sl_clip<ras_conv_dbl>::line_to for one outline edge (xa,ya) -> (xb,yb); clip box [0,W] x [0,H]
struct Clip { double x1, y1, x2, y2; };
OCTA_AGG_HD unsigned clip_flags(const Clip& c, double x, double y) {
    return (unsigned)(x > c.x2) | ((unsigned)(y > c.y2) << 1) | ((unsigned)(x < c.x1) << 2) | ((unsigned)(y < c.y1) << 3);
}
OCTA_AGG_HD unsigned clip_flags_y(const Clip& c, double y) { return ((unsigned)(y > c.y2) << 1) | ((unsigned)(y < c.y1) << 3); }
OCTA_AGG_HD int xi(double v) { return iround(v * SUB_SCALE); }

OCTA_AGG_HD int line_clip_y(const Clip& c, double x1, double y1, double x2, double y2, unsigned f1, unsigned f2, Line* out) {
    f1 &= 10; f2 &= 10;
    if ((f1 | f2) == 0) { out->x1 = xi(x1); out->y1 = xi(y1); out->x2 = xi(x2); out->y2 = xi(y2); return 1; }
    if (f1 == f2) return 0;
    double tx1 = x1, ty1 = y1, tx2 = x2, ty2 = y2;
    if (f1 & 8) { tx1 = x1 + (c.y1 - y1) * (x2 - x1) / (y2 - y1); ty1 = c.y1; }
    if (f1 & 2) { tx1 = x1 + (c.y2 - y1) * (x2 - x1) / (y2 - y1); ty1 = c.y2; }
    if (f2 & 8) { tx2 = x1 + (c.y1 - y1) * (x2 - x1) / (y2 - y1); ty2 = c.y1; }
    if (f2 & 2) { tx2 = x1 + (c.y2 - y1) * (x2 - x1) / (y2 - y1); ty2 = c.y2; }
    out->x1 = xi(tx1); out->y1 = xi(ty1); out->x2 = xi(tx2); out->y2 = xi(ty2);
    return 1;
}

// returns the number of integer lines written to out[0..3)
OCTA_AGG_HD int clip_edge(const Clip& c, double x1, double y1, double x2, double y2, Line* out) {
    const unsigned f1 = clip_flags(c, x1, y1), f2 = clip_flags(c, x2, y2);
    if ((f1 & 10) == (f2 & 10) && (f1 & 10) != 0) return 0;       // invisible by y
    double y3, y4;
    unsigned f3, f4;
    int n = 0;
    switch (((f1 & 5) << 1) | (f2 & 5)) {
    case 0: n += line_clip_y(c, x1, y1, x2, y2, f1, f2, out + n); break;
    case 1:
        y3 = y1 + (c.x2 - x1) * (y2 - y1) / (x2 - x1); f3 = clip_flags_y(c, y3);
        n += line_clip_y(c, x1, y1, c.x2, y3, f1, f3, out + n); n += line_clip_y(c, c.x2, y3, c.x2, y2, f3, f2, out + n); break;
    case 2:
        y3 = y1 + (c.x2 - x1) * (y2 - y1) / (x2 - x1); f3 = clip_flags_y(c, y3);
        n += line_clip_y(c, c.x2, y1, c.x2, y3, f1, f3, out + n); n += line_clip_y(c, c.x2, y3, x2, y2, f3, f2, out + n); break;
    case 3: n += line_clip_y(c, c.x2, y1, c.x2, y2, f1, f2, out + n); break;
    case 4:
        y3 = y1 + (c.x1 - x1) * (y2 - y1) / (x2 - x1); f3 = clip_flags_y(c, y3);
        n += line_clip_y(c, x1, y1, c.x1, y3, f1, f3, out + n); n += line_clip_y(c, c.x1, y3, c.x1, y2, f3, f2, out + n); break;
    case 6:
        y3 = y1 + (c.x2 - x1) * (y2 - y1) / (x2 - x1); y4 = y1 + (c.x1 - x1) * (y2 - y1) / (x2 - x1);
        f3 = clip_flags_y(c, y3); f4 = clip_flags_y(c, y4);
        n += line_clip_y(c, c.x2, y1, c.x2, y3, f1, f3, out + n); n += line_clip_y(c, c.x2, y3, c.x1, y4, f3, f4, out + n);
        n += line_clip_y(c, c.x1, y4, c.x1, y2, f4, f2, out + n); break;
    case 8:
        y3 = y1 + (c.x1 - x1) * (y2 - y1) / (x2 - x1); f3 = clip_flags_y(c, y3);
        n += line_clip_y(c, c.x1, y1, c.x1, y3, f1, f3, out + n); n += line_clip_y(c, c.x1, y3, x2, y2, f3, f2, out + n); break;
    case 9:
        y3 = y1 + (c.x1 - x1) * (y2 - y1) / (x2 - x1); y4 = y1 + (c.x2 - x1) * (y2 - y1) / (x2 - x1);
        f3 = clip_flags_y(c, y3); f4 = clip_flags_y(c, y4);
        n += line_clip_y(c, c.x1, y1, c.x1, y3, f1, f3, out + n); n += line_clip_y(c, c.x1, y3, c.x2, y4, f3, f4, out + n);
        n += line_clip_y(c, c.x2, y4, c.x2, y2, f4, f2, out + n); break;
    case 12: n += line_clip_y(c, c.x1, y1, c.x1, y2, f1, f2, out + n); break;
    }
    return n;
}

// ---- rasterizer_cells_aa::render_hline: the piece (x1,y1) -> (x2,y2) of a line inside one scanline (y1, y2 = 0..256 fractions);
// add(ex, cover, area) receives the contribution to cell ex
template <class Add>
OCTA_AGG_HD void render_hline(int x1, int y1, int x2, int y2, Add& add) {
    int ex1 = x1 >> SUB_SHIFT;
    const int ex2 = x2 >> SUB_SHIFT;
    const int fx1 = x1 & SUB_MASK, fx2 = x2 & SUB_MASK;
    int delta, p, first, dx, incr, lift, mod, rem;
    if (y1 == y2) return;
    if (ex1 == ex2) {
        delta = y2 - y1;
        add(ex1, delta, (fx1 + fx2) * delta);
        return;
    }
    p = (SUB_SCALE - fx1) * (y2 - y1);
    first = SUB_SCALE;
    incr = 1;
    dx = x2 - x1;
    if (dx < 0) { p = fx1 * (y2 - y1); first = 0; incr = -1; dx = -dx; }
    delta = p / dx;
    mod = p % dx;
    if (mod < 0) { delta--; mod += dx; }
    add(ex1, delta, (fx1 + first) * delta);
    ex1 += incr;
    y1 += delta;
    if (ex1 != ex2) {
        p = SUB_SCALE * (y2 - y1 + delta);
        lift = p / dx;
        rem = p % dx;
        if (rem < 0) { lift--; rem += dx; }
        mod -= dx;
        while (ex1 != ex2) {
            delta = lift;
            mod += rem;
            if (mod >= 0) { mod -= dx; delta++; }
            add(ex1, delta, SUB_SCALE * delta);
            y1 += delta;
            ex1 += incr;
        }
    }
    delta = y2 - y1;
    add(ex1, delta, (fx2 + SUB_SCALE - first) * delta);
}

// ---- rasterizer_cells_aa::line restricted to scanline `ey`.  Agg's loop reaches the scanline boundary that lies `dist`
// sub-pixels of y away from y1 at x1 + floor(dist * dx / |dy|) (its delta / mod / lift / rem recurrence is exactly this floor
// division), entering a scanline at fraction 0 (going down) or 256 (going up) and leaving at the opposite one.  One code path
// for both directions (the lanes of a warp hold lines of either direction).
template <class Add>
OCTA_AGG_HD void line_row(const Line& L, int ey, Add& add) {
    const int x1 = L.x1, y1 = L.y1, x2 = L.x2, y2 = L.y2;
    const int ey1 = y1 >> SUB_SHIFT, ey2 = y2 >> SUB_SHIFT;
    if (ey1 <= ey2 ? (ey < ey1 || ey > ey2) : (ey < ey2 || ey > ey1)) return;
    const int fy1 = y1 & SUB_MASK, fy2 = y2 & SUB_MASK;
    if (ey1 == ey2) { render_hline(x1, fy1, x2, fy2, add); return; }
    const int dx = x2 - x1;
    const int dy = y2 - y1;
    const bool up = dy < 0;
    const int first = up ? 0 : SUB_SCALE;              // fraction at which the line leaves a scanline
    if (dx == 0) {
        const int ex = x1 >> SUB_SHIFT;
        const int two_fx = (x1 - (ex << SUB_SHIFT)) << 1;
        int delta;
        if (ey == ey1) delta = first - fy1;
        else if (ey == ey2) delta = fy2 - SUB_SCALE + first;
        else delta = first + first - SUB_SCALE;
        add(ex, delta, two_fx * delta);
        return;
    }
    const int ady = up ? -dy : dy;
    const int j = up ? ey1 - ey : ey - ey1;            // scanlines since the start
    // distance (sub-pixels of y) from y1 to the boundary through which the line ENTERS scanline ey (j >= 1) / LEAVES it
    const long long d_in = up ? (long long)fy1 + (long long)(j - 1) * SUB_SCALE : (long long)j * SUB_SCALE - fy1;
    const long long d_out = d_in + SUB_SCALE;
    const int xa = j == 0 ? x1 : x1 + (int)floordiv(d_in * dx, ady);
    const int ya = j == 0 ? fy1 : SUB_SCALE - first;
    const bool last = ey == ey2;
    const int xb = last ? x2 : x1 + (int)floordiv((j == 0 ? (up ? (long long)fy1 : (long long)SUB_SCALE - fy1) : d_out) * dx, ady);
    const int yb = last ? fy2 : first;
    render_hline(xa, ya, xb, yb, add);
}

// rasterizer_scanline_aa::calculate_alpha (fill_non_zero, gamma = identity); v = (running cover << 9) - area of the cell
OCTA_AGG_HD int calc_alpha(int v) {
    int c = v >> (SUB_SHIFT * 2 + 1 - 8);
    if (c < 0) c = -c;
    return c > 255 ? 255 : c;
}

// renderer_scanline_aa_solid + pixfmt_rgba (fixed_blender_rgba_plain) for opaque white on an opaque gray pixel v
OCTA_AGG_HD unsigned char blend_cover(unsigned char v, int cover) {
    if (!cover) return v;
    const unsigned alpha = (255u * ((unsigned)cover + 1)) >> 8;
    if (alpha == 255) return 255;
    if (!alpha) return v;
    const unsigned r = (unsigned)v * 255u;
    const unsigned na = ((alpha + 255u) << 8) - alpha * 255u;
    const unsigned num = (((255u << 8) - r) * alpha) + (r << 8);
    // num < 2^24 and na < 2^16 are exact in float; the rounded quotient is off by at most one, which the check repairs
    unsigned q = (unsigned)((float)num / (float)na);
    if (q * na > num) --q;
    else if ((q + 1) * na <= num) ++q;
    return (unsigned char)q;
}

}  // namespace agg
}  // namespace octa
