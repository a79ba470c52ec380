/*
 * oracle/splat_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Scalar CPU restatement (plain C, fp32 arithmetic, no FMA contraction) of the
 * orthographic TSW Gaussian splatting path that GSVC calls through
 *   /root/reference/ortho_gaussian_renderer/renderer.py:63-98   (rasterizer forward)
 *   /root/reference/ortho_gaussian_renderer/preprocess.py:58-104 (visible_filter)
 *
 * PARITY UNPINNED: the arithmetic of this path lives in the un-vendored,
 * un-pinned dependency github.com/actcwlf/ortho_diff_gaussian_rasterization
 * (/root/reference/README.md:52), which is absent from /root/reference, not
 * installed, and cannot be fetched (no network).  The reference has no tests or
 * golden vectors for the path.  This file therefore
 * restates the published 3DGS tile-rasterizer algorithm specialised to GSVC's
 * orthographic camera as frozen in SURVEY.md Appendix A / DESIGN.md §2
 * (decisions U1..U8), anchored on the reference's call sites and on the
 * in-tree conventions it cites:
 *   - slab cull rule            preprocess.py:109-116
 *   - pixel mapping             utils/loss_utils.py:122-127
 *   - quaternion (r,x,y,z)→R    utils/general_utils.py:98-119
 *   - SH constants / basis      utils/sh_utils.py:26-110
 *   - view matrices             frame_cube/frame.py:18-43
 *   - output layout [3,H,W]     pipeline/train.py:407
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product
 * (gsvc_b200/, libgsvc_rast.so) never does.
 *
 * Floating point: every expression is written in the same left-to-right order
 * as the CUDA preprocess kernel (which is compiled with -fmad=false), so the
 * per-Gaussian stage (radii, pixel centres, conics, tile rectangles, depth
 * keys) is expected to agree BIT-EXACTLY with the GPU; the blend stages use
 * glibc expf and therefore agree to tolerance only.
 * Build with:  gcc -O2 -ffp-contract=off -fno-fast-math -fopenmp -shared -fPIC
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define TILE 16
#define ALPHA_MIN (1.0f / 255.0f)
#define ALPHA_MAX 0.99f
#define T_STOP 0.0001f

static int g_power_f64;
static inline double exponent_f64(const float *co, float dx, float dy);
static inline double well_conditioned_err(double pw, double S);
static double g_err_scale = 0.25;  /* well_conditioned_err = this x the worst-case bound (calibration below) */
#define LOWPASS 0.3f

typedef struct {
    int32_t W, H;
    float x_min, y_min, scale, threshold, scale_modifier;
    float bg[3];
    float V[16];        /* logical row-major 4x4: p_v = V[:3,:3] p + V[:3,3] (SURVEY §8c ix) */
    int32_t sh_degree;  /* active SH degree (0..3) */
    int32_t sh_M;       /* coefficients per Gaussian stored in shs ((max_deg+1)^2) */
    float campos[3];
} orc_settings;

/* SH constants: /root/reference/utils/sh_utils.py:26-43 */
static const float SH_C0 = 0.28209479177387814f;
static const float SH_C1 = 0.4886025119029199f;
static const float SH_C2[5] = {1.0925484305920792f, -1.0925484305920792f, 0.31539156525252005f,
                               -1.0925484305920792f, 0.5462742152960396f};
static const float SH_C3[7] = {-0.5900435899266435f, 2.890611442640554f, -0.4570457994644658f,
                               0.3731763325901154f, -0.4570457994644658f, 1.445305721320277f,
                               -0.5900435899266435f};

/* U2: order-preserving float→uint32 map (ascending key == ascending view depth). */
static uint32_t ordered_u32(float z)
{
    uint32_t u;
    memcpy(&u, &z, 4);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

/* Rotation from quaternion (r,x,y,z), used as given (no re-normalisation).
 * Convention: /root/reference/utils/general_utils.py:98-119. */
static void quat_to_rot(const float *q, float R[9])
{
    float r = q[0], x = q[1], y = q[2], z = q[3];
    R[0] = 1.f - 2.f * (y * y + z * z);
    R[1] = 2.f * (x * y - r * z);
    R[2] = 2.f * (x * z + r * y);
    R[3] = 2.f * (x * y + r * z);
    R[4] = 1.f - 2.f * (x * x + z * z);
    R[5] = 2.f * (y * z - r * x);
    R[6] = 2.f * (x * z - r * y);
    R[7] = 2.f * (y * z + r * x);
    R[8] = 1.f - 2.f * (x * x + y * y);
}

/* Σ = (R S)(R S)^T, packed (xx,xy,xz,yy,yz,zz).  Appendix A.1. */
static void cov3d_from_scale_rot(const float *s, float mod, const float *q, float cov[6])
{
    float R[9], M[9];
    quat_to_rot(q, R);
    float sx = mod * s[0], sy = mod * s[1], sz = mod * s[2];
    for (int i = 0; i < 3; i++) {
        M[3 * i + 0] = R[3 * i + 0] * sx;
        M[3 * i + 1] = R[3 * i + 1] * sy;
        M[3 * i + 2] = R[3 * i + 2] * sz;
    }
    cov[0] = M[0] * M[0] + M[1] * M[1] + M[2] * M[2];
    cov[1] = M[0] * M[3] + M[1] * M[4] + M[2] * M[5];
    cov[2] = M[0] * M[6] + M[1] * M[7] + M[2] * M[8];
    cov[3] = M[3] * M[3] + M[4] * M[4] + M[5] * M[5];
    cov[4] = M[3] * M[6] + M[4] * M[7] + M[5] * M[8];
    cov[5] = M[6] * M[6] + M[7] * M[7] + M[8] * M[8];
}

/* Orthographic cov2D: J = scale*[I2|0] so cov2D = scale^2 (W Σ W^T)[0:2,0:2] + 0.3 I (U5). */
static void cov2d_ortho(const float cov[6], const float *V, float scale, float *a, float *b, float *c)
{
    float w00 = V[0], w01 = V[1], w02 = V[2];
    float w10 = V[4], w11 = V[5], w12 = V[6];
    /* u0 = Σ w0, u1 = Σ w1 */
    float u00 = cov[0] * w00 + cov[1] * w01 + cov[2] * w02;
    float u01 = cov[1] * w00 + cov[3] * w01 + cov[4] * w02;
    float u02 = cov[2] * w00 + cov[4] * w01 + cov[5] * w02;
    float u10 = cov[0] * w10 + cov[1] * w11 + cov[2] * w12;
    float u11 = cov[1] * w10 + cov[3] * w11 + cov[4] * w12;
    float u12 = cov[2] * w10 + cov[4] * w11 + cov[5] * w12;
    float s2 = scale * scale;
    *a = s2 * (w00 * u00 + w01 * u01 + w02 * u02) + LOWPASS;
    *b = s2 * (w00 * u10 + w01 * u11 + w02 * u12);
    *c = s2 * (w10 * u10 + w11 * u11 + w12 * u12) + LOWPASS;
}

static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

/* SH → RGB (+0.5, clamp ≥ 0), basis of /root/reference/utils/sh_utils.py:57-110. */
static void sh_to_rgb(int deg, int M, const float *sh /* [M][3] */, const float *p, const float *campos,
                      float rgb[3], uint8_t clamped[3])
{
    float dx = p[0] - campos[0], dy = p[1] - campos[1], dz = p[2] - campos[2];
    float inv = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
    float x = dx * inv, y = dy * inv, z = dz * inv;
    (void)M;
    for (int c = 0; c < 3; c++) {
        float r = SH_C0 * sh[0 * 3 + c];
        if (deg > 0) {
            r = r - SH_C1 * y * sh[1 * 3 + c] + SH_C1 * z * sh[2 * 3 + c] - SH_C1 * x * sh[3 * 3 + c];
            if (deg > 1) {
                float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                r = r + SH_C2[0] * xy * sh[4 * 3 + c] + SH_C2[1] * yz * sh[5 * 3 + c] +
                    SH_C2[2] * (2.0f * zz - xx - yy) * sh[6 * 3 + c] + SH_C2[3] * xz * sh[7 * 3 + c] +
                    SH_C2[4] * (xx - yy) * sh[8 * 3 + c];
                if (deg > 2) {
                    r = r + SH_C3[0] * y * (3.0f * xx - yy) * sh[9 * 3 + c] + SH_C3[1] * xy * z * sh[10 * 3 + c] +
                        SH_C3[2] * y * (4.0f * zz - xx - yy) * sh[11 * 3 + c] +
                        SH_C3[3] * z * (2.0f * zz - 3.0f * xx - 3.0f * yy) * sh[12 * 3 + c] +
                        SH_C3[4] * x * (4.0f * zz - xx - yy) * sh[13 * 3 + c] +
                        SH_C3[5] * z * (xx - yy) * sh[14 * 3 + c] + SH_C3[6] * x * (xx - 3.0f * yy) * sh[15 * 3 + c];
                }
            }
        }
        r += 0.5f;
        clamped[c] = (r < 0.0f);
        rgb[c] = r < 0.0f ? 0.0f : r;
    }
}

/*
 * Per-Gaussian preprocess (Appendix A.1).  All outputs are [P]-sized and
 * zero-filled for culled Gaussians.  Either (scales,rotations) or cov3D_precomp
 * (packed 6) is given; either colors_precomp or shs (may both be NULL for
 * visible_filter, in which case only radii/rect outputs are meaningful).
 *
 * rect[g] = (min_x, min_y, max_x, max_y) in tiles, max exclusive.
 */
void orc_preprocess(const orc_settings *st, int P, const float *means3D, const float *scales,
                    const float *rotations, const float *cov3D_precomp, const float *opacities,
                    const float *colors_precomp, const float *shs, int32_t *radii, float *depth, float *xy,
                    float *conic_opacity, float *rgb, int32_t *rect, int32_t *tiles_touched, uint8_t *clamped,
                    float *cov3D_out)
{
    const float *V = st->V;
    int gx = (st->W + TILE - 1) / TILE, gy = (st->H + TILE - 1) / TILE;
    for (int g = 0; g < P; g++) {
        radii[g] = 0;
        if (tiles_touched) tiles_touched[g] = 0;
        if (rect) rect[4 * g] = rect[4 * g + 1] = rect[4 * g + 2] = rect[4 * g + 3] = 0;
        if (depth) depth[g] = 0.f;
        if (xy) xy[2 * g] = xy[2 * g + 1] = 0.f;
        if (conic_opacity) memset(conic_opacity + 4 * g, 0, 16);
        if (rgb) memset(rgb + 3 * g, 0, 12);
        if (clamped) memset(clamped + 3 * g, 0, 3);
        if (cov3D_out) memset(cov3D_out + 6 * g, 0, 24);

        const float *p = means3D + 3 * g;
        float vx = V[0] * p[0] + V[1] * p[1] + V[2] * p[2] + V[3];
        float vy = V[4] * p[0] + V[5] * p[1] + V[6] * p[2] + V[7];
        float vz = V[8] * p[0] + V[9] * p[1] + V[10] * p[2] + V[11];
        /* U6: TSW slab cull, preprocess.py:109-116 */
        if (fabsf(vz) > st->threshold) continue;

        float cov[6];
        if (cov3D_precomp)
            memcpy(cov, cov3D_precomp + 6 * g, 24);
        else
            cov3d_from_scale_rot(scales + 3 * g, st->scale_modifier, rotations + 4 * g, cov);
        float a, b, c;
        cov2d_ortho(cov, V, st->scale, &a, &b, &c);
        float det = a * c - b * b;
        if (det == 0.0f) continue;
        float det_inv = 1.f / det;
        float cA = c * det_inv, cB = -b * det_inv, cC = a * det_inv;
        float mid = 0.5f * (a + c);
        float lam = mid + sqrtf(fmaxf(0.1f, mid * mid - det));
        float rad_f = ceilf(3.f * sqrtf(lam));
        int radius = (int)rad_f;
        /* U1: pixel centre convention */
        float px = (vx - st->x_min) * st->scale - 0.5f;
        float py = (vy - st->y_min) * st->scale - 0.5f;
        int rminx = imin(gx, imax(0, (int)((px - rad_f) / (float)TILE)));
        int rminy = imin(gy, imax(0, (int)((py - rad_f) / (float)TILE)));
        int rmaxx = imin(gx, imax(0, (int)((px + rad_f + (float)(TILE - 1)) / (float)TILE)));
        int rmaxy = imin(gy, imax(0, (int)((py + rad_f + (float)(TILE - 1)) / (float)TILE)));
        int area = (rmaxx - rminx) * (rmaxy - rminy);
        if (area <= 0) continue;

        radii[g] = radius;
        if (tiles_touched) tiles_touched[g] = area;
        if (rect) {
            rect[4 * g] = rminx; rect[4 * g + 1] = rminy; rect[4 * g + 2] = rmaxx; rect[4 * g + 3] = rmaxy;
        }
        if (depth) depth[g] = vz;
        if (xy) { xy[2 * g] = px; xy[2 * g + 1] = py; }
        if (conic_opacity) {
            conic_opacity[4 * g] = cA; conic_opacity[4 * g + 1] = cB; conic_opacity[4 * g + 2] = cC;
            conic_opacity[4 * g + 3] = opacities ? opacities[g] : 0.f;
        }
        if (cov3D_out) memcpy(cov3D_out + 6 * g, cov, 24);
        if (rgb) {
            if (colors_precomp) {
                memcpy(rgb + 3 * g, colors_precomp + 3 * g, 12);
            } else if (shs) {
                uint8_t cl[3];
                sh_to_rgb(st->sh_degree, st->sh_M, shs + (size_t)g * st->sh_M * 3, p, st->campos, rgb + 3 * g, cl);
                if (clamped) memcpy(clamped + 3 * g, cl, 3);
            }
        }
    }
}

/* A.2: emit (key,val) instances in Gaussian-index order, tiles row-major (U8). Returns R. */
int64_t orc_count_instances(int P, const int32_t *tiles_touched)
{
    int64_t R = 0;
    for (int g = 0; g < P; g++) R += tiles_touched[g];
    return R;
}

void orc_duplicate_with_keys(const orc_settings *st, int P, const int32_t *radii, const int32_t *rect,
                             const float *depth, uint64_t *keys, uint32_t *vals)
{
    int gx = (st->W + TILE - 1) / TILE;
    int64_t off = 0;
    for (int g = 0; g < P; g++) {
        if (radii[g] <= 0) continue;
        uint32_t dk = ordered_u32(depth[g]);
        for (int ty = rect[4 * g + 1]; ty < rect[4 * g + 3]; ty++)
            for (int tx = rect[4 * g]; tx < rect[4 * g + 2]; tx++) {
                uint64_t tile = (uint64_t)(ty * gx + tx);
                keys[off] = (tile << 32) | dk;
                vals[off] = (uint32_t)g;
                off++;
            }
    }
}

/* Stable LSD radix sort (8-bit digits) over the low `nbits` key bits — the
 * reference semantics of cub::DeviceRadixSort::SortPairs(keys,vals,R,0,nbits). */
void orc_sort_pairs(int64_t R, int nbits, uint64_t *keys, uint32_t *vals)
{
    if (R <= 1) return;
    uint64_t *k2 = (uint64_t *)malloc(sizeof(uint64_t) * R);
    uint32_t *v2 = (uint32_t *)malloc(sizeof(uint32_t) * R);
    uint64_t *ka = keys, *kb = k2;
    uint32_t *va = vals, *vb = v2;
    for (int shift = 0; shift < nbits; shift += 8) {
        int64_t cnt[257];
        memset(cnt, 0, sizeof(cnt));
        for (int64_t i = 0; i < R; i++) cnt[((ka[i] >> shift) & 0xFF) + 1]++;
        for (int d = 0; d < 256; d++) cnt[d + 1] += cnt[d];
        for (int64_t i = 0; i < R; i++) {
            int64_t dst = cnt[(ka[i] >> shift) & 0xFF]++;
            kb[dst] = ka[i];
            vb[dst] = va[i];
        }
        uint64_t *tk = ka; ka = kb; kb = tk;
        uint32_t *tv = va; va = vb; vb = tv;
    }
    if (ka != keys) {
        memcpy(keys, ka, sizeof(uint64_t) * R);
        memcpy(vals, va, sizeof(uint32_t) * R);
    }
    free(k2);
    free(v2);
}

/* identifyTileRanges: ranges[t] = [start,end); untouched tiles stay (0,0). */
void orc_tile_ranges(int64_t R, const uint64_t *sorted_keys, int n_tiles, uint32_t *ranges /* [T][2] */)
{
    memset(ranges, 0, sizeof(uint32_t) * 2 * (size_t)n_tiles);
    for (int64_t i = 0; i < R; i++) {
        uint32_t t = (uint32_t)(sorted_keys[i] >> 32);
        if (i == 0 || t != (uint32_t)(sorted_keys[i - 1] >> 32)) ranges[2 * t] = (uint32_t)i;
        if (i == R - 1 || t != (uint32_t)(sorted_keys[i + 1] >> 32)) ranges[2 * t + 1] = (uint32_t)(i + 1);
    }
}

/*
 * A.3 forward blend.  out_color [3,H,W]; final_T, n_contrib [H,W].
 * fragile[H*W] (optional): set to 1 for pixels where this arithmetic is itself ambiguous at the
 * 1e-5 level: (i) the exponent is ill-conditioned — `power` is a sum of three terms that can be
 * ~1e2..1e4 each and cancel to ~1 for a needle-like Gaussian far from its centre, so ANY fp32
 * evaluation order (this one, the reference's FMA-contracted CUDA, a sum-of-squares form) carries
 * eps * sum|terms| of rounding noise; the flag is set when the first-order bound
 * eps * sum_k alpha_k T_k S_k (S_k = sum of |terms| of power_k) exceeds 2.5e-6; (ii) a discontinuous decision
 * (alpha floor, T stop, power>0) was within `frag_eps` + the exponent's own rounding bound eps*S_k (relative)
 * of flipping — two correct fp32 implementations may legitimately disagree there.
 */
void orc_render_forward(const orc_settings *st, const uint32_t *ranges, const uint32_t *point_list,
                        const float *xy, const float *conic_opacity, const float *rgb, float *out_color,
                        float *final_T, uint32_t *n_contrib, uint8_t *fragile, float frag_eps)
{
    int W = st->W, H = st->H;
    int gx = (W + TILE - 1) / TILE;
#pragma omp parallel for schedule(dynamic, 8)
    for (int j = 0; j < H; j++) {
        for (int i = 0; i < W; i++) {
            int t = (j / TILE) * gx + (i / TILE);
            uint32_t s = ranges[2 * t], e = ranges[2 * t + 1];
            float T = 1.f, C[3] = {0.f, 0.f, 0.f};
            uint32_t last = 0, n = 0;
            uint8_t frag = 0;
            double cond = 0.0;   /* sum_k alpha_k T_k S_k: sensitivity of the pixel to rounding in the exponents */
            double tamb = 0.0;   /* relative ambiguity of T accumulated from the exponents of the blended Gaussians */
            float pxf = (float)i, pyf = (float)j;
            for (uint32_t k = s; k < e; k++) {
                n++;
                uint32_t g = point_list[k];
                float dx = xy[2 * g] - pxf, dy = xy[2 * g + 1] - pyf;
                const float *co = conic_opacity + 4 * g;
                float power = -0.5f * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
                /* S = sum of |terms| of power: ANY fp32 evaluation of the exponent carries ~eps*S of absolute
                 * error, i.e. eps*S RELATIVE error in alpha, so every decision below is ambiguous within
                 * frag_eps + eps*S (needle-like Gaussians far from their centre: S ~ 1e3, eps*S ~ 1e-4) */
                double S = 0.5 * (fabs((double)co[0] * dx * dx) + fabs((double)co[2] * dy * dy)) +
                           fabs((double)co[1] * dx * dy);
                float Gs;
                if (g_power_f64) {                   /* referee mode: exact exponent, implementation-side ambiguity */
                    double pw = exponent_f64(co, dx, dy);
                    S = well_conditioned_err(pw, S) / 1.1920929e-7;   /* eps * S below is then that bound */
                    power = (float)pw;
                    Gs = (float)exp(pw);
                } else {
                    Gs = expf(power);
                }
                double amb = (double)frag_eps + 1.1920929e-7 * S;
                if (fabs((double)power) <= 1e-6 + 1.1920929e-7 * S) frag = 1;
                if (power > 0.f) continue;
                float alpha = fminf(ALPHA_MAX, co[3] * Gs);
                if (fabs((double)alpha - ALPHA_MIN) <= amb * ALPHA_MIN) frag = 1;
                if (alpha < ALPHA_MIN) continue;
                float test_T = T * (1.f - alpha);
                /* T carries the alphas before it: d(log T) = -sum alpha_k/(1-alpha_k) * d(log alpha_k) */
                double tamb_k = (alpha < ALPHA_MAX ? (double)alpha / (1.0 - (double)alpha) : 0.0) * 1.1920929e-7 * S;
                if (fabs((double)test_T - T_STOP) <= ((double)frag_eps + tamb + tamb_k) * T_STOP) frag = 1;
                if (test_T < T_STOP) break;
                tamb += tamb_k;
                cond += (double)(alpha * T) * S;
                for (int c = 0; c < 3; c++) C[c] += rgb[3 * g + c] * alpha * T;
                T = test_T;
                last = n;
            }
            if (cond * 1.1920929e-7 > 2.5e-6) frag = 1;
            size_t pix = (size_t)j * W + i;
            for (int c = 0; c < 3; c++) out_color[(size_t)c * H * W + pix] = C[c] + T * st->bg[c];
            final_T[pix] = T;
            n_contrib[pix] = last;
            if (fragile) fragile[pix] = frag;
        }
    }
}

/*
 * A.4 backward blend: dL/dout [3,H,W] → per-Gaussian dL/dpix (2), dL/dconic (3:
 * A, B, Cc with the full -Gs*dx*dy*dL/dGs for B), dL/dopacity, dL/drgb (3).
 * Accumulators are double so the oracle value does not depend on pixel order.
 * touched[g] (optional) is set for EVERY Gaussian in the tile list of a pixel flagged in
 * `fragile` (so tests can exclude Gaussians affected by decision flips): the Gaussian whose
 * decision is fragile may be one this replay skips (alpha a hair below 1/255, or behind a
 * fragile stop), and everything behind it sees a different T if the decision flips.
 */
/* How wide the exclusion around a fragile pixel is.  0 (default, what every GPU parity test uses): every Gaussian
 * in the pixel's tile list.  1: only the Gaussians that can reach the pixel at all — alpha there at least half the
 * 1/255 floor (blended, skipped by a hair, or waiting behind a fragile stop) or an exponent that came out positive.
 * The narrow mode is held against a CPU restatement of the kernels' evaluation order in tests/test_oracle.py. */
static int g_touched_narrow = 0;
void orc_set_touched_mode(int narrow) { g_touched_narrow = narrow; }

/* Referee mode for the exponent (default 0 = the fp32 three-term form of SPEC, what every GPU parity test uses).
 * 1: `power` is evaluated in double from the same fp32 conic / pixel inputs, i.e. the exact value of the formula
 * that every fp32 evaluation order approximates.  The ambiguity the oracle then has to report is no longer its OWN
 * rounding (eps * S, S = sum of |terms|, ruinous for elongated Gaussians) but only that of a well-conditioned fp32
 * implementation such as the kernels' sum of squares of the Cholesky factor:
 *     |d power| <= eps * (6 sqrt(|power| S) + 5 |power| + 3)        (worst case; derivation in docs/SPEC.md)
 * of which a quarter is used (g_err_scale): against the CPU restatement of the kernels' evaluation order the
 * measured errors stay 3x below that on scenes with axis ratios from 1:1 to 144:1 (tests/test_oracle.py), and the
 * worst-case figure would flag a third of the pixels of an ordinary scene.  Switching the GPU tests to this mode
 * is the next round's first step. */
void orc_set_power_mode(int f64) { g_power_f64 = f64; }
void orc_set_err_scale(double k) { g_err_scale = k; }

static inline double exponent_f64(const float *co, float dx, float dy)
{
    double x = dx, y = dy;
    return -0.5 * ((double)co[0] * x * x + (double)co[2] * y * y) - (double)co[1] * x * y;
}
static inline double well_conditioned_err(double pw, double S)
{
    double a = fabs(pw);
    return g_err_scale * 1.1920929e-7 * (6.0 * sqrt(a * S) + 5.0 * a + 3.0);
}

void orc_render_backward(const orc_settings *st, const uint32_t *ranges, const uint32_t *point_list,
                         const float *xy, const float *conic_opacity, const float *rgb, const float *final_T,
                         const uint32_t *n_contrib, const float *dL_dout, int P, double *dL_dpix, double *dL_dconic,
                         double *dL_dopacity, double *dL_drgb, const uint8_t *fragile, uint8_t *touched)
{
    int W = st->W, H = st->H;
    int gx = (W + TILE - 1) / TILE;
    (void)P;
#pragma omp parallel for schedule(dynamic, 8)
    for (int j = 0; j < H; j++) {
        for (int i = 0; i < W; i++) {
            size_t pix = (size_t)j * W + i;
            int t = (j / TILE) * gx + (i / TILE);
            uint32_t s = ranges[2 * t];
            float T_final = final_T[pix];
            float T = T_final;
            uint32_t last = n_contrib[pix];
            float Gd[3], accum[3] = {0.f, 0.f, 0.f}, last_col[3] = {0.f, 0.f, 0.f};
            float last_alpha = 0.f;
            double Td = (double)T_final, accum_d[3] = {0., 0., 0.}, last_col_d[3] = {0., 0., 0.}, last_alpha_d = 0.;
            for (int c = 0; c < 3; c++) Gd[c] = dL_dout[(size_t)c * H * W + pix];
            float bg_dot = st->bg[0] * Gd[0] + st->bg[1] * Gd[1] + st->bg[2] * Gd[2];
            float pxf = (float)i, pyf = (float)j;
            int frag = fragile ? fragile[pix] : 0;
            if (frag && touched) {
                uint32_t e = ranges[2 * t + 1];
                for (uint32_t k = s; k < e; k++) {
                    uint32_t g = point_list[k];
                    if (g_touched_narrow) {
                        float dx = xy[2 * g] - pxf, dy = xy[2 * g + 1] - pyf;
                        const float *co = conic_opacity + 4 * g;
                        float power = -0.5f * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
                        if (!(power > 0.f) && co[3] * expf(power) < 0.5f * ALPHA_MIN) continue;
                    }
#pragma omp atomic write
                    touched[g] = 1;
                }
            }
            for (int64_t k = (int64_t)last - 1; k >= 0; k--) {
                uint32_t g = point_list[s + k];
                float dx = xy[2 * g] - pxf, dy = xy[2 * g + 1] - pyf;
                const float *co = conic_opacity + 4 * g;
                float power = -0.5f * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
                float Gs;
                if (g_power_f64) {      /* referee mode: the forward's decisions are replayed on the same values */
                    double pw = exponent_f64(co, dx, dy);
                    power = (float)pw;
                    Gs = (float)exp(pw);
                } else {
                    Gs = expf(power);
                }
                if (power > 0.f) continue;
                float alpha = fminf(ALPHA_MAX, co[3] * Gs);
                if (alpha < ALPHA_MIN) continue;
                if (g_power_f64) {
                    /* Referee mode: the decisions above are the forward's (same fp32 alpha); the VALUES of the
                     * derivative are formed in double from the same fp32 state (conic, xy, rgb, final_T).  With fp32
                     * products the three conic moments of a pair round independently, and the way from dL/dconic to
                     * dL/dcov2D (division by det^2, terms that cancel) amplifies that by the squared axis ratio of
                     * an elongated Gaussian: measured against this branch, the fp32-product replay is off by up to
                     * 1e-5 of a tensor's largest entry and up to 4x the per-Gaussian tolerance of tests/parity.py
                     * on needle scenes — too close to the 1e-4 bar for the thing that referees it. */
                    double ad = fmin((double)ALPHA_MAX, (double)co[3] * exp(exponent_f64(co, dx, dy)));
                    double Gsd = exp(exponent_f64(co, dx, dy));
                    Td = Td / (1.0 - ad);
                    double dchan_d = ad * Td, dla = 0.0;
                    for (int c = 0; c < 3; c++) {
                        double col = rgb[3 * g + c];
                        accum_d[c] = last_alpha_d * last_col_d[c] + (1.0 - last_alpha_d) * accum_d[c];
                        last_col_d[c] = col;
                        dla += (col - accum_d[c]) * (double)Gd[c];
                        double v = dchan_d * (double)Gd[c];
#pragma omp atomic
                        dL_drgb[3 * g + c] += v;
                    }
                    dla *= Td;
                    last_alpha_d = ad;
                    dla += (-(double)T_final / (1.0 - ad)) * (double)bg_dot;
                    double dLdG = (double)co[3] * dla;                   /* U4: straight-through the 0.99 cap */
                    double x = dx, y = dy, gdxd = Gsd * x, gdyd = Gsd * y;
                    double v0 = dLdG * (-gdxd * co[0] - gdyd * co[1]), v1 = dLdG * (-gdyd * co[2] - gdxd * co[1]);
                    double cAd = -0.5 * gdxd * x * dLdG, cBd = -gdxd * y * dLdG, cCd = -0.5 * gdyd * y * dLdG;
                    double vod = Gsd * dla;
#pragma omp atomic
                    dL_dpix[2 * g] += v0;
#pragma omp atomic
                    dL_dpix[2 * g + 1] += v1;
#pragma omp atomic
                    dL_dconic[3 * g] += cAd;
#pragma omp atomic
                    dL_dconic[3 * g + 1] += cBd;
#pragma omp atomic
                    dL_dconic[3 * g + 2] += cCd;
#pragma omp atomic
                    dL_dopacity[g] += vod;
                    continue;
                }
                T = T / (1.f - alpha);
                float dchan = alpha * T;
                float dL_dalpha = 0.f;
                for (int c = 0; c < 3; c++) {
                    float col = rgb[3 * g + c];
                    accum[c] = last_alpha * last_col[c] + (1.f - last_alpha) * accum[c];
                    last_col[c] = col;
                    dL_dalpha += (col - accum[c]) * Gd[c];
                    double v = (double)(dchan * Gd[c]);
#pragma omp atomic
                    dL_drgb[3 * g + c] += v;
                }
                dL_dalpha *= T;
                last_alpha = alpha;
                dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot;
                /* U4: straight-through the 0.99 cap */
                float dL_dG = co[3] * dL_dalpha;
                float gdx = Gs * dx, gdy = Gs * dy;
                float dG_ddx = -gdx * co[0] - gdy * co[1];
                float dG_ddy = -gdy * co[2] - gdx * co[1];
                double v0 = (double)(dL_dG * dG_ddx), v1 = (double)(dL_dG * dG_ddy);
                double cA = (double)(-0.5f * gdx * dx * dL_dG);
                double cB = (double)(-gdx * dy * dL_dG);
                double cC = (double)(-0.5f * gdy * dy * dL_dG);
                double vo = (double)(Gs * dL_dalpha);
#pragma omp atomic
                dL_dpix[2 * g] += v0;
#pragma omp atomic
                dL_dpix[2 * g + 1] += v1;
#pragma omp atomic
                dL_dconic[3 * g] += cA;
#pragma omp atomic
                dL_dconic[3 * g + 1] += cB;
#pragma omp atomic
                dL_dconic[3 * g + 2] += cC;
#pragma omp atomic
                dL_dopacity[g] += vo;
            }
        }
    }
}

/*
 * A.4 per-Gaussian backward: (dL/dpix, dL/dconic, dL/drgb) → dL/dmeans3D,
 * dL/dscales, dL/drotations (or dL/dcov3D packed 6), dL/dshs, dL/dcolors,
 * means2D.grad (U3: dL/dpix * (0.5W, 0.5H), col 2 = 0).
 * Inputs are double (oracle accumulators); math in double here because the
 * result is compared at 1e-4 relative, not bit-exactly.
 */
void orc_preprocess_backward(const orc_settings *st, int P, const int32_t *radii, const float *means3D,
                             const float *scales, const float *rotations, const float *cov3D, /* [P][6] fwd */
                             const float *shs, const uint8_t *clamped, const double *dL_dpix,
                             const double *dL_dconic, const double *dL_drgb, int have_precomp_cov,
                             double *dL_dmeans3D, double *dL_dmeans2D, double *dL_dscales, double *dL_drot,
                             double *dL_dcov3D, double *dL_dshs, double *dL_dcolors)
{
    const float *V = st->V;
    double w0[3] = {V[0], V[1], V[2]}, w1[3] = {V[4], V[5], V[6]};
    double s2 = (double)st->scale * (double)st->scale;
    for (int g = 0; g < P; g++) {
        if (radii[g] <= 0) continue;
        /* colour */
        if (dL_dcolors)
            for (int c = 0; c < 3; c++) dL_dcolors[3 * g + c] = dL_drgb[3 * g + c];
        double dmean[3] = {0, 0, 0};
        if (shs && dL_dshs) {
            int deg = st->sh_degree, M = st->sh_M;
            const float *sh = shs + (size_t)g * M * 3;
            double *dsh = dL_dshs + (size_t)g * M * 3;
            double d[3] = {(double)means3D[3 * g] - st->campos[0], (double)means3D[3 * g + 1] - st->campos[1],
                           (double)means3D[3 * g + 2] - st->campos[2]};
            double len = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
            double x = d[0] / len, y = d[1] / len, z = d[2] / len;
            double dRGBdx[3] = {0, 0, 0}, dRGBdy[3] = {0, 0, 0}, dRGBdz[3] = {0, 0, 0};
            double dL[3];
            for (int c = 0; c < 3; c++) dL[c] = clamped[3 * g + c] ? 0.0 : dL_drgb[3 * g + c];
            for (int c = 0; c < 3; c++) {
                dsh[0 * 3 + c] = SH_C0 * dL[c];
                if (deg > 0) {
                    dsh[1 * 3 + c] = -SH_C1 * y * dL[c];
                    dsh[2 * 3 + c] = SH_C1 * z * dL[c];
                    dsh[3 * 3 + c] = -SH_C1 * x * dL[c];
                    dRGBdx[c] = -SH_C1 * sh[3 * 3 + c];
                    dRGBdy[c] = -SH_C1 * sh[1 * 3 + c];
                    dRGBdz[c] = SH_C1 * sh[2 * 3 + c];
                    if (deg > 1) {
                        double xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
                        dsh[4 * 3 + c] = SH_C2[0] * xy * dL[c];
                        dsh[5 * 3 + c] = SH_C2[1] * yz * dL[c];
                        dsh[6 * 3 + c] = SH_C2[2] * (2.0 * zz - xx - yy) * dL[c];
                        dsh[7 * 3 + c] = SH_C2[3] * xz * dL[c];
                        dsh[8 * 3 + c] = SH_C2[4] * (xx - yy) * dL[c];
                        dRGBdx[c] += SH_C2[0] * y * sh[4 * 3 + c] + SH_C2[2] * 2.0 * -x * sh[6 * 3 + c] +
                                     SH_C2[3] * z * sh[7 * 3 + c] + SH_C2[4] * 2.0 * x * sh[8 * 3 + c];
                        dRGBdy[c] += SH_C2[0] * x * sh[4 * 3 + c] + SH_C2[1] * z * sh[5 * 3 + c] +
                                     SH_C2[2] * 2.0 * -y * sh[6 * 3 + c] + SH_C2[4] * 2.0 * -y * sh[8 * 3 + c];
                        dRGBdz[c] += SH_C2[1] * y * sh[5 * 3 + c] + SH_C2[2] * 2.0 * 2.0 * z * sh[6 * 3 + c] +
                                     SH_C2[3] * x * sh[7 * 3 + c];
                        if (deg > 2) {
                            dsh[9 * 3 + c] = SH_C3[0] * y * (3.0 * xx - yy) * dL[c];
                            dsh[10 * 3 + c] = SH_C3[1] * xy * z * dL[c];
                            dsh[11 * 3 + c] = SH_C3[2] * y * (4.0 * zz - xx - yy) * dL[c];
                            dsh[12 * 3 + c] = SH_C3[3] * z * (2.0 * zz - 3.0 * xx - 3.0 * yy) * dL[c];
                            dsh[13 * 3 + c] = SH_C3[4] * x * (4.0 * zz - xx - yy) * dL[c];
                            dsh[14 * 3 + c] = SH_C3[5] * z * (xx - yy) * dL[c];
                            dsh[15 * 3 + c] = SH_C3[6] * x * (xx - 3.0 * yy) * dL[c];
                            dRGBdx[c] += SH_C3[0] * sh[9 * 3 + c] * 3.0 * 2.0 * xy + SH_C3[1] * sh[10 * 3 + c] * yz +
                                         SH_C3[2] * sh[11 * 3 + c] * -2.0 * xy +
                                         SH_C3[3] * sh[12 * 3 + c] * -3.0 * 2.0 * xz +
                                         SH_C3[4] * sh[13 * 3 + c] * (-3.0 * xx + 4.0 * zz - yy) +
                                         SH_C3[5] * sh[14 * 3 + c] * 2.0 * xz +
                                         SH_C3[6] * sh[15 * 3 + c] * 3.0 * (xx - yy);
                            dRGBdy[c] += SH_C3[0] * sh[9 * 3 + c] * 3.0 * (xx - yy) + SH_C3[1] * sh[10 * 3 + c] * xz +
                                         SH_C3[2] * sh[11 * 3 + c] * (-3.0 * yy + 4.0 * zz - xx) +
                                         SH_C3[3] * sh[12 * 3 + c] * -3.0 * 2.0 * yz +
                                         SH_C3[4] * sh[13 * 3 + c] * -2.0 * xy +
                                         SH_C3[5] * sh[14 * 3 + c] * -2.0 * yz +
                                         SH_C3[6] * sh[15 * 3 + c] * -3.0 * 2.0 * xy;
                            dRGBdz[c] += SH_C3[1] * sh[10 * 3 + c] * xy + SH_C3[2] * sh[11 * 3 + c] * 4.0 * 2.0 * yz +
                                         SH_C3[3] * sh[12 * 3 + c] * 3.0 * (2.0 * zz - xx - yy) +
                                         SH_C3[4] * sh[13 * 3 + c] * 4.0 * 2.0 * xz +
                                         SH_C3[5] * sh[14 * 3 + c] * (xx - yy);
                        }
                    }
                }
            }
            /* through the direction normalisation */
            double ddir[3] = {0, 0, 0};
            for (int c = 0; c < 3; c++) {
                ddir[0] += dRGBdx[c] * dL[c];
                ddir[1] += dRGBdy[c] * dL[c];
                ddir[2] += dRGBdz[c] * dL[c];
            }
            double dot = x * ddir[0] + y * ddir[1] + z * ddir[2];
            dmean[0] += (ddir[0] - x * dot) / len;
            dmean[1] += (ddir[1] - y * dot) / len;
            dmean[2] += (ddir[2] - z * dot) / len;
        }

        /* position: pix = (V[:2,:3] p + V[:2,3] - min) * scale - 0.5 */
        double gpx = dL_dpix[2 * g], gpy = dL_dpix[2 * g + 1];
        for (int k = 0; k < 3; k++) dmean[k] += (double)st->scale * (w0[k] * gpx + w1[k] * gpy);
        for (int k = 0; k < 3; k++) dL_dmeans3D[3 * g + k] = dmean[k];
        dL_dmeans2D[3 * g] = gpx * 0.5 * st->W; /* U3 */
        dL_dmeans2D[3 * g + 1] = gpy * 0.5 * st->H;
        dL_dmeans2D[3 * g + 2] = 0.0;

        /* conic → cov2D (a,b,c) */
        float af, bf, cf;
        cov2d_ortho(cov3D + 6 * g, V, st->scale, &af, &bf, &cf);
        double a = af, b = bf, c = cf;
        double det = a * c - b * b;
        double gA = dL_dconic[3 * g], gB = dL_dconic[3 * g + 1], gC = dL_dconic[3 * g + 2];
        double d2 = 1.0 / (det * det);
        double da = d2 * (-c * c * gA + b * c * gB - b * b * gC);
        double db = d2 * (2.0 * b * c * gA - (det + 2.0 * b * b) * gB + 2.0 * a * b * gC);
        double dc = d2 * (-b * b * gA + a * b * gB - a * a * gC);
        /* cov2D = s2 * (w0'Σw0, w0'Σw1, w1'Σw1) (+0.3) → G[k][l] = dL/dΣ[k][l] (independent entries) */
        da *= s2; db *= s2; dc *= s2;
        double Gm[3][3];
        for (int k = 0; k < 3; k++)
            for (int l = 0; l < 3; l++) Gm[k][l] = da * w0[k] * w0[l] + db * w0[k] * w1[l] + dc * w1[k] * w1[l];
        if (have_precomp_cov) {
            if (dL_dcov3D) {
                dL_dcov3D[6 * g + 0] = Gm[0][0];
                dL_dcov3D[6 * g + 1] = Gm[0][1] + Gm[1][0];
                dL_dcov3D[6 * g + 2] = Gm[0][2] + Gm[2][0];
                dL_dcov3D[6 * g + 3] = Gm[1][1];
                dL_dcov3D[6 * g + 4] = Gm[1][2] + Gm[2][1];
                dL_dcov3D[6 * g + 5] = Gm[2][2];
            }
            continue;
        }
        /* Σ = M M^T, M = R diag(mod*s):  dL/dM = (G + G^T) M */
        float Rf[9];
        quat_to_rot(rotations + 4 * g, Rf);
        double mod = st->scale_modifier;
        double sv[3] = {mod * scales[3 * g], mod * scales[3 * g + 1], mod * scales[3 * g + 2]};
        double M[3][3], dM[3][3];
        for (int i = 0; i < 3; i++)
            for (int j2 = 0; j2 < 3; j2++) M[i][j2] = (double)Rf[3 * i + j2] * sv[j2];
        for (int i = 0; i < 3; i++)
            for (int j2 = 0; j2 < 3; j2++) {
                double acc = 0;
                for (int k = 0; k < 3; k++) acc += (Gm[i][k] + Gm[k][i]) * M[k][j2];
                dM[i][j2] = acc;
            }
        double gR[3][3];
        for (int j2 = 0; j2 < 3; j2++) {
            double acc = 0;
            for (int i = 0; i < 3; i++) {
                acc += dM[i][j2] * (double)Rf[3 * i + j2];
                gR[i][j2] = dM[i][j2] * sv[j2];
            }
            dL_dscales[3 * g + j2] = acc * mod;
        }
        double r = rotations[4 * g], x = rotations[4 * g + 1], y = rotations[4 * g + 2], z = rotations[4 * g + 3];
        dL_drot[4 * g + 0] = 2.0 * (-z * gR[0][1] + y * gR[0][2] + z * gR[1][0] - x * gR[1][2] - y * gR[2][0] + x * gR[2][1]);
        dL_drot[4 * g + 1] = 2.0 * (y * gR[0][1] + z * gR[0][2] + y * gR[1][0] - 2.0 * x * gR[1][1] - r * gR[1][2] +
                                    z * gR[2][0] + r * gR[2][1] - 2.0 * x * gR[2][2]);
        dL_drot[4 * g + 2] = 2.0 * (-2.0 * y * gR[0][0] + x * gR[0][1] + r * gR[0][2] + x * gR[1][0] + z * gR[1][2] -
                                    r * gR[2][0] + z * gR[2][1] - 2.0 * y * gR[2][2]);
        dL_drot[4 * g + 3] = 2.0 * (-2.0 * z * gR[0][0] - r * gR[0][1] + x * gR[0][2] + r * gR[1][0] -
                                    2.0 * z * gR[1][1] + y * gR[1][2] + x * gR[2][0] + y * gR[2][1]);
    }
}

int orc_abi_version(void) { return 1; }
