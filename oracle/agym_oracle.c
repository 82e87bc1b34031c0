/*
 * agym_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, single-threaded CPU restatement of the arithmetic on Active-Gym's
 * observation hot path.  It exists only so that tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg can check / time the CUDA path against it.  Nothing under
 * active_gym_b200/ may include, link or call it.
 *
 * Parity status: PINNED.  Every function here is checked (tests/test_oracle_golden.py)
 * against fixtures in tests/golden/ that were produced by running the UNMODIFIED
 * reference source (fov_env.py / atari_env.py / dmc_env.py) under simulator stubs with
 * this container's OpenCV 4.13 / torch 2.11 / torchvision 0.26 (oracle/make_golden.py),
 * and — when /root/reference is present — against the reference itself
 * (tests/test_oracle_vs_reference.py).
 * One input is NOT pinned: ALE's palette RGB->gray inside getScreenGrayscale() lives
 * in the absent third-party atari-py (setup.py:13, version unpinned); the reference
 * boundary is the gray screen, so the RGB luma used for 210x160x3 synthetic frames is
 * a declared stand-in (OpenCV's RGB2GRAY fixed-point formula).
 *
 * Reference lines restated (paths relative to /root/reference/active_gym/):
 *   or_cv2_resize_linear_u8   atari_env.py:74   cv2.resize(gray, obs_size, INTER_LINEAR)
 *   or_luma_u8                dmc_env.py:182    cv2.cvtColor(obs, COLOR_BGR2GRAY)
 *   or_aa_resize_*            fov_env.py:120,182,248,278,366-368  torchvision Resize
 *                             == ATen upsample_bilinear2d_aa (antialias=True)
 *   or_update_loc             fov_env.py:166-170,193-199 (fixed), :270-274,314-324 (flexible)
 *   or_ingest_atari           atari_env.py:80-82,91,111-114,121-133,143
 *   or_ingest_dmc             dmc_env.py:175-183,193-195,206-207,228-230
 *   or_observe_fixed          fov_env.py:172-185
 *   or_observe_peripheral     fov_env.py:375-388
 *   or_observe_flexible       fov_env.py:276-298
 *
 * Data model (shared with the CUDA path so the two can be compared byte for byte):
 *   ring  u8  [N][K][S_h][S_w]   frame stack, slot `head[n]` holds the NEWEST frame and
 *                                logical order oldest->newest is (head+1)%K ... head
 *   head  i32 [N]
 *   loc   i32 [N][2]  (row, col) of the fovea's upper-left corner
 *   res   i32 [N][2]  (rows, cols) of the flexible fovea
 * Reference observations are normalised floats f32(u)/255; the oracle works on the
 * u8 value u (SURVEY.md §8 "u8-equivalent") and returns resampled pixels as doubles in
 * u8 units (= 255 * reference value) so the caller can bound the rounding error.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define OR_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------
 * cv2.resize(src u8 1-channel, (dw, dh), INTER_LINEAR): 11-bit fixed point, horizontal
 * pass to int32 then vertical pass (OpenCV imgproc resize.cpp: resizeGeneric_ /
 * HResizeLinear / VResizeLinear<uchar,int,short,FixedPtCast<...,22>>).
 * ---------------------------------------------------------------------------------- */
static void or_cv2_axis(int n_src, int n_dst, int zero_frac_low, int *ofs, short *c0, short *c1)
{
    double scale = 1.0 / ((double)n_dst / (double)n_src);
    for (int d = 0; d < n_dst; ++d) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        f -= (float)s;
        if (zero_frac_low) { /* x axis: the fraction is zeroed at both borders */
            if (s < 0) { s = 0; f = 0.f; }
            if (s >= n_src - 1) { s = n_src - 1; f = 0.f; }
        }
        ofs[d] = s;
        c0[d] = (short)lrintf((1.f - f) * 2048.f); /* cvRound: round half to even */
        c1[d] = (short)lrintf(f * 2048.f);
    }
}

static inline int or_clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

OR_API void or_cv2_resize_linear_u8(const uint8_t *src, int sh, int sw, uint8_t *dst, int dh, int dw)
{
    int *xo = malloc(sizeof(int) * dw), *yo = malloc(sizeof(int) * dh);
    short *xa = malloc(2 * dw), *xb = malloc(2 * dw), *ya = malloc(2 * dh), *yb = malloc(2 * dh);
    int *h0 = malloc(sizeof(int) * dw), *h1 = malloc(sizeof(int) * dw);
    or_cv2_axis(sw, dw, 1, xo, xa, xb);
    or_cv2_axis(sh, dh, 0, yo, ya, yb);
    for (int y = 0; y < dh; ++y) {
        const uint8_t *r0 = src + (size_t)or_clampi(yo[y], 0, sh - 1) * sw;
        const uint8_t *r1 = src + (size_t)or_clampi(yo[y] + 1, 0, sh - 1) * sw;
        for (int x = 0; x < dw; ++x) {
            int s = xo[x], s1 = s + 1 < sw ? s + 1 : sw - 1;
            h0[x] = r0[s] * xa[x] + r0[s1] * xb[x];
            h1[x] = r1[s] * xa[x] + r1[s1] * xb[x];
        }
        for (int x = 0; x < dw; ++x) {
            int v = (((ya[y] * (h0[x] >> 4)) >> 16) + ((yb[y] * (h1[x] >> 4)) >> 16) + 2) >> 2;
            dst[(size_t)y * dw + x] = (uint8_t)or_clampi(v, 0, 255);
        }
    }
    free(xo); free(yo); free(xa); free(xb); free(ya); free(yb); free(h0); free(h1);
}

/* cv2.cvtColor(8UC3 -> GRAY): (w0*c0 + w1*c1 + w2*c2 + 2^14) >> 15 with the 15-bit
 * weights {B:3735, G:19235, R:9798}.  dmc_env.py:182 feeds an RGB image to BGR2GRAY, so
 * channel 0 (R) gets 3735 there; the caller passes the weights per channel position. */
OR_API void or_luma_u8(const uint8_t *src3, size_t npix, int w0, int w1, int w2, uint8_t *dst)
{
    for (size_t i = 0; i < npix; ++i)
        dst[i] = (uint8_t)((w0 * src3[3 * i] + w1 * src3[3 * i + 1] + w2 * src3[3 * i + 2] + 16384) >> 15);
}

/* ------------------------------------------------------------------------------------
 * torchvision.transforms.Resize((oh, ow)) on a float tensor == F.interpolate(bilinear,
 * align_corners=False, antialias=True) == ATen separable_upsample_generic_Nd_kernel_impl
 * with the triangle filter: W pass, then H pass, a pass is skipped when in == out.
 * Weights follow ATen _compute_indices_min_size_weights_aa.  `use_f32` evaluates in
 * float (the reference's DMC steady state is float32), else double (Atari, float64).
 * ---------------------------------------------------------------------------------- */
typedef struct { int n_out, maxk; int *xmin, *xsize; double *w; } or_aa_axis_t;

/* antialias=False: what torchvision's Resize did on tensors before antialias=True became its default (SURVEY.md
 * section 8c, library-version sensitivity) == F.interpolate(bilinear, align_corners=False, antialias=False) == ATen
 * upsample_bilinear2d: src = scale * (dst + 0.5) - 0.5 clamped at 0, taps {floor(src), +1}, weights {1 - l, l}. */
static int g_or_antialias = 1;
OR_API void or_set_antialias(int on) { g_or_antialias = on ? 1 : 0; }

static void or_bilinear_axis_build(or_aa_axis_t *t, int n_in, int n_out, int use_f32)
{
    t->n_out = n_out; t->maxk = 2;
    t->xmin = malloc(sizeof(int) * n_out); t->xsize = malloc(sizeof(int) * n_out);
    t->w = calloc((size_t)n_out * 2, sizeof(double));
    for (int i = 0; i < n_out; ++i) {
        double l1; int i0;
        if (use_f32) {
            float scale = (float)n_in / (float)n_out;
            float src = scale * ((float)i + 0.5f) - 0.5f; if (src < 0.f) src = 0.f;
            i0 = (int)src; if (i0 > n_in - 1) i0 = n_in - 1;
            l1 = (double)(src - (float)i0);
            t->w[2 * i] = (double)(1.0f - (float)l1);
        } else {
            double scale = (double)n_in / (double)n_out;
            double src = scale * (i + 0.5) - 0.5; if (src < 0.0) src = 0.0;
            i0 = (int)src; if (i0 > n_in - 1) i0 = n_in - 1;
            l1 = src - i0;
            t->w[2 * i] = 1.0 - l1;
        }
        t->xmin[i] = i0;
        if (i0 < n_in - 1) { t->xsize[i] = 2; t->w[2 * i + 1] = l1; }
        else { t->xsize[i] = 1; t->w[2 * i] = 1.0; }
    }
}

static void or_aa_axis_build(or_aa_axis_t *t, int n_in, int n_out, int use_f32)
{
    if (!g_or_antialias) { or_bilinear_axis_build(t, n_in, n_out, use_f32); return; }
    double scale_d = (double)n_in / (double)n_out;
    float scale_f = (float)n_in / (float)n_out;
    double support_d = scale_d >= 1.0 ? scale_d : 1.0;
    float support_f = scale_f >= 1.0f ? scale_f : 1.0f;
    int maxk = (int)ceil(use_f32 ? (double)support_f : support_d) * 2 + 1;
    t->n_out = n_out; t->maxk = maxk;
    t->xmin = malloc(sizeof(int) * n_out); t->xsize = malloc(sizeof(int) * n_out);
    t->w = calloc((size_t)n_out * maxk, sizeof(double));
    for (int i = 0; i < n_out; ++i) {
        double *w = t->w + (size_t)i * maxk;
        int xmin, xsize;
        if (use_f32) {
            float center = scale_f * ((float)i + 0.5f);
            float inv = scale_f >= 1.0f ? 1.0f / scale_f : 1.0f, total = 0.f;
            /* C++ promotion order of the ATen expression: float op, then + 0.5 (double) */
            xmin = (int)((double)(center - support_f) + 0.5); if (xmin < 0) xmin = 0;
            xsize = (int)((double)(center + support_f) + 0.5); if (xsize > n_in) xsize = n_in;
            xsize -= xmin; xsize = or_clampi(xsize, 0, maxk);
            for (int j = 0; j < xsize; ++j) {
                float x = (float)(((double)((float)(j + xmin) - center) + 0.5) * (double)inv);
                float v = fabsf(x) < 1.0f ? 1.0f - fabsf(x) : 0.0f;
                w[j] = v; total += v;
            }
            if (total != 0.f) for (int j = 0; j < xsize; ++j) w[j] = (float)((float)w[j] / total);
        } else {
            double center = scale_d * (i + 0.5);
            double inv = scale_d >= 1.0 ? 1.0 / scale_d : 1.0, total = 0.0;
            xmin = (int)(center - support_d + 0.5); if (xmin < 0) xmin = 0;
            xsize = (int)(center + support_d + 0.5); if (xsize > n_in) xsize = n_in;
            xsize -= xmin; xsize = or_clampi(xsize, 0, maxk);
            for (int j = 0; j < xsize; ++j) {
                double x = ((j + xmin) - center + 0.5) * inv;
                double v = fabs(x) < 1.0 ? 1.0 - fabs(x) : 0.0;
                w[j] = v; total += v;
            }
            if (total != 0.0) for (int j = 0; j < xsize; ++j) w[j] /= total;
        }
        t->xmin[i] = xmin; t->xsize[i] = xsize;
    }
}

static void or_aa_axis_free(or_aa_axis_t *t) { free(t->xmin); free(t->xsize); free(t->w); }

/* src/dst are planes with explicit row strides (elements). */
static void or_aa_resize_plane(const double *src, int ih, int iw, int sstride,
                               double *dst, int oh, int ow, int dstride, int use_f32)
{
    const double *cur = src; int cur_stride = sstride;
    double *tmp = NULL;
    if (iw != ow) { /* horizontal pass first (contiguous dim) */
        or_aa_axis_t ax; or_aa_axis_build(&ax, iw, ow, use_f32);
        int need_v = (ih != oh);
        double *out = need_v ? (tmp = malloc(sizeof(double) * (size_t)ih * ow)) : dst;
        int ostride = need_v ? ow : dstride;
        for (int y = 0; y < ih; ++y)
            for (int x = 0; x < ow; ++x) {
                const double *w = ax.w + (size_t)x * ax.maxk;
                const double *s = cur + (size_t)y * cur_stride + ax.xmin[x];
                if (use_f32) { float a = 0.f; for (int j = 0; j < ax.xsize[x]; ++j) a += (float)s[j] * (float)w[j]; out[(size_t)y * ostride + x] = a; }
                else { double a = 0.0; for (int j = 0; j < ax.xsize[x]; ++j) a += s[j] * w[j]; out[(size_t)y * ostride + x] = a; }
            }
        or_aa_axis_free(&ax);
        cur = out; cur_stride = ostride;
    }
    if (ih != oh) {
        or_aa_axis_t ay; or_aa_axis_build(&ay, ih, oh, use_f32);
        for (int y = 0; y < oh; ++y) {
            const double *w = ay.w + (size_t)y * ay.maxk;
            for (int x = 0; x < ow; ++x) {
                const double *s = cur + (size_t)ay.xmin[y] * cur_stride + x;
                if (use_f32) { float a = 0.f; for (int j = 0; j < ay.xsize[y]; ++j) a += (float)s[(size_t)j * cur_stride] * (float)w[j]; dst[(size_t)y * dstride + x] = a; }
                else { double a = 0.0; for (int j = 0; j < ay.xsize[y]; ++j) a += s[(size_t)j * cur_stride] * w[j]; dst[(size_t)y * dstride + x] = a; }
            }
        }
        or_aa_axis_free(&ay);
    } else if (iw == ow) {
        for (int y = 0; y < oh; ++y) memcpy(dst + (size_t)y * dstride, src + (size_t)y * sstride, sizeof(double) * ow);
    }
    free(tmp);
}

OR_API void or_aa_resize_f64(const double *src, int ih, int iw, double *dst, int oh, int ow, int use_f32)
{
    or_aa_resize_plane(src, ih, iw, iw, dst, oh, ow, ow, use_f32);
}

/* ------------------------------------------------------------------------------------
 * fov_loc / fov_res update.  np.clip first, then np.rint (round half to even):
 *   absolute: loc = rint(clip(a, 0, S - f))                         fov_env.py:166-167,193-195
 *   relative: d = rint(clip(a, lo, hi)); loc = rint(clip(loc+d, 0, S - f))   :169-170,196-199
 * Flexible (fov_env.py:300-324): action_type 0 moves the window as above but clamps with
 * S - res (:270-271); action_type 1 sets res = action (no clip, no rint: the reference
 * needs integers there) then re-clamps loc (:322-324).
 * `res` == NULL means the fixed fovea (window = fov).  `atype` == NULL means all FOV_LOC.
 * ---------------------------------------------------------------------------------- */
static inline double or_clipd(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }

OR_API void or_update_loc(const double *action, const int32_t *atype, int32_t *loc, int32_t *res, int n,
                          int relative, double lo, double hi, int sh, int sw, int fh, int fw)
{
    const int S[2] = {sh, sw}, F[2] = {fh, fw};
    for (int e = 0; e < n; ++e) {
        int t = atype ? atype[e] : 0;
        if (t == 1) {
            for (int a = 0; a < 2; ++a) res[2 * e + a] = (int32_t)action[2 * e + a];
            for (int a = 0; a < 2; ++a)
                loc[2 * e + a] = (int32_t)nearbyint(or_clipd((double)loc[2 * e + a], 0.0, (double)(S[a] - res[2 * e + a])));
            continue;
        }
        for (int a = 0; a < 2; ++a) {
            double win = res ? (double)res[2 * e + a] : (double)F[a];
            double top = (double)S[a] - win;
            double v = action[2 * e + a];
            if (relative) v = (double)loc[2 * e + a] + nearbyint(or_clipd(v, lo, hi));
            loc[2 * e + a] = (int32_t)nearbyint(or_clipd(v, 0.0, top));
        }
    }
}

/* ------------------------------------------------------------------------------------
 * Ingest: the base env's per-step observation code.
 * flags[n]: bit0 frame A valid (atari_env.py:125-126, t==2), bit1 frame B valid (:127-128,
 * t==3; an early `done` leaves a slot zero, :129-131), bit2 hard reset = zero-fill the
 * stack first (:80-82,91), bit3 env idle (nothing pushed).  A reset pushes one
 * un-pooled frame (:111-112) = flags A only.  Pooled frame = max(A or 0, B or 0) (:132).
 * channels: 1 = gray screen (the reference's ALE boundary) or 3 = RGB + luma weights.
 * ---------------------------------------------------------------------------------- */
static void or_push(uint8_t *ring, int32_t *head, int K, size_t plane, int hard_reset)
{
    if (hard_reset) memset(ring, 0, plane * K);
    *head = (*head + 1) % K;
}

OR_API void or_ingest_atari(const uint8_t *fa, const uint8_t *fb, const uint8_t *flags,
                            uint8_t *ring, int32_t *head, int n, int K, int rh, int rw, int ch,
                            int sh, int sw, int w0, int w1, int w2)
{
    size_t fsz = (size_t)rh * rw * ch, plane = (size_t)sh * sw;
    uint8_t *gray = malloc((size_t)rh * rw), *ra = malloc(plane), *rb = malloc(plane);
    for (int e = 0; e < n; ++e) {
        int fl = flags[e];
        if (fl & 8) continue;
        memset(ra, 0, plane); memset(rb, 0, plane);
        for (int which = 0; which < 2; ++which) {
            if (!(fl & (1 << which))) continue;
            const uint8_t *f = (which ? fb : fa) + fsz * e;
            const uint8_t *g = f;
            if (ch == 3) { or_luma_u8(f, (size_t)rh * rw, w0, w1, w2, gray); g = gray; }
            or_cv2_resize_linear_u8(g, rh, rw, which ? rb : ra, sh, sw);
        }
        uint8_t *r = ring + plane * K * e;
        or_push(r, head + e, K, plane, fl & 4);
        uint8_t *slot = r + plane * head[e];
        for (size_t i = 0; i < plane; ++i) slot[i] = ra[i] > rb[i] ? ra[i] : rb[i];
    }
    free(gray); free(ra); free(rb);
}

/* DMC: one rendered frame per step at obs_size, luma, push (no max-pool, no resize). */
OR_API void or_ingest_dmc(const uint8_t *f, const uint8_t *flags, uint8_t *ring, int32_t *head,
                          int n, int K, int sh, int sw, int w0, int w1, int w2)
{
    size_t plane = (size_t)sh * sw;
    for (int e = 0; e < n; ++e) {
        int fl = flags[e];
        if (fl & 8) continue;
        uint8_t *r = ring + plane * K * e;
        or_push(r, head + e, K, plane, fl & 4);
        or_luma_u8(f + plane * 3 * e, plane, w0, w1, w2, r + plane * head[e]);
    }
}

/* np.stack(state_buffer): oldest -> newest (atari_env.py:143, dmc_env.py:230). */
OR_API void or_stack(const uint8_t *ring, const int32_t *head, uint8_t *out, int n, int K, int sh, int sw)
{
    size_t plane = (size_t)sh * sw;
    for (int e = 0; e < n; ++e)
        for (int k = 0; k < K; ++k)
            memcpy(out + plane * ((size_t)K * e + k), ring + plane * ((size_t)K * e + (head[e] + 1 + k) % K), plane);
}

/* ------------------------------------------------------------------------------------
 * FixedFovealEnv._get_fov_state (fov_env.py:172-185).
 * variant 0: crop -> out_u8 [N][K][fh][fw]
 * variant 1: mask_out -> out_u8 [N][K][S][S] zero with the crop pasted in place
 * variant 2: resize_to_full -> out_f64 [N][K][S][S] in u8 units (torchvision Resize of
 *            the crop to obs_size; upscaling, so the antialias flag has no effect)
 * ---------------------------------------------------------------------------------- */
OR_API void or_observe_fixed(const uint8_t *ring, const int32_t *head, const int32_t *loc,
                             uint8_t *out_u8, double *out_f64, int n, int K, int sh, int sw,
                             int fh, int fw, int variant, int use_f32)
{
    size_t plane = (size_t)sh * sw;
    double *crop = malloc(sizeof(double) * fh * fw);
    if (variant == 1) memset(out_u8, 0, plane * K * n);
    for (int e = 0; e < n; ++e)
        for (int k = 0; k < K; ++k) {
            const uint8_t *src = ring + plane * ((size_t)K * e + (head[e] + 1 + k) % K);
            int r0 = loc[2 * e], c0 = loc[2 * e + 1];
            size_t ok = (size_t)K * e + k;
            for (int y = 0; y < fh; ++y)
                for (int x = 0; x < fw; ++x) {
                    uint8_t v = src[(size_t)(r0 + y) * sw + c0 + x];
                    if (variant == 0) out_u8[ok * fh * fw + (size_t)y * fw + x] = v;
                    else if (variant == 1) out_u8[ok * plane + (size_t)(r0 + y) * sw + c0 + x] = v;
                    else crop[(size_t)y * fw + x] = v;
                }
            if (variant == 2) or_aa_resize_plane(crop, fh, fw, fw, out_f64 + ok * plane, sh, sw, sw, use_f32);
        }
    free(crop);
}

/* FixedFovealPeripheralEnv._get_fov_state (fov_env.py:379-388):
 * out = Resize(S)(Resize(p)(full)); out[fovea window] = full[fovea window].
 * out_f64 [N][K][S][S] in u8 units; pasted pixels are exact integers. */
OR_API void or_observe_peripheral(const uint8_t *ring, const int32_t *head, const int32_t *loc,
                                  double *out_f64, int n, int K, int sh, int sw, int fh, int fw,
                                  int ph, int pw, int use_f32)
{
    size_t plane = (size_t)sh * sw;
    double *full = malloc(sizeof(double) * plane), *small = malloc(sizeof(double) * ph * pw);
    for (int e = 0; e < n; ++e)
        for (int k = 0; k < K; ++k) {
            const uint8_t *src = ring + plane * ((size_t)K * e + (head[e] + 1 + k) % K);
            double *dst = out_f64 + plane * ((size_t)K * e + k);
            for (size_t i = 0; i < plane; ++i) full[i] = src[i];
            or_aa_resize_plane(full, sh, sw, sw, small, ph, pw, pw, use_f32);
            or_aa_resize_plane(small, ph, pw, pw, dst, sh, sw, sw, use_f32);
            int r0 = loc[2 * e], c0 = loc[2 * e + 1];
            for (int y = 0; y < fh; ++y)
                for (int x = 0; x < fw; ++x)
                    dst[(size_t)(r0 + y) * sw + c0 + x] = src[(size_t)(r0 + y) * sw + c0 + x];
        }
    free(full); free(small);
}

/* FlexibleFovealEnv._get_fov_state (fov_env.py:283-298).  Crop rh x rw at loc; iff
 * rh > fh (row dimension only, :286) blur = Resize((fh,fw)) then Resize((rh,rw)) (:276-280).
 * variant 0: padded crop -> out_f64 [N][K][P_h][P_w], patch in the top-left corner, rest 0
 *            (the reference returns a variable-shape (K,rh,rw) array)
 * variant 1: mask_out -> out_f64 [N][K][S][S], patch pasted at loc, rest 0
 * variant 2: resize_to_full -> out_f64 [N][K][S][S] = Resize(S)(patch)
 * All values in u8 units; un-blurred patches are exact integers. */
OR_API void or_observe_flexible(const uint8_t *ring, const int32_t *head, const int32_t *loc,
                                const int32_t *res, double *out_f64, int n, int K, int sh, int sw,
                                int fh, int fw, int variant, int pad_h, int pad_w, int use_f32)
{
    size_t plane = (size_t)sh * sw;
    size_t oplane = variant == 0 ? (size_t)pad_h * pad_w : plane;
    int ow = variant == 0 ? pad_w : sw;
    double *patch = malloc(sizeof(double) * plane), *mid = malloc(sizeof(double) * fh * fw);
    double *blur = malloc(sizeof(double) * plane);
    memset(out_f64, 0, sizeof(double) * oplane * K * n);
    for (int e = 0; e < n; ++e) {
        int rh = res[2 * e], rw = res[2 * e + 1], r0 = loc[2 * e], c0 = loc[2 * e + 1];
        for (int k = 0; k < K; ++k) {
            const uint8_t *src = ring + plane * ((size_t)K * e + (head[e] + 1 + k) % K);
            double *dst = out_f64 + oplane * ((size_t)K * e + k);
            for (int y = 0; y < rh; ++y)
                for (int x = 0; x < rw; ++x) patch[(size_t)y * rw + x] = src[(size_t)(r0 + y) * sw + c0 + x];
            const double *p = patch;
            if (rh > fh) {
                or_aa_resize_plane(patch, rh, rw, rw, mid, fh, fw, fw, use_f32);
                or_aa_resize_plane(mid, fh, fw, fw, blur, rh, rw, rw, use_f32);
                p = blur;
            }
            if (variant == 2) or_aa_resize_plane(p, rh, rw, rw, dst, sh, sw, sw, use_f32);
            else {
                int oy = variant == 1 ? r0 : 0, ox = variant == 1 ? c0 : 0;
                for (int y = 0; y < rh; ++y)
                    for (int x = 0; x < rw; ++x) dst[(size_t)(oy + y) * ow + ox + x] = p[(size_t)y * rw + x];
            }
        }
    }
    free(patch); free(mid); free(blur);
}
