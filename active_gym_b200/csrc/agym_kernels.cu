// sm_100a kernels of the active-perception observation path.
//
// Every kernel here is a byte/integer streaming kernel bounded by HBM bandwidth or by
// instruction issue — there is no dense contraction on this path, so no tensor cores.
// Mapping: one CTA per environment; the env's source tile is staged in shared memory with
// coalesced 16-byte loads, all variable-offset (fovea) accesses are served from shared memory
// or from L2-resident rows, and outputs leave as whole 4/16-byte words.
//
// Reference semantics (file:line relative to /root/reference/active_gym/) are cited at each
// kernel; the arithmetic must stay bit-identical to oracle/agym_oracle.c for the integer
// stages (cv2 resize, luma, max, stack, crop, mask, paste) and within 0.5 u8 LSB + fp32
// round-off of it for the antialiased resamples.
#include "agym_kernels.cuh"
#ifndef AGYM_EXPERIMENT
#define AGYM_EXPERIMENT 0
#endif

#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "../../include/agym_b200.h"

namespace agym {

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ uint4 ld_stream128(const void *p) {
    // read-once data (raw simulator frames): bypass L1 allocation
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

struct FastDiv {
    uint32_t magic;
    int32_t d;
    __device__ __forceinline__ explicit FastDiv(int32_t dd) : magic(0xFFFFFFFFu / (uint32_t)dd + 1u), d(dd) {}
    // multiplier from a table of 0xFFFFFFFF / d + 1 (d < 256) instead of an integer division
    __device__ __forceinline__ FastDiv(int32_t dd, const uint32_t *table) : magic(table[dd]), d(dd) {}
    // exact for n * d < 2^32 (all indices here are < 2^24, divisors < 2^8)
    __device__ __forceinline__ int32_t div(int32_t n) const { return d == 1 ? n : (int32_t)__umulhi((uint32_t)n, magic); }
};

__device__ __forceinline__ int clip_rint(double v, double lo, double hi) {
    // np.rint(np.clip(v, lo, hi)): clip first, then round half to even (fov_env.py:166-170)
    return __double2int_rn(fmin(fmax(v, lo), hi));
}

__device__ __forceinline__ uint32_t quant_u8(float v) {
    int q = __float2int_rn(v);
    return (uint32_t)min(max(q, 0), 255);
}

// fov_loc update of the fixed fovea (fov_env.py:187-199), split into the global loads and the
// arithmetic so that a persistent kernel can issue the loads one env ahead.
struct LocIn {
    double a0, a1;
    int r, c, mode;
};

__device__ __forceinline__ LocIn load_loc_in(int n, const double *action, const uint8_t *ctrl, const int32_t *loc) {
    LocIn v;
    v.mode = ctrl ? ctrl[n] : AGYM_FOV_APPLY;
    v.r = loc[2 * n];
    v.c = loc[2 * n + 1];
    v.a0 = action ? action[2 * n] : 0.0;
    v.a1 = action ? action[2 * n + 1] : 0.0;
    return v;
}

__device__ __forceinline__ void apply_loc(const DevPlan &p, const LocIn &v, int &r, int &c) {
    r = v.r;
    c = v.c;
    if (v.mode == AGYM_FOV_RESET) {
        r = p.init_r;
        c = p.init_c;
    } else if (v.mode == AGYM_FOV_APPLY) {
        double a0 = v.a0, a1 = v.a1;
        if (p.relative) {
            a0 = (double)(r + clip_rint(a0, p.lo, p.hi));
            a1 = (double)(c + clip_rint(a1, p.lo, p.hi));
        }
        r = clip_rint(a0, 0.0, (double)(p.S_h - p.f_h));
        c = clip_rint(a1, 0.0, (double)(p.S_w - p.f_w));
    }
}

// STORE: write the new loc back (one thread per env); otherwise only compute it.
template <bool STORE = true>
__device__ __forceinline__ void update_loc_fixed(const DevPlan &p, int n, const double *action, const uint8_t *ctrl,
                                                 int32_t *loc, int &r, int &c) {
    apply_loc(p, load_loc_in(n, action, ctrl, loc), r, c);
    if (STORE) {
        loc[2 * n] = r;
        loc[2 * n + 1] = c;
    }
}

// ------------------------------------------------------------------ separable resamples
// dst[r][c] = sum_j w[c][j] * src[r][xmin[c] + j]      (pass along the contiguous axis)
template <typename SrcT>
__device__ __forceinline__ void resample_w(const SrcT *src, int sstride, float *dst, int dstride, int rows,
                                           const AxisRef &ax, int tid, int nt) {
    const int total = rows * ax.n_out;
    const FastDiv fd(ax.n_out);
    for (int i = tid; i < total; i += nt) {
        const int r = fd.div(i), c = i - r * ax.n_out;
        const SrcT *s = src + r * sstride + __ldg(ax.xmin + c);
        const float *w = ax.w + c * ax.taps;
        float acc = 0.f;
        for (int j = 0; j < ax.taps; ++j) acc = fmaf(__ldg(w + j), (float)s[j], acc);
        dst[r * dstride + c] = acc;
    }
}

// dst[r][c] = sum_j w[r][j] * src[xmin[r] + j][c]      (pass along the strided axis)
template <typename SrcT>
__device__ __forceinline__ void resample_h(const SrcT *src, int sstride, float *dst, int dstride, int cols,
                                           const AxisRef &ax, int tid, int nt) {
    const int total = ax.n_out * cols;
    const FastDiv fd(cols);
    for (int i = tid; i < total; i += nt) {
        const int r = fd.div(i), c = i - r * cols;
        const SrcT *s = src + __ldg(ax.xmin + r) * sstride + c;
        const float *w = ax.w + r * ax.taps;
        float acc = 0.f;
        for (int j = 0; j < ax.taps; ++j) acc = fmaf(__ldg(w + j), (float)s[j * sstride], acc);
        dst[r * dstride + c] = acc;
    }
}

// One output word (4 pixels of row y starting at column x0) of a strided-axis pass.
__device__ __forceinline__ uint32_t resample_h_word(const float *src, int sstride, const AxisRef &ax, int y, int x0) {
    const float *s = src + __ldg(ax.xmin + y) * sstride + x0;
    const float *w = ax.w + y * ax.taps;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (int j = 0; j < ax.taps; ++j) {
        const float wj = __ldg(w + j);
        const float *sj = s + j * sstride;
        a0 = fmaf(wj, sj[0], a0);
        a1 = fmaf(wj, sj[1], a1);
        a2 = fmaf(wj, sj[2], a2);
        a3 = fmaf(wj, sj[3], a3);
    }
    return quant_u8(a0) | (quant_u8(a1) << 8) | (quant_u8(a2) << 16) | (quant_u8(a3) << 24);
}

// byte mask of the columns [c0, c1) inside the 4-byte word that starts at column x0
__device__ __forceinline__ uint32_t word_mask(int x0, int c0, int c1) {
    const int lo = max(c0 - x0, 0), hi = min(c1 - x0, 4);
    if (lo >= hi) return 0u;
    const uint32_t upto_hi = hi >= 4 ? 0xFFFFFFFFu : ((1u << (8 * hi)) - 1u);
    return upto_hi & ~((1u << (8 * lo)) - 1u);
}

// 16 interleaved 3-channel pixels (48 bytes, 12 words) -> 16 luma bytes.
// Y = (lw0 c0 + lw1 c1 + lw2 c2 + 16384) >> 15 (cv2's 15-bit luma, dmc_env.py:182) evaluated as
// (2 (lw . c) + 32768) >> 16, so that Y is byte 2 of the accumulator: one PRMT gathers a pixel's three
// bytes, two IDP.2A (16-bit weights x 8-bit channels) form the sum and PRMTs pack four results.
// w01 = 2 lw0 | 2 lw1 << 16, w2 = 2 lw2.
__device__ __forceinline__ uint4 luma16(const uint32_t (&w)[12], uint32_t w01, uint32_t w2) {
    uint32_t acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int j = 3 * i, wi = j >> 2, b = j & 3;  // compile time after unrolling
        // bytes b, b+1, b+2 of the pair (w[wi], w[wi + 1]); b <= 1 stays inside one word
        const uint32_t hi = wi + 1 < 12 ? w[wi + 1] : 0u;
        const uint32_t q = __byte_perm(w[wi], hi, (uint32_t)(b | (b + 1) << 4 | (b + 2) << 8 | 0x4000));
        acc[i] = __dp2a_hi(w2, q, __dp2a_lo(w01, q, 32768u));  // w2's upper half is 0: byte 3 of q does not matter
    }
    uint32_t o[4];
#pragma unroll
    for (int m = 0; m < 4; ++m)
        o[m] = __byte_perm(__byte_perm(acc[4 * m], acc[4 * m + 1], 0x0062), __byte_perm(acc[4 * m + 2], acc[4 * m + 3], 0x0062), 0x5410);
    return make_uint4(o[0], o[1], o[2], o[3]);
}

__device__ __forceinline__ size_t align16(size_t v) { return (v + 15) & ~size_t(15); }

// Squeeze of one obs frame held in shared memory (u8) into the peripheral cache slot:
// Resize(peripheral_res) = W pass then H pass (fov_env.py:367).  Needs t1[S_h * p_w] floats.
__device__ void squeeze_to_cache(const DevPlan &p, const uint8_t *s_frame, float *s_t1, float *dst_global, int tid, int nt) {
    resample_w<uint8_t>(s_frame, p.S_w, s_t1, p.p_w, p.S_h, p.sq_w, tid, nt);
    __syncthreads();
    resample_h<float>(s_t1, p.p_w, dst_global, p.p_w, p.p_w, p.sq_h, tid, nt);
}

// ------------------------------------------------------------------------ ingest: Atari
// AtariEnv._get_state + the frame logic of _step/_reset (atari_env.py:73-75, 80-82, 91,
// 111-114, 121-133): gray -> cv2.resize(INTER_LINEAR) for frame A and frame B separately,
// max of the two resized frames, push.  Shared memory holds, per frame, only the two source
// rows every output row samples (raw rows the resize never reads are not fetched).
template <int CH>
__global__ void __launch_bounds__(kThreads) k_ingest_atari(const __grid_constant__ DevPlan p,
                                                           const uint8_t *__restrict__ fa,
                                                           const uint8_t *__restrict__ fb,
                                                           const uint8_t *__restrict__ flags, uint8_t *__restrict__ ring,
                                                           int32_t *__restrict__ head, float *__restrict__ pcache) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int n = blockIdx.x, tid = threadIdx.x;
    const int fl = flags[n];
    if (fl & AGYM_FLAG_IDLE) return;
    const int slot = (head[n] + 1) % p.K;

    int32_t *t_xs0 = reinterpret_cast<int32_t *>(smem);
    int32_t *t_xs1 = t_xs0 + p.S_w, *t_xcf = t_xs1 + p.S_w;
    int32_t *t_ys0 = t_xcf + p.S_w, *t_ys1 = t_ys0 + p.S_h, *t_ycf = t_ys1 + p.S_h;
    uint8_t *s_gray = smem + align16(sizeof(int32_t) * 3 * (p.S_w + p.S_h));
    const int rows2 = 2 * p.S_h;                       // staged rows per frame
    const size_t gray_bytes = (size_t)rows2 * p.raw_w; // per frame
    uint8_t *s_frame = s_gray + align16(2 * gray_bytes);
    float *s_t1 = reinterpret_cast<float *>(s_frame + align16(p.plane));

    for (int i = tid; i < p.S_w; i += kThreads) {
        t_xs0[i] = p.cx_s0[i]; t_xs1[i] = p.cx_s1[i]; t_xcf[i] = p.cx_coef[i];
    }
    for (int i = tid; i < p.S_h; i += kThreads) {
        t_ys0[i] = p.cy_s0[i]; t_ys1[i] = p.cy_s1[i]; t_ycf[i] = p.cy_coef[i];
    }
    __syncthreads();  // also orders every thread's read of head[n] before the update below
    if (tid == 0) head[n] = slot;

    // ---- stage: gray rows of the valid frames -> shared memory
    const int vpr = p.raw_w / 16;  // 16-pixel groups per row
    const FastDiv fd_vpr(vpr);
    const size_t frame_bytes = (size_t)p.raw_h * p.raw_w * CH;
#pragma unroll 1
    for (int fr = 0; fr < 2; ++fr) {
        if (!(fl & (1 << fr))) continue;
        const uint8_t *src = (fr ? fb : fa) + frame_bytes * n;
        uint8_t *dst = s_gray + gray_bytes * fr;
#pragma unroll 2
        for (int t = tid; t < rows2 * vpr; t += kThreads) {
            const int sr = fd_vpr.div(t), g = t - sr * vpr;
            const int srow = (sr & 1) ? t_ys1[sr >> 1] : t_ys0[sr >> 1];
            uint4 o;
            if (CH == 1) {
                o = ld_stream128(src + (size_t)srow * p.raw_w + 16 * g);
            } else {
                const uint8_t *q = src + ((size_t)srow * p.raw_w + 16 * g) * 3;
                uint32_t w[12];
                *reinterpret_cast<uint4 *>(w) = ld_stream128(q);
                *reinterpret_cast<uint4 *>(w + 4) = ld_stream128(q + 16);
                *reinterpret_cast<uint4 *>(w + 8) = ld_stream128(q + 32);
                o = luma16(w, (2u * p.lw0) | ((2u * p.lw1) << 16), 2u * p.lw2);
            }
            *reinterpret_cast<uint4 *>(dst + (size_t)sr * p.raw_w + 16 * g) = o;
        }
    }
    // hard reset: the other K-1 slots (and their cache entries) become zero frames
    if (fl & AGYM_FLAG_HARD_RESET) {
        const int words = p.plane / 4;
        for (int k = 0; k < p.K; ++k) {
            if (k == slot) continue;
            uint32_t *z = reinterpret_cast<uint32_t *>(ring + ((size_t)n * p.K + k) * p.plane);
            for (int i = tid; i < words; i += kThreads) z[i] = 0u;
            if (pcache) {
                float *zc = pcache + ((size_t)n * p.K + k) * p.p_h * p.p_w;
                for (int i = tid; i < p.p_h * p.p_w; i += kThreads) zc[i] = 0.f;
            }
        }
    }
    __syncthreads();

    // ---- resize both frames (11-bit fixed point, horizontal then vertical), max, store
    const int wpr = p.S_w / 4;  // output words per row
    const FastDiv fd_wpr(wpr);
    uint32_t *out_words = reinterpret_cast<uint32_t *>(ring + ((size_t)n * p.K + slot) * p.plane);
    uint32_t *frame_words = reinterpret_cast<uint32_t *>(s_frame);
    for (int t = tid; t < p.plane / 4; t += kThreads) {
        const int y = fd_wpr.div(t), q = t - y * wpr;
        const int ycf = t_ycf[y];
        const int b0 = ycf & 0xffff, b1 = ycf >> 16;
        uint32_t word = 0u;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int x = 4 * q + i;
            const int s0 = t_xs0[x], s1 = t_xs1[x], xcf = t_xcf[x];
            const int c0 = xcf & 0xffff, c1 = xcf >> 16;
            int m = 0;
#pragma unroll
            for (int fr = 0; fr < 2; ++fr) {
                if (!(fl & (1 << fr))) continue;
                const uint8_t *r0 = s_gray + gray_bytes * fr + (size_t)(2 * y) * p.raw_w;
                const uint8_t *r1 = r0 + p.raw_w;
                const int h0 = r0[s0] * c0 + r0[s1] * c1;
                const int h1 = r1[s0] * c0 + r1[s1] * c1;
                const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
                m = max(m, v);
            }
            word |= (uint32_t)min(m, 255) << (8 * i);
        }
        out_words[t] = word;
        if (pcache) frame_words[t] = word;
    }
    if (pcache) {  // uniform branch
        __syncthreads();
        squeeze_to_cache(p, s_frame, s_t1, pcache + ((size_t)n * p.K + slot) * p.p_h * p.p_w, tid, kThreads);
    }
}

// Squeeze along W straight from a u8 frame in shared memory (fov_env.py:367, W pass): every
// output column reads one 16-byte window, realigned with funnel shifts so that the byte ->
// float conversions use compile-time byte selectors (I2F.U8 Rx.Bn).
__device__ __forceinline__ void squeeze_w_fast(const DevPlan &p, const uint8_t *s_frame, float *s_t1, int tid, int nt) {
    const int total = p.S_h * p.p_w;
    const FastDiv fd(p.p_w);
    const int nq = p.sqw_taps4 >> 2;
    for (int t = tid; t < total; t += nt) {
        const int y = fd.div(t), i = t - y * p.p_w;
        const int2 o = __ldg(p.sqw_ofs + i);
        const uint32_t *src = reinterpret_cast<const uint32_t *>(s_frame + y * p.S_w + o.x);
        const uint32_t w0 = src[0], w1 = src[1], w2 = src[2], w3 = src[3];
        const uint32_t a[4] = {__funnelshift_r(w0, w1, o.y), __funnelshift_r(w1, w2, o.y), __funnelshift_r(w2, w3, o.y),
                               w3 >> o.y};
        const float4 *wt = reinterpret_cast<const float4 *>(p.sqw_w + i * p.sqw_taps4);
        float acc = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (q < nq) {
                const float4 w = __ldg(wt + q);
                const uint32_t v = a[q];
                acc = fmaf(w.x, (float)(v & 0xffu), acc);
                acc = fmaf(w.y, (float)((v >> 8) & 0xffu), acc);
                acc = fmaf(w.z, (float)((v >> 16) & 0xffu), acc);
                acc = fmaf(w.w, (float)(v >> 24), acc);
            }
        }
        s_t1[t] = acc;
    }
}

// Fast ingest (geometry checked at plan creation): the horizontal pass as IDP.2A on byte pairs
// picked with PRMT from two aligned words, the vertical pass as two IMAD.HI; one thread owns
// two adjacent output columns and walks a segment of output rows.
template <int CH>
__global__ void __launch_bounds__(kThreads) k_ingest_atari_fast(const __grid_constant__ DevPlan p,
                                                                const uint8_t *__restrict__ fa,
                                                                const uint8_t *__restrict__ fb,
                                                                const uint8_t *__restrict__ flags,
                                                                uint8_t *__restrict__ ring, int32_t *__restrict__ head,
                                                                float *__restrict__ pcache) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int n = blockIdx.x, tid = threadIdx.x;
    const int fl = flags[n];
    if (fl & AGYM_FLAG_IDLE) return;
    const int slot = (head[n] + 1) % p.K;
    const int rows2 = 2 * p.S_h;
    const size_t gray_bytes = (size_t)rows2 * p.raw_w;
    // one frame's sampled rows at a time: half the shared memory, twice the CTAs per SM to hide the
    // latency of the global loads (this kernel has no prefetch pipeline)
    uint8_t *s_gray = smem;
    uint8_t *s_frame = s_gray + align16(gray_bytes + 16);
    float *s_t1 = reinterpret_cast<float *>(s_frame + align16(p.plane + 16));
    __syncthreads();  // every thread has read head[n]
    if (tid == 0) head[n] = slot;

    if (fl & AGYM_FLAG_HARD_RESET) {
        const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
        for (int k = 0; k < p.K; ++k) {
            if (k == slot) continue;
            uint4 *z = reinterpret_cast<uint4 *>(ring + ((size_t)n * p.K + k) * p.plane);
            for (int i = tid; i < p.plane / 16; i += kThreads) z[i] = z4;
            if (pcache) {
                float *zc = pcache + ((size_t)n * p.K + k) * p.p_h * p.p_w;
                for (int i = tid; i < p.p_h * p.p_w; i += kThreads) zc[i] = 0.f;
            }
        }
    }
    if (!(fl & 3))  // no frame at all (game over before t == 2): a zero frame (atari_env.py:121,132)
        for (int i = tid; i < p.plane / 16; i += kThreads) reinterpret_cast<uint4 *>(s_frame)[i] = make_uint4(0u, 0u, 0u, 0u);

    const int vpr = p.raw_w / 16;
    const FastDiv fd_vpr(vpr);
    const size_t frame_bytes = (size_t)p.raw_h * p.raw_w * CH;
    const int pairs = p.S_w >> 1, segs = kThreads / pairs;
    const int g = tid / pairs, pi = tid - g * pairs;
    const int4 px = g < segs ? __ldg(p.cx_pair + pi) : make_int4(0, 0, 0, 0);  // {aligned byte offset, PRMT selector, coef(x0), coef(x0+1)}
    const int rows_per = (p.S_h + segs - 1) / segs;
    const int y_begin = g * rows_per, y_end = g < segs ? min(p.S_h, (g + 1) * rows_per) : y_begin;
    bool first = true;
#pragma unroll 1
    for (int fr = 0; fr < 2; ++fr) {
        if (!(fl & (1 << fr))) continue;
        const uint8_t *src = (fr ? fb : fa) + frame_bytes * n;
#pragma unroll 4
        for (int t = tid; t < rows2 * vpr; t += kThreads) {
            const int sr = fd_vpr.div(t), gg = t - sr * vpr;
            const int srow = __ldg(((sr & 1) ? p.cy_s1 : p.cy_s0) + (sr >> 1));
            uint4 o;
            if (CH == 1) {
                o = ld_stream128(src + (size_t)srow * p.raw_w + 16 * gg);
            } else {
                const uint8_t *q = src + ((size_t)srow * p.raw_w + 16 * gg) * 3;
                uint32_t w[12];
                *reinterpret_cast<uint4 *>(w) = ld_stream128(q);
                *reinterpret_cast<uint4 *>(w + 4) = ld_stream128(q + 16);
                *reinterpret_cast<uint4 *>(w + 8) = ld_stream128(q + 32);
                o = luma16(w, (2u * p.lw0) | ((2u * p.lw1) << 16), 2u * p.lw2);
            }
            *reinterpret_cast<uint4 *>(s_gray + (size_t)sr * p.raw_w + 16 * gg) = o;
        }
        __syncthreads();
        for (int y = y_begin; y < y_end; ++y) {
            const int2 bs = __ldg(p.cy_bs + y);  // {b0 << 16, b1 << 16}
            const uint8_t *r0 = s_gray + (size_t)(2 * y) * p.raw_w + px.x;
            const uint32_t a0 = *reinterpret_cast<const uint32_t *>(r0);
            const uint32_t a1 = *reinterpret_cast<const uint32_t *>(r0 + 4);
            const uint32_t b0 = *reinterpret_cast<const uint32_t *>(r0 + p.raw_w);
            const uint32_t b1 = *reinterpret_cast<const uint32_t *>(r0 + p.raw_w + 4);
            const uint32_t qa = __byte_perm(a0, a1, (uint32_t)px.y), qb = __byte_perm(b0, b1, (uint32_t)px.y);
            const uint32_t h00 = __dp2a_lo((uint32_t)px.z, qa, 0u), h01 = __dp2a_hi((uint32_t)px.w, qa, 0u);
            const uint32_t h10 = __dp2a_lo((uint32_t)px.z, qb, 0u), h11 = __dp2a_hi((uint32_t)px.w, qb, 0u);
            // never above 255: b0 + b1 = 2048 and h >> 4 <= 32640
            uint32_t v0 = (__umulhi((uint32_t)bs.x, h00 >> 4) + __umulhi((uint32_t)bs.y, h10 >> 4) + 2u) >> 2;
            uint32_t v1 = (__umulhi((uint32_t)bs.x, h01 >> 4) + __umulhi((uint32_t)bs.y, h11 >> 4) + 2u) >> 2;
            uint16_t *o = reinterpret_cast<uint16_t *>(s_frame + y * p.S_w + 2 * pi);
            if (!first) {  // max with the other frame's resized pixel (atari_env.py:132)
                const uint32_t prev = *o;
                v0 = max(v0, prev & 0xffu);
                v1 = max(v1, prev >> 8);
            }
            *o = (uint16_t)(v0 | (v1 << 8));
        }
        first = false;
        __syncthreads();
    }
    uint4 *out4 = reinterpret_cast<uint4 *>(ring + ((size_t)n * p.K + slot) * p.plane);
    for (int i = tid; i < p.plane / 16; i += kThreads) out4[i] = reinterpret_cast<const uint4 *>(s_frame)[i];
    if (pcache) {  // uniform
        float *dst = pcache + ((size_t)n * p.K + slot) * p.p_h * p.p_w;
        if (p.fast_squeeze) squeeze_w_fast(p, s_frame, s_t1, tid, kThreads);
        else resample_w<uint8_t>(s_frame, p.S_w, s_t1, p.p_w, p.S_h, p.sq_w, tid, kThreads);
        __syncthreads();
        resample_h<float>(s_t1, p.p_w, dst, p.p_w, p.p_w, p.sq_h, tid, kThreads);
    }
}

// ---- TMA (cp.async.bulk) + mbarrier helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
constexpr uint32_t kBackoffNs = 2000;  // suspend-time hint of the producer's try_wait
template <bool BACKOFF = false>
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        if (BACKOFF) {
            // a lone producer lane must not burn issue slots while it waits: let the hardware suspend
            // the thread on the barrier (suspend-time hint) instead of polling it
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done)
                : "r"(smem_u32(bar)), "r"(parity), "r"(kBackoffNs)
                : "memory");
        } else {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done)
                : "r"(smem_u32(bar)), "r"(parity)
                : "memory");
        }
    }
}
// 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// 4-D tiled tensor copy global -> shared through a CUtensorMap (TMA; SASS: UTMALDG), completing on an mbarrier
__device__ __forceinline__ void tensor_g2s_4d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}

// 1-D bulk copy shared -> global (TMA store; SASS: UBLKCP), tracked by bulk async-groups
__device__ __forceinline__ void bulk_s2g(void *gmem_dst, const void *smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the issuing thread's bulk stores, all but the newest N groups, have finished READING shared memory
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
// generic-proxy writes to shared memory become visible to the async proxy (TMA)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// barrier among the consumer warps only (the producer warp never joins it)
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kThreads) : "memory"); }

// Persistent, TMA-fed ingest for gray frames.  Every CTA walks the env batch in "units" of R
// output rows; for each unit ONE cp.async.bulk per frame brings the contiguous span of source
// rows the unit samples into a ring of shared-memory stages, tracked by full/empty mbarriers
// (the canonical TMA producer/consumer pipeline): warp 8 is the producer, warps 0-7 resize.
// HBM streaming and the fixed-point arithmetic therefore overlap inside every CTA.
// RAW_W / S_W > 0 bake the strides of the standard geometry (210x160 -> 84x84) into the
// instruction immediates; 0 = take them from the plan.
constexpr int kIngestThreads = kThreads + 32;

// TM: the standard 2.5x vertical scale samples raw rows {5m, 5m+1} (even output rows) and {5m+3, 5m+4} (odd ones)
// and never row 5m+2.  The frames are then viewed as a 4-D tensor [env][period of 5 rows][row in period][row bytes]
// and a unit's rows arrive as TWO tiled tensor copies per frame (boxes of 2 rows x R/2 periods at row 0 and at
// row 3 of the period): the unsampled fifth of every frame never leaves HBM, with as few copies as before.
template <int RAW_W, int S_W, int CH, bool TM, int NS>  // RAW_W: BYTES per raw row (pixels * CH) when baked in; NS: stages
__global__ void __launch_bounds__(kIngestThreads, 3) k_ingest_atari_tma(const __grid_constant__ DevPlan p,
                                                                        const uint8_t *__restrict__ fa,
                                                                        const uint8_t *__restrict__ fb,
                                                                        const uint8_t *__restrict__ flags,
                                                                        uint8_t *__restrict__ ring,
                                                                        int32_t *__restrict__ head,
                                                                        float *__restrict__ pcache, int units,
                                                                        int span_rows,
                                                                        const __grid_constant__ CUtensorMap tma,
                                                                        const __grid_constant__ CUtensorMap tmb) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint8_t *smem = smem_raw + (TM ? ((128u - (smem_u32(smem_raw) & 127u)) & 127u) : 0u);   // tensor copies land on 128-byte lines
    __shared__ __align__(8) uint64_t full[NS], empty[NS];
    constexpr int kEnvWin = 64;
    __shared__ int s_envfl[kEnvWin], s_envhd[kEnvWin];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = p.N, K = p.K;
    const int raw_w = RAW_W ? RAW_W : p.raw_w * CH;               // bytes per raw row
    const int S_w = S_W ? S_W : p.S_w;
    const int R = p.S_h / units;                          // output rows per unit
    const int box_bytes = R * raw_w;                      // TM: the R/2 x 2 rows of one tensor copy ...
    const int block_bytes = (box_bytes + 127) & ~127;     // ... which must land on a 128-byte line
    const int frame_stride = TM ? 2 * block_bytes + 128 : span_rows * raw_w + 16;  // one frame's staged rows of a unit (+ pad)
    const int stage_bytes = (2 * frame_stride + 15) & ~15;
    uint8_t *stages = smem;
    uint8_t *s_frame = stages + NS * stage_bytes;
    float *s_t1 = reinterpret_cast<float *>(s_frame + align16(p.plane + 16));
    int4 *s_row = reinterpret_cast<int4 *>(s_t1 + (pcache ? p.S_h * p.p_w : 0));  // [S_h] {ofs0, ofs1, b0<<16, b1<<16}
    int2 *s_span = reinterpret_cast<int2 *>(s_row + p.S_h);                       // [units] {first row, bytes}
    float *s_sqw = reinterpret_cast<float *>(s_span + ((units + 1) & ~1));        // [p_w][taps4], 16-byte aligned
    uint32_t *s_sqq = reinterpret_cast<uint32_t *>(s_sqw + (pcache ? p.p_w * 16 : 0));  // [p_w][8] fixed-point W weights
    float *s_sqh = reinterpret_cast<float *>(s_sqq + (pcache ? p.p_w * 8 : 0));         // [p_h][taps] H-pass weights
    int32_t *s_sqx = reinterpret_cast<int32_t *>(s_sqh + (pcache ? p.p_h * p.sq_h.taps : 0));  // [p_h] first source row
    const size_t frame_bytes = (size_t)p.raw_h * raw_w;

    if (tid == 0) {
        for (int i = 0; i < NS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], kThreads / 32); }
        mbar_fence_init();
    }
    for (int y = tid; y < p.S_h; y += kIngestThreads) {
        const int2 bs = __ldg(p.cy_bs + y);
        if (TM) {   // even rows of the unit in the first block, odd rows in the second, two staged rows each
            const int yy = y - (y / R) * R;
            const int o = ((yy & 1) ? block_bytes : 0) + (yy >> 1) * 2 * raw_w;
            s_row[y] = make_int4(o, o + raw_w, bs.x, bs.y);
        } else {
            const int lo = __ldg(p.cy_s0 + (y / R) * R);     // first source row of this row's unit
            s_row[y] = make_int4((__ldg(p.cy_s0 + y) - lo) * raw_w, (__ldg(p.cy_s1 + y) - lo) * raw_w, bs.x, bs.y);
        }
    }
    for (int u = tid; u < units; u += kIngestThreads) {
        const int lo = __ldg(p.cy_s0 + u * R), hi = __ldg(p.cy_s1 + u * R + R - 1);
        s_span[u] = make_int2(lo, (hi - lo + 1) * raw_w);
    }
    if (pcache && p.fast_squeeze)
        for (int i = tid; i < p.p_w * p.sqw_taps4; i += kIngestThreads) s_sqw[i] = __ldg(p.sqw_w + i);
    if (pcache) {
        for (int i = tid; i < p.p_w * 8; i += kIngestThreads) s_sqq[i] = p.squeeze_q ? __ldg(p.sqw_q + i) : 0u;
        for (int i = tid; i < p.p_h * p.sq_h.taps; i += kIngestThreads) s_sqh[i] = __ldg(p.sq_h.w + i);
        for (int i = tid; i < p.p_h; i += kIngestThreads) s_sqx[i] = __ldg(p.sq_h.xmin + i);
    }
    __syncthreads();

    // unit `it` of this CTA: env = blockIdx.x + (it / units) * gridDim.x, part = it % units
    const int my_envs = (N - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int my_units = my_envs * units;

    if (warp == kThreads / 32) {
        // ------------------------------------------------------------------ producer warp
        if (lane == 0) {
            int n = blockIdx.x, part = 0, st = 0, ph = 1;  // ph: parity of the empty-barrier phase to wait for
            for (int it = 0; it < my_units; ++it) {
                if (it >= NS) mbar_wait<true>(&empty[st], ph);
                const int fl = flags[n];
                const int2 span = s_span[part];
                const int nvalid = (fl & AGYM_FLAG_IDLE) ? 0 : __popc(fl & 3);
                uint8_t *dst = stages + st * stage_bytes;
                if (TM) {
                    mbar_expect_tx(&full[st], (uint32_t)(nvalid * 2 * box_bytes));
                    const int m0 = (R >> 1) * part;
                    if (nvalid && (fl & AGYM_FLAG_FRAME_A)) {
                        tensor_g2s_4d(dst, &tma, &full[st], 0, 0, m0, n);
                        tensor_g2s_4d(dst + block_bytes, &tma, &full[st], 0, 3, m0, n);
                    }
                    if (nvalid && (fl & AGYM_FLAG_FRAME_B)) {
                        tensor_g2s_4d(dst + frame_stride, &tmb, &full[st], 0, 0, m0, n);
                        tensor_g2s_4d(dst + frame_stride + block_bytes, &tmb, &full[st], 0, 3, m0, n);
                    }
                } else {
                    mbar_expect_tx(&full[st], (uint32_t)(nvalid * span.y));
                    if (nvalid) {
                        const size_t off = frame_bytes * n + (size_t)span.x * raw_w;
                        if (fl & AGYM_FLAG_FRAME_A) bulk_g2s(dst, fa + off, span.y, &full[st]);
                        if (fl & AGYM_FLAG_FRAME_B) bulk_g2s(dst + frame_stride, fb + off, span.y, &full[st]);
                    }
                }
                if (++part == units) { part = 0; n += gridDim.x; }
                if (++st == NS) { st = 0; ph ^= 1; }
            }
        }
        return;
    }

    // ---------------------------------------------------------------------- consumer warps
    const int pairs = S_w >> 1, segs = kThreads / pairs;
    const int g = tid / pairs, pi = tid - g * pairs;
    const bool worker = g < segs;
    const int4 px = worker ? __ldg(p.cx_pair + pi) : make_int4(0, 0, 0, 0);
    // RGB: the four source pixels of this column pair lie among s0 .. s0 + 3 (checked when the plan was made):
    // 12 bytes starting at byte 3 * s0 of the row
    const int s0px = px.x + (px.y & 7);
    const int tap_off = CH == 1 ? px.x : ((3 * s0px) & ~3);
    const uint32_t rgb_sh = (uint32_t)((3 * s0px) & 3) * 8u;
    uint32_t rgb_sel = 0u;
    if (CH == 3) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int rel = ((px.y >> (4 * k)) & 7) - (px.y & 7);  // 0..3: which of the four lumas
            rgb_sel |= (uint32_t)(rel < 2 ? rel : rel + 2) << (4 * k);
        }
    }
    const uint32_t lw_a = (2u * p.lw0) | ((2u * p.lw1) << 16), lw_b = 2u * p.lw2;
    const uint32_t lw_c = (2u * p.lw0) << 16, lw_d = (2u * p.lw1) | ((2u * p.lw2) << 16);
    // the 4 source bytes {s0(x0), s1(x0), s0(x0+1), s1(x0+1)} of one raw row, as the IDP.2A operand
    auto tap4 = [&](const uint8_t *row) -> uint32_t {
        const uint32_t *w = reinterpret_cast<const uint32_t *>(row);
        if (CH == 1) return __byte_perm(w[0], w[1], (uint32_t)px.y);
        // cv2 luma Y = (lw . c + 16384) >> 15 = byte 2 of 2 (lw . c) + 32768, for 4 consecutive RGB pixels:
        // v0 = R0 G0 B0 R1, v1 = G1 B1 R2 G2, v2 = B2 R3 G3 B3
        const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3];
        const uint32_t v0 = __funnelshift_r(w0, w1, rgb_sh), v1 = __funnelshift_r(w1, w2, rgb_sh), v2 = __funnelshift_r(w2, w3, rgb_sh);
        const uint32_t y0 = __dp2a_hi(lw_b, v0, __dp2a_lo(lw_a, v0, 32768u));
        const uint32_t y1 = __dp2a_lo(lw_d, v1, __dp2a_hi(lw_c, v0, 32768u));
        const uint32_t y2 = __dp2a_lo(lw_b, v2, __dp2a_hi(lw_a, v1, 32768u));
        const uint32_t y3 = __dp2a_hi(lw_d, v2, __dp2a_lo(lw_c, v2, 32768u));
        return __byte_perm(__byte_perm(y0, y1, 0x0062), __byte_perm(y2, y3, 0x0062), rgb_sel);
    };
    const int rows_per = (R + segs - 1) / segs;
    const int yy_begin = g * rows_per, yy_end = worker ? min(R, yy_begin + rows_per) : yy_begin;
    // squeeze along W: one thread = one output column i, rows in passes
    const int sq_rows = p.fast_squeeze ? kThreads / p.p_w : 0;
    const int sq_i = sq_rows ? tid % p.p_w : 0, sq_y0 = sq_rows ? tid / p.p_w : 0;
    const bool sq_worker = pcache && sq_rows && sq_y0 < sq_rows;
    const int nq = p.sqw_taps4 >> 2;
    const int2 sq_o = sq_worker ? __ldg(p.sqw_ofs + sq_i) : make_int2(0, 0);
    const float4 *sqw4 = reinterpret_cast<const float4 *>(s_sqw + sq_i * p.sqw_taps4);
    const uint4 *sqq4 = reinterpret_cast<const uint4 *>(s_sqq) + 2 * sq_i;  // fixed-point weights of this column

    int slot = 0, fl = 0;
    for (int it = 0, n = blockIdx.x, part = 0, st = 0, ph = 0, je = 0; it < my_units; ++it) {
        if (part == 0) {
            // flags / head of this CTA's next kEnvWin envs are fetched together into shared memory: a load per
            // env, used at once, left every warp waiting ~a microsecond of DRAM latency per env
            if ((je & (kEnvWin - 1)) == 0) {
                consumer_sync();
                if (tid < kEnvWin && (long long)n + (long long)tid * gridDim.x < N) {
                    s_envfl[tid] = flags[n + tid * gridDim.x];
                    s_envhd[tid] = head[n + tid * gridDim.x];
                }
                consumer_sync();
            }
            fl = s_envfl[je & (kEnvWin - 1)];
            slot = s_envhd[je & (kEnvWin - 1)] + 1;
            slot -= slot >= K ? K : 0;
            ++je;
        }
        const bool idle = fl & AGYM_FLAG_IDLE;
        mbar_wait(&full[st], ph);
        if (!idle && yy_begin < yy_end) {
            const uint8_t *base = stages + st * stage_bytes + tap_off;
            const int4 *rw = s_row + part * R + yy_begin;
            uint8_t *o = s_frame + (part * R + yy_begin) * S_w + 2 * pi;
            if ((fl & 3) == 3) {  // both frames (the steady state)
#pragma unroll 2
                for (int yy = yy_begin; yy < yy_end; ++yy, ++rw, o += S_w) {
                    const int4 t = *rw;
                    const uint8_t *ra = base + t.x, *rb = base + t.y;
                    // max of the two frames taken before the final (x + 2) >> 2, which is monotone; the
                    // result cannot exceed 255 (b0 + b1 = 2048, h >> 4 <= 32640), so cv2's saturate is a no-op
                    uint32_t m0 = 0u, m1 = 0u;
#pragma unroll
                    for (int fr = 0; fr < 2; ++fr) {
                        const uint32_t qa = tap4(ra + fr * frame_stride), qb = tap4(rb + fr * frame_stride);
                        const uint32_t h00 = __dp2a_lo((uint32_t)px.z, qa, 0u), h01 = __dp2a_hi((uint32_t)px.w, qa, 0u);
                        const uint32_t h10 = __dp2a_lo((uint32_t)px.z, qb, 0u), h11 = __dp2a_hi((uint32_t)px.w, qb, 0u);
                        m0 = max(m0, __umulhi((uint32_t)t.z, h00 >> 4) + __umulhi((uint32_t)t.w, h10 >> 4));
                        m1 = max(m1, __umulhi((uint32_t)t.z, h01 >> 4) + __umulhi((uint32_t)t.w, h11 >> 4));
                    }
                    *reinterpret_cast<uint16_t *>(o) = (uint16_t)(((m0 + 2u) >> 2) | (((m1 + 2u) >> 2) << 8));
                }
            } else {  // resets / early game-over: one frame or none
                for (int yy = yy_begin; yy < yy_end; ++yy, ++rw, o += S_w) {
                    const int4 t = *rw;
                    uint32_t m0 = 0u, m1 = 0u;
                    for (int fr = 0; fr < 2; ++fr) {
                        if (!(fl & (1 << fr))) continue;
                        const uint8_t *ra = base + t.x + fr * frame_stride, *rb = base + t.y + fr * frame_stride;
                        const uint32_t qa = tap4(ra), qb = tap4(rb);
                        const uint32_t h00 = __dp2a_lo((uint32_t)px.z, qa, 0u), h01 = __dp2a_hi((uint32_t)px.w, qa, 0u);
                        const uint32_t h10 = __dp2a_lo((uint32_t)px.z, qb, 0u), h11 = __dp2a_hi((uint32_t)px.w, qb, 0u);
                        m0 = max(m0, (__umulhi((uint32_t)t.z, h00 >> 4) + __umulhi((uint32_t)t.w, h10 >> 4) + 2u) >> 2);
                        m1 = max(m1, (__umulhi((uint32_t)t.z, h01 >> 4) + __umulhi((uint32_t)t.w, h11 >> 4) + 2u) >> 2);
                    }
                    *reinterpret_cast<uint16_t *>(o) = (uint16_t)(min(m0, 255u) | (min(m1, 255u) << 8));
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[st]);  // this warp is done with the stage
        if (++st == NS) { st = 0; ph ^= 1; }
        if (++part < units) continue;
        part = 0;
        const int n_done = n;
        n += gridDim.x;
        if (idle) continue;
        {
            const int n = n_done;
        consumer_sync();  // the whole 84x84 frame is in s_frame
        if (tid == 0) head[n] = slot;
        if (fl & AGYM_FLAG_HARD_RESET) {
            const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
            for (int k = 0; k < K; ++k) {
                if (k == slot) continue;
                uint4 *z = reinterpret_cast<uint4 *>(ring + ((size_t)n * K + k) * p.plane);
                for (int i = tid; i < p.plane / 16; i += kThreads) z[i] = z4;
                if (pcache) {
                    float *zc = pcache + ((size_t)n * K + k) * p.p_h * p.p_w;
                    for (int i = tid; i < p.p_h * p.p_w; i += kThreads) zc[i] = 0.f;
                }
            }
        }
        uint4 *out4 = reinterpret_cast<uint4 *>(ring + ((size_t)n * K + slot) * p.plane);
        for (int i = tid; i < p.plane / 16; i += kThreads) out4[i] = reinterpret_cast<const uint4 *>(s_frame)[i];
        if (pcache) {  // uniform
            if (sq_rows && p.squeeze_q) {
                // W pass in 16-bit fixed point: 8 IDP.2A over the aligned 16-byte window, exact integer sum
                if (sq_worker) {
                    const uint4 qa = sqq4[0], qb = sqq4[1];  // live only across this loop
                    for (int y = sq_y0; y < p.S_h; y += sq_rows) {
                        const uint32_t *src = reinterpret_cast<const uint32_t *>(s_frame + y * S_w + sq_o.x);
                        uint32_t acc = __dp2a_lo(qa.x, src[0], 0u);
                        acc = __dp2a_hi(qa.y, src[0], acc);
                        acc = __dp2a_lo(qa.z, src[1], acc);
                        acc = __dp2a_hi(qa.w, src[1], acc);
                        acc = __dp2a_lo(qb.x, src[2], acc);
                        acc = __dp2a_hi(qb.y, src[2], acc);
                        acc = __dp2a_lo(qb.z, src[3], acc);
                        acc = __dp2a_hi(qb.w, src[3], acc);
                        s_t1[y * p.p_w + sq_i] = (float)acc * (1.f / 131072.f);
                    }
                }
            } else if (sq_rows) {
                if (sq_worker) {
                    for (int y = sq_y0; y < p.S_h; y += sq_rows) {
                        const uint32_t *src = reinterpret_cast<const uint32_t *>(s_frame + y * S_w + sq_o.x);
                        const uint32_t w0 = src[0], w1 = src[1], w2 = src[2], w3 = src[3];
                        const uint32_t a[4] = {__funnelshift_r(w0, w1, sq_o.y), __funnelshift_r(w1, w2, sq_o.y),
                                               __funnelshift_r(w2, w3, sq_o.y), w3 >> sq_o.y};
                        float acc = 0.f;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            if (q < nq) {
                                const uint32_t v = a[q];
                                const float4 w = sqw4[q];
                                acc = fmaf(w.x, (float)(v & 0xffu), acc);
                                acc = fmaf(w.y, (float)((v >> 8) & 0xffu), acc);
                                acc = fmaf(w.z, (float)((v >> 16) & 0xffu), acc);
                                acc = fmaf(w.w, (float)(v >> 24), acc);
                            }
                        }
                        s_t1[y * p.p_w + sq_i] = acc;
                    }
                }
            } else {
                resample_w<uint8_t>(s_frame, S_w, s_t1, p.p_w, p.S_h, p.sq_w, tid, kThreads);
            }
            consumer_sync();
            {   // H pass: out[i][j] = sum_t wh[i][t] * t1[xmin[i] + t][j], weights from shared memory
                float *dst = pcache + ((size_t)n * K + slot) * p.p_h * p.p_w;
                const int pw = p.p_w, taps = p.sq_h.taps, total = p.p_h * pw;
                auto one = [&](int o, int i, int j) {
                    const float *w = s_sqh + i * taps;
                    const float *t = s_t1 + s_sqx[i] * pw + j;
                    float acc0 = 0.f, acc1 = 0.f;
                    int tt = 0;
                    for (; tt + 1 < taps; tt += 2) {
                        acc0 = fmaf(w[tt], t[tt * pw], acc0);
                        acc1 = fmaf(w[tt + 1], t[(tt + 1) * pw], acc1);
                    }
                    if (tt < taps) acc0 = fmaf(w[tt], t[tt * pw], acc0);
                    dst[o] = acc0 + acc1;
                };
                if ((pw & 3) == 0) {
                    // four adjacent columns per thread: one LDS.128 of t1 and one weight per tap serve 4 FFMA
                    const int nq = pw >> 2;
                    const FastDiv fd_nq(nq);
                    for (int o = tid; o < p.p_h * nq; o += kThreads) {
                        const int i = fd_nq.div(o), j = 4 * (o - i * nq);
                        const float *w = s_sqh + i * taps;
                        const float *t = s_t1 + s_sqx[i] * pw + j;
                        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                        for (int tt = 0; tt < taps; ++tt) {
                            const float4 v = *reinterpret_cast<const float4 *>(t + tt * pw);
                            const float wt = w[tt];
                            acc.x = fmaf(wt, v.x, acc.x); acc.y = fmaf(wt, v.y, acc.y);
                            acc.z = fmaf(wt, v.z, acc.z); acc.w = fmaf(wt, v.w, acc.w);
                        }
                        *reinterpret_cast<float4 *>(dst + i * pw + j) = acc;
                    }
                } else {
                    const FastDiv fd_pw(pw);
                    for (int o = tid; o < total; o += kThreads) {
                        const int i = fd_pw.div(o);
                        one(o, i, o - i * pw);
                    }
                }
            }
        }
        // with the cache, no barrier here: the H pass reads only s_t1, which the next env rewrites after its own
        // 'frame complete' barrier, and s_frame was last read before the barrier between the two passes
        if (!pcache) consumer_sync();  // s_frame is rewritten by the next env's first unit
        }
    }
}

// -------------------------------------------------------------------------- ingest: DMC
// DMCEnv._get_obs pixel/grey branch + stack logic (dmc_env.py:175-183, 193-195, 206-207,
// 228-230): 15-bit luma of the frame rendered at obs_size, pushed as is (no max-pool).
__global__ void __launch_bounds__(kThreads) k_ingest_dmc(const __grid_constant__ DevPlan p,
                                                         const uint8_t *__restrict__ f, const uint8_t *__restrict__ flags,
                                                         uint8_t *__restrict__ ring, int32_t *__restrict__ head,
                                                         float *__restrict__ pcache) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int n = blockIdx.x, tid = threadIdx.x;
    const uint8_t *src = f + (size_t)n * p.plane * 3;
    const int nvec = p.plane / 16;
    // the frame words of this thread's first two 16-pixel groups are requested before flags / head are known
    // (an idle env wastes them): one DRAM round trip per CTA instead of two
    uint32_t w0[12], w1[12];
    const int t0 = tid, t1 = tid + kThreads;
    if (t0 < nvec) {
        const uint8_t *q = src + (size_t)t0 * 48;
        *reinterpret_cast<uint4 *>(w0) = ld_stream128(q);
        *reinterpret_cast<uint4 *>(w0 + 4) = ld_stream128(q + 16);
        *reinterpret_cast<uint4 *>(w0 + 8) = ld_stream128(q + 32);
    }
    if (t1 < nvec) {
        const uint8_t *q = src + (size_t)t1 * 48;
        *reinterpret_cast<uint4 *>(w1) = ld_stream128(q);
        *reinterpret_cast<uint4 *>(w1 + 4) = ld_stream128(q + 16);
        *reinterpret_cast<uint4 *>(w1 + 8) = ld_stream128(q + 32);
    }
    const int fl = flags[n];
    const int hd = head[n];
    if (fl & AGYM_FLAG_IDLE) return;
    const int slot = (hd + 1) % p.K;
    uint8_t *s_frame = smem;
    float *s_t1 = reinterpret_cast<float *>(smem + align16(p.plane));
    __syncthreads();
    if (tid == 0) head[n] = slot;

    uint4 *dst = reinterpret_cast<uint4 *>(ring + ((size_t)n * p.K + slot) * p.plane);
    const uint32_t lw01 = (2u * p.lw0) | ((2u * p.lw1) << 16), lw2 = 2u * p.lw2;
    if (t0 < nvec) {
        const uint4 o = luma16(w0, lw01, lw2);
        dst[t0] = o;
        if (pcache) reinterpret_cast<uint4 *>(s_frame)[t0] = o;
    }
    if (t1 < nvec) {
        const uint4 o = luma16(w1, lw01, lw2);
        dst[t1] = o;
        if (pcache) reinterpret_cast<uint4 *>(s_frame)[t1] = o;
    }
    for (int t = tid + 2 * kThreads; t < nvec; t += kThreads) {  // larger observations
        uint32_t w[12];
        const uint8_t *q = src + (size_t)t * 48;
        *reinterpret_cast<uint4 *>(w) = ld_stream128(q);
        *reinterpret_cast<uint4 *>(w + 4) = ld_stream128(q + 16);
        *reinterpret_cast<uint4 *>(w + 8) = ld_stream128(q + 32);
        const uint4 o = luma16(w, lw01, lw2);
        dst[t] = o;
        if (pcache) reinterpret_cast<uint4 *>(s_frame)[t] = o;
    }
    if (fl & AGYM_FLAG_HARD_RESET) {
        for (int k = 0; k < p.K; ++k) {
            if (k == slot) continue;
            uint4 *z = reinterpret_cast<uint4 *>(ring + ((size_t)n * p.K + k) * p.plane);
            for (int i = tid; i < p.plane / 16; i += kThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
            if (pcache) {
                float *zc = pcache + ((size_t)n * p.K + k) * p.p_h * p.p_w;
                for (int i = tid; i < p.p_h * p.p_w; i += kThreads) zc[i] = 0.f;
            }
        }
    }
    if (pcache) {
        __syncthreads();
        squeeze_to_cache(p, s_frame, s_t1, pcache + ((size_t)n * p.K + slot) * p.p_h * p.p_w, tid, kThreads);
    }
}

// ------------------------------------------------------------------------------- stack
// np.stack(state_buffer) (atari_env.py:143, dmc_env.py:230): oldest -> newest.
__global__ void __launch_bounds__(kThreads) k_stack(const __grid_constant__ DevPlan p, const uint8_t *__restrict__ ring,
                                                    const int32_t *__restrict__ head, uint8_t *__restrict__ out) {
    const int n = blockIdx.x;
    const int h = head[n];
    const int vpp = p.plane / 16;
    for (int k = 0; k < p.K; ++k) {
        const uint4 *src = reinterpret_cast<const uint4 *>(ring + ((size_t)n * p.K + (h + 1 + k) % p.K) * p.plane);
        uint4 *dst = reinterpret_cast<uint4 *>(out + ((size_t)n * p.K + k) * p.plane);
        for (int i = threadIdx.x; i < vpp; i += kThreads) dst[i] = __ldg(src + i);
    }
}

// ------------------------------------------------------------------------ observe: fixed
// FixedFovealEnv._fov_step + _get_fov_state (fov_env.py:166-203).
template <int VARIANT>
__global__ void __launch_bounds__(kThreads) k_observe_fixed(const __grid_constant__ DevPlan p,
                                                            const uint8_t *__restrict__ ring,
                                                            const int32_t *__restrict__ head,
                                                            const double *__restrict__ action,
                                                            const uint8_t *__restrict__ ctrl, int32_t *__restrict__ loc,
                                                            uint8_t *__restrict__ out) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ int s_loc[2];
    const int n = blockIdx.x, tid = threadIdx.x;
    if (tid == 0) {
        int r, c;
        update_loc_fixed(p, n, action, ctrl, loc, r, c);
        s_loc[0] = r; s_loc[1] = c;
    }
    __syncthreads();
    const int r0 = s_loc[0], c0 = s_loc[1];
    const int h = head[n];
    const uint8_t *env_ring = ring + (size_t)n * p.K * p.plane;

    if (VARIANT == AGYM_OUT_CROP) {
        // (K, f_h, f_w) packed; bytes gathered from the (L2-resident) ring rows
        const int per_k = p.f_h * p.f_w, total = p.K * per_k;
        const FastDiv fd_k(per_k), fd_w(p.f_w);
        uint8_t *dst = out + (size_t)n * total;
        if ((total & 3) == 0) {
            for (int t = tid; t < total / 4; t += kThreads) {
                uint32_t word = 0u;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int b = 4 * t + i;
                    const int k = fd_k.div(b), rem = b - k * per_k;
                    const int y = fd_w.div(rem), x = rem - y * p.f_w;
                    const uint8_t *src = env_ring + (size_t)((h + 1 + k) % p.K) * p.plane;
                    word |= (uint32_t)__ldg(src + (r0 + y) * p.S_w + c0 + x) << (8 * i);
                }
                reinterpret_cast<uint32_t *>(dst)[t] = word;
            }
        } else {
            for (int b = tid; b < total; b += kThreads) {
                const int k = fd_k.div(b), rem = b - k * per_k;
                const int y = fd_w.div(rem), x = rem - y * p.f_w;
                const uint8_t *src = env_ring + (size_t)((h + 1 + k) % p.K) * p.plane;
                dst[b] = __ldg(src + (r0 + y) * p.S_w + c0 + x);
            }
        }
    } else if (VARIANT == AGYM_OUT_MASK) {
        // (K, S_h, S_w): the ring word where it lies inside the fovea, zero elsewhere
        const int wpr = p.S_w / 4, wpp = p.plane / 4;
        const FastDiv fd_wpr(wpr);
        for (int k = 0; k < p.K; ++k) {
            const uint32_t *src = reinterpret_cast<const uint32_t *>(env_ring + (size_t)((h + 1 + k) % p.K) * p.plane);
            uint32_t *dst = reinterpret_cast<uint32_t *>(out + ((size_t)n * p.K + k) * p.plane);
            for (int t = tid; t < wpp; t += kThreads) {
                const int y = fd_wpr.div(t), x0 = 4 * (t - y * wpr);
                uint32_t m = 0u;
                if (y >= r0 && y < r0 + p.f_h) m = word_mask(x0, c0, c0 + p.f_w);
                dst[t] = m ? (__ldg(src + t) & m) : 0u;
            }
        }
    } else {
        // resize_to_full: Resize(obs_size) of the crop, W pass then H pass (fov_env.py:120,182)
        float *s_crop = reinterpret_cast<float *>(smem);
        float *s_t = s_crop + p.f_h * p.f_w;
        const int wpr = p.S_w / 4, wpp = p.plane / 4;
        const FastDiv fd_fw(p.f_w), fd_wpr(wpr);
        for (int k = 0; k < p.K; ++k) {
            const uint8_t *src = env_ring + (size_t)((h + 1 + k) % p.K) * p.plane;
            for (int i = tid; i < p.f_h * p.f_w; i += kThreads) {
                const int y = fd_fw.div(i), x = i - y * p.f_w;
                s_crop[i] = (float)__ldg(src + (r0 + y) * p.S_w + c0 + x);
            }
            __syncthreads();
            resample_w<float>(s_crop, p.f_w, s_t, p.S_w, p.f_h, p.full_w, tid, kThreads);
            __syncthreads();
            uint32_t *dst = reinterpret_cast<uint32_t *>(out + ((size_t)n * p.K + k) * p.plane);
            for (int t = tid; t < wpp; t += kThreads) {
                const int y = fd_wpr.div(t), x0 = 4 * (t - y * wpr);
                dst[t] = resample_h_word(s_t, p.S_w, p.full_h, y, x0);
            }
            __syncthreads();
        }
    }
}

// Crop variant, one WARP per env (8 envs per CTA): a (K, f_h, f_w) observation is only a few KB, so
// a CTA per env spends its life in the latency chain loc update -> loads -> stores; here every lane
// owns ~total/128 output words whose byte gathers are all independent and in flight together.
__global__ void __launch_bounds__(kThreads) k_observe_fixed_crop_warp(const __grid_constant__ DevPlan p,
                                                                      const uint8_t *__restrict__ ring,
                                                                      const int32_t *__restrict__ head,
                                                                      const double *__restrict__ action,
                                                                      const uint8_t *__restrict__ ctrl,
                                                                      int32_t *__restrict__ loc, uint8_t *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int n = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
    if (n >= p.N) return;
    int r0 = 0, c0 = 0;
    if (lane == 0) update_loc_fixed(p, n, action, ctrl, loc, r0, c0);
    r0 = __shfl_sync(0xffffffffu, r0, 0);
    c0 = __shfl_sync(0xffffffffu, c0, 0);
    const int h = head[n];
    const int per_k = p.f_h * p.f_w, words = (p.K * per_k) >> 2;
    const FastDiv fd_k(per_k), fd_w(p.f_w);
    const uint8_t *env_ring = ring + (size_t)n * p.K * p.plane + (size_t)r0 * p.S_w + c0;
    uint32_t *dst = reinterpret_cast<uint32_t *>(out + (size_t)n * p.K * per_k);
#pragma unroll 4
    for (int t = lane; t < words; t += 32) {
        const int b = 4 * t;
        int k = fd_k.div(b), rem = b - k * per_k;
        int y = fd_w.div(rem), x = rem - y * p.f_w;
        int slot = h + 1 + k;
        slot -= slot >= p.K ? p.K : 0;
        const uint8_t *src = env_ring + (size_t)slot * p.plane + y * p.S_w;
        uint32_t word = 0u;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            word |= (uint32_t)__ldg(src + x) << (8 * i);
            if (++x == p.f_w) {  // next fovea row, possibly next frame
                x = 0;
                src += p.S_w;
                if (++y == p.f_h) {
                    y = 0;
                    ++k;
                    slot = slot + 1 == p.K ? 0 : slot + 1;
                    src = env_ring + (size_t)slot * p.plane;
                }
            }
        }
        dst[t] = word;
    }
}

// ------------------------------------------------------------------- observe: peripheral
// FixedFovealPeripheralEnv._get_fov_state (fov_env.py:375-388):
//   out = Resize(obs)(Resize(peripheral_res)(full)); out[fovea] = full[fovea].
// CACHED: the squeeze of every ring slot was stored at ingest time (it does not depend on
// fov_loc), so only the expand + paste remain per step.
template <bool CACHED>
__global__ void __launch_bounds__(kThreads) k_observe_peripheral(const __grid_constant__ DevPlan p,
                                                                 const uint8_t *__restrict__ ring,
                                                                 const int32_t *__restrict__ head,
                                                                 const float *__restrict__ pcache,
                                                                 const double *__restrict__ action,
                                                                 const uint8_t *__restrict__ ctrl,
                                                                 int32_t *__restrict__ loc, uint8_t *__restrict__ out) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ int s_loc[2];
    const int n = blockIdx.x, tid = threadIdx.x;
    if (tid == 0) {
        int r, c;
        update_loc_fixed(p, n, action, ctrl, loc, r, c);
        s_loc[0] = r; s_loc[1] = c;
    }
    __syncthreads();
    const int r0 = s_loc[0], c0 = s_loc[1];
    const int h = head[n];

    // shared: frame u8 [plane] | t1 f32 [S_h][p_w] | sq f32 [p_h][p_w] | t2 f32 [p_h][S_w]
    uint8_t *s_frame = smem;
    float *s_t1 = reinterpret_cast<float *>(smem + align16(p.plane));
    float *s_sq = s_t1 + p.S_h * p.p_w;
    float *s_t2 = s_sq + p.p_h * p.p_w;
    const int wpr = p.S_w / 4, wpp = p.plane / 4;
    const FastDiv fd_wpr(wpr);

    for (int k = 0; k < p.K; ++k) {
        const int slot = (h + 1 + k) % p.K;
        const uint8_t *src = ring + ((size_t)n * p.K + slot) * p.plane;
        if (!CACHED) {
            for (int i = tid; i < p.plane / 16; i += kThreads)
                reinterpret_cast<uint4 *>(s_frame)[i] = __ldg(reinterpret_cast<const uint4 *>(src) + i);
            __syncthreads();
            resample_w<uint8_t>(s_frame, p.S_w, s_t1, p.p_w, p.S_h, p.sq_w, tid, kThreads);
            __syncthreads();
            resample_h<float>(s_t1, p.p_w, s_sq, p.p_w, p.p_w, p.sq_h, tid, kThreads);
            __syncthreads();
        } else {
            const float *c = pcache + ((size_t)n * p.K + slot) * p.p_h * p.p_w;
            for (int i = tid; i < p.p_h * p.p_w; i += kThreads) s_sq[i] = __ldg(c + i);
            __syncthreads();
        }
        resample_w<float>(s_sq, p.p_w, s_t2, p.S_w, p.p_h, p.ex_w, tid, kThreads);
        __syncthreads();
        uint32_t *dst = reinterpret_cast<uint32_t *>(out + ((size_t)n * p.K + k) * p.plane);
        const uint32_t *sharp_g = reinterpret_cast<const uint32_t *>(src);
        const uint32_t *sharp_s = reinterpret_cast<const uint32_t *>(s_frame);
        for (int t = tid; t < wpp; t += kThreads) {
            const int y = fd_wpr.div(t), x0 = 4 * (t - y * wpr);
            uint32_t v = resample_h_word(s_t2, p.S_w, p.ex_h, y, x0);
            if (y >= r0 && y < r0 + p.f_h) {
                const uint32_t m = word_mask(x0, c0, c0 + p.f_w);
                if (m) v = (v & ~m) | ((CACHED ? __ldg(sharp_g + t) : sharp_s[t]) & m);
            }
            dst[t] = v;
        }
        __syncthreads();
    }
}

// Fast peripheral observe (cached squeeze, bilinear-upsample expand): both expand passes are
// two-tap lerps r = b + w0 * (a - b).  The cached squeeze is biased by 49152.5 on load; lerp
// weights sum to one, so the bias rides through both passes and the rounded pixel
// floor(v + 0.5) ends up in byte 1 of the float's bit pattern (ulp there is 2^-8, so the total
// evaluation error is < 0.01 u8 LSB) — no float->int conversion, one PRMT tree per 4 pixels.
constexpr float kBias = 49152.5f;

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async4_s(uint32_t smem_addr, const void *gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_addr), "l"(gmem_src));
}
// 4-byte cp.async with compile-time byte offsets folded into the instruction's immediates
template <int SOFF, int GOFF>
__device__ __forceinline__ void cp_async4_imm(uint32_t smem_addr, const void *gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0 + %2], [%1 + %3], 4;" ::"r"(smem_addr), "l"(gmem_src), "n"(SOFF), "n"(GOFF));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- packed fp32 pairs (Blackwell FFMA2 / FADD2: two fp32 lanes per issue slot)
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, uint32_t &lo, uint32_t &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t fsub2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// Persistent CTAs (a few per SM) walk the env batch.  Per env:
//   prefetch  the cached squeeze of its K ring slots and the ring words under its fovea arrive by
//             cp.async into a double buffer, one env ahead of the arithmetic; thread 0 also applies
//             the NEXT-but-one env's sensory action to fov_loc (fov_env.py:187-199), so no separate
//             launch is needed and the prefetch knows where the fovea will be;
//   phase A   W-expand: T[k][j][x] = w0[x]*sq[k][j][i0[x]] + w1[x]*sq[k][j][i0[x]+1] + bias for the
//             K*p_h squeezed rows, two columns per thread (FFMA2), kept in shared memory;
//   phase B   H-expand + quantise + paste: one thread owns 4 adjacent columns and a segment of
//             output rows for KG frames; it holds rows b = T[j0+1] and d = T[j0]-T[j0+1] in
//             registers (reloaded only when the source row changes, ~every 4th output row) and
//             emits one 4-pixel word per frame and row: 2 FFMA2 + 3 PRMT + 1 STG.
// The bias 49152.5 makes the rounded pixel floor(v + 0.5) appear in byte 1 of the float's bit
// pattern (ulp there is 2^-8: total evaluation error < 0.01 u8 LSB), so quantisation needs no
// float->int conversion.
template <int KG, int PW>  // KG: frames per thread (K % KG == 0); PW: words per plane if known at compile time, else 0
__global__ void __launch_bounds__(128) k_observe_peripheral_v2(const __grid_constant__ DevPlan p,
                                                               const uint8_t *__restrict__ ring,
                                                               const int32_t *__restrict__ head,
                                                               const float *__restrict__ pcache,
                                                               const double *__restrict__ action,
                                                               const uint8_t *__restrict__ ctrl,
                                                               int32_t *__restrict__ loc, uint8_t *__restrict__ out,
                                                               int ysegs) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ int s_loc[2][2];
    const int tid = threadIdx.x, nt = blockDim.x;
    const int pp = p.p_h * p.p_w, K = p.K, N = p.N, f_h = p.f_h, S_w = p.S_w;
    const int quads = S_w >> 2;
    const uint32_t wpp = PW ? (uint32_t)PW : ((uint32_t)p.plane >> 2);  // words per plane
    const int nw_max = (p.f_w + 3) / 4 + 1;
    const int sq_words = (K * pp + 3) & ~3, buf_words = sq_words + ((K * f_h * nw_max + 3) & ~3);
    const int trows = K * p.p_h;
    float *bufs = reinterpret_cast<float *>(smem);  // [2]{ sq [K][p_h][p_w] | fov [K][f_h][nw_max] }
    float *s_T = bufs + 2 * buf_words;              // [K * p_h][S_w], biased
    int2 *s_row = reinterpret_cast<int2 *>(s_T + trows * S_w);  // [S_h + 1] {source row j0, bits of w0}
    const uint32_t *ring_w = reinterpret_cast<const uint32_t *>(ring);
    uint32_t *out_w = reinterpret_cast<uint32_t *>(out);
    const int G = gridDim.x;

    // ---- per-thread constants
    // phase A: column pair cp, rows part, part + parts, ...
    const int npairs = S_w >> 1, parts = nt / npairs;
    const int a_part = tid / npairs, cp = tid - a_part * npairs;
    const bool a_active = a_part < parts;
    int a_i0 = 0, a_i1 = 0;
    uint64_t a_w0 = 0, a_w1 = 0;
    if (a_active) {
        a_i0 = __ldg(p.exw_i0 + 2 * cp);
        a_i1 = __ldg(p.exw_i0 + 2 * cp + 1);
        a_w0 = pack2(__ldg(p.exw_w0 + 2 * cp), __ldg(p.exw_w0 + 2 * cp + 1));
        a_w1 = pack2(__ldg(p.exw_w1 + 2 * cp), __ldg(p.exw_w1 + 2 * cp + 1));
    }
    const uint64_t bias2 = pack2(kBias, kBias);
    // phase B: column quad q, row segment g
    const int g = tid / quads, q = tid - g * quads;
    const bool active = g < ysegs;
    const int rows_per = (p.S_h + ysegs - 1) / ysegs;
    const int y_begin = g * rows_per, y_end = active ? min(p.S_h, y_begin + rows_per) : y_begin;
    const int pp4 = pp >> 2, frows = K * f_h;
    for (int i = tid; i < p.S_h; i += nt) s_row[i] = make_int2(__ldg(p.exh_i0 + i), __float_as_int(__ldg(p.exh_w0 + i)));
    if (tid == 0) s_row[p.S_h] = make_int2(-1, 0);  // sentinel: ends the last run of rows

    auto issue = [&](int env, int lr, int lc, int hh, int b) {
        float *sq = bufs + b * buf_words;
        uint32_t *fv = reinterpret_cast<uint32_t *>(sq + sq_words);
        const size_t slot0 = (size_t)env * K;
        for (int k = 0; k < K; ++k) {  // cached squeeze: logical frame k <- ring slot (hh+1+k)%K
            int slot = hh + 1 + k;
            slot -= slot >= K ? K : 0;
            const float *src = pcache + (slot0 + slot) * pp;
            for (int c = tid; c < pp4; c += nt) cp_async16(sq + k * pp + 4 * c, src + 4 * c);
        }
        const int wq0 = lc >> 2, nw = ((lc + p.f_w - 1) >> 2) - wq0 + 1;
        for (int r = tid; r < frows; r += nt) {  // one fovea row (k, yy) per thread and pass
            const int k = r / f_h, yy = r - k * f_h;
            int slot = hh + 1 + k;
            slot -= slot >= K ? K : 0;
            const uint32_t *src = ring_w + (slot0 + slot) * wpp + (uint32_t)(lr + yy) * quads + wq0;
            uint32_t *dst = fv + r * nw_max;
            for (int w = 0; w < nw; ++w) cp_async4(dst + w, src + w);
        }
    };
    // fov_loc of env `env` after this step's sensory action; thread 0 only
    auto next_loc = [&](int env, int slot) {
        int r = 0, c = 0;
        if (env < N) update_loc_fixed<true>(p, env, action, ctrl, loc, r, c);
        s_loc[slot][0] = r;
        s_loc[slot][1] = c;
    };

    int e = blockIdx.x;
    if (tid == 0) { next_loc(e, 0); next_loc(e + G, 1); }
    __syncthreads();
    int lr = s_loc[0][0], lc = s_loc[0][1], lr1 = s_loc[1][0], lc1 = s_loc[1][1];
    int hh = 0, hh1 = 0;
    if (e < N) { hh = head[e]; issue(e, lr, lc, hh, 0); }
    cp_async_commit();
    if (e + G < N) hh1 = head[e + G];
    __syncthreads();  // s_loc is rewritten below

    for (int it = 0; e < N; e += G, ++it) {
        const int en = e + G;
        if (en < N) issue(en, lr1, lc1, hh1, (it + 1) & 1);
        cp_async_commit();
        if (tid == 0) next_loc(en + G, it & 1);
        int hh2 = 0;
        if (en + G < N) hh2 = head[en + G];
        cp_async_wait<1>();
        __syncthreads();

        const float *sq = bufs + (it & 1) * buf_words;
        // ---- phase A: W-expand every squeezed row into s_T
        if (a_active) {
            const float *s = sq + a_part * p.p_w;
            float *t = s_T + a_part * S_w + 2 * cp;
            const int sstep = parts * p.p_w, tstep = parts * S_w;
#pragma unroll 4
            for (int r = a_part; r < trows; r += parts, s += sstep, t += tstep) {
                const uint64_t u = pack2(s[a_i0], s[a_i1]), v = pack2(s[a_i0 + 1], s[a_i1 + 1]);
                *reinterpret_cast<uint64_t *>(t) = ffma2(a_w0, u, ffma2(a_w1, v, bias2));
            }
        }
        __syncthreads();

        // ---- phase B: H-expand, quantise, paste the fovea, store
        if (active) {
            const uint32_t *fvb = reinterpret_cast<const uint32_t *>(sq + sq_words);
            const uint32_t fov_mask = word_mask(4 * q, lc, lc + p.f_w);
            const int rf = fov_mask ? lr : (1 << 29);  // this column quad never meets the fovea
            const int fstride = f_h * nw_max;          // words between two frames' fovea tiles
            for (int k0 = 0; k0 < K; k0 += KG) {
                uint32_t *orow = out_w + ((size_t)e * K + k0) * wpp + (uint32_t)y_begin * quads + q;
                const uint32_t *fv = fvb + (k0 * f_h - rf) * nw_max + (q - (lc >> 2));
                const float *tk = s_T + (size_t)k0 * p.p_h * S_w + 4 * q;
                const int kstride = p.p_h * S_w;
                int y = y_begin, j_have = -2;
                const int2 *rp = s_row + y_begin;
                int2 rw = *rp;
                uint64_t b0[KG], b1[KG], d0[KG], d1[KG];
                while (y < y_end) {
                    const int j0 = rw.x;
#pragma unroll
                    for (int kk = 0; kk < KG; ++kk) {
                        const float *tj = tk + kk * kstride + j0 * S_w;
                        uint64_t a0, a1;
                        if (j0 == j_have + 1) { a0 = b0[kk]; a1 = b1[kk]; }
                        else { const ulonglong2 a = *reinterpret_cast<const ulonglong2 *>(tj); a0 = a.x; a1 = a.y; }
                        const ulonglong2 b = *reinterpret_cast<const ulonglong2 *>(tj + S_w);
                        b0[kk] = b.x; b1[kk] = b.y;
                        d0[kk] = fsub2(a0, b.x); d1[kk] = fsub2(a1, b.y);
                    }
                    j_have = j0;
                    do {
                        const float w0 = __int_as_float(rw.y);
                        const uint64_t w2 = pack2(w0, w0);
                        rw = *++rp;  // next output row (sentinel past the end)
                        uint32_t word[KG];
#pragma unroll
                        for (int kk = 0; kk < KG; ++kk) {
                            uint32_t u0, u1, u2, u3;
                            unpack2(ffma2(w2, d0[kk], b0[kk]), u0, u1);
                            unpack2(ffma2(w2, d1[kk], b1[kk]), u2, u3);
                            word[kk] = __byte_perm(__byte_perm(u0, u1, 0x0051), __byte_perm(u2, u3, 0x0051), 0x5410);
                        }
                        if ((unsigned)(y - rf) < (unsigned)f_h) {  // fovea rows: paste the sharp bytes (fov_env.py:385-386)
                            const uint32_t *sh = fv + y * nw_max;
#pragma unroll
                            for (int kk = 0; kk < KG; ++kk) word[kk] = (word[kk] & ~fov_mask) | (sh[kk * fstride] & fov_mask);
                        }
#pragma unroll
                        for (int kk = 0; kk < KG; ++kk) orow[kk * wpp] = word[kk];
                        orow += quads;
                        ++y;
                    } while (rw.x == j0 && y < y_end);
                }
            }
        }
        __syncthreads();  // s_T, s_loc and the buffer just read are rewritten next
        lr = lr1; lc = lc1; hh = hh1; hh1 = hh2;
        lr1 = s_loc[it & 1][0]; lc1 = s_loc[it & 1][1];
    }
}

// Standard geometry (obs 84x84, periphery 20x20): the H-expand pattern is known at compile time —
// output row y = 21 g + r reads squeezed rows 5 g - 1 + t(r), 5 g + t(r) with t(r) = src(r) + 1
// (clamped at the frame border) — so the row loop is fully unrolled: register-resident rows,
// immediate offsets, no index arithmetic.  The plan checks the host tables against this pattern
// before the kernel is used; the WEIGHTS always come from the host tables (ATen's values).
//
// One thread = one 4-pixel column quad q of one frame k over one 21-row segment g (21 x K x 4
// threads per env).  It W-expands the 7 squeezed rows it needs straight from the cached squeeze
// in shared memory — three adjacent samples s0..s2 cover its four columns, so
//   T[c] = a[c] * s0 + b[c] * s1 + g[c] * s2 + bias      (one of a[c], g[c] is zero)
// is 6 FFMA2 per row with per-thread constant weights — then H-expands, quantises, pastes the
// fovea and stores one word per row: 2 FFMA2 + 3 PRMT + (LDS + LOP3 on fovea rows) + 1 STG.
// The bias 49152.5 puts the rounded pixel floor(v + 0.5) into byte 1 of the float (ulp 2^-8,
// total evaluation error < 0.01 u8 LSB), so there is no float->int conversion.
// Persistent CTAs, one block-wide barrier per env.  Inputs run two envs ahead through three
// shared-memory buffers: for env e+2 one elected thread issues a single TMA bulk copy
// (cp.async.bulk, completion on an mbarrier) of the env's cached squeeze (K x 400 f32, contiguous),
// and every thread cp.asyncs exactly the ring words it will itself paste (same (row, quad) as its
// output words: immediate offsets, no index arithmetic).  Outputs are assembled in a
// double-buffered shared-memory tile and leave as ONE TMA bulk store per env (28,224 contiguous
// bytes) — 4-byte stores straight to global ran the write path at 40 % of HBM bandwidth.
// Every 32 iterations warp 0 applies the sensory actions of the CTA's next 32 envs to fov_loc,
// one env per lane (fov_env.py:187-199).
struct StdGeom {
    static constexpr int S = 84, P = 20, Q = 21, SEG = 4, R = 21, SPAN = 7;
    // floor((40 i - 64) / 168): source index of output i relative to the 20-sample axis
    __host__ __device__ static constexpr int src(int i) { return (40 * i - 64 + 168 * 4) / 168 - 4; }
};

// rows r of a 21-row segment whose bit is set in `rows`: stage ring word (row r, this quad) at fovea-tile row r
template <int NW, int R0>
__device__ __forceinline__ void prefetch_rows_imm(uint32_t rows, uint32_t dst, const uint32_t *src) {
    if constexpr (R0 < StdGeom::R) {
        if (rows & (1u << R0)) cp_async4_imm<R0 * (NW ? NW : 1) * 4, R0 * StdGeom::Q * 4>(dst, src);
        prefetch_rows_imm<NW, R0 + 1>(rows, dst, src);
    }
}

template <int K, int NW>  // NW: words per staged fovea row, (f_w + 3) / 4 + 1, when baked in; 0 = from the plan
__global__ void __launch_bounds__(((StdGeom::Q * StdGeom::SEG * K + 31) / 32) * 32, 2)
    k_observe_peripheral_std(const __grid_constant__ DevPlan p, const __grid_constant__ ExpandStd ew,
                             const uint8_t *__restrict__ ring, const int32_t *__restrict__ head,
                             const float *__restrict__ pcache, const double *__restrict__ action,
                             const uint8_t *__restrict__ ctrl, int32_t *__restrict__ loc, uint8_t *__restrict__ out) {
    using Gm = StdGeom;
    constexpr int S = Gm::S, P = Gm::P, Q = Gm::Q, R = Gm::R;
    constexpr int NB = Q * Gm::SEG * K;
    constexpr int PP = P * P, PLANE_W = S * S / 4;
    constexpr int NBUF = 3, DIST = 2, LOC_RING = 128;
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ int4 s_loc[LOC_RING];  // {fov row, fov col, head, -} of the CTA's iteration j at [j % LOC_RING]
    __shared__ __align__(8) uint64_t full[NBUF];
    const int tid = threadIdx.x;
    const int N = p.N, f_h = p.f_h, f_w = p.f_w, G = gridDim.x;
    const int nw_max = NW ? NW : (f_w + 3) / 4 + 1;
    const int buf_words = K * PP + ((K * f_h * nw_max + 3) & ~3);
    float *bufs = reinterpret_cast<float *>(smem);  // [NBUF]{ sq [K slots][P][P] | fov [K][f_h][nw_max] }
    uint32_t *tiles = reinterpret_cast<uint32_t *>(bufs + NBUF * buf_words);  // [2][K][S][S / 4] output words
    const uint32_t *ring_w = reinterpret_cast<const uint32_t *>(ring);

    // fov_loc (after this step's action) and head of iterations [it0, it0 + 32), one per lane (warp 0)
    auto loc_batch = [&](int it0) {
        const int j = it0 + (tid & 31);
        // j * G cannot overflow: a CTA runs at most N / G + 1 iterations and batches reach 64 past that
        const int env = j <= N / G + 1 ? (int)blockIdx.x + j * G : N;
        int r = 0, c = 0, hh = 0;
        if (env < N) {
            update_loc_fixed<true>(p, env, action, ctrl, loc, r, c);
            hh = head[env];
        }
        s_loc[j & (LOC_RING - 1)] = make_int4(r, c, hh, 0);
    };
    if (tid == 0) {
        for (int i = 0; i < NBUF; ++i) mbar_init(&full[i], 1);
        mbar_fence_init();
    }
    if (tid < 32) loc_batch(0);

    const bool active = tid < NB;
    const int q = tid % Q, t2 = tid / Q;
    // which (frame, segment) group the t2-th run of 21 threads works on: an order found by search that keeps
    // the two or three groups sharing a warp on different shared-memory banks when they store their tile words
    // (word offset 1764 k + 441 g + 21 r + q); in k-major order every warp's stores were 2-way conflicts
    constexpr int kOrd4[16] = {7, 4, 13, 3, 0, 1, 9, 6, 10, 14, 11, 12, 8, 5, 15, 2};
    constexpr int kOrd3[12] = {11, 8, 5, 2, 9, 6, 3, 0, 10, 4, 1, 7};
    const int grp = !active ? 0 : (K == 4 ? kOrd4[t2 & 15] : (K == 3 ? kOrd3[t2 % 12] : t2));
    const int g = grp % Gm::SEG, k = grp / Gm::SEG;
    // W pass: samples base .. base + 2 of a squeezed row cover this quad's columns
    int base;
    uint64_t wa01, wa23, wb01, wb23, wc01, wc23;  // weights of s0 / s1 / s2 for columns (0,1) and (2,3)
    {
        const int i00 = __ldg(p.exw_i0 + 4 * q);
        base = min(i00, P - 3);
        float wa[4], wb[4], wc[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int sel = __ldg(p.exw_i0 + 4 * q + c) - base;  // 0 or 1
            const float w0 = __ldg(p.exw_w0 + 4 * q + c), w1 = __ldg(p.exw_w1 + 4 * q + c);
            wa[c] = sel == 0 ? w0 : 0.f;
            wb[c] = sel == 0 ? w1 : w0;
            wc[c] = sel == 0 ? 0.f : w1;
        }
        wa01 = pack2(wa[0], wa[1]); wa23 = pack2(wa[2], wa[3]);
        wb01 = pack2(wb[0], wb[1]); wb23 = pack2(wb[2], wb[3]);
        wc01 = pack2(wc[0], wc[1]); wc23 = pack2(wc[2], wc[3]);
    }
    const uint64_t bias2 = pack2(kBias, kBias);
    // squeezed rows 5g-1 .. 5g+5, clamped to the frame: float offsets inside one staged slot
    const int row_first = 5 * g - 1;
    const int so_t0 = (row_first < 0 ? 0 : row_first) * P + base;                 // t = 0
    const int so_mid = row_first * P + base;                                      // t = 1..5 at + t * P
    const int so_t6 = (row_first + 6 > P - 1 ? P - 1 : row_first + 6) * P + base;  // t = 6
    const uint32_t bufs_s = smem_u32(bufs);
    const uint32_t word0 = (uint32_t)(g * R) * Q + q;  // this thread's first output word inside a plane

    // one bit per row r of this segment that meets the fovea at (lr, lc); 0 if the quad misses it
    auto fovea_rows = [&](int lr, int lc, uint32_t &mask) {
        mask = word_mask(4 * q, lc, lc + f_w);
        const int r_lo = lr - g * R;
        const int m_lo = max(r_lo, 0), m_hi = min(r_lo + f_h, R);
        return (mask && m_hi > m_lo) ? (((1u << m_hi) - 1u) & ~((1u << m_lo) - 1u)) : 0u;
    };
    // staged position (word index inside a buffer's fovea tile) of this thread's row r = 0
    auto fovea_pos = [&](int lr, int lc) { return (k * f_h + g * R - lr) * nw_max + (q - (lc >> 2)); };
    // prefetch for iteration j (env) into buffer j % NBUF
    auto prefetch = [&](int env, int j) {
        // buffer j % NBUF: its last tenant (iteration j - NBUF) was consumed before the barrier that
        // ended iteration j - DIST - 1, which every thread has passed
        const int b = j % NBUF;
        if (tid == 0) {
            mbar_expect_tx(&full[b], K * PP * 4);
            bulk_g2s(bufs + b * buf_words, pcache + (size_t)env * (K * PP), K * PP * 4, &full[b]);
        }
        if (active) {
            const int4 lh = s_loc[j & (LOC_RING - 1)];
            uint32_t mask;
            const uint32_t rows = fovea_rows(lh.x, lh.y, mask);
            if (rows) {
                int slot = lh.z + 1 + k;
                slot -= slot >= K ? K : 0;
                const uint32_t *src = ring_w + ((size_t)env * K + slot) * PLANE_W + word0;
                const uint32_t dst = bufs_s + (uint32_t)(b * buf_words + K * PP + fovea_pos(lh.x, lh.y)) * 4u;
                if (NW) {
                    prefetch_rows_imm<NW, 0>(rows, dst, src);
                } else {
#pragma unroll
                    for (int r = 0; r < R; ++r)
                        if (rows & (1u << r)) cp_async4_s(dst + (uint32_t)(r * nw_max) * 4u, src + r * Q);
                }
            }
        }
    };
    // W-expanded, biased squeezed row at float offset `so` of the staged squeeze
    auto t_row = [&](const float *sq, int so, uint64_t &t01, uint64_t &t23) {
        const float s0 = sq[so], s1 = sq[so + 1], s2 = sq[so + 2];
        const uint64_t p0 = pack2(s0, s0), p1 = pack2(s1, s1), p2 = pack2(s2, s2);
        t01 = ffma2(p2, wc01, ffma2(p1, wb01, ffma2(p0, wa01, bias2)));
        t23 = ffma2(p2, wc23, ffma2(p1, wb23, ffma2(p0, wa23, bias2)));
    };

    __syncthreads();  // s_loc of the first 32 iterations, mbarriers initialised
    {
        const int e0 = blockIdx.x;
#pragma unroll
        for (int j = 0; j < DIST; ++j) {
            if (e0 + j * G < N) prefetch(e0 + j * G, j);
            cp_async_commit();
        }
    }
    int it = 0;
    for (int e = blockIdx.x; e < N; e += G, ++it) {
        if (e + DIST * G < N) prefetch(e + DIST * G, it + DIST);
        cp_async_commit();
        if ((it & 31) == 0 && tid < 32) loc_batch(it + 32);
        const int b = it % NBUF;
        cp_async_wait<DIST>();                       // this thread's own fovea words of env e
        mbar_wait(&full[b], (it / NBUF) & 1);        // the env's cached squeeze (TMA)
        uint32_t *tile = tiles + (it & 1) * (K * PLANE_W);
        if (active) {
            const int4 lh = s_loc[it & (LOC_RING - 1)];
            uint32_t fov_mask;
            const uint32_t rows = fovea_rows(lh.x, lh.y, fov_mask);
            int slot = lh.z + 1 + k;
            slot -= slot >= K ? K : 0;
            const float *sq = bufs + b * buf_words + slot * PP;
            const uint32_t *sh = reinterpret_cast<const uint32_t *>(bufs + b * buf_words + K * PP) + fovea_pos(lh.x, lh.y);
            uint32_t *o = tile + k * PLANE_W + word0;
            uint64_t a0, a1, b0, b1, d0 = 0, d1 = 0;
            t_row(sq, so_t0, a0, a1);
            t_row(sq, so_mid + P, b0, b1);
            int t_have = 0;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int t = Gm::src(r) + 1;  // compile time: 0,0,1,1,1,1,2,...
                if (t != t_have) {              // resolved at compile time after unrolling
                    a0 = b0; a1 = b1;
                    t_row(sq, t + 1 == Gm::SPAN - 1 ? so_t6 : so_mid + (t + 1) * P, b0, b1);
                    t_have = t;
                }
                if (r == 0 || Gm::src(r) != Gm::src(r - 1)) { d0 = fsub2(a0, b0); d1 = fsub2(a1, b1); }
                // H weight of row r: the same for every segment (the plan checked it); at the clamped
                // border rows a == b, so the weight does not matter there
                const uint64_t w2 = pack2(ew.hw[r], ew.hw[r]);
                uint32_t u0, u1, u2, u3;
                unpack2(ffma2(w2, d0, b0), u0, u1);
                unpack2(ffma2(w2, d1, b1), u2, u3);
                uint32_t word = __byte_perm(__byte_perm(u0, u1, 0x0051), __byte_perm(u2, u3, 0x0051), 0x5410);
                if (rows & (1u << r))  // fovea rows: paste the sharp bytes (fov_env.py:385-386)
                    word = (word & ~fov_mask) | (sh[r * nw_max] & fov_mask);
                o[r * Q] = word;
            }
        }
        // the tile leaves as one TMA store; the store issued two iterations ago has finished reading
        // this iteration's tile buffer's twin before anyone writes it again (next iteration)
        fence_async_smem();
        if (tid == 0) bulk_wait_read<0>();
        __syncthreads();
        if (tid == 0) {
            bulk_s2g(out + (size_t)e * (K * PLANE_W * 4), tile, K * PLANE_W * 4);
            bulk_commit();
        }
    }
    if (tid == 0) bulk_wait_read<0>();  // shared memory must outlive the last store's reads
}

// --------------------------------------------------------------------- observe: flexible
// FlexibleFovealEnv._fov_step + _get_fov_state (fov_env.py:270-330).
__device__ __forceinline__ AxisRef flex_axis(const DevPlan &p, int axis, int family, int r, int n_in) {
    const FlexEntry e = p.flex[(axis * 3 + family) * (p.S_max + 1) + r];
    AxisRef a;
    a.xmin = p.pool_i + e.xmin_off;
    a.w = reinterpret_cast<const float *>(p.pool_i + e.w_off);
    a.n_in = n_in; a.n_out = e.n_out; a.taps = e.taps;
    return a;
}

template <int VARIANT>
__global__ void __launch_bounds__(kThreads) k_observe_flexible(const __grid_constant__ DevPlan p,
                                                               const uint8_t *__restrict__ ring,
                                                               const int32_t *__restrict__ head,
                                                               const double *__restrict__ action,
                                                               const int32_t *__restrict__ atype,
                                                               const uint8_t *__restrict__ ctrl,
                                                               int32_t *__restrict__ loc, int32_t *__restrict__ res,
                                                               int pad_h, int pad_w, uint8_t *__restrict__ out) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ int s_win[4];
    const int n = blockIdx.x, tid = threadIdx.x;
    if (tid == 0) {
        const int mode = ctrl ? ctrl[n] : AGYM_FOV_APPLY;
        int r = loc[2 * n], c = loc[2 * n + 1], rh = res[2 * n], rw = res[2 * n + 1];
        if (mode == AGYM_FOV_RESET) {
            r = p.init_r; c = p.init_c; rh = p.f_h; rw = p.f_w;
        } else if (mode == AGYM_FOV_APPLY) {
            const double a0 = action[2 * n], a1 = action[2 * n + 1];
            const int t = atype ? atype[n] : AGYM_ATYPE_FOV_LOC;
            if (t == AGYM_ATYPE_FOV_RES) {
                // fov_res = action (fov_env.py:323); the reference raises for res > obs, the
                // device clamps to [1, S] instead (the Python layer validates beforehand)
                rh = min(max((int)a0, 1), p.S_h);
                rw = min(max((int)a1, 1), p.S_w);
                r = clip_rint((double)r, 0.0, (double)(p.S_h - rh));
                c = clip_rint((double)c, 0.0, (double)(p.S_w - rw));
            } else {
                double v0 = a0, v1 = a1;
                if (p.relative) {
                    v0 = (double)(r + clip_rint(a0, p.lo, p.hi));
                    v1 = (double)(c + clip_rint(a1, p.lo, p.hi));
                }
                r = clip_rint(v0, 0.0, (double)(p.S_h - rh));
                c = clip_rint(v1, 0.0, (double)(p.S_w - rw));
            }
        }
        loc[2 * n] = r; loc[2 * n + 1] = c; res[2 * n] = rh; res[2 * n + 1] = rw;
        s_win[0] = r; s_win[1] = c; s_win[2] = rh; s_win[3] = rw;
    }
    __syncthreads();
    const int r0 = s_win[0], c0 = s_win[1], rh = s_win[2], rw = s_win[3];
    const int h = head[n];
    const bool blur = rh > p.f_h;  // row dimension only (fov_env.py:286)

    float *bufA = reinterpret_cast<float *>(smem);
    float *bufB = bufA + p.plane;
    const FastDiv fd_rw(rw);
    const int oh = VARIANT == AGYM_OUT_CROP ? pad_h : p.S_h, ow = VARIANT == AGYM_OUT_CROP ? pad_w : p.S_w;
    const int oy = VARIANT == AGYM_OUT_MASK ? r0 : 0, ox = VARIANT == AGYM_OUT_MASK ? c0 : 0;
    const int wpr = ow / 4, wpp = oh * ow / 4;
    const FastDiv fd_wpr(wpr);

    for (int k = 0; k < p.K; ++k) {
        const uint8_t *src = ring + ((size_t)n * p.K + (h + 1 + k) % p.K) * p.plane;
        for (int i = tid; i < rh * rw; i += kThreads) {
            const int y = fd_rw.div(i), x = i - y * rw;
            bufA[i] = (float)__ldg(src + (r0 + y) * p.S_w + c0 + x);
        }
        __syncthreads();
        if (blur) {  // Resize(fov_size) then Resize(fov_res) (fov_env.py:276-280)
            resample_w<float>(bufA, rw, bufB, p.f_w, rh, flex_axis(p, 1, 0, rw, rw), tid, kThreads);
            __syncthreads();
            resample_h<float>(bufB, p.f_w, bufA, p.f_w, p.f_w, flex_axis(p, 0, 0, rh, rh), tid, kThreads);
            __syncthreads();
            resample_w<float>(bufA, p.f_w, bufB, rw, p.f_h, flex_axis(p, 1, 1, rw, p.f_w), tid, kThreads);
            __syncthreads();
            resample_h<float>(bufB, rw, bufA, rw, rw, flex_axis(p, 0, 1, rh, p.f_h), tid, kThreads);
            __syncthreads();
        }
        uint32_t *dst = reinterpret_cast<uint32_t *>(out + ((size_t)n * p.K + k) * oh * ow);
        if (VARIANT == AGYM_OUT_RESIZE_FULL) {
            resample_w<float>(bufA, rw, bufB, p.S_w, rh, flex_axis(p, 1, 2, rw, rw), tid, kThreads);
            __syncthreads();
            const AxisRef ah = flex_axis(p, 0, 2, rh, rh);
            for (int t = tid; t < wpp; t += kThreads) {
                const int y = fd_wpr.div(t), x0 = 4 * (t - y * wpr);
                dst[t] = resample_h_word(bufB, p.S_w, ah, y, x0);
            }
        } else {
            for (int t = tid; t < wpp; t += kThreads) {
                const int y = fd_wpr.div(t), x0 = 4 * (t - y * wpr);
                uint32_t word = 0u;
                const int py = y - oy;
                if (py >= 0 && py < rh) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int px = x0 + i - ox;
                        if (px >= 0 && px < rw) word |= quant_u8(bufA[py * rw + px]) << (8 * i);
                    }
                }
                dst[t] = word;
            }
        }
        __syncthreads();
    }
}

// Fast flexible fovea for the paste-type outputs (mask_out in place, or the zero-padded crop):
// the blur Resize(fov_size) -> Resize(fov_res) (fov_env.py:276-280) is applied as ONE banded operator
// per axis (host-composed, see build_blur_axis), i.e. two passes instead of four; the env's K windows
// are staged once as aligned words; the output frame is assembled in a shared-memory tile (zeros +
// window) and leaves as one TMA bulk store.
template <int VARIANT>
__global__ void __launch_bounds__(kThreads) k_observe_flexible_fast(const __grid_constant__ DevPlan p,
                                                                    const uint8_t *__restrict__ ring,
                                                                    const int32_t *__restrict__ head,
                                                                    const double *__restrict__ action,
                                                                    const int32_t *__restrict__ atype,
                                                                    const uint8_t *__restrict__ ctrl,
                                                                    int32_t *__restrict__ loc, int32_t *__restrict__ res,
                                                                    int oh, int ow, int t1_cap, uint8_t *__restrict__ out) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ int s_win[4];
    const int n = blockIdx.x, tid = threadIdx.x;
    const int K = p.K, quads = p.S_w >> 2, xwm = quads + 1;
    const int tile_bytes = K * oh * ow;
    uint32_t *s_tile = reinterpret_cast<uint32_t *>(smem);                       // [K][oh][ow] bytes
    uint32_t *s_x = s_tile + (tile_bytes >> 2);                                  // [K][S_h][xwm] window words
    float *s_t1 = reinterpret_cast<float *>(s_x + K * p.S_h * xwm);              // [rh][rw]
    float *s_ww = s_t1 + t1_cap;                                                 // [S_w][blur_tmax]
    float *s_wh = s_ww + p.S_w * p.blur_tmax;                                    // [S_h][blur_tmax]
    int32_t *s_xw = reinterpret_cast<int32_t *>(s_wh + p.S_h * p.blur_tmax);     // [S_w] first tap
    int32_t *s_xh = s_xw + p.S_w;                                                // [S_h]
    if (tid == 0) {
        const int mode = ctrl ? ctrl[n] : AGYM_FOV_APPLY;
        int r = loc[2 * n], c = loc[2 * n + 1], rh = res[2 * n], rw = res[2 * n + 1];
        if (mode == AGYM_FOV_RESET) {
            r = p.init_r; c = p.init_c; rh = p.f_h; rw = p.f_w;
        } else if (mode == AGYM_FOV_APPLY) {
            const double a0 = action[2 * n], a1 = action[2 * n + 1];
            const int t = atype ? atype[n] : AGYM_ATYPE_FOV_LOC;
            if (t == AGYM_ATYPE_FOV_RES) {  // fov_res = action, then re-clamp loc (fov_env.py:322-324)
                rh = min(max((int)a0, 1), p.S_h);
                rw = min(max((int)a1, 1), p.S_w);
                r = clip_rint((double)r, 0.0, (double)(p.S_h - rh));
                c = clip_rint((double)c, 0.0, (double)(p.S_w - rw));
            } else {
                double v0 = a0, v1 = a1;
                if (p.relative) {
                    v0 = (double)(r + clip_rint(a0, p.lo, p.hi));
                    v1 = (double)(c + clip_rint(a1, p.lo, p.hi));
                }
                r = clip_rint(v0, 0.0, (double)(p.S_h - rh));
                c = clip_rint(v1, 0.0, (double)(p.S_w - rw));
            }
        }
        loc[2 * n] = r; loc[2 * n + 1] = c; res[2 * n] = rh; res[2 * n + 1] = rw;
        s_win[0] = r; s_win[1] = c; s_win[2] = rh; s_win[3] = rw;
    }
    {   // zero frame (everything outside the window stays zero)
        const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
        for (int i = tid; i < (tile_bytes >> 4); i += kThreads) reinterpret_cast<uint4 *>(s_tile)[i] = z4;
    }
    __syncthreads();
    const int r0 = s_win[0], c0 = s_win[1], rh = s_win[2], rw = s_win[3];
    const int h = head[n];
    const bool blur = rh > p.f_h;  // row dimension only (fov_env.py:286)
    const int oy = VARIANT == AGYM_OUT_MASK ? r0 : 0, ox = VARIANT == AGYM_OUT_MASK ? c0 : 0;
    // crop output may be narrower / shorter than the window only if the caller's pad is: clip
    const int vh = min(rh, oh - oy), vw = min(rw, ow - ox);

    // ---- stage the K windows as aligned words: row y of frame k at s_x[(k * S_h + y) * xwm ...]
    // (byte loads straight from the ring were measured 10 % slower)
    const int wq0 = c0 >> 2, nwx = ((c0 + rw - 1) >> 2) - wq0 + 1, cb = c0 & 3;
    {
        const uint32_t *ring_w = reinterpret_cast<const uint32_t *>(ring) + (size_t)n * K * (p.plane >> 2);
        const FastDiv fd_w(nwx), fd_h(rh);
        for (int i = tid; i < K * rh * nwx; i += kThreads) {
            const int row = fd_w.div(i), w = i - row * nwx;
            const int k = fd_h.div(row), y = row - k * rh;
            int slot = h + 1 + k;
            slot -= slot >= K ? K : 0;
            s_x[(k * p.S_h + y) * xwm + w] = __ldg(ring_w + (size_t)slot * (p.plane >> 2) + (r0 + y) * quads + wq0 + w);
        }
    }
    int tw = 1, th = 1;
    if (blur) {  // this env's two banded operators -> shared memory
        const FlexEntry ew = p.flexb[(p.S_max + 1) + rw], eh = p.flexb[rh];
        tw = ew.taps; th = eh.taps;
        const float *gw = reinterpret_cast<const float *>(p.pool_i + ew.w_off), *gh = reinterpret_cast<const float *>(p.pool_i + eh.w_off);
        for (int i = tid; i < rw * tw; i += kThreads) s_ww[i] = __ldg(gw + i);
        for (int i = tid; i < rh * th; i += kThreads) s_wh[i] = __ldg(gh + i);
        for (int i = tid; i < rw; i += kThreads) s_xw[i] = __ldg(p.pool_i + ew.xmin_off + i);
        for (int i = tid; i < rh; i += kThreads) s_xh[i] = __ldg(p.pool_i + eh.xmin_off + i);
    }
    __syncthreads();
    const uint8_t *xb = reinterpret_cast<const uint8_t *>(s_x) + cb;
    uint8_t *tb = reinterpret_cast<uint8_t *>(s_tile);
    const FastDiv fd_rw(rw);
    if (!blur) {  // the window itself, bit exact
        const FastDiv fd_rh(rh);
        for (int i = tid; i < K * rh * rw; i += kThreads) {
            const int row = fd_rw.div(i), x = i - row * rw;
            const int k = fd_rh.div(row), y = row - k * rh;
            if (y < vh && x < vw) tb[(k * oh + oy + y) * ow + ox + x] = xb[((k * p.S_h + y) * xwm) * 4 + x];
        }
    } else {
        // t1 rows are padded to a multiple of 4 floats so that the H pass reads float4; as many frames
        // per pass as fit (all K for windows up to ~50 x 52), so an env costs 2 barriers instead of 2K
        const int rwp = (rw + 3) & ~3, per = rh * rwp;
        const int kg = K * per <= t1_cap ? K : 1;
        const FastDiv fd_fr(rh * rw), fd_q(rwp >> 2), fd_frq(rh * (rwp >> 2));
        for (int k0 = 0; k0 < K; k0 += kg) {
            // W pass: t1[y][x] = sum_t Mw[x][t] * X[y][xw[x] + t]
            for (int i = tid; i < kg * rh * rw; i += kThreads) {
                const int kk = fd_fr.div(i), rem = i - kk * rh * rw;
                const int y = fd_rw.div(rem), x = rem - y * rw;
                const uint8_t *src = xb + (((k0 + kk) * p.S_h + y) * xwm) * 4 + s_xw[x];
                const float *w = s_ww + x * tw;
                float acc = 0.f;
                for (int t = 0; t < tw; ++t) acc = fmaf(w[t], (float)src[t], acc);
                s_t1[kk * per + y * rwp + x] = acc;
            }
            __syncthreads();
            // H pass, 4 columns per thread + quantise + paste: out[y][x] = sum_t Mh[y][t] * t1[xh[y] + t][x]
            for (int i = tid; i < kg * rh * (rwp >> 2); i += kThreads) {
                const int kk = fd_frq.div(i), rem = i - kk * rh * (rwp >> 2);
                const int y = fd_q.div(rem), x0 = 4 * (rem - y * (rwp >> 2));
                const float *src = s_t1 + kk * per + s_xh[y] * rwp + x0;
                const float *w = s_wh + y * th;
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int t = 0; t < th; ++t) {
                    const float4 v = *reinterpret_cast<const float4 *>(src + t * rwp);
                    const float wt = w[t];
                    acc.x = fmaf(wt, v.x, acc.x); acc.y = fmaf(wt, v.y, acc.y);
                    acc.z = fmaf(wt, v.z, acc.z); acc.w = fmaf(wt, v.w, acc.w);
                }
                if (y < vh) {
                    uint8_t *o = tb + ((k0 + kk) * oh + oy + y) * ow + ox + x0;
                    if (x0 < vw) o[0] = (uint8_t)quant_u8(acc.x);
                    if (x0 + 1 < vw) o[1] = (uint8_t)quant_u8(acc.y);
                    if (x0 + 2 < vw) o[2] = (uint8_t)quant_u8(acc.z);
                    if (x0 + 3 < vw) o[3] = (uint8_t)quant_u8(acc.w);
                }
            }
            __syncthreads();
        }
    }
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
        bulk_s2g(out + (size_t)n * tile_bytes, s_tile, (uint32_t)tile_bytes);
        bulk_commit();
        bulk_wait_read<0>();
    }
}

// Crop variant, second version: still one warp per env, but the K windows are first staged in the warp's
// own shared-memory slice as ALIGNED words (25 independent 4-byte loads per lane instead of 85 byte loads in
// five dependent rounds: one DRAM round trip per env), and the output words are then gathered from shared
// memory through a CTA-wide table of byte offsets that is the same for every env (it only depends on K, f).
// Requires f_h * f_w % 4 == 0 is NOT needed: the table is flat over the K * f_h * f_w output bytes.
constexpr int kCropWarps = 8;
__global__ void __launch_bounds__(kCropWarps * 32) k_observe_fixed_crop_v2(const __grid_constant__ DevPlan p,
                                                                           const uint8_t *__restrict__ ring,
                                                                           const int32_t *__restrict__ head,
                                                                           const double *__restrict__ action,
                                                                           const uint8_t *__restrict__ ctrl,
                                                                           int32_t *__restrict__ loc, uint8_t *__restrict__ out,
                                                                           int nwx) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K = p.K, per_k = p.f_h * p.f_w, words = (K * per_k) >> 2, rows = K * p.f_h, quads = p.S_w >> 2;
    const int stride = nwx * 4;                                   // staged bytes per window row
    uint2 *s_off = reinterpret_cast<uint2 *>(smem);               // [words] 4 x u16: staged byte offset of each output byte
    uint32_t *s_win = reinterpret_cast<uint32_t *>(smem + align16((size_t)words * 8)) + warp * rows * nwx;
    const int n = blockIdx.x * kCropWarps + warp;
    const bool valid = n < p.N;
    // the loads of the fov update go out first: their latency hides behind the table build
    LocIn li;
    li.a0 = li.a1 = 0.0; li.r = li.c = 0; li.mode = AGYM_FOV_KEEP;
    if (valid && lane == 0) li = load_loc_in(n, action, ctrl, loc);
    const int h = valid ? head[n] : 0;
    {
        const FastDiv fd_w(p.f_w);
        for (int t = tid; t < words; t += kCropWarps * 32) {
            uint32_t o[4];
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int idx = 4 * t + b, row = fd_w.div(idx), x = idx - row * p.f_w;  // row = k * f_h + y
                o[b] = (uint32_t)(row * stride + x);
            }
            s_off[t] = make_uint2(o[0] | (o[1] << 16), o[2] | (o[3] << 16));
        }
    }
    int r0 = 0, c0 = 0;
    if (valid && lane == 0) {
        apply_loc(p, li, r0, c0);
        loc[2 * n] = r0;
        loc[2 * n + 1] = c0;
    }
    r0 = __shfl_sync(0xffffffffu, r0, 0);
    c0 = __shfl_sync(0xffffffffu, c0, 0);
    const int wq0 = c0 >> 2, cb = c0 & 3;
    if (valid) {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(ring) + (size_t)n * K * (p.plane >> 2) + r0 * quads + wq0;
        const FastDiv fd_n(nwx), fd_h(p.f_h);
        const int wmax = quads - 1 - wq0;   // last word of a frame row: the spare staged word must not run past the ring
#pragma unroll 4
        for (int i = lane; i < rows * nwx; i += 32) {
            const int row = fd_n.div(i), w = i - row * nwx;
            const int k = fd_h.div(row), y = row - k * p.f_h;
            int slot = h + 1 + k;
            slot -= slot >= K ? K : 0;
            cp_async4(s_win + i, src + slot * (p.plane >> 2) + y * quads + min(w, wmax));
        }
    }
    cp_async_commit();
    __syncthreads();   // offset table complete
    if (!valid) return;
    cp_async_wait<0>();
    __syncwarp();
    const uint8_t *wb = reinterpret_cast<const uint8_t *>(s_win) + cb;
    uint32_t *dst = reinterpret_cast<uint32_t *>(out + (size_t)n * K * per_k);
#pragma unroll 4
    for (int t = lane; t < words; t += 32) {
        const uint2 o = s_off[t];
        const uint32_t b0 = wb[o.x & 0xffffu], b1 = wb[o.x >> 16], b2 = wb[o.y & 0xffffu], b3 = wb[o.y >> 16];
        dst[t] = __byte_perm(__byte_perm(b0, b1, 0x0040), __byte_perm(b2, b3, 0x0040), 0x5410);
    }
}

// Persistent flexible fovea (mask_out in place, or the zero-padded crop): the successor of
// k_observe_flexible_fast.  Two CTAs per SM pull envs from a device counter (windows of 20..50 pixels,
// blurred or not, make an env's cost vary 3x: a static split left a quarter of the SM time idle) and
// keep three things in flight:
//   * thread 0 claims the env two iterations ahead and applies its sensory action (fov_env.py:314-330):
//     the atomic, the loads and the arithmetic are spread over the phases of the current env;
//   * the K windows of the next env (and its W operator) travel by cp.async during the current H pass;
//   * the finished frame tile leaves as one TMA bulk store while the next env is computed.
// The blur Resize(fov_size) -> Resize(fov_res) (fov_env.py:276-280) is one banded operator per axis.
// Along W it runs on the staged bytes in 16-bit fixed point: 8 taps = 4 IDP.2A on a funnel-shifted
// 8-byte window, two rows per thread, weights * 2^16 summing to 2^16 exactly (<= 255 * taps / 2^17 LSB
// from the fp64 weights: 0.01 LSB for the 5-tap operators of windows up to 50).  Along H it is packed
// fp32 (FFMA2) on four columns per thread; results are rounded to nearest-even with the 1.5 * 2^23 bias
// and leave as whole words.
#ifndef AGYM_FLEX_THREADS
#define AGYM_FLEX_THREADS 256
#endif
constexpr int kFlexThreads = AGYM_FLEX_THREADS;

__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

struct FlexGeom {
    int rh, rwp, vh, nq, ow4, oh, wlo, oy;
    uint32_t m_first, m_last;  // byte masks of the first / last output word of a window row
};

// H pass over kc frames: item = (frame row, output word); s_wh2 holds every weight twice (FFMA2 operand)
template <int TH>
__device__ __forceinline__ void flex_hpass(const FlexGeom &g, const float *s_t1, const uint64_t *s_wh2, const int32_t *s_xh,
                                           uint32_t *s_tile, int k0, int kc, int th, int tid, const uint32_t *s_magic) {
    const FastDiv fd_nq(g.nq, s_magic), fd_rh(g.rh, s_magic);
    const int nrows = kc * g.rh, dr = fd_nq.div(kFlexThreads), dq = kFlexThreads - dr * g.nq;
    const uint64_t rne2 = pack2(12582912.f, 12582912.f);  // 1.5 * 2^23: v + bias has rint(v) (half to even) in its low byte
    int row = fd_nq.div(tid), q = tid - row * g.nq;
    while (row < nrows) {
        const int kk = fd_rh.div(row), y = row - kk * g.rh;
        if (y < g.vh) {
            const float *src = s_t1 + (kk * g.rh + s_xh[y]) * g.rwp + 4 * q;
            const uint64_t *w = s_wh2 + y * th;
            uint64_t a01 = 0ull, a23 = 0ull;
            if (TH > 0) {
#pragma unroll
                for (int t = 0; t < TH; ++t) {
                    const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(src + t * g.rwp);
                    const uint64_t wt = w[t];
                    a01 = ffma2(v.x, wt, a01);
                    a23 = ffma2(v.y, wt, a23);
                }
            } else {
                for (int t = 0; t < th; ++t) {
                    const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(src + t * g.rwp);
                    const uint64_t wt = w[t];
                    a01 = ffma2(v.x, wt, a01);
                    a23 = ffma2(v.y, wt, a23);
                }
            }
            uint32_t b0, b1, b2, b3;
            unpack2(fadd2(a01, rne2), b0, b1);
            unpack2(fadd2(a23, rne2), b2, b3);
            uint32_t word = __byte_perm(__byte_perm(b0, b1, 0x0040), __byte_perm(b2, b3, 0x0040), 0x5410);
            if (q == 0) word &= g.m_first;
            if (q == g.nq - 1) word &= g.m_last;
            s_tile[((k0 + kk) * g.oh + g.oy + y) * g.ow4 + g.wlo + q] = word;
        }
        q += dq; row += dr;
        if (q >= g.nq) { q -= g.nq; ++row; }
    }
}

// W pass over nrows staged rows: t1[row][sb + x] = 2^-16 * sum_t q[x][t] * X[row][xw[x] + t]; a thread takes
// column x of rows rp and rp + ceil(nrows / 2) (same weights and shift, two independent IDP.2A chains)
template <int NH>
__device__ __forceinline__ void flex_wpass(const uint32_t *xrow0, int nwxp, const int32_t *s_xw, const uint32_t *s_wq,
                                           float *s_t1, int rwp, int sb, int cb, int rw, int nrows, int tid, const uint32_t *s_magic) {
    const FastDiv fd_rw(rw, s_magic);
    const int nrp = (nrows + 1) >> 1, dr = fd_rw.div(kFlexThreads), dx = kFlexThreads - dr * rw;
    int rp = fd_rw.div(tid), x = tid - rp * rw;
    while (rp < nrp) {
        const int b = cb + s_xw[x];
        const uint32_t sh = (uint32_t)(b & 3) * 8u;
        const uint32_t *sp0 = xrow0 + rp * nwxp + (b >> 2), *sp1 = sp0 + nrp * nwxp;
        const bool two = rp + nrp < nrows;
        const uint4 *wq = reinterpret_cast<const uint4 *>(s_wq) + x * NH;
        uint32_t acc0 = 0u, acc1 = 0u;
#pragma unroll
        for (int hh = 0; hh < NH; ++hh) {
            const uint4 q = wq[hh];
            {
                const uint32_t a0 = sp0[2 * hh], a1 = sp0[2 * hh + 1], a2 = sp0[2 * hh + 2];
                const uint32_t lo = __funnelshift_r(a0, a1, sh), hi = __funnelshift_r(a1, a2, sh);
                acc0 = __dp2a_lo(q.x, lo, acc0); acc0 = __dp2a_hi(q.y, lo, acc0);
                acc0 = __dp2a_lo(q.z, hi, acc0); acc0 = __dp2a_hi(q.w, hi, acc0);
            }
            if (two) {
                const uint32_t a0 = sp1[2 * hh], a1 = sp1[2 * hh + 1], a2 = sp1[2 * hh + 2];
                const uint32_t lo = __funnelshift_r(a0, a1, sh), hi = __funnelshift_r(a1, a2, sh);
                acc1 = __dp2a_lo(q.x, lo, acc1); acc1 = __dp2a_hi(q.y, lo, acc1);
                acc1 = __dp2a_lo(q.z, hi, acc1); acc1 = __dp2a_hi(q.w, hi, acc1);
            }
        }
        float *d = s_t1 + rp * rwp + sb + x;
        d[0] = (float)acc0 * (1.f / 65536.f);
        if (two) d[nrp * rwp] = (float)acc1 * (1.f / 65536.f);
        x += dx; rp += dr;
        if (x >= rw) { x -= rw; ++rp; }
    }
}

template <int VARIANT>
__global__ void __launch_bounds__(kFlexThreads + 32, 2) k_observe_flexible_v3(const __grid_constant__ DevPlan p,
                                                                          const uint8_t *__restrict__ ring,
                                                                          const int32_t *__restrict__ head,
                                                                          const double *__restrict__ action,
                                                                          const int32_t *__restrict__ atype,
                                                                          const uint8_t *__restrict__ ctrl,
                                                                          int32_t *__restrict__ loc, int32_t *__restrict__ res,
                                                                          int oh, int ow, int t1_cap, uint8_t *__restrict__ out,
                                                                          int *__restrict__ counters) {
    extern __shared__ __align__(16) uint8_t smem[];
    constexpr int kWin = 4;
    // per claimed env: index, window, ring head, and where its operators live in the pool
    __shared__ int s_en[kWin], s_er[kWin], s_ec[kWin], s_erh[kWin], s_erw[kWin], s_ehd[kWin];
    __shared__ int s_eth[kWin], s_ehw[kWin], s_ehx[kWin], s_enh[kWin], s_eqw[kWin], s_eqx[kWin];
    __shared__ uint32_t s_magic[256];              // FastDiv multipliers of 1..255: a division per divisor and env otherwise
    const int tid = threadIdx.x;
    if (tid < 256) s_magic[tid] = tid ? 0xFFFFFFFFu / (uint32_t)tid + 1u : 0u;
    const bool worker = tid < kFlexThreads;        // warps 0 .. 7/11 compute, the last warp is the control warp
    const bool boss = tid == kFlexThreads;         // its lane 0: env claims, fov updates, TMA stores
    const int K = p.K, quads = p.S_w >> 2, xcap = quads + 2, plane4 = p.plane >> 2;
    const int tile_bytes = K * oh * ow;
    uint32_t *s_tile = reinterpret_cast<uint32_t *>(smem);                          // [K][oh][ow] bytes
    uint32_t *s_x = s_tile + (tile_bytes >> 2);                                     // [K * rh][nwx + 2] window words
    float *s_t1 = reinterpret_cast<float *>(smem + align16((size_t)tile_bytes + 4 * (size_t)K * p.S_h * xcap));  // [kg * rh][rwp]
    uint32_t *s_wq = reinterpret_cast<uint32_t *>(s_t1 + t1_cap);                   // [rw][halves][4]; t1_cap % 4 == 0
    uint64_t *s_wh2 = reinterpret_cast<uint64_t *>(s_wq + p.S_w * 8);               // [rh][th] {w, w}
    int32_t *s_xw = reinterpret_cast<int32_t *>(s_wh2 + p.S_h * p.blur_tmax);       // [rw] first tap
    int32_t *s_xh = s_xw + p.S_w;                                                   // [rh]
    const int N = p.N;

    // ---- boss: claim an env, load what its fov update needs, apply it (three steps, spread over an iteration)
    struct Pend { int n, mode, r, c, rh, rw, hd, t; double a0, a1; };
    // envs blockIdx.x and blockIdx.x + G are this CTA's without asking; the counter hands out the rest.  A claim
    // is only issued here: its value is looked at a barrier later, so nobody waits for the atomic's round trip
    bool more = true;
    const int n_static = 2 * (int)gridDim.x;
    auto claim = [&]() { return more ? n_static + atomicAdd(counters, 1) : N; };
    auto load_env = [&](int n, Pend &q) {
        q.n = n;
        if (n >= N) return;
        q.mode = ctrl ? ctrl[n] : AGYM_FOV_APPLY;
        q.r = loc[2 * n]; q.c = loc[2 * n + 1]; q.rh = res[2 * n]; q.rw = res[2 * n + 1];
        q.hd = head[n];
        q.a0 = action ? action[2 * n] : 0.0; q.a1 = action ? action[2 * n + 1] : 0.0;
        q.t = atype ? atype[n] : AGYM_ATYPE_FOV_LOC;
    };
    auto finish_env = [&](const Pend &q, int e) {
        s_en[e] = q.n;
        if (q.n >= N) return;
        int r = q.r, c = q.c, rh = q.rh, rw = q.rw;
        if (q.mode == AGYM_FOV_RESET) {
            r = p.init_r; c = p.init_c; rh = p.f_h; rw = p.f_w;
        } else if (q.mode == AGYM_FOV_APPLY) {
            if (q.t == AGYM_ATYPE_FOV_RES) {  // fov_res = action, then re-clamp loc (fov_env.py:322-324)
                rh = min(max((int)q.a0, 1), p.S_h);
                rw = min(max((int)q.a1, 1), p.S_w);
                r = clip_rint((double)r, 0.0, (double)(p.S_h - rh));
                c = clip_rint((double)c, 0.0, (double)(p.S_w - rw));
            } else {
                double v0 = q.a0, v1 = q.a1;
                if (p.relative) {
                    v0 = (double)(r + clip_rint(q.a0, p.lo, p.hi));
                    v1 = (double)(c + clip_rint(q.a1, p.lo, p.hi));
                }
                r = clip_rint(v0, 0.0, (double)(p.S_h - rh));
                c = clip_rint(v1, 0.0, (double)(p.S_w - rw));
            }
        }
        const int n = q.n;
        loc[2 * n] = r; loc[2 * n + 1] = c; res[2 * n] = rh; res[2 * n + 1] = rw;
        s_er[e] = r; s_ec[e] = c; s_erh[e] = rh; s_erw[e] = rw; s_ehd[e] = q.hd;
        if (rh > p.f_h) {
            const FlexEntry eh = p.flexh2[rh], ew = p.flexq[rw];
            s_eth[e] = eh.taps; s_ehw[e] = eh.w_off; s_ehx[e] = eh.xmin_off;
            s_enh[e] = ew.taps; s_eqw[e] = ew.w_off; s_eqx[e] = ew.xmin_off;
        }
    };
    // ---- workers: the K windows of entry e as aligned words, and its W operator, by cp.async (one group)
    auto prefetch = [&](int e) {
        const int n = s_en[e];
        if (n < N) {
            const int r0 = s_er[e], c0 = s_ec[e], rh = s_erh[e], rw = s_erw[e], h = s_ehd[e];
            const int wq0 = c0 >> 2, nwx = ((c0 + rw - 1) >> 2) - wq0 + 1, nwxp = nwx + 2;
            const uint32_t *src = reinterpret_cast<const uint32_t *>(ring) + (size_t)n * K * plane4 + r0 * quads + wq0;
            const FastDiv fd_w(nwx, s_magic), fd_h(rh, s_magic);
            const int nrows = K * rh, dr = fd_w.div(kFlexThreads), dw = kFlexThreads - dr * nwx;
            int row = fd_w.div(tid), w = tid - row * nwx;
            while (row < nrows) {
                const int k = fd_h.div(row), y = row - k * rh;
                int slot = h + 1 + k;
                slot -= slot >= K ? K : 0;
                cp_async4(s_x + row * nwxp + w, src + slot * plane4 + y * quads + w);
                w += dw; row += dr;
                if (w >= nwx) { w -= nwx; ++row; }
            }
            if (rh > p.f_h) {
                const int32_t *gq = p.pool_i + s_eqw[e], *gx = p.pool_i + s_eqx[e];
                for (int i = tid; i < rw * s_enh[e]; i += kFlexThreads) cp_async16(s_wq + 4 * i, gq + 4 * i);
                for (int i = tid; i < rw; i += kFlexThreads) cp_async4(s_xw + i, gx + i);
            }
        }
        cp_async_commit();
    };

    if (boss) {
        Pend q0, q1;
        load_env(min((int)blockIdx.x, N), q0);
        load_env(min((int)(blockIdx.x + gridDim.x), N), q1);
        finish_env(q0, 0);
        finish_env(q1, 1);
        more = q1.n < N;
    }
    __syncthreads();
    if (worker) prefetch(0);

    for (int j = 0;; ++j) {
        const int e = j & (kWin - 1), n = s_en[e];
        if (n >= N) break;
        int n2 = N;
        if (boss) n2 = claim();                  // the env two iterations ahead; the result is not needed before #1
        const int r0 = s_er[e], c0 = s_ec[e], rh = s_erh[e], rw = s_erw[e];
        const bool blur = rh > p.f_h;  // row dimension only (fov_env.py:286)
        const int oy = VARIANT == AGYM_OUT_MASK ? r0 : 0, ox = VARIANT == AGYM_OUT_MASK ? c0 : 0;
        const int cb = c0 & 3, sb = ox & 3;
        const int vw = min(rw, ow - ox);
        FlexGeom g;
        g.rh = rh; g.oy = oy; g.oh = oh; g.ow4 = ow >> 2; g.wlo = ox >> 2;
        g.vh = min(rh, oh - oy);
        g.nq = ((sb + vw - 1) >> 2) + 1;
        g.rwp = (sb + rw + 3) & ~3;
        g.m_first = word_mask(0, sb, sb + vw);
        g.m_last = word_mask(4 * (g.nq - 1), sb, sb + vw);
        const int nwxp = ((c0 + rw - 1) >> 2) - (c0 >> 2) + 3;
        int th = 1, nh = 1;
        if (blur) {  // this env's H operator; s_wh2 / s_xh were last read before the previous env's final barrier
            th = s_eth[e];
            nh = s_enh[e];
            if (worker) {
                const int32_t *gh = p.pool_i + s_ehw[e], *gx = p.pool_i + s_ehx[e];
                for (int i = tid; i < (rh * th + 1) >> 1; i += kFlexThreads) cp_async16(s_wh2 + 2 * i, gh + 4 * i);
                for (int i = tid; i < rh; i += kFlexThreads) cp_async4(s_xh + i, gx + i);
            }
        }
        cp_async_commit();
        cp_async_wait<1>();                    // this thread's part of the windows (and W operator) has landed
        if (boss) bulk_wait_read<0>();         // the previous tile has been read by the TMA store
        __syncthreads();                       // #1
        Pend pend;
        pend.n = N;
        if (boss) {
            if (n2 >= N) { n2 = N; more = false; }
            load_env(n2, pend);
        }
        if (worker) {   // zero frame (everything outside the window stays zero)
            const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
            for (int i = tid; i < (tile_bytes >> 4); i += kFlexThreads) reinterpret_cast<uint4 *>(s_tile)[i] = z4;
        }
        const int per = blur ? rh * g.rwp : rh * g.nq;
        const int kg = min(K, t1_cap / per);
        for (int k0 = 0; k0 < K; k0 += kg) {
            const int kc = min(kg, K - k0);
            const bool last = k0 + kc >= K;
            if (k0) __syncthreads();           // the previous group's H pass has read t1
            if (!worker) {
            } else if (blur) {
                if (nh == 1) flex_wpass<1>(s_x + k0 * rh * nwxp, nwxp, s_xw, s_wq, s_t1, g.rwp, sb, cb, rw, kc * rh, tid, s_magic);
                else flex_wpass<2>(s_x + k0 * rh * nwxp, nwxp, s_xw, s_wq, s_t1, g.rwp, sb, cb, rw, kc * rh, tid, s_magic);
            } else {
                // the window itself, bit exact: output words (bytes outside the window masked to zero) parked in t1
                const FastDiv fd_nq(g.nq, s_magic);
                const uint32_t sh = (uint32_t)(cb - sb) * 8u;
                uint32_t *t1w = reinterpret_cast<uint32_t *>(s_t1);
                for (int i = tid; i < kc * rh * g.nq; i += kFlexThreads) {
                    const int row = fd_nq.div(i), q = i - row * g.nq;
                    const uint32_t *sp = s_x + (k0 * rh + row) * nwxp + q;
                    uint32_t word = __funnelshift_r(sp[0], sp[1], sh);
                    if (q == 0) word &= g.m_first;
                    if (q == g.nq - 1) word &= g.m_last;
                    t1w[i] = word;
                }
            }
            if (last) cp_async_wait<0>();      // H operator
            __syncthreads();                   // #2: t1 complete; after the last group s_x / s_wq / s_xw are free
            if (last && boss) finish_env(pend, (j + 2) & (kWin - 1));  // read by the prefetch after the NEXT env's #2
            if (!worker) continue;
            if (last) prefetch((j + 1) & (kWin - 1));
            if (blur) {
                switch (th) {
                    case 3: flex_hpass<3>(g, s_t1, s_wh2, s_xh, s_tile, k0, kc, th, tid, s_magic); break;
                    case 4: flex_hpass<4>(g, s_t1, s_wh2, s_xh, s_tile, k0, kc, th, tid, s_magic); break;
                    case 5: flex_hpass<5>(g, s_t1, s_wh2, s_xh, s_tile, k0, kc, th, tid, s_magic); break;
                    case 6: flex_hpass<6>(g, s_t1, s_wh2, s_xh, s_tile, k0, kc, th, tid, s_magic); break;
                    default: flex_hpass<0>(g, s_t1, s_wh2, s_xh, s_tile, k0, kc, th, tid, s_magic); break;
                }
            } else {
                const FastDiv fd_nq(g.nq, s_magic), fd_rh(rh, s_magic);
                const uint32_t *t1w = reinterpret_cast<const uint32_t *>(s_t1);
                for (int i = tid; i < kc * rh * g.nq; i += kFlexThreads) {
                    const int row = fd_nq.div(i), q = i - row * g.nq;
                    const int kk = fd_rh.div(row), y = row - kk * rh;
                    if (y < g.vh) s_tile[((k0 + kk) * oh + oy + y) * g.ow4 + g.wlo + q] = t1w[i];
                }
            }
        }
        fence_async_smem();
        __syncthreads();                       // #3: tile complete
        if (boss) {
            bulk_s2g(out + (size_t)n * tile_bytes, s_tile, (uint32_t)tile_bytes);
            bulk_commit();
        }
    }
    cp_async_wait<0>();
    if (boss) {
        bulk_wait_read<0>();
        // the last CTA to leave re-arms the counters for the next launch
        __threadfence();
        if (atomicAdd(counters + 1, 1) == (int)gridDim.x - 1) {
            atomicExch(counters, 0);
            atomicExch(counters + 1, 0);
        }
    }
}

// ---------------------------------------------------------------------------- normalise
// The reference hands the agent float32(u) / 255 (atari_env.py:75, dmc_env.py:183).  Consumer-side
// convenience: u8 observations -> normalised f32 (bit-identical to the reference's value: IEEE
// division), f16 or bf16, 16 pixels per thread, 16-byte loads and stores.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
template <int DT>  // 0 f32, 1 f16, 2 bf16
__global__ void k_normalize(const uint4 *__restrict__ src, void *__restrict__ dst, size_t n_vec) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_vec; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = ld_stream128(src + i);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        float f[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) f[j] = __fdiv_rn((float)((w[j >> 2] >> (8 * (j & 3))) & 0xffu), 255.f);
        if (DT == 0) {
            float4 *o = reinterpret_cast<float4 *>(dst) + 4 * i;
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
        } else {
            uint32_t h[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (DT == 1) {
                    const __half2 t = __floats2half2_rn(f[2 * j], f[2 * j + 1]);
                    h[j] = *reinterpret_cast<const uint32_t *>(&t);
                } else {
                    const __nv_bfloat162 t = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
                    h[j] = *reinterpret_cast<const uint32_t *>(&t);
                }
            }
            uint4 *o = reinterpret_cast<uint4 *>(dst) + 2 * i;
            o[0] = make_uint4(h[0], h[1], h[2], h[3]);
            o[1] = make_uint4(h[4], h[5], h[6], h[7]);
        }
    }
}

// ------------------------------------------------------------------------------ synth
__global__ void k_synth(uint4 *dst, size_t n_vec, uint64_t seed) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_vec; i += (size_t)gridDim.x * blockDim.x) {
        uint64_t z = seed + i * 0x9E3779B97F4A7C15ull;
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            z += 0x9E3779B97F4A7C15ull;
            uint64_t x = z;
            x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
            x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
            x ^= x >> 31;
            o[2 * j] = (uint32_t)x; o[2 * j + 1] = (uint32_t)(x >> 32);
        }
        dst[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

template <typename F>
cudaError_t set_smem(F func, size_t bytes) {
    if (bytes > 48 * 1024)
        return cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    return cudaSuccess;
}

size_t a16(size_t v) { return (v + 15) & ~size_t(15); }

// AGYM_NO_STD=1 forces the table-driven peripheral kernel even for the standard geometry
const bool g_disable_std = getenv("AGYM_NO_STD") != nullptr;
// AGYM_NO_TMA=1 forces the non-persistent ingest kernel (A/B comparisons, debugging)
const bool g_disable_tma = getenv("AGYM_NO_TMA") != nullptr;
// AGYM_FLEX_OLD=1 forces the one-CTA-per-env flexible kernel (A/B comparisons)
const bool g_flex_old = getenv("AGYM_FLEX_OLD") != nullptr;
// AGYM_CROP_OLD=1 forces the byte-gather crop kernel (A/B comparisons)
const bool g_crop_old = getenv("AGYM_CROP_OLD") != nullptr;
// AGYM_INGEST_UNITS=n: units (shared-memory stages) per env of the TMA ingest kernel (tuning)
const int g_units = getenv("AGYM_INGEST_UNITS") ? atoi(getenv("AGYM_INGEST_UNITS")) : 0;

}  // namespace

// --------------------------------------------------------------------------- launchers
namespace {
// cuTensorMapEncodeTiled through the runtime's driver entry point (the library does not link libcuda)
using EncodeTiledFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                   const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void *f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            f = nullptr;
        return reinterpret_cast<EncodeTiledFn>(f);
    }();
    return fn;
}
// frames [N][raw_h][rowb bytes] as [N][raw_h / 5][5][rowb]; box = 2 rows of R/2 consecutive periods of one env
bool encode_period5(CUtensorMap *m, const uint8_t *frames, int rowb, int raw_h, int N, int R) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return false;
    CUtensorMapDataType dt = CU_TENSOR_MAP_DATA_TYPE_UINT8;
    int esz = 1;
    if (rowb > 256) { dt = CU_TENSOR_MAP_DATA_TYPE_UINT16; esz = 2; }
    if (rowb > 512) { dt = CU_TENSOR_MAP_DATA_TYPE_UINT32; esz = 4; }
    if (rowb % (16 * 1) != 0 || rowb / esz > 256 || R % 2 != 0 || R / 2 > 256) return false;
    const cuuint64_t dims[4] = {(cuuint64_t)(rowb / esz), 5, (cuuint64_t)(raw_h / 5), (cuuint64_t)N};
    const cuuint64_t strides[3] = {(cuuint64_t)rowb, (cuuint64_t)5 * rowb, (cuuint64_t)raw_h * rowb};
    const cuuint32_t box[4] = {(cuuint32_t)(rowb / esz), 2, (cuuint32_t)(R / 2), 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    return enc(m, dt, 4, const_cast<uint8_t *>(frames), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// AGYM_INGEST_STAGES=3: three stages in the TMA ingest ring (tuning)
const int g_stages = getenv("AGYM_INGEST_STAGES") ? atoi(getenv("AGYM_INGEST_STAGES")) : 0;
// AGYM_NO_TM=1: contiguous bulk copies instead of the strided tensor copies in the TMA ingest kernel (A/B)
const bool g_disable_tm = getenv("AGYM_NO_TM") != nullptr;
}  // namespace

cudaError_t launch_ingest_atari(const DevPlan &p, const uint8_t *fa, const uint8_t *fb, const uint8_t *flags,
                                uint8_t *ring, int32_t *head, float *pcache, cudaStream_t st) {
    size_t smem = a16(sizeof(int32_t) * 3 * (p.S_w + p.S_h)) + a16(2 * (size_t)2 * p.S_h * p.raw_w);
    if (pcache) smem += a16(p.plane) + sizeof(float) * p.S_h * p.p_w;
    cudaError_t e;
    // units per env: index into the plan's span table (gray: 3 units of 28 rows; RGB: 7 units of 12 rows = two rows
    // for each of the 6 row segments of the consumer warps, and three CTAs still fit an SM)
    // With the gap-free stages of the tensor-copy path two units of 42 rows fit three CTAs per SM for gray frames
    // (7 rows for each row segment, fewer per-unit prologues): 0.239 ms instead of 0.254 ms at three units.
    const int rowb = p.raw_w * p.raw_c;
    auto pick_units = [&](int want) {
        int ui = std::min(std::max(want, 1), 8);
        while (ui > 1 && p.tma_span_rows[ui - 1] == 0) --ui;
        if (p.tma_span_rows[ui - 1] == 0)
            for (ui = 8; ui > 1 && p.tma_span_rows[ui - 1] == 0;) --ui;
        return ui;
    };
    CUtensorMap tma, tmb;
    std::memset(&tma, 0, sizeof(tma));
    std::memset(&tmb, 0, sizeof(tmb));
    auto try_tm = [&](int units) {
        const int R = p.S_h / units;
        return p.tma_period5 && !g_disable_tm && p.S_h % units == 0 && R % 2 == 0 &&
               (reinterpret_cast<uintptr_t>(fa) & 15) == 0 && (reinterpret_cast<uintptr_t>(fb) & 15) == 0 &&
               encode_period5(&tma, fa, rowb, p.raw_h, p.N, R) && encode_period5(&tmb, fb, rowb, p.raw_h, p.N, R);
    };
    int ui = pick_units(g_units ? g_units : (p.raw_c == 3 ? 7 : 2));
    bool tm = p.tma_span_rows[ui - 1] > 0 && try_tm(ui);
    if (!tm && !g_units && p.raw_c == 1) {   // contiguous copies: three units of 28 rows
        ui = pick_units(3);
        tm = p.tma_span_rows[ui - 1] > 0 && try_tm(ui);
    }
    const bool tma_ok = p.raw_c == 1 || (p.raw_c == 3 && p.fast_ingest_rgb);
    if (p.fast_ingest && tma_ok && !g_disable_tma && p.tma_span_rows[ui - 1] > 0) {
        const int units = ui, span_rows = p.tma_span_rows[ui - 1];
        const int R = p.S_h / units;
        const size_t stage = tm ? a16(2 * (2 * (((size_t)R * rowb + 127) & ~size_t(127)) + 128)) : a16(2 * ((size_t)span_rows * rowb + 16));
        const bool std_geom = p.raw_w == 160 && p.S_w == 84;
        // stages of the shared-memory ring: the gap-free gray stages are small enough for three at 3 CTAs per SM
        const int ns = (tm && std_geom && (g_stages ? g_stages == 3 : (p.raw_c == 1 && units >= 3))) ? 3 : 2;
        size_t fs = (tm ? 128 : 0) + ns * stage + a16(p.plane + 16) + 16 * (size_t)p.S_h + 8 * (size_t)((units + 1) & ~1);
        if (pcache) fs += sizeof(float) * ((size_t)p.S_h * p.p_w + (size_t)p.p_w * 24 + (size_t)p.p_h * p.sq_h.taps + (size_t)p.p_h);
        int dev = 0, sms = 148, occ = 1;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
#define AGYM_LAUNCH_TMA(...)                                                                                        \
    {                                                                                                               \
        if ((e = set_smem(k_ingest_atari_tma<__VA_ARGS__>, fs)) != cudaSuccess) return e;                           \
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_ingest_atari_tma<__VA_ARGS__>, kIngestThreads, fs);   \
        k_ingest_atari_tma<__VA_ARGS__><<<std::min(p.N, sms * std::max(occ, 1)), kIngestThreads, fs, st>>>(         \
            p, fa, fb, flags, ring, head, pcache, units, span_rows, tma, tmb);                                      \
    }
        if (p.raw_c == 3) {
            if (std_geom && tm && ns == 3) AGYM_LAUNCH_TMA(480, 84, 3, true, 3)
            else if (std_geom && tm) AGYM_LAUNCH_TMA(480, 84, 3, true, 2)
            else if (std_geom) AGYM_LAUNCH_TMA(480, 84, 3, false, 2)
            else if (tm) AGYM_LAUNCH_TMA(0, 0, 3, true, 2)
            else AGYM_LAUNCH_TMA(0, 0, 3, false, 2)
        } else {
            if (std_geom && tm && ns == 3) AGYM_LAUNCH_TMA(160, 84, 1, true, 3)
            else if (std_geom && tm) AGYM_LAUNCH_TMA(160, 84, 1, true, 2)
            else if (std_geom) AGYM_LAUNCH_TMA(160, 84, 1, false, 2)
            else if (tm) AGYM_LAUNCH_TMA(0, 0, 1, true, 2)
            else AGYM_LAUNCH_TMA(0, 0, 1, false, 2)
        }
#undef AGYM_LAUNCH_TMA
        return cudaGetLastError();
    }
    if (p.fast_ingest) {
        size_t fs = a16((size_t)2 * p.S_h * p.raw_w + 16) + a16(p.plane + 16);
        if (pcache) fs += sizeof(float) * p.S_h * p.p_w;
        if (p.raw_c == 1) {
            if ((e = set_smem(k_ingest_atari_fast<1>, fs)) != cudaSuccess) return e;
            k_ingest_atari_fast<1><<<p.N, kThreads, fs, st>>>(p, fa, fb, flags, ring, head, pcache);
        } else {
            if ((e = set_smem(k_ingest_atari_fast<3>, fs)) != cudaSuccess) return e;
            k_ingest_atari_fast<3><<<p.N, kThreads, fs, st>>>(p, fa, fb, flags, ring, head, pcache);
        }
        return cudaGetLastError();
    }
    if (p.raw_c == 1) {
        if ((e = set_smem(k_ingest_atari<1>, smem)) != cudaSuccess) return e;
        k_ingest_atari<1><<<p.N, kThreads, smem, st>>>(p, fa, fb, flags, ring, head, pcache);
    } else {
        if ((e = set_smem(k_ingest_atari<3>, smem)) != cudaSuccess) return e;
        k_ingest_atari<3><<<p.N, kThreads, smem, st>>>(p, fa, fb, flags, ring, head, pcache);
    }
    return cudaGetLastError();
}

cudaError_t launch_ingest_dmc(const DevPlan &p, const uint8_t *f, const uint8_t *flags, uint8_t *ring, int32_t *head,
                              float *pcache, cudaStream_t st) {
    size_t smem = pcache ? a16(p.plane) + sizeof(float) * p.S_h * p.p_w : 0;
    cudaError_t e;
    if ((e = set_smem(k_ingest_dmc, smem)) != cudaSuccess) return e;
    k_ingest_dmc<<<p.N, kThreads, smem, st>>>(p, f, flags, ring, head, pcache);
    return cudaGetLastError();
}

cudaError_t launch_stack(const DevPlan &p, const uint8_t *ring, const int32_t *head, uint8_t *out, cudaStream_t st) {
    k_stack<<<p.N, kThreads, 0, st>>>(p, ring, head, out);
    return cudaGetLastError();
}

cudaError_t launch_observe_fixed(const DevPlan &p, const uint8_t *ring, const int32_t *head, const double *action,
                                 const uint8_t *ctrl, int32_t *loc, int variant, uint8_t *out, cudaStream_t st) {
    cudaError_t e;
    const int crop_nwx = (p.f_w + 2) / 4 + 1;  // aligned words that cover f_w bytes at any byte offset
    const size_t crop_smem = a16((size_t)(p.K * p.f_h * p.f_w / 4) * 8) + (size_t)kCropWarps * p.K * p.f_h * crop_nwx * 4;
    if (variant == AGYM_OUT_CROP && (p.K * p.f_h * p.f_w) % 4 == 0 && !g_disable_std && !g_crop_old &&
        crop_smem <= 64 * 1024 && p.K * p.f_h * crop_nwx * 4 + 4 < 65536) {
        if ((e = set_smem(k_observe_fixed_crop_v2, crop_smem)) != cudaSuccess) return e;
        k_observe_fixed_crop_v2<<<(p.N + kCropWarps - 1) / kCropWarps, kCropWarps * 32, crop_smem, st>>>(p, ring, head, action, ctrl, loc, out, crop_nwx);
    } else if (variant == AGYM_OUT_CROP && (p.K * p.f_h * p.f_w) % 4 == 0 && !g_disable_std) {
        k_observe_fixed_crop_warp<<<(p.N + kThreads / 32 - 1) / (kThreads / 32), kThreads, 0, st>>>(p, ring, head, action, ctrl, loc, out);
    } else if (variant == AGYM_OUT_CROP) {
        k_observe_fixed<AGYM_OUT_CROP><<<p.N, kThreads, 0, st>>>(p, ring, head, action, ctrl, loc, out);
    } else if (variant == AGYM_OUT_MASK) {
        k_observe_fixed<AGYM_OUT_MASK><<<p.N, kThreads, 0, st>>>(p, ring, head, action, ctrl, loc, out);
    } else {
        const size_t smem = sizeof(float) * ((size_t)p.f_h * p.f_w + (size_t)p.f_h * p.S_w);
        if ((e = set_smem(k_observe_fixed<AGYM_OUT_RESIZE_FULL>, smem)) != cudaSuccess) return e;
        k_observe_fixed<AGYM_OUT_RESIZE_FULL><<<p.N, kThreads, smem, st>>>(p, ring, head, action, ctrl, loc, out);
    }
    return cudaGetLastError();
}

cudaError_t launch_observe_peripheral(const DevPlan &p, const ExpandStd *ew, const uint8_t *ring, const int32_t *head,
                                      const float *pcache, const double *action, const uint8_t *ctrl, int32_t *loc,
                                      uint8_t *out, cudaStream_t st) {
    const size_t smem = a16(p.plane) + sizeof(float) * ((size_t)p.S_h * p.p_w + (size_t)p.p_h * p.p_w + (size_t)p.p_h * p.S_w);
    cudaError_t e;
    const int quads = p.S_w / 4;
    if (pcache && ew && ew->ok && (p.K == 4 || p.K == 3) && !g_disable_std && p.N < (1 << 30) &&
        (reinterpret_cast<uintptr_t>(out) & 15) == 0 && (reinterpret_cast<uintptr_t>(pcache) & 15) == 0) {  // TMA bulk copies
        const int nw_max = (p.f_w + 3) / 4 + 1;
        const size_t fov_words = ((size_t)p.K * p.f_h * nw_max + 3) & ~size_t(3);
        const size_t fs = 4 * (3 * ((size_t)p.K * 400 + fov_words) + 2 * (size_t)p.K * 1764);
        int dev = 0, sms = 148, occ = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
#define AGYM_LAUNCH_STD(KK, NW)                                                                                    \
    {                                                                                                              \
    if (fs <= 220 * 1024) {                                                                                        \
        const int threads = ((21 * 4 * KK + 31) / 32) * 32;                                                        \
        if ((e = set_smem(k_observe_peripheral_std<KK, NW>, fs)) != cudaSuccess) return e;                         \
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_observe_peripheral_std<KK, NW>, threads, fs);        \
        if (occ >= 1) {                                                                                            \
            k_observe_peripheral_std<KK, NW><<<std::min(p.N, sms * occ), threads, fs, st>>>(                       \
                p, *ew, ring, head, pcache, action, ctrl, loc, out);                                               \
            return cudaGetLastError();                                                                             \
        }                                                                                                          \
    }                                                                                                              \
    }
        if (p.K == 4 && nw_max == 9) AGYM_LAUNCH_STD(4, 9)
        else if (p.K == 4) AGYM_LAUNCH_STD(4, 0)
        else if (nw_max == 9) AGYM_LAUNCH_STD(3, 9)
        else AGYM_LAUNCH_STD(3, 0)
#undef AGYM_LAUNCH_STD
    }
    if (pcache && p.fast_expand && quads <= 64 && (p.p_h * p.p_w) % 4 == 0) {
        // rows per thread segment: a multiple of the expand pattern's period when S_h allows it
        // (84 -> 20 repeats every 21 rows), so the lanes of a warp switch source rows together
        int ysegs = std::max(1, 128 / quads);
        {
            int g = p.S_h, b2 = p.p_h;
            while (b2) { const int t = g % b2; g = b2; b2 = t; }   // gcd(S_h, p_h)
            const int period = p.S_h / g;
            while (ysegs > 1 && (p.S_h % ysegs != 0 || (p.S_h / ysegs) % period != 0)) --ysegs;
        }
        // phase B needs quads * ysegs threads, phase A at least one thread per column pair
        const int threads = ((std::max(quads * ysegs, 2 * quads) + 31) / 32) * 32;
        const int nw_max = (p.f_w + 3) / 4 + 1;
        const size_t buf_words = (((size_t)p.K * p.p_h * p.p_w + 3) & ~size_t(3)) + (((size_t)p.K * p.f_h * nw_max + 3) & ~size_t(3));
        const size_t fs = 4 * (2 * buf_words + (size_t)p.K * p.p_h * p.S_w) + 8 * ((size_t)p.S_h + 1);
        int dev = 0, sms = 148, occ = 1;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
#define AGYM_LAUNCH_PF(KG, PW)                                                                                     \
    {                                                                                                              \
        if ((e = set_smem(k_observe_peripheral_v2<KG, PW>, fs)) != cudaSuccess) return e;                          \
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_observe_peripheral_v2<KG, PW>, threads, fs);         \
        if (occ < 1) return cudaErrorInvalidConfiguration;                                                         \
        k_observe_peripheral_v2<KG, PW><<<std::min(p.N, sms * occ), threads, fs, st>>>(p, ring, head, pcache, action, \
                                                                                      ctrl, loc, out, ysegs);     \
    }
        if (p.K == 4 && p.plane == 7056) AGYM_LAUNCH_PF(4, 1764)
        else if (p.K == 3 && p.plane == 7056) AGYM_LAUNCH_PF(3, 1764)
        else if (p.K % 4 == 0) AGYM_LAUNCH_PF(4, 0)
        else if (p.K % 3 == 0) AGYM_LAUNCH_PF(3, 0)
        else if (p.K % 2 == 0) AGYM_LAUNCH_PF(2, 0)
        else AGYM_LAUNCH_PF(1, 0)
#undef AGYM_LAUNCH_PF
        return cudaGetLastError();
    }
    if (pcache) {
        if ((e = set_smem(k_observe_peripheral<true>, smem)) != cudaSuccess) return e;
        k_observe_peripheral<true><<<p.N, kThreads, smem, st>>>(p, ring, head, pcache, action, ctrl, loc, out);
    } else {
        if ((e = set_smem(k_observe_peripheral<false>, smem)) != cudaSuccess) return e;
        k_observe_peripheral<false><<<p.N, kThreads, smem, st>>>(p, ring, head, pcache, action, ctrl, loc, out);
    }
    return cudaGetLastError();
}

cudaError_t launch_observe_flexible(const DevPlan &p, const uint8_t *ring, const int32_t *head, const double *action,
                                    const int32_t *atype, const uint8_t *ctrl, int32_t *loc, int32_t *res, int variant,
                                    int pad_h, int pad_w, uint8_t *out, cudaStream_t st) {
    cudaError_t e;
    if (variant != AGYM_OUT_RESIZE_FULL && p.flexb && p.flexq && !g_disable_std && !g_flex_old) {
        // persistent kernel, 2 CTAs per SM: whatever the fixed buffers leave of ~113 KB goes to t1
        const int oh = variant == AGYM_OUT_CROP ? pad_h : p.S_h, ow = variant == AGYM_OUT_CROP ? pad_w : p.S_w;
        const size_t tile = (size_t)p.K * oh * ow;
        const size_t fixed = a16(tile + 4 * (size_t)p.K * p.S_h * (p.S_w / 4 + 2)) +
                             4 * ((size_t)p.S_w * 8 + 2 * (size_t)p.S_h * p.blur_tmax + p.S_w + p.S_h) + 16;
        const size_t budget = 114000;  // + static shared memory + 1 KB reserved per CTA: two CTAs per SM (233,472 B)
        const size_t one = (size_t)p.S_h * (p.S_w + 4), all = (size_t)p.K * one;   // floats: one / all K frames of the largest window
        const size_t room = budget > fixed ? ((budget - fixed) / 4) & ~size_t(3) : 0;
        const size_t t1_cap = std::min(all, room);
        if (tile % 16 == 0 && ow % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 && t1_cap >= one &&
            (size_t)p.K * p.f_h * (p.S_w / 4 + 1) <= t1_cap) {
            const size_t fs = fixed + 4 * t1_cap;
            int dev = 0, sms = 148;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            const int grid = std::min(p.N, 2 * sms);
            if (variant == AGYM_OUT_CROP) {
                if ((e = set_smem(k_observe_flexible_v3<AGYM_OUT_CROP>, fs)) != cudaSuccess) return e;
                k_observe_flexible_v3<AGYM_OUT_CROP><<<grid, kFlexThreads + 32, fs, st>>>(p, ring, head, action, atype, ctrl, loc, res, oh, ow, (int)t1_cap, out, p.flex_counters);
            } else {
                if ((e = set_smem(k_observe_flexible_v3<AGYM_OUT_MASK>, fs)) != cudaSuccess) return e;
                k_observe_flexible_v3<AGYM_OUT_MASK><<<grid, kFlexThreads + 32, fs, st>>>(p, ring, head, action, atype, ctrl, loc, res, oh, ow, (int)t1_cap, out, p.flex_counters);
            }
            return cudaGetLastError();
        }
    }
    if (variant != AGYM_OUT_RESIZE_FULL && p.flexb && !g_disable_std) {
        const int oh = variant == AGYM_OUT_CROP ? pad_h : p.S_h, ow = variant == AGYM_OUT_CROP ? pad_w : p.S_w;
        const size_t tile = (size_t)p.K * oh * ow;
        // t1: one padded window of any size (S_h x S_w), or all K frames of windows up to ~50 x 52
        const int t1_cap = std::max(p.S_h * ((p.S_w + 3) & ~3), std::min(p.K * 50 * 52, p.K * p.S_h * ((p.S_w + 3) & ~3)));
        const size_t fs = tile + 4 * ((size_t)p.K * p.S_h * (p.S_w / 4 + 1) + (size_t)t1_cap +
                                      (size_t)(p.S_w + p.S_h) * p.blur_tmax + p.S_w + p.S_h);
        if (tile % 16 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 && fs <= 200 * 1024) {
            if (variant == AGYM_OUT_CROP) {
                if ((e = set_smem(k_observe_flexible_fast<AGYM_OUT_CROP>, fs)) != cudaSuccess) return e;
                k_observe_flexible_fast<AGYM_OUT_CROP><<<p.N, kThreads, fs, st>>>(p, ring, head, action, atype, ctrl, loc, res, oh, ow, t1_cap, out);
            } else {
                if ((e = set_smem(k_observe_flexible_fast<AGYM_OUT_MASK>, fs)) != cudaSuccess) return e;
                k_observe_flexible_fast<AGYM_OUT_MASK><<<p.N, kThreads, fs, st>>>(p, ring, head, action, atype, ctrl, loc, res, oh, ow, t1_cap, out);
            }
            return cudaGetLastError();
        }
    }
    const size_t smem = sizeof(float) * 2 * (size_t)p.plane;
#define AGYM_LAUNCH_FLEX(V)                                                                              \
    if ((e = set_smem(k_observe_flexible<V>, smem)) != cudaSuccess) return e;                            \
    k_observe_flexible<V><<<p.N, kThreads, smem, st>>>(p, ring, head, action, atype, ctrl, loc, res, pad_h, pad_w, out);
    if (variant == AGYM_OUT_CROP) { AGYM_LAUNCH_FLEX(AGYM_OUT_CROP) }
    else if (variant == AGYM_OUT_MASK) { AGYM_LAUNCH_FLEX(AGYM_OUT_MASK) }
    else { AGYM_LAUNCH_FLEX(AGYM_OUT_RESIZE_FULL) }
#undef AGYM_LAUNCH_FLEX
    return cudaGetLastError();
}

cudaError_t launch_normalize(const uint8_t *src, size_t n, int dtype, void *dst, cudaStream_t st) {
    const size_t n_vec = n / 16;
    if (n_vec == 0) return cudaSuccess;
    const int blocks = (int)std::min<size_t>((n_vec + 255) / 256, 148 * 16);
    const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
    if (dtype == 0) k_normalize<0><<<blocks, 256, 0, st>>>(s4, dst, n_vec);
    else if (dtype == 1) k_normalize<1><<<blocks, 256, 0, st>>>(s4, dst, n_vec);
    else k_normalize<2><<<blocks, 256, 0, st>>>(s4, dst, n_vec);
    return cudaGetLastError();
}

cudaError_t launch_synth(uint8_t *dst, size_t n, uint64_t seed, cudaStream_t st) {
    const size_t n_vec = n / 16;
    if (n_vec == 0) return cudaSuccess;
    const int blocks = (int)((n_vec + 255) / 256 < 148 * 8 ? (n_vec + 255) / 256 : 148 * 8);
    k_synth<<<blocks, 256, 0, st>>>(reinterpret_cast<uint4 *>(dst), n_vec, seed);
    return cudaGetLastError();
}

}  // namespace agym
