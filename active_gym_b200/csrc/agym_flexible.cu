// sm_100a kernels: flexible fovea.
#include "agym_device.cuh"

namespace agym {

namespace {

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
            const double a0 = action ? action[2 * n] : 0.0, a1 = action ? action[2 * n + 1] : 0.0;
            const int t = atype ? atype[n] : AGYM_ATYPE_FOV_LOC;
            if (t == AGYM_ATYPE_FOV_RES) {
                res_from_action(p, a0, a1, rh, rw);  // fov_res = action (fov_env.py:323)
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
            const double a0 = action ? action[2 * n] : 0.0, a1 = action ? action[2 * n + 1] : 0.0;
            const int t = atype ? atype[n] : AGYM_ATYPE_FOV_LOC;
            if (t == AGYM_ATYPE_FOV_RES) {  // fov_res = action, then re-clamp loc (fov_env.py:322-324)
                res_from_action(p, a0, a1, rh, rw);
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


struct FlexGeom {
    int rh, rwp, vh, nq, ow4, oh, wlo, oy;
    uint32_t m_first, m_last;  // byte masks of the first / last output word of a window row
};

// H pass over kc frames.  A thread OWNS an output word (window row y, word q): the row's weights, its first source row
// and the word's byte mask stay in registers while the thread walks through the kc frames (pointer increments only) —
// 25 instead of 70 instructions per output word for a 5-tap row.  s_wh2 holds every weight twice (FFMA2 operand).
template <int TH>
__device__ __forceinline__ void flex_hpass(const FlexGeom &g, const float *s_t1, const uint64_t *s_wh2, const int32_t *s_xh,
                                           uint32_t *s_tile, int k0, int kc, int th, int tid, const uint32_t *s_magic) {
    const FastDiv fd_nq(g.nq, s_magic);
    const int pairs = g.vh * g.nq;                        // rows past the frame's edge (y >= vh) produce nothing
    const int fstride = g.rh * g.rwp, tstride = g.oh * g.ow4;
    const uint64_t rne2 = pack2(12582912.f, 12582912.f);  // 1.5 * 2^23: v + bias has rint(v) (half to even) in its low byte
    for (int pr = tid; pr < pairs; pr += kFlexThreads) {
        const int y = fd_nq.div(pr), q = pr - y * g.nq;
        const float *src = s_t1 + s_xh[y] * g.rwp + 4 * q;
        uint32_t *dst = s_tile + (k0 * g.oh + g.oy + y) * g.ow4 + g.wlo + q;
        uint32_t mask = 0xffffffffu;
        if (q == 0) mask &= g.m_first;
        if (q == g.nq - 1) mask &= g.m_last;
        const uint64_t *w = s_wh2 + y * th;
        uint64_t wr[TH > 0 ? TH : 1];
        if (TH > 0) {
#pragma unroll
            for (int t = 0; t < TH; ++t) wr[t] = w[t];
        }
        for (int kk = 0; kk < kc; ++kk, src += fstride, dst += tstride) {
            uint64_t a01 = 0ull, a23 = 0ull;
            if (TH > 0) {
#pragma unroll
                for (int t = 0; t < TH; ++t) {
                    const ulonglong2 v = *reinterpret_cast<const ulonglong2 *>(src + t * g.rwp);
                    a01 = ffma2(v.x, wr[t], a01);
                    a23 = ffma2(v.y, wr[t], a23);
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
            *dst = __byte_perm(__byte_perm(b0, b1, 0x0040), __byte_perm(b2, b3, 0x0040), 0x5410) & mask;
        }
    }
}

// W pass over nrows staged rows: t1[row][sb + x] = 2^-16 * sum_t q[x][t] * X[row][xw[x] + t].  A thread OWNS output
// column x (its weights, first tap and shift stay in registers for the whole env) and walks down the rows of its row
// group, two rows per iteration (two independent IDP.2A chains): 28 instructions per two outputs instead of the 57 of
// the item-per-iteration form, whose index arithmetic and weight reloads were a quarter of the kernel (ncu source view).
// The staged rows are the full-width strips of the K frames (strip_w words apart, rows row_w words apart; xrow0 = the
// first window row of frame k0, c0 = the window's first column): row r of the group is row r % rh of frame r / rh.
template <int NH>
__device__ __forceinline__ void flex_wpass(const uint32_t *xrow0, int rh, int row_w, int strip_w, const int32_t *s_xw, const uint32_t *s_wq,
                                           float *s_t1, int rwp, int sb, int c0, int rw, int nrows, int tid, const uint32_t *s_magic) {
    const FastDiv fd_rw(rw, s_magic), fd_rh(rh, s_magic);
    const int kfix = strip_w - rh * row_w;
    const int G = fd_rw.div(kFlexThreads);           // row groups: kFlexThreads / rw >= 3 (rw <= 84)
    const int g = fd_rw.div(tid), x = tid - g * rw;
    if (g >= G) return;                              // the kFlexThreads - G * rw threads past the last group
    const int b = c0 + s_xw[x];
    const uint32_t sh = (uint32_t)(b & 3) * 8u;
    uint4 q[NH];
#pragma unroll
    for (int hh = 0; hh < NH; ++hh) q[hh] = reinterpret_cast<const uint4 *>(s_wq)[x * NH + hh];
    const uint32_t *sp0 = xrow0 + (b >> 2);
    float *d = s_t1 + g * rwp + sb + x;
    const int step_d = G * rwp;
    for (int r = g; r < nrows; r += 2 * G, d += 2 * step_d) {
        const bool two = r + G < nrows;
        const int r1 = two ? r + G : r;                  // the second row of the iteration (the first again past the end)
        const uint32_t *sp = sp0 + r * row_w + fd_rh.div(r) * kfix;
        const uint32_t *sp1 = sp0 + r1 * row_w + fd_rh.div(r1) * kfix;
        uint32_t acc0 = 0u, acc1 = 0u;
#pragma unroll
        for (int hh = 0; hh < NH; ++hh) {
            const uint32_t a0 = sp[2 * hh], a1 = sp[2 * hh + 1], a2 = sp[2 * hh + 2];
            const uint32_t c0 = sp1[2 * hh], c1 = sp1[2 * hh + 1], c2 = sp1[2 * hh + 2];
            const uint32_t lo0 = __funnelshift_r(a0, a1, sh), hi0 = __funnelshift_r(a1, a2, sh);
            const uint32_t lo1 = __funnelshift_r(c0, c1, sh), hi1 = __funnelshift_r(c1, c2, sh);
            acc0 = __dp2a_lo(q[hh].x, lo0, acc0); acc0 = __dp2a_hi(q[hh].y, lo0, acc0);
            acc1 = __dp2a_lo(q[hh].x, lo1, acc1); acc1 = __dp2a_hi(q[hh].y, lo1, acc1);
            acc0 = __dp2a_lo(q[hh].z, hi0, acc0); acc0 = __dp2a_hi(q[hh].w, hi0, acc0);
            acc1 = __dp2a_lo(q[hh].z, hi1, acc1); acc1 = __dp2a_hi(q[hh].w, hi1, acc1);
        }
        d[0] = (float)acc0 * (1.f / 65536.f);
        if (two) d[step_d] = (float)acc1 * (1.f / 65536.f);
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
    __shared__ __align__(8) uint64_t s_bar;        // completion of an env's K strip copies (phase = env count & 1)
    const int tid = threadIdx.x;
    if (tid < 256) s_magic[tid] = tid ? 0xFFFFFFFFu / (uint32_t)tid + 1u : 0u;
    if (tid == kFlexThreads) { mbar_init(&s_bar, 1); mbar_fence_init(); }
    const bool worker = tid < kFlexThreads;        // warps 0 .. 7/11 compute, the last warp is the control warp
    const bool boss = tid == kFlexThreads;         // its lane 0: env claims, fov updates, TMA stores
    const int K = p.K, quads = p.S_w >> 2, xcap = quads + 2, strip_w = (p.plane + 16) >> 2;
    const int tile_bytes = K * oh * ow;
    uint32_t *s_tile = reinterpret_cast<uint32_t *>(smem);                          // [K][oh][ow] bytes
    uint32_t *s_x = s_tile + (tile_bytes >> 2);                                     // [K][strip_w]: the windows' full-width row strips
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
        double a0 = q.a0, a1 = q.a1;
        // the control thread's fp64 chain must stay inside its branch: without this fence the compiler speculates the
        // ~60 fp64 instructions above `if (boss)` and every warp executes them for every env (ncu source view)
        asm volatile("" : "+d"(a0), "+d"(a1), "+r"(r), "+r"(c));
        if (q.mode == AGYM_FOV_RESET) {
            r = p.init_r; c = p.init_c; rh = p.f_h; rw = p.f_w;
        } else if (q.mode == AGYM_FOV_APPLY) {
            if (q.t == AGYM_ATYPE_FOV_RES) {  // fov_res = action, then re-clamp loc (fov_env.py:322-324)
                res_from_action(p, a0, a1, rh, rw);
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
        const int n = q.n;
        loc[2 * n] = r; loc[2 * n + 1] = c; res[2 * n] = rh; res[2 * n + 1] = rw;
        s_er[e] = r; s_ec[e] = c; s_erh[e] = rh; s_erw[e] = rw; s_ehd[e] = q.hd;
        if (rh > p.f_h) {
            const FlexEntry eh = p.flexh2[rh], ew = p.flexq[rw];
            s_eth[e] = eh.taps; s_ehw[e] = eh.w_off; s_ehx[e] = eh.xmin_off;
            s_enh[e] = ew.taps; s_eqw[e] = ew.w_off; s_eqx[e] = ew.xmin_off;
        }
    };
    // ---- boss: the K windows of entry e as the full-width strips of their rows, one TMA bulk copy per frame (a strip is
    // contiguous in the ring; its 16-byte-aligned hull starts `lead` bytes early, the same for every frame because a
    // plane is a multiple of 16 bytes).  No thread spends instructions on the windows; the strips complete on s_bar.
    auto fetch_strips = [&](int e) {
        const int n = s_en[e];
        if (n >= N) return;
        const int r0 = s_er[e], rh = s_erh[e], h = s_ehd[e];
        const uint32_t first = (uint32_t)(r0 * p.S_w), lead = first & 15u;
        const uint32_t bytes = (lead + (uint32_t)(rh * p.S_w) + 15u) & ~15u;
        const uint8_t *src = ring + (size_t)n * K * p.plane + (first - lead);
        mbar_expect_tx(&s_bar, bytes * (uint32_t)K);
        for (int k = 0; k < K; ++k) {
            int slot = h + 1 + k;
            slot -= slot >= K ? K : 0;
            bulk_g2s(s_x + k * strip_w, src + (size_t)slot * p.plane, bytes, &s_bar);
        }
    };
    // ---- workers: the W operator of entry e by cp.async (one group)
    auto prefetch = [&](int e) {
        const int n = s_en[e];
        if (n < N && s_erh[e] > p.f_h) {
            const int rw = s_erw[e];
            const int32_t *gq = p.pool_i + s_eqw[e], *gx = p.pool_i + s_eqx[e];
            for (int i = tid; i < rw * s_enh[e]; i += kFlexThreads) cp_async16(s_wq + 4 * i, gq + 4 * i);
            for (int i = tid; i < rw; i += kFlexThreads) cp_async4(s_xw + i, gx + i);
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
        fetch_strips(0);
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
        const uint32_t *s_xs = s_x + ((r0 * p.S_w & 15) >> 2);   // first window row of frame 0 (behind the hull's lead)
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
        cp_async_wait<1>();                    // this thread's part of the W operator has landed
        if (worker) mbar_wait(&s_bar, (uint32_t)j & 1u);   // the strips of this env (the j-th of this CTA) have landed
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
                if (nh == 1) flex_wpass<1>(s_xs + k0 * strip_w, rh, quads, strip_w, s_xw, s_wq, s_t1, g.rwp, sb, c0, rw, kc * rh, tid, s_magic);
                else flex_wpass<2>(s_xs + k0 * strip_w, rh, quads, strip_w, s_xw, s_wq, s_t1, g.rwp, sb, c0, rw, kc * rh, tid, s_magic);
            } else {
                // the window itself, bit exact: output words (bytes outside the window masked to zero) parked in t1
                const FastDiv fd_nq(g.nq, s_magic), fd_rh(rh, s_magic);
                const uint32_t sh = (uint32_t)(cb - sb) * 8u;
                uint32_t *t1w = reinterpret_cast<uint32_t *>(s_t1);
                const uint32_t *sp0 = s_xs + k0 * strip_w + (c0 >> 2);
                for (int i = tid; i < kc * rh * g.nq; i += kFlexThreads) {
                    const int row = fd_nq.div(i), q = i - row * g.nq;
                    const int kk = fd_rh.div(row), y = row - kk * rh;
                    const uint32_t *sp = sp0 + kk * strip_w + y * quads + q;
                    uint32_t word = __funnelshift_r(sp[0], sp[1], sh);
                    if (q == 0) word &= g.m_first;
                    if (q == g.nq - 1) word &= g.m_last;
                    t1w[i] = word;
                }
            }
            if (last) cp_async_wait<0>();      // H operator
            __syncthreads();                   // #2: t1 complete; after the last group s_x / s_wq / s_xw are free
            if (last && boss) {
                fetch_strips((j + 1) & (kWin - 1));                    // s_x is free: the next env's windows
                finish_env(pend, (j + 2) & (kWin - 1));                // read by the prefetch after the NEXT env's #2
            }
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
        if (p.norm_out) {   // uniform: the normalised copy of the finished tile; the tile is next written after barrier #1 of
                            // the following env, which comes after this loop in every thread's program order
            const uint4 *t4 = reinterpret_cast<const uint4 *>(s_tile);
            const int nv = tile_bytes >> 4;
            for (int i = threadIdx.x; i < nv; i += (int)blockDim.x) norm_store16(p.norm_dt, t4[i], p.norm_out, (size_t)n * nv + i);
        }
    }
    cp_async_wait<0>();
    if (boss) {
        bulk_wait_read<0>();
    }
}


}  // namespace

// --------------------------------------------------------------------------- launchers
cudaError_t launch_observe_flexible(const DevPlan &p, const uint8_t *ring, const int32_t *head, const double *action,
                                    const int32_t *atype, const uint8_t *ctrl, int32_t *loc, int32_t *res, int variant,
                                    int pad_h, int pad_w, uint8_t *out, int32_t *err, void *norm_out, int norm_dt,
                                    cudaStream_t st) {
    cudaError_t e;
    DevPlan q = p;
    q.err = err;
    q.norm_out = norm_out; q.norm_dt = norm_dt;
    const size_t out_bytes = (size_t)p.N * p.K * (variant == AGYM_OUT_CROP ? (size_t)pad_h * pad_w : (size_t)p.plane);
    auto finish = [&](cudaError_t r) {   // kernels without the fused normalised store: a separate pass over the u8 output
        if (r != cudaSuccess || !norm_out) return r;
        if (out_bytes % 16 != 0) return cudaErrorInvalidValue;
        return launch_normalize(out, out_bytes, norm_dt, norm_out, st);
    };
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
            p.plane % 16 == 0 && p.S_w % 4 == 0 && (reinterpret_cast<uintptr_t>(ring) & 15) == 0 &&   // strip copies
            (size_t)p.K * p.f_h * (p.S_w / 4 + 1) <= t1_cap) {
            const size_t fs = fixed + 4 * t1_cap;
            int dev = 0, sms = 148;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            const int grid = std::min(p.N, 2 * sms);
            // the env-claim counter starts every launch at zero (stream-ordered; a launch that faulted cannot leave it armed)
            if ((e = cudaMemsetAsync(p.flex_counters, 0, 16, st)) != cudaSuccess) return e;
            if (variant == AGYM_OUT_CROP) {
                if ((e = set_smem(k_observe_flexible_v3<AGYM_OUT_CROP>, fs)) != cudaSuccess) return e;
                k_observe_flexible_v3<AGYM_OUT_CROP><<<grid, kFlexThreads + 32, fs, st>>>(q, ring, head, action, atype, ctrl, loc, res, oh, ow, (int)t1_cap, out, p.flex_counters);
            } else {
                if ((e = set_smem(k_observe_flexible_v3<AGYM_OUT_MASK>, fs)) != cudaSuccess) return e;
                k_observe_flexible_v3<AGYM_OUT_MASK><<<grid, kFlexThreads + 32, fs, st>>>(q, ring, head, action, atype, ctrl, loc, res, oh, ow, (int)t1_cap, out, p.flex_counters);
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
                k_observe_flexible_fast<AGYM_OUT_CROP><<<p.N, kThreads, fs, st>>>(q, ring, head, action, atype, ctrl, loc, res, oh, ow, t1_cap, out);
            } else {
                if ((e = set_smem(k_observe_flexible_fast<AGYM_OUT_MASK>, fs)) != cudaSuccess) return e;
                k_observe_flexible_fast<AGYM_OUT_MASK><<<p.N, kThreads, fs, st>>>(q, ring, head, action, atype, ctrl, loc, res, oh, ow, t1_cap, out);
            }
            return finish(cudaGetLastError());
        }
    }
    const size_t smem = sizeof(float) * 2 * (size_t)p.plane;
#define AGYM_LAUNCH_FLEX(V)                                                                              \
    if ((e = set_smem(k_observe_flexible<V>, smem)) != cudaSuccess) return e;                            \
    k_observe_flexible<V><<<p.N, kThreads, smem, st>>>(q, ring, head, action, atype, ctrl, loc, res, pad_h, pad_w, out);
    if (variant == AGYM_OUT_CROP) { AGYM_LAUNCH_FLEX(AGYM_OUT_CROP) }
    else if (variant == AGYM_OUT_MASK) { AGYM_LAUNCH_FLEX(AGYM_OUT_MASK) }
    else { AGYM_LAUNCH_FLEX(AGYM_OUT_RESIZE_FULL) }
#undef AGYM_LAUNCH_FLEX
    return finish(cudaGetLastError());
}


}  // namespace agym
