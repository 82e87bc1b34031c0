// Host-side coefficient tables (see agym_tables.h).
#include "agym_tables.h"

#include <algorithm>
#include <cmath>

namespace agym {

// OpenCV resize.cpp, INTER_LINEAR, 8-bit source: per destination index d the source position
// is computed in double, narrowed to float, split into floor + fraction, and the fraction is
// turned into two 11-bit weights with cvRound (round half to even).
Cv2Axis build_cv2_axis(int n_src, int n_dst, bool zero_frac_at_border) {
    Cv2Axis ax;
    ax.s0.resize(n_dst); ax.s1.resize(n_dst); ax.coef.resize(n_dst);
    const double inv_scale = static_cast<double>(n_dst) / static_cast<double>(n_src);
    const double scale = 1.0 / inv_scale;
    for (int d = 0; d < n_dst; ++d) {
        float pos = static_cast<float>((d + 0.5) * scale - 0.5);
        int s = static_cast<int>(std::floor(pos));
        float frac = pos - static_cast<float>(s);
        if (zero_frac_at_border) {
            if (s < 0) { s = 0; frac = 0.f; }
            if (s >= n_src - 1) { s = n_src - 1; frac = 0.f; }
        }
        const int c0 = static_cast<int>(std::lrintf((1.f - frac) * 2048.f));
        const int c1 = static_cast<int>(std::lrintf(frac * 2048.f));
        ax.s0[d] = std::min(std::max(s, 0), n_src - 1);
        ax.s1[d] = std::min(std::max(s + 1, 0), n_src - 1);
        ax.coef[d] = (c0 & 0xffff) | (c1 << 16);
    }
    return ax;
}

// ATen UpSampleKernel.cpp, _compute_indices_min_size_weights_aa with the bilinear (triangle)
// filter, evaluated in double (the Atari reference path is float64).
AaAxis build_aa_axis(int n_in, int n_out, bool antialias) {
    AaAxis ax;
    ax.n_in = n_in; ax.n_out = n_out;
    ax.xmin.resize(n_out);
    if (n_in == n_out) {  // ATen skips a pass whose size does not change
        ax.taps = 1;
        ax.w.assign(n_out, 1.0f);
        for (int i = 0; i < n_out; ++i) ax.xmin[i] = i;
        return ax;
    }
    const double scale = static_cast<double>(n_in) / static_cast<double>(n_out);
    // without antialiasing the triangle filter keeps its unit support when downscaling: taps {floor(src), floor(src) + 1}
    // of src = scale * (i + 0.5) - 0.5 with weights {1 - frac, frac}, i.e. ATen's upsample_bilinear2d
    const double support = (antialias && scale >= 1.0) ? scale : 1.0;
    const double inv = (antialias && scale >= 1.0) ? 1.0 / scale : 1.0;
    const int max_taps = static_cast<int>(std::ceil(support)) * 2 + 1;
    std::vector<int> xsize(n_out);
    std::vector<double> wd(static_cast<size_t>(n_out) * max_taps, 0.0);
    int taps = 1;
    for (int i = 0; i < n_out; ++i) {
        const double center = scale * (i + 0.5);
        int lo = static_cast<int>(center - support + 0.5);
        lo = std::max(lo, 0);
        int hi = static_cast<int>(center + support + 0.5);
        hi = std::min(hi, n_in);
        int cnt = std::min(std::max(hi - lo, 0), max_taps);
        double total = 0.0;
        double *row = wd.data() + static_cast<size_t>(i) * max_taps;
        for (int j = 0; j < cnt; ++j) {
            const double x = std::fabs((j + lo - center + 0.5) * inv);
            row[j] = x < 1.0 ? 1.0 - x : 0.0;
            total += row[j];
        }
        if (total != 0.0)
            for (int j = 0; j < cnt; ++j) row[j] /= total;
        // drop exact-zero tails so the device loop is as short as the filter really is
        while (cnt > 1 && row[cnt - 1] == 0.0) --cnt;
        ax.xmin[i] = lo; xsize[i] = cnt;
        taps = std::max(taps, cnt);
    }
    taps = std::min(taps, n_in);
    ax.taps = taps;
    ax.w.assign(static_cast<size_t>(n_out) * taps, 0.f);
    for (int i = 0; i < n_out; ++i) {
        const double *row = wd.data() + static_cast<size_t>(i) * max_taps;
        int lo = ax.xmin[i];
        int shift = 0;  // keep lo + taps <= n_in: move the window left, pad zeros in front
        if (lo + taps > n_in) { shift = lo + taps - n_in; lo -= shift; }
        for (int j = 0; j < xsize[i] && j + shift < taps; ++j)
            ax.w[static_cast<size_t>(i) * taps + j + shift] = static_cast<float>(row[j]);
        ax.xmin[i] = lo;
    }
    return ax;
}

AaAxis build_blur_axis(int r, int f, bool antialias) {
    const AaAxis down = build_aa_axis(r, f, antialias), up = build_aa_axis(f, r, antialias);
    std::vector<double> dense(static_cast<size_t>(r) * r, 0.0);
    for (int y = 0; y < r; ++y)
        for (int a = 0; a < up.taps; ++a) {
            const double wa = up.w[static_cast<size_t>(y) * up.taps + a];
            if (wa == 0.0) continue;
            const int j = up.xmin[y] + a;
            for (int b = 0; b < down.taps; ++b)
                dense[static_cast<size_t>(y) * r + down.xmin[j] + b] += wa * down.w[static_cast<size_t>(j) * down.taps + b];
        }
    AaAxis ax;
    ax.n_in = r; ax.n_out = r;
    ax.xmin.resize(r);
    std::vector<int> last(r);
    int taps = 1;
    for (int y = 0; y < r; ++y) {
        int lo = 0, hi = r - 1;
        while (lo < r - 1 && dense[static_cast<size_t>(y) * r + lo] == 0.0) ++lo;
        while (hi > lo && dense[static_cast<size_t>(y) * r + hi] == 0.0) --hi;
        ax.xmin[y] = lo; last[y] = hi;
        taps = std::max(taps, hi - lo + 1);
    }
    ax.taps = taps;
    ax.w.assign(static_cast<size_t>(r) * taps, 0.f);
    for (int y = 0; y < r; ++y) {
        int lo = ax.xmin[y];
        if (lo + taps > r) lo = r - taps;  // keep the window inside: zeros in front
        for (int x = ax.xmin[y]; x <= last[y]; ++x)
            ax.w[static_cast<size_t>(y) * taps + (x - lo)] = static_cast<float>(dense[static_cast<size_t>(y) * r + x]);
        ax.xmin[y] = lo;
    }
    return ax;
}

std::vector<uint32_t> quantize_axis_q16(const AaAxis &ax, int *halves) {
    const int nh = (ax.taps + 7) / 8;
    if (halves) *halves = nh;
    std::vector<uint32_t> q(static_cast<size_t>(ax.n_out) * nh * 4, 0u);
    std::vector<int64_t> v(static_cast<size_t>(nh) * 8);
    std::vector<double> frac(static_cast<size_t>(nh) * 8);
    for (int x = 0; x < ax.n_out; ++x) {
        std::fill(v.begin(), v.end(), 0);
        std::fill(frac.begin(), frac.end(), 0.0);
        int64_t sum = 0;
        for (int t = 0; t < ax.taps; ++t) {
            const double w = static_cast<double>(ax.w[static_cast<size_t>(x) * ax.taps + t]) * 65536.0;
            v[t] = static_cast<int64_t>(std::floor(w));
            frac[t] = w - std::floor(w);
            sum += v[t];
        }
        for (int64_t left = 65536 - sum; left > 0; --left) {  // hand the missing units to the largest remainders
            int best = 0;
            for (int t = 1; t < ax.taps; ++t)
                if (frac[t] > frac[best]) best = t;
            if (frac[best] <= 0.0) break;
            ++v[best];
            frac[best] = -1.0;
        }
        for (int t = 0; t < nh * 8; ++t) v[t] = std::min<int64_t>(std::max<int64_t>(v[t], 0), 65535);
        for (int t = 0; t < nh * 8; t += 2)
            q[static_cast<size_t>(x) * nh * 4 + t / 2] = static_cast<uint32_t>(v[t]) | (static_cast<uint32_t>(v[t + 1]) << 16);
    }
    return q;
}

}  // namespace agym
