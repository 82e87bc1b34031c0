// Host-side coefficient tables for the observation kernels.
//
// Everything a kernel multiplies a pixel by is computed here, on the host, once per plan,
// with the same scalar arithmetic the reference's libraries use, so that no device float op
// takes part in coefficient generation:
//   * Cv2Axis  — OpenCV's INTER_LINEAR 11-bit fixed-point coefficients, the resize behind
//                AtariEnv._get_state (atari_env.py:74).
//   * AaAxis   — ATen's antialiased-bilinear (triangle filter) weights, the resample behind
//                torchvision Resize in the foveal wrappers (fov_env.py:120,248,278,366-368).
#pragma once
#include <cstdint>
#include <vector>

namespace agym {

// One axis of cv2.resize(INTER_LINEAR) on 8-bit data.
struct Cv2Axis {
    std::vector<int32_t> s0, s1;   // the two source indices per destination index (clamped)
    std::vector<int32_t> coef;     // c0 | (c1 << 16), each an 11-bit fixed-point weight
};
// `zero_frac_at_border`: OpenCV zeroes the fraction at the borders on the x axis only.
Cv2Axis build_cv2_axis(int n_src, int n_dst, bool zero_frac_at_border);

// One axis of an antialiased bilinear resample n_in -> n_out.  Every destination index reads
// `taps` consecutive sources starting at xmin[i]; rows are zero-padded and, near the right
// border, shifted left so that xmin[i] + taps <= n_in always holds (no bounds checks on device).
struct AaAxis {
    int n_in = 0, n_out = 0, taps = 0;
    std::vector<int32_t> xmin;     // [n_out]
    std::vector<float> w;          // [n_out][taps]
};
// antialias = false: plain bilinear (ATen upsample_bilinear2d, align_corners=False) — what torchvision's Resize did on
// tensors before antialias=True became the default; two taps per destination index.
AaAxis build_aa_axis(int n_in, int n_out, bool antialias = true);

// The flexible fovea's blur along one axis, Resize(f) followed by Resize(r) (fov_env.py:276-280), as ONE
// banded r x r operator M = A(f -> r) * B(r -> f): the reference keeps floats between the two resamples,
// so the composition is the same linear map (weights multiplied and summed in double).
AaAxis build_blur_axis(int r, int f, bool antialias = true);

// The same operator in 16-bit fixed point for IDP.2A (k_observe_flexible_v3's W pass): per output index
// `halves` groups of 8 weights (taps 0-7, 8-15, ...) starting at xmin, scaled by 2^16 and rounded by largest
// remainder so that every row sums to 2^16 exactly (a lone weight of 1.0 is stored as 65535); two weights per
// 32-bit word, low half first.  Error against the float weights: < 2^-16 per tap, zero on constant images.
std::vector<uint32_t> quantize_axis_q16(const AaAxis &ax, int *halves);

}  // namespace agym
