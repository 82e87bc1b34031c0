// sm_100a kernels: consumer-side normalisation and the synthetic frame generator.
#include "agym_device.cuh"

namespace agym {

namespace {

// ---------------------------------------------------------------------------- normalise
// The reference hands the agent float32(u) / 255 (atari_env.py:75, dmc_env.py:183).  Consumer-side
// convenience: u8 observations -> normalised f32 (bit-identical to the reference's value, the IEEE quotient:
// norm_u8 in agym_device.cuh, checked for all 256 inputs; a division per pixel made this kernel issue-bound at
// 2.3 TB/s), f16 or bf16, 16 pixels per thread, 16-byte loads and stores.
template <int DT>  // 0 f32, 1 f16, 2 bf16
__global__ void k_normalize(const uint4 *__restrict__ src, void *__restrict__ dst, size_t n_vec) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n_vec; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 v = ld_stream128(src + i);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
        float f[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) f[j] = norm_u8((w[j >> 2] >> (8 * (j & 3))) & 0xffu);   // == IEEE u / 255 for every u8 (tested)
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

// ----------------------------------------------------------------------------- record
// RecordWrapper's episode counters (fov_env.py:15-67: reset zeroes cumulative_reward / ep_len, step adds one
// step and info["raw_reward"]) and the fov_loc / fov_res trace of save_transition (fov_env.py:152-154, 160-163,
// 205-207, 216-219, 253-256, 262-265, 332-335, 351-354: appended on reset and on every step that is not done),
// for N envs at once.  One thread per env.
__global__ void k_record_step(int n, int is_reset, const double *__restrict__ raw_reward, const uint8_t *__restrict__ done,
                              const uint8_t *__restrict__ reset_mask, long long *__restrict__ ep_len,
                              double *__restrict__ cum_reward, const int32_t *__restrict__ loc,
                              const int32_t *__restrict__ res, int32_t *__restrict__ trace_row) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    long long len = ep_len[i];
    double cum = cum_reward[i];
    int valid;
    if (is_reset) {
        const bool mine = !reset_mask || reset_mask[i];
        if (mine) { len = 0; cum = 0.0; }
        valid = mine ? 1 : 0;
    } else {
        len += 1;
        cum += raw_reward ? raw_reward[i] : 0.0;
        valid = (done && done[i]) ? 0 : 1;
    }
    ep_len[i] = len;
    cum_reward[i] = cum;
    if (trace_row) {
        int32_t *t = trace_row + (size_t)i * 6;
        t[0] = loc ? loc[2 * i] : 0; t[1] = loc ? loc[2 * i + 1] : 0;
        t[2] = res ? res[2 * i] : 0; t[3] = res ? res[2 * i + 1] : 0;
        t[4] = (int32_t)len; t[5] = valid;
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


}  // namespace

// --------------------------------------------------------------------------- launchers
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

cudaError_t launch_record_step(int n, int is_reset, const double *raw_reward, const uint8_t *done, const uint8_t *reset_mask,
                               int64_t *ep_len, double *cum_reward, const int32_t *loc, const int32_t *res,
                               int32_t *trace_row, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    k_record_step<<<(n + 255) / 256, 256, 0, st>>>(n, is_reset, raw_reward, done, reset_mask,
                                                   reinterpret_cast<long long *>(ep_len), cum_reward, loc, res, trace_row);
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
