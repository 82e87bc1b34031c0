// sm_100a kernel: Atari ingest for the STANDARD geometry only (gray 210x160 screens -> 84x84, the configuration of
// BASELINE configs[3] and of every ALE game): AtariEnv._get_state + the frame logic of _step / _reset
// (atari_env.py:73-75, 80-82, 91, 111-114, 121-133).  Same arithmetic as k_ingest_atari_tma (agym_ingest.cu) — both
// frames resized with cv2's 11-bit fixed-point INTER_LINEAR, then max, then pushed into the ring — but every stride,
// row offset and vertical weight is a compile-time constant (43 instead of 52 instructions per column pair and row,
// no per-row table lookups, the squeeze loops fully unrolled).  Measured at 16,384 envs: 0.236 -> 0.182 ms, i.e. the
// kernel now runs at the HBM roofline (1.11 GB of real DRAM traffic at 6.1 TB/s) instead of at the issue limit:
//
//   * the vertical weights of the 2.5x scale are (512, 1536) on even and (1536, 512) on odd output rows, so
//     ((b * (h >> 4)) >> 16) is h >> 11 for one tap and one IMAD.HI for the other;
//   * segment g of the 6 row segments (42 column pairs each) resizes the output rows b + 6 g of a 42-row unit
//     (b = 0, 2, 4, 1, 3, 5 in turn) and, last, one of rows 36 .. 41: all staged-row offsets are immediates;
//   * raw rows arrive by strided tensor copies (TMA, period-5 view, see agym_ingest.cu).  PITCH = 176 (opt-in,
//     AGYM_INGEST_STD=176) makes the box 176 bytes wide although a raw row has 160: the 16 bytes past the row end
//     are out of bounds for the tensor map and are zero-filled without being read.  With that pitch tap rows six
//     rows apart lie 8 banks apart (6 * 176 = 1056 B = 264 words), right behind the 40-word row of the neighbouring
//     segment, and a warp that straddles two segments reads ONE contiguous run of banks (1.02 instead of 1.63
//     shared-memory wavefronts per LDS).  At the HBM bound this buys nothing (0.189 vs 0.188 ms), so the dense
//     160-byte pitch (10 % less shared memory) is the default.
//
// The plan verifies on the host that the cv2 tables follow exactly this pattern (DevPlan::std_gray) before the kernel
// is ever selected; every other geometry keeps the table-driven kernels of agym_ingest.cu.
#include "agym_device.cuh"

namespace agym {

namespace {

constexpr int kStdThreads = kThreads + 32;  // 8 resize warps + the TMA producer warp
constexpr int kRawB = 160, kS = 84, kR = 42, kPairs = 42, kSegs = 6, kPlane = kS * kS;
constexpr int kPW = 20;                     // peripheral squeeze width / height handled by the cache path

template <int PITCH>
struct StdLayout {
    static constexpr int kBlk = (kR * PITCH + 127) & ~127;                 // one tensor-copy box: 21 periods x 2 rows
    static constexpr int kFrame = 2 * kBlk;                                // even-row block + odd-row block (the 128-byte rounding of a
                                                                           // block also absorbs the 4-byte over-read of the last column pair)
    static constexpr int kStage = 2 * kFrame;                              // both frames
    static constexpr int kBox = kR * PITCH;                                // bytes one tensor copy signals (fill included)
};

// output row (inside a 42-row unit) that segment g resizes in its r-th turn
__host__ __device__ constexpr int std_row_base(int r) { return r == 0 ? 0 : r == 1 ? 2 : r == 2 ? 4 : r == 3 ? 1 : r == 4 ? 3 : 5; }
__device__ __forceinline__ int std_last_row(int g) { return g < 3 ? 40 - 2 * g : 47 - 2 * g; }  // 40 38 36 41 39 37

template <int PITCH, bool PC, int NS>
__global__ void __launch_bounds__(kStdThreads, 3) k_ingest_gray_std(const __grid_constant__ DevPlan p,
                                                                    const uint8_t *__restrict__ flags,
                                                                    uint8_t *__restrict__ ring, int32_t *__restrict__ head,
                                                                    float *__restrict__ pcache,
                                                                    const __grid_constant__ CUtensorMap tma,
                                                                    const __grid_constant__ CUtensorMap tmb) {
    using L = StdLayout<PITCH>;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint8_t *smem = smem_raw + ((128u - (smem_u32(smem_raw) & 127u)) & 127u);   // tensor copies land on 128-byte lines
    __shared__ __align__(8) uint64_t full[NS], empty[NS];
    constexpr int kEnvWin = 32;
    __shared__ int s_envfl[kEnvWin], s_envhd[kEnvWin];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = p.N, K = p.K;
    uint8_t *stages = smem;
    uint8_t *s_frame = stages + NS * L::kStage;
    float *s_t1 = reinterpret_cast<float *>(s_frame + kPlane);                  // [84][20] W-pass result
    uint32_t *s_sqq = reinterpret_cast<uint32_t *>(s_t1 + (PC ? kS * kPW : 0)); // [20][8] fixed-point W weights
    float *s_sqh = reinterpret_cast<float *>(s_sqq + (PC ? kPW * 8 : 0));       // [20][taps] H-pass weights
    int32_t *s_sqx = reinterpret_cast<int32_t *>(s_sqh + (PC ? kPW * p.sq_h.taps : 0));  // [20] first source row

    pdl_launch_dependents();
    if (tid == 0) {
        for (int i = 0; i < NS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], kThreads / 32); }
        mbar_fence_init();
    }
    if (PC) {
        for (int i = tid; i < kPW * 8; i += kStdThreads) s_sqq[i] = __ldg(p.sqw_q + i);
        for (int i = tid; i < kPW * p.sq_h.taps; i += kStdThreads) s_sqh[i] = __ldg(p.sq_h.w + i);
        for (int i = tid; i < kPW; i += kStdThreads) s_sqx[i] = __ldg(p.sq_h.xmin + i);
    }
    __syncthreads();
    pdl_wait();   // frames, flags, ring, head, pcache: not before the previous kernel in the stream has finished

    // unit `it` of this CTA: env = blockIdx.x + (it / 2) * gridDim.x, part = it & 1 (output rows 42 part .. 42 part + 41)
    const int my_envs = (N - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int my_units = my_envs * 2;
    if (warp == kThreads / 32) {
        // ------------------------------------------------------------------ producer warp
        if (lane == 0) {
            int n = blockIdx.x, part = 0, st = 0, ph = 1;  // ph: parity of the empty-barrier phase to wait for
            for (int it = 0; it < my_units; ++it) {
                if (it >= NS) mbar_wait<true>(&empty[st], ph);
                const int fl = flags[n];
                const int nvalid = (fl & AGYM_FLAG_IDLE) ? 0 : __popc(fl & 3);
                uint8_t *dst = stages + st * L::kStage;
                mbar_expect_tx(&full[st], (uint32_t)(nvalid * 2 * L::kBox));
                const int m0 = (kR >> 1) * part;
                if (nvalid && (fl & AGYM_FLAG_FRAME_A)) {
                    tensor_g2s_4d(dst, &tma, &full[st], 0, 0, m0, n);
                    tensor_g2s_4d(dst + L::kBlk, &tma, &full[st], 0, 3, m0, n);
                }
                if (nvalid && (fl & AGYM_FLAG_FRAME_B)) {
                    tensor_g2s_4d(dst + L::kFrame, &tmb, &full[st], 0, 0, m0, n);
                    tensor_g2s_4d(dst + L::kFrame + L::kBlk, &tmb, &full[st], 0, 3, m0, n);
                }
                if (++part == 2) { part = 0; n += gridDim.x; }
                if (++st == NS) { st = 0; ph ^= 1; }
            }
        }
        return;
    }

    // ---------------------------------------------------------------------- consumer warps
    const int g = tid / kPairs, pi = tid - g * kPairs;
    const bool worker = g < kSegs;
    const int4 px = worker ? __ldg(p.cx_pair + pi) : make_int4(0, 0, 0, 0);   // {aligned byte offset, PRMT selector, coef(x0), coef(x0+1)}
    const uint32_t sel = (uint32_t)px.y, c0 = (uint32_t)px.z, c1 = (uint32_t)px.w;
    // staged-row offsets of this thread (relative to a stage): turns 0 .. 5 are immediates on `tb`, the last turn has its own
    const int tb = px.x + g * 6 * PITCH;
    const int y6 = std_last_row(worker ? g : 0);
    const int t6 = px.x + (y6 & 1) * L::kBlk + (y6 >> 1) * 2 * PITCH;
    // the 512-weighted tap is the upper raw row of an even output row and the lower one of an odd output row
    const int t6_lo = t6 + ((y6 & 1) ? PITCH : 0), t6_hi = t6 + ((y6 & 1) ? 0 : PITCH);
    const int ob = 2 * pi + g * 6 * kS, ob6 = 2 * pi + y6 * kS;
    constexpr uint32_t kW1536 = 1536u << 16;

    // the 4 source bytes {s0(x0), s1(x0), s0(x0+1), s1(x0+1)} of one raw row, as the IDP.2A operand
    auto tap4 = [&](const uint8_t *row) -> uint32_t {
        const uint32_t *w = reinterpret_cast<const uint32_t *>(row);
        return __byte_perm(w[0], w[1], sel);
    };
    // one output column pair of one row from both frames: lo = the 512-weighted raw row, hi = the 1536-weighted one
    auto row_both = [&](const uint8_t *lo, const uint8_t *hi, uint8_t *dst) {
        uint32_t m0 = 0u, m1 = 0u;
#pragma unroll
        for (int fr = 0; fr < 2; ++fr) {
            const uint32_t qa = tap4(lo + fr * L::kFrame), qb = tap4(hi + fr * L::kFrame);
            const uint32_t a0 = __dp2a_lo(c0, qa, 0u), a1 = __dp2a_hi(c1, qa, 0u);
            const uint32_t b0 = __dp2a_lo(c0, qb, 0u), b1 = __dp2a_hi(c1, qb, 0u);
            // (512 (a >> 4)) >> 16 == a >> 11; the max of the two frames is taken before the monotone (x + 2) >> 2,
            // and the sum cannot exceed 4 * 255 + 1, so cv2's saturate is a no-op
            m0 = max(m0, (a0 >> 11) + __umulhi(kW1536, b0 >> 4));
            m1 = max(m1, (a1 >> 11) + __umulhi(kW1536, b1 >> 4));
        }
        *reinterpret_cast<uint16_t *>(dst) = (uint16_t)(((m0 + 2u) >> 2) | (((m1 + 2u) >> 2) << 8));
    };
    // resets / early game-over: one frame or none
    auto row_some = [&](const uint8_t *lo, const uint8_t *hi, uint8_t *dst, int fl) {
        uint32_t m0 = 0u, m1 = 0u;
        for (int fr = 0; fr < 2; ++fr) {
            if (!(fl & (1 << fr))) continue;
            const uint32_t qa = tap4(lo + fr * L::kFrame), qb = tap4(hi + fr * L::kFrame);
            const uint32_t a0 = __dp2a_lo(c0, qa, 0u), a1 = __dp2a_hi(c1, qa, 0u);
            const uint32_t b0 = __dp2a_lo(c0, qb, 0u), b1 = __dp2a_hi(c1, qb, 0u);
            m0 = max(m0, (a0 >> 11) + __umulhi(kW1536, b0 >> 4));
            m1 = max(m1, (a1 >> 11) + __umulhi(kW1536, b1 >> 4));
        }
        *reinterpret_cast<uint16_t *>(dst) = (uint16_t)(((m0 + 2u) >> 2) | (((m1 + 2u) >> 2) << 8));
    };

    // squeeze along W (cache path): one thread = one output column, rows sq_y0 + 12 k
    const int sq_i = tid % kPW, sq_y0 = tid / kPW;
    const bool sq_worker = PC && sq_y0 < 12;
    const int sq_ofs = (sq_worker ? __ldg(p.sqw_ofs + sq_i).x : 0) + sq_y0 * kS;
    const uint4 *sqq4 = reinterpret_cast<const uint4 *>(s_sqq) + 2 * sq_i;

    int slot = 0, fl = 0;
    for (int it = 0, n = blockIdx.x, part = 0, st = 0, ph = 0, je = 0; it < my_units; ++it) {
        if (part == 0) {
            // flags / head of this CTA's next kEnvWin envs are fetched together into shared memory
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
        if (!idle && worker) {
            const uint8_t *sb = stages + st * L::kStage;
            uint8_t *of = s_frame + part * (kR * kS);
            if ((fl & 3) == 3) {  // both frames (the steady state)
#pragma unroll
                for (int r = 0; r < 6; ++r) {
                    const int b = std_row_base(r);
                    const int o = (b & 1) * L::kBlk + (b >> 1) * 2 * PITCH;
                    const uint8_t *t = sb + tb + o;
                    row_both(t + ((b & 1) ? PITCH : 0), t + ((b & 1) ? 0 : PITCH), of + ob + b * kS);
                }
                row_both(sb + t6_lo, sb + t6_hi, of + ob6);
            } else {
#pragma unroll 1
                for (int r = 0; r < 6; ++r) {
                    const int b = std_row_base(r);
                    const int o = (b & 1) * L::kBlk + (b >> 1) * 2 * PITCH;
                    const uint8_t *t = sb + tb + o;
                    row_some(t + ((b & 1) ? PITCH : 0), t + ((b & 1) ? 0 : PITCH), of + ob + b * kS, fl);
                }
                row_some(sb + t6_lo, sb + t6_hi, of + ob6, fl);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[st]);  // this warp is done with the stage
        if (++st == NS) { st = 0; ph ^= 1; }
        if (++part < 2) continue;
        part = 0;
        const int n_done = n;
        n += gridDim.x;
        if (idle) continue;
        consumer_sync();  // the whole 84x84 frame is in s_frame
        if (tid == 0) head[n_done] = slot;
        if (fl & AGYM_FLAG_HARD_RESET) {
            const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
            for (int k = 0; k < K; ++k) {
                if (k == slot) continue;
                uint4 *z = reinterpret_cast<uint4 *>(ring + ((size_t)n_done * K + k) * kPlane);
                for (int i = tid; i < kPlane / 16; i += kThreads) z[i] = z4;
                if (PC) {
                    float *zc = pcache + ((size_t)n_done * K + k) * (kPW * kPW);
                    for (int i = tid; i < kPW * kPW; i += kThreads) zc[i] = 0.f;
                }
            }
        }
        {
            uint4 *out4 = reinterpret_cast<uint4 *>(ring + ((size_t)n_done * K + slot) * kPlane);
            const uint4 *src4 = reinterpret_cast<const uint4 *>(s_frame);
            out4[tid] = src4[tid];
            if (tid + kThreads < kPlane / 16) out4[tid + kThreads] = src4[tid + kThreads];
        }
        if (PC) {
            // W pass in 16-bit fixed point: 8 IDP.2A over the aligned 16-byte window, exact integer sum
            if (sq_worker) {
                const uint4 qa = sqq4[0], qb = sqq4[1];
#pragma unroll
                for (int k = 0; k < 7; ++k) {
                    const uint32_t *src = reinterpret_cast<const uint32_t *>(s_frame + sq_ofs + k * 12 * kS);
                    uint32_t acc = __dp2a_lo(qa.x, src[0], 0u);
                    acc = __dp2a_hi(qa.y, src[0], acc);
                    acc = __dp2a_lo(qa.z, src[1], acc);
                    acc = __dp2a_hi(qa.w, src[1], acc);
                    acc = __dp2a_lo(qb.x, src[2], acc);
                    acc = __dp2a_hi(qb.y, src[2], acc);
                    acc = __dp2a_lo(qb.z, src[3], acc);
                    acc = __dp2a_hi(qb.w, src[3], acc);
                    s_t1[tid + k * 12 * kPW] = (float)acc * (1.f / 131072.f);
                }
            }
            consumer_sync();
            // H pass: out[i][j..j+3] = sum_t wh[i][t] * t1[xmin[i] + t][j..j+3]; 100 threads, one float4 each
            if (tid < kPW * (kPW / 4)) {
                float *dst = pcache + ((size_t)n_done * K + slot) * (kPW * kPW);
                const int taps = p.sq_h.taps;
                const int i = tid / (kPW / 4), j = 4 * (tid - i * (kPW / 4));
                const float *w = s_sqh + i * taps;
                const float *t = s_t1 + s_sqx[i] * kPW + j;
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int tt = 0; tt < taps; ++tt) {
                    const float4 v = *reinterpret_cast<const float4 *>(t + tt * kPW);
                    const float wt = w[tt];
                    acc.x = fmaf(wt, v.x, acc.x); acc.y = fmaf(wt, v.y, acc.y);
                    acc.z = fmaf(wt, v.z, acc.z); acc.w = fmaf(wt, v.w, acc.w);
                }
                *reinterpret_cast<float4 *>(dst + i * kPW + j) = acc;
            }
            // no barrier here: the H pass reads only s_t1, which the next env rewrites after its own 'frame complete'
            // barrier, and s_frame was last read before the barrier between the two passes
        } else {
            consumer_sync();  // s_frame is rewritten by the next env's first unit
        }
    }
}

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

// AGYM_TM_L2 = 0 | 64 | 128 | 256: L2 promotion of the tensor copies (tuning).  Measured at 16,384 envs: 64 B 0.182 ms,
// none 0.188, 128 B 0.189, 256 B 0.198 — the boxes are 320 contiguous bytes on 32-byte offsets, wider promotion over-reads.
const int g_tm_l2 = getenv("AGYM_TM_L2") ? atoi(getenv("AGYM_TM_L2")) : 64;

// gray frames [N][210][160] as [N][42 periods][5 rows][160 bytes]; box = `pitch` bytes (>= 160, the excess is
// out of bounds and zero-filled) x 2 rows x 21 periods of one env
bool encode_std(CUtensorMap *m, const uint8_t *frames, int N, int pitch) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return false;
    const cuuint64_t dims[4] = {(cuuint64_t)kRawB, 5, 42, (cuuint64_t)N};
    const cuuint64_t strides[3] = {(cuuint64_t)kRawB, (cuuint64_t)5 * kRawB, (cuuint64_t)210 * kRawB};
    const cuuint32_t box[4] = {(cuuint32_t)pitch, 2, (cuuint32_t)(kR / 2), 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUtensorMapL2promotion l2 = g_tm_l2 == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                                      : g_tm_l2 == 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                      : g_tm_l2 == 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                                                       : CU_TENSOR_MAP_L2_PROMOTION_L2_64B;
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, const_cast<uint8_t *>(frames), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// AGYM_INGEST_STD = 0: never use this kernel; 176: conflict-free 176-byte pitch (see the header); default 160
const int g_std_mode = getenv("AGYM_INGEST_STD") ? atoi(getenv("AGYM_INGEST_STD")) : 160;

template <int PITCH, bool PC>
cudaError_t launch_std(const DevPlan &p, const uint8_t *flags, uint8_t *ring, int32_t *head, float *pcache,
                       const CUtensorMap &tma, const CUtensorMap &tmb, cudaStream_t st) {
    using L = StdLayout<PITCH>;
    constexpr int NS = 2;
    size_t fs = 128 + (size_t)NS * L::kStage + kPlane;
    if (PC) fs += sizeof(float) * ((size_t)kS * kPW + (size_t)kPW * 8 + (size_t)kPW * p.sq_h.taps + kPW);
    fs = a16(fs);
    cudaError_t e;
    auto kern = k_ingest_gray_std<PITCH, PC, NS>;
    if ((e = set_smem(kern, fs)) != cudaSuccess) return e;
    int dev = 0, sms = 148, occ = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kStdThreads, fs);
    return launch_pdl(kern, dim3(std::min(p.N, sms * std::max(occ, 1))), dim3(kStdThreads), fs, st, p, flags, ring, head, pcache, tma, tmb);
}

}  // namespace

// Returns cudaErrorNotSupported when the call is not eligible (the caller then takes the table-driven kernels).
cudaError_t launch_ingest_gray_std(const DevPlan &p, const uint8_t *fa, const uint8_t *fb, const uint8_t *flags,
                                   uint8_t *ring, int32_t *head, float *pcache, cudaStream_t st) {
    if (!p.std_gray || g_std_mode == 0) return cudaErrorNotSupported;
    if (pcache && !(p.p_h == kPW && p.p_w == kPW && p.squeeze_q && p.sq_h.taps <= 16)) return cudaErrorNotSupported;
    if ((reinterpret_cast<uintptr_t>(fa) & 15) || (reinterpret_cast<uintptr_t>(fb) & 15) ||
        (reinterpret_cast<uintptr_t>(ring) & 15) || (pcache && (reinterpret_cast<uintptr_t>(pcache) & 15)))
        return cudaErrorNotSupported;
    const int pitch = g_std_mode == 176 ? 176 : 160;
    CUtensorMap tma, tmb;
    std::memset(&tma, 0, sizeof(tma));
    std::memset(&tmb, 0, sizeof(tmb));
    if (!encode_std(&tma, fa, p.N, pitch) || !encode_std(&tmb, fb, p.N, pitch)) return cudaErrorNotSupported;
    if (pitch == 160)
        return pcache ? launch_std<160, true>(p, flags, ring, head, pcache, tma, tmb, st)
                      : launch_std<160, false>(p, flags, ring, head, pcache, tma, tmb, st);
    return pcache ? launch_std<176, true>(p, flags, ring, head, pcache, tma, tmb, st)
                  : launch_std<176, false>(p, flags, ring, head, pcache, tma, tmb, st);
}

}  // namespace agym
