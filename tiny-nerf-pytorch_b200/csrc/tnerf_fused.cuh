// Shared pieces of the tensor-core fused kernels (forward: tnerf_fused.cu, train: tnerf_train.cu).
#pragma once
#include "tnerf_internal.cuh"
#include "tnerf_ptx.cuh"

namespace tnerf {
using namespace ptx;

// ================================================================================================
// Operand images.  Every fp16 operand matrix X(r, k) (r = M or N index, k = reduction index) is stored
// in the SWIZZLE_NONE canonical form  idx(r,k) = ((k/8)*R + r)*8 + k%8  (R rows): 8x8 core matrices
// of 128 contiguous bytes, K-major view: LBO = R*16 B, SBO = 128 B.  The same bytes read as the
// transposed operand (r' = k, k' = r) are the MN-major canonical form with LBO = 128 B, SBO = R*16 B,
// which is what the backward kernel uses for dgrad/wgrad without a second copy.
__host__ __device__ inline long long img_idx(int r, int k, int R) { return ((long long)(k >> 3) * R + r) * 8 + (k & 7); }

enum { SEG_ACT = 0, SEG_X = 1, SEG_ONES = 2 };

struct LayerPlan {
    uint32_t b_off;      // byte offset of this layer's B image inside the packed image
    uint32_t idesc;
    uint16_t N;
    uint8_t nseg;
    uint8_t seg_kind[3];
    uint8_t seg_steps[3];   // K=16 steps per segment
};

struct FusedPlan {
    int depth, D, Kx, L, include_input, H;
    int bias_in_x;          // x layers carry their bias in the pad column D of the encoding
    uint32_t image_bytes;
    LayerPlan layer[kMaxDepth + 1];   // hidden layers, then the head
};

struct FwdParams {
    RaySource rs;
    long long n_rays, n_units;
    int S, G, R, white;
    float near_, far_;
    const float* jitter;
    float *comp, *depth, *acc, *weights, *rays_d_out;
    const __half* image;
    long long* debug;          // optional clock64 phase stamps of CTA 0 (tools/trace_fwd.py)
    FusedPlan plan;
};

// ------------------------------------------------------------------------------------------------
// Fourier features of one point into 64 fp16 slots (packed pairs), column order of
// src/encoding.py:27-33 ([x, sin f0, cos f0, sin f1, ...], 3 axes each).  sin/cos of the base
// frequency use a 2-term Cody-Waite reduction (|x| < ~50) + minimax polynomials; higher octaves use
// the double-angle recurrence s' = 2sc, c' = (c-s)(c+s): max abs error 5e-5 at 2^9, below fp16 rounding.
__device__ __forceinline__ void sincos_small(float x, float& s, float& c) {
    const float n = rintf(x * 0.6366197723675814f);
    float r = fmaf(-n, 1.5707963705062866f, x);
    r = fmaf(-n, -4.371138828673793e-08f, r);
    const float r2 = r * r;
    const float sp = fmaf(fmaf(fmaf(-1.9515295891e-4f, r2, 8.3321608736e-3f), r2, -1.6666654611e-1f) * r2, r, r);
    const float cp = fmaf(fmaf(fmaf(2.443315711809948e-5f, r2, -1.388731625493765e-3f), r2, 4.166664568298827e-2f) * r2, r2,
                          fmaf(-0.5f, r2, 1.f));
    const int q = (int)n & 3;
    const float ss = (q & 1) ? cp : sp, cc = (q & 1) ? sp : cp;
    s = (q & 2) ? -ss : ss;
    c = ((q + 1) & 2) ? -cc : cc;
}

template <int KX, bool INC>
__device__ __forceinline__ void encode_point(const float p[3], int L, uint32_t (&pk)[KX / 2]) {
    float f[KX];
#pragma unroll
    for (int i = 0; i < KX; ++i) f[i] = 0.f;
    constexpr int base = INC ? 3 : 0;
    if (INC) {
#pragma unroll
        for (int a = 0; a < 3; ++a) f[a] = p[a];
    }
    float s[3], c[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) sincos_small(p[a], s[a], c[a]);
#pragma unroll
    for (int k = 0; k < 10; ++k) {
        const bool on = k < L;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            if (base + 6 * k + 3 + a < KX - 1) {
                f[base + 6 * k + a] = on ? s[a] : 0.f;
                f[base + 6 * k + 3 + a] = on ? c[a] : 0.f;
            }
            const float s2 = s[a] + s[a];
            const float cn = (c[a] - s[a]) * (c[a] + s[a]);
            s[a] = s2 * c[a];
            c[a] = cn;
        }
    }
    f[KX - 1] = 1.f;   // constant-1 pad column: carries the bias of the x-consuming layers
#pragma unroll
    for (int i = 0; i < KX / 2; ++i) pk[i] = pack_h2(f[2 * i], f[2 * i + 1]);
}


// ---- Fourier features, packed as they are produced (keeps the live set small) ---------------------
template <int KX, bool INC, bool ALL_ON>
__device__ __forceinline__ void encode_stream_impl(const float p[3], int L, uint32_t (&pk)[KX / 2]) {
    constexpr int base = INC ? 3 : 0;
    float pend = 0.f;
    auto put = [&](int i, float v) {
        if (i & 1) pk[i >> 1] = pack_h2(pend, v); else pend = v;
    };
    if (INC) {
#pragma unroll
        for (int a = 0; a < 3; ++a) put(a, p[a]);
    }
    float s[3], c[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) sincos_small(p[a], s[a], c[a]);
    int next = base;
#pragma unroll
    for (int k = 0; k < 10; ++k) {
        if (base + 6 * k + 5 < KX - 1) {
            const bool on = ALL_ON || k < L;
#pragma unroll
            for (int a = 0; a < 3; ++a) put(base + 6 * k + a, on ? s[a] : 0.f);
#pragma unroll
            for (int a = 0; a < 3; ++a) put(base + 6 * k + 3 + a, on ? c[a] : 0.f);
            next = base + 6 * k + 6;
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const float s2 = s[a] + s[a];
                const float cn = (c[a] - s[a]) * (c[a] + s[a]);
                s[a] = s2 * c[a];
                c[a] = cn;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < KX; ++i)
        if (i >= next && i < KX - 1) put(i, 0.f);
    put(KX - 1, 1.f);          // constant-1 column: bias of the layers that consume the encoding
}
// number of octaves that fit in KX-1 feature columns; when L covers all of them the per-octave on/off selects disappear
template <int KX, bool INC>
__device__ __forceinline__ void encode_stream(const float p[3], int L, uint32_t (&pk)[KX / 2]) {
    constexpr int fit = (KX - 1 - (INC ? 3 : 0)) / 6;
    if (L >= (fit < 10 ? fit : 10)) encode_stream_impl<KX, INC, true>(p, L, pk);
    else encode_stream_impl<KX, INC, false>(p, L, pk);
}

// ------------------------------------------------------------------------------------------------
// In-place refresh of the fp16 operand image by the optimiser: parameter i of the flat fp32 vector (state_dict order)
// owns one image element (weights, biases carried by the encoding's constant-1 column) or two (biases carried by the
// "ones" K-step as an fp16 hi/lo pair).  Same conventions as pack_weights_kernel (tnerf_fused.cu); padding never changes.
enum { RP_W_HID = 0, RP_B_HID = 1, RP_W_SIG = 2, RP_B_SIG = 3, RP_W_RGB = 4, RP_B_RGB = 5 };
struct RepackMap {
    int n_tensors, H, valid;
    long long off[2 * kMaxDepth + 5];          // flat offset of every tensor (+ total)
    uint8_t kind[2 * kMaxDepth + 4], layer[2 * kMaxDepth + 4];
    int fan[kMaxDepth + 1], act_len[kMaxDepth + 1], xstart[kMaxDepth + 1], bias_k[kMaxDepth + 1], bias_hilo[kMaxDepth + 1], N[kMaxDepth + 1];
    uint32_t img_off[kMaxDepth + 1];           // in halfs
    __half* image;
};
__device__ __forceinline__ void repack_param(const RepackMap& mp, long long i, float val) {
    int t = 0;
    while (t + 1 < mp.n_tensors && mp.off[t + 1] <= i) ++t;
    const int loc = (int)(i - mp.off[t]), kind = mp.kind[t], l = mp.layer[t], N = mp.N[l];
    __half* img = mp.image + mp.img_off[l];
    int n, kk;
    bool bias = false;
    if (kind == RP_W_HID) { n = loc / mp.fan[l]; const int j = loc - n * mp.fan[l]; kk = j < mp.act_len[l] ? j : mp.xstart[l] + (j - mp.act_len[l]); }
    else if (kind == RP_W_SIG) { n = 0; kk = loc; }
    else if (kind == RP_W_RGB) { n = 1 + loc / mp.H; kk = loc % mp.H; }
    else { bias = true; n = (kind == RP_B_HID) ? loc : (kind == RP_B_SIG ? 0 : 1 + loc); kk = mp.bias_k[l]; }
    const __half hi = __float2half_rn(val);
    img[img_idx(n, kk, N)] = hi;
    if (bias && mp.bias_hilo[l]) img[img_idx(n, kk + 1, N)] = __float2half_rn(val - __half2float(hi));
}
bool build_repack_map(const tnerf_handle* h, RepackMap& mp);
// GradScaler semantics of the optimiser kernels (src/train.py:81,126-128), all on the device.  state (16 floats, 64-byte aligned):
// [0] loss scale, [1] clean steps since the scale last changed, [2] optimiser steps applied, [8..15] as four doubles: beta1^t and
// beta2^t, one pair per call parity (a kernel reads pair call&1 and writes the other, so blocks never race with the updater).
struct ScalerArgs {
    float* state;              // NULL: no scaler -- the step is always applied, bias corrections come from the host's step count
    const float* found;        // this call's overflow flag (non-zero: skip the step); the exchange kernel reads it from the vectors
    float* clear;              // flag to clear for the NEXT call (the other parity slot)
    float growth, backoff;
    int interval;
    unsigned int call;
};
int launch_adam_fused(float* p, float* g, float* m, float* v, long long n, long long n_clear, int step, float lr, float b1, float b2,
                      float eps, float* tail_out, const RepackMap& mp, const ScalerArgs& sc, cudaStream_t s);
int launch_adam_gather(float* p, float* g, float* m, float* v, long long n, long long n_clear, int step, float lr, float b1, float b2,
                       float eps, float* tail_out, const RepackMap& mp, const ScalerArgs& sc, const GatherPlan& plan, float* gsum, cudaStream_t s);
int launch_allreduce_adam(float* p, float* m, float* v, long long n, const float* const* peer_grads, unsigned int* const* peer_flags,
                          int world, int rank, unsigned int epoch, int step, float lr, float b1, float b2, float eps, float* reduced_out,
                          float* zero_next, const RepackMap& mp, const ScalerArgs& sc, long long timeout_cycles, cudaStream_t s);

int launch_allreduce_adam_gather(float* p, float* m, float* v, long long n, const float* const* peer_sums, unsigned int* const* peer_flags,
                                 int world, int rank, unsigned int epoch, int step, float lr, float b1, float b2, float eps, float* reduced_out,
                                 float* zero_next, const RepackMap& mp, const ScalerArgs& sc, const GatherPlan& plan, int sum_total,
                                 long long timeout_cycles, cudaStream_t s);
bool build_plan(const tnerf_handle* h, FusedPlan& pl);
int fused_render_fwd_fast(const FwdParams& p, int grid, cudaStream_t s);   // tnerf_fused_fast.cu (n_samples % 32 == 0)

// ---- unrolled MMA issue -------------------------------------------------------------------------
// A rolled issue loop costs ~120 cycles per tcgen05.mma (tools/umma_rate.py); unrolled, the issuer
// sustains the 64-cycle N=128 rate.  lo/hi are the two 32-bit words of the B (or A) shared-memory descriptor;
// only the start-address field of `lo` advances.
template <int STEPS>
__device__ __forceinline__ void issue_ts(uint32_t d, uint32_t a, uint32_t& b_lo, uint32_t b_hi, uint32_t b_adv, uint32_t idesc, uint32_t& acc) {
#pragma unroll
    for (int j = 0; j < STEPS; ++j)
        mma_ts(d, a + 8 * j, ((uint64_t)b_hi << 32) | (b_lo + j * b_adv), idesc, j == 0 ? acc : 1u);
    b_lo += STEPS * b_adv;
    acc = 1;
}
__device__ __forceinline__ void issue_ts_n(int steps, uint32_t d, uint32_t a, uint32_t& b_lo, uint32_t b_hi, uint32_t b_adv, uint32_t idesc, uint32_t& acc) {
    switch (steps) {
        case 0: break;
        case 1: issue_ts<1>(d, a, b_lo, b_hi, b_adv, idesc, acc); break;
        case 2: issue_ts<2>(d, a, b_lo, b_hi, b_adv, idesc, acc); break;
        case 3: issue_ts<3>(d, a, b_lo, b_hi, b_adv, idesc, acc); break;
        case 4: issue_ts<4>(d, a, b_lo, b_hi, b_adv, idesc, acc); break;
        case 8: issue_ts<8>(d, a, b_lo, b_hi, b_adv, idesc, acc); break;
        default:
            for (int j = 0; j < steps; ++j) issue_ts<1>(d, a + 8 * j, b_lo, b_hi, b_adv, idesc, acc);
    }
}
template <int STEPS>
__device__ __forceinline__ void issue_ss(uint32_t d, uint32_t& a_lo, uint32_t a_hi, uint32_t a_adv, uint32_t& b_lo, uint32_t b_hi, uint32_t b_adv,
                                         uint32_t idesc, uint32_t& acc) {
#pragma unroll
    for (int j = 0; j < STEPS; ++j)
        mma_ss(d, ((uint64_t)a_hi << 32) | (a_lo + j * a_adv), ((uint64_t)b_hi << 32) | (b_lo + j * b_adv), idesc, j == 0 ? acc : 1u);
    a_lo += STEPS * a_adv;
    b_lo += STEPS * b_adv;
    acc = 1;
}
__device__ __forceinline__ void issue_ss_n(int steps, uint32_t d, uint32_t& a_lo, uint32_t a_hi, uint32_t a_adv, uint32_t& b_lo, uint32_t b_hi,
                                           uint32_t b_adv, uint32_t idesc, uint32_t& acc) {
    switch (steps) {
        case 0: break;
        case 1: issue_ss<1>(d, a_lo, a_hi, a_adv, b_lo, b_hi, b_adv, idesc, acc); break;
        case 2: issue_ss<2>(d, a_lo, a_hi, a_adv, b_lo, b_hi, b_adv, idesc, acc); break;
        case 3: issue_ss<3>(d, a_lo, a_hi, a_adv, b_lo, b_hi, b_adv, idesc, acc); break;
        case 4: issue_ss<4>(d, a_lo, a_hi, a_adv, b_lo, b_hi, b_adv, idesc, acc); break;
        case 8: issue_ss<8>(d, a_lo, a_hi, a_adv, b_lo, b_hi, b_adv, idesc, acc); break;
        default:
            for (int j = 0; j < steps; ++j) issue_ss<1>(d, a_lo, a_hi, a_adv, b_lo, b_hi, b_adv, idesc, acc);
    }
}

}  // namespace tnerf
