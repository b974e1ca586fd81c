// Shared pieces of the fused training kernels (tnerf_train.cu: one 128-sample tile per CTA, streamed weights;
// tnerf_train2.cu: two independent 64-sample streams per CTA, resident weights).
#pragma once
#include "tnerf_fused.cuh"

namespace tnerf {

struct SlabMap {     // offsets (floats) inside one per-CTA slab, TMEM-native [col][128 rows]
    int dw0, dw1, dw2, dw3, dwh, db1, db3, hb, total;
};

struct TrainParams {
    RaySource rs;
    long long n_rays, n_tiles;
    int S, R, white, Kx, L, include_input;
    float near_, far_;
    const float* jitter;
    const float* target;       // MSE mode when non-NULL
    const float *gC, *gD, *gA; // upstream gradients otherwise
    float inv_denom, scale;
    const float* scale_dev;    // optional device-side loss scale (overrides scale)
    float *comp, *loss_sum, *slabs;
    const __half* image;
    SlabMap sm;
    // two-stream kernel only: fp32 biases added in the accumulator drains (layers 1, 3) and by the compositing warps (heads)
    const float *b1, *b3, *b_sigma, *b_rgb;
    long long* debug;          // optional clock64 phase stamps of CTA 0 (tools/trace_train.py)
    int bulk_reduce;           // two-stream kernel: add the CTA's gradients into ONE vector (slabs[0 .. sm.total)) with bulk async reductions
    int sync_streams;          // the two streams of a CTA keep the same tile phase (dW1 halves drained where they are produced); -1 = host default
};

__device__ __forceinline__ uint32_t pack_sat_h2(float a, float b) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
__device__ __forceinline__ uint32_t relu_mask(uint32_t h) {   // 0xFFFF per half where the post-ReLU activation is > 0
    const __half2 hv = *reinterpret_cast<const __half2*>(&h);
    return __hgt2_mask(hv, __float2half2_rn(0.f));
}
// composite forward + loss gradient + reverse scan for the rays of one tile (src/volume.py:18-44 and its
// backward).  stage[] holds (sigma, r, g, b) per sample on entry and
// (dL/dsigma, dL/dr, dL/dg, dL/db) on exit.
template <int NW>
__device__ __forceinline__ float composite_tile(const TrainParams& p, float4* st, const float* sz, long long ray0, int warp_q, int lane) {
    const int S = p.S;
    float loss_part = 0.f;
    for (int rr = warp_q; rr < p.R; rr += NW) {
        const long long ray = ray0 + rr;
        float4* s4 = st + rr * S;
        const float* zz = sz + rr * S;
        const int nchunk = (S + 31) >> 5;
        if (ray >= p.n_rays) {
            for (int i = lane; i < S; i += 32) s4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            continue;
        }
        float o[3], d[3];
        load_ray(p.rs, ray, o, d);
        const float dn = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        // forward
        float T_carry = 1.f, cr = 0.f, cg = 0.f, cb = 0.f, asum = 0.f;
        float T_chunk[4];
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
            if (ch < nchunk) {
                T_chunk[ch] = T_carry;
                const int i = ch * 32 + lane;
                const bool ok = i < S;
                float alpha = 0.f, q = 1.f;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ok) {
                    v = s4[i];
                    const float gap = ((i == S - 1) ? kLastDelta : (zz[i + 1] - zz[i])) * dn;
                    alpha = 1.f - expf(-v.x * gap);
                    q = 1.f - alpha + kEpsT;
                }
                float incl = q;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const float up = __shfl_up_sync(0xffffffffu, incl, off);
                    if (lane >= off) incl *= up;
                }
                float excl = __shfl_up_sync(0xffffffffu, incl, 1);
                if (lane == 0) excl = 1.f;
                const float w = alpha * (T_carry * excl);
                cr += w * v.y; cg += w * v.z; cb += w * v.w; asum += w;
                T_carry *= __shfl_sync(0xffffffffu, incl, 31);
            }
        }
        cr = warp_sum(cr); cg = warp_sum(cg); cb = warp_sum(cb); asum = warp_sum(asum);
        const float bg = p.white ? 1.f - asum : 0.f;
        const float C0 = cr + bg, C1 = cg + bg, C2 = cb + bg;
        float g0, g1, g2, gd = 0.f, ga = 0.f;
        if (p.target) {
            const float e0 = C0 - p.target[3 * ray], e1 = C1 - p.target[3 * ray + 1], e2 = C2 - p.target[3 * ray + 2];
            g0 = 2.f * e0 * p.inv_denom; g1 = 2.f * e1 * p.inv_denom; g2 = 2.f * e2 * p.inv_denom;
            if (lane == 0) loss_part += (e0 * e0 + e1 * e1 + e2 * e2) * p.inv_denom;
        } else {
            g0 = p.gC ? p.gC[3 * ray] : 0.f; g1 = p.gC ? p.gC[3 * ray + 1] : 0.f; g2 = p.gC ? p.gC[3 * ray + 2] : 0.f;
            gd = p.gD ? p.gD[ray] : 0.f; ga = p.gA ? p.gA[ray] : 0.f;
        }
        if (p.comp && lane == 0) { p.comp[3 * ray] = C0; p.comp[3 * ray + 1] = C1; p.comp[3 * ray + 2] = C2; }
        const float gconst = ga - (p.white ? (g0 + g1 + g2) : 0.f);
        // reverse scan
        float R_carry = 0.f;
#pragma unroll
        for (int ch = 3; ch >= 0; --ch) {
            if (ch < nchunk) {
                const int i = ch * 32 + lane;
                const bool ok = i < S;
                float zi = 0.f, e = 1.f, gap = 0.f;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ok) {
                    v = s4[i];
                    zi = zz[i];
                    gap = ((i == S - 1) ? kLastDelta : (zz[i + 1] - zi)) * dn;
                    e = expf(-v.x * gap);
                }
                const float alpha = 1.f - e, q = ok ? (1.f - alpha + kEpsT) : 1.f;
                const float g = ok ? (g0 * v.y + g1 * v.z + g2 * v.w + gd * zi + gconst) : 0.f;
                float incl = q;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const float up = __shfl_up_sync(0xffffffffu, incl, off);
                    if (lane >= off) incl *= up;
                }
                float excl = __shfl_up_sync(0xffffffffu, incl, 1);
                if (lane == 0) excl = 1.f;
                const float T = T_chunk[ch] * excl;
                float Aa = ok ? g * alpha : 0.f, Qq = q;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const float An = __shfl_down_sync(0xffffffffu, Aa, off);
                    const float Qn = __shfl_down_sync(0xffffffffu, Qq, off);
                    if (lane + off < 32) { Aa = fmaf(Qq, An, Aa); Qq *= Qn; }
                }
                const float Rprev = fmaf(Qq, R_carry, Aa);
                float Ri = __shfl_down_sync(0xffffffffu, Rprev, 1);
                if (lane == 31) Ri = R_carry;
                if (ok) {
                    const float w = alpha * T;
                    s4[i] = make_float4(T * (g - Ri) * gap * e, w * g0, w * g1, w * g2);
                }
                R_carry = __shfl_sync(0xffffffffu, Rprev, 0);
            }
        }
    }
    return loss_part;
}


// slab reduction (tnerf_train.cu)
struct ReduceArgs {
    const float* slabs;
    int n_slabs;
    SlabMap sm;
    int D, Kx;
    long long off_w[4], off_b[4], off_ws, off_bs, off_wc, off_bc;   // flat offsets of layers.{0..3}.{weight,bias}, sigma, rgb
    float inv_scale;
    const float* scale_dev;
    float* grads;
    int zero_after;            // clear the (single) slab after reading: it is the accumulation target of the next launch
};
int fused_train2(tnerf_handle* h, const FusedPlan& fp, const TrainParams& p, int Kx, int grid, cudaStream_t s);   // tnerf_train2.cu

}  // namespace tnerf
