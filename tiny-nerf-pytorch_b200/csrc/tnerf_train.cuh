// Shared pieces of the fused training step (tnerf_train2.cu: the kernel, two independent 64-sample streams per CTA, resident
// weights; tnerf_train.cu: host side and the gradient scatter).
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
    int unscale;               // bulk-reduction flush only: the flush divides by the loss scale, so the ONE gradient vector holds unscaled sums
                               // (the optimiser launch then gathers from it directly: no scatter kernel, tnerf_train_fwd_bwd with grads = NULL)
    const int* tile_order;     // optional permutation of the CTAs for the tile dealing (tnerf_set_tile_order); NULL = identity
    float* found;              // optional overflow flag (GradScaler's found_inf): set to 1 when a head gradient leaves the fp16-safe range or is not finite
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
    float* found;              // optional overflow flag: set when a reduced gradient is not finite
};
int fused_train2(tnerf_handle* h, const FusedPlan& fp, const TrainParams& p, int Kx, int grid, cudaStream_t s);   // tnerf_train2.cu

}  // namespace tnerf
