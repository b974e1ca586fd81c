// Shared device-side math of the TinyNeRF ray engine (sm_100a).
// Each helper cites the reference lines whose arithmetic it reproduces
// (paths relative to the reference repository avihaig/tiny-nerf-pytorch).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace tnerf {

constexpr int kMaxDepth = 8;
constexpr float kEpsT = 1e-10f;     // src/volume.py:30
constexpr float kLastDelta = 1e10f; // src/volume.py:20

struct RaySource {          // device-side mirror of tnerf_ray_source
    const float* rays_o;
    long long o_stride;
    const float* rays_d;
    const float* c2w;
    int H, W;
    float focal;
    const long long* pixel_index;
    long long first_ray;
    unsigned long long jitter_seed, jitter_step;   // seed != 0 and no jitter tensor: stratified jitter drawn in-kernel (jitter_uniform below)
    long long frame_rays;   // > 0: pose batch -- ray i belongs to pose i / frame_rays (c2w = [n_poses][16]), pixel first_ray + i % frame_rays
};

// t_i of torch.linspace(0,1,S) in fp32, bit for bit (src/sampling.py:16):
// step = fl(1/(S-1)); first half step*i, second half 1 - step*(S-1-i) with a single rounding.
__device__ __forceinline__ float linspace01(int i, int S) {
    if (S <= 1) return 0.f;
    const float step = __fdiv_rn(1.f, (float)(S - 1));
    return (i < S / 2) ? __fmul_rn(step, (float)i) : __fmaf_rn(-step, (float)(S - 1 - i), 1.f);
}

// z_i = near*(1-t) + far*t with separate roundings (src/sampling.py:17)
__device__ __forceinline__ float depth_bin(int i, int S, float near_, float far_) {
    const float t = linspace01(i, S);
    return __fadd_rn(__fmul_rn(near_, __fsub_rn(1.f, t)), __fmul_rn(far_, t));
}

// jittered depth: bins bounded by mid-points, first/last bin clamped (src/sampling.py:21-25)
__device__ __forceinline__ float depth_sample(int i, int S, float near_, float far_, float u, bool jittered) {
    const float zc = depth_bin(i, S, near_, far_);
    if (!jittered) return zc;
    const float lo = (i == 0) ? zc : __fmul_rn(0.5f, __fadd_rn(depth_bin(i - 1, S, near_, far_), zc));
    const float hi = (i == S - 1) ? zc : __fmul_rn(0.5f, __fadd_rn(zc, depth_bin(i + 1, S, near_, far_)));
    return __fadd_rn(lo, __fmul_rn(__fsub_rn(hi, lo), u));
}

// In-kernel stratified jitter (src/sampling.py:24 draws torch.rand_like(z_vals) on the device): counter-based Philox4x32-10 keyed by
// (seed, step) with the counter (sample, ray): u(seed, step, ray, sample) is a pure function, so the fused kernels, the fill kernel
// of tnerf_jitter_fill and any sharding of the rays see the same numbers.  Uniform in [0, 1) with 24 bits like torch.rand (fp32).
__device__ __forceinline__ uint32_t philox4x32_10_x(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return c0;
}
__device__ __forceinline__ float jitter_uniform(unsigned long long seed, unsigned long long step, long long ray, int sample) {
    const uint32_t x = philox4x32_10_x((uint32_t)sample, (uint32_t)ray, (uint32_t)((unsigned long long)ray >> 32), (uint32_t)step,
                                       (uint32_t)seed, (uint32_t)(seed >> 32) ^ (uint32_t)(step >> 32));
    return (float)(x >> 8) * 5.9604644775390625e-8f;      // 2^-24
}

// camera-space pixel direction rotated to world and normalised (src/rays.py:21-31)
__device__ __forceinline__ void pixel_ray(long long k, int H, int W, float focal, const float* __restrict__ c2w,
                                          float& dx, float& dy, float& dz) {
    const int col = (int)(k % W), row = (int)(k / W);
    const float cx = __fdiv_rn((float)col - (float)W * 0.5f, focal);
    const float cy = -__fdiv_rn((float)row - (float)H * 0.5f, focal);
    const float cz = -1.f;
    float wx = fmaf(cz, c2w[2], fmaf(cy, c2w[1], cx * c2w[0]));
    float wy = fmaf(cz, c2w[6], fmaf(cy, c2w[5], cx * c2w[4]));
    float wz = fmaf(cz, c2w[10], fmaf(cy, c2w[9], cx * c2w[8]));
    const float n = fmaxf(sqrtf(fmaf(wz, wz, fmaf(wy, wy, wx * wx))), 1e-12f);
    dx = __fdiv_rn(wx, n); dy = __fdiv_rn(wy, n); dz = __fdiv_rn(wz, n);
}

// origin + direction of ray i of a RaySource
__device__ __forceinline__ void load_ray(const RaySource& rs, long long i, float o[3], float d[3]) {
    if (rs.rays_d) {
        d[0] = rs.rays_d[3 * i]; d[1] = rs.rays_d[3 * i + 1]; d[2] = rs.rays_d[3 * i + 2];
        const float* po = rs.rays_o + rs.o_stride * i;
        o[0] = po[0]; o[1] = po[1]; o[2] = po[2];
    } else {
        const long long f = rs.frame_rays ? i / rs.frame_rays : 0;
        const float* cm = rs.c2w + 16 * f;
        const long long k = rs.pixel_index ? rs.pixel_index[i] : rs.first_ray + (i - f * rs.frame_rays);
        pixel_ray(k, rs.H, rs.W, rs.focal, cm, d[0], d[1], d[2]);
        o[0] = cm[3]; o[1] = cm[7]; o[2] = cm[11];
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace tnerf
