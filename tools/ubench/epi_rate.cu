// Micro-benchmarks for the epilogue building blocks (developer tool): cycles per warp instruction of
// cvt.rn.relu.f16x2.f32, FADD, and tcgen05.ld / tcgen05.st round trips, for 1..4 warps per SM sub-partition.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../tiny-nerf-pytorch_b200/csrc/tnerf_ptx.cuh"
using namespace tnerf::ptx;

__global__ void k_cvt(float* out, long long* cyc, int iters) {
    float a[16];
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
    uint32_t acc = 0;
    __syncthreads();
    const long long c0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; i += 2) { acc ^= pack_relu_h2(a[i], a[i + 1]); a[i] += 1.0f; }
    }
    const long long c1 = clock64();
    out[threadIdx.x] = __uint_as_float(acc) + a[0];
    if (threadIdx.x % 32 == 0) cyc[threadIdx.x / 32] = c1 - c0;
}
__global__ void k_fadd(float* out, long long* cyc, int iters) {
    float a[16];
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
    __syncthreads();
    const long long c0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] += 1.0f;
    }
    const long long c1 = clock64();
    float s = 0; for (int i = 0; i < 16; ++i) s += a[i];
    out[threadIdx.x] = s;
    if (threadIdx.x % 32 == 0) cyc[threadIdx.x / 32] = c1 - c0;
}
__global__ void k_tmem(float* out, long long* cyc, int iters, int mode) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) { tmem_alloc(smem_u32(&slot), 512); tmem_relinquish(); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t t = slot + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 128;
    uint32_t v[32];
    for (int i = 0; i < 32; ++i) v[i] = i;
    uint32_t h[16];
    for (int i = 0; i < 16; ++i) h[i] = i;
    tmem_st16(t, h); tmem_st16(t + 16, h); tc_wait_st();
    __syncthreads();
    const long long c0 = clock64();
    uint32_t acc = 0;
    for (int it = 0; it < iters; ++it) {
        if (mode == 0) { tmem_ld32(t, v); tc_wait_ld(); acc += v[0]; }                       // ld latency (dependent)
        else if (mode == 1) { tmem_ld32(t, v); tmem_ld32(t + 32, v); tmem_ld32(t + 64, v); tmem_ld32(t + 96, v); tc_wait_ld(); acc += v[0]; }  // 4 loads per wait
        else if (mode == 2) { h[0] = acc; tmem_st16(t, h); tc_wait_st(); acc += 1; }        // st round trip
        else { h[0] = acc; tmem_st16(t, h); tmem_st16(t + 16, h); tmem_st16(t + 32, h); tmem_st16(t + 48, h); tc_wait_st(); acc += 1; }
    }
    const long long c1 = clock64();
    out[threadIdx.x] = __uint_as_float(acc);
    if (threadIdx.x % 32 == 0) cyc[threadIdx.x / 32] = c1 - c0;
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(slot, 512);
}
int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 4096 * 4); cudaMallocManaged(&cyc, 64 * 8);
    const int iters = 1000;
    for (int nw : {4, 8, 16}) {
        k_cvt<<<1, nw * 32>>>(out, cyc, iters); cudaDeviceSynchronize();
        printf("cvt.rn.relu.f16x2.f32 (+1 FADD each): %2d warps/SM: %.2f cycles per (cvt+fadd) per warp\n", nw, (double)cyc[0] / (iters * 8));
        k_fadd<<<1, nw * 32>>>(out, cyc, iters); cudaDeviceSynchronize();
        printf("FADD                                : %2d warps/SM: %.2f cycles per FADD per warp\n", nw, (double)cyc[0] / (iters * 16));
    }
    for (int nw : {4, 8}) for (int mode = 0; mode < 4; ++mode) {
        k_tmem<<<1, nw * 32>>>(out, cyc, 200, mode); cudaDeviceSynchronize();
        const char* nm[] = {"tcgen05.ld x32 + wait", "4 x tcgen05.ld x32 + wait", "tcgen05.st x16 + wait", "4 x tcgen05.st x16 + wait"};
        printf("%-28s %2d warps/SM: %.1f cycles per iteration (err %d)\n", nm[mode], nw, (double)cyc[0] / 200, (int)cudaGetLastError());
    }
    return 0;
}
