// Tensor-core fused hot path (sm_100a): per 128-sample tile
//   rays -> stratified depths -> Fourier features (registers) -> MLP as tcgen05 GEMMs (activations in
//   tensor memory, weights resident in shared memory, fp32 accumulators in tensor memory) ->
//   sigma/rgb heads -> alpha-compositing scan.
// Replaces the reference sequence src/train.py:51-56 (render) without materialising points, encodings
// or activations in HBM.  See DESIGN.md sections 3-5 for layouts and the roofline of each kernel.
#include "tnerf_fused.cuh"

namespace tnerf {

// ------------------------------------------------------------------------------------------------
// weight packing: fp32 parameters -> fp16 operand image.  One thread per image element.
struct PackArgs {
    const float* W[kMaxDepth + 2];
    const float* b[kMaxDepth + 2];
    FusedPlan plan;
    int fan[kMaxDepth + 1];
    int skip_layer;      // index of the layer that consumes [h, x], or -1
};

__global__ void pack_weights_kernel(PackArgs a, __half* __restrict__ img) {
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long total = a.plan.image_bytes / 2;
    if (e >= total) return;
    const int nl = a.plan.depth + 1;
    int l = 0;
    while (l + 1 < nl && (long long)a.plan.layer[l + 1].b_off / 2 <= e) ++l;
    const LayerPlan& lp = a.plan.layer[l];
    const long long loc = e - lp.b_off / 2;
    const int N = lp.N;
    const int k = (int)((loc / 8) / N) * 8 + (int)(loc % 8);
    const int n = (int)((loc / 8) % N);
    // walk the K segments of this layer to find what forward-K index k means
    int kk = k;
    float v = 0.f;
    const bool head = (l == a.plan.depth);
    for (int sgi = 0; sgi < lp.nseg; ++sgi) {
        const int len = lp.seg_steps[sgi] * 16;
        if (kk < len) {
            const int kind = lp.seg_kind[sgi];
            if (head) {
                // rows: 0 = sigma, 1..3 = rgb, rest zero
                if (n < 4) {
                    const float* W = (n == 0) ? a.W[a.plan.depth] : a.W[a.plan.depth + 1] + (long long)(n - 1) * a.plan.H;
                    const float bb = (n == 0) ? a.b[a.plan.depth][0] : a.b[a.plan.depth + 1][n - 1];
                    if (kind == SEG_ACT) v = W[kk];
                    else if (kind == SEG_ONES) {
                        const float hi = __half2float(__float2half_rn(bb));
                        v = (kk == 0) ? hi : (kk == 1 ? bb - hi : 0.f);
                    }
                }
            } else {
                const float* W = a.W[l] + (long long)n * a.fan[l];
                const float bb = a.b[l][n];
                if (kind == SEG_ACT) v = W[kk];
                else if (kind == SEG_X) {
                    const int xo = (l == 0) ? 0 : a.plan.H;
                    if (kk < a.plan.D) v = W[xo + kk];
                    else if (kk == a.plan.Kx - 1 && a.plan.bias_in_x) v = bb;   // constant-1 column of the encoding
                } else {
                    const float hi = __half2float(__float2half_rn(bb));
                    v = (kk == 0) ? hi : (kk == 1 ? bb - hi : 0.f);
                }
            }
            break;
        }
        kk -= len;
    }
    img[e] = __float2half_rn(v);
}

constexpr int MAX_G = 8;       // tiles of 128 samples per work unit of the render kernels (lcm(n_samples, 128) rows)

// ================================================================================================
// UMMA self-test: D(128 x N) = A(128 x K) B(N x K)^T through the descriptor conventions above.
//   mode & 3: 0 = A,B K-major from smem; 1 = A from tensor memory; 2 = B MN-major; 3 = A MN-major
//   mode & 16: swap LBO/SBO (convention probe); mode & 32: swap fp16 halves in the TMEM A packing
__global__ void __launch_bounds__(128, 1) umma_selftest_kernel(const float* __restrict__ a, const float* __restrict__ b, int N, int K,
                                                              int mode, float* __restrict__ d) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __half* As = reinterpret_cast<__half*>(smem);
    __half* Bs = reinterpret_cast<__half*>(smem + 128 * K * 2);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int m = mode & 3;
    if (t == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    if (warp == 0) { tmem_alloc(smem_u32(&slot), 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    // stage operands
    for (int e = t; e < 128 * K; e += 128) {
        int mm, kk;
        if (m == 3) { kk = e / 128; mm = e % 128; } else { mm = e / K; kk = e % K; }
        const float v = a[e];
        if (m == 3) As[img_idx(kk, mm, K)] = __float2half_rn(v);       // MN-major: rows = k, "k" = m
        else As[img_idx(mm, kk, 128)] = __float2half_rn(v);
    }
    for (int e = t; e < N * K; e += 128) {
        int nn, kk;
        if (m == 2) { kk = e / N; nn = e % N; } else { nn = e / K; kk = e % K; }
        const float v = b[e];
        if (m == 2) Bs[img_idx(kk, nn, K)] = __float2half_rn(v);
        else Bs[img_idx(nn, kk, N)] = __float2half_rn(v);
    }
    if (m == 1) {
        const uint32_t tw = tmem + 256 + ((uint32_t)(warp * 32) << 16);
        for (int c0 = 0; c0 < K / 2; c0 += 8) {
            uint32_t r[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float lo = a[(warp * 32 + lane) * K + 2 * (c0 + i)], hi = a[(warp * 32 + lane) * K + 2 * (c0 + i) + 1];
                r[i] = (mode & 32) ? pack_h2(hi, lo) : pack_h2(lo, hi);
            }
            tmem_st8(tw + c0, r);
        }
        tc_wait_st();
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (t == 0) {
        const uint32_t idesc = make_idesc_f16(128, N, m == 3, m == 2);
        const uint32_t sa = smem_u32(As), sb = smem_u32(Bs);
        for (int j = 0; j < K / 16; ++j) {
            uint32_t a_lbo, a_sbo, a_adv, b_lbo, b_sbo, b_adv;
            if (m == 3) { a_lbo = 128; a_sbo = K * 16; a_adv = 256; } else { a_lbo = 128 * 16; a_sbo = 128; a_adv = 128 * 32; }
            if (m == 2) { b_lbo = 128; b_sbo = K * 16; b_adv = 256; } else { b_lbo = N * 16; b_sbo = 128; b_adv = N * 32; }
            if (mode & 16) { uint32_t x = a_lbo; a_lbo = a_sbo; a_sbo = x; x = b_lbo; b_lbo = b_sbo; b_sbo = x; }
            const uint64_t bd = make_desc(sb + j * b_adv, b_lbo, b_sbo);
            if (m == 1) mma_ts(tmem, tmem + 256 + j * 8, bd, idesc, j > 0);
            else mma_ss(tmem, make_desc(sa + j * a_adv, a_lbo, a_sbo), bd, idesc, j > 0);
        }
        tc_commit(smem_u32(&bar));
    }
    mbar_wait(smem_u32(&bar), 0);
    tc_fence_after();
    const uint32_t tw = tmem + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tw + c0, v);
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i) d[(warp * 32 + lane) * N + c0 + i] = __uint_as_float(v[i]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

int umma_selftest(const float* a, const float* b, int n, int k, int mode, float* d, cudaStream_t s) {
    if (n < 16 || n > 256 || n % 16 || k < 16 || k > 128 || k % 16) { set_error("umma_selftest: need 16<=N<=256 (x16), 16<=K<=128 (x16)"); return -1; }
    const size_t smem = (size_t)(128 + n) * k * 2;
    cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    umma_selftest_kernel<<<1, 128, smem, s>>>(a, b, n, k, mode, d);
    return count_launch();
}


// ------------------------------------------------------------------------------------------------
// MMA rate probe: `reps` back-to-back tcgen05.mma (M=128, N, K=16) from one thread, timed with clock64 from
// first issue to commit completion.  variant 0: A from TMEM, one accumulator; 1: A from SMEM; 2: A from TMEM,
// two accumulators alternating; 3: A from TMEM, fully unrolled by 8.  out[0] = cycles, out[1] = issue-only cycles.
__global__ void __launch_bounds__(128, 1) umma_rate_kernel(int N, int reps, int variant, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int t = threadIdx.x, warp = t >> 5;
    for (int i = t; i < (128 * 16 + 256 * 16) / 2; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;
    if (t == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    if (warp == 0) { tmem_alloc(smem_u32(&slot), 512); tmem_relinquish(); }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (t == 0) {
        const uint32_t idesc = make_idesc_f16(128, N, 0, 0);
        const uint32_t sa = smem_u32(smem), sb = sa + 4096;
        const uint64_t ad = make_desc(sa, 2048, 128), bd = make_desc(sb, N * 16, 128);
        const long long c0 = clock64();
        if (variant == 3) {
            for (int i = 0; i < reps; i += 8) {
#pragma unroll
                for (int j = 0; j < 8; ++j) mma_ts(tmem, tmem + 448, bd, idesc, 1);
            }
        } else {
            for (int i = 0; i < reps; ++i) {
                if (variant == 1) mma_ss(tmem, ad, bd, idesc, 1);
                else if (variant == 2) mma_ts(tmem + ((i & 1) ? 0 : 0) + (N <= 128 ? (i & 1) * N : 0), tmem + 448, bd, idesc, 1);
                else mma_ts(tmem, tmem + 448, bd, idesc, 1);
            }
        }
        const long long c1 = clock64();
        tc_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0);
        const long long c2 = clock64();
        out[0] = c2 - c0; out[1] = c1 - c0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// variant >= 8: (variant - 8 + 1) warps issue concurrently, warp-uniform code with one elected lane (the pattern of the
// two-stream training kernel), A and B from shared memory, one accumulator per warp.  out[2w] = cycles to completion of
// warp w's MMAs, out[2w+1] = issue-only cycles.
template <bool A_TMEM, bool STREAM_B>
__global__ void __launch_bounds__(256, 1) umma_rate2_kernel(int N, int reps, int nw, int flags, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar[4];
    __shared__ uint32_t slot;
    __shared__ volatile int stop;
    const int t = threadIdx.x, warp = t >> 5;
    for (int i = t; i < 80 * 1024 / 4; i += 256) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;
    if (t == 0) { for (int w = 0; w < 4; ++w) mbar_init(smem_u32(&bar[w]), 1); fence_barrier_init(); stop = 0; }
    if (warp == 0) { tmem_alloc(smem_u32(&slot), 512); tmem_relinquish(); }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    const bool background = flags & 2;
    if (warp < nw) {
        const uint32_t idesc = make_idesc_f16(128, N, 0, 0);
        const uint32_t sa = smem_u32(smem), sb = sa + 4096;
        const uint32_t alo = ((sa >> 4) & 0x3FFFu) | (128u << 16), blo = ((sb >> 4) & 0x3FFFu) | ((uint32_t)N << 16), hi = 8u | (1u << 14);
        const uint32_t d = tmem + (N * nw <= 256 ? warp * N : 0);
        const long long c0 = clock64();
        for (int i = 0; i < reps; i += 8) {
            if (elect_one()) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    // STREAM_B: a different 128-row K-step of B every MMA (walks 32 KB of shared memory like a weight matrix)
                    const uint32_t bj = STREAM_B ? blo + j * (uint32_t)(N * 2) : blo;
                    if (A_TMEM) mma_ts(d, tmem + 448 + 8 * (j & 3), ((uint64_t)hi << 32) | bj, idesc, 1);
                    else mma_ss(d, ((uint64_t)hi << 32) | alo, ((uint64_t)hi << 32) | bj, idesc, 1);
                }
            }
            __syncwarp();
        }
        const long long c1 = clock64();
        if (elect_one()) tc_commit(smem_u32(&bar[warp]));
        __syncwarp();
        mbar_wait(smem_u32(&bar[warp]), 0);
        const long long c2 = clock64();
        if ((t & 31) == 0) { out[2 * warp] = c2 - c0; out[2 * warp + 1] = c1 - c0; }
        if (warp == 0 && (t & 31) == 0) stop = 1;
    } else if (warp >= 4 && background) {
        // epilogue-like tensor-memory traffic on columns the MMAs do not touch: ld 128 columns, st 64 columns, repeat
        const uint32_t tl = tmem + ((uint32_t)((warp & 3) * 32) << 16) + 256;
        uint32_t v[32], h[16];
        for (int i = 0; i < 16; ++i) h[i] = i;
        while (!stop) {
            tmem_ld32(tl, v); tmem_ld32(tl + 32, v); tmem_ld32(tl + 64, v); tmem_ld32(tl + 96, v);
            tc_wait_ld();
            h[0] = v[0];
            tmem_st16(tl + 128, h); tmem_st16(tl + 144, h); tmem_st16(tl + 160, h); tmem_st16(tl + 176, h);
            tc_wait_st();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

int umma_rate(int n, int reps, int variant, long long* out, cudaStream_t s) {
    if (n < 16 || n > 256 || n % 16 || reps < 8 || reps % 8) { set_error("umma_rate: bad arguments"); return -1; }
    if (variant >= 8) {     // variant = 8 + (warps - 1) + 16 * flags   (flags: 1 = A from tensor memory, 2 = background tcgen05.ld/st traffic)
        const int nw = ((variant - 8) & 3) + 1, flags = (variant - 8) >> 4;
        auto kern = (flags & 4) ? ((flags & 1) ? umma_rate2_kernel<true, true> : umma_rate2_kernel<false, true>)
                                : ((flags & 1) ? umma_rate2_kernel<true, false> : umma_rate2_kernel<false, false>);
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
        kern<<<1, 256, 80 * 1024, s>>>(n, reps, nw, flags, out);
    } else
        umma_rate_kernel<<<1, 128, 16384, s>>>(n, reps, variant, out);
    return count_launch();
}

// ================================================================================================
// host side
static int round_up(int v, int m) { return (v + m - 1) / m * m; }

bool build_plan(const tnerf_handle* h, FusedPlan& pl) {
    if (h->hidden != 128 || h->depth < 1 || h->depth > kMaxDepth) return false;
    int L, inc;
    if (h->num_freqs >= 0 && h->in_dim == 6 * h->num_freqs + 3) { L = h->num_freqs; inc = 1; }
    else if (h->num_freqs >= 0 && h->in_dim == 6 * h->num_freqs) { L = h->num_freqs; inc = 0; }
    else return false;
    if (L < 0 || L > 10) return false;
    pl = FusedPlan{};
    pl.depth = h->depth; pl.D = h->in_dim; pl.L = L; pl.include_input = inc; pl.H = 128;
    pl.Kx = round_up(h->in_dim + 1, 16);
    if (pl.Kx > 64) return false;
    pl.bias_in_x = 1;    // Kx always leaves at least one pad column for the constant 1
    uint32_t off = 0;
    const int skip_layer = (h->skip_at >= 1 && h->skip_at <= h->depth - 1) ? h->skip_at : -1;
    for (int l = 0; l <= h->depth; ++l) {
        LayerPlan& lp = pl.layer[l];
        lp.b_off = off;
        const bool head = (l == h->depth);
        lp.N = head ? 16 : 128;
        lp.idesc = make_idesc_f16(128, lp.N, 0, 0);
        int ns = 0, ksteps = 0;
        if (l == 0) { lp.seg_kind[ns] = SEG_X; lp.seg_steps[ns++] = pl.Kx / 16; }
        else {
            lp.seg_kind[ns] = SEG_ACT; lp.seg_steps[ns++] = 8;
            if (l == skip_layer) { lp.seg_kind[ns] = SEG_X; lp.seg_steps[ns++] = pl.Kx / 16; }
            else { lp.seg_kind[ns] = SEG_ONES; lp.seg_steps[ns++] = 1; }
        }
        lp.nseg = ns;
        for (int i = 0; i < ns; ++i) ksteps += lp.seg_steps[i];
        off += (uint32_t)ksteps * 16u * lp.N * 2u;
    }
    pl.image_bytes = off;
    return true;
}

bool build_repack_map(const tnerf_handle* h, RepackMap& mp) {
    FusedPlan pl;
    mp = RepackMap{};
    if (!build_plan(h, pl) || !h->packed) return false;
    mp.n_tensors = h->n_params; mp.H = h->hidden; mp.image = reinterpret_cast<__half*>(h->packed);
    for (int t = 0; t < h->n_params; ++t) mp.off[t] = h->offsets[t];
    mp.off[h->n_params] = h->param_count;
    for (int l = 0; l <= h->depth; ++l) {
        const LayerPlan& lp = pl.layer[l];
        mp.N[l] = lp.N; mp.img_off[l] = lp.b_off / 2; mp.fan[l] = l < h->depth ? h->layer_in[l] : h->hidden;
        int k0 = 0;
        mp.act_len[l] = 0; mp.xstart[l] = 0; mp.bias_k[l] = -1;
        for (int sgi = 0; sgi < lp.nseg; ++sgi) {
            const int kind = lp.seg_kind[sgi], len = lp.seg_steps[sgi] * 16;
            if (kind == SEG_ACT) mp.act_len[l] = h->hidden;
            else if (kind == SEG_X) { mp.xstart[l] = k0; if (pl.bias_in_x) { mp.bias_k[l] = k0 + pl.Kx - 1; mp.bias_hilo[l] = 0; } }
            else { mp.bias_k[l] = k0; mp.bias_hilo[l] = 1; }
            k0 += len;
        }
        if (mp.bias_k[l] < 0) return false;
    }
    for (int l = 0; l < h->depth; ++l) { mp.kind[2 * l] = RP_W_HID; mp.layer[2 * l] = (uint8_t)l; mp.kind[2 * l + 1] = RP_B_HID; mp.layer[2 * l + 1] = (uint8_t)l; }
    const int d = h->depth;
    mp.kind[2 * d] = RP_W_SIG; mp.kind[2 * d + 1] = RP_B_SIG; mp.kind[2 * d + 2] = RP_W_RGB; mp.kind[2 * d + 3] = RP_B_RGB;
    for (int t = 2 * d; t < 2 * d + 4; ++t) mp.layer[t] = (uint8_t)d;
    mp.valid = 1;
    return true;
}

bool fused_shape_supported(const tnerf_handle* h) { FusedPlan pl; return build_plan(h, pl) || wide_shape_supported(h); }

int fused_pack_weights(tnerf_handle* h, cudaStream_t s) {
    if (wide_shape_supported(h)) return wide_pack_weights(h, s);
    FusedPlan pl;
    if (!build_plan(h, pl)) { set_error("fused path: unsupported MLP shape (need hidden=128, in_dim=6L(+3)<=63, depth<=8)"); return -2; }
    if (h->params.empty()) { set_error("pack_weights: parameters not bound"); return -3; }
    if (h->packed_bytes < pl.image_bytes) {
        if (h->packed) cudaFree(h->packed);
        cudaError_t e = cudaMalloc(&h->packed, pl.image_bytes);
        if (e != cudaSuccess) { set_error("cudaMalloc(packed image) failed"); return (int)e; }
        h->packed_bytes = pl.image_bytes;
    }
    PackArgs a{};
    a.plan = pl;
    for (int l = 0; l < h->depth; ++l) { a.W[l] = h->params[2 * l]; a.b[l] = h->params[2 * l + 1]; a.fan[l] = h->layer_in[l]; }
    a.W[h->depth] = h->params[2 * h->depth]; a.b[h->depth] = h->params[2 * h->depth + 1];
    a.W[h->depth + 1] = h->params[2 * h->depth + 2]; a.b[h->depth + 1] = h->params[2 * h->depth + 3];
    const long long total = pl.image_bytes / 2;
    pack_weights_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(a, reinterpret_cast<__half*>(h->packed));
    return count_launch();
}

static long long gcd_ll(long long a, long long b) { while (b) { long long t = a % b; a = b; b = t; } return a; }

// shapes the tensor-core render kernels cover; everything else runs on the exact fp32 path (tnerf_api.cu)
bool fused_render_shape_ok(const tnerf_handle* h, int S, bool want_weights) {
    if (S < 1 || S % 32 != 0) return false;
    const long long G = S / gcd_ll(S, 128);
    if (wide_shape_supported(h)) return G <= 4 && !want_weights;
    FusedPlan pl;
    return build_plan(h, pl) && G <= MAX_G && !(want_weights && G > 1);
}

int fused_render_fwd(tnerf_handle* h, const RaySource& rs, long long n, float nr, float fr, int S, const float* jitter, int white,
                     float* comp, float* depth, float* acc, float* weights, float* rays_d_out, cudaStream_t s) {
    if (n <= 0) return 0;
    if (wide_shape_supported(h)) return fused_render_fwd_wide(h, rs, n, nr, fr, S, jitter, white, comp, depth, acc, weights, rays_d_out, s);
    FwdParams p{};
    if (!build_plan(h, p.plan)) { set_error("fused path: unsupported MLP shape"); return -2; }
    if (!h->packed) { set_error("fused path: tnerf_pack_weights has not been called"); return -3; }
    const long long g = gcd_ll(S, 128);
    p.G = (int)(S / g); p.R = (int)(128 / g);
    if (S < 1 || p.G > MAX_G) { set_error("fused path: n_samples must satisfy n_samples/gcd(n_samples,128) <= 8"); return -4; }
    p.rs = rs; p.n_rays = n; p.n_units = (n + p.R - 1) / p.R; p.S = S; p.white = white; p.near_ = nr; p.far_ = fr;
    p.jitter = jitter; p.comp = comp; p.depth = depth; p.acc = acc; p.weights = weights; p.rays_d_out = rays_d_out;
    p.image = reinterpret_cast<const __half*>(h->packed);
    p.debug = reinterpret_cast<long long*>(h->debug);
    long long grid = (p.n_units + 1) / 2;
    if (grid > h->sm_count) grid = h->sm_count;
    if (!(S % 32 == 0 && !(weights && p.G > 1))) { set_error("fused path: shape not covered by the tensor-core render kernel"); return -5; }
    return fused_render_fwd_fast(p, (int)grid, s);
}

}  // namespace tnerf
