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

// ------------------------------------------------------------------------------------------------
// Fused forward kernel.  288 threads: warps 0-3 = warpgroup A, warps 4-7 = warpgroup B (each owns one
// 128-row tile at a time: thread <-> sample row <-> TMEM lane), warp 8 = MMA issuer + weight loader.
constexpr int FWD_THREADS = 320;   // 2 x 4 row warps + one MMA-issuer warp per warpgroup
constexpr int TM_ACC = 0, TM_ACT = 128, TM_X = 192, TM_HEAD = 224, TM_ONES = 240, TM_WG_STRIDE = 256;
constexpr int MAX_G = 8;

struct FwdSmem {   // trailing part of dynamic smem (after the weight image)
    float4 stage[2][MAX_G * 128];   // per warpgroup: (sigma, r, g, b) per sample of the current unit
    float stage_z[2][MAX_G * 128];
    uint64_t bar_w, bar_a[2], bar_acc[2];
    uint32_t tmem_slot;
};

__device__ __forceinline__ void composite_unit(const FwdParams& p, const float4* st, const float* sz, long long ray0, int warp_q,
                                               int lane) {
    const int S = p.S;
    for (int rr = warp_q; rr < p.R; rr += 4) {
        const long long ray = ray0 + rr;
        if (ray >= p.n_rays) break;
        float o[3], d[3];
        load_ray(p.rs, ray, o, d);
        const float dn = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        const float4* s4 = st + rr * S;
        const float* zz = sz + rr * S;
        float T_carry = 1.f, cr = 0.f, cg = 0.f, cb = 0.f, dsum = 0.f, asum = 0.f;
        for (int base = 0; base < S; base += 32) {
            const int i = base + lane;
            const bool ok = i < S;
            float zi = 0.f, alpha = 0.f, q = 1.f;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ok) {
                v = s4[i];
                zi = zz[i];
                const float gap = ((i == S - 1) ? kLastDelta : (zz[i + 1] - zi)) * dn;
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
            if (ok) {
                if (p.weights) p.weights[ray * S + i] = w;
                cr += w * v.y; cg += w * v.z; cb += w * v.w;
                dsum += w * zi; asum += w;
            }
            T_carry *= __shfl_sync(0xffffffffu, incl, 31);
        }
        cr = warp_sum(cr); cg = warp_sum(cg); cb = warp_sum(cb); dsum = warp_sum(dsum); asum = warp_sum(asum);
        if (lane == 0) {
            const float bg = p.white ? 1.f - asum : 0.f;
            p.comp[3 * ray] = cr + bg; p.comp[3 * ray + 1] = cg + bg; p.comp[3 * ray + 2] = cb + bg;
            if (p.depth) p.depth[ray] = dsum;
            if (p.acc) p.acc[ray] = asum;
            if (p.rays_d_out) { p.rays_d_out[3 * ray] = d[0]; p.rays_d_out[3 * ray + 1] = d[1]; p.rays_d_out[3 * ray + 2] = d[2]; }
        }
    }
}

template <int KX>
__global__ void __launch_bounds__(FWD_THREADS, 1) fused_fwd_kernel(const __grid_constant__ FwdParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    FwdSmem& sm = *reinterpret_cast<FwdSmem*>(smem + ((p.plan.image_bytes + 1023u) & ~1023u));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t bar_w = smem_u32(&sm.bar_w);

    if (warp == 8 && lane == 0) {
        mbar_init(bar_w, 1);
        for (int w = 0; w < 2; ++w) { mbar_init(smem_u32(&sm.bar_a[w]), 128); mbar_init(smem_u32(&sm.bar_acc[w]), 1); }
        fence_barrier_init();
    }
    if (warp == 0) { tmem_alloc(smem_u32(&sm.tmem_slot), 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_slot;
    long long* dbg = (p.debug && blockIdx.x == 0 && warp != 9) ? p.debug : nullptr;
    int dbg_n = 0;
#define STAMP(base) do { if (dbg && dbg_n < 500) dbg[(base) + dbg_n++] = clock64(); } while (0)

    // units are dealt round-robin: (cta, warpgroup) pair j takes units j, j + 2*grid, ...
    const long long stride = 2LL * gridDim.x;

    if (warp >= 8) {
        // ------------------------------ MMA issuers: warp 8 serves warpgroup 0, warp 9 warpgroup 1 ------------------------------
        const int w = warp - 8;
        if (lane == 0) {
            if (w == 0) {
                mbar_expect_tx(bar_w, p.plan.image_bytes);
                uint32_t off = 0;
                while (off < p.plan.image_bytes) {
                    const uint32_t n = min(32768u, p.plan.image_bytes - off);
                    bulk_g2s(smem_u32(smem) + off, reinterpret_cast<const uint8_t*>(p.image) + off, n, bar_w);
                    off += n;
                }
            }
            mbar_wait(bar_w, 0);
            uint32_t phase = 0;
            const uint32_t wbase = smem_u32(smem);
            const uint32_t tw = tmem + w * TM_WG_STRIDE;
            const uint32_t bar_a = smem_u32(&sm.bar_a[w]), bar_acc = smem_u32(&sm.bar_acc[w]);
            for (long long u = 2LL * blockIdx.x + w; u < p.n_units; u += stride) {
                for (int g = 0; g < p.G; ++g) {
                    for (int step = 0; step <= p.plan.depth; ++step) {
                        const LayerPlan& lp = p.plan.layer[step];
                        const uint32_t idesc = lp.idesc, N = lp.N;
                        const uint32_t b_adv = (N * 32u) >> 4;
                        uint32_t b_lo = (((wbase + lp.b_off) >> 4) & 0x3FFFu) | (((N * 16u) >> 4) << 16);
                        const uint32_t b_hi = (128u >> 4) | (1u << 14);          // SBO | descriptor version
                        const uint32_t d_t = tw + ((step == p.plan.depth) ? TM_HEAD : TM_ACC);
                        const int nseg = lp.nseg;
                        uint32_t seg_a[3];
                        int seg_n[3];
#pragma unroll
                        for (int sgi = 0; sgi < 3; ++sgi) {
                            const uint32_t kind = lp.seg_kind[sgi];
                            seg_a[sgi] = tw + (kind == SEG_ACT ? TM_ACT : kind == SEG_X ? TM_X : TM_ONES);
                            seg_n[sgi] = sgi < nseg ? lp.seg_steps[sgi] : 0;
                        }
                        mbar_wait(bar_a, phase);
                        phase ^= 1;
                        tc_fence_after();
                        STAMP(512);
                        uint32_t acc = 0;
#pragma unroll
                        for (int sgi = 0; sgi < 3; ++sgi) issue_ts_n(seg_n[sgi], d_t, seg_a[sgi], b_lo, b_hi, b_adv, idesc, acc);
                        tc_commit(bar_acc);
                        STAMP(512);
                    }
                }
            }
        }
        __syncwarp();
    } else {
        // ------------------------------ sample / epilogue warpgroups ------------------------------
        const int wg = warp >> 2, q = warp & 3, row = q * 32 + lane;
        const uint32_t tw = tmem + wg * TM_WG_STRIDE + ((uint32_t)(q * 32) << 16);
        const uint32_t bar_a = smem_u32(&sm.bar_a[wg]), bar_acc = smem_u32(&sm.bar_acc[wg]);
        float4* st = sm.stage[wg];
        float* sz = sm.stage_z[wg];
        {   // constant-one chunk used to add biases inside the GEMM: A[:,0] = A[:,1] = 1
            uint32_t ones[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) ones[i] = 0u;
            ones[0] = 0x3C003C00u;
            tmem_st8(tw + TM_ONES, ones);
        }
        uint32_t phase = 0;
        const bool jit = p.jitter != nullptr;
        if (!(wg == 0 && row == 0)) dbg = nullptr;
        for (long long u = 2LL * blockIdx.x + wg; u < p.n_units; u += stride) {
            const long long ray0 = u * p.R;
            for (int g = 0; g < p.G; ++g) {
                STAMP(0);
                const int urow = g * 128 + row;            // row inside the unit
                const long long ray = ray0 + urow / p.S;
                const int si = urow % p.S;
                const bool valid = ray < p.n_rays;
                float pt[3] = {0.f, 0.f, 0.f};
                float z = 0.f;
                if (valid) {
                    float o[3], d[3];
                    load_ray(p.rs, ray, o, d);
                    z = depth_sample(si, p.S, p.near_, p.far_, jit ? p.jitter[ray * p.S + si] : 0.f, jit);
#pragma unroll
                    for (int c = 0; c < 3; ++c) pt[c] = __fadd_rn(o[c], __fmul_rn(d[c], z));
                }
                sz[urow] = z;
                {
                    uint32_t pk[KX / 2];
                    if (p.plan.include_input) encode_point<KX, true>(pt, p.plan.L, pk);
                    else encode_point<KX, false>(pt, p.plan.L, pk);
#pragma unroll
                    for (int c = 0; c < KX / 32; ++c) {
                        uint32_t chunk[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) chunk[i] = pk[c * 16 + i];
                        tmem_st16(tw + TM_X + c * 16, chunk);
                    }
                    if (KX % 32) {
                        uint32_t chunk[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) chunk[i] = pk[(KX / 32) * 16 + i];
                        tmem_st8(tw + TM_X + (KX / 32) * 16, chunk);
                    }
                }
                tc_wait_st();
                tc_fence_before();
                mbar_arrive(bar_a);
                STAMP(0);
                for (int l = 0; l < p.plan.depth; ++l) {
                    mbar_wait(bar_acc, phase);
                    phase ^= 1;
                    tc_fence_after();
                    STAMP(0);
#pragma unroll
                    for (int c = 0; c < 4; c += 2) {
                        uint32_t v0[32], v1[32];
                        tmem_ld32(tw + TM_ACC + c * 32, v0);
                        tmem_ld32(tw + TM_ACC + c * 32 + 32, v1);
                        tc_wait_ld();
                        uint32_t h0[16], h1[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) h0[i] = pack_relu_h2(__uint_as_float(v0[2 * i]), __uint_as_float(v0[2 * i + 1]));
                        tmem_st16(tw + TM_ACT + c * 16, h0);
#pragma unroll
                        for (int i = 0; i < 16; ++i) h1[i] = pack_relu_h2(__uint_as_float(v1[2 * i]), __uint_as_float(v1[2 * i + 1]));
                        tmem_st16(tw + TM_ACT + c * 16 + 16, h1);
                    }
                    tc_wait_st();
                    tc_fence_before();
                    mbar_arrive(bar_a);
                    STAMP(0);
                }
                mbar_wait(bar_acc, phase);
                phase ^= 1;
                tc_fence_after();
                STAMP(0);
                {
                    uint32_t v[4];
                    tmem_ld4(tw + TM_HEAD, v);
                    tc_wait_ld();
                    float4 o4;
                    o4.x = fmaxf(__uint_as_float(v[0]), 0.f);
                    o4.y = 1.f / (1.f + __expf(-__uint_as_float(v[1])));
                    o4.z = 1.f / (1.f + __expf(-__uint_as_float(v[2])));
                    o4.w = 1.f / (1.f + __expf(-__uint_as_float(v[3])));
                    st[urow] = o4;
                }
            }
            STAMP(0);
            bar_sync(1 + wg, 128);
            composite_unit(p, st, sz, ray0, q, lane);
            bar_sync(1 + wg, 128);
            STAMP(0);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

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
    const size_t smem = ((p.plan.image_bytes + 1023u) & ~1023u) + sizeof(FwdSmem);
    long long grid = (p.n_units + 1) / 2;
    if (grid > h->sm_count) grid = h->sm_count;
    if (S % 32 == 0 && !(weights && p.G > 1)) return fused_render_fwd_fast(p, (int)grid, s);
    auto kern = p.plan.Kx == 64 ? fused_fwd_kernel<64> : p.plan.Kx == 48 ? fused_fwd_kernel<48> : p.plan.Kx == 32 ? fused_fwd_kernel<32> : fused_fwd_kernel<16>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("fused fwd: shared memory request rejected"); return (int)e; }
    kern<<<(unsigned)grid, FWD_THREADS, smem, s>>>(p);
    return count_launch();
}

}  // namespace tnerf
