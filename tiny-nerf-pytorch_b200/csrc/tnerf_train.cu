// Fused training kernel (sm_100a): forward + MSE (or upstream gradients) + backward of one 128-sample
// tile entirely on chip, all GEMMs on tcgen05.  Replaces src/train.py:114-126 of the reference
// (stratified_samples -> encoder -> model -> volume_render -> loss -> backward) with ONE launch.
//
// Per tile (128 samples = 128/S whole rays), see DESIGN.md section 5:
//   recompute   X -> H0 -> H1 -> H2 -> H3 -> heads      (activations as fp16 operand images in smem)
//   composite   forward scan, loss gradient, reverse scan (warp per ray, fp32)
//   backward    head dgrad/wgrad, then per layer: wgrad (fp32 accumulators resident in tensor memory
//               or registers for the whole kernel), dgrad, ReLU mask, in-place dZ image
// Weights are streamed per layer from the packed fp16 image (L2-resident) through a 2 x 48 KB
// bulk-copy ring; weight gradients leave the SM once, at kernel end, as one coalesced slab per CTA.
//
// Warp roles: warps 0-3 = row warps (thread <-> sample row <-> TMEM lane; in wgrad drains thread <->
// output-feature row), warp 4 = MMA issuer (table driven), warp 5 = weight loader.
#include <cstdlib>
#include "tnerf_train.cuh"

namespace tnerf {

constexpr int TR_THREADS = 192;
constexpr int TC_DW2 = 0, TC_DW3 = 192, TC_DW0 = 320, TC_WORK = 384, TC_TEMP = 448;
constexpr uint32_t WBUF_BYTES = 49152;
constexpr uint32_t SM_WBUF = 0, SM_X = 98304, SM_SLOT_A = 114688, SM_SLOT_B = 147456, SM_SLOT_C = 180224, SM_DZH = 212992,
                   SM_ONES = 217088, SM_MISC = 221184;
constexpr int MAX_JOBS = 28, N_WLOADS = 9;

struct Seg {
    uint32_t a_off, b_off;                    // byte offsets (a: from smem base; b: from smem base or weight-ring slot)
    uint16_t a_lbo, a_sbo, a_adv, b_lbo, b_sbo, b_adv;   // bytes >> 4
    uint8_t steps, b_w;
    uint16_t pad_;
};
struct Job {
    Seg seg[2];
    uint32_t idesc;
    uint16_t d_col;
    uint8_t nseg, wait_op, w_acquire, w_release, resident, commit;
};
struct WLoad { uint32_t src_off, bytes; };
struct TrainPlan {
    Job job[MAX_JOBS];
    WLoad wl[N_WLOADS];
    int njobs;
};

struct TrainMisc {
    float4 stage[128];
    float stage_z[128];
    uint64_t bar_op, bar_acc, bar_full[2], bar_empty[2];
    uint32_t tmem_slot;
};

// ---- row-warp helpers ----------------------------------------------------------------------------
// forward drain of a 64-column half: relu, fp16, K-major operand image (chunk = 8 features = 16 B per row)
__device__ __forceinline__ void drain_fwd_half(uint32_t tw_work, uint8_t* slot, int half, int row) {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld32(tw_work + c * 32, v);
        tc_wait_ld();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint4 o;
            o.x = pack_relu_h2(__uint_as_float(v[8 * j + 0]), __uint_as_float(v[8 * j + 1]));
            o.y = pack_relu_h2(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3]));
            o.z = pack_relu_h2(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5]));
            o.w = pack_relu_h2(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7]));
            *reinterpret_cast<uint4*>(slot + ((size_t)((half * 8 + c * 4 + j) * 128 + row) << 4)) = o;
        }
    }
}
// backward drain: dZ = dH * (H > 0), written over H (same bytes, same thread)
__device__ __forceinline__ void drain_bwd_half(uint32_t tw_work, uint8_t* slot, int half, int row) {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld32(tw_work + c * 32, v);
        tc_wait_ld();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint4* p = reinterpret_cast<uint4*>(slot + ((size_t)((half * 8 + c * 4 + j) * 128 + row) << 4));
            const uint4 h = *p;
            uint4 o;
            o.x = pack_sat_h2(__uint_as_float(v[8 * j + 0]), __uint_as_float(v[8 * j + 1])) & relu_mask(h.x);
            o.y = pack_sat_h2(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3])) & relu_mask(h.y);
            o.z = pack_sat_h2(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5])) & relu_mask(h.z);
            o.w = pack_sat_h2(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7])) & relu_mask(h.w);
            *p = o;
        }
    }
}

template <int KX>
__global__ void __launch_bounds__(TR_THREADS, 1) fused_train_kernel(const __grid_constant__ TrainParams p,
                                                                     const __grid_constant__ TrainPlan plan) {
    extern __shared__ __align__(1024) uint8_t smem[];
    TrainMisc& ms = *reinterpret_cast<TrainMisc*>(smem + SM_MISC);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar_op = smem_u32(&ms.bar_op), bar_acc = smem_u32(&ms.bar_acc);

    if (warp == 4 && lane == 0) {
        mbar_init(bar_op, 128);
        mbar_init(bar_acc, 1);
        for (int b = 0; b < 2; ++b) { mbar_init(smem_u32(&ms.bar_full[b]), 1); mbar_init(smem_u32(&ms.bar_empty[b]), 1); }
        fence_barrier_init();
    }
    if (warp == 0) { tmem_alloc(smem_u32(&ms.tmem_slot), 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = ms.tmem_slot;
    const long long n_my = (p.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;   // tiles blockIdx.x, +grid, ...

    if (warp == 5) {
        // ------------------------------ weight loader ------------------------------
        if (lane == 0) {
            uint32_t g = 0;
            for (long long t = 0; t < n_my; ++t) {
                for (int l = 0; l < N_WLOADS; ++l, ++g) {
                    const uint32_t buf = g & 1;
                    if (g >= 2) mbar_wait(smem_u32(&ms.bar_empty[buf]), ((g >> 1) - 1) & 1);
                    const uint32_t full = smem_u32(&ms.bar_full[buf]);
                    mbar_expect_tx(full, plan.wl[l].bytes);
                    uint32_t off = 0;
                    while (off < plan.wl[l].bytes) {
                        const uint32_t n = min(16384u, plan.wl[l].bytes - off);
                        bulk_g2s(sbase + SM_WBUF + buf * WBUF_BYTES + off,
                                 reinterpret_cast<const uint8_t*>(p.image) + plan.wl[l].src_off + off, n, full);
                        off += n;
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 4) {
        // ------------------------------ MMA issuer ------------------------------
        if (lane == 0) {
            uint32_t ph_op = 0, g = 0, cur_buf = 0;
            for (long long t = 0; t < n_my; ++t) {
                for (int j = 0; j < plan.njobs; ++j) {
                    const Job& jb = plan.job[j];
                    if (jb.wait_op) { mbar_wait(bar_op, ph_op); ph_op ^= 1; tc_fence_after(); }
                    if (jb.w_acquire) {
                        cur_buf = g & 1;
                        mbar_wait(smem_u32(&ms.bar_full[cur_buf]), (g >> 1) & 1);
                        ++g;
                    }
                    const uint32_t d_t = tmem + jb.d_col;
                    const uint32_t idesc = jb.idesc;
                    uint32_t acc = (jb.resident && t > 0) ? 1u : 0u;
                    const uint32_t wb = sbase + SM_WBUF + cur_buf * WBUF_BYTES;
#pragma unroll
                    for (int s = 0; s < 2; ++s) {
                        if (s < jb.nseg) {
                            const Seg& sg = jb.seg[s];
                            // descriptor words: lo = start>>4 | LBO<<16 ; hi = SBO | version<<14 ; only lo advances
                            uint32_t a_lo = (((sbase + sg.a_off) >> 4) & 0x3FFFu) | ((uint32_t)sg.a_lbo << 16);
                            uint32_t b_lo = ((((sg.b_w ? wb : sbase) + sg.b_off) >> 4) & 0x3FFFu) | ((uint32_t)sg.b_lbo << 16);
                            const uint32_t a_hi = (uint32_t)sg.a_sbo | (1u << 14), b_hi = (uint32_t)sg.b_sbo | (1u << 14);
                            issue_ss_n(sg.steps, d_t, a_lo, a_hi, sg.a_adv, b_lo, b_hi, sg.b_adv, idesc, acc);
                        }
                    }
                    if (jb.w_release) tc_commit(smem_u32(&ms.bar_empty[cur_buf]));
                    if (jb.commit) tc_commit(bar_acc);
                }
            }
        }
        __syncwarp();
    } else {
        // ------------------------------ row warps ------------------------------
        const int q = warp, row = q * 32 + lane;
        const uint32_t tw = tmem + ((uint32_t)(q * 32) << 16);
        uint8_t* slotA = smem + SM_SLOT_A;
        uint8_t* slotB = smem + SM_SLOT_B;
        uint8_t* slotC = smem + SM_SLOT_C;
        uint32_t ph_acc = 0;
        float dw1[128];
#pragma unroll
        for (int i = 0; i < 128; ++i) dw1[i] = 0.f;
        float dwh[4] = {0.f, 0.f, 0.f, 0.f}, db1 = 0.f, db3 = 0.f, hb[4] = {0.f, 0.f, 0.f, 0.f}, loss_acc = 0.f;
        // constant images: ONES (k/n = 0,1 -> 1.0) and the zero half of the dZh image
        *reinterpret_cast<uint4*>(smem + SM_ONES + ((size_t)row << 4)) = make_uint4(0x3C003C00u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(smem + SM_ONES + ((size_t)(128 + row) << 4)) = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(smem + SM_DZH + ((size_t)(128 + row) << 4)) = make_uint4(0u, 0u, 0u, 0u);
        const bool jit = p.jitter != nullptr;
        const float gscale = p.scale_dev ? *p.scale_dev : p.scale;

#define WAIT_ACC() do { mbar_wait(bar_acc, ph_acc); ph_acc ^= 1; tc_fence_after(); } while (0)
#define SIGNAL_OP() do { fence_proxy_async(); tc_fence_before(); mbar_arrive(bar_op); } while (0)

        for (long long t = 0; t < n_my; ++t) {
            const long long tile = blockIdx.x + t * gridDim.x;
            const long long ray0 = tile * p.R;
            const long long ray = ray0 + row / p.S;
            const int si = row % p.S;
            const bool valid = ray < p.n_rays;
            {   // sample point + Fourier features -> X image
                float pt[3] = {0.f, 0.f, 0.f};
                float z = 0.f;
                if (valid) {
                    float o[3], d[3];
                    load_ray(p.rs, ray, o, d);
                    z = depth_sample(si, p.S, p.near_, p.far_, jit ? p.jitter[ray * p.S + si] : 0.f, jit);
#pragma unroll
                    for (int c = 0; c < 3; ++c) pt[c] = __fadd_rn(o[c], __fmul_rn(d[c], z));
                }
                ms.stage_z[row] = z;
                uint32_t pk[KX / 2];
                if (p.include_input) encode_point<KX, true>(pt, p.L, pk); else encode_point<KX, false>(pt, p.L, pk);
#pragma unroll
                for (int c = 0; c < KX / 8; ++c)
                    *reinterpret_cast<uint4*>(smem + SM_X + ((size_t)(c * 128 + row) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
            }
            SIGNAL_OP();
            // forward recompute: H0 -> A, H1 -> B, H2 -> C, H3 -> A
            WAIT_ACC(); drain_fwd_half(tw + TC_WORK, slotA, 0, row); SIGNAL_OP();
            WAIT_ACC(); drain_fwd_half(tw + TC_WORK, slotA, 1, row); SIGNAL_OP();
            WAIT_ACC(); drain_fwd_half(tw + TC_WORK, slotB, 0, row); SIGNAL_OP();
            WAIT_ACC(); drain_fwd_half(tw + TC_WORK, slotB, 1, row); SIGNAL_OP();
            WAIT_ACC(); drain_fwd_half(tw + TC_WORK, slotC, 0, row); SIGNAL_OP();
            WAIT_ACC(); drain_fwd_half(tw + TC_WORK, slotC, 1, row); SIGNAL_OP();
            WAIT_ACC(); drain_fwd_half(tw + TC_WORK, slotA, 0, row); SIGNAL_OP();
            WAIT_ACC(); drain_fwd_half(tw + TC_WORK, slotA, 1, row); SIGNAL_OP();
            // heads + composite
            WAIT_ACC();
            float4 own;
            {
                uint32_t v[4];
                tmem_ld4(tw + TC_TEMP, v);
                tc_wait_ld();
                own.x = fmaxf(__uint_as_float(v[0]), 0.f);
                own.y = 1.f / (1.f + __expf(-__uint_as_float(v[1])));
                own.z = 1.f / (1.f + __expf(-__uint_as_float(v[2])));
                own.w = 1.f / (1.f + __expf(-__uint_as_float(v[3])));
                ms.stage[row] = own;
            }
            bar_sync(1, 128);
            loss_acc += composite_tile<4>(p, ms.stage, ms.stage_z, ray0, q, lane);
            bar_sync(1, 128);
            {
                const float4 g = ms.stage[row];
                const float s0 = (own.x > 0.f) ? g.x * gscale : 0.f;
                const float s1 = g.y * own.y * (1.f - own.y) * gscale;
                const float s2 = g.z * own.z * (1.f - own.z) * gscale;
                const float s3 = g.w * own.w * (1.f - own.w) * gscale;
                *reinterpret_cast<uint4*>(smem + SM_DZH + ((size_t)row << 4)) = make_uint4(pack_sat_h2(s0, s1), pack_sat_h2(s2, s3), 0u, 0u);
                const float r0 = warp_sum(s0), r1 = warp_sum(s1), r2 = warp_sum(s2), r3 = warp_sum(s3);
                hb[0] += r0; hb[1] += r1; hb[2] += r2; hb[3] += r3;
            }
            SIGNAL_OP();
            // head wgrad (thread <-> feature row) + head dgrad half 0
            WAIT_ACC();
            {
                uint32_t v[4];
                tmem_ld4(tw + TC_TEMP + 16, v);
                tc_wait_ld();
#pragma unroll
                for (int i = 0; i < 4; ++i) dwh[i] += __uint_as_float(v[i]);
            }
            drain_bwd_half(tw + TC_WORK, slotA, 0, row); SIGNAL_OP();
            WAIT_ACC(); drain_bwd_half(tw + TC_WORK, slotA, 1, row); SIGNAL_OP();
            // layer 3: wgrad (resident), bias grad, dgrad -> dZ2 over H2
            WAIT_ACC();
            { uint32_t v[4]; tmem_ld4(tw + TC_TEMP + 32, v); tc_wait_ld(); db3 += __uint_as_float(v[0]); }
            drain_bwd_half(tw + TC_WORK, slotC, 0, row); SIGNAL_OP();
            WAIT_ACC(); drain_bwd_half(tw + TC_WORK, slotC, 1, row); SIGNAL_OP();
            // layer 2: wgrads (resident), dgrad -> dZ1 over H1
            WAIT_ACC(); drain_bwd_half(tw + TC_WORK, slotB, 0, row); SIGNAL_OP();
            WAIT_ACC(); drain_bwd_half(tw + TC_WORK, slotB, 1, row); SIGNAL_OP();
            // recompute H0 -> A
            WAIT_ACC(); drain_fwd_half(tw + TC_WORK, slotA, 0, row); SIGNAL_OP();
            WAIT_ACC(); drain_fwd_half(tw + TC_WORK, slotA, 1, row); SIGNAL_OP();
            // layer 1 wgrad -> registers (thread <-> output-feature row)
            WAIT_ACC();
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t v[32];
                tmem_ld32(tw + TC_WORK + c * 32, v);
                tc_wait_ld();
#pragma unroll
                for (int i = 0; i < 32; ++i) dw1[c * 32 + i] += __uint_as_float(v[i]);
            }
            SIGNAL_OP();
            // layer 1 bias grad + dgrad -> dZ0 over H0
            WAIT_ACC();
            { uint32_t v[4]; tmem_ld4(tw + TC_TEMP + 48, v); tc_wait_ld(); db1 += __uint_as_float(v[0]); }
            drain_bwd_half(tw + TC_WORK, slotA, 0, row); SIGNAL_OP();
            WAIT_ACC(); drain_bwd_half(tw + TC_WORK, slotA, 1, row); SIGNAL_OP();
            // layer 0 wgrad (resident); tile done when it has been issued and completed
            WAIT_ACC();
        }
#undef WAIT_ACC
#undef SIGNAL_OP
        // ---- flush this CTA's weight-gradient slab (coalesced: consecutive rows) ----
        float* slab = p.slabs + (size_t)blockIdx.x * p.sm.total;
        auto flush_tmem = [&](int tcol, int ncols, int off) {
            for (int c0 = 0; c0 < ncols; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(tw + tcol + c0, v);
                tc_wait_ld();
#pragma unroll
                for (int i = 0; i < 16; ++i) slab[off + (c0 + i) * 128 + row] = __uint_as_float(v[i]);
            }
        };
        flush_tmem(TC_DW0, KX, p.sm.dw0);
        flush_tmem(TC_DW2, 128 + KX, p.sm.dw2);
        flush_tmem(TC_DW3, 128, p.sm.dw3);
#pragma unroll
        for (int i = 0; i < 128; ++i) slab[p.sm.dw1 + i * 128 + row] = dw1[i];
#pragma unroll
        for (int i = 0; i < 4; ++i) slab[p.sm.dwh + i * 128 + row] = dwh[i];
        slab[p.sm.db1 + row] = db1;
        slab[p.sm.db3 + row] = db3;
        // head biases and loss: per-warp partials (identical in every lane after warp_sum) -> one slot per warp
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) slab[p.sm.hb + q * 4 + i] = hb[i];
            if (p.loss_sum && loss_acc != 0.f) atomicAdd(p.loss_sum, loss_acc);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------
// slab reduction: grads[flat] += (sum over CTAs of slab[.]) / scale, scattering the TMEM-native
// [col][row] slab layout into the state_dict-ordered flat gradient.
__global__ void reduce_slabs_kernel(ReduceArgs a) {
    // launched with programmatic stream serialisation: the grid is set up while the training kernel drains; wait for its results here
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= a.sm.total) return;
    float s = 0.f;
    for (int i = 0; i < a.n_slabs; ++i) s += a.slabs[(size_t)i * a.sm.total + e];
    if (a.zero_after) const_cast<float*>(a.slabs)[e] = 0.f;
    s *= a.scale_dev ? 1.f / *a.scale_dev : a.inv_scale;
    long long dst = -1;
    const int fan2 = 128 + a.D;
    if (e >= a.sm.hb) {
        const int o = (e - a.sm.hb) & 3;                   // 4 warps x 4 head biases
        if (e - a.sm.hb < 16) dst = (o == 0) ? a.off_bs : a.off_bc + (o - 1);
    } else if (e >= a.sm.db3) dst = a.off_b[3] + (e - a.sm.db3);
    else if (e >= a.sm.db1) dst = a.off_b[1] + (e - a.sm.db1);
    else if (e >= a.sm.dwh) {
        const int o = (e - a.sm.dwh) / 128, f = (e - a.sm.dwh) % 128;
        dst = (o == 0) ? a.off_ws + f : a.off_wc + (long long)(o - 1) * 128 + f;
    } else if (e >= a.sm.dw3) {
        const int c = (e - a.sm.dw3) / 128, m = (e - a.sm.dw3) % 128;
        dst = a.off_w[3] + (long long)m * 128 + c;
    } else if (e >= a.sm.dw2) {
        const int c = (e - a.sm.dw2) / 128, m = (e - a.sm.dw2) % 128;
        if (c < 128) dst = a.off_w[2] + (long long)m * fan2 + c;
        else if (c - 128 < a.D) dst = a.off_w[2] + (long long)m * fan2 + c;
        else if (c - 128 == a.Kx - 1) dst = a.off_b[2] + m;
    } else if (e >= a.sm.dw1) {
        const int c = (e - a.sm.dw1) / 128, m = (e - a.sm.dw1) % 128;
        dst = a.off_w[1] + (long long)m * 128 + c;
    } else {
        const int c = (e - a.sm.dw0) / 128, m = (e - a.sm.dw0) % 128;
        if (c < a.D) dst = a.off_w[0] + (long long)m * a.D + c;
        else if (c == a.Kx - 1) dst = a.off_b[0] + m;
    }
    if (dst >= 0) {
        if (e >= a.sm.hb) atomicAdd(a.grads + dst, s);     // four per-warp partials land on the same element
        else a.grads[dst] += s;
    }
}

// ------------------------------------------------------------------------------------------------
// host: job table
static Seg seg_kmajorA(uint32_t a_off, int steps) {
    Seg s{}; s.a_off = a_off; s.a_lbo = 2048 >> 4; s.a_sbo = 128 >> 4; s.a_adv = 4096 >> 4; s.steps = (uint8_t)steps; return s;
}
static Seg seg_mnA(uint32_t a_off, int steps) {
    Seg s{}; s.a_off = a_off; s.a_lbo = 128 >> 4; s.a_sbo = 2048 >> 4; s.a_adv = 256 >> 4; s.steps = (uint8_t)steps; return s;
}
static void b_kmajorW(Seg& s, uint32_t off, int R) { s.b_off = off; s.b_lbo = (uint16_t)(R * 16 >> 4); s.b_sbo = 128 >> 4; s.b_adv = (uint16_t)(R * 32 >> 4); s.b_w = 1; }
static void b_mnW(Seg& s, uint32_t off, int R) { s.b_off = off; s.b_lbo = 128 >> 4; s.b_sbo = (uint16_t)(R * 16 >> 4); s.b_adv = 256 >> 4; s.b_w = 1; }
static void b_mnS(Seg& s, uint32_t off) { s.b_off = off; s.b_lbo = 128 >> 4; s.b_sbo = 2048 >> 4; s.b_adv = 256 >> 4; s.b_w = 0; }

static bool build_train_plan(const FusedPlan& fp, TrainPlan& tp) {
    if (fp.depth != 4 || fp.H != 128) return false;
    // the skip must feed layer 2: layer plans 0:[X] 1:[ACT,ONES] 2:[ACT,X] 3:[ACT,ONES] head:[ACT,ONES]
    if (fp.layer[2].nseg != 2 || fp.layer[2].seg_kind[1] != SEG_X || fp.layer[1].seg_kind[1] != SEG_ONES || fp.layer[3].seg_kind[1] != SEG_ONES) return false;
    const int Kx = fp.Kx, xs = Kx / 16;
    tp = TrainPlan{};
    const uint32_t o0 = fp.layer[0].b_off, o1 = fp.layer[1].b_off, o2 = fp.layer[2].b_off, o3 = fp.layer[3].b_off, oh = fp.layer[4].b_off;
    tp.wl[0] = {o0, (uint32_t)Kx * 256u};
    tp.wl[1] = {o1, 144u * 256u};
    tp.wl[2] = {o2, (uint32_t)(128 + Kx) * 256u};
    tp.wl[3] = {o3, 144u * 256u};
    tp.wl[4] = {oh, 144u * 32u};
    tp.wl[5] = {o3, 32768u};
    tp.wl[6] = {o2, 32768u};
    tp.wl[7] = {o0, (uint32_t)Kx * 256u};
    tp.wl[8] = {o1, 32768u};
    int n = 0;
    auto add = [&](Job j) { tp.job[n++] = j; };
    auto fwd_half = [&](uint32_t a_off, int a_steps, int second /*0 none, 1 ONES, 2 X*/, int h, bool first_use, bool last_use) {
        Job j{};
        j.seg[0] = seg_kmajorA(a_off, a_steps);
        b_kmajorW(j.seg[0], 1024u * h, 128);
        j.nseg = 1;
        if (second) {
            j.seg[1] = seg_kmajorA(second == 1 ? SM_ONES : SM_X, second == 1 ? 1 : xs);
            b_kmajorW(j.seg[1], 16u * 2048u + 1024u * h, 128);
            j.nseg = 2;
        }
        j.idesc = make_idesc_f16(128, 64, 0, 0);
        j.d_col = TC_WORK; j.wait_op = 1; j.w_acquire = first_use; j.w_release = last_use; j.commit = 1;
        return j;
    };
    auto dgrad_half = [&](uint32_t a_off, int h, bool wait_op, bool first_use, bool last_use) {
        Job j{};
        j.seg[0] = seg_kmajorA(a_off, 8);
        b_mnW(j.seg[0], 16384u * h, 128);
        j.nseg = 1; j.idesc = make_idesc_f16(128, 64, 0, 1);
        j.d_col = TC_WORK; j.wait_op = wait_op; j.w_acquire = first_use; j.w_release = last_use; j.commit = 1;
        return j;
    };
    auto wgrad = [&](uint32_t dz_off, uint32_t in_off, int N, int d_col, bool resident, bool wait_op, bool commit) {
        Job j{};
        j.seg[0] = seg_mnA(dz_off, 8);
        b_mnS(j.seg[0], in_off);
        j.nseg = 1; j.idesc = make_idesc_f16(128, N, 1, 1);
        j.d_col = (uint16_t)d_col; j.resident = resident; j.wait_op = wait_op; j.commit = commit;
        return j;
    };
    // forward recompute
    add(fwd_half(SM_X, xs, 0, 0, true, false));       add(fwd_half(SM_X, xs, 0, 1, false, true));
    add(fwd_half(SM_SLOT_A, 8, 1, 0, true, false));   add(fwd_half(SM_SLOT_A, 8, 1, 1, false, true));
    add(fwd_half(SM_SLOT_B, 8, 2, 0, true, false));   add(fwd_half(SM_SLOT_B, 8, 2, 1, false, true));
    add(fwd_half(SM_SLOT_C, 8, 1, 0, true, false));   add(fwd_half(SM_SLOT_C, 8, 1, 1, false, true));
    {   // heads: [H3 | ONES] x WH (R = 16) -> TEMP[0..15]
        Job j{};
        j.seg[0] = seg_kmajorA(SM_SLOT_A, 8); b_kmajorW(j.seg[0], 0, 16);
        j.seg[1] = seg_kmajorA(SM_ONES, 1);   b_kmajorW(j.seg[1], 16u * 256u, 16);
        j.nseg = 2; j.idesc = make_idesc_f16(128, 16, 0, 0); j.d_col = TC_TEMP; j.wait_op = 1; j.w_acquire = 1; j.commit = 1;
        add(j);
    }
    add(wgrad(SM_SLOT_A, SM_DZH, 16, TC_TEMP + 16, false, true, false));     // head wgrad: H3^T dZh
    for (int h = 0; h < 2; ++h) {   // head dgrad: dZh x WH^T
        Job j{};
        j.seg[0] = seg_kmajorA(SM_DZH, 1);
        j.seg[0].b_off = 2048u * h; j.seg[0].b_lbo = 128 >> 4; j.seg[0].b_sbo = 256 >> 4; j.seg[0].b_adv = 256 >> 4; j.seg[0].b_w = 1;
        j.nseg = 1; j.idesc = make_idesc_f16(128, 64, 0, 1); j.d_col = TC_WORK; j.wait_op = (h == 1); j.w_release = (h == 1); j.commit = 1;
        add(j);
    }
    // layer 3
    add(wgrad(SM_SLOT_A, SM_SLOT_C, 128, TC_DW3, true, true, false));
    add(wgrad(SM_SLOT_A, SM_ONES, 16, TC_TEMP + 32, false, false, false));
    add(dgrad_half(SM_SLOT_A, 0, false, true, false));  add(dgrad_half(SM_SLOT_A, 1, true, false, true));
    // layer 2
    add(wgrad(SM_SLOT_C, SM_SLOT_B, 128, TC_DW2, true, true, false));
    add(wgrad(SM_SLOT_C, SM_X, Kx, TC_DW2 + 128, true, false, false));
    add(dgrad_half(SM_SLOT_C, 0, false, true, false));  add(dgrad_half(SM_SLOT_C, 1, true, false, true));
    // recompute H0
    add(fwd_half(SM_X, xs, 0, 0, true, false));       add(fwd_half(SM_X, xs, 0, 1, false, true));
    // layer 1
    add(wgrad(SM_SLOT_B, SM_SLOT_A, 128, TC_WORK, false, true, true));
    add(wgrad(SM_SLOT_B, SM_ONES, 16, TC_TEMP + 48, false, true, false));
    add(dgrad_half(SM_SLOT_B, 0, false, true, false));  add(dgrad_half(SM_SLOT_B, 1, true, false, true));
    // layer 0
    add(wgrad(SM_SLOT_A, SM_X, Kx, TC_DW0, true, true, true));
    tp.njobs = n;
    return n <= MAX_JOBS;
}

int fused_train(tnerf_handle* h, const RaySource& rs, long long n, float nr, float fr, int S, const float* jitter, int white,
                const float* target, float loss_denom, const float* gC, const float* gD, const float* gA, const float* gW,
                float grad_scale, const float* grad_scale_dev, float* comp, float* loss_sum, float* grads, cudaStream_t s) {
    if (n <= 0) return 0;
    FusedPlan fp;
    if (!build_plan(h, fp)) { set_error("fused train: unsupported MLP shape"); return -2; }
    static thread_local TrainPlan tp;
    if (!build_train_plan(fp, tp)) { set_error("fused train: tensor-core backward needs depth=4, skip_at=2, hidden=128"); return -2; }
    if (!h->packed) { set_error("fused train: tnerf_pack_weights has not been called"); return -3; }
    if (S < 1 || S > 128 || 128 % S) { set_error("fused train: n_samples must divide 128"); return -4; }
    if (gW) { set_error("fused train: upstream gradient w.r.t. weights is not supported on the fused path"); return -5; }
    TrainParams p{};
    p.rs = rs; p.n_rays = n; p.S = S; p.R = 128 / S; p.n_tiles = (n + p.R - 1) / p.R; p.white = white; p.Kx = fp.Kx; p.L = fp.L;
    p.include_input = fp.include_input; p.near_ = nr; p.far_ = fr; p.jitter = jitter; p.target = target; p.gC = gC; p.gD = gD; p.gA = gA;
    p.comp = comp; p.loss_sum = loss_sum; p.image = reinterpret_cast<const __half*>(h->packed);
    p.b1 = h->params[3]; p.b3 = h->params[7]; p.b_sigma = h->params[9]; p.b_rgb = h->params[11];
    p.debug = reinterpret_cast<long long*>(h->debug);
    if (target) {
        p.inv_denom = 1.f / loss_denom;
        // power-of-two loss scale keeping dZ inside fp16's normal range: gC*scale = 2 (C-t) * 2^k/denom with 2^k/denom in [64,128)
        p.scale = exp2f(ceilf(log2f(loss_denom)) + 6.f);
    } else {
        p.inv_denom = 0.f;
        p.scale = grad_scale > 0.f ? grad_scale : 1.f;
        p.scale_dev = grad_scale_dev;
    }
    const int Kx = fp.Kx;
    SlabMap& sm = p.sm;
    int off = 0;
    sm.dw0 = off; off += Kx * 128;
    sm.dw1 = off; off += 128 * 128;
    sm.dw2 = off; off += (128 + Kx) * 128;
    sm.dw3 = off; off += 128 * 128;
    sm.dwh = off; off += 4 * 128;
    sm.db1 = off; off += 128;
    sm.db3 = off; off += 128;
    sm.hb = off; off += 16;
    sm.total = off;
    const size_t need = (size_t)h->sm_count * sm.total * sizeof(float);
    if (h->slab_bytes < need) {
        if (h->slabs) cudaFree(h->slabs);
        cudaError_t e = cudaMalloc(&h->slabs, need);
        if (e != cudaSuccess) { set_error("cudaMalloc(gradient slabs) failed"); h->slabs = nullptr; h->slab_bytes = 0; return (int)e; }
        h->slab_bytes = need;
        cudaMemsetAsync(h->slabs, 0, need, s);      // slab 0 doubles as the bulk-reduction target: zero on entry, cleared by the scatter kernel
        h->slab0_zero = true;
    }
    p.slabs = reinterpret_cast<float*>(h->slabs);
    long long grid;
    static const bool force_v1 = getenv("TNERF_TRAIN_V1") != nullptr;
    // Default mode: streams in phase + gradients added into ONE vector by bulk async reductions (fastest; fp32 sums differ in the last
    // bits between runs).  TNERF_TRAIN_SYNC=0 = the run-to-run reproducible mode: streams half a tile apart, per-CTA slabs summed in a
    // fixed order.  TNERF_BULK_REDUCE=0/1 overrides the flush alone.
    const char* sy_env = getenv("TNERF_TRAIN_SYNC");
    const char* bk_env = getenv("TNERF_BULK_REDUCE");
    const int sync = sy_env ? (sy_env[0] == '1') : 1;
    const bool bulk = bk_env ? atoi(bk_env) != 0 : (sync != 0);
    bool used_bulk = false;
    if ((64 % S == 0 || S == 128) && !force_v1) {
        // two-stream kernel (tnerf_train2.cu): 64-sample tiles; at 128 samples the two streams of a CTA carry the two halves of one ray
        if (S == 128) { p.R = 1; p.n_tiles = 2 * n; }
        else { p.R = 64 / S; p.n_tiles = (n + p.R - 1) / p.R; }
        grid = (p.n_tiles + 1) / 2 < h->sm_count ? (p.n_tiles + 1) / 2 : h->sm_count;
        p.bulk_reduce = bulk ? 1 : 0;
        p.sync_streams = sync;
        used_bulk = bulk;
        if (bulk && !h->slab0_zero) cudaMemsetAsync(h->slabs, 0, (size_t)sm.total * sizeof(float), s);
        if (int rc = fused_train2(h, fp, p, Kx, (int)grid, s)) return rc;
    } else {
        grid = p.n_tiles < h->sm_count ? p.n_tiles : h->sm_count;
        const size_t smem = SM_MISC + sizeof(TrainMisc);
        auto kern = Kx == 64 ? fused_train_kernel<64> : Kx == 48 ? fused_train_kernel<48> : Kx == 32 ? fused_train_kernel<32> : fused_train_kernel<16>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("fused train: shared memory request rejected"); return (int)e; }
        kern<<<(unsigned)grid, TR_THREADS, smem, s>>>(p, tp);
        if (int rc = count_launch()) return rc;
    }
    h->slab0_zero = used_bulk;          // the scatter kernel leaves slab 0 cleared in bulk mode; otherwise it holds a CTA's partial sums
    ReduceArgs ra{};
    ra.slabs = p.slabs; ra.n_slabs = used_bulk ? 1 : (int)grid; ra.zero_after = used_bulk ? 1 : 0; ra.sm = sm; ra.D = fp.D; ra.Kx = Kx; ra.inv_scale = 1.f / p.scale; ra.scale_dev = p.scale_dev; ra.grads = grads;
    for (int l = 0; l < 4; ++l) { ra.off_w[l] = h->offsets[2 * l]; ra.off_b[l] = h->offsets[2 * l + 1]; }
    ra.off_ws = h->offsets[8]; ra.off_bs = h->offsets[9]; ra.off_wc = h->offsets[10]; ra.off_bc = h->offsets[11];
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)((sm.total + 255) / 256)); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = (pdl_mask() & 2) ? 1 : 0;
    cudaLaunchKernelEx(&cfg, reduce_slabs_kernel, ra);
    return count_launch();
}

}  // namespace tnerf
