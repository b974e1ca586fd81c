// Host side of the fused training step (sm_100a): tnerf_train_fwd_bwd / tnerf_render_bwd on the tensor-core path =
// the two-stream kernel of tnerf_train2.cu (forward recompute + composite + loss gradient + full backward of
// src/train.py:114-126 in ONE launch) followed by the gradient scatter below.  Also here: the device-side choice of the loss
// scale for upstream-gradient calls and the overflow flag of the GradScaler semantics (src/train.py:81,126-128).
#include <cstdlib>
#include <vector>
#include "tnerf_train.cuh"

namespace tnerf {

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
    if (a.found && !isfinite(s)) *a.found = 1.f;       // GradScaler's found_inf (every writer stores the same value)
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

// Upstream-gradient calls (tnerf_render_bwd) without a caller-provided scale: a power of two that brings the largest upstream
// gradient to ~64, chosen on the device (no host synchronisation, one small launch instead of a dozen tensor operations).
__global__ void upstream_scale_kernel(const float* __restrict__ gC, const float* __restrict__ gD, const float* __restrict__ gA, long long n,
                                      float* __restrict__ scale_out) {
    float m = 0.f;
    for (long long i = threadIdx.x; i < 5 * n; i += blockDim.x) {
        float v = 0.f;
        if (i < 3 * n) { if (gC) v = gC[i]; }
        else if (i < 4 * n) { if (gD) v = gD[i - 3 * n]; }
        else if (gA) v = gA[i - 4 * n];
        m = fmaxf(m, fabsf(v));             // fmaxf drops NaNs: a non-finite upstream gradient surfaces in the gradients, not here
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    __shared__ float part[32];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = threadIdx.x < (blockDim.x >> 5) ? part[threadIdx.x] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (threadIdx.x == 0) {
            const float amax = fminf(fmaxf(m, 1e-30f), 3.0e38f);
            *scale_out = fminf(fmaxf(exp2f(floorf(log2f(64.f / amax))), 5.9604645e-8f), 1.0995116e12f);   // [2^-24, 2^40]
        }
    }
}

// non-finite check of the fp32 path's gradient vector (the tensor-core path raises the flag inside its own kernels)
__global__ void found_inf_kernel(const float* __restrict__ g, long long n, float* __restrict__ found) {
    bool bad = false;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) bad |= !isfinite(g[i]);
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) *found = 1.f;
}
int launch_found_inf(const float* g, long long n, float* found, cudaStream_t s) {
    if (n <= 0 || !found) return 0;
    unsigned nb = (unsigned)((n + 1023) / 1024); if (nb > 592) nb = 592;
    found_inf_kernel<<<nb, 256, 0, s>>>(g, n, found);
    return count_launch();
}

static void make_slab_map(int Kx, SlabMap& sm) {
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
}
// elements of the training kernel's sum vector for this model, and the gather plan (tnerf_sum_elems: callers that provide the vector)
int fused_train_sum_elems(tnerf_handle* h) {
    FusedPlan fp;
    if (!build_plan(h, fp) || fp.depth != 4 || fp.H != 128) return -1;
    SlabMap sm;
    make_slab_map(fp.Kx, sm);
    h->sum_total = sm.total;
    return sm.total;
}

// where every parameter's gradient sits in the tensor-memory-ordered sum vector: the inverse of reduce_slabs_kernel's scatter as a
// short list of blocks and vectors (tnerf_train_fwd_bwd with grads = NULL leaves the gradient there; the gathering optimiser launch
// transposes the weight blocks through shared memory so that both sides are coalesced)
static int build_gather_plan(tnerf_handle* h, const SlabMap& sm, int D, int Kx) {
    GatherPlan& g = h->gplan;
    if (g.valid && g.n == h->param_count) return 0;
    g = GatherPlan{};
    const long long* off = h->offsets.data();
    int tiles = 0;
    auto seg = [&](int sb, int C, int ld, long long pb) {
        g.seg[g.n_seg++] = GatherSeg{sb, C, ld, tiles, pb};
        tiles += ((C + 31) / 32) * 4;
    };
    seg(sm.dw0, D, D, off[0]);
    seg(sm.dw1, 128, 128, off[2]);
    seg(sm.dw2, 128, 128 + D, off[4]);
    seg(sm.dw2 + 128 * 128, D, 128 + D, off[4] + 128);
    seg(sm.dw3, 128, 128, off[6]);
    g.n_tiles = tiles;
    int first = 0;
    auto vec = [&](int sb, int n, long long pb) { g.vec[g.n_vec++] = GatherVec{sb, n, first, pb}; first += n; };
    vec(sm.dw0 + (Kx - 1) * 128, 128, off[1]);            // layer-0 bias = the encoding's constant-1 column
    vec(sm.db1, 128, off[3]);
    vec(sm.dw2 + (128 + Kx - 1) * 128, 128, off[5]);      // layer-2 bias likewise
    vec(sm.db3, 128, off[7]);
    vec(sm.dwh, 128, off[8]);                             // sigma weight
    vec(sm.dwh + 128, 384, off[10]);                      // rgb weight, [3][128]
    g.n_vec_elems = first;
    g.hb = sm.hb;
    g.pb_hb[0] = off[9]; g.pb_hb[1] = off[11]; g.pb_hb[2] = off[11] + 1; g.pb_hb[3] = off[11] + 2;
    long long covered = 4;
    for (int i = 0; i < g.n_seg; ++i) covered += 128LL * g.seg[i].C;
    for (int i = 0; i < g.n_vec; ++i) covered += g.vec[i].n;
    if (covered != h->param_count) { set_error("fused train: gather plan does not cover the parameter vector (unexpected layout)"); return -6; }
    g.n = h->param_count;
    g.valid = 1;
    return 0;
}

int fused_train(tnerf_handle* h, const RaySource& rs, long long n, float nr, float fr, int S, const float* jitter, int white,
                const float* target, float loss_denom, const float* gC, const float* gD, const float* gA, const float* gW,
                float grad_scale, const float* grad_scale_dev, float* found, float* comp, float* loss_sum, float* grads, cudaStream_t s) {
    if (n <= 0) return 0;
    FusedPlan fp;
    if (!build_plan(h, fp)) { set_error("fused train: unsupported MLP shape"); return -2; }
    if (fp.depth != 4 || fp.H != 128 || fp.layer[2].nseg != 2 || fp.layer[2].seg_kind[1] != SEG_X || fp.layer[1].seg_kind[1] != SEG_ONES ||
        fp.layer[3].seg_kind[1] != SEG_ONES) {
        set_error("fused train: tensor-core backward needs depth=4, skip_at=2, hidden=128");
        return -2;
    }
    if (!h->packed) { set_error("fused train: tnerf_pack_weights has not been called"); return -3; }
    if (S < 1 || S > 128 || 128 % S) { set_error("fused train: n_samples must divide 128"); return -4; }
    if (gW) { set_error("fused train: upstream gradient w.r.t. weights is not supported on the fused path"); return -5; }
    TrainParams p{};
    p.rs = rs; p.n_rays = n; p.S = S; p.white = white; p.Kx = fp.Kx; p.L = fp.L;
    p.include_input = fp.include_input; p.near_ = nr; p.far_ = fr; p.jitter = jitter; p.target = target; p.gC = gC; p.gD = gD; p.gA = gA;
    p.comp = comp; p.loss_sum = loss_sum; p.image = reinterpret_cast<const __half*>(h->packed);
    p.b1 = h->params[3]; p.b3 = h->params[7]; p.b_sigma = h->params[9]; p.b_rgb = h->params[11];
    p.debug = reinterpret_cast<long long*>(h->debug);
    p.found = found;
    p.scale_dev = grad_scale_dev;
    if (target) {
        p.inv_denom = 1.f / loss_denom;
        // power-of-two loss scale keeping dZ inside fp16's normal range: gC*scale = 2 (C-t) * 2^k/denom with 2^k/denom in [64,128)
        // (a caller-owned device scale -- the GradScaler state of tnerf_optimizer_step -- overrides it)
        p.scale = exp2f(ceilf(log2f(loss_denom)) + 6.f);
    } else {
        p.inv_denom = 0.f;
        p.scale = grad_scale > 0.f ? grad_scale : 1.f;
        if (!grad_scale_dev && !(grad_scale > 0.f)) {     // automatic: from the largest upstream gradient, on the device
            if (!h->auto_scale) {
                cudaError_t e = cudaMalloc(&h->auto_scale, sizeof(float));
                if (e != cudaSuccess) { set_error("cudaMalloc(loss scale) failed"); h->auto_scale = nullptr; return (int)e; }
            }
            upstream_scale_kernel<<<1, 1024, 0, s>>>(gC, gD, gA, n, h->auto_scale);
            if (int rc = count_launch()) return rc;
            p.scale_dev = h->auto_scale;
        }
    }
    const int Kx = fp.Kx;
    SlabMap& sm = p.sm;
    make_slab_map(Kx, sm);
    h->sum_total = sm.total;
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
    const bool ext = grads == nullptr && h->ext_sum != nullptr;      // caller-owned sum vector (zero on entry, cleared by the caller's exchange)
    if (ext) p.slabs = h->ext_sum;
    // Default mode: streams in phase + gradients added into ONE vector by bulk async reductions (fastest; fp32 sums differ in the last
    // bits between runs).  Option train_sync = 0 (tnerf_set_option / TNERF_TRAIN_SYNC at handle creation) = the run-to-run reproducible mode: streams half a tile apart, per-CTA slabs summed in a
    // fixed order.  Option bulk_reduce = 0/1 overrides the flush alone.
    const int sync = h->opt_train_sync >= 0 ? (h->opt_train_sync != 0) : 1;
    const bool bulk = h->opt_bulk_reduce >= 0 ? h->opt_bulk_reduce != 0 : (sync != 0);
    // two-stream kernel (tnerf_train2.cu): 64-sample tiles; at 128 samples the two streams of a CTA carry the two halves of one ray
    if (S == 128) { p.R = 1; p.n_tiles = 2 * n; }
    else { p.R = 64 / S; p.n_tiles = (n + p.R - 1) / p.R; }
    const long long grid = (p.n_tiles + 1) / 2 < h->sm_count ? (p.n_tiles + 1) / 2 : h->sm_count;
    p.bulk_reduce = bulk ? 1 : 0;
    p.sync_streams = sync;
    // measured SM speeds (tnerf_set_tile_order): the slowest SMs get the shorter allotments; not in the reproducible schedule
    p.tile_order = (sync && h->tile_order && h->tile_order_n == (int)grid) ? h->tile_order : nullptr;
    // grads == NULL: leave the (unscaled) sum in the ONE vector for the optimiser launch to gather from -- no scatter kernel
    const bool leave = grads == nullptr;
    if (leave) {
        if (!bulk) { set_error("fused train: grads = NULL needs the one-vector gradient flush (option bulk_reduce)"); return -7; }
        if (int rc = build_gather_plan(h, sm, fp.D, Kx)) return rc;
        p.unscale = 1;
    }
    if (!leave && h->slab_pending && !ext) { set_error("fused train: a gradient sum is pending (tnerf_train_fwd_bwd with grads = NULL): run the gathering tnerf_optimizer_step first"); return -8; }
    if (bulk && !ext && !h->slab0_zero && !h->slab_pending) cudaMemsetAsync(h->slabs, 0, (size_t)sm.total * sizeof(float), s);
    if (int rc = fused_train2(h, fp, p, Kx, (int)grid, s)) return rc;
    if (ext) return 0;
    if (leave) { h->slab0_zero = false; h->slab_pending = true; return 0; }     // pending: cleared element by element by the gathering optimiser launch
    h->slab0_zero = bulk;               // the scatter kernel leaves slab 0 cleared in bulk mode; otherwise it holds a CTA's partial sums
    ReduceArgs ra{};
    ra.slabs = p.slabs; ra.n_slabs = bulk ? 1 : (int)grid; ra.zero_after = bulk ? 1 : 0; ra.sm = sm; ra.D = fp.D; ra.Kx = Kx; ra.inv_scale = 1.f / p.scale; ra.scale_dev = p.scale_dev; ra.grads = grads;
    ra.found = found;
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
