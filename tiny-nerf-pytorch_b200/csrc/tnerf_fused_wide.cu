// Fused forward (render) kernel for the wide MLP of BASELINE config 4: hidden = 256, depth 4, skip into layer 2
// (src/nerf.py:10-41 with hidden=256), n_samples % 32 == 0 (C4: 192).
//
// The fp16 weights of this model are 458 KB: they do not fit one SM's shared memory, and streaming them per 128-sample tile
// would need ~64 B/cycle/SM from L2 (the chip's bulk-copy limit is ~43).  So the kernel runs on CTA PAIRS
// (`tcgen05.mma.cta_group::2`, cluster of two SMs of one TPC): one instruction computes D[256 x N] where each CTA of the pair
// owns 128 rows (its own 128-sample tile: A operand = its activations in ITS tensor memory) and supplies HALF of the B columns
// from ITS shared memory.  Each SM therefore keeps only half of every weight matrix resident (229 KB), nothing is streamed,
// and both tensor cores run at the full M=128 rate.
//
//   * the 256 output features of a layer are produced in four QUARTERS of 64 (N=64 instructions; per CTA a [K x 32] slice of
//     the weights); accumulators ping-pong between two 64-column buffers, so the epilogue of quarter q (tensor memory ->
//     +bias -> relu -> fp16 -> tensor memory) runs under the MMAs of quarter q+1;
//   * activations ping-pong between two 128-column fp16 buffers P/Q (layer l reads one, its epilogue fills the other);
//     the first twelve K-steps of a layer's first quarter only need quarters 0-2 of the previous layer and are issued while
//     quarter 3 is still in its epilogue: the tensor pipe never waits for a whole layer to drain;
//   * biases of layers 0/2 ride on the encoding's constant-1 column, those of layers 1/3 are added in fp32 by the epilogue;
//   * the two heads (sigma, rgb: 4 x 256) are one more pair-MMA (N = 32, 4 useful columns).  Only 3 KB of shared memory are
//     left beside the weights, so the head operand is stored with OVERLAPPING core matrices: 4 rows of 16 B per K-octet
//     (64 B apart instead of 128), the 8-row-group stride is 0.  Rows 4-7 / 8-15 of the operand then alias real head weights
//     of other K-octets (finite garbage in accumulator columns nobody reads); 2.1 KB instead of 8 KB;
//   * the biases of layers 1/3 and of the heads sit in the CONSTANT bank (uniform-register operands of the epilogue's adds);
//   * rays, depths, Fourier features and compositing run in a separate sample warpgroup, as in tnerf_fused_fast.cu.
//
// Tensor memory per CTA (columns): ACC0 0-63, ACC1 64-127, P 128-255, Q 256-383, X 384-415, HEAD 416-447.
// Warps per CTA: 0-7 epilogue (warp w: lanes 32(w%4).., columns 32(w/4).. of each 64-column quarter), 8-11 samples,
// 12 weight loader + (leader CTA only) MMA issuer (13-15 idle: setmaxnreg is per warpgroup).
// Cross-CTA signalling: "operand ready" barriers live in the leader CTA and collect one arrival per warp of BOTH CTAs
// (remote mbarrier arrive); MMA completion is multicast to the same barrier in both CTAs by tcgen05.commit.
#include <mutex>
#include <type_traits>
#include "tnerf_fused.cuh"

namespace tnerf {
namespace wide {

constexpr int THREADS = 512;      // warps 13-15 exist only so that warp 12's warpgroup can take part in the register re-distribution
constexpr int C_ACC = 0, C_P = 128, C_Q = 256, C_X = 384, C_HEAD = 416;
constexpr uint32_t HEAD_BYTES = 2112;   // 32 K-octets x 64 B + the 64 B the last overlapping core matrix reaches into
constexpr int TAIL_FLOATS = 1540;
// b1[256], b3[256], head weights [256][4] (sigma, r, g, b; kept for tools, the kernel's heads are an MMA), head biases [4] of the model
// whose kernel runs next on this device:
// compile-time offsets, so every bias / head weight is a constant-bank OPERAND of an FADD / FFMA (no load instruction at all).
// The host re-uploads the table (stream ordered, behind an event of the last kernel that used it) when another handle or a
// re-packed model comes along.
__constant__ float c_tail[TAIL_FLOATS];
constexpr int MAX_CHUNKS = 16;

struct Params {
    RaySource rs;
    long long n_rays, n_units;
    int S, G, R, white;
    float near_, far_;
    const float* jitter;
    float *comp, *depth, *acc, *rays_d_out;
    const uint8_t* image;      // [2 CTA ranks][image_bytes] fp16 operand slices, then the fp32 tail
    uint32_t image_bytes;
    int L, include_input;
    long long* debug;
};

struct Smem {
    float part[2][MAX_CHUNKS][6];     // [unit parity][chunk] = {P, sum w r, sum w g, sum w b, sum w z, sum w}
    uint64_t bar_w, bar_x, bar_e[2], bar_acc[2], bar_xfree, bar_head, bar_hfree;
    uint32_t tmem_slot;
};

// byte offset of quarter q of layer l inside one CTA's image
template <int KX>
__host__ __device__ constexpr uint32_t w_off(int l, int q) {
    constexpr uint32_t s0 = KX * 64u, s1 = 16384u, s2 = (256u + KX) * 64u;
    return l == 0 ? q * s0 : l == 1 ? 4 * s0 + q * s1 : l == 2 ? 4 * (s0 + s1) + q * s2 : 4 * (s0 + s1 + s2) + q * s1;
}
template <int KX>
__host__ __device__ constexpr uint32_t main_bytes_of() { return 4u * (KX * 64u + 16384u + (256u + KX) * 64u + 16384u); }
template <int KX>
__host__ __device__ constexpr uint32_t image_bytes_of() { return main_bytes_of<KX>() + HEAD_BYTES; }

// ---- pair / cluster primitives ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t smem_slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_slot), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[256 x N] (+)= A[tensor memory of each CTA, 128 rows] * B[shared memory of each CTA, N/2 columns]
__device__ __forceinline__ void mma_ts2(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate));
}
// arrive(1) on the barrier at this shared-memory offset in BOTH CTAs once all MMAs issued so far have completed
__device__ __forceinline__ void tc_commit2(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t)3)
                 : "memory");
}
// arrive on the barrier at the same offset in CTA `rank` of the pair
__device__ __forceinline__ void mbar_arrive_cta(uint32_t bar, uint32_t rank) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(bar),
        "r"(rank)
        : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}

// ---- weight image --------------------------------------------------------------------------------------------------------
// Per CTA rank r, layer l, quarter q: B slice [K x 32] of features 64q + 32r + n, K-major canonical (R = 32 rows):
// element (n, k) at ((k/8)*32 + n)*8 + k%8.  K order: layer 0 = encoding (constant-1 column KX-1 carries the bias);
// layer 2 = 256 activations, then the encoding (same bias convention); layers 1, 3 = 256 activations.
// Then the head operand: element (n < 4, k) at byte (k/8)*64 + n*16 + (k%8)*2 (n = 0 sigma, 1..3 rgb).
struct PackArgs {
    const float* W[4]; const float* b[4];
    const float *Wsig, *bsig, *Wrgb, *brgb;
    int D, KX;
    uint32_t image_bytes;
};
__global__ void pack_wide_kernel(PackArgs a, uint8_t* __restrict__ out) {
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long halfs = a.image_bytes / 2;
    if (e < 2 * halfs) {
        const int r = (int)(e / halfs);
        long long loc = e - r * halfs;
        const int KX = a.KX;
        const long long s0 = KX * 32, s1 = 8192, s2 = (256 + KX) * 32;
        const long long main_halfs = 4 * (s0 + s1 + s2 + s1);
        if (loc >= main_halfs) {
            const int byte = (int)(loc - main_halfs) * 2, ko = byte / 64, n = (byte % 64) / 16, k = ko * 8 + (byte % 16) / 2;
            float v = 0.f;
            if (ko < 32) v = n == 0 ? a.Wsig[k] : a.Wrgb[(n - 1) * 256 + k];
            reinterpret_cast<__half*>(out)[e] = __float2half_rn(v);
            return;
        }
        int l, q;
        if (loc < 4 * s0) { l = 0; q = (int)(loc / s0); loc -= q * s0; }
        else if ((loc -= 4 * s0) < 4 * s1) { l = 1; q = (int)(loc / s1); loc -= q * s1; }
        else if ((loc -= 4 * s1) < 4 * s2) { l = 2; q = (int)(loc / s2); loc -= q * s2; }
        else { loc -= 4 * s2; l = 3; q = (int)(loc / s1); loc -= q * s1; }
        const int k = (int)(loc / 256) * 8 + (int)(loc % 8);
        const int n = (int)((loc / 8) % 32);
        const int f = 64 * q + 32 * r + n;
        const int fan = l == 0 ? a.D : l == 2 ? 256 + a.D : 256;
        const float* W = a.W[l] + (long long)f * fan;
        float v = 0.f;
        if (l == 0) v = k < a.D ? W[k] : (k == KX - 1 ? a.b[0][f] : 0.f);
        else if (l == 2) {
            if (k < 256) v = W[k];
            else { const int kx = k - 256; v = kx < a.D ? W[256 + kx] : (kx == KX - 1 ? a.b[2][f] : 0.f); }
        } else v = W[k];
        reinterpret_cast<__half*>(out)[e] = __float2half_rn(v);
    } else {
        const long long t = e - 2 * halfs;
        float* tail = reinterpret_cast<float*>(out + 2ull * a.image_bytes);
        if (t < 256) tail[t] = a.b[1][t];
        else if (t < 512) tail[t] = a.b[3][t - 256];
        else if (t < 1536) { const int k = (int)(t - 512) / 4, c = (int)(t - 512) % 4; tail[t] = c == 0 ? a.Wsig[k] : a.Wrgb[(c - 1) * 256 + k]; }
        else if (t < 1540) { const int c = (int)(t - 1536); tail[t] = c == 0 ? a.bsig[0] : a.brgb[c - 1]; }
    }
}

// -DWIDE_DEBUG_WAITS: every barrier wait is bounded; a wait that times out records (site, parity) of the first failure of its
// warp at debug[512 + (blockIdx.x*9 + warp)*2] and FALLS THROUGH, so a protocol bug yields a finished kernel and a trace
// instead of a hang (results are garbage then).  Developer builds only.
#ifdef WIDE_DEBUG_WAITS
__device__ __forceinline__ void dbg_wait(uint32_t bar, uint32_t parity, long long* debug, int site, uint32_t& budget) {
    uint32_t n = 0;
    while (true) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return;
        if (++n > budget) break;
    }
    if (debug && (threadIdx.x & 31) == 0 && blockIdx.x < 16) {
        long long* slot = debug + 512 + (blockIdx.x * 13 + min((int)(threadIdx.x >> 5), 12)) * 2;
        if (slot[0] == 0) { slot[0] = site; slot[1] = parity; }
    }
    budget = 64;
}
#define WWAIT(bar, parity, site) dbg_wait(bar, parity, p.debug, site, wbudget)
#define WWAITC(bar, parity, site) dbg_wait(bar, parity, p.debug, site, wbudget)
#else
#define WWAIT(bar, parity, site) mbar_wait(bar, parity)
#define WWAITC(bar, parity, site) mbar_wait_cluster(bar, parity)
#endif

#define WSTAMP() do { if (dbg && dbg_n < 250) dbg[dbg_n++] = clock64(); } while (0)

struct PreIn {          // prefetched inputs of one sample row
    float o[3], d[3], u0, u1;
    long long ray;
    int si;
    bool valid;
};

template <int KX>
__global__ void __cluster_dims__(2, 1, 1) __maxnreg__(128) fused_fwd_wide_kernel(const __grid_constant__ Params p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr uint32_t IMG = image_bytes_of<KX>();
    Smem& sm = *reinterpret_cast<Smem*>(smem + IMG);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
#ifdef WIDE_DEBUG_WAITS
    uint32_t wbudget = 1u << 21;
#endif
    const uint32_t bar_w = smem_u32(&sm.bar_w), bar_x = smem_u32(&sm.bar_x), bar_xfree = smem_u32(&sm.bar_xfree),
                   bar_head = smem_u32(&sm.bar_head), bar_hfree = smem_u32(&sm.bar_hfree);
    const uint32_t bar_e0 = smem_u32(&sm.bar_e[0]), bar_e1 = smem_u32(&sm.bar_e[1]);
    const uint32_t bar_acc0 = smem_u32(&sm.bar_acc[0]), bar_acc1 = smem_u32(&sm.bar_acc[1]);

    if (warp == 12 && lane == 0) {
        mbar_init(bar_w, 1);
        mbar_init(bar_x, 8);            // one arrival per sample warp of both CTAs (leader's copy is the one waited on)
        mbar_init(bar_e0, 16); mbar_init(bar_e1, 16);       // one per epilogue warp of both CTAs
        mbar_init(bar_acc0, 1); mbar_init(bar_acc1, 1);
        mbar_init(bar_xfree, 1);
        mbar_init(bar_head, 1); mbar_init(bar_hfree, 8);    // head MMA commit; one arrival per sample warp of both CTAs
        fence_barrier_init();
    }
    if (warp == 0) { tmem_alloc2(smem_u32(&sm.tmem_slot), 512); tmem_relinquish2(); }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                 // barriers of both CTAs exist before any remote arrive / multicast commit
    tc_fence_after();
    const uint32_t tmem = sm.tmem_slot;

    const long long n_pairs = gridDim.x >> 1, pair = blockIdx.x >> 1;
    // unit (R whole rays = G tiles of 128 samples) j of this CTA: (j*n_pairs + pair)*2 + rank; both CTAs of a pair run the same
    // number of tiles (a CTA whose unit does not exist processes empty rows)
    const long long u_first = 2 * pair, u_stride = 2 * n_pairs;

    if (warp >= 12) {
        // ------------------------------ weight loader (both CTAs) + MMA issuer (leader): warp 12 ------------------------------
        TN_SETMAXNREG_DEC(56);           // executed by the whole warpgroup (warps 12-15)
        if (warp == 12) {
        if (lane == 0) {
            mbar_expect_tx(bar_w, IMG);
            const uint8_t* src = p.image + (size_t)rank * IMG;
            for (uint32_t off = 0; off < IMG; off += 32768u) bulk_g2s(smem_u32(smem) + off, src + off, min(32768u, IMG - off), bar_w);
        }
        __syncwarp();
        WWAIT(bar_w, 0, 1);
        // the peer's weights must be resident too before the first pair-MMA reads them: it arrives on the leader's bar_x only after
        // its own load (its sample warps wait for bar_w as well, below)
        if (rank == 0) {
            constexpr int XS = KX / 16;
            const uint32_t idesc = make_idesc_f16(256, 64, 0, 0);
            const uint32_t hi = (128u >> 4) | (1u << 14);                 // SBO = 128 B, descriptor version 1
            const uint32_t lbo = (512u >> 4) << 16;                       // LBO = 32 rows * 16 B
            const uint32_t tP = tmem + C_P, tQ = tmem + C_Q, tX = tmem + C_X;
            const uint32_t idesc_h = make_idesc_f16(256, 32, 0, 0);
            const uint32_t hi_h = 0u | (1u << 14);                        // SBO = 0: rows 8-15 of each CTA's head slice alias rows 0-7
            const uint32_t lbo_h = (64u >> 4) << 16;                      // LBO = 64 B: overlapping core matrices, 4 real rows each
            uint32_t ph_x = 0, ph_e0 = 0, ph_e1 = 0, ph_hf = 0;
            bool first_head = true;
#define W_TS(STEPS, d, a, blo, accum) do { _Pragma("unroll") for (int j_ = 0; j_ < (STEPS); ++j_) \
        mma_ts2(d, (a) + 8 * j_, ((uint64_t)hi << 32) | ((blo) + j_ * 64u), idesc, (j_ == 0) ? (accum) : 1u); } while (0)
#define WAIT_E0() do { WWAITC(bar_e0, ph_e0, 3); ph_e0 ^= 1; tc_fence_after(); } while (0)
#define WAIT_E1() do { WWAITC(bar_e1, ph_e1, 4); ph_e1 ^= 1; tc_fence_after(); } while (0)
            for (long long u = u_first; u < p.n_units; u += u_stride) {
                for (int g = 0; g < p.G; ++g) {
                    uint32_t sb = smem_u32(smem) >> 4;
                    asm volatile("" : "+r"(sb));          // descriptors are rebuilt per tile, not kept live across it
                    const uint32_t wb = sb + lbo;
                    const uint32_t A0 = tmem + C_ACC, A1 = tmem + C_ACC + 64;
                    // ---- layer 0: X -> P ----
                    WWAITC(bar_x, ph_x, 2); ph_x ^= 1; tc_fence_after();
                    if (elect_one()) { W_TS(XS, A0, tX, wb + (w_off<KX>(0, 0) >> 4), 0u); tc_commit2(bar_acc0); }
                    __syncwarp();
                    if (elect_one()) { W_TS(XS, A1, tX, wb + (w_off<KX>(0, 1) >> 4), 0u); tc_commit2(bar_acc1); }
                    __syncwarp();
                    WAIT_E0();
                    if (elect_one()) { W_TS(XS, A0, tX, wb + (w_off<KX>(0, 2) >> 4), 0u); tc_commit2(bar_acc0); }
                    __syncwarp();
                    WAIT_E1();
                    if (elect_one()) { W_TS(XS, A1, tX, wb + (w_off<KX>(0, 3) >> 4), 0u); tc_commit2(bar_acc1); }
                    __syncwarp();
                    // ---- layers 1..3: activations (previous layer) [+ X for the skip layer] ----
#pragma unroll
                    for (int l = 1; l <= 3; ++l) {
                        const uint32_t tA = (l == 2) ? tQ : tP;
                        // quarter 0: the first 12 K-steps read features 0-191 of the previous layer (its quarters 0-2)
                        WAIT_E0();
                        if (elect_one()) {
                            W_TS(12, A0, tA, wb + (w_off<KX>(l, 0) >> 4), 0u);
                            if (l == 2) W_TS(XS, A0, tX, wb + (w_off<KX>(l, 0) >> 4) + 16 * 64u, 1u);
                        }
                        __syncwarp();
                        WAIT_E1();
                        if (elect_one()) { W_TS(4, A0, tA + 96, wb + (w_off<KX>(l, 0) >> 4) + 12 * 64u, 1u); tc_commit2(bar_acc0); }
                        __syncwarp();
                        if (elect_one()) {
                            W_TS(16, A1, tA, wb + (w_off<KX>(l, 1) >> 4), 0u);
                            if (l == 2) W_TS(XS, A1, tX, wb + (w_off<KX>(l, 1) >> 4) + 16 * 64u, 1u);
                            tc_commit2(bar_acc1);
                        }
                        __syncwarp();
                        WAIT_E0();
                        if (elect_one()) {
                            W_TS(16, A0, tA, wb + (w_off<KX>(l, 2) >> 4), 0u);
                            if (l == 2) W_TS(XS, A0, tX, wb + (w_off<KX>(l, 2) >> 4) + 16 * 64u, 1u);
                            tc_commit2(bar_acc0);
                        }
                        __syncwarp();
                        WAIT_E1();
                        if (elect_one()) {
                            W_TS(16, A1, tA, wb + (w_off<KX>(l, 3) >> 4), 0u);
                            if (l == 2) W_TS(XS, A1, tX, wb + (w_off<KX>(l, 3) >> 4) + 16 * 64u, 1u);
                            tc_commit2(bar_acc1);
                            if (l == 2) tc_commit2(bar_xfree);      // last reader of this tile's encoding
                        }
                        __syncwarp();
                    }
                    // ---- heads: Q (layer 3 activations) x [256 x 4] -> HEAD ----
                    WAIT_E0();
                    WAIT_E1();
                    if (!first_head) { WWAITC(bar_hfree, ph_hf, 5); ph_hf ^= 1; tc_fence_after(); }
                    first_head = false;
                    if (elect_one()) {
                        const uint32_t hb = sb + (main_bytes_of<KX>() >> 4) + lbo_h;
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            mma_ts2(tmem + C_HEAD, tQ + 8 * j, ((uint64_t)hi_h << 32) | (hb + j * 8u), idesc_h, j ? 1u : 0u);
                        tc_commit2(bar_head);
                    }
                    __syncwarp();
                }
            }
#undef W_TS
#undef WAIT_E0
#undef WAIT_E1
        }
        }
    } else if (warp < 8) {
        // ------------------------------ epilogue warps: lanes 32(w%4).., columns 32(w/4).. of every 64-column quarter ------------------------------
        TN_SETMAXNREG_DEC(112);      // register pool = 128 x 512 threads: 8 x 112 + 4 x 168 + 4 x 56 <= 16 x 128
        auto epilogue = [&](auto CHC) {
        constexpr int ch = decltype(CHC)::value;       // column half: compile time, so the constant-bank offsets are immediates
        const int q4 = warp & 3;
        const uint32_t tw = tmem + ((uint32_t)(q4 * 32) << 16);
        uint32_t ph_acc0 = 0, ph_acc1 = 0;
        const float* ct = c_tail;
        long long* dbg = (p.debug && blockIdx.x == 0 && warp == 0 && lane == 0) ? p.debug + 256 : nullptr;
        int dbg_n = 0;
        for (long long u = u_first; u < p.n_units; u += u_stride) {
            for (int g = 0; g < p.G; ++g) {
#pragma unroll
                for (int l = 0; l < 4; ++l) {
                    const uint32_t dst = tw + ((l & 1) ? C_Q : C_P) + ch * 16;
                    const float* bias = ct + (l == 3 ? 256 : 0) + ch * 32;          // b1 / b3 (layers 0, 2: bias inside the GEMM)
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        WSTAMP();
                        if (q & 1) { WWAIT(bar_acc1, ph_acc1, 10 + l * 4 + q); ph_acc1 ^= 1; } else { WWAIT(bar_acc0, ph_acc0, 10 + l * 4 + q); ph_acc0 ^= 1; }
                        tc_fence_after();
                        WSTAMP();
                        uint32_t v[32];
                        tmem_ld32(tw + C_ACC + (q & 1) * 64 + ch * 32, v);
                        tc_wait_ld();
                        if (l & 1) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + bias[q * 64 + i]);
                        }
                        uint32_t h[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) h[i] = pack_relu_h2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
                        tmem_st16(dst + q * 32, h);
                        tc_wait_st();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cta((q & 1) ? bar_e1 : bar_e0, 0);     // accumulator buffer drained + activations stored
                    }
                }
            }
        }
        };
        if (warp < 4) epilogue(std::integral_constant<int, 0>{}); else epilogue(std::integral_constant<int, 1>{});
    } else {
        // ------------------------------ sample warpgroup ------------------------------
        TN_SETMAXNREG_INC(168);
        const int q = warp & 3, row = q * 32 + lane;
        const uint32_t tw = tmem + ((uint32_t)(q * 32) << 16);
        const int S = p.S, cpr = S >> 5;                 // chunks (warps) per ray
        const bool jit = p.jitter != nullptr;
        const bool camera = p.rs.rays_d == nullptr;
        float cam[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) cam[i] = camera ? p.rs.c2w[i] : 0.f;
        const float lin_step = (S > 1) ? __fdiv_rn(1.f, (float)(S - 1)) : 0.f;
        const float inv_focal = camera ? __frcp_rn(p.rs.focal) : 0.f, half_w = (float)p.rs.W * 0.5f, half_h = (float)p.rs.H * 0.5f;
        const float near_ = p.near_, far_ = p.far_;
        long long* dbg = (p.debug && blockIdx.x == 0 && warp == 8 && lane == 0) ? p.debug : nullptr;
        int dbg_n = 0;

        auto bin = [&](int i) -> float {      // bit-exact torch.linspace / z formula (src/sampling.py:16-17)
            const float t = (S <= 1) ? 0.f : ((i < S / 2) ? __fmul_rn(lin_step, (float)i) : __fmaf_rn(-lin_step, (float)(S - 1 - i), 1.f));
            return __fadd_rn(__fmul_rn(near_, __fsub_rn(1.f, t)), __fmul_rn(far_, t));
        };
        auto zsample = [&](int i, float uu) -> float {
            const float zc = bin(i);
            if (!jit) return zc;
            const float lo = (i == 0) ? zc : __fmul_rn(0.5f, __fadd_rn(bin(i - 1), zc));
            const float hi = (i == S - 1) ? zc : __fmul_rn(0.5f, __fadd_rn(zc, bin(i + 1)));
            return __fadd_rn(lo, __fmul_rn(__fsub_rn(hi, lo), uu));
        };
        auto prefetch = [&](long long u, int g) -> PreIn {
            PreIn in;
            const int chunk = g * 4 + q;
            in.ray = u * p.R + chunk / cpr;
            in.si = (chunk % cpr) * 32 + lane;
            in.valid = u < p.n_units && in.ray < p.n_rays;
            in.u0 = in.u1 = 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) { in.o[c] = 0.f; in.d[c] = 0.f; }
            if (in.valid) {
                if (jit) {
                    in.u0 = p.jitter[in.ray * S + in.si];
                    in.u1 = (in.si + 1 < S) ? p.jitter[in.ray * S + in.si + 1] : 0.f;
                }
                if (!camera) {
                    const float* po = p.rs.rays_o + p.rs.o_stride * in.ray;
#pragma unroll
                    for (int c = 0; c < 3; ++c) { in.d[c] = p.rs.rays_d[3 * in.ray + c]; in.o[c] = po[c]; }
                } else {
                    float cm[12];
                    unsigned local = (unsigned)in.ray;
                    if (p.rs.frame_rays) {        // pose batch: this row's frame and its camera (uniform per warp: a warp = 32 samples of one ray)
                        const unsigned fr = (unsigned)in.ray / (unsigned)p.rs.frame_rays;
                        local = (unsigned)in.ray - fr * (unsigned)p.rs.frame_rays;
                        const float4* cp = reinterpret_cast<const float4*>(p.rs.c2w + 16 * fr);
                        const float4 r0 = __ldg(cp), r1 = __ldg(cp + 1), r2 = __ldg(cp + 2);
                        cm[0] = r0.x; cm[1] = r0.y; cm[2] = r0.z; cm[3] = r0.w; cm[4] = r1.x; cm[5] = r1.y; cm[6] = r1.z; cm[7] = r1.w;
                        cm[8] = r2.x; cm[9] = r2.y; cm[10] = r2.z; cm[11] = r2.w;
                    } else {
#pragma unroll
                        for (int i = 0; i < 12; ++i) cm[i] = cam[i];
                    }
                    const long long k = p.rs.pixel_index ? p.rs.pixel_index[in.ray] : p.rs.first_ray + local;
                    const unsigned kk = (unsigned)k, Wd = (unsigned)p.rs.W;
                    const unsigned prow = kk / Wd, pcol = kk - prow * Wd;
                    // src/rays.py:21-31 with the divisions turned into multiplications by once-computed reciprocals
                    // (<= 3 ulp on the direction, inside the 1e-6 bar of the stand-alone get_rays kernel)
                    const float cx = ((float)pcol - half_w) * inv_focal;
                    const float cy = -((float)prow - half_h) * inv_focal;
                    const float wx = fmaf(-1.f, cm[2], fmaf(cy, cm[1], cx * cm[0]));
                    const float wy = fmaf(-1.f, cm[6], fmaf(cy, cm[5], cx * cm[4]));
                    const float wz = fmaf(-1.f, cm[10], fmaf(cy, cm[9], cx * cm[8]));
                    const float inv_n = rsqrtf(fmaxf(fmaf(wz, wz, fmaf(wy, wy, wx * wx)), 1e-24f));
                    in.d[0] = wx * inv_n; in.d[1] = wy * inv_n; in.d[2] = wz * inv_n;
                    in.o[0] = cm[3]; in.o[1] = cm[7]; in.o[2] = cm[11];
                }
            }
            return in;
        };
        uint32_t pk[KX / 2];
        // depth, point, features (registers); returns z and gap*|d| of this sample (src/volume.py:18-23)
        auto encode = [&](const PreIn& in, float& z_out, float& gapdn_out) {
            float pt[3] = {0.f, 0.f, 0.f};
            float z = 0.f, gd = 0.f;
            if (in.valid) {
                z = zsample(in.si, in.u0);
                const float znext = (in.si == S - 1) ? 0.f : zsample(in.si + 1, in.u1);
                const float dn = sqrtf(in.d[0] * in.d[0] + in.d[1] * in.d[1] + in.d[2] * in.d[2]);
                gd = ((in.si == S - 1) ? kLastDelta : (znext - z)) * dn;
#pragma unroll
                for (int c = 0; c < 3; ++c) pt[c] = __fadd_rn(in.o[c], __fmul_rn(in.d[c], z));
                if (p.rays_d_out && in.si == 0) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) p.rays_d_out[3 * in.ray + c] = in.d[c];
                }
            }
            z_out = z; gapdn_out = gd;
            if (p.include_input) encode_stream<KX, true>(pt, p.L, pk); else encode_stream<KX, false>(pt, p.L, pk);
        };
        auto store_x = [&]() {          // features -> tensor memory (A operand of layer 0 and of the skip layer)
#pragma unroll
            for (int c = 0; c < KX / 32; ++c) {
                uint32_t chunk[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) chunk[i] = pk[c * 16 + i];
                tmem_st16(tw + C_X + c * 16, chunk);
            }
            if (KX % 32) {
                uint32_t chunk[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) chunk[i] = pk[(KX / 32) * 16 + i];
                tmem_st8(tw + C_X + (KX / 32) * 16, chunk);
            }
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cta(bar_x, 0);
        };

        WWAIT(bar_w, 0, 40);             // this CTA's weight slices are resident before it reports its first tile ready
        uint32_t ph_head = 0, ph_xfree = 0;
        int parity = 0;
        // three tiles in flight on these threads, as in tnerf_fused_fast.cu: tile t (in the layers, composited when its heads
        // arrive), tile t+1 (features stored once tile t's last reader of the encoding is done), tile t+2 (in registers)
        struct Samp { float z, gd; long long ray; int si; bool valid; };
        auto next_tile = [&](long long& u_, int& g_) { if (++g_ == p.G) { g_ = 0; u_ += u_stride; } };
        // the loop bounds follow the LEADER's unit (u_first); this CTA's own unit is u + rank
        long long u = u_first;
        int g = 0;
        if (u < p.n_units) {
            Samp cur, nxt, nn;
            {
                const PreIn in = prefetch(u + rank, g);
                encode(in, cur.z, cur.gd);
                cur.ray = in.ray; cur.si = in.si; cur.valid = in.valid;
                store_x();
            }
            long long u1 = u; int g1 = g; next_tile(u1, g1);
            bool has1 = u1 < p.n_units;
            nxt = cur;
            if (has1) {
                const PreIn in = prefetch(u1 + rank, g1);
                encode(in, nxt.z, nxt.gd);
                nxt.ray = in.ray; nxt.si = in.si; nxt.valid = in.valid;
            }
            while (true) {
                long long u2 = u1; int g2 = g1; next_tile(u2, g2);
                const bool has2 = has1 && u2 < p.n_units;
                WSTAMP();
                WWAIT(bar_xfree, ph_xfree, 41);          // tile t's last reader of the encoding has completed
                ph_xfree ^= 1;
                tc_fence_after();
                if (has1) store_x();                     // features of tile t+1
                WSTAMP();
                nn = nxt;
                if (has2) {
                    const PreIn in = prefetch(u2 + rank, g2);
                    encode(in, nn.z, nn.gd);             // features of tile t+2 stay in registers
                    nn.ray = in.ray; nn.si = in.si; nn.valid = in.valid;
                }
                WSTAMP();
                WWAIT(bar_head, ph_head, 42);
                ph_head ^= 1;
                tc_fence_after();
                WSTAMP();
                uint32_t hv[4];
                tmem_ld4(tw + C_HEAD, hv);
                tc_wait_ld();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cta(bar_hfree, 0);
                const float sigma = fmaxf(__uint_as_float(hv[0]) + c_tail[1536], 0.f);
                const float cr = __fdividef(1.f, 1.f + __expf(-(__uint_as_float(hv[1]) + c_tail[1537])));
                const float cg = __fdividef(1.f, 1.f + __expf(-(__uint_as_float(hv[2]) + c_tail[1538])));
                const float cb = __fdividef(1.f, 1.f + __expf(-(__uint_as_float(hv[3]) + c_tail[1539])));
                // chunk-local compositing (this warp = 32 consecutive samples of one ray), src/volume.py:26-41
                const float alpha = cur.valid ? 1.f - __expf(-sigma * cur.gd) : 0.f;
                const float qv = 1.f - alpha + kEpsT;
                float incl = qv;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const float up = __shfl_up_sync(0xffffffffu, incl, off);
                    if (lane >= off) incl *= up;
                }
                float excl = __shfl_up_sync(0xffffffffu, incl, 1);
                if (lane == 0) excl = 1.f;
                const float wl = alpha * excl;
                const float s0 = warp_sum(wl * cr), s1 = warp_sum(wl * cg), s2 = warp_sum(wl * cb), s3 = warp_sum(wl * cur.z), s4 = warp_sum(wl);
                const float P = __shfl_sync(0xffffffffu, incl, 31);
                const int chunk = g * 4 + q;
                if (lane == 0) {
                    float* pp = sm.part[parity][chunk];
                    pp[0] = P; pp[1] = s0; pp[2] = s1; pp[3] = s2; pp[4] = s3; pp[5] = s4;
                }
                if (g == p.G - 1) {
                    bar_sync(1, 128);
                    if (row < p.R) {
                        const long long ray = (u + rank) * p.R + row;
                        if (u + rank < p.n_units && ray < p.n_rays) {
                            float T = 1.f, C0 = 0.f, C1 = 0.f, C2 = 0.f, Dd = 0.f, A = 0.f;
                            for (int c = 0; c < cpr; ++c) {
                                const float* pp = sm.part[parity][row * cpr + c];
                                C0 = fmaf(T, pp[1], C0); C1 = fmaf(T, pp[2], C1); C2 = fmaf(T, pp[3], C2);
                                Dd = fmaf(T, pp[4], Dd); A = fmaf(T, pp[5], A);
                                T *= pp[0];
                            }
                            const float bg = p.white ? 1.f - A : 0.f;
                            p.comp[3 * ray] = C0 + bg; p.comp[3 * ray + 1] = C1 + bg; p.comp[3 * ray + 2] = C2 + bg;
                            if (p.depth) p.depth[ray] = Dd;
                            if (p.acc) p.acc[ray] = A;
                        }
                    }
                    parity ^= 1;
                }
                WSTAMP();
                if (!has1) break;
                u = u1; g = g1; u1 = u2; g1 = g2; has1 = has2; cur = nxt; nxt = nn;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                 // the peer may still be reading this CTA's barriers / tensor memory through the pair MMAs
    if (warp == 0) tmem_dealloc2(tmem, 512);
}

static int kx_of(const tnerf_handle* h, int& L, int& inc) {
    if (h->num_freqs >= 0 && h->in_dim == 6 * h->num_freqs + 3) { L = h->num_freqs; inc = 1; }
    else if (h->num_freqs >= 0 && h->in_dim == 6 * h->num_freqs) { L = h->num_freqs; inc = 0; }
    else return 0;
    if (L < 0 || L > 10) return 0;
    const int kx = (h->in_dim + 1 + 15) / 16 * 16;
    return kx <= 64 ? kx : 0;
}

}  // namespace wide

// The constant table belongs to one (handle, pack version) at a time per device.  A launch for another owner first waits (on its
// stream) for the last kernel that read the table, then uploads its own 6 KB: stream ordered, no host synchronisation.
struct TableState {
    const tnerf_handle* owner = nullptr; long long version = -1;
    cudaEvent_t last_use = nullptr;       // last kernel that read the table (a new owner's upload waits for it)
    cudaEvent_t uploaded = nullptr;       // the current owner's upload; a launch on ANOTHER stream waits for it (same owner, same version)
    cudaStream_t upload_stream = nullptr;
};
static std::mutex g_table_mu;
static TableState g_table[64];
void wide_release(tnerf_handle* h) {
    std::lock_guard<std::mutex> lk(g_table_mu);
    for (TableState& t : g_table) if (t.owner == h) { t.owner = nullptr; t.version = -1; }
}

bool wide_shape_supported(const tnerf_handle* h) {
    int L, inc;
    return h->hidden == 256 && h->depth == 4 && h->skip_at == 2 && wide::kx_of(h, L, inc) > 0;
}

static uint32_t wide_image_bytes(int kx) { return 4u * ((uint32_t)kx * 64u + 16384u + (256u + (uint32_t)kx) * 64u + 16384u) + wide::HEAD_BYTES; }

int wide_pack_weights(tnerf_handle* h, cudaStream_t s) {
    int L, inc;
    const int kx = wide::kx_of(h, L, inc);
    if (!wide_shape_supported(h) || !kx) { set_error("wide fused path: unsupported MLP shape (need hidden=256, depth=4, skip_at=2, in_dim=6L(+3)<=63)"); return -2; }
    if (h->params.empty()) { set_error("pack_weights: parameters not bound"); return -3; }
    const uint32_t img = wide_image_bytes(kx);
    const size_t need = 2ull * img + wide::TAIL_FLOATS * sizeof(float);
    if (h->packed_bytes < need) {
        if (h->packed) cudaFree(h->packed);
        cudaError_t e = cudaMalloc(&h->packed, need);
        if (e != cudaSuccess) { set_error("cudaMalloc(packed image) failed"); return (int)e; }
        h->packed_bytes = need;
    }
    wide::PackArgs a{};
    for (int l = 0; l < 4; ++l) { a.W[l] = h->params[2 * l]; a.b[l] = h->params[2 * l + 1]; }
    a.Wsig = h->params[8]; a.bsig = h->params[9]; a.Wrgb = h->params[10]; a.brgb = h->params[11];
    a.D = h->in_dim; a.KX = kx; a.image_bytes = img;
    const long long total = (long long)img + wide::TAIL_FLOATS;      // img halfs per rank * 2 ranks = img elements, then the tail
    wide::pack_wide_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(a, reinterpret_cast<uint8_t*>(h->packed));
    ++h->wide_version;
    return count_launch();
}

int fused_render_fwd_wide(tnerf_handle* h, const RaySource& rs, long long n, float nr, float fr, int S, const float* jitter, int white,
                          float* comp, float* depth, float* acc, float* weights, float* rays_d_out, cudaStream_t s) {
    int L, inc;
    const int kx = wide::kx_of(h, L, inc);
    if (!wide_shape_supported(h) || !kx) { set_error("wide fused path: unsupported MLP shape"); return -2; }
    if (!h->packed || h->wide_version == 0) { set_error("fused path: tnerf_pack_weights has not been called"); return -3; }
    if (weights) { set_error("wide fused path: per-sample weights output is not available (use the fp32 path)"); return -4; }
    if (S < 32 || S % 32) { set_error("wide fused path: n_samples must be a multiple of 32"); return -4; }
    long long a = S, b = 128;
    while (b) { const long long t = a % b; a = b; b = t; }
    wide::Params p{};
    p.G = (int)(S / a); p.R = (int)(128 / a);
    if (p.G * 4 > wide::MAX_CHUNKS) { set_error("wide fused path: n_samples/gcd(n_samples,128) must be <= 4"); return -4; }
    p.rs = rs; p.n_rays = n; p.n_units = (n + p.R - 1) / p.R; p.S = S; p.white = white; p.near_ = nr; p.far_ = fr;
    p.jitter = jitter; p.comp = comp; p.depth = depth; p.acc = acc; p.rays_d_out = rays_d_out;
    p.image_bytes = wide_image_bytes(kx);
    p.image = reinterpret_cast<const uint8_t*>(h->packed);
    p.L = L; p.include_input = inc;
    p.debug = reinterpret_cast<long long*>(h->debug);
    const size_t smem = p.image_bytes + sizeof(wide::Smem);
    long long pairs = (p.n_units + 1) / 2;
    if (pairs > h->sm_count / 2) pairs = h->sm_count / 2;
    auto kern = kx == 64 ? wide::fused_fwd_wide_kernel<64> : kx == 48 ? wide::fused_fwd_wide_kernel<48>
              : kx == 32 ? wide::fused_fwd_wide_kernel<32> : wide::fused_fwd_wide_kernel<16>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("wide fused fwd: shared memory request rejected"); return (int)e; }
    std::lock_guard<std::mutex> lk(g_table_mu);
    TableState& t = g_table[h->device & 63];
    if (!t.last_use && cudaEventCreateWithFlags(&t.last_use, cudaEventDisableTiming) != cudaSuccess) { set_error("wide fused fwd: cudaEventCreate failed"); return -6; }
    if (!t.uploaded && cudaEventCreateWithFlags(&t.uploaded, cudaEventDisableTiming) != cudaSuccess) { set_error("wide fused fwd: cudaEventCreate failed"); return -6; }
    if (t.owner != h || t.version != h->wide_version) {
        if (t.owner) cudaStreamWaitEvent(s, t.last_use, 0);
        e = cudaMemcpyToSymbolAsync(wide::c_tail, p.image + 2ull * p.image_bytes, wide::TAIL_FLOATS * sizeof(float), 0, cudaMemcpyDeviceToDevice, s);
        if (e != cudaSuccess) { set_error("wide fused fwd: constant-bank upload failed"); return (int)e; }
        t.owner = h; t.version = h->wide_version;
        cudaEventRecord(t.uploaded, s);
        t.upload_stream = s;
    } else if (t.upload_stream != s) {
        cudaStreamWaitEvent(s, t.uploaded, 0);        // same table, other stream: the upload was only ordered on the stream that made it
    }
    kern<<<(unsigned)(2 * pairs), wide::THREADS, smem, s>>>(p);
    cudaEventRecord(t.last_use, s);
    return count_launch();
}

}  // namespace tnerf
