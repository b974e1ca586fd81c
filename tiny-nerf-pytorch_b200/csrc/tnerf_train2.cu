// Fused training kernel (sm_100a): forward recompute + composite + loss gradient + full backward of src/train.py:114-126 in ONE
// launch.  Design (DESIGN.md sections 5.2 and 5.5):
//
//   * feature-major ("transposed") GEMMs: every layer is  D[feature x sample] = W[feature x k] . H[k x sample],
//     so the weights are the M=128 operand and a tile may hold any number of samples.  A tile is 64 samples
//     (N = 64), which halves the activation footprint and lets TWO independent tiles ("streams") be in flight
//     per CTA: while one stream's accumulator is being drained the other stream's GEMM owns the tensor pipe.
//   * the fp16 weight image (132 KB) is RESIDENT in shared memory (one bulk copy per CTA), nothing is streamed;
//   * per stream: X (8 KB) + two 16 KB activation slots; H1 is parked in registers of the drain threads between
//     its two uses, H0 is recomputed (K = 64), dZ_l overwrites H_l in place; layer 0 of the next tile is issued early from a
//     second copy of its features staged in the Q slot;
//   * biases of layers 1/3 are added in fp32 by the drain threads (thread <-> feature row), their gradients are
//     row sums of the same threads; layers 0/2 keep their bias in the constant-1 column of the encoding;
//   * tensor memory: dW3 (128 cols) + dW2 (128+Kx) + dW0 (Kx) stay resident, one 64-column accumulator per
//     stream; dW1 lives in registers, half of its columns in each drain warpgroup;
//   * gradient flush: bulk async reductions into ONE vector (divided by the loss scale on the way when the optimiser launch
//     gathers from it); dW3 / dW2 leave early, from the sample warps, while the last tile finishes.
//
// Warp roles (512 threads): warpgroup 0/1 = drain threads of stream 0/1 (thread <-> feature <-> TMEM lane);
// warps 8-9 / 10-11 = sample threads of stream 0/1 (rays, depths, jitter, Fourier features, compositing fwd+bwd, early flush);
// warp 12 = MMA issuer of BOTH streams (default schedule; warps 12/13 one stream each in the reproducible schedule); warp 14 lane 0
// loads the weights.
// Developer switches (compile time, tools/build_variant.sh): T2_SOLO (stream 1 idle: length of one stream's chain), T2_NO_MERGE (an
// issuer warp per stream in every schedule), T2_NO_EARLY (no early flush), T2_FLUSH_NBUF (staging chunks of the final flush).
#include <cstdlib>
#include <type_traits>
#include "tnerf_train.cuh"

namespace tnerf {
namespace t2 {

#ifndef T2_FLUSH_NBUF
#define T2_FLUSH_NBUF 6
#endif
constexpr int THREADS = 512;
constexpr int C_DW3 = 0, C_DW2 = 128, C_DW0 = 320, C_D = 384;   // tensor-memory columns; accumulator of stream s at C_D + 64 s
constexpr uint32_t S_W0 = 0, S_W1 = 16384, S_W2 = 49152, S_W3 = 98304, S_WH = 131072;
constexpr uint32_t S_P0 = 135168, S_Q0 = 151552, S_P1 = 167936, S_Q1 = 184320;   // order matters: see the head GEMM
constexpr uint32_t S_X0 = 200704, S_X1 = 208896, S_DZH0 = 217088, S_DZH1 = 219136, S_MISC = 221184;

struct Misc {
    float xch[2][48];          // cross-warp carries of the compositing scans (a ray spans 2 warps at 64 samples, all 4 at 128)
    uint64_t bar_w, bar_x[2], bar_in[2], bar_d[2], bar_head[2], bar_dzh[2], bar_g[2][2], bar_gfree[2], bar_xfree[2], bar_wg[2], bar_dread[2], bar_xq[2], bar_qfree[2], bar_fin[2];
    uint32_t tmem_slot;
};

__device__ __forceinline__ long long gtimer() { long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define T2_STAMP() do { if (dbg && dbg_n < 251) dbg[dbg_n++] = clock64(); } while (0)

struct WCopy { uint32_t src, dst, bytes; };
struct Extra { WCopy w[5]; };

// ---- operand descriptors --------------------------------------------------------------------------
// image of X(r, c), R rows: byte((c/8)*R + r)*16 + (c%8)*2.  "kmaj": r is the M/N index, c the reduction index;
// "mnmaj": c is the M/N index, r the reduction index.
struct Op { uint32_t lo, hi, adv; };
// sb16 = (shared-memory base) >> 4 (the base is 1024-byte aligned and below 256 KB, so the start-address field is a plain
// add of compile-time constants: the issuer re-materialises descriptors instead of keeping dozens of them in registers)
__device__ __forceinline__ Op kmaj(uint32_t sb16, uint32_t off, uint32_t R) { return {sb16 + ((off >> 4) + (R << 16)), 8u | (1u << 14), R * 2u}; }
__device__ __forceinline__ Op mnmaj(uint32_t sb16, uint32_t off, uint32_t R) { return {sb16 + ((off >> 4) + (8u << 16)), R | (1u << 14), 16u}; }

template <int STEPS>
__device__ __forceinline__ void gemm(uint32_t d, const Op a, const Op b, uint32_t idesc, uint32_t acc) {
    const uint32_t alo = a.lo, blo = b.lo;
#pragma unroll
    for (int j = 0; j < STEPS; ++j)
        mma_ss(d, ((uint64_t)a.hi << 32) | (alo + j * a.adv), ((uint64_t)b.hi << 32) | (blo + j * b.adv), idesc, j == 0 ? acc : 1u);
}

// ---- drains (thread <-> feature row f; accumulator columns = the 64 samples of the tile) ------------
__device__ __forceinline__ float h2sum(uint32_t h) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&h));
    return a.x + a.y;
}
// ReLU masks.  A drain thread remembers which of its row's 64 activations were positive as two 32-bit words (one per 32-sample
// half): pair k of a half (samples 2k, 2k+1 = one packed fp16x2 word) owns bits 15-k (low half) and 31-k (high half).  The
// backward drains expand them again instead of re-reading the activation slot: those 16 KB reads per layer competed with the
// operand fetches of the weight-gradient GEMMs for the shared-memory pipe (7 % of the step).
__device__ __forceinline__ uint32_t mask_collect(uint32_t acc, uint32_t h2, int k) {    // h2: two non-negative fp16 values
    const uint32_t t = h2 + 0x7FFF7FFFu;                 // bit 15 / 31 set iff the low / high half is non-zero (no carry between halves)
    return acc | ((t >> k) & (0x80008000u >> k));
}
__device__ __forceinline__ uint32_t mask_expand(uint32_t m, int k) {                    // 0xFFFF per half whose activation was positive
    return ((m >> (15 - k)) & 0x00010001u) * 0xFFFFu;
}
// backward drain, two-phase: dZ = dH * (H > 0); the masked, packed dZ row is computed into registers while the wgrad GEMM that still reads the slot
// (H as its operand) runs; drain_store() writes it once that GEMM has committed.  `dread_bar` (optional) is arrived on as
// soon as the accumulator has been read, so a GEMM that only needs the accumulator may be issued under the rest of the drain.
__device__ __forceinline__ void drain_bwd_compute(uint32_t D, uint32_t mask_lo, uint32_t mask_hi, uint32_t (&o)[32], uint32_t dread_bar) {
    uint32_t va[2][32];
    tmem_ld32(D, va[0]);
    tc_wait_ld();
    tmem_ld32(D + 32, va[1]);                 // second half in flight while the first is processed
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        if (c == 1) { tc_wait_ld(); if (dread_bar) { tc_fence_before(); __syncwarp(); if ((threadIdx.x & 31) == 0) mbar_arrive(dread_bar); } }
        const uint32_t (&v)[32] = va[c];
        const uint32_t m = c ? mask_hi : mask_lo;
#pragma unroll
        for (int k = 0; k < 16; ++k)
            o[c * 16 + k] = pack_sat_h2(__uint_as_float(v[2 * k]), __uint_as_float(v[2 * k + 1])) & mask_expand(m, k);
    }
}
__device__ __forceinline__ void drain_store(uint8_t* slot, int f, const uint32_t (&o)[32]) {
#pragma unroll
    for (int c = 0; c < 8; ++c)
        *reinterpret_cast<uint4*>(slot + ((size_t)(c * 128 + f) << 4)) = make_uint4(o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
}

// five / two shuffles of one scan level issued back to back into DISTINCT registers: written as separate __shfl_up_sync calls the
// register allocator funnelled them through one register (shuffle, consume, shuffle, ...), five dependent shuffle latencies per level
__device__ __forceinline__ void shfl_up5(float (&o)[5], float i0, float i1, float i2, float i3, float i4, int off, int width) {
    const int c = (32 - width) << 8;
    asm volatile("shfl.sync.up.b32 %0, %5, %10, %11, 0xffffffff;\n\tshfl.sync.up.b32 %1, %6, %10, %11, 0xffffffff;\n\t"
                 "shfl.sync.up.b32 %2, %7, %10, %11, 0xffffffff;\n\tshfl.sync.up.b32 %3, %8, %10, %11, 0xffffffff;\n\t"
                 "shfl.sync.up.b32 %4, %9, %10, %11, 0xffffffff;"
                 : "=f"(o[0]), "=f"(o[1]), "=f"(o[2]), "=f"(o[3]), "=f"(o[4]) : "f"(i0), "f"(i1), "f"(i2), "f"(i3), "f"(i4), "r"(off), "r"(c));
}
__device__ __forceinline__ void shfl_down2(float (&o)[2], float i0, float i1, int off, int width) {
    const int c = ((32 - width) << 8) | 0x1f;
    asm volatile("shfl.sync.down.b32 %0, %2, %4, %5, 0xffffffff;\n\tshfl.sync.down.b32 %1, %3, %4, %5, 0xffffffff;"
                 : "=f"(o[0]), "=f"(o[1]) : "f"(i0), "f"(i1), "r"(off), "r"(c));
}

// ---- MMA issue: the tile program of ONE stream as fourteen numbered operations (wait for the operands, issue, commit) ----
// The whole warp runs this code with warp-uniform values; one elected lane issues (operands stay in uniform registers).
// Descriptors are rebuilt from an opaque copy of the base every tile: hoisted out of the loop they would occupy ~100 registers,
// rebuilt they are one uniform add each.  Two ways to run it:
//   * one issuer warp per stream (streams half a tile apart, TNERF_TRAIN_SYNC=0): each warp loops over its own operations;
//   * ONE warp for both streams (streams in phase, the default): operation k of stream 0, then operation k of stream 1.  With two
//     issuer warps the tensor pipe alternates between their instructions, both batches of a step complete together at TWICE the
//     batch time, both accumulators are drained together and the next step collides again -- a self-sustaining lock step.  Issued
//     batch after batch, stream 0's accumulator is complete one batch time before stream 1's, its drain and its next batch start that
//     much earlier, and from then on every batch has the pipe to itself (the streams stay one batch apart).
constexpr int N_OPS = 14;
template <int KX, int SS>
struct StreamIssuer {
    static constexpr int XS = KX / 16;
    static constexpr uint32_t P = S_P0, Q = S_Q0, X = S_X0, DZH = S_DZH0;   // byte offsets of stream 0; the stream enters through the bases
    static_assert(S_P1 - S_P0 == 32768 && S_Q1 - S_Q0 == 32768 && S_X1 - S_X0 == 8192 && S_DZH1 - S_DZH0 == 2048, "stream strides");
    enum { PH_X = 1, PH_IN = 2, PH_DZH = 4, PH_GF = 8, PH_DR = 16, PH_XQ = 64 };
    uint32_t mb, tmem, ph = 0;      // shared-memory address of Misc, tensor-memory base, phase bits of the barriers this warp waits on
    long long* dbg;
    int dbg_n = 0;
    __device__ __forceinline__ StreamIssuer(Misc& ms, uint32_t tmem_, long long* dbg_) : mb(smem_u32(&ms)), tmem(tmem_), dbg(dbg_) {}
#define T2_BAR(field) (mb + (uint32_t)offsetof(Misc, field) + 8u * SS)
    __device__ __forceinline__ void wait(uint32_t bar, uint32_t bit) {
        T2_STAMP();
        mbar_wait(bar, (ph & bit) ? 1u : 0u);
        ph ^= bit;
        tc_fence_after();
        T2_STAMP();
    }
#define T2_ISSUE(...) do { if (elect_one()) { __VA_ARGS__ } __syncwarp(); } while (0)
    // layer 0 of a tile is issued EARLY, at the end of the previous tile (the first one from here): the encoding buffer X is single
    // and stays busy until the tile's last GEMM (dW0), so the sample warps stage a second copy of the next tile's features in the Q
    // slot as soon as its last reader (dH0) has completed; F0 reads that copy while the drain threads still store dZ0 and dW0 runs.
    // The regular X buffer is refilled after dW0 and is first needed by layer 2.
    __device__ __forceinline__ void first(uint32_t sb, long long n) {
        if (n <= 0) return;
        const uint32_t D = tmem + C_D + 64 * SS;
        wait(T2_BAR(bar_xq), PH_XQ);
        T2_ISSUE(gemm<XS>(D, kmaj(sb, S_W0, 128), kmaj(sb + SS * (32768u >> 4), Q, 64), make_idesc_f16(128, 64, 0, 0), 0); tc_commit(T2_BAR(bar_d)););
    }
    // "X buffer refilled" (needed from layer 2 on) is looked at right after layer 1 has been issued for both streams: the issuer would
    // wait for the layer-1 drain then anyway, so the ~70 cycles even a completed mbarrier costs are off the serial path
    __device__ __forceinline__ void wait_x() { wait(T2_BAR(bar_x), PH_X); }
    template <int OP>
    __device__ __forceinline__ void op(uint32_t sb, long long t, long long n) {
        constexpr uint32_t i64kk = make_idesc_f16(128, 64, 0, 0), i64kt = make_idesc_f16(128, 64, 0, 1), i64tk = make_idesc_f16(128, 64, 1, 0),
                           i64tt = make_idesc_f16(128, 64, 1, 1), i16tk = make_idesc_f16(128, 16, 1, 0), i16kt = make_idesc_f16(128, 16, 0, 1),
                           i128kk = make_idesc_f16(128, 128, 0, 0), iXkt = make_idesc_f16(128, KX, 0, 1);
        const uint32_t D = tmem + C_D + 64 * SS;
        const uint32_t sbA = sb + SS * (32768u >> 4), sbX = sb + SS * (8192u >> 4), sbZ = sb + SS * (2048u >> 4), sbH = sb + SS * (16384u >> 4);
        // operands (descriptor words); weights are rows = output features
        const Op aW0 = kmaj(sb, S_W0, 128), aW1 = kmaj(sb, S_W1, 128), aW2h = kmaj(sb, S_W2, 128),
                 aW2x = kmaj(sb, S_W2 + 32768, 128), aW3 = kmaj(sb, S_W3, 128);
        const Op aW1t = mnmaj(sb, S_W1, 128), aW2t = mnmaj(sb, S_W2, 128), aW3t = mnmaj(sb, S_W3, 128);
        const Op bWH = kmaj(sb, S_WH, 16), aWHt = mnmaj(sb, S_WH, 16);
        const Op bXk = kmaj(sbX, X, 64), bXt = mnmaj(sbX, X, 64);
        const Op bP = mnmaj(sbA, P, 128), bQ = mnmaj(sbA, Q, 128);   // [sample x feature] readers (forward / dgrad)
        const Op aP = kmaj(sbA, P, 128), aQ = kmaj(sbA, Q, 128);     // [feature x sample] readers (wgrad)
        const Op bP_lo = kmaj(sbA, P, 128), bP_hi = kmaj(sbA, P + 1024, 128);
        // head GEMM reads H3 as the M operand with samples on the rows: 128 rows are fetched, the tile's 64 samples
        // land on accumulator lanes 64 s .. 64 s + 63 (stream 1 starts one slot early), the other rows are ignored
        const Op aQhead = mnmaj(sbH, Q, 128);                        // stream 1: Q1 - 16384 = Q0 + 16384
        const Op bDZHt = mnmaj(sbZ, DZH, 64), bDZHk = kmaj(sbZ, DZH, 64);
        const Op bXq = kmaj(sbA, Q, 64);                             // staged copy of the NEXT tile's features
        const uint32_t bar_in = T2_BAR(bar_in), bar_d = T2_BAR(bar_d), bar_wg = T2_BAR(bar_wg), bar_gfree = T2_BAR(bar_gfree);
        if constexpr (OP == 0) {            // F1: H1 (F0 was issued early)
            wait(bar_in, PH_IN);
            T2_ISSUE(gemm<8>(D, aW1, bP, i64kt, 0); tc_commit(bar_d););
        } else if constexpr (OP == 1) {     // F2: H2 (needs the refilled X buffer: wait_x() below)
            wait(bar_in, PH_IN);
            T2_ISSUE(gemm<8>(D, aW2h, bQ, i64kt, 0); gemm<XS>(D, aW2x, bXk, i64kk, 1); tc_commit(bar_d););
        } else if constexpr (OP == 2) {     // F3: H3
            wait(bar_in, PH_IN);
            T2_ISSUE(gemm<8>(D, aW3, bP, i64kt, 0); tc_commit(bar_d););
        } else if constexpr (OP == 3) {     // heads (samples on lanes)
            wait(bar_in, PH_IN);
            T2_ISSUE(gemm<8>(D, aQhead, bWH, i16tk, 0); tc_commit(T2_BAR(bar_head)););
        } else if constexpr (OP == 4) {     // head dgrad -> dH3 (the tile's serial chain continues through it)
            wait(T2_BAR(bar_dzh), PH_DZH);
            T2_ISSUE(gemm<1>(D, aWHt, bDZHk, i64tk, 0); tc_commit(bar_d););
        } else if constexpr (OP == 5) {     // head wgrad: H3 . dZh -- behind the dgrad, into four columns of the accumulator the drain
            wait(T2_BAR(bar_dread), PH_DR); // threads have just read dH3 out of: it runs while they turn dH3 into dZ3 (which they store
            T2_ISSUE(gemm<4>(D + 16, aQ, bDZHt, i16kt, 0); tc_commit(bar_d););   // over H3 only after this GEMM has committed)
        } else if constexpr (OP == 6) {
            // layers 3 and 2: dgrad first on its own barrier, the wgrad GEMMs behind it -- the drain threads turn dH into the
            // packed dZ row while the wgrad still reads the slot, and store once the wgrad has committed
            wait(bar_in, PH_IN);
            T2_ISSUE(gemm<8>(D, aW3t, bQ, i64tt, 0); tc_commit(bar_d);                                  // dH2
                     gemm<4>(tmem + C_DW3, aQ, aP, i128kk, 1); tc_commit(bar_wg););                     // dW3 += dZ3 . H2^T
        } else if constexpr (OP == 7) {
            wait(bar_in, PH_IN);
            T2_ISSUE(gemm<8>(D, aW2t, bP, i64tt, 0); tc_commit(bar_d);                                  // dH1
                     gemm<4>(tmem + C_DW2, aP, aQ, i128kk, 1);                                          // dW2[:, :128] += dZ2 . H1^T
                     gemm<4>(tmem + C_DW2 + 128, aP, bXt, iXkt, 1); tc_commit(bar_wg););                // dW2[:, 128:] += dZ2 . X^T
        } else if constexpr (OP == 8) {     // recompute H0 under the dZ1 drain
            wait(T2_BAR(bar_dread), PH_DR);                                                             // dH1 has been read out
            T2_ISSUE(gemm<XS>(D, aW0, bXk, i64kk, 0); tc_commit(bar_d););
        } else if constexpr (OP == 9) {     // dW1[:, :64] partial
            // dZ1 stored (step 8) AND the recomputed H0 stored (step 9): step 8 does not signal at all -- the same threads store H0
            // afterwards and signal once.  (Two signals on ONE mbarrier here would let two phases complete unobserved, which alias
            // to "not complete": a deadlock found by tools/stress_train.py in round 1; a second barrier cost a wait per tile.)
            wait(bar_in, PH_IN);
            T2_ISSUE(gemm<4>(D, aQ, bP_lo, i64kk, 0); tc_commit(mb + (uint32_t)offsetof(Misc, bar_g) + 16u * SS););
        } else if constexpr (OP == 10) {    // dW1[:, 64:] partial
            wait(bar_gfree, PH_GF);
            T2_ISSUE(gemm<4>(D, aQ, bP_hi, i64kk, 0); tc_commit(mb + (uint32_t)offsetof(Misc, bar_g) + 16u * SS + 8u););
        } else if constexpr (OP == 11) {    // dH0; Q is free behind it
            wait(bar_gfree, PH_GF);
            T2_ISSUE(gemm<8>(D, aW1t, bQ, i64tt, 0); tc_commit(bar_d); tc_commit(T2_BAR(bar_qfree)););
        } else if constexpr (OP == 12) {    // F0 of the NEXT tile
            if (t + 1 < n) {
                wait(T2_BAR(bar_xq), PH_XQ);        // next tile's features staged in Q (2 sample warps) AND dH0 read out of the accumulator (4 drain warps)
                T2_ISSUE(gemm<XS>(D, aW0, bXq, i64kk, 0); tc_commit(bar_d););
            }
        } else {                            // dW0 += dZ0 . X^T ; tile done
            wait(bar_in, PH_IN);
            T2_ISSUE(gemm<4>(tmem + C_DW0, aP, bXt, iXkt, 1); tc_commit(T2_BAR(bar_xfree)););
        }
    }
#undef T2_ISSUE
#undef T2_BAR
};
template <int OP, int KX, int SS>
__device__ __forceinline__ void run_ops(StreamIssuer<KX, SS>& a, uint32_t sb, long long t, long long n) {
    if constexpr (OP < N_OPS) {
        a.template op<OP>(sb, t, n);
        if constexpr (OP == 0) a.wait_x();
        run_ops<OP + 1>(a, sb, t, n);
    }
}
template <int OP, int KX>
__device__ __forceinline__ void run_ops2(StreamIssuer<KX, 0>& a, StreamIssuer<KX, 1>& b, uint32_t sb, long long t, long long n0, long long n1) {
    if constexpr (OP < N_OPS) {
        a.template op<OP>(sb, t, n0);
        if (t < n1) b.template op<OP>(sb, t, n1);
        if constexpr (OP == 0) { a.wait_x(); if (t < n1) b.wait_x(); }
        // last tile: everything that adds into dW3 (operation 6) / dW2 (operation 7) has been issued by this thread -- one commit
        // tells the sample warps that the accumulator is final (early gradient flush, see the end of their role)
        if constexpr (OP == 6 || OP == 7) {
            if (t == n0 - 1) { if (elect_one()) tc_commit(a.mb + (uint32_t)offsetof(Misc, bar_fin) + 8u * (OP - 6)); __syncwarp(); }
        }
        run_ops2<OP + 1>(a, b, sb, t, n0, n1);
    }
}
__device__ __forceinline__ uint32_t opaque_base(uint32_t sbase) {
    uint32_t sb = sbase >> 4;
    asm volatile("" : "+r"(sb));
    return sb;
}
template <int KX, int SS>
__device__ __forceinline__ void issuer_loop(Misc& ms, uint32_t sbase, uint32_t tmem, long long n_tiles_mine, long long* dbg) {
    StreamIssuer<KX, SS> a(ms, tmem, dbg);
    mbar_wait(smem_u32(&ms.bar_w), 0);
    a.first(opaque_base(sbase), n_tiles_mine);
    for (long long t = 0; t < n_tiles_mine; ++t) run_ops<0>(a, opaque_base(sbase), t, n_tiles_mine);
}
// both streams from one warp; n0 >= n1 (tiles are dealt round-robin, stream 0 first)
template <int KX>
__device__ __forceinline__ void issuer_loop_merged(Misc& ms, uint32_t sbase, uint32_t tmem, long long n0, long long n1, long long* dbg) {
    StreamIssuer<KX, 0> a(ms, tmem, dbg);
    StreamIssuer<KX, 1> b(ms, tmem, nullptr);
    mbar_wait(smem_u32(&ms.bar_w), 0);
    a.first(opaque_base(sbase), n0);
    b.first(opaque_base(sbase), n1);
    for (long long t = 0; t < n0; ++t) run_ops2<0>(a, b, opaque_base(sbase), t, n0, n1);
}

template <int KX, bool UNROLL>
__global__ void __launch_bounds__(THREADS, 1) fused_train2_kernel(const __grid_constant__ TrainParams p, const __grid_constant__ Extra ex) {
    extern __shared__ __align__(1024) uint8_t smem[];
    Misc& ms = *reinterpret_cast<Misc*>(smem + S_MISC);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wg = warp >> 2;
    const uint32_t sbase = smem_u32(smem);
    if (p.debug && threadIdx.x == 0) p.debug[1024 + 4 * blockIdx.x] = gtimer();      // per-CTA lifetime (tools/hot_cold.py)

    if (warp == 12 && lane == 0) {
        mbar_init(smem_u32(&ms.bar_w), 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&ms.bar_x[s]), 2);             // thread groups signal with ONE arrival per warp (after __syncwarp):
                                                              // 128 arrivals on one barrier serialise for a few hundred cycles
            mbar_init(smem_u32(&ms.bar_in[s]), 4);
            mbar_init(smem_u32(&ms.bar_d[s]), 1);
            mbar_init(smem_u32(&ms.bar_head[s]), 1);
            mbar_init(smem_u32(&ms.bar_dzh[s]), 2);
            mbar_init(smem_u32(&ms.bar_g[s][0]), 1);
            mbar_init(smem_u32(&ms.bar_g[s][1]), 1);
            mbar_init(smem_u32(&ms.bar_gfree[s]), 4);
            mbar_init(smem_u32(&ms.bar_xfree[s]), 1);
            mbar_init(smem_u32(&ms.bar_wg[s]), 1);
            mbar_init(smem_u32(&ms.bar_dread[s]), 4);
            mbar_init(smem_u32(&ms.bar_xq[s]), 6);            // 2 sample warps (features staged) + 4 drain warps (accumulator read; once up front for the first tile)
            mbar_init(smem_u32(&ms.bar_qfree[s]), 1);
            mbar_init(smem_u32(&ms.bar_fin[s]), 1);
        }
        fence_barrier_init();
    }
    if (warp == 0) { tmem_alloc(smem_u32(&ms.tmem_slot), 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = ms.tmem_slot;
    if (warp < 8) {   // zero the resident weight-gradient accumulators (both streams accumulate into them from the first tile)
        const uint32_t tl = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        uint32_t z[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) z[i] = 0u;
        for (int c0 = (warp >> 2) * 192; c0 < (warp >> 2) * 192 + 192; c0 += 16) tmem_st16(tl + c0, z);
        tc_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // Launched with programmatic stream serialisation: everything above (barriers, tensor-memory allocation and clearing) ran under the
    // tail of the previous kernel in the stream (the optimiser); its results -- the refreshed operand image, the cleared loss slot -- are
    // needed from here on.
    asm volatile("griddepcontrol.wait;" ::: "memory");

    // tiles are dealt round-robin over (cta, stream) pairs
#ifdef T2_SOLO      // experiment: stream 0 alone carries the CTA's tiles (length of one stream's chain without pipe collisions)
    const long long nstreams = gridDim.x;
    long long n_my[2];
    n_my[0] = (blockIdx.x < p.n_tiles) ? (p.n_tiles - blockIdx.x + nstreams - 1) / nstreams : 0;
    n_my[1] = 0;
#else
    const long long nstreams = 2LL * gridDim.x;
    long long n_my[2];
    // dealing index of this CTA: the identity, or the caller's permutation (the slowest SMs take the highest indices = one tile less
    // when the tiles do not divide evenly, tnerf_set_tile_order); a constant table, not written by the previous kernel
    const long long vb = p.tile_order ? (long long)__ldg(p.tile_order + blockIdx.x) : (long long)blockIdx.x;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const long long j = 2LL * vb + s;
        n_my[s] = (j < p.n_tiles) ? (p.n_tiles - j + nstreams - 1) / nstreams : 0;
    }
#endif
    float* slab = p.slabs + (size_t)blockIdx.x * p.sm.total;
    // early gradient flush (dW3 / dW2 by the sample warps under the last tile's tail): needs the single issuer warp (its commits tell
    // when the accumulators are final) and the one-vector flush
#if defined(T2_NO_MERGE) || defined(T2_NO_EARLY)
    const bool early_flush = false;
#else
    const bool early_flush = (p.S == 128 || p.sync_streams > 0) && p.bulk_reduce && n_my[0] > 0;
#endif

    if (wg == 3) {
        TN_SETMAXNREG_DEC(32);
        if (warp == 14 && lane == 0) {
            const uint32_t bar_w = smem_u32(&ms.bar_w);
            uint32_t total = 0;
            for (int i = 0; i < 5; ++i) total += ex.w[i].bytes;
            mbar_expect_tx(bar_w, total);
            for (int i = 0; i < 5; ++i) {
                uint32_t off = 0;
                while (off < ex.w[i].bytes) {
                    const uint32_t n = min(16384u, ex.w[i].bytes - off);
                    bulk_g2s(sbase + ex.w[i].dst + off, reinterpret_cast<const uint8_t*>(p.image) + ex.w[i].src + off, n, bar_w);
                    off += n;
                }
            }
        }
        long long* idbg = (p.debug && blockIdx.x == 0 && lane == 0) ? p.debug + 256 : nullptr;
#ifdef T2_NO_MERGE
        if (false) {
#else
        if (p.S == 128 || p.sync_streams > 0) {       // streams in phase: one warp issues for both (see StreamIssuer)
#endif
            if (warp == 12) issuer_loop_merged<KX>(ms, sbase, tmem, n_my[0], n_my[1], idbg);
        } else {
            if (warp == 12) issuer_loop<KX, 0>(ms, sbase, tmem, n_my[0], idbg);
            if (warp == 13) issuer_loop<KX, 1>(ms, sbase, tmem, n_my[1], nullptr);
        }
        __syncthreads();                                          // (A)
        __syncthreads();                                          // (B)
    } else if (wg == 2) {
        // ------------------------------ sample threads: warps 8-9 stream 0, warps 10-11 stream 1 ------------------------------
        TN_SETMAXNREG_DEC(96);
        const int s = (warp - 8) >> 1, wp = (warp - 8) & 1, i = wp * 32 + lane;
        const uint32_t Dh = tmem + ((uint32_t)((warp & 3) * 32) << 16) + C_D + 64 * s;
        uint8_t* X = smem + (s ? S_X1 : S_X0);
        uint8_t* DZH = smem + (s ? S_DZH1 : S_DZH0);
        const uint32_t bar_x = smem_u32(&ms.bar_x[s]), bar_head = smem_u32(&ms.bar_head[s]), bar_dzh = smem_u32(&ms.bar_dzh[s]),
                       bar_xfree = smem_u32(&ms.bar_xfree[s]), bar_xq = smem_u32(&ms.bar_xq[s]), bar_qfree = smem_u32(&ms.bar_qfree[s]);
        uint8_t* XQ = smem + (s ? S_Q1 : S_Q0);              // early copy of the next tile's features (layer 0 is issued ahead)
        // n_samples = 128 ("in-phase" mode): the two streams of the CTA carry the two halves of ONE ray, the four sample warps
        // composite it together (chain of four 32-sample chunks); otherwise a ray lives inside one stream's tile
        const bool inphase = p.S == 128;
        const int nchain = inphase ? 4 : (p.S == 64 ? 2 : 1);          // warps a ray spans
        const int cw = inphase ? 2 * s + wp : wp;                       // this warp's position in the chain
        const int cbar = inphase ? 1 : 1 + s, cthreads = inphase ? 128 : 64;
        float* xch = inphase ? ms.xch[0] : ms.xch[s];
        *reinterpret_cast<uint4*>(DZH + ((size_t)(64 + i) << 4)) = make_uint4(0u, 0u, 0u, 0u);   // head columns 8..15 stay zero
        const bool jit_rng = p.jitter == nullptr && p.rs.jitter_seed != 0;       // stratified jitter drawn in-kernel (Philox)
        const bool jit = p.jitter != nullptr || jit_rng;
        const float gscale = p.scale_dev ? *p.scale_dev : p.scale;
        const float inv_g = p.unscale ? __frcp_rn(gscale) : 1.f;      // power-of-two scales: exact
        const float bs = p.b_sigma[0], br = p.b_rgb[0], bg = p.b_rgb[1], bb = p.b_rgb[2];
        float hb[4] = {0.f, 0.f, 0.f, 0.f}, loss_acc = 0.f;
        bool overflow = false;          // a scaled head gradient left the fp16-safe range (or is not finite): GradScaler's found_inf
#ifdef T2_SOLO
        const long long j0 = blockIdx.x;
#else
        const long long j0 = 2LL * vb + s;
#endif
        const int S = p.S, W = S < 32 ? S : 32, sl = lane & (W - 1);
        const bool camera = p.rs.rays_d == nullptr;
        uint32_t pk[KX / 2];
        // rays (src/rays.py:21-31), depth (sampling.py), Fourier features of sample i of `tile`; returns z and delta*|d| (volume.py:18-23)
        auto encode = [&](long long tile, float& z_out, float& gap_out) {
            // everything only this lambda needs (pose, sampling constants) is fetched / derived HERE, once per tile, instead of living
            // in registers across the compositing: the sample warps run on 96 registers and the scans need room to pipeline
            float cam[12];
#pragma unroll
            for (int k = 0; k < 12; ++k) cam[k] = camera ? __ldg(p.rs.c2w + k) : 0.f;
            const float lin_step = (S > 1) ? __fdiv_rn(1.f, (float)(S - 1)) : 0.f;
            const float inv_focal = camera ? __frcp_rn(p.rs.focal) : 0.f, half_w = (float)p.rs.W * 0.5f, half_h = (float)p.rs.H * 0.5f;
            const float near_ = p.near_, far_ = p.far_;
            auto bin = [&](int k) -> float {      // bit-exact torch.linspace / z formula (src/sampling.py:16-17)
                const float t = (S <= 1) ? 0.f : ((k < S / 2) ? __fmul_rn(lin_step, (float)k) : __fmaf_rn(-lin_step, (float)(S - 1 - k), 1.f));
                return __fadd_rn(__fmul_rn(near_, __fsub_rn(1.f, t)), __fmul_rn(far_, t));
            };
            auto zsample = [&](int k, float uu) -> float {   // src/sampling.py:21-25
                const float zc = bin(k);
                if (!jit) return zc;
                const float lo = (k == 0) ? zc : __fmul_rn(0.5f, __fadd_rn(bin(k - 1), zc));
                const float hi = (k == S - 1) ? zc : __fmul_rn(0.5f, __fadd_rn(zc, bin(k + 1)));
                return __fadd_rn(lo, __fmul_rn(__fsub_rn(hi, lo), uu));
            };
            const long long ray = inphase ? (tile >> 1) : tile * p.R + i / S;
            const int si = inphase ? (int)(tile & 1) * 64 + i : i % S;
            float pt[3] = {0.f, 0.f, 0.f};
            float z = 0.f, gd = 0.f;
            if (ray < p.n_rays) {
                float o[3], d[3];
                if (!camera) load_ray(p.rs, ray, o, d);
                else {
                    const long long k = p.rs.pixel_index ? p.rs.pixel_index[ray] : p.rs.first_ray + ray;
                    const unsigned kk = (unsigned)k, Wd = (unsigned)p.rs.W;
                    const unsigned prow = kk / Wd, pcol = kk - prow * Wd;
                    // src/rays.py:21-31 with the divisions turned into multiplications by once-computed reciprocals
                    // (<= 3 ulp on the direction, inside the 1e-6 bar of the stand-alone get_rays kernel)
                    const float cx = ((float)pcol - half_w) * inv_focal;
                    const float cy = -((float)prow - half_h) * inv_focal;
                    const float wx = fmaf(-1.f, cam[2], fmaf(cy, cam[1], cx * cam[0]));
                    const float wy = fmaf(-1.f, cam[6], fmaf(cy, cam[5], cx * cam[4]));
                    const float wz = fmaf(-1.f, cam[10], fmaf(cy, cam[9], cx * cam[8]));
                    const float inv_n = rsqrtf(fmaxf(fmaf(wz, wz, fmaf(wy, wy, wx * wx)), 1e-24f));
                    d[0] = wx * inv_n; d[1] = wy * inv_n; d[2] = wz * inv_n;
                    o[0] = cam[3]; o[1] = cam[7]; o[2] = cam[11];
                }
                float u0 = 0.f, u1 = 0.f;
                if (jit_rng) {
                    u0 = jitter_uniform(p.rs.jitter_seed, p.rs.jitter_step, ray, si);
                    if (si + 1 < S) u1 = jitter_uniform(p.rs.jitter_seed, p.rs.jitter_step, ray, si + 1);
                } else if (jit) { u0 = p.jitter[ray * S + si]; if (si + 1 < S) u1 = p.jitter[ray * S + si + 1]; }
                z = zsample(si, u0);
                const float dn = sqrtf(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
                gd = ((si == S - 1) ? kLastDelta : (zsample(si + 1, u1) - z)) * dn;
#pragma unroll
                for (int c = 0; c < 3; ++c) pt[c] = __fadd_rn(o[c], __fmul_rn(d[c], z));
            }
            z_out = z; gap_out = gd;
            if (p.include_input) encode_stream<KX, true>(pt, p.L, pk); else encode_stream<KX, false>(pt, p.L, pk);
        };
        auto store_x = [&](uint8_t* dst, uint32_t bar) {
#pragma unroll
            for (int c = 0; c < KX / 8; ++c)
                *reinterpret_cast<uint4*>(dst + ((size_t)(c * 64 + i) << 4)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar);
        };
        uint32_t ph_head = 0, ph_xfree = 0, ph_qfree = 0;
        float z_cur = 0.f, gap_cur = 0.f;
        long long* dbg = (p.debug && blockIdx.x == 0 && warp == 8 && lane == 0) ? p.debug + 512 : nullptr;
        int dbg_n = 0;
        // iteration t = -1 only encodes and stores the first tile: encode() and store_x() have ONE call site each (code size)
#pragma unroll 1
        for (long long t = -1; t < n_my[s]; ++t) {
            const long long tile = j0 + t * nstreams;
            const bool more = t + 1 < n_my[s];
            float z_next = 0.f, gap_next = 0.f;
            T2_STAMP();
            if (t >= 0) {
            // per-ray inputs of the loss are fetched before the heads are ready
            const long long ray = inphase ? (tile >> 1) : tile * p.R + i / S;
            const bool valid = ray < p.n_rays;
            float t0 = 0.f, t1 = 0.f, t2 = 0.f, gd = 0.f, ga = 0.f;
            if (valid) {
                if (p.target) { t0 = p.target[3 * ray]; t1 = p.target[3 * ray + 1]; t2 = p.target[3 * ray + 2]; }
                else {
                    if (p.gC) { t0 = p.gC[3 * ray]; t1 = p.gC[3 * ray + 1]; t2 = p.gC[3 * ray + 2]; }
                    if (p.gD) gd = p.gD[ray];
                    if (p.gA) ga = p.gA[ray];
                }
            }
            mbar_wait(bar_head, ph_head); ph_head ^= 1;
            T2_STAMP();
            tc_fence_after();
            float4 own;
            {
                uint32_t v[4];
                tmem_ld4(Dh, v);
                tc_wait_ld();
                own.x = fmaxf(__uint_as_float(v[0]) + bs, 0.f);
                own.y = __fdividef(1.f, 1.f + __expf(-(__uint_as_float(v[1]) + br)));
                own.z = __fdividef(1.f, 1.f + __expf(-(__uint_as_float(v[2]) + bg)));
                own.w = __fdividef(1.f, 1.f + __expf(-(__uint_as_float(v[3]) + bb)));
            }
            // ---- compositing forward + loss gradient + reverse scan, one sample per thread, all in registers
            //      (src/volume.py:18-44 and its backward).  The body is specialised on the number of
            //      samples per ray (64 and 128 at compile time: no width tests around the shuffles, unrolled chunk stitching);
            //      other counts take the generic instance.
            //      Forward, generic instance: ONE inclusive scan of the per-sample maps (T, C) -> (T q, C + T alpha c) gives the
            //      transmittance in front of every sample and, in the segment's last lane, the four ray sums (colour, opacity).
            //      64 / 128 samples: product scan of q + one integer REDUX per ray sum (see below).
            //      Reverse: suffix composition of R -> g alpha + q R (division-free). ----
            T2_STAMP();                                              // heads read and activated
            auto composite = [&](auto SC) __attribute__((always_inline)) {
                constexpr int SCT = decltype(SC)::value;
                const int Wc = SCT ? (SCT < 32 ? SCT : 32) : W;
                const int slc = SCT >= 32 ? lane : sl;
                const int nch = SCT == 128 ? 4 : (SCT == 64 ? 2 : nchain);
                const float e = valid ? __expf(-own.x * gap_cur) : 1.f;
                const float alpha = valid ? 1.f - e : 0.f;
                const float q = valid ? 1.f - alpha + kEpsT : 1.f;
                float Qi = q, A0, A1, A2, A3;
                float excl;
                if constexpr (SCT >= 32) {
                    // a warp holds one 32-sample chunk of ONE ray: product scan of q alone (five one-shuffle levels), then the four
                    // chunk sums of the weights relative to the chunk start as 2^-30 fixed point -- one REDUX each instead of five
                    // more shuffle levels of four values (terms and sums lie in [0, 1]: exact to 2^-30, below fp32 rounding)
#pragma unroll
                    for (int off = 1; off < 32; off <<= 1) {
                        const float qu = __shfl_up_sync(0xffffffffu, Qi, off);
                        if (lane >= off) Qi *= qu;
                    }
                    excl = __shfl_up_sync(0xffffffffu, Qi, 1);
                    if (lane == 0) excl = 1.f;
                    const float wl = alpha * excl * 1073741824.f;
                    const int i0 = __float2int_rn(wl * own.y), i1 = __float2int_rn(wl * own.z), i2 = __float2int_rn(wl * own.w), i3 = __float2int_rn(wl);
                    A0 = (float)__reduce_add_sync(0xffffffffu, i0) * 9.313225746154785e-10f;
                    A1 = (float)__reduce_add_sync(0xffffffffu, i1) * 9.313225746154785e-10f;
                    A2 = (float)__reduce_add_sync(0xffffffffu, i2) * 9.313225746154785e-10f;
                    A3 = (float)__reduce_add_sync(0xffffffffu, i3) * 9.313225746154785e-10f;
                    if (__any_sync(0xffffffffu, !(wl * own.y * own.z * own.w == wl * own.y * own.z * own.w)))     // non-finite heads stay visible in the loss
                        A0 = A1 = A2 = A3 = __int_as_float(0x7fc00000);
                } else {
                    A0 = alpha * own.y; A1 = alpha * own.z; A2 = alpha * own.w; A3 = alpha;
#pragma unroll
                    for (int off = 1; off < 32; off <<= 1) {
                        if (off < Wc) {
                            float up[5];
                            shfl_up5(up, Qi, A0, A1, A2, A3, off, Wc);
                            const float qu = up[0], a0 = up[1], a1 = up[2], a2 = up[3], a3 = up[4];
                            if (slc >= off) {      // earlier segment (qu, a) followed by this one: (qu Q, a + qu A)
                                A0 = fmaf(qu, A0, a0); A1 = fmaf(qu, A1, a1); A2 = fmaf(qu, A2, a2); A3 = fmaf(qu, A3, a3);
                                Qi *= qu;
                            }
                        }
                    }
                    excl = __shfl_up_sync(0xffffffffu, Qi, 1, Wc);
                    if (slc == 0) excl = 1.f;
                }
                T2_STAMP();                                          // forward scan done
                float c0, c1, c2, asum, Tc = 1.f;
                if (nch > 1) {       // stitch the chunks of the ray: chunk c enters with T = product of the earlier chunks' transmittances
                    if (lane == 31) { float* x = xch + cw * 8; x[0] = Qi; x[1] = A0; x[2] = A1; x[3] = A2; x[4] = A3; }
                    bar_sync(cbar, cthreads);
                    float T = 1.f;
                    c0 = c1 = c2 = asum = 0.f;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        if (c < nch) {
                            const float* x = xch + c * 8;
                            if (c == cw) Tc = T;
                            c0 = fmaf(T, x[1], c0); c1 = fmaf(T, x[2], c1); c2 = fmaf(T, x[3], c2); asum = fmaf(T, x[4], asum);
                            T *= x[0];
                        }
                    }
                } else {             // the segment's last lane holds the ray sums
                    c0 = __shfl_sync(0xffffffffu, A0, Wc - 1, Wc); c1 = __shfl_sync(0xffffffffu, A1, Wc - 1, Wc);
                    c2 = __shfl_sync(0xffffffffu, A2, Wc - 1, Wc); asum = __shfl_sync(0xffffffffu, A3, Wc - 1, Wc);
                }
                T2_STAMP();                                          // ray sums known
                const float Ti = Tc * excl, w = alpha * Ti;
                const float bgc = p.white ? 1.f - asum : 0.f;
                const float C0 = c0 + bgc, C1 = c1 + bgc, C2 = c2 + bgc;
                float g0 = 0.f, g1 = 0.f, g2 = 0.f;
                const bool leader = slc == 0 && cw == 0;
                if (valid) {
                    if (p.target) {
                        const float e0 = C0 - t0, e1 = C1 - t1, e2 = C2 - t2;
                        g0 = 2.f * e0 * p.inv_denom; g1 = 2.f * e1 * p.inv_denom; g2 = 2.f * e2 * p.inv_denom;
                        if (leader) loss_acc += (e0 * e0 + e1 * e1 + e2 * e2) * p.inv_denom;
                    } else { g0 = t0; g1 = t1; g2 = t2; }
                    if (p.comp && leader) { p.comp[3 * ray] = C0; p.comp[3 * ray + 1] = C1; p.comp[3 * ray + 2] = C2; }
                }
                const float gconst = ga - (p.white ? (g0 + g1 + g2) : 0.f);
                const float g = valid ? (g0 * own.y + g1 * own.z + g2 * own.w + gd * z_cur + gconst) : 0.f;
                float Aa = g * alpha, Qq = q;        // suffix composition of the maps R -> g a + q R
                float Rc = 0.f;
                T2_STAMP();                                          // loss gradient ready
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    if (SCT >= 32 || off < Wc) {
                        float dn[2];
                        shfl_down2(dn, Aa, Qq, off, Wc);
                        const float An = dn[0], Qn = dn[1];
                        if (slc + off < Wc) { Aa = fmaf(Qq, An, Aa); Qq *= Qn; }
                    }
                }
                if (nch > 1) {       // R entering this chunk from behind = the later chunks' affine maps applied to 0, last chunk first
                    if (lane == 0) { xch[32 + 2 * cw] = Aa; xch[33 + 2 * cw] = Qq; }
                    bar_sync(cbar, cthreads);
#pragma unroll
                    for (int c = 3; c > 0; --c)
                        if (c < nch && c > cw) Rc = fmaf(xch[33 + 2 * c], Rc, xch[32 + 2 * c]);
                }
                T2_STAMP();                                          // reverse scan + stitch done
                const float Rprev = fmaf(Qq, Rc, Aa);
                float Ri = __shfl_down_sync(0xffffffffu, Rprev, 1, Wc);
                if (slc == Wc - 1) Ri = Rc;
                const float dsig = Ti * (g - Ri) * gap_cur * e;
                const float s0 = (own.x > 0.f) ? dsig * gscale : 0.f;
                const float s1 = w * g0 * own.y * (1.f - own.y) * gscale;
                const float s2 = w * g1 * own.z * (1.f - own.z) * gscale;
                const float s3 = w * g2 * own.w * (1.f - own.w) * gscale;
                *reinterpret_cast<uint4*>(DZH + ((size_t)i << 4)) = make_uint4(pack_sat_h2(s0, s1), pack_sat_h2(s2, s3), 0u, 0u);
                fence_proxy_async();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_dzh);
                T2_STAMP();
                hb[0] += s0; hb[1] += s1; hb[2] += s2; hb[3] += s3;      // head bias gradients: per-thread partials, reduced at the end
                // the hidden layers amplify |dZ| by a few units per layer at most: 2^10 at the heads keeps every fp16 operand of the
                // backward chain far from 65504 (cvt.satfinite would clip silently); a NaN fails its comparison too
                overflow |= !(fabsf(s0) <= 1024.f) || !(fabsf(s1) <= 1024.f) || !(fabsf(s2) <= 1024.f) || !(fabsf(s3) <= 1024.f);   // (fmaxf would drop a NaN)
            };
            if (S == 64) composite(std::integral_constant<int, 64>{});
            else if (S == 128) composite(std::integral_constant<int, 128>{});
            else composite(std::integral_constant<int, 0>{});
            }
            // Fourier features of the NEXT tile, computed under the backward GEMMs of this one.  (Computed before the compositing they
            // stayed live through it -- 32 registers -- and the compiler, short of registers, funnelled every shuffle of the scans
            // through one register: five dependent shuffle latencies per level instead of one.)
            T2_STAMP();
            if (more) encode(tile + nstreams, z_next, gap_next);
            T2_STAMP();
            if (!more) break;                                    // last tile: nothing to store (and the early gradient flush is waiting)
            if (t >= 0) {
                mbar_wait(bar_qfree, ph_qfree); ph_qfree ^= 1;   // dH0 has completed: the Q slot is free
                T2_STAMP();
            }
            store_x(XQ, bar_xq);                                 // early copy for layer 0 of the next tile
            if (t >= 0) {
                mbar_wait(bar_xfree, ph_xfree); ph_xfree ^= 1;   // last GEMM of the tile has completed: X may be replaced
                T2_STAMP();
            }
            store_x(X, bar_x); z_cur = z_next; gap_cur = gap_next;
        }
        loss_acc = warp_sum(loss_acc);
#pragma unroll
        for (int k = 0; k < 4; ++k) hb[k] = warp_sum(hb[k]);
        if (early_flush) {
            // ---- early gradient flush: dW3 and dW2 (160 of the CTA's 265 KB) are final well before the last tile ends (after its
            //      operations 6 / 7) and these four warps -- one per lane quadrant -- have nothing left to do.  The bulk-reduction
            //      engine moves ~23 B/clk per SM, so the flush after the last GEMM was 6.4 us of every launch; started here it runs
            //      under the rest of the last tile.  Staging: the W3 weight matrix (32 KB = two 32-column chunks) -- its last reader
            //      is operation 6 of the last tile, the very event that makes dW3 final.  (Two 4 KB buffers in the spare 10 KB were
            //      tried first: 8 KB in flight against the engine's ~1 k-cycle latency is 8 B/clk -- slower than no early flush.) ----
            const uint32_t tl = tmem + ((uint32_t)((warp & 3) * 32) << 16);
            const int f = (warp & 3) * 32 + lane;
            float* stage = reinterpret_cast<float*>(smem + S_W3);
            const bool issuer_thread = warp == 8 && lane == 0;
            int nchunk = 0;
            auto push = [&](int tcol, int ncols, int off) {       // ncols is a multiple of 16
                for (int c0 = 0; c0 < ncols; c0 += 32) {
                    const int nc = ncols - c0 < 32 ? ncols - c0 : 32;
                    if (nchunk >= 2) { if (issuer_thread) bulk_wait_group_read<1>(); bar_sync(3, 128); }
                    float* st = stage + (nchunk & 1) * 4096;
                    for (int c1 = 0; c1 < nc; c1 += 16) {
                        uint32_t v[16];
                        tmem_ld16(tl + tcol + c0 + c1, v);
                        tc_wait_ld();
#pragma unroll
                        for (int k = 0; k < 16; ++k) st[(c1 + k) * 128 + f] = __uint_as_float(v[k]) * inv_g;
                    }
                    fence_proxy_async();
                    bar_sync(3, 128);
                    if (issuer_thread) { bulk_reduce_add_f32(p.slabs + off + c0 * 128, smem_u32(st), (uint32_t)nc * 512u); bulk_commit_group(); }
                    ++nchunk;
                }
            };
            mbar_wait(smem_u32(&ms.bar_fin[0]), 0);
            tc_fence_after();
            push(C_DW3, 128, p.sm.dw3);
            mbar_wait(smem_u32(&ms.bar_fin[1]), 0);
            tc_fence_after();
            push(C_DW2, 128 + KX, p.sm.dw2);
            if (issuer_thread) bulk_wait_group<0>();
        }
        tc_fence_before();
        __syncthreads();                                          // (A) every GEMM of both streams has completed
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (p.bulk_reduce) atomicAdd(p.slabs + p.sm.hb + (warp - 8) * 4 + k, hb[k] * inv_g);
                else slab[p.sm.hb + (warp - 8) * 4 + k] = hb[k];
            }
            if (p.loss_sum && loss_acc != 0.f) atomicAdd(p.loss_sum, loss_acc);
        }
        if (p.found && __any_sync(0xffffffffu, overflow) && lane == 0) *p.found = 1.f;
        __syncthreads();                                          // (B)
    } else {
        // ------------------------------ drain threads of stream s (thread <-> feature row) ------------------------------
        TN_SETMAXNREG_INC(192);
        const int s = wg, f = (warp & 3) * 32 + lane;
        const uint32_t tl = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        const uint32_t D_own = tl + C_D + 64 * s, D_oth = tl + C_D + 64 * (1 - s);
        uint8_t* P = smem + (s ? S_P1 : S_P0);
        uint8_t* Q = smem + (s ? S_Q1 : S_Q0);
        const uint32_t bar_in = smem_u32(&ms.bar_in[s]), bar_d = smem_u32(&ms.bar_d[s]), bar_wg = smem_u32(&ms.bar_wg[s]), bar_dread = smem_u32(&ms.bar_dread[s]), bar_xfree = smem_u32(&ms.bar_xfree[s]), bar_xq = smem_u32(&ms.bar_xq[s]);
        const uint32_t bar_g_own = smem_u32(&ms.bar_g[s][s]), bar_g_oth = smem_u32(&ms.bar_g[1 - s][s]);
        const uint32_t bar_gfree_own = smem_u32(&ms.bar_gfree[s]), bar_gfree_oth = smem_u32(&ms.bar_gfree[1 - s]);
        float dw1[64];                  // dW1[f][64 s + j]: this warpgroup's half of the columns, BOTH streams
#pragma unroll
        for (int j = 0; j < 64; ++j) dw1[j] = 0.f;
        uint32_t stash[32];             // H1 row of this feature, parked between layer 2 forward and layer 2 wgrad
#pragma unroll
        for (int j = 0; j < 32; ++j) stash[j] = 0u;
        uint32_t mk0a = 0u, mk0b = 0u, mk1a = 0u, mk1b = 0u, mk2a = 0u, mk2b = 0u, mk3a = 0u, mk3b = 0u;   // ReLU masks of H0 (recomputed), H1, H2, H3
        float dwh[4] = {0.f, 0.f, 0.f, 0.f}, db1 = 0.f, db3 = 0.f;
        const float b1 = p.b1[f], b3 = p.b3[f];
        const float inv_g = p.unscale ? __frcp_rn(p.scale_dev ? *p.scale_dev : p.scale) : 1.f;
        uint32_t ph_d = 0, ph_g_own = 0, ph_g_oth = 0, ph_wg = 0, ph_xf = 0;
        long long g_oth_left = n_my[1 - s];
        const bool inphase = p.S == 128 || p.sync_streams > 0;
        long long* dbg = (p.debug && blockIdx.x == 0 && (warp & 3) == 0 && lane == 0) ? p.debug + (s ? 768 : 0) : nullptr;
        int dbg_n = 0;
        if (dbg) dbg[251] = gtimer();

        // ---- the tile program of a drain warpgroup, as a rolled step machine: every drain body exists ONCE in the kernel image
        //      (unrolled per step the kernel's hot code was ~75 KB per tile against a 32 KB instruction cache, and most SMs ran
        //      20 % slower than the few that happened to hit).  Steps of one tile:
        //        0-3  forward drains H0->P, H1->Q (+ parked in registers), H2->P, H3->Q
        //        4    [staggered streams] the other stream's dW1 half: fixed rendezvous where this warpgroup would otherwise sleep
        //             through its own compositing; keeps the streams half a tile apart, blocking waits only
        //        5    (nothing)          6  dZ3 from the head input gradient; the head weight gradient (4 columns), issued behind it, is
        //             collected before dZ3 replaces H3 (Q); bias gradient of layer 3
        //        7    dZ2 over H2 (P) once dW3 has committed, H1 back into Q
        //        8    dZ1 over H1 (Q) once dW2 has committed, bias gradient of layer 1; releases the accumulator early (H0 recompute)
        //        9    recomputed H0 -> P                         10  dW1 halves (own; in-phase mode also the other stream's, see below)
        //        11   dZ0 over H0 (P)
        //      after the last tile: the other stream's remaining dW1 halves.
        // In-phase mode (n_samples = 128, both streams in the same tile phase): the dW1 halves are drained where they are produced,
        // in the order stream 0 first half (WG0), stream 0 second half (WG1) | stream 1 first half (WG0), stream 1 second half (WG1).
        enum { K_FWD = 0, K_G = 1, K_BWD = 3 };
#define T2_SIGNAL(bar) do { fence_proxy_async(); tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive(bar); } while (0)
        // one step of the tile program; returns 0 = next step, 1 = stay in this step (tail: more foreign halves), 2 = leave the tile.
        // ROLLED (default): called from a rolled loop, each drain body exists once (instruction-cache footprint: the two streams of a
        // CTA are half a tile apart and run different steps at the same time).  UNROLL: the twelve calls of a full tile are unrolled and
        // specialised per step: 15 % faster when the streams run in phase (n_samples = 128, BASELINE config 5); the host picks.
        auto do_step = [&](const int step, const bool tail, const long long t) __attribute__((always_inline)) -> int {
                if (step == 5) return 0;                                // (the head weight gradient is collected inside step 6)
                const int kind = (step == 4 || step == 10) ? K_G : (step <= 3 || step == 9) ? K_FWD : K_BWD;
                uint8_t* slot = ((0x14A >> step) & 1) ? Q : P;          // Q for steps 1, 3, 6, 8
                if (tail && step == 11) return 2;
                if (kind == K_G) {
                    // job list: bit j of `own_mask` = job j drains this stream's accumulator, else the other stream's
                    int njobs = 0, own_mask = 0;
                    if (tail) njobs = (int)(g_oth_left > 2 ? 2 : g_oth_left);                       // (re-entered below until none are left)
                    else if (step == 4) njobs = (!inphase && (s == 1 || t >= 1) && g_oth_left > 0) ? 1 : 0;
                    else if (!inphase) { njobs = 1; own_mask = 1; }
                    else { njobs = g_oth_left > 0 ? 2 : 1; own_mask = (njobs == 1) ? 1 : (s == 0 ? 1 : 2); }
#pragma unroll 1
                    for (int j = 0; j < njobs; ++j) {
                        const bool own = (own_mask >> j) & 1;
                        const uint32_t bar = own ? bar_g_own : bar_g_oth;
                        uint32_t& ph = own ? ph_g_own : ph_g_oth;
                        T2_STAMP();
                        mbar_wait(bar, ph);
                        ph ^= 1;
                        tc_fence_after();
                        T2_STAMP();
                        const uint32_t D = own ? D_own : D_oth;
                        uint32_t v[2][32];
                        tmem_ld32(D, v[0]);
                        tmem_ld32(D + 32, v[1]);
                        tc_wait_ld();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(own ? bar_gfree_own : bar_gfree_oth);     // the accumulator is released as soon as it has been read
#pragma unroll
                        for (int c = 0; c < 2; ++c)
#pragma unroll
                            for (int i = 0; i < 32; ++i) dw1[c * 32 + i] += __uint_as_float(v[c][i]);
                        if (!own) --g_oth_left;
                    }
                    return (tail && g_oth_left > 0) ? 1 : 0;                                         // stay in the tail step
                }
                T2_STAMP();
                mbar_wait(bar_d, ph_d);
                ph_d ^= 1;
                tc_fence_after();
                T2_STAMP();
                if (kind == K_FWD) {
                    const bool biased = step == 1 || step == 3;       // layers 1 / 3: fp32 bias added here (uniform branch: no cost elsewhere)
                    const float bias = (step == 1) ? b1 : b3;
                    const bool park = step == 1;
                    uint32_t va[2][32];
                    uint32_t o[32];
                    tmem_ld32(D_own, va[0]);
                    tc_wait_ld();
                    tmem_ld32(D_own + 32, va[1]);
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        if (c == 1) tc_wait_ld();
                        else if (step == 0 && t > 0) { mbar_wait(bar_xfree, ph_xf); ph_xf ^= 1; }   // layer 0 ran ahead: dW0 of the previous tile still read P
                        uint32_t (&v)[32] = va[c];
                        if (biased) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + bias);
                        }
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int w = (c * 4 + j) * 4;
                            o[w + 0] = pack_relu_h2(__uint_as_float(v[8 * j + 0]), __uint_as_float(v[8 * j + 1]));
                            o[w + 1] = pack_relu_h2(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3]));
                            o[w + 2] = pack_relu_h2(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5]));
                            o[w + 3] = pack_relu_h2(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7]));
                            *reinterpret_cast<uint4*>(slot + ((size_t)((c * 4 + j) * 128 + f) << 4)) = make_uint4(o[w], o[w + 1], o[w + 2], o[w + 3]);
                        }
                    }
                    T2_SIGNAL(bar_in);
                    // bookkeeping AFTER the hand-off (off the tile's serial chain, under the next GEMM): ReLU masks, parked copy
                    if (step != 0) {          // the first H0 is recomputed before its mask is needed
                        uint32_t mk[2] = {0u, 0u};
#pragma unroll
                        for (int c = 0; c < 2; ++c)
#pragma unroll
                            for (int k = 0; k < 16; ++k) mk[c] = mask_collect(mk[c], o[c * 16 + k], k);
                        asm volatile("" : "+r"(mk[0]), "+r"(mk[1]));     // pin: sunk to its use half a tile later, the packed row stays live (spills)
                        if (step == 1) { mk1a = mk[0]; mk1b = mk[1]; } else if (step == 2) { mk2a = mk[0]; mk2b = mk[1]; }
                        else if (step == 3) { mk3a = mk[0]; mk3b = mk[1]; } else if (step == 9) { mk0a = mk[0]; mk0b = mk[1]; }
                    }
                    if (park) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) stash[i] = o[i];
                    }
                } else {
                    // dZ = dH * (H > 0): computed into registers while the weight-gradient GEMM that still reads the slot runs,
                    // stored once that GEMM has committed (steps 7, 8); steps 6 and 11 store at once
                    uint32_t o[32];
                    const uint32_t ma = step == 6 ? mk3a : step == 7 ? mk2a : step == 8 ? mk1a : mk0a;
                    const uint32_t mb = step == 6 ? mk3b : step == 7 ? mk2b : step == 8 ? mk1b : mk0b;
                    drain_bwd_compute(D_own, ma, mb, o, (step == 6 || step == 8) ? bar_dread : step == 11 ? bar_xq : 0u);
                    if (step == 7 || step == 8) { T2_STAMP(); mbar_wait(bar_wg, ph_wg); ph_wg ^= 1; T2_STAMP(); }
                    if (step == 6) {
                        // the head weight-gradient GEMM was issued behind "dH3 read" and still reads H3 from the slot: wait for it,
                        // collect its four columns, then dZ3 may replace H3
                        T2_STAMP(); mbar_wait(bar_d, ph_d); ph_d ^= 1; tc_fence_after(); T2_STAMP();
                        uint32_t v[4];
                        tmem_ld4(D_own + 16, v);
                        tc_wait_ld();
#pragma unroll
                        for (int k = 0; k < 4; ++k) dwh[k] += __uint_as_float(v[k]);
                    }
                    drain_store(slot, f, o);
                    if (step == 7) drain_store(Q, f, stash);                                         // H1 back into Q (dZ3 is dead)
                    if (step != 8) T2_SIGNAL(bar_in);          // dZ1 (step 8) is covered by the signal of step 9: same threads, stored before
                    if (step == 6 || step == 8) {                     // bias gradient of layer 3 / 1 = row sum of dZ3 / dZ1 (after the hand-off)
                        float sum = 0.f;
#pragma unroll
                        for (int i = 0; i < 32; i += 4) sum += (h2sum(o[i]) + h2sum(o[i + 1])) + (h2sum(o[i + 2]) + h2sum(o[i + 3]));
                        asm volatile("" : "+f"(sum));
                        if (step == 6) db3 += sum; else db1 += sum;
                    }
                }
                return 0;
        };
        if (lane == 0) mbar_arrive(bar_xq);       // the first tile's layer 0 has no accumulator to wait for
#pragma unroll 1
        for (long long t = 0; t <= n_my[s]; ++t) {
            const bool tail = t == n_my[s];
            if (UNROLL && !tail) {
#pragma unroll
                for (int step = 0; step < 12; ++step) do_step(step, false, t);
            } else {
#pragma unroll 1
                for (int step = tail ? 10 : 0; step < 12; ++step) {
                    const int r = do_step(step, tail, t);
                    if (r == 2) break;
                    if (r == 1) --step;
                }
            }
        }
#undef T2_SIGNAL
        if (dbg) dbg[253] = clock64();
        if (p.debug && threadIdx.x == 0) p.debug[1025 + 4 * blockIdx.x] = gtimer();
        tc_fence_before();
        __syncthreads();                                          // (A)
        tc_fence_after();
        if (dbg) dbg[254] = clock64();
        if (p.debug && threadIdx.x == 0) p.debug[1026 + 4 * blockIdx.x] = gtimer();
        if (p.bulk_reduce) {
            // ---- add this CTA's weight gradients into the ONE global vector: 32-column chunks (16 KB) are staged in the now idle
            //      activation slots and handed to the bulk-copy engine as reductions (cp.reduce.async.bulk .add.f32: the adds run
            //      in L2, no 39 MB of slabs, no reduce kernel).  Two staging halves per warpgroup; the issuing thread waits for
            //      the engine to have READ a half before it is refilled. ----
            float* G = p.slabs;
            // every GEMM has completed: the whole shared memory is free.  NBUF staging chunks of 16 KB per warpgroup (weights and
            // slots region): a bulk reduction has a latency of a few thousand cycles, its throughput grows with the bytes in flight
            constexpr int NBUF = T2_FLUSH_NBUF;
            static_assert(2 * NBUF * 16384 <= S_MISC, "staging exceeds the shared memory below Misc");
            float* stage = reinterpret_cast<float*>(smem + s * NBUF * 16384);
            const bool issuer_thread = (warp & 3) == 0 && lane == 0;
            int nchunk = 0;
            auto begin_chunk = [&]() -> float* {
                if (nchunk >= NBUF) { if (issuer_thread) bulk_wait_group_read<NBUF - 1>(); bar_sync(4 + s, 128); }
                return stage + (nchunk % NBUF) * 4096;
            };
            auto end_chunk = [&](int off, int ncols) {
                fence_proxy_async();
                bar_sync(4 + s, 128);
                if (issuer_thread) { bulk_reduce_add_f32(G + off, smem_u32(stage + (nchunk % NBUF) * 4096), (uint32_t)ncols * 512u); bulk_commit_group(); }
                ++nchunk;
            };
            auto push_tmem = [&](int tcol, int ncols, int off) {           // ncols is a multiple of 16
                for (int c0 = 0; c0 < ncols; c0 += 32) {
                    const int nc = ncols - c0 < 32 ? ncols - c0 : 32;
                    float* st = begin_chunk();
                    for (int c1 = 0; c1 < nc; c1 += 16) {
                        uint32_t v[16];
                        tmem_ld16(tl + tcol + c0 + c1, v);
                        tc_wait_ld();
#pragma unroll
                        for (int k = 0; k < 16; ++k) st[(c1 + k) * 128 + f] = __uint_as_float(v[k]) * inv_g;
                    }
                    end_chunk(off + c0 * 128, nc);
                }
            };
            if (early_flush) {               // dW3 / dW2 are already on their way (sample warps); dW0 is split between the warpgroups
                constexpr int N0 = KX >= 32 ? ((KX / 2 + 15) & ~15) : KX;
                if (s == 0) push_tmem(C_DW0, N0, p.sm.dw0);
                else if (KX > N0) push_tmem(C_DW0 + N0, KX - N0, p.sm.dw0 + N0 * 128);
            } else if (s == 0) { push_tmem(C_DW3, 128, p.sm.dw3); push_tmem(C_DW0, KX, p.sm.dw0); }
            else push_tmem(C_DW2, 128 + KX, p.sm.dw2);
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
                float* st = begin_chunk();
#pragma unroll
                for (int j = 0; j < 32; ++j) st[j * 128 + f] = dw1[h2 * 32 + j] * inv_g;
                end_chunk(p.sm.dw1 + (64 * s + 32 * h2) * 128, 32);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) atomicAdd(G + p.sm.dwh + k * 128 + f, dwh[k] * inv_g);
            atomicAdd(G + p.sm.db1 + f, db1 * inv_g);
            atomicAdd(G + p.sm.db3 + f, db3 * inv_g);
            if (issuer_thread) bulk_wait_group<0>();
            __syncthreads();                                      // (B)
            if (dbg) { dbg[255] = clock64(); dbg[252] = gtimer(); }
            tc_fence_before();
        } else {
        // ---- flush this CTA's weight-gradient slab (coalesced: consecutive rows) ----
        auto flush_tmem = [&](int tcol, int ncols, int off) {
            for (int c0 = 0; c0 < ncols; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(tl + tcol + c0, v);
                tc_wait_ld();
#pragma unroll
                for (int k = 0; k < 16; ++k) slab[off + (c0 + k) * 128 + f] = __uint_as_float(v[k]);
            }
        };
        if (s == 0) { flush_tmem(C_DW3, 128, p.sm.dw3); flush_tmem(C_DW0, KX, p.sm.dw0); }
        else flush_tmem(C_DW2, 128 + KX, p.sm.dw2);
#pragma unroll
        for (int j = 0; j < 64; ++j) slab[p.sm.dw1 + (64 * s + j) * 128 + f] = dw1[j];
        float* xch = reinterpret_cast<float*>(smem + S_P0);     // activation slots are free now
        if (s == 1) {
#pragma unroll
            for (int k = 0; k < 4; ++k) xch[k * 128 + f] = dwh[k];
            xch[4 * 128 + f] = db1; xch[5 * 128 + f] = db3;
        }
        __syncthreads();                                          // (B) stream 1's partial sums are visible
        if (s == 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) slab[p.sm.dwh + k * 128 + f] = dwh[k] + xch[k * 128 + f];
            slab[p.sm.db1 + f] = db1 + xch[4 * 128 + f];
            slab[p.sm.db3 + f] = db3 + xch[5 * 128 + f];
        }
        if (dbg) { dbg[255] = clock64(); dbg[252] = gtimer(); }
        tc_fence_before();
        }
    }
    __syncthreads();                                              // (C)
    if (warp == 0) tmem_dealloc(tmem, 512);
    if (p.debug && threadIdx.x == 0) p.debug[1027 + 4 * blockIdx.x] = gtimer();
}

}  // namespace t2

int fused_train2(tnerf_handle* h, const FusedPlan& fp, const TrainParams& p, int Kx, int grid, cudaStream_t s) {
    t2::Extra ex{};
    // sub-ranges of the packed forward image (tnerf_fused.cu): the first `fan-in` reduction columns of every layer
    // ([k/8][n][k%8] layout puts them first); the bias steps of layers 1/3 and of the heads are not used here
    const uint32_t kx = (uint32_t)Kx;
    ex.w[0] = {fp.layer[0].b_off, t2::S_W0, kx * 256u};
    ex.w[1] = {fp.layer[1].b_off, t2::S_W1, 32768u};
    ex.w[2] = {fp.layer[2].b_off, t2::S_W2, (128u + kx) * 256u};
    ex.w[3] = {fp.layer[3].b_off, t2::S_W3, 32768u};
    ex.w[4] = {fp.layer[4].b_off, t2::S_WH, 4096u};
    const size_t smem = t2::S_MISC + sizeof(t2::Misc);
    // Schedule of the two streams of a CTA and tile program (measured: tools/variant_crossover.py, tools/sweep_c5.py):
    //   * IN PHASE + UNROLLED (default): both streams run the same step of their tiles at the same time (each drain warpgroup drains
    //     its dW1 half for both streams where the halves are produced), so both warpgroups execute the same per-step specialised code
    //     and the larger image costs nothing: 5-8 % faster than the alternative from 100 to 65 536 rays at 16-64 samples, 15 % at 128
    //     (where in-phase is the only mode: the two streams carry the halves of one ray);
    //   * half a tile apart + ROLLED: the previous default (the foreign dW1 half is drained in the middle of the own tile, each drain
    //     body exists once because the streams execute different steps at the same time); kept as the comparison arm
    //     (TNERF_TRAIN_SYNC=0).
    // Option unroll_from = tiles per stream from which the unrolled program runs (tuning / tests).
    TrainParams q = p;
    const long long per_stream = q.n_tiles / (2ll * grid);
    if (q.sync_streams < 0) q.sync_streams = 1;
    const bool unroll = h->opt_unroll_from >= 0 ? per_stream >= h->opt_unroll_from : (q.sync_streams || q.S == 128);
    auto kern = unroll ? (Kx == 64 ? t2::fused_train2_kernel<64, true> : Kx == 48 ? t2::fused_train2_kernel<48, true>
                          : Kx == 32 ? t2::fused_train2_kernel<32, true> : t2::fused_train2_kernel<16, true>)
                       : (Kx == 64 ? t2::fused_train2_kernel<64, false> : Kx == 48 ? t2::fused_train2_kernel<48, false>
                          : Kx == 32 ? t2::fused_train2_kernel<32, false> : t2::fused_train2_kernel<16, false>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("fused train (two-stream): shared memory request rejected"); return (int)e; }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(t2::THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = (pdl_mask() & 1) ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kern, q, ex);
    return count_launch();
}

}  // namespace tnerf
