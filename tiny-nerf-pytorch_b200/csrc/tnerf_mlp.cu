// fp32 TinyNeRF MLP forward / backward as a sequence of tiled FFMA GEMMs (src/nerf.py:29-41 and the
// autograd graph it induces).  This is the stand-alone drop-in for `model(xenc)` and the exact-mode
// building block; the tensor-core path is in tnerf_fused.cu.
#include "tnerf_internal.cuh"

namespace tnerf {

static inline bool skip_into(const tnerf_handle* h, int layer) {
    // layer `layer` consumes [h_{layer-1}, x] when the concat happened after layer skip_at-1
    return layer >= 1 && layer == h->skip_at && h->skip_at <= h->depth - 1;
}

#define TN_TRY(expr) do { int _e = (expr); if (_e) return _e; } while (0)

int mlp_forward_f32(tnerf_handle* h, const float* x, long long n, float* rgb, float* sigma, float* acts, float* tmp,
                    cudaStream_t s) {
    const int H = h->hidden, D = h->in_dim;
    const float* prev = x;
    for (int i = 0; i < h->depth; ++i) {
        float* out = acts ? acts + (long long)i * n * H : tmp + (long long)(i & 1) * n * H;
        const float* W = h->params[2 * i];
        const float* b = h->params[2 * i + 1];
        const int fan = h->layer_in[i];
        GemmArgs g;
        g.A = prev; g.lda = (i == 0) ? D : H; g.B = W; g.ldb = fan; g.C = out; g.ldc = H;
        g.M = n; g.N = H; g.K = (i == 0) ? D : H;
        if (skip_into(h, i)) {
            TN_TRY(launch_gemm(g, true, true, s));                 // h part, no bias / activation yet
            GemmArgs g2 = g;
            g2.A = x; g2.lda = D; g2.B = W + H; g2.K = D; g2.bias = b; g2.flags = GEMM_ACCUM | GEMM_RELU;
            TN_TRY(launch_gemm(g2, true, true, s));
        } else {
            g.bias = b; g.flags = GEMM_RELU;
            TN_TRY(launch_gemm(g, true, true, s));
        }
        prev = out;
    }
    const int P = 2 * h->depth;
    GemmArgs gs;
    gs.A = prev; gs.lda = (h->depth == 0) ? D : H; gs.K = gs.lda; gs.M = n;
    gs.B = h->params[P]; gs.ldb = gs.K; gs.bias = h->params[P + 1]; gs.C = sigma; gs.ldc = 1; gs.N = 1; gs.flags = GEMM_RELU;
    TN_TRY(launch_gemm(gs, true, true, s));
    GemmArgs gc = gs;
    gc.B = h->params[P + 2]; gc.bias = h->params[P + 3]; gc.C = rgb; gc.ldc = 3; gc.N = 3; gc.flags = GEMM_SIGMOID;
    TN_TRY(launch_gemm(gc, true, true, s));
    return 0;
}

long long mlp_bwd_scratch_floats(const tnerf_handle* h, long long n) { return n * (2LL * h->hidden + 4); }

int mlp_backward_f32(tnerf_handle* h, const float* x, long long n, const float* acts, const float* rgb, const float* sigma,
                     const float* g_rgb, const float* g_sigma, float* grads, float* g_x, float* scratch, cudaStream_t s) {
    const int H = h->hidden, D = h->in_dim, P = 2 * h->depth;
    const long long KCH = 4096;
    float* dzs = scratch;
    float* dzr = scratch + n;
    float* buf[2] = {scratch + 4 * n, scratch + 4 * n + n * (long long)H};
    TN_TRY(launch_head_grad(rgb, sigma, g_rgb, g_sigma, n, dzs, dzr, s));
    const float* h_last = acts + (long long)(h->depth - 1) * n * H;
    // head weight / bias gradients
    for (int head = 0; head < 2; ++head) {
        const int rows = head ? 3 : 1;
        const float* dz = head ? dzr : dzs;
        GemmArgs w;
        w.A = dz; w.lda = rows; w.B = h_last; w.ldb = H; w.C = grads + h->offsets[P + 2 * head]; w.ldc = H;
        w.M = rows; w.N = H; w.K = n; w.flags = GEMM_ATOMIC; w.k_chunk = KCH;
        TN_TRY(launch_gemm(w, false, false, s));
        TN_TRY(launch_colsum(dz, n, rows, rows, grads + h->offsets[P + 2 * head + 1], s));
    }
    // dH_last = dzs * W_sigma + dzr * W_rgb, masked by relu
    {
        GemmArgs d;
        d.A = dzs; d.lda = 1; d.K = 1; d.B = h->params[P]; d.ldb = H; d.C = buf[0]; d.ldc = H; d.M = n; d.N = H;
        TN_TRY(launch_gemm(d, true, false, s));
        d.A = dzr; d.lda = 3; d.K = 3; d.B = h->params[P + 2]; d.flags = GEMM_ACCUM; d.mask = h_last; d.ldm = H;
        TN_TRY(launch_gemm(d, true, false, s));
    }
    int cur = 0;
    bool gx_written = false;
    for (int i = h->depth - 1; i >= 0; --i) {
        const float* dZ = buf[cur];
        const float* W = h->params[2 * i];
        const int fan = h->layer_in[i];
        const float* prev = (i == 0) ? x : acts + (long long)(i - 1) * n * H;
        const int Kp = (i == 0) ? D : H;
        GemmArgs w;
        w.A = dZ; w.lda = H; w.B = prev; w.ldb = Kp; w.C = grads + h->offsets[2 * i]; w.ldc = fan;
        w.M = H; w.N = Kp; w.K = n; w.flags = GEMM_ATOMIC; w.k_chunk = KCH;
        TN_TRY(launch_gemm(w, false, false, s));
        if (skip_into(h, i)) {
            GemmArgs w2 = w;
            w2.B = x; w2.ldb = D; w2.C = grads + h->offsets[2 * i] + H; w2.N = D;
            TN_TRY(launch_gemm(w2, false, false, s));
            if (g_x) {
                GemmArgs d;
                d.A = dZ; d.lda = H; d.K = H; d.B = W + H; d.ldb = fan; d.C = g_x; d.ldc = D; d.M = n; d.N = D;
                d.flags = gx_written ? GEMM_ACCUM : 0;
                TN_TRY(launch_gemm(d, true, false, s));
                gx_written = true;
            }
        }
        TN_TRY(launch_colsum(dZ, n, H, H, grads + h->offsets[2 * i + 1], s));
        if (i > 0) {
            GemmArgs d;
            d.A = dZ; d.lda = H; d.K = H; d.B = W; d.ldb = fan; d.C = buf[cur ^ 1]; d.ldc = H; d.M = n; d.N = H;
            d.mask = prev; d.ldm = H;
            TN_TRY(launch_gemm(d, true, false, s));
            cur ^= 1;
        } else if (g_x) {
            GemmArgs d;
            d.A = dZ; d.lda = H; d.K = H; d.B = W; d.ldb = fan; d.C = g_x; d.ldc = D; d.M = n; d.N = D;
            d.flags = gx_written ? GEMM_ACCUM : 0;
            TN_TRY(launch_gemm(d, true, false, s));
            gx_written = true;
        }
    }
    return 0;
}

}  // namespace tnerf
