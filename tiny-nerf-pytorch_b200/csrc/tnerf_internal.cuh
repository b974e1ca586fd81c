// Internal declarations shared by the translation units of libtnerf.so.
#pragma once
#include <atomic>
#include <string>
#include <vector>
#include "tnerf_common.cuh"

namespace tnerf {

enum { GEMM_ACCUM = 1, GEMM_RELU = 2, GEMM_SIGMOID = 4, GEMM_ATOMIC = 8 };

struct GemmArgs {
    const float* A = nullptr; long long lda = 0;
    const float* B = nullptr; long long ldb = 0;
    float* C = nullptr;       long long ldc = 0;
    long long M = 0; int N = 0; long long K = 0;
    const float* bias = nullptr;
    const float* mask = nullptr; long long ldm = 0;
    int flags = 0;
    long long k_chunk = 0;
};

extern std::atomic<long long> g_launches;
// which launches use programmatic stream serialisation: 1 = training kernel, 2 = gradient scatter, 4 = optimiser (TNERF_PDL, default 7)
int pdl_mask();
void set_error(const std::string& msg);
// records a launch, returns the pending cudaError_t (0 if none)
int count_launch();

// How the optimiser launch reads the gradient straight out of the training kernel's sum vector (tensor-memory order: element
// (column c, row m) of a block at c * 128 + m) -- tnerf_train_fwd_bwd with grads = NULL, tnerf_optimizer_step with repack bit 1.
struct GatherSeg { int sb, C, ld, tile0; long long pb; };   // weight block: param[pb + m * ld + c] <- sum[sb + c * 128 + m], c < C, m < 128; 32 x 32 tiles
struct GatherVec { int sb, n, first; long long pb; };       // vector: param[pb + j] <- sum[sb + j], j < n (biases, head weights); `first` = position in the vector pass
struct GatherPlan {
    int valid = 0, n_seg = 0, n_vec = 0, n_tiles = 0, n_vec_elems = 0, hb = 0;
    GatherSeg seg[6];
    GatherVec vec[8];
    long long pb_hb[4];                                     // head biases: param[pb_hb[o]] <- sum[hb + o] + sum[hb + o + 4] + sum[hb + o + 8] + sum[hb + o + 12]
    long long n = 0;
};

struct Workspace {   // grow-only device scratch owned by a handle
    void* ptr = nullptr; size_t bytes = 0;
    int reserve(size_t need);
    void release();
};

}  // namespace tnerf

struct tnerf_handle {
    int device = 0, in_dim = 0, hidden = 0, depth = 0, skip_at = 0;
    int n_params = 0;
    long long param_count = 0;
    std::vector<const float*> params;     // 2*depth+4 device pointers, state_dict order
    std::vector<long long> offsets;       // offset of each param in the flat gradient vector
    std::vector<int> layer_in;            // fan-in of each hidden layer
    tnerf::Workspace ws;                  // fp32 path scratch
    // tensor-core path
    void* packed = nullptr; size_t packed_bytes = 0;   // fp16 operand image (device)
    void* slabs = nullptr;  size_t slab_bytes = 0;     // per-CTA partial weight gradients
    bool slab0_zero = false;                           // slab 0 is all zeros (it is the accumulation target of the bulk-reduction mode)
    bool slab_pending = false;                         // slab 0 holds an unscaled gradient sum waiting for the gathering optimiser launch
    tnerf::GatherPlan gplan;                           // built by tnerf_train_fwd_bwd with grads = NULL (tnerf_train.cu)
    float* ext_sum = nullptr;                          // caller-owned sum vector (tnerf_set_sum_buffer: peer-mapped memory of a multi-rank step)
    int sum_total = 0;                                 // elements of the sum vector for this model (0 until known)
    int sm_count = 0;
    long long wide_version = 0;           // bumped by every pack of the hidden=256 image (the kernel's constant table follows it)
    bool fused_ok = false;
    int num_freqs = 0;                    // (in_dim-3)/6 when in_dim = 3+6L
    void* debug = nullptr;                // optional device buffer (1024 int64) for kernel phase stamps
    const int* tile_order = nullptr; int tile_order_n = 0;   // caller-owned device permutation of the training kernel's CTAs (tnerf_set_tile_order)
    // training-kernel schedule (tnerf_set_option; defaults from TNERF_TRAIN_SYNC / TNERF_BULK_REDUCE / TNERF_TRAIN_UNROLL_FROM, read
    // ONCE when the handle is created): -1 = built-in choice
    int opt_train_sync = -1, opt_bulk_reduce = -1, opt_unroll_from = -1;
    float* jitter_scratch = nullptr; size_t jitter_scratch_bytes = 0;   // fp32 path: materialised in-kernel jitter
    float* auto_scale = nullptr;          // device float: loss scale chosen from the upstream gradients (tnerf_render_bwd, automatic mode)
};

namespace tnerf {
int launch_get_rays(int H, int W, float focal, const float* c2w, long long first, long long n, float* ro, float* rd, cudaStream_t s);
int launch_gather3(const long long* idx, long long n, long long n_src, const float* sa, float* da, const float* sb, float* db, const float* sc, float* dc, cudaStream_t s);
int launch_stratified(const float* ro, long long os, const float* rd, long long n, int S, float nr, float fr, const float* nray, const float* fray, const float* jit, float* z, float* pts, cudaStream_t s);
int launch_posenc(const float* x, long long n, int L, int inc, float* out, cudaStream_t s);
int launch_posenc_bwd(const float* x, const float* g, long long n, int L, int inc, float* gx, cudaStream_t s);
int launch_gemm(const GemmArgs& g, bool a_kc, bool b_kc, cudaStream_t s);
int launch_colsum(const float* a, long long rows, int cols, long long lda, float* out, cudaStream_t s);
int launch_head_grad(const float* rgb, const float* sigma, const float* g_rgb, const float* g_sigma, long long n, float* dzs, float* dzr, cudaStream_t s);
int launch_composite_fwd(const float* rgb, const float* sigma, const float* z, long long zs, const float* rd, long long n, int S, int white, float* comp, float* depth, float* acc, float* w, cudaStream_t s);
int launch_composite_bwd(const float* rgb, const float* sigma, const float* z, long long zs, const float* rd, long long n, int S, int white, const float* gC, const float* gD, const float* gA, const float* gW, float* g_rgb, float* g_sigma, cudaStream_t s);
int launch_mse_psnr(const float* a, const float* b, long long n, float* out2, cudaStream_t s);
int launch_adam(float* p, const float* g, float* m, float* v, long long n, int step, float lr, float b1, float b2, float eps, float inv_scale, const int* found_inf, cudaStream_t s);
int launch_check_finite(const float* g, long long n, int* flag, cudaStream_t s);
int launch_jitter_fill(unsigned long long seed, unsigned long long step, long long n, int S, float* out, cudaStream_t s);

int launch_mse_grad(const float* c, const float* t, long long n3, float inv_denom, float* gC, float* loss, cudaStream_t s);

// fp32 MLP (tnerf_mlp.cu)
int mlp_forward_f32(tnerf_handle* h, const float* x, long long n, float* rgb, float* sigma, float* acts, float* tmp, cudaStream_t s);
int mlp_backward_f32(tnerf_handle* h, const float* x, long long n, const float* acts, const float* rgb, const float* sigma,
                     const float* g_rgb, const float* g_sigma, float* grads, float* g_x, float* scratch, cudaStream_t s);
long long mlp_bwd_scratch_floats(const tnerf_handle* h, long long n);

// tensor-core fused path (tnerf_fused.cu)
bool fused_shape_supported(const tnerf_handle* h);
int fused_pack_weights(tnerf_handle* h, cudaStream_t s);
int fused_render_fwd(tnerf_handle* h, const RaySource& rs, long long n, float nr, float fr, int S, const float* jitter, int white,
                     float* comp, float* depth, float* acc, float* weights, float* rays_d_out, cudaStream_t s);
int fused_train(tnerf_handle* h, const RaySource& rs, long long n, float nr, float fr, int S, const float* jitter, int white,
                const float* target, float loss_denom, const float* gC, const float* gD, const float* gA, const float* gW,
                float grad_scale, const float* grad_scale_dev, float* found, float* comp, float* loss_sum, float* grads, cudaStream_t s);
int launch_found_inf(const float* g, long long n, float* found, cudaStream_t s);
// wide MLP (hidden = 256) on CTA pairs (tnerf_fused_wide.cu)
bool wide_shape_supported(const tnerf_handle* h);
bool fused_render_shape_ok(const tnerf_handle* h, int S, bool want_weights);   // tnerf_fused.cu
int fused_train_sum_elems(tnerf_handle* h);                                    // tnerf_train.cu
int wide_pack_weights(tnerf_handle* h, cudaStream_t s);
void wide_release(tnerf_handle* h);
int fused_render_fwd_wide(tnerf_handle* h, const RaySource& rs, long long n, float nr, float fr, int S, const float* jitter, int white,
                          float* comp, float* depth, float* acc, float* weights, float* rays_d_out, cudaStream_t s);
int umma_rate(int n, int reps, int variant, long long* out, cudaStream_t s);
int umma_selftest(const float* a, const float* b, int n, int k, int mode, float* d, cudaStream_t s);
}  // namespace tnerf
