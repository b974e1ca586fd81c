// C ABI of libtnerf.so (declared in include/tnerf.h).  Argument checking, the handle, and the fp32
// composition of the fused entry points out of the stand-alone kernels (TNERF_PREC_F32_SIMT).
#include "../../include/tnerf.h"
#include "tnerf_fused.cuh"

#include <cstdio>
#include <cstring>
#include <cstdint>
#include <cstdlib>

namespace tnerf {

std::atomic<long long> g_launches{0};
int pdl_mask() { static const int mask = [] { const char* e = getenv("TNERF_PDL"); return e ? atoi(e) : 7; }(); return mask; }   // read once per process
static thread_local std::string t_error;
void set_error(const std::string& msg) { t_error = msg; }
int count_launch() {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error(std::string("kernel launch failed: ") + cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}
int Workspace::reserve(size_t need) {
    if (need <= bytes) return 0;
    if (ptr) cudaFree(ptr);
    ptr = nullptr; bytes = 0;
    size_t want = need + need / 4;
    cudaError_t e = cudaMalloc(&ptr, want);
    if (e != cudaSuccess) { set_error("workspace cudaMalloc failed"); return (int)e; }
    bytes = want;
    return 0;
}
void Workspace::release() { if (ptr) cudaFree(ptr); ptr = nullptr; bytes = 0; }

static int bad(const char* msg) { set_error(msg); return -1; }
// Every entry point that takes a handle runs on the handle's device whatever the caller's current device is, and leaves the
// caller's device selected on return (torch modules work on any device; so must their replacements).
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int want) {
        if (want < 0) return;
        int cur = -1;
        if (cudaGetDevice(&cur) == cudaSuccess && cur != want && cudaSetDevice(want) == cudaSuccess) prev = cur;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define TN_ON_DEVICE(h) DeviceGuard device_guard_((h) ? (h)->device : -1)
// pointer-only entry points: the device that owns one of the (required) buffers
static int device_of(const void* p) {
    cudaPointerAttributes a{};
    if (!p || cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return -1; }
    return (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) ? a.device : -1;
}
#define TN_ON_DEVICE_OF(ptr) DeviceGuard device_guard_(device_of(ptr))
static RaySource to_device_source(const tnerf_ray_source* r) {
    RaySource s;
    s.rays_o = r->rays_o; s.o_stride = r->o_stride; s.rays_d = r->rays_d; s.c2w = r->c2w; s.H = r->H; s.W = r->W;
    s.focal = r->focal; s.pixel_index = r->pixel_index; s.first_ray = r->first_ray; s.frame_rays = 0;
    s.jitter_seed = r->jitter_seed; s.jitter_step = r->jitter_step;
    return s;
}
static int check_source(const tnerf_ray_source* r) {
    if (!r) return bad("ray source is NULL");
    if (r->rays_d) { if (!r->rays_o) return bad("rays_o is NULL"); return 0; }
    if (!r->c2w || r->H <= 0 || r->W <= 0 || !(r->focal > 0.f)) return bad("camera ray source needs c2w, H, W, focal");
    return 0;
}

// materialise rays of a source into dense (n,3) buffers (used by the fp32 composition only)
__global__ void expand_rays_kernel(RaySource rs, long long first, long long n, float* __restrict__ ro, float* __restrict__ rd) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float o[3], d[3];
    load_ray(rs, first + i, o, d);
#pragma unroll
    for (int c = 0; c < 3; ++c) { ro[3 * i + c] = o[c]; rd[3 * i + c] = d[c]; }
}

// fp32 exact-mode pipeline over chunks of rays; mode 0 = render, 1 = backward from upstream grads, 2 = MSE train step
struct F32Job {
    int mode;
    const float *target, *gC, *gD, *gA, *gW;
    float loss_denom;
    float *comp, *depth, *acc, *weights, *rays_d_out, *loss_sum, *grads;
};

static int run_f32(tnerf_handle* h, const RaySource& rs, long long n, float nr, float fr, int S, const float* jitter, int white,
                   const F32Job& job, cudaStream_t s) {
    if (S > 256 && job.mode != 0) return bad("fp32 backward path supports n_samples <= 256");
    const int H = h->hidden, D = h->in_dim, depth = h->depth;
    const int L = h->num_freqs, inc = (h->in_dim == 6 * L + 3);
    if (L < 0 || (h->in_dim != 6 * L + 3 && h->in_dim != 6 * L)) return bad("handle in_dim is not 6L or 6L+3: set the encoding with tnerf_set_encoding");
    long long chunk = 4096;
    while (chunk > 64 && chunk * S > (1 << 19)) chunk >>= 1;
    const long long m = chunk * S;
    // floats: ro 3c, rd 3c, z m, pts 3m, enc D m, acts depth*H*m, rgb 3m, sigma m, comp/depth/acc 5c, gC 3c, g_rgb 3m, g_sigma m, scratch
    const bool bw = job.mode != 0;
    size_t fl = 6 * chunk + m + 3 * m + (size_t)D * m + (size_t)(bw ? depth : 2) * H * m + 4 * m + 8 * chunk;
    if (bw) fl += 3 * chunk + 4 * m + (size_t)mlp_bwd_scratch_floats(h, m);
    int e = h->ws.reserve(fl * sizeof(float));
    if (e) return e;
    float* w = reinterpret_cast<float*>(h->ws.ptr);
    float* ro = w; w += 3 * chunk;
    float* rd = w; w += 3 * chunk;
    float* z = w; w += m;
    float* pts = w; w += 3 * m;
    float* enc = w; w += (size_t)D * m;
    float* acts = w; w += (size_t)(bw ? depth : 2) * H * m;
    float* rgb = w; w += 3 * m;
    float* sig = w; w += m;
    float* c_comp = w; w += 3 * chunk;
    float* c_depth = w; w += chunk;
    float* c_acc = w; w += chunk;
    w += 3 * chunk;
    float *gC = nullptr, *g_rgb = nullptr, *g_sig = nullptr, *scratch = nullptr;
    if (bw) { gC = w; w += 3 * chunk; g_rgb = w; w += 3 * m; g_sig = w; w += m; scratch = w; }

    for (long long a = 0; a < n; a += chunk) {
        const long long c = (n - a < chunk) ? n - a : chunk;
        expand_rays_kernel<<<(unsigned)((c + 255) / 256), 256, 0, s>>>(rs, a, c, ro, rd);
        if ((e = count_launch())) return e;
        if ((e = launch_stratified(ro, 3, rd, c, S, nr, fr, nullptr, nullptr, jitter ? jitter + a * S : nullptr, z, pts, s))) return e;
        if ((e = launch_posenc(pts, c * S, L, inc, enc, s))) return e;
        if ((e = mlp_forward_f32(h, enc, c * S, rgb, sig, bw ? acts : nullptr, acts, s))) return e;
        float* o_comp = job.comp ? job.comp + 3 * a : c_comp;
        if ((e = launch_composite_fwd(rgb, sig, z, S, rd, c, S, white, o_comp, job.depth ? job.depth + a : c_depth,
                                      job.acc ? job.acc + a : c_acc, job.weights ? job.weights + a * S : nullptr, s))) return e;
        if (job.rays_d_out) cudaMemcpyAsync(job.rays_d_out + 3 * a, rd, 3 * c * sizeof(float), cudaMemcpyDeviceToDevice, s);
        if (!bw) continue;
        const float* up_c = job.gC ? job.gC + 3 * a : nullptr;
        if (job.mode == 2) {
            if ((e = launch_mse_grad(o_comp, job.target + 3 * a, 3 * c, 1.f / job.loss_denom, gC, job.loss_sum, s))) return e;
            up_c = gC;
        }
        if ((e = launch_composite_bwd(rgb, sig, z, S, rd, c, S, white, up_c, job.gD ? job.gD + a : nullptr, job.gA ? job.gA + a : nullptr,
                                      job.gW ? job.gW + a * S : nullptr, g_rgb, g_sig, s))) return e;
        if ((e = mlp_backward_f32(h, enc, c * S, acts, rgb, sig, g_rgb, g_sig, job.grads, nullptr, scratch, s))) return e;
    }
    return 0;
}

}  // namespace tnerf

using namespace tnerf;

extern "C" {

int tnerf_abi_version(void) { return TNERF_ABI_VERSION; }
const char* tnerf_last_error(void) { return t_error.c_str(); }
long long tnerf_launch_count(void) { return g_launches.load(); }

int tnerf_get_rays(int H, int W, float focal, const float* c2w, long long first_ray, long long n_rays, float* rays_o,
                   float* rays_d, void* stream) {
    TN_ON_DEVICE_OF(rays_d);
    if (H <= 0 || W <= 0 || !(focal > 0.f) || !c2w || !rays_d || first_ray < 0 || n_rays < 0 || first_ray + n_rays > (long long)H * W)
        return bad("tnerf_get_rays: invalid argument");
    return launch_get_rays(H, W, focal, c2w, first_ray, n_rays, rays_o, rays_d, (cudaStream_t)stream);
}

int tnerf_gather3(const long long* index, long long n, long long n_src, const float* sa, float* da, const float* sb, float* db,
                  const float* sc, float* dc, void* stream) {
    TN_ON_DEVICE_OF(index);
    if (n == 0) return 0;
    if (!index || n < 0 || (sa && !da) || (sb && !db) || (sc && !dc)) return bad("tnerf_gather3: invalid argument");
    return launch_gather3(index, n, n_src, sa, da, sb, db, sc, dc, (cudaStream_t)stream);
}

int tnerf_stratified(const float* rays_o, long long o_stride, const float* rays_d, long long n_rays, int n_samples, float near_,
                     float far_, const float* near_ray, const float* far_ray, const float* jitter, float* z_vals, float* pts,
                     void* stream) {
    TN_ON_DEVICE_OF(z_vals ? (const void*)z_vals : (const void*)pts);
    if (n_rays == 0) return 0;
    if (n_rays < 0 || n_samples < 1 || (pts && (!rays_o || !rays_d))) return bad("tnerf_stratified: invalid argument");
    return launch_stratified(rays_o, o_stride, rays_d, n_rays, n_samples, near_, far_, near_ray, far_ray, jitter, z_vals, pts,
                             (cudaStream_t)stream);
}

int tnerf_posenc(const float* x, long long n_pts, int num_freqs, int include_input, float* out, void* stream) {
    TN_ON_DEVICE_OF(out);
    if (n_pts == 0) return 0;
    if (!x || !out || n_pts < 0 || num_freqs < 0 || num_freqs > 30) return bad("tnerf_posenc: invalid argument");
    return launch_posenc(x, n_pts, num_freqs, include_input, out, (cudaStream_t)stream);
}
int tnerf_posenc_bwd(const float* x, const float* g_out, long long n_pts, int num_freqs, int include_input, float* g_x, void* stream) {
    TN_ON_DEVICE_OF(g_x);
    if (n_pts == 0) return 0;
    if (!x || !g_out || !g_x || n_pts < 0 || num_freqs < 0 || num_freqs > 30) return bad("tnerf_posenc_bwd: invalid argument");
    return launch_posenc_bwd(x, g_out, n_pts, num_freqs, include_input, g_x, (cudaStream_t)stream);
}

int tnerf_create(tnerf_handle** out, int device, int in_dim, int hidden, int depth, int skip_at) {
    if (!out || in_dim < 1 || hidden < 1 || depth < 1 || depth > kMaxDepth) return bad("tnerf_create: invalid shape (1 <= depth <= 8)");
    if (skip_at == depth) return bad("tnerf_create: skip_at == depth leaves the heads with hidden+in_dim inputs (the reference errors too)");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || device < 0 || device >= n_dev) return bad("tnerf_create: no such CUDA device");
    DeviceGuard device_guard_(device);          // the caller's current device is left as it was
    tnerf_handle* h = new tnerf_handle();
    h->device = device; h->in_dim = in_dim; h->hidden = hidden; h->depth = depth; h->skip_at = skip_at;
    h->n_params = 2 * depth + 4;
    long long off = 0;
    int last = in_dim;
    for (int i = 0; i < depth; ++i) {
        h->layer_in.push_back(last);
        h->offsets.push_back(off); off += (long long)hidden * last;
        h->offsets.push_back(off); off += hidden;
        last = (i == skip_at - 1) ? hidden + in_dim : hidden;
    }
    h->offsets.push_back(off); off += hidden;       // sigma.0.weight
    h->offsets.push_back(off); off += 1;            // sigma.0.bias
    h->offsets.push_back(off); off += 3LL * hidden; // rgb.0.weight
    h->offsets.push_back(off); off += 3;            // rgb.0.bias
    h->param_count = off;
    h->num_freqs = (in_dim >= 3 && (in_dim - 3) % 6 == 0) ? (in_dim - 3) / 6 : (in_dim % 6 == 0 ? in_dim / 6 : -1);
    cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device);
    h->fused_ok = fused_shape_supported(h);
    // developer knobs are environment DEFAULTS read once here (a getenv per launch is host time on a 160 us step); tnerf_set_option changes them
    if (const char* e = getenv("TNERF_TRAIN_SYNC")) h->opt_train_sync = atoi(e);
    if (const char* e = getenv("TNERF_BULK_REDUCE")) h->opt_bulk_reduce = atoi(e);
    if (const char* e = getenv("TNERF_TRAIN_UNROLL_FROM")) h->opt_unroll_from = atoi(e);
    *out = h;
    return 0;
}
void tnerf_destroy(tnerf_handle* h) {
    TN_ON_DEVICE(h);
    if (!h) return;
    wide_release(h);
    h->ws.release();
    if (h->packed) cudaFree(h->packed);
    if (h->slabs) cudaFree(h->slabs);
    if (h->auto_scale) cudaFree(h->auto_scale);
    if (h->jitter_scratch) cudaFree(h->jitter_scratch);
    delete h;
}
int tnerf_set_encoding(tnerf_handle* h, int num_freqs, int include_input) {
    if (!h || num_freqs < 0 || h->in_dim != 6 * num_freqs + (include_input ? 3 : 0)) return bad("tnerf_set_encoding: in_dim mismatch");
    h->num_freqs = num_freqs;
    h->fused_ok = fused_shape_supported(h);
    return 0;
}
int tnerf_set_option(tnerf_handle* h, const char* name, int value) {
    if (!h || !name) return bad("tnerf_set_option: NULL argument");
    const std::string n(name);
    if (n == "train_sync") h->opt_train_sync = value;
    else if (n == "bulk_reduce") h->opt_bulk_reduce = value;
    else if (n == "unroll_from") h->opt_unroll_from = value;
    else return bad("tnerf_set_option: unknown option");
    return 0;
}
long long tnerf_sum_elems(tnerf_handle* h) { return h ? fused_train_sum_elems(h) : -1; }
int tnerf_set_sum_buffer(tnerf_handle* h, float* buf) {
    if (!h) return bad("tnerf_set_sum_buffer: NULL handle");
    if (buf && (reinterpret_cast<uintptr_t>(buf) & 15)) return bad("tnerf_set_sum_buffer: the vector must be 16-byte aligned");
    h->ext_sum = buf;
    return 0;
}
int tnerf_clear_sum(tnerf_handle* h, void* stream) {
    TN_ON_DEVICE(h);
    if (!h) return bad("tnerf_clear_sum: NULL handle");
    if (h->slabs && h->sum_total > 0) cudaMemsetAsync(h->slabs, 0, (size_t)h->sum_total * sizeof(float), (cudaStream_t)stream);
    h->slab_pending = false; h->slab0_zero = h->slabs != nullptr;
    return 0;
}
int tnerf_get_option(const tnerf_handle* h, const char* name) {
    if (!h || !name) return -1;
    const std::string n(name);
    const int sync = h->opt_train_sync >= 0 ? (h->opt_train_sync != 0) : 1;
    if (n == "train_sync") return sync;
    if (n == "bulk_reduce") return h->opt_bulk_reduce >= 0 ? (h->opt_bulk_reduce != 0) : sync;
    if (n == "unroll_from") return h->opt_unroll_from;
    return -1;
}
int tnerf_set_tile_order(tnerf_handle* h, const int* order_dev, int n) {
    if (!h || n < 0 || (n > 0 && !order_dev)) return bad("tnerf_set_tile_order: NULL handle / negative length / NULL table");
    h->tile_order = n > 0 ? order_dev : nullptr;
    h->tile_order_n = n;
    return 0;
}
int tnerf_set_debug_buffer(tnerf_handle* h, void* buf) {
    if (!h) return bad("NULL handle");
    h->debug = buf;
    return 0;
}
int tnerf_bind_params(tnerf_handle* h, const float* const* params_host, int n_params) {
    if (!h || !params_host || n_params != h->n_params) return bad("tnerf_bind_params: expected 2*depth+4 pointers");
    h->params.assign(params_host, params_host + n_params);
    for (auto p : h->params) if (!p) return bad("tnerf_bind_params: NULL parameter");
    return 0;
}
long long tnerf_param_count(const tnerf_handle* h) { return h ? h->param_count : -1; }
int tnerf_fused_supported(const tnerf_handle* h) { return h && h->fused_ok; }
int tnerf_pack_weights(tnerf_handle* h, void* stream) {
    TN_ON_DEVICE(h);
    if (!h) return bad("NULL handle");
    return fused_pack_weights(h, (cudaStream_t)stream);
}

int tnerf_mlp_fwd(tnerf_handle* h, const float* x, long long n, float* rgb, float* sigma, float* acts, void* stream) {
    TN_ON_DEVICE(h);
    if (n == 0) return 0;
    if (!h || h->params.empty() || !x || !rgb || !sigma || n < 0) return bad("tnerf_mlp_fwd: invalid argument / params not bound");
    float* tmp = nullptr;
    if (!acts) {
        int e = h->ws.reserve((size_t)2 * n * h->hidden * sizeof(float));
        if (e) return e;
        tmp = reinterpret_cast<float*>(h->ws.ptr);
    }
    return mlp_forward_f32(h, x, n, rgb, sigma, acts, tmp, (cudaStream_t)stream);
}
long long tnerf_mlp_bwd_scratch_floats(const tnerf_handle* h, long long n) { return h ? mlp_bwd_scratch_floats(h, n) : -1; }
int tnerf_mlp_bwd(tnerf_handle* h, const float* x, long long n, const float* acts, const float* rgb, const float* sigma,
                  const float* g_rgb, const float* g_sigma, float* grads, float* g_x, float* scratch, void* stream) {
    TN_ON_DEVICE(h);
    if (n == 0) return 0;
    if (!h || h->params.empty() || !x || !acts || !rgb || !sigma || !grads || !scratch || n < 0) return bad("tnerf_mlp_bwd: invalid argument");
    return mlp_backward_f32(h, x, n, acts, rgb, sigma, g_rgb, g_sigma, grads, g_x, scratch, (cudaStream_t)stream);
}

int tnerf_composite_fwd(const float* rgb, const float* sigma, const float* z_vals, long long z_stride, const float* rays_d,
                        long long n_rays, int n_samples, int white_bkgd, float* comp_rgb, float* depth, float* acc, float* weights,
                        void* stream) {
    TN_ON_DEVICE_OF(comp_rgb);
    if (n_rays == 0) return 0;
    if (!rgb || !sigma || !z_vals || !rays_d || !comp_rgb || n_rays < 0 || n_samples < 1) return bad("tnerf_composite_fwd: invalid argument");
    return launch_composite_fwd(rgb, sigma, z_vals, z_stride, rays_d, n_rays, n_samples, white_bkgd, comp_rgb, depth, acc, weights,
                                (cudaStream_t)stream);
}
int tnerf_composite_bwd(const float* rgb, const float* sigma, const float* z_vals, long long z_stride, const float* rays_d,
                        long long n_rays, int n_samples, int white_bkgd, const float* g_comp, const float* g_depth,
                        const float* g_acc, const float* g_weights, float* g_rgb, float* g_sigma, void* stream) {
    TN_ON_DEVICE_OF(rgb);
    if (n_rays == 0) return 0;
    if (!rgb || !sigma || !z_vals || !rays_d || n_rays < 0 || n_samples < 1 || n_samples > 256)
        return bad("tnerf_composite_bwd: invalid argument (n_samples <= 256)");
    return launch_composite_bwd(rgb, sigma, z_vals, z_stride, rays_d, n_rays, n_samples, white_bkgd, g_comp, g_depth, g_acc, g_weights,
                                g_rgb, g_sigma, (cudaStream_t)stream);
}

int tnerf_render_fwd(tnerf_handle* h, const tnerf_ray_source* rays_host, long long n_rays, float near_, float far_, int n_samples,
                     const float* jitter, int white_bkgd, int precision, float* comp_rgb, float* depth, float* acc, float* weights,
                     float* rays_d_out, void* stream) {
    TN_ON_DEVICE(h);
    if (n_rays == 0) return 0;
    if (!h || h->params.empty() || !comp_rgb || n_rays < 0 || n_samples < 1) return bad("tnerf_render_fwd: invalid argument / params not bound");
    if (int e = check_source(rays_host)) return e;
    const RaySource rs = to_device_source(rays_host);
    // tensor-core kernels where they cover the shape (n_samples a multiple of 32, ...); other shapes take the exact fp32 path
    if (precision == TNERF_PREC_F16_TC && fused_render_shape_ok(h, n_samples, weights != nullptr))
        return fused_render_fwd(h, rs, n_rays, near_, far_, n_samples, jitter, white_bkgd, comp_rgb, depth, acc, weights, rays_d_out,
                                (cudaStream_t)stream);
    F32Job job{};
    job.mode = 0; job.comp = comp_rgb; job.depth = depth; job.acc = acc; job.weights = weights; job.rays_d_out = rays_d_out;
    return run_f32(h, rs, n_rays, near_, far_, n_samples, jitter, white_bkgd, job, (cudaStream_t)stream);
}

int tnerf_render_frames(tnerf_handle* h, const float* poses, int n_poses, int H, int W, float focal, long long first_ray,
                        long long rays_per_pose, float near_, float far_, int n_samples, int white_bkgd, int precision, float* comp_rgb,
                        float* depth, float* acc, void* stream) {
    TN_ON_DEVICE(h);
    if (n_poses == 0 || rays_per_pose == 0) return 0;
    if (!h || h->params.empty() || !poses || !comp_rgb || n_poses < 0 || H <= 0 || W <= 0 || !(focal > 0.f) || first_ray < 0 ||
        rays_per_pose < 0 || first_ray + rays_per_pose > (long long)H * W || n_samples < 1)
        return bad("tnerf_render_frames: invalid argument / params not bound");
    const long long total = (long long)n_poses * rays_per_pose;
    RaySource rs{};
    rs.c2w = poses; rs.H = H; rs.W = W; rs.focal = focal; rs.first_ray = first_ray; rs.frame_rays = rays_per_pose;
    // one launch for the whole pose batch where the kernel indexes the pose per ray (role-split render kernels: n_samples % 32 == 0);
    // other shapes: one launch per pose from here
    const bool tc = precision == TNERF_PREC_F16_TC && fused_render_shape_ok(h, n_samples, false);
    const bool one_launch = tc && total < (1ll << 31);
    if (one_launch)
        return fused_render_fwd(h, rs, total, near_, far_, n_samples, nullptr, white_bkgd, comp_rgb, depth, acc, nullptr, nullptr, (cudaStream_t)stream);
    rs.frame_rays = 0;
    for (int f = 0; f < n_poses; ++f) {
        rs.c2w = poses + 16 * f;
        float* c = comp_rgb + 3 * f * rays_per_pose;
        float* d = depth ? depth + f * rays_per_pose : nullptr;
        float* a = acc ? acc + f * rays_per_pose : nullptr;
        int e;
        if (tc)
            e = fused_render_fwd(h, rs, rays_per_pose, near_, far_, n_samples, nullptr, white_bkgd, c, d, a, nullptr, nullptr, (cudaStream_t)stream);
        else {
            F32Job job{};
            job.mode = 0; job.comp = c; job.depth = d; job.acc = a;
            e = run_f32(h, rs, rays_per_pose, near_, far_, n_samples, nullptr, white_bkgd, job, (cudaStream_t)stream);
        }
        if (e) return e;
    }
    return 0;
}

int tnerf_render_bwd(tnerf_handle* h, const tnerf_ray_source* rays_host, long long n_rays, float near_, float far_, int n_samples,
                     const float* jitter, int white_bkgd, int precision, const float* g_comp, const float* g_depth, const float* g_acc,
                     const float* g_weights, float grad_scale, const float* grad_scale_dev, float* grads, void* stream) {
    TN_ON_DEVICE(h);
    if (n_rays == 0) return 0;
    if (!h || h->params.empty() || !grads || n_rays < 0 || n_samples < 1) return bad("tnerf_render_bwd: invalid argument");
    if (int e = check_source(rays_host)) return e;
    const RaySource rs = to_device_source(rays_host);
    if (precision == TNERF_PREC_F16_TC)
        return fused_train(h, rs, n_rays, near_, far_, n_samples, jitter, white_bkgd, nullptr, 1.f, g_comp, g_depth, g_acc, g_weights,
                           grad_scale, grad_scale_dev, nullptr, nullptr, nullptr, grads, (cudaStream_t)stream);
    F32Job job{};
    job.mode = 1; job.gC = g_comp; job.gD = g_depth; job.gA = g_acc; job.gW = g_weights; job.grads = grads;
    return run_f32(h, rs, n_rays, near_, far_, n_samples, jitter, white_bkgd, job, (cudaStream_t)stream);
}

int tnerf_train_fwd_bwd(tnerf_handle* h, const tnerf_ray_source* rays_host, const float* target, long long n_rays, float near_,
                        float far_, int n_samples, const float* jitter, int white_bkgd, int precision, float loss_denom,
                        float* comp_rgb, float* loss_sum, float* grads, const float* loss_scale_dev, float* found_inf, void* stream) {
    TN_ON_DEVICE(h);
    if (n_rays == 0) return 0;
    if (!h || h->params.empty() || !target || !loss_sum || n_rays < 0 || n_samples < 1 || !(loss_denom > 0.f))
        return bad("tnerf_train_fwd_bwd: invalid argument");
    if (!grads && precision != TNERF_PREC_F16_TC) return bad("tnerf_train_fwd_bwd: grads = NULL (gradient left for the gathering optimiser launch) is a tensor-core path feature");
    if (int e = check_source(rays_host)) return e;
    const RaySource rs = to_device_source(rays_host);
    if (precision == TNERF_PREC_F16_TC)
        return fused_train(h, rs, n_rays, near_, far_, n_samples, jitter, white_bkgd, target, loss_denom, nullptr, nullptr, nullptr,
                           nullptr, 0.f, loss_scale_dev, found_inf, comp_rgb, loss_sum, grads, (cudaStream_t)stream);
    // fp32 path: no loss scale is needed; the overflow flag is the plain non-finite test of the gradient vector and the loss
    F32Job job{};
    job.mode = 2; job.target = target; job.loss_denom = loss_denom; job.comp = comp_rgb; job.loss_sum = loss_sum; job.grads = grads;
    if (!jitter && rs.jitter_seed) {      // in-kernel jitter requested: the fp32 composition reads the same numbers from a scratch tensor
        const size_t bytes = (size_t)n_rays * n_samples * sizeof(float);
        if (h->jitter_scratch_bytes < bytes) {
            if (h->jitter_scratch) cudaFree(h->jitter_scratch);
            h->jitter_scratch = nullptr; h->jitter_scratch_bytes = 0;
            if (cudaMalloc(&h->jitter_scratch, bytes) != cudaSuccess) return bad("cudaMalloc(jitter scratch) failed");
            h->jitter_scratch_bytes = bytes;
        }
        if (int e = launch_jitter_fill(rs.jitter_seed, rs.jitter_step, n_rays, n_samples, h->jitter_scratch, (cudaStream_t)stream)) return e;
        jitter = h->jitter_scratch;
    }
    if (int e = run_f32(h, rs, n_rays, near_, far_, n_samples, jitter, white_bkgd, job, (cudaStream_t)stream)) return e;
    if (found_inf) {
        if (int e = launch_found_inf(grads, h->param_count, found_inf, (cudaStream_t)stream)) return e;
        return launch_found_inf(loss_sum, 1, found_inf, (cudaStream_t)stream);
    }
    return 0;
}

int tnerf_mse_psnr(const float* pred, const float* target, long long n, float* out2, void* stream) {
    TN_ON_DEVICE_OF(out2);
    if (!pred || !target || !out2 || n < 0) return bad("tnerf_mse_psnr: invalid argument");
    return launch_mse_psnr(pred, target, n, out2, (cudaStream_t)stream);
}
int tnerf_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n, int step, float lr, float beta1,
                    float beta2, float eps, float inv_scale, const int* found_inf, void* stream) {
    TN_ON_DEVICE_OF(params);
    if (!params || !grads || !exp_avg || !exp_avg_sq || n < 0 || step < 1) return bad("tnerf_adam_step: invalid argument");
    return launch_adam(params, grads, exp_avg, exp_avg_sq, n, step, lr, beta1, beta2, eps, inv_scale, found_inf, (cudaStream_t)stream);
}
long long tnerf_packed_image_copy(const tnerf_handle* h, void* dst, long long dst_bytes, void* stream) {
    TN_ON_DEVICE(h);
    if (!h || !h->packed) { set_error("tnerf_packed_image_copy: no packed image"); return -1; }
    FusedPlan pl;
    if (!build_plan(h, pl)) { set_error("tnerf_packed_image_copy: unsupported shape"); return -2; }
    if (dst) {
        if (dst_bytes < (long long)pl.image_bytes) { set_error("tnerf_packed_image_copy: destination too small"); return -3; }
        cudaMemcpyAsync(dst, h->packed, pl.image_bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
    }
    return (long long)pl.image_bytes;
}
// the flat vector must be the handle's bound parameters laid out back to back (state_dict order) for the image refresh
static bool params_are_flat(const tnerf_handle* h, const float* flat) {
    if (!h || h->params.empty()) return false;
    for (int t = 0; t < h->n_params; ++t)
        if (h->params[t] != flat + h->offsets[t]) return false;
    return true;
}
static int to_scaler_args(const tnerf_scaler* sc, ScalerArgs& out) {
    out = ScalerArgs{};
    if (!sc) return 0;
    if (!sc->state || (reinterpret_cast<uintptr_t>(sc->state) & 7) || !(sc->growth_factor >= 1.f) || !(sc->backoff_factor > 0.f && sc->backoff_factor <= 1.f) ||
        sc->growth_interval < 1)
        return bad("tnerf_scaler: need an 8-byte aligned state of 16 floats, growth >= 1, 0 < backoff <= 1, interval >= 1");
    out.state = sc->state; out.found = sc->found_inf; out.clear = sc->clear_next; out.growth = sc->growth_factor; out.backoff = sc->backoff_factor;
    out.interval = sc->growth_interval; out.call = sc->call;
    return 0;
}
int tnerf_optimizer_step(tnerf_handle* h, float* params, float* grads, float* exp_avg, float* exp_avg_sq, long long n,
                         long long n_clear, int step, float lr, float beta1, float beta2, float eps, float* tail_out, int repack,
                         const tnerf_scaler* scaler_host, void* stream) {
    DeviceGuard device_guard_(h ? h->device : device_of(params));
    if (!params || !grads || !exp_avg || !exp_avg_sq || n < 0 || n_clear < n || step < 1) return bad("tnerf_optimizer_step: invalid argument");
    const bool gather = (repack & 2) != 0;      // the gradient is the handle's pending sum (tnerf_train_fwd_bwd with grads = NULL)
    if (gather && (!h || !h->gplan.valid || h->gplan.n != n || !h->slabs || n_clear > n + 64))
        return bad("tnerf_optimizer_step: no pending gradient sum (call tnerf_train_fwd_bwd with grads = NULL first)");
    repack &= 1;
    RepackMap mp{};
    if (repack) {
        if (!params_are_flat(h, params) || n != h->param_count) return bad("tnerf_optimizer_step: repack needs the handle's parameters bound as one flat vector");
        if (!build_repack_map(h, mp)) return bad("tnerf_optimizer_step: no packed image to refresh (call tnerf_pack_weights once first)");
    }
    ScalerArgs sc;
    if (int e = to_scaler_args(scaler_host, sc)) return e;
    if (gather) {
        h->slab_pending = false; h->slab0_zero = true;      // every element the training kernel writes is read and cleared by this launch
        return launch_adam_gather(params, grads, exp_avg, exp_avg_sq, n, n_clear, step, lr, beta1, beta2, eps, tail_out, mp, sc, h->gplan,
                                  reinterpret_cast<float*>(h->slabs), (cudaStream_t)stream);
    }
    return launch_adam_fused(params, grads, exp_avg, exp_avg_sq, n, n_clear, step, lr, beta1, beta2, eps, tail_out, mp, sc, (cudaStream_t)stream);
}
int tnerf_allreduce_adam_step(tnerf_handle* h, float* params, float* exp_avg, float* exp_avg_sq, long long n, const float* const* peer_grads,
                              unsigned int* const* peer_flags, int world, int rank, unsigned int epoch, int step, float lr,
                              float beta1, float beta2, float eps, float* reduced_out, float* zero_next, int repack,
                              const tnerf_scaler* scaler_host, void* stream) {
    DeviceGuard device_guard_(h ? h->device : device_of(params));
    if (!params || !exp_avg || !exp_avg_sq || !peer_grads || !peer_flags || n < 0 || step < 1) return bad("tnerf_allreduce_adam_step: invalid argument");
    if (world < 1 || world > 8 || rank < 0 || rank >= world || epoch == 0) return bad("tnerf_allreduce_adam_step: need 1 <= world <= 8, 0 <= rank < world, epoch >= 1");
    for (int r = 0; r < world; ++r)
        if (!peer_grads[r] || !peer_flags[r] || (reinterpret_cast<uintptr_t>(peer_grads[r]) & 15)) return bad("tnerf_allreduce_adam_step: NULL or unaligned (16 B) peer pointer");
    const bool gather = (repack & 2) != 0;      // peer_grads[r] are the ranks' SUM vectors [sum | loss | flag] (tnerf_set_sum_buffer + grads = NULL)
    if (gather && (!h || !h->gplan.valid || h->gplan.n != n || h->sum_total <= 0))
        return bad("tnerf_allreduce_adam_step: gather mode needs a preceding tnerf_train_fwd_bwd with grads = NULL on this handle");
    repack &= 1;
    RepackMap mp{};
    if (repack) {
        if (!params_are_flat(h, params) || n != h->param_count) return bad("tnerf_allreduce_adam_step: repack needs the handle's parameters bound as one flat vector");
        if (!build_repack_map(h, mp)) return bad("tnerf_allreduce_adam_step: no packed image to refresh (call tnerf_pack_weights once first)");
    }
    ScalerArgs sc;
    if (int e = to_scaler_args(scaler_host, sc)) return e;
    // a peer that has not published its vector after this long is reported by a trap instead of a hung device: TNERF_PEER_TIMEOUT_S
    // seconds (default 60; 0 = wait forever) -- rank-local work between steps (checkpoints, previews) must stay below it
    static const long long timeout_cycles = [] {
        const char* e = getenv("TNERF_PEER_TIMEOUT_S");
        const double sec = e ? atof(e) : 60.0;
        return sec > 0.0 ? (long long)(sec * 2.0e9) : 0LL;
    }();
    if (gather)
        return launch_allreduce_adam_gather(params, exp_avg, exp_avg_sq, n, peer_grads, peer_flags, world, rank, epoch, step, lr, beta1, beta2, eps,
                                            reduced_out, zero_next, mp, sc, h->gplan, h->sum_total, timeout_cycles, (cudaStream_t)stream);
    return launch_allreduce_adam(params, exp_avg, exp_avg_sq, n, peer_grads, peer_flags, world, rank, epoch, step, lr, beta1, beta2, eps,
                                 reduced_out, zero_next, mp, sc, timeout_cycles, (cudaStream_t)stream);
}
int tnerf_jitter_fill(unsigned long long seed, unsigned long long step, long long n_rays, int n_samples, float* out, void* stream) {
    TN_ON_DEVICE_OF(out);
    if (!out || n_rays < 0 || n_samples < 1 || seed == 0) return bad("tnerf_jitter_fill: invalid argument (seed must be non-zero)");
    return launch_jitter_fill(seed, step, n_rays, n_samples, out, (cudaStream_t)stream);
}
int tnerf_check_finite(const float* grads, long long n, int* found_inf, void* stream) {
    TN_ON_DEVICE_OF(grads);
    if (!grads || !found_inf || n < 0) return bad("tnerf_check_finite: invalid argument");
    return launch_check_finite(grads, n, found_inf, (cudaStream_t)stream);
}
int tnerf_umma_rate(int n, int reps, int variant, long long* out2, void* stream) {
    if (!out2) return bad("tnerf_umma_rate: invalid argument");
    return umma_rate(n, reps, variant, out2, (cudaStream_t)stream);
}
int tnerf_umma_selftest(const float* a, const float* b, int n, int k, int mode, float* d, void* stream) {
    if (!a || !b || !d) return bad("tnerf_umma_selftest: invalid argument");
    return umma_selftest(a, b, n, k, mode, d, (cudaStream_t)stream);
}

}  // extern "C"
