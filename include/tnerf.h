/*
 * tnerf.h -- C ABI of the B200-native TinyNeRF ray engine (libtnerf.so, sm_100a only).
 *
 * The reference (avihaig/tiny-nerf-pytorch) has no FFI layer: its boundary is the Python call
 * surface of the flat modules rays/sampling/encoding/nerf/volume.  The Python
 * modules under tiny-nerf-pytorch_b200/ keep that surface and call THIS library through ctypes;
 * INTEGRATION.md shows the binding a reference maintainer would add.  Each entry point cites the
 * reference code it replaces (paths relative to the reference repository).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; the caller owns and sizes
 *     every buffer (the library allocates device memory only inside a tnerf_handle);
 *   - `stream` is a cudaStream_t passed as void*; all calls are asynchronous on it, none synchronises;
 *   - return value: 0 ok, <0 invalid argument / unsupported shape, >0 a cudaError_t;
 *     tnerf_last_error() gives the message of the calling thread's last failure;
 *   - float tensors are fp32, densely packed row-major unless a stride argument says otherwise;
 *   - no call has a CPU fallback.
 */
#ifndef TNERF_H_
#define TNERF_H_

#ifdef __cplusplus
extern "C" {
#endif

#define TNERF_ABI_VERSION 2

typedef struct tnerf_handle tnerf_handle;

/* precision of the fused MLP path */
enum { TNERF_PREC_F16_TC = 0,   /* fp16 operands, fp32 accumulate, tcgen05 tensor cores     */
       TNERF_PREC_F32_SIMT = 1  /* fp32 FFMA everywhere (exact mode / on-device fp32 oracle) */ };

int         tnerf_abi_version(void);
const char* tnerf_last_error(void);
/* number of kernels this library has launched in the calling process (bench.py's gpu_launches) */
long long   tnerf_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * a1  rays.get_rays                                  (src/rays.py:3-33)
 * Rays [first_ray, first_ray+n_rays) of an HxW pinhole camera, ray k = row*W + col.
 * c2w: 16 floats row-major (device).  rays_o/rays_d: (n_rays,3).  rays_o may be NULL (the
 * reference returns a broadcast view of c2w[:3,3]; the Python side builds the same view).
 */
int tnerf_get_rays(int H, int W, float focal, const float* c2w, long long first_ray, long long n_rays,
                   float* rays_o, float* rays_d, void* stream);

/* a2  ray/pixel gather of the training loop          (src/train.py:108-112)
 * out[i,:] = src[index[i],:] for three (n_src,3) tables at once (any of the src/dst pairs may be NULL). */
int tnerf_gather3(const long long* index, long long n, long long n_src,
                  const float* src_a, float* dst_a, const float* src_b, float* dst_b,
                  const float* src_c, float* dst_c, void* stream);

/* ------------------------------------------------------------------------------------------
 * a3  sampling.stratified_samples                    (src/sampling.py:3-28)
 * rays_o has row stride o_stride floats (0 = one origin broadcast to all rays, 3 = dense).
 * near/far scalars are used unless near_ray/far_ray (n_rays) are non-NULL.
 * jitter (n_rays,n_samples) uniform [0,1) or NULL for the non-randomised path.
 * z_vals (n_rays,n_samples), pts (n_rays,n_samples,3); either output may be NULL.
 */
int tnerf_stratified(const float* rays_o, long long o_stride, const float* rays_d, long long n_rays,
                     int n_samples, float near_, float far_, const float* near_ray, const float* far_ray,
                     const float* jitter, float* z_vals, float* pts, void* stream);

/* ------------------------------------------------------------------------------------------
 * a4  encoding.PositionalEncoding.forward            (src/encoding.py:21-33)
 * x (n_pts,3) -> out (n_pts, 6*num_freqs + 3*include_input), column 3+6k+3*{sin,cos}+axis.
 * The backward accumulates d(out)/d(x) for callers that differentiate w.r.t. positions.
 */
int tnerf_posenc(const float* x, long long n_pts, int num_freqs, int include_input, float* out, void* stream);
int tnerf_posenc_bwd(const float* x, const float* g_out, long long n_pts, int num_freqs, int include_input,
                     float* g_x, void* stream);

/* ------------------------------------------------------------------------------------------
 * a5  nerf.TinyNeRF                                  (src/nerf.py:4-41)
 * A handle describes one MLP: `depth` Linear(hidden)+ReLU layers, the input re-appended after
 * layer skip_at-1 (behind h), heads sigma=ReLU(Linear(hidden,1)), rgb=Sigmoid(Linear(hidden,3)).
 * Parameters are bound as 2*depth+4 fp32 device pointers in state_dict order:
 *   layers.0.weight, layers.0.bias, ..., sigma.0.weight, sigma.0.bias, rgb.0.weight, rgb.0.bias
 * and stay owned by the caller (torch.optim mutates them in place).
 */
int  tnerf_create(tnerf_handle** out, int device, int in_dim, int hidden, int depth, int skip_at);
void tnerf_destroy(tnerf_handle* h);
int  tnerf_bind_params(tnerf_handle* h, const float* const* params_host, int n_params);
long long tnerf_param_count(const tnerf_handle* h);
/* The fused entry points generate the Fourier features themselves.  By default the encoding is
 * inferred from in_dim (6L+3 -> include_input); call this to state it explicitly. */
int  tnerf_set_encoding(tnerf_handle* h, int num_freqs, int include_input);
/* Schedule options of the fused training kernel (value -1 = built-in choice; the environment variables TNERF_TRAIN_SYNC,
 * TNERF_BULK_REDUCE and TNERF_TRAIN_UNROLL_FROM give the defaults when the handle is created):
 *   "train_sync"  1 = the two tile streams of a CTA run in phase (fastest; fp32 gradient sums differ in the last bits between
 *                 runs), 0 = half a tile apart, fixed accumulation order: run-to-run REPRODUCIBLE gradients;
 *   "bulk_reduce" 1 = CTAs add their gradients into one vector with bulk async reductions, 0 = per-CTA slabs summed in order;
 *   "unroll_from" tiles per stream from which the unrolled tile program is used (tuning / tests). */
int  tnerf_set_option(tnerf_handle* h, const char* name, int value);
/* the value in EFFECT for an option ("train_sync", "bulk_reduce": 0 / 1 after defaults; "unroll_from": as set); -1 = unknown name */
int  tnerf_get_option(const tnerf_handle* h, const char* name);
/* The sum vector of the fused training kernel (hidden 128, depth 4 models): tnerf_sum_elems = its length in floats (-1: no such kernel for
 * this model); tnerf_set_sum_buffer = let tnerf_train_fwd_bwd (grads = NULL) accumulate into a CALLER-owned vector of that length
 * (16-byte aligned, zero on entry, e.g. peer-mapped memory of a multi-rank step: tnerf_allreduce_adam_step with bit 1 of `repack`
 * then sums the ranks' vectors [sum | loss | overflow flag] directly); NULL = the handle's own vector again. */
long long tnerf_sum_elems(tnerf_handle* h);
/* drop a gradient sum left in the handle's own vector by tnerf_train_fwd_bwd(grads = NULL) without applying it */
int  tnerf_clear_sum(tnerf_handle* h, void* stream);
int  tnerf_set_sum_buffer(tnerf_handle* h, float* buf);
/* Tile dealing of the fused training kernel.  The kernel deals its 64-sample tiles round-robin over (CTA, stream) pairs, so when the
 * tile count is not a multiple of 2 x CTAs the pairs with the highest indices run one tile less.  SMs of one GPU do not run at exactly
 * the same speed (a few per cent, stable per device); `order_dev` (caller-owned DEVICE array of n int32, a permutation of 0..n-1, kept
 * alive by the caller) gives CTA b the dealing index order_dev[b]: a caller that has timed the CTAs (tnerf_set_debug_buffer: stamps
 * 1024 + 4 b = start, 1025 + 4 b = end of the tile loop of CTA b, globaltimer ns) hands the shorter allotments to the slowest SMs.
 * Used only when n equals the launch's CTA count and the schedule is not the reproducible one (train_sync = 0); any permutation
 * gives the same gradient up to fp32 summation order.  n = 0 restores the identity.  The table lives in device memory and is NOT
 * validated: entries that are not a permutation of 0..n-1 make CTAs process the same tiles twice and others never (a wrong gradient,
 * no fault) -- build it on the host from a sort, as engine.Trainer.calibrate_tile_order does. */
int  tnerf_set_tile_order(tnerf_handle* h, const int* order_dev, int n);
/* Developer hook: a device buffer of 2048 int64 that receives clock64() phase stamps of CTA 0 of the fused
 * forward kernel (tools/trace_fwd.py); NULL disables it. */
int  tnerf_set_debug_buffer(tnerf_handle* h, void* buf);
/* 1 when the tcgen05 fused kernels support this handle's (in_dim, hidden, depth, skip_at) */
int  tnerf_fused_supported(const tnerf_handle* h);
/* Re-derive the packed fp16 operand image used by the tensor-core kernels from the bound fp32
 * parameters.  Call after every optimiser step (cheap: one small kernel). */
int  tnerf_pack_weights(tnerf_handle* h, void* stream);

/* Unfused fp32 forward: x (n,in_dim) -> rgb (n,3), sigma (n,1).  If `acts` is non-NULL it receives
 * the post-ReLU hidden activations (depth, n, hidden) needed by tnerf_mlp_bwd. */
int tnerf_mlp_fwd(tnerf_handle* h, const float* x, long long n, float* rgb, float* sigma, float* acts,
                  void* stream);
long long tnerf_mlp_bwd_scratch_floats(const tnerf_handle* h, long long n);
/* Unfused fp32 backward.  grads: flat (param_count) in state_dict order, ACCUMULATED into (caller
 * zeroes).  g_x (n,in_dim) may be NULL.  scratch: tnerf_mlp_bwd_scratch_floats(h,n) floats. */
int tnerf_mlp_bwd(tnerf_handle* h, const float* x, long long n, const float* acts, const float* rgb,
                  const float* sigma, const float* g_rgb, const float* g_sigma, float* grads, float* g_x,
                  float* scratch, void* stream);

/* ------------------------------------------------------------------------------------------
 * a6  volume.volume_render                           (src/volume.py:3-44)
 * rgb (n,S,3), sigma (n,S), z_vals (n,S; row stride z_stride floats, 0 = one row shared by all
 * rays), rays_d (n,3) -> comp_rgb (n,3), depth (n), acc (n), weights (n,S) or NULL.
 */
int tnerf_composite_fwd(const float* rgb, const float* sigma, const float* z_vals, long long z_stride,
                        const float* rays_d, long long n_rays, int n_samples, int white_bkgd,
                        float* comp_rgb, float* depth, float* acc, float* weights, void* stream);
/* Backward w.r.t. rgb and sigma from upstream g_comp (n,3), g_depth (n), g_acc (n), g_weights (n,S);
 * any upstream pointer may be NULL (= zero). */
int tnerf_composite_bwd(const float* rgb, const float* sigma, const float* z_vals, long long z_stride,
                        const float* rays_d, long long n_rays, int n_samples, int white_bkgd,
                        const float* g_comp, const float* g_depth, const float* g_acc, const float* g_weights,
                        float* g_rgb, float* g_sigma, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused hot path: a3 -> a4 -> a5 -> a6 in one kernel   (src/train.py:51-56 and :114-121)
 * Per-sample points, encodings and activations stay on chip.
 * ray source: (rays_o,o_stride,rays_d) as in tnerf_stratified, OR rays_d == NULL and
 * (c2w, H, W, focal, pixel_index|first_ray) to generate rays in-kernel (a1 fused in as well).
 */
typedef struct tnerf_ray_source {
    const float* rays_o;      /* may be NULL when c2w is given */
    long long    o_stride;
    const float* rays_d;      /* NULL -> generate from the camera below */
    const float* c2w;         /* 16 floats */
    int          H, W;
    float        focal;
    const long long* pixel_index; /* optional (n_rays) pixel ids k=row*W+col; NULL -> first_ray+i */
    long long    first_ray;
    /* stratified jitter of tnerf_train_fwd_bwd drawn IN-KERNEL (src/sampling.py:24 draws it on the device too): used when the
     * call's jitter tensor is NULL and jitter_seed != 0.  u(seed, step, ray, sample) = Philox4x32-10 keyed by (seed, step), counter
     * (sample, ray), 24-bit uniform in [0,1) -- the same numbers tnerf_jitter_fill writes, whatever the launch geometry.  Saves
     * 4*n_samples bytes per ray of HBM reads (86 % of the step's input) and of host-to-device traffic.  0 = off. */
    unsigned long long jitter_seed, jitter_step;
} tnerf_ray_source;

int tnerf_render_fwd(tnerf_handle* h, const tnerf_ray_source* rays_host, long long n_rays,
                     float near_, float far_, int n_samples, const float* jitter, int white_bkgd,
                     int precision, float* comp_rgb, float* depth, float* acc, float* weights,
                     float* rays_d_out, void* stream);

/* (f) N3  pose-batched rendering: the frame loop of src/make_gif.py:22-27 (render_one per spiral pose) as ONE call.
 * poses = n_poses row-major 4x4 camera-to-world matrices (device); every pose renders the pixel range
 * [first_ray, first_ray + rays_per_pose) of an H x W frame (the whole frame: 0, H*W).  Outputs are
 * (n_poses * rays_per_pose, 3|1|1), pose-major.  One kernel launch for the whole batch on the tensor-core path
 * when n_samples % 32 == 0 (the kernel indexes the pose per ray), otherwise one launch per pose from inside the call. */
int tnerf_render_frames(tnerf_handle* h, const float* poses, int n_poses, int H, int W, float focal,
                        long long first_ray, long long rays_per_pose, float near_, float far_, int n_samples,
                        int white_bkgd, int precision, float* comp_rgb, float* depth, float* acc, void* stream);

/* Backward of tnerf_render_fwd w.r.t. the MLP parameters with activations recomputed on chip
 * (implicit backward of src/train.py:126).  Upstream grads as in tnerf_composite_bwd.
 * grads (param_count) is ACCUMULATED into.  The tensor-core path carries gradients as fp16 operands:
 * they are multiplied by a power-of-two loss scale on entry and divided on exit -- grad_scale_dev
 * (1 device float) if non-NULL, else grad_scale if > 0, else (grad_scale <= 0) a scale chosen ON THE DEVICE
 * from the largest upstream gradient (2^floor(log2(64/max|g|)), one small launch, no host synchronisation). */
int tnerf_render_bwd(tnerf_handle* h, const tnerf_ray_source* rays_host, long long n_rays,
                     float near_, float far_, int n_samples, const float* jitter, int white_bkgd,
                     int precision, const float* g_comp, const float* g_depth, const float* g_acc,
                     const float* g_weights, float grad_scale, const float* grad_scale_dev, float* grads,
                     void* stream);

/* Whole training step body: forward, MSE against target (n,3), backward   (src/train.py:114-126)
 * loss_denom: the divisor of the summed squared error (3*n_rays for one process, 3*global rays
 * under ray-sharded data parallel).  Outputs: comp_rgb (n,3) or NULL, loss_sum (1 float,
 * ACCUMULATED: sum of squared errors / loss_denom), grads (param_count) ACCUMULATED.
 * GradScaler semantics of src/train.py:81,126 (scaler.scale(loss).backward()) on the device: loss_scale_dev (1 device float, e.g.
 * tnerf_scaler.state) replaces the built-in power-of-two loss scale of the tensor-core path (NULL = built-in); found_inf (1 device
 * float or NULL) is SET TO 1 when the step overflowed -- a scaled head gradient beyond 2^10 (the fp16 operands of the backward chain
 * would saturate), or a non-finite loss / gradient.  It is never cleared here.
 * grads = NULL (tensor-core path, "bulk_reduce" in effect): the gradient is LEFT in the handle as the training kernel's one sum
 * vector (already divided by the loss scale, accumulating over calls) and the next tnerf_optimizer_step with bit 1 of `repack` set
 * gathers it from there -- the gradient scatter launch disappears from the step. */
int tnerf_train_fwd_bwd(tnerf_handle* h, const tnerf_ray_source* rays_host, const float* target,
                        long long n_rays, float near_, float far_, int n_samples, const float* jitter,
                        int white_bkgd, int precision, float loss_denom, float* comp_rgb, float* loss_sum,
                        float* grads, const float* loss_scale_dev, float* found_inf, void* stream);

/* ------------------------------------------------------------------------------------------
 * a7  loss + PSNR                                    (src/train.py:122-123, src/utils.py:14-15)
 * out[0] = mean((pred-target)^2) over n floats, out[1] = -10*log10(max(out[0],1e-10)). */
int tnerf_mse_psnr(const float* pred, const float* target, long long n, float* out2, void* stream);

/* a9  optimiser                                      (src/train.py:80,125-128)
 * torch.optim.Adam semantics on a flat parameter vector, fused with the GradScaler unscale:
 * g = grads*inv_scale; if found_inf (device int, may be NULL) is non-zero the step is skipped. */
int tnerf_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n,
                    int step, float lr, float beta1, float beta2, float eps, float inv_scale,
                    const int* found_inf, void* stream);
/* a9  torch.amp.GradScaler.step()/update() (src/train.py:81,127-128) as device-resident state consumed by the optimiser entry
 * points below -- no host synchronisation, no extra launch.  state: 16 floats on the device, 8-byte aligned:
 *   [0] loss scale            [1] clean steps since the scale last changed      [2] optimiser steps actually applied
 *   [8..15] four doubles: (beta1^t, beta2^t) for call parity 0 and 1 -- initialise both pairs to beta^t of the restored step
 *           count (1.0, 1.0 for a fresh optimiser) and [0] to the initial scale.
 * Per call: if the overflow flag is non-zero the step is SKIPPED (parameters and moments untouched, the step count does not
 * advance) and the scale is multiplied by backoff_factor; otherwise Adam is applied with the bias corrections of the DEVICE step
 * count, and after growth_interval clean steps in a row the scale is multiplied by growth_factor. */
typedef struct tnerf_scaler {
    float*       state;
    const float* found_inf;      /* this call's overflow flag (device float; tnerf_train_fwd_bwd sets it).  tnerf_allreduce_adam_step
                                    ignores it: there the flag is element n+1 of the exchanged vectors                              */
    float*       clear_next;     /* device float cleared for the NEXT call, or NULL: keep two flags and alternate them by `call`     */
    float        growth_factor, backoff_factor;   /* torch defaults: 2, 0.5 */
    int          growth_interval;                 /* torch default: 2000    */
    unsigned int call;           /* strictly increasing per optimiser call (skipped or not); its parity selects the beta-power pair */
} tnerf_scaler;

/* a9 + (f) N1: the optimiser step of the training loop as ONE launch: Adam as above (inv_scale = 1) on the flat parameter
 * vector, the gradient vector cleared for the next step (grads[0 .. n_clear), n_clear >= n so a trailing loss slot is cleared
 * too; their old values go to tail_out if non-NULL) and -- bit 0 of repack -- the fp16 operand image of the tensor-core kernels refreshed in place (replaces memset +
 * tnerf_adam_step + tnerf_pack_weights; needs the handle's parameters bound as views of `params` in state_dict order).
 * Bit 1 of repack: grads[0 .. n) is NOT read -- the gradient is gathered from the sum vector a preceding tnerf_train_fwd_bwd with
 * grads = NULL left in the handle (and cleared there); grads[n .. n_clear), e.g. the loss slot, is read and cleared as usual.
 * scaler_host (HOST struct, may be NULL): GradScaler semantics as described at tnerf_scaler; `step` is then only used when
 * scaler_host is NULL. */
int tnerf_optimizer_step(tnerf_handle* h, float* params, float* grads, float* exp_avg, float* exp_avg_sq, long long n,
                         long long n_clear, int step, float lr, float beta1, float beta2, float eps, float* tail_out, int repack,
                         const tnerf_scaler* scaler_host, void* stream);

/* (e) multi-GPU exchange step (new work defined by BASELINE config 3; the reference is single-process):
 * one-shot all-reduce(sum) of the ranks' flat [gradient(n) | loss(1)] vectors over NVLink peer memory fused with the
 * optimiser step above -- what DDP would do with ncclAllReduce + optimizer.step() (src/train.py:126-127 per rank).
 * peer_grads[r] / peer_flags[r]: HOST arrays of `world` DEVICE pointers, peer-mapped into this process (e.g. CUDA IPC or
 * torch symmetric memory): rank r's vector for this epoch (16-byte aligned, padded to a multiple of 4 floats) and rank r's flag
 * array (>= world uint32, zeroed once).  epoch: same strictly increasing value (>= 1) on every rank for the same step; vectors
 * must be double-buffered by epoch parity.  The sum is formed in rank order, so every rank computes bit-identical parameters.
 * With scaler_host the vectors are [gradient(n) | loss | overflow flag]: the flags are summed too, so every rank takes the same
 * skip / apply decision.  reduced_out (vector length or NULL) receives the reduced vector; zero_next (same length or NULL) = this
 * rank's OTHER-parity vector, cleared here for the next step; repack bit 0 as in tnerf_optimizer_step (h may be NULL when 0); bit 1:
 * peer_grads[r] are the ranks' SUM vectors [sum(tnerf_sum_elems) | loss | overflow flag] left by tnerf_train_fwd_bwd (grads = NULL,
 * tnerf_set_sum_buffer) -- no gradient-scatter launch; reduced_out still receives [gradient(n) | loss | flag] in parameter order.  A peer
 * that has not arrived after TNERF_PEER_TIMEOUT_S seconds (environment, default 60, 0 = wait forever) traps instead of hanging
 * the device. */
int tnerf_allreduce_adam_step(tnerf_handle* h, float* params, float* exp_avg, float* exp_avg_sq, long long n,
                              const float* const* peer_grads, unsigned int* const* peer_flags, int world, int rank,
                              unsigned int epoch, int step, float lr, float beta1, float beta2, float eps,
                              float* reduced_out, float* zero_next, int repack, const tnerf_scaler* scaler_host, void* stream);
/* the (n_rays, n_samples) jitter tensor that tnerf_train_fwd_bwd draws in-kernel for (jitter_seed, jitter_step): parity runs feed it
 * to the oracle / to the explicit-tensor path */
int tnerf_jitter_fill(unsigned long long seed, unsigned long long step, long long n_rays, int n_samples, float* out, void* stream);
/* sets *found_inf (device int) to 1 if any grad is non-finite (device-side GradScaler check) */
int tnerf_check_finite(const float* grads, long long n, int* found_inf, void* stream);

/* test hook: size in bytes of the packed fp16 operand image (negative on error); if dst is non-NULL the image is also copied
 * there (device to device) -- used to check the optimiser's in-place refresh against a full tnerf_pack_weights */
long long tnerf_packed_image_copy(const tnerf_handle* h, void* dst, long long dst_bytes, void* stream);
/* test hook: D(128xN) = A(128xK) * B(NxK)^T through the same tcgen05 descriptors the fused
 * kernels use.  mode 0: A from shared memory, 1: A from tensor memory, 2: B MN-major. */
/* developer probe: cycles for `reps` back-to-back 128 x n x 16 MMAs (out2[0] = to completion, out2[1] = issue only) */
int tnerf_umma_rate(int n, int reps, int variant, long long* out2, void* stream);
int tnerf_umma_selftest(const float* a, const float* b, int n, int k, int mode, float* d, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TNERF_H_ */
