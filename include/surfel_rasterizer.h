/*
 * surfel_rasterizer.h — C ABI of the B200-native differentiable 2D-surfel rasterizer
 * (libsurfel_b200.so, built from 2d-gaussian-splatting_b200/csrc by nvcc for sm_100a).
 *
 * This is the drop-in boundary for the native module of hbb1/diff-surfel-rasterization
 * (pinned by /root/reference/.SUBMODULES.json:10-14; its C++/CUDA sources are NOT vendored in
 * /root/reference, so the citations below are to the reference's own call sites and to
 * SURVEY.md §8(b), which records the upstream pybind11 signatures being replaced):
 *
 *   upstream _C.rasterize_gaussians(bg, means3D, colors, opacity, scales, rotations,
 *       scale_modifier, transMat_precomp, viewmatrix, projmatrix, tan_fovx, tan_fovy, H, W, sh,
 *       degree, campos, prefiltered, debug) -> (num_rendered, color, others, radii, geomBuffer,
 *       binningBuffer, imgBuffer)
 *     == surfel_forward_preprocess()  [preprocess + tile-count scan -> num_rendered]
 *      + surfel_forward_render()      [duplicateWithKeys, radix sort, tile ranges, blend]
 *     reference call site: /root/reference/gaussian_renderer/__init__.py:37-53, :97-106
 *   upstream _C.rasterize_gaussians_backward(...) -> 8 gradient tensors
 *     == surfel_backward()
 *     reference consumers: /root/reference/train.py:90, :127-128,
 *                          /root/reference/scene/gaussian_model.py:405-407
 *   upstream _C.mark_visible(means3D, viewmatrix, projmatrix) -> bool tensor
 *     == surfel_mark_visible()
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - all tensors are contiguous float32 unless stated; "absent" optional inputs are NULL
 *     (upstream passes empty tensors);
 *   - ownership: the caller (PyTorch on the Python side) owns every buffer, including the three
 *     opaque workspaces (geometry / binning / image state) whose sizes the *_bytes() functions
 *     return; forward fills them, backward reads them, nothing is retained across calls;
 *   - every launch is ordered on the cudaStream_t passed as `stream` (void*); no call synchronises
 *     the device; nothing about a CALL is retained.  Process-global state, all of it optional tooling:
 *     the last-error string (thread-local), the binning-variant switch (surfel_set_variant /
 *     SURFEL_SORT), the launch counter and the per-stage profiling switch (surfel_profile_*), and
 *     the per-device "function attribute set" flags of kernels that opt into large shared memory;
 *   - return value: 0 on success, non-zero on failure with surfel_last_error() describing it
 *     (the Python wrapper raises RuntimeError, like upstream's AT_ERROR path).
 */
#ifndef SURFEL_RASTERIZER_H_
#define SURFEL_RASTERIZER_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SURFEL_ABI_VERSION 3
#define SURFEL_MAX_OUT_REPLICAS 8

/* Mirrors GaussianRasterizationSettings (fields constructed at
 * /root/reference/gaussian_renderer/__init__.py:37-51) plus the tile-row band used by the
 * multi-GPU tile-band partition.  tile_row_begin == tile_row_end == 0 => full frame. */
typedef struct surfel_settings {
    int32_t image_height;
    int32_t image_width;
    float tanfovx;
    float tanfovy;
    float scale_modifier;
    int32_t sh_degree;
    int32_t prefiltered;
    int32_t debug;
    int32_t tile_row_begin;
    int32_t tile_row_end;
    const float* bg;          /* (3)    device */
    const float* viewmatrix;  /* (4,4)  device, row-vector convention (scene/cameras.py:56) */
    const float* projmatrix;  /* (4,4)  device, viewmatrix @ P^T      (scene/cameras.py:57-58) */
    const float* campos;      /* (3)    device                        (scene/cameras.py:59) */
    /* Distance, in floats, between consecutive planes of out_color / out_others (forward) and of
     * dL_dout_color / dL_dout_others (backward); 0 = image_height * image_width (contiguous (C,H,W), what
     * upstream allocates).  A larger stride lets the tile-band mode render straight into a frame padded to
     * equal bands, which an in-place all-gather then completes — no staging or stitch copies. */
    int64_t out_plane_stride;
    int64_t grad_plane_stride;
    /* Forward outputs written to REPLICATED frames (multi-GPU tile-band mode): when
     * out_replica_count > 0 the render kernel stores every output value of its band to each address
     * out_replica_base[r] + 4 * (plane * out_plane_stride + y * W + x), plane = 0..2 for out_color and 3..9 for
     * out_others, INSTEAD of to out_color / out_others (which must still be valid pointers laid out the same way:
     * out_others == out_color + 3 * out_plane_stride).  The addresses are device pointers valid in this process:
     * peer mappings of the other GPUs' frames and this GPU's own frame (symmetric memory over NVLink), or ONE
     * NVSwitch multicast address that fans a single store out to all of them.  The exchange of the band outputs
     * is thereby done by the stores of the kernel that produces them; the caller only has to run a cross-GPU
     * barrier before any rank reads rows outside its own band.  0 = off (default). */
    int32_t out_replica_count;
    /* Backward, SH inputs only.  1 = surfel_backward() does NOT write dL_dsh; instead dL_dcolors (P,3, required)
     * receives the gradient of the splat's SH colour with the forward's clamp mask applied — the 3 numbers the
     * (P,M,3) SH gradient is a rank-1 expansion of.  A multi-GPU caller sums those 3 floats per splat across
     * ranks (16 floats per splat in all instead of 61) and then calls surfel_sh_grad_expand() once.  0 = off. */
    int32_t sh_grad_deferred;
    uint64_t out_replica_base[SURFEL_MAX_OUT_REPLICAS];
} surfel_settings_t;

int surfel_abi_version(void);
const char* surfel_last_error(void);

/* Selects between the two binning implementations (both sm_100a, bit-identical output):
 * "sort" = "bucket" (default: tile buckets + per-tile sort) | "radix" (device-wide CUB-free onesweep).
 * Environment default: SURFEL_SORT.  Process-global; do not change it between a forward and its backward. */
int surfel_set_variant(const char* name, const char* value);
/* 1 iff R passed to surfel_forward_render / surfel_backward may be an upper bound ("capacity") of the
 * true instance count, which lets the caller launch stage 2 before it has read R back. */
int surfel_accepts_capacity(void);

/* Workspace sizes (bytes).  R = number of (splat, tile) instances ("num_rendered"). */
size_t surfel_geom_bytes(int P);
size_t surfel_image_bytes(int W, int H);
size_t surfel_binning_bytes(size_t R, int W, int H);

/* Byte offsets of the sub-arrays inside the workspaces, for tests and debugging.
 *   geom   : out[0]=render records (P x 128 B: adjugate of T about the splat's screen position, opacity,
 *            normal, rgb, det T, Tw, culling boxes), [1]=tiles_touched u32, [2]=offsets u32 (inclusive),
 *            [3]=clamped u8 (bit c = channel c clamped), [4]=counters u32 ([1] = R),
 *            [5]=transform records (P x 48 B: transMat[9], xy[2], view depth)
 *   binning: out[0]=keys_unsorted u64, [1]=vals_unsorted u32, [2]=keys_sorted u64,
 *            [3]=vals_sorted u32 (the per-tile point list), [4]=ranges uint2 per tile
 *            (unsorted and sorted regions coincide when the sort runs an even number of passes)
 *   image  : out[0]=accum f32 (final_T, M1, M2 planes), [1]=n_contrib u32 (last, median planes) */
int surfel_geom_offsets(int P, size_t* out6);
int surfel_binning_offsets(size_t R, int W, int H, size_t* out5);
int surfel_image_offsets(int W, int H, size_t* out2);

/* Forward, stage 1: preprocess every splat and scan tiles_touched.  Writes radii (P) int32 and the
 * geometry workspace.  If image_ws != NULL the per-tile instance counts are accumulated there in the
 * same launch (fused count for the tile-bucketed binning; pass tile_counts_ready = 1 downstream).  The instance count R is left in the workspace and, if
 * num_rendered_host != NULL, delivered there in stream order: for pinned, device-mapped host memory
 * (cudaHostAlloc / torch pin_memory) the kernel stores it directly (no copy-engine transfer that
 * could queue behind a bulk download on another stream); otherwise by a 4-byte cudaMemcpyAsync.  The
 * caller synchronises the stream (or an event) before reading it to size the binning workspace. */
int surfel_forward_preprocess(const surfel_settings_t* s, int P, int M, const float* means3D,
                              const float* opacities, const float* scales, const float* rotations,
                              const float* transMat_precomp, const float* shs,
                              const float* colors_precomp, int32_t* radii, void* geom_ws, void* image_ws,
                              uint32_t* num_rendered_host, void* stream);

/* Forward, stage 2: emit keys, sort, find tile ranges, blend.  out_color (3,H,W), out_others
 * (7,H,W): 0 = sum w*depth, 1 = alpha, 2-4 = view-space normal, 5 = median depth, 6 = distortion
 * (channel order consumed at /root/reference/gaussian_renderer/__init__.py:118-135). */
int surfel_forward_render(const surfel_settings_t* s, int P, uint32_t R, const int32_t* radii,
                          const void* geom_ws, void* binning_ws, void* image_ws, int tile_counts_ready,
                          float* out_color, float* out_others, void* stream);

/* The two halves of stage 2, exposed separately for parity tests. */
int surfel_bin_duplicate(const surfel_settings_t* s, int P, uint32_t R, const void* geom_ws,
                         const int32_t* radii, void* binning_ws, void* stream);
int surfel_bin_sort(const surfel_settings_t* s, uint32_t R, void* binning_ws, void* stream);
/* Production binning used by surfel_forward_render: counting scatter of (depth|idx) pairs into tile
 * buckets + per-tile shared-memory sort; fills ranges and the point list (vals_sorted), and the
 * sorted keys too when write_keys != 0.  Result is identical to surfel_bin_duplicate + surfel_bin_sort
 * (the CUB-free device-wide radix sort), which stays selectable with SURFEL_SORT=radix. */
int surfel_bin_bucket(const surfel_settings_t* s, int P, uint32_t R, const void* geom_ws,
                      const int32_t* radii, void* binning_ws, const void* image_ws_with_counts,
                      int write_keys, void* stream);
int surfel_render_forward(const surfel_settings_t* s, uint32_t R, const void* geom_ws,
                          const void* binning_ws, void* image_ws, float* out_color,
                          float* out_others, void* stream);

/* Backward.  dL_dout_color (3,H,W), dL_dout_others (7,H,W).  grad_scratch: P *
 * surfel_grad_scratch_floats() floats (zeroed here).  Outputs are written for every splat (zeros where culled), so they may be uninitialised:
 *   dL_dmeans2D (P,3) [densification proxy in .xy, SURVEY A.5], dL_dcolors (P,3),
 *   dL_dopacity (P,1), dL_dmeans3D (P,3), dL_dtransMat (P,9), dL_dsh (P,M,3),
 *   dL_dscales (P,2), dL_drotations (P,4).  Optional outputs may be NULL when the matching input
 *   is absent.  lowpass_depth_quirk: 1 = the published upstream kernel's depth gradient in the low-pass
 *   branch, dL_dTw += (s.x, s.y, 1) * dL_dz (what callers should pass: the reference trains on it);
 *   0 = the exact derivative of the forward there, (0, 0, 1) * dL_dz.  See DESIGN.md. */
int surfel_grad_scratch_floats(void);
int surfel_backward(const surfel_settings_t* s, int P, int M, uint32_t R, const float* means3D,
                    const float* scales, const float* rotations, const float* transMat_precomp,
                    const float* shs, int has_colors_precomp, const int32_t* radii,
                    const void* geom_ws, const void* binning_ws, const void* image_ws,
                    const float* dL_dout_color, const float* dL_dout_others, float* grad_scratch,
                    float* dL_dmeans2D, float* dL_dcolors, float* dL_dopacity, float* dL_dmeans3D,
                    float* dL_dtransMat, float* dL_dsh, float* dL_dscales, float* dL_drotations,
                    int lowpass_depth_quirk, void* stream);

/* dL_dsh (P,M,3) = basis_k(normalize(means3D - campos)) * dL_dcolors[c] for k < (sh_degree+1)^2, zero beyond:
 * the expansion surfel_backward() skips when surfel_settings.sh_grad_deferred = 1.  Rows of splats whose
 * colour gradient is exactly zero are zero. */
int surfel_sh_grad_expand(int P, int M, int sh_degree, const float* means3D, const float* campos,
                          const float* dL_dcolors, float* dL_dsh, void* stream);

/* GaussianRasterizer.markVisible: near-plane test (present: P bytes, 0/1). */
int surfel_mark_visible(int P, const float* means3D, const float* viewmatrix,
                        const float* projmatrix, uint8_t* present, void* stream);

/* Stand-alone CUB-free stable radix sort of (u64 key, u32 value) pairs on key bits [0,end_bit).
 * Data starts in A; *result_in_b tells where the sorted pairs are (buffers ping-pong per pass). */
size_t surfel_sort_temp_bytes(size_t n);
int surfel_sort_pairs(uint64_t* keys_a, uint32_t* vals_a, uint64_t* keys_b, uint32_t* vals_b,
                      size_t n, int end_bit, void* temp, int* result_in_b, void* stream);

/* OPT-IN fused post-process of the op's allmap: what the reference's render() does
 * with ~10 PyTorch kernels per direction at /root/reference/gaussian_renderer/__init__.py:118-147 and
 * /root/reference/utils/point_utils.py:9-37.  rot (9): n_world = n_view . rot (= world_view[:3,:3]^T);
 * rays (12): 3x3 pixel->world ray matrix (row-major, dir = (x,y,1).M) followed by the camera centre.
 * Outputs: rend_normal (3,H,W), surf_depth (1,H,W), surf_normal (3,H,W).  Backward: cotangents of the
 * three outputs (any may be NULL), tmp6 = (6,H,W) scratch, g_allmap (7,H,W) fully written. */
int surfel_post_forward(int W, int H, float depth_ratio, const float* allmap, const float* rot,
                        const float* rays, float* rend_normal, float* surf_depth, float* surf_normal,
                        void* stream);
int surfel_post_backward(int W, int H, float depth_ratio, const float* allmap, const float* rot,
                         const float* rays, const float* surf_depth, const float* g_rend_normal,
                         const float* g_surf_depth, const float* g_surf_normal, float* tmp6,
                         float* g_allmap, void* stream);

/* OPT-IN fused photometric loss: (1-l)*L1 + l*(1-SSIM) of
 * /root/reference/train.py:73-74 with /root/reference/utils/loss_utils.py:6-7, :43-73 (11x11 Gaussian
 * window, sigma 1.5, zero padding, per channel).  forward: sums2[0] = sum |img-gt|, sums2[1] = sum of
 * the SSIM map (doubles, device) and the three (C,H,W) derivative maps the backward consumes.
 * backward: gscale2 (device, 2 floats) = dL/d(sums2); g_img (C,H,W) fully written. */
int surfel_l1_ssim_forward(int C, int H, int W, const float* img, const float* gt, float* dmu1,
                           float* ds11, float* ds12, double* sums2, void* stream);
int surfel_l1_ssim_backward(int C, int H, int W, const float* img, const float* gt, const float* dmu1,
                            const float* ds11, const float* ds12, const float* gscale2, float* g_img,
                            void* stream);

/* ---- the parameter update after the backward -------------------------------
 * surfel_adam_step replaces torch.optim.Adam(l, lr=0.0, eps=1e-15).step() of the reference
 * (/root/reference/scene/gaussian_model.py:148-166, /root/reference/train.py:138-140): ONE launch
 * updates every parameter group in place (param, exp_avg, exp_avg_sq), arithmetic as torch's
 * single-tensor Adam (no weight decay, no amsgrad).  The host folds the step count into
 *   step_size = lr / (1 - beta1^step),   bias2_sqrt = sqrt(1 - beta2^step)
 * (in double, as torch does); betas and eps are doubles so that 1 - beta is rounded once.
 * surfel_densify_stats replaces /root/reference/train.py:125-128 (max_radii2D update) and
 * /root/reference/scene/gaussian_model.py:405-407 (add_densification_stats): where radii > 0,
 *   max_radii2D = max(max_radii2D, radii); xyz_gradient_accum += |means2D_grad (3)|; denom += 1.
 * max_radii2D may be NULL. */
#define SURFEL_ADAM_MAX_GROUPS 8
typedef struct surfel_adam_group {
    float* param;            /* n floats, updated in place */
    const float* grad;       /* n floats */
    float* exp_avg;          /* n floats, updated in place */
    float* exp_avg_sq;       /* n floats, updated in place */
    long long n;
    float step_size;
    float bias2_sqrt;
    int aligned16;           /* filled in by the library */
} surfel_adam_group_t;
int surfel_adam_step(int n_groups, const surfel_adam_group_t* groups, double beta1, double beta2, double eps,
                     void* stream);
int surfel_densify_stats(int P, const int32_t* radii, const float* means2D_grad, float* xyz_gradient_accum,
                         float* denom, float* max_radii2D, void* stream);

/* ---- the model's on-disk format either side of the path ------------------
 * The reference saves / loads a trained model as a binary little-endian PLY with one row of 61
 * float32 per splat, pre-activation, SH channel-major (/root/reference/scene/gaussian_model.py:176-209
 * save_ply, :215-255 load_ply):  x y z nx ny nz f_dc_0..2 f_rest_0..44 opacity scale_0..1 rot_0..3.
 * surfel_ply_unpack turns raw rows (device memory, `row_floats` floats each, any property order)
 * into the rasterizer's inputs in one pass.  columns[58] gives, for every target float, its column
 * in the row; target order: 0..2 xyz | 3..50 shs[k][c] (coefficient-major (16,3): k = 0 is f_dc_c,
 * k >= 1 is f_rest_{c*15 + k-1}) | 51 opacity | 52..53 scale | 54..57 rot (w,x,y,z).
 * activate = 1 applies what the reference's getters apply before the op (gaussian_model.py:35-41,
 * :95-115): opacity = sigmoid, scale = exp, rot = q / max(|q|, 1e-12); activate = 0 leaves the
 * stored parameters.  surfel_ply_pack is the inverse of save_ply's gather: parameters
 * (xyz (P,3), features_dc (P,1,3), features_rest (P,15,3), opacity (P,1), scaling (P,2),
 * rotation (P,4)) -> rows in the reference's column order with zero normals. */
#define SURFEL_PLY_ROW_FLOATS 61
#define SURFEL_PLY_TARGETS 58
#define SURFEL_PLY_MAX_ROW_FLOATS 127
int surfel_ply_unpack(int P, int row_floats, const float* rows, const int32_t* columns, int activate,
                      float* means3D, float* shs, float* opacities, float* scales, float* rotations, void* stream);
int surfel_ply_pack(int P, const float* xyz, const float* features_dc, const float* features_rest,
                    const float* opacity, const float* scaling, const float* rotation, float* rows, void* stream);

/* Instrumentation used by bench.py: number of kernels this library has launched in this process,
 * and optional per-stage CUDA-event timing (events recorded on the launching stream around each
 * kernel while enabled; surfel_profile_read() waits for them and returns summed ms / launch counts
 * per stage since the previous read). */
unsigned long long surfel_launch_count(void);
void surfel_profile_enable(int on);
int surfel_profile_num_stages(void);
const char* surfel_profile_stage_name(int stage);
int surfel_profile_read(double* ms_out, int* count_out);

#ifdef __cplusplus
}
#endif
#endif /* SURFEL_RASTERIZER_H_ */
