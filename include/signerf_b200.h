/* signerf_b200 — C ABI of the B200-native SIGNeRF reference-sheet hot path.
 *
 * The reference (cgtuebingen/SIGNeRF) has NO FFI: its hot path is Python calling nerfstudio
 * (SURVEY.md §8b).  Every entry point below therefore cites the reference *Python* interface
 * it replaces; the binding a maintainer adds on the reference side is the ctypes stub shown
 * in INTEGRATION.md (mirrored by signerf_b200/_lib.py).
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross the boundary;
 *   - every `d_` pointer is DEVICE memory owned by the caller, every `h_` pointer HOST memory;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *   - return value: SGN_OK (0) or a negative SgnStatus; sgn_last_error() gives the message of
 *     the last failure on the calling thread.  Nothing throws across the ABI;
 *   - the library allocates only inside opaque handles (sgn_*_create / sgn_*_destroy);
 *   - re-entrant per handle; safe to call from a non-main Python thread.
 */
#ifndef SIGNERF_B200_H
#define SIGNERF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum SgnStatus {
  SGN_OK = 0,
  SGN_ERR_INVALID_ARG = -1,
  SGN_ERR_UNSUPPORTED = -2, /* architecture/shape the sm_100a kernels are not specialised for */
  SGN_ERR_CUDA = -3,
  SGN_ERR_NO_DEVICE = -4
} SgnStatus;

const char* sgn_last_error(void);
/* ABI version; bumped on any signature change. */
int sgn_abi_version(void);
/* Number of kernel launches issued by this library on the calling process so far
 * (bench.py reports the delta over the timed region as `gpu_launches`). */
uint64_t sgn_launch_count(void);
/* Process-wide tuning knobs (set between launches, not thread-safe against concurrent launches):
 *   "render_ctas_per_sm" 1..4 (default 4): resident renderer CTAs per SM; 1 leaves room for a co-resident GEMM CTA
 *   "gemm_pair_stages"   5|6  (default 6): pipeline stages of the CTA-pair GEMM (5 = 160 KB shared memory)
 * (Measured on B200, scratch/overlap.py: co-running a 1-CTA/SM renderer with the GEMMs on two streams is slower than
 * running them back to back - 194 ms against 145 ms - so bench.py keeps the defaults and a single stream.) */
int sgn_set_option(const char* name, int value);

/* ------------------------------------------------------------------ field (A4/A5) */
/* One multiresolution hash grid = nerfstudio HashEncoding, torch-fallback semantics
 * (always hashed, floor/ceil corners, primes {1, 2654435761, 805459861}).               */
typedef struct SgnHashGrid {
  const float* d_table;    /* [num_levels * 2^log2_size, 2] fp32, caller-owned, kept by reference */
  const float* h_scalings; /* [num_levels] per-level resolution, EXACTLY the fp32 values of
                              HashEncoding.scalings (never recomputed here)      */
  int num_levels;          /* <= 16 */
  int log2_size;           /* table rows per level = 2^log2_size */
} SgnHashGrid;

/* nn.Linear: y = W x + b, W row-major [out_dim, in_dim], host fp32. */
typedef struct SgnLinear {
  const float* h_weight;
  const float* h_bias;
  int in_dim, out_dim;
} SgnLinear;

/* nerfacto field + proposal networks (nerfstudio fields/nerfacto_field.py, density_fields.py)
 * as SIGNeRFModel(NerfactoModel) holds them (reference signerf/signerf.py:27-39).            */
typedef struct SgnFieldDesc {
  SgnHashGrid grid;           /* main: 16 levels, 2^19, F=2 */
  SgnLinear base[2];          /* 32->64->16 (ReLU between, no out activation) */
  SgnLinear head[3];          /* 63->64->64->3 (ReLU, Sigmoid); input = [SH16, geo15, app32] */
  const float* h_appearance;  /* [32] mean appearance embedding used in eval */
  float average_init_density; /* SIGNeRF: 0.01 (signerf_config.py:35) */
  int num_proposals;          /* 0 (flat mode only) or 2 */
  SgnHashGrid prop_grid[2];
  SgnLinear prop_mlp[2][2];   /* 10->16->1 each */
} SgnFieldDesc;

typedef struct SgnField SgnField;
int sgn_field_create(const SgnFieldDesc* desc, SgnField** out);
void sgn_field_destroy(SgnField* f);

/* ------------------------------------------------------------------ K1: render (A1-A6) */
#define SGN_MLP_FP16_MMA 0 /* fp16 mma.sync tensor-core MLP, fp32 accumulate (fast path)  */
#define SGN_MLP_FP32 1     /* fp32 CUDA-core MLP (parity / debugging path)                  */

typedef struct SgnRenderOpts {
  float near_plane, far_plane; /* NearFarCollider(0.05, 1000) */
  int mode;                    /* 0 = flat: S piecewise-lin-disp bins straight into the main field;
                                  1 = cascade: ProposalNetworkSampler 256 -> 96 -> 48            */
  int num_samples;             /* flat: S;  cascade: samples into the main field (48)            */
  int num_prop_samples[2];     /* cascade only: {256, 96}                                        */
  int mlp_mode;                /* SGN_MLP_* */
  const float* h_bins;         /* optional [S+1] (flat) / [num_prop_samples[0]+1] (cascade) EUCLIDEAN
                                  bin edges shared by all rays, computed by the host with the same
                                  torch expression the reference uses; NULL = computed in-library */
} SgnRenderOpts;

/* Replaces, for V views at once, reference `DatasetGenerator.render_camera`'s
 *   camera.generate_rays(camera_indices=0, ...)            datasetgenerator.py:691
 *   graph.get_outputs_for_camera_ray_bundle(bundle)        datasetgenerator.py:694
 * and returns the two outputs it consumes (:700-701) plus accumulation.
 *   d_c2w  [V,3,4] camera-to-world, d_intr [V,4] = fx,fy,cx,cy   (device)
 *   d_rgb  [V,H,W,3], d_depth [V,H,W,1], d_acc [V,H,W,1] or NULL (device)
 * Ray id inside a view is y*W + x (row-major), pixel centres at +0.5.                        */
int sgn_render_views(const SgnField* f, const float* d_c2w, const float* d_intr, int V, int H, int W,
                     const SgnRenderOpts* opts, float* d_rgb, float* d_depth, float* d_acc, void* stream);

/* The second half alone, for a caller that already holds the rays: nerfstudio's
 * `Model.get_outputs_for_camera_ray_bundle(camera_ray_bundle)` (datasetgenerator.py:694) on a RayBundle's
 * `origins` / `directions` (unit length, as `Cameras.generate_rays` returns them), flattened row-major.
 *   d_origins, d_directions [N,3] fp32; d_rgb [N,3], d_depth [N], d_acc [N] or NULL (device).
 * Same sampling / field / compositing as sgn_render_views (near / far planes from opts: NearFarCollider). */
int sgn_render_rays(const SgnField* f, const float* d_origins, const float* d_directions, int64_t N,
                    const SgnRenderOpts* opts, float* d_rgb, float* d_depth, float* d_acc, void* stream);

/* Same call with HOST buffers: H2D of the cameras and D2H of the images happen inside
 * (this is the end-to-end entry bench.py times as `e2e`).                                    */
int sgn_render_views_host(const SgnField* f, const float* h_c2w, const float* h_intr, int V, int H, int W,
                          const SgnRenderOpts* opts, float* h_rgb, float* h_depth, float* h_acc);

/* Ray generation only (A2): nerfstudio Cameras.generate_rays semantics.
 *   d_origins/d_directions [V,H,W,3]; d_pixel_area/d_dir_norm [V,H,W,1] or NULL.             */
int sgn_generate_rays(const float* d_c2w, const float* d_intr, int V, int H, int W, float* d_origins,
                      float* d_directions, float* d_pixel_area, float* d_dir_norm, void* stream);

/* Hash-grid probe (A4) for parity tests: positions already in [0,1]^3.
 *   which: 0 = main grid, 1/2 = proposal grid 0/1
 *   d_indices [N, L, 8] int64 table rows (corner order of HashEncoding.pytorch_fwd), or NULL
 *   d_features [N, 2L] fp32 trilinear features, or NULL                                       */
int sgn_hash_encode(const SgnField* f, int which, const float* d_positions01, int64_t N,
                    int64_t* d_indices, float* d_features, void* stream);

/* Field probe (A5) for parity tests: world-space positions + unit directions ->
 * d_density [N], d_rgb [N,3] with the selected MLP mode.                                      */
int sgn_field_eval(const SgnField* f, const float* d_positions, const float* d_directions, int64_t N,
                   int mlp_mode, float* d_density, float* d_rgb, void* stream);

/* ------------------------------------------------------------------ K2/K3: mask + condition (A7-A9) */
typedef struct SgnMaskOpts {
  float aabb[6];           /* min xyz, max xyz (DatasetGenerator.aabb) */
  int inverse_mask;        /* DatasetGeneratorConfig.inverse_mask */
  int dilate_w, dilate_h;  /* mask_dialation (50,50); 0,0 = none. cv2 MORPH_ELLIPSE semantics */
  float depth_radius;      /* additional_depth_radius (0.1) */
  int use_manual_depth;    /* manual_depth is not None */
  float manual_min, manual_max;
} SgnMaskOpts;

/* Replaces reference render_camera's masking_mode == "aabb" branch, datasetgenerator.py:758-818
 * (intersect_with_aabb utils/intersection.py:5-56, cv2.dilate :775-778, min/max :786-792,
 * condition :809-810, not-visible zeros :813-818) for V views without any host round trip.
 *   d_depth [V,H,W,1] in; d_mask [V,H,W,1] uint8 0/1 out; d_cond [V,H,W,1] fp32 out
 *   d_stats [V,4] fp32 out or NULL: {is_visible, min_depth, max_depth, visible_count}         */
int sgn_mask_condition(const float* d_c2w, const float* d_intr, int V, int H, int W, const float* d_depth,
                       const SgnMaskOpts* opts, uint8_t* d_mask, float* d_cond, float* d_stats,
                       void* stream);

/* The same branch with combine_shape_with_depth=True (datasetgenerator.py:794-807): mask / min / max exactly as above;
 * where the proxy mesh is in front of the NeRF surface ((shape_depth < depth) & (shape_depth > 0)) the condition takes
 * the R channel of the mesh's colour image / 255 instead of the normalised NeRF depth, then 1 - clamp(.,0,1).
 *   d_shape_depth [V,H,W,1] fp32 (0 = empty), d_shape_color [V,H,W,3] uint8 = Renderer.render_camera's two outputs
 *   (signerf/renderer/renderer.py:149-196).                                                      */
int sgn_mask_condition_combined(const float* d_c2w, const float* d_intr, int V, int H, int W, const float* d_depth,
                                const float* d_shape_depth, const uint8_t* d_shape_color, const SgnMaskOpts* opts,
                                uint8_t* d_mask, float* d_cond, float* d_stats, void* stream);

/* cv2.dilate(mask, getStructuringElement(MORPH_ELLIPSE, (kw, kh))) > 0 on V binary images.    */
int sgn_dilate_ellipse(const uint8_t* d_in, int V, int H, int W, int kw, int kh, uint8_t* d_out,
                       void* stream);

/* ------------------------------------------------------------------ K4: sheet (A10, A14) */
/* F.interpolate(bilinear, align_corners=False) of V tiles [V,H,W,C] to (h,w) and paste of tile v
 * at grid cell (v0+v) of a rows x cols sheet with `border` px between tiles
 * (datasetgenerator.py:526-539, :633-646).  threshold >= 0: write (value > threshold) as 0/1
 * (the mask path, `> 0.5`).  src_u8: source is uint8 0/1 (mask) instead of fp32.               */
int sgn_sheet_paste(const void* d_src, int src_u8, int V, int H, int W, int C, float* d_sheet,
                    int sheet_h, int sheet_w, int rows, int cols, int border, int tile_h, int tile_w,
                    int first_cell, float threshold, void* stream);

/* Inverse: cut tile `cell` out of the sheet and bilinear-resize to (H,W) (datasetgenerator.py:570-585,
 * :652-659), optionally blending `edited*mask + base*(1-mask)` first (:562, :656).             */
int sgn_sheet_cut(const float* d_sheet, int sheet_h, int sheet_w, int C, int rows, int cols, int border,
                  int tile_h, int tile_w, int cell, float* d_out, int H, int W, void* stream);

/* out = edited*mask + base*(1-mask), mask [npix] broadcast over C channels
 * (datasetgenerator.py:562 sheet blend, :656 tile blend).                                     */
int sgn_blend_masked(const float* d_edited, const float* d_base, const float* d_mask, int64_t npix, int C,
                     float* d_out, void* stream);

/* Multi-GPU exchange fused with the tile packing ("fused NVLink peer stores"): this rank's views
 * (rank, rank+world, ...) of each grid g are written as packed pixels [r,g,b,depth,cond,mask] (6 fp32) directly into
 * the tile buffer of grid g's owner, d_peer_ptrs[g] -> [v_loc*world, H, W, 6], a peer-mapped device pointer
 * (torch symmetric memory / cudaIpc).  Inputs are [G, v_loc, H, W, C] contiguous.  The caller orders it between two
 * cross-rank barriers; no reduction is involved, so results equal the NCCL all-gather bit for bit. */
int sgn_scatter_tiles_peer(const float* d_rgb, const float* d_depth, const float* d_cond, const uint8_t* d_mask, int G,
                           int v_loc, int H, int W, int world, int rank, const int64_t* d_peer_ptrs, void* stream);

/* tensor_to_image quantisation (utils/image_tensor_converter.py:22-23,29-30):
 * uint8(x*255) by truncation with wrap-around, exactly numpy's float32 -> uint8 cast.          */
int sgn_quantize_u8(const float* d_in, int64_t n, uint8_t* d_out, void* stream);

/* ------------------------------------------------------------------ K5-K9: SDXL / ControlNet UNet step (A11-A13) */
/* The reference reaches this arithmetic over HTTP: Diffuser._diffuse_remote_sdwebui_controlnet POSTs the sheet to an
 * A1111 SD-WebUI server (signerf/diffuser/diffuser.py:116-195) which runs sgm's UNetModel + the ControlNet extension.
 * The in-process replacement plugs into the hook the reference leaves open, Diffuser._custom (diffuser.py:102-113);
 * signerf_b200/unet.py walks the UNet graph and calls the operators below.
 * Activations are NHWC ("tokens x channels"): the fp32 residual stream [B*H*W, C] and fp16 GEMM operands. */

/* Epilogue of the tensor-core contractions:  y = acc + bias[n] + rowbias[m / rows_per_batch][n] + residual[m][n]. */
typedef struct SgnEpilogue {
  const float* d_bias;     /* [N] or NULL */
  const float* d_rowbias;  /* [M / rows_per_batch, N] or NULL: ResBlock's per-image emb_layers output */
  int rows_per_batch;
  const float* d_residual; /* [M, ldo] fp32 or NULL; may alias the output (in-place residual add) */
  int64_t ldo;             /* output (and residual) row stride in elements; 0 = dense */
  int out_f16;             /* 0: fp32 output, 1: fp16 output */
  int geglu;               /* 1: weight rows interleaved (value_j, gate_j): out[m][j] = value * gelu(gate), fp16 [M, N/2]
                              (sgm GEGLU: `x, gate = proj(x).chunk(2); x * F.gelu(gate)`) */
  int nchw;                /* conv only: store fp32 NCHW [B, N, H, W] (the UNet's 4-channel output conv) */
  int act_silu;            /* 1: y = silu(y) before the store (ControlNet input_hint_block) */
} SgnEpilogue;

/* nn.Linear / 1x1 conv on tcgen05: out[M,N] = A[M,K] . W[N,K]^T (+ epilogue).  A, W fp16 row-major with row strides
 * lda / ldw (elements, multiples of 8); fp32 accumulation in TMEM. */
int sgn_gemm_f16(const void* d_a, int64_t lda, const void* d_w, int64_t ldw, int M, int N, int K,
                 const SgnEpilogue* ep, void* d_out, void* stream);

/* Host-side probe of the tile schedule sgn_gemm_f16 would use (no launch, works without a GPU: 148 SMs assumed then):
 * h_plan[6] = {block_n, cluster size, tiles, tail split, work items, scheduling units}. */
int sgn_gemm_plan(int M, int N, int K, int residual_f32, int* h_plan);

/* 3x3 / stride 1 / pad 1 convolution as an implicit GEMM on tcgen05.  d_x fp16 NHWC [B,H,W,C] (C % 64 == 0),
 * d_w fp16 [N, 9*C] with k = (ky*3 + kx)*C + c  (torch weight.permute(0,2,3,1)); out [B*H*W, N]. */
int sgn_conv3x3_f16(const void* d_x, const void* d_w, int B, int H, int W, int C, int N, const SgnEpilogue* ep,
                    void* d_out, void* stream);

/* sgm CrossAttention core (head_dim 64): out = softmax(Q K^T * scale) V per (image, head), on tcgen05.
 * Q [B*T_q, ldq], K / V [B*T_kv, ldk / ldv], out [B*T_q, ldo], all fp16; head h uses columns [64h, 64h+64) of each
 * (so Q/K/V may be column slices of one fused projection output).  Any T_q, T_kv >= 1 (77-token context included). */
int sgn_attention_f16(const void* d_q, int64_t ldq, const void* d_k, int64_t ldk, const void* d_v, int64_t ldv,
                      int B, int heads, int T_q, int T_kv, float scale, void* d_out, int64_t ldo, void* stream);

/* Same, with a caller-owned workspace: (query tile, head, image) work items that would only fill part of a last wave
 * of CTAs are cut into key ranges whose partial (O, m, l) the last-arriving CTA combines (4.32 waves at T = 4 096 x 20
 * heads x 2 images on 148 SMs become 4 + 0.34).  sgn_attention_workspace_bytes = bytes this shape needs on the current
 * device (0: no split, d_ws may be NULL); the workspace is scratch, nothing survives the call. */
int64_t sgn_attention_workspace_bytes(int B, int heads, int T_q, int T_kv);
int sgn_attention_f16_ws(const void* d_q, int64_t ldq, const void* d_k, int64_t ldk, const void* d_v, int64_t ldv,
                         int B, int heads, int T_q, int T_kv, float scale, void* d_out, int64_t ldo, void* d_ws,
                         int64_t ws_bytes, void* stream);

/* GroupNorm (sgm `normalization` = GroupNorm32(32, C), eps 1e-5; SpatialTransformer.norm eps 1e-6) + optional SiLU.
 * d_x fp32 [B, HW, C] -> d_out fp16 [B, HW, C].  d_ws: scratch of sgn_group_norm_ws_doubles(B, HW, groups) doubles
 * (per-chunk partial sums, reduced in a fixed order: results are run-to-run bit-identical). */
int64_t sgn_group_norm_ws_doubles(int B, int HW, int groups);
int sgn_group_norm_f16(const float* d_x, int B, int HW, int C, int groups, float eps, const float* d_gamma,
                       const float* d_beta, int act_silu, double* d_ws, void* d_out, void* stream);
/* Same, with the result split into two fp16 halves: d_out fp16 [B, HW, 2C] = [hi | lo], y = hi + lo to 2^-22.  A
 * contraction against the weights repeated along K ([W | W], per tap for a conv) then reproduces the fp32 operator to
 * fp32 rounding for fp16-representable weights (the VAE's "exact" mode, signerf_b200/vae.py). */
int sgn_group_norm_split_f16(const float* d_x, int B, int HW, int C, int groups, float eps, const float* d_gamma,
                             const float* d_beta, int act_silu, double* d_ws, void* d_out, void* stream);
/* nn.LayerNorm(C) over the last dimension: fp32 [M, C] -> fp16 [M, C]. */
int sgn_layer_norm_f16(const float* d_x, int64_t M, int C, float eps, const float* d_gamma, const float* d_beta,
                       void* d_out, void* stream);
/* fp32 -> fp16 cast of n elements (n % 4 == 0). */
int sgn_cast_f16(const float* d_x, int64_t n, void* d_out, void* stream);
/* sgm Upsample: F.interpolate(scale_factor=2, mode="nearest"); fp32 NHWC [B,H,W,C] -> fp16 NHWC [B,2H,2W,C]. */
int sgn_upsample2x_f16(const float* d_x, int B, int H, int W, int C, void* d_out, void* stream);
/* torch.cat([a, b + scale*b2], dim=channel) on NHWC fp32: the UNet decoder's skip concat with the ControlNet residual
 * folded in (sd-webui-controlnet hook: `h = cat([h, hs.pop() + control.pop()])`).  d_b2 may be NULL. */
int sgn_concat_f32(const float* d_a, int Ca, const float* d_b, const float* d_b2, float scale, int Cb, int64_t P,
                   float* d_out, void* stream);
/* Same, also emitting the result as fp16 [P, Ca+Cb] (operand of the ResBlock's 1x1 shortcut GEMM). */
int sgn_concat_f32_f16(const float* d_a, int Ca, const float* d_b, const float* d_b2, float scale, int Cb, int64_t P,
                       float* d_out, void* d_out16, void* stream);
/* y += a * x (ControlNet middle-block residual). */
int sgn_axpy_f32(const float* d_x, float a, int64_t n, float* d_y, void* stream);
/* im2col of sgm Downsample (3x3 / stride 2 / pad 1): fp32 NHWC [B,H,W,C] -> fp16 [B*Ho*Wo, 9*C], feeding sgn_gemm_f16. */
int sgn_im2col3x3_s2_f16(const float* d_x, int B, int H, int W, int C, void* d_out, void* stream);
/* Direct fp32 3x3 / pad 1 conv for channel counts the tensor-core path does not take (UNet conv_in 4->320, ControlNet
 * input_hint_block 3->16->...->256).  d_x fp32 NHWC or NCHW (in_nchw), d_w fp32 [Cout,3,3,Cin]; out NHWC fp32 / fp16
 * = act(conv + bias + residual[b % res_batch]) (ControlNet: `h = conv_in(x) + guided_hint`, one hint per CFG pair;
 * res_batch 0 = B). */
int sgn_conv3x3_direct(const float* d_x, int in_nchw, const float* d_w, const float* d_bias, const float* d_residual,
                       int res_batch, int B, int H, int W, int Cin, int Cout, int stride, int act_silu, int out_f16,
                       void* d_out, void* stream);
/* out[b] = act_out(W act_in(x[b]) + bias (+ residual[b])) for B <= 8 rows, fp32 (time_embed, label_emb, emb_layers). */
int sgn_linear_small(const float* d_x, const float* d_w, const float* d_bias, const float* d_residual, int B, int N,
                     int K, int silu_in, int silu_out, float* d_out, void* stream);
/* The `emb_layers` projections of every ResBlock of a network in one launch ([EXT] sgm ResBlock.forward:
 * `emb_out = self.emb_layers(emb)`, which depends on emb only): d_w [N, K] / d_bias [N] are the blocks' weights concatenated
 * along the rows, d_seg_offsets int32 [num_segments + 1] the row range of every block; block j's result is the contiguous
 * [B, N_j] matrix at d_out + B * seg[j] - bit for bit what sgn_linear_small gives per block. */
int sgn_linear_small_segments(const float* d_x, const float* d_w, const float* d_bias, int B, int N, int K, int silu_in,
                              const int32_t* d_seg_offsets, int num_segments, float* d_out, void* stream);
/* sgm timestep_embedding(t, dim, max_period=10000): [cos | sin]. */
int sgn_timestep_embedding(const float* d_t, int B, int dim, float* d_out, void* stream);

/* K9 (A13): one sampler update on latents [B,C,H,W] fp32, fused: classifier-free guidance over d_eps = [cond | uncond]
 * (2*B*C*H*W), k-diffusion eps -> denoised, A1111 inpaint blend `denoised = init*mask + (1-mask)*denoised`
 * (d_mask [B,1,H,W], NULL = no blend) and the Euler-ancestral step x' = x + d*(sigma_down - sigma) + noise*sigma_up
 * (d_noise NULL = no noise). d_denoised may be NULL. */
int sgn_cfg_euler_step(const float* d_x, const float* d_eps, const float* d_init, const float* d_mask,
                       const float* d_noise, int B, int C, int H, int W, float cfg_scale, float sigma, float sigma_down,
                       float sigma_up, float* d_x_out, float* d_denoised, void* stream);

/* out[r*n + i] = scale * x[i], r < repeats: A1111 CFGDenoiser's x_in = cat([x, x]) * c_in. */
int sgn_scale_repeat_f32(const float* d_x, int64_t n, float scale, int repeats, float* d_out, void* stream);
/* Sheet -> denoiser conditioning (diffuser.py:121-130, :146-166): d_hint [3,Hs,Ws] = uint8(cond*255)/255 replicated
 * (tensor_to_image truncation, preprocessor "none"); d_lat_mask [Hs/8,Ws/8] = 1 - round(8x8 box mean of the mask). */
int sgn_sheet_to_conditioning(const float* d_cond, const float* d_mask, int Hs, int Ws, float* d_hint,
                              float* d_lat_mask, void* stream);

/* im2col of a 3x3 / pad 1 / stride 1|2 conv with the fp32 input split into two fp16 halves (x = hi + lo exactly to
 * 2^-22): d_out fp16 [B*Ho*Wo, 2*Kp], Kp = round_up(9*C, 8), row = [hi(k) for k < Kp | lo(k) for k < Kp],
 * k = (ky*3+kx)*C + c.  A GEMM against [W | W] then reproduces the fp32 convolution to fp32 rounding on the tensor
 * cores (weights fp16-representable).  Used for the small-channel convs of ControlNet's input_hint_block.
 * d_x fp32 NCHW (in_nchw) or NHWC. */
int sgn_im2col3x3_split_f16(const float* d_x, int in_nchw, int B, int H, int W, int C, int stride, void* d_out,
                            void* stream);
/* 3x3 / pad 1 conv for (Cin, Cout) in {(16,16), (16,32), (32,32)} at stride 1 and (16,32) at stride 2 (ControlNet
 * input_hint_block at the sheet resolution) on mma.sync with the fp32 input split into fp16 hi + lo in shared memory: the fp32 convolution to fp32
 * rounding for fp16-representable weights.  d_x fp32 NHWC, d_w16 fp16 [Cout, 9*Cin] (k = (ky*3+kx)*Cin + c), out NHWC
 * fp32 / fp16 = act(conv + bias). */
int sgn_conv3x3_small_tc(const float* d_x, const void* d_w16, const float* d_bias, int B, int H, int W, int Cin, int Cout,
                         int stride, int act_silu, int out_f16, void* d_out, void* stream);

/* ------------------------------------------------------------------ VAE + A1111 inpaint pre / post */
/* What the A1111 server does around the denoising loop for the request of signerf/diffuser/diffuser.py:132-169
 * (mask_blur 4, inpainting_fill 1, inpaint_full_res 0; A1111 modules/processing.py StableDiffusionProcessingImg2Img.init,
 * decode_latent_batch, apply_overlay).  The autoencoder's convolutions / linears run on sgn_conv3x3_f16 / sgn_gemm_f16;
 * signerf_b200/vae.py walks the ldm Encoder / Decoder graphs.  The integer image operators are bit-exact against
 * cv2 4.13 / Pillow 12.2, which is where A1111 gets them from (tests/golden/inpaint.npz). */

/* ldm Downsample (`F.pad(x, (0,1,0,1)); Conv2d(3, stride 2, padding 0)`): im2col of fp32 NHWC [B,H,W,C] ->
 * fp16 [B*Ho*Wo, 9*C], Ho = (H-2)/2 + 1, input pixel (2oy+ky, 2ox+kx), zero beyond the bottom / right edge. */
int sgn_im2col3x3_s2_asym_f16(const float* d_x, int B, int H, int W, int C, void* d_out, void* stream);
/* hi / lo split producers for the VAE's exact mode: fp32 [M,C] -> fp16 [M,2C]; nearest x2 upsample fp32 NHWC ->
 * fp16 [B,2H,2W,2C]; the asymmetric stride-2 im2col with rows [hi(9C) | lo(9C)]. */
int sgn_split_f16(const float* d_x, int64_t M, int C, void* d_out, void* stream);
int sgn_upsample2x_split_f16(const float* d_x, int B, int H, int W, int C, void* d_out, void* stream);
int sgn_im2col3x3_s2_asym_split_f16(const float* d_x, int B, int H, int W, int C, void* d_out, void* stream);
/* ldm AttnBlock: out[m][:] = softmax(scale * scores[m][:]) as fp16; scores fp32 [M,N] (N % 4 == 0) from sgn_gemm_f16. */
int sgn_softmax_rows_f16(const float* d_scores, int64_t M, int N, float scale, void* d_out, void* stream);
/* quant_conv / post_quant_conv: 1x1 convolution over <= 16 channels, NCHW fp32; h_w [Cout,Cin], h_bias [Cout] on the
 * HOST (they travel as kernel parameters); the input is multiplied by in_scale first (decode: 1 / scale_factor). */
int sgn_pointwise_nchw(const float* d_x, const float* h_w, const float* h_bias, int B, int Cin, int Cout, int64_t HW,
                       float in_scale, float* d_out, void* stream);
/* DiagonalGaussianDistribution: d_moments [B,2Z,HW] = (mean | logvar); out [B,Z,HW] = scale * (mean + exp(0.5 *
 * clamp(logvar,-30,20)) * noise); d_noise NULL = the mode. */
int sgn_vae_sample_latent(const float* d_moments, const float* d_noise, int B, int Z, int64_t HW, float scale,
                          float* d_out, void* stream);
/* uint8 HWC [H,W,3] -> fp32 NCHW [1,3,H,W] = 2 * (v / 255) - 1, and back: uint8(255 * clamp((x + 1) / 2, 0, 1)). */
int sgn_u8_to_vae_input(const uint8_t* d_img, int H, int W, float* d_out, void* stream);
int sgn_vae_output_to_u8(const float* d_x, int H, int W, uint8_t* d_out, void* stream);

/* cv2.GaussianBlur(src, (ksize,1) | (1,ksize), sigma) on CV_8U with BORDER_REFLECT_101, bit-exact: OpenCV's Q8 taps
 * (error diffusion towards the centre, sum exactly 256), result (sum + 128) >> 8.  ksize odd, <= 63. */
int sgn_gaussian_blur_u8(const uint8_t* d_in, int H, int W, int ksize, double sigma, int horizontal, uint8_t* d_out,
                         void* stream);
/* The ksize Q8 taps sgn_gaussian_blur_u8 uses (host-side probe for the parity tests). */
int sgn_gaussian_kernel_q8(int ksize, double sigma, int* h_taps);
/* PIL.Image.resize((w,h), BICUBIC) of one 8-bit band, bit-exact (Pillow Resample.c: 22-bit fixed-point taps, horizontal
 * pass through uint8, then vertical).  d_ws: sgn_pil_resize_ws_bytes(H,W,h,w) bytes of device scratch. */
int64_t sgn_pil_resize_ws_bytes(int H, int W, int h, int w);
int sgn_pil_resize_bicubic_u8(const uint8_t* d_in, int H, int W, int h, int w, void* d_ws, uint8_t* d_out, void* stream);
/* mask_for_overlay = clip(2 * blurred, 0, 255). */
int sgn_inpaint_overlay_mask_u8(const uint8_t* d_blurred, int64_t n, uint8_t* d_overlay, void* stream);
/* keep[i] = 1 - round(lat_u8[i] / 255): A1111's `mask = 1 - latmask`, the d_mask of sgn_cfg_euler_step. */
int sgn_latent_keep_mask(const uint8_t* d_lat_u8, int64_t n, float* d_keep, void* stream);
/* apply_overlay: PIL paste of the original through the inverted overlay mask + alpha_composite onto the generated image
 * (Pillow's integer arithmetic); uint8 HWC in, uint8 HWC and / or fp32 HWC (= image_to_tensor: v / 255) out. */
int sgn_overlay_composite_u8(const uint8_t* d_generated, const uint8_t* d_original, const uint8_t* d_overlay_mask, int H,
                             int W, uint8_t* d_out_u8, float* d_out_f32, void* stream);

/* ------------------------------------------------------------------ proxy-mesh depth + shape masking */
/* Replaces reference Renderer.render_camera's depth output (signerf/renderer/renderer.py:149-196: pyrender
 * OffscreenRenderer + IntrinsicsCamera(znear 1e-4, zfar 10) over the trimesh loaded in setup(), :64-131) for V views:
 * OpenGL rasterisation rules (pixel centres at +0.5, 8 sub-pixel bits, top-left rule, LESS on a 24-bit depth buffer,
 * back-face culling when cull_back), the Blender -> OpenGL axis swap of :134-146 applied to the camera poses, pyrender's
 * buffer -> metric depth conversion; 0 = empty.
 *   d_vertices [Nv,3] fp32 object space, d_faces [Nf,3] int32; h_model: 4x4 row-major double on the HOST = the object
 *   pose already in OpenGL axes (convert @ [Rz Ry Rx diag(10 scale) | position]); d_c2w [V,3,4], d_intr [V,4] as K1;
 *   d_ws: sgn_rasterize_ws_bytes(Nv,V,H,W) bytes of 16-byte aligned scratch; d_depth [V,H,W,1] fp32 out. */
int64_t sgn_rasterize_ws_bytes(int Nv, int V, int H, int W);
int sgn_rasterize_depth(const float* d_vertices, const int32_t* d_faces, int Nv, int Nf, const double* h_model,
                        const float* d_c2w, const float* d_intr, int V, int H, int W, double znear, double zfar, int cull_back,
                        void* d_ws, float* d_depth, void* stream);
/* The colour image `Renderer.render_camera` returns next to the depth (renderer.py:187-196): the scene holds one mesh lit
 * by ambient light 1.0 only (renderer.py:129-130), so every covered pixel has the material's flat base colour and every
 * other pixel pyrender's background.  d_depth [npix] fp32 (0 = empty) -> d_color [npix,3] uint8; h_fg / h_bg: 3 bytes on
 * the HOST. */
int sgn_shape_color_u8(const float* d_depth, int64_t npix, const uint8_t* h_fg_rgb, const uint8_t* h_bg_rgb,
                       uint8_t* d_color, void* stream);
/* render_camera's masking_mode == "shape" branch (datasetgenerator.py:711-757) for V views: visible = (proxy < depth) &
 * (proxy > 0) (inverted when inverse_mask), cv2.dilate, min over visible proxy depths - radius, max over the whole proxy
 * image + radius (or manual_depth), condition = 1 - clamp(visible * proxy_n + ~visible * depth_n).  Same outputs /
 * stats as sgn_mask_condition; SgnMaskOpts.aabb is ignored. */
int sgn_mask_condition_shape(const float* d_proxy_depth, const float* d_depth, int V, int H, int W, const SgnMaskOpts* o,
                             uint8_t* d_mask, float* d_cond, float* d_stats, void* stream);

/* ------------------------------------------------------------------ prompt conditioning ("also" clause)
 * The A1111 server the reference posts to (diffuser.py:132-180, `prompt` / `negative_prompt`) encodes the prompt with
 * SDXL's two text transformers once per request: CLIP ViT-L/14 (12 layers, width 768, quick_gelu) and OpenCLIP ViT-bigG/14
 * (32 layers, width 1280, gelu), both causal over 77 tokens, head_dim 64.  The layers run on sgn_layer_norm_f16 /
 * sgn_gemm_f16 plus the three entry points below (signerf_b200/text_encoder.py walks them). */
/* softmax(Q K^T * scale + causal mask) V for self-attention over T <= 80 tokens (query i sees keys <= i); layout as
 * sgn_attention_f16. */
int sgn_attention_causal_f16(const void* d_q, int64_t ldq, const void* d_k, int64_t ldk, const void* d_v, int64_t ldv,
                             int B, int heads, int T, float scale, void* d_out, int64_t ldo, void* stream);
/* fp32 [n] -> fp16 [n]: mode 0 quick_gelu x * sigmoid(1.702 x), mode 1 exact (erf) GELU. */
int sgn_act_f16(const float* d_x, int64_t n, int mode, void* d_out, void* stream);
/* CLIPTextEmbeddings: out[r] = token_table[ids[r]] + position_table[r % T]; ids int32 [rows], tables fp32 [vocab | T, width]. */
int sgn_embed_tokens(const int32_t* d_ids, const float* d_token_table, const float* d_position_table, int rows, int T,
                     int width, int vocab, float* d_out, void* stream);

/* ------------------------------------------------------------------ the NeRF fine-tune step
 * `SIGNeRFModel` trains with nerfacto's forward and `get_loss_dict` (signerf/signerf.py:41-82) between two dataset
 * generations.  These entry points are the main field's forward + backward on a batch of rays, the image loss and the
 * optimizer update, fp32 end to end (gradients = torch autograd through nerfstudio's torch modules).
 *
 * Parameters: the hash table stays the caller's tensor (SgnHashGrid.d_table, updated in place by sgn_adam_step); the MLPs
 * live in the field's fp32 parameter block of sgn_mlp_param_count() floats, layout (row-major [out][in], nn.Linear):
 *   w_base0[64*32] w_base1[16*64] w_head0[64*32] w_head1[64*64] w_head2[3*64] b_base0[64] b_base1[16] b_head0[64]
 *   b_head1[64] b_head2[4] avg_density pad[3]
 * w_head0's 32 input columns are [SH 16 | unused | geo 15]; b_head0 carries the mean appearance embedding folded in
 * (b + W_app . mean(embedding)), as the eval renderer uses it.  Gradient buffers use the same layouts and are ACCUMULATED
 * into (zero them per step).  LPIPS's pretrained network stays a host-side torch term. */
int64_t sgn_mlp_param_count(void);
int sgn_field_mlp_params(const SgnField* f, float** d_params);
/* Re-derives the fp16 tensor-core fragments (and the feature scale) the renderer uses from the fp32 parameter block and
 * the current hash table - call after optimizer steps, before the next sgn_render_*.  Synchronises the stream. */
int sgn_field_refresh(SgnField* f, void* stream);
/* Forward of N rays x S samples through the main field: d_bins [S+1] euclidean bin edges shared by all rays, or
 * d_ray_bins [N,S+1] (exactly one non-NULL).  d_head_bias [N,64] or NULL: per-ray bias of the head's first layer
 * (sgn_appearance_bias: training with per-image appearance embeddings; NULL = the folded mean embedding, eval semantics).
 * Outputs: d_sigma [N,S], d_color [N,S,3] (kept by the caller for the backward), d_rgb [N,3] = sum w c + c_last (1 - sum w)
 * without the eval clamp, d_acc [N] or NULL. */
int sgn_train_forward(const SgnField* f, const float* d_origins, const float* d_directions, int64_t N, const float* d_bins,
                      const float* d_ray_bins, int S, const float* d_head_bias, float* d_sigma, float* d_color, float* d_rgb,
                      float* d_acc, void* stream);
/* Backward: d_grad_rgb [N,3] = dL/drgb and, optionally, d_grad_weights [N,S] = the gradient of loss terms that read the
 * final level's weights directly (sgn_distortion_loss) and d_grad_geo [N,S,15] = dL / d(geo features) of the normal
 * prediction branch (sgn_train_normals_backward) -> d_grad_table [L*T,2], d_grad_mlp [sgn_mlp_param_count()] and, when
 * d_head_bias was given, d_grad_head_bias [N,64] (all +=).  d_ws: sgn_train_ws_bytes(N, S) bytes, 16-byte aligned. */
int64_t sgn_train_ws_bytes(int64_t N, int S);
int sgn_train_backward(const SgnField* f, const float* d_origins, const float* d_directions, int64_t N, const float* d_bins,
                       const float* d_ray_bins, int S, const float* d_head_bias, const float* d_sigma, const float* d_color,
                       const float* d_grad_rgb, const float* d_grad_weights, const float* d_grad_geo, float* d_grad_table,
                       float* d_grad_mlp, float* d_grad_head_bias, void* d_ws, int64_t ws_bytes, void* stream);
/* Per-image appearance embedding while training ([EXT] NerfactoField.get_outputs: embedding_appearance(camera_indices)
 * concatenated to the head's input): head_bias[ray] = bias + W_app . E[camera[ray]], with W_app [64,32] the head's first
 * layer's appearance columns, bias [64], E [num_images,32], camera indices int32 [N]; the backward accumulates (+=) the
 * gradients of all three from d_grad_head_bias [N,64]. */
int sgn_appearance_bias(const float* d_w_app, const float* d_bias, const float* d_embedding, int num_images,
                        const int32_t* d_camera_indices, int64_t N, float* d_head_bias, void* stream);
int sgn_appearance_bias_backward(const float* d_w_app, const float* d_embedding, int num_images, const int32_t* d_camera_indices,
                                 const float* d_grad_head_bias, int64_t N, float* d_grad_w_app, float* d_grad_bias,
                                 float* d_grad_embedding, void* stream);

/* --- the normal regularisers of `predict_normals=True` (signerf_config.py:33; loss terms signerf/signerf.py:69-80) ---
 * Parameter block of the prediction branch ([EXT] NerfactoField.mlp_pred_normals + PredNormalsFieldHead), fp32,
 * sgn_pred_normals_param_count() = 10 308 floats, 16-byte aligned, nn.Linear row-major [out][in]:
 *   w1[64*64] w2[64*64] w_head[3*64] w0[64*27] b0[64] b1[64] b2[64] b_head[3] pad[1]
 * (input of layer 0: NeRFEncoding of the raw position, 12 values, then the 15 geo features).
 * Forward on the final level's bins d_ray_bins [N,S+1]: d_normals [N,S,3] = -normalize(d density_logit / d p), the analytic
 * normals ([EXT] Field.get_normals: no graph, constants of the step); d_pred_normals [N,S,3] = the branch's output. */
int64_t sgn_pred_normals_param_count(void);
int sgn_train_normals_forward(const SgnField* f, const float* d_pn_params, const float* d_origins, const float* d_directions,
                              int64_t N, const float* d_ray_bins, int S, float* d_normals, float* d_pred_normals, void* stream);
/* [EXT] losses.py orientation_loss / pred_normal_loss on the (detached) weights [N,S], means over the rays times the
 * multipliers: d_loss_* [1] += the terms; d_grad_pred [N,S,3] = d pred_normal term / d pred_normals (overwritten).  The
 * orientation term has no gradient (weights and analytic normals carry no graph). */
int sgn_normal_losses(const float* d_weights, const float* d_normals, const float* d_pred_normals, const float* d_directions,
                      int64_t N, int S, float orientation_mult, float pred_normal_mult, float* d_loss_orientation,
                      float* d_loss_pred_normal, float* d_grad_pred, void* stream);
/* Backward of the prediction branch: d_grad_pred -> d_grad_pn_params [sgn_pred_normals_param_count()] (+=) and d_grad_geo
 * [N,S,15] (overwritten; feed it to sgn_train_backward).  d_ws: sgn_train_normals_ws_bytes(N, S) bytes, 16-byte aligned. */
int64_t sgn_train_normals_ws_bytes(int64_t N, int S);
int sgn_train_normals_backward(const SgnField* f, const float* d_pn_params, const float* d_origins, const float* d_directions,
                               int64_t N, const float* d_ray_bins, int S, const float* d_grad_pred, float* d_grad_pn_params,
                               float* d_grad_geo, void* d_ws, int64_t ws_bytes, void* stream);

/* The three calls above in ONE pass over the samples (what the trainer uses): analytic normals, both loss terms
 * (d_loss_* [1] +=) and the backward of the prediction branch from the detached weights d_weights [N,S] - the separate
 * forward and loss kernels would gather and run the base MLP a second time. */
int sgn_train_normals_step(const SgnField* f, const float* d_pn_params, const float* d_origins, const float* d_directions,
                           int64_t N, const float* d_ray_bins, int S, const float* d_weights, float orientation_mult,
                           float pred_normal_mult, float* d_loss_orientation, float* d_loss_pred_normal,
                           float* d_grad_pn_params, float* d_grad_geo, void* d_ws, int64_t ws_bytes, void* stream);

/* --- the proposal half of the training step ([EXT] nerfstudio ProposalNetworkSampler in training mode, losses.py
 * interlevel_loss / distortion_loss as signerf/signerf.py:62-68 adds them to the loss dict) ---
 * A proposal network's trainable parameters: its 5-level hash table (the caller's SgnHashGrid.d_table) and
 * sgn_prop_param_count() = 193 contiguous floats inside the field, layout w0[16*10] b0[16] w1[16] b1 (row-major nn.Linear). */
int64_t sgn_prop_param_count(void);
int sgn_field_prop_params(const SgnField* f, int level, float** d_params);
/* Samples of one training batch, all caller-allocated: level 0 = the S0 stratified initial bins, level 1 / 2 = the PDF
 * re-samplings (S1, S2 bins).  d_spacing / d_euclid [N, S_l + 1] bin edges in the spacing domain / in metres; d_sigma /
 * d_weights [N, S_l] of the proposal network evaluated on level l (l = 0, 1). */
typedef struct SgnTrainSamples {
  float* d_spacing[3];
  float* d_euclid[3];
  float* d_sigma[2];
  float* d_weights[2];
} SgnTrainSamples;
/* ProposalNetworkSampler.generate_ray_samples while training: d_jitter [3,N] uniform draws in [0,1) (initial sampler, PDF
 * level 1, PDF level 2; one per ray = single_jitter) or NULL for the eval bins (bin centres).  The draws are inputs so that
 * the oracle and this path sample the same bins.  anneal: the proposal weights are raised to this power before each
 * re-sampling (`use_proposal_weight_anneal`: bias(step / 1000, slope 10), 1 past the warm-up and in eval - the reference's
 * trainer resets the step count, signerf_trainer.py:321-325, so a fine-tune starts at 0 = uniform re-sampling).
 * d_ws: sgn_train_sample_ws_bytes bytes, 16-byte aligned. */
int64_t sgn_train_sample_ws_bytes(int64_t N, int S0, int S1);
int sgn_train_sample(const SgnField* f, const float* d_origins, const float* d_directions, int64_t N, int S0, int S1, int S2,
                     float near_plane, float far_plane, const float* d_jitter, float anneal, const SgnTrainSamples* out,
                     void* d_ws, int64_t ws_bytes, void* stream);
/* RaySamples.get_weights: d_euclid [N,S+1], d_sigma [N,S] -> d_weights [N,S]. */
int sgn_weights_from_density(const float* d_euclid, const float* d_sigma, int64_t N, int S, float* d_weights, void* stream);
/* losses.py lossfun_outer of one proposal level against the final level's histogram (weights detached), mean over
 * N * S_final, times mult: d_loss [1] += the term, d_grad_weights_prop [N,S_prop] = its gradient (overwritten).
 * d_ws: N * (S_prop + 1) floats. */
int sgn_interlevel_loss(const float* d_spacing_final, const float* d_weights_final, int S_final, const float* d_spacing_prop,
                        const float* d_weights_prop, int S_prop, int64_t N, float mult, float* d_loss,
                        float* d_grad_weights_prop, void* d_ws, int64_t ws_bytes, void* stream);
/* losses.py distortion_loss on the final level (mip-NeRF 360 eq. 15), mean over rays, times mult: d_loss [1] += the term,
 * d_grad_weights [N,S] = its gradient (overwritten; NULL to skip) - feed it to sgn_train_backward. */
int sgn_distortion_loss(const float* d_spacing, const float* d_weights, int64_t N, int S, float mult, float* d_loss,
                        float* d_grad_weights, void* stream);
/* Backward of proposal network `level` on its N x S samples: d_grad_weights [N,S] -> d_grad_table [5*T,2],
 * d_grad_mlp [sgn_prop_param_count()] (+=).  d_ws: N * S floats. */
int sgn_prop_backward(const SgnField* f, int level, const float* d_origins, const float* d_directions, int64_t N, int S,
                      const float* d_euclid, const float* d_sigma, const float* d_grad_weights, float* d_grad_table,
                      float* d_grad_mlp, void* d_ws, int64_t ws_bytes, void* stream);
/* `rgb_loss` of signerf/signerf.py:36-47: nerfstudio L1Loss (l1 = 1) or MSELoss (l1 = 0) = mean over all n elements;
 * d_loss [1]; d_grad [n] = d loss / d pred, or NULL. */
int sgn_rgb_loss(const float* d_pred, const float* d_target, int64_t n, int l1, float* d_loss, float* d_grad, void* stream);
/* torch.optim.Adam as nerfstudio's AdamOptimizerConfig sets it up (signerf_config.py:43-50: lr 1e-2, eps 1e-15, betas
 * 0.9 / 0.999, no weight decay); `step` counts from 1. */
int sgn_adam_step(float* d_param, const float* d_grad, float* d_m, float* d_v, int64_t n, float lr, float beta1, float beta2,
                  float eps, int step, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SIGNERF_B200_H */
