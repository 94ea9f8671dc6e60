/*
 * nvo_b200 — C ABI of the B200-native (sm_100a) NeRF mapping hot path.
 *
 * Drop-in boundary for the operator surface that NeRF-VO's mapping thread reaches through
 * nerfstudio -> tinycudann (reference paths relative to /root/reference):
 *   - nerf_vo/thirdparty/tiny_cuda_nn/include/tiny-cuda-nn/cpp_api.h:86-111   (tcnn::cpp::Module:
 *     inference / forward / backward on raw device pointers + cudaStream_t) and
 *   - nerf_vo/thirdparty/tiny_cuda_nn/bindings/torch/tinycudann/bindings.cpp:282-336 (what pybind exposes),
 * widened to the nerfstudio torch-path operators the field/sampler/renderer call
 * (NS = nerf_vo/thirdparty/nerfstudio/nerfstudio): each entry point cites the function it replaces.
 *
 * Conventions
 *   - plain C types only; every pointer is a DEVICE pointer unless its name starts with h_;
 *   - the caller (PyTorch) owns every buffer; the library allocates nothing and keeps no state
 *     except a thread-local error string;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises;
 *   - return value 0 = success; non-zero = failure, message via nvo_last_error()
 *     (mirrors CHECK_THROW -> std::runtime_error of bindings.cpp:54-55; the Python shim raises RuntimeError);
 *   - arithmetic follows the reference's *torch* implementation (the parity oracle): always-hashed levels,
 *     floor/ceil corners, no +0.5 offset, biased Linear layers (SURVEY.md §8 a-notes);
 *   - gradient outputs documented as "accumulate" are added into (atomics); the caller zero-fills.
 */
#ifndef NVO_B200_H
#define NVO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NVO_MAX_LEVELS 32
#define NVO_MAX_LAYERS 6

/* dtype tags for the table / feature buffers */
/* NVO_F16_TMH (grid forward output only): fp16 in the tensor-core MLP's tile-major operand layout, see nvo_mlp_tc_forward */
/* NVO_F32_TMF (dy of nvo_grid_backward / nvo_grid_backward_input only): fp32 tile-major, see nvo_mlp_tc_backward's dx */
enum { NVO_F32 = 0, NVO_F16 = 1, NVO_F16_TMH = 2, NVO_F32_TMF = 3 };
/* activations (tcnn network_config "activation"/"output_activation", NS/field_components/mlp.py:34-58) */
/* NVO_ACT_TRUNC_EXP: exp forward, backward g*exp(clamp(x,-15,15)) (NS/field_components/activations.py:28-41) */
enum { NVO_ACT_NONE = 0, NVO_ACT_RELU = 1, NVO_ACT_SIGMOID = 2, NVO_ACT_TANH = 3, NVO_ACT_EXP = 4, NVO_ACT_TRUNC_EXP = 5 };

const char* nvo_last_error(void);
int nvo_version(void);
/* bindings.cpp:285 batch_size_granularity(); ours is the tcgen05 M tile */
int nvo_batch_size_granularity(void);
/* number of entry-point calls that enqueued GPU work since the library was loaded (bench.py's gpu_launches evidence) */
int64_t nvo_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * Multiresolution hash grid — replaces NS/field_components/encodings.py:405-465 (HashEncoding.pytorch_fwd,
 * hash_fn :405-422) and tcnn kernel_grid / kernel_grid_backward / kernel_grid_backward_input
 * (tiny-cuda-nn/encodings/grid.h:48-349).  F (features per level) is 2.
 * table: [n_levels * 2^log2_T, 2] row-major, level-major (encodings.py:351,381).
 * scalings: the fp32 per-level scale computed by the caller exactly as encodings.py:347-349.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int32_t n_levels;
    int32_t log2_T;
    int32_t table_dtype; /* NVO_F32 | NVO_F16 */
    int32_t out_dtype;   /* NVO_F32 | NVO_F16 : dtype of y / dy (row-major [n,2L]); NVO_F16_TMH: y as TMH tiles, forward only */
    float scalings[NVO_MAX_LEVELS];
} nvo_grid_desc;

/* y[n, 2L] = encode(x[n,3]) */
int nvo_grid_forward(const nvo_grid_desc* d, void* stream, int64_t n, const float* x, const void* table, void* y);
/* dtable[L*T,2] (fp32, accumulate) += scatter(dy[n,2L]) */
int nvo_grid_backward(const nvo_grid_desc* d, void* stream, int64_t n, const float* x, const void* dy, float* dtable);
/* dx[n,3] = d(sum(y*dy))/dx  (re-gathers the table instead of storing dy_dx, cf. grid.h:322-349) */
int nvo_grid_backward_input(const nvo_grid_desc* d, void* stream, int64_t n, const float* x, const void* table, const void* dy, float* dx);
/* nvo_grid_forward with out_dtype NVO_F16_TMH that also saves d(feature)/d(x) (tcnn's `dy_dx`, grid.h:160-211, stored when input
 * gradients are prepared): jac = fp16 [tile][chunk][axis 0..2][row 0..127][8], the derivative of the chunk's 8 features w.r.t. the
 * level-scaled coordinate (NOT yet multiplied by scale_l); 3 x the TMH feature buffer's size. */
int nvo_grid_forward_jac(const nvo_grid_desc* d, void* stream, int64_t n, const float* x, const void* table, void* y, void* jac);
/* dx[n,3] = sum_l scale_l * jac_l^T dy_l from the saved derivatives; dy = fp32 tile-major [tile][2L][128] (NVO_F32_TMF).
 * normalize_scale != 0: dx is written as normalize_scale * v / max(|v|, eps) (density-gradient normals, NS/fields/base_field.py:97-99). */
int nvo_grid_jac_dx(const nvo_grid_desc* d, void* stream, int64_t n, const void* jac, const float* dy, float normalize_scale, float eps, float* dx);
/* idx[n, L, 8] int64: table row of every corner in the reference's corner order (encodings.py:435-442). Test hook. */
int nvo_grid_indices(const nvo_grid_desc* d, void* stream, int64_t n, const float* x, int64_t* idx);

/* ---------------------------------------------------------------------------------------------
 * MLP — replaces NS/field_components/mlp.py:160-179 (MLP.pytorch_fwd) and tcnn FullyFusedMLP
 * (tiny-cuda-nn/src/fully_fused_mlp.cu:499-557 forward, :150-258 backward, :760-830 weight gradients).
 * params: torch layout, fp32: for each layer W[out,in] row-major then b[out], concatenated.
 * acts[l] is the activation applied to the output of layer l (hidden layers: ReLU; last: the output activation).
 * Per-layer activations let a trailing Linear head (e.g. PredNormalsFieldHead, NS/field_components/field_heads.py:189-204)
 * run inside the same fused kernel.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    int32_t n_layers;
    int32_t in_dim;
    int32_t dims[NVO_MAX_LAYERS]; /* output width of each layer */
    int32_t acts[NVO_MAX_LAYERS]; /* NVO_ACT_* per layer */
} nvo_mlp_desc;

int64_t nvo_mlp_n_params(const nvo_mlp_desc* d);
/* floats of scratch per sample the training forward writes (post-activation outputs of every layer) */
int64_t nvo_mlp_saved_per_sample(const nvo_mlp_desc* d);
/* y[n, dims[last]]; saved (nullable => inference) [n, saved_per_sample];
 * row_mask (nullable) [n] of 0/1: y is multiplied by it after the output activation (the `density * selector` of
 * NS/fields/nerfacto_field.py:222, density_fields.py:115). */
int nvo_mlp_forward(const nvo_mlp_desc* d, void* stream, int64_t n, const float* x, const float* params, const float* row_mask, float* y,
                    float* saved);
/* dx (nullable) [n,in_dim]; dparams (nullable, accumulate) [n_params] */
int nvo_mlp_backward(const nvo_mlp_desc* d, void* stream, int64_t n, const float* x, const float* params, const float* saved,
                     const float* y, const float* row_mask, const float* dy, float* dx, float* dparams);

/* Tensor-core path (tcgen05.mma, accumulators in TMEM) for networks whose widths are <= 64 and depth <= 4 — the three
 * field MLPs.  Same semantics and parameter layout as nvo_mlp_forward/backward with fp16 operands and fp32 accumulation.
 *   x16     the input in TMH layout ("tile-major half"): tiles of 128 rows, each a contiguous in_pad*256-byte block
 *           [chunk = col/8][row][8 halfs] (the UMMA canonical operand layout, moved to shared memory by one bulk copy);
 *           in_pad = nvo_mlp_tc_in_pad(d) (in_dim rounded up to 16); ceil(n/128) tiles, padding columns / rows ZERO.
 *           nvo_cast_pad_f16 produces it from fp32 [n,in_dim] row-major; the hash grid and the field assembly write it directly.
 *   wimage  the packed fp16 weight image of nvo_mlp_tc_wimage_bytes(d) bytes, refreshed with nvo_mlp_tc_pack_weights
 *           whenever `params` (torch layout, fp32, as nvo_mlp_forward) change;
 *   saved   opaque forward context of nvo_mlp_tc_saved_bytes(d, n) bytes (fp16 hidden activations, TMH layout);
 *   scratch one float the backward uses for its device-side gradient scale (max|dy| -> power-of-two loss scale, the
 *           device analogue of tinycudann's loss_scale, modules.py:174); dy_absmax_hint > 0 supplies max|dy| from the
 *           caller instead (e.g. 1 for a one-hot seed) and skips the reduction pass; dy_absmax_hint < 0: `scratch` already holds
 *           max|dy| as a float bit pattern, written by the kernel that produced dy (nvo_field_assemble_backward);
 *   y, dy, dparams fp32 row-major exactly as the SIMT entry points;
 *   dx      fp32 in TMF layout ("tile-major float"): [ceil(n/128)][in_dim][128], i.e. column-major inside each 128-row tile, so
 *           the kernel's row-per-thread epilogue stores coalesced and the per-sample consumers (nvo_grid_backward with
 *           out_dtype NVO_F32_TMF, nvo_field_assemble_backward with tmf != 0) read coalesced; nvo_tmf_to_rows converts to [n,in_dim]. */
int nvo_mlp_tc_in_pad(const nvo_mlp_desc* d);
int64_t nvo_mlp_tc_saved_bytes(const nvo_mlp_desc* d, int64_t n);
int64_t nvo_mlp_tc_wimage_bytes(const nvo_mlp_desc* d);
int nvo_cast_pad_f16(void* stream, int64_t n, int32_t in_dim, int32_t kpad, const float* x, void* out);
int nvo_mlp_tc_pack_weights(const nvo_mlp_desc* d, void* stream, const float* params, void* wimage);
int nvo_tmf_to_rows(void* stream, int64_t n, int32_t K, const float* src, float* dst);
int nvo_mlp_tc_forward(const nvo_mlp_desc* d, void* stream, int64_t n, const void* x16, const void* wimage, const float* row_mask, float* y,
                       void* saved);
int nvo_mlp_tc_backward(const nvo_mlp_desc* d, void* stream, int64_t n, const void* x16, const void* wimage, const void* saved, const float* y,
                        const float* row_mask, const float* dy, float dy_absmax_hint, float* scratch, float* dx, float* dparams);
/* nvo_mlp_tc_backward on input / saved-activation tiles that are embedded in larger per-tile records (the fused field kernels' saved tiles):
 * consecutive tiles lie x_tile_bytes / saved_tile_bytes apart. */
int nvo_mlp_tc_backward_strided(const nvo_mlp_desc* d, void* stream, int64_t n, const void* x16, int64_t x_tile_bytes, const void* wimage,
                                const void* saved, int64_t saved_tile_bytes, const float* y, const float* row_mask, const float* dy,
                                float dy_absmax_hint, float* scratch, float* dx, float* dparams);

/* ---------------------------------------------------------------------------------------------
 * Fused proposal density field — replaces HashMLPDensityField.density_fn / get_density
 * (NS/fields/density_fields.py:93-116, NS/fields/base_field.py:48-68) as ProposalNetworkSampler calls it
 * (NS/model_components/ray_samplers.py:604-610): position -> SceneContraction -> (x+2)/4 -> selector -> hash grid (5 levels)
 * -> Linear(10,16)+ReLU -> Linear(16,1) -> trunc_exp * selector in one kernel; nothing but the density (and, for
 * training, the 2L interpolated features) is written to HBM.
 * Sample points: `positions` [B*S,3] when non-NULL, else o + d*(start+end)/2 from (origins, directions, starts, ends, stride).
 * params: the MLP in torch layout (W0[16,10], b0[16], W1[1,16], b1[1]), fp32.
 * feat (nullable => inference): opaque buffer of nvo_prop_density_feat_floats(L, B*S) floats saved for the backward.  nvo_prop_density_supported tells the host
 * whether a (levels, hidden width, layers) combination has a fused kernel; others use the unfused operators.
 * backward: dtable (fp32, accumulate, nullable), dparams (accumulate, nullable) from ddensity [B*S].
 * ------------------------------------------------------------------------------------------- */
int nvo_prop_density_supported(int32_t n_levels, int32_t hidden, int32_t n_layers);
/* floats of the saved-feature buffer for n samples (level-major, padded to 128-sample blocks, permuted inside a block) */
int64_t nvo_prop_density_feat_floats(int32_t n_levels, int64_t n);
/* slot in [0,4): which constant-memory bank holds this network's MLP parameters while its kernels run (the call copies
 * `params` there first, device to device, on `stream`); networks evaluated concurrently on different streams need distinct slots.
 * params == NULL in the forward / backward calls: the bank already holds the network — the caller has run nvo_prop_density_upload(stream',
 * slot, params) (the forward's and the backward's copy), ordered the call behind it, and nothing else has used the slot since.  The
 * mapping step uploads both proposal networks at its start, next to the first sampling kernel, instead of in front of each of the four
 * launches that use them. */
int nvo_prop_density_upload(void* stream, int32_t slot, const float* params);
int nvo_prop_density_forward(const nvo_grid_desc* g, int32_t hidden, int32_t slot, void* stream, int64_t B, int32_t S, const float* origins,
                             const float* directions, const float* starts, const float* ends, int64_t stride, const float* positions, const void* table,
                             const float* params, float* density, float* feat);
int nvo_prop_density_backward(const nvo_grid_desc* g, int32_t hidden, int32_t slot, void* stream, int64_t B, int32_t S, const float* origins,
                              const float* directions, const float* starts, const float* ends, int64_t stride, const float* positions,
                              const float* params, const float* feat, const float* ddensity, float* dtable, float* dparams);
/* The same backward split in two: this call runs the MLP part only (weight gradients into dparams) and writes the feature gradients as
 * fp32 tile-major [tile][2L][128] (NVO_F32_TMF) plus the normalised, selector-masked positions xq [ceil(n/128)*128, 3]; the caller then
 * scatters them with nvo_grid_backward(out_dtype = NVO_F32_TMF), whose long-run kernel merges equal-cell runs over 16 samples of a ray. */
int nvo_prop_density_backward_split(const nvo_grid_desc* g, int32_t hidden, int32_t slot, void* stream, int64_t B, int32_t S, const float* params,
                                    const float* feat, const float* ddensity, float* dparams, float* dfeat_tmf, float* xq);

/* ---------------------------------------------------------------------------------------------
 * Field element-wise operators.
 * ------------------------------------------------------------------------------------------- */
/* SceneContraction(order=inf) + (x+2)/4 + selector masking (NS/field_components/spatial_distortions.py:67-69,
 * NS/fields/nerfacto_field.py:201-209): pos[n,3] -> x[n,3] (zeroed where outside (0,1)), selector[n] (0/1 as float).
 * contract=0 skips the contraction and maps through the aabb instead: x = (pos - aabb_min)/(aabb_max-aabb_min). */
int nvo_contract_forward(void* stream, int64_t n, const float* pos, float* x, float* selector);
/* nvo_sample_positions followed by nvo_contract_forward in one pass (Frustums.get_positions, NS/cameras/rays.py:49-58, then
 * SceneContraction + normalisation + selector, spatial_distortions.py:67-69, nerfacto_field.py:201-209): positions[B*S,3] (world),
 * x[B*S,3] (normalised, selector-masked), selector[B*S].  starts / ends: Euclidean interval edges with row stride `stride`. */
int nvo_sample_positions_contract(void* stream, int64_t B, int32_t S, const float* origins, const float* directions, const float* starts,
                                  const float* ends, int64_t stride, float* positions, float* x, float* selector);
/* SHEncoding degree 4 on the vector as given (NS/utils/math.py:29-93): d[n,3] -> out[n,16] */
int nvo_sh4_forward(void* stream, int64_t n, const float* d, float* out);
/* NeRFEncoding torch path (NS/field_components/encodings.py:170-176): sin(cat[u, u+pi/2]), u = 2*pi*x_i*2^k;
 * x[n,in_dim] -> out[n, in_dim*n_freq*2] */
int nvo_frequency_forward(void* stream, int64_t n, int32_t in_dim, int32_t n_freq, const float* x, float* out);
/* trunc_exp (NS/field_components/activations.py:28-41) */
int nvo_trunc_exp_forward(void* stream, int64_t n, const float* x, float* y);
int nvo_trunc_exp_backward(void* stream, int64_t n, const float* x, const float* dy, float* dx);
/* out = scale * v / max(|v|, eps) on [n,3] (F.normalize; scale=-1 gives base_field.py:99 normals) and its backward */
int nvo_normalize3_forward(void* stream, int64_t n, const float* v, float scale, float eps, float* out);
int nvo_normalize3_backward(void* stream, int64_t n, const float* v, const float* dout, float scale, float eps, float* dv);
/* Fused input assembly of NerfactoField.get_outputs (NS/fields/nerfacto_field.py:225-297) for B rays x S samples:
 *   density[n]   = trunc_exp(h[:,0]) * selector                      (:214-222)
 *   head_in[n,63]= [SH16((dir+1)/2) | h[:,1:16] | appearance32]      (:286-293, base_field.py:142)
 *   pn_in[n,27]  = [posenc12(pos) | h[:,1:16]]                       (:278-281)
 * appearance = embedding[cam_idx[ray]] (training) or the caller-provided mean vector when cam_idx is NULL (eval;
 * embedding then points at ONE 32-vector).  h[n,16] is mlp_base's output.  pn_in may be NULL.
 * f16_padded != 0: head_in / pn_in are fp16 TMH tiles zero-padded to 64 / 32 columns and to ceil(B*S/128) whole tiles (the
 * operand format of nvo_mlp_tc_forward). */
int nvo_field_assemble_forward(void* stream, int64_t B, int32_t S, const float* h, const float* selector, const float* directions, const float* pos,
                               const int64_t* cam_idx, const float* embedding, int32_t f16_padded, float* density, void* head_in, void* pn_in);
/* backward: dh[n,16] (overwritten) and dembedding[K,32] (accumulate, nullable) from ddensity[n] (nullable), dhead_in[n,63], dpn_in[n,27] (nullable) */
/* tmf != 0: dhead_in / dpn_in are in the TMF layout nvo_mlp_tc_backward writes ([tile][63 | 27][128]) instead of row-major */
/* dh_absmax (nullable, one float, zero-filled by the caller): receives max|dh| (atomic max on the bit pattern), which nvo_mlp_tc_backward
 * accepts in `scratch` with dy_absmax_hint < 0 — the gradient-scale reduction pass over dh then disappears from the backward chain */
int nvo_field_assemble_backward(void* stream, int64_t B, int32_t S, const float* h, const float* selector, const int64_t* cam_idx,
                                const float* ddensity, const float* dhead_in, const float* dpn_in, int32_t tmf, float* dh, float* dembedding,
                                float* dh_absmax);

/* ---------------------------------------------------------------------------------------------
 * Per-ray operators.  B rays, S samples per ray.  Sample intervals are stored as bin EDGES:
 *   sdist [B,S+1] spacing-space edges in [0,1] (RaySamples.spacing_starts/ends, NS/cameras/rays.py:113-116),
 *   ebins [B,S+1] euclidean edges (Frustums.starts = ebins[:, :-1], ends = ebins[:, 1:]).
 * Operators that read euclidean intervals take (starts, ends, stride): element (r,i) is starts[r*stride+i] /
 * ends[r*stride+i]; for compact edges pass (ebins, ebins+1, S+1), for separate [B,S] arrays pass stride S.
 * ------------------------------------------------------------------------------------------- */

/* UniformLinDispPiecewiseSampler (NS/model_components/ray_samplers.py:78-128,225-248).
 * base_bins[S+1] = torch.linspace(0,1,S+1) from the caller (bit-exact); jitter[B] in [0,1) or NULL (eval). */
int nvo_sample_uniform(void* stream, int64_t B, int32_t S, const float* base_bins, const float* jitter, const float* nears, const float* fars,
                       float* sdist, float* ebins);

/* Frustums.get_positions (NS/cameras/rays.py:49-58): pos[B,S,3] = o + d*(start+end)/2 */
int nvo_sample_positions(void* stream, int64_t B, int32_t S, const float* origins, const float* directions, const float* starts, const float* ends,
                         int64_t stride, float* pos);

/* RaySamples.get_weights (NS/cameras/rays.py:128-150) and its backward w.r.t. density */
int nvo_weights_forward(void* stream, int64_t B, int32_t S, const float* starts, const float* ends, int64_t stride, const float* density,
                        float* weights);
int nvo_weights_backward(void* stream, int64_t B, int32_t S, const float* starts, const float* ends, int64_t stride, const float* density,
                         const float* dweights, float* ddensity);

/* PDFSampler.generate_ray_samples (ray_samplers.py:276-372) incl. the anneal pow of ProposalNetworkSampler (:602).
 * u_base[S_out+1]: linspace(0, 1-1/n, n) (+1/(2n) already added by the caller in eval); jitter[B] or NULL.
 * anneal_dev (nullable): device float that overrides `anneal` — the trainer's CUDA graph reads the schedule value
 * (NS/models/nerfacto.py:256-278) from it on every replay.
 * inds (nullable) int32 [B,S_out+1]: raw searchsorted(cdf,u,right) result (test hook; bit-exact contract). */
int nvo_pdf_resample(void* stream, int64_t B, int32_t S_in, int32_t S_out, const float* weights, const float* sdist_in, const float* u_base,
                     const float* jitter, float anneal, const float* anneal_dev, float histogram_padding, const float* nears, const float* fars,
                     float* sdist_out, float* ebins_out, int32_t* inds);

/* RGBRenderer('last_sample') + AccumulationRenderer + DepthRenderer(expected|median) + NormalsRenderer/NormalsShader
 * (NS/model_components/renderers.py:70-117,199-230,287-315,333-381,427-447; shaders.py:56-77).
 * Inputs nullable where noted; every output nullable.  eval_mode!=0 applies nan_to_num+clamp to rgb (:223-229).
 * depth_expected is written UNCLIPPED; minmax[2] (initialised by the call) receives min/max of the mid-steps
 * over the whole batch, nvo_clip_depth applies renderers.py:379 afterwards. */
int nvo_render_forward(void* stream, int64_t B, int32_t S, int32_t eval_mode, const float* starts, const float* ends, int64_t stride,
                       const float* weights, const float* rgb /*[B,S,3]*/,
                       const float* normals /*nullable*/, const float* pred_normals /*nullable*/, float* out_rgb /*[B,3]*/, float* out_acc /*[B]*/,
                       float* out_depth_expected /*[B]*/, float* minmax /*[2]*/, float* out_depth_median /*[B]*/, int32_t* out_median_idx /*[B]*/,
                       float* out_normals /*[B,3]*/, float* out_pred_normals /*[B,3]*/);
int nvo_clip_depth(void* stream, int64_t B, const float* minmax, float* depth);
/* backward of nvo_render_forward (training mode). Any d_out_* may be NULL (= zero). dweights [B,S] is OVERWRITTEN
 * unless accumulate_dweights!=0; drgb [B,S,3] and dpred_normals [B,S,3] (nullable) are overwritten. */
int nvo_render_backward(void* stream, int64_t B, int32_t S, const float* starts, const float* ends, int64_t stride, const float* weights,
                        const float* rgb, const float* normals,
                        const float* pred_normals, const float* d_out_rgb, const float* d_out_acc, const float* d_out_depth_expected,
                        const float* minmax, const float* d_out_normals, const float* d_out_pred_normals, int32_t accumulate_dweights,
                        float* dweights, float* drgb, float* dpred_normals);

/* ---------------------------------------------------------------------------------------------
 * Losses (NS/model_components/losses.py). Each *_forward adds its per-batch MEAN (already divided by the element
 * count the reference averages over) into loss[0] (accumulate, fp32); each *_backward ADDS scale * (*dscale) * dL/dweights
 * into dweights: scale = host-side multiplier, dscale = nullable DEVICE scalar (the upstream autograd gradient), so
 * no host synchronisation is needed to chain losses.
 * ------------------------------------------------------------------------------------------- */
/* distortion_loss / lossfun_distortion (losses.py:134-153) on weights [B,S], sdist [B,S+1] */
int nvo_distortion_loss_forward(void* stream, int64_t B, int32_t S, const float* weights, const float* sdist, float* loss);
int nvo_distortion_loss_backward(void* stream, int64_t B, int32_t S, const float* weights, const float* sdist, const float* dscale, float scale,
                                 float* dweights);
/* interlevel_loss for ONE proposal level (losses.py:52-130): c/w = final level (S), cp/wp = proposal level (Sp).
 * idx_lo/idx_hi (nullable) int32 [B,S]: clamped searchsorted results (test hook). Gradient flows to wp only. */
int nvo_interlevel_loss_forward(void* stream, int64_t B, int32_t S, int32_t Sp, const float* w, const float* c, const float* wp, const float* cp,
                                float* loss, int32_t* idx_lo, int32_t* idx_hi);
int nvo_interlevel_loss_backward(void* stream, int64_t B, int32_t S, int32_t Sp, const float* w, const float* c, const float* wp, const float* cp,
                                 const float* dscale, float scale, float* dwp);
/* ds_nerf_depth_loss through depth_loss(is_euclidean=False) (losses.py:224-246,288-324): depth_gt[B], directions_norm[B] */
int nvo_depth_loss_forward(void* stream, int64_t B, int32_t S, const float* weights, const float* starts, const float* ends, int64_t stride,
                           const float* depth_gt,
                           const float* directions_norm, float sigma, float* loss);
int nvo_depth_loss_backward(void* stream, int64_t B, int32_t S, const float* weights, const float* starts, const float* ends, int64_t stride,
                            const float* depth_gt,
                            const float* directions_norm, float sigma, const float* dscale, float scale, float* dweights);
/* MSELoss on [B,3] (NS/models/nerfacto.py:362) and monosdf_normal_loss (losses.py:327-342); d_pred is overwritten with scale*grad */
int nvo_mse_loss(void* stream, int64_t n, const float* pred, const float* target, float scale, float* loss, float* d_pred);
int nvo_normal_loss(void* stream, int64_t B, const float* pred /*[B,3]*/, const float* gt /*[B,3]*/, float scale, float* loss, float* d_pred);
/* Every loss of the NeRF-VO mapping step (nerf_vo/mapping/nerfstudio.py:71-82; NS/models/nerfacto.py:354-381, depth_nerfacto.py:79-125,
 * nerf_vo/mapping/nerfstudio_utils.py:333-350) and its gradient for grad_output == 1 in ONE launch: w_l [B,S_l] weights of proposal level 0,
 * proposal level 1 and the field (l = 2), sdist_l [B,S_l+1] normalised bin edges, starts_l / ends_l Euclidean interval ends (row stride_l;
 * read only with depth_gt), rgb / rgb_gt [B,3], normals / normal_gt [B,3] (nullable), depth_gt / directions_norm [B] (nullable).
 * terms[5] (zeroed by the caller) receives the unweighted batch means [rgb, interlevel, distortion, depth summed over the levels, normal],
 * total[1] = sum mult_i terms_i (written by the last CTA; ticket: one zero-initialised uint32 the kernel leaves zero), dw_l [B,S_l],
 * d_rgb [B,3], d_normals [B,3] (nullable) are WRITTEN with mult * d term / d input. */
int nvo_step_losses(void* stream, int64_t B, int32_t S0, int32_t S1, int32_t S2, const float* w0, const float* w1, const float* w2,
                    const float* sdist0, const float* sdist1, const float* sdist2, const float* starts0, const float* ends0, int64_t stride0,
                    const float* starts1, const float* ends1, int64_t stride1, const float* starts2, const float* ends2, int64_t stride2,
                    const float* rgb, const float* rgb_gt, const float* normals, const float* normal_gt, const float* depth_gt,
                    const float* directions_norm, float sigma, float mult_rgb, float mult_interlevel, float mult_distortion, float mult_depth,
                    float mult_normal, float* terms, float* total, void* ticket, float* dw0, float* dw1, float* dw2, float* d_rgb,
                    float* d_normals);

/* ---------------------------------------------------------------------------------------------
 * Step prologue: pixel sampling + pixel gather + ray generation + camera-pose correction in ONE launch —
 * replaces DynamicDataManager.next_train (nerf_vo/mapping/nerfstudio_utils.py:295-300): DynamicDataset.get_dataset (:133-155,
 * which re-solves R^-1 n over EVERY frame each step), PixelSampler.collate_image_dataset_batch (NS/data/pixel_samplers.py:103-106,
 * 170-219, including its c/y/x `.cpu()` sync), RayGenerator.forward (NS/model_components/ray_generators.py:40-57),
 * Cameras.generate_rays perspective branch (NS/cameras/cameras.py:596-654,780-785,865-912) and
 * CameraOptimizer.apply_to_raybundle (NS/cameras/camera_optimizers.py:108-147).
 *   u[B,3]            uniform [0,1) draws (torch.rand); indices = trunc(u * [K,H,W]) in fp32, bit-exact
 *   K / K_dev         number of active keyframes; K_dev (nullable) = device int32 read by the kernel instead of K (clamped to [1,K]):
 *                     a captured CUDA graph then follows the keyframes the mapping thread inserts (nerfstudio_utils.py:203-241)
 *   intrinsics[*,4]   fx fy cx cy per frame; extrinsics[*,4,4] camera-to-world (row-major; only rows 0..2 are read)
 *   frames_color[K,H,W,3], frames_depth[K,H,W,1], frames_normal[K,H,W,3] (camera-frame normals; NULL = no normal target)
 *   pose_adjustment[*,6] (translation, rotation tangent) or NULL; pose_mode NVO_POSE_OFF / SO3XR3 / SE3
 * outputs: indices[B,3] int64 (camera,row,col), camera_indices[B,1] int64, origins/directions[B,3] (pose-corrected),
 *   directions_norm[B,1], pixel_area[B,1], rgb[B,3], depth[B,1], normal[B,3] = (R^-1 n + 1)/2 (nullable with frames_normal),
 *   directions_raw[B,3] (nullable): the uncorrected unit directions, the saved input of nvo_pose_correction_backward.
 * ------------------------------------------------------------------------------------------- */
enum { NVO_POSE_OFF = 0, NVO_POSE_SO3XR3 = 1, NVO_POSE_SE3 = 2 };
int nvo_batch_prologue(void* stream, int64_t B, int32_t K, const int32_t* K_dev, int32_t H, int32_t W, const float* u, const float* intrinsics,
                       const float* extrinsics,
                       const float* frames_color, const float* frames_depth, const float* frames_normal, const float* pose_adjustment,
                       int32_t pose_mode, int64_t* indices, int64_t* camera_indices, float* origins, float* directions, float* directions_norm,
                       float* pixel_area, float* rgb, float* depth, float* normal, float* directions_raw);
/* Cameras.generate_rays for given pixels (indices[n,3] int64) or, with indices == NULL, for every pixel of frame `cam` in row-major
 * order (n = H*W; the evaluation bundle of Cameras.generate_rays(camera_indices=cam, keep_shape=True), evaluation/nerf_renderer.py:140-147).
 * No pose correction (eval mode, NS/models/nerfacto.py:290). */
int nvo_generate_rays(void* stream, int64_t n, int32_t cam, int32_t W, const int64_t* indices, const float* intrinsics, const float* extrinsics,
                      int64_t* camera_indices, float* origins, float* directions, float* directions_norm, float* pixel_area);
/* exp_map_SO3xR3 / exp_map_SE3 (NS/cameras/lie_groups.py:25-60,63-120): tangent[n,6] -> matrices[n,3,4] (CameraOptimizer.forward) */
int nvo_pose_exp_map(void* stream, int64_t n, int32_t pose_mode, const float* tangent, float* matrices);
/* CameraOptimizer.apply_to_raybundle (NS/cameras/camera_optimizers.py:142-147): origins_out = origins + t(camera), directions_out =
 * R(camera) directions, with [R|t] = exp map of pose_adjustment[camera_indices[i]]. */
int nvo_pose_apply(void* stream, int64_t B, int32_t pose_mode, const int64_t* camera_indices, const float* pose_adjustment, const float* origins,
                   const float* directions, float* origins_out, float* directions_out);
/* backward of the pose correction: d_pose[K,6] += d(origins, directions)/d(pose_adjustment) applied to (d_origins, d_directions)[B,3].
 * scratch[K,12] must be zero on entry (per-camera cotangent of [R|t], accumulated by reductions). */
int nvo_pose_correction_backward(void* stream, int64_t B, int32_t K, int32_t pose_mode, const int64_t* camera_indices, const float* directions_raw,
                                 const float* d_origins, const float* d_directions, const float* pose_adjustment, float* scratch, float* d_pose);

/* ---------------------------------------------------------------------------------------------
 * Evaluation frame output — replaces the host-side numpy of NerfstudioRenderer.render_frame
 * (evaluation/nerf_renderer.py:160-167) and NeRFRenderer-driven passes of evaluation/renderer.py:79-97,113-121.
 * ------------------------------------------------------------------------------------------- */
/* color[n,3] uint8 = (uint8)(rgb*255) (numpy astype: truncation); depth_out[n] = depth[n] / directions_norm[n] (directions_norm
 * NULL = plain copy, the non-depth-supervised branch); depth16[n] uint16 = (uint16)((depth_out*scale_a)*scale_b) in fp32, nullable. */
int nvo_frame_finalize(void* stream, int64_t n, const float* rgb, const float* depth, const float* directions_norm, float scale_a, float scale_b,
                       void* color, float* depth_out, void* depth16);
/* sums[3] (double, accumulate; caller zero-fills) += { sum depth_gt, sum depth_pred, count } over pixels with
 * 0 < depth_gt < 5 and 0 < depth_pred < 5 (evaluation/renderer.py:88-93: the per-keyframe depth-scale alignment). */
int nvo_depth_scale_sums(void* stream, int64_t n, const float* depth_gt, const float* depth_pred, void* sums);
/* Point-cloud export for Poisson meshing (evaluation/nerf_renderer.py:170-209 calling NS/exporter/exporter_utils.py:78-231, lines 130-180 and
 * 222-226): points[n,3] = origins + directions * depth; keep[n] (uint8) = accumulation > 0.5 and, when box_min / box_max (HOST float[3],
 * both or neither) are given, box_min < point < box_max on every axis; normals[n,3] = normals_coded * 2 - 1 (nullable pair), negated where
 * reorient != 0 and dot(direction, normal) > 0.  The caller compacts by `keep`. */
int nvo_point_cloud(void* stream, int64_t n, const float* origins, const float* directions, const float* depth, const float* accumulation,
                    const float* normals_coded, const float* box_min, const float* box_max, int32_t reorient, float* points, float* normals,
                    void* keep);

/* ---------------------------------------------------------------------------------------------
 * The nerfacto field's three networks as ONE persistent tcgen05 kernel per direction (csrc/field_tc.cu): NerfactoField.get_density's
 * mlp_base + get_outputs' mlp_head / mlp_pred_normals + PredNormalsFieldHead + Field.get_normals (NS/fields/nerfacto_field.py:199-297,
 * NS/fields/base_field.py:80-133, NS/field_components/field_heads.py:189-204) — the tensor-core twin of tiny-cuda-nn's FullyFusedMLP
 * chain (TCNN/src/fully_fused_mlp.cu:499-557) for NeRF-VO's fixed architecture: base 32->64->16, head 63->64->64->3 (sigmoid),
 * pred-normals 27->64->64->64->3 (tanh, normalize); fp16 operands, fp32 accumulation.
 *   nvo_field_pack_weights: torch-layout fp32 parameters of the three networks ([W0, b0, W1, b1, ...]; pn_params = the three
 *     mlp_pred_normals layers followed by PredNormalsFieldHead's Linear, nullable) -> the kernels' fp16 weight image
 *     (nvo_field_wimage_bytes() bytes); `grid` supplies the 16 level scales folded into the normals chain.
 *   nvo_field_forward: feat16 = the main grid's TMH feature tiles (nvo_grid_forward / _jac), jac = its saved derivatives (nullable with
 *     `normals`), positions[n,3] world sample positions (position encoding), directions[B,3], cam_idx[B] int64 or NULL (then
 *     `embedding` is one 32-vector: eval), selector[n].  Outputs: density[n], rgb[n,3], pred_normals[n,3] (nullable: network skipped),
 *     normals[n,3] (nullable), h0[n] raw density and pn_raw[n,3] (nullable; saved for the backward), saved = nvo_field_saved_bytes(n,
 *     save_pn) bytes of fp16 activations for the backward (nullable: inference); save_pn = 0 leaves the pred-normals activations out
 *     (512 instead of 960 B per sample): NeRF-VO trains with pred_normal_loss_mult = 0, that network receives no gradient.
 * ------------------------------------------------------------------------------------------- */
int64_t nvo_field_wimage_bytes(void);
int64_t nvo_field_saved_bytes(int64_t n, int32_t save_pn);
int nvo_field_pack_weights(void* stream, const nvo_grid_desc* grid, const float* base_params, const float* head_params, const float* pn_params,
                           void* wimage);
/* nvo_field_backward: d loss / d(density[n], rgb[n,3]) (+ dpn_in: fp32 TMF [tiles][27][128], the gradient w.r.t. the pred-normals network's
 * input from nvo_mlp_tc_backward_strided, nullable) -> dfeat: fp32 TMF [tiles][32][128] gradient w.r.t. the hash features (input of
 * nvo_grid_backward), and ACCUMULATES the parameter gradients: dbase_params / dhead_params (flat, torch layout), dembedding ([K,32] rows by
 * cam_idx, or one 32-vector when cam_idx is NULL; nullable).  rgb / h0 / saved / feat16: the forward's outputs; scratch: two device floats.
 * ddirections[B,3] (nullable, with directions[B,3]): accumulates d loss / d ray direction through the SH encoding (camera-pose optimisation).
 * Needs S >= 32 (a warp's 32 rows then span at most two rays: the embedding gradient is reduced per ray before its atomics). */
int nvo_field_backward(void* stream, int64_t B, int32_t S, const void* feat16, const void* saved, int32_t save_pn, const void* wimage,
                       const float* rgb, const float* h0, const float* selector, const int64_t* cam_idx, const float* ddensity, const float* drgb,
                       const float* dpn_in, float* scratch, float* dfeat, float* dbase_params, float* dhead_params, float* dembedding,
                       const float* directions, float* ddirections);
int nvo_field_forward(void* stream, int64_t B, int32_t S, const void* feat16, const void* jac, const float* positions, const float* directions,
                      const int64_t* cam_idx, const float* embedding, const float* selector, const void* wimage, float* density, float* rgb,
                      float* pred_normals, float* normals, float* h0, float* pn_raw, void* saved, int32_t save_pn);

/* d loss / d (ray origins, ray directions)[B,3] (accumulated with atomics; either nullable) from dx[B*S,3] = d loss / d x, x = the normalised,
 * selector-masked contracted sample positions the hash grids read (NS/cameras/rays.py:49-58, spatial_distortions.py:67-69,
 * nerfacto_field.py:204-209): the path the reference's autograd takes from the field back to CameraOptimizer.apply_to_raybundle. */
int nvo_position_backward(void* stream, int64_t B, int32_t S, const float* origins, const float* directions, const float* starts,
                          const float* ends, int64_t stride, const float* dx, float* d_origins, float* d_directions);
/* CameraOptimizer.get_loss_dict (NS/cameras/camera_optimizers.py:149-155): loss += mean_k ||t_k|| trans_l2_penalty + mean_k ||r_k|| rot_l2_penalty;
 * d_pose[K,6] (nullable) += scale * its gradient (zero at a zero vector, like torch's norm backward). */
int nvo_pose_regularizer(void* stream, int32_t K, const float* pose_adjustment, float trans_l2_penalty, float rot_l2_penalty, float scale,
                         float* loss, float* d_pose);
/* nvo_adam_step with ExponentialDecayScheduler's learning rate (no warm-up; NS/engine/schedulers.py:122-138) evaluated on the device from
 * the step counter: lr = exp(log(lr_init) (1 - t) + log(lr_final) t), t = min(step / max_steps, 1) — the "camera_opt" group. */
int nvo_adam_step_decay(void* stream, int64_t n, float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int32_t* step,
                        double lr_init, double lr_final, int32_t max_steps, double beta1, double beta2, double eps, float grad_scale);

/* ---------------------------------------------------------------------------------------------
 * Fused dense Adam over a flat fp32 buffer (torch.optim.Adam semantics; NS/engine/optimizers.py:138-150,
 * nerf_vo/mapping/nerfstudio.py:84-100). step[1] is a DEVICE int32 counter (number of steps taken so far), read for the
 * bias correction and incremented by the call, so the launch is CUDA-graph replayable. grad_scale multiplies the
 * gradient first (1/world_size after a sum-allreduce).  lr / betas / eps are doubles: the scalar constants (1 - beta, bias
 * corrections, step size) are evaluated in double precision as torch's python floats are, then rounded to fp32 once.
 * ------------------------------------------------------------------------------------------- */
int nvo_adam_step(void* stream, int64_t n, float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int32_t* step, double lr,
                  double beta1, double beta2, double eps, float grad_scale);

/* ---------------------------------------------------------------------------------------------
 * Data-parallel exchange fused with the optimizer over NVLink peer memory (csrc/exchange.cu) — replaces the
 * DistributedDataParallel gradient all-reduce + dense Adam the reference would run for world_size > 1
 * (NS/pipelines/base_pipeline.py:281-283, NS/engine/optimizers.py:138-150).
 * One launch per rank and step: barrier(gradients ready) -> reduce-scatter by peer loads -> Adam on the rank's own
 * slice (moments exist only there) -> all-gather by peer stores -> barrier(replicas written).
 * h_peer_params / h_peer_grads / h_peer_flags: HOST arrays of `world` device pointers (entry k = rank k's buffer as
 * mapped into THIS process; entry `rank` = the local allocation). Flag pads are nvo_exchange_flag_words() int32, zeroed.
 * n (floats, multiple of 4) is the flat buffer length; exp_avg_slice / exp_avg_sq_slice hold nvo_exchange_slice() floats.
 * step[1]: device int32 step counter, identical on every rank, incremented by the call (epoch of the barriers).
 * Spin waits are bounded; on timeout flag word 33 of the local pad becomes non-zero (1: ready barrier, 2: done barrier).
 * ------------------------------------------------------------------------------------------- */
int nvo_exchange_flag_words(void);
int64_t nvo_exchange_slice(int64_t n, int32_t rank, int32_t world, int64_t* lo, int64_t* hi);
int nvo_adam_exchange_step(void* stream, int64_t n, int32_t rank, int32_t world, const void* h_peer_params, const void* h_peer_grads,
                           const void* h_peer_flags, float* exp_avg_slice, float* exp_avg_sq_slice, int32_t* step, double lr, double beta1,
                           double beta2, double eps, float grad_scale);
/* The same exchange restricted to the flat range [offset, offset + n) — one optimizer parameter group (NS/engine/optimizers.py:138-150
 * keeps one Adam per group; nerfacto's groups are "fields" and "proposal_networks", NS/models/nerfacto.py:244-249).  `phase`
 * (0 .. 2) selects the group's own block of flags, so the groups' exchanges of one step may be in flight concurrently; moments
 * hold nvo_exchange_slice(n, ...) floats; `step` is the group's own counter (a group that received no gradient is not stepped).
 * ctas_per_sm > 0 caps the persistent grid (0 = as many CTAs as fit): a launch that overlaps other kernels leaves them the registers. */
int nvo_adam_exchange_group(void* stream, int64_t offset, int64_t n, int32_t phase, int32_t rank, int32_t world, const void* h_peer_params,
                            const void* h_peer_grads, const void* h_peer_flags, float* exp_avg_slice, float* exp_avg_sq_slice, int32_t* step,
                            double lr, double beta1, double beta2, double eps, float grad_scale, int32_t ctas_per_sm);
/* Two parameter groups stepped in the same launch (one pair of barriers instead of two; the flags of phase 0 and group A's counter as
 * barrier epoch): what the trainer uses when both groups' gradients are complete at the same time. */
/* nvo_adam_exchange_group under ExponentialDecayScheduler (lr_init -> lr_final over max_steps, evaluated on the device from the group's step
 * counter): the "camera_opt" group of a data-parallel run. */
int nvo_adam_exchange_group_decay(void* stream, int64_t offset, int64_t n, int32_t phase, int32_t rank, int32_t world, const void* h_peer_params,
                                  const void* h_peer_grads, const void* h_peer_flags, float* exp_avg_slice, float* exp_avg_sq_slice, int32_t* step,
                                  double lr_init, double lr_final, int32_t max_steps, double beta1, double beta2, double eps, float grad_scale,
                                  int32_t ctas_per_sm);
int nvo_adam_exchange_groups2(void* stream, int64_t offset_a, int64_t n_a, float* exp_avg_a, float* exp_avg_sq_a, int32_t* step_a, int64_t offset_b,
                              int64_t n_b, float* exp_avg_b, float* exp_avg_sq_b, int32_t* step_b, int32_t rank, int32_t world,
                              const void* h_peer_params, const void* h_peer_grads, const void* h_peer_flags, double lr, double beta1, double beta2,
                              double eps, float grad_scale);
/* peer-visible allocations: cudaMalloc (zero-filled) + CUDA IPC export / import; handles are 64 opaque bytes */
int nvo_peer_alloc(int64_t bytes, void* h_ptr_out, void* h_handle64_out);
int nvo_peer_open(const void* h_handle64, void* h_ptr_out);
int nvo_peer_close(void* ptr);
int nvo_peer_free(void* ptr);

#ifdef __cplusplus
}
#endif
#endif /* NVO_B200_H */
