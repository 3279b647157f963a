/*
 * b2s.h - C ABI of libb2s: B200-native (sm_100a) kernels for the DiffSinger acoustic / variance
 *         sampling hot path (reference: vsingerxiaoice-rwkv/xiaoicesing-io, an OpenVPI DiffSinger fork).
 *
 * The reference has NO native code and NO FFI for this path: its boundary is
 * Python (backbone registry modules/backbones/__init__.py:6-18, GaussianDiffusion / RectifiedFlow in
 * modules/core/ddpm.py:55-383 and reflow.py:13-144).  This header is therefore the NEW seam that sits
 * directly under that Python surface; every entry point names the reference code it replaces.
 * The Python binding a maintainer adds is a ctypes stub (xiaoicesing_io_b200/_cabi.py).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the name says host;
 *   - the caller owns every buffer; no entry point allocates, synchronises or keeps state, so all of them
 *     are CUDA-graph capturable on the stream passed as the last argument (a cudaStream_t as void*);
 *   - activations are TIME-MAJOR: a "frame matrix" is [rows = B*T, channels] row-major, row r = b*T + t.
 *     (The reference keeps [B, channels, T]; its final op is a transpose to [B, T, M], ddpm.py:350.)
 *   - return value 0 = success, negative = error (b2s_last_error() gives the text, thread-local).
 */
#ifndef B2S_H_
#define B2S_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2S_ABI_VERSION 5

#define B2S_OK 0
#define B2S_ERR_INVALID_ARGUMENT (-1)
#define B2S_ERR_CUDA (-2)
#define B2S_ERR_UNSUPPORTED (-3)

/* activation codes for b2s_linear_f32 */
#define B2S_ACT_NONE 0
#define B2S_ACT_RELU 1 /* wavenet.py:88,98 */
#define B2S_ACT_MISH 2 /* wavenet.py:60 */
#define B2S_ACT_GELU 3 /* exact erf GELU, lynxnet.py:107,143 */
#define B2S_ACT_SILU 4
#define B2S_ACT_LRELU 5 /* leaky_relu(x, 0.1), nsf_hifigan/models.py:15 */

int b2s_abi_version(void);
const char* b2s_last_error(void);

/* ------------------------------------------------------------------------------------------------
 * Layout helpers
 * ---------------------------------------------------------------------------------------------- */

/* in [batch, rows, cols] -> out [batch, cols, rows] (fp32).  Replaces the .transpose() calls at
 * ddpm.py:350,357,370 / reflow.py:44,58,137 when crossing between the reference layout [B, M, T]
 * and the time-major internal layout. */
int b2s_transpose_f32(const float* in, float* out, int batch, int rows, int cols, void* stream);

/* norm_spec / denorm_spec (ddpm.py:379-383, reflow.py:140-144) fused with the layout change between the caller's spec tensor
 * [B, T, M] (F = 1) or [B, F, T, M] and the sampler's time-major state [B*T, F*M] (the reference transposes to [B, F, M, T] at
 * ddpm.py:370-373 and back at :350).  spec_min / spec_max: [F*M] fp32 (the registered buffers, flattened).
 *   norm:   state[(b*T + t), f*M + m] = (spec[b, f, t, m] - min) / (max - min) * 2 - 1
 *   denorm: spec[b, f, t, m] = (state[(b*T + t), f*M + m] + 1) / 2 * (max - min) + min */
int b2s_spec_norm_f32(const float* spec, const float* spec_min, const float* spec_max, float* state, int B, int F, int T, int M,
                      void* stream);
int b2s_spec_denorm_f32(const float* state, const float* spec_min, const float* spec_max, float* spec, int B, int F, int T, int M,
                        void* stream);

/* ------------------------------------------------------------------------------------------------
 * Sampler update: dst = sum_i coef[i] * src[i]   (one vectorised elementwise kernel)
 *
 * Covers p_sample (ddpm.py:149-156), p_sample_ddim (:158-167), p_sample_plms (:169-204), q_sample
 * (:206-210), the DPM-Solver++ 2M updates (dpm_solver_pytorch.py:547-580, 796-831) incl. the
 * eps -> x0 conversion (:434-442), the UniPC-bh2 predictor / corrector (uni_pc.py:548-567) and the
 * Euler / RK stages (reflow.py:66-102).  coef is a DEVICE array so a captured graph can be replayed.
 * n_src <= 8.  dst may alias any src.
 * ---------------------------------------------------------------------------------------------- */
int b2s_sampler_lincomb_f32(float* dst, const float* const* srcs_host, const float* coef, int n_src,
                            int64_t n, void* stream);

/* Same update, and in the same launch: out_h [n] = dst rounded to bf16 (bf16 != 0) or fp16 - the 16-bit A operand the
 * tensor-core denoiser reads next (ddpm.py:149-156 feeds x' straight back into denoise_fn, :141) - and flags[0..n_flags)
 * = 0, re-arming the tile flags of the persistent denoiser kernel (replaces b2s_cast_f32_h_reset after an update). */
int b2s_sampler_lincomb_f32_h(float* dst, const float* const* srcs_host, const float* coef, int n_src, int64_t n,
                              void* out_h, int bf16, int* flags, int n_flags, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Step embedding:  SinusoidalPosEmb (common_layers.py:266-278)
 * t [n] fp32 (the model time: int step index, (t-1/N)*N, or 1000*t) -> out [n, dim]
 * ---------------------------------------------------------------------------------------------- */
int b2s_sinusoid_f32(const float* t, float* out, int n, int dim, void* stream);

/* ------------------------------------------------------------------------------------------------
 * fp32 reference-precision path (CUDA cores): out[M,N] = act(alpha * A[M,K] . W[N,K]^T + bias[N])
 * Replaces nn.Linear / Conv1d(k=1): wavenet.py:34,35,58-62,86,97,99; lynxnet.py:71,72,104-109,154.
 * If y is non-null also writes y = out + dvec[b * d_stride + n]  (b = row / T), the next layer's
 * pre-added step embedding (wavenet.py:36).  K % 4 == 0, lda/ldw/ldo % 4 == 0.
 * ---------------------------------------------------------------------------------------------- */
int b2s_linear_f32(const float* A, int lda, const float* W, int ldw, const float* bias, float* out, int ldo,
                   int M, int N, int K, float alpha, int act, float* y, const float* dvec, int d_stride, int T,
                   void* stream);

/* WaveNet residual block, first half (wavenet.py:36-42): z = sigmoid(g) * tanh(f),
 * [g|f] = dilated_conv_k3(y) + cond, y = x + step embedding (zero outside [0,T) per utterance, H1).
 *   y      [B*T, C]        pre-added activations
 *   Wd     [2C, 3C]        packed: row 2j = gate j, 2j+1 = filter j; column = tap*C + c
 *   cond   [B*T, ld_cond]  hoisted conditioner projection of THIS layer (+ both biases), same interleave
 *   z      [B*T, C]
 * C % 16 == 0. */
int b2s_wavenet_gate_f32(const float* y, const float* Wd, const float* cond, int ld_cond, float* z, int B, int T,
                         int C, int dilation, void* stream);

/* WaveNet residual block, second half (wavenet.py:44-48): [r|s] = Wo z + bo;
 * x <- (x + r)/sqrt(2);  y_next <- x + dvec_next (if y_next != NULL);  skip <- s (first) or skip + s.
 *   Wo [2C, C] rows 0..C-1 residual, C..2C-1 skip (reference order). */
int b2s_wavenet_out_f32(const float* z, const float* Wo, const float* bo, float* x, float* y_next, float* skip,
                        const float* dvec_next, int d_stride, int first_layer, int B, int T, int C, void* stream);

/* LYNXNet (lynxnet.py:76-87, 52-62) fp32 building blocks ------------------------------------------ */

/* u = x + cond + d;  res = strong_cond ? x + cond : x (written back to x);  h = LayerNorm_C(u)*gamma+beta */
int b2s_lynx_prenorm_f32(float* x, const float* cond, int ld_cond, const float* dvec, int d_stride,
                         const float* gamma, const float* beta, float* h, int B, int T, int C, int strong_cond,
                         void* stream);
/* plain LayerNorm over channels (final norm, lynxnet.py:151) */
int b2s_layernorm_f32(const float* x, const float* gamma, const float* beta, float* h, int rows, int C,
                      void* stream);
/* g = out * silu(gate), [out|gate] = W h + b; W packed [2*inner, C] interleaved (row 2j = out j, 2j+1 = gate j) */
int b2s_lynx_glu_f32(const float* h, const float* W, const float* bias, float* g, int rows, int C, int inner,
                     void* stream);
/* depthwise conv along time (k taps, zero pad k/2 per utterance) + bias + activation
 * act: 0 = PReLU(slope[inner]), B2S_ACT_SILU, B2S_ACT_RELU.  Wdw [inner, k]. */
int b2s_lynx_dwconv_f32(const float* g, const float* Wdw, const float* bias, const float* slope, float* p, int B,
                        int T, int inner, int ksize, int act, void* stream);
/* x <- W p + b + x   (W [C, inner]) */
int b2s_linear_residual_f32(const float* p, const float* W, const float* bias, float* x, int rows, int C,
                            int inner, void* stream);

/* ------------------------------------------------------------------------------------------------
 * 16-bit tensor-core path (tcgen05.mma with TMEM accumulators, TMA-staged operands; b2s_tc_gemm.cu).
 * "_h" pointers are 16-bit arrays: bf16 when the trailing `bf16` flag is 1, IEEE fp16 when 0.
 * Accumulation, the residual stream x, the skip sum and every bias / step-embedding row stay fp32.
 * Operand requirements: 16B-aligned bases, leading dimensions that are multiples of 8 elements.
 * ---------------------------------------------------------------------------------------------- */

/* out = fp32 -> 16-bit cast (feeds the TMA-staged A operands; n elements) */
int b2s_cast_f32_h(const float* in, void* out_h, int64_t n, int bf16, void* stream);
/* same, and zeroes `flags[0..n_flags)` (the tile flags of b2s_tc_wavenet_stack / _denoiser) in the same launch */
int b2s_cast_f32_h_reset(const float* in, void* out_h, int64_t n, int* flags, int n_flags, int bf16, void* stream);

/* Same contract as b2s_linear_f32 with 16-bit A [rows, K] / W [N, K]:
 *   v = act(alpha * A.W^T + bias);  out_f32 (nullable) <- v;  out_h (nullable) <- v;
 *   y_h (nullable) <- v + dvec[(row / T) * d_stride + n]   (wavenet.py:36, next layer's pre-added embedding)
 * Replaces wavenet.py:35,86-88,97-99 / lynxnet.py:72,141-143,154 on the tensor cores. */
int b2s_tc_linear(const void* A_h, int lda, int rows, int T, const void* W_h, int ldw, const float* bias, int N, int K,
                  float alpha, int act, float* out_f32, int ldo, void* out_h, int ldoh, void* y_h, int ldy,
                  const float* dvec, int d_stride, int bf16, void* stream);

/* Hoisted conditioner projection of ALL layers as ONE tensor-core GEMM (wavenet.py:35 / lynxnet.py:77-82; precedent:
 * the reference's own ONNX exporter, utils/onnx_helper.py:231-314):  table[l][row][n] = cond[row,:] . Wc[l*N2 + n,:] + bc.
 * The table is LAYER-MAJOR [L][rows][N2] so that one layer's slab is contiguous in HBM (it is streamed once per
 * denoiser evaluation).  N2 = 2C (WaveNet, gate/filter interleaved) or C (LYNXNet); N2 % 32 == 0. */
int b2s_tc_cond_table(const void* cond_h, int rows, const void* Wc_h, const float* bc, int L, int N2, int H,
                      void* table_h, int bf16, void* stream);

/* Same GEMM, table written in the TILE/CHUNK-MAJOR layout the fused WaveNet kernels read with fully coalesced 512-byte
 * warp loads (thread = frame row):  element (layer l, utterance b, frame t, packed column n) lives at
 *   ((((l*B + b)*tpb + t/128) * (N2/32) + n/32) * 4 + (n%32)/8) * 1024 + (t%128)*8 + n%8        [16-bit elements]
 * with tpb = ceil(T/128) rounded up to even; size L*B*tpb*128*N2 elements (rows >= T of the last tile are never read). */
int b2s_tc_cond_table_tiled(const void* cond_h, int B, int T, const void* Wc_h, const float* bc, int L, int N2, int H,
                            void* table_h, int bf16, void* stream);

/* b2s_wavenet_gate_f32 on the tensor cores: implicit-GEMM dilated conv (3 TMA tiles per K slab at time
 * offsets -d, 0, +d; out-of-bounds zero fill = the per-utterance zero padding), epilogue adds the hoisted
 * 16-bit conditioner projection and applies sigmoid*tanh (wavenet.py:38-42).  C % 64 == 0. */
int b2s_tc_wavenet_gate(const void* y_h, const void* Wd_h, const void* cond_h, int ld_cond, void* z_h, int B, int T,
                        int C, int dilation, int bf16, void* stream);

/* b2s_wavenet_out_f32 on the tensor cores (wavenet.py:44-48).  x / skip fp32 in place; y_next_h 16-bit;
 * skip_h (nullable): 16-bit copy of the running skip sum (the head GEMM's A operand after the last layer). */
int b2s_tc_wavenet_out(const void* z_h, const void* Wo_h, const float* bo, float* x, void* y_next_h, float* skip,
                       void* skip_h, const float* dvec_next, int d_stride, int first_layer, int B, int T, int C,
                       int bf16, void* stream);

/* Deferred skip sum (channel widths without a fused stack kernel, e.g. C = 512 / 192): the L skip outputs are never
 * accumulated layer by layer.  Every layer's z_l is kept (b2s_tc_wavenet_gate into slab l of a [L][rows][C] buffer, or
 * b2s_tc_wavenet_gate_ld into column block l of a [rows, ld_z = L*C] one), b2s_tc_wavenet_res applies only the residual rows of
 * output_projection (wavenet.py:44-48: x <- (x + W_res z + b)/sqrt(2), y_next = x + d_next), and after the last layer ONE GEMM
 * with K = L*C (b2s_tc_skip_sum, or b2s_tc_linear on the [rows, L*C] layout) computes sum_l W_skip,l z_l (wavenet.py:96) - the
 * same FLOPs, but no fp32 skip read-modify-write per layer and half the columns in every per-layer GEMM. */
int b2s_tc_wavenet_gate_ld(const void* y_h, const void* Wd_h, const void* cond_h, int ld_cond, void* z_h, int ld_z, int B, int T,
                           int C, int dilation, int bf16, void* stream);
int b2s_tc_wavenet_res(const void* z_h, int ld_z, const void* Wres_h, const float* b_res, float* x, void* y_next_h,
                       const float* dvec_next, int d_stride, int B, int T, int C, int bf16, void* stream);
/* The deferred skip GEMM over a LAYER-MAJOR z buffer [L][rows][C] (every layer's z rows stay contiguous, as the gate GEMM
 * writes them): out_h [rows, C] = sum_l z_l W_skip,l^T + bias, Wcat_h [C, L*C] with column l*C + k = W_skip,l[:, k]; the K loop
 * walks a 3-D tensor map {C, rows, L}. */
int b2s_tc_skip_sum(const void* z_all_h, const void* Wcat_h, const float* bias, void* out_h, int rows, int C, int L, int bf16,
                    void* stream);

/* ONE fused kernel per WaveNet residual layer (wavenet.py:33-48), residual channels C = 256:
 * dilated conv (implicit GEMM) -> + hoisted cond -> sigmoid*tanh -> z kept in shared memory -> output projection
 * -> x <- (x + r)/sqrt2, y_next <- x + dvec_next, skip (+)= s.  y_next_h must be a different buffer than y_h
 * (neighbouring tiles still read the halo of y_h).  Returns B2S_ERR_UNSUPPORTED for other C: call
 * b2s_tc_wavenet_gate + b2s_tc_wavenet_out instead.  Launched with programmatic dependent launch so that
 * consecutive layers overlap prologue and tail. */
int b2s_tc_wavenet_layer(const void* y_h, const void* Wd_h, const void* cond_h, int ld_cond, const void* Wo_h,
                         const float* bo, float* x, void* y_next_h, float* skip, void* skip_h, const float* dvec_next,
                         int d_stride, int first_layer, int B, int T, int C, int dilation, int bf16, void* stream);

/* The WHOLE residual stack (all L layers, wavenet.py:92-94) as ONE persistent kernel, C = 256: one CTA per 128-frame
 * tile, every tile resident at once (B * 2*ceil(ceil(T/128)/2) <= b2s_tc_wavenet_stack_max_tiles(), else
 * B2S_ERR_UNSUPPORTED: split the batch by utterance).  Layer-to-layer hand-off of y between neighbouring tiles goes
 * through release/acquire flags in `flags` (int32 [B * tiles], must be ZERO at launch) instead of kernel boundaries.
 *   y0_h      layer 0's input y = x + d_0 (written by the stem); y1_h the ping-pong partner
 *   Wd_h [L,2C,3C], Wo_h [L,2C,C], bo [L,2C]; cond_h: the table of b2s_tc_cond_table_tiled for THIS call's B and T,
 *   cond_layer_stride = B*tpb*128*2C elements (ld_cond is ignored)
 *   dvec: this evaluation's step-embedding row, layer l at dvec + b*d_stride + l*C;  dilations_host: L ints (HOST)
 * b2s_tc_wavenet_stack_max_tiles(): CTAs of this kernel the CURRENT device can hold at once = 2 x cudaOccupancyMaxActiveClusters for
 * its launch configuration (MPS / MIG / green-context limits included), not the SM count. */
int b2s_tc_wavenet_stack_max_tiles(void);
int b2s_has_experiments(void);   /* 1: built with B2S_BUILD_EXPERIMENTS=1 */
int b2s_tc_wavenet_stack(void* y0_h, void* y1_h, const void* Wd_h, const void* cond_h, int ld_cond,
                         int64_t cond_layer_stride, const void* Wo_h, const float* bo, float* x, float* skip, void* skip_h,
                         const float* dvec, int d_stride, const int* dilations_host, int L, int B, int T, int C, int* flags,
                         int bf16, void* stream);

#ifdef B2S_EXPERIMENTS   /* measured-and-rejected variants; exported only by B2S_BUILD_EXPERIMENTS=1 builds */
/* The residual stack with the GEMMs TRANSPOSED (channels on the tensor-core M axis, frames on N), C = 256: the frame tile
 * NT is 32, 48, 64 or 80 instead of 128, so a small batch still fills the SMs (16 x 690 frames = 144 tiles of 80); the fp32
 * residual stream lives in registers and the skip sum in TMEM for all L layers (wavenet.py:33-48, 92-96).
 *   tiles      b2s_tc_wavenet_stack_t_tiles(B, T, NT) = B * ceil(T / NT) rounded up to even; must be <= ..._max_tiles()
 *   Wd_h       [L, 2C, 3C] rows in BLOCK-PLANAR order: row 256h + 128g + c = (g ? filter : gate) of channel 128h + c
 *   cond_t     b2s_tc_cond_retile() of the layer-major table [L, B*T, 2C] computed with the same row order
 *   Wo_h, bo   [L, 2C, C], [L, 2C] in the reference's order (residual | skip)
 *   x          [B*T, C] fp32 stem output (read once); y0_h = x + d_0 (16-bit) on entry, y1_h its ping-pong partner
 *   skip_h     out: sum over layers of the skip outputs (+ biases), 16-bit [B*T, C]  (the head divides by sqrt(L))
 *   flags      int32 [tiles], ZERO at launch;  dilations (HOST) <= min(16, NT) */
int b2s_tc_wavenet_stack_t_tiles(int B, int T, int NT);
int b2s_tc_cond_retile(const void* table_h, int L, int B, int T, int n2, int NT, void* out_h, void* stream);
int b2s_tc_wavenet_stack_t(void* y0_h, void* y1_h, const void* Wd_h, const void* cond_t, const void* Wo_h, const float* bo,
                           const float* x, void* skip_h, const float* dvec, int d_stride, const int* dilations_host, int L,
                           int B, int T, int C, int NT, int* flags, int bf16, void* stream);

#endif /* B2S_EXPERIMENTS */

/* ONE launch per denoiser evaluation (wavenet.py:75-107 without the step-embedding MLP, which is hoisted into the step
 * table): b2s_tc_wavenet_stack plus, inside the same persistent kernel, the stem x = relu(W_in x_in + b_in), y_0 = x + d_0
 * (wavenet.py:86-88, :36) before the first layer and the head out = W_fin relu(W_sp skip/sqrt(L) + b_sp) + b_fin
 * (wavenet.py:96-99) after the last one (the skip sum and the hidden tile go through shared memory, never through HBM).
 *   xin_h [B*T, MF] 16-bit sampler state; Win_h [C, ld_win >= MF]; Wsp_h [C, C]; Wfin_h [MF, C]; out [B*T, MF] fp32.
 * MF = in_dims * n_feats, a multiple of 8, <= 256.  Same residency rule as b2s_tc_wavenet_stack. */
int b2s_tc_wavenet_denoiser(const void* xin_h, int MF, const void* Win_h, int ld_win, const float* b_in, void* y0_h, void* y1_h,
                            const void* Wd_h, const void* cond_h, int64_t cond_layer_stride, const void* Wo_h, const float* bo,
                            float* x, float* skip, const float* dvec, int d_stride, const int* dilations_host, int L,
                            const void* Wsp_h, const float* b_sp, const void* Wfin_h, const float* b_fin, float* out, int B, int T,
                            int C, int* flags, int bf16, void* stream);

#ifdef B2S_EXPERIMENTS
/* b2s_tc_wavenet_denoiser + the sampler update that consumes its output, in the same launch (ancestral / DDIM-type steps,
 * ddpm.py:149-167): x' = sum_i coef[i] * src_i for n_terms <= 3 terms in the order given, where srcs_host[i] == NULL stands
 * for THIS evaluation's output (eps_hat) and the other sources are fp32 [B*T, MF] buffers (the state x, the step's noise);
 * same fma order as b2s_sampler_lincomb_f32, so results are bit-identical to evaluation + update as two launches.
 * x_out (fp32, may alias a source) and x_out_h (16-bit: the next evaluation's xin_h; may alias this launch's xin_h - a tile
 * only overwrites the rows it alone has read) receive x'; the evaluation's output itself is not stored.
 * flags_next: the tile flags of the NEXT launch (or NULL): zeroed here, so consecutive launches alternate two flag buffers
 * and need no reset kernel in between. */
int b2s_tc_wavenet_denoiser_update(const void* xin_h, int MF, const void* Win_h, int ld_win, const float* b_in, void* y0_h,
                                   void* y1_h, const void* Wd_h, const void* cond_h, int64_t cond_layer_stride, const void* Wo_h,
                                   const float* bo, float* x, float* skip, const float* dvec, int d_stride,
                                   const int* dilations_host, int L, const void* Wsp_h, const float* b_sp, const void* Wfin_h,
                                   const float* b_fin, int B, int T, int C, int* flags, int* flags_next, int n_terms,
                                   const float* const* srcs_host, const float* coef, float* x_out, void* x_out_h, int bf16,
                                   void* stream);

#endif /* B2S_EXPERIMENTS */

/* Third design of the whole-stack kernel (round 2; csrc/b2s_tc_wavenet3.cu), C = 256, dilations <= b2s_tc_wavenet_stack3_halo():
 * stem + all L residual layers in ONE persistent launch of cta_group::2 pairs (two neighbouring 128-frame tiles form one 256-row
 * MMA).  The layer input y stays RESIDENT in shared memory (only 16 edge rows per side travel through yedge0_h / yedge1_h and the
 * tile flags), the fp32 residual stream stays in TMEM (wavenet.py:47-48 folded into the accumulator: Wres_h carries 2^(l/2)), and
 * the skip half of output_projection is deferred: every layer's gated tile z_l is TMA-stored to z_all_h [L][rows][C]
 * (z_layer_stride elements between layers) for ONE b2s_tc_skip_sum afterwards (wavenet.py:96).
 *   xin_h [B*T, MF] 16-bit; Win_h [C, ld_win]; b_in [C]; Wd_h [L,2C,3C] (gate/filter interleaved rows, column = tap*C + c)
 *   cond_h: table of b2s_tc_cond_table_tiled for THIS B and T; cond_layer_stride = B*tpb*128*2C elements
 *   Wres_h [L,C,C]: 2^(l/2) * output_projection.weight[:C];  bsum [L,C] fp32: sum_{k<l} 2^(k/2) * output_projection.bias_k[:C]
 *   dvec: this evaluation's step-embedding row, layer l at dvec + b*d_stride + l*C;  dilations_host: L ints (HOST)
 *   yedge0_h, yedge1_h: 16-bit [B*T, C] scratch (only the first / last 16 rows of every 128-frame tile are touched)
 *   flags: int32 [B * tiles], ZERO at launch.  Every tile must be resident: checked against cudaOccupancyMaxActiveClusters.
 *   lens: NULL, or DEVICE int32 [B]: utterance b has lens[b] <= T valid frames (a RAGGED batch padded to T).  Frames at or beyond
 *         lens[b] are treated exactly like frames beyond T - the conv's zero padding (wavenet.py:22-28) - so every valid frame of
 *         utterance b gets the bits it would get in a batch of its own with T = lens[b]; outputs of the padded frames are undefined. */
int b2s_tc_wavenet_stack3_halo(void);
/* tiles (2*ceil(ceil(T/128)/2) per utterance) ONE launch can hold on the current device for utterances of T frames: cluster size =
 * the largest of 8, 6, 4, 2 dividing the tiles per utterance (halo rows inside a cluster go through distributed shared memory), as
 * many clusters as cudaOccupancyMaxActiveClusters reports.  The host splits larger batches by utterance. */
int b2s_tc_wavenet_stack3_max_tiles(int T, int bf16);
int b2s_tc_wavenet_stack3(const void* xin_h, int MF, const void* Win_h, int ld_win, const float* b_in, const void* Wd_h,
                          const void* cond_h, int64_t cond_layer_stride, const void* Wres_h, const float* bsum, const float* dvec,
                          int d_stride, const int* dilations_host, int L, void* yedge0_h, void* yedge1_h, void* z_all_h,
                          int64_t z_layer_stride, int B, int T, int C, int* flags, const int* lens, int bf16, void* stream);

/* b2s_tc_wavenet_stack3 plus, INSIDE the same launch, the deferred skip sum and the head (wavenet.py:96-99): behind the layer tiles
 * the grid carries one CTA pair per four tiles (on the SMs a 16 x 690-frame batch leaves idle) that accumulates
 * S = sum_l z_l Wskip_l^T in TMEM as soon as the layer tiles publish their z tiles (zflags), then out = W_fin relu(W_sp (S + bss) /
 * sqrt(L) + b_sp) + b_fin.  One launch per denoiser evaluation.
 *   Wskip_h [L,C,C] = output_projection.weight[C:2C] per layer; bss [C] = sum_l output_projection.bias_l[C:2C]
 *   Wsp_h [C,C], b_sp [C]; Wfin_h [MF,C], b_fin [MF]; out [B*T, MF] fp32; MF a multiple of 16
 *   flags, zflags: int32 [B * tiles] each, ZERO at launch
 * b2s_tc_wavenet_denoiser3_max_utterances(T, bf16): utterances of T frames one launch can hold (0: use b2s_tc_wavenet_stack3). */
int b2s_tc_wavenet_denoiser3_max_utterances(int T, int bf16);
int b2s_tc_wavenet_denoiser3(const void* xin_h, int MF, const void* Win_h, int ld_win, const float* b_in, const void* Wd_h,
                             const void* cond_h, int64_t cond_layer_stride, const void* Wres_h, const float* bsum, const float* dvec,
                             int d_stride, const int* dilations_host, int L, void* yedge0_h, void* yedge1_h, void* z_all_h,
                             int64_t z_layer_stride, const void* Wskip_h, const float* bss, const void* Wsp_h, const float* b_sp,
                             const void* Wfin_h, const float* b_fin, float* out, int B, int T, int C, int* flags, int* zflags,
                             const int* lens, int bf16, void* stream);
/* The same launch as one of SEVERAL utterance groups of ONE evaluation issued back to back on the stream (batches larger than
 * b2s_tc_wavenet_denoiser3_max_utterances).  chain bit 0: this launch follows another group of the evaluation, bit 1: another group
 * follows.  Same results; a following group's layer kernel does not wait for this group's skip / head tail (it waits at its end
 * instead, so stream-order completion still holds for the kernel after the last group). */
int b2s_tc_wavenet_denoiser3_chained(const void* xin_h, int MF, const void* Win_h, int ld_win, const float* b_in, const void* Wd_h,
                             const void* cond_h, int64_t cond_layer_stride, const void* Wres_h, const float* bsum, const float* dvec,
                             int d_stride, const int* dilations_host, int L, void* yedge0_h, void* yedge1_h, void* z_all_h,
                             int64_t z_layer_stride, const void* Wskip_h, const float* bss, const void* Wsp_h, const float* b_sp,
                             const void* Wfin_h, const float* b_fin, float* out, int B, int T, int C, int* flags, int* zflags,
                             const int* lens, int bf16, int chain, void* stream);

/* LYNXNet pointwise convs on the tensor cores (lynxnet.py:55-56, 60): SwiGLU up-projection and the
 * down-projection with the residual add into the fp32 stream. */
int b2s_tc_lynx_glu(const void* h_h, const void* W_h, const float* bias, void* g_h, int rows, int C, int inner, int bf16,
                    void* stream);
int b2s_tc_linear_residual(const void* p_h, const void* W_h, const float* bias, float* x, int rows, int C, int inner,
                           int bf16, void* stream);
/* Same, plus the NEXT layer's front_cond_inject (strong_cond, lynxnet.py:77-82): x <- (x + W p + b) + cond_next, cond_next a
 * 16-bit [rows, ld_cond] slab of the hoisted conditioner projection.  The next b2s_lynx_prenorm_h is then called with
 * cond_h = NULL (it only adds the step embedding) and neither reads the table nor writes x back.  Bit-identical to the
 * unfolded sequence. */
int b2s_tc_linear_residual_cond(const void* p_h, const void* W_h, const float* bias, float* x, const void* cond_next_h,
                                int ld_cond, int rows, int C, int inner, int bf16, void* stream);

/* 16-bit LYNXNet layer helpers (HBM-bound): fused (x + cond + d) -> residual write-back -> LayerNorm with 16-bit
 * cond table in / 16-bit h out (lynxnet.py:76-84, 54), plain LayerNorm (lynxnet.py:151), depthwise conv + activation
 * with 16-bit in / out and fp32 math (lynxnet.py:57-58).  Same argument meaning as the _f32 entry points, except that
 * b2s_lynx_dwconv_h takes the depthwise weights K-MAJOR: WdwT [ksize][inner] (coalesced loads). */
int b2s_lynx_prenorm_h(float* x, const void* cond_h /* may be NULL: cond already folded in */, int ld_cond, const float* dvec,
                       int d_stride, const float* gamma, const float* beta, void* h_h, int B, int T, int C, int strong_cond,
                       int bf16, void* stream);
int b2s_layernorm_h(const float* x, const float* gamma, const float* beta, void* h_h, int rows, int C, int bf16,
                    void* stream);
int b2s_lynx_dwconv_h(const void* g_h, const float* WdwT, const float* bias, const float* slope, void* p_h, int B, int T,
                      int inner, int ksize, int act /* 0: PReLU(slope), B2S_ACT_*: that activation, < 0: none */, int bf16,
                      void* stream);

/* ---- ConvNeXt aux decoder (the producer of x_start for shallow diffusion; reference modules/aux_decoder/convnext.py:58-87) ----
 * b2s_tc_conv1d: dense Conv1d along time, stride 1, 'same' zero padding per utterance, as ONE tcgen05 GEMM over ksize taps
 *   (inconv / outconv, convnext.py:64-67, :73-76).  a_h [B, T, Cin] 16-bit (Cin % 64 == 0), W_h [N, ksize * Cin] with column
 *   tap * Cin + c, out = act(conv + bias) as fp32 rows (out_f32, ldo) and / or 16-bit rows (out_h, ldoh).
 * b2s_layernorm_hh: LayerNorm over the channels of 16-bit rows -> 16-bit rows, fp32 statistics, eps given (convnext.py:27, :44).
 * b2s_tc_linear_residual_scaled: x <- x + gamma * (p W^T + bias) on the fp32 residual stream (pwconv2 + layer scale + residual,
 *   convnext.py:49-57), gamma may be NULL; x_h (may be NULL) receives the new x as 16-bit rows for the next block's depthwise conv. */
int b2s_tc_conv1d(const void* a_h, const void* W_h, const float* bias, float* out_f32, int ldo, void* out_h, int ldoh, int B, int T,
                  int Cin, int N, int ksize, int act, int bf16, void* stream);
int b2s_layernorm_hh(const void* in_h, const float* gamma, const float* beta, void* out_h, int rows, int C, float eps, int bf16,
                     void* stream);
int b2s_tc_linear_residual_scaled(const void* p_h, const void* W_h, const float* bias, const float* gamma, float* x, void* x_h,
                                  int rows, int C, int inner, int bf16, void* stream);

/* ---- FastSpeech2 acoustic encoder (the producer of the condition tensor; reference modules/fastspeech/acoustic_encoder.py:79-109,
 * modules/fastspeech/tts_modules.py:353-428, modules/commons/common_layers.py:152-263; rotary-position configuration) ----
 * Token-rate tensors are rows r = b * L + l (L = padded tokens per utterance), frame-rate tensors rows b * T + t.  The GEMMs of the
 * transformer layers use b2s_tc_linear / b2s_tc_conv1d / b2s_tc_linear_residual; these are the remaining pieces:
 * b2s_enc_mel2ph_to_dur  dur[b, l] = number of frames with mel2ph == l + 1                              (tts_modules.py:345-351)
 * b2s_enc_embed          x = (sqrt(H) E[token] + dur w_dur + b_dur) * keep, keep = (token != 0)         (tts_modules.py:385-389, :415)
 * b2s_enc_rope           rotary embedding of the q and k thirds of qkv [rows, 3H] fp32, in place; freqs [H / heads / 2]
 * b2s_enc_attention      softmax(q k^T / sqrt(d), padded keys masked) v per (utterance, head) -> 16-bit rows [rows, H]
 * b2s_enc_mask_rows      x[r, :] = 0 where keep[r] == 0
 * b2s_enc_layernorm_mask enc[b, 1 + l, :] = LayerNorm(x[b, l, :]) * keep, enc[b, 0, :] = 0   (table [B, L + 1, H], acoustic_encoder.py:89)
 * b2s_enc_assemble       cond[b, t, :] = enc[b, mel2ph[b, t], :] + spk[b, :] + sum_i (val_i[b, t] w_i + bias_i); val_0 is f0 and enters
 *                        as log(1 + f0 / 700); embeddings [n_var_first, n_var_first + n_var) are summed among themselves first (:61-107). */
int b2s_enc_mel2ph_to_dur(const int64_t* mel2ph, float* dur, int B, int T, int L, void* stream);
int b2s_enc_embed(const int64_t* tokens, const float* dur, const float* E, const float* w_dur, const float* b_dur, float* x, float* keep,
                  int rows, int H, int vocab, void* stream);
int b2s_enc_rope(float* qkv, const float* freqs, int B, int L, int H, int num_heads, void* stream);
int b2s_enc_attention(const float* qkv, const float* keep, void* out_h, int B, int L, int H, int num_heads, int bf16, void* stream);
int b2s_enc_mask_rows(float* x, const float* keep, int rows, int H, void* stream);
int b2s_enc_layernorm_mask(const float* x, const float* gamma, const float* beta, const float* keep, float* enc, int B, int L, int H,
                           float eps, void* stream);
int b2s_enc_assemble(const float* enc, const int64_t* mel2ph, const float* spk /* may be NULL; [B, H], or [B, T, H] with spk_per_frame */,
                     int spk_per_frame, const float* const* vals_host,
                     const float* const* w_host, const float* const* bias_host, int n, int n_var_first, int n_var, float* cond, int B,
                     int T, int L, int H, void* stream);

/* ---- NSF-HiFiGAN vocoder (mel + f0 -> waveform, the step AFTER the sampling loop; reference modules/vocoders/nsf_hifigan.py:57-69,
 * modules/nsf_hifigan/models.py:206-289) ----
 * Activations are time-major rows r = b * T_i + t of the stage's sample rate, channels zero-padded to a multiple of 64 (Cp).
 * b2s_tc_conv1d_dil       b2s_tc_conv1d with a dilation: conv_pre (:269), the residual blocks' first convs (:62-64, act =
 *                         B2S_ACT_LRELU), and the transposed convs (:272) as 3-tap convs over u * Cp output columns
 * b2s_tc_conv1d_residual  x <- x_src + conv(a) + bias on the fp32 stream, y_h <- leaky_relu(x, y_slope) 16-bit (:64-66, :92-94);
 *                         x_src NULL = x; y_h NULL = no copy; y_h must not be the conv's input
 * b2s_voc_phase           phase[b, t] = fmod(sum_{t' < t} wrap(f0[b, t'] / sr * upp), 1), wrap(v) = fmod(v + 0.5, 1) - 0.5 (:138-141;
 *                         mini_nsf adds the chirp term of :256)
 * b2s_voc_source          out[b, n] of the harmonic source at the waveform rate: dim = harmonics + 1 > 0: SineGen + SourceModuleHnNSF
 *                         (:142-147, :160-166, :197-200) with the two random draws given (rand_ini [dim], noise [B, T * upp, dim]);
 *                         dim == 0: Generator.fastsinegen (:252-262)
 * b2s_voc_source_add      x[r, :] += bias + sum_j Wt[j, :] * src[b, t * stride - pad + j] (noise_convs / source_conv: Conv1d with one
 *                         input channel, :273-278); lx_h <- leaky_relu(x, slope) 16-bit; ksize == 0: only the copy.  Wt [ksize, Cp]
 * b2s_voc_avg_act         out_h <- leaky_relu((x_0 + ... + x_{n-1}) / n, slope) 16-bit, n <= 4 residual blocks (:279-285, :271)
 * b2s_voc_post            wav[b, t] = tanh(conv_post(leaky_relu(mean of the blocks, slope))) in fp32, W [ksize, C] (:285-288)
 * b2s_cast_scale_f32_h    16-bit(in * scale): log10 -> ln mel (vocoders/nsf_hifigan.py:60-64) */
int b2s_tc_conv1d_dil(const void* a_h, const void* W_h, const float* bias, float* out_f32, int ldo, void* out_h, int ldoh, int B, int T,
                      int Cin, int N, int ksize, int dil, int act, int bf16, void* stream);
int b2s_tc_conv1d_residual(const void* a_h, const void* W_h, const float* bias, const float* x_src, float* x, void* y_h, float y_slope,
                           int B, int T, int Cin, int N, int ksize, int dil, int bf16, void* stream);
int b2s_voc_phase(const float* f0, float* phase, int B, int T, float sr, int upp, int mini_nsf, void* stream);
int b2s_voc_source(const float* f0, const float* phase, const float* rand_ini, const float* noise, const float* w, const float* bias,
                   float* out, int B, int T, int upp, int dim, float sr, float sine_amp, float noise_std, float voiced_threshold,
                   void* stream);
int b2s_voc_source_add(float* x, void* lx_h, const float* src, const float* Wt, const float* bias, int B, int T, int Cp, int ksize,
                       int stride, int pad, int n_src, float slope, int bf16, void* stream);
int b2s_voc_avg_act(const float* const* xs_host, int n_blocks, void* out_h, int64_t n, float slope, int bf16, void* stream);
int b2s_voc_post(const float* const* xs_host, int n_blocks, const float* W, const float* b0, float* wav, int B, int T, int C, int Cp,
                 int ksize, float slope, void* stream);
int b2s_cast_scale_f32_h(const float* in, void* out_h, int64_t n, float scale, int bf16, void* stream);

/* ---- training-branch losses, forward values (validation during training; reference modules/losses/diff_loss.py:17-37,
 * modules/losses/reflow_loss.py:18-50) ----
 * out[0] = mean over [B, F, M, T] of w_b * loss(a * m, b * m): loss = |.| (l1) or (.)^2; m = mask[b, t, 0 or bin] (non_padding, NULL: none);
 * w_b = the log-normal weight of t_weights[b] (reflow_loss.py:26-33; NULL: 1).  workspace: b2s_masked_loss_workspace_bytes() bytes.
 * Deterministic (fixed summation order), double accumulation across threads. */
int b2s_masked_loss_workspace_bytes(void);
int b2s_masked_loss_f32(const float* a, const float* b, const float* mask, int mask_m, const float* t_weights, int B, int F, int M, int T,
                        int l1, void* workspace, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B2S_H_ */
