/* pcgc_b200.h -- C ABI of libpcgc_b200.so: the B200 (sm_100a) implementation of PCGCv1's per-cube
 * compress/decompress hot path.  Plain pointers and sizes only; no torch / C++ types.
 *
 * Every entry point names the reference interface it replaces (paths into NJUVISION/PCGCv1).
 * The reference has no FFI of its own (it is Python on TF 1.13); the binding a maintainer adds
 * is the ctypes stub shown in INTEGRATION.md (pcgcv1_b200/_lib.py is that stub).
 *
 * Conventions
 *   - every function returns 0 on success or a negative pcgc_status; pcgc_last_error(ctx) gives text.
 *   - "dev" pointers are CUDA device pointers on the ctx's device, "host" pointers are CPU memory.
 *   - tensors are dense, channels-last NDHWC float32 unless stated (the reference's layout).
 *   - all device work is enqueued on the ctx stream (pcgc_set_stream); nothing synchronises unless
 *     the function returns a host value (documented per function).
 *   - a ctx is not thread-safe; different ctxs (one per GPU) are independent.  Host coder entry
 *     points (pcgc_range_*) take no ctx and are re-entrant.
 */
#ifndef PCGC_B200_H_
#define PCGC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCGC_B200_ABI_VERSION 1

typedef struct pcgc_ctx pcgc_ctx;

typedef enum {
  PCGC_OK = 0,
  PCGC_ERR_BAD_ARG = -1,    /* null pointer, unknown enum, shape mismatch */
  PCGC_ERR_BAD_RANGE = -2,  /* symbol range unsupported (single-symbol alphabet, N > PCGC_MAX_SYMBOLS) or k out of range */
  PCGC_ERR_CUDA = -3,       /* a CUDA call failed; text in pcgc_last_error */
  PCGC_ERR_OOM = -4,
  PCGC_ERR_NOT_READY = -5,  /* weights of a required layer were never loaded */
  PCGC_ERR_OVERFLOW = -6,   /* output buffer too small */
  PCGC_ERR_CORRUPT = -7     /* bitstream inconsistent with the CDFs */
} pcgc_status;

/* Nets = the reference's Keras models (checkpoint top-level keys, transform.py:107-111). */
typedef enum {
  PCGC_NET_VOX_ANALYSIS = 0,   /* models/model_voxception.py:71-144  AnalysisTransform  */
  PCGC_NET_VOX_SYNTHESIS = 1,  /* models/model_voxception.py:147-214 SynthesisTransform */
  PCGC_NET_HYPER_ENCODER = 2,  /* models/model_voxception.py:217-252 HyperEncoder       */
  PCGC_NET_HYPER_DECODER = 3,  /* models/model_voxception.py:255-308 HyperDecoder       */
  PCGC_NET_SIMPLE_ANALYSIS = 4,  /* models/model_simple.py:12-51  AnalysisTransform  */
  PCGC_NET_SIMPLE_SYNTHESIS = 5, /* models/model_simple.py:54-95  SynthesisTransform */
  PCGC_NET_COUNT = 6
} pcgc_net;

typedef enum { PCGC_DTYPE_U8 = 0, PCGC_DTYPE_F32 = 1, PCGC_DTYPE_F64 = 2 } pcgc_dtype;

/* Conv engines.  AUTO = tcgen05 implicit GEMM where a layer qualifies, FP32 CUDA-core kernel elsewhere. */
typedef enum { PCGC_ENGINE_AUTO = 0, PCGC_ENGINE_FFMA = 1, PCGC_ENGINE_UMMA = 2 } pcgc_engine;

#define PCGC_MAX_SYMBOLS 64   /* largest max_v-min_v+1 the conditional CDF kernels accept */

/* ---- context ------------------------------------------------------------------------------- */
int pcgc_abi_version(void);
int pcgc_create(pcgc_ctx** out, int device);
void pcgc_destroy(pcgc_ctx* ctx);
const char* pcgc_last_error(const pcgc_ctx* ctx);
/* stream: a cudaStream_t cast to void* (NULL = legacy default stream). */
int pcgc_set_stream(pcgc_ctx* ctx, void* stream);
int pcgc_set_engine(pcgc_ctx* ctx, int engine);
/* number of kernels this library has launched on ctx since creation (bench "gpu_launches"). */
int64_t pcgc_launch_count(const pcgc_ctx* ctx);
int pcgc_synchronize(pcgc_ctx* ctx);
/* Pipelining aid.  Several entry points synchronise the stream only to read the device-side error flag (bad symbol range,
 * tcgen05 barrier timeout ...).  With deferred checks ON they return right after enqueueing; the caller MUST call
 * pcgc_synchronize (which reports the flag) before it consumes any result.  Default OFF. */
int pcgc_set_deferred_checks(pcgc_ctx* ctx, int on);
/* Measurement aid (bench.py roofline): when on, every kernel launch group is bracketed by CUDA events
 * on the ctx stream.  pcgc_profile_report synchronises, writes a JSON array of
 * {"tag","count","ms","flops","bytes"} (algorithmic work per tag) into buf and clears the records. */
int pcgc_profile_enable(pcgc_ctx* ctx, int on);
int pcgc_profile_report(pcgc_ctx* ctx, char* buf, int64_t cap);

/* ---- weights (replaces tf.train.Checkpoint.restore, transform.py:36-38,70-72,107-112,214-218) -- */
/* kernel: HOST float32 in the Keras layout ([k,k,k,Cin,Cout]; Conv3DTranspose [k,k,k,Cout,Cin]),
 * kshape = its 5 dims, bias: HOST float32 [Cout] or NULL for use_bias=False layers.
 * layer = the Keras name= of the layer ("conv_in", "vrn1_1_conv1_1", "up_1", ...). */
int pcgc_load_conv(pcgc_ctx* ctx, int net, const char* layer, const float* kernel, const int64_t kshape[5],
                   const float* bias);
/* EntropyBottleneck variables (models/entropy_model.py:42-68), HOST float32:
 * matrices = matrix_0..3 concatenated ([C,3,1],[C,3,3],[C,3,3],[C,1,3]), biases = bais_0..3
 * ([C,3,1]x3,[C,1,1]), factors = factor_0..3 (same shapes as biases).  slot 0 = "estimator"
 * of the checkpoint; slot 1 is free for a second bottleneck. */
int pcgc_load_bottleneck(pcgc_ctx* ctx, int slot, int channels, const float* matrices, const float* biases,
                         const float* factors);

/* Test hook for the tcgen05 engine: one 3x3x3 stride-1 SAME conv (+bias, optional ReLU) of a float32 NDHWC
 * batch [B,n,n,n,cin] (cin in {8,16,32,64}, cout <= 64, n in {16,32,64}) with a HOST Keras kernel
 * [3,3,3,cin,cout]; converts to the engine's split-bf16 format, runs the UMMA kernel, writes float32
 * [B,n,n,n,cout].  wt = 1 | 2 | 4 selects the y-banded form of the kernel (each M row produces wt output lines).
 * Synchronises.  Compared against a plain FP32 conv in tests/test_gpu_umma.py. */
int pcgc_debug_conv3_umma(pcgc_ctx* ctx, const float* in_dev, int n, int cin, int cout, const float* kernel_host,
                          const float* bias_host, int relu, int B, int wt, float* out_dev);

/* ---- transforms (dev pointers) --------------------------------------------------------------- */
/* AnalysisTransform()(x), one call for B cubes instead of tf.map_fn(parallel_iterations=1)
 * (transform.py:42-48,116-122).  cubes: [B,64,64,64,1] of `dtype`; y: [B,16,16,16,16] (voxception)
 * or [B,8,8,8,32] (simple).  net = PCGC_NET_VOX_ANALYSIS | PCGC_NET_SIMPLE_ANALYSIS. */
int pcgc_analysis(pcgc_ctx* ctx, int net, const void* cubes_dev, int dtype, int B, float* y_dev);
/* SynthesisTransform()(y) (transform.py:79-84,181-183,251-256): logits [B,64,64,64,1]. */
int pcgc_synthesis(pcgc_ctx* ctx, int net, const float* y_dev, int B, float* logits_dev);
/* HyperEncoder()(y) (transform.py:125-131): y [B,16,16,16,16] -> z [B,8,8,8,8]. */
int pcgc_hyper_encode(pcgc_ctx* ctx, const float* y_dev, int B, float* z_dev);
/* HyperDecoder()(z_hat) followed by scales = max(scales, 1e-9) (transform.py:138-146,225-233):
 * z_hat [B,8,8,8,8] -> loc, scale [B,16,16,16,16]; scale = max(|.|, scale_floor).
 * Bit-reproducible across calls, batch sizes, devices of the same type, pcgc_set_engine settings and the PCGC_UMMA_* tuning
 * switches (one pinned kernel program: the outputs become integer CDF tables on both sides of the stream). */
int pcgc_hyper_decode(pcgc_ctx* ctx, const float* z_hat_dev, int B, float scale_floor, float* loc_dev,
                      float* scale_dev);

/* ---- factorized entropy model (models/entropy_model.py) ---------------------------------------- */
/* EntropyBottleneck.__call__(x, training=False) (:153-181): x_hat = round-half-even(x),
 * p = max(|sigmoid(s*u)-sigmoid(s*l)|, bound).  x: [n_vox, C] (C fastest).  Optional outputs
 * (NULL to skip): x_hat_dev, p_dev [n_vox*C]; bits_dev (double, sum(log2 p) * -1);
 * minmax_dev int32[2] = {floor(min x_hat), ceil(max x_hat)} (:249-250).  No host sync. */
int pcgc_factorized_quantize_likelihood(pcgc_ctx* ctx, int slot, const float* x_dev, int64_t n_vox, int C,
                                        float likelihood_bound, float* x_hat_dev, float* p_dev,
                                        double* bits_dev, int32_t* minmax_dev);
/* EntropyBottleneck._get_cdf (:183-221): int32 cdf [C, N+1] to HOST, N = max_v-min_v+1 >= 2.
 * Synchronises the ctx stream. */
int pcgc_factorized_cdf(pcgc_ctx* ctx, int slot, int min_v, int max_v, float likelihood_bound, int precision,
                        int32_t* cdf_host);

/* ---- conditional (Laplace) entropy model (models/conditional_entropy_model.py) ------------------ */
/* SymmetricConditional.__call__(y, loc, scale, training=False) (:71-93) per cube.
 * y, loc, scale: [B, E] (E elements per cube, 65536 for voxception).  Optional outputs:
 * y_hat_dev, p_dev [B,E]; bits_dev double[B]; minmax_dev int32[2B] = {min_v, max_v} per cube
 * (:153-154).  No host sync. */
int pcgc_laplace_quantize_likelihood(pcgc_ctx* ctx, const float* y_dev, const float* loc_dev,
                                     const float* scale_dev, int B, int64_t E, float likelihood_bound,
                                     float* y_hat_dev, float* p_dev, double* bits_dev, int32_t* minmax_dev);
/* Makes per-cube symbol ranges codable and storable before the tables are built: min_v <- min(min_v, 0), max_v <- max(max_v, 0)
 * (the container packs max*16 - min with min <= 0 <= max, dataprocess/inout_bitstream.py:95-96) and max_v <- 1 for the
 * one-symbol range {0} (pmf_to_quantized_cdf needs two symbols, entropy_model.py:192-193).  minmax_dev int32[2*n_pairs]. */
int pcgc_widen_symbol_ranges(pcgc_ctx* ctx, int32_t* minmax_dev, int n_pairs);
/* Encoder side of SymmetricConditional._get_cdf + range_encode's table lookup (:95-124,142-161):
 * for every element builds its quantised CDF row over min_v[b]..max_v[b] and emits only the
 * interval of the element's own symbol: interval = lower | (upper-lower-1) << 16.
 * y_hat: [B,E] rounded latents; minmax_dev int32[2B] (device, e.g. from the call above). */
int pcgc_laplace_intervals(pcgc_ctx* ctx, const float* y_hat_dev, const float* loc_dev, const float* scale_dev,
                           int B, int64_t E, const int32_t* minmax_dev, float likelihood_bound, int precision,
                           uint32_t* intervals_dev);
/* Decoder side (:186-195): full rows.  Cube b's rows start at row_offset[b] (in uint16 units,
 * HOST array of B+1 prefix sums of E*N_b) and hold cdf[0..N_b-1] as uint16 (cdf[N_b] = 2^precision
 * is implied).  minmax_host int32[2B]. */
int pcgc_laplace_cdf(pcgc_ctx* ctx, const float* loc_dev, const float* scale_dev, int B, int64_t E,
                     const int32_t* minmax_host, float likelihood_bound, int precision,
                     const int64_t* row_offset_host, uint16_t* cdf_dev);

/* Test hook: the DEVICE copy of the 16-bit normaliser on given pmf rows (device float32 [rows,N], 2 <= N <=
 * PCGC_MAX_SYMBOLS) -> device int32 cdf [rows,N+1]; must equal pcgc_pmf_to_quantized_cdf bit for bit.  Synchronises. */
int pcgc_debug_quantize_pmf(pcgc_ctx* ctx, const float* pmf_dev, int64_t rows, int N, int precision, int32_t* cdf_dev);

/* "noise" quantisation of the training graph (entropy_model.py:105-107, conditional_entropy_model.py:62-64): when noise = 1
 * the two *_quantize_likelihood entry points return x + U(-1/2, 1/2) (counter-based Philox4x32-10 keyed by `seed`, element
 * index as the counter: reproducible, independent of the launch shape) and the likelihood AT that value instead of round(x);
 * minmax outputs are then meaningless.  noise = 0 (default) is "symbols" = round half to even. */
int pcgc_set_quantize_mode(pcgc_ctx* ctx, int noise, uint64_t seed);
/* Test hook (host, no ctx): the four draws of elements 4*v .. 4*v+3 for `seed` as the kernels make them (the conditional
 * model's launch uses seed + 0x9E3779B97F4A7C15 so that y and z do not share a stream). */
int pcgc_debug_noise(uint64_t seed, uint64_t v, float* out4);

/* ---- GPU-side range coding of the per-cube strings (models/conditional_entropy_model.py:126-201) ----
 * All pointers are DEVICE pointers, nothing synchronises; device-side failures (overflow, bad range) set the ctx error flag
 * that pcgc_synchronize reports.  The strings are byte-identical to the host coder's (same state machine source). */
/* pcgc_laplace_cdf with device-resident headers: minmax_dev int32[2B], row_offset_dev int64[B+1] (uint16 units);
 * rows_total = row_offset[B] (only used for the profile's byte count). */
int pcgc_laplace_cdf_dev(pcgc_ctx* ctx, const float* loc_dev, const float* scale_dev, int B, int64_t E,
                         const int32_t* minmax_dev, const int64_t* row_offset_dev, double rows_total, float likelihood_bound,
                         int precision, uint16_t* cdf_dev);
/* range_encode of B cubes from pcgc_laplace_intervals' output (precision 16, E % 32 == 0): cube b is coded into
 * scratch_dev + b*stride (E <= 65536; stride a multiple of 16 and >= 6*E + 32: the string, at most one 16-bit word per symbol,
 * followed by the encoder's 32-bit digit sums), lens_dev[b] receives its length, then the strings are concatenated into
 * packed_dev (capacity cap; 2*E + 2 bytes per cube always suffice) with offsets_dev int64[B+1] (offsets[B] = total bytes). */
int pcgc_range_encode_intervals_dev(pcgc_ctx* ctx, const uint32_t* intervals_dev, int B, int64_t E, int precision,
                                    uint8_t* scratch_dev, int64_t stride, int64_t* lens_dev, uint8_t* packed_dev, int64_t cap,
                                    int64_t* offsets_dev);
/* range_decode of B cubes: string b = packed_dev[offsets[b] .. offsets[b+1]), rows as written by pcgc_laplace_cdf(_dev),
 * y_hat_dev float32 [B,E] = symbol + min_v (conditional_entropy_model.py:196-199).  E % 32 == 0; max_n = the largest
 * N_b = max_v - min_v + 1 of the call (<= PCGC_MAX_SYMBOLS; sizes the shared-memory row window). */
int pcgc_range_decode_rows_dev(pcgc_ctx* ctx, const uint8_t* packed_dev, const int64_t* offsets_dev, int B, int64_t E,
                               const uint16_t* rows_dev, const int64_t* row_offset_dev, double rows_total, const int32_t* minmax_dev,
                               int max_n, int precision, float* y_hat_dev);

/* ---- host twins of the CDF builders (bit-identical tables: det_math.h + cdf_norm.h are shared by host and device) ---------
 * Interoperability / test aids: a CPU that has (loc, scale) or the bottleneck parameters can build exactly the tables the GPU
 * coded with, and decode the stream with pcgc_range_decode_rows / pcgc_range_decode.  Not on the product path. */
int pcgc_host_laplace_cdf(const float* loc, const float* scale, int B, int64_t E, const int32_t* minmax, float likelihood_bound,
                          int precision, const int64_t* row_offset, uint16_t* rows, int threads);
int pcgc_factorized_cdf_host(pcgc_ctx* ctx, int slot, int min_v, int max_v, float likelihood_bound, int precision,
                             int32_t* cdf_host);

/* ---- top-k occupancy classification (dataprocess/inout_points.py:147-179) ---------------------- */
/* select_voxels: per cube k = ks[b] (caller computes int(rho * n_points)); threshold = k-th
 * largest logit; mask = logits >= threshold (ties kept).  k == 0 reproduces the reference's
 * values[-0] quirk (threshold = smallest logit > -2.0).  Outputs: mask_dev uint8 [B,V],
 * thres_dev float[B] (optional), count_dev int32[B] (optional, voxels set). */
int pcgc_topk_select(pcgc_ctx* ctx, const float* logits_dev, int B, int64_t V, const int32_t* ks_dev,
                     uint8_t* mask_dev, float* thres_dev, int32_t* count_dev);
/* fixed_thres bypass of select_voxels (:157-160). */
int pcgc_threshold_select(pcgc_ctx* ctx, const float* logits_dev, int B, int64_t V, float thres,
                          uint8_t* mask_dev, int32_t* count_dev);

/* ---- host range coder (replaces coder_ops.range_encode / range_decode / pmf_to_quantized_cdf,
 *      models/entropy_model.py:218,258,298; models/conditional_entropy_model.py:122,161,195) ------ */
/* pmf float32 [rows,N] -> cdf int32 [rows,N+1]. */
int pcgc_pmf_to_quantized_cdf(const float* pmf, int64_t rows, int N, int precision, int32_t* cdf);
/* Symbols sym[i] coded with row (i % cdf_rows) of cdf [cdf_rows, N+1] (the EntropyBottleneck
 * broadcast [1,C,N+1] against data [M,C]).  out/cap: caller buffer; *len receives the size. */
int pcgc_range_encode(const int16_t* sym, int64_t n, const int32_t* cdf, int cdf_rows, int N, int precision,
                      uint8_t* out, int64_t cap, int64_t* len);
int pcgc_range_decode(const uint8_t* data, int64_t nbytes, int64_t n, const int32_t* cdf, int cdf_rows, int N,
                      int precision, int16_t* sym);
/* As pcgc_range_decode, publishing progress: after every `step` symbols (and at the end) the count of symbols already
 * written to sym is stored to *progress (release order; -1 on a bad argument), so a second thread can start on the head of
 * one long string -- decompress_hyper starts the hyper decoder on the first cubes while the rest of z is still decoded. */
int pcgc_range_decode_progress(const uint8_t* data, int64_t nbytes, int64_t n, const int32_t* cdf, int cdf_rows, int N,
                               int precision, int16_t* sym, int64_t* progress, int64_t step);
/* Per-element intervals from pcgc_laplace_intervals -> one string. */
int pcgc_range_encode_intervals(const uint32_t* intervals, int64_t n, int precision, uint8_t* out, int64_t cap,
                                int64_t* len);
/* Per-element rows from pcgc_laplace_cdf (N uint16 per element, last implied) -> symbols. */
int pcgc_range_decode_rows(const uint8_t* data, int64_t nbytes, int64_t n, const uint16_t* rows, int N,
                           int precision, int16_t* sym);
/* Many independent strings at once on a host thread pool (threads <= 0: all cores).
 * Cube b: intervals + b*E, output written at out + b*stride, length in lens[b]. */
int pcgc_range_encode_intervals_batch(const uint32_t* intervals, int B, int64_t E, int precision, uint8_t* out,
                                      int64_t stride, int64_t* lens, int threads);
int pcgc_range_decode_rows_batch(const uint8_t* const* data, const int64_t* nbytes, int B, int64_t E,
                                 const uint16_t* rows, const int64_t* row_offset, const int32_t* minmax,
                                 int precision, int16_t* sym, int threads);

/* Same, writing y_hat = symbol + min_v as float32 [B,E] (conditional_entropy_model.py:196-199). */
int pcgc_range_decode_rows_batch_f32(const uint8_t* const* data, const int64_t* nbytes, int B, int64_t E,
                                     const uint16_t* rows, const int64_t* row_offset, const int32_t* minmax,
                                     int precision, float* y_hat, int threads);

/* ---- training step primitives (train_hyper.py:184-214, loss.py:8-33) -------------------------------
 * Exact FP32 on CUDA cores, deterministic (fixed reduction orders, no atomics).  All pointers are DEVICE pointers, float32
 * NDHWC activations, kernels in the Keras layouts ([k,k,k,Cin,Cout]; Conv3DTranspose [k,k,k,Cout,Cin]); nothing synchronises.
 * pcgcv1_b200/training.py strings them into the reference's training graph (torch.autograd is only the tape). */
/* y = act(conv(x) + bias): x [B,n,n,n,cin] -> out [B,m,m,m,cout], m = n/stride (conv) or n*stride (transposed); k odd, SAME. */
int pcgc_train_conv_forward(pcgc_ctx* ctx, const float* x, int B, int n, int cin, int cout, int k, int stride, int transposed,
                            const float* w_keras, const float* bias, int relu, float* out);
/* dx [B,n,n,n,cin] from the output gradient g [B,m,m,m,cout] of the same layer (n = the layer's INPUT grid). */
int pcgc_train_conv_dgrad(pcgc_ctx* ctx, const float* g, int B, int n, int cin, int cout, int k, int stride, int transposed,
                          const float* w_keras, float* dx);
/* dw (the layer's Keras layout) and db [cout] (NULL for use_bias=False layers) from the layer input x and g. */
int pcgc_train_conv_wgrad(pcgc_ctx* ctx, const float* x, const float* g, int B, int n, int cin, int cout, int k, int stride,
                          int transposed, float* dw, float* db);
/* out = g where y > 0 else 0 (gradient through a fused ReLU; y = the layer output). */
int pcgc_train_relu_backward(pcgc_ctx* ctx, const float* g, const float* y, int64_t n, float* out);
/* _VoxceptionResNet merge (model_voxception.py:64-67): out = relu(x + concat[t12, t23]) over nvox voxels of c channels. */
int pcgc_train_vrn_merge(pcgc_ctx* ctx, const float* x, const float* t12, const float* t23, int64_t nvox, int c, float* out);
int pcgc_train_vrn_merge_backward(pcgc_ctx* ctx, const float* g, const float* out, int64_t nvox, int c, float* gx, float* g12, float* g23);
/* scale = max(|s|, floor) (model_voxception.py:308, train_hyper.py:191) and its gradient. */
int pcgc_train_abs_floor(pcgc_ctx* ctx, const float* s, int64_t n, float floor_v, float* out);
int pcgc_train_abs_floor_backward(pcgc_ctx* ctx, const float* g, const float* s, int64_t n, float floor_v, float* out);
/* Gradients of coef * sum(log max(likelihood, bound)) w.r.t. the (noisy) latents, loc and scale of the conditional model. */
int pcgc_train_laplace_backward(pcgc_ctx* ctx, const float* y_t, const float* loc, const float* scale, int64_t n, float bound, float coef,
                                float* gy, float* gloc, float* gscale);
/* EntropyBottleneck(z, training=True) from the RAW variables (matrices / biases / factors concatenated in the pcgc_load_bottleneck
 * order): z_t = z + U(-1/2, 1/2) (Philox, `seed`), logsum_dev double[1] = sum(log max(likelihood, bound)). */
int pcgc_train_factorized_forward(pcgc_ctx* ctx, const float* matrices, const float* biases, const float* factors, int C, const float* z,
                                  int64_t nvox, uint64_t seed, float bound, float* z_t, double* logsum_dev);
/* The same for the EntropyBottleneck on z_t [nvox, C], with the gradients of its RAW variables (the pcgc_load_bottleneck order). */
int pcgc_train_factorized_backward(pcgc_ctx* ctx, const float* matrices, const float* biases, const float* factors, int C, const float* z_t,
                                   int64_t nvox, float bound, float coef, float* gz, float* gmatrices, float* gbiases, float* gfactors);
/* get_bce_loss (loss.py:8-33): sums_dev double[4] = {sum_empty -log(1-occ), sum_full -log(occ), #empty, #full}. */
int pcgc_train_bce(pcgc_ctx* ctx, const float* logits, const uint8_t* label, int64_t n, double* sums_dev);
int pcgc_train_bce_backward(pcgc_ctx* ctx, const float* logits, const uint8_t* label, int64_t n, const double* sums_dev, float w_empty,
                            float w_full, float* g);
/* get_focal_loss (loss.py:83-93; BASELINE config 5's "focal occupancy loss") on y_pred = sigmoid(logits), label in {0, 1}:
 * sums_dev double[2] = {-sum(alpha (1 - pt_1)^gamma log pt_1), -sum((1 - alpha) pt_0^gamma log(1 - pt_0))} with the reference's
 * clip to [1e-3, .999] (the constant "other" branch of each tf.where included); the loss is their sum.  The backward entry
 * writes g = weight * d loss / d logits (zero where the clip is active).  gamma >= 1. */
int pcgc_train_focal(pcgc_ctx* ctx, const float* logits, const uint8_t* label, int64_t n, float gamma, float alpha, double* sums_dev);
int pcgc_train_focal_backward(pcgc_ctx* ctx, const float* logits, const uint8_t* label, int64_t n, float gamma, float alpha, float weight,
                              float* g);
/* tf.train.AdamOptimizer update with the bias-corrected step lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t). */
int pcgc_train_adam(pcgc_ctx* ctx, float* p, const float* g, float* m, float* v, int64_t n, float lr_t, float beta1, float beta2, float eps);

/* ---- point-cloud I/O either side of the codec ----------------------- */
/* Multithreaded memcpy of a large host buffer (a pinned staging buffer -> the NumPy array handed to the caller; first-touch
 * page faults of the fresh destination are spread over the threads).  No reference counterpart: plumbing of this build. */
int pcgc_host_copy(void* dst, const void* src, int64_t nbytes, int threads);
/* load_ply_data (dataprocess/inout_points.py:8-28): every line of the ASCII text whose first three single-space
 * separated tokens parse as floats is a point, truncated to int32; other lines (header, comments) are skipped.
 * xyz: HOST int32 [cap,3]; *n = points found (set even when PCGC_ERR_OVERFLOW says cap was too small).  A line with a
 * numeric first token but fewer than three tokens returns PCGC_ERR_CORRUPT (the reference raises IndexError there). */
int pcgc_ply_parse(const char* text, int64_t nbytes, int32_t* xyz, int64_t cap, int64_t* n, int threads);
/* write_ply_data (inout_points.py:30-46) for integer coordinates: header + "x y z\n" per point, byte-identical to the
 * reference's output.  out: HOST buffer of at least 160 + 36*n bytes. */
int pcgc_ply_format(const int32_t* xyz, int64_t n, char* out, int64_t cap, int64_t* len, int threads);
/* The cube partition of load_points (inout_points.py:50-90): cube = point // cube_size, local = point % cube_size, cubes
 * with fewer than min_num points dropped (a single-point cube counts as 3, the reference's 1-D shape quirk), kept cubes
 * ordered by x + y*step + z*step^2 with step = max kept cube coordinate + 1, file order inside a cube.
 * Outputs (HOST, caller-allocated): local_sorted int16 [local_cap,3] (points of the kept cubes, grouped in sorted cube
 * order; local_cap = n suffices unless negative coordinates make the reference repeat a cube -- PCGC_ERR_OVERFLOW then
 * reports the needed size in *n_points), cube_pos_seen int64 [n,3] (kept cubes in first-appearance order = the reference's cube_positions
 * return value), cube_pos_sorted int64 [n,3], counts_sorted int64 [n], *n_cubes, *n_points. */
int pcgc_partition_points(const int32_t* xyz, int64_t n, int cube_size, int min_num, int16_t* local_sorted, int64_t local_cap,
                          int64_t* cube_pos_seen, int64_t* cube_pos_sorted, int64_t* counts_sorted, int64_t* n_cubes,
                          int64_t* n_points);
/* points2voxels (inout_points.py:116-132) on the device: local_dev int16 [n,3] grouped per cube, offsets_host int64
 * [B+1] -> cubes_dev uint8 [B,S,S,S] (0/1).  Coordinates outside [0,S) return an error (the reference raises
 * IndexError).  Synchronises (error flag). */
int pcgc_voxelize(pcgc_ctx* ctx, const int16_t* local_dev, const int64_t* offsets_host, int B, int S, uint8_t* cubes_dev);
/* voxels2points (inout_points.py:134-143) on the device: mask_dev uint8 [B,S,S,S] -> counts_dev int32 [B], points_dev
 * int16 [cap,3] (non-zero voxels as (d,h,w), lexicographic per cube, cubes in order = np.where order) and *total_dev
 * (int64, may exceed cap: only the first cap points are written).  S must be a multiple of 16.  Does not synchronise. */
int pcgc_extract_points(pcgc_ctx* ctx, const uint8_t* mask_dev, int B, int S, int32_t* counts_dev, int16_t* points_dev,
                        int64_t cap, int64_t* total_dev);

#ifdef __cplusplus
}
#endif
#endif /* PCGC_B200_H_ */
