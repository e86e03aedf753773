/*
 * sudormrf_b200.h — C ABI of the B200-native SuDoRM-RF forward path.
 *
 * The reference (etzinis/sudo_rm_rf) has no native/FFI layer: its boundary is
 * the Python nn.Module contract.  This header is the C-ABI that sits UNDER the
 * Python mirror of that contract (sudo_rm_rf_b200/improved_sudormrf.py etc.);
 * each entry point names the reference interface it replaces (file:line in
 * /root/reference).  Plain pointers and sizes only; no torch types.
 *
 * Ownership: the caller owns every byte (parameters, packed weights,
 * workspace, inputs, outputs are caller-allocated device buffers); the library
 * never allocates or frees device memory and keeps no mutable global state
 * beyond one-time cudaFuncSetAttribute calls.  All work is enqueued on the
 * stream passed in; nothing synchronises.  Fully re-entrant (nn.DataParallel
 * calls forward from one host thread per device).
 *
 * Errors: integer return codes, 0 = OK, negative = failure (see
 * sdr_error_string).  No C++ exceptions cross the ABI.
 */
#ifndef SUDORMRF_B200_H
#define SUDORMRF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDR_ABI_VERSION 2

/* return codes */
#define SDR_OK                 0
#define SDR_ERR_BAD_CONFIG    -1   /* constructor arguments the kernels cannot run */
#define SDR_ERR_BAD_ARGUMENT  -2   /* null pointer, wrong count, B/T <= 0 ... */
#define SDR_ERR_WORKSPACE     -3   /* workspace / packed buffer too small */
#define SDR_ERR_CUDA          -4   /* a CUDA call or launch failed */
#define SDR_ERR_UNSUPPORTED   -5   /* valid for the reference, not implemented here */

/* Constructor arguments of the reference models:
 *   improved_sudormrf.py:224-231   SuDORMRF.__init__
 *   groupcomm_sudormrf_v2.py:232-241 GroupCommSudoRmRf.__init__
 *   causal_improved_sudormrf_v3.py:121-129 CausalSuDORMRF.__init__
 *   sudormrf.py:186-193 SuDORMRF.__init__ (the original model)             */
typedef struct {
    int32_t variant;            /* 0 = improved SuDORMRF, 1 = GroupCommSudoRmRf, 2 = CausalSuDORMRF,
                                   3 = the original SuDORMRF (sudormrf.py) */
    int32_t in_audio_channels;  /* 1 for improved */
    int32_t out_channels;
    int32_t in_channels;
    int32_t num_blocks;
    int32_t upsampling_depth;
    int32_t enc_kernel_size;
    int32_t enc_num_basis;
    int32_t num_sources;
    int32_t group_size;         /* ignored for improved */
} sdr_config;

typedef void* sdr_stream;       /* a cudaStream_t (CUstream); NULL = legacy default stream */

int         sdr_abi_version(void);
const char* sdr_error_string(int code);

/* Number of parameter tensors in the reference's state_dict() order
 * (improved_sudormrf.py:247-281, :170-196; groupcomm_sudormrf_v2.py:262-299,
 * :347-354, :401-403; causal_improved_sudormrf_v3.py:146-189, :71-96) and the
 * element count of parameter i.  For the causal model the caller passes
 * skipinit_gain already multiplied by the block's alpha and proj_1x1's weight
 * divided by its beta (both 1.0 in the reference's constructor, :165-174).
 * Original model (variant 3, sudormrf.py:211-264, :134-162): every state_dict
 * entry in order EXCEPT the trailing ln_mask_in.{weight,bias}, which forward
 * never reads (:264).                                                         */
int     sdr_num_params(const sdr_config* cfg);
int64_t sdr_param_numel(const sdr_config* cfg, int index);

/* Padded length rule of pad_to_appropriate_length (improved_sudormrf.py:303-310;
 * variant 3: sudormrf.py:206-209,283-293, multiples of lcm(hop, 2^depth)).     */
int64_t sdr_padded_length(const sdr_config* cfg, int64_t T);

/* Packed weights: one flat device buffer holding every parameter plus derived
 * layouts (decoder weight as a [S*K, S*N] matrix, ...).  A single buffer so the
 * multi-GPU driver can broadcast it with ONE ncclBroadcast.  Replaces the
 * per-forward module replication of nn.DataParallel
 * (run_improved_sudormrf.py:118).                                            */
size_t sdr_packed_weight_bytes(const sdr_config* cfg);
int    sdr_pack_weights(const sdr_config* cfg,
                        const float* const* params, /* host array of n device pointers, state_dict order, fp32 contiguous */
                        int n_params,
                        void* packed, size_t packed_bytes, sdr_stream stream);

/* Caller-allocated scratch for a forward at batch B, length T (all
 * intermediates + the GlobLN statistics).                                    */
size_t sdr_workspace_bytes(const sdr_config* cfg, int B, int64_t T);

/* SuDORMRF.forward (improved_sudormrf.py:283-301) /
 * GroupCommSudoRmRf.forward (groupcomm_sudormrf_v2.py:302-322) /
 * CausalSuDORMRF.forward (causal_improved_sudormrf_v3.py:191-211) /
 * the original SuDORMRF.forward (sudormrf.py:266-292), optionally
 * followed by mixture_consistency.apply(..., 'uniform')
 * (mixture_consistency.py:14-36; only when in_audio_channels == 1).
 *   mixture: device [B, in_audio_channels, T] fp32 contiguous
 *   out:     device [B, num_sources*in_audio_channels, T] fp32 contiguous   */
int sdr_forward(const sdr_config* cfg, const void* packed,
                const float* mixture, float* out, int B, int64_t T,
                int apply_mixture_consistency,
                void* workspace, size_t workspace_bytes, sdr_stream stream);

/* Number of kernels one sdr_forward enqueues (the benchmark's gpu_launches claim): for T = 32000 samples, and for
 * any length (the depthwise levels run as one pass or level by level depending on the padded length).          */
int sdr_forward_launch_count(const sdr_config* cfg);
int sdr_forward_launch_count_at(const sdr_config* cfg, int64_t T);

/* Same call with HOST buffers (pinned for real asynchrony): H2D copy of the
 * mixture, forward, D2H copy of the estimates, all on `stream`.  `dev_io` is a
 * device staging buffer of sdr_host_staging_bytes().  This is the end-to-end
 * entry the benchmark's `e2e` figure times.                                  */
size_t sdr_host_staging_bytes(const sdr_config* cfg, int B, int64_t T);
int sdr_forward_host(const sdr_config* cfg, const void* packed,
                     const float* host_mixture, float* host_out, int B, int64_t T,
                     int apply_mixture_consistency,
                     void* dev_io, size_t dev_io_bytes,
                     void* workspace, size_t workspace_bytes, sdr_stream stream);

/* mixture_consistency.apply (mixture_consistency.py:14-36).
 * weights_type: 0 = 'uniform', 1 = 'magsq'.  est/out [B,S,T], mix [B,1,T].
 * `scratch` (device, >= B*S doubles) is only used by 'magsq'.               */
int sdr_mixture_consistency(const float* est, const float* mix, float* out,
                            int B, int S, int64_t T, int weights_type,
                            void* scratch, sdr_stream stream);

/* ---- per-stage entry points (stage-level parity tests; same kernels the
 *      forward launches).  "Deferred GlobLN": a producer stores its RAW output
 *      and accumulates per-sample (sum, sum of squares) in fp64 into `stats`
 *      ([samples][2] doubles, zero-initialised by the caller); the consumer
 *      applies gamma*(x-mean)*rstd+beta (+PReLU) while loading.
 *      GlobLN = improved_sudormrf.py:30-47.                                  */

/* description of a deferred normalisation applied to an input while loading */
typedef struct {
    const double* stats;   /* [samples][2] or NULL = no normalisation */
    const float*  gamma;   /* [C] */
    const float*  beta;    /* [C] */
    const float*  prelu;   /* 1 element ([C] when prelu_per_channel), or NULL = no activation */
    double        count;   /* elements per sample the statistics were taken over */
    int32_t       prelu_per_channel;   /* 0: nn.PReLU() (one shared slope); 1: nn.PReLU(C) (sudormrf.py:33,71) */
} sdr_norm_in;

/* nn.Conv1d(A, N, k, stride=k/2, padding=k/2, bias=False) on the zero-padded
 * waveform (improved_sudormrf.py:247-251,286,303-314).
 * wav [B,A,T] -> enc [B,N,L] (L = padded_length/hop), stats over (N,L).      */
int sdr_encoder(const float* wav, const float* weight, float* enc, double* stats,
                int B, int A, int64_t T, int N, int K, int L, sdr_stream stream);

/* The same encoder on the tensor cores: the tcgen05 GEMM kernel with a "window" operand producer
 * (A[position, tap] = wav[hop*position + tap - pad], built in registers: no im2col in HBM); the
 * weight is converted once to pre-swizzled bf16 hi/lo images (taps zero-padded to 64).  N >= 32. */
size_t sdr_encoder_mma_packed_bytes(int N, int A, int K);
int sdr_encoder_mma_pack(const float* weight, int N, int A, int K, void* packed, sdr_stream stream);
int sdr_encoder_mma(const float* wav, const void* packed_w, float* enc, double* stats,
                    int B, int A, int64_t T, int N, int K, int L, sdr_stream stream);

/* 1x1 Conv1d as a GEMM: y[b,m,l] = sum_k W[m,k] f(x[b,k,l]) + bias[m]
 * (+ residual[b,m,l]); f = deferred norm/PReLU.  epilogue 0: plain,
 * 1: relu(y) * gate[b, m % gate_channels, l] (mask path, improved_sudormrf.py:296-298).
 * Replaces bottleneck / proj_1x1.conv / res_conv / mask_net.1 (+decoder GEMM). */
int sdr_pointwise(const float* x, const sdr_norm_in* fin, const float* W, const float* bias,
                  const float* residual, const float* gate, int gate_channels,
                  float* y, double* stats_out,
                  int samples, int M, int Kc, int L, int epilogue, sdr_stream stream);

/* Tensor-core (tcgen05, bf16x3 split, fp32 accumulate) variant of sdr_pointwise for
 * channel counts that fill a tile: M % 128 == 0 and Kc % 64 == 0.  The weight is
 * first converted to pre-swizzled bf16 hi/lo shared-memory images (bulk-TMA source);
 * sdr_pointwise_mma_packed_bytes returns 0 when the shape is not eligible.           */
size_t sdr_pointwise_mma_packed_bytes(int M, int Kc);
int sdr_pointwise_mma_pack(const float* W, int M, int Kc, void* packed, sdr_stream stream);
int sdr_pointwise_mma(const float* x, const sdr_norm_in* fin, const void* packed_w, const float* bias,
                      const float* residual, const float* gate, int gate_channels,
                      float* y, double* stats_out,
                      int samples, int M, int Kc, int L, int epilogue, sdr_stream stream);

/* depthwise Conv1d(k=5, padding=2, stride 1|2, groups=C) on the deferred-normalised
 * input (improved_sudormrf.py:178-189,206-211).  x [samples,C,Lin] ->
 * y [samples,C,Lout] raw + stats.                                            */
int sdr_depthwise(const float* x, const sdr_norm_in* fin, const float* w5, const float* bias,
                  float* y, double* stats_out,
                  int samples, int C, int Lin, int stride, sdr_stream stream);

/* The whole depthwise pyramid of a U-ConvBlock (improved_sudormrf.py:205-216) in one pass over the projection
 * output y [samples,C,L] (fin: its GlobLN + PReLU).  Levels d >= 1 are affine in their inputs, so the kernel
 * chains RAW stride-2 convolutions without waiting for any statistics: z[0] receives level 0's output, z[d]
 * (d >= 1) the raw chain R_d [samples,C,L>>d]; a second small kernel reproduces every level's GlobLN from row
 * statistics and leaves the coefficients of the merge in `scratch` (sdr_pyramid_scratch_bytes; 0 = shape not
 * eligible: D < 4, L % 16, L >> (D-1) < 6, or rows too long for shared memory -> use sdr_depthwise / sdr_merge).
 * w5[d] [C][5], bias[d] [C]: level d's depthwise conv; gamma[d] / beta[d] [C]: spp_dw[d].norm.
 * stats0: zeroed (sum, sumsq) slot per sample for level 0's output.
 * sdr_merge_pyramid then writes m[c,t] = sum_d GLN_d(z_d)[c, t>>d] (+ its statistics), reading z and scratch. */
size_t sdr_pyramid_scratch_bytes(int samples, int C, int D, int L);
int sdr_depthwise_pyramid(const float* y, const sdr_norm_in* fin, const float* const* w5, const float* const* bias,
                          const float* const* gamma, const float* const* beta, float* const* z, double* stats0,
                          void* scratch, int D, int samples, int C, int L, sdr_stream stream);
int sdr_merge_pyramid(const float* const* z, const void* scratch, int D, float* m, double* stats_out,
                      int samples, int C, int L, sdr_stream stream);

/* The depthwise stage of the causal U-ConvBlock (causal_improved_sudormrf_v3.py:106-116) in one pass: PReLU of
 * proj_1x1 on load (slope_in), D levels of [causally masked 21-tap depthwise conv (stride 1, then 2) + bias + PReLU]
 * kept in shared memory, nearest up-sampling and adds, m[c,t] = sum_d o_d[c, t>>d].  No normalisation layers exist
 * in this block, so nothing crosses CTAs.  y, m [samples,C,L]; w21[d] [C][1][21] in the reference's layout (the 10
 * taps the causal mask zeroes, :21-27, are not read); bias[d] [C]; slope_in, slope[d]: one float each.
 * L % 4 == 0 and L % 2^D == 0 (every padded length is), else SDR_ERR_UNSUPPORTED.                              */
int sdr_causal_pyramid(const float* y, const float* slope_in, const float* const* w21, const float* const* bias,
                       const float* const* slope, float* m, int D, int samples, int C, int L, sdr_stream stream);

/* nearest x2 up-sampling + skip adds, closed form
 * m[c,t] = sum_d norm_d(z_d)[c, t>>d] (improved_sudormrf.py:214-216).        */
int sdr_merge(const float* const* z, const sdr_norm_in* fins, int depth,
              float* m, double* stats_out, int samples, int C, int L, sdr_stream stream);

/* TAC (groupcomm_sudormrf_v2.py:356-384) up to (not including) its GlobLN:
 * x [B,G,n,L] -> o [B,G,n,L] raw + stats per (b,g).  params = the 9 TAC
 * tensors before TAC_norm, state_dict order.                                 */
int sdr_tac(const float* x, const float* const* params, float* o, double* stats_out,
            int B, int G, int n, int L, sdr_stream stream);

/* ConvTranspose1d overlap-add + crop (+ uniform mixture consistency):
 * frames [B, SA*K, L] -> out [B, SA, T] (improved_sudormrf.py:272-279,300-301). */
int sdr_overlap_add(const float* frames, const float* mix_or_null, float* out,
                    int B, int SA, int K, int L, int64_t T, sdr_stream stream);

/* ---- stages that only the original model has (sudormrf.py) ---------------- */

/* Tail of the original UBlock (sudormrf.py:184-186) up to the statistics module_act needs:
 *   x[b,c,l] <- GN_e(e)[b,c,l] + f(x[b,c,l])      in place, (sum, sumsq) of the new x into stats_out
 * e [samples,C,L]: raw conv_1x1_exp.conv output, fe its GroupNorm (stats + weight/bias, no activation);
 * fx: how the block input is read: NULL stats = plain tensor (first block), else module_act of the previous
 * block (GroupNorm + per-channel PReLU) applied on load.                                                   */
int sdr_residual_norm(const float* e, const sdr_norm_in* fe, float* x, const sdr_norm_in* fx, double* stats_out,
                      int samples, int C, int L, sdr_stream stream);

/* Masks of the original model (sudormrf.py:285-289): logits [B,S,N,L] (the (N+1)x1 Conv2d output) ->
 * softmax over the S sources (sigmoid when S == 1) times the encoder output enc [B,N,L]; in place is allowed. */
int sdr_softmax_gate(const float* logits, const float* enc, float* out, int B, int S, int N, int L, sdr_stream stream);

/* ---- the steps either side of the forward -------- */

/* Per-row (mean, unbiased std) of a waveform batch: wav [rows, T] ->
 * mean_std [rows][2] fp32 (README.md:101-102: `x.mean(-1)`, `x.std(-1)`).
 * scratch: rows * 2 doubles of device memory.                                */
int sdr_utterance_stats(const float* wav, float* mean_std, int rows, int64_t T,
                        void* scratch, sdr_stream stream);

/* The README inference recipe as ONE call (README.md:100-114):
 *   m, s = wav.mean(-1), wav.std(-1);  x = (wav - m) / (s + 1e-9)
 *   est  = model(x.unsqueeze(1)) * s + m
 *   [est = mixture_consistency.apply(est, x.unsqueeze(1))]      (uniform)
 * wav [B,1,T] -> out [B,S,T].  Mono models only (in_audio_channels == 1).
 * Workspace: sdr_separate_workspace_bytes (forward workspace + the
 * normalised copy of the batch).                                             */
size_t sdr_separate_workspace_bytes(const sdr_config* cfg, int B, int64_t T);
int sdr_separate(const sdr_config* cfg, const void* packed, const float* wav, float* out,
                 int B, int64_t T, int apply_mixture_consistency,
                 void* workspace, size_t workspace_bytes, sdr_stream stream);

/* The same recipe for a RAGGED batch: B utterances of different lengths that
 * share one padded length T (a multiple of hop * 2^depth, sdr_padded_length),
 * stored zero-padded as wav [B,1,T]; lengths[b] (device, int64) = true length.
 * Statistics use the first lengths[b] samples only, the padding stays zero,
 * i.e. every row is computed exactly as the reference computes that utterance
 * alone (improved_sudormrf.py:303-318 pads it to the same T).  rescale = 0
 * skips `est * std + mean` (utils/simple_whamr_evaluation.py:141-148 evaluates
 * the normalised estimates).  out [B,S,T]: row b is valid up to lengths[b].   */
int sdr_separate_ragged(const sdr_config* cfg, const void* packed, const float* wav,
                        const int64_t* lengths, float* out, int B, int64_t T,
                        int apply_mixture_consistency, int rescale,
                        void* workspace, size_t workspace_bytes, sdr_stream stream);

/* Permutation-invariant SI-SDR of a batch (dnn/losses/sisdr.py:66-194,
 * PermInvariantSISDR.forward with return_individual_results=True,
 * backward_loss=False): est, target [B,S,T], mixture [B,1,T] (needed only for
 * improvement != 0) -> best[b] = max over permutations of the source-mean
 * SI-SNR (minus the batch-mean SI-SNR of the mixture when improvement != 0),
 * perm_index[b] = index of that permutation in itertools.permutations(range(S))
 * order.  1 <= S <= 4.  eps as in the reference's forward (default 1e-9).     */
/* Pairwise negative SNR / SI-SDR / SD-SDR of a batch (dnn/losses/sisdr.py:372-457, PairwiseNegSDR.forward):
 * out[b, i, j] = -sdr(estimate i, target j), [B, S, S] fp32; sdr_type 0 = "snr", 1 = "sisdr", 2 = "sdsdr".
 * One fp64 Gram pass + a finalize kernel; scratch: sdr_pit_sisdr_scratch_bytes(B, S).  1 <= S <= 4.        */
int sdr_pairwise_neg_sdr(const float* est, const float* target, float* out, int B, int S, int64_t T, int sdr_type,
                         int zero_mean, int take_log, void* scratch, sdr_stream stream);

size_t sdr_pit_sisdr_scratch_bytes(int B, int S);
int sdr_pit_sisdr(const float* est, const float* target, const float* mixture_or_null,
                  float* best, int32_t* perm_index, int B, int S, int64_t T,
                  int zero_mean, int improvement, double eps,
                  void* scratch, sdr_stream stream);

/* StabilizedPermInvSISDRMetric.forward (dnn/losses/sisdr.py:460-591; backward_loss=False,
 * return_individual_results=True), the validation metric of run_fuss_separation.py:111-131: n_est estimated sources
 * scored against n_act <= n_est actual ones with the stabilised SI-SNR
 *     rho^2 = <e,t>^2 / (<e,e><t,t> + eps),  10 log10((rho^2 + eps) / (1 - rho^2 + eps)),
 * best[b] = max over itertools.permutations(range(n_est), r=n_act) of the source mean (minus, for improvement != 0,
 * the batch mean of the same figure for the mixture = sum of the targets), perm_index[b] = index of that assignment.
 * est [B, est_rows, T], target [B, n_act, T]; est_rows == n_est, or est_rows > n_est == 1: single_source mode, the
 * rows are summed first (:576-577).  1 <= n_act <= n_est <= 4.  eps as in the reference's forward (1e-9).            */
size_t sdr_stabilized_sisdr_scratch_bytes(int B, int n_est, int n_act);
int sdr_stabilized_sisdr(const float* est, const float* target, float* best, int32_t* perm_index,
                         int B, int est_rows, int n_est, int n_act, int64_t T,
                         int zero_mean, int improvement, double eps, void* scratch, sdr_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* SUDORMRF_B200_H */
