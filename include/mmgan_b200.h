/* mmgan_b200 -- C ABI of the B200 (sm_100a) CUDA library behind the MM-GAN / GAN-DES hot path.
 *
 * The reference (marja-w/gan-des-midi-music-gen) has no FFI layer: its boundary is the Python API
 * (nn.Module / Dataset / generate_piano_roll).  The Python mirror in gan-des-midi-music-gen_b200/
 * keeps that API and calls these entry points through ctypes; a maintainer of the reference binds
 * them the same way.  Each entry point cites the reference code it replaces
 * (paths relative to the reference root).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the comment says HOST; the caller owns all buffers,
 *     including `workspace` (size from the matching *_workspace_bytes); kernels never allocate.
 *   - `stream` is a cudaStream_t (CUstream) passed as void*; calls are asynchronous, stream-ordered,
 *     re-entrant across streams, and never synchronise the host.
 *   - return value: 0 ok; <0 invalid argument (-1), unsupported shape (-2), workspace too small (-3);
 *     >0 a cudaError_t.  mmg_last_error() returns a thread-local message.  Nothing throws.
 *   - tensors are contiguous: activations NCHW or (rows, features); Linear weight (out,in);
 *     Conv2d weight (Co,Ci,kh,kw); ConvTranspose2d weight (Ci,Co,kh,kw).
 *   - `act`: 0 none, 1 LeakyReLU(0.2), 2 ReLU, 3 sigmoid.
 */
#ifndef MMGAN_B200_H
#define MMGAN_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

int mmg_abi_version(void);
const char* mmg_last_error(void);
uint64_t mmg_launch_count(void);      /* kernels launched by this library so far (process-wide) */

/* ---- piano-roll rasteriser: MMGAN_MIDI_DES/datasets.py:27-54 (generate_piano_roll raster core) ----
 * Events are the post-mido stream: dt[i] seconds (float64), meta[i] = kind | pitch<<8 | velocity<<16 with
 * kind 0 other / 1 note_on / 2 note_off; song s owns events [offsets[s], offsets[s+1]).
 * sequence_length < 0 means None (end+20, datasets.py:14-15).  out: (n_songs, 2, 128, Wout), plane 0 =
 * piano_roll, plane 1 = durations; out_dtype 0 float32, 1 bfloat16, 2 uint8 (saturating).
 * status (may be NULL): per song, bit0 = negative time step met (dt < 0, outside the contract),
 * bit1 = note with pitch >= 128 met (the reference's IndexError path). */
int mmg_raster_out_width(int start, int end);                          /* datasets.py:49-54 re-slice */
size_t mmg_raster_workspace_bytes(int64_t n_songs, int64_t total_events);
/* time steps of the workspace path: 0 (default) speculate-and-verify parallel prefix sums, the exact sequential chain only for the songs that
 * need it (their count accumulates in the last 8 bytes of the workspace, uint64); 1 = the sequential chain for every song.  Bit-exact both. */
int mmg_raster_set_mode(int mode);
int mmg_raster_piano_roll(const double* dt, const uint32_t* meta, const int64_t* offsets, int64_t n_songs,
                          int64_t total_events, int sequence_length, int start, int end, int out_dtype, void* out,
                          int32_t* status, void* workspace, size_t ws_bytes, void* stream);

/* ---- Standard MIDI File reader (host code, no device work): what `for msg in mido.MidiFile(path)` yields, datasets.py:18,34 (mido 1.3.2:
 * merge_tracks by absolute tick, stable; end_of_track metas dropped and one re-appended at the last tick; delta seconds with the RUNNING tempo,
 * ticks * (tempo * 1e-6 / ticks_per_beat)).  data / len = the file image.  Outputs for message i: dt[i] seconds, meta[i] in the rasteriser's
 * record format (above), abs_tick[i] (may be NULL); capacity >= mmg_smf_max_messages(len) always suffices.  tempo_tick / tempo_us (may be NULL):
 * the set_tempo messages in stream order (for the beat grid); *n_tempo their count.  -1: not an SMF / malformed (the message says where),
 * -2: SMPTE division or a type-2 file (mido raises for those too), -3: capacity too small. */
int64_t mmg_smf_max_messages(size_t len);
int mmg_smf_parse(const unsigned char* data, size_t len, double* dt, uint32_t* meta, int64_t* abs_tick, int64_t capacity, int64_t* n_messages,
                  int* ticks_per_beat, int64_t* tempo_tick, int32_t* tempo_us, int64_t tempo_capacity, int64_t* n_tempo);
/* quarter-note beat grid along that tempo map up to last_tick (stand-in for pretty_midi.get_beats, datasets.py:57): last_tick / ticks_per_beat + 1
 * values (= *n_beats; -3 when capacity is smaller) */
int mmg_smf_beat_grid(const int64_t* tempo_tick, const int32_t* tempo_us, int64_t n_tempo, int ticks_per_beat, int64_t last_tick, double* beats,
                      int64_t capacity, int64_t* n_beats);

/* ---- simulator log -> note-event stream (host code, all cores): MMGAN_MIDI_DES/sim_log_to_midi.py:13-277 (MidiGenerator, LogLineProcessor,
 * process_adjsim_log up to the generate_piano_roll call) followed by mido's playback of the saved track (datasets.py:34) = the message stream the
 * rasteriser above consumes.  log = the DES log lines, '\n'-separated; instruments / note_levels = int(v) of the reference's arguments; gen2 = the beat
 * generator's output row (>= 6 values) as float32 (gen2_is_f32: products are formed in single precision like numpy float32 scalars) or float64;
 * generate as in process_adjsim_log.  dt / meta: room for mmg_simlog_max_messages() (512) messages per song.  -1 "Error in processing log file" where
 * the reference raises it (unknown server id, non-integer customer id, modulo by zero ...).  The batch form runs n_threads host threads (0 = all
 * cores), takes rows per song and returns the streams packed: song s at [offsets[s], offsets[s+1]). */
int mmg_simlog_max_messages(void);
int mmg_simlog_to_events(const char* log, size_t log_len, const int64_t* instruments, int n_instruments, const int64_t* note_levels, int n_note_levels,
                         const void* gen2, int gen2_len, int gen2_is_f32, int generate, double* dt, uint32_t* meta, int64_t capacity, int64_t* n_messages);
int mmg_simlog_batch_to_events(const char* logs, const int64_t* log_offsets, int64_t n_songs, const int64_t* instruments, int n_instruments,
                               const int64_t* note_levels, int n_note_levels, const void* gen2, int gen2_len, int gen2_is_f32, int generate, double* dt,
                               uint32_t* meta, int64_t* offsets, int n_threads);

/* ---- losses / optimiser ----
 * nn.BCEWithLogitsLoss() mean (network_tests.py:248,304-306,313; SIMNN.py:257): loss[0] (+)= mean(l_i);
 * dlogits = (sigmoid(x) - y) * gscale * (*gscale_dev if not NULL).  targets NULL -> constant target. */
int mmg_bce_logits_f32(const float* logits, const float* targets, float target_const, int64_t n, float* loss,
                       int accumulate, float* dlogits, float gscale, const float* gscale_dev, void* stream);
int mmg_fill_scalar_f32(float* dst, const float* src_dev, int64_t n, void* stream);   /* dst[:] = *src_dev */
int mmg_zero(void* dst, size_t bytes, void* stream);                                   /* cudaMemsetAsync(dst, 0, bytes) */
int mmg_sum_f32(const float* x, int64_t n, float* out, int accumulate, void* stream);  /* out[0] (+)= sum(x) */
/* dz = dy * act'(y), y = act(z) */
int mmg_act_bwd_f32(const float* y, const float* dy, float* dz, int64_t n, int act, void* stream);
/* torch.optim.Adam (network_tests.py:253-254,308,315; SIMNN.py:258-259,316,331), one vectorised launch per
 * <= 48 tensors.  ptrs: HOST array of 4*n_tensors device pointers [params | grads | exp_avg | exp_avg_sq];
 * sizes: HOST array.  step is the 1-based step count AFTER increment; grads are multiplied by grad_scale. */
int mmg_adam_multi_tensor_f32(int n_tensors, void* const* ptrs, const int64_t* sizes, float lr, float beta1,
                              float beta2, float eps, int64_t step, float grad_scale, void* stream);

/* the same update with [lr, beta1, beta2, eps] (hyper_dev, fp32) and the count of updates applied so far (step_dev, int64,
 * incremented by the call) in DEVICE memory, so the launch can be replayed from a CUDA graph; at most 48 tensors. */
int mmg_adam_multi_tensor_dev_f32(int n_tensors, void* const* ptrs, const int64_t* sizes, const float* hyper_dev,
                                  int64_t* step_dev, float grad_scale, void* stream);

/* ---- fp32 layers (full-precision path) ----
 * nn.Linear (network_tests.py:77,139,154; SIMNN.py:127-128): y = act(x.w^T + b) */
int mmg_linear_fwd_f32(const float* x, const float* w, const float* b, float* y, int64_t M, int64_t N, int64_t K,
                       int act, void* stream);
int mmg_linear_bwd_f32(const float* x, const float* w, const float* dy, float* dx, float* dw, float* db, int64_t M,
                       int64_t N, int64_t K, int accumulate, void* stream);
/* nn.BatchNorm1d / BatchNorm2d over (N, C, HW) fused with the following activation
 * (network_tests.py:78-79; SIMNN.py:85-87,104-108).  Training: batch mean / biased variance, running stats
 * updated with the unbiased variance (run_* may be NULL). */
size_t mmg_bn_workspace_bytes(int64_t C);
int mmg_bn_fwd_train_f32(const float* z, const float* gamma, const float* beta, float* run_mean, float* run_var,
                         float* y, float* save_mean, float* save_invstd, int64_t N, int64_t C, int64_t HW,
                         float momentum, float eps, int act, void* workspace, size_t ws_bytes, void* stream);
int mmg_bn_fwd_eval_f32(const float* z, const float* gamma, const float* beta, const float* run_mean,
                        const float* run_var, float* y, int64_t N, int64_t C, int64_t HW, float eps, int act,
                        void* stream);
int mmg_bn_bwd_f32(const float* z, const float* dy, const float* gamma, const float* beta, const float* save_mean,
                   const float* save_invstd, float* dz, float* dgamma, float* dbeta, int64_t N, int64_t C,
                   int64_t HW, int act, int accumulate, void* workspace, size_t ws_bytes, void* stream);
/* nn.Conv2d (network_tests.py:150-151; SIMNN.py:123-124) and, through the data gradient,
 * nn.ConvTranspose2d (SIMNN.py:70-84). */
int mmg_conv2d_fwd_f32(const float* x, const float* w, const float* b, float* y, int N, int Ci, int H, int W, int Co,
                       int kh, int kw, int stride, int pad, int act, void* stream);
int mmg_conv2d_bwd_data_f32(const float* dy, const float* w, const float* b, float* dx, int N, int Ci, int H, int W,
                            int Co, int kh, int kw, int stride, int pad, int act, void* stream);
int mmg_conv2d_bwd_weight_f32(const float* x, const float* dy, float* dw, float* db, int N, int Ci, int H, int W,
                              int Co, int kh, int kw, int stride, int pad, int accumulate, void* stream);
/* nn.MaxPool2d(2,2) (SIMNN.py:125,138-139) */
int mmg_maxpool2_fwd_f32(const float* x, float* y, uint8_t* idx, int64_t NC, int H, int W, void* stream);
int mmg_maxpool2_bwd_f32(const float* dy, const uint8_t* idx, float* dx, int64_t NC, int H, int W, void* stream);

/* ---- GAN-DES contractions on the tensor cores (GAN_DES/SIMNN.py:70-84 ConvTranspose stack, :123-142 conv / fc layers) ----
 * mmg_gemm_tc: C[m][n] (+)= sum_k A(m,k) B(n,k), fp32 accumulation in TMEM (tcgen05.mma fed by TMA).  A is K-major ([M][K], lda =
 * row pitch in elements) or MN-major ([K][M]); B likewise.  dtype 0 = bf16, 1 = fp32 read as tf32 (K-major operands only).  Bases and row
 * pitches must be 16-byte aligned; M / N / K tails are zero-filled by TMA.  trans_out: element (m, n) is stored at
 * C[(m / inner) * N * inner + n * inner + m % inner] (inner = pixels per image gives NCHW, inner = M a plain transpose), else at
 * C[m * ldc + n]; trans_out = 2: C is a bf16 row-major matrix (plain stores only).  split_k > 1 needs atomic = 1 and a zeroed C (bias / act then through mmg_bias_act_inplace_f32). */
int mmg_gemm_tc(const void* A, int a_mn, long long lda, const void* B, int b_mn, long long ldb, float* C, long long ldc, int M, int N, int K, int dtype,
                int split_k, int trans_out, long long inner, int atomic, const float* bias, int bias_on_m, int act, void* stream);
/* fp32 (d0,d1,d2) -> bf16 dst[i_pa * a_stride + i_pb * pitch + i_pc] with the dimensions permuted to (pa, pb, pc); columns [n_pc, pitch)
 * are zeros; a_stride = 0 means n_pb * pitch */
int mmg_pack_bf16(const float* src, void* dst, int d0, int d1, int d2, int pa, int pb, int pc, long long pitch, long long a_stride, void* stream);
/* NCHW fp32 -> bf16 rows [B*OH*OW][pitch]: columns (c, ky, kx), column Ci*kh*kw = 1 when ones_col (bias / bias gradient ride the GEMMs) */
int mmg_im2col_bf16(const float* x, void* col, int B, int Ci, int H, int W, int kh, int kw, int stride, int pad, int pitch, int ones_col, void* stream);
/* transposed convolution, second half: col fp32 [B*Hin*Win][ldc >= Co*kh*kw] -> act(y) NCHW [B][Co][Hout][Wout], gather form (no atomics) */
int mmg_col2im_f32(const float* col, float* y, int B, int Co, int Hin, int Win, int kh, int kw, int stride, int pad, long long ldc, int act, void* stream);
int mmg_transpose_f32(const float* src, float* dst, int rows, int cols, long long lds, void* stream);
int mmg_colsum_f32(const float* src, float* dst, int rows, int cols, void* stream);                 /* dst[c] = sum_r src[r][c] */
int mmg_bias_act_inplace_f32(float* y, const float* bias, long long rows, int cols, int act, void* stream);
/* backward of MaxPool2d(2,2)(ReLU(z)) in one pass (SIMNN.py:138-139): dyp / idx / yp are the pooled gradient, the argmax codes of
 * mmg_maxpool2_fwd_f32 and the pooled output; writes dz fp32 NCHW (B,C,H,W) and / or dzt bf16 [C][Pp >= B*H*W] (either may be NULL) */
/* MaxPool2d(2,2)(ReLU(conv2d(x, w, b, stride 1, pad))) for a tiny stencil in ONE kernel (SIMNN.py:123,138: Conv2d(1, 16, kernel_size=2) has
 * K = 4: a stencil, not a GEMM).  Built for Ci = 1, 2 x 2, Co <= 32; anything else returns MMG_EUNSUPPORTED (use mmg_im2col_bf16 + mmg_gemm_tc).
 * yp / idx: (B, Co, OH/2, OW/2) pooled output and argmax codes (as mmg_maxpool2_fwd_f32), OH = H + 2 pad - kh + 1. */
int mmg_conv_small_relu_pool_f32(const float* x, const float* w, const float* bias, float* yp, uint8_t* idx, int B, int Ci, int H, int W, int Co, int kh,
                                 int kw, int pad, void* stream);
int mmg_pool_relu_bwd(const float* dyp, const uint8_t* idx, const float* yp, float* dz, void* dzt, void* dzn, int B, int C, int H, int W, long long Pp,
                      void* stream);
/* second half of a stride-1 convolution's data gradient: the tap columns dcol[b*OH*OW + p][(ky*kw + kx)*Ci + ci] (bf16, produced by mmg_gemm_tc
 * with trans_out = 2 from dz in NHWC rows and the weights as [(ky,kx,ci)][oc]) -> dx fp32 NCHW (B, Ci, H, W), gather form.  Built for Ci = 16. */
int mmg_conv_dgrad_gather(const void* dcol, float* dx, int B, int Ci, int H, int W, int kh, int kw, int pad, long long ldc, void* stream);

/* ---- GAN-DES mel front end (GAN_DES/util.py:37-61: torchaudio MelSpectrogram + AmplitudeToDB) ----
 * mmg_stft_power_f32: wave (B, L) fp32 (row pitch wave_pitch) -> power [B*T][pitch >= 1025] fp32, T = 1 + L / hop: centred frames with
 * reflect padding, periodic Hann window, 2048-point FFT, |X|^2.  Only n_fft = 2048.  The mel projection is mmg_gemm_tc (dtype 1, tf32) with
 * the transposed filter bank as B and trans_out / inner = T; mmg_power_to_db_f32 then gives 10 log10(max(x, 1e-10)) floored at the
 * spectrogram's maximum - top_db (top_db < 0: no floor) for n_spectrograms blocks of n values. */
int mmg_stft_power_f32(const float* wave, int B, long long L, long long wave_pitch, int n_fft, int hop, float* power, int pitch, void* stream);
int mmg_power_to_db_f32(const float* x, float* out, int n_spectrograms, long long n, float top_db, void* stream);

/* ---- bf16 tensor-core discriminator (DiscriminatorCNN, network_tests.py:147-160 and its autograd backward) ----
 * Activations live in padded space-to-depth layouts (see csrc/disc_tc.cu): XS (B*1690, 8) is the input, P1 (B*429, 64)
 * the conv1 activations, A2 / DZ2 (B*429, 32) the conv2 activations / their gradient, DZ1C (B*1690, 16) the conv1
 * pre-activation gradient (junk rows zero: allocate zeroed); all bf16.  x is (B,2,128,50) uint8 (x_dtype 2) or float32 (0).
 * `packed` = mmg_disc_packed_weights_bytes() bytes filled by mmg_disc_pack_weights from the fp32 nn.Parameters.
 * P1's pad cells must be zero (allocate zeroed, reuse).  logits must be initialised (fc bias) before conv2_fwd.
 * Gradient outputs are fp32, in the reference's parameter layouts, and are ACCUMULATED into (+=). */
size_t mmg_disc_packed_weights_bytes(void);
int mmg_disc_pack_weights(const float* conv1_w, const float* conv2_w, const float* fc_w, void* packed, void* stream);
int mmg_disc_xs_pack(const void* x, int x_dtype, void* xs, int64_t B, void* stream);
int mmg_disc_conv1_fwd(const void* xs, const void* packed, const float* conv1_b, void* p1, int64_t B, void* stream);
int mmg_disc_conv2_fwd(const void* p1, const void* packed, const float* conv2_b, void* a2, float* logits, int64_t B, void* stream);
int mmg_disc_fc_bwd(const void* a2, const float* dlogit, const void* packed, void* dz2, float* dfc_w, float* dconv2_b, int64_t B, void* stream);
int mmg_disc_conv2_wgrad(const void* p1, const void* dz2, float* dconv2_w, int64_t B, void* stream);
int mmg_disc_conv2_dgrad(const void* dz2, const void* packed, const void* p1, void* dz1c, float* dconv1_b, int64_t B, void* stream);
int mmg_disc_conv1_wgrad(const void* xs, const void* dz1c, float* dconv1_w, int64_t B, void* stream);

/* The whole backward of one pass in ONE persistent kernel (csrc/disc_tc_fused.cu): same inputs as the four calls above
 * (xs, p1, a2 as left by the forward, dlogit), dz2 / dz1c stay in shared memory, all six gradients are accumulated (+=). */
int mmg_disc_bwd_fused(const void* xs, const void* p1, const void* a2, const float* dlogit, const void* packed, float* dconv1_w,
                       float* dconv1_b, float* dconv2_w, float* dconv2_b, float* dfc_w, float* dfc_b, int64_t B, void* stream);

/* The whole forward of one pass in ONE persistent kernel (csrc/disc_tc_fused.cu): xs_pack + conv1_fwd + conv2_fwd with P1 kept in
 * shared memory between the two convolutions; writes xs / p1 / a2 for the backward and logits (fc bias included). */
int mmg_disc_fwd_fused(const void* x, int x_dtype, const void* packed, const float* conv1_b, const float* conv2_b, const float* fc_b,
                       void* xs, void* p1, void* a2, float* logits, int64_t B, void* stream);
/* The same with a gather: sample b of the pass is row x_index[b] (int64, device memory) of x, so the real rolls are read straight out of
 * an HBM-resident training set by sampler index (what DataLoader(MaestroDatasetPickle(..., device), shuffle=True) delivers,
 * network_tests.py:229-230,281) without materialising the batch.  x_index == NULL: rows 0..B-1. */
int mmg_disc_fwd_fused_gather(const void* x, int x_dtype, const int64_t* x_index, const void* packed, const float* conv1_b, const float* conv2_b,
                              const float* fc_b, void* xs, void* p1, void* a2, float* logits, int64_t B, void* stream);

/* ONE persistent kernel for a whole discriminator pass of the loop body (csrc/disc_tc_pass.cu): DiscriminatorCNN.forward
 * (network_tests.py:156-160) -> nn.BCEWithLogitsLoss against the constant target of the pass (network_tests.py:248,304-306,313)
 * -> backward (network_tests.py:307,314).  dlogit[b] = (sigmoid(logit[b]) - target) / loss_rows needs nothing from other samples, so
 * XS / P1 / A2 / DZ2 / DZ1 never leave the SM: HBM traffic per sample is the 12.8 KB uint8 roll.  x / x_dtype / x_index as in
 * mmg_disc_fwd_fused_gather; x_rows > 0 = number of rows of x: an x_index entry outside [0, x_rows) reads row 0 and sets the int32 flag at
 * workspace + mmg_disc_pass_workspace_bytes() - 128 (torch.index_select would raise).  loss_rows = rows behind the loss mean (0 = B).  logits (B,) and loss (loss[0] += mean BCE of the pass) may be
 * NULL; the six fp32 gradients are ACCUMULATED (+=).  workspace: mmg_disc_pass_workspace_bytes() bytes, 128-byte aligned (per-CTA scratch
 * slots that stay in L2).  mmg_disc_pass_fused_dbg additionally takes a HOST-MAPPED int array of 4 * 148 progress words (NULL = off). */
size_t mmg_disc_pass_workspace_bytes(void);
int mmg_disc_pass_set_flags(int flags);   /* debugging aid, process-wide: bit 0 = biases added in the epilogues instead of by the bias MMAs */
int mmg_disc_pass_fused(const void* x, int x_dtype, const int64_t* x_index, int64_t x_rows, const void* packed, const float* conv1_b, const float* conv2_b,
                        const float* fc_b, float target, int64_t loss_rows, float* logits, float* loss, float* dconv1_w, float* dconv1_b,
                        float* dconv2_w, float* dconv2_b, float* dfc_w, float* dfc_b, void* workspace, size_t ws_bytes, int64_t B, void* stream);
int mmg_disc_pass_fused_dbg(const void* x, int x_dtype, const int64_t* x_index, int64_t x_rows, const void* packed, const float* conv1_b, const float* conv2_b,
                            const float* fc_b, float target, int64_t loss_rows, float* logits, float* loss, float* dconv1_w, float* dconv1_b,
                            float* dconv2_w, float* dconv2_b, float* dfc_w, float* dfc_b, void* workspace, size_t ws_bytes, int64_t B, void* stream,
                            int* dbg);

/* ---- bf16 tensor-core generator blocks ([Linear -> BatchNorm1d -> Sigmoid], network_tests.py:75-80, as used by Generator
 * :58-90 and BeatGenerator :93-123) ----
 * One call = one Linear layer as a tcgen05 GEMM (bf16 operands, fp32 accumulate) with the neighbouring BatchNorm + sigmoid
 * folded into its prologue / epilogue (csrc/gen_tc.cu).  Input rows are x0 (M,k0) [|| x1 (M,k1)] fp32:
 *   in_mode 0: used as they are (first layer; the torch.cat of network_tests.py:86,122 is never materialised);
 *   in_mode 1: x0 holds the PREVIOUS layer's pre-activations z; a = sigmoid(BN(z)) with batch statistics from in_sums
 *              (sum z | sum z^2, fp64 [2][K]) and in_gamma / in_beta; if update_running, in_run_mean / in_run_var are updated
 *              (momentum, unbiased variance) exactly once;
 *   in_mode 2: the same with the running statistics (eval).
 * w_packed = mmg_gen_pack_weight(weight (N,K)); bias (N,) fp32.  Outputs (any subset, NULL = skip):
 *   z_out (M,N) fp32 pre-activations;  out_sums fp64 [2][N] += column sums of z and z^2 (zero them before the first layer call
 *   of a forward);  y_out (M,N) fp32 = sigmoid(BN(z)) with out_mode 1 = batch statistics y_sums (complete sums of THIS layer,
 *   i.e. a previous call accumulated them) or 2 = running statistics; out_run_* updated if update_running and out_mode 1.
 * Train-mode statistics over M <= 1 rows fail with -1 and torch's "Expected more than 1 value per channel" message. */
typedef struct mmg_gen_layer_args {
    const float* x0; const float* x1; int k0; int k1;
    int in_mode; const double* in_sums; const float* in_gamma; const float* in_beta; float* in_run_mean; float* in_run_var;
    const void* w_packed; const float* bias; int N;
    float* z_out; double* out_sums; float* y_out;
    int out_mode; const double* y_sums; const float* out_gamma; const float* out_beta; float* out_run_mean; float* out_run_var;
    float momentum; float eps; int update_running; int64_t M;
    int64_t stat_count;   /* rows behind in_sums / y_sums; 0 = M.  Data parallel SyncBN: the caller all-reduces the sums and passes the global batch */
} mmg_gen_layer_args;
size_t mmg_gen_packed_weight_bytes(int N, int K);
int mmg_gen_pack_weight(const float* w, int N, int K, void* packed, void* stream);
int mmg_gen_layer_fwd(const mmg_gen_layer_args* args, void* stream);
int mmg_gen_set_worker_groups(int groups);   /* tuning aid, process-wide: 2 or 4 groups of four builder / epilogue warps per CTA of mmg_gen_layer_fwd (default 4); returns the previous value */
/* Batch statistics of a wide layer fed by a narrow one (the generators' 64 -> 4096 output block, network_tests.py:71,78) without a GEMM pass:
 * with a = sigmoid(BN(z_prev)) (the layer's bf16 input operand), s = sum_r a_r and G = sum_r a_r a_r^T, the column sums of z = a W^T + b are
 * w_n.s + M b_n and those of z^2 are w_n^T G w_n + 2 b_n w_n.s + M b_n^2 (fp64).  Writes out_sums [2][N] for the M local rows; K <= 64;
 * in_sums / stat_count as in mmg_gen_layer_args (in_mode 1).  workspace: mmg_gen_layer_stats_gram_workspace() bytes. */
size_t mmg_gen_layer_stats_gram_workspace(void);
int mmg_gen_layer_stats_gram(const float* z_prev, int64_t M, int K, const double* in_sums, int64_t stat_count, const float* in_gamma,
                             const float* in_beta, float eps, const float* weight, const float* bias, int N, double* out_sums, void* workspace,
                             size_t ws_bytes, void* stream);
/* The three hidden [Linear -> BatchNorm1d -> Sigmoid] blocks of a generator (network_tests.py:68-71,75-80 / :103-106) in ONE cooperative launch, train
 * mode with the local batch statistics (what mmg_gen_layer_fwd does in three launches): one CTA per 128-row tile keeps the pre-activations in TMEM,
 * a grid-wide barrier per layer completes the column sums.  x0 / x1 as in mmg_gen_layer_args (in_mode 0); w_packed[l] = mmg_gen_pack_weight of
 * layer l; sums[l] fp64 [2][N_l] and the 4-byte barrier word are ZEROED by the caller before every call; z_out (M, N[2]) fp32 = the last hidden
 * layer's pre-activations (input of the output block's mmg_gen_layer_fwd, in_mode 1); run_mean / run_var updated once if update_running.
 * gram_part (NULL = skip): workspace of mmg_gen_layer_stats_gram_workspace() bytes receiving the per-CTA Gram partials of the output block's
 * statistics -- follow with mmg_gen_layer_stats_gram_finish(nparts = ceil(M / 128), ...).  stat_count must be 0 or M.
 * mmg_gen_hidden_fused_supported: 1 when (M rows, k_in input features, widths N3[3]) fits (at most one row tile per SM, N_l <= 256, shared memory). */
typedef struct mmg_gen_hidden_args {
    const float* x0; const float* x1; int k0; int k1;
    const void* w_packed[3]; const float* bias[3]; int N[3];
    const float* gamma[3]; const float* beta[3]; float* run_mean[3]; float* run_var[3];
    double* sums[3];
    float* z_out; double* gram_part; unsigned int* barrier;
    float momentum; float eps; int update_running; int64_t M; int64_t stat_count;
} mmg_gen_hidden_args;
int mmg_gen_hidden_fused_supported(int64_t M, int k_in, const int* N3, int with_gram);
int mmg_gen_hidden_fused(const mmg_gen_hidden_args* args, void* stream);
int mmg_gen_layer_stats_gram_finish(int nparts, const float* weight, const float* bias, int N, int K, int64_t M, double* out_sums, void* workspace,
                                    size_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MMGAN_B200_H */
