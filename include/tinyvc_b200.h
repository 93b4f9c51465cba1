/* tinyvc_b200 -- C-ABI of the B200-native TinyVC inference path.
 *
 * The reference (uthree/tinyvc) is pure Python on torch and has no FFI of its own; the
 * boundary it offers is its Python class surface.  Each entry point below
 * replaces the arithmetic behind one of those Python callables -- cited as
 * /root/reference-relative file:line -- and is what `tinyvc_b200/tinyvc/*.py` (the drop-in
 * mirror of `module.tinyvc`) binds with ctypes.  INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - plain C: opaque handles, raw pointers, sizes; no torch types.
 *   - every tensor pointer is a DEVICE pointer to contiguous fp32 in the reference's
 *     channels-first layout [B][C][T] unless stated otherwise; the caller owns all buffers.
 *   - `workspace` is caller-provided scratch (device, 256-byte aligned) of at least
 *     `*_workspace_bytes(...)` bytes; no entry point allocates on the hot path.
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*); calls are asynchronous.
 *   - return value 0 = ok; otherwise tvc_last_error() gives a thread-local message.  No entry
 *     point throws, and none falls back to the CPU.
 *   - frame = 480 samples @ 24 kHz; an utterance has Lf frames and L = 480*Lf samples.
 */
#ifndef TINYVC_B200_H
#define TINYVC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tvc_decoder* tvc_decoder_t;
typedef struct tvc_encoder* tvc_encoder_t;
typedef struct tvc_index* tvc_index_t;

const char* tvc_last_error(void);
const char* tvc_version(void);
/* Runtime switches.  ("conv_impl","tc"|"fp32"): Decoder.infer on the tcgen05 tensor-core plan (default)
 * or on the exact-fp32 CUDA-core plan.  ("encoder_impl","tc"|"fp32"): the same choice for tvc_encoder_forward;
 * ("pitch_impl","fp32"|"tc"): its PitchEstimator stack alone (default fp32: f0 feeds the oscillator's phase integrator).
 * ("graphs","1"|"0"): tvc_decoder_infer replays a captured CUDA
 * graph when it is called again with the same buffers (default on).  ("fused_up","1"|"0"): the 24-channel
 * Upsample block as one fused kernel (default) or five conv launches.  ("pdl","0"|"1"): programmatic dependent
 * launch (default on).  ("profile","0"|"1"): per-launcher event timing.  ("nvtx","0"|"1"): NVTX ranges per launcher.
 * Returns non-zero for unknown keys.                                                                */
int tvc_set_option(const char* key, const char* value);
/* Number of kernels this library has launched in this process (bench.py's `gpu_launches`). */
unsigned long long tvc_launch_count(void);
/* With option ("profile","1"): per-launcher CUDA-event times since the last report, as a JSON
 * object {"name": {"launches": n, "ms": total}} written to buf.  Synchronises the device.      */
int tvc_profile_report(char* buf, size_t n);
/* Measurement aid (bench.py's `fp32_flop_fraction` denominator): sustained CUDA-core FP32 FMA rate of the current
 * device in TFLOP/s, from a register-resident FMA micro-benchmark timed with CUDA events.            */
int tvc_measure_fp32_peak(double* tflops, void* stream);
/* Multi-GPU (tinyvc_b200/shard.py, one process per GPU, one box): a result buffer on the owner's GPU that the kernels of
 * every other rank's GPU store into directly over NVLink / NVSwitch (tvc_decoder_infer's last kernel writes its waveform
 * block into it, so the sharded mode has no gather step).
 *   tvc_peer_alloc : owner.  cudaMalloc on the current device + its CUDA IPC handle (64 bytes, to be sent to the peers).
 *   tvc_peer_open  : peer.   Maps the owner's buffer into the CURRENT device's address space (cudaIpcOpenMemHandle with lazy
 *                    peer access), which is what makes it addressable by this device's kernels.
 *   tvc_peer_close / tvc_peer_free : undo the two.                                                                     */
int tvc_peer_alloc(size_t bytes, void** ptr, unsigned char handle[64]);
int tvc_peer_open(const unsigned char handle[64], void** ptr);
int tvc_peer_close(void* ptr);
int tvc_peer_free(void* ptr);

/* ---- parameter contract: flat fp32 buffers in torch state_dict() order ------------------- */
/* kind: 0 = Decoder (module/tinyvc/decoder.py:236-251), 1 = Encoder (encoder.py:100-106).     */
int tvc_param_count(int kind);
const char* tvc_param_name(int kind, int i);
int64_t tvc_param_numel(int kind, int i);
int64_t tvc_param_total(int kind);

/* ---- Decoder (module/tinyvc/decoder.py) --------------------------------------------------- */
/* `params`: host or device pointer to tvc_param_total(0) floats (replaces
 * decoder.load_state_dict, infer.py:36-37).  Weights are repacked once into kernel layout.   */
int tvc_decoder_create(const float* params, int64_t numel, tvc_decoder_t* out);
int tvc_decoder_destroy(tvc_decoder_t h);
/* Scratch bytes for this shape.  tvc_decoder_workspace_bytes fits every decoder entry point under any option;
 * tvc_decoder_infer_workspace_bytes is what tvc_decoder_infer alone needs under the options in force when it is called
 * (tvc_set_option: conv_impl, fused_up, fuse_down, pad limits -- ask again after changing one): the tensor-core plan
 * needs 540 B per output sample, the exact-fp32 plan 862.                                                              */
size_t tvc_decoder_workspace_bytes(int B, int Lf);
size_t tvc_decoder_infer_workspace_bytes(int B, int Lf);

/* Decoder.infer(content, f0, energy)  (decoder.py:253-257).
 *   content [B,768,Lf]  f0 [B,1,Lf]  energy [B,1,L]  rand01 [B,961,Lf]  ->  out [B,L]
 * `rand01` is the uniform [0,1) draw that decoder.py:78 takes from torch's generator; passing
 * it in makes the noise branch reproducible against the CPU reference.  NULL: the draw is made
 * inside the noise kernel (Philox-4x32-10 keyed by tvc_decoder_seed, one fresh tensor per call). */
int tvc_decoder_infer(tvc_decoder_t h, const float* content, const float* f0, const float* energy,
                      const float* rand01, float* out, int B, int Lf, void* workspace,
                      size_t workspace_bytes, void* stream);

/* The same with output pruning: the caller reads only out[b][out_t0, out_t1) of every utterance (StreamInfer.audio_callback
 * keeps y[-9600:-3840] of its 13 440-sample window, module/infer/stream.py:75).  The fused full-rate block skips the windows
 * that produce none of those samples; the samples inside the range are bit-identical to tvc_decoder_infer's, the samples
 * outside it are left unspecified.  0 <= out_t0 < out_t1 <= 480 * Lf.                                                     */
int tvc_decoder_infer_range(tvc_decoder_t h, const float* content, const float* f0, const float* energy,
                            const float* rand01, float* out, int B, int Lf, int64_t out_t0, int64_t out_t1,
                            void* workspace, size_t workspace_bytes, void* stream);

/* Test / planning aid (host arithmetic only, no GPU work): the rows [wa[i], wb[i]) of every utterance that Upsample level i of the
 * FilterNet (i = 0 .. 4: rates L/240, L/80, L/20, L/5, L) works on when tvc_decoder_infer_range is asked for [out_t0, out_t1).
 * Level 4 (the fused block) always reports its full length: it prunes by 426-sample windows itself.                        */
int tvc_decoder_plan_windows(int Lf, int64_t out_t0, int64_t out_t1, int32_t* wa, int32_t* wb);

/* Seed of the in-kernel noise draw (the role torch.manual_seed plays for decoder.py:78).      */
int tvc_decoder_seed(tvc_decoder_t h, uint64_t seed, void* stream);

/* SourceNet.forward (decoder.py:126-134): -> amps [B,15,Lf], kernel [B,961,Lf].              */
int tvc_source_net(tvc_decoder_t h, const float* content, const float* f0, const float* energy,
                   float* amps, float* kernel, int B, int Lf, void* workspace, size_t workspace_bytes,
                   void* stream);

/* Decoder.dsp (decoder.py:259-266; oscillate_harmonics :24-54, oscillate_noise :63-85):
 *   f0 [B,1,Lf], amps [B,15,Lf], kernel [B,961,Lf], rand01 [B,961,Lf] -> source [B,16,L].     */
int tvc_dsp(tvc_decoder_t h, const float* f0, const float* amps, const float* kernel,
            const float* rand01, float* source, int B, int Lf, void* workspace, size_t workspace_bytes,
            void* stream);

/* FilterNet.forward (decoder.py:222-233): source [B,16,L] -> out [B,1,L].                    */
int tvc_filter_net(tvc_decoder_t h, const float* content, const float* f0, const float* energy,
                   const float* source, float* out, int B, int Lf, void* workspace,
                   size_t workspace_bytes, void* stream);

/* oscillate_harmonics phase only (decoder.py:39-50): theta [B,15,L]; parity probe.           */
int tvc_harmonic_theta(const float* f0, float* theta, int B, int Lf, void* stream);

/* ---- Encoder (module/tinyvc/encoder.py) --------------------------------------------------- */
int tvc_encoder_create(const float* params, int64_t numel, tvc_encoder_t* out);
int tvc_encoder_destroy(tvc_encoder_t h);
size_t tvc_encoder_workspace_bytes(int B, int Lf);

/* Encoder.forward / Encoder.infer (encoder.py:108-116).  spec [B,961,Lf] ->
 *   z [B,768,Lf] (SSLFeatureEstimator, :89-97), logits [B,512,Lf] (PitchEstimator.forward,
 *   :33-38) and f0 [B,1,Lf] (PitchEstimator.decode, :61-67).  Any output pointer may be NULL. */
int tvc_encoder_forward(tvc_encoder_t h, const float* spec, float* z, float* logits, float* f0,
                        int B, int Lf, void* workspace, size_t workspace_bytes, void* stream);

/* PitchEstimator.decode (encoder.py:61-67) on caller-provided logits [B,512,Lf] (k = 4).     */
int tvc_pitch_decode(const float* logits, float* f0, int B, int Lf, void* stream);

/* ---- kNN content match (module/tinyvc/feature_retrieval.py:15-33) ------------------------- */
/* metric: 0 = 'cos', 1 = 'IP', 2 = 'L2'.  `index` is the index.pt tensor [1,768,N]
 * (extract_index.py:58), device pointer; a normalised copy and a row-major copy are built once. */
int tvc_index_create(const float* index, int N, int metric, tvc_index_t* out);
int tvc_index_destroy(tvc_index_t h);
size_t tvc_match_workspace_bytes(tvc_index_t h, int B, int Lf);
/* match_features(source, reference, k, alpha, metrics): source [B,768,Lf] -> out [B,768,Lf];
 * idx_out (nullable) int32 [B,Lf,k] receives topk indices in descending-similarity order.    */
int tvc_match_features(tvc_index_t h, const float* source, float* out, int32_t* idx_out, int B, int Lf,
                       int k, float alpha, void* workspace, size_t workspace_bytes, void* stream);

/* ---- signal front end (module/utils) ------------------------------------------------------ */
/* spectrogram (utils/spectrogram.py:8-15): wf [B,L] (L multiple of 480) -> spec [B,961,Lf].  */
size_t tvc_spectrogram_workspace_bytes(int B, int L);
int tvc_spectrogram(const float* wf, float* spec, int B, int L, void* workspace, size_t workspace_bytes,
                    void* stream);
/* estimate_energy (utils/energy_estimation.py:9-14): wf [B,L] -> energy [B,1,L].             */
size_t tvc_energy_workspace_bytes(int B, int L);
int tvc_estimate_energy(const float* wf, float* energy, int B, int L, void* workspace, size_t workspace_bytes,
                        void* stream);
/* Sample-rate conversion of input files (infer.py:45-46,63-64: torchaudio.functional.resample(wf, sr, 24000) with its
 * defaults: sinc_interp_hann, lowpass_filter_width 6, rolloff 0.99).  wf [B,L] -> out [B, tvc_resample_length(L, ...)]
 * (= ceil(new_freq * L / orig_freq)); the polyphase filter bank is built once per (device, orig_freq, new_freq).        */
int64_t tvc_resample_length(int64_t L, int orig_freq, int new_freq);
int tvc_resample(const float* wf, float* out, int B, int64_t L, int orig_freq, int new_freq, void* stream);
/* shift_frequency (utils/pitch_shift.py:5-15): n elements, in place allowed.                 */
int tvc_shift_frequency(const float* f0, float* out, int64_t n, float semitones, void* stream);

/* ---- streaming (module/infer/stream.py:68-95) --------------------------------------------- */
/* SOLA search + cross-fade for S independent streams.  y [S,y_len] is the converted window;
 * sola_buf [S,cross] is read and replaced; out_block [S,block]; shift_out int32 [S].
 * fade_in [cross] as built by StreamInfer.init_buffer (stream.py:61).                        */
int tvc_sola(const float* y, int y_len, float* sola_buf, const float* fade_in, float* out_block,
             int32_t* shift_out, int S, int block, int cross, int search, int delay, void* stream);

/* The same tick with the reference's `use_phase_vocoder=True` cross-fade (stream.py:83-89): the first
 * `cross` output samples are phase_vocoder(old tail, new head, fade_out, fade_in) (stream.py:9-26)
 * instead of the linear cross-fade.  Needs block >= cross.                                         */
size_t tvc_sola_pv_workspace_bytes(int S, int cross);
int tvc_sola_pv(const float* y, int y_len, float* sola_buf, const float* fade_in, float* out_block,
                int32_t* shift_out, int S, int block, int cross, int search, int delay, void* workspace,
                size_t workspace_bytes, void* stream);
/* phase_vocoder(a, b, fade_out = 1 - fade_in, fade_in) (stream.py:9-26) for S independent pairs:
 * a, b, out [S,n]; fade_in [n].                                                                    */
size_t tvc_phase_vocoder_workspace_bytes(int S, int n);
int tvc_phase_vocoder(const float* a, const float* b, const float* fade_in, float* out, int S, int n,
                      void* workspace, size_t workspace_bytes, void* stream);

/* ---- parity probes (tests only) ------------------------------------------------------------ */
/* One dense Conv1d through the tcgen05 tensor-core kernel (csrc/tc_conv.cu) with channels-first fp32
 * device tensors at the boundary, as nn.Conv1d(Cin, Cout, K, dilation=dil, padding=dil*(K-1)/2,
 * padding_mode='replicate') evaluates it (decoder.py:143-146,165-171).  `w` [Cout][Cin][K], `bias`
 * [Cout], `aux_w`, `aux_b` are HOST pointers in torch layout.  aux_mode 0 = none; 1 = a 1x1 conv of
 * aux_x accumulated into the output (Downsample.down_res, decoder.py:143,157); 2 = FiLM
 * (decoder.py:88-97): aux_w [2][Cout][aux_cin] = to_scale rows then to_shift rows.  res (nullable)
 * is added last.  epi_act/out_act: 0 none, 1 leaky_relu(0.1), 2 GELU, 3 ELU+1.  y receives the fp32
 * output, y_planes the re-split bf16 hi+lo copy with out_act applied (both [B][Cout][T], nullable). */
int tvc_tc_conv_probe(const float* x, const float* w, const float* bias, int B, int T, int Cin, int Cout, int K,
                      int dil, const float* aux_x, const float* aux_w, const float* aux_b, int aux_cin,
                      int aux_mode, const float* res, int epi_act, int out_act, int NT, float* y, float* y_planes,
                      void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TINYVC_B200_H */
