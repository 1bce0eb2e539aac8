/*
 * roar_sup.h -- C ABI of libroar_sup.so: B200 (sm_100a) supplementary-data extraction for Roar.
 *
 * The reference (AshwinSankar17/Roar) has no FFI/plugin layer on this path: the arithmetic is
 * inline Python in TTSDataset / FilterbankFeatures calling torch.stft, librosa and
 * torch.special.gammaln on the CPU.  Each entry point below replaces one of those inline call
 * sites (cited as file:line relative to the reference root); INTEGRATION.md shows the ctypes
 * binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - plain C, no torch types; every buffer is owned by the caller (PyTorch tensors in practice);
 *     the library never allocates outputs and never frees inputs;
 *   - pointers named d_* are DEVICE pointers on the handle's device, everything else is host;
 *   - all work is enqueued on the given stream (a cudaStream_t passed as void*), no implicit sync;
 *   - return value 0 = success, negative = error (roar_sup_last_error() gives the text);
 *   - a handle is immutable after create: safe to use from several host threads with distinct
 *     streams and workspaces;
 *   - there is no CPU fallback: without a CUDA device roar_sup_create fails.
 *
 * Ragged batches are PACKED: utterance i occupies d_audio[sample_off[i] .. sample_off[i]+sample_len[i]).
 * Offsets that are multiples of 4 samples (16 B) enable the TMA bulk-copy path; any offset works.
 */
#ifndef ROAR_SUP_H
#define ROAR_SUP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ROAR_SUP_ABI_VERSION 2

enum roar_sup_window { ROAR_WIN_HANN = 0, ROAR_WIN_HAMMING = 1, ROAR_WIN_BLACKMAN = 2,
                       ROAR_WIN_BARTLETT = 3, ROAR_WIN_NONE = 4 };
enum roar_sup_log_mode { ROAR_LOG_NONE = 0, ROAR_LOG_CLAMP = 1, ROAR_LOG_ADD = 2 };
enum roar_sup_normalize { ROAR_NORM_NONE = 0, ROAR_NORM_PER_FEATURE = 1, ROAR_NORM_ALL_FEATURES = 2 };

enum roar_sup_error {
  ROAR_OK = 0,
  ROAR_ERR_INVALID_ARG = -1,
  ROAR_ERR_CUDA = -2,
  ROAR_ERR_NO_DEVICE = -3,
  ROAR_ERR_WORKSPACE = -4,
  ROAR_ERR_UNSUPPORTED = -5
};

/* One immutable configuration per handle.  Field meaning follows the reference's keyword
 * arguments: TTSDataset.__init__ (roar/collections/tts/data/dataset.py:71-365) and
 * FilterbankFeatures.__init__ (roar/collections/asr/parts/preprocessing/features.py:196-345). */
typedef struct roar_sup_config {
  int32_t struct_size;      /* = sizeof(roar_sup_config); ABI check */
  int32_t sample_rate;
  int32_t n_fft;            /* power of two, 256..4096 */
  int32_t win_length;       /* <= n_fft; window is zero-padded centred like torch.stft */
  int32_t hop_length;
  int32_t window;           /* enum roar_sup_window, symmetric (periodic=False) */
  int32_t n_mels;
  int32_t mel_norm;         /* 1 = slaney area norm, 0 = none */
  double  fmin;             /* mel lowfreq */
  double  fmax;             /* mel highfreq; <= 0 means sample_rate/2 */
  double  spec_floor;       /* added under the sqrt: 1e-9 (TTSDataset.get_spec, dataset.py:529);
                               0 or 1e-5 (FilterbankFeatures guard, features.py:408-410) */
  double  mag_power;        /* 1.0 magnitude (TTSDataset); 2.0 power (FilterbankFeatures default) */
  int32_t log_mode;         /* enum roar_sup_log_mode */
  int32_t exact_pad;        /* FilterbankFeatures exact_pad: center=False + reflect pad (n_fft-hop)/2 */
  double  log_guard;        /* clamp floor or additive guard */
  int32_t has_preemph;      /* FilterbankFeatures preemph is not None */
  int32_t normalize;        /* enum roar_sup_normalize (roar_fbank_forward only) */
  double  preemph;
  double  pad_value;        /* roar_fbank_forward: fill beyond seq_len */
  int32_t pad_to;           /* roar_fbank_forward: T padded to a multiple (0 = off) */
  int32_t energy_mode;      /* 0: L2 norm of the linear spectrum (TTSDataset, dataset.py:751-753);
                               1: L2 norm of the output features over the mel axis (EnergyFeaturizer,
                               tts/parts/preprocessing/features.py:283-302) */
  /* pYIN (librosa.pyin keyword arguments as called at dataset.py:696-703) */
  double  pitch_fmin;
  double  pitch_fmax;
  int32_t pyin_frame_length;  /* reference passes win_length; 0 = mel-only handle (no pYIN tables, e.g. FilterbankFeatures) */
  int32_t pyin_win_length;    /* 0 => frame_length/2 */
  int32_t pyin_hop_length;    /* 0 => frame_length/4 (the reference does not pass hop_length) */
  int32_t n_thresholds;       /* 100 */
  double  beta_a, beta_b;     /* (2, 18) */
  double  boltzmann_parameter;/* 2 */
  double  resolution;         /* 0.1 */
  double  max_transition_rate;/* 35.92 */
  double  switch_prob;        /* 0.01 */
  double  no_trough_prob;     /* 0.01 */
} roar_sup_config;

typedef struct roar_sup_handle roar_sup_handle;

/* kernel ids for the optional per-kernel timing below */
enum roar_sup_kernel { ROAR_K_TILE_OFFSETS = 0, ROAR_K_STFT_MEL = 1, ROAR_K_PYIN_CMND = 2, ROAR_K_PYIN_PROBS = 3,
                       ROAR_K_LEN_SORT = 4, ROAR_K_VITERBI = 5, ROAR_K_BACKTRACK = 6, ROAR_K_PRIOR = 7,
                       ROAR_K_STATS = 8, ROAR_K_FBANK_NORM = 9, ROAR_K_PYIN_ENERGY = 10, ROAR_K_PCM16 = 11, ROAR_K_COUNT = 12 };  /* roar_sup_trim is not timed */

/* Fill cfg with the reference's extraction defaults
 * (scripts/dataset_processing/tts/rasa/ds_conf/ds_for_fastpitch_align.yaml:12-27). */
void roar_sup_config_default(roar_sup_config* cfg);

int  roar_sup_abi_version(void);
const char* roar_sup_last_error(void);

/* Build tables (window, twiddles, sparse mel rows, beta/boltzmann tables, banded log-transition
 * rows) on the host in float64 and upload them to `device`. */
int  roar_sup_create(const roar_sup_config* cfg, int device, roar_sup_handle** out);
void roar_sup_destroy(roar_sup_handle* h);

/* Diagnostics (bench.py's roofline): when on, every kernel launch is bracketed by CUDA events on the
 * launching stream; _read synchronises them and returns accumulated milliseconds and launch counts
 * per enum roar_sup_kernel (arrays of ROAR_K_COUNT).  Not thread-safe; leave off in production. */
int  roar_sup_set_profiling(roar_sup_handle* h, int on);
int  roar_sup_profile_read(roar_sup_handle* h, double* ms_out, int64_t* count_out, int reset);

/* Diagnostic builds only (-DROAR_VIT_STATS): per-category step counts of the Viterbi kernel since the last
 * reset, out[n] uint64 ([0] sparse steps, [1] dense steps, [2] unvoiced band scans, [3] voiced band scans,
 * [4] unvoiced list walks, [5] voiced list walks, [6] candidate evaluations, [7] utterance-steps in uniform mode).
 * The regular build returns zeros. */
int  roar_sup_debug_counters(uint64_t* out, int32_t n, int reset);

/* Frame counts. STFT: 1 + L/hop (center) -- TTSDataset.get_spec, dataset.py:524-530;
 * pYIN: 1 + L/pyin_hop -- librosa.pyin center=True. */
int64_t roar_sup_num_frames(const roar_sup_handle* h, int64_t n_samples);
int64_t roar_sup_pyin_num_frames(const roar_sup_handle* h, int64_t n_samples);
/* pYIN geometry: [0]=min_period [1]=max_period [2]=n_pitch_bins [3]=transition_width
 * [4]=hop [5]=win [6]=max_candidates_per_frame [7]=unique transition rows */
int  roar_sup_pyin_geometry(const roar_sup_handle* h, int32_t out8[8]);

/* Host-side table access (no device work): used by `.filter_banks`
 * (features.py:380-382) and by the CPU tests that pin the tables to the oracle. */
int  roar_sup_host_mel_filterbank(const roar_sup_config* cfg, float* out /* [n_mels, n_fft/2+1] */);
int  roar_sup_host_window(const roar_sup_config* cfg, float* out /* [n_fft] */);
/* dense [2*npb, 2*npb] float64 log-transition matrix rebuilt from the banded device tables */
int  roar_sup_host_pyin_log_transition(const roar_sup_config* cfg, double* out, int64_t n);
int  roar_sup_host_pyin_beta_probs(const roar_sup_config* cfg, double* out /* [n_thresholds] */);

/* Bytes of device workspace that serve every entry point taking a workspace (the maximum over their
 * layouts): roar_sup_pyin for total_pyin_frames frames, roar_sup_logmel_energy, and
 * roar_fbank_forward / roar_fbank_backward for a batch of n_utts rows. */
size_t roar_sup_workspace_bytes(const roar_sup_handle* h, int32_t n_utts, int64_t total_samples,
                                int64_t total_pyin_frames);
/* Bytes roar_fbank_forward / roar_fbank_backward need for a batch of B rows. */
size_t roar_fbank_workspace_bytes(const roar_sup_handle* h, int32_t B);

/* Sample-rate conversion of a packed batch on the device: replaces the resampling step of AudioSegment.__init__
 * (asr/parts/preprocessing/segment.py:68-75; the reference calls librosa.core.resample, res_type soxr_hq).
 * Polyphase FIR with scipy.signal.resample_poly's arithmetic: out[n] = sum_k taps[(n + n_pre_remove) * down -
 * n_pre_pad - k * up] * in[k]; the host designs `taps` (low-pass, multiplied by `up`) and the output lengths
 * (roar_b200/resample.py).  A different filter than soxr's: parity unpinned (oracle/resample.py). */
int roar_sup_resample(roar_sup_handle* h, const float* d_in, const int64_t* d_in_off, const int32_t* d_in_len,
                      int32_t n_utts, int32_t max_out_len, int32_t up, int32_t down, const float* d_taps,
                      int32_t n_taps, int32_t n_pre_pad, int32_t n_pre_remove, float* d_out,
                      const int64_t* d_out_off, const int32_t* d_out_len, void* stream);

/* Small metadata (offsets, lengths, prefix sums) host -> device by a kernel that reads the page-locked host block
 * directly, so it never queues behind large copies on the copy engine.  16-byte aligned pointers and size;
 * falls back to cudaMemcpyAsync when the host block is not page-locked. */
int roar_sup_upload(roar_sup_handle* h, const void* host_pinned, void* d_dst, size_t bytes, void* stream);

/* 16-bit PCM ingest: replaces the integer -> float32 step of AudioSegment._convert_samples_to_float32
 * (asr/parts/preprocessing/segment.py:140-153): d_audio[i] = d_pcm[i] / 2^15, exact.  The packed-batch
 * layout (sample_off / sample_len) applies to d_audio unchanged, so a wav corpus travels host -> device
 * at 2 bytes per sample. */
int roar_sup_pcm16_to_f32(roar_sup_handle* h, const int16_t* d_pcm, int64_t n_samples, float* d_audio,
                          void* stream);

/* log-mel + energy: replaces TTSDataset.get_spec / get_log_mel / the energy line
 * (dataset.py:524-537, 751-753; torch.stft at :324-333).
 *   d_frame_off[n_utts+1]: prefix sum of T_i = roar_sup_num_frames(L_i)
 *   d_logmel: utterance i is a contiguous [n_mels, T_i] row-major block at n_mels*frame_off[i]
 *   d_energy: [sum T_i]
 * Either output may be NULL. */
int roar_sup_logmel_energy(roar_sup_handle* h, const float* d_audio, const int64_t* d_sample_off,
                           const int32_t* d_sample_len, int32_t n_utts, const int64_t* d_frame_off,
                           int64_t total_frames, float* d_logmel, float* d_energy,
                           void* d_workspace, size_t workspace_bytes, void* stream);

/* pYIN: replaces librosa.pyin(audio, fmin, fmax, frame_length=win_length, sr, fill_na=0.0)
 * (dataset.py:696-703).  Outputs are float32 like the tensors the reference saves
 * (dataset.py:704-708): f0 (0 where unvoiced), voiced flag (0/1), voiced probability.
 *   d_frame_off[n_utts+1]: prefix sum of roar_sup_pyin_num_frames(L_i) */
int roar_sup_pyin(roar_sup_handle* h, const float* d_audio, const int64_t* d_sample_off,
                  const int32_t* d_sample_len, int32_t n_utts, const int64_t* d_frame_off,
                  int64_t total_frames, int32_t max_frames_per_utt, float* d_f0, float* d_voiced_flag,
                  float* d_voiced_prob, void* d_workspace, size_t workspace_bytes, void* stream);

/* Alignment prior: replaces beta_binomial_prior_distribution(phoneme_count, mel_count, scaling)
 * (roar/collections/tts/parts/utils/tts_dataset_utils.py:128-149).
 *   utterance i: [mel_len[i], text_len[i]] row-major float32 block at d_out_off[i] */
int roar_sup_align_prior(roar_sup_handle* h, const int32_t* d_text_len, const int32_t* d_mel_len,
                         int32_t n_utts, const int64_t* d_out_off, int32_t max_mel_len,
                         double scaling_factor, float* d_prior, void* stream);

/* Interpolated prior: replaces BetaBinomialInterpolator.__call__(mel_len, text_len)
 * (tts_dataset_utils.py:69-92; used when `use_beta_binomial_interpolator: true`): exact prior at sizes
 * rounded to multiples of (round_mel_len_to, round_text_len_to) = (50, 10) by default, resampled to
 * [mel_len, text_len] with scipy.ndimage.zoom(order=1) semantics.  Same output layout as above. */
int roar_sup_align_prior_interp(roar_sup_handle* h, const int32_t* d_text_len, const int32_t* d_mel_len,
                                int32_t n_utts, const int64_t* d_out_off, int32_t max_mel_len,
                                int32_t round_mel_len_to, int32_t round_text_len_to, float* d_prior,
                                void* stream);

/* Silence trimming: replaces librosa.effects.trim(samples, top_db, ref, frame_length, hop_length) as called by
 * AudioSegment (asr/parts/preprocessing/segment.py:76-88; TTSDataset forwards trim_* at dataset.py:613-617).
 * Writes [start, end) sample indices per utterance (relative to the utterance start); the caller narrows
 * sample_off / sample_len, no audio is copied.  ref_value <= 0 selects ref = np.max (the reference default).
 * Utterances of more than 24 576 trim frames are rejected. */
int roar_sup_trim(roar_sup_handle* h, const float* d_audio, const int64_t* d_sample_off,
                  const int32_t* d_sample_len, int32_t n_utts, int32_t max_samples_per_utt, double top_db,
                  double ref_value, int32_t frame_length, int32_t hop_length, int64_t* d_start, int64_t* d_end,
                  void* stream);

/* Partial pitch statistics over f0 != 0: replaces get_pitch_stats
 * (scripts/dataset_processing/tts/extract_sup_data.py:8-13, 29-30).
 *   d_out (float64) [n_groups, 5]: sum, sum of squares, count, min, max.  `_init` sets
 *   (0, 0, 0, +inf, 0); the partial calls ACCUMULATE, so several chunks / calls add up; ranks then
 *   all-reduce the five numbers (SUM, SUM, SUM, MIN, MAX). */
int roar_sup_pitch_partials_init(roar_sup_handle* h, double* d_out, int32_t n_groups, void* stream);
int roar_sup_pitch_partials(roar_sup_handle* h, const float* d_f0, int64_t n, double* d_out5,
                            void* stream);
/* Same per group: d_group[n_utts] int32 in [0, n_groups) (speaker ids,
 * compute_speaker_stats.py:105-132); utterance i owns f0[frame_off[i] .. frame_off[i+1]). */
int roar_sup_pitch_partials_grouped(roar_sup_handle* h, const float* d_f0, const int64_t* d_frame_off,
                                    const int32_t* d_group, int32_t n_utts, int32_t n_groups,
                                    double* d_out, void* stream);

/* FilterbankFeatures.forward (eval, no grad): replaces features.py:384-461 incl. normalize_batch
 * (:25-61).  d_x [B, Lmax] row-major, d_len[B] int64 samples ->
 *   d_out [B, n_mels, Tpad] with Tpad = roar_fbank_out_frames(h, Lmax); d_out_len[B] int64. */
int64_t roar_fbank_out_frames(const roar_sup_handle* h, int64_t Lmax);
int roar_fbank_forward(roar_sup_handle* h, const float* d_x, const int64_t* d_len, int32_t B,
                       int64_t Lmax, float* d_out, int64_t* d_out_len, void* d_workspace,
                       size_t workspace_bytes, void* stream);

/* Backward of roar_fbank_forward for FilterbankFeatures(use_grads=True) (features.py:329-331,408-410; used by
 * the mel losses of tts/models/jets.py:175-177, hifigan.py:56-58, bigvgan.py:57-59, roar_tts.py:174-176):
 * d_grad_out [B, n_mels, Tpad] -> d_grad_x [B, Lmax] (overwritten).  The forward spectrum is recomputed.
 * Requires normalize = NONE and no pre-emphasis (the reference's grad configurations). */
int roar_fbank_backward(roar_sup_handle* h, const float* d_x, const int64_t* d_len, int32_t B, int64_t Lmax,
                        const float* d_grad_out, float* d_grad_x, void* d_workspace, size_t workspace_bytes,
                        void* stream);

/* ---- host-side I/O of the extraction run (native threads, no device work) -------------------------------
 * wav decoding: replaces AudioSegment.from_file for RIFF/WAVE input (asr/parts/preprocessing/segment.py:156-278,
 * reached through WaveformFeaturizer.process, asr/parts/preprocessing/features.py:137-170): PCM 8/16/24/32 bit and
 * IEEE float 32/64, any channel count, `offset` / `duration` as frame ranges. */
typedef struct roar_wav_info {
  int32_t sample_rate;   /* -1: the file could not be probed */
  int32_t channels;
  int32_t bits;
  int32_t format;        /* 1 = integer PCM, 3 = IEEE float */
  int64_t n_frames;
  int64_t data_offset;   /* byte offset of the first frame */
} roar_wav_info;
/* Fills out[n]; returns the number of files that failed (their sample_rate is -1; roar_sup_last_error names one). */
int roar_sup_wav_probe_batch(const char* const* paths, int32_t n, roar_wav_info* out, int32_t n_threads);
/* Reads frames [first_frame[i], first_frame[i] + n_frames[i]) of file i to dst + dst_off[i] (elements):
 *   as_pcm16 = 1: raw int16 (16-bit mono PCM files only) -- pair with roar_sup_pcm16_to_f32;
 *   as_pcm16 = 0: float32 mono the way soundfile's float read scales (x / 2^(bits-1)); `channel` = -1 averages
 *                 the channels (channel_selector="average"), otherwise selects one.
 * dst is typically pinned host memory.  Returns the number of failed files. */
int roar_sup_wav_read_batch(const char* const* paths, const roar_wav_info* info, int32_t n,
                            const int64_t* first_frame, const int64_t* n_frames, int32_t channel,
                            int32_t as_pcm16, void* dst, const int64_t* dst_off, int32_t n_threads);
/* Cache writer: replaces torch.save(tensor, path) of the cached sup data (tts/data/dataset.py:656-657, 704-708,
 * 752-753).  File i holds the contiguous float32 CPU tensor base[elem_off[i] ...] of shape shape3[3*i .. 3*i+rank[i])
 * in torch's zip serialization (loadable by torch.load, weights_only or not); written to a temporary name and
 * renamed into place, so an interrupted run never leaves a truncated cache file.  Returns the number of failures. */
int roar_sup_pt_write_batch(const float* base, int32_t n_files, const int64_t* elem_off, const int32_t* rank,
                            const int64_t* shape3, const char* const* paths, int32_t n_threads);

#ifdef __cplusplus
}
#endif
#endif /* ROAR_SUP_H */
