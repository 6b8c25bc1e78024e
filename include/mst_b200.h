/*
 * mst_b200.h -- C ABI of the B200-native preprocessing / inversion hot path.
 *
 * Drop-in boundary for silburt/ML_Music_Style_Transfer.  The reference is pure Python; the
 * entry points below are what a ctypes / cffi stub (or the torch custom ops shipped in
 * ml_music_style_transfer_b200/csrc/torch_ops.cpp) binds in place of the third-party calls the
 * reference makes on its hot path.  Every function cites the reference call it replaces
 * (paths relative to the reference repository root).
 *
 * Conventions
 *   - plain C types only; no torch / C++ types cross this boundary;
 *   - `const T* d_*` arguments are CALLER-OWNED DEVICE pointers on the current CUDA device,
 *     `h_*` arguments are HOST pointers;
 *   - every launch is asynchronous on `stream` (a cudaStream_t passed as void*; NULL = legacy
 *     default stream); nothing here synchronises the device except *_create / *_destroy;
 *   - return value: 0 = MST_OK, negative = error, text via mst_last_error() (thread-local);
 *   - there is no CPU fallback: without a CUDA device every compute entry returns MST_ERR_CUDA.
 *   - n_fft = 2048 (the only value the reference uses: preprocess.py:25, inference.py:105) runs the tuned
 *     warp-per-frame kernels; every other power of two in [64, 16384] runs the general path (one CTA per
 *     frame, shared-memory FFT) behind the same entry points, every spectrum then has n_fft/2 + 1 bins;
 *     anything else returns MST_ERR_UNSUPPORTED.  Mel inversion is n_fft = 2048 only.
 */
#ifndef MST_B200_H_
#define MST_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MST_OK 0
#define MST_ERR_INVALID (-1)
#define MST_ERR_UNSUPPORTED (-2)
#define MST_ERR_CUDA (-3)
#define MST_ERR_WORKSPACE (-4)

/* STFT epilogues (fused into the FFT kernel). */
#define MST_OUT_COMPLEX 0     /* float2 per bin: librosa.stft itself                          */
#define MST_OUT_MAGNITUDE 1   /* |S|          (tests/plot_spec.py:17, Griffin-Lim input)       */
#define MST_OUT_POWER 2       /* |S|^2        (librosa melspectrogram's S, plot_spec.py:20)    */
#define MST_OUT_LOG1P_POWER 3 /* log1p(|S|^2) (preprocess.py:49)                               */

/* Memory layout of per-clip (bins x frames) matrices. */
#define MST_LAYOUT_FRAME_MAJOR 0 /* [T][K]: frame contiguous == librosa's Fortran-ordered (K,T) */
#define MST_LAYOUT_BIN_MAJOR 1   /* [K][T]: C-ordered (K,T), what np.array(spec_list) yields    */

#define MST_PAD_REFLECT 0  /* librosa <= 0.9 default, np.pad(mode='reflect')  */
#define MST_PAD_CONSTANT 1 /* librosa >= 0.10 default, zero padding           */

#define MST_DTYPE_I8 0
#define MST_DTYPE_F32 1
#define MST_DTYPE_F64 2

typedef void* mst_stream_t;

const char* mst_last_error(void);
int mst_version(void);
/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
int64_t mst_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * Clip batch descriptor: ragged set of clips inside one device audio buffer.
 * Replaces the Python loop + slicing of preprocess.py:63-75 (`audio[step*ws*stride : ... +
 * n_samples_per_chunk]`): chunk c is (offset = c*131072, length = 219904) and chunks may overlap.
 * Frame count per clip follows librosa.stft(center=True): T = 1 + length / hop.
 * Host arrays are copied; the descriptor owns small device tables.
 * ------------------------------------------------------------------------------------------- */
typedef struct mst_batch mst_batch_t;
int mst_batch_create(int n_clips, const int64_t* h_clip_offsets, const int64_t* h_clip_lengths,
                     int n_fft, int hop, int pad_mode, mst_batch_t** out);
/* Batch described by frames per clip (Griffin-Lim side): length_c = hop * (T_c - 1), clips packed
 * back to back in the output waveform buffer. */
int mst_batch_create_from_frames(int n_clips, const int64_t* h_frames_per_clip, int n_fft, int hop,
                                 int pad_mode, mst_batch_t** out);
/* The same with librosa's win_length (<= n_fft): the periodic Hann window of that length, centre-padded to n_fft, is the
 * analysis window of librosa.stft(..., win_length=) and the analysis + synthesis window (and window-sum-square envelope)
 * of librosa.griffinlim(..., win_length=) [model/inference.py:110 passes win_length=n_fft]. */
int mst_batch_create_ex(int n_clips, const int64_t* h_clip_offsets, const int64_t* h_clip_lengths, int n_fft, int hop,
                        int win_length, int pad_mode, mst_batch_t** out);
int mst_batch_create_from_frames_ex(int n_clips, const int64_t* h_frames_per_clip, int n_fft, int hop, int win_length,
                                    int pad_mode, mst_batch_t** out);
void mst_batch_destroy(mst_batch_t* b);
int mst_batch_n_clips(const mst_batch_t* b);
int64_t mst_batch_total_frames(const mst_batch_t* b);
int64_t mst_batch_total_samples(const mst_batch_t* b); /* sum of clip lengths */
int64_t mst_batch_audio_extent(const mst_batch_t* b);  /* max(clip offset + length): floats the audio buffer must hold */
int mst_batch_device(const mst_batch_t* b);            /* CUDA device the descriptor tables live on */
int mst_batch_n_fft(const mst_batch_t* b);             /* FFT size: every spectrum of this batch has n_fft/2 + 1 bins */
int64_t mst_batch_clip_frames(const mst_batch_t* b, int clip);
int64_t mst_batch_frame_offset(const mst_batch_t* b, int clip); /* prefix sum of frames */

/* ---------------------------------------------------------------------------------------------
 * P1: framing + Hann window + rFFT(2048) + fused epilogue.
 * Replaces librosa.stft(y, n_fft=2048, hop_length=hop) [preprocess.py:48, plot_spec.py:14] and the
 * NumPy epilogue np.log1p(np.abs(spec)**2) [preprocess.py:49].
 * d_out: FRAME_MAJOR -> rows are global frame ids (clip c starts at row frame_offset(c)), 1025
 *        elements per row; BIN_MAJOR -> clip c occupies [1025][T_c] starting at element
 *        frame_offset(c)*1025.  Element = float2 for MST_OUT_COMPLEX, float otherwise.
 * ------------------------------------------------------------------------------------------- */
int mst_stft_f32(const float* d_audio, const mst_batch_t* batch, int out_mode, int layout, void* d_out,
                 mst_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * P2: mel filterbank + projection.
 * mst_mel_filterbank_f32 replaces librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax, htk=False,
 * norm='slaney') (float64 maths, float32 result, [n_mels][1+n_fft/2] row-major, host memory).
 * A mel plan owns the device form of a filterbank: its non-zero band per 64-bin slice, split into
 * bf16 (hi, lo) and stored as the shared-memory image the tcgen05 projection kernel consumes.  mst_stft_mel_f32 replaces librosa.feature.melspectrogram(y=, sr=, n_fft=, hop_length=)
 * [plot_spec.py:20; preprocess.py:55] = mel_basis @ |stft|^2, with optional log1p (this build's
 * "log-mel", following the log1p convention of preprocess.py:49).
 * d_out: FRAME_MAJOR [total_frames][n_mels] or BIN_MAJOR per clip [n_mels][T_c].
 * ------------------------------------------------------------------------------------------- */
int mst_mel_filterbank_f32(int sr, int n_fft, int n_mels, double fmin, double fmax, float* h_weights);
typedef struct mst_mel_plan mst_mel_plan_t;
int mst_mel_plan_create(const float* h_weights, int n_mels, int n_bins, mst_mel_plan_t** out);
void mst_mel_plan_destroy(mst_mel_plan_t* p);
/* Workspace for the ring of split-precision power-spectrum rows between the STFT and the projection kernel: sized for
 * THIS batch (at most 8 waves of 128-frame projection tiles, never more rows than the batch has frames); query it with the
 * batch you are going to pass to mst_stft_mel_f32. */
size_t mst_stft_mel_workspace_bytes(const mst_batch_t* batch, const mst_mel_plan_t* plan);
int mst_stft_mel_f32(const float* d_audio, const mst_batch_t* batch, const mst_mel_plan_t* plan,
                     int apply_log1p, int layout, float* d_out, void* d_workspace, size_t workspace_bytes,
                     mst_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * P3: MIDI notes -> 128-pitch piano roll -> binarise -> on/off map -> chunks / audio-rate planes.
 * Replaces pretty_midi.PrettyMIDI(...).get_piano_roll(fs=hp.wps).T + the NumPy lines of
 * preprocess.py:146-155 (duplicated at model/inference.py:40-49).  Notes are SoA device arrays
 * for a batch of pieces; piece p owns notes [h_note_offsets[p], h_note_offsets[p+1]).
 * Column indices are int(start*fs), int(end*fs) on IEEE double products (truncate toward zero).
 * ------------------------------------------------------------------------------------------- */
/* T_p = int(fs * max_end_p) for every piece (pretty_midi: roll width); d_rows_out is int64[n_pieces]. */
int mst_pianoroll_count_rows(const double* d_end, const int64_t* d_note_offsets, int n_pieces, int fs,
                             int64_t* d_rows_out, mst_stream_t stream);
/* d_row_offsets: int64[n_pieces+1] prefix sum of T_p.  Outputs are [sum T][128], time-major as the
 * reference's `.T`: d_roll (uint8 in {0,1}), d_onoff (int8 in {-1,0,1}), optional d_velsum (int32,
 * the un-binarised get_piano_roll value; may be NULL). */
int mst_pianoroll_rasterize(const int32_t* d_pitch, const int32_t* d_velocity, const double* d_start,
                            const double* d_end, const int64_t* d_note_offsets, int n_pieces,
                            const int64_t* d_row_offsets, int64_t total_rows, int64_t total_notes, int fs,
                            uint8_t* d_roll, int8_t* d_onoff, int32_t* d_velsum, mst_stream_t stream);
/* Same, followed by pretty_midi >= 0.2.9's CC64 sustain rule (get_piano_roll(pedal_threshold=64)): inside each
 * pedal-down span [d_span_start, d_span_end) (columns of piece d_span_piece) every pitch keeps the running maximum of
 * its velocity sum, then roll / onoff are derived from the sustained sums.  Spans are computed by the host from the
 * CC64 events (int(cc.time * fs) on the host is already exact).  d_velsum is required when n_spans > 0. */
int mst_pianoroll_rasterize_sustain(const int32_t* d_pitch, const int32_t* d_velocity, const double* d_start,
                                    const double* d_end, const int64_t* d_note_offsets, int n_pieces,
                                    const int64_t* d_row_offsets, int64_t total_rows, int64_t total_notes, int fs,
                                    const int32_t* d_span_piece, const int64_t* d_span_start, const int64_t* d_span_end,
                                    int n_spans, uint8_t* d_roll, int8_t* d_onoff, int32_t* d_velsum,
                                    mst_stream_t stream);
/* PrettyMIDI.get_piano_roll over a whole file (preprocess.py:146-147, inference.py:40-41 call it on files that may
 * hold several instruments): the per-instrument velocity sums (d_velsum, already through the CC64 rule, one "piece" per
 * instrument, rows d_inst_row_offsets) get their pitch bends applied and are summed, in instrument order, into the
 * widest roll of their file.  Bend segment = one (start_bend, end_bend) pair of pretty_midi's loop with |pitch| >= 1:
 * columns [c0, c1), bend_int, bend_decimal d and m1 = 1 - d (float64, evaluated by the host exactly as Python does),
 * positive = (pitch >= 0).  Segments of an instrument are sorted and disjoint (d_seg_offsets: int32[n_instruments+1]).
 * Outputs: d_out_f64 (optional, [total_rows][128] float64 = get_piano_roll(fs).T), d_roll = (value != 0), d_onoff. */
typedef struct mst_bend_segment {
  int64_t c0, c1;
  double d, m1;
  int32_t bend_int;
  int32_t positive;
} mst_bend_segment_t;
int mst_pianoroll_merge_instruments(const int32_t* d_velsum, const int64_t* d_inst_row_offsets, const int32_t* d_inst_is_drum,
                                    int n_instruments, const int32_t* d_file_inst_offsets, const int64_t* d_file_row_offsets,
                                    int n_files, int64_t total_rows, const int32_t* d_seg_offsets, const void* d_segments,
                                    double* d_out_f64, uint8_t* d_roll, int8_t* d_onoff, mst_stream_t stream);
/* preprocess.py:80-96: out[c][j][p] = plane[c*stride_rows + j][p], j < chunk_rows; rows beyond
 * n_rows read as 0.  in: int8/uint8 plane; out_dtype MST_DTYPE_{I8,F32,F64} (reference: float64). */
int mst_pianoroll_chunks(const void* d_plane, int64_t n_rows, int num_chunks, int chunk_rows, int stride_rows,
                         int out_dtype, void* d_out, mst_stream_t stream);
/* README.md:19-20 (design note only in the reference; SURVEY section 8a P3d defines it):
 * out[p][k][n] = plane[row_offsets[p] + (n*fs)/sr][pitch_lo + k] for n < N_p, 0 where the column
 * is >= T_p.  Piece p's block starts at element n_keys * d_sample_offsets[p]. */
int mst_pianoroll_upsample(const void* d_plane, const int64_t* d_row_offsets, const int64_t* d_sample_offsets,
                           int n_pieces, int64_t total_samples, int fs, int sr, int pitch_lo, int n_keys,
                           int out_dtype, void* d_out, mst_stream_t stream);
/* The same for the roll and its on/off map in ONE launch (the two planes that condition the model, README.md:19-20,27):
 * the index arithmetic is shared.  d_out_roll / d_out_onoff must agree in alignment modulo 16 bytes. */
int mst_pianoroll_upsample_pair(const void* d_roll, const void* d_onoff, const int64_t* d_row_offsets,
                                const int64_t* d_sample_offsets, int n_pieces, int64_t total_samples, int fs, int sr,
                                int pitch_lo, int n_keys, int out_dtype, void* d_out_roll, void* d_out_onoff,
                                mst_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * Mel inversion: librosa.feature.inverse.mel_to_stft, the first half of the
 * librosa.feature.inverse.mel_to_audio(M, n_iter=300, sr=hp.sr, n_fft=hp.n_fft, hop_length=hp.ws) call kept as a
 * comment at tests/test_griffinlim.py:24:  S = nnls(mel_basis, M) ** (1 / power), then Griffin-Lim (P4) on S.
 * Every frame is an independent non-negative least-squares problem min ||A x - m||, x >= 0 (A = the filterbank): start
 * point max(pinv(A) m, 0) as in librosa.util.nnls, then accelerated projected gradient to `tol` (relative residual) or
 * `max_iter`.  The plan factorises A A^T once on the host (double precision).
 * d_mel: FRAME_MAJOR [total_frames][n_mels] or BIN_MAJOR per clip [n_mels][T_c]; `batch` from
 * mst_batch_create_from_frames; d_S_out: [total_frames][1025] frame-major magnitudes (what mst_griffinlim_f32 consumes in
 * place). */
typedef struct mst_mel_inverse_plan mst_mel_inverse_plan_t;
int mst_mel_inverse_plan_create(const float* h_weights, int n_mels, int n_bins, mst_mel_inverse_plan_t** out);
void mst_mel_inverse_plan_destroy(mst_mel_inverse_plan_t* p);
int mst_mel_to_stft_f32(const float* d_mel, int mel_layout, const mst_batch_t* batch, const mst_mel_inverse_plan_t* plan,
                        float power, int max_iter, float tol, float* d_S_out, mst_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * P4: Griffin-Lim phase reconstruction (fast Griffin-Lim, momentum; momentum=0 is the classic
 * loop kept as a comment at model/inference.py:131-154).
 * Replaces AudioSynthesizer.griffinlim [model/inference.py:105-110]:
 *   magnitude = sqrt(expm1(clip(spec, 0, 20)))            (s_is_log1p_power != 0)
 *   librosa.griffinlim(magnitude, n_iter, window='hann', win_length=2048, hop_length=hop)
 * and the call at tests/test_griffinlim.py:23.
 * batch: created with mst_batch_create_from_frames.  d_S: magnitudes (or log1p-power) per clip in
 * `s_layout`.  d_init_phase: uniform [0,1) field, same layout as S (angles = exp(2*pi*i*u), what
 * librosa draws from RandomState.rand); NULL -> counter-based device RNG seeded with `seed`
 * (init_mode 0) or angles = 1 (init_mode 1, librosa init=None).
 * d_y_out: packed waveforms, clip c at sample offset sum_{j<c} hop*(T_j-1).
 * State (previous iterate, overlap-add accumulators) lives in the caller's workspace in HBM.
 * ------------------------------------------------------------------------------------------- */
size_t mst_griffinlim_workspace_bytes(const mst_batch_t* batch); /* worst case over layouts */
/* Exact size for a given input form: frame-major magnitudes are used in place (no 4 B/bin copy: 16 B/bin of state). */
size_t mst_griffinlim_workspace_bytes_ex(const mst_batch_t* batch, int s_layout, int s_is_log1p_power);
int mst_griffinlim_f32(const float* d_S, int s_layout, int s_is_log1p_power, const mst_batch_t* batch,
                       int n_iter, float momentum, const float* d_init_phase, int init_mode, uint64_t seed,
                       float* d_y_out, void* d_workspace, size_t workspace_bytes, mst_stream_t stream);

/* Spectral convergence of waveforms against target magnitudes, the normalised form of the Frobenius loss the
 * reference prints at model/inference.py:149-150: per clip c, d_num[c] = sum (|STFT(y_c)| - S_c)^2 and
 * d_den[c] = sum S_c^2 (double; SC_c = sqrt(num/den)).  `batch` describes the clips of d_y (mst_batch_create);
 * d_S holds one (1025 x T_c) block per clip in `s_layout`.  Fused into the STFT kernel: no spectrogram is written. */
int mst_spectral_convergence_f32(const float* d_y, const mst_batch_t* batch, const float* d_S, int s_layout,
                                 double* d_num, double* d_den, mst_stream_t stream);

/* ---------------------------------------------------------------------------------------------
 * "Next" row: the resampling half of librosa.load(path, sr=hp.sr)
 * [preprocess.py:106, model/inference.py:54, tests/test_griffinlim.py:16] = resampy 'kaiser_best'
 * (64 zero crossings, Kaiser beta 14.7697, roll-off 0.9476, 512-per-crossing table, linear interpolation).
 * mst_resample_length = ceil(n_in * sr_out / sr_in) (librosa pads resampy's floor() length with zeros).
 * d_out must hold mst_resample_length(...) floats.
 * ------------------------------------------------------------------------------------------- */
int64_t mst_resample_length(int64_t n_in, int sr_in, int sr_out);
int mst_resample_f32(const float* d_in, int64_t n_in, int sr_in, int sr_out, float* d_out, mst_stream_t stream);
/* librosa.to_mono (np.mean over channels, float32) of the interleaved frames a WAV file stores: d_out[t] =
 * (sum_c d_interleaved[t * channels + c]) / channels. */
int mst_mono_mix_f32(const float* d_interleaved, int64_t n_frames, int channels, float* d_out, mst_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MST_B200_H_ */
