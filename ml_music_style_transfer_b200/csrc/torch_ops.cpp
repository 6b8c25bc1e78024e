// PyTorch custom ops (namespace mst_b200) over the C ABI in include/mst_b200.h.
// PyTorch is plumbing here: it owns device memory and the current stream; every op forwards raw
// pointers to libmst_b200.so.  Compute ops are registered for the CUDA dispatch key only -- there is
// no CPU implementation to fall back to.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/library.h>
#include <torch/types.h>
#include "../../include/mst_b200.h"

namespace {

void check(int rc, const char* what) {
  TORCH_CHECK(rc == MST_OK, what, " failed (", rc, "): ", mst_last_error());
}
mst_stream_t cur_stream() { return reinterpret_cast<mst_stream_t>(at::cuda::getCurrentCUDAStream().stream()); }
mst_batch_t* as_batch(int64_t h) { TORCH_CHECK(h != 0, "null batch handle"); return reinterpret_cast<mst_batch_t*>(h); }
// A batch handle indexes the audio / waveform tensor it is used with: refuse tensors that are too small or live on
// another device instead of letting a kernel read out of bounds.
void check_batch_audio(const mst_batch_t* b, const at::Tensor& t, const char* name) {
  TORCH_CHECK(mst_batch_device(b) == (int)t.get_device(), "the batch handle was built on cuda:", mst_batch_device(b), " but ",
              name, " lives on cuda:", (int)t.get_device());
  TORCH_CHECK(t.numel() >= mst_batch_audio_extent(b), name, " has ", t.numel(), " samples but the batch's clips reach ",
              mst_batch_audio_extent(b));
}
mst_mel_plan_t* as_plan(int64_t h) { TORCH_CHECK(h != 0, "null mel plan handle"); return reinterpret_cast<mst_mel_plan_t*>(h); }
void want(const at::Tensor& t, at::ScalarType st, const char* name) {
  TORCH_CHECK(t.is_cuda(), name, " must be a CUDA tensor");
  TORCH_CHECK(t.scalar_type() == st, name, " has the wrong dtype");
  TORCH_CHECK(t.is_contiguous(), name, " must be contiguous");
}
void want_cpu_i64(const at::Tensor& t, const char* name) {
  TORCH_CHECK(t.device().is_cpu() && t.scalar_type() == at::kLong && t.is_contiguous() && t.dim() == 1, name,
              " must be a contiguous 1-D int64 CPU tensor");
}

// ---- handles (host side) -------------------------------------------------------------------
int64_t batch_create(const at::Tensor& offsets, const at::Tensor& lengths, int64_t n_fft, int64_t hop, int64_t pad_mode,
                     int64_t device, int64_t win_length) {
  want_cpu_i64(offsets, "clip_offsets");
  want_cpu_i64(lengths, "clip_lengths");
  TORCH_CHECK(offsets.numel() == lengths.numel(), "offsets / lengths size mismatch");
  c10::cuda::CUDAGuard guard((c10::DeviceIndex)device);
  mst_batch_t* b = nullptr;
  check(mst_batch_create_ex((int)offsets.numel(), offsets.data_ptr<int64_t>(), lengths.data_ptr<int64_t>(), (int)n_fft,
                            (int)hop, (int)win_length, (int)pad_mode, &b), "mst_batch_create");
  return reinterpret_cast<int64_t>(b);
}
int64_t batch_create_from_frames(const at::Tensor& frames, int64_t n_fft, int64_t hop, int64_t pad_mode, int64_t device,
                                 int64_t win_length) {
  want_cpu_i64(frames, "frames_per_clip");
  c10::cuda::CUDAGuard guard((c10::DeviceIndex)device);
  mst_batch_t* b = nullptr;
  check(mst_batch_create_from_frames_ex((int)frames.numel(), frames.data_ptr<int64_t>(), (int)n_fft, (int)hop,
                                        (int)win_length, (int)pad_mode, &b), "mst_batch_create_from_frames");
  return reinterpret_cast<int64_t>(b);
}
void batch_destroy(int64_t h) { if (h) mst_batch_destroy(reinterpret_cast<mst_batch_t*>(h)); }
int64_t batch_total_frames(int64_t h) { return mst_batch_total_frames(as_batch(h)); }
int64_t batch_total_samples(int64_t h) { return mst_batch_total_samples(as_batch(h)); }
int64_t batch_clip_frames(int64_t h, int64_t c) { return mst_batch_clip_frames(as_batch(h), (int)c); }

at::Tensor mel_filterbank(int64_t sr, int64_t n_fft, int64_t n_mels, double fmin, double fmax) {
  at::Tensor W = at::empty({n_mels, 1 + n_fft / 2}, at::TensorOptions().dtype(at::kFloat).device(at::kCPU));
  check(mst_mel_filterbank_f32((int)sr, (int)n_fft, (int)n_mels, fmin, fmax, W.data_ptr<float>()), "mst_mel_filterbank_f32");
  return W;
}
int64_t mel_plan_create(const at::Tensor& W, int64_t device) {
  TORCH_CHECK(W.device().is_cpu() && W.scalar_type() == at::kFloat && W.is_contiguous() && W.dim() == 2,
              "mel weights must be a contiguous 2-D float32 CPU tensor");
  c10::cuda::CUDAGuard guard((c10::DeviceIndex)device);
  mst_mel_plan_t* p = nullptr;
  check(mst_mel_plan_create(W.data_ptr<float>(), (int)W.size(0), (int)W.size(1), &p), "mst_mel_plan_create");
  return reinterpret_cast<int64_t>(p);
}
void mel_plan_destroy(int64_t h) { if (h) mst_mel_plan_destroy(reinterpret_cast<mst_mel_plan_t*>(h)); }
int64_t mel_inverse_plan_create(const at::Tensor& W, int64_t device) {
  TORCH_CHECK(W.device().is_cpu() && W.scalar_type() == at::kFloat && W.is_contiguous() && W.dim() == 2,
              "mel weights must be a contiguous 2-D float32 CPU tensor");
  c10::cuda::CUDAGuard guard((c10::DeviceIndex)device);
  mst_mel_inverse_plan_t* p = nullptr;
  check(mst_mel_inverse_plan_create(W.data_ptr<float>(), (int)W.size(0), (int)W.size(1), &p), "mst_mel_inverse_plan_create");
  return reinterpret_cast<int64_t>(p);
}
void mel_inverse_plan_destroy(int64_t h) { if (h) mst_mel_inverse_plan_destroy(reinterpret_cast<mst_mel_inverse_plan_t*>(h)); }
int64_t launch_count() { return mst_launch_count(); }

// ---- device ops --------------------------------------------------------------------------------
at::Tensor stft(const at::Tensor& audio, int64_t batch, int64_t out_mode, int64_t layout) {
  want(audio, at::kFloat, "audio");
  c10::cuda::CUDAGuard guard(audio.device());
  mst_batch_t* b = as_batch(batch);
  check_batch_audio(b, audio, "audio");
  const int64_t F = mst_batch_total_frames(b);
  const int64_t K = mst_batch_n_fft(b) / 2 + 1;
  at::Tensor out = out_mode == MST_OUT_COMPLEX ? at::empty({F, K}, audio.options().dtype(at::kComplexFloat))
                                               : at::empty({F * K}, audio.options());
  check(mst_stft_f32(audio.data_ptr<float>(), b, (int)out_mode, (int)layout, out.data_ptr(), cur_stream()), "mst_stft_f32");
  return out;
}

at::Tensor stft_mel(const at::Tensor& audio, int64_t batch, int64_t plan, int64_t n_mels, bool apply_log1p, int64_t layout) {
  want(audio, at::kFloat, "audio");
  c10::cuda::CUDAGuard guard(audio.device());
  mst_batch_t* b = as_batch(batch);
  mst_mel_plan_t* p = as_plan(plan);
  check_batch_audio(b, audio, "audio");
  const int64_t F = mst_batch_total_frames(b);
  at::Tensor out = at::empty({F * n_mels}, audio.options());
  const size_t ws_bytes = mst_stft_mel_workspace_bytes(b, p);
  at::Tensor ws = at::empty({(int64_t)ws_bytes + 256}, audio.options().dtype(at::kByte));
  check(mst_stft_mel_f32(audio.data_ptr<float>(), b, p, apply_log1p ? 1 : 0, (int)layout, out.data_ptr<float>(),
                         ws.data_ptr(), ws_bytes, cur_stream()), "mst_stft_mel_f32");
  return out;
}

at::Tensor pianoroll_count_rows(const at::Tensor& end, const at::Tensor& note_offsets, int64_t fs) {
  want(end, at::kDouble, "end");
  want(note_offsets, at::kLong, "note_offsets");
  c10::cuda::CUDAGuard guard(end.device());
  const int64_t n_pieces = note_offsets.numel() - 1;
  at::Tensor rows = at::empty({n_pieces}, end.options().dtype(at::kLong));
  check(mst_pianoroll_count_rows(end.data_ptr<double>(), note_offsets.data_ptr<int64_t>(), (int)n_pieces, (int)fs,
                                 rows.data_ptr<int64_t>(), cur_stream()), "mst_pianoroll_count_rows");
  return rows;
}

std::tuple<at::Tensor, at::Tensor, at::Tensor> pianoroll_rasterize(const at::Tensor& pitch, const at::Tensor& velocity,
                                                                   const at::Tensor& start, const at::Tensor& end,
                                                                   const at::Tensor& note_offsets,
                                                                   const at::Tensor& row_offsets, int64_t total_rows,
                                                                   int64_t fs, bool want_velsum,
                                                                   const c10::optional<at::Tensor>& span_piece,
                                                                   const c10::optional<at::Tensor>& span_start,
                                                                   const c10::optional<at::Tensor>& span_end) {
  want(pitch, at::kInt, "pitch"); want(velocity, at::kInt, "velocity");
  want(start, at::kDouble, "start"); want(end, at::kDouble, "end");
  want(note_offsets, at::kLong, "note_offsets"); want(row_offsets, at::kLong, "row_offsets");
  c10::cuda::CUDAGuard guard(pitch.device());
  const int64_t n_pieces = note_offsets.numel() - 1;
  TORCH_CHECK(row_offsets.numel() == n_pieces + 1, "row_offsets must have n_pieces + 1 entries");
  at::Tensor roll = at::empty({total_rows, 128}, pitch.options().dtype(at::kByte));
  at::Tensor onoff = at::empty({total_rows, 128}, pitch.options().dtype(at::kChar));
  const int64_t n_spans = span_piece.has_value() ? span_piece->numel() : 0;
  if (n_spans > 0) {
    TORCH_CHECK(span_start.has_value() && span_end.has_value(), "span_start / span_end missing");
    want(*span_piece, at::kInt, "span_piece"); want(*span_start, at::kLong, "span_start"); want(*span_end, at::kLong, "span_end");
    TORCH_CHECK(span_start->numel() == n_spans && span_end->numel() == n_spans, "span arrays differ in length");
    want_velsum = true;
  }
  at::Tensor velsum = want_velsum ? at::empty({total_rows, 128}, pitch.options().dtype(at::kInt))
                                  : at::empty({0}, pitch.options().dtype(at::kInt));
  check(mst_pianoroll_rasterize_sustain(pitch.data_ptr<int32_t>(), velocity.data_ptr<int32_t>(), start.data_ptr<double>(),
                                        end.data_ptr<double>(), note_offsets.data_ptr<int64_t>(), (int)n_pieces,
                                        row_offsets.data_ptr<int64_t>(), total_rows, pitch.numel(), (int)fs,
                                        n_spans ? span_piece->data_ptr<int32_t>() : nullptr,
                                        n_spans ? span_start->data_ptr<int64_t>() : nullptr,
                                        n_spans ? span_end->data_ptr<int64_t>() : nullptr, (int)n_spans,
                                        roll.data_ptr<uint8_t>(), onoff.data_ptr<int8_t>(),
                                        want_velsum ? velsum.data_ptr<int32_t>() : nullptr, cur_stream()),
        "mst_pianoroll_rasterize_sustain");
  return {roll, onoff, velsum};
}

std::tuple<at::Tensor, at::Tensor, at::Tensor> pianoroll_merge_instruments(
    const at::Tensor& velsum, const at::Tensor& inst_row_offsets, const at::Tensor& inst_is_drum,
    const at::Tensor& file_inst_offsets, const at::Tensor& file_row_offsets, int64_t total_rows,
    const at::Tensor& seg_offsets, const at::Tensor& segments, bool want_f64) {
  want(velsum, at::kInt, "velsum"); want(inst_row_offsets, at::kLong, "inst_row_offsets");
  want(inst_is_drum, at::kInt, "inst_is_drum"); want(file_inst_offsets, at::kInt, "file_inst_offsets");
  want(file_row_offsets, at::kLong, "file_row_offsets"); want(seg_offsets, at::kInt, "seg_offsets");
  want(segments, at::kByte, "segments");
  c10::cuda::CUDAGuard guard(velsum.device());
  const int64_t n_inst = inst_is_drum.numel(), n_files = file_row_offsets.numel() - 1;
  TORCH_CHECK(inst_row_offsets.numel() == n_inst + 1 && seg_offsets.numel() == n_inst + 1 &&
              file_inst_offsets.numel() == n_files + 1, "offset arrays do not match the instrument / file counts");
  TORCH_CHECK(segments.numel() % (int64_t)sizeof(mst_bend_segment_t) == 0, "segments must be packed mst_bend_segment_t");
  at::Tensor out = want_f64 ? at::empty({total_rows, 128}, velsum.options().dtype(at::kDouble))
                            : at::empty({0}, velsum.options().dtype(at::kDouble));
  at::Tensor roll = at::empty({total_rows, 128}, velsum.options().dtype(at::kByte));
  at::Tensor onoff = at::empty({total_rows, 128}, velsum.options().dtype(at::kChar));
  check(mst_pianoroll_merge_instruments(velsum.data_ptr<int32_t>(), inst_row_offsets.data_ptr<int64_t>(),
                                        inst_is_drum.data_ptr<int32_t>(), (int)n_inst, file_inst_offsets.data_ptr<int32_t>(),
                                        file_row_offsets.data_ptr<int64_t>(), (int)n_files, total_rows,
                                        seg_offsets.data_ptr<int32_t>(), segments.data_ptr(),
                                        want_f64 ? out.data_ptr<double>() : nullptr, roll.data_ptr<uint8_t>(),
                                        onoff.data_ptr<int8_t>(), cur_stream()), "mst_pianoroll_merge_instruments");
  return {roll, onoff, out};
}

at::ScalarType dtype_of(int64_t code) {
  switch (code) {
    case MST_DTYPE_I8: return at::kChar;
    case MST_DTYPE_F32: return at::kFloat;
    case MST_DTYPE_F64: return at::kDouble;
  }
  TORCH_CHECK(false, "bad dtype code ", code);
}

at::Tensor pianoroll_chunks(const at::Tensor& plane, int64_t num_chunks, int64_t chunk_rows, int64_t stride_rows,
                            int64_t out_dtype) {
  TORCH_CHECK(plane.is_cuda() && plane.is_contiguous() && plane.dim() == 2 && plane.size(1) == 128 &&
              (plane.scalar_type() == at::kByte || plane.scalar_type() == at::kChar), "plane must be a CUDA (T,128) int8/uint8 tensor");
  c10::cuda::CUDAGuard guard(plane.device());
  at::Tensor out = at::empty({num_chunks, chunk_rows, 128}, plane.options().dtype(dtype_of(out_dtype)));
  check(mst_pianoroll_chunks(plane.data_ptr(), plane.size(0), (int)num_chunks, (int)chunk_rows, (int)stride_rows,
                             (int)out_dtype, out.data_ptr(), cur_stream()), "mst_pianoroll_chunks");
  return out;
}

at::Tensor pianoroll_upsample(const at::Tensor& plane, const at::Tensor& row_offsets, const at::Tensor& sample_offsets,
                              int64_t total_samples, int64_t fs, int64_t sr, int64_t pitch_lo, int64_t n_keys,
                              int64_t out_dtype) {
  TORCH_CHECK(plane.is_cuda() && plane.is_contiguous() && plane.dim() == 2 && plane.size(1) == 128 &&
              (plane.scalar_type() == at::kByte || plane.scalar_type() == at::kChar), "plane must be a CUDA (T,128) int8/uint8 tensor");
  want(row_offsets, at::kLong, "row_offsets"); want(sample_offsets, at::kLong, "sample_offsets");
  c10::cuda::CUDAGuard guard(plane.device());
  at::Tensor out = at::empty({n_keys * total_samples}, plane.options().dtype(dtype_of(out_dtype)));
  check(mst_pianoroll_upsample(plane.data_ptr(), row_offsets.data_ptr<int64_t>(), sample_offsets.data_ptr<int64_t>(),
                               (int)(row_offsets.numel() - 1), total_samples, (int)fs, (int)sr, (int)pitch_lo, (int)n_keys,
                               (int)out_dtype, out.data_ptr(), cur_stream()), "mst_pianoroll_upsample");
  return out;
}

std::tuple<at::Tensor, at::Tensor> pianoroll_upsample_pair(const at::Tensor& roll, const at::Tensor& onoff,
                                                           const at::Tensor& row_offsets, const at::Tensor& sample_offsets,
                                                           int64_t total_samples, int64_t fs, int64_t sr, int64_t pitch_lo,
                                                           int64_t n_keys, int64_t out_dtype) {
  for (const at::Tensor* t : {&roll, &onoff})
    TORCH_CHECK(t->is_cuda() && t->is_contiguous() && t->dim() == 2 && t->size(1) == 128 &&
                (t->scalar_type() == at::kByte || t->scalar_type() == at::kChar), "planes must be CUDA (T,128) int8/uint8 tensors");
  TORCH_CHECK(roll.size(0) == onoff.size(0) && roll.get_device() == onoff.get_device(), "roll / onoff do not match");
  want(row_offsets, at::kLong, "row_offsets"); want(sample_offsets, at::kLong, "sample_offsets");
  c10::cuda::CUDAGuard guard(roll.device());
  at::Tensor a = at::empty({n_keys * total_samples}, roll.options().dtype(dtype_of(out_dtype)));
  at::Tensor b = at::empty({n_keys * total_samples}, roll.options().dtype(dtype_of(out_dtype)));
  check(mst_pianoroll_upsample_pair(roll.data_ptr(), onoff.data_ptr(), row_offsets.data_ptr<int64_t>(),
                                    sample_offsets.data_ptr<int64_t>(), (int)(row_offsets.numel() - 1), total_samples, (int)fs,
                                    (int)sr, (int)pitch_lo, (int)n_keys, (int)out_dtype, a.data_ptr(), b.data_ptr(), cur_stream()),
        "mst_pianoroll_upsample_pair");
  return {a, b};
}

at::Tensor griffinlim(const at::Tensor& S, int64_t s_layout, bool s_is_log1p_power, int64_t batch, int64_t n_iter,
                      double momentum, const c10::optional<at::Tensor>& init_phase, int64_t init_mode, int64_t seed) {
  want(S, at::kFloat, "S");
  c10::cuda::CUDAGuard guard(S.device());
  mst_batch_t* b = as_batch(batch);
  TORCH_CHECK(mst_batch_device(b) == (int)S.get_device(), "the batch handle was built on cuda:", mst_batch_device(b),
              " but S lives on cuda:", (int)S.get_device());
  const int64_t K = mst_batch_n_fft(b) / 2 + 1;
  TORCH_CHECK(S.numel() == mst_batch_total_frames(b) * K, "S has ", S.numel(), " elements, batch expects ",
              mst_batch_total_frames(b) * K);
  const float* phase = nullptr;
  if (init_phase.has_value()) {
    want(*init_phase, at::kFloat, "init_phase");
    TORCH_CHECK(init_phase->numel() == S.numel(), "init_phase must match S");
    phase = init_phase->data_ptr<float>();
  }
  at::Tensor y = at::empty({mst_batch_total_samples(b)}, S.options());
  const size_t ws_bytes = mst_griffinlim_workspace_bytes_ex(b, (int)s_layout, s_is_log1p_power ? 1 : 0);
  at::Tensor ws = at::empty({(int64_t)ws_bytes}, S.options().dtype(at::kByte));
  check(mst_griffinlim_f32(S.data_ptr<float>(), (int)s_layout, s_is_log1p_power ? 1 : 0, b, (int)n_iter, (float)momentum,
                           phase, (int)init_mode, (uint64_t)seed, y.data_ptr<float>(), ws.data_ptr(), ws_bytes, cur_stream()),
        "mst_griffinlim_f32");
  return y;
}

at::Tensor mel_to_stft(const at::Tensor& mel, int64_t mel_layout, int64_t batch, int64_t plan, int64_t n_mels, double power,
                       int64_t max_iter, double tol) {
  want(mel, at::kFloat, "mel");
  c10::cuda::CUDAGuard guard(mel.device());
  mst_batch_t* b = as_batch(batch);
  TORCH_CHECK(plan != 0, "null mel inverse plan handle");
  TORCH_CHECK(mst_batch_device(b) == (int)mel.get_device(), "the batch handle lives on another device");
  TORCH_CHECK(mel.numel() == mst_batch_total_frames(b) * n_mels, "mel has ", mel.numel(), " elements, batch expects ",
              mst_batch_total_frames(b) * n_mels);
  at::Tensor S = at::empty({mst_batch_total_frames(b), 1025}, mel.options());
  check(mst_mel_to_stft_f32(mel.data_ptr<float>(), (int)mel_layout, b, reinterpret_cast<mst_mel_inverse_plan_t*>(plan),
                            (float)power, (int)max_iter, (float)tol, S.data_ptr<float>(), cur_stream()), "mst_mel_to_stft_f32");
  return S;
}

at::Tensor spectral_convergence(const at::Tensor& y, int64_t batch, const at::Tensor& S, int64_t s_layout) {
  want(y, at::kFloat, "y");
  want(S, at::kFloat, "S");
  c10::cuda::CUDAGuard guard(y.device());
  mst_batch_t* b = as_batch(batch);
  check_batch_audio(b, y, "y");
  TORCH_CHECK(S.get_device() == y.get_device(), "S and y live on different devices");
  TORCH_CHECK(S.numel() == mst_batch_total_frames(b) * (mst_batch_n_fft(b) / 2 + 1), "S does not match the batch");
  const int64_t n = mst_batch_n_clips(b);
  at::Tensor sums = at::empty({2, n}, y.options().dtype(at::kDouble));
  check(mst_spectral_convergence_f32(y.data_ptr<float>(), b, S.data_ptr<float>(), (int)s_layout, sums.data_ptr<double>(),
                                     sums.data_ptr<double>() + n, cur_stream()), "mst_spectral_convergence_f32");
  return at::sqrt(sums[0] / sums[1]);
}

at::Tensor resample(const at::Tensor& x, int64_t sr_in, int64_t sr_out) {
  want(x, at::kFloat, "x");
  TORCH_CHECK(x.dim() == 1, "resample expects a 1-D mono signal");
  c10::cuda::CUDAGuard guard(x.device());
  const int64_t n = mst_resample_length(x.numel(), (int)sr_in, (int)sr_out);
  TORCH_CHECK(n >= 0, "bad resample arguments");
  at::Tensor y = at::empty({n}, x.options());
  check(mst_resample_f32(x.data_ptr<float>(), x.numel(), (int)sr_in, (int)sr_out, y.data_ptr<float>(), cur_stream()),
        "mst_resample_f32");
  return y;
}

at::Tensor mono_mix(const at::Tensor& frames) {
  want(frames, at::kFloat, "frames");
  TORCH_CHECK(frames.dim() == 2, "mono_mix expects (n_frames, channels) interleaved samples");
  c10::cuda::CUDAGuard guard(frames.device());
  at::Tensor y = at::empty({frames.size(0)}, frames.options());
  check(mst_mono_mix_f32(frames.data_ptr<float>(), frames.size(0), (int)frames.size(1), y.data_ptr<float>(), cur_stream()),
        "mst_mono_mix_f32");
  return y;
}

}  // namespace

TORCH_LIBRARY(mst_b200, m) {
  // host-side handles / helpers (no dispatch key: they take no device tensors)
  m.def("batch_create(Tensor clip_offsets, Tensor clip_lengths, int n_fft, int hop, int pad_mode, int device, int win_length) -> int", &batch_create);
  m.def("batch_create_from_frames(Tensor frames_per_clip, int n_fft, int hop, int pad_mode, int device, int win_length) -> int", &batch_create_from_frames);
  m.def("batch_destroy(int handle) -> ()", &batch_destroy);
  m.def("batch_total_frames(int handle) -> int", &batch_total_frames);
  m.def("batch_total_samples(int handle) -> int", &batch_total_samples);
  m.def("batch_clip_frames(int handle, int clip) -> int", &batch_clip_frames);
  m.def("mel_filterbank(int sr, int n_fft, int n_mels, float fmin, float fmax) -> Tensor", &mel_filterbank);
  m.def("mel_plan_create(Tensor weights, int device) -> int", &mel_plan_create);
  m.def("mel_plan_destroy(int handle) -> ()", &mel_plan_destroy);
  m.def("mel_inverse_plan_create(Tensor weights, int device) -> int", &mel_inverse_plan_create);
  m.def("mel_inverse_plan_destroy(int handle) -> ()", &mel_inverse_plan_destroy);
  m.def("launch_count() -> int", &launch_count);
  // device ops: CUDA implementations only
  m.def("stft(Tensor audio, int batch, int out_mode, int layout) -> Tensor");
  m.def("stft_mel(Tensor audio, int batch, int plan, int n_mels, bool apply_log1p, int layout) -> Tensor");
  m.def("pianoroll_count_rows(Tensor end, Tensor note_offsets, int fs) -> Tensor");
  m.def("pianoroll_rasterize(Tensor pitch, Tensor velocity, Tensor start, Tensor end, Tensor note_offsets, "
        "Tensor row_offsets, int total_rows, int fs, bool want_velsum, Tensor? span_piece, Tensor? span_start, "
        "Tensor? span_end) -> (Tensor, Tensor, Tensor)");
  m.def("pianoroll_merge_instruments(Tensor velsum, Tensor inst_row_offsets, Tensor inst_is_drum, Tensor file_inst_offsets, "
        "Tensor file_row_offsets, int total_rows, Tensor seg_offsets, Tensor segments, bool want_f64) -> (Tensor, Tensor, Tensor)");
  m.def("pianoroll_chunks(Tensor plane, int num_chunks, int chunk_rows, int stride_rows, int out_dtype) -> Tensor");
  m.def("pianoroll_upsample(Tensor plane, Tensor row_offsets, Tensor sample_offsets, int total_samples, int fs, int sr, "
        "int pitch_lo, int n_keys, int out_dtype) -> Tensor");
  m.def("pianoroll_upsample_pair(Tensor roll, Tensor onoff, Tensor row_offsets, Tensor sample_offsets, int total_samples, "
        "int fs, int sr, int pitch_lo, int n_keys, int out_dtype) -> (Tensor, Tensor)");
  m.def("resample(Tensor x, int sr_in, int sr_out) -> Tensor");
  m.def("mono_mix(Tensor frames) -> Tensor");
  m.def("spectral_convergence(Tensor y, int batch, Tensor S, int s_layout) -> Tensor");
  m.def("mel_to_stft(Tensor mel, int mel_layout, int batch, int plan, int n_mels, float power, int max_iter, float tol) -> Tensor");
  m.def("griffinlim(Tensor S, int s_layout, bool s_is_log1p_power, int batch, int n_iter, float momentum, "
        "Tensor? init_phase, int init_mode, int seed) -> Tensor");
}

TORCH_LIBRARY_IMPL(mst_b200, CUDA, m) {
  m.impl("stft", &stft);
  m.impl("stft_mel", &stft_mel);
  m.impl("pianoroll_count_rows", &pianoroll_count_rows);
  m.impl("pianoroll_rasterize", &pianoroll_rasterize);
  m.impl("pianoroll_merge_instruments", &pianoroll_merge_instruments);
  m.impl("pianoroll_chunks", &pianoroll_chunks);
  m.impl("pianoroll_upsample", &pianoroll_upsample);
  m.impl("pianoroll_upsample_pair", &pianoroll_upsample_pair);
  m.impl("griffinlim", &griffinlim);
  m.impl("resample", &resample);
  m.impl("mono_mix", &mono_mix);
  m.impl("spectral_convergence", &spectral_convergence);
  m.impl("mel_to_stft", &mel_to_stft);
}
