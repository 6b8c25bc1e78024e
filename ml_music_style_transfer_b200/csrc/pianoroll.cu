// P3: MIDI notes -> 128-pitch piano roll -> binarise -> on/off -> chunks / audio-rate planes.
// Integer / byte work, bit-exact against pretty_midi's get_piano_roll (notes only) followed by the
// NumPy lines of reference preprocessing/preprocess.py:146-155 and :80-96; the audio-rate
// hold-replication is the README-only step (README.md:19-20) as defined in SURVEY section 8a P3d.
// All of it is HBM-bound: the frame-rate roll is tiny, the audio-rate planes are a pure write stream
// (128-bit coalesced stores).
#include <algorithm>
#include "mst_common.cuh"

namespace mst {

// int(t * fs) as Python evaluates it: IEEE double product (no FMA contraction), truncation toward zero.
__device__ __forceinline__ int64_t col_of(double t, int fs) { return (int64_t)__dmul_rn(t, (double)fs); }

__global__ void count_rows_kernel(const double* __restrict__ end, const int64_t* __restrict__ note_off, int n_pieces,
                                  int fs, int64_t* __restrict__ rows) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n_pieces) return;
  const int64_t a = note_off[warp], b = note_off[warp + 1];
  double m = 0.0;
  bool any = false;
  for (int64_t i = a + lane; i < b; i += 32) {
    const double e = end[i];
    m = any ? fmax(m, e) : e;
    any = true;
  }
  for (int o = 16; o; o >>= 1) {
    const double mo = __shfl_xor_sync(0xffffffffu, m, o);
    const bool ao = __shfl_xor_sync(0xffffffffu, (int)any, o);
    if (ao) { m = any ? fmax(m, mo) : mo; any = true; }
  }
  if (lane == 0) {
    int64_t T = any ? col_of(m, fs) : 0;  // int(fs * end_time)
    rows[warp] = T < 0 ? 0 : T;
  }
}

// One warp per note: mark [int(start*fs), int(end*fs)) in the time-major roll of its piece.
__global__ void rasterize_kernel(const int32_t* __restrict__ pitch, const int32_t* __restrict__ vel,
                                 const double* __restrict__ start, const double* __restrict__ end,
                                 const int64_t* __restrict__ note_off, int n_pieces,
                                 const int64_t* __restrict__ row_off, int64_t total_notes, int fs,
                                 uint8_t* __restrict__ roll, int32_t* __restrict__ velsum) {
  const int64_t note = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (note >= total_notes) return;
  // piece = last p with note_off[p] <= note
  int lo = 0, hi = n_pieces;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (note_off[mid] <= note) lo = mid; else hi = mid;
  }
  const int64_t r0 = row_off[lo], T = row_off[lo + 1] - r0;
  const int p = pitch[note], v = vel[note];
  if (p < 0 || p > 127 || v == 0) return;
  int64_t s = col_of(start[note], fs), e = col_of(end[note], fs);
  if (s < 0) s = 0;  // pretty_midi never produces negative note times
  if (e > T) e = T;  // NumPy slice clipping
  for (int64_t c = s + lane; c < e; c += 32) {
    const int64_t idx = (r0 + c) * 128 + p;
    roll[idx] = 1;
    if (velsum) atomicAdd(velsum + idx, v);
  }
}

// CC64 sustain (pretty_midi >= 0.2.9 get_piano_roll, pedal_threshold=64): inside every pedal-down span [a, b) each
// pitch keeps the running maximum of its velocity sum ("np.maximum.accumulate(subpr, axis=1)").  One thread per
// (span, pitch); consecutive threads touch consecutive pitches, so every column step is one coalesced 512-byte row.
__global__ void sustain_kernel(int32_t* __restrict__ velsum, const int64_t* __restrict__ row_off,
                               const int32_t* __restrict__ span_piece, const int64_t* __restrict__ span_start,
                               const int64_t* __restrict__ span_end, int n_spans) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int span = (int)(gid >> 7), p = (int)(gid & 127);
  if (span >= n_spans) return;
  const int piece = span_piece[span];
  const int64_t r0 = row_off[piece], T = row_off[piece + 1] - r0;
  int64_t a = span_start[span], b = span_end[span];
  if (a < 0) a = 0;
  if (b > T) b = T;  // NumPy slice clipping
  int32_t run = 0;
  for (int64_t c = a; c < b; ++c) {
    int32_t* cell = velsum + (r0 + c) * 128 + p;
    const int32_t v = *cell;
    run = c == a ? v : max(run, v);
    *cell = run;
  }
}

// roll = (velsum != 0)  (preprocess.py:148)
__global__ void binarize_kernel(const int32_t* __restrict__ velsum, int64_t n, uint8_t* __restrict__ roll) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    roll[i] = velsum[i] != 0 ? 1 : 0;
}

// onoff[i] = roll[i] - roll[i-1] with a zero row before the first row of each piece (preprocess.py:149-155).
__global__ void onoff_kernel(const uint8_t* __restrict__ roll, const int64_t* __restrict__ row_off, int n_pieces,
                             int8_t* __restrict__ onoff) {
  for (int p = blockIdx.y; p < n_pieces; p += gridDim.y) {
    const int64_t r0 = row_off[p], T = row_off[p + 1] - r0;
    const uint4* src = reinterpret_cast<const uint4*>(roll + r0 * 128);
    uint4* dst = reinterpret_cast<uint4*>(onoff + r0 * 128);
    const int64_t nvec = T * 8;  // 128 bytes per row = 8 x 16 B
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
      const uint4 cur = src[i];
      uint4 prev = make_uint4(0, 0, 0, 0);
      if (i >= 8) prev = src[i - 8];
      uint4 o;
      // bytes are 0/1: per-byte subtraction without borrow across lanes -> __vsub4
      o.x = __vsub4(cur.x, prev.x); o.y = __vsub4(cur.y, prev.y);
      o.z = __vsub4(cur.z, prev.z); o.w = __vsub4(cur.w, prev.w);
      dst[i] = o;
    }
  }
}

// PrettyMIDI.get_piano_roll over several instruments (pretty_midi 0.2.9): every instrument's velocity-sum roll (already
// sustained) gets its pitch bends applied -- each bend segment [c0, c1) shifts the rows by the integer part of the bend
// and linearly interpolates by the fractional part, in float64 with NumPy's rounding sequence (two products, one sum,
// no FMA) -- and the instruments of a file are then summed, in instrument order, into the widest roll.  Drum
// instruments contribute zeros (but their width counts).  One thread per (file row, pitch): a column of an instrument
// lies in at most one bend segment, so every output cell is a pure function of the un-bent column.
struct BendSeg {
  int64_t c0, c1;   // columns [c0, c1) of the instrument's roll
  double d, m1;     // bend_decimal and (1 - bend_decimal), both evaluated on the host in float64
  int32_t bi;       // bend_int (signed)
  int32_t positive; // start_bend.pitch >= 0
};

__device__ __forceinline__ double bent_value(const int32_t* __restrict__ col, int r, const BendSeg& sg) {
  // B[q] = piano_roll[q - bi] where that row exists (the shifted copy), 0 elsewhere
  auto B = [&](int q) -> double {
    const int src = q - sg.bi;
    if (sg.bi > 0 && q < sg.bi) return 0.0;
    if (sg.bi < 0 && q >= 128 + sg.bi) return 0.0;
    return (src >= 0 && src < 128) ? (double)col[src] : 0.0;
  };
  if (sg.positive) {
    if (r == 0) return B(0);
    return __dadd_rn(__dmul_rn(sg.m1, B(r)), __dmul_rn(sg.d, B(r - 1)));
  }
  if (r == 127) return B(127);
  return __dadd_rn(__dmul_rn(sg.m1, B(r)), __dmul_rn(sg.d, B(r + 1)));
}

__global__ void merge_instruments_kernel(const int32_t* __restrict__ velsum, const int64_t* __restrict__ inst_row_off,
                                         const int32_t* __restrict__ inst_is_drum, const int32_t* __restrict__ file_inst_off,
                                         const int64_t* __restrict__ file_row_off, int n_files,
                                         const int32_t* __restrict__ seg_off, const BendSeg* __restrict__ segs,
                                         int64_t total_rows, double* __restrict__ out_f64, uint8_t* __restrict__ roll) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t row = gid >> 7;
  const int p = (int)(gid & 127);
  if (row >= total_rows) return;
  int lo = 0, hi = n_files;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (file_row_off[mid] <= row) lo = mid; else hi = mid;
  }
  const int64_t col = row - file_row_off[lo];
  double acc = 0.0;
  for (int i = file_inst_off[lo]; i < file_inst_off[lo + 1]; ++i) {
    const int64_t r0 = inst_row_off[i], T = inst_row_off[i + 1] - r0;
    if (col >= T) continue;
    double v = 0.0;
    if (!inst_is_drum[i]) {
      const int32_t* c = velsum + (r0 + col) * 128;
      int a = seg_off[i], b = seg_off[i + 1];  // segments are sorted and disjoint: find the one with c0 <= col < c1
      const BendSeg* hit = nullptr;
      while (a < b) {
        const int m = (a + b) >> 1;
        if (segs[m].c1 <= col) a = m + 1;
        else if (segs[m].c0 > col) b = m;
        else { hit = segs + m; break; }
      }
      v = hit ? bent_value(c, p, *hit) : (double)c[p];
    }
    acc = __dadd_rn(acc, v);
  }
  if (out_f64) out_f64[row * 128 + p] = acc;
  roll[row * 128 + p] = acc != 0.0 ? 1 : 0;
}

template <typename OUT>
__global__ void chunks_kernel(const int8_t* __restrict__ plane, int64_t n_rows, int num_chunks, int chunk_rows,
                              int stride_rows, OUT* __restrict__ out) {
  const int64_t total = (int64_t)num_chunks * chunk_rows * 128;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int p = (int)(i & 127);
    const int64_t rj = i >> 7;
    const int64_t c = rj / chunk_rows, j = rj - c * chunk_rows;
    const int64_t row = c * stride_rows + j;
    out[i] = row < n_rows ? (OUT)plane[row * 128 + p] : (OUT)0;
  }
}

// Audio-rate hold-replication: out[k][n] = plane[(n*fs)/sr][pitch_lo + k], for ONE or TWO planes (roll + on/off) per
// launch so that the index arithmetic is done once.  One warp owns a chunk of 32 x kUpVec 16-byte vectors of one key row
// (4096 int8 / 1024 float samples): fully coalesced 128-bit stores.  The (few) roll columns a chunk can touch are fetched
// once, one or two bytes per lane and plane, and every vector's current / next column value then comes from a warp
// shuffle -- no dependent global load inside the store loop.  A piano roll is mostly runs: when all staged columns of a
// chunk hold the same value (a held or a silent key -- the common case) the chunk is eight plain vector stores of that
// value, no per-vector arithmetic at all; otherwise the column / remainder pair is advanced incrementally (ONE division
// per chunk) and each vector is split at its single column crossing with a byte mask (the number of samples left in
// the current column comes from a multiply-high by a precomputed reciprocal of fs).  Rows need not start 16-byte
// aligned: scalar head / tail elements are written by the warp that owns chunk 0.  If one vector could span more than
// two columns (fs * EPV > sr), or a chunk more than 62 columns, a generic loop is used.
constexpr int kUpVec = 8;  // vectors per lane per chunk

__device__ __forceinline__ unsigned shl_clamp(unsigned x, unsigned n) {  // PTX shl: shift amounts >= 32 give 0
  unsigned r;
  asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(n));
  return r;
}

template <typename OUT>
__device__ __forceinline__ uint4 make_vector(int8_t cur, int8_t nxt, int e_cross) {
  constexpr int EPV = 16 / sizeof(OUT);
  uint4 out;
  if (sizeof(OUT) == 1) {
    // bytes [0, e_cross) come from cur, the rest from nxt:  w = n4 ^ ((c4 ^ n4) & low_bytes(k)),  k = e_cross - 4*i
    const unsigned c4 = __byte_perm((unsigned)(uint8_t)cur, 0u, 0x0000), n4 = __byte_perm((unsigned)(uint8_t)nxt, 0u, 0x0000);
    const unsigned d = c4 ^ n4;
    unsigned w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int bits = 8 * e_cross - 32 * i;
      const unsigned m = shl_clamp(1u, (unsigned)(bits < 0 ? 0 : bits)) - 1u;
      w[i] = n4 ^ (d & m);
    }
    out = make_uint4(w[0], w[1], w[2], w[3]);
  } else {
    OUT vals[EPV];
#pragma unroll
    for (int e = 0; e < EPV; ++e) vals[e] = (OUT)(e < e_cross ? cur : nxt);
    out = *reinterpret_cast<const uint4*>(vals);
  }
  return out;
}

template <typename OUT>
__device__ __forceinline__ uint4 splat_vector(int8_t v) {
  constexpr int EPV = 16 / sizeof(OUT);
  if (sizeof(OUT) == 1) {
    const unsigned c4 = __byte_perm((unsigned)(uint8_t)v, 0u, 0x0000);
    return make_uint4(c4, c4, c4, c4);
  }
  OUT vals[EPV];
#pragma unroll
  for (int e = 0; e < EPV; ++e) vals[e] = (OUT)v;
  return *reinterpret_cast<const uint4*>(vals);
}

// 4 CTAs x 256 threads per SM (<= 64 registers): the kernel waits on its two staged byte loads per chunk (ncu: 60 % of the
// stall samples), so resident warps are what hides them.  Measured at 4096 clips (profiles/r2_ab_upsample.log): no cap
// (78 registers, 3 CTAs) 5.31 TB/s; cap 4: 5.80; cap 5 / 6 / 8 (spills): 5.61 / 5.54 / 4.76; prefetching the next chunk's
// bytes before the stores of the current one: 5.57 with the cap, 4.91 without -- the cap alone wins.
template <typename OUT, int NP>
__global__ void __launch_bounds__(256, 4) upsample_kernel(const int8_t* __restrict__ plane0, const int8_t* __restrict__ plane1,
                                                       const int64_t* __restrict__ row_off, const int64_t* __restrict__ samp_off,
                                                       int n_pieces, int fs, int sr, int pitch_lo, int n_keys,
                                                       OUT* __restrict__ out0, OUT* __restrict__ out1) {
  constexpr int EPV = 16 / sizeof(OUT);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warps_per_cta = blockDim.x >> 5;
  const bool simple = (int64_t)fs * EPV <= sr;                // at most one column crossing per vector
  const unsigned fs_magic = (unsigned)(0x100000000ull / (unsigned)fs) + 1u;  // exact ceil-div by fs for numerators < 2^32/fs
  const int step_col = (32 * EPV * fs) / sr, step_rem = (32 * EPV * fs) % sr;  // advance of 32 vectors
  const int span_cols = (int)(((int64_t)32 * kUpVec * EPV * fs) / sr) + 2;     // columns one chunk can touch (upper bound)
  const bool staged = simple && span_cols <= 62;                               // they fit the 64 staged columns
  const int8_t* planes[2] = {plane0, NP > 1 ? plane1 : plane0};
  OUT* outs[2] = {out0, NP > 1 ? out1 : out0};
  for (int piece = blockIdx.z; piece < n_pieces; piece += gridDim.z) {
    const int64_t r0 = row_off[piece], T = row_off[piece + 1] - r0;
    const int64_t N = samp_off[piece + 1] - samp_off[piece];
    for (int k = blockIdx.y; k < n_keys; k += gridDim.y) {
      const int64_t src_off = r0 * 128 + pitch_lo + k;
      const int64_t row_base = samp_off[piece] * n_keys + (int64_t)k * N;
      // both outputs share the misalignment of their rows (the host checks that the two bases agree mod 16)
      const int64_t mis = (int64_t)(((16 - (reinterpret_cast<uintptr_t>(out0 + row_base) & 15)) & 15) / sizeof(OUT));
      const int64_t head = mis < N ? mis : N;
      const int64_t nvec = (N - head) / EPV;
      const int64_t n_chunks = (nvec + 32 * kUpVec - 1) / (32 * kUpVec);
#ifndef MST_UP_ROWPTR
#define MST_UP_ROWPTR 1
#endif
      // vector-aligned start of this key row in both outputs, and the row's first roll byte in both planes: computed once
      // per row and kept opaque (otherwise ptxas re-derives them from the kernel parameters in front of every store)
      OUT* vrow[2] = {out0 + row_base + head, (NP > 1 ? out1 : out0) + row_base + head};
      const int8_t* srow[2] = {plane0 + src_off, (NP > 1 ? plane1 : plane0) + src_off};
#if MST_UP_ROWPTR
      asm volatile("" : "+l"(vrow[0]), "+l"(vrow[1]), "+l"(srow[0]), "+l"(srow[1]));
#endif
      for (int64_t chunk = (int64_t)blockIdx.x * warps_per_cta + warp; chunk < (n_chunks > 0 ? n_chunks : 1);
           chunk += (int64_t)gridDim.x * warps_per_cta) {
        if (chunk == 0) {
          // scalar head and tail of the row
          const int64_t tail0 = head + nvec * EPV;
#pragma unroll
          for (int q = 0; q < NP; ++q) {
            const int8_t* src = planes[q] + src_off;
            OUT* row = outs[q] + row_base;
            for (int64_t n = lane; n < head; n += 32) {
              const int64_t col = (n * fs) / sr;
              row[n] = (OUT)(col < T ? src[col * 128] : (int8_t)0);
            }
            for (int64_t n = tail0 + lane; n < N; n += 32) {
              const int64_t col = (n * fs) / sr;
              row[n] = (OUT)(col < T ? src[col * 128] : (int8_t)0);
            }
          }
        }
        int64_t v = chunk * (32 * kUpVec) + lane;
        // column / remainder of lane 0's first sample by ONE 64-bit division per warp, the other lanes by a 32-bit one
        const int64_t prod0 = (head + chunk * (32 * kUpVec) * EPV) * fs;
        const int64_t c0 = prod0 / sr;
        const unsigned t0 = (unsigned)(prod0 - c0 * sr) + (unsigned)(lane * EPV * fs);
        int64_t col = c0 + t0 / (unsigned)sr;
        int rem = (int)(t0 % (unsigned)sr);
        if (staged) {
          int b0 = 0, b1 = 0;   // staged columns c0 + lane and c0 + 32 + lane: byte q = plane q
#pragma unroll
          for (int q = 0; q < NP; ++q) {
            const int8_t* src = srow[q];
            const int x0 = c0 + lane < T ? (int)(uint8_t)src[(c0 + lane) * 128] : 0;
            const int x1 = c0 + 32 + lane < T ? (int)(uint8_t)src[(c0 + 32 + lane) * 128] : 0;
            b0 |= x0 << (8 * q);
            b1 |= x1 << (8 * q);
          }
          // uniform chunk: every column this chunk can touch holds the same value(s)
          const int first = __shfl_sync(0xffffffffu, b0, 0);
          const bool same = (lane >= span_cols || b0 == first) && (lane + 32 >= span_cols || b1 == first);
          if (__all_sync(0xffffffffu, same)) {
#pragma unroll
            for (int q = 0; q < NP; ++q) {
              const uint4 val = splat_vector<OUT>((int8_t)((first >> (8 * q)) & 0xff));
              OUT* dst = vrow[q];
#pragma unroll
              for (int u = 0; u < kUpVec; ++u)
                if (v + 32 * u < nvec) *reinterpret_cast<uint4*>(dst + (v + 32 * u) * EPV) = val;
            }
            continue;
          }
#pragma unroll
          for (int u = 0; u < kUpVec; ++u) {
            int rel = (int)(col - c0);
            rel = rel < 0 ? 0 : (rel > 62 ? 62 : rel);
            const int lo0 = __shfl_sync(0xffffffffu, b0, rel & 31), hi0 = __shfl_sync(0xffffffffu, b1, rel & 31);
            const int lo1 = __shfl_sync(0xffffffffu, b0, (rel + 1) & 31), hi1 = __shfl_sync(0xffffffffu, b1, (rel + 1) & 31);
            const int cur2 = rel < 32 ? lo0 : hi0;
            const int nxt2 = rel + 1 < 32 ? lo1 : hi1;
            const int e_cross = (int)__umulhi((unsigned)(sr - rem + fs - 1), fs_magic);
            if (v < nvec) {
#pragma unroll
              for (int q = 0; q < NP; ++q)
                *reinterpret_cast<uint4*>(vrow[q] + v * EPV) =
                    make_vector<OUT>((int8_t)((cur2 >> (8 * q)) & 0xff), (int8_t)((nxt2 >> (8 * q)) & 0xff), e_cross);
            }
            v += 32;
            col += step_col;
            rem += step_rem;
            if (rem >= sr) { rem -= sr; ++col; }
          }
          continue;
        }
#pragma unroll
        for (int u = 0; u < kUpVec; ++u) {
          if (v >= nvec) break;
#pragma unroll
          for (int q = 0; q < NP; ++q) {
            const int8_t* src = srow[q];
            OUT* dst = vrow[q] + v * EPV;
            const int8_t cur = col < T ? src[col * 128] : (int8_t)0;
            if (simple) {
              const int8_t nxt = col + 1 < T ? src[(col + 1) * 128] : (int8_t)0;
              // samples e with rem + e*fs < sr stay in `col`: e_cross = ceil((sr - rem) / fs)
              const int e_cross = (int)__umulhi((unsigned)(sr - rem + fs - 1), fs_magic);
              *reinterpret_cast<uint4*>(dst) = make_vector<OUT>(cur, nxt, e_cross);
            } else {
              OUT vals[EPV];
              int64_t c2 = col;
              int r2 = rem;
              int8_t cv = cur;
#pragma unroll 1
              for (int e = 0; e < EPV; ++e) {
                vals[e] = (OUT)cv;
                r2 += fs;
                if (r2 >= sr) {
                  do { r2 -= sr; ++c2; } while (r2 >= sr);
                  cv = c2 < T ? src[c2 * 128] : (int8_t)0;
                }
              }
              *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(vals);
            }
          }
          v += 32;
          col += step_col;
          rem += step_rem;
          if (rem >= sr) { rem -= sr; ++col; }
        }
      }
    }
  }
}

}  // namespace mst

using namespace mst;

extern "C" {

int mst_pianoroll_count_rows(const double* d_end, const int64_t* d_note_offsets, int n_pieces, int fs,
                             int64_t* d_rows_out, mst_stream_t stream) {
  if (!d_end || !d_note_offsets || !d_rows_out) return fail(MST_ERR_INVALID, "null argument");
  if (n_pieces <= 0 || fs <= 0) return fail(MST_ERR_INVALID, "n_pieces and fs must be positive");
  const int threads = 256, warps_per_block = threads / 32;
  count_rows_kernel<<<(n_pieces + warps_per_block - 1) / warps_per_block, threads, 0,
                      reinterpret_cast<cudaStream_t>(stream)>>>(d_end, d_note_offsets, n_pieces, fs, d_rows_out);
  MST_CUDA_OK(cudaGetLastError());
  count_launch();
  return MST_OK;
}

int mst_pianoroll_rasterize_sustain(const int32_t* d_pitch, const int32_t* d_velocity, const double* d_start,
                                    const double* d_end, const int64_t* d_note_offsets, int n_pieces,
                                    const int64_t* d_row_offsets, int64_t total_rows, int64_t total_notes, int fs,
                                    const int32_t* d_span_piece, const int64_t* d_span_start, const int64_t* d_span_end,
                                    int n_spans, uint8_t* d_roll, int8_t* d_onoff, int32_t* d_velsum,
                                    mst_stream_t stream) {
  if (n_spans < 0 || (n_spans > 0 && (!d_span_piece || !d_span_start || !d_span_end || !d_velsum)))
    return fail(MST_ERR_INVALID, "sustain spans need span arrays and a velocity-sum buffer");
  int rc = mst_pianoroll_rasterize(d_pitch, d_velocity, d_start, d_end, d_note_offsets, n_pieces, d_row_offsets, total_rows,
                                   total_notes, fs, d_roll, d_onoff, d_velsum, stream);
  if (rc || n_spans == 0 || total_rows == 0) return rc;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int64_t threads = (int64_t)n_spans * 128;
  sustain_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(d_velsum, d_row_offsets, d_span_piece, d_span_start,
                                                                  d_span_end, n_spans);
  MST_CUDA_OK(cudaGetLastError());
  const int64_t n = total_rows * 128;
  binarize_kernel<<<(unsigned)std::min<int64_t>((n + 255) / 256, 148 * 16), 256, 0, s>>>(d_velsum, n, d_roll);
  MST_CUDA_OK(cudaGetLastError());
  dim3 grid(64, (unsigned)std::min(n_pieces, 65535));
  onoff_kernel<<<grid, 256, 0, s>>>(d_roll, d_row_offsets, n_pieces, d_onoff);
  MST_CUDA_OK(cudaGetLastError());
  count_launch(3);
  return MST_OK;
}

int mst_pianoroll_rasterize(const int32_t* d_pitch, const int32_t* d_velocity, const double* d_start,
                            const double* d_end, const int64_t* d_note_offsets, int n_pieces,
                            const int64_t* d_row_offsets, int64_t total_rows, int64_t total_notes, int fs,
                            uint8_t* d_roll, int8_t* d_onoff, int32_t* d_velsum, mst_stream_t stream) {
  if (n_pieces <= 0 || fs <= 0 || total_rows < 0 || total_notes < 0) return fail(MST_ERR_INVALID, "bad sizes");
  if (total_rows == 0) return MST_OK;  // every piece rounds to zero columns: empty rolls, nothing to launch
  if (!d_note_offsets || !d_row_offsets || !d_roll || !d_onoff ||
      (total_notes > 0 && (!d_pitch || !d_velocity || !d_start || !d_end)))
    return fail(MST_ERR_INVALID, "null argument");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (reinterpret_cast<uintptr_t>(d_roll) & 15 || reinterpret_cast<uintptr_t>(d_onoff) & 15)
    return fail(MST_ERR_INVALID, "roll / onoff must be 16-byte aligned");
  MST_CUDA_OK(cudaMemsetAsync(d_roll, 0, (size_t)total_rows * 128, s));
  if (d_velsum) MST_CUDA_OK(cudaMemsetAsync(d_velsum, 0, (size_t)total_rows * 128 * sizeof(int32_t), s));
  if (total_notes > 0) {
    const int threads = 256;
    const int64_t blocks = (total_notes * 32 + threads - 1) / threads;
    rasterize_kernel<<<(unsigned)blocks, threads, 0, s>>>(d_pitch, d_velocity, d_start, d_end, d_note_offsets, n_pieces,
                                                           d_row_offsets, total_notes, fs, d_roll, d_velsum);
    MST_CUDA_OK(cudaGetLastError());
    count_launch();
  }
  dim3 grid(64, (unsigned)std::min(n_pieces, 65535));
  onoff_kernel<<<grid, 256, 0, s>>>(d_roll, d_row_offsets, n_pieces, d_onoff);
  MST_CUDA_OK(cudaGetLastError());
  count_launch();
  return MST_OK;
}

int mst_pianoroll_merge_instruments(const int32_t* d_velsum, const int64_t* d_inst_row_offsets, const int32_t* d_inst_is_drum,
                                    int n_instruments, const int32_t* d_file_inst_offsets, const int64_t* d_file_row_offsets,
                                    int n_files, int64_t total_rows, const int32_t* d_seg_offsets, const void* d_segments,
                                    double* d_out_f64, uint8_t* d_roll, int8_t* d_onoff, mst_stream_t stream) {
  if (n_files <= 0 || n_instruments < 0 || total_rows < 0) return fail(MST_ERR_INVALID, "bad sizes");
  if (total_rows == 0) return MST_OK;
  if (!d_inst_row_offsets || !d_inst_is_drum || !d_file_inst_offsets || !d_file_row_offsets || !d_seg_offsets || !d_roll ||
      !d_onoff || (!d_velsum && n_instruments > 0))
    return fail(MST_ERR_INVALID, "null argument");
  static_assert(sizeof(BendSeg) == 40, "BendSeg layout is part of the ABI (include/mst_b200.h: mst_bend_segment_t)");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int64_t threads = total_rows * 128;
  merge_instruments_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(
      d_velsum, d_inst_row_offsets, d_inst_is_drum, d_file_inst_offsets, d_file_row_offsets, n_files, d_seg_offsets,
      reinterpret_cast<const BendSeg*>(d_segments), total_rows, d_out_f64, d_roll);
  MST_CUDA_OK(cudaGetLastError());
  dim3 grid(64, (unsigned)std::min(n_files, 65535));
  onoff_kernel<<<grid, 256, 0, s>>>(d_roll, d_file_row_offsets, n_files, d_onoff);
  MST_CUDA_OK(cudaGetLastError());
  count_launch(2);
  return MST_OK;
}

int mst_pianoroll_chunks(const void* d_plane, int64_t n_rows, int num_chunks, int chunk_rows, int stride_rows,
                         int out_dtype, void* d_out, mst_stream_t stream) {
  if (num_chunks < 0 || chunk_rows <= 0 || stride_rows <= 0 || n_rows < 0) return fail(MST_ERR_INVALID, "bad sizes");
  if (num_chunks == 0) return MST_OK;  // empty result (np.array([]) in the reference): nothing to launch
  if (!d_out || (!d_plane && n_rows > 0)) return fail(MST_ERR_INVALID, "null argument");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int64_t total = (int64_t)num_chunks * chunk_rows * 128;
  const unsigned blocks = (unsigned)std::min<int64_t>((total + 255) / 256, 148 * 16);
  const int8_t* src = reinterpret_cast<const int8_t*>(d_plane);
  switch (out_dtype) {
    case MST_DTYPE_I8: chunks_kernel<int8_t><<<blocks, 256, 0, s>>>(src, n_rows, num_chunks, chunk_rows, stride_rows, (int8_t*)d_out); break;
    case MST_DTYPE_F32: chunks_kernel<float><<<blocks, 256, 0, s>>>(src, n_rows, num_chunks, chunk_rows, stride_rows, (float*)d_out); break;
    case MST_DTYPE_F64: chunks_kernel<double><<<blocks, 256, 0, s>>>(src, n_rows, num_chunks, chunk_rows, stride_rows, (double*)d_out); break;
    default: return fail(MST_ERR_INVALID, "bad out_dtype %d", out_dtype);
  }
  MST_CUDA_OK(cudaGetLastError());
  count_launch();
  return MST_OK;
}

static int launch_upsample(const void* d_plane0, const void* d_plane1, const int64_t* d_row_offsets,
                           const int64_t* d_sample_offsets, int n_pieces, int64_t total_samples, int fs, int sr, int pitch_lo,
                           int n_keys, int out_dtype, void* d_out0, void* d_out1, mst_stream_t stream) {
  if (n_pieces <= 0 || fs <= 0 || sr <= 0 || pitch_lo < 0 || n_keys <= 0 || pitch_lo + n_keys > 128)
    return fail(MST_ERR_INVALID, "bad upsample geometry");
  if ((int64_t)32 * 16 * fs + sr >= ((int64_t)1 << 32)) return fail(MST_ERR_INVALID, "fs / sr too large");
  if (total_samples <= 0) return MST_OK;
  if (!d_row_offsets || !d_sample_offsets || !d_out0) return fail(MST_ERR_INVALID, "null argument");
  // d_plane may be NULL only when every piece has an empty roll (all output samples are then zero)
  const bool two = d_out1 != nullptr;
  if (two && ((reinterpret_cast<uintptr_t>(d_out0) ^ reinterpret_cast<uintptr_t>(d_out1)) & 15))
    return fail(MST_ERR_INVALID, "the two output planes must have the same alignment modulo 16 bytes");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  const int64_t avg = total_samples / n_pieces + 1;
  const int epv = out_dtype == MST_DTYPE_I8 ? 16 : 4;
  const int64_t chunks = (avg / epv + 32 * kUpVec - 1) / (32 * kUpVec);  // warp-chunks per key row (average piece)
  // each CTA (8 warps) walks up to 4 rounds of chunks of its row when the grid is large enough to fill the GPU anyway
  // (measured: 4 / 8 / 16 / 32 / 64 chunks per CTA -> 3.7 / 4.5 / 5.0 / 5.3 / 5.3 TB/s)
  const int64_t rows = (int64_t)n_keys * n_pieces;
  int64_t per_cta = rows >= 8 * 148 ? 32 : 8;
  const unsigned gx = (unsigned)std::max<int64_t>(1, std::min<int64_t>(64, (chunks + per_cta - 1) / per_cta));
  dim3 grid(gx, (unsigned)n_keys, (unsigned)std::min(n_pieces, 65535));
  const int8_t* p0 = reinterpret_cast<const int8_t*>(d_plane0);
  const int8_t* p1 = reinterpret_cast<const int8_t*>(d_plane1);
  switch (out_dtype) {
    case MST_DTYPE_I8:
      if (two) upsample_kernel<int8_t, 2><<<grid, 256, 0, s>>>(p0, p1, d_row_offsets, d_sample_offsets, n_pieces, fs, sr, pitch_lo, n_keys, (int8_t*)d_out0, (int8_t*)d_out1);
      else upsample_kernel<int8_t, 1><<<grid, 256, 0, s>>>(p0, p0, d_row_offsets, d_sample_offsets, n_pieces, fs, sr, pitch_lo, n_keys, (int8_t*)d_out0, (int8_t*)d_out0);
      break;
    case MST_DTYPE_F32:
      if (two) upsample_kernel<float, 2><<<grid, 256, 0, s>>>(p0, p1, d_row_offsets, d_sample_offsets, n_pieces, fs, sr, pitch_lo, n_keys, (float*)d_out0, (float*)d_out1);
      else upsample_kernel<float, 1><<<grid, 256, 0, s>>>(p0, p0, d_row_offsets, d_sample_offsets, n_pieces, fs, sr, pitch_lo, n_keys, (float*)d_out0, (float*)d_out0);
      break;
    default: return fail(MST_ERR_UNSUPPORTED, "upsample out_dtype must be int8 or float32");
  }
  MST_CUDA_OK(cudaGetLastError());
  count_launch();
  return MST_OK;
}

int mst_pianoroll_upsample(const void* d_plane, const int64_t* d_row_offsets, const int64_t* d_sample_offsets,
                           int n_pieces, int64_t total_samples, int fs, int sr, int pitch_lo, int n_keys, int out_dtype,
                           void* d_out, mst_stream_t stream) {
  return launch_upsample(d_plane, nullptr, d_row_offsets, d_sample_offsets, n_pieces, total_samples, fs, sr, pitch_lo, n_keys,
                         out_dtype, d_out, nullptr, stream);
}

int mst_pianoroll_upsample_pair(const void* d_roll, const void* d_onoff, const int64_t* d_row_offsets,
                                const int64_t* d_sample_offsets, int n_pieces, int64_t total_samples, int fs, int sr,
                                int pitch_lo, int n_keys, int out_dtype, void* d_out_roll, void* d_out_onoff,
                                mst_stream_t stream) {
  if (total_samples > 0 && !d_out_onoff) return fail(MST_ERR_INVALID, "null argument");
  return launch_upsample(d_roll, d_onoff, d_row_offsets, d_sample_offsets, n_pieces, total_samples, fs, sr, pitch_lo, n_keys,
                         out_dtype, d_out_roll, d_out_onoff, stream);
}

}  // extern "C"
