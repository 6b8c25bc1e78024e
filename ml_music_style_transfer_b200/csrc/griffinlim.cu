// P4: Griffin-Lim phase reconstruction (librosa.griffinlim semantics; reference model/inference.py:105-110,
// tests/test_griffinlim.py:23, classic loop at model/inference.py:131-154 == momentum 0).
//
// One kernel launch per iteration.  Launch j (j = 1..n_iter) does, per frame (one warp each):
//   re-analysis   rebuilt_j = rfft(window * reflect_pad(y_{j-1})[t*hop : t*hop+2048])
//   re-projection angles_j  = (rebuilt_j - a*rebuilt_{j-1}) / (|.| + 1e-16),  a = momentum / (1 + momentum)
//   synthesis     y_j      += window * irfft(S * angles_j)            (overlap-add)
// which is librosa's loop rotated by half an iteration: the launch before the first one synthesises y_0 from the
// initial phase, and y_{n_iter} is librosa's final istft(S * angles).  State resident in HBM between launches:
// the previous iterate rebuilt_{j-1} (complex64, read then overwritten in place by the same thread) and three
// rotating overlap-add accumulators (read y_{j-1}, accumulate y_j, zero the one the next launch accumulates into).
// The phase itself never touches HBM: it is recomputed from rebuilt_j and rebuilt_{j-1} in registers.
//
// Overlap-add: the frames of a tile (kWarpsPerCta = 8) are summed in shared memory in frame order, hop-block by hop-block.  Blocks that
// only this tile touches are written with plain 128-bit stores; the blocks a tile shares with its neighbour go to the
// global accumulator as 128-bit vector reductions (red.global.add.v4.f32).  At most two tiles touch a shared sample
// (hop >= 256) and the shared region was zeroed one launch earlier, so the result does not depend on the order the two
// adds land in (a + b == b + a): runs are reproducible.
//
// Large batches: one launch per iteration (persistent CTAs walk the tiles).  Small batches (every tile resident at once,
// e.g. a single clip): gl_persistent_kernel runs the whole loop in ONE cooperative launch with a grid-wide barrier
// between iterations.
#include <algorithm>
#include <atomic>
#include "fft_warp.cuh"
#include "mst_common.cuh"

namespace mst {

struct GlParams {
  const ClipDesc* clips;
  const int32_t* tile_clip;
  int total_tiles;
  int hop;
  int hop_shift;           // log2(hop) when hop is a power of two, else -1
  int q_full;              // 2047 / hop: frames t in [q_full, T-1-q_full] see the periodic interior envelope
  int pad_mode;
  Tables tabs;
  const float* wq;         // [2048] analysis window * interior 1/window-sum-square (period hop)
  const float* S;          // [total_frames][1025] magnitudes, frame-major
  float2* tprev;           // [total_frames][1025] previous rebuilt spectrum
  const float* inv_wss;    // 1 / window-sum-square envelopes (edge frames)
  const int64_t* wss_off;  // per clip offset into inv_wss
  const float* acc_in;     // y_{j-1} (un-normalised overlap-add sums)
  float* acc_out;          // y_j accumulator (zero on entry)
  float* acc_zero;         // accumulator of launch j+1: zeroed here
  float alpha;             // momentum / (1 + momentum)
  int first_iter;          // rebuilt_0 = 0: skip the tprev read
  int last_iter;           // nobody reads rebuilt_{n_iter}: skip the tprev write
  // init launch only
  const float* init_phase; // uniform [0,1) field or NULL
  int phase_layout;        // layout of init_phase
  int init_mode;           // 0 random (device RNG when init_phase == NULL), 1 angles = 1
  unsigned long long seed;
};

__device__ __forceinline__ float uniform_hash(unsigned long long seed, unsigned long long idx) {
  // counter-based generator for the random initial phase (librosa draws it from NumPy's global RNG, i.e. it is
  // unspecified): two rounds of 32-bit multiply-xorshift mixing (murmur3 / "lowbias32" finalisers) over the 64-bit element
  // index and the seed -> 24-bit mantissa uniform in [0,1).  All 32-bit: 64-bit multiplies cost 4-6 IMADs each and the
  // initial-synthesis launch was spending more time here than a full iteration does.
  unsigned x = (unsigned)idx * 0x9E3779B1u ^ ((unsigned)(idx >> 32) + 0x7F4A7C15u) * 0x85EBCA77u ^ (unsigned)seed;
  x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
  x += (unsigned)(seed >> 32) * 0x27D4EB2Fu + 0x165667B1u;
  x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
  return (float)(x >> 8) * (1.0f / 16777216.0f);
}

// Inverse transform of one frame's spectrum (mirror layout), synthesis window (1/1024 folded in), and park the 2048
// samples in this warp's slot.
__device__ __forceinline__ void synthesize_frame(float2 (&y)[32], float2 mid, float2* scratch, const float2* s_tw1024,
                                                 const float2* s_twp, const float* s_wsyn, int lane) {
  float2 v[32];
  irfft2048_warp(y, mid, v, scratch, s_tw1024, s_twp, lane);
  __syncwarp();
  const float2* w2 = reinterpret_cast<const float2*>(s_wsyn);
#pragma unroll
  for (int r = 0; r < 32; ++r) {
    scratch[lane + 32 * r] = pk_mul(v[br5(r)], w2[lane + 32 * r]);  // samples 2m, 2m+1
  }
}

constexpr int kSlot = 2 * kScratchPerWarp;  // floats between consecutive frame slots

// Overlap-add of one tile when hop divides n_fft (R = n_fft / hop frames overlap every sample): thread j walks the
// span hop-block by hop-block; block b receives frames f = b-R+1 .. b, summed in increasing f.
__device__ __forceinline__ void red_add_v4(float* addr, float4 v) {
  // one 128-bit reduction instead of four 32-bit ones (sm_90+): the atomic path is LSU-issue bound per lane
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Overlap-add of one tile when hop divides n_fft (R = n_fft / hop frames overlap every sample).  The span is walked
// hop-block by hop-block, four samples per thread; block b receives frames f = b-R+1 .. b, summed in increasing f.
// Blocks b < R-1 are shared with the previous tile of the clip and blocks b >= nvalid with the next one: those go to
// the accumulator as 128-bit reductions (the buffer was zeroed one launch earlier); every other block belongs to this
// tile alone and is written with a plain 128-bit store -- no read-modify-write in L2 and nothing to zero.
template <int R>
__device__ __forceinline__ void ola_blocks(const float* s_slots, int nvalid, int hop_shift, bool first, bool last,
                                           float* dst) {
  const int n_blocks = nvalid - 1 + R;
  const int q_shift = hop_shift - 2;  // float4 items per hop block = 2^q_shift (hop is a power of two >= 128 here)
  const int items = n_blocks << q_shift;
  for (int it = threadIdx.x; it < items; it += blockDim.x) {
    const int b = it >> q_shift, j = (it & ((1 << q_shift) - 1)) << 2;
    const float* sp = s_slots + j;
    float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = R - 1; r >= 0; --r) {  // increasing frame index f = b - r
      const int f = b - r;
      if (f >= 0 && f < nvalid) {
        const float4 v = *reinterpret_cast<const float4*>(sp + f * kSlot + (r << hop_shift));
        const float2 lo = pk_add(make_float2(sum.x, sum.y), make_float2(v.x, v.y));
        const float2 hi = pk_add(make_float2(sum.z, sum.w), make_float2(v.z, v.w));
        sum = make_float4(lo.x, lo.y, hi.x, hi.y);
      }
    }
    float* p = dst + (b << hop_shift) + j;
    const bool shared = (!first && b < R - 1) || (!last && b >= nvalid);
    if (shared) red_add_v4(p, sum);
    else *reinterpret_cast<float4*>(p) = sum;
  }
}

// Sum the tile's frame slots in frame order and add the span to the global accumulator; zero, in the accumulator of
// the NEXT launch, the region this tile shares with its successor.
__device__ __forceinline__ void overlap_add_tile(const float* s_slots, const ClipDesc& cd, int t0, int hop, int hop_shift,
                                                 float* __restrict__ acc_out, float* __restrict__ acc_zero) {
  const int nvalid = min(kWarpsPerCta, cd.frames - t0);
  const bool first = t0 == 0, last = t0 + kWarpsPerCta >= cd.frames;
  float* dst = acc_out + cd.acc_offset + (int64_t)t0 * hop;
  const bool blocks = hop_shift >= 7 && hop_shift <= 10;
  if (hop_shift == 9) {
    ola_blocks<4>(s_slots, nvalid, hop_shift, first, last, dst);
  } else if (hop_shift == 8) {
    ola_blocks<8>(s_slots, nvalid, hop_shift, first, last, dst);
  } else if (hop_shift == 10) {
    ola_blocks<2>(s_slots, nvalid, hop_shift, first, last, dst);
  } else if (hop_shift == 7) {
    ola_blocks<16>(s_slots, nvalid, hop_shift, first, last, dst);
  } else {
    const int span = (nvalid - 1) * hop + kNfft;
    for (int p = threadIdx.x; p < span; p += blockDim.x) {
      int f_lo = p - (kNfft - 1);
      f_lo = f_lo > 0 ? (f_lo + hop - 1) / hop : 0;
      const int f_hi = min(nvalid - 1, p / hop);
      float sum = 0.0f;
      for (int f = f_lo; f <= f_hi; ++f) sum += s_slots[f * kSlot + p - f * hop];
      atomicAdd(dst + p, sum);
    }
  }
  if (acc_zero) {
    float* z = acc_zero + cd.acc_offset;
    if (blocks) {
      // only the region shared with the next tile is accumulated by two tiles: [(t0+nv)*hop, (t0+nv)*hop + n_fft - hop)
      if (!last) {
        const int64_t z0 = (int64_t)(t0 + nvalid) * hop, z1 = z0 + kNfft - hop;
        float4* z4 = reinterpret_cast<float4*>(z);
        for (int64_t p = (z0 >> 2) + threadIdx.x; p < (z1 >> 2); p += blockDim.x) z4[p] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else {
      const int64_t acc_len = kNfft + (int64_t)hop * (cd.frames - 1);
      const int64_t z0 = (int64_t)t0 * hop;
      const int64_t z1 = last ? acc_len : z0 + (int64_t)kWarpsPerCta * hop;
      for (int64_t p = z0 + threadIdx.x; p < z1; p += blockDim.x) z[p] = 0.0f;
    }
  }
}

constexpr size_t kGlSmemBytes = 8192 + sizeof(float2) * kTwpCount + 8192 + 8192 + sizeof(float2) * kScratchPerWarp * kWarpsPerCta;
constexpr int kSpecStride = 1032;  // float2 elements per tprev row: 8256 B, 16-byte aligned rows for cp.async

__device__ __forceinline__ float unit_scale(float ax, float ay) {
  // librosa: angles /= |angles| + 1e-16.  1 / (|a| + 1e-16) and rsqrt(|a|^2 + 1e-32) agree to float32 precision for
  // every |a| that is not itself ~1e-16 (where S * angles is inaudible either way), and both map a = 0 to 0.
  // rsqrt.approx (one MUFU.RSQ, 2 ulp): the argument is >= 1e-32, a normal number, so the denormal pre-/post-scaling that
  // rsqrtf() wraps around the same instruction (3-4 extra instructions per bin) has nothing to do here.
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fmaf(ax, ax, fmaf(ay, ay, 1e-32f))));
  return r;
}

// streaming read that does not displace the overlap-add accumulator lines from L1
__device__ __forceinline__ float ld_stream(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

// Shared-memory views of one CTA (tables staged once, one scratch tile per warp).
struct GlSmem {
  float2* tw1024;
  float2* twp;
  float* wq;
  float* wsyn;
  float2* scratch_all;
};

__device__ __forceinline__ GlSmem gl_smem_setup(unsigned char* smem_raw, const GlParams& P) {
  GlSmem m;
  m.tw1024 = reinterpret_cast<float2*>(smem_raw);
  m.twp = m.tw1024 + 1024;
  m.wq = reinterpret_cast<float*>(m.twp + kTwpCount);
  m.wsyn = m.wq + kNfft;
  m.scratch_all = reinterpret_cast<float2*>(m.wsyn + kNfft);
  stage_table(m.tw1024, P.tabs.tw1024, 512);
  stage_table(m.twp, P.tabs.twp, kTwpCount / 2);
  stage_table(m.wq, P.wq, 512);
  stage_table(m.wsyn, P.tabs.wsyn, 512);
  __syncthreads();
  return m;
}

// One tile (8 consecutive frames of clip c, one warp each) of one launch: synthesis only (INIT) or re-analysis ->
// re-projection -> synthesis, then the tile's overlap-add.  COHERENT: accumulator reads bypass L1 (needed when several
// iterations run inside one persistent kernel and another SM wrote the accumulator since this SM last read it).
template <bool INIT, bool FIRST, bool COHERENT>
__device__ __forceinline__ void gl_tile(const GlParams& P, const GlSmem& m, int c, const ClipDesc& cd, int tile) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kb = mirror_base(lane);
  float2* scratch = m.scratch_all + warp * kScratchPerWarp;
  float2* s_tw1024 = m.tw1024;
  float2* s_twp = m.twp;
  float* s_wq = m.wq;
  float* s_wsyn = m.wsyn;
  float2* s_scratch_all = m.scratch_all;
  const int t0 = (tile - cd.tile_offset) * kWarpsPerCta;
  const int t = t0 + warp;
  const bool active = t < cd.frames;
  // Barrier B2 ("every warp is done reading the slots of the previous tile's overlap-add") is executed exactly once per
  // warp and tile, at the latest point before this warp writes its scratch tile again: interior frames first request
  // their 32 accumulator loads, so the wait overlaps the load latency.
#ifdef MST_GL_B2_EARLY
  __syncthreads();
  constexpr bool kLateB2 = false;
#else
  constexpr bool kLateB2 = true;
#endif
    if (!active && kLateB2) __syncthreads();
    if (active) {
      const int64_t frame = cd.frame_offset + t;
      const float* Srow = P.S + frame * kBins;
      float2* trow = P.tprev + frame * kSpecStride;
      // Start the HBM -> L2 fetch of this frame's target magnitudes now; they are consumed after the forward FFT.
      for (int i = lane * 128; i < kBins * 4; i += 32 * 128) prefetch_l2(reinterpret_cast<const char*>(Srow) + i);
      float2 y[32];
      float2 mid = make_float2(0.0f, 0.0f);
      if (INIT) {
        if (kLateB2) __syncthreads();
        // y_0 = istft(S * exp(2*pi*i*u)).  All target magnitudes are requested first (33 loads in flight per lane), the
        // phases are computed while they arrive: the one-load-one-use form left a DRAM round trip exposed per bin
        // (ncu: long_scoreboard 5.7 cycles per issued instruction in this launch).
        float s_in[33];
#pragma unroll
        for (int j = 0; j < 33; ++j) {
          if (j == 32 && lane != 0) break;
          s_in[j] = ld_stream(Srow + (j < 32 ? mirror_bin(lane, kb, j) : 512));
        }
#pragma unroll
        for (int j = 0; j < 33; ++j) {
          if (j == 32 && lane != 0) break;
          const int k = j < 32 ? mirror_bin(lane, kb, j) : 512;
          float sn = 0.0f, cs = 1.0f;
          if (P.init_mode == 0) {
            float u;
            if (P.init_phase) {
              const int64_t idx = P.phase_layout == MST_LAYOUT_FRAME_MAJOR
                                      ? frame * kBins + k
                                      : cd.frame_offset * kBins + (int64_t)k * cd.frames + t;
              u = __ldg(P.init_phase + idx);
            } else {
              u = uniform_hash(P.seed, (unsigned long long)(frame * kBins + k));
            }
            // exp(2*pi*i*u): reduce to [-1/2, 1/2) turns, then the hardware sin/cos (absolute error ~5e-7)
            const float turns = u - rintf(u);
            __sincosf(6.283185307179586f * turns, &sn, &cs);
          }
          const float2 val = pk_mul(pk_bcast(s_in[j]), make_float2(cs, sn));
          if (j < 32) y[j] = val; else mid = val;
        }
      } else {
        // ---- re-analysis of y_{j-1}: reflect-padded frame, normalised by the window-sum-square envelope ----
        float2 v[32];
        const int64_t L = cd.length;                         // hop * (T - 1)
        const float* acc = P.acc_in + cd.acc_offset;         // acc[n + 1024] holds sample n before normalisation
        const int64_t base = (int64_t)t * P.hop - kHalf;     // first sample index of the frame (may be negative)
        const bool full = t >= P.q_full && t <= cd.frames - 1 - P.q_full && base >= 0 && base + kNfft <= L &&
                          ((base & 1) == 0);
        if (full) {
          // interior frame: the envelope is periodic in hop and pre-multiplied into the analysis window
          const float2* a2 = reinterpret_cast<const float2*>(acc + base + kHalf) + lane;
          const float2* w2 = reinterpret_cast<const float2*>(s_wq) + lane;
          if (kLateB2) {
#pragma unroll
            for (int r = 0; r < 32; ++r) v[r] = COHERENT ? __ldcg(a2 + 32 * r) : a2[32 * r];
            __syncthreads();
#pragma unroll
            for (int r = 0; r < 32; ++r) v[r] = pk_mul(v[r], w2[32 * r]);
          } else {
#pragma unroll
            for (int r = 0; r < 32; ++r) {
              const float2 a = COHERENT ? __ldcg(a2 + 32 * r) : a2[32 * r];
              v[r] = pk_mul(a, w2[32 * r]);
            }
          }
        } else {
          if (kLateB2) __syncthreads();
          // edge frame (rare): stage through this warp's scratch with a compact loop
          const float* iw = P.inv_wss + __ldg(P.wss_off + c);
          float* sf = reinterpret_cast<float*>(scratch);
#pragma unroll 1
          for (int jj = lane; jj < kNfft; jj += 32) {
            int64_t n = base + jj;
            bool ok = true;
            if (n < 0) {
              if (P.pad_mode == MST_PAD_REFLECT) n = -n; else ok = false;
            } else if (n >= L) {
              if (P.pad_mode == MST_PAD_REFLECT) n = 2 * (L - 1) - n; else ok = false;
            }
            sf[jj] = ok ? (COHERENT ? __ldcg(acc + n + kHalf) : acc[n + kHalf]) * __ldg(iw + n + kHalf) * __ldg(P.tabs.window + jj) : 0.0f;
          }
          __syncwarp();
#pragma unroll
          for (int r = 0; r < 32; ++r) v[r] = scratch[32 * r + lane];
          __syncwarp();
        }
        fft1024_front<-1>(v, scratch, s_tw1024, lane);
        if (!FIRST) {
          // previous iterate of this frame: HBM -> this warp's (now free) scratch tile, asynchronously, while the
          // second FFT pass and the split butterflies run
          const char* src = reinterpret_cast<const char*>(trow);
          char* dst = reinterpret_cast<char*>(scratch);
#pragma unroll
          for (int i = 0; i < 17; ++i) {
            const int o16 = (lane + 32 * i) * 16;
            if (o16 < kSpecStride * 8) cp_async16(dst + o16, src + o16);
          }
          cp_async_commit();
        }
        fft32<-1>(v);
        rfft_split(v, y, &mid, s_twp, lane);
        if (!FIRST) {
          cp_async_wait_all();
          __syncwarp();
        }
        // ---- re-projection: momentum update, unit-modulus phase, target magnitude ----
        const float* SA = Srow + lane;
        const float* SB = Srow + kb;
        float2* TA = trow + lane;
        float2* TB = trow + kb;
#pragma unroll
        for (int g8 = 0; g8 < 32; g8 += 8) {
          float sm[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) sm[i] = ld_stream((g8 < 16 ? SA : SB) + 32 * (g8 + i));
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int j = g8 + i;
            float2 tp = make_float2(0.0f, 0.0f);
            if (!FIRST) tp = scratch[mirror_bin(lane, kb, j)];
            if (!P.last_iter) (j < 16 ? TA : TB)[32 * j] = y[j];
            const float2 a = pk_fma(pk_bcast(-P.alpha), tp, y[j]);
            y[j] = pk_mul(pk_bcast(sm[i] * unit_scale(a.x, a.y)), a);
          }
        }
        if (lane == 0) {
          float2 tp = make_float2(0.0f, 0.0f);
          if (!FIRST) tp = scratch[512];
          if (!P.last_iter) trow[512] = mid;
          const float2 a = pk_fma(pk_bcast(-P.alpha), tp, mid);
          mid = pk_mul(pk_bcast(ld_stream(Srow + 512) * unit_scale(a.x, a.y)), a);
        }
        __syncwarp();  // everyone is done reading the staged previous iterate before the inverse FFT reuses the tile
      }
      synthesize_frame(y, mid, scratch, s_tw1024, s_twp, s_wsyn, lane);
    }
    __syncthreads();
    overlap_add_tile(reinterpret_cast<const float*>(s_scratch_all), cd, t0, P.hop, P.hop_shift, P.acc_out, P.acc_zero);
}

template <bool INIT, bool FIRST>
__global__ void __launch_bounds__(kWarpsPerCta * 32, kCtasPerSm) gl_kernel(GlParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const GlSmem m = gl_smem_setup(smem_raw, P);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  int c = blockIdx.x < P.total_tiles ? __ldg(P.tile_clip + blockIdx.x) : 0;
  ClipDesc cd = P.clips[c];
  for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
    // descriptor of this CTA's next tile: fetched now (one tile ahead), used at the bottom of the loop
    const int nt = tile + gridDim.x;
    const int c_next = nt < P.total_tiles ? __ldg(P.tile_clip + nt) : c;
    const ClipDesc cd_next = P.clips[c_next];
    if (!INIT && warp == 0 && nt < P.total_tiles) {  // pull the accumulator span of the next tile towards L2
      const float* a = P.acc_in + cd_next.acc_offset + (int64_t)(nt - cd_next.tile_offset) * kWarpsPerCta * P.hop;
      const int span = (kWarpsPerCta - 1) * P.hop + kNfft;
      for (int i = lane * 32; i < span; i += 32 * 32) prefetch_l2(a + i);
    }
    gl_tile<INIT, FIRST, false>(P, m, c, cd, tile);
    c = c_next;
    cd = cd_next;
  }
}

// Small batches (every tile resident at once, e.g. ONE 30 s clip = 162 tiles): all launches of the loop collapse into
// one cooperative persistent kernel -- tables are staged once, each CTA keeps its tile, iterations are separated by a
// grid-wide barrier instead of a launch, and the final normalisation runs in the same kernel.
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int& target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += gridDim.x;
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned int seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
    } while (seen < target);
    __threadfence();
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kWarpsPerCta * 32, kCtasPerSm)
gl_persistent_kernel(GlParams P, int n_iter, float* acc0, float* acc1, float* acc2, unsigned int* barrier, float* y_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const GlSmem m = gl_smem_setup(smem_raw, P);
  const int tile = blockIdx.x;  // grid == total_tiles, all co-resident (cooperative launch)
  const int c = __ldg(P.tile_clip + tile);
  const ClipDesc cd = P.clips[c];
  float* acc[3] = {acc0, acc1, acc2};
  unsigned int target = 0;

  P.acc_in = nullptr; P.acc_out = acc[0]; P.acc_zero = nullptr;
  gl_tile<true, false, true>(P, m, c, cd, tile);
  grid_barrier(barrier, target);
  int cur = 0;
  for (int j = 1; j <= n_iter; ++j) {
    P.acc_in = acc[cur];
    P.acc_out = acc[(cur + 1) % 3];
    P.acc_zero = j == n_iter ? nullptr : acc[(cur + 2) % 3];
    P.last_iter = (j == n_iter);
    gl_tile<false, false, true>(P, m, c, cd, tile);  // tprev starts zeroed, so iteration 1 needs no special case
    grid_barrier(barrier, target);
    cur = (cur + 1) % 3;
  }
  // y[n] = acc[n + 1024] / wss[n + 1024] for this tile's share of the clip
  {
    const int t0 = (tile - cd.tile_offset) * kWarpsPerCta;
    const bool last = t0 + kWarpsPerCta >= cd.frames;
    const int64_t n0 = (int64_t)t0 * P.hop;
    const int64_t n1 = last ? cd.length : min((int64_t)(t0 + kWarpsPerCta) * P.hop, cd.length);
    const float* a = acc[cur] + cd.acc_offset + kHalf;
    const float* w = P.inv_wss + __ldg(P.wss_off + c) + kHalf;
    float* dst = y_out + cd.sample_offset;
    for (int64_t n = n0 + threadIdx.x; n < n1; n += blockDim.x) dst[n] = __ldcg(a + n) * __ldg(w + n);
  }
}

// y[n] = acc[n + 1024] / wss[n + 1024]  (librosa.istft normalisation + centre trim)
__global__ void gl_finalize_kernel(const ClipDesc* __restrict__ clips, int n_clips, const float* __restrict__ acc,
                                   const float* __restrict__ inv_wss, const int64_t* __restrict__ wss_off,
                                   float* __restrict__ y) {
  const int c = blockIdx.y;
  const ClipDesc cd = clips[c];
  const float* a = acc + cd.acc_offset + kHalf;
  const float* w = inv_wss + wss_off[c] + kHalf;
  float* dst = y + cd.sample_offset;
  for (int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; n < cd.length; n += (int64_t)gridDim.x * blockDim.x)
    dst[n] = a[n] * __ldg(w + n);
}

// S ingest: any layout / log1p-power -> frame-major magnitudes.  One 32x32 tile per CTA (tiled transpose).
__global__ void gl_ingest_kernel(const float* __restrict__ S_in, int layout, int is_log1p_power,
                                 const ClipDesc* __restrict__ clips, float* __restrict__ S_out) {
  __shared__ float tile[32][33];
  const int c = blockIdx.z;
  const ClipDesc cd = clips[c];
  const int T = cd.frames;
  const float* src = S_in + cd.frame_offset * kBins;
  float* dst = S_out + cd.frame_offset * kBins;
  for (int tb = blockIdx.x * 32; tb < T; tb += gridDim.x * 32) {
    const int kb = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
      float val = 0.0f;
      if (layout == MST_LAYOUT_BIN_MAJOR) {
        const int k = kb + i, t = tb + threadIdx.x;  // coalesced along t
        if (k < kBins && t < T) val = src[(int64_t)k * T + t];
      } else {
        const int t = tb + i, k = kb + threadIdx.x;  // coalesced along k
        if (k < kBins && t < T) val = src[(int64_t)t * kBins + k];
      }
      if (is_log1p_power) val = sqrtf(expm1f(fminf(fmaxf(val, 0.0f), 20.0f)));  // inference.py:109
      tile[i][threadIdx.x] = val;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
      const int t = tb + i, k = kb + threadIdx.x;
      if (k < kBins && t < T) dst[(int64_t)t * kBins + k] = layout == MST_LAYOUT_BIN_MAJOR ? tile[threadIdx.x][i] : tile[i][threadIdx.x];
    }
    __syncthreads();
  }
}

}  // namespace mst

using namespace mst;

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

extern "C" {

size_t mst_griffinlim_workspace_bytes_ex(const mst_batch_t* b, int s_layout, int s_is_log1p_power) {
  if (!b) return 0;
  if (b->n_fft != kNfft) return generic_gl_workspace_bytes(b);
  const size_t spec = (size_t)b->total_frames * kBins;
  const bool needs_copy = s_layout != MST_LAYOUT_FRAME_MAJOR || s_is_log1p_power;  // else the caller's S is used in place
  return align_up((size_t)b->total_frames * kSpecStride * sizeof(float2), 256) +
         (needs_copy ? align_up(spec * sizeof(float), 256) : 0) +
         3 * align_up((size_t)b->total_acc * sizeof(float), 256) + 256;
}

size_t mst_griffinlim_workspace_bytes(const mst_batch_t* b) {
  return mst_griffinlim_workspace_bytes_ex(b, MST_LAYOUT_BIN_MAJOR, 1);  // worst case: transposed magnitude copy included
}

int mst_griffinlim_f32(const float* d_S, int s_layout, int s_is_log1p_power, const mst_batch_t* b, int n_iter,
                       float momentum, const float* d_init_phase, int init_mode, uint64_t seed, float* d_y_out,
                       void* d_workspace, size_t workspace_bytes, mst_stream_t stream) {
  if (!d_S || !b || !d_y_out || !d_workspace) return fail(MST_ERR_INVALID, "mst_griffinlim_f32: null argument");
  if (!b->from_frames || !b->d_inv_wss)
    return fail(MST_ERR_INVALID, "mst_griffinlim_f32 needs a batch made by mst_batch_create_from_frames");
  if (n_iter < 0) return fail(MST_ERR_INVALID, "n_iter must be >= 0");
  if (s_layout != MST_LAYOUT_FRAME_MAJOR && s_layout != MST_LAYOUT_BIN_MAJOR) return fail(MST_ERR_INVALID, "bad layout");
  if (init_mode != 0 && init_mode != 1) return fail(MST_ERR_INVALID, "bad init_mode");
  const size_t need = mst_griffinlim_workspace_bytes_ex(b, s_layout, s_is_log1p_power);
  if (workspace_bytes < need) return fail(MST_ERR_WORKSPACE, "workspace too small: %zu < %zu", workspace_bytes, need);
  if (reinterpret_cast<uintptr_t>(d_workspace) & 255) return fail(MST_ERR_INVALID, "workspace must be 256-byte aligned");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (b->n_fft != kNfft)  // general path (generic_fft.cu)
    return generic_griffinlim(d_S, s_layout, s_is_log1p_power, b, n_iter, momentum, d_init_phase, init_mode, seed, d_y_out,
                              d_workspace, s);
  Tables tabs;
  int rc = get_tables(&tabs);
  if (rc) return rc;
  if (b->d_window) { tabs.window = b->d_window; tabs.wsyn = b->d_wsyn; }  // win_length < n_fft

  const size_t spec = (size_t)b->total_frames * kBins;
  char* ws = reinterpret_cast<char*>(d_workspace);
  float2* tprev = reinterpret_cast<float2*>(ws); ws += align_up((size_t)b->total_frames * kSpecStride * sizeof(float2), 256);
  const bool needs_copy = s_layout != MST_LAYOUT_FRAME_MAJOR || s_is_log1p_power;
  float* S_t = reinterpret_cast<float*>(ws);
  if (needs_copy) ws += align_up(spec * sizeof(float), 256);
  const size_t acc_bytes = align_up((size_t)b->total_acc * sizeof(float), 256);
  float* acc[3];
  for (int i = 0; i < 3; ++i) { acc[i] = reinterpret_cast<float*>(ws); ws += acc_bytes; }

  const float* S_use = d_S;
  if (s_layout != MST_LAYOUT_FRAME_MAJOR || s_is_log1p_power) {
    int max_T = 0;
    for (int c = 0; c < b->n_clips; ++c) max_T = std::max(max_T, (int)b->h_clips[c].frames);
    // gridDim.z is limited to 65535 clips per launch
    for (int c0 = 0; c0 < b->n_clips; c0 += 65535) {
      const int nc = std::min(65535, b->n_clips - c0);
      dim3 grid((unsigned)std::min(64, (max_T + 31) / 32), (kBins + 31) / 32, (unsigned)nc);
      gl_ingest_kernel<<<grid, dim3(32, 8), 0, s>>>(d_S, s_layout, s_is_log1p_power, b->d_clips + c0, S_t);
      MST_CUDA_OK(cudaGetLastError());
      count_launch();
    }
    S_use = S_t;
  }

  int dev = 0, sms = 0;
  MST_CUDA_OK(cudaGetDevice(&dev));
  MST_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const size_t smem = kGlSmemBytes;
  if (dev < 0 || dev >= 64) return fail(MST_ERR_INVALID, "device index %d out of range", dev);
  static std::atomic<bool> attr_set[64];  // setting the attribute twice is harmless; the flag itself must not race
  if (!attr_set[dev].load(std::memory_order_acquire)) {
    MST_CUDA_OK(cudaFuncSetAttribute(gl_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MST_CUDA_OK(cudaFuncSetAttribute(gl_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MST_CUDA_OK(cudaFuncSetAttribute(gl_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set[dev].store(true, std::memory_order_release);
  }
  const int grid = std::min(b->total_tiles, kCtasPerSm * sms);

  MST_CUDA_OK(cudaMemsetAsync(acc[0], 0, acc_bytes, s));
  MST_CUDA_OK(cudaMemsetAsync(acc[1], 0, acc_bytes, s));

  GlParams P{};
  P.clips = b->d_clips; P.tile_clip = b->d_tile_clip; P.total_tiles = b->total_tiles;
  P.hop = b->hop; P.pad_mode = b->pad_mode; P.tabs = tabs;
  P.hop_shift = -1;
  for (int sft = 0; sft < 12; ++sft) if ((1 << sft) == b->hop) P.hop_shift = sft;
  P.q_full = (kNfft - 1) / b->hop;
  P.wq = b->d_wq;
  P.S = S_use; P.tprev = tprev; P.inv_wss = b->d_inv_wss; P.wss_off = b->d_wss_offset;
  P.alpha = momentum / (1.0f + momentum);
  P.init_phase = d_init_phase; P.phase_layout = s_layout; P.init_mode = init_mode; P.seed = seed;

  // Small batch: every tile can be resident at once -> one cooperative persistent kernel for the whole loop.
  {
    static std::atomic<int> coop_ok[64];  // 0 unknown, 1 usable, -1 not
    if (coop_ok[dev].load(std::memory_order_acquire) == 0) {
      int coop = 0, per_sm = 0;
      cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
      if (coop && cudaFuncSetAttribute(gl_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess &&
          cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gl_persistent_kernel, kWarpsPerCta * 32, smem) == cudaSuccess &&
          per_sm >= kCtasPerSm)
        coop_ok[dev].store(1, std::memory_order_release);
      else
        coop_ok[dev].store(-1, std::memory_order_release);
      cudaGetLastError();
    }
    if (coop_ok[dev].load(std::memory_order_acquire) == 1 && b->total_tiles <= kCtasPerSm * sms) {
      unsigned int* barrier = reinterpret_cast<unsigned int*>(ws);  // the 256 spare bytes at the end of the workspace
      MST_CUDA_OK(cudaMemsetAsync(barrier, 0, 256, s));
      MST_CUDA_OK(cudaMemsetAsync(tprev, 0, (size_t)b->total_frames * kSpecStride * sizeof(float2), s));
      P.acc_in = nullptr; P.acc_out = acc[0]; P.acc_zero = nullptr; P.first_iter = 0;
      int n_it = n_iter;
      void* args[] = {&P, &n_it, &acc[0], &acc[1], &acc[2], &barrier, &d_y_out};
      MST_CUDA_OK(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(gl_persistent_kernel), dim3(b->total_tiles),
                                              dim3(kWarpsPerCta * 32), args, smem, s));
      count_launch();
      return MST_OK;
    }
  }

  // launch 0: y_0 from the initial phase, accumulated into acc[0]
  P.acc_in = nullptr; P.acc_out = acc[0]; P.acc_zero = nullptr; P.first_iter = 1;
  gl_kernel<true, false><<<grid, kWarpsPerCta * 32, smem, s>>>(P);
  MST_CUDA_OK(cudaGetLastError());
  count_launch();
  int cur = 0;
  for (int j = 1; j <= n_iter; ++j) {
    P.acc_in = acc[cur];
    P.acc_out = acc[(cur + 1) % 3];
    P.acc_zero = j == n_iter ? nullptr : acc[(cur + 2) % 3];  // the launch after the last one accumulates nothing
    P.first_iter = (j == 1);
    P.last_iter = (j == n_iter);
    if (j == 1) gl_kernel<false, true><<<grid, kWarpsPerCta * 32, smem, s>>>(P);
    else gl_kernel<false, false><<<grid, kWarpsPerCta * 32, smem, s>>>(P);
    MST_CUDA_OK(cudaGetLastError());
    count_launch();
    cur = (cur + 1) % 3;
  }
  {
    int64_t max_len = 0;
    for (int c = 0; c < b->n_clips; ++c) max_len = std::max(max_len, b->h_clips[c].length);
    for (int c0 = 0; c0 < b->n_clips; c0 += 65535) {
      const int nc = std::min(65535, b->n_clips - c0);
      dim3 fgrid((unsigned)std::max<int64_t>(1, std::min<int64_t>(64, (max_len + 1023) / 1024)), (unsigned)nc);
      gl_finalize_kernel<<<fgrid, 256, 0, s>>>(b->d_clips + c0, nc, acc[cur], b->d_inv_wss, b->d_wss_offset + c0, d_y_out);
      MST_CUDA_OK(cudaGetLastError());
      count_launch();
    }
  }
  return MST_OK;
}

}  // extern "C"
