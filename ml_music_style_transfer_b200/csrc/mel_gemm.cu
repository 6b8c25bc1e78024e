// P2: mel filterbank projection on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM, operands
// staged in shared memory by TMA).  Replaces `np.dot(mel_basis, |stft|**2)` inside librosa.feature.melspectrogram
// (reference tests/plot_spec.py:20; preprocessing/preprocess.py:55).
//
// mel[frame][m] = sum_k W[m][k] * P[frame][k] is a dense contraction, but W (128 x 1025, Slaney triangles) is 98.5 %
// zeros and *banded*: a 64-bin slice of the spectrum only touches 16..48 consecutive mel rows.  The kernel therefore
// runs one small UMMA per K-slice against the band of W that is non-zero there:
//
//     D[128 frames x N_s mels]  +=  A[128 frames x 64 bins] . B_s[N_s mels x 64 bins]^T        (M=128, N=N_s, K=4x16)
//
// with D a column window [n0_s, n0_s + N_s) of one 128-lane TMEM accumulator.  The banded filterbank (~92 KB in
// split-bf16 form) is resident in shared memory for the life of the persistent CTA, so the only stream is the power
// spectrum itself, which the STFT kernel leaves in an L2-sized ring as split bf16 (hi + lo).
//
// Precision: float32 parity (<= 1e-4 relative) is kept with an error-compensated split, three passes per slice:
//     P*W ~= P_hi*W_hi + P_lo*W_hi + P_hi*W_lo          (x_hi = bf16(x), x_lo = bf16(x - x_hi); fp32 accumulation in TMEM)
// whose dropped term P_lo*W_lo is ~2^-18 relative.  All operands are non-negative, so there is no cancellation.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2..5 = epilogue (tcgen05.ld -> log1p -> global).  Two TMEM accumulators alternate between tiles so the
// epilogue of tile i overlaps the MMAs of tile i+1.
#include <cuda.h>
#include <cuda_bf16.h>
#include <algorithm>
#include <vector>
#include "fft_warp.cuh"
#include "mel_plan.cuh"
#include "mst_common.cuh"

namespace mst {

constexpr int kGemmThreads = 192;
constexpr int kTileFrames = 128;                    // UMMA M
constexpr int kStages = 3;                          // A-operand ring depth
constexpr int kStageBytes = 2 * kTileFrames * 128;  // P_hi + P_lo tile of one K-slice: 2 x 16 KB

// ---- raw PTX wrappers ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int x, int y) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y)
      : "memory");
}
// 1-D bulk asynchronous copy global -> shared (TMA engine, no tensor map), completion on an mbarrier.
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// K-major, SWIZZLE_128B operand descriptor: rows of 128 bytes, 8-row groups 1024 bytes apart (SBO), version 1.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);  // start address
  d |= (uint64_t)(1024 >> 4) << 32;             // stride byte offset between 8-row core-matrix groups
  d |= (uint64_t)1 << 46;                       // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: D = F32, A = B = BF16, both K-major, M = 128, N = n.
__device__ __forceinline__ uint32_t umma_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTileFrames >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, "
      "%24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_zero32(uint32_t taddr) {
  const uint32_t z = 0;
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, "
      "%1, %1, %1, %1};" ::"r"(taddr),
      "r"(z)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

struct GemmParams {
  MelSlices slices;              // banded filterbank geometry
  const __nv_bfloat16* w_hi;     // [sum N_s][64] banded filterbank, high part, ALREADY in the swizzled shared-memory image
  const __nv_bfloat16* w_lo;     // same, low part
  int w_rows;                    // sum N_s
  int n_mels, n_cols;            // mel rows, TMEM columns per accumulator (n_mels rounded up to 32)
  int n_rows;                    // valid frames in this chunk of the ring
  int64_t g0;                    // global frame id of ring row 0
  const ClipDesc* clips;
  int n_clips;
  int uniform_frames;            // frames per clip when the batch is uniform, else 0
  int apply_log1p, layout;
  float* out;
};

__global__ void __launch_bounds__(kGemmThreads, 1)
mel_gemm_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo, GemmParams P) {
  extern __shared__ unsigned char smem_raw[];
  // SWIZZLE_128B operand tiles need 1024-byte alignment in the shared window
  unsigned char* smem = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  // [ A ring: kStages x (hi 16 KB | lo 16 KB) ][ W_hi band | W_lo band ][ barriers ]
  unsigned char* s_a = smem;
  unsigned char* s_whi = smem + kStages * kStageBytes;
  unsigned char* s_wlo = s_whi + (size_t)P.w_rows * 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_wlo + (size_t)P.w_rows * 128);
  uint64_t* full = bars;                      // [kStages] TMA -> MMA
  uint64_t* empty = bars + kStages;           // [kStages] MMA -> TMA
  uint64_t* tmem_full = bars + 2 * kStages;   // [2] MMA -> epilogue
  uint64_t* tmem_empty = tmem_full + 2;       // [2] epilogue -> MMA
  uint64_t* wbar = tmem_empty + 2;            // [1] filterbank image landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wbar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (P.n_rows + kTileFrames - 1) / kTileFrames;
  const uint32_t tmem_cols = 2 * P.n_cols;  // 256 or 512: power of two >= 32

  // ---- one-time setup -------------------------------------------------------------------------------------------
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(tmem_full + i, 1); mbar_init(tmem_empty + i, 4); }
    mbar_init(wbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // banded filterbank: the plan stores it as the exact shared-memory image (UMMA K-major SWIZZLE_128B: 16-byte
    // chunk c of row r at c ^ (r & 7)), so two bulk copies bring in all ~92 KB without touching registers
    const uint32_t wbytes = (uint32_t)P.w_rows * 128;
    mbar_arrive_expect_tx(wbar, 2 * wbytes);
    bulk_load(s_whi, P.w_hi, wbytes, wbar);
    bulk_load(s_wlo, P.w_lo, wbytes, wbar);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to UMMA
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer ==========================================================================================
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int s = 0; s < P.slices.n_slices; ++s) {
          if (P.slices.n[s] == 0) continue;
          const int stage = it % kStages;
          mbar_wait(empty + stage, ((it / kStages) & 1) ^ 1);
          mbar_arrive_expect_tx(full + stage, kStageBytes);
          unsigned char* dst = s_a + stage * kStageBytes;
          tma_load_2d(dst, &map_hi, full + stage, s * 64, tile * kTileFrames);
          tma_load_2d(dst + kTileFrames * 128, &map_lo, full + stage, s * 64, tile * kTileFrames);
          ++it;
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) ===============================================================================
    if (lane == 0) {
      uint32_t it = 0, local_tile = 0;
      mbar_wait(wbar, 0);  // filterbank image resident
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++local_tile) {
        const int buf = local_tile & 1;
        mbar_wait(tmem_empty + buf, (local_tile >> 1) & 1);  // accumulator drained and re-zeroed by the epilogue
        tc_fence_after();
        for (int s = 0; s < P.slices.n_slices; ++s) {
          const int n = P.slices.n[s];
          if (n == 0) continue;
          const int stage = it % kStages;
          mbar_wait(full + stage, (it / kStages) & 1);
          tc_fence_after();
          const uint32_t a_hi = smem_u32(s_a + stage * kStageBytes), a_lo = a_hi + kTileFrames * 128;
          const uint32_t b_hi = smem_u32(s_whi) + P.slices.row[s] * 128, b_lo = smem_u32(s_wlo) + P.slices.row[s] * 128;
          const uint32_t d = tmem_base + buf * P.n_cols + P.slices.n0[s];
          const uint32_t idesc = umma_idesc(n);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {  // 4 x UMMA_K(16 bf16 = 32 bytes) per 64-bin slice
            const uint32_t ko = kk * 32;
            umma_bf16(d, umma_desc(a_hi + ko), umma_desc(b_hi + ko), idesc, 1);
            umma_bf16(d, umma_desc(a_lo + ko), umma_desc(b_hi + ko), idesc, 1);
            umma_bf16(d, umma_desc(a_hi + ko), umma_desc(b_lo + ko), idesc, 1);
          }
          umma_commit(empty + stage);  // frees the A stage when these MMAs have read it
          ++it;
        }
        umma_commit(tmem_full + buf);  // accumulator complete -> epilogue
      }
    }
    __syncwarp();
  } else {
    // ===== epilogue warps 2..5: TMEM lane quarter (warp & 3), one frame per thread ================================
    const int q = warp & 3;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    // zero both accumulators once, then hand them to the MMA issuer
    for (int buf = 0; buf < 2; ++buf)
      for (int c = 0; c < P.n_cols; c += 32) tmem_zero32(lane_base + buf * P.n_cols + c);
    tmem_wait_st();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) { mbar_arrive(tmem_empty + 0); mbar_arrive(tmem_empty + 1); }

    uint32_t local_tile = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++local_tile) {
      const int buf = local_tile & 1;
      const int row = tile * kTileFrames + q * 32 + lane;  // ring row of this thread's frame
      const bool valid = row < P.n_rows;
      const int64_t g = P.g0 + row;
      // locate the frame inside its clip (needed for the bin-major layout)
      int64_t out_base = 0, out_stride = 1;
      if (valid) {
        if (P.layout == MST_LAYOUT_FRAME_MAJOR) {
          out_base = g * P.n_mels;
          out_stride = 1;
        } else {
          int c;
          if (P.uniform_frames > 0) {
            c = (int)(g / P.uniform_frames);
          } else {
            int lo = 0, hi = P.n_clips;
            while (hi - lo > 1) {
              const int mid = (lo + hi) >> 1;
              if (P.clips[mid].frame_offset <= g) lo = mid; else hi = mid;
            }
            c = lo;
          }
          const ClipDesc cd = P.clips[c];
          out_base = cd.frame_offset * P.n_mels + (g - cd.frame_offset);
          out_stride = cd.frames;
        }
      }
      mbar_wait(tmem_full + buf, (local_tile >> 1) & 1);
      tc_fence_after();
      for (int c0 = 0; c0 < P.n_cols; c0 += 32) {
        uint32_t r[32];
        const uint32_t taddr = lane_base + buf * P.n_cols + c0;
        tmem_ld32(taddr, r);
        tmem_zero32(taddr);  // leave the columns zeroed for the tile after next
        if (valid) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int m = c0 + i;
            if (m < P.n_mels) {
              float v = __uint_as_float(r[i]);
              if (P.apply_log1p) v = fast_log1p(v);
              P.out[out_base + (int64_t)m * out_stride] = v;
            }
          }
        }
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty + buf);
    }
  }

  // ---- teardown ---------------------------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// ---- host side ------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int get_encode_fn(EncodeTiledFn* out) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    MST_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || !p) return fail(MST_ERR_CUDA, "cuTensorMapEncodeTiled not available");
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  *out = fn;
  return MST_OK;
}

static int make_ring_map(CUtensorMap* map, void* base, int64_t rows) {
  EncodeTiledFn enc;
  int rc = get_encode_fn(&enc);
  if (rc) return rc;
  const cuuint64_t dims[2] = {(cuuint64_t)kSpecPad, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)kSpecPad * sizeof(__nv_bfloat16)};
  const cuuint32_t box[2] = {64, (cuuint32_t)kTileFrames};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(MST_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return MST_OK;
}

size_t mel_gemm_smem_bytes(const mst_mel_plan* plan) {
  return (size_t)kStages * kStageBytes + 2 * (size_t)plan->w_rows * 128 + 16 * sizeof(uint64_t) + 1024;
}

int launch_mel_gemm(const mst_mel_plan* plan, const mst_batch* b, void* ring_hi, void* ring_lo, int64_t ring_rows,
                    int n_rows, int64_t g0, int apply_log1p, int layout, float* out, cudaStream_t stream) {
  CUtensorMap map_hi, map_lo;
  int rc = make_ring_map(&map_hi, ring_hi, ring_rows);
  if (rc) return rc;
  rc = make_ring_map(&map_lo, ring_lo, ring_rows);
  if (rc) return rc;
  int dev = 0, sms = 0;
  MST_CUDA_OK(cudaGetDevice(&dev));
  MST_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const size_t smem = mel_gemm_smem_bytes(plan);
  static size_t attr_smem[64] = {0};
  if (attr_smem[dev] < smem) {
    MST_CUDA_OK(cudaFuncSetAttribute(mel_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_smem[dev] = smem;
  }
  GemmParams P{};
  P.slices = plan->slices;
  P.w_hi = plan->d_band_hi;
  P.w_lo = plan->d_band_lo;
  P.w_rows = plan->w_rows;
  P.n_mels = plan->n_mels;
  P.n_cols = plan->n_mels <= 128 ? 128 : 256;
  P.n_rows = n_rows;
  P.g0 = g0;
  P.clips = b->d_clips;
  P.n_clips = b->n_clips;
  P.uniform_frames = b->uniform_frames;
  P.apply_log1p = apply_log1p;
  P.layout = layout;
  P.out = out;
  const int n_tiles = (n_rows + kTileFrames - 1) / kTileFrames;
  mel_gemm_kernel<<<std::min(n_tiles, sms), kGemmThreads, smem, stream>>>(map_hi, map_lo, P);
  MST_CUDA_OK(cudaGetLastError());
  count_launch();
  return MST_OK;
}

}  // namespace mst
