// Microbenchmarks behind two design decisions (build: python -m ml_music_style_transfer_b200.build --tools;
// run on the GPU box: tools/ubench > gpurun_out/ubench.jsonl).
//
//  1. What does a WRITE-ONLY stream reach on this B200?  The piano-roll upsample (P3d) writes 254 GB per step and reads
//     next to nothing, so its ceiling is the write bandwidth, not the copy bandwidth of MEASURED_PEAKS.json.  Variants:
//     cudaMemsetAsync, a grid-stride st.global.v4 kernel (default / .cs / L1::no_allocate), and shared-memory-staged bulk
//     stores (cp.async.bulk.global.shared::cta, the TMA store path); plus read-only and copy for reference.
//  2. Does the packed fp32x2 FMA of sm_100 (PTX fma.rn.f32x2 -> SASS FFMA2) raise flops per ISSUE SLOT?  The FFT
//     butterflies are issue-bound; a complex butterfly is 3 FFMA2 instead of 6 FFMA.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int MODE>
__global__ void store_v4(uint4* __restrict__ p, size_t n_vec, unsigned v) {
  const uint4 val = make_uint4(v, v, v, v);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (size_t)gridDim.x * blockDim.x) {
    if (MODE == 0) p[i] = val;
    if (MODE == 1) asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p + i), "r"(val.x), "r"(val.y), "r"(val.z), "r"(val.w) : "memory");
    if (MODE == 2) asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p + i), "r"(val.x), "r"(val.y), "r"(val.z), "r"(val.w) : "memory");
  }
}

// each thread writes UNROLL consecutive-by-warp vectors per trip: more stores in flight per thread
template <int UNROLL>
__global__ void store_v4_unrolled(uint4* __restrict__ p, size_t n_vec, unsigned v) {
  const uint4 val = make_uint4(v, v, v, v);
  const size_t stride = (size_t)gridDim.x * blockDim.x * UNROLL;
  for (size_t base = (size_t)blockIdx.x * blockDim.x * UNROLL + threadIdx.x; base < n_vec; base += stride) {
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const size_t i = base + (size_t)u * blockDim.x;
      if (i < n_vec) p[i] = val;
    }
  }
}

// shared-memory staged bulk store: each CTA owns a TILE-byte shared buffer filled once; one thread streams it out
template <int TILE>
__global__ void store_bulk(unsigned char* __restrict__ p, size_t n_bytes, unsigned v) {
  extern __shared__ __align__(128) unsigned char tile[];
  for (int i = threadIdx.x; i < TILE / 16; i += blockDim.x) reinterpret_cast<uint4*>(tile)[i] = make_uint4(v, v, v, v);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(tile);
    const size_t n_tiles = n_bytes / TILE;
    int pending = 0;
    for (size_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(p + t * TILE), "r"(s), "r"(TILE) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      if (++pending >= 8) { asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory"); pending = 4; }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

__global__ void read_v4(const uint4* __restrict__ p, size_t n_vec, unsigned* out) {
  unsigned acc = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (size_t)gridDim.x * blockDim.x) {
    const uint4 v = p[i];
    acc ^= v.x ^ v.y ^ v.z ^ v.w;
  }
  if (acc == 0x12345678u) *out = acc;
}

__global__ void copy_v4(const uint4* __restrict__ a, uint4* __restrict__ b, size_t n_vec) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}

// ---- FMA issue-rate kernels: 16 independent accumulator chains per thread, 8 warps x 2 CTAs per SM ------------------
template <int MODE>
__global__ void __launch_bounds__(256) fma_rate(float* out, int iters, float a, float b) {
  float2 acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
  const float2 m = make_float2(a, a), c = make_float2(b, b);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) {           // scalar FFMA, register operands
        acc[i].x = fmaf(acc[i].x, m.x, c.x);
        acc[i].y = fmaf(acc[i].y, m.y, c.y);
      } else if (MODE == 1) {    // scalar FFMA, immediate multiplier
        acc[i].x = fmaf(acc[i].x, 0.999f, c.x);
        acc[i].y = fmaf(acc[i].y, 0.999f, c.y);
      } else if (MODE == 2) {    // packed FFMA2, register operands
        unsigned long long ra, rm, rc;
        asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(acc[i].x), "f"(acc[i].y));
        asm("mov.b64 %0, {%1, %2};" : "=l"(rm) : "f"(m.x), "f"(m.y));
        asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
        asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(ra) : "l"(ra), "l"(rm), "l"(rc));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[i].x), "=f"(acc[i].y) : "l"(ra));
      } else {                   // packed FFMA2, immediate multiplier
        unsigned long long ra, rm, rc;
        asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(acc[i].x), "f"(acc[i].y));
        asm("mov.b64 %0, {%1, %2};" : "=l"(rm) : "f"(0.999f), "f"(0.999f));
        asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
        asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(ra) : "l"(ra), "l"(rm), "l"(rc));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[i].x), "=f"(acc[i].y) : "l"(ra));
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i].x + acc[i].y;
  if (s == 123.456f) out[0] = s;
}

template <typename F>
static float time_ms(F f, int reps) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) f();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) f();
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  CK(cudaGetLastError());
  return ms / reps;
}

int main() {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const size_t bytes = (size_t)4 << 30;  // 4 GiB >> 126 MB L2
  unsigned char *a, *b;
  unsigned* flag;
  CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes)); CK(cudaMalloc(&flag, 4));
  CK(cudaMemset(a, 1, bytes)); CK(cudaMemset(b, 2, bytes));
  const size_t n_vec = bytes / 16;
  auto report = [&](const char* name, float ms, double moved) {
    printf("{\"bench\": \"%s\", \"ms\": %.4f, \"gbs\": %.1f}\n", name, ms, moved / (ms * 1e-3) / 1e9);
    fflush(stdout);
  };
  report("cudaMemsetAsync", time_ms([&] { CK(cudaMemsetAsync(a, 3, bytes)); }, 10), (double)bytes);
  for (int ctas_per_sm : {2, 4, 8, 16}) {
    char nm[96];
    snprintf(nm, sizeof nm, "st.global.v4 grid=%dxSM block=256", ctas_per_sm);
    report(nm, time_ms([&] { store_v4<0><<<sms * ctas_per_sm, 256>>>((uint4*)a, n_vec, 5u); }, 10), (double)bytes);
  }
  report("st.global.cs.v4 grid=8xSM", time_ms([&] { store_v4<1><<<sms * 8, 256>>>((uint4*)a, n_vec, 5u); }, 10), (double)bytes);
  report("st.global.L1::no_allocate.v4 grid=8xSM", time_ms([&] { store_v4<2><<<sms * 8, 256>>>((uint4*)a, n_vec, 5u); }, 10), (double)bytes);
  report("st.global.v4 unroll8 grid=4xSM", time_ms([&] { store_v4_unrolled<8><<<sms * 4, 256>>>((uint4*)a, n_vec, 5u); }, 10), (double)bytes);
  report("st.global.v4 one-shot grid (1 vec/thread)", time_ms([&] { store_v4<0><<<(unsigned)(n_vec / 256), 256>>>((uint4*)a, n_vec, 5u); }, 10), (double)bytes);
  CK(cudaFuncSetAttribute(store_bulk<32768>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
  CK(cudaFuncSetAttribute(store_bulk<65536>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  report("cp.async.bulk smem->global 32KB tiles grid=4xSM", time_ms([&] { store_bulk<32768><<<sms * 4, 128, 32768>>>(a, bytes, 7u); }, 10), (double)bytes);
  report("cp.async.bulk smem->global 64KB tiles grid=2xSM", time_ms([&] { store_bulk<65536><<<sms * 2, 128, 65536>>>(a, bytes, 7u); }, 10), (double)bytes);
  report("ld.global.v4 read-only grid=8xSM", time_ms([&] { read_v4<<<sms * 8, 256>>>((const uint4*)a, n_vec, flag); }, 10), (double)bytes);
  report("copy v4 (read+write bytes) grid=8xSM", time_ms([&] { copy_v4<<<sms * 8, 256>>>((const uint4*)a, (uint4*)b, n_vec); }, 10), 2.0 * bytes);
  report("cudaMemcpyAsync D2D (read+write bytes)", time_ms([&] { CK(cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice)); }, 10), 2.0 * bytes);

  // FMA issue rate
  float* out;
  CK(cudaMalloc(&out, 4));
  const int iters = 4096;
  int clk_khz = 0;
  CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  auto fma_report = [&](const char* name, float ms, double fma_per_thread) {
    const double threads = (double)sms * 2 * 256;
    const double tflops = 2.0 * fma_per_thread * threads / (ms * 1e-3) / 1e12;
    printf("{\"bench\": \"%s\", \"ms\": %.4f, \"tflops\": %.2f, \"fma_per_clk_per_sm_at_max_clock\": %.1f}\n", name, ms, tflops,
           fma_per_thread * threads / (ms * 1e-3) / sms / (clk_khz * 1e3));
    fflush(stdout);
  };
  const double fpt = (double)iters * 16 * 2;
  fma_report("FFMA reg x reg", time_ms([&] { fma_rate<0><<<sms * 2, 256>>>(out, iters, 0.999f, 0.001f); }, 5), fpt);
  fma_report("FFMA reg x imm", time_ms([&] { fma_rate<1><<<sms * 2, 256>>>(out, iters, 0.999f, 0.001f); }, 5), fpt);
  fma_report("FFMA2 reg x reg (packed f32x2)", time_ms([&] { fma_rate<2><<<sms * 2, 256>>>(out, iters, 0.999f, 0.001f); }, 5), fpt);
  fma_report("FFMA2 reg x imm (packed f32x2)", time_ms([&] { fma_rate<3><<<sms * 2, 256>>>(out, iters, 0.999f, 0.001f); }, 5), fpt);
  return 0;
}
