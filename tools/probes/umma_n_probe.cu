// Does a cta_group::2 tcgen05.mma (M = 256, kind::f16) cost time proportional to N?  Decides whether the transposed K1
// (N = 240 = 10 slots, no pad rows) can win over N = 256.  One CTA pair per 2 SMs, the leader issues a long
// chain of MMAs over zero-filled shared-memory operand tiles (SWIZZLE_128B, 3 x 4 k-slices per "k-block" as K1 does) into
// two alternating accumulators, commits every k-block, waits at the end.  Build + run:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I egohmr_b200/csrc -o /tmp/umma_n_probe tools/probes/umma_n_probe.cu && /tmp/umma_n_probe
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"

using namespace ehb;

template <int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) probe(int kblocks, int stages, int* sink) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = raw + ((1024u - (ptx::smem_u32(raw) & 1023u)) & 1023u);
  constexpr int STAGE = 2 * 16384 + 2 * 16384;   // A hi/lo 128 rows x 128 B, B hi/lo (N/2 <= 128 rows) x 128 B
  __shared__ uint64_t bar_done, bar_kb;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  const bool leader = ptx::cluster_ctarank() == 0;
  for (int i = threadIdx.x; i < stages * STAGE / 16; i += blockDim.x) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar_done, 1);
    ptx::mbar_init(&bar_kb, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc_2sm(&tmem_slot, 512);
    ptx::tmem_relinquish_2sm();
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::cluster_sync();
  ptx::tc_fence_after_sync();
  const uint32_t tmem = tmem_slot;
  if (warp == 1 && leader) {
    constexpr uint32_t idesc = ptx::make_idesc_f16_f32(256, N);
    int stage = 0;
    for (int kb = 0; kb < kblocks; ++kb) {
      const uint32_t sa = ptx::smem_u32(sm + stage * STAGE);
      const uint32_t tacc = tmem + ((kb >> 4) & 1) * 256;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint32_t koff = ks * 32;
        const uint64_t a_hi = ptx::make_kmajor_desc<128>(sa + koff), a_lo = ptx::make_kmajor_desc<128>(sa + 16384 + koff);
        const uint64_t b_hi = ptx::make_kmajor_desc<128>(sa + 32768 + koff), b_lo = ptx::make_kmajor_desc<128>(sa + 49152 + koff);
        ptx::umma_f16_2sm_elect(tacc, a_hi, b_hi, idesc, (kb & 15) | ks ? 1u : 0u);
        ptx::umma_f16_2sm_elect(tacc, a_hi, b_lo, idesc, 1u);
        ptx::umma_f16_2sm_elect(tacc, a_lo, b_hi, idesc, 1u);
      }
      ptx::umma_commit_2sm_mc_elect(&bar_kb, 0b01);   // as K1: one commit per k-block (nobody waits on it here)
      __syncwarp();
      if (++stage == stages) stage = 0;
    }
    ptx::umma_commit_2sm_mc_elect(&bar_done, 0b11);
    __syncwarp();
  }
  ptx::mbar_wait(&bar_done, 0);
  ptx::tc_fence_after_sync();
  if (threadIdx.x < 32) {
    float v[16];
    ptx::tmem_ld_32x32b_x16(tmem, v);
    ptx::tmem_ld_wait();
    if (v[0] != 0.f) atomicAdd(sink, 1);
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::cluster_sync();
  if (warp == 0) ptx::tmem_dealloc_2sm(tmem, 512);
}

template <int N>
static void run(int sms, int* sink) {
  const int kblocks = 4096, stages = 3;
  const size_t smem = 1024 + stages * 65536;
  cudaFuncSetAttribute(probe<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int grid = (sms / 2) * 2;
  probe<N><<<grid, 128, smem>>>(64, stages, sink);
  cudaDeviceSynchronize();
  float best = 1e9f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    probe<N><<<grid, 128, smem>>>(kblocks, stages, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    best = ms < best ? ms : best;
  }
  cudaError_t e = cudaGetLastError();
  const double mmas = 12.0 * kblocks;
  const double tflops = 2.0 * 256 * N * 16 * mmas * (grid / 2) / (best * 1e-3) / 1e12;
  printf("{\"N\": %d, \"ms\": %.4f, \"ns_per_mma\": %.2f, \"issued_tflops\": %.1f, \"err\": \"%s\"}\n", N, best,
         best * 1e6 / mmas, tflops, cudaGetErrorString(e));
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  int* sink;
  cudaMalloc(&sink, 4);
  cudaMemset(sink, 0, 4);
  run<256>(prop.multiProcessorCount, sink);
  run<240>(prop.multiProcessorCount, sink);
  run<224>(prop.multiProcessorCount, sink);
  run<192>(prop.multiProcessorCount, sink);
  run<128>(prop.multiProcessorCount, sink);
  run<96>(prop.multiProcessorCount, sink);
  run<64>(prop.multiProcessorCount, sink);
  run<32>(prop.multiProcessorCount, sink);
  run<256>(prop.multiProcessorCount, sink);
  return 0;
}
