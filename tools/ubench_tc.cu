// Micro-benchmarks that size the K4 epilogue (development aid, not part of libvfr):
//   (a) tcgen05.ld throughput per SM for 4 / 8 / 16 reader warps and x16 / x32 / x64 shapes
//   (b) tcgen05.mma (SS, kind::f16, M=128) cycles per instruction for N = 48 ... 256
//   (c) both at once (does the TMEM read port slow the MMA?)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_tc ubench_tc.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>
#include <algorithm>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ uint64_t make_desc(const void* smem_ptr) {
  const uint32_t addr = smem_u32(smem_ptr);
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3fff);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int X>
__device__ __forceinline__ uint32_t tmem_ld(uint32_t taddr);
template <>
__device__ __forceinline__ uint32_t tmem_ld<16>(uint32_t taddr) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  uint32_t x = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) x ^= r[i];
  return x;
}
template <>
__device__ __forceinline__ uint32_t tmem_ld<32>(uint32_t taddr) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  uint32_t x = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) x ^= r[i];
  return x;
}
template <>
__device__ __forceinline__ uint32_t tmem_ld<64>(uint32_t taddr) {
  return tmem_ld<32>(taddr) ^ tmem_ld<32>(taddr + 32);   // two x32 loads in flight before the wait
}

// mode bit 0: readers active, bit 1: MMA active
template <int X>
__global__ void __launch_bounds__(64 + 512, 1)
ubench_kernel(int mode, int reader_warps, int ld_iters, int mma_n, int mma_count, int wait_every, long long* out_ld,
              long long* out_mma, uint32_t* sink, int chains = 2) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + (uint32_t)i * 0x00010001u;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    if (lane == 0 && (mode & 2)) {
      const uint32_t idesc = (1u << 4) | ((wait_every & 1) ? ((1u << 7) | (1u << 10)) : 0u) | ((uint32_t)(mma_n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);  // wait_every odd: bf16, even: fp16
      const uint64_t adesc = make_desc(smem), bdesc = make_desc(smem + 16384);
      const long long t0 = clock64();
      // `chains` accumulators in rotation: 1 = every MMA accumulates into the result of the one before it
      if (chains == 2) for (int i = 0; i < mma_count; ++i) tc_mma(tmem_base + (uint32_t)((i & 1) * 256), adesc + (uint64_t)((i & 3) * 2), bdesc + (uint64_t)((i & 3) * 2), idesc, i > 1);
      else {
        // (no runtime division in the issuing thread: its instruction latency would bound the loop)
        uint32_t c = 0;
        for (int i = 0; i < mma_count; ++i) {
          tc_mma(tmem_base + c * 128u, adesc + (uint64_t)((i & 3) * 2), bdesc + (uint64_t)((i & 3) * 2), idesc, i >= chains);
          c = (c + 1 == (uint32_t)chains) ? 0u : c + 1;
        }
      }
      tc_commit(&bar);
      mbar_wait(&bar, 0);
      const long long t1 = clock64();
      out_mma[blockIdx.x] = t1 - t0;
      if (mma_count == 7) {   // cost of waiting on an ALREADY COMPLETE barrier, and of a commit + wait round trip
        const long long w0 = clock64();
        for (int i = 0; i < 64; ++i) mbar_wait(&bar, 0);
        const long long w1 = clock64();
        long long acc = 0;
        uint32_t ph = 1;
        for (int i = 0; i < 16; ++i) {
          tc_mma(tmem_base, adesc, bdesc, idesc, 0);
          const long long c0 = clock64();
          tc_commit(&bar);
          mbar_wait(&bar, ph);
          acc += clock64() - c0;
          ph ^= 1;
        }
        if (blockIdx.x == 0) printf("wait on complete barrier: %.1f cyc; 1 MMA + commit + wait: %.1f cyc\n", (double)(w1 - w0) / 64.0, (double)acc / 16.0);
      }
    }
  } else if (warp >= 2 && warp < 2 + reader_warps && (mode & 1)) {
    const int quarter = warp & 3;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
    uint32_t x = 0;
    const int ncols = 512 - X;
    __syncwarp();
    const long long t0 = clock64();
    for (int i = 0; i < ld_iters; ++i) {
      x ^= tmem_ld<X>(lane_addr + (uint32_t)((i * X + (warp >> 2) * 64) % ncols));
      if ((i + 1) % wait_every == 0) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    const long long t1 = clock64();
    if (lane == 0) out_ld[blockIdx.x * 16 + (warp - 2)] = t1 - t0;
    if (x == 0xdeadbeef) sink[0] = x;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}

template <int X>
static void run(int mode, int reader_warps, int ld_iters, int mma_n, int mma_count, int wait_every, int chains = 2) {
  long long *d_ld, *d_mma;
  uint32_t* sink;
  cudaMalloc(&d_ld, 148 * 16 * sizeof(long long));
  cudaMalloc(&d_mma, 148 * sizeof(long long));
  cudaMalloc(&sink, 4);
  cudaMemset(d_ld, 0, 148 * 16 * sizeof(long long));
  cudaMemset(d_mma, 0, 148 * sizeof(long long));
  const int smem = 64 * 1024;
  cudaFuncSetAttribute(ubench_kernel<X>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) ubench_kernel<X><<<148, 64 + reader_warps * 32, smem>>>(mode, reader_warps, ld_iters, mma_n, mma_count, wait_every, d_ld, d_mma, sink, chains);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); exit(1); }
  std::vector<long long> ld(148 * 16), mma(148);
  cudaMemcpy(ld.data(), d_ld, ld.size() * sizeof(long long), cudaMemcpyDeviceToHost);
  cudaMemcpy(mma.data(), d_mma, mma.size() * sizeof(long long), cudaMemcpyDeviceToHost);
  long long ld_max = 0, mma_max = 0;
  for (int b = 0; b < 148; ++b) {
    for (int w = 0; w < reader_warps; ++w) ld_max = std::max(ld_max, ld[b * 16 + w]);
    mma_max = std::max(mma_max, mma[b]);
  }
  printf("mode=%d X=%3d readers=%2d wait_every=%d", mode, X, reader_warps, wait_every);
  if (mode & 1) {
    const double bytes = (double)reader_warps * ld_iters * X * 32 * 4;
    printf("  | ld: %lld cyc, %.1f B/clk/SM, %.2f cyc per x%d load per warp", ld_max, bytes / (double)ld_max, (double)ld_max / ld_iters, X);
  }
  if (mode & 2) printf("  | mma N=%d chains=%d: %.1f cyc per MMA (floor %.0f)", mma_n, chains, (double)mma_max / mma_count, 128.0 * mma_n / 256.0);
  printf("\n");
  cudaFree(d_ld); cudaFree(d_mma); cudaFree(sink);
}

int main(int argc, char** argv) {
  if (argc > 1 && argv[1][0] == 'd') {   // dependent accumulation: how long does an MMA take that accumulates into the previous one's result?
    for (int n : {64, 96, 128, 256})
      for (int chains : {1, 2, 3, 4})
        if (n * chains <= 512 && !(chains == 2 && n > 128)) run<32>(2, 4, 0, n, 4000, 2, chains);
    for (int chains : {1, 2, 4}) run<32>(3, 8, 4000, 128, 8000, 2, chains);
    return 0;
  }
  if (argc > 1) { run<32>(2, 4, 0, 256, 7, 1); for (int n : {128, 240, 256}) { run<32>(2, 4, 0, n, 4000, 1); run<32>(2, 4, 0, n, 4000, 2); } for (int readers : {8}) { run<32>(3, readers, 4000, 256, 8000, 1); run<32>(3, readers, 4000, 256, 8000, 2); } return 0; }
  for (int readers : {4, 8, 16}) {
    for (int we : {1, 2}) {
      run<16>(1, readers, 4000, 240, 0, we);
      run<32>(1, readers, 4000, 240, 0, we);
      run<64>(1, readers, 2000, 240, 0, we);
    }
  }
  for (int n : {48, 96, 128, 192, 240, 256}) run<32>(2, 4, 0, n, 4000, 1);
  for (int readers : {4, 8, 16}) run<32>(3, readers, 4000, 240, 8000, 1);
  return 0;
}
