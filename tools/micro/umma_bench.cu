// Developer microbenchmark: cost of tcgen05.mma streams of small shapes issued by one thread (cycles per MMA).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_bench umma_bench.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../rosettafold-pytorch_b200/csrc/rfk_common.cuh"
using namespace rfk;

__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
               "r"(a), "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}

// mode 0: SS (A, B from smem), mode 1: TS (A from TMEM). group: MMAs per commit. reps: groups.
template <int MODE, int ELECT_ONCE>
__global__ void __launch_bounds__(128, 1) bench(int N, int group, int reps, int wait_each, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar = base + 65536, slot = base + 65536 + 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 16384; i += 128) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) {
    const uint64_t da = umma_desc_sw128(base), db = umma_desc_sw128(base + 32768);
    const uint32_t idesc = umma_idesc_bf16(128, N);
    uint32_t par = 0;
    const long long t0 = clock64();
    if (ELECT_ONCE) {
      // CUTLASS style: ONE election, the elected thread runs the whole loop (waits included)
      uint32_t el;
      asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(el));
      if (el) {
        for (int r = 0; r < reps; ++r) {
          if (group == 4) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (MODE == 0) umma_bf16(0, da + 2 * k, db + 2 * k, idesc, 1);
              else mma_ts(0, 256 + 8 * k, db + 2 * k, idesc, 1);
            }
          } else if (group == 16) {
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              if (MODE == 0) umma_bf16(0, da + 2 * (k & 3), db + 2 * (k & 3), idesc, 1);
              else mma_ts(0, 256 + 8 * (k & 3), db + 2 * (k & 3), idesc, 1);
            }
          } else {
            for (int k = 0; k < group; ++k) {
              if (MODE == 0) umma_bf16(0, da + 2 * (k & 3), db + 2 * (k & 3), idesc, 1);
              else mma_ts(0, 256 + 8 * (k & 3), db + 2 * (k & 3), idesc, 1);
            }
          }
          if (wait_each || r == reps - 1) {
            umma_commit(bar);
            mbar_wait(bar, par);
            par ^= 1;
          }
        }
      }
      __syncwarp();
    } else {
    for (int r = 0; r < reps; ++r) {
      uint32_t el;
      asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(el));
      if (el) {
      if (group == 4) {  // unrolled, constant operand offsets
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (MODE == 0) umma_bf16(0, da + 2 * k, db + 2 * k, idesc, 1);
          else mma_ts(0, 256 + 8 * k, db + 2 * k, idesc, 1);
        }
      } else if (group == 16) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          if (MODE == 0) umma_bf16(0, da + 2 * (k & 3), db + 2 * (k & 3), idesc, 1);
          else mma_ts(0, 256 + 8 * (k & 3), db + 2 * (k & 3), idesc, 1);
        }
      } else {
        for (int k = 0; k < group; ++k) {
          if (MODE == 0) umma_bf16(0, da + 2 * (k & 3), db + 2 * (k & 3), idesc, 1);
          else mma_ts(0, 256 + 8 * (k & 3), db + 2 * (k & 3), idesc, 1);
        }
      }
      if (wait_each || r == reps - 1) umma_commit(bar);
      }
      __syncwarp();
      if (wait_each || r == reps - 1) {
        mbar_wait(bar, par);
        par ^= 1;
      }
    }
    }
    const long long t1 = clock64();
    if (lane == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(0, 512); }
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  const int smem = 65536 + 1024 + 1024;
  cudaFuncSetAttribute(bench<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(bench<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(bench<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(bench<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int Ns[] = {16, 80, 128};
  for (int once = 0; once < 2; ++once)
  for (int mode = 0; mode < 2; ++mode)
    for (int N : Ns)
      for (int wait_each = 0; wait_each < 2; ++wait_each) {
        for (int group : {4, 16, 5}) {
        const int reps = 2000;
        for (int it = 0; it < 2; ++it) {
          if (once == 0) {
            if (mode == 0) bench<0, 0><<<1, 128, smem>>>(N, group, reps, wait_each, d);
            else bench<1, 0><<<1, 128, smem>>>(N, group, reps, wait_each, d);
          } else {
            if (mode == 0) bench<0, 1><<<1, 128, smem>>>(N, group, reps, wait_each, d);
            else bench<1, 1><<<1, 128, smem>>>(N, group, reps, wait_each, d);
          }
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        }
        long long c;
        cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
        printf("%s %s M=128 N=%3d K=16  group=%d %s: %.1f cycles / MMA (math floor %d)\n", once ? "elect-once    " : "elect-per-group", mode ? "TS" : "SS", N, group,
               wait_each ? "commit+wait per group" : "back-to-back          ", (double)c / (group * reps), 128 * N / 256);
        }
      }
  return 0;
}
