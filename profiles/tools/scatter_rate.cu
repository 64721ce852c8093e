// scatter_rate.cu -- how fast does one B200 take the kinds of scattered accesses the aggregation kernels are made of?
// Each test: 2 CTAs x 512 threads per SM, every lane of every warp goes to its own random address (4 independent
// accesses in flight per thread), the table fits L2 (32 MB) unless stated.  Prints lane-operations per SM-cycle and G ops/s.
// build + run:  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/scatter_rate profiles/tools/scatter_rate.cu && /tmp/scatter_rate
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t x) {
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL; x ^= x >> 27; x *= 0x94d049bb133111ebULL; x ^= x >> 31;
  return x;
}
constexpr int ITERS = 128, UNROLL = 4;

template <int KIND>
__global__ void __launch_bounds__(512, 2) k(uint4* table, uint64_t mask, unsigned* cursors, uint4* out, unsigned long long* sink, unsigned ncur, unsigned cstride) {
  __shared__ unsigned long long sm[4096];
  for (int i = threadIdx.x; i < 4096; i += 512) sm[i] = i;
  __syncthreads();
  const uint64_t tid = (uint64_t)blockIdx.x * 512 + threadIdx.x;
  unsigned long long acc = 0;
  for (int it = 0; it < ITERS; it += UNROLL) {
    uint64_t a[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) a[u] = mix(tid * 1315423911ull + (uint64_t)(it + u)) & mask;
    if (KIND == 0) {  // 32-byte load (two 16-byte halves of one sector)
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        unsigned long long v0, v1, v2, v3;
        asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(v0), "=l"(v1), "=l"(v2), "=l"(v3) : "l"(table + 2 * (a[u] >> 1)));
        acc += v0 ^ v3;
      }
    } else if (KIND == 1) {  // 64-bit reduction (no return)
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) atomicAdd(reinterpret_cast<unsigned long long*>(table + a[u]), 1ull);
    } else if (KIND == 2) {  // returning 32-bit atomic on one of ncur cursors (cstride words apart) + dependent 16-byte store
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const unsigned p = (unsigned)(a[u] % ncur);
        const unsigned pos = atomicAdd(cursors + cstride * p, 1u);
        out[((size_t)p * 2048u + (pos & 2047u)) & ((1ull << 28) - 1)] = make_uint4((unsigned)a[u], pos, p, 0u);
      }
    } else if (KIND == 9) {  // the 16-byte store into the partition space alone (position from a private counter)
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const unsigned p = (unsigned)(a[u] % ncur);
        out[((size_t)p * 2048u + ((unsigned)(tid + it + u) & 2047u)) & ((1ull << 28) - 1)] = make_uint4((unsigned)a[u], 1u, p, 0u);
      }
    } else if (KIND == 10) {  // returning atomic + dependent 16-byte store into an L2-resident space (32 MB)
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const unsigned p = (unsigned)(a[u] % ncur);
        const unsigned pos = atomicAdd(cursors + cstride * p, 1u);
        table[((size_t)p * 128u + (pos & 127u)) & mask] = make_uint4((unsigned)a[u], pos, p, 0u);
      }
    } else if (KIND == 11) {  // returning atomic + dependent store, partitions spread evenly over `mask + 1` slots of the partition space
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const unsigned p = (unsigned)(a[u] % ncur);
        const unsigned pos = atomicAdd(cursors + cstride * p, 1u);
        const uint64_t per = (mask + 1) / ncur;  // slots per partition (>= 2048 here)
        out[(size_t)p * per + (pos % (unsigned)per)] = make_uint4((unsigned)a[u], pos, p, 0u);
      }
    } else if (KIND == 12) {  // the same with the partitions interleaved in chunks of 256 entries (4 KB): all partitions fill at the same
      // pace, so the stores of any moment fall into one row of chunks (ncur x 4 KB) whatever the total size is
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const unsigned p = (unsigned)(a[u] % ncur);
        const unsigned pos = atomicAdd(cursors + cstride * p, 1u);
        out[(((size_t)(pos >> 8) * ncur + p) << 8) + (pos & 255u)] = make_uint4((unsigned)a[u], pos, p, 0u);
      }
    } else if (KIND == 7) {  // the returning atomic alone
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) acc += atomicAdd(cursors + cstride * (unsigned)(a[u] % ncur), 1u);
    } else if (KIND == 8) {  // a reduction (no return) on the cursors
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) atomicAdd(cursors + cstride * (unsigned)(a[u] % ncur), 1u);
    } else if (KIND == 3) {  // 16-byte store only
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) table[a[u]] = make_uint4((unsigned)a[u], 1u, 2u, 3u);
    } else if (KIND == 4) {  // shared memory: 64-bit load at a random word
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) acc += *reinterpret_cast<volatile unsigned long long*>(&sm[a[u] & 4095u]);
    } else if (KIND == 5) {  // shared memory: 64-bit compare-and-swap at a random word
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) acc += atomicCAS(&sm[a[u] & 4095u], 0ull, a[u] | 1ull);
    } else if (KIND == 6) {  // three 16-byte loads of a 48-byte record, consecutive records per lane (the record read of pass 1)
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const uint4* r = table + 3 * (((uint64_t)(it + u) * 512 * gridDim.x + tid) % (mask / 3));
        const uint4 x = __ldg(r), y = __ldg(r + 1), z = __ldg(r + 2);
        acc += x.x ^ y.y ^ z.z;
      }
    }
  }
  if (acc == 0x1234567ull) *sink = acc;
}

template <int KIND>
void run(const char* what, uint4* table, uint64_t mask, unsigned* cursors, uint4* out, unsigned long long* sink, int sms, double mhz, unsigned ncur = 16384, unsigned cstride = 8) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e9f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaMemset(cursors, 0, (size_t)1 << 26);
    cudaEventRecord(e0);
    k<KIND><<<sms * 2, 512>>>(table, mask, cursors, out, sink, ncur, cstride);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  const double ops = (double)sms * 2 * 512 * ITERS;
  printf("%-78s %8.1f us  %6.1f G ops/s  %5.2f lane-ops per SM-cycle (%.2f cycles per lane-op)\n", what, best * 1e3, ops / best / 1e6,
         ops / (best * 1e-3 * mhz * 1e6 * sms), best * 1e-3 * mhz * 1e6 * sms / ops);
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double mhz = khz / 1000.0;
  const uint64_t n16 = 1ull << 21;  // 2 M x 16 B = 32 MB
  uint4 *table, *out;
  unsigned* cursors;
  unsigned long long* sink;
  cudaMalloc(&table, n16 * 16);
  cudaMemset(table, 0, n16 * 16);
  cudaMalloc(&out, (size_t)16384 * 16384 * 16);  // 4 GB of partition space
  cudaMalloc(&cursors, (size_t)1 << 26);
  cudaMalloc(&sink, 8);
  printf("%s, %d SMs, %.0f MHz nominal; %d lane-ops per thread, %d in flight\n", p.name, sms, mhz, ITERS, UNROLL);
  run<0>("global: 32-byte load, random sector of a 32-MB table", table, n16 - 1, cursors, out, sink, sms, mhz);
  run<6>("global: 48-byte record read (3 x LDG.128 at 48-byte stride), streaming", table, n16 - 1, cursors, out, sink, sms, mhz);
  run<1>("global: 64-bit reduction (RED), random word of a 32-MB table", table, n16 - 1, cursors, out, sink, sms, mhz);
  run<2>("global: returning atomic on 1 of 16384 cursors (32 B apart) + dependent 16-byte store", table, n16 - 1, cursors, out, sink, sms, mhz);
  run<7>("global: returning atomic alone, 16384 cursors 32 B apart", table, n16 - 1, cursors, out, sink, sms, mhz, 16384, 8);
  run<7>("global: returning atomic alone, 16384 cursors 128 B apart", table, n16 - 1, cursors, out, sink, sms, mhz, 16384, 32);
  run<7>("global: returning atomic alone, 131072 cursors 32 B apart", table, n16 - 1, cursors, out, sink, sms, mhz, 131072, 8);
  run<7>("global: returning atomic alone, 1048576 cursors 32 B apart", table, n16 - 1, cursors, out, sink, sms, mhz, 1048576, 8);
  run<7>("global: returning atomic alone, 2048 cursors 32 B apart", table, n16 - 1, cursors, out, sink, sms, mhz, 2048, 8);
  run<8>("global: reduction (no return), 16384 cursors 32 B apart", table, n16 - 1, cursors, out, sink, sms, mhz, 16384, 8);
  run<2>("global: returning atomic on 1 of 131072 cursors (32 B apart) + dependent 16-byte store", table, n16 - 1, cursors, out, sink, sms, mhz, 131072, 8);
  run<11>("global: returning atomic + dependent store, 16384 partitions over  64 MB", table, (1ull << 22) - 1, cursors, out, sink, sms, mhz, 16384, 8);
  run<11>("global: returning atomic + dependent store, 16384 partitions over 256 MB", table, (1ull << 24) - 1, cursors, out, sink, sms, mhz, 16384, 8);
  run<11>("global: returning atomic + dependent store, 16384 partitions over 512 MB", table, (1ull << 25) - 1, cursors, out, sink, sms, mhz, 16384, 8);
  run<11>("global: returning atomic + dependent store, 16384 partitions over   1 GB", table, (1ull << 26) - 1, cursors, out, sink, sms, mhz, 16384, 8);
  run<11>("global: returning atomic + dependent store, 16384 partitions over   2 GB", table, (1ull << 27) - 1, cursors, out, sink, sms, mhz, 16384, 8);
  run<11>("global: returning atomic + dependent store, 16384 partitions over   4 GB", table, (1ull << 28) - 1, cursors, out, sink, sms, mhz, 16384, 8);
  run<12>("global: returning atomic + dependent store, 16384 partitions interleaved in 4-KB chunks", table, n16 - 1, cursors, out, sink, sms, mhz, 16384, 8);
  run<9>("global: 16-byte store into 4 GB of partition space alone (16384 partitions)", table, n16 - 1, cursors, out, sink, sms, mhz, 16384, 8);
  run<10>("global: returning atomic + dependent 16-byte store into 32 MB (16384 partitions)", table, n16 - 1, cursors, out, sink, sms, mhz, 16384, 8);
  run<3>("global: 16-byte store, random slot of a 32-MB table", table, n16 - 1, cursors, out, sink, sms, mhz);
  run<4>("shared: 64-bit load, random word of 32 KB", table, n16 - 1, cursors, out, sink, sms, mhz);
  run<5>("shared: 64-bit compare-and-swap, random word of 32 KB", table, n16 - 1, cursors, out, sink, sms, mhz);
  return 0;
}
