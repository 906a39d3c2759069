// Micro-benchmark (tuning helper, not part of the product): throughput of random 8-byte gathers from an array
// much larger than L2, for different load flavours.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t mix(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }
template <int MODE> __device__ __forceinline__ double ld(const double* p) {
  double v;
  if (MODE == 0) v = *p;
  else if (MODE == 1) v = __ldg(p);
  else if (MODE == 2) asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p));
  else if (MODE == 3) asm volatile("ld.global.cv.f64 %0, [%1];" : "=d"(v) : "l"(p));
  else if (MODE == 4) asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  else if (MODE == 5) asm volatile("ld.global.L2::64B.f64 %0, [%1];" : "=d"(v) : "l"(p));
  else if (MODE == 6) asm volatile("ld.global.nc.L1::no_allocate.L2::64B.f64 %0, [%1];" : "=d"(v) : "l"(p));
  else if (MODE == 7) asm volatile("ld.global.lu.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
template <int MODE, int PAIR> __global__ void k(const double* a, uint64_t n, uint64_t row, double* out, int per) {
  uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  double acc = 0;
  for (int i = 0; i < per; i += 4) {
    uint64_t i0 = mix(t * per + i) % (n - row - 2), i1 = mix(t * per + i + 1) % (n - row - 2), i2 = mix(t * per + i + 2) % (n - row - 2), i3 = mix(t * per + i + 3) % (n - row - 2);
    double v0 = ld<MODE>(a + i0), v1 = ld<MODE>(a + i1), v2 = ld<MODE>(a + i2), v3 = ld<MODE>(a + i3);
    if (PAIR) { v0 += ld<MODE>(a + i0 + 1) + ld<MODE>(a + i0 + row) + ld<MODE>(a + i0 + row + 1);
                v1 += ld<MODE>(a + i1 + 1) + ld<MODE>(a + i1 + row) + ld<MODE>(a + i1 + row + 1);
                v2 += ld<MODE>(a + i2 + 1) + ld<MODE>(a + i2 + row) + ld<MODE>(a + i2 + row + 1);
                v3 += ld<MODE>(a + i3 + 1) + ld<MODE>(a + i3 + row) + ld<MODE>(a + i3 + row + 1); }
    acc += v0 + v1 + v2 + v3;
  }
  out[t] = acc;
}
__global__ void k256(const double* a, uint64_t n, uint64_t row, double* out, int per) {
  uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  double acc = 0;
  for (int i = 0; i < per; i += 4) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint64_t i0 = mix(t * per + i + j) % (n - row - 8);
      const double* p0 = a + (i0 & ~3ull);
      double x0, x1, x2, x3, y0, y1, y2, y3;
      asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(x0), "=d"(x1), "=d"(x2), "=d"(x3) : "l"(p0));
      asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(y0), "=d"(y1), "=d"(y2), "=d"(y3) : "l"(p0 + row));
      const int o = (int)(i0 & 3);
      double a0 = o == 0 ? x0 : o == 1 ? x1 : o == 2 ? x2 : x3, a1 = o == 0 ? x1 : o == 1 ? x2 : x3;
      double b0 = o == 0 ? y0 : o == 1 ? y1 : o == 2 ? y2 : y3, b1 = o == 0 ? y1 : o == 1 ? y2 : y3;
      if (o == 3) { a1 = __ldg(p0 + 4); b1 = __ldg(p0 + row + 4); }
      acc += a0 + a1 + b0 + b1;
    }
  }
  out[t] = acc;
}
template <int MODE, int PAIR> void run(const double* a, uint64_t n, double* out, const char* name, uint64_t row = 16384) {
  const int threads = 256, blocks = 148 * 64, per = 32;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE, PAIR><<<blocks, threads>>>(a, n, row, out, per);
  cudaEventRecord(e0);
  for (int r = 0; r < 3; ++r) k<MODE, PAIR><<<blocks, threads>>>(a, n, row, out, per);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
  double pts = (double)blocks * threads * per;
  printf("%-34s pair=%d  %.3f ms  %.2f G gather-points/s  err=%s\n", name, PAIR, ms, pts / ms * 1e-6, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  const uint64_t n = 16384ull * 16384ull;
  double *a, *out; cudaMalloc(&a, n * 8); cudaMalloc(&out, 148 * 64 * 256 * 8); cudaMemset(a, 0, n * 8);
  run<0, 0>(a, n, out, "ld.global"); run<1, 0>(a, n, out, "__ldg (nc)"); run<2, 0>(a, n, out, "ld.global.cg"); run<3, 0>(a, n, out, "ld.global.cv");
  run<4, 0>(a, n, out, "nc.L1::no_allocate"); run<5, 0>(a, n, out, "L2::64B"); run<6, 0>(a, n, out, "nc.no_allocate.L2::64B"); run<7, 0>(a, n, out, "ld.global.lu");
  run<0, 1>(a, n, out, "ld.global"); run<1, 1>(a, n, out, "__ldg (nc)"); run<2, 1>(a, n, out, "ld.global.cg"); run<4, 1>(a, n, out, "nc.L1::no_allocate");
  {
    const int threads = 256, blocks = 148 * 64, per = 32;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k256<<<blocks, threads>>>(a, n, 16384, out, per);
    cudaEventRecord(e0);
    for (int r = 0; r < 3; ++r) k256<<<blocks, threads>>>(a, n, 16384, out, per);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 3;
    printf("%-34s pair=1  %.3f ms  %.2f G gather-points/s  err=%s\n", "256-bit sector loads", ms, (double)blocks * threads * per / ms * 1e-6, cudaGetErrorString(cudaGetLastError()));
  }
  run<1, 1>(a, n, out, "__ldg row=16384+16", 16384 + 16); run<1, 1>(a, n, out, "__ldg row=16384+32", 16384 + 32); run<1, 1>(a, n, out, "__ldg row=16384+256", 16384 + 256); run<1, 1>(a, n, out, "__ldg row=16", 16); run<1, 1>(a, n, out, "__ldg row=4096", 4096); run<1, 1>(a, n, out, "__ldg row=1000003", 1000003);
  return 0;
}
