// mcb_probe.cu -- measured ceiling of the operation that bounds the photon kernels: fully divergent 4-byte gathers,
// one 32-byte sector per lane per load, several loads in flight per lane.
//
// The flux kernels march rays through a packed f32 extinction field; every crossing is one gather whose address
// depends on the ray alone, so the lanes of a warp touch 32 different sectors.  On the L2-resident fields (C1-C3) HBM
// is idle and what limits the kernel is the rate at which an SM can push L1 misses through its L1TEX->XBAR port and
// the L2 slices can look them up (ncu, round 1).  This kernel measures that rate directly so that the roofline in
// bench.py is a fraction of a MEASURED peak: same launch shape as the flux kernels (persistent, 128 threads, a given
// number of CTAs per SM), `inFlight` independent loads per lane per iteration at pseudo-random addresses inside a
// buffer of `bytes` bytes (16 MB: L2-resident like the C3 field; 78 MB: the C5 field; >= 512 MB: HBM-bound), next
// to nothing else (3 integer instructions per address).  A negative `inFlight` issues 16-byte loads (float4) instead.
#include <cuda_runtime.h>
#include <stdint.h>

namespace mcbprobe {

template <int INFLIGHT>
__global__ void __launch_bounds__(128) gather_kernel(const float *__restrict__ buf, uint32_t nWords, int iterations, float *sink) {
  uint32_t s = (blockIdx.x * 128u + threadIdx.x) * 2654435761u + 12345u;
  float acc = 0.0f;
  for (int it = 0; it < iterations; ++it) {
    float v[INFLIGHT];
#pragma unroll
    for (int k = 0; k < INFLIGHT; ++k) {
      s = s * 1664525u + 1013904223u;                       // LCG: the high bits are well mixed
      v[k] = __ldg(buf + __umulhi(s, nWords));
    }
#pragma unroll
    for (int k = 0; k < INFLIGHT; ++k) acc += v[k];
  }
  if (acc == 1.2345e-30f) *sink = acc;                      // keeps the loads alive
}

// the same with 16-byte loads (one aligned float4 per lane per load): does a wider gather cost more than a 4-byte one?
template <int INFLIGHT>
__global__ void __launch_bounds__(128) gather4_kernel(const float4 *__restrict__ buf, uint32_t nQuads, int iterations, float *sink) {
  uint32_t s = (blockIdx.x * 128u + threadIdx.x) * 2654435761u + 12345u;
  float acc = 0.0f;
  for (int it = 0; it < iterations; ++it) {
    float4 v[INFLIGHT];
#pragma unroll
    for (int k = 0; k < INFLIGHT; ++k) {
      s = s * 1664525u + 1013904223u;
      v[k] = __ldg(buf + __umulhi(s, nQuads));
    }
#pragma unroll
    for (int k = 0; k < INFLIGHT; ++k) acc += v[k].x + v[k].y + v[k].z + v[k].w;
  }
  if (acc == 1.2345e-30f) *sink = acc;
}

__global__ void fill_kernel(float *buf, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) buf[i] = 1.0f;
}

}  // namespace mcbprobe

// returns gathers per second (each gather = one 32-byte sector request), or a negative CUDA error code
double mcb_run_gather_probe(size_t bytes, int inFlight, int blocksPerSM, int iterations, int numSMs, cudaStream_t stream) {
  float *buf = nullptr, *sink = nullptr;
  const size_t n = bytes / sizeof(float);
  if (n < 1024 || n >= (1ull << 32)) return -1.0;
  if (cudaMalloc((void **)&buf, n * sizeof(float)) != cudaSuccess) return -2.0;
  if (cudaMalloc((void **)&sink, sizeof(float)) != cudaSuccess) { cudaFree(buf); return -2.0; }
  mcbprobe::fill_kernel<<<numSMs * 8, 256, 0, stream>>>(buf, n);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int blocks = numSMs * blocksPerSM;
  auto go = [&](int iters) {
    if (inFlight < 0) {                                     // 16-byte loads
      if (inFlight == -1) mcbprobe::gather4_kernel<1><<<blocks, 128, 0, stream>>>((const float4 *)buf, (uint32_t)(n / 4), iters, sink);
      else if (inFlight == -4) mcbprobe::gather4_kernel<4><<<blocks, 128, 0, stream>>>((const float4 *)buf, (uint32_t)(n / 4), iters, sink);
      else { mcbprobe::gather4_kernel<8><<<blocks, 128, 0, stream>>>((const float4 *)buf, (uint32_t)(n / 4), iters, sink); inFlight = -8; }
      return;
    }
    switch (inFlight) {
      case 1: mcbprobe::gather_kernel<1><<<blocks, 128, 0, stream>>>(buf, (uint32_t)n, iters, sink); break;
      case 2: mcbprobe::gather_kernel<2><<<blocks, 128, 0, stream>>>(buf, (uint32_t)n, iters, sink); break;
      case 4: mcbprobe::gather_kernel<4><<<blocks, 128, 0, stream>>>(buf, (uint32_t)n, iters, sink); break;
      case 16: mcbprobe::gather_kernel<16><<<blocks, 128, 0, stream>>>(buf, (uint32_t)n, iters, sink); break;
      default: mcbprobe::gather_kernel<8><<<blocks, 128, 0, stream>>>(buf, (uint32_t)n, iters, sink); inFlight = 8; break;
    }
  };
  go(iterations / 8 + 1);                                   // warm-up: page the buffer into L2 where it fits
  double best = 0.0;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0, stream);
    go(iterations);
    cudaEventRecord(e1, stream);
    if (cudaEventSynchronize(e1) != cudaSuccess) { best = -3.0; break; }
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double rate = (double)blocks * 128.0 * (double)iterations * (double)(inFlight < 0 ? -inFlight : inFlight) / (ms * 1e-3);
    if (rate > best) best = rate;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(buf); cudaFree(sink);
  if (cudaGetLastError() != cudaSuccess) return -4.0;
  return best;
}
