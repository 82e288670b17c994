// Shared-memory load cost of the access patterns the step kernel uses (not shipped):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_patterns tools/ubench/lds_patterns.cu && ./lds_patterns
// Every SM runs 16 warps (4 per sub-partition) that do nothing but LDS of one width with one address pattern; reported:
// cycles per warp-level load instruction per SM = the number of shared-memory wavefronts the pattern costs
// (the data pipe delivers one 128-byte wavefront per cycle per SM).
#include <cuda_runtime.h>
#include <stdio.h>
#define ITERS 4096
#define TYPE_STRIDE 1056   // sizeof(DsTypeDev)

template <int W> struct V;
template <> struct V<4> { typedef float T; };
template <> struct V<8> { typedef float2 T; };
template <> struct V<16> { typedef float4 T; };
__device__ float sum(float a) { return a; }
__device__ float sum(float2 a) { return a.x + a.y; }
__device__ float sum(float4 a) { return a.x + a.y + a.z + a.w; }

// PAT 0: all lanes one address   1: per-type structs, lanes = slots r t r t r t r t h h h h h h h h (x2 envs)
// PAT 2: consecutive rows (lane l -> row l)   3: slot-major table: lane l -> entry (l & 15)
// PAT 4: two types alternating per lane (quads only)
template <int W, int PAT>
__global__ void __launch_bounds__(512) k(float* out, long long* clk, int off) {
  extern __shared__ __align__(16) unsigned char sm[];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = i * 1e-3f;
  __syncthreads();
  const int l = threadIdx.x & 31, slot = l & 15;
  int base;
  if (PAT == 0) base = 0;
  else if (PAT == 1) base = (slot >= 8 ? 2 : (slot & 1)) * TYPE_STRIDE;
  else if (PAT == 2) base = l * W;
  else if (PAT == 3) base = slot * W;
  else base = (slot & 1) * TYPE_STRIDE;
  typedef typename V<W>::T T;
  float acc = 0.f;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const T v = *reinterpret_cast<const T*>(sm + base + ((u * 16 * 33 + off) & 1023 & ~15));
      acc += sum(v);
    }
    off += 16;
  }
  long long t1 = clock64();
  if (acc == 123.456f) out[0] = acc;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int W, int PAT>
static void run(const char* name, int sms) {
  float* d; long long* c;
  cudaMalloc(&d, 4); cudaMalloc(&c, sizeof(long long) * sms);
  cudaFuncSetAttribute(k<W, PAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  for (int r = 0; r < 2; ++r) { k<W, PAT><<<sms, 512, 65536>>>(d, c, 0); cudaDeviceSynchronize(); }
  long long h[1024];
  cudaMemcpy(h, c, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
  double mean = 0; for (int i = 0; i < sms; ++i) mean += h[i]; mean /= sms;
  printf("LDS.%-3d %-52s %.2f cycles / warp-load / SM\n", W * 8, name, mean / (16.0 * ITERS * 8));
  cudaFree(d); cudaFree(c);
}

int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
#define ALL(P, N) run<4, P>(N, sms); run<8, P>(N, sms); run<16, P>(N, sms);
  ALL(0, "all lanes one address (homogeneous swarm)")
  ALL(1, "per-type structs, mixed warp r t r t r t r t h x 8")
  ALL(4, "per-type structs, two quad types alternating")
  ALL(2, "consecutive rows (downwash snapshot)")
  ALL(3, "slot-major table, lane -> entry lane & 15")
  return 0;
}
