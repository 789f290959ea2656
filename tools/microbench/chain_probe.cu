// chain_probe.cu -- how many independent IMAD.WIDE.U32.X carry chains does one
// warp need in flight, at W warps per SM sub-partition, to keep the multiplier
// pipe busy?  Each chain is a serial mad.lo.cc/madc.hi.cc sequence (the shape of
// a CIOS row); C chains are interleaved by ptxas.  Prints ms and the fraction of
// the 4-cycles-per-warp-instruction pipe rate.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ void mad_wide_cc(uint64_t& d, uint32_t a, uint32_t b) {
  asm volatile("{\n\t.reg .u32 cl, ch, dl, dh;\n\tmov.b64 {cl, ch}, %0;\n\t"
               "mad.lo.cc.u32 dl, %1, %2, cl;\n\tmadc.hi.cc.u32 dh, %1, %2, ch;\n\t"
               "mov.b64 %0, {dl, dh};\n\t}" : "+l"(d) : "r"(a), "r"(b));
}
__device__ __forceinline__ void madc_wide_cc(uint64_t& d, uint32_t a, uint32_t b) {
  asm volatile("{\n\t.reg .u32 cl, ch, dl, dh;\n\tmov.b64 {cl, ch}, %0;\n\t"
               "madc.lo.cc.u32 dl, %1, %2, cl;\n\tmadc.hi.cc.u32 dh, %1, %2, ch;\n\t"
               "mov.b64 %0, {dl, dh};\n\t}" : "+l"(d) : "r"(a), "r"(b));
}

template <int C, int LEN>
__global__ void __launch_bounds__(128) chains(uint32_t* out, uint32_t a, uint32_t b, int iters,
                                              unsigned long long* cyc) {
  const long long c0 = clock64();
  uint64_t acc[C][LEN];
  uint32_t sink[C];
#pragma unroll
  for (int c = 0; c < C; c++) {
    sink[c] = c;
#pragma unroll
    for (int j = 0; j < LEN; j++) acc[c][j] = ((uint64_t)(threadIdx.x + c) << 32) | (blockIdx.x + j);
  }
  const uint32_t m = a + threadIdx.x;
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int c = 0; c < C; c++) {
      mad_wide_cc(acc[c][0], m, b);
#pragma unroll
      for (int j = 1; j < LEN; j++) madc_wide_cc(acc[c][j], m, b);
      asm volatile("addc.u32 %0, %0, 0;" : "+r"(sink[c]));
    }
  }
  uint32_t s = m;
#pragma unroll
  for (int c = 0; c < C; c++) {
    s ^= sink[c];
#pragma unroll
    for (int j = 0; j < LEN; j++) s ^= (uint32_t)acc[c][j] ^ (uint32_t)(acc[c][j] >> 32);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  // SM cycles this warp was running (clock64 counts at the SM clock)
  if ((threadIdx.x & 31) == 0) atomicMax(cyc, (unsigned long long)(clock64() - c0));
}

template <int C, int LEN>
void run(int warps_per_sm, uint32_t* d, int sms, unsigned long long* d_cyc) {
  const int iters = 20000 / (C * LEN) * 8;
  const int blocks = sms * warps_per_sm / 4;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  chains<C, LEN><<<blocks, 128>>>(d, 3, 5, iters, d_cyc);
  cudaMemset(d_cyc, 0, 8);
  cudaEventRecord(e0);
  chains<C, LEN><<<blocks, 128>>>(d, 3, 5, iters, d_cyc);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  unsigned long long cycles = 0;
  cudaMemcpy(&cycles, d_cyc, 8, cudaMemcpyDeviceToHost);
  const double wide = (double)iters * C * LEN;              // per warp
  const double cyc = (double)cycles;                        // SM cycles of the slowest warp
  const double per_smsp = wide * warps_per_sm / 4.0;        // warp-instructions per sub-partition
  printf("{\"warps_per_sm\": %d, \"chains\": %d, \"len\": %d, \"ms\": %.3f, \"cycles_per_wide_per_warp\": %.2f, "
         "\"pipe_frac\": %.3f}\n", warps_per_sm, C, LEN, ms, cyc / wide, per_smsp * 4.0 / cyc);
}

int main() {
  cudaDeviceProp pr;
  cudaGetDeviceProperties(&pr, 0);
  uint32_t* d;
  cudaMalloc(&d, 64 << 20);
  unsigned long long* ghz;  // device counter: max cycles over warps
  cudaMalloc(&ghz, 8);
  for (int w : {4, 8, 12, 16}) {
    run<1, 16>(w, d, pr.multiProcessorCount, ghz);
    run<2, 16>(w, d, pr.multiProcessorCount, ghz);
    run<3, 16>(w, d, pr.multiProcessorCount, ghz);
    run<4, 16>(w, d, pr.multiProcessorCount, ghz);
    run<6, 16>(w, d, pr.multiProcessorCount, ghz);
    run<8, 8>(w, d, pr.multiProcessorCount, ghz);
  }
  return 0;
}
