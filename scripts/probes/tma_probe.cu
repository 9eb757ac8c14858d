// Probe of TMA tile-mode constraints on sm_100a: one 64x128 bf16 box, 128B swizzle.
// usage: tma_probe cols rows ld c0 c1   -> prints OK (and a checksum) or the CUDA error
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

__global__ void probe(const __grid_constant__ CUtensorMap tm, int c0, int c1, float* out) {
  extern __shared__ __align__(1024) unsigned char sm[];
  unsigned base = (static_cast<unsigned>(__cvta_generic_to_shared(sm)) + 1023u) & ~1023u;
  unsigned bar = base + 16384;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(16384) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(base), "l"(&tm), "r"(bar), "r"(c0), "r"(c1) : "memory");
    unsigned ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(0) : "memory");
    }
  }
  __syncthreads();
  const __nv_bfloat16* t = reinterpret_cast<const __nv_bfloat16*>(sm + (base - static_cast<unsigned>(__cvta_generic_to_shared(sm))));
  float s = 0.f;
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) s += __bfloat162float(t[i]);
  atomicAdd(out, s);
}

int main(int argc, char** argv) {
  long cols = atol(argv[1]), rows = atol(argv[2]), ld = atol(argv[3]);
  int c0 = atoi(argv[4]), c1 = atoi(argv[5]);
  std::vector<__nv_bfloat16> h(ld * rows);
  for (long i = 0; i < ld * rows; ++i) h[i] = __float2bfloat16(1.0f);
  __nv_bfloat16* d; float* o;
  cudaMalloc(&d, ld * rows * 2); cudaMalloc(&o, 4); cudaMemset(o, 0, 4);
  cudaMemcpy(d, h.data(), ld * rows * 2, cudaMemcpyHostToDevice);
  void* sym = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
  typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                         CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  CUtensorMap tm;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows}; cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, 128}; cuuint32_t es[2] = {1, 1};
  CUresult r = ((Fn)sym)(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("cols=%ld rows=%ld ld=%ld c0=%d c1=%d: encode failed %d\n", cols, rows, ld, c0, c1, (int)r); return 0; }
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 20000);
  probe<<<1, 128, 20000>>>(tm, c0, c1, o);
  cudaError_t e = cudaDeviceSynchronize();
  float sum = 0; if (e == cudaSuccess) cudaMemcpy(&sum, o, 4, cudaMemcpyDeviceToHost);
  printf("cols=%ld rows=%ld ld=%ld c0=%d c1=%d: %s  in-bounds elements loaded=%.0f\n", cols, rows, ld, c0, c1, cudaGetErrorString(e), sum);
  return 0;
}
