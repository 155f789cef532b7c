// standalone TMA 3D tile load probe (development aid)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>
#include <cstdlib>
typedef unsigned long long u64;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
extern __shared__ __align__(128) unsigned char dsm[];
template <typename L, int RFP>
__global__ void k(const __grid_constant__ CUtensorMap tmap, int c0, int c1, int c2, L* out, uint32_t* info) {
  L* lab = reinterpret_cast<L*>(dsm);
  u64* bar = reinterpret_cast<u64*>(dsm + ((sizeof(L) * RFP * 81 + 127) / 128) * 128);
  if (threadIdx.x == 0) {
    info[0] = smem_u32(lab); info[1] = smem_u32(bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"((uint32_t)(sizeof(L) * RFP * 81)) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(lab)), "l"(&tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
  }
  uint32_t done;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(smem_u32(bar)), "r"(0) : "memory");
  } while (!done);
  __syncthreads();
  for (int i = threadIdx.x; i < RFP * 81; i += blockDim.x) out[i] = lab[i];
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
template <typename L, int RFP>
int run(EncodeTiledFn enc, CUtensorMapDataType dt, uint32_t nf, uint32_t nm, uint32_t ns, int c0, int c1, int c2) {
  size_t n = (size_t)nf * nm * ns;
  std::vector<L> h(n);
  for (size_t i = 0; i < n; ++i) h[i] = (L)(i + 1);
  L* d; cudaMalloc(&d, n * sizeof(L)); cudaMemcpy(d, h.data(), n * sizeof(L), cudaMemcpyHostToDevice);
  L* out; cudaMalloc(&out, sizeof(L) * RFP * 81); uint32_t* info; cudaMalloc(&info, 8);
  CUtensorMap tm; memset(&tm, 0, sizeof(tm));
  cuuint64_t dims[3] = {nf, nm, ns}; cuuint64_t strides[2] = {(cuuint64_t)nf * sizeof(L), (cuuint64_t)nf * nm * sizeof(L)};
  cuuint32_t box[3] = {RFP, 9, 9}; cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(&tm, dt, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("L=%zu RFP=%d dims=(%u,%u,%u) c=(%d,%d,%d): encode rc=%d\n", sizeof(L), RFP, nf, nm, ns, c0, c1, c2, (int)r);
  if (r != CUDA_SUCCESS) return 1;
  size_t smem = ((sizeof(L) * RFP * 81 + 127) / 128) * 128 + 64;
  cudaFuncSetAttribute(k<L, RFP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<L, RFP><<<1, 256, smem>>>(tm, c0, c1, c2, out, info);
  cudaError_t e = cudaDeviceSynchronize();
  printf("  kernel: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 2;
  std::vector<L> o(RFP * 81); uint32_t hi[2];
  cudaMemcpy(o.data(), out, sizeof(L) * RFP * 81, cudaMemcpyDeviceToHost); cudaMemcpy(hi, info, 8, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int s = 0; s < 9; ++s) for (int m = 0; m < 9; ++m) for (int f = 0; f < RFP; ++f) {
    long jf = c0 + f, jm = c1 + m, js = c2 + s;
    L want = 0;
    if (jf >= 0 && jf < nf && jm >= 0 && jm < nm && js >= 0 && js < ns) want = (L)(((size_t)js * nm + jm) * nf + jf + 1);
    if (o[(s * 9 + m) * RFP + f] != want) ++bad;
  }
  printf("  smem lab=0x%x bar=0x%x mismatches=%d\n", hi[0], hi[1], bad);
  return bad;
}
int main(int argc, char** argv) {
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)p;
  int eb = atoi(argv[1]), c0 = atoi(argv[2]), c1 = atoi(argv[3]), c2 = atoi(argv[4]);
  uint32_t nf = argc > 5 ? atoi(argv[5]) : 64;
  if (eb == 4) return run<uint32_t, 36>(enc, CU_TENSOR_MAP_DATA_TYPE_UINT32, nf, 40, 36, c0, c1, c2);
  if (eb == 8) return run<u64, 34>(enc, CU_TENSOR_MAP_DATA_TYPE_UINT64, nf, 40, 36, c0, c1, c2);
  if (eb == 1) return run<uint8_t, 48>(enc, CU_TENSOR_MAP_DATA_TYPE_UINT8, nf, 40, 36, c0, c1, c2);
  if (eb == 2) return run<uint16_t, 40>(enc, CU_TENSOR_MAP_DATA_TYPE_UINT16, nf, 40, 36, c0, c1, c2);
  return 0;
}
