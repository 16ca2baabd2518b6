// Probe: which TMA descriptor / call variants work on this B200 (debug aid, not part of the product).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cstdlib>
typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int RANK>
__global__ void k(const __grid_constant__ CUtensorMap tmap, int bytes, int x, int y, int z, uint32_t* out, int nout) {
  extern __shared__ __align__(1024) uint8_t sm[];
  uint64_t* bar = (uint64_t*)sm;
  uint8_t* dst = sm + 1024;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(bytes) : "memory");
    if (RANK == 3)
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(s32(dst)), "l"(&tmap), "r"(s32(bar)), "r"(x), "r"(y), "r"(z) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(s32(dst)), "l"(&tmap), "r"(s32(bar)), "r"(x), "r"(y) : "memory");
  }
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(s32(bar)), "r"(0) : "memory");
  for (int i = threadIdx.x; i < nout; i += blockDim.x) out[i] = ((uint32_t*)dst)[i];
}
int main(int argc, char** argv) {
  int only = argc > 1 ? atoi(argv[1]) : -1;
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  PFN enc = (PFN)fn;
  int W = 752, H = 480, B = 2;
  std::vector<uint8_t> h((size_t)W * H * B);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)(i * 7 + (i >> 8));
  uint8_t* d; cudaMalloc(&d, h.size()); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
  uint32_t* out; cudaMalloc(&out, 1 << 20);
  struct V { const char* name; int rank; CUtensorMapDataType dt; int esz; int box0, box1; int x, y; };
  V vs[] = {
    {"3D u32 box192x36 (current)", 3, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, 192, 36, -1, -2},
    {"3D u32 box188x36 x=0", 3, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, 188, 36, 0, 0},
    {"3D u32 box64x36 x=-1", 3, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, 64, 36, -1, -2},
    {"3D u32 box32x36", 3, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, 32, 36, -1, -2},
    {"2D u32 box192x36", 2, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, 192, 36, -1, -2},
    {"2D u32 box64x36", 2, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, 64, 36, -1, -2},
    {"3D u8 box256x36", 3, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, 256, 36, -2, -2},
    {"2D u8 box256x36", 2, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, 256, 36, -2, -2},
    {"3D u32 box192x32 y=0", 3, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, 192, 32, 0, 0},
    {"2D u8 box64x64 x=0", 2, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, 64, 64, 0, 0},
    {"2D u32 box32x32 x=0", 2, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, 32, 32, 0, 0},
    {"3D u32 box32x32 x=0", 3, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, 32, 32, 0, 0},
    {"3D u32 box188x36 x=0 y=-2", 3, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, 188, 36, 0, -2},
    {"3D u32 box128x36 x=-1 y=-2", 3, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, 128, 36, -1, -2},
    {"3D u32 box128x36 x=+1", 3, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, 128, 36, 1, 3},
    {"3D u32 box128x36 x=+4", 3, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, 128, 36, 4, 3},
    {"3D u32 box128x36 x=-4", 3, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, 128, 36, -4, -2},
    {"3D u32 box196x36 x=-4", 3, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, 196, 36, -4, -2},
    {"3D u8 box256x36 x=-16", 3, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, 256, 36, -16, -2},
    {"3D u8 box256x36 x=+3", 3, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, 256, 36, 3, 2},
    {"3D u32 box128x36 x=100 (tail oob)", 3, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, 128, 36, 100, 470},
    {"3D u32 box128x36 x=101 (tail oob)", 3, CU_TENSOR_MAP_DATA_TYPE_UINT32, 4, 128, 36, 101, 470},
  };
  int vi = -1;
  for (auto& v : vs) {
    ++vi; if (only >= 0 && vi != only) continue;
    CUtensorMap m;
    cuuint64_t gd[3] = {(cuuint64_t)(W / v.esz), (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t gs[2] = {(cuuint64_t)W, (cuuint64_t)W * H};
    cuuint32_t box[3] = {(cuuint32_t)v.box0, (cuuint32_t)v.box1, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&m, v.dt, v.rank, d, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    int bytes = v.box0 * v.esz * v.box1;
    size_t smem = 1024 + bytes;
    cudaError_t e;
    if (v.rank == 3) { cudaFuncSetAttribute(k<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000); k<3><<<1, 128, smem>>>(m, bytes, v.x, v.y, 1, out, bytes / 4); }
    else { cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100000); k<2><<<1, 128, smem>>>(m, bytes, v.x, v.y, 0, out, bytes / 4); }
    e = cudaDeviceSynchronize();
    std::vector<uint32_t> ho(bytes / 4);
    int good = -1;
    if (e == cudaSuccess) {
      cudaMemcpy(ho.data(), out, bytes, cudaMemcpyDeviceToHost);
      // check a value: row 5 of the box, element 10
      int row = 5, el = 10;
      int X = (v.x + el) * v.esz, Y = v.y + row, Z = (v.rank == 3) ? 1 : 0;
      uint32_t expect = 0;
      if (X >= 0 && Y >= 0) { const uint8_t* p = &h[(size_t)Z * W * H + (size_t)Y * W + X]; if (v.esz == 4) expect = *(const uint32_t*)p; }
      uint32_t got = ho[row * (v.box0 * v.esz / 4) + (el * v.esz) / 4];
      good = (v.esz == 4) ? (got == expect) : 2;
    }
    printf("%-32s encode=%d run=%s check=%d\n", v.name, (int)r, cudaGetErrorString(e), good);
    if (e != cudaSuccess) { cudaDeviceReset(); cudaMalloc(&d, h.size()); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice); cudaMalloc(&out, 1 << 20); }
  }
  return 0;
}
