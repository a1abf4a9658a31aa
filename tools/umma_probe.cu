// Probe: can a tcgen05 K-major SWIZZLE_128B A operand be a SHIFTED VIEW of a TMA-loaded halo tile?
//
// A halo tile of (TH+2) x (TW+2) pixels x 32 channels (one 128-byte row per pixel, TMA SWIZZLE_128B) sits in shared memory.
// The A operand of filter tap (kh, kw) is the 128 pixels (ty + kh, tx + kw), ty < 16, tx < 8: 16 groups of 8 consecutive 128-byte
// rows, group stride = (TW+2) * 128 B = 1280 B (SBO), start address = base + (kh * (TW+2) + kw) * 128 -- not 1024-aligned, so the
// descriptor's base_offset field (bits [49,52)) must carry (start >> 7) & 7 if the hardware derives the XOR pattern from the row
// index relative to the start address.  Prints, per tap and per base-offset policy, the max |error| against the CPU result.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I structure_knowledge_distillation_b200/csrc -I include \
//          tools/umma_probe.cu -o tools/umma_probe.bin -lcuda
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "sm100_ptx.cuh"

using namespace skd;

constexpr int TH = 16, TW = 8, HW2 = TW + 2, HH2 = TH + 2, ROWS = HH2 * HW2;   // 180 halo pixels
constexpr int N = 64, K = 32;

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, float* out, int kh, int kw, int policy) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* sa = smem;                                   // 180 x 128 B = 23040 B
  uint8_t* sb = smem + 23552;                           // 1024-aligned (23 * 1024)
  uint64_t* full = reinterpret_cast<uint64_t*>(sb + N * 128);
  uint64_t* done = full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { ptx::mbar_init(full, 1); ptx::mbar_init(done, 1); ptx::fence_barrier_init(); }
  if (warp == 0) ptx::tmem_alloc<64>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (threadIdx.x == 0) {
    ptx::mbar_expect_tx(full, ROWS * 128 + N * 128);
    ptx::tma_load_2d(sa, &tmap_a, full, 0, 0);
    ptx::tma_load_2d(sb, &tmap_b, full, 0, 0);
    ptx::mbar_wait(full, 0);
    ptx::tc_fence_after();
    const uint32_t idesc = ptx::make_idesc_tf32(128, N, 0, 0);
    const uint32_t a0 = ptx::smem_u32(sa) + (uint32_t)((kh * HW2 + kw) * 128);
    const uint32_t b0 = ptx::smem_u32(sb);
    for (int kk = 0; kk < K / 8; ++kk) {
      uint64_t da = ptx::make_smem_desc_sw128(a0 + kk * 32, 16, HW2 * 128);
      if (policy == 1) da |= (uint64_t)((a0 >> 7) & 7u) << 49;
      const uint64_t db = ptx::make_smem_desc_sw128(b0 + kk * 32, 16, 1024);
      ptx::mma_tf32(tmem, da, db, idesc, kk != 0 ? 1u : 0u);
    }
    ptx::mma_commit(done);
  }
  __syncthreads();
  ptx::mbar_wait(done, 0);
  ptx::tc_fence_after();
  const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
  for (int ch = 0; ch < N / 32; ++ch) {
    uint32_t v[32];
    ptx::tmem_ld_32x32(taddr + ch * 32, v);
    ptx::tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * N + ch * 32 + j] = __uint_as_float(v[j]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc<64>(tmem); }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess || !fp) { printf("no driver entry\n"); return 1; }
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fp);
  std::vector<float> ha(ROWS * K), hb(N * K);
  srand(1);
  for (auto& v : ha) v = (float)(rand() % 17 - 8);                 // small integers: exact in TF32
  for (auto& v : hb) v = (float)(rand() % 9 - 4);
  float *da, *db, *dout;
  cudaMalloc(&da, ha.size() * 4); cudaMalloc(&db, hb.size() * 4); cudaMalloc(&dout, 128 * N * 4);
  cudaMemcpy(da, ha.data(), ha.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(db, hb.data(), hb.size() * 4, cudaMemcpyHostToDevice);
  CUtensorMap ta, tb;
  {
    cuuint64_t dims[2] = {K, ROWS}; cuuint64_t strides[1] = {K * 4}; cuuint32_t box[2] = {K, ROWS}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&ta, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, da, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode A failed %d\n", (int)r); return 1; }
  }
  {
    cuuint64_t dims[2] = {K, N}; cuuint64_t strides[1] = {K * 4}; cuuint32_t box[2] = {K, N}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tb, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, db, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode B failed %d\n", (int)r); return 1; }
  }
  const int smem = 23552 + N * 128 + 1024 + 256;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> hout(128 * N);
  for (int policy = 0; policy < 2; ++policy) {
    int ok_taps = 0;
    for (int kh = 0; kh < 3; ++kh) for (int kw = 0; kw < 3; ++kw) {
      cudaMemset(dout, 0, 128 * N * 4);
      probe_kernel<<<1, 128, smem>>>(ta, tb, dout, kh, kw, policy);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("policy %d tap (%d,%d): CUDA error %s\n", policy, kh, kw, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(hout.data(), dout, hout.size() * 4, cudaMemcpyDeviceToHost);
      double maxerr = 0;
      for (int m = 0; m < 128; ++m) {
        const int ty = m / TW, tx = m % TW;
        const int row = (ty + kh) * HW2 + tx + kw;
        for (int n = 0; n < N; ++n) {
          double ref = 0;
          for (int k = 0; k < K; ++k) ref += (double)ha[row * K + k] * hb[n * K + k];
          const double d = fabs(ref - hout[m * N + n]);
          if (d > maxerr) maxerr = d;
        }
      }
      printf("policy %d (base_offset %s) tap (%d,%d) start&1023=%4d  max|err| = %g\n", policy, policy ? "=(start>>7)&7" : "=0", kh, kw,
             ((kh * HW2 + kw) * 128) & 1023, maxerr);
      if (maxerr == 0) ++ok_taps;
    }
    printf("policy %d: %d / 9 taps exact\n", policy, ok_taps);
  }
  return 0;
}
