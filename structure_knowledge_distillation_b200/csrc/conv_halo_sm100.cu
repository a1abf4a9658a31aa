// tcgen05 3x3 / stride-1 / pad-1 convolution for SMALL channel counts (Cin, Cout <= 128): the student's stem and layer1, their
// data gradients, and the teacher's 64/128-wide 3x3 layers (networks/pspnet_combine.py:118-132, 33-45).
//
// Why a second kernel: the general implicit-GEMM kernel (conv_sm100.cu) fetches one 128-pixel x 32-channel A box per (tap, chunk)
// K step -- every activation line travels L2 -> shared memory NINE times.  With Cin = 64 the tensor core needs only ~1.9 us per
// 128 x 64 tile but the 18 A boxes (288 KB) plus 18 weight boxes take ~6.7 us: those layers ran at 120-350 TFLOP/s, L2-bound.
//
// Here the input HALO tile of a 16 x 8 pixel output tile -- 18 x 10 pixels x 32 channels, one 128-byte swizzled row per pixel --
// is loaded ONCE per channel chunk (one TMA box, 23 KB) and the nine filter taps read it through SHIFTED shared-memory matrix
// descriptors: tap (kh, kw) is the operand that starts (kh * 10 + kw) * 128 bytes into the halo tile, its 16 groups of 8 pixel rows
// 1280 bytes (one halo row) apart (SBO).  The 128B-swizzle XOR is a function of the absolute shared-memory address, so a start
// address that is not 1024-byte aligned needs nothing else (tools/umma_probe.cu checks all nine shifts bit-exactly on the GPU).
// A traffic per tile drops from 288 KB to 46 KB (Cin = 64); weights still stream per (tap, chunk) through a TMA ring.
//
// Warp roles as in conv_sm100.cu: warp 0 TMA producer, warp 1 MMA issuer (kind::tf32, M = 128, N = BLOCK_N, K = 8; optional
// split-precision 3xTF32 with lo-part tiles), warps 2-9 epilogue (tcgen05.ld -> scale/shift/activation -> swizzled staging ->
// TMA store of {32 ch, 8 px, 16 rows} boxes, hardware-clipped at the image edge).  K loop: channel chunk outermost, taps inside,
// so a halo slot is released after its nine taps and the next tile's chunk can land while this tile's later chunks compute.
//
// Weights: every CTA needs the same 9 * Cin/32 weight boxes for every tile.  With one CTA per SM fetching them itself, 148 SMs hammer
// the same few L2 lines (ncu: 376 us for 8 x 64 x 256 x 512 -> 64 with the A traffic already cut to 1/6: each 8 KB weight box took
// 0.38 us to arrive -- hot-line latency, not bandwidth).  So the CTAs form CLUSTERS of 4: the leader issues each weight box once with
// TMA multicast into the shared memory of all four, every CTA's MMA warp releases the stage on the leader's barrier
// (tcgen05.commit ... multicast::cluster).  The CTAs of a cluster walk the same number of tiles (a ragged last round runs phantom
// tiles on zero-filled halos).  MEASURED: slower than every CTA fetching for itself (the leader can refill a stage only when all
// four CTAs have released it: lock-step) -- kept as an option (skd_set_conv_halo bit 2), off by default.
//
// What does work where it fits (Cin, Cout <= 64, TF32: 144 KB of weights): WRES -- the whole filter bank is loaded ONCE per persistent
// CTA and stays in shared memory; only halo tiles stream.  The weight ring was latency-bound, not bandwidth-bound: 8 stages x 8 KB in
// flight against ~2.5 us of loaded L2 latency is ~25 GB/s per SM where the tensor core wants ~80.
#include <cuda.h>

#include "common.cuh"
#include "skd.h"
#include "sm100_ptx.cuh"

using namespace skd;

namespace {

constexpr int kTH = 16, kTW = 8;                       // output tile: 16 rows x 8 pixels = 128 GEMM rows
constexpr int kHW = kTW + 2, kHH = kTH + 2;            // halo tile 18 x 10
constexpr int kHaloBytes = kHH * kHW * 128;            // 23 040
constexpr int kSlotBytes = 23 * 1024;                  // padded to a multiple of 1024 (swizzle atom alignment of the next slot)
constexpr int kThreads = 320;
constexpr int kMaxSlots = 8, kMaxStages = 8;
constexpr int kOutStageBytes = 128 * 32 * 4;

struct HaloArgs {
  int N, H, W, Cin, Cout;
  int tiles_x, tiles_y, m_tiles, k_chunks;
  int passes;                      // 1: TF32; 3: split precision (hi/lo tiles of both operands)
  int nslots, nstages;             // halo slots (one channel chunk each) and weight stages in use
  const float* scale; const float* shift; int act; float slope; int round_out;
};

template <int BLOCK_N, int CL, bool WRES>
__global__ void __launch_bounds__(kThreads, 1)
conv3x3_halo_sm100_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                          const __grid_constant__ CUtensorMap tmap_y, const __grid_constant__ CUtensorMap tmap_x2,
                          const __grid_constant__ CUtensorMap tmap_w2, const HaloArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw_addr = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  const bool split3 = (a.passes == 3);
  const int parts = split3 ? 2 : 1;
  const int slot_stride = kSlotBytes * parts;                     // {hi, lo} halo tiles of one chunk
  constexpr int kBBytes = BLOCK_N * 128;
  const int stage_stride = kBBytes * parts;
  uint8_t* sA = smem;
  uint8_t* sB = sA + a.nslots * slot_stride;
  uint8_t* out_stage = sB + a.nstages * stage_stride;              // 2 x 16 KB
  uint64_t* a_full = reinterpret_cast<uint64_t*>(out_stage + 2 * kOutStageBytes);
  uint64_t* a_empty = a_full + kMaxSlots;
  uint64_t* b_full = a_empty + kMaxSlots;
  uint64_t* b_empty = b_full + kMaxStages;
  uint64_t* b_empty_all = b_empty + kMaxStages;                    // leader's: one arrival per CTA of the cluster
  uint64_t* tmem_full = b_empty_all + kMaxStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* s_scale = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(a_full) + 512);
  const uint32_t cta_rank = CL > 1 ? ptx::cluster_ctarank() : 0u;
  const int iters = (a.m_tiles + (int)gridDim.x - 1) / (int)gridDim.x;      // the same for every CTA: ragged rounds run phantom tiles
  float* s_shift = s_scale + BLOCK_N;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_x); ptx::prefetch_tmap(&tmap_w); ptx::prefetch_tmap(&tmap_y);
    for (int s = 0; s < kMaxSlots; ++s) { ptx::mbar_init(&a_full[s], 1); ptx::mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < kMaxStages; ++s) { ptx::mbar_init(&b_full[s], 1); ptx::mbar_init(&b_empty[s], 1); ptx::mbar_init(&b_empty_all[s], CL); }
    for (int s = 0; s < 2; ++s) { ptx::mbar_init(&tmem_full[s], 1); ptx::mbar_init(&tmem_empty[s], 8); }
    ptx::fence_barrier_init();
  }
  // TMEM: 2 accumulator stages x kSub sub-accumulators x BLOCK_N columns = all 512.  Split precision spreads the filter taps of a tile
  // over the kSub sub-accumulators and adds them in the epilogue with round-to-nearest fp32 adds: the tensor core TRUNCATES when it adds
  // into its accumulator (~2^-24 relative, biased, per MMA), and with 216 MMAs per tile that truncation -- not the operand split -- was
  // what was left of the error of the "fp32-grade" layers (~1e-5), which train-mode BN then amplifies ~50x into the pair-wise loss.
  constexpr int kSub = 256 / BLOCK_N;                  // 4 (BLOCK_N 64) or 2 (128)
  constexpr int kTmemCols = 512;
  const int nsub = split3 ? kSub : 1;
  if (warp == 1) ptx::tmem_alloc<kTmemCols>(tmem_base_slot);
  ptx::tc_fence_before();
  if constexpr (CL > 1) ptx::cluster_sync(); else __syncthreads();     // peers' barriers exist before anything is multicast to them
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  const int tiles_per_img = a.tiles_x * a.tiles_y;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int slot = 0; uint32_t sphase = 0; int stage = 0; uint32_t bphase = 0;
      if constexpr (WRES) {                                            // the whole filter bank, once: box (kc, tap) at index kc * 9 + tap
        ptx::mbar_expect_tx(&b_full[0], (uint32_t)(9 * a.k_chunks * stage_stride));
        for (int kc = 0; kc < a.k_chunks; ++kc)
          for (int tap = 0; tap < 9; ++tap) {
            uint8_t* pb = sB + (kc * 9 + tap) * stage_stride;
            ptx::tma_load_3d(pb, &tmap_w, &b_full[0], kc * 32, tap, 0);
            if (split3) ptx::tma_load_3d(pb + kBBytes, &tmap_w2, &b_full[0], kc * 32, tap, 0);
          }
      }
      for (int it = 0; it < iters; ++it) {
        const int tile = blockIdx.x + it * gridDim.x;
        // phantom tile (ragged last round of a cluster): image index N is out of range -> the halo is all hardware zero fill
        const int img = tile < a.m_tiles ? tile / tiles_per_img : a.N, rem = tile < a.m_tiles ? tile - img * tiles_per_img : 0;
        const int ty = rem / a.tiles_x, tx = rem - ty * a.tiles_x;
        const int x0 = tx * kTW - 1, y0 = ty * kTH - 1;                 // halo origin (may be -1: hardware zero fill = padding)
        for (int kc = 0; kc < a.k_chunks; ++kc) {
          ptx::mbar_wait(&a_empty[slot], sphase ^ 1);
          uint8_t* pa = sA + slot * slot_stride;
          ptx::mbar_expect_tx(&a_full[slot], (uint32_t)(kHaloBytes * parts));
          ptx::tma_load_4d(pa, &tmap_x, &a_full[slot], kc * 32, x0, y0, img);
          if (split3) ptx::tma_load_4d(pa + kSlotBytes, &tmap_x2, &a_full[slot], kc * 32, x0, y0, img);
          if (++slot == a.nslots) { slot = 0; sphase ^= 1; }
          for (int tap = 0; tap < (WRES ? 0 : 9); ++tap) {
            ptx::mbar_wait(&b_empty[stage], bphase ^ 1);                  // this CTA is done with the stage: arm its barrier
            uint8_t* pb = sB + stage * stage_stride;
            ptx::mbar_expect_tx(&b_full[stage], (uint32_t)stage_stride);
            if constexpr (CL > 1) {
              if (cta_rank == 0) {                                          // the leader fetches the box once for the whole cluster
                ptx::mbar_wait(&b_empty_all[stage], bphase ^ 1);            // every CTA's MMAs have released the stage
                constexpr uint16_t mask = (uint16_t)((1u << CL) - 1u);
                ptx::tma_load_3d_mcast(pb, &tmap_w, &b_full[stage], kc * 32, tap, 0, mask);
                if (split3) ptx::tma_load_3d_mcast(pb + kBBytes, &tmap_w2, &b_full[stage], kc * 32, tap, 0, mask);
              }
            } else {
              ptx::tma_load_3d(pb, &tmap_w, &b_full[stage], kc * 32, tap, 0);
              if (split3) ptx::tma_load_3d(pb + kBBytes, &tmap_w2, &b_full[stage], kc * 32, tap, 0);
            }
            if (++stage == a.nstages) { stage = 0; bphase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = ptx::make_idesc_tf32(128, BLOCK_N, 0, 0);
    int slot = 0; uint32_t sphase = 0; int stage = 0; uint32_t bphase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    if constexpr (WRES) {
      if (lane == 0) { ptx::mbar_wait(&b_full[0], 0); ptx::tc_fence_after(); }       // resident weights have landed
      __syncwarp();
    }
    for (int it = 0; it < iters; ++it) {
      if (lane == 0) ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      __syncwarp();
      ptx::tc_fence_after();
      const uint32_t tmem_acc = tmem_base + (uint32_t)(acc * kSub * BLOCK_N);
      uint32_t used = 0;                                                            // sub-accumulators that already hold a partial sum
      for (int kc = 0; kc < a.k_chunks; ++kc) {
        if (lane == 0) {
          ptx::mbar_wait(&a_full[slot], sphase);
          ptx::tc_fence_after();
        }
        __syncwarp();
        const uint32_t sa0 = ptx::smem_u32(sA + slot * slot_stride);
        for (int tap = 0; tap < 9; ++tap) {
          if (lane == 0) {
            if constexpr (!WRES) {
              ptx::mbar_wait(&b_full[stage], bphase);
              ptx::tc_fence_after();
            }
            const int kh = tap / 3, kw = tap - kh * 3;
            const uint32_t sa = sa0 + (uint32_t)((kh * kHW + kw) * 128);          // shifted view of the halo tile
            const uint32_t sb = ptx::smem_u32(sB + (WRES ? kc * 9 + tap : stage) * stage_stride);
            const int sub = tap % nsub;
            const uint32_t tmem_d = tmem_acc + (uint32_t)(sub * BLOCK_N);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint64_t da = ptx::make_smem_desc_sw128(sa + kk * 32, 16, kHW * 128);   // 8-pixel groups one halo row apart
              const uint64_t db = ptx::make_smem_desc_sw128(sb + kk * 32, 16, 1024);
              ptx::mma_tf32(tmem_d, da, db, idesc, ((used >> sub) & 1u) | (kk != 0 ? 1u : 0u));
              if (split3) {
                ptx::mma_tf32(tmem_d, ptx::make_smem_desc_sw128(sa + kSlotBytes + kk * 32, 16, kHW * 128), db, idesc, 1u);
                ptx::mma_tf32(tmem_d, da, ptx::make_smem_desc_sw128(sb + kBBytes + kk * 32, 16, 1024), idesc, 1u);
              }
            }
            used |= 1u << sub;
            if constexpr (!WRES) ptx::mma_commit(&b_empty[stage]);
            if constexpr (CL > 1) ptx::mma_commit_mcast(&b_empty_all[stage], (uint16_t)1);   // ... and tell the leader
            if (tap == 8) {
              ptx::mma_commit(&a_empty[slot]);                                     // the halo slot is free after its nine taps
              if (kc == a.k_chunks - 1) ptx::mma_commit(&tmem_full[acc]);
            }
          }
          __syncwarp();
          if (++stage == a.nstages) { stage = 0; bphase ^= 1; }
        }
        if (++slot == a.nslots) { slot = 0; sphase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ===================== epilogue: two groups of four warps, alternate 32-column chunks =====================
    const int ew = warp - 2;
    const int quarter = warp & 3;
    const int group = ew >> 2;
    const int row = quarter * 32 + lane;                                           // accumulator row = tile pixel (ty * 8 + tx)
    const int etid = ew * 32 + lane;
    const bool is_store_leader = ((ew & 3) == 0 && lane == 0);
    uint8_t* stg = out_stage + group * kOutStageBytes;
    int acc = 0; uint32_t acc_phase = 0;
    // per-channel affine (one N tile: loaded once)
    if (etid < BLOCK_N) {
      s_scale[etid] = (etid < a.Cout) ? (a.scale ? __ldg(a.scale + etid) : 1.f) : 0.f;
      s_shift[etid] = (etid < a.Cout && a.shift) ? __ldg(a.shift + etid) : 0.f;
    }
    ptx::named_bar_sync(3, 256);
    for (int it = 0; it < iters; ++it) {
      const int tile = blockIdx.x + it * gridDim.x;
      const bool real = tile < a.m_tiles;
      const int img = real ? tile / tiles_per_img : 0, rem = real ? tile - img * tiles_per_img : 0;
      const int ty = rem / a.tiles_x, tx = rem - ty * a.tiles_x;
      ptx::mbar_wait(&tmem_full[acc], acc_phase);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * kSub * BLOCK_N);
#pragma unroll 1
      for (int ch = group; ch < BLOCK_N / 32; ch += 2) {
        const int c0 = ch * 32;
        if (c0 >= a.Cout || !real) break;                                          // phantom tile: nothing to store
        uint32_t r[32];
        ptx::tmem_ld_32x32(taddr + ch * 32, r);
        ptx::tmem_ld_wait();
        for (int sb2 = 1; sb2 < nsub; ++sb2) {                                     // split precision: add the other sub-accumulators (RN fp32)
          uint32_t r2[32];
          ptx::tmem_ld_32x32(taddr + sb2 * BLOCK_N + ch * 32, r2);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
        }
        float v[32];
        const float4* sc4 = reinterpret_cast<const float4*>(s_scale + c0);
        const float4* sh4 = reinterpret_cast<const float4*>(s_shift + c0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 sc = sc4[j], sh = sh4[j];
          v[4 * j + 0] = fmaf(__uint_as_float(r[4 * j + 0]), sc.x, sh.x);
          v[4 * j + 1] = fmaf(__uint_as_float(r[4 * j + 1]), sc.y, sh.y);
          v[4 * j + 2] = fmaf(__uint_as_float(r[4 * j + 2]), sc.z, sh.z);
          v[4 * j + 3] = fmaf(__uint_as_float(r[4 * j + 3]), sc.w, sh.w);
        }
        if (a.act == ACT_RELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        } else if (a.act == ACT_LEAKY) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = v[j] < 0.f ? v[j] * a.slope : v[j];
        }
        if (a.round_out) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = ptx::round_tf32(v[j]);
        }
        if (is_store_leader) ptx::bulk_wait_read<0>();                             // the group's previous store has read the buffer
        ptx::named_bar_sync(1 + group, 128);
        uint8_t* srow = stg + row * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(srow + ((j ^ (row & 7)) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        ptx::fence_proxy_async();
        ptx::named_bar_sync(1 + group, 128);
        if (is_store_leader) {
          ptx::tma_store_4d(&tmap_y, stg, c0, tx * kTW, ty * kTH, img);           // {32 ch, 8 px, 16 rows}: clipped at the image edge
          ptx::bulk_commit();
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (is_store_leader) ptx::bulk_wait<0>();
  }

  ptx::tc_fence_before();
  if constexpr (CL > 1) ptx::cluster_sync(); else __syncthreads();     // peers may still multicast into / signal this CTA's shared memory
  if (warp == 1) { ptx::tc_fence_after(); ptx::tmem_dealloc<kTmemCols>(tmem_base); }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr; static bool tried = false;
  if (!tried) {
    tried = true; void* p = nullptr; cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

bool encode(CUtensorMap* m, int rank, const void* base, const cuuint64_t* dims, const cuuint64_t* strides, const cuuint32_t* box, bool tf32) {
  cuuint32_t es[4] = {1, 1, 1, 1};
  return get_encode()(m, tf32 ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), dims,
                      strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct HaloPlan { int bn, k_chunks, nslots, nstages, smem, wres; };
int g_halo_wres = 1;

// shared-memory budget: halo slots (one per channel chunk, x2 for split precision) + weight stages + 2 staging tiles + barriers
bool make_plan(int Cin, int Cout, int passes, HaloPlan* p) {
  if (Cin % 32 || Cin < 32 || Cin > 128 || Cout % 4 || Cout < 4 || Cout > 128) return false;
  p->bn = Cout > 64 ? 128 : 64;
  p->k_chunks = Cin / 32;
  const int parts = passes == 3 ? 2 : 1;
  const int slot = kSlotBytes * parts, stage = p->bn * 128 * parts;
  const int fixed = 2 * kOutStageBytes + 1024 /*align*/ + 512 /*barriers*/ + 2 * p->bn * 4 + 512;
  const int avail = 232448 - fixed;
  p->wres = 0;
  if (g_halo_wres && avail - 9 * p->k_chunks * stage >= p->k_chunks * slot) {     // the filter bank fits beside one halo tile: keep it resident
    int nslots = (avail - 9 * p->k_chunks * stage) / slot;
    if (nslots > 2 * p->k_chunks) nslots = 2 * p->k_chunks;
    if (nslots > kMaxSlots) nslots = kMaxSlots;
    p->wres = 1; p->nslots = nslots; p->nstages = 9 * p->k_chunks;
    p->smem = nslots * slot + p->nstages * stage + fixed;
    return true;
  }
  int nstages = 4;
  if (avail - nstages * stage < p->k_chunks * slot) nstages = 3;
  int nslots = (avail - nstages * stage) / slot;
  if (nslots < p->k_chunks) return false;
  if (nslots > 2 * p->k_chunks) nslots = 2 * p->k_chunks;
  if (nslots > kMaxSlots) nslots = kMaxSlots;
  nstages = (avail - nslots * slot) / stage;
  if (nstages > kMaxStages) nstages = kMaxStages;
  p->nslots = nslots; p->nstages = nstages;
  p->smem = nslots * slot + nstages * stage + fixed;
  return nstages >= 3;
}

template <int BLOCK_N, int CL, bool WRES>
int launch(const CUtensorMap& tx, const CUtensorMap& tw, const CUtensorMap& ty, const CUtensorMap& tx2, const CUtensorMap& tw2, const HaloArgs& a,
           int smem, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_halo_sm100_kernel<BLOCK_N, CL, WRES>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
    if (e != cudaSuccess) { set_error("skd_conv3x3_halo_sm100(attr)", e); return 0; }
    attr = true;
  }
  int grid = a.m_tiles < kNumSMs ? a.m_tiles : kNumSMs;
  if (CL > 1) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((kNumSMs / CL) * CL); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    // persistent grid = the clusters that can be resident at once (a cluster lives inside one GPC: with 18-SM GPCs not all 37
    // clusters of 4 fit; the left-over would start as a second wave after the first finished its whole tile loop)
    static int max_clusters = 0;
    if (max_clusters == 0) {
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, conv3x3_halo_sm100_kernel<BLOCK_N, CL, WRES>, &cfg) != cudaSuccess || n <= 0) { cudaGetLastError(); n = kNumSMs / CL; }
      max_clusters = n;
    }
    grid = max_clusters * CL;
    if (grid > (kNumSMs / CL) * CL) grid = (kNumSMs / CL) * CL;
    cfg.gridDim = dim3(grid);
    cudaError_t e = cudaLaunchKernelEx(&cfg, conv3x3_halo_sm100_kernel<BLOCK_N, CL, WRES>, tx, tw, ty, tx2, tw2, a);
    if (e != cudaSuccess) { set_error("skd_conv3x3_halo_sm100(cluster launch)", e); return 0; }
  } else {
    conv3x3_halo_sm100_kernel<BLOCK_N, CL, WRES><<<grid, kThreads, smem, st>>>(tx, tw, ty, tx2, tw2, a);
  }
  return finish("skd_conv3x3_halo_sm100");
}

}  // namespace

namespace skd {
int g_halo_cluster = 0;            // weight multicast across clusters of 4: measured SLOWER (lock-step release through the leader); opt-in via skd_set_conv_halo(5)
int g_conv_halo = 1;   // skd_set_conv_halo(0): every shape through the general implicit-GEMM kernel

bool conv3x3_halo_supported(int Cin, int Cout, int passes) {
  HaloPlan p;
  return g_conv_halo && get_encode() && make_plan(Cin, Cout, passes, &p);
}

// y = act(scale * conv3x3(x, w) + shift), stride 1, pad 1, dilation 1.  x / y dense NHWC with pitches ldx / ldy; w [Cout][3][3][Cin].
int conv3x3_halo_launch(int N, int H, int W, int Cin, int Cout, const float* x, const float* x_lo, int ldx, const float* w, const float* w_lo,
                        float* y, int ldy, const float* scale, const float* shift, int act, float slope, int round_tf32, cudaStream_t st) {
  const char* who = "skd_conv3x3_halo_sm100";
  HaloPlan p;
  const int passes = (x_lo && w_lo) ? 3 : 1;
  if (!make_plan(Cin, Cout, passes, &p)) { set_error_msg(who, "unsupported channel counts for the halo kernel"); return 0; }
  if (ldx % 4 || ldy % 4 || ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(y)) & 15)) {
    set_error_msg(who, "pitches must be multiples of 4 floats and pointers 16-byte aligned (TMA)"); return 0;
  }
  if (N <= 0 || H <= 0 || W <= 0) return 1;
  CUtensorMap tx, tw, ty, tx2, tw2;
  {
    cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)ldx * 4, (cuuint64_t)W * ldx * 4, (cuuint64_t)H * W * ldx * 4};
    cuuint32_t box[4] = {32, (cuuint32_t)kHW, (cuuint32_t)kHH, 1};
    if (!encode(&tx, 4, x, dims, strides, box, g_tf32_tma_type != 0)) { set_error_msg(who, "cuTensorMapEncodeTiled(x) failed"); return 0; }
    tx2 = tx;
    if (passes == 3 && !encode(&tx2, 4, x_lo, dims, strides, box, g_tf32_tma_type != 0)) { set_error_msg(who, "cuTensorMapEncodeTiled(x_lo) failed"); return 0; }
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)Cin, 9, (cuuint64_t)Cout};
    cuuint64_t strides[2] = {(cuuint64_t)Cin * 4, (cuuint64_t)9 * Cin * 4};
    cuuint32_t box[3] = {32, 1, (cuuint32_t)p.bn};
    if (!encode(&tw, 3, w, dims, strides, box, g_tf32_tma_type != 0)) { set_error_msg(who, "cuTensorMapEncodeTiled(w) failed"); return 0; }
    tw2 = tw;
    if (passes == 3 && !encode(&tw2, 3, w_lo, dims, strides, box, g_tf32_tma_type != 0)) { set_error_msg(who, "cuTensorMapEncodeTiled(w_lo) failed"); return 0; }
  }
  {
    cuuint64_t dims[4] = {(cuuint64_t)Cout, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)ldy * 4, (cuuint64_t)W * ldy * 4, (cuuint64_t)H * W * ldy * 4};
    cuuint32_t box[4] = {32, (cuuint32_t)kTW, (cuuint32_t)kTH, 1};
    if (!encode(&ty, 4, y, dims, strides, box, false)) { set_error_msg(who, "cuTensorMapEncodeTiled(y) failed"); return 0; }
  }
  HaloArgs a;
  a.N = N; a.H = H; a.W = W; a.Cin = Cin; a.Cout = Cout;
  a.tiles_x = (W + kTW - 1) / kTW; a.tiles_y = (H + kTH - 1) / kTH; a.m_tiles = N * a.tiles_x * a.tiles_y; a.k_chunks = p.k_chunks;
  a.passes = passes; a.nslots = p.nslots; a.nstages = p.nstages;
  a.scale = scale; a.shift = shift; a.act = act; a.slope = slope; a.round_out = round_tf32;
  // clusters of 4 share every weight box through TMA multicast; a handful of tiles is not worth a cluster
  if (p.wres) return p.bn == 128 ? launch<128, 1, true>(tx, tw, ty, tx2, tw2, a, p.smem, st) : launch<64, 1, true>(tx, tw, ty, tx2, tw2, a, p.smem, st);
  const bool mc = g_halo_cluster && a.m_tiles >= 2 * kNumSMs;
  if (p.bn == 128) return mc ? launch<128, 4, false>(tx, tw, ty, tx2, tw2, a, p.smem, st) : launch<128, 1, false>(tx, tw, ty, tx2, tw2, a, p.smem, st);
  return mc ? launch<64, 4, false>(tx, tw, ty, tx2, tw2, a, p.smem, st) : launch<64, 1, false>(tx, tw, ty, tx2, tw2, a, p.smem, st);
}
}  // namespace skd

// bit 0: halo kernel on; bit 1: NO resident weights (always the streaming ring); bit 2: weight multicast over clusters of 4
extern "C" void skd_set_conv_halo(int on) { skd::g_conv_halo = (on & 1) ? 1 : 0; g_halo_wres = (on & 2) ? 0 : 1; skd::g_halo_cluster = (on & 4) ? 1 : 0; }

extern "C" int skd_conv3x3_halo_sm100(int N, int H, int W, int Cin, int Cout, const float* x, const float* x_lo, int ldx, const float* w,
                                      const float* w_lo, float* y, int ldy, const float* scale, const float* shift, int act, float slope,
                                      cudaStream_t st) {
  if ((x_lo == nullptr) != (w_lo == nullptr)) { set_error_msg("skd_conv3x3_halo_sm100", "x_lo and w_lo must be given together"); return 0; }
  return skd::conv3x3_halo_launch(N, H, W, Cin, Cout, x, x_lo, ldx, w, w_lo, y, ldy, scale, shift, act, slope, 0, st);
}
