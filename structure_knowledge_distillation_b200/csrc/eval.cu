// Evaluation kernels (networks/evaluate.py:75-206): bilinear (align_corners) up-sampling of the class scores of one tile
// accumulated into the full-image score buffer (predict_sliding :86-113 / predict_whole :115-122), then arg-max and the confusion
// matrix (get_confusion_matrix :143-160, ignore label 255) -- no score tensor ever goes to the host.
#include "common.cuh"
#include "skd.h"

using namespace skd;

namespace {

// full[(y1+y)*W + x1+x][c] += bilinear(logits)[c][y][x] for y < vh, x < vw (the tile may be padded beyond the image: only the valid
// part is accumulated, evaluate.py:106-110)
__global__ void __launch_bounds__(256)
upsample_acc_kernel(int C, int h, int w, const float* __restrict__ L, long long sc, long long sp, int TH, int TW, int vh, int vw,
                    float* __restrict__ full, int W, int y1, int x1) {
  const long long total = (long long)vh * vw * C;
  const float ry = TH > 1 ? (float)(h - 1) / (float)(TH - 1) : 0.f, rx = TW > 1 ? (float)(w - 1) / (float)(TW - 1) : 0.f;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int c = (int)(i % C); const long long r = i / C;
    const int x = (int)(r % vw), y = (int)(r / vw);
    const float fy = y * ry, fx = x * rx;
    const int y0 = min((int)fy, h - 1), x0 = min((int)fx, w - 1);
    const int y1i = min(y0 + 1, h - 1), x1i = min(x0 + 1, w - 1);
    const float ly = fy - y0, lx = fx - x0;
    const float* p = L + c * sc;
    const float v00 = __ldg(p + ((long long)y0 * w + x0) * sp), v01 = __ldg(p + ((long long)y0 * w + x1i) * sp);
    const float v10 = __ldg(p + ((long long)y1i * w + x0) * sp), v11 = __ldg(p + ((long long)y1i * w + x1i) * sp);
    const float v = (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
    full[((long long)(y1 + y) * W + x1 + x) * C + c] += v;
  }
}

__global__ void __launch_bounds__(256)
argmax_confusion_kernel(int H, int W, int C, const float* __restrict__ full, const long long* __restrict__ gt, long long gt_row, int vh,
                        int vw, int ignore, unsigned long long* conf, unsigned char* __restrict__ pred) {
  const long long total = (long long)vh * vw;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int x = (int)(i % vw), y = (int)(i / vw);
    const float* p = full + ((long long)y * W + x) * C;
    int best = 0; float bv = p[0];
    for (int c = 1; c < C; ++c) { const float v = p[c]; if (v > bv) { bv = v; best = c; } }     // first maximum, like np.argmax
    if (pred) pred[(long long)y * W + x] = (unsigned char)best;
    if (gt) {
      const long long g = gt[(long long)y * gt_row + x];
      if (g != ignore && g >= 0 && g < C) atomicAdd(conf + g * C + best, 1ull);
    }
  }
}

int blocks_for(long long n) { long long b = (n + 255) / 256; if (b > kNumSMs * 8) b = kNumSMs * 8; if (b < 1) b = 1; return (int)b; }

}  // namespace

extern "C" int skd_eval_upsample_accumulate(int C, int h, int w, const float* logits, long long sc, long long sp, int tile_h, int tile_w,
                                            int valid_h, int valid_w, float* full, int W, int y1, int x1, cudaStream_t st) {
  if (valid_h <= 0 || valid_w <= 0) return 1;
  upsample_acc_kernel<<<blocks_for((long long)valid_h * valid_w * C), 256, 0, st>>>(C, h, w, logits, sc, sp, tile_h, tile_w, valid_h, valid_w, full,
                                                                                   W, y1, x1);
  return finish("skd_eval_upsample_accumulate");
}

extern "C" int skd_eval_argmax_confusion(int H, int W, int C, const float* full, const long long* gt, long long gt_row, int valid_h,
                                         int valid_w, int ignore_index, long long* confusion, unsigned char* pred, cudaStream_t st) {
  if (valid_h <= 0 || valid_w <= 0) return 1;
  argmax_confusion_kernel<<<blocks_for((long long)valid_h * valid_w), 256, 0, st>>>(H, W, C, full, gt, gt_row, valid_h, valid_w, ignore_index,
                                                                                   reinterpret_cast<unsigned long long*>(confusion), pred);
  return finish("skd_eval_argmax_confusion");
}
