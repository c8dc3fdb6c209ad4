// Device side of the sliding-window / multi-scale evaluation driver (reference: engine/evaluator.py:306-431, the immediate
// caller of the hot path in BASELINE config 5).  Two streaming kernels replace the host work around the model call:
//   eval_pack_crop_kernel        uint8 image -> normalised fp32 CHW network crop (float64 normalisation, zero padding of the
//                                raw image and of the normalised crop, tile cut, optional horizontal flip)
//   eval_accumulate_scale_kernel exp(logits) of all tiles of one scale -> canvas sum -> margin slice -> bilinear resize to the
//                                original size -> += fp64 accumulator over scales
// (argmax + confusion matrix: cmx_argmax_confusion in head.cu).  Both are HBM-bound byte/float movers: one thread per output
// element, consecutive threads on consecutive x, every byte read once (tile overlaps hit L2).
#include "common.cuh"
#include "../../include/cmx_b200.h"
#include <atomic>

extern std::atomic<long long> g_cmx_launches;

__global__ void __launch_bounds__(256) eval_pack_crop_kernel(const uint8_t* __restrict__ img, int rows, int cols, int ch, int pad_top,
                                                            int pad_left, int s_y, int s_x, int win_h, int win_w, int out_top,
                                                            int out_left, double m0, double m1, double m2, double d0, double d1,
                                                            double d2, int flip, float* __restrict__ out, int crop_h, int crop_w) {
  pdl_trigger();
  const long total = (long)ch * crop_h * crop_w;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int x = (int)(idx % crop_w);
  const int y = (int)((idx / crop_w) % crop_h);
  const int c = (int)(idx / ((long)crop_w * crop_h));
  const int xs = flip ? crop_w - 1 - x : x;   // the flipped crop's column x is the unflipped crop's column crop_w - 1 - x
  const int wy = y - out_top, wx = xs - out_left;
  float v = 0.f;
  if (wy >= 0 && wy < win_h && wx >= 0 && wx < win_w) {
    const int iy = s_y + wy - pad_top, ix = s_x + wx - pad_left;
    double p = 0.0;   // black border of the zero-padded raw image
    if (iy >= 0 && iy < rows && ix >= 0 && ix < cols) p = (double)img[((long)iy * cols + ix) * ch + c];
    const double mean = c == 0 ? m0 : (c == 1 ? m1 : m2), sd = c == 0 ? d0 : (c == 1 ? d1 : d2);
    v = (float)((p / 255.0 - mean) / sd);   // IEEE double arithmetic: bit-identical to the numpy expression
  }
  out[idx] = v;
}

CMX_API int cmx_eval_pack_crop(const uint8_t* img, int rows, int cols, int ch, int pad_top, int pad_left, int s_y, int s_x,
                               int win_h, int win_w, int out_top, int out_left, double mean0, double mean1, double mean2,
                               double std0, double std1, double std2, int flip, float* out, int crop_h, int crop_w, void* stream) {
  CMX_REQUIRE(ch == 1 || ch == 3, "eval_pack_crop: ch=%d (1 or 3)", ch);
  CMX_REQUIRE(rows > 0 && cols > 0 && crop_h > 0 && crop_w > 0, "eval_pack_crop: empty image or crop");
  const long total = (long)ch * crop_h * crop_w;
  eval_pack_crop_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(img, rows, cols, ch, pad_top, pad_left, s_y, s_x, win_h,
                                                                            win_w, out_top, out_left, mean0, mean1, mean2, std0,
                                                                            std1, std2, flip, out, crop_h, crop_w);
  g_cmx_launches++;
  CMX_CHECK_LAUNCH("eval_pack_crop");
  return 0;
}

// score of the scale at (c, y, x) of its rows x cols map
__device__ __forceinline__ float eval_score_at(const float* __restrict__ logits, const float* __restrict__ lflip, int crop_h, int crop_w,
                                               int ncls, const CmxEvalTile* __restrict__ tiles, int n_tiles, int m_top, int m_left,
                                               int c, int y, int x) {
  const int cy = y + m_top, cx = x + m_left;
  float acc = 0.f;
  for (int t = 0; t < n_tiles; t++) {
    const CmxEvalTile tl = tiles[t];
    if (cy < tl.s_y || cy >= tl.e_y || cx < tl.s_x || cx >= tl.e_x) continue;
    const int ty = cy - tl.s_y + tl.tm_top, tx = cx - tl.s_x + tl.tm_left;
    const long base = ((long)tl.crop * ncls + c) * crop_h + ty;
    float l = logits[base * crop_w + tx];
    if (lflip) l = __fadd_rn(l, lflip[base * crop_w + (crop_w - 1 - tx)]);
    acc = __fadd_rn(acc, expf(l));
  }
  return acc;
}

__global__ void __launch_bounds__(256) eval_accumulate_scale_kernel(const float* __restrict__ logits, const float* __restrict__ lflip,
                                                                   int crop_h, int crop_w, int ncls,
                                                                   const CmxEvalTile* __restrict__ tiles, int n_tiles, int m_top,
                                                                   int m_left, int rows, int cols, double* __restrict__ processed,
                                                                   int ori_rows, int ori_cols, double sy, double sx) {
  pdl_trigger();
  const long total = (long)ncls * ori_rows * ori_cols;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int ox = (int)(idx % ori_cols);
  const int oy = (int)((idx / ori_cols) % ori_rows);
  const int c = (int)(idx / ((long)ori_cols * ori_rows));
  float out;
  if (rows == ori_rows && cols == ori_cols) {
    out = eval_score_at(logits, lflip, crop_h, crop_w, ncls, tiles, n_tiles, m_top, m_left, c, oy, ox);
  } else {
    // cv2.resize INTER_LINEAR, float path: src = (float)((dst + 0.5) * scale - 0.5) with scale in double; floor; clamp with
    // zero weight at the borders
    float fy = (float)(((double)oy + 0.5) * sy - 0.5), fx = (float)(((double)ox + 0.5) * sx - 0.5);
    int y0 = (int)floorf(fy), x0 = (int)floorf(fx);
    float wy = fy - (float)y0, wx = fx - (float)x0;
    if (y0 < 0) { y0 = 0; wy = 0.f; }
    if (y0 >= rows - 1) { y0 = rows - 1; wy = 0.f; }
    if (x0 < 0) { x0 = 0; wx = 0.f; }
    if (x0 >= cols - 1) { x0 = cols - 1; wx = 0.f; }
    const int y1 = y0 + (y0 < rows - 1 ? 1 : 0), x1 = x0 + (x0 < cols - 1 ? 1 : 0);
    const float v00 = eval_score_at(logits, lflip, crop_h, crop_w, ncls, tiles, n_tiles, m_top, m_left, c, y0, x0);
    const float v01 = eval_score_at(logits, lflip, crop_h, crop_w, ncls, tiles, n_tiles, m_top, m_left, c, y0, x1);
    const float v10 = eval_score_at(logits, lflip, crop_h, crop_w, ncls, tiles, n_tiles, m_top, m_left, c, y1, x0);
    const float v11 = eval_score_at(logits, lflip, crop_h, crop_w, ncls, tiles, n_tiles, m_top, m_left, c, y1, x1);
    // horizontal pass on the two rows, then the vertical blend (separate fp32 roundings, no contraction)
    const float h0 = __fadd_rn(__fmul_rn(v00, 1.f - wx), __fmul_rn(v01, wx));
    const float h1 = __fadd_rn(__fmul_rn(v10, 1.f - wx), __fmul_rn(v11, wx));
    out = __fadd_rn(__fmul_rn(h0, 1.f - wy), __fmul_rn(h1, wy));
  }
  processed[idx] += (double)out;
}

CMX_API int cmx_eval_accumulate_scale(const float* logits, const float* logits_flip, int crop_h, int crop_w, int ncls,
                                      const CmxEvalTile* tiles, int n_tiles, int m_top, int m_left, int rows, int cols,
                                      double* processed, int ori_rows, int ori_cols, void* stream) {
  CMX_REQUIRE(n_tiles >= 1 && rows > 0 && cols > 0 && ori_rows > 0 && ori_cols > 0 && ncls > 0, "eval_accumulate_scale: bad sizes");
  const long total = (long)ncls * ori_rows * ori_cols;
  const double sy = (double)rows / (double)ori_rows, sx = (double)cols / (double)ori_cols;
  eval_accumulate_scale_kernel<<<cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(logits, logits_flip, crop_h, crop_w, ncls, tiles,
                                                                                   n_tiles, m_top, m_left, rows, cols, processed,
                                                                                   ori_rows, ori_cols, sy, sx);
  g_cmx_launches++;
  CMX_CHECK_LAUNCH("eval_accumulate_scale");
  return 0;
}
