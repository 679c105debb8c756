// inflate.cu -- K2: map_inflate (replaces map_inflate_node::MapCallback / InflateMap /
// CreateInflatedOccupancyGrid, MI/src/map_inflate_node.cpp:65-161).
//
// Reference pipeline: occupancy -> gray = 255 - round(occ*255/100) (unknown -1 -> 100), vertical flip ->
// cv::threshold(OTSU) -> invert -> cv::dilate(MORPH_ELLIPSE (2r+1)^2) -> invert -> flip back ->
// occ' = 100 - round(gray*100/255).  The two flips cancel (the element is symmetric under both), so the
// device works in occupancy order:
//   1. 256-bin histogram of the occupancy bytes (shared-memory atomics, 16-byte loads);
//   2. one thread maps it through the occupancy->gray table and runs OpenCV's Otsu scan in fp64
//      (getThreshVal_Otsu_8u), then publishes the predicate table "gray(v) <= threshold";
//   3. row pass: distance along x to the nearest obstacle cell (edt_row_kernel's scan, shared with K1);
//   4. column pass: occ'(x,y) = 100 iff some |dy| <= r has g(x, y+dy) <= dx(dy), dx = OpenCV's ellipse
//      half-widths (computed on the host with the same cvRound/sqrt as getStructuringElement), else 0.
#include <math.h>

#include "hpp_internal.cuh"

struct pred_lut {
  uint32_t w[8];
};

// defined in edt.cu: the row pass (K1 pass 1) with an arbitrary obstacle predicate table
hpp_status hpp_row_dist_launch(hpp_ctx *c, const int8_t *occ_dev, int W, int H, const void *lut_dev, uint16_t *g,
                               float *obstacle, uint32_t *any_obstacle);

__global__ void __launch_bounds__(256) occ_hist_kernel(const int8_t *__restrict__ occ, size_t n,
                                                       unsigned int *__restrict__ hist) {
  __shared__ unsigned int h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  // the caller's pointer may be any byte address (a view into a larger buffer): bytes up to the first 16-byte
  // boundary and after the last whole 16-byte group are read one by one, the rest with 16-byte loads
  size_t head = (size_t)((16 - ((uintptr_t)occ & 15)) & 15);
  if (head > n) head = n;
  const size_t n16 = (n - head) / 16;
  const int4 *p = (const int4 *)(occ + head);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
    int4 v = __ldg(p + i);
    uint8_t b[16];
    *(int4 *)b = v;
#pragma unroll
    for (int k = 0; k < 16; k++) atomicAdd(&h[b[k]], 1u);
  }
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < head; i += stride) atomicAdd(&h[(uint8_t)occ[i]], 1u);
  for (size_t i = head + n16 * 16 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    atomicAdd(&h[(uint8_t)occ[i]], 1u);
  __syncthreads();
  if (h[threadIdx.x]) atomicAdd(&hist[threadIdx.x], h[threadIdx.x]);
}

// gray_of[v] for v = occupancy byte as uint8 (MI/src/map_inflate_node.cpp:94-97)
struct gray_table {
  uint8_t g[256];
};

__global__ void otsu_kernel(const unsigned int *__restrict__ occ_hist, gray_table gt, double n_px,
                            pred_lut *__restrict__ lut_out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  // histogram of gray values
  unsigned int h[256];
  for (int i = 0; i < 256; i++) h[i] = 0;
  for (int v = 0; v < 256; v++) h[gt.g[v]] += occ_hist[v];
  // OpenCV getThreshVal_Otsu_8u (imgproc/src/thresh.cpp), called at MI/src/map_inflate_node.cpp:127
  double mu = 0, scale = 1. / n_px;
  for (int i = 0; i < 256; i++) mu += i * (double)h[i];
  mu *= scale;
  double mu1 = 0, q1 = 0, max_sigma = 0, max_val = 0;
  const double FLT_EPS = 1.1920928955078125e-7;
  for (int i = 0; i < 256; i++) {
    double p_i = h[i] * scale;
    mu1 *= q1;
    q1 += p_i;
    double q2 = 1. - q1;
    if (fmin(q1, q2) < FLT_EPS || fmax(q1, q2) > 1. - FLT_EPS) continue;
    mu1 = (mu1 + i * p_i) / q1;
    double mu2 = (mu - q1 * mu1) / q2;
    double sigma = q1 * q2 * (mu1 - mu2) * (mu1 - mu2);
    if (sigma > max_sigma) {
      max_sigma = sigma;
      max_val = i;
    }
  }
  const int thr = (int)max_val;
  // obstacle (white after the invert) <=> !(gray > thr)
  pred_lut l;
  for (int k = 0; k < 8; k++) l.w[k] = 0;
  for (int v = 0; v < 256; v++)
    if (!(gt.g[v] > thr)) l.w[v >> 5] |= 1u << (v & 31);
  *lut_out = l;
}

#define INF_MAX_R 255
struct ellipse_rows {
  int r;
  uint16_t dx[2 * INF_MAX_R + 1];
};

__global__ void __launch_bounds__(256) inflate_col_kernel(const uint16_t *__restrict__ g, int W, int H,
                                                          const ellipse_rows *__restrict__ el_dev,
                                                          int8_t *__restrict__ out, int8_t hit_value) {
  __shared__ uint16_t dxs[2 * INF_MAX_R + 1];
  const int r = el_dev->r;
  for (int i = threadIdx.x; i < 2 * r + 1; i += blockDim.x) dxs[i] = el_dev->dx[i];
  __syncthreads();
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  int hit = 0;
  const int lo = max(0, y - r), hi = min(H - 1, y + r);  // dilate border: outside pixels never win
  for (int yy = lo; yy <= hi && !hit; yy++) hit = g[(size_t)yy * W + x] <= dxs[yy - y + r];
  // dilated 255 -> inverted 0 -> occ' = 100 - round(0) = 100; else 100 - round(255*100/255) = 0  (:151-157)
  out[(size_t)y * W + x] = hit ? hit_value : (int8_t)0;  // hit_value: 100 (occupancy out) or -1 = 255 (cv::dilate out)
}

hpp_status hpp_inflate_launch(hpp_ctx *c, const int8_t *occ_dev, uint32_t W, uint32_t H, double resolution,
                              double inflation_radius, int8_t *out_dev) {
  const size_t n = (size_t)W * H;
  int r = (int)round(inflation_radius / resolution);  // MI/src/map_inflate_node.cpp:118
  if (r < 1) r = 1;                                    // :120-123
  if (r > INF_MAX_R) HPP_FAIL(c, HPP_ERR_UNSUPPORTED, "inflation radius of %d cells > %d", r, INF_MAX_R);
  if (W >= 65535u) HPP_FAIL(c, HPP_ERR_UNSUPPORTED, "inflate supports W < 65535");
  // scratch slot 2: g rows (uint16) | hist (256 u32) | lut | ellipse
  void *scr;
  const size_t g_bytes = (n * 2 + 255) & ~(size_t)255;
  hpp_status s = hpp_scratch(c, 2, g_bytes + 1024 + 256 + sizeof(ellipse_rows) + 256, &scr);
  if (s != HPP_OK) return s;
  uint16_t *d_g = (uint16_t *)scr;
  unsigned int *d_hist = (unsigned int *)((uint8_t *)scr + g_bytes);
  pred_lut *d_lut = (pred_lut *)((uint8_t *)scr + g_bytes + 1024);
  uint32_t *d_any = (uint32_t *)((uint8_t *)scr + g_bytes + 1024 + 64);
  ellipse_rows *d_el = (ellipse_rows *)((uint8_t *)scr + g_bytes + 1024 + 256);

  gray_table gt;
  for (int v = 0; v < 256; v++) {
    int occ_prob = (int8_t)v;
    occ_prob = (occ_prob < 0) ? 100 : occ_prob;  // :95
    gt.g[v] = (uint8_t)(int)(255 - round(occ_prob * 255.0 / 100.0));  // :96
  }
  // cv::getStructuringElement(MORPH_ELLIPSE): dx = cvRound(c * sqrt((r*r - dy*dy) * inv_r2)), c = r (hazard H8)
  void *hp;
  s = hpp_pinned(c, 1, sizeof(ellipse_rows), &hp);
  if (s != HPP_OK) return s;
  HPP_CUDA(c, cudaStreamSynchronize(c->stream));  // the pinned slot may still feed an earlier copy
  ellipse_rows *el = (ellipse_rows *)hp;
  el->r = r;
  const double inv_r2 = 1. / ((double)r * r);
  for (int i = 0; i < 2 * r + 1; i++) {
    int dy = i - r;
    el->dx[i] = (uint16_t)(int)nearbyint(r * sqrt((r * r - dy * dy) * inv_r2));
  }
  HPP_CUDA(c, cudaMemcpyAsync(d_el, el, sizeof(ellipse_rows), cudaMemcpyHostToDevice, c->stream));
  HPP_CUDA(c, cudaMemsetAsync(d_hist, 0, 1024 + 256, c->stream));
  int grid = (int)((n / 4096 + 1) < (size_t)c->sm_count * 8 ? (n / 4096 + 1) : (size_t)c->sm_count * 8);
  occ_hist_kernel<<<grid, 256, 0, c->stream>>>(occ_dev, n, d_hist);
  HPP_LAUNCHED(c);
  otsu_kernel<<<1, 32, 0, c->stream>>>(d_hist, gt, (double)W * (double)H, d_lut);
  HPP_LAUNCHED(c);
  s = hpp_row_dist_launch(c, occ_dev, (int)W, (int)H, d_lut, d_g, nullptr, d_any);
  if (s != HPP_OK) return s;
  dim3 cgrid(hpp_div_up(W, 256), H);
  inflate_col_kernel<<<cgrid, 256, 0, c->stream>>>(d_g, (int)W, (int)H, d_el, out_dev, (int8_t)100);
  HPP_LAUNCHED(c);
  return HPP_OK;
}

extern "C" hpp_status hpp_inflate_dev(hpp_ctx *c, const int8_t *occ_dev, uint32_t W, uint32_t H, double resolution,
                                      double inflation_radius, int8_t *out_dev) {
  if (!c || !occ_dev || !out_dev || W == 0 || H == 0 || !(resolution > 0)) return HPP_ERR_ARG;
  HPP_CUDA(c, cudaSetDevice(c->device));
  return hpp_inflate_launch(c, occ_dev, W, H, resolution, inflation_radius, out_dev);
}

extern "C" hpp_status hpp_inflate(hpp_ctx *c, const int8_t *occ, uint32_t W, uint32_t H, double resolution,
                                  double inflation_radius, int8_t *out) {
  if (!c || !occ || !out || W == 0 || H == 0 || !(resolution > 0)) return HPP_ERR_ARG;
  HPP_CUDA(c, cudaSetDevice(c->device));
  const size_t n = (size_t)W * H;
  void *d_in, *d_out;
  hpp_status s = hpp_scratch(c, 0, n, &d_in);
  if (s != HPP_OK) return s;
  s = hpp_scratch(c, 1, n, &d_out);
  if (s != HPP_OK) return s;
  HPP_CUDA(c, cudaMemcpyAsync(d_in, occ, n, cudaMemcpyHostToDevice, c->stream));
  s = hpp_inflate_launch(c, (const int8_t *)d_in, W, H, resolution, inflation_radius, (int8_t *)d_out);
  if (s != HPP_OK) return s;
  HPP_CUDA(c, cudaMemcpyAsync(out, d_out, n, cudaMemcpyDeviceToHost, c->stream));
  HPP_CUDA(c, cudaStreamSynchronize(c->stream));
  return HPP_OK;
}

// ------------------------------------------------------------------ the two OpenCV calls of InflateMap, as such
// MI/src/map_inflate_node.cpp:115-139 is `threshold(OTSU) -> bitwise_not -> dilate(ellipse) -> bitwise_not` on a gray
// cv::Mat.  host/dropin/mi/opencv2/opencv.hpp maps exactly those calls onto the two entry points below, so the node's
// source compiles and runs unmodified with the image work on the device (same kernels as hpp_inflate).
__global__ void __launch_bounds__(256) lut_binarize_kernel(const uint8_t *__restrict__ src, size_t n,
                                                           const pred_lut *__restrict__ lut, uint8_t *__restrict__ dst) {
  __shared__ pred_lut l;
  if (threadIdx.x < 8) l.w[threadIdx.x] = lut->w[threadIdx.x];
  __syncthreads();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const unsigned v = src[i];
    dst[i] = ((l.w[v >> 5] >> (v & 31)) & 1u) ? 0 : 255;  // lut = "gray <= threshold"
  }
}

__global__ void otsu_threshold_out_kernel(const pred_lut *__restrict__ lut, uint32_t *__restrict__ thr_out) {
  // the predicate table is "v <= thr" for the identity gray table: thr = number of set bits - 1
  int cnt = 0;
  for (int k = 0; k < 8; k++) cnt += __popc(lut->w[k]);
  *thr_out = (uint32_t)(cnt - 1);
}

// cv::threshold(src, dst, 0, 255, THRESH_BINARY | THRESH_OTSU) on an 8-bit single-channel image (host buffers)
extern "C" hpp_status hpp_cv_threshold_otsu(hpp_ctx *c, const uint8_t *src, uint32_t W, uint32_t H, uint8_t *dst,
                                            double *threshold_out) {
  if (!c || !src || !dst || W == 0 || H == 0) return HPP_ERR_ARG;
  HPP_CUDA(c, cudaSetDevice(c->device));
  const size_t n = (size_t)W * H;
  void *d_in, *d_out, *scr;
  hpp_status s = hpp_scratch(c, 0, n, &d_in);
  if (s != HPP_OK) return s;
  s = hpp_scratch(c, 1, n, &d_out);
  if (s != HPP_OK) return s;
  s = hpp_scratch(c, 2, 1024 + 256 + 64, &scr);
  if (s != HPP_OK) return s;
  unsigned int *d_hist = (unsigned int *)scr;
  pred_lut *d_lut = (pred_lut *)((uint8_t *)scr + 1024);
  uint32_t *d_thr = (uint32_t *)((uint8_t *)scr + 1024 + 64);
  gray_table gt;
  for (int v = 0; v < 256; v++) gt.g[v] = (uint8_t)v;  // the image already holds gray values
  HPP_CUDA(c, cudaMemcpyAsync(d_in, src, n, cudaMemcpyHostToDevice, c->stream));
  HPP_CUDA(c, cudaMemsetAsync(d_hist, 0, 1024 + 256 + 64, c->stream));
  int grid = (int)((n / 4096 + 1) < (size_t)c->sm_count * 8 ? (n / 4096 + 1) : (size_t)c->sm_count * 8);
  occ_hist_kernel<<<grid, 256, 0, c->stream>>>((const int8_t *)d_in, n, d_hist);
  HPP_LAUNCHED(c);
  otsu_kernel<<<1, 32, 0, c->stream>>>(d_hist, gt, (double)W * (double)H, d_lut);
  HPP_LAUNCHED(c);
  lut_binarize_kernel<<<grid, 256, 0, c->stream>>>((const uint8_t *)d_in, n, d_lut, (uint8_t *)d_out);
  HPP_LAUNCHED(c);
  otsu_threshold_out_kernel<<<1, 1, 0, c->stream>>>(d_lut, d_thr);
  HPP_LAUNCHED(c);
  HPP_CUDA(c, cudaMemcpyAsync(dst, d_out, n, cudaMemcpyDeviceToHost, c->stream));
  HPP_CUDA(c, cudaMemcpyAsync(c->h_flags + 52, d_thr, 4, cudaMemcpyDeviceToHost, c->stream));
  HPP_CUDA(c, cudaStreamSynchronize(c->stream));
  if (threshold_out) *threshold_out = (double)c->h_flags[52];
  return HPP_OK;
}

// cv::dilate(src, dst, getStructuringElement(MORPH_ELLIPSE, Size(2r+1, 2r+1), Point(r, r))) on a BINARY 8-bit image
// (pixels are 0 or non-zero, as cv::threshold leaves them; non-zero pixels come out as 255): host buffers
extern "C" hpp_status hpp_cv_dilate_ellipse(hpp_ctx *c, const uint8_t *src, uint32_t W, uint32_t H, int32_t r,
                                            int32_t src_complemented, uint8_t *dst) {
  if (!c || !src || !dst || W == 0 || H == 0 || r < 0) return HPP_ERR_ARG;
  // dilating the set pixels by the ellipse = inflating "occupied" cells: the inflate pipeline with a predicate table
  // that marks every non-zero byte, and 100/0 mapped to 255/0 on the way out
  if (r > INF_MAX_R) HPP_FAIL(c, HPP_ERR_UNSUPPORTED, "structuring element radius %d > %d", r, INF_MAX_R);
  if (W >= 65535u) HPP_FAIL(c, HPP_ERR_UNSUPPORTED, "dilate supports W < 65535");
  HPP_CUDA(c, cudaSetDevice(c->device));
  const size_t n = (size_t)W * H;
  void *d_in, *d_out, *scr;
  hpp_status s = hpp_scratch(c, 0, n, &d_in);
  if (s != HPP_OK) return s;
  s = hpp_scratch(c, 1, n, &d_out);
  if (s != HPP_OK) return s;
  const size_t g_bytes = (n * 2 + 255) & ~(size_t)255;
  s = hpp_scratch(c, 2, g_bytes + 1024 + 256 + sizeof(ellipse_rows) + 256, &scr);
  if (s != HPP_OK) return s;
  uint16_t *d_g = (uint16_t *)scr;
  pred_lut *d_lut = (pred_lut *)((uint8_t *)scr + g_bytes + 1024);
  uint32_t *d_any = (uint32_t *)((uint8_t *)scr + g_bytes + 1024 + 64);
  ellipse_rows *d_el = (ellipse_rows *)((uint8_t *)scr + g_bytes + 1024 + 256);
  void *hp;
  s = hpp_pinned(c, 1, sizeof(ellipse_rows) + sizeof(pred_lut), &hp);
  if (s != HPP_OK) return s;
  HPP_CUDA(c, cudaStreamSynchronize(c->stream));
  ellipse_rows *el = (ellipse_rows *)hp;
  pred_lut *lut = (pred_lut *)(el + 1);
  el->r = r;
  const double inv_r2 = r ? 1. / ((double)r * r) : 0.;
  for (int i = 0; i < 2 * r + 1; i++) {
    int dy = i - r;
    el->dx[i] = (uint16_t)(int)nearbyint(r * sqrt((r * r - dy * dy) * inv_r2));
  }
  if (src_complemented) {  // the bytes are the complement of the image (a cv::bitwise_not not yet applied): 255 = unset
    for (int k = 0; k < 8; k++) lut->w[k] = 0xffffffffu;
    lut->w[7] &= ~(1u << 31);
  } else {                 // every byte but 0 is a set pixel
    for (int k = 0; k < 8; k++) lut->w[k] = 0xffffffffu;
    lut->w[0] &= ~1u;
  }
  HPP_CUDA(c, cudaMemcpyAsync(d_el, el, sizeof(ellipse_rows), cudaMemcpyHostToDevice, c->stream));
  HPP_CUDA(c, cudaMemcpyAsync(d_lut, lut, sizeof(pred_lut), cudaMemcpyHostToDevice, c->stream));
  HPP_CUDA(c, cudaMemsetAsync(d_any, 0, 4, c->stream));
  HPP_CUDA(c, cudaMemcpyAsync(d_in, src, n, cudaMemcpyHostToDevice, c->stream));
  s = hpp_row_dist_launch(c, (const int8_t *)d_in, (int)W, (int)H, d_lut, d_g, nullptr, d_any);
  if (s != HPP_OK) return s;
  dim3 cgrid(hpp_div_up(W, 256), H);
  inflate_col_kernel<<<cgrid, 256, 0, c->stream>>>(d_g, (int)W, (int)H, d_el, (int8_t *)d_out, (int8_t)-1);
  HPP_LAUNCHED(c);
  HPP_CUDA(c, cudaMemcpyAsync(dst, d_out, n, cudaMemcpyDeviceToHost, c->stream));
  HPP_CUDA(c, cudaStreamSynchronize(c->stream));
  return HPP_OK;
}

// ------------------------------------------------------------------ image -> OccupancyGrid
// grid_map_cv addLayerFromImage<unsigned char,1>(image, layer, map, lowerValue = 255, upperValue = 0, alpha 1):
//   mapValue = lowerValue + (upperValue - lowerValue) * ((float)imageValue / 255.f), grid_map index (i, j) = pixel (i, j);
// grid_map_ros toOccupancyGrid(map, layer, dataMin = 0, dataMax = 255): value = (mapValue - 0) / 255, clamped to [0,1],
//   * 100, stored (float -> int8, truncation) at data[nCells - 1 - (j*rows + i)], info.width = rows, info.height = cols.
// (GP/src/grid_publisher.cpp:32-50,78-86; the arithmetic is grid_map's, an un-vendored dependency: restated as recalled.)
__global__ void __launch_bounds__(256) image_to_occ_kernel(const uint8_t *__restrict__ img, int rows, int cols,
                                                           int8_t *__restrict__ occ) {
  const size_t n = (size_t)rows * cols;
  for (size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
    // k = y * rows + x  (coalesced writes); pixel (i, j) = (rows-1-x, cols-1-y)
    const int x = (int)(k % rows), y = (int)(k / rows);
    const int i = rows - 1 - x, j = cols - 1 - y;
    const float v = (float)img[(size_t)i * cols + j];
    const float map_value = __fadd_rn(255.0f, __fmul_rn(-255.0f, __fdiv_rn(v, 255.0f)));
    float val = __fdiv_rn(__fsub_rn(map_value, 0.0f), 255.0f);
    val = fminf(fmaxf(0.0f, val), 1.0f);
    occ[k] = (int8_t)__fadd_rn(0.0f, __fmul_rn(val, 100.0f));
  }
}

extern "C" hpp_status hpp_image_to_occupancy_dev(hpp_ctx *c, const uint8_t *img_dev, uint32_t rows, uint32_t cols,
                                                 int8_t *occ_dev) {
  if (!c || !img_dev || !occ_dev || rows == 0 || cols == 0) return HPP_ERR_ARG;
  HPP_CUDA(c, cudaSetDevice(c->device));
  const size_t n = (size_t)rows * cols;
  const long long want = hpp_div_up((long long)n, 256);
  const int grid = (int)(want < (long long)c->sm_count * 16 ? want : (long long)c->sm_count * 16);
  image_to_occ_kernel<<<grid, 256, 0, c->stream>>>(img_dev, (int)rows, (int)cols, occ_dev);
  HPP_LAUNCHED(c);
  return HPP_OK;
}

extern "C" hpp_status hpp_image_to_occupancy(hpp_ctx *c, const uint8_t *img, uint32_t rows, uint32_t cols, int8_t *occ) {
  if (!c || !img || !occ || rows == 0 || cols == 0) return HPP_ERR_ARG;
  HPP_CUDA(c, cudaSetDevice(c->device));
  const size_t n = (size_t)rows * cols;
  void *d_in, *d_out;
  hpp_status s = hpp_scratch(c, 0, n, &d_in);
  if (s != HPP_OK) return s;
  s = hpp_scratch(c, 1, n, &d_out);
  if (s != HPP_OK) return s;
  HPP_CUDA(c, cudaMemcpyAsync(d_in, img, n, cudaMemcpyHostToDevice, c->stream));
  s = hpp_image_to_occupancy_dev(c, (const uint8_t *)d_in, rows, cols, (int8_t *)d_out);
  if (s != HPP_OK) return s;
  HPP_CUDA(c, cudaMemcpyAsync(occ, d_out, n, cudaMemcpyDeviceToHost, c->stream));
  HPP_CUDA(c, cudaStreamSynchronize(c->stream));
  return HPP_OK;
}
