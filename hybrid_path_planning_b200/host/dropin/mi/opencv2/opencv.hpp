// opencv2/opencv.hpp of the map_inflate drop-in: put this directory BEFORE the real OpenCV on
// the include path of the map_inflate package and MI/src/map_inflate_node.cpp compiles and runs unmodified with its
// image work on the B200.  Only what that node uses exists here (MI/src/map_inflate_node.cpp:86-157):
//   cv::Mat (8-bit single channel: Mat(rows, cols, CV_8UC1), at<uchar>, clone), cv::Size, cv::Point,
//   cv::threshold(.., THRESH_BINARY | THRESH_OTSU)      -> hpp_cv_threshold_otsu   (histogram + Otsu + binarise on the device)
//   cv::dilate(.., getStructuringElement(MORPH_ELLIPSE)) -> hpp_cv_dilate_ellipse  (row distances + column OR on the device)
//   cv::bitwise_not                                       -> a flag on the Mat (applied lazily, no pass over the image)
// Any other OpenCV use is a compile error here by design: this header replaces OpenCV for this one node only.
#pragma once
#include <cmath>
#include <cstdint>
#include <iostream>  // the real opencv2/opencv.hpp pulls these in, and the node relies on it (cout, string)
#include <memory>
#include <string>
#include <stdexcept>
#include <vector>

#include "GridMapShell.h"  // hpp_host::thread_ctx / check, the C ABI

typedef unsigned char uchar;
#define CV_8UC1 0

namespace cv {

struct Size {
  int width = 0, height = 0;
  Size() = default;
  Size(int w, int h) : width(w), height(h) {}
};
struct Point {
  int x = 0, y = 0;
  Point() = default;
  Point(int x_, int y_) : x(x_), y(y_) {}
};

enum { THRESH_BINARY = 0, THRESH_OTSU = 8 };
enum { MORPH_RECT = 0, MORPH_CROSS = 1, MORPH_ELLIPSE = 2 };

class Mat {
 public:
  int rows = 0, cols = 0;
  Mat() = default;
  Mat(int r, int c, int type) : rows(r), cols(c), d_(std::make_shared<std::vector<uchar>>((std::size_t)r * c)) {
    if (type != CV_8UC1) throw std::runtime_error("hpp cv shim: only CV_8UC1 images");
  }
  bool empty() const { return !d_ || d_->empty(); }
  Mat clone() const {
    Mat m;
    m.rows = rows; m.cols = cols; m.inverted_ = inverted_; m.ellipse_r_ = ellipse_r_;
    if (d_) m.d_ = std::make_shared<std::vector<uchar>>(*d_);
    return m;
  }
  template <class T> T &at(int r, int c) {
    static_assert(sizeof(T) == 1, "hpp cv shim: 8-bit images only");
    settle();
    return reinterpret_cast<T &>((*d_)[(std::size_t)r * cols + c]);
  }
  template <class T> T at(int r, int c) const {
    static_assert(sizeof(T) == 1, "hpp cv shim: 8-bit images only");
    const uchar v = (*d_)[(std::size_t)r * cols + c];
    return (T)(inverted_ ? (uchar)~v : v);
  }
  // shim internals
  const uchar *raw() const { return d_ ? d_->data() : nullptr; }
  uchar *raw() { return d_ ? d_->data() : nullptr; }
  bool inverted() const { return inverted_; }
  void set_inverted(bool v) { inverted_ = v; }
  int ellipse_radius() const { return ellipse_r_; }
  void set_ellipse_radius(int r) { ellipse_r_ = r; }
  void settle() {  // apply a pending bitwise_not before the pixels are written through at<>()
    if (inverted_ && d_) {
      if (d_.use_count() > 1) d_ = std::make_shared<std::vector<uchar>>(*d_);
      for (auto &v : *d_) v = (uchar)~v;
      inverted_ = false;
    }
  }
  void share(const Mat &o) { rows = o.rows; cols = o.cols; d_ = o.d_; inverted_ = o.inverted_; ellipse_r_ = o.ellipse_r_; }
  void alloc(int r, int c) { rows = r; cols = c; d_ = std::make_shared<std::vector<uchar>>((std::size_t)r * c); inverted_ = false; }

 private:
  std::shared_ptr<std::vector<uchar>> d_;
  bool inverted_ = false;   // the stored bytes are the complement of the image (pending cv::bitwise_not)
  int ellipse_r_ = -1;      // >= 0: this Mat is getStructuringElement(MORPH_ELLIPSE, (2r+1)^2, (r, r))
};

// cv::threshold(src, dst, 0, 255, THRESH_BINARY | THRESH_OTSU) (MI/src/map_inflate_node.cpp:127)
inline double threshold(const Mat &src, Mat &dst, double /*thresh*/, double maxval, int type) {
  if (type != (THRESH_BINARY | THRESH_OTSU) || maxval != 255 || src.inverted())
    throw std::runtime_error("hpp cv shim: threshold supports (0, 255, THRESH_BINARY | THRESH_OTSU) on a plain image");
  hpp_ctx *ctx = hpp_host::thread_ctx();
  Mat out;
  out.alloc(src.rows, src.cols);
  double thr = 0;
  hpp_host::check(ctx, hpp_cv_threshold_otsu(ctx, src.raw(), (uint32_t)src.cols, (uint32_t)src.rows, out.raw(), &thr),
                  "hpp_cv_threshold_otsu");
  dst = out;
  return thr;
}

// cv::bitwise_not(src, dst) (:131, :138): O(1), the complement is applied when the pixels are next consumed
inline void bitwise_not(const Mat &src, Mat &dst) {
  Mat out;
  out.share(src);
  out.set_inverted(!src.inverted());
  dst = out;
}

// cv::getStructuringElement(MORPH_ELLIPSE, Size(2r+1, 2r+1), Point(r, r)) (:134)
inline Mat getStructuringElement(int shape, Size ksize, Point anchor = Point(-1, -1)) {
  const int r = ksize.width / 2;
  if (shape != MORPH_ELLIPSE || ksize.width != ksize.height || ksize.width != 2 * r + 1 ||
      !((anchor.x == r && anchor.y == r) || (anchor.x == -1 && anchor.y == -1)))
    throw std::runtime_error("hpp cv shim: getStructuringElement supports a centred MORPH_ELLIPSE of odd square size");
  Mat el(ksize.height, ksize.width, CV_8UC1);
  // the element's pixels as OpenCV rasterises them (not read by dilate below, kept for code that inspects the element)
  const double inv_r2 = r ? 1. / ((double)r * r) : 0;
  for (int i = 0; i < ksize.height; i++) {
    const int dy = i - r, dx = (int)std::nearbyint(r * std::sqrt((r * r - dy * dy) * inv_r2));
    for (int j = 0; j < ksize.width; j++) el.at<uchar>(i, j) = (j >= r - dx && j <= r + dx) ? 1 : 0;
  }
  el.set_ellipse_radius(r);
  return el;
}

// cv::dilate(src, dst, element) for a binary image and the ellipse element above (:135)
inline void dilate(const Mat &src, Mat &dst, const Mat &element) {
  if (element.ellipse_radius() < 0) throw std::runtime_error("hpp cv shim: dilate needs the MORPH_ELLIPSE element");
  hpp_ctx *ctx = hpp_host::thread_ctx();
  Mat out;
  out.alloc(src.rows, src.cols);
  // a pending complement of src (the bitwise_not in front of the dilation) is folded into the device call
  hpp_host::check(ctx, hpp_cv_dilate_ellipse(ctx, src.raw(), (uint32_t)src.cols, (uint32_t)src.rows,
                                             element.ellipse_radius(), src.inverted() ? 1 : 0, out.raw()),
                  "hpp_cv_dilate_ellipse");
  dst = out;
}

}  // namespace cv
