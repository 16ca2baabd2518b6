// TEST INFRASTRUCTURE — a stand-in for <opencv2/opencv.hpp>, just large enough to compile the reference's
// monocular_pose_estimator_lib/src/{led_detector,pose_estimator}.cpp UNMODIFIED (OpenCV's C++ headers are not installed in
// this image; only the Python binding cv2 4.13 is).  Not OpenCV code: container types written from the public API, and the
// seven imgproc/calib3d entry points LEDDetector::findLeds calls (led_detector.cpp:44,51,57,67,68,72,97) and the four calls of
// Visualization::createVisualizationImage (visualization.cpp:49-104: projectPoints, line, circle, rectangle) are forwarded to
// C callbacks registered at run time (cv_shim::callbacks(), set by oracle/ref_pose.py to the cv2 functions of the same
// name).  So the reference's own findLeds source drives the real OpenCV kernels.  Never included by the product library.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstring>
#include <memory>
#include <type_traits>
#include <vector>

#define CV_PI 3.1415926535897932384626433832795
#define CV_8U 0
#define CV_8UC1 0
#define CV_64F 6
#define CV_8UC3 16
#define CV_64FC1 6
#define CV_RGB(r, g, b) cv::Scalar((b), (g), (r), 0)
#define CV_GRAY2RGB 8
#define CV_GRAY2BGR 8
#define CV_RETR_EXTERNAL 0
#define CV_CHAIN_APPROX_NONE 1

namespace cv {

typedef unsigned char uchar;
enum { THRESH_BINARY = 0, THRESH_BINARY_INV = 1, THRESH_TRUNC = 2, THRESH_TOZERO = 3, THRESH_TOZERO_INV = 4 };
enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4, BORDER_DEFAULT = 4 };
enum { RETR_EXTERNAL = 0, CHAIN_APPROX_NONE = 1 };

namespace shim_detail {
// cv::saturate_cast: floating -> floating and integer -> anything are plain casts; floating -> int rounds to nearest even (cvRound)
template <typename T, typename U, bool kRound> struct Sat { static T run(U v) { return (T)v; } };
template <typename T, typename U> struct Sat<T, U, true> { static T run(U v) { return (T)std::lrint((double)v); } };
template <typename T, typename U> inline T sat(U v) { return Sat<T, U, std::is_integral<T>::value && std::is_floating_point<U>::value>::run(v); }
}
template <typename T> struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T x_, T y_) : x(x_), y(y_) {}
  template <typename U> Point_(const Point_<U>& o) : x(shim_detail::sat<T, U>(o.x)), y(shim_detail::sat<T, U>(o.y)) {}
  Point_ operator+(const Point_& o) const { return Point_((T)(x + o.x), (T)(y + o.y)); }
  Point_ operator-(const Point_& o) const { return Point_((T)(x - o.x), (T)(y - o.y)); }
};
typedef Point_<int> Point;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;
template <typename T> struct Point3_ { T x, y, z; Point3_() : x(0), y(0), z(0) {} Point3_(T a, T b, T c) : x(a), y(b), z(c) {} };
typedef Point3_<float> Point3f;

struct Size { int width, height; Size() : width(0), height(0) {} Size(int w, int h) : width(w), height(h) {} };
struct Rect {
  int x, y, width, height;
  Rect() : x(0), y(0), width(0), height(0) {}
  Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {}
};
struct Scalar { double val[4]; Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; } };
struct Moments { double m00, m10, m01, m20, m11, m02, m30, m21, m12, m03; Moments() { std::memset(this, 0, sizeof(*this)); } };

// A reference-counted 2-D array header: CV_8UC1, CV_8UC3 or CV_64FC1, row stride `step` in bytes.
class Mat {
 public:
  int rows, cols, type_;
  size_t step;
  uchar* data;
  Mat() : rows(0), cols(0), type_(CV_8UC1), step(0), data(nullptr) {}
  Mat(int r, int c, int type) { create(r, c, type); }
  Mat(int r, int c, int type, void* external, size_t step_ = 0) : rows(r), cols(c), type_(type), step(step_ ? step_ : (size_t)c * elem(type)), data((uchar*)external) {}
  void create(int r, int c, int type) {
    rows = r; cols = c; type_ = type; step = (size_t)c * elem(type);
    owner_.reset(new std::vector<uchar>((size_t)r * step, 0)); data = owner_->data();
  }
  static size_t elem(int type) { return type == CV_64F ? 8 : (type == CV_8UC3 ? 3 : 1); }
  static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }     // create() zero-fills
  int channels() const { return type_ == CV_8UC3 ? 3 : 1; }
  Mat operator()(const Rect& roi) const {                       // a view, no copy (as cv::Mat::operator())
    Mat m; m.rows = roi.height; m.cols = roi.width; m.type_ = type_; m.step = step; m.owner_ = owner_;
    m.data = data + (size_t)roi.y * step + (size_t)roi.x * elem(type_);
    return m;
  }
  Mat clone() const {
    Mat m(rows, cols, type_);
    for (int i = 0; i < rows; ++i) std::memcpy(m.data + (size_t)i * m.step, data + (size_t)i * step, (size_t)cols * elem(type_));
    return m;
  }
  Size size() const { return Size(cols, rows); }
  int type() const { return type_; }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  template <typename T> T& at(int i, int j) { return *(T*)(data + (size_t)i * step + (size_t)j * sizeof(T)); }
  template <typename T> const T& at(int i, int j) const { return *(const T*)(data + (size_t)i * step + (size_t)j * sizeof(T)); }
 private:
  std::shared_ptr<std::vector<uchar> > owner_;
};

// cv::cvtColor for the one conversion the ROS node makes (GRAY2RGB, monocular_pose_estimator.cpp:201; in place allowed)
inline void cvtColor(const Mat& src, Mat& dst, int code) {
  if (code != CV_GRAY2RGB || src.type() != CV_8UC1) return;
  Mat out(src.rows, src.cols, CV_8UC3);
  for (int r = 0; r < src.rows; ++r)
    for (int c = 0; c < src.cols; ++c) {
      const uchar v = src.data[(size_t)r * src.step + (size_t)c];
      uchar* o = out.data + (size_t)r * out.step + (size_t)3 * c;
      o[0] = v; o[1] = v; o[2] = v;
    }
  dst = out;
}

struct NoArray {};
inline NoArray noArray() { return NoArray(); }

double threshold(const Mat& src, Mat& dst, double thresh, double maxval, int type);
void GaussianBlur(const Mat& src, Mat& dst, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_DEFAULT);
void findContours(const Mat& image, std::vector<std::vector<Point> >& contours, int mode, int method);
double contourArea(const std::vector<Point>& contour, bool oriented = false);
Rect boundingRect(const std::vector<Point>& contour);
Moments moments(const std::vector<Point>& contour, bool binaryImage = false);
void undistortPoints(const std::vector<Point2f>& src, std::vector<Point2f>& dst, const Mat& cameraMatrix,
                     const std::vector<double>& distCoeffs, NoArray R, const Mat& P);
// what Visualization::createVisualizationImage calls (visualization.cpp:49-104)
void projectPoints(const std::vector<Point3f>& objectPoints, const Mat& rvec, const Mat& tvec, const Mat& cameraMatrix,
                   const std::vector<double>& distCoeffs, std::vector<Point2f>& imagePoints);
void line(Mat& img, Point pt1, Point pt2, const Scalar& color, int thickness = 1);
void circle(Mat& img, Point center, int radius, const Scalar& color, int thickness = 1);
void rectangle(Mat& img, Rect rec, const Scalar& color, int thickness = 1);

}  // namespace cv

// ---- run-time binding of the forwarded calls (C ABI, filled in from Python with ctypes callbacks around cv2)
extern "C" {
struct cv_shim_callbacks {
  // dst is rows x cols, contiguous
  void (*threshold)(const unsigned char* src, int rows, int cols, long step, double thresh, double maxval, int type, unsigned char* dst);
  void (*gaussian_blur)(const unsigned char* src, int rows, int cols, long step, double sigma_x, double sigma_y, int border, unsigned char* dst);
  // returns the number of contours; *counts -> int[n_contours], *points -> int[2*sum(counts)] (x,y), valid until the next call
  int (*find_contours)(const unsigned char* img, int rows, int cols, long step, int mode, int method, const int** counts, const int** points);
  double (*contour_area)(const int* pts, int n);
  void (*bounding_rect)(const int* pts, int n, int out_xywh[4]);
  void (*moments)(const int* pts, int n, double out10[10]);
  void (*undistort_points)(const float* src, int n, const double K[9], const double* D, int nD, const double P[9], float* dst);
  void (*project_points)(const float* xyz, int n, const double rvec[3], const double tvec[3], const double K[9], const double* D, int nD, float* out_xy);
  // what: 0 line (x1,y1,x2,y2), 1 circle (cx,cy,radius,-), 2 rectangle (x,y,w,h); drawn in place
  void (*draw)(unsigned char* img, int rows, int cols, long step, int channels, int what, const int geom[4], const double color[4], int thickness);
};
}
namespace cv_shim { cv_shim_callbacks& callbacks(); }
