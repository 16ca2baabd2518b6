// Header-only C++ shim: re-creates the class API of monocular_pose_estimator_lib (LEDDetector, PoseEstimator) on top
// of the C ABI in mpe_b200.h, so that the reference's callers (MPENode: monocular_pose_estimator/src/
// monocular_pose_estimator.cpp:84,110-120,159,163-164,222-233; the nodelet wraps MPENode) compile against it unchanged.
//
//   * With Eigen and OpenCV headers available the reference's own typedefs are used (datatypes.h:38-52, cv::Mat, ...).
//   * Without them (this build image has neither) minimal stand-ins with the same member syntax are used, so the shim can
//     still be compiled and exercised (tests/cpp/shim_demo.cpp); define MPE_SHIM_FORCE_STANDIN to force that mode.
//
// PoseEstimator::estimateBodyPose (pose_estimator.cpp:62-147) has two implementations here (setDeviceLoop):
//   * device loop (default): ONE call per image, mpe_streams_step with a single stream — the whole state machine and the
//     estimator state live on the GPU (CUDA-graph replay), the members are refreshed from the returned record;
//   * stage by stage: the state machine and the tiny sequential helpers of tracking mode (predictPose, determineROI,
//     findCorrespondences) run on the host exactly as in the reference and call one CUDA stage at a time
//     (mpe_find_leds / mpe_initialise / mpe_check_correspondences / mpe_optimise_pose) — for callers that drive the public
//     stage methods and setters themselves.
#ifndef MPE_B200_SHIM_H_
#define MPE_B200_SHIM_H_

#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>

#include "../mpe_b200.h"

#if !defined(MPE_SHIM_FORCE_STANDIN) && defined(__has_include)
#if __has_include(<Eigen/Dense>) && (__has_include(<opencv2/core.hpp>) || __has_include(<opencv2/opencv.hpp>))
#define MPE_SHIM_REAL_TYPES 1
#endif
#endif

#ifdef MPE_SHIM_REAL_TYPES
#include <Eigen/Dense>
#if __has_include(<opencv2/core.hpp>)
#include <opencv2/core.hpp>
#if __has_include(<opencv2/imgproc.hpp>) && __has_include(<opencv2/calib3d.hpp>) && !defined(MPE_SHIM_NO_OPENCV_DRAWING)
#include <opencv2/calib3d.hpp>
#include <opencv2/imgproc.hpp>
#define MPE_SHIM_OPENCV_DRAWING 1      // augmentImage draws the reference's overlay (visualization.cpp) with OpenCV, on the host
#endif
#else
#include <opencv2/opencv.hpp>           // one-header installations: define MPE_SHIM_OPENCV_DRAWING if projectPoints / line / circle / rectangle exist
#endif
namespace monocular_pose_estimator {
// datatypes.h:38-52 (all of them, so that sources written against the reference's header compile unchanged) + led_detector.h:39
typedef Eigen::Matrix<double, 6, 6> Matrix6d;
typedef Eigen::Matrix<double, 2, 6> Matrix2x6d;
typedef Eigen::Matrix<double, 3, 4> Matrix3x4d;
typedef Eigen::Matrix<double, Eigen::Dynamic, Eigen::Dynamic> MatrixXYd;
typedef Eigen::Matrix<unsigned, Eigen::Dynamic, Eigen::Dynamic> MatrixXYu;
typedef Eigen::Matrix<double, 6, 1> Vector6d;
typedef Eigen::Matrix<unsigned, 3, 1> Vector3u;
typedef Eigen::Matrix<unsigned, 4, 1> Vector4u;
typedef Eigen::Matrix<unsigned, Eigen::Dynamic, 1> VectorXu;
typedef Eigen::Matrix<unsigned, Eigen::Dynamic, 2> VectorXuPairs;
typedef Eigen::Matrix<double, 1, Eigen::Dynamic> RowXd;
typedef Eigen::Matrix<unsigned, 1, Eigen::Dynamic> RowXu;
typedef Eigen::Matrix<Eigen::Vector2d, Eigen::Dynamic, 1> List2DPoints;
typedef Eigen::Matrix<Eigen::Vector3d, Eigen::Dynamic, 1> List3DPoints;
typedef Eigen::Matrix<Eigen::Vector4d, Eigen::Dynamic, 1> List4DPoints;
typedef Eigen::Matrix4d Matrix4dT;
typedef cv::Mat ImageT;
typedef cv::Mat CameraMatT;
typedef cv::Rect RectT;
typedef cv::Size SizeT;
typedef cv::Point2f Point2fT;
inline double cam_at(const CameraMatT& K, int r, int c) { return K.at<double>(r, c); }
}  // namespace monocular_pose_estimator
#else
namespace monocular_pose_estimator {
namespace standin {
template <int N> struct Vec {
  double v[N];
  Vec() { for (int i = 0; i < N; ++i) v[i] = 0; }
  double& operator()(int i) { return v[i]; }
  double operator()(int i) const { return v[i]; }
};
template <typename T> struct List {
  std::vector<T> d;
  size_t size() const { return d.size(); }
  void resize(size_t n) { d.resize(n); }
  T& operator()(size_t i) { return d[i]; }
  const T& operator()(size_t i) const { return d[i]; }
};
template <int R, int C> struct Mat {   // column-major like Eigen
  double m[R * C];
  Mat() { for (int i = 0; i < R * C; ++i) m[i] = 0; }
  double& operator()(int r, int c) { return m[c * R + r]; }
  double operator()(int r, int c) const { return m[c * R + r]; }
};
template <int N> struct UVec { unsigned v[N]; UVec() { for (int i = 0; i < N; ++i) v[i] = 0; } unsigned& operator()(int i) { return v[i]; } unsigned operator()(int i) const { return v[i]; } };
template <typename T> struct DynMat {   // column-major like Eigen
  std::vector<T> d; int r = 0, c = 0;
  int rows() const { return r; } int cols() const { return c; }
  void resize(int rows, int cols) { r = rows; c = cols; d.assign((size_t)rows * cols, T()); }
  T& operator()(int i, int j) { return d[(size_t)j * r + i]; }
  const T& operator()(int i, int j) const { return d[(size_t)j * r + i]; }
};
struct Pairs {
  std::vector<unsigned> d; int n = 0;
  int rows() const { return n; }
  void resize(int r, int) { n = r; d.assign((size_t)r * 2, 0u); }
  unsigned& operator()(int r, int c) { return d[(size_t)r * 2 + c]; }
  unsigned operator()(int r, int c) const { return d[(size_t)r * 2 + c]; }
};
struct Image { const uint8_t* data = nullptr; int rows = 0, cols = 0; size_t step = 0; };
struct Camera { double k[9]; };   // row-major 3x3
struct Rect { int x = 0, y = 0, width = 0, height = 0; Rect() {} Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {} };
struct Size { int width = 0, height = 0; Size() {} Size(int w, int h) : width(w), height(h) {} };
struct Point2f { float x = 0, y = 0; Point2f() {} Point2f(float x_, float y_) : x(x_), y(y_) {} };
}  // namespace standin
typedef standin::Mat<6, 6> Matrix6d;
typedef standin::Mat<2, 6> Matrix2x6d;
typedef standin::Mat<3, 4> Matrix3x4d;
typedef standin::Vec<6> Vector6d;
typedef standin::Pairs VectorXuPairs;
typedef standin::List<standin::Vec<2> > List2DPoints;
typedef standin::List<standin::Vec<3> > List3DPoints;
typedef standin::List<standin::Vec<4> > List4DPoints;
typedef standin::List<double> RowXd;
typedef standin::List<unsigned> RowXu;
typedef standin::List<unsigned> VectorXu;
typedef standin::UVec<3> Vector3u;
typedef standin::UVec<4> Vector4u;
typedef standin::DynMat<double> MatrixXYd;
typedef standin::DynMat<unsigned> MatrixXYu;
typedef standin::Mat<4, 4> Matrix4dT;
typedef standin::Image ImageT;
typedef standin::Camera CameraMatT;
typedef standin::Rect RectT;
typedef standin::Size SizeT;
typedef standin::Point2f Point2fT;
inline double cam_at(const CameraMatT& K, int r, int c) { return K.k[3 * r + c]; }
}  // namespace monocular_pose_estimator
#endif

namespace monocular_pose_estimator {

namespace detail {
inline void check(mpe_ctx* c, int rc, const char* what) {
  if (rc != MPE_OK) throw std::runtime_error(std::string(what) + ": " + (c ? mpe_last_error(c) : "no context"));
}
#ifdef MPE_SHIM_OPENCV_DRAWING
// Visualization::createVisualizationImage (visualization.cpp:57-104) + projectOrientationVectorsOnImage (:37-55): body axes in
// red / green / blue from the projected trivector, a circle per detection centre, the region of interest.  `pose` row-major 4x4.
inline void draw_overlay(ImageT& image, const double pose[16], const CameraMatT& K, const std::vector<double>& D, const RectT& roi,
                         const std::vector<Point2fT>& centers) {
  const double len = 0.075;                                       // orientation_vector_length
  const double O[4][4] = {{0, len, 0, 0}, {0, 0, len, 0}, {0, 0, 0, len}, {1, 1, 1, 1}};   // columns: origin, x, y, z axis tips
  std::vector<cv::Point3f> pts(4);
  for (int c = 0; c < 4; ++c) {
    double v[3];
    for (int r = 0; r < 3; ++r) {                                 // transform * orientation_vector_points, k ascending
      double acc = pose[4 * r] * O[0][c];
      for (int k = 1; k < 4; ++k) acc = acc + pose[4 * r + k] * O[k][c];
      v[r] = acc;
    }
    pts[c] = cv::Point3f((float)v[0], (float)v[1], (float)v[2]);
  }
  std::vector<cv::Point2f> proj;
  cv::Mat rvec = cv::Mat::zeros(3, 1, CV_64F), tvec = cv::Mat::zeros(3, 1, CV_64F);
  cv::projectPoints(pts, rvec, tvec, K, D, proj);
  cv::line(image, proj[0], proj[1], CV_RGB(255, 0, 0), 2);
  cv::line(image, proj[0], proj[2], CV_RGB(0, 255, 0), 2);
  cv::line(image, proj[0], proj[3], CV_RGB(0, 0, 255), 2);
  for (size_t i = 0; i < centers.size(); ++i) cv::circle(image, centers[i], 10, CV_RGB(255, 0, 0), 2);
  cv::rectangle(image, roi, CV_RGB(0, 0, 255), 2);
}
#endif
inline void camera_to_rowmajor(const CameraMatT& K, double out[9]) {
  for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) out[3 * r + c] = cam_at(K, r, c);
}
}  // namespace detail

class LEDDetector {
 public:
  // led_detector.h:84-88.  Uses (and lazily creates) a process-wide context sized for the image; the PoseEstimator
  // shim below uses its own context instead.
  static void findLeds(const ImageT& image, RectT ROI, const int& threshold_value, const double& gaussian_sigma, const double& min_blob_area,
                       const double& max_blob_area, const double& max_width_height_distortion, const double& max_circular_distortion,
                       List2DPoints& pixel_positions, std::vector<Point2fT>& distorted_detection_centers, const CameraMatT& camera_matrix_K,
                       const std::vector<double>& camera_distortion_coeffs) {
    mpe_ctx* ctx = defaultContext(image.cols, image.rows);
    mpe_params p; std::memset(&p, 0, sizeof(p));
    p.threshold_value = threshold_value; p.gaussian_sigma = gaussian_sigma; p.min_blob_area = min_blob_area; p.max_blob_area = max_blob_area;
    p.max_width_height_distortion = max_width_height_distortion; p.max_circular_distortion = max_circular_distortion;
    p.back_projection_pixel_tolerance = 3; p.nearest_neighbour_pixel_tolerance = 5; p.certainty_threshold = 0.75; p.valid_correspondence_threshold = 0.7;
    detail::check(ctx, mpe_set_params(ctx, &p), "mpe_set_params");
    double K[9]; detail::camera_to_rowmajor(camera_matrix_K, K);
    detail::check(ctx, mpe_set_camera(ctx, K, camera_distortion_coeffs.data(), (int)camera_distortion_coeffs.size()), "mpe_set_camera");
    findLedsWith(ctx, image, ROI, pixel_positions, distorted_detection_centers);
  }

  static void findLedsWith(mpe_ctx* ctx, const ImageT& image, RectT ROI, List2DPoints& pixel_positions, std::vector<Point2fT>& centers) {
    double px[2 * MPE_MAX_BLOBS]; float ce[2 * MPE_MAX_BLOBS]; int n = 0, flags = 0;
    mpe_rect r; r.x = ROI.x; r.y = ROI.y; r.width = ROI.width; r.height = ROI.height;
    detail::check(ctx, mpe_find_leds(ctx, image.data, (int)(size_t)image.step, image.cols, image.rows, r, px, ce, &n, &flags), "mpe_find_leds");
    centers.clear();
    for (int i = 0; i < n; ++i) centers.push_back(Point2fT(ce[2 * i], ce[2 * i + 1]));       // led_detector.cpp:89
    if (n > 0) {                                                                               // :91 — untouched when nothing is found
      pixel_positions.resize(n);
      for (int i = 0; i < n; ++i) { pixel_positions(i)(0) = px[2 * i]; pixel_positions(i)(1) = px[2 * i + 1]; }
    }
  }

  // led_detector.cpp:114-179 with distortPoints :181-224 — computed by the library's host helper, the very function the device
  // loop runs per stream (csrc/tracking_math.cuh), so stage mode and device loop cannot drift apart
  static RectT determineROI(List2DPoints pixel_positions, SizeT image_size, const int border_size, const CameraMatT& K, const std::vector<double>& D) {
    std::vector<double> px(2 * (size_t)pixel_positions.size());
    for (unsigned i = 0; i < pixel_positions.size(); ++i) { px[2 * i] = pixel_positions(i)(0); px[2 * i + 1] = pixel_positions(i)(1); }
    double Kr[9]; detail::camera_to_rowmajor(K, Kr);
    mpe_rect r;
    detail::check(nullptr, mpe_host_determine_roi(px.data(), (int)pixel_positions.size(), image_size.width, image_size.height, border_size, Kr,
                                                  D.data(), (int)D.size(), &r), "mpe_host_determine_roi");
    return RectT(r.x, r.y, r.width, r.height);
  }

 private:
  static mpe_ctx* defaultContext(int w, int h) {
    static mpe_ctx* ctx = nullptr; static int cw = 0, ch = 0;
    if (!ctx || w > cw || h > ch) {
      if (ctx) mpe_destroy(ctx);
      cw = w > cw ? w : cw; ch = h > ch ? h : ch;
      if (mpe_create(&ctx, 0, 1, cw, ch) != MPE_OK) { ctx = nullptr; throw std::runtime_error("mpe_create failed: no CUDA device? (there is no CPU fallback)"); }
    }
    return ctx;
  }
};

class PoseEstimator {
 public:
  // public tunables, written directly by the caller (pose_estimator.h:82-91)
  CameraMatT camera_matrix_K_;
  std::vector<double> camera_distortion_coeffs_;
  int detection_threshold_value_;
  double gaussian_sigma_, min_blob_area_, max_blob_area_, max_width_height_distortion_, max_circular_distortion_;
  unsigned roi_border_thickness_;

  PoseEstimator() : ctx_(nullptr), ctx_w_(0), ctx_h_(0) {                      // pose_estimator.cpp:34-42
    back_projection_pixel_tolerance_ = 3; nearest_neighbour_pixel_tolerance_ = 5; certainty_threshold_ = 0.75;
    valid_correspondence_threshold_ = 0.7; it_since_initialized_ = 0; histogram_threshold_ = 0; pose_updated_ = false;
    detection_threshold_value_ = 0; gaussian_sigma_ = 0.6; min_blob_area_ = 0; max_blob_area_ = 0; max_width_height_distortion_ = 0;
    max_circular_distortion_ = 0; roi_border_thickness_ = 0; current_time_ = previous_time_ = predicted_time_ = 0;
    identity(current_pose_); identity(previous_pose_); identity(predicted_pose_);
    for (int i = 0; i < 36; ++i) cov_[i] = 0;
  }
  ~PoseEstimator() { if (ctx_) mpe_destroy(ctx_); }
  PoseEstimator(const PoseEstimator&) = delete;
  PoseEstimator& operator=(const PoseEstimator&) = delete;

  void setMarkerPositions(List4DPoints positions) {                             // :50-55
    object_points_ = positions;
    predicted_pixel_positions_.resize(object_points_.size());
    unsigned n = (unsigned)object_points_.size();
    histogram_threshold_ = numCombinations(n, 3);
    markers_dirty_ = true;
  }
  List4DPoints getMarkerPositions() { return object_points_; }
  // pose_estimator.cpp:44-48 -> Visualization::createVisualizationImage (visualization.cpp:57-104): the debug overlay MPENode
  // publishes when somebody subscribes (monocular_pose_estimator.cpp:198-214).  Host side, OpenCV drawing, like the reference —
  // it is next to the pose path, not part of it; what it draws (pose, region of interest, detection centres) came back from the
  // device in the result record.  Without OpenCV's drawing API (stand-in types) the image is left as it is.
#ifdef MPE_SHIM_OPENCV_DRAWING
  void augmentImage(ImageT& image) {
    detail::draw_overlay(image, predicted_pose_, camera_matrix_K_, camera_distortion_coeffs_, region_of_interest_, distorted_detection_centers_);
  }
#else
  void augmentImage(ImageT& /*image*/) {}
#endif
  void setPredictedPose(const Matrix4dT& pose, double time) { fromEigen(pose, predicted_pose_); predicted_time_ = time; }
  Matrix4dT getPredictedPose() { return toEigen(predicted_pose_); }
  Matrix6d getPoseCovariance() { Matrix6d c; for (int r = 0; r < 6; ++r) for (int q = 0; q < 6; ++q) c(r, q) = cov_[6 * r + q]; return c; }
  void setImagePoints(List2DPoints points) { image_points_ = points; }
  List2DPoints getImagePoints() { return image_points_; }
  void setPredictedPixels(List2DPoints points) { predicted_pixel_positions_ = points; }
  List2DPoints getPredictedPixelPositions() { return predicted_pixel_positions_; }
  void setCorrespondences(VectorXuPairs corrs) { correspondences_ = corrs; }
  VectorXuPairs getCorrespondences() { return correspondences_; }
  void setBackProjectionPixelTolerance(double t) { back_projection_pixel_tolerance_ = t; }
  double getBackProjectionPixelTolerance() { return back_projection_pixel_tolerance_; }
  void setNearestNeighbourPixelTolerance(double t) { nearest_neighbour_pixel_tolerance_ = t; }
  double getNearestNeighbourPixelTolerance() { return nearest_neighbour_pixel_tolerance_; }
  void setCertaintyThreshold(double t) { certainty_threshold_ = t; }
  double getCertaintyThreshold() { return certainty_threshold_; }
  void setValidCorrespondenceThreshold(double t) { valid_correspondence_threshold_ = t; }
  double getValidCorrespondenceThreshold() { return valid_correspondence_threshold_; }
  void setHistogramThreshold(unsigned t) { histogram_threshold_ = t; }
  unsigned getHistogramThreshold() { return histogram_threshold_; }
  void setPredictedTime(double t) { predicted_time_ = t; }
  double getPredictedTime() { return predicted_time_; }
  unsigned lastGaussNewtonIterations() const { return last_gn_iters_; }

  // Where the per-frame state machine runs.  true (default): one mpe_streams_step call per image — predictWithROI, findLeds,
  // findCorrespondences / checkCorrespondences / initialise and optimisePose all run on the GPU with the estimator state
  // resident there (replayed as a single CUDA graph), and the members below are refreshed from the returned record; this is
  // what MPENode needs (it only calls estimateBodyPose and the getters).  false: the state machine runs here, stage by stage
  // through mpe_find_leds / mpe_initialise / mpe_check_correspondences / mpe_optimise_pose, so that a caller may interleave
  // its own calls to the public stage methods and setters (setPredictedPose, setCorrespondences, ...) between frames.
  void setDeviceLoop(bool on) { device_loop_ = on; }
  bool deviceLoop() const { return device_loop_; }

  // pose_estimator.cpp:62-147
  bool estimateBodyPose(ImageT image, double time_to_predict) {
    pose_updated_ = false; too_many_detections_ = false;
    ensureContext(image.cols, image.rows);
    push();   // the caller may have changed the public fields since the last frame
    if (device_loop_) return estimateBodyPoseOnDevice(image, time_to_predict);
    List2DPoints detected;
    if (it_since_initialized_ < 1) {
      setPredictedTime(time_to_predict);
      region_of_interest_ = RectT(0, 0, image.cols, image.rows);
      LEDDetector::findLedsWith(ctx_, image, region_of_interest_, detected, distorted_detection_centers_);
      if (detected.size() >= min_num_leds_detected_) {
        setImagePoints(detected);
        if (initialise() == 1) optimiseAndUpdatePose(time_to_predict);
      }
    } else {
      predictWithROI(time_to_predict, image);
      LEDDetector::findLedsWith(ctx_, image, region_of_interest_, detected, distorted_detection_centers_);
      bool repeat_check = true; unsigned num_loops = 0;
      do {
        num_loops++;
        if (detected.size() >= min_num_leds_detected_) {
          setImagePoints(detected);
          findCorrespondencesAndPredictPose(time_to_predict);
          repeat_check = false;
        } else if (num_loops < 2) {
          region_of_interest_ = RectT(0, 0, image.cols, image.rows);
          LEDDetector::findLedsWith(ctx_, image, region_of_interest_, detected, distorted_detection_centers_);
        } else {
          repeat_check = false;
        }
      } while (repeat_check);
    }
    return pose_updated_;
  }

  // estimateBodyPose as one device step (mpe_streams_step with a single stream)
  bool estimateBodyPoseOnDevice(const ImageT& image, double time_to_predict) {
    mpe_result r;
    detail::check(ctx_, mpe_streams_step(ctx_, image.data, (int)(size_t)image.step, (long long)image.step * image.rows, image.cols, image.rows, 1,
                                         &time_to_predict, &r), "mpe_streams_step");
    predicted_time_ = time_to_predict;
    too_many_detections_ = (r.flags & MPE_F_TOO_MANY_DET) != 0;
    region_of_interest_ = RectT(r.roi.x, r.roi.y, r.roi.width, r.roi.height);
    const int nd = r.n_det < MPE_MAX_DET ? r.n_det : MPE_MAX_DET;
    if (nd > 0) {                                                               // pixel_positions untouched when nothing was found (led_detector.cpp:91)
      distorted_detection_centers_.resize(nd);
      List2DPoints det; det.resize(nd);
      for (int i = 0; i < nd; ++i) {
        det(i)(0) = r.det[2 * i]; det(i)(1) = r.det[2 * i + 1];
        distorted_detection_centers_[i] = Point2fT(r.centers[2 * i], r.centers[2 * i + 1]);
      }
      if (nd >= (int)min_num_leds_detected_) image_points_ = det;
    }
    correspondences_.resize(r.n_corr, 2);
    for (int i = 0; i < r.n_corr; ++i) { correspondences_(i, 0) = r.corr[2 * i]; correspondences_(i, 1) = r.corr[2 * i + 1]; }
    std::memcpy(predicted_pose_, r.pose, sizeof(r.pose));
    last_gn_iters_ = (unsigned)r.gn_iters;
    if (r.updated) {                                                            // optimiseAndUpdatePose bookkeeping (:802-812, :794-800)
      std::memcpy(cov_, r.cov, sizeof(r.cov));
      if (it_since_initialized_ < 2) it_since_initialized_++;
      updatePose();
      pose_updated_ = true;
    }
    return pose_updated_;
  }

  unsigned initialise() {                                                       // :544-721 (K2 + K3a + Kabsch)
    push();
    if (image_points_.size() > (size_t)MPE_MAX_DET) { too_many_detections_ = true; return 0; }   // capacity of the sweep's tables: no pose, no exception
    std::vector<double> det; flatten(image_points_, det);
    uint32_t corr[2 * MPE_MAX_LEDS]; int k = 0, ok = 0; double pose[16];
    detail::check(ctx_, mpe_initialise(ctx_, det.data(), (int)image_points_.size(), nullptr, corr, &k, pose, &ok), "mpe_initialise");
    correspondences_.resize(k, 2);
    for (int i = 0; i < k; ++i) { correspondences_(i, 0) = corr[2 * i]; correspondences_(i, 1) = corr[2 * i + 1]; }
    if (ok) std::memcpy(predicted_pose_, pose, sizeof(pose));
    return (unsigned)ok;
  }
  unsigned checkCorrespondences() {                                             // :394-542
    push();
    if (correspondences_.rows() < 4) return 0;
    if (image_points_.size() > (size_t)MPE_MAX_DET) { compactToMatchedDetections(); if (image_points_.size() < 4) return 0; }
    std::vector<double> det; flatten(image_points_, det);
    std::vector<uint32_t> corr; flattenCorr(corr);
    int ok = 0; double pose[16];
    detail::check(ctx_, mpe_check_correspondences(ctx_, det.data(), (int)image_points_.size(), corr.data(), correspondences_.rows(), pose, &ok), "mpe_check_correspondences");
    if (ok) std::memcpy(predicted_pose_, pose, sizeof(pose));
    return (unsigned)ok;
  }
  void optimisePose() {                                                         // :733-792
    push();
    std::vector<double> det; flatten(image_points_, det);
    std::vector<uint32_t> corr; flattenCorr(corr);
    int it = 0;
    detail::check(ctx_, mpe_optimise_pose(ctx_, det.data(), (int)image_points_.size(), corr.data(), correspondences_.rows(), predicted_pose_, cov_, &it), "mpe_optimise_pose");
    last_gn_iters_ = (unsigned)it;
  }
  void updatePose() {                                                           // :794-800
    std::memcpy(previous_pose_, current_pose_, sizeof(current_pose_)); std::memcpy(current_pose_, predicted_pose_, sizeof(current_pose_));
    previous_time_ = current_time_; current_time_ = predicted_time_;
  }
  void optimiseAndUpdatePose(double&) {                                         // :802-812
    optimisePose();
    if (it_since_initialized_ < 2) it_since_initialized_++;
    updatePose();
    pose_updated_ = true;
  }
  void predictWithROI(double& time_to_predict, const ImageT& image) {           // :814-829
    if (it_since_initialized_ >= 2) predictPose(time_to_predict); else setPredictedTime(time_to_predict);
    predictMarkerPositionsInImage();
    region_of_interest_ = LEDDetector::determineROI(getPredictedPixelPositions(), SizeT(image.cols, image.rows), (int)roi_border_thickness_,
                                                   camera_matrix_K_, camera_distortion_coeffs_);
  }
  void findCorrespondencesAndPredictPose(double& time_to_predict) {             // :831-848
    findCorrespondences();
    if (checkCorrespondences() == 1) optimiseAndUpdatePose(time_to_predict);
    else if (initialise() == 1) optimiseAndUpdatePose(time_to_predict);
  }
  void predictMarkerPositionsInImage() {                                        // :270-276 with project2d :251-268 ((K|0)*T first)
    predicted_pixel_positions_.resize(object_points_.size());
    double K[9]; detail::camera_to_rowmajor(camera_matrix_K_, K);
    std::vector<double> xyz(3 * (size_t)object_points_.size()), px(2 * (size_t)object_points_.size());
    for (unsigned i = 0; i < object_points_.size(); ++i) for (int k = 0; k < 3; ++k) xyz[3 * i + k] = object_points_(i)(k);
    detail::check(ctx_, mpe_host_project_markers(K, predicted_pose_, xyz.data(), (int)object_points_.size(), px.data()), "mpe_host_project_markers");
    for (unsigned i = 0; i < object_points_.size(); ++i) { predicted_pixel_positions_(i)(0) = px[2 * i]; predicted_pixel_positions_(i)(1) = px[2 * i + 1]; }
  }
  void findCorrespondences() {                                                  // :372-392 (+ :862-906)
    std::vector<unsigned> rows;
    for (unsigned i = 0; i < predicted_pixel_positions_.size(); ++i) {
      double best = INFINITY; unsigned bj = 0;
      for (unsigned j = 0; j < image_points_.size(); ++j) {
        const double dx = predicted_pixel_positions_(i)(0) - image_points_(j)(0), dy = predicted_pixel_positions_(i)(1) - image_points_(j)(1);
        const double d2 = dx * dx + dy * dy;
        if (d2 < best) { best = d2; bj = j + 1; }
      }
      if (std::sqrt(best) <= nearest_neighbour_pixel_tolerance_) { rows.push_back(i + 1); rows.push_back(bj); }
    }
    correspondences_.resize((int)rows.size() / 2, 2);
    for (size_t i = 0; i < rows.size() / 2; ++i) { correspondences_((int)i, 0) = rows[2 * i]; correspondences_((int)i, 1) = rows[2 * i + 1]; }
    if (image_points_.size() > (size_t)MPE_MAX_DET) compactToMatchedDetections();
  }
  // Capacity path, same as the device loop (csrc/k4_tracking.cu): the brute-force tables hold MPE_MAX_DET detections; with more,
  // the detections that are some LED's nearest neighbour are compacted to the front (ascending) and the rows renumbered.
  void compactToMatchedDetections() {
    std::vector<int> remap(image_points_.size(), 0);
    for (int r = 0; r < correspondences_.rows(); ++r) remap[correspondences_(r, 1) - 1] = 1;
    List2DPoints kept; kept.resize(image_points_.size());
    size_t m = 0;
    for (size_t j = 0; j < image_points_.size(); ++j) if (remap[j]) { kept(m) = image_points_(j); remap[j] = (int)++m; }
    List2DPoints out; out.resize(m);
    for (size_t j = 0; j < m; ++j) out(j) = kept(j);
    image_points_ = out;
    for (int r = 0; r < correspondences_.rows(); ++r) correspondences_(r, 1) = (unsigned)remap[correspondences_(r, 1) - 1];
    too_many_detections_ = true;
  }
  bool tooManyDetections() const { return too_many_detections_; }
  void predictPose(double time_to_predict) {                                    // :232-244 (general 4x4 inverse, log map, exp map)
    predicted_time_ = time_to_predict;
    double out[16];
    detail::check(ctx_, mpe_host_predict_pose(previous_pose_, current_pose_, previous_time_, current_time_, time_to_predict, out), "mpe_host_predict_pose");
    std::memcpy(predicted_pose_, out, sizeof(out));
  }
  static void exponentialMap(const double twist[6], double T[16]) { mpe_host_exponential_map(twist, T); }   // :962-994
  static void logarithmMap(const double T[16], double xi[6]) { mpe_host_logarithm_map(T, xi); }             // :996-1064

  const RectT& regionOfInterest() const { return region_of_interest_; }
  const double* predictedPoseRowMajor() const { return predicted_pose_; }

 private:
  static const unsigned min_num_leds_detected_ = 4;                             // pose_estimator.h:78
  static unsigned factorial(int N) { return (N == 1 || N == 0) ? 1u : factorial(N - 1) * (unsigned)N; }          // combinations.cpp:34-40
  static unsigned numCombinations(unsigned N, unsigned K) { return (N >= K) ? factorial((int)N) / (factorial((int)K) * factorial((int)(N - K))) : 0u; }
  static void identity(double* T) { for (int i = 0; i < 16; ++i) T[i] = (i % 5 == 0) ? 1.0 : 0.0; }
  static Matrix4dT toEigen(const double* T) { Matrix4dT m; for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) m(r, c) = T[4 * r + c]; return m; }
  static void fromEigen(const Matrix4dT& m, double* T) { for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) T[4 * r + c] = m(r, c); }
  static void flatten(const List2DPoints& l, std::vector<double>& out) { out.resize(2 * l.size()); for (unsigned i = 0; i < l.size(); ++i) { out[2 * i] = l(i)(0); out[2 * i + 1] = l(i)(1); } }
  void flattenCorr(std::vector<uint32_t>& out) const { out.resize(2 * (size_t)correspondences_.rows()); for (int i = 0; i < correspondences_.rows(); ++i) { out[2 * i] = correspondences_(i, 0); out[2 * i + 1] = correspondences_(i, 1); } }

  void ensureContext(int w, int h) {
    if (ctx_ && w <= ctx_w_ && h <= ctx_h_) return;
    if (ctx_) mpe_destroy(ctx_);
    ctx_ = nullptr; ctx_w_ = w; ctx_h_ = h;
    if (mpe_create(&ctx_, 0, 1, w, h) != MPE_OK) { ctx_ = nullptr; throw std::runtime_error("mpe_create failed: no CUDA device? (there is no CPU fallback)"); }
    markers_dirty_ = true; pushed_valid_ = false;
  }
  void push() {   // the caller writes the public fields directly, so configuration is pushed before every device stage
    if (!ctx_) ensureContext(752, 480);
    mpe_params p; std::memset(&p, 0, sizeof(p));
    p.threshold_value = detection_threshold_value_; p.roi_border_thickness = (int)roi_border_thickness_; p.gaussian_sigma = gaussian_sigma_;
    p.min_blob_area = min_blob_area_; p.max_blob_area = max_blob_area_; p.max_width_height_distortion = max_width_height_distortion_;
    p.max_circular_distortion = max_circular_distortion_; p.back_projection_pixel_tolerance = back_projection_pixel_tolerance_;
    p.nearest_neighbour_pixel_tolerance = nearest_neighbour_pixel_tolerance_; p.certainty_threshold = certainty_threshold_;
    p.valid_correspondence_threshold = valid_correspondence_threshold_;
    // only what changed is pushed: every configuration call invalidates the captured CUDA graph of the device loop
    if (!pushed_valid_ || std::memcmp(&p, &pushed_params_, sizeof(p)) != 0) {
      detail::check(ctx_, mpe_set_params(ctx_, &p), "mpe_set_params");
      pushed_params_ = p;
    }
    double K[9]; detail::camera_to_rowmajor(camera_matrix_K_, K);
    if (!pushed_valid_ || std::memcmp(K, pushed_K_, sizeof(K)) != 0 || camera_distortion_coeffs_ != pushed_D_) {
      detail::check(ctx_, mpe_set_camera(ctx_, K, camera_distortion_coeffs_.data(), (int)camera_distortion_coeffs_.size()), "mpe_set_camera");
      std::memcpy(pushed_K_, K, sizeof(K)); pushed_D_ = camera_distortion_coeffs_;
    }
    if (markers_dirty_) {
      std::vector<double> xyz(3 * object_points_.size());
      for (unsigned i = 0; i < object_points_.size(); ++i) for (int k = 0; k < 3; ++k) xyz[3 * i + k] = object_points_(i)(k);
      detail::check(ctx_, mpe_set_markers(ctx_, xyz.data(), (int)object_points_.size()), "mpe_set_markers");
      markers_dirty_ = false;
    }
    if (!pushed_valid_ || histogram_threshold_ != pushed_hist_) {
      detail::check(ctx_, mpe_set_histogram_threshold(ctx_, histogram_threshold_), "mpe_set_histogram_threshold");
      pushed_hist_ = histogram_threshold_;
    }
    pushed_valid_ = true;
  }

  mpe_ctx* ctx_; int ctx_w_, ctx_h_; bool markers_dirty_ = true;
  bool device_loop_ = true, pushed_valid_ = false, too_many_detections_ = false;
  mpe_params pushed_params_; double pushed_K_[9]; std::vector<double> pushed_D_; unsigned pushed_hist_ = 0;
  double current_pose_[16], previous_pose_[16], predicted_pose_[16], cov_[36];   // row-major
  double current_time_, previous_time_, predicted_time_;
  List4DPoints object_points_;
  List2DPoints image_points_, predicted_pixel_positions_;
  VectorXuPairs correspondences_;
  double back_projection_pixel_tolerance_, nearest_neighbour_pixel_tolerance_, certainty_threshold_, valid_correspondence_threshold_;
  unsigned histogram_threshold_, it_since_initialized_, last_gn_iters_ = 0;
  std::vector<Point2fT> distorted_detection_centers_;
  RectT region_of_interest_;
  bool pose_updated_;
};

}  // namespace monocular_pose_estimator
#endif  // MPE_B200_SHIM_H_
