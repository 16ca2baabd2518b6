// TEST INFRASTRUCTURE — C entry points around the UNMODIFIED reference pose path.  oracle/Makefile compiles
// /root/reference/monocular_pose_estimator_lib/src/{p3p,combinations,pose_estimator,led_detector}.cpp where they lie, against
// the stand-ins oracle/eigen_shim (Eigen) and oracle/cv_shim (OpenCV types; imgproc calls forwarded to cv2 callbacks), plus
// this file, into oracle/_ref/libref_pose.so.  No reference source is copied or edited: private members are reached by
// compiling THIS translation unit with `private` spelled `public` around the reference headers (layout is unchanged).
//
// The calling convention mirrors the mpeo_* functions of oracle/pose_oracle.cpp (row-major 4x4 poses, n x 2 point lists,
// 1-based (LED, detection) correspondence rows) so that tests can drive both with the same code.
// Only tests/ (and the fixture generators under tests/golden/) load this library.
#include <cmath>
#include <complex>
#include <cstring>
#include <iostream>
#include <memory>
#include <vector>
#include <algorithm>
#include <Eigen/Dense>
#include <Eigen/Geometry>
#include <opencv2/opencv.hpp>

#define private public
#include "monocular_pose_estimator_lib/pose_estimator.h"
#undef private

using namespace monocular_pose_estimator;

// ---- OpenCV entry points forwarded to the registered callbacks -----------------------------------------------------
namespace cv_shim {
cv_shim_callbacks& callbacks() { static cv_shim_callbacks c = {}; return c; }
}
namespace cv {
static std::vector<int> flat(const std::vector<Point>& c) {
  std::vector<int> f(2 * c.size());
  for (size_t i = 0; i < c.size(); ++i) { f[2 * i] = c[i].x; f[2 * i + 1] = c[i].y; }
  return f;
}
double threshold(const Mat& src, Mat& dst, double thresh, double maxval, int type) {
  dst.create(src.rows, src.cols, CV_8UC1);
  if (src.rows > 0 && src.cols > 0) cv_shim::callbacks().threshold(src.data, src.rows, src.cols, (long)src.step, thresh, maxval, type, dst.data);
  return thresh;
}
void GaussianBlur(const Mat& src, Mat& dst, Size ksize, double sigmaX, double sigmaY, int borderType) {
  (void)ksize;   // the reference passes (0,0): derived from sigma inside OpenCV
  dst.create(src.rows, src.cols, CV_8UC1);
  if (src.rows > 0 && src.cols > 0) cv_shim::callbacks().gaussian_blur(src.data, src.rows, src.cols, (long)src.step, sigmaX, sigmaY, borderType, dst.data);
}
void findContours(const Mat& image, std::vector<std::vector<Point> >& contours, int mode, int method) {
  contours.clear();
  if (image.rows == 0 || image.cols == 0) return;
  const int* counts = nullptr; const int* pts = nullptr;
  int n = cv_shim::callbacks().find_contours(image.data, image.rows, image.cols, (long)image.step, mode, method, &counts, &pts);
  size_t k = 0;
  for (int i = 0; i < n; ++i) {
    std::vector<Point> c((size_t)counts[i]);
    for (int j = 0; j < counts[i]; ++j, ++k) c[(size_t)j] = Point(pts[2 * k], pts[2 * k + 1]);
    contours.push_back(c);
  }
}
double contourArea(const std::vector<Point>& contour, bool) { std::vector<int> f = flat(contour); return cv_shim::callbacks().contour_area(f.data(), (int)contour.size()); }
Rect boundingRect(const std::vector<Point>& contour) {
  std::vector<int> f = flat(contour); int r[4];
  cv_shim::callbacks().bounding_rect(f.data(), (int)contour.size(), r);
  return Rect(r[0], r[1], r[2], r[3]);
}
Moments moments(const std::vector<Point>& contour, bool) {
  std::vector<int> f = flat(contour); double m[10];
  cv_shim::callbacks().moments(f.data(), (int)contour.size(), m);
  Moments mu; mu.m00 = m[0]; mu.m10 = m[1]; mu.m01 = m[2]; mu.m20 = m[3]; mu.m11 = m[4]; mu.m02 = m[5]; mu.m30 = m[6]; mu.m21 = m[7]; mu.m12 = m[8]; mu.m03 = m[9];
  return mu;
}
void undistortPoints(const std::vector<Point2f>& src, std::vector<Point2f>& dst, const Mat& K, const std::vector<double>& D, NoArray, const Mat& P) {
  double k[9], p[9];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { k[3 * i + j] = K.at<double>(i, j); p[3 * i + j] = P.at<double>(i, j); }
  std::vector<float> in(2 * src.size()), out(2 * src.size());
  for (size_t i = 0; i < src.size(); ++i) { in[2 * i] = src[i].x; in[2 * i + 1] = src[i].y; }
  cv_shim::callbacks().undistort_points(in.data(), (int)src.size(), k, D.data(), (int)D.size(), p, out.data());
  dst.resize(src.size());
  for (size_t i = 0; i < src.size(); ++i) dst[i] = Point2f(out[2 * i], out[2 * i + 1]);
}
void projectPoints(const std::vector<Point3f>& pts, const Mat& rvec, const Mat& tvec, const Mat& K, const std::vector<double>& D,
                   std::vector<Point2f>& out) {
  double k[9], r[3], t[3];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) k[3 * i + j] = K.at<double>(i, j);
  for (int i = 0; i < 3; ++i) { r[i] = rvec.at<double>(i, 0); t[i] = tvec.at<double>(i, 0); }
  std::vector<float> in(3 * pts.size()), o(2 * pts.size());
  for (size_t i = 0; i < pts.size(); ++i) { in[3 * i] = pts[i].x; in[3 * i + 1] = pts[i].y; in[3 * i + 2] = pts[i].z; }
  cv_shim::callbacks().project_points(in.data(), (int)pts.size(), r, t, k, D.data(), (int)D.size(), o.data());
  out.resize(pts.size());
  for (size_t i = 0; i < pts.size(); ++i) out[i] = Point2f(o[2 * i], o[2 * i + 1]);
}
static void draw(Mat& img, int what, int a, int b, int c, int d, const Scalar& color, int thickness) {
  const int g[4] = {a, b, c, d};
  cv_shim::callbacks().draw(img.data, img.rows, img.cols, (long)img.step, img.channels(), what, g, color.val, thickness);
}
void line(Mat& img, Point p1, Point p2, const Scalar& color, int thickness) { draw(img, 0, p1.x, p1.y, p2.x, p2.y, color, thickness); }
void circle(Mat& img, Point c, int radius, const Scalar& color, int thickness) { draw(img, 1, c.x, c.y, radius, 0, color, thickness); }
void rectangle(Mat& img, Rect r, const Scalar& color, int thickness) { draw(img, 2, r.x, r.y, r.width, r.height, color, thickness); }
}  // namespace cv

namespace {

struct Ref {
  PoseEstimator pe;
  std::vector<unsigned> last_hist;   // n_det x n_obj row-major, hist_corr as it was handed to correspondencesFromHistogram
  unsigned last_gn_iterations = 0;
  Ref() {
    pe.camera_matrix_K_ = cv::Mat(3, 3, CV_64F);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) pe.camera_matrix_K_.at<double>(i, j) = (i == j) ? 1.0 : 0.0;
    // members the reference constructor leaves unset (pose_estimator.cpp:34-42); the ROS node sets them before any frame
    pe.current_pose_.setIdentity(); pe.previous_pose_.setIdentity(); pe.predicted_pose_.setIdentity(); pe.pose_covariance_.setZero();
    pe.current_time_ = pe.previous_time_ = pe.predicted_time_ = 0; pe.histogram_threshold_ = 0; pe.pose_updated_ = false;
    pe.detection_threshold_value_ = 140; pe.gaussian_sigma_ = 0.6; pe.min_blob_area_ = 10; pe.max_blob_area_ = 200;
    pe.max_width_height_distortion_ = 0.5; pe.max_circular_distortion_ = 0.5; pe.roi_border_thickness_ = 20;
  }
};

Eigen::Matrix4d m4_from_rowmajor(const double in[16]) { Eigen::Matrix4d m; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) m(i, j) = in[4 * i + j]; return m; }
void m4_to_rowmajor(const Eigen::Matrix4d& m, double out[16]) { for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) out[4 * i + j] = m(i, j); }
List2DPoints list2(const double* p, int n) { List2DPoints l(n); for (int i = 0; i < n; ++i) { l(i)(0) = p[2 * i]; l(i)(1) = p[2 * i + 1]; } return l; }

}  // namespace

extern "C" {

void mper_set_cv_callbacks(const cv_shim_callbacks* c) { cv_shim::callbacks() = *c; }

// Visualization::createVisualizationImage (visualization.cpp:57-104) on a 3-channel 8-bit image, in place
void mper_create_visualization_image(unsigned char* img, int rows, int cols, long step, const double pose[16], const double K[9],
                                     const double* D, int nD, const int roi_xywh[4], const float* centers, int n_centers) {
  cv::Mat image(rows, cols, CV_8UC3, img, (size_t)step);
  cv::Mat Km(3, 3, CV_64F);
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Km.at<double>(i, j) = K[3 * i + j];
  std::vector<double> Dv(D, D + nD);
  std::vector<cv::Point2f> c((size_t)n_centers);
  for (int i = 0; i < n_centers; ++i) c[(size_t)i] = cv::Point2f(centers[2 * i], centers[2 * i + 1]);
  Visualization::createVisualizationImage(image, m4_from_rowmajor(pose), Km, Dv, cv::Rect(roi_xywh[0], roi_xywh[1], roi_xywh[2], roi_xywh[3]), c);
}

int mper_p3p(const double f[9], const double P[9], double sol[48]) {
  Eigen::Matrix3d fv, wp;
  for (int k = 0; k < 3; ++k) for (int r = 0; r < 3; ++r) { fv(r, k) = f[3 * k + r]; wp(r, k) = P[3 * k + r]; }
  Eigen::Matrix<Eigen::Matrix<double, 3, 4>, 4, 1> s;
  for (int i = 0; i < 48; ++i) sol[i] = 0;
  int rc = P3P::computePoses(fv, wp, s);
  if (rc == 0) for (int i = 0; i < 4; ++i) for (int r = 0; r < 3; ++r) for (int c = 0; c < 4; ++c) sol[12 * i + 4 * r + c] = s(i)(r, c);
  return rc;
}
int mper_solve_quartic(const double factors[5], double roots[4]) {
  Eigen::Matrix<double, 5, 1> f; Eigen::Matrix<double, 4, 1> r;
  for (int i = 0; i < 5; ++i) f(i) = factors[i];
  int rc = P3P::solveQuartic(f, r);
  for (int i = 0; i < 4; ++i) roots[i] = r(i);
  return rc;
}

// Combinations (combinations.cpp:42-244): tables row-major, 1-based as the reference returns them
static int copy_table(const MatrixXYu& t, unsigned* out, int cap_rows, int* n_cols) {
  *n_cols = (int)t.cols();
  if ((int)t.rows() > cap_rows) return -(int)t.rows();
  for (int i = 0; i < (int)t.rows(); ++i) for (int j = 0; j < (int)t.cols(); ++j) out[i * t.cols() + j] = t(i, j);
  return (int)t.rows();
}
int mper_combinations_no_replacement(unsigned N, unsigned K, unsigned* out, int cap_rows, int* n_cols) { return copy_table(Combinations::combinationsNoReplacement(N, K), out, cap_rows, n_cols); }
int mper_permutations_no_replacement(unsigned N, unsigned K, unsigned* out, int cap_rows, int* n_cols) { return copy_table(Combinations::permutationsNoReplacement(N, K), out, cap_rows, n_cols); }
unsigned mper_num_combinations(unsigned N, unsigned K) { return Combinations::numCombinations(N, K); }
unsigned mper_num_permutations(unsigned N, unsigned K) { return Combinations::numPermutations(N, K); }

void* mper_create() { return new Ref(); }
void mper_destroy(void* h) { delete (Ref*)h; }

void mper_set_camera(void* h, const double K[9], const double* D, int nD) {
  PoseEstimator& pe = ((Ref*)h)->pe;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) pe.camera_matrix_K_.at<double>(i, j) = K[3 * i + j];
  pe.camera_distortion_coeffs_.assign(D, D + nD);
}
void mper_set_markers(void* h, const double* xyz, int n) {
  List4DPoints m(n);
  for (int i = 0; i < n; ++i) { m(i)(0) = xyz[3 * i]; m(i)(1) = xyz[3 * i + 1]; m(i)(2) = xyz[3 * i + 2]; m(i)(3) = 1.0; }
  ((Ref*)h)->pe.setMarkerPositions(m);
}
void mper_set_params(void* h, double back_proj_tol, double nn_tol, double certainty_thr, double valid_corr_thr) {
  PoseEstimator& pe = ((Ref*)h)->pe;
  pe.setBackProjectionPixelTolerance(back_proj_tol); pe.setNearestNeighbourPixelTolerance(nn_tol);
  pe.setCertaintyThreshold(certainty_thr); pe.setValidCorrespondenceThreshold(valid_corr_thr);
}
void mper_set_detector_params(void* h, int threshold, double sigma, double min_area, double max_area, double max_wh, double max_circ, unsigned roi_border) {
  PoseEstimator& pe = ((Ref*)h)->pe;
  pe.detection_threshold_value_ = threshold; pe.gaussian_sigma_ = sigma; pe.min_blob_area_ = min_area; pe.max_blob_area_ = max_area;
  pe.max_width_height_distortion_ = max_wh; pe.max_circular_distortion_ = max_circ; pe.roi_border_thickness_ = roi_border;
}
void mper_set_histogram_threshold(void* h, unsigned t) { ((Ref*)h)->pe.setHistogramThreshold(t); }
unsigned mper_get_histogram_threshold(void* h) { return ((Ref*)h)->pe.getHistogramThreshold(); }
void mper_set_image_points(void* h, const double* pts, int n) { ((Ref*)h)->pe.setImagePoints(list2(pts, n)); }
int mper_get_image_points(void* h, double* out) {
  PoseEstimator& pe = ((Ref*)h)->pe;
  for (int i = 0; i < (int)pe.image_points_.size(); ++i) { out[2 * i] = pe.image_points_(i)(0); out[2 * i + 1] = pe.image_points_(i)(1); }
  return (int)pe.image_points_.size();
}
int mper_get_image_vectors(void* h, double* out) {
  PoseEstimator& pe = ((Ref*)h)->pe;
  for (int i = 0; i < (int)pe.image_vectors_.size(); ++i) for (int k = 0; k < 3; ++k) out[3 * i + k] = pe.image_vectors_(i)(k);
  return (int)pe.image_vectors_.size();
}

static void arm_hist(Ref* r) {
  Eigen::shim::Hooks& hk = Eigen::shim::hooks();
  hk.armed = true; hk.first_unsigned_max_coeff.clear(); hk.first_rows = hk.first_cols = 0;
  (void)r;
}
static void harvest_hist(Ref* r) {
  Eigen::shim::Hooks& hk = Eigen::shim::hooks();
  const int nd = (int)r->pe.image_points_.size(), no = (int)r->pe.object_points_.size();
  r->last_hist.assign((size_t)nd * no, 0u);            // never decoded => the histogram was all zero (pose_estimator.cpp:704)
  if (!hk.armed && hk.first_rows == nd && hk.first_cols == no)
    for (int i = 0; i < nd; ++i) for (int j = 0; j < no; ++j) r->last_hist[(size_t)i * no + j] = hk.first_unsigned_max_coeff[(size_t)j * nd + i];
  hk.armed = false;
}
unsigned mper_initialise(void* h) {
  Ref* r = (Ref*)h;
  arm_hist(r);
  unsigned ok = r->pe.initialise();
  harvest_hist(r);
  return ok;
}
int mper_get_histogram(void* h, unsigned* out) {
  Ref* r = (Ref*)h;
  std::copy(r->last_hist.begin(), r->last_hist.end(), out);
  return (int)r->last_hist.size();
}
int mper_get_correspondences(void* h, unsigned* out) {
  PoseEstimator& pe = ((Ref*)h)->pe;
  for (int i = 0; i < (int)pe.correspondences_.rows(); ++i) { out[2 * i] = pe.correspondences_(i, 0); out[2 * i + 1] = pe.correspondences_(i, 1); }
  return (int)pe.correspondences_.rows();
}
void mper_set_correspondences(void* h, const unsigned* c, int n) {
  VectorXuPairs v(n, 2);
  for (int i = 0; i < n; ++i) { v(i, 0) = c[2 * i]; v(i, 1) = c[2 * i + 1]; }
  ((Ref*)h)->pe.setCorrespondences(v);
}
// correspondencesFromHistogram (pose_estimator.cpp:344-370) on a caller-supplied histogram (row-major n_det x n_obj)
int mper_correspondences_from_histogram(void* h, const unsigned* hist, int n_det, int n_obj, unsigned* out) {
  MatrixXYu m(n_det, n_obj);
  for (int i = 0; i < n_det; ++i) for (int j = 0; j < n_obj; ++j) m(i, j) = hist[i * n_obj + j];
  VectorXuPairs c = ((Ref*)h)->pe.correspondencesFromHistogram(m);
  for (int i = 0; i < (int)c.rows(); ++i) { out[2 * i] = c(i, 0); out[2 * i + 1] = c(i, 1); }
  return (int)c.rows();
}
// calculateMinDistancesAndPairs (pose_estimator.cpp:862-906)
void mper_min_distances_and_pairs(void* h, const double* a, int na, const double* b, int nb, unsigned* pairs_out, double* dist_out) {
  Eigen::VectorXd d;
  VectorXuPairs p = ((Ref*)h)->pe.calculateMinDistancesAndPairs(list2(a, na), list2(b, nb), d);
  for (int i = 0; i < na; ++i) { pairs_out[2 * i] = p(i, 0); pairs_out[2 * i + 1] = p(i, 1); dist_out[i] = d(i); }
}
unsigned mper_check_correspondences(void* h) { return ((Ref*)h)->pe.checkCorrespondences(); }
void mper_find_correspondences(void* h) { ((Ref*)h)->pe.findCorrespondences(); }
int mper_optimise_pose(void* h) {
  Ref* r = (Ref*)h;
  unsigned long before = Eigen::shim::hooks().ldlt_calls;
  r->pe.optimisePose();
  r->last_gn_iterations = (unsigned)(Eigen::shim::hooks().ldlt_calls - before);
  return (int)r->last_gn_iterations;
}
void mper_optimise_and_update_pose(void* h) {
  Ref* r = (Ref*)h;
  unsigned long before = Eigen::shim::hooks().ldlt_calls;
  double t = r->pe.predicted_time_;
  r->pe.optimiseAndUpdatePose(t);
  r->last_gn_iterations = (unsigned)(Eigen::shim::hooks().ldlt_calls - before);
}
int mper_last_gn_iterations(void* h) { return (int)((Ref*)h)->last_gn_iterations; }
void mper_update_pose(void* h) { ((Ref*)h)->pe.updatePose(); }
void mper_get_predicted_pose(void* h, double out[16]) { m4_to_rowmajor(((Ref*)h)->pe.getPredictedPose(), out); }
void mper_set_predicted_pose(void* h, const double in[16], double time) { ((Ref*)h)->pe.setPredictedPose(m4_from_rowmajor(in), time); }
void mper_get_current_pose(void* h, double out[16]) { m4_to_rowmajor(((Ref*)h)->pe.current_pose_, out); }
void mper_get_previous_pose(void* h, double out[16]) { m4_to_rowmajor(((Ref*)h)->pe.previous_pose_, out); }
void mper_set_state(void* h, const double cur[16], const double prev[16], double cur_t, double prev_t, unsigned it_since_init) {
  PoseEstimator& pe = ((Ref*)h)->pe;
  pe.current_pose_ = m4_from_rowmajor(cur); pe.previous_pose_ = m4_from_rowmajor(prev);
  pe.current_time_ = cur_t; pe.previous_time_ = prev_t; pe.it_since_initialized_ = it_since_init;
}
void mper_get_covariance(void* h, double out[36]) { Matrix6d c = ((Ref*)h)->pe.getPoseCovariance(); for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) out[6 * i + j] = c(i, j); }
void mper_set_predicted_time(void* h, double t) { ((Ref*)h)->pe.setPredictedTime(t); }
double mper_get_predicted_time(void* h) { return ((Ref*)h)->pe.getPredictedTime(); }
unsigned mper_it_since_initialized(void* h) { return ((Ref*)h)->pe.it_since_initialized_; }
void mper_predict_pose(void* h, double t) { ((Ref*)h)->pe.predictPose(t); }
void mper_predict_marker_positions(void* h) { ((Ref*)h)->pe.predictMarkerPositionsInImage(); }
int mper_get_predicted_pixels(void* h, double* out) {
  List2DPoints p = ((Ref*)h)->pe.getPredictedPixelPositions();
  for (int i = 0; i < (int)p.size(); ++i) { out[2 * i] = p(i)(0); out[2 * i + 1] = p(i)(1); }
  return (int)p.size();
}
void mper_set_predicted_pixels(void* h, const double* p, int n) { ((Ref*)h)->pe.setPredictedPixels(list2(p, n)); }
void mper_determine_roi(void* h, int img_w, int img_h, int border, int roi[4]) {
  PoseEstimator& pe = ((Ref*)h)->pe;
  cv::Rect r = LEDDetector::determineROI(pe.getPredictedPixelPositions(), cv::Size(img_w, img_h), border, pe.camera_matrix_K_, pe.camera_distortion_coeffs_);
  roi[0] = r.x; roi[1] = r.y; roi[2] = r.width; roi[3] = r.height;
}
// project2d (pose_estimator.cpp:251-268) is `inline` in the reference's .cpp and therefore not linkable from here; it is
// reached through predictMarkerPositionsInImage (:270-276) on a copy of the estimator that holds the one point.
void mper_project2d(void* h, const double point[4], const double T[16], double out[2]) {
  PoseEstimator tmp = ((Ref*)h)->pe;
  List4DPoints one(1); for (int i = 0; i < 4; ++i) one(0)(i) = point[i];
  tmp.object_points_ = one; tmp.predicted_pixel_positions_.resize(1);   // (setMarkerPositions would evaluate C(1,3))
  tmp.predicted_pose_ = m4_from_rowmajor(T);
  tmp.predictMarkerPositionsInImage();
  out[0] = tmp.predicted_pixel_positions_(0)(0); out[1] = tmp.predicted_pixel_positions_(0)(1);
}
void mper_exponential_map(const double twist[6], double out[16]) {
  PoseEstimator pe; Vector6d t; for (int i = 0; i < 6; ++i) t(i) = twist[i];
  m4_to_rowmajor(pe.exponentialMap(t), out);
}
void mper_logarithm_map(const double T[16], double xi[6]) {
  PoseEstimator pe; Vector6d x = pe.logarithmMap(m4_from_rowmajor(T));
  for (int i = 0; i < 6; ++i) xi[i] = x(i);
}
// computeJacobian (pose_estimator.cpp:932-960); out row-major 2x6
void mper_compute_jacobian(const double T[16], const double point[4], const double focal[2], double out[12]) {
  PoseEstimator pe; Eigen::Vector4d p; for (int i = 0; i < 4; ++i) p(i) = point[i];
  Eigen::Vector2d f; f(0) = focal[0]; f(1) = focal[1];
  Matrix2x6d J = pe.computeJacobian(m4_from_rowmajor(T), p, f);
  for (int i = 0; i < 2; ++i) for (int j = 0; j < 6; ++j) out[6 * i + j] = J(i, j);
}
// computeTransformation / Kabsch (pose_estimator.cpp:908-930); points n x 3 row-major
void mper_compute_transformation(const double* object_pts, const double* reprojected_pts, int n, double out[16]) {
  PoseEstimator pe; MatrixXYd a(3, n), b(3, n);
  for (int j = 0; j < n; ++j) for (int i = 0; i < 3; ++i) { a(i, j) = object_pts[3 * j + i]; b(i, j) = reprojected_pts[3 * j + i]; }
  m4_to_rowmajor(pe.computeTransformation(a, b), out);
}
void mper_distort_point(void* h, float x, float y, float out[2]) {
  PoseEstimator& pe = ((Ref*)h)->pe;
  std::vector<cv::Point2f> src(1, cv::Point2f(x, y)), dst;
  LEDDetector::distortPoints(src, dst, pe.camera_matrix_K_, pe.camera_distortion_coeffs_);
  out[0] = dst[0].x; out[1] = dst[0].y;
}

// LEDDetector::findLeds (led_detector.cpp:35-112).  *n_px_inout: in = current length of px_out (the reference leaves
// pixel_positions untouched when nothing is found, :91), out = its length afterwards.  Returns the number of distorted centres.
int mper_find_leds(const unsigned char* img, int rows, int cols, long step, const int roi[4], int threshold, double sigma,
                   double min_area, double max_area, double max_wh, double max_circ, const double K[9], const double* D, int nD,
                   double* px_out, int* n_px_inout, float* centers_out) {
  cv::Mat image(rows, cols, CV_8UC1, (void*)img, (size_t)step), Km(3, 3, CV_64F);
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Km.at<double>(i, j) = K[3 * i + j];
  std::vector<double> Dv(D, D + nD);
  List2DPoints px(*n_px_inout);
  for (int i = 0; i < *n_px_inout; ++i) { px(i)(0) = px_out[2 * i]; px(i)(1) = px_out[2 * i + 1]; }
  std::vector<cv::Point2f> centers;
  LEDDetector::findLeds(image, cv::Rect(roi[0], roi[1], roi[2], roi[3]), threshold, sigma, min_area, max_area, max_wh, max_circ, px, centers, Km, Dv);
  for (int i = 0; i < (int)px.size(); ++i) { px_out[2 * i] = px(i)(0); px_out[2 * i + 1] = px(i)(1); }
  *n_px_inout = (int)px.size();
  for (size_t i = 0; i < centers.size(); ++i) { centers_out[2 * i] = centers[i].x; centers_out[2 * i + 1] = centers[i].y; }
  return (int)centers.size();
}

// PoseEstimator::estimateBodyPose (pose_estimator.cpp:62-147), the whole per-frame state machine, unmodified.
int mper_estimate_body_pose(void* h, const unsigned char* img, int rows, int cols, long step, double time_to_predict) {
  Ref* r = (Ref*)h;
  cv::Mat image(rows, cols, CV_8UC1, (void*)img, (size_t)step);
  unsigned long before = Eigen::shim::hooks().ldlt_calls;
  arm_hist(r);
  bool ok = r->pe.estimateBodyPose(image, time_to_predict);
  harvest_hist(r);
  r->last_gn_iterations = (unsigned)(Eigen::shim::hooks().ldlt_calls - before);
  return ok ? 1 : 0;
}
void mper_get_roi(void* h, int roi[4]) { const cv::Rect& r = ((Ref*)h)->pe.region_of_interest_; roi[0] = r.x; roi[1] = r.y; roi[2] = r.width; roi[3] = r.height; }
int mper_get_distorted_centers(void* h, float* out) {
  const std::vector<cv::Point2f>& c = ((Ref*)h)->pe.distorted_detection_centers_;
  for (size_t i = 0; i < c.size(); ++i) { out[2 * i] = c[i].x; out[2 * i + 1] = c[i].y; }
  return (int)c.size();
}

}  // extern "C"
