// CPU test driver: the overlay code of the C++ shim (PoseEstimator::augmentImage -> detail::draw_overlay) behind a C entry point,
// compiled in real-types mode against the Eigen / OpenCV stand-ins of oracle/ with the drawing API switched on.  The stand-ins
// forward projectPoints / line / circle / rectangle to cv2 (callbacks registered by oracle/ref_pose.py), exactly as they do for
// the reference's own visualization.cpp, so the two images can be compared byte for byte (tests/test_oracle_pose_ref.py).
#define MPE_SHIM_OPENCV_DRAWING 1
#include "monocular_pose_estimator_b200/shim.h"

extern "C" void shim_draw_overlay(unsigned char* img, int rows, int cols, long step, const double pose[16], const double K[9], const double* D,
                                  int nD, const int roi_xywh[4], const float* centers, int n_centers) {
  using namespace monocular_pose_estimator;
  cv::Mat image(rows, cols, CV_8UC3, img, (size_t)step);
  cv::Mat Km(3, 3, CV_64F);
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Km.at<double>(i, j) = K[3 * i + j];
  std::vector<double> Dv(D, D + nD);
  std::vector<cv::Point2f> c((size_t)n_centers);
  for (int i = 0; i < n_centers; ++i) c[(size_t)i] = cv::Point2f(centers[2 * i], centers[2 * i + 1]);
  detail::draw_overlay(image, pose, Km, Dv, cv::Rect(roi_xywh[0], roi_xywh[1], roi_xywh[2], roi_xywh[3]), c);
}
