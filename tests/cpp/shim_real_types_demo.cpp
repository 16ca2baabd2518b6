// The C++ class shim compiled in its REAL-TYPES mode (MPE_SHIM_REAL_TYPES: Eigen::Matrix..., cv::Mat, cv::Rect, ...), i.e. with the
// very type names monocular_pose_estimator_lib's callers use (datatypes.h:38-52, pose_estimator.h:82-91), and driven exactly
// like MPENode drives the reference (monocular_pose_estimator/src/monocular_pose_estimator.cpp:103-126 cameraInfoCallback,
// :133-214 imageCallback incl. augmentImage, :220-236 dynamicParametersCallback).  Real Eigen/OpenCV are not installed in this
// image: the test build points the include path at the stand-ins under oracle/eigen_shim and oracle/cv_shim (test
// infrastructure), which is also what the unmodified reference sources are compiled against in oracle/_ref.
// Same input / output format as shim_demo.cpp.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>
#include "monocular_pose_estimator_b200/shim.h"

#ifndef MPE_SHIM_REAL_TYPES
#error "this translation unit must see <Eigen/Dense> and <opencv2/...> (stand-ins: -Ioracle/eigen_shim -Ioracle/cv_shim)"
#endif

using namespace monocular_pose_estimator;

// every typedef of datatypes.h must exist
static_assert(sizeof(Matrix6d) > 0 && sizeof(Matrix2x6d) > 0 && sizeof(Matrix3x4d) > 0 && sizeof(MatrixXYd) > 0 && sizeof(MatrixXYu) > 0 &&
              sizeof(Vector6d) > 0 && sizeof(Vector3u) > 0 && sizeof(Vector4u) > 0 && sizeof(VectorXu) > 0 && sizeof(VectorXuPairs) > 0 &&
              sizeof(RowXd) > 0 && sizeof(RowXu) > 0 && sizeof(List2DPoints) > 0 && sizeof(List3DPoints) > 0 && sizeof(List4DPoints) > 0, "datatypes.h");

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: shim_real_types_demo scene.bin [stage]\n"); return 2; }
  const bool stage_mode = argc > 2 && std::string(argv[2]) == "stage";
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror("open"); return 2; }
  int hdr[4];
  if (fread(hdr, sizeof(int), 4, f) != 4) return 2;
  const int n_frames = hdr[0], w = hdr[1], h = hdr[2], n_leds = hdr[3];
  double K[9], D[5], params[11];
  std::vector<double> markers(3 * n_leds), times(n_frames);
  if (fread(K, 8, 9, f) != 9 || fread(D, 8, 5, f) != 5 || fread(markers.data(), 8, markers.size(), f) != markers.size() ||
      fread(params, 8, 11, f) != 11 || fread(times.data(), 8, times.size(), f) != times.size()) return 2;
  std::vector<uint8_t> frames((size_t)n_frames * w * h);
  if (fread(frames.data(), 1, frames.size(), f) != frames.size()) return 2;
  fclose(f);

  try {
    PoseEstimator trackable_object_;
    trackable_object_.setDeviceLoop(!stage_mode);
    // cameraInfoCallback (:110-120)
    trackable_object_.camera_matrix_K_ = cv::Mat(3, 3, CV_64F);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) trackable_object_.camera_matrix_K_.at<double>(i, j) = K[3 * i + j];
    trackable_object_.camera_distortion_coeffs_.assign(D, D + 5);
    // dynamicParametersCallback (:222-233)
    trackable_object_.detection_threshold_value_ = (int)params[0];
    trackable_object_.gaussian_sigma_ = params[1];
    trackable_object_.min_blob_area_ = params[2];
    trackable_object_.max_blob_area_ = params[3];
    trackable_object_.max_width_height_distortion_ = params[4];
    trackable_object_.max_circular_distortion_ = params[5];
    trackable_object_.roi_border_thickness_ = (unsigned)params[10];
    trackable_object_.setBackProjectionPixelTolerance(params[6]);
    trackable_object_.setNearestNeighbourPixelTolerance(params[7]);
    trackable_object_.setCertaintyThreshold(params[8]);
    trackable_object_.setValidCorrespondenceThreshold(params[9]);
    // constructor (:63-84)
    List4DPoints positions_of_markers_on_object;
    positions_of_markers_on_object.resize(n_leds);
    for (int i = 0; i < n_leds; ++i) {
      Eigen::Matrix<double, 4, 1> temp_point;
      temp_point(0) = markers[3 * i]; temp_point(1) = markers[3 * i + 1]; temp_point(2) = markers[3 * i + 2]; temp_point(3) = 1;
      positions_of_markers_on_object(i) = temp_point;
    }
    trackable_object_.setMarkerPositions(positions_of_markers_on_object);
    for (int fi = 0; fi < n_frames; ++fi) {
      cv::Mat image(h, w, CV_8UC1, frames.data() + (size_t)fi * w * h, (size_t)w);      // cv_bridge::toCvCopy(..., MONO8)->image
      const bool found_body_pose = trackable_object_.estimateBodyPose(image, times[fi]);   // :159
      Eigen::Matrix4d transform = trackable_object_.getPredictedPose();                    // :163
      Matrix6d cov = trackable_object_.getPoseCovariance();                                // :164
      (void)cov;
      cv::Mat visualized_image = image.clone();
      trackable_object_.augmentImage(visualized_image);                                    // :204
      const cv::Rect& r = trackable_object_.regionOfInterest();
      printf("%d %d %d %d %d %d %u", fi, found_body_pose ? 1 : 0, r.x, r.y, r.width, r.height, trackable_object_.lastGaussNewtonIterations());
      for (int a = 0; a < 4; ++a) for (int b = 0; b < 4; ++b) printf(" %.17g", transform(a, b));
      printf("\n");
    }
  } catch (const std::exception& e) {
    fprintf(stderr, "shim_real_types_demo: %s\n", e.what());
    return 1;
  }
  return 0;
}
